#include <stdarg.h>
#include <atomic>
#include <vector>

#include "common.cuh"

namespace b200 {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n); }

struct ProfRec { cudaEvent_t a, b; double flops; int* sweeps_host; };
static bool g_prof = false;
static std::vector<ProfRec> g_pending;
static std::vector<ProfRec> g_free;
static cudaEvent_t g_cur_start;
static ProfRec g_cur;

bool profile_on() { return g_prof; }
void profile_begin(cudaStream_t s) {
  if (!g_prof) return;
  if (!g_free.empty()) { g_cur = g_free.back(); g_free.pop_back(); }
  else {
    cudaEventCreate(&g_cur.a);
    cudaEventCreate(&g_cur.b);
    cudaMallocHost((void**)&g_cur.sweeps_host, sizeof(int));
  }
  cudaEventRecord(g_cur.a, s);
}
void profile_end(cudaStream_t s, double flops, const int* sweeps_dev) {
  if (!g_prof) return;
  cudaEventRecord(g_cur.b, s);
  cudaMemcpyAsync(g_cur.sweeps_host, sweeps_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
  g_cur.flops = flops;
  g_pending.push_back(g_cur);
}
}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int b200_abi_version(void) { return 1; }
uint64_t b200_launch_count(void) { return b200::g_launches.load(); }
int b200_profile_enable(int on) { b200::g_prof = (on != 0); return B200_OK; }
int b200_profile_read(double* kernel_ms, double* algorithmic_flops, uint64_t* launches,
                      uint64_t* sweeps) {
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  double ms = 0.0, fl = 0.0;
  uint64_t n = 0, sw = 0;
  for (auto& r : b200::g_pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; fl += r.flops; ++n; sw += (uint64_t)*r.sweeps_host; }
    b200::g_free.push_back(r);
  }
  b200::g_pending.clear();
  if (kernel_ms) *kernel_ms = ms;
  if (algorithmic_flops) *algorithmic_flops = fl;
  if (launches) *launches = n;
  if (sweeps) *sweeps = sw;
  return B200_OK;
}
}
