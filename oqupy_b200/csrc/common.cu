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

struct ProfRec { cudaEvent_t a, b; double flops; int slot; };
static bool g_prof = false;
static std::vector<ProfRec> g_pending;
static std::vector<ProfRec> g_free;
static ProfRec g_cur;
// sweeps of every profiled launch land in ONE pinned ring (allocated once: a pinned
// allocation per launch would cost more than the kernels being timed)
static const int kProfSlots = 1 << 16;
static int* g_sweeps_host = nullptr;
static int g_next_slot = 0;

bool profile_on() { return g_prof; }
void profile_begin(cudaStream_t s) {
  if (!g_prof) return;
  if (!g_sweeps_host) cudaMallocHost((void**)&g_sweeps_host, kProfSlots * sizeof(int));
  if (!g_free.empty()) { g_cur = g_free.back(); g_free.pop_back(); }
  else {
    cudaEventCreate(&g_cur.a);
    cudaEventCreate(&g_cur.b);
  }
  cudaEventRecord(g_cur.a, s);
}
void profile_end(cudaStream_t s, double flops, const int* sweeps_dev) {
  if (!g_prof) return;
  cudaEventRecord(g_cur.b, s);
  g_cur.slot = g_next_slot;
  g_next_slot = (g_next_slot + 1) % kProfSlots;
  cudaMemcpyAsync(g_sweeps_host + g_cur.slot, sweeps_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
  g_cur.flops = flops;
  g_pending.push_back(g_cur);
}
}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int b200_abi_version(void) { return 1; }
uint64_t b200_launch_count(void) { return b200::g_launches.load(); }
int b200_profile_enable(int on) {
  b200::g_prof = (on != 0);
  if (on) {   // create the event pairs up front: not inside somebody's timed region
    if (!b200::g_sweeps_host)
      cudaMallocHost((void**)&b200::g_sweeps_host, b200::kProfSlots * sizeof(int));
    while (b200::g_free.size() < 12288) {
      b200::ProfRec r;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      r.flops = 0.0; r.slot = 0;
      b200::g_free.push_back(r);
    }
  }
  return B200_OK;
}
int b200_profile_read(double* kernel_ms, double* algorithmic_flops, uint64_t* launches,
                      uint64_t* sweeps) {
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  double ms = 0.0, fl = 0.0;
  uint64_t n = 0, sw = 0;
  for (auto& r : b200::g_pending) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; fl += r.flops; ++n; sw += (uint64_t)b200::g_sweeps_host[r.slot]; }
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
