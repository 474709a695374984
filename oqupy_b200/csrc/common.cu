#include <stdarg.h>
#include <atomic>
#include <mutex>
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

struct ProfRec { cudaEvent_t a, b; double flops; int slot; int kind; };
static std::mutex g_prof_mu;   // the library is driven from several host threads (ensembles)
static bool g_prof = false;
static std::vector<ProfRec> g_pending;
static std::vector<ProfRec> g_free;
static thread_local ProfRec g_cur;
// sweeps of every profiled launch land in ONE pinned ring (allocated once: a pinned
// allocation per launch would cost more than the kernels being timed)
static const int kProfSlots = 1 << 16;
static int* g_sweeps_host = nullptr;
static int g_next_slot = 0;

bool profile_on() { return g_prof; }
void profile_begin(cudaStream_t s, int kind) {
  if (!g_prof) return;
  {
    std::lock_guard<std::mutex> g(g_prof_mu);
    if (!g_sweeps_host) cudaMallocHost((void**)&g_sweeps_host, kProfSlots * sizeof(int));
    if (!g_free.empty()) { g_cur = g_free.back(); g_free.pop_back(); }
    else {
      cudaEventCreate(&g_cur.a);
      cudaEventCreate(&g_cur.b);
    }
  }
  g_cur.kind = kind;
  cudaEventRecord(g_cur.a, s);
}
void profile_end(cudaStream_t s, int kind, double flops, const int* sweeps_dev) {
  if (!g_prof) return;
  cudaEventRecord(g_cur.b, s);
  std::lock_guard<std::mutex> g(g_prof_mu);
  g_cur.slot = -1;
  if (sweeps_dev) {
    g_cur.slot = g_next_slot;
    g_next_slot = (g_next_slot + 1) % kProfSlots;
    cudaMemcpyAsync(g_sweeps_host + g_cur.slot, sweeps_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
  }
  g_cur.flops = flops;
  g_cur.kind = kind;
  g_pending.push_back(g_cur);
}
}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int b200_abi_version(void) { return 2; }
uint64_t b200_launch_count(void) { return b200::g_launches.load(); }
int b200_profile_enable(int on) {
  std::lock_guard<std::mutex> g(b200::g_prof_mu);
  b200::g_prof = (on != 0);
  if (on) {   // create the event pairs up front: not inside somebody's timed region
    if (!b200::g_sweeps_host)
      cudaMallocHost((void**)&b200::g_sweeps_host, b200::kProfSlots * sizeof(int));
    while (b200::g_free.size() < 12288) {
      b200::ProfRec r;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      r.flops = 0.0; r.slot = 0; r.kind = 0;
      b200::g_free.push_back(r);
    }
  }
  return B200_OK;
}
int b200_profile_read_kinds(double* ms4, uint64_t* launches4, double* algorithmic_flops,
                            uint64_t* sweeps) {
  B200_CUDA_CHECK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(b200::g_prof_mu);
  double ms[4] = {0, 0, 0, 0}, fl = 0.0;
  uint64_t n[4] = {0, 0, 0, 0}, sw = 0;
  for (auto& r : b200::g_pending) {
    float t = 0.f;
    const int kd = (r.kind >= 0 && r.kind < 4) ? r.kind : 3;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[kd] += t; fl += r.flops; ++n[kd];
      if (r.slot >= 0) sw += (uint64_t)b200::g_sweeps_host[r.slot];
    }
    b200::g_free.push_back(r);
  }
  b200::g_pending.clear();
  for (int k = 0; k < 4; ++k) {
    if (ms4) ms4[k] = ms[k];
    if (launches4) launches4[k] = n[k];
  }
  if (algorithmic_flops) *algorithmic_flops = fl;
  if (sweeps) *sweeps = sw;
  return B200_OK;
}
int b200_profile_read(double* kernel_ms, double* algorithmic_flops, uint64_t* launches,
                      uint64_t* sweeps) {
  double ms4[4];
  uint64_t n4[4];
  const int rc = b200_profile_read_kinds(ms4, n4, algorithmic_flops, sweeps);
  if (rc != B200_OK) return rc;
  if (kernel_ms) *kernel_ms = ms4[0];
  if (launches) *launches = n4[0];
  return B200_OK;
}
}
