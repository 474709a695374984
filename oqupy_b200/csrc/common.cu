#include <stdarg.h>
#include <atomic>

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
}  // namespace b200

extern "C" {
const char* b200_last_error(void) { return b200::g_err; }
int b200_abi_version(void) { return 1; }
uint64_t b200_launch_count(void) { return b200::g_launches.load(); }
}
