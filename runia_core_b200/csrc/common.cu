#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace runia {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int finish_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return RUNIA_OK;
}

}  // namespace runia

extern "C" {
int runia_b200_abi_version(void) { return RUNIA_B200_ABI_VERSION; }
const char *runia_b200_last_error(void) { return runia::g_err; }
int64_t runia_b200_launch_count(void) { return runia::g_launches.load(); }
}
