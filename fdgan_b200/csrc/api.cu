// Error reporting, launch accounting and version of the fdgan_b200 C ABI.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace fdg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return FDG_ECUDA;
  }
  count_launch(1);
  return FDG_OK;
}

}  // namespace fdg

extern "C" {

const char* fdg_last_error(void) { return fdg::g_err; }

int fdg_version(void) { return 100; }

int64_t fdg_launch_count(void) { return fdg::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
