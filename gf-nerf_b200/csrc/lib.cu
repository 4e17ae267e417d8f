// Library-level C-ABI: error string, version, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace gf {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace gf

extern "C" {

const char* gf_last_error(void) { return gf::g_err; }
const char* gf_version(void) { return "gfnerf_b200 0.1.0 sm_100a"; }
int64_t gf_launch_count(void) { return gf::g_launches.load(); }

}  // extern "C"
