// cabi.cu -- host-side plumbing shared by the extern "C" entry points of libdagb200.so.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dagb200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *where) {
  set_error("%s: CUDA error %d (%s)", where, (int)e, cudaGetErrorString(e));
  return (int)e;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace dagb200

extern "C" int dagb200_version(void) { return DAGB200_VERSION; }
extern "C" const char *dagb200_last_error(void) { return dagb200::g_err; }
