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

// ---- optional per-kernel timing: CUDA events recorded on the launch stream around each kernel -------------
static int g_profile = 0;
static cudaEvent_t g_ev[kProfMarks];
static bool g_ev_init = false;
static bool g_ev_set[kProfMarks];

void prof_mark(int idx, cudaStream_t st) {
  if (!g_profile || idx < 0 || idx >= kProfMarks) return;
  if (!g_ev_init) {
    for (int i = 0; i < kProfMarks; i++) { cudaEventCreate(&g_ev[i]); g_ev_set[i] = false; }
    g_ev_init = true;
  }
  cudaEventRecord(g_ev[idx], st);
  g_ev_set[idx] = true;
}

}  // namespace dagb200

extern "C" void dagb200_set_profile(int on) {
  dagb200::g_profile = on ? 1 : 0;
  if (dagb200::g_ev_init) for (int i = 0; i < dagb200::kProfMarks; i++) dagb200::g_ev_set[i] = false;
}

// ms[0] = dag_prep_kernel, ms[1] = blocked alpha/beta kernel (or the log-domain one), ms[2] = grad_match,
// ms[3] = grad_links, ms[4] = Viterbi kernel of the most recent calls; -1 where nothing was recorded.
extern "C" int dagb200_get_profile(float *ms, int n) {
  using namespace dagb200;
  static const int span[5][2] = {{0, 1}, {1, 2}, {3, 4}, {4, 5}, {6, 7}};
  for (int i = 0; i < n && i < 5; i++) {
    ms[i] = -1.f;
    const int a = span[i][0], b = span[i][1];
    if (g_ev_init && g_ev_set[a] && g_ev_set[b]) {
      cudaEventSynchronize(g_ev[b]);
      float t = 0.f;
      if (cudaEventElapsedTime(&t, g_ev[a], g_ev[b]) == cudaSuccess) ms[i] = t;
    }
  }
  return 0;
}

extern "C" int dagb200_version(void) { return DAGB200_VERSION; }
extern "C" const char *dagb200_last_error(void) { return dagb200::g_err; }
