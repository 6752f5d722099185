// dag_grad.cu -- gradients of the DAG log-marginal w.r.t. emissions and transitions, for sm_100a.
//
// Replaces calculate_grad_match_all_kernel (reference dag_loss.cu:378-401) and
// calculate_grad_links_kernel (reference dag_loss.cu:432-485).
//
//   gm[b,t,j] = exp(alpha + beta - match - Z) * go[b]          (0 where match or Z is +-inf)
//   gl[b,i,k] = go[b] * sum_{t=0}^{Tn-2} exp(alpha[t,i] + beta[t+1,i+k+1] + links[i,k] - Z)
//
// v1 grad_links: the reference launches one 4-lane group per (i,k) walking alpha/beta down a COLUMN
// (stride L floats, uncoalesced, every beta element re-read ~L times from L2).  Here a CTA owns 8
// source vertices; their alpha columns are staged once in shared memory, lanes map to destination
// vertices so beta[t+1][n] and the grad_links row are contiguous across the warp, and each beta load
// feeds 8 accumulators.  Both outputs are written exactly once including the zero padding (the
// reference needs at::zeros launches first).
#include "common.cuh"

namespace dagb200 {

template <typename T>
__global__ void __launch_bounds__(256)
grad_match_kernel(const T *__restrict__ go, const T *__restrict__ alpha, const T *__restrict__ beta,
                  const T *__restrict__ match, T *__restrict__ gm, int64_t lat, int64_t total) {
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(x / lat);
    const T Z = __ldg(beta + (int64_t)b * lat);
    const T m = match[x];
    T r = (T)0;
    if (!isinf(m) && !isinf(Z)) r = acc_exp(alpha[x] + beta[x] - m - Z) * __ldg(go + b);
    gm[x] = r;
  }
}

// float4 fast path (lat % 4 == 0 and 16-byte aligned bases)
__global__ void __launch_bounds__(256)
grad_match_kernel_v4(const float *__restrict__ go, const float4 *__restrict__ alpha, const float4 *__restrict__ beta,
                     const float4 *__restrict__ match, float4 *__restrict__ gm, const float *__restrict__ beta_s,
                     int64_t lat4, int64_t total4) {
  for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total4; x += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(x / lat4);
    const float Z = __ldg(beta_s + (int64_t)b * lat4 * 4);
    const float g = __ldg(go + b);
    const float4 m = __ldcs(match + x), a = __ldcs(alpha + x), be = __ldcs(beta + x);
    const bool zinf = isinf(Z);
    float4 r;
    r.x = (zinf || isinf(m.x)) ? 0.f : expf(a.x + be.x - m.x - Z) * g;
    r.y = (zinf || isinf(m.y)) ? 0.f : expf(a.y + be.y - m.y - Z) * g;
    r.z = (zinf || isinf(m.z)) ? 0.f : expf(a.z + be.z - m.z - Z) * g;
    r.w = (zinf || isinf(m.w)) ? 0.f : expf(a.w + be.w - m.w - Z) * g;
    __stcs(gm + x, r);
  }
}

constexpr int kGlRows = 8;      // source vertices per CTA
constexpr int kGlThreads = 256; // destination vertices per sweep

template <typename T>
__global__ void __launch_bounds__(kGlThreads)
grad_links_kernel(const T *__restrict__ go, const T *__restrict__ alpha, const T *__restrict__ beta,
                  const T *__restrict__ links, const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen,
                  T *__restrict__ gl, int M, int L, int Tl) {
  extern __shared__ __align__(16) unsigned char gl_smem_raw[];
  T *acol = reinterpret_cast<T *>(gl_smem_raw);  // [M][kGlRows] alpha columns i0..i0+7
  const int b = blockIdx.y;
  const int i0 = blockIdx.x * kGlRows;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t lat = (int64_t)M * L;
  const T *a = alpha + b * lat;
  const T *be = beta + b * lat;
  const T *E = links + (int64_t)b * L * Tl;
  T *g = gl + (int64_t)b * L * Tl;
  const T Z = be[0];
  const T gout = go[b];
  const bool dead = isinf(Z) || O > L || Tn > M || Tn < 2 || O < 2;
  const int steps = dead ? 0 : Tn - 1;

  for (int x = threadIdx.x; x < steps * kGlRows; x += kGlThreads) {
    const int t = x / kGlRows, ii = x % kGlRows;
    acol[x] = (i0 + ii < L) ? a[(int64_t)t * L + i0 + ii] : neg_inf<T>();
  }
  __syncthreads();

  // sweep destinations n = i0+1+c*256+tid; row ii sees k = n-(i0+ii)-1
  const int kmax = Tl - 1 + (kGlRows - 1);
  for (int c0 = 0; c0 <= kmax; c0 += kGlThreads) {
    const int n = i0 + 1 + c0 + threadIdx.x;
    T extra[kGlRows];
    T acc[kGlRows];
    bool any = false;
#pragma unroll
    for (int ii = 0; ii < kGlRows; ii++) {
      const int i = i0 + ii, k = n - i - 1;
      acc[ii] = (T)0;
      extra[ii] = neg_inf<T>();
      if (!dead && i < O && k >= 0 && k < Tl && n < O) { extra[ii] = E[(int64_t)i * Tl + k] - Z; any = true; }
    }
    if (any) {
      const T *bcol = be + n;
      for (int t = 0; t < steps; t++) {
        const T bv = __ldg(bcol + (int64_t)(t + 1) * L);
        const T *ar = acol + t * kGlRows;
#pragma unroll
        for (int ii = 0; ii < kGlRows; ii++) acc[ii] += fast_exp(ar[ii] + bv + extra[ii]);
      }
    }
#pragma unroll
    for (int ii = 0; ii < kGlRows; ii++) {
      const int i = i0 + ii, k = n - i - 1;
      if (i < L && k >= 0 && k < Tl) g[(int64_t)i * Tl + k] = acc[ii] * gout;
    }
  }
}

}  // namespace dagb200

namespace dagb200 {
int g_exact_mode = 0;
}  // namespace dagb200

using namespace dagb200;

extern "C" void dagb200_set_exact(int on) { g_exact_mode = on ? 1 : 0; }
extern "C" int dagb200_get_exact(void) { return g_exact_mode; }

namespace dagb200 {
size_t grad4_workspace_bytes(int B, int M, int L);
bool grad4_supported(int M, int L);
int launch_grad4(const float *go, const float *alpha, const float *beta, const float *match, const float *links,
                 const int64_t *olen, const int64_t *tlen, float *gm, float *gl, int B, int M, int L, int Tl,
                 void *workspace, cudaStream_t st);
}  // namespace dagb200

extern "C" size_t dagb200_dag_loss_backward_workspace_bytes(int B, int M, int L, int T) {
  (void)T;
  if (B <= 0 || M < 2 || L < 1 || !dagb200::grad4_supported(M, L)) return 0;
  return dagb200::grad4_workspace_bytes(B, M, L);
}

extern "C" int dagb200_dag_loss_backward(const void *grad_output, const void *alpha, const void *beta,
                                         const void *match, const void *links, const int64_t *output_length,
                                         const int64_t *target_length, void *grad_match, void *grad_links, int dtype,
                                         int B, int M, int L, int T, int config1, int config2, void *stream);

// Same contract as dagb200_dag_loss_backward plus a device scratch of dagb200_dag_loss_backward_workspace_bytes()
// bytes: with it (fp32, not in exact mode) the two-pass plane kernels of dag_grad4.cu run.
extern "C" int dagb200_dag_loss_backward_ws(const void *grad_output, const void *alpha, const void *beta,
                                            const void *match, const void *links, const int64_t *output_length,
                                            const int64_t *target_length, void *grad_match, void *grad_links, int dtype,
                                            int B, int M, int L, int T, int config1, int config2, void *workspace,
                                            size_t workspace_bytes, void *stream) {
  using namespace dagb200;
  const bool fast = dtype == DAGB200_F32 && !g_exact_mode && workspace && B > 0 && M >= 2 && L >= 1 && T >= 1 &&
                    grad4_supported(M, L) && workspace_bytes >= grad4_workspace_bytes(B, M, L) &&
                    config1 >= 1 && config1 <= 2 && config2 >= 1 && config2 <= 3 && grad_output && alpha && beta && match &&
                    links && output_length && target_length && grad_match && grad_links &&
                    (int64_t)L * T < (1ll << 31) && (int64_t)M * L < (1ll << 31) && B < 65536 && M < 65536;
  if (!fast)
    return dagb200_dag_loss_backward(grad_output, alpha, beta, match, links, output_length, target_length, grad_match,
                                     grad_links, dtype, B, M, L, T, config1, config2, stream);
  return launch_grad4((const float *)grad_output, (const float *)alpha, (const float *)beta, (const float *)match,
                      (const float *)links, output_length, target_length, (float *)grad_match, (float *)grad_links, B, M,
                      L, T, workspace, (cudaStream_t)stream);
}

extern "C" int dagb200_dag_loss_backward(const void *grad_output, const void *alpha, const void *beta,
                                         const void *match, const void *links, const int64_t *output_length,
                                         const int64_t *target_length, void *grad_match, void *grad_links, int dtype,
                                         int B, int M, int L, int T, int config1, int config2, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && M >= 1 && L >= 1 && T >= 1, DAGB200_EINVAL, "dag_loss_backward: bad sizes");
  DAGB200_CHECK_ARG(config1 >= 1 && config1 <= 2, DAGB200_EINVAL, "config1 should be 1~2");
  DAGB200_CHECK_ARG(config2 >= 1 && config2 <= 3, DAGB200_EINVAL, "config2 should be 1~3");
  DAGB200_CHECK_ARG(dtype == DAGB200_F32 || dtype == DAGB200_F64, DAGB200_EDTYPE,
                    "dag_loss_backward: lattice dtype must be float32 or float64 (got %d)", dtype);
  if (B == 0) return 0;
  DAGB200_CHECK_ARG(grad_output && alpha && beta && match && links && output_length && target_length && grad_match && grad_links,
                    DAGB200_EINVAL, "dag_loss_backward: null pointer");
  DAGB200_CHECK_ARG((int64_t)L * T < (1ll << 31) && (int64_t)M * L < (1ll << 31) && B < 65536, DAGB200_ELIMIT,
                    "dag_loss_backward: lattice too large");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t lat = (int64_t)M * L, total = lat * B;
  const int blocks = sm_count() * 8;
  prof_mark(3, st);
  if (dtype == DAGB200_F32) {
    const bool v4 = (lat % 4 == 0) && (((uintptr_t)alpha | (uintptr_t)beta | (uintptr_t)match | (uintptr_t)grad_match) & 15) == 0;
    if (v4)
      grad_match_kernel_v4<<<blocks, 256, 0, st>>>((const float *)grad_output, (const float4 *)alpha, (const float4 *)beta,
                                                  (const float4 *)match, (float4 *)grad_match, (const float *)beta, lat / 4, total / 4);
    else
      grad_match_kernel<float><<<blocks, 256, 0, st>>>((const float *)grad_output, (const float *)alpha, (const float *)beta,
                                                      (const float *)match, (float *)grad_match, lat, total);
    DAGB200_CHECK_LAUNCH("grad_match_kernel");
    prof_mark(4, st);
    {
      const size_t smem = (size_t)M * kGlRows * sizeof(float);
      DAGB200_CHECK_ARG(smem <= 200 * 1024, DAGB200_ELIMIT, "dag_loss_backward: M=%d too large", M);
      if (smem > 48 * 1024) cudaFuncSetAttribute(grad_links_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((L + kGlRows - 1) / kGlRows, B);
      grad_links_kernel<float><<<grid, kGlThreads, smem, st>>>((const float *)grad_output, (const float *)alpha, (const float *)beta,
                                                              (const float *)links, output_length, target_length, (float *)grad_links, M, L, T);
      DAGB200_CHECK_LAUNCH("grad_links_kernel");
    }
    prof_mark(5, st);
  } else {
    grad_match_kernel<double><<<blocks, 256, 0, st>>>((const double *)grad_output, (const double *)alpha, (const double *)beta,
                                                     (const double *)match, (double *)grad_match, lat, total);
    DAGB200_CHECK_LAUNCH("grad_match_kernel<double>");
    const size_t smem = (size_t)M * kGlRows * sizeof(double);
    DAGB200_CHECK_ARG(smem <= 200 * 1024, DAGB200_ELIMIT, "dag_loss_backward: M=%d too large", M);
    if (smem > 48 * 1024) cudaFuncSetAttribute(grad_links_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((L + kGlRows - 1) / kGlRows, B);
    grad_links_kernel<double><<<grid, kGlThreads, smem, st>>>((const double *)grad_output, (const double *)alpha, (const double *)beta,
                                                             (const double *)links, output_length, target_length, (double *)grad_links, M, L, T);
    DAGB200_CHECK_LAUNCH("grad_links_kernel<double>");
  }
  return 0;
}
