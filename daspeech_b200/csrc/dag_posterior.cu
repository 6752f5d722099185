// dag_posterior.cu -- alignment posterior of the S2S criterion, for sm_100a.
//
// The reference computes P(a_t = j | x, y) from the lattices that dag_loss_with_alpha_beta returns with five torch ops
// (DASpeech/criterions/s2s_dag_fastspeech2_loss.py:259-261): alpha + beta, logsumexp_keepdim over the vertices
// (custom_ops/dag_loss.py:303-311: max, masked exp / sum, log), the subtraction, exp, NaN -> 0 for the rows that are
// entirely -inf, and a cast to the feature dtype -- seven passes over [B, M, L] temporaries.  Here one CTA owns one
// (utterance, target) row: alpha and beta are read ONCE with 128-bit loads, the row stays in registers for the max and
// the sum, and the posterior is written once in the requested dtype.  HBM-bound: 8 + sizeof(out) bytes per cell.
// The contraction with the decoder features that follows (torch.matmul, :262) stays on cuBLAS.
#include "common.cuh"

namespace dagb200 {

constexpr int kPostThreads = 256;

template <typename T> __device__ __forceinline__ T post_cast(float v);
template <> __device__ __forceinline__ float post_cast<float>(float v) { return v; }
template <> __device__ __forceinline__ __half post_cast<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 post_cast<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float post_block_reduce(float v, bool is_max, float *red) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = red[l & 7];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float x = __shfl_xor_sync(0xffffffffu, r, o);
    r = is_max ? fmaxf(r, x) : (r + x);
  }
  return r;
}

// NV float4 per thread: L <= NV * 1024, L % 4 == 0, 16-byte aligned rows
template <typename OutT, int NV>
__global__ void __launch_bounds__(kPostThreads)
dag_posterior_vec_kernel(const float *__restrict__ alpha, const float *__restrict__ beta, OutT *__restrict__ score, int L) {
  __shared__ float red[16];
  const int64_t row = blockIdx.x;
  const float4 *a4 = reinterpret_cast<const float4 *>(alpha + row * L);
  const float4 *b4 = reinterpret_cast<const float4 *>(beta + row * L);
  const int nvec = L >> 2;
  float x[NV][4];
  float mx = neg_inf_f();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const int c = k * kPostThreads + threadIdx.x;
    if (c < nvec) {
      const float4 a = __ldcs(a4 + c), b = __ldcs(b4 + c);
      x[k][0] = a.x + b.x; x[k][1] = a.y + b.y; x[k][2] = a.z + b.z; x[k][3] = a.w + b.w;
#pragma unroll
      for (int e = 0; e < 4; e++) mx = fmaxf(mx, x[k][e]);      // fmaxf ignores a NaN operand ((-inf) + (+inf) cannot occur)
    } else {
#pragma unroll
      for (int e = 0; e < 4; e++) x[k][e] = neg_inf_f();
    }
  }
  mx = post_block_reduce(mx, true, red);
  const bool empty = !(mx > neg_inf_f());       // a row without any finite cell: the reference's NaN -> 0
  float sum = 0.f;
  const float mo2 = (empty ? 0.f : mx) * kLog2e;        // exp(x - m) = 2^(x log2e - m log2e): one FFMA + one MUFU.EX2
#pragma unroll
  for (int k = 0; k < NV; k++)
#pragma unroll
    for (int e = 0; e < 4; e++) {
      float pz;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pz) : "f"(fmaf(x[k][e], kLog2e, -mo2)));
      x[k][e] = empty ? 0.f : pz;
      sum += x[k][e];
    }
  sum = post_block_reduce(sum, false, red + 8);
  const float inv = empty ? 0.f : __fdividef(1.f, sum);
  OutT *o = score + row * L;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const int c = k * kPostThreads + threadIdx.x;
    if (c < nvec) {
      if (sizeof(OutT) == 4) {
        reinterpret_cast<float4 *>(o)[c] = make_float4(x[k][0] * inv, x[k][1] * inv, x[k][2] * inv, x[k][3] * inv);
      } else {
        OutT v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = post_cast<OutT>(x[k][e] * inv);
        reinterpret_cast<uint2 *>(o)[c] = *reinterpret_cast<const uint2 *>(v);
      }
    }
  }
}

// any L / alignment: three passes, the second and third hit L1 / L2
template <typename OutT>
__global__ void __launch_bounds__(kPostThreads)
dag_posterior_stream_kernel(const float *__restrict__ alpha, const float *__restrict__ beta, OutT *__restrict__ score, int L) {
  __shared__ float red[16];
  const int64_t row = blockIdx.x;
  const float *a = alpha + row * L, *b = beta + row * L;
  float mx = neg_inf_f();
  for (int j = threadIdx.x; j < L; j += kPostThreads) mx = fmaxf(mx, a[j] + b[j]);
  mx = post_block_reduce(mx, true, red);
  const bool empty = !(mx > neg_inf_f());
  float sum = 0.f;
  if (!empty)
    for (int j = threadIdx.x; j < L; j += kPostThreads) sum += __expf(a[j] + b[j] - mx);
  sum = post_block_reduce(sum, false, red + 8);
  const float inv = empty ? 0.f : __fdividef(1.f, sum);
  OutT *o = score + row * L;
  for (int j = threadIdx.x; j < L; j += kPostThreads) o[j] = post_cast<OutT>(empty ? 0.f : __expf(a[j] + b[j] - mx) * inv);
}

template <typename OutT>
static int launch_posterior(const float *alpha, const float *beta, OutT *score, int B, int M, int L, cudaStream_t st) {
  const unsigned rows = (unsigned)((int64_t)B * M);
  const bool aligned = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(alpha) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(beta) & 15) == 0) && ((reinterpret_cast<uintptr_t>(score) & 15) == 0);
  const int need = (L / 4 + kPostThreads - 1) / kPostThreads;
  if (aligned && need <= 1) dag_posterior_vec_kernel<OutT, 1><<<rows, kPostThreads, 0, st>>>(alpha, beta, score, L);
  else if (aligned && need <= 2) dag_posterior_vec_kernel<OutT, 2><<<rows, kPostThreads, 0, st>>>(alpha, beta, score, L);
  else if (aligned && need <= 4) dag_posterior_vec_kernel<OutT, 4><<<rows, kPostThreads, 0, st>>>(alpha, beta, score, L);
  else dag_posterior_stream_kernel<OutT><<<rows, kPostThreads, 0, st>>>(alpha, beta, score, L);
  DAGB200_CHECK_LAUNCH("dag_posterior_kernel");
  return 0;
}

}  // namespace dagb200

using namespace dagb200;

extern "C" int dagb200_dag_posterior(const float *alpha, const float *beta, void *score, int out_dtype, int B, int M, int L,
                                     void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && M >= 0 && L >= 1, DAGB200_EINVAL, "dag_posterior: bad sizes B=%d M=%d L=%d", B, M, L);
  if ((int64_t)B * M == 0) return 0;
  DAGB200_CHECK_ARG(alpha && beta && score, DAGB200_EINVAL, "dag_posterior: null pointer");
  DAGB200_CHECK_ARG((int64_t)B * M < (1ll << 31), DAGB200_ELIMIT, "dag_posterior: B*M too large");
  cudaStream_t st = (cudaStream_t)stream;
  switch (out_dtype) {
    case DAGB200_F32: return launch_posterior<float>(alpha, beta, (float *)score, B, M, L, st);
    case DAGB200_F16: return launch_posterior<__half>(alpha, beta, (__half *)score, B, M, L, st);
    case DAGB200_BF16: return launch_posterior<__nv_bfloat16>(alpha, beta, (__nv_bfloat16 *)score, B, M, L, st);
    default:
      set_error("dag_posterior: unsupported output dtype %d", out_dtype);
      return DAGB200_EDTYPE;
  }
}
