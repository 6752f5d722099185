// dag_glat.cu -- GLAT force-emit masking of the emission plane, for sm_100a.
//
// Between logsoftmax_gather and dag_loss the criterion pins the glanced vertices to their aligned target
// (DASpeech/criterions/nat_dag_loss.py:130-132):
//     prev = keep_word_mask.unsqueeze(1)
//     match_all = match_all.masked_fill(prev, 0) + match_all.masked_fill(~matchmask, -inf).masked_fill(~prev, 0).detach()
// i.e. out[b,t,j] = match[b,t,j]                       where vertex j was not glanced (gradient passes),
//                 = match[b,t,j] (no gradient) / -inf  where it was: kept on its aligned target t, excluded elsewhere.
// Five torch ops and as many [B,M,L] temporaries; here one streaming pass (9 bytes per cell), and the same kernel
// with `backward` set is the gradient (grad * !glanced).  HBM-bound.
#include "common.cuh"

namespace dagb200 {

constexpr int kGlatRows = 4;

__global__ void __launch_bounds__(256)
glat_force_emit_kernel(const float *__restrict__ match, const unsigned char *__restrict__ matchmask,
                       const unsigned char *__restrict__ keep, float *__restrict__ out, int M, int L, int backward, bool vec) {
  const int b = blockIdx.z, t = blockIdx.y;
  const int64_t row = ((int64_t)b * M + t) * L;
  const unsigned char *kp = keep + (int64_t)b * L;
  const float ninf = neg_inf_f();
  if (vec) {
    // kGlatRows target rows per CTA, all loads first: four independent 16-byte loads in flight per thread
    const int j = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (j >= L) return;
    const uchar4 k = *reinterpret_cast<const uchar4 *>(kp + j);
    const int t0 = blockIdx.y * kGlatRows;
    float4 m[kGlatRows];
    uchar4 a[kGlatRows];
#pragma unroll
    for (int r = 0; r < kGlatRows; r++) {
      const int64_t off = ((int64_t)b * M + min(t0 + r, M - 1)) * L + j;
      m[r] = __ldcs(reinterpret_cast<const float4 *>(match + off));
      if (!backward) a[r] = *reinterpret_cast<const uchar4 *>(matchmask + off);
    }
#pragma unroll
    for (int r = 0; r < kGlatRows; r++) {
      if (t0 + r >= M) break;
      float4 o;
      if (backward) {
        o = make_float4(k.x ? 0.f : m[r].x, k.y ? 0.f : m[r].y, k.z ? 0.f : m[r].z, k.w ? 0.f : m[r].w);
      } else {
        o = make_float4(k.x ? (a[r].x ? m[r].x : ninf) : m[r].x, k.y ? (a[r].y ? m[r].y : ninf) : m[r].y,
                        k.z ? (a[r].z ? m[r].z : ninf) : m[r].z, k.w ? (a[r].w ? m[r].w : ninf) : m[r].w);
      }
      *reinterpret_cast<float4 *>(out + ((int64_t)b * M + t0 + r) * L + j) = o;
    }
  } else {
    for (int j = blockIdx.x * 256 + threadIdx.x; j < L; j += gridDim.x * 256) {
      const float m = match[row + j];
      const bool k = kp[j] != 0;
      out[row + j] = backward ? (k ? 0.f : m) : (k ? (matchmask[row + j] ? m : ninf) : m);
    }
  }
}

}  // namespace dagb200

using namespace dagb200;

extern "C" int dagb200_glat_force_emit(const float *match, const unsigned char *matchmask, const unsigned char *keep_word_mask,
                                       float *out, int B, int M, int L, int backward, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && M >= 0 && L >= 0, DAGB200_EINVAL, "glat_force_emit: bad sizes B=%d M=%d L=%d", B, M, L);
  if ((int64_t)B * M * L == 0) return 0;
  DAGB200_CHECK_ARG(match && keep_word_mask && out && (matchmask || backward), DAGB200_EINVAL, "glat_force_emit: null pointer");
  DAGB200_CHECK_ARG(M < 65536 && B < 65536, DAGB200_ELIMIT, "glat_force_emit: B and M must be below 65536");
  const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(match) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(keep_word_mask) & 3) == 0) &&
                   (backward || (reinterpret_cast<uintptr_t>(matchmask) & 3) == 0);
  const int per = vec ? 1024 : 256;
  dim3 grid((unsigned)((L + per - 1) / per), (unsigned)(vec ? (M + kGlatRows - 1) / kGlatRows : M), (unsigned)B);
  glat_force_emit_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(match, matchmask, keep_word_mask, out, M, L, backward, vec);
  DAGB200_CHECK_LAUNCH("glat_force_emit_kernel");
  return 0;
}
