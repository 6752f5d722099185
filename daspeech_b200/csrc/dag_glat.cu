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

// ---- what the glancing pass derives from the Viterbi path (DASpeech/criterions/nat_dag_loss.py:223-227):
//     predict_align_mask = path >= 0
//     matchmask = zeros(B, M + 1, L, bool).scatter_(1, path.unsqueeze(1) + 1, 1)[:, 1:]      matchmask[b,t,j] = (path[b,j] == t)
//     oracle    = tgt_tokens.gather(-1, path.clip(min=0))
//     same_num  = ((pred_tokens == oracle) & predict_align_mask).sum(1)
// A [B, M+1, L] zero fill, a scatter, a slice, a gather and a reduction; here one kernel: the path row of the utterance
// sits in shared memory, every thread writes 16 mask bytes per store (the plane is written exactly once), and the first
// CTA of an utterance also emits the oracle tokens, the int64 path and the match count.
constexpr int kGaThreads = 256;
constexpr int kGaRows = 16;      // target rows per CTA

__global__ void __launch_bounds__(kGaThreads)
glat_alignment_kernel(const int32_t *__restrict__ path, const int64_t *__restrict__ tgt, int64_t tsb, int64_t tss,
                      const int64_t *__restrict__ pred, unsigned char *__restrict__ matchmask, int64_t *__restrict__ oracle,
                      int64_t *__restrict__ path64, unsigned char *__restrict__ align_mask, int64_t *__restrict__ same_num,
                      int M, int L) {
  extern __shared__ int32_t ga_path[];      // [L rounded up to 16]
  __shared__ int ga_red[kGaThreads / 32];
  const int b = blockIdx.y;
  const int32_t *prow = path + (int64_t)b * L;
  const int L16 = (L + 15) & ~15;
  for (int j = threadIdx.x; j < L16; j += kGaThreads) ga_path[j] = j < L ? prow[j] : -2;
  __syncthreads();
  const int t0 = blockIdx.x * kGaRows, t1 = min(M, t0 + kGaRows);
  unsigned char *mrow = matchmask + ((int64_t)b * M + t0) * L;
  if ((L & 15) == 0 && (reinterpret_cast<uintptr_t>(matchmask) & 15) == 0) {
    const int per_row = L >> 4;
    for (int x = threadIdx.x; x < (t1 - t0) * per_row; x += kGaThreads) {
      const int t = t0 + x / per_row, j = (x % per_row) << 4;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int4 p4 = *reinterpret_cast<const int4 *>(ga_path + j + 4 * q);
        w[q] = (p4.x == t ? 1u : 0u) | (p4.y == t ? 0x100u : 0u) | (p4.z == t ? 0x10000u : 0u) | (p4.w == t ? 0x1000000u : 0u);
      }
      __stcs(reinterpret_cast<uint4 *>(mrow + (int64_t)(t - t0) * L + j), make_uint4(w[0], w[1], w[2], w[3]));
    }
  } else {
    for (int x = threadIdx.x; x < (t1 - t0) * L; x += kGaThreads) {
      const int t = t0 + x / L, j = x % L;
      mrow[(int64_t)(t - t0) * L + j] = ga_path[j] == t ? 1 : 0;
    }
  }
  if (blockIdx.x == 0) {
    int same = 0;
    for (int j = threadIdx.x; j < L; j += kGaThreads) {
      const int pj = ga_path[j];
      const int64_t o = tgt[b * tsb + (int64_t)max(pj, 0) * tss];
      if (oracle) oracle[(int64_t)b * L + j] = o;
      if (path64) path64[(int64_t)b * L + j] = pj;
      if (align_mask) align_mask[(int64_t)b * L + j] = pj >= 0 ? 1 : 0;
      if (pred && pj >= 0 && pred[(int64_t)b * L + j] == o) same++;
    }
    same = __reduce_add_sync(0xffffffffu, same);
    if ((threadIdx.x & 31) == 0) ga_red[threadIdx.x >> 5] = same;
    __syncthreads();
    if (threadIdx.x == 0 && same_num) {
      int tot = 0;
      for (int w = 0; w < kGaThreads / 32; w++) tot += ga_red[w];
      same_num[b] = tot;
    }
  }
}

}  // namespace dagb200

using namespace dagb200;

extern "C" int dagb200_glat_alignment(const int32_t *path, const int64_t *tgt_tokens, int64_t tsb, int64_t tss,
                                      const int64_t *pred_tokens, unsigned char *matchmask, int64_t *oracle, int64_t *path64,
                                      unsigned char *align_mask, int64_t *same_num, int B, int M, int L, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && M >= 1 && L >= 1, DAGB200_EINVAL, "glat_alignment: bad sizes B=%d M=%d L=%d", B, M, L);
  if (B == 0) return 0;
  DAGB200_CHECK_ARG(path && tgt_tokens && matchmask, DAGB200_EINVAL, "glat_alignment: null pointer");
  DAGB200_CHECK_ARG(B < 65536 && (size_t)((L + 15) & ~15) * 4 <= 200 * 1024, DAGB200_ELIMIT, "glat_alignment: B=%d L=%d too large", B, L);
  const size_t smem = (size_t)((L + 15) & ~15) * sizeof(int32_t);
  if (smem > 48 * 1024) cudaFuncSetAttribute(glat_alignment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)((M + kGaRows - 1) / kGaRows), (unsigned)B);
  glat_alignment_kernel<<<grid, kGaThreads, smem, (cudaStream_t)stream>>>(path, tgt_tokens, tsb, tss, pred_tokens, matchmask, oracle,
                                                                       path64, align_mask, same_num, M, L);
  DAGB200_CHECK_LAUNCH("glat_alignment_kernel");
  return 0;
}

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
