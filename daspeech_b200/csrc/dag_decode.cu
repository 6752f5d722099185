// dag_decode.cu -- inference decoding over the DAG, for sm_100a.
//
// Replaces the decoding loops of DASpeech/models/s2s_conformer_dag_fastspeech2.py:211-304 -- Python walks over
// `.tolist()`-ed back-pointers, one host synchronisation per batch and a Python loop per token:
//   * greedy / lookahead (:218-244): next[i] = argmax_j (links[i][j] + beta * max_y log P(y | v_j)), then the chain
//     0 -> next[0] -> ... up to the last vertex, emitting the arg-max token of every visited vertex, consecutive
//     duplicates and pad removed.  One kernel: the arg-max rows in parallel, the walk by one thread per utterance.
//   * viterbi / jointviterbi (:245-304): the free-length max-plus recurrence is EXACTLY the max-plus lattice of the
//     training-time alignment with a length-independent emission plane (daspeech_b200/decode.py builds it and runs the
//     wave kernel of dag_viterbi3.cu with the lattice output); this file holds what follows: the end-transition, the
//     length penalty, the choice of the length, the backtrace -- back-pointers recomputed for the cells on the path, with
//     torch.max's tie-break (first = smallest source vertex) -- and the token de-duplication.
#include "common.cuh"

namespace dagb200 {

constexpr int kDecThreads = 256;

// value-then-smaller-index arg-max over the block; every thread gets the result
__device__ __forceinline__ void block_argmax_first(float &v, int &ix, float *sv, int *si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
    if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
  }
  __syncthreads();
  if (lane == 0) { sv[warp] = v; si[warp] = ix; }
  __syncthreads();
  v = sv[0]; ix = si[0];
  for (int w = 1; w < kDecThreads / 32; w++)
    if (sv[w] > v || (sv[w] == v && si[w] < ix)) { v = sv[w]; ix = si[w]; }
}

// ---- greedy / lookahead ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDecThreads)
decode_lookahead_kernel(const float *__restrict__ links, const float *__restrict__ vlogit, const int64_t *__restrict__ vtoken,
                        const int64_t *__restrict__ olen, float beta, int64_t pad, int L, int Tl,
                        int64_t *__restrict__ out_tok, int32_t *__restrict__ out_vtx, int32_t *__restrict__ out_len) {
  extern __shared__ int dec_next[];                   // [L] successor of every vertex
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int O = min((int)olen[b], L);
  const float *E = links + (int64_t)b * L * Tl;
  const float *lg = vlogit + (int64_t)b * L;
  const float ninf = neg_inf_f();
  // the reference takes the arg-max over ALL vertices of the dense row (-inf where there is no transition): the first
  // index of the maximum, i.e. 0 when the row has no finite entry
  for (int i = warp; i < L; i += kDecThreads / 32) {
    float bv = ninf; int bj = 0;
    for (int k = lane; k < Tl && i + k + 1 < L; k += 32) {
      const int j = i + k + 1;
      const float x = E[(int64_t)i * Tl + k] + (beta != 0.f ? lg[j] * beta : 0.f);
      if (x > bv) { bv = x; bj = j; }                  // ascending j inside a lane: the first maximum stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ov > bv || (ov == bv && oj < bj && ov > ninf)) { bv = ov; bj = oj; }
    }
    if (lane == 0) dec_next[i] = bv > ninf ? bj : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t *tok = out_tok + (int64_t)b * L;
    int32_t *vtx = out_vtx + (int64_t)b * L;
    int64_t last = vtoken[(int64_t)b * L];            // <bos>: the token of vertex 0, without a feature
    int nt = 0, nf = 0;
    tok[nt++] = last;
    int j = 0;
    for (int guard = 0; guard < L && j != O - 1; guard++) {
      j = dec_next[j];
      const int64_t now = vtoken[(int64_t)b * L + j];
      if (now != pad && now != last) { tok[nt++] = now; vtx[nf++] = j; }
      last = now;
      if (j == 0) break;                               // a vertex without successors: the reference would loop forever
    }
    out_len[2 * b] = nt;
    out_len[2 * b + 1] = nf;
    for (int x = nt; x < L; x++) tok[x] = pad;
    for (int x = nf; x < L; x++) vtx[x] = -1;
  }
}

// ---- viterbi / jointviterbi: everything after the max-plus recurrence ----------------------------------------------
// lattice [B][S+1][L]: row s+1 = the reference's scores[s] (row 0 is the start row of the alignment kernel).
__global__ void __launch_bounds__(kDecThreads)
decode_viterbi_finish_kernel(const float *__restrict__ lattice, const float *__restrict__ links,
                             const int64_t *__restrict__ vtoken, const int64_t *__restrict__ olen, float viterbibeta,
                             int64_t pad, int S, int L, int Tl, int64_t *__restrict__ out_tok, int32_t *__restrict__ out_vtx,
                             int32_t *__restrict__ out_len, int32_t *__restrict__ path_scratch) {
  __shared__ float sv[kDecThreads / 32];
  __shared__ int si[kDecThreads / 32];
  __shared__ float s_best;
  __shared__ int s_len, s_start;
  const int b = blockIdx.x;
  const int O = min((int)olen[b], L);
  const float *lat = lattice + (int64_t)b * (S + 1) * L;
  const float *E = links + (int64_t)b * L * Tl;
  const float ninf = neg_inf_f();
  // score of ending after s+1 vertices: max_j (scores[s][j] + links[j][O-1]) / (s+1)^viterbibeta; the first maximum over s
  if (threadIdx.x == 0) { s_best = ninf; s_len = 1; s_start = 0; }
  for (int s = 0; s < S; s++) {
    float bv = ninf; int bj = 0x7fffffff;
    for (int j = threadIdx.x; j < L; j += kDecThreads) {
      const int k = O - 1 - j - 1;
      const float last = (k >= 0 && k < Tl) ? E[(int64_t)j * Tl + k] : ninf;      // dense links[j][O-1]
      const float x = lat[(int64_t)(s + 1) * L + j] + last;
      if (x > bv) { bv = x; bj = j; }
    }
    if (!(bv > ninf)) bj = 0;
    block_argmax_first(bv, bj, sv, si);
    if (threadIdx.x == 0) {
      const float sc = bv / powf((float)(s + 1), viterbibeta);
      if (sc > s_best || s == 0) { s_best = sc; s_len = s + 1; s_start = bv > ninf ? bj : 0; }
    }
    __syncthreads();
  }
  // backtrace: vertex of step s from the vertex of step s+1, arg-max recomputed (first = smallest source on ties)
  int32_t *pth = path_scratch + (int64_t)b * S;
  const int len = s_len;
  int j = s_start;
  if (threadIdx.x == 0) pth[len - 1] = j;
  for (int s = len - 1; s >= 1; s--) {
    float bv = ninf; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < j; i += kDecThreads) {
      const int k = j - i - 1;
      if (k < Tl) {
        const float x = lat[(int64_t)s * L + i] + E[(int64_t)i * Tl + k];       // scores[s-1][i] + links[i][j]
        if (x > bv) { bv = x; bi = i; }
      }
    }
    if (!(bv > ninf)) bi = 0x7fffffff;
    block_argmax_first(bv, bi, sv, si);
    j = bv > ninf ? bi : 0;                                                      // torch.max over an all -inf column: index 0
    if (threadIdx.x == 0) pth[s - 1] = j;
    __syncthreads();
  }
  // tokens along the path (the reference builds them back to front: a vertex is kept when its token differs from the
  // token of the vertex AFTER it), pad removed; written front to back
  if (threadIdx.x == 0) {
    int64_t *tok = out_tok + (int64_t)b * L;
    int32_t *vtx = out_vtx + (int64_t)b * L;
    int n = 0;
    for (int s = 0; s < len; s++) {
      const int v = pth[s];
      const int64_t now = vtoken[(int64_t)b * L + v];
      bool keep;
      if (s == len - 1) keep = true;                                             // the last vertex of the path is always kept
      else keep = now != pad && now != vtoken[(int64_t)b * L + pth[s + 1]];
      if (keep) { tok[n] = now; vtx[n] = v; n++; }
    }
    out_len[2 * b] = n;
    out_len[2 * b + 1] = n;
    for (int x = n; x < L; x++) { tok[x] = pad; vtx[x] = -1; }
  }
}

}  // namespace dagb200

using namespace dagb200;

extern "C" int dagb200_decode_lookahead(const float *links, const float *vertex_logit, const int64_t *vertex_token,
                                        const int64_t *output_length, float beta, int64_t pad, int B, int L, int T,
                                        int64_t *out_tokens, int32_t *out_vertices, int32_t *out_lengths, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && L >= 2 && T >= 1, DAGB200_EINVAL, "decode_lookahead: bad sizes B=%d L=%d T=%d", B, L, T);
  if (B == 0) return 0;
  DAGB200_CHECK_ARG(links && vertex_logit && vertex_token && output_length && out_tokens && out_vertices && out_lengths,
                    DAGB200_EINVAL, "decode_lookahead: null pointer");
  const size_t smem = (size_t)L * sizeof(int);
  DAGB200_CHECK_ARG(smem <= 200 * 1024, DAGB200_ELIMIT, "decode_lookahead: L=%d too large", L);
  if (smem > 48 * 1024) cudaFuncSetAttribute(decode_lookahead_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  decode_lookahead_kernel<<<B, kDecThreads, smem, (cudaStream_t)stream>>>(links, vertex_logit, vertex_token, output_length, beta,
                                                                        pad, L, T, out_tokens, out_vertices, out_lengths);
  DAGB200_CHECK_LAUNCH("decode_lookahead_kernel");
  return 0;
}

extern "C" int dagb200_decode_viterbi_finish(const float *lattice, const float *links, const int64_t *vertex_token,
                                             const int64_t *output_length, float viterbibeta, int64_t pad, int B, int S,
                                             int L, int T, int64_t *out_tokens, int32_t *out_vertices, int32_t *out_lengths,
                                             int32_t *path_scratch, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && L >= 2 && T >= 1 && S >= 1, DAGB200_EINVAL, "decode_viterbi_finish: bad sizes");
  if (B == 0) return 0;
  DAGB200_CHECK_ARG(lattice && links && vertex_token && output_length && out_tokens && out_vertices && out_lengths && path_scratch,
                    DAGB200_EINVAL, "decode_viterbi_finish: null pointer");
  decode_viterbi_finish_kernel<<<B, kDecThreads, 0, (cudaStream_t)stream>>>(lattice, links, vertex_token, output_length,
                                                                          viterbibeta, pad, S, L, T, out_tokens, out_vertices,
                                                                          out_lengths, path_scratch);
  DAGB200_CHECK_LAUNCH("decode_viterbi_finish_kernel");
  return 0;
}
