// dag_dp.cu -- log-domain forward (alpha) / backward (beta) recurrences and the max-plus Viterbi over the
// target-token x graph-vertex lattice, for sm_100a.
//
// Replaces calculate_alpha_kernel / calculate_beta_kernel (reference dag_loss.cu:40-140, 178-274) and
// calculate_maxalpha_kernel + calculate_backtrace_kernel (reference dag_best_alignment.cu:39-130, 170-185).
//
// v1 organisation (exact reference semantics, one pass per cell, no inter-CTA spin-waits):
//   * one CTA per (utterance, direction); alpha and beta chains of the same call run concurrently as
//     blockIdx.y = 0 / 1 of ONE launch (the reference uses two launches and a per-call side stream);
//   * the previous lattice row lives in shared memory (double buffered); each step reads the row's
//     transition log-probs straight from global/L2 with fully coalesced warp accesses:
//       - alpha "pulls through pushes": lane <-> destination vertex j, the warp walks source vertex i
//         uniformly, so links[i][j-i-1] is contiguous across lanes (the reference walks an
//         anti-diagonal here: 4-byte loads at ~4 KB stride);
//       - beta: warp <-> source vertex j, lanes <-> successor k, links[j][k] contiguous, combined with a
//         warp-shuffle log-sum-exp merge;
//   * single-pass online log-sum-exp (one MUFU.EX2 per edge) instead of the reference's max pass + sum pass;
//   * every output row is written once, coalesced, including the -inf padding the reference produces
//     with separate at::zeros + fill_ launches.
#include <cstdlib>

#include "common.cuh"

namespace dagb200 {

template <typename T> struct SmemRows {
  __device__ static T *get() {
    extern __shared__ __align__(16) unsigned char dag_smem_raw[];
    return reinterpret_cast<T *>(dag_smem_raw);
  }
};

__device__ __forceinline__ int check_lengths(int O, int Tn, int L, int M) {
  if (Tn < 2 || O < 2) return DAGB200_ST_LEN_LT2;
  if (O < Tn) return DAGB200_ST_GRAPH_SMALL;
  if (O > L || Tn > M) return DAGB200_ST_GRAPH_SMALL;
  return DAGB200_ST_OK;
}

template <typename T, int THREADS>
__device__ void fill_rows(T *dst, int64_t n) {
  const T ninf = neg_inf<T>();
  for (int64_t x = threadIdx.x; x < n; x += THREADS) dst[x] = ninf;
}

// ------------------------------------------------------------------------------------------------
// alpha[t][j] = match[t][j] + LSE_{d=1..min(j,T)} (alpha[t-1][j-d] + links[j-d][d-1])   (dag_loss.cu:71-131)
template <typename T, int THREADS>
__device__ void alpha_chain(const T *__restrict__ match, const T *__restrict__ links, T *__restrict__ alpha,
                            int O, int Tn, int M, int L, int Tl) {
  T *prev = SmemRows<T>::get();
  T *cur = prev + L;
  const T ninf = neg_inf<T>();
  const int lane = threadIdx.x & 31;

  for (int j = threadIdx.x; j < L; j += THREADS) {
    T v = (j == 0) ? match[0] : ninf;
    prev[j] = v;
    alpha[j] = v;
  }
  __syncthreads();

  for (int t = 1; t < Tn; t++) {
    const T *mrow = match + (int64_t)t * L;
    T *arow = alpha + (int64_t)t * L;
    for (int j0 = threadIdx.x - lane; j0 < L; j0 += THREADS) {
      const int j = j0 + lane;
      T val = ninf;
      // warp-uniform source range: all lanes walk the same i so that links[i][j-i-1] coalesces
      const int jhi = min(j0 + 31, O - 1);            // largest destination in this warp
      const int i_begin = max(t - 1, j0 - Tl);
      if (jhi >= t && j0 < O) {
        const bool jvalid = (j >= t) && (j < O);
        const int ilo = max(t - 1, j - Tl);           // this lane's own lower bound
        Lse<T> a0, a1, a2, a3;
        a0.init(); a1.init(); a2.init(); a3.init();
        const T *e = links + (int64_t)(j - 1);        // &links[i][j-i-1] = e + i*(Tl-1)
        const int64_t es = (int64_t)Tl - 1;
        int i = i_begin;
        for (; i + 4 <= jhi; i += 4) {
          T x0 = ninf, x1 = ninf, x2 = ninf, x3 = ninf;
          if (jvalid && i + 0 >= ilo && i + 0 < j) x0 = prev[i + 0] + __ldg(e + (i + 0) * es);
          if (jvalid && i + 1 >= ilo && i + 1 < j) x1 = prev[i + 1] + __ldg(e + (i + 1) * es);
          if (jvalid && i + 2 >= ilo && i + 2 < j) x2 = prev[i + 2] + __ldg(e + (i + 2) * es);
          if (jvalid && i + 3 >= ilo && i + 3 < j) x3 = prev[i + 3] + __ldg(e + (i + 3) * es);
          a0.add(x0); a1.add(x1); a2.add(x2); a3.add(x3);
        }
        for (; i < jhi; i++) {
          T x0 = ninf;
          if (jvalid && i >= ilo && i < j) x0 = prev[i] + __ldg(e + i * es);
          a0.add(x0);
        }
        if (jvalid) {
          a0.merge(a1); a2.merge(a3); a0.merge(a2);
          val = a0.finish(mrow[j]);
        }
      }
      if (j < L) { cur[j] = val; arow[j] = val; }
    }
    __syncthreads();
    T *tmp = prev; prev = cur; cur = tmp;
  }
  fill_rows<T, THREADS>(alpha + (int64_t)Tn * L, (int64_t)(M - Tn) * L);
}

// beta[t][j] = match[t][j] + LSE_{d=1..min(O-1-j,T)} (beta[t+1][j+d] + links[j][d-1])      (dag_loss.cu:206-265)
template <typename T, int THREADS>
__device__ void beta_chain(const T *__restrict__ match, const T *__restrict__ links, T *__restrict__ beta,
                           int O, int Tn, int M, int L, int Tl) {
  T *next = SmemRows<T>::get();
  T *cur = next + L;
  const T ninf = neg_inf<T>();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = THREADS / 32;

  fill_rows<T, THREADS>(beta + (int64_t)Tn * L, (int64_t)(M - Tn) * L);
  {
    T *brow = beta + (int64_t)(Tn - 1) * L;
    for (int j = threadIdx.x; j < L; j += THREADS) {
      T v = (j == O - 1) ? match[(int64_t)(Tn - 1) * L + j] : ninf;
      next[j] = v;
      brow[j] = v;
    }
  }
  __syncthreads();

  for (int t = Tn - 2; t >= 0; t--) {
    const T *mrow = match + (int64_t)t * L;
    // cells right of jmax cannot reach (Tn-1, O-1) any more: every successor is -inf, so is the cell
    const int jmax = O - 1 - (Tn - 1 - t);
    for (int j = threadIdx.x; j < L; j += THREADS)
      if (j < t || j > jmax) cur[j] = ninf;
    for (int j = t + warp; j <= jmax; j += NW) {
      const int n = min(O - 1 - j, Tl);
      const T *e = links + (int64_t)j * Tl;
      const T *bn = next + j + 1;
      Lse<T> a0, a1;
      a0.init(); a1.init();
      int k = lane;
      for (; k + 32 < n; k += 64) {
        T x0 = bn[k] + __ldg(e + k);
        T x1 = bn[k + 32] + __ldg(e + k + 32);
        a0.add(x0); a1.add(x1);
      }
      if (k < n) a0.add(bn[k] + __ldg(e + k));
      a0.merge(a1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Lse<T> other;
        other.m = __shfl_xor_sync(0xffffffffu, a0.m, o);
        other.s = __shfl_xor_sync(0xffffffffu, a0.s, o);
        a0.merge(other);
      }
      if (lane == 0) cur[j] = a0.finish(mrow[j]);
    }
    __syncthreads();
    T *brow = beta + (int64_t)t * L;
    for (int j = threadIdx.x; j < L; j += THREADS) brow[j] = cur[j];
    // the next iteration only READS the row just produced and overwrites the older one, which every warp
    // finished reading before the barrier above -- no second barrier needed
    T *tmp = next; next = cur; cur = tmp;
  }
}

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
dag_alpha_beta_kernel(const T *__restrict__ match, const T *__restrict__ links, const int64_t *__restrict__ olen,
                      const int64_t *__restrict__ tlen, T *__restrict__ alpha, T *__restrict__ beta, int M, int L,
                      int Tl, int32_t *__restrict__ status) {
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t lat = (int64_t)M * L;
  T *dst = (is_beta ? beta : alpha) + b * lat;
  const int st = check_lengths(O, Tn, L, M);
  if (st != DAGB200_ST_OK) {
    fill_rows<T, THREADS>(dst, lat);
    if (status && threadIdx.x == 0 && !is_beta) status[b] = st;
    return;
  }
  if (status && threadIdx.x == 0 && !is_beta) status[b] = DAGB200_ST_OK;
  const T *m = match + b * lat;
  const T *e = links + (int64_t)b * L * Tl;
  if (is_beta) beta_chain<T, THREADS>(m, e, dst, O, Tn, M, L, Tl);
  else alpha_chain<T, THREADS>(m, e, dst, O, Tn, M, L, Tl);
}

// ------------------------------------------------------------------------------------------------
// Viterbi: same recurrence with (max, argmax).  Candidate value is ONE fp32/fp64 add, the cell value one
// more add, exactly as the reference (dag_best_alignment.cu:102,114), so values and -- through the
// explicit tie-break key below -- arg-max indices are reproducible bit for bit.
// Tie-break of the reference (dag_best_alignment.cu:100-111): candidate d lives in lane (d-1) % W; a lane
// keeps its first strict maximum (smallest d); lanes merge through a shuffle-down tree with strict '>',
// which makes the lane with the smaller BIT-REVERSED index win.  key = bitrev(lane) << 16 | d, smaller wins.
__device__ __forceinline__ int tie_key(int d, int wmask, int wshift) {
  return (int)((__brev((unsigned)((d - 1) & wmask)) >> wshift) << 16) | d;
}

template <typename T> struct Best {
  T v; int key;
  __device__ __forceinline__ void init() { v = neg_inf<T>(); key = -1; }
  __device__ __forceinline__ void add(T x, int k) {
    if (x > v || (x == v && k < key)) { v = x; key = k; }
  }
};

template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
dag_viterbi_kernel(const T *__restrict__ match, const T *__restrict__ links, const int64_t *__restrict__ olen,
                   const int64_t *__restrict__ tlen, T *__restrict__ alpha_out, uint16_t *__restrict__ trace,
                   int32_t *__restrict__ path, int M, int L, int Tl, int wbits, int32_t *__restrict__ status) {
  const int b = blockIdx.x;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t lat = (int64_t)M * L;
  const T ninf = neg_inf<T>();
  const int lane = threadIdx.x & 31;
  T *alpha = alpha_out ? alpha_out + b * lat : nullptr;
  uint16_t *tr = trace + b * lat;
  int32_t *prow = path + (int64_t)b * L;
  for (int j = threadIdx.x; j < L; j += THREADS) prow[j] = -1;

  int st = check_lengths(O, Tn, L, M);
  if (st == DAGB200_ST_OK && (int64_t)(Tn - 1) * Tl + 1 < O) st = DAGB200_ST_TOO_SHORT;
  if (st != DAGB200_ST_OK) {
    if (alpha) fill_rows<T, THREADS>(alpha, lat);
    if (status && threadIdx.x == 0) status[b] = st;
    return;
  }
  const T *m = match + b * lat;
  const T *E = links + (int64_t)b * L * Tl;
  T *prev = SmemRows<T>::get();
  T *cur = prev + L;
  const int wmask = (1 << wbits) - 1, wshift = 32 - wbits;

  for (int j = threadIdx.x; j < L; j += THREADS) {
    T v = (j == 0) ? m[0] : ninf;
    prev[j] = v;
    if (alpha) alpha[j] = v;
  }
  __syncthreads();

  for (int t = 1; t < Tn; t++) {
    const T *mrow = m + (int64_t)t * L;
    uint16_t *trow = tr + (int64_t)t * L;
    for (int j0 = threadIdx.x - lane; j0 < L; j0 += THREADS) {
      const int j = j0 + lane;
      T val = ninf;
      int dbest = 0;
      const int jhi = min(j0 + 31, O - 1);
      const int i_begin = max(t - 1, j0 - Tl);
      if (jhi >= t && j0 < O) {
        const bool jvalid = (j >= t) && (j < O);
        const int ilo = max(t - 1, j - Tl);
        Best<T> b0, b1, b2, b3;
        b0.init(); b1.init(); b2.init(); b3.init();
        const T *e = E + (int64_t)(j - 1);
        const int64_t es = (int64_t)Tl - 1;
        int i = i_begin;
        for (; i + 4 <= jhi; i += 4) {
          T x0 = ninf, x1 = ninf, x2 = ninf, x3 = ninf;
          if (jvalid && i + 0 >= ilo && i + 0 < j) x0 = prev[i + 0] + __ldg(e + (i + 0) * es);
          if (jvalid && i + 1 >= ilo && i + 1 < j) x1 = prev[i + 1] + __ldg(e + (i + 1) * es);
          if (jvalid && i + 2 >= ilo && i + 2 < j) x2 = prev[i + 2] + __ldg(e + (i + 2) * es);
          if (jvalid && i + 3 >= ilo && i + 3 < j) x3 = prev[i + 3] + __ldg(e + (i + 3) * es);
          b0.add(x0, tie_key(j - i - 0, wmask, wshift));
          b1.add(x1, tie_key(j - i - 1, wmask, wshift));
          b2.add(x2, tie_key(j - i - 2, wmask, wshift));
          b3.add(x3, tie_key(j - i - 3, wmask, wshift));
        }
        for (; i < jhi; i++) {
          T x0 = ninf;
          if (jvalid && i >= ilo && i < j) x0 = prev[i] + __ldg(e + i * es);
          b0.add(x0, tie_key(j - i, wmask, wshift));
        }
        if (jvalid) {
          b0.add(b1.v, b1.key); b2.add(b3.v, b3.key); b0.add(b2.v, b2.key);
          // -inf candidates never win (the reference starts from maxval=-inf, maxidx=-1 with strict '>')
          if (b0.v > ninf) dbest = b0.key & 0xffff;
          val = b0.v + mrow[j];
        }
      }
      if (j < L) {
        cur[j] = val;
        trow[j] = (uint16_t)dbest;
        if (alpha) alpha[(int64_t)t * L + j] = val;
      }
    }
    __syncthreads();
    T *tmp = prev; prev = cur; cur = tmp;
  }
  if (alpha) fill_rows<T, THREADS>(alpha + (int64_t)Tn * L, (int64_t)(M - Tn) * L);

  // backtrace (dag_best_alignment.cu:178-184); trace rows were written by this CTA (barrier above)
  if (threadIdx.x == 0) {
    int code = DAGB200_ST_OK;
    if (!(prev[O - 1] > ninf)) {
      code = DAGB200_ST_NO_PATH;
    } else {
      int pos = O - 1;
      for (int i = Tn - 1; i >= 0; i--) {
        prow[pos] = i;
        if (i == 0) break;
        const int d = tr[(int64_t)i * L + pos];
        if (d == 0) { code = DAGB200_ST_NO_PATH; break; }
        pos -= d;
      }
    }
    if (status) status[b] = code;
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
static int launch_alpha_beta(const T *match, const T *links, const int64_t *olen, const int64_t *tlen, T *alpha,
                             T *beta, int B, int M, int L, int Tl, bool grad, int32_t *status, cudaStream_t st) {
  const size_t smem = 2 * (size_t)L * sizeof(T);
  dim3 grid(B, grad ? 2 : 1);
  prof_mark(1, st);
#define LAUNCH_AB(TH)                                                                                          \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      cudaFuncSetAttribute(dag_alpha_beta_kernel<T, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    dag_alpha_beta_kernel<T, TH><<<grid, TH, smem, st>>>(match, links, olen, tlen, alpha, beta, M, L, Tl, status); \
  } while (0)
  if (L <= 256) LAUNCH_AB(256);
  else if (L <= 512) LAUNCH_AB(512);
  else LAUNCH_AB(1024);
#undef LAUNCH_AB
  DAGB200_CHECK_LAUNCH("dag_alpha_beta_kernel");
  prof_mark(2, st);
  return 0;
}

template <typename T>
static int launch_viterbi(const T *match, const T *links, const int64_t *olen, const int64_t *tlen, T *alpha,
                          uint16_t *trace, int32_t *path, int B, int M, int L, int Tl, int wbits, int32_t *status,
                          cudaStream_t st) {
  const size_t smem = 2 * (size_t)L * sizeof(T);
  prof_mark(6, st);
#define LAUNCH_V(TH)                                                                                            \
  do {                                                                                                          \
    if (smem > 48 * 1024)                                                                                       \
      cudaFuncSetAttribute(dag_viterbi_kernel<T, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
    dag_viterbi_kernel<T, TH><<<B, TH, smem, st>>>(match, links, olen, tlen, alpha, trace, path, M, L, Tl, wbits, status); \
  } while (0)
  if (L <= 256) LAUNCH_V(256);
  else if (L <= 512) LAUNCH_V(512);
  else LAUNCH_V(1024);
#undef LAUNCH_V
  DAGB200_CHECK_LAUNCH("dag_viterbi_kernel");
  prof_mark(7, st);
  return 0;
}

static int check_dp_args(const char *who, const void *match, const void *links, const int64_t *olen,
                         const int64_t *tlen, int dtype, int B, int M, int L, int Tl) {
  DAGB200_CHECK_ARG(B >= 0 && M >= 1 && L >= 1 && Tl >= 1, DAGB200_EINVAL, "%s: bad sizes B=%d M=%d L=%d T=%d", who, B, M, L, Tl);
  DAGB200_CHECK_ARG(B == 0 || (match && links && olen && tlen), DAGB200_EINVAL, "%s: null pointer", who);
  DAGB200_CHECK_ARG(dtype == DAGB200_F32 || dtype == DAGB200_F64, DAGB200_EDTYPE,
                    "%s: lattice dtype must be float32 or float64 (got %d)", who, dtype);
  DAGB200_CHECK_ARG((int64_t)L * Tl < (1ll << 31) && (int64_t)M * L < (1ll << 31) && L < 65536, DAGB200_ELIMIT,
                    "%s: lattice too large (L=%d T=%d M=%d)", who, L, Tl, M);
  DAGB200_CHECK_ARG(2 * (size_t)L * (dtype == DAGB200_F64 ? 8 : 4) <= 200 * 1024, DAGB200_ELIMIT,
                    "%s: L=%d exceeds the shared-memory row buffers", who, L);
  return 0;
}

}  // namespace dagb200

using namespace dagb200;

namespace dagb200 {
extern int g_exact_mode;
bool dp4_supported(int M, int L);
size_t dp4_workspace_bytes(int B, int M, int L);
int launch_alpha_beta_tcgen05(const float *match, const float *links, const int64_t *olen, const int64_t *tlen,
                              float *alpha, float *beta, int B, int M, int L, int Tl, bool grad, void *workspace,
                              int32_t *status, cudaStream_t st);
}  // namespace dagb200

extern "C" size_t dagb200_dag_loss_workspace_bytes(int B, int M, int L, int T) {
  (void)T;
  if (B <= 0 || M < 1 || L < 1 || !dp4_supported(M, L)) return 0;
  return dp4_workspace_bytes(B, M, L);
}

extern "C" int dagb200_dag_loss(const void *match, const void *links, const int64_t *output_length,
                                const int64_t *target_length, void *alpha, void *beta, int dtype, int B, int M,
                                int L, int T, int require_gradient, int config, void *workspace,
                                size_t workspace_bytes, int32_t *status, void *stream) {
  int rc = check_dp_args("dag_loss", match, links, output_length, target_length, dtype, B, M, L, T);
  if (rc) return rc;
  DAGB200_CHECK_ARG(config >= 1 && config <= 4, DAGB200_EINVAL, "config should be 1~4");
  if (B == 0) return 0;
  DAGB200_CHECK_ARG(alpha && beta, DAGB200_EINVAL, "dag_loss: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = require_gradient != 0;
  const size_t esz = dtype == DAGB200_F64 ? 8 : 4;
  if (!grad) {
    // the reference returns its at::zeros beta untouched when no gradient is required (dag_loss.cu:340,355)
    cudaError_t e = cudaMemsetAsync(beta, 0, (size_t)B * M * L * esz, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(beta)");
  }
  // fp32 with a workspace: blocked recurrences, far sums on tcgen05 (dag_dp4.cu).  Everything else -- fp64, exact
  // mode, no workspace, lattices beyond the shared memory of the blocked kernel (L > ~1760) -- runs the exact
  // log-domain kernels below.
  if (dtype == DAGB200_F32 && !g_exact_mode && workspace && dp4_supported(M, L) &&
      workspace_bytes >= dp4_workspace_bytes(B, M, L))
    return launch_alpha_beta_tcgen05((const float *)match, (const float *)links, output_length, target_length,
                                     (float *)alpha, (float *)beta, B, M, L, T, grad, workspace, status, st);
  if (dtype == DAGB200_F32)
    return launch_alpha_beta<float>((const float *)match, (const float *)links, output_length, target_length,
                                    (float *)alpha, (float *)beta, B, M, L, T, grad, status, st);
  return launch_alpha_beta<double>((const double *)match, (const double *)links, output_length, target_length,
                                   (double *)alpha, (double *)beta, B, M, L, T, grad, status, st);
}

namespace dagb200 {
size_t vit3_hand_bytes(int B, int L);
int launch_viterbi_wave(const float *match, const float *links, const int64_t *olen, const int64_t *tlen, float *lattice,
                        int32_t *path, float *hand_g, int full_lattice, int B, int M, int L, int Tl, int32_t *status,
                        cudaStream_t st);
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}  // namespace dagb200

// workspace: [trace, uint16 per cell (exact kernels)] [lattice plane, for callers that do not want alpha back]
//            [hand-down rows of the wave kernel]
extern "C" size_t dagb200_best_alignment_workspace_bytes(int B, int M, int L, int T) {
  (void)T;
  if (B <= 0 || M < 1 || L < 1) return 0;
  const size_t cells = (size_t)B * M * L;
  return align256(cells * sizeof(uint16_t)) + align256(cells * sizeof(float)) + vit3_hand_bytes(B, L);
}

extern "C" int dagb200_dag_best_alignment(const void *match, const void *links, const int64_t *output_length,
                                          const int64_t *target_length, void *alpha, int32_t *path, int dtype, int B,
                                          int M, int L, int T, int config, void *workspace, size_t workspace_bytes,
                                          int32_t *status, void *stream) {
  int rc = check_dp_args("dag_best_alignment", match, links, output_length, target_length, dtype, B, M, L, T);
  if (rc) return rc;
  DAGB200_CHECK_ARG(config >= 1 && config <= 4, DAGB200_EINVAL, "config should be 1~4");
  if (B == 0) return 0;
  DAGB200_CHECK_ARG(path, DAGB200_EINVAL, "dag_best_alignment: null path");
  DAGB200_CHECK_ARG(workspace && workspace_bytes >= dagb200_best_alignment_workspace_bytes(B, M, L, T),
                    DAGB200_EWORKSPACE, "dag_best_alignment: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int wbits = config + 1;  // config 1..4 -> TRANS_BLOCK_SIZE 4/8/16/32 (dag_best_alignment.cu:243-246)
  const size_t cells = (size_t)B * M * L;
  unsigned char *ws = (unsigned char *)workspace;
  if (dtype == DAGB200_F32 && config == 1 && !g_exact_mode && 2 * (int64_t)B <= 2147483647) {
    // wave-pipelined blocked kernel (dag_viterbi3.cu); without an alpha output only the cells that can reach the end
    // cell are computed, in a scratch lattice
    float *lattice = alpha ? (float *)alpha : reinterpret_cast<float *>(ws + align256(cells * sizeof(uint16_t)));
    float *hand = reinterpret_cast<float *>(ws + align256(cells * sizeof(uint16_t)) + align256(cells * sizeof(float)));
    return launch_viterbi_wave((const float *)match, (const float *)links, output_length, target_length, lattice, path,
                               hand, alpha ? 1 : 0, B, M, L, T, status, st);
  }
  if (dtype == DAGB200_F32)
    return launch_viterbi<float>((const float *)match, (const float *)links, output_length, target_length,
                                 (float *)alpha, (uint16_t *)workspace, path, B, M, L, T, wbits, status, st);
  return launch_viterbi<double>((const double *)match, (const double *)links, output_length, target_length,
                                (double *)alpha, (uint16_t *)workspace, path, B, M, L, T, wbits, status, st);
}
