// dag_dp2.cu -- blocked forward (alpha) / backward (beta) recurrences of the DAG loss for sm_100a.
//
// Replaces calculate_alpha_kernel / calculate_beta_kernel (reference dag_loss.cu:40-140, 178-274) on the fp32
// path.  Same recurrence, reorganised so that the transition plane is touched M/R times instead of M times and
// the predecessor sum of all "far" vertices runs on the tensor cores:
//
//   vertices in blocks of 32 (index q in sweep order), target rows in chunks of R steps; tile = (block, chunk)
//   far part   X[t, j in J] = sum_{i in earlier blocks} exp(a[t-1,i] - f[t,J]) * P'[i,j]
//              bf16 hi/lo split of both operands (3 mma.sync.m16n8k16 per product -> ~2^-16 relative error),
//              fp32 accumulation; f[t,J] = max of the far predecessors' log-values, so every operand is <= 1
//   near part  the 32x32 diagonal block: a 32-lane serial chain per tile, fp32, own frame g = max of the block's
//              previous row; the two parts are combined in the log domain:
//              a[t,j] = match[t,j] + L + log(X e^{f-L} + S e^{g-L}),  L = max(f, g)
//   schedule   tile (q, c) depends on (q' < q, c) and (q, c-1): anti-diagonal waves inside ONE CTA per
//              (utterance, direction), __syncthreads between the GEMM phase and the chain phase of a wave.
//              No inter-CTA communication (the reference spin-waits between CTAs on a global queue).
//
// Flush-to-zero contract (DESIGN.md "Numerics"): a far predecessor whose log-value is > 87 nats below the best far
// predecessor of the same 32-vertex block, or an in-block predecessor > 87 nats below the block's best, contributes
// exactly 0 (the log-domain reference would carry it at < e^-87 relative weight).
#include "common.cuh"
#include "dag_tiles.cuh"

namespace dagb200 {

constexpr int kDp2Threads = 512;
constexpr int kDp2Warps = kDp2Threads / 32;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// order-preserving float <-> int map so that the warp maximum is ONE redux.sync
__device__ __forceinline__ int f2ord(float x) {
  int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_max_redux(float x) { return ord2f(__reduce_max_sync(0xffffffffu, f2ord(x))); }

__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t *>(&h);
  lo = *reinterpret_cast<uint32_t *>(&l);
}

struct Dp2Smem {
  float *xbuf;   // [TPW][R][32]   far sums of the tiles of the current batch
  float *fbuf;   // [TPW][R]       their frames
  float *rmtab;  // [M][NB]        per (row, block q) maximum of the (shifted) log-values
  float *abuf;   // [warps][32]    chain broadcast buffer
  float *rmax;   // [NB*32]        per-source-vertex transition maximum
};

// One direction of one utterance.  R = rows per chunk (multiple of 16).
template <bool BETA, int R>
__device__ void blocked_chain(const float *__restrict__ match, float *__restrict__ lat, const unsigned char *__restrict__ ws,
                              const TileLayout &lay, const Dp2Smem &sm, int O, int Tn, int M, int L, int Tl) {
  constexpr int WPT = R / 16;             // warps per tile in the GEMM phase
  constexpr int TPW = kDp2Warps / WPT;    // tiles per batch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const float ninf = neg_inf_f();
  const int NB = lay.NB;
  const int NBv = (O + kBlk - 1) / kBlk;
  const int nsteps = Tn - 1;
  const int NCv = (nsteps + R - 1) / R;
  const int band = band_blocks(Tl);
  const float *g_rmax = reinterpret_cast<const float *>(ws + lay.off_rmax);
  const float *diag = reinterpret_cast<const float *>(ws + (BETA ? lay.off_diagB : lay.off_diagA));
  const uint4 *tiles = reinterpret_cast<const uint4 *>(ws + (BETA ? lay.off_tilesB : lay.off_tilesA));

  // ---- prologue: -inf padding, the seed row, per-vertex maxima, row-maximum table ------------------------
  for (int x = threadIdx.x; x < NB * kBlk; x += kDp2Threads) sm.rmax[x] = (x < O) ? g_rmax[x] : ninf;
  for (int x = threadIdx.x; x < M * NB; x += kDp2Threads) sm.rmtab[x] = ninf;
  {
    // rows >= Tn entirely, and columns beyond the last valid block of rows < Tn
    const int64_t tail0 = (int64_t)Tn * L;
    for (int64_t x = tail0 + threadIdx.x; x < (int64_t)M * L; x += kDp2Threads) lat[x] = ninf;
    const int c0 = NBv * kBlk;
    if (c0 < L) {
      const int wcols = L - c0;
      for (int x = threadIdx.x; x < Tn * wcols; x += kDp2Threads) lat[(int64_t)(x / wcols) * L + c0 + x % wcols] = ninf;
    }
    const int seed_row = BETA ? Tn - 1 : 0, seed_col = BETA ? O - 1 : 0;
    float *row = lat + (int64_t)seed_row * L;
    for (int j = threadIdx.x; j < min(L, c0); j += kDp2Threads) row[j] = (j == seed_col) ? match[(int64_t)seed_row * L + j] : ninf;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int seed_row = BETA ? Tn - 1 : 0, seed_col = BETA ? O - 1 : 0;
    const int Jb = seed_col / kBlk;
    const int q = BETA ? NBv - 1 - Jb : Jb;
    float v = match[(int64_t)seed_row * L + seed_col];
    if (!BETA) v += sm.rmax[seed_col];
    sm.rmtab[seed_row * NB + q] = v;
  }
  __syncthreads();

  // ---- anti-diagonal waves ----------------------------------------------------------------------------
  const int nwaves = NBv + NCv - 1;
  for (int w = 0; w < nwaves; w++) {
    const int c_lo = max(0, w - NBv + 1), c_hi = min(NCv - 1, w);
    for (int cb = c_lo; cb <= c_hi; cb += TPW) {
      // ================= GEMM phase: far predecessors through the tensor cores =================
      {
        const int ts = warp / WPT, sl = warp % WPT;
        const int c = cb + ts;
        if (c <= c_hi) {
          const int q = w - c;
          const int J = BETA ? NBv - 1 - q : q;
          const int r0 = 16 * sl + gid, r1 = r0 + 8;
          const int s0 = c * R + r0, s1 = c * R + r1;
          const bool v0 = s0 < nsteps, v1 = s1 < nsteps;
          const int tp0 = BETA ? Tn - 1 - s0 : s0, tp1 = BETA ? Tn - 1 - s1 : s1;   // previous-row index
          const int qlo = max(0, q - band);
          float f0 = ninf, f1 = ninf;
          if (v0) for (int qq = qlo; qq < q; qq++) f0 = fmaxf(f0, sm.rmtab[tp0 * NB + qq]);
          if (v1) for (int qq = qlo; qq < q; qq++) f1 = fmaxf(f1, sm.rmtab[tp1 * NB + qq]);
          float acc[4][4];
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[a][e] = 0.f;
          // any row of this 16-row slice with a finite far frame?
          const bool live = __any_sync(0xffffffffu, (v0 && f0 > ninf) || (v1 && f1 > ninf));
          if (live) {
            const float fs0 = (f0 > ninf) ? f0 * kLog2e : 0.f, fs1 = (f1 > ninf) ? f1 * kLog2e : 0.f;
            const float *row0 = lat + (int64_t)(v0 ? tp0 : 0) * L, *row1 = lat + (int64_t)(v1 ? tp1 : 0) * L;
            for (int qq = qlo; qq < q; qq++) {
              const int Js = BETA ? NBv - 1 - qq : qq;
              // skip source blocks whose previous-row maxima are all -inf for this slice
              const float m0 = v0 ? sm.rmtab[tp0 * NB + qq] : ninf, m1 = v1 ? sm.rmtab[tp1 * NB + qq] : ninf;
              if (!__any_sync(0xffffffffu, m0 > ninf || m1 > ninf)) continue;
              const uint4 *tp = tiles + (BETA ? lay.idxB(J, Js) : lay.idxA(Js, J)) * (kTileBytes / 16);
              uint4 u[8];
#pragma unroll
              for (int k8 = 0; k8 < 8; k8++) u[k8] = __ldg(tp + k8 * 32 + lane);
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                const int col = kBlk * Js + 16 * ks + 2 * tig;
                float x[8];
                // rows r0 / r1, columns col, col+1, col+8, col+9
                x[0] = (v0 && col < L) ? row0[col] : ninf;         x[1] = (v0 && col + 1 < L) ? row0[col + 1] : ninf;
                x[2] = (v1 && col < L) ? row1[col] : ninf;         x[3] = (v1 && col + 1 < L) ? row1[col + 1] : ninf;
                x[4] = (v0 && col + 8 < L) ? row0[col + 8] : ninf; x[5] = (v0 && col + 9 < L) ? row0[col + 9] : ninf;
                x[6] = (v1 && col + 8 < L) ? row1[col + 8] : ninf; x[7] = (v1 && col + 9 < L) ? row1[col + 9] : ninf;
                if (!BETA) {
                  const float ra = sm.rmax[col], rb = sm.rmax[col + 1], rc = sm.rmax[col + 8], rd = sm.rmax[col + 9];
                  x[0] += ra; x[1] += rb; x[2] += ra; x[3] += rb; x[4] += rc; x[5] += rd; x[6] += rc; x[7] += rd;
                }
                x[0] = exp2f(fmaf(x[0], kLog2e, -fs0)); x[1] = exp2f(fmaf(x[1], kLog2e, -fs0));
                x[2] = exp2f(fmaf(x[2], kLog2e, -fs1)); x[3] = exp2f(fmaf(x[3], kLog2e, -fs1));
                x[4] = exp2f(fmaf(x[4], kLog2e, -fs0)); x[5] = exp2f(fmaf(x[5], kLog2e, -fs0));
                x[6] = exp2f(fmaf(x[6], kLog2e, -fs1)); x[7] = exp2f(fmaf(x[7], kLog2e, -fs1));
                uint32_t ahi[4], alo[4];
                split_bf16x2(x[0], x[1], ahi[0], alo[0]);
                split_bf16x2(x[2], x[3], ahi[1], alo[1]);
                split_bf16x2(x[4], x[5], ahi[2], alo[2]);
                split_bf16x2(x[6], x[7], ahi[3], alo[3]);
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                  const uint4 &h = u[2 * ks + (nt >> 1)], &l = u[4 + 2 * ks + (nt >> 1)];
                  const uint32_t bh0 = (nt & 1) ? h.z : h.x, bh1 = (nt & 1) ? h.w : h.y;
                  const uint32_t bl0 = (nt & 1) ? l.z : l.x, bl1 = (nt & 1) ? l.w : l.y;
                  mma_bf16_16816(acc[nt], ahi, bh0, bh1);
                  mma_bf16_16816(acc[nt], alo, bh0, bh1);
                  mma_bf16_16816(acc[nt], ahi, bl0, bl1);
                }
              }
            }
          }
          float *xb = sm.xbuf + (size_t)ts * R * kBlk;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            *reinterpret_cast<float2 *>(xb + r0 * kBlk + 8 * nt + 2 * tig) = make_float2(acc[nt][0], acc[nt][1]);
            *reinterpret_cast<float2 *>(xb + r1 * kBlk + 8 * nt + 2 * tig) = make_float2(acc[nt][2], acc[nt][3]);
          }
          if (tig == 0) { sm.fbuf[ts * R + r0] = f0; sm.fbuf[ts * R + r1] = f1; }
        }
      }
      __syncthreads();
      // ================= chain phase: the 32x32 diagonal block, one warp per tile =================
      {
        const int ts = warp / WPT;
        const int c = cb + ts;
        const int cw = (WPT >= 4) ? (ts % WPT) : (WPT == 2 ? ((ts >> 1) & 1) : 0);  // spread chain warps over SMSPs
        if (c <= c_hi && (warp % WPT) == cw) {
          const int q = w - c;
          const int J = BETA ? NBv - 1 - q : q;
          const int j = kBlk * J + lane;
          const bool jin = j < L;
          float pd[kBlk];
          {
            const float *d = diag + (size_t)J * kBlk * kBlk + lane;
#pragma unroll
            for (int ii = 0; ii < kBlk; ii++) pd[ii] = __ldg(d + ii * kBlk);
          }
          const float rmj = sm.rmax[kBlk * J + lane];
          const int sbeg = c * R, send = min(nsteps, sbeg + R);
          const int tpf = BETA ? Tn - 1 - sbeg : sbeg;
          float prev = jin ? lat[(int64_t)tpf * L + j] : ninf;
          if (!BETA) prev += rmj;
          float *ab = sm.abuf + warp * kBlk;
          const float *xb = sm.xbuf + (size_t)ts * R * kBlk + lane;
          const float *fb = sm.fbuf + ts * R;
          // emission prefetch ring (4 steps ahead)
          float mring[4];
#pragma unroll
          for (int p = 0; p < 4; p++) {
            const int s = sbeg + p;
            const int t = BETA ? Tn - 2 - s : 1 + s;
            mring[p] = (s < send && jin) ? __ldg(match + (int64_t)t * L + j) : ninf;
          }
          float g = warp_max_redux(prev);
          for (int sb = sbeg; sb < send; sb += 4) {
#pragma unroll
            for (int p = 0; p < 4; p++) {
              const int s = sb + p;
              if (s < send) {
                const int t = BETA ? Tn - 2 - s : 1 + s;
                const float mt = mring[p];
                {
                  const int s4 = s + 4;
                  const int t4 = BETA ? Tn - 2 - s4 : 1 + s4;
                  mring[p] = (s4 < send && jin) ? __ldg(match + (int64_t)t4 * L + j) : ninf;
                }
                const float gs = (g > ninf) ? g : 0.f;
                const float ah = exp2f((prev - gs) * kLog2e);
                __syncwarp();
                ab[lane] = ah;
                __syncwarp();
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int i4 = 0; i4 < kBlk; i4 += 4) {
                  const float4 a4 = *reinterpret_cast<const float4 *>(ab + i4);
                  s0 = fmaf(a4.x, pd[i4 + 0], s0); s1 = fmaf(a4.y, pd[i4 + 1], s1);
                  s2 = fmaf(a4.z, pd[i4 + 2], s2); s3 = fmaf(a4.w, pd[i4 + 3], s3);
                }
                const float S = (s0 + s1) + (s2 + s3);
                const int r = s - sbeg;
                const float X = xb[r * kBlk];
                const float f = fb[r];
                const float Lm = fmaxf(f, g);
                const float Ls = (Lm > ninf) ? Lm : 0.f;
                const float ef = (f > ninf) ? exp2f((f - Ls) * kLog2e) : 0.f;
                const float eg = (g > ninf) ? exp2f((gs - Ls) * kLog2e) : 0.f;
                const float tot = fmaf(X, ef, S * eg);
                const bool valid = jin && j >= t && j < O;
                float out = ninf;
                if (valid && tot > 0.f) out = (BETA ? mt + rmj : mt) + (Ls + __logf(tot));
                if (jin) lat[(int64_t)t * L + j] = out;
                prev = BETA ? out : out + rmj;
                g = warp_max_redux(prev);
                if (lane == 0) sm.rmtab[t * NB + q] = g;
              }
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

template <int R>
__global__ void __launch_bounds__(kDp2Threads, 1)
dag_alpha_beta_blocked_kernel(const float *__restrict__ match, const int64_t *__restrict__ olen,
                              const int64_t *__restrict__ tlen, float *__restrict__ alpha, float *__restrict__ beta,
                              const unsigned char *__restrict__ ws, int M, int L, int Tl, TileLayout lay,
                              int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char dp2_smem[];
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t latsz = (int64_t)M * L;
  float *dst = (is_beta ? beta : alpha) + b * latsz;
  int st = DAGB200_ST_OK;
  if (Tn < 2 || O < 2) st = DAGB200_ST_LEN_LT2;
  else if (O < Tn || O > L || Tn > M) st = DAGB200_ST_GRAPH_SMALL;
  if (st != DAGB200_ST_OK) {
    for (int64_t x = threadIdx.x; x < latsz; x += kDp2Threads) dst[x] = neg_inf_f();
    if (status && threadIdx.x == 0 && !is_beta) status[b] = st;
    return;
  }
  if (status && threadIdx.x == 0 && !is_beta) status[b] = DAGB200_ST_OK;
  constexpr int TPW = kDp2Warps / (R / 16);
  Dp2Smem sm;
  float *p = reinterpret_cast<float *>(dp2_smem);
  sm.xbuf = p;  p += TPW * R * kBlk;
  sm.fbuf = p;  p += TPW * R;
  sm.abuf = p;  p += kDp2Warps * kBlk;
  sm.rmax = p;  p += lay.NB * kBlk;
  sm.rmtab = p;
  const float *m = match + b * latsz;
  const unsigned char *wsb = ws + (size_t)b * lay.sample_bytes;
  if (is_beta) blocked_chain<true, R>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl);
  else blocked_chain<false, R>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl);
}

size_t dp2_smem_bytes(int R, int M, int L) {
  TileLayout lay = TileLayout::make(L);
  const int TPW = kDp2Warps / (R / 16);
  return sizeof(float) * ((size_t)TPW * R * kBlk + (size_t)TPW * R + kDp2Warps * kBlk + (size_t)lay.NB * kBlk + (size_t)M * lay.NB);
}

int launch_dag_prep(const float *links, const int64_t *olen, void *workspace, int B, int L, int Tl, cudaStream_t st);

size_t dp2_workspace_bytes(int B, int L) { return TileLayout::make(L).sample_bytes * (size_t)B; }

bool dp2_supported(int M, int L) { return dp2_smem_bytes(64, M, L) <= 200 * 1024 && L >= 1; }

int launch_alpha_beta_blocked(const float *match, const float *links, const int64_t *olen, const int64_t *tlen,
                              float *alpha, float *beta, int B, int M, int L, int Tl, bool grad, void *workspace,
                              int32_t *status, cudaStream_t st) {
  int rc = launch_dag_prep(links, olen, workspace, B, L, Tl, st);
  if (rc) return rc;
  TileLayout lay = TileLayout::make(L);
  dim3 grid(B, grad ? 2 : 1);
  // rows per chunk: long targets amortise the transition tiles over 64 rows, short ones keep more waves
  const int R = (M > 96) ? 64 : (M > 40 ? 32 : 16);
  const size_t smem = dp2_smem_bytes(R, M, L);
#define LAUNCH_DP2(RR)                                                                                          \
  do {                                                                                                          \
    cudaFuncSetAttribute(dag_alpha_beta_blocked_kernel<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    dag_alpha_beta_blocked_kernel<RR><<<grid, kDp2Threads, smem, st>>>(match, olen, tlen, alpha, beta,          \
                                                                      (const unsigned char *)workspace, M, L, Tl, lay, status); \
  } while (0)
  if (R == 64) LAUNCH_DP2(64);
  else if (R == 32) LAUNCH_DP2(32);
  else LAUNCH_DP2(16);
#undef LAUNCH_DP2
  DAGB200_CHECK_LAUNCH("dag_alpha_beta_blocked_kernel");
  return 0;
}

}  // namespace dagb200
