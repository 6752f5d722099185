// dag_dp2.cu -- blocked forward (alpha) / backward (beta) recurrences of the DAG loss for sm_100a.
//
// Replaces calculate_alpha_kernel / calculate_beta_kernel (reference dag_loss.cu:40-140, 178-274) on the fp32
// path.  Same recurrence, reorganised so that the transition plane is touched M/32 times instead of M times,
// the predecessor sum of all "far" vertices runs on the tensor cores, and the serial depth of a tile is its 32
// COLUMNS instead of its rows:
//
//   vertices in blocks of 32 (index q in sweep order), target rows in chunks of 32 steps; tile = (block, chunk)
//   far part   X[t, j in J] = sum_{i in earlier blocks} exp2(a2[t-1,i] - FI[t,J]) * P'[i,j]
//              a2 = log2 of the predecessor's outgoing mass, FI = integer upper bound of the far predecessors'
//              a2 (every operand <= 1, the scaling an exact power of two); both operands are split into bf16
//              hi/lo (3 mma.sync.m16n8k16 per product, ~2^-16 relative error), fp32 accumulation
//   near part  the 32x32 diagonal block.  Cell (t, j) only depends on cells with smaller t AND smaller j, so a
//              warp maps LANES TO ROWS and sweeps the 32 columns serially: at column j every lane r forms the
//              in-block predecessor sum of its own row for the NEXT row (FMAs on registers), hands it to lane
//              r+1 with one shuffle, and combines what it received with its far sum and emission.  All values
//              are block floating point -- a mantissa per cell, an integer exponent per (row, 8-column group),
//              all lane-private: no cross-lane reduction and nothing transcendental on the dependency path.
//   schedule   tile (q, c) depends on (q' < q, c) and (q, c-1): anti-diagonal waves inside ONE CTA per
//              (utterance, direction), __syncthreads between the GEMM phase and the chain phase of a wave.
//              No inter-CTA communication (the reference spin-waits between CTAs on a global queue).
//
// Flush-to-zero contract (DESIGN.md "Numerics"): a predecessor contributes exactly 0 when its mass is more than
// 2^-126 below (far part) the largest far predecessor of the 32-vertex destination block / (near part) the largest
// in-block predecessor group of the cell; a transition when it is > 87 nats below the best transition of its
// source vertex; an emission when it is > 87 nats below the best emission of its (row, 8-column group).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "dag_tiles.cuh"

namespace dagb200 {

constexpr int kDp2Threads = 512;
constexpr int kDp2Warps = kDp2Threads / 32;
constexpr int kRows = 32;                       // target rows per chunk (= lanes of a chain warp)
constexpr int kWpt = kRows / 16;                // warps per tile in the GEMM phase
constexpr int kTpw = kDp2Warps / kWpt;          // tiles per batch
constexpr int kPitch = 33;                      // padded row pitch of 32x32 fp32 tiles in shared memory
constexpr int kNegBig = -(1 << 20);             // "empty" integer frame
constexpr float kLn2Hi = 0.693359375f;          // 355/512, 9 significant bits: k * kLn2Hi is exact for |k| < 2^15
constexpr float kLn2Lo = -2.12194440e-4f;       // ln2 - kLn2Hi

__device__ long long g_dp2_dbg[8];

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 2^d for integer d <= 127 (0 below the normal range)
__device__ __forceinline__ float pow2i(int d) { return d < -126 ? 0.f : __int_as_float((d + 127) << 23); }
// unbiased exponent of a positive normal float
__device__ __forceinline__ int fexp(float v) { return (int)((__float_as_uint(v) >> 23) & 0xff) - 127; }

__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t *>(&h);
  lo = *reinterpret_cast<uint32_t *>(&l);
}

__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gsrc) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- TMA bulk copy (global -> shared) completing on an mbarrier -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void group_barrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Fragment tile of the A operand (32 rows x 32 K, bf16 hi/lo), unit = slice*4 + ks*2 + hl, 32 uint4 per unit:
// lane (gid, tig) holds a0 = (row gid, k 2tig..+1), a1 = (row gid+8, same k), a2 = (row gid, k+8..+9), a3 = (gid+8, k+8..).
__device__ __forceinline__ void write_frag_elem(uint4 *tile, int row, int k, float v) {
  const int slice = row >> 4, r16 = row & 15, ks = k >> 4, k16 = k & 15;
  const int gid = r16 & 7, reg = (r16 >> 3) | ((k16 >> 3) << 1), tig = (k16 & 7) >> 1, half = k16 & 1;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  __nv_bfloat16 *ph = reinterpret_cast<__nv_bfloat16 *>(tile + (slice * 4 + ks * 2 + 0) * 32 + gid * 4 + tig);
  __nv_bfloat16 *pl = reinterpret_cast<__nv_bfloat16 *>(tile + (slice * 4 + ks * 2 + 1) * 32 + gid * 4 + tig);
  ph[reg * 2 + half] = h;
  pl[reg * 2 + half] = l;
}

struct Dp2Smem {
  float *xbuf;   // [kTpw][32][kPitch] far sums of the tiles of the current batch
  int *fbuf;     // [kTpw][32]         their integer frames
  float *ut;     // [kTpw][32][32]     diagonal-block weights of the tiles of the current batch ([cj][ci])
  float *io;     // [kTpw][32][kPitch] emissions in, lattice values out
  int *rmtab;    // [M][NB]            per (row, block q) integer upper bound of log2(outgoing mass)
  float *rmax;   // [NB*32]            per-source-vertex transition maximum
  float *stm;    // [NB][32]           chain state carried between chunks: mantissas of the chunk's last row
  int *stf;      // [NB]               ... and their common frame
  uint4 *tstage; // [kTpw][2][256]     transition tiles in flight (TMA destination, double buffered per tile group)
  uint64_t *mbar;// [kTpw][2]          their completion barriers
};

// ---------------------------------------------------------------------------------------------------------
// Diagonal-block sweep, lanes = rows.  The 32 columns are processed in four groups of 8 (runtime loop, so the
// code stays small enough for the instruction cache); the mantissas of the current group live in registers,
// those of the completed groups in my row of the far-sum tile (shared memory: a cell's slot holds its far sum
// until the cell is computed, its mantissa afterwards) under ONE common frame CF.
struct ChainRow {
  int CF;      // frame of the completed groups (kNegBig: nothing yet)
  int maxe;    // integer upper bound of log2(outgoing mass) over my row
};

// K = column inside the group (compile time), G = group (runtime)
template <bool BETA, int K>
__device__ __forceinline__ void chain_column(float (&cur)[8], int &gexpG, ChainRow &st, int G, const float *utw,
                                             float *mrow, float *iow, int lane, int FI, float d0m, int d0f,
                                             const float (&ew)[8], const float (&mm)[8], int KE, bool rowvalid,
                                             int jbase, int t, int O, const float *rmax_blk) {
  const int cj = 8 * G + K;
  const float *urow = utw + cj * kBlk;                 // U[ci][cj] for ci = 0..31 (0 for ci >= cj)
  // (1) predecessor sum of MY row for the next row's column cj
  float pdone = 0.f;                                   // completed groups, frame st.CF
  for (int g = 0; g < G; g++) {
    const float4 u0 = *reinterpret_cast<const float4 *>(urow + 8 * g);
    const float4 u1 = *reinterpret_cast<const float4 *>(urow + 8 * g + 4);
    // slot of sweep column ci in my row: its vertex offset (alpha: ci, beta: 31 - ci)
    const float *mr = BETA ? mrow + (kBlk - 1 - 8 * g) : mrow + 8 * g;
    constexpr int D = BETA ? -1 : 1;
    pdone = fmaf(mr[0 * D], u0.x, pdone); pdone = fmaf(mr[1 * D], u0.y, pdone);
    pdone = fmaf(mr[2 * D], u0.z, pdone); pdone = fmaf(mr[3 * D], u0.w, pdone);
    pdone = fmaf(mr[4 * D], u1.x, pdone); pdone = fmaf(mr[5 * D], u1.y, pdone);
    pdone = fmaf(mr[6 * D], u1.z, pdone); pdone = fmaf(mr[7 * D], u1.w, pdone);
  }
  float pcur = 0.f;                                    // current group, frame gexpG
  if (K > 0) {
    const float4 u0 = *reinterpret_cast<const float4 *>(urow + 8 * G);
    const float4 u1 = *reinterpret_cast<const float4 *>(urow + 8 * G + 4);
    const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
    for (int k = 0; k < K; k++) pcur = fmaf(cur[k], uu[k], pcur);
  }
  const int df = max(st.CF, (K > 0) ? gexpG : kNegBig);
  const float dm = fmaf(pdone, pow2i(st.CF - df), (K > 0) ? pcur * pow2i(gexpG - df) : 0.f);
  // (2) hand it to the next row; row 0 of the chunk takes the sum formed from the previous chunk's last row
  float rm = __shfl_up_sync(0xffffffffu, dm, 1);
  int rf = __shfl_up_sync(0xffffffffu, df, 1);
  const float zm = __shfl_sync(0xffffffffu, d0m, cj);
  const int zf = __shfl_sync(0xffffffffu, d0f, cj);
  if (lane == 0) { rm = zm; rf = zf; }
  // (3) combine with the far sum (frame FI) -> total incoming mass of the cell, frame Lm
  const int jj = BETA ? (kBlk - 1 - cj) : cj;
  const float X = mrow[jj];                            // far sum of this cell (its slot later takes the mantissa)
  const int Lm = max(FI, rf);
  const float tot = fmaf(X, pow2i(FI - Lm), rm * pow2i(rf - Lm));
  // (4) lattice value (off the dependency path)
  const int j = jbase + jj;
  const bool valid = rowvalid && j >= t && j < O;
  float out = neg_inf_f();
  if (valid && tot > 0.f) {
    const float fl = (float)Lm;
    out = (mm[K] + fmaf(__log2f(tot), 0.6931471805599453f, fl * kLn2Lo)) + fl * kLn2Hi;
    if (BETA) out += rmax_blk[jj];
  }
  iow[jj] = out;
  // (5) outgoing mass = tot * emission * best transition, as a mantissa in the current group's frame
  const float v = tot * ew[K];                         // frame Lm + KE ; ew = 0 for invalid cells
  float mnew = 0.f;
  if (v > 0.f) {
    const int fr = Lm + KE;
    const int e = fexp(v);
    const float vn = v * pow2i(-e);                    // normalised to [1, 2)
    st.maxe = max(st.maxe, fr + e + 1);
    if (gexpG == kNegBig) {
      gexpG = fr + e;
      mnew = vn;
    } else {
      const int shift = fr + e - gexpG;
      if (shift > 100) {                               // group frame far too low for this value: re-frame (rare)
#pragma unroll
        for (int k = 0; k < K; k++) cur[k] *= pow2i(-shift);
        gexpG += shift;
        mnew = vn;
      } else {
        mnew = vn * pow2i(shift);
      }
    }
  }
  cur[K] = mnew;
}

template <bool BETA>
__device__ __forceinline__ void chain_sweep(ChainRow &st, const float *utw, float *mrow, float *iow, int lane, int FI,
                                            float d0m, int d0f, bool rowvalid, int jbase, int t, int O,
                                            const float *rmax_blk) {
#pragma unroll 1
  for (int G = 0; G < 4; G++) {
    // emissions of this row for the 8 columns of the group: scaled exponentials and their integer frame
    float mm[8], ew[8], w2[8];
    float wmax = neg_inf_f();
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int cj = 8 * G + k;
      const int jj = BETA ? (kBlk - 1 - cj) : cj;
      const int j = jbase + jj;
      const bool valid = rowvalid && j >= t && j < O;
      mm[k] = iow[jj];
      w2[k] = valid ? (mm[k] + rmax_blk[jj]) * kLog2e : neg_inf_f();
      wmax = fmaxf(wmax, w2[k]);
    }
    const int KE = wmax > -1.0e30f ? (int)ceilf(wmax) : kNegBig;
#pragma unroll
    for (int k = 0; k < 8; k++) ew[k] = (KE > kNegBig) ? exp2f(w2[k] - (float)KE) : 0.f;
    float cur[8];
    int gexpG = kNegBig;
    chain_column<BETA, 0>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 1>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 2>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 3>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 4>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 5>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 6>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    chain_column<BETA, 7>(cur, gexpG, st, G, utw, mrow, iow, lane, FI, d0m, d0f, ew, mm, KE, rowvalid, jbase, t, O, rmax_blk);
    // retire the group: bring everything completed so far to the common frame max(CF, gexpG)
    const int nf = max(st.CF, gexpG);
    if (nf > st.CF && st.CF > kNegBig) {
      const float sc = pow2i(st.CF - nf);
      for (int ci = 0; ci < 8 * G; ci++) mrow[BETA ? kBlk - 1 - ci : ci] *= sc;
    }
    const float sg = (gexpG > kNegBig) ? pow2i(gexpG - nf) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) mrow[BETA ? kBlk - 1 - (8 * G + k) : 8 * G + k] = cur[k] * sg;
    st.CF = nf;
  }
}

// One direction of one utterance.
template <bool BETA>
__device__ void blocked_chain(const float *__restrict__ match, float *__restrict__ lat, unsigned char *__restrict__ ws,
                              const TileLayout &lay, const Dp2Smem &sm, int O, int Tn, int M, int L, int Tl, bool dbg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const float ninf = neg_inf_f();
  const int NB = lay.NB;
  const int NBv = (O + kBlk - 1) / kBlk;
  const int nsteps = Tn - 1;
  const int NCv = (nsteps + kRows - 1) / kRows;
  const int band = band_blocks(Tl);
  const float *g_rmax = reinterpret_cast<const float *>(ws + lay.off_rmax);
  const float *diag = reinterpret_cast<const float *>(ws + (BETA ? lay.off_diagB : lay.off_diagA));
  const uint4 *tiles = reinterpret_cast<const uint4 *>(ws + (BETA ? lay.off_tilesB : lay.off_tilesA));
  uint4 *afrag = reinterpret_cast<uint4 *>(ws + (BETA ? lay.off_afragB : lay.off_afragA));   // [chunk][q] 256 uint4

  // ---- prologue: -inf padding, the seed row, per-vertex maxima, frame table, chain state ------------------
  for (int x = threadIdx.x; x < NB * kBlk; x += kDp2Threads) {
    sm.rmax[x] = (x < O) ? g_rmax[x] : ninf;
    sm.stm[x] = 0.f;
  }
  for (int x = threadIdx.x; x < NB * 4; x += kDp2Threads) sm.stf[x] = kNegBig;
  for (int x = threadIdx.x; x < NBv * 256; x += kDp2Threads) afrag[x] = make_uint4(0u, 0u, 0u, 0u);   // chunk 0
  if (threadIdx.x < kTpw * 2) mbar_init(sm.mbar + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int x = threadIdx.x; x < M * NB; x += kDp2Threads) sm.rmtab[x] = kNegBig;
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
    // seed: outgoing mass of the single start cell, as mantissa * 2^F (sweep-order column index ci)
    const int seed_row = BETA ? Tn - 1 : 0, seed_col = BETA ? O - 1 : 0;
    const int Jb = seed_col / kBlk, jj = seed_col % kBlk;
    const int ci = BETA ? kBlk - 1 - jj : jj;
    const int q = BETA ? NBv - 1 - Jb : Jb;
    float v = match[(int64_t)seed_row * L + seed_col];
    if (!BETA) v += sm.rmax[seed_col];
    const float v2 = v * kLog2e;
    if (v2 > -1.0e30f) {
      const int F = (int)ceilf(v2);
      const float mant0 = exp2f(v2 - (float)F);  // in (0.5, 1]
      sm.stm[q * kBlk + ci] = mant0;
      sm.stf[q] = F;
      sm.rmtab[seed_row * NB + q] = F + 1;
      // row 0 of the chunk-0 fragment tile of block q: value mant0/2 in frame F+1, at K index = vertex offset jj
      write_frag_elem(afrag + (size_t)q * 256, 0, jj, 0.5f * mant0);
    }
  }
  __syncthreads();

  // ---- anti-diagonal waves ----------------------------------------------------------------------------
  uint32_t tuse = 0;   // tiles consumed so far by my tile group (stage = tuse & 1, mbarrier parity = (tuse >> 1) & 1)
  const int nwaves = NBv + NCv - 1;
  for (int w = 0; w < nwaves; w++) {
    const int c_lo = max(0, w - NBv + 1), c_hi = min(NCv - 1, w);
    for (int cb = c_lo; cb <= c_hi; cb += kTpw) {
      long long tdbg0 = dbg ? clock64() : 0;
      // ================= GEMM phase: far predecessors through the tensor cores =================
      {
        const int ts = warp / kWpt, sl = warp % kWpt;
        const int c = cb + ts;
        if (c <= c_hi) {
          const int q = w - c;
          const int J = BETA ? NBv - 1 - q : q;
          const int r0 = 16 * sl + gid, r1 = r0 + 8;
          const int s0 = c * kRows + r0, s1 = c * kRows + r1;
          const bool v0 = s0 < nsteps, v1 = s1 < nsteps;
          const int tp0 = BETA ? Tn - 1 - s0 : s0, tp1 = BETA ? Tn - 1 - s1 : s1;   // previous-row index
          const bool wd = dbg && !BETA && blockIdx.x == 0 && warp == 0 && lane == 0;
          long long tw0 = wd ? clock64() : 0;
          // stage this tile's diagonal-block weights and emissions for the chain phase: asynchronous
          // global->shared copies issued now, landed by the end of the MMA loop (2 warps x 16 rows)
          {
            const float *d = diag + (size_t)J * kBlk * kBlk;
            float *utw = sm.ut + (size_t)ts * kBlk * kBlk;
            float *iow = sm.io + (size_t)ts * kRows * kPitch;
            const int j = kBlk * J + lane;
            for (int rr = sl; rr < kBlk; rr += kWpt) {
              cp_async_f32(utw + rr * kBlk + lane, d + rr * kBlk + lane);
              const int s = c * kRows + rr;
              const int t = BETA ? Tn - 2 - s : 1 + s;
              if (s < nsteps && j < L) cp_async_f32(iow + rr * kPitch + lane, match + (int64_t)t * L + j);
              else iow[rr * kPitch + lane] = ninf;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
          }
          const int qlo = max(0, q - band);
          // far frames of the 32 rows of this tile and the set of source blocks holding any mass: lanes <-> source
          // blocks (two passes if the band spans more than 32 blocks), one independent smem read + redux per row
          unsigned long long livemask = 0ull;
          int F0 = kNegBig, F1 = kNegBig;
          {
            int *fbw = sm.fbuf + ts * kRows;
            const int nsrc = q - qlo;
            bool any_lo = false, any_hi = false;
            for (int rr = 0; rr < kRows; rr++) {
              const int sR = c * kRows + rr;
              const int tpR = BETA ? Tn - 1 - sR : sR;
              int v = kNegBig, v2 = kNegBig;
              if (sR < nsteps) {
                if (lane < nsrc) v = sm.rmtab[tpR * NB + qlo + lane];
                if (lane + 32 < nsrc) v2 = sm.rmtab[tpR * NB + qlo + lane + 32];
              }
              any_lo = any_lo || v > kNegBig;
              any_hi = any_hi || v2 > kNegBig;
              const int fr = __reduce_max_sync(0xffffffffu, max(v, v2));
              if (rr == r0) F0 = fr;
              if (rr == r1) F1 = fr;
              if (sl == 0 && lane == 0) fbw[rr] = fr;
            }
            livemask = (unsigned long long)__ballot_sync(0xffffffffu, any_lo) |
                       ((unsigned long long)__ballot_sync(0xffffffffu, any_hi) << 32);
          }
          float acc[4][4];
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[a][e] = 0.f;
          const bool live = livemask != 0ull;
          long long tw1 = wd ? clock64() : 0;
          if (live) {
            // A operand of one source block for my 16-row slice: cached fragments (hi ks0, lo ks0, hi ks1, lo ks1),
            // written by the chain warp that produced those rows; plain loads (same CTA, ordered by the barriers)
            auto load_frag = [&](int qn, uint4 (&f)[4]) {
              const uint4 *ft = afrag + ((size_t)c * NBv + qn) * 256 + (sl * 4) * 32 + lane;
              f[0] = ft[0]; f[1] = ft[32]; f[2] = ft[64]; f[3] = ft[96];
            };
            auto tile_ptr = [&](int Js) {
              return tiles + (BETA ? lay.idxB(J, Js) : lay.idxA(Js, J)) * (kTileBytes / 16);
            };
            uint4 *stg = sm.tstage + (size_t)ts * 2 * 256;
            uint64_t *mb = sm.mbar + ts * 2;
            const bool producer = (sl == 0) && (lane == 0);
            unsigned long long mk = livemask;
            // consume one source block: rescale the cached fragments from the source rows' frames to this tile's far
            // frames (exact powers of two), then 3 split products per n-tile against the staged transition tile
            auto consume = [&](int qs, const uint4 (&f)[4], int stage) {
              const uint4 *tsm = stg + stage * 256;
              const int Fs0 = v0 ? sm.rmtab[tp0 * NB + qs] : kNegBig, Fs1 = v1 ? sm.rmtab[tp1 * NB + qs] : kNegBig;
              const float s0 = (Fs0 > kNegBig) ? pow2i(Fs0 - F0) : 0.f, s1 = (Fs1 > kNegBig) ? pow2i(Fs1 - F1) : 0.f;
              const __nv_bfloat162 sc0 = __floats2bfloat162_rn(s0, s0), sc1 = __floats2bfloat162_rn(s1, s1);
              auto scale = [&](uint32_t v, const __nv_bfloat162 &sc) -> uint32_t {
                __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162 *>(&v), sc);
                return *reinterpret_cast<uint32_t *>(&r);
              };
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                const uint4 fh = f[2 * ks], fl = f[2 * ks + 1];
                uint32_t ahi[4], alo[4];
                ahi[0] = scale(fh.x, sc0); ahi[1] = scale(fh.y, sc1); ahi[2] = scale(fh.z, sc0); ahi[3] = scale(fh.w, sc1);
                alo[0] = scale(fl.x, sc0); alo[1] = scale(fl.y, sc1); alo[2] = scale(fl.z, sc0); alo[3] = scale(fl.w, sc1);
                const uint4 h0 = tsm[(2 * ks) * 32 + lane], h1 = tsm[(2 * ks + 1) * 32 + lane];
                const uint4 l0 = tsm[(4 + 2 * ks) * 32 + lane], l1 = tsm[(4 + 2 * ks + 1) * 32 + lane];
                mma_bf16_16816(acc[0], ahi, h0.x, h0.y); mma_bf16_16816(acc[1], ahi, h0.z, h0.w);
                mma_bf16_16816(acc[2], ahi, h1.x, h1.y); mma_bf16_16816(acc[3], ahi, h1.z, h1.w);
                mma_bf16_16816(acc[0], alo, h0.x, h0.y); mma_bf16_16816(acc[1], alo, h0.z, h0.w);
                mma_bf16_16816(acc[2], alo, h1.x, h1.y); mma_bf16_16816(acc[3], alo, h1.z, h1.w);
                mma_bf16_16816(acc[0], ahi, l0.x, l0.y); mma_bf16_16816(acc[1], ahi, l0.z, l0.w);
                mma_bf16_16816(acc[2], ahi, l1.x, l1.y); mma_bf16_16816(acc[3], ahi, l1.z, l1.w);
              }
            };
            auto next_block = [&]() -> int {   // pops the next live source block (sweep index), -1 when done
              if (!mk) return -1;
              const int r = qlo + __ffsll((long long)mk) - 1;
              mk &= mk - 1;
              return r;
            };
            auto issue = [&](int qn, uint4 (&f)[4], int stage) {   // TMA the tile, register-load the fragments
              const int Jn = BETA ? NBv - 1 - qn : qn;
              if (producer) { mbar_expect_tx(mb + stage, kTileBytes); bulk_g2s(stg + stage * 256, tile_ptr(Jn), kTileBytes, mb + stage); }
              load_frag(qn, f);
            };
            // software pipeline, unrolled by two so that the two fragment buffers ping-pong without register copies
            uint4 fa[4], fb4[4];
            int qa = next_block(), qb;
            issue(qa, fa, tuse & 1);
            while (true) {
              qb = next_block();
              if (qb >= 0) issue(qb, fb4, (tuse + 1) & 1);
              mbar_wait(mb + (tuse & 1), (tuse >> 1) & 1);
              consume(qa, fa, tuse & 1);
              tuse++;
              group_barrier(1 + ts, 64);   // both warps of the tile are done with the stage before it is refilled
              if (qb < 0) break;
              qa = next_block();
              if (qa >= 0) issue(qa, fa, (tuse + 1) & 1);
              mbar_wait(mb + (tuse & 1), (tuse >> 1) & 1);
              consume(qb, fb4, tuse & 1);
              tuse++;
              group_barrier(1 + ts, 64);
              if (qa < 0) break;
            }
          }
          long long tw2 = wd ? clock64() : 0;
          if (wd) { g_dp2_dbg[4] += tw1 - tw0; g_dp2_dbg[5] += tw2 - tw1; }
          float *xb = sm.xbuf + (size_t)ts * kRows * kPitch;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int n = 8 * nt + 2 * tig;
            xb[r0 * kPitch + n] = acc[nt][0]; xb[r0 * kPitch + n + 1] = acc[nt][1];
            xb[r1 * kPitch + n] = acc[nt][2]; xb[r1 * kPitch + n + 1] = acc[nt][3];
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
        }
      }
      __syncthreads();
      long long tdbg1 = dbg ? clock64() : 0;
      // ================= chain phase: the 32x32 diagonal block, one warp per tile, lanes = rows ============
      {
        const int ts = warp / kWpt;
        const int c = cb + ts;
        const int cw = (ts >> 1) & 1;  // spread the chain warps over the four SM sub-partitions
        if (c <= c_hi && (warp % kWpt) == cw) {
          const int q = w - c;
          const int J = BETA ? NBv - 1 - q : q;
          const int jbase = kBlk * J;
          const int s = c * kRows + lane;                      // my row's step
          const bool rowvalid = s < nsteps;
          const int t = BETA ? Tn - 2 - s : 1 + s;
          const float *utw = sm.ut + (size_t)ts * kBlk * kBlk;
          float *iow = sm.io + (size_t)ts * kRows * kPitch + lane * kPitch;
          const int FI = sm.fbuf[ts * kRows + lane];
          const float *rmax_blk = sm.rmax + jbase;

          // predecessor sums formed from the previous chunk's last row (the "row -1" of this tile): lane cj
          // computes column cj, handed to lane 0 column by column inside chain_column
          float d0m = 0.f;
          int d0f = kNegBig;
          {
            const float *pm = sm.stm + q * kBlk;
            d0f = sm.stf[q];
#pragma unroll
            for (int c4 = 0; c4 < kBlk; c4 += 4) {  // U[ci][lane] (row `lane` of the staged weights), 0 for ci >= lane
              const float4 u4 = *reinterpret_cast<const float4 *>(utw + lane * kBlk + c4);
              const float4 m4 = *reinterpret_cast<const float4 *>(pm + c4);
              d0m = fmaf(m4.x, u4.x, d0m); d0m = fmaf(m4.y, u4.y, d0m); d0m = fmaf(m4.z, u4.z, d0m); d0m = fmaf(m4.w, u4.w, d0m);
            }
            if (!(d0m > 0.f)) d0f = kNegBig;
          }
          // my row of the far-sum tile doubles as the mantissa row (slot = vertex offset)
          float *mrow = sm.xbuf + ((size_t)ts * kRows + lane) * kPitch;
          ChainRow st;
          st.CF = kNegBig;
          st.maxe = kNegBig;
          chain_sweep<BETA>(st, utw, mrow, iow, lane, FI, d0m, d0f, rowvalid, jbase, t, O, rmax_blk);

          // frame table entry of my row, chain state of the chunk's last valid row (one frame per block)
          if (rowvalid) sm.rmtab[t * NB + q] = st.maxe;
          const int rl = min(kRows, nsteps - c * kRows) - 1;
          if (lane == rl) {
#pragma unroll
            for (int ci = 0; ci < kBlk; ci++) sm.stm[q * kBlk + ci] = mrow[BETA ? kBlk - 1 - ci : ci];
            sm.stf[q] = st.CF;
          }
          // publish my row as A-operand for the tiles that will consume it as a previous row: normalised by the
          // row frame (value < 1), through shared memory into fragment order.  Producer step s feeds consumer row
          // (s + 1): rows 1..31 of this chunk's fragment tile, and row 0 of the next chunk's.
          {
            float *vt = sm.xbuf + (size_t)ts * kRows * kPitch;   // slots are already indexed by vertex offset = K index
            const float nsc = (st.maxe > kNegBig && st.CF > kNegBig) ? pow2i(st.CF - st.maxe) : 0.f;
#pragma unroll
            for (int k = 0; k < kBlk; k++) mrow[k] *= nsc;
            __syncwarp();
            constexpr int VP = kPitch;
            uint4 *ft = afrag + ((size_t)c * NBv + q) * 256;
            const int fgid = lane >> 2, ftig = lane & 3;
#pragma unroll
            for (int slice = 0; slice < 2; slice++)
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                // consumer rows rho = 16*slice + {fgid, fgid+8}  <- producer rows rho-1
                const int rA = 16 * slice + fgid - 1, rB = rA + 8;
                const int k0 = 16 * ks + 2 * ftig;
                float e[8];
                e[0] = rA >= 0 ? vt[rA * VP + k0] : 0.f;     e[1] = rA >= 0 ? vt[rA * VP + k0 + 1] : 0.f;
                e[2] = vt[rB * VP + k0];                      e[3] = vt[rB * VP + k0 + 1];
                e[4] = rA >= 0 ? vt[rA * VP + k0 + 8] : 0.f; e[5] = rA >= 0 ? vt[rA * VP + k0 + 9] : 0.f;
                e[6] = vt[rB * VP + k0 + 8];                  e[7] = vt[rB * VP + k0 + 9];
                uint4 hi, lo;
                split_bf16x2(e[0], e[1], hi.x, lo.x);
                split_bf16x2(e[2], e[3], hi.y, lo.y);
                split_bf16x2(e[4], e[5], hi.z, lo.z);
                split_bf16x2(e[6], e[7], hi.w, lo.w);
                if (slice == 0 && fgid == 0) {
                  // row 0 of this tile belongs to the previous chunk's producer (or the seed): keep a0 / a2
                  uint32_t *ph = reinterpret_cast<uint32_t *>(ft + (slice * 4 + ks * 2 + 0) * 32 + lane);
                  uint32_t *pl = reinterpret_cast<uint32_t *>(ft + (slice * 4 + ks * 2 + 1) * 32 + lane);
                  ph[1] = hi.y; ph[3] = hi.w; pl[1] = lo.y; pl[3] = lo.w;
                } else {
                  ft[(slice * 4 + ks * 2 + 0) * 32 + lane] = hi;
                  ft[(slice * 4 + ks * 2 + 1) * 32 + lane] = lo;
                }
              }
            // my chunk's last row is row 0 of the next chunk's tile
            if (c + 1 < NCv && lane < 4) {
              uint4 *fn = afrag + ((size_t)(c + 1) * NBv + q) * 256;
#pragma unroll
              for (int ks = 0; ks < 2; ks++) {
                const int k0 = 16 * ks + 2 * lane;   // lane = tig of (gid 0)
                uint32_t h0, l0, h2, l2;
                split_bf16x2(vt[31 * VP + k0], vt[31 * VP + k0 + 1], h0, l0);
                split_bf16x2(vt[31 * VP + k0 + 8], vt[31 * VP + k0 + 9], h2, l2);
                uint32_t *ph = reinterpret_cast<uint32_t *>(fn + (ks * 2 + 0) * 32 + lane);
                uint32_t *pl = reinterpret_cast<uint32_t *>(fn + (ks * 2 + 1) * 32 + lane);
                ph[0] = h0; ph[2] = h2; pl[0] = l0; pl[2] = l2;
              }
            }
          }
          __syncwarp();
          // lattice values: coalesced row writes (lane = column)
          const float *iot = sm.io + (size_t)ts * kRows * kPitch;
          const int j = jbase + lane;
          if (j < L) {
            for (int rr = 0; rr <= rl; rr++) {
              const int sr = c * kRows + rr;
              const int tr = BETA ? Tn - 2 - sr : 1 + sr;
              lat[(int64_t)tr * L + j] = iot[rr * kPitch + lane];
            }
          }
        }
      }
      __syncthreads();
      if (dbg && threadIdx.x == 0 && blockIdx.x == 0) {
        long long tdbg2 = clock64();
        g_dp2_dbg[BETA ? 2 : 0] += tdbg1 - tdbg0;
        g_dp2_dbg[BETA ? 3 : 1] += tdbg2 - tdbg1;
      }
    }
  }
}

__global__ void __launch_bounds__(kDp2Threads, 1)
dag_alpha_beta_blocked_kernel(const float *__restrict__ match, const int64_t *__restrict__ olen,
                              const int64_t *__restrict__ tlen, float *__restrict__ alpha, float *__restrict__ beta,
                              unsigned char *__restrict__ ws, int M, int L, int Tl, TileLayout lay,
                              int32_t *__restrict__ status, int dbg) {
  extern __shared__ __align__(128) unsigned char dp2_smem[];
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
  Dp2Smem sm;
  float *p = reinterpret_cast<float *>(dp2_smem);
  sm.tstage = reinterpret_cast<uint4 *>(p);  p += kTpw * 2 * 1024;
  sm.mbar = reinterpret_cast<uint64_t *>(p); p += kTpw * 2 * 2;
  sm.ut = p;    p += kTpw * kBlk * kBlk;
  sm.xbuf = p;  p += kTpw * kRows * kPitch;
  sm.io = p;    p += kTpw * kRows * kPitch;
  sm.fbuf = reinterpret_cast<int *>(p);  p += kTpw * kRows;
  sm.rmax = p;  p += lay.NB * kBlk;
  sm.stm = p;   p += lay.NB * kBlk;
  sm.stf = reinterpret_cast<int *>(p);  p += lay.NB * 4;
  sm.rmtab = reinterpret_cast<int *>(p);
  const float *m = match + b * latsz;
  unsigned char *wsb = ws + (size_t)b * lay.sample_bytes;
  if (is_beta) blocked_chain<true>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl, dbg != 0);
  else blocked_chain<false>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl, dbg != 0);
}

size_t dp2_smem_bytes(int M, int L) {
  TileLayout lay = TileLayout::make(L, M);
  return sizeof(float) * ((size_t)kTpw * 2 * 1024 + (size_t)kTpw * 4 + (size_t)kTpw * kBlk * kBlk + 2 * (size_t)kTpw * kRows * kPitch + (size_t)kTpw * kRows +
                          2 * (size_t)lay.NB * kBlk + (size_t)lay.NB * 4 + (size_t)M * lay.NB);
}

int launch_dag_prep(const float *links, const int64_t *olen, void *workspace, int B, int M, int L, int Tl, int fmt,
                    cudaStream_t st);

size_t dp2_workspace_bytes(int B, int M, int L) { return TileLayout::make(L, M).sample_bytes * (size_t)B; }

bool dp2_supported(int M, int L) { return dp2_smem_bytes(M, L) <= 226 * 1024 && L >= 1 && (L + kBlk - 1) / kBlk <= 64; }

int launch_alpha_beta_blocked(const float *match, const float *links, const int64_t *olen, const int64_t *tlen,
                              float *alpha, float *beta, int B, int M, int L, int Tl, bool grad, void *workspace,
                              int32_t *status, cudaStream_t st) {
  prof_mark(0, st);
  int rc = launch_dag_prep(links, olen, workspace, B, M, L, Tl, 0, st);
  if (rc) return rc;
  prof_mark(1, st);
  TileLayout lay = TileLayout::make(L, M);
  dim3 grid(B, grad ? 2 : 1);
  const size_t smem = dp2_smem_bytes(M, L);
  static const bool dbg = getenv("DAGB200_DP2_DEBUG") != nullptr;
  cudaFuncSetAttribute(dag_alpha_beta_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dag_alpha_beta_blocked_kernel<<<grid, kDp2Threads, smem, st>>>(match, olen, tlen, alpha, beta,
                                                                (unsigned char *)workspace, M, L, Tl, lay, status, dbg ? 1 : 0);
  DAGB200_CHECK_LAUNCH("dag_alpha_beta_blocked_kernel");
  prof_mark(2, st);
  if (dbg) {
    long long h[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_dp2_dbg, sizeof(h));
    fprintf(stderr, "[dp2 dbg] cumulative cycles CTA0: alpha gemm %lld chain %lld | beta gemm %lld chain %lld | warp0: setup %lld loop %lld | reframes alpha %lld beta %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  return 0;
}

}  // namespace dagb200
