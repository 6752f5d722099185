// dag_viterbi2.cu -- blocked max-plus (Viterbi) recurrence with back-pointers, for sm_100a.
//
// Replaces calculate_maxalpha_kernel + calculate_backtrace_kernel (reference dag_best_alignment.cu:39-130,
// 170-185) on the fp32 / config 1 path.  Arithmetic is the reference's, operation for operation: a candidate is
// ONE fp32 add (previous value + transition), the cell value one more add (+ emission), so lattice values are
// bit-identical; the arg-max reproduces the reference's tie-break (TRANS_BLOCK_SIZE = 4: candidate delta lives in
// lane (delta-1)%4, a lane keeps its first strict maximum = smallest delta, lanes merge with priority 0,2,1,3)
// without ever comparing keys in the inner loop:
//   * candidates are visited in DESCENDING delta and kept in four per-class maxima updated with '>=', so inside a
//     class the smallest delta wins; classes are then merged in priority order with strict '>'.
// Organisation = dag_dp2.cu: vertices in blocks of 32, rows in chunks of 32, anti-diagonal waves inside one CTA
// per utterance.  Far predecessors: lanes = destination columns, the 32x32 transition tile column lives in
// registers and is reused for 16 rows (the reference re-reads every transition for every row); near
// predecessors: lanes = rows, serial sweep over the 32 columns, one shuffle per column.
#include <cstdlib>

#include "common.cuh"

namespace dagb200 {

constexpr int kV2Threads = 512;
constexpr int kV2Warps = kV2Threads / 32;
constexpr int kV2Tpw = kV2Warps / 2;      // tiles per batch (2 warps x 16 rows per tile in the far phase)
constexpr int kV2Pitch = 33;
constexpr int kVB = 32;

__device__ __forceinline__ int rank4(int delta) {  // priority of the class of `delta`: classes 0,2,1,3 -> 0,1,2,3
  const int cl = (delta - 1) & 3;
  return ((cl & 1) << 1) | (cl >> 1);
}
// is candidate (v1, d1) preferred over (v2, d2)?  (-inf never wins)
__device__ __forceinline__ bool better(float v1, int d1, float v2, int d2) {
  if (v1 > v2) return true;
  if (v1 < v2 || !(v1 > neg_inf_f())) return false;
  const int r1 = rank4(d1), r2 = rank4(d2);
  return r1 < r2 || (r1 == r2 && d1 < d2);
}
__device__ __forceinline__ void cp_async_f32_v(float *smem_dst, const float *gsrc) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(gsrc) : "memory");
}

struct V2Smem {
  float *xv;     // [tiles][32][pitch] best far candidate value per (row, column)
  int *xd;       // [tiles][32][pitch] its delta; overwritten column by column with the chosen delta (trace)
  float *ed;     // [tiles][32][32]    diagonal transition block, [cj][ci]
  float *io;     // [tiles][32][pitch] emissions in, cell values out
  float *vs;     // [warps][16][32]    previous-row values of the current source block (far phase)
  unsigned char *flag;  // [M][NB]     1 if block row has a finite value
};

// one column of the diagonal sweep (lanes = rows)
template <int CJ>
__device__ __forceinline__ void vit_column(float (&vrow)[kVB], const float *edw, const float (&mm)[8], float *iow,
                                           const float *xvw, int *xdw, int lane, float d0v, int d0d,
                                           bool rowvalid, int jbase, int t, int O, bool &anyfin) {
  const float ninf = neg_inf_f();
  // best in-block predecessor of MY row for the next row's column CJ; classes are compile-time here
  float bv[4] = {ninf, ninf, ninf, ninf};
  int bd[4] = {0, 0, 0, 0};
#pragma unroll
  for (int c4 = 0; c4 < CJ; c4 += 4) {
    const float4 e4 = *reinterpret_cast<const float4 *>(edw + CJ * kVB + c4);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int ci = c4 + k;
      if (ci < CJ) {
        const int delta = CJ - ci, cl = (delta - 1) & 3;
        const float x = vrow[ci] + (k == 0 ? e4.x : k == 1 ? e4.y : k == 2 ? e4.z : e4.w);
        if (x >= bv[cl]) { bv[cl] = x; bd[cl] = delta; }   // descending delta: later = smaller delta wins ties
      }
    }
  }
  float nv = bv[0]; int nd = bd[0];
  if (bv[2] > nv) { nv = bv[2]; nd = bd[2]; }
  if (bv[1] > nv) { nv = bv[1]; nd = bd[1]; }
  if (bv[3] > nv) { nv = bv[3]; nd = bd[3]; }
  // hand to the next row
  float rv = __shfl_up_sync(0xffffffffu, nv, 1);
  int rd = __shfl_up_sync(0xffffffffu, nd, 1);
  const float zv = __shfl_sync(0xffffffffu, d0v, CJ);
  const int zd = __shfl_sync(0xffffffffu, d0d, CJ);
  if (lane == 0) { rv = zv; rd = zd; }
  // merge with the far candidate
  const float fv = xvw[CJ];
  const int fd = xdw[CJ];
  float best = fv; int bdl = fd;
  if (better(rv, rd, fv, fd)) { best = rv; bdl = rd; }
  const int j = jbase + CJ;
  const bool valid = rowvalid && j >= t && j < O;
  float val = ninf; int dl = 0;
  if (valid) {
    val = best + mm[CJ & 7];
    if (best > ninf) dl = bdl;
  }
  vrow[CJ] = val;
  iow[CJ] = val;
  xdw[CJ] = dl;
  anyfin = anyfin || (val > ninf);
}

template <int CJ0>
__device__ __forceinline__ void vit_group(float (&vrow)[kVB], const float *edw, float *iow, const float *xvw,
                                          int *xdw, int lane, float d0v, int d0d, bool rowvalid, int jbase, int t,
                                          int O, bool &anyfin) {
  float mm[8];
#pragma unroll
  for (int k = 0; k < 8; k++) mm[k] = iow[CJ0 + k];
  vit_column<CJ0 + 0>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 1>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 2>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 3>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 4>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 5>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 6>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
  vit_column<CJ0 + 7>(vrow, edw, mm, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
}

// ---- value-only variants (cluster kernel): the forward sweep keeps NO back-pointers -- max is exact in any order, so
// the lattice is bit-identical -- and the arg-max with the reference's tie-break is recomputed during the backtrace
// for the cells on the path only (M cells instead of M x L).
// order-preserving float <-> int32 key (an involution): lets shared-memory integer atomics take the float maximum
__device__ __forceinline__ int f2key(float x) { const int b = __float_as_int(x); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

template <int CJ>
__device__ __forceinline__ void vitv_column(float (&vrow)[kVB], const float *edw, const float (&mm)[8], float *iow,
                                            const int *xvw, int lane, float d0v, bool rowvalid, int jbase, int t,
                                            int O, bool &anyfin) {
  const float ninf = neg_inf_f();
  float n0 = ninf, n1 = ninf;      // best in-block predecessor of MY row for the next row's column CJ
#pragma unroll
  for (int c4 = 0; c4 < CJ; c4 += 4) {
    const float4 e4 = *reinterpret_cast<const float4 *>(edw + CJ * kVB + c4);
    if (c4 + 0 < CJ) n0 = fmaxf(n0, vrow[c4 + 0] + e4.x);
    if (c4 + 1 < CJ) n1 = fmaxf(n1, vrow[c4 + 1] + e4.y);
    if (c4 + 2 < CJ) n0 = fmaxf(n0, vrow[c4 + 2] + e4.z);
    if (c4 + 3 < CJ) n1 = fmaxf(n1, vrow[c4 + 3] + e4.w);
  }
  float rv = __shfl_up_sync(0xffffffffu, fmaxf(n0, n1), 1);
  const float zv = __shfl_sync(0xffffffffu, d0v, CJ);
  if (lane == 0) rv = zv;
  const float best = fmaxf(rv, key2f(xvw[CJ]));
  const int j = jbase + CJ;
  const bool valid = rowvalid && j >= t && j < O;
  const float val = valid ? best + mm[CJ & 7] : ninf;
  vrow[CJ] = val;
  iow[CJ] = val;
  anyfin = anyfin || (val > ninf);
}
template <int CJ0>
__device__ __forceinline__ void vitv_group(float (&vrow)[kVB], const float *edw, float *iow, const int *xvw, int lane,
                                           float d0v, bool rowvalid, int jbase, int t, int O, bool &anyfin) {
  float mm[8];
#pragma unroll
  for (int k = 0; k < 8; k++) mm[k] = iow[CJ0 + k];
  vitv_column<CJ0 + 0>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 1>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 2>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 3>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 4>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 5>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 6>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
  vitv_column<CJ0 + 7>(vrow, edw, mm, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
}

__global__ void __launch_bounds__(kV2Threads, 1)
dag_viterbi_blocked_kernel(const float *__restrict__ match, const float *__restrict__ links,
                           const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen,
                           float *__restrict__ lattice, uint16_t *__restrict__ trace, int32_t *__restrict__ path,
                           int M, int L, int Tl, int NB, int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char v2_smem[];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t latsz = (int64_t)M * L;
  const float ninf = neg_inf_f();
  float *lat = lattice + b * latsz;
  uint16_t *trg = trace + b * latsz;
  int32_t *prow = path + (int64_t)b * L;
  const float *m = match + b * latsz;
  const float *E = links + (int64_t)b * L * Tl;
  for (int j = threadIdx.x; j < L; j += kV2Threads) prow[j] = -1;

  int st = DAGB200_ST_OK;
  if (Tn < 2 || O < 2) st = DAGB200_ST_LEN_LT2;
  else if (O < Tn || O > L || Tn > M) st = DAGB200_ST_GRAPH_SMALL;
  else if ((int64_t)(Tn - 1) * Tl + 1 < O) st = DAGB200_ST_TOO_SHORT;
  if (st != DAGB200_ST_OK) {
    for (int64_t x = threadIdx.x; x < latsz; x += kV2Threads) lat[x] = ninf;
    if (status && threadIdx.x == 0) status[b] = st;
    return;
  }

  V2Smem sm;
  {
    float *p = reinterpret_cast<float *>(v2_smem);
    sm.ed = p;  p += kV2Tpw * kVB * kVB;
    sm.xv = p;  p += kV2Tpw * kVB * kV2Pitch;
    sm.xd = reinterpret_cast<int *>(p);  p += kV2Tpw * kVB * kV2Pitch;
    sm.io = p;  p += kV2Tpw * kVB * kV2Pitch;
    sm.vs = p;  p += kV2Warps * 16 * kVB;
    sm.flag = reinterpret_cast<unsigned char *>(p);
  }
  const int NBv = (O + kVB - 1) / kVB;
  const int nsteps = Tn - 1;
  const int NCv = (nsteps + kVB - 1) / kVB;
  const int band = 1 + (Tl - 1) / kVB;

  // prologue: padding, seed row, flags
  for (int x = threadIdx.x; x < M * NB; x += kV2Threads) sm.flag[x] = 0;
  {
    const int64_t tail0 = (int64_t)Tn * L;
    for (int64_t x = tail0 + threadIdx.x; x < latsz; x += kV2Threads) lat[x] = ninf;
    const int c0 = NBv * kVB;
    if (c0 < L) {
      const int wcols = L - c0;
      for (int x = threadIdx.x; x < Tn * wcols; x += kV2Threads) lat[(int64_t)(x / wcols) * L + c0 + x % wcols] = ninf;
    }
    for (int j = threadIdx.x; j < min(L, c0); j += kV2Threads) lat[j] = (j == 0) ? m[0] : ninf;
  }
  __syncthreads();
  if (threadIdx.x == 0 && m[0] > ninf) sm.flag[0] = 1;
  __syncthreads();

  const int nwaves = NBv + NCv - 1;
  for (int w = 0; w < nwaves; w++) {
    const int c_lo = max(0, w - NBv + 1), c_hi = min(NCv - 1, w);
    for (int cb = c_lo; cb <= c_hi; cb += kV2Tpw) {
      // ======================= far phase: lanes = destination columns, 16 rows per warp =======================
      {
        const int ts = warp >> 1, sl = warp & 1;
        const int c = cb + ts;
        if (c <= c_hi) {
          const int J = w - c;
          const int j = kVB * J + lane;
          // stage the diagonal transition block ([cj][ci]) and the emissions of this tile asynchronously
          {
            float *edw = sm.ed + (size_t)ts * kVB * kVB;
            float *iow = sm.io + (size_t)ts * kVB * kV2Pitch;
            for (int rr = sl; rr < kVB; rr += 2) {
              // row rr of the diagonal block = source vertex ci = rr, lane = destination cj
              const int i = kVB * J + rr, k = lane - rr - 1;
              if (k >= 0 && k < Tl && i < O && j < O) cp_async_f32_v(edw + lane * kVB + rr, E + (int64_t)i * Tl + k);
              else edw[lane * kVB + rr] = ninf;
              const int s = c * kVB + rr;
              if (s < nsteps && j < L) cp_async_f32_v(iow + rr * kV2Pitch + lane, m + (int64_t)(1 + s) * L + j);
              else iow[rr * kV2Pitch + lane] = ninf;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
          }
          float *xvw = sm.xv + (size_t)ts * kVB * kV2Pitch + (16 * sl) * kV2Pitch + lane;
          int *xdw = sm.xd + (size_t)ts * kVB * kV2Pitch + (16 * sl) * kV2Pitch + lane;
#pragma unroll
          for (int rr = 0; rr < 16; rr++) { xvw[rr * kV2Pitch] = ninf; xdw[rr * kV2Pitch] = 0; }
          float *vsw = sm.vs + (size_t)warp * 16 * kVB;
          const int s_first = c * kVB + 16 * sl;           // previous-row index of my first row
          const int qlo = max(0, J - band);
          for (int I = qlo; I < J; I++) {
            // skip source blocks without a finite value on any of my 16 previous rows
            bool any = false;
            if (lane < 16 && s_first + lane < nsteps) any = sm.flag[(s_first + lane) * NB + I] != 0;
            if (!__any_sync(0xffffffffu, any)) continue;
            // my column of the transition tile: E[32I+ii][j - (32I+ii) - 1]
            float ecol[kVB];
#pragma unroll
            for (int ii = 0; ii < kVB; ii++) {
              const int i = kVB * I + ii, k = j - i - 1;
              ecol[ii] = (k < Tl && j < O) ? __ldg(E + (int64_t)i * Tl + k) : ninf;   // i < 32J <= O always here
            }
            __syncwarp();
#pragma unroll 4
            for (int rr = 0; rr < 16; rr++) {
              const int tp = s_first + rr;
              vsw[rr * kVB + lane] = (tp < nsteps) ? lat[(int64_t)tp * L + kVB * I + lane] : ninf;
            }
            __syncwarp();
            const int dbase = kVB * (J - I) + lane;        // delta of source ii is dbase - ii
            // class of slot u = ii & 3 for this lane: (dbase - u - 1) & 3 ; merge order = priority 0,2,1,3
            int ord[4];   // ord[r] = slot whose class has priority r: class = bitrev2(r), slot = (dbase-1-class) & 3
#pragma unroll
            for (int r4 = 0; r4 < 4; r4++) ord[r4] = (dbase - 1 - (((r4 & 1) << 1) | (r4 >> 1))) & 3;
            for (int rr = 0; rr < 16; rr++) {
              float bv0 = ninf, bv1 = ninf, bv2 = ninf, bv3 = ninf;
              int bi0 = 0, bi1 = 0, bi2 = 0, bi3 = 0;
#pragma unroll
              for (int c4 = 0; c4 < kVB; c4 += 4) {
                const float4 a4 = *reinterpret_cast<const float4 *>(vsw + rr * kVB + c4);
                float x;
                x = a4.x + ecol[c4 + 0]; if (x >= bv0) { bv0 = x; bi0 = c4 + 0; }
                x = a4.y + ecol[c4 + 1]; if (x >= bv1) { bv1 = x; bi1 = c4 + 1; }
                x = a4.z + ecol[c4 + 2]; if (x >= bv2) { bv2 = x; bi2 = c4 + 2; }
                x = a4.w + ecol[c4 + 3]; if (x >= bv3) { bv3 = x; bi3 = c4 + 3; }
              }
              const float bvs[4] = {bv0, bv1, bv2, bv3};
              const int bis[4] = {bi0, bi1, bi2, bi3};
              // slots in priority order, strict '>'
              float nv = ninf; int ni = 0;
#pragma unroll
              for (int r4 = 0; r4 < 4; r4++) {
                const int u = ord[r4];
                const float v = (u == 0) ? bvs[0] : (u == 1) ? bvs[1] : (u == 2) ? bvs[2] : bvs[3];
                const int ii = (u == 0) ? bis[0] : (u == 1) ? bis[1] : (u == 2) ? bis[2] : bis[3];
                if (r4 == 0 || v > nv) { nv = v; ni = ii; }
              }
              const int nd = dbase - ni;
              const float rv = xvw[rr * kV2Pitch];
              const int rd = xdw[rr * kV2Pitch];
              // this block is nearer than everything accumulated so far: it wins ties unless its class ranks lower
              if (nv > rv || (nv == rv && nv > ninf && rank4(nd) <= rank4(rd))) {
                xvw[rr * kV2Pitch] = nv;
                xdw[rr * kV2Pitch] = nd;
              }
            }
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
        }
      }
      __syncthreads();
      // ======================= chain phase: lanes = rows, serial over the 32 columns =======================
      {
        const int ts = warp >> 1;
        const int c = cb + ts;
        const int cw = (ts >> 1) & 1;
        if (c <= c_hi && (warp & 1) == cw) {
          const int J = w - c;
          const int jbase = kVB * J;
          const int s = c * kVB + lane;
          const bool rowvalid = s < nsteps;
          const int t = 1 + s;
          const float *edw = sm.ed + (size_t)ts * kVB * kVB;
          float *iow = sm.io + (size_t)ts * kVB * kV2Pitch + lane * kV2Pitch;
          const float *xvw = sm.xv + (size_t)ts * kVB * kV2Pitch + lane * kV2Pitch;
          int *xdw = sm.xd + (size_t)ts * kVB * kV2Pitch + lane * kV2Pitch;
          // "row -1": best in-block predecessor formed from the last row of the previous chunk; lane = column
          float d0v = ninf; int d0d = 0;
          {
            const int tp = c * kVB;   // previous-row index of the chunk's first row
            const float pv = (jbase + lane < L) ? lat[(int64_t)tp * L + jbase + lane] : ninf;
            for (int c4 = 0; c4 < kVB; c4 += 4) {   // any order: `better` is a total order
              const float4 e4 = *reinterpret_cast<const float4 *>(edw + lane * kVB + c4);   // -inf for ci >= lane
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int ci = c4 + k;
                const float x = __shfl_sync(0xffffffffu, pv, ci) + (k == 0 ? e4.x : k == 1 ? e4.y : k == 2 ? e4.z : e4.w);
                const int delta = lane - ci;
                if (ci < lane && better(x, delta, d0v, d0d)) { d0v = x; d0d = delta; }
              }
            }
          }
          float vrow[kVB];
          bool anyfin = false;
          vit_group<0>(vrow, edw, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
          vit_group<8>(vrow, edw, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
          vit_group<16>(vrow, edw, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
          vit_group<24>(vrow, edw, iow, xvw, xdw, lane, d0v, d0d, rowvalid, jbase, t, O, anyfin);
          if (rowvalid) sm.flag[t * NB + J] = anyfin ? 1 : 0;
          __syncwarp();
          // coalesced writes of the tile: cell values and back-pointers (lane = column)
          const int rl = min(kVB, nsteps - c * kVB) - 1;
          const float *iot = sm.io + (size_t)ts * kVB * kV2Pitch;
          const int *trt = sm.xd + (size_t)ts * kVB * kV2Pitch;
          const int j = jbase + lane;
          if (j < L) {
            for (int rr = 0; rr <= rl; rr++) {
              const int64_t off = (int64_t)(1 + c * kVB + rr) * L + j;
              lat[off] = iot[rr * kV2Pitch + lane];
              trg[off] = (uint16_t)trt[rr * kV2Pitch + lane];
            }
          }
        }
      }
      __syncthreads();
    }
  }

  // backtrace (dag_best_alignment.cu:178-184)
  if (threadIdx.x == 0) {
    int code = DAGB200_ST_OK;
    if (!(lat[(int64_t)(Tn - 1) * L + O - 1] > ninf)) {
      code = DAGB200_ST_NO_PATH;
    } else {
      int pos = O - 1;
      for (int i = Tn - 1; i >= 0; i--) {
        prow[pos] = i;
        if (i == 0) break;
        const int d = trg[(int64_t)i * L + pos];
        if (d == 0) { code = DAGB200_ST_NO_PATH; break; }
        pos -= d;
      }
    }
    if (status) status[b] = code;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Two CTAs per utterance (one thread-block cluster): the tiles of an anti-diagonal wave are independent, so the
// chunks of a wave are split by parity between the two CTAs of the cluster -- 128 CTAs instead of 64 at B = 64.
// Everything a tile needs from another tile already travels through global memory (lattice rows, back-pointers);
// the per-(row, block) "has a finite value" flags move to global memory as well, and one cluster barrier
// (release / acquire) per wave publishes a wave's results to both CTAs.  Far phase: 4 warps x 8 rows per tile.
constexpr int kVcTpw = kV2Warps / 4;     // tiles per batch per CTA

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kV2Threads, 1)
dag_viterbi_cluster_kernel(const float *__restrict__ match, const float *__restrict__ links,
                           const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen,
                           float *lattice, int32_t *__restrict__ path,
                           unsigned char *gflags, int M, int L, int Tl, int NB, int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char v2_smem[];
  const int b = blockIdx.x >> 1, rank = blockIdx.x & 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t latsz = (int64_t)M * L;
  const float ninf = neg_inf_f();
  float *lat = lattice + b * latsz;
  int32_t *prow = path + (int64_t)b * L;
  const float *m = match + b * latsz;
  const float *E = links + (int64_t)b * L * Tl;
  unsigned char *flag = gflags + (size_t)b * M * NB;
  if (rank == 0)
    for (int j = threadIdx.x; j < L; j += kV2Threads) prow[j] = -1;

  int st = DAGB200_ST_OK;
  if (Tn < 2 || O < 2) st = DAGB200_ST_LEN_LT2;
  else if (O < Tn || O > L || Tn > M) st = DAGB200_ST_GRAPH_SMALL;
  else if ((int64_t)(Tn - 1) * Tl + 1 < O) st = DAGB200_ST_TOO_SHORT;
  if (st != DAGB200_ST_OK) {   // both CTAs of the cluster leave together: no barrier is pending
    if (rank == 0) {
      for (int64_t x = threadIdx.x; x < latsz; x += kV2Threads) lat[x] = ninf;
      if (status && threadIdx.x == 0) status[b] = st;
    }
    return;
  }

  float *s_ed, *s_io, *s_vs;
  int *s_xk;     // [tiles][32][pitch] best far candidate per (row, column) as an order-preserving key
  {
    float *p = reinterpret_cast<float *>(v2_smem);
    s_ed = p;  p += kVcTpw * kVB * kVB;
    s_xk = reinterpret_cast<int *>(p);  p += kVcTpw * kVB * kV2Pitch;
    s_io = p;  p += kVcTpw * kVB * kV2Pitch;
    s_vs = p;    // [warps][2][32][32] previous-row values of a unit's source block, double-buffered
  }
  const int NBv = (O + kVB - 1) / kVB;
  const int nsteps = Tn - 1;
  const int NCv = (nsteps + kVB - 1) / kVB;
  const int band = 1 + (Tl - 1) / kVB;

  // prologue (split between the two CTAs): flags, padding, seed row
  if (rank == 1) {
    for (int x = threadIdx.x; x < M * NB; x += kV2Threads) flag[x] = 0;
    __syncthreads();
    if (threadIdx.x == 0 && m[0] > ninf) flag[0] = 1;
  } else {
    const int64_t tail0 = (int64_t)Tn * L;
    for (int64_t x = tail0 + threadIdx.x; x < latsz; x += kV2Threads) lat[x] = ninf;
    const int c0 = NBv * kVB;
    if (c0 < L) {
      const int wcols = L - c0;
      for (int x = threadIdx.x; x < Tn * wcols; x += kV2Threads) lat[(int64_t)(x / wcols) * L + c0 + x % wcols] = ninf;
    }
    for (int j = threadIdx.x; j < min(L, c0); j += kV2Threads) lat[j] = (j == 0) ? m[0] : ninf;
  }
  cluster_sync_all();

  const int nwaves = NBv + NCv - 1;
  for (int w = 0; w < nwaves; w++) {
    const int c_lo = max(0, w - NBv + 1), c_hi = min(NCv - 1, w);
    const int c_first = c_lo + ((c_lo ^ rank) & 1);          // my chunks: parity = rank
    for (int cb = c_first; cb <= c_hi; cb += 2 * kVcTpw) {
      // ======================= staging: diagonal transition blocks and emissions of the batch's tiles ==========
      {
        const int ts = warp >> 2, sl = warp & 3;
        const int c = cb + 2 * ts;
        if (c <= c_hi) {
          const int J = w - c;
          const int j = kVB * J + lane;
          float *edw = s_ed + (size_t)ts * kVB * kVB;
          float *iow = s_io + (size_t)ts * kVB * kV2Pitch;
          for (int rr = sl; rr < kVB; rr += 4) {
            const int i = kVB * J + rr, k = lane - rr - 1;
            if (k >= 0 && k < Tl && i < O && j < O) cp_async_f32_v(edw + lane * kVB + rr, E + (int64_t)i * Tl + k);
            else edw[lane * kVB + rr] = ninf;
            const int s = c * kVB + rr;
            if (s < nsteps && j < L) cp_async_f32_v(iow + rr * kV2Pitch + lane, m + (int64_t)(1 + s) * L + j);
            else iow[rr * kV2Pitch + lane] = ninf;
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          int *xk = s_xk + (size_t)ts * kVB * kV2Pitch + (8 * sl) * kV2Pitch + lane;
#pragma unroll
          for (int rr = 0; rr < 8; rr++) xk[rr * kV2Pitch] = f2key(ninf);
        }
      }
      __syncthreads();
      // ======================= far phase: units = (tile, source block), dealt round-robin to the 16 warps ========
      // A unit covers all 32 rows of its tile: the transition column of the lane's destination vertex is fetched once
      // per unit (not once per row quarter), every warp gets the same number of units whatever the tiles' depths, and
      // the per-(row, column) maxima of the units meet in shared memory through integer atomics on order-preserving keys.
      {
        int nun[kVcTpw], total = 0;
#pragma unroll
        for (int ts = 0; ts < kVcTpw; ts++) {
          const int c = cb + 2 * ts, J = w - c;
          nun[ts] = (c <= c_hi) ? J - max(0, J - band) : 0;
          total += nun[ts];
        }
        // few units (banded transitions, first waves): split every unit into four row quarters so that all warps work
        const int sf = (total < kV2Warps) ? 4 : 1;
        const int nrow = kVB / sf;
        const int nunits = total * sf;
        float *vs0 = s_vs + (size_t)warp * 2 * kVB * kVB;    // two slabs per warp: the next unit's rows land while this one is scored
        const bool vec_ok = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(lat) & 15) == 0);
        // unit -> (tile, source block, first row)
        auto decode = [&](int u, int &ts, int &I, int &row0) {
          int r = u / sf;
          row0 = (u % sf) * nrow;
          ts = 0;
#pragma unroll
          for (int x = 0; x < kVcTpw - 1; x++)
            if (ts == x && r >= nun[x]) { r -= nun[x]; ts = x + 1; }
          const int J = w - (cb + 2 * ts);
          I = max(0, J - band) + r;
        };
        // start fetching a unit: liveness flag of one row per lane (consumed later) and the rows themselves
        auto stage = [&](int u, float *slab, bool &any) {
          int ts, I, row0;
          decode(u, ts, I, row0);
          const int tp0 = (cb + 2 * ts) * kVB + row0;        // previous-row index of the unit's first row
          any = (lane < nrow) && (tp0 + lane < nsteps) && __ldcg(flag + (tp0 + lane) * NB + I) != 0;
          if (vec_ok) {
            const int q = lane & 7;
            for (int rr = lane >> 3; rr < nrow; rr += 4) {
              float *dst = slab + rr * kVB + 4 * q;
              if (tp0 + rr < nsteps) {
                const uint32_t a = (uint32_t)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(lat + (int64_t)(tp0 + rr) * L + kVB * I + 4 * q) : "memory");
              } else {
                *reinterpret_cast<float4 *>(dst) = make_float4(ninf, ninf, ninf, ninf);
              }
            }
          } else {
            const float *lp = lat + (int64_t)tp0 * L + kVB * I + lane;
            for (int rr = 0; rr < nrow; rr++) slab[rr * kVB + lane] = (tp0 + rr < nsteps) ? __ldcg(lp + (int64_t)rr * L) : ninf;
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int buf = 0;
        bool any_cur = false, any_next = false;
        if (warp < nunits) stage(warp, vs0, any_cur);
        for (int u = warp; u < nunits; u += kV2Warps) {
          const bool more = u + kV2Warps < nunits;
          if (more) stage(u + kV2Warps, vs0 + (buf ^ 1) * kVB * kVB, any_next);
          if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          if (__any_sync(0xffffffffu, any_cur)) {
            int ts, I, row0;
            decode(u, ts, I, row0);
            const int J = w - (cb + 2 * ts);
            const int j = kVB * J + lane;
            float ecol[kVB];
            {
              const int k0 = j - kVB * I - 1;                // transition index from the block's first source vertex; >= 31
              const float *ep = E + (int64_t)(kVB * I) * Tl + k0;
              const bool jok = j < O;
#pragma unroll
              for (int ii = 0; ii < kVB; ii++) ecol[ii] = (jok && k0 - ii < Tl) ? __ldg(ep + (int64_t)ii * (Tl - 1)) : ninf;
            }
            const float *vsw = vs0 + buf * kVB * kVB;
            int *xk = s_xk + (size_t)ts * kVB * kV2Pitch + row0 * kV2Pitch + lane;
#pragma unroll 4
            for (int rr = 0; rr < nrow; rr++) {
              float b0 = ninf, b1 = ninf;
#pragma unroll
              for (int c4 = 0; c4 < kVB; c4 += 4) {
                const float4 a4 = *reinterpret_cast<const float4 *>(vsw + rr * kVB + c4);
                b0 = fmaxf(fmaxf(b0, a4.x + ecol[c4 + 0]), a4.y + ecol[c4 + 1]);
                b1 = fmaxf(fmaxf(b1, a4.z + ecol[c4 + 2]), a4.w + ecol[c4 + 3]);
              }
              const float bb = fmaxf(b0, b1);
              if (bb > ninf) atomicMax(xk + rr * kV2Pitch, f2key(bb));
            }
          }
          __syncwarp();
          any_cur = any_next;
          buf ^= 1;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
      __syncthreads();
      // ======================= chain phase: lanes = rows, serial over the 32 columns =======================
      {
        const int ts = warp >> 2;
        const int c = cb + 2 * ts;
        if (c <= c_hi && (warp & 3) == (ts & 3)) {
          const int J = w - c;
          const int jbase = kVB * J;
          const int s = c * kVB + lane;
          const bool rowvalid = s < nsteps;
          const int t = 1 + s;
          const float *edw = s_ed + (size_t)ts * kVB * kVB;
          float *iow = s_io + (size_t)ts * kVB * kV2Pitch + lane * kV2Pitch;
          const int *xvw = s_xk + (size_t)ts * kVB * kV2Pitch + lane * kV2Pitch;
          float d0v = ninf;
          {
            const int tp = c * kVB;   // previous-row index of the chunk's first row (written by the other CTA)
            const float pv = (jbase + lane < L) ? __ldcg(lat + (int64_t)tp * L + jbase + lane) : ninf;
            for (int c4 = 0; c4 < kVB; c4 += 4) {
              const float4 e4 = *reinterpret_cast<const float4 *>(edw + lane * kVB + c4);
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const int ci = c4 + k;
                const float x = __shfl_sync(0xffffffffu, pv, ci) + (k == 0 ? e4.x : k == 1 ? e4.y : k == 2 ? e4.z : e4.w);
                if (ci < lane) d0v = fmaxf(d0v, x);
              }
            }
          }
          float vrow[kVB];
          bool anyfin = false;
          vitv_group<0>(vrow, edw, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
          vitv_group<8>(vrow, edw, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
          vitv_group<16>(vrow, edw, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
          vitv_group<24>(vrow, edw, iow, xvw, lane, d0v, rowvalid, jbase, t, O, anyfin);
          if (rowvalid) flag[t * NB + J] = anyfin ? 1 : 0;
          __syncwarp();
          const int rl = min(kVB, nsteps - c * kVB) - 1;
          const float *iot = s_io + (size_t)ts * kVB * kV2Pitch;
          const int j = jbase + lane;
          if (j < L) {
            for (int rr = 0; rr <= rl; rr++) {
              const int64_t off = (int64_t)(1 + c * kVB + rr) * L + j;
              lat[off] = iot[rr * kV2Pitch + lane];
            }
          }
        }
      }
      __syncthreads();
    }
    cluster_sync_all();    // wave w is complete in both CTAs and visible to both
  }

  // backtrace (dag_best_alignment.cu:178-184) with the back-pointers recomputed on the way: for the cell (i, pos) of the
  // path, all kV2Threads threads of CTA 0 score its candidates delta = 1 .. min(pos, Tl) -- ONE fp32 add each, as in the
  // forward sweep -- and reduce them with the reference's order (value, then class priority, then smaller delta).
  if (rank == 0) {
    float *s_bv = reinterpret_cast<float *>(v2_smem);          // [warps]
    int *s_bd = reinterpret_cast<int *>(v2_smem) + kV2Warps;   // [warps]
    int *s_pos = s_bd + kV2Warps;                              // [1]
    __syncthreads();
    int code = DAGB200_ST_OK;
    int pos = O - 1;
    if (!(__ldcg(lat + (int64_t)(Tn - 1) * L + O - 1) > ninf)) {
      code = DAGB200_ST_NO_PATH;
    } else {
      for (int i = Tn - 1; i >= 1; i--) {
        if (threadIdx.x == 0) prow[pos] = i;
        const float *prev = lat + (int64_t)(i - 1) * L;
        const int dmax = min(pos, Tl);
        float bv = ninf; int bd = 0;
        for (int d = dmax - (int)threadIdx.x; d >= 1; d -= kV2Threads) {    // descending delta inside a thread
          const int src = pos - d;
          const float x = __ldcg(prev + src) + __ldg(E + (int64_t)src * Tl + d - 1);
          if (better(x, d, bv, bd)) { bv = x; bd = d; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int od = __shfl_xor_sync(0xffffffffu, bd, o);
          if (better(ov, od, bv, bd)) { bv = ov; bd = od; }
        }
        if (lane == 0) { s_bv[warp] = bv; s_bd[warp] = bd; }
        __syncthreads();
        if (warp == 0) {
          bv = (lane < kV2Warps) ? s_bv[lane] : ninf;
          bd = (lane < kV2Warps) ? s_bd[lane] : 0;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int od = __shfl_xor_sync(0xffffffffu, bd, o);
            if (better(ov, od, bv, bd)) { bv = ov; bd = od; }
          }
          if (lane == 0) s_pos[0] = (bv > ninf) ? bd : 0;
        }
        __syncthreads();
        const int d = s_pos[0];
        if (d == 0) { code = DAGB200_ST_NO_PATH; break; }
        pos -= d;
      }
      if (code == DAGB200_ST_OK && threadIdx.x == 0) prow[pos] = 0;
    }
    if (status && threadIdx.x == 0) status[b] = code;
  }
}

size_t vitc_smem_bytes() {
  return sizeof(float) * ((size_t)kVcTpw * kVB * kVB + 2 * (size_t)kVcTpw * kVB * kV2Pitch + (size_t)kV2Warps * 2 * kVB * kVB) + 16;
}

size_t vit2_smem_bytes(int M, int L) {
  const int NB = (L + kVB - 1) / kVB;
  return sizeof(float) * ((size_t)kV2Tpw * kVB * kVB + 3 * (size_t)kV2Tpw * kVB * kV2Pitch + (size_t)kV2Warps * 16 * kVB) +
         (size_t)M * NB + 16;
}
bool vit2_supported(int M, int L) { return vit2_smem_bytes(M, L) <= 200 * 1024; }

int launch_viterbi_blocked(const float *match, const float *links, const int64_t *olen, const int64_t *tlen,
                           float *lattice, uint16_t *trace, int32_t *path, unsigned char *flags, int B, int M, int L,
                           int Tl, int32_t *status, cudaStream_t st) {
  const int NB = (L + kVB - 1) / kVB;
  static const int vit_version = getenv("DAGB200_VIT") ? atoi(getenv("DAGB200_VIT")) : 3;
  if (vit_version >= 3 && flags && 2 * (int64_t)B <= 2147483647) {
    const size_t smemc = vitc_smem_bytes();
    cudaFuncSetAttribute(dag_viterbi_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemc);
    prof_mark(6, st);
    dag_viterbi_cluster_kernel<<<2 * B, kV2Threads, smemc, st>>>(match, links, olen, tlen, lattice, path, flags, M, L, Tl, NB,
                                                                status);
    DAGB200_CHECK_LAUNCH("dag_viterbi_cluster_kernel");
    prof_mark(7, st);
    return 0;
  }
  const size_t smem = vit2_smem_bytes(M, L);
  cudaFuncSetAttribute(dag_viterbi_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  prof_mark(6, st);
  dag_viterbi_blocked_kernel<<<B, kV2Threads, smem, st>>>(match, links, olen, tlen, lattice, trace, path, M, L, Tl, NB, status);
  DAGB200_CHECK_LAUNCH("dag_viterbi_blocked_kernel");
  prof_mark(7, st);
  return 0;
}

}  // namespace dagb200
