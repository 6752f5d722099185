// dag_viterbi3.cu -- blocked max-plus (Viterbi) recurrence, wave-pipelined, for sm_100a (fp32, config 1).
//
// Replaces calculate_maxalpha_kernel + calculate_backtrace_kernel (reference dag_best_alignment.cu:39-130, 170-185).
// Arithmetic is the reference's, operation for operation: a candidate is ONE fp32 add (previous value + transition),
// the cell value one more add (+ emission); `max` is exact in any order, so the lattice values are bit-identical.  The
// forward sweep keeps no back-pointers: the arg-max (with the reference's tie-break: value, then the bit-reversed
// priority 0,2,1,3 of (delta-1) mod 4, then the smaller delta) is recomputed during the backtrace for the cells on the
// path only.
//
// Organisation.  Vertices in blocks of 32, target rows in chunks of 32, tile = (chunk c, block J).  A pass covers 8
// chunks (256 rows); the chunks of a pass are split by parity between the TWO CTAs of a thread-block cluster (one
// cluster per utterance).  Tiles are processed in anti-diagonal steps (step = J + local chunk), ONE cluster barrier per
// step.  Inside a step two kinds of warps run concurrently:
//   * chain warps (4 per CTA, one per chunk, lanes = rows): the tile of the CURRENT step.  For column cj the lane takes
//     max(far sums, what the row above hands down) + emission, then computes what its own row hands to the row below:
//     the max over the 32 sources of the PREVIOUS block (its own values of the step before, still in registers) and the
//     already finished columns of this block.  The last row's hand-down goes to the chunk below through global memory.
//   * far warps (12 per CTA, lanes = destination columns): the far-predecessor maxima of the NEXT step's tiles -- all
//     source blocks up to J-2, which were finished at least one step earlier, so nothing inside a step depends on the
//     chain warps.  A unit = (tile, source block): the transition column of the lane's vertex lives in 32 registers, the
//     32x32 tile of previous-row values is staged by cp.async (double-buffered), 32 running maxima stay in registers
//     across the units of a tile and meet the other warps' maxima in shared memory (integer atomicMax on
//     order-preserving keys) once per tile.  Units are dealt in contiguous ranges, the same number to every far warp.
// When the caller does not ask for the lattice (the Python wrapper never does, dag_loss.py:227-230) only cells that can
// still reach the end cell are computed: column j of row t is skipped when O-1-j < Tn-1-t.
#include <cstdlib>

#include "common.cuh"

namespace dagb200 {
namespace v3 {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kChain = 4;                 // chain warps per CTA = chunks of a pass owned by a CTA
constexpr int kFar = kWarps - kChain;     // far warps per CTA
constexpr int kB = 32;                    // block / chunk edge
constexpr int kPitch = 33;
constexpr int kPassChunks = 2 * kChain;   // chunks per pass (both CTAs)

__device__ __forceinline__ int f2key(float x) { const int b = __float_as_int(x); return b ^ ((b >> 31) & 0x7fffffff); }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(float *smem_dst, const float *gsrc) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ int rank4(int delta) {  // priority of the class of `delta`: classes 0,2,1,3 -> 0,1,2,3
  const int cl = (delta - 1) & 3;
  return ((cl & 1) << 1) | (cl >> 1);
}
__device__ __forceinline__ bool better(float v1, int d1, float v2, int d2) {   // (-inf never wins)
  if (v1 > v2) return true;
  if (v1 < v2 || !(v1 > neg_inf_f())) return false;
  const int r1 = rank4(d1), r2 = rank4(d2);
  return r1 < r2 || (r1 == r2 && d1 < d2);
}

struct Geo {
  int O, Tn, L, Tl, NBv, NCv, nsteps, band;
  bool full;
  // last block of chunk c that holds a cell still able to reach the end (plus the column the row below needs)
  __device__ __forceinline__ int jhi(int c) const {
    if (full) return NBv - 1;
    const int tmax = min(kB * c + kB, Tn - 1);
    const int hi = O - 1 - (Tn - 1 - tmax) + 1;
    return min(NBv - 1, hi / kB);
  }
  __device__ __forceinline__ bool tile_alive(int c, int J) const { return c < NCv && J >= c && J <= jhi(c); }
  // far sources of tile (c, J): blocks Ilo .. J-2
  __device__ __forceinline__ int ilo(int c, int J) const { return max(c, J - band); }
};

// ---- chain warp: one column of the tile (lanes = rows); everything indexed by compile-time CJ -------------------------
template <int CJ>
__device__ __forceinline__ void chain_column(float (&vrow)[kB], const float (&vprev)[kB], const float *ed, const float *ep,
                                             const float *iow, const int *xkw, const float *handin, float *handout,
                                             int lane, bool rowvalid, int jbase, int t, int O, int jmax, bool prev_on) {
  const float ninf = neg_inf_f();
  // what the row above hands down for this column; mine is computed below and travels one lane down
  float n0 = ninf, n1 = ninf;
  if (prev_on) {
#pragma unroll
    for (int c4 = 0; c4 < kB; c4 += 4) {
      const float4 e4 = *reinterpret_cast<const float4 *>(ep + CJ * kB + c4);
      n0 = max3(n0, vprev[c4 + 0] + e4.x, vprev[c4 + 1] + e4.y);
      n1 = max3(n1, vprev[c4 + 2] + e4.z, vprev[c4 + 3] + e4.w);
    }
  }
#pragma unroll
  for (int c4 = 0; c4 < CJ; c4 += 4) {
    const float4 e4 = *reinterpret_cast<const float4 *>(ed + CJ * kB + c4);
    if (c4 + 0 < CJ) n0 = fmaxf(n0, vrow[c4 + 0] + e4.x);
    if (c4 + 1 < CJ) n1 = fmaxf(n1, vrow[c4 + 1] + e4.y);
    if (c4 + 2 < CJ) n0 = fmaxf(n0, vrow[c4 + 2] + e4.z);
    if (c4 + 3 < CJ) n1 = fmaxf(n1, vrow[c4 + 3] + e4.w);
  }
  const float n = fmaxf(n0, n1);
  float rv = __shfl_up_sync(0xffffffffu, n, 1);
  if (lane == 0) rv = handin[CJ];
  if (lane == kB - 1) handout[CJ] = n;
  const float best = fmaxf(rv, key2f(xkw[CJ]));
  const int j = jbase + CJ;
  const bool valid = rowvalid && j >= t && j < O && j <= jmax;
  vrow[CJ] = valid ? best + iow[CJ] : ninf;
}
template <int CJ0>
__device__ __forceinline__ void chain_group(float (&vrow)[kB], const float (&vprev)[kB], const float *ed, const float *ep,
                                            const float *iow, const int *xkw, const float *handin, float *handout, int lane,
                                            bool rowvalid, int jbase, int t, int O, int jmax, bool prev_on) {
  chain_column<CJ0 + 0>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 1>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 2>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 3>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 4>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 5>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 6>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
  chain_column<CJ0 + 7>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, jbase, t, O, jmax, prev_on);
}

// stage the operands of the chain tile (c, J): diagonal block [cj][ci], previous-block transitions [cj][ci] (source
// block J-1), emissions [row][cj]; lanes = destination columns, 4-byte asynchronous copies (any Tl / alignment)
__device__ __forceinline__ void chain_stage(const Geo &g, const float *__restrict__ m, const float *__restrict__ E, float *ed,
                                            float *ep, float *io, int c, int J, int lane) {
  const float ninf = neg_inf_f();
  const int j = kB * J + lane;
  const bool jok = j < g.O;
#pragma unroll 4
  for (int rr = 0; rr < kB; rr++) {
    {   // diagonal block: source vertex 32J + rr, destination j
      const int i = kB * J + rr, k = lane - rr - 1;
      if (k >= 0 && k < g.Tl && jok) cp_async4(ed + lane * kB + rr, E + (int64_t)i * g.Tl + k);   // i < j < O
      else ed[lane * kB + rr] = ninf;
    }
    {   // previous block: source vertex 32(J-1) + rr
      const int i = kB * (J - 1) + rr, k = kB + lane - rr - 1;
      if (J > 0 && k < g.Tl && jok) cp_async4(ep + lane * kB + rr, E + (int64_t)i * g.Tl + k);
      else ep[lane * kB + rr] = ninf;
    }
    {   // emissions of row t = 32c + 1 + rr
      const int s = c * kB + rr;
      if (s < g.nsteps && j < g.L) cp_async4(io + rr * kPitch + lane, m + (int64_t)(1 + s) * g.L + j);
      else io[rr * kPitch + lane] = ninf;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dag_viterbi_wave_kernel(const float *__restrict__ match, const float *__restrict__ links,
                        const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen, float *lattice,
                        int32_t *__restrict__ path, float *hand_g, int M, int L, int Tl, int NB, int full_lattice,
                        int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char v3_smem[];
  const int b = blockIdx.x >> 1, rank = blockIdx.x & 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t latsz = (int64_t)M * L;
  const float ninf = neg_inf_f();
  float *lat = lattice + b * latsz;
  int32_t *prow = path + (int64_t)b * L;
  const float *m = match + b * latsz;
  const float *E = links + (int64_t)b * L * Tl;
  float *HG = hand_g + (size_t)b * kPassChunks * NB * kB;     // [consumer chunk slot][vertex]
  if (rank == 0)
    for (int j = threadIdx.x; j < L; j += kThreads) prow[j] = -1;

  int st = DAGB200_ST_OK;
  if (Tn < 2 || O < 2) st = DAGB200_ST_LEN_LT2;
  else if (O < Tn || O > L || Tn > M) st = DAGB200_ST_GRAPH_SMALL;
  else if ((int64_t)(Tn - 1) * Tl + 1 < O) st = DAGB200_ST_TOO_SHORT;
  if (st != DAGB200_ST_OK) {   // both CTAs of the cluster leave together: no barrier is pending
    if (rank == 0) {
      if (full_lattice)
        for (int64_t x = threadIdx.x; x < latsz; x += kThreads) lat[x] = ninf;
      if (status && threadIdx.x == 0) status[b] = st;
    }
    return;
  }

  Geo g;
  g.O = O; g.Tn = Tn; g.L = L; g.Tl = Tl;
  g.NBv = (O + kB - 1) / kB;
  g.nsteps = Tn - 1;
  g.NCv = (g.nsteps + kB - 1) / kB;
  g.band = 1 + (Tl - 1) / kB;
  g.full = full_lattice != 0;
  const int NP = (g.NCv + kPassChunks - 1) / kPassChunks;

  // shared memory
  float *s_ed, *s_ep, *s_io, *s_hand, *s_vs;
  int *s_xk;
  {
    float *p = reinterpret_cast<float *>(v3_smem);
    s_ed = p;   p += kChain * kB * kB;                 // [tile][cj][ci]
    s_ep = p;   p += kChain * kB * kB;                 // [tile][cj][ci]
    s_io = p;   p += kChain * kB * kPitch;             // [tile][row][cj] emissions in, cell values out
    s_xk = reinterpret_cast<int *>(p);  p += 2 * kChain * kB * kPitch;   // [step parity][tile][row][cj] far maxima (keys)
    s_hand = p; p += kChain * 2 * kB;                  // [tile][in | out][cj]
    s_vs = p;                                          // [far warp][2][32][32] previous-row values of a unit
  }

  // ---- prologue (split between the two CTAs) ----------------------------------------------------------------
  for (int x = threadIdx.x; x < 2 * kChain * kB * kPitch; x += kThreads) s_xk[x] = f2key(ninf);
  if (g.full) {
    // everything outside the computed tiles is -inf; the tiles overwrite their part after the barrier below
    const int64_t half = (latsz + 1) / 2;
    const int64_t x0 = rank * half, x1 = min(latsz, x0 + half);
    for (int64_t x = x0 + threadIdx.x; x < x1; x += kThreads) lat[x] = ninf;
  }
  cluster_sync_all();
  if (rank == 0) {
    // seed row (t = 0): only vertex 0 carries a value; what it hands to row 1 is one add per column
    const float v0 = m[0];
    for (int j = threadIdx.x; j < g.NBv * kB; j += kThreads) {
      if (j < L) lat[j] = (j == 0) ? v0 : ninf;
      HG[j] = (j >= 1 && j < O && j - 1 < Tl) ? v0 + E[j - 1] : ninf;
    }
  }
  cluster_sync_all();

  const bool is_chain = warp < kChain;
  const int fw = warp - kChain;                       // far warp index
  float *vs0 = s_vs + (size_t)(fw < 0 ? 0 : fw) * 2 * kB * kB;
  const bool vec_ok = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(lat) & 15) == 0);

  for (int p = 0; p < NP; p++) {
    const int cpass = min(kPassChunks, g.NCv - kPassChunks * p);    // chunks of this pass
    const int nst = g.NBv + cpass - 1;                               // anti-diagonal steps
    float vprev[kB];
#pragma unroll
    for (int k = 0; k < kB; k++) vprev[k] = ninf;
    bool staged = false;
    // tiles of this pass need J >= c >= 8p: the steps before 8p - 1 are empty
    for (int sg = kPassChunks * p - 1; sg < nst; sg++) {
      if (is_chain) {
        // ======================= chain warp: tile (c, J = sg - lc) of this step ==================================
        const int ts = warp, lc = 2 * ts + rank, c = kPassChunks * p + lc;
        const int J = sg - lc;
        float *ed = s_ed + ts * kB * kB, *ep = s_ep + ts * kB * kB, *io = s_io + ts * kB * kPitch;
        float *handin = s_hand + ts * 2 * kB, *handout = handin + kB;
        int *xk = s_xk + ((sg & 1) * kChain + ts) * kB * kPitch;
        if (sg >= 0 && lc < cpass && g.tile_alive(c, J)) {
          if (!staged) chain_stage(g, m, E, ed, ep, io, c, J, lane);
          handin[lane] = __ldcg(HG + (size_t)lc * NB * kB + kB * J + lane);     // posted by the chunk above, an earlier step
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncwarp();
          const int s = c * kB + lane;
          const bool rowvalid = s < g.nsteps;
          const int t = 1 + s;
          const int jmax = g.full ? L : (O - 1 - (Tn - 1 - t));
          const float *iow = io + lane * kPitch;
          const int *xkw = xk + lane * kPitch;
          const bool prev_on = J > c;                // the previous block holds cells of this chunk
          float vrow[kB];
          chain_group<0>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, kB * J, t, O, jmax, prev_on);
          chain_group<8>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, kB * J, t, O, jmax, prev_on);
          chain_group<16>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, kB * J, t, O, jmax, prev_on);
          chain_group<24>(vrow, vprev, ed, ep, iow, xkw, handin, handout, lane, rowvalid, kB * J, t, O, jmax, prev_on);
          __syncwarp();
          // reset the far maxima for the step after next; cell values -> staging tile -> lattice rows (coalesced)
#pragma unroll
          for (int k = 0; k < kB; k++) {
            xk[lane * kPitch + k] = f2key(ninf);
            io[lane * kPitch + k] = vrow[k];
            vprev[k] = vrow[k];
          }
          __syncwarp();
          {
            const int rl = min(kB, g.nsteps - c * kB);
            const int j = kB * J + lane;
            if (j < L)
              for (int rr = 0; rr < rl; rr++) __stcg(lat + (int64_t)(1 + c * kB + rr) * L + j, io[rr * kPitch + lane]);
            // what my last row hands to the chunk below (next local chunk: the other CTA; last chunk: next pass)
            const int slot = (lc + 1) % kPassChunks;
            __stcg(HG + (size_t)slot * NB * kB + kB * J + lane, handout[lane]);
          }
          __syncwarp();
          // operands of my next tile
          staged = false;
          if (g.tile_alive(c, J + 1)) { chain_stage(g, m, E, ed, ep, io, c, J + 1, lane); staged = true; }
        }
      } else {
        // ======================= far warps: far maxima of the NEXT step's tiles ===================================
        const int sn = sg + 1;
        if (sn < nst) {
          int nun[kChain], total = 0;
#pragma unroll
          for (int ts = 0; ts < kChain; ts++) {
            const int lc = 2 * ts + rank, c = kPassChunks * p + lc, J = sn - lc;
            nun[ts] = (lc < cpass && g.tile_alive(c, J)) ? max(0, J - 2 - g.ilo(c, J) + 1) : 0;
            total += nun[ts];
          }
          const int u0 = (int)(((int64_t)total * fw) / kFar), u1 = (int)(((int64_t)total * (fw + 1)) / kFar);
          if (u1 > u0) {
            auto decode = [&](int u, int &ts, int &I) {
              ts = 0;
              int r = u;
#pragma unroll
              for (int x = 0; x < kChain - 1; x++)
                if (ts == x && r >= nun[x]) { r -= nun[x]; ts = x + 1; }
              const int lc = 2 * ts + rank, c = kPassChunks * p + lc, J = sn - lc;
              I = g.ilo(c, J) + r;
            };
            auto stage = [&](int u, float *slab) {
              int ts, I;
              decode(u, ts, I);
              const int c = kPassChunks * p + 2 * ts + rank;
              const int tp0 = c * kB;                       // previous-row index of the tile's first row
              if (vec_ok) {
                const int q = lane & 7;
                for (int rr = lane >> 3; rr < kB; rr += 4) {
                  float *dst = slab + rr * kB + 4 * q;
                  if (tp0 + rr < g.nsteps) cp_async16_cg(dst, lat + (int64_t)(tp0 + rr) * L + kB * I + 4 * q);
                  else *reinterpret_cast<float4 *>(dst) = make_float4(ninf, ninf, ninf, ninf);
                }
              } else {
                const float *lp = lat + (int64_t)tp0 * L + kB * I + lane;       // 32 I + lane < 32 J <= O - 1 < L
                for (int rr = 0; rr < kB; rr++) slab[rr * kB + lane] = (tp0 + rr < g.nsteps) ? __ldcg(lp + (int64_t)rr * L) : ninf;
              }
              asm volatile("cp.async.commit_group;" ::: "memory");
            };
            float acc[kB];
#pragma unroll
            for (int k = 0; k < kB; k++) acc[k] = ninf;
            int cur_ts = -1;
            auto flush = [&](int ts) {
              int *xk = s_xk + ((sn & 1) * kChain + ts) * kB * kPitch + lane;
#pragma unroll
              for (int k = 0; k < kB; k++) {
                if (acc[k] > ninf) atomicMax(xk + k * kPitch, f2key(acc[k]));
                acc[k] = ninf;
              }
            };
            int buf = 0;
            stage(u0, vs0);
            for (int u = u0; u < u1; u++) {
              const bool more = u + 1 < u1;
              if (more) stage(u + 1, vs0 + (buf ^ 1) * kB * kB);
              int ts, I;
              decode(u, ts, I);
              if (ts != cur_ts) { if (cur_ts >= 0) flush(cur_ts); cur_ts = ts; }
              const int J = sn - (2 * ts + rank);
              const int j = kB * J + lane;
              float ecol[kB];
              {
                const int k0 = j - kB * I - 1;                // transition index from the block's first source vertex; >= 32
                const float *epn = E + (int64_t)(kB * I) * Tl + k0;
                const bool jok = j < O;
#pragma unroll
                for (int ii = 0; ii < kB; ii++) ecol[ii] = (jok && k0 - ii < Tl) ? __ldg(epn + (int64_t)ii * (Tl - 1)) : ninf;
              }
              if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
              else asm volatile("cp.async.wait_group 0;" ::: "memory");
              __syncwarp();
              const float *vsw = vs0 + buf * kB * kB;
#pragma unroll
              for (int rr = 0; rr < kB; rr++) {
                float b0 = acc[rr], b1 = ninf;
#pragma unroll
                for (int c4 = 0; c4 < kB; c4 += 4) {
                  const float4 a4 = *reinterpret_cast<const float4 *>(vsw + rr * kB + c4);
                  b0 = max3(b0, a4.x + ecol[c4 + 0], a4.y + ecol[c4 + 1]);
                  b1 = max3(b1, a4.z + ecol[c4 + 2], a4.w + ecol[c4 + 3]);
                }
                acc[rr] = fmaxf(b0, b1);
              }
              __syncwarp();
              buf ^= 1;
            }
            flush(cur_ts);
          }
        }
      }
      cluster_sync_all();      // the step is complete in both CTAs and visible to both
    }
  }

  // ---- backtrace (dag_best_alignment.cu:178-184) with the back-pointers recomputed on the way: for the cell (i, pos) of
  // the path all threads of CTA 0 score its candidates delta = 1 .. min(pos, Tl) that hold a lattice cell (source
  // vertex >= i-1; the others are -inf) -- ONE fp32 add each, as in the forward sweep -- and reduce them with the
  // reference's order (value, then class priority, then smaller delta).
  if (rank == 0) {
    float *s_bv = reinterpret_cast<float *>(v3_smem);
    int *s_bd = reinterpret_cast<int *>(v3_smem) + kWarps;
    int *s_pos = s_bd + kWarps;
    __syncthreads();
    int code = DAGB200_ST_OK;
    int pos = O - 1;
    if (!(__ldcg(lat + (int64_t)(Tn - 1) * L + O - 1) > ninf)) {
      code = DAGB200_ST_NO_PATH;
    } else {
      for (int i = Tn - 1; i >= 1; i--) {
        if (threadIdx.x == 0) prow[pos] = i;
        const float *prev = lat + (int64_t)(i - 1) * L;
        const int dmax = min(min(pos, Tl), pos - (i - 1));
        float bv = ninf; int bd = 0;
        for (int d = dmax - (int)threadIdx.x; d >= 1; d -= kThreads) {    // descending delta inside a thread
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
          bv = (lane < kWarps) ? s_bv[lane] : ninf;
          bd = (lane < kWarps) ? s_bd[lane] : 0;
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

}  // namespace v3

size_t vit3_smem_bytes() {
  using namespace v3;
  return sizeof(float) * ((size_t)2 * kChain * kB * kB + (size_t)3 * kChain * kB * kPitch + kChain * 2 * kB +
                          (size_t)kFar * 2 * kB * kB) + 16;
}
size_t vit3_hand_bytes(int B, int L) {
  const int NB = (L + 31) / 32;
  return (((size_t)B * v3::kPassChunks * NB * 32 * sizeof(float)) + 255) & ~(size_t)255;
}

int launch_viterbi_wave(const float *match, const float *links, const int64_t *olen, const int64_t *tlen, float *lattice,
                        int32_t *path, float *hand_g, int full_lattice, int B, int M, int L, int Tl, int32_t *status,
                        cudaStream_t st) {
  using namespace v3;
  const int NB = (L + kB - 1) / kB;
  const size_t smem = vit3_smem_bytes();
  cudaFuncSetAttribute(dag_viterbi_wave_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  prof_mark(6, st);
  dag_viterbi_wave_kernel<<<2 * B, kThreads, smem, st>>>(match, links, olen, tlen, lattice, path, hand_g, M, L, Tl, NB,
                                                        full_lattice, status);
  DAGB200_CHECK_LAUNCH("dag_viterbi_wave_kernel");
  prof_mark(7, st);
  return 0;
}

}  // namespace dagb200
