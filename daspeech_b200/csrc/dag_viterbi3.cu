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
// chunks (256 rows); the chunks of a pass are split {0,3,4,7} / {1,2,5,6} between the TWO CTAs of a thread-block
// cluster (one cluster per utterance; both CTAs then own the same number of far units in every step).  Tiles are
// processed in anti-diagonal steps (step = J + local chunk), ONE cluster barrier per step.  Per CTA and step:
//   * a team of four warps per tile (the chain warp + three helpers) first adds the PREVIOUS block J-1 -- the rows the
//     chain warp produced one step earlier, kept in shared memory, plus the last row of the chunk above -- to the far
//     maxima, eight rows per warp, exactly like a far unit (lanes = destination columns).  One named barrier later
//     the chain warp (lanes = rows) sweeps the 32 columns of the diagonal block in push form: cell = max(far maxima,
//     what the row above hands down) + emission, then the cell pushes into what its own row hands to the row below for
//     the later columns -- one add and one max on the column-to-column dependency path.  The last row's hand-down goes
//     to the chunk below through global memory.  Meanwhile one helper stages the operands of the team's next tile
//     (double-buffered), so no load latency sits at the head of a step.
//   * all sixteen warps then pull far units from a shared-memory queue: the far-predecessor maxima of the NEXT step's
//     tiles from source blocks up to J-2, which were finished at least one step earlier -- nothing inside a step depends
//     on the chain warps.  A unit = (tile, half a source block): the transition column of the lane's vertex lives in 16
//     registers, the 32 x 16 tile of previous-row values is staged by cp.async (double-buffered) and read as broadcast
//     float4s one row ahead of the arithmetic, 32 running maxima stay in registers across the units of a tile and meet
//     the other warps' maxima in shared memory (integer atomicMax on order-preserving keys) once per tile.
// Cost model (tools/ubench3.cu, ubench4.cu on B200): one candidate = one FADD + half an FMNMX3, and FMNMX3 issues at
// half rate, i.e. 2 issue slots per edge whatever the idiom (float or integer min/max); a broadcast LDS.128 issues at
// most once per ~15 cycles per warp.  The kernel is issue-bound, not HBM-bound: DESIGN.md section 4.6.
// When the caller does not ask for the lattice (the Python wrapper never does, dag_loss.py:227-230) only cells that can
// still reach the end cell are computed: column j of row t is skipped when O-1-j < Tn-1-t.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace dagb200 {
namespace v3 {

#ifdef DAGB200_V3_TIMING
__device__ long long g_v3t[64][8];   // per step: [0] chain busy, [1] far busy (warp 4), [2] step length (warp 4), [3] units of warp 4, [4] far busy warp 15
#endif

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kChain = 4;                 // chain warps per CTA = chunks of a pass owned by a CTA
constexpr int kHelp = 3;                  // helper warps per tile team
constexpr int kHalf = 16;                 // source vertices per far unit
constexpr int kB = 32;                    // block / chunk edge
constexpr int kPitch = 33;
constexpr int kEdPitch = 36;               // 16-byte aligned rows, 4 banks apart
constexpr int kIoPitch = 36;
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

// local chunk of tile slot ts in CTA `rank`: {0,3,4,7} / {1,2,5,6} -- both CTAs own the same number of far units in
// every step (a tile of local chunk lc has step - 2 lc - 1 of them)
__device__ __forceinline__ int chunk_of(int ts, int rank) { return 2 * ts + ((ts ^ rank) & 1); }

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

// ---- chain warp, column sweep (lanes = rows), push formulation: hd[k] = what my row hands to the row below for column
// k from the columns finished so far; the cell of column CJ pushes into hd[CJ+1 ..].  The only operations on the
// column-to-column dependency path are one add + one max (hd[CJ+1]), the shuffle, one max and one add.
template <int CJ>
__device__ __forceinline__ void chain_column(float (&hd)[kB], const float *ed, float *iow, const float (&mm)[8],
                                             const float (&fx)[8], const float (&hin)[8], const float (&e1)[8],
                                             float *handout, int lane, uint32_t vmask) {
  const float ninf = neg_inf_f();
  const float n = hd[CJ];
  float rv = __shfl_up_sync(0xffffffffu, n, 1);
  if (lane == 0) rv = hin[CJ & 7];
  if (lane == kB - 1) handout[CJ] = n;
  const float best = fmaxf(rv, fx[CJ & 7]);
  const float v = ((vmask >> CJ) & 1u) ? best + mm[CJ & 7] : ninf;
  if (CJ + 1 < kB) hd[(CJ + 1) & (kB - 1)] = fmaxf(hd[(CJ + 1) & (kB - 1)], v + e1[CJ & 7]);   // the next column first
  iow[CJ] = v;
  // push into the later columns: ed[CJ][k] = transition from vertex CJ of this block to vertex k (source-major tile)
#pragma unroll
  for (int k4 = (CJ + 2) & ~3; k4 < kB; k4 += 4) {
    const float4 e4 = *reinterpret_cast<const float4 *>(ed + CJ * kEdPitch + k4);
    if (k4 + 0 > CJ + 1) hd[k4 + 0] = fmaxf(hd[k4 + 0], v + e4.x);
    if (k4 + 1 > CJ + 1) hd[k4 + 1] = fmaxf(hd[k4 + 1], v + e4.y);
    if (k4 + 2 > CJ + 1) hd[k4 + 2] = fmaxf(hd[k4 + 2], v + e4.z);
    if (k4 + 3 > CJ + 1) hd[k4 + 3] = fmaxf(hd[k4 + 3], v + e4.w);
  }
}
// eight columns: their emissions, far maxima, hand-ins and next-column transitions are fetched before the first needs them
template <int CJ0>
__device__ __forceinline__ void chain_group(float (&hd)[kB], const float *ed, float *iow, const int *xkw,
                                            const float *handin, float *handout, int lane, uint32_t vmask) {
  float mm[8], fx[8], hin[8], e1[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    mm[k] = iow[CJ0 + k];
    fx[k] = key2f(xkw[CJ0 + k]);
    hin[k] = handin[CJ0 + k];
    e1[k] = (CJ0 + k + 1 < kB) ? ed[(CJ0 + k) * kEdPitch + ((CJ0 + k + 1) & (kB - 1))] : neg_inf_f();
  }
  chain_column<CJ0 + 0>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 1>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 2>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 3>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 4>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 5>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 6>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
  chain_column<CJ0 + 7>(hd, ed, iow, mm, fx, hin, e1, handout, lane, vmask);
}

// stage the operands of the chain tile (c, J) asynchronously.  Diagonal block, source-major [ci][cj]: row ci holds the
// first 31 - ci transitions of links row 32 J + ci (one coalesced, conflict-free 4-byte copy per row and lane); emissions
// [row][cj]: 16-byte copies when the lattice rows are 16-byte aligned.
__device__ __forceinline__ void chain_stage(const Geo &g, const float *__restrict__ m, const float *__restrict__ E, float *ed,
                                            float *io, int c, int J, int lane, bool vec_m) {
  const float ninf = neg_inf_f();
  const int j = kB * J + lane;
  const bool jok = j < g.O;
  const float *erow = E + (int64_t)(kB * J) * g.Tl - 1;      // erow[ci * (Tl - 1) + lane] = E[32 J + ci][lane - ci - 1]
#pragma unroll 8
  for (int rr = 0; rr < kB; rr++) {
    const int k = lane - rr - 1;
    if (k >= 0 && k < g.Tl && jok) cp_async4(ed + rr * kEdPitch + lane, erow + (int64_t)rr * (g.Tl - 1) + lane);   // 32 J + rr < j < O
    else ed[rr * kEdPitch + lane] = ninf;
  }
  const int s0 = c * kB;
  if (vec_m && kB * J + kB <= g.L) {
    const int q = lane & 7;
#pragma unroll
    for (int rr = lane >> 3; rr < kB; rr += 4) {
      float *dst = io + rr * kIoPitch + 4 * q;
      if (s0 + rr < g.nsteps) cp_async16_cg(dst, m + (int64_t)(1 + s0 + rr) * g.L + kB * J + 4 * q);
      else *reinterpret_cast<float4 *>(dst) = make_float4(ninf, ninf, ninf, ninf);
    }
  } else {
#pragma unroll 8
    for (int rr = 0; rr < kB; rr++) {
      if (s0 + rr < g.nsteps && j < g.L) cp_async4(io + rr * kIoPitch + lane, m + (int64_t)(1 + s0 + rr) * g.L + j);
      else io[rr * kIoPitch + lane] = ninf;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ void team_sync(int ts) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + ts), "r"(32 * (1 + kHelp)) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dag_viterbi_wave_kernel(const float *__restrict__ match, const float *__restrict__ links,
                        const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen, float *lattice,
                        int32_t *__restrict__ path, float *hand_g, int M, int L, int Tl, int NB, int full_lattice,
                        int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char v3_smem[];
#ifdef DAGB200_V3_TIMING
  const long long tkernel0 = clock64();
#endif
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
  float *s_ed, *s_pv, *s_io, *s_hand, *s_vs;
  int *s_xk, *s_queue;
  {
    float *p = reinterpret_cast<float *>(v3_smem);
    s_ed = p;   p += 2 * kChain * kB * kEdPitch;       // [block parity][tile][ci][cj] diagonal transition block (source-major)
    s_io = p;   p += 2 * kChain * kB * kIoPitch;       // [block parity][tile][row][cj] emissions in, cell values out
    s_pv = p;   p += kChain * kB * kEdPitch;           // [tile][source row][ci] values of the previous block: row 0 = last row of
                                                       // the chunk above, rows 1..31 = my rows 0..30
    s_xk = reinterpret_cast<int *>(p);  p += 2 * kChain * kB * kPitch;   // [step parity][tile][row][cj] far maxima (keys)
    s_hand = p; p += kChain * 2 * kB;                  // [tile][in | out][cj]
    s_queue = reinterpret_cast<int *>(p);  p += 4;     // [step parity] next far unit
    s_vs = p;                                          // [warp][2][32 rows][16 sources] previous-row values of a unit
  }

  // ---- prologue (split between the two CTAs) ----------------------------------------------------------------
  for (int x = threadIdx.x; x < 2 * kChain * kB * kPitch; x += kThreads) s_xk[x] = f2key(ninf);
  if (threadIdx.x < 4) s_queue[threadIdx.x] = 0;
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

  // team of a tile slot: the chain warp (sub 0) and three helpers (sub 1..3)
  const int ts = warp < kChain ? warp : (warp - kChain) / kHelp;
  const int sub = warp < kChain ? 0 : (warp - kChain) % kHelp + 1;
  const int lc = chunk_of(ts, rank);
  float *vs0 = s_vs + (size_t)warp * 2 * kB * kHalf;
  const bool vec_ok = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(lat) & 15) == 0);
  const bool vec_m = ((L & 3) == 0) && ((reinterpret_cast<uintptr_t>(m) & 15) == 0);
  float *pv = s_pv + ts * kB * kEdPitch;
  float *handin = s_hand + ts * 2 * kB, *handout = handin + kB;

  for (int p = 0; p < NP; p++) {
    const int cpass = min(kPassChunks, g.NCv - kPassChunks * p);    // chunks of this pass
    const int nst = g.NBv + cpass - 1;                               // anti-diagonal steps
    const int c = kPassChunks * p + lc;
    // tiles of this pass need J >= c >= 8p: the steps before 8p - 1 are empty
    for (int sg = kPassChunks * p - 1; sg < nst; sg++) {
#ifdef DAGB200_V3_TIMING
      const long long tstep0 = clock64();
      const bool tlog = blockIdx.x == 0 && lane == 0 && sg >= 0 && sg < 64;
#endif
      const int J = sg - lc;
      if (sg >= 0 && lc < cpass && g.tile_alive(c, J)) {
        // ======================= the team's tile (c, J) of this step ================================================
        float *ed = s_ed + ((J & 1) * kChain + ts) * kB * kEdPitch, *io = s_io + ((J & 1) * kChain + ts) * kB * kIoPitch;
        int *xk = s_xk + ((sg & 1) * kChain + ts) * kB * kPitch;
        if (sub == 1 && J == c) chain_stage(g, m, E, ed, io, c, J, lane, vec_m);       // first tile of the chunk: nothing staged yet
        if (sub == 0) handin[lane] = __ldcg(HG + (size_t)lc * NB * kB + kB * J + lane);  // posted by the chunk above, an earlier step
        if (J > c) {
          // ---- phase A (lanes = destination columns, eight rows per warp of the team): the previous block J-1 as one
          // more far unit.  Its values are the rows the chain warp produced in the step before plus the last row of the
          // chunk above (row 0, fetched by the warp that uses it).
          const int j = kB * J + lane;
          float ecol[kB];
          {
            const int k0 = kB + lane - 1;              // transition index from the previous block's first vertex
            const float *epn = E + (int64_t)(kB * (J - 1)) * Tl + k0;
            const bool jok = j < O;
#pragma unroll
            for (int ii = 0; ii < kB; ii++) ecol[ii] = (jok && k0 - ii < Tl) ? __ldg(epn + (int64_t)ii * (Tl - 1)) : ninf;
          }
          if (sub == 0) {
            pv[lane] = __ldcg(lat + (int64_t)(c * kB) * L + kB * (J - 1) + lane);
            __syncwarp();
          }
          int *xkc = xk + (8 * sub) * kPitch + lane;
          const float *pvr = pv + (8 * sub) * kEdPitch;
#pragma unroll
          for (int rr = 0; rr < 8; rr++) {
            float b0 = key2f(xkc[rr * kPitch]), b1 = ninf;
#pragma unroll
            for (int c4 = 0; c4 < kB; c4 += 4) {
              const float4 a4 = *reinterpret_cast<const float4 *>(pvr + rr * kEdPitch + c4);
              b0 = max3(b0, a4.x + ecol[c4 + 0], a4.y + ecol[c4 + 1]);
              b1 = max3(b1, a4.z + ecol[c4 + 2], a4.w + ecol[c4 + 3]);
            }
            xkc[rr * kPitch] = f2key(fmaxf(b0, b1));
          }
        }
        if (sub == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");     // this tile's operands (staged one step ago)
        team_sync(ts);
#ifdef DAGB200_V3_TIMING
        { long long tq = clock64(); if (tlog && warp == 0) g_v3t[sg][3] = tq - tstep0; }
#endif
        if (sub == 0) {
          // ---- phase B (lanes = rows): column sweep over the diagonal block
          const int s = c * kB + lane;
          const int t = 1 + s;
          const int jmax = g.full ? L : (O - 1 - (Tn - 1 - t));
          uint32_t vmask = 0;
          {
            const int lo = max(t, kB * J) - kB * J, hi = min(min(O - 1, jmax), kB * J + kB - 1) - kB * J;
            if (s < g.nsteps && hi >= lo) vmask = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
          }
          float *iow = io + lane * kIoPitch;
          const int *xkw = xk + lane * kPitch;
          {
            float hd[kB];
#pragma unroll
            for (int k = 0; k < kB; k++) hd[k] = ninf;
            chain_group<0>(hd, ed, iow, xkw, handin, handout, lane, vmask);
            chain_group<8>(hd, ed, iow, xkw, handin, handout, lane, vmask);
            chain_group<16>(hd, ed, iow, xkw, handin, handout, lane, vmask);
            chain_group<24>(hd, ed, iow, xkw, handin, handout, lane, vmask);
          }
          __syncwarp();
#ifdef DAGB200_V3_TIMING
          { long long tq = clock64(); if (tlog && warp == 0) g_v3t[sg][7] = tq - tstep0; }
#endif
          // ---- phase C: reset the far maxima for the step after next; my rows become source rows 1..31 of the next
          // block's phase A; cell values -> lattice rows (coalesced)
#pragma unroll
          for (int k = 0; k < kB; k++) xk[lane * kPitch + k] = f2key(ninf);
          if (lane < kB - 1) {
#pragma unroll
            for (int k = 0; k < kB; k += 4)
              *reinterpret_cast<float4 *>(pv + (lane + 1) * kEdPitch + k) = *reinterpret_cast<const float4 *>(iow + k);
          }
          {
            const int rl = min(kB, g.nsteps - c * kB);
            const int j = kB * J + lane;
            if (j < L) {
#pragma unroll 8
              for (int rr = 0; rr < kB; rr++)
                if (rr < rl) __stcg(lat + (int64_t)(1 + c * kB + rr) * L + j, io[rr * kIoPitch + lane]);
            }
            // what my last row hands to the chunk below (next local chunk: the other CTA or this one; last chunk: next pass)
            const int slot = (lc + 1) % kPassChunks;
            __stcg(HG + (size_t)slot * NB * kB + kB * J + lane, handout[lane]);
          }
          __syncwarp();
#ifdef DAGB200_V3_TIMING
          { long long tq = clock64(); if (tlog && warp == 0) g_v3t[sg][1] = tq - tstep0; }
#endif
        } else if (sub == 1) {
          // operands of the team's next tile into the other buffer, while the chain warp sweeps this one
          if (g.tile_alive(c, J + 1))
            chain_stage(g, m, E, s_ed + (((J + 1) & 1) * kChain + ts) * kB * kEdPitch,
                        s_io + (((J + 1) & 1) * kChain + ts) * kB * kIoPitch, c, J + 1, lane, vec_m);
        } else if (sub == 2) {
          // the transition column of the next tile's phase A towards L2: row 32 J + lane of the links plane,
          // transitions 31 - lane .. 62 - lane
          if (g.tile_alive(c, J + 1) && kB * J + lane < O) {
            const float *pf = E + (int64_t)(kB * J + lane) * Tl + (kB - 1 - lane);
#pragma unroll
            for (int x = 0; x < 40; x += 8)
              if (kB - 1 - lane + min(x, 31) < Tl) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + min(x, 31)));
          }
        }
      }
      {
        // ======================= every warp: far maxima of the NEXT step's tiles ===================================
        // Units (tile, source block, half) are pulled from a per-step counter in shared memory.
        const int sn = sg + 1;
        if (sn < nst) {
          int nun[kChain], total = 0;
#pragma unroll
          for (int x = 0; x < kChain; x++) {
            const int lcx = chunk_of(x, rank), cx = kPassChunks * p + lcx, Jx = sn - lcx;
            nun[x] = (lcx < cpass && g.tile_alive(cx, Jx)) ? 2 * max(0, Jx - 2 - g.ilo(cx, Jx) + 1) : 0;
            total += nun[x];
          }
          int *qnext = s_queue + (sn & 1);
          auto grab = [&]() {
            int u = 0;
            if (lane == 0) u = atomicAdd(qnext, 1);
            return __shfl_sync(0xffffffffu, u, 0);
          };
          // unit -> (tile slot, first source vertex of the half block)
          auto decode = [&](int u, int &tx, int &i0) {
            tx = 0;
            int r = u;
#pragma unroll
            for (int x = 0; x < kChain - 1; x++)
              if (tx == x && r >= nun[x]) { r -= nun[x]; tx = x + 1; }
            const int lcx = chunk_of(tx, rank), cx = kPassChunks * p + lcx, Jx = sn - lcx;
            i0 = kB * g.ilo(cx, Jx) + kHalf * r;
          };
          auto stage = [&](int u, float *slab) {
            int tx, i0;
            decode(u, tx, i0);
            const int tp0 = (kPassChunks * p + chunk_of(tx, rank)) * kB;       // previous-row index of the tile's first row
            if (vec_ok) {
              const int q = lane & 3;
#pragma unroll
              for (int rr = lane >> 2; rr < kB; rr += 8) {
                float *dst = slab + rr * kHalf + 4 * q;
                if (tp0 + rr < g.nsteps) cp_async16_cg(dst, lat + (int64_t)(tp0 + rr) * L + i0 + 4 * q);
                else *reinterpret_cast<float4 *>(dst) = make_float4(ninf, ninf, ninf, ninf);
              }
            } else {
              const int q = lane & 15;
              const float *lp = lat + (int64_t)tp0 * L + i0 + q;              // i0 + q < 32 J <= O - 1 < L
              for (int rr = lane >> 4; rr < kB; rr += 2) slab[rr * kHalf + q] = (tp0 + rr < g.nsteps) ? __ldcg(lp + (int64_t)rr * L) : ninf;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
          };
          int cur = (total > 0) ? grab() : total;
          if (cur < total) {
            float acc[kB];
#pragma unroll
            for (int k = 0; k < kB; k++) acc[k] = ninf;
            int buf = 0;
            stage(cur, vs0);
            while (cur < total) {
              const int nxt = grab();
              const bool more = nxt < total;
              if (more) stage(nxt, vs0 + (buf ^ 1) * kB * kHalf);
              int tx, i0;
              decode(cur, tx, i0);
              const int j = kB * (sn - chunk_of(tx, rank)) + lane;
              float ecol[kHalf];
              {
                const int k0 = j - i0 - 1;                    // transition index from the half block's first source vertex; >= 32
                const float *epn = E + (int64_t)i0 * Tl + k0;
                const bool jok = j < O;
#pragma unroll
                for (int ii = 0; ii < kHalf; ii++) ecol[ii] = (jok && k0 - ii < Tl) ? __ldg(epn + (int64_t)ii * (Tl - 1)) : ninf;
              }
              if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
              else asm volatile("cp.async.wait_group 0;" ::: "memory");
              __syncwarp();
              const float4 *vsw = reinterpret_cast<const float4 *>(vs0 + buf * kB * kHalf);
              // the four broadcast loads of a row are issued one row ahead of the arithmetic that consumes them
              float4 a0 = vsw[0], a1 = vsw[1], a2 = vsw[2], a3 = vsw[3];
#pragma unroll
              for (int rr = 0; rr < kB; rr++) {
                float4 n0 = a0, n1 = a1, n2 = a2, n3 = a3;
                if (rr + 1 < kB) { n0 = vsw[4 * (rr + 1)]; n1 = vsw[4 * (rr + 1) + 1]; n2 = vsw[4 * (rr + 1) + 2]; n3 = vsw[4 * (rr + 1) + 3]; }
                float b0 = acc[rr], b1;
                b0 = max3(b0, a0.x + ecol[0], a0.y + ecol[1]);
                b1 = fmaxf(a0.z + ecol[2], a0.w + ecol[3]);
                b0 = max3(b0, a1.x + ecol[4], a1.y + ecol[5]);
                b1 = max3(b1, a1.z + ecol[6], a1.w + ecol[7]);
                b0 = max3(b0, a2.x + ecol[8], a2.y + ecol[9]);
                b1 = max3(b1, a2.z + ecol[10], a2.w + ecol[11]);
                b0 = max3(b0, a3.x + ecol[12], a3.y + ecol[13]);
                b1 = max3(b1, a3.z + ecol[14], a3.w + ecol[15]);
                acc[rr] = fmaxf(b0, b1);
                a0 = n0; a1 = n1; a2 = n2; a3 = n3;
              }
              __syncwarp();
              buf ^= 1;
              // my next unit belongs to another tile (or there is none): the maxima of this tile meet the other warps'
              int txn = -1, i0n = 0;
              if (more) decode(nxt, txn, i0n);
              if (txn != tx) {
                int *xk = s_xk + ((sn & 1) * kChain + tx) * kB * kPitch + lane;
#pragma unroll
                for (int k = 0; k < kB; k++) {
                  if (acc[k] > ninf) atomicMax(xk + k * kPitch, f2key(acc[k]));
                  acc[k] = ninf;
                }
              }
              cur = nxt;
            }
          }
        }
        if (threadIdx.x == 0) s_queue[sg & 1] = 0;     // the counter of the step after next (nobody touches it in this step)
      }
#ifdef DAGB200_V3_TIMING
      const long long tbusy = clock64() - tstep0;
#endif
      cluster_sync_all();      // the step is complete in both CTAs and visible to both
#ifdef DAGB200_V3_TIMING
      if (tlog) {
        if (warp == 0) g_v3t[sg][0] = tbusy;
        if (warp == 4) { g_v3t[sg][2] = clock64() - tstep0; }
        if (warp == 15) g_v3t[sg][4] = tbusy;
        if (warp == 2) g_v3t[sg][5] = tbusy;
      }
#endif
    }
  }

  // ---- backtrace (dag_best_alignment.cu:178-184) with the back-pointers recomputed on the way: for the cell (i, pos) of
  // the path all threads of CTA 0 score its candidates delta = 1 .. min(pos, Tl) that hold a lattice cell (source
  // vertex >= i-1; the others are -inf) -- ONE fp32 add each, as in the forward sweep -- and reduce them with the
  // reference's order: value, then class priority 0,2,1,3 of (delta-1) mod 4, then smaller delta.  The reduction is two
  // warp-wide integer reductions (REDUX) on (value key) and (priority << 16 | delta) per level and ONE block barrier per
  // step: every warp repeats the second level, so no result has to be broadcast.
  if (rank == 0) {
    int *s_k1 = reinterpret_cast<int *>(v3_smem);            // [2 (step parity)][warps] best value key of a warp
    int *s_k2 = s_k1 + 2 * kWarps;                           // [2][warps] its (priority << 16 | delta)
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
        // my candidates: delta = dmax - tid, dmax - tid - 512, ... (descending); all loads first
        float bv = ninf; int bd = 0;
        for (int d0 = dmax - (int)threadIdx.x; d0 >= 1; d0 -= 2 * kThreads) {
          const int d1 = d0 - kThreads;
          const float *e0 = E + (int64_t)(pos - d0) * Tl + d0 - 1;
          const float *e1 = E + (int64_t)(pos - max(d1, 1)) * Tl + max(d1, 1) - 1;
          const float v0 = __ldcg(prev + pos - d0), t0 = __ldg(e0);
          const float v1 = d1 >= 1 ? __ldcg(prev + pos - d1) : ninf, t1 = d1 >= 1 ? __ldg(e1) : 0.f;
          // the next row asks the same source vertices for a transition a few columns to the left (the path moves left
          // by the delta chosen here) and for their values one row up: pull both towards L2 now
#ifndef DAGB200_BT_PREFETCH
#define DAGB200_BT_PREFETCH 0
#endif
          if (DAGB200_BT_PREFETCH >= 1 && i >= 2) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(e0 - min(d0 - 1, 8)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(prev - L + pos - d0));
            if (d1 >= 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(e1 - min(d1 - 1, 8)));
            if (DAGB200_BT_PREFETCH >= 2) {
              asm volatile("prefetch.global.L2 [%0];" ::"l"(e0 - min(d0 - 1, 16)));
              if (d1 >= 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(e1 - min(d1 - 1, 16)));
              if (i >= 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(prev - 2 * L + pos - d0));
            }
          }
          const float x0 = v0 + t0, x1 = v1 + t1;
          if (better(x0, d0, bv, bd)) { bv = x0; bd = d0; }
          if (d1 >= 1 && better(x1, d1, bv, bd)) { bv = x1; bd = d1; }
        }
        // level 1: the warp's best value, then the preferred delta among the lanes that hold it
        const int par = (i & 1) * kWarps;
        {
          const int k1 = f2key(bv + 0.0f);                 // (-0 and +0 are one value)
          const int w1 = __reduce_max_sync(0xffffffffu, k1);
          const int k2 = (k1 == w1 && bv > ninf) ? ((rank4(bd) << 16) | bd) : 0x7fffffff;
          const int w2 = __reduce_min_sync(0xffffffffu, k2);
          if (lane == 0) { s_k1[par + warp] = w1; s_k2[par + warp] = w2; }
        }
        __syncthreads();
        // level 2, in every warp
        int dsel;
        {
          const int k1 = lane < kWarps ? s_k1[par + lane] : f2key(ninf);
          const int k2r = lane < kWarps ? s_k2[par + lane] : 0x7fffffff;
          const int w1 = __reduce_max_sync(0xffffffffu, k1);
          const int k2 = (k1 == w1) ? k2r : 0x7fffffff;
          const int w2 = __reduce_min_sync(0xffffffffu, k2);
          dsel = (w2 == 0x7fffffff) ? 0 : (w2 & 0xffff);
        }
        if (dsel == 0) { code = DAGB200_ST_NO_PATH; break; }
        pos -= dsel;
      }
      if (code == DAGB200_ST_OK && threadIdx.x == 0) prow[pos] = 0;
    }
    if (status && threadIdx.x == 0) status[b] = code;
  }
}

}  // namespace v3

size_t vit3_smem_bytes() {
  using namespace v3;
  return sizeof(float) * ((size_t)5 * kChain * kB * kEdPitch + (size_t)2 * kChain * kB * kPitch + kChain * 2 * kB + 4 +
                          (size_t)kWarps * 2 * kB * kHalf) + 16;
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
#ifdef DAGB200_V3_TIMING
  {
    static int calls = 0;
    if (++calls == 3) {
      cudaStreamSynchronize(st);
      long long h[64][8];
      cudaMemcpyFromSymbol(h, g_v3t, sizeof(h));
      printf("[v3 timing, CTA 0] step: warp0 busy, warp2 busy, warp15 busy, step length | chain warp 0, since step start: team barrier passed, -, sweep done, stored\n");
      printf("  forward (prologue + steps) %lld cycles, backtrace %lld cycles (candidate loads of thread 0: %lld over 3 launches)\n", h[62][0], h[62][1], h[62][2]);
      for (int i = 0; i < 40; i++) printf("  %2d  %7lld %7lld %7lld %7lld | %7lld %7lld %7lld %7lld\n", i, h[i][0], h[i][5], h[i][4], h[i][2], h[i][3], h[i][6], h[i][7], h[i][1]);
    }
  }
#endif
  prof_mark(7, st);
  return 0;
}

}  // namespace dagb200
