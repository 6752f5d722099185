// dag_dp3.cu -- column-major blocked forward (alpha) / backward (beta) recurrences of the DAG loss for sm_100a.
//
// Replaces calculate_alpha_kernel / calculate_beta_kernel (reference dag_loss.cu:40-140, 178-274) on the fp32
// path; successor of the anti-diagonal kernel in dag_dp2.cu (kept as the fallback for very long graphs).
//
//   vertices in blocks of 32 (index q in sweep order), target rows in chunks of 32; a PASS = 8 consecutive chunks
//   (256 rows).  Inside a pass the blocks are swept in order q = 0, 1, ...; for every block
//     phase 1   8 "chain" warps (one per chunk) run the 32x32 diagonal block of their chunk, lanes = rows, columns
//               serially, in fp64: cell (r, c) = (far[r][c] + handed[r-1][c]) * emission, then the cell pushes its
//               mass into the predecessor sums of the later columns of its own row (DFMA).  One fp64 frame
//               (power of two) per tile, so nothing on the dependency path but SHFL + DADD + DMUL + DFMA.  Chunk
//               c+1 follows chunk c one column behind (its row 0 needs the last row of chunk c), handed through
//               shared memory with release/acquire counters.
//               MEANWHILE the 8 "GEMM" warps (32 rows each) accumulate the far predecessor sums of block q+1 from
//               the blocks < q on the tensor cores: mma.sync.m16n8k16 bf16, both operands split hi/lo (3 MMAs per
//               product), A operand = cached previous-row masses (bf16 hi/lo fragments, one integer frame per
//               (row, block)), B operand = transition tiles streamed ONCE per pass through a TMA ring.
//     phase 2   the GEMM warps add the contribution of block q itself (A operand built straight from the chain
//               warps' shared-memory rows) and hand the far sums of block q+1 to the chain warps, while the chain
//               warps publish block q as A-operand fragments and write the lattice rows.
//
// Flush-to-zero contract (DESIGN.md "Numerics"): a far predecessor contributes exactly 0 when its mass is more
// than 2^-126 below the largest far predecessor of the 32-vertex destination block; a transition when it is > 87
// nats below the best transition of its source vertex; inside a tile everything is fp64 relative to one frame
// (contributions more than ~2^-1000 below it vanish).  In-block predecessors of EARLIER 8-column groups enter
// with 21 significant bits (2^-21 relative), below the 2^-16 of the split bf16 products.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "dag_tiles.cuh"

namespace dagb200 {
namespace dp3 {

constexpr int kThreads = 512;
constexpr int kCW = 8;                          // chain warps = chunks per pass (GEMM warps: the other 8)
constexpr int kStages = 6;                      // transition-tile ring
constexpr int kPitch = 33;                      // padded row pitch of 32x32 fp32 tiles in shared memory
constexpr int kTileF = 32 * kPitch;             // floats per staged 32x32 tile
constexpr int kNegBig = -(1 << 20);             // "empty" integer frame
constexpr int kEv = 34;                         // progress events per tile: anchor, then one per column (+1 spare)
constexpr float kLn2Hi = 0.693359375f;          // 355/512: k * kLn2Hi is exact for |k| < 2^15
constexpr float kLn2Lo = -2.12194440e-4f;       // ln2 - kLn2Hi

__device__ long long g_dbg[8];

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^d as float for d <= 127; exactly 0 below the normal range
__device__ __forceinline__ float pow2i(int d) { return __int_as_float((max(d, -127) + 127) << 23); }
// 2^d as double, exactly 0 for d <= -1023, clamped above
__device__ __forceinline__ double pow2d(int d) { return __hiloint2double((min(max(d, -1023), 1023) + 1023) << 20, 0); }

__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t *>(&h);
  lo = *reinterpret_cast<uint32_t *>(&l);
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ int ld_acquire_s32(const int *p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int *p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// one element of a fragment tile of the A operand (see dag_tiles.cuh / dag_dp2.cu)
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

struct Smem {
  uint4 *ring;      // [kStages][256]      transition tiles in flight (TMA destination)
  double *ut;       // [2][32*32]          fp64 push table of the current / next block
  float *xbuf;      // [2][kCW][kTileF]    far sums in -> masses out (normalised fp32 after phase 1)
  float *io;        // [2][kCW][kTileF]    emissions in -> lattice values out
  double *hand;     // [kCW][32]           predecessor sums of a chunk's last row, handed to the next chunk
  int *fbuf;        // [kCW][32]           far frames of the rows of the current tile
  int *prog;        // [kCW]               progress counters of the chain warps (events)
  int *tanchor;     // [kCW]               fp64 frame of each chain warp's current tile
  float *rmax;      // [NB*32]             per-source-vertex transition maximum
  int *rmtab;       // [257][NB]           per (row of the pass, block) integer upper bound of log2(outgoing mass)
  uint64_t *full;   // [kStages]
  uint64_t *empty;  // [kStages]
  uint64_t *ubar;   // [2]
};

struct Geo {
  int O, Tn, M, L, NB, NBv, nsteps, NCv, NP, band;
  bool dbg;
};

// a tile without a single lattice cell (below the diagonal j >= t): nothing flows through it
template <bool BETA>
__device__ __forceinline__ bool tile_geo_dead(const Geo &g, int c, int J) {
  const int smax = min(c * 32 + 31, g.nsteps - 1);
  const int tmin = BETA ? g.Tn - 2 - smax : 1 + c * 32;
  return min(kBlk * J + 31, g.O - 1) < tmin;
}

// ---------------------------------------------------------------------------------------------------------
// phase 1 of a chain warp: the 32x32 diagonal block of (chunk c, block q)
template <bool BETA>
__device__ __forceinline__ void chain_phase1(const Geo &g, const Smem &sm, const float *__restrict__ match,
                                             double *__restrict__ passd, int *__restrict__ passf, int p, int q,
                                             int cw, int lane, int ustep) {
  const int c = p * kCW + cw;
  const int J = BETA ? g.NBv - 1 - q : q;
  const int jbase = kBlk * J;
  const int s = c * 32 + lane;
  const bool rowvalid = s < g.nsteps;
  const int t = BETA ? g.Tn - 2 - s : 1 + s;
  float *xb = sm.xbuf + ((size_t)(q & 1) * kCW + cw) * kTileF;
  float *iob = sm.io + ((size_t)(q & 1) * kCW + cw) * kTileF;
  float *mrow = xb + lane * kPitch;
  float *iow = iob + lane * kPitch;
  const float ninf = neg_inf_f();
  const bool feeds_next = (c + 1 < g.NCv);       // somebody consumes my last row
  const bool to_pass = feeds_next && cw == kCW - 1;

  // (1) start fetching the emissions of the next block into the other buffer
  if (q + 1 < g.NBv) {
    const int Jn = BETA ? g.NBv - 2 - q : q + 1;
    float *ion = sm.io + ((size_t)((q + 1) & 1) * kCW + cw) * kTileF;
    const int j = kBlk * Jn + lane;
    for (int rr = 0; rr < 32; rr++) {
      const int sr = c * 32 + rr;
      const int tr = BETA ? g.Tn - 2 - sr : 1 + sr;
      if (sr < g.nsteps && j < g.L) cp_async_f32(ion + rr * kPitch + lane, match + (int64_t)tr * g.L + j);
      else ion[rr * kPitch + lane] = ninf;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // (2) frames: far frames of my rows, the frame handed from the chunk above, the fp64 frame of this tile
  const int FI = sm.fbuf[cw * 32 + lane];
  const int maxFI = __reduce_max_sync(0xffffffffu, FI);
  int Fh;
  double hv = 0.0;                               // cw == 0: lane cj holds the handed sum of column cj
  int known = 0;                                 // cw > 0: last progress value seen of the warp above
  const int evbase = q * kEv;
  if (cw == 0) {
    Fh = passf[(p & 1) * g.NB + q];
    hv = passd[((size_t)(p & 1) * g.NB + q) * 32 + lane];
  } else {
    do { known = ld_acquire_s32(sm.prog + cw - 1); } while (known < evbase + 1);
    Fh = sm.tanchor[cw - 1];
  }
  const bool dead = tile_geo_dead<BETA>(g, c, J) || max(maxFI, Fh) <= kNegBig;
  const int Ft = dead ? kNegBig : max(maxFI, Fh - 600);
  if (feeds_next && !to_pass) {
    if (lane == 0) sm.tanchor[cw] = Ft;
    if (dead) sm.hand[cw * 32 + lane] = 0.0;
    __syncwarp();
    if (lane == 0) st_release_s32(sm.prog + cw, dead ? evbase + kEv : evbase + 1);
  }
  if (to_pass) {
    if (lane == 0) passf[((p + 1) & 1) * g.NB + q] = Ft;
    if (dead) passd[((size_t)((p + 1) & 1) * g.NB + q) * 32 + lane] = 0.0;
  }
  // emissions of THIS block have landed (the group committed above may still be in flight)
  asm volatile("cp.async.wait_group 1;" ::: "memory");
  __syncwarp();
  if (dead) {   // nothing reaches this tile: -inf lattice values, zero masses
#pragma unroll 4
    for (int k = 0; k < 32; k++) { iow[k] = ninf; mrow[k] = 0.f; }
    sm.rmtab[(cw * 32 + lane + 1) * g.NB + q] = kNegBig;
    __syncwarp();
    return;
  }
  // (3) the fp64 push table of this block
  mbar_wait(sm.ubar + (ustep & 1), (ustep >> 1) & 1);
  const double *ut = sm.ut + (size_t)(ustep & 1) * 1024;
  const double hs = pow2d(Fh - Ft);              // handed sums are in the frame of the tile above
  const double xs = pow2d(FI - Ft);              // far sums of my row are in the frame FI
  const float *rmax_blk = sm.rmax + jbase;
  int maxhi = 0;

  // (4) column sweep: four groups of 8 columns (runtime loop keeps the code small)
  const bool swdbg = g.dbg && blockIdx.x == 0 && lane == 0 && (cw == 0 || cw == 7);
  const long long sw0 = swdbg ? clock64() : 0;
#pragma unroll 1
  for (int G = 0; G < 4; G++) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.0;
    // predecessor sums from the completed groups of my own row (masses re-read with 21 significant bits)
    for (int ci = 0; ci < 8 * G; ci++) {
      const int w = __float_as_int(mrow[BETA ? 31 - ci : ci]);
      const double md = __hiloint2double(w, 0);
      const double2 *ur = reinterpret_cast<const double2 *>(ut + ci * 32 + 8 * G);
      const double2 u0 = ur[0], u1 = ur[1], u2 = ur[2], u3 = ur[3];
      a[0] = fma(md, u0.x, a[0]); a[1] = fma(md, u0.y, a[1]); a[2] = fma(md, u1.x, a[2]); a[3] = fma(md, u1.y, a[3]);
      a[4] = fma(md, u2.x, a[4]); a[5] = fma(md, u2.y, a[5]); a[6] = fma(md, u3.x, a[6]); a[7] = fma(md, u3.y, a[7]);
    }
#pragma unroll
    for (int K = 0; K < 8; K++) {
      const int cj = 8 * G + K;
      const int jj = BETA ? 31 - cj : cj;
      const int j = jbase + jj;
      const bool valid = rowvalid && j >= t && j < g.O;
      // what the row above hands to this column (lane 0: from the chunk above / the previous pass)
      double z;
      if (cw == 0) {
        z = __shfl_sync(0xffffffffu, hv, cj) * hs;
      } else {
        while (known < evbase + 2 + cj) known = ld_acquire_s32(sm.prog + cw - 1);
        z = sm.hand[(cw - 1) * 32 + cj] * hs;
      }
      // hand my own last row to the chunk below
      if (feeds_next && lane == 31) {
        if (to_pass) {
          passd[((size_t)((p + 1) & 1) * g.NB + q) * 32 + cj] = a[K];
        } else {
          sm.hand[cw * 32 + cj] = a[K];
          st_release_s32(sm.prog + cw, evbase + 2 + cj);
        }
      }
      double rm = __shfl_up_sync(0xffffffffu, a[K], 1);
      if (lane == 0) rm = z;
      const float X = mrow[jj];
      const double tot = fma((double)X, xs, rm);
      // emission weight exp(match + rmax) as a double (0 for cells outside the lattice)
      const float em = iow[jj];
      const float rmx = rmax_blk[jj];
      const float w2 = (em + rmx) * kLog2e;
      double ewd = 0.0;
      if (valid && w2 > -1.0e30f) {
        const float wf = floorf(w2);
        ewd = (double)exp2f(w2 - wf) * pow2d((int)wf);
      }
      const double m = tot * ewd;
      // lattice value (off the dependency path): log of tot through its exponent and leading mantissa bits
      float out = ninf;
      const int thi = __double2hiint(tot);
      if (valid && thi >= 0x00100000) {
        const int tlo = __double2loint(tot);
        const int e2 = (thi >> 20) - 1023 + Ft;
        const float mant = __int_as_float(0x3f800000 | ((thi & 0xfffff) << 3) | ((unsigned)tlo >> 29));
        const float fl = (float)e2;
        out = (em + fmaf(__log2f(mant), 0.6931471805599453f, fl * kLn2Lo)) + fl * kLn2Hi;
        if (BETA) out += rmx;
      }
      iow[jj] = out;
      const int mhi = __double2hiint(m);
      mrow[jj] = __int_as_float(mhi);             // mass, 21 significant bits, frame Ft
      maxhi = max(maxhi, mhi);
      // push into the later columns of my row
      if (K < 7) {
        const double *ur = ut + cj * 32 + 8 * G;
#pragma unroll
        for (int k2 = K + 1; k2 < 8; k2++) a[k2] = fma(m, ur[k2], a[k2]);
      }
    }
  }
  if (swdbg) atomicAdd((unsigned long long *)&g_dbg[4 + (cw == 7 ? 1 : 0) + (BETA ? 2 : 0)], (unsigned long long)(clock64() - sw0));
  // (5) row frame and normalised masses (value / 2^maxe < 1) for the A-operand fragments
  const int emf = maxhi >> 20;                   // biased exponent of the largest mass of my row (0: none)
  const int maxe = (emf > 0) ? Ft + emf - 1022 : kNegBig;
  sm.rmtab[(cw * 32 + lane + 1) * g.NB + q] = maxe;
#pragma unroll 8
  for (int k = 0; k < 32; k++) {
    const int w = __float_as_int(mrow[k]);
    const int we = w >> 20;
    const int fe = we - emf + 126;
    const int bits = (we > 0 && fe > 0) ? ((fe << 23) | ((w & 0xfffff) << 3)) : 0;
    mrow[k] = __int_as_float(bits);
  }
  __syncwarp();
}

// phase 2 of a chain warp: publish the block as A-operand fragments, write the lattice rows
template <bool BETA>
__device__ __forceinline__ void chain_phase2(const Geo &g, const Smem &sm, float *__restrict__ lat,
                                             uint4 *__restrict__ afrag, int p, int q, int cw, int lane) {
  const int c = p * kCW + cw;
  const int J = BETA ? g.NBv - 1 - q : q;
  const float *vt = sm.xbuf + ((size_t)(q & 1) * kCW + cw) * kTileF;   // slots indexed by vertex offset = K index
  const float *iot = sm.io + ((size_t)(q & 1) * kCW + cw) * kTileF;
  constexpr int VP = kPitch;
  uint4 *ft = afrag + ((size_t)c * g.NBv + q) * 256;
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
        // row 0 of this tile belongs to the chunk above (or the seed): keep a0 / a2
        uint32_t *ph = reinterpret_cast<uint32_t *>(ft + (slice * 4 + ks * 2 + 0) * 32 + lane);
        uint32_t *pl = reinterpret_cast<uint32_t *>(ft + (slice * 4 + ks * 2 + 1) * 32 + lane);
        ph[1] = hi.y; ph[3] = hi.w; pl[1] = lo.y; pl[3] = lo.w;
      } else {
        ft[(slice * 4 + ks * 2 + 0) * 32 + lane] = hi;
        ft[(slice * 4 + ks * 2 + 1) * 32 + lane] = lo;
      }
    }
  // my chunk's last row is row 0 of the next chunk's tile
  if (c + 1 < g.NCv && lane < 4) {
    uint4 *fn = afrag + ((size_t)(c + 1) * g.NBv + q) * 256;
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
  // lattice values: coalesced row writes (lane = column)
  const int j = kBlk * J + lane;
  const int rl = min(32, g.nsteps - c * 32) - 1;
  if (j < g.L) {
    for (int rr = 0; rr <= rl; rr++) {
      const int sr = c * 32 + rr;
      const int tr = BETA ? g.Tn - 2 - sr : 1 + sr;
      lat[(int64_t)tr * g.L + j] = iot[rr * kPitch + lane];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// GEMM warps.  Accumulators: 32 rows (two 16-row slices) x 32 destination columns.
struct TileSeq {       // static order of the transition tiles of a pass: destination J = 1.., sources qlo(J)..J-1
  int J, qs;
  __device__ __forceinline__ void reset(int band) { J = 1; qs = max(0, 1 - band); }
  __device__ __forceinline__ void advance(int NBv, int band) {
    qs++;
    if (qs > J - 1) {
      J++;
      if (J >= NBv) J = 1;
      qs = max(0, J - band);
    }
  }
};

struct Ring {
  int n;          // tiles consumed so far (all GEMM warps)
  int issued;     // tiles issued so far (producer lane only)
  int total;
  TileSeq next;   // next tile to issue (producer)
  bool dbg;
  long long t_prod, t_full, t_comp, t_p2;
};

template <bool BETA>
__device__ __forceinline__ void ring_produce(Ring &r, const Geo &g, const Smem &sm, const uint4 *__restrict__ tiles,
                                             const TileLayout &lay, bool producer) {
  // keep up to kStages tiles in flight: before consuming tile n, issue everything up to tile n + kStages - 1
  if (!producer) return;
  while (r.issued < r.total && r.issued < r.n + kStages) {
    const int st = r.issued % kStages;
    if (r.issued >= kStages) mbar_wait(sm.empty + st, ((r.issued / kStages) - 1) & 1);
    const int Jd = BETA ? g.NBv - 1 - r.next.J : r.next.J;      // vertex-block index of the destination (sweep J)
    const int Js = BETA ? g.NBv - 1 - r.next.qs : r.next.qs;    // ... of the source
    const uint4 *src = tiles + (BETA ? lay.idxB(Jd, Js) : lay.idxA(Js, Jd)) * (kTileBytes / 16);
    mbar_expect_tx(sm.full + st, kTileBytes);
    bulk_g2s(sm.ring + (size_t)st * 256, src, kTileBytes, sm.full + st);
    r.next.advance(g.NBv, g.band);
    r.issued++;
  }
}

// 12 MMAs per 16-row slice for one k16 step: (ahi, alo) x (h0, h1 | l0, l1)
__device__ __forceinline__ void mma_half(float (&acc)[2][4][4], const uint32_t (&ahi)[2][4], const uint32_t (&alo)[2][4],
                                         const uint4 &h0, const uint4 &h1, const uint4 &l0, const uint4 &l1) {
#pragma unroll
  for (int sl = 0; sl < 2; sl++) {
    mma_bf16_16816(acc[sl][0], ahi[sl], h0.x, h0.y); mma_bf16_16816(acc[sl][1], ahi[sl], h0.z, h0.w);
    mma_bf16_16816(acc[sl][2], ahi[sl], h1.x, h1.y); mma_bf16_16816(acc[sl][3], ahi[sl], h1.z, h1.w);
  }
#pragma unroll
  for (int sl = 0; sl < 2; sl++) {
    mma_bf16_16816(acc[sl][0], alo[sl], h0.x, h0.y); mma_bf16_16816(acc[sl][1], alo[sl], h0.z, h0.w);
    mma_bf16_16816(acc[sl][2], alo[sl], h1.x, h1.y); mma_bf16_16816(acc[sl][3], alo[sl], h1.z, h1.w);
  }
#pragma unroll
  for (int sl = 0; sl < 2; sl++) {
    mma_bf16_16816(acc[sl][0], ahi[sl], l0.x, l0.y); mma_bf16_16816(acc[sl][1], ahi[sl], l0.z, l0.w);
    mma_bf16_16816(acc[sl][2], ahi[sl], l1.x, l1.y); mma_bf16_16816(acc[sl][3], ahi[sl], l1.z, l1.w);
  }
}

__device__ __forceinline__ uint32_t hmul2_u32(uint32_t v, const __nv_bfloat162 &sc) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162 *>(&v), sc);
  return *reinterpret_cast<uint32_t *>(&r);
}

// rescale the fragments {sl0 hi, sl0 lo, sl1 hi, sl1 lo} of one k16 step by the per-row powers of two and multiply
__device__ __forceinline__ void consume_half(float (&acc)[2][4][4], const uint4 (&A)[4], const __nv_bfloat162 (&sc)[4],
                                             const uint4 *tsm, int ks, int lane) {
  uint32_t ahi[2][4], alo[2][4];
#pragma unroll
  for (int sl = 0; sl < 2; sl++) {
    const uint4 fh = A[2 * sl], fl = A[2 * sl + 1];
    ahi[sl][0] = hmul2_u32(fh.x, sc[2 * sl]); ahi[sl][1] = hmul2_u32(fh.y, sc[2 * sl + 1]);
    ahi[sl][2] = hmul2_u32(fh.z, sc[2 * sl]); ahi[sl][3] = hmul2_u32(fh.w, sc[2 * sl + 1]);
    alo[sl][0] = hmul2_u32(fl.x, sc[2 * sl]); alo[sl][1] = hmul2_u32(fl.y, sc[2 * sl + 1]);
    alo[sl][2] = hmul2_u32(fl.z, sc[2 * sl]); alo[sl][3] = hmul2_u32(fl.w, sc[2 * sl + 1]);
  }
  const uint4 h0 = tsm[(2 * ks) * 32 + lane], h1 = tsm[(2 * ks + 1) * 32 + lane];
  const uint4 l0 = tsm[(4 + 2 * ks) * 32 + lane], l1 = tsm[(4 + 2 * ks + 1) * 32 + lane];
  mma_half(acc, ahi, alo, h0, h1, l0, l1);
}

struct GemmState {
  float acc[2][4][4];
  int Fp[4];        // frame of the accumulators, per row (index 2*slice + half)
};

// far predecessors of destination block J (sweep index) from the source blocks qlo .. J-2  (phase 1)
template <bool BETA>
__device__ __forceinline__ void gemm_phase1(GemmState &gs, Ring &ring, const Geo &g, const Smem &sm,
                                            const uint4 *__restrict__ tiles, const uint4 *__restrict__ afrag,
                                            const TileLayout &lay, int p, int J, int gw, int lane, bool producer) {
  const int c = p * kCW + gw;
  const bool active = c < g.NCv && !tile_geo_dead<BETA>(g, c, BETA ? g.NBv - 1 - J : J);
  const int gid = lane >> 2;
  const int qlo = max(0, J - g.band);
  const int nsrc = (J - 1) - qlo;                 // sources handled here
#pragma unroll
  for (int sl = 0; sl < 2; sl++)
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int e = 0; e < 4; e++) gs.acc[sl][a][e] = 0.f;
  int lr[4];
  bool rv[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = 16 * (i >> 1) + gid + 8 * (i & 1);
    lr[i] = (gw * 32 + r) * g.NB;
    rv[i] = active && (c * 32 + r) < g.nsteps;
  }
  if (qlo > 0) {   // banded transitions: the window of sources moves, rescan its frames
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int f = kNegBig;
      if (rv[i]) for (int qs = qlo; qs <= J - 2; qs++) f = max(f, sm.rmtab[lr[i] + qs]);
      gs.Fp[i] = f;
    }
  }
  if (nsrc <= 0) return;
  const uint4 *fbase = afrag + (size_t)c * g.NBv * 256 + lane;
  uint4 A0[4], A1[4];
  auto load_half = [&](uint4 (&A)[4], int qs, int ks) {
    const uint4 *ft = fbase + (size_t)qs * 256;
    A[0] = ft[(0 * 4 + ks * 2 + 0) * 32]; A[1] = ft[(0 * 4 + ks * 2 + 1) * 32];
    A[2] = ft[(1 * 4 + ks * 2 + 0) * 32]; A[3] = ft[(1 * 4 + ks * 2 + 1) * 32];
  };
  if (active) { load_half(A0, qlo, 0); load_half(A1, qlo, 1); }
  for (int qs = qlo; qs <= J - 2; qs++) {
    long long c0 = ring.dbg ? clock64() : 0;
    ring_produce<BETA>(ring, g, sm, tiles, lay, producer);
    __syncwarp();
    long long c1 = ring.dbg ? clock64() : 0;
    const int st = ring.n % kStages;
    mbar_wait(sm.full + st, (ring.n / kStages) & 1);
    long long c2 = ring.dbg ? clock64() : 0;
    if (ring.dbg) { ring.t_prod += c1 - c0; ring.t_full += c2 - c1; }
    if (active) {
      __nv_bfloat162 sc[4];
      bool any = false;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int Fs = rv[i] ? sm.rmtab[lr[i] + qs] : kNegBig;
        const float s = (Fs > kNegBig) ? pow2i(Fs - gs.Fp[i]) : 0.f;
        any = any || (s != 0.f);
        sc[i] = __floats2bfloat162_rn(s, s);
      }
      const uint4 *tsm = sm.ring + (size_t)st * 256;
      const bool work = __any_sync(0xffffffffu, any);
      if (work) consume_half(gs.acc, A0, sc, tsm, 0, lane);
      if (qs + 1 <= J - 2) load_half(A0, qs + 1, 0);
      if (work) consume_half(gs.acc, A1, sc, tsm, 1, lane);
      if (qs + 1 <= J - 2) load_half(A1, qs + 1, 1);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.empty + st);
    ring.n++;
    if (ring.dbg) ring.t_comp += clock64() - c2;
  }
}

// phase 2: add the source block J-1 (just finished by the chain warps), hand the far sums of block J over
template <bool BETA>
__device__ __forceinline__ void gemm_phase2(GemmState &gs, Ring &ring, const Geo &g, const Smem &sm,
                                            const uint4 *__restrict__ tiles, const uint4 *__restrict__ afrag,
                                            const TileLayout &lay, int p, int J, int gw, int lane, bool producer) {
  const int c = p * kCW + gw;
  const bool active = c < g.NCv && !tile_geo_dead<BETA>(g, c, BETA ? g.NBv - 1 - J : J);
  const int gid = lane >> 2, tig = lane & 3;
  const int qsrc = J - 1;
  ring_produce<BETA>(ring, g, sm, tiles, lay, producer);
  __syncwarp();
  const int st = ring.n % kStages;
  mbar_wait(sm.full + st, (ring.n / kStages) & 1);
  if (active) {
    int F[4];
    __nv_bfloat162 sc[4];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = 16 * (i >> 1) + gid + 8 * (i & 1);
      const bool rv = (c * 32 + r) < g.nsteps;
      const int Fl = rv ? sm.rmtab[(gw * 32 + r) * g.NB + qsrc] : kNegBig;
      F[i] = max(gs.Fp[i], Fl);
      const float rs = pow2i(gs.Fp[i] - F[i]);     // rescale what has been accumulated so far (exact)
      const int sl = i >> 1, h = i & 1;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) { gs.acc[sl][nt][2 * h] *= rs; gs.acc[sl][nt][2 * h + 1] *= rs; }
      const float s = (Fl > kNegBig) ? pow2i(Fl - F[i]) : 0.f;
      any = any || (s != 0.f);
      sc[i] = __floats2bfloat162_rn(s, s);
    }
    if (__any_sync(0xffffffffu, any)) {
      // A operand straight from the chain warps' rows: consumer row rho <- producer row rho - 1
      const float *vt = sm.xbuf + ((size_t)(qsrc & 1) * kCW + gw) * kTileF;
      const float *vprev = sm.xbuf + ((size_t)(qsrc & 1) * kCW + (gw > 0 ? gw - 1 : 0)) * kTileF + 31 * kPitch;
      const uint4 *ft0 = afrag + ((size_t)c * g.NBv + qsrc) * 256 + lane;   // row 0 of the first chunk of a pass
      const uint4 *tsm = sm.ring + (size_t)st * 256;
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        uint4 A[4];
#pragma unroll
        for (int sl = 0; sl < 2; sl++) {
          const int rA = 16 * sl + gid - 1, rB = rA + 8;
          const int k0 = 16 * ks + 2 * tig;
          float e[8];
          if (rA >= 0) {
            e[0] = vt[rA * kPitch + k0]; e[1] = vt[rA * kPitch + k0 + 1];
            e[4] = vt[rA * kPitch + k0 + 8]; e[5] = vt[rA * kPitch + k0 + 9];
          } else if (gw > 0) {
            e[0] = vprev[k0]; e[1] = vprev[k0 + 1]; e[4] = vprev[k0 + 8]; e[5] = vprev[k0 + 9];
          } else {
            e[0] = e[1] = e[4] = e[5] = 0.f;       // replaced below by the stored fragments of the previous pass
          }
          e[2] = vt[rB * kPitch + k0]; e[3] = vt[rB * kPitch + k0 + 1];
          e[6] = vt[rB * kPitch + k0 + 8]; e[7] = vt[rB * kPitch + k0 + 9];
          uint4 hi, lo;
          split_bf16x2(e[0], e[1], hi.x, lo.x);
          split_bf16x2(e[2], e[3], hi.y, lo.y);
          split_bf16x2(e[4], e[5], hi.z, lo.z);
          split_bf16x2(e[6], e[7], hi.w, lo.w);
          if (sl == 0 && gw == 0 && gid == 0) {
            const uint32_t *ph = reinterpret_cast<const uint32_t *>(ft0 + (ks * 2 + 0) * 32);
            const uint32_t *pl = reinterpret_cast<const uint32_t *>(ft0 + (ks * 2 + 1) * 32);
            hi.x = ph[0]; hi.z = ph[2]; lo.x = pl[0]; lo.z = pl[2];
          }
          A[2 * sl] = hi; A[2 * sl + 1] = lo;
        }
        consume_half(gs.acc, A, sc, tsm, ks, lane);
      }
    }
    // far sums and frames of block J for the chain warps
    float *xo = sm.xbuf + ((size_t)(J & 1) * kCW + gw) * kTileF;
#pragma unroll
    for (int sl = 0; sl < 2; sl++) {
      const int r0 = 16 * sl + gid, r1 = r0 + 8;
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int n = 8 * nt + 2 * tig;
        xo[r0 * kPitch + n] = gs.acc[sl][nt][0]; xo[r0 * kPitch + n + 1] = gs.acc[sl][nt][1];
        xo[r1 * kPitch + n] = gs.acc[sl][nt][2]; xo[r1 * kPitch + n + 1] = gs.acc[sl][nt][3];
      }
    }
    if (tig == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) sm.fbuf[gw * 32 + 16 * (i >> 1) + gid + 8 * (i & 1)] = F[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++) gs.Fp[i] = F[i];
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(sm.empty + st);
  ring.n++;
}

// ---------------------------------------------------------------------------------------------------------
// One direction of one utterance.
template <bool BETA>
__device__ void colmajor_dir(const float *__restrict__ match, float *__restrict__ lat, unsigned char *__restrict__ ws,
                             const TileLayout &lay, const Smem &sm, int O, int Tn, int M, int L, int Tl, bool dbg) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float ninf = neg_inf_f();
  Geo g;
  g.O = O; g.Tn = Tn; g.M = M; g.L = L; g.NB = lay.NB;
  g.NBv = (O + kBlk - 1) / kBlk;
  g.nsteps = Tn - 1;
  g.NCv = (g.nsteps + 31) / 32;
  g.NP = (g.NCv + kCW - 1) / kCW;
  g.band = band_blocks(Tl);
  g.dbg = dbg;
  const float *g_rmax = reinterpret_cast<const float *>(ws + lay.off_rmax);
  const double *push = reinterpret_cast<const double *>(ws + (BETA ? lay.off_pushB : lay.off_pushA));
  const uint4 *tiles = reinterpret_cast<const uint4 *>(ws + (BETA ? lay.off_tilesB : lay.off_tilesA));
  uint4 *afrag = reinterpret_cast<uint4 *>(ws + (BETA ? lay.off_afragB : lay.off_afragA));
  double *passd = reinterpret_cast<double *>(ws + (BETA ? lay.off_passB : lay.off_passA));
  int *passf = reinterpret_cast<int *>(ws + (BETA ? lay.off_passfB : lay.off_passfA));

  // ---- prologue ----------------------------------------------------------------------------------------
  for (int x = threadIdx.x; x < g.NB * kBlk; x += kThreads) sm.rmax[x] = (x < O) ? g_rmax[x] : ninf;
  for (int x = threadIdx.x; x < 257 * g.NB; x += kThreads) sm.rmtab[x] = kNegBig;
  for (int x = threadIdx.x; x < g.NBv * 256; x += kThreads) afrag[x] = make_uint4(0u, 0u, 0u, 0u);   // chunk 0
  for (int x = threadIdx.x; x < g.NB * 32; x += kThreads) passd[x] = 0.0;                             // parity 0
  for (int x = threadIdx.x; x < g.NB; x += kThreads) passf[x] = kNegBig;
  if (threadIdx.x < kStages) { mbar_init(sm.full + threadIdx.x, 1); mbar_init(sm.empty + threadIdx.x, kCW); }
  if (threadIdx.x < 2) mbar_init(sm.ubar + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  {
    // -inf padding: rows >= Tn entirely, columns beyond the last valid block of rows < Tn; the seed row
    const int64_t tail0 = (int64_t)Tn * L;
    for (int64_t x = tail0 + threadIdx.x; x < (int64_t)M * L; x += kThreads) lat[x] = ninf;
    const int c0 = g.NBv * kBlk;
    if (c0 < L) {
      const int wcols = L - c0;
      for (int x = threadIdx.x; x < Tn * wcols; x += kThreads) lat[(int64_t)(x / wcols) * L + c0 + x % wcols] = ninf;
    }
    const int seed_row = BETA ? Tn - 1 : 0, seed_col = BETA ? O - 1 : 0;
    float *row = lat + (int64_t)seed_row * L;
    for (int j = threadIdx.x; j < min(L, c0); j += kThreads) row[j] = (j == seed_col) ? match[(int64_t)seed_row * L + j] : ninf;
  }
  __syncthreads();
  if (warp == 0) {
    // seed: outgoing mass of the single start cell = mant0 * 2^F; it is the "row above" of chunk 0
    const int seed_row = BETA ? Tn - 1 : 0, seed_col = BETA ? O - 1 : 0;
    const int Jb = seed_col / kBlk, jj = seed_col % kBlk;
    const int ci = BETA ? kBlk - 1 - jj : jj;
    const int q = BETA ? g.NBv - 1 - Jb : Jb;
    float v = match[(int64_t)seed_row * L + seed_col];
    if (!BETA) v += sm.rmax[seed_col];   // alpha masses carry the best transition of their own vertex
    const float v2 = v * kLog2e;
    if (v2 > -1.0e30f) {
      const int F = (int)ceilf(v2);
      const float mant0 = exp2f(v2 - (float)F);  // in (0.5, 1]
      passd[(size_t)q * 32 + lane] = (double)mant0 * push[((size_t)Jb * 32 + ci) * 32 + lane];
      if (lane == 0) {
        passf[q] = F;
        sm.rmtab[0 * g.NB + q] = F + 1;
        // row 0 of the chunk-0 fragment tile of block q: value mant0/2 in frame F+1, at K index = vertex offset jj
        write_frag_elem(afrag + (size_t)q * 256, 0, jj, 0.5f * mant0);
      }
    }
  }
  __syncthreads();

  // ---- passes ------------------------------------------------------------------------------------------
  Ring ring;
  ring.n = 0; ring.issued = 0;
  ring.dbg = dbg && blockIdx.x == 0 && (threadIdx.x == kCW * 32 || threadIdx.x == kCW * 32 + 7 * 32);
  ring.t_prod = ring.t_full = ring.t_comp = ring.t_p2 = 0;
  {
    int tt = 0;
    for (int J = 1; J < g.NBv; J++) tt += J - max(0, J - g.band);
    ring.total = tt * g.NP;
  }
  ring.next.reset(g.band);
  const bool is_gemm = warp >= kCW;
  const int cw = warp & (kCW - 1);
  long long t_p1 = 0, t_p2 = 0;
  // The two roles run separate copies of the (pass, block) loops so that neither carries the other's registers;
  // they meet at the CTA barrier (same barrier id, same count, two program locations).
  auto cta_sync = [] { asm volatile("bar.sync 0;" ::: "memory"); };

  if (is_gemm) {
    // ================================ GEMM warps ================================
    const bool producer = (warp == kCW) && (lane == 0);
    const int gt = threadIdx.x - kCW * 32;
    GemmState gs;
    for (int p = 0; p < g.NP; p++) {
      if (p > 0) {
        for (int x = gt; x < g.NB; x += kCW * 32) sm.rmtab[x] = sm.rmtab[256 * g.NB + x];
      }
      if (gt < kCW) sm.prog[gt] = 0;
      {
        float *xo = sm.xbuf + (size_t)cw * kTileF;          // far sums of block 0: none
        for (int x = lane; x < kTileF; x += 32) xo[x] = 0.f;
        sm.fbuf[cw * 32 + lane] = kNegBig;
#pragma unroll
        for (int i = 0; i < 4; i++) gs.Fp[i] = kNegBig;
        if (producer && p == 0) {
          const int J0 = BETA ? g.NBv - 1 : 0;
          mbar_expect_tx(sm.ubar + 0, 8192);
          bulk_g2s(sm.ut, push + (size_t)J0 * 1024, 8192, sm.ubar + 0);
        }
      }
      cta_sync();
      for (int q = 0; q < g.NBv; q++) {
        const int ustep = p * g.NBv + q;
        if (producer && (q + 1 < g.NBv || p + 1 < g.NP)) {   // push table of the next block
          const int un = ustep + 1;
          const int qn = (q + 1 < g.NBv) ? q + 1 : 0;
          const int Jn = BETA ? g.NBv - 1 - qn : qn;
          mbar_expect_tx(sm.ubar + (un & 1), 8192);
          bulk_g2s(sm.ut + (size_t)(un & 1) * 1024, push + (size_t)Jn * 1024, 8192, sm.ubar + (un & 1));
        }
        if (q + 1 < g.NBv) gemm_phase1<BETA>(gs, ring, g, sm, tiles, afrag, lay, p, q + 1, cw, lane, producer);
        cta_sync();
        long long cp2 = ring.dbg ? clock64() : 0;
        if (q + 1 < g.NBv) gemm_phase2<BETA>(gs, ring, g, sm, tiles, afrag, lay, p, q + 1, cw, lane, producer);
        if (ring.dbg) ring.t_p2 += clock64() - cp2;
        cta_sync();
      }
    }
    if (ring.dbg) printf("[dp3 gemm warp %d %s] produce %lld  wait-full %lld  compute %lld  phase2 %lld\n", warp,
                         BETA ? "beta" : "alpha", ring.t_prod, ring.t_full, ring.t_comp, ring.t_p2);
  } else {
    // ================================ chain warps ================================
    for (int p = 0; p < g.NP; p++) {
      const int c = p * kCW + cw;
      const bool active = c < g.NCv;
      if (active) {     // emissions of block 0
        const int J0 = BETA ? g.NBv - 1 : 0;
        float *ion = sm.io + (size_t)cw * kTileF;
        const int j = kBlk * J0 + lane;
        for (int rr = 0; rr < 32; rr++) {
          const int sr = c * 32 + rr;
          const int tr = BETA ? Tn - 2 - sr : 1 + sr;
          if (sr < g.nsteps && j < L) cp_async_f32(ion + rr * kPitch + lane, match + (int64_t)tr * L + j);
          else ion[rr * kPitch + lane] = ninf;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      cta_sync();
      for (int q = 0; q < g.NBv; q++) {
        const int ustep = p * g.NBv + q;
        long long t0 = dbg ? clock64() : 0;
        if (active) chain_phase1<BETA>(g, sm, match, passd, passf, p, q, cw, lane, ustep);
        long long t1 = dbg ? clock64() : 0;
        cta_sync();
        long long t2 = dbg ? clock64() : 0;
        if (active) chain_phase2<BETA>(g, sm, lat, afrag, p, q, cw, lane);
        if (dbg) { t_p1 += t1 - t0; t_p2 += clock64() - t2; }
        cta_sync();
      }
    }
  }
  if (dbg && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 7 * 32))
    printf("[dp3 chain warp %d %s] phase1 %lld  phase2 %lld\n", warp, BETA ? "beta" : "alpha", t_p1, t_p2);
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd((unsigned long long *)&g_dbg[BETA ? 2 : 0], (unsigned long long)t_p1);
    atomicAdd((unsigned long long *)&g_dbg[BETA ? 3 : 1], (unsigned long long)t_p2);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
dag_alpha_beta_colmajor_kernel(const float *__restrict__ match, const int64_t *__restrict__ olen,
                               const int64_t *__restrict__ tlen, float *__restrict__ alpha, float *__restrict__ beta,
                               unsigned char *__restrict__ ws, int M, int L, int Tl, TileLayout lay,
                               int32_t *__restrict__ status, int dbg) {
  extern __shared__ __align__(128) unsigned char dp3_smem[];
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t latsz = (int64_t)M * L;
  float *dst = (is_beta ? beta : alpha) + b * latsz;
  int st = DAGB200_ST_OK;
  if (Tn < 2 || O < 2) st = DAGB200_ST_LEN_LT2;
  else if (O < Tn || O > L || Tn > M) st = DAGB200_ST_GRAPH_SMALL;
  if (st != DAGB200_ST_OK) {
    for (int64_t x = threadIdx.x; x < latsz; x += kThreads) dst[x] = neg_inf_f();
    if (status && threadIdx.x == 0 && !is_beta) status[b] = st;
    return;
  }
  if (status && threadIdx.x == 0 && !is_beta) status[b] = DAGB200_ST_OK;
  Smem sm;
  unsigned char *p = dp3_smem;
  sm.ring = reinterpret_cast<uint4 *>(p);   p += (size_t)kStages * kTileBytes;
  sm.ut = reinterpret_cast<double *>(p);    p += 2 * 8192;
  sm.hand = reinterpret_cast<double *>(p);  p += kCW * 32 * sizeof(double);
  sm.full = reinterpret_cast<uint64_t *>(p);  p += kStages * 8;
  sm.empty = reinterpret_cast<uint64_t *>(p); p += kStages * 8;
  sm.ubar = reinterpret_cast<uint64_t *>(p);  p += 2 * 8;
  sm.xbuf = reinterpret_cast<float *>(p);   p += (size_t)2 * kCW * kTileF * 4;
  sm.io = reinterpret_cast<float *>(p);     p += (size_t)2 * kCW * kTileF * 4;
  sm.fbuf = reinterpret_cast<int *>(p);     p += kCW * 32 * 4;
  sm.prog = reinterpret_cast<int *>(p);     p += kCW * 4;
  sm.tanchor = reinterpret_cast<int *>(p);  p += kCW * 4;
  sm.rmax = reinterpret_cast<float *>(p);   p += (size_t)lay.NB * kBlk * 4;
  sm.rmtab = reinterpret_cast<int *>(p);
  const float *m = match + b * latsz;
  unsigned char *wsb = ws + (size_t)b * lay.sample_bytes;
  if (is_beta) colmajor_dir<true>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl, dbg != 0);
  else colmajor_dir<false>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl, dbg != 0);
}

}  // namespace dp3

size_t dp3_smem_bytes(int M, int L) {
  using namespace dp3;
  TileLayout lay = TileLayout::make(L, M);
  return (size_t)kStages * kTileBytes + 2 * 8192 + kCW * 32 * 8 + (2 * kStages + 2) * 8 +
         (size_t)4 * kCW * kTileF * 4 + kCW * 32 * 4 + 2 * kCW * 4 + (size_t)lay.NB * kBlk * 4 + (size_t)257 * lay.NB * 4;
}

bool dp3_supported(int M, int L) { return L >= 1 && M >= 2 && dp3_smem_bytes(M, L) <= 227 * 1024; }

int launch_dag_prep(const float *links, const int64_t *olen, void *workspace, int B, int M, int L, int Tl, int fmt,
                    cudaStream_t st);

int launch_alpha_beta_colmajor(const float *match, const float *links, const int64_t *olen, const int64_t *tlen,
                               float *alpha, float *beta, int B, int M, int L, int Tl, bool grad, void *workspace,
                               int32_t *status, cudaStream_t st) {
  using namespace dp3;
  prof_mark(0, st);
  int rc = launch_dag_prep(links, olen, workspace, B, M, L, Tl, 0, st);
  if (rc) return rc;
  prof_mark(1, st);
  TileLayout lay = TileLayout::make(L, M);
  dim3 grid(B, grad ? 2 : 1);
  const size_t smem = dp3_smem_bytes(M, L);
  static const bool dbg = getenv("DAGB200_DP3_DEBUG") != nullptr;
  cudaFuncSetAttribute(dag_alpha_beta_colmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dag_alpha_beta_colmajor_kernel<<<grid, kThreads, smem, st>>>(match, olen, tlen, alpha, beta, (unsigned char *)workspace,
                                                               M, L, Tl, lay, status, dbg ? 1 : 0);
  DAGB200_CHECK_LAUNCH("dag_alpha_beta_colmajor_kernel");
  prof_mark(2, st);
  if (dbg) {
    long long h[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_dbg, sizeof(h));
    fprintf(stderr, "[dp3 dbg] cumulative cycles CTA0 thread0: alpha phase1 %lld phase2 %lld | beta phase1 %lld phase2 %lld | "
            "column sweeps: alpha w0 %lld w7 %lld beta w0 %lld w7 %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  }
  return 0;
}

}  // namespace dagb200
