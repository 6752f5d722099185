// dag_dp4.cu -- column-major blocked alpha / beta recurrences with the far-predecessor GEMM on tcgen05 (sm_100a).
//
// Vertices in blocks of 32 swept in order, target rows in chunks of 32, a pass = 8 chunks = 256 rows; 8 chain warps run
// the fp64 diagonal blocks of the 8 chunks one column apart (lanes = rows, push formulation), two phases per vertex
// block; the far predecessor sums run on the 5th-generation tensor cores:
//
//   D[128 rows x 32 destination vertices] (TMEM, fp32) = A[128 x 32] (smem, bf16, K-major) * B[32 x 32]^T (smem, bf16)
//   tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 32, K = 16, canonical no-swizzle core-matrix layout (8 rows x 16
//   bytes); both operands split bf16 hi/lo -> 3 MMAs per k16 step, 6 per (source block, 128-row tile); ONE thread
//   issues; completion through tcgen05.commit -> mbarrier.
//   * A operand = previous-row masses, published by the chain warps (lanes = rows, so a row's 8 consecutive vertices
//     are one 16-byte store) in the canonical layout: to global memory for the later destination blocks (streamed back
//     by TMA bulk copies through a 3-stage ring together with the transition tile) and to shared memory for the very
//     next block (phase 2).  One integer frame per (row, source block) as before.
//   * every (destination, source block) product lands in its own TMEM accumulator (two 32-column slots, ping-pong);
//     8 epilogue warps (thread = row) read it back with tcgen05.ld.32x32b, scale it by the exact power of two
//     2^(frame of the source row - running frame) and add it to the row's far sums in registers (online maximum), then
//     hand the sums to the chain warps through shared memory.
//   The per-(row, source block) power-of-two rescale lives on the accumulator side, which is what makes the tcgen05
//   form possible (the A operand is never rewritten per use).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "dag_tiles.cuh"

namespace dagb200 {
namespace dp4 {

constexpr int kThreads = 576;                  // 8 chain warps + 8 epilogue warps + MMA issuer warp + TMA producer warp
constexpr int kCW = 8;                          // chain warps = chunks per pass (GEMM warps: the other 8)
constexpr int kStages = 3;                      // operand ring: {A slab 16 KB, transition tile 4 KB}
constexpr int kStageBytes = 20480;
constexpr int kTSlots = 16;                     // TMEM accumulator slots of 32 columns (all 512 columns)
constexpr int kIssuerWarp = 16;
constexpr int kProducerWarp = 17;
constexpr int kPitch = 33;                      // padded row pitch of 32x32 fp32 tiles in shared memory
constexpr int kTileF = 32 * kPitch;             // floats per staged 32x32 tile
constexpr int kNegBig = -(1 << 20);             // "empty" integer frame
constexpr int kEv = 34;                         // progress events per tile: anchor, then one per column (+1 spare)
constexpr float kLn2Hi = 0.693359375f;          // 355/512: k * kLn2Hi is exact for |k| < 2^15
constexpr float kLn2Lo = -2.12194440e-4f;       // ln2 - kLn2Hi

__device__ long long g_dbg[8];
__device__ long long g_dbg2[16];
__device__ long long g_tl[6][32][4];
__device__ long long g_ev[4][64][4];   // item event log of one block (debug)
__device__ int g_evq = 27;
__device__ long long g_ep[2][32][4];
__device__ int g_evbase[4];

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
// the same with an L2 eviction-priority hint (createpolicy)
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_fractional_evict_last(float fraction) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(pol) : "f"(fraction));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
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
// 64-bit shared-memory mailbox word (value + tag in the sign bit): one relaxed scalar access each way, no fence
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const void *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.cta.shared.b64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(void *p, unsigned long long v) {
  asm volatile("st.relaxed.cta.shared.b64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v));
}
__device__ __forceinline__ void st_release_s32(int *p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// ---- tcgen05 / TMEM helpers (encodings verified stand-alone in tools/tcgen05_probe.cu) ----------------------
// shared-memory matrix descriptor, K-major, no swizzle: start address, LBO = distance between the two 8-element K
// core matrices of one MMA, SBO = distance between 8-row groups (units of 16 bytes), version = 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::f16: D fp32, A / B bf16, both K-major, M = 128, N = 32
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// writer-side fences towards the async proxy by state space (the generic form above costs a GPU-scope MEMBAR)
__device__ __forceinline__ void proxy_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// thread = row: 32 consecutive fp32 columns of my TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}

struct Smem {
  unsigned char *ring;  // [kStages][kStageBytes] operand ring: A slab [2 planes][4 k-cores][128 rows][16 B], then the
                        //                        transition tile [2 planes][4 k-cores][32 vertices][16 B]
  unsigned char *afresh;// [2 planes][4 k-cores][256 rows][16 B]  the block just finished, as the A operand of phase 2
  double *ut;       // [32*32]             fp64 push table of the current block
  float *xbuf;      // [kCW][kTileF]       far sums in -> masses out (normalised fp32 after phase 1)
  float *io;        // [2][kCW][kTileF]    emission weights in (high words of fp64) -> incoming masses out (high words);
                    //                     written / drained by the epilogue warps one block ahead / behind the chain
  float *rmxs;      // [2][kCW][32]        per-column transition maximum folded into the emission weights (alpha: 0 where a vertex has no successor)
  double *hand;     // [2][kCW + 1][32]    row c = what chunk c receives (predecessor sums of the last row of the chunk above; row 0: of
                    //                     the previous pass, posted by chunk 0 itself from global memory): mailbox words
                    //                     (the sums are >= 0; the sign bit carries the tag (step >> 1) & 1, slot = step & 1)
  int *fbuf;        // [kCW][32]           far frames of the rows of the current tile
  int *prog;        // [kCW]               progress counters of the chain warps (events)
  int *tanchor;     // [kCW]               fp64 frame of each chain warp's current tile
  short *rmtab;     // [257][NB]           per (row of the pass, block) integer upper bound of log2(outgoing mass)
  uint64_t *full;   // [kStages]           ring stage filled (TMA)
  uint64_t *empty;  // [kStages]           ring stage consumed (tcgen05.commit)
  uint64_t *tfull;  // [kTSlots]           TMEM accumulator slot written (tcgen05.commit)
  uint64_t *tempty; // [kTSlots]           (unused: replaced by `done`)
  int *done;        // [kCW]               per epilogue warp: item count up to which its accumulator reads are complete
                    //                     (the issuer polls it only when its cached minimum is too small for the slot it
                    //                     wants to overwrite -- an mbarrier try-wait per item cost ~90 cycles on the issue path)
  uint64_t *ubar;   // [1]
  uint32_t *tmem;   // TMEM base address
};

struct Geo {
  int O, Tn, M, L, NB, NBv, nsteps, NCv, NP, band, Mr;
  bool hint;                 // L2 eviction hints on the operand stream (see colmajor_dir)
  int qpin;
  uint64_t pol_a, pol_t;
  bool dbg;
  int dbgmode;
};

// a tile without a single lattice cell (below the diagonal j >= t): nothing flows through it
template <bool BETA>
__device__ __forceinline__ bool tile_geo_dead(const Geo &g, int c, int J) {
  const int smax = min(c * 32 + 31, g.nsteps - 1);
  const int tmin = BETA ? g.Tn - 2 - smax : 1 + c * 32;
  return min(kBlk * J + 31, g.O - 1) < tmin;
}

// frames are stored as int16 (|frame| < 32767 covers lattice values down to -22 000 nats); -32768 = no mass
__device__ __forceinline__ void rm_store(short *p, int v) { *p = (short)(v <= kNegBig ? -32768 : max(min(v, 32767), -32767)); }
__device__ __forceinline__ int rm_load(const short *p) { const int v = *p; return v == -32768 ? kNegBig : v; }

// high word of the fp64 emission weight exp(em + rmx) (21 significant bits; 0 when it vanishes)
__device__ __forceinline__ int ew_word(float em, float rmx) {
  // branch-free, so that the 16 independent weights of a pre-pass batch interleave
  const float w2 = (em + rmx) * kLog2e;
  const float wc = fmaxf(w2, -2000.f);                        // -inf / NaN -> far below the fp64 range -> 0
  const float wf = floorf(wc);
  float fb;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(fb) : "f"(wc - wf));   // [1, 2]
  const int fbi = __float_as_int(fb);
  const int iw = (int)wf + ((fbi >> 23) - 127);
  const int word = ((iw + 1023) << 20) | ((fbi & 0x7fffff) >> 3);
  return (iw > -1023 && iw < 1024) ? word : 0;
}

// ---- epilogue-warp side jobs (lanes = columns): emission weights of a block BEFORE the chain warps reach it, lattice
// values of a block AFTER they left it.  Both touch global memory, off the chain warps' critical path.
// Split in two so that no global-memory latency sits at the head of a phase: `stage` starts asynchronous copies
// (cp.async, 4 bytes each: any L) of the raw emissions into the tile and returns the column's transition maximum;
// `convert` (same warp, after the phase's accumulator takes) turns them into weights in place.
template <bool BETA>
__device__ __forceinline__ float epi_prepass_stage(const Geo &g, const Smem &sm, const float *__restrict__ match,
                                                   const float *__restrict__ g_rmax, int p, int qn, int ew, int lane) {
  const int c = p * kCW + ew;
  const int Jn = BETA ? g.NBv - 1 - qn : qn;
  const int jn = kBlk * Jn + lane;
  float *iot = sm.io + ((size_t)(qn & 1) * kCW + ew) * kTileF;
  const float rmxn = jn < g.O ? __ldg(g_rmax + jn) : neg_inf_f();
  const int jc = min(jn, g.L - 1);
#pragma unroll 8
  for (int rr = 0; rr < 32; rr++) {
    const int sr = c * 32 + rr;
    const int tr = min(max(BETA ? g.Tn - 2 - sr : 1 + sr, 0), g.M - 1);
    cp_async_f32(iot + rr * kPitch + lane, match + (int64_t)tr * g.L + jc);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return rmxn;
}
template <bool BETA>
__device__ __forceinline__ void epi_prepass_convert(const Geo &g, const Smem &sm, float rmxn, int p, int qn, int ew, int lane) {
  const int c = p * kCW + ew;
  const int Jn = BETA ? g.NBv - 1 - qn : qn;
  const int jn = kBlk * Jn + lane;
  float *iot = sm.io + ((size_t)(qn & 1) * kCW + ew) * kTileF;
  // alpha masses carry exp(rmax) of their own vertex; a vertex without successors (rmax = -inf) still has a forward
  // value and its mass meets only zero transitions, so its weight is taken without the factor
  if (!BETA && jn < g.O && rmxn == neg_inf_f()) rmxn = 0.f;
  sm.rmxs[((qn & 1) * kCW + ew) * 32 + lane] = rmxn;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll 1
  for (int h = 0; h < 2; h++) {          // two batches of 16 independent weights
    float em[16];
#pragma unroll
    for (int r = 0; r < 16; r++) em[r] = iot[(16 * h + r) * kPitch + lane];
#pragma unroll
    for (int r = 0; r < 16; r++) {
      const int rr = 16 * h + r, sr = c * 32 + rr;
      const int tr = BETA ? g.Tn - 2 - sr : 1 + sr;
      const int word = ew_word(em[r], rmxn);
      iot[rr * kPitch + lane] = __int_as_float((sr < g.nsteps && jn >= tr && jn < g.O) ? word : 0);
    }
  }
}
// The row of block q that lane `lane` of chain warp cw just finished (normalised masses v[0..31], K index = vertex
// offset) becomes consumer row cr = 32 cw + lane + 1 of the pass: bf16 hi/lo, 16 bytes per 8-vertex core, into the
// shared-memory operand of phase 2 (rows < 256) and into the global operand store (all later destination blocks).
__device__ __forceinline__ void publish_row(const Geo &g, const Smem &sm, unsigned char *__restrict__ aop, const float *mrow,
                                            int p, int q, int cw, int lane, bool zero) {
  const int cr = 32 * cw + lane + 1;
  const int grow = 256 * p + cr;
#pragma unroll
  for (int kc = 0; kc < 4; kc++) {
    uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
    if (!zero) {
      split_bf16x2(mrow[8 * kc + 0], mrow[8 * kc + 1], hi.x, lo.x);
      split_bf16x2(mrow[8 * kc + 2], mrow[8 * kc + 3], hi.y, lo.y);
      split_bf16x2(mrow[8 * kc + 4], mrow[8 * kc + 5], hi.z, lo.z);
      split_bf16x2(mrow[8 * kc + 6], mrow[8 * kc + 7], hi.w, lo.w);
    }
    if (cr < 256) {
      *reinterpret_cast<uint4 *>(sm.afresh + ((size_t)(0 * 4 + kc) * 256 + cr) * 16) = hi;
      *reinterpret_cast<uint4 *>(sm.afresh + ((size_t)(1 * 4 + kc) * 256 + cr) * 16) = lo;
    }
    if (grow < g.Mr) {
      unsigned char *slab = aop + ((size_t)q * (g.Mr >> 7) + (grow >> 7)) * 16384;   // [q][128-row tile]: [plane][kc][row][16 B]
      *reinterpret_cast<uint4 *>(slab + ((size_t)(0 * 4 + kc) * 128 + (grow & 127)) * 16) = hi;
      *reinterpret_cast<uint4 *>(slab + ((size_t)(1 * 4 + kc) * 128 + (grow & 127)) * 16) = lo;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// phase 1 of a chain warp: the 32x32 diagonal block of (chunk c, block q)
template <bool BETA>
__device__ __forceinline__ void chain_phase1(const Geo &g, const Smem &sm, const float *__restrict__ match,
                                             double *__restrict__ passd, int *__restrict__ passf,
                                             unsigned char *__restrict__ aop, int p, int q, int cw, int lane, int ustep,
                                             int &Fh_pre, double &hv_pre) {
  const int c = p * kCW + cw;
  const int J = BETA ? g.NBv - 1 - q : q;
  const int jbase = kBlk * J;
  const int s = c * 32 + lane;
  const bool rowvalid = s < g.nsteps;
  const int t = BETA ? g.Tn - 2 - s : 1 + s;
  float *xb = sm.xbuf + (size_t)cw * kTileF;
  float *iob = sm.io + ((size_t)(q & 1) * kCW + cw) * kTileF;
  float *mrow = xb + lane * kPitch;
  float *iow = iob + lane * kPitch;
  const float ninf = neg_inf_f();
  const bool feeds_next = (c + 1 < g.NCv);       // somebody consumes my last row
  const bool to_pass = feeds_next && cw == kCW - 1;

  const long long sgE = g.dbg ? clock64() : 0;
  // (2) frames: far frames of my rows, the frame handed from the chunk above, the fp64 frame of this tile
  const int FI = sm.fbuf[cw * 32 + lane];
  const int maxFI = __reduce_max_sync(0xffffffffu, FI);
  int Fh;
  double hv = 0.0;                               // cw == 0: lane cj holds the handed sum of column cj
  int known = 0;                                 // cw > 0: last progress value seen of the warp above
  const int evbase = q * kEv;
  unsigned long long *hand_r = reinterpret_cast<unsigned long long *>(sm.hand) + ((ustep & 1) * (kCW + 1) + cw) * 32;
  unsigned long long *hand_w = hand_r + 32;
  const unsigned long long tagbit = (unsigned long long)((ustep >> 1) & 1) << 63;
  if (cw == 0) {
    // handed in from the previous pass (or the seed) through global memory; fetched one block ahead
    Fh = Fh_pre;
    hv = hv_pre;
    st_relaxed_u64(hand_r + lane, (unsigned long long)__double_as_longlong(hv) | tagbit);   // same path as the other chunks
    __syncwarp();
    int pn = p, qn = q + 1;
    if (qn >= g.NBv) { pn = p + 1; qn = 0; }
    if (pn < g.NP) {
      Fh_pre = passf[(pn & 1) * g.NB + qn];
      hv_pre = passd[((size_t)(pn & 1) * g.NB + qn) * 32 + lane];
    }
  } else {
    do { known = ld_acquire_s32(sm.prog + cw - 1); } while (known < evbase + 1);
    Fh = sm.tanchor[cw - 1];
  }
  const bool dead = tile_geo_dead<BETA>(g, c, J) || max(maxFI, Fh) <= kNegBig;
  const int Ft = dead ? kNegBig : max(maxFI, Fh - 600);
  if (feeds_next && !to_pass) {
    if (lane == 0) sm.tanchor[cw] = Ft;
    if (dead) st_relaxed_u64(hand_w + lane, tagbit);          // tagged zeros
    __syncwarp();
    if (lane == 0) st_release_s32(sm.prog + cw, evbase + 1);
  }
  if (to_pass) {
    if (lane == 0) passf[((p + 1) & 1) * g.NB + q] = Ft;
    if (dead) passd[((size_t)((p + 1) & 1) * g.NB + q) * 32 + lane] = 0.0;
  }
  __syncwarp();
  if (dead) {   // nothing reaches this tile: -inf lattice values, zero masses
#pragma unroll 4
    for (int k = 0; k < 32; k++) { iow[k] = neg_inf_f(); mrow[k] = 0.f; }
    rm_store(sm.rmtab + (cw * 32 + lane + 1) * g.NB + q, kNegBig);
    publish_row(g, sm, aop, mrow, p, q, cw, lane, true);
    proxy_fence_async_smem();
    proxy_fence_async_global();
    __syncwarp();
    return;
  }
  // (3) the fp64 push table of this block
  const bool swdbg = g.dbg && blockIdx.x == 0 && lane == 0 && (cw == 0 || cw == 7);
  const long long sg0 = swdbg ? clock64() : 0;
  mbar_wait(sm.ubar, ustep & 1);
  long long sgA = 0, sgB = 0;
  const long long sg1 = swdbg ? clock64() : 0;
  const double *ut = sm.ut;
  const double hs = pow2d(Fh - Ft);              // handed sums are in the frame of the tile above
  const double xs = pow2d(FI - Ft);              // far sums of my row are in the frame FI
  int maxhi = 0;

  // (4) column sweep: four groups of 8 columns (runtime loop keeps the code small)
  const long long sw0 = swdbg ? clock64() : 0;
#pragma unroll 1
  for (int G = 0; G < 4; G++) {
    const long long sa0 = swdbg ? clock64() : 0;
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.0;
    // predecessor sums from the completed groups of my own row (masses re-read with 21 significant bits)
#pragma unroll 4
    for (int ci = 0; ci < 8 * G; ci++) {
      const int w = __float_as_int(mrow[BETA ? 31 - ci : ci]);
      const double md = __hiloint2double(w, 0);
      const double2 *ur = reinterpret_cast<const double2 *>(ut + ci * 32 + 8 * G);
      const double2 u0 = ur[0], u1 = ur[1], u2 = ur[2], u3 = ur[3];
      a[0] = fma(md, u0.x, a[0]); a[1] = fma(md, u0.y, a[1]); a[2] = fma(md, u1.x, a[2]); a[3] = fma(md, u1.y, a[3]);
      a[4] = fma(md, u2.x, a[4]); a[5] = fma(md, u2.y, a[5]); a[6] = fma(md, u3.x, a[6]); a[7] = fma(md, u3.y, a[7]);
    }
    if (swdbg) sgA += clock64() - sa0;
    unsigned long long zraw = ld_relaxed_u64(hand_r + 8 * G);
    // far sums (in the tile frame) and emission weights of the group: off the column-to-column dependency path
    double fd[8];
    int ewd[8];
#pragma unroll
    for (int K = 0; K < 8; K++) {
      const int jj = BETA ? 31 - (8 * G + K) : 8 * G + K;
      fd[K] = (double)mrow[jj] * xs;
      ewd[K] = __float_as_int(iow[jj]);
    }
#pragma unroll
    for (int K = 0; K < 8; K++) {
      const int cj = 8 * G + K;
      const int jj = BETA ? 31 - cj : cj;
      // hand my own last row to the chunk below
      if (feeds_next && lane == 31) {
        if (to_pass) passd[((size_t)((p + 1) & 1) * g.NB + q) * 32 + cj] = a[K];
        else st_relaxed_u64(hand_w + cj, (unsigned long long)__double_as_longlong(a[K]) | tagbit);
      }
      // what the row above hands to this column (lane 0 uses it)
      while ((zraw ^ tagbit) >> 63) zraw = ld_relaxed_u64(hand_r + cj);
      const double z = __longlong_as_double((long long)(zraw & 0x7fffffffffffffffull)) * hs;
      if (K < 7) zraw = ld_relaxed_u64(hand_r + cj + 1);       // usually already there: the chunk above runs ahead
      double rm = __shfl_up_sync(0xffffffffu, a[K], 1);
      if (lane == 0) rm = z;
      const double m = (fd[K] + rm) * __hiloint2double(ewd[K], 0);   // emission weight: 0 outside the lattice
      const int mhi = __double2hiint(m);
      mrow[jj] = __int_as_float(mhi);                  // outgoing mass, 21 significant bits, frame Ft
      maxhi = max(maxhi, mhi);
      // push into the later columns of my row
      if (K < 7) {
        const double *ur = ut + cj * 32 + 8 * G;
#pragma unroll
        for (int k2 = K + 1; k2 < 8; k2++) a[k2] = fma(m, ur[k2], a[k2]);
      }
    }
  }
  const long long sg2 = swdbg ? clock64() : 0;
  // (5) row frame and normalised masses (value / 2^maxe < 1) for the A-operand fragments
  const int emf = maxhi >> 20;                   // biased exponent of the largest mass of my row (0: none)
  const int maxe = (emf > 0) ? Ft + emf - 1022 : kNegBig;
  rm_store(sm.rmtab + (cw * 32 + lane + 1) * g.NB + q, maxe);
  const float *rmxw = sm.rmxs + ((q & 1) * kCW + cw) * 32;
#pragma unroll 1
  for (int k0 = 0; k0 < 32; k0 += 8) {           // batches of 8: all loads, the arithmetic, all stores
    int w[8];
    float rx[8], out[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { w[k] = __float_as_int(mrow[k0 + k]); rx[k] = BETA ? 0.f : rmxw[k0 + k]; }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int we = w[k] >> 20;
      // lattice value = log(outgoing mass) - rmax (alpha) / log(outgoing mass) (beta), off the dependency path; branch-free
      const float fl = (float)(we - 1023 + Ft);
      float lg;
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(__int_as_float(0x3f800000 | ((w[k] & 0xfffff) << 3))));   // mantissa in [1, 2)
      float o = fmaf(lg, 0.6931471805599453f, fl * kLn2Lo) + fl * kLn2Hi;
      if (!BETA) o -= rx[k];
      out[k] = (we > 0) ? o : neg_inf_f();
      const int fe = we - emf + 126;
      w[k] = (we > 0 && fe > 0) ? ((fe << 23) | ((w[k] & 0xfffff) << 3)) : 0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { iow[k0 + k] = out[k]; mrow[k0 + k] = __int_as_float(w[k]); }
  }
  // (6) the row as A operand of the tensor cores (shared memory for the next block, global memory for the later ones)
  const long long sg3 = swdbg ? clock64() : 0;
  publish_row(g, sm, aop, mrow, p, q, cw, lane, false);
  const long long sg4 = swdbg ? clock64() : 0;
  proxy_fence_async_smem();
  proxy_fence_async_global();
  __syncwarp();
  if (swdbg) {
    const int o = (cw == 7 ? 8 : 0);
    atomicAdd((unsigned long long *)&g_dbg2[o + 0], (unsigned long long)(sg1 - sg0));      // wait push table
    atomicAdd((unsigned long long *)&g_dbg2[o + 1], (unsigned long long)sgA);              // re-entry loops
    atomicAdd((unsigned long long *)&g_dbg2[o + 2], (unsigned long long)(sg2 - sw0 - sgA));// column steps
    atomicAdd((unsigned long long *)&g_dbg2[o + 3], (unsigned long long)(sg3 - sg2));      // normalisation
    atomicAdd((unsigned long long *)&g_dbg2[o + 4], (unsigned long long)(sg4 - sg3));      // publish
    atomicAdd((unsigned long long *)&g_dbg2[o + 5], (unsigned long long)(clock64() - sg4));// fence
    atomicAdd((unsigned long long *)&g_dbg2[o + 6], (unsigned long long)(sg0 - sgE));      // entry .. push-table wait
  }
}

// phase 2 of a chain warp: the lattice rows of the block (lanes = columns, coalesced)
template <bool BETA>
__device__ __forceinline__ void chain_phase2(const Geo &g, const Smem &sm, float *__restrict__ lat, int p, int q, int cw,
                                             int lane) {
  const int c = p * kCW + cw;
  const int J = BETA ? g.NBv - 1 - q : q;
  const float *iot = sm.io + ((size_t)(q & 1) * kCW + cw) * kTileF;
  const int j = kBlk * J + lane;
  const int rl = min(32, g.nsteps - c * 32) - 1;
  if (j < g.L) {
#pragma unroll 4
    for (int rr = 0; rr <= rl; rr++) {
      const int sr = c * 32 + rr;
      const int tr = BETA ? g.Tn - 2 - sr : 1 + sr;
      lat[(int64_t)tr * g.L + j] = iot[rr * kPitch + lane];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Tensor-core side.  Work items in a fixed order: for every destination block J = 1 .. NBv-1 of a pass, the far source
// blocks qs = qlo(J) .. J-2 (operands from the ring), then the source J-1 (A operand = sm.afresh); every source once
// per active 128-row tile mt.  Item n uses ring stage n % kStages and TMEM slot n & 1.
struct ItemSeq {
  int p, J, qs, mt;
  bool done;
  __device__ __forceinline__ void start(const Geo &g) { p = 0; J = 1; qs = max(0, 1 - g.band); mt = 0; done = (g.NBv < 2); }
  __device__ __forceinline__ int ntiles(const Geo &g) const { return (g.NCv - kCW * p > 4) ? 2 : 1; }
  __device__ __forceinline__ void advance(const Geo &g) {
    if (++mt < ntiles(g)) return;
    mt = 0;
    if (++qs <= J - 1) return;
    if (++J >= g.NBv) {
      J = 1;
      if (++p >= g.NP) { done = true; return; }
    }
    qs = max(0, J - g.band);
  }
};

// TMA producer (one thread): loads items in order while (a) the ring has room, (b) the source block of the item is
// complete (far sources <= safe_q; the last item of a destination only needs its transition tile) and (c) the slot
// it waits for is freed by MMAs that are issued before the next CTA barrier (`limit` = items consumed by then + ring).
struct Producer {
  ItemSeq next;
  int issued;
};
template <bool BETA>
__device__ __forceinline__ void producer_run(Producer &pr, const Geo &g, const Smem &sm, const unsigned char *__restrict__ aop,
                                             const unsigned char *__restrict__ tiles, const TileLayout &lay, int cur_p,
                                             int safe_q, int limit) {
  while (!pr.next.done && pr.issued < limit) {
    const ItemSeq &it = pr.next;
    const bool last = (it.qs == it.J - 1);
    if (it.p > cur_p) break;
    if (!last && it.qs > safe_q) break;
    const int st = pr.issued % kStages;
    if (pr.issued >= kStages) {       // the stage was last used by item issued - kStages: its MMAs must be complete
      const int prev = pr.issued - kStages;
      mbar_wait(sm.tfull + prev % kTSlots, (prev / kTSlots) & 1);
    }
    unsigned char *dst = sm.ring + (size_t)st * kStageBytes;
    const int Jd = BETA ? g.NBv - 1 - it.J : it.J;           // vertex-block index of the destination (sweep J)
    const int Js = BETA ? g.NBv - 1 - it.qs : it.qs;         // ... of the source
    const unsigned char *tsrc = tiles + (BETA ? lay.idxB(Jd, Js) : lay.idxA(Js, Jd)) * (size_t)kTileBytes;
    mbar_expect_tx(sm.full + st, last ? 4096u : (uint32_t)kStageBytes);
    const unsigned char *asrc = aop + ((size_t)it.qs * (g.Mr >> 7) + 2 * it.p + it.mt) * 16384;
    if (g.hint) {
      bulk_g2s_hint(dst + 16384, tsrc, 4096, sm.full + st, g.pol_t);
      if (!last) bulk_g2s_hint(dst, asrc, 16384, sm.full + st, it.qs < g.qpin ? g.pol_a : g.pol_t);
    } else {
      bulk_g2s(dst + 16384, tsrc, 4096, sm.full + st);
      if (!last) bulk_g2s(dst, asrc, 16384, sm.full + st);
    }
    pr.next.advance(g);
    pr.issued++;
  }
}

// MMA issuer (one thread): the 6 MMAs of item n (A: ring stage or sm.afresh), then the two commits
struct Issuer {
  int n;             // items whose MMAs have been issued
  bool dbg, logq;
  int logbase;
  int st, stpar;     // ring stage of the next item and its mbarrier parity
  int cmin;          // cached minimum of Smem::done over the epilogue warps
  long long t_full, t_tempty, t_p1, t_p2;
};
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred;
}
// The whole issuer warp runs this with warp-uniform values (the descriptors then live in uniform registers and the
// UTCHMMA issue needs no per-instruction vector->uniform transfer); one elected lane issues.
__device__ __forceinline__ void issuer_mma(Issuer &is, const Smem &sm, uint32_t tmem_base, uint32_t ring_u32, uint32_t afresh_u32,
                                           bool from_afresh, int mt) {
  const int st = is.st, pp = is.n % kTSlots;
  long long c0 = is.dbg ? clock64() : 0;
  mbar_wait(sm.full + st, is.stpar);
  long long c1 = is.dbg ? clock64() : 0;
  {
    // slot pp was last used by item n - kTSlots: every epilogue warp must be past that item (a warp counts the items of
    // the other 128-row tile as soon as it skips them)
    const int need = is.n - kTSlots + 1;
    while (is.cmin < need) {
      const volatile int *dn = sm.done;
      is.cmin = min(min(min(dn[0], dn[1]), min(dn[2], dn[3])), min(min(dn[4], dn[5]), min(dn[6], dn[7])));
    }
  }
  long long c2 = is.dbg ? clock64() : 0;
  if (is.dbg) { is.t_full += c1 - c0; is.t_tempty += c2 - c1; }
  tc_fence_after();
  const uint32_t stage = ring_u32 + st * kStageBytes;
  const uint32_t d = tmem_base + pp * 32;
  // descriptors differ from a per-kernel constant only in the 14-bit start address field (bits 0..13, 16-byte units)
  const uint64_t dA = from_afresh ? umma_desc(0, 4096, 128) : umma_desc(0, 2048, 128);
  const uint64_t dB = umma_desc(0, 512, 128);
  const uint32_t a0 = (from_afresh ? afresh_u32 + mt * 2048 : stage) >> 4;
  const uint32_t aplane = (from_afresh ? 16384u : 8192u) >> 4, akstep = (from_afresh ? 8192u : 4096u) >> 4;
  const uint32_t b0 = (stage + 16384) >> 4;
  const long long c3 = is.dbg ? clock64() : 0;
  long long c4 = 0, c5 = 0;
  if (elect_one()) {
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
      // shared-memory addresses are < 256 KB: the 14-bit field never carries
      const uint64_t ahi = dA | (uint64_t)(a0 + ks * akstep);
      const uint64_t alo = dA | (uint64_t)(a0 + aplane + ks * akstep);
      const uint64_t bhi = dB | (uint64_t)(b0 + ks * 64);
      const uint64_t blo = dB | (uint64_t)(b0 + 128 + ks * 64);
      umma_f16(d, ahi, bhi, ks > 0 ? 1u : 0u);
      umma_f16(d, alo, bhi, 1u);
      umma_f16(d, ahi, blo, 1u);
    }
    if (is.dbg) c4 = clock64();
    // ONE commit per item: "accumulator slot complete" also tells the producer that the item's ring stage has been read
    // (it waits on the slot barrier of the item that used the stage last)
    umma_commit(sm.tfull + pp);
    if (is.dbg) c5 = clock64();
  }
  __syncwarp();
  if (is.dbg && is.logq) {
    const int i = is.n - is.logbase;
    if (i >= 0 && i < 64) { g_ev[0][i][0] = c0; g_ev[0][i][1] = c1; g_ev[0][i][2] = c2; g_ev[0][i][3] = clock64();
                            g_ev[3][i][1] = c3; g_ev[3][i][2] = c4; g_ev[3][i][3] = c5; }
  }
  is.n++;
  if (++is.st == kStages) { is.st = 0; is.stpar ^= 1; }
}

// epilogue warp state: far sums of my row for the current destination block, online maximum of the source frames
struct Epi {
  float acc[32];
  int F;
  int n;            // items seen so far (all tiles), mirrors Issuer::n
  int logbase;
};

// take one item: wait for its accumulator, read my row, release the slot, add with the power-of-two scale
__device__ __forceinline__ void epi_take(Epi &e, const Smem &sm, uint32_t tmem_base, int quarter, int Fs, bool mine, int logrow = -1) {
  const int pp = e.n % kTSlots, k = e.n / kTSlots;
  const int li = e.n - e.logbase;
  e.n++;
  if (!mine) {
    if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile int *>(sm.done + (threadIdx.x >> 5) - kCW) = e.n;
    return;
  }
  const bool lg = logrow >= 0 && li >= 0 && li < 64;
  if (lg) g_ev[logrow][li][0] = clock64();
  mbar_wait(sm.tfull + pp, k & 1);
  if (lg) g_ev[logrow][li][1] = clock64();
  tc_fence_after();
  float v[32];
  tmem_ld32(tmem_base + pp * 32 + ((uint32_t)(quarter * 32) << 16), v);
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) *reinterpret_cast<volatile int *>(sm.done + (threadIdx.x >> 5) - kCW) = e.n;   // reads of items < e.n complete
  if (lg) g_ev[logrow][li][2] = clock64();
  if (Fs > kNegBig) {
    if (Fs > e.F) {   // the new source block sets the frame: acc = acc * 2^(F - Fs) + v  (the scale is 0 when acc is empty)
      const float rs = pow2i(e.F - Fs);
#pragma unroll
      for (int j = 0; j < 32; j++) e.acc[j] = fmaf(e.acc[j], rs, v[j]);
      e.F = Fs;
    } else {
      const float sc = pow2i(Fs - e.F);
#pragma unroll
      for (int j = 0; j < 32; j++) e.acc[j] = fmaf(v[j], sc, e.acc[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// One direction of one utterance.
template <bool BETA, bool DBG>
__device__ void colmajor_dir(const float *__restrict__ match, float *__restrict__ lat, unsigned char *__restrict__ ws,
                             const TileLayout &lay, const Smem &sm, int O, int Tn, int M, int L, int Tl, int dbgi, float l2frac) {
  // DBG is a compile-time switch: the production kernel carries none of the timers below
  const bool dbg = DBG && (dbgi & 1) != 0;  // bit 0: all role timers and item logs; bit 1: only the per-block barrier timeline
  const bool tlm = DBG && (dbgi & 3) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float ninf = neg_inf_f();
  Geo g;
  g.O = O; g.Tn = Tn; g.M = M; g.L = L; g.NB = lay.NB;
  g.NBv = (O + kBlk - 1) / kBlk;
  g.nsteps = Tn - 1;
  g.NCv = (g.nsteps + 31) / 32;
  g.NP = (g.NCv + kCW - 1) / kCW;
  g.band = band_blocks(Tl);
  g.Mr = lay.Mr;
  g.dbg = dbg;
  g.dbgmode = dbgi;
  // The A slabs of all source blocks are re-streamed for every later destination block: a cyclic pattern over ~1 MB per
  // CTA (128 MB per launch at C2, above the 126 MB L2) that LRU turns into misses -- ncu: 1.65 GB of DRAM reads per
  // launch against 0.34 GB of compulsory input, and the late blocks run at the speed of that stream.  The slab of source
  // block qs is read by 30 - qs destinations, so the slabs of the FIRST `qpin` source blocks are requested evict_last
  // (they stay), everything else -- the later slabs and the transition tiles, used by two consecutive items and never
  // again -- evict_first.  l2frac = 0 switches the hints off; qpin = l2frac when >= 1 (DAGB200_DP4_L2FRAC).
  g.hint = l2frac > 0.f;
  g.qpin = (int)l2frac;
  g.pol_a = g.hint ? l2_policy_fractional_evict_last(1.0f) : 0ull;
  g.pol_t = g.hint ? l2_policy_evict_first() : 0ull;
  const float *g_rmax = reinterpret_cast<const float *>(ws + lay.off_rmax);
  const double *push = reinterpret_cast<const double *>(ws + (BETA ? lay.off_pushB : lay.off_pushA));
  const unsigned char *tiles = ws + (BETA ? lay.off_tilesB : lay.off_tilesA);
  unsigned char *aop = ws + (BETA ? lay.off_aopB : lay.off_aopA);
  double *passd = reinterpret_cast<double *>(ws + (BETA ? lay.off_passB : lay.off_passA));
  int *passf = reinterpret_cast<int *>(ws + (BETA ? lay.off_passfB : lay.off_passfA));

  // ---- prologue ----------------------------------------------------------------------------------------
  for (int x = threadIdx.x; x < 257 * g.NB; x += kThreads) sm.rmtab[x] = (short)-32768;
  for (int x = threadIdx.x; x < g.NB * 8; x += kThreads)               // consumer row 0 (the seed row): zeros
    *reinterpret_cast<uint4 *>(aop + (size_t)(x >> 3) * (g.Mr >> 7) * 16384 + (size_t)(x & 7) * 2048) = make_uint4(0u, 0u, 0u, 0u);
  for (int x = threadIdx.x; x < g.NB * 32; x += kThreads) passd[x] = 0.0;                             // parity 0
  for (int x = threadIdx.x; x < g.NB; x += kThreads) passf[x] = kNegBig;
  for (int x = threadIdx.x; x < 2 * (kCW + 1) * 32; x += kThreads)    // mailboxes: tag 1 = nothing posted for steps 0, 1
    reinterpret_cast<unsigned long long *>(sm.hand)[x] = 1ull << 63;
  if (threadIdx.x < kCW) sm.done[threadIdx.x] = 0;
  if (threadIdx.x < kStages) { mbar_init(sm.full + threadIdx.x, 1); mbar_init(sm.empty + threadIdx.x, 1); }
  if (threadIdx.x < kTSlots) { mbar_init(sm.tfull + threadIdx.x, 1); mbar_init(sm.tempty + threadIdx.x, 4); }
  if (threadIdx.x == 0) mbar_init(sm.ubar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == kIssuerWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(sm.tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm.tmem;
  if (warp == 0) {
    // seed: outgoing mass of the single start cell = mant0 * 2^F; it is the "row above" of chunk 0
    const int seed_row = BETA ? Tn - 1 : 0, seed_col = BETA ? O - 1 : 0;
    const int Jb = seed_col / kBlk, jj = seed_col % kBlk;
    const int ci = BETA ? kBlk - 1 - jj : jj;
    const int q = BETA ? g.NBv - 1 - Jb : Jb;
    float v = match[(int64_t)seed_row * L + seed_col];
    if (!BETA) v += g_rmax[seed_col];    // alpha masses carry the best transition of their own vertex
    const float v2 = v * kLog2e;
    if (v2 > -1.0e30f) {
      const int F = (int)ceilf(v2);
      const float mant0 = exp2f(v2 - (float)F);  // in (0.5, 1]
      passd[(size_t)q * 32 + lane] = (double)mant0 * push[((size_t)Jb * 32 + ci) * 32 + lane];
      if (lane == 0) {
        passf[q] = F;
        rm_store(sm.rmtab + q, F + 1);
        // consumer row 0 of block q: value mant0/2 in frame F+1 at K index = vertex offset jj
        const __nv_bfloat16 h = __float2bfloat16_rn(0.5f * mant0);
        const __nv_bfloat16 l = __float2bfloat16_rn(0.5f * mant0 - __bfloat162float(h));
        const int kc = jj >> 3, e = jj & 7;
        unsigned char *slab = aop + (size_t)q * (g.Mr >> 7) * 16384;
        reinterpret_cast<__nv_bfloat16 *>(slab + (size_t)(0 * 4 + kc) * 2048)[e] = h;
        reinterpret_cast<__nv_bfloat16 *>(slab + (size_t)(1 * 4 + kc) * 2048)[e] = l;
      }
    }
    proxy_fence_async();
  }
  __syncthreads();

  // ---- passes ------------------------------------------------------------------------------------------
  const int cw = warp & (kCW - 1);
  long long t_p1 = 0, t_p2 = 0;
  // The roles run separate copies of the (pass, block) loops; they meet at the CTA barrier.
  auto cta_sync = [] { asm volatile("bar.sync 0;" ::: "memory"); };
  long long t_b1 = 0, t_b2 = 0;
  // debug timeline: arrival / release clocks of six observer threads at both barriers of every block (CTA 0, alpha)
  const int tl_role = (!tlm || BETA || blockIdx.x != 0 || lane != 0) ? -1
                      : (warp == 0 ? 0 : warp == 7 ? 1 : warp == 8 ? 2 : warp == 15 ? 3 : warp == kIssuerWarp ? 4 : warp == kProducerWarp ? 5 : -1);
  auto cta_sync1 = [&](int q) {
    if (tl_role >= 0 && q < 32) g_tl[tl_role][q][0] = clock64();
    asm volatile("bar.sync 0;" ::: "memory");
    if (tl_role >= 0 && q < 32) g_tl[tl_role][q][1] = clock64();
  };
  auto cta_sync2 = [&](int q) {
    if (tl_role >= 0 && q < 32) g_tl[tl_role][q][2] = clock64();
    asm volatile("bar.sync 0;" ::: "memory");
    if (tl_role >= 0 && q < 32) g_tl[tl_role][q][3] = clock64();
  };

  if (warp == kProducerWarp) {
    // ================================ TMA producer (one thread) ========================================
    Producer pr;
    pr.next.start(g);
    pr.issued = 0;
    int consumed = 0;                              // items whose MMAs are issued before the next CTA barrier
    for (int p = 0; p < g.NP; p++) {
      const int ntl = (g.NCv - kCW * p > 4) ? 2 : 1;
      if (lane == 0 && p == 0) {
        const int J0 = BETA ? g.NBv - 1 : 0;
        mbar_expect_tx(sm.ubar, 8192);
        bulk_g2s(sm.ut, push + (size_t)J0 * 1024, 8192, sm.ubar);
      }
      cta_sync();
      for (int q = 0; q < g.NBv; q++) {
        const int J = q + 1;
        const bool have = J < g.NBv;
        if (lane == 0) {
          if (have) consumed += max(0, (J - 1) - max(0, J - g.band)) * ntl;
          producer_run<BETA>(pr, g, sm, aop, tiles, lay, p, q - 1, consumed + kStages);
        }
        __syncwarp();
        cta_sync1(q);
        if (lane == 0) {
          if (have || p + 1 < g.NP) {   // push table of the next block (the chain warps are done with this one)
            const int qn = have ? q + 1 : 0;
            const int Jn = BETA ? g.NBv - 1 - qn : qn;
            mbar_expect_tx(sm.ubar, 8192);
            bulk_g2s(sm.ut, push + (size_t)Jn * 1024, 8192, sm.ubar);
          }
          if (have) consumed += ntl;
          producer_run<BETA>(pr, g, sm, aop, tiles, lay, p, q, consumed + kStages);
        }
        __syncwarp();
        cta_sync2(q);
      }
    }
  } else if (warp == kIssuerWarp) {
    // ================================ MMA issuer (one warp, one elected lane issues) =====================
    Issuer is;
    is.n = 0;
    is.dbg = dbg && blockIdx.x == 0 && lane == 0;
    is.t_full = is.t_tempty = is.t_p1 = is.t_p2 = 0;
    is.logq = false; is.logbase = 0;
    is.st = 0; is.stpar = 0; is.cmin = 0;
    // everything that shapes the item sequence as warp-uniform values
    const uint32_t ring_u32 = __shfl_sync(0xffffffffu, smem_u32(sm.ring), 0);
    const uint32_t afresh_u32 = __shfl_sync(0xffffffffu, smem_u32(sm.afresh), 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const int uNP = __shfl_sync(0xffffffffu, g.NP, 0), uNBv = __shfl_sync(0xffffffffu, g.NBv, 0);
    const int uNCv = __shfl_sync(0xffffffffu, g.NCv, 0), uband = __shfl_sync(0xffffffffu, g.band, 0);
    for (int p = 0; p < uNP; p++) {
      const int ntl = (uNCv - kCW * p > 4) ? 2 : 1;
      cta_sync();
      for (int q = 0; q < uNBv; q++) {
        const int J = q + 1;
        const bool have = J < uNBv;
        long long i0 = is.dbg ? clock64() : 0;
        is.logq = is.dbg && !BETA && q == g_evq;
        if (is.logq) { is.logbase = is.n; g_evbase[0] = is.n; g_ev[3][0][0] = i0; }
        if (have) {
          for (int qs = max(0, J - uband); qs <= J - 2; qs++)
            for (int mt = 0; mt < ntl; mt++) issuer_mma(is, sm, tmem_u, ring_u32, afresh_u32, false, mt);
        }
        long long i1 = is.dbg ? clock64() : 0;
        cta_sync1(q);
        long long i2 = is.dbg ? clock64() : 0;
        is.logq = false;
        if (have) {
          tc_fence_after();
          for (int mt = 0; mt < ntl; mt++) issuer_mma(is, sm, tmem_u, ring_u32, afresh_u32, true, mt);
        }
        if (is.dbg) { is.t_p1 += i1 - i0; is.t_p2 += clock64() - i2; }
        cta_sync2(q);
      }
    }
    if (is.dbg) printf("[dp4 issuer %s] items %d  phase1 %lld  phase2 %lld  wait-full %lld  wait-tmem-empty %lld  bar1 %lld bar2 %lld\n",
                       BETA ? "beta" : "alpha", is.n, is.t_p1, is.t_p2, is.t_full, is.t_tempty, t_b1, t_b2);
  } else if (warp >= kCW) {
    // ================================ epilogue warps: thread = row ======================================
    const int ew = warp - kCW;                    // rows 32 ew .. 32 ew + 31 of the pass
    const int mymt = ew >> 2, quarter = ew & 3;   // TMEM lanes 32 quarter .. (this warp's lane window: warp % 4)
    Epi e;
    e.n = 0;
    e.logbase = 0;
    const bool edbg = dbg && blockIdx.x == 0 && lane == 0 && (ew == 0 || ew == 7);
    long long t_post = 0, t_pre = 0, t_take = 0;
    const int et = threadIdx.x - kCW * 32;
    for (int p = 0; p < g.NP; p++) {
      const int ntl = (g.NCv - kCW * p > 4) ? 2 : 1;
      const int c = p * kCW + ew;
      const bool active = c < g.NCv;
      const bool rowvalid = (c * 32 + lane) < g.nsteps;
      if (p > 0) {
        for (int x = et; x < g.NB; x += kCW * 32) sm.rmtab[x] = sm.rmtab[256 * g.NB + x];
      }
      if (et < kCW) sm.prog[et] = 0;
      {
        float *xo = sm.xbuf + (size_t)ew * kTileF;          // far sums of block 0: none
        for (int x = lane; x < kTileF; x += 32) xo[x] = 0.f;
        sm.fbuf[ew * 32 + lane] = kNegBig;
      }
      if (active) {                                                          // emission weights of block 0
        const float r0 = epi_prepass_stage<BETA>(g, sm, match, g_rmax, p, 0, ew, lane);
        epi_prepass_convert<BETA>(g, sm, r0, p, 0, ew, lane);
      }
      cta_sync();
      for (int q = 0; q < g.NBv; q++) {
        const int J = q + 1;
        const bool have = J < g.NBv;
        if (active && q + 2 < g.NBv) {   // pull the emission rows the next pre-pass reads towards L2
          const int J2 = BETA ? g.NBv - 3 - q : q + 2;
          const int sr = c * 32 + lane;
          const int tr = BETA ? g.Tn - 2 - sr : 1 + sr;
          if (sr < g.nsteps && kBlk * J2 < g.L) asm volatile("prefetch.global.L2 [%0];" ::"l"(match + (int64_t)tr * g.L + kBlk * J2));
        }
        long long e0 = edbg ? clock64() : 0;
        const int epr = (edbg && !BETA && q < 32) ? (ew == 0 ? 0 : 1) : -1;
        if (epr >= 0) g_ep[epr][q][0] = e0;
        float rmx_next = 0.f;
        bool converted = false;
        if (active && have) rmx_next = epi_prepass_stage<BETA>(g, sm, match, g_rmax, p, q + 1, ew, lane);   // raw emissions of the next block
        if (edbg) t_post += clock64() - e0;
        long long e2 = edbg ? clock64() : 0;
        if (epr >= 0) g_ep[epr][q][1] = e2;
        const bool tile_on = active && have && !tile_geo_dead<BETA>(g, c, BETA ? g.NBv - 1 - J : J);
        if (have) {
          // consumer row 0 of the pass (the last row of the previous pass / the seed) of block q for phase 2
          if (ew == 0 && lane < 8) {
            const int pl = lane >> 2, kc = lane & 3;
            const uint4 v = *reinterpret_cast<const uint4 *>(aop + ((size_t)q * (g.Mr >> 7) + 2 * p) * 16384 + (size_t)(pl * 4 + kc) * 2048);
            *reinterpret_cast<uint4 *>(sm.afresh + ((size_t)(pl * 4 + kc) * 256) * 16) = v;
            proxy_fence_async_smem();
          }
#pragma unroll
          for (int j = 0; j < 32; j++) e.acc[j] = 0.f;
          e.F = kNegBig;
          if (epr >= 0) g_ep[epr][q][2] = clock64();
          for (int qs = max(0, J - g.band); qs <= J - 2; qs++) {
            // the weights of the next block are converted in the middle of the takes (their raw emissions have landed by
            // then, and the tail of the phase stays short); phases with few items convert after the loop
            if (!converted && qs == max(0, J - g.band) + 6) {
              if (active) epi_prepass_convert<BETA>(g, sm, rmx_next, p, q + 1, ew, lane);
              converted = true;
            }
            for (int mt = 0; mt < ntl; mt++) {
              const bool mine = (mt == mymt);
              const int Fs = (mine && tile_on && rowvalid) ? rm_load(sm.rmtab + (ew * 32 + lane) * g.NB + qs) : kNegBig;
              const int logrow = (dbg && !BETA && blockIdx.x == 0 && lane == 0 && q == g_evq && (ew == 0 || ew == 4)) ? (ew == 0 ? 1 : 2) : -1;
              if (logrow >= 0 && qs == max(0, J - g.band) && mt == 0) e.logbase = e.n;
              epi_take(e, sm, tmem_base, quarter, Fs, mine, logrow);
              if (logrow >= 0) g_ev[logrow][e.n - 1 - e.logbase < 64 ? e.n - 1 - e.logbase : 63][3] = clock64();
            }
          }
        }
        if (edbg) t_take += clock64() - e2;
        if (epr >= 0) g_ep[epr][q][3] = clock64();
        {
          const long long e3 = edbg ? clock64() : 0;
          if (active && have && !converted) epi_prepass_convert<BETA>(g, sm, rmx_next, p, q + 1, ew, lane);  // ... into weights
          if (edbg) t_pre += clock64() - e3;
        }
        cta_sync1(q);
        if (have) {
          for (int mt = 0; mt < ntl; mt++) {
            const bool mine = (mt == mymt);
            const int Fs = (mine && tile_on && rowvalid) ? rm_load(sm.rmtab + (ew * 32 + lane) * g.NB + (J - 1)) : kNegBig;
            epi_take(e, sm, tmem_base, quarter, Fs, mine);
          }
          if (active) {
            float *xo = sm.xbuf + (size_t)ew * kTileF + lane * kPitch;
#pragma unroll
            for (int j = 0; j < 32; j++) xo[j] = e.acc[j];
            sm.fbuf[ew * 32 + lane] = e.F;
          }
        }
        cta_sync2(q);
      }
    }
    if (edbg) printf("[dp4 epilogue warp %d %s] post %lld  pre %lld  takes(phase 1) %lld  bar1 %lld bar2 %lld\n", ew, BETA ? "beta" : "alpha", t_post, t_pre, t_take, t_b1, t_b2);
  } else {
    // ================================ chain warps ================================
    int Fh_pre = kNegBig;
    double hv_pre = 0.0;
    if (cw == 0) { Fh_pre = passf[0]; hv_pre = passd[lane]; }      // (pass 0, block 0): written by the prologue
    for (int p = 0; p < g.NP; p++) {
      const int c = p * kCW + cw;
      const bool active = c < g.NCv;
      cta_sync();
      for (int q = 0; q < g.NBv; q++) {
        const int ustep = p * g.NBv + q;
        long long t0 = dbg ? clock64() : 0;
        if (active) chain_phase1<BETA>(g, sm, match, passd, passf, aop, p, q, cw, lane, ustep, Fh_pre, hv_pre);
        long long t1 = dbg ? clock64() : 0;
        cta_sync1(q);
        if (active) chain_phase2<BETA>(g, sm, lat, p, q, cw, lane);
        if (dbg) t_p1 += t1 - t0;
        cta_sync2(q);
      }
    }
  }
  if (dbg && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 7 * 32)) {
    const int o = (warp == 7 ? 8 : 0);
    printf("[dp4 chain warp %d %s] bar1 %lld bar2 %lld\n", warp, BETA ? "beta" : "alpha", t_b1, t_b2);
    printf("[dp4 chain warp %d %s] phase1 %lld  phase2 %lld | entry %lld wait-ut %lld reentry %lld columns %lld norm %lld publish %lld fence %lld (alpha+beta)\n",
           warp, BETA ? "beta" : "alpha", t_p1, t_p2, g_dbg2[o + 6], g_dbg2[o + 0], g_dbg2[o + 1], g_dbg2[o + 2], g_dbg2[o + 3], g_dbg2[o + 4], g_dbg2[o + 5]);
  }
  tc_fence_before();
  __syncthreads();
  if (tlm && !dbg && !BETA && blockIdx.x == 0 && threadIdx.x == 0) {
    printf("[dp4 light timeline alpha] per block: issuer's phase-1 length (barrier to barrier), arrivals at barrier 1 relative to its release: chain0 chain7 epi0 epi7 issuer producer; phase-2 length\n");
    for (int q = 1; q < min(g.NBv, 32); q++) {
      const long long rel1 = g_tl[4][q][1], st1 = g_tl[4][q - 1][3];
      printf("  q %2d  p1 %6lld | %6lld %6lld %6lld %6lld %6lld %6lld | p2 %6lld\n", q, rel1 - st1, g_tl[0][q][0] - rel1, g_tl[1][q][0] - rel1,
             g_tl[2][q][0] - rel1, g_tl[3][q][0] - rel1, g_tl[4][q][0] - rel1, g_tl[5][q][0] - rel1, g_tl[4][q][3] - rel1);
    }
  }
  if (dbg && !BETA && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long z0 = g_ev[3][0][0];
    printf("[dp4 items of block %d] issuer: start, full ok, tmem-empty ok, issued | epilogue (owner warp): start wait, tfull ok, ld done+released, fma done\n", g_evq);
    for (int i = 0; i < 56; i++) {
      const int r = (i & 1) ? 2 : 1;
      printf("  item %2d  %6lld %6lld %6lld [pre-issue %6lld mma-issued %6lld committed %6lld] %6lld | %6lld %6lld %6lld %6lld\n", i, g_ev[0][i][0] - z0, g_ev[0][i][1] - z0, g_ev[0][i][2] - z0,
             g_ev[3][i][1] - z0, g_ev[3][i][2] - z0, g_ev[3][i][3] - z0, g_ev[0][i][3] - z0,
             g_ev[r][i][0] - z0, g_ev[r][i][1] - z0, g_ev[r][i][2] - z0, g_ev[r][i][3] - z0);
    }
    printf("[dp4 epilogue warps 0 / 7 per block] prepass start, prepass end, takes start, takes end (relative to chain 0 leaving barrier 2 of the previous block)\n");
    for (int q = 1; q < min(g.NBv, 32); q++) {
      const long long st = g_tl[4][q - 1][3];
      printf("  q %2d  %6lld %6lld %6lld %6lld | %6lld %6lld %6lld %6lld\n", q, g_ep[0][q][0] - st, g_ep[0][q][1] - st, g_ep[0][q][2] - st, g_ep[0][q][3] - st,
             g_ep[1][q][0] - st, g_ep[1][q][1] - st, g_ep[1][q][2] - st, g_ep[1][q][3] - st);
    }
    printf("[dp4 timeline alpha] per block: phase-1 length, then arrival at barrier 1 relative to its release (chain0 chain7 epi0 epi7 issuer producer), phase-2 length, same for barrier 2\n");
    for (int q = 0; q < min(g.NBv, 32); q++) {
      const long long rel1 = g_tl[0][q][1], rel2 = g_tl[0][q][3];
      const long long start1 = q > 0 ? g_tl[0][q - 1][3] : rel1;
      printf("  q %2d  p1 %6lld | %6lld %6lld %6lld %6lld %6lld %6lld | p2 %6lld | %6lld %6lld %6lld %6lld %6lld %6lld\n", q, rel1 - start1,
             g_tl[0][q][0] - rel1, g_tl[1][q][0] - rel1, g_tl[2][q][0] - rel1, g_tl[3][q][0] - rel1, g_tl[4][q][0] - rel1, g_tl[5][q][0] - rel1,
             rel2 - rel1,
             g_tl[0][q][2] - rel2, g_tl[1][q][2] - rel2, g_tl[2][q][2] - rel2, g_tl[3][q][2] - rel2, g_tl[4][q][2] - rel2, g_tl[5][q][2] - rel2);
    }
  }
  if (warp == kIssuerWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

template <bool DBG>
__device__ __forceinline__ void alpha_beta_body(const float *__restrict__ match, const int64_t *__restrict__ olen,
                                                const int64_t *__restrict__ tlen, float *__restrict__ alpha, float *__restrict__ beta,
                                                unsigned char *__restrict__ ws, int M, int L, int Tl, const TileLayout &lay,
                                                int32_t *__restrict__ status, int dbg, float l2frac) {
  extern __shared__ __align__(128) unsigned char dp4_smem[];
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
  unsigned char *p = dp4_smem;
  sm.ring = p;                               p += (size_t)kStages * kStageBytes;
  sm.afresh = p;                             p += 32768;
  sm.ut = reinterpret_cast<double *>(p);     p += 8192;
  sm.hand = reinterpret_cast<double *>(p);   p += 2 * (kCW + 1) * 32 * sizeof(double);
  sm.full = reinterpret_cast<uint64_t *>(p);   p += kStages * 8;
  sm.empty = reinterpret_cast<uint64_t *>(p);  p += kStages * 8;
  sm.tfull = reinterpret_cast<uint64_t *>(p);  p += kTSlots * 8;
  sm.tempty = reinterpret_cast<uint64_t *>(p); p += kTSlots * 8;
  sm.ubar = reinterpret_cast<uint64_t *>(p);   p += 2 * 8;
  sm.tmem = reinterpret_cast<uint32_t *>(p);   p += 16;
  sm.done = reinterpret_cast<int *>(p);        p += kCW * 4;
  sm.xbuf = reinterpret_cast<float *>(p);    p += (size_t)kCW * kTileF * 4;
  sm.io = reinterpret_cast<float *>(p);      p += (size_t)2 * kCW * kTileF * 4;
  sm.rmxs = reinterpret_cast<float *>(p);    p += 2 * kCW * 32 * 4;
  sm.fbuf = reinterpret_cast<int *>(p);      p += kCW * 32 * 4;
  sm.prog = reinterpret_cast<int *>(p);      p += kCW * 4;
  sm.tanchor = reinterpret_cast<int *>(p);   p += kCW * 4;
  sm.rmtab = reinterpret_cast<short *>(p);
  const float *m = match + b * latsz;
  unsigned char *wsb = ws + (size_t)b * lay.sample_bytes;
  if (is_beta) colmajor_dir<true, DBG>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl, dbg, l2frac);
  else colmajor_dir<false, DBG>(m, dst, wsb, lay, sm, O, Tn, M, L, Tl, dbg, l2frac);
}

__global__ void __launch_bounds__(kThreads, 1)
dag_alpha_beta_tcgen05_kernel(const float *__restrict__ match, const int64_t *__restrict__ olen,
                              const int64_t *__restrict__ tlen, float *__restrict__ alpha, float *__restrict__ beta,
                              unsigned char *__restrict__ ws, int M, int L, int Tl, TileLayout lay,
                              int32_t *__restrict__ status, float l2frac) {
  alpha_beta_body<false>(match, olen, tlen, alpha, beta, ws, M, L, Tl, lay, status, 0, l2frac);
}
#ifdef DAGB200_DEBUG_KERNELS
// the same kernel with the in-kernel timers / timelines compiled in (build with -DDAGB200_DEBUG_KERNELS, run with
// DAGB200_DP4_DEBUG != 0); not part of the product library
__global__ void __launch_bounds__(kThreads, 1)
dag_alpha_beta_tcgen05_debug_kernel(const float *__restrict__ match, const int64_t *__restrict__ olen,
                                    const int64_t *__restrict__ tlen, float *__restrict__ alpha, float *__restrict__ beta,
                                    unsigned char *__restrict__ ws, int M, int L, int Tl, TileLayout lay,
                                    int32_t *__restrict__ status, int dbg, float l2frac) {
  alpha_beta_body<true>(match, olen, tlen, alpha, beta, ws, M, L, Tl, lay, status, dbg, l2frac);
}
#endif

}  // namespace dp4

size_t dp4_smem_bytes(int M, int L) {
  using namespace dp4;
  TileLayout lay = TileLayout::make(L, M);
  return (size_t)kStages * kStageBytes + 32768 + 8192 + 2 * (kCW + 1) * 32 * 8 + (2 * kStages + 2 * kTSlots + 2) * 8 + 16 + kCW * 4 +
         (size_t)3 * kCW * kTileF * 4 + 2 * kCW * 32 * 4 + kCW * 32 * 4 + 2 * kCW * 4 + (size_t)257 * lay.NB * 2 + 64;
}

bool dp4_supported(int M, int L) { return L >= 1 && M >= 2 && dp4_smem_bytes(M, L) <= 227 * 1024; }

size_t dp4_workspace_bytes(int B, int M, int L) { return TileLayout::make(L, M).sample_bytes * (size_t)B; }

int launch_dag_prep(const float *links, const int64_t *olen, void *workspace, int B, int M, int L, int Tl,
                    cudaStream_t st);

int launch_alpha_beta_tcgen05(const float *match, const float *links, const int64_t *olen, const int64_t *tlen,
                              float *alpha, float *beta, int B, int M, int L, int Tl, bool grad, void *workspace,
                              int32_t *status, cudaStream_t st) {
  using namespace dp4;
  prof_mark(0, st);
  int rc = launch_dag_prep(links, olen, workspace, B, M, L, Tl, st);
  if (rc) return rc;
  prof_mark(1, st);
  TileLayout lay = TileLayout::make(L, M);
  dim3 grid(B, grad ? 2 : 1);
  const size_t smem = dp4_smem_bytes(M, L);
  // number of leading source blocks whose A slabs are requested evict_last (about 50 MB pinned at B = 64; measured at C2:
  // DRAM reads 1.64 -> 1.09 GB, L2 hit rate 27 -> 42 %, 631 -> 607 us); DAGB200_DP4_L2FRAC=0 switches the hints off
  static const float l2env = getenv("DAGB200_DP4_L2FRAC") ? (float)atof(getenv("DAGB200_DP4_L2FRAC")) : -1.f;
  const float l2frac = l2env >= 0.f ? l2env : (float)(B <= 64 ? 12 : (768 / B > 2 ? 768 / B : 2));
#ifdef DAGB200_DEBUG_KERNELS
  static const int dbg = getenv("DAGB200_DP4_DEBUG") ? atoi(getenv("DAGB200_DP4_DEBUG")) : 0;
  if (dbg) {
    cudaFuncSetAttribute(dag_alpha_beta_tcgen05_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dag_alpha_beta_tcgen05_debug_kernel<<<grid, kThreads, smem, st>>>(match, olen, tlen, alpha, beta, (unsigned char *)workspace,
                                                                      M, L, Tl, lay, status, dbg, l2frac);
    DAGB200_CHECK_LAUNCH("dag_alpha_beta_tcgen05_debug_kernel");
    prof_mark(2, st);
    return 0;
  }
#endif
  cudaFuncSetAttribute(dag_alpha_beta_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dag_alpha_beta_tcgen05_kernel<<<grid, kThreads, smem, st>>>(match, olen, tlen, alpha, beta, (unsigned char *)workspace,
                                                              M, L, Tl, lay, status, l2frac);
  DAGB200_CHECK_LAUNCH("dag_alpha_beta_tcgen05_kernel");
  prof_mark(2, st);
  return 0;
}

}  // namespace dagb200
