// dag_links.cu -- the transition log-probabilities of the DAG straight from the link heads' queries and keys
// (SURVEY section 8(f) rank 1): the forward of `extract_links` + `extract_valid_links`
// (DASpeech/models/s2t_conformer_dag.py:171-212, 140-155) as one kernel on tcgen05.
//
// Reference, per utterance b, vertex i, head c (H heads of F features), successor k (j = i + k + 1 < O_b, k < T):
//   s[c][k]  = q[b,i,c,:] . key[b,j,c,:] / sqrt(F)                      einsum("bicf,bjcf->bijc") -> [B,L,L,H] fp32 (2.1 GB at C2)
//   lp[c][k] = s[c][k] - logsumexp_k s[c][k]                             gather of the band, masked log_softmax over successors
//   links[k] = logsumexp_c (lp[c][k] + log_gates[b,i,c])                 mixture over the heads
// several [B,L,T,H] temporaries.  Here nothing is materialised but the [B,L,T] result, [B,H,L] row normalisers and the
// operands converted once (bf16 hi + lo, in the layout the MMAs read):
//   links_convert_kernel          q, key -> operand stores (scale 1/sqrt(F) log2e folded into q)
//   extract_links_tcgen05_kernel<0>  one CTA per (128-vertex row tile, head): S = Q K^T tile by tile (128 x 64, K = F) on the
//           tensor cores -- tcgen05.mma kind::f16, both operands split bf16 hi/lo (3 MMAs per k16 step: 2^-16 relative,
//           fp32 accumulate in TMEM) -- and the online maximum / sum of every row (thread = row = TMEM lane) -> lse[c];
//   extract_links_tcgen05_kernel<1>  one CTA per (row tile, pair of 64-vertex destination blocks): the same tiles again for
//           every head, P[j] += exp(s - lse[c] + log_gate[c]); links = log P, written into the band links[i][j-i-1].
// The scores are recomputed instead of stored.  Tile operands arrive by cp.async while the previous tile is in its
// epilogue; one thread issues the MMAs, completion through tcgen05.commit -> mbarrier.
// Mixture weights below e^-87 of a row's total flush to 0, i.e. such a transition comes out as -inf where the reference
// returns a finite value below -87 (the same contract as the blocked recurrences).
#include <algorithm>
#include <cstdint>

#include "common.cuh"
#include "../../include/dagb200.h"

namespace dagb200 {
namespace lk {

constexpr int kRows = 128;      // source vertices per CTA = M of the MMA = TMEM lanes
constexpr int kCols = 64;       // destination vertices per tile = N of the MMA
constexpr int kGroup = 2;       // destination blocks per CTA of the mixture launch
constexpr int kThreads = 160;   // warps 0-3: rows (epilogue), warp 4: MMA issue; all five stage operands
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// shared-memory matrix descriptor, K-major, no swizzle (start address, LBO = distance between the two 8-element K core
// matrices of one MMA, SBO = distance between 8-row groups; 16-byte units; version 1 = Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::f16: D fp32, A / B bf16, both K-major, M = 128, N = 64
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
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
__device__ __forceinline__ void proxy_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
    float2 hf = __bfloat1622float2(hh);
    __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
    h[e] = *reinterpret_cast<uint32_t *>(&hh);
    l[e] = *reinterpret_cast<uint32_t *>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// Operand store: the projections converted ONCE into the layout the MMAs read -- per (utterance, head)
// [plane bf16 hi | lo][k-core][Lp vertices][8 bf16 along K] (Lp = L rounded up to the row tile, zero rows beyond L), the
// 1/sqrt(F) log2e scale folded into Q.  A tile operand is then 2 F/8 contiguous runs of nrows x 16 bytes, copied into
// shared memory by cp.async (no registers, no conversion in the tile loop) while the previous tile is in its epilogue.
__global__ void __launch_bounds__(256)
links_convert_kernel(const float *__restrict__ src, unsigned char *__restrict__ dst, int L, int Lp, int H, int F, float scale,
                     size_t nitems) {
  const int F8 = F >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x; item < nitems; item += stride) {
    // item = ((b * H + c) * F8 + kc) * Lp + v : consecutive threads -> consecutive vertices (coalesced 16-byte stores)
    const int v = (int)(item % Lp);
    size_t rest = item / Lp;
    const int kc = (int)(rest % F8); rest /= F8;
    const int c = (int)(rest % H);
    const size_t b = rest / H;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; e++) x[e] = 0.f;
    if (v < L) {
      const float4 *p = reinterpret_cast<const float4 *>(src + ((b * L + v) * H + c) * F + 8 * kc);
      const float4 a = __ldg(p), bb = __ldg(p + 1);
      x[0] = a.x * scale; x[1] = a.y * scale; x[2] = a.z * scale; x[3] = a.w * scale;
      x[4] = bb.x * scale; x[5] = bb.y * scale; x[6] = bb.z * scale; x[7] = bb.w * scale;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    unsigned char *o = dst + ((((b * H + c) * 2) * F8 + kc) * (size_t)Lp + v) * 16;
    *reinterpret_cast<uint4 *>(o) = hi;
    *reinterpret_cast<uint4 *>(o + (size_t)F8 * Lp * 16) = lo;
  }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// start the copy of `nrows` vertices from `v0` of one (utterance, head) operand store into an MMA operand buffer
__device__ __forceinline__ void stage_async(unsigned char *dst, const unsigned char *__restrict__ conv, int v0, int nrows,
                                            int Lp, int F8) {
  const int nitems = 2 * F8 * nrows;
  for (int item = threadIdx.x; item < nitems; item += kThreads) {
    const int pk = item / nrows, r = item - pk * nrows;       // pk = plane * F8 + kc
    cp_async16(dst + ((size_t)pk * nrows + r) * 16, conv + ((size_t)pk * Lp + v0 + r) * 16);
  }
}

// S[128 x 64] = Q K^T into TMEM (one thread issues), completion on `bar`
__device__ __forceinline__ void issue_tile(uint32_t tmem_d, uint32_t q_u32, uint32_t k_u32, int F, uint64_t *bar) {
  const int F8 = F >> 3;
  const uint64_t dA = umma_desc(0, kRows * 16, 128), dB = umma_desc(0, kCols * 16, 128);
  const uint32_t aplane = (uint32_t)(F8 * kRows * 16) >> 4, bplane = (uint32_t)(F8 * kCols * 16) >> 4;
  const uint32_t akstep = (uint32_t)(2 * kRows * 16) >> 4, bkstep = (uint32_t)(2 * kCols * 16) >> 4;
  const uint32_t a0 = q_u32 >> 4, b0 = k_u32 >> 4;
  for (int ks = 0; ks < (F >> 4); ks++) {
    const uint64_t ahi = dA | (uint64_t)(a0 + ks * akstep), alo = dA | (uint64_t)(a0 + aplane + ks * akstep);
    const uint64_t bhi = dB | (uint64_t)(b0 + ks * bkstep), blo = dB | (uint64_t)(b0 + bplane + ks * bkstep);
    umma_f16(tmem_d, ahi, bhi, ks > 0 ? 1u : 0u);
    umma_f16(tmem_d, alo, bhi, 1u);
    umma_f16(tmem_d, ahi, blo, 1u);
  }
  umma_commit(bar);
}

// MODE 0 (grid: row tile, head, utterance): pass 1 of ONE head -> stats[b][c][i].
// MODE 1 (grid: row tile x column group, utterance): pass 2 of kGroup destination blocks, all heads -> links.
// Two launches instead of one CTA per row tile doing everything: the first row tile of an utterance has 16 destination
// blocks x 8 heads x 2 passes = 256 tiles against 32 for the last one, and the kernel ran as long as its heaviest CTA.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 3)
extract_links_tcgen05_kernel(const unsigned char *__restrict__ qconv, const unsigned char *__restrict__ kconv,
                             const float *__restrict__ log_gates, const int64_t *__restrict__ olen, float *__restrict__ stats,
                             float *__restrict__ links, int L, int Lp, int H, int F, int T, int ngroups) {
  extern __shared__ __align__(128) unsigned char lk_smem[];
  const int b = MODE == 0 ? blockIdx.z : blockIdx.y;
  const int rt = MODE == 0 ? blockIdx.x : blockIdx.x / ngroups;
  const int grp = MODE == 0 ? 0 : blockIdx.x % ngroups;
  const int i0 = rt * kRows;
  const int O = min((int)olen[b], L);
  if (i0 >= O - 1) return;                       // no vertex of this tile has a successor: the rows stay -inf
  // destination blocks that hold a successor of some vertex of the tile: j in [i0 + 1, min(O - 1, i0 + 127 + T)]
  const int jb_all_lo = (i0 + 1) / kCols, jb_all_hi = min(O - 1, i0 + kRows - 1 + T) / kCols;
  const int jb_lo = MODE == 0 ? jb_all_lo : jb_all_lo + grp * kGroup;
  const int jb_hi = MODE == 0 ? jb_all_hi : min(jb_all_hi, jb_lo + kGroup - 1);
  if (jb_lo > jb_hi) return;                     // no block in this group
  const int F8 = F >> 3;
  unsigned char *qs = lk_smem;                                   // [2][F8][128][16 B]
  unsigned char *ks = qs + (size_t)2 * F8 * kRows * 16;          // [2][F8][64][16 B]
  float *lse = reinterpret_cast<float *>(ks + (size_t)2 * F8 * kCols * 16);   // [H][128]: log2-domain normaliser - log2 gate
  uint64_t *bar = reinterpret_cast<uint64_t *>(lse + (size_t)H * kRows);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t head_bytes = (size_t)2 * F8 * Lp * 16;            // one (utterance, head) operand store
  const unsigned char *qb = qconv + (size_t)b * H * head_bytes, *kb = kconv + (size_t)b * H * head_bytes;

  // tiles of this CTA in order: MODE 0: (my head, jb_lo..jb_hi); MODE 1: for every block of the group, every head
  const int ntiles = MODE == 0 ? jb_hi - jb_lo + 1 : (jb_hi - jb_lo + 1) * H;
  auto tile_c = [&](int t) { return MODE == 0 ? (int)blockIdx.y : t % H; };
  auto tile_jb = [&](int t) { return MODE == 0 ? jb_lo + t : jb_lo + t / H; };
  auto stage_tile = [&](int t) {                 // asynchronous: returns at once
    const int c = tile_c(t);
    if (MODE == 1 || t == 0) stage_async(qs, qb + (size_t)c * head_bytes, i0, kRows, Lp, F8);
    stage_async(ks, kb + (size_t)c * head_bytes, tile_jb(t) * kCols, kCols, Lp, F8);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_tile(0);

  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (MODE == 1) {                                 // the statistics of my rows, all heads (written by the MODE 0 launch)
    for (int x = threadIdx.x; x < H * kRows; x += kThreads) {
      const int c = x / kRows, rr = x - c * kRows;
      lse[x] = (i0 + rr < L) ? __ldg(stats + ((size_t)b * H + c) * L + i0 + rr) : __int_as_float(0x7f800000);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t q_u32 = smem_u32(qs), k_u32 = smem_u32(ks);
  uint32_t phase = 0;

  const int r = threadIdx.x;                     // row of the tile (threads 0..127)
  const int i = i0 + r;
  const bool rowlive = r < kRows && i < O - 1;   // has at least one successor (T >= 1)
  const uint32_t lanebase = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  float m = neg_inf_f(), l = 0.f;                // MODE 0: online maximum / sum of my row over its successors
  float P[MODE == 1 ? kCols : 1];                // MODE 1: mixture weights of the current destination block
#pragma unroll
  for (int n = 0; n < (MODE == 1 ? kCols : 1); n++) P[n] = 0.f;

  for (int t = 0; t < ntiles; t++) {
    const int c = tile_c(t), jb = tile_jb(t);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    proxy_fence_async_smem();
    tc_fence_before();
    __syncthreads();                               // operands landed; every row is done with the previous accumulator
    if (warp == 4 && lane == 0) {
      tc_fence_after();
      issue_tile(tmem, q_u32, k_u32, F, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (t + 1 < ntiles) stage_tile(t + 1);         // the MMAs have read the buffers: refill them under the epilogue
    if (warp < 4) {
      if (MODE == 0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          float v[32];
          tmem_ld32(lanebase + 32 * h, v);
          const int jbase = jb * kCols + 32 * h;
          float tm = neg_inf_f();
#pragma unroll
          for (int n = 0; n < 32; n++) {
            const int j = jbase + n;
            const bool ok = rowlive && j > i && j < O && j - i - 1 < T;
            v[n] = ok ? v[n] : neg_inf_f();
            tm = fmaxf(tm, v[n]);
          }
          if (tm > neg_inf_f()) {
            const float mn = fmaxf(m, tm);
            float sum = 0.f;
#pragma unroll
            for (int n = 0; n < 32; n++) sum += ex2(v[n] - mn);     // ex2(-inf) = 0
            l = l * ex2(m - mn) + sum;
            m = mn;
          }
        }
      } else {
        const float z = lse[c * kRows + r];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          float v[32];
          tmem_ld32(lanebase + 32 * h, v);
#pragma unroll
          for (int n = 0; n < 32; n++) P[32 * h + n] += ex2(v[n] - z);   // z = +inf for a dead row / head: adds 0
        }
        if (c == H - 1) {                          // the block is complete: ln of the mixture into the band
          if (rowlive) {
            float *row = links + ((size_t)b * L + i) * T;
#pragma unroll
            for (int n = 0; n < kCols; n++) {
              const int j = jb * kCols + n;
              if (j > i && j < O && j - i - 1 < T) row[j - i - 1] = P[n] > 0.f ? log2f(P[n]) * kLn2 : neg_inf_f();
            }
          }
#pragma unroll
          for (int n = 0; n < kCols; n++) P[n] = 0.f;
        }
      }
    }
  }
  if (MODE == 0 && warp < 4 && i < L) {
    // log2 of the head's normaliser minus the log2 gate: the mixture pass subtracts it from the score
    const int c = blockIdx.y;
    const float lg = rowlive ? __ldg(log_gates + ((size_t)b * L + i) * H + c) * kLog2e : 0.f;
    stats[((size_t)b * H + c) * L + i] = (rowlive && l > 0.f) ? m + log2f(l) - lg : __int_as_float(0x7f800000);   // +inf: adds 0
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

__global__ void fill_neg_inf_kernel(float *__restrict__ p, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = neg_inf_f();
}

}  // namespace lk
}  // namespace dagb200

using namespace dagb200;

// q, key: [B][L][H][F] fp32 (the reshaped outputs of query_linear / key_linear), log_gates: [B][L][H] fp32
// (log_softmax of gate_linear), output_length[b] = number of non-pad positions, links: [B][L][T] fp32, every element
// written; workspace: dagb200_extract_links_workspace_bytes(B, L, H, F) bytes of device scratch (the converted operand
// stores and the per-head row normalisers).  F a multiple of 16, 16 <= F <= 128; H <= 64.
extern "C" size_t dagb200_extract_links_workspace_bytes(int B, int L, int H, int F) {
  if (B <= 0 || L <= 0 || H <= 0 || F <= 0) return 0;
  const size_t Lp = (size_t)(L + lk::kRows - 1) / lk::kRows * lk::kRows;
  const size_t conv = (size_t)B * H * Lp * F * 4;              // hi + lo planes of bf16 = 4 bytes per element
  const size_t stats = ((size_t)B * H * L * 4 + 255) / 256 * 256;
  return 2 * conv + stats;
}

extern "C" int dagb200_extract_links(const float *q, const float *key, const float *log_gates, const int64_t *output_length,
                                     float *links, int B, int L, int H, int F, int T, void *workspace, size_t workspace_bytes,
                                     void *stream) {
  using namespace lk;
  if (B < 0 || L < 1 || H < 1 || H > 64 || F < 16 || F > 128 || (F & 15) || T < 0 || B > 65535) {
    set_error("extract_links: bad shape (F a multiple of 16 in [16, 128], 1 <= H <= 64)");
    return DAGB200_EINVAL;
  }
  if (B == 0 || T == 0) return 0;
  if (!workspace || workspace_bytes < dagb200_extract_links_workspace_bytes(B, L, H, F)) {
    set_error("extract_links: workspace of %zu bytes required", dagb200_extract_links_workspace_bytes(B, L, H, F));
    return DAGB200_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * L * T;
  fill_neg_inf_kernel<<<(int)std::min<size_t>((n + 1023) / 1024, (size_t)8 * sm_count()), 256, 0, st>>>(links, n);
  const int F8 = F >> 3;
  const int Lp = (L + kRows - 1) / kRows * kRows;
  const size_t conv = (size_t)B * H * Lp * F * 4;
  unsigned char *qconv = (unsigned char *)workspace, *kconv = qconv + conv;
  float *stats = reinterpret_cast<float *>(kconv + conv);
  {
    const size_t nitems = (size_t)B * H * F8 * Lp;
    const int grid = (int)std::min<size_t>((nitems + 255) / 256, (size_t)16 * sm_count());
    links_convert_kernel<<<grid, 256, 0, st>>>(q, qconv, L, Lp, H, F, (1.f / sqrtf((float)F)) * lk::kLog2e, nitems);
    links_convert_kernel<<<grid, 256, 0, st>>>(key, kconv, L, Lp, H, F, 1.f, nitems);
    DAGB200_CHECK_LAUNCH("links_convert_kernel");
  }
  const size_t smem = (size_t)2 * F8 * (kRows + kCols) * 16 + (size_t)H * kRows * 4 + 64;
  cudaFuncSetAttribute(extract_links_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(extract_links_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rts = Lp / kRows;
  const int ngroups = ((L + kCols - 1) / kCols + kGroup - 1) / kGroup;
  extract_links_tcgen05_kernel<0><<<dim3(rts, H, B), kThreads, smem, st>>>(qconv, kconv, log_gates, output_length, stats,
                                                                          links, L, Lp, H, F, T, ngroups);
  DAGB200_CHECK_LAUNCH("extract_links_tcgen05_kernel<0>");
  extract_links_tcgen05_kernel<1><<<dim3(rts * ngroups, B), kThreads, smem, st>>>(qconv, kconv, log_gates, output_length,
                                                                                 stats, links, L, Lp, H, F, T, ngroups);
  DAGB200_CHECK_LAUNCH("extract_links_tcgen05_kernel<1>");
  return 0;
}
