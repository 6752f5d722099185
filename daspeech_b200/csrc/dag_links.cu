// dag_links.cu -- the transition log-probabilities of the DAG straight from the link heads' queries and keys
// (SURVEY section 8(f) rank 1): the forward of `extract_links` + `extract_valid_links`
// (DASpeech/models/s2t_conformer_dag.py:171-212, 140-155) as one kernel on tcgen05.
//
// Reference, per utterance b, vertex i, head c (H heads of F features), successor k (j = i + k + 1 < O_b, k < T):
//   s[c][k]  = q[b,i,c,:] . key[b,j,c,:] / sqrt(F)                      einsum("bicf,bjcf->bijc") -> [B,L,L,H] fp32 (2.1 GB at C2)
//   lp[c][k] = s[c][k] - logsumexp_k s[c][k]                             gather of the band, masked log_softmax over successors
//   links[k] = logsumexp_c (lp[c][k] + log_gates[b,i,c])                 mixture over the heads
// several [B,L,T,H] temporaries.  Here a CTA owns 128 consecutive vertices of one utterance and nothing is materialised
// but the [B,L,T] result and [B,H,L] row normalisers (two launches: one CTA per (row tile, head), then one per (row
// tile, pair of destination blocks)):
//   pass 1  for every head: S = Q K^T tile by tile (128 x 64, K = F) on the tensor cores -- tcgen05.mma kind::f16 with
//           both operands split bf16 hi/lo (3 MMAs per k16 step: 2^-16 relative, fp32 accumulate in TMEM) -- and the
//           online maximum / sum of every row (thread = row = TMEM lane) -> lse[c] per row;
//   pass 2  for every 64-column block, for every head: the same tile again, P[j] += exp(s - lse[c] + log_gate[c]);
//           links = log P, written into the banded layout links[i][j-i-1].
// The scores are recomputed instead of stored (12 MMAs of 128x64x16 per tile: the kernel is bound by the exponentials
// and the operand staging, not by the tensor pipe).  Operands are converted from the fp32 projections while they are
// staged into the canonical K-major core-matrix layout; one thread issues, completion through tcgen05.commit -> mbarrier.
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

// `nrows` vertices starting at `v0` of head c as an MMA operand: [plane hi|lo][k-core][row][8 bf16 along K], scaled.
// Four items (32 bytes each) are requested per thread before the first one is converted: the staging of a tile is one
// L2 latency, not one per item (the first version converted item by item and spent two thirds of a tile in here).
__device__ __forceinline__ void stage_operand(unsigned char *dst, const float *__restrict__ src, int v0, int nrows, int L,
                                              int H, int F, int c, float scale) {
  constexpr int NB = 4;
  const int F8 = F >> 3;
  const int nitems = nrows * F8;
  const size_t plane = (size_t)F8 * nrows * 16;
  for (int base = threadIdx.x; base < nitems; base += kThreads * NB) {
    float4 a[NB], b[NB];
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const int item = base + u * kThreads;
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      b[u] = a[u];
      if (item < nitems) {
        const int kc = item / nrows, r = item - kc * nrows;   // consecutive threads: consecutive rows (conflict-free stores)
        const int v = v0 + r;
        if (v < L) {
          const float4 *p = reinterpret_cast<const float4 *>(src + ((size_t)v * H + c) * F + 8 * kc);
          a[u] = __ldg(p);
          b[u] = __ldg(p + 1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const int item = base + u * kThreads;
      if (item < nitems) {
        const int kc = item / nrows, r = item - kc * nrows;
        const float x[8] = {a[u].x * scale, a[u].y * scale, a[u].z * scale, a[u].w * scale,
                            b[u].x * scale, b[u].y * scale, b[u].z * scale, b[u].w * scale};
        uint4 hi, lo;
        split8(x, hi, lo);
        unsigned char *q = dst + ((size_t)kc * nrows + r) * 16;
        *reinterpret_cast<uint4 *>(q) = hi;
        *reinterpret_cast<uint4 *>(q + plane) = lo;
      }
    }
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
constexpr int kGroup = 2;
template <int MODE>
__global__ void __launch_bounds__(kThreads)
extract_links_tcgen05_kernel(const float *__restrict__ q, const float *__restrict__ key, const float *__restrict__ log_gates,
                             const int64_t *__restrict__ olen, float *__restrict__ stats, float *__restrict__ links, int L,
                             int H, int F, int T, int ngroups) {
  extern __shared__ __align__(128) unsigned char lk_smem[];
  const int b = MODE == 0 ? blockIdx.z : blockIdx.y;
  const int rt = MODE == 0 ? blockIdx.x : blockIdx.x / ngroups;
  const int grp = MODE == 0 ? 0 : blockIdx.x % ngroups;
  const int i0 = rt * kRows;
  const int O = min((int)olen[b], L);
  if (i0 >= O - 1) return;                       // no vertex of this tile has a successor: the rows stay -inf
  if (MODE == 1 && (i0 + 1) / kCols + grp * kGroup > min(O - 1, i0 + kRows - 1 + T) / kCols) return;   // no block in this group
  const int F8 = F >> 3;
  unsigned char *qs = lk_smem;                                   // [2][F8][128][16 B]
  unsigned char *ks = qs + (size_t)2 * F8 * kRows * 16;          // [2][F8][64][16 B]
  float *lse = reinterpret_cast<float *>(ks + (size_t)2 * F8 * kCols * 16);   // [H][128]: log2-domain normaliser - log2 gate
  uint64_t *bar = reinterpret_cast<uint64_t *>(lse + (size_t)H * kRows);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *qb = q + (size_t)b * L * H * F, *kb = key + (size_t)b * L * H * F;

  if (threadIdx.x == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
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
  // destination blocks that hold a successor of some vertex of the tile: j in [i0 + 1, min(O - 1, i0 + 127 + T)]
  const int jb_all_lo = (i0 + 1) / kCols, jb_all_hi = min(O - 1, i0 + kRows - 1 + T) / kCols;
  const int jb_lo = MODE == 0 ? jb_all_lo : jb_all_lo + grp * kGroup;
  const int jb_hi = MODE == 0 ? jb_all_hi : min(jb_all_hi, jb_lo + kGroup - 1);
  const float scale = rsqrtf((float)F) * kLog2e;  // scores in log2 units
  const uint32_t lanebase = tmem + ((uint32_t)((warp & 3) * 32) << 16);

  // one tile: stage K, run the MMAs, wait; on return the accumulator of (head c, block jb) sits in TMEM
  auto run_tile = [&](int c, int jb) {
    stage_operand(ks, kb, jb * kCols, kCols, L, H, F, c, 1.f);
    proxy_fence_async_smem();
    tc_fence_before();
    __syncthreads();                               // operands staged; every row is done with the previous accumulator
    if (warp == 4 && lane == 0) {
      tc_fence_after();
      issue_tile(tmem, q_u32, k_u32, F, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
  };

  // ---- pass 1: per head, the online maximum / sum of the row over its successors ---------------------------------
  for (int c = (MODE == 0 ? (int)blockIdx.y : H); c < (MODE == 0 ? (int)blockIdx.y + 1 : H); c++) {
    __syncthreads();                               // the previous head's MMAs are complete (everybody waited on `bar`)
    stage_operand(qs, qb, i0, kRows, L, H, F, c, scale);
    float m = neg_inf_f(), l = 0.f;
    for (int jb = jb_lo; jb <= jb_hi; jb++) {
      run_tile(c, jb);
      if (warp < 4) {
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
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < 32; n++) s += ex2(v[n] - mn);     // ex2(-inf) = 0
            l = l * ex2(m - mn) + s;
            m = mn;
          }
        }
      }
    }
    if (warp < 4) {
      // log2 of the head's normaliser minus the log2 gate: pass 2 subtracts it from the score
      const float lg = (rowlive) ? __ldg(log_gates + ((size_t)b * L + i) * H + c) * kLog2e : 0.f;
      if (i < L) stats[((size_t)b * H + c) * L + i] = (rowlive && l > 0.f) ? m + log2f(l) - lg : __int_as_float(0x7f800000);   // +inf: contributes 0
    }
  }
  if (MODE == 1) {                                 // the statistics of my rows, all heads (written by the MODE 0 launch)
    for (int x = threadIdx.x; x < H * kRows; x += kThreads) {
      const int c = x / kRows, rr = x - c * kRows;
      lse[x] = (i0 + rr < L) ? __ldg(stats + ((size_t)b * H + c) * L + i0 + rr) : __int_as_float(0x7f800000);
    }
  }
  // ---- pass 2: per destination block, the mixture over the heads ---------------------------------------------------
  for (int jb = jb_lo; MODE == 1 && jb <= jb_hi; jb++) {
    float P[kCols];
#pragma unroll
    for (int n = 0; n < kCols; n++) P[n] = 0.f;
    for (int c = 0; c < H; c++) {
      __syncthreads();
      stage_operand(qs, qb, i0, kRows, L, H, F, c, scale);
      run_tile(c, jb);
      if (warp < 4) {
        const float z = lse[c * kRows + r];
#pragma unroll
        for (int h = 0; h < 2; h++) {
          float v[32];
          tmem_ld32(lanebase + 32 * h, v);
#pragma unroll
          for (int n = 0; n < 32; n++) P[32 * h + n] += ex2(v[n] - z);   // z = +inf for a dead row / head: adds 0
        }
      }
    }
    if (warp < 4 && rowlive) {
      float *row = links + ((size_t)b * L + i) * T;
#pragma unroll
      for (int n = 0; n < kCols; n++) {
        const int j = jb * kCols + n;
        if (j > i && j < O && j - i - 1 < T) row[j - i - 1] = P[n] > 0.f ? log2f(P[n]) * kLn2 : neg_inf_f();
      }
    }
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
// written; stats: fp32 [B][H][L] scratch (per-head row normalisers).  F a multiple of 16, 16 <= F <= 128; H <= 64.
extern "C" int dagb200_extract_links(const float *q, const float *key, const float *log_gates, const int64_t *output_length,
                                     float *stats, float *links, int B, int L, int H, int F, int T, void *stream) {
  using namespace lk;
  if (B < 0 || L < 1 || H < 1 || H > 64 || F < 16 || F > 128 || (F & 15) || T < 0 || B > 65535 || !stats) {
    set_error("extract_links: bad shape (F a multiple of 16 in [16, 128], 1 <= H <= 64)");
    return DAGB200_EINVAL;
  }
  if (B == 0 || T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * L * T;
  fill_neg_inf_kernel<<<(int)std::min<size_t>((n + 1023) / 1024, (size_t)8 * sm_count()), 256, 0, st>>>(links, n);
  const int F8 = F >> 3;
  const size_t smem = (size_t)2 * F8 * (kRows + kCols) * 16 + (size_t)H * kRows * 4 + 64;
  cudaFuncSetAttribute(extract_links_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(extract_links_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int rts = (L + kRows - 1) / kRows;
  const int ngroups = ((L + kCols - 1) / kCols + kGroup - 1) / kGroup;
  extract_links_tcgen05_kernel<0><<<dim3(rts, H, B), kThreads, smem, st>>>(q, key, log_gates, output_length, stats, links,
                                                                          L, H, F, T, ngroups);
  DAGB200_CHECK_LAUNCH("extract_links_tcgen05_kernel<0>");
  extract_links_tcgen05_kernel<1><<<dim3(rts * ngroups, B), kThreads, smem, st>>>(q, key, log_gates, output_length, stats,
                                                                                 links, L, H, F, T, ngroups);
  DAGB200_CHECK_LAUNCH("extract_links_tcgen05_kernel<1>");
  return 0;
}
