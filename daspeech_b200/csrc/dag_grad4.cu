// dag_grad4.cu -- backward of the DAG loss for sm_100a (fp32 path, needs a workspace): the transition gradient as a
// tcgen05 contraction with the A operand in tensor memory.
//
// Replaces calculate_grad_match_all_kernel + calculate_grad_links_kernel (reference dag_loss.cu:378-401, 432-485).
//
//   gm[t,j] = exp(alpha + beta - match - Z) * go
//   gl[i,k] = go * exp(links[i,k]) * sum_t exp(alpha[t,i] + beta[t+1,n] - Z),   n = i + k + 1
//
//   pass 1  grad_planes4_kernel  one streaming pass over alpha, beta, match: writes gm and the two operands of the
//             contraction over the target index t, bf16 hi + lo with one integer frame per (row, 32-vertex block):
//               A[t,i] = exp2(alpha[t,i] log2e - FA[t,I]),   B[t,n] = exp2(beta[t+1,n] log2e - FB[t+1,N])
//             already in the layouts the contraction consumes, 16 rows (one K step of the MMA) at a time:
//               A: [chunk][vertex][hi | lo][16 rows]                       -- 64 contiguous bytes per (chunk, vertex)
//               B: [chunk][32-vertex block][hi | lo][k-core][vertex][8 rows] -- the canonical K-major core-matrix layout
//                                                                            of tcgen05 (no swizzle), 2 KB per tile
//   pass 2  grad_fmax_kernel     Fmax[I,N] = max_t (FA[t,I] + FB[t+1,N]): the frame of a block pair's sum
//   pass 3  grad_links_tcgen05_kernel   a CTA owns a 128 x 128 (source x destination) tile = 4 x 4 block pairs, two CTAs
//             per SM (256 TMEM columns each) so that one's epilogue overlaps the other's contraction:
//             D[128 x 128] (TMEM, fp32) += A'[128 x 16] * B[16 x 128], 16 rows per step.  For a block pair every row t
//             has ONE scale 2^(FA[t,I] + FB[t+1,N] - Fmax[I,N]), an exact power of two that depends on the K index, so
//             it cannot live on the accumulator side: four warps (thread = source vertex = TMEM lane, warp = block I)
//             read their rows of A straight from global memory (64 contiguous bytes per thread), multiply them by the
//             four scale columns (packed HMUL2, the scale pairs are computed once per warp) and write the four scaled
//             copies into TENSOR MEMORY (tcgen05.st, double-buffered); the MMAs take A from there
//             (tcgen05.mma ... [d], [a], b-desc) and B from shared memory, where it arrives by one bulk copy per step
//             (3-stage mbarrier ring).  bf16 hi/lo split on both operands: 3 MMAs per product (M128 N32 K16), 12 per
//             step, one issuing thread.  The links tile of the epilogue is staged in shared memory by cp.async while
//             the contraction runs; the row owners multiply their accumulators by exp2(links log2e + Fmax - Z log2e) in
//             place, and one coalesced pass per row writes grad_links once, including the zero padding.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace dagb200 {
namespace g4 {

constexpr int kNegBig = -(1 << 20);
constexpr int kKc = 16;                            // rows per step (K of one MMA)
constexpr int kBI = 128, kBN = 128;                // tile: source vertices x destination vertices (two CTAs per SM:
                                                   // one's epilogue overlaps the other's contraction)
constexpr int kNBt = kBN / 32;                     // destination blocks per tile
constexpr int kThreads = 192;                      // 4 scale/epilogue warps + MMA issuer + bulk-copy producer
constexpr int kStages = 3;
constexpr int kStageBytes = kNBt * 2048;           // B tiles of one step
constexpr int kCsPitch = kBN + 1;                  // float pitch of the staged output tile
constexpr double kL2E_D = 1.4426950408889634074;
constexpr float kL2E = 1.4426950408889634074f;
constexpr float kL2E_LO = (float)(kL2E_D - (double)kL2E);

struct Planes {
  int Lp, NBp, Mc;             // padded row length (multiple of 256), 32-vertex blocks per row, 16-row chunks
  size_t off_a, off_b, off_fa, off_fb, off_fmax, sample_a, sample_b, bytes;
  __host__ __device__ static inline Planes make(int B, int M, int L) {
    Planes p;
    p.Lp = (L + 255) / 256 * 256;
    p.NBp = p.Lp / 32;
    p.Mc = (M + kKc - 1) / kKc;
    p.sample_a = (size_t)p.Mc * p.Lp * 64;            // bytes per utterance
    p.sample_b = (size_t)p.Mc * p.NBp * 2048;
    size_t o = 0;
    p.off_a = o; o += (size_t)B * p.sample_a;
    p.off_b = o; o += (size_t)B * p.sample_b;
    o = (o + 255) & ~(size_t)255;
    p.off_fa = o; o += (size_t)B * (M + 1) * p.NBp * sizeof(int);
    o = (o + 255) & ~(size_t)255;
    p.off_fb = o; o += (size_t)B * (M + 1) * p.NBp * sizeof(int);
    o = (o + 255) & ~(size_t)255;
    p.off_fmax = o; o += (size_t)B * p.NBp * p.NBp * sizeof(int);
    p.bytes = (o + 255) & ~(size_t)255;
    return p;
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// x * log2(e) - F with the product carried to ~2^-45 (the lattice values are ~1e3: a plain fp32 product would add
// 1e-4 of relative error to the exponential)
__device__ __forceinline__ float scaled_log2(float x, float p, int F) {
  const float e1 = fmaf(x, kL2E, -p);             // exact residual of p = x * kL2E
  return (p - (float)F) + fmaf(x, kL2E_LO, e1);
}

// ---------------------------------------------------------------------------------------------------------
// pass 1.  CTA = (utterance, chunk of 16 rows, 128 vertices); warp w owns rows w and w + 8 of the chunk (and warp 0 the
// beta row that follows the chunk: the B operand of chunk c holds beta rows 16c+1 .. 16c+16); lane = 4 consecutive
// vertices, 8 lanes = one 32-vertex block.
constexpr int kPlThreads = 288;                    // 8 warps x 2 rows + one warp for the beta row that follows the chunk
__global__ void __launch_bounds__(kPlThreads)
grad_planes4_kernel(const float *__restrict__ go, const float *__restrict__ alpha, const float *__restrict__ beta,
                    const float *__restrict__ match, float *__restrict__ gm, unsigned char *__restrict__ ws, Planes pl,
                    int M, int L) {
  __shared__ __align__(16) __nv_bfloat16 sa[2][kKc][128 + 8];     // [plane][row][vertex]: written 4 vertices at a time, read by column
  __shared__ __align__(16) __nv_bfloat16 sb[2][kKc][128 + 8];     // [plane][row - 1][vertex]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, c = blockIdx.y, vg = blockIdx.x;
  const int i = vg * 128 + 4 * lane;
  const float ninf = neg_inf_f();
  const float Z = __ldg(beta + (int64_t)b * M * L);
  const float g = __ldg(go + b);
  const bool zinf = isinf(Z);
  const bool vecok = (L % 4 == 0) && (i + 3 < L) &&
                     ((((uintptr_t)alpha | (uintptr_t)beta | (uintptr_t)match | (uintptr_t)gm) & 15) == 0);
  int *FAg = reinterpret_cast<int *>(ws + pl.off_fa) + (size_t)b * (M + 1) * pl.NBp;
  int *FBg = reinterpret_cast<int *>(ws + pl.off_fb) + (size_t)b * (M + 1) * pl.NBp;
  // rows: r = 0..15 both lattices (+ emission gradient), r = 16 beta only
  for (int r = (warp < 8 ? warp : kKc); r <= kKc; r += 8) {
    if ((r == kKc) != (warp == 8)) break;
    const int t = c * kKc + r;
    const bool rowin = t < M;
    const bool full = r < kKc;                         // alpha / match / gm are handled for this row
    const int64_t row = ((int64_t)b * M + t) * L;
    float a[4], be[4], m[4];
#pragma unroll
    for (int e = 0; e < 4; e++) { a[e] = ninf; be[e] = ninf; m[e] = ninf; }
    if (rowin) {
      if (vecok) {
        const float4 b4 = __ldcs(reinterpret_cast<const float4 *>(beta + row + i));
        be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
        if (full) {
          const float4 a4 = __ldcs(reinterpret_cast<const float4 *>(alpha + row + i));
          const float4 m4 = __ldcs(reinterpret_cast<const float4 *>(match + row + i));
          a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
          m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const bool in = i + e < L;
          be[e] = in ? beta[row + i + e] : ninf;
          if (full) { a[e] = in ? alpha[row + i + e] : ninf; m[e] = in ? match[row + i + e] : ninf; }
        }
      }
    }
    if (full && rowin) {
      // emission gradient (reference dag_loss.cu:395-399)
      float rr[4];
#pragma unroll
      for (int e = 0; e < 4; e++) rr[e] = (zinf || isinf(m[e])) ? 0.f : expf(a[e] + be[e] - m[e] - Z) * g;
      if (vecok) {
        __stcs(reinterpret_cast<float4 *>(gm + row + i), make_float4(rr[0], rr[1], rr[2], rr[3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; e++)
          if (i + e < L) gm[row + i + e] = rr[e];
      }
    }
    // operands.  A vertex takes part in the A operand when its alpha is finite and in the B operand when its beta is;
    // vertices without posterior mass (either lattice -inf at their own cell) are dropped from both, so that the frames
    // follow the vertices that matter.  The beta-only row (r = 16) has no alpha at hand: its liveness is beta's alone
    // (its alpha-dead vertices only meet zero products: alpha = -inf upstream).
    float pa[4], pb[4];
    float mxa = ninf, mxb = ninf;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const bool lb = be[e] > ninf && be[e] < 3.0e38f;
      const bool la = a[e] > ninf && a[e] < 3.0e38f;
      pa[e] = (full && la && lb) ? a[e] * kL2E : ninf;
      pb[e] = (lb && (la || !full)) ? be[e] * kL2E : ninf;
      mxa = fmaxf(mxa, pa[e]);
      mxb = fmaxf(mxb, pb[e]);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      mxa = fmaxf(mxa, __shfl_xor_sync(0xffffffffu, mxa, o));
      mxb = fmaxf(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
    }
    const int FA = mxa > ninf ? (int)ceilf(mxa) : kNegBig;
    const int FB = mxb > ninf ? (int)ceilf(mxb) : kNegBig;
    {
      __nv_bfloat16 ah[4], al[4], bh[4], bl[4];
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const float va = pa[e] > ninf ? exp2f(scaled_log2(a[e], pa[e], FA)) : 0.f;
        const float vb = pb[e] > ninf ? exp2f(scaled_log2(be[e], pb[e], FB)) : 0.f;
        ah[e] = __float2bfloat16_rn(va); al[e] = __float2bfloat16_rn(va - __bfloat162float(ah[e]));
        bh[e] = __float2bfloat16_rn(vb); bl[e] = __float2bfloat16_rn(vb - __bfloat162float(bh[e]));
      }
      auto pack4 = [](const __nv_bfloat16 (&x)[4]) {
        const __nv_bfloat162 p0 = __halves2bfloat162(x[0], x[1]), p1 = __halves2bfloat162(x[2], x[3]);
        return make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
      };
      if (full) {
        *reinterpret_cast<uint2 *>(&sa[0][r][4 * lane]) = pack4(ah);
        *reinterpret_cast<uint2 *>(&sa[1][r][4 * lane]) = pack4(al);
      }
      if (r >= 1) {                                      // beta row t is row t - 1 of the B operand
        *reinterpret_cast<uint2 *>(&sb[0][r - 1][4 * lane]) = pack4(bh);
        *reinterpret_cast<uint2 *>(&sb[1][r - 1][4 * lane]) = pack4(bl);
      }
    }
    if ((lane & 7) == 0 && t <= M) {
      const size_t fo = (size_t)t * pl.NBp + vg * 4 + (lane >> 3);
      if (full) FAg[fo] = FA;
      // a beta row is seen by two CTAs (as the row after a chunk and as the first row of the next one): its frame is
      // written by the one that also uses the alpha liveness -- both liveness rules give frames that cover the row, but
      // the operand and the frame must come from the same rule, and the operand of row t lives in the chunk that holds
      // it as row t - 1
      if (r >= 1) FBg[fo] = FB;
    }
  }
  __syncthreads();
  // A: [chunk][vertex][plane][16 rows]: 32 bytes per (vertex, plane)
  if (threadIdx.x < 256) {
    unsigned char *Ab = ws + pl.off_a + (size_t)b * pl.sample_a + ((size_t)c * pl.Lp + vg * 128) * 64;
    const int v = threadIdx.x >> 1, plane = threadIdx.x & 1;
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const __nv_bfloat162 pr = __halves2bfloat162(sa[plane][2 * q][v], sa[plane][2 * q + 1][v]);
      w[q] = *reinterpret_cast<const uint32_t *>(&pr);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(Ab + (size_t)v * 64 + plane * 32);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  // B: [chunk][block][plane][k-core][vertex][8 rows]: 16 bytes per (block, plane, k-core, vertex)
  {
    unsigned char *Bb = ws + pl.off_b + (size_t)b * pl.sample_b + ((size_t)c * pl.NBp + vg * 4) * 2048;
    for (int x = threadIdx.x; x < 512; x += kPlThreads) {
      const int blk = x >> 7, plane = (x >> 6) & 1, kc = (x >> 5) & 1, n = x & 31;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const __nv_bfloat162 pr = __halves2bfloat162(sb[plane][8 * kc + 2 * q][blk * 32 + n], sb[plane][8 * kc + 2 * q + 1][blk * 32 + n]);
        w[q] = *reinterpret_cast<const uint32_t *>(&pr);
      }
      *reinterpret_cast<uint4 *>(Bb + (size_t)blk * 2048 + ((plane * 2 + kc) * 32 + n) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// pass 2: Fmax[I][N] = max_t (FA[t][I] + FB[t+1][N]) over the rows that carry mass on both sides.  One CTA per
// (utterance, source block): the frames travel through shared memory 64 rows at a time, thread = destination block.
constexpr int kFmRows = 64;
__global__ void __launch_bounds__(128)
grad_fmax_kernel(const int64_t *__restrict__ tlen, unsigned char *__restrict__ ws, Planes pl, int M) {
  extern __shared__ int fm_smem[];                     // [kFmRows] FA of my block | [kFmRows][NBp] FB
  int *sfa = fm_smem, *sfb = fm_smem + kFmRows;
  const int b = blockIdx.y, I = blockIdx.x;
  const int Tn = (int)tlen[b];
  const int nsteps = (Tn >= 2 && Tn <= M) ? Tn - 1 : 0;
  const int *FA = reinterpret_cast<const int *>(ws + pl.off_fa) + (size_t)b * (M + 1) * pl.NBp;
  const int *FB = reinterpret_cast<const int *>(ws + pl.off_fb) + (size_t)b * (M + 1) * pl.NBp;
  int *out = reinterpret_cast<int *>(ws + pl.off_fmax) + ((size_t)b * pl.NBp + I) * pl.NBp;
  for (int N0 = 0; N0 < pl.NBp; N0 += 128) {
    const int N = N0 + threadIdx.x;
    int fmx = kNegBig;
    for (int t0 = 0; t0 < nsteps; t0 += kFmRows) {
      const int nr = min(kFmRows, nsteps - t0);
      __syncthreads();
      for (int x = threadIdx.x; x < nr; x += 128) sfa[x] = FA[(size_t)(t0 + x) * pl.NBp + I];
      for (int x = threadIdx.x; x < nr * pl.NBp; x += 128) sfb[x] = FB[(size_t)(t0 + 1) * pl.NBp + x];
      __syncthreads();
      if (N < pl.NBp && N >= I) {
#pragma unroll 4
        for (int r = 0; r < nr; r++) {
          const int fa = sfa[r], fb = sfb[r * pl.NBp + N];
          if (fa > kNegBig && fb > kNegBig) fmx = max(fmx, fa + fb);
        }
      }
    }
    if (N < pl.NBp) out[N] = fmx;
  }
}

// ---- tcgen05 / mbarrier helpers (encodings verified stand-alone in tools/tcgen05_probe.cu and tcgen05_ts_probe.cu) --
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
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
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t v, uint32_t sc) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162 *>(&v), *reinterpret_cast<const __nv_bfloat162 *>(&sc));
  return *reinterpret_cast<uint32_t *>(&r);
}
// 16 consecutive 32-bit columns of my TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&w)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]),
        "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
}
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

#ifdef DAGB200_G4_TIMING
__device__ long long g_g4t[16];
#define G4T(i, v) do { if (tlog) atomicAdd((unsigned long long *)&g_g4t[i], (unsigned long long)(v)); } while (0)
#else
#define G4T(i, v) do { } while (0)
#endif

// TMEM (256 columns per CTA): accumulator 0 .. 127 (destination vertex), scaled A copies 128 + 64 * buffer + 16 * block + 8 * plane + (0..7)
constexpr uint32_t kTmemA = kBN;
constexpr uint32_t kTmemCols = 256;

__global__ void __launch_bounds__(kThreads, 2)
grad_links_tcgen05_kernel(const float *__restrict__ go, const float *__restrict__ beta, const float *__restrict__ links,
                          const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen, float *__restrict__ gl,
                          const unsigned char *__restrict__ ws, Planes pl, int M, int L, int Tl, int NNt) {
  extern __shared__ __align__(128) unsigned char g4_smem[];
  const int b = blockIdx.y;
  // row tiles in descending order: the last ones carry the most zero padding (transitions beyond the graph) and start first
  const int It = (int)(gridDim.x / NNt) - 1 - blockIdx.x / NNt, Nt = blockIdx.x % NNt;
  const int i0 = It * kBI, n0 = Nt * kBN;
  if (n0 + kBN - 1 <= i0) return;                         // no destination after a source: nothing stored here
  if (n0 - (i0 + kBI - 1) - 1 >= Tl) return;              // entirely beyond the transition band: no storage
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef DAGB200_G4_TIMING
  const bool tlog = b == 0 && blockIdx.x == gridDim.x - NNt + 1 && lane == 0;
  const long long tk0 = clock64();
#endif
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const float *E = links + (int64_t)b * L * Tl;
  float *g = gl + (int64_t)b * L * Tl;
  const float Z = beta[(int64_t)b * M * L];
  const float gout = go[b];
  const bool dead = isinf(Z) || O > L || Tn > M || Tn < 2 || O < 2;
  const int nsteps = dead ? 0 : Tn - 1;
  const bool compute = nsteps > 0 && i0 < O && n0 < O;
  const int nchunks = compute ? (nsteps + kKc - 1) / kKc : 0;

  unsigned char *ring = g4_smem;                                                        // [kStages][kStageBytes]
  float *cs = reinterpret_cast<float *>(g4_smem + kStages * kStageBytes);               // [128][kCsPitch]
  uint32_t *sct = reinterpret_cast<uint32_t *>(cs + kBI * kCsPitch);                    // [4 warps][blocks][8 row pairs] scale pairs
  float *dpair = reinterpret_cast<float *>(sct + 4 * 8 * kNBt);                         // [4][blocks] epilogue exponents
  uint64_t *bfull = reinterpret_cast<uint64_t *>(dpair + 4 * kNBt);                     // [kStages]
  uint64_t *bempty = bfull + kStages;                                                   // [kStages]
  uint64_t *afull = bempty + kStages;                                                   // [2]
  uint64_t *aempty = afull + 2;                                                         // [2]
  uint64_t *dfull = aempty + 2;                                                         // [1]
  uint32_t *tmem_s = reinterpret_cast<uint32_t *>(dfull + 1);

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(bfull + s, 1); mbar_init(bempty + s, 1); }
    for (int s = 0; s < 2; s++) { mbar_init(afull + s, 4); mbar_init(aempty + s, 1); }
    mbar_init(dfull, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_s)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#ifdef DAGB200_G4_TIMING
  if (warp == 0) G4T(11, clock64() - tk0);
#endif
  const uint32_t tmem = *tmem_s;
  const int *FA = reinterpret_cast<const int *>(ws + pl.off_fa) + (size_t)b * (M + 1) * pl.NBp;
  const int *FB = reinterpret_cast<const int *>(ws + pl.off_fb) + (size_t)b * (M + 1) * pl.NBp;
  const int *FM = reinterpret_cast<const int *>(ws + pl.off_fmax) + (size_t)b * pl.NBp * pl.NBp;

  if (compute) {
    if (warp == 5) {
      // ================================ bulk-copy producer (one thread): the eight B tiles of a step ==============
      if (lane == 0) {
        const unsigned char *Bb = ws + pl.off_b + (size_t)b * pl.sample_b + (size_t)(n0 / 32) * 2048;
        for (int c = 0; c < nchunks; c++) {
          const int st = c % kStages;
          if (c >= kStages) mbar_wait(bempty + st, ((c / kStages) - 1) & 1);
          mbar_expect_tx(bfull + st, kStageBytes);
          bulk_g2s(ring + (size_t)st * kStageBytes, Bb + (size_t)c * pl.NBp * 2048, kStageBytes, bfull + st);
        }
      }
    } else if (warp == 4) {
      // ================================ MMA issuer (one elected lane) =============================================
      const uint32_t ring_u = smem_u32(ring);
      for (int c = 0; c < nchunks; c++) {
        const int st = c % kStages, ab = c & 1;
#ifdef DAGB200_G4_TIMING
        const long long ti0 = clock64();
#endif
        mbar_wait(bfull + st, (c / kStages) & 1);
#ifdef DAGB200_G4_TIMING
        const long long ti1 = clock64();
#endif
        mbar_wait(afull + ab, (c >> 1) & 1);
#ifdef DAGB200_G4_TIMING
        const long long ti2 = clock64();
        G4T(0, ti1 - ti0); G4T(1, ti2 - ti1);
#endif
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ta = tmem + kTmemA + 16 * kNBt * ab;
#pragma unroll
          for (int nb = 0; nb < kNBt; nb++) {
            const uint32_t bt = ring_u + st * kStageBytes + nb * 2048;             // [plane][k-core][32][16 B]
            const uint64_t bhi = umma_desc(bt, 512, 128), blo = umma_desc(bt + 1024, 512, 128);
            const uint32_t d = tmem + nb * 32;
            umma_ts(d, ta + 16 * nb, bhi, c > 0 ? 1u : 0u);                          // hi * hi
            umma_ts(d, ta + 16 * nb + 8, bhi, 1u);                                   // lo * hi
            umma_ts(d, ta + 16 * nb, blo, 1u);                                       // hi * lo
          }
          umma_commit(bempty + st);
          umma_commit(aempty + ab);
          if (c == nchunks - 1) umma_commit(dfull);
        }
        __syncwarp();
#ifdef DAGB200_G4_TIMING
        G4T(2, clock64() - ti2);
#endif
      }
    } else {
      // ================================ scale warps: thread = source vertex, warp = source block ====================
      const int I = i0 / 32 + warp;                    // my block
      const int i = i0 + tid;                          // my vertex
      const unsigned char *Ab = ws + pl.off_a + (size_t)b * pl.sample_a + (size_t)i * 64;
      uint32_t *sc = sct + warp * 8 * kNBt;
      // epilogue exponent of my block's eight pairs and the scale bases
      int fmx = kNegBig;
      if (lane < kNBt) {
        const int N = n0 / 32 + lane;
        fmx = (N < pl.NBp && I < pl.NBp) ? __ldg(FM + (size_t)I * pl.NBp + N) : kNegBig;
        dpair[warp * kNBt + lane] = fmx > kNegBig ? (float)((double)fmx - (double)Z * kL2E_D) : 0.f;
      }
      // lane -> (destination block nbl, row pair rp) of the scale table
      const int nbl = lane >> 3, rp = lane & 7;
      const int fmx_l = __shfl_sync(0xffffffffu, fmx, nbl);
      const int Nl = n0 / 32 + nbl;
      // the links tile of the epilogue travels to shared memory (the staging tile of the output) while the contraction
      // runs: 4-byte asynchronous copies (the rows are not 16-byte aligned), lanes = consecutive transitions
      {
        const int n = n0 + tid;                              // my destination vertex (column of the tile)
        const int ii_hi = min(min(kBI, O - i0), n - i0);     // rows with ir < O and k = n - ir - 1 >= 0
        const int ii_lo = max(0, n - Tl - i0);               // ... and k < Tl
        const bool col_on = n < O;
        float *dst = cs + tid;
        const float *src = E + (int64_t)i0 * Tl + (n - i0 - 1);          // row ii: src + ii * (Tl - 1)
        const uint32_t dst_u = smem_u32(dst);
#pragma unroll 4
        for (int ii = 0; ii < kBI; ii++) {
          if (col_on && ii >= ii_lo && ii < ii_hi) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_u + (uint32_t)(ii * kCsPitch * 4)), "l"(src + (int64_t)ii * (Tl - 1)) : "memory");
          } else {
            dst[ii * kCsPitch] = neg_inf_f();
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      uint4 raw[4];
      {
        const uint4 *src = reinterpret_cast<const uint4 *>(Ab);
#pragma unroll
        for (int q = 0; q < 4; q++) raw[q] = __ldg(src + q);
      }
      // scale pairs of a step: 2^(FA[t,I] + FB[t+1,N] - Fmax[I,N]) as bf16 (exact); 0 where a side has no mass or the
      // row is more than 2^-126 below the pair's best row.  The frames are loaded one step ahead.
      int fan[2], fbn[2];
      auto load_frames = [&](int c) {
#pragma unroll
        for (int x = 0; x < 2; x++) {
          const int t = c * kKc + 2 * rp + x;
          const bool ok = t < nsteps && Nl < pl.NBp;
          fan[x] = ok ? __ldg(FA + (size_t)t * pl.NBp + I) : kNegBig;
          fbn[x] = ok ? __ldg(FB + (size_t)(t + 1) * pl.NBp + Nl) : kNegBig;
        }
      };
      load_frames(0);
#ifdef DAGB200_G4_TIMING
      if (warp == 0) G4T(10, clock64() - tk0);
#endif
      for (int c = 0; c < nchunks; c++) {
        const int ab = c & 1;
#ifdef DAGB200_G4_TIMING
        const long long ts0 = clock64();
#endif
        uint32_t pk = 0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int fa = fan[e], fb = fbn[e];
          const int d = (fa > kNegBig && fb > kNegBig && fmx_l > kNegBig) ? fa + fb - fmx_l : kNegBig;
          const uint32_t bits = d >= -126 ? (uint32_t)(d + 127) << 7 : 0u;
          pk |= bits << (16 * e);
        }
        __syncwarp();                                   // the previous step's readers are done with the table
        sc[nbl * 8 + rp] = pk;
        __syncwarp();
        if (c + 1 < nchunks) load_frames(c + 1);
        // my 16 rows of A for this step (hi: raw[0..1], lo: raw[2..3]); the next step's are requested now
        uint32_t av[16];
        av[0] = raw[0].x; av[1] = raw[0].y; av[2] = raw[0].z; av[3] = raw[0].w;
        av[4] = raw[1].x; av[5] = raw[1].y; av[6] = raw[1].z; av[7] = raw[1].w;
        av[8] = raw[2].x; av[9] = raw[2].y; av[10] = raw[2].z; av[11] = raw[2].w;
        av[12] = raw[3].x; av[13] = raw[3].y; av[14] = raw[3].z; av[15] = raw[3].w;
        if (c + 1 < nchunks) {
          const uint4 *src = reinterpret_cast<const uint4 *>(Ab + (size_t)(c + 1) * pl.Lp * 64);
#pragma unroll
          for (int q = 0; q < 4; q++) raw[q] = __ldg(src + q);
        }
        // the MMAs that read this buffer two steps ago are complete
#ifdef DAGB200_G4_TIMING
        const long long ts1 = clock64();
#endif
        if (c >= 2) mbar_wait(aempty + ab, ((c >> 1) - 1) & 1);
#ifdef DAGB200_G4_TIMING
        const long long ts2 = clock64();
        if (warp == 0) { G4T(3, ts1 - ts0); G4T(4, ts2 - ts1); }
#endif
        tc_fence_after();
        const uint32_t ta = tmem + kTmemA + 16 * kNBt * ab + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int nb = 0; nb < kNBt; nb++) {
          const uint4 s0 = *reinterpret_cast<const uint4 *>(sc + nb * 8), s1 = *reinterpret_cast<const uint4 *>(sc + nb * 8 + 4);
          const uint32_t sp[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          uint32_t w[16];
#pragma unroll
          for (int q = 0; q < 8; q++) { w[q] = hmul2_u32(av[q], sp[q]); w[8 + q] = hmul2_u32(av[8 + q], sp[q]); }
          tmem_st16(ta + 16 * nb, w);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull + ab);
#ifdef DAGB200_G4_TIMING
        if (warp == 0) G4T(5, clock64() - ts2);
#endif
      }
#ifdef DAGB200_G4_TIMING
      const long long tm0 = clock64();
      if (warp == 0) G4T(6, tm0 - tk0);
#endif
      // ---- accumulators (thread = row) x exp2(links log2e + Fmax - Z log2e) x go, in place over the staged links tile
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");          // every scale thread's share of the links tile has landed
      mbar_wait(dfull, 0);
      tc_fence_after();
#pragma unroll 1
      for (int nb = 0; nb < kNBt; nb++) {
        float v[32];
        tmem_ld32(tmem + nb * 32 + ((uint32_t)(warp * 32) << 16), v);
        const float dp = dpair[warp * kNBt + nb];
        float *row = cs + tid * kCsPitch + nb * 32;
#pragma unroll
        for (int j = 0; j < 32; j++) row[j] = gout * exp2f(fmaf(row[j], kL2E, dp)) * v[j];
      }
      tc_fence_before();
#ifdef DAGB200_G4_TIMING
      if (warp == 0) G4T(7, clock64() - tm0);
#endif
    }
  }
  __syncthreads();
#ifdef DAGB200_G4_TIMING
  const long long te0 = clock64();
#endif

  // ---- epilogue: gl = go * exp2(links * log2e + Fmax - Z log2e) * G, one coalesced write per row, zeros elsewhere.
  const bool last_col = (n0 + kBN >= L);
  for (int ii = warp; ii < kBI; ii += kThreads / 32) {
    const int i = i0 + ii;
    if (i >= L) break;
    const bool rowon = compute && i < O;
    const int kb = n0 + lane - i - 1;                        // transition index of my first column
    float *gp = g + (int64_t)i * Tl + kb;
    const float *cp = cs + ii * kCsPitch + lane;
#pragma unroll
    for (int h = 0; h < kNBt; h++) {
      const int k = kb + 32 * h;
      if (k >= 0 && k < Tl) gp[32 * h] = (rowon && n0 + lane + 32 * h < O) ? cp[32 * h] : 0.f;
    }
    if (last_col) {  // transitions that point beyond the padded graph: k >= L-1-i
      float *grow = g + (int64_t)i * Tl;
      for (int k = max(0, n0 + kBN - i - 1) + lane; k < Tl; k += 32) grow[k] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef DAGB200_G4_TIMING
  if (warp == 0) { G4T(8, clock64() - te0); G4T(9, clock64() - tk0); }
#endif
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

size_t links_smem_bytes() {
  return (size_t)kStages * kStageBytes + sizeof(float) * kBI * kCsPitch + 4 * 8 * kNBt * 4 + 4 * kNBt * 4 + (2 * kStages + 5) * 8 + 64;
}

}  // namespace g4

size_t grad4_workspace_bytes(int B, int M, int L) { return g4::Planes::make(B, M, L).bytes; }
bool grad4_supported(int M, int L) { return M >= 2 && L >= 1; }

int launch_grad4(const float *go, const float *alpha, const float *beta, const float *match, const float *links,
                 const int64_t *olen, const int64_t *tlen, float *gm, float *gl, int B, int M, int L, int Tl,
                 void *workspace, cudaStream_t st) {
  using namespace g4;
  const Planes pl = Planes::make(B, M, L);
  prof_mark(3, st);
  {
    dim3 grid(pl.Lp / 128, pl.Mc, B);
    grad_planes4_kernel<<<grid, kPlThreads, 0, st>>>(go, alpha, beta, match, gm, (unsigned char *)workspace, pl, M, L);
    DAGB200_CHECK_LAUNCH("grad_planes4_kernel");
  }
  prof_mark(4, st);
  {
    dim3 grid(pl.NBp, B);
    const size_t smem = (size_t)kFmRows * (1 + pl.NBp) * sizeof(int);
    if (smem > 48 * 1024) cudaFuncSetAttribute(grad_fmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    grad_fmax_kernel<<<grid, 128, smem, st>>>(tlen, (unsigned char *)workspace, pl, M);
    DAGB200_CHECK_LAUNCH("grad_fmax_kernel");
  }
  {
    const int NI = (L + kBI - 1) / kBI, NN = (L + kBN - 1) / kBN;
    const size_t smem = links_smem_bytes();
    cudaFuncSetAttribute(grad_links_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(NI * NN, B);
    grad_links_tcgen05_kernel<<<grid, kThreads, smem, st>>>(go, beta, links, olen, tlen, gl, (const unsigned char *)workspace,
                                                           pl, M, L, Tl, NN);
    DAGB200_CHECK_LAUNCH("grad_links_tcgen05_kernel");
#ifdef DAGB200_G4_TIMING
    {
      static int calls = 0;
      if (++calls == 3) {
        cudaStreamSynchronize(st);
        long long h[16];
        cudaMemcpyFromSymbol(h, g_g4t, sizeof(h));
        printf("[g4 timing, tile 1 of utterance 0, sums over 3 launches] issuer: wait B %lld, wait A %lld, issue %lld | scale warp 0: table+loads %lld, wait buffer %lld, scale+st %lld | main loop %lld, acc->smem %lld, epilogue %lld, total %lld | after alloc+sync %lld, at loop start %lld\n",
               h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[11], h[10]);
      }
    }
#endif
  }
  prof_mark(5, st);
  return 0;
}

}  // namespace dagb200
