// dag_grad3.cu -- backward of the DAG loss in two passes, for sm_100a (fp32 path, needs a workspace).
//
// Replaces calculate_grad_match_all_kernel + calculate_grad_links_kernel (reference dag_loss.cu:378-401, 432-485).
//
//   pass 1  grad_planes_kernel   one streaming pass over alpha, beta, match (float4, coalesced):
//             gm[t,j]  = exp(alpha + beta - match - Z) * go              (the emission gradient, written once)
//             A[t,i]   = exp2(alpha[t,i] * log2e - FA[t,I])  bf16 hi + lo, FA = integer frame of the 32-vertex block I
//             B[t,n]   = exp2(beta[t,n]  * log2e - FB[t,N])  bf16 hi + lo, FB likewise
//             (vertices with alpha = -inf or beta = -inf carry no posterior mass: their entries are 0 and they
//              do not take part in the frames, so a frame follows the vertices that matter)
//   pass 2  grad_links_planes_kernel   gl[i,k] = go * exp(links[i,k]) * sum_t exp(alpha[t,i] + beta[t+1,n] - Z), n = i+k+1,
//             as a tensor-core contraction over the target index t: a CTA owns a 128 x 64 (source x destination)
//             tile, a warp one 32 x 32 pair of vertex blocks (I, N).  For that pair every row t has ONE scale
//             2^(FA[t,I] + FB[t+1,N] - Fmax[I,N]) -- an exact power of two applied to the A fragments (HMUL2) --
//             so the operands are generated once per lattice cell instead of once per tile (the previous kernel
//             spent 1 exp per 64 MACs in every one of the 36 tiles of an utterance and needed two exponent levels).
//             bf16 hi/lo split on both operands: 3 mma.sync.m16n8k16 per product (~2^-16), fp32 accumulation;
//             operand chunks of 16 rows arrive by cp.async (3 stages); the epilogue multiplies by
//             exp2(links * log2e + Fmax - Z * log2e) while streaming the links tile and writes grad_links once,
//             coalesced, including the zero padding.
#include "common.cuh"

namespace dagb200 {
namespace g3 {

constexpr int kNegBig = -(1 << 20);
constexpr int kBI = 128, kBN = 64, kKc = 16, kThreads = 256, kStages = 3;
constexpr int kPA = kBI + 8, kPB = kBN + 8;       // bf16 row pitch (odd multiple of 16 bytes: conflict-free ldmatrix)
constexpr int kCP = kBN + 4;                      // float pitch of the staged output tile
constexpr double kL2E_D = 1.4426950408889634074;
constexpr float kL2E = 1.4426950408889634074f;
constexpr float kL2E_LO = (float)(kL2E_D - (double)kL2E);

struct Planes {
  int Lp, NBp;                 // padded row length (multiple of 128), 32-vertex blocks per row
  size_t plane_elems;          // B * M * Lp
  size_t off_ahi, off_alo, off_bhi, off_blo, off_fa, off_fb, bytes;
  __host__ __device__ static inline Planes make(int B, int M, int L) {
    Planes p;
    p.Lp = (L + 127) / 128 * 128;
    p.NBp = p.Lp / 32;
    p.plane_elems = (size_t)B * M * p.Lp;
    size_t o = 0;
    p.off_ahi = o; o += p.plane_elems * 2;
    p.off_alo = o; o += p.plane_elems * 2;
    p.off_bhi = o; o += p.plane_elems * 2;
    p.off_blo = o; o += p.plane_elems * 2;
    o = (o + 255) & ~(size_t)255;
    p.off_fa = o; o += (size_t)B * M * p.NBp * sizeof(int);
    o = (o + 255) & ~(size_t)255;
    p.off_fb = o; o += (size_t)B * M * p.NBp * sizeof(int);
    p.bytes = (o + 255) & ~(size_t)255;
    return p;
  }
};

__device__ __forceinline__ void split4(const float (&v)[4], uint2 &hi, uint2 &lo) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  __nv_bfloat162 l0 = __floats2bfloat162_rn(v[0] - f0.x, v[1] - f0.y), l1 = __floats2bfloat162_rn(v[2] - f1.x, v[3] - f1.y);
  hi.x = *reinterpret_cast<uint32_t *>(&h0); hi.y = *reinterpret_cast<uint32_t *>(&h1);
  lo.x = *reinterpret_cast<uint32_t *>(&l0); lo.y = *reinterpret_cast<uint32_t *>(&l1);
}

// x * log2(e) - F with the product carried to ~2^-45 (the lattice values are ~1e3: a plain fp32 product would add
// 1e-4 of relative error to the exponential)
__device__ __forceinline__ float scaled_log2(float x, float p, int F) {
  const float e1 = fmaf(x, kL2E, -p);             // exact residual of p = x * kL2E
  return (p - (float)F) + fmaf(x, kL2E_LO, e1);
}

// one warp per (utterance, row, 128 vertices); lane = 4 consecutive vertices, 8 lanes = one 32-vertex block
__global__ void __launch_bounds__(256)
grad_planes_kernel(const float *__restrict__ go, const float *__restrict__ alpha, const float *__restrict__ beta,
                   const float *__restrict__ match, float *__restrict__ gm, unsigned char *__restrict__ ws, Planes pl,
                   int M, int L) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, t = blockIdx.y;
  const int blk = blockIdx.x * 8 + warp;            // 128-vertex block
  if (blk * 128 >= pl.Lp) return;
  const int i = blk * 128 + 4 * lane;
  const int64_t row = ((int64_t)b * M + t) * L;
  const float ninf = neg_inf_f();
  const float Z = __ldg(beta + (int64_t)b * M * L);
  const float g = __ldg(go + b);
  const bool zinf = isinf(Z);
  float a[4], be[4], m[4];
  const bool vec = (L % 4 == 0) && (i + 3 < L) &&
                   ((((uintptr_t)alpha | (uintptr_t)beta | (uintptr_t)match | (uintptr_t)gm) & 15) == 0);
  if (vec) {
    const float4 a4 = __ldcs(reinterpret_cast<const float4 *>(alpha + row + i));
    const float4 b4 = __ldcs(reinterpret_cast<const float4 *>(beta + row + i));
    const float4 m4 = __ldcs(reinterpret_cast<const float4 *>(match + row + i));
    a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
    be[0] = b4.x; be[1] = b4.y; be[2] = b4.z; be[3] = b4.w;
    m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const bool in = i + e < L;
      a[e] = in ? alpha[row + i + e] : ninf;
      be[e] = in ? beta[row + i + e] : ninf;
      m[e] = in ? match[row + i + e] : ninf;
    }
  }
  // emission gradient (reference dag_loss.cu:395-399)
  float r[4];
#pragma unroll
  for (int e = 0; e < 4; e++) r[e] = (zinf || isinf(m[e])) ? 0.f : expf(a[e] + be[e] - m[e] - Z) * g;
  if (vec) {
    __stcs(reinterpret_cast<float4 *>(gm + row + i), make_float4(r[0], r[1], r[2], r[3]));
  } else {
#pragma unroll
    for (int e = 0; e < 4; e++)
      if (i + e < L) gm[row + i + e] = r[e];
  }
  // operand planes
  float pa[4], pb[4];
  float mxa = ninf, mxb = ninf;
#pragma unroll
  for (int e = 0; e < 4; e++) {
    const bool live = a[e] > ninf && be[e] > ninf && a[e] < 3.0e38f && be[e] < 3.0e38f;
    pa[e] = live ? a[e] * kL2E : ninf;
    pb[e] = live ? be[e] * kL2E : ninf;
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
  float va[4], vb[4];
#pragma unroll
  for (int e = 0; e < 4; e++) {
    va[e] = pa[e] > ninf ? exp2f(scaled_log2(a[e], pa[e], FA)) : 0.f;
    vb[e] = pb[e] > ninf ? exp2f(scaled_log2(be[e], pb[e], FB)) : 0.f;
  }
  uint2 hi, lo;
  const size_t off = ((size_t)b * M + t) * pl.Lp + i;
  split4(va, hi, lo);
  *reinterpret_cast<uint2 *>(ws + pl.off_ahi + off * 2) = hi;
  *reinterpret_cast<uint2 *>(ws + pl.off_alo + off * 2) = lo;
  split4(vb, hi, lo);
  *reinterpret_cast<uint2 *>(ws + pl.off_bhi + off * 2) = hi;
  *reinterpret_cast<uint2 *>(ws + pl.off_blo + off * 2) = lo;
  if ((lane & 7) == 0) {
    const size_t fo = ((size_t)b * M + t) * pl.NBp + blk * 4 + (lane >> 3);
    reinterpret_cast<int *>(ws + pl.off_fa)[fo] = FA;
    reinterpret_cast<int *>(ws + pl.off_fb)[fo] = FB;
  }
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t v, uint32_t sc) {
  __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162 *>(&v), *reinterpret_cast<const __nv_bfloat162 *>(&sc));
  return *reinterpret_cast<uint32_t *>(&r);
}

struct Stage {
  __nv_bfloat16 ahi[kKc][kPA], alo[kKc][kPA], bhi[kKc][kPB], blo[kKc][kPB];
};

__global__ void __launch_bounds__(kThreads, 3)
grad_links_planes_kernel(const float *__restrict__ go, const float *__restrict__ beta, const float *__restrict__ links,
                         const int64_t *__restrict__ olen, const int64_t *__restrict__ tlen, float *__restrict__ gl,
                         const unsigned char *__restrict__ ws, Planes pl, int M, int L, int Tl, int NN, int Mp) {
  extern __shared__ __align__(16) unsigned char g3_smem[];
  const int b = blockIdx.y;
  const int I = blockIdx.x / NN, N = blockIdx.x % NN;
  const int i0 = I * kBI, n0 = N * kBN;
  if (n0 + kBN - 1 <= i0) return;                         // no destination after a source: nothing stored here
  if (n0 - (i0 + kBI - 1) - 1 >= Tl) return;              // entirely beyond the transition band: no storage
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const float *E = links + (int64_t)b * L * Tl;
  float *g = gl + (int64_t)b * L * Tl;
  const float Z = beta[(int64_t)b * M * L];
  const float gout = go[b];
  const bool dead = isinf(Z) || O > L || Tn > M || Tn < 2 || O < 2;
  const int nsteps = dead ? 0 : Tn - 1;

  // shared memory: operand stages | per-warp scale tables [8][Mp] bf16 | per-pair epilogue exponents
  Stage *stg = reinterpret_cast<Stage *>(g3_smem);
  int *fsum = reinterpret_cast<int *>(g3_smem + (kStages - 1) * sizeof(Stage));                 // [8][Mp], aliases the last
                                                                                                // stage (free in the prologue)
  unsigned char *stab = g3_smem + kStages * sizeof(Stage);                                      // [8][Mp] biased exponents
  float *lks = reinterpret_cast<float *>(stab + 8 * Mp);                                        // [128][64] links tile
  float *dpair = lks + kBI * kBN;                                                               // [8]
  if ((size_t)8 * Mp * sizeof(int) > sizeof(Stage)) fsum = reinterpret_cast<int *>(dpair + 8);  // long targets: own region
  float *cs = reinterpret_cast<float *>(g3_smem);                                  // [128][kCP] after the K loop

  const bool compute = nsteps > 0 && i0 < O && n0 < O;
  const int wi = warp >> 1, wn = warp & 1;               // my pair of vertex blocks inside the tile
  float acc[2][4][4];
#pragma unroll
  for (int x = 0; x < 2; x++)
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[x][y][e] = 0.f;

  if (compute) {
    const __nv_bfloat16 *Ahi = reinterpret_cast<const __nv_bfloat16 *>(ws + pl.off_ahi) + (size_t)b * M * pl.Lp;
    const __nv_bfloat16 *Alo = reinterpret_cast<const __nv_bfloat16 *>(ws + pl.off_alo) + (size_t)b * M * pl.Lp;
    const __nv_bfloat16 *Bhi = reinterpret_cast<const __nv_bfloat16 *>(ws + pl.off_bhi) + (size_t)b * M * pl.Lp;
    const __nv_bfloat16 *Blo = reinterpret_cast<const __nv_bfloat16 *>(ws + pl.off_blo) + (size_t)b * M * pl.Lp;
    const int *FA = reinterpret_cast<const int *>(ws + pl.off_fa) + (size_t)b * M * pl.NBp;
    const int *FB = reinterpret_cast<const int *>(ws + pl.off_fb) + (size_t)b * M * pl.NBp;
    const int nchunks = (nsteps + kKc - 1) / kKc;
    // my links tile (read by the epilogue) starts travelling to shared memory now: 128 rows x 64 transitions, 4-byte
    // cp.async granules (the rows are not 16-byte aligned), first commit group
    for (int x = tid; x < kBI * kBN; x += kThreads) {
      const int ii = x >> 6, nn = x & 63;
      const int i = i0 + ii, n = n0 + nn, k = n - i - 1;
      if (i < O && n < O && k >= 0 && k < Tl) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(lks + x);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(E + (int64_t)i * Tl + k) : "memory");
      } else {
        lks[x] = neg_inf_f();
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // one chunk of 16 target rows: A rows t, B rows t + 1, 768 16-byte cp.async granules = 3 per thread with fixed
    // (row, plane, column) roles, so the addressing is done once
    const __nv_bfloat16 *gsrc[3];
    uint32_t sdst[3];
    int grow_[3];                      // global row offset of the granule relative to the chunk's first row
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const int x = tid + r * kThreads;
      const int tr = x / 48, w = x % 48;
      Stage &S0 = stg[0];
      if (w < 32) {
        const int plane = w >> 4, gq = w & 15;
        gsrc[r] = (plane ? Alo : Ahi) + (size_t)tr * pl.Lp + i0 + gq * 8;
        sdst[r] = (uint32_t)__cvta_generic_to_shared((plane ? S0.alo[tr] : S0.ahi[tr]) + gq * 8);
        grow_[r] = tr;
      } else {
        const int plane = (w - 32) >> 3, gq = (w - 32) & 7;
        gsrc[r] = (plane ? Blo : Bhi) + (size_t)(tr + 1) * pl.Lp + n0 + gq * 8;
        sdst[r] = (uint32_t)__cvta_generic_to_shared((plane ? S0.blo[tr] : S0.bhi[tr]) + gq * 8);
        grow_[r] = tr + 1;
      }
    }
    auto stage_chunk = [&](int c, int s) {
      const size_t goff = (size_t)c * kKc * pl.Lp;
      const uint32_t soff = (uint32_t)(s * sizeof(Stage));
#pragma unroll
      for (int r = 0; r < 3; r++) {
        if (c * kKc + grow_[r] < M) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst[r] + soff), "l"(gsrc[r] + goff) : "memory");
        } else {
          asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(sdst[r] + soff), "r"(0) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int c = 0; c < kStages - 1; c++) {
      if (c < nchunks) stage_chunk(c, c);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // per-pair scale table: 2^(FA[t] + FB[t+1] - Fmax) as bf16 (exact), and the epilogue exponent Fmax - Z log2e.
    // One batched pass over the frames of my pair (global -> per-warp shared table), then the conversion in place.
    {
      const int blkI = i0 / 32 + wi, blkN = n0 / 32 + wn;
      int *sums = fsum + warp * Mp;
      unsigned char *st = stab + warp * Mp;
      int fmx = kNegBig;
      const int *pa = FA + blkI + (size_t)lane * pl.NBp, *pb = FB + blkN + (size_t)(lane + 1) * pl.NBp;
      const size_t step = (size_t)32 * pl.NBp;
#pragma unroll 4
      for (int t = lane; t < nsteps; t += 32, pa += step, pb += step) {
        const int x = __ldg(pa), y = __ldg(pb);
        const int sum = (x > kNegBig && y > kNegBig) ? x + y : kNegBig;
        sums[t] = sum;
        fmx = max(fmx, sum);
      }
      fmx = __reduce_max_sync(0xffffffffu, fmx);
      const int mpad = (nsteps + kKc - 1) / kKc * kKc;
      for (int t = lane; t < mpad; t += 32) {
        unsigned char bits = 0;                                // biased bf16 exponent of the scale (0: scale 0)
        if (t < nsteps) {
          const int d = sums[t] - fmx;                         // <= 0 (or hugely negative: no mass)
          if (d >= -126) bits = (unsigned char)(d + 127);
        }
        st[t] = bits;
      }
      if (lane == 0) dpair[warp] = fmx > kNegBig ? (float)((double)fmx - (double)Z * kL2E_D) : 0.f;
      __syncwarp();
    }
    __syncthreads();   // the frame sums alias the last operand stage, which the loop below starts filling
    // a pair entirely on or below the diagonal, beyond the graph or without any live row contributes nothing
    const bool pair_on = (n0 + 32 * wn + 31 > i0 + 32 * wi) && (i0 + 32 * wi < O) && (n0 + 32 * wn < O);

    for (int c = 0; c < nchunks; c++) {
      if (c + kStages - 1 < nchunks) stage_chunk(c + kStages - 1, (c + kStages - 1) % kStages);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
      __syncthreads();
      if (pair_on) {
        const Stage &S = stg[c % kStages];
        const int tig = lane & 3;
        const unsigned char *st = stab + warp * Mp + c * kKc;
        const uint32_t s_lo = ((uint32_t)st[2 * tig] << 7) | ((uint32_t)st[2 * tig + 1] << 23);
        const uint32_t s_hi = ((uint32_t)st[2 * tig + 8] << 7) | ((uint32_t)st[2 * tig + 9] << 23);
        if (__any_sync(0xffffffffu, (s_lo | s_hi) != 0u)) {
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int np = 0; np < 2; np++) {
            const int nb = 32 * wn + 16 * np;
            const int krow = ((lane >> 3) & 1) * 8 + (lane & 7), ncol = nb + (lane >> 4) * 8;
            uint32_t r[4];
            ldmatrix_x4_trans(r, &S.bhi[krow][ncol]);
            bh[2 * np][0] = r[0]; bh[2 * np][1] = r[1]; bh[2 * np + 1][0] = r[2]; bh[2 * np + 1][1] = r[3];
            ldmatrix_x4_trans(r, &S.blo[krow][ncol]);
            bl[2 * np][0] = r[0]; bl[2 * np][1] = r[1]; bl[2 * np + 1][0] = r[2]; bl[2 * np + 1][1] = r[3];
          }
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            const int mb = 32 * wi + 16 * mt;
            const int krow = (lane >> 4) * 8 + (lane & 7), mcol = mb + ((lane >> 3) & 1) * 8;
            uint32_t ah[4], al[4];
            ldmatrix_x4_trans(ah, &S.ahi[krow][mcol]);
            ldmatrix_x4_trans(al, &S.alo[krow][mcol]);
            ah[0] = hmul2_u32(ah[0], s_lo); ah[1] = hmul2_u32(ah[1], s_lo); ah[2] = hmul2_u32(ah[2], s_hi); ah[3] = hmul2_u32(ah[3], s_hi);
            al[0] = hmul2_u32(al[0], s_lo); al[1] = hmul2_u32(al[1], s_lo); al[2] = hmul2_u32(al[2], s_hi); al[3] = hmul2_u32(al[3], s_hi);
#pragma unroll
            for (int nt = 0; nt < 4; nt++) mma_bf16(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
#pragma unroll
            for (int nt = 0; nt < 4; nt++) mma_bf16(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
#pragma unroll
            for (int nt = 0; nt < 4; nt++) mma_bf16(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
          }
        }
      }
      __syncthreads();
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // stage the output tile (aliases the operand stages: all warps are past the last barrier)
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int m = 32 * wi + 16 * mt + gid, n = 32 * wn + 8 * nt + 2 * tig;
        *reinterpret_cast<float2 *>(cs + m * kCP + n) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<float2 *>(cs + (m + 8) * kCP + n) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
    __syncthreads();
  }

  // ---- epilogue: gl = go * exp2(links * log2e + Fmax - Z log2e) * G, one coalesced write per row, zeros elsewhere.
  // A warp owns 16 rows; their 32 transition values are fetched first (independent loads), then combined.
  const bool last_col = (n0 + kBN >= L);
#pragma unroll
  for (int r = 0; r < 16; r++) {
    const int ii = warp + 8 * r, i = i0 + ii;
    if (i >= L) break;
    float *grow = g + (int64_t)i * Tl;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int nn = lane + 32 * h;
      const int n = n0 + nn, k = n - i - 1;
      if (k < 0 || k >= Tl) continue;
      float v = 0.f;
      if (compute && i < O && n < O) v = gout * exp2f(fmaf(lks[ii * kBN + nn], kL2E, dpair[(ii >> 5) * 2 + h])) * cs[ii * kCP + nn];
      grow[k] = v;
    }
    if (last_col) {  // transitions that point beyond the padded graph: k >= L-1-i
      for (int k = max(0, n0 + kBN - i - 1) + lane; k < Tl; k += 32) grow[k] = 0.f;
    }
  }
}

int padded_rows(int M) { return (M + 15) / 16 * 16; }
size_t mma_smem_bytes(int M) {
  const int Mp = padded_rows(M);
  size_t a = kStages * sizeof(Stage) + (size_t)8 * Mp + sizeof(float) * kBI * kBN + 8 * sizeof(float) + 64;
  if ((size_t)8 * Mp * sizeof(int) > sizeof(Stage)) a += (size_t)8 * Mp * sizeof(int);
  const size_t c = sizeof(float) * kBI * kCP;
  return a > c ? a : c;
}

}  // namespace g3

size_t grad3_workspace_bytes(int B, int M, int L) { return g3::Planes::make(B, M, L).bytes; }
bool grad3_supported(int M, int L) { return M >= 2 && L >= 1 && g3::mma_smem_bytes(M) <= 100 * 1024; }

int launch_grad3(const float *go, const float *alpha, const float *beta, const float *match, const float *links,
                 const int64_t *olen, const int64_t *tlen, float *gm, float *gl, int B, int M, int L, int Tl,
                 void *workspace, cudaStream_t st) {
  using namespace g3;
  const Planes pl = Planes::make(B, M, L);
  prof_mark(3, st);
  {
    dim3 grid((pl.Lp / 128 + 7) / 8, M, B);
    grad_planes_kernel<<<grid, 256, 0, st>>>(go, alpha, beta, match, gm, (unsigned char *)workspace, pl, M, L);
    DAGB200_CHECK_LAUNCH("grad_planes_kernel");
  }
  prof_mark(4, st);
  {
    const int NI = (L + kBI - 1) / kBI, NN = (L + kBN - 1) / kBN;
    const size_t smem = mma_smem_bytes(M);
    cudaFuncSetAttribute(grad_links_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(NI * NN, B);
    grad_links_planes_kernel<<<grid, kThreads, smem, st>>>(go, beta, links, olen, tlen, gl, (const unsigned char *)workspace,
                                                          pl, M, L, Tl, NN, padded_rows(M));
    DAGB200_CHECK_LAUNCH("grad_links_planes_kernel");
  }
  prof_mark(5, st);
  return 0;
}

}  // namespace dagb200
