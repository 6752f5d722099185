// dag_grad2.cu -- transition gradient of the DAG loss as a tensor-core contraction, for sm_100a.
//
// Replaces calculate_grad_links_kernel (reference dag_loss.cu:432-485) on the fp32 path:
//
//   gl[b,i,k] = go[b] * sum_{t=0}^{Tn-2} exp(alpha[t,i] + beta[t+1,n] + links[i,k] - Z),   n = i+k+1
//             = go[b] * exp(links[i,k] - Emax) * G[i,n],      G = A^T B  (contraction over the target index t)
//   A[t,i] = exp(alpha[t,i] - u[t])            u[t]  = max over the LIVE (beta finite) vertices of the 128-block
//   B[t,n] = exp(beta[t+1,n] + u[t] + Emax - Z)                Emax = max of the tile's transitions
//
// The reference spends one exp per (i,k,t) -- 8.4e9 at C2 -- walking alpha/beta down columns.  Here a CTA owns a
// 128 x 128 (source x destination) tile, generates the two operand panels once per 16 target rows (1 exp per 64
// MACs), splits them into bf16 hi/lo (3 mma.sync.m16n8k16 per product, ~2^-16 relative error, fp32 accumulation)
// and multiplies by exp(links) in the epilogue while streaming the links tile and writing grad_links once,
// coalesced, including the zero padding.
//
// Dynamic range: alpha spans > 87 nats inside a 128-vertex window on tight lattices, so the operands are stacked
// along K in two exponent levels 60 nats apart (A level 1 holds exp(x+60) for x < -60, B level 1 holds
// exp(y-60)); the level cancels inside each product.  Usable window: 147 nats below the best live vertex.
#include "common.cuh"

namespace dagb200 {

constexpr int kG2Tile = 128;
constexpr int kG2Threads = 256;
constexpr int kG2Kc = 16;            // target rows per K chunk
constexpr int kG2Pitch = 136;        // bf16 elements per operand row (128 + 8: conflict-free ldmatrix)
constexpr int kG2CPitch = 132;       // floats per row of the staged output tile
constexpr float kG2Level = 60.f;

__device__ __forceinline__ void mma_bf16_g2(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
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

__device__ __forceinline__ void split_bf16x2_g2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<uint32_t *>(&h);
  lo = *reinterpret_cast<uint32_t *>(&l);
}

struct G2Planes {            // [level][hi/lo][kG2Kc][kG2Pitch]
  __nv_bfloat16 v[2][2][kG2Kc][kG2Pitch];
};

__global__ void __launch_bounds__(kG2Threads, 2)
grad_links_mma_kernel(const float *__restrict__ go, const float *__restrict__ alpha, const float *__restrict__ beta,
                      const float *__restrict__ links, const int64_t *__restrict__ olen,
                      const int64_t *__restrict__ tlen, float *__restrict__ gl, int M, int L, int Tl, int NI) {
  extern __shared__ __align__(16) unsigned char g2_smem[];
  const int b = blockIdx.y;
  const int I = blockIdx.x / NI, J = blockIdx.x % NI;
  if (J < I) return;
  const int i0 = I * kG2Tile, n0 = J * kG2Tile;
  if (n0 - (i0 + kG2Tile - 1) - 1 >= Tl) return;       // tile entirely beyond the transition band: no storage
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int O = (int)olen[b], Tn = (int)tlen[b];
  const int64_t lat = (int64_t)M * L;
  const float *a = alpha + b * lat, *be = beta + b * lat;
  const float *E = links + (int64_t)b * L * Tl;
  float *g = gl + (int64_t)b * L * Tl;
  const float Z = be[0];
  const float gout = go[b];
  const float ninf = neg_inf_f();
  const bool dead = isinf(Z) || O > L || Tn > M || Tn < 2 || O < 2;

  // carve shared memory: two staging buffers (cp.async, one chunk ahead), frames, operand planes; the output tile
  // aliases all of it after the K loop
  constexpr int kStageFloats = 3 * kG2Kc * kG2Tile;             // alpha[t][i-blk], beta[t][i-blk], beta[t+1][n-blk]
  float *stage_base = reinterpret_cast<float *>(g2_smem);        // [2][3][16][128]
  float *u_s = stage_base + 2 * kStageFloats;                    // [16]
  int *lev1_s = reinterpret_cast<int *>(u_s + 16);               // [16] row needs the second exponent level
  float *red_s = u_s + 32;                                       // [16]
  G2Planes *pa = reinterpret_cast<G2Planes *>(u_s + 48);
  G2Planes *pb = pa + 1;
  float *cs = reinterpret_cast<float *>(g2_smem);                // [128][kG2CPitch] after the loop

  // ---- tile maximum of the transitions (valid entries only) -------------------------------------------
  bool compute = !dead && i0 < O && n0 < O && (n0 + kG2Tile - 1 > i0);
  float emax = ninf;
  if (compute) {  // pull the tile's transition rows towards L2: they are read now (maximum) and again in the epilogue
    for (int x = tid; x < kG2Tile * 5; x += kG2Threads) {
      const int ii = x / 5, seg = x % 5;
      const int i = i0 + ii;
      const int k = max(0, n0 - i - 1) + seg * 32;
      if (i < O && k < Tl && k < n0 + kG2Tile - i - 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(E + (int64_t)i * Tl + k));
    }
  }
  if (compute) {
    for (int ii = warp; ii < kG2Tile; ii += kG2Threads / 32) {
      const int i = i0 + ii;
      if (i >= O) break;
      const int klo = max(0, n0 - i - 1), khi = min(min(Tl, n0 + kG2Tile - i - 1), O - i - 1);
      const float *row = E + (int64_t)i * Tl;
      for (int k = klo + lane; k < khi; k += 32) emax = fmaxf(emax, __ldg(row + k));
    }
    emax = warp_max(emax);
    if (lane == 0) red_s[warp] = emax;
    __syncthreads();
    emax = red_s[0];
#pragma unroll
    for (int w = 1; w < kG2Threads / 32; w++) emax = fmaxf(emax, red_s[w]);
    __syncthreads();
    if (emax == ninf) compute = false;   // no usable transition in this tile
  }

  float acc[4][4][4];
#pragma unroll
  for (int x = 0; x < 4; x++)
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[x][y][e] = 0.f;

  if (compute) {
    const int wi = warp >> 2, wn = warp & 3;     // warp tile: 64 sources x 32 destinations
    const int nsteps = Tn - 1;
    const float shift = emax - Z;
    const bool vec16 = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(be)) % 16 == 0);
    // asynchronous staging of one 16-row chunk: alpha[t][i-block], beta[t][i-block] (liveness), beta[t+1][n-block]
    auto stage_chunk = [&](int t0, int buf) {
      float *sa = stage_base + buf * kStageFloats, *sbi = sa + kG2Kc * kG2Tile, *sbn = sbi + kG2Kc * kG2Tile;
      if (vec16) {
        for (int x = tid; x < kG2Kc * kG2Tile / 4; x += kG2Threads) {
          const int tr = x >> 5, c = (x & 31) * 4;
          const int t = t0 + tr;
          const bool tv = t < nsteps;
          const int i = i0 + c, n = n0 + c;
          const uint32_t da = (uint32_t)__cvta_generic_to_shared(sa + tr * kG2Tile + c);
          const uint32_t db = (uint32_t)__cvta_generic_to_shared(sbi + tr * kG2Tile + c);
          const uint32_t dn = (uint32_t)__cvta_generic_to_shared(sbn + tr * kG2Tile + c);
          if (tv && i + 3 < L) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(a + (int64_t)t * L + i) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db), "l"(be + (int64_t)t * L + i) : "memory");
          } else {
            *reinterpret_cast<float4 *>(sa + tr * kG2Tile + c) = make_float4(ninf, ninf, ninf, ninf);
            *reinterpret_cast<float4 *>(sbi + tr * kG2Tile + c) = make_float4(ninf, ninf, ninf, ninf);
          }
          if (tv && n + 3 < L)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dn), "l"(be + (int64_t)(t + 1) * L + n) : "memory");
          else
            *reinterpret_cast<float4 *>(sbn + tr * kG2Tile + c) = make_float4(ninf, ninf, ninf, ninf);
        }
      } else {
        for (int x = tid; x < kG2Kc * kG2Tile; x += kG2Threads) {
          const int tr = x >> 7, c = x & 127;
          const int t = t0 + tr;
          const bool tv = t < nsteps;
          const int i = i0 + c, n = n0 + c;
          sa[x] = (tv && i < L) ? a[(int64_t)t * L + i] : ninf;
          sbi[x] = (tv && i < L) ? be[(int64_t)t * L + i] : ninf;
          sbn[x] = (tv && n < L) ? be[(int64_t)(t + 1) * L + n] : ninf;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_chunk(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < nsteps; t0 += kG2Kc, buf ^= 1) {
      // next chunk into the other buffer (its last readers passed the barrier after the previous generation)
      if (t0 + kG2Kc < nsteps) stage_chunk(t0 + kG2Kc, buf ^ 1);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      const float *stage_a = stage_base + buf * kStageFloats, *stage_bi = stage_a + kG2Kc * kG2Tile,
                  *stage_bn = stage_bi + kG2Kc * kG2Tile;
      // frame u[t] = max alpha over live vertices of the block (2 rows per warp); does any live vertex sit more
      // than one level below it?
#pragma unroll
      for (int rr = 0; rr < 2; rr++) {
        const int tr = warp * 2 + rr;
        float m = ninf, mn = __int_as_float(0x7f800000);
#pragma unroll
        for (int c = lane; c < kG2Tile; c += 32) {
          const int i = i0 + c;
          if (i < O && stage_bi[tr * kG2Tile + c] > ninf) {
            const float v = stage_a[tr * kG2Tile + c];
            m = fmaxf(m, v);
            if (v > ninf) mn = fminf(mn, v);
          }
        }
        m = warp_max(m);
        mn = -warp_max(-mn);
        if (lane == 0) { u_s[tr] = m; lev1_s[tr] = (m > ninf && mn < m - kG2Level) ? 1 : 0; }
      }
      __syncthreads();
      bool need1 = false;
#pragma unroll
      for (int tr = 0; tr < kG2Kc; tr++) need1 = need1 || lev1_s[tr] != 0;
      // operand planes: two exponent levels, bf16 hi/lo; 4 consecutive vertices per thread-iteration
      for (int x = tid; x < kG2Kc * kG2Tile / 4; x += kG2Threads) {
        const int tr = x >> 5, c = (x & 31) * 4;
        const float u = u_s[tr];
        const float4 av = *reinterpret_cast<const float4 *>(stage_a + tr * kG2Tile + c);
        const float4 lv = *reinterpret_cast<const float4 *>(stage_bi + tr * kG2Tile + c);
        const float4 bv = *reinterpret_cast<const float4 *>(stage_bn + tr * kG2Tile + c);
        const float aa[4] = {av.x, av.y, av.z, av.w}, ll[4] = {lv.x, lv.y, lv.z, lv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
        float a0[4], a1[4], b0[4], b1[4];
        const bool rowlive = u > ninf;
        const float ub = u + shift;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          // A: exp(alpha - u) for live vertices, split over the two levels
          const float xa = (rowlive && i0 + c + e < O && ll[e] > ninf) ? aa[e] - u : ninf;
          const bool hi_level = xa >= -kG2Level;
          const float ea = __expf(hi_level ? xa : xa + kG2Level);        // exp(-inf) = 0
          a0[e] = hi_level ? ea : 0.f;
          a1[e] = hi_level ? 0.f : ea;
          // B: exp(beta + u + Emax - Z), clamped
          const float y = (rowlive && n0 + c + e < O) ? bb[e] + ub : ninf;
          b0[e] = __expf(fminf(y, 80.f));
          b1[e] = need1 ? __expf(fminf(y - kG2Level, 80.f)) : 0.f;
        }
        uint2 h, l;
        split_bf16x2_g2(a0[0], a0[1], h.x, l.x); split_bf16x2_g2(a0[2], a0[3], h.y, l.y);
        *reinterpret_cast<uint2 *>(&pa->v[0][0][tr][c]) = h; *reinterpret_cast<uint2 *>(&pa->v[0][1][tr][c]) = l;
        split_bf16x2_g2(b0[0], b0[1], h.x, l.x); split_bf16x2_g2(b0[2], b0[3], h.y, l.y);
        *reinterpret_cast<uint2 *>(&pb->v[0][0][tr][c]) = h; *reinterpret_cast<uint2 *>(&pb->v[0][1][tr][c]) = l;
        if (need1) {
          split_bf16x2_g2(a1[0], a1[1], h.x, l.x); split_bf16x2_g2(a1[2], a1[3], h.y, l.y);
          *reinterpret_cast<uint2 *>(&pa->v[1][0][tr][c]) = h; *reinterpret_cast<uint2 *>(&pa->v[1][1][tr][c]) = l;
          split_bf16x2_g2(b1[0], b1[1], h.x, l.x); split_bf16x2_g2(b1[2], b1[3], h.y, l.y);
          *reinterpret_cast<uint2 *>(&pb->v[1][0][tr][c]) = h; *reinterpret_cast<uint2 *>(&pb->v[1][1][tr][c]) = l;
        }
      }
      __syncthreads();
      // tensor-core contraction over the 16 rows of this chunk (x2 levels, x3 split products)
#pragma unroll
      for (int lev = 0; lev < 2; lev++) {
        if (lev == 1 && !need1) break;
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int np = 0; np < 2; np++) {
          const int nb = 32 * wn + 16 * np;
          const int krow = ((lane >> 3) & 1) * 8 + (lane & 7), ncol = nb + (lane >> 4) * 8;
          uint32_t r[4];
          ldmatrix_x4_trans(r, &pb->v[lev][0][krow][ncol]);
          bh[2 * np][0] = r[0]; bh[2 * np][1] = r[1]; bh[2 * np + 1][0] = r[2]; bh[2 * np + 1][1] = r[3];
          ldmatrix_x4_trans(r, &pb->v[lev][1][krow][ncol]);
          bl[2 * np][0] = r[0]; bl[2 * np][1] = r[1]; bl[2 * np + 1][0] = r[2]; bl[2 * np + 1][1] = r[3];
        }
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
          const int mb = 64 * wi + 16 * mt;
          const int krow = (lane >> 4) * 8 + (lane & 7), mcol = mb + ((lane >> 3) & 1) * 8;
          uint32_t ah[4], al[4];
          ldmatrix_x4_trans(ah, &pa->v[lev][0][krow][mcol]);
          ldmatrix_x4_trans(al, &pa->v[lev][1][krow][mcol]);
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            mma_bf16_g2(acc[mt][nt], ah, bh[nt][0], bh[nt][1]);
            mma_bf16_g2(acc[mt][nt], al, bh[nt][0], bh[nt][1]);
            mma_bf16_g2(acc[mt][nt], ah, bl[nt][0], bl[nt][1]);
          }
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // stage the output tile (aliases the operand buffers: all warps are past the last barrier)
    const int gid = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int m = 64 * wi + 16 * mt + gid, n = 32 * wn + 8 * nt + 2 * tig;
        *reinterpret_cast<float2 *>(cs + m * kG2CPitch + n) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
        *reinterpret_cast<float2 *>(cs + (m + 8) * kG2CPitch + n) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
      }
    __syncthreads();
  }

  // ---- epilogue: gl = go * exp(links - Emax) * G, one coalesced write per row, zeros elsewhere -----------
  const bool last_col = (n0 + kG2Tile >= L);
  for (int ii = warp; ii < kG2Tile; ii += kG2Threads / 32) {
    const int i = i0 + ii;
    if (i >= L) break;
    const float *erow = E + (int64_t)i * Tl;
    float *grow = g + (int64_t)i * Tl;
    for (int nn = lane; nn < kG2Tile; nn += 32) {
      const int n = n0 + nn, k = n - i - 1;
      if (k < 0 || k >= Tl) continue;
      float v = 0.f;
      if (compute && i < O && n < O) v = gout * __expf(__ldg(erow + k) - emax) * cs[ii * kG2CPitch + nn];
      grow[k] = v;
    }
    if (last_col) {  // transitions that point beyond the padded graph: k >= L-1-i
      for (int k = max(0, n0 + kG2Tile - i - 1) + lane; k < Tl; k += 32) grow[k] = 0.f;
    }
  }
}

size_t g2_smem_bytes() {
  const size_t stage = sizeof(float) * (2 * 3 * kG2Kc * kG2Tile + 48) + 2 * sizeof(G2Planes);
  const size_t cst = sizeof(float) * kG2Tile * kG2CPitch;
  return stage > cst ? stage : cst;
}

int launch_grad_links_mma(const float *go, const float *alpha, const float *beta, const float *links,
                          const int64_t *olen, const int64_t *tlen, float *gl, int B, int M, int L, int Tl,
                          cudaStream_t st) {
  const int NI = (L + kG2Tile - 1) / kG2Tile;
  const size_t smem = g2_smem_bytes();
  cudaFuncSetAttribute(grad_links_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid(NI * NI, B);
  grad_links_mma_kernel<<<grid, kG2Threads, smem, st>>>(go, alpha, beta, links, olen, tlen, gl, M, L, Tl, NI);
  DAGB200_CHECK_LAUNCH("grad_links_mma_kernel");
  return 0;
}

}  // namespace dagb200
