// dag_prep.cu -- one streaming pass over the transition plane: links[b][i][k] (fp32 log-probs, read once,
// coalesced along k) -> P'[i][j] = exp(links[i][j-i-1] - rmax[i]) laid out as the operand tiles of the blocked
// recurrences (layout: dag_tiles.cuh).  The row maximum rmax[i] is folded back in by the consumers, so a
// transition only underflows when it is > 87 nats below the best transition of the SAME source vertex.
#include "common.cuh"
#include "dag_tiles.cuh"

namespace dagb200 {

__device__ __forceinline__ uint32_t pack_bf16_pair(float lo_elem, float hi_elem) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo_elem, hi_elem);  // .x = low half
  return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float bf16_hi_part(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// pass 1: per-source-vertex maximum over the valid successors (one warp per vertex, coalesced along k)
__global__ void __launch_bounds__(256)
dag_rowmax_kernel(const float *__restrict__ links, const int64_t *__restrict__ olen, unsigned char *__restrict__ ws,
                  int L, int Tl, TileLayout lay) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + warp;
  if (i >= lay.NB * kBlk) return;
  const int O = min((int)olen[b], L);
  float mx = neg_inf_f();
  if (i < O) {
    const int kmax = min(Tl, O - 1 - i);
    const float *row = links + ((int64_t)b * L + i) * Tl;
    float m0 = mx, m1 = mx, m2 = mx, m3 = mx;
    int k = lane;
    for (; k + 96 < kmax; k += 128) {
      m0 = fmaxf(m0, __ldg(row + k)); m1 = fmaxf(m1, __ldg(row + k + 32));
      m2 = fmaxf(m2, __ldg(row + k + 64)); m3 = fmaxf(m3, __ldg(row + k + 96));
    }
    for (; k < kmax; k += 32) m0 = fmaxf(m0, __ldg(row + k));
    mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
  mx = warp_max(mx);
  if (lane == 0) reinterpret_cast<float *>(ws + (size_t)b * lay.sample_bytes + lay.off_rmax)[i] = mx;
}

// pass 2: one CTA (4 warps) per 32x32 tile (I <= J): P' = exp(links - rmax) into the operand layouts
__global__ void __launch_bounds__(128)
dag_tiles_kernel(const float *__restrict__ links, const int64_t *__restrict__ olen, unsigned char *__restrict__ ws,
                 int L, int Tl, TileLayout lay) {
  __shared__ float tile[kBlk][kBlk + 1];
  const int I = blockIdx.y, b = blockIdx.z;
  const int J = I + blockIdx.x;                    // blockIdx.x = block distance, 0 .. band
  const int O = min((int)olen[b], L);
  if (32 * I >= O) return;                         // block beyond the graph: never read by the recurrences
  const int NBv = (O + kBlk - 1) / kBlk;
  if (J > min(NBv - 1, I + band_blocks(Tl))) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float *E = links + (int64_t)b * L * Tl;
  unsigned char *base = ws + (size_t)b * lay.sample_bytes;
  const float *g_rmax = reinterpret_cast<const float *>(base + lay.off_rmax);
  // the 32x32 tile of P' in shared memory (row = source ii, column = destination jj); 8 rows per warp, all loads first
  float v[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int ii = warp * 8 + r;
    const int i = 32 * I + ii, j = 32 * J + lane, k = j - i - 1;
    v[r] = (i < O && j < O && k >= 0 && k < Tl) ? __ldg(E + (int64_t)i * Tl + k) : neg_inf_f();
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int ii = warp * 8 + r;
    const int i = 32 * I + ii;
    const float rm = i < O ? g_rmax[i] : neg_inf_f();
    tile[ii][lane] = __expf(v[r] - (rm == neg_inf_f() ? 0.f : rm));   // exp(-inf) = 0
  }
  __syncthreads();
  if (J == I) {
    // fp64 push tables [ci][cj] = weight of sweep column ci for the later column cj of the same block
    double *pA = reinterpret_cast<double *>(base + lay.off_pushA) + (size_t)I * kBlk * kBlk;
    double *pB = reinterpret_cast<double *>(base + lay.off_pushB) + (size_t)I * kBlk * kBlk;
    for (int r = warp; r < kBlk; r += 4) {
      pA[r * kBlk + lane] = (double)tile[r][lane];                            // P'[ci][cj]
      pB[r * kBlk + lane] = (double)tile[kBlk - 1 - lane][kBlk - 1 - r];      // P'[31-cj][31-ci]
    }
  } else {
    // canonical K-major core-matrix layout of tcgen05 (dag_dp4.cu): [plane hi|lo][k-core][n = 32][8 bf16 along K];
    // alpha direction: K = source, N = destination; beta direction: K = destination, N = source
    uint4 *tA = reinterpret_cast<uint4 *>(base + lay.off_tilesA + lay.idxA(I, J) * kTileBytes);
    uint4 *tB = reinterpret_cast<uint4 *>(base + lay.off_tilesB + lay.idxB(I, J) * kTileBytes);
    for (int x = threadIdx.x; x < 128; x += 128) {
      const int kc = x >> 5, n = x & 31;
      float a[8], bq[8];
#pragma unroll
      for (int e = 0; e < 8; e++) { a[e] = tile[8 * kc + e][n]; bq[e] = tile[n][8 * kc + e]; }
      uint4 ha, la, hb, lb;
      auto split2 = [](float x0, float x1, uint32_t &h, uint32_t &l) {
        const float h0 = bf16_hi_part(x0), h1 = bf16_hi_part(x1);
        h = pack_bf16_pair(h0, h1);
        l = pack_bf16_pair(x0 - h0, x1 - h1);
      };
      split2(a[0], a[1], ha.x, la.x); split2(a[2], a[3], ha.y, la.y); split2(a[4], a[5], ha.z, la.z); split2(a[6], a[7], ha.w, la.w);
      split2(bq[0], bq[1], hb.x, lb.x); split2(bq[2], bq[3], hb.y, lb.y); split2(bq[4], bq[5], hb.z, lb.z); split2(bq[6], bq[7], hb.w, lb.w);
      tA[kc * 32 + n] = ha; tA[128 + kc * 32 + n] = la;
      tB[kc * 32 + n] = hb; tB[128 + kc * 32 + n] = lb;
    }
  }
}

int launch_dag_prep(const float *links, const int64_t *olen, void *workspace, int B, int M, int L, int Tl,
                    cudaStream_t st) {
  TileLayout lay = TileLayout::make(L, M);
  {
    dim3 grid((lay.NB * kBlk + 7) / 8, B);
    dag_rowmax_kernel<<<grid, 256, 0, st>>>(links, olen, (unsigned char *)workspace, L, Tl, lay);
    DAGB200_CHECK_LAUNCH("dag_rowmax_kernel");
  }
  {
    const int nd = min(lay.NB, band_blocks(Tl) + 1);
    dim3 grid(nd, lay.NB, B);
    dag_tiles_kernel<<<grid, 128, 0, st>>>(links, olen, (unsigned char *)workspace, L, Tl, lay);
    DAGB200_CHECK_LAUNCH("dag_tiles_kernel");
  }
  return 0;
}

}  // namespace dagb200
