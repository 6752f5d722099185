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

// reg(nt, r) of the fragment order, from an operand tile addressed Bop(k, n)
template <bool LO, typename F>
__device__ __forceinline__ uint32_t frag_reg(F bop, int ks, int nt, int r, int gid, int tig) {
  const int k = 16 * ks + 2 * tig + 8 * r, n = 8 * nt + gid;
  float x0 = bop(k, n), x1 = bop(k + 1, n);
  if (LO) { x0 -= bf16_hi_part(x0); x1 -= bf16_hi_part(x1); }
  return pack_bf16_pair(x0, x1);
}

__global__ void __launch_bounds__(256)
dag_prep_kernel(const float *__restrict__ links, const int64_t *__restrict__ olen, unsigned char *__restrict__ ws,
                int L, int Tl, TileLayout lay) {
  __shared__ float tile[kBlk][kBlk + 1];
  __shared__ float s_rmax[kBlk];
  const int I = blockIdx.x, b = blockIdx.y;
  const int O = min((int)olen[b], L);
  if (32 * I >= O) return;  // block beyond the graph: never read by the recurrences
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float *E = links + (int64_t)b * L * Tl;
  unsigned char *base = ws + (size_t)b * lay.sample_bytes;
  float *g_rmax = reinterpret_cast<float *>(base + lay.off_rmax);

  // phase 1: row maxima over the valid successors of each source vertex of this block
  for (int ii = warp; ii < kBlk; ii += 8) {
    const int i = 32 * I + ii;
    float mx = neg_inf_f();
    if (i < O) {
      const int kmax = min(Tl, O - 1 - i);
      const float *row = E + (int64_t)i * Tl;
      for (int k = lane; k < kmax; k += 32) mx = fmaxf(mx, __ldg(row + k));
    }
    mx = warp_max(mx);
    if (lane == 0) { s_rmax[ii] = mx; g_rmax[i < lay.NB * kBlk ? i : 0] = mx; }
  }
  __syncthreads();

  const int NBv = (O + kBlk - 1) / kBlk;
  const int Jend = min(NBv - 1, I + band_blocks(Tl));
  const int gid = lane >> 2, tig = lane & 3;
  for (int J = I; J <= Jend; J++) {
    // phase 2a: the 32x32 tile of P' in shared memory (row = source ii, column = destination jj)
    for (int ii = warp; ii < kBlk; ii += 8) {
      const int i = 32 * I + ii, j = 32 * J + lane, k = j - i - 1;
      float p = 0.f;
      if (i < O && j < O && k >= 0 && k < Tl) {
        const float rm = s_rmax[ii];
        p = __expf(__ldg(E + (int64_t)i * Tl + k) - (rm == neg_inf_f() ? 0.f : rm));
      }
      tile[ii][lane] = p;
    }
    __syncthreads();
    if (J == I) {
      float *dA = reinterpret_cast<float *>(base + lay.off_diagA) + (size_t)I * kBlk * kBlk;
      float *dB = reinterpret_cast<float *>(base + lay.off_diagB) + (size_t)I * kBlk * kBlk;
      // [cj][ci] = weight of in-block predecessor ci for cell cj, both in sweep order (alpha: ascending vertex,
      // beta: descending vertex), zero for ci >= cj
      for (int r = warp; r < kBlk; r += 8) {
        dA[r * kBlk + lane] = tile[lane][r];                       // P'[ci][cj]
        dB[r * kBlk + lane] = tile[kBlk - 1 - r][kBlk - 1 - lane];  // P'[31-cj][31-ci]
      }
      // fp64 push tables of the column-major kernel: [ci][cj] = weight of sweep column ci for the later column cj
      double *pA = reinterpret_cast<double *>(base + lay.off_pushA) + (size_t)I * kBlk * kBlk;
      double *pB = reinterpret_cast<double *>(base + lay.off_pushB) + (size_t)I * kBlk * kBlk;
      for (int r = warp; r < kBlk; r += 8) {
        pA[r * kBlk + lane] = (double)tile[r][lane];                          // P'[ci][cj]
        pB[r * kBlk + lane] = (double)tile[kBlk - 1 - lane][kBlk - 1 - r];    // P'[31-cj][31-ci]
      }
    } else {
      const int q = warp;  // 8 warps <-> 8 units
      const int qq = q & 3, ks = qq >> 1, nt0 = 2 * (qq & 1);
      uint4 *tA = reinterpret_cast<uint4 *>(base + lay.off_tilesA + lay.idxA(I, J) * kTileBytes);
      uint4 *tB = reinterpret_cast<uint4 *>(base + lay.off_tilesB + lay.idxB(I, J) * kTileBytes);
      auto bopA = [&](int k, int n) { return tile[k][n]; };  // K = source, N = destination
      auto bopB = [&](int k, int n) { return tile[n][k]; };  // K = destination, N = source
      uint4 ua, ub;
      if (q < 4) {
        ua.x = frag_reg<false>(bopA, ks, nt0, 0, gid, tig); ua.y = frag_reg<false>(bopA, ks, nt0, 1, gid, tig);
        ua.z = frag_reg<false>(bopA, ks, nt0 + 1, 0, gid, tig); ua.w = frag_reg<false>(bopA, ks, nt0 + 1, 1, gid, tig);
        ub.x = frag_reg<false>(bopB, ks, nt0, 0, gid, tig); ub.y = frag_reg<false>(bopB, ks, nt0, 1, gid, tig);
        ub.z = frag_reg<false>(bopB, ks, nt0 + 1, 0, gid, tig); ub.w = frag_reg<false>(bopB, ks, nt0 + 1, 1, gid, tig);
      } else {
        ua.x = frag_reg<true>(bopA, ks, nt0, 0, gid, tig); ua.y = frag_reg<true>(bopA, ks, nt0, 1, gid, tig);
        ua.z = frag_reg<true>(bopA, ks, nt0 + 1, 0, gid, tig); ua.w = frag_reg<true>(bopA, ks, nt0 + 1, 1, gid, tig);
        ub.x = frag_reg<true>(bopB, ks, nt0, 0, gid, tig); ub.y = frag_reg<true>(bopB, ks, nt0, 1, gid, tig);
        ub.z = frag_reg<true>(bopB, ks, nt0 + 1, 0, gid, tig); ub.w = frag_reg<true>(bopB, ks, nt0 + 1, 1, gid, tig);
      }
      tA[q * 32 + lane] = ua;
      tB[q * 32 + lane] = ub;
    }
    __syncthreads();
  }
}

int launch_dag_prep(const float *links, const int64_t *olen, void *workspace, int B, int M, int L, int Tl, cudaStream_t st) {
  TileLayout lay = TileLayout::make(L, M);
  dim3 grid(lay.NB, B);
  dag_prep_kernel<<<grid, 256, 0, st>>>(links, olen, (unsigned char *)workspace, L, Tl, lay);
  DAGB200_CHECK_LAUNCH("dag_prep_kernel");
  return 0;
}

}  // namespace dagb200
