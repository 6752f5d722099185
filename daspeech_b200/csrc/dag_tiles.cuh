// dag_tiles.cuh -- layout of the per-call transition-probability scratch shared by the precompute kernels
// (dag_prep.cu) and the blocked recurrences (dag_dp4.cu).
//
// Vertices are grouped in blocks of 32.  For every utterance the scratch holds, for NB = ceil(L/32):
//   rmax  [NB*32]            per-source-vertex max_k links[i][k] over valid successors (-inf if none)
//   tilesA[NB(NB-1)/2] 4 KB  off-diagonal tile (I<J) as the B operand of tcgen05.mma with K = source vertex,
//                            N = destination vertex: canonical K-major core-matrix layout
//                            [plane bf16 hi | lo][k-core of 8 vertices][n = 32][8 bf16 along K]
//   tilesB[NB(NB-1)/2] 4 KB  the same tile with K = destination vertex, N = source vertex (beta direction)
//   pushA / pushB [NB][32][32] fp64  in-block push tables [ci][cj] (sweep order): weight of the already computed
//                            column ci for the later column cj of the same block
//   passA / passB, passfA / passfB   the row handed from the last chunk of a 256-row pass to the first chunk of the
//                            next one: [2 (pass parity)][NB][32] fp64 predecessor sums + [2][NB] integer frames
//   aopA / aopB              written by the recurrences themselves: previous-row masses as the A operand,
//                            [source block][128-row tile][plane][k-core][row][16 bytes]
// with P'[i][j] = exp(links[i][j-i-1] - rmax[i])  (0 outside the band / beyond the graph).
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace dagb200 {

constexpr int kBlk = 32;            // vertices per block
constexpr int kTileBytes = 4096;    // one off-diagonal operand tile (hi + lo planes)

struct TileLayout {
  int NB;                // blocks per utterance
  int NC;                // 32-row chunks per utterance
  int Mr;                // consumer rows of the A-operand store (multiple of 256: whole passes)
  size_t off_rmax, off_tilesA, off_tilesB, off_pushA, off_pushB, off_passA, off_passB, off_passfA, off_passfB,
      off_aopA, off_aopB, sample_bytes;
  __host__ __device__ static inline TileLayout make(int L, int M = 2) {
    TileLayout t;
    t.NB = (L + kBlk - 1) / kBlk;
    t.NC = (M - 1 + kBlk - 1) / kBlk;
    if (t.NC < 1) t.NC = 1;
    const size_t ntri = (size_t)t.NB * (t.NB - 1) / 2;
    size_t o = 0;
    t.off_rmax = o;   o += (size_t)t.NB * kBlk * sizeof(float);
    o = (o + 255) & ~(size_t)255;
    t.off_tilesA = o; o += ntri * kTileBytes;
    t.off_tilesB = o; o += ntri * kTileBytes;
    t.off_pushA = o;  o += (size_t)t.NB * kBlk * kBlk * sizeof(double);
    t.off_pushB = o;  o += (size_t)t.NB * kBlk * kBlk * sizeof(double);
    t.off_passA = o;  o += (size_t)2 * t.NB * kBlk * sizeof(double);
    t.off_passB = o;  o += (size_t)2 * t.NB * kBlk * sizeof(double);
    t.off_passfA = o; o += (((size_t)2 * t.NB * sizeof(int)) + 255) & ~(size_t)255;
    t.off_passfB = o; o += (((size_t)2 * t.NB * sizeof(int)) + 255) & ~(size_t)255;
    o = (o + 255) & ~(size_t)255;
    t.Mr = ((t.NC + 7) / 8) * 256;
    t.off_aopA = o;   o += (size_t)2 * t.NB * 4 * t.Mr * 16;
    t.off_aopB = o;   o += (size_t)2 * t.NB * 4 * t.Mr * 16;
    t.sample_bytes = (o + 255) & ~(size_t)255;
    return t;
  }
  // tile (I < J) for the alpha direction: the panel of destination block J is contiguous in I
  __host__ __device__ inline size_t idxA(int I, int J) const { return (size_t)J * (J - 1) / 2 + I; }
  // tile (Jb < Nb) for the beta direction: the panel of source block Jb is contiguous in Nb
  __host__ __device__ inline size_t idxB(int Jb, int Nb) const {
    return (size_t)Jb * (2 * NB - Jb - 1) / 2 + (Nb - Jb - 1);
  }
};

// number of 32-blocks a band of T transitions reaches beyond the adjacent block
__host__ __device__ inline int band_blocks(int T) { return 1 + (T - 1) / kBlk; }

}  // namespace dagb200
