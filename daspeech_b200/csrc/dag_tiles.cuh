// dag_tiles.cuh -- layout of the per-call transition-probability scratch shared by the precompute kernel
// (dag_prep.cu) and the blocked recurrences (dag_dp2.cu).
//
// Vertices are grouped in blocks of 32.  For every utterance the scratch holds, for NB = ceil(L/32):
//   rmax  [NB*32]            per-source-vertex max_k links[i][k] over valid successors (-inf if none)
//   diagA [NB][32][32] fp32  [cj][ci] = P'[32J+ci][32J+cj]        in-block predecessor weights of the alpha chain
//   diagB [NB][32][32] fp32  [cj][ci] = P'[32J+31-cj][32J+31-ci]  the same for beta, in its (descending) sweep order
//   tilesA[NB(NB-1)/2] 4 KB  off-diagonal tile (I<J) as the B operand of mma.m16n8k16 with K = source vertex,
//                            N = destination vertex, bf16 hi plane then bf16 lo plane, in FRAGMENT order
//   tilesB[NB(NB-1)/2] 4 KB  the same tile as the B operand with K = destination vertex, N = source vertex
//   afragA/afragB [NC][NB] 4 KB  written by the recurrences themselves: the previous-row masses of (chunk, block)
//                            as the A operand (rows = the chunk's 32 previous rows, K = vertex), normalised per row
//                            by the integer frame of the frame table, bf16 hi/lo, fragment order:
//                            unit = slice*4 + ks*2 + hl (slice = 16-row half), .x.y.z.w = a0..a3 of mma.m16n8k16
// with P'[i][j] = exp(links[i][j-i-1] - rmax[i])  (0 outside the band / beyond the graph).
//
// Fragment order of one 32(K) x 32(N) operand tile Bop[k][n]: eight uint4 "units" q, unit q is stored as 32
// consecutive uint4 (one per lane -> a warp load of a unit is one fully coalesced 512-byte access).
//   q < 4: hi plane, q >= 4: lo plane;  qq = q & 3;  ks = qq >> 1 (k16 step);  nt0 = 2 * (qq & 1)
//   .x = reg(nt0, 0)  .y = reg(nt0, 1)  .z = reg(nt0+1, 0)  .w = reg(nt0+1, 1)
//   reg(nt, r) = { Bop[16ks + 2tig + 8r][8nt + gid] (low half), Bop[16ks + 2tig + 8r + 1][8nt + gid] (high half) }
//   gid = lane >> 2, tig = lane & 3   -- exactly the b0/b1 registers of mma.sync.m16n8k16.row.col
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace dagb200 {

constexpr int kBlk = 32;            // vertices per block
constexpr int kTileBytes = 4096;    // one off-diagonal operand tile (hi + lo planes)

struct TileLayout {
  int NB;                // blocks per utterance
  int NC;                // 32-row chunks per utterance
  size_t off_rmax, off_diagA, off_diagB, off_tilesA, off_tilesB, off_afragA, off_afragB, sample_bytes;
  // column-major kernel (dag_dp3.cu): fp64 in-block push tables [NB][ci][cj] (sweep order, weight of the already
  // computed column ci for the later column cj) and the row handed from the last chunk of a pass to the first
  // chunk of the next one: [2 (pass parity)][NB][32] fp64 predecessor sums + [2][NB] integer anchors
  size_t off_pushA, off_pushB, off_passA, off_passB, off_passfA, off_passfB;
  // tcgen05 kernel (dag_dp4.cu): previous-row masses as the A operand in the canonical K-major core-matrix layout:
  // [2 planes (bf16 hi, lo)][NB*4 k-cores of 8 vertices (sweep order)][Mr consumer rows][16 bytes]
  int Mr;
  size_t off_aopA, off_aopB;
  __host__ __device__ static inline TileLayout make(int L, int M = 2) {
    TileLayout t;
    t.NB = (L + kBlk - 1) / kBlk;
    t.NC = (M - 1 + kBlk - 1) / kBlk;
    if (t.NC < 1) t.NC = 1;
    const size_t ntri = (size_t)t.NB * (t.NB - 1) / 2;
    size_t o = 0;
    t.off_rmax = o;   o += (size_t)t.NB * kBlk * sizeof(float);
    t.off_diagA = o;  o += (size_t)t.NB * kBlk * kBlk * sizeof(float);
    t.off_diagB = o;  o += (size_t)t.NB * kBlk * kBlk * sizeof(float);
    t.off_tilesA = o; o += ntri * kTileBytes;
    t.off_tilesB = o; o += ntri * kTileBytes;
    // A-operand fragment cache of the recurrences: [chunk][block in sweep order] 4 KB each, per direction
    t.off_afragA = o; o += (size_t)t.NC * t.NB * kTileBytes;
    t.off_afragB = o; o += (size_t)t.NC * t.NB * kTileBytes;
    o = (o + 255) & ~(size_t)255;
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
  // tile (I < J) for the alpha direction: the J-1... panel of destination block J is contiguous in I
  __host__ __device__ inline size_t idxA(int I, int J) const { return (size_t)J * (J - 1) / 2 + I; }
  // tile (Jb < Nb) for the beta direction: the panel of source block Jb is contiguous in Nb
  __host__ __device__ inline size_t idxB(int Jb, int Nb) const {
    return (size_t)Jb * (2 * NB - Jb - 1) / 2 + (Nb - Jb - 1);
  }
};

// number of 32-blocks a band of T transitions reaches beyond the adjacent block
__host__ __device__ inline int band_blocks(int T) { return 1 + (T - 1) / kBlk; }

}  // namespace dagb200
