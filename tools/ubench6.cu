// The chain warp's column step in isolation (dag_dp4.cu chain_phase1): per column, lanes = rows:
//   rm = shfl_up(a[K]); m = (fd + rm) * ew; mrow[K] = hi(m); a[K+1..7] += m * u[K][K+1..7]   (fp64, u from shared memory)
// How does the time per column scale with the number of warps doing this on one SM (1, 2, 4, 8; + 8 idle-ish others)?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>   // 0: full, 1: push weights in registers (no LDS), 2: no shuffle (rm = a[K] of the same lane)
__global__ void k(double *out, int iters, long long *cyc, int slot) {
  __shared__ double ut[32 * 32];
  __shared__ float mrow[16][33];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) ut[i] = 1.0 / (1.0 + (i & 31) + (i >> 5));
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 1e-3 * (lane + i);
  const double fd = out[0] + 1e-9 * lane, ew = out[1] + 0.999;
  double ureg[8];
#pragma unroll
  for (int i = 0; i < 8; i++) ureg[i] = ut[i];
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const int G = it & 3;
#pragma unroll
    for (int K = 0; K < 8; K++) {
      double rm = MODE == 2 ? a[K] : __shfl_up_sync(0xffffffffu, a[K], 1);
      if (lane == 0) rm = fd;
      const double m = (fd + rm) * ew;
      mrow[warp][K] = __int_as_float(__double2hiint(m));
      const double *ur = ut + (8 * G + K) * 32 + 8 * G;
#pragma unroll
      for (int k2 = K + 1; k2 < 8; k2++) a[k2] = fma(m, MODE == 1 ? ureg[k2] : ur[k2], a[k2]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = a[i] * 1e-3 + 1e-3;
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s + mrow[warp][lane & 7];
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[slot] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 256); cudaMemset(out, 0, 1 << 22);
  const int iters = 4096;
  const char *names[3] = {"full", "weights in registers", "no shuffle"};
  for (int warps = 1; warps <= 16; warps *= 2) {
    k<0><<<148, 32 * warps>>>(out, iters, cyc, 0); k<1><<<148, 32 * warps>>>(out, iters, cyc, 1); k<2><<<148, 32 * warps>>>(out, iters, cyc, 2);
    cudaDeviceSynchronize();
    long long h[3]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 3; m++) printf("warps/SM %2d  %-22s %.1f cycles per column\n", warps, names[m], (double)h[m] / (iters * 8.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
