// FP64 throughput of one SM on B200: cycles per warp-level DFMA / DMUL / DADD with 1..16 warps resident,
// 8 independent accumulators per thread (latency hidden within a warp); and the dependent-chain latency.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, int iters, long long *cyc) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = 1.0 + threadIdx.x * 1e-3 + i;
  const double x = out[0] + 1.0000001, y = out[1] + 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = fma(a[i], x, y);
      if (MODE == 1) a[i] = a[i] * x;
      if (MODE == 2) a[i] = a[i] + y;
      if (MODE == 3) a[0] = fma(a[0], x, y);     // dependent chain
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 22);
  const int iters = 2048;
  const char *names[4] = {"DFMA", "DMUL", "DADD", "DFMA dependent"};
  for (int warps = 1; warps <= 16; warps *= 2) {
    k<0><<<148, 32 * warps>>>(out, iters, cyc); k<1><<<148, 32 * warps>>>(out, iters, cyc);
    k<2><<<148, 32 * warps>>>(out, iters, cyc); k<3><<<148, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 4; m++)
      printf("warps/SM %2d  %-16s %.2f cycles per warp instruction (per warp)  -> %.2f warp-instr/clk/SM\n", warps, names[m],
             (double)h[m] / (iters * 8.0), warps * iters * 8.0 / (double)h[m]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
