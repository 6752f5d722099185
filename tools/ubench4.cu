// per-SM throughput of candidate max-reduction idioms: cycles per edge (one FADD + max) per scheduler
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = -(threadIdx.x * 0.5f + i);
  float x = out[0] - 1.f, y = out[1] - 2.f;
  unsigned acc0 = 0xff800000u, acc1 = 0xff800000u;
  float f0 = -1e30f, f1 = -1e30f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float c0 = a[i] + x, c1 = a[i + 1] + y;      // two candidates
      if (MODE == 0) { f0 = fmaxf(fmaxf(f0, c0), c1); }                                         // FMNMX3
      if (MODE == 1) { f0 = fmaxf(f0, c0); f1 = fmaxf(f1, c1); }                                // 2 x FMNMX
      if (MODE == 2) { acc0 = __vimin3_u32(acc0, __float_as_uint(c0), __float_as_uint(c1)); }   // VIMNMX3.U32
      if (MODE == 3) { acc0 = min(acc0, __float_as_uint(c0)); acc1 = min(acc1, __float_as_uint(c1)); }  // 2 x VIMNMX
      if (MODE == 4) { acc0 = __vimin3_u32(acc0, __float_as_uint(c0), __float_as_uint(c1)); f0 = fmaxf(f0, a[i] + y); } // mixed pipes
    }
    x -= 1e-3f; y -= 1e-3f;
  }
  long long t1 = clock64();
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + __uint_as_float(acc0) + __uint_as_float(acc1);
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 22);
  const int iters = 4096;
  const char *names[5] = {"2 FADD + FMNMX3", "2 FADD + 2 FMNMX", "2 FADD + VIMNMX3.U32", "2 FADD + 2 VIMNMX.U32", "3 FADD + VIMNMX3 + FMNMX"};
  for (int warps = 4; warps <= 16; warps *= 2) {
    k<0><<<148, 32 * warps>>>(out, iters, cyc); k<1><<<148, 32 * warps>>>(out, iters, cyc); k<2><<<148, 32 * warps>>>(out, iters, cyc);
    k<3><<<148, 32 * warps>>>(out, iters, cyc); k<4><<<148, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 5; m++)
      printf("warps/SM %2d  %-26s %.2f scheduler-cycles per candidate pair\n", warps, names[m], (double)h[m] / (iters * 8.0) / (warps / 4.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
