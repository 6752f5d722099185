// throughput of FADD / FMNMX / FMNMX3 / LDS.128-broadcast per SM (one CTA per SM, W warps), cycles per warp-instruction
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
  __shared__ float4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 0.5f + i;
  float x = out[0], y = out[1];
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (MODE == 0) a[i] = a[i] + x;
      if (MODE == 1) a[i] = fmaxf(a[i], x);
      if (MODE == 2) a[i] = fmaxf(fmaxf(a[i], x), y);
      if (MODE == 3) { float4 v = sm[(it * 16 + i) & 1023]; a[i] += v.x + v.y; }   // broadcast LDS.128 + 2 FADD
      if (MODE == 4) { a[i] = fmaxf(fmaxf(a[i], a[(i + 1) & 15] + x), a[(i + 2) & 15] + y); }  // 2 FADD + FMNMX3
    }
    x += 1e-30f; y -= 1e-30f;
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[2 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
}
int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 64); cudaMemset(out, 0, 1 << 22);
  const int iters = 4096;
  const char *names[5] = {"FADD", "FMNMX", "FMNMX3", "LDS.128 bcast + 2 FADD", "2 FADD + FMNMX3"};
  for (int warps = 1; warps <= 16; warps *= 2) {
    k<0><<<148, 32 * warps>>>(out, iters, cyc); k<1><<<148, 32 * warps>>>(out, iters, cyc); k<2><<<148, 32 * warps>>>(out, iters, cyc);
    k<3><<<148, 32 * warps>>>(out, iters, cyc); k<4><<<148, 32 * warps>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    for (int m = 0; m < 5; m++)
      printf("warps/SM %2d  %-24s %.2f cycles per loop-op per warp, %.2f SM-cycles per warp-op\n", warps, names[m], (double)h[m] / (iters * 16.0), (double)h[m] / (iters * 16.0) / warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
