// ubench2.cu -- fp64 issue rate / dependent latencies (decides whether the in-block chain can run in fp64).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench2 tools/ubench2.cu && gpurun_out/ubench2
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(double *out, int iters, long long *cyc) {
  double acc[16];
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
  double x = threadIdx.x * 1e-6;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {          // throughput: 16 independent DFMA
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], 1.0001, x);
    } else if (MODE == 1) {   // latency: 16 dependent DFMA
#pragma unroll
      for (int i = 0; i < 16; i++) acc[0] = fma(acc[0], 1.0001, x);
    } else if (MODE == 2) {   // dependent: shfl(64-bit) + DADD
#pragma unroll
      for (int i = 0; i < 16; i++) acc[0] = __shfl_up_sync(0xffffffffu, acc[0], 1) + x;
    } else if (MODE == 3) {   // dependent chain of one column step: shfl64 + DADD + DMUL + DFMA
#pragma unroll
      for (int i = 0; i < 16; i++) {
        double rm = __shfl_up_sync(0xffffffffu, acc[0], 1);
        double tot = rm + acc[1];
        double m = tot * 0.999;
        acc[0] = fma(m, 0.5, acc[2]);
      }
    } else if (MODE == 4) {   // dependent FFMA latency
      float f = (float)acc[0];
#pragma unroll
      for (int i = 0; i < 16; i++) f = fmaf(f, 1.0001f, (float)x);
      acc[0] = f;
    } else if (MODE == 5) {   // dependent: shfl32 + FADD
      float f = (float)acc[0];
#pragma unroll
      for (int i = 0; i < 16; i++) f = __shfl_up_sync(0xffffffffu, f, 1) + (float)x;
      acc[0] = f;
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int threads, double ops_per_iter) {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  int iters = 5000;
  k<MODE><<<148, threads>>>(out, 100, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s threads=%4d  %.3f ms  cycles/iter=%.1f  per-SM thread-ops/clk=%.2f\n", name, threads, ms,
         (double)c / iters, ops_per_iter * threads * iters / (double)c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {32, 128, 256, 512}) {
    run<0>("DFMA x16 independent", th, 16);
    run<1>("DFMA x16 dependent", th, 16);
    run<2>("SHFL64+DADD x16 dependent", th, 16);
    run<3>("SHFL64+DADD+DMUL+DFMA x16 dependent", th, 16);
    run<4>("FFMA x16 dependent", th, 16);
    run<5>("SHFL32+FADD x16 dependent", th, 16);
  }
  return 0;
}
