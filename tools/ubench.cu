// ubench.cu -- per-SM issue rates that drive the DP kernel design (run on the B200 box).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void k(float *out, int iters, long long *cyc) {
  float acc[8][4];
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc[i][j] = threadIdx.x * 1e-3f + i;
  uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b[2] = {0x3f803f80u, 0x3f803f80u};
  float x = threadIdx.x * 1e-6f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) mma_bf16(acc[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; i++) mma_tf32(acc[i], a, b);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(acc[i][j], 1.0001f, x);
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = exp2f(acc[i][j] * 0.999f);
    } else if (MODE == 4) {
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(__shfl_sync(0xffffffffu, acc[i][j], (i * 4 + j) & 31), 1.0001f, x);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int threads, double ops_per_iter_per_thread_or_warp, bool per_warp) {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  int iters = 20000;
  k<MODE><<<148, threads>>>(out, 100, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  double units = per_warp ? threads / 32.0 : threads;
  double ops = ops_per_iter_per_thread_or_warp * units * iters;  // per SM
  printf("%-28s threads=%4d  %.3f ms  cycles=%lld  per-SM ops/clk=%.1f  chip=%.3e ops/s\n", name, threads, ms, c,
         ops / (double)c, ops * 148 / (ms * 1e-3));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int th : {128, 256, 512, 1024}) {
    run<0>("mma.sync bf16 m16n8k16 (MAC)", th, 8.0 * 2048, true);
    run<1>("mma.sync tf32 m16n8k8 (MAC)", th, 8.0 * 1024, true);
    run<2>("FFMA (FMA)", th, 32.0, false);
    run<3>("FMUL+MUFU.EX2 (exp)", th, 32.0, false);
    run<4>("SHFL+FFMA (pairs)", th, 32.0, false);
  }
  return 0;
}
