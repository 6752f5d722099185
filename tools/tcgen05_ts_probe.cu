// tcgen05_ts_probe.cu -- stand-alone check of the A-from-TMEM form planned for the transition-gradient contraction:
// D[128 x 32] (TMEM, fp32) += A[128 x 16] (TMEM, bf16 pairs, lane = row, 8 columns) * B[32 x 16]^T (smem, bf16, K-major,
// no swizzle).  The A operand is written by the threads themselves (tcgen05.st.32x32b.x8, thread = row).  Verifies the
// numerics against the host and times the issue rate of back-to-back MMAs on different accumulator columns.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
constexpr int kM = 128, kN = 32, kK = 16, kNB = 8;      // 8 independent N-blocks (different B tiles, different D columns)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__global__ void __launch_bounds__(128, 1)
probe(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ B, float *__restrict__ D, long long *cyc) {
  __shared__ __align__(128) __nv_bfloat16 sb[kNB][(kK / 8) * kN * 8];     // per N-block: [k-core][n][8 along K]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int x = tid; x < kNB * kN * kK; x += 128) {
    const int nb = x / (kN * kK), n = (x / kK) % kN, k = x % kK;
    sb[nb][((k / 8) * kN + n) * 8 + (k % 8)] = B[x];
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t tmem_a = tmem + 256;             // A operand: columns 256 .. 263
  // my row of A: 16 bf16 = 8 packed words (element k in the low half of word k / 2 when k is even)
  uint32_t w[8];
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const __nv_bfloat162 p = __halves2bfloat162(A[tid * kK + 2 * q], A[tid * kK + 2 * q + 1]);
    w[q] = *reinterpret_cast<const uint32_t *>(&p);
  }
  const uint32_t ta = tmem_a + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(ta), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int rep = 0; rep < 12; rep++)
#pragma unroll
      for (int nb = 0; nb < kNB; nb++) {
        const uint64_t bd = make_desc(smem_u32(sb[nb]), kN * 16, 128);
        mma_f16_ts(tmem + nb * 32, tmem_a, bd, rep > 0 ? 1u : 0u);
      }
    t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0);
  if (tid == 0) { t2 = clock64(); cyc[0] = t1 - t0; cyc[1] = t2 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int nb = 0; nb < kNB; nb++) {
    uint32_t v[32];
    const uint32_t taddr = tmem + nb * 32 + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; j++) D[(nb * kM + tid) * kN + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  const int nA = kM * kK, nB = kNB * kN * kK;
  __nv_bfloat16 *hA = new __nv_bfloat16[nA], *hB = new __nv_bfloat16[nB];
  float *fA = new float[nA], *fB = new float[nB];
  srand(1);
  for (int i = 0; i < nA; i++) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < nB; i++) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float *dD; long long *dc;
  cudaMalloc(&dA, nA * 2); cudaMalloc(&dB, nB * 2); cudaMalloc(&dD, kNB * kM * kN * 4); cudaMalloc(&dc, 64);
  cudaMemcpy(dA, hA, nA * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, nB * 2, cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dA, dB, dD, dc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  float *hD = new float[kNB * kM * kN];
  cudaMemcpy(hD, dD, kNB * kM * kN * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int nb = 0; nb < kNB; nb++)
    for (int r = 0; r < kM; r++)
      for (int n = 0; n < kN; n++) {
        double s = 0;
        for (int k = 0; k < kK; k++) s += (double)fA[r * kK + k] * fB[(nb * kN + n) * kK + k];
        s *= 12;
        maxerr = fmax(maxerr, fabs(s - hD[(nb * kM + r) * kN + n])); maxref = fmax(maxref, fabs(s));
      }
  printf("max |err| = %.3e (max |ref| = %.3f) -> %s\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  long long c[2]; cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
  printf("96 MMAs (A from TMEM, M128 N32 K16): issue %.1f cycles each, %.1f cycles each until complete\n", c[0] / 96.0, c[1] / 96.0);
  return 0;
}
