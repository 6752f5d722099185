// tcgen05_probe.cu -- stand-alone check of the tcgen05.mma path planned for the far-predecessor GEMM (round 2):
// D[128 x 32] (TMEM, fp32) = A[128 x K] (smem, bf16, K-major) * B[32 x K]^T (smem, bf16, K-major), no swizzle,
// canonical core-matrix layout (8 rows x 16 bytes), one thread issues, completion through tcgen05.commit -> mbarrier,
// epilogue = tcgen05.ld 32x32b (thread = row).  Verifies numerics against the host and times the issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tcgen05_probe tools/tcgen05_probe.cu && /tmp/tcgen05_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

constexpr int kM = 128, kK = 64;                   // K = 64 -> 4 MMAs of K = 16; N is a template parameter

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: start address, LBO (distance between the two 8-element K
// core matrices of one MMA), SBO (distance between 8-row groups), all in units of 16 bytes; version = 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                          // version_ = 1
  return d;                                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M = 128, N = 32
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                 // c_format = F32
  d |= 1u << 7;                 // a_format = BF16
  d |= 1u << 10;                // b_format = BF16
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
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

template <int kN>
__global__ void __launch_bounds__(128, 1)
probe(const __nv_bfloat16 *__restrict__ A, const __nv_bfloat16 *__restrict__ B, float *__restrict__ D, int reps,
      long long *cyc) {
  // canonical layout: element (row r, k) at ((k / 8) * rows + r) * 16 bytes + (k % 8) * 2
  __shared__ __align__(128) __nv_bfloat16 sa[(kK / 8) * kM * 8];
  __shared__ __align__(128) __nv_bfloat16 sb[(kK / 8) * kN * 8];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int x = tid; x < kM * kK; x += 128) {
    const int r = x / kK, k = x % kK;
    sa[((k / 8) * kM + r) * 8 + (k % 8)] = A[r * kK + k];
  }
  for (int x = tid; x < kN * kK; x += 128) {
    const int n = x / kK, k = x % kK;
    sb[((k / 8) * kN + n) * 8 + (k % 8)] = B[n * kK + k];
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  // make the generic-proxy writes of the operands visible to the async (tensor-core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc(kM, kN);
  uint32_t phase = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; rep++) {
    if (tid == 0) {
#pragma unroll
      for (int ks = 0; ks < kK / 16; ks++) {
        const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * kM * 16, kM * 16, 128);
        const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * kN * 16, kN * 16, 128);
        mma_f16_ss(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  long long t1 = clock64();
  // issue throughput: 96 MMAs back to back, one commit
  long long t2 = 0, t3 = 0;
  if (reps > 1) {
    t2 = clock64();
    if (tid == 0) {
      for (int it = 0; it < 24; it++) {
#pragma unroll
        for (int ks = 0; ks < kK / 16; ks++) {
          const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * kM * 16, kM * 16, 128);
          const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * kN * 16, kN * 16, 128);
          mma_f16_ss(tmem, ad, bd, idesc, 1u);
        }
      }
      t3 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) { cyc[1] = t3 - t2; cyc[2] = clock64() - t2; }
  }
  // issuer-style loop (as in dag_dp4.cu): the whole warp 0 iterates; per item one elected lane issues 6 MMAs + ONE commit,
  // the warp reconverges; commits go to a ring of 8 mbarriers that nobody waits on except every 8th item
  __shared__ __align__(8) uint64_t ring[8];
  if (reps > 1) {
    if (tid < 8) mbar_init(&ring[tid], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      const long long t4 = clock64();
      for (int it = 0; it < 256; it++) {
        const int sl = it & 7;
        if (it >= 8) mbar_wait(&ring[sl], ((it >> 3) - 1) & 1);      // slot reuse: item it - 8 complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
        if (pred) {
#pragma unroll
          for (int m = 0; m < 6; m++) {
            const int ks = m & 1;
            const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * kM * 16, kM * 16, 128);
            const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * kN * 16, kN * 16, 128);
            mma_f16_ss(tmem + (uint32_t)(sl & 3) * 32, ad, bd, idesc, m > 0 ? 1u : 0u);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ring[sl])) : "memory");
        }
        __syncwarp();
      }
      const long long t5 = clock64();
      for (int sl = 0; sl < 8; sl++) mbar_wait(&ring[sl], 1);         // 256 items: 32 per slot -> last phase parity 1
      if (tid == 0) { cyc[3] = t5 - t4; }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // epilogue: thread = row (warp w reads TMEM lanes 32w .. 32w+31), 32 columns
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
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
  for (int j = 0; j < 32; j++) D[tid * kN + j] = __uint_as_float(v[j]);
  if (tid == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int kN>
int run() {
  const int nA = kM * kK, nB = kN * kK;
  __nv_bfloat16 *hA = new __nv_bfloat16[nA], *hB = new __nv_bfloat16[nB];
  float *fA = new float[nA], *fB = new float[nB];
  srand(1);
  for (int i = 0; i < nA; i++) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < nB; i++) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float *dD; long long *dc;
  cudaMalloc(&dA, nA * 2); cudaMalloc(&dB, nB * 2); cudaMalloc(&dD, kM * kN * 4); cudaMalloc(&dc, 64);
  cudaMemcpy(dA, hA, nA * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, nB * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, kM * kN * 4);
  probe<kN><<<1, 128>>>(dA, dB, dD, 1, dc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("N = %d launch: %s\n", kN, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  float *hD = new float[kM * kN];
  cudaMemcpy(hD, dD, kM * kN * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < kM; r++)
    for (int n = 0; n < 32; n++) {     // the epilogue of the probe reads back the first 32 columns
      double s = 0;
      for (int k = 0; k < kK; k++) s += (double)fA[r * kK + k] * fB[n * kK + k];
      maxerr = fmax(maxerr, fabs(s - hD[r * kN + n])); maxref = fmax(maxref, fabs(s));
    }
  printf("  max |err| = %.3e (max |ref| = %.3f) -> %s\n", maxerr, maxref, maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  for (int reps : {1000}) {
    probe<kN><<<1, 128>>>(dA, dB, dD, reps, dc);
    cudaDeviceSynchronize();
    long long c[4]; cudaMemcpy(c, dc, 32, cudaMemcpyDeviceToHost);
    printf("  reps %d: %.1f cycles per (4 x MMA M128 N%d K16 + commit + wait); 96 MMAs back to back: issue %.1f cycles each, %.1f cycles each until complete\n",
           reps, (double)c[0] / reps, kN, c[1] / 96.0, c[2] / 96.0);
    printf("  issuer-style loop (warp-wide, elect, 6 MMAs + 1 commit per item, 8-slot ring): %.1f cycles per item\n", c[3] / 256.0);
  }
  return 0;
}

int main() {
  int rc = 0;
  rc |= run<32>();
  rc |= run<64>();
  rc |= run<128>();
  return rc;
}
