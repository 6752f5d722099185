// lsg.cu -- fused vocabulary log-softmax + gather, forward and backward, for sm_100a.
//
// Replaces logsoftmax_gather_kernel (reference logsoftmax_gather.cu:256-310, three scalar-load passes
// over every row) and the torch mul_/scatter_add_ pair of DagLogsoftmaxGatherFunc.backward
// (reference dag_loss.py:293-295).
//
// Forward, register-resident path: one 256-thread CTA per vocabulary row.  The row is read ONCE with
// 128-bit loads into registers (NV x 16 B per thread), reduced (max, then sum of exp) with warp
// shuffles + one shared-memory hop, the S target logits are gathered from L2 before anything is
// overwritten, and -- when a gradient is required -- the probabilities are written back in place
// with 128-bit stores.  HBM traffic = one read + one write of the logits plane + the gathered plane,
// which is the algorithmic minimum.
//
// Rows that do not fit the register path (V not a multiple of the vector width, huge V, fp64) take a
// streaming three-pass kernel whose 2nd/3rd pass hit L2.
#include <cstdlib>

#include "common.cuh"

namespace dagb200 {

constexpr int kLsgThreads = 256;

template <typename T> struct VecTraits;
template <> struct VecTraits<float> {
  static constexpr int kElems = 4;
  __device__ static __forceinline__ void unpack(const uint4 &u, float *f) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ uint4 pack(const float *f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  __device__ static __forceinline__ float to_float(float v) { return v; }
  __device__ static __forceinline__ float from_float(float v) { return v; }
  __device__ static __forceinline__ uint4 scale(const uint4 &u, float sc) {
    return make_uint4(__float_as_uint(__uint_as_float(u.x) * sc), __float_as_uint(__uint_as_float(u.y) * sc),
                      __float_as_uint(__uint_as_float(u.z) * sc), __float_as_uint(__uint_as_float(u.w) * sc));
  }
};
template <> struct VecTraits<__half> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void unpack(const uint4 &u, float *f) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 p = __half22float2(h[i]); f[2 * i] = p.x; f[2 * i + 1] = p.y; }
  }
  __device__ static __forceinline__ uint4 pack(const float *f) {
    uint4 u; __half2 *h = reinterpret_cast<__half2 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    return u;
  }
  __device__ static __forceinline__ float to_float(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
  // element-wise product with a scalar IN the 16-bit type (one rounding per element, as a torch half multiply)
  __device__ static __forceinline__ uint4 scale(const uint4 &u, float sc) {
    const __half2 s2 = __float2half2_rn(sc);
    uint4 r;
    const __half2 *a = reinterpret_cast<const __half2 *>(&u);
    __half2 *o = reinterpret_cast<__half2 *>(&r);
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = __hmul2(a[i], s2);
    return r;
  }
};
template <> struct VecTraits<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static __forceinline__ void unpack(const uint4 &u, float *f) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 p = __bfloat1622float2(h[i]); f[2 * i] = p.x; f[2 * i + 1] = p.y; }
  }
  __device__ static __forceinline__ uint4 pack(const float *f) {
    uint4 u; __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
  }
  __device__ static __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ uint4 scale(const uint4 &u, float sc) {
    const __nv_bfloat162 s2 = __float2bfloat162_rn(sc);
    uint4 r;
    const __nv_bfloat162 *a = reinterpret_cast<const __nv_bfloat162 *>(&u);
    __nv_bfloat162 *o = reinterpret_cast<__nv_bfloat162 *>(&r);
#pragma unroll
    for (int i = 0; i < 4; i++) o[i] = __hmul2(a[i], s2);
    return r;
  }
};

// block-wide all-reduce over 256 threads (8 warps); `red` is 8 floats of shared memory per use
template <bool IS_MAX> __device__ __forceinline__ float block_allreduce(float v, float *red) {
  v = IS_MAX ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = red[l & 7];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    float x = __shfl_xor_sync(0xffffffffu, r, o);
    r = IS_MAX ? fmaxf(r, x) : (r + x);
  }
  return r;
}

// block-wide arg-max over 256 threads: larger value wins, then the smaller index (first occurrence, as torch.argmax on
// CUDA returns); `redv` / `redi` are 8 entries of shared memory each
__device__ __forceinline__ void block_allreduce_argmax(float &v, int &i, float *redv, int *redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { redv[w] = v; redi[w] = i; }
  __syncthreads();
  v = redv[l & 7];
  i = redi[l & 7];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}

// arg-max over the vocabulary for rows the TMA-staged kernel does not take (any V / alignment / dtype): one CTA per row
template <typename T>
__global__ void __launch_bounds__(kLsgThreads)
lsg_argmax_kernel(const T *__restrict__ logits, int64_t *__restrict__ amax, int V) {
  __shared__ float redv[8];
  __shared__ int redi[8];
  const T *x = logits + (int64_t)blockIdx.x * V;
  float v = neg_inf_f();
  int i = 0x7fffffff;
  for (int c = threadIdx.x; c < V; c += kLsgThreads) {
    const float f = (float)x[c];
    if (f > v || i == 0x7fffffff) { v = f; i = c; }
  }
  block_allreduce_argmax(v, i, redv, redi);
  if (threadIdx.x == 0) amax[blockIdx.x] = i;
}

// ---------------------------------------------------------------------------------------------
// forward, register-resident: NV 16-byte vectors per thread (V <= NV * 256 * kElems)
template <typename T, int NV, bool GRAD>
__global__ void __launch_bounds__(kLsgThreads)
lsg_fwd_reg_kernel(T *__restrict__ logits, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
                   float *__restrict__ out, int64_t osb, int64_t osl, int64_t oss, int L, int V, int S) {
  using VT = VecTraits<T>;
  constexpr int E = VT::kElems;
  extern __shared__ float smem[];  // [S] gathered raw logits, then 16 floats of reduction scratch
  float *sel = smem;
  float *red = smem + S;

  const int64_t row = blockIdx.x;
  const int b = (int)(row / L), l = (int)(row % L);
  T *x = logits + row * (int64_t)V;
  const int nvec = V / E;

  float v[NV][E];
  float tmax = neg_inf_f();
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const int c = k * kLsgThreads + threadIdx.x;
    if (c < nvec) {
      uint4 u = __ldcs(reinterpret_cast<const uint4 *>(x) + c);  // streaming: read once
      VT::unpack(u, v[k]);
#pragma unroll
      for (int e = 0; e < E; e++) tmax = fmaxf(tmax, v[k][e]);
    } else {
#pragma unroll
      for (int e = 0; e < E; e++) v[k][e] = neg_inf_f();
    }
  }
  // gather the raw target logits before any in-place store (barriers below order them)
  const int64_t *ib = idx + b * isb + l * isl;
  for (int s = threadIdx.x; s < S; s += kLsgThreads) sel[s] = VT::to_float(x[ib[s * iss]]);

  const float rmax = block_allreduce<true>(tmax, red);
  float tsum = 0.f;
  const float moff = (rmax == neg_inf_f()) ? 0.f : rmax;
#pragma unroll
  for (int k = 0; k < NV; k++)
#pragma unroll
    for (int e = 0; e < E; e++) { v[k][e] = __expf(v[k][e] - moff); tsum += v[k][e]; }
  const float rsum = block_allreduce<false>(tsum, red + 8);
  const float lsum = __logf(rsum);

  float *ob = out + b * osb + l * osl;
  for (int s = threadIdx.x; s < S; s += kLsgThreads) ob[s * oss] = (sel[s] - rmax) - lsum;

  if (GRAD) {
    const float inv = __fdividef(1.f, rsum);
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = k * kLsgThreads + threadIdx.x;
      if (c < nvec) {
#pragma unroll
        for (int e = 0; e < E; e++) v[k][e] *= inv;
        reinterpret_cast<uint4 *>(x)[c] = VT::pack(v[k]);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// forward, TMA-staged: one CTA walks kLsgGroup consecutive vertices (rows) of one utterance.  Rows are bulk-copied
// (cp.async.bulk, one 16-byte-aligned row per copy) into a ring of kLsgStages shared-memory stages by one thread,
// so the next rows are in flight while the current one is reduced -- no registers are tied up by loads in flight.
// The S target logits are gathered from the staged row (no second trip to L2), and the results of the group are
// written together: with the transposed [B,S,L] result layout that is one 32-byte segment per target instead of
// eight scattered 4-byte stores.
constexpr int kLsgGroup = 8;
constexpr int kLsgStages = 3;

__device__ __forceinline__ uint32_t lsg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lsg_mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(lsg_smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void lsg_load_row(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lsg_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(lsg_smem_u32(dst)), "l"(src), "r"(bytes), "r"(lsg_smem_u32(bar)) : "memory");
}

template <typename T, int NV, bool GRAD, bool AMAX>
__global__ void __launch_bounds__(kLsgThreads)
lsg_fwd_tma_kernel(T *__restrict__ logits, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
                   float *__restrict__ out, int64_t osb, int64_t osl, int64_t oss, int L, int V, int S, int groups,
                   int64_t *__restrict__ amax) {
  using VT = VecTraits<T>;
  constexpr int E = VT::kElems;
  extern __shared__ __align__(128) unsigned char lsg_smem[];
  const uint32_t rowbytes = (uint32_t)V * sizeof(T);
  unsigned char *stages = lsg_smem;                                            // [kLsgStages][rowbytes]
  float *outb = reinterpret_cast<float *>(lsg_smem + (size_t)kLsgStages * rowbytes);   // [S][kLsgGroup + 1]
  int *idxs = reinterpret_cast<int *>(outb + (size_t)S * (kLsgGroup + 1));     // [S]
  float *red = reinterpret_cast<float *>(idxs + S);                            // [16]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + 16);                     // [kLsgStages]
  int *redi = reinterpret_cast<int *>(full + kLsgStages);                      // [8] (arg-max variant)

  const int b = blockIdx.x / groups, l0 = (blockIdx.x % groups) * kLsgGroup;
  const int nrows = min(kLsgGroup, L - l0);
  T *xg = logits + ((int64_t)b * L + l0) * (int64_t)V;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kLsgStages; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsg_smem_u32(full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int r = 0; r < min(kLsgStages, nrows); r++)
      lsg_load_row(stages + (size_t)r * rowbytes, xg + (int64_t)r * V, rowbytes, full + r);
  }
  const bool shared_idx = (isl == 0);
  if (shared_idx) {
    const int64_t *ib = idx + b * isb;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) idxs[s] = (int)ib[s * iss];
  }
  __syncthreads();
  const int nvec = V / E;

  for (int r = 0; r < nrows; r++) {
    const int st = r % kLsgStages;
    const T *xs = reinterpret_cast<const T *>(stages + (size_t)st * rowbytes);
    if (!shared_idx) {
      const int64_t *ib = idx + b * isb + (l0 + r) * isl;
      for (int s = threadIdx.x; s < S; s += kLsgThreads) idxs[s] = (int)ib[s * iss];
      __syncthreads();
    }
    lsg_mbar_wait(full + st, (r / kLsgStages) & 1);
    float v[NV][E];
    float tmax = neg_inf_f();
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = k * kLsgThreads + threadIdx.x;
      if (c < nvec) {
        const uint4 u = reinterpret_cast<const uint4 *>(xs)[c];
        VT::unpack(u, v[k]);
#pragma unroll
        for (int e = 0; e < E; e++) tmax = fmaxf(tmax, v[k][e]);
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) v[k][e] = neg_inf_f();
      }
    }
    // raw target logits from the staged row (outb doubles as their parking place)
    for (int s = threadIdx.x; s < S; s += kLsgThreads) outb[s * (kLsgGroup + 1) + r] = VT::to_float(xs[idxs[s]]);
    const float rmax = block_allreduce<true>(tmax, red);
    if (AMAX) {
      // first index holding the row maximum: the smallest one among my elements, then a warp-wide integer minimum;
      // the 8 warp results meet after the barrier of the sum reduction below
      int tidx = 0x7fffffff;
#pragma unroll
      for (int k = NV - 1; k >= 0; k--) {
        const int c = k * kLsgThreads + threadIdx.x;
#pragma unroll
        for (int e = E - 1; e >= 0; e--)
          if (c < nvec && v[k][e] == rmax) tidx = c * E + e;
      }
      tidx = __reduce_min_sync(0xffffffffu, tidx);
      if ((threadIdx.x & 31) == 0) redi[threadIdx.x >> 5] = tidx;
    }
    // every thread is past its reads of this stage: refill it with the row kLsgStages ahead
    if (threadIdx.x == 0 && r + kLsgStages < nrows) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      lsg_load_row(stages + (size_t)st * rowbytes, xg + (int64_t)(r + kLsgStages) * V, rowbytes, full + st);
    }
    float tsum = 0.f;
    const float moff = (rmax == neg_inf_f()) ? 0.f : rmax;
    // exp(x - m) = 2^(x log2e - m log2e): one FFMA + one MUFU.EX2 per element (the kernel is issue-bound for 16-bit rows:
    // ncu showed 82 % issue utilisation with __expf's subtract, multiply, range fix-up and MUFU)
    const float mo2 = moff * kLog2e;
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
      for (int e = 0; e < E; e++) {
        float p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(fmaf(v[k][e], kLog2e, -mo2)));
        v[k][e] = p;
        tsum += p;
      }
    const float rsum = block_allreduce<false>(tsum, red + 8);
    if (AMAX && threadIdx.x == 0) {
      int best = redi[0];
#pragma unroll
      for (int w = 1; w < kLsgThreads / 32; w++) best = min(best, redi[w]);
      amax[(int64_t)b * L + l0 + r] = best;
    }
    const float lsum = __logf(rsum);
    for (int s = threadIdx.x; s < S; s += kLsgThreads) {
      float *o = outb + s * (kLsgGroup + 1) + r;
      *o = (*o - rmax) - lsum;
    }
    if (GRAD) {
      const float inv = __fdividef(1.f, rsum);
      T *x = xg + (int64_t)r * V;
#pragma unroll
      for (int k = 0; k < NV; k++) {
        const int c = k * kLsgThreads + threadIdx.x;
        if (c < nvec) {
#pragma unroll
          for (int e = 0; e < E; e++) v[k][e] *= inv;
          reinterpret_cast<uint4 *>(x)[c] = VT::pack(v[k]);
        }
      }
    }
  }
  __syncthreads();
  // results of the group
  float *ob = out + b * osb + (int64_t)l0 * osl;
  const bool vec = (osl == 1) && nrows == kLsgGroup && ((oss & 3) == 0) && ((osb & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (vec) {
    for (int s = threadIdx.x; s < S; s += kLsgThreads) {
      const float *o = outb + s * (kLsgGroup + 1);
      float4 *dst = reinterpret_cast<float4 *>(ob + s * oss);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  } else if (oss == 1) {
    for (int r = 0; r < nrows; r++)
      for (int s = threadIdx.x; s < S; s += kLsgThreads) ob[r * osl + s] = outb[s * (kLsgGroup + 1) + r];
  } else {
    for (int x = threadIdx.x; x < S * nrows; x += kLsgThreads) {
      const int s = x / nrows, r = x % nrows;
      ob[r * osl + s * oss] = outb[s * (kLsgGroup + 1) + r];
    }
  }
}

static size_t lsg_fwd_tma_smem(size_t rowbytes, int S) {
  return (size_t)kLsgStages * rowbytes + (size_t)S * (kLsgGroup + 1) * 4 + (size_t)S * 4 + 16 * 4 + kLsgStages * 8 + 8 * 4 + 16;
}

template <typename T, int NV>
static int launch_fwd_tma(T *logits, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss, float *out,
                          int64_t osb, int64_t osl, int64_t oss, int B, int L, int V, int S, bool grad,
                          cudaStream_t st, int64_t *amax) {
  const size_t smem = lsg_fwd_tma_smem((size_t)V * sizeof(T), S);
  const int groups = (L + kLsgGroup - 1) / kLsgGroup;
  const unsigned grid = (unsigned)((int64_t)B * groups);
#define DAGB200_LSG_LAUNCH(G, A)                                                                                      \
  do {                                                                                                                \
    cudaFuncSetAttribute(lsg_fwd_tma_kernel<T, NV, G, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   \
    lsg_fwd_tma_kernel<T, NV, G, A><<<grid, kLsgThreads, smem, st>>>(logits, idx, isb, isl, iss, out, osb, osl, oss,  \
                                                                     L, V, S, groups, amax);                          \
  } while (0)
  if (grad) { if (amax) DAGB200_LSG_LAUNCH(true, true); else DAGB200_LSG_LAUNCH(true, false); }
  else { if (amax) DAGB200_LSG_LAUNCH(false, true); else DAGB200_LSG_LAUNCH(false, false); }
#undef DAGB200_LSG_LAUNCH
  DAGB200_CHECK_LAUNCH("lsg_fwd_tma_kernel");
  return 0;
}

// forward, streaming fallback (any V / alignment); T2 = compute type
template <typename T, typename C, bool GRAD>
__global__ void __launch_bounds__(kLsgThreads)
lsg_fwd_stream_kernel(T *__restrict__ logits, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
                      C *__restrict__ out, int64_t osb, int64_t osl, int64_t oss, int L, int V, int S, int64_t rows) {
  __shared__ C red[kLsgThreads / 32];
  __shared__ C bcast[2];
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / L), l = (int)(row % L);
    T *x = logits + row * (int64_t)V;
    C tmax = neg_inf<C>();
    for (int c = threadIdx.x; c < V; c += kLsgThreads) tmax = fmax(tmax, (C)x[c]);
    tmax = warp_max(tmax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tmax;
    __syncthreads();
    if (threadIdx.x == 0) { C m = red[0]; for (int w = 1; w < kLsgThreads / 32; w++) m = fmax(m, red[w]); bcast[0] = m; }
    __syncthreads();
    const C rmax = bcast[0];
    const C moff = isinf(rmax) ? (C)0 : rmax;
    C tsum = 0;
    for (int c = threadIdx.x; c < V; c += kLsgThreads) tsum += acc_exp((C)x[c] - moff);
    tsum = warp_sum(tsum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tsum;
    __syncthreads();
    if (threadIdx.x == 0) { C s = 0; for (int w = 0; w < kLsgThreads / 32; w++) s += red[w]; bcast[1] = s; }
    __syncthreads();
    const C rsum = bcast[1];
    const C lsum = acc_log(rsum);
    const int64_t *ib = idx + b * isb + l * isl;
    C *ob = out + b * osb + l * osl;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) ob[s * oss] = ((C)x[ib[s * iss]] - rmax) - lsum;
    if (GRAD) {
      __syncthreads();  // all gathers done before the row is overwritten
      for (int c = threadIdx.x; c < V; c += kLsgThreads) x[c] = (T)(acc_exp((C)x[c] - moff) / rsum);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// backward: row staged in shared memory as fp32, scatter via shared-memory atomics, one write.
template <typename T, typename C>
__global__ void __launch_bounds__(kLsgThreads)
lsg_bwd_kernel(T *__restrict__ probs, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
               const C *__restrict__ gout, int64_t gsb, int64_t gsl, int64_t gss, int L, int V, int S, int64_t rows) {
  extern __shared__ unsigned char smem_raw[];
  C *rowbuf = reinterpret_cast<C *>(smem_raw);  // [V]
  __shared__ C red[kLsgThreads / 32];
  __shared__ C bcast;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / L), l = (int)(row % L);
    T *x = probs + row * (int64_t)V;
    const C *gb = gout + b * gsb + l * gsl;
    C part = 0;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) part += gb[s * gss];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) { C s = 0; for (int w = 0; w < kLsgThreads / 32; w++) s += red[w]; bcast = s; }
    __syncthreads();
    // the reference casts -sum to the logits dtype before the multiply (dag_loss.py:294)
    const C neg = (C)(T)(-bcast);
    for (int c = threadIdx.x; c < V; c += kLsgThreads) rowbuf[c] = (C)(T)((C)x[c] * neg);
    __syncthreads();
    const int64_t *ib = idx + b * isb + l * isl;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) atomicAdd(&rowbuf[ib[s * iss]], (C)(T)gb[s * gss]);
    __syncthreads();
    for (int c = threadIdx.x; c < V; c += kLsgThreads) x[c] = (T)rowbuf[c];
    __syncthreads();
  }
}

// backward for rows beyond the shared-memory staging limits (V > ~51 k elements; > 25 k in fp64): no vocabulary
// limit, as the reference's mul_ + scatter_add_ (dag_loss.py:294-295).  The row is scaled in place in global memory,
// then the S targets are added with global atomics in the row's dtype (one rounding per addition, as scatter_add_).
template <typename T, typename C>
__global__ void __launch_bounds__(kLsgThreads)
lsg_bwd_global_kernel(T *__restrict__ probs, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
                      const C *__restrict__ gout, int64_t gsb, int64_t gsl, int64_t gss, int L, int V, int S, int64_t rows) {
  __shared__ C red[kLsgThreads / 32];
  __shared__ C bcast;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = (int)(row / L), l = (int)(row % L);
    T *x = probs + row * (int64_t)V;
    const C *gb = gout + b * gsb + l * gsl;
    C part = 0;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) part += gb[s * gss];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) { C t = 0; for (int w = 0; w < kLsgThreads / 32; w++) t += red[w]; bcast = t; }
    __syncthreads();
    const C neg = (C)(T)(-bcast);
    for (int c = threadIdx.x; c < V; c += kLsgThreads) x[c] = (T)((C)x[c] * neg);
    __syncthreads();   // the scaled row is visible to the whole CTA before the scatter
    const int64_t *ib = idx + b * isb + l * isl;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) atomicAdd(&x[ib[s * iss]], (T)gb[s * gss]);
    __syncthreads();
  }
}

// vectorised backward for 16-bit / fp32 rows with V % kElems == 0
template <typename T>
__global__ void __launch_bounds__(kLsgThreads)
lsg_bwd_vec_kernel(T *__restrict__ probs, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
                   const float *__restrict__ gout, int64_t gsb, int64_t gsl, int64_t gss, int L, int V, int S) {
  using VT = VecTraits<T>;
  constexpr int E = VT::kElems;
  extern __shared__ float rowf[];  // [V] + 16
  float *red = rowf + V;
  const int64_t row = blockIdx.x;
  const int b = (int)(row / L), l = (int)(row % L);
  T *x = probs + row * (int64_t)V;
  const float *gb = gout + b * gsb + l * gsl;
  float part = 0.f;
  for (int s = threadIdx.x; s < S; s += kLsgThreads) part += gb[s * gss];
  const float tot = block_allreduce<false>(part, red);
  const float neg = VT::to_float(VT::from_float(-tot));
  const int nvec = V / E;
  for (int c = threadIdx.x; c < nvec; c += kLsgThreads) {
    uint4 u = __ldcs(reinterpret_cast<const uint4 *>(x) + c);
    float f[E];
    VT::unpack(u, f);
#pragma unroll
    for (int e = 0; e < E; e++) rowf[c * E + e] = VT::to_float(VT::from_float(f[e] * neg));
  }
  __syncthreads();
  const int64_t *ib = idx + b * isb + l * isl;
  for (int s = threadIdx.x; s < S; s += kLsgThreads)
    atomicAdd(&rowf[ib[s * iss]], VT::to_float(VT::from_float(gb[s * gss])));
  __syncthreads();
  for (int c = threadIdx.x; c < nvec; c += kLsgThreads) {
    float f[E];
#pragma unroll
    for (int e = 0; e < E; e++) f[e] = rowf[c * E + e];
    reinterpret_cast<uint4 *>(x)[c] = VT::pack(f);
  }
}


// ---------------------------------------------------------------------------------------------
// backward, TMA-staged, same organisation as lsg_fwd_tma_kernel: a CTA owns kLsgGroup consecutive vertices of one
// utterance; the incoming gradients of the group are fetched once (32-byte segments of the transposed [B,S,L]
// buffer), the probability rows stream through a ring of bulk copies, every row is scaled into an fp32 copy in
// shared memory (double-buffered), takes the scatter as shared-memory atomics and is written back once.
template <typename T, int NV, int STAGES>
__global__ void __launch_bounds__(kLsgThreads)
lsg_bwd_tma_kernel(T *__restrict__ probs, const int64_t *__restrict__ idx, int64_t isb, int64_t isl, int64_t iss,
                   const float *__restrict__ gout, int64_t gsb, int64_t gsl, int64_t gss, int L, int V, int S, int groups) {
  using VT = VecTraits<T>;
  constexpr int E = VT::kElems;
  extern __shared__ __align__(128) unsigned char lsg_smem[];
  const uint32_t rowbytes = (uint32_t)V * sizeof(T);
  unsigned char *stages = lsg_smem;                                            // [STAGES][rowbytes]
  float *rowf = reinterpret_cast<float *>(lsg_smem + (size_t)STAGES * rowbytes);       // [2][V]
  float *gbuf = rowf + (size_t)2 * V;                                          // [S][kLsgGroup + 1]
  int *idxs = reinterpret_cast<int *>(gbuf + (size_t)S * (kLsgGroup + 1));     // [S]
  float *red = reinterpret_cast<float *>(idxs + S);                            // [8 warps][kLsgGroup] + [kLsgGroup]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + 9 * kLsgGroup);          // [STAGES]

  const int b = blockIdx.x / groups, l0 = (blockIdx.x % groups) * kLsgGroup;
  const int nrows = min(kLsgGroup, L - l0);
  T *xg = probs + ((int64_t)b * L + l0) * (int64_t)V;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsg_smem_u32(full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int r = 0; r < min(STAGES, nrows); r++)
      lsg_load_row(stages + (size_t)r * rowbytes, xg + (int64_t)r * V, rowbytes, full + r);
  }
  const bool shared_idx = (isl == 0);
  if (shared_idx) {
    const int64_t *ib = idx + b * isb;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) idxs[s] = (int)ib[s * iss];
  }
  // incoming gradients of the group and their per-vertex sums
  const float *gb = gout + b * gsb + (int64_t)l0 * gsl;
  const bool vec = (gsl == 1) && nrows == kLsgGroup && ((gss & 3) == 0) && ((gsb & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
  float part[kLsgGroup];
#pragma unroll
  for (int r = 0; r < kLsgGroup; r++) part[r] = 0.f;
  for (int s = threadIdx.x; s < S; s += kLsgThreads) {
    float g[kLsgGroup];
    if (vec) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(gb + s * gss));
      const float4 c = __ldg(reinterpret_cast<const float4 *>(gb + s * gss) + 1);
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = c.x; g[5] = c.y; g[6] = c.z; g[7] = c.w;
    } else {
#pragma unroll
      for (int r = 0; r < kLsgGroup; r++) g[r] = (r < nrows) ? gb[r * gsl + s * gss] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kLsgGroup; r++) { gbuf[s * (kLsgGroup + 1) + r] = g[r]; part[r] += g[r]; }
  }
#pragma unroll
  for (int r = 0; r < kLsgGroup; r++) part[r] = warp_sum(part[r]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int r = 0; r < kLsgGroup; r++) red[(threadIdx.x >> 5) * kLsgGroup + r] = part[r];
  }
  __syncthreads();
  if (threadIdx.x < kLsgGroup) {
    float t = 0.f;
    for (int w = 0; w < kLsgThreads / 32; w++) t += red[w * kLsgGroup + threadIdx.x];
    // the reference casts -sum to the logits dtype before the multiply (dag_loss.py:294)
    red[8 * kLsgGroup + threadIdx.x] = VT::to_float(VT::from_float(-t));
  }
  __syncthreads();
  const int nvec = V / E;

  for (int r = 0; r < nrows; r++) {
    const int st = r % STAGES;
    const T *xs = reinterpret_cast<const T *>(stages + (size_t)st * rowbytes);
    float *rf = rowf + (size_t)(r & 1) * V;
    if (!shared_idx) {
      const int64_t *ib = idx + b * isb + (l0 + r) * isl;
      for (int s = threadIdx.x; s < S; s += kLsgThreads) idxs[s] = (int)ib[s * iss];   // ordered by the barrier below
    }
    const float neg = red[8 * kLsgGroup + r];
    lsg_mbar_wait(full + st, (r / STAGES) & 1);
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = k * kLsgThreads + threadIdx.x;
      if (c < nvec) {
        const uint4 u = reinterpret_cast<const uint4 *>(xs)[c];
        float f[E];
        VT::unpack(u, f);
#pragma unroll
        for (int e = 0; e < E; e++) f[e] = VT::to_float(VT::from_float(f[e] * neg));
        if (E == 8) {
          reinterpret_cast<float4 *>(rf)[2 * c] = make_float4(f[0], f[1], f[2], f[3]);
          reinterpret_cast<float4 *>(rf)[2 * c + 1] = make_float4(f[4 % E], f[5 % E], f[6 % E], f[7 % E]);
        } else {
          reinterpret_cast<float4 *>(rf)[c] = make_float4(f[0], f[1], f[2], f[3]);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && r + STAGES < nrows) {   // every thread is past its reads of this stage
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      lsg_load_row(stages + (size_t)st * rowbytes, xg + (int64_t)(r + STAGES) * V, rowbytes, full + st);
    }
    for (int s = threadIdx.x; s < S; s += kLsgThreads)
      atomicAdd(&rf[idxs[s]], VT::to_float(VT::from_float(gbuf[s * (kLsgGroup + 1) + r])));
    __syncthreads();
    T *x = xg + (int64_t)r * V;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = k * kLsgThreads + threadIdx.x;
      if (c < nvec) {
        float f[E];
        if (E == 8) {
          const float4 a = reinterpret_cast<const float4 *>(rf)[2 * c], d = reinterpret_cast<const float4 *>(rf)[2 * c + 1];
          f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4 % E] = d.x; f[5 % E] = d.y; f[6 % E] = d.z; f[7 % E] = d.w;
        } else {
          const float4 a = reinterpret_cast<const float4 *>(rf)[c];
          f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        }
        reinterpret_cast<uint4 *>(x)[c] = VT::pack(f);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------
// backward for 16-bit rows with a shared (stride-0) index row: the scatter is applied while the row is in registers.
// The S targets of the utterance are bucketed once per CTA by the 16-byte vector they fall into (counting sort in
// shared memory); per row a thread scales its vectors, adds the gradients of the targets in its buckets (usually none)
// and writes -- no fp32 copy of the row in shared memory, one barrier per row, twice the CTAs per SM.
template <typename T, int NV>
__global__ void __launch_bounds__(kLsgThreads)
lsg_bwd_tma16_kernel(T *__restrict__ probs, const int64_t *__restrict__ idx, int64_t isb, int64_t iss,
                     const float *__restrict__ gout, int64_t gsb, int64_t gsl, int64_t gss, int L, int V, int S, int groups) {
  using VT = VecTraits<T>;
  constexpr int E = VT::kElems;
  constexpr int STAGES = kLsgStages;
  extern __shared__ __align__(128) unsigned char lsg_smem[];
  const uint32_t rowbytes = (uint32_t)V * sizeof(T);
  const int nvec = V / E;
  unsigned char *stages = lsg_smem;                                            // [STAGES][rowbytes]
  uint64_t *full = reinterpret_cast<uint64_t *>(lsg_smem + (size_t)STAGES * rowbytes);  // [STAGES] (+ pad to 32 bytes)
  float *gbuf = reinterpret_cast<float *>(full + 4);                           // [S][kLsgGroup + 1]
  int *idxs = reinterpret_cast<int *>(gbuf + (size_t)S * (kLsgGroup + 1));     // [S]
  int *boff = idxs + S;                                                        // [nvec + 1] bucket offsets
  int *bfill = boff + nvec + 1;                                                // [nvec]     fill cursors
  int *blist = bfill + nvec;                                                   // [S]        targets sorted by bucket
  float *red = reinterpret_cast<float *>(blist + S);                           // [8][kLsgGroup] + [kLsgGroup]

  const int b = blockIdx.x / groups, l0 = (blockIdx.x % groups) * kLsgGroup;
  const int nrows = min(kLsgGroup, L - l0);
  T *xg = probs + ((int64_t)b * L + l0) * (int64_t)V;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; i++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lsg_smem_u32(full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int r = 0; r < min(STAGES, nrows); r++)
      lsg_load_row(stages + (size_t)r * rowbytes, xg + (int64_t)r * V, rowbytes, full + r);
  }
  {
    const int64_t *ib = idx + b * isb;
    for (int s = threadIdx.x; s < S; s += kLsgThreads) idxs[s] = (int)ib[s * iss];
    for (int c = threadIdx.x; c <= nvec; c += kLsgThreads) boff[c] = 0;
    for (int c = threadIdx.x; c < nvec; c += kLsgThreads) bfill[c] = 0;
  }
  // incoming gradients of the group and their per-vertex sums
  const float *gb = gout + b * gsb + (int64_t)l0 * gsl;
  const bool vec = (gsl == 1) && nrows == kLsgGroup && ((gss & 3) == 0) && ((gsb & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
  float part[kLsgGroup];
#pragma unroll
  for (int r = 0; r < kLsgGroup; r++) part[r] = 0.f;
  for (int s = threadIdx.x; s < S; s += kLsgThreads) {
    float g[kLsgGroup];
    if (vec) {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(gb + s * gss));
      const float4 c = __ldg(reinterpret_cast<const float4 *>(gb + s * gss) + 1);
      g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = c.x; g[5] = c.y; g[6] = c.z; g[7] = c.w;
    } else {
#pragma unroll
      for (int r = 0; r < kLsgGroup; r++) g[r] = (r < nrows) ? gb[r * gsl + s * gss] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kLsgGroup; r++) {
      part[r] += g[r];
      gbuf[s * (kLsgGroup + 1) + r] = VT::to_float(VT::from_float(g[r]));   // the reference scatters the gradient in the row's dtype
    }
  }
#pragma unroll
  for (int r = 0; r < kLsgGroup; r++) part[r] = warp_sum(part[r]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int r = 0; r < kLsgGroup; r++) red[(threadIdx.x >> 5) * kLsgGroup + r] = part[r];
  }
  __syncthreads();
  if (threadIdx.x < kLsgGroup) {
    float t = 0.f;
    for (int w = 0; w < kLsgThreads / 32; w++) t += red[w * kLsgGroup + threadIdx.x];
    red[8 * kLsgGroup + threadIdx.x] = VT::to_float(VT::from_float(-t));   // dag_loss.py:294: -sum in the row's dtype
  }
  // counting sort of the targets by vector: histogram, exclusive scan, fill
  for (int s = threadIdx.x; s < S; s += kLsgThreads) atomicAdd(&boff[idxs[s] / E + 1], 1);
  __syncthreads();
  {
    // inclusive scan of boff[1 .. nvec] by one warp (nvec <= 2048: 64 entries per lane at most)
    if (threadIdx.x < 32) {
      const int per = (nvec + 31) / 32;
      const int lo = 1 + threadIdx.x * per, hi = min(nvec + 1, lo + per);
      int sum = 0;
      for (int c = lo; c < hi; c++) sum += boff[c];
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)threadIdx.x >= o) incl += v;
      }
      int run = incl - sum;
      for (int c = lo; c < hi; c++) { run += boff[c]; boff[c] = run; }
    }
  }
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += kLsgThreads) {
    const int c = idxs[s] / E;
    blist[boff[c] + atomicAdd(&bfill[c], 1)] = s;
  }
  __syncthreads();

  for (int r = 0; r < nrows; r++) {
    const int st = r % STAGES;
    const T *xs = reinterpret_cast<const T *>(stages + (size_t)st * rowbytes);
    const float neg = red[8 * kLsgGroup + r];
    lsg_mbar_wait(full + st, (r / STAGES) & 1);
    uint4 u[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = k * kLsgThreads + threadIdx.x;
      if (c < nvec) u[k] = reinterpret_cast<const uint4 *>(xs)[c];
    }
    __syncthreads();
    if (threadIdx.x == 0 && r + STAGES < nrows) {   // every thread is past its reads of this stage
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      lsg_load_row(stages + (size_t)st * rowbytes, xg + (int64_t)(r + STAGES) * V, rowbytes, full + st);
    }
    T *x = xg + (int64_t)r * V;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c = k * kLsgThreads + threadIdx.x;
      if (c < nvec) {
        // probs * (-sum) in the row's dtype: one packed multiply per pair, rounded once like the reference's multiply
        // (dag_loss.py:294); only the few vectors that hold a target leave the packed form for the scatter
        uint4 o = VT::scale(u[k], neg);
        const int q0 = boff[c], q1 = boff[c + 1];
        if (q1 > q0) {
          float f[E];
          VT::unpack(o, f);
          for (int q = q0; q < q1; q++) {
            const int s = blist[q];
            const int col = idxs[s] - c * E;
            const float gv = gbuf[s * (kLsgGroup + 1) + r];
#pragma unroll
            for (int e = 0; e < E; e++) f[e] += (e == col) ? gv : 0.f;
          }
          o = VT::pack(f);
        }
        reinterpret_cast<uint4 *>(x)[c] = o;
      }
    }
  }
}

static size_t lsg_bwd_tma16_smem(size_t rowbytes, int nvec, int S) {
  return (size_t)kLsgStages * rowbytes + (size_t)S * (kLsgGroup + 1) * 4 + (size_t)S * 4 + (size_t)(2 * nvec + 1) * 4 + (size_t)S * 4 +
         9 * kLsgGroup * 4 + 32 + 16;
}

template <typename T, int NV>
static int launch_bwd_tma16(T *probs, const int64_t *idx, int64_t isb, int64_t iss, const float *gout, int64_t gsb,
                            int64_t gsl, int64_t gss, int B, int L, int V, int S, cudaStream_t st) {
  const size_t smem = lsg_bwd_tma16_smem((size_t)V * sizeof(T), V / VecTraits<T>::kElems, S);
  const int groups = (L + kLsgGroup - 1) / kLsgGroup;
  const unsigned grid = (unsigned)((int64_t)B * groups);
  cudaFuncSetAttribute(lsg_bwd_tma16_kernel<T, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  lsg_bwd_tma16_kernel<T, NV><<<grid, kLsgThreads, smem, st>>>(probs, idx, isb, iss, gout, gsb, gsl, gss, L, V, S, groups);
  DAGB200_CHECK_LAUNCH("lsg_bwd_tma16_kernel");
  return 0;
}

template <int STAGES> static size_t lsg_bwd_tma_smem(size_t rowbytes, int V, int S) {
  return (size_t)STAGES * rowbytes + (size_t)2 * V * 4 + (size_t)S * (kLsgGroup + 1) * 4 + (size_t)S * 4 + 9 * kLsgGroup * 4 +
         STAGES * 8 + 16;
}

template <typename T, int NV>
static int launch_bwd_tma(T *probs, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss, const float *gout,
                          int64_t gsb, int64_t gsl, int64_t gss, int B, int L, int V, int S, cudaStream_t st) {
  constexpr int STAGES = sizeof(T) == 4 ? 2 : 3;
  const size_t smem = lsg_bwd_tma_smem<STAGES>((size_t)V * sizeof(T), V, S);
  const int groups = (L + kLsgGroup - 1) / kLsgGroup;
  const unsigned grid = (unsigned)((int64_t)B * groups);
  cudaFuncSetAttribute(lsg_bwd_tma_kernel<T, NV, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  lsg_bwd_tma_kernel<T, NV, STAGES><<<grid, kLsgThreads, smem, st>>>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, L, V, S, groups);
  DAGB200_CHECK_LAUNCH("lsg_bwd_tma_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------------------
template <typename T, int NV>
static int launch_fwd_reg(T *logits, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss, float *out,
                          int64_t osb, int64_t osl, int64_t oss, int B, int L, int V, int S, bool grad,
                          cudaStream_t st) {
  const size_t smem = (size_t)(S + 16) * sizeof(float);
  const unsigned grid = (unsigned)((int64_t)B * L);
  if (grad)
    lsg_fwd_reg_kernel<T, NV, true><<<grid, kLsgThreads, smem, st>>>(logits, idx, isb, isl, iss, out, osb, osl, oss, L, V, S);
  else
    lsg_fwd_reg_kernel<T, NV, false><<<grid, kLsgThreads, smem, st>>>(logits, idx, isb, isl, iss, out, osb, osl, oss, L, V, S);
  DAGB200_CHECK_LAUNCH("lsg_fwd_reg_kernel");
  return 0;
}

template <typename T>
static int lsg_fwd_dispatch16(T *logits, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss, float *out,
                              int64_t osb, int64_t osl, int64_t oss, int B, int L, int V, int S, bool grad,
                              cudaStream_t st, int64_t *amax = nullptr) {
  constexpr int E = VecTraits<T>::kElems;
  const bool aligned = (V % E == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int nvec = V / E;
  const int need = (nvec + kLsgThreads - 1) / kLsgThreads;
  // TMA-staged path: rows of <= 32 KB that keep at least two CTAs per SM resident
  static const bool no_tma = getenv("DAGB200_LSG_NOTMA") != nullptr;
  if (!no_tma && aligned && need <= 8 && (size_t)V * sizeof(T) <= 32768 && lsg_fwd_tma_smem((size_t)V * sizeof(T), S) <= 110 * 1024) {
    if (need <= 1) return launch_fwd_tma<T, 1>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st, amax);
    if (need <= 2) return launch_fwd_tma<T, 2>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st, amax);
    if (need <= 4) return launch_fwd_tma<T, 4>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st, amax);
    return launch_fwd_tma<T, 8>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st, amax);
  }
  if (amax) {   // not a TMA-staged shape: the arg-max runs as its own pass BEFORE the row may be overwritten
    lsg_argmax_kernel<T><<<(unsigned)((int64_t)B * L), kLsgThreads, 0, st>>>(logits, amax, V);
    DAGB200_CHECK_LAUNCH("lsg_argmax_kernel");
  }
  if (aligned && need <= 16 && S <= 8192) {
    if (need <= 1) return launch_fwd_reg<T, 1>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st);
    if (need <= 2) return launch_fwd_reg<T, 2>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st);
    if (need <= 4) return launch_fwd_reg<T, 4>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st);
    if (need <= 8) return launch_fwd_reg<T, 8>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st);
    return launch_fwd_reg<T, 16>(logits, idx, isb, isl, iss, out, osb, osl, oss, B, L, V, S, grad, st);
  }
  const int64_t rows = (int64_t)B * L;
  const unsigned grid = (unsigned)(rows < (int64_t)sm_count() * 16 ? rows : (int64_t)sm_count() * 16);
  if (grad)
    lsg_fwd_stream_kernel<T, float, true><<<grid, kLsgThreads, 0, st>>>(logits, idx, isb, isl, iss, out, osb, osl, oss, L, V, S, rows);
  else
    lsg_fwd_stream_kernel<T, float, false><<<grid, kLsgThreads, 0, st>>>(logits, idx, isb, isl, iss, out, osb, osl, oss, L, V, S, rows);
  DAGB200_CHECK_LAUNCH("lsg_fwd_stream_kernel");
  return 0;
}

template <typename T>
static int lsg_bwd_dispatch16(T *probs, const int64_t *idx, int64_t isb, int64_t isl, int64_t iss, const float *gout,
                              int64_t gsb, int64_t gsl, int64_t gss, int B, int L, int V, int S, cudaStream_t st) {
  constexpr int E = VecTraits<T>::kElems;
  const int64_t rows = (int64_t)B * L;
  const bool aligned = (V % E == 0) && ((reinterpret_cast<uintptr_t>(probs) & 15) == 0);
  const size_t smem = (size_t)(V + 16) * sizeof(float);
  if (smem > 200 * 1024) {   // the fp32 copy of a row does not fit in shared memory: global-memory backward, any V
    const unsigned grid = (unsigned)(rows < (int64_t)sm_count() * 8 ? rows : (int64_t)sm_count() * 8);
    lsg_bwd_global_kernel<T, float><<<grid, kLsgThreads, 0, st>>>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, L, V, S, rows);
    DAGB200_CHECK_LAUNCH("lsg_bwd_global_kernel");
    return 0;
  }
  static const bool no_tma = getenv("DAGB200_LSG_NOTMA") != nullptr;
  const int need = (V / E + kLsgThreads - 1) / kLsgThreads;
  // 16-bit rows with the criterion's expanded (stride-0) index view: scatter applied in registers
  if (!no_tma && aligned && sizeof(T) == 2 && isl == 0 && need <= 4 && V / E <= 2048 &&
      lsg_bwd_tma16_smem((size_t)V * sizeof(T), V / E, S) <= 72 * 1024) {
    if (need <= 1) return launch_bwd_tma16<T, 1>(probs, idx, isb, iss, gout, gsb, gsl, gss, B, L, V, S, st);
    if (need <= 2) return launch_bwd_tma16<T, 2>(probs, idx, isb, iss, gout, gsb, gsl, gss, B, L, V, S, st);
    return launch_bwd_tma16<T, 4>(probs, idx, isb, iss, gout, gsb, gsl, gss, B, L, V, S, st);
  }
  if (!no_tma && aligned && need <= 8 && (size_t)V * sizeof(T) <= 32768 &&
      lsg_bwd_tma_smem<sizeof(T) == 4 ? 2 : 3>((size_t)V * sizeof(T), V, S) <= 110 * 1024) {
    if (need <= 1) return launch_bwd_tma<T, 1>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, B, L, V, S, st);
    if (need <= 2) return launch_bwd_tma<T, 2>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, B, L, V, S, st);
    if (need <= 4) return launch_bwd_tma<T, 4>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, B, L, V, S, st);
    return launch_bwd_tma<T, 8>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, B, L, V, S, st);
  }
  if (aligned) {
    cudaFuncSetAttribute(lsg_bwd_vec_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    lsg_bwd_vec_kernel<T><<<(unsigned)rows, kLsgThreads, smem, st>>>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, L, V, S);
    DAGB200_CHECK_LAUNCH("lsg_bwd_vec_kernel");
    return 0;
  }
  cudaFuncSetAttribute(lsg_bwd_kernel<T, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const unsigned grid = (unsigned)(rows < (int64_t)sm_count() * 8 ? rows : (int64_t)sm_count() * 8);
  lsg_bwd_kernel<T, float><<<grid, kLsgThreads, (size_t)V * sizeof(float), st>>>(probs, idx, isb, isl, iss, gout, gsb, gsl, gss, L, V, S, rows);
  DAGB200_CHECK_LAUNCH("lsg_bwd_kernel");
  return 0;
}

}  // namespace dagb200

using namespace dagb200;

extern "C" int dagb200_logsoftmax_gather(void *logits, int dtype, const int64_t *idx, int64_t isb, int64_t isl,
                                         int64_t iss, void *out, int64_t osb, int64_t osl, int64_t oss, int B, int L,
                                         int V, int S, int require_gradient, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && L >= 0 && V >= 1 && S >= 0, DAGB200_EINVAL,
                    "logsoftmax_gather: bad sizes B=%d L=%d V=%d S=%d", B, L, V, S);
  if ((int64_t)B * L == 0) return 0;
  DAGB200_CHECK_ARG(logits && out && (idx || S == 0), DAGB200_EINVAL, "logsoftmax_gather: null pointer");
  DAGB200_CHECK_ARG((int64_t)B * L < (1ll << 31), DAGB200_ELIMIT, "logsoftmax_gather: B*L too large");
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = require_gradient != 0;
  switch (dtype) {
    case DAGB200_F32:
      return lsg_fwd_dispatch16<float>((float *)logits, idx, isb, isl, iss, (float *)out, osb, osl, oss, B, L, V, S, grad, st);
    case DAGB200_F16:
      return lsg_fwd_dispatch16<__half>((__half *)logits, idx, isb, isl, iss, (float *)out, osb, osl, oss, B, L, V, S, grad, st);
    case DAGB200_BF16:
      return lsg_fwd_dispatch16<__nv_bfloat16>((__nv_bfloat16 *)logits, idx, isb, isl, iss, (float *)out, osb, osl, oss, B, L, V, S, grad, st);
    case DAGB200_F64: {
      const int64_t rows = (int64_t)B * L;
      const unsigned grid = (unsigned)(rows < (int64_t)sm_count() * 16 ? rows : (int64_t)sm_count() * 16);
      if (grad)
        lsg_fwd_stream_kernel<double, double, true><<<grid, kLsgThreads, 0, st>>>((double *)logits, idx, isb, isl, iss, (double *)out, osb, osl, oss, L, V, S, rows);
      else
        lsg_fwd_stream_kernel<double, double, false><<<grid, kLsgThreads, 0, st>>>((double *)logits, idx, isb, isl, iss, (double *)out, osb, osl, oss, L, V, S, rows);
      DAGB200_CHECK_LAUNCH("lsg_fwd_stream_kernel<double>");
      return 0;
    }
    default:
      set_error("logsoftmax_gather: unsupported dtype %d", dtype);
      return DAGB200_EDTYPE;
  }
}

extern "C" int dagb200_logsoftmax_gather_argmax(void *logits, int dtype, const int64_t *idx, int64_t isb, int64_t isl,
                                                int64_t iss, void *out, int64_t osb, int64_t osl, int64_t oss,
                                                int64_t *argmax, int B, int L, int V, int S, int require_gradient,
                                                void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && L >= 0 && V >= 1 && S >= 0, DAGB200_EINVAL,
                    "logsoftmax_gather_argmax: bad sizes B=%d L=%d V=%d S=%d", B, L, V, S);
  if ((int64_t)B * L == 0) return 0;
  DAGB200_CHECK_ARG(logits && out && argmax && (idx || S == 0), DAGB200_EINVAL, "logsoftmax_gather_argmax: null pointer");
  DAGB200_CHECK_ARG((int64_t)B * L < (1ll << 31), DAGB200_ELIMIT, "logsoftmax_gather_argmax: B*L too large");
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = require_gradient != 0;
  switch (dtype) {
    case DAGB200_F32:
      return lsg_fwd_dispatch16<float>((float *)logits, idx, isb, isl, iss, (float *)out, osb, osl, oss, B, L, V, S, grad, st, argmax);
    case DAGB200_F16:
      return lsg_fwd_dispatch16<__half>((__half *)logits, idx, isb, isl, iss, (float *)out, osb, osl, oss, B, L, V, S, grad, st, argmax);
    case DAGB200_BF16:
      return lsg_fwd_dispatch16<__nv_bfloat16>((__nv_bfloat16 *)logits, idx, isb, isl, iss, (float *)out, osb, osl, oss, B, L, V, S, grad, st, argmax);
    default:
      set_error("logsoftmax_gather_argmax: unsupported dtype %d (fp32 / fp16 / bf16)", dtype);
      return DAGB200_EDTYPE;
  }
}

extern "C" int dagb200_logsoftmax_gather_backward(void *probs, int dtype, const int64_t *idx, int64_t isb, int64_t isl,
                                                  int64_t iss, const void *gout, int64_t gsb, int64_t gsl, int64_t gss,
                                                  int B, int L, int V, int S, void *stream) {
  DAGB200_CHECK_ARG(B >= 0 && L >= 0 && V >= 1 && S >= 0, DAGB200_EINVAL,
                    "logsoftmax_gather_backward: bad sizes B=%d L=%d V=%d S=%d", B, L, V, S);
  if ((int64_t)B * L == 0) return 0;
  DAGB200_CHECK_ARG(probs && (gout || S == 0) && (idx || S == 0), DAGB200_EINVAL, "logsoftmax_gather_backward: null pointer");
  DAGB200_CHECK_ARG((int64_t)B * L < (1ll << 31), DAGB200_ELIMIT, "logsoftmax_gather_backward: B*L too large");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case DAGB200_F32:
      return lsg_bwd_dispatch16<float>((float *)probs, idx, isb, isl, iss, (const float *)gout, gsb, gsl, gss, B, L, V, S, st);
    case DAGB200_F16:
      return lsg_bwd_dispatch16<__half>((__half *)probs, idx, isb, isl, iss, (const float *)gout, gsb, gsl, gss, B, L, V, S, st);
    case DAGB200_BF16:
      return lsg_bwd_dispatch16<__nv_bfloat16>((__nv_bfloat16 *)probs, idx, isb, isl, iss, (const float *)gout, gsb, gsl, gss, B, L, V, S, st);
    case DAGB200_F64: {
      const int64_t rows = (int64_t)B * L;
      const size_t smem = (size_t)V * sizeof(double);
      const unsigned grid = (unsigned)(rows < (int64_t)sm_count() * 8 ? rows : (int64_t)sm_count() * 8);
      if (smem > 200 * 1024) {
        lsg_bwd_global_kernel<double, double><<<grid, kLsgThreads, 0, st>>>((double *)probs, idx, isb, isl, iss, (const double *)gout, gsb, gsl, gss, L, V, S, rows);
        DAGB200_CHECK_LAUNCH("lsg_bwd_global_kernel<double>");
        return 0;
      }
      cudaFuncSetAttribute(lsg_bwd_kernel<double, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      lsg_bwd_kernel<double, double><<<grid, kLsgThreads, smem, st>>>((double *)probs, idx, isb, isl, iss, (const double *)gout, gsb, gsl, gss, L, V, S, rows);
      DAGB200_CHECK_LAUNCH("lsg_bwd_kernel<double>");
      return 0;
    }
    default:
      set_error("logsoftmax_gather_backward: unsupported dtype %d", dtype);
      return DAGB200_EDTYPE;
  }
}
