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
                              cudaStream_t st) {
  constexpr int E = VecTraits<T>::kElems;
  const bool aligned = (V % E == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int nvec = V / E;
  const int need = (nvec + kLsgThreads - 1) / kLsgThreads;
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
  if (smem > 200 * 1024) { set_error("logsoftmax_gather_backward: V=%d too large for the staged row", V); return DAGB200_ELIMIT; }
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
      DAGB200_CHECK_ARG(smem <= 200 * 1024, DAGB200_ELIMIT, "logsoftmax_gather_backward: V=%d too large (fp64)", V);
      cudaFuncSetAttribute(lsg_bwd_kernel<double, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      const unsigned grid = (unsigned)(rows < (int64_t)sm_count() * 8 ? rows : (int64_t)sm_count() * 8);
      lsg_bwd_kernel<double, double><<<grid, kLsgThreads, smem, st>>>((double *)probs, idx, isb, isl, iss, (const double *)gout, gsb, gsl, gss, L, V, S, rows);
      DAGB200_CHECK_LAUNCH("lsg_bwd_kernel<double>");
      return 0;
    }
    default:
      set_error("logsoftmax_gather_backward: unsupported dtype %d", dtype);
      return DAGB200_EDTYPE;
  }
}
