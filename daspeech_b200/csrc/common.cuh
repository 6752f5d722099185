// common.cuh -- shared device helpers for libdagb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/dagb200.h"

namespace dagb200 {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float neg_inf_f() { return __int_as_float(0xff800000); }

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return __int_as_float(0xff800000); }
template <> __device__ __forceinline__ double neg_inf<double>() { return __longlong_as_double(0xfff0000000000000ULL); }

// exp(x) for x <= 0 through MUFU.EX2 (fp32) / libm (fp64)
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * kLog2e); }
__device__ __forceinline__ double fast_exp(double x) { return exp(x); }
__device__ __forceinline__ float acc_log(float x) { return logf(x); }
__device__ __forceinline__ double acc_log(double x) { return log(x); }
__device__ __forceinline__ float acc_exp(float x) { return expf(x); }
__device__ __forceinline__ double acc_exp(double x) { return exp(x); }

// ---- online log-sum-exp accumulator: value = m + log(s), s relative to running max m ----------
// m starts at a large finite negative sentinel so that (-inf) candidates need no special casing:
// x=-inf -> d=-inf -> e=0 -> s unchanged.  One MUFU per candidate.
template <typename T> struct Lse {
  T m, s;
  __device__ __forceinline__ void init() { m = (T)-3.0e38; s = (T)0; }
  __device__ __forceinline__ void add(T x) {
    T d = x - m;
    T e = fast_exp(-fabs(d));
    s = (d > (T)0) ? fma(s, e, (T)1) : (s + e);
    m = fmax(m, x);
  }
  __device__ __forceinline__ void merge(const Lse &o) {
    T d = o.m - m;
    T e = fast_exp(-fabs(d));
    s = (d > (T)0) ? fma(s, e, o.s) : fma(o.s, e, s);
    m = fmax(m, o.m);
  }
  // reference rule (dag_loss.cu:113-127): all candidates -inf -> -inf and NO emission added
  __device__ __forceinline__ T finish(T emission) const {
    if (s == (T)0) return neg_inf<T>();
    if (isinf(m)) return m;
    return acc_log(s) + m + emission;
  }
};
template <> __device__ __forceinline__ void Lse<double>::init() { m = -1.0e300; s = 0.0; }

template <typename T> __device__ __forceinline__ T shfl_xor(T v, int lane_mask) {
  return __shfl_xor_sync(0xffffffffu, v, lane_mask);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// thread-local error message plumbing (host side)
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *where);

#define DAGB200_CHECK_ARG(cond, code, ...)  \
  do {                                      \
    if (!(cond)) {                          \
      dagb200::set_error(__VA_ARGS__);      \
      return (code);                        \
    }                                       \
  } while (0)

#define DAGB200_CHECK_LAUNCH(where)                               \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return dagb200::cuda_fail(e__, where); \
  } while (0)

int sm_count();

constexpr int kProfMarks = 8;
void prof_mark(int idx, cudaStream_t st);   // no-op unless dagb200_set_profile(1)

}  // namespace dagb200
