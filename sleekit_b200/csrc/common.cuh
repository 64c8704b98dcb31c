// Shared device/host helpers for the sleekit_b200 kernels (sm_100a).
// Everything numeric here follows the reference's op-by-op rounding
// (sleekit/codebook.py:43-95, scaling.py:58-81): separately rounded IEEE ops,
// never a fused multiply-add and never a reciprocal-multiply in place of a divide.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sleekit_b200.h"

namespace slk {

void set_error(const char* fmt, ...);
void note_launch();  // bumps the process-wide kernel-launch counter (slk_launch_count)

#define SLK_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::slk::set_error(__VA_ARGS__);  \
      return SLK_ERR_ARG;             \
    }                                 \
  } while (0)

#define SLK_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ::slk::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),    \
                       __FILE__, __LINE__);                                        \
      return SLK_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

#define SLK_LAUNCH_CHECK()                                                         \
  do {                                                                             \
    ::slk::note_launch();                                                          \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      ::slk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),\
                       __FILE__, __LINE__);                                        \
      return SLK_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

int sm_count();   // of the CURRENT device (cached per device)

// One-time cudaFuncSetAttribute(MaxDynamicSharedMemorySize) PER DEVICE: the attribute belongs to the
// device's context, so a process that works on several GPUs must set it on each of them.
#define SLK_SMEM_ATTR_ONCE(kern, bytes)                                                              \
  do {                                                                                               \
    static unsigned long long done__ = 0;                                                            \
    int dev__ = 0;                                                                                   \
    SLK_CUDA(cudaGetDevice(&dev__));                                                                 \
    if (!((done__ >> (dev__ & 63)) & 1ull)) {                                                        \
      SLK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      done__ |= 1ull << (dev__ & 63);                                                                \
    }                                                                                                \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// IEEE op wrappers, one rounding each, immune to -fmad contraction.
// ---------------------------------------------------------------------------
template <typename T> struct Ieee;
template <> struct Ieee<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float rint(float a) { return rintf(a); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Ieee<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double rint(double a) { return ::rint(a); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

// ---------------------------------------------------------------------------
// Device view of a codebook.  Uniform grids carry zero/step already rounded to
// the compute dtype (NEP 50: the Python float meets an fp32 array as fp32).
// ---------------------------------------------------------------------------
template <typename T>
struct DevGrid {
  int kind;   // 0 uniform, 1 table
  int size;
  T zero;
  T step;
  const float* values;
  const float* limits;
};

template <typename T>
static inline DevGrid<T> make_grid(const slk_codebook* cb) {
  DevGrid<T> g;
  g.kind = cb->kind;
  g.size = cb->size;
  g.zero = (T)cb->lo;
  g.step = (T)cb->step;
  g.values = cb->values;
  g.limits = cb->limits;
  return g;
}

static inline int check_codebook(const slk_codebook* cb) {
  SLK_REQUIRE(cb != nullptr, "codebook is NULL");
  SLK_REQUIRE(cb->kind == 0 || cb->kind == 1, "codebook kind %d not in {0,1}", cb->kind);
  SLK_REQUIRE(cb->size >= 1, "codebook size %d", cb->size);
  if (cb->kind == 0) {
    SLK_REQUIRE(cb->size >= 2 && cb->step > 0, "uniform codebook needs size >= 2 and step > 0");
  } else {
    SLK_REQUIRE(cb->values != nullptr && (cb->size == 1 || cb->limits != nullptr),
                "table codebook needs values and limits");
  }
  return SLK_OK;
}

// number of limits <= x  == np.digitize(x, limits) for ascending limits
template <typename T>
__device__ __forceinline__ int table_bin(const float* __restrict__ limits, int nlim, T x) {
  if (nlim <= 32) {
    int k = 0;
    for (int i = 0; i < nlim; ++i) k += ((T)__ldg(limits + i) <= x) ? 1 : 0;
    return k;
  }
  int lo = 0, hi = nlim;  // first index with limit > x
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((T)__ldg(limits + mid) <= x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Slot (integer-valued T) a uniform grid assigns to x; shift is 0, +1 or -1.
template <typename T>
__device__ __forceinline__ T uniform_slot(const DevGrid<T>& g, T x, int mode) {
  typedef Ieee<T> F;
  T t = F::div(F::sub(x, g.zero), g.step);
  T lo = (T)0, hi = (T)(g.size - 1);
  if (mode == SLK_UP) { t = F::add(t, (T)1); lo = (T)1; }
  else if (mode == SLK_DOWN) { t = F::sub(t, (T)1); hi = (T)(g.size - 2); }
  T k = F::rint(t);
  k = k < lo ? lo : k;   // np.clip: NaN propagates, like the comparisons here
  k = k > hi ? hi : k;
  return k;
}

template <typename T>
__device__ __forceinline__ T uniform_value_of_slot(const DevGrid<T>& g, T k) {
  typedef Ieee<T> F;
  return F::add(F::mul(k, g.step), g.zero);
}

// index a table grid assigns in the given mode (value = values[idx])
template <typename T>
__device__ __forceinline__ int table_index(const DevGrid<T>& g, T x, int mode) {
  int k = table_bin<T>(g.limits, g.size - 1, x);
  if (mode == SLK_UP) k = k + 1 < g.size ? k + 1 : g.size - 1;
  else if (mode == SLK_DOWN) k = k > 0 ? k - 1 : 0;
  return k;
}

// Nearest value, generic over the two kinds (fp32 hot path).
__device__ __forceinline__ float grid_value(const DevGrid<float>& g, float x) {
  if (g.kind == 0) return uniform_value_of_slot<float>(g, uniform_slot<float>(g, x, SLK_NEAREST));
  return __ldg(g.values + table_index<float>(g, x, SLK_NEAREST));
}

__device__ __forceinline__ int grid_index(const DevGrid<float>& g, float x) {
  if (g.kind == 0) return (int)uniform_slot<float>(g, x, SLK_NEAREST);
  return table_index<float>(g, x, SLK_NEAREST);
}

__device__ __forceinline__ float grid_value_of_index(const DevGrid<float>& g, int k) {
  if (g.kind == 0) return uniform_value_of_slot<float>(g, (float)k);
  return __ldg(g.values + k);
}

// ---------------------------------------------------------------------------
// Correctly rounded division by a divisor that is reused many times.
// With y = RN(1/d) obtained once by a true IEEE divide, two FMA residual corrections of
// q = a*y give RN(a/d) (Markstein's theorem: the first correction makes q faithful, the second
// one correctly rounded), for every normal-range quotient unless d's significand is all ones.
// `ok` is false for such divisors (and for extreme exponents); callers then use the plain divide.
// Range: exact for quotients with exponent in [-100, 100] (verified exhaustively); a quotient
// below 2^-102 may be one ulp off because the residual a - d*q underflows.  The kernels only
// divide (i) weights by row scales and (ii) "x - zero" by the codebook step, (iii) codewords by
// reciprocal scales: tiny quotients of (i) are absorbed by the following "x - zero" (|zero| ~ 1),
// (ii) and (iii) cannot be that small for scales >= 1e-16 * min_factor.
// 5 FMA-pipe ops instead of the ~20-instruction IEEE divide sequence; checked exhaustively
// against __fdiv_rn on the device by slk_selftest_fastdiv_f32 (tests/test_gpu_parity.py).
// ---------------------------------------------------------------------------
struct FastDivF { float d, y; int ok; };

__device__ __forceinline__ FastDivF make_fastdiv(float d) {
  FastDivF f;
  f.d = d;
  f.y = __fdiv_rn(1.0f, d);
  const unsigned bits = __float_as_uint(d);
  const int ex = (int)((bits >> 23) & 0xffu) - 127;
  f.ok = (ex > -60 && ex < 60 && (bits & 0x7fffffu) != 0x7fffffu) ? 1 : 0;
  return f;
}

__device__ __forceinline__ float fastdiv_core(float a, float d, float y) {
  float q = __fmul_rn(a, y);
  float r = __fmaf_rn(-d, q, a);
  q = __fmaf_rn(r, y, q);
  r = __fmaf_rn(-d, q, a);
  return __fmaf_rn(r, y, q);
}

__device__ __forceinline__ float fastdiv(float a, const FastDivF& f) {
  return f.ok ? fastdiv_core(a, f.d, f.y) : __fdiv_rn(a, f.d);
}

struct FastDivD { double d, y; int ok; };

__device__ __forceinline__ FastDivD make_fastdiv(double d) {
  FastDivD f;
  f.d = d;
  f.y = __ddiv_rn(1.0, d);
  const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
  const int ex = (int)((bits >> 52) & 0x7ffull) - 1023;
  f.ok = (ex > -400 && ex < 400 && (bits & 0xfffffffffffffull) != 0xfffffffffffffull) ? 1 : 0;
  return f;
}

__device__ __forceinline__ double fastdiv(double a, const FastDivD& f) {
  if (!f.ok) return __ddiv_rn(a, f.d);
  double q = __dmul_rn(a, f.y);
  double r = __fma_rn(-f.d, q, a);
  q = __fma_rn(r, f.y, q);
  r = __fma_rn(-f.d, q, a);
  return __fma_rn(r, f.y, q);
}

// Uniform-grid rounding with the divide by `step` done through a FastDivF (same results).
__device__ __forceinline__ float uniform_value_fast(const DevGrid<float>& g, const FastDivF& fstep, float x) {
  float t = fastdiv(__fsub_rn(x, g.zero), fstep);
  float k = rintf(t);
  const float hi = (float)(g.size - 1);
  k = k < 0.0f ? 0.0f : k;
  k = k > hi ? hi : k;
  return __fadd_rn(__fmul_rn(k, g.step), g.zero);
}

// ---------------------------------------------------------------------------
// Breakpoints of a uniform codebook in the scaled domain: X[k] = min{x : slot(x) >= k} for
// slot(x) = clip(rint((x - zero) / step), 0, C-1)   (codebook.py:60-62), k = 1..C-1, +inf beyond.
// slot() is a composition of monotone correctly rounded fp32 operations, so X[k] is found exactly
// by bisection over the ordered fp32 values.  It runs on the HOST, once per call (15 x 32 steps):
// IEEE fp32 subtract / divide / nearbyint there give the same results as the device chain (whose
// reciprocal-based divide is the correctly rounded quotient), and the table travels to the
// kernels by value.
// ---------------------------------------------------------------------------
struct GridBreaks { float X[16]; };

static inline GridBreaks make_breaks(const slk_codebook* cb) {
  GridBreaks b;
  const float inf = __builtin_inff();
  for (int k = 0; k < 16; ++k) b.X[k] = inf;
  if (cb->kind != 0 || cb->size > 16) return b;
  const volatile float zero = (float)cb->lo, step = (float)cb->step;
  const float top = (float)(cb->size - 1);
  auto ord = [](float x) { int32_t i; __builtin_memcpy(&i, &x, 4); return (int64_t)(i >= 0 ? i : (int32_t)(0x80000000u - (uint32_t)i)); };
  auto unord = [](int64_t o) { int32_t i = (int32_t)o; i = i >= 0 ? i : (int32_t)(0x80000000u - (uint32_t)i); float x; __builtin_memcpy(&x, &i, 4); return x; };
  auto slot = [&](float x) {
    volatile float d = x - zero;          // volatile: one rounding per operation, no contraction
    volatile float t = d / step;
    float kk = __builtin_nearbyintf(t);
    kk = kk < 0.0f ? 0.0f : kk;
    return kk > top ? top : kk;
  };
  for (int k = 1; k < cb->size; ++k) {
    int64_t lo = ord(-3.402823466e+38f), hi = ord(3.402823466e+38f);   // slots 0 and C-1
    while (hi - lo > 1) {
      const int64_t mid = lo + ((hi - lo) >> 1);
      if (slot(unord(mid)) >= (float)k) hi = mid; else lo = mid;
    }
    b.X[k] = unord(hi);
  }
  return b;
}

// ---------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T u = __shfl_xor_sync(0xffffffffu, v, o); v = u < v ? u : v; }
  return v;
}

// Block-wide sum; `scratch` must hold >= 32 T.  Result valid in every thread.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T t = lane < nw ? scratch[lane] : (T)0;
  return warp_sum(t);
}

}  // namespace slk
