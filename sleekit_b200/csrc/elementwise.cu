// K4: codebook rounding, plus the HBM-bound elementwise / reduction helpers of the
// hot path (row scales, axis scaling, bias removal, column permutation, ordering keys).
// All of them are one pass over their input, 128-bit vectorised where the layout allows,
// with grids sized in multiples of the SM count.
#include "common.cuh"
#include <stdlib.h>

#include <stdarg.h>

namespace slk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_launches = 0;
void note_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& c = cached[dev & 63];
  if (c == 0) c = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  return c;
}

static inline int stream_grid(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = ceil_div(work_items, per_block);
  int64_t cap = (int64_t)sm_count() * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------
// rounding
// ---------------------------------------------------------------------------
template <typename T, typename VT>
__device__ __forceinline__ void round_one(const DevGrid<T>& g, T x, int mode, VT* out_val,
                                          void* out_idx, int idx_bytes, int64_t i) {
  if (g.kind == 0) {
    T k = uniform_slot<T>(g, x, mode);
    if (out_val) out_val[i] = (VT)uniform_value_of_slot<T>(g, k);
    if (out_idx) {
      unsigned u = (unsigned)k;
      if (idx_bytes == 1) ((uint8_t*)out_idx)[i] = (uint8_t)u;
      else if (idx_bytes == 2) ((uint16_t*)out_idx)[i] = (uint16_t)u;
      else ((uint32_t*)out_idx)[i] = u;
    }
  } else {
    int k = table_index<T>(g, x, mode);
    if (out_val) out_val[i] = (VT)__ldg(g.values + k);
    if (out_idx) {
      if (idx_bytes == 1) ((uint8_t*)out_idx)[i] = (uint8_t)k;
      else if (idx_bytes == 2) ((uint16_t*)out_idx)[i] = (uint16_t)k;
      else ((uint32_t*)out_idx)[i] = (uint32_t)k;
    }
  }
}

// Index-only nearest rounding through the codebook's exact breakpoints (uniform codebooks of <= 8
// entries): idx(w) >= k  <=>  w >= X[k]  with X found on the host by bisection over the reference's own
// op chain (make_breaks), so a 3-level compare tree gives the reference's index bit for bit -- 9
// instructions per value instead of ~17 (subtract, exact divide, rint, two clips, float->int, pack);
// NaN compares false everywhere and yields 0, as the conversion of the clipped NaN does.
// 16 values per thread and step: four 128-bit loads in flight, one 128-bit store.
__global__ void __launch_bounds__(256) round_f32_index_tree_kernel(const float4* __restrict__ x, int64_t nquad,
                                                                   GridBreaks brk, int size,
                                                                   uint4* __restrict__ out_idx8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float X[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) X[k] = (k >= 1 && k < size) ? brk.X[k] : __int_as_float(0x7f800000);   // +inf: never reached
  auto idx = [&](float w) -> uint32_t {
    const bool b2 = w >= X[4];
    const float t1 = b2 ? X[6] : X[2];
    const float xa = b2 ? X[7] : X[3], xb = b2 ? X[5] : X[1];
    const bool b1 = w >= t1;
    const float t0 = b1 ? xa : xb;
    const bool b0 = w >= t0;
    return (b2 ? 4u : 0u) + (b1 ? 2u : 0u) + (b0 ? 1u : 0u);
  };
  auto pack = [&](const float4& v) -> uint32_t {
    return idx(v.x) | (idx(v.y) << 8) | (idx(v.z) << 16) | (idx(v.w) << 24);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquad; i += stride) {
    const float4 a = __ldcs(x + 4 * i), b = __ldcs(x + 4 * i + 1), c = __ldcs(x + 4 * i + 2), d = __ldcs(x + 4 * i + 3);
    out_idx8[i] = make_uint4(pack(a), pack(b), pack(c), pack(d));
  }
}

// fp32 fast path: 4 values per thread per step, 128-bit loads and stores.
__global__ void __launch_bounds__(256) round_f32_vec4_kernel(const float4* __restrict__ x, int64_t nvec,
                                                             DevGrid<float> g, int mode,
                                                             float4* __restrict__ out_val,
                                                             uint32_t* __restrict__ out_idx8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // the divide by the codebook step goes through the exact reciprocal scheme (common.cuh: the
  // correctly rounded quotient, 5 FMA-pipe operations instead of the ~20 of the IEEE divide
  // sequence), which is what keeps this kernel on the HBM roofline instead of the ALU
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);
  const float slo = mode == SLK_UP ? 1.0f : 0.0f;
  const float shi = (float)(g.size - (mode == SLK_DOWN ? 2 : 1));
  const float shift = mode == SLK_UP ? 1.0f : (mode == SLK_DOWN ? -1.0f : 0.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 v = __ldcs(x + i);
    float in[4] = {v.x, v.y, v.z, v.w};
    float val[4];
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (g.kind == 0 && fstep.ok) {
        float t = fastdiv_core(__fsub_rn(in[k], g.zero), fstep.d, fstep.y);     // codebook.py:47,60,71,83
        if (mode != SLK_NEAREST) t = __fadd_rn(t, shift);                        // codebook.py:73,85
        float s = rintf(t);
        s = s < slo ? slo : s;      // np.clip: NaN propagates, like the comparisons here
        s = s > shi ? shi : s;
        val[k] = uniform_value_of_slot<float>(g, s);
        packed |= ((uint32_t)s & 0xffu) << (8 * k);
      } else if (g.kind == 0) {
        float s = uniform_slot<float>(g, in[k], mode);
        val[k] = uniform_value_of_slot<float>(g, s);
        packed |= ((uint32_t)s & 0xffu) << (8 * k);
      } else {
        int s = table_index<float>(g, in[k], mode);
        val[k] = __ldg(g.values + s);
        packed |= ((uint32_t)s & 0xffu) << (8 * k);
      }
    }
    if (out_val) __stcs(out_val + i, make_float4(val[0], val[1], val[2], val[3]));
    if (out_idx8) out_idx8[i] = packed;
  }
}

template <typename T, typename VT>
__global__ void __launch_bounds__(256) round_scalar_kernel(const T* __restrict__ x, int64_t begin, int64_t count,
                                                           DevGrid<T> g, int mode, VT* out_val,
                                                           void* out_idx, int idx_bytes) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    round_one<T, VT>(g, x[i], mode, out_val, out_idx, idx_bytes, i);
}

static inline int index_bytes(int size) { return size <= 256 ? 1 : (size <= 65536 ? 2 : 4); }

template <typename T>
static int round_impl(const T* x, int64_t count, const slk_codebook* cb, int mode, void* out_val,
                      void* out_idx, cudaStream_t st) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(mode >= 0 && mode <= 2, "round mode %d", mode);
  SLK_REQUIRE(count >= 0, "negative count");
  if (count == 0) return SLK_OK;
  SLK_REQUIRE(x != nullptr, "x is NULL");
  SLK_REQUIRE(cb->kind == 1 || cb->size >= 2, "uniform codebook of size < 2");
  if (cb->kind == 0 && mode != SLK_NEAREST) SLK_REQUIRE(cb->size >= 2, "up/down need size >= 2");
  DevGrid<T> g = make_grid<T>(cb);
  const int ib = index_bytes(cb->size);
  int64_t done = 0;
  if (sizeof(T) == 4 && ib == 1) {
    bool aligned = ((uintptr_t)x % 16 == 0) && (!out_val || (uintptr_t)out_val % 16 == 0) &&
                   (!out_idx || (uintptr_t)out_idx % 4 == 0);
    int64_t nvec = aligned ? count / 4 : 0;
    int64_t vdone = 0;                       // float4 groups already rounded
    static int tree = -1;   // SLK_K4_TREE=0: arithmetic index rounding instead of the compare tree (A/B testing)
    if (tree < 0) {
      const char* ev = getenv("SLK_K4_TREE");
      tree = (ev && ev[0] == '0') ? 0 : 1;
    }
    if (tree && nvec >= 4 && !out_val && out_idx && cb->kind == 0 && cb->size <= 8 && mode == SLK_NEAREST &&
        (uintptr_t)out_idx % 16 == 0) {
      // index only: compare tree over the exact breakpoints, 16 values per thread
      const int64_t nquad = nvec / 4;
      round_f32_index_tree_kernel<<<stream_grid(nquad, 256), 256, 0, st>>>((const float4*)x, nquad, make_breaks(cb),
                                                                            (int)cb->size, (uint4*)out_idx);
      SLK_LAUNCH_CHECK();
      vdone = nquad * 4;
    }
    if (nvec > vdone) {
      round_f32_vec4_kernel<<<stream_grid(nvec - vdone, 256), 256, 0, st>>>(
          (const float4*)x + vdone, nvec - vdone, *(DevGrid<float>*)&g, mode,
          out_val ? (float4*)out_val + vdone : nullptr, out_idx ? (uint32_t*)out_idx + vdone : nullptr);
      SLK_LAUNCH_CHECK();
    }
    done = nvec * 4;
  }
  if (done < count) {
    if (sizeof(T) == 8 && cb->kind == 1)
      round_scalar_kernel<T, float><<<stream_grid(count - done, 256), 256, 0, st>>>(
          x, done, count, g, mode, (float*)out_val, out_idx, ib);
    else
      round_scalar_kernel<T, T><<<stream_grid(count - done, 256), 256, 0, st>>>(
          x, done, count, g, mode, (T*)out_val, out_idx, ib);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

// ---------------------------------------------------------------------------
// axis scaling: out[o, a, i] = x[o, a, i] / s[a]   (or / (1/s[a]))
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) scale_axis_kernel(const T* __restrict__ x, int64_t total, int64_t len,
                                                         int64_t inner, const T* __restrict__ s, int mode,
                                                         T* __restrict__ out) {
  typedef Ieee<T> F;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t a = (i / inner) % len;
    T d = s[a];
    if (mode == 1) d = F::div((T)1, d);
    out[i] = F::div(x[i], d);
  }
}

// rows contiguous (inner == n, one divisor per row): 128-bit path for fp32
__global__ void __launch_bounds__(256) scale_rows_vec4_kernel(const float4* __restrict__ x, int64_t r, int64_t nvec,
                                                              const float* __restrict__ s, int mode,
                                                              float4* __restrict__ out) {
  const int64_t total = r * nvec;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float d = __ldg(s + i / nvec);
    if (mode == 1) d = __fdiv_rn(1.0f, d);
    float4 v = __ldcs(x + i);
    v.x = __fdiv_rn(v.x, d); v.y = __fdiv_rn(v.y, d); v.z = __fdiv_rn(v.z, d); v.w = __fdiv_rn(v.w, d);
    __stcs(out + i, v);
  }
}

template <typename T>
static int scale_axis_impl(const T* x, int64_t outer, int64_t len, int64_t inner, const T* s, int mode,
                           T* out, cudaStream_t st) {
  SLK_REQUIRE(outer >= 0 && len >= 0 && inner >= 0, "negative extent");
  SLK_REQUIRE(mode == 0 || mode == 1, "scale mode %d", mode);
  int64_t total = outer * len * inner;
  if (total == 0) return SLK_OK;
  SLK_REQUIRE(x && s && out, "NULL pointer");
  if (sizeof(T) == 4 && inner % 4 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)out % 16 == 0 && outer == 1) {
    scale_rows_vec4_kernel<<<stream_grid(total / 4, 256), 256, 0, st>>>(
        (const float4*)x, len, inner / 4, (const float*)s, mode, (float4*)out);
  } else {
    scale_axis_kernel<T><<<stream_grid(total, 256), 256, 0, st>>>(x, total, len, inner, s, mode, out);
  }
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

// ---------------------------------------------------------------------------
// per-row scales
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) row_noclip_scale_kernel(const T* __restrict__ w, int64_t r, int64_t n,
                                                               T cb_min, T cb_max, T* __restrict__ out) {
  typedef Ieee<T> F;
  __shared__ T smin[32], smax[32];
  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    const T* p = w + row * n;
    T lo = p[0], hi = p[0];
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      T v = p[j];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    lo = warp_min(lo); hi = warp_max(hi);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x < 32) {
      int nw = (blockDim.x + 31) >> 5;
      lo = threadIdx.x < nw ? smin[threadIdx.x] : smin[0];
      hi = threadIdx.x < nw ? smax[threadIdx.x] : smax[0];
      lo = warp_min(lo); hi = warp_max(hi);
      if (threadIdx.x == 0) {
        // scaling.py:53-54: max(max/maxcode, min/mincode), floored at float32(1e-16)
        T a = F::div(hi, cb_max), b = F::div(lo, cb_min);
        T s = a > b ? a : b;
        T fl = (T)1.0e-16f;
        out[row] = s > fl ? s : fl;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) row_rms_scale_kernel(const T* __restrict__ w, int64_t r, int64_t n,
                                                            T* __restrict__ out) {
  typedef Ieee<T> F;
  __shared__ double scratch[32];
  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    const T* p = w + row * n;
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      T v = p[j];
      acc += (double)F::mul(v, v);
    }
    acc = block_sum<double>(acc, scratch);
    if (threadIdx.x == 0) {
      // scaling.py:40-41: sqrt(max(mean(x^2), 1e-16)); the mean is rounded to the data dtype
      T m = F::div((T)acc, (T)n);
      T fl = (T)1.0e-16;
      m = m > fl ? m : fl;
      out[row] = F::sqrt(m);
    }
  }
}

// ---------------------------------------------------------------------------
// H - outer(m, m)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) remove_bias_kernel(const T* __restrict__ h, const T* __restrict__ m, int64_t n,
                                                          T* __restrict__ out) {
  typedef Ieee<T> F;
  const int64_t total = n * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t a = i / n, b = i - a * n;
    out[i] = F::sub(h[i], F::mul(m[a], m[b]));
  }
}

// ---------------------------------------------------------------------------
// column gather / scatter of a [r, n] matrix
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) permute_cols_kernel(const float* __restrict__ src, int64_t r, int64_t n,
                                                           const int64_t* __restrict__ idx, int scatter,
                                                           float* __restrict__ dst) {
  const int64_t total = r * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t row = i / n, j = i - row * n;
    int64_t k = __ldg(idx + j);
    if (scatter) dst[row * n + k] = src[i];
    else dst[i] = __ldg(src + row * n + k);
  }
}

// Row scaling fused with the column permutation (the two passes on either side of the sweep):
//   gather : dst[row, j]      = src[row, idx[j]] / s[row]            scaling.py:73 then obq.py:202
//   scatter: dst[row, idx[j]] = src[row, j] / (1 / s[row])           obq.py:212-213 then scaling.py:80
// Same separately rounded divides as slk_scale_axis_f32; one pass over W instead of two.
__global__ void __launch_bounds__(256) scale_permute_cols_kernel(const float* __restrict__ src, int64_t r, int64_t n,
                                                                 const int64_t* __restrict__ idx,
                                                                 const float* __restrict__ s, int scatter,
                                                                 float* __restrict__ dst) {
  const int64_t total = r * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t row = i / n, j = i - row * n;
    const int64_t k = idx ? __ldg(idx + j) : j;
    float d = __ldg(s + row);
    if (scatter) {
      d = __fdiv_rn(1.0f, d);
      dst[row * n + k] = __fdiv_rn(src[i], d);
    } else {
      dst[i] = __fdiv_rn(__ldg(src + row * n + k), d);
    }
  }
}

// ---------------------------------------------------------------------------
// ordering helpers
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) damp_value_kernel(const float* __restrict__ h, int64_t n, float damp,
                                                         float* __restrict__ out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += (double)h[j * (n + 1)];
  acc = block_sum<double>(acc, scratch);
  if (threadIdx.x == 0) {
    // obq.py:198: damp (Python float, weak) * float32 mean -> float32
    float mean = __fdiv_rn((float)acc, (float)n);
    out[0] = __fmul_rn(damp, mean);
  }
}

// one thread per column, rows added in order (numpy's axis-0 reduction order)
__global__ void __launch_bounds__(128) col_resid_sums_kernel(const float* __restrict__ w, int64_t r, int64_t n,
                                                             DevGrid<float> g, int mode, float* __restrict__ out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float acc = 0.0f;
  for (int64_t i = 0; i < r; ++i) {
    float x = __ldg(w + i * n + j);
    float d = __fsub_rn(grid_value(g, x), x);
    float t = mode == 0 ? fabsf(d) : __fmul_rn(d, d);
    acc = __fadd_rn(acc, t);
  }
  out[j] = acc;
}

__global__ void __launch_bounds__(256) order_keys_kernel(const float* __restrict__ h, int64_t n,
                                                         const float* __restrict__ dampval,
                                                         const float* __restrict__ colsum, double* __restrict__ keys) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double d = (double)h[j * (n + 1)];
  if (dampval) d = __dadd_rn(d, (double)dampval[0]);
  double k = -d;
  if (colsum) k = __dmul_rn(k, (double)colsum[j]);
  keys[j] = k;
}

// Stable rank sort: order[rank(j)] = j with rank = #{i : key_i < key_j or (== and i < j)}.
// One WARP per key: the lanes split the n candidates and a shuffle reduction adds the counts, so
// the serial chain is n/32 compares (n = 3072: 96) and the grid has n/8 CTAs.  Keys are mapped to
// order-preserving unsigned integers first (-0 = +0, NaN last, as numpy sorts).  No scratch.
__device__ __forceinline__ unsigned long long sortable_key(double k) {
  if (k != k) return ~0ull;
  k = __dadd_rn(k, 0.0);                                   // -0.0 -> +0.0
  const unsigned long long b = (unsigned long long)__double_as_longlong(k);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) argsort_rank_kernel(const double* __restrict__ keys, int64_t n,
                                                           int64_t* __restrict__ order) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  const unsigned long long uj = sortable_key(__ldg(keys + j));
  int cnt = 0;
  for (int64_t i = lane; i < n; i += 32) {
    const unsigned long long ui = sortable_key(__ldg(keys + i));
    cnt += (ui < uj || (ui == uj && i < j)) ? 1 : 0;
  }
  cnt = warp_sum(cnt);
  if (lane == 0) order[cnt] = j;
}

// ---------------------------------------------------------------------------
// act_order = "pivot": greedy pivoted-Cholesky ordering                         obq.py:140-166
// The reference swaps rows and columns of a working copy; here the matrix stays in place and only
// the order array is swapped: position p holds index order[p], the pivot is the first position
// p >= k with the largest |L[order[p]][order[p]]| (np.argmax), and the trailing update
//     L[i][j] -= (L[piv][i] * L[piv][j]) / L[piv][piv]        (obq.py:160-161, three roundings)
// runs over the still-active indices.  Same fp64 operations on the same operands as numpy, hence
// the same order.  Two launches per step (select: one CTA; update: the grid).
// ---------------------------------------------------------------------------
struct PivotState { int piv; int pad; double lpp; };

__global__ void __launch_bounds__(256) pivot_init_kernel(const double* __restrict__ h, int64_t n, double* __restrict__ L,
                                                         int64_t* __restrict__ order, int* __restrict__ active) {
  const int64_t total = n * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    L[t] = h[t];
    if (t < n) { order[t] = t; active[t] = 1; }
  }
}

__global__ void __launch_bounds__(1024) pivot_select_kernel(const double* __restrict__ L, int64_t n, int64_t k,
                                                            int64_t* __restrict__ order, int* __restrict__ active,
                                                            PivotState* __restrict__ stp) {
  __shared__ double bv[32];
  __shared__ int64_t bp[32];
  double best = -1.0;
  int64_t bpos = INT64_MAX;
  for (int64_t p = k + threadIdx.x; p < n; p += blockDim.x) {
    const int64_t i = order[p];
    const double v = fabs(L[i * n + i]);
    if (bpos == INT64_MAX || v > best) { best = v; bpos = p; }     // p ascends: the first maximum stays
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long op = __shfl_xor_sync(0xffffffffu, (long long)bpos, o);
    if (op != INT64_MAX && (bpos == INT64_MAX || ov > best || (ov == best && op < bpos))) { best = ov; bpos = op; }
  }
  if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bp[threadIdx.x >> 5] = bpos; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (bp[w] != INT64_MAX && (bpos == INT64_MAX || bv[w] > best || (bv[w] == best && bp[w] < bpos))) { best = bv[w]; bpos = bp[w]; }
    const int64_t piv = order[bpos];
    order[bpos] = order[k];                      // obq.py:157
    order[k] = piv;
    active[piv] = 0;
    stp->piv = (int)piv;
    stp->lpp = L[piv * n + piv];
  }
}

__global__ void __launch_bounds__(256) pivot_update_kernel(double* __restrict__ L, int64_t n, const int* __restrict__ active,
                                                           const PivotState* __restrict__ stp) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t piv = stp->piv;
  const double lpp = stp->lpp;
  const bool jact = j < n && active[j];
  const double bj = jact ? L[piv * n + j] : 0.0;
  const int64_t i0 = (int64_t)blockIdx.y * 16;
  for (int64_t i = i0; i < i0 + 16 && i < n; ++i) {
    if (!active[i] || !jact) continue;
    const double bi = L[piv * n + i];
    const double t = __ddiv_rn(__dmul_rn(bi, bj), lpp);                      // np.outer(b, b) / L[k, k]
    L[i * n + j] = __dsub_rn(L[i * n + j], t);
  }
}

// Exhaustive check of fastdiv_core against the IEEE divide: for each divisor, all 2^32 dividends.
// Counted: results that differ (sign of zero ignored) while the true quotient is zero or has an
// exponent in [-100, 100].  (Below 2^-102 the residual a - d*q is not representable and the
// scheme can be one ulp off; every use in the kernels feeds such quotients into "x - zero" with
// |zero| ~ 1, which absorbs them -- see common.cuh.)
__global__ void __launch_bounds__(256) selftest_fastdiv_kernel(const float* __restrict__ divisors,
                                                               unsigned long long* __restrict__ mismatches,
                                                               unsigned long long stride_limit) {
  const float d = divisors[blockIdx.y];
  const FastDivF f = make_fastdiv(d);
  unsigned long long bad = 0;
  const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < stride_limit; i += step) {
    const float a = __uint_as_float((unsigned)i);
    if (!(fabsf(a) <= 3.0e38f)) continue;  // skip inf / nan
    if (a != 0.0f && fabsf(a) < 7.8886091e-31f) continue;  // dividends below 2^-100: residual underflows
    const float want = __fdiv_rn(a, d);
    const float aw = fabsf(want);
    if (!(aw == 0.0f || (aw >= 7.8886091e-31f && aw <= 1.2676506e30f))) continue;
    const float got = f.ok ? fastdiv_core(a, f.d, f.y) : want;
    if (!(got == want)) ++bad;
  }
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches + blockIdx.y, bad);
}

__global__ void __launch_bounds__(256) mean_kernel_f32(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += (double)v[j];
  acc = block_sum<double>(acc, scratch);
  if (threadIdx.x == 0) out[0] = __fdiv_rn((float)acc, (float)n);
}
__global__ void __launch_bounds__(256) mean_kernel_f64(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) acc += v[j];
  acc = block_sum<double>(acc, scratch);
  if (threadIdx.x == 0) out[0] = acc / (double)n;
}

// delta[row] = sum_j (w - wq)[row, j] * mean[j]
__global__ void __launch_bounds__(256) bias_delta_kernel(const float* __restrict__ w, const float* __restrict__ wq,
                                                         const float* __restrict__ mean, int64_t r, int64_t n,
                                                         float* __restrict__ delta) {
  __shared__ double scratch[32];
  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      float d = __fsub_rn(w[row * n + j], wq[row * n + j]);
      acc += (double)__fmul_rn(d, __ldg(mean + j));
    }
    acc = block_sum<double>(acc, scratch);
    if (threadIdx.x == 0) delta[row] = (float)acc;
  }
}

// out[row] = sum_j h[j] * e[row, j]^2  (h NULL: plain sum of squares) -- the None / 1-D branches of
// _compute_mse (scaling.py:84-95); products and squares rounded as numpy rounds them, sums in fp64.
// TE: dtype of e (the square is rounded in it, as np.square(E) is); TH: promoted dtype of h and of the result.
template <typename TE, typename TH>
__global__ void __launch_bounds__(256) row_wsq_kernel(const TE* __restrict__ e, const TH* __restrict__ h, int64_t r, int64_t n,
                                                      TH* __restrict__ out) {
  __shared__ double scratch[32];
  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    double acc = 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      const TE v = e[row * n + j];
      const TE sq = Ieee<TE>::mul(v, v);
      acc += (double)(h ? Ieee<TH>::mul(h[j], (TH)sq) : (TH)sq);
    }
    acc = block_sum<double>(acc, scratch);
    if (threadIdx.x == 0) out[row] = (TH)acc;
  }
}

// H[i][j] = H[j][i] for the 64x64 tiles (ti > tj) that lie below the block diagonal of the symmetric upload
// / unpack (tiles of one copy block share floor(32-row tile index / tpb)).  Grid (nt64, nt64): CTA (bx, by)
// with by > bx mirrors tile (bx, by) into (by, bx) through shared memory, so that both the read and the write
// are 256-byte row segments; CTAs of one grid row write one band of 64 matrix rows (consecutive DRAM pages).
// (The first version used 32x32 tiles and a linear tile id decoded by a loop: 22 ms at n = 28672 -- TLB-
// and decode-bound; this one moves the same 3.2 GB in about a millisecond.)
__global__ void __launch_bounds__(256) mirror_upper_kernel(float* __restrict__ h, int64_t n, int64_t tpb) {
  const int64_t tj = blockIdx.x, ti = blockIdx.y;        // 64-row tile indices; source (tj, ti), destination (ti, tj)
  if (ti < tj) return;                                   // ti == tj: a copy-block boundary may cut a diagonal tile
  // a 64-tile spans two 32-row tiles: skip it only if all four 32x32 sub-tiles are inside a diagonal copy block
  const int64_t blk_lo_i = (2 * ti) / tpb, blk_hi_i = (2 * ti + 1) / tpb, blk_lo_j = (2 * tj) / tpb, blk_hi_j = (2 * tj + 1) / tpb;
  const bool all_inside = blk_lo_i == blk_hi_i && blk_lo_j == blk_hi_j && blk_lo_i == blk_lo_j;
  if (all_inside) return;
  __shared__ float tile[64][65];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const int64_t a = tj * 64 + i, b = ti * 64 + tx;     // source: the tile above the diagonal
    tile[i][tx] = (a < n && b < n) ? h[a * n + b] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 4) {
    const int64_t a = ti * 64 + i, b = tj * 64 + tx;
    if (a < n && b < n) {
      // inside a diagonal copy block the element was copied as it is: leave it (only whole 32x32 sub-tiles can be)
      const bool inside = (a / 32) / tpb == (b / 32) / tpb;
      if (!inside && a > b) h[a * n + b] = tile[tx][i];
    }
  }
}

// Block-upper-triangle packing of a symmetric matrix (the exchange format of the sample-sharded
// statistics, dist.allreduce_statistics): block row b (rows [b*bs, b*bs+rows)) contributes its columns
// [b*bs, n) as rows x (n - b*bs) contiguous floats, block rows one after the other -- the layout
// slk_upload_symmetric_f32 sends over PCIe.  pack: packed = scale * H; unpack: H (upper block rows) =
// scale * packed, the caller mirrors.  blockIdx.y = block row.
template <bool UNPACK>
__global__ void __launch_bounds__(256) sym_pack_kernel(float* __restrict__ h, int64_t n, int64_t bs, float scale,
                                                       float* __restrict__ packed) {
  const int64_t b = blockIdx.y, r0 = b * bs;
  const int64_t rows = (r0 + bs < n) ? bs : n - r0, width = n - r0;
  // offset of this block row in the packed buffer: sum over earlier block rows of bs * (n - b' * bs)
  const int64_t off = b * bs * n - bs * bs * (b * (b - 1) / 2);
  const int64_t total = rows * width;
  float* pk = packed + off;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t i = t / width, j = t - i * width;
    float* hp = h + (r0 + i) * n + r0 + j;
    if (UNPACK) *hp = __fmul_rn(scale, pk[t]);
    else pk[t] = __fmul_rn(scale, *hp);
  }
}

}  // namespace slk

using namespace slk;

template <typename T>
static int row_noclip_impl(const T* w, int64_t r, int64_t n, double cb_min, double cb_max, T* out, void* stream) {
  SLK_REQUIRE(r >= 0 && n >= 1, "bad shape [%lld, %lld]", (long long)r, (long long)n);
  SLK_REQUIRE(cb_min < 0 && cb_max > 0, "Codebook should have both negative and positive values.");
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(w && out, "NULL pointer");
  int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  row_noclip_scale_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(w, r, n, (T)cb_min, (T)cb_max, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

template <typename T>
static int row_rms_impl(const T* w, int64_t r, int64_t n, T* out, void* stream) {
  SLK_REQUIRE(r >= 0 && n >= 1, "bad shape");
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(w && out, "NULL pointer");
  int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  row_rms_scale_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(w, r, n, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

template <typename T>
static int remove_bias_impl(const T* h, const T* m, int64_t n, T* out, void* stream) {
  SLK_REQUIRE(n >= 0, "negative n");
  if (n == 0) return SLK_OK;
  SLK_REQUIRE(h && m && out, "NULL pointer");
  remove_bias_kernel<T><<<stream_grid(n * n, 256), 256, 0, (cudaStream_t)stream>>>(h, m, n, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

extern "C" {

int slk_abi_version(void) { return SLK_ABI_VERSION; }
const char* slk_last_error(void) { return slk::g_err; }
int64_t slk_launch_count(void) { return (int64_t)__atomic_load_n(&slk::g_launches, __ATOMIC_RELAXED); }

int slk_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  SLK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  SLK_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count_host) *sm_count_host = p.multiProcessorCount;
  if (cc_major_host) *cc_major_host = p.major;
  if (cc_minor_host) *cc_minor_host = p.minor;
  return SLK_OK;
}

int slk_round_f32(const float* x, int64_t count, const slk_codebook* cb, int mode, float* out_val,
                  void* out_idx, void* stream) {
  return round_impl<float>(x, count, cb, mode, out_val, out_idx, (cudaStream_t)stream);
}
int slk_round_f64(const double* x, int64_t count, const slk_codebook* cb, int mode, void* out_val,
                  void* out_idx, void* stream) {
  return round_impl<double>(x, count, cb, mode, out_val, out_idx, (cudaStream_t)stream);
}

int slk_scale_axis_f32(const float* x, int64_t outer, int64_t len, int64_t inner, const float* s, int mode,
                       float* out, void* stream) {
  return scale_axis_impl<float>(x, outer, len, inner, s, mode, out, (cudaStream_t)stream);
}
int slk_scale_axis_f64(const double* x, int64_t outer, int64_t len, int64_t inner, const double* s, int mode,
                       double* out, void* stream) {
  return scale_axis_impl<double>(x, outer, len, inner, s, mode, out, (cudaStream_t)stream);
}

int slk_row_noclip_scale_f32(const float* w, int64_t r, int64_t n, double cb_min, double cb_max, float* out,
                             void* stream) {
  return row_noclip_impl<float>(w, r, n, cb_min, cb_max, out, stream);
}
int slk_row_noclip_scale_f64(const double* w, int64_t r, int64_t n, double cb_min, double cb_max, double* out,
                             void* stream) {
  return row_noclip_impl<double>(w, r, n, cb_min, cb_max, out, stream);
}

int slk_row_rms_scale_f32(const float* w, int64_t r, int64_t n, float* out, void* stream) {
  return row_rms_impl<float>(w, r, n, out, stream);
}
int slk_row_rms_scale_f64(const double* w, int64_t r, int64_t n, double* out, void* stream) {
  return row_rms_impl<double>(w, r, n, out, stream);
}

int slk_remove_input_bias_f32(const float* h, const float* m, int64_t n, float* out, void* stream) {
  return remove_bias_impl<float>(h, m, n, out, stream);
}
int slk_remove_input_bias_f64(const double* h, const double* m, int64_t n, double* out, void* stream) {
  return remove_bias_impl<double>(h, m, n, out, stream);
}

int slk_damp_value_f32(const float* h, int64_t n, double damp, float* out_dampval, void* stream) {
  SLK_REQUIRE(h && out_dampval && n >= 1, "bad arguments");
  damp_value_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(h, n, (float)damp, out_dampval);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_col_resid_sums_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb, int mode, float* out,
                           void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(w && out && r >= 0 && n >= 1 && (mode == 0 || mode == 1), "bad arguments");
  col_resid_sums_kernel<<<(int)ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(w, r, n, make_grid<float>(cb), mode, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_order_keys(const float* h, int64_t n, const float* dampval, const float* colsum, double* keys,
                   void* stream) {
  SLK_REQUIRE(h && keys && n >= 1, "bad arguments");
  order_keys_kernel<<<(int)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(h, n, dampval, colsum, keys);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_argsort_f64(const double* keys, int64_t n, int64_t* order, void* stream) {
  SLK_REQUIRE(keys && order && n >= 1, "bad arguments");
  argsort_rank_kernel<<<(int)ceil_div(n, 8), 256, 0, (cudaStream_t)stream>>>(keys, n, order);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_permute_cols_f32(const float* src, int64_t r, int64_t n, const int64_t* idx, int scatter, float* dst,
                         void* stream) {
  SLK_REQUIRE(r >= 0 && n >= 0, "negative extent");
  if (r * n == 0) return SLK_OK;
  SLK_REQUIRE(src && idx && dst && src != dst, "bad pointers");
  permute_cols_kernel<<<stream_grid(r * n, 256), 256, 0, (cudaStream_t)stream>>>(src, r, n, idx, scatter, dst);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_scale_permute_cols_f32(const float* src, int64_t r, int64_t n, const int64_t* idx, const float* s, int scatter,
                               float* dst, void* stream) {
  SLK_REQUIRE(r >= 0 && n >= 0, "negative extent");
  if (r * n == 0) return SLK_OK;
  SLK_REQUIRE(src && s && dst && src != dst, "bad pointers");
  scale_permute_cols_kernel<<<stream_grid(r * n, 256), 256, 0, (cudaStream_t)stream>>>(src, r, n, idx, s, scatter, dst);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

/* Upload of a SYMMETRIC host matrix (the Hessian H = X^T X / n, statistics.py:87) moving only what is
 * needed over PCIe: the matrix is cut into block rows of `bs` rows (bs a multiple of 32); block row b
 * sends its columns [b*bs, n) -- the diagonal block and everything to its right -- as one strided
 * copy, and a device kernel mirrors the 32x32 tiles above the block diagonal into the tiles below it.
 * (nb + 1) / (2 nb) of the bytes cross the bus.  Only enqueues work (capturable in a CUDA graph). */
int slk_upload_symmetric_f32(const float* h_host, float* h_dev, int64_t n, int64_t bs, void* stream) {
  SLK_REQUIRE(h_host && h_dev && n >= 1 && bs >= 32 && bs % 32 == 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t r0 = 0; r0 < n; r0 += bs) {
    const int64_t rows = (r0 + bs < n) ? bs : n - r0;
    SLK_CUDA(cudaMemcpy2DAsync(h_dev + r0 * n + r0, (size_t)n * 4, h_host + r0 * n + r0, (size_t)n * 4,
                               (size_t)(n - r0) * 4, (size_t)rows, cudaMemcpyHostToDevice, st));
  }
  if (bs < n) {
    const unsigned nt64 = (unsigned)((n + 63) / 64);
    mirror_upper_kernel<<<dim3(nt64, nt64), 256, 0, st>>>(h_dev, n, bs / 32);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

/* Pack / unpack the block upper triangle of a symmetric device matrix (layout and byte count of
 * slk_upload_symmetric_bytes) with a scale factor; unpack also mirrors the strictly lower blocks.
 * Exchange format of the sample-sharded statistics: each rank packs count_rank / count_total * H, ONE
 * all-reduce (sum) over ~half the bytes of H runs in place on the packed buffer, unpack rebuilds H. */
int slk_sym_pack_f32(const float* h, int64_t n, int64_t bs, float scale, float* packed, void* stream) {
  SLK_REQUIRE(h && packed && n >= 1 && bs >= 32 && bs % 32 == 0, "bad arguments");
  const int64_t nb = ceil_div(n, bs);
  dim3 grid((unsigned)stream_grid(bs * n, 256, 4), (unsigned)nb);
  sym_pack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(h), n, bs, scale, packed);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_sym_unpack_f32(const float* packed, int64_t n, int64_t bs, float scale, float* h, void* stream) {
  SLK_REQUIRE(h && packed && n >= 1 && bs >= 32 && bs % 32 == 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nb = ceil_div(n, bs);
  dim3 grid((unsigned)stream_grid(bs * n, 256, 4), (unsigned)nb);
  sym_pack_kernel<true><<<grid, 256, 0, st>>>(h, n, bs, scale, const_cast<float*>(packed));
  SLK_LAUNCH_CHECK();
  if (bs < n) {
    const unsigned nt64 = (unsigned)((n + 63) / 64);
    mirror_upper_kernel<<<dim3(nt64, nt64), 256, 0, st>>>(h, n, bs / 32);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

/* The two halves of slk_upload_symmetric_f32 for callers that keep a dedicated copy stream: the DMA part
 * (block rows of the upper triangle) and the device mirror, to be enqueued on different streams with an
 * event between them, so that the copy queue never waits for a kernel. */
int slk_upload_symmetric_copy_f32(const float* h_host, float* h_dev, int64_t n, int64_t bs, void* stream) {
  SLK_REQUIRE(h_host && h_dev && n >= 1 && bs >= 32 && bs % 32 == 0, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t r0 = 0; r0 < n; r0 += bs) {
    const int64_t rows = (r0 + bs < n) ? bs : n - r0;
    SLK_CUDA(cudaMemcpy2DAsync(h_dev + r0 * n + r0, (size_t)n * 4, h_host + r0 * n + r0, (size_t)n * 4,
                               (size_t)(n - r0) * 4, (size_t)rows, cudaMemcpyHostToDevice, st));
  }
  return SLK_OK;
}

int slk_mirror_symmetric_f32(float* h_dev, int64_t n, int64_t bs, void* stream) {
  SLK_REQUIRE(h_dev && n >= 1 && bs >= 32 && bs % 32 == 0, "bad arguments");
  if (bs < n) {
    const unsigned nt64 = (unsigned)((n + 63) / 64);
    mirror_upper_kernel<<<dim3(nt64, nt64), 256, 0, (cudaStream_t)stream>>>(h_dev, n, bs / 32);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

size_t slk_upload_symmetric_bytes(int64_t n, int64_t bs) {
  size_t total = 0;
  for (int64_t r0 = 0; r0 < n; r0 += bs) total += (size_t)((r0 + bs < n) ? bs : n - r0) * (size_t)(n - r0) * 4;
  return total;
}

size_t slk_pivot_order_ws_bytes(int64_t n) {
  return (size_t)n * n * sizeof(double) + (size_t)n * sizeof(int) + sizeof(PivotState) + 512;
}

int slk_pivot_order_f64(const double* h, int64_t n, void* ws, size_t ws_bytes, int64_t* order, void* stream) {
  SLK_REQUIRE(h && order && n >= 1, "bad arguments");
  SLK_REQUIRE(ws && ws_bytes >= slk_pivot_order_ws_bytes(n), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* L = (double*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  PivotState* stp = (PivotState*)(L + n * n);
  int* active = (int*)(stp + 1);
  pivot_init_kernel<<<stream_grid(n * n, 256), 256, 0, st>>>(h, n, L, order, active);
  SLK_LAUNCH_CHECK();
  const dim3 ugrid((unsigned)ceil_div(n, 256), (unsigned)ceil_div(n, 16));
  for (int64_t k = 0; k < n; ++k) {
    pivot_select_kernel<<<1, 1024, 0, st>>>(L, n, k, order, active, stp);
    SLK_LAUNCH_CHECK();
    if (k + 1 < n) {
      pivot_update_kernel<<<ugrid, 256, 0, st>>>(L, n, active, stp);
      SLK_LAUNCH_CHECK();
    }
  }
  return SLK_OK;
}

/* host-only: the exact breakpoints the kernels use for a uniform codebook (tests) */
int slk_codebook_breaks_host(const slk_codebook* cb, float* out16_host) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(out16_host != nullptr, "NULL output");
  const GridBreaks b = make_breaks(cb);
  for (int k = 0; k < 16; ++k) out16_host[k] = b.X[k];
  return SLK_OK;
}

int slk_mean_f32(const float* v, int64_t count, float* out, void* stream) {
  SLK_REQUIRE(v && out && count >= 1, "bad arguments");
  mean_kernel_f32<<<1, 256, 0, (cudaStream_t)stream>>>(v, count, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
int slk_mean_f64(const double* v, int64_t count, double* out, void* stream) {
  SLK_REQUIRE(v && out && count >= 1, "bad arguments");
  mean_kernel_f64<<<1, 256, 0, (cudaStream_t)stream>>>(v, count, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_selftest_fastdiv_f32(const float* divisors, int32_t count, uint64_t* mismatches, void* stream) {
  SLK_REQUIRE(divisors && mismatches && count >= 1, "bad arguments");
  SLK_CUDA(cudaMemsetAsync(mismatches, 0, (size_t)count * sizeof(uint64_t), (cudaStream_t)stream));
  dim3 grid((unsigned)(sm_count() * 8), (unsigned)count);
  selftest_fastdiv_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(divisors, (unsigned long long*)mismatches, 1ull << 32);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

/* _compute_mse, H None or 1-D (scaling.py:84-95): out[row] = sum_j h[j] e[row, j]^2 (h may be NULL) */
int slk_row_wsq_f32(const float* e, const float* h, int64_t r, int64_t n, float* out, void* stream) {
  SLK_REQUIRE(e && out && r >= 1 && n >= 1, "bad arguments");
  int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  row_wsq_kernel<float, float><<<grid, 256, 0, (cudaStream_t)stream>>>(e, h, r, n, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
int slk_row_wsq_f64(const double* e, const double* h, int64_t r, int64_t n, double* out, void* stream) {
  SLK_REQUIRE(e && out && r >= 1 && n >= 1, "bad arguments");
  int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  row_wsq_kernel<double, double><<<grid, 256, 0, (cudaStream_t)stream>>>(e, h, r, n, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
/* fp32 residuals under an fp64 diagonal (the "diagN" / hessianN penalties promote H, scaling.py:222): the square
   is rounded in fp32 as np.square(E) is, the product and the sum are fp64 */
int slk_row_wsq_f32_h64(const float* e, const double* h, int64_t r, int64_t n, double* out, void* stream) {
  SLK_REQUIRE(e && h && out && r >= 1 && n >= 1, "bad arguments");
  int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  row_wsq_kernel<float, double><<<grid, 256, 0, (cudaStream_t)stream>>>(e, h, r, n, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

int slk_bias_delta_f32(const float* w, const float* wq, const float* mean, int64_t r, int64_t n, float* delta,
                       void* stream) {
  SLK_REQUIRE(w && wq && mean && delta && r >= 1 && n >= 1, "bad arguments");
  int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  bias_delta_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, wq, mean, r, n, delta);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

}  // extern "C"
