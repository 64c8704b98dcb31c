// K1 / K6 / K7-init: the dense contractions of the hot path, on the shared GEMM core.
//   slk_hessian_accum_f32      Sleekit.add_batch                 statistics.py:76-87
//   slk_hweighted_error_*      channelwise_error / _compute_mse  obq.py:89-95, scaling.py:91-95
//   slk_gain_*                 compute_gain                      obq.py:220-231
//   slk_scale_search_fullh_f32 compute_min_mse_scaling, 2-D H    scaling.py:98-134
#include "gemm.cuh"
#include "tc_gemm.cuh"

#include <cuda_bf16.h>

// slk_set_option("fullh_topk", 0 | 4 | 8 | 16): candidates per row of the screened full-H search (0: every grid
// point is evaluated exactly); ("fullh_bn", 128 | 256): tile width of the screening product
// ("fullh_bf16", 0 | 1): screening product on TF32 (one kind::tf32 pass) or BF16 (kind::f16) operands
int64_t g_opt_fullh_topk = 8;
int64_t g_opt_fullh_bn = 256;
int64_t g_opt_fullh_bf16 = 1;
int64_t g_opt_fullh_ctas = 2;      // ("fullh_ctas", 1 | 2): CTAs per SM of the screening product
int64_t g_opt_fullh_compact = 1;   // ("fullh_compact", 0 | 1): candidates within 2^-5 of the best-ranked one only, compacted

namespace slk {

// out[m] = sum over column tiles of part[m, t], fixed order -> deterministic
template <typename T>
__global__ void __launch_bounds__(256) rowdot_reduce_kernel(const T* __restrict__ part, int64_t rows, int64_t tiles,
                                                            T* __restrict__ out) {
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  T s = (T)0;
  for (int64_t t = 0; t < tiles; ++t) s += part[m * tiles + t];
  out[m] = s;
}

template <typename T>
static int hweighted_error_impl(const T* w, const T* q, const T* h, int64_t r, int64_t n, void* ws,
                                size_t ws_bytes, T* out, cudaStream_t st) {
  SLK_REQUIRE(r >= 0 && n >= 1, "bad shape");
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(w && h && out, "NULL pointer");
  const int64_t tiles = rowdot_tiles<T>(n);
  SLK_REQUIRE(ws && ws_bytes >= (size_t)(r * tiles) * sizeof(T), "workspace too small");
  GemmParams<T> p = gemm_params<T>(w, n, h, n, (T*)ws, 0, r, n, n);
  p.A2 = q;
  int rc = gemm_launch<T, false, false, EPI_ROWDOT>(p, 1, st);
  if (rc) return rc;
  rowdot_reduce_kernel<T><<<(int)ceil_div(r, 256), 256, 0, st>>>((const T*)ws, r, tiles, out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

template <typename T>
static int gain_impl(const T* w, const T* q, const T* h, const T* cand, int64_t r, int64_t n, T* out,
                     cudaStream_t st) {
  SLK_REQUIRE(w && q && h && cand && out && r >= 1 && n >= 1, "bad arguments");
  // acc = (q - w) @ H ; the epilogue applies obq.py:231 with diag(H) read in place (stride n+1)
  GemmParams<T> p = gemm_params<T>(q, n, h, n, out, n, r, n, n);
  p.A2 = w;
  p.x0 = cand; p.x1 = q; p.ldx = n;
  p.x2 = h; p.x2_stride = n + 1;
  return gemm_launch<T, false, false, EPI_GAIN>(p, 1, st);
}

__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ x, int64_t S, int64_t n, int64_t ldx,
                                                      float* __restrict__ mean, float keep, float count) {
  // 32 columns x 32 row groups per CTA: coalesced 128-byte row segments, fp64 partial sums per row
  // group, combined in a fixed order (deterministic), rounded once (statistics.py:84-85)
  __shared__ double part[32][33];
  const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + cx;
  double acc = 0.0;
  if (j < n)
    for (int64_t i = rg; i < S; i += 32) acc += (double)__ldg(x + i * ldx + j);
  part[rg][cx] = acc;
  __syncthreads();
  if (rg == 0 && j < n) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += part[k][cx];
    mean[j] = __fadd_rn(__fmul_rn(mean[j], keep), __fdiv_rn((float)t, count));
  }
}

// ---- full-H scale search pieces --------------------------------------------
// resid[(g - g0) * r + row, j] = descale(quant(w / (f_g * init))) - w   (scaling.py:128-130 -> 73-80)
// cand != nullptr: slot gi of a row evaluates grid point cand[gi * r + row] instead of g0 + gi (< 0: empty slot,
// zeros).  SCREEN: resid receives the residual rounded to the nearest TF32 value (operand of the one-pass product).
// One CTA per (grid point, row) pair at a time: scale and reciprocal once per pair, float4 traffic when the row
// pitch allows it (the pass is then bound by its stores, not by index arithmetic).
// SCREEN: 0 exact residuals (+ TF32 parts), 1 residuals rounded to TF32, 2 residuals rounded to bf16 (resid then
// points to a bf16 matrix of the same shape)
// The three divides (by the scale, by the codebook step, by the reciprocal scale) go through the correctly
// rounded reciprocal scheme of common.cuh: same results as the IEEE divides, a quarter of the instructions.
struct ResidDiv { FastDivF scale, rs, step; };

template <int SCREEN>
__device__ __forceinline__ float grid_resid_value(const DevGrid<float>& g, float x, const ResidDiv& d, bool live) {
  float e = 0.0f;
  if (live) {
    const float v = g.kind == 0 ? uniform_value_fast(g, d.step, fastdiv(x, d.scale)) : grid_value(g, __fdiv_rn(x, d.scale.d));
    e = __fsub_rn(fastdiv(v, d.rs), x);
  }
  return SCREEN == 1 ? __uint_as_float((__float_as_uint(e) + 0x1000u) & 0xffffe000u) : e;
}

template <typename TE, int SCREEN = 0>
__global__ void __launch_bounds__(128) grid_resid_kernel(const float* __restrict__ w, int64_t r, int64_t n,
                                                         DevGrid<float> g, const float* __restrict__ factors,
                                                         int g0, int gcount, const float* __restrict__ init,
                                                         TE* __restrict__ resid, float* __restrict__ rhi = nullptr,
                                                         float* __restrict__ rlo = nullptr,
                                                         const int* __restrict__ cand = nullptr,
                                                         const int* __restrict__ pair_row = nullptr,
                                                         const int* __restrict__ pair_count = nullptr) {
  // pair_row != nullptr: a compacted list of *pair_count (row, grid point) pairs: pair_row[i], cand[i]
  const int64_t pairs = pair_row ? (int64_t)__ldg(pair_count) : (int64_t)gcount * r;
  const bool vec = sizeof(TE) == 4 && (n & 3) == 0 && (((uintptr_t)w | (uintptr_t)resid | (uintptr_t)rhi | (uintptr_t)rlo) & 15) == 0;
  for (int64_t pr = blockIdx.x; pr < pairs; pr += gridDim.x) {
    const int gi = (int)(pr / r);
    const int64_t row = pair_row ? (int64_t)__ldg(pair_row + pr) : pr - (int64_t)gi * r;
    const int gp = cand ? __ldg(cand + pr) : g0 + gi;
    const bool live = gp >= 0;
    float scale = 1.0f;
    if (live) scale = __fmul_rn(__ldg(factors + gp), __ldg(init + row));
    ResidDiv dv;
    dv.scale = make_fastdiv(scale);
    dv.rs = make_fastdiv(dv.scale.y);             // y = RN(1 / scale): the reciprocal scale of scaling.py:80
    dv.step = make_fastdiv(g.step);
    const float* wr = w + row * n;
    TE* out = resid + pr * n;
    if (vec) {
      const float4* w4 = reinterpret_cast<const float4*>(wr);
      for (int64_t j = threadIdx.x; j < (n >> 2); j += blockDim.x) {
        const float4 x = __ldg(w4 + j);
        float4 e;
        e.x = grid_resid_value<SCREEN>(g, x.x, dv, live);
        e.y = grid_resid_value<SCREEN>(g, x.y, dv, live);
        e.z = grid_resid_value<SCREEN>(g, x.z, dv, live);
        e.w = grid_resid_value<SCREEN>(g, x.w, dv, live);
        if (SCREEN == 2) {
          __nv_bfloat162 lo2 = __floats2bfloat162_rn(e.x, e.y), hi2 = __floats2bfloat162_rn(e.z, e.w);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo2);
          pk.y = *reinterpret_cast<uint32_t*>(&hi2);
          reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(resid) + pr * n)[j] = pk;
          continue;
        }
        reinterpret_cast<float4*>(out)[j] = e;
        if (!SCREEN && rhi) {                    // TF32 parts for the tensor-core product
          float4 hh, ll;
          split_tf32(e.x, hh.x, ll.x);
          split_tf32(e.y, hh.y, ll.y);
          split_tf32(e.z, hh.z, ll.z);
          split_tf32(e.w, hh.w, ll.w);
          reinterpret_cast<float4*>(rhi + pr * n)[j] = hh;
          reinterpret_cast<float4*>(rlo + pr * n)[j] = ll;
        }
      }
    } else {
      for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const float e = grid_resid_value<SCREEN>(g, __ldg(wr + j), dv, live);
        if (SCREEN == 2) {
          reinterpret_cast<__nv_bfloat16*>(resid)[pr * n + j] = __float2bfloat16_rn(e);
          continue;
        }
        out[j] = (TE)e;
        if (!SCREEN && rhi) {
          float hh, ll;
          split_tf32(e, hh, ll);
          rhi[pr * n + j] = hh;
          rlo[pr * n + j] = ll;
        }
      }
    }
  }
}

// strict '<', grid order, best error kept in fp32 (scaling.py:125-134)
template <typename TE>
__global__ void __launch_bounds__(256) grid_argmin_kernel(const TE* __restrict__ err, int64_t r, int g0, int gcount,
                                                          const float* __restrict__ factors,
                                                          float* __restrict__ best_err, float* __restrict__ best_f) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  float be = best_err[row], bf = best_f[row];
  for (int gi = 0; gi < gcount; ++gi) {
    TE e = err[(int64_t)gi * r + row];
    if (e < (TE)be) { be = (float)e; bf = __ldg(factors + g0 + gi); }
  }
  best_err[row] = be;
  best_f[row] = bf;
}

__global__ void __launch_bounds__(256) fill_inf_kernel(float* a, float* b, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = __int_as_float(0x7f800000); b[i] = __int_as_float(0x7f800000); }
}

// ---- screened full-H search ---------------------------------------------------------------------------------
// The reference evaluates e_g H e_g^T for every grid point g and keeps the first minimum (scaling.py:125-134).
// Only the arg-min matters, so all G points are first ranked by a single kind::tf32 pass (~1e-3 relative, the
// error largely common to the grid points of a row: same H, same weights), and only the TOPK best-ranked points
// of each row are evaluated by the fp32-faithful 3xTF32 product -- the same kernel and the same per-row
// arithmetic as the unscreened path, so the chosen scale and its error are bit-identical whenever the true
// minimum is among the TOPK candidates (neighbouring grid points differ by 1e-3..1e-2 relative around the
// minimum; TOPK = 8 of 100 leaves a wide margin, test_fullh_screening_*).
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ x, int64_t total, __nv_bfloat16* __restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}

template <int TOPK>
__global__ void __launch_bounds__(128) screen_topk_kernel(const float* __restrict__ err, int64_t r, int G,
                                                          int* __restrict__ cand) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  float be[TOPK];
  int bg[TOPK];
#pragma unroll
  for (int j = 0; j < TOPK; ++j) { be[j] = __int_as_float(0x7f800000); bg[j] = -1; }
  for (int gi = 0; gi < G; ++gi) {
    const float e = err[(int64_t)gi * r + row];
    if (e < be[TOPK - 1]) {   // strict: on ties the lower grid index stays; inf / NaN never enter (as in grid_argmin_kernel)
      be[TOPK - 1] = e; bg[TOPK - 1] = gi;
#pragma unroll
      for (int j = TOPK - 1; j > 0; --j) {
        const bool sw = be[j] < be[j - 1];
        const float te = sw ? be[j - 1] : be[j];
        const int tg = sw ? bg[j - 1] : bg[j];
        be[j - 1] = sw ? be[j] : be[j - 1];
        bg[j - 1] = sw ? bg[j] : bg[j - 1];
        be[j] = te; bg[j] = tg;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < TOPK; ++j) cand[(int64_t)j * r + row] = bg[j];
}

// Compacted candidates: a row keeps the (at most TOPK) best-ranked grid points whose screening error lies within
// `tau` of its best one -- with a ranking error eps, the true minimum is within (1 + eps) / (1 - eps) of the best
// screening value, so tau = 2^-5 against eps <= 5e-3 (bf16) cuts nothing that could win; typically 2-3 points
// remain of 8.  count[row] = kept candidates (>= 1 when the row has a finite error at all).
template <int TOPK>
__global__ void __launch_bounds__(128) screen_select_kernel(const float* __restrict__ err, int64_t r, int G, float tau,
                                                            int* __restrict__ cand, int* __restrict__ count) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  float be[TOPK];
  int bg[TOPK];
#pragma unroll
  for (int j = 0; j < TOPK; ++j) { be[j] = __int_as_float(0x7f800000); bg[j] = -1; }
  for (int gi = 0; gi < G; ++gi) {
    const float e = err[(int64_t)gi * r + row];
    if (e < be[TOPK - 1]) {
      be[TOPK - 1] = e; bg[TOPK - 1] = gi;
#pragma unroll
      for (int j = TOPK - 1; j > 0; --j) {
        const bool sw = be[j] < be[j - 1];
        const float te = sw ? be[j - 1] : be[j];
        const int tg = sw ? bg[j - 1] : bg[j];
        be[j - 1] = sw ? be[j] : be[j - 1];
        bg[j - 1] = sw ? bg[j] : bg[j - 1];
        be[j] = te; bg[j] = tg;
      }
    }
  }
  const float thr = __fadd_rn(be[0], __fmul_rn(tau, fabsf(be[0])));
  int c = 0;
#pragma unroll
  for (int j = 0; j < TOPK; ++j) {
    const bool keep = bg[j] >= 0 && (j == 0 || be[j] <= thr);
    cand[(int64_t)j * r + row] = keep ? bg[j] : -1;     // sorted by screening error: the kept ones are a prefix
    c += keep ? 1 : 0;
  }
  count[row] = c;
}

// offsets[row] = min(capacity, sum of count[0..row)), offsets[r] = total pairs (one CTA: r is a layer's row count),
// then the pair lists in row order
__global__ void __launch_bounds__(1024) screen_pairs_kernel(const int* __restrict__ count, const int* __restrict__ cand,
                                                            int64_t r, int capacity, int* __restrict__ offsets,
                                                            int* __restrict__ pair_row, int* __restrict__ pair_g) {
  __shared__ int part[1024];
  const int t = threadIdx.x;
  const int64_t per = (r + 1023) / 1024;
  const int64_t lo = t * per, hi = lo + per < r ? lo + per : r;
  int s = 0;
  for (int64_t i = lo; i < hi; ++i) s += count[i];
  part[t] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {        // inclusive scan of the per-thread sums
    const int v = t >= d ? part[t - d] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = t ? part[t - 1] : 0;
  for (int64_t i = lo; i < hi; ++i) {
    const int b = run < capacity ? run : capacity;
    run += count[i];
    const int e = run < capacity ? run : capacity;
    offsets[i] = b;
    for (int k = b; k < e; ++k) {
      pair_row[k] = (int)i;
      pair_g[k] = cand[(int64_t)(k - b) * r + i];
    }
  }
  if (t == 1023) offsets[r] = part[1023] < capacity ? part[1023] : capacity;
}

__global__ void __launch_bounds__(256) pairs_argmin_kernel(const float* __restrict__ err, const int* __restrict__ pair_g,
                                                           const int* __restrict__ offsets, int64_t r,
                                                           const float* __restrict__ factors, float* __restrict__ best_err,
                                                           float* __restrict__ best_f) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  float be = best_err[row];
  int bg = -1;
  for (int i = offsets[row]; i < offsets[row + 1]; ++i) {
    const float e = err[i];
    const int gp = pair_g[i];
    if (e < be || (e == be && bg >= 0 && gp < bg)) { be = e; bg = gp; }
  }
  best_err[row] = be;
  if (bg >= 0) best_f[row] = __ldg(factors + bg);
}

// first minimum in grid order over the evaluated candidates (strict '<'; equal errors: the lower grid index)
__global__ void __launch_bounds__(256) cand_argmin_kernel(const float* __restrict__ err, const int* __restrict__ cand,
                                                          int64_t r, int slots, const float* __restrict__ factors,
                                                          float* __restrict__ best_err, float* __restrict__ best_f,
                                                          int* __restrict__ best_g) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  float be = best_err[row];
  int bg = best_g[row];
  for (int s = 0; s < slots; ++s) {
    const int gp = cand[(int64_t)s * r + row];
    if (gp < 0) continue;
    const float e = err[(int64_t)s * r + row];
    if (e < be || (e == be && bg >= 0 && gp < bg)) { be = e; bg = gp; }
  }
  best_err[row] = be;
  best_g[row] = bg;
  if (bg >= 0) best_f[row] = __ldg(factors + bg);
}

__global__ void __launch_bounds__(256) fill_int_kernel(int* a, int v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

__global__ void __launch_bounds__(256) finish_scale_kernel(const float* __restrict__ init, const float* __restrict__ best_f,
                                                           const float* __restrict__ best_err, int64_t r,
                                                           float* __restrict__ out_scale, float* __restrict__ out_err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  out_scale[i] = __fmul_rn(init[i], best_f[i]);
  if (out_err) out_err[i] = best_err[i];
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

constexpr int FULLH_TOPK_MAX = 16;

// how many grid points are evaluated per GEMM launch
static inline int fullh_chunk(int64_t r, int64_t n, int G, size_t elem) {
  const size_t budget = (size_t)1 << 30;  // <= 1 GiB of residuals per chunk
  int64_t per = (int64_t)(budget / ((size_t)r * n * elem));
  if (per < 1) per = 1;
  if (per > G) per = G;
  return (int)per;
}

}  // namespace slk

using namespace slk;

extern "C" {

size_t slk_hweighted_error_ws_bytes(int64_t r, int64_t n, int32_t elem_bytes) {
  int64_t tiles = elem_bytes == 8 ? rowdot_tiles<double>(n) : rowdot_tiles<float>(n);
  size_t bytes = align256((size_t)(r * tiles) * (size_t)elem_bytes + 256);
  if (elem_bytes == 4) bytes += tc_gemm_ws_bytes(r, n, n);   // hi/lo operand copies of the tensor-core path
  return bytes;
}

int slk_hweighted_error_f32(const float* w, const float* q, const float* h, int64_t r, int64_t n, void* ws,
                            size_t ws_bytes, float* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (r >= 1 && n >= 32 && w && h && out && ws && tc_gemm_usable(w, n, h, n) && (!q || (uintptr_t)q % 16 == 0) &&
      ws_bytes >= slk_hweighted_error_ws_bytes(r, n, 4)) {
    // tensor-core path: ((W-Q) @ H) via tcgen05 (H symmetric: H is its own K-major B operand),
    // the row dot with (W-Q) fused in the epilogue, then the fixed-order reduce over column tiles
    const int64_t tiles = ceil_div(n, TC_TILE_N);
    const size_t part_bytes = align256((size_t)(r * rowdot_tiles<float>(n)) * 4 + 256);
    TcParams p;
    p.C = (float*)ws; p.ldc = tiles; p.R = w; p.R2 = q; p.ldr = n; p.M = r; p.N = n; p.K = n;
    p.alpha = 1.0f; p.keep = 0.0f; p.count = 1.0f; p.error_flag = nullptr;
    int rc = tc_gemm_f32(TC_ROWDOT, w, q, n, h, n, p, (char*)ws + part_bytes, ws_bytes - part_bytes, st);
    if (rc) return rc;
    rowdot_reduce_kernel<float><<<(int)ceil_div(r, 256), 256, 0, st>>>((const float*)ws, r, tiles, out);
    SLK_LAUNCH_CHECK();
    return SLK_OK;
  }
  return hweighted_error_impl<float>(w, q, h, r, n, ws, ws_bytes, out, st);
}
int slk_hweighted_error_f64(const double* w, const double* q, const double* h, int64_t r, int64_t n, void* ws,
                            size_t ws_bytes, double* out, void* stream) {
  return hweighted_error_impl<double>(w, q, h, r, n, ws, ws_bytes, out, (cudaStream_t)stream);
}

int slk_gain_f32(const float* w, const float* q, const float* h, const float* cand, int64_t r, int64_t n,
                 float* out, void* stream) {
  return gain_impl<float>(w, q, h, cand, r, n, out, (cudaStream_t)stream);
}
int slk_gain_f64(const double* w, const double* q, const double* h, const double* cand, int64_t r, int64_t n,
                 double* out, void* stream) {
  return gain_impl<double>(w, q, h, cand, r, n, out, (cudaStream_t)stream);
}

size_t slk_hessian_accum_ws_bytes(int64_t S, int64_t n) { return tc_gemm_at_ws_bytes(n, S) + 256; }

int slk_hessian_accum_f32(const float* x, int64_t S, int64_t n, int64_t ldx, float* hess, float* mean, double keep,
                          double new_count, void* ws, size_t ws_bytes, void* stream) {
  SLK_REQUIRE(x && hess && mean && S >= 1 && n >= 1 && ldx >= n, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  colsum_kernel<<<(int)ceil_div(n, 32), 1024, 0, st>>>(x, S, n, ldx, mean, (float)keep, (float)new_count);
  SLK_LAUNCH_CHECK();
  if (n >= 64 && S >= 32 && ws && ws_bytes >= slk_hessian_accum_ws_bytes(S, n) && tc_gemm_usable(hess, 4, hess, 4)) {
    // tensor-core path: X is split and transposed once into K-major hi/lo [n, S]; both operands of
    // X^T X read that same copy; H = H*keep + D/count fused in the epilogue (statistics.py:82-87)
    TcParams p;
    p.C = hess; p.ldc = n; p.R = nullptr; p.R2 = nullptr; p.ldr = 0; p.M = n; p.N = n; p.K = S;
    p.alpha = 1.0f; p.keep = (float)keep; p.count = (float)new_count; p.error_flag = nullptr;
    return tc_gemm_at_f32(TC_HESS_SYM, x, ldx, p, ws, ws_bytes, st);
  }
  // H = H*keep + X^T X / count : A = X^T (stored [K=S, M=n]), B = X ([K=S, N=n])
  GemmParams<float> p = gemm_params<float>(x, ldx, x, ldx, hess, n, n, n, S);
  p.keep = (float)keep;
  p.count = (float)new_count;
  return gemm_launch<float, true, false, EPI_HESS>(p, 1, st);
}

size_t slk_scale_search_fullh_ws_bytes(int64_t r, int64_t n, int32_t G, int32_t h_dtype) {
  size_t elem = h_dtype == 2 ? 8 : 4;
  int chunk = fullh_chunk(r, n, G, elem);
  int64_t tiles = h_dtype == 2 ? rowdot_tiles<double>(n) : rowdot_tiles<float>(n);
  size_t bytes = 0;
  bytes += align256((size_t)chunk * r * n * elem);       // residuals
  bytes += align256((size_t)chunk * r * tiles * elem);   // row-dot partials
  bytes += align256((size_t)chunk * r * elem);           // errors
  bytes += 3 * align256((size_t)r * sizeof(float));      // init, best_err, best_f
  if (h_dtype == 1 && n % 4 == 0) {                      // tensor-core path: TF32 parts of the residuals and of H
    bytes += 2 * align256((size_t)chunk * r * n * 4) + 2 * align256((size_t)n * n * 4);
    // screening: errors of all grid points, candidate lists, best index, row-dot partials of 2 * chunk points
    bytes += align256((size_t)G * r * 4) + align256((size_t)FULLH_TOPK_MAX * r * 4) + align256((size_t)r * 4) +
             align256((size_t)2 * chunk * r * ceil_div(n, TC_TILE_N) * 4) + align256((size_t)n * n * 2) +
             3 * align256((size_t)(FULLH_TOPK_MAX * r + 2) * 4);   // row counts / offsets, pair rows, pair grid points
  }
  return bytes;
}

int slk_scale_search_fullh_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb, const float* factors,
                               int32_t G, const void* h, int32_t h_dtype, void* ws, size_t ws_bytes,
                               float* out_scale, float* out_err, void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(w && factors && h && out_scale && r >= 1 && n >= 1 && G >= 1, "bad arguments");
  SLK_REQUIRE(h_dtype == 1 || h_dtype == 2, "h_dtype %d", h_dtype);
  SLK_REQUIRE(ws && ws_bytes >= slk_scale_search_fullh_ws_bytes(r, n, G, h_dtype), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t elem = h_dtype == 2 ? 8 : 4;
  const int chunk = fullh_chunk(r, n, G, elem);
  const int64_t tiles = h_dtype == 2 ? rowdot_tiles<double>(n) : rowdot_tiles<float>(n);
  char* base = (char*)ws;
  void* resid = base; base += align256((size_t)chunk * r * n * elem);
  void* part = base; base += align256((size_t)chunk * r * tiles * elem);
  void* errs = base; base += align256((size_t)chunk * r * elem);
  float* init = (float*)base; base += align256((size_t)r * sizeof(float));
  float* best_err = (float*)base; base += align256((size_t)r * sizeof(float));
  float* best_f = (float*)base; base += align256((size_t)r * sizeof(float));
  // tensor-core path (fp32 H): residuals are written with their TF32 parts, H is split once
  const bool tc = h_dtype == 1 && n % 4 == 0 && n >= 32 && tc_gemm_usable(resid, n, h, n);
  float *rhi = nullptr, *rlo = nullptr, *hhi = nullptr, *hlo = nullptr;
  if (tc) {
    rhi = (float*)base; base += align256((size_t)chunk * r * n * 4);
    rlo = (float*)base; base += align256((size_t)chunk * r * n * 4);
    hhi = (float*)base; base += align256((size_t)n * n * 4);
    hlo = (float*)base; base += align256((size_t)n * n * 4);
    rc = tc_split_f32((const float*)h, n, n, n, n, hhi, hlo, st);
    if (rc) return rc;
  }
  const int topk = tc ? (int)g_opt_fullh_topk : 0;
  const bool screen = tc && (topk == 4 || topk == 8 || topk == 16) && G >= 4 * topk;
  float *errs_all = nullptr, *part_s = nullptr;
  int *cand = nullptr, *best_g = nullptr;
  __nv_bfloat16* hbf = nullptr;
  int *offsets = nullptr, *pair_row = nullptr, *pair_g = nullptr;
  if (screen) {
    errs_all = (float*)base; base += align256((size_t)G * r * 4);
    cand = (int*)base; base += align256((size_t)FULLH_TOPK_MAX * r * 4);
    best_g = (int*)base; base += align256((size_t)r * 4);
    part_s = (float*)base; base += align256((size_t)2 * chunk * r * ceil_div(n, TC_TILE_N) * 4);
    hbf = (__nv_bfloat16*)base; base += align256((size_t)n * n * 2);
    offsets = (int*)base; base += align256((size_t)(FULLH_TOPK_MAX * r + 2) * 4);
    pair_row = (int*)base; base += align256((size_t)(FULLH_TOPK_MAX * r + 2) * 4);
    pair_g = (int*)base;
  }

  rc = slk_row_noclip_scale_f32(w, r, n, cb->lo, cb->hi, init, stream);
  if (rc) return rc;
  fill_inf_kernel<<<(int)ceil_div(r, 256), 256, 0, st>>>(best_err, best_f, r);
  SLK_LAUNCH_CHECK();
  DevGrid<float> g = make_grid<float>(cb);
  auto resid_blocks = [](int64_t pairs) {           // one CTA per (grid point, row) pair, grid-stride beyond 64 per SM
    const int64_t cap = (int64_t)sm_count() * 64;
    return (int)(pairs < cap ? pairs : cap);
  };
  if (screen) {
    // 1. rank all grid points with one TF32 pass: TF32-rounded residuals (one array over the rhi | rlo region,
    //    2 * chunk grid points at a time) against the truncated H
    const int bn = g_opt_fullh_bn == 128 ? 128 : 256;
    const bool bf16 = g_opt_fullh_bf16 != 0 && n % 8 == 0;
    const int64_t tiles_s = ceil_div(n, bn);
    const int chunk_s = 2 * chunk < G ? 2 * chunk : G;
    if (bf16) {
      to_bf16_kernel<<<resid_blocks(ceil_div(n * n, 256)), 256, 0, st>>>((const float*)h, n * n, hbf);
      SLK_LAUNCH_CHECK();
    }
    for (int g0 = 0; g0 < G; g0 += chunk_s) {
      const int gc = (G - g0) < chunk_s ? (G - g0) : chunk_s;
      const int64_t rows = (int64_t)gc * r;
      if (bf16) grid_resid_kernel<float, 2><<<resid_blocks(rows), 128, 0, st>>>(w, r, n, g, factors, g0, gc, init, rhi);
      else grid_resid_kernel<float, 1><<<resid_blocks(rows), 128, 0, st>>>(w, r, n, g, factors, g0, gc, init, rhi);
      SLK_LAUNCH_CHECK();
      TcParams tp;
      tp.C = part_s; tp.ldc = tiles_s; tp.R = rhi; tp.R2 = nullptr; tp.ldr = n;
      tp.M = rows; tp.N = n; tp.K = n; tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
      rc = bf16 ? tc_gemm_screen_bf16(rhi, n, hbf, n, tp, bn, (int)g_opt_fullh_ctas, st)
                : tc_gemm_screen_f32(rhi, n, hhi, n, tp, bn, (int)g_opt_fullh_ctas, st);
      if (rc) return rc;
      rowdot_reduce_kernel<float><<<(int)ceil_div(rows, 256), 256, 0, st>>>(part_s, rows, tiles_s, errs_all + (int64_t)g0 * r);
      SLK_LAUNCH_CHECK();
    }
    if (g_opt_fullh_compact && topk <= chunk && (int64_t)topk * r < (1ll << 30)) {
      // 2a. compacted candidates (see screen_select_kernel): capacity topk * r pairs, GEMM row tiles beyond the
      //     device-side pair count exit at once
      const int cap = (int)((int64_t)topk * r);
      int* count = best_g;
      screen_select_kernel<FULLH_TOPK_MAX><<<(int)ceil_div(r, 128), 128, 0, st>>>(errs_all, r, G, 0.03125f, cand, count);
      SLK_LAUNCH_CHECK();
      screen_pairs_kernel<<<1, 1024, 0, st>>>(count, cand, r, cap, offsets, pair_row, pair_g);
      SLK_LAUNCH_CHECK();
      grid_resid_kernel<float><<<resid_blocks(cap), 128, 0, st>>>(w, r, n, g, factors, 0, topk, init, (float*)resid, rhi, rlo,
                                                                  pair_g, pair_row, offsets + r);
      SLK_LAUNCH_CHECK();
      TcParams tp;
      tp.C = (float*)part; tp.ldc = ceil_div(n, TC_TILE_N); tp.R = (const float*)resid; tp.R2 = nullptr; tp.ldr = n;
      tp.M = cap; tp.N = n; tp.K = n; tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
      tp.m_limit = offsets + r;
      rc = tc_gemm_presplit_f32(TC_ROWDOT, rhi, rlo, n, hhi, hlo, n, tp, st);
      if (rc) return rc;
      rowdot_reduce_kernel<float><<<(int)ceil_div(cap, 256), 256, 0, st>>>((const float*)part, cap, ceil_div(n, TC_TILE_N),
                                                                          (float*)errs);
      SLK_LAUNCH_CHECK();
      pairs_argmin_kernel<<<(int)ceil_div(r, 256), 256, 0, st>>>((const float*)errs, pair_g, offsets, r, factors, best_err,
                                                                 best_f);
      SLK_LAUNCH_CHECK();
      finish_scale_kernel<<<(int)ceil_div(r, 256), 256, 0, st>>>(init, best_f, best_err, r, out_scale, out_err);
      SLK_LAUNCH_CHECK();
      return SLK_OK;
    }
    if (topk == 4) screen_topk_kernel<4><<<(int)ceil_div(r, 128), 128, 0, st>>>(errs_all, r, G, cand);
    else if (topk == 8) screen_topk_kernel<8><<<(int)ceil_div(r, 128), 128, 0, st>>>(errs_all, r, G, cand);
    else screen_topk_kernel<16><<<(int)ceil_div(r, 128), 128, 0, st>>>(errs_all, r, G, cand);
    SLK_LAUNCH_CHECK();
    fill_int_kernel<<<(int)ceil_div(r, 256), 256, 0, st>>>(best_g, -1, r);
    SLK_LAUNCH_CHECK();
    // 2. the candidates, exactly as the unscreened path evaluates a grid point
    for (int s0 = 0; s0 < topk; s0 += chunk) {
      const int sc = (topk - s0) < chunk ? (topk - s0) : chunk;
      const int64_t rows = (int64_t)sc * r;
      const int* cs = cand + (int64_t)s0 * r;
      grid_resid_kernel<float><<<resid_blocks(rows), 128, 0, st>>>(w, r, n, g, factors, 0, sc, init, (float*)resid,
                                                                      rhi, rlo, cs);
      SLK_LAUNCH_CHECK();
      TcParams tp;
      tp.C = (float*)part; tp.ldc = ceil_div(n, TC_TILE_N); tp.R = (const float*)resid; tp.R2 = nullptr; tp.ldr = n;
      tp.M = rows; tp.N = n; tp.K = n; tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
      rc = tc_gemm_presplit_f32(TC_ROWDOT, rhi, rlo, n, hhi, hlo, n, tp, st);
      if (rc) return rc;
      rowdot_reduce_kernel<float><<<(int)ceil_div(rows, 256), 256, 0, st>>>((const float*)part, rows, ceil_div(n, TC_TILE_N),
                                                                           (float*)errs);
      SLK_LAUNCH_CHECK();
      cand_argmin_kernel<<<(int)ceil_div(r, 256), 256, 0, st>>>((const float*)errs, cs, r, sc, factors, best_err, best_f,
                                                                best_g);
      SLK_LAUNCH_CHECK();
    }
  }
  for (int g0 = 0; g0 < G && !screen; g0 += chunk) {
    const int gc = (G - g0) < chunk ? (G - g0) : chunk;
    const int64_t rows = (int64_t)gc * r;
    const int blocks = resid_blocks(rows);
    if (h_dtype == 1) {
      grid_resid_kernel<float><<<blocks, 128, 0, st>>>(w, r, n, g, factors, g0, gc, init, (float*)resid, rhi, rlo);
      SLK_LAUNCH_CHECK();
      if (tc) {
        // (E H) . E per row on tcgen05: H symmetric, hence its own K-major B operand; row dot in the epilogue
        TcParams tp;
        tp.C = (float*)part; tp.ldc = ceil_div(n, TC_TILE_N); tp.R = (const float*)resid; tp.R2 = nullptr; tp.ldr = n;
        tp.M = rows; tp.N = n; tp.K = n; tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
        rc = tc_gemm_presplit_f32(TC_ROWDOT, rhi, rlo, n, hhi, hlo, n, tp, st);
      } else {
        GemmParams<float> p = gemm_params<float>((const float*)resid, n, (const float*)h, n, (float*)part, 0, rows, n, n);
        rc = gemm_launch<float, false, false, EPI_ROWDOT>(p, 1, st);
      }
      if (rc) return rc;
      rowdot_reduce_kernel<float><<<(int)ceil_div(rows, 256), 256, 0, st>>>((const float*)part, rows,
                                                                           tc ? ceil_div(n, TC_TILE_N) : tiles, (float*)errs);
      SLK_LAUNCH_CHECK();
      grid_argmin_kernel<float><<<(int)ceil_div(r, 256), 256, 0, st>>>((const float*)errs, r, g0, gc, factors, best_err, best_f);
    } else {
      grid_resid_kernel<double><<<blocks, 128, 0, st>>>(w, r, n, g, factors, g0, gc, init, (double*)resid);
      SLK_LAUNCH_CHECK();
      GemmParams<double> p = gemm_params<double>((const double*)resid, n, (const double*)h, n, (double*)part, 0, rows, n, n);
      rc = gemm_launch<double, false, false, EPI_ROWDOT>(p, 1, st);
      if (rc) return rc;
      rowdot_reduce_kernel<double><<<(int)ceil_div(rows, 256), 256, 0, st>>>((const double*)part, rows, tiles, (double*)errs);
      SLK_LAUNCH_CHECK();
      grid_argmin_kernel<double><<<(int)ceil_div(r, 256), 256, 0, st>>>((const double*)errs, r, g0, gc, factors, best_err, best_f);
    }
    SLK_LAUNCH_CHECK();
  }
  finish_scale_kernel<<<(int)ceil_div(r, 256), 256, 0, st>>>(init, best_f, best_err, r, out_scale, out_err);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

}  // extern "C"
