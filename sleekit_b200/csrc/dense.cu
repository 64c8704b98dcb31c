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
// screening value, so tau = 2^-5 against eps <= 1e-2 (bf16; tests/test_fullh_rank_model.py) cuts nothing that could win; typically 2-3 points
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
    cand[(int64_t)j * r + row] = bg[j];                 // sorted by ranking value: the kept ones are a prefix, the
    c += keep ? 1 : 0;                                  // next one is the best-ranked EXCLUDED point (fullh_certify)
  }
  count[row] = c;
}

// offsets[row] = pairs before the row, offsets[r] = total pairs (one CTA: r is a layer's row count), then the pair
// lists in row order.  If the rows ask for more than `capacity` pairs in total (flat error landscapes, e.g. many
// all-zero rows whose grid points all tie), every row is cut to its capacity / r best-ranked candidates, so the
// total fits and no row is left without its best ones.
__global__ void __launch_bounds__(1024) screen_pairs_kernel(const int* __restrict__ count, const int* __restrict__ cand,
                                                            int64_t r, int capacity, int* __restrict__ offsets,
                                                            int* __restrict__ pair_row, int* __restrict__ pair_g) {
  __shared__ int part[1024];
  __shared__ int s_limit;
  const int t = threadIdx.x;
  const int64_t per = (r + 1023) / 1024;
  const int64_t lo = t * per, hi = lo + per < r ? lo + per : r;
  const int per_row = (int)(capacity / r) > 0 ? (int)(capacity / r) : 1;
  int limit = 0x7fffffff;
  for (int pass = 0; pass < 2; ++pass) {
    int s = 0;
    for (int64_t i = lo; i < hi; ++i) s += min(count[i], limit);
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {        // inclusive scan of the per-thread sums
      const int v = t >= d ? part[t - d] : 0;
      __syncthreads();
      part[t] += v;
      __syncthreads();
    }
    if (t == 0) s_limit = part[1023] > capacity ? per_row : limit;
    __syncthreads();
    const bool again = s_limit != limit;
    limit = s_limit;
    if (!again) break;                          // uniform: s_limit is the same for every thread
    __syncthreads();                            // part[] is rewritten by the second pass
  }
  int run = t ? part[t - 1] : 0;
  for (int64_t i = lo; i < hi; ++i) {
    const int c = min(count[i], limit);
    offsets[i] = run;
    for (int k = 0; k < c; ++k) {
      pair_row[run + k] = (int)i;
      pair_g[run + k] = cand[(int64_t)k * r + i];
    }
    run += c;
  }
  if (t == 1023) offsets[r] = part[1023];
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

// Certificate of the screened search: every grid point that was NOT evaluated exactly has a ranking value >= `next`
// (the best-ranked excluded one; the last list entry when the 16-entry list itself was full).  With a ranking error
// of at most eps relative, such a point's exact error is >= next - eps * |next|; if that is not below the row's
// exact minimum, the minimum over the candidates is the minimum over all G points.  Rows for which this cannot be
// shown are counted (0 on every tested input; a caller that sees a non-zero count can rerun with "fullh_topk" 0).
__global__ void __launch_bounds__(256) fullh_certify_kernel(const float* __restrict__ rank, const int* __restrict__ cand,
                                                            const int* __restrict__ offsets, int64_t r, int G, int topk_max,
                                                            float eps, const float* __restrict__ best_err,
                                                            int* __restrict__ uncertified) {
  int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  const int kept = offsets[row + 1] - offsets[row];
  if (kept == 0 || kept >= G) return;                    // nothing finite to choose from / everything evaluated
  const int slot = kept < topk_max ? kept : topk_max - 1;
  const int gp = cand[(int64_t)slot * r + row];
  if (gp < 0) return;                                    // fewer finite ranking values than list entries: all evaluated
  const float next = rank[(int64_t)gp * r + row];
  if (__fsub_rn(next, __fmul_rn(eps, fabsf(next))) < best_err[row]) atomicAdd(uncertified, 1);
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

// Workspace of the full-H search; one layout for the size query (ws == nullptr) and the entry point.
struct FullhWs {
  int chunk;                // grid points per exact GEMM launch
  int64_t tiles;            // row-dot partials per row of the CUDA-core GEMM
  void *resid, *part, *errs;
  float *init, *best_err, *best_f;
  float *rhi, *rlo, *hhi, *hlo;                  // tensor-core path: TF32 parts of the residuals and of H
  float *errs_all, *part_s;                      // screening: ranking values [G, r], their row-dot partials
  int *cand, *best_g, *offsets, *pair_row, *pair_g;
  __nv_bfloat16* hbf;
  size_t bytes;
};

static FullhWs fullh_layout(void* ws, int64_t r, int64_t n, int G, int h_dtype) {
  FullhWs L = {};
  const size_t elem = h_dtype == 2 ? 8 : 4;
  L.chunk = fullh_chunk(r, n, G, elem);
  L.tiles = h_dtype == 2 ? rowdot_tiles<double>(n) : rowdot_tiles<float>(n);
  size_t off = 0;
  auto take = [&](size_t b) {
    char* p = ws ? (char*)ws + off : nullptr;
    off += align256(b);
    return p;
  };
  const size_t chunk = (size_t)L.chunk, pairs = (size_t)FULLH_TOPK_MAX * r + 2;
  L.resid = take(chunk * r * n * elem);
  L.part = take(chunk * r * L.tiles * elem);
  L.errs = take(chunk * r * elem);
  L.init = (float*)take((size_t)r * 4);
  L.best_err = (float*)take((size_t)r * 4);
  L.best_f = (float*)take((size_t)r * 4);
  if (h_dtype == 1 && n % 4 == 0) {
    L.rhi = (float*)take(chunk * r * n * 4);     // rhi | rlo: also the 2 * chunk grid points of a ranking pass
    L.rlo = (float*)take(chunk * r * n * 4);
    L.hhi = (float*)take((size_t)n * n * 4);
    L.hlo = (float*)take((size_t)n * n * 4);
    L.errs_all = (float*)take((size_t)G * r * 4);
    L.cand = (int*)take((size_t)FULLH_TOPK_MAX * r * 4);
    L.best_g = (int*)take((size_t)r * 4);
    L.part_s = (float*)take(2 * chunk * r * ceil_div(n, TC_TILE_N) * 4);
    L.hbf = (__nv_bfloat16*)take((size_t)n * n * 2);
    L.offsets = (int*)take(pairs * 4);
    L.pair_row = (int*)take(pairs * 4);
    L.pair_g = (int*)take(pairs * 4);
  }
  L.bytes = off;
  return L;
}

struct FullhCtx {
  const float* w; int64_t r, n; DevGrid<float> g; const float* factors; int G; const void* h; int h_dtype;
  bool tc; cudaStream_t st; FullhWs L;
  int* uncertified;          // optional device counter of rows without a certificate (fullh_certify_kernel)
};

static inline int resid_blocks(int64_t pairs) {    // one CTA per (grid point, row) pair, grid-stride beyond 64 per SM
  const int64_t cap = (int64_t)sm_count() * 64;
  return (int)(pairs < cap ? pairs : cap);
}

// errs[0 .. rows) = e H e^T of the fp32 residual rows in L.resid / rhi / rlo: the fp32-faithful product with the
// row dot in its epilogue (H symmetric: its own K-major B operand), partials reduced in fixed order.  m_limit:
// device-side row count (row tiles beyond it exit).
static int fullh_exact_errors(const FullhCtx& c, int64_t rows, const int* m_limit) {
  const int64_t tiles = ceil_div(c.n, TC_TILE_N);
  TcParams tp;
  tp.C = (float*)c.L.part; tp.ldc = tiles; tp.R = (const float*)c.L.resid; tp.R2 = nullptr; tp.ldr = c.n;
  tp.M = rows; tp.N = c.n; tp.K = c.n; tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
  tp.m_limit = m_limit;
  int rc = tc_gemm_presplit_f32(TC_ROWDOT, c.L.rhi, c.L.rlo, c.n, c.L.hhi, c.L.hlo, c.n, tp, c.st);
  if (rc) return rc;
  rowdot_reduce_kernel<float><<<(int)ceil_div(rows, 256), 256, 0, c.st>>>((const float*)c.L.part, rows, tiles,
                                                                          (float*)c.L.errs);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

// 1. ranking values of all G grid points into L.errs_all [G, r]: bf16 residuals against a bf16 H (kind::f16), or,
//    when n % 8 != 0, TF32-rounded residuals against the truncated H (one kind::tf32 pass); 2 * chunk grid
//    points per launch, in the rhi | rlo region
static int fullh_rank(const FullhCtx& c) {
  const FullhWs& L = c.L;
  const int bn = g_opt_fullh_bn == 128 ? 128 : 256, ctas = (int)g_opt_fullh_ctas;
  const bool bf16 = g_opt_fullh_bf16 != 0 && c.n % 8 == 0;
  const int64_t tiles_s = ceil_div(c.n, bn);
  const int chunk_s = 2 * L.chunk < c.G ? 2 * L.chunk : c.G;
  if (bf16) {
    to_bf16_kernel<<<resid_blocks(ceil_div(c.n * c.n, 256)), 256, 0, c.st>>>((const float*)c.h, c.n * c.n, L.hbf);
    SLK_LAUNCH_CHECK();
  }
  for (int g0 = 0; g0 < c.G; g0 += chunk_s) {
    const int gc = (c.G - g0) < chunk_s ? (c.G - g0) : chunk_s;
    const int64_t rows = (int64_t)gc * c.r;
    if (bf16)
      grid_resid_kernel<float, 2><<<resid_blocks(rows), 128, 0, c.st>>>(c.w, c.r, c.n, c.g, c.factors, g0, gc, L.init, L.rhi);
    else
      grid_resid_kernel<float, 1><<<resid_blocks(rows), 128, 0, c.st>>>(c.w, c.r, c.n, c.g, c.factors, g0, gc, L.init, L.rhi);
    SLK_LAUNCH_CHECK();
    TcParams tp;
    tp.C = L.part_s; tp.ldc = tiles_s; tp.R = L.rhi; tp.R2 = nullptr; tp.ldr = c.n;
    tp.M = rows; tp.N = c.n; tp.K = c.n; tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
    int rc = bf16 ? tc_gemm_screen_bf16(L.rhi, c.n, L.hbf, c.n, tp, bn, ctas, c.st)
                  : tc_gemm_screen_f32(L.rhi, c.n, L.hhi, c.n, tp, bn, ctas, c.st);
    if (rc) return rc;
    rowdot_reduce_kernel<float><<<(int)ceil_div(rows, 256), 256, 0, c.st>>>(L.part_s, rows, tiles_s,
                                                                            L.errs_all + (int64_t)g0 * c.r);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

// 2a. compacted candidates (screen_select_kernel): at most topk * r (row, grid point) pairs evaluated exactly by one
//     product whose row tiles beyond the device-side pair count exit at once
static int fullh_eval_pairs(const FullhCtx& c, int topk) {
  const FullhWs& L = c.L;
  const int cap = (int)((int64_t)topk * c.r);
  int* count = L.best_g;
  screen_select_kernel<FULLH_TOPK_MAX><<<(int)ceil_div(c.r, 128), 128, 0, c.st>>>(L.errs_all, c.r, c.G, 0.03125f, L.cand, count);
  SLK_LAUNCH_CHECK();
  screen_pairs_kernel<<<1, 1024, 0, c.st>>>(count, L.cand, c.r, cap, L.offsets, L.pair_row, L.pair_g);
  SLK_LAUNCH_CHECK();
  const int* npairs = L.offsets + c.r;
  grid_resid_kernel<float><<<resid_blocks(cap), 128, 0, c.st>>>(c.w, c.r, c.n, c.g, c.factors, 0, topk, L.init, (float*)L.resid,
                                                                L.rhi, L.rlo, L.pair_g, L.pair_row, npairs);
  SLK_LAUNCH_CHECK();
  int rc = fullh_exact_errors(c, cap, npairs);
  if (rc) return rc;
  pairs_argmin_kernel<<<(int)ceil_div(c.r, 256), 256, 0, c.st>>>((const float*)L.errs, L.pair_g, L.offsets, c.r, c.factors,
                                                                 L.best_err, L.best_f);
  SLK_LAUNCH_CHECK();
  if (c.uncertified) {
    fullh_certify_kernel<<<(int)ceil_div(c.r, 256), 256, 0, c.st>>>(L.errs_all, L.cand, L.offsets, c.r, c.G, FULLH_TOPK_MAX,
                                                                    0.015625f, L.best_err, c.uncertified);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

// 2b. a fixed topk candidates per row, `chunk` slots per product (also the path of layers whose residuals do not
//     fit topk at a time)
static int fullh_eval_slots(const FullhCtx& c, int topk) {
  const FullhWs& L = c.L;
  const int rb = (int)ceil_div(c.r, 128);
  if (topk == 4) screen_topk_kernel<4><<<rb, 128, 0, c.st>>>(L.errs_all, c.r, c.G, L.cand);
  else if (topk == 8) screen_topk_kernel<8><<<rb, 128, 0, c.st>>>(L.errs_all, c.r, c.G, L.cand);
  else screen_topk_kernel<16><<<rb, 128, 0, c.st>>>(L.errs_all, c.r, c.G, L.cand);
  SLK_LAUNCH_CHECK();
  fill_int_kernel<<<(int)ceil_div(c.r, 256), 256, 0, c.st>>>(L.best_g, -1, c.r);
  SLK_LAUNCH_CHECK();
  for (int s0 = 0; s0 < topk; s0 += L.chunk) {
    const int sc = (topk - s0) < L.chunk ? (topk - s0) : L.chunk;
    const int64_t rows = (int64_t)sc * c.r;
    const int* cs = L.cand + (int64_t)s0 * c.r;
    grid_resid_kernel<float><<<resid_blocks(rows), 128, 0, c.st>>>(c.w, c.r, c.n, c.g, c.factors, 0, sc, L.init, (float*)L.resid,
                                                                   L.rhi, L.rlo, cs);
    SLK_LAUNCH_CHECK();
    int rc = fullh_exact_errors(c, rows, nullptr);
    if (rc) return rc;
    cand_argmin_kernel<<<(int)ceil_div(c.r, 256), 256, 0, c.st>>>((const float*)L.errs, cs, c.r, sc, c.factors, L.best_err,
                                                                  L.best_f, L.best_g);
    SLK_LAUNCH_CHECK();
  }
  if (c.uncertified) {                                   // no certificate on this path: -1 = not checked
    fill_int_kernel<<<1, 256, 0, c.st>>>(c.uncertified, -1, 1);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

// every grid point evaluated exactly, `chunk` points per product: the reference's loop (scaling.py:125-134)
static int fullh_all_points(const FullhCtx& c) {
  const FullhWs& L = c.L;
  const int rb = (int)ceil_div(c.r, 256);
  for (int g0 = 0; g0 < c.G; g0 += L.chunk) {
    const int gc = (c.G - g0) < L.chunk ? (c.G - g0) : L.chunk;
    const int64_t rows = (int64_t)gc * c.r;
    const int blocks = resid_blocks(rows);
    int rc;
    if (c.h_dtype == 1) {
      grid_resid_kernel<float><<<blocks, 128, 0, c.st>>>(c.w, c.r, c.n, c.g, c.factors, g0, gc, L.init, (float*)L.resid,
                                                         c.tc ? L.rhi : nullptr, c.tc ? L.rlo : nullptr);
      SLK_LAUNCH_CHECK();
      if (c.tc) {
        rc = fullh_exact_errors(c, rows, nullptr);
        if (rc) return rc;
      } else {
        GemmParams<float> p = gemm_params<float>((const float*)L.resid, c.n, (const float*)c.h, c.n, (float*)L.part, 0, rows,
                                                 c.n, c.n);
        rc = gemm_launch<float, false, false, EPI_ROWDOT>(p, 1, c.st);
        if (rc) return rc;
        rowdot_reduce_kernel<float><<<(int)ceil_div(rows, 256), 256, 0, c.st>>>((const float*)L.part, rows, L.tiles,
                                                                                (float*)L.errs);
        SLK_LAUNCH_CHECK();
      }
      grid_argmin_kernel<float><<<rb, 256, 0, c.st>>>((const float*)L.errs, c.r, g0, gc, c.factors, L.best_err, L.best_f);
    } else {
      grid_resid_kernel<double><<<blocks, 128, 0, c.st>>>(c.w, c.r, c.n, c.g, c.factors, g0, gc, L.init, (double*)L.resid);
      SLK_LAUNCH_CHECK();
      GemmParams<double> p = gemm_params<double>((const double*)L.resid, c.n, (const double*)c.h, c.n, (double*)L.part, 0,
                                                 rows, c.n, c.n);
      rc = gemm_launch<double, false, false, EPI_ROWDOT>(p, 1, c.st);
      if (rc) return rc;
      rowdot_reduce_kernel<double><<<(int)ceil_div(rows, 256), 256, 0, c.st>>>((const double*)L.part, rows, L.tiles,
                                                                               (double*)L.errs);
      SLK_LAUNCH_CHECK();
      grid_argmin_kernel<double><<<rb, 256, 0, c.st>>>((const double*)L.errs, c.r, g0, gc, c.factors, L.best_err, L.best_f);
    }
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
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
  return fullh_layout(nullptr, r, n, G, h_dtype).bytes;
}

int slk_scale_search_fullh_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb, const float* factors,
                               int32_t G, const void* h, int32_t h_dtype, void* ws, size_t ws_bytes,
                               float* out_scale, float* out_err, void* stream) {
  return slk_scale_search_fullh_checked_f32(w, r, n, cb, factors, G, h, h_dtype, ws, ws_bytes, out_scale, out_err, nullptr,
                                            stream);
}

int slk_scale_search_fullh_checked_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb, const float* factors,
                                       int32_t G, const void* h, int32_t h_dtype, void* ws, size_t ws_bytes,
                                       float* out_scale, float* out_err, int32_t* uncertified_rows, void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(w && factors && h && out_scale && r >= 1 && n >= 1 && G >= 1, "bad arguments");
  SLK_REQUIRE(h_dtype == 1 || h_dtype == 2, "h_dtype %d", h_dtype);
  FullhCtx c;
  c.L = fullh_layout(ws, r, n, G, h_dtype);
  SLK_REQUIRE(ws && ws_bytes >= c.L.bytes, "workspace too small");
  c.w = w; c.r = r; c.n = n; c.g = make_grid<float>(cb); c.factors = factors; c.G = G; c.h = h; c.h_dtype = h_dtype;
  c.st = (cudaStream_t)stream;
  c.uncertified = uncertified_rows;
  if (uncertified_rows) SLK_CUDA(cudaMemsetAsync(uncertified_rows, 0, sizeof(int32_t), c.st));
  // tensor-core path (fp32 H): residuals are written with their TF32 parts, H is split once
  c.tc = c.L.rhi != nullptr && n >= 32 && tc_gemm_usable(c.L.resid, n, h, n);
  if (c.tc) {
    rc = tc_split_f32((const float*)h, n, n, n, n, c.L.hhi, c.L.hlo, c.st);
    if (rc) return rc;
  }
  rc = slk_row_noclip_scale_f32(w, r, n, cb->lo, cb->hi, c.L.init, stream);
  if (rc) return rc;
  fill_inf_kernel<<<(int)ceil_div(r, 256), 256, 0, c.st>>>(c.L.best_err, c.L.best_f, r);
  SLK_LAUNCH_CHECK();
  const int topk = c.tc ? (int)g_opt_fullh_topk : 0;
  if ((topk == 4 || topk == 8 || topk == 16) && G >= 4 * topk) {
    rc = fullh_rank(c);
    if (rc) return rc;
    const bool compact = g_opt_fullh_compact && topk <= c.L.chunk && (int64_t)topk * r < (1ll << 30);
    rc = compact ? fullh_eval_pairs(c, topk) : fullh_eval_slots(c, topk);
  } else {
    rc = fullh_all_points(c);
  }
  if (rc) return rc;
  finish_scale_kernel<<<(int)ceil_div(r, 256), 256, 0, c.st>>>(c.L.init, c.L.best_f, c.L.best_err, r, out_scale, out_err);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

}  // extern "C"
