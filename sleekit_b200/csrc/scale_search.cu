// K5: fused scale-grid search for the MSE and diagonal-H criteria.
//   compute_min_mse_scaling (H None or 1-D)        scaling.py:98-134
//   -> compute_non_saturating_scaling              scaling.py:44-55
//   -> quantize_with_scaling / apply_scaling        scaling.py:58-81, 21-25
//   -> _compute_mse                                 scaling.py:84-90
// The reference makes G (=100) passes over W, each ~12 numpy passes.  Here one CTA owns one
// row: the row is read from HBM once into shared memory, every thread owns one grid point and
// walks the whole row (shared-memory broadcast reads, no cross-thread reduction), so HBM sees
// 4 bytes per weight in total and the kernel is bound by the fp32 pipe (three IEEE divides per
// weight and grid point).  Every arithmetic step is the reference's, separately rounded:
//   scale = f*init; rs = 1/scale; x = w/scale; v = quant(x); dq = v/rs; e = dq - w; err += h*e*e
#include "common.cuh"

#include <stdlib.h>

namespace slk {

// Error of one row at one grid point.  FAST: the three divides (by scale, by the codebook step,
// by rs) go through fastdiv_core -- same correctly rounded quotients, 5 FMA-pipe ops each.
template <typename HT, bool HAS_H, bool FAST>
__device__ __forceinline__ HT row_error(const float* __restrict__ rowp, int64_t n, const DevGrid<float>& g,
                                        const FastDivF& fs, const FastDivF& fr, const FastDivF& fstep,
                                        const HT* __restrict__ hdiag) {
  const float hi = (float)(g.size - 1);
  HT acc[4] = {(HT)0, (HT)0, (HT)0, (HT)0};
  auto one = [&](float w) -> float {
    float v;
    if (FAST) {
      const float x = fastdiv_core(w, fs.d, fs.y);                          // scaling.py:73
      float k = rintf(fastdiv_core(__fsub_rn(x, g.zero), fstep.d, fstep.y)); // codebook.py:60-62
      k = k < 0.0f ? 0.0f : k;
      k = k > hi ? hi : k;
      v = __fadd_rn(__fmul_rn(k, g.step), g.zero);                          // codebook.py:63-64
      v = fastdiv_core(v, fr.d, fr.y);                                      // scaling.py:80
    } else {
      v = __fdiv_rn(grid_value(g, __fdiv_rn(w, fs.d)), fr.d);
    }
    const float e = __fsub_rn(v, w);                                        // scaling.py:130
    return __fmul_rn(e, e);
  };
  int64_t j = 0;
  for (; j + 4 <= n; j += 4) {
    float e2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) e2[u] = one(rowp[j + u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] += HAS_H ? (HT)e2[u] * __ldg(hdiag + j + u) : (HT)e2[u];
  }
  for (; j < n; ++j) {
    const float e2 = one(rowp[j]);
    acc[0] += HAS_H ? (HT)e2 * __ldg(hdiag + j) : (HT)e2;
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// ---- threshold-table form (uniform codebooks of <= 16 entries, fp32 errors) ------------------------
// For a fixed (row, grid point) the chain  w -> w/scale -> (x - zero)/step -> rint -> clip  is a
// composition of monotone non-decreasing correctly rounded operations, hence the code index is a
// non-decreasing step function of w with C-1 breakpoints.  Per row and grid point the breakpoints
// T_k = min{w : idx(w) >= k} are found EXACTLY by bisection over the ordered fp32 values using the
// reference's own op chain, and the C de-quantised values (k*step + zero)/rs are computed once with
// that same chain; both go to shared memory.  The main loop is then transposed with respect to the
// direct kernel: a thread owns 8 WEIGHTS (in registers) and walks the grid points; per weight and
// grid point it forms a biased index estimate with one FMA (k_a in {k-1, k}: the bias 0.5 exceeds
// the estimate's error by four orders of magnitude), looks T[k_a+1] up, corrects, looks the value
// up, and accumulates h*(v-w)^2:  ~13 issue slots instead of the ~36 of three exact divides, split
// over the FMA, ALU and LSU pipes.  Every weight receives exactly the value the reference's chain
// gives it (non-NaN inputs), so the chosen grid point is unchanged.
constexpr int TAB_MAXC = 16;
constexpr int TAB_PITCH = 18;     // T[0..C] then padding; all lanes read one grid point's table -> distinct banks
constexpr int TAB_THREADS = 128;
constexpr int TAB_MAXG = 128;    // grid points per row held in shared memory (23 KB per CTA: 9 CTAs per SM)

__device__ __forceinline__ int f32_ord(float x) {
  const int i = __float_as_int(x);
  return i >= 0 ? i : (int)(0x80000000u - (unsigned)i);
}
__device__ __forceinline__ float f32_unord(int o) {
  return __int_as_float(o >= 0 ? o : (int)(0x80000000u - (unsigned)o));
}

struct TabSmem {
  float T[TAB_MAXG][TAB_PITCH];   // T[g][k], k = 1..C-1 breakpoints, T[g][C] = +inf
  float D[TAB_MAXG][TAB_PITCH];   // D[g][k] = de-quantised value of code k
  float ea[TAB_MAXG], eb[TAB_MAXG];     // index estimate t_a = w*ea + eb  (eb already holds the -0.5 bias)
  float scale[TAB_MAXG], rs[TAB_MAXG];
  float errs[TAB_THREADS / 32][TAB_MAXG];
  float X[TAB_MAXC];              // X[k] = min{x : slot(x) >= k}: breakpoints of the codebook itself
  float red_lo[4], red_hi[4];
  float init;
  int bad;                         // some grid point of this row cannot use the estimate -> direct form
};

template <bool HAS_H, int TAB_WPT>
__global__ void __launch_bounds__(TAB_THREADS) scale_search_tab_kernel(const float* __restrict__ w, int64_t r, int64_t n,
                                                                       DevGrid<float> g, GridBreaks brk, float cb_min,
                                                                       float cb_max, const float* __restrict__ factors,
                                                                       int G, const float* __restrict__ hdiag,
                                                                       float* __restrict__ out_scale,
                                                                       float* __restrict__ out_err,
                                                                       float* __restrict__ out_init) {
  extern __shared__ __align__(16) unsigned char tab_raw[];
  TabSmem& sm = *reinterpret_cast<TabSmem*>(tab_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = g.size;
  const float top = (float)(C - 1);
  const FastDivF fstep = make_fastdiv(g.step);

  // breakpoints of the codebook in the scaled domain (exact; found on the host, make_breaks)
  if (tid < TAB_MAXC) sm.X[tid] = brk.X[tid];
  __syncthreads();

  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    const float* grow = w + row * n;
    // ---- pass 1: range of the row -> non-saturating scale (scaling.py:44-55) ---------------------
    float lo = __ldg(grow), hi = lo;
    for (int64_t j = tid; j < n; j += TAB_THREADS) {
      const float v = __ldg(grow + j);
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if (lane == 0) { sm.red_lo[warp] = lo; sm.red_hi[warp] = hi; }
    if (tid == 0) sm.bad = 0;
    __syncthreads();
    lo = fminf(fminf(sm.red_lo[0], sm.red_lo[1]), fminf(sm.red_lo[2], sm.red_lo[3]));
    hi = fmaxf(fmaxf(sm.red_hi[0], sm.red_hi[1]), fmaxf(sm.red_hi[2], sm.red_hi[3]));
    float init;
    {
      const float a = __fdiv_rn(hi, cb_max), b = __fdiv_rn(lo, cb_min);  // scaling.py:53
      const float s = a > b ? a : b;
      init = s > 1.0e-16f ? s : 1.0e-16f;                                // scaling.py:54
    }
    const float wmax = fmaxf(fabsf(lo), fabsf(hi));
    // ---- pass 2a: per grid point constants -----------------------------------------------------------
    for (int gi = tid; gi < G; gi += TAB_THREADS) {
      const float scale = __fmul_rn(__ldg(factors + gi), init);          // scaling.py:128
      const float rsv = __fdiv_rn(1.0f, scale);                          // scaling.py:80
      sm.scale[gi] = scale;
      sm.rs[gi] = rsv;
      const float ea = __fdiv_rn(1.0f, __fmul_rn(scale, g.step));
      const float eb = __fsub_rn(__fdiv_rn(-g.zero, g.step), 0.5f);
      sm.ea[gi] = ea;
      sm.eb[gi] = eb;
      const float reach = __fadd_rn(__fmul_rn(wmax, fabsf(ea)), fabsf(eb));
      if (!(scale > 0.0f) || !(reach < 2.0e6f)) sm.bad = 1;             // also catches NaN / inf
      sm.T[gi][C] = __int_as_float(0x7f800000);
      for (int wp = 0; wp < TAB_THREADS / 32; ++wp) sm.errs[wp][gi] = 0.0f;
    }
    __syncthreads();
    const bool direct = sm.bad != 0;
    if (direct) {
      // rare: the direct op chain, one grid point per thread (same code as scale_search_kernel)
      for (int gi = tid; gi < G; gi += TAB_THREADS) {
        const FastDivF fs = make_fastdiv(sm.scale[gi]), fr = make_fastdiv(sm.rs[gi]);
        sm.errs[0][gi] = (fstep.ok && fs.ok && fr.ok)
                             ? row_error<float, HAS_H, true>(grow, n, g, fs, fr, fstep, hdiag)
                             : row_error<float, HAS_H, false>(grow, n, g, fs, fr, fstep, hdiag);
      }
    } else {
      // ---- pass 2b: exact tables, one (grid point, code) task at a time ----------------------------
      // idx(w) >= k  <=>  RN(w/scale) >= X[k], so T[k] = min{w : RN(w/scale) >= X[k]}: it lies within a
      // few ulps of X[k]*scale; bracket, verify, bisect (full range if the bracket ever fails).
      for (int task = tid; task < G * C; task += TAB_THREADS) {
        const int k = task / G, gi = task - k * G;          // heavy and light codes spread over the threads
        const float scale = sm.scale[gi];
        const FastDivF fs = make_fastdiv(scale), fr = make_fastdiv(sm.rs[gi]);
        const float v = __fadd_rn(__fmul_rn((float)k, g.step), g.zero);                                 // codebook.py:63-64
        sm.D[gi][k] = fr.ok ? fastdiv_core(v, fr.d, fr.y) : __fdiv_rn(v, fr.d);                         // scaling.py:80
        if (k == 0) continue;
        const float xk = sm.X[k];
        auto reaches = [&](float x0) -> bool {
          return (fs.ok ? fastdiv_core(x0, fs.d, fs.y) : __fdiv_rn(x0, fs.d)) >= xk;                    // scaling.py:73
        };
        const int og = f32_ord(__fmul_rn(xk, scale));
        long long blo = (long long)og - 4, bhi = (long long)og + 4;
        if (blo < f32_ord(-3.402823466e+38f) || bhi > f32_ord(3.402823466e+38f) ||
            reaches(f32_unord((int)blo)) || !reaches(f32_unord((int)bhi))) {
          blo = f32_ord(-3.402823466e+38f);
          bhi = f32_ord(3.402823466e+38f);
        }
        while (bhi - blo > 1) {
          const long long mid = blo + ((bhi - blo) >> 1);
          if (reaches(f32_unord((int)mid))) bhi = mid; else blo = mid;
        }
        sm.T[gi][k] = f32_unord((int)bhi);
      }
      __syncthreads();
      // ---- pass 2c: weights in registers, walk the grid points -----------------------------------------
      for (int64_t c0 = 0; c0 < n; c0 += (int64_t)TAB_THREADS * TAB_WPT) {
        float wv[TAB_WPT], hv[TAB_WPT];
#pragma unroll
        for (int m = 0; m < TAB_WPT; ++m) {
          const int64_t j = c0 + tid + (int64_t)m * TAB_THREADS;
          const bool in = j < n;
          wv[m] = in ? __ldg(grow + j) : 0.0f;
          hv[m] = in ? (HAS_H ? __ldg(hdiag + j) : 1.0f) : 0.0f;
        }
#pragma unroll 2
        for (int gi = 0; gi < G; ++gi) {
          const float ea = sm.ea[gi], eb = sm.eb[gi];
          const float* Tg = sm.T[gi];
          const float* Dg = sm.D[gi];
          float part = 0.0f;
#pragma unroll
          for (int m = 0; m < TAB_WPT; ++m) {
            const float ta = __fmaf_rn(wv[m], ea, eb);
            const float u = __fadd_rn(ta, 12582912.0f);                    // integer part lands in the low mantissa bits
            int ka = __float_as_int(u) - 0x4B400000;
            ka = __vimin_s32_relu(ka, C - 1);                               // clip to [0, C-1]
            const int kx = (wv[m] >= Tg[ka + 1]) ? ka + 1 : ka;             // exact index
            const float e = __fsub_rn(Dg[kx], wv[m]);                       // scaling.py:130
            part = __fmaf_rn(__fmul_rn(e, e), hv[m], part);
          }
          part = warp_sum(part);
          if (lane == 0) sm.errs[warp][gi] += part;
        }
      }
    }
    __syncthreads();
    // ---- pass 3: first strict minimum in grid order, best kept in fp32 (scaling.py:125-134) --------
    if (tid == 0) {
      float best = __int_as_float(0x7f800000), pick = __int_as_float(0x7f800000);
      for (int gi = 0; gi < G; ++gi) {
        const float e = direct ? sm.errs[0][gi]
                               : __fadd_rn(__fadd_rn(sm.errs[0][gi], sm.errs[1][gi]), __fadd_rn(sm.errs[2][gi], sm.errs[3][gi]));
        if (e < best) { best = e; pick = __ldg(factors + gi); }
      }
      out_scale[row] = __fmul_rn(init, pick);
      if (out_err) out_err[row] = best;
      if (out_init) out_init[row] = init;
    }
    __syncthreads();
  }
}

template <typename HT, bool HAS_H>
__global__ void __launch_bounds__(128) scale_search_kernel(const float* __restrict__ w, int64_t r, int64_t n,
                                                           DevGrid<float> g, float cb_min, float cb_max,
                                                           const float* __restrict__ factors, int G,
                                                           const HT* __restrict__ hdiag, int row_in_smem,
                                                           float* __restrict__ out_scale, float* __restrict__ out_err,
                                                           float* __restrict__ out_init) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HT* errs = (HT*)smem_raw;                                   // [G]
  float* srow = (float*)(smem_raw + (((size_t)G * sizeof(HT) + 15) & ~(size_t)15));  // [n] if row_in_smem
  __shared__ float red_lo[4], red_hi[4];
  __shared__ float s_init;
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);

  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    const float* grow = w + row * n;
    // ---- pass 1: stage the row, find its range -------------------------------
    float lo = __ldg(grow), hi = lo;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      float v = __ldg(grow + j);
      if (row_in_smem) srow[j] = v;
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { red_lo[threadIdx.x >> 5] = lo; red_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
        lo = red_lo[k] < lo ? red_lo[k] : lo;
        hi = red_hi[k] > hi ? red_hi[k] : hi;
      }
      float a = __fdiv_rn(hi, cb_max), b = __fdiv_rn(lo, cb_min);  // scaling.py:53
      float s = a > b ? a : b;
      s_init = s > 1.0e-16f ? s : 1.0e-16f;                        // scaling.py:54
    }
    __syncthreads();
    const float init = s_init;
    const float* rowp = row_in_smem ? srow : grow;

    // ---- pass 2: one grid point per thread ------------------------------------
    for (int gi = threadIdx.x; gi < G; gi += blockDim.x) {
      const float scale = __fmul_rn(__ldg(factors + gi), init);    // scaling.py:128
      const float rs = __fdiv_rn(1.0f, scale);                     // scaling.py:80
      const FastDivF fs = make_fastdiv(scale), fr = make_fastdiv(rs);
      if (g.kind == 0 && fstep.ok && fs.ok && fr.ok)
        errs[gi] = row_error<HT, HAS_H, true>(rowp, n, g, fs, fr, fstep, hdiag);
      else
        errs[gi] = row_error<HT, HAS_H, false>(rowp, n, g, fs, fr, fstep, hdiag);
    }
    __syncthreads();
    // ---- pass 3: first strict minimum in grid order, best kept in fp32 (scaling.py:125-134)
    if (threadIdx.x == 0) {
      float best = __int_as_float(0x7f800000), pick = __int_as_float(0x7f800000);
      for (int gi = 0; gi < G; ++gi) {
        HT e = errs[gi];
        if (e < (HT)best) { best = (float)e; pick = __ldg(factors + gi); }
      }
      out_scale[row] = __fmul_rn(init, pick);
      if (out_err) out_err[row] = best;
      if (out_init) out_init[row] = init;
    }
    __syncthreads();
  }
}

}  // namespace slk

using namespace slk;

static int g_search_direct = -1;

/* development aid / tests: 1 forces the direct op-chain kernel, 0 the default (threshold tables) */
extern "C" int slk_debug_scale_search_direct(int on) {
  g_search_direct = on ? 1 : 0;
  return SLK_OK;
}

extern "C" int slk_scale_search_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb,
                                    const float* factors, int32_t G, const void* hdiag, int32_t h_dtype,
                                    float* out_scale, float* out_err, float* out_init, void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(r >= 0 && n >= 1 && G >= 1, "bad shape r=%lld n=%lld G=%d", (long long)r, (long long)n, G);
  SLK_REQUIRE(cb->lo < 0 && cb->hi > 0, "Codebook should have both negative and positive values.");
  SLK_REQUIRE(h_dtype >= 0 && h_dtype <= 2 && (h_dtype == 0) == (hdiag == nullptr), "hdiag / h_dtype mismatch");
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(w && factors && out_scale, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = h_dtype == 2 ? 8 : 4;
  const size_t err_bytes = ((size_t)G * esz + 15) & ~(size_t)15;
  size_t smem = err_bytes + (size_t)n * 4;
  int row_in_smem = 1;
  if (smem > 200 * 1024) { smem = err_bytes; row_in_smem = 0; }
  SLK_REQUIRE(smem <= 200 * 1024, "grid of %d points does not fit in shared memory", G);
  const int grid = (int)(r < (int64_t)sm_count() * 16 ? r : (int64_t)sm_count() * 16);
  DevGrid<float> g = make_grid<float>(cb);
  const float cmin = (float)cb->lo, cmax = (float)cb->hi;
  if (g_search_direct < 0) {   // SLK_SEARCH_TABLE=0 forces the direct op chain (A/B testing)
    const char* ev = getenv("SLK_SEARCH_TABLE");
    g_search_direct = (ev && ev[0] == '0') ? 1 : 0;
  }
  const int steps_off = g_search_direct;
  if (!steps_off && cb->kind == 0 && cb->size <= TAB_MAXC && G <= TAB_MAXG && h_dtype != 2) {
    // threshold-table form; grid: every SM holds several CTAs so that one CTA's (latency-bound) table
    // construction overlaps the others' (throughput-bound) main loops
    const int tgrid = (int)(r < (int64_t)sm_count() * 12 ? r : (int64_t)sm_count() * 12);
    const GridBreaks brk = make_breaks(cb);
    // weights per thread and chunk: the candidate with the least padding
    int wpt = 8;
    {
      int64_t best_waste = -1;
      for (int cand : {8, 6, 4}) {
        const int64_t span = (int64_t)TAB_THREADS * cand;
        const int64_t waste = ceil_div(n, span) * span - n;
        if (best_waste < 0 || waste < best_waste) { best_waste = waste; wpt = cand; }
      }
    }
#define SLK_LAUNCH_TAB(HAS, WPT)                                                                                     \
    scale_search_tab_kernel<HAS, WPT><<<tgrid, TAB_THREADS, sizeof(TabSmem), st>>>(                                    \
        w, r, n, g, brk, cmin, cmax, factors, G, (const float*)hdiag, out_scale, out_err, out_init)
    if (h_dtype == 1) {
      if (wpt == 8) SLK_LAUNCH_TAB(true, 8); else if (wpt == 6) SLK_LAUNCH_TAB(true, 6); else SLK_LAUNCH_TAB(true, 4);
    } else {
      if (wpt == 8) SLK_LAUNCH_TAB(false, 8); else if (wpt == 6) SLK_LAUNCH_TAB(false, 6); else SLK_LAUNCH_TAB(false, 4);
    }
#undef SLK_LAUNCH_TAB
    SLK_LAUNCH_CHECK();
    return SLK_OK;
  }
#define SLK_LAUNCH_SS(HT, HAS)                                                                           \
  do {                                                                                                   \
    if (smem > 48 * 1024)                                                                                \
      SLK_CUDA(cudaFuncSetAttribute(scale_search_kernel<HT, HAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    scale_search_kernel<HT, HAS><<<grid, 128, smem, st>>>(w, r, n, g, cmin, cmax, factors, G, (const HT*)hdiag, \
                                                           row_in_smem, out_scale, out_err, out_init);   \
  } while (0)
  if (h_dtype == 0) SLK_LAUNCH_SS(float, false);
  else if (h_dtype == 1) SLK_LAUNCH_SS(float, true);
  else SLK_LAUNCH_SS(double, true);
#undef SLK_LAUNCH_SS
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
