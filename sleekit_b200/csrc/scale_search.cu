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
