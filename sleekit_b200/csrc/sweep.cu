// K3: the GPTQ / OBQ column sweep.
//   _quantize_opt_core   (leaf, <= 32 columns)      obq.py:106-118
//   _quantize_opt_block  (8-ary lazy batching)      obq.py:121-137
// Leaf: one warp per row, lane j holds column a+j.  Column i is broadcast with a shuffle,
// quantised, its scaled residual r = (w - q) / U[i,i] is formed in fp64 (as numpy does: the
// divisor is an fp64 scalar) and lanes j > i take q_j <- fp32(fp64(q_j) - r * U[i,j]) -- the
// exact op sequence of obq.py:114-118, so given the same fp64 U a leaf is bit-identical to
// the reference.  Trailing updates Q[:, b:end] -= E[:, a:b] @ U[a:b, b:end] are fp32 GEMMs
// with exact-product fmaf accumulation on the fp32 rounding of U (parity-safe).
#include "gemm.cuh"

namespace slk {

__global__ void __launch_bounds__(128) sweep_leaf_kernel(float* __restrict__ Q, float* __restrict__ E, int64_t r,
                                                         int64_t n, int a, int width, const double* __restrict__ U,
                                                         DevGrid<float> g) {
  __shared__ double Us[32][33];
  for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
    int i = t >> 5, j = t & 31;
    Us[i][j] = (i < width && j < width) ? U[(int64_t)(a + i) * n + (a + j)] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < r;
       row += (int64_t)gridDim.x * warps_per_block) {
    float* qrow = Q + row * n + a;
    float q = lane < width ? qrow[lane] : 0.0f;
    float e = 0.0f;
    for (int i = 0; i < width; ++i) {
      const float w = __shfl_sync(0xffffffffu, q, i);
      const float qq = grid_value(g, w);
      const double res = __ddiv_rn((double)__fsub_rn(w, qq), Us[i][i]);   // obq.py:114
      if (lane == i) { q = qq; e = (float)res; }                          // obq.py:115-116
      else if (lane > i) q = (float)__dsub_rn((double)q, __dmul_rn(res, Us[i][lane]));  // obq.py:118
    }
    if (lane < width) {
      qrow[lane] = q;
      E[row * n + a + lane] = e;
    }
  }
}

struct SweepCtx {
  float* Q; float* E; int64_t r, n;
  const double* u64; const float* u32;
  DevGrid<float> g;
  int leaf, fanout;
  cudaStream_t st;
};

// obq.py:121-137, same recursion; launches are issued in the reference's order
static int sweep_range(const SweepCtx& c, int64_t a, int64_t b) {
  const int64_t size = b - a;
  if (size <= c.leaf) {
    const int warps = 4;
    int64_t blocks = ceil_div(c.r, warps);
    int64_t cap = (int64_t)sm_count() * 16;
    sweep_leaf_kernel<<<(int)(blocks < cap ? blocks : cap), warps * 32, 0, c.st>>>(c.Q, c.E, c.r, c.n, (int)a,
                                                                                     (int)size, c.u64, c.g);
    SLK_LAUNCH_CHECK();
    return SLK_OK;
  }
  int64_t width = (size + c.fanout - 1) / c.fanout;
  if (width < c.leaf) width = c.leaf;
  for (int64_t s = a; s < b; s += width) {
    const int64_t e = (s + width) < b ? (s + width) : b;
    int rc = sweep_range(c, s, e);
    if (rc) return rc;
    if (e < b) {
      // Q[:, e:b] -= E[:, s:e] @ U[s:e, e:b]
      GemmParams<float> p = gemm_params<float>(c.E + s, c.n, c.u32 + s * c.n + e, c.n, c.Q + e, c.n, c.r, b - e, e - s);
      p.alpha = -1.0f;
      rc = gemm_launch<float, false, false, EPI_ACCUM>(p, 1, c.st);
      if (rc) return rc;
    }
  }
  return SLK_OK;
}

}  // namespace slk

using namespace slk;

extern "C" int slk_gptq_sweep_f32(float* q, float* e, int64_t r, int64_t n, const double* u64, const float* u32,
                                  const slk_codebook* cb, int32_t leaf, int32_t fanout, void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(r >= 0 && n >= 1, "bad shape");
  SLK_REQUIRE(leaf >= 1 && leaf <= 32, "leaf width %d not in [1, 32]", leaf);
  SLK_REQUIRE(fanout >= 2, "fanout %d < 2", fanout);
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(q && e && u64 && u32, "NULL pointer");
  SweepCtx c;
  c.Q = q; c.E = e; c.r = r; c.n = n; c.u64 = u64; c.u32 = u32;
  c.g = make_grid<float>(cb);
  c.leaf = leaf; c.fanout = fanout; c.st = (cudaStream_t)stream;
  return sweep_range(c, 0, n);
}
