// K3: the GPTQ / OBQ column sweep.
//   _quantize_opt_core   (leaf, <= 32 columns)      obq.py:106-118
//   _quantize_opt_block  (8-ary lazy batching)      obq.py:121-137
// Leaf: one thread per row with its 32 columns in registers.  Column i is quantised, its scaled
// residual r = (w - q) / U[i,i] is formed in fp64 (as numpy does: the divisor is an fp64 scalar)
// and columns j > i take q_j <- fp32(fp64(q_j) - r * U[i,j]) -- the exact op sequence of
// obq.py:114-118, so given the same fp64 U a leaf is bit-identical to the reference.  Trailing updates Q[:, b:end] -= E[:, a:b] @ U[a:b, b:end] are fp32 GEMMs
// with exact-product fmaf accumulation on the fp32 rounding of U (parity-safe).
#include "gemm.cuh"

namespace slk {

constexpr int LEAF_ROWS = 64;  // rows (threads) per CTA of the leaf kernel

// Thread-per-row leaf: each thread keeps its 32 columns in registers and walks them in order.
// All threads of a warp read the same U[i][j] (shared-memory broadcast), so there is no
// cross-lane traffic at all and the only serial chain is the one the algorithm imposes
// (column i+1 needs column i's residual).  Arithmetic per column, as obq.py:110-118:
//   q = quant(w);  res = fp64(w - q) / U[i,i];  E[:, i] = fp32(res);
//   Q[:, j] = fp32(fp64(Q[:, j]) - res * U[i, j])   for j > i
// The two divides (by the codebook step and by U[i,i]) use the exact reciprocal scheme of
// common.cuh, so the results are the IEEE quotients.
__global__ void __launch_bounds__(LEAF_ROWS) sweep_leaf_kernel(float* __restrict__ Q, float* __restrict__ E, int64_t r,
                                                               int64_t n, int a, int width,
                                                               const double* __restrict__ U, DevGrid<float> g) {
  __shared__ double Us[32][32];
  __shared__ double Uy[32];   // RN(1 / U[i][i])
  __shared__ int Uok[32];
  for (int t = threadIdx.x; t < 32 * 32; t += blockDim.x) {
    const int i = t >> 5, j = t & 31;
    double v = (i < width && j < width) ? U[(int64_t)(a + i) * n + (a + j)] : (i == j ? 1.0 : 0.0);
    Us[i][j] = v;
    if (i == j) {
      const FastDivD f = make_fastdiv(v);
      Uy[i] = f.y;
      Uok[i] = f.ok;
    }
  }
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);
  const bool fastq = (g.kind == 0) && fstep.ok;
  float* qrow = Q + row * n + a;
  float* erow = E + row * n + a;
  float q[32], e[32];
  const bool vec = (width == 32) && ((((uintptr_t)qrow) & 15) == 0) && ((((uintptr_t)erow) & 15) == 0);
  if (vec) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = reinterpret_cast<const float4*>(qrow)[k];
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) q[k] = k < width ? qrow[k] : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float w = q[i];
    const float qq = fastq ? uniform_value_fast(g, fstep, w) : grid_value(g, w);
    const double num = (double)__fsub_rn(w, qq);
    FastDivD fd;
    fd.d = Us[i][i]; fd.y = Uy[i]; fd.ok = Uok[i];
    const double res = fastdiv(num, fd);                                   // obq.py:114
    q[i] = qq;                                                             // obq.py:116
    e[i] = (float)res;                                                     // obq.py:115
#pragma unroll
    for (int j = i + 1; j < 32; ++j)
      q[j] = (float)__dsub_rn((double)q[j], __dmul_rn(res, Us[i][j]));     // obq.py:118
  }
  if (vec) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      reinterpret_cast<float4*>(qrow)[k] = make_float4(q[4 * k], q[4 * k + 1], q[4 * k + 2], q[4 * k + 3]);
      reinterpret_cast<float4*>(erow)[k] = make_float4(e[4 * k], e[4 * k + 1], e[4 * k + 2], e[4 * k + 3]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k)
      if (k < width) { qrow[k] = q[k]; erow[k] = e[k]; }
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
    sweep_leaf_kernel<<<(int)ceil_div(c.r, LEAF_ROWS), LEAF_ROWS, 0, c.st>>>(c.Q, c.E, c.r, c.n, (int)a, (int)size,
                                                                              c.u64, c.g);
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
