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

// Thread-per-row leaf: each thread keeps the live columns of its row in registers as a window
// that rotates by one per column, so every register index is static while the column loop is a
// real loop (small code: a fully unrolled 32x32 body is instruction-fetch bound).  All threads
// of a warp read the same U[i][j] (shared-memory broadcast); there is no cross-lane traffic and
// the only serial chain is the algorithm's own (column i+1 needs column i's residual).
// Arithmetic per column, as obq.py:110-118:
//   q = quant(w);  res = fp64(w - q) / U[i,i];  E[:, i] = fp32(res);
//   Q[:, j] = fp32(fp64(Q[:, j]) - res * U[i, j])   for j > i
// The two divides (by the codebook step and by U[i,i]) use the exact reciprocal scheme of
// common.cuh, so the results are the IEEE quotients.
struct LeafShared {
  double U[32][64];   // block of the factor, zero-padded to 64 columns (window reads run past 32)
  double Uy[32];      // RN(1 / U[i][i])
  int Uok[32];
};

template <int WIN>
__device__ __forceinline__ void leaf_phase(float (&q)[32], const LeafShared& sh, int i0, int width,
                                           const DevGrid<float>& g, const FastDivF& fstep, bool fastq,
                                           float* __restrict__ qrow, float* __restrict__ erow) {
#pragma unroll 1
  for (int t = 0; t < 8; ++t) {
    const int i = i0 + t;
    if (i >= width) return;
    const float w = q[0];
    const float qq = fastq ? uniform_value_fast(g, fstep, w) : grid_value(g, w);
    FastDivD fd;
    fd.d = sh.U[i][i]; fd.y = sh.Uy[i]; fd.ok = sh.Uok[i];
    const double res = fastdiv((double)__fsub_rn(w, qq), fd);               // obq.py:114
    qrow[i] = qq;                                                           // obq.py:116
    erow[i] = (float)res;                                                   // obq.py:115
    const double* urow = &sh.U[i][i];
#pragma unroll
    for (int j = 1; j < WIN; ++j)
      q[j - 1] = (float)__dsub_rn((double)q[j], __dmul_rn(res, urow[j]));   // obq.py:118
  }
}

__global__ void __launch_bounds__(LEAF_ROWS) sweep_leaf_kernel(float* __restrict__ Q, float* __restrict__ E, int64_t r,
                                                               int64_t n, int a, int width,
                                                               const double* __restrict__ U, DevGrid<float> g) {
  __shared__ LeafShared sh;
  {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = threadIdx.x + k * LEAF_ROWS, i = t >> 5, j = t & 31;
      v[k] = (i < width && j < width) ? __ldg(U + (int64_t)(a + i) * n + (a + j)) : ((i == j) ? 1.0 : 0.0);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = threadIdx.x + k * LEAF_ROWS, i = t >> 5, j = t & 31;
      sh.U[i][j] = v[k];
      sh.U[i][32 + j] = 0.0;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const FastDivD f = make_fastdiv(sh.U[threadIdx.x][threadIdx.x]);
    sh.Uy[threadIdx.x] = f.y;
    sh.Uok[threadIdx.x] = f.ok;
  }
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);
  const bool fastq = (g.kind == 0) && fstep.ok;
  float* qrow = Q + row * n + a;
  float* erow = E + row * n + a;
  float q[32];
  if ((width == 32) && ((((uintptr_t)qrow) & 15) == 0)) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = reinterpret_cast<const float4*>(qrow)[k];
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) q[k] = k < width ? qrow[k] : 0.0f;
  }
  leaf_phase<32>(q, sh, 0, width, g, fstep, fastq, qrow, erow);
  leaf_phase<24>(q, sh, 8, width, g, fstep, fastq, qrow, erow);
  leaf_phase<16>(q, sh, 16, width, g, fstep, fastq, qrow, erow);
  leaf_phase<8>(q, sh, 24, width, g, fstep, fastq, qrow, erow);
}

// ---- fp32 leaf (default) -------------------------------------------------------------------------
// Same structure, all-fp32 arithmetic on the fp32 rounding of U: res = (w - q) / U[i,i] (correctly
// rounded fp32 quotient), Q[:, j] = fma(-res, U[i, j], Q[:, j]).  On B200 the fp64<->fp32 converts
// of the exact leaf run on the XU pipe at a fraction of the FMA rate (ncu: XU at 111 % of its
// sustained peak, 33 us per 32 columns); this version has none.  SURVEY 7.3 H1 measured all-fp32
// sweeps at >= 99.998 % code agreement with the reference; tests assert the 1e-3 error bar.
struct LeafShared32 {
  float U[32][64];
  float Uy[32];
  int Uok[32];
};

template <int WIN>
__device__ __forceinline__ void leaf_phase32(float (&q)[32], const LeafShared32& sh, int i0, int width,
                                             const DevGrid<float>& g, const FastDivF& fstep, bool fastq,
                                             float* __restrict__ qrow, float* __restrict__ erow) {
#pragma unroll 1
  for (int t = 0; t < 8; ++t) {
    const int i = i0 + t;
    if (i >= width) return;
    const float w = q[0];
    const float qq = fastq ? uniform_value_fast(g, fstep, w) : grid_value(g, w);
    FastDivF fd;
    fd.d = sh.U[i][i]; fd.y = sh.Uy[i]; fd.ok = sh.Uok[i];
    const float res = fastdiv(__fsub_rn(w, qq), fd);
    qrow[i] = qq;
    erow[i] = res;
    const float* urow = &sh.U[i][i];
#pragma unroll
    for (int j = 1; j < WIN; ++j) q[j - 1] = __fmaf_rn(-res, urow[j], q[j]);
  }
}

__global__ void __launch_bounds__(LEAF_ROWS) sweep_leaf32_kernel(float* __restrict__ Q, float* __restrict__ E, int64_t r,
                                                                 int64_t n, int a, int width,
                                                                 const float* __restrict__ U, DevGrid<float> g) {
  __shared__ LeafShared32 sh;
  {
    // 32x32 block of U: every thread issues its 16 loads before using any (one memory latency,
    // not 32 in a row); columns 32..63 are the zero padding the rotating window reads past.
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = threadIdx.x + k * LEAF_ROWS, i = t >> 5, j = t & 31;
      v[k] = (i < width && j < width) ? __ldg(U + (int64_t)(a + i) * n + (a + j)) : ((i == j) ? 1.0f : 0.0f);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int t = threadIdx.x + k * LEAF_ROWS, i = t >> 5, j = t & 31;
      sh.U[i][j] = v[k];
      sh.U[i][32 + j] = 0.0f;
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const FastDivF f = make_fastdiv(sh.U[threadIdx.x][threadIdx.x]);
    sh.Uy[threadIdx.x] = f.y;
    sh.Uok[threadIdx.x] = f.ok;
  }
  __syncthreads();
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r) return;
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);
  const bool fastq = (g.kind == 0) && fstep.ok;
  float* qrow = Q + row * n + a;
  float* erow = E + row * n + a;
  float q[32];
  if ((width == 32) && ((((uintptr_t)qrow) & 15) == 0)) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = reinterpret_cast<const float4*>(qrow)[k];
      q[4 * k] = v.x; q[4 * k + 1] = v.y; q[4 * k + 2] = v.z; q[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) q[k] = k < width ? qrow[k] : 0.0f;
  }
  leaf_phase32<32>(q, sh, 0, width, g, fstep, fastq, qrow, erow);
  leaf_phase32<24>(q, sh, 8, width, g, fstep, fastq, qrow, erow);
  leaf_phase32<16>(q, sh, 16, width, g, fstep, fastq, qrow, erow);
  leaf_phase32<8>(q, sh, 24, width, g, fstep, fastq, qrow, erow);
}

struct SweepCtx {
  float* Q; float* E; int64_t r, n;
  const double* u64; const float* u32;
  DevGrid<float> g;
  int leaf, fanout;
  int exact_leaf;
  cudaStream_t st;
};

// obq.py:121-137, same recursion; launches are issued in the reference's order
static int sweep_range(const SweepCtx& c, int64_t a, int64_t b) {
  const int64_t size = b - a;
  if (size <= c.leaf) {
    if (c.exact_leaf)
      sweep_leaf_kernel<<<(int)ceil_div(c.r, LEAF_ROWS), LEAF_ROWS, 0, c.st>>>(c.Q, c.E, c.r, c.n, (int)a, (int)size,
                                                                                c.u64, c.g);
    else
      sweep_leaf32_kernel<<<(int)ceil_div(c.r, LEAF_ROWS), LEAF_ROWS, 0, c.st>>>(c.Q, c.E, c.r, c.n, (int)a,
                                                                                  (int)size, c.u32, c.g);
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
                                  const slk_codebook* cb, int32_t leaf, int32_t fanout, int32_t exact_leaf,
                                  void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(r >= 0 && n >= 1, "bad shape");
  SLK_REQUIRE(leaf >= 1 && leaf <= 32, "leaf width %d not in [1, 32]", leaf);
  SLK_REQUIRE(fanout >= 2, "fanout %d < 2", fanout);
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(q && e && u32 && (u64 || !exact_leaf), "NULL pointer");
  SweepCtx c;
  c.Q = q; c.E = e; c.r = r; c.n = n; c.u64 = u64; c.u32 = u32;
  c.g = make_grid<float>(cb);
  c.leaf = leaf; c.fanout = fanout; c.exact_leaf = exact_leaf; c.st = (cudaStream_t)stream;
  return sweep_range(c, 0, n);
}
