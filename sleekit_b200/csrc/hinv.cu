// K2: dampening + column permutation + fp64 factor of the inverse Hessian.
//   H_opt = H + damp*mean(diag H)*I ; H_opt[order][:, order]      obq.py:198-204
//   compute_hessian_chol: U = flip(inv(cholesky(flip(H_opt))))     obq.py:38-55
// Everything here is fp64, as in the reference (LAPACK dpotrf + dgesv there).
//
// Layout in the workspace (all [npad, npad] fp64, npad = n rounded up to 64):
//   A    flip(H_opt permuted), identity on the padding; overwritten by its lower factor L
//   Li   L^-1, built from 64x64 diagonal-block inverses by recursive doubling
//   T    scratch for the doubling products
// Steps: gather -> for every 64-wide panel { diag block factor+inverse (1 CTA, smem) ; panel
// solve as a GEMM with the block inverse ; trailing symmetric update (lower tiles only) }
// -> log2(npad/64) levels of batched triangular products -> flip + (fp64|fp32) store.
#include "dgemm_mma.cuh"

namespace slk {

constexpr int NB = 64;

// fp64 products of K2: FP64 tensor path (DMMA) when the operands qualify (always inside K2: all
// extents are multiples of 64), DFMA tiles otherwise.
template <bool B_T, int EPI>
static int dgemm(const GemmParams<double>& p, int batch, cudaStream_t st) {
  if (dgemm_mma_ok(p)) return dgemm_mma_launch<B_T, EPI>(p, batch, st);
  return gemm_launch<double, false, B_T, EPI>(p, batch, st);
}

// A[i, j] = flip(H_opt[order][:, order])[i, j], identity outside n
template <typename TS>
__global__ void __launch_bounds__(256) hinv_gather_kernel(const TS* __restrict__ h, int64_t n, int64_t npad,
                                                          const int64_t* __restrict__ order,
                                                          const float* __restrict__ dampval, double* __restrict__ A) {
  const int64_t total = npad * npad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double damp = dampval ? (double)dampval[0] : 0.0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int64_t i = t / npad, j = t - i * npad;
    double v;
    if (i < n && j < n) {
      int64_t si = n - 1 - i, sj = n - 1 - j;
      if (order) { si = __ldg(order + si); sj = __ldg(order + sj); }
      v = (double)h[si * n + sj];
      if (i == j) v = __dadd_rn(v, damp);  // obq.py:198: fp32 H promoted, fp64 add
    } else {
      v = (i == j) ? 1.0 : 0.0;
    }
    A[t] = v;
  }
}

// ---- 64x64 diagonal block: factor + inverse, warp-level ----------------------------------------
// The block is split 2x2 into 32x32 tiles.  A 32x32 Cholesky and a 32x32 triangular inverse each
// run in ONE warp with a row per lane and only __syncwarp between the 32 dependent steps (no
// block barriers); the four 32^3 products between them use all 128 threads on shared memory.
constexpr int HB = 32;
constexpr int LDP = NB + 1;  // padded leading dimension of the smem blocks

// Cholesky of the 32x32 tile at (r0, r0) of M (lower, in place) and its inverse into the same
// tile of X.  Executed by one full warp; tiles stay in shared memory and every loop is a real
// loop (a fully unrolled register version runs ~10^4 instructions exactly once and is
// instruction-fetch bound; measured 150 us per panel).
//  - factor: left-looking, lane i owns row i.  Step j: s_i = A[i][j] - sum_{k<j} L[i][k] L[j][k]
//    (reads only, no shared-memory read-modify-write), lane j's s is the pivot, broadcast by
//    shuffle; one __syncwarp per column.
//  - inverse: lane j solves L x = e_j by forward substitution on its own column of X; lanes do
//    not communicate at all.
__device__ __noinline__ void chol32_and_inverse(double (*M)[LDP], double (*X)[LDP], int r0, int lane,
                                                int32_t* info, int64_t gcol0) {
  __shared__ double rdiag[HB];                 // 1 / L[j][j]
  double* mrow = &M[r0 + lane][r0];
  for (int j = 0; j < HB; ++j) {
    const double* jrow = &M[r0 + j][r0];
    double s0 = mrow[j], s1 = 0.0;
    int k = 0;
    for (; k + 1 < j; k += 2) {
      s0 = __fma_rn(-mrow[k], jrow[k], s0);
      s1 = __fma_rn(-mrow[k + 1], jrow[k + 1], s1);
    }
    if (k < j) s0 = __fma_rn(-mrow[k], jrow[k], s0);
    const double s = s0 + s1;
    const double piv = __shfl_sync(0xffffffffu, s, j);
    if (lane == 0 && !(piv > 0.0)) atomicCAS(info, 0, (int32_t)(gcol0 + j + 1));
    const double y = rsqrt(piv);
    __syncwarp();
    if (lane == j) { mrow[j] = piv * y; rdiag[j] = y; }   // sqrt(piv) and its reciprocal
    else if (lane > j) mrow[j] = s * y;
    else mrow[j] = 0.0;                        // upper part of L
    __syncwarp();
  }
  // X = inv(L): lane c owns column c.  x_i = (delta_ic - sum_{k<i} L[i][k] x_k) / L[i][i], with
  // x_k = 0 for k < c, so the sum may start at k = c.
  double* xcol = &X[r0][r0 + lane];            // element i of this lane's column: xcol[i * LDP]
  for (int i = 0; i < HB; ++i) {
    const double* irow = &M[r0 + i][r0];
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
    if (i > lane) {
      int k = lane;
      for (; k + 1 < i; k += 2) {
        s0 = __fma_rn(-irow[k], xcol[k * LDP], s0);
        s1 = __fma_rn(-irow[k + 1], xcol[(k + 1) * LDP], s1);
      }
      if (k < i) s0 = __fma_rn(-irow[k], xcol[k * LDP], s0);
    }
    xcol[i * LDP] = i >= lane ? (s0 + s1) * rdiag[i] : 0.0;
  }
  __syncwarp();
}

// C(32x32 at cr,cc) = beta*C + alpha * A(32x32 at ar,ac) * op(B)(32x32), op = transpose if TB.
// All 128 threads; each owns a 2x4 micro tile.  Barriers on entry (operands ready) and between
// the reads and the in-place write.
template <bool TB>
__device__ __forceinline__ void mm32(double (*C)[LDP], int cr, int cc, double (*A)[LDP], int ar, int ac,
                                     double (*B)[LDP], int br, int bc, double alpha, double beta, int tid) {
  const int ri = (tid >> 3) * 2, cj = (tid & 7) * 4;
  double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  __syncthreads();
#pragma unroll 8
  for (int k = 0; k < HB; ++k) {
    const double a0 = A[ar + ri][ac + k], a1 = A[ar + ri + 1][ac + k];
    double bv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) bv[c] = TB ? B[br + cj + c][bc + k] : B[br + k][bc + cj + c];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      acc[0][c] = __fma_rn(a0, bv[c], acc[0][c]);
      acc[1][c] = __fma_rn(a1, bv[c], acc[1][c]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double* dst = &C[cr + ri + rr][cc + cj + c];
      *dst = beta == 0.0 ? alpha * acc[rr][c] : beta * (*dst) + alpha * acc[rr][c];
    }
}

// Factor the 64x64 block held in L (lower, in place) and leave its inverse in X.  128 threads.
__device__ __forceinline__ void factor64(double (*L)[LDP], double (*X)[LDP], int tid, int32_t* info, int64_t gcol0) {
  const int warp = tid >> 5, lane = tid & 31;
  if (warp == 0) chol32_and_inverse(L, X, 0, lane, info, gcol0);               // L11, Y11
  mm32<true>(L, HB, 0, L, HB, 0, X, 0, 0, 1.0, 0.0, tid);                       // L21 = A21 Y11^T
  mm32<true>(L, HB, HB, L, HB, 0, L, HB, 0, -1.0, 1.0, tid);                    // A22 -= L21 L21^T
  __syncthreads();
  if (warp == 0) chol32_and_inverse(L, X, HB, lane, info, gcol0 + HB);         // L22, Y22
  mm32<false>(X, HB, 0, X, HB, HB, L, HB, 0, 1.0, 0.0, tid);                    // T = Y22 L21   (into X21)
  mm32<false>(X, HB, 0, X, HB, 0, X, 0, 0, -1.0, 0.0, tid);                     // X21 = -T Y11
  __syncthreads();
  for (int t = tid; t < HB * HB; t += 128) {                                     // zero the upper-right tiles
    const int i = t / HB, j = t % HB;
    L[i][HB + j] = 0.0;
    X[i][HB + j] = 0.0;
  }
  __syncthreads();
}

// Panel step k0: every CTA factors the diagonal block (redundantly -- it is cheaper than a
// dependent launch); CTA 0 stores L_kk and inv(L_kk); CTA b > 0 solves its 64-row slab of the
// panel, A[k0+64b : +64, k0 : k0+64] <- slab @ inv(L_kk)^T, in place.
__global__ void __launch_bounds__(128) hinv_panel_kernel(double* __restrict__ A, double* __restrict__ Li, int64_t ld,
                                                         int64_t k0, int32_t* __restrict__ info) {
  extern __shared__ __align__(16) double diag_smem[];
  double (*L)[LDP] = (double (*)[LDP])diag_smem;
  double (*X)[LDP] = (double (*)[LDP])(diag_smem + NB * LDP);
  const int tid = threadIdx.x;
  // All global loads are issued up front, 32 per thread in flight (one memory latency instead of
  // 32 in a row): the diagonal block, and for slab CTAs the slab too, which then sits in registers
  // while the block is factored.
  const double* D = A + k0 * ld + k0;
  double* S = A + (k0 + (int64_t)blockIdx.x * NB) * ld + k0;   // blockIdx.x == 0: the diagonal block itself
  double dv[32], sv[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int t = tid + k * 128;
    dv[k] = D[(t >> 6) * ld + (t & 63)];
  }
  if (blockIdx.x != 0) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int t = tid + k * 128;
      sv[k] = S[(t >> 6) * ld + (t & 63)];
    }
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int t = tid + k * 128;
    L[t >> 6][t & 63] = dv[k];
  }
  __syncthreads();
  factor64(L, X, tid, info, k0);
  if (blockIdx.x == 0) {
    // L_kk itself is never needed again (the other CTAs of this launch are still reading the
    // unfactored block, so it must not be overwritten here); only its inverse is kept.
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int t = tid + k * 128;
      Li[(k0 + (t >> 6)) * ld + k0 + (t & 63)] = X[t >> 6][t & 63];
    }
    return;
  }
  // slab solve: S <- S @ X^T, X lower triangular
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int t = tid + k * 128;
    L[t >> 6][t & 63] = sv[k];
  }
  __syncthreads();
  const int ri = (tid >> 3) * 4, cj = (tid & 7) * 8;
  double acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int k = 0; k < NB; ++k) {
    double av[4], bv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = L[ri + i][k];
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = X[cj + j][k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = __fma_rn(av[i], bv[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) S[(ri + i) * ld + cj + j] = acc[i][j];
}

// U[i, j] = Li[n-1-i, n-1-j] on and above the diagonal, 0 below
__global__ void __launch_bounds__(256) hinv_flip_kernel(const double* __restrict__ Li, int64_t n, int64_t ld,
                                                        double* __restrict__ u64, float* __restrict__ u32) {
  const int64_t total = n * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    int64_t i = t / n, j = t - i * n;
    double v = (j >= i) ? Li[(n - 1 - i) * ld + (n - 1 - j)] : 0.0;
    if (u64) u64[t] = v;
    if (u32) u32[t] = (float)v;
  }
}

static inline int64_t pad64(int64_t n) { return (n + NB - 1) / NB * NB; }

static int hinv_factor_and_invert(double* A, double* Li, double* T, int64_t n, int64_t npad, double* u64, float* u32,
                                  int32_t* info, cudaStream_t st) {
  const int64_t ld = npad;
  SLK_CUDA(cudaMemsetAsync(Li, 0, (size_t)npad * npad * sizeof(double), st));
  SLK_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  const int64_t nblk = npad / NB;
  const size_t diag_smem_bytes = (size_t)2 * NB * LDP * sizeof(double);
  SLK_SMEM_ATTR_ONCE(hinv_panel_kernel, (int)diag_smem_bytes);
  for (int64_t k = 0; k < nblk; ++k) {
    const int64_t k0 = k * NB;
    const int64_t rest = npad - (k0 + NB);
    // diagonal block factor + inverse, fused with the panel solve (one CTA per 64-row slab)
    hinv_panel_kernel<<<(unsigned)(1 + rest / NB), 128, diag_smem_bytes, st>>>(A, Li, ld, k0, info);
    SLK_LAUNCH_CHECK();
    if (rest <= 0) break;
    // trailing update: A[k0+NB:, k0+NB:] -= P @ P^T, lower tiles only
    {
      const double* P = A + (k0 + NB) * ld + k0;
      GemmParams<double> p = gemm_params<double>(P, ld, P, ld, A + (k0 + NB) * ld + (k0 + NB), ld, rest, rest, NB);
      p.alpha = -1.0;
      p.lower_only = 1;
      int rc = dgemm<true, EPI_ACCUM>(p, 1, st);
      if (rc) return rc;
    }
  }
  // recursive doubling: inv([[A,0],[C,B]]) = [[Ai,0],[-Bi C Ai, Bi]]
  for (int64_t s = NB; s < npad; s *= 2) {
    const int64_t pairs = ceil_div(npad - s, 2 * s);  // pairs that have a lower block
    if (pairs <= 0) break;
    const int64_t stride = 2 * s * (ld + 1);
    const int64_t b0_last = (pairs - 1) * 2 * s + s;
    const int64_t m_last = (npad - b0_last) < s ? (npad - b0_last) : s;
    {
      GemmParams<double> p = gemm_params<double>(A + s * ld, ld, Li, ld, T + s * ld, ld, s, s, s);
      p.strideA = p.strideB = p.strideC = stride;
      p.M_last = m_last; p.K_last = s;
      p.k_lo_from_n = 1;
      int rc = dgemm<false, EPI_STORE>(p, (int)pairs, st);
      if (rc) return rc;
    }
    {
      GemmParams<double> p = gemm_params<double>(Li + s * (ld + 1), ld, T + s * ld, ld, Li + s * ld, ld, s, s, s);
      p.strideA = p.strideB = p.strideC = stride;
      p.M_last = m_last; p.K_last = m_last;
      p.alpha = -1.0;
      p.k_hi_from_m = 1;
      int rc = dgemm<false, EPI_STORE>(p, (int)pairs, st);
      if (rc) return rc;
    }
  }
  if (u64 || u32) {
    int64_t blocks = ceil_div(n * n, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    hinv_flip_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(Li, n, ld, u64, u32);
    SLK_LAUNCH_CHECK();
  }
  return SLK_OK;
}

template <typename TS>
static int hinv_impl(const TS* h, int64_t n, const int64_t* order, const float* dampval, void* ws, size_t ws_bytes,
                     double* u64, float* u32, int32_t* info, cudaStream_t st) {
  SLK_REQUIRE(h && info && n >= 1, "bad arguments");
  SLK_REQUIRE(ws && ws_bytes >= slk_hinv_ws_bytes(n), "workspace too small");
  const int64_t npad = pad64(n);
  double* A = (double*)ws;
  double* Li = A + npad * npad;
  double* T = Li + npad * npad;
  int64_t blocks = ceil_div(npad * npad, 256);
  int64_t cap = (int64_t)sm_count() * 16;
  hinv_gather_kernel<TS><<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(h, n, npad, order, dampval, A);
  SLK_LAUNCH_CHECK();
  return hinv_factor_and_invert(A, Li, T, n, npad, u64, u32, info, st);
}

}  // namespace slk

using namespace slk;

extern "C" {

size_t slk_hinv_ws_bytes(int64_t n) {
  const int64_t npad = pad64(n);
  return (size_t)3 * npad * npad * sizeof(double) + 256;
}

int slk_hinv_from_f32(const float* h, int64_t n, const int64_t* order, const float* dampval, void* ws,
                      size_t ws_bytes, double* u64, float* u32, int32_t* info, void* stream) {
  return hinv_impl<float>(h, n, order, dampval, ws, ws_bytes, u64, u32, info, (cudaStream_t)stream);
}

int slk_hinv_from_f64(const double* h, int64_t n, void* ws, size_t ws_bytes, double* u64, float* u32,
                      int32_t* info, void* stream) {
  return hinv_impl<double>(h, n, nullptr, nullptr, ws, ws_bytes, u64, u32, info, (cudaStream_t)stream);
}

}  // extern "C"
