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
#include "gemm.cuh"

namespace slk {

constexpr int NB = 64;

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

// One CTA: factor the 64x64 diagonal block at (k0, k0) in place (lower), write its inverse into Li.
__global__ void __launch_bounds__(256) hinv_diag_kernel(double* __restrict__ A, double* __restrict__ Li, int64_t ld,
                                                        int64_t k0, int32_t* __restrict__ info) {
  extern __shared__ __align__(16) double diag_smem[];
  double (*L)[NB + 1] = (double (*)[NB + 1])diag_smem;
  double (*X)[NB + 1] = (double (*)[NB + 1])(diag_smem + NB * (NB + 1));
  const int tid = threadIdx.x;
  for (int t = tid; t < NB * NB; t += 256) {
    int i = t / NB, j = t % NB;
    L[i][j] = A[(k0 + i) * ld + k0 + j];
  }
  __syncthreads();
  // right-looking Cholesky, lower triangle
  for (int j = 0; j < NB; ++j) {
    const double piv = L[j][j];
    if (tid == 0 && !(piv > 0.0)) atomicCAS(info, 0, (int32_t)(k0 + j + 1));
    const double d = __dsqrt_rn(piv);
    __syncthreads();
    if (tid < NB) {
      if (tid == j) L[j][j] = d;
      else if (tid > j) L[tid][j] = __ddiv_rn(L[tid][j], d);
    }
    __syncthreads();
    // trailing update of the lower triangle: L[i][c] -= L[i][j] * L[c][j], j < c <= i
    const int rem = NB - 1 - j;
    for (int t = tid; t < rem * rem; t += 256) {
      int i = j + 1 + t / rem, c = j + 1 + t % rem;
      if (c <= i) L[i][c] = __fma_rn(-L[i][j], L[c][j], L[i][c]);
    }
    __syncthreads();
  }
  // inverse of the lower-triangular block: column c handled by 4 lanes splitting the k-sum
  {
    const int c = tid >> 2, part = tid & 3;
    for (int i = 0; i < NB; ++i) {
      double s = 0.0;
      if (i > c) {
        for (int k = c + part; k < i; k += 4) s = __fma_rn(L[i][k], X[k][c], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (part == 0) {
        double v;
        if (i < c) v = 0.0;
        else if (i == c) v = __ddiv_rn(1.0, L[i][i]);
        else v = __ddiv_rn(-s, L[i][i]);
        X[i][c] = v;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int t = tid; t < NB * NB; t += 256) {
    int i = t / NB, j = t % NB;
    A[(k0 + i) * ld + k0 + j] = (j <= i) ? L[i][j] : 0.0;
    Li[(k0 + i) * ld + k0 + j] = X[i][j];
  }
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
  const size_t diag_smem_bytes = (size_t)2 * NB * (NB + 1) * sizeof(double);
  SLK_CUDA(cudaFuncSetAttribute(hinv_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem_bytes));
  for (int64_t k = 0; k < nblk; ++k) {
    const int64_t k0 = k * NB;
    hinv_diag_kernel<<<1, 256, diag_smem_bytes, st>>>(A, Li, ld, k0, info);
    SLK_LAUNCH_CHECK();
    const int64_t rest = npad - (k0 + NB);
    if (rest <= 0) break;
    // panel solve: A[k0+NB:, k0:k0+NB] <- A[...] @ inv(L_kk)^T   (in place: one CTA owns full rows)
    {
      double* P = A + (k0 + NB) * ld + k0;
      GemmParams<double> p = gemm_params<double>(P, ld, Li + k0 * ld + k0, ld, P, ld, rest, NB, NB);
      int rc = gemm_launch<double, false, true, EPI_STORE>(p, 1, st);
      if (rc) return rc;
    }
    // trailing update: A[k0+NB:, k0+NB:] -= P @ P^T, lower tiles only
    {
      const double* P = A + (k0 + NB) * ld + k0;
      GemmParams<double> p = gemm_params<double>(P, ld, P, ld, A + (k0 + NB) * ld + (k0 + NB), ld, rest, rest, NB);
      p.alpha = -1.0;
      p.lower_only = 1;
      int rc = gemm_launch<double, false, true, EPI_ACCUM>(p, 1, st);
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
      int rc = gemm_launch<double, false, false, EPI_STORE>(p, (int)pairs, st);
      if (rc) return rc;
    }
    {
      GemmParams<double> p = gemm_params<double>(Li + s * (ld + 1), ld, T + s * ld, ld, Li + s * ld, ld, s, s, s);
      p.strideA = p.strideB = p.strideC = stride;
      p.M_last = m_last; p.K_last = m_last;
      p.alpha = -1.0;
      p.k_hi_from_m = 1;
      int rc = gemm_launch<double, false, false, EPI_STORE>(p, (int)pairs, st);
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
