// CUDA-core tiled GEMM core shared by the dense phases of the hot path (fp32 and fp64):
//   K1  X^T X            (statistics.py:87)        A transposed, epilogue keep*C + acc/count
//   K2  fp64 panel / trailing / inverse products (obq.py:38-55)
//   K3  trailing update  Q[:, e:] -= E[:, i:e] @ U[i:e, e:]   (obq.py:137)
//   K6  ((W-Q) @ H) * (W-Q) row sums              (obq.py:95, scaling.py:95)
//   K7  (Q-W) @ H        (obq.py:231)
// fp32 products use explicit fmaf (exact-product accumulate), so results are fp32-faithful;
// the tcgen05 3xTF32 path in tc_gemm.cu is checked against this one.
//
// Tiling: 256 threads as 16x16; each thread owns a (2H x 2H) micro tile split in two halves
// 16H apart so that shared-memory reads are 128-bit and conflict-free.  H = 4 for fp32
// (128x128x8 CTA tile), H = 2 for fp64 (64x64x8).  Global loads are register-prefetched one
// k-tile ahead and double-buffered in shared memory (one barrier per k-tile).
#pragma once

#include "common.cuh"

namespace slk {

enum GemmEpilogue {
  EPI_STORE = 0,    // C = alpha * acc
  EPI_ACCUM = 1,    // C = C + alpha * acc
  EPI_HESS = 2,     // C = C * keep + acc / count          (statistics.py:82-87)
  EPI_ROWDOT = 3,   // part[m, tile_n] = sum_n acc[m, n] * a(m, n)   (needs N == K)
  EPI_GAIN = 4      // C = -D^2 * diag[n] - 2 * acc * D,  D = x0[m,n] - x1[m,n]  (obq.py:231)
};

template <typename T>
struct GemmParams {
  const T* A; int64_t lda;    // A_T ? [K, M] : [M, K]
  const T* A2;                // optional: operand is A - A2 (same layout), or NULL
  const T* B; int64_t ldb;    // B_T ? [N, K] : [K, N]
  T* C; int64_t ldc;          // [M, N]
  int64_t M, N, K;
  T alpha;
  T keep, count;              // EPI_HESS
  const T* x0; const T* x1; const T* x2; int64_t ldx; int64_t x2_stride;  // EPI_GAIN: cand, q, diag(H)
  // batching over blockIdx.z (uniform strides); the last batch may be smaller
  int64_t strideA, strideB, strideC;
  int64_t M_last, K_last;     // extents of the last batch (== M, K when not ragged)
  // triangular structure: skip k < n0 (B lower-tri, i.e. B[k,n]==0 for k<n) / k >= m0+BM
  int k_lo_from_n, k_hi_from_m;
  int lower_only;             // skip tiles entirely above the diagonal (n0 > m0 + BM - 1)
  int negate_a_diff;          // operand is A2 - A instead of A - A2
};

template <typename T> struct GemmTile { };
template <> struct GemmTile<float> { static constexpr int H = 4; static constexpr int PAD = 4; };
template <> struct GemmTile<double> { static constexpr int H = 2; static constexpr int PAD = 4; };

template <typename T, int H, bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(256) gemm_kernel(GemmParams<T> p) {
  constexpr int BM = 32 * H, BN = 32 * H, BK = 8;
  constexpr int LDS_ = BM + GemmTile<T>::PAD;
  constexpr int PER = BM * BK / 256;  // elements of each operand tile per thread
  typedef Ieee<T> F;

  __shared__ __align__(16) T As[2][BK][LDS_];
  __shared__ __align__(16) T Bs[2][BK][LDS_];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int z = blockIdx.z;
  const bool last = (z == (int)gridDim.z - 1);
  const int64_t M = last ? p.M_last : p.M;
  const int64_t K = last ? p.K_last : p.K;
  const int64_t N = p.N;
  if (m0 >= M || n0 >= N) return;
  if (p.lower_only && n0 > m0 + BM - 1) return;

  const T* __restrict__ A = p.A + (int64_t)z * p.strideA;
  const T* __restrict__ A2 = p.A2 ? p.A2 + (int64_t)z * p.strideA : nullptr;
  const T* __restrict__ B = p.B + (int64_t)z * p.strideB;
  T* __restrict__ C = p.C ? p.C + (int64_t)z * p.strideC : nullptr;

  int64_t kbeg = 0, kend = K;
  if (p.k_lo_from_n) kbeg = (n0 / BK) * BK;
  if (p.k_hi_from_m) { int64_t e = m0 + BM; kend = e < K ? e : K; }
  if (kbeg > kend) kbeg = kend;

  T acc[2 * H][2 * H];
#pragma unroll
  for (int i = 0; i < 2 * H; ++i)
#pragma unroll
    for (int j = 0; j < 2 * H; ++j) acc[i][j] = (T)0;

  T ra[PER], rb[PER];

  auto load_a = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int mm, kk;
      if (A_T) { mm = tid % BM; kk = tid / BM + (256 / BM) * i; }
      else { kk = tid % BK; mm = tid / BK + (256 / BK) * i; }
      int64_t m = m0 + mm, k = k0 + kk;
      T v = (T)0;
      if (m < M && k < kend) {
        int64_t off = A_T ? k * p.lda + m : m * p.lda + k;
        v = __ldg(A + off);
        if (A2) { T w = __ldg(A2 + off); v = p.negate_a_diff ? F::sub(w, v) : F::sub(v, w); }
      }
      ra[i] = v;
    }
  };
  auto load_b = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int nn, kk;
      if (B_T) { kk = tid % BK; nn = tid / BK + (256 / BK) * i; }
      else { nn = tid % BN; kk = tid / BN + (256 / BN) * i; }
      int64_t n = n0 + nn, k = k0 + kk;
      T v = (T)0;
      if (n < N && k < kend) v = __ldg(B + (B_T ? n * p.ldb + k : k * p.ldb + n));
      rb[i] = v;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      int mm, kk;
      if (A_T) { mm = tid % BM; kk = tid / BM + (256 / BM) * i; }
      else { kk = tid % BK; mm = tid / BK + (256 / BK) * i; }
      As[buf][kk][mm] = ra[i];
      int nn, kb;
      if (B_T) { kb = tid % BK; nn = tid / BK + (256 / BK) * i; }
      else { nn = tid % BN; kb = tid / BN + (256 / BN) * i; }
      Bs[buf][kb][nn] = rb[i];
    }
  };

  const int64_t nk = (kend - kbeg + BK - 1) / BK;
  if (nk > 0) {
    load_a(kbeg); load_b(kbeg);
    stash(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) { load_a(kbeg + (kt + 1) * BK); load_b(kbeg + (kt + 1) * BK); }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[2 * H], b[2 * H];
#pragma unroll
      for (int i = 0; i < H; ++i) {
        a[i] = As[buf][kk][ty * H + i];
        a[H + i] = As[buf][kk][16 * H + ty * H + i];
        b[i] = Bs[buf][kk][tx * H + i];
        b[H + i] = Bs[buf][kk][16 * H + tx * H + i];
      }
#pragma unroll
      for (int i = 0; i < 2 * H; ++i)
#pragma unroll
        for (int j = 0; j < 2 * H; ++j) acc[i][j] = F::fma(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) { stash(buf ^ 1); }
    __syncthreads();
  }

  // ---- epilogue ------------------------------------------------------------
  if (EPI == EPI_ROWDOT) {
    // part[m * gridDim.x + blockIdx.x] = sum over this tile's columns of acc * a(m, n)
#pragma unroll
    for (int i = 0; i < 2 * H; ++i) {
      int64_t m = m0 + (i < H ? ty * H + i : 16 * H + ty * H + (i - H));
      T s = (T)0;
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 2 * H; ++j) {
          int64_t n = n0 + (j < H ? tx * H + j : 16 * H + tx * H + (j - H));
          if (n < N) {
            int64_t off = A_T ? n * p.lda + m : m * p.lda + n;
            T v = __ldg(A + off);
            if (A2) { T w = __ldg(A2 + off); v = p.negate_a_diff ? F::sub(w, v) : F::sub(v, w); }
            s = F::fma(acc[i][j], v, s);
          }
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tx == 0 && m < M) C[m * (int64_t)gridDim.x + blockIdx.x] = s;
    }
    return;
  }
  // Read-modify-write epilogues first gather every C element they need (all loads in flight at
  // once -- interleaving each load with its store serialises 64 L2 round trips per thread), then
  // compute and store.
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    T old[H][2 * H];
    if (EPI == EPI_ACCUM || EPI == EPI_HESS) {
#pragma unroll
      for (int ii = 0; ii < H; ++ii) {
        const int64_t m = m0 + half * 16 * H + ty * H + ii;
#pragma unroll
        for (int j = 0; j < 2 * H; ++j) {
          const int64_t n = n0 + (j < H ? tx * H + j : 16 * H + tx * H + (j - H));
          old[ii][j] = (m < M && n < N) ? __ldcg(C + m * p.ldc + n) : (T)0;
        }
      }
    }
#pragma unroll
    for (int ii = 0; ii < H; ++ii) {
      const int i = half * H + ii;
      const int64_t m = m0 + half * 16 * H + ty * H + ii;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 2 * H; ++j) {
        const int64_t n = n0 + (j < H ? tx * H + j : 16 * H + tx * H + (j - H));
        if (n >= N) continue;
        T* c = C + m * p.ldc + n;
        const T v = acc[i][j];
        if (EPI == EPI_STORE) *c = F::mul(p.alpha, v);
        else if (EPI == EPI_ACCUM) *c = F::add(old[ii][j], F::mul(p.alpha, v));
        else if (EPI == EPI_HESS) *c = F::add(F::mul(old[ii][j], p.keep), F::div(v, p.count));
        else if (EPI == EPI_GAIN) {
          T d = F::sub(__ldg(p.x0 + m * p.ldx + n), __ldg(p.x1 + m * p.ldx + n));
          T t1 = F::mul(-F::mul(d, d), __ldg(p.x2 + n * p.x2_stride));
          T t2 = F::mul(F::mul((T)2, v), d);
          *c = F::sub(t1, t2);
        }
      }
    }
  }
}

template <typename T>
static inline GemmParams<T> gemm_params(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc,
                                        int64_t M, int64_t N, int64_t K) {
  GemmParams<T> p;
  p.A = A; p.lda = lda; p.A2 = nullptr; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.M = M; p.N = N; p.K = K; p.alpha = (T)1; p.keep = (T)0; p.count = (T)1;
  p.x0 = p.x1 = p.x2 = nullptr; p.ldx = 0; p.x2_stride = 1;
  p.strideA = p.strideB = p.strideC = 0; p.M_last = M; p.K_last = K;
  p.k_lo_from_n = p.k_hi_from_m = p.lower_only = p.negate_a_diff = 0;
  return p;
}

template <typename T> static inline int gemm_bm() { return 32 * GemmTile<T>::H; }

template <typename T, bool A_T, bool B_T, int EPI>
static inline int gemm_launch(const GemmParams<T>& p, int batch, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || batch <= 0) return SLK_OK;
  // fp64: 8x8 register tiles (128x128 CTA tile, 4 DFMA per shared-memory load) once the problem is
  // large enough to fill the GPU with such tiles; 4x4 (64x64) otherwise.  fp32 always 8x8.
  constexpr int HS = GemmTile<T>::H;
  if (false && sizeof(T) == 8 && EPI != EPI_ROWDOT && p.M >= 1024 && p.N >= 1024) {  // measured slower on B200 (1 CTA/SM)
    constexpr int HL = 4;
    dim3 grid((unsigned)ceil_div(p.N, 32 * HL), (unsigned)ceil_div(p.M, 32 * HL), (unsigned)batch);
    gemm_kernel<T, HL, A_T, B_T, EPI><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)ceil_div(p.N, 32 * HS), (unsigned)ceil_div(p.M, 32 * HS), (unsigned)batch);
    gemm_kernel<T, HS, A_T, B_T, EPI><<<grid, 256, 0, st>>>(p);
  }
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

// number of column tiles the ROWDOT epilogue writes per row
template <typename T> static inline int64_t rowdot_tiles(int64_t n) { return ceil_div(n, gemm_bm<T>()); }

}  // namespace slk
