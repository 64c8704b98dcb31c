// fp64 GEMM on the FP64 tensor path (mma.sync m8n8k4 f64) for K2's trailing updates and inverse
// products (tcgen05 has no fp64 kind).  Same parameter block and structural options as gemm.cuh
// (batching over blockIdx.z, ragged last batch, triangular k-range clipping, lower tiles only),
// restricted to what K2 needs: M, N multiples of 64, K multiple of 16, 16-byte aligned operands.
//
//   C[M, N] (=|+=) alpha * A[M, K] * op(B),   A row-major (k contiguous)
//   B_T = true : B stored [N, K] (k contiguous)     -- SYRK  A22 -= P P^T
//   B_T = false: B stored [K, N] (n contiguous)     -- doubling products
//
// CTA = 64x64 tile, 4 warps, each warp a 32x32 sub-tile = 4x4 DMMA tiles (32 fp64 accumulators per
// thread).  Operand tiles of 16 k-columns go through a 3-stage cp.async ring; the shared-memory
// pitches (20 / 68 doubles) make every fragment load conflict-free.  Per 16-k tile and warp:
// 32 LDS.64 feed 64 DMMAs (0.5 B of shared-memory traffic per FMA, 8x less than the 4x4
// register-tiled DFMA kernel, which was shared-memory bound at ~1/3 of the FP64 peak).
#pragma once

#include "gemm.cuh"

namespace slk {

constexpr int DM_BM = 64, DM_BN = 64, DM_BK = 16, DM_ST = 3;
constexpr int DM_LDA = DM_BK + 4;    // pitch of k-contiguous tiles (A, and B when B_T)
constexpr int DM_LDBN = DM_BN + 4;   // pitch of n-contiguous B tiles

struct DmSmem {
  double A[DM_ST][DM_BM * DM_LDA];
  double B[DM_ST][DM_BK * DM_LDBN > DM_BN * DM_LDA ? DM_BK * DM_LDBN : DM_BN * DM_LDA];
};

__device__ __forceinline__ void dm_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}

__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <bool B_T, int EPI>
__global__ void __launch_bounds__(128) dgemm_mma_kernel(GemmParams<double> p) {
  extern __shared__ __align__(16) unsigned char dm_raw[];
  DmSmem& sm = *reinterpret_cast<DmSmem*>(dm_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.y * DM_BM, n0 = (int64_t)blockIdx.x * DM_BN;
  const int z = blockIdx.z;
  const bool last = (z == (int)gridDim.z - 1);
  const int64_t M = last ? p.M_last : p.M;
  const int64_t K = last ? p.K_last : p.K;
  if (m0 >= M || n0 >= p.N) return;
  if (p.lower_only && n0 > m0 + DM_BM - 1) return;
  const double* __restrict__ A = p.A + (int64_t)z * p.strideA;
  const double* __restrict__ B = p.B + (int64_t)z * p.strideB;
  double* __restrict__ C = p.C + (int64_t)z * p.strideC;

  int64_t kbeg = 0, kend = K;
  if (p.k_lo_from_n) kbeg = (n0 / DM_BK) * DM_BK;
  if (p.k_hi_from_m) { const int64_t e = m0 + DM_BM; kend = e < K ? e : K; }
  if (kbeg > kend) kbeg = kend;
  const int nk = (int)((kend - kbeg) / DM_BK);

  auto load_tile = [&](int t) {
    const int st = t % DM_ST;
    const int64_t k0 = kbeg + (int64_t)t * DM_BK;
    // A tile: 64 rows x 16 k = 512 x 16-byte pieces, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128, row = idx >> 3, piece = idx & 7;
      dm_cp16(&sm.A[st][row * DM_LDA + piece * 2], A + (m0 + row) * p.lda + k0 + piece * 2);
    }
    if (B_T) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 128, row = idx >> 3, piece = idx & 7;
        dm_cp16(&sm.B[st][row * DM_LDA + piece * 2], B + (n0 + row) * p.ldb + k0 + piece * 2);
      }
    } else {
      // B tile: 16 k-rows x 64 n = 512 pieces
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 128, row = idx >> 5, piece = idx & 31;
        dm_cp16(&sm.B[st][row * DM_LDBN + piece * 2], B + (k0 + row) * p.ldb + n0 + piece * 2);
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;   // fragment row (or column) and k within the 8x4 / 4x8 fragment

  for (int t = 0; t < DM_ST - 1; ++t) {
    if (t < nk) load_tile(t);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int t = 0; t < nk; ++t) {
    asm volatile("cp.async.wait_group %0;" ::"n"(DM_ST - 2) : "memory");
    __syncthreads();
    if (t + DM_ST - 1 < nk) load_tile(t + DM_ST - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int st = t % DM_ST;
    const double* As = sm.A[st];
    const double* Bs = sm.B[st];
#pragma unroll
    for (int ks = 0; ks < DM_BK; ks += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(wm + 8 * i + fr) * DM_LDA + ks + fk];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = B_T ? Bs[(wn + 8 * j + fr) * DM_LDA + ks + fk] : Bs[(ks + fk) * DM_LDBN + wn + 8 * j + fr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma8x8x4(acc[i][j], a[i], b[j]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // C fragment: thread holds row fr, columns 2*fk, 2*fk+1 of each 8x8 tile
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + 8 * i + fr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + wn + 8 * j + 2 * fk;
      double2* c = reinterpret_cast<double2*>(C + m * p.ldc + n);
      double2 o;
      if (EPI == EPI_ACCUM) {
        o = *c;
        o.x = __dadd_rn(o.x, __dmul_rn(p.alpha, acc[i][j][0]));
        o.y = __dadd_rn(o.y, __dmul_rn(p.alpha, acc[i][j][1]));
      } else {
        o.x = __dmul_rn(p.alpha, acc[i][j][0]);
        o.y = __dmul_rn(p.alpha, acc[i][j][1]);
      }
      *c = o;
    }
  }
}

static inline bool dgemm_mma_ok(const GemmParams<double>& p) {
  auto a16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  return p.M % 64 == 0 && p.N % 64 == 0 && p.K % 16 == 0 && p.M_last % 64 == 0 && p.K_last % 16 == 0 && !p.A2 &&
         p.lda % 2 == 0 && p.ldb % 2 == 0 && p.ldc % 2 == 0 && a16(p.A) && a16(p.B) && a16(p.C) &&
         p.strideA % 2 == 0 && p.strideB % 2 == 0 && p.strideC % 2 == 0;
}

template <bool B_T, int EPI>
static inline int dgemm_mma_launch(const GemmParams<double>& p, int batch, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0 || batch <= 0) return SLK_OK;
  auto kern = dgemm_mma_kernel<B_T, EPI>;
  SLK_SMEM_ATTR_ONCE(kern, (int)sizeof(DmSmem));
  dim3 grid((unsigned)(p.N / DM_BN), (unsigned)(p.M / DM_BM), (unsigned)batch);
  kern<<<grid, 128, sizeof(DmSmem), st>>>(p);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

}  // namespace slk
