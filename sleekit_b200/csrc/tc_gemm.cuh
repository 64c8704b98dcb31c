// Interface of the tcgen05 GEMM core (tc_gemm.cu) for the other translation units.
#pragma once

#include "common.cuh"

namespace slk {

enum TcEpilogue { TC_STORE = 0, TC_ACCUM = 1, TC_HESS = 2, TC_ROWDOT = 3, TC_HESS_SYM = 4, TC_SYM_PART = 5 };

struct TcParams {
  float* C; int64_t ldc;         // TC_STORE / TC_ACCUM / TC_HESS: [M, N]; TC_ROWDOT: partials [M, tiles_n]
  const float* R; const float* R2; int64_t ldr;   // TC_ROWDOT: rows of (R - R2) are dotted with the product rows
  int64_t M, N, K;
  float alpha, keep, count;
  int* error_flag;               // set (and the kernel traps) if a barrier wait exceeds its budget
  int kb_per_split = 0;          // TC_SYM_PART: k-blocks (of 32) per blockIdx.z; plane z of C is C + z*M*ldc
  const int* m_limit = nullptr;  // device scalar: row tiles starting at or beyond *m_limit exit at once (a row count only
                                 // known on the device; M is then the capacity the grid is sized for)
};

constexpr int TC_TILE_N = 128;

// true when the tensor-core path can take these operands (16-byte aligned, pitches multiple of 4,
// driver exposes cuTensorMapEncodeTiled, not disabled by SLK_DISABLE_TC=1)
bool tc_gemm_usable(const void* a, int64_t lda, const void* b, int64_t ldb);
size_t tc_gemm_ws_bytes(int64_t M, int64_t N, int64_t K);
// D = (A [- A2]) * B^T; A [M, K], B [N, K] fp32 K-major
int tc_gemm_f32(int epi, const float* A, const float* A2, int64_t lda, const float* B, int64_t ldb, TcParams p,
                void* ws, size_t ws_bytes, cudaStream_t st);
// same from operands already split into TF32 hi / lo parts (see split_tf32 below)
int tc_gemm_presplit_f32(int epi, const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi,
                         const float* b_lo, int64_t ldb, TcParams p, cudaStream_t st);
// one kind::tf32 pass of TF32-representable operands with the row-dot epilogue (tiles of bn = 128 | 256 columns,
// ctas = 1 | 2 CTAs per SM): ~1e-3 relative, for screening only
int tc_gemm_screen_f32(const float* a, int64_t lda, const float* b, int64_t ldb, TcParams p, int bn, int ctas,
                       cudaStream_t st);
// the same with bf16 operands (kind::f16; pitches in elements, multiples of 8; p.R = the bf16 rows): ~4e-3 relative
int tc_gemm_screen_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, TcParams p, int bn, int ctas,
                        cudaStream_t st);
int tc_split_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, int64_t ldo, float* hi, float* lo,
                 cudaStream_t st);
// the split the tensor path expects: hi = v with 13 low mantissa bits cleared, lo = RN_tf32(v - hi)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
  const float l = __fsub_rn(v, hi);
  lo = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
}
// same with A given transposed: At [K, M] (pitch ldat); used by K1 (X^T X with X [S, n])
int tc_gemm_at_f32(int epi, const float* At, int64_t ldat, TcParams p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t tc_gemm_at_ws_bytes(int64_t M, int64_t K);

}  // namespace slk
