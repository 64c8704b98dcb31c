// K7: best-first local search.
//   LocalSearchQuantizer / quantize_local_search            obq.py:234-358
//   compute_gain                                            obq.py:220-231
// The reference keeps six [r, n] arrays and patches both gain matrices after every flip
// (obq.py:299-336).  Rows never interact, so here one CTA owns one row and keeps only
//   codes[n] (uint16)  and  p[n] = ((Q - W) @ H)[row]  (fp32)
// in shared memory; gains are recomputed from p on every move,
//   gain(j, D) = -D^2 * H[j,j] - 2 * p[j] * D       (obq.py:231)
// the best up / best down flips are found with a block arg-max (first index wins ties, as
// np.argmax), the reference's selection rule (obq.py:339-342) picks one, and p is advanced by
// the one row of H the flip touches: p += delta * H[b, :].  Per move and row the traffic is one
// row of H (4n bytes, L2-resident across rows) instead of ~20n bytes of gain patches.
#include "gemm.cuh"
#include "tc_gemm.cuh"

namespace slk {

__global__ void __launch_bounds__(256) ls_diag_kernel(const float* __restrict__ h, int64_t n, float* __restrict__ d) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) d[j] = h[j * (n + 1)];
}

struct Best { float v; int i; };

__device__ __forceinline__ Best best_of(Best a, Best b) {
  // larger value wins; on equal values the smaller index wins
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = best_of(x, y);
  }
  return x;
}

__global__ void __launch_bounds__(256) local_search_kernel(float* __restrict__ Q, float* __restrict__ P,
                                                           const float* __restrict__ H, const float* __restrict__ hdiag,
                                                           int64_t r, int64_t n, DevGrid<float> g, int moves,
                                                           int keep_p, const float* __restrict__ W = nullptr,
                                                           float2* __restrict__ err_sums = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* p = (float*)smem_raw;                       // [n]
  uint16_t* code = (uint16_t*)(p + n);               // [n]
  __shared__ Best red_up[8], red_dn[8];
  __shared__ float red_err[8];
  __shared__ int s_col, s_newk;
  __shared__ float s_delta;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float NEG_INF = __int_as_float(0xff800000);
  const int top = g.size - 1;

  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    for (int64_t j = tid; j < n; j += blockDim.x) {
      p[j] = P[row * n + j];
      code[j] = (uint16_t)grid_index(g, Q[row * n + j]);
    }
    __syncthreads();
    for (int mv = 0; mv < moves; ++mv) {
      Best bu, bd;
      bu.v = NEG_INF; bu.i = 0x7fffffff; bd = bu;
      for (int64_t j = tid; j < n; j += blockDim.x) {
        const int k = code[j];
        const float v = grid_value_of_index(g, k);
        const float hj = __ldg(hdiag + j);
        const float p2 = __fmul_rn(2.0f, p[j]);
        const float du = k < top ? __fsub_rn(grid_value_of_index(g, k + 1), v) : 0.0f;
        const float dd = k > 0 ? __fsub_rn(grid_value_of_index(g, k - 1), v) : 0.0f;
        const float gu = __fsub_rn(__fmul_rn(-__fmul_rn(du, du), hj), __fmul_rn(p2, du));
        const float gd = __fsub_rn(__fmul_rn(-__fmul_rn(dd, dd), hj), __fmul_rn(p2, dd));
        if (gu > bu.v) { bu.v = gu; bu.i = (int)j; }
        if (gd > bd.v) { bd.v = gd; bd.i = (int)j; }
      }
      bu = warp_best(bu); bd = warp_best(bd);
      if (lane == 0) { red_up[wid] = bu; red_dn[wid] = bd; }
      __syncthreads();
      if (tid == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { bu = best_of(bu, red_up[k]); bd = best_of(bd, red_dn[k]); }
        const bool up = (bu.v > bd.v) && (bu.v > 0.0f);      // obq.py:341
        const bool down = !up && (bd.v > 0.0f);               // obq.py:342
        if (up || down) {
          const int col = up ? bu.i : bd.i;
          const int k = code[col];
          const int nk = up ? k + 1 : k - 1;
          s_col = col; s_newk = nk;
          s_delta = __fsub_rn(grid_value_of_index(g, nk), grid_value_of_index(g, k));
        } else {
          s_col = -1;
        }
      }
      __syncthreads();
      const int col = s_col;
      if (col < 0) break;  // local optimum: later moves would not change this row either
      const float delta = s_delta;
      const float* hrow = H + (int64_t)col * n;
      for (int64_t j = tid; j < n; j += blockDim.x) p[j] = __fmaf_rn(delta, __ldg(hrow + j), p[j]);
      if (tid == 0) code[col] = (uint16_t)s_newk;
      __syncthreads();
    }
    float part = 0.0f;
    for (int64_t j = tid; j < n; j += blockDim.x) {
      const float v = grid_value_of_index(g, code[j]);
      Q[row * n + j] = v;
      if (keep_p) P[row * n + j] = p[j];             // (Q - W) H of the row after its moves, for a later resume
      if (err_sums) part = __fmaf_rn(p[j], __fsub_rn(v, __ldg(W + row * n + j)), part);
    }
    if (err_sums) {
      // channelwise_error of the row after its moves (obq.py:89-95): ((Q - W) H) . (Q - W) = p . (Q - W); p is kept
      // current by every move, so the 2 r n^2 product of K6 is not needed again.  Fixed order: deterministic.
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
      if (lane == 0) red_err[wid] = part;
      __syncthreads();
      if (tid == 0) {
        float tot = red_err[0];
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) tot = __fadd_rn(tot, red_err[k]);
        err_sums[row] = make_float2(tot, 0.0f);
      }
    }
    __syncthreads();
  }
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace slk

using namespace slk;

extern "C" {

size_t slk_local_search_ws_bytes(int64_t r, int64_t n) {
  size_t bytes = align256((size_t)r * n * sizeof(float)) + align256((size_t)n * sizeof(float));
  if (n % 4 == 0 && n >= 32) bytes += tc_gemm_ws_bytes(r, n, n);   // TF32 parts for the tensor-core init product
  return bytes;
}

static int local_search_impl(const float* w, float* q, const float* h, int64_t r, int64_t n, const slk_codebook* cb,
                             int32_t moves, void* ws, size_t ws_bytes, int resume, int keep_p, void* stream,
                             float* err_sums = nullptr);

int slk_local_search_f32(const float* w, float* q, const float* h, int64_t r, int64_t n, const slk_codebook* cb,
                         int32_t moves, void* ws, size_t ws_bytes, void* stream) {
  return local_search_impl(w, q, h, r, n, cb, moves, ws, ws_bytes, 0, 0, stream);
}

/* The same, also returning err_sums [r, 2] = (channelwise_error of the row after its moves under h, 0): the layout
   slk_sweep_error_f32 takes (row scales and the mean are applied there).  obq.py:89-95 without a second product. */
int slk_local_search_err_f32(const float* w, float* q, const float* h, int64_t r, int64_t n, const slk_codebook* cb,
                             int32_t moves, void* ws, size_t ws_bytes, float* err_sums, void* stream) {
  SLK_REQUIRE(err_sums != nullptr && moves >= 1, "err_sums needs at least one move");
  return local_search_impl(w, q, h, r, n, cb, moves, ws, ws_bytes, 0, 0, stream, err_sums);
}

/* The same moves in instalments (LocalSearchQuantizer.do_move, obq.py:338-346, one move per call): the
   workspace keeps P = (Q - W) H between calls; resume = 0 forms it (2 r n^2 flop), resume = 1 continues from
   the P a previous call left for the SAME q, w, h -- a move then costs one pass over the gains, not a GEMM. */
int slk_local_search_step_f32(const float* w, float* q, const float* h, int64_t r, int64_t n, const slk_codebook* cb,
                              int32_t moves, void* ws, size_t ws_bytes, int32_t resume, void* stream) {
  return local_search_impl(w, q, h, r, n, cb, moves, ws, ws_bytes, resume ? 1 : 0, 1, stream);
}

}  // extern "C"

static int local_search_impl(const float* w, float* q, const float* h, int64_t r, int64_t n, const slk_codebook* cb,
                             int32_t moves, void* ws, size_t ws_bytes, int resume, int keep_p, void* stream,
                             float* err_sums) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(r >= 0 && n >= 1 && moves >= 0, "bad arguments");
  SLK_REQUIRE(cb->size <= 65536, "local search supports codebooks up to 65536 entries");
  if (r == 0 || moves == 0) return SLK_OK;
  SLK_REQUIRE(w && q && h, "NULL pointer");
  SLK_REQUIRE(ws && ws_bytes >= slk_local_search_ws_bytes(r, n), "workspace too small");
  const size_t smem = (size_t)n * 6 + 16;
  SLK_REQUIRE(smem <= 220 * 1024, "row of %lld columns does not fit in shared memory", (long long)n);
  cudaStream_t st = (cudaStream_t)stream;
  float* P = (float*)ws;
  float* hdiag = (float*)((char*)ws + align256((size_t)r * n * sizeof(float)));
  ls_diag_kernel<<<(int)ceil_div(n, 256), 256, 0, st>>>(h, n, hdiag);
  SLK_LAUNCH_CHECK();
  // P = (Q - W) @ H                                           obq.py:229-231
  void* gws = (char*)ws + align256((size_t)r * n * sizeof(float)) + align256((size_t)n * sizeof(float));
  if (resume) {
    rc = SLK_OK;                                               // P of the previous instalment is still valid
  } else if (n % 4 == 0 && n >= 32 && tc_gemm_usable(q, n, h, n) && (uintptr_t)w % 16 == 0) {
    // tcgen05, fp32-faithful 3xTF32; H is symmetric, hence its own K-major B operand
    TcParams tp;
    tp.C = P; tp.ldc = n; tp.R = nullptr; tp.R2 = nullptr; tp.ldr = 0; tp.M = r; tp.N = n; tp.K = n;
    tp.alpha = 1.0f; tp.keep = 0.0f; tp.count = 1.0f; tp.error_flag = nullptr;
    rc = tc_gemm_f32(TC_STORE, q, w, n, h, n, tp, gws, tc_gemm_ws_bytes(r, n, n), st);
  } else {
    GemmParams<float> p = gemm_params<float>(q, n, h, n, P, n, r, n, n);
    p.A2 = w;
    rc = gemm_launch<float, false, false, EPI_STORE>(p, 1, st);
  }
  if (rc) return rc;
  if (smem > 48 * 1024)
    SLK_CUDA(cudaFuncSetAttribute(local_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(r < (int64_t)sm_count() * 8 ? r : (int64_t)sm_count() * 8);
  local_search_kernel<<<grid, 256, smem, st>>>(q, P, h, hdiag, r, n, make_grid<float>(cb), moves, keep_p, w,
                                               reinterpret_cast<float2*>(err_sums));
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
