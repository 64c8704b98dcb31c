/*
 * sleekit_b200 -- C ABI of the B200 (sm_100a) implementation of sleekit's
 * layer-wise quantization hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - matrices are dense row-major; W is [r, n] (r output rows, n input
 *     features), H is [n, n], X is [S, n];
 *   - every entry point only ENQUEUES work on `stream` (a cudaStream_t passed
 *     as void*) and returns 0, or a negative slk_status on a usage / launch
 *     error; slk_last_error() gives the thread-local message;
 *   - nothing here allocates or frees caller-visible memory: scratch is
 *     sized by the matching *_ws_bytes() query and passed in;
 *   - no C++ types, no exceptions, no torch types cross this boundary.
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the reference repo, Coloquinte/sleekit).
 */
#ifndef SLEEKIT_B200_H
#define SLEEKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLK_ABI_VERSION 1

typedef enum slk_status {
  SLK_OK = 0,
  SLK_ERR_ARG = -1,    /* bad shape / null pointer / unsupported option */
  SLK_ERR_CUDA = -2,   /* a CUDA runtime call failed */
  SLK_ERR_WS = -3      /* workspace too small */
} slk_status;

/* Rounding direction (sleekit/codebook.py:43-95, 155-188). */
typedef enum slk_round_mode { SLK_NEAREST = 0, SLK_UP = 1, SLK_DOWN = 2 } slk_round_mode;

/* Codebook passed by host pointer; `values` / `limits` are device pointers.
 * kind 0: UniformCodebook(size, lo, hi)                       codebook.py:4-41
 * kind 1: Codebook(values, limits)                            codebook.py:98-113 */
typedef struct slk_codebook {
  int32_t kind;
  int32_t size;
  double lo;            /* codebook.min(); uniform: also the zero point            */
  double hi;            /* codebook.max()                                           */
  double step;          /* uniform: (hi-lo)/(size-1) as a Python float, else 0      */
  const float* values;  /* table: [size] ascending, fp32                            */
  const float* limits;  /* table: [size-1] bin limits, fp32                         */
} slk_codebook;

int slk_abi_version(void);
const char* slk_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t slk_launch_count(void);
/* Fills sm_count / compute capability of the current device. */
int slk_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ---- K4: codebook rounding ------------------------------------------------
 * UniformCodebook.quantize_{index,value,up,down}   codebook.py:43-95
 * Codebook.quantize_{index,value,up,down}          codebook.py:155-188
 * out_val / out_idx may each be NULL.  out_idx element width is 1 byte for
 * size <= 256, 2 bytes for size <= 65536, else 4 (codebook.py:50-54).
 * Table codebooks always produce fp32 values (as the reference does). */
int slk_round_f32(const float* x, int64_t count, const slk_codebook* cb_host, int mode,
                  float* out_val, void* out_idx, void* stream);
int slk_round_f64(const double* x, int64_t count, const slk_codebook* cb_host, int mode,
                  void* out_val, void* out_idx, void* stream);

/* ---- scaling primitives ---------------------------------------------------
 * x viewed as [outer, len, inner]; s has `len` entries.
 * mode 0: out = x / s            apply_scaling              scaling.py:21-25
 * mode 1: out = x / (1 / s)      the de-scaling of quantize_with_scaling, scaling.py:80 */
int slk_scale_axis_f32(const float* x, int64_t outer, int64_t len, int64_t inner, const float* s,
                       int mode, float* out, void* stream);
int slk_scale_axis_f64(const double* x, int64_t outer, int64_t len, int64_t inner, const double* s,
                       int mode, double* out, void* stream);
/* compute_non_saturating_scaling on a [r, n] view, axis 0   scaling.py:44-55 */
int slk_row_noclip_scale_f32(const float* w, int64_t r, int64_t n, double cb_min, double cb_max,
                             float* out, void* stream);
int slk_row_noclip_scale_f64(const double* w, int64_t r, int64_t n, double cb_min, double cb_max,
                             double* out, void* stream);
/* compute_norm_scaling on a [r, n] view, axis 0             scaling.py:35-41 */
int slk_row_rms_scale_f32(const float* w, int64_t r, int64_t n, float* out, void* stream);
int slk_row_rms_scale_f64(const double* w, int64_t r, int64_t n, double* out, void* stream);

/* ---- K5: fused scale-grid search (MSE / diagonal-H) -------------------------
 * compute_min_mse_scaling with H None or 1-D                scaling.py:98-134
 * factors: [G] fp32 grid (np.linspace(..., dtype=float32), made by the caller)
 * hdiag: NULL, or [n] of fp32 (h_dtype 1) / fp64 (h_dtype 2)
 * out_scale [r] = init * best factor; out_err [r] (may be NULL) best error;
 * out_init [r] (may be NULL) the non-saturating scale. */
int slk_scale_search_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb_host,
                         const float* factors, int32_t G, const void* hdiag, int32_t h_dtype,
                         float* out_scale, float* out_err, float* out_init, void* stream);

/* Tests / A-B runs: 1 forces the direct op-chain kernel of slk_scale_search_f32, 0 restores the
 * default (exact threshold tables for uniform codebooks of <= 16 entries; identical results). */
int slk_debug_scale_search_direct(int on);

/* Host-only (no CUDA call): the exact breakpoints of a uniform codebook of <= 16 entries in the
 * scaled domain, X[k] = min{x : quantize_index(x) >= k} (codebook.py:43-54), k = 1..size-1, +inf
 * elsewhere; out16_host is a HOST array of 16 floats.  The scale-search tables and the sweep's
 * leaf round through these (idx >= k <=> x >= X[k]). */
int slk_codebook_breaks_host(const slk_codebook* cb_host, float* out16_host);

/* ---- host -> device upload of a symmetric Hessian ------------------------------------------
 * The experiments hand every layer's H = X^T X / n (statistics.py:87, read back from the data/ tree at
 * experiments/compare.py:37-53) to the hot path as a host array; it is symmetric, so only its block
 * upper triangle needs to cross PCIe.  h_host: page-locked [n, n]; h_dev: [n, n]; bs: rows per block
 * row, a multiple of 32 (bs >= n: plain full copy).  Strided async copies + one mirror kernel on
 * `stream`; slk_upload_symmetric_bytes returns the bytes that cross the bus. */
int slk_upload_symmetric_f32(const float* h_host, float* h_dev, int64_t n, int64_t bs, void* stream);
/* The two halves of slk_upload_symmetric_f32 for callers with a dedicated copy stream (copy on it, mirror on
 * the consumer's stream after an event): the DMA queue then never waits for a kernel. */
int slk_upload_symmetric_copy_f32(const float* h_host, float* h_dev, int64_t n, int64_t bs, void* stream);
int slk_mirror_symmetric_f32(float* h_dev, int64_t n, int64_t bs, void* stream);
size_t slk_upload_symmetric_bytes(int64_t n, int64_t bs);

/* ---- K6: H-weighted error ---------------------------------------------------
 * channelwise_error  ((W-Q) @ H * (W-Q)).sum(-1)            obq.py:89-95
 * _compute_mse with a 2-D H                                 scaling.py:91-95
 * q may be NULL, in which case `w` already holds the residual. */
size_t slk_hweighted_error_ws_bytes(int64_t r, int64_t n, int32_t elem_bytes);
int slk_hweighted_error_f32(const float* w, const float* q, const float* h, int64_t r, int64_t n,
                            void* ws, size_t ws_bytes, float* out, void* stream);
int slk_hweighted_error_f64(const double* w, const double* q, const double* h, int64_t r, int64_t n,
                            void* ws, size_t ws_bytes, double* out, void* stream);
/* mean of a vector (quantization_error's .mean())           obq.py:98-103 */
int slk_mean_f32(const float* v, int64_t count, float* out, void* stream);
int slk_mean_f64(const double* v, int64_t count, double* out, void* stream);

/* compute_gain                                               obq.py:220-231 */
int slk_gain_f32(const float* w, const float* q, const float* h, const float* cand, int64_t r,
                 int64_t n, float* out, void* stream);
int slk_gain_f64(const double* w, const double* q, const double* h, const double* cand, int64_t r,
                 int64_t n, double* out, void* stream);

/* full-H scale-grid search: compute_min_mse_scaling with a 2-D H  scaling.py:98-134
 * h_dtype 1: H fp32, 2: H fp64 (errors then accumulate in fp64).  With an fp32 H (n % 4 == 0) all G grid points
 * are ranked by one low-precision tensor pass and only the best-ranked ones of each row are evaluated by the
 * fp32-faithful product: the same scales and errors, bit for bit, as evaluating every point (options below). */
size_t slk_scale_search_fullh_ws_bytes(int64_t r, int64_t n, int32_t G, int32_t h_dtype);
int slk_scale_search_fullh_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb_host,
                               const float* factors, int32_t G, const void* h, int32_t h_dtype,
                               void* ws, size_t ws_bytes, float* out_scale, float* out_err,
                               void* stream);
/* The same; uncertified_rows (device int32, may be NULL) receives the number of rows for which the screened search
 * could not show, from the ranking values of the grid points it did not evaluate exactly, that its minimum is the
 * minimum over all G points (0 when every point is evaluated; 0 on every input tested so far). */
int slk_scale_search_fullh_checked_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb_host,
                                       const float* factors, int32_t G, const void* h, int32_t h_dtype,
                                       void* ws, size_t ws_bytes, float* out_scale, float* out_err,
                                       int32_t* uncertified_rows, void* stream);

/* ---- K1: calibration statistics ---------------------------------------------
 * Sleekit.add_batch                                         statistics.py:76-87
 * x: [S, n] with row stride ldx (samples are rows).  Updates, in place,
 *   mean = mean*keep + colsum(x)/new_count ;  hess = hess*keep + x^T x/new_count
 * keep = old_count/(old_count+S), new_count = old_count+S (both given by caller). */
size_t slk_hessian_accum_ws_bytes(int64_t S, int64_t n);
int slk_hessian_accum_f32(const float* x, int64_t S, int64_t n, int64_t ldx, float* hess,
                          float* mean, double keep, double new_count, void* ws, size_t ws_bytes,
                          void* stream);
/* remove_input_bias  H - outer(m, m)                        obq.py:14-25 */
int slk_remove_input_bias_f32(const float* h, const float* m, int64_t n, float* out, void* stream);
int slk_remove_input_bias_f64(const double* h, const double* m, int64_t n, double* out, void* stream);

/* ---- ordering -----------------------------------------------------------------
 * damp * mean(diag H) as an fp32 scalar                      obq.py:198
 * add_mode 0: writes dampval only.                                              */
int slk_damp_value_f32(const float* h, int64_t n, double damp, float* out_dampval, void* stream);
/* column sums over rows of |q(w)-w| (mode 0) or (q(w)-w)^2 (mode 1), fp32,
 * rows added in order                                        obq.py:60-69 */
int slk_col_resid_sums_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb_host,
                           int mode, float* out, void* stream);
/* keys[j] = -(double(H[j,j]) + dampval) [* colsum[j]]        obq.py:64,69,81 */
int slk_order_keys(const float* h, int64_t n, const float* dampval, const float* colsum,
                   double* keys, void* stream);
/* stable ascending argsort of fp64 keys -> int64 order        obq.py:64,69,81 */
int slk_argsort_f64(const double* keys, int64_t n, int64_t* order, void* stream);
/* act_order = "pivot": greedy pivoted-Cholesky ordering of an fp64 matrix    obq.py:140-166
 * (same fp64 operations on the same operands as the reference, hence the same order). */
size_t slk_pivot_order_ws_bytes(int64_t n);
int slk_pivot_order_f64(const double* h, int64_t n, void* ws, size_t ws_bytes, int64_t* order,
                        void* stream);
/* dst[:, j] = src[:, idx[j]] (gather) or dst[:, idx[j]] = src[:, j] (scatter)  obq.py:202,212-213 */
int slk_permute_cols_f32(const float* src, int64_t r, int64_t n, const int64_t* idx, int scatter,
                         float* dst, void* stream);

/* The two passes on either side of the sweep, fused (one pass over W each):
 * scatter 0: dst[:, j] = src[:, idx[j]] / s[row]            scaling.py:73 + obq.py:202
 * scatter 1: dst[:, idx[j]] = src[:, j] / (1 / s[row])      obq.py:212-213 + scaling.py:80
 * idx may be NULL (identity).  Same separately rounded divides as slk_scale_axis_f32. */
int slk_scale_permute_cols_f32(const float* src, int64_t r, int64_t n, const int64_t* idx,
                               const float* s, int scatter, float* dst, void* stream);

/* ---- K2: damp + permute + fp64 factor of the inverse --------------------------
 * H_opt = H + dampval*I; H_opt[order][:, order]; compute_hessian_chol
 *                                                             obq.py:198-205, 38-55
 * order may be NULL (identity), dampval may be NULL (0).  u64 / u32 may each be
 * NULL.  info (int32, device) receives 0, or 1 + the first pivot (in factor
 * order) that was not positive -> the Python layer raises LinAlgError. */
size_t slk_hinv_ws_bytes(int64_t n);
int slk_hinv_from_f32(const float* h, int64_t n, const int64_t* order, const float* dampval,
                      void* ws, size_t ws_bytes, double* u64, float* u32, int32_t* info,
                      void* stream);
int slk_hinv_from_f64(const double* h, int64_t n, void* ws, size_t ws_bytes, double* u64,
                      float* u32, int32_t* info, void* stream);

/* K2 without the triangular inverse (the form the fused quantize_opt path uses):
 * R = flip(cholesky(flip(H_opt[order][:, order]))) in fp64 (obq.py:46-50), i.e. H_opt = R R^T and
 * compute_hessian_chol's U = R^-1.  Outputs: r32 [n, n] fp32 rounding of R (upper, zeros below),
 * rt_hi / rt_lo [n, n] (may both be NULL) its transpose split into TF32 parts, hi + lo = R^T
 * (only the part on and below the diagonal is written: the K-major operand of the sweep's
 * tensor-core GEMMs) and ud32 [ceil(n/32), 32, 32] fp32: the
 * inverses of R's 32x32 diagonal blocks (= diagonal blocks of U; identity-padded for a ragged last
 * block).  One tile-task kernel, no dependent launches. */
size_t slk_chol_factor_ws_bytes(int64_t n);
int slk_chol_factor_f32(const float* h, int64_t n, const int64_t* order, const float* dampval,
                        void* ws, size_t ws_bytes, float* r32, float* rt_hi, float* rt_lo,
                        float* ud32, int32_t* info, void* stream);

/* The same factorisation for `njobs` matrices of ONE size in one launch sequence: the tile tasks of all
 * matrices share one ticket queue (ticket t -> matrix t % njobs), so a CTA whose next tile is not ready
 * works on another matrix instead of spinning -- what a layer set of many small, independent Hessians
 * needs (experiments/compare.py:50-135 loops over layers; obq.py:38-55 per layer).  Every pointer
 * argument is a HOST array of njobs device pointers (order / dampval / rt_hi / rt_lo: the array or
 * single entries may be NULL); ws[k] holds slk_chol_factor_ws_bytes(n) bytes. */
int slk_chol_factor_batched_f32(int32_t njobs, const float* const* h_host, int64_t n,
                                const int64_t* const* order_host, const float* const* dampval_host,
                                void* const* ws_host, float* const* r32_host, float* const* rt_hi_host,
                                float* const* rt_lo_host, float* const* ud32_host,
                                int32_t* const* info_host, void* stream);

/* One matrix factored by nranks GPUs of a node (one process per GPU; SURVEY 8 f-3, obq.py:38-55 for
 * the n = 28672 Hessian of BASELINE config 5).  Tile row i belongs to rank i % nranks; a finished tile
 * is pushed into every peer's workspace by NVLink peer stores + a system-scope flag, so consumers only
 * read local memory and the transfer overlaps the arithmetic; no collective inside.  Three
 * stream-ordered steps; the caller puts a barrier over the ranks ON THE SAME STREAM (an all-reduce of
 * one element) after the gather (every rank's flags are clear before a peer pushes) and after the
 * factor (all pushes have landed before the export reads).  ws: slk_chol_factor_ws_bytes(n) bytes from
 * slk_peer_alloc; peer_ws_host: HOST array of nranks device pointers (slk_peer_open), entry `rank`
 * being ws itself.  Every rank ends up with the complete factor; info must be max-reduced by the caller. */
int slk_chol_dist_gather_f32(const float* h, int64_t n, const int64_t* order, const float* dampval,
                             void* ws, size_t ws_bytes, int32_t* info, void* stream);
int slk_chol_dist_factor(int64_t n, void* ws, int32_t nranks, int32_t rank, void* const* peer_ws_host,
                         int32_t* info, void* stream);
int slk_chol_dist_export_f32(int64_t n, void* ws, float* r32, float* rt_hi, float* rt_lo, float* ud32,
                             void* stream);

/* Peer-visible device memory for slk_chol_dist_* (CUDA IPC): the only calls of this library that own
 * device memory.  handle_host: 64 bytes (cudaIpcMemHandle_t) exchanged between the ranks' processes. */
int slk_peer_alloc(size_t bytes, void** dptr_host, void* handle_host);
int slk_peer_open(const void* handle_host, void** dptr_host);
int slk_peer_close(void* dptr);
int slk_peer_free(void* dptr);

/* In-place sum all-reduce of count floats (count % 4 == 0) over nranks peer-visible buffers (peers_host: HOST
 * array of device pointers from slk_peer_alloc / slk_peer_open, entry `rank` = this rank's own): rank r sums
 * the r-th slice over all ranks with NVLink peer loads and stores the result into every copy.  The caller puts
 * a barrier over the ranks on the same stream before and after.  Exchange step of the sample-sharded statistics
 * (Sleekit.add_batch summed over ranks, statistics.py:76-87). */
int slk_peer_allreduce_f32(void* const* peers_host, int32_t nranks, int32_t rank, int64_t count, void* stream);

/* Block-upper-triangle packing of a symmetric matrix (layout / size of slk_upload_symmetric_bytes):
 * the exchange format of the sample-sharded statistics (statistics.py:76-87 summed over ranks): pack
 * scale * H, all-reduce the packed buffer in place, unpack (+ mirror) into H. */
int slk_sym_pack_f32(const float* h, int64_t n, int64_t bs, float scale, float* packed, void* stream);
int slk_sym_unpack_f32(const float* packed, int64_t n, int64_t bs, float scale, float* h, void* stream);

/* Streams for the layer-set driver: one real CUDA stream per independent layer, with a priority
 * (0 default, negative = higher, clamped to the device's range). */
int slk_stream_create(int priority, void** stream_host);
int slk_stream_destroy(void* stream);
/* Process-wide tuning knobs, by name: "sweep_ctas" = CTAs wanted per macro-block sweep launch (0 = the
 * single-layer default; a layer set that runs many sweeps side by side prefers fewer, taller CTAs).
 * Full-H search: "fullh_topk" = exactly evaluated candidates per row, 0 | 4 | 8 | 16 (0: no screening, every grid
 * point is evaluated; default 8), "fullh_compact" 1 | 0 (only the candidates within 2^-5 of the best-ranked one,
 * as a compacted list of at most topk per row on average | a fixed topk per row), "fullh_bf16" 1 | 0 (ranking
 * pass on bf16 | TF32 operands), "fullh_bn" 256 | 128 and "fullh_ctas" 2 | 1 (tile width and CTAs per SM of the
 * ranking product). */
int slk_set_option(const char* name, int64_t value);

/* Development aid: stores %globaltimer (ns, uint64) into *slot when the stream reaches the call. */
int slk_debug_timestamp(void* slot, void* stream);
/* Development aid: per-tile-task trace of the Cholesky kernel (8 int64 per task); NULL disables. */
int slk_debug_chol_trace(void* buf);
/* Same for the macro-block sweep kernel: 8 int64 phase clocks per 32-column block of CTA 0. */
int slk_debug_sweep_trace(void* buf);

/* ---- K3: GPTQ / OBQ sweep ------------------------------------------------------
 * _quantize_opt_block / _quantize_opt_core                    obq.py:106-137
 * q: [r, n] in: scaled, column-permuted weights; out: quantized values.
 * e: [r, n] out: scaled residuals.  u32: fp32 rounding of the factor.  leaf <= 32.
 * exact_leaf = 0: all-fp32 leaf arithmetic (default, fast; u64 may be NULL);
 * exact_leaf = 1: the leaf reproduces the reference's mixed fp64/fp32 op sequence bit for bit
 *                 from the fp64 factor u64 (obq.py:114-118). */
int slk_gptq_sweep_f32(float* q, float* e, int64_t r, int64_t n, const double* u64,
                       const float* u32, const slk_codebook* cb_host, int32_t leaf,
                       int32_t fanout, int32_t exact_leaf, void* stream);

/* Same sweep from the Cholesky factor (slk_chol_factor_f32) instead of the inverse factor:
 * block J is formed as W[:, J] + ((W-Q)[:, :a] R[:a, J]) U_JJ, which equals obq.py:137's
 * W[:, J] - E[:, :a] U[:a, J].  q in/out as above; d [r, n] out: W - Q (scaled, permuted domain). */
/* Columns are cut into macro blocks of 256; inside one the fused kernel propagates locally,
 * between them one tcgen05 GEMM (fp32-faithful 3xTF32) accumulates D[:, s:e] R[s:e, e:] for all
 * later columns (needs rt_hi / rt_lo and the workspace; without them, or for n < 512, one fused
 * launch).  The leaf writes D already split into TF32 parts, so a macro block costs two launches. */
size_t slk_gptq_sweep_r_ws_bytes(int64_t r, int64_t n);
int slk_gptq_sweep_r_f32(float* q, float* d, int64_t r, int64_t n, const float* r32,
                         const float* rt_hi, const float* rt_lo, const float* ud32,
                         const slk_codebook* cb_host, void* ws, size_t ws_bytes, void* stream);
/* The same sweep, also returning err_sums [r, 2] = per row (sum_i E_i^2, sum_i (W-Q)_i^2), E_i =
 * obq.py:114's scaled residual.  Since W - Q = E U and U H_opt U^T = I, (W-Q) H_opt (W-Q)^T =
 * sum E^2 exactly, so channelwise_error (obq.py:89-95) under the undamped H the factor came from is
 * sum E^2 - damp_abs * sum (W-Q)^2: the layer error falls out of the sweep without the 2 r n^2
 * product.  slk_sweep_error_f32 finishes it: rows_out[row] = row_scale[row]^2 * (that)  (the error
 * of the de-scaled weights; row_scale may be NULL), mean_out = their mean (quantization_error,
 * obq.py:98-103); damp_abs is the device scalar of slk_damp_value_f32 (NULL: 0). */
int slk_gptq_sweep_r_err_f32(float* q, float* d, int64_t r, int64_t n, const float* r32,
                             const float* rt_hi, const float* rt_lo, const float* ud32,
                             const slk_codebook* cb_host, void* ws, size_t ws_bytes,
                             float* err_sums, void* stream);
int slk_sweep_error_f32(const float* err_sums, const float* row_scale, const float* damp_abs,
                        int64_t r, float* rows_out, float* mean_out, void* stream);

/* ---- K7: best-first local search ------------------------------------------------
 * quantize_local_search / LocalSearchQuantizer                obq.py:234-358
 * q: [r, n] in/out quantized values (must lie on the codebook). */
size_t slk_local_search_ws_bytes(int64_t r, int64_t n);
int slk_local_search_f32(const float* w, float* q, const float* h, int64_t r, int64_t n,
                         const slk_codebook* cb_host, int32_t moves, void* ws, size_t ws_bytes,
                         void* stream);
/* slk_local_search_f32 that also returns err_sums [r, 2] = (channelwise_error of each row after its moves under
 * h (obq.py:89-95), 0) -- the row-sum layout of slk_sweep_error_f32, which applies row scales and the mean.
 * The search keeps p = (Q - W) H current, so the error is p . (Q - W): no second 2 r n^2 product. */
int slk_local_search_err_f32(const float* w, float* q, const float* h, int64_t r, int64_t n,
                             const slk_codebook* cb_host, int32_t moves, void* ws, size_t ws_bytes,
                             float* err_sums, void* stream);
/* The same moves in instalments (LocalSearchQuantizer.do_move, obq.py:338-346: one move per call): the
 * workspace keeps P = (Q - W) H between calls.  resume = 0 forms it, resume = 1 continues from the P the
 * previous call left for the same q, w, h, so a move costs one pass over the gains instead of a GEMM. */
int slk_local_search_step_f32(const float* w, float* q, const float* h, int64_t r, int64_t n,
                              const slk_codebook* cb_host, int32_t moves, void* ws, size_t ws_bytes,
                              int32_t resume, void* stream);

/* bias correction  bias += ((W - Wq) * mean).sum(1)          statistics.py:187-190 */
/* _compute_mse with H None or 1-D                              scaling.py:84-95
 * out[row] = sum_j h[j] * e[row, j]^2 (h NULL: sum of squares); the 2-D case is slk_hweighted_error_*. */
int slk_row_wsq_f32(const float* e, const float* h, int64_t r, int64_t n, float* out, void* stream);
int slk_row_wsq_f64(const double* e, const double* h, int64_t r, int64_t n, double* out, void* stream);
int slk_row_wsq_f32_h64(const float* e, const double* h, int64_t r, int64_t n, double* out, void* stream);

int slk_bias_delta_f32(const float* w, const float* wq, const float* mean, int64_t r, int64_t n,
                       float* delta, void* stream);

/* ---- tensor-core GEMM core (tcgen05 + TMEM + TMA, 3xTF32 split, fp32-faithful) -----------------
 * D[M, N] = (A [- A2])[M, K] * B[N, K]^T, both operands fp32 K-major with pitches lda / ldb.
 * epilogue 0: C = alpha*D        1: C += alpha*D (obq.py:137 with alpha = -1)
 *          2: C = C*keep + D/count (statistics.py:82-87)
 *          3: C[m, tile] = sum_n D[m, n] * rowdot[m, n] over each 128-column tile (obq.py:95)
 * Used internally by the K1 / K3 / K6 / K7 entry points when shapes and alignment allow
 * (pitches multiple of 4, 16-byte aligned bases); exported for tests and benchmarks. */
size_t slk_tc_gemm_ws_bytes(int64_t M, int64_t N, int64_t K);
int slk_tc_gemm_f32(int32_t epilogue, const float* a, const float* a2, int64_t lda, const float* b,
                    int64_t ldb, float* c, int64_t ldc, const float* rowdot, int64_t ldr, int64_t M,
                    int64_t N, int64_t K, float alpha, float keep, float count, void* ws,
                    size_t ws_bytes, int32_t* error_flag, void* stream);

/* Self-test (tests only): for each of `count` divisors, compares the kernels' fast exact divide
 * with the IEEE divide over all 2^32 dividends; mismatches[i] must come back 0. */
int slk_selftest_fastdiv_f32(const float* divisors, int32_t count, uint64_t* mismatches, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLEEKIT_B200_H */
