// K3: the GPTQ / OBQ column sweep.
//   _quantize_opt_core   (leaf, <= 32 columns)      obq.py:106-118
//   _quantize_opt_block  (8-ary lazy batching)      obq.py:121-137
// Two entry points:
//  * slk_gptq_sweep_r_f32 (what quantize_opt / compute_obq_scaling / Sleekit.quantize run): the
//    sweep from the Cholesky factor R (H_opt = R R^T, chol_dag.cu) -- macro blocks of 256 columns
//    swept by sweep_macro_kernel (shared-memory resident, four lanes per row, look-ahead product),
//    one tcgen05 GEMM per macro block (two levels for n > 4096) pushing D = W - Q to the later
//    columns.  All fp32 (parity-safe).
//  * slk_gptq_sweep_f32 (the public _quantize_opt_core/_block with a given Hinv): the sweep from
//    the inverse factor U; exact_leaf = 1 reproduces the reference's mixed fp64/fp32 leaf bit for
//    bit: column i is quantised, its scaled residual r = (w - q) / U[i,i] is formed in fp64 (as
//    numpy does: the divisor is an fp64 scalar) and columns j > i take
//    q_j <- fp32(fp64(q_j) - r * U[i,j]) -- the op sequence of obq.py:114-118; trailing updates
//    Q[:, b:end] -= E[:, a:b] @ U[a:b, b:end] are fp32 GEMMs with exact-product fmaf accumulation.
#include "gemm.cuh"
#include "tc_gemm.cuh"

#include <stdlib.h>

static int64_t g_opt_sweep_ctas = 0;     // slk_set_option("sweep_ctas", v)
extern int64_t g_opt_fullh_topk, g_opt_fullh_bn, g_opt_fullh_bf16, g_opt_fullh_ctas, g_opt_fullh_compact;   // dense.cu

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
                                             float* __restrict__ qrow, float* __restrict__ erow,
                                             const float* __restrict__ w0 = nullptr, float* dsm = nullptr,
                                             float* esd = nullptr) {
#pragma unroll 1
  for (int t = 0; t < 8; ++t) {
    const int i = i0 + t;
    if (i >= width) return;
    const float w = q[0];
    const float qq = fastq ? uniform_value_fast(g, fstep, w) : grid_value(g, w);
    float res;
    if (w0) {
      res = __fmul_rn(__fsub_rn(w, qq), sh.Uy[i]);   // R form: Uy = R[i][i] = 1 / U[i][i]
    } else {
      FastDivF fd;
      fd.d = sh.U[i][i]; fd.y = sh.Uy[i]; fd.ok = sh.Uok[i];
      res = fastdiv(__fsub_rn(w, qq), fd);
    }
    qrow[i] = qq;
    const float dv = w0 ? __fsub_rn(w0[i], qq) : res;   // R form: D = W - Q of the ORIGINAL weight
    erow[i] = dv;
    if (dsm) dsm[i] = dv;
    if (esd) {                                         // row sums of res^2 and D^2 (error identity, below)
      esd[0] = __fmaf_rn(res, res, esd[0]);
      esd[1] = __fmaf_rn(dv, dv, esd[1]);
    }
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

// ---- fused sweep: one launch per layer --------------------------------------------------------
// Rows never interact, so a CTA that owns R rows can walk all columns by itself: for every
// 32-column block it (1) forms the block left-looking, Qb = W[:, blk] - E[:, :a] @ U[:a, blk]
// (all 256 threads; E and U stream through a 4-stage cp.async ring, E being this CTA's own rows
// written earlier and read back through L2), (2) sweeps the block with one thread per row (the
// fp32 leaf above).  No trailing read-modify-write of Q, no dependent launches: 1 launch instead
// of the 47 (n = 768) to 255 (n = 3072) of the recursion.  Same algebra as obq.py:121-137 -- the
// propagated terms are summed per block instead of per recursion level.
// R (8, 16 or 32 rows per CTA) is chosen so that the grid covers the SMs; the 256 threads are
// R x 8 column groups x KG = 32/R k-groups (split-K inside the CTA, reduced through shared memory).
constexpr int FT = 256;    // threads per CTA
constexpr int FST = 4;     // cp.async stages

template <int R>
struct FusedSmem {
  static constexpr int KG = 32 / R;
  float E[FST][KG][R][36];    // [k-group][row][k], rows padded to 36 floats (16-byte aligned, conflict free)
  float U[FST][KG][32][32];   // [k-group][k][col]
  float red[KG][R][33];       // split-K partial sums
  float Qs[R][33];
  float W0[R][33];            // R form: the original weights of the block
  LeafShared32 leaf;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// RFORM = false: E holds the scaled residuals and U the inverse factor (obq.py:121-137 as written).
// RFORM = true : U is the Cholesky factor R (H_opt = R R^T), Ud the 32x32 diagonal-block inverses
//                (chol_dag.cu), E holds D = W - Q, and the block is formed as
//                W[:, J] + (D[:, :a] R[:a, J]) Ud_J  -- same algebra, no full inverse.
// The kernel sweeps the column range [c0, c1) (c0 a multiple of 32) and forms the propagated term
// from the columns [c0, a) only; what the columns before c0 contribute arrives in Pacc (R form:
// Pacc[:, J] = D[:, :c0] R[:c0, J], accumulated by tensor-core GEMMs between macro blocks).
template <int R, bool RFORM>
__global__ void __launch_bounds__(FT) sweep_fused_kernel(float* __restrict__ Q, float* __restrict__ E, int64_t r, int64_t n,
                                                         const float* __restrict__ U, const float* __restrict__ Ud,
                                                         DevGrid<float> g, int64_t c0, int64_t c1,
                                                         const float* __restrict__ Pacc, float* __restrict__ Esum) {
  typedef FusedSmem<R> SM;
  constexpr int KG = SM::KG;
  constexpr int KSUP = 32 * KG;                       // k covered by one ring stage
  extern __shared__ __align__(16) unsigned char fused_raw[];
  SM& sm = *reinterpret_cast<SM*>(fused_raw);
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int kg = tid / (R * 8), lt = tid % (R * 8);
  const int lrow = lt >> 3, seg = lt & 7;
  const int64_t grow = row0 + lrow;
  const bool aligned = (n % 4 == 0) && ((((uintptr_t)E) & 15) == 0) && ((((uintptr_t)U) & 15) == 0);
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);
  const bool fastq = (g.kind == 0) && fstep.ok;
  float esd[2] = {0.f, 0.f};                            // this row's sums of res^2 and D^2 (leaf threads)

  for (int64_t a = c0; a < c1; a += 32) {
    const int width = (int)((c1 - a) < 32 ? (c1 - a) : 32);
    const bool fullw = aligned && width == 32;
    // operands of the leaf are requested now and parked in registers: their latency hides
    // behind the product loop
    float ud[4], wq[4], pa[4] = {0.f, 0.f, 0.f, 0.f};
    float rdiag = 1.0f;
    if (RFORM && tid < 32) rdiag = (tid < width) ? __ldg(U + (a + tid) * n + (a + tid)) : 1.0f;
    {
      const int i = tid >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = seg * 4 + j;
        if (RFORM) ud[j] = __ldg(Ud + (a / 32) * 1024 + i * 32 + c);
        else ud[j] = (i < width && c < width) ? __ldg(U + (a + i) * n + (a + c)) : ((i == c) ? 1.0f : 0.0f);
        wq[j] = (kg == 0 && grow < r && c < width) ? __ldcg(Q + grow * n + a + c) : 0.0f;
        if (RFORM && Pacc) pa[j] = (kg == 0 && grow < r && c < width) ? __ldcg(Pacc + grow * n + a + c) : 0.0f;
      }
    }
    // ---- (1) Qb = W[:, a:a+32] - E[:, :a] @ U[:a, a:a+32] -----------------------------------
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int nsup = (int)((a - c0 + KSUP - 1) / KSUP);
    auto load_stage = [&](int c) {
      const int st = c % FST;
      const int64_t k0 = c0 + (int64_t)c * KSUP;
      const int64_t ke = k0 + 32 * kg;               // this thread's k-group for the E piece
      float* de = &sm.E[st][kg][lrow][seg * 4];
      if (ke < a) {
        if (aligned && grow < r) cp_async16(de, E + grow * n + ke + seg * 4);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) de[j] = grow < r ? __ldcg(E + grow * n + ke + seg * 4 + j) : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < KG; ++j) {                  // U: KG pieces per thread, k-row tid/8 of group j
        const int64_t ku = k0 + 32 * j;
        if (ku < a) {
          float* du = &sm.U[st][j][tid >> 3][seg * 4];
          const float* su = U + (ku + (tid >> 3)) * n + a + seg * 4;
          if (fullw) cp_async16(du, su);
          else {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) du[q4] = (seg * 4 + q4 < width) ? __ldg(su + q4) : 0.0f;
          }
        }
      }
    };
    for (int c = 0; c < FST - 1; ++c) {
      if (c < nsup) load_stage(c);
      cp_async_commit();
    }
    for (int c = 0; c < nsup; ++c) {
      cp_async_wait<FST - 2>();
      __syncthreads();
      if (c + FST - 1 < nsup) load_stage(c + FST - 1);
      cp_async_commit();
      const int st = c % FST;
      if (c0 + (int64_t)c * KSUP + 32 * kg < a) {
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
          const float e = sm.E[st][kg][lrow][kk];
          const float4 u = *reinterpret_cast<const float4*>(&sm.U[st][kg][kk][seg * 4]);
          acc[0] = __fmaf_rn(e, u.x, acc[0]);
          acc[1] = __fmaf_rn(e, u.y, acc[1]);
          acc[2] = __fmaf_rn(e, u.z, acc[2]);
          acc[3] = __fmaf_rn(e, u.w, acc[3]);
        }
      }
    }
    cp_async_wait<0>();
    {
      const int i = tid >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sm.leaf.U[i][seg * 4 + j] = ud[j];
        sm.leaf.U[i][32 + seg * 4 + j] = 0.0f;
        sm.red[kg][lrow][seg * 4 + j] = acc[j];
      }
    }
    __syncthreads();
    if (kg == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = sm.red[0][lrow][seg * 4 + j];
#pragma unroll
        for (int q = 1; q < KG; ++q) t = __fadd_rn(t, sm.red[q][lrow][seg * 4 + j]);
        if (RFORM) {
          t = __fadd_rn(t, pa[j]);                 // + what the columns before c0 contribute
          sm.Qs[lrow][seg * 4 + j] = t;            // P = D[:, :a] R[:a, J], multiplied by Ud_J below
          sm.W0[lrow][seg * 4 + j] = wq[j];
        } else {
          sm.Qs[lrow][seg * 4 + j] = __fsub_rn(wq[j], t);
        }
      }
    }
    if (RFORM) {
      __syncthreads();
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      if (kg == 0) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float pv = sm.Qs[lrow][c];
          const float4 u = *reinterpret_cast<const float4*>(&sm.leaf.U[c][seg * 4]);
          s4[0] = __fmaf_rn(pv, u.x, s4[0]);
          s4[1] = __fmaf_rn(pv, u.y, s4[1]);
          s4[2] = __fmaf_rn(pv, u.z, s4[2]);
          s4[3] = __fmaf_rn(pv, u.w, s4[3]);
        }
      }
      __syncthreads();
      if (kg == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) sm.Qs[lrow][seg * 4 + j] = __fadd_rn(wq[j], s4[j]);
      }
    }
    if (tid < 32) {
      if (RFORM) {
        sm.leaf.Uy[tid] = rdiag;
      } else {
        const FastDivF f = make_fastdiv(sm.leaf.U[tid][tid]);
        sm.leaf.Uy[tid] = f.y;
        sm.leaf.Uok[tid] = f.ok;
      }
    }
    __syncthreads();
    // ---- (2) leaf: one thread per row ----------------------------------------------------------
    if (tid < R && row0 + tid < r) {
      float q[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) q[k] = sm.Qs[tid][k];
      float* qrow = Q + (row0 + tid) * n + a;
      float* erow = E + (row0 + tid) * n + a;
      const float* w0 = RFORM ? &sm.W0[tid][0] : nullptr;
      float* ep = (RFORM && Esum) ? esd : nullptr;
      leaf_phase32<32>(q, sm.leaf, 0, width, g, fstep, fastq, qrow, erow, w0, nullptr, ep);
      leaf_phase32<24>(q, sm.leaf, 8, width, g, fstep, fastq, qrow, erow, w0, nullptr, ep);
      leaf_phase32<16>(q, sm.leaf, 16, width, g, fstep, fastq, qrow, erow, w0, nullptr, ep);
      leaf_phase32<8>(q, sm.leaf, 24, width, g, fstep, fastq, qrow, erow, w0, nullptr, ep);
    }
    __syncthreads();   // E of this block is visible to the whole CTA before the next block reads it
  }
  if (RFORM && Esum && tid < R && row0 + tid < r) {
    // launches of one sweep follow each other on one stream and every row has one writer: plain
    // read-modify-write, fixed order, deterministic
    float2* p = reinterpret_cast<float2*>(Esum) + (row0 + tid);
    float2 v = make_float2(0.f, 0.f);
    if (c0 > 0) v = *p;
    v.x = __fadd_rn(v.x, esd[0]);
    v.y = __fadd_rn(v.y, esd[1]);
    *p = v;
  }
}

template <int R, bool RFORM>
static int launch_fused(float* q, float* e, int64_t r, int64_t n, const float* u32, const float* ud, const DevGrid<float>& g,
                        cudaStream_t st, int64_t c0 = 0, int64_t c1 = -1, const float* pacc = nullptr,
                        float* esum = nullptr) {
  if (c1 < 0) c1 = n;
  auto kern = sweep_fused_kernel<R, RFORM>;
  SLK_SMEM_ATTR_ONCE(kern, (int)sizeof(FusedSmem<R>));
  kern<<<(unsigned)ceil_div(r, R), FT, sizeof(FusedSmem<R>), st>>>(q, e, r, n, u32, ud, g, c0, c1, pacc, esum);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
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
  static int fused_mode = -1;   // SLK_SWEEP_FUSED=0 forces the multi-launch recursion (A/B testing)
  if (fused_mode < 0) {
    const char* ev = getenv("SLK_SWEEP_FUSED");
    fused_mode = (ev && ev[0] == '0') ? 0 : 1;
  }
  if (!exact_leaf && leaf == 32 && fused_mode) {
    // rows per CTA: as many as still give every SM about two CTAs
    const DevGrid<float> g = make_grid<float>(cb);
    const int64_t want = 2 * (int64_t)sm_count();
    if (r >= 32 * want) return launch_fused<32, false>(q, e, r, n, u32, nullptr, g, (cudaStream_t)stream);
    if (r >= 16 * want) return launch_fused<16, false>(q, e, r, n, u32, nullptr, g, (cudaStream_t)stream);
    return launch_fused<8, false>(q, e, r, n, u32, nullptr, g, (cudaStream_t)stream);
  }
  SweepCtx c;
  c.Q = q; c.E = e; c.r = r; c.n = n; c.u64 = u64; c.u32 = u32;
  c.g = make_grid<float>(cb);
  c.leaf = leaf; c.fanout = fanout; c.exact_leaf = exact_leaf; c.st = (cudaStream_t)stream;
  return sweep_range(c, 0, n);
}

// ---- macro-block kernel (R form) ------------------------------------------------------------------
// Sweeps ONE macro block [c0, c1), c1 - c0 <= MB_COLS, with everything the dependent chain touches
// resident in shared memory: the block's D (written by the leaf, read by the products), and the
// R panel R[c0:a, J] of the NEXT 32-column block, prefetched with cp.async while the current block
// is being swept; the leaf operands of the next block (W, Pacc, Ud, diag R) are prefetched into
// registers the same way.  Per 32 columns the chain is: local product from shared memory ->
// 32x32 multiply by Ud_J -> leaf.  No global-memory latency sits on it.
// Leaf of the macro kernel: FOUR lanes per row, 8 columns each.  For every group of 8 columns
// (A) the owning lane walks its 8 columns alone -- round, residual res = (w - q) * R[i][i], update
// of its remaining columns -- so the dependent chain never crosses a lane; (B) the 8 residuals are
// shuffled to the row's lanes and the lanes that own later columns apply them (64 independent
// FMAs).  obq.py:110-118 in the R form; D = W - q goes to shared memory and to HBM.
// MODE 0: any codebook (grid_value); 1: uniform codebook, exact-reciprocal arithmetic;
// 2: uniform codebook of <= 8 entries through its exact breakpoints X (idx >= k <=> w >= X[k], found
//    by bisection with the reference's op chain at kernel start) and values V: a 3-level compare /
//    select tree, 6 dependent instructions instead of 12 on the column chain, identical results.
template <int MODE>
__device__ __forceinline__ void leaf_rows4(float (&q)[8], const LeafShared32& sh, const DevGrid<float>& g,
                                           const FastDivF& fstep, int width, int part, int lane, bool rowok,
                                           float* __restrict__ qrow, float* __restrict__ drow,
                                           float* __restrict__ dhi, float* __restrict__ dlo,
                                           const float (&w0)[8], float* __restrict__ dsm, float* __restrict__ esm,
                                           const float* __restrict__ XV = nullptr, long long* tr = nullptr) {
  float X[8], V[8];
  float se = 0.0f;
  if (MODE == 2) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { X[k] = XV[k]; V[k] = XV[8 + k]; }
  }
  const float top = (float)(g.size - 1);
  float qv[8], dv8[8];                                   // this lane's eight columns: stored after all four phases
#pragma unroll
  for (int t = 0; t < 8; ++t) qv[t] = dv8[t] = 0.0f;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    if (p * 8 >= width) break;
    float resv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const long long tA = tr ? clock64() : 0;
    if (part == p) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int i = p * 8 + t;
        const float w = q[t];
        float qq;
        if (MODE == 2) {
          const bool b2 = w >= X[4];
          const float t1 = b2 ? X[6] : X[2];
          const float xa = b2 ? X[7] : X[3], xb = b2 ? X[5] : X[1];
          const float va = b2 ? V[7] : V[3], vb = b2 ? V[5] : V[1], vc = b2 ? V[6] : V[2], vd = b2 ? V[4] : V[0];
          const bool b1 = w >= t1;
          const float t0 = b1 ? xa : xb;
          const float c1 = b1 ? va : vb, c0 = b1 ? vc : vd;
          qq = (w >= t0) ? c1 : c0;
        } else if (MODE == 1) {
          // codebook.py:60-64 with the divide by the step through the exact reciprocal scheme and
          // rint through the 1.5*2^23 constant (identical after the clip for every finite input)
          const float tt = fastdiv_core(__fsub_rn(w, g.zero), fstep.d, fstep.y);
          float k = __fsub_rn(__fadd_rn(tt, 12582912.0f), 12582912.0f);
          k = fminf(fmaxf(k, 0.0f), top);
          qq = __fadd_rn(__fmul_rn(k, g.step), g.zero);
        } else {
          qq = grid_value(g, w);
        }
        const float res = __fmul_rn(__fsub_rn(w, qq), sh.Uy[i]);
        const bool in = i < width;
        resv[t] = in ? res : 0.0f;                       // padding columns carry nothing (and add nothing to se)
        if (t < 7) {
          const float4 ua = *reinterpret_cast<const float4*>(&sh.U[i][p * 8]);
          const float4 ub = *reinterpret_cast<const float4*>(&sh.U[i][p * 8 + 4]);
          const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
          for (int c = t + 1; c < 8; ++c) q[c] = __fmaf_rn(-res, u[c], q[c]);
        }
        qv[t] = qq;
        dv8[t] = in ? __fsub_rn(w0[t], qq) : 0.0f;
      }
      if (tr && p == 0) tr[8] = clock64() - tA;          // development trace: the bare 8-column chain of phase 0
    }
    if (tr && p == 0) tr[9] = clock64() - tA;            // ... + its stores
    __syncwarp();
    const long long tB = tr ? clock64() : 0;
    const int owner_lane = (lane & ~3) | p;
#pragma unroll
    for (int t = 0; t < 8; ++t) resv[t] = __shfl_sync(0xffffffffu, resv[t], owner_lane);
    if (tr && p == 0) tr[10] = clock64() - tB;           // the eight shuffles
    // error identity (slk_gptq_sweep_r_err_f32): sum of res^2, off the owner's chain -- every lane of
    // the row holds the 8 residuals now and keeps the same running sum
#pragma unroll
    for (int t = 0; t < 8; ++t) se = __fmaf_rn(resv[t], resv[t], se);
    if (part > p) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 ua = *reinterpret_cast<const float4*>(&sh.U[p * 8 + t][part * 8]);
        const float4 ub = *reinterpret_cast<const float4*>(&sh.U[p * 8 + t][part * 8 + 4]);
        const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = __fmaf_rn(-resv[t], u[c], q[c]);
      }
    }
    if (tr) { tr[6] += tB - tA; tr[7] += clock64() - tB; }   // development trace: owner walk / shuffle + update
  }
  // Everything a lane produced is stored AFTER the row's dependent chain, by all four lanes at once, as
  // 128-bit accesses with one predicate per quad (the macro path runs with n % 4 == 0 and 256-column macro
  // blocks, so a quad of columns is inside the block or outside it as a whole): no store, predicate or
  // branch sits between two links of the chain, and the stores of the four lanes overlap.
  if (part * 8 < width) {
    float4* dsm4 = reinterpret_cast<float4*>(dsm + part * 8);
    dsm4[0] = make_float4(dv8[0], dv8[1], dv8[2], dv8[3]);
    dsm4[1] = make_float4(dv8[4], dv8[5], dv8[6], dv8[7]);
    if (rowok) {
#pragma unroll
      for (int h4 = 0; h4 < 2; ++h4) {
        if (part * 8 + h4 * 4 < width) {
          const int o = part * 8 + h4 * 4;
          *reinterpret_cast<float4*>(qrow + o) = make_float4(qv[h4 * 4], qv[h4 * 4 + 1], qv[h4 * 4 + 2], qv[h4 * 4 + 3]);
          *reinterpret_cast<float4*>(drow + o) = make_float4(dv8[h4 * 4], dv8[h4 * 4 + 1], dv8[h4 * 4 + 2], dv8[h4 * 4 + 3]);
          if (dhi) {                                     // TF32 parts for the macro-block GEMM
            float hh[4], ll[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_tf32(dv8[h4 * 4 + j], hh[j], ll[j]);
            *reinterpret_cast<float4*>(dhi + o) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<float4*>(dlo + o) = make_float4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
      }
    }
  }
  esm[0] = __fadd_rn(esm[0], se);                        // this lane's own shared-memory slot
}

// The same leaf as ONE branch-free instruction stream (round 2).  leaf_rows4 above alternates two divergent
// regions per phase -- the owner lane's 8-column walk, then 8 shuffles and the other lanes' 64 FMAs -- and a warp
// issues in order, so the owner's chain stalls (50 cycles per column, tools/lat_chain.cu) hide nothing and the
// update sits between two phases: 1.3-1.5 k cycles per phase measured, 400 of them the chain.  Here every lane
// executes every column step: it rounds ITS column t (only the owner's result is used: the others are masked to
// zero, and a multiply-add with a zero factor leaves a value unchanged bit for bit), the residual is broadcast
// with one shuffle, and the lanes that own later columns apply the residual of the PREVIOUS step -- independent
// FMAs the compiler schedules into the chain's stall slots.  Same operations on the same operands in the same
// order for every weight as leaf_rows4: identical results.
template <int MODE>
__device__ __forceinline__ void leaf_rows4_pipelined(float (&q)[8], const LeafShared32& sh, const DevGrid<float>& g,
                                                     const FastDivF& fstep, int width, int part, int lane, bool rowok,
                                                     float* __restrict__ qrow, float* __restrict__ drow,
                                                     float* __restrict__ dhi, float* __restrict__ dlo,
                                                     const float (&w0)[8], float* __restrict__ dsm,
                                                     float* __restrict__ esm, const float* __restrict__ XV = nullptr,
                                                     long long* tr = nullptr) {
  float X[8], V[8];
  float se = 0.0f;
  if (MODE == 2) {
#pragma unroll
    for (int k = 0; k < 8; ++k) { X[k] = XV[k]; V[k] = XV[8 + k]; }
  }
  const float top = (float)(g.size - 1);
  float qv[8], dv8[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) qv[t] = dv8[t] = 0.0f;
  const long long tA = tr ? clock64() : 0;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    if (p * 8 >= width) break;
    const bool own = part == p;
    const bool later = part > p;
    const int owner_lane = (lane & ~3) | p;
    const float* urow_own = &sh.U[p * 8][p * 8];          // + t * 64: row p*8+t, the phase's own 8 columns
    const float* urow_lat = &sh.U[p * 8][part * 8];       // + t * 64: row p*8+t, this lane's 8 columns
    float rq = 0.0f;                                     // the previous step's residual, applied one step late
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int i = p * 8 + t;
      const float w = q[t];
      float qq;
      if (MODE == 2) {
        const bool b2 = w >= X[4];
        const float t1 = b2 ? X[6] : X[2];
        const float xa = b2 ? X[7] : X[3], xb = b2 ? X[5] : X[1];
        const float va = b2 ? V[7] : V[3], vb = b2 ? V[5] : V[1], vc = b2 ? V[6] : V[2], vd = b2 ? V[4] : V[0];
        const bool b1 = w >= t1;
        const float t0 = b1 ? xa : xb;
        const float c1 = b1 ? va : vb, c0 = b1 ? vc : vd;
        qq = (w >= t0) ? c1 : c0;
      } else if (MODE == 1) {
        const float tt = fastdiv_core(__fsub_rn(w, g.zero), fstep.d, fstep.y);
        float k = __fsub_rn(__fadd_rn(tt, 12582912.0f), 12582912.0f);
        k = fminf(fmaxf(k, 0.0f), top);
        qq = __fadd_rn(__fmul_rn(k, g.step), g.zero);
      } else {
        qq = grid_value(g, w);
      }
      const float res = __fmul_rn(__fsub_rn(w, qq), sh.Uy[i]);
      const bool live = own && i < width;                // padding columns and the other lanes carry nothing
      const float res_m = live ? res : 0.0f;
      if (t < 7) {
        const float4 ua = *reinterpret_cast<const float4*>(urow_own + t * 64);
        const float4 ub = *reinterpret_cast<const float4*>(urow_own + t * 64 + 4);
        const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
        for (int c = t + 1; c < 8; ++c) q[c] = __fmaf_rn(-res_m, u[c], q[c]);
      }
      qv[t] = own ? qq : qv[t];
      dv8[t] = own ? (i < width ? __fsub_rn(w0[t], qq) : 0.0f) : dv8[t];
      // the later lanes apply the PREVIOUS step's residual (already broadcast) ...
      if (t > 0) {
        const float rm = later ? rq : 0.0f;
        const float4 ua = *reinterpret_cast<const float4*>(urow_lat + (t - 1) * 64);
        const float4 ub = *reinterpret_cast<const float4*>(urow_lat + (t - 1) * 64 + 4);
        const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) q[c] = __fmaf_rn(-rm, u[c], q[c]);
      }
      // ... while this step's residual travels
      rq = __shfl_sync(0xffffffffu, res_m, owner_lane);
      se = __fmaf_rn(rq, rq, se);
    }
    {                                                    // the phase's last residual
      const float rm = later ? rq : 0.0f;
      const float4 ua = *reinterpret_cast<const float4*>(urow_lat + 7 * 64);
      const float4 ub = *reinterpret_cast<const float4*>(urow_lat + 7 * 64 + 4);
      const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) q[c] = __fmaf_rn(-rm, u[c], q[c]);
    }
  }
  if (tr) { tr[6] = clock64() - tA; tr[7] = 0; }
  if (part * 8 < width) {
    float4* dsm4 = reinterpret_cast<float4*>(dsm + part * 8);
    dsm4[0] = make_float4(dv8[0], dv8[1], dv8[2], dv8[3]);
    dsm4[1] = make_float4(dv8[4], dv8[5], dv8[6], dv8[7]);
    if (rowok) {
#pragma unroll
      for (int h4 = 0; h4 < 2; ++h4) {
        if (part * 8 + h4 * 4 < width) {
          const int o = part * 8 + h4 * 4;
          *reinterpret_cast<float4*>(qrow + o) = make_float4(qv[h4 * 4], qv[h4 * 4 + 1], qv[h4 * 4 + 2], qv[h4 * 4 + 3]);
          *reinterpret_cast<float4*>(drow + o) = make_float4(dv8[h4 * 4], dv8[h4 * 4 + 1], dv8[h4 * 4 + 2], dv8[h4 * 4 + 3]);
          if (dhi) {
            float hh[4], ll[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_tf32(dv8[h4 * 4 + j], hh[j], ll[j]);
            *reinterpret_cast<float4*>(dhi + o) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<float4*>(dlo + o) = make_float4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
      }
    }
  }
  esm[0] = __fadd_rn(esm[0], se);
}

constexpr int MB_COLS = 256;
constexpr int MB_PITCH = MB_COLS + 4;
__device__ long long* g_sweep_trace = nullptr;   // development aid: per-block phase clocks of CTA 0

// R panels of the macro kernel.  Default: two full buffers (the look-ahead of block a reads the
// panel of block a+32 while the tail product of block a reads the last 32 rows of its own panel).
// COMPACT (default since round 2; SLK_SWEEP_COMPACT=0 selects the double buffer): only the last 32 rows of
// the current panel are ever read again, so one full buffer plus two 32-row tail buffers do -- 20 KB less
// per CTA (R = 8/16: three CTAs per SM instead of two, R = 32: two instead of one); the prefetch then has
// to be issued after the barrier that ends the previous look-ahead.  Measured on B200: a single layer's
// sweep is ~10 % slower (the prefetch starts later), a layer set's pass 1-2 % faster, [8192, 28672] 58 -> 55 ms.
template <bool COMPACT>
struct PanelBufs;
template <>
struct PanelBufs<false> {
  float Rs[2][MB_COLS - 32][32];            // R[c0 + k][a + col], k < a - c0
};
template <>
struct PanelBufs<true> {
  float Rs[1][MB_COLS - 32][32];            // rows k < a - c0 of the NEXT block's panel (look-ahead)
  float Rt[2][32][32];                      // the last 32 rows of each panel (tail product)
};

template <int R, bool COMPACT = false>
struct MacroSmem {
  static constexpr int KG = 32 / R;
  static constexpr int NLEAF = 4 * R;                       // threads of the leaf
  static constexpr int NSLOT = 8 * R;                       // (row, 4-column) output slots of a block
  static constexpr int HELP = FT - NLEAF;                   // threads free while the leaf runs
  // ONE k group for the look-ahead whatever the tile height: the order in which a row's products are
  // summed must not depend on R, so that a row gets the same bits however the rows are cut into CTAs
  // (row slices, row-sharded runs and the layer-set driver's taller tiles all agree with the full run)
  static constexpr int KGH = 1;
  static constexpr int SPT = HELP >= NSLOT ? 1 : NSLOT / HELP;   // slots per helper thread
  float Dm[R][MB_PITCH];                    // D = W - Q of this macro block
  PanelBufs<COMPACT> pb;
  float red[4][R][33];                      // tail product: four chains of 8 k-steps, summed in a fixed order
  float red2[KGH][R][33];                   // look-ahead part of the next block's product
  float XV[16];                             // codebook breakpoints X[0..7] (X[0] unused) and values V[0..7]
  float Qs[R][33];
  float W0[R][33];
  float esd[NLEAF];                         // per leaf lane: running sum of res^2 of its row
  LeafShared32 leaf;
};

template <int R, bool COMPACT, bool PIPE>
__global__ void __launch_bounds__(FT, (COMPACT && !PIPE) ? (R == 32 ? 2 : 3) : 2)
sweep_macro_kernel(float* __restrict__ Q, float* __restrict__ D, int64_t r, int64_t n,
                                                         const float* __restrict__ Rf, const float* __restrict__ Ud,
                                                         DevGrid<float> g, int64_t c0, int64_t c1,
                                                         const float* __restrict__ Pacc, float* __restrict__ Dhi,
                                                         float* __restrict__ Dlo, GridBreaks brk,
                                                         float* __restrict__ Esum) {
  typedef MacroSmem<R, COMPACT> SM;
  constexpr int KG = SM::KG;
  extern __shared__ __align__(16) unsigned char macro_raw[];
  SM& sm = *reinterpret_cast<SM*>(macro_raw);
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int kg = tid / (R * 8), lt = tid % (R * 8);
  const int lrow = lt >> 3, seg = lt & 7;
  const int64_t grow = row0 + lrow;
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);
  const bool fastq = (g.kind == 0) && fstep.ok;

  // leaf operands of a block, fetched one block ahead
  struct Pre { float ud[4], wq[4], pa[4], rdiag; };
  auto fetch = [&](int64_t a, Pre& p) {
    const int width = (int)((c1 - a) < 32 ? (c1 - a) : 32);
    const int i = tid >> 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = seg * 4 + j;
      p.ud[j] = __ldg(Ud + (a / 32) * 1024 + i * 32 + c);
      const bool in = (kg == 0 && grow < r && c < width);
      p.wq[j] = in ? __ldcg(Q + grow * n + a + c) : 0.0f;
      p.pa[j] = (in && Pacc) ? __ldcg(Pacc + grow * n + a + c) : 0.0f;
    }
    p.rdiag = (tid < width) ? __ldg(Rf + (a + tid) * n + (a + tid)) : 1.0f;
  };
  // R panel of block a: rows c0 .. a-1, columns a .. a+31 (zero beyond c1)
  auto prefetch_panel = [&](int64_t a, int buf) {
    // issued by the helper threads only (the ones that read it first, in the look-ahead): the leaf threads
    // then never wait for the panel -- in the compact layout it is requested one barrier later and used to
    // cost the chain several hundred cycles per block at the CTA-wide wait
    if (tid < SM::NLEAF) return;
    const int rows = (int)(a - c0);
    const int width = (int)((c1 - a) < 32 ? (c1 - a) : 32);
    for (int t = tid - SM::NLEAF; t < rows * 8; t += SM::HELP) {
      const int k = t >> 3, piece = t & 7;
      float* dst;
      if constexpr (COMPACT) dst = (k < rows - 32) ? &sm.pb.Rs[0][k][piece * 4] : &sm.pb.Rt[buf][k - (rows - 32)][piece * 4];
      else dst = &sm.pb.Rs[buf][k][piece * 4];
      const float* src = Rf + (c0 + k) * n + a + piece * 4;
      if (piece * 4 + 4 <= width) cp_async16(dst, src);
      else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = (piece * 4 + j < width) ? __ldg(src + j) : 0.0f;
      }
    }
    cp_async_commit();
  };

  // Schedule per 32-column block J (panel = R[c0:a, J], landed one block earlier):
  //   (1) tail of the product: the 32 columns of D the previous leaf produced, all threads
  //   (2) + the look-ahead part computed during the previous leaf (red2) + Pacc, times Ud_J
  //   (3) leaf on 4R threads  ||  the other threads form the next block's product over all
  //       columns of D that already exist (look-ahead), so only 32 k-steps remain on the chain
  for (int t = tid; t < SM::KGH * R * 33; t += FT) (&sm.red2[0][0][0])[t] = 0.0f;
  const bool tree = fastq && g.size <= 8;
  if (tree && tid >= FT - 8) {
    // breakpoints (exact; found on the host, make_breaks) and values (codebook.py:63-64) of the codebook
    const int k = tid - (FT - 8);
    sm.XV[k] = brk.X[k];
    const int kv = k < g.size ? k : g.size - 1;
    sm.XV[8 + k] = __fadd_rn(__fmul_rn((float)kv, g.step), g.zero);
  }
  Pre cur, nxt;
  fetch(c0, cur);
  int buf = 0;
  if (tid < SM::NLEAF) sm.esd[tid] = 0.0f;             // only ever touched by its own lane
  for (int64_t a = c0; a < c1; a += 32, buf ^= 1) {
    const int width = (int)((c1 - a) < 32 ? (c1 - a) : 32);
    const int ka = (int)(a - c0);                      // columns of D that exist when this block starts
    const bool has_next = a + 32 < c1;
    if (has_next) fetch(a + 32, nxt);
    long long* tr = (g_sweep_trace && blockIdx.x == 0 && tid == 0) ? g_sweep_trace + ((a - c0) / 32) * 16 : nullptr;
    if (tr) tr[0] = clock64();
    __syncthreads();                                   // previous leaf's D, the look-ahead sums and the panel are visible
    if (tr) tr[1] = clock64();
    // ---- (1) tail: D[:, a-32:a] R[a-32:a, J]: four chains of 8 k-steps each (chain c = k 8c..8c+7), dealt to
    // the KG thread groups, summed ((c0 + c1) + c2) + c3 -- the same association for every tile height R
    constexpr int CPG = 4 / KG;                        // chains per thread group
    float acc[CPG][4];
#pragma unroll
    for (int c = 0; c < CPG; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    if (ka > 0) {
#pragma unroll
      for (int c = 0; c < CPG; ++c) {
        const int koff = (kg * CPG + c) * 8;           // within the last 32 columns of D
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const float e = sm.Dm[lrow][ka - 32 + koff + kk];
          const float* rp;
          if constexpr (COMPACT) rp = &sm.pb.Rt[buf][koff + kk][seg * 4];
          else rp = &sm.pb.Rs[buf][ka - 32 + koff + kk][seg * 4];
          const float4 u = *reinterpret_cast<const float4*>(rp);
          acc[c][0] = __fmaf_rn(e, u.x, acc[c][0]);
          acc[c][1] = __fmaf_rn(e, u.y, acc[c][1]);
          acc[c][2] = __fmaf_rn(e, u.z, acc[c][2]);
          acc[c][3] = __fmaf_rn(e, u.w, acc[c][3]);
        }
      }
    }
    {
      const int i = tid >> 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sm.leaf.U[i][seg * 4 + j] = cur.ud[j];
        sm.leaf.U[i][32 + seg * 4 + j] = 0.0f;
      }
      if (tid < 32) sm.leaf.Uy[tid] = cur.rdiag;
    }
    if constexpr (KG == 1) {
      // one thread group holds all four chains: same association ((c0 + c1) + c2) + c3, in registers,
      // and one CTA barrier less per block
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[0][j];
#pragma unroll
        for (int q = 1; q < 4; ++q) t = __fadd_rn(t, acc[q][j]);
#pragma unroll
        for (int q = 0; q < SM::KGH; ++q) t = __fadd_rn(t, sm.red2[q][lrow][seg * 4 + j]);
        sm.Qs[lrow][seg * 4 + j] = __fadd_rn(t, cur.pa[j]);
        sm.W0[lrow][seg * 4 + j] = cur.wq[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < CPG; ++c) sm.red[kg * CPG + c][lrow][seg * 4 + j] = acc[c][j];
      __syncthreads();
      if (kg == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = sm.red[0][lrow][seg * 4 + j];
#pragma unroll
          for (int q = 1; q < 4; ++q) t = __fadd_rn(t, sm.red[q][lrow][seg * 4 + j]);
#pragma unroll
          for (int q = 0; q < SM::KGH; ++q) t = __fadd_rn(t, sm.red2[q][lrow][seg * 4 + j]);
          sm.Qs[lrow][seg * 4 + j] = __fadd_rn(t, cur.pa[j]);
          sm.W0[lrow][seg * 4 + j] = cur.wq[j];
        }
      }
    }
    __syncthreads();
    // ---- (2) block = W + P Ud_J --------------------------------------------------------------------
    if (tr) tr[2] = clock64();
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    if (kg == 0) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float pv = sm.Qs[lrow][c];
        const float4 u = *reinterpret_cast<const float4*>(&sm.leaf.U[c][seg * 4]);
        s4[0] = __fmaf_rn(pv, u.x, s4[0]);
        s4[1] = __fmaf_rn(pv, u.y, s4[1]);
        s4[2] = __fmaf_rn(pv, u.z, s4[2]);
        s4[3] = __fmaf_rn(pv, u.w, s4[3]);
      }
    }
    __syncthreads();
    if (kg == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) sm.Qs[lrow][seg * 4 + j] = __fadd_rn(cur.wq[j], s4[j]);
    }
    __syncthreads();
    if (tr) tr[3] = clock64();
    if (tid < SM::NLEAF) {
      // ---- (3a) leaf: four lanes per row (leaf_rows4) -----------------------------------------------
      const int lr = tid >> 2, part = tid & 3, lane = tid & 31;
      const bool rowok = row0 + lr < r;
      float q[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) q[c] = sm.Qs[lr][part * 8 + c];
      float* qrow = Q + (row0 + lr) * n + a;
      float* drow = D + (row0 + lr) * n + a;
      float w0[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) w0[c] = sm.W0[lr][part * 8 + c];
      float* dsm = &sm.Dm[lr][ka];
      float* dhi = Dhi ? Dhi + (row0 + lr) * n + a : nullptr;
      float* dlo = Dhi ? Dlo + (row0 + lr) * n + a : nullptr;
      if (tr) { tr[6] = 0; tr[7] = 0; tr[11] = clock64(); }
      if constexpr (PIPE) {
        if (tree) leaf_rows4_pipelined<2>(q, sm.leaf, g, fstep, width, part, lane, rowok, qrow, drow, dhi, dlo, w0, dsm, &sm.esd[tid], sm.XV, tr);
        else if (fastq) leaf_rows4_pipelined<1>(q, sm.leaf, g, fstep, width, part, lane, rowok, qrow, drow, dhi, dlo, w0, dsm, &sm.esd[tid]);
        else leaf_rows4_pipelined<0>(q, sm.leaf, g, fstep, width, part, lane, rowok, qrow, drow, dhi, dlo, w0, dsm, &sm.esd[tid]);
      } else {
        if (tree) leaf_rows4<2>(q, sm.leaf, g, fstep, width, part, lane, rowok, qrow, drow, dhi, dlo, w0, dsm, &sm.esd[tid], sm.XV, tr);
        else if (fastq) leaf_rows4<1>(q, sm.leaf, g, fstep, width, part, lane, rowok, qrow, drow, dhi, dlo, w0, dsm, &sm.esd[tid]);
        else leaf_rows4<0>(q, sm.leaf, g, fstep, width, part, lane, rowok, qrow, drow, dhi, dlo, w0, dsm, &sm.esd[tid]);
      }
    } else if (has_next) {
      // ---- (3b) look-ahead: D[:, c0:a] R[c0:a, J+1] on the threads the leaf does not use -----------
      // The R panels are requested TWO blocks ahead, by the helper threads, at the end of this phase (below):
      // neither the request loop nor the wait sits on the chain.  Here: the panel of block a + 32, requested
      // one iteration ago, must have landed (this thread's pieces, then every helper's).
      cp_async_wait<0>();
      asm volatile("bar.sync 1, %0;" ::"n"(SM::HELP) : "memory");   // a barrier of the helper threads only
      const int h = tid - SM::NLEAF;
      const int gh = SM::SPT == 1 ? h / SM::NSLOT : 0;
      if (gh < SM::KGH) {
#pragma unroll
        for (int sp = 0; sp < SM::SPT; ++sp) {
          const int slot = (SM::SPT == 1 ? h % SM::NSLOT : h + sp * SM::HELP);
          const int hr = slot >> 3, hs = slot & 7;
          float la[4] = {0.f, 0.f, 0.f, 0.f};
          for (int kb = gh * 32; kb < ka; kb += 32 * SM::KGH) {
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) {
              const float e = sm.Dm[hr][kb + kk];
              const float4 u = *reinterpret_cast<const float4*>(&sm.pb.Rs[COMPACT ? 0 : (buf ^ 1)][kb + kk][hs * 4]);
              la[0] = __fmaf_rn(e, u.x, la[0]);
              la[1] = __fmaf_rn(e, u.y, la[1]);
              la[2] = __fmaf_rn(e, u.z, la[2]);
              la[3] = __fmaf_rn(e, u.w, la[3]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) sm.red2[gh][hr][hs * 4 + j] = la[j];
        }
      }
      // panel requests for later blocks.  Block a + 64's panel goes where block a's was: its rows k < a + 32 - c0
      // into the look-ahead buffer this phase has just finished reading (hence the helpers' barrier), its last
      // 32 rows into the tail buffer the tail product of this iteration read long ago.  The first iteration also
      // requests block c0 + 32's panel (32 rows, tail buffer only) and waits for it, because the next iteration's
      // tail product -- all threads, before any helper wait -- reads it.
      asm volatile("bar.sync 1, %0;" ::"n"(SM::HELP) : "memory");
      if (a == c0) prefetch_panel(a + 32, buf ^ 1);
      if (a + 64 < c1) prefetch_panel(a + 64, buf);
      if (a == c0) { if (a + 64 < c1) cp_async_wait<1>(); else cp_async_wait<0>(); }
    }
    if (tr) { tr[4] = clock64(); tr[11] = tr[4] - tr[11]; }   // tr[11]: the leaf call alone (thread 0)
    cur = nxt;
    if (tr) tr[5] = clock64();
    // the barrier at the top of the next iteration orders Dm / Qs / leaf / red2 reuse
  }
  if (Esum) {
    // sum of res^2: lane 0 of each row's four leaf lanes holds it.  sum of D^2: the macro block's D is
    // still in shared memory -- FT / R threads per row, fixed-order shuffle tree.  One writer per row
    // and value; the macro-block launches of a sweep follow each other on one stream, so a plain
    // read-modify-write is deterministic.
    __syncthreads();
    if (tid < SM::NLEAF && (tid & 3) == 0) {
      const int64_t row = row0 + (tid >> 2);
      if (row < r) {
        float* p = Esum + 2 * row;
        const float prev = c0 > 0 ? *p : 0.0f;
        *p = __fadd_rn(prev, sm.esd[tid]);
      }
    }
    constexpr int TPR = FT / R;                        // 32, 16 or 8 threads per row
    const int drow_ = tid / TPR, sub = tid % TPR;
    const int cols = (int)(c1 - c0);
    float sd = 0.0f;
    for (int c = sub; c < cols; c += TPR) {
      const float dv = sm.Dm[drow_][c];
      sd = __fmaf_rn(dv, dv, sd);
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) sd = __fadd_rn(sd, __shfl_xor_sync(0xffffffffu, sd, o));
    if (sub == 0 && row0 + drow_ < r) {
      float* p = Esum + 2 * (row0 + drow_) + 1;
      const float prev = c0 > 0 ? *p : 0.0f;
      *p = __fadd_rn(prev, sd);
    }
  }
}

extern "C" int slk_set_option(const char* name, int64_t value) {
  SLK_REQUIRE(name, "NULL option name");
  if (strcmp(name, "sweep_ctas") == 0) { g_opt_sweep_ctas = value > 0 ? value : 0; return SLK_OK; }
  if (strcmp(name, "fullh_topk") == 0) {
    SLK_REQUIRE(value == 0 || value == 4 || value == 8 || value == 16, "fullh_topk must be 0, 4, 8 or 16");
    g_opt_fullh_topk = value;
    return SLK_OK;
  }
  if (strcmp(name, "fullh_ctas") == 0) { g_opt_fullh_ctas = value == 1 ? 1 : 2; return SLK_OK; }
  if (strcmp(name, "fullh_compact") == 0) { g_opt_fullh_compact = value != 0; return SLK_OK; }
  if (strcmp(name, "fullh_bf16") == 0) { g_opt_fullh_bf16 = value != 0; return SLK_OK; }
  if (strcmp(name, "fullh_bn") == 0) {
    SLK_REQUIRE(value == 128 || value == 256, "fullh_bn must be 128 or 256");
    g_opt_fullh_bn = value;
    return SLK_OK;
  }
  set_error("unknown option %s", name);
  return SLK_ERR_ARG;
}

extern "C" int slk_debug_sweep_trace(void* buf) {
  long long* p = (long long*)buf;
  SLK_CUDA(cudaMemcpyToSymbol(g_sweep_trace, &p, sizeof(p)));
  return SLK_OK;
}

// ---- CUDA-core push for narrow layers -------------------------------------------------------------
// P[:, 0:N] += D[:, 0:K] Rm[0:K, 0:N]  (the macro-block GEMM of the sweep) as a light fp32 FMA kernel:
// 64 x 64 tiles, 256 threads, 16 accumulators, ~17 KB of shared memory.  A tensor-core tile of
// tc_gemm.cu owns a whole SM (192 KB of shared memory, 36 K registers), so inside a layer SET it can
// not run beside a resident Cholesky CTA of another layer and the sweep stalls at its first push; for
// narrow layers (n <= 1024: K = 256, N <= 768, a few hundred MFLOP) this kernel costs microseconds,
// is exact fp32 FMA arithmetic in k order (deterministic) and shares an SM with anything.
constexpr int PS_BM = 64, PS_BN = 64, PS_BK = 16;
__global__ void __launch_bounds__(256, 4) push_simt_kernel(const float* __restrict__ D, int64_t ldd,
                                                           const float* __restrict__ Rm, int64_t ldr,
                                                           float* __restrict__ P, int64_t ldp, int64_t M, int64_t N,
                                                           int64_t K) {
  __shared__ __align__(16) float As[PS_BK][PS_BM + 4];
  __shared__ __align__(16) float Bs[PS_BK][PS_BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * PS_BM, n0 = (int64_t)blockIdx.x * PS_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int ar = tid >> 2, ak = (tid & 3) * 4;          // A: row ar, k quad ak
  const int bk = tid >> 4, bn = (tid & 15) * 4;         // B: k row bk, column quad bn
  for (int64_t k0 = 0; k0 < K; k0 += PS_BK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + ar < M) {
      const float* src = D + (m0 + ar) * ldd + k0 + ak;
      if (k0 + ak + 3 < K) a = *reinterpret_cast<const float4*>(src);
      else {
        if (k0 + ak < K) a.x = src[0];
        if (k0 + ak + 1 < K) a.y = src[1];
        if (k0 + ak + 2 < K) a.z = src[2];
      }
    }
    if (k0 + bk < K) {
      const float* src = Rm + (k0 + bk) * ldr + n0 + bn;
      if (n0 + bn + 3 < N) b = *reinterpret_cast<const float4*>(src);
      else {
        if (n0 + bn < N) b.x = src[0];
        if (n0 + bn + 1 < N) b.y = src[1];
        if (n0 + bn + 2 < N) b.z = src[2];
      }
    }
    __syncthreads();                                    // the previous tile has been consumed
    As[ak][ar] = a.x; As[ak + 1][ar] = a.y; As[ak + 2][ar] = a.z; As[ak + 3][ar] = a.w;
    *reinterpret_cast<float4*>(&Bs[bk][bn]) = b;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PS_BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ax[4] = {av.x, av.y, av.z, av.w}, bx[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(ax[i], bx[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= M) continue;
    float* dst = P + row * ldp + n0 + tx * 4;
    if (n0 + tx * 4 + 3 < N) {
      float4 c = *reinterpret_cast<float4*>(dst);
      c.x = __fadd_rn(c.x, acc[i][0]); c.y = __fadd_rn(c.y, acc[i][1]);
      c.z = __fadd_rn(c.z, acc[i][2]); c.w = __fadd_rn(c.w, acc[i][3]);
      *reinterpret_cast<float4*>(dst) = c;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n0 + tx * 4 + j < N) dst[j] = __fadd_rn(dst[j], acc[i][j]);
    }
  }
}

template <int R, bool COMPACT, bool PIPE>
static int launch_macro_v(float* q, float* d, int64_t r, int64_t n, const float* r32, const float* ud, const DevGrid<float>& g,
                          cudaStream_t st, int64_t c0, int64_t c1, const float* pacc, float* dhi, float* dlo,
                          const GridBreaks& xv, float* esum) {
  auto kern = sweep_macro_kernel<R, COMPACT, PIPE>;
  SLK_SMEM_ATTR_ONCE(kern, (int)sizeof(MacroSmem<R, COMPACT>));
  kern<<<(unsigned)ceil_div(r, R), FT, sizeof(MacroSmem<R, COMPACT>), st>>>(q, d, r, n, r32, ud, g, c0, c1, pacc, dhi, dlo,
                                                                           xv, esum);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

template <int R>
static int launch_macro(float* q, float* d, int64_t r, int64_t n, const float* r32, const float* ud, const DevGrid<float>& g,
                        cudaStream_t st, int64_t c0, int64_t c1, const float* pacc, float* dhi, float* dlo,
                        const GridBreaks& xv, float* esum) {
  static int compact = -1;   // SLK_SWEEP_COMPACT=0/1: double-buffered / single R-panel buffer (PanelBufs)
  if (compact < 0) {
    const char* ev = getenv("SLK_SWEEP_COMPACT");
    compact = (ev && ev[0] == '0') ? 0 : 1;   // default on: one more CTA per SM (measured)
  }
  static int pipe = -1;      // SLK_LEAF_PIPE=0: the phase-by-phase leaf (leaf_rows4) instead of the branch-free pipelined one
  if (pipe < 0) {
    const char* ev = getenv("SLK_LEAF_PIPE");
    pipe = (ev && ev[0] == '0') ? 0 : 1;
  }
  if (pipe) {
    if (compact) return launch_macro_v<R, true, true>(q, d, r, n, r32, ud, g, st, c0, c1, pacc, dhi, dlo, xv, esum);
    return launch_macro_v<R, false, true>(q, d, r, n, r32, ud, g, st, c0, c1, pacc, dhi, dlo, xv, esum);
  }
  if (compact) return launch_macro_v<R, true, false>(q, d, r, n, r32, ud, g, st, c0, c1, pacc, dhi, dlo, xv, esum);
  return launch_macro_v<R, false, false>(q, d, r, n, r32, ud, g, st, c0, c1, pacc, dhi, dlo, xv, esum);
}

// R form of the sweep: r32 = Cholesky factor R (upper, H_opt = R R^T), rt32 = its transpose, ud32 =
// [ceil(n/32), 32, 32] inverses of its diagonal blocks (slk_chol_factor_f32).  d receives W - Q.
// Lazy batching (obq.py:121-137's idea, two levels): the columns are cut into macro blocks of
// SWEEP_MB; inside one the fused kernel propagates locally, between them ONE tensor-core GEMM
// (tcgen05, fp32-faithful 3xTF32) pushes the finished block to every later column:
//     Pacc[:, e:] += D[:, s:e] R[s:e, e:]
static constexpr int64_t SWEEP_MB = MB_COLS;
static constexpr int64_t SWEEP_SUPER = 2048;

static bool sweep_macro_ok(int64_t r, int64_t n, const float* d, const float* rt_hi, const float* rt_lo) {
  static int off = -1;
  if (off < 0) {
    const char* ev = getenv("SLK_SWEEP_MACRO");
    off = (ev && ev[0] == '0') ? 1 : 0;
  }
  return !off && rt_hi != nullptr && rt_lo != nullptr && n >= 2 * SWEEP_MB && tc_gemm_usable(rt_hi, n, rt_lo, n);
}

extern "C" size_t slk_gptq_sweep_r_ws_bytes(int64_t r, int64_t n) {
  if (n < 2 * SWEEP_MB || n % 4 != 0) return 256;
  return (size_t)3 * r * n * sizeof(float) + 1024;     // Pacc, D_hi, D_lo
}

extern "C" int slk_gptq_sweep_r_f32(float* q, float* d, int64_t r, int64_t n, const float* r32, const float* rt_hi,
                                    const float* rt_lo, const float* ud32, const slk_codebook* cb, void* ws,
                                    size_t ws_bytes, void* stream) {
  return slk_gptq_sweep_r_err_f32(q, d, r, n, r32, rt_hi, rt_lo, ud32, cb, ws, ws_bytes, nullptr, stream);
}

// The same sweep, also returning per row the sums err_sums[row] = (sum_i res_i^2, sum_i D_i^2).  With
// res_i = (w'_i - q_i) R_ii = obq.py:114's scaled residual E_i one has W - Q = E U and U H_opt U^T = I,
// hence  (W-Q) H_opt (W-Q)^T = sum_i E_i^2  exactly, and the layer error of obq.py:89-95 under the
// UNDAMPED H the factor was formed from is  sum_i E_i^2 - damp_abs * sum_i D_i^2
// (slk_sweep_error_f32): the 2 r n^2 flop product of K6 is not needed after a sweep.
extern "C" int slk_gptq_sweep_r_err_f32(float* q, float* d, int64_t r, int64_t n, const float* r32, const float* rt_hi,
                                        const float* rt_lo, const float* ud32, const slk_codebook* cb, void* ws,
                                        size_t ws_bytes, float* err_sums, void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(r >= 0 && n >= 1, "bad shape");
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(q && d && r32 && ud32, "NULL pointer");
  const DevGrid<float> g = make_grid<float>(cb);
  cudaStream_t st = (cudaStream_t)stream;
  // rows per CTA: the largest tile that still gives a CTA to every third SM -- wide tiles reuse each
  // R element for more rows and hold fewer SM slots while they wait on the column chain (measured
  // on the 72-layer set: 23.8 -> 22.2 ms per pass against +7 % on a single layer's sweep)
  static int64_t want_env = -1;   // SLK_SWEEP_CTAS: CTAs wanted per launch (experiments); default 2/3 of the SMs
  if (want_env < 0) {
    const char* ev = getenv("SLK_SWEEP_CTAS");
    want_env = ev ? atoll(ev) : 0;
  }
  const int64_t want = want_env > 0 ? want_env : (g_opt_sweep_ctas > 0 ? g_opt_sweep_ctas
                                                  : ((int64_t)sm_count() / 48) * 16);   // 48 on a B200: 768 rows -> 16 per CTA
  const GridBreaks xv = make_breaks(cb);        // codebook breakpoints for the leaf's compare tree
  auto fused = [&](int64_t c0, int64_t c1, const float* pacc, float* dhi, float* dlo) -> int {
    if (n % 4 == 0 && c1 - c0 <= MB_COLS && (((uintptr_t)r32) & 15) == 0) {
      if (r >= 32 * want) return launch_macro<32>(q, d, r, n, r32, ud32, g, st, c0, c1, pacc, dhi, dlo, xv, err_sums);
      if (r >= 16 * want) return launch_macro<16>(q, d, r, n, r32, ud32, g, st, c0, c1, pacc, dhi, dlo, xv, err_sums);
      return launch_macro<8>(q, d, r, n, r32, ud32, g, st, c0, c1, pacc, dhi, dlo, xv, err_sums);
    }
    if (r >= 32 * want) return launch_fused<32, true>(q, d, r, n, r32, ud32, g, st, c0, c1, pacc, err_sums);
    if (r >= 16 * want) return launch_fused<16, true>(q, d, r, n, r32, ud32, g, st, c0, c1, pacc, err_sums);
    return launch_fused<8, true>(q, d, r, n, r32, ud32, g, st, c0, c1, pacc, err_sums);
  };
  if (!sweep_macro_ok(r, n, d, rt_hi, rt_lo) || !ws || ws_bytes < slk_gptq_sweep_r_ws_bytes(r, n))
    return fused(0, n, nullptr, nullptr, nullptr);
  float* pacc = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* dhi = pacc + (size_t)r * n;
  float* dlo = dhi + (size_t)r * n;
  SLK_CUDA(cudaMemsetAsync(pacc, 0, (size_t)r * n * sizeof(float), st));
  // narrow layers push on the CUDA cores (push_simt_kernel); SLK_SWEEP_SIMT_N: widest such layer (0 = never)
  static int64_t simt_n = -1;
  if (simt_n < 0) {
    const char* ev = getenv("SLK_SWEEP_SIMT_N");
    simt_n = ev ? atoll(ev) : 1024;
  }
  const bool simt = n <= simt_n && (((uintptr_t)d) & 15) == 0;
  auto push = [&](int64_t k0, int64_t k1, int64_t c0, int64_t c1) -> int {   // Pacc[:, c0:c1] += D[:, k0:k1] R[k0:k1, c0:c1]
    if (simt) {
      dim3 grid((unsigned)ceil_div(c1 - c0, PS_BN), (unsigned)ceil_div(r, PS_BM));
      push_simt_kernel<<<grid, 256, 0, st>>>(d + k0, n, r32 + k0 * n + c0, n, pacc + c0, n, r, c1 - c0, k1 - k0);
      SLK_LAUNCH_CHECK();
      return SLK_OK;
    }
    TcParams p;
    p.C = pacc + c0; p.ldc = n; p.R = nullptr; p.R2 = nullptr; p.ldr = 0;
    p.M = r; p.N = c1 - c0; p.K = k1 - k0;
    p.alpha = 1.0f; p.keep = 0.0f; p.count = 1.0f; p.error_flag = nullptr;
    return tc_gemm_presplit_f32(TC_ACCUM, dhi + k0, dlo + k0, n, rt_hi + c0 * n + k0, rt_lo + c0 * n + k0, n, p, st);
  };
  // Two levels of lazy batching for wide layers (the read-modify-write of Pacc is HBM traffic:
  // r (n - e) 8 bytes per GEMM): inside a super block of SWEEP_SUPER columns the macro-block GEMMs
  // only reach to its end; one K = SWEEP_SUPER GEMM then pushes the whole super block to the rest.
  const int64_t super = n > 2 * SWEEP_SUPER ? SWEEP_SUPER : n;
  for (int64_t S0 = 0; S0 < n; S0 += super) {
    const int64_t S1 = (S0 + super) < n ? (S0 + super) : n;
    for (int64_t s0 = S0; s0 < S1; s0 += SWEEP_MB) {
      const int64_t e0 = (s0 + SWEEP_MB) < S1 ? (s0 + SWEEP_MB) : S1;
      rc = fused(s0, e0, pacc, (e0 < n && !simt) ? dhi : nullptr, dlo);
      if (rc) return rc;
      if (e0 < S1 && (rc = push(s0, e0, e0, S1))) return rc;
    }
    if (S1 < n && (rc = push(S0, S1, S1, n))) return rc;
  }
  return SLK_OK;
}

// ---- layer error from the sweep's row sums ----------------------------------------------------------
// rows_out[row] = scale[row]^2 * (sum E^2 - damp_abs * sum D^2)   (channelwise_error, obq.py:89-95, of the
// de-scaled result: W - Q_descaled = scale * D);  mean_out = their mean (quantization_error, obq.py:98-103).
namespace slk {
__global__ void __launch_bounds__(256) sweep_error_kernel(const float2* __restrict__ sums, const float* __restrict__ scale,
                                                          const float* __restrict__ dampval, int64_t r,
                                                          float* __restrict__ rows_out, float* __restrict__ mean_out) {
  __shared__ double scratch[32];
  const float lam = dampval ? __ldg(dampval) : 0.0f;
  double acc = 0.0;
  for (int64_t j = threadIdx.x; j < r; j += blockDim.x) {
    const float2 v = sums[j];
    float e = __fsub_rn(v.x, __fmul_rn(lam, v.y));
    if (scale) {
      const float s = scale[j];
      e = __fmul_rn(e, __fmul_rn(s, s));
    }
    if (rows_out) rows_out[j] = e;
    acc += (double)e;
  }
  acc = block_sum<double>(acc, scratch);
  if (threadIdx.x == 0 && mean_out) mean_out[0] = __fdiv_rn((float)acc, (float)r);
}
}  // namespace slk

extern "C" int slk_sweep_error_f32(const float* err_sums, const float* row_scale, const float* damp_abs, int64_t r,
                                   float* rows_out, float* mean_out, void* stream) {
  SLK_REQUIRE(err_sums && r >= 1 && (rows_out || mean_out), "bad arguments");
  sweep_error_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(err_sums), row_scale, damp_abs, r,
                                                          rows_out, mean_out);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
