// Tensor-core GEMM for the dense phases: tcgen05.mma (kind::tf32) with TMEM accumulators, operands
// staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), fp32-faithful through a
// 3xTF32 split.
//
//   D[M, N] = A[M, K] * B[N, K]^T          (both operands K-major fp32, row pitch multiple of 16 B)
//
// Precision: a single TF32 pass changes ~1 % of the GPTQ codes; the reference's
// GEMMs are fp32 (sgemm) or better.  Every operand x is therefore split once, in HBM, into
//   hi = x with the 13 low mantissa bits cleared (what kind::tf32 would read anyway)
//   lo = x - hi  (exact), rounded to nearest TF32
// and each k-step issues three MMAs: lo*hi and hi*lo into a cross-term TMEM tile, hi*hi into the
// main one.  The dropped lo*lo term is <= 2^-20 relative, i.e. fp32 round-off level.
//
// Kernel shape: one 128 x BN output tile per CTA, 192 threads = 6 warps:
//   warp 0      TMA producer (one elected lane): 4 tiles per stage (A_hi, A_lo, B_hi, B_lo)
//   warp 1      TMEM allocation + MMA issue (one elected lane), tcgen05.commit -> mbarriers
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 16 columns at a time, fused epilogue, global store
// Pipeline: STAGES-deep ring of {full, empty} mbarriers between TMA and MMA; a 2-deep ring of
// {hi_full, hi_empty} between MMA and epilogue (see TC_KC).  TMEM: 2 x BN columns of hi*hi
// accumulators + BN columns of cross terms, 128 lanes each.
#include "tc_gemm.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace slk {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;            // 32 fp32 = 128 bytes = one swizzle row
constexpr int TC_UMMA_K = 8;         // kind::tf32: 32 bytes of K per instruction
constexpr int TC_THREADS = 192;


// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline must fail loudly, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag) {
  for (uint32_t spin = 0; spin < (1u << 26); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  if (error_flag) atomicExch(error_flag, 1);
  __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor of a K-major, 128B-swizzled tile (rows of 128 bytes, 8-row groups
// 1024 bytes apart): start address, LBO (unused for swizzled K-major) = 1, SBO = 1024 B, version 1,
// layout SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::tf32 instruction descriptor: D fp32, A/B TF32, both K-major, M x N (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with BF16 operands, fp32 accumulation: a 128-byte swizzle row holds 64 elements, one instruction 16
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN, int STAGES, int PASSES = 3>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 4;
  // PASSES == 3: {A_hi, A_lo, B_hi, B_lo} per stage; PASSES == 1 (screening pass): {A, B}
  static constexpr int B_OFF = PASSES == 3 ? 2 * A_BYTES : A_BYTES;
  static constexpr int STAGE_BYTES = PASSES == 3 ? 2 * A_BYTES + 2 * B_BYTES : A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// K-blocks (of 32) whose hi*hi products share one TMEM accumulation before being drained.
// Measured on B200: the tensor core's fp32 accumulator truncates, so the error of a long TMEM
// accumulation grows linearly with the number of MMAs (2.3e-5 relative at K = 3072, vs 2.4e-6 for
// an fp32 FMA GEMM).  The hi*hi products are therefore accumulated in TMEM only over TC_KC
// k-blocks (16 MMAs), drained, and summed in fp32 registers with round-to-nearest adds by the
// epilogue warps while the next chunk runs in the other TMEM buffer; the cross terms (2^-11
// smaller, so their truncation is irrelevant) keep one TMEM tile for the whole K loop.
constexpr int TC_KC = 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// PASSES == 1: a single kind::tf32 product of the operands as given (the caller rounds them to TF32), accumulated
// over the whole K range in one TMEM tile -- ~1e-3 relative, a third of the tensor time and half of the operand
// traffic: the screening pass of the full-H scale search (dense.cu), never a result on its own.
// BF16 (with PASSES == 1): the operands are bf16 matrices (TMA boxes of 64 elements = the same 128-byte rows),
// the product is kind::f16 -- half the operand bytes and twice the tensor rate of the TF32 pass, ~4e-3 relative.
template <int BN, int STAGES, int EPI, int PASSES = 3, bool BF16 = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, TcParams p) {
  typedef TcSmem<BN, STAGES, PASSES> SM;
  static_assert(PASSES == 1 || PASSES == 3, "one TF32 pass or the 3xTF32 split");
  static_assert((PASSES == 3 ? 3 : 1) * BN <= 512, "two hi buffers and one lo tile must fit the 512 TMEM columns");
  static_assert(PASSES == 3 || EPI == TC_ROWDOT, "the single-pass product only feeds the row-dot screening");
  static_assert(!BF16 || PASSES == 1, "bf16 operands: screening pass only");
  constexpr int BK = BF16 ? 2 * TC_BK : TC_BK;   // elements per 128-byte k-block
  // the single-pass tile allocates only its own accumulator columns, so that two CTAs can share an SM (one's
  // epilogue under the other's main loop) when their shared memory allows it
  constexpr uint32_t TMEM_COLS = PASSES == 1 ? (uint32_t)BN : 512u;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* tiles = (uint8_t*)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);   // SW128 needs 1024 B alignment
  uint64_t* bars = (uint64_t*)(tiles + STAGES * SM::STAGE_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* hi_full = bars + 2 * STAGES;     // [2]       MMA -> epilogue (chunk accumulated)
  uint64_t* hi_empty = bars + 2 * STAGES + 2;  // [2]     epilogue -> MMA (buffer drained)
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  if ((EPI == TC_HESS_SYM || EPI == TC_SYM_PART) && n0 + BN <= m0) return;   // strictly-lower tile of a symmetric product
  if (p.m_limit != nullptr && m0 >= __ldg(p.m_limit)) return;                // beyond the device-side row count
  const int num_kb_all = (int)((p.K + BK - 1) / BK);
  // split-K (TC_SYM_PART): this CTA covers k-blocks [kb0, kb0 + num_kb) and writes plane blockIdx.z
  const int kb0 = EPI == TC_SYM_PART ? (int)blockIdx.z * p.kb_per_split : 0;
  const int num_kb = EPI == TC_SYM_PART ? max(0, min(p.kb_per_split, num_kb_all - kb0)) : num_kb_all;
  const int num_chunks = PASSES == 1 ? (num_kb > 0 ? 1 : 0) : (num_kb + TC_KC - 1) / TC_KC;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(hi_full + b, 1); mbar_init(hi_empty + b, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_lo = tmem_base + 2 * BN;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----------------------------------------------------------------------
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(empty_bar + s, ph ^ 1u, p.error_flag);
        uint8_t* st = tiles + s * SM::STAGE_BYTES;
        mbar_expect_tx(full_bar + s, SM::STAGE_BYTES);
        const int k0 = (kb0 + kb) * BK;
        tma_load_2d(st, &map_a_hi, full_bar + s, k0, m0);
        tma_load_2d(st + SM::B_OFF, &map_b_hi, full_bar + s, k0, n0);
        if (PASSES == 3) {
          tma_load_2d(st + SM::A_BYTES, &map_a_lo, full_bar + s, k0, m0);
          tma_load_2d(st + SM::B_OFF + SM::B_BYTES, &map_b_lo, full_bar + s, k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ------------------------------------------------------------------------
      constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(TC_BM, BN) : umma_idesc_tf32(TC_BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        const int chunk = PASSES == 1 ? 0 : kb / TC_KC, buf = chunk & 1;
        const bool chunk_first = PASSES == 1 ? kb == 0 : kb % TC_KC == 0;
        const bool chunk_last = PASSES == 1 ? kb == num_kb - 1 : (kb % TC_KC == TC_KC - 1 || kb == num_kb - 1);
        if (chunk_first) {   // new chunk: its TMEM buffer must have been drained
          mbar_wait(hi_empty + buf, ((uint32_t)(chunk >> 1) & 1u) ^ 1u, p.error_flag);
          tc_fence_after();
        }
        mbar_wait(full_bar + s, ph, p.error_flag);
        tc_fence_after();
        const uint32_t st = smem_u32(tiles + s * SM::STAGE_BYTES);
        const uint64_t a_hi = umma_desc_sw128(st), a_lo = umma_desc_sw128(st + SM::A_BYTES);
        const uint64_t b_hi = umma_desc_sw128(st + SM::B_OFF);
        const uint64_t b_lo = umma_desc_sw128(st + SM::B_OFF + SM::B_BYTES);
        const uint32_t d_hi = tmem_base + (uint32_t)(buf * BN);
#pragma unroll
        for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
          const uint64_t adv = (uint64_t)((k * TC_UMMA_K * 4) >> 4);   // +32 bytes along K inside the swizzle row
          if (PASSES == 3) {
            tc_mma_tf32(tmem_lo, a_lo + adv, b_hi + adv, idesc, (kb == 0 && k == 0) ? 0u : 1u);
            tc_mma_tf32(tmem_lo, a_hi + adv, b_lo + adv, idesc, 1u);
          }
          if (BF16) tc_mma_bf16(d_hi, a_hi + adv, b_hi + adv, idesc, (chunk_first && k == 0) ? 0u : 1u);
          else tc_mma_tf32(d_hi, a_hi + adv, b_hi + adv, idesc, (chunk_first && k == 0) ? 0u : 1u);
        }
        tc_commit(empty_bar + s);            // frees the smem stage once these MMAs have read it
        if (chunk_last) tc_commit(hi_full + buf);   // chunk complete
      }
    }
  } else {
    // ---- epilogue warps: TMEM lane quarter = warp % 4 ------------------------------------------
    const int quarter = warp & 3;
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const int64_t row = (int64_t)m0 + quarter * 32 + lane;
    const bool row_ok = row < p.M;
    float run[PASSES == 3 ? BN : 1];
    if (PASSES == 3) {
#pragma unroll
      for (int j = 0; j < (PASSES == 3 ? BN : 1); ++j) run[j] = 0.0f;
      for (int chunk = 0; chunk < num_chunks; ++chunk) {
        const int buf = chunk & 1;
        mbar_wait(hi_full + buf, (uint32_t)(chunk >> 1) & 1u, p.error_flag);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < (PASSES == 3 ? BN : 0); c0 += 16) {
          float v[16];
          tc_ld16(tmem_base + lane_sel + (uint32_t)(buf * BN + c0), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) run[c0 + j] = __fadd_rn(run[c0 + j], v[j]);
        }
        tc_fence_before();
        mbar_arrive(hi_empty + buf);
      }
    } else if (num_chunks > 0) {   // the one accumulation over the whole K range
      mbar_wait(hi_full, 0u, p.error_flag);
      tc_fence_after();
    }
    // the last hi_full commit also covers every cross-term MMA
    float dot = 0.0f;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tc_ld16((PASSES == 3 ? tmem_lo : tmem_base) + lane_sel + (uint32_t)c0, v);
      if (num_kb == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0f;       // empty k range: TMEM was never written
      }
      const int64_t col = (int64_t)n0 + c0;
      if (!row_ok || col >= p.N) continue;
      if (PASSES == 3) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __fadd_rn(run[c0 + j], v[j]);
      }
      if (EPI == TC_HESS_SYM) {
        // symmetric accumulate: only n >= m is computed here; each value is written to (m, n) and
        // mirrored to (n, m) (coalesced: consecutive lanes are consecutive m), so H stays
        // symmetric to the bit and the tiles below the diagonal are never launched
        float* cc = p.C + row * p.ldc + col;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int64_t n = col + j;
          if (n < p.N && n >= row) {
            const float o = __fadd_rn(__fmul_rn(cc[j], p.keep), __fdiv_rn(v[j], p.count));
            cc[j] = o;
            p.C[n * p.ldc + row] = o;
          }
        }
      } else if (EPI == TC_SYM_PART) {
        float* cc = p.C + (int64_t)blockIdx.z * p.M * p.ldc + row * p.ldc + col;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < p.N) cc[j] = v[j];
      } else if (EPI == TC_ROWDOT && BF16) {
        const __nv_bfloat16* rr = reinterpret_cast<const __nv_bfloat16*>(p.R) + row * p.ldr + col;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < p.N) dot = __fmaf_rn(v[j], __bfloat162float(rr[j]), dot);
      } else if (EPI == TC_ROWDOT) {
        const float* rr = p.R + row * p.ldr + col;
        const float* r2 = p.R2 ? p.R2 + row * p.ldr + col : nullptr;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < p.N) {
            float e = __ldg(rr + j);
            if (r2) e = __fsub_rn(e, __ldg(r2 + j));
            dot = __fmaf_rn(v[j], e, dot);
          }
      } else {
        float* cc = p.C + row * p.ldc + col;
        const bool vec = (col + 16 <= p.N) && ((((uintptr_t)cc) & 15) == 0);
        if (vec) {
          float4* c4 = reinterpret_cast<float4*>(cc);
          float4 o[4];
          if (EPI != TC_STORE) {
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = c4[q];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float* of = reinterpret_cast<float*>(&o[q]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = v[4 * q + j];
              if (EPI == TC_STORE) of[j] = __fmul_rn(p.alpha, a);
              else if (EPI == TC_ACCUM) of[j] = __fadd_rn(of[j], __fmul_rn(p.alpha, a));
              else of[j] = __fadd_rn(__fmul_rn(of[j], p.keep), __fdiv_rn(a, p.count));
            }
            c4[q] = o[q];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (col + j >= p.N) break;
            const float a = v[j];
            if (EPI == TC_STORE) cc[j] = __fmul_rn(p.alpha, a);
            else if (EPI == TC_ACCUM) cc[j] = __fadd_rn(cc[j], __fmul_rn(p.alpha, a));
            else cc[j] = __fadd_rn(__fmul_rn(cc[j], p.keep), __fdiv_rn(a, p.count));
          }
        }
      }
    }
    if (EPI == TC_ROWDOT && row_ok) p.C[row * p.ldc + blockIdx.x] = dot;
    tc_fence_before();
  }
  __syncwarp();   // the single-lane roles rejoin their warps before the block barrier
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---- hi / lo split ------------------------------------------------------------------------------
// hi = x (- y) with the 13 low mantissa bits cleared; lo = (x - hi) rounded to nearest TF32.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         int64_t rows, int64_t cols, int64_t ldx, int64_t ldo,
                                                         float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t total = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t r = t / cols, c = t - r * cols;
    float v = x[r * ldx + c];
    if (y) v = __fsub_rn(v, y[r * ldx + c]);
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const float l = __fsub_rn(v, h);
    hi[r * ldo + c] = h;
    lo[r * ldo + c] = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
  }
}

// transposing variant for K1: in[S, n] (row pitch ldx) -> hi/lo [n, S] (row pitch ldo), 32x32 smem tiles
__global__ void __launch_bounds__(256) split_tf32_transpose_kernel(const float* __restrict__ x, int64_t S, int64_t n,
                                                                   int64_t ldx, int64_t ldo, float* __restrict__ hi,
                                                                   float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int64_t s0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int64_t s = s0 + i, c = c0 + tx;
    tile[i][tx] = (s < S && c < n) ? x[s * ldx + c] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t c = c0 + i, s = s0 + tx;
    if (c < n && s < S) {
      const float v = tile[tx][i];
      const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
      const float l = __fsub_rn(v, h);
      hi[c * ldo + s] = h;
      lo[c * ldo + s] = __uint_as_float((__float_as_uint(l) + 0x1000u) & 0xffffe000u);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// rows x cols fp32, row pitch ld elements; box = box_rows x 32 columns, 128B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    bool bf16 = false) {
  EncodeTiledFn fn = encode_fn();
  SLK_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 2 * TC_BK : TC_BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SLK_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)rc);
  return SLK_OK;
}

bool tc_gemm_usable(const void* a, int64_t lda, const void* b, int64_t ldb) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("SLK_DISABLE_TC");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (disabled) return false;
  return ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && (lda % 4 == 0) && (ldb % 4 == 0) && encode_fn() != nullptr;
}

template <int BN, int STAGES, int EPI, int PASSES = 3, bool BF16 = false>
static int tc_launch_s(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo,
                       int64_t ldb, const TcParams& p, cudaStream_t st, int splits = 1) {
  typedef TcSmem<BN, STAGES, PASSES> SM;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a_hi, p.M, p.K, lda, TC_BM, BF16))) return rc;
  if ((rc = make_map(&ma_lo, a_lo, p.M, p.K, lda, TC_BM, BF16))) return rc;
  if ((rc = make_map(&mb_hi, b_hi, p.N, p.K, ldb, BN, BF16))) return rc;
  if ((rc = make_map(&mb_lo, b_lo, p.N, p.K, ldb, BN, BF16))) return rc;
  auto kern = tc_gemm_kernel<BN, STAGES, EPI, PASSES, BF16>;
  SLK_SMEM_ATTR_ONCE(kern, SM::TOTAL);
  dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, TC_BM), (unsigned)splits);
  kern<<<grid, TC_THREADS, SM::TOTAL, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

template <int BN, int EPI>
static int tc_launch(const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi, const float* b_lo,
                     int64_t ldb, const TcParams& p, cudaStream_t st, int splits = 1) {
  return tc_launch_s<BN, (BN == 128 ? 3 : 4), EPI>(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, st, splits);
}

static inline int split_grid(int64_t total) {
  int64_t b = ceil_div(total, 256), cap = (int64_t)sm_count() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// D = A * B^T with the epilogue `epi`; A [M, K] and B [N, K] fp32 K-major.  ws holds the hi/lo copies.
int tc_gemm_f32(int epi, const float* A, const float* A2, int64_t lda, const float* B, int64_t ldb, TcParams p,
                void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t Kp = (p.K + 3) / 4 * 4;
  const size_t a_elems = (size_t)p.M * Kp, b_elems = (size_t)p.N * Kp;
  SLK_REQUIRE(ws && ws_bytes >= (2 * a_elems + 2 * b_elems) * sizeof(float) + 1024, "tc_gemm workspace too small");
  float* a_hi = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* a_lo = a_hi + a_elems;
  float* b_hi = a_lo + a_elems;
  float* b_lo = b_hi + b_elems;
  split_tf32_kernel<<<split_grid(p.M * p.K), 256, 0, st>>>(A, A2, p.M, p.K, lda, Kp, a_hi, a_lo);
  SLK_LAUNCH_CHECK();
  split_tf32_kernel<<<split_grid(p.N * p.K), 256, 0, st>>>(B, nullptr, p.N, p.K, ldb, Kp, b_hi, b_lo);
  SLK_LAUNCH_CHECK();
  switch (epi) {
    case TC_STORE: return tc_launch<128, TC_STORE>(a_hi, a_lo, Kp, b_hi, b_lo, Kp, p, st);
    case TC_ACCUM: return tc_launch<128, TC_ACCUM>(a_hi, a_lo, Kp, b_hi, b_lo, Kp, p, st);
    case TC_HESS: return tc_launch<128, TC_HESS>(a_hi, a_lo, Kp, b_hi, b_lo, Kp, p, st);
    case TC_ROWDOT: return tc_launch<128, TC_ROWDOT>(a_hi, a_lo, Kp, b_hi, b_lo, Kp, p, st);
  }
  SLK_REQUIRE(false, "unknown tc epilogue %d", epi);
  return SLK_ERR_ARG;
}

// Same product from operands that are already split (hi = value with the 13 low mantissa bits
// cleared, lo = RN_tf32(value - hi)); no workspace, one launch.
int tc_gemm_presplit_f32(int epi, const float* a_hi, const float* a_lo, int64_t lda, const float* b_hi,
                         const float* b_lo, int64_t ldb, TcParams p, cudaStream_t st) {
  switch (epi) {
    case TC_STORE: return tc_launch<128, TC_STORE>(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, st);
    case TC_ACCUM: {
      // The sweep's macro-block pushes have K = 256: eight k-blocks.  Two pipeline stages (129 KB of shared
      // memory instead of 193 KB) then cost nothing and let the tile share an SM with one sweep CTA of
      // another layer instead of waiting for an empty SM (SLK_TC_STAGES2=0: always three stages).
      static int two = -1;
      if (two < 0) {
        const char* e = getenv("SLK_TC_STAGES2");
        two = (e && e[0] == '0') ? 0 : 1;
      }
      if (two && p.K <= 512) return tc_launch_s<128, 2, TC_ACCUM>(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, st);
      return tc_launch<128, TC_ACCUM>(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, st);
    }
    case TC_ROWDOT: return tc_launch<128, TC_ROWDOT>(a_hi, a_lo, lda, b_hi, b_lo, ldb, p, st);
  }
  SLK_REQUIRE(false, "tc_gemm_presplit: unsupported epilogue %d", epi);
  return SLK_ERR_ARG;
}

// Screening product (see the PASSES == 1 note at the kernel): rows of (A B^T) dotted with the rows of p.R,
// partials per column tile of width bn (128 or 256) in p.C [M, ceil(N / bn)].  A and B hold TF32-representable
// values.  ctas = 1: six (bn = 128) or four (bn = 256) 32 / 48 KB stages, one CTA per SM; ctas = 2: half the
// stages (96 KB), two CTAs per SM.  (Measured and dropped: two 128-row tiles per CTA sharing each B tile --
// 64 KB per 256 x 256 x 64 block, one CTA per SM -- 3.7 ms against 3.4 ms at [102400, 4096] x 4096: its
// epilogue cannot overlap a main loop, the 512 TMEM columns being taken by the one tile pair.)
int tc_gemm_screen_f32(const float* a, int64_t lda, const float* b, int64_t ldb, TcParams p, int bn, int ctas,
                       cudaStream_t st) {
  if (bn == 256 && ctas == 2) return tc_launch_s<256, 2, TC_ROWDOT, 1>(a, a, lda, b, b, ldb, p, st);
  if (bn == 256) return tc_launch_s<256, 4, TC_ROWDOT, 1>(a, a, lda, b, b, ldb, p, st);
  SLK_REQUIRE(bn == 128, "tc_gemm_screen: tile width %d", bn);
  if (ctas == 2) return tc_launch_s<128, 3, TC_ROWDOT, 1>(a, a, lda, b, b, ldb, p, st);
  return tc_launch_s<128, 6, TC_ROWDOT, 1>(a, a, lda, b, b, ldb, p, st);
}
// same with bf16 operands (pitches in elements, multiples of 8; p.R: the bf16 rows, pitch p.ldr elements)
int tc_gemm_screen_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, TcParams p, int bn, int ctas,
                        cudaStream_t st) {
  SLK_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "tc_gemm_screen_bf16: row pitches must be multiples of 16 bytes");
  if (bn == 256 && ctas == 2) return tc_launch_s<256, 2, TC_ROWDOT, 1, true>(a, a, lda, b, b, ldb, p, st);
  if (bn == 256) return tc_launch_s<256, 4, TC_ROWDOT, 1, true>(a, a, lda, b, b, ldb, p, st);
  SLK_REQUIRE(bn == 128, "tc_gemm_screen: tile width %d", bn);
  if (ctas == 2) return tc_launch_s<128, 3, TC_ROWDOT, 1, true>(a, a, lda, b, b, ldb, p, st);
  return tc_launch_s<128, 6, TC_ROWDOT, 1, true>(a, a, lda, b, b, ldb, p, st);
}

// hi / lo TF32 parts of a [rows, cols] matrix (row pitch ld) into two matrices of pitch ldo
int tc_split_f32(const float* x, int64_t rows, int64_t cols, int64_t ld, int64_t ldo, float* hi, float* lo,
                 cudaStream_t st) {
  split_tf32_kernel<<<split_grid(rows * cols), 256, 0, st>>>(x, nullptr, rows, cols, ld, ldo, hi, lo);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

size_t tc_gemm_ws_bytes(int64_t M, int64_t N, int64_t K) {
  const int64_t Kp = (K + 3) / 4 * 4;
  return (size_t)(2 * M * Kp + 2 * N * Kp) * sizeof(float) + 1024;
}

// Split-K reduction of the symmetric Hessian product: planes are added in order (deterministic),
// then H = H*keep + D/count is applied to (m, n) and mirrored to (n, m) (statistics.py:82-87).
__global__ void __launch_bounds__(256) sym_part_reduce_kernel(const float* __restrict__ part, int splits, int64_t n,
                                                              float* __restrict__ H, float keep, float count) {
  const int64_t total = n * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t m = t / n, c = t - m * n;
    if (c < m) continue;
    float d = part[t];
    for (int z = 1; z < splits; ++z) d = __fadd_rn(d, part[(int64_t)z * total + t]);
    const float o = __fadd_rn(__fmul_rn(H[t], keep), __fdiv_rn(d, count));
    H[t] = o;
    H[c * n + m] = o;
  }
}

static inline int sym_splits(int64_t M, int64_t K) {
  const int64_t tm = ceil_div(M, TC_BM);
  const int64_t tiles = tm * (tm + 1) / 2;
  const int64_t kb = ceil_div(K, TC_BK);
  int s = (int)((int64_t)sm_count() / (tiles > 0 ? tiles : 1));
  if (s > 8) s = 8;
  if ((int64_t)s > kb / 8) s = (int)(kb / 8);      // at least 8 k-blocks (two drained chunks) per split
  return s < 2 ? 1 : s;
}

// D = At^T * At (M = N = columns of At): one transposing split feeds both operands.
int tc_gemm_at_f32(int epi, const float* At, int64_t ldat, TcParams p, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t Kp = (p.K + 3) / 4 * 4;
  const size_t elems = (size_t)p.M * Kp;
  SLK_REQUIRE(p.M == p.N, "tc_gemm_at: square output expected");
  SLK_REQUIRE(ws && ws_bytes >= 2 * elems * sizeof(float) + 1024, "tc_gemm_at workspace too small");
  float* hi = (float*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  float* lo = hi + elems;
  dim3 grid((unsigned)ceil_div(p.M, 32), (unsigned)ceil_div(p.K, 32));
  split_tf32_transpose_kernel<<<grid, 256, 0, st>>>(At, p.K, p.M, ldat, Kp, hi, lo);
  SLK_LAUNCH_CHECK();
  if (epi == TC_HESS_SYM) {
    // few output tiles (small n): split K over blockIdx.z so that the grid covers the SMs
    const int splits = sym_splits(p.M, p.K);
    if (splits > 1 && ws_bytes >= tc_gemm_at_ws_bytes(p.M, p.K)) {
      float* part = lo + elems;
      TcParams q = p;
      q.C = part; q.ldc = p.N;
      q.kb_per_split = (int)ceil_div(ceil_div(p.K, TC_BK), splits);
      int rc = tc_launch<128, TC_SYM_PART>(hi, lo, Kp, hi, lo, Kp, q, st, splits);
      if (rc) return rc;
      sym_part_reduce_kernel<<<split_grid(p.M * p.N), 256, 0, st>>>(part, splits, p.M, p.C, p.keep, p.count);
      SLK_LAUNCH_CHECK();
      return SLK_OK;
    }
  }
  switch (epi) {
    case TC_STORE: return tc_launch<128, TC_STORE>(hi, lo, Kp, hi, lo, Kp, p, st);
    case TC_HESS: return tc_launch<128, TC_HESS>(hi, lo, Kp, hi, lo, Kp, p, st);
    case TC_HESS_SYM: return tc_launch<128, TC_HESS_SYM>(hi, lo, Kp, hi, lo, Kp, p, st);
  }
  SLK_REQUIRE(false, "tc_gemm_at: unsupported epilogue %d", epi);
  return SLK_ERR_ARG;
}

size_t tc_gemm_at_ws_bytes(int64_t M, int64_t K) {
  const int64_t Kp = (K + 3) / 4 * 4;
  const int splits = sym_splits(M, K);
  return (size_t)(2 * M * Kp + (splits > 1 ? (int64_t)splits * M * M : 0)) * sizeof(float) + 1024;
}

}  // namespace slk

using namespace slk;

extern "C" {

/* Generic entry: D = op(A [- A2]) * B^T, see include/sleekit_b200.h */
size_t slk_tc_gemm_ws_bytes(int64_t M, int64_t N, int64_t K) { return tc_gemm_ws_bytes(M, N, K); }

int slk_tc_gemm_f32(int32_t epilogue, const float* a, const float* a2, int64_t lda, const float* b, int64_t ldb,
                    float* c, int64_t ldc, const float* rowdot, int64_t ldr, int64_t M, int64_t N, int64_t K,
                    float alpha, float keep, float count, void* ws, size_t ws_bytes, int32_t* error_flag,
                    void* stream) {
  SLK_REQUIRE(a && b && c && M >= 1 && N >= 1 && K >= 1, "bad arguments");
  SLK_REQUIRE(epilogue >= 0 && epilogue <= 3, "epilogue %d", epilogue);
  SLK_REQUIRE(epilogue != TC_ROWDOT || rowdot != nullptr, "row-dot epilogue needs the row matrix");
  SLK_REQUIRE(encode_fn() != nullptr, "TMA descriptors unavailable (driver too old?)");
  TcParams p;
  p.C = c; p.ldc = ldc; p.R = rowdot; p.R2 = nullptr; p.ldr = ldr; p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.keep = keep; p.count = count; p.error_flag = error_flag;
  return tc_gemm_f32(epilogue, a, a2, lda, b, ldb, p, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
