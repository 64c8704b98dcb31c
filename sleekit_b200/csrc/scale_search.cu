// K5: fused scale-grid search for the MSE and diagonal-H criteria.
//   compute_min_mse_scaling (H None or 1-D)        scaling.py:98-134
//   -> compute_non_saturating_scaling              scaling.py:44-55
//   -> quantize_with_scaling / apply_scaling        scaling.py:58-81, 21-25
//   -> _compute_mse                                 scaling.py:84-90
// The reference makes G (=100) passes over W, each ~12 numpy passes.  Here one CTA owns one
// row: the row is read from HBM once into shared memory, every thread owns one grid point and
// walks the whole row (shared-memory broadcast reads, no cross-thread reduction), so HBM sees
// 4 bytes per weight in total and the kernel is bound by the fp32 pipe (three IEEE divides per
// weight and grid point).  Every arithmetic step is the reference's, separately rounded:
//   scale = f*init; rs = 1/scale; x = w/scale; v = quant(x); dq = v/rs; e = dq - w; err += h*e*e
#include "common.cuh"

#include <stdlib.h>

namespace slk {

// Error of one row at one grid point.  FAST: the three divides (by scale, by the codebook step,
// by rs) go through fastdiv_core -- same correctly rounded quotients, 5 FMA-pipe ops each.
template <typename HT, bool HAS_H, bool FAST>
__device__ __forceinline__ HT row_error(const float* __restrict__ rowp, int64_t n, const DevGrid<float>& g,
                                        const FastDivF& fs, const FastDivF& fr, const FastDivF& fstep,
                                        const HT* __restrict__ hdiag) {
  const float hi = (float)(g.size - 1);
  HT acc[4] = {(HT)0, (HT)0, (HT)0, (HT)0};
  auto one = [&](float w) -> float {
    float v;
    if (FAST) {
      const float x = fastdiv_core(w, fs.d, fs.y);                          // scaling.py:73
      float k = rintf(fastdiv_core(__fsub_rn(x, g.zero), fstep.d, fstep.y)); // codebook.py:60-62
      k = k < 0.0f ? 0.0f : k;
      k = k > hi ? hi : k;
      v = __fadd_rn(__fmul_rn(k, g.step), g.zero);                          // codebook.py:63-64
      v = fastdiv_core(v, fr.d, fr.y);                                      // scaling.py:80
    } else {
      v = __fdiv_rn(grid_value(g, __fdiv_rn(w, fs.d)), fr.d);
    }
    const float e = __fsub_rn(v, w);                                        // scaling.py:130
    return __fmul_rn(e, e);
  };
  int64_t j = 0;
  for (; j + 4 <= n; j += 4) {
    float e2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) e2[u] = one(rowp[j + u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[u] += HAS_H ? (HT)e2[u] * __ldg(hdiag + j + u) : (HT)e2[u];
  }
  for (; j < n; ++j) {
    const float e2 = one(rowp[j]);
    acc[0] += HAS_H ? (HT)e2 * __ldg(hdiag + j) : (HT)e2;
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// ---- threshold-table form (uniform codebooks of <= 16 entries, fp32 errors) ------------------------
// For a fixed (row, grid point) the chain  w -> w/scale -> (x - zero)/step -> rint -> clip  is a
// composition of monotone non-decreasing correctly rounded operations, hence the code index is a
// non-decreasing step function of w with C-1 breakpoints.  Per row and grid point the breakpoints
// T_k = min{w : idx(w) >= k} are found EXACTLY by bisection over the ordered fp32 values using the
// reference's own op chain, and the C de-quantised values (k*step + zero)/rs are computed once with
// that same chain; both go to shared memory.  The main loop is then transposed with respect to the
// direct kernel: a thread owns 8 WEIGHTS (in registers) and walks the grid points; per weight and
// grid point it forms a biased index estimate with one FMA (k_a in {k-1, k}: the bias 0.5 exceeds
// the estimate's error by four orders of magnitude), looks T[k_a+1] up, corrects, looks the value
// up, and accumulates h*(v-w)^2:  ~13 issue slots instead of the ~36 of three exact divides, split
// over the FMA, ALU and LSU pipes.  Every weight receives exactly the value the reference's chain
// gives it (non-NaN inputs), so the chosen grid point is unchanged.
//
// Walk form of the main loop (default).  Along an ascending grid the thresholds move away from zero
// (T_k ~ X_k * scale), so the exact index of a weight can only step towards the centre of the
// codebook, and by at most one code per grid point when the grid is fine enough.  Both facts are
// statements about the exact tables and are CHECKED on them per row (two comparisons per (grid point,
// code) and class: see `walk conditions` below; a row that fails uses the estimate loop).  Negative
// weights are served by a mirrored copy of the tables (below), so that for every weight the index
// steps DOWN and one code path serves both signs.  The thread keeps, per weight, the shared-memory
// address of the table entry that decides its next move, found once at the first grid point with the
// estimate; per (weight, grid point): threshold load, compare, predicated address step, value load
// (both loads with immediate offsets), sub, mul, fma -- 7 issue slots instead of 14.  The 16
// per-grid-point partial sums of a block of grid points are reduced across the warp by recursive
// halving (16 shuffles for 16 sums instead of 80).  Same exact indices, hence the same values and
// the same chosen grid point.
constexpr int TAB_MAXC = 16;
constexpr int TAB_THREADS = 128;
constexpr int TAB_MAXG = 128;     // grid points per row held in shared memory

__device__ __forceinline__ int f32_ord(float x) {
  const int i = __float_as_int(x);
  return i >= 0 ? i : (int)(0x80000000u - (unsigned)i);
}
__device__ __forceinline__ float f32_unord(int o) {
  return __int_as_float(o >= 0 ? o : (int)(0x80000000u - (unsigned)o));
}

// Shared memory: a fixed header, then two table regions of G rows each (dynamic, sized by G).
// Row g of a region: T[0..C] (T[0] = -inf, T[C] = +inf) then, PITCH floats later, the C values.
// Region 0 serves the weights w >= 0 as they are.  Region 1 serves the weights w < 0 MIRRORED:
// the thread holds w' = -w > 0, and with  T'[j] = nextup(-T[C-j]),  D'[j] = -D[C-1-j]  one has
//   idx'(w') = C-1-idx(w)   (idx(w) >= k <=> w >= T[k] <=> w' <= -T[k] <=> not (w' >= nextup(-T[k])))
//   D'[idx'(w')] - w' = -(D[idx(w)] - w)                       -- the same square, bit for bit.
// Both classes therefore run the same code: the index only ever steps DOWN along an ascending grid,
// the move test is w' < T*[g][k], and the value sits at a compile-time offset from the threshold.
// PITCH = 9 (C <= 8) or 17 (C <= 16); region 1 starts 16 banks away from region 0, so that for
// C <= 8 the two classes of one warp never collide on a bank.
struct TabHead {
  float ea[TAB_MAXG], eb[TAB_MAXG];     // index estimate t_a = w*ea + eb  (eb already holds the -0.5 bias)
  float scale[TAB_MAXG], rs[TAB_MAXG];  // rs = RN(1/scale): scaling.py:80's factor AND the reciprocal for dividing by scale
  float rr[TAB_MAXG];                   // RN(1/rs): reciprocal for dividing by rs
  int ok[TAB_MAXG];                     // bit 0 / 1: the exact-reciprocal scheme applies to scale / rs (common.cuh)
  float errs[TAB_THREADS / 32][TAB_MAXG];
  float X[TAB_MAXC];              // X[k] = min{x : slot(x) >= k}: breakpoints of the codebook itself
  float red_lo[4], red_hi[4];
  float init;
  int bad;                         // some grid point of this row cannot use the estimate -> direct form
  int nowalk;                      // the walk conditions do not hold for this row -> estimate loop
  int pad[5];
};
static_assert(sizeof(TabHead) % 128 == 0, "table regions must start on a bank-0 boundary");

template <int PITCH>
__host__ __device__ constexpr int tab_region_floats(int G) {
  // G rows of 2*PITCH floats, rounded so that the next region starts 16 banks further
  return ((G * 2 * PITCH + 31) / 32) * 32 + 16;
}
template <int PITCH>
static size_t tab_smem_bytes(int G) { return sizeof(TabHead) + (size_t)2 * tab_region_floats<PITCH>(G) * sizeof(float); }

// Sum over the warp of 16 values per lane by recursive halving: returns, in every lane, the total of
// v[lane & 15] over the 32 lanes.  Fixed order: deterministic.
__device__ __forceinline__ float warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, h));
    }
  }
  return __fadd_rn(v[0], __shfl_xor_sync(0xffffffffu, v[0], 16));
}

template <bool HAS_H, int TAB_WPT, int PITCH, bool PACK2 = true>
__global__ void __launch_bounds__(TAB_THREADS, 8) scale_search_tab_kernel(const float* __restrict__ w, int64_t r, int64_t n,
                                                                          DevGrid<float> g, GridBreaks brk, float cb_min,
                                                                          float cb_max, const float* __restrict__ factors,
                                                                          int G, const float* __restrict__ hdiag,
                                                                          float* __restrict__ out_scale,
                                                                          float* __restrict__ out_err,
                                                                          float* __restrict__ out_init, int walk_allowed) {
  extern __shared__ __align__(128) unsigned char tab_raw[];
  TabHead& sm = *reinterpret_cast<TabHead*>(tab_raw);
  constexpr int RP = 2 * PITCH;                         // floats per table row
  float* const tab0 = reinterpret_cast<float*>(tab_raw + sizeof(TabHead));
  float* const tab1 = tab0 + tab_region_floats<PITCH>(G);
  auto T0 = [&](int gi, int k) -> float& { return tab0[gi * RP + k]; };
  auto D0 = [&](int gi, int k) -> float& { return tab0[gi * RP + PITCH + k]; };
  auto T1 = [&](int gi, int k) -> float& { return tab1[gi * RP + k]; };
  auto D1 = [&](int gi, int k) -> float& { return tab1[gi * RP + PITCH + k]; };
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = g.size;
  const FastDivF fstep = make_fastdiv(g.step);
  const float pinf = __int_as_float(0x7f800000), ninf = __int_as_float(0xff800000);

  // breakpoints of the codebook in the scaled domain (exact; found on the host, make_breaks)
  if (tid < TAB_MAXC) sm.X[tid] = brk.X[tid];
  __syncthreads();

  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    const float* grow = w + row * n;
    // ---- pass 1: range of the row -> non-saturating scale (scaling.py:44-55) ---------------------
    float lo = __ldg(grow), hi = lo;
    for (int64_t j = tid; j < n; j += TAB_THREADS) {
      const float v = __ldg(grow + j);
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if (lane == 0) { sm.red_lo[warp] = lo; sm.red_hi[warp] = hi; }
    if (tid == 0) { sm.bad = 0; sm.nowalk = walk_allowed ? 0 : 1; }
    __syncthreads();
    lo = fminf(fminf(sm.red_lo[0], sm.red_lo[1]), fminf(sm.red_lo[2], sm.red_lo[3]));
    hi = fmaxf(fmaxf(sm.red_hi[0], sm.red_hi[1]), fmaxf(sm.red_hi[2], sm.red_hi[3]));
    float init;
    {
      const float a = __fdiv_rn(hi, cb_max), b = __fdiv_rn(lo, cb_min);  // scaling.py:53
      const float s = a > b ? a : b;
      init = s > 1.0e-16f ? s : 1.0e-16f;                                // scaling.py:54
    }
    const float wmax = fmaxf(fabsf(lo), fabsf(hi));
    // ---- pass 2a: per grid point constants -----------------------------------------------------------
    for (int gi = tid; gi < G; gi += TAB_THREADS) {
      const float scale = __fmul_rn(__ldg(factors + gi), init);          // scaling.py:128
      const float rsv = __fdiv_rn(1.0f, scale);                          // scaling.py:80
      sm.scale[gi] = scale;
      sm.rs[gi] = rsv;
      {
        const FastDivF fs = make_fastdiv(scale), fr = make_fastdiv(rsv);   // fs.y == rsv
        sm.rr[gi] = fr.y;
        sm.ok[gi] = fs.ok | (fr.ok << 1);
      }
      const float ea = __fdiv_rn(1.0f, __fmul_rn(scale, g.step));
      const float eb = __fsub_rn(__fdiv_rn(-g.zero, g.step), 0.5f);
      sm.ea[gi] = ea;
      sm.eb[gi] = eb;
      const float reach = __fadd_rn(__fmul_rn(wmax, fabsf(ea)), fabsf(eb));
      if (!(scale > 0.0f) || !(reach < 2.0e6f)) sm.bad = 1;             // also catches NaN / inf
      T0(gi, C) = pinf; T0(gi, 0) = ninf;
      T1(gi, C) = pinf; T1(gi, 0) = ninf;
      for (int wp = 0; wp < TAB_THREADS / 32; ++wp) sm.errs[wp][gi] = 0.0f;
    }
    __syncthreads();
    const bool direct = sm.bad != 0;
    if (direct) {
      // rare: the direct op chain, one grid point per thread (same code as scale_search_kernel)
      for (int gi = tid; gi < G; gi += TAB_THREADS) {
        const FastDivF fs = make_fastdiv(sm.scale[gi]), fr = make_fastdiv(sm.rs[gi]);
        sm.errs[0][gi] = (fstep.ok && fs.ok && fr.ok)
                             ? row_error<float, HAS_H, true>(grow, n, g, fs, fr, fstep, hdiag)
                             : row_error<float, HAS_H, false>(grow, n, g, fs, fr, fstep, hdiag);
      }
    } else {
      // ---- pass 2b: exact tables, one (grid point, code) task at a time ----------------------------
      // idx(w) >= k  <=>  RN(w/scale) >= X[k], so T[k] = min{w : RN(w/scale) >= X[k]}: it lies within a
      // few ulps of X[k]*scale; bracket, verify, bisect (full range if the bracket ever fails).
      // Task (gi, k): k = 0 only de-quantises.  CP2 = 8 or 16 codes per grid point slot, so that the task
      // index splits by shifts; the code of a thread rotates with the round (k = 0 tasks are light).
      constexpr int CP2 = PITCH - 1, CSH = (CP2 == 8 ? 3 : 4);
      for (int t0 = 0; t0 < G * CP2; t0 += TAB_THREADS) {
        const int gi = (t0 + tid) >> CSH, k = (tid + (t0 / TAB_THREADS)) & (CP2 - 1);
        if (gi >= G || k >= C) continue;
        const float scale = sm.scale[gi], rsv = sm.rs[gi], rrv = sm.rr[gi];
        const int okb = sm.ok[gi];
        const float v = __fadd_rn(__fmul_rn((float)k, g.step), g.zero);                                 // codebook.py:63-64
        const float dqv = (okb & 2) ? fastdiv_core(v, rsv, rrv) : __fdiv_rn(v, rsv);                    // scaling.py:80
        D0(gi, k) = dqv;
        D1(gi, C - 1 - k) = -dqv;
        if (k == 0) continue;
        const float xk = sm.X[k];
        auto reaches = [&](int o) -> bool {
          const float x0 = f32_unord(o);
          return ((okb & 1) ? fastdiv_core(x0, scale, rsv) : __fdiv_rn(x0, scale)) >= xk;               // scaling.py:73
        };
        // `reaches` is monotone in the ordered-integer key; the threshold lies within an ulp or two of
        // X[k]*scale: step from there (typically two evaluations), full bisection if that ever fails
        const int omin = f32_ord(-3.402823466e+38f), omax = f32_ord(3.402823466e+38f);
        int o = f32_ord(__fmul_rn(xk, scale));
        bool found = false;
        if (o > omin + 8 && o < omax - 8) {
          if (reaches(o)) {
            int steps = 0;
            while (steps < 6 && reaches(o - 1)) { --o; ++steps; }
            found = steps < 6;
          } else {
            int steps = 0;
            while (steps < 6 && !reaches(o + 1)) { ++o; ++steps; }
            ++o;
            found = steps < 6;
          }
        }
        if (!found) {
          long long blo = omin, bhi = omax;
          while (bhi - blo > 1) {
            const long long mid = blo + ((bhi - blo) >> 1);
            if (reaches((int)mid)) bhi = mid; else blo = mid;
          }
          o = (int)bhi;
        }
        const float tk = f32_unord(o);
        T0(gi, k) = tk;
        // mirrored threshold: the smallest float above -T[k] (in the ordered-integer domain -0 and +0
        // coincide, so the successor of the key of -T[k] is the next larger VALUE)
        const int om = f32_ord(-tk);
        T1(gi, C - k) = om == 0x7f7fffff ? pinf : f32_unord(om + 1);
      }
      __syncthreads();
      // ---- walk conditions, on the exact tables of both classes (idx_g(x) = max{j : T[g][j] <= x}) -----
      //   x >= 0:  idx_{g+1}(x) in {idx_g(x) - 1, idx_g(x)}     <=  (a) T[g][j] <= 0 or T[g+1][j] >= T[g][j]
      //                                                            (b) T[g+1][j-1] <= max(T[g][j], 0)
      if (!sm.nowalk) {
        for (int task = tid; task < 2 * (G - 1); task += TAB_THREADS) {
          const int gi = task >> 1;
          const float* tb = (task & 1) ? tab1 : tab0;
          bool ok = true;
          for (int j = 1; j < C; ++j) {
            const float tg = tb[gi * RP + j], tn = tb[(gi + 1) * RP + j];
            ok = ok && (tg <= 0.0f || tn >= tg) && (tb[(gi + 1) * RP + j - 1] <= fmaxf(tg, 0.0f));
          }
          if (!ok) sm.nowalk = 1;
        }
      }
      __syncthreads();
      if (!sm.nowalk) {
        // ---- pass 2c (walk form): weights and table pointers in registers ------------------------------
        // shared-memory BYTE addresses are tracked directly, so that the table loads are
        // LDS [register + immediate] with the grid point (and the value offset) folded into the immediate
        const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(tab0);
        const uint32_t base1 = (uint32_t)__cvta_generic_to_shared(tab1);
        auto lds = [](uint32_t addr) -> float {
          float v;
          asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));   // tables are read-only in this phase
          return v;
        };
        constexpr int GP = RP * 4;                            // bytes between consecutive grid points
        constexpr int DQ = PITCH * 4;                         // bytes from a threshold to its value
        for (int64_t c0 = 0; c0 < n; c0 += (int64_t)TAB_THREADS * TAB_WPT) {
          float wv[TAB_WPT], hv[TAB_WPT];
          uint32_t pt[TAB_WPT];         // &T*[g][k] of the weight's class at the block's first grid point
#pragma unroll
          for (int m = 0; m < TAB_WPT; ++m) {
            const int64_t j = c0 + tid + (int64_t)m * TAB_THREADS;
            const bool in = j < n;
            const float wj = in ? __ldg(grow + j) : 0.0f;
            hv[m] = in ? (HAS_H ? __ldg(hdiag + j) : 1.0f) : 0.0f;
            // exact index at the first grid point (estimate + one comparison, as in the estimate loop)
            const float ta = __fmaf_rn(wj, sm.ea[0], sm.eb[0]);
            int ka = __float_as_int(__fadd_rn(ta, 12582912.0f)) - 0x4B400000;
            ka = __vimin_s32_relu(ka, C - 1);
            const int kx = (wj >= T0(0, ka + 1)) ? ka + 1 : ka;
            const bool neg = wj < 0.0f;
            wv[m] = neg ? -wj : wj;
            pt[m] = neg ? base1 + 4u * (uint32_t)(C - 1 - kx) : base0 + 4u * (uint32_t)kx;
          }
          auto eval = [&](int m, int u, float acc) -> float {
            const float t = lds(pt[m] + u * GP);
            if (wv[m] < t) pt[m] -= 4;                                           // one step towards the centre
            const float e = __fsub_rn(lds(pt[m] + (u * GP + DQ)), wv[m]);        // scaling.py:130 (sign-mirrored for w < 0)
            return __fmaf_rn(__fmul_rn(e, e), hv[m], acc);
          };
          // packed form: two weights per arithmetic instruction (FADD2 / FMUL2 / FFMA2 on sm_100a), the
          // same per-lane IEEE operations; the thread's partial sum becomes (even weights) + (odd weights)
          auto pack2 = [](float lo_, float hi_) -> unsigned long long {
            unsigned long long v;
            asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo_), "f"(hi_));
            return v;
          };
          auto eval2 = [&](int m, int u, unsigned long long acc2) -> unsigned long long {
            const float t0 = lds(pt[m] + u * GP), t1 = lds(pt[m + 1] + u * GP);
            if (wv[m] < t0) pt[m] -= 4;
            if (wv[m + 1] < t1) pt[m + 1] -= 4;
            const unsigned long long d2 = pack2(lds(pt[m] + (u * GP + DQ)), lds(pt[m + 1] + (u * GP + DQ)));
            unsigned long long e2, s2, r2;
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(e2) : "l"(d2), "l"(pack2(wv[m], wv[m + 1])));
            asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(s2) : "l"(e2));
            asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(s2), "l"(pack2(hv[m], hv[m + 1])), "l"(acc2));
            return r2;
          };
          auto sum2 = [](unsigned long long v) -> float {
            float a_, b_;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a_), "=f"(b_) : "l"(v));
            return __fadd_rn(a_, b_);
          };
          int gb = 0;
          for (; gb + 16 <= G; gb += 16) {
            float part[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              if (PACK2) {
                unsigned long long acc2 = 0ull;
#pragma unroll
                for (int m = 0; m < TAB_WPT; m += 2) acc2 = eval2(m, u, acc2);
                part[u] = sum2(acc2);
              } else {
                float acc = 0.0f;
#pragma unroll
                for (int m = 0; m < TAB_WPT; ++m) acc = eval(m, u, acc);
                part[u] = acc;
              }
            }
#pragma unroll
            for (int m = 0; m < TAB_WPT; ++m) pt[m] += 16 * GP;
            const float tot = warp_sum16(part, lane);
            if (lane < 16) sm.errs[warp][gb + lane] += tot;
          }
          for (; gb < G; ++gb) {
            float acc = 0.0f;
#pragma unroll
            for (int m = 0; m < TAB_WPT; ++m) acc = eval(m, 0, acc);
#pragma unroll
            for (int m = 0; m < TAB_WPT; ++m) pt[m] += GP;
            acc = warp_sum(acc);
            if (lane == 0) sm.errs[warp][gb] += acc;
          }
        }
      } else
      // ---- pass 2c (estimate form): weights in registers, walk the grid points ------------------------
      for (int64_t c0 = 0; c0 < n; c0 += (int64_t)TAB_THREADS * TAB_WPT) {
        float wv[TAB_WPT], hv[TAB_WPT];
#pragma unroll
        for (int m = 0; m < TAB_WPT; ++m) {
          const int64_t j = c0 + tid + (int64_t)m * TAB_THREADS;
          const bool in = j < n;
          wv[m] = in ? __ldg(grow + j) : 0.0f;
          hv[m] = in ? (HAS_H ? __ldg(hdiag + j) : 1.0f) : 0.0f;
        }
#pragma unroll 2
        for (int gi = 0; gi < G; ++gi) {
          const float ea = sm.ea[gi], eb = sm.eb[gi];
          const float* Tg = &T0(gi, 0);
          const float* Dg = &D0(gi, 0);
          float part = 0.0f;
#pragma unroll
          for (int m = 0; m < TAB_WPT; ++m) {
            const float ta = __fmaf_rn(wv[m], ea, eb);
            const float u = __fadd_rn(ta, 12582912.0f);                    // integer part lands in the low mantissa bits
            int ka = __float_as_int(u) - 0x4B400000;
            ka = __vimin_s32_relu(ka, C - 1);                               // clip to [0, C-1]
            const int kx = (wv[m] >= Tg[ka + 1]) ? ka + 1 : ka;             // exact index
            const float e = __fsub_rn(Dg[kx], wv[m]);                       // scaling.py:130
            part = __fmaf_rn(__fmul_rn(e, e), hv[m], part);
          }
          part = warp_sum(part);
          if (lane == 0) sm.errs[warp][gi] += part;
        }
      }
    }
    __syncthreads();
    // ---- pass 3: first strict minimum in grid order, best kept in fp32 (scaling.py:125-134) --------
    if (warp == 0) {
      // every lane scans its grid points in ascending order (strict <), then the lanes are merged:
      // smaller error wins, equal errors -> smaller grid index = the first strict minimum in grid order
      float best = pinf;
      int bi = 0x7fffffff;
      for (int gi = lane; gi < G; gi += 32) {
        const float e = direct ? sm.errs[0][gi]
                               : __fadd_rn(__fadd_rn(sm.errs[0][gi], sm.errs[1][gi]), __fadd_rn(sm.errs[2][gi], sm.errs[3][gi]));
        if (e < best) { best = e; bi = gi; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float eo = __shfl_xor_sync(0xffffffffu, best, o);
        const int io = __shfl_xor_sync(0xffffffffu, bi, o);
        if (eo < best || (eo == best && io < bi)) { best = eo; bi = io; }
      }
      if (lane == 0) {
        const float pick = bi < G ? __ldg(factors + bi) : pinf;     // no finite error at all: inf, as the serial scan
        out_scale[row] = __fmul_rn(init, pick);
        if (out_err) out_err[row] = best;
        if (out_init) out_init[row] = init;
      }
    }
    __syncthreads();
  }
}

template <typename HT, bool HAS_H>
__global__ void __launch_bounds__(128) scale_search_kernel(const float* __restrict__ w, int64_t r, int64_t n,
                                                           DevGrid<float> g, float cb_min, float cb_max,
                                                           const float* __restrict__ factors, int G,
                                                           const HT* __restrict__ hdiag, int row_in_smem,
                                                           float* __restrict__ out_scale, float* __restrict__ out_err,
                                                           float* __restrict__ out_init) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HT* errs = (HT*)smem_raw;                                   // [G]
  float* srow = (float*)(smem_raw + (((size_t)G * sizeof(HT) + 15) & ~(size_t)15));  // [n] if row_in_smem
  __shared__ float red_lo[4], red_hi[4];
  __shared__ float s_init;
  const FastDivF fstep = make_fastdiv(g.kind == 0 ? g.step : 1.0f);

  for (int64_t row = blockIdx.x; row < r; row += gridDim.x) {
    const float* grow = w + row * n;
    // ---- pass 1: stage the row, find its range -------------------------------
    float lo = __ldg(grow), hi = lo;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
      float v = __ldg(grow + j);
      if (row_in_smem) srow[j] = v;
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { red_lo[threadIdx.x >> 5] = lo; red_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
        lo = red_lo[k] < lo ? red_lo[k] : lo;
        hi = red_hi[k] > hi ? red_hi[k] : hi;
      }
      float a = __fdiv_rn(hi, cb_max), b = __fdiv_rn(lo, cb_min);  // scaling.py:53
      float s = a > b ? a : b;
      s_init = s > 1.0e-16f ? s : 1.0e-16f;                        // scaling.py:54
    }
    __syncthreads();
    const float init = s_init;
    const float* rowp = row_in_smem ? srow : grow;

    // ---- pass 2: one grid point per thread ------------------------------------
    for (int gi = threadIdx.x; gi < G; gi += blockDim.x) {
      const float scale = __fmul_rn(__ldg(factors + gi), init);    // scaling.py:128
      const float rs = __fdiv_rn(1.0f, scale);                     // scaling.py:80
      const FastDivF fs = make_fastdiv(scale), fr = make_fastdiv(rs);
      if (g.kind == 0 && fstep.ok && fs.ok && fr.ok)
        errs[gi] = row_error<HT, HAS_H, true>(rowp, n, g, fs, fr, fstep, hdiag);
      else
        errs[gi] = row_error<HT, HAS_H, false>(rowp, n, g, fs, fr, fstep, hdiag);
    }
    __syncthreads();
    // ---- pass 3: first strict minimum in grid order, best kept in fp32 (scaling.py:125-134)
    if (threadIdx.x == 0) {
      float best = __int_as_float(0x7f800000), pick = __int_as_float(0x7f800000);
      for (int gi = 0; gi < G; ++gi) {
        HT e = errs[gi];
        if (e < (HT)best) { best = (float)e; pick = __ldg(factors + gi); }
      }
      out_scale[row] = __fmul_rn(init, pick);
      if (out_err) out_err[row] = best;
      if (out_init) out_init[row] = init;
    }
    __syncthreads();
  }
}

}  // namespace slk

using namespace slk;

static int g_search_direct = -1;

/* development aid / tests: 1 forces the direct op-chain kernel, 2 the threshold tables with the
   estimate loop (no walk), 0 the default (threshold tables, walk form) */
extern "C" int slk_debug_scale_search_direct(int on) {
  g_search_direct = on == 2 ? 2 : (on ? 1 : 0);
  return SLK_OK;
}

extern "C" int slk_scale_search_f32(const float* w, int64_t r, int64_t n, const slk_codebook* cb,
                                    const float* factors, int32_t G, const void* hdiag, int32_t h_dtype,
                                    float* out_scale, float* out_err, float* out_init, void* stream) {
  int rc = check_codebook(cb);
  if (rc) return rc;
  SLK_REQUIRE(r >= 0 && n >= 1 && G >= 1, "bad shape r=%lld n=%lld G=%d", (long long)r, (long long)n, G);
  SLK_REQUIRE(cb->lo < 0 && cb->hi > 0, "Codebook should have both negative and positive values.");
  SLK_REQUIRE(h_dtype >= 0 && h_dtype <= 2 && (h_dtype == 0) == (hdiag == nullptr), "hdiag / h_dtype mismatch");
  if (r == 0) return SLK_OK;
  SLK_REQUIRE(w && factors && out_scale, "NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esz = h_dtype == 2 ? 8 : 4;
  const size_t err_bytes = ((size_t)G * esz + 15) & ~(size_t)15;
  size_t smem = err_bytes + (size_t)n * 4;
  int row_in_smem = 1;
  if (smem > 200 * 1024) { smem = err_bytes; row_in_smem = 0; }
  SLK_REQUIRE(smem <= 200 * 1024, "grid of %d points does not fit in shared memory", G);
  const int grid = (int)(r < (int64_t)sm_count() * 16 ? r : (int64_t)sm_count() * 16);
  DevGrid<float> g = make_grid<float>(cb);
  const float cmin = (float)cb->lo, cmax = (float)cb->hi;
  if (g_search_direct < 0) {   // SLK_SEARCH_TABLE=0 forces the direct op chain (A/B testing)
    const char* ev = getenv("SLK_SEARCH_TABLE");
    g_search_direct = (ev && ev[0] == '0') ? 1 : 0;
  }
  const int steps_off = g_search_direct == 1;
  const int walk_allowed = g_search_direct == 0;
  if (!steps_off && cb->kind == 0 && cb->size <= TAB_MAXC && G <= TAB_MAXG && h_dtype != 2) {
    // threshold-table form; grid: 1.5x the CTAs that are resident at once (8 per SM: 64 registers x 128
    // threads, ~21 KB of tables), each looping over rows, so that one CTA's latency-bound table
    // construction overlaps the others' throughput-bound main loops (measured on [3072,768]: 88.9 us
    // with 12 CTAs per SM, 92.1 us with 8)
    const int tgrid = (int)(r < (int64_t)sm_count() * 12 ? r : (int64_t)sm_count() * 12);
    const GridBreaks brk = make_breaks(cb);
    static int pack2 = -1;   // SLK_SEARCH_PACK2=0: scalar walk loop instead of the packed f32x2 one (A/B testing)
    if (pack2 < 0) {
      const char* ev = getenv("SLK_SEARCH_PACK2");
      pack2 = (ev && ev[0] == '0') ? 0 : 1;
    }
    // weights per thread and chunk: the candidate with the least padding
    int wpt = 8;
    {
      int64_t best_waste = -1;
      for (int cand : {8, 6, 4}) {
        const int64_t span = (int64_t)TAB_THREADS * cand;
        const int64_t waste = ceil_div(n, span) * span - n;
        if (best_waste < 0 || waste < best_waste) { best_waste = waste; wpt = cand; }
      }
    }
#define SLK_LAUNCH_TAB3(HAS, WPT, PITCH, P2)                                                                         \
    do {                                                                                                             \
      const size_t tb = tab_smem_bytes<PITCH>(G);                                                                    \
      if (tb > 48 * 1024)                                                                                            \
        SLK_CUDA(cudaFuncSetAttribute(scale_search_tab_kernel<HAS, WPT, PITCH, P2>,                                  \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb));                        \
      scale_search_tab_kernel<HAS, WPT, PITCH, P2><<<tgrid, TAB_THREADS, tb, st>>>(                                  \
          w, r, n, g, brk, cmin, cmax, factors, G, (const float*)hdiag, out_scale, out_err, out_init, walk_allowed); \
    } while (0)
#define SLK_LAUNCH_TAB2(HAS, WPT, PITCH)                                                                             \
    do { if (pack2) SLK_LAUNCH_TAB3(HAS, WPT, PITCH, true); else SLK_LAUNCH_TAB3(HAS, WPT, PITCH, false); } while (0)
#define SLK_LAUNCH_TAB(HAS, WPT)                                                                                     \
    do { if (cb->size <= 8) SLK_LAUNCH_TAB2(HAS, WPT, 9); else SLK_LAUNCH_TAB2(HAS, WPT, 17); } while (0)
    if (h_dtype == 1) {
      if (wpt == 8) SLK_LAUNCH_TAB(true, 8); else if (wpt == 6) SLK_LAUNCH_TAB(true, 6); else SLK_LAUNCH_TAB(true, 4);
    } else {
      if (wpt == 8) SLK_LAUNCH_TAB(false, 8); else if (wpt == 6) SLK_LAUNCH_TAB(false, 6); else SLK_LAUNCH_TAB(false, 4);
    }
#undef SLK_LAUNCH_TAB2
#undef SLK_LAUNCH_TAB3
#undef SLK_LAUNCH_TAB
    SLK_LAUNCH_CHECK();
    return SLK_OK;
  }
#define SLK_LAUNCH_SS(HT, HAS)                                                                           \
  do {                                                                                                   \
    if (smem > 48 * 1024)                                                                                \
      SLK_CUDA(cudaFuncSetAttribute(scale_search_kernel<HT, HAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    scale_search_kernel<HT, HAS><<<grid, 128, smem, st>>>(w, r, n, g, cmin, cmax, factors, G, (const HT*)hdiag, \
                                                           row_in_smem, out_scale, out_err, out_init);   \
  } while (0)
  if (h_dtype == 0) SLK_LAUNCH_SS(float, false);
  else if (h_dtype == 1) SLK_LAUNCH_SS(float, true);
  else SLK_LAUNCH_SS(double, true);
#undef SLK_LAUNCH_SS
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}
