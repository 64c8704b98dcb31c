// K2 (fused-path form): dampening + column permutation + fp64 Cholesky factor, WITHOUT the n^3/3
// triangular inverse.
//   H_opt = H + damp*mean(diag H)*I ; H_opt[order][:, order]                  obq.py:198-204
//   compute_hessian_chol builds U = flip(inv(cholesky(flip(H_opt))))            obq.py:38-55
// With L = cholesky(flip(H_opt)) and R = flip(L) (upper, H_opt = R R^T) one has U = R^-1, and the
// sweep's propagated term is  E[:, :a] U[:a, J] = -(W-Q)[:, :a] R[:a, J] U_JJ  with U_JJ the inverse
// of the 32x32 diagonal block R_JJ (verified against the reference's own loop).  The
// sweep (sweep.cu, R form) therefore needs only R (fp32) and the diagonal-block inverses; the
// public compute_hessian_chol keeps the full inverse (hinv.cu).
//
// The factorisation is ONE launch: a tile Cholesky (64x64 tiles, left-looking) whose tile tasks are
// handed out by an atomic ticket in column-major order.  Task (i, j), i >= j:
//     S = A(i,j) - sum_{k<j} L(i,k) L(j,k)^T        DMMA m8n8k4, operands through a cp.async ring
//     i == j:  L(j,j) = chol(S), X_j = L(j,j)^-1     in shared memory (8-column register panels,
//                                                     DMMA trailing updates, inverse by doubling)
//     i >  j:  L(i,j) = S X_j^T                      DMMA
// and publishes a per-tile ready flag (release); consumers poll the flags of the tiles they read
// (acquire).  A task only ever waits for tasks with a SMALLER ticket, and a ticket is only taken by
// a CTA that is already running, so the scheme cannot deadlock whatever else shares the GPU and
// however few CTAs are resident (the decoupled look-back argument).  No dependent launches: the
// n/64-panel launch chain of the right-looking version (hinv.cu) becomes in-kernel flag latency.
//
// Two generalisations of the same kernel (round 2):
//  * BATCH (slk_chol_factor_batched_f32): several matrices of one size share ONE ticket queue, ticket t
//    -> (matrix t % B, task t / B).  A CTA whose next tile is not ready is then almost always handed a
//    tile of another matrix instead of spinning: the 72 layers of an OPT-125M pass are bound by the
//    FP64 pipe instead of by 72 separate chains of diagonal tiles that each hold SM slots while they wait.
//  * MULTI-GPU (slk_chol_dist_*): tile rows are dealt to the ranks block-cyclically; a rank runs the tile
//    tasks of its own rows in the same column-major order and PUSHES every finished tile (and the
//    inverse of a diagonal tile) into the same place of every peer's workspace through NVLink peer
//    stores, followed by a system-scope release of the tile's flag on that peer.  Consumers only ever
//    read local memory and local flags.  No collective inside the factorisation, the transfer of
//    tile (i, j) overlaps the arithmetic of every other tile, and the deadlock argument above carries
//    over (a task still only waits for tasks that precede it in the global column-major order, and
//    every rank works through its own tasks in that order).
#include "common.cuh"

#include <stdlib.h>
#include "tc_gemm.cuh"

namespace slk {

constexpr int CT = 64;            // tile edge
constexpr int CP = 68;            // shared-memory pitch (doubles): 68 = 4 (mod 16) makes every DMMA fragment
                                  // access touch each 8-byte bank pair at most twice (the minimum)
constexpr int CBK = 16;           // k per ring stage
constexpr int CLD = CBK + 4;      // ring pitch
constexpr int CST = 4;            // ring stages
constexpr int CTH = 256;          // threads per CTA (8 warps: 2 x 4 warp tiles of 32 x 16)

struct CholSmem {
  union {
    struct { double A[CST][CT * CLD]; double B[CST][CT * CLD]; } ring;          // 80 KB
    struct { double M[CT * CP]; double X[CT * CP]; double T[32 * CP]; } f;      // 85 KB
  };
  double rd[CT];                  // 1 / L[c][c]
  int ticket;
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <bool MULTI>
__device__ __forceinline__ void wait_ready(const int* flag) {
  if (MULTI) { while (ld_acquire_sys(flag) == 0) __nanosleep(64); }
  else { while (ld_acquire_gpu(flag) == 0) __nanosleep(32); }
}

constexpr int CHOL_MAXJ = 64;     // matrices per batched launch
constexpr int CHOL_MAXP = 8;      // ranks of a distributed factorisation

// One launch: `njobs` matrices of T x T tiles (workspaces laid out as chol_ws_layout says), or ONE
// matrix shared by `nranks` GPUs (peer[q] = base of rank q's workspace, same layout, peer-mapped).
struct CholDagParams {
  int njobs, T;
  int nranks, rank;
  int rowblock;          // multi-GPU: tile rows are dealt to the ranks in blocks of `rowblock` consecutive rows
  int64_t ld;
  int* ticket;
  double* A[CHOL_MAXJ];
  double* Dinv[CHOL_MAXJ];
  int* ready[CHOL_MAXJ];
  int32_t* info[CHOL_MAXJ];
  double* peerA[CHOL_MAXP];
  double* peerDinv[CHOL_MAXP];
  int* peerReady[CHOL_MAXP];
};
__device__ __forceinline__ void cd_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cd_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// A[i, j] (tiles on and below the diagonal) of flip(H_opt[order][:, order]) with the identity padding in FRONT:
// flipped index i <-> sweep column npad-1-i, so that 32-column sweep blocks coincide with 32-row
// blocks of the factor for every n.  Also clears the flags, the ticket and info.
template <typename TS>
struct CholGatherParams {
  const TS* h[CHOL_MAXJ];
  const int64_t* order[CHOL_MAXJ];
  const float* dampval[CHOL_MAXJ];
  double* A[CHOL_MAXJ];
  int* flags[CHOL_MAXJ];
  int32_t* info[CHOL_MAXJ];
};

template <typename TS>
__global__ void __launch_bounds__(256) chol_gather_kernel(const __grid_constant__ CholGatherParams<TS> P, int64_t n,
                                                          int64_t npad, int64_t nflags) {
  const int job = blockIdx.y;
  const TS* __restrict__ h = P.h[job];
  const int64_t* __restrict__ order = P.order[job];
  double* __restrict__ A = P.A[job];
  int* __restrict__ flags = P.flags[job];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t t = t0; t < nflags; t += stride) flags[t] = 0;
  if (t0 == 0) *P.info[job] = 0;
  const double damp = P.dampval[job] ? (double)P.dampval[job][0] : 0.0;
  const int64_t total = npad * npad;
  for (int64_t t = t0; t < total; t += stride) {
    const int64_t i = t / npad, j = t - i * npad;
    if (j / CT > i / CT) continue;             // tiles above the diagonal are never read
    const int64_t a = npad - 1 - i, b = npad - 1 - j;
    double v;
    if (a < n && b < n) {
      const int64_t sa = order ? __ldg(order + a) : a, sb = order ? __ldg(order + b) : b;
      v = (double)h[sa * n + sb];
      if (i == j) v = __dadd_rn(v, damp);      // obq.py:198: fp32 H promoted, fp64 add
    } else {
      v = (i == j) ? 1.0 : 0.0;
    }
    A[t] = v;
  }
}

// ---- 64x64 diagonal tile: factor + inverse in shared memory ---------------------------------------
// M holds the lower triangle column-major (M[c*CP + r] = S[r][c], r >= c) so that a lane that owns
// a row reads consecutive addresses.

// Panel of 8 columns, one warp.  Every lane keeps the 8x8 diagonal block in registers and factors
// it redundantly, so the dependent chain per column is rsqrt -> multiply -> fma with no shuffle
// on it; the lane's own two rows (lane, lane+32) are carried along right-looking with the block's
// multipliers taken from those registers.
__device__ __forceinline__ void chol_panel8(double* M, double* rd, int b, int lane, int32_t* info, int64_t gcol0) {
  const int c0 = 8 * b;
  double dg[8][8];          // dg[c][r] = block(r, c), r >= c
  double v0[8], v1[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
#pragma unroll
    for (int r = c; r < 8; ++r) dg[c][r] = M[(c0 + c) * CP + c0 + r];
    v0[c] = M[(c0 + c) * CP + lane];
    v1[c] = M[(c0 + c) * CP + lane + 32];
  }
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double piv = dg[j][j];
    bad = bad || !(piv > 0.0);
    const double y = rsqrt(piv);
#pragma unroll
    for (int r = j + 1; r < 8; ++r) dg[j][r] = __dmul_rn(dg[j][r], y);        // L(r, j)
#pragma unroll
    for (int c = j + 1; c < 8; ++c)
#pragma unroll
      for (int r = c; r < 8; ++r) dg[c][r] = __fma_rn(-dg[j][r], dg[j][c], dg[c][r]);
    v0[j] = __dmul_rn(v0[j], y);                 // the pivot row's own entry becomes piv*y = sqrt(piv)
    v1[j] = __dmul_rn(v1[j], y);
#pragma unroll
    for (int c = j + 1; c < 8; ++c) {
      v0[c] = __fma_rn(-v0[j], dg[j][c], v0[c]);
      v1[c] = __fma_rn(-v1[j], dg[j][c], v1[c]);
    }
    if (lane == j) rd[c0 + j] = y;
  }
  if (lane == 0 && bad) atomicCAS(info, 0, (int32_t)(gcol0 + c0 + 1));
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    M[(c0 + c) * CP + lane] = v0[c];
    M[(c0 + c) * CP + lane + 32] = v1[c];
  }
}

// Rank-8 update of one 8x8 tile (ti, tk), b < tk <= ti, by the panel b just factored (two DMMAs).
__device__ __forceinline__ void chol_trailing8_tile(double* M, int abase, int ti, int tk, int fr, int fk) {
  double* cp = &M[(8 * tk + 2 * fk) * CP + 8 * ti + fr];
  double c[2] = {cp[0], cp[CP]};
  const double a0 = M[abase + 8 * ti], a1 = M[abase + 8 * ti + 4 * CP];
  const double b0 = M[abase + 8 * tk], b1 = M[abase + 8 * tk + 4 * CP];
  cd_dmma(c, -a0, b0);
  cd_dmma(c, -a1, b1);
  cp[0] = c[0];
  cp[CP] = c[1];
}

// wr = rank of this warp among the 7 helper warps.  first: only the column block of panel b+1 (tile
// (b+1+wr, b+1)); else: the remaining tiles (tk >= b+2), dealt round-robin.
__device__ __forceinline__ void chol_trailing8_part(double* M, int b, bool first, int wr, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
  const int abase = (8 * b + fk) * CP + fr;
  if (first) {
    const int ti = b + 1 + wr;
    if (ti < 8) chol_trailing8_tile(M, abase, ti, b + 1, fr, fk);
    return;
  }
  int cm = 0;
  for (int ti = b + 2; ti < 8; ++ti)
    for (int tk = b + 2; tk <= ti; ++tk) {
      if (cm == wr) chol_trailing8_tile(M, abase, ti, tk, fr, fk);
      if (++cm == 7) cm = 0;
    }
}

// C(i, n) = alpha * sum_k A(i, k) B(k, n) on 8x8 DMMA tiles in shared memory, `nprob` independent
// problems of size s x s x s (s = 8 << lg) laid out with the given strides.  Tiles are dealt to the
// 8 warps; a warp that owns two tiles runs them interleaved (two independent accumulator chains).
__device__ __forceinline__ void smem_mm_batched(double* C, int sci, int scn, int spc, const double* A, int sai, int sak,
                                                int spa, const double* B, int sbk, int sbn, int spb, int lg, int nprob,
                                                double alpha, int warp, int lane) {
  const int fr = lane >> 2, fk = lane & 3;
  const int s = 8 << lg, tmask = (1 << lg) - 1;          // tiles per side = 1 << lg
  const int total = nprob << (2 * lg);
  auto setup = [&](int t, const double*& Ap, const double*& Bp, double*& Cp) {
    const int p = t >> (2 * lg), tt = t & ((1 << (2 * lg)) - 1);
    const int i0 = (tt >> lg) * 8, n0 = (tt & tmask) * 8;
    Ap = A + p * spa + (i0 + fr) * sai + fk * sak;
    Bp = B + p * spb + fk * sbk + (n0 + fr) * sbn;
    Cp = C + p * spc + (i0 + fr) * sci + (n0 + 2 * fk) * scn;
  };
  for (int t = warp; t < total; t += 2 * (CTH / 32)) {
    const int t2 = t + CTH / 32;
    const bool two = t2 < total;
    const double *A0, *B0, *A1, *B1;
    double *C0, *C1;
    setup(t, A0, B0, C0);
    setup(two ? t2 : t, A1, B1, C1);
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
#pragma unroll 4
    for (int k0 = 0; k0 < s; k0 += 4) {
      const double a0 = A0[k0 * sak], b0 = B0[k0 * sbk];
      const double a1 = A1[k0 * sak], b1 = B1[k0 * sbk];
      cd_dmma(c0, a0, b0);
      cd_dmma(c1, a1, b1);
    }
    C0[0] = __dmul_rn(alpha, c0[0]);
    C0[scn] = __dmul_rn(alpha, c0[1]);
    if (two) {
      C1[0] = __dmul_rn(alpha, c1[0]);
      C1[scn] = __dmul_rn(alpha, c1[1]);
    }
  }
}

// In: sm.f.M lower triangle (column-major).  Out: L in sm.f.M (same layout, entries with r < c are
// garbage), X = L^-1 row-major in sm.f.X (upper part zero).  All 256 threads.
__device__ __forceinline__ void chol_factor_tile(CholSmem& sm, int tid, int32_t* info, int64_t gcol0,
                                                 long long* tr = nullptr) {
  const int warp = tid >> 5, lane = tid & 31;
  double* M = sm.f.M;
  double* X = sm.f.X;
  for (int t = tid; t < CT * CP; t += CTH) X[t] = 0.0;
  long long tp = 0, tt = 0, c0 = tr ? clock64() : 0;
  // Look-ahead schedule: after panel b, warps 1..7 first update only the column block of panel b+1
  // (one tile each), then warp 0 factors panel b+1 while warps 1..7 apply panel b to the remaining
  // tiles -- the bulk of the trailing update runs beside the rsqrt chain instead of after it.
  if (warp == 0) chol_panel8(M, sm.rd, 0, lane, info, gcol0);
  __syncthreads();
  for (int b = 0; b < 7; ++b) {
    if (warp > 0) chol_trailing8_part(M, b, true, warp - 1, lane);
    __syncthreads();
    if (tr) { const long long c1 = clock64(); tt += c1 - c0; c0 = c1; }
    if (warp == 0) chol_panel8(M, sm.rd, b + 1, lane, info, gcol0);
    else if (b < 6) chol_trailing8_part(M, b, false, warp - 1, lane);
    __syncthreads();
    if (tr) { const long long c1 = clock64(); tp += c1 - c0; c0 = c1; }
  }
  if (tr) { tr[8] = tp; tr[9] = tt; tr[10] = clock64(); }
  // inverses of the 8x8 diagonal blocks: thread c solves L_bb x = e_cc by forward substitution
  if (tid < CT) {
    const int b8 = tid & ~7, cc = tid & 7;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double s = (i == cc) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; ++k) s = __fma_rn(-M[(b8 + k) * CP + b8 + i], x[k], s);
      x[i] = (i >= cc) ? __dmul_rn(s, sm.rd[b8 + i]) : 0.0;
      X[(b8 + i) * CP + b8 + cc] = x[i];
    }
  }
  __syncthreads();
  if (tr) tr[11] = clock64();
  // doubling: inv([[A,0],[C,B]]) = [[Ai,0],[-Bi C Ai, Bi]]
  for (int lg = 0; lg < 3; ++lg) {
    const int s = 8 << lg, pairs = CT / (2 * s);
    // T_p = L21 X11
    smem_mm_batched(sm.f.T, CP, 1, s * CP, M + s, 1, CP, 2 * s * (CP + 1), X, CP, 1, 2 * s * (CP + 1), lg, pairs, 1.0,
                    warp, lane);
    __syncthreads();
    // X21 = -X22 T_p
    smem_mm_batched(X + s * CP, CP, 1, 2 * s * (CP + 1), X + s * (CP + 1), CP, 1, 2 * s * (CP + 1), sm.f.T, CP, 1, s * CP,
                    lg, pairs, -1.0, warp, lane);
    __syncthreads();
  }
  if (tr) tr[12] = clock64();
}

// Optional per-task trace (development aid, tools/chol_trace.py): 16 int64 per task --
// i, j, globaltimer at start / end, clock64 after the ticket / k loop / tile math / publish.
__device__ long long* g_chol_trace = nullptr;
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---- the tile-task kernel ----------------------------------------------------------------------------
// Per matrix: A [T*64, ld] fp64, lower triangle in, L out; Dinv [T, 64, 64] inverses of the diagonal
// tiles; ready [T*T] flags.  P.ticket: the launch's ticket counter.
// Ticket t -> matrix t % njobs, local ticket t / njobs; local tickets enumerate, column by column, the
// tiles (i, j), i >= j, whose row this rank owns (i % nranks == rank; every row when nranks == 1).
template <bool MULTI>
__global__ void __launch_bounds__(CTH, 2) chol_dag_kernel(const __grid_constant__ CholDagParams P) {
  extern __shared__ __align__(16) unsigned char chol_raw[];
  CholSmem& sm = *reinterpret_cast<CholSmem*>(chol_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fk = lane & 3;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
  const int T = P.T;
  const int64_t ld = P.ld;
  const int NR = MULTI ? P.nranks : 1, RK = MULTI ? P.rank : 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) sm.ticket = atomicAdd(P.ticket, 1);
    __syncthreads();
    const int tk = sm.ticket;
    const int job = MULTI ? 0 : tk % P.njobs;
    int rem = MULTI ? tk : tk / P.njobs;
    // Rows are dealt block-cyclically: row i belongs to rank (i / RB) % NR (RB = 1, NR = 1 on one GPU).  The chain
    // diag(j) -> solve(j+1, j) -> diag(j+1) then stays on one GPU for RB columns before it crosses NVLink.
    // below(x) = owned rows < x; the m-th owned row is (m / RB) * RB * NR + RK * RB + m % RB.
    const int RB = MULTI ? P.rowblock : 1;
    auto below = [&](int x) {
      const int per = RB * NR, r_ = x % per - RK * RB;
      return (x / per) * RB + (r_ < 0 ? 0 : (r_ > RB ? RB : r_));
    };
    const int owned = below(T);
    int j = 0, i = 0;
    for (;; ++j) {
      if (j >= T) return;                                   // past this rank's last task
      const int b = below(j), cnt = owned - b;              // owned rows at or below the diagonal of column j
      if (rem < cnt) {
        const int m = b + rem;
        i = (m / RB) * RB * NR + RK * RB + m % RB;
        break;
      }
      rem -= cnt;
    }
    const bool diag = (i == j);
    double* __restrict__ A = P.A[job];
    double* __restrict__ Dinv = P.Dinv[job];
    int* ready = P.ready[job];
    long long* tr = (g_chol_trace && tid == 0) ? g_chol_trace + (int64_t)tk * 16 : nullptr;
    if (tr) { tr[0] = i; tr[1] = j; tr[2] = gtimer(); tr[4] = clock64(); }
    const double* Ai = A + (int64_t)i * CT * ld;
    const double* Aj = A + (int64_t)j * CT * ld;
    const int64_t tile_off = (int64_t)i * CT * ld + (int64_t)j * CT;
    double* Cij = A + tile_off;

    double acc[4][2][2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) acc[ii][jj][0] = acc[ii][jj][1] = 0.0;

    const int nt = 4 * j;   // ring steps: 64-wide k tiles split in 4
    auto load = [&](int s) {
      const int st = s % CST;
      const int64_t k0 = (int64_t)(s >> 2) * CT + (s & 3) * CBK;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * CTH, row = idx >> 3, piece = idx & 7;
        cd_cp16(&sm.ring.A[st][row * CLD + piece * 2], Ai + (int64_t)row * ld + k0 + piece * 2);
        if (!diag) cd_cp16(&sm.ring.B[st][row * CLD + piece * 2], Aj + (int64_t)row * ld + k0 + piece * 2);
      }
    };
    if (nt > 0) {
      if (tid == 0) {
        wait_ready<MULTI>(ready + (int64_t)i * T);
        if (!diag) wait_ready<MULTI>(ready + (int64_t)j * T);
      }
      __syncthreads();
    }
    for (int s = 0; s < CST - 1; ++s) {
      if (s < nt) load(s);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int s = 0; s < nt; ++s) {
      const int sn = s + CST - 1;
      if (tid == 0 && sn < nt && (sn & 3) == 0) {
        wait_ready<MULTI>(ready + (int64_t)i * T + (sn >> 2));
        if (!diag) wait_ready<MULTI>(ready + (int64_t)j * T + (sn >> 2));
      }
      asm volatile("cp.async.wait_group %0;" ::"n"(CST - 2) : "memory");
      __syncthreads();
      if (sn < nt) load(sn);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const double* As = sm.ring.A[s % CST];
      const double* Bs = diag ? As : sm.ring.B[s % CST];
#pragma unroll
      for (int ks = 0; ks < CBK; ks += 4) {
        double a[4], b[2];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) a[ii] = As[(wm + 8 * ii + fr) * CLD + ks + fk];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) b[jj] = Bs[(wn + 8 * jj + fr) * CLD + ks + fk];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) cd_dmma(acc[ii][jj], a[ii], b[jj]);
      }
    }
    // the tile itself, in accumulator-fragment layout (rows wm+8ii+fr, columns wn+8jj+2fk, +1): loaded AFTER the
    // k loop (nobody else writes A(i,j) before this task publishes it), so that its 32 registers are free for
    // the loop's operand prefetch; the load latency overlaps the ring's last wait and the barrier
    double2 cij[4][2];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int row = wm + 8 * ii + fr, col = wn + 8 * jj + 2 * fk;
        cij[ii][jj] = *reinterpret_cast<const double2*>(Cij + (int64_t)row * ld + col);
      }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // the ring is dead: its storage becomes M / X / T
    if (tr) tr[5] = clock64();

    if (diag) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int row = wm + 8 * ii + fr, col = wn + 8 * jj + 2 * fk;
          sm.f.M[col * CP + row] = __dsub_rn(cij[ii][jj].x, acc[ii][jj][0]);
          sm.f.M[(col + 1) * CP + row] = __dsub_rn(cij[ii][jj].y, acc[ii][jj][1]);
        }
      __syncthreads();
      chol_factor_tile(sm, tid, P.info[job], (int64_t)j * CT, tr);
      // X_j is what the panel solves below wait for: publish it first; L(j,j) itself is only read
      // by the export kernel after this launch
      const int64_t xoff = (int64_t)j * CT * CT;
      if (MULTI) {
        // Local copy and local flag first (gpu-scope release: the solves of this GPU's own rows go on at once),
        // then the peers' copies and their flags.  The CTA barrier orders every thread's stores before the
        // flag writers' release stores, and a release is cumulative (the pattern of a grid barrier: one
        // thread fences for the CTA), so no per-thread system fence sits on the chain of diagonal tiles.
        for (int e = tid; e < CT * CT / 2; e += CTH) {
          const int r = e >> 5, c2 = (e & 31) * 2;
          *reinterpret_cast<double2*>(Dinv + xoff + r * CT + c2) = *reinterpret_cast<const double2*>(&sm.f.X[r * CP + c2]);
        }
        __syncthreads();
        if (tid == 0) st_release_gpu(ready + (int64_t)i * T + j, 1);
        for (int e = tid; e < CT * CT / 2; e += CTH) {
          const int r = e >> 5, c2 = (e & 31) * 2;
          const double2 v = *reinterpret_cast<const double2*>(&sm.f.X[r * CP + c2]);
          for (int q = 0; q < NR; ++q)
            if (q != RK) *reinterpret_cast<double2*>(P.peerDinv[q] + xoff + r * CT + c2) = v;
        }
        __syncthreads();
        if (tid < NR && tid != RK) st_release_sys(P.peerReady[tid] + (int64_t)i * T + j, 1);
        for (int e = tid; e < CT * CT; e += CTH) {
          const int r = e >> 6, c = e & 63;
          const double v = (r >= c) ? sm.f.M[c * CP + r] : 0.0;
          for (int q = 0; q < NR; ++q) P.peerA[q][tile_off + (int64_t)r * ld + c] = v;
        }
      } else {
        double* Xg = Dinv + xoff;
        for (int e = tid; e < CT * CT; e += CTH) Xg[e] = sm.f.X[(e >> 6) * CP + (e & 63)];
        // publish: the CTA barrier orders every thread's stores before thread 0's release store, and a
        // gpu-scope release is cumulative -- no per-thread fence (the pattern of a split-K semaphore)
        __syncthreads();
        if (tid == 0) st_release_gpu(ready + (int64_t)i * T + j, 1);
        for (int e = tid; e < CT * CT; e += CTH) {
          const int r = e >> 6, c = e & 63;
          Cij[(int64_t)r * ld + c] = (r >= c) ? sm.f.M[c * CP + r] : 0.0;
        }
      }
      if (tr) { tr[6] = clock64(); tr[7] = tr[6]; tr[3] = gtimer(); }
      continue;
    } else {
      // S row-major into M, X_j into X, L(i,j) = S X_j^T
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int row = wm + 8 * ii + fr, col = wn + 8 * jj + 2 * fk;
          double2 v;
          v.x = __dsub_rn(cij[ii][jj].x, acc[ii][jj][0]);
          v.y = __dsub_rn(cij[ii][jj].y, acc[ii][jj][1]);
          *reinterpret_cast<double2*>(&sm.f.M[row * CP + col]) = v;
        }
      if (tid == 0) wait_ready<MULTI>(ready + (int64_t)j * T + j);
      __syncthreads();
      const double* Xg = Dinv + (int64_t)j * CT * CT;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = tid + q * CTH, row = idx >> 5, piece = idx & 31;
        cd_cp16(&sm.f.X[row * CP + piece * 2], Xg + row * CT + piece * 2);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      double out[4][2][2];
#pragma unroll
      for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) out[ii][jj][0] = out[ii][jj][1] = 0.0;
      // X_j is lower triangular: column tile wn..wn+15 only needs k < wn + 16
      for (int ks = 0; ks < wn + 16; ks += 4) {
        double a[4], b[2];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) a[ii] = sm.f.M[(wm + 8 * ii + fr) * CP + ks + fk];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) b[jj] = sm.f.X[(wn + 8 * jj + fr) * CP + ks + fk];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) cd_dmma(out[ii][jj], a[ii], b[jj]);
      }
      if (MULTI) {
        // stage the tile in shared memory so that the pushes to the peers are row-contiguous 16-byte
        // stores (a warp writes one 512-byte row of the tile per instruction)
        __syncthreads();                                   // every warp is done reading M and X
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int row = wm + 8 * ii + fr, col = wn + 8 * jj + 2 * fk;
            *reinterpret_cast<double2*>(&sm.f.M[row * CP + col]) = make_double2(out[ii][jj][0], out[ii][jj][1]);
          }
        __syncthreads();
        for (int e = tid; e < CT * CT / 2; e += CTH) {           // local copy + local flag first
          const int r = e >> 5, c2 = (e & 31) * 2;
          *reinterpret_cast<double2*>(Cij + (int64_t)r * ld + c2) = *reinterpret_cast<const double2*>(&sm.f.M[r * CP + c2]);
        }
        __syncthreads();
        if (tid == 0) st_release_gpu(ready + (int64_t)i * T + j, 1);
        for (int e = tid; e < CT * CT / 2; e += CTH) {
          const int r = e >> 5, c2 = (e & 31) * 2;
          const double2 v = *reinterpret_cast<const double2*>(&sm.f.M[r * CP + c2]);
          for (int q = 0; q < NR; ++q)
            if (q != RK) *reinterpret_cast<double2*>(P.peerA[q] + tile_off + (int64_t)r * ld + c2) = v;
        }
      } else {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int row = wm + 8 * ii + fr, col = wn + 8 * jj + 2 * fk;
            *reinterpret_cast<double2*>(Cij + (int64_t)row * ld + col) = make_double2(out[ii][jj][0], out[ii][jj][1]);
          }
      }
    }
    if (tr) tr[6] = clock64();
    __syncthreads();
    if (MULTI) {
      if (tid < NR && tid != RK) st_release_sys(P.peerReady[tid] + (int64_t)i * T + j, 1);
    } else {
      if (tid == 0) st_release_gpu(ready + (int64_t)i * T + j, 1);
    }
    if (tr) { tr[7] = clock64(); tr[3] = gtimer(); }
  }
}

// R32[a, b] = L[npad-1-a, npad-1-b] for b >= a (0 below); rt_hi + rt_lo = R32^T on and below its
// diagonal, split into TF32 parts (the part above is never read and left untouched); Ud[blk][p][q] = inverse of the 32x32 diagonal
// block R[32blk.., 32blk..] = flip of the matching diagonal block of the tile inverses.
// One CTA per 32x32 tile pair (ta <= tb) and matrix (blockIdx.z); the transpose goes through shared
// memory so that all global accesses are coalesced.
struct CholExportParams {
  const double* A[CHOL_MAXJ];
  const double* Dinv[CHOL_MAXJ];
  float* r32[CHOL_MAXJ];
  float* rt_hi[CHOL_MAXJ];
  float* rt_lo[CHOL_MAXJ];
  float* ud32[CHOL_MAXJ];
};

__global__ void __launch_bounds__(256) chol_export_kernel(const __grid_constant__ CholExportParams P, int64_t n,
                                                          int64_t npad) {
  __shared__ float tile[32][33];
  const int job = blockIdx.z;
  const double* __restrict__ A = P.A[job];
  const double* __restrict__ Dinv = P.Dinv[job];
  float* __restrict__ r32 = P.r32[job];
  float* __restrict__ rt_hi = P.rt_hi[job];
  float* __restrict__ rt_lo = P.rt_lo[job];
  float* __restrict__ ud32 = P.ud32[job];
  const int64_t nt = (n + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  // grid (nt, nt, jobs): CTA (tb, ta) with tb >= ta (a linear tile id decoded by a loop cost O(nt) per CTA:
  // milliseconds at n = 28672)
  const int64_t tb = blockIdx.x, ta = blockIdx.y;
  if (tb < ta || tb >= nt) return;
  const int64_t a0 = ta * 32, b0 = tb * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t a = a0 + i, b = b0 + tx;
    float v = 0.0f;
    if (a < n && b < n && b >= a) v = (float)A[(npad - 1 - a) * npad + (npad - 1 - b)];
    tile[i][tx] = v;
    if (a < n && b < n) r32[a * n + b] = v;
    if (tb > ta) {                                   // mirror tile of R32 is all zeros
      const int64_t a2 = b0 + i, b2 = a0 + tx;
      if (a2 < n && b2 < n) r32[a2 * n + b2] = 0.0f;
    }
  }
  __syncthreads();
  if (rt_hi) {
    for (int i = ty; i < 32; i += 8) {
      const int64_t row = b0 + i, col = a0 + tx;     // RT[b][a] = R[a][b], already split for the tensor path
      if (row < n && col < n) {
        float h, l;
        split_tf32(tile[tx][i], h, l);
        rt_hi[row * n + col] = h;
        rt_lo[row * n + col] = l;
      }
    }
  }
  if (ta == tb) {
    const int64_t fb = npad / 32 - 1 - ta;           // 32-block index in factor order
    const int64_t tl = fb >> 1;
    const int sub = (int)(fb & 1) * 32;
    for (int i = ty; i < 32; i += 8)
      ud32[ta * 1024 + i * 32 + tx] = (float)Dinv[tl * CT * CT + (int64_t)(sub + 31 - i) * CT + (sub + 31 - tx)];
  }
}

static inline int64_t cpad64(int64_t n) { return (n + CT - 1) / CT * CT; }

// workspace of one matrix: A [npad, npad] fp64 | Dinv [T, 64, 64] fp64 | ready [T*T] int | ticket
struct ChWs { double* A; double* Dinv; int* ready; int* ticket; };
static inline ChWs chol_ws_layout(void* ws, int64_t n) {
  const int64_t npad = cpad64(n), T = npad / CT;
  ChWs w;
  w.A = (double*)ws;
  w.Dinv = w.A + npad * npad;
  w.ready = (int*)(w.Dinv + T * CT * CT);
  w.ticket = w.ready + T * T;
  return w;
}

static int env_int(const char* name, int dflt) {
  const char* ev = getenv(name);
  return ev ? atoi(ev) : dflt;
}

// (1) gather: flip + permutation + damp into the fp64 workspaces; clears flags, tickets and info
static int chol_gather(int njobs, const float* const* h, int64_t n, const int64_t* const* order,
                       const float* const* dampval, void* const* ws, int32_t* const* info, cudaStream_t st) {
  const int64_t npad = cpad64(n), T = npad / CT;
  CholGatherParams<float> G;
  for (int k = 0; k < njobs; ++k) {
    const ChWs w = chol_ws_layout(ws[k], n);
    G.h[k] = h[k]; G.order[k] = order ? order[k] : nullptr; G.dampval[k] = dampval ? dampval[k] : nullptr;
    G.A[k] = w.A; G.flags[k] = w.ready; G.info[k] = info[k];
  }
  const int64_t cap = (int64_t)sm_count() * 16 / (njobs > 16 ? 16 : njobs) + 1;
  int64_t blocks = ceil_div(npad * npad, 256);
  dim3 grid((unsigned)(blocks < cap ? blocks : cap), (unsigned)njobs);
  chol_gather_kernel<float><<<grid, 256, 0, st>>>(G, n, npad, T * T + 1);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

// (2) the tile-task kernel over `njobs` matrices (nranks == 1) or one matrix over nranks GPUs
static int chol_dag(int njobs, int64_t n, void* const* ws, int32_t* const* info, int nranks, int rank,
                    void* const* peer_ws, cudaStream_t st) {
  const int64_t npad = cpad64(n);
  const int T = (int)(npad / CT);
  CholDagParams P;
  P.njobs = njobs; P.T = T; P.nranks = nranks; P.rank = rank; P.ld = npad;
  // rows per ownership block of the distributed factorisation (SLK_CHOL_DIST_BLOCK; 1 = plain cyclic rows)
  static int rowblock = -2;
  if (rowblock == -2) rowblock = env_int("SLK_CHOL_DIST_BLOCK", 1);
  P.rowblock = rowblock < 1 ? 1 : rowblock;
  for (int k = 0; k < njobs; ++k) {
    const ChWs w = chol_ws_layout(ws[k], n);
    P.A[k] = w.A; P.Dinv[k] = w.Dinv; P.ready[k] = w.ready; P.info[k] = info[k];
    if (k == 0) P.ticket = w.ticket;
  }
  for (int q = 0; q < nranks && nranks > 1; ++q) {
    const ChWs w = chol_ws_layout(peer_ws[q], n);
    P.peerA[q] = w.A; P.peerDinv[q] = w.Dinv; P.peerReady[q] = w.ready;
  }
  const int64_t ntasks = (int64_t)njobs * T * (T + 1) / 2;
  const int64_t sms = sm_count();
  int64_t grid;
  size_t smem = sizeof(CholSmem);
  if (nranks > 1) {
    // every rank owns ~1/P of the tasks of every column; the FP64 pipe is the bound
    // (SLK_CHOL_DIST_CTAS: CTAs per SM x 100, experiments)
    static int dist_pct = -2;
    if (dist_pct == -2) dist_pct = env_int("SLK_CHOL_DIST_CTAS", 200);
    grid = sms * dist_pct / 100;
    if (grid < 1) grid = 1;
  } else if (njobs == 1) {
    // CTAs: a small factorisation is bound by the chain of diagonal tiles, not by the number of CTAs
    // (n = 768: the same 0.28 ms with 11 to 26 CTAs), and every CTA beyond the useful ones only holds
    // an SM slot while it spins; mid-sized ones get 1.5 CTAs per tile row (n = 3072: 1.14 ms with 2 per
    // row, 1.29 ms with 1.5, 1.65 ms with 1), n >= 4096 -- bound by the FP64 pipe, and big enough to
    // own the GPU -- 2 per row.  SLK_CHOL_GRID overrides the percentage (experiments).
    static int grid_pct = -2;
    if (grid_pct == -2) grid_pct = env_int("SLK_CHOL_GRID", 0);
    const int pct = grid_pct >= 50 ? grid_pct : (T <= 16 ? 100 : (T < 64 ? 150 : 200));
    grid = (int64_t)T * pct / 100 + 2;
    if (grid > 2 * sms) grid = 2 * sms;
  } else {
    // batch: CTAs do not spin (the next ticket is another matrix's tile), so the grid is a throughput
    // choice.  Default ONE CTA per SM with the dynamic shared memory padded past half an SM's, which
    // leaves half of the register file and ~110 KB of shared memory of every SM to the latency-bound
    // sweeps and the issue-bound scale searches of the other layers that run beside the factorisations
    // (SLK_CHOL_BATCH_CTAS: CTAs per SM x 100; SLK_CHOL_BATCH_SMEM_KB: dynamic shared memory per CTA).
    static int ctas_pct = -2, smem_kb = -2;
    if (ctas_pct == -2) ctas_pct = env_int("SLK_CHOL_BATCH_CTAS", 100);
    if (smem_kb == -2) smem_kb = env_int("SLK_CHOL_BATCH_SMEM_KB", 116);
    grid = sms * ctas_pct / 100;
    if (grid < 1) grid = 1;
    if ((size_t)smem_kb * 1024 > smem) smem = (size_t)smem_kb * 1024;
    if (smem > 227 * 1024) smem = 227 * 1024;
  }
  if (grid > ntasks) grid = ntasks;
  if (nranks > 1) {
    SLK_CUDA(cudaFuncSetAttribute(chol_dag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chol_dag_kernel<true><<<(unsigned)grid, CTH, smem, st>>>(P);
  } else {
    SLK_CUDA(cudaFuncSetAttribute(chol_dag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    chol_dag_kernel<false><<<(unsigned)grid, CTH, smem, st>>>(P);
  }
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

// (3) export: R (fp32), R^T as TF32 parts, 32x32 diagonal-block inverses
static int chol_export(int njobs, int64_t n, void* const* ws, float* const* r32, float* const* rt_hi,
                       float* const* rt_lo, float* const* ud32, cudaStream_t st) {
  const int64_t npad = cpad64(n);
  CholExportParams E;
  for (int k = 0; k < njobs; ++k) {
    const ChWs w = chol_ws_layout(ws[k], n);
    E.A[k] = w.A; E.Dinv[k] = w.Dinv; E.r32[k] = r32[k]; E.ud32[k] = ud32[k];
    const bool split = rt_hi && rt_lo && rt_hi[k] && rt_lo[k];
    E.rt_hi[k] = split ? rt_hi[k] : nullptr;
    E.rt_lo[k] = split ? rt_lo[k] : nullptr;
  }
  const int64_t nt32 = (n + 31) / 32;
  dim3 grid((unsigned)nt32, (unsigned)nt32, (unsigned)njobs);
  chol_export_kernel<<<grid, 256, 0, st>>>(E, n, npad);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

}  // namespace slk

using namespace slk;

extern "C" {

/* development aid: per-task trace buffer (8 int64 per tile task), NULL disables */
int slk_debug_chol_trace(void* buf) {
  long long* p = (long long*)buf;
  SLK_CUDA(cudaMemcpyToSymbol(g_chol_trace, &p, sizeof(p)));
  return SLK_OK;
}

size_t slk_chol_factor_ws_bytes(int64_t n) {
  const int64_t npad = cpad64(n);
  const int64_t T = npad / CT;
  return (size_t)(npad * npad + T * CT * CT) * sizeof(double) + (size_t)(T * T + 1) * sizeof(int) + 256;
}

int slk_chol_factor_f32(const float* h, int64_t n, const int64_t* order, const float* dampval, void* ws,
                        size_t ws_bytes, float* r32, float* rt_hi, float* rt_lo, float* ud32, int32_t* info,
                        void* stream) {
  SLK_REQUIRE(h && info && r32 && ud32 && n >= 1, "bad arguments");
  SLK_REQUIRE(ws && ws_bytes >= slk_chol_factor_ws_bytes(n), "workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = chol_gather(1, &h, n, &order, &dampval, &ws, &info, st))) return rc;
  if ((rc = chol_dag(1, n, &ws, &info, 1, 0, nullptr, st))) return rc;
  return chol_export(1, n, &ws, &r32, &rt_hi, &rt_lo, &ud32, st);
}

/* The same factorisation for `njobs` matrices of ONE size n in one launch sequence (gather, tile-task
   kernel with a shared ticket queue, export).  Every argument is a HOST array of njobs device pointers
   (order / dampval / rt_hi / rt_lo: the array or single entries may be NULL); ws[k] holds
   slk_chol_factor_ws_bytes(n) bytes.  More than 64 matrices are processed in launches of 64. */
int slk_chol_factor_batched_f32(int32_t njobs, const float* const* h, int64_t n, const int64_t* const* order,
                                const float* const* dampval, void* const* ws, float* const* r32,
                                float* const* rt_hi, float* const* rt_lo, float* const* ud32, int32_t* const* info,
                                void* stream) {
  SLK_REQUIRE(njobs >= 1 && h && ws && r32 && ud32 && info && n >= 1, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int k0 = 0; k0 < njobs; k0 += CHOL_MAXJ) {
    const int nb = (njobs - k0) < CHOL_MAXJ ? (njobs - k0) : CHOL_MAXJ;
    for (int k = k0; k < k0 + nb; ++k) SLK_REQUIRE(h[k] && ws[k] && r32[k] && ud32[k] && info[k], "NULL pointer in job %d", k);
    int rc;
    if ((rc = chol_gather(nb, h + k0, n, order ? order + k0 : nullptr, dampval ? dampval + k0 : nullptr, ws + k0, info + k0, st)))
      return rc;
    if ((rc = chol_dag(nb, n, ws + k0, info + k0, 1, 0, nullptr, st))) return rc;
    if ((rc = chol_export(nb, n, ws + k0, r32 + k0, rt_hi ? rt_hi + k0 : nullptr, rt_lo ? rt_lo + k0 : nullptr, ud32 + k0, st)))
      return rc;
  }
  return SLK_OK;
}

/* One matrix factored by `nranks` GPUs (one process each).  Three stream-ordered steps; the caller
   separates them with a barrier over the ranks ON THE SAME STREAM (e.g. an NCCL all-reduce of one
   element): after slk_chol_dist_gather_f32 (every rank's flags are cleared before any peer pushes a
   tile) and after slk_chol_dist_factor (every peer's pushes have landed before the export reads them).
   ws: this rank's workspace (slk_chol_factor_ws_bytes(n) bytes) in memory the peers can write
   (slk_peer_alloc + slk_peer_open); peer_ws: HOST array of nranks device pointers, peer_ws[rank] == ws.
   Every rank gathers the full matrix, factors the tile rows i with i % nranks == rank, and ends up
   holding the complete factor, so the export (and everything after it) is local. */
int slk_chol_dist_gather_f32(const float* h, int64_t n, const int64_t* order, const float* dampval, void* ws,
                             size_t ws_bytes, int32_t* info, void* stream) {
  SLK_REQUIRE(h && info && n >= 1, "bad arguments");
  SLK_REQUIRE(ws && ws_bytes >= slk_chol_factor_ws_bytes(n), "workspace too small");
  return chol_gather(1, &h, n, &order, &dampval, &ws, &info, (cudaStream_t)stream);
}

int slk_chol_dist_factor(int64_t n, void* ws, int32_t nranks, int32_t rank, void* const* peer_ws, int32_t* info,
                         void* stream) {
  SLK_REQUIRE(ws && info && n >= 1 && nranks >= 1 && nranks <= CHOL_MAXP && rank >= 0 && rank < nranks, "bad arguments");
  SLK_REQUIRE(nranks == 1 || (peer_ws && peer_ws[rank] == ws), "peer_ws[rank] must be this rank's workspace");
  return chol_dag(1, n, &ws, &info, nranks, rank, peer_ws, (cudaStream_t)stream);
}

int slk_chol_dist_export_f32(int64_t n, void* ws, float* r32, float* rt_hi, float* rt_lo, float* ud32, void* stream) {
  SLK_REQUIRE(ws && r32 && ud32 && n >= 1, "bad arguments");
  return chol_export(1, n, &ws, &r32, &rt_hi, &rt_lo, &ud32, (cudaStream_t)stream);
}

}  // extern "C"
