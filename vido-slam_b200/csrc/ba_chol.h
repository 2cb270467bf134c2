// ba_chol.h -- dense Cholesky solve of the reduced camera system of the window BA, in shared memory, by ONE CTA of 512
// threads (device code shared by ba_window.cu and tools/chol_probe2.cu).
//
// Replaces LinearSolverCSparse::solve on the sliding-window graph (3rdparty/g2o/g2o/solvers/linear_solver_csparse.h:108-141):
// the reference factors the full (poses + points) H with CSparse; here the points are eliminated first (Schur complement)
// and the reduced 6W x 6W system is factored densely.  Same linear system, same positive-definiteness test (all pivots > 0).
//
// Structure (right-looking LL^T, block size 8 = one FP64 tensor-core tile):
//   * the trailing matrix lives in REGISTERS for the whole factorisation: every 8x8 tile is owned by one of the 12 "bulk"
//     warps as an m8n8k4 accumulator fragment (2 doubles per lane); the rank-8 trailing update of a tile is two
//     mma.sync.m8n8k4.f64 (DMMA) whose A/B fragments come from the 8-wide panel buffer Lp.  Only the tiles of the next
//     block column are written back to shared memory after a step (they are final), so the shared-memory traffic per step
//     is the panel, not the trailing matrix (the scalar version of round 1 was bound by exactly that traffic).
//   * warp 0 is the "chain" warp: the pivot chain (8 x (rsqrt + dependent FMA) per block) is the critical path of the
//     factorisation, so warp 0 runs one block ahead -- it solves the 8 panel rows of block jb+1, applies their update to
//     the diagonal block jb+1 and factors it while the bulk warps are still busy with step jb.
//   * one named barrier (panel complete; the chain warp only arrives) and one CTA barrier per block step.
//   * the right-hand side rides along as row np of the matrix (forward substitution = part of the panel steps); the
//     8x8 inverses of the diagonal factors are computed off the critical path and turn the back substitution into
//     matrix-vector products.
#pragma once
#include <cuda_runtime.h>

#define BC_THREADS 512
#define BC_LPS 12        // row stride of the panel buffer in doubles: the 8 rows x 4 doubles of a fragment load then touch every
                         // bank exactly twice (stride 8 or 9 would give 4-way conflicts)
#define BC_MAXT 18       // block rows at the largest window (W = 24: n = 144)
#define BC_BULK_WARPS 12 // warps 0,1,2, 4,5,6, 8,9,10, 12,13,14: the scheduler of warps 3,7,11,15 serves the chain (15) and the helper (11)
                         // only (measured: with bulk warps on that scheduler the chain's step grows from 1650 to 2100-2500 cycles)
#define BC_MAX_SLOTS 13  // tiles per bulk warp at BC_MAXT: ceil(152 / 12)
#define BC_ACTIVE ((BC_BULK_WARPS + 1) * 32)   // threads taking part in the factorisation barriers

struct CholSm {
  double* S;     // (np + 1) x ld: lower triangle of the system, row np = right-hand side
  double* Lp;    // np x BC_LPS: panel (block column of L) of the current step
  double* dinv;  // np: 1 / L_jj
  double* Linv;  // T x 64: inverses of the diagonal factors (row-major 8x8, lower)
  int n, np, ld; // n = 6W unknowns, np = n rounded up to a multiple of 8 (identity padding), ld = np + 1 (odd)
};

__host__ __device__ inline int bc_np(int n) { return (n + 7) & ~7; }
__host__ __device__ inline size_t bc_smem_doubles(int n) {
  const int np = bc_np(n), ld = np + 1;
  return (size_t)(np + 1) * ld + (size_t)np * BC_LPS + np + (size_t)(np / 8) * 64;
}
__device__ __forceinline__ void bc_carve(double* base, int n, CholSm& cs) {
  cs.n = n; cs.np = bc_np(n); cs.ld = cs.np + 1;
  cs.S = base;
  cs.Lp = cs.S + (size_t)(cs.np + 1) * cs.ld;
  cs.dinv = cs.Lp + (size_t)cs.np * BC_LPS;
  cs.Linv = cs.dinv + cs.np;
}

// 1/sqrt(d) for a positive normal d: MUFU.RSQ64H + one third-order correction, without the library's special-case
// subroutine (a CALL inside a latency-critical chain makes the compiler park live values in local memory)
__device__ __forceinline__ double bc_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(d, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

__device__ __forceinline__ void bc_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 8x8 Cholesky of the diagonal block at D (row stride ld) by ONE lane, in registers, right-looking so that the next pivot
// is ready as early as possible.  Writes L in place and 1/L_jj to dinv.  false = not positive definite.
__device__ __forceinline__ bool bc_chol8(double* D, int ld, double* dinv) {
  double a[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c <= r; c++) a[r][c] = D[r * ld + c];
  bool ok = true;
  double rs[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const double d = a[j][j];
    ok = ok && (d > 0);
    rs[j] = bc_rsqrt(d);
    a[j][j] = d * rs[j];
#pragma unroll
    for (int i = j + 1; i < 8; i++) a[i][j] *= rs[j];
#pragma unroll
    for (int c = j + 1; c < 8; c++)
#pragma unroll
      for (int i = c; i < 8; i++) a[i][c] = fma(-a[i][j], a[c][j], a[i][c]);
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
#pragma unroll
    for (int c = 0; c <= r; c++) D[r * ld + c] = a[r][c];
    dinv[r] = rs[r];
  }
  return ok;
}

// row <- row * L11^-T (8-step forward chain) for the row starting at `row` (8 entries); L11 at D, reciprocals in dv
__device__ __forceinline__ void bc_trsm_row(const double* D, int ld, const double* dv, double* row, double* lp /* or nullptr */) {
  double o[8];
#pragma unroll
  for (int k = 0; k < 8; k++) o[k] = row[k];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double s = o[k];
#pragma unroll
    for (int m = 0; m < k; m++) s = fma(-o[m], D[k * ld + m], s);
    o[k] = s * dv[k];
  }
#pragma unroll
  for (int k = 0; k < 8; k++) row[k] = o[k];
  if (lp) {
#pragma unroll
    for (int k = 0; k < 8; k++) lp[k] = o[k];
  }
}

// Factor S = L L^T in place (lower triangle) and carry the right-hand side (row np) through the forward substitution.
// Called by all BC_THREADS threads of the CTA.  s_bad: shared flag, set when a pivot is not positive.
// tk: optional cycle counters (probe): [0] diag0, [1] chain work per step, [2] chain waiting at the step barrier, [3] bulk
// panel, [4] bulk trailing, [5] bulk waiting, [6] chain trsm, [7] chain diagonal update; [16 + 8 jb ..] per-step trace.
//
// Warp roles (warp specialisation: the register files of the roles never coexist):
//   chain  (warp 15)  the pivot chain, one block ahead of everybody else: panel rows of block jb+1 (8 lanes, scalar), their
//                     update of the diagonal block jb+1 (two DMMA), its 8x8 factorisation (one lane, in registers);
//   helper (warp 11)  the inverse of every freshly factored diagonal block (8 lanes, one 8-step chain), which turns the
//                     panel of the bulk warps and the back substitution into matrix products;
//   bulk   (12 warps: 0,1,2, 4,5,6, 8,9,10, 12,13,14)  panel L21 = A21 L11^-T as two DMMA per 8-row tile, trailing update
//                     of the register-resident tiles (two DMMA per tile, two tiles interleaved), write-back of final tiles.
// The scheduler of warps 3, 7, 11, 15 serves the chain (and the short helper bursts) only; the issue arbiter prefers high
// warp ids.  Barriers: STEP (chain + bulk, end of a step), PANEL (panel complete; the chain only arrives), LINV (chain ->
// helper: block factored), LREADY (helper -> bulk: inverse stored).
#define BC_BAR_PANEL 1
#define BC_BAR_STEP 2
#define BC_BAR_BULK 3
#define BC_BAR_LINV 4
#define BC_BAR_LREADY 5
#define BC_CHAIN_WARP 15
#define BC_HELPER_WARP 11
__device__ __forceinline__ void bc_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(BC_ACTIVE) : "memory"); }
__device__ __forceinline__ void bc_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(BC_ACTIVE) : "memory"); }

// first tile id of block column c (tiles (R, C), R >= C >= 1, without (1,1), numbered column by column)
__device__ __forceinline__ int bc_tile_off(int c, int T) {
  return (c <= 1) ? 0 : (T - 2) + (c - 2) * T - ((c - 1) * c / 2 - 1);
}

// What the factorisation needs in shared memory before it starts, copied from the global copy Sg (same layout): block
// column 0 (the first panel, with the diagonal block 0), the diagonal block 1 (the chain warp's first look-ahead) and the
// right-hand side row.  Everything else goes straight from Sg into the accumulator registers of the bulk warps.
__device__ __forceinline__ void bc_stage(const CholSm cs, const double* __restrict__ Sg) {
  const int np = cs.np, ld = cs.ld;
  for (int i = threadIdx.x; i < 8 * np; i += BC_THREADS) {
    const int r = i >> 3, c = i & 7;
    if (c <= r) cs.S[r * ld + c] = Sg[r * ld + c];
  }
  for (int i = threadIdx.x; i < np; i += BC_THREADS) cs.S[np * ld + i] = Sg[np * ld + i];
  if (np > 8 && threadIdx.x >= BC_THREADS - 64) {
    const int i = threadIdx.x - (BC_THREADS - 64), r = 8 + (i >> 3), c = 8 + (i & 7);
    if (c <= r) cs.S[r * ld + c] = Sg[r * ld + c];
  }
  __syncthreads();
}

__device__ __forceinline__ void bc_factor(const CholSm cs, int* s_bad, long long* tk, const double* __restrict__ Sg = nullptr) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = cs.np, ld = cs.ld, T = np >> 3;
  double* const S = cs.S;
  double* const Lp = cs.Lp;
  const int fr = lane >> 2, fk = lane & 3;
  if (warp == BC_CHAIN_WARP) {
    // =========================== chain warp ===========================
    long long t0 = 0;
    if (tk) t0 = clock64();
    if (lane == 0) *s_bad = bc_chol8(S, ld, cs.dinv) ? 0 : 1;
    __syncwarp();
    asm volatile("bar.arrive %0, 64;" ::"n"(BC_BAR_LINV) : "memory");
    bc_bar_sync(BC_BAR_STEP);   // block 0 is factored, the bulk warps have loaded their tiles
    if (tk && lane == 0) { const long long t1 = clock64(); tk[0] += t1 - t0; t0 = t1; }
    for (int jb = 0; jb < T; jb++) {
      if (*s_bad) break;
      const int j0 = 8 * jb;
      const double* D = S + j0 * ld + j0;
      const double* dv = cs.dinv + j0;
      const bool more = jb + 1 < T;
      if (more && lane < 8) {   // panel rows of block jb+1
        const int i = j0 + 8 + lane;
        bc_trsm_row(D, ld, dv, S + i * ld + j0, Lp + i * BC_LPS);
      }
      __syncwarp();
      bc_bar_arrive(BC_BAR_PANEL);
      long long t6 = 0;
      if (tk && lane == 0) { t6 = clock64(); tk[6] += t6 - t0; }
      if (more) {
        // update of the diagonal block jb+1, D -= L L^T, as two DMMA: the A and the B fragment of L(jb+1, jb) are the same
        // registers (B[k][c] = L(c, k)).  The tile is symmetric; its upper half is carried along and never read.
        double* dp = S + (j0 + 8 + fr) * ld + j0 + 8 + 2 * fk;
        const double* pl = Lp + (j0 + 8 + fr) * BC_LPS + fk;
        double d0 = dp[0], d1 = dp[1];
        const double a0 = pl[0], a1 = pl[4];
        bc_dmma(d0, d1, -a0, a0);
        bc_dmma(d0, d1, -a1, a1);
        dp[0] = d0; dp[1] = d1;
        __syncwarp();
        if (tk && lane == 0) tk[7] += clock64() - t6;
        if (lane == 0 && !bc_chol8(S + (j0 + 8) * ld + j0 + 8, ld, cs.dinv + j0 + 8)) *s_bad = 1;
        __syncwarp();
        asm volatile("bar.arrive %0, 64;" ::"n"(BC_BAR_LINV) : "memory");
      }
      if (tk && lane == 0) { const long long t1 = clock64(); tk[1] += t1 - t0; tk[16 + 8 * jb] = t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_STEP);
      if (tk && lane == 0) { const long long t1 = clock64(); tk[2] += t1 - t0; tk[17 + 8 * jb] = t1 - t0; t0 = t1; }
    }
  } else if (warp == BC_HELPER_WARP) {
    // =========================== helper warp ===========================
    for (int jb = 0; jb < T; jb++) {
      asm volatile("bar.sync %0, 64;" ::"n"(BC_BAR_LINV) : "memory");
      const int bad = *s_bad;
      if (!bad && lane < 8) {   // X = L11^-1, column c by lane c
        const int c = lane, j0 = 8 * jb;
        const double* D = S + j0 * ld + j0;
        const double* dv = cs.dinv + j0;
        double x[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
          double sum = (m == c) ? 1.0 : 0.0;
#pragma unroll
          for (int k = 0; k < m; k++) sum = fma(-D[m * ld + k], (k >= c) ? x[k] : 0.0, sum);
          x[m] = (m >= c) ? sum * dv[m] : 0.0;
        }
#pragma unroll
        for (int m = 0; m < 8; m++) cs.Linv[jb * 64 + m * 8 + c] = x[m];
      }
      __syncwarp();
      bc_bar_arrive(BC_BAR_LREADY);
      if (bad) break;
    }
  } else if ((warp & 3) != 3) {
    // =========================== bulk warps ===========================
    const int bw = warp - (warp >> 2), btid = bw * 32 + lane;   // bulk warp / thread index
    // tile ownership: tile id = bw + 12 * slot, ids run column by column, so the tiles still active at a step are the
    // slots >= smin(step) of every warp: the update below jumps into an unrolled slot sequence (no per-slot activity test
    // for retired tiles).  packed = R | C << 8.
    double c0[BC_MAX_SLOTS + 1], c1[BC_MAX_SLOTS + 1];
    int tRC[BC_MAX_SLOTS + 1];
    const int ntiles = bc_tile_off(T, T);
    const int nsl = (ntiles > bw) ? (ntiles - bw + BC_BULK_WARPS - 1) / BC_BULK_WARPS : 0;
    {  // tile table: one thread per tile decodes its (R, C); the panel buffer is free until the first step
      int* tab = (int*)Lp;
      if (btid < BC_BULK_WARPS * (BC_MAX_SLOTS + 1)) {
        int t = btid, c = 1, R = -1;
        while (c < T) {
          const int cnt = (c == 1) ? T - 2 : T - c;
          if (t < cnt) { R = ((c == 1) ? 2 : c) + t; break; }
          t -= cnt; c++;
        }
        tab[btid] = (R >= 0) ? (R | (c << 8)) : 0;
      }
      asm volatile("bar.sync %0, %1;" ::"n"(BC_BAR_BULK), "n"(BC_BULK_WARPS * 32) : "memory");
#pragma unroll
      for (int s = 0; s <= BC_MAX_SLOTS; s++) tRC[s] = tab[bw + BC_BULK_WARPS * s];
      asm volatile("bar.sync %0, %1;" ::"n"(BC_BAR_BULK), "n"(BC_BULK_WARPS * 32) : "memory");
    }
    const double* src = Sg ? Sg : S;
#pragma unroll
    for (int s = 0; s <= BC_MAX_SLOTS; s++) {
      const int R = tRC[s] & 255, c = tRC[s] >> 8;
      c0[s] = 0; c1[s] = 0;
      if (s < nsl) {
        const int row = 8 * R + fr, col = 8 * c + 2 * fk;
        c0[s] = (row >= col) ? src[row * ld + col] : src[col * ld + row];
        c1[s] = (row >= col + 1) ? src[row * ld + col + 1] : src[(col + 1) * ld + row];
      }
    }
    bc_bar_sync(BC_BAR_STEP);
    long long t0 = 0;
    if (tk && btid == 0) t0 = clock64();
    const double* const lpf = Lp + fr * BC_LPS + fk;   // fragment base of this lane
    for (int jb = 0; jb < T; jb++) {
      bc_bar_sync(BC_BAR_LREADY);
      if (*s_bad) break;
      const int j0 = 8 * jb;
      const bool more = jb + 1 < T;
      const double* Li = cs.Linv + jb * 64;
      // ---- panel: L21 = A21 L11^-T for the 8-row tiles of the blocks >= jb+2 (two DMMA each), and the right-hand side row
      {
        const double b0 = Li[fr * 8 + fk], b1 = Li[fr * 8 + 4 + fk];   // B[k][c] = Linv[c][k]
        for (int rt = jb + 2 + bw; rt < T; rt += BC_BULK_WARPS) {
          double* ar = S + (8 * rt + fr) * ld + j0;
          double r0 = 0, r1 = 0;
          bc_dmma(r0, r1, ar[fk], b0);
          bc_dmma(r0, r1, ar[4 + fk], b1);
          __syncwarp();
          ar[2 * fk] = r0; ar[2 * fk + 1] = r1;
          double* lp = Lp + (8 * rt + fr) * BC_LPS + 2 * fk;
          lp[0] = r0; lp[1] = r1;
        }
        if (bw == BC_BULK_WARPS - 1 && lane < 8) {
          double* yr = S + np * ld + j0;
          double sa = 0, sb = 0;
#pragma unroll
          for (int m = 0; m < 8; m += 2) {
            sa = fma(yr[m], Li[lane * 8 + m], sa);
            sb = fma(yr[m + 1], Li[lane * 8 + m + 1], sb);
          }
          __syncwarp(0xffu);
          yr[lane] = sa + sb;
        }
      }
      if (tk && btid == 0) { const long long t1 = clock64(); tk[3] += t1 - t0; tk[18 + 8 * jb] = t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_PANEL);
      if (tk && btid == 0) { const long long t1 = clock64(); tk[19 + 8 * jb] = t1 - t0; }
      if (more) {
        // ---- trailing update of the owned tiles, C -= L(R,jb) L(C,jb)^T as two DMMA (k = 0..3, 4..7), two tiles interleaved
        //      (the two DMMA of a tile depend on each other); tiles that received their last update go back to shared
        //      memory: block column jb+1 (the next panel) and the diagonal tile jb+2.  Active tile ids at step jb: >= fa(jb)
        //      = first id of column jb+1, +1 for its diagonal tile (the chain warp owns it by now); the ones below fa(jb+1)
        //      are final.  The pair that contains slot smin may start one slot early: that tile is retired, its registers
        //      are dead, updating them is harmless (it is not written back: s >= smin).
        const int fa = (jb == 0) ? 0 : bc_tile_off(jb + 1, T) + 1;
        const int fn = bc_tile_off(jb + 2, T) + 1;
        const int smin = (fa > bw) ? (fa - bw + BC_BULK_WARPS - 1) / BC_BULK_WARPS : 0;
        const int sfl = (fn > bw) ? (fn - bw + BC_BULK_WARPS - 1) / BC_BULK_WARPS : 0;
#define BC_UPD2(s)                                                                                         \
  case (s) / 2: {                                                                                          \
    if ((s) >= nsl) break;                                                                                 \
    const int Ra_ = tRC[s] & 255, Ca_ = tRC[s] >> 8, Rb_ = tRC[(s) + 1] & 255, Cb_ = tRC[(s) + 1] >> 8;    \
    const double* paa_ = lpf + Ra_ * (8 * BC_LPS);                                                         \
    const double* pba_ = lpf + Ca_ * (8 * BC_LPS);                                                         \
    const double* pab_ = lpf + Rb_ * (8 * BC_LPS);                                                         \
    const double* pbb_ = lpf + Cb_ * (8 * BC_LPS);                                                         \
    const double a0_ = paa_[0], a1_ = paa_[4], b0_ = pba_[0], b1_ = pba_[4];                               \
    const double e0_ = pab_[0], e1_ = pab_[4], f0_ = pbb_[0], f1_ = pbb_[4];                               \
    bc_dmma(c0[s], c1[s], -a0_, b0_);                                                                      \
    bc_dmma(c0[(s) + 1], c1[(s) + 1], -e0_, f0_);                                                          \
    bc_dmma(c0[s], c1[s], -a1_, b1_);                                                                      \
    bc_dmma(c0[(s) + 1], c1[(s) + 1], -e1_, f1_);                                                          \
    if ((s) >= smin && (s) < sfl) {                                                                        \
      double* dst_ = S + (8 * Ra_ + fr) * ld + 8 * Ca_ + 2 * fk;                                           \
      dst_[0] = c0[s]; dst_[1] = c1[s];                                                                    \
    }                                                                                                      \
    if ((s) + 1 < sfl && (s) + 1 < nsl) {                                                                  \
      double* dst_ = S + (8 * Rb_ + fr) * ld + 8 * Cb_ + 2 * fk;                                           \
      dst_[0] = c0[(s) + 1]; dst_[1] = c1[(s) + 1];                                                        \
    }                                                                                                      \
  }
        switch (smin >> 1) {
          BC_UPD2(0) BC_UPD2(2) BC_UPD2(4) BC_UPD2(6) BC_UPD2(8) BC_UPD2(10) BC_UPD2(12)
          default: break;
        }
#undef BC_UPD2
        if (bw == BC_BULK_WARPS - 1) {  // right-hand side row: y(c) -= L(rhs, jb) . L(c, jb)
          const double* lr = S + np * ld + j0;
          const double r0 = lr[0], r1 = lr[1], r2 = lr[2], r3 = lr[3], r4 = lr[4], r5 = lr[5], r6 = lr[6], r7 = lr[7];
          for (int c = j0 + 8 + lane; c < np; c += 32) {
            const double* lc = Lp + c * BC_LPS;
            const double s01 = fma(r1, lc[1], r0 * lc[0]), s23 = fma(r3, lc[3], r2 * lc[2]);
            const double s45 = fma(r5, lc[5], r4 * lc[4]), s67 = fma(r7, lc[7], r6 * lc[6]);
            S[np * ld + c] -= (s01 + s23) + (s45 + s67);
          }
        }
      }
      if (tk && btid == 0) { const long long t1 = clock64(); tk[4] += t1 - t0; tk[20 + 8 * jb] = t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_STEP);
      if (tk && btid == 0) { const long long t1 = clock64(); tk[5] += t1 - t0; tk[21 + 8 * jb] = t1 - t0; t0 = t1; }
    }
  }
  __syncthreads();
}

// Back substitution L^T x = y by the chain warp (the other warps return at once): x overwrites row np.  Per block step
// lane k (mod 8) forms unknown k as a row of the matrix-vector product with the precomputed inverse of the diagonal
// factor, a shuffle broadcast hands the 8 unknowns to every lane, and the lanes update their rows above the block.  The
// rows of L a lane needs for that are loaded BEFORE the unknowns are formed (they do not depend on them).
__device__ __forceinline__ void bc_backsolve(const CholSm cs) {
  if ((threadIdx.x >> 5) != BC_CHAIN_WARP) return;
  const int lane = threadIdx.x & 31, np = cs.np, ld = cs.ld, T = np >> 3;
  double* const ys = cs.S + (size_t)np * ld;
  const double* const S = cs.S;
  constexpr int NCH = (BC_MAXT * 8 + 31) / 32;
  const int k8 = lane & 7;
  for (int jb = T - 1; jb >= 0; jb--) {
    const int j0 = 8 * jb;
    const double* Li = cs.Linv + jb * 64;
    double lv[NCH][8], yv[NCH];
#pragma unroll
    for (int u = 0; u < NCH; u++) {
      const int i = lane + 32 * u;
      const bool on = i < j0;
      const double* col = S + j0 * ld + (on ? i : 0);
      yv[u] = on ? ys[i] : 0.0;
#pragma unroll
      for (int k = 0; k < 8; k++) lv[u][k] = on ? col[k * ld] : 0.0;
    }
    // x_k = sum_{m >= k} Linv[m][k] y_m (two partial sums to halve the dependent chain); Linv is stored with its zeros
    double sa = 0, sb = 0;
#pragma unroll
    for (int m = 0; m < 8; m += 2) {
      sa = fma(Li[m * 8 + k8], ys[j0 + m], sa);
      sb = fma(Li[(m + 1) * 8 + k8], ys[j0 + m + 1], sb);
    }
    const double xk = sa + sb;
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = __shfl_sync(0xffffffffu, xk, k);
#pragma unroll
    for (int u = 0; u < NCH; u++) {
      const int i = lane + 32 * u;
      double ta = yv[u], tb = 0;
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        ta = fma(-lv[u][k], x[k], ta);
        tb = fma(-lv[u][k + 1], x[k + 1], tb);
      }
      if (i < j0) ys[i] = ta + tb;
    }
    if (lane < 8) ys[j0 + lane] = xk;
    __syncwarp();
  }
}
