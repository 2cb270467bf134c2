// ba_chol.h -- dense Cholesky solve of the reduced camera system of the window BA, in shared memory, by ONE CTA of 512
// threads (device code shared by ba_window.cu and tools/chol_probe2.cu).
//
// Replaces LinearSolverCSparse::solve on the sliding-window graph (3rdparty/g2o/g2o/solvers/linear_solver_csparse.h:108-141):
// the reference factors the full (poses + points) H with CSparse; here the points are eliminated first (Schur complement)
// and the reduced 6W x 6W system is factored densely.  Same linear system, same positive-definiteness test (all pivots > 0).
//
// Structure (LL^T, block size 8 = one FP64 tensor-core tile, mma.sync.m8n8k4.f64 = SASS DMMA.8x8x4):
//   * warp 15 is the "chain" warp: the pivot chain (8 x (rsqrt + dependent FMA) per block) is the critical path of the
//     factorisation, so it runs one block ahead -- it solves the 8 panel rows of block jb+1, applies their update to the
//     diagonal block jb+1 (two DMMA) and factors it (one lane, in registers) while the bulk warps are busy with step jb;
//   * twelve "bulk" warps form the panel (thread per row) and update tiles LEFT-looking: at step jb only the tiles of block
//     column jb+1 and the diagonal tile jb+2 are brought up to date, with all panels 0..jb, as chains of DMMA over four
//     accumulator pairs; the tile's original value comes from the global copy of the system, fetched a step ahead;
//   * warp 11 inverts every freshly factored diagonal block off the critical path: the back substitution then is a
//     matrix-vector product per block;
//   * named barriers with immediate ids (panel complete; end of step; block factored) instead of CTA barriers;
//   * the right-hand side rides along as row np of the matrix (forward substitution = part of the panel steps).
// Measured alternatives are recorded next to the code they lost against (right-looking register-resident trailing matrix, panel
// as a DMMA product with the inverse factor, 14 bulk warps, software-pipelined tile loop, one barrier per 32-row chunk in the
// back substitution): profiles/r2_chol_probe.txt, tools/chol_probe2.cu.
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
// helper: block factored).
#define BC_BAR_PANEL 1
#define BC_BAR_STEP 2
#define BC_BAR_BULK 3
#define BC_BAR_LINV 4
#define BC_CHAIN_WARP 15
#define BC_HELPER_WARP 11
// barrier id and thread count are immediates (a register operand takes the slow dispatch path of BAR)
#define bc_bar_sync(id) asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(BC_ACTIVE) : "memory")
#define bc_bar_arrive(id) asm volatile("bar.arrive %0, %1;" ::"n"(id), "n"(BC_ACTIVE) : "memory")

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
      if (bad) break;
    }
  } else if ((warp & 3) != 3) {
    // =========================== bulk warps ===========================
    // LEFT-looking for everything the chain does not need at once: at step jb only the tiles that become the next panel
    // (block column jb+1, rows >= jb+2) and the diagonal tile jb+2 are brought up to date, with ALL the panels 0..jb.  The
    // updates with the panels 0..jb-1 (already final in S) run while the helper inverts the diagonal factor and the panel of
    // this step is formed; only the last rank-8 update (panel jb, from the conflict-free buffer Lp) follows the panel.
    // A right-looking version (every trailing tile resident in registers, updated at every step) was measured first: its work
    // is front-loaded (104 tiles at step 0, ~2300 cycles against ~1650 of the chain warp) and the first 8 of 15 steps were
    // bound by it; left-looking the busiest step has 56 tile updates spread over 12 warps, most of them off the critical path.
    const int bw = warp - (warp >> 2), btid = bw * 32 + lane;   // bulk warp / thread index
    // The original value of a tile (from the global copy Sg when the caller staged only the first panel) is fetched one step
    // AHEAD of its use: an L2 round trip is several hundred cycles, more than the early steps' few DMMA can hide.
    const double* const src = Sg ? Sg : S;
    double on0[2] = {0, 0}, on1[2] = {0, 0};
    auto fetch = [&](int jn) {
      const int ncol = T - jn - 2 > 0 ? T - jn - 2 : 0, nt = ncol + (jn + 2 < T ? 1 : 0);
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int t = bw + BC_BULK_WARPS * q;
        if (t >= nt) continue;
        const int R = (t < ncol) ? jn + 2 + t : jn + 2, C = (t < ncol) ? jn + 1 : jn + 2;
        const int row = 8 * R + fr, col = 8 * C + 2 * fk;
        on0[q] = (row >= col) ? src[row * ld + col] : src[col * ld + row];
        on1[q] = (row >= col + 1) ? src[row * ld + col + 1] : src[(col + 1) * ld + row];
      }
    };
    fetch(0);
    bc_bar_sync(BC_BAR_STEP);
    long long t0 = 0;
    if (tk && btid == 0) t0 = clock64();
    const double* const lpf = Lp + fr * BC_LPS + fk;   // fragment base of this lane in the panel buffer
    for (int jb = 0; jb < T; jb++) {
      const int j0 = 8 * jb;
      const bool more = jb + 1 < T;
      if (*s_bad) break;
      // ---- panel: rows of the blocks >= jb+2 and the right-hand side row against L11 (thread per row).  (Forming the panel as
      //      a DMMA product with the inverse of L11 was measured too: the bulk warps then wait for that inverse at every step.)
      {
        const double* D = S + j0 * ld + j0;
        const double* dv = cs.dinv + j0;
        const int i = j0 + 16 + btid;
        if (i < np) bc_trsm_row(D, ld, dv, S + i * ld + j0, Lp + i * BC_LPS);
        else if (btid == BC_BULK_WARPS * 32 - 1) bc_trsm_row(D, ld, dv, S + np * ld + j0, nullptr);
      }
      if (tk && btid == 0) tk[22 + 8 * jb] = clock64() - t0;   // probe: warp 0 behind the panel rows
      // ---- tiles of this step: (R, jb+1) for R = jb+2 .. T-1, then the diagonal tile (jb+2, jb+2); at most two per warp
      const int ncol = T - jb - 2 > 0 ? T - jb - 2 : 0, nt = ncol + (jb + 2 < T ? 1 : 0);
      double c0[2] = {0, 0}, c1[2] = {0, 0};
      int tR[2] = {-1, -1}, tC[2] = {0, 0};
      const double oc0[2] = {on0[0], on0[1]}, oc1[2] = {on1[0], on1[1]};
      if (jb + 1 < T) fetch(jb + 1);
      if (tk && btid == 0) tk[23 + 8 * jb] = clock64() - t0;   // probe: ... and behind the prefetch of the next step's tiles
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int t = bw + BC_BULK_WARPS * q;
        if (t >= nt) continue;
        const int R = (t < ncol) ? jb + 2 + t : jb + 2, C = (t < ncol) ? jb + 1 : jb + 2;
        tR[q] = R; tC[q] = C;
        const double o0 = oc0[q], o1 = oc1[q];   // the tile's original value, fetched during the previous step
        double a0 = 0, a1 = 0;
        double b0 = 0, b1 = 0, e0 = 0, e1 = 0, g0 = 0, g1 = 0;   // four accumulator pairs: the 2 jb DMMA of a tile form four
                                                                  // chains instead of one (late steps: one tile, many panels)
        const double* pa = S + (8 * R + fr) * ld + fk;
        const double* pb = S + (8 * C + fr) * ld + fk;
        int k = 0;
        for (; k + 1 < jb; k += 2) {
          const double x0 = pa[8 * k], x1 = pa[8 * k + 4], y0 = pb[8 * k], y1 = pb[8 * k + 4];
          const double u0 = pa[8 * k + 8], u1 = pa[8 * k + 12], v0 = pb[8 * k + 8], v1 = pb[8 * k + 12];
          bc_dmma(a0, a1, -x0, y0);
          bc_dmma(b0, b1, -u0, v0);
          bc_dmma(e0, e1, -x1, y1);
          bc_dmma(g0, g1, -u1, v1);
        }
        if (k < jb) {
          const double x0 = pa[8 * k], x1 = pa[8 * k + 4], y0 = pb[8 * k], y1 = pb[8 * k + 4];
          bc_dmma(a0, a1, -x0, y0);
          bc_dmma(b0, b1, -x1, y1);
        }
        a0 += e0 + o0; a1 += e1 + o1; b0 += g0; b1 += g1;
        c0[q] = a0 + b0; c1[q] = a1 + b1;
      }
      if (tk && btid == 0) { const long long t1 = clock64(); tk[3] += t1 - t0; tk[18 + 8 * jb] = t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_PANEL);
      if (tk && btid == 0) { const long long t1 = clock64(); tk[19 + 8 * jb] = t1 - t0; }
      if (more) {
        // ---- the last update of this step's tiles (panel jb, from Lp) and their write-back
#pragma unroll
        for (int q = 0; q < 2; q++) {
          if (tR[q] < 0) continue;
          const double* pa = lpf + tR[q] * (8 * BC_LPS);
          const double* pb = lpf + tC[q] * (8 * BC_LPS);
          const double x0 = pa[0], x1 = pa[4], y0 = pb[0], y1 = pb[4];
          double z0 = 0, z1 = 0;
          bc_dmma(c0[q], c1[q], -x0, y0);
          bc_dmma(z0, z1, -x1, y1);
          double* dst = S + (8 * tR[q] + fr) * ld + 8 * tC[q] + 2 * fk;
          dst[0] = c0[q] + z0; dst[1] = c1[q] + z1;
        }
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

// Back substitution L^T x = y by FIVE warps (11..15; the other warps return at once): x overwrites row np.
// Warp u owns the rows 32u .. 32u+31 (np <= 144 at the largest window) and keeps their y in registers (lane = row).  Per block step, from the last block up:
// the warp that owns the block gathers its 8 values with shuffles, forms the 8 unknowns as a matrix-vector product with the
// precomputed inverse of the diagonal factor and publishes them in a 2 x 8 shared buffer (the dinv array, dead by now); one
// named barrier later every warp subtracts their contribution from its rows above the block -- 8 FMAs in two chains, the
// rows of L loaded before the barrier.  The single-warp version this replaces issued ~120 instructions per step and was
// bound by that (650 cycles per step measured); here a step is ~35 instructions per warp on its own scheduler.
#define BC_BAR_BACK 5
__device__ __forceinline__ void bc_backsolve(const CholSm cs) {
  const int warp = threadIdx.x >> 5;
  if (warp < 11) return;
  const int lane = threadIdx.x & 31, u = warp - 11, np = cs.np, ld = cs.ld, T = np >> 3;
  double* const ys = cs.S + (size_t)np * ld;
  const double* const S = cs.S;
  double* const xbuf = cs.dinv;
  const int i = 32 * u + lane;
  const bool mine = i < np;
  double y = mine ? ys[i] : 0.0;
  const int k8 = lane & 7;
  for (int jb = T - 1; jb >= 0; jb--) {
    const int j0 = 8 * jb;
    const bool above = i < j0;   // this row still receives the block's contribution
    double lv[8];
    {
      const double* col = S + (size_t)j0 * ld + (above ? i : 0);
#pragma unroll
      for (int k = 0; k < 8; k++) lv[k] = above ? col[k * ld] : 0.0;
    }
    if (u == (j0 >> 5)) {
      const double* Li = cs.Linv + jb * 64;
      const int l0 = j0 & 31;
      double sa = 0, sb = 0;
#pragma unroll
      for (int m = 0; m < 8; m += 2) {   // x_k = sum_{m >= k} Linv[m][k] y_m; Linv is stored with its zeros
        const double y0 = __shfl_sync(0xffffffffu, y, l0 + m), y1 = __shfl_sync(0xffffffffu, y, l0 + m + 1);
        sa = fma(Li[m * 8 + k8], y0, sa);
        sb = fma(Li[(m + 1) * 8 + k8], y1, sb);
      }
      const double xk = sa + sb;
      if (lane < 8) xbuf[8 * (jb & 1) + lane] = xk;
      if (lane >= l0 && lane < l0 + 8) y = xk;   // l0 is a multiple of 8: lane l0 + k formed unknown k
    }
    asm volatile("bar.sync %0, 160;" ::"n"(BC_BAR_BACK) : "memory");
    const double* xb = xbuf + 8 * (jb & 1);
    double ta = 0, tb = 0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      ta = fma(lv[k], xb[k], ta);
      tb = fma(lv[k + 1], xb[k + 1], tb);
    }
    y -= ta + tb;
  }
  if (mine) ys[i] = y;
}
