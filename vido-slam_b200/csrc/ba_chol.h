// ba_chol.h -- dense Cholesky solve of the reduced camera system of the window BA, in shared memory, by ONE CTA of 512
// threads (device code shared by ba_window.cu and tools/chol_probe2.cu).
//
// Replaces LinearSolverCSparse::solve on the sliding-window graph (3rdparty/g2o/g2o/solvers/linear_solver_csparse.h:108-141):
// the reference factors the full (poses + points) H with CSparse; here the points are eliminated first (Schur complement)
// and the reduced 6W x 6W system is factored densely.  Same linear system, same positive-definiteness test (all pivots > 0).
//
// Structure (right-looking LL^T, block size 8 = one FP64 tensor-core tile):
//   * the trailing matrix lives in REGISTERS for the whole factorisation: every 8x8 tile is owned by one of the 15 "bulk"
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
#define BC_BULK_WARPS 15
#define BC_MAX_SLOTS 11  // tiles per bulk warp at BC_MAXT: ceil(152 / 15)

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
// tk: optional cycle counters of thread 0 (chain warp) and thread 32 (bulk) for the probe: [0] diag0, [1] chain work per
// step, [2] chain waiting at the step barrier, [3] bulk panel, [4] bulk trailing, [5] bulk waiting.
// The chain warp and the bulk warps run two separate loops (warp specialisation: the register files of the two roles --
// the 8x8 block being factored vs. the accumulator tiles -- never coexist) that meet at two named barriers per step.
#define BC_BAR_PANEL 1
#define BC_BAR_STEP 2
__device__ __forceinline__ void bc_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(BC_THREADS) : "memory"); }
__device__ __forceinline__ void bc_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(BC_THREADS) : "memory"); }

__device__ __forceinline__ void bc_factor(const CholSm cs, int* s_bad, long long* tk) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = cs.np, ld = cs.ld, T = np >> 3;
  double* const S = cs.S;
  double* const Lp = cs.Lp;
  if (tid == 0) *s_bad = 0;
  if (warp == 0) {
    // =========================== chain warp ===========================
    __syncwarp();
    long long t0 = 0;
    if (tk) t0 = clock64();
    if (lane == 0 && !bc_chol8(S, ld, cs.dinv)) *s_bad = 1;
    bc_bar_sync(BC_BAR_STEP);   // the bulk warps have loaded their tiles, block 0 is factored
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
      if (more) {
        // update of the diagonal block jb+1: 36 entries of the lower triangle over the 32 lanes (lanes 0..3 take a second
        // one); 8-term dot as a tree
        for (int e = lane; e < 36; e += 32) {
          const int r = (e >= 1) + (e >= 3) + (e >= 6) + (e >= 10) + (e >= 15) + (e >= 21) + (e >= 28), c = e - ((r * (r + 1)) >> 1);
          const double* lr = Lp + (j0 + 8 + r) * BC_LPS;
          const double* lc = Lp + (j0 + 8 + c) * BC_LPS;
          const double s01 = fma(lr[1], lc[1], lr[0] * lc[0]), s23 = fma(lr[3], lc[3], lr[2] * lc[2]);
          const double s45 = fma(lr[5], lc[5], lr[4] * lc[4]), s67 = fma(lr[7], lc[7], lr[6] * lc[6]);
          double* dst = S + (j0 + 8 + r) * ld + j0 + 8 + c;
          *dst = *dst - ((s01 + s23) + (s45 + s67));
        }
        __syncwarp();
        if (lane == 0 && !bc_chol8(S + (j0 + 8) * ld + j0 + 8, ld, cs.dinv + j0 + 8)) *s_bad = 1;
      }
      if (tk && lane == 0) { const long long t1 = clock64(); tk[1] += t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_STEP);
      if (tk && lane == 0) { const long long t1 = clock64(); tk[2] += t1 - t0; t0 = t1; }
    }
  } else {
    // =========================== bulk warps ===========================
    // tile ownership: tiles (R, C), R >= C >= 1, without (1,1), numbered column by column and dealt round-robin, so that
    // the tiles still active at any step are spread evenly over the warps.  packed = R | C << 8, -1 = no tile.
    double c0[BC_MAX_SLOTS], c1[BC_MAX_SLOTS];
    int tRC[BC_MAX_SLOTS];
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int s = 0; s < BC_MAX_SLOTS; s++) {
      int t = (warp - 1) + BC_BULK_WARPS * s;
      int c = 1, R = -1;
      while (c < T) {
        const int cnt = (c == 1) ? T - 2 : T - c;
        if (t < cnt) { R = ((c == 1) ? 2 : c) + t; break; }
        t -= cnt; c++;
      }
      tRC[s] = (R >= 0) ? (R | (c << 8)) : -1;
      c0[s] = 0; c1[s] = 0;
      if (R >= 0) {
        const int row = 8 * R + fr, col = 8 * c + 2 * fk;
        c0[s] = (row >= col) ? S[row * ld + col] : S[col * ld + row];
        c1[s] = (row >= col + 1) ? S[row * ld + col + 1] : S[(col + 1) * ld + row];
      }
    }
    bc_bar_sync(BC_BAR_STEP);
    long long t0 = 0;
    if (tk && tid == 32) t0 = clock64();
    for (int jb = 0; jb < T; jb++) {
      if (*s_bad) break;
      const int j0 = 8 * jb;
      const double* D = S + j0 * ld + j0;
      const double* dv = cs.dinv + j0;
      const bool more = jb + 1 < T;
      // ---- panel rows of the blocks >= jb+2 and the right-hand side row
      {
        const int b = tid - 32, i = j0 + 16 + b;
        if (i < np) bc_trsm_row(D, ld, dv, S + i * ld + j0, Lp + i * BC_LPS);
        else if (b == BC_THREADS - 33) bc_trsm_row(D, ld, dv, S + np * ld + j0, nullptr);
      }
      if (warp == BC_BULK_WARPS && lane >= 8 && lane < 16) {
        // inverse of the diagonal factor (column c by one lane) for the back substitution: X = L11^-1
        const int c = lane - 8;
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
      if (tk && tid == 32) { const long long t1 = clock64(); tk[3] += t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_PANEL);
      if (more) {
        // ---- trailing update of the owned tiles: C -= L(R,jb) L(C,jb)^T as two DMMA (k = 0..3, 4..7)
#pragma unroll
        for (int s = 0; s < BC_MAX_SLOTS; s++) {
          const int R = tRC[s] & 255, C = tRC[s] >> 8;
          if (tRC[s] < 0 || C <= jb || (R == C && C == jb + 1)) continue;
          const double* pa = Lp + (8 * R + fr) * BC_LPS + fk;
          const double* pb = Lp + (8 * C + fr) * BC_LPS + fk;
          const double a0 = -pa[0], a1 = -pa[4], b0 = pb[0], b1 = pb[4];
          bc_dmma(c0[s], c1[s], a0, b0);
          bc_dmma(c0[s], c1[s], a1, b1);
        }
        if (warp == BC_BULK_WARPS) {  // right-hand side row: y(c) -= L(rhs, jb) . L(c, jb)
          const double* lr = S + np * ld + j0;
          const double r0 = lr[0], r1 = lr[1], r2 = lr[2], r3 = lr[3], r4 = lr[4], r5 = lr[5], r6 = lr[6], r7 = lr[7];
          for (int c = j0 + 8 + lane; c < np; c += 32) {
            const double* lc = Lp + c * BC_LPS;
            const double s01 = fma(r1, lc[1], r0 * lc[0]), s23 = fma(r3, lc[3], r2 * lc[2]);
            const double s45 = fma(r5, lc[5], r4 * lc[4]), s67 = fma(r7, lc[7], r6 * lc[6]);
            S[np * ld + c] -= (s01 + s23) + (s45 + s67);
          }
        }
        // ---- tiles that are final now go back to shared memory: block column jb+1 (next panel) and the diagonal tile jb+2
#pragma unroll
        for (int s = 0; s < BC_MAX_SLOTS; s++) {
          const int R = tRC[s] & 255, C = tRC[s] >> 8;
          if (tRC[s] < 0) continue;
          if ((C == jb + 1 && R > C) || (R == C && C == jb + 2)) {
            const int row = 8 * R + fr, col = 8 * C + 2 * fk;
            S[row * ld + col] = c0[s];
            S[row * ld + col + 1] = c1[s];
          }
        }
      }
      if (tk && tid == 32) { const long long t1 = clock64(); tk[4] += t1 - t0; t0 = t1; }
      bc_bar_sync(BC_BAR_STEP);
      if (tk && tid == 32) { const long long t1 = clock64(); tk[5] += t1 - t0; t0 = t1; }
    }
  }
  __syncthreads();
}

// Back substitution L^T x = y by warp 0 (the other warps return at once): x overwrites row np.  Per block step the 8
// unknowns are a matrix-vector product with the precomputed inverse of the diagonal factor; every lane then updates its
// rows above the block.
__device__ __forceinline__ void bc_backsolve(const CholSm cs) {
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x, np = cs.np, ld = cs.ld, T = np >> 3;
  double* const ys = cs.S + (size_t)np * ld;
  const double* const S = cs.S;
  for (int jb = T - 1; jb >= 0; jb--) {
    const int j0 = 8 * jb;
    const double* Li = cs.Linv + jb * 64;
    // rows of L of this block, at the columns this lane owns (issued before the solve: independent of it)
    constexpr int NCH = (BC_MAXT * 8 + 31) / 32;
    double x[8];
    {
      double y[8];
#pragma unroll
      for (int m = 0; m < 8; m++) y[m] = ys[j0 + m];
      // x = L11^-T y: x_k = sum_{m >= k} Linv[m][k] y_m
#pragma unroll
      for (int k = 0; k < 8; k++) {
        double s = 0;
#pragma unroll
        for (int m = k; m < 8; m++) s = fma(Li[m * 8 + k], y[m], s);
        x[k] = s;
      }
    }
#pragma unroll
    for (int u = 0; u < NCH; u++) {
      const int i = lane + 32 * u;
      if (i < j0) {
        double t = ys[i];
#pragma unroll
        for (int k = 0; k < 8; k++) t = fma(-S[(j0 + k) * ld + i], x[k], t);
        ys[i] = t;
      }
    }
    if (lane < 8) {
      double xv = x[0];
#pragma unroll
      for (int k = 1; k < 8; k++) xv = (lane == k) ? x[k] : xv;
      ys[j0 + lane] = xv;
    }
    __syncwarp();
  }
}
