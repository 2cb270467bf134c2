// Stand-alone timing harness for the shared-memory blocked Cholesky of the window BA (debug tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chol_probe tools/chol_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// 1/sqrt(d) for a positive normal d: the library fast path (MUFU.RSQ64H + one third-order correction) without its
// special-case subroutine -- a CALL inside a latency-critical chain makes the compiler park live values in local memory
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(d, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// 6x6 Cholesky of a diagonal block by one lane, straight-line scalar code (no local arrays, no subroutine calls: either
// would put local-memory round trips into this latency-critical chain).  Writes L in place and 1/L_jj to dinv.
__device__ __forceinline__ bool chol_diag6(double* Ls, int ld, int j0, double* dinv) {
  double* r0 = Ls + (size_t)j0 * ld + j0;
  double *r1 = r0 + ld, *r2 = r1 + ld, *r3 = r2 + ld, *r4 = r3 + ld, *r5 = r4 + ld;
  const double a00 = r0[0];
  const double a10 = r1[0], a11 = r1[1];
  const double a20 = r2[0], a21 = r2[1], a22 = r2[2];
  const double a30 = r3[0], a31 = r3[1], a32 = r3[2], a33 = r3[3];
  const double a40 = r4[0], a41 = r4[1], a42 = r4[2], a43 = r4[3], a44 = r4[4];
  const double a50 = r5[0], a51 = r5[1], a52 = r5[2], a53 = r5[3], a54 = r5[4], a55 = r5[5];
  const double i0 = rsqrt_pos(a00), l00 = a00 * i0;
  const double l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0, l40 = a40 * i0, l50 = a50 * i0;
  const double d1 = a11 - l10 * l10;
  const double i1 = rsqrt_pos(d1), l11 = d1 * i1;
  const double l21 = (a21 - l20 * l10) * i1, l31 = (a31 - l30 * l10) * i1, l41 = (a41 - l40 * l10) * i1, l51 = (a51 - l50 * l10) * i1;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  const double i2 = rsqrt_pos(d2), l22 = d2 * i2;
  const double l32 = (a32 - l30 * l20 - l31 * l21) * i2, l42 = (a42 - l40 * l20 - l41 * l21) * i2, l52 = (a52 - l50 * l20 - l51 * l21) * i2;
  const double d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
  const double i3 = rsqrt_pos(d3), l33 = d3 * i3;
  const double l43 = (a43 - l40 * l30 - l41 * l31 - l42 * l32) * i3, l53 = (a53 - l50 * l30 - l51 * l31 - l52 * l32) * i3;
  const double d4 = a44 - l40 * l40 - l41 * l41 - l42 * l42 - l43 * l43;
  const double i4 = rsqrt_pos(d4), l44 = d4 * i4;
  const double l54 = (a54 - l50 * l40 - l51 * l41 - l52 * l42 - l53 * l43) * i4;
  const double d5 = a55 - l50 * l50 - l51 * l51 - l52 * l52 - l53 * l53 - l54 * l54;
  const double i5 = rsqrt_pos(d5), l55 = d5 * i5;
  if (!(a00 > 0 && d1 > 0 && d2 > 0 && d3 > 0 && d4 > 0 && d5 > 0)) return false;
  r0[0] = l00;
  r1[0] = l10; r1[1] = l11;
  r2[0] = l20; r2[1] = l21; r2[2] = l22;
  r3[0] = l30; r3[1] = l31; r3[2] = l32; r3[3] = l33;
  r4[0] = l40; r4[1] = l41; r4[2] = l42; r4[3] = l43; r4[4] = l44;
  r5[0] = l50; r5[1] = l51; r5[2] = l52; r5[3] = l53; r5[4] = l54; r5[5] = l55;
  dinv[0] = i0; dinv[1] = i1; dinv[2] = i2; dinv[3] = i3; dinv[4] = i4; dinv[5] = i5;
  return true;
}

// variant 0: as in ba_kernels.cu (look-ahead, warp 0 factors next diag);  variant 1: the 6x6 factor is done by 6 lanes
// column-parallel with shuffles (no local arrays)
template <int VARIANT>
__global__ void __launch_bounds__(256, 1) chol_kernel(const double* S, const double* b, double* x, int nb, long long* cyc, int reps) {
  extern __shared__ double Ls[];
  const int n = 6 * nb, ld = n + 1, tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  double* Li = Ls + (size_t)(n + 1) * ld;
  __shared__ int s_bad;
  long long c_diag = 0, c_panel = 0, c_trail = 0, c_back = 0, c_tot = 0, c_w = 0;
  for (int rep = 0; rep < reps; rep++) {
    for (int i = tid; i < n * n; i += nt) { const int r = i / n, c = i - r * n; if (c <= r) Ls[r * ld + c] = S[i]; }
    double* ys = Ls + (size_t)n * ld;
    for (int i = tid; i < n; i += nt) ys[i] = b[i];
    if (tid == 0) s_bad = 0;
    __syncthreads();
    long long t0 = clock64(), t1;
    const long long tstart = t0;
    if (tid == 0 && !chol_diag6(Ls, ld, 0, Li)) s_bad = 1;
    __syncthreads();
    for (int jb = 0; jb < nb; jb++) {
      if (s_bad) break;
      const int j0 = 6 * jb;
      t0 = clock64();
      {
        const double* D = Ls + j0 * ld + j0;
        const double* dv = Li + 6 * jb;
        const double l10 = D[ld], l20 = D[2 * ld], l21 = D[2 * ld + 1], l30 = D[3 * ld], l31 = D[3 * ld + 1], l32 = D[3 * ld + 2];
        const double l40 = D[4 * ld], l41 = D[4 * ld + 1], l42 = D[4 * ld + 2], l43 = D[4 * ld + 3];
        const double l50 = D[5 * ld], l51 = D[5 * ld + 1], l52 = D[5 * ld + 2], l53 = D[5 * ld + 3], l54 = D[5 * ld + 4];
        const double d0 = dv[0], d1 = dv[1], d2 = dv[2], d3 = dv[3], d4 = dv[4], d5 = dv[5];
        for (int i = j0 + 6 + tid; i <= n; i += nt) {
          double* row = Ls + i * ld + j0;
          const double o0 = row[0] * d0;
          const double o1 = (row[1] - o0 * l10) * d1;
          const double o2 = (row[2] - o0 * l20 - o1 * l21) * d2;
          const double o3 = (row[3] - o0 * l30 - o1 * l31 - o2 * l32) * d3;
          const double o4 = (row[4] - o0 * l40 - o1 * l41 - o2 * l42 - o3 * l43) * d4;
          const double o5 = (row[5] - o0 * l50 - o1 * l51 - o2 * l52 - o3 * l53 - o4 * l54) * d5;
          row[0] = o0; row[1] = o1; row[2] = o2; row[3] = o3; row[4] = o4; row[5] = o5;
        }
      }
      __syncthreads();
      t1 = clock64(); c_panel += t1 - t0; t0 = t1;
      const int m = n - j0 - 6;
      if (warp == 0) {
        if (jb + 1 < nb) {
          if ((VARIANT & 16) && lane < 21) {
            // row of lane l in the packed lower triangle: 3 bits per lane in one constant; the 6-term dot as a tree
            const int r = (int)((0x5b6db2491b6d2448ull >> (3 * lane)) & 7ull), c = lane - ((r * (r + 1)) >> 1);
            const double* lr = Ls + (j0 + 6 + r) * ld + j0;
            const double* lc = Ls + (j0 + 6 + c) * ld + j0;
            double* dst = Ls + (j0 + 6 + r) * ld + j0 + 6 + c;
            const double s01 = fma(lr[1], lc[1], lr[0] * lc[0]), s23 = fma(lr[3], lc[3], lr[2] * lc[2]), s45 = fma(lr[5], lc[5], lr[4] * lc[4]);
            *dst = *dst - ((s01 + s23) + s45);
          }
          if (!(VARIANT & 16) && lane < 21) {
            int r = 0, c = lane;
            while (c > r) { c -= r + 1; r++; }
            const double* lr = Ls + (j0 + 6 + r) * ld + j0;
            const double* lc = Ls + (j0 + 6 + c) * ld + j0;
            Ls[(j0 + 6 + r) * ld + j0 + 6 + c] -= lr[0] * lc[0] + lr[1] * lc[1] + lr[2] * lc[2] + lr[3] * lc[3] + lr[4] * lc[4] + lr[5] * lc[5];
          }
          __syncwarp();
          if (!(VARIANT & 1)) {
            if (lane == 0 && !chol_diag6(Ls, ld, j0 + 6, Li + 6 * (jb + 1))) s_bad = 1;
          } else {
            // lane r (< 6) owns row r of the block; column j: pivot from lane j, broadcast by shuffle
            const int j1 = j0 + 6;
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0;
            if (lane < 6) {
              const double* row = Ls + (j1 + lane) * ld + j1;
              a0 = row[0]; if (lane >= 1) a1 = row[1]; if (lane >= 2) a2 = row[2]; if (lane >= 3) a3 = row[3]; if (lane >= 4) a4 = row[4]; if (lane >= 5) a5 = row[5];
            }
            bool ok = true;
            double inv, p;
#define STEP(J, AJ)                                                        \
            p = __shfl_sync(0xffffffffu, AJ, J);                             \
            ok = ok && (p > 0);                                              \
            inv = rsqrt_pos(p);                                                \
            AJ = (lane == J) ? p * inv : AJ * inv;                           \
            if (lane == J) Li[6 * (jb + 1) + J] = inv;
            STEP(0, a0)
            { const double l1 = __shfl_sync(0xffffffffu, a0, 1), l2 = __shfl_sync(0xffffffffu, a0, 2), l3 = __shfl_sync(0xffffffffu, a0, 3), l4 = __shfl_sync(0xffffffffu, a0, 4), l5 = __shfl_sync(0xffffffffu, a0, 5);
              a1 -= a0 * l1; a2 -= a0 * l2; a3 -= a0 * l3; a4 -= a0 * l4; a5 -= a0 * l5; }
            STEP(1, a1)
            { const double l2 = __shfl_sync(0xffffffffu, a1, 2), l3 = __shfl_sync(0xffffffffu, a1, 3), l4 = __shfl_sync(0xffffffffu, a1, 4), l5 = __shfl_sync(0xffffffffu, a1, 5);
              a2 -= a1 * l2; a3 -= a1 * l3; a4 -= a1 * l4; a5 -= a1 * l5; }
            STEP(2, a2)
            { const double l3 = __shfl_sync(0xffffffffu, a2, 3), l4 = __shfl_sync(0xffffffffu, a2, 4), l5 = __shfl_sync(0xffffffffu, a2, 5);
              a3 -= a2 * l3; a4 -= a2 * l4; a5 -= a2 * l5; }
            STEP(3, a3)
            { const double l4 = __shfl_sync(0xffffffffu, a3, 4), l5 = __shfl_sync(0xffffffffu, a3, 5);
              a4 -= a3 * l4; a5 -= a3 * l5; }
            STEP(4, a4)
            { const double l5 = __shfl_sync(0xffffffffu, a4, 5);
              a5 -= a4 * l5; }
            STEP(5, a5)
#undef STEP
            if (lane < 6) {
              double* row = Ls + (j1 + lane) * ld + j1;
              row[0] = a0; if (lane >= 1) row[1] = a1; if (lane >= 2) row[2] = a2; if (lane >= 3) row[3] = a3; if (lane >= 4) row[4] = a4; if (lane >= 5) row[5] = a5;
            }
            if (!ok && lane == 0) s_bad = 1;
          }
        }
        t1 = clock64(); c_diag += t1 - t0; t0 = t1;
      } else {
        const long long tw0 = clock64();
        const int ngroups = (m - 5 + 7) >> 3;
        // VARIANT & 2: the warp that shares warp 0's scheduler (warp 4) stays idle, so the critical chain owns its FP64 pipe
        const int widx = (VARIANT & 2) ? (warp < 4 ? warp - 1 : warp - 2) : warp - 1;
        const int nwork = (VARIANT & 2) ? nwarp - 2 : nwarp - 1;
        const bool idle = (VARIANT & 2) && warp == 4;
        if (VARIANT & 64) {
          // 8-row x 32-column tiles dealt round-robin to the workers (the triangle makes the groups unequal: 1..4 tiles)
          int t = 0;
          for (int g = 0; g < ngroups && !idle; g++) {
            const int r0 = 6 + 8 * g;
            const int rmax = min(r0 + 7, m), cmax = min(rmax, m - 1);
            const int nch = (cmax >> 5) + 1;
            bool have = false;
            double rv[8][6];
            for (int ch = 0; ch < nch; ch++, t++) {
              if (t % nwork != widx) continue;
              if (!have) {
                have = true;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                  const int r = min(r0 + i, m);
                  const double* lr = Ls + (j0 + 6 + r) * ld + j0;
#pragma unroll
                  for (int k = 0; k < 6; k++) rv[i][k] = lr[k];
                }
              }
              const int c = lane + 32 * ch;
              if (c <= cmax) {
                const double* lc = Ls + (j0 + 6 + c) * ld + j0;
                const double l0 = lc[0], l1 = lc[1], l2 = lc[2], l3 = lc[3], l4 = lc[4], l5 = lc[5];
                double cv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) { const int r = r0 + i; cv[i] = (r <= m && c <= r) ? Ls[(j0 + 6 + r) * ld + j0 + 6 + c] : 0.0; }
#pragma unroll
                for (int i = 0; i < 8; i++) cv[i] -= rv[i][0] * l0 + rv[i][1] * l1 + rv[i][2] * l2 + rv[i][3] * l3 + rv[i][4] * l4 + rv[i][5] * l5;
#pragma unroll
                for (int i = 0; i < 8; i++) { const int r = r0 + i; if (r <= m && c <= r) Ls[(j0 + 6 + r) * ld + j0 + 6 + c] = cv[i]; }
              }
            }
          }
        } else
        for (int pass = 0; pass * nwork < ngroups && !idle; pass++) {
          int g;
          if (VARIANT & 128) {   // heaviest groups first, alternating direction over the workers
            const int gp = pass * nwork + ((pass & 1) ? nwork - 1 - widx : widx);
            if (gp >= ngroups) continue;
            g = ngroups - 1 - gp;
          } else {
            g = widx + pass * nwork;
            if (g >= ngroups) continue;
          }
          const int r0 = 6 + 8 * g;
          double rv[8][6];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int r = min(r0 + i, m);
            const double* lr = Ls + (j0 + 6 + r) * ld + j0;
#pragma unroll
            for (int k = 0; k < 6; k++) rv[i][k] = lr[k];
          }
          const int rmax = min(r0 + 7, m), cmax = min(rmax, m - 1);
          for (int c = lane; c <= cmax; c += 32) {
            const double* lc = Ls + (j0 + 6 + c) * ld + j0;
            const double l0 = lc[0], l1 = lc[1], l2 = lc[2], l3 = lc[3], l4 = lc[4], l5 = lc[5];
            double cv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { const int r = r0 + i; cv[i] = (r <= m && c <= r) ? Ls[(j0 + 6 + r) * ld + j0 + 6 + c] : 0.0; }
#pragma unroll
            for (int i = 0; i < 8; i++) cv[i] -= rv[i][0] * l0 + rv[i][1] * l1 + rv[i][2] * l2 + rv[i][3] * l3 + rv[i][4] * l4 + rv[i][5] * l5;
#pragma unroll
            for (int i = 0; i < 8; i++) { const int r = r0 + i; if (r <= m && c <= r) Ls[(j0 + 6 + r) * ld + j0 + 6 + c] = cv[i]; }
          }
        }
        c_w += clock64() - tw0;
      }
      __syncthreads();
      t1 = clock64(); c_trail += t1 - t0; t0 = t1;
    }
    __syncthreads();
    t0 = clock64();
    if ((VARIANT & 8) && !s_bad) {
      // inverse of every diagonal block (lower triangular 6x6), one thread per block: Linv stored row-major in Lv[36*jb]
      double* Lv = Li + 6 * nb;
      if (tid < nb) {
        const double* D = Ls + (6 * tid) * ld + 6 * tid;
        const double* dv = Li + 6 * tid;
        double* V = Lv + 36 * tid;
        for (int c = 0; c < 6; c++) {
          V[7 * c] = dv[c];
          for (int r = c + 1; r < 6; r++) {
            double sacc = 0;
            for (int k = c; k < r; k++) sacc += D[r * ld + k] * V[6 * k + c];
            V[6 * r + c] = -sacc * dv[r];
          }
        }
      }
      __syncthreads();
    }
    if (!s_bad && warp == 0 && !(VARIANT & (12 | 32))) {
      for (int jb = nb - 1; jb >= 0; jb--) {
        const int j0 = 6 * jb;
        const double* D = Ls + j0 * ld + j0;
        const double* dv = Li + 6 * jb;
        double x0, x1, x2, x3, x4, x5;
        {
          const double y0 = ys[j0], y1 = ys[j0 + 1], y2 = ys[j0 + 2], y3 = ys[j0 + 3], y4 = ys[j0 + 4], y5 = ys[j0 + 5];
          x5 = y5 * dv[5];
          x4 = (y4 - D[5 * ld + 4] * x5) * dv[4];
          x3 = (y3 - D[5 * ld + 3] * x5 - D[4 * ld + 3] * x4) * dv[3];
          x2 = (y2 - D[5 * ld + 2] * x5 - D[4 * ld + 2] * x4 - D[3 * ld + 2] * x3) * dv[2];
          x1 = (y1 - D[5 * ld + 1] * x5 - D[4 * ld + 1] * x4 - D[3 * ld + 1] * x3 - D[2 * ld + 1] * x2) * dv[1];
          x0 = (y0 - D[5 * ld] * x5 - D[4 * ld] * x4 - D[3 * ld] * x3 - D[2 * ld] * x2 - D[ld] * x1) * dv[0];
        }
        if (lane < 6) ys[j0 + lane] = lane == 0 ? x0 : lane == 1 ? x1 : lane == 2 ? x2 : lane == 3 ? x3 : lane == 4 ? x4 : x5;
#pragma unroll 5
        for (int i = lane; i < j0; i += 32)
          ys[i] -= D[i - j0] * x0 + D[ld + i - j0] * x1 + D[2 * ld + i - j0] * x2 + D[3 * ld + i - j0] * x3 + D[4 * ld + i - j0] * x4 + D[5 * ld + i - j0] * x5;
        __syncwarp();
      }
    }
    if (!s_bad && warp == 0 && (VARIANT & 32)) {
      // every load of a block step is issued before the 6-step solve (the stores of the previous step forbid the compiler to
      // hoist them itself), the rows above are updated as independent chains that end with the unknown known last
      for (int jb = nb - 1; jb >= 0; jb--) {
        const int j0 = 6 * jb;
        const double* D = Ls + j0 * ld + j0;
        const double* dv = Li + 6 * jb;
        const double y0 = ys[j0], y1 = ys[j0 + 1], y2 = ys[j0 + 2], y3 = ys[j0 + 3], y4 = ys[j0 + 4], y5 = ys[j0 + 5];
        const double d0 = dv[0], d1 = dv[1], d2 = dv[2], d3 = dv[3], d4 = dv[4], d5 = dv[5];
        const double l10 = D[ld], l20 = D[2 * ld], l21 = D[2 * ld + 1], l30 = D[3 * ld], l31 = D[3 * ld + 1], l32 = D[3 * ld + 2];
        const double l40 = D[4 * ld], l41 = D[4 * ld + 1], l42 = D[4 * ld + 2], l43 = D[4 * ld + 3];
        const double l50 = D[5 * ld], l51 = D[5 * ld + 1], l52 = D[5 * ld + 2], l53 = D[5 * ld + 3], l54 = D[5 * ld + 4];
        double dl[4][6], yv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = lane + 32 * u;
          const bool on = i < j0;
          yv[u] = on ? ys[i] : 0.0;
#pragma unroll
          for (int k = 0; k < 6; k++) dl[u][k] = on ? D[k * ld + i - j0] : 0.0;
        }
        const double x5 = y5 * d5;
        const double x4 = (y4 - l54 * x5) * d4;
        const double x3 = ((y3 - l53 * x5) - l43 * x4) * d3;
        const double x2 = (((y2 - l52 * x5) - l42 * x4) - l32 * x3) * d2;
        const double x1 = ((((y1 - l51 * x5) - l41 * x4) - l31 * x3) - l21 * x2) * d1;
        const double x0 = (((((y0 - l50 * x5) - l40 * x4) - l30 * x3) - l20 * x2) - l10 * x1) * d0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = lane + 32 * u;
          double t = yv[u];
          t = fma(-dl[u][5], x5, t); t = fma(-dl[u][4], x4, t); t = fma(-dl[u][3], x3, t);
          t = fma(-dl[u][2], x2, t); t = fma(-dl[u][1], x1, t); t = fma(-dl[u][0], x0, t);
          if (i < j0) ys[i] = t;
        }
        if (lane < 6) {
          double xv = x0;
          xv = lane == 1 ? x1 : xv; xv = lane == 2 ? x2 : xv; xv = lane == 3 ? x3 : xv; xv = lane == 4 ? x4 : xv; xv = lane == 5 ? x5 : xv;
          ys[j0 + lane] = xv;
        }
        __syncwarp();
      }
    }
    if (!s_bad && warp == 0 && (VARIANT & 12)) {
      // y kept in registers: lane l owns rows l, l+32, l+64, l+96; the 6 entries of the current block are fetched by shuffles,
      // every lane solves the block redundantly; the update of a lane's rows starts with the unknown that is known first
      double yv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { const int i = lane + 32 * u; yv[u] = (i < n) ? ys[i] : 0.0; }
      const double* Lv = Li + 6 * nb;
      for (int jb = nb - 1; jb >= 0; jb--) {
        const int j0 = 6 * jb;
        const double* D = Ls + j0 * ld + j0;
        const double* dv = Li + 6 * jb;
        // loads that do not depend on the previous block: issue first
        double dl[4][6];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = lane + 32 * u;
#pragma unroll
          for (int k = 0; k < 6; k++) dl[u][k] = (i < j0) ? D[k * ld + i - j0] : 0.0;
        }
        double yb[6];
#pragma unroll
        for (int k = 0; k < 6; k++) {
          const int i = j0 + k, u = i >> 5;
          const double v = u == 0 ? yv[0] : u == 1 ? yv[1] : u == 2 ? yv[2] : yv[3];
          yb[k] = __shfl_sync(0xffffffffu, v, i & 31);
        }
        double x0, x1, x2, x3, x4, x5;
        if (VARIANT & 8) {
          const double* V = Lv + 36 * jb;   // x = Linv^T y: x_c = sum_{r >= c} V[r][c] y_r, independent dot products
          x5 = V[35] * yb[5];
          x4 = V[28] * yb[4] + V[34] * yb[5];
          x3 = (V[21] * yb[3] + V[27] * yb[4]) + V[33] * yb[5];
          x2 = (V[14] * yb[2] + V[20] * yb[3]) + (V[26] * yb[4] + V[32] * yb[5]);
          x1 = (V[7] * yb[1] + V[13] * yb[2]) + (V[19] * yb[3] + V[25] * yb[4]) + V[31] * yb[5];
          x0 = (V[0] * yb[0] + V[6] * yb[1]) + (V[12] * yb[2] + V[18] * yb[3]) + (V[24] * yb[4] + V[30] * yb[5]);
        } else {
          x5 = yb[5] * dv[5];
          x4 = (yb[4] - D[5 * ld + 4] * x5) * dv[4];
          x3 = (yb[3] - D[5 * ld + 3] * x5 - D[4 * ld + 3] * x4) * dv[3];
          x2 = (yb[2] - D[5 * ld + 2] * x5 - D[4 * ld + 2] * x4 - D[3 * ld + 2] * x3) * dv[2];
          x1 = (yb[1] - D[5 * ld + 1] * x5 - D[4 * ld + 1] * x4 - D[3 * ld + 1] * x3 - D[2 * ld + 1] * x2) * dv[1];
          x0 = (yb[0] - D[5 * ld] * x5 - D[4 * ld] * x4 - D[3 * ld] * x3 - D[2 * ld] * x2 - D[ld] * x1) * dv[0];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = lane + 32 * u;
          double t = yv[u];
          if (VARIANT & 8) t -= ((dl[u][5] * x5 + dl[u][4] * x4) + (dl[u][3] * x3 + dl[u][2] * x2)) + (dl[u][1] * x1 + dl[u][0] * x0);
          else { t -= dl[u][5] * x5; t -= dl[u][4] * x4; t -= dl[u][3] * x3; t -= dl[u][2] * x2; t -= dl[u][1] * x1; t -= dl[u][0] * x0; }
          if (i >= j0 && i < j0 + 6) { const int k = i - j0; t = k == 0 ? x0 : k == 1 ? x1 : k == 2 ? x2 : k == 3 ? x3 : k == 4 ? x4 : x5; }
          yv[u] = t;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) { const int i = lane + 32 * u; if (i < n) ys[i] = yv[u]; }
    }
    __syncthreads();
    t1 = clock64(); c_back += t1 - t0;
    c_tot += t1 - tstart;
    for (int i = tid; i < n; i += nt) x[i] = ys[i];
    __syncthreads();
  }
  if (tid == 32) cyc[9] = c_w / reps;
  if (tid == 224) cyc[10] = c_w / reps;
  if (tid == 0) { cyc[0] = c_tot / reps; cyc[1] = c_panel / reps; cyc[2] = c_diag / reps; cyc[3] = c_trail / reps; cyc[4] = c_back / reps; }
  // isolated pieces, no other warp active
  __syncthreads();
  if (warp == 0) {
    const int ld2 = 7;
    double* B = Ls;  // 6x6 block, ld 7
    long long acc1 = 0, acc2 = 0, acc3 = 0;
    for (int rep = 0; rep < 32; rep++) {
      if (lane == 0) for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++) B[r * ld2 + c] = (r == c) ? 50.0 + r : 0.5 + 0.1 * (r + c);
      __syncwarp();
      long long t0 = clock64();
      bool okk = true;
      if (lane == 0) okk = chol_diag6(B, ld2, 0, Li);
      __syncwarp();
      long long t1 = clock64();
      acc1 += t1 - t0;
      // dependent chain of 6 rsqrt + 6 mul only
      double v = B[0] + (okk ? 0.0 : 1.0);
      t0 = clock64();
#pragma unroll
      for (int k = 0; k < 6; k++) { const double iv = rsqrt(v); v = fma(v, iv, 40.0); }
      t1 = clock64();
      acc2 += t1 - t0;
      B[40] = v;
      // 21 LDS + 21 STS by one lane
      t0 = clock64();
      if (lane == 0) { double s2 = 0; for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++) s2 += B[r * ld2 + c]; for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++) B[r * ld2 + c] = s2 + r; }
      __syncwarp();
      t1 = clock64();
      acc3 += t1 - t0;
    }
    if (tid == 0) { cyc[6] = acc1 / 32; cyc[7] = acc2 / 32; cyc[8] = acc3 / 32; }
  }
  __syncthreads();
  if (tid == 0) {  // cost of a globaltimer read
    long long t0 = clock64();
    unsigned long long s = 0;
    for (int i = 0; i < 64; i++) s += gtime();
    long long t1 = clock64();
    cyc[5] = (t1 - t0) / 64 + (s == 1);
  }
}

int main() {
  const int nb = 20, n = 6 * nb;
  std::vector<double> A(n * n), S(n * n, 0.0), b(n), xr(n);
  srand(1);
  for (auto& v : A) v = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += A[i * n + k] * A[j * n + k]; S[i * n + j] = s + (i == j ? n : 0); }
  for (int i = 0; i < n; i++) b[i] = rand() / (double)RAND_MAX;
  double *dS, *db, *dx; long long* cyc;
  cudaMalloc(&dS, n * n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dx, n * 8); cudaMallocManaged(&cyc, 128);
  cudaMemcpy(dS, S.data(), n * n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
  const size_t smem = sizeof(double) * ((size_t)(n + 1) * (n + 1) + 48 * nb);
  const int variants[] = {0, 48, 128, 176, 178};
  for (int vi = 0; vi < 5; vi++) {
    const int variant = variants[vi];
#define RUNV(V) case V: cudaFuncSetAttribute(chol_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); chol_kernel<V><<<1, 256, smem>>>(dS, db, dx, nb, cyc, 20); break;
    switch (variant) { RUNV(0) RUNV(48) RUNV(128) RUNV(176) RUNV(178) }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(xr.data(), dx, n * 8, cudaMemcpyDeviceToHost);
    double res = 0;
    for (int i = 0; i < n; i++) { double s = -b[i]; for (int j = 0; j < n; j++) s += S[i * n + j] * xr[j]; res = fmax(res, fabs(s)); }
    printf("variant %d: total=%lld cycles (panel=%lld diag(warp0)=%lld trail+wait=%lld backsub=%lld) residual=%.3e trailing own: warp1=%lld warp7=%lld | isolated: chol_diag6=%lld\n",
           variant, cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], res, cyc[9], cyc[10], cyc[6]);
  }
  return 0;
}
