// pnp_kernels.cu -- initial camera / object model: PnP-RANSAC hypotheses scored in parallel + the constant-velocity
// alternative, winner = more inliers.
//
// Replaces Tracking::GetInitModelCam (src/Tracking.cc:1914-2028) and GetInitModelObj (:2030-2162).  The reference
// delegates the RANSAC to cv::solvePnPRansac(500 its, 0.4 px, 0.98, SOLVEPNP_P3P) of un-vendored OpenCV, whose RNG and
// minimal solver cannot be matched bit for bit; the deterministic variant specified in oracle/vido_oracle.h is
// implemented here: counter-based 4-point samples, Gauss-Newton minimal solve from the motion-model pose, OpenCV's
// adaptive iteration count replayed sequentially over the precomputed scores, refit on the consensus set.
// All `iters` hypotheses are evaluated concurrently (one warp each); the sequential "best so far / shrinking
// iteration budget" logic is replayed by one thread, which gives exactly the sequential algorithm's answer.
#include <cstring>

#include "ctx.h"

#define PNP_THREADS 256

struct PnpPose {
  double R[9], t[3];
};

struct PnpArgs {
  int n, M, iters, no_mm;   // no_mm: GetInitModelObj without a previous object motion (the RANSAC model always wins)
  const float* cur_xy;   // [n][2]
  const float* pts3d;    // [n][3]
  const int* good;       // [M] indices with valid depth
  const float* Tcw_motion;  // [16]
  float fx, fy, cx, cy, thr, confidence;
  // scratch
  PnpPose* hyp;   // [iters]
  int* hyp_cnt;   // [iters] (-1: degenerate sample)
  // outputs
  float* Tcw_out;   // [16]
  int* inlier_ids;  // [n]
  int* result;      // n_inliers, winner, ransac_inliers, mm_inliers
  int* tmp_ids;     // [n] scratch of the selection kernel
};

__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ void quat_to_R9(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}

// residual and Jacobian rows of one correspondence; returns false when the point is behind the camera
__device__ __forceinline__ bool pnp_row(const PnpPose& T, const float* X3, const float* uv, double fx, double fy, double cx,
                                        double cy, double* r, double* J0, double* J1) {
  const double X[3] = {X3[0], X3[1], X3[2]};
  double p[3];
#pragma unroll
  for (int k = 0; k < 3; k++) p[k] = T.R[3 * k] * X[0] + T.R[3 * k + 1] * X[1] + T.R[3 * k + 2] * X[2] + T.t[k];
  if (!(p[2] > 1e-9)) return false;
  const double iz = 1.0 / p[2], x = p[0] * iz, y = p[1] * iz;
  r[0] = fx * x + cx - uv[0];
  r[1] = fy * y + cy - uv[1];
  if (J0) {
    const double a00 = fx * iz, a02 = -fx * x * iz, a11 = fy * iz, a12 = -fy * y * iz;
    J0[0] = a02 * p[1]; J0[1] = a00 * p[2] - a02 * p[0]; J0[2] = -a00 * p[1]; J0[3] = a00; J0[4] = 0; J0[5] = a02;
    J1[0] = -a11 * p[2] + a12 * p[1]; J1[1] = -a12 * p[0]; J1[2] = a11 * p[0]; J1[3] = 0; J1[4] = a11; J1[5] = a12;
  }
  return true;
}

// H d = b by LDL^T, straight-line scalar code with exactly the operation order of the reference loops (the file is built
// with -fmad=false: results are bit-identical to the CPU restatement); false if a pivot is <= 1e-12
__device__ __forceinline__ bool ldlt6_solve(const double* H, const double* b, double* d) {
  double v0 = H[0];
  if (!(v0 > 1e-12)) return false;
  const double D0 = v0;
  double l10 = H[6];
  l10 = l10 / v0;
  double l20 = H[12];
  l20 = l20 / v0;
  double l30 = H[18];
  l30 = l30 / v0;
  double l40 = H[24];
  l40 = l40 / v0;
  double l50 = H[30];
  l50 = l50 / v0;
  double v1 = H[7];
  v1 -= l10 * l10 * D0;
  if (!(v1 > 1e-12)) return false;
  const double D1 = v1;
  double l21 = H[13];
  l21 -= l20 * l10 * D0;
  l21 = l21 / v1;
  double l31 = H[19];
  l31 -= l30 * l10 * D0;
  l31 = l31 / v1;
  double l41 = H[25];
  l41 -= l40 * l10 * D0;
  l41 = l41 / v1;
  double l51 = H[31];
  l51 -= l50 * l10 * D0;
  l51 = l51 / v1;
  double v2 = H[14];
  v2 -= l20 * l20 * D0;
  v2 -= l21 * l21 * D1;
  if (!(v2 > 1e-12)) return false;
  const double D2 = v2;
  double l32 = H[20];
  l32 -= l30 * l20 * D0;
  l32 -= l31 * l21 * D1;
  l32 = l32 / v2;
  double l42 = H[26];
  l42 -= l40 * l20 * D0;
  l42 -= l41 * l21 * D1;
  l42 = l42 / v2;
  double l52 = H[32];
  l52 -= l50 * l20 * D0;
  l52 -= l51 * l21 * D1;
  l52 = l52 / v2;
  double v3 = H[21];
  v3 -= l30 * l30 * D0;
  v3 -= l31 * l31 * D1;
  v3 -= l32 * l32 * D2;
  if (!(v3 > 1e-12)) return false;
  const double D3 = v3;
  double l43 = H[27];
  l43 -= l40 * l30 * D0;
  l43 -= l41 * l31 * D1;
  l43 -= l42 * l32 * D2;
  l43 = l43 / v3;
  double l53 = H[33];
  l53 -= l50 * l30 * D0;
  l53 -= l51 * l31 * D1;
  l53 -= l52 * l32 * D2;
  l53 = l53 / v3;
  double v4 = H[28];
  v4 -= l40 * l40 * D0;
  v4 -= l41 * l41 * D1;
  v4 -= l42 * l42 * D2;
  v4 -= l43 * l43 * D3;
  if (!(v4 > 1e-12)) return false;
  const double D4 = v4;
  double l54 = H[34];
  l54 -= l50 * l40 * D0;
  l54 -= l51 * l41 * D1;
  l54 -= l52 * l42 * D2;
  l54 -= l53 * l43 * D3;
  l54 = l54 / v4;
  double v5 = H[35];
  v5 -= l50 * l50 * D0;
  v5 -= l51 * l51 * D1;
  v5 -= l52 * l52 * D2;
  v5 -= l53 * l53 * D3;
  v5 -= l54 * l54 * D4;
  if (!(v5 > 1e-12)) return false;
  const double D5 = v5;
  double y0 = b[0];
  double y1 = b[1];
  y1 -= l10 * y0;
  double y2 = b[2];
  y2 -= l20 * y0;
  y2 -= l21 * y1;
  double y3 = b[3];
  y3 -= l30 * y0;
  y3 -= l31 * y1;
  y3 -= l32 * y2;
  double y4 = b[4];
  y4 -= l40 * y0;
  y4 -= l41 * y1;
  y4 -= l42 * y2;
  y4 -= l43 * y3;
  double y5 = b[5];
  y5 -= l50 * y0;
  y5 -= l51 * y1;
  y5 -= l52 * y2;
  y5 -= l53 * y3;
  y5 -= l54 * y4;
  y0 /= D0;
  y1 /= D1;
  y2 /= D2;
  y3 /= D3;
  y4 /= D4;
  y5 /= D5;
  double d5 = y5;
  double d4 = y4;
  d4 -= l54 * d5;
  double d3 = y3;
  d3 -= l43 * d4;
  d3 -= l53 * d5;
  double d2 = y2;
  d2 -= l32 * d3;
  d2 -= l42 * d4;
  d2 -= l52 * d5;
  double d1 = y1;
  d1 -= l21 * d2;
  d1 -= l31 * d3;
  d1 -= l41 * d4;
  d1 -= l51 * d5;
  double d0 = y0;
  d0 -= l10 * d1;
  d0 -= l20 * d2;
  d0 -= l30 * d3;
  d0 -= l40 * d4;
  d0 -= l50 * d5;
  d[0] = d0; d[1] = d1; d[2] = d2; d[3] = d3; d[4] = d4; d[5] = d5;
  return true;
}

// solve H d = b (6x6, LDL^T) and apply the left-multiplicative update; false if H is not positive definite
__device__ __forceinline__ bool pnp_step(const double* H, const double* b, PnpPose& T) {
  double d[6];
  if (!ldlt6_solve(H, b, d)) return false;
  double q[4] = {1.0, 0.5 * d[0], 0.5 * d[1], 0.5 * d[2]};
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= nq;
  double dR[9], Rn[9], tn[3];
  quat_to_R9(q, dR);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Rn[3 * r + c] = dR[3 * r] * T.R[c] + dR[3 * r + 1] * T.R[3 + c] + dR[3 * r + 2] * T.R[6 + c];
    tn[r] = dR[3 * r] * T.t[0] + dR[3 * r + 1] * T.t[1] + dR[3 * r + 2] * T.t[2] + d[3 + r];
  }
  for (int k = 0; k < 9; k++) T.R[k] = Rn[k];
  for (int k = 0; k < 3; k++) T.t[k] = tn[k];
  return true;
}

__device__ __forceinline__ bool pnp_inlier(const PnpPose& T, const float* X3, const float* uv, double fx, double fy, double cx,
                                           double cy, double thr) {
  double r[2];
  if (!pnp_row(T, X3, uv, fx, fy, cx, cy, r, nullptr, nullptr)) return false;
  return sqrt(r[0] * r[0] + r[1] * r[1]) < thr;
}

// ---- kernel 1: one warp per hypothesis
// PnpArgs are laid out in PNP_ARGS_SLOT-byte slots: problem k of a batched launch is blockIdx.y (hypotheses) / blockIdx.x (select)
static constexpr size_t PNP_ARGS_SLOT = 256;
__device__ __forceinline__ const PnpArgs& pnp_args(const PnpArgs* ap, int k) {
  return *(const PnpArgs*)((const char*)ap + (size_t)k * PNP_ARGS_SLOT);
}

__global__ void __launch_bounds__(PNP_THREADS) pnp_hypotheses_kernel(const PnpArgs* __restrict__ ap) {
  const PnpArgs a = pnp_args(ap, blockIdx.y);
  const int lane = threadIdx.x & 31;
  const int iter = blockIdx.x * (PNP_THREADS / 32) + (threadIdx.x >> 5);
  if (iter >= a.iters || a.M < 4) return;
  const double fx = a.fx, fy = a.fy, cx = a.cx, cy = a.cy;
  PnpPose T;
  int ok = 1;
  if (lane == 0) {
    int s[4];
    unsigned long long c = (unsigned long long)iter << 8;
    for (int j = 0; j < 4; j++) {
      while (true) {
        int v = (int)(splitmix(c++) % (unsigned long long)a.M);
        bool dup = false;
        for (int k = 0; k < j; k++) dup |= (s[k] == v);
        if (!dup) { s[j] = v; break; }
      }
    }
    const float* Tm = a.Tcw_motion;
    for (int r = 0; r < 3; r++) { for (int c2 = 0; c2 < 3; c2++) T.R[3 * r + c2] = Tm[4 * r + c2]; T.t[r] = Tm[4 * r + 3]; }
    for (int it = 0; it < 6 && ok; it++) {
      double H[36], b[6];
      for (int k = 0; k < 36; k++) H[k] = 0;
      for (int k = 0; k < 6; k++) b[k] = 0;
      for (int k = 0; k < 4 && ok; k++) {
        const int i = a.good[s[k]];
        double r[2], J0[6], J1[6];
        if (!pnp_row(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, r, J0, J1)) { ok = 0; break; }
        for (int p = 0; p < 6; p++) {
          b[p] -= J0[p] * r[0] + J1[p] * r[1];
          for (int q = 0; q < 6; q++) H[6 * p + q] += J0[p] * J0[q] + J1[p] * J1[q];
        }
      }
      if (ok && !pnp_step(H, b, T)) ok = 0;
    }
    a.hyp[iter] = T;
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) {
    if (lane == 0) a.hyp_cnt[iter] = -1;
    return;
  }
#pragma unroll
  for (int k = 0; k < 9; k++) T.R[k] = __shfl_sync(0xffffffffu, T.R[k], 0);
#pragma unroll
  for (int k = 0; k < 3; k++) T.t[k] = __shfl_sync(0xffffffffu, T.t[k], 0);
  int cnt = 0;
  for (int k = lane; k < a.M; k += 32) {
    const int i = a.good[k];
    cnt += pnp_inlier(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, (double)a.thr) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) a.hyp_cnt[iter] = cnt;
}

__device__ int ransac_update_iters(double p, double ep, int modelPoints, int maxIters) {  // cv::RANSACUpdateNumIters
  p = fmin(fmax(p, 0.), 1.);
  ep = fmin(fmax(ep, 0.), 1.);
  double num = fmax(1. - p, 2.220446049250313e-16);
  double denom = 1. - pow(1. - ep, (double)modelPoints);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= maxIters * (-denom)) ? maxIters : (int)llrint(num / denom);
}

// ---- kernel 2: replay the sequential selection, consensus set, refit, motion model, winner (one CTA)
__global__ void __launch_bounds__(PNP_THREADS) pnp_select_kernel(const PnpArgs* __restrict__ ap) {
  const PnpArgs a = pnp_args(ap, blockIdx.x);
  int* __restrict__ tmp_ids = a.tmp_ids;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double fx = a.fx, fy = a.fy, cx = a.cx, cy = a.cy;
  __shared__ int s_best, s_cnt, s_warp[PNP_THREADS / 32], s_base, s_ok, s_nr, s_nm;
  __shared__ PnpPose T;
  __shared__ double red[(PNP_THREADS / 32) * 27 + 32];
  if (tid == 0) {
    int best = -1, best_cnt = 0, niters = a.iters;
    if (a.M >= 4) {
      for (int iter = 0; iter < niters; iter++) {
        const int cnt = a.hyp_cnt[iter];
        if (cnt > max(best_cnt, 3)) {
          best_cnt = cnt;
          best = iter;
          niters = ransac_update_iters((double)a.confidence, (double)(a.M - cnt) / a.M, 4, niters);
        }
      }
    }
    s_best = best;
    if (best >= 0) T = a.hyp[best];
    s_nr = 0;
  }
  __syncthreads();
  int nr = 0;
  if (s_best >= 0) {
    // consensus set of the best hypothesis, ascending positions (ordered compaction into tmp_ids)
    s_base = 0;
    __syncthreads();
    for (int k0 = 0; k0 < a.M; k0 += PNP_THREADS) {
      const int k = k0 + tid;
      bool in = false;
      if (k < a.M) { const int i = a.good[k]; in = pnp_inlier(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, (double)a.thr); }
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      int off = s_base;
      for (int w = 0; w < warp; w++) off += s_warp[w];
      if (in) tmp_ids[off + __popc(m & ((1u << lane) - 1u))] = k;
      __syncthreads();
      if (tid == 0) { int t = 0; for (int w = 0; w < PNP_THREADS / 32; w++) t += s_warp[w]; s_base += t; }
      __syncthreads();
    }
    nr = s_base;
    // refit on the consensus set: 10 Gauss-Newton steps, block-reduced normal equations
    for (int it = 0; it < 10; it++) {
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; k++) acc[k] = 0;
      int bad = 0;
      for (int k = tid; k < nr; k += PNP_THREADS) {
        const int i = a.good[tmp_ids[k]];
        double r[2], J0[6], J1[6];
        if (!pnp_row(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, r, J0, J1)) { bad = 1; continue; }
        int idx = 0;
#pragma unroll
        for (int p = 0; p < 6; p++) {
          acc[21 + p] -= J0[p] * r[0] + J1[p] * r[1];
#pragma unroll
          for (int q = p; q < 6; q++) acc[idx++] += J0[p] * J0[q] + J1[p] * J1[q];
        }
      }
      bad = __syncthreads_or(bad);
#pragma unroll
      for (int k = 0; k < 27; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (lane == 0)
        for (int k = 0; k < 27; k++) red[warp * 27 + k] = acc[k];
      __syncthreads();
      if (tid == 0) {
        double H[36], b[6];
        int idx = 0;
        for (int p = 0; p < 6; p++) {
          double s = 0;
          for (int w = 0; w < PNP_THREADS / 32; w++) s += red[w * 27 + 21 + p];
          b[p] = s;
          for (int q = p; q < 6; q++) {
            double h = 0;
            for (int w = 0; w < PNP_THREADS / 32; w++) h += red[w * 27 + idx];
            H[6 * p + q] = h; H[6 * q + p] = h;
            idx++;
          }
        }
        PnpPose Tn = T;
        s_ok = (!bad && pnp_step(H, b, Tn)) ? 1 : 0;
        if (s_ok) T = Tn;
      }
      __syncthreads();
      if (!s_ok) break;  // the refit is abandoned as a whole when a step fails (oracle: keep the unrefined model)
    }
  }
  // NOTE: a failed refit must fall back to the un-refitted hypothesis
  if (tid == 0 && s_best >= 0 && !s_ok) T = a.hyp[s_best];
  __syncthreads();
  // ---- constant-velocity model in float32 (src/Tracking.cc:1980-2001), ordered compaction into inlier_ids
  const float* Tm = a.Tcw_motion;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < a.n; k0 += PNP_THREADS) {
    const int i = k0 + tid;
    bool in = false;
    if (i < a.n) {
      const float X[3] = {a.pts3d[3 * i], a.pts3d[3 * i + 1], a.pts3d[3 * i + 2]};
      float pc[3];
      for (int r = 0; r < 3; r++)
        pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tm[4 * r], X[0]), __fmul_rn(Tm[4 * r + 1], X[1])), __fmul_rn(Tm[4 * r + 2], X[2])), Tm[4 * r + 3]);
      const float invz = (float)(1.0 / (double)pc[2]);
      const float u = __fadd_rn(__fmul_rn(__fmul_rn(a.fx, pc[0]), invz), a.cx);
      const float v = __fadd_rn(__fmul_rn(__fmul_rn(a.fy, pc[1]), invz), a.cy);
      const float du = __fsub_rn(a.cur_xy[2 * i], u), dv = __fsub_rn(a.cur_xy[2 * i + 1], v);
      const float rpe = __fsqrt_rn(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
      in = rpe < a.thr;
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_warp[w];
    if (in) a.inlier_ids[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < PNP_THREADS / 32; w++) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  const int nm = s_base;
  // ---- winner (src/Tracking.cc:2006-2025)
  if (a.no_mm || nr > nm) {
    for (int k = tid; k < nr; k += PNP_THREADS) a.inlier_ids[k] = tmp_ids[k];  // position inside the valid list, as the reference does
    if (tid == 0) {
      if (s_best >= 0) {
        for (int r = 0; r < 3; r++) {
          for (int c = 0; c < 3; c++) a.Tcw_out[4 * r + c] = (float)T.R[3 * r + c];
          a.Tcw_out[4 * r + 3] = (float)T.t[r];
        }
        a.Tcw_out[12] = 0.f; a.Tcw_out[13] = 0.f; a.Tcw_out[14] = 0.f; a.Tcw_out[15] = 1.f;
      } else {
        for (int k = 0; k < 16; k++) a.Tcw_out[k] = Tm[k];   // no consensus at all: the seed pose, zero inliers
      }
      a.result[0] = nr; a.result[1] = 0;
    }
  } else if (tid == 0) {
    for (int k = 0; k < 16; k++) a.Tcw_out[k] = Tm[k];
    a.result[0] = nm; a.result[1] = 1;
  }
  if (tid == 0) { a.result[2] = nr; a.result[3] = nm; }
}

// =========================================================================================================
static constexpr int PNP_MAX_BATCH = 8;   // problems per launch (the objects of one frame)

struct PnpWorkspace {
  int capN = 0, capIters = 0;
  // one pinned staging block each way: [PnpArgs slots | per problem: Tcw_motion | cur_xy | pts3d | good] -> device,
  // [per problem: result | Tcw_out | ids] <- device
  char *d_in = nullptr, *h_in = nullptr, *d_out = nullptr, *h_out = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
  int *tmp = nullptr, *cnt = nullptr;   // [PNP_MAX_BATCH][capN], [PNP_MAX_BATCH][capIters]
  PnpPose* hyp = nullptr;               // [PNP_MAX_BATCH][capIters]
  PnpPose* c_hyp = nullptr;             // the device-chained tracker's own hypothesis scratch (it may run beside a host-driven batch)
  int* c_cnt = nullptr;
  // device-chained camera problem (pnp_chain_enqueue): arguments per state buffer, inputs built on the device, outputs
  char* c_block = nullptr;
  PnpArgs* c_args[2] = {nullptr, nullptr};
  float *c_p3d = nullptr, *c_tm = nullptr, *c_T = nullptr;
  int *c_good = nullptr, *c_ids = nullptr, *c_res = nullptr, *c_tmp = nullptr;
  int c_cap = 0;
  bool c_ready[2] = {false, false};
};

int pnp_setup(vido_ctx* ctx, int capN, int capIters) {
  static_assert(sizeof(PnpArgs) <= PNP_ARGS_SLOT, "PnpArgs slot");
  PnpWorkspace* ws = new PnpWorkspace();
  ctx->pnp = ws;
  ws->capN = capN; ws->capIters = capIters;
  // the batch shares one data region of capN points in total (a frame's objects partition its features)
  ws->in_bytes = PNP_ARGS_SLOT * PNP_MAX_BATCH + PNP_MAX_BATCH * (sizeof(float) * 16 + 64) + sizeof(float) * 5 * (size_t)capN + sizeof(int) * (size_t)capN;
  ws->out_bytes = PNP_MAX_BATCH * (sizeof(int) * 4 + sizeof(float) * 16 + 64) + sizeof(int) * (size_t)capN;
  VIDO_CUDA(cudaMalloc(&ws->d_in, ws->in_bytes));
  VIDO_CUDA(cudaMalloc(&ws->d_out, ws->out_bytes));
  VIDO_CUDA(cudaMallocHost(&ws->h_in, ws->in_bytes));
  VIDO_CUDA(cudaMallocHost(&ws->h_out, ws->out_bytes));
  VIDO_CUDA(cudaMalloc(&ws->tmp, sizeof(int) * (size_t)capN));
  VIDO_CUDA(cudaMalloc(&ws->cnt, sizeof(int) * (size_t)capIters * PNP_MAX_BATCH));
  VIDO_CUDA(cudaMalloc(&ws->hyp, sizeof(PnpPose) * (size_t)capIters * PNP_MAX_BATCH));
  return VIDO_OK;
}

void pnp_teardown(vido_ctx* ctx) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  if (!ws) return;
  cudaFree(ws->d_in); cudaFree(ws->d_out); cudaFreeHost(ws->h_in); cudaFreeHost(ws->h_out);
  cudaFree(ws->tmp); cudaFree(ws->cnt); cudaFree(ws->hyp); cudaFree(ws->c_block); cudaFree(ws->c_hyp); cudaFree(ws->c_cnt);
  delete ws;
  ctx->pnp = nullptr;
}

// up to PNP_MAX_BATCH problems in one pair of launches (one H2D block, one D2H block, one synchronisation)
static int pnp_batch(vido_ctx* ctx, vido_pnp_problem* ps, int nb) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  cudaStream_t s = ctx->stream;
  auto a16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t in_off = PNP_ARGS_SLOT * (size_t)nb, out_off = 0, tmp_off = 0;
  size_t o_res[PNP_MAX_BATCH], o_T[PNP_MAX_BATCH], o_ids[PNP_MAX_BATCH];
  int max_iters = 1, any_ransac = 0;
  for (int k = 0; k < nb; k++) {
    vido_pnp_problem* p = &ps[k];
    const size_t n = (size_t)p->n;
    const size_t o_tm = in_off, o_cur = o_tm + sizeof(float) * 16, o_pts = a16(o_cur + sizeof(float) * 2 * n), o_good = a16(o_pts + sizeof(float) * 3 * n);
    if (o_good + sizeof(int) * n > ws->in_bytes || tmp_off + n > (size_t)ws->capN) { ctx->err = "PnP batch exceeds capacity"; return VIDO_ERR_CAPACITY; }
    int* hgood = (int*)(ws->h_in + o_good);
    int M = 0;
    for (int i = 0; i < p->n; i++)
      if (!p->valid || p->valid[i]) hgood[M++] = i;
    in_off = a16(o_good + sizeof(int) * (size_t)M);
    o_res[k] = out_off; o_T[k] = out_off + sizeof(int) * 4; o_ids[k] = o_T[k] + sizeof(float) * 16;
    out_off = a16(o_ids[k] + sizeof(int) * n);
    if (out_off > ws->out_bytes) { ctx->err = "PnP batch exceeds capacity"; return VIDO_ERR_CAPACITY; }
    PnpArgs a;
    memset(&a, 0, sizeof a);
    a.n = p->n; a.M = M; a.iters = p->iters; a.no_mm = p->no_motion_model;
    a.Tcw_motion = (const float*)(ws->d_in + o_tm); a.cur_xy = (const float*)(ws->d_in + o_cur);
    a.pts3d = (const float*)(ws->d_in + o_pts); a.good = (const int*)(ws->d_in + o_good);
    a.fx = p->fx; a.fy = p->fy; a.cx = p->cx; a.cy = p->cy; a.thr = p->reproj_err; a.confidence = p->confidence;
    a.hyp = ws->hyp + (size_t)k * ws->capIters; a.hyp_cnt = ws->cnt + (size_t)k * ws->capIters;
    a.result = (int*)(ws->d_out + o_res[k]); a.Tcw_out = (float*)(ws->d_out + o_T[k]); a.inlier_ids = (int*)(ws->d_out + o_ids[k]);
    a.tmp_ids = ws->tmp + tmp_off;
    tmp_off += n;
    memcpy(ws->h_in + PNP_ARGS_SLOT * (size_t)k, &a, sizeof a);
    memcpy(ws->h_in + o_tm, p->Tcw_motion, sizeof(float) * 16);
    if (n) {
      memcpy(ws->h_in + o_cur, p->cur_xy, sizeof(float) * 2 * n);
      memcpy(ws->h_in + o_pts, p->pts3d, sizeof(float) * 3 * n);
    }
    max_iters = std::max(max_iters, p->iters);
    any_ransac |= (M >= 4);
  }
  VIDO_CUDA(cudaMemcpyAsync(ws->d_in, ws->h_in, in_off, cudaMemcpyHostToDevice, s));
  const PnpArgs* d_args = (const PnpArgs*)ws->d_in;
  cudaEventRecord(ctx->ev0, s);
  if (any_ransac) {
    pnp_hypotheses_kernel<<<dim3((max_iters + PNP_THREADS / 32 - 1) / (PNP_THREADS / 32), nb), PNP_THREADS, 0, s>>>(d_args);
    ctx->launches++;
  }
  pnp_select_kernel<<<nb, PNP_THREADS, 0, s>>>(d_args);
  cudaEventRecord(ctx->ev1, s);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  // the inlier lists are at most n ints each: fetch them with the results instead of paying a second round trip
  VIDO_CUDA(cudaMemcpyAsync(ws->h_out, ws->d_out, out_off, cudaMemcpyDeviceToHost, s));
  if (ctx->idle_work) { std::function<void()> f; f.swap(ctx->idle_work); f(); }   // host work hidden behind the kernels
  VIDO_CUDA(cudaStreamSynchronize(s));
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) { ctx->t_ms[1] += ms; ctx->t_n[1]++; }
  }
  for (int k = 0; k < nb; k++) {
    vido_pnp_problem* p = &ps[k];
    const int* res = (const int*)(ws->h_out + o_res[k]);
    p->n_inliers = res[0]; p->winner = res[1]; p->ransac_inliers = res[2]; p->mm_inliers = res[3];
    memcpy(p->Tcw_out, ws->h_out + o_T[k], sizeof(float) * 16);
    if (res[0] > 0 && p->inlier_ids) memcpy(p->inlier_ids, ws->h_out + o_ids[k], sizeof(int) * (size_t)res[0]);
  }
  return VIDO_OK;
}

int pnp_init_model_batch(vido_ctx* ctx, vido_pnp_problem* ps, int nproblems) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  for (int k = 0; k < nproblems; k++)
    if (ps[k].n < 0 || ps[k].n > ws->capN || ps[k].iters > ws->capIters || ps[k].iters < 1) { ctx->err = "PnP problem exceeds capacity"; return VIDO_ERR_CAPACITY; }
  // greedy packing: as many consecutive problems per launch as the shared point capacity allows
  int k0 = 0;
  while (k0 < nproblems) {
    int nb = 0;
    size_t pts = 0;
    while (k0 + nb < nproblems && nb < PNP_MAX_BATCH && (nb == 0 || pts + (size_t)ps[k0 + nb].n <= (size_t)ws->capN)) { pts += (size_t)ps[k0 + nb].n; nb++; }
    const int rc = pnp_batch(ctx, ps + k0, nb);
    if (rc) return rc;
    k0 += nb;
  }
  return VIDO_OK;
}

int pnp_init_model_host(vido_ctx* ctx, vido_pnp_problem* p) { return pnp_init_model_batch(ctx, p, 1); }

// =========================================================================================================
// device-chained camera init model: the problem is built from the device-resident tracker state (no host round trip)
// =========================================================================================================
// Tracking::GetInitModelCam prologue (src/Tracking.cc:1914-1965): 3-D points of the last frame through Twl (float, cv::Mat
// gemm rounding: double accumulation, one rounding, then the float translation add), valid = depth >= 0, the constant-
// velocity pose mVelocity * Tcw_last.  One CTA; writes n / M into the argument block of the two PnP kernels.
__global__ void __launch_bounds__(1024) pnp_chain_prep_kernel(PnpArgs* __restrict__ args, int32_t* __restrict__ hdr, const float* __restrict__ Tcw,
                                                              const float* __restrict__ vel, const float* __restrict__ keys,
                                                              const float* __restrict__ depth, float fx, float fy, float cx, float cy,
                                                              float* __restrict__ p3d, int* __restrict__ good, float* __restrict__ tm) {
  __shared__ float Twl[16];
  __shared__ int s_warp[32], s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = hdr[0];
  const bool skip = n < 2;   // "if (Ns < 2) skipped" of the per-frame driver: lost tracking, the frame is not processed
  if (tid == 0) {
    hdr[2] = skip ? 1 : 0;
    s_base = 0;
    for (int k = 0; k < 16; k++) Twl[k] = 0.f;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Twl[4 * r + c] = Tcw[4 * c + r];
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += (double)(-Twl[4 * r + k]) * (double)Tcw[4 * k + 3];
      Twl[4 * r + 3] = (float)s;
    }
    Twl[15] = 1.f;
  }
  if (tid < 16) {
    float v = Tcw[tid];
    if (hdr[1]) {   // mVelocity * Tcw_last, double accumulation, one rounding (cv::Mat CV_32F product)
      const int r = tid >> 2, c = tid & 3;
      double s = 0;
      for (int k = 0; k < 4; k++) s += (double)vel[4 * r + k] * (double)Tcw[4 * k + c];
      v = (float)s;
    }
    tm[tid] = v;
  }
  __syncthreads();
  const float invfx = __fdiv_rn(1.0f, fx), invfy = __fdiv_rn(1.0f, fy);
  const int nn = skip ? 0 : n;
  for (int i0 = 0; i0 < nn; i0 += 1024) {
    const int i = i0 + tid;
    bool ok = false;
    if (i < nn) {
      const float z = depth[i];
      ok = !(z < 0);
      float o[3] = {0.f, 0.f, 0.f};
      if (ok) {
        const float xc[3] = {__fmul_rn(__fmul_rn(__fsub_rn(keys[2 * i], cx), z), invfx), __fmul_rn(__fmul_rn(__fsub_rn(keys[2 * i + 1], cy), z), invfy), z};
        for (int r = 0; r < 3; r++)
          o[r] = __fadd_rn((float)((double)Twl[4 * r] * xc[0] + (double)Twl[4 * r + 1] * xc[1] + (double)Twl[4 * r + 2] * xc[2]), Twl[4 * r + 3]);
      }
      p3d[3 * i] = o[0]; p3d[3 * i + 1] = o[1]; p3d[3 * i + 2] = o[2];
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_warp[w];
    if (ok) good[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < 32; w++) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  if (tid == 0) { args->n = nn; args->M = s_base; }
}

int pnp_chain_setup(vido_ctx* ctx, int cap) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  if (cap > ws->capN) { ctx->err = "chain capacity exceeds the PnP capacity"; return VIDO_ERR_CAPACITY; }
  ws->c_cap = cap;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_args = 0, o_p3d = o_args + 2 * PNP_ARGS_SLOT, o_tm = o_p3d + al(12 * (size_t)cap), o_T = o_tm + 256, o_good = o_T + 256,
               o_ids = o_good + al(4 * (size_t)cap), o_res = o_ids + al(4 * (size_t)cap), o_tmp = o_res + 256, total = o_tmp + al(4 * (size_t)cap);
  VIDO_CUDA(cudaMalloc(&ws->c_block, total));
  VIDO_CUDA(cudaMemset(ws->c_block, 0, total));
  VIDO_CUDA(cudaMalloc(&ws->c_cnt, sizeof(int) * (size_t)ws->capIters));
  VIDO_CUDA(cudaMalloc(&ws->c_hyp, sizeof(PnpPose) * (size_t)ws->capIters));
  ws->c_args[0] = (PnpArgs*)(ws->c_block + o_args); ws->c_args[1] = (PnpArgs*)(ws->c_block + o_args + PNP_ARGS_SLOT);
  ws->c_p3d = (float*)(ws->c_block + o_p3d); ws->c_tm = (float*)(ws->c_block + o_tm); ws->c_T = (float*)(ws->c_block + o_T);
  ws->c_good = (int*)(ws->c_block + o_good); ws->c_ids = (int*)(ws->c_block + o_ids); ws->c_res = (int*)(ws->c_block + o_res);
  ws->c_tmp = (int*)(ws->c_block + o_tmp);
  ws->c_ready[0] = ws->c_ready[1] = false;
  return VIDO_OK;
}

// queue (no synchronisation): prologue, 500 hypotheses, selection.  `which` = index of the state buffer `st` (its pointers are
// baked into that buffer's argument block the first time it is used).
int pnp_chain_enqueue(vido_ctx* ctx, const ChainStateDev& st, int which, ChainPnpOut* out) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  cudaStream_t s = ctx->stream;
  const vido_config& c = ctx->cfg;
  vido_pnp_problem dp;
  vido_pnp_default_params(&dp);
  if (!ws->c_ready[which]) {
    PnpArgs a;
    memset(&a, 0, sizeof a);
    a.iters = dp.iters; a.no_mm = 0;
    a.cur_xy = st.corres; a.pts3d = ws->c_p3d; a.good = ws->c_good; a.Tcw_motion = ws->c_tm;
    a.fx = c.fx; a.fy = c.fy; a.cx = c.cx; a.cy = c.cy; a.thr = dp.reproj_err; a.confidence = dp.confidence;
    a.hyp = ws->c_hyp; a.hyp_cnt = ws->c_cnt;
    a.Tcw_out = ws->c_T; a.inlier_ids = ws->c_ids; a.result = ws->c_res; a.tmp_ids = ws->c_tmp;
    VIDO_CUDA(cudaMemcpyAsync(ws->c_args[which], &a, sizeof a, cudaMemcpyHostToDevice, s));   // pageable source: staged before return
    ws->c_ready[which] = true;
  }
  pnp_chain_prep_kernel<<<1, 1024, 0, s>>>(ws->c_args[which], st.hdr, st.Tcw, st.vel, st.keys, st.depth, c.fx, c.fy, c.cx, c.cy, ws->c_p3d,
                                           ws->c_good, ws->c_tm);
  pnp_hypotheses_kernel<<<dim3((dp.iters + PNP_THREADS / 32 - 1) / (PNP_THREADS / 32), 1), PNP_THREADS, 0, s>>>(ws->c_args[which]);
  pnp_select_kernel<<<1, PNP_THREADS, 0, s>>>(ws->c_args[which]);
  ctx->launches += 3;
  VIDO_CUDA(cudaGetLastError());
  out->T = ws->c_T; out->ids = ws->c_ids; out->res = ws->c_res;
  return VIDO_OK;
}
