// poseopt_kernels.cu -- per-frame joint optical-flow + pose optimisation, one CTA per problem, the four
// optimisation rounds and every LM iteration inside one launch.
//
// Replaces Optimizer::PoseOptimizationFlow2Cam (src/Optimizer.cc:2622-2824) and, with another initial transform,
// Optimizer::PoseOptimizationFlow2 (:3037-3253):
//   VertexSE3Expmap (T <- exp(dx) T)      g2o/types/types_six_dof_expmap.h:67-85, se3quat.h:228-262
//   EdgeSE3ProjectFlow2                   g2o/types/types_six_dof_expmap.h:436-476, types_six_dof_expmap.cpp:805-845
//   EdgeFlowPrior                         g2o/types/types_six_dof_expmap.h:414-432
//   Schur complement over the 2-dof flow vertices + dense 6x6 solve   g2o/core/block_solver.hpp:367-486,
//                                         g2o/solvers/linear_solver_dense.h:65-116
// Because both Jacobians w.r.t. the flow vertex are the identity, H_ll = (w + w_prior) I and the Schur complement
// collapses to a weighted sum of J^T J: one 27-value block reduction (warp shuffles) per solve.
#include <algorithm>
#include <cstring>

#include "ctx.h"
#include "lm_device.h"

#define PO_THREADS 256

struct PoArgs {
  int n;
  const float* obs_xy;   // [n][2]
  const float* flow_xy;  // [n][2]
  const float* depth;    // [n]
  const float* Tcw_init; // [16]
  const float* Tcw_last; // [16]
  float fx, fy, cx, cy;
  double info_f, info_p, delta;
  float th0, th1;        // chi2 gates: round 0, rounds 1..3
  int rounds, its;
  // workspace
  double* Xw;      // [3n]
  double* flow;    // [2][2n]
  double* xl;      // [2n] last solved flow increment
  double* eProj;   // [2n]
  int* level;      // [n]
  // outputs
  float* Tcw_out;  // [16]
  float* flow_out; // [n][2]
  int* inlier;     // [n]
  int* n_inliers;  // [1]
  LmCtl* ctl;      // [rounds] final controller state per round
  LmRec* rec;      // [rounds][VIDO_LM_REC]
};

struct PoseQ {
  double q[4];  // w x y z
  double t[3];
};

__device__ __forceinline__ void q_rot(const double* q, const double* v, double* o) {
  const double ux = q[1], uy = q[2], uz = q[3], w = q[0];
  double cx = 2 * (uy * v[2] - uz * v[1]), cy = 2 * (uz * v[0] - ux * v[2]), cz = 2 * (ux * v[1] - uy * v[0]);
  o[0] = v[0] + w * cx + (uy * cz - uz * cy);
  o[1] = v[1] + w * cy + (uz * cx - ux * cz);
  o[2] = v[2] + w * cz + (ux * cy - uy * cx);
}

__device__ void quat_from_R_dev(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > (i == 0 ? R[0] : R[4])) i = 2;
    if (i == 0) {
      t = sqrt(R[0] - R[4] - R[8] + 1.0);
      q[1] = 0.5 * t; t = 0.5 / t;
      q[0] = (R[7] - R[5]) * t; q[2] = (R[3] + R[1]) * t; q[3] = (R[6] + R[2]) * t;
    } else if (i == 1) {
      t = sqrt(R[4] - R[8] - R[0] + 1.0);
      q[2] = 0.5 * t; t = 0.5 / t;
      q[0] = (R[2] - R[6]) * t; q[3] = (R[7] + R[5]) * t; q[1] = (R[1] + R[3]) * t;
    } else {
      t = sqrt(R[8] - R[0] - R[4] + 1.0);
      q[3] = 0.5 * t; t = 0.5 / t;
      q[0] = (R[3] - R[1]) * t; q[1] = (R[2] + R[6]) * t; q[2] = (R[5] + R[7]) * t;
    }
  }
  if (q[0] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// T <- exp(u) * T, u = [omega, upsilon] (SE3Quat::exp, se3quat.h:228-262)
__device__ void se3_exp_mul(const double* u, const PoseQ& T, PoseQ& o) {
  const double wx = u[0], wy = u[1], wz = u[2];
  const double theta = sqrt(wx * wx + wy * wy + wz * wz);
  const double Om[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double Om2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Om2[3 * i + j] = Om[3 * i] * Om[j] + Om[3 * i + 1] * Om[3 + j] + Om[3 * i + 2] * Om[6 + j];
  double R[9], V[9];
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + Om[i] + Om2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * Om[i] + b * Om2[i];
      V[i] = I + b * Om[i] + c * Om2[i];
    }
  }
  double dq[4], dt[3];
  quat_from_R_dev(R, dq);
  for (int i = 0; i < 3; i++) dt[i] = V[3 * i] * u[3] + V[3 * i + 1] * u[4] + V[3 * i + 2] * u[5];
  double rt[3];
  q_rot(dq, T.t, rt);
  o.t[0] = dt[0] + rt[0]; o.t[1] = dt[1] + rt[1]; o.t[2] = dt[2] + rt[2];
  double q[4];
  q[0] = dq[0] * T.q[0] - dq[1] * T.q[1] - dq[2] * T.q[2] - dq[3] * T.q[3];
  q[1] = dq[0] * T.q[1] + dq[1] * T.q[0] + dq[2] * T.q[3] - dq[3] * T.q[2];
  q[2] = dq[0] * T.q[2] + dq[2] * T.q[0] + dq[3] * T.q[1] - dq[1] * T.q[3];
  q[3] = dq[0] * T.q[3] + dq[3] * T.q[0] + dq[1] * T.q[2] - dq[2] * T.q[1];
  if (q[0] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) o.q[i] = q[i] / n;
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__device__ __forceinline__ void bsum(double* v, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = wsum(v[k]);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < NV; k++) sm[warp * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < nw; w++) s += sm[w * NV + threadIdx.x];
    sm[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ void huber_dev(double e, double delta, double& rho0, double& w) {
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho0 = e; w = 1.0; }
  else { const double s = sqrt(e); rho0 = 2 * s * delta - dsqr; w = delta / s; }
}

// projection error and (optionally) the 2x6 Jacobian w.r.t. the pose at transform T
__device__ __forceinline__ void proj_edge(const PoArgs& a, const PoseQ& T, int i, const double* fl, double* e, double* J) {
  double pc[3];
  q_rot(T.q, a.Xw + 3 * (size_t)i, pc);
  pc[0] += T.t[0]; pc[1] += T.t[1]; pc[2] += T.t[2];
  const double fx = a.fx, fy = a.fy;
  e[0] = ((double)a.obs_xy[2 * i] + fl[0]) - (pc[0] / pc[2] * fx + (double)a.cx);
  e[1] = ((double)a.obs_xy[2 * i + 1] + fl[1]) - (pc[1] / pc[2] * fy + (double)a.cy);
  if (J) {
    const double x = pc[0], y = pc[1], z = pc[2], z2 = z * z;
    J[0] = x * y / z2 * fx; J[1] = -(1 + (x * x / z2)) * fx; J[2] = y / z * fx; J[3] = -1. / z * fx; J[4] = 0; J[5] = x / z2 * fx;
    J[6] = (1 + y * y / z2) * fy; J[7] = -x * y / z2 * fy; J[8] = -x / z * fy; J[9] = 0; J[10] = -1. / z * fy; J[11] = y / z2 * fy;
  }
}

__global__ void __launch_bounds__(PO_THREADS) poseopt_flow2_kernel(const PoArgs* __restrict__ problems) {
  const PoArgs a = problems[blockIdx.x];
  const int n = a.n, tid = threadIdx.x;
  __shared__ double red[(PO_THREADS / 32) * 27 + 32];
  __shared__ PoseQ T[2];
  __shared__ PoseQ Tinit;
  __shared__ LmCtl ctl;
  __shared__ double Hpp[36], bp[6], xp[6];
  __shared__ int s_robust, s_nbad;

  if (n < 3) {  // "if(nInitialCorrespondences<3) return 0;" -- nothing is touched
    for (int i = tid; i < 16; i += PO_THREADS) a.Tcw_out[i] = a.Tcw_init[i];
    for (int i = tid; i < n; i += PO_THREADS) {
      a.flow_out[2 * i] = a.flow_xy[2 * i]; a.flow_out[2 * i + 1] = a.flow_xy[2 * i + 1];
      a.inlier[i] = 1;
    }
    if (tid == 0) *a.n_inliers = 0;
    return;
  }
  // ---- setup: Twl in float like the reference's cv::Mat code, world points in double
  if (tid == 0) {
    const float* L = a.Tcw_init;
    double R[9] = {L[0], L[1], L[2], L[4], L[5], L[6], L[8], L[9], L[10]};
    quat_from_R_dev(R, Tinit.q);
    Tinit.t[0] = L[3]; Tinit.t[1] = L[7]; Tinit.t[2] = L[11];
    s_robust = 1;
    for (int k = 0; k < 6; k++) xp[k] = 0;
  }
  {
    const float* L = a.Tcw_last;
    float Rwl[9], twl[3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Rwl[3 * r + c] = L[4 * c + r];
    for (int r = 0; r < 3; r++) {
      float s = 0.f;
      for (int k = 0; k < 3; k++) s = __fadd_rn(s, __fmul_rn(-Rwl[3 * r + k], L[4 * k + 3]));
      twl[r] = s;
    }
    for (int i = tid; i < n; i += PO_THREADS) {
      const double ox = a.obs_xy[2 * i], oy = a.obs_xy[2 * i + 1], d = a.depth[i];
      const double X[3] = {(ox - (double)a.cx) * d / (double)a.fx, (oy - (double)a.cy) * d / (double)a.fy, d};
      for (int r = 0; r < 3; r++)
        a.Xw[3 * (size_t)i + r] = (double)Rwl[3 * r] * X[0] + (double)Rwl[3 * r + 1] * X[1] + (double)Rwl[3 * r + 2] * X[2] + (double)twl[r];
      a.flow[2 * i] = a.flow_xy[2 * i]; a.flow[2 * i + 1] = a.flow_xy[2 * i + 1];
      a.xl[2 * i] = 0; a.xl[2 * i + 1] = 0;
      a.level[i] = 0;
    }
  }
  __syncthreads();

  double* const flowbuf[2] = {a.flow, a.flow + 2 * (size_t)n};
  for (int round = 0; round < a.rounds; round++) {
    if (tid == 0) {
      lm_reset(&ctl);
      T[0] = Tinit;
    }
    __syncthreads();
    const int robust = s_robust;
    // ---- robust chi2 at the start state (buffer 0)
    {
      double chi = 0;
      for (int i = tid; i < n; i += PO_THREADS) {
        const double* fl = flowbuf[0] + 2 * i;
        if (a.level[i] == 0) {
          double e[2], r0, w;
          proj_edge(a, T[0], i, fl, e, nullptr);
          a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
          const double c = (e[0] * e[0] + e[1] * e[1]) * a.info_f;
          if (robust) { huber_dev(c, a.delta, r0, w); chi += r0; } else chi += c;
        }
        const double p0 = fl[0] - (double)a.flow_xy[2 * i], p1 = fl[1] - (double)a.flow_xy[2 * i + 1];
        chi += (p0 * p0 + p1 * p1) * a.info_p;
      }
      double v[1] = {chi};
      bsum<1>(v, red);
      if (tid == 0) ctl.currentChi = red[0];
      __syncthreads();
    }
    for (int it = 0; it < a.its; it++) {
      if (ctl.stop_flag || !ctl.ok) break;
      const int cur = ctl.cur;
      const PoseQ Tc = T[cur];
      const double* fcur = flowbuf[cur];
      double* ftrial = flowbuf[cur ^ 1];
      // ---- build: H_pp, b_p and the diagonal maximum
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; k++) acc[k] = 0;
      double mx = 0;
      for (int i = tid; i < n; i += PO_THREADS) {
        double h = a.info_p;
        if (a.level[i] == 0) {
          double e[2], J[12], r0, w = 1.0;
          proj_edge(a, Tc, i, fcur + 2 * i, e, J);
          a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
          if (robust) huber_dev((e[0] * e[0] + e[1] * e[1]) * a.info_f, a.delta, r0, w);
          w *= a.info_f;
          h += w;
          int idx = 0;
#pragma unroll
          for (int r = 0; r < 6; r++) {
            acc[21 + r] -= w * (J[r] * e[0] + J[6 + r] * e[1]);
#pragma unroll
            for (int c = r; c < 6; c++) acc[idx++] += w * (J[r] * J[c] + J[6 + r] * J[6 + c]);
          }
        }
        mx = fmax(mx, h);
      }
      bsum<27>(acc, red);
      if (tid == 0) {
        int idx = 0;
        for (int r = 0; r < 6; r++) {
          bp[r] = red[21 + r];
          for (int c = r; c < 6; c++) { Hpp[6 * r + c] = red[idx]; Hpp[6 * c + r] = red[idx]; idx++; }
        }
      }
      __syncthreads();
      {  // block max
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        if (tid == 0) {
          double m = 0;
          for (int w = 0; w < PO_THREADS / 32; w++) m = fmax(m, red[w]);
          for (int r = 0; r < 6; r++) m = fmax(m, fabs(Hpp[7 * r]));
          lm_begin_iteration(&ctl, it, m, -1.0);
        }
        __syncthreads();
      }
      // ---- trials
      while (true) {
        const double lambda = ctl.lambda;
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0;
        for (int i = tid; i < n; i += PO_THREADS) {
          if (a.level[i] != 0) continue;
          double e[2], J[12], r0, w = 1.0;
          proj_edge(a, Tc, i, fcur + 2 * i, e, J);
          if (robust) huber_dev((e[0] * e[0] + e[1] * e[1]) * a.info_f, a.delta, r0, w);
          w *= a.info_f;
          const double h = w + a.info_p + lambda;
          const double p0 = fcur[2 * i] - (double)a.flow_xy[2 * i], p1 = fcur[2 * i + 1] - (double)a.flow_xy[2 * i + 1];
          const double bl0 = -w * e[0] - a.info_p * p0, bl1 = -w * e[1] - a.info_p * p1;
          const double s = w * w / h, g = w / h;
          int idx = 0;
#pragma unroll
          for (int r = 0; r < 6; r++) {
            acc[21 + r] += g * (J[r] * bl0 + J[6 + r] * bl1);
#pragma unroll
            for (int c = r; c < 6; c++) acc[idx++] += s * (J[r] * J[c] + J[6 + r] * J[6 + c]);
          }
        }
        bsum<27>(acc, red);
        if (tid == 0) {  // reduced 6x6 system, LDL^T with positivity test
          double S[36], bs[6];
          int idx = 0;
          for (int r = 0; r < 6; r++) {
            bs[r] = bp[r] - red[21 + r];
            for (int c = r; c < 6; c++) {
              const double v = Hpp[6 * r + c] - red[idx++] + ((r == c) ? lambda : 0.0);
              S[6 * r + c] = v; S[6 * c + r] = v;
            }
          }
          double Lm[36], D[6];
          for (int k = 0; k < 36; k++) Lm[k] = 0;
          bool ok = true;
          for (int j = 0; j < 6 && ok; j++) {
            double d = S[7 * j];
            for (int k = 0; k < j; k++) d -= Lm[6 * j + k] * Lm[6 * j + k] * D[k];
            if (!(d > 0)) { ok = false; break; }
            D[j] = d;
            Lm[7 * j] = 1;
            for (int i2 = j + 1; i2 < 6; i2++) {
              double s2 = S[6 * i2 + j];
              for (int k = 0; k < j; k++) s2 -= Lm[6 * i2 + k] * Lm[6 * j + k] * D[k];
              Lm[6 * i2 + j] = s2 / d;
            }
          }
          if (ok) {
            double y[6];
            for (int i2 = 0; i2 < 6; i2++) {
              double s2 = bs[i2];
              for (int k = 0; k < i2; k++) s2 -= Lm[6 * i2 + k] * y[k];
              y[i2] = s2;
            }
            for (int i2 = 0; i2 < 6; i2++) y[i2] /= D[i2];
            for (int i2 = 5; i2 >= 0; i2--) {
              double s2 = y[i2];
              for (int k = i2 + 1; k < 6; k++) s2 -= Lm[6 * k + i2] * xp[k];
              xp[i2] = s2;
            }
          }
          ctl.fail = ok ? 0 : 1;  // on failure x keeps its previous content
          se3_exp_mul(xp, Tc, T[cur ^ 1]);
        }
        __syncthreads();
        const int failed = ctl.fail;
        const PoseQ Tt = T[cur ^ 1];
        // ---- flow increments, trial state, chi2 and scale
        double chi = 0, scale = 0;
        for (int i = tid; i < n; i += PO_THREADS) {
          const double p0 = fcur[2 * i] - (double)a.flow_xy[2 * i], p1 = fcur[2 * i + 1] - (double)a.flow_xy[2 * i + 1];
          double bl0 = -a.info_p * p0, bl1 = -a.info_p * p1, h = a.info_p + lambda;
          double x0, x1;
          if (a.level[i] == 0) {
            double e[2], J[12], r0, w = 1.0;
            proj_edge(a, Tc, i, fcur + 2 * i, e, J);
            if (robust) huber_dev((e[0] * e[0] + e[1] * e[1]) * a.info_f, a.delta, r0, w);
            w *= a.info_f;
            h += w;
            bl0 -= w * e[0]; bl1 -= w * e[1];
            double jx0 = 0, jx1 = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) { jx0 += J[r] * xp[r]; jx1 += J[6 + r] * xp[r]; }
            x0 = (bl0 - w * jx0) / h;
            x1 = (bl1 - w * jx1) / h;
          } else {
            x0 = bl0 / h;
            x1 = bl1 / h;
          }
          if (failed) { x0 = a.xl[2 * i]; x1 = a.xl[2 * i + 1]; }
          else { a.xl[2 * i] = x0; a.xl[2 * i + 1] = x1; }
          const double f0 = fcur[2 * i] + x0, f1 = fcur[2 * i + 1] + x1;
          ftrial[2 * i] = f0; ftrial[2 * i + 1] = f1;
          scale += x0 * (lambda * x0 + bl0) + x1 * (lambda * x1 + bl1);
          if (a.level[i] == 0) {
            double e[2], r0, w;
            const double ft[2] = {f0, f1};
            proj_edge(a, Tt, i, ft, e, nullptr);
            a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
            const double c = (e[0] * e[0] + e[1] * e[1]) * a.info_f;
            if (robust) { huber_dev(c, a.delta, r0, w); chi += r0; } else chi += c;
          }
          const double q0 = f0 - (double)a.flow_xy[2 * i], q1 = f1 - (double)a.flow_xy[2 * i + 1];
          chi += (q0 * q0 + q1 * q1) * a.info_p;
        }
        double v[2] = {chi, scale};
        bsum<2>(v, red);
        if (tid == 0) {
          double sc = red[1];
          for (int r = 0; r < 6; r++) sc += xp[r] * (lambda * xp[r] + bp[r]);
          lm_trial(&ctl, red[0], sc, failed);
        }
        __syncthreads();
        if (!lm_more_trials(&ctl)) break;
      }
      if (tid == 0) lm_end_iteration(&ctl, it, -1.0, a.rec ? a.rec + (size_t)round * VIDO_LM_REC : nullptr);
      __syncthreads();
    }
    // ---- classify (src/Optimizer.cc:2751-2794); accepted state -> buffer 0 for the next round
    const int cur = ctl.cur;
    const float th = (round == 0) ? a.th0 : a.th1;
    int bad = 0;
    for (int i = tid; i < n; i += PO_THREADS) {
      const double f0 = flowbuf[cur][2 * i], f1 = flowbuf[cur][2 * i + 1];
      if (a.level[i] != 0) {  // demoted edges are re-evaluated at the final estimate
        double e[2];
        const double ft[2] = {f0, f1};
        proj_edge(a, T[cur], i, ft, e, nullptr);
        a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
      }
      const float chi2 = (float)((a.eProj[2 * i] * a.eProj[2 * i] + a.eProj[2 * i + 1] * a.eProj[2 * i + 1]) * a.info_f);
      if (chi2 > th) { a.level[i] = 1; bad++; }
      else a.level[i] = 0;
      flowbuf[0][2 * i] = f0; flowbuf[0][2 * i + 1] = f1;
    }
    {
      double v[1] = {(double)bad};
      bsum<1>(v, red);
      if (tid == 0) {
        s_nbad = (int)(red[0] + 0.5);
        if (round == 2) s_robust = 0;
        if (a.ctl) a.ctl[round] = ctl;
        T[0] = T[cur];  // keep the final pose of this round (overwritten by the reset unless it is the last round)
      }
      __syncthreads();
    }
  }
  // ---- outputs: pose (SE3Quat -> homogeneous -> float), refined flows, inlier flags
  if (tid == 0) {
    const PoseQ& F = T[0];
    const double w = F.q[0], x = F.q[1], y = F.q[2], z = F.q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x,
                 txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) a.Tcw_out[4 * r + c] = (float)R[3 * r + c];
      a.Tcw_out[4 * r + 3] = (float)F.t[r];
    }
    a.Tcw_out[12] = 0.f; a.Tcw_out[13] = 0.f; a.Tcw_out[14] = 0.f; a.Tcw_out[15] = 1.f;
    *a.n_inliers = n - s_nbad;
  }
  for (int i = tid; i < n; i += PO_THREADS) {
    a.flow_out[2 * i] = (float)flowbuf[0][2 * i];
    a.flow_out[2 * i + 1] = (float)flowbuf[0][2 * i + 1];
    a.inlier[i] = a.level[i] == 0;
  }
}


// =========================================================================================================
// Fast path (n <= PO_CACHE_CAP): 512 threads; the linearisation of every edge (error, robust weight, the ten non-zero
// Jacobian entries) is computed ONCE per LM iteration and kept in shared memory (struct-of-arrays, conflict-free), so the
// damped Schur pass of every trial and the flow back-substitution read 13 doubles instead of re-evaluating the
// projection (two divisions, a quaternion rotation) -- 2 projections per iteration instead of 5.  The 27-value block
// reductions use halving exchanges (31 shuffles instead of 135), the 6x6 solve is straight-line register code.
// Same arithmetic as the generic kernel above up to summation order.
// =========================================================================================================
#define PO_FAST_THREADS 512
#define PO_CACHE_CAP 1536
#define PO_NC 13  // e0 e1 w J0 J1 J2 J3 J5 J6 J7 J8 J10 J11

__device__ __forceinline__ double rsqrt_pos_po(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(d, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// 6x6 SPD solve S x = b by LL^T in straight-line scalar code (registers only); false if a pivot is not positive
__device__ __forceinline__ bool solve6_spd(const double* S, const double* b, double* x) {
  const double a00 = S[0];
  const double a10 = S[6], a11 = S[7];
  const double a20 = S[12], a21 = S[13], a22 = S[14];
  const double a30 = S[18], a31 = S[19], a32 = S[20], a33 = S[21];
  const double a40 = S[24], a41 = S[25], a42 = S[26], a43 = S[27], a44 = S[28];
  const double a50 = S[30], a51 = S[31], a52 = S[32], a53 = S[33], a54 = S[34], a55 = S[35];
  const double d0 = a00;
  const double i0 = rsqrt_pos_po(d0);
  const double l10 = (a10) * i0;
  const double l20 = (a20) * i0;
  const double l30 = (a30) * i0;
  const double l40 = (a40) * i0;
  const double l50 = (a50) * i0;
  const double d1 = a11 - l10 * l10;
  const double i1 = rsqrt_pos_po(d1);
  const double l21 = (a21 - l20 * l10) * i1;
  const double l31 = (a31 - l30 * l10) * i1;
  const double l41 = (a41 - l40 * l10) * i1;
  const double l51 = (a51 - l50 * l10) * i1;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  const double i2 = rsqrt_pos_po(d2);
  const double l32 = (a32 - l30 * l20 - l31 * l21) * i2;
  const double l42 = (a42 - l40 * l20 - l41 * l21) * i2;
  const double l52 = (a52 - l50 * l20 - l51 * l21) * i2;
  const double d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
  const double i3 = rsqrt_pos_po(d3);
  const double l43 = (a43 - l40 * l30 - l41 * l31 - l42 * l32) * i3;
  const double l53 = (a53 - l50 * l30 - l51 * l31 - l52 * l32) * i3;
  const double d4 = a44 - l40 * l40 - l41 * l41 - l42 * l42 - l43 * l43;
  const double i4 = rsqrt_pos_po(d4);
  const double l54 = (a54 - l50 * l40 - l51 * l41 - l52 * l42 - l53 * l43) * i4;
  const double d5 = a55 - l50 * l50 - l51 * l51 - l52 * l52 - l53 * l53 - l54 * l54;
  const double i5 = rsqrt_pos_po(d5);
  if (!(d0 > 0 && d1 > 0 && d2 > 0 && d3 > 0 && d4 > 0 && d5 > 0)) return false;
  const double y0 = (b[0]) * i0;
  const double y1 = (b[1] - l10 * y0) * i1;
  const double y2 = (b[2] - l20 * y0 - l21 * y1) * i2;
  const double y3 = (b[3] - l30 * y0 - l31 * y1 - l32 * y2) * i3;
  const double y4 = (b[4] - l40 * y0 - l41 * y1 - l42 * y2 - l43 * y3) * i4;
  const double y5 = (b[5] - l50 * y0 - l51 * y1 - l52 * y2 - l53 * y3 - l54 * y4) * i5;
  const double x5 = (y5) * i5;
  const double x4 = (y4 - l54 * x5) * i4;
  const double x3 = (y3 - l53 * x5 - l43 * x4) * i3;
  const double x2 = (y2 - l52 * x5 - l42 * x4 - l32 * x3) * i2;
  const double x1 = (y1 - l51 * x5 - l41 * x4 - l31 * x3 - l21 * x2) * i1;
  const double x0 = (y0 - l50 * x5 - l40 * x4 - l30 * x3 - l20 * x2 - l10 * x1) * i0;
  x[0] = x0; x[1] = x1; x[2] = x2; x[3] = x3; x[4] = x4; x[5] = x5;
  return true;
}

// sum of 32 per-lane values over the warp with 31 shuffles; afterwards v[0] of lane L is the total of value L
__device__ __forceinline__ void warp_sum32(double* v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 16, m = 16; h >= 1; h >>= 1, m >>= 1) {
    const bool hi = (lane & m) != 0;
#pragma unroll
    for (int k = 0; k < h; k++) {
      const double send = hi ? v[k] : v[k + h], keep = hi ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
}

// block sum of 27 values per thread -> sm[0..26] (v is clobbered); sm must hold 32 * (#warps) doubles
__device__ __forceinline__ void bsum27(double* v, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  warp_sum32(v);
  __syncthreads();
  sm[warp * 32 + lane] = v[0];
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0;
    for (int w = 0; w < nw; w++) s += sm[w * 32 + threadIdx.x];
    v[0] = s;
  }
  __syncthreads();
  if (threadIdx.x < 32) sm[threadIdx.x] = v[0];
  __syncthreads();
}

__global__ void __launch_bounds__(PO_FAST_THREADS) poseopt_flow2_fast_kernel(const PoArgs* __restrict__ problems) {
  const PoArgs a = problems[blockIdx.x];
  const int n = a.n, tid = threadIdx.x, NT = PO_FAST_THREADS;
  extern __shared__ double cache[];  // [PO_NC][n]
  __shared__ double red[(PO_FAST_THREADS / 32) * 32];
  __shared__ PoseQ T[2];
  __shared__ PoseQ Tinit;
  __shared__ LmCtl ctl;
  __shared__ double Hpp[36], bp[6], xp[6], Ss[36], bs[6];
  __shared__ int s_robust, s_nbad;

  if (n < 3) {  // "if(nInitialCorrespondences<3) return 0;" -- nothing is touched
    for (int i = tid; i < 16; i += NT) a.Tcw_out[i] = a.Tcw_init[i];
    for (int i = tid; i < n; i += NT) {
      a.flow_out[2 * i] = a.flow_xy[2 * i]; a.flow_out[2 * i + 1] = a.flow_xy[2 * i + 1];
      a.inlier[i] = 1;
    }
    if (tid == 0) *a.n_inliers = 0;
    return;
  }
  if (tid == 0) {
    const float* L = a.Tcw_init;
    double R[9] = {L[0], L[1], L[2], L[4], L[5], L[6], L[8], L[9], L[10]};
    quat_from_R_dev(R, Tinit.q);
    Tinit.t[0] = L[3]; Tinit.t[1] = L[7]; Tinit.t[2] = L[11];
    s_robust = 1;
    for (int k = 0; k < 6; k++) xp[k] = 0;
  }
  {
    const float* L = a.Tcw_last;
    float Rwl[9], twl[3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Rwl[3 * r + c] = L[4 * c + r];
    for (int r = 0; r < 3; r++) {
      float s = 0.f;
      for (int k = 0; k < 3; k++) s = __fadd_rn(s, __fmul_rn(-Rwl[3 * r + k], L[4 * k + 3]));
      twl[r] = s;
    }
    for (int i = tid; i < n; i += NT) {
      const double ox = a.obs_xy[2 * i], oy = a.obs_xy[2 * i + 1], d = a.depth[i];
      const double X[3] = {(ox - (double)a.cx) * d / (double)a.fx, (oy - (double)a.cy) * d / (double)a.fy, d};
      for (int r = 0; r < 3; r++)
        a.Xw[3 * (size_t)i + r] = (double)Rwl[3 * r] * X[0] + (double)Rwl[3 * r + 1] * X[1] + (double)Rwl[3 * r + 2] * X[2] + (double)twl[r];
      a.flow[2 * i] = a.flow_xy[2 * i]; a.flow[2 * i + 1] = a.flow_xy[2 * i + 1];
      a.xl[2 * i] = 0; a.xl[2 * i + 1] = 0;
      a.level[i] = 0;
    }
  }
  __syncthreads();

  double* const flowbuf[2] = {a.flow, a.flow + 2 * (size_t)n};
  double* const c_e0 = cache, *const c_e1 = cache + n, *const c_w = cache + 2 * (size_t)n, *const c_J = cache + 3 * (size_t)n;  // c_J[k*n + i], k < 10
  for (int round = 0; round < a.rounds; round++) {
    if (tid == 0) {
      lm_reset(&ctl);
      T[0] = Tinit;
    }
    __syncthreads();
    const int robust = s_robust;
    {  // robust chi2 at the start state (buffer 0)
      double v[32];
#pragma unroll
      for (int k = 0; k < 32; k++) v[k] = 0;
      const PoseQ T0 = T[0];
      for (int i = tid; i < n; i += NT) {
        const double* fl = flowbuf[0] + 2 * i;
        if (a.level[i] == 0) {
          double e[2], r0, w;
          proj_edge(a, T0, i, fl, e, nullptr);
          a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
          const double c = (e[0] * e[0] + e[1] * e[1]) * a.info_f;
          if (robust) { huber_dev(c, a.delta, r0, w); v[0] += r0; } else v[0] += c;
        }
        const double p0 = fl[0] - (double)a.flow_xy[2 * i], p1 = fl[1] - (double)a.flow_xy[2 * i + 1];
        v[0] += (p0 * p0 + p1 * p1) * a.info_p;
      }
      bsum27(v, red);
      if (tid == 0) ctl.currentChi = red[0];
      __syncthreads();
    }
    for (int it = 0; it < a.its; it++) {
      if (ctl.stop_flag || !ctl.ok) break;
      const int cur = ctl.cur;
      const PoseQ Tc = T[cur];
      const double* fcur = flowbuf[cur];
      double* ftrial = flowbuf[cur ^ 1];
      // ---- build: linearise every active edge once (cached), H_pp, b_p and the diagonal maximum
      double acc[32];
#pragma unroll
      for (int k = 0; k < 32; k++) acc[k] = 0;
      double mx = 0;
      for (int i = tid; i < n; i += NT) {
        double h = a.info_p;
        if (a.level[i] == 0) {
          double e[2], J[12], r0, w = 1.0;
          proj_edge(a, Tc, i, fcur + 2 * i, e, J);
          a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
          if (robust) huber_dev((e[0] * e[0] + e[1] * e[1]) * a.info_f, a.delta, r0, w);
          w *= a.info_f;
          h += w;
          c_e0[i] = e[0]; c_e1[i] = e[1]; c_w[i] = w;
          c_J[i] = J[0]; c_J[n + i] = J[1]; c_J[2 * n + i] = J[2]; c_J[3 * n + i] = J[3]; c_J[4 * n + i] = J[5];
          c_J[5 * n + i] = J[6]; c_J[6 * n + i] = J[7]; c_J[7 * n + i] = J[8]; c_J[8 * n + i] = J[10]; c_J[9 * n + i] = J[11];
          int idx = 0;
#pragma unroll
          for (int r = 0; r < 6; r++) {
            acc[21 + r] -= w * (J[r] * e[0] + J[6 + r] * e[1]);
#pragma unroll
            for (int c = r; c < 6; c++) acc[idx++] += w * (J[r] * J[c] + J[6 + r] * J[6 + c]);
          }
        }
        mx = fmax(mx, h);
      }
      bsum27(acc, red);
      if (tid == 0) {
        int idx = 0;
        for (int r = 0; r < 6; r++) {
          bp[r] = red[21 + r];
          for (int c = r; c < 6; c++) { Hpp[6 * r + c] = red[idx]; Hpp[6 * c + r] = red[idx]; idx++; }
        }
      }
      __syncthreads();
      {  // block max
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        if (tid == 0) {
          double m = 0;
          for (int w = 0; w < NT / 32; w++) m = fmax(m, red[w]);
          for (int r = 0; r < 6; r++) m = fmax(m, fabs(Hpp[7 * r]));
          lm_begin_iteration(&ctl, it, m, -1.0);
        }
        __syncthreads();
      }
      // ---- trials
      while (true) {
        const double lambda = ctl.lambda;
#pragma unroll
        for (int k = 0; k < 32; k++) acc[k] = 0;
        for (int i = tid; i < n; i += NT) {
          if (a.level[i] != 0) continue;
          const double e0 = c_e0[i], e1 = c_e1[i], w = c_w[i];
          const double J[12] = {c_J[i], c_J[n + i], c_J[2 * n + i], c_J[3 * n + i], 0.0, c_J[4 * n + i],
                                c_J[5 * n + i], c_J[6 * n + i], c_J[7 * n + i], 0.0, c_J[8 * n + i], c_J[9 * n + i]};
          const double h = w + a.info_p + lambda;
          const double p0 = fcur[2 * i] - (double)a.flow_xy[2 * i], p1 = fcur[2 * i + 1] - (double)a.flow_xy[2 * i + 1];
          const double bl0 = -w * e0 - a.info_p * p0, bl1 = -w * e1 - a.info_p * p1;
          const double s = w * w / h, g = w / h;
          int idx = 0;
#pragma unroll
          for (int r = 0; r < 6; r++) {
            acc[21 + r] += g * (J[r] * bl0 + J[6 + r] * bl1);
#pragma unroll
            for (int c = r; c < 6; c++) acc[idx++] += s * (J[r] * J[c] + J[6 + r] * J[6 + c]);
          }
        }
        bsum27(acc, red);
        if (tid < 36) {  // reduced 6x6 system
          const int r = tid / 6, c = tid - 6 * r;
          const int lo = r < c ? r : c, hi = r < c ? c : r;
          const int idx = lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
          Ss[tid] = Hpp[tid] - red[idx] + ((r == c) ? lambda : 0.0);
          if (tid < 6) bs[tid] = bp[tid] - red[21 + tid];
        }
        __syncthreads();
        if (tid == 0) {
          double x[6];
          const bool ok = solve6_spd(Ss, bs, x);
          if (ok) for (int k = 0; k < 6; k++) xp[k] = x[k];
          ctl.fail = ok ? 0 : 1;  // on failure x keeps its previous content
          se3_exp_mul(xp, Tc, T[cur ^ 1]);
        }
        __syncthreads();
        const int failed = ctl.fail;
        const PoseQ Tt = T[cur ^ 1];
        double xr[6];
#pragma unroll
        for (int k = 0; k < 6; k++) xr[k] = xp[k];
        // ---- flow increments, trial state, chi2 and scale
        double v[32];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = 0;
        for (int i = tid; i < n; i += NT) {
          const double p0 = fcur[2 * i] - (double)a.flow_xy[2 * i], p1 = fcur[2 * i + 1] - (double)a.flow_xy[2 * i + 1];
          double bl0 = -a.info_p * p0, bl1 = -a.info_p * p1, h = a.info_p + lambda;
          double x0, x1;
          const bool active = a.level[i] == 0;
          if (active) {
            const double e0 = c_e0[i], e1 = c_e1[i], w = c_w[i];
            h += w;
            bl0 -= w * e0; bl1 -= w * e1;
            const double jx0 = c_J[i] * xr[0] + c_J[n + i] * xr[1] + c_J[2 * n + i] * xr[2] + c_J[3 * n + i] * xr[3] + c_J[4 * n + i] * xr[5];
            const double jx1 = c_J[5 * n + i] * xr[0] + c_J[6 * n + i] * xr[1] + c_J[7 * n + i] * xr[2] + c_J[8 * n + i] * xr[4] + c_J[9 * n + i] * xr[5];
            x0 = (bl0 - w * jx0) / h;
            x1 = (bl1 - w * jx1) / h;
          } else {
            x0 = bl0 / h;
            x1 = bl1 / h;
          }
          if (failed) { x0 = a.xl[2 * i]; x1 = a.xl[2 * i + 1]; }
          else { a.xl[2 * i] = x0; a.xl[2 * i + 1] = x1; }
          const double f0 = fcur[2 * i] + x0, f1 = fcur[2 * i + 1] + x1;
          ftrial[2 * i] = f0; ftrial[2 * i + 1] = f1;
          v[1] += x0 * (lambda * x0 + bl0) + x1 * (lambda * x1 + bl1);
          if (active) {
            double e[2], r0, w;
            const double ft[2] = {f0, f1};
            proj_edge(a, Tt, i, ft, e, nullptr);
            a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
            const double c = (e[0] * e[0] + e[1] * e[1]) * a.info_f;
            if (robust) { huber_dev(c, a.delta, r0, w); v[0] += r0; } else v[0] += c;
          }
          const double q0 = f0 - (double)a.flow_xy[2 * i], q1 = f1 - (double)a.flow_xy[2 * i + 1];
          v[0] += (q0 * q0 + q1 * q1) * a.info_p;
        }
        bsum27(v, red);
        if (tid == 0) {
          double sc = red[1];
          for (int r = 0; r < 6; r++) sc += xp[r] * (lambda * xp[r] + bp[r]);
          lm_trial(&ctl, red[0], sc, failed);
        }
        __syncthreads();
        if (!lm_more_trials(&ctl)) break;
      }
      if (tid == 0) lm_end_iteration(&ctl, it, -1.0, a.rec ? a.rec + (size_t)round * VIDO_LM_REC : nullptr);
      __syncthreads();
    }
    // ---- classify (src/Optimizer.cc:2751-2794); accepted state -> buffer 0 for the next round
    const int cur = ctl.cur;
    const float th = (round == 0) ? a.th0 : a.th1;
    double vb[32];
#pragma unroll
    for (int k = 0; k < 32; k++) vb[k] = 0;
    const PoseQ Tf = T[cur];
    for (int i = tid; i < n; i += NT) {
      const double f0 = flowbuf[cur][2 * i], f1 = flowbuf[cur][2 * i + 1];
      if (a.level[i] != 0) {  // demoted edges are re-evaluated at the final estimate
        double e[2];
        const double ft[2] = {f0, f1};
        proj_edge(a, Tf, i, ft, e, nullptr);
        a.eProj[2 * i] = e[0]; a.eProj[2 * i + 1] = e[1];
      }
      const float chi2 = (float)((a.eProj[2 * i] * a.eProj[2 * i] + a.eProj[2 * i + 1] * a.eProj[2 * i + 1]) * a.info_f);
      if (chi2 > th) { a.level[i] = 1; vb[0] += 1.0; }
      else a.level[i] = 0;
      flowbuf[0][2 * i] = f0; flowbuf[0][2 * i + 1] = f1;
    }
    bsum27(vb, red);
    if (tid == 0) {
      s_nbad = (int)(red[0] + 0.5);
      if (round == 2) s_robust = 0;
      if (a.ctl) a.ctl[round] = ctl;
      T[0] = T[cur];  // keep the final pose of this round (overwritten by the reset unless it is the last round)
    }
    __syncthreads();
  }
  // ---- outputs: pose (SE3Quat -> homogeneous -> float), refined flows, inlier flags
  if (tid == 0) {
    const PoseQ& F = T[0];
    const double w = F.q[0], x = F.q[1], y = F.q[2], z = F.q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x,
                 txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) a.Tcw_out[4 * r + c] = (float)R[3 * r + c];
      a.Tcw_out[4 * r + 3] = (float)F.t[r];
    }
    a.Tcw_out[12] = 0.f; a.Tcw_out[13] = 0.f; a.Tcw_out[14] = 0.f; a.Tcw_out[15] = 1.f;
    *a.n_inliers = n - s_nbad;
  }
  for (int i = tid; i < n; i += NT) {
    a.flow_out[2 * i] = (float)flowbuf[0][2 * i];
    a.flow_out[2 * i + 1] = (float)flowbuf[0][2 * i + 1];
    a.inlier[i] = a.level[i] == 0;
  }
}

// =========================================================================================================
// host side
// =========================================================================================================
struct PoWorkspace {
  int capN = 0, capProblems = 0;
  // input block (one H2D): PoArgs[capProblems] then per problem obs | flow | depth | Tinit | Tlast
  // output block (one D2H): per problem Tout | n_inliers | flow_out | inlier, then LmCtl[4] per problem, then records
  char *d_in = nullptr, *d_out = nullptr, *h_in = nullptr, *h_out = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
  double *Xw = nullptr, *flow = nullptr, *xl = nullptr, *eProj = nullptr;
  int* level = nullptr;
  // device-chained camera problem (po_chain_enqueue)
  char* c_block = nullptr;
  double *c_Xw = nullptr, *c_flowd = nullptr, *c_xl = nullptr, *c_eProj = nullptr;   // chain-private scratch
  int* c_level = nullptr;
  PoArgs* c_args[2] = {nullptr, nullptr};
  float *c_obs = nullptr, *c_flow = nullptr, *c_dep = nullptr, *c_T = nullptr, *c_flowout = nullptr;
  int *c_inl = nullptr, *c_ninl = nullptr, *c_n = nullptr;
  LmCtl* c_ctl = nullptr;
  int c_cap = 0;
  bool c_ready[2] = {false, false};
};

static size_t pal(size_t v) { return (v + 63) & ~(size_t)63; }
template <class T>
static T* pcarve(char*& p, size_t n) { T* r = (T*)p; p += pal(sizeof(T) * n); return r; }

int po_setup(vido_ctx* ctx, int capN, int capProblems) {
  PoWorkspace* ws = new PoWorkspace();
  ctx->po = ws;
  ws->capN = capN;
  ws->capProblems = capProblems;
  const size_t N = (size_t)capN * capProblems;
  ws->in_bytes = pal(sizeof(PoArgs) * capProblems) + capProblems * (pal(8 * (size_t)capN) * 2 + pal(4 * (size_t)capN) + 2 * pal(64));
  ws->out_bytes = capProblems * (pal(64) + pal(4) + pal(8 * (size_t)capN) + pal(4 * (size_t)capN)) + pal(sizeof(LmCtl) * 4 * capProblems) +
                  pal(sizeof(LmRec) * VIDO_LM_REC * 4 * capProblems);
  VIDO_CUDA(cudaMalloc(&ws->d_in, ws->in_bytes));
  VIDO_CUDA(cudaMalloc(&ws->d_out, ws->out_bytes));
  VIDO_CUDA(cudaMallocHost(&ws->h_in, ws->in_bytes));
  VIDO_CUDA(cudaMallocHost(&ws->h_out, ws->out_bytes));
  VIDO_CUDA(cudaMalloc(&ws->Xw, sizeof(double) * 3 * N));
  VIDO_CUDA(cudaMalloc(&ws->flow, sizeof(double) * 4 * N));
  VIDO_CUDA(cudaMalloc(&ws->xl, sizeof(double) * 2 * N));
  VIDO_CUDA(cudaMalloc(&ws->eProj, sizeof(double) * 2 * N));
  VIDO_CUDA(cudaMalloc(&ws->level, sizeof(int) * N));
  VIDO_CUDA(cudaFuncSetAttribute(poseopt_flow2_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * PO_NC * PO_CACHE_CAP)));
  return VIDO_OK;
}

void po_teardown(vido_ctx* ctx) {
  PoWorkspace* ws = (PoWorkspace*)ctx->po;
  if (!ws) return;
  cudaFree(ws->d_in); cudaFree(ws->d_out); cudaFreeHost(ws->h_in); cudaFreeHost(ws->h_out);
  cudaFree(ws->Xw); cudaFree(ws->flow); cudaFree(ws->xl); cudaFree(ws->eProj); cudaFree(ws->level); cudaFree(ws->c_block);
  cudaFree(ws->c_Xw); cudaFree(ws->c_flowd); cudaFree(ws->c_xl); cudaFree(ws->c_eProj); cudaFree(ws->c_level);
  delete ws;
  ctx->po = nullptr;
}

int po_flow2_host(vido_ctx* ctx, vido_poseopt_problem* prs, int nproblems, vido_lm_stats* stats) {
  PoWorkspace* ws = (PoWorkspace*)ctx->po;
  if (nproblems < 1 || nproblems > ws->capProblems) { ctx->err = "too many pose problems"; return VIDO_ERR_CAPACITY; }
  cudaStream_t s = ctx->stream;
  // identical carving of the host (pinned) and device blocks
  char *hi = ws->h_in, *di = ws->d_in, *ho = ws->h_out, *dq = ws->d_out;
  PoArgs* h_args = pcarve<PoArgs>(hi, nproblems);
  PoArgs* d_args = pcarve<PoArgs>(di, nproblems);
  struct Out { float* T; int* ninl; float* flow; int* inl; };
  std::vector<Out> outs(nproblems);
  for (int k = 0; k < nproblems; k++) {
    vido_poseopt_problem& p = prs[k];
    if (p.n > ws->capN || p.n < 0 || p.rounds > 4 || p.rounds < 1) { ctx->err = "pose problem exceeds capacity"; return VIDO_ERR_CAPACITY; }
    const size_t off = (size_t)k * ws->capN;
    PoArgs& a = h_args[k];
    memset(&a, 0, sizeof a);
    a.n = p.n;
    float* h_obs = pcarve<float>(hi, 2 * (size_t)p.n);   a.obs_xy = pcarve<float>(di, 2 * (size_t)p.n);
    float* h_flow = pcarve<float>(hi, 2 * (size_t)p.n);  a.flow_xy = pcarve<float>(di, 2 * (size_t)p.n);
    float* h_dep = pcarve<float>(hi, p.n);               a.depth = pcarve<float>(di, p.n);
    float* h_Ti = pcarve<float>(hi, 16);                 a.Tcw_init = pcarve<float>(di, 16);
    float* h_Tl = pcarve<float>(hi, 16);                 a.Tcw_last = pcarve<float>(di, 16);
    if (p.n) {
      memcpy(h_obs, p.obs_xy, sizeof(float) * 2 * p.n);
      memcpy(h_flow, p.flow_xy, sizeof(float) * 2 * p.n);
      memcpy(h_dep, p.depth, sizeof(float) * p.n);
    }
    memcpy(h_Ti, p.Tcw_init, sizeof(float) * 16);
    memcpy(h_Tl, p.Tcw_last, sizeof(float) * 16);
    a.fx = p.fx; a.fy = p.fy; a.cx = p.cx; a.cy = p.cy;
    a.info_f = (p.info_flow == 0.1f) ? 0.1 : (double)p.info_flow;   // Matrix2d literals of the reference are doubles
    a.info_p = (p.info_prior == 0.3f) ? 0.3 : (p.info_prior == 0.5f ? 0.5 : (double)p.info_prior);
    a.delta = (double)sqrtf(p.rp_thres);
    a.th0 = p.rp_thres; a.th1 = p.chi2_th;
    a.rounds = p.rounds; a.its = p.its;
    a.Xw = ws->Xw + 3 * off; a.flow = ws->flow + 4 * off; a.xl = ws->xl + 2 * off; a.eProj = ws->eProj + 2 * off;
    a.level = ws->level + off;
    outs[k].T = pcarve<float>(ho, 16);                    a.Tcw_out = pcarve<float>(dq, 16);
    outs[k].ninl = pcarve<int>(ho, 1);                    a.n_inliers = pcarve<int>(dq, 1);
    outs[k].flow = pcarve<float>(ho, 2 * (size_t)p.n);    a.flow_out = pcarve<float>(dq, 2 * (size_t)p.n);
    outs[k].inl = pcarve<int>(ho, p.n);                   a.inlier = pcarve<int>(dq, p.n);
  }
  LmCtl* h_ctl = pcarve<LmCtl>(ho, 4 * (size_t)nproblems);
  LmCtl* d_ctl = pcarve<LmCtl>(dq, 4 * (size_t)nproblems);
  const size_t out_small = (size_t)(ho - ws->h_out);
  LmRec* h_rec = pcarve<LmRec>(ho, (size_t)4 * VIDO_LM_REC * nproblems);
  LmRec* d_rec = pcarve<LmRec>(dq, (size_t)4 * VIDO_LM_REC * nproblems);
  for (int k = 0; k < nproblems; k++) {
    h_args[k].ctl = d_ctl + 4 * k;
    h_args[k].rec = stats ? d_rec + (size_t)4 * VIDO_LM_REC * k : nullptr;
  }
  VIDO_CUDA(cudaMemcpyAsync(ws->d_in, ws->h_in, (size_t)(hi - ws->h_in), cudaMemcpyHostToDevice, s));
  cudaEventRecord(ctx->ev0, s);
  int nmax = 0;
  for (int k = 0; k < nproblems; k++) nmax = std::max(nmax, prs[k].n);
  if (nmax <= PO_CACHE_CAP)
    poseopt_flow2_fast_kernel<<<nproblems, PO_FAST_THREADS, sizeof(double) * PO_NC * (size_t)std::max(nmax, 1), s>>>(d_args);
  else
    poseopt_flow2_kernel<<<nproblems, PO_THREADS, 0, s>>>(d_args);
  cudaEventRecord(ctx->ev1, s);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  VIDO_CUDA(cudaMemcpyAsync(ws->h_out, ws->d_out, stats ? (size_t)(ho - ws->h_out) : out_small, cudaMemcpyDeviceToHost, s));
  if (ctx->idle_work) { std::function<void()> f; f.swap(ctx->idle_work); f(); }   // host work hidden behind the kernel
  VIDO_CUDA(cudaStreamSynchronize(s));
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) { ctx->t_ms[2] += ms; ctx->t_n[2]++; }
  }
  for (int k = 0; k < nproblems; k++) {
    vido_poseopt_problem& p = prs[k];
    memcpy(p.Tcw_out, outs[k].T, sizeof(float) * 16);
    p.n_inliers = *outs[k].ninl;
    if (p.n && p.flow_out) memcpy(p.flow_out, outs[k].flow, sizeof(float) * 2 * p.n);
    if (p.n && p.inlier) memcpy(p.inlier, outs[k].inl, sizeof(int) * p.n);
  }
  if (stats) {
    for (int k = 0; k < nproblems; k++)
      for (int r = 0; r < prs[k].rounds; r++) {
        vido_lm_stats& st = stats[4 * k + r];
        const LmCtl& c = h_ctl[4 * k + r];
        st.iterations = c.iterations; st.n_records = c.n_records; st.total_trials = c.total_trials;
        for (int i = 0; i < c.n_records && i < VIDO_LM_MAX_RECORDS; i++) {
          const LmRec& rr = h_rec[((size_t)4 * k + r) * VIDO_LM_REC + i];
          st.rec[i].chi2 = rr.chi2; st.rec[i].lambda = rr.lambda; st.rec[i].trials = rr.trials;
        }
      }
  }
  return VIDO_OK;
}

// =========================================================================================================
// device-chained camera pose optimisation: the problem is gathered on the device from the init-model inliers
// =========================================================================================================
// Tracking::Track between GetInitModelCam and PoseOptimizationFlow2Cam (src/Tracking.cc:1137-1160): the inliers' last-frame
// key points, flows and depths, in inlier order
__global__ void __launch_bounds__(1024) po_chain_prep_kernel(PoArgs* __restrict__ args, const int32_t* __restrict__ pnp_res,
                                                             const int32_t* __restrict__ ids, const float* __restrict__ keys,
                                                             const float* __restrict__ flow, const float* __restrict__ depth,
                                                             float* __restrict__ obs, float* __restrict__ fl, float* __restrict__ dep,
                                                             int32_t* __restrict__ n_out) {
  const int n = pnp_res[0];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int k = ids[i];
    obs[2 * i] = keys[2 * k]; obs[2 * i + 1] = keys[2 * k + 1];
    fl[2 * i] = flow[2 * k]; fl[2 * i + 1] = flow[2 * k + 1];
    dep[i] = depth[k];
  }
  if (threadIdx.x == 0) { args->n = n; *n_out = n; }
}

int po_chain_setup(vido_ctx* ctx, int cap) {
  PoWorkspace* ws = (PoWorkspace*)ctx->po;
  if (cap > ws->capN || cap > PO_CACHE_CAP) { ctx->err = "chain capacity exceeds the pose-optimisation capacity"; return VIDO_ERR_CAPACITY; }
  ws->c_cap = cap;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_args = 0, o_obs = o_args + al(2 * sizeof(PoArgs)), o_flow = o_obs + al(8 * (size_t)cap), o_dep = o_flow + al(8 * (size_t)cap),
               o_T = o_dep + al(4 * (size_t)cap), o_fo = o_T + 256, o_inl = o_fo + al(8 * (size_t)cap), o_ninl = o_inl + al(4 * (size_t)cap),
               o_n = o_ninl + 256, o_ctl = o_n + 256, total = o_ctl + al(sizeof(LmCtl) * 4);
  VIDO_CUDA(cudaMalloc(&ws->c_block, total));
  VIDO_CUDA(cudaMemset(ws->c_block, 0, total));
  // the chain's own per-vertex scratch: it may run beside a host-driven batch (the objects of a frame)
  VIDO_CUDA(cudaMalloc(&ws->c_Xw, sizeof(double) * 3 * (size_t)cap));
  VIDO_CUDA(cudaMalloc(&ws->c_flowd, sizeof(double) * 4 * (size_t)cap));
  VIDO_CUDA(cudaMalloc(&ws->c_xl, sizeof(double) * 2 * (size_t)cap));
  VIDO_CUDA(cudaMalloc(&ws->c_eProj, sizeof(double) * 2 * (size_t)cap));
  VIDO_CUDA(cudaMalloc(&ws->c_level, sizeof(int) * (size_t)cap));
  ws->c_args[0] = (PoArgs*)(ws->c_block + o_args); ws->c_args[1] = ws->c_args[0] + 1;
  ws->c_obs = (float*)(ws->c_block + o_obs); ws->c_flow = (float*)(ws->c_block + o_flow); ws->c_dep = (float*)(ws->c_block + o_dep);
  ws->c_T = (float*)(ws->c_block + o_T); ws->c_flowout = (float*)(ws->c_block + o_fo); ws->c_inl = (int*)(ws->c_block + o_inl);
  ws->c_ninl = (int*)(ws->c_block + o_ninl); ws->c_n = (int*)(ws->c_block + o_n); ws->c_ctl = (LmCtl*)(ws->c_block + o_ctl);
  ws->c_ready[0] = ws->c_ready[1] = false;
  return VIDO_OK;
}

int po_chain_enqueue(vido_ctx* ctx, const ChainStateDev& st, int which, const ChainPnpOut& pnp, ChainPoOut* out) {
  PoWorkspace* ws = (PoWorkspace*)ctx->po;
  cudaStream_t s = ctx->stream;
  const vido_config& c = ctx->cfg;
  if (!ws->c_ready[which]) {
    vido_poseopt_problem p;
    vido_poseopt_default_params(&p);
    PoArgs a;
    memset(&a, 0, sizeof a);
    a.obs_xy = ws->c_obs; a.flow_xy = ws->c_flow; a.depth = ws->c_dep; a.Tcw_init = pnp.T; a.Tcw_last = st.Tcw;
    a.fx = c.fx; a.fy = c.fy; a.cx = c.cx; a.cy = c.cy;
    a.info_f = (p.info_flow == 0.1f) ? 0.1 : (double)p.info_flow;
    a.info_p = (p.info_prior == 0.3f) ? 0.3 : (p.info_prior == 0.5f ? 0.5 : (double)p.info_prior);
    a.delta = (double)sqrtf(p.rp_thres);
    a.th0 = p.rp_thres; a.th1 = p.chi2_th;
    a.rounds = p.rounds; a.its = p.its;
    a.Xw = ws->c_Xw; a.flow = ws->c_flowd; a.xl = ws->c_xl; a.eProj = ws->c_eProj; a.level = ws->c_level;
    a.Tcw_out = ws->c_T; a.n_inliers = ws->c_ninl; a.flow_out = ws->c_flowout; a.inlier = ws->c_inl;
    a.ctl = ws->c_ctl; a.rec = nullptr;
    VIDO_CUDA(cudaMemcpyAsync(ws->c_args[which], &a, sizeof a, cudaMemcpyHostToDevice, s));
    ws->c_ready[which] = true;
  }
  po_chain_prep_kernel<<<1, 1024, 0, s>>>(ws->c_args[which], pnp.res, pnp.ids, st.keys, st.flow, st.depth, ws->c_obs, ws->c_flow, ws->c_dep, ws->c_n);
  poseopt_flow2_fast_kernel<<<1, PO_FAST_THREADS, sizeof(double) * PO_NC * (size_t)ws->c_cap, s>>>(ws->c_args[which]);
  ctx->launches += 2;
  VIDO_CUDA(cudaGetLastError());
  out->T = ws->c_T; out->flow = ws->c_flowout; out->inl = ws->c_inl; out->ninl = ws->c_ninl; out->n = ws->c_n;
  return VIDO_OK;
}
