// projopt_kernels.cu -- reprojection-only pose optimisation (the bJoint == false branch of Tracking::Track), one CTA per problem,
// the whole Levenberg-Marquardt optimisation in one launch.
//
// Replaces:
//   Optimizer::PoseOptimizationNew      src/Optimizer.cc:2180-2334  (kind 0: EdgeSE3ProjectXYZOnlyPose, Huber sqrt(rp_thres), 100 its)
//   Optimizer::PoseOptimizationObjMot   src/Optimizer.cc:2826-3035  (kind 1: EdgeSE3ProjectXYZOnlyObjMotion, P = K Tcw, 200 its)
//   edge math                           g2o/types/types_six_dof_expmap.cpp:266-296, 394-441
//   VertexSE3Expmap::oplusImpl          g2o/types/types_six_dof_expmap.h:81-84
//   BlockSolver_6_3 + LinearSolverDense (6x6), LM driver: lm_device.h
// Per LM step the edges are swept once by all threads (2x6 Jacobian, robust weight), the 27 sums of the normal equations are
// reduced in a fixed order, one thread solves the 6x6 system and steps the LM state machine.  The outlier test after the
// optimisation uses the errors of the LAST trial state, like the reference (g2o keeps the edge errors of the last
// computeActiveErrors call, also when that trial was rejected).
#include <algorithm>
#include <cstring>
#include <vector>

#include "ctx.h"
#include "lm_device.h"
#include "se3q_math.h"

namespace {

using se3q::PoseQ;
constexpr int PJ_THREADS = 256;

struct ProjArgs {
  int n, kind, its;
  const float* obs;   // [n][2]
  const float* pts;   // [n][3]
  PoseQ Tinit;
  double fx, fy, cx, cy, delta, P[12];
  float th;
  float* T_out;       // [16]
  int* inlier;        // [n]
  int* n_inliers;
  LmCtl* ctl;
  LmRec* rec;
};

__device__ __forceinline__ void proj_eval(const ProjArgs& a, const PoseQ& T, int i, double* e, double* J) {
  const double X[3] = {(double)a.pts[3 * i], (double)a.pts[3 * i + 1], (double)a.pts[3 * i + 2]};
  double pc[3];
  se3q::q_rot(T.q, X, pc);
  pc[0] += T.t[0]; pc[1] += T.t[1]; pc[2] += T.t[2];
  const double x = pc[0], y = pc[1], z = pc[2];
  if (a.kind == 0) {
    e[0] = (double)a.obs[2 * i] - (x / z * a.fx + a.cx);
    e[1] = (double)a.obs[2 * i + 1] - (y / z * a.fy + a.cy);
    if (!J) return;
    const double invz = 1.0 / z, invz_2 = invz * invz;
    J[0] = x * y * invz_2 * a.fx; J[1] = -(1 + (x * x * invz_2)) * a.fx; J[2] = y * invz * a.fx;
    J[3] = -invz * a.fx;          J[4] = 0;                              J[5] = x * invz_2 * a.fx;
    J[6] = (1 + y * y * invz_2) * a.fy; J[7] = -x * y * invz_2 * a.fy;   J[8] = -x * invz * a.fy;
    J[9] = 0;                     J[10] = -invz * a.fy;                  J[11] = y * invz_2 * a.fy;
    return;
  }
  const double* P = a.P;
  const double m1 = P[0] * x + P[1] * y + P[2] * z + P[3], m2 = P[4] * x + P[5] * y + P[6] * z + P[7],
               m3 = P[8] * x + P[9] * y + P[10] * z + P[11];
  const double invm3 = 1.0 / m3;
  e[0] = (double)a.obs[2 * i] - m1 * invm3;
  e[1] = (double)a.obs[2 * i + 1] - m2 * invm3;
  if (!J) return;
  const double invm3_2 = invm3 * invm3;
  double t[6];
  for (int c = 0; c < 3; c++) {
    t[c] = invm3_2 * (P[c] * m3 - P[8 + c] * m1);
    t[3 + c] = invm3_2 * (P[4 + c] * m3 - P[8 + c] * m2);
  }
  for (int r = 0; r < 2; r++) {
    const double* tr = t + 3 * r;
    J[6 * r + 0] = -1.0 * (y * tr[2] - z * tr[1]);
    J[6 * r + 1] = -1.0 * (z * tr[0] - x * tr[2]);
    J[6 * r + 2] = -1.0 * (x * tr[1] - y * tr[0]);
    J[6 * r + 3] = -1.0 * tr[0]; J[6 * r + 4] = -1.0 * tr[1]; J[6 * r + 5] = -1.0 * tr[2];
  }
}

template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* sm) {   // fixed order; result in sm[0..NV)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < NV; k++) sm[(warp + 1) * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < PJ_THREADS / 32; w++) s += sm[(w + 1) * NV + threadIdx.x];
    sm[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ double chi2_at(const ProjArgs& a, const PoseQ& T, double* sm) {
  double acc[1] = {0};
  for (int i = threadIdx.x; i < a.n; i += PJ_THREADS) {
    double e[2];
    proj_eval(a, T, i, e, nullptr);
    const double c = e[0] * e[0] + e[1] * e[1];
    if (a.kind == 0) {
      const double dsqr = a.delta * a.delta;
      acc[0] += c <= dsqr ? c : 2 * sqrt(c) * a.delta - dsqr;
    } else acc[0] += c;
  }
  block_sum<1>(acc, sm);
  return sm[0];
}

__global__ void __launch_bounds__(PJ_THREADS) projopt_kernel(const ProjArgs* __restrict__ args) {
  const ProjArgs& a = args[blockIdx.x];
  __shared__ double sm[(PJ_THREADS / 32 + 1) * 27];
  __shared__ PoseQ T[2], Tlast;
  __shared__ double xs[6], bsv[6];
  __shared__ int s_fail;
  LmCtl* c = a.ctl;
  const int tid = threadIdx.x;
  if (tid == 0) {
    lm_reset(c);
    T[0] = a.Tinit; T[1] = a.Tinit; Tlast = a.Tinit;
    for (int k = 0; k < 6; k++) xs[k] = 0;
  }
  __syncthreads();
  for (int it = 0; it < a.its; it++) {
    if (c->stop_flag || !c->ok) break;
    const int cur = c->cur;
    if (it == 0) {
      const double chi = chi2_at(a, T[cur], sm);
      if (tid == 0) c->currentChi = chi;
      __syncthreads();
    }
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    for (int i = tid; i < a.n; i += PJ_THREADS) {
      double e[2], J[12];
      proj_eval(a, T[cur], i, e, J);
      double w = 1.0;
      if (a.kind == 0) {
        const double cc = e[0] * e[0] + e[1] * e[1], dsqr = a.delta * a.delta;
        if (cc > dsqr) w = a.delta / sqrt(cc);
      }
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        acc[21 + r] += -w * (J[r] * e[0] + J[6 + r] * e[1]);
#pragma unroll
        for (int q = r; q < 6; q++) acc[idx++] += w * (J[r] * J[q] + J[6 + r] * J[6 + q]);
      }
    }
    block_sum<27>(acc, sm);
    __shared__ double H[36];
    if (tid == 0) {
      int idx = 0;
      double md = 0;
      for (int r = 0; r < 6; r++)
        for (int q = r; q < 6; q++) { H[6 * r + q] = sm[idx]; H[6 * q + r] = sm[idx]; idx++; }
      for (int r = 0; r < 6; r++) { bsv[r] = sm[21 + r]; md = fmax(md, fabs(H[7 * r])); }
      lm_begin_iteration(c, it, md, -1.0);
    }
    __syncthreads();
    while (true) {
      const double lambda = c->lambda;
      if (tid == 0) {
        // LDL^T with the positivity test of the dense solver; a failure leaves x as it was
        double S[36], L[36], D[6], y[6];
        for (int k = 0; k < 36; k++) { S[k] = H[k] + ((k % 7 == 0) ? lambda : 0.0); L[k] = 0; }
        bool ok = true;
        for (int j = 0; j < 6 && ok; j++) {
          double dd = S[7 * j];
          for (int k = 0; k < j; k++) dd -= L[6 * j + k] * L[6 * j + k] * D[k];
          if (!(dd > 0)) { ok = false; break; }
          D[j] = dd; L[7 * j] = 1;
          for (int i = j + 1; i < 6; i++) {
            double s2 = S[6 * i + j];
            for (int k = 0; k < j; k++) s2 -= L[6 * i + k] * L[6 * j + k] * D[k];
            L[6 * i + j] = s2 / dd;
          }
        }
        if (ok) {
          for (int i = 0; i < 6; i++) { double s2 = bsv[i]; for (int k = 0; k < i; k++) s2 -= L[6 * i + k] * y[k]; y[i] = s2; }
          for (int i = 0; i < 6; i++) y[i] /= D[i];
          for (int i = 5; i >= 0; i--) { double s2 = y[i]; for (int k = i + 1; k < 6; k++) s2 -= L[6 * k + i] * xs[k]; xs[i] = s2; }
        }
        s_fail = ok ? 0 : 1;
        se3q::se3_exp_mul(xs, T[cur], T[cur ^ 1]);
        Tlast = T[cur ^ 1];
      }
      __syncthreads();
      const double chi = chi2_at(a, T[cur ^ 1], sm);
      if (tid == 0) {
        double sc = 0;
        for (int j = 0; j < 6; j++) sc += xs[j] * (lambda * xs[j] + bsv[j]);
        lm_trial(c, chi, sc, s_fail);
      }
      __syncthreads();
      if (!lm_more_trials(c)) break;
    }
    if (tid == 0) lm_end_iteration(c, it, -1.0, a.rec);
    __syncthreads();
  }
  // outliers: chi2 of the last evaluated state against rp_thres (float comparison like the reference)
  int bad = 0;
  for (int i = tid; i < a.n; i += PJ_THREADS) {
    double e[2];
    proj_eval(a, Tlast, i, e, nullptr);
    const float chi2 = (float)(e[0] * e[0] + e[1] * e[1]);
    const int out = chi2 > a.th ? 1 : 0;
    a.inlier[i] = out ? 0 : 1;
    bad += out;
  }
  double accb[1] = {(double)bad};
  block_sum<1>(accb, sm);
  if (tid == 0) {
    *a.n_inliers = a.n - (int)sm[0];
    const PoseQ& Tf = T[c->cur];
    const double w = Tf.q[0], x = Tf.q[1], y = Tf.q[2], z = Tf.q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    float* o = a.T_out;
    o[0] = (float)(1 - (tyy + tzz)); o[1] = (float)(txy - twz);       o[2] = (float)(txz + twy);       o[3] = (float)Tf.t[0];
    o[4] = (float)(txy + twz);       o[5] = (float)(1 - (txx + tzz)); o[6] = (float)(tyz - twx);       o[7] = (float)Tf.t[1];
    o[8] = (float)(txz - twy);       o[9] = (float)(tyz + twx);       o[10] = (float)(1 - (txx + tyy)); o[11] = (float)Tf.t[2];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  }
}

void quat_from_f32(const float* T, PoseQ& o) {   // Converter::toSE3Quat
  const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double q[4];
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0); q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    double v[3];
    v[i] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[3 * k + j] - R[3 * j + k]) * t;
    v[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    v[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
  }
  if (q[0] < 0) for (int m = 0; m < 4; m++) q[m] = -q[m];
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int m = 0; m < 4; m++) o.q[m] = q[m] / n;
  o.t[0] = T[3]; o.t[1] = T[7]; o.t[2] = T[11];
}

}  // namespace

void projopt_default_params(vido_projopt_problem* p, int kind) { p->kind = kind; p->rp_thres = 0.01f; p->its = kind == 0 ? 100 : 200; }

int projopt_host(vido_ctx* ctx, vido_projopt_problem* prs, int nproblems, vido_lm_stats* stats) {
  if (nproblems < 1) return VIDO_OK;
  cudaStream_t s = ctx->stream;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = al(sizeof(ProjArgs) * (size_t)nproblems);
  std::vector<size_t> o_obs(nproblems), o_pts(nproblems), o_T(nproblems), o_inl(nproblems), o_n(nproblems), o_ctl(nproblems), o_rec(nproblems);
  std::vector<int> active;
  for (int k = 0; k < nproblems; k++) {
    vido_projopt_problem& p = prs[k];
    if (p.n < 0 || (p.kind != 0 && p.kind != 1)) { ctx->err = "bad reprojection problem"; return VIDO_ERR_ARG; }
    if (p.n < 3) {   // "if(nInitialCorrespondences<3) return": kind 0 leaves the pose, kind 1 returns identity
      if (p.kind == 0) memcpy(p.T_out, p.T_init, sizeof(float) * 16);
      else { memset(p.T_out, 0, sizeof(float) * 16); p.T_out[0] = p.T_out[5] = p.T_out[10] = p.T_out[15] = 1.f; }
      for (int i = 0; i < p.n; i++) if (p.inlier) p.inlier[i] = 1;
      p.n_inliers = 0;
      if (stats) { stats[k].iterations = -1; stats[k].n_records = 0; stats[k].total_trials = 0; }
      continue;
    }
    active.push_back(k);
    const size_t n = (size_t)p.n;
    o_obs[k] = off; off += al(8 * n); o_pts[k] = off; off += al(12 * n);
    o_T[k] = off; off += al(64); o_inl[k] = off; off += al(4 * n); o_n[k] = off; off += al(4);
    o_ctl[k] = off; off += al(sizeof(LmCtl)); o_rec[k] = off; off += al(sizeof(LmRec) * VIDO_LM_REC);
  }
  if (active.empty()) return VIDO_OK;
  char* base = (char*)vido_scratch(ctx, 1, off);
  if (!base) { ctx->err = "projection-only optimisation: device allocation failed"; return VIDO_ERR_CUDA; }
  std::vector<char> host(off, 0);
  ProjArgs* ha = (ProjArgs*)host.data();
  for (size_t q = 0; q < active.size(); q++) {
    const int k = active[q];
    vido_projopt_problem& p = prs[k];
    ProjArgs& a = ha[q];
    memset(&a, 0, sizeof a);
    a.n = p.n; a.kind = p.kind; a.its = p.its;
    a.obs = (const float*)(base + o_obs[k]); a.pts = (const float*)(base + o_pts[k]);
    memcpy(host.data() + o_obs[k], p.obs_xy, 8 * (size_t)p.n);
    memcpy(host.data() + o_pts[k], p.pts3d, 12 * (size_t)p.n);
    quat_from_f32(p.T_init, a.Tinit);
    a.fx = p.fx; a.fy = p.fy; a.cx = p.cx; a.cy = p.cy;
    a.delta = (double)sqrtf(p.rp_thres); a.th = p.rp_thres;
    memcpy(a.P, p.P, sizeof a.P);
    a.T_out = (float*)(base + o_T[k]); a.inlier = (int*)(base + o_inl[k]); a.n_inliers = (int*)(base + o_n[k]);
    a.ctl = (LmCtl*)(base + o_ctl[k]); a.rec = (LmRec*)(base + o_rec[k]);
  }
  int rc = VIDO_OK;
  do {
    if (cudaMemcpyAsync(base, host.data(), off, cudaMemcpyHostToDevice, s) != cudaSuccess) { ctx->err = "projopt: upload failed"; rc = VIDO_ERR_CUDA; break; }
    projopt_kernel<<<(int)active.size(), PJ_THREADS, 0, s>>>((const ProjArgs*)base);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) { ctx->err = "projopt: launch failed"; rc = VIDO_ERR_CUDA; break; }
    if (cudaMemcpyAsync(host.data(), base, off, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
      ctx->err = "projopt: solve failed"; rc = VIDO_ERR_CUDA; break;
    }
    for (size_t q = 0; q < active.size(); q++) {
      const int k = active[q];
      vido_projopt_problem& p = prs[k];
      memcpy(p.T_out, host.data() + o_T[k], sizeof(float) * 16);
      if (p.inlier) memcpy(p.inlier, host.data() + o_inl[k], sizeof(int) * (size_t)p.n);
      p.n_inliers = *(const int*)(host.data() + o_n[k]);
      if (stats) {
        const LmCtl& c = *(const LmCtl*)(host.data() + o_ctl[k]);
        const LmRec* rec = (const LmRec*)(host.data() + o_rec[k]);
        stats[k].iterations = c.iterations; stats[k].n_records = c.n_records; stats[k].total_trials = c.total_trials;
        for (int i = 0; i < c.n_records && i < VIDO_LM_MAX_RECORDS; i++) { stats[k].rec[i].chi2 = rec[i].chi2; stats[k].rec[i].lambda = rec[i].lambda; stats[k].rec[i].trials = rec[i].trials; }
      }
    }
  } while (0);
  cudaStreamSynchronize(s);   // (the scratch belongs to the context)
  return rc;
}
