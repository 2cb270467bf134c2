// poseopt_oracle.cc -- CPU restatement of the per-frame joint flow+pose optimisation.  TEST INFRASTRUCTURE ONLY.
//
//   Optimizer::PoseOptimizationFlow2Cam        src/Optimizer.cc:2622-2824
//   VertexSE3Expmap::oplusImpl                 g2o/types/types_six_dof_expmap.h:81-84 (T <- exp(update) * T)
//   VertexSBAFlow                              g2o/types/types_sba.h:78-95
//   EdgeSE3ProjectFlow2 error / Jacobians      g2o/types/types_six_dof_expmap.h:436-476, types_six_dof_expmap.cpp:805-845
//   EdgeFlowPrior                              g2o/types/types_six_dof_expmap.h:414-432, types_six_dof_expmap.cpp:772-776
//   BlockSolver Schur path (flows marginalised) g2o/core/block_solver.hpp:367-486; LinearSolverDense (Eigen LDLT; here a
//   pivot-free LDL^T with the same positivity test)  g2o/solvers/linear_solver_dense.h:65-116
#include <cmath>
#include <cstring>
#include <vector>

#include "g2o_math.h"
#include "lm_oracle.h"
#include "vido_oracle.h"

namespace vo {

struct SE3Q { Quat q; V3 t; };

static V3 map_point(const SE3Q& T, const V3& p) { return quat_rotate(T.q, p) + T.t; }

struct Flow2System {
  int N = 0;
  SE3Q T;
  std::vector<double> flow;               // 2N
  std::vector<V3> Xw;
  std::vector<double> obs, flow0;         // 2N each
  std::vector<int> level;                 // 0 active, 1 outlier
  bool robust = true;
  double fx, fy, cx, cy, info_f, info_p, delta;
  std::vector<double> eProj, ePrior;      // 2N each (eProj only refreshed for active edges)
  // system: pose block, flow blocks, cross blocks
  double Hpp[36], bp[6];
  std::vector<double> Hll, bl, Hpl;       // N*4, N*2, N*12 (6x2)
  std::vector<double> x;                  // 6 + 2N
  std::vector<SE3Q> bkT;
  std::vector<std::vector<double>> bkF;

  int num_vertices() const { return 1 + N; }
  void proj_error(int i, double* e) const {
    V3 pc = map_point(T, Xw[i]);
    e[0] = (obs[2 * i] + flow[2 * i]) - (pc.x / pc.z * fx + cx);
    e[1] = (obs[2 * i + 1] + flow[2 * i + 1]) - (pc.y / pc.z * fy + cy);
  }
  void compute_errors() {
    for (int i = 0; i < N; i++) {
      if (level[i] == 0) proj_error(i, &eProj[2 * i]);
      ePrior[2 * i] = flow[2 * i] - flow0[2 * i];
      ePrior[2 * i + 1] = flow[2 * i + 1] - flow0[2 * i + 1];
    }
  }
  double robust_chi2() const {
    double chi = 0, rho[3];
    for (int i = 0; i < N; i++) {
      if (level[i] == 0) {
        double c = (eProj[2 * i] * eProj[2 * i] + eProj[2 * i + 1] * eProj[2 * i + 1]) * info_f;
        if (robust) { huber(c, delta, rho); chi += rho[0]; } else chi += c;
      }
      chi += (ePrior[2 * i] * ePrior[2 * i] + ePrior[2 * i + 1] * ePrior[2 * i + 1]) * info_p;
    }
    return chi;
  }
  void build_system() {
    memset(Hpp, 0, sizeof Hpp);
    memset(bp, 0, sizeof bp);
    std::fill(Hll.begin(), Hll.end(), 0.0);
    std::fill(bl.begin(), bl.end(), 0.0);
    std::fill(Hpl.begin(), Hpl.end(), 0.0);
    double rho[3];
    for (int i = 0; i < N; i++) {
      if (level[i] == 0) {
        V3 pc = map_point(T, Xw[i]);
        const double x = pc.x, y = pc.y, z = pc.z, z2 = z * z;
        double J[2][6];
        J[0][0] = x * y / z2 * fx; J[0][1] = -(1 + (x * x / z2)) * fx; J[0][2] = y / z * fx;
        J[0][3] = -1. / z * fx;    J[0][4] = 0;                         J[0][5] = x / z2 * fx;
        J[1][0] = (1 + y * y / z2) * fy; J[1][1] = -x * y / z2 * fy;   J[1][2] = -x / z * fy;
        J[1][3] = 0;               J[1][4] = -1. / z * fy;             J[1][5] = y / z2 * fy;
        const double* e = &eProj[2 * i];
        double w = info_f;
        if (robust) { huber((e[0] * e[0] + e[1] * e[1]) * info_f, delta, rho); w *= rho[1]; }
        // flow vertex (Xi, identity Jacobian) and pose vertex (Xj)
        Hll[4 * i] += w; Hll[4 * i + 3] += w;
        bl[2 * i] += -w * e[0]; bl[2 * i + 1] += -w * e[1];
        for (int r = 0; r < 6; r++) {
          bp[r] += -w * (J[0][r] * e[0] + J[1][r] * e[1]);
          for (int c = 0; c < 6; c++) Hpp[6 * r + c] += w * (J[0][r] * J[0][c] + J[1][r] * J[1][c]);
          Hpl[12 * i + 2 * r] += w * J[0][r];
          Hpl[12 * i + 2 * r + 1] += w * J[1][r];
        }
      }
      const double* e = &ePrior[2 * i];
      Hll[4 * i] += info_p; Hll[4 * i + 3] += info_p;
      bl[2 * i] += -info_p * e[0]; bl[2 * i + 1] += -info_p * e[1];
    }
  }
  double max_diag() const {
    double m = 0;
    for (int r = 0; r < 6; r++) m = std::max(m, std::fabs(Hpp[7 * r]));
    for (int i = 0; i < N; i++) m = std::max(m, std::max(std::fabs(Hll[4 * i]), std::fabs(Hll[4 * i + 3])));
    return m;
  }
  bool solve(double lambda) {
    double S[36], bs[6];
    memcpy(S, Hpp, sizeof S);
    memcpy(bs, bp, sizeof bs);
    for (int r = 0; r < 6; r++) S[7 * r] += lambda;
    std::vector<double> Dinv(4 * (size_t)N);
    for (int i = 0; i < N; i++) {
      const double a = Hll[4 * i] + lambda, b = Hll[4 * i + 1], c = Hll[4 * i + 2], d = Hll[4 * i + 3] + lambda;
      const double det = a * d - b * c;
      double* Di = &Dinv[4 * i];
      Di[0] = d / det; Di[1] = -b / det; Di[2] = -c / det; Di[3] = a / det;
      if (level[i] != 0) continue;  // no cross block for demoted edges
      const double* B = &Hpl[12 * i];
      double BD[12];
      for (int r = 0; r < 6; r++) {
        BD[2 * r] = B[2 * r] * Di[0] + B[2 * r + 1] * Di[2];
        BD[2 * r + 1] = B[2 * r] * Di[1] + B[2 * r + 1] * Di[3];
      }
      const double db0 = Di[0] * bl[2 * i] + Di[1] * bl[2 * i + 1], db1 = Di[2] * bl[2 * i] + Di[3] * bl[2 * i + 1];
      for (int r = 0; r < 6; r++) {
        bs[r] -= B[2 * r] * db0 + B[2 * r + 1] * db1;
        for (int c2 = 0; c2 < 6; c2++) S[6 * r + c2] -= BD[2 * r] * B[2 * c2] + BD[2 * r + 1] * B[2 * c2 + 1];
      }
    }
    // LDL^T, positive test
    double L[36] = {0}, D[6];
    bool ok = true;
    for (int j = 0; j < 6 && ok; j++) {
      double d = S[7 * j];
      for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k] * D[k];
      if (!(d > 0)) { ok = false; break; }
      D[j] = d;
      L[7 * j] = 1;
      for (int i = j + 1; i < 6; i++) {
        double s = S[6 * i + j];
        for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k] * D[k];
        L[6 * i + j] = s / d;
      }
    }
    if (!ok) return false;  // x keeps its previous content (block_solver.hpp:455-456 returns before touching x_l)
    double y[6];
    for (int i = 0; i < 6; i++) {
      double s = bs[i];
      for (int k = 0; k < i; k++) s -= L[6 * i + k] * y[k];
      y[i] = s;
    }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) {
      double s = y[i];
      for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * x[k];
      x[i] = s;
    }
    for (int i = 0; i < N; i++) {
      double c0 = bl[2 * i], c1 = bl[2 * i + 1];
      if (level[i] == 0) {
        const double* B = &Hpl[12 * i];
        for (int r = 0; r < 6; r++) { c0 -= B[2 * r] * x[r]; c1 -= B[2 * r + 1] * x[r]; }
      }
      const double* Di = &Dinv[4 * i];
      x[6 + 2 * i] = Di[0] * c0 + Di[1] * c1;
      x[6 + 2 * i + 1] = Di[2] * c0 + Di[3] * c1;
    }
    return true;
  }
  void update() {
    Quat dq; V3 dt;
    se3_exp(&x[0], dq, dt);
    SE3Q n;
    n.t = dt + quat_rotate(dq, T.t);      // SE3Quat::operator* (se3quat.h:100-106)
    n.q = quat_normalized(quat_mul(dq, T.q));
    T = n;
    for (int i = 0; i < 2 * N; i++) flow[i] += x[6 + i];
  }
  void push() { bkT.push_back(T); bkF.push_back(flow); }
  void pop() { T = bkT.back(); flow = bkF.back(); bkT.pop_back(); bkF.pop_back(); }
  void discard_top() { bkT.pop_back(); bkF.pop_back(); }
  double compute_scale(double lambda) const {
    double s = 0;
    for (int j = 0; j < 6; j++) s += x[j] * (lambda * x[j] + bp[j]);
    for (int j = 0; j < 2 * N; j++) s += x[6 + j] * (lambda * x[6 + j] + bl[j]);
    return s;
  }
};

// ---- one SE3Expmap vertex, n reprojection edges (PoseOptimizationNew / PoseOptimizationObjMot)
struct ProjSystem {
  int N = 0, kind = 0;
  SE3Q T;
  std::vector<V3> Xw;
  std::vector<double> obs, err;
  bool robust = true;
  double fx, fy, cx, cy, delta, P[12];
  double H[36], b[6], x[6];
  std::vector<SE3Q> bk;

  int num_vertices() const { return 1; }
  void project(const V3& pc, double* uv) const {
    if (kind == 0) { uv[0] = pc.x / pc.z * fx + cx; uv[1] = pc.y / pc.z * fy + cy; return; }
    const double m1 = P[0] * pc.x + P[1] * pc.y + P[2] * pc.z + P[3], m2 = P[4] * pc.x + P[5] * pc.y + P[6] * pc.z + P[7],
                 m3 = P[8] * pc.x + P[9] * pc.y + P[10] * pc.z + P[11];
    const double invm3 = 1.0 / m3;
    uv[0] = m1 * invm3; uv[1] = m2 * invm3;
  }
  void edge_error(int i, double* e) const {
    double uv[2];
    project(map_point(T, Xw[i]), uv);
    e[0] = obs[2 * i] - uv[0]; e[1] = obs[2 * i + 1] - uv[1];
  }
  void compute_errors() { for (int i = 0; i < N; i++) edge_error(i, &err[2 * i]); }
  double robust_chi2() const {
    double chi = 0, rho[3];
    for (int i = 0; i < N; i++) {
      const double c = err[2 * i] * err[2 * i] + err[2 * i + 1] * err[2 * i + 1];
      if (robust) { huber(c, delta, rho); chi += rho[0]; } else chi += c;
    }
    return chi;
  }
  void jacobian(int i, double J[2][6]) const {
    const V3 pc = map_point(T, Xw[i]);
    const double x = pc.x, y = pc.y, z = pc.z;
    if (kind == 0) {
      const double invz = 1.0 / z, invz_2 = invz * invz;
      J[0][0] = x * y * invz_2 * fx; J[0][1] = -(1 + (x * x * invz_2)) * fx; J[0][2] = y * invz * fx;
      J[0][3] = -invz * fx;          J[0][4] = 0;                            J[0][5] = x * invz_2 * fx;
      J[1][0] = (1 + y * y * invz_2) * fy; J[1][1] = -x * y * invz_2 * fy;   J[1][2] = -x * invz * fy;
      J[1][3] = 0;                   J[1][4] = -invz * fy;                   J[1][5] = y * invz_2 * fy;
      return;
    }
    const double m1 = P[0] * x + P[1] * y + P[2] * z + P[3], m2 = P[4] * x + P[5] * y + P[6] * z + P[7],
                 m3 = P[8] * x + P[9] * y + P[10] * z + P[11];
    const double invm3 = 1.0 / m3, invm3_2 = invm3 * invm3;
    double t[2][3];
    for (int c = 0; c < 3; c++) {
      t[0][c] = invm3_2 * (P[c] * m3 - P[8 + c] * m1);
      t[1][c] = invm3_2 * (P[4 + c] * m3 - P[8 + c] * m2);
    }
    for (int r = 0; r < 2; r++) {
      J[r][0] = -1.0 * (y * t[r][2] - z * t[r][1]);
      J[r][1] = -1.0 * (z * t[r][0] - x * t[r][2]);
      J[r][2] = -1.0 * (x * t[r][1] - y * t[r][0]);
      J[r][3] = -1.0 * t[r][0]; J[r][4] = -1.0 * t[r][1]; J[r][5] = -1.0 * t[r][2];
    }
  }
  void build_system() {
    memset(H, 0, sizeof H);
    memset(b, 0, sizeof b);
    double rho[3];
    for (int i = 0; i < N; i++) {
      double J[2][6];
      jacobian(i, J);
      const double* e = &err[2 * i];
      double w = 1.0;
      if (robust) { huber(e[0] * e[0] + e[1] * e[1], delta, rho); w = rho[1]; }
      for (int r = 0; r < 6; r++) {
        b[r] += -w * (J[0][r] * e[0] + J[1][r] * e[1]);
        for (int c = 0; c < 6; c++) H[6 * r + c] += w * (J[0][r] * J[0][c] + J[1][r] * J[1][c]);
      }
    }
  }
  double max_diag() const { double m = 0; for (int r = 0; r < 6; r++) m = std::max(m, std::fabs(H[7 * r])); return m; }
  bool solve(double lambda) {
    double S[36], L[36] = {0}, D[6], y[6];
    memcpy(S, H, sizeof S);
    for (int r = 0; r < 6; r++) S[7 * r] += lambda;
    for (int j = 0; j < 6; j++) {
      double d = S[7 * j];
      for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k] * D[k];
      if (!(d > 0)) return false;   // x keeps its previous content
      D[j] = d; L[7 * j] = 1;
      for (int i = j + 1; i < 6; i++) {
        double s = S[6 * i + j];
        for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k] * D[k];
        L[6 * i + j] = s / d;
      }
    }
    for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[6 * i + k] * y[k]; y[i] = s; }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * x[k]; x[i] = s; }
    return true;
  }
  void update() {
    Quat dq; V3 dt;
    se3_exp(x, dq, dt);
    SE3Q n;
    n.t = dt + quat_rotate(dq, T.t);
    n.q = quat_normalized(quat_mul(dq, T.q));
    T = n;
  }
  void push() { bk.push_back(T); }
  void pop() { T = bk.back(); bk.pop_back(); }
  void discard_top() { bk.pop_back(); }
  double compute_scale(double lambda) const { double s = 0; for (int j = 0; j < 6; j++) s += x[j] * (lambda * x[j] + b[j]); return s; }
};

static SE3Q se3q_from_f32(const float* T) {
  M3 R = {{T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]}};
  return {quat_normalized(quat_from_R(R)), {T[3], T[7], T[11]}};
}

}  // namespace vo

using namespace vo;

extern "C" {

void vo_poseopt_default_params(vo_poseopt_problem* p) {
  p->info_flow = 0.1f; p->info_prior = 0.3f; p->rp_thres = 0.04f; p->chi2_th = 5.991f;
  p->rounds = 4; p->its = 100;
}

int vo_poseopt_flow2cam(vo_poseopt_problem* p, vo_lm_stats* stats) {
  const int N = p->n;
  Flow2System S;
  S.N = N;
  S.fx = p->fx; S.fy = p->fy; S.cx = p->cx; S.cy = p->cy;
  S.info_f = (double)p->info_flow;   // Matrix2d << 0.1 ... (double literals; the struct carries them as float)
  S.info_p = (double)p->info_prior;
  if (p->info_flow == 0.1f) S.info_f = 0.1;
  if (p->info_prior == 0.3f) S.info_p = 0.3;
  S.delta = (double)sqrtf(p->rp_thres);  // const float deltaMono = sqrt(rp_thres)
  S.flow.resize(2 * (size_t)N); S.flow0.resize(2 * (size_t)N); S.obs.resize(2 * (size_t)N);
  S.Xw.resize(N); S.level.assign(N, 0);
  S.eProj.assign(2 * (size_t)N, 0.0); S.ePrior.assign(2 * (size_t)N, 0.0);
  S.Hll.assign(4 * (size_t)N, 0.0); S.bl.assign(2 * (size_t)N, 0.0); S.Hpl.assign(12 * (size_t)N, 0.0);
  S.x.assign(6 + 2 * (size_t)N, 0.0);
  // Twl from the float32 pose of the last frame: Rwl = Rlw^T, twl = -Rlw^T tlw in float (cv::Mat), then to double
  const float* L = p->Tcw_last;
  float Rwl[9], twl[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Rwl[3 * r + c] = L[4 * c + r];
  for (int r = 0; r < 3; r++) {
    float s = 0.f;  // (-Rlw.t()) * tlw : negate the matrix first, then multiply-accumulate in float
    for (int k = 0; k < 3; k++) s += (-Rwl[3 * r + k]) * L[4 * k + 3];
    twl[r] = s;
  }
  for (int i = 0; i < N; i++) {
    S.obs[2 * i] = p->obs_xy[2 * i]; S.obs[2 * i + 1] = p->obs_xy[2 * i + 1];
    S.flow0[2 * i] = S.flow[2 * i] = p->flow_xy[2 * i];
    S.flow0[2 * i + 1] = S.flow[2 * i + 1] = p->flow_xy[2 * i + 1];
    const double depth = p->depth[i];
    V3 Xc = {(S.obs[2 * i] - S.cx) * depth / S.fx, (S.obs[2 * i + 1] - S.cy) * depth / S.fy, depth};
    M3 R = {{Rwl[0], Rwl[1], Rwl[2], Rwl[3], Rwl[4], Rwl[5], Rwl[6], Rwl[7], Rwl[8]}};
    S.Xw[i] = mul(R, Xc) + V3{twl[0], twl[1], twl[2]};
  }
  const SE3Q init = se3q_from_f32(p->Tcw_init);
  S.T = init;
  int nBad = 0;
  if (N >= 3) {
    const float chi2Mono[4] = {p->rp_thres, p->chi2_th, p->chi2_th, p->chi2_th};
    for (int it = 0; it < p->rounds; it++) {
      S.T = init;
      S.bkT.clear(); S.bkF.clear();
      lm_optimize(S, p->its, -1.0, -1.0, stats ? &stats[it] : nullptr);
      nBad = 0;
      for (int i = 0; i < N; i++) {
        if (S.level[i] != 0) S.proj_error(i, &S.eProj[2 * i]);
        const float chi2 = (float)((S.eProj[2 * i] * S.eProj[2 * i] + S.eProj[2 * i + 1] * S.eProj[2 * i + 1]) * S.info_f);
        const float th = chi2Mono[it < 4 ? it : 3];
        if (chi2 > th) { S.level[i] = 1; nBad++; }
        else S.level[i] = 0;
      }
      if (it == 2) S.robust = false;
      if (2 * N < 5) break;
    }
  }
  if (N < 3) {  // "if(nInitialCorrespondences<3) return 0;" -- nothing is touched
    memcpy(p->Tcw_out, p->Tcw_init, sizeof(float) * 16);
    for (int i = 0; i < N; i++) {
      if (p->flow_out) { p->flow_out[2 * i] = p->flow_xy[2 * i]; p->flow_out[2 * i + 1] = p->flow_xy[2 * i + 1]; }
      if (p->inlier) p->inlier[i] = 1;
    }
    return 0;
  }
  // recover pose: SE3Quat -> homogeneous double -> float
  M3 R = quat_to_R(S.T.q);
  float* o = p->Tcw_out;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) o[4 * r + c] = (float)R.m[3 * r + c];
  }
  o[3] = (float)S.T.t.x; o[7] = (float)S.T.t.y; o[11] = (float)S.T.t.z;
  o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
  for (int i = 0; i < N; i++) {
    if (p->flow_out) { p->flow_out[2 * i] = (float)S.flow[2 * i]; p->flow_out[2 * i + 1] = (float)S.flow[2 * i + 1]; }
    if (p->inlier) p->inlier[i] = (N >= 3) ? (S.level[i] == 0) : 1;
  }
  return (N >= 3) ? N - nBad : 0;
}

void vo_projopt_default_params(vo_projopt_problem* p, int kind) {
  p->kind = kind; p->rp_thres = 0.01f; p->its = kind == 0 ? 100 : 200;
}

int vo_pose_opt_proj(vo_projopt_problem* p, vo_lm_stats* stats) {
  const int N = p->n;
  if (N < 3) {  // "if(nInitialCorrespondences<3) return" -- kind 0 leaves the pose, kind 1 returns identity
    if (p->kind == 0) memcpy(p->T_out, p->T_init, sizeof(float) * 16);
    else { memset(p->T_out, 0, sizeof(float) * 16); p->T_out[0] = p->T_out[5] = p->T_out[10] = p->T_out[15] = 1.f; }
    for (int i = 0; i < N; i++) if (p->inlier) p->inlier[i] = 1;
    p->n_inliers = 0;
    if (stats) { stats->iterations = -1; stats->n_records = 0; stats->total_trials = 0; }
    return 0;
  }
  ProjSystem S;
  S.N = N; S.kind = p->kind;
  S.fx = p->fx; S.fy = p->fy; S.cx = p->cx; S.cy = p->cy;
  memcpy(S.P, p->P, sizeof S.P);
  S.robust = p->kind == 0;
  S.delta = (double)sqrtf(p->rp_thres);
  S.Xw.resize(N); S.obs.resize(2 * (size_t)N); S.err.assign(2 * (size_t)N, 0.0);
  for (int i = 0; i < N; i++) {
    S.Xw[i] = {p->pts3d[3 * i], p->pts3d[3 * i + 1], p->pts3d[3 * i + 2]};
    S.obs[2 * i] = p->obs_xy[2 * i]; S.obs[2 * i + 1] = p->obs_xy[2 * i + 1];
  }
  memset(S.x, 0, sizeof S.x);
  S.T = se3q_from_f32(p->T_init);
  lm_optimize(S, p->its, -1.0, -1.0, stats);
  int nBad = 0;
  for (int i = 0; i < N; i++) {
    const float chi2 = (float)(S.err[2 * i] * S.err[2 * i] + S.err[2 * i + 1] * S.err[2 * i + 1]);   // errors of the last evaluation
    const bool out = chi2 > p->rp_thres;
    if (p->inlier) p->inlier[i] = out ? 0 : 1;
    nBad += out ? 1 : 0;
  }
  M3 R = quat_to_R(S.T.q);
  float* o = p->T_out;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[4 * r + c] = (float)R.m[3 * r + c];
  o[3] = (float)S.T.t.x; o[7] = (float)S.T.t.y; o[11] = (float)S.T.t.z;
  o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
  p->n_inliers = N - nBad;
  return p->n_inliers;
}

}  // extern "C"
