// pnp_oracle.cc -- CPU restatement of Tracking::GetInitModelCam (src/Tracking.cc:1914-2028).  TEST INFRASTRUCTURE ONLY.
// The RANSAC part is a deterministic stand-in for cv::solvePnPRansac (see vido_oracle.h); the motion-model
// comparison and the winner rule follow the reference line by line (float32 arithmetic).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "vido_oracle.h"

namespace {

struct Pose { double R[9], t[3]; };

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// 4 distinct indices of [0,M) for hypothesis `iter`
void sample4(int iter, int M, int* idx) {
  uint64_t c = (uint64_t)iter << 8;
  for (int j = 0; j < 4; j++) {
    while (true) {
      int v = (int)(splitmix(c++) % (uint64_t)M);
      bool dup = false;
      for (int k = 0; k < j; k++) dup |= (idx[k] == v);
      if (!dup) { idx[j] = v; break; }
    }
  }
}

void quat_to_R(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}

// Gauss-Newton PnP: minimise sum |pi(R X + t) - uv|^2 over the listed points, left-multiplicative update
// delta = (omega, upsilon): R <- R(dq) R, t <- R(dq) t + upsilon, dq = normalise(1, omega/2).
bool gn_pnp(const float* pts, const float* uv, const int* ids, int m, double fx, double fy, double cx, double cy, int its,
            Pose& T) {
  for (int it = 0; it < its; it++) {
    double H[36] = {0}, b[6] = {0};
    for (int k = 0; k < m; k++) {
      const int i = ids[k];
      const double X[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
      double p[3];
      for (int r = 0; r < 3; r++) p[r] = T.R[3 * r] * X[0] + T.R[3 * r + 1] * X[1] + T.R[3 * r + 2] * X[2] + T.t[r];
      if (!(p[2] > 1e-9)) return false;
      const double iz = 1.0 / p[2], x = p[0] * iz, y = p[1] * iz;
      const double r0 = fx * x + cx - uv[2 * i], r1 = fy * y + cy - uv[2 * i + 1];
      // d pi / d p
      const double a00 = fx * iz, a02 = -fx * x * iz, a11 = fy * iz, a12 = -fy * y * iz;
      // d p / d (omega, upsilon) = [-[p]x | I]
      double J0[6], J1[6];
      J0[0] = a02 * p[1];                 J0[1] = a00 * p[2] - a02 * p[0];  J0[2] = -a00 * p[1];
      J0[3] = a00; J0[4] = 0; J0[5] = a02;
      J1[0] = -a11 * p[2] + a12 * p[1];   J1[1] = -a12 * p[0];              J1[2] = a11 * p[0];
      J1[3] = 0; J1[4] = a11; J1[5] = a12;
      for (int r = 0; r < 6; r++) {
        b[r] -= J0[r] * r0 + J1[r] * r1;
        for (int c = 0; c < 6; c++) H[6 * r + c] += J0[r] * J0[c] + J1[r] * J1[c];
      }
    }
    // LDL^T
    double L[36] = {0}, D[6], y[6], d[6];
    for (int j = 0; j < 6; j++) {
      double v = H[7 * j];
      for (int k = 0; k < j; k++) v -= L[6 * j + k] * L[6 * j + k] * D[k];
      if (!(v > 1e-12)) return false;
      D[j] = v;
      L[7 * j] = 1;
      for (int i = j + 1; i < 6; i++) {
        double s = H[6 * i + j];
        for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k] * D[k];
        L[6 * i + j] = s / v;
      }
    }
    for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[6 * i + k] * y[k]; y[i] = s; }
    for (int i = 0; i < 6; i++) y[i] /= D[i];
    for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * d[k]; d[i] = s; }
    double q[4] = {1.0, 0.5 * d[0], 0.5 * d[1], 0.5 * d[2]};
    const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; k++) q[k] /= nq;
    double dR[9], Rn[9], tn[3];
    quat_to_R(q, dR);
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) Rn[3 * r + c] = dR[3 * r] * T.R[c] + dR[3 * r + 1] * T.R[3 + c] + dR[3 * r + 2] * T.R[6 + c];
      tn[r] = dR[3 * r] * T.t[0] + dR[3 * r + 1] * T.t[1] + dR[3 * r + 2] * T.t[2] + d[3 + r];
    }
    memcpy(T.R, Rn, sizeof Rn);
    memcpy(T.t, tn, sizeof tn);
  }
  return true;
}

int count_inliers(const Pose& T, const float* pts, const float* uv, const int* good, int M, double fx, double fy, double cx,
                  double cy, double thr, std::vector<int>* out) {
  int c = 0;
  if (out) out->clear();
  for (int k = 0; k < M; k++) {
    const int i = good[k];
    const double X[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    double p[3];
    for (int r = 0; r < 3; r++) p[r] = T.R[3 * r] * X[0] + T.R[3 * r + 1] * X[1] + T.R[3 * r + 2] * X[2] + T.t[r];
    if (!(p[2] > 1e-9)) continue;
    const double du = fx * p[0] / p[2] + cx - uv[2 * i], dv = fy * p[1] / p[2] + cy - uv[2 * i + 1];
    if (std::sqrt(du * du + dv * dv) < thr) { c++; if (out) out->push_back(k); }
  }
  return c;
}

// cv::RANSACUpdateNumIters
int update_num_iters(double p, double ep, int modelPoints, int maxIters) {
  p = std::fmax(p, 0.); p = std::fmin(p, 1.);
  ep = std::fmax(ep, 0.); ep = std::fmin(ep, 1.);
  double num = std::fmax(1. - p, 2.220446049250313e-16);
  double denom = 1. - std::pow(1. - ep, modelPoints);
  if (denom < 2.2250738585072014e-308) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= maxIters * (-denom) ? maxIters : (int)std::lrint(num / denom);
}

}  // namespace

extern "C" {

void vo_pnp_default_params(vo_pnp_problem* p) { p->iters = 500; p->reproj_err = 0.4f; p->confidence = 0.98f; }

int vo_init_model_cam(vo_pnp_problem* p) {
  const int N = p->n;
  const double fx = p->fx, fy = p->fy, cx = p->cx, cy = p->cy;
  std::vector<int> good;
  for (int i = 0; i < N; i++)
    if (!p->valid || p->valid[i]) good.push_back(i);
  const int M = (int)good.size();
  // ---- RANSAC
  Pose T0;
  const float* Tm = p->Tcw_motion;
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T0.R[3 * r + c] = Tm[4 * r + c]; T0.t[r] = Tm[4 * r + 3]; }
  Pose best = T0;
  std::vector<int> best_in;  // positions into `good`
  int best_cnt = 0;
  if (M >= 4) {
    int niters = p->iters;
    for (int iter = 0; iter < niters; iter++) {
      int s[4], ids[4];
      sample4(iter, M, s);
      for (int k = 0; k < 4; k++) ids[k] = good[s[k]];
      Pose T = T0;
      if (!gn_pnp(p->pts3d, p->cur_xy, ids, 4, fx, fy, cx, cy, 6, T)) continue;
      const int cnt = count_inliers(T, p->pts3d, p->cur_xy, good.data(), M, fx, fy, cx, cy, (double)p->reproj_err, nullptr);
      if (cnt > std::max(best_cnt, 3)) {
        best_cnt = cnt;
        best = T;
        niters = update_num_iters((double)p->confidence, (double)(M - cnt) / M, 4, niters);
      }
    }
    if (best_cnt > 0) {
      count_inliers(best, p->pts3d, p->cur_xy, good.data(), M, fx, fy, cx, cy, (double)p->reproj_err, &best_in);
      std::vector<int> ids(best_in.size());
      for (size_t k = 0; k < best_in.size(); k++) ids[k] = good[best_in[k]];
      Pose T = best;
      if (gn_pnp(p->pts3d, p->cur_xy, ids.data(), (int)ids.size(), fx, fy, cx, cy, 10, T)) best = T;  // refit on the consensus set
    }
  }
  p->ransac_inliers = (int)best_in.size();
  // ---- constant-velocity model, float32 like the reference (src/Tracking.cc:1980-2001)
  std::vector<int> mm;
  for (int i = 0; i < N; i++) {
    const float X[3] = {p->pts3d[3 * i], p->pts3d[3 * i + 1], p->pts3d[3 * i + 2]};
    float pc[3];
    for (int r = 0; r < 3; r++) pc[r] = Tm[4 * r] * X[0] + Tm[4 * r + 1] * X[1] + Tm[4 * r + 2] * X[2] + Tm[4 * r + 3];
    const float invz = (float)(1.0 / pc[2]);
    const float u = p->fx * pc[0] * invz + p->cx, v = p->fy * pc[1] * invz + p->cy;
    const float du = p->cur_xy[2 * i] - u, dv = p->cur_xy[2 * i + 1] - v;
    const float rpe = std::sqrt(du * du + dv * dv);
    if (rpe < p->reproj_err) mm.push_back(i);
  }
  p->mm_inliers = (int)mm.size();
  if (p->no_motion_model || (int)best_in.size() > (int)mm.size()) {  // GetInitModelObj without PreObjID (:2143-2151)
    p->winner = 0;
    float* o = p->Tcw_out;
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) o[4 * r + c] = (float)best.R[3 * r + c]; o[4 * r + 3] = (float)best.t[r]; }
    o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
    p->n_inliers = (int)best_in.size();
    // the reference indexes MatchId with the position inside the *good* arrays (:2008-2010); kept as is
    for (size_t k = 0; k < best_in.size(); k++) p->inlier_ids[k] = best_in[k];
  } else {
    p->winner = 1;
    memcpy(p->Tcw_out, Tm, sizeof(float) * 16);
    p->n_inliers = (int)mm.size();
    for (size_t k = 0; k < mm.size(); k++) p->inlier_ids[k] = mm[k];
  }
  return p->n_inliers;
}

}  // extern "C"
