// ba_oracle.cc -- CPU restatement of the sliding-window graph optimisation.  TEST INFRASTRUCTURE ONLY.
//
//   Optimizer::PartialBatchOptimization      src/Optimizer.cc:43-1228 (graph build :220-362, optimize :806, write-back :1056-1142)
//   EdgeSE3 computeError / linearizeOplus    g2o/types/edge_se3.cpp:77-105, isometry3d_gradients.h:192-262 (6-argument overload)
//   EdgeSE3PointXYZ                          g2o/types/edge_se3_pointxyz.cpp:99-140 (offset parameter = identity)
//   VertexSE3::oplusImpl                     g2o/types/vertex_se3.h:105-114
//   constructQuadraticForm (Huber, rho[1] weighting, second-order term dropped)   g2o/core/base_binary_edge.hpp:55-120, base_edge.h:96-102
//   BlockSolverX + LinearSolverCSparse: no vertex is marginalised, so the whole H (poses + points) is
//   Cholesky-factored (block_solver.hpp:354-365, linear_solver_csparse.h:108-141).  CSparse/AMD is un-vendored; the
//   restatement fixes the elimination order "points first, then poses", which is what AMD yields for this arrow
//   pattern, and runs the same up-looking-equivalent LL^T: 3x3 point pivots, L_pl = H_pl L_ll^-T, dense LL^T of the
//   remaining pose block.  A non-positive pivot = Cholesky failure (csparse_extension.cpp:112), reported like the
//   reference (x = b, trial rejected).
#include <algorithm>
#include <cstdio>
#include <vector>

#include "g2o_math.h"
#include "lm_oracle.h"
#include "vido_oracle.h"

namespace vo {

// ------------------------------------------------------------------ edge math
static void skewT2(const V3& v, M3& S) {  // isometry3d_gradients.h:49-55 (factor 2 included)
  double x = 2 * v.x, y = 2 * v.y, z = 2 * v.z;
  S = {{0, -z, y, z, 0, -x, -y, x, 0}};
}

// M (9x3, column c = column-major vec of A * S_c) multiplied by dq_dR (3x9)
static void dq_times(const double dq[3][9], const M3 prod[3], double out[3][3]) {
  for (int c = 0; c < 3; c++) {
    double vec[9];
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) vec[3 * j + i] = prod[c].m[3 * i + j];  // column-major
    for (int a = 0; a < 3; a++) {
      double s = 0;
      for (int k = 0; k < 9; k++) s += dq[a][k] * vec[k];
      out[a][c] = s;
    }
  }
}

void edge_se3(const Iso& Xi, const Iso& Xj, const Iso& Z, double err[6], double Ji[6][6], double Jj[6][6]) {
  const Iso A = inverse(Z);
  const Iso B = mul(inverse(Xi), Xj);
  const Iso E = mul(A, B);
  V3 cq = compact_quat(E.R);
  err[0] = E.t.x; err[1] = E.t.y; err[2] = E.t.z; err[3] = cq.x; err[4] = cq.y; err[5] = cq.z;
  if (!Ji) return;
  double dq[3][9];
  dq_dR(E.R, dq);
  memset(Ji, 0, sizeof(double) * 36);
  memset(Jj, 0, sizeof(double) * 36);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Ji[i][j] = -A.R.m[3 * i + j];  // dte/dti = -Ra
      Jj[i][j] = E.R.m[3 * i + j];   // dte/dtj = Re
    }
  M3 S;
  skewT2(B.t, S);
  M3 RaS = mul(A.R, S);  // dte/dqi
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Ji[i][3 + j] = RaS.m[3 * i + j];
  {  // dre/dqi: skewT(Sxt,Syt,Szt, Rb)
    const double* r = B.R.m;
    const double r11 = 2 * r[0], r12 = 2 * r[1], r13 = 2 * r[2], r21 = 2 * r[3], r22 = 2 * r[4], r23 = 2 * r[5],
                 r31 = 2 * r[6], r32 = 2 * r[7], r33 = 2 * r[8];
    M3 Sx = {{0, 0, 0, r31, r32, r33, -r21, -r22, -r23}};
    M3 Sy = {{-r31, -r32, -r33, 0, 0, 0, r11, r12, r13}};
    M3 Sz = {{r21, r22, r23, -r11, -r12, -r13, 0, 0, 0}};
    M3 prod[3] = {mul(A.R, Sx), mul(A.R, Sy), mul(A.R, Sz)};
    double o[3][3];
    dq_times(dq, prod, o);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Ji[3 + i][3 + j] = o[i][j];
  }
  {  // dre/dqj: skew(Sx,Sy,Sz, Identity)
    M3 Sx = {{0, 0, 0, 0, 0, -2, 0, 2, 0}};
    M3 Sy = {{0, 0, 2, 0, 0, 0, -2, 0, 0}};
    M3 Sz = {{0, -2, 0, 2, 0, 0, 0, 0, 0}};
    M3 prod[3] = {mul(E.R, Sx), mul(E.R, Sy), mul(E.R, Sz)};
    double o[3][3];
    dq_times(dq, prod, o);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Jj[3 + i][3 + j] = o[i][j];
  }
}

void edge_se3_pointxyz(const Iso& X, const V3& p, const V3& z, double err[3], double Ji[3][6], double Jj[3][3]) {
  const Iso w2n = inverse(X);
  const V3 zc = apply(w2n, p);
  err[0] = zc.x - z.x; err[1] = zc.y - z.y; err[2] = zc.z - z.z;
  if (!Ji) return;
  memset(Ji, 0, sizeof(double) * 18);
  Ji[0][0] = Ji[1][1] = Ji[2][2] = -1;
  Ji[0][4] = -2 * zc.z; Ji[0][5] = 2 * zc.y;
  Ji[1][3] = 2 * zc.z;  Ji[1][5] = -2 * zc.x;
  Ji[2][3] = -2 * zc.y; Ji[2][4] = 2 * zc.x;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Jj[i][j] = w2n.R.m[3 * i + j];
}

Iso se3_oplus(const Iso& X, const double u[6]) {
  Iso inc;
  inc.R = from_compact_quat({u[3], u[4], u[5]});
  inc.t = {u[0], u[1], u[2]};
  return mul(X, inc);
}

// ------------------------------------------------------------------ the window system
struct BASystem {
  int W = 0, P = 0, M = 0;
  std::vector<Iso> X, Z;
  std::vector<V3> pts, meas;
  std::vector<int> op, ol;
  double info_cam, info_3d, d_cam, d_3d;
  std::vector<double> eSE3, ePt;            // errors
  std::vector<double> Hpp, Hll, Hpl, b, x;  // Hpp dense (6W)^2 row-major, Hll P*9, Hpl M*18 (6x3), b/x poses then points
  std::vector<int> pt_start, pt_obs;        // point -> observation list
  std::vector<std::vector<Iso>> bkX;
  std::vector<std::vector<V3>> bkP;

  int num_vertices() const { return W + P; }
  int dim() const { return 6 * W + 3 * P; }

  void compute_errors() {
    for (int i = 0; i + 1 < W; i++) edge_se3(X[i], X[i + 1], Z[i], &eSE3[6 * i], nullptr, nullptr);
    for (int o = 0; o < M; o++) edge_se3_pointxyz(X[op[o]], pts[ol[o]], meas[o], &ePt[3 * o], nullptr, nullptr);
  }
  double robust_chi2() const {
    double chi = 0, rho[3];
    for (int i = 0; i + 1 < W; i++) {
      double e = 0;
      for (int k = 0; k < 6; k++) e += eSE3[6 * i + k] * eSE3[6 * i + k];
      huber(e * info_cam, d_cam, rho);
      chi += rho[0];
    }
    for (int o = 0; o < M; o++) {
      double e = ePt[3 * o] * ePt[3 * o] + ePt[3 * o + 1] * ePt[3 * o + 1] + ePt[3 * o + 2] * ePt[3 * o + 2];
      huber(e * info_3d, d_3d, rho);
      chi += rho[0];
    }
    return chi;
  }
  void build_system() {
    const int n = 6 * W;
    std::fill(Hpp.begin(), Hpp.end(), 0.0);
    std::fill(Hll.begin(), Hll.end(), 0.0);
    std::fill(Hpl.begin(), Hpl.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    double rho[3];
    for (int i = 0; i + 1 < W; i++) {
      double e[6], Ji[6][6], Jj[6][6];
      edge_se3(X[i], X[i + 1], Z[i], e, Ji, Jj);
      double chi = 0;
      for (int k = 0; k < 6; k++) chi += e[k] * e[k];
      huber(chi * info_cam, d_cam, rho);
      const double w = rho[1] * info_cam;
      double* J[2] = {&Ji[0][0], &Jj[0][0]};
      for (int a = 0; a < 2; a++) {
        for (int r = 0; r < 6; r++) {
          double s = 0;
          for (int k = 0; k < 6; k++) s += J[a][6 * k + r] * e[k];
          b[6 * (i + a) + r] += -w * s;
        }
        for (int c = a; c < 2; c++)
          for (int r = 0; r < 6; r++)
            for (int q = 0; q < 6; q++) {
              double s = 0;
              for (int k = 0; k < 6; k++) s += J[a][6 * k + r] * J[c][6 * k + q];
              Hpp[(size_t)(6 * (i + a) + r) * n + 6 * (i + c) + q] += w * s;
              if (c != a) Hpp[(size_t)(6 * (i + c) + q) * n + 6 * (i + a) + r] += w * s;
            }
      }
    }
    for (int o = 0; o < M; o++) {
      double e[3], Ji[3][6], Jj[3][3];
      const int pi = op[o], li = ol[o];
      edge_se3_pointxyz(X[pi], pts[li], meas[o], e, Ji, Jj);
      const double chi = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
      huber(chi * info_3d, d_3d, rho);
      const double w = rho[1] * info_3d;
      for (int r = 0; r < 6; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Ji[k][r] * e[k];
        b[6 * pi + r] += -w * s;
        for (int q = 0; q < 6; q++) {
          double h = 0;
          for (int k = 0; k < 3; k++) h += Ji[k][r] * Ji[k][q];
          Hpp[(size_t)(6 * pi + r) * n + 6 * pi + q] += w * h;
        }
        for (int q = 0; q < 3; q++) {
          double h = 0;
          for (int k = 0; k < 3; k++) h += Ji[k][r] * Jj[k][q];
          Hpl[18 * (size_t)o + 3 * r + q] += w * h;
        }
      }
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Jj[k][r] * e[k];
        b[n + 3 * li + r] += -w * s;
        for (int q = 0; q < 3; q++) {
          double h = 0;
          for (int k = 0; k < 3; k++) h += Jj[k][r] * Jj[k][q];
          Hll[9 * (size_t)li + 3 * r + q] += w * h;
        }
      }
    }
  }
  double max_diag() const {
    const int n = 6 * W;
    double m = 0;
    for (int i = 0; i < n; i++) m = std::max(m, std::fabs(Hpp[(size_t)i * n + i]));
    for (int l = 0; l < P; l++)
      for (int k = 0; k < 3; k++) m = std::max(m, std::fabs(Hll[9 * (size_t)l + 4 * k]));
    return m;
  }
  bool solve(double lambda) {
    const int n = 6 * W;
    std::vector<double> S(Hpp), bp(b.begin(), b.begin() + n);
    std::vector<double> Lll((size_t)P * 9), cl((size_t)P * 3), Y((size_t)M * 18);
    for (int i = 0; i < n; i++) S[(size_t)i * n + i] += lambda;
    bool ok = true;
    for (int l = 0; l < P && ok; l++) {
      double D[9];
      for (int k = 0; k < 9; k++) D[k] = Hll[9 * (size_t)l + k];
      D[0] += lambda; D[4] += lambda; D[8] += lambda;
      double* L = &Lll[9 * (size_t)l];
      memset(L, 0, sizeof(double) * 9);
      for (int j = 0; j < 3 && ok; j++) {
        double d = D[3 * j + j];
        for (int k = 0; k < j; k++) d -= L[3 * j + k] * L[3 * j + k];
        if (d <= 0) { ok = false; break; }
        L[3 * j + j] = std::sqrt(d);
        for (int i = j + 1; i < 3; i++) {
          double s = D[3 * i + j];
          for (int k = 0; k < j; k++) s -= L[3 * i + k] * L[3 * j + k];
          L[3 * i + j] = s / L[3 * j + j];
        }
      }
      if (!ok) break;
      // c_l = L^-1 b_l
      double* c = &cl[3 * (size_t)l];
      const double* bl = &b[n + 3 * (size_t)l];
      c[0] = bl[0] / L[0];
      c[1] = (bl[1] - L[3] * c[0]) / L[4];
      c[2] = (bl[2] - L[6] * c[0] - L[7] * c[1]) / L[8];
      // Y_o = Hpl_o L^-T  (6x3)
      for (int k = pt_start[l]; k < pt_start[l + 1]; k++) {
        const int o = pt_obs[k];
        const double* Hb = &Hpl[18 * (size_t)o];
        double* y = &Y[18 * (size_t)o];
        for (int r = 0; r < 6; r++) {
          y[3 * r + 0] = Hb[3 * r] / L[0];
          y[3 * r + 1] = (Hb[3 * r + 1] - y[3 * r] * L[3]) / L[4];
          y[3 * r + 2] = (Hb[3 * r + 2] - y[3 * r] * L[6] - y[3 * r + 1] * L[7]) / L[8];
        }
      }
      for (int k1 = pt_start[l]; k1 < pt_start[l + 1]; k1++) {
        const int o1 = pt_obs[k1], p1 = op[o1];
        const double* y1 = &Y[18 * (size_t)o1];
        for (int r = 0; r < 6; r++) bp[6 * p1 + r] -= y1[3 * r] * c[0] + y1[3 * r + 1] * c[1] + y1[3 * r + 2] * c[2];
        for (int k2 = pt_start[l]; k2 < pt_start[l + 1]; k2++) {
          const int o2 = pt_obs[k2], p2 = op[o2];
          const double* y2 = &Y[18 * (size_t)o2];
          for (int r = 0; r < 6; r++)
            for (int q = 0; q < 6; q++)
              S[(size_t)(6 * p1 + r) * n + 6 * p2 + q] -= y1[3 * r] * y2[3 * q] + y1[3 * r + 1] * y2[3 * q + 1] + y1[3 * r + 2] * y2[3 * q + 2];
        }
      }
    }
    if (ok) {  // dense LL^T of the pose block
      for (int j = 0; j < n && ok; j++) {
        double d = S[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= S[(size_t)j * n + k] * S[(size_t)j * n + k];
        if (d <= 0) { ok = false; break; }
        const double ljj = std::sqrt(d);
        S[(size_t)j * n + j] = ljj;
        for (int i = j + 1; i < n; i++) {
          double s = S[(size_t)i * n + j];
          for (int k = 0; k < j; k++) s -= S[(size_t)i * n + k] * S[(size_t)j * n + k];
          S[(size_t)i * n + j] = s / ljj;
        }
      }
    }
    if (!ok) {
      x = b;  // LinearSolverCSparse::solve copies b into x before the (failed) factorisation
      return false;
    }
    for (int i = 0; i < n; i++) {
      double s = bp[i];
      for (int k = 0; k < i; k++) s -= S[(size_t)i * n + k] * x[k];
      x[i] = s / S[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = x[i];
      for (int k = i + 1; k < n; k++) s -= S[(size_t)k * n + i] * x[k];
      x[i] = s / S[(size_t)i * n + i];
    }
    for (int l = 0; l < P; l++) {
      double c[3] = {cl[3 * (size_t)l], cl[3 * (size_t)l + 1], cl[3 * (size_t)l + 2]};
      for (int k = pt_start[l]; k < pt_start[l + 1]; k++) {
        const int o = pt_obs[k], p = op[o];
        const double* y = &Y[18 * (size_t)o];
        for (int r = 0; r < 6; r++) {
          c[0] -= y[3 * r] * x[6 * p + r];
          c[1] -= y[3 * r + 1] * x[6 * p + r];
          c[2] -= y[3 * r + 2] * x[6 * p + r];
        }
      }
      const double* L = &Lll[9 * (size_t)l];
      double* xl = &x[n + 3 * (size_t)l];
      xl[2] = c[2] / L[8];
      xl[1] = (c[1] - L[7] * xl[2]) / L[4];
      xl[0] = (c[0] - L[3] * xl[1] - L[6] * xl[2]) / L[0];
    }
    return true;
  }
  void update() {
    for (int i = 0; i < W; i++) X[i] = se3_oplus(X[i], &x[6 * i]);
    for (int l = 0; l < P; l++) {
      const double* u = &x[6 * W + 3 * (size_t)l];
      pts[l] = pts[l] + V3{u[0], u[1], u[2]};
    }
  }
  void push() { bkX.push_back(X); bkP.push_back(pts); }
  void pop() { X = bkX.back(); pts = bkP.back(); bkX.pop_back(); bkP.pop_back(); }
  void discard_top() { bkX.pop_back(); bkP.pop_back(); }
  double compute_scale(double lambda) const {
    double s = 0;
    for (int j = 0; j < dim(); j++) s += x[j] * (lambda * x[j] + b[j]);
    return s;
  }
};

}  // namespace vo

using namespace vo;

extern "C" {

void vo_ba_default_params(vo_ba_problem* p) {
  p->max_iterations = 100;
  p->sigma2_cam = 0.0001f;
  p->sigma2_3d = 16.f;
  p->huber_cam = 0.01f;
  p->huber_3d = 0.01f;
  p->gain_threshold = 1e-3f;
  p->fix_first = 0;
}

int vo_ba_partial(vo_ba_problem* p, vo_lm_stats* stats) {
  BASystem S;
  S.W = p->n_poses; S.P = p->n_points; S.M = p->n_obs;
  S.info_cam = 1.0 / (double)p->sigma2_cam;  // Identity/sigma2 with float sigma2 (src/Optimizer.cc:192,255)
  S.info_3d = 1.0 / (double)p->sigma2_3d;
  S.d_cam = (double)p->huber_cam;
  S.d_3d = (double)p->huber_3d;
  S.X.resize(S.W); S.Z.resize(std::max(S.W - 1, 0)); S.pts.resize(S.P); S.meas.resize(S.M);
  S.op.assign(p->obs_pose, p->obs_pose + S.M);
  S.ol.assign(p->obs_point, p->obs_point + S.M);
  for (int i = 0; i < S.W; i++) S.X[i] = iso_from_f32(p->poses + 16 * i);
  for (int i = 0; i + 1 < S.W; i++) S.Z[i] = iso_from_f32(p->rel_motion + 16 * i);
  for (int l = 0; l < S.P; l++) S.pts[l] = {p->points[3 * l], p->points[3 * l + 1], p->points[3 * l + 2]};
  for (int o = 0; o < S.M; o++) S.meas[o] = {p->obs_xyz[3 * o], p->obs_xyz[3 * o + 1], p->obs_xyz[3 * o + 2]};
  S.eSE3.assign(6 * (size_t)std::max(S.W - 1, 0), 0.0);
  S.ePt.assign(3 * (size_t)S.M, 0.0);
  S.Hpp.assign((size_t)36 * S.W * S.W, 0.0);
  S.Hll.assign((size_t)9 * S.P, 0.0);
  S.Hpl.assign((size_t)18 * S.M, 0.0);
  S.b.assign(S.dim(), 0.0);
  S.x.assign(S.dim(), 0.0);
  S.pt_start.assign(S.P + 1, 0);
  for (int o = 0; o < S.M; o++) S.pt_start[S.ol[o] + 1]++;
  for (int l = 0; l < S.P; l++) S.pt_start[l + 1] += S.pt_start[l];
  S.pt_obs.resize(S.M);
  {
    std::vector<int> fill(S.pt_start.begin(), S.pt_start.end() - 1);
    for (int o = 0; o < S.M; o++) S.pt_obs[fill[S.ol[o]]++] = o;
  }
  int its = lm_optimize(S, p->max_iterations, (double)p->gain_threshold, -1.0, stats);
  // write back (src/Optimizer.cc:1056-1142): float32 poses, relative motions from the float32 poses, points
  for (int i = 0; i < S.W; i++) iso_to_f32(S.X[i], p->poses + 16 * i);
  for (int i = 1; i < S.W; i++) {
    // Converter::toInvMatrix(pose[i-1]) * pose[i] in float32 (cv::Mat arithmetic)
    const float* A = p->poses + 16 * (i - 1);
    const float* Bm = p->poses + 16 * i;
    float Ai[16] = {0};
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Ai[4 * r + c] = A[4 * c + r];
    for (int r = 0; r < 3; r++) Ai[4 * r + 3] = -(Ai[4 * r] * A[3] + Ai[4 * r + 1] * A[7] + Ai[4 * r + 2] * A[11]);
    Ai[15] = 1.f;
    float* out = p->rel_motion + 16 * (i - 1);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        float s = 0.f;
        for (int k = 0; k < 4; k++) s += Ai[4 * r + k] * Bm[4 * k + c];
        out[4 * r + c] = s;
      }
  }
  for (int l = 0; l < S.P; l++) {
    p->points[3 * l] = (float)S.pts[l].x;
    p->points[3 * l + 1] = (float)S.pts[l].y;
    p->points[3 * l + 2] = (float)S.pts[l].z;
  }
  return its;
}

static Iso iso_from_d(const double* X) {
  Iso r;
  for (int i = 0; i < 9; i++) r.R.m[i] = X[i];
  r.t = {X[9], X[10], X[11]};
  return r;
}

void vo_edge_se3(const double* Xi, const double* Xj, const double* Z, double err[6], double Ji[36], double Jj[36]) {
  double a[6][6], b[6][6];
  edge_se3(iso_from_d(Xi), iso_from_d(Xj), iso_from_d(Z), err, a, b);
  memcpy(Ji, a, sizeof a);
  memcpy(Jj, b, sizeof b);
}

void vo_se3_oplus(const double* X, const double* u, double* Xout) {
  Iso r = se3_oplus(iso_from_d(X), u);
  for (int i = 0; i < 9; i++) Xout[i] = r.R.m[i];
  Xout[9] = r.t.x; Xout[10] = r.t.y; Xout[11] = r.t.z;
}

void vo_edge_se3_pointxyz(const double* X, const double* p, const double* z, double err[3], double Ji[18], double Jj[9]) {
  double a[3][6], b[3][3];
  edge_se3_pointxyz(iso_from_d(X), {p[0], p[1], p[2]}, {z[0], z[1], z[2]}, err, a, b);
  memcpy(Ji, a, sizeof a);
  memcpy(Jj, b, sizeof b);
}

}  // extern "C"
