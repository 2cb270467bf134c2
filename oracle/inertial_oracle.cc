// inertial_oracle.cc -- CPU restatement of the inertial-only optimisation (VIO initialisation).  TEST INFRASTRUCTURE ONLY.
//
//   Optimizer::InertialOptimization (gravity direction, scale, biases, velocities; poses fixed)   src/Optimizer.cc:2441-2620
//   EdgeInertialGS (ctor: information from the preintegrated covariance; computeError; linearizeOplus)  src/G2oTypes.cc:357-482
//   EdgePriorAcc / EdgePriorGyro  include/G2oTypes.h:578-624, src/G2oTypes.cc:526-538  (error = prior - estimate with a +I
//                                 Jacobian -- kept as written)
//   VertexVelocity / VertexGyroBias / VertexAccBias (additive), VertexGDir (Rwg <- Rwg Exp(u0,u1,0)), VertexScale (s <- s e^u)
//                                 include/G2oTypes.h:144-284
//   IMU::Preintegrated::GetDeltaRotation / GetDeltaVelocity / GetDeltaPosition / GetDeltaBias    src/ImuTypes.cc:339-368 (float32)
//   ExpSO3 / LogSO3 / RightJacobianSO3 / InverseRightJacobianSO3 (double)                        src/G2oTypes.cc:541-613
// Third-party arithmetic that is un-vendored and therefore "parity unpinned" (tolerance-checked): cv::Mat::inv(DECOMP_SVD) of
// the float32 9x9 covariance and Eigen::SelfAdjointEigenSolver are restated with a cyclic Jacobi eigen-decomposition in double;
// cv::SVDecomp in NormalizeRotation is the orthogonal polar factor (Newton iteration); LinearSolverEigen (SimplicialLDLT) is a
// dense LL^T of the same matrix.  The unknown order is g2o's vertex-id order: velocities, gyro bias, acc bias, gravity
// direction, scale.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "lm_oracle.h"
#include "vido_oracle.h"

namespace {

const double GRAVITY_VALUE = (double)9.79f;  // include/ImuTypes.h:29 (a float constant)

void mul33(const double* a, const double* b, double* o) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(o, t, sizeof t);
}
void mul33t(const double* a, const double* b, double* o) {  // a^T b
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
  memcpy(o, t, sizeof t);
}
void mulv(const double* a, const double* v, double* o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = a[3 * i] * v[0] + a[3 * i + 1] * v[1] + a[3 * i + 2] * v[2];
  memcpy(o, t, sizeof t);
}
void multv(const double* a, const double* v, double* o) {  // a^T v
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = a[i] * v[0] + a[3 + i] * v[1] + a[6 + i] * v[2];
  memcpy(o, t, sizeof t);
}
void inv33(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// IMU::NormalizeRotation on a float32 matrix: polar factor, result rounded to float32
void normalize_rotation_f(const double* R, double* out) {
  double X[9];
  for (int k = 0; k < 9; k++) X[k] = (double)(float)R[k];
  for (int it = 0; it < 8; it++) {
    double Xi[9];
    inv33(X, Xi);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) X[3 * i + j] = 0.5 * (X[3 * i + j] + Xi[3 * j + i]);
  }
  for (int k = 0; k < 9; k++) out[k] = (double)(float)X[k];
}
void skew(const double* v, double* W) {
  W[0] = 0; W[1] = -v[2]; W[2] = v[1]; W[3] = v[2]; W[4] = 0; W[5] = -v[0]; W[6] = -v[1]; W[7] = v[0]; W[8] = 0;
}
// ExpSO3 (double, src/G2oTypes.cc:546-562): Rodrigues, then NormalizeRotation through a float32 cv::Mat
void exp_so3(double x, double y, double z, double* R) {
  const double d2 = x * x + y * y + z * z, d = std::sqrt(d2);
  const double v[3] = {x, y, z};
  double W[9], W2[9], res[9];
  skew(v, W);
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    res[k] = d < 1e-5 ? I + W[k] + 0.5 * W2[k] : I + W[k] * std::sin(d) / d + W2[k] * (1.0 - std::cos(d)) / d2;
  }
  normalize_rotation_f(res, R);
}
// ExpSO3 (float32, src/ImuTypes.cc:38-50): entries rounded to float32, no normalisation
void exp_so3_f(float x, float y, float z, double* R) {
  const float d2 = x * x + y * y + z * z;
  const float d = std::sqrt(d2);
  const double W[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  double W2[9];
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    const double v = d < 1e-4f ? I + W[k] + 0.5 * (double)(float)W2[k]
                               : I + W[k] * std::sin((double)d) / d + (double)(float)W2[k] * (1.0 - std::cos((double)d)) / d2;
    R[k] = (double)(float)v;
  }
}
void log_so3(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (tr - 1.0) * 0.5;
  if (costheta > 1 || costheta < -1) return;
  const double theta = std::acos(costheta), s = std::sin(theta);
  if (std::fabs(s) < 1e-5) return;
  for (int k = 0; k < 3; k++) w[k] = theta * w[k] / s;
}
void inv_right_jacobian(const double* v, double* J) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = std::sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    J[k] = d < 1e-5 ? I : I + W[k] / 2 + W2[k] * (1.0 / d2 - (1.0 + std::cos(d)) / (2.0 * d * std::sin(d)));
  }
}
void right_jacobian(const double* v, double* J) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = std::sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    J[k] = d < 1e-5 ? I : I - W[k] * (1.0 - std::cos(d)) / d2 + W2[k] * (d - std::sin(d)) / (d2 * d);
  }
}

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (row-major): A = V diag(w) V^T
void jacobi_eig(int n, std::vector<double>& A, std::vector<double>& V, std::vector<double>& w) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
    if (off == 0) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq; A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk; A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A[(size_t)i * n + i];
}

// EdgeInertialGS ctor (src/G2oTypes.cc:357-376): Info = C[0:9,0:9]^-1 (float32, SVD), symmetrised, eigenvalues < 1e-12 zeroed
void edge_information(const float* C15, double* Info) {
  std::vector<double> A(81), V, w;
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) A[9 * r + c] = 0.5 * ((double)C15[15 * r + c] + (double)C15[15 * c + r]);
  jacobi_eig(9, A, V, w);
  double sum = 0;
  for (int i = 0; i < 9; i++) sum += std::fabs(w[i]);
  const double thr = (double)FLT_EPSILON * 2 * sum;   // cv::SVD::backSubst threshold for float32 input
  float invf[81];
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++)
        if (std::fabs(w[k]) > thr) s += V[9 * r + k] * V[9 * c + k] / w[k];
      invf[9 * r + c] = (float)s;
    }
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) A[9 * r + c] = ((double)invf[9 * r + c] + (double)invf[9 * c + r]) / 2;
  jacobi_eig(9, A, V, w);
  for (int i = 0; i < 9; i++)
    if (w[i] < 1e-12) w[i] = 0;
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++) s += V[9 * r + k] * w[k] * V[9 * c + k];
      Info[9 * r + c] = s;
    }
}

struct InertialSys {
  int N = 0, dim = 0, mode = 0;
  std::vector<double> Rwb, twb, V;     // [N][9], [N][3], [N][3]
  double bg[3], ba[3], Rwg[9], s;
  const vo_imu_preint* pre = nullptr;  // [N-1]
  const float* blin = nullptr;         // [N-1][6]
  std::vector<double> Info;            // [N-1][81]
  double priorG, priorA;
  std::vector<double> err;             // [N-1][9]
  std::vector<double> H, b, x;
  struct Bk { std::vector<double> V; double bg[3], ba[3], Rwg[9], s; };
  std::vector<Bk> stack;

  int num_vertices() const { return N + 4; }

  // delta getters with the updated bias (float32 arithmetic of src/ImuTypes.cc:339-368)
  void deltas(int e, double* dR, double* dV, double* dP, double* dbg_d) const {
    const vo_imu_preint& p = pre[e];
    const float* bl = blin + 6 * (size_t)e;
    float dbg[3], dba[3];
    for (int k = 0; k < 3; k++) { dba[k] = (float)ba[k] - bl[k]; dbg[k] = (float)bg[k] - bl[3 + k]; }
    if (dbg_d) for (int k = 0; k < 3; k++) dbg_d[k] = dbg[k];
    float rj[3];
    for (int i = 0; i < 3; i++) rj[i] = (float)((double)p.JRg[3 * i] * dbg[0] + (double)p.JRg[3 * i + 1] * dbg[1] + (double)p.JRg[3 * i + 2] * dbg[2]);
    double E[9], Rd[9], M[9];
    exp_so3_f(rj[0], rj[1], rj[2], E);
    for (int k = 0; k < 9; k++) Rd[k] = p.dR[k];
    mul33(Rd, E, M);
    normalize_rotation_f(M, dR);
    for (int i = 0; i < 3; i++) {
      const float t1 = (float)((double)p.JVg[3 * i] * dbg[0] + (double)p.JVg[3 * i + 1] * dbg[1] + (double)p.JVg[3 * i + 2] * dbg[2]);
      const float t2 = (float)((double)p.JVa[3 * i] * dba[0] + (double)p.JVa[3 * i + 1] * dba[1] + (double)p.JVa[3 * i + 2] * dba[2]);
      dV[i] = (double)(float)((float)(p.dV[i] + t1) + t2);
      const float u1 = (float)((double)p.JPg[3 * i] * dbg[0] + (double)p.JPg[3 * i + 1] * dbg[1] + (double)p.JPg[3 * i + 2] * dbg[2]);
      const float u2 = (float)((double)p.JPa[3 * i] * dba[0] + (double)p.JPa[3 * i + 1] * dba[1] + (double)p.JPa[3 * i + 2] * dba[2]);
      dP[i] = (double)(float)((float)(p.dP[i] + u1) + u2);
    }
  }
  void edge(int e, double* er9, double* J /* 9 x 15 or null */) const {
    const double* R1 = &Rwb[9 * (size_t)e];
    const double* R2 = &Rwb[9 * (size_t)(e + 1)];
    const double* t1 = &twb[3 * (size_t)e];
    const double* t2 = &twb[3 * (size_t)(e + 1)];
    const double* V1 = &V[3 * (size_t)e];
    const double* V2 = &V[3 * (size_t)(e + 1)];
    const double dt = (double)pre[e].dT;
    const double gI[3] = {0, 0, -GRAVITY_VALUE};
    double g[3], dR[9], dV[3], dP[3], dbg[3];
    mulv(Rwg, gI, g);
    deltas(e, dR, dV, dP, dbg);
    double R12[9], eR[9], er[3];
    mul33t(R1, R2, R12);
    mul33t(dR, R12, eR);
    log_so3(eR, er);
    double a[3], c[3], ra[3], rc[3];
    for (int k = 0; k < 3; k++) {
      a[k] = s * (V2[k] - V1[k]) - g[k] * dt;
      c[k] = s * (t2[k] - t1[k] - V1[k] * dt) - g[k] * dt * dt / 2;
    }
    multv(R1, a, ra);
    multv(R1, c, rc);
    for (int k = 0; k < 3; k++) { er9[k] = er[k]; er9[3 + k] = ra[k] - dV[k]; er9[6 + k] = rc[k] - dP[k]; }
    if (!J) return;
    memset(J, 0, sizeof(double) * 9 * 15);
    auto put = [&](int r0, int c0, const double* M3, double f) {
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) J[15 * (r0 + r) + c0 + q] = f * M3[3 * r + q];
    };
    double Rbw1[9];
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) Rbw1[3 * r + q] = R1[3 * q + r];
    double invJr[9];
    inv_right_jacobian(er, invJr);
    // velocity 1 (cols 0..2)
    put(3, 0, Rbw1, -s);
    put(6, 0, Rbw1, -s * dt);
    // gyro bias (cols 3..5): -invJr eR^T Jr(JRg dbg) JRg ; -JVg ; -JPg
    {
      double JRg[9], JVg[9], JPg[9], v[3], Jr[9], T1[9], T2[9], T3[9];
      for (int k = 0; k < 9; k++) { JRg[k] = pre[e].JRg[k]; JVg[k] = pre[e].JVg[k]; JPg[k] = pre[e].JPg[k]; }
      mulv(JRg, dbg, v);
      right_jacobian(v, Jr);
      mul33t(eR, Jr, T1);
      mul33(T1, JRg, T2);
      mul33(invJr, T2, T3);
      put(0, 3, T3, -1.0);
      put(3, 3, JVg, -1.0);
      put(6, 3, JPg, -1.0);
    }
    // acc bias (cols 6..8)
    {
      double JVa[9], JPa[9];
      for (int k = 0; k < 9; k++) { JVa[k] = pre[e].JVa[k]; JPa[k] = pre[e].JPa[k]; }
      put(3, 6, JVa, -1.0);
      put(6, 6, JPa, -1.0);
    }
    // velocity 2 (cols 9..11)
    put(3, 9, Rbw1, s);
    // gravity direction (cols 12..13): dGdTheta = Rwg * Gm, Gm(0,1) = -G, Gm(1,0) = G
    {
      double dG[6];  // 3 x 2
      for (int r = 0; r < 3; r++) { dG[2 * r] = Rwg[3 * r + 1] * GRAVITY_VALUE; dG[2 * r + 1] = -Rwg[3 * r] * GRAVITY_VALUE; }
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 2; q++) {
          double v = 0;
          for (int k = 0; k < 3; k++) v += Rbw1[3 * r + k] * dG[2 * k + q];
          J[15 * (3 + r) + 12 + q] = -v * dt;
          J[15 * (6 + r) + 12 + q] = -0.5 * v * dt * dt;
        }
    }
    // scale (col 14)
    {
      double dv[3], dp[3], o1[3], o2[3];
      for (int k = 0; k < 3; k++) { dv[k] = V2[k] - V1[k]; dp[k] = t2[k] - t1[k] - V1[k] * dt; }
      multv(R1, dv, o1);
      multv(R1, dp, o2);
      for (int r = 0; r < 3; r++) { J[15 * (3 + r) + 14] = o1[r]; J[15 * (6 + r) + 14] = o2[r]; }
    }
  }
  int gcol(int e, int lc) const {  // local column -> global index
    if (lc < 3) return 3 * e + lc;
    if (lc < 9) return 3 * N + (lc - 3);
    if (lc < 12) return 3 * (e + 1) + (lc - 9);
    return 3 * N + 6 + (lc - 12);
  }
  void compute_errors() {
    for (int e = 0; e + 1 < N; e++) edge(e, &err[9 * (size_t)e], nullptr);
  }
  double robust_chi2() const {
    double chi = 0;
    for (int e = 0; e + 1 < N; e++) {
      const double* r = &err[9 * (size_t)e];
      const double* I9 = &Info[81 * (size_t)e];
      for (int i = 0; i < 9; i++)
        for (int j = 0; j < 9; j++) chi += r[i] * I9[9 * i + j] * r[j];
    }
    if (mode == 0)
      for (int k = 0; k < 3; k++) chi += priorA * ba[k] * ba[k] + priorG * bg[k] * bg[k];
    return chi;
  }
  void build_system() {
    std::fill(H.begin(), H.end(), 0.0);
    std::fill(b.begin(), b.end(), 0.0);
    std::vector<double> J(9 * 15), OJ(9 * 15);
    for (int e = 0; e + 1 < N; e++) {
      double r[9], Or[9];
      edge(e, r, J.data());
      const double* I9 = &Info[81 * (size_t)e];
      for (int i = 0; i < 9; i++) {
        double sr = 0;
        for (int k = 0; k < 9; k++) sr += I9[9 * i + k] * r[k];
        Or[i] = sr;
        for (int c = 0; c < 15; c++) {
          double sj = 0;
          for (int k = 0; k < 9; k++) sj += I9[9 * i + k] * J[15 * k + c];
          OJ[15 * i + c] = sj;
        }
      }
      for (int c1 = 0; c1 < 15; c1++) {
        const int g1 = gcol(e, c1);
        double sb = 0;
        for (int k = 0; k < 9; k++) sb += J[15 * k + c1] * Or[k];
        b[g1] -= sb;
        for (int c2 = 0; c2 < 15; c2++) {
          double sh = 0;
          for (int k = 0; k < 9; k++) sh += J[15 * k + c1] * OJ[15 * k + c2];
          H[(size_t)g1 * dim + gcol(e, c2)] += sh;
        }
      }
    }
    // priors: error = 0 - estimate, Jacobian +I (as written in the reference)
    for (int k = 0; k < 3 && mode == 0; k++) {
      H[(size_t)(3 * N + k) * dim + 3 * N + k] += priorG;
      b[3 * N + k] -= priorG * (0.0 - bg[k]);
      H[(size_t)(3 * N + 3 + k) * dim + 3 * N + 3 + k] += priorA;
      b[3 * N + 3 + k] -= priorA * (0.0 - ba[k]);
    }
  }
  double max_diag() const {
    double m = 0;
    for (int i = (mode == 1 ? 3 * N + 6 : 0); i < dim; i++) m = std::max(m, std::fabs(H[(size_t)i * dim + i]));
    return m;
  }
  bool solve3(double lambda) {   // mode 1: only gravity direction (2) and scale (1) are free
    const int o = 3 * N + 6;
    double A[9], L[9] = {0}, y[3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A[3 * r + c] = H[(size_t)(o + r) * dim + o + c] + (r == c ? lambda : 0.0);
    for (int j = 0; j < 3; j++) {
      double d = A[3 * j + j];
      for (int k = 0; k < j; k++) d -= L[3 * j + k] * L[3 * j + k];
      if (d <= 0) return false;
      L[3 * j + j] = std::sqrt(d);
      for (int i = j + 1; i < 3; i++) {
        double s2 = A[3 * i + j];
        for (int k = 0; k < j; k++) s2 -= L[3 * i + k] * L[3 * j + k];
        L[3 * i + j] = s2 / L[3 * j + j];
      }
    }
    for (int i = 0; i < 3; i++) { double s2 = b[o + i]; for (int k = 0; k < i; k++) s2 -= L[3 * i + k] * y[k]; y[i] = s2 / L[3 * i + i]; }
    for (int i = 2; i >= 0; i--) { double s2 = y[i]; for (int k = i + 1; k < 3; k++) s2 -= L[3 * k + i] * x[o + k]; x[o + i] = s2 / L[3 * i + i]; }
    return true;
  }
  bool solve(double lambda) {
    if (mode == 1) return solve3(lambda);
    std::vector<double> L(H);
    for (int i = 0; i < dim; i++) L[(size_t)i * dim + i] += lambda;
    for (int j = 0; j < dim; j++) {
      double d = L[(size_t)j * dim + j];
      for (int k = 0; k < j; k++) d -= L[(size_t)j * dim + k] * L[(size_t)j * dim + k];
      if (d <= 0) return false;  // x keeps the previous solution (linear_solver_eigen.h:94-125)
      const double ljj = std::sqrt(d);
      L[(size_t)j * dim + j] = ljj;
      for (int i = j + 1; i < dim; i++) {
        double s2 = L[(size_t)i * dim + j];
        for (int k = 0; k < j; k++) s2 -= L[(size_t)i * dim + k] * L[(size_t)j * dim + k];
        L[(size_t)i * dim + j] = s2 / ljj;
      }
    }
    for (int i = 0; i < dim; i++) {
      double s2 = b[i];
      for (int k = 0; k < i; k++) s2 -= L[(size_t)i * dim + k] * x[k];
      x[i] = s2 / L[(size_t)i * dim + i];
    }
    for (int i = dim - 1; i >= 0; i--) {
      double s2 = x[i];
      for (int k = i + 1; k < dim; k++) s2 -= L[(size_t)k * dim + i] * x[k];
      x[i] = s2 / L[(size_t)i * dim + i];
    }
    return true;
  }
  void update() {
    for (int i = 0; i < 3 * N; i++) V[i] += x[i];
    for (int k = 0; k < 3; k++) { bg[k] += x[3 * N + k]; ba[k] += x[3 * N + 3 + k]; }
    double E[9], R[9];
    exp_so3(x[3 * N + 6], x[3 * N + 7], 0.0, E);
    mul33(Rwg, E, R);
    memcpy(Rwg, R, sizeof R);
    s = s * std::exp(x[3 * N + 8]);
  }
  void push() { Bk k; k.V = V; memcpy(k.bg, bg, sizeof bg); memcpy(k.ba, ba, sizeof ba); memcpy(k.Rwg, Rwg, sizeof Rwg); k.s = s; stack.push_back(k); }
  void pop() { Bk& k = stack.back(); V = k.V; memcpy(bg, k.bg, sizeof bg); memcpy(ba, k.ba, sizeof ba); memcpy(Rwg, k.Rwg, sizeof Rwg); s = k.s; stack.pop_back(); }
  void discard_top() { stack.pop_back(); }
  double compute_scale(double lambda) const {
    double sc = 0;
    for (int j = (mode == 1 ? 3 * N + 6 : 0); j < dim; j++) sc += x[j] * (lambda * x[j] + b[j]);
    return sc;
  }
};

}  // namespace

extern "C" {

void vo_inertial_default_params(vo_inertial_problem* p) { p->prior_g = 1e2f; p->prior_a = 1e9f; p->its = 200; p->mode = 0; }

void vo_inertial_edge_information(const float* C15, double* info81) { edge_information(C15, info81); }

int vo_inertial_optimization(vo_inertial_problem* p, vo_lm_stats* stats) {
  InertialSys S;
  S.N = p->n_frames;
  S.mode = p->mode;
  if (S.N < 2) { if (stats) { stats->iterations = -1; stats->n_records = 0; stats->total_trials = 0; } return -1; }
  S.dim = 3 * S.N + 9;
  S.Rwb.resize(9 * (size_t)S.N); S.twb.resize(3 * (size_t)S.N); S.V.resize(3 * (size_t)S.N);
  for (size_t i = 0; i < S.Rwb.size(); i++) S.Rwb[i] = p->Rwb[i];
  for (size_t i = 0; i < S.twb.size(); i++) { S.twb[i] = p->twb[i]; S.V[i] = p->velocity[i]; }
  for (int k = 0; k < 3; k++) { S.bg[k] = p->bg[k]; S.ba[k] = p->ba[k]; }
  for (int k = 0; k < 9; k++) S.Rwg[k] = p->Rwg[k];
  S.s = p->scale;
  S.pre = p->preint; S.blin = p->bias_lin;
  S.priorG = (double)p->prior_g; S.priorA = (double)p->prior_a;
  S.Info.resize(81 * (size_t)(S.N - 1));
  for (int e = 0; e + 1 < S.N; e++) edge_information(p->preint[e].C, &S.Info[81 * (size_t)e]);
  S.err.assign(9 * (size_t)(S.N - 1), 0.0);
  S.H.assign((size_t)S.dim * S.dim, 0.0); S.b.assign(S.dim, 0.0); S.x.assign(S.dim, 0.0);
  const int its = vo::lm_optimize(S, p->its, -1.0, (p->mode == 0 && p->prior_g != 0.f) ? 1e3 : -1.0, stats);
  for (size_t i = 0; i < S.V.size(); i++) p->velocity[i] = (float)S.V[i];
  for (int k = 0; k < 3; k++) { p->bg[k] = S.bg[k]; p->ba[k] = S.ba[k]; }
  for (int k = 0; k < 9; k++) p->Rwg[k] = S.Rwg[k];
  p->scale = S.s;
  return its;
}

// IMU::Preintegrated::GetUpdatedDeltaRotation / Velocity / Position (src/ImuTypes.cc:370-386) for db = (dbg, dba), and
// IMU::ExpSO3(float) / NormalizeRotation for the tracker's VIO glue
void vo_imu_updated_deltas(const vo_imu_preint* p, const float* dbg, const float* dba, float* dR, float* dV, float* dP) {
  float rj[3];
  for (int i = 0; i < 3; i++) rj[i] = (float)((double)p->JRg[3 * i] * dbg[0] + (double)p->JRg[3 * i + 1] * dbg[1] + (double)p->JRg[3 * i + 2] * dbg[2]);
  double E[9], Rd[9], M[9], Rn[9];
  exp_so3_f(rj[0], rj[1], rj[2], E);
  for (int k = 0; k < 9; k++) Rd[k] = p->dR[k];
  mul33(Rd, E, M);
  normalize_rotation_f(M, Rn);
  for (int k = 0; k < 9; k++) dR[k] = (float)Rn[k];
  for (int i = 0; i < 3; i++) {
    const float t1 = (float)((double)p->JVg[3 * i] * dbg[0] + (double)p->JVg[3 * i + 1] * dbg[1] + (double)p->JVg[3 * i + 2] * dbg[2]);
    const float t2 = (float)((double)p->JVa[3 * i] * dba[0] + (double)p->JVa[3 * i + 1] * dba[1] + (double)p->JVa[3 * i + 2] * dba[2]);
    dV[i] = (float)((float)(p->dV[i] + t1) + t2);
    const float u1 = (float)((double)p->JPg[3 * i] * dbg[0] + (double)p->JPg[3 * i + 1] * dbg[1] + (double)p->JPg[3 * i + 2] * dbg[2]);
    const float u2 = (float)((double)p->JPa[3 * i] * dba[0] + (double)p->JPa[3 * i + 1] * dba[1] + (double)p->JPa[3 * i + 2] * dba[2]);
    dP[i] = (float)((float)(p->dP[i] + u1) + u2);
  }
}
void vo_imu_exp_so3_f(const float* w, float* R) {
  double Rd[9];
  exp_so3_f(w[0], w[1], w[2], Rd);
  for (int k = 0; k < 9; k++) R[k] = (float)Rd[k];
}

}  // extern "C"
