// ba_math.h -- FP64 pose / edge math of the graph-optimisation kernels (host+device inline functions so that the
// CPU test-suite can exercise exactly the code the kernels run).
//
// Replaces, for the sliding-window graph of Optimizer::PartialBatchOptimization (src/Optimizer.cc:43-1228):
//   g2o::VertexSE3::oplusImpl                         g2o/types/vertex_se3.h:105-114
//   g2o::EdgeSE3::computeError / linearizeOplus       g2o/types/edge_se3.cpp:77-105 (+ isometry3d_gradients.h:192-262)
//   g2o::EdgeSE3PointXYZ::computeError/linearizeOplus g2o/types/edge_se3_pointxyz.cpp:99-140
//   g2o::RobustKernelHuber::robustify                 g2o/core/robust_kernel_impl.cpp:78-91
//   Converter::toSE3Quat / toCvSE3 float<->double     src/Converter.cc:29-39,84-100
// The EdgeSE3 Jacobians are written in closed quaternion form (d vec(q_E (x) dq)/d v) instead of the
// reference's dq/dR chain rule; both are the exact derivative of the same error w.r.t. the same increment.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VHD __host__ __device__ __forceinline__
#else
#define VHD inline
#endif

namespace vb {

struct Pose {  // rotation (row-major) and translation of an isometry
  double R[9];
  double t[3];
};

VHD void mat3_mul(const double* a, const double* b, double* o) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
VHD void mat3t_mul(const double* a, const double* b, double* o) {  // a^T b
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
// out = a^-1 * b
VHD void pose_inv_mul(const Pose& a, const Pose& b, Pose& o) {
  mat3t_mul(a.R, b.R, o.R);
  const double d0 = b.t[0] - a.t[0], d1 = b.t[1] - a.t[1], d2 = b.t[2] - a.t[2];
  for (int i = 0; i < 3; i++) o.t[i] = a.R[i] * d0 + a.R[3 + i] * d1 + a.R[6 + i] * d2;
}
VHD void pose_mul(const Pose& a, const Pose& b, Pose& o) {
  mat3_mul(a.R, b.R, o.R);
  for (int i = 0; i < 3; i++) o.t[i] = a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2] + a.t[i];
}

// rotation matrix -> quaternion (w,x,y,z), branch structure of Eigen::Quaterniond(Matrix3d)
VHD void quat_from_R(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t;
    q[2] = (R[2] - R[6]) * t;
    q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > (i == 0 ? R[0] : R[4])) i = 2;
    // the three cases spelled out with constant indices (a run-time index would push R and q into local memory)
    if (i == 0) {  // j = 1, k = 2
      t = sqrt(R[0] - R[4] - R[8] + 1.0);
      q[1] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (R[7] - R[5]) * t;
      q[2] = (R[3] + R[1]) * t;
      q[3] = (R[6] + R[2]) * t;
    } else if (i == 1) {  // j = 2, k = 0
      t = sqrt(R[4] - R[8] - R[0] + 1.0);
      q[2] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (R[2] - R[6]) * t;
      q[3] = (R[7] + R[5]) * t;
      q[1] = (R[1] + R[3]) * t;
    } else {  // j = 0, k = 1
      t = sqrt(R[8] - R[0] - R[4] + 1.0);
      q[3] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (R[3] - R[1]) * t;
      q[1] = (R[2] + R[6]) * t;
      q[2] = (R[5] + R[7]) * t;
    }
  }
}
VHD void R_from_quat(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
VHD void quat_unit(double* q, bool positive_w) {
  if (positive_w && q[0] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// float 4x4 (Map::vmCameraPose / vmRigidMotion) -> Pose through a unit quaternion, like Converter::toSE3Quat
VHD void pose_from_f32(const float* T, Pose& X) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double q[4];
  quat_from_R(R, q);
  quat_unit(q, true);
  R_from_quat(q, X.R);
  X.t[0] = T[3]; X.t[1] = T[7]; X.t[2] = T[11];
}
// Pose -> float 4x4 through getEstimateData's quaternion (src/Optimizer.cc:1058-1069)
VHD void pose_to_f32(const Pose& X, float* T) {
  double q[4], R[9];
  quat_from_R(X.R, q);
  quat_unit(q, false);
  R_from_quat(q, R);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R[3 * i + j];
    T[4 * i + 3] = (float)X.t[i];
  }
  T[12] = 0.f; T[13] = 0.f; T[14] = 0.f; T[15] = 1.f;
}

// X <- X * [t = u[0..2], R = R(q), q = (sqrt(1-|v|^2), v = u[3..5])]  (identity rotation when |v| > 1)
VHD void pose_oplus(const Pose& X, const double* u, Pose& o) {
  Pose inc;
  const double w2 = 1 - (u[3] * u[3] + u[4] * u[4] + u[5] * u[5]);
  if (w2 < 0) {
    for (int i = 0; i < 9; i++) inc.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  } else {
    const double q[4] = {sqrt(w2), u[3], u[4], u[5]};
    R_from_quat(q, inc.R);
  }
  inc.t[0] = u[0]; inc.t[1] = u[1]; inc.t[2] = u[2];
  pose_mul(X, inc, o);
}

// Huber: rho(e) and weight rho'(e)
VHD void huber(double e, double delta, double& rho0, double& w) {
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho0 = e; w = 1.0; }
  else {
    const double s = sqrt(e);
    rho0 = 2 * s * delta - dsqr;
    w = delta / s;
  }
}

// EdgeSE3: e = [t_E ; vec(q_E)], E = Zinv * Xi^-1 * Xj.  Ji/Jj are 6x6 row-major; pass nullptr to skip.
VHD void edge_se3(const Pose& Xi, const Pose& Xj, const Pose& Zinv, double* e, double* Ji, double* Jj) {
  Pose B, E;
  pose_inv_mul(Xi, Xj, B);
  pose_mul(Zinv, B, E);
  double q[4];
  quat_from_R(E.R, q);
  quat_unit(q, true);
  e[0] = E.t[0]; e[1] = E.t[1]; e[2] = E.t[2];
  e[3] = q[1]; e[4] = q[2]; e[5] = q[3];
  if (!Ji) return;
  for (int i = 0; i < 36; i++) { Ji[i] = 0; Jj[i] = 0; }
  // translation rows: dt/du_i = -Ra, dt/dv_i = Ra * 2[tb]x, dt/du_j = Re
  const double tb0 = 2 * B.t[0], tb1 = 2 * B.t[1], tb2 = 2 * B.t[2];
  const double S[9] = {0, -tb2, tb1, tb2, 0, -tb0, -tb1, tb0, 0};
  double RaS[9];
  mat3_mul(Zinv.R, S, RaS);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Ji[6 * i + j] = -Zinv.R[3 * i + j];
      Ji[6 * i + 3 + j] = RaS[3 * i + j];
      Jj[6 * i + j] = E.R[3 * i + j];
    }
  // rotation rows: q' = q (x) (1, v)  =>  d vec/dv = w I + [vec]x ; for Xi the increment is -Rb^T v_i
  const double G[9] = {q[0], -q[3], q[2], q[3], q[0], -q[1], -q[2], q[1], q[0]};
  double GRbT[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) GRbT[3 * i + j] = G[3 * i] * B.R[3 * j] + G[3 * i + 1] * B.R[3 * j + 1] + G[3 * i + 2] * B.R[3 * j + 2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Jj[6 * (3 + i) + 3 + j] = G[3 * i + j];
      Ji[6 * (3 + i) + 3 + j] = -GRbT[3 * i + j];
    }
}

// EdgeSE3PointXYZ (offset = identity): zc = X^-1 p, e = zc - z.
VHD void edge_xyz(const Pose& X, const double* p, const double* z, double* zc, double* e) {
  const double d0 = p[0] - X.t[0], d1 = p[1] - X.t[1], d2 = p[2] - X.t[2];
  for (int i = 0; i < 3; i++) {
    zc[i] = X.R[i] * d0 + X.R[3 + i] * d1 + X.R[6 + i] * d2;
    e[i] = zc[i] - z[i];
  }
}
// J_pose (3x6) = [-I | Q(zc)], Q = 2 * [[0,-z,y],[z,0,-x],[-y,x,0]];  J_point (3x3) = R^T

}  // namespace vb
