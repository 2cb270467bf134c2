// g2o_math.h -- small fixed-size double math for the oracle (TEST INFRASTRUCTURE ONLY, see vido_oracle.h).
// Restates the Eigen / g2o slam3d helpers the reference's optimisation path uses:
//   Eigen::Quaterniond(Matrix3d), toRotationMatrix        (Eigen3, un-vendored: published algorithm)
//   g2o::internal::normalize / toCompactQuaternion / fromCompactQuaternion  g2o/types/isometry3d_mappings.cpp:44-108
//   g2o::SE3Quat (ctor, normalizeRotation, exp, map)       g2o/types/se3quat.h:40-300
//   compute_dq_dR                                          g2o/types/dquat2mat.cpp:35-84 (+ maxima-generated partials)
#pragma once
#include <cmath>
#include <cstring>

namespace vo {

struct V3 { double x, y, z; };
struct M3 { double m[9]; };  // row-major
struct Iso { M3 R; V3 t; };  // Eigen::Isometry3d (rotation block taken verbatim, extractRotation)
struct Quat { double w, x, y, z; };

inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, const V3& a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline M3 m3_identity() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[3 * i + j] = a.m[3 * i] * b.m[j] + a.m[3 * i + 1] * b.m[3 + j] + a.m[3 * i + 2] * b.m[6 + j];
  return r;
}
inline V3 mul(const M3& a, const V3& v) {
  return {a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
          a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
inline M3 transpose(const M3& a) { return {{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
inline Iso iso_identity() { return {m3_identity(), {0, 0, 0}}; }
inline Iso mul(const Iso& a, const Iso& b) { return {mul(a.R, b.R), mul(a.R, b.t) + a.t}; }
inline Iso inverse(const Iso& a) {  // Eigen Isometry inverse: R^T, -R^T t
  M3 Rt = transpose(a.R);
  V3 t = mul(Rt, a.t);
  return {Rt, {-t.x, -t.y, -t.z}};
}
inline V3 apply(const Iso& a, const V3& p) { return mul(a.R, p) + a.t; }

// Eigen::Quaterniond(const Matrix3d&)
inline Quat quat_from_R(const M3& R) {
  auto m = [&](int i, int j) { return R.m[3 * i + j]; };
  Quat q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t;
    q.y = (m(0, 2) - m(2, 0)) * t;
    q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t;
    v[j] = (m(j, i) + m(i, j)) * t;
    v[k] = (m(k, i) + m(i, k)) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
// Eigen QuaternionBase::toRotationMatrix
inline M3 quat_to_R(const Quat& q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  return {{1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)}};
}
inline Quat quat_mul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
// SE3Quat::normalizeRotation / g2o::internal::normalize: w >= 0, unit norm
inline Quat quat_normalized(Quat q) {
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return {q.w / n, q.x / n, q.y / n, q.z / n};
}
// internal::toCompactQuaternion (isometry3d_mappings.cpp:78-83)
inline V3 compact_quat(const M3& R) {
  Quat q = quat_normalized(quat_from_R(R));
  return {q.x, q.y, q.z};
}
// internal::fromCompactQuaternion (isometry3d_mappings.cpp:85-92)
inline M3 from_compact_quat(const V3& v) {
  double w = 1 - (v.x * v.x + v.y * v.y + v.z * v.z);
  if (w < 0) return m3_identity();
  w = std::sqrt(w);
  return quat_to_R({w, v.x, v.y, v.z});
}
// Converter::toSE3Quat(cv::Mat float 4x4) then SE3Quat -> Isometry3 (src/Converter.cc:29-39, se3quat.h:56-58,283-288)
inline Iso iso_from_f32(const float* T) {
  M3 R = {{T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]}};
  Quat q = quat_normalized(quat_from_R(R));
  return {quat_to_R(q), {T[3], T[7], T[11]}};
}
// VertexSE3::getEstimateData -> Quaterniond(qw,qx,qy,qz).matrix() -> Converter::toCvSE3 (src/Optimizer.cc:1058-1069)
inline void iso_to_f32(const Iso& X, float* T) {
  Quat q = quat_from_R(X.R);
  double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);  // toVectorQT: q.normalize() (no sign flip)
  q = {q.w / n, q.x / n, q.y / n, q.z / n};
  M3 R = quat_to_R(q);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[4 * i + j] = (float)R.m[3 * i + j];
  }
  T[3] = (float)X.t.x; T[7] = (float)X.t.y; T[11] = (float)X.t.z;
  T[12] = T[13] = T[14] = 0.f; T[15] = 1.f;
}

// d(compact quaternion)/d(R) -- 3 x 9, columns ordered r00,r10,r20,r01,r11,r21,r02,r12,r22 (column-major R),
// i.e. the exact partials of Eigen's matrix->quaternion branches (dquat2mat.cpp:35-84); sign flipped when qw <= 0.
inline void dq_dR(const M3& R, double J[3][9]) {
  auto r = [&](int i, int j) { return R.m[3 * i + j]; };
  memset(J, 0, sizeof(double) * 27);
  const double tr = r(0, 0) + r(1, 1) + r(2, 2);
  auto col = [](int i, int j) { return 3 * j + i; };  // column index of r_ij
  double qw;
  if (tr > 0) {
    double S = std::sqrt(tr + 1.0) * 2;  // 4 qw
    qw = 0.25 * S;
    double q = qw, a = 0.25 / q, c = -0.03125 / (q * q * q);
    // qx = (r21 - r12)/(4qw), qw = sqrt(1+tr)/2  =>  dqx/dr_ii = -(r21-r12)/(32 qw^3) ...
    double dx = (r(2, 1) - r(1, 2)) * c, dy = (r(0, 2) - r(2, 0)) * c, dz = (r(1, 0) - r(0, 1)) * c;
    for (int d = 0; d < 3; d++) { J[0][col(d, d)] = dx; J[1][col(d, d)] = dy; J[2][col(d, d)] = dz; }
    J[0][col(2, 1)] = a; J[0][col(1, 2)] = -a;
    J[1][col(0, 2)] = a; J[1][col(2, 0)] = -a;
    J[2][col(1, 0)] = a; J[2][col(0, 1)] = -a;
  } else {
    int i = 0;
    if ((r(0, 0) > r(1, 1)) & (r(0, 0) > r(2, 2))) i = 0;
    else if (r(1, 1) > r(2, 2)) i = 1;
    else i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double S = std::sqrt(1.0 + r(i, i) - r(j, j) - r(k, k)) * 2;  // 4 q_i
    qw = (r(k, j) - r(j, k)) / S;
    double qi = 0.25 * S;
    // q_i = S/4, dq_i/dr_ii = 1/(8 q_i) *... : q_i = sqrt(1+rii-rjj-rkk)/2
    double dqi = 0.125 / qi;  // d q_i / d r_ii ; -dqi for r_jj, r_kk
    J[i][col(i, i)] = dqi; J[i][col(j, j)] = -dqi; J[i][col(k, k)] = -dqi;
    // q_j = (r_ji + r_ij)/(4 q_i); q_k = (r_ki + r_ik)/(4 q_i)
    double a = 0.25 / qi;
    double cj = -(r(j, i) + r(i, j)) / (4 * qi * qi), ck = -(r(k, i) + r(i, k)) / (4 * qi * qi);
    J[j][col(j, i)] = a; J[j][col(i, j)] = a;
    J[j][col(i, i)] = cj * dqi; J[j][col(j, j)] = -cj * dqi; J[j][col(k, k)] = -cj * dqi;
    J[k][col(k, i)] = a; J[k][col(i, k)] = a;
    J[k][col(i, i)] = ck * dqi; J[k][col(j, j)] = -ck * dqi; J[k][col(k, k)] = -ck * dqi;
  }
  if (qw <= 0)
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 9; b++) J[a][b] = -J[a][b];
}

// SE3Quat::exp (se3quat.h:228-262): update = [omega, upsilon]
inline void se3_exp(const double u[6], Quat& q, V3& t) {
  V3 om = {u[0], u[1], u[2]}, up = {u[3], u[4], u[5]};
  double theta = std::sqrt(dot(om, om));
  M3 Om = {{0, -om.z, om.y, om.z, 0, -om.x, -om.y, om.x, 0}};
  M3 Om2 = mul(Om, Om);
  M3 R, V;
  M3 I = m3_identity();
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) R.m[i] = I.m[i] + Om.m[i] + Om2.m[i];
    V = R;
  } else {
    double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta);
    double c = (theta - std::sin(theta)) / std::pow(theta, 3);
    for (int i = 0; i < 9; i++) {
      R.m[i] = I.m[i] + a * Om.m[i] + b * Om2.m[i];
      V.m[i] = I.m[i] + b * Om.m[i] + c * Om2.m[i];
    }
  }
  q = quat_normalized(quat_from_R(R));
  t = mul(V, up);
}
inline V3 quat_rotate(const Quat& q, const V3& v) {  // Eigen: q * v  (via rotation matrix expansion)
  V3 u = {q.x, q.y, q.z};
  V3 uv = {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
  uv = 2.0 * uv;
  V3 uuv = {u.y * uv.z - u.z * uv.y, u.z * uv.x - u.x * uv.z, u.x * uv.y - u.y * uv.x};
  return v + (q.w * uv) + uuv;
}

// Huber kernel, g2o/core/robust_kernel_impl.cpp:78-91
inline void huber(double e, double delta, double rho[3]) {
  double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
  else {
    double sqrte = std::sqrt(e);
    rho[0] = 2 * sqrte * delta - dsqr;
    rho[1] = delta / sqrte;
    rho[2] = -0.5 * rho[1] / e;
  }
}

}  // namespace vo
