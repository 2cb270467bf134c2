// se3q_math.h -- SE3Quat device helpers shared by the reprojection-only optimiser (same formulas as poseopt_kernels.cu):
//   g2o::SE3Quat::exp / operator* / map      g2o/types/se3quat.h:100-106, 228-262
//   Eigen::Quaterniond(Matrix3d) branches    (un-vendored Eigen3: published algorithm)
#pragma once
#include <math.h>

namespace se3q {

struct PoseQ {
  double q[4];  // w x y z
  double t[3];
};

static __device__ __forceinline__ void q_rot(const double* q, const double* v, double* o) {
  const double ux = q[1], uy = q[2], uz = q[3], w = q[0];
  double cx = 2 * (uy * v[2] - uz * v[1]), cy = 2 * (uz * v[0] - ux * v[2]), cz = 2 * (ux * v[1] - uy * v[0]);
  o[0] = v[0] + w * cx + (uy * cz - uz * cy);
  o[1] = v[1] + w * cy + (uz * cx - ux * cz);
  o[2] = v[2] + w * cz + (ux * cy - uy * cx);
}

static __device__ void quat_from_R_dev(const double* R, double* q) {
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
static __device__ void se3_exp_mul(const double* u, const PoseQ& T, PoseQ& o) {
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


}  // namespace se3q
