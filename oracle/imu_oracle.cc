// imu_oracle.cc -- CPU restatement of the IMU preintegration.  TEST INFRASTRUCTURE ONLY.
//   Tracking::PreintegrateIMU               src/Tracking.cc:784-887 (queue selection, end-point interpolation, mid-point rule)
//   IMU::Preintegrated::IntegrateNewMeasurement  src/ImuTypes.cc:245-300
//   IMU::IntegratedRotation                 src/ImuTypes.cc:143-168
//   IMU::NormalizeRotation (cv::SVDecomp, U*Vt)  src/ImuTypes.cc:20-26 -> orthogonal polar factor (same matrix)
// float32 state, every cv::Mat expression evaluated in double and rounded once ("parity unpinned": OpenCV's exact
// rounding order inside MatExpr is not reproducible; tolerance-checked at 1e-5).
#include <cmath>
#include <cstring>
#include <vector>

#include "vido_oracle.h"

namespace {

typedef double M33[9];

void mul33(const double* a, const double* b, double* o) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(o, t, sizeof t);
}
void f2d(const float* a, double* o, int n) { for (int i = 0; i < n; i++) o[i] = a[i]; }
void d2f(const double* a, float* o, int n) { for (int i = 0; i < n; i++) o[i] = (float)a[i]; }

void inv33(const double* m, double* o) {
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  const double id = 1.0 / det;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// orthogonal polar factor U*Vt of a near-rotation (Newton iteration X <- (X + X^-T)/2)
void normalize_rotation(const double* R, float* out) {
  double X[9];
  memcpy(X, R, sizeof X);
  for (int it = 0; it < 8; it++) {
    double Xi[9];
    inv33(X, Xi);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) X[3 * i + j] = 0.5 * (X[3 * i + j] + Xi[3 * j + i]);
  }
  d2f(X, out, 9);
}

struct State {
  float dT, dR[9], dV[3], dP[3], JRg[9], JVg[9], JVa[9], JPg[9], JPa[9], C[225], avgA[3], avgW[3];
};

void integrate(State& s, const float* bias, const float* Nga, const float* NgaWalk, const float a3[3], const float w3[3], float dt) {
  const double acc[3] = {(double)(float)(a3[0] - bias[0]), (double)(float)(a3[1] - bias[1]), (double)(float)(a3[2] - bias[2])};
  const double accW[3] = {(double)(float)(w3[0] - bias[3]), (double)(float)(w3[1] - bias[4]), (double)(float)(w3[2] - bias[5])};
  double dR[9], dV[3], dP[3], JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
  f2d(s.dR, dR, 9); f2d(s.dV, dV, 3); f2d(s.dP, dP, 3); f2d(s.JRg, JRg, 9); f2d(s.JVg, JVg, 9); f2d(s.JVa, JVa, 9);
  f2d(s.JPg, JPg, 9); f2d(s.JPa, JPa, 9);
  const double t = dt, T = s.dT;
  double Ra[3];
  for (int i = 0; i < 3; i++) Ra[i] = dR[3 * i] * acc[0] + dR[3 * i + 1] * acc[1] + dR[3 * i + 2] * acc[2];
  for (int i = 0; i < 3; i++) {
    s.avgA[i] = (float)((T * (double)s.avgA[i] + Ra[i] * t) / (T + t));
    s.avgW[i] = (float)((T * (double)s.avgW[i] + accW[i] * t) / (T + t));
  }
  for (int i = 0; i < 3; i++) {
    s.dP[i] = (float)(dP[i] + dV[i] * t + 0.5 * Ra[i] * t * t);
    s.dV[i] = (float)(dV[i] + Ra[i] * t);
  }
  const double Wacc[9] = {0, -acc[2], acc[1], acc[2], 0, -acc[0], -acc[1], acc[0], 0};
  double RW[9], RWJ[9];
  mul33(dR, Wacc, RW);
  mul33(RW, JRg, RWJ);
  double A[81], B[54];
  memset(A, 0, sizeof A); memset(B, 0, sizeof B);
  for (int i = 0; i < 9; i++) A[10 * i] = 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[9 * (3 + i) + j] = (float)(-RW[3 * i + j] * t);
      A[9 * (6 + i) + j] = (float)(-0.5 * RW[3 * i + j] * t * t);
      A[9 * (6 + i) + 3 + j] = (i == j) ? (double)(float)t : 0.0;
      B[6 * (3 + i) + 3 + j] = (float)(dR[3 * i + j] * t);
      B[6 * (6 + i) + 3 + j] = (float)(0.5 * dR[3 * i + j] * t * t);
    }
  for (int k = 0; k < 9; k++) {
    s.JPa[k] = (float)(JPa[k] + JVa[k] * t - 0.5 * dR[k] * t * t);
    s.JPg[k] = (float)(JPg[k] + JVg[k] * t - 0.5 * RWJ[k] * t * t);
    s.JVa[k] = (float)(JVa[k] - dR[k] * t);
    s.JVg[k] = (float)(JVg[k] - RWJ[k] * t);
  }
  // IntegratedRotation
  const float x = (float)((w3[0] - bias[3]) * dt), y = (float)((w3[1] - bias[4]) * dt), z = (float)((w3[2] - bias[5]) * dt);
  const float d2 = x * x + y * y + z * z;
  const float d = std::sqrt(d2);
  const double W[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  double W2[9], dRi[9], rJ[9];
  mul33(W, W, W2);
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (d < 1e-4f) {
    for (int k = 0; k < 9; k++) { dRi[k] = (float)(I[k] + W[k]); rJ[k] = I[k]; }
  } else {
    const double sd = std::sin((double)d), cd = std::cos((double)d);
    for (int k = 0; k < 9; k++) {
      dRi[k] = (float)(I[k] + W[k] * sd / d + W2[k] * (1.0 - cd) / d2);
      rJ[k] = (float)(I[k] - W[k] * (1.0 - cd) / d2 + W2[k] * (d - sd) / ((double)d2 * d));
    }
  }
  double RdR[9];
  mul33(dR, dRi, RdR);
  for (int k = 0; k < 9; k++) RdR[k] = (float)RdR[k];
  normalize_rotation(RdR, s.dR);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[9 * i + j] = dRi[3 * j + i];
      B[6 * i + j] = (float)(rJ[3 * i + j] * t);
    }
  // C[0:9,0:9] = A C A^T + B Nga B^T ; C[9:15,9:15] += NgaWalk
  double C9[81], AC[81], NB[54];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) C9[9 * i + j] = s.C[15 * i + j];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      double v = 0;
      for (int k = 0; k < 9; k++) v += A[9 * i + k] * C9[9 * k + j];
      AC[9 * i + j] = (float)v;
    }
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 6; j++) {
      double v = 0;
      for (int k = 0; k < 6; k++) v += B[6 * i + k] * (double)Nga[6 * k + j];
      NB[6 * i + j] = (float)v;
    }
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      double v1 = 0, v2 = 0;
      for (int k = 0; k < 9; k++) v1 += AC[9 * i + k] * A[9 * j + k];
      for (int k = 0; k < 6; k++) v2 += NB[6 * i + k] * B[6 * j + k];
      s.C[15 * i + j] = (float)((double)(float)v1 + (double)(float)v2);
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) s.C[15 * (9 + i) + 9 + j] = (float)((double)s.C[15 * (9 + i) + 9 + j] + (double)NgaWalk[6 * i + j]);
  // JRg = dRi^T JRg - rightJ dt
  double tJ[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double v = 0;
      for (int k = 0; k < 3; k++) v += dRi[3 * k + i] * JRg[3 * k + j];
      tJ[3 * i + j] = v - rJ[3 * i + j] * t;
    }
  d2f(tJ, s.JRg, 9);
  s.dT = (float)(T + t);
}

}  // namespace

extern "C" int vo_imu_preintegrate(const vo_imu_sample* q, int n, double t_prev, double t_cur, const float* bias,
                                   const float* noise, vo_imu_preint* out) {
  // queue selection (src/Tracking.cc:806-838)
  std::vector<vo_imu_sample> v;
  int consumed = 0;
  for (int i = 0; i < n; i++) {
    const long double t = q[i].t;
    if (t < (long double)t_prev - 0.001L) { consumed++; continue; }
    else if (t < (long double)t_cur - 0.001L) { v.push_back(q[i]); consumed++; }
    else { v.push_back(q[i]); break; }
  }
  State s;
  memset(&s, 0, sizeof s);
  s.dR[0] = s.dR[4] = s.dR[8] = 1.f;
  float Nga[36] = {0}, NgaWalk[36] = {0};
  for (int i = 0; i < 3; i++) {
    Nga[7 * i] = noise[0] * noise[0]; Nga[7 * (3 + i)] = noise[1] * noise[1];
    NgaWalk[7 * i] = noise[2] * noise[2]; NgaWalk[7 * (3 + i)] = noise[3] * noise[3];
  }
  const int m = (int)v.size() - 1;
  for (int i = 0; i < m; i++) {
    float tstep, acc[3], ang[3];
    const float* a0 = &v[i].ax; const float* a1 = &v[i + 1].ax;
    const float* w0 = &v[i].wx; const float* w1 = &v[i + 1].wx;
    if (i == 0 && i < m - 1) {
      const float tab = (float)(v[i + 1].t - v[i].t), tini = (float)(v[i].t - t_prev);
      for (int k = 0; k < 3; k++) {
        acc[k] = (a0[k] + a1[k] - (a1[k] - a0[k]) * (tini / tab)) * 0.5f;
        ang[k] = (w0[k] + w1[k] - (w1[k] - w0[k]) * (tini / tab)) * 0.5f;
      }
      tstep = (float)(v[i + 1].t - t_prev);
    } else if (i < m - 1) {
      for (int k = 0; k < 3; k++) { acc[k] = (a0[k] + a1[k]) * 0.5f; ang[k] = (w0[k] + w1[k]) * 0.5f; }
      tstep = (float)(v[i + 1].t - v[i].t);
    } else if (i > 0 && i == m - 1) {
      const float tab = (float)(v[i + 1].t - v[i].t), tend = (float)(v[i + 1].t - t_cur);
      for (int k = 0; k < 3; k++) {
        acc[k] = (a0[k] + a1[k] - (a1[k] - a0[k]) * (tend / tab)) * 0.5f;
        ang[k] = (w0[k] + w1[k] - (w1[k] - w0[k]) * (tend / tab)) * 0.5f;
      }
      tstep = (float)(t_cur - v[i].t);
    } else {
      for (int k = 0; k < 3; k++) { acc[k] = a0[k]; ang[k] = w0[k]; }
      tstep = (float)(t_cur - t_prev);
    }
    integrate(s, bias, Nga, NgaWalk, acc, ang, tstep);
  }
  out->dT = s.dT;
  memcpy(out->dR, s.dR, sizeof s.dR); memcpy(out->dV, s.dV, sizeof s.dV); memcpy(out->dP, s.dP, sizeof s.dP);
  memcpy(out->JRg, s.JRg, sizeof s.JRg); memcpy(out->JVg, s.JVg, sizeof s.JVg); memcpy(out->JVa, s.JVa, sizeof s.JVa);
  memcpy(out->JPg, s.JPg, sizeof s.JPg); memcpy(out->JPa, s.JPa, sizeof s.JPa); memcpy(out->C, s.C, sizeof s.C);
  memcpy(out->avgA, s.avgA, sizeof s.avgA); memcpy(out->avgW, s.avgW, sizeof s.avgW);
  out->n_steps = m > 0 ? m : 0;
  out->n_consumed = consumed;
  return out->n_steps;
}
