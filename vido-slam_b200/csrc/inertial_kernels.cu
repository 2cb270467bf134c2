// inertial_kernels.cu -- inertial-only optimisation of the VIO initialisation on the GPU (one launch per optimisation).
//
// Replaces Optimizer::InertialOptimization (src/Optimizer.cc:2441-2620) and the g2o types it instantiates:
//   EdgeInertialGS (information from the preintegrated covariance, computeError, linearizeOplus)  src/G2oTypes.cc:357-482
//   EdgePriorAcc / EdgePriorGyro  include/G2oTypes.h:578-624, src/G2oTypes.cc:526-538 (error = prior - estimate, Jacobian +I as written)
//   VertexVelocity / VertexGyroBias / VertexAccBias / VertexGDir / VertexScale                    include/G2oTypes.h:144-284
//   IMU::Preintegrated::GetDeltaRotation / Velocity / Position (float32 arithmetic)               src/ImuTypes.cc:339-368
//   ExpSO3 / LogSO3 / RightJacobianSO3 / InverseRightJacobianSO3                                  src/G2oTypes.cc:541-613
//   LM driver: lm_device.h (user lambda 1e3, 200 iterations, no terminate action)
// Poses are fixed.  Unknowns in g2o's vertex-id order: one velocity per frame, gyro bias, acc bias, gravity direction (2),
// scale (1).  H is block-tridiagonal in the velocities (an edge joins consecutive frames) with a dense 9-wide border, so
// (H + lambda I) x = b is solved exactly by eliminating the velocities along the chain (3x3 pivots carrying a 9x3 border block)
// and factoring the remaining 9x9 -- the same pivot order as a Cholesky of the full matrix (the reference: SimplicialLDLT).
// One CTA: threads linearise the edges in parallel, one thread runs the O(N) chain; everything else is reductions.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <vector>

#include "ctx.h"
#include "lm_device.h"

namespace {

constexpr int IN_THREADS = 256;
#define GRAVITY_VALUE ((double)9.79f)   // include/ImuTypes.h:29 (a float constant)

struct InertialArgs {
  int N, its, mode;
  const double* Rwb;   // [N][9]
  const double* twb;   // [N][3]
  const vido_imu_preint* pre;  // [N-1]
  const float* blin;   // [N-1][6]
  const double* Info;  // [N-1][81]
  double priorG, priorA, user_lambda;
  // state, double-buffered: V [2][N][3]; glob [2][16] = bg(3) ba(3) Rwg(9) s(1)
  double* V;
  double* glob;
  // workspace
  double* He;    // [N-1][225] per-edge J^T Omega J (15 x 15 local columns: V1 3 | bg 3 | ba 3 | V2 3 | gdir 2 | scale 1)
  double* be;    // [N-1][15]
  double* D;     // [N][9] velocity diagonal blocks
  double* O;     // [N][9] coupling V_i - V_{i+1}
  double* Bv;    // [N][27] velocity - border (3 x 9)
  double* bv;    // [N][3]
  double* x;     // [3N + 9]
  double* L;     // [N][9] chain factors, M [N][9], Y [N][27], c [N][3]
  double* M;
  double* Y;
  double* cc;
  LmCtl* ctl;
  LmRec* rec;
};

__device__ __forceinline__ void mul33(const double* a, const double* b, double* o) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  for (int k = 0; k < 9; k++) o[k] = t[k];
}
__device__ __forceinline__ void mul33t(const double* a, const double* b, double* o) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
  for (int k = 0; k < 9; k++) o[k] = t[k];
}
__device__ __forceinline__ void mulv(const double* a, const double* v, double* o) {
  const double t0 = a[0] * v[0] + a[1] * v[1] + a[2] * v[2], t1 = a[3] * v[0] + a[4] * v[1] + a[5] * v[2],
               t2 = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
__device__ __forceinline__ void multv(const double* a, const double* v, double* o) {
  const double t0 = a[0] * v[0] + a[3] * v[1] + a[6] * v[2], t1 = a[1] * v[0] + a[4] * v[1] + a[7] * v[2],
               t2 = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
  o[0] = t0; o[1] = t1; o[2] = t2;
}
__device__ void inv33(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// IMU::NormalizeRotation of a float32 matrix (cv::SVDecomp U*Vt = orthogonal polar factor), result rounded to float32
__device__ void normalize_rotation_f(const double* R, double* out) {
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
__device__ __forceinline__ void skew(const double* v, double* W) {
  W[0] = 0; W[1] = -v[2]; W[2] = v[1]; W[3] = v[2]; W[4] = 0; W[5] = -v[0]; W[6] = -v[1]; W[7] = v[0]; W[8] = 0;
}
__device__ void exp_so3(double x, double y, double z, double* R) {   // src/G2oTypes.cc:546-562
  const double d2 = x * x + y * y + z * z, d = sqrt(d2);
  const double v[3] = {x, y, z};
  double W[9], W2[9], res[9];
  skew(v, W);
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    res[k] = d < 1e-5 ? I + W[k] + 0.5 * W2[k] : I + W[k] * sin(d) / d + W2[k] * (1.0 - cos(d)) / d2;
  }
  normalize_rotation_f(res, R);
}
__device__ void exp_so3_f(float x, float y, float z, double* R) {    // src/ImuTypes.cc:38-50, entries rounded to float32
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float d = __fsqrt_rn(d2);
  const double W[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  double W2[9];
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    const double v = d < 1e-4f ? I + W[k] + 0.5 * (double)(float)W2[k]
                               : I + W[k] * sin((double)d) / d + (double)(float)W2[k] * (1.0 - cos((double)d)) / d2;
    R[k] = (double)(float)v;
  }
}
__device__ void log_so3(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (tr - 1.0) * 0.5;
  if (costheta > 1 || costheta < -1) return;
  const double theta = acos(costheta), s = sin(theta);
  if (fabs(s) < 1e-5) return;
  for (int k = 0; k < 3; k++) w[k] = theta * w[k] / s;
}
__device__ void inv_right_jacobian(const double* v, double* J) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    J[k] = d < 1e-5 ? I : I + W[k] / 2 + W2[k] * (1.0 / d2 - (1.0 + cos(d)) / (2.0 * d * sin(d)));
  }
}
__device__ void right_jacobian(const double* v, double* J) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mul33(W, W, W2);
  for (int k = 0; k < 9; k++) {
    const double I = (k % 4 == 0) ? 1.0 : 0.0;
    J[k] = d < 1e-5 ? I : I - W[k] * (1.0 - cos(d)) / d2 + W2[k] * (d - sin(d)) / (d2 * d);
  }
}

// products of float32 values are exact in double; "(float)" marks the single rounding of a cv::Mat float expression
__device__ void deltas(const InertialArgs& a, const double* glob, int e, double* dR, double* dV, double* dP, double* dbg_d) {
  const vido_imu_preint& p = a.pre[e];
  const float* bl = a.blin + 6 * (size_t)e;
  float dbg[3], dba[3];
  for (int k = 0; k < 3; k++) { dba[k] = __fsub_rn((float)glob[3 + k], bl[k]); dbg[k] = __fsub_rn((float)glob[k], bl[3 + k]); }
  for (int k = 0; k < 3; k++) dbg_d[k] = dbg[k];
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
    dV[i] = (double)__fadd_rn(__fadd_rn(p.dV[i], t1), t2);
    const float u1 = (float)((double)p.JPg[3 * i] * dbg[0] + (double)p.JPg[3 * i + 1] * dbg[1] + (double)p.JPg[3 * i + 2] * dbg[2]);
    const float u2 = (float)((double)p.JPa[3 * i] * dba[0] + (double)p.JPa[3 * i + 1] * dba[1] + (double)p.JPa[3 * i + 2] * dba[2]);
    dP[i] = (double)__fadd_rn(__fadd_rn(p.dP[i], u1), u2);
  }
}

// EdgeInertialGS::computeError (+ linearizeOplus when J != nullptr; J is 9 x 15, local columns as in InertialArgs::He)
__device__ void edge_eval(const InertialArgs& a, const double* V, const double* glob, int e, double* er9, double* J) {
  const double* R1 = a.Rwb + 9 * (size_t)e;
  const double* R2 = a.Rwb + 9 * (size_t)(e + 1);
  const double* t1 = a.twb + 3 * (size_t)e;
  const double* t2 = a.twb + 3 * (size_t)(e + 1);
  const double* V1 = V + 3 * (size_t)e;
  const double* V2 = V + 3 * (size_t)(e + 1);
  const double* Rwg = glob + 6;
  const double s = glob[15];
  const double dt = (double)a.pre[e].dT;
  const double gI[3] = {0, 0, -GRAVITY_VALUE};
  double g[3], dR[9], dV[3], dP[3], dbg[3];
  mulv(Rwg, gI, g);
  deltas(a, glob, e, dR, dV, dP, dbg);
  double R12[9], eR[9], er[3];
  mul33t(R1, R2, R12);
  mul33t(dR, R12, eR);
  log_so3(eR, er);
  double av[3], cv[3], ra[3], rc[3];
  for (int k = 0; k < 3; k++) {
    av[k] = s * (V2[k] - V1[k]) - g[k] * dt;
    cv[k] = s * (t2[k] - t1[k] - V1[k] * dt) - g[k] * dt * dt / 2;
  }
  multv(R1, av, ra);
  multv(R1, cv, rc);
  for (int k = 0; k < 3; k++) { er9[k] = er[k]; er9[3 + k] = ra[k] - dV[k]; er9[6 + k] = rc[k] - dP[k]; }
  if (!J) return;
  for (int k = 0; k < 135; k++) J[k] = 0;
  double Rbw1[9];
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) Rbw1[3 * r + q] = R1[3 * q + r];
  double invJr[9];
  inv_right_jacobian(er, invJr);
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) {
      J[15 * (3 + r) + q] = -s * Rbw1[3 * r + q];            // velocity 1
      J[15 * (6 + r) + q] = -s * dt * Rbw1[3 * r + q];
      J[15 * (3 + r) + 9 + q] = s * Rbw1[3 * r + q];         // velocity 2
      J[15 * (3 + r) + 3 + q] = -(double)a.pre[e].JVg[3 * r + q];   // gyro bias
      J[15 * (6 + r) + 3 + q] = -(double)a.pre[e].JPg[3 * r + q];
      J[15 * (3 + r) + 6 + q] = -(double)a.pre[e].JVa[3 * r + q];   // acc bias
      J[15 * (6 + r) + 6 + q] = -(double)a.pre[e].JPa[3 * r + q];
    }
  {
    double JRg[9], v[3], Jr[9], T1[9], T2[9], T3[9];
    for (int k = 0; k < 9; k++) JRg[k] = a.pre[e].JRg[k];
    mulv(JRg, dbg, v);
    right_jacobian(v, Jr);
    mul33t(eR, Jr, T1);
    mul33(T1, JRg, T2);
    mul33(invJr, T2, T3);
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 3; q++) J[15 * r + 3 + q] = -T3[3 * r + q];
  }
  {
    double dG[6];
    for (int r = 0; r < 3; r++) { dG[2 * r] = Rwg[3 * r + 1] * GRAVITY_VALUE; dG[2 * r + 1] = -Rwg[3 * r] * GRAVITY_VALUE; }
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 2; q++) {
        double v = 0;
        for (int k = 0; k < 3; k++) v += Rbw1[3 * r + k] * dG[2 * k + q];
        J[15 * (3 + r) + 12 + q] = -v * dt;
        J[15 * (6 + r) + 12 + q] = -0.5 * v * dt * dt;
      }
  }
  {
    double dv[3], dp[3], o1[3], o2[3];
    for (int k = 0; k < 3; k++) { dv[k] = V2[k] - V1[k]; dp[k] = t2[k] - t1[k] - V1[k] * dt; }
    multv(R1, dv, o1);
    multv(R1, dp, o2);
    for (int r = 0; r < 3; r++) { J[15 * (3 + r) + 14] = o1[r]; J[15 * (6 + r) + 14] = o2[r]; }
  }
}

// deterministic block reduction (fixed order), result broadcast to all threads
__device__ double block_sum(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
  for (int w = 0; w < IN_THREADS / 32; w++) s += sm[w];
  return s;
}

__device__ double chi2_of(const InertialArgs& a, int sel, double* sm) {
  const double* V = a.V + (size_t)sel * 3 * a.N;
  const double* glob = a.glob + 16 * sel;
  double acc = 0;
  for (int e = threadIdx.x; e + 1 < a.N; e += IN_THREADS) {
    double r[9];
    edge_eval(a, V, glob, e, r, nullptr);
    const double* I9 = a.Info + 81 * (size_t)e;
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) acc += r[i] * I9[9 * i + j] * r[j];
  }
  if (threadIdx.x == 0 && a.mode == 0)
    for (int k = 0; k < 3; k++) acc += a.priorA * glob[3 + k] * glob[3 + k] + a.priorG * glob[k] * glob[k];
  return block_sum(acc, sm);
}

__device__ __forceinline__ bool chol3(const double* D, double* L) {
  double v = D[0];
  if (!(v > 0)) return false;
  L[0] = sqrt(v); L[1] = 0; L[2] = 0;
  L[3] = D[3] / L[0]; L[6] = D[6] / L[0];
  v = D[4] - L[3] * L[3];
  if (!(v > 0)) return false;
  L[4] = sqrt(v); L[5] = 0;
  L[7] = (D[7] - L[6] * L[3]) / L[4];
  v = D[8] - L[6] * L[6] - L[7] * L[7];
  if (!(v > 0)) return false;
  L[8] = sqrt(v);
  return true;
}

__global__ void __launch_bounds__(IN_THREADS) inertial_opt_kernel(InertialArgs a) {
  __shared__ double sm[IN_THREADS / 32];
  __shared__ double Hbb[81], bb[9], Sbb[81];
  __shared__ int s_fail;
  LmCtl* c = a.ctl;
  const int tid = threadIdx.x, N = a.N, E = N - 1;
  if (tid == 0) lm_reset(c);
  __syncthreads();
  for (int it = 0; it < a.its; it++) {
    if (c->stop_flag || !c->ok) break;
    const int cur = c->cur;
    if (it == 0) {
      const double chi = chi2_of(a, cur, sm);
      if (tid == 0) c->currentChi = chi;
    }
    const double* V = a.V + (size_t)cur * 3 * N;
    const double* glob = a.glob + 16 * cur;
    // ---- buildSystem: per-edge 15x15 blocks, then the block-tridiagonal + border assembly
    for (int e = tid; e < E; e += IN_THREADS) {
      double r[9], J[135], Or[9];
      edge_eval(a, V, glob, e, r, J);
      const double* I9 = a.Info + 81 * (size_t)e;
      double* He = a.He + 225 * (size_t)e;
      double* be = a.be + 15 * (size_t)e;
      for (int i = 0; i < 9; i++) {
        double sr = 0;
        for (int k = 0; k < 9; k++) sr += I9[9 * i + k] * r[k];
        Or[i] = sr;
      }
      for (int c1 = 0; c1 < 15; c1++) {
        double sb = 0;
        for (int k = 0; k < 9; k++) sb += J[15 * k + c1] * Or[k];
        be[c1] = -sb;
      }
      for (int c2 = 0; c2 < 15; c2++) {
        double oj[9];
        for (int i = 0; i < 9; i++) {
          double sj = 0;
          for (int k = 0; k < 9; k++) sj += I9[9 * i + k] * J[15 * k + c2];
          oj[i] = sj;
        }
        for (int c1 = 0; c1 < 15; c1++) {
          double sh = 0;
          for (int k = 0; k < 9; k++) sh += J[15 * k + c1] * oj[k];
          He[15 * c1 + c2] = sh;
        }
      }
    }
    __syncthreads();
    // local border columns: bg 3..5, ba 6..8, gdir 12..13, scale 14  ->  border index 0..8
    const int bcol[9] = {3, 4, 5, 6, 7, 8, 12, 13, 14};
    for (int i = tid; i < N; i += IN_THREADS) {
      double* D = a.D + 9 * (size_t)i;
      double* O = a.O + 9 * (size_t)i;
      double* Bv = a.Bv + 27 * (size_t)i;
      double* bv = a.bv + 3 * (size_t)i;
      for (int k = 0; k < 9; k++) { D[k] = 0; O[k] = 0; }
      for (int k = 0; k < 27; k++) Bv[k] = 0;
      bv[0] = bv[1] = bv[2] = 0;
      if (i > 0) {          // V2 role in edge i-1 (local columns 9..11)
        const double* He = a.He + 225 * (size_t)(i - 1);
        const double* be = a.be + 15 * (size_t)(i - 1);
        for (int r = 0; r < 3; r++) {
          bv[r] += be[9 + r];
          for (int q = 0; q < 3; q++) D[3 * r + q] += He[15 * (9 + r) + 9 + q];
          for (int q = 0; q < 9; q++) Bv[9 * r + q] += He[15 * (9 + r) + bcol[q]];
        }
      }
      if (i < E) {          // V1 role in edge i (local columns 0..2)
        const double* He = a.He + 225 * (size_t)i;
        const double* be = a.be + 15 * (size_t)i;
        for (int r = 0; r < 3; r++) {
          bv[r] += be[r];
          for (int q = 0; q < 3; q++) { D[3 * r + q] += He[15 * r + q]; O[3 * r + q] = He[15 * r + 9 + q]; }
          for (int q = 0; q < 9; q++) Bv[9 * r + q] += He[15 * r + bcol[q]];
        }
      }
    }
    if (tid < 90) {  // border block and rhs: fixed-order sums over the edges
      double s = 0;
      if (tid < 81) {
        const int r = bcol[tid / 9], q = bcol[tid % 9];
        for (int e = 0; e < E; e++) s += a.He[225 * (size_t)e + 15 * r + q];
        if (tid / 9 == tid % 9 && a.mode == 0) { if (tid / 9 < 3) s += a.priorG; else if (tid / 9 < 6) s += a.priorA; }
        Hbb[tid] = s;
      } else {
        const int r = tid - 81;
        for (int e = 0; e < E; e++) s += a.be[15 * (size_t)e + bcol[r]];
        if (a.mode == 0) {
          if (r < 3) s -= a.priorG * (0.0 - glob[r]);
          else if (r < 6) s -= a.priorA * (0.0 - glob[r]);
        }
        bb[r] = s;
      }
    }
    __syncthreads();
    if (it == 0 && !(a.user_lambda > 0)) {
      double m = 0;
      if (a.mode == 0)
        for (int i = tid; i < 3 * N; i += IN_THREADS) m = fmax(m, fabs(a.D[9 * (size_t)(i / 3) + 4 * (i % 3)]));
      if (tid < 9 && (a.mode == 0 || tid >= 6)) m = fmax(m, fabs(Hbb[10 * tid]));
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      __syncthreads();
      if ((tid & 31) == 0) sm[tid >> 5] = m;
      __syncthreads();
      if (tid == 0) { for (int w = 1; w < IN_THREADS / 32; w++) m = fmax(m, sm[w]); lm_begin_iteration(c, it, fmax(m, sm[0]), a.user_lambda); }
    } else if (tid == 0) lm_begin_iteration(c, it, 0.0, a.user_lambda);
    __syncthreads();
    // ---- trials
    while (true) {
      const double lambda = c->lambda;
      if (tid == 0 && a.mode == 1) {
        // only gravity direction and scale are free: 3x3 system on the border entries 6..8, every other increment stays 0
        s_fail = 0;
        double A3[9], L3[9], y3[3], x3[3] = {a.x[3 * (size_t)N + 6], a.x[3 * (size_t)N + 7], a.x[3 * (size_t)N + 8]};
        for (int r = 0; r < 3; r++)
          for (int q = 0; q < 3; q++) { A3[3 * r + q] = Hbb[9 * (6 + r) + 6 + q] + (r == q ? lambda : 0.0); L3[3 * r + q] = 0; }
        for (int j = 0; j < 3 && !s_fail; j++) {
          double dd = A3[3 * j + j];
          for (int k = 0; k < j; k++) dd -= L3[3 * j + k] * L3[3 * j + k];
          if (!(dd > 0)) { s_fail = 1; break; }
          L3[3 * j + j] = sqrt(dd);
          for (int i = j + 1; i < 3; i++) {
            double s2 = A3[3 * i + j];
            for (int k = 0; k < j; k++) s2 -= L3[3 * i + k] * L3[3 * j + k];
            L3[3 * i + j] = s2 / L3[3 * j + j];
          }
        }
        if (!s_fail) {
          for (int i = 0; i < 3; i++) { double s2 = bb[6 + i]; for (int k = 0; k < i; k++) s2 -= L3[3 * i + k] * y3[k]; y3[i] = s2 / L3[3 * i + i]; }
          for (int i = 2; i >= 0; i--) { double s2 = y3[i]; for (int k = i + 1; k < 3; k++) s2 -= L3[3 * k + i] * x3[k]; x3[i] = s2 / L3[3 * i + i]; }
          for (int k = 0; k < 3; k++) a.x[3 * (size_t)N + 6 + k] = x3[k];
        }
      }
      if (tid == 0 && a.mode == 0) {
        // forward elimination of the velocity chain; border Schur complement in Sbb / x[3N..]
        s_fail = 0;
        for (int k = 0; k < 81; k++) Sbb[k] = Hbb[k] + ((k / 9 == k % 9) ? lambda : 0.0);
        double bbw[9];
        for (int k = 0; k < 9; k++) bbw[k] = bb[k];
        double Lp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, cprev[3] = {0, 0, 0}, Yp[27];
        for (int k = 0; k < 27; k++) Yp[k] = 0;
        for (int i = 0; i < N && !s_fail; i++) {
          double Dm[9], Mm[9];
          for (int k = 0; k < 9; k++) { Dm[k] = a.D[9 * (size_t)i + k] + ((k % 4 == 0) ? lambda : 0.0); Mm[k] = 0; }
          if (i > 0) {
            const double* Op = a.O + 9 * (size_t)(i - 1);   // T_{i-1,i}; T_{i,i-1} = Op^T
            for (int r = 0; r < 3; r++) {
              const double a0 = Op[r], a1 = Op[3 + r], a2 = Op[6 + r];
              const double y0 = a0 / Lp[0];
              const double y1 = (a1 - y0 * Lp[3]) / Lp[4];
              const double y2 = (a2 - y0 * Lp[6] - y1 * Lp[7]) / Lp[8];
              Mm[3 * r] = y0; Mm[3 * r + 1] = y1; Mm[3 * r + 2] = y2;
            }
            for (int r = 0; r < 3; r++)
              for (int q = 0; q < 3; q++) Dm[3 * r + q] -= Mm[3 * r] * Mm[3 * q] + Mm[3 * r + 1] * Mm[3 * q + 1] + Mm[3 * r + 2] * Mm[3 * q + 2];
          }
          double Lc[9];
          if (!chol3(Dm, Lc)) { s_fail = 1; break; }
          // Y_i (9 x 3) = (Bv_i^T - Y_{i-1} M^T) L^-T ; c_i = L^-1 (b_i - M c_{i-1})
          double Yc[27];
          const double* Bv = a.Bv + 27 * (size_t)i;
          for (int r = 0; r < 9; r++) {
            const double y0 = Bv[r] - (Yp[3 * r] * Mm[0] + Yp[3 * r + 1] * Mm[1] + Yp[3 * r + 2] * Mm[2]);
            const double y1 = Bv[9 + r] - (Yp[3 * r] * Mm[3] + Yp[3 * r + 1] * Mm[4] + Yp[3 * r + 2] * Mm[5]);
            const double y2 = Bv[18 + r] - (Yp[3 * r] * Mm[6] + Yp[3 * r + 1] * Mm[7] + Yp[3 * r + 2] * Mm[8]);
            const double z0 = y0 / Lc[0];
            const double z1 = (y1 - z0 * Lc[3]) / Lc[4];
            const double z2 = (y2 - z0 * Lc[6] - z1 * Lc[7]) / Lc[8];
            Yc[3 * r] = z0; Yc[3 * r + 1] = z1; Yc[3 * r + 2] = z2;
          }
          const double* bvi = a.bv + 3 * (size_t)i;
          double rb[3], ck[3];
          for (int r = 0; r < 3; r++) rb[r] = bvi[r] - (Mm[3 * r] * cprev[0] + Mm[3 * r + 1] * cprev[1] + Mm[3 * r + 2] * cprev[2]);
          ck[0] = rb[0] / Lc[0];
          ck[1] = (rb[1] - Lc[3] * ck[0]) / Lc[4];
          ck[2] = (rb[2] - Lc[6] * ck[0] - Lc[7] * ck[1]) / Lc[8];
          for (int r = 0; r < 9; r++) {
            bbw[r] -= Yc[3 * r] * ck[0] + Yc[3 * r + 1] * ck[1] + Yc[3 * r + 2] * ck[2];
            for (int q = 0; q < 9; q++) Sbb[9 * r + q] -= Yc[3 * r] * Yc[3 * q] + Yc[3 * r + 1] * Yc[3 * q + 1] + Yc[3 * r + 2] * Yc[3 * q + 2];
          }
          for (int k = 0; k < 9; k++) { a.L[9 * (size_t)i + k] = Lc[k]; a.M[9 * (size_t)i + k] = Mm[k]; Lp[k] = Lc[k]; }
          for (int k = 0; k < 27; k++) { a.Y[27 * (size_t)i + k] = Yc[k]; Yp[k] = Yc[k]; }
          for (int k = 0; k < 3; k++) { a.cc[3 * (size_t)i + k] = ck[k]; cprev[k] = ck[k]; }
        }
        if (!s_fail) {   // 9 x 9 Cholesky of the border, solve, back-substitute the velocities
          double Lb[81];
          for (int k = 0; k < 81; k++) Lb[k] = 0;
          for (int j = 0; j < 9 && !s_fail; j++) {
            double d = Sbb[10 * j];
            for (int k = 0; k < j; k++) d -= Lb[9 * j + k] * Lb[9 * j + k];
            if (!(d > 0)) { s_fail = 1; break; }
            Lb[10 * j] = sqrt(d);
            for (int i = j + 1; i < 9; i++) {
              double s2 = Sbb[9 * i + j];
              for (int k = 0; k < j; k++) s2 -= Lb[9 * i + k] * Lb[9 * j + k];
              Lb[9 * i + j] = s2 / Lb[10 * j];
            }
          }
          if (!s_fail) {
            double xb[9];
            for (int i = 0; i < 9; i++) {
              double s2 = bbw[i];
              for (int k = 0; k < i; k++) s2 -= Lb[9 * i + k] * xb[k];
              xb[i] = s2 / Lb[10 * i];
            }
            for (int i = 8; i >= 0; i--) {
              double s2 = xb[i];
              for (int k = i + 1; k < 9; k++) s2 -= Lb[9 * k + i] * xb[k];
              xb[i] = s2 / Lb[10 * i];
            }
            for (int k = 0; k < 9; k++) a.x[3 * (size_t)N + k] = xb[k];
            double xn[3] = {0, 0, 0};
            for (int i = N - 1; i >= 0; i--) {
              const double* Lc = a.L + 9 * (size_t)i;
              const double* Yc = a.Y + 27 * (size_t)i;
              double r[3] = {a.cc[3 * (size_t)i], a.cc[3 * (size_t)i + 1], a.cc[3 * (size_t)i + 2]};
              for (int q = 0; q < 9; q++) { r[0] -= Yc[3 * q] * xb[q]; r[1] -= Yc[3 * q + 1] * xb[q]; r[2] -= Yc[3 * q + 2] * xb[q]; }
              if (i + 1 < N) {
                const double* Mn = a.M + 9 * (size_t)(i + 1);
                for (int q = 0; q < 3; q++) r[q] -= Mn[q] * xn[0] + Mn[3 + q] * xn[1] + Mn[6 + q] * xn[2];
              }
              xn[2] = r[2] / Lc[8];
              xn[1] = (r[1] - Lc[7] * xn[2]) / Lc[4];
              xn[0] = (r[0] - Lc[3] * xn[1] - Lc[6] * xn[2]) / Lc[0];
              a.x[3 * (size_t)i] = xn[0]; a.x[3 * (size_t)i + 1] = xn[1]; a.x[3 * (size_t)i + 2] = xn[2];
            }
          }
        }
        // a failed factorisation leaves x as it was (linear_solver_eigen.h:94-125)
      }
      __syncthreads();
      // ---- update into the trial buffer, x^T (lambda x + b), chi2 at the trial state
      double* Vt = a.V + (size_t)(cur ^ 1) * 3 * N;
      double* gt = a.glob + 16 * (cur ^ 1);
      double sc = 0;
      for (int i = tid; i < 3 * N; i += IN_THREADS) {
        const double xi = a.mode == 0 ? a.x[i] : 0.0;
        Vt[i] = V[i] + xi;
        sc += xi * (lambda * xi + a.bv[i]);
      }
      if (tid == 0) {
        const double* xb = a.x + 3 * (size_t)N;
        for (int k = 0; k < 6; k++) gt[k] = glob[k] + (a.mode == 0 ? xb[k] : 0.0);
        double Ex[9];
        exp_so3(xb[6], xb[7], 0.0, Ex);
        mul33(glob + 6, Ex, gt + 6);
        gt[15] = glob[15] * exp(xb[8]);
        for (int k = (a.mode == 0 ? 0 : 6); k < 9; k++) sc += xb[k] * (lambda * xb[k] + bb[k]);
      }
      const double scale = block_sum(sc, sm);
      const double chi = chi2_of(a, cur ^ 1, sm);
      if (tid == 0) lm_trial(c, chi, scale, s_fail);
      __syncthreads();
      if (!lm_more_trials(c)) break;
    }
    if (tid == 0) lm_end_iteration(c, it, -1.0, a.rec);
    __syncthreads();
  }
}

// cyclic Jacobi eigen-decomposition (host), A = V diag(w) V^T
void jacobi9(double* A, double* V, double* w) {
  const int n = 9;
  for (int i = 0; i < 81; i++) V[i] = (i % 10 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
    if (off == 0) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (apq == 0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double cs = 1 / sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = cs * akp - sn * akq; A[k * n + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = cs * apk - sn * aqk; A[q * n + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = cs * vkp - sn * vkq; V[k * n + q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

// EdgeInertialGS ctor (src/G2oTypes.cc:363-375): float32 SVD inverse of C[0:9,0:9], symmetrised, eigenvalues < 1e-12 zeroed
void edge_information(const float* C15, double* Info) {
  double A[81], V[81], w[9];
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) A[9 * r + c] = 0.5 * ((double)C15[15 * r + c] + (double)C15[15 * c + r]);
  jacobi9(A, V, w);
  double sum = 0;
  for (int i = 0; i < 9; i++) sum += fabs(w[i]);
  const double thr = (double)FLT_EPSILON * 2 * sum;
  float invf[81];
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++)
        if (fabs(w[k]) > thr) s += V[9 * r + k] * V[9 * c + k] / w[k];
      invf[9 * r + c] = (float)s;
    }
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) A[9 * r + c] = ((double)invf[9 * r + c] + (double)invf[9 * c + r]) / 2;
  jacobi9(A, V, w);
  for (int i = 0; i < 9; i++)
    if (w[i] < 1e-12) w[i] = 0;
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++) s += V[9 * r + k] * w[k] * V[9 * c + k];
      Info[9 * r + c] = s;
    }
}

}  // namespace

void inertial_default_params(vido_inertial_problem* p) { p->prior_g = 1e2f; p->prior_a = 1e9f; p->its = 200; p->mode = 0; }

int inertial_opt_host(vido_ctx* ctx, vido_inertial_problem* p, vido_lm_stats* st) {
  if (st) { st->iterations = -1; st->n_records = 0; st->total_trials = 0; }
  const int N = p->n_frames;
  if (N < 2) return VIDO_OK;
  cudaStream_t s = ctx->stream;
  const size_t E = (size_t)N - 1;
  // host staging: doubles of the float32 poses, information matrices, initial state
  std::vector<double> Rwb(9 * (size_t)N), twb(3 * (size_t)N), V(2 * 3 * (size_t)N), glob(32, 0.0), Info(81 * E);
  for (size_t i = 0; i < Rwb.size(); i++) Rwb[i] = p->Rwb[i];
  for (size_t i = 0; i < twb.size(); i++) { twb[i] = p->twb[i]; V[i] = p->velocity[i]; }
  for (int k = 0; k < 3; k++) { glob[k] = p->bg[k]; glob[3 + k] = p->ba[k]; }
  for (int k = 0; k < 9; k++) glob[6 + k] = p->Rwg[k];
  glob[15] = p->scale;
  for (size_t e = 0; e < E; e++) edge_information(p->preint[e].C, &Info[81 * e]);
  // one allocation, carved
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += al(bytes); return o; };
  const size_t o_Rwb = take(8 * 9 * N), o_twb = take(8 * 3 * N), o_pre = take(sizeof(vido_imu_preint) * E), o_blin = take(4 * 6 * E),
               o_Info = take(8 * 81 * E), o_V = take(8 * 6 * N), o_glob = take(8 * 32), o_He = take(8 * 225 * E), o_be = take(8 * 15 * E),
               o_D = take(8 * 9 * N), o_O = take(8 * 9 * N), o_Bv = take(8 * 27 * N), o_bv = take(8 * 3 * N), o_x = take(8 * (3 * N + 9)),
               o_L = take(8 * 9 * N), o_M = take(8 * 9 * N), o_Y = take(8 * 27 * N), o_cc = take(8 * 3 * N), o_ctl = take(sizeof(LmCtl)),
               o_rec = take(sizeof(LmRec) * VIDO_LM_REC);
  char* base = (char*)vido_scratch(ctx, 0, off);
  if (!base) { ctx->err = "inertial: device allocation failed"; return VIDO_ERR_CUDA; }
  int rc = VIDO_OK;
  do {
#define CP(o, src, bytes) if (cudaMemcpyAsync(base + (o), src, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { ctx->err = "inertial: upload failed"; rc = VIDO_ERR_CUDA; break; }
    CP(o_Rwb, Rwb.data(), 8 * 9 * (size_t)N); CP(o_twb, twb.data(), 8 * 3 * (size_t)N); CP(o_pre, p->preint, sizeof(vido_imu_preint) * E);
    CP(o_blin, p->bias_lin, 4 * 6 * E); CP(o_Info, Info.data(), 8 * 81 * E); CP(o_V, V.data(), 8 * 6 * (size_t)N); CP(o_glob, glob.data(), 8 * 32);
#undef CP
    if (cudaMemsetAsync(base + o_x, 0, 8 * (3 * (size_t)N + 9), s) != cudaSuccess) { ctx->err = "inertial: memset failed"; rc = VIDO_ERR_CUDA; break; }
    InertialArgs a;
    memset(&a, 0, sizeof a);
    a.N = N; a.its = p->its; a.mode = p->mode;
    a.Rwb = (const double*)(base + o_Rwb); a.twb = (const double*)(base + o_twb); a.pre = (const vido_imu_preint*)(base + o_pre);
    a.blin = (const float*)(base + o_blin); a.Info = (const double*)(base + o_Info);
    a.priorG = (double)p->prior_g; a.priorA = (double)p->prior_a; a.user_lambda = (p->mode == 0 && p->prior_g != 0.f) ? 1e3 : -1.0;   // :2456-2458
    a.V = (double*)(base + o_V); a.glob = (double*)(base + o_glob);
    a.He = (double*)(base + o_He); a.be = (double*)(base + o_be); a.D = (double*)(base + o_D); a.O = (double*)(base + o_O);
    a.Bv = (double*)(base + o_Bv); a.bv = (double*)(base + o_bv); a.x = (double*)(base + o_x); a.L = (double*)(base + o_L);
    a.M = (double*)(base + o_M); a.Y = (double*)(base + o_Y); a.cc = (double*)(base + o_cc);
    a.ctl = (LmCtl*)(base + o_ctl); a.rec = (LmRec*)(base + o_rec);
    inertial_opt_kernel<<<1, IN_THREADS, 0, s>>>(a);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) { ctx->err = "inertial: launch failed"; rc = VIDO_ERR_CUDA; break; }
    LmCtl c;
    std::vector<LmRec> rec(VIDO_LM_REC);
    if (cudaMemcpyAsync(&c, base + o_ctl, sizeof c, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(rec.data(), base + o_rec, sizeof(LmRec) * VIDO_LM_REC, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(V.data(), base + o_V, 8 * 6 * (size_t)N, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaMemcpyAsync(glob.data(), base + o_glob, 8 * 32, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { ctx->err = "inertial: solve failed"; rc = VIDO_ERR_CUDA; break; }
    const double* Vc = V.data() + (size_t)c.cur * 3 * N;
    const double* gc = glob.data() + 16 * c.cur;
    for (size_t i = 0; i < 3 * (size_t)N; i++) p->velocity[i] = (float)Vc[i];
    for (int k = 0; k < 3; k++) { p->bg[k] = gc[k]; p->ba[k] = gc[3 + k]; }
    for (int k = 0; k < 9; k++) p->Rwg[k] = gc[6 + k];
    p->scale = gc[15];
    if (st) {
      st->iterations = c.iterations; st->n_records = c.n_records; st->total_trials = c.total_trials;
      for (int k = 0; k < c.n_records && k < VIDO_LM_MAX_RECORDS; k++) { st->rec[k].chi2 = rec[k].chi2; st->rec[k].lambda = rec[k].lambda; st->rec[k].trials = rec[k].trials; }
    }
  } while (0);
  cudaStreamSynchronize(s);   // (the scratch belongs to the context)
  return rc;
}
