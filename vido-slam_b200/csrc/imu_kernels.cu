// imu_kernels.cu -- IMU preintegration, one thread per (frame interval) job, batched over frames / sequences.
//
// Replaces Tracking::PreintegrateIMU (src/Tracking.cc:784-887: queue selection, end-point interpolation, mid-point
// rule) and IMU::Preintegrated::IntegrateNewMeasurement (src/ImuTypes.cc:245-300) with IMU::IntegratedRotation
// (:143-168).  State is float32 like the reference's cv::Mat fields; every matrix expression is evaluated in double
// and rounded once (cv::gemm on CV_32F accumulates in double).  NormalizeRotation's SVD (U*Vt) is replaced by the
// orthogonal polar factor, which is the same matrix.  The recurrence is sequential (~20 steps per frame at 200 Hz /
// 10 fps), so the parallelism is across jobs: this stage is latency-bound, not roofline-meaningful.
#include <cstring>

#include "ctx.h"

struct ImuState {
  float dT, dR[9], dV[3], dP[3], JRg[9], JVg[9], JVa[9], JPg[9], JPa[9], avgA[3], avgW[3];
};

__device__ __forceinline__ void mm33(const double* a, const double* b, double* o) {
  double t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
#pragma unroll
  for (int k = 0; k < 9; k++) o[k] = t[k];
}

__device__ __forceinline__ void inv33d(const double* m, double* o) {
  const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  const double id = 1.0 / det;
  o[0] = (m[4] * m[8] - m[5] * m[7]) * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = (m[5] * m[6] - m[3] * m[8]) * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = (m[3] * m[7] - m[4] * m[6]) * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

__device__ void imu_step(ImuState& s, float* C /* 225, global */, const float* bias, const float* Nga, const float* NgaWalk,
                         const float* a3, const float* w3, float dt) {
  const double acc[3] = {(double)(a3[0] - bias[0]), (double)(a3[1] - bias[1]), (double)(a3[2] - bias[2])};
  const double accW[3] = {(double)(w3[0] - bias[3]), (double)(w3[1] - bias[4]), (double)(w3[2] - bias[5])};
  double dR[9], JRg[9], RW[9], RWJ[9];
  for (int k = 0; k < 9; k++) { dR[k] = s.dR[k]; JRg[k] = s.JRg[k]; }
  const double t = dt, T = s.dT;
  double Ra[3];
  for (int i = 0; i < 3; i++) Ra[i] = dR[3 * i] * acc[0] + dR[3 * i + 1] * acc[1] + dR[3 * i + 2] * acc[2];
  for (int i = 0; i < 3; i++) {
    s.avgA[i] = (float)((T * (double)s.avgA[i] + Ra[i] * t) / (T + t));
    s.avgW[i] = (float)((T * (double)s.avgW[i] + accW[i] * t) / (T + t));
    const double dV = s.dV[i], dP = s.dP[i];
    s.dP[i] = (float)(dP + dV * t + 0.5 * Ra[i] * t * t);
    s.dV[i] = (float)(dV + Ra[i] * t);
  }
  const double Wacc[9] = {0, -acc[2], acc[1], acc[2], 0, -acc[0], -acc[1], acc[0], 0};
  mm33(dR, Wacc, RW);
  mm33(RW, JRg, RWJ);
  double A[81], B[54];
  for (int k = 0; k < 81; k++) A[k] = 0;
  for (int k = 0; k < 54; k++) B[k] = 0;
  for (int i = 0; i < 9; i++) A[10 * i] = 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[9 * (3 + i) + j] = (float)(-RW[3 * i + j] * t);
      A[9 * (6 + i) + j] = (float)(-0.5 * RW[3 * i + j] * t * t);
      A[9 * (6 + i) + 3 + j] = (i == j) ? (double)dt : 0.0;
      B[6 * (3 + i) + 3 + j] = (float)(dR[3 * i + j] * t);
      B[6 * (6 + i) + 3 + j] = (float)(0.5 * dR[3 * i + j] * t * t);
    }
  for (int k = 0; k < 9; k++) {
    const double JPa = s.JPa[k], JVa = s.JVa[k], JPg = s.JPg[k], JVg = s.JVg[k];
    s.JPa[k] = (float)(JPa + JVa * t - 0.5 * dR[k] * t * t);
    s.JPg[k] = (float)(JPg + JVg * t - 0.5 * RWJ[k] * t * t);
    s.JVa[k] = (float)(JVa - dR[k] * t);
    s.JVg[k] = (float)(JVg - RWJ[k] * t);
  }
  const float x = (w3[0] - bias[3]) * dt, y = (w3[1] - bias[4]) * dt, z = (w3[2] - bias[5]) * dt;
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  const float d = __fsqrt_rn(d2);
  const double W[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  double W2[9], dRi[9], rJ[9];
  mm33(W, W, W2);
  if (d < 1e-4f) {
    for (int k = 0; k < 9; k++) { const double I = (k % 4 == 0) ? 1.0 : 0.0; dRi[k] = (float)(I + W[k]); rJ[k] = I; }
  } else {
    const double sd = sin((double)d), cd = cos((double)d);
    for (int k = 0; k < 9; k++) {
      const double I = (k % 4 == 0) ? 1.0 : 0.0;
      dRi[k] = (float)(I + W[k] * sd / d + W2[k] * (1.0 - cd) / d2);
      rJ[k] = (float)(I - W[k] * (1.0 - cd) / d2 + W2[k] * (d - sd) / ((double)d2 * d));
    }
  }
  double X[9];
  mm33(dR, dRi, X);
  for (int k = 0; k < 9; k++) X[k] = (float)X[k];
  for (int it = 0; it < 8; it++) {  // orthogonal polar factor
    double Xi[9];
    inv33d(X, Xi);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) X[3 * i + j] = 0.5 * (X[3 * i + j] + Xi[3 * j + i]);
  }
  for (int k = 0; k < 9; k++) s.dR[k] = (float)X[k];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      A[9 * i + j] = dRi[3 * j + i];
      B[6 * i + j] = (float)(rJ[3 * i + j] * t);
    }
  // covariance: C[0:9,0:9] = A C A^T + B Nga B^T, C[9:15,9:15] += NgaWalk
  double AC[81], NB[54];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      double v = 0;
      for (int k = 0; k < 9; k++) v += A[9 * i + k] * (double)C[15 * k + j];
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
      C[15 * i + j] = (float)((double)(float)v1 + (double)(float)v2);
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) C[15 * (9 + i) + 9 + j] = (float)((double)C[15 * (9 + i) + 9 + j] + (double)NgaWalk[6 * i + j]);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double v = 0;
      for (int k = 0; k < 3; k++) v += dRi[3 * k + i] * JRg[3 * k + j];
      s.JRg[3 * i + j] = (float)(v - rJ[3 * i + j] * t);
    }
  s.dT = (float)(T + t);
}

__global__ void imu_preint_kernel(const vido_imu_sample* __restrict__ q, int n_all, const int32_t* __restrict__ nvis,
                                  const double* __restrict__ t_prev, const double* __restrict__ t_cur, int njobs,
                                  const float* __restrict__ bias, float ng, float na, float ngw, float naw,
                                  vido_imu_preint* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  const int n = nvis ? min(nvis[j], n_all) : n_all;   // samples delivered before this frame (the queue the reference would see)
  const double tp = t_prev[j], tc = t_cur[j];
  vido_imu_preint* o = out + j;
  float Nga[36], NgaWalk[36];
  for (int k = 0; k < 36; k++) { Nga[k] = 0; NgaWalk[k] = 0; }
  for (int i = 0; i < 3; i++) {
    Nga[7 * i] = ng * ng; Nga[7 * (3 + i)] = na * na;
    NgaWalk[7 * i] = ngw * ngw; NgaWalk[7 * (3 + i)] = naw * naw;
  }
  ImuState s;
  memset(&s, 0, sizeof s);
  s.dR[0] = s.dR[4] = s.dR[8] = 1.f;
  for (int k = 0; k < 225; k++) o->C[k] = 0.f;
  // queue selection: first sample with t >= t_prev - 1 ms ... first sample with t >= t_cur - 1 ms (inclusive)
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (q[mid].t < tp - 0.001) lo = mid + 1; else hi = mid; }
  const int i0 = lo;
  int i1 = i0;
  while (i1 < n && q[i1].t < tc - 0.001) i1++;
  const int last = (i1 < n) ? i1 : n - 1;  // index of the last selected sample
  const int m = last - i0;                 // number of integration steps
  const float* b = bias + 6 * j;
  for (int i = 0; i < m; i++) {
    const vido_imu_sample s0 = q[i0 + i], s1 = q[i0 + i + 1];
    const float a0[3] = {s0.ax, s0.ay, s0.az}, a1[3] = {s1.ax, s1.ay, s1.az};
    const float w0[3] = {s0.wx, s0.wy, s0.wz}, w1[3] = {s1.wx, s1.wy, s1.wz};
    float acc[3], ang[3], tstep;
    if (i == 0 && i < m - 1) {
      const float tab = (float)(s1.t - s0.t), tini = (float)(s0.t - tp);
      const float r = __fdiv_rn(tini, tab);
      for (int k = 0; k < 3; k++) {
        acc[k] = __fmul_rn(__fsub_rn(__fadd_rn(a0[k], a1[k]), __fmul_rn(__fsub_rn(a1[k], a0[k]), r)), 0.5f);
        ang[k] = __fmul_rn(__fsub_rn(__fadd_rn(w0[k], w1[k]), __fmul_rn(__fsub_rn(w1[k], w0[k]), r)), 0.5f);
      }
      tstep = (float)(s1.t - tp);
    } else if (i < m - 1) {
      for (int k = 0; k < 3; k++) { acc[k] = __fmul_rn(__fadd_rn(a0[k], a1[k]), 0.5f); ang[k] = __fmul_rn(__fadd_rn(w0[k], w1[k]), 0.5f); }
      tstep = (float)(s1.t - s0.t);
    } else if (i > 0) {
      const float tab = (float)(s1.t - s0.t), tend = (float)(s1.t - tc);
      const float r = __fdiv_rn(tend, tab);
      for (int k = 0; k < 3; k++) {
        acc[k] = __fmul_rn(__fsub_rn(__fadd_rn(a0[k], a1[k]), __fmul_rn(__fsub_rn(a1[k], a0[k]), r)), 0.5f);
        ang[k] = __fmul_rn(__fsub_rn(__fadd_rn(w0[k], w1[k]), __fmul_rn(__fsub_rn(w1[k], w0[k]), r)), 0.5f);
      }
      tstep = (float)(tc - s0.t);
    } else {
      for (int k = 0; k < 3; k++) { acc[k] = a0[k]; ang[k] = w0[k]; }
      tstep = (float)(tc - tp);
    }
    imu_step(s, o->C, b, Nga, NgaWalk, acc, ang, tstep);
  }
  o->dT = s.dT;
  for (int k = 0; k < 9; k++) { o->dR[k] = s.dR[k]; o->JRg[k] = s.JRg[k]; o->JVg[k] = s.JVg[k]; o->JVa[k] = s.JVa[k]; o->JPg[k] = s.JPg[k]; o->JPa[k] = s.JPa[k]; }
  for (int k = 0; k < 3; k++) { o->dV[k] = s.dV[k]; o->dP[k] = s.dP[k]; o->avgA[k] = s.avgA[k]; o->avgW[k] = s.avgW[k]; }
  o->n_steps = m > 0 ? m : 0;
  o->n_consumed = (i1 < n) ? i1 : n;
}

int imu_preintegrate_host(vido_ctx* ctx, const vido_imu_sample* samples, int n, const double* t_prev, const double* t_cur,
                          int njobs, const float* bias, const float* noise, vido_imu_preint* out, const int32_t* nvis) {
  if (n < 0 || njobs < 1) return VIDO_ERR_ARG;
  cudaStream_t s = ctx->stream;
  // one carved block of the context's grow-only scratch (stream-ordered allocations come from a pool that hands its memory back to
  // the OS at every synchronisation by default: five of them per call made this entry point erratic, 0.1 .. 20 ms)
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_s = 0, o_t = o_s + al(sizeof(vido_imu_sample) * (size_t)std::max(n, 1)), o_b = o_t + al(sizeof(double) * 2 * njobs),
               o_o = o_b + al(sizeof(float) * 6 * njobs), o_n = o_o + al(sizeof(vido_imu_preint) * njobs), total = o_n + al(sizeof(int32_t) * njobs);
  char* base = (char*)vido_scratch(ctx, 2, total);
  if (!base) { ctx->err = "imu preintegration: device allocation failed"; return VIDO_ERR_CUDA; }
  vido_imu_sample* d_s = (vido_imu_sample*)(base + o_s); double* d_t = (double*)(base + o_t); float* d_b = (float*)(base + o_b);
  vido_imu_preint* d_o = (vido_imu_preint*)(base + o_o); int32_t* d_n = nullptr;
  if (nvis) {
    d_n = (int32_t*)(base + o_n);
    VIDO_CUDA(cudaMemcpyAsync(d_n, nvis, sizeof(int32_t) * njobs, cudaMemcpyHostToDevice, s));
  }
  if (n) VIDO_CUDA(cudaMemcpyAsync(d_s, samples, sizeof(vido_imu_sample) * n, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(d_t, t_prev, sizeof(double) * njobs, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(d_t + njobs, t_cur, sizeof(double) * njobs, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(d_b, bias, sizeof(float) * 6 * njobs, cudaMemcpyHostToDevice, s));
  imu_preint_kernel<<<(njobs + 63) / 64, 64, 0, s>>>(d_s, n, d_n, d_t, d_t + njobs, njobs, d_b, noise[0], noise[1], noise[2], noise[3], d_o);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  VIDO_CUDA(cudaMemcpyAsync(out, d_o, sizeof(vido_imu_preint) * njobs, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  return VIDO_OK;
}
