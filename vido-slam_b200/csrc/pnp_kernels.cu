// pnp_kernels.cu -- initial camera / object model: PnP-RANSAC hypotheses scored in parallel + the constant-velocity
// alternative, winner = more inliers.
//
// Replaces Tracking::GetInitModelCam (src/Tracking.cc:1914-2028) and GetInitModelObj (:2030-2162).  The reference
// delegates the RANSAC to cv::solvePnPRansac(500 its, 0.4 px, 0.98, SOLVEPNP_P3P) of un-vendored OpenCV, whose RNG and
// minimal solver cannot be matched bit for bit; the deterministic variant specified in oracle/vido_oracle.h is
// implemented here: counter-based 4-point samples, Gauss-Newton minimal solve from the motion-model pose, OpenCV's
// adaptive iteration count replayed sequentially over the precomputed scores, refit on the consensus set.
// All `iters` hypotheses are evaluated concurrently (one warp each); the sequential "best so far / shrinking
// iteration budget" logic is replayed by one thread, which gives exactly the sequential algorithm's answer.
#include <cstring>

#include "ctx.h"

#define PNP_THREADS 256

struct PnpPose {
  double R[9], t[3];
};

struct PnpArgs {
  int n, M, iters;
  const float* cur_xy;   // [n][2]
  const float* pts3d;    // [n][3]
  const int* good;       // [M] indices with valid depth
  const float* Tcw_motion;  // [16]
  float fx, fy, cx, cy, thr, confidence;
  // scratch
  PnpPose* hyp;   // [iters]
  int* hyp_cnt;   // [iters] (-1: degenerate sample)
  // outputs
  float* Tcw_out;   // [16]
  int* inlier_ids;  // [n]
  int* result;      // n_inliers, winner, ransac_inliers, mm_inliers
};

__device__ __forceinline__ unsigned long long splitmix(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ void quat_to_R9(const double* q, double* R) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}

// residual and Jacobian rows of one correspondence; returns false when the point is behind the camera
__device__ __forceinline__ bool pnp_row(const PnpPose& T, const float* X3, const float* uv, double fx, double fy, double cx,
                                        double cy, double* r, double* J0, double* J1) {
  const double X[3] = {X3[0], X3[1], X3[2]};
  double p[3];
#pragma unroll
  for (int k = 0; k < 3; k++) p[k] = T.R[3 * k] * X[0] + T.R[3 * k + 1] * X[1] + T.R[3 * k + 2] * X[2] + T.t[k];
  if (!(p[2] > 1e-9)) return false;
  const double iz = 1.0 / p[2], x = p[0] * iz, y = p[1] * iz;
  r[0] = fx * x + cx - uv[0];
  r[1] = fy * y + cy - uv[1];
  if (J0) {
    const double a00 = fx * iz, a02 = -fx * x * iz, a11 = fy * iz, a12 = -fy * y * iz;
    J0[0] = a02 * p[1]; J0[1] = a00 * p[2] - a02 * p[0]; J0[2] = -a00 * p[1]; J0[3] = a00; J0[4] = 0; J0[5] = a02;
    J1[0] = -a11 * p[2] + a12 * p[1]; J1[1] = -a12 * p[0]; J1[2] = a11 * p[0]; J1[3] = 0; J1[4] = a11; J1[5] = a12;
  }
  return true;
}

// solve H d = b (6x6, LDL^T) and apply the left-multiplicative update; false if H is not positive definite
__device__ bool pnp_step(const double* H, const double* b, PnpPose& T) {
  double L[36], D[6], y[6], d[6];
  for (int k = 0; k < 36; k++) L[k] = 0;
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
  const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= nq;
  double dR[9], Rn[9], tn[3];
  quat_to_R9(q, dR);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Rn[3 * r + c] = dR[3 * r] * T.R[c] + dR[3 * r + 1] * T.R[3 + c] + dR[3 * r + 2] * T.R[6 + c];
    tn[r] = dR[3 * r] * T.t[0] + dR[3 * r + 1] * T.t[1] + dR[3 * r + 2] * T.t[2] + d[3 + r];
  }
  for (int k = 0; k < 9; k++) T.R[k] = Rn[k];
  for (int k = 0; k < 3; k++) T.t[k] = tn[k];
  return true;
}

__device__ __forceinline__ bool pnp_inlier(const PnpPose& T, const float* X3, const float* uv, double fx, double fy, double cx,
                                           double cy, double thr) {
  double r[2];
  if (!pnp_row(T, X3, uv, fx, fy, cx, cy, r, nullptr, nullptr)) return false;
  return sqrt(r[0] * r[0] + r[1] * r[1]) < thr;
}

// ---- kernel 1: one warp per hypothesis
__global__ void __launch_bounds__(PNP_THREADS) pnp_hypotheses_kernel(const PnpArgs* __restrict__ ap) {
  const PnpArgs a = *ap;
  const int lane = threadIdx.x & 31;
  const int iter = blockIdx.x * (PNP_THREADS / 32) + (threadIdx.x >> 5);
  if (iter >= a.iters || a.M < 4) return;
  const double fx = a.fx, fy = a.fy, cx = a.cx, cy = a.cy;
  PnpPose T;
  int ok = 1;
  if (lane == 0) {
    int s[4];
    unsigned long long c = (unsigned long long)iter << 8;
    for (int j = 0; j < 4; j++) {
      while (true) {
        int v = (int)(splitmix(c++) % (unsigned long long)a.M);
        bool dup = false;
        for (int k = 0; k < j; k++) dup |= (s[k] == v);
        if (!dup) { s[j] = v; break; }
      }
    }
    const float* Tm = a.Tcw_motion;
    for (int r = 0; r < 3; r++) { for (int c2 = 0; c2 < 3; c2++) T.R[3 * r + c2] = Tm[4 * r + c2]; T.t[r] = Tm[4 * r + 3]; }
    for (int it = 0; it < 6 && ok; it++) {
      double H[36], b[6];
      for (int k = 0; k < 36; k++) H[k] = 0;
      for (int k = 0; k < 6; k++) b[k] = 0;
      for (int k = 0; k < 4 && ok; k++) {
        const int i = a.good[s[k]];
        double r[2], J0[6], J1[6];
        if (!pnp_row(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, r, J0, J1)) { ok = 0; break; }
        for (int p = 0; p < 6; p++) {
          b[p] -= J0[p] * r[0] + J1[p] * r[1];
          for (int q = 0; q < 6; q++) H[6 * p + q] += J0[p] * J0[q] + J1[p] * J1[q];
        }
      }
      if (ok && !pnp_step(H, b, T)) ok = 0;
    }
    a.hyp[iter] = T;
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok) {
    if (lane == 0) a.hyp_cnt[iter] = -1;
    return;
  }
#pragma unroll
  for (int k = 0; k < 9; k++) T.R[k] = __shfl_sync(0xffffffffu, T.R[k], 0);
#pragma unroll
  for (int k = 0; k < 3; k++) T.t[k] = __shfl_sync(0xffffffffu, T.t[k], 0);
  int cnt = 0;
  for (int k = lane; k < a.M; k += 32) {
    const int i = a.good[k];
    cnt += pnp_inlier(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, (double)a.thr) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) a.hyp_cnt[iter] = cnt;
}

__device__ int ransac_update_iters(double p, double ep, int modelPoints, int maxIters) {  // cv::RANSACUpdateNumIters
  p = fmin(fmax(p, 0.), 1.);
  ep = fmin(fmax(ep, 0.), 1.);
  double num = fmax(1. - p, 2.220446049250313e-16);
  double denom = 1. - pow(1. - ep, (double)modelPoints);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= maxIters * (-denom)) ? maxIters : (int)llrint(num / denom);
}

// ---- kernel 2: replay the sequential selection, consensus set, refit, motion model, winner (one CTA)
__global__ void __launch_bounds__(PNP_THREADS) pnp_select_kernel(const PnpArgs* __restrict__ ap, int* __restrict__ tmp_ids) {
  const PnpArgs a = *ap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double fx = a.fx, fy = a.fy, cx = a.cx, cy = a.cy;
  __shared__ int s_best, s_cnt, s_warp[PNP_THREADS / 32], s_base, s_ok, s_nr, s_nm;
  __shared__ PnpPose T;
  __shared__ double red[(PNP_THREADS / 32) * 27 + 32];
  if (tid == 0) {
    int best = -1, best_cnt = 0, niters = a.iters;
    if (a.M >= 4) {
      for (int iter = 0; iter < niters; iter++) {
        const int cnt = a.hyp_cnt[iter];
        if (cnt > max(best_cnt, 3)) {
          best_cnt = cnt;
          best = iter;
          niters = ransac_update_iters((double)a.confidence, (double)(a.M - cnt) / a.M, 4, niters);
        }
      }
    }
    s_best = best;
    if (best >= 0) T = a.hyp[best];
    s_nr = 0;
  }
  __syncthreads();
  int nr = 0;
  if (s_best >= 0) {
    // consensus set of the best hypothesis, ascending positions (ordered compaction into tmp_ids)
    s_base = 0;
    __syncthreads();
    for (int k0 = 0; k0 < a.M; k0 += PNP_THREADS) {
      const int k = k0 + tid;
      bool in = false;
      if (k < a.M) { const int i = a.good[k]; in = pnp_inlier(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, (double)a.thr); }
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (lane == 0) s_warp[warp] = __popc(m);
      __syncthreads();
      int off = s_base;
      for (int w = 0; w < warp; w++) off += s_warp[w];
      if (in) tmp_ids[off + __popc(m & ((1u << lane) - 1u))] = k;
      __syncthreads();
      if (tid == 0) { int t = 0; for (int w = 0; w < PNP_THREADS / 32; w++) t += s_warp[w]; s_base += t; }
      __syncthreads();
    }
    nr = s_base;
    // refit on the consensus set: 10 Gauss-Newton steps, block-reduced normal equations
    for (int it = 0; it < 10; it++) {
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; k++) acc[k] = 0;
      int bad = 0;
      for (int k = tid; k < nr; k += PNP_THREADS) {
        const int i = a.good[tmp_ids[k]];
        double r[2], J0[6], J1[6];
        if (!pnp_row(T, a.pts3d + 3 * i, a.cur_xy + 2 * i, fx, fy, cx, cy, r, J0, J1)) { bad = 1; continue; }
        int idx = 0;
#pragma unroll
        for (int p = 0; p < 6; p++) {
          acc[21 + p] -= J0[p] * r[0] + J1[p] * r[1];
#pragma unroll
          for (int q = p; q < 6; q++) acc[idx++] += J0[p] * J0[q] + J1[p] * J1[q];
        }
      }
      bad = __syncthreads_or(bad);
#pragma unroll
      for (int k = 0; k < 27; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      if (lane == 0)
        for (int k = 0; k < 27; k++) red[warp * 27 + k] = acc[k];
      __syncthreads();
      if (tid == 0) {
        double H[36], b[6];
        int idx = 0;
        for (int p = 0; p < 6; p++) {
          double s = 0;
          for (int w = 0; w < PNP_THREADS / 32; w++) s += red[w * 27 + 21 + p];
          b[p] = s;
          for (int q = p; q < 6; q++) {
            double h = 0;
            for (int w = 0; w < PNP_THREADS / 32; w++) h += red[w * 27 + idx];
            H[6 * p + q] = h; H[6 * q + p] = h;
            idx++;
          }
        }
        PnpPose Tn = T;
        s_ok = (!bad && pnp_step(H, b, Tn)) ? 1 : 0;
        if (s_ok) T = Tn;
      }
      __syncthreads();
      if (!s_ok) break;  // the refit is abandoned as a whole when a step fails (oracle: keep the unrefined model)
    }
  }
  // NOTE: a failed refit must fall back to the un-refitted hypothesis
  if (tid == 0 && s_best >= 0 && !s_ok) T = a.hyp[s_best];
  __syncthreads();
  // ---- constant-velocity model in float32 (src/Tracking.cc:1980-2001), ordered compaction into inlier_ids
  const float* Tm = a.Tcw_motion;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int k0 = 0; k0 < a.n; k0 += PNP_THREADS) {
    const int i = k0 + tid;
    bool in = false;
    if (i < a.n) {
      const float X[3] = {a.pts3d[3 * i], a.pts3d[3 * i + 1], a.pts3d[3 * i + 2]};
      float pc[3];
      for (int r = 0; r < 3; r++)
        pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tm[4 * r], X[0]), __fmul_rn(Tm[4 * r + 1], X[1])), __fmul_rn(Tm[4 * r + 2], X[2])), Tm[4 * r + 3]);
      const float invz = (float)(1.0 / (double)pc[2]);
      const float u = __fadd_rn(__fmul_rn(__fmul_rn(a.fx, pc[0]), invz), a.cx);
      const float v = __fadd_rn(__fmul_rn(__fmul_rn(a.fy, pc[1]), invz), a.cy);
      const float du = __fsub_rn(a.cur_xy[2 * i], u), dv = __fsub_rn(a.cur_xy[2 * i + 1], v);
      const float rpe = __fsqrt_rn(__fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv)));
      in = rpe < a.thr;
    }
    const unsigned m = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; w++) off += s_warp[w];
    if (in) a.inlier_ids[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < PNP_THREADS / 32; w++) t += s_warp[w]; s_base += t; }
    __syncthreads();
  }
  const int nm = s_base;
  // ---- winner (src/Tracking.cc:2006-2025)
  if (nr > nm) {
    for (int k = tid; k < nr; k += PNP_THREADS) a.inlier_ids[k] = tmp_ids[k];  // position inside the valid list, as the reference does
    if (tid == 0) {
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) a.Tcw_out[4 * r + c] = (float)T.R[3 * r + c];
        a.Tcw_out[4 * r + 3] = (float)T.t[r];
      }
      a.Tcw_out[12] = 0.f; a.Tcw_out[13] = 0.f; a.Tcw_out[14] = 0.f; a.Tcw_out[15] = 1.f;
      a.result[0] = nr; a.result[1] = 0;
    }
  } else if (tid == 0) {
    for (int k = 0; k < 16; k++) a.Tcw_out[k] = Tm[k];
    a.result[0] = nm; a.result[1] = 1;
  }
  if (tid == 0) { a.result[2] = nr; a.result[3] = nm; }
}

// =========================================================================================================
struct PnpWorkspace {
  int capN = 0, capIters = 0;
  PnpArgs* d_args = nullptr;
  float *cur, *pts, *Tm, *Tout;
  int *good, *ids, *tmp, *cnt, *result;
  PnpPose* hyp;
};

int pnp_setup(vido_ctx* ctx, int capN, int capIters) {
  PnpWorkspace* ws = new PnpWorkspace();
  ctx->pnp = ws;
  ws->capN = capN; ws->capIters = capIters;
  VIDO_CUDA(cudaMalloc(&ws->d_args, sizeof(PnpArgs)));
  VIDO_CUDA(cudaMalloc(&ws->cur, sizeof(float) * 2 * capN));
  VIDO_CUDA(cudaMalloc(&ws->pts, sizeof(float) * 3 * capN));
  VIDO_CUDA(cudaMalloc(&ws->Tm, sizeof(float) * 16));
  VIDO_CUDA(cudaMalloc(&ws->Tout, sizeof(float) * 16));
  VIDO_CUDA(cudaMalloc(&ws->good, sizeof(int) * capN));
  VIDO_CUDA(cudaMalloc(&ws->ids, sizeof(int) * capN));
  VIDO_CUDA(cudaMalloc(&ws->tmp, sizeof(int) * capN));
  VIDO_CUDA(cudaMalloc(&ws->cnt, sizeof(int) * capIters));
  VIDO_CUDA(cudaMalloc(&ws->result, sizeof(int) * 4));
  VIDO_CUDA(cudaMalloc(&ws->hyp, sizeof(PnpPose) * capIters));
  return VIDO_OK;
}

void pnp_teardown(vido_ctx* ctx) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  if (!ws) return;
  cudaFree(ws->d_args); cudaFree(ws->cur); cudaFree(ws->pts); cudaFree(ws->Tm); cudaFree(ws->Tout); cudaFree(ws->good);
  cudaFree(ws->ids); cudaFree(ws->tmp); cudaFree(ws->cnt); cudaFree(ws->result); cudaFree(ws->hyp);
  delete ws;
  ctx->pnp = nullptr;
}

int pnp_init_model_host(vido_ctx* ctx, vido_pnp_problem* p) {
  PnpWorkspace* ws = (PnpWorkspace*)ctx->pnp;
  if (p->n < 0 || p->n > ws->capN || p->iters > ws->capIters || p->iters < 1) { ctx->err = "PnP problem exceeds capacity"; return VIDO_ERR_CAPACITY; }
  cudaStream_t s = ctx->stream;
  std::vector<int> good;
  good.reserve(p->n);
  for (int i = 0; i < p->n; i++)
    if (!p->valid || p->valid[i]) good.push_back(i);
  PnpArgs a;
  memset(&a, 0, sizeof a);
  a.n = p->n; a.M = (int)good.size(); a.iters = p->iters;
  a.cur_xy = ws->cur; a.pts3d = ws->pts; a.good = ws->good; a.Tcw_motion = ws->Tm;
  a.fx = p->fx; a.fy = p->fy; a.cx = p->cx; a.cy = p->cy; a.thr = p->reproj_err; a.confidence = p->confidence;
  a.hyp = ws->hyp; a.hyp_cnt = ws->cnt; a.Tcw_out = ws->Tout; a.inlier_ids = ws->ids; a.result = ws->result;
  if (p->n) {
    VIDO_CUDA(cudaMemcpyAsync(ws->cur, p->cur_xy, sizeof(float) * 2 * p->n, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(ws->pts, p->pts3d, sizeof(float) * 3 * p->n, cudaMemcpyHostToDevice, s));
  }
  if (a.M) VIDO_CUDA(cudaMemcpyAsync(ws->good, good.data(), sizeof(int) * a.M, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(ws->Tm, p->Tcw_motion, sizeof(float) * 16, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(ws->d_args, &a, sizeof a, cudaMemcpyHostToDevice, s));
  cudaEventRecord(ctx->ev0, s);
  if (a.M >= 4) {
    pnp_hypotheses_kernel<<<(a.iters + PNP_THREADS / 32 - 1) / (PNP_THREADS / 32), PNP_THREADS, 0, s>>>(ws->d_args);
    ctx->launches++;
  }
  pnp_select_kernel<<<1, PNP_THREADS, 0, s>>>(ws->d_args, ws->tmp);
  cudaEventRecord(ctx->ev1, s);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  int res[4];
  VIDO_CUDA(cudaMemcpyAsync(res, ws->result, sizeof res, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaMemcpyAsync(p->Tcw_out, ws->Tout, sizeof(float) * 16, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) { ctx->t_ms[1] += ms; ctx->t_n[1]++; }
  }
  p->n_inliers = res[0]; p->winner = res[1]; p->ransac_inliers = res[2]; p->mm_inliers = res[3];
  if (res[0] > 0 && p->inlier_ids) {
    VIDO_CUDA(cudaMemcpyAsync(p->inlier_ids, ws->ids, sizeof(int) * res[0], cudaMemcpyDeviceToHost, s));
    VIDO_CUDA(cudaStreamSynchronize(s));
  }
  return VIDO_OK;
}
