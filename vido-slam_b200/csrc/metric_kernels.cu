// metric_kernels.cu -- Tracking::GetMetricError (src/Tracking.cc:3531-3674, bRMSError = false) as a device-side evaluation:
// one thread per relative camera pose / per estimated object motion computes the translation and rotation error with the
// reference's float32 cv::Mat arithmetic (4x4 products accumulated in double and rounded once, Converter::toInvMatrix
// inverses, the reference's clamped trace rule); the per-item errors come back and are averaged in the reference's order
// (sequential float32 sums).  Step after the hot path: no reference kernel exists.
#include <cmath>
#include <vector>

#include "ctx.h"

namespace {

__device__ void mmul44(const float* A, const float* B, float* C) {
  float o[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += (double)A[4 * r + k] * (double)B[4 * k + c];
      o[4 * r + c] = (float)s;
    }
  for (int k = 0; k < 16; k++) C[k] = o[k];
}
__device__ void minv44(const float* T, float* Ti) {  // Converter::toInvMatrix (src/Converter.cc:155-170)
  float o[16];
  for (int k = 0; k < 16; k++) o[k] = 0.f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[4 * r + c] = T[4 * c + r];
  for (int r = 0; r < 3; r++) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)(-o[4 * r + k]) * (double)T[4 * k + 3];
    o[4 * r + 3] = (float)s;
  }
  o[15] = 1.f;
  for (int k = 0; k < 16; k++) Ti[k] = o[k];
}
__device__ void pose_error(const float* E, float* t_err, float* r_err) {
  *t_err = sqrtf(E[3] * E[3] + E[7] * E[7] + E[11] * E[11]);
  float tr = 0;
  for (int j = 0; j < 3; j++) {
    const float d = E[5 * j];
    if (d > 1.0) tr = (float)((double)tr + 1.0 - ((double)d - 1.0));
    else tr = tr + d;
  }
  *r_err = (float)(acos(((double)tr - 1.0) / 2.0) * 180.0 / 3.1415926);
}

// items [0, n_cam): camera frame pairs (i = item + 1); items [n_cam, n_cam + n_obj): object motions
__global__ void metric_kernel(const float* __restrict__ cam, const float* __restrict__ cam_gt, int n_cam,
                              const float* __restrict__ mot, const float* __restrict__ pose_pre, const float* __restrict__ mot_gt,
                              int n_obj, float* __restrict__ out /* [n_cam + n_obj][2] */) {
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= n_cam + n_obj) return;
  float A[16], B[16], Cm[16], E[16];
  if (it < n_cam) {
    const int i = it + 1;
    minv44(cam + 16 * (size_t)(i - 1), A);
    mmul44(cam + 16 * (size_t)i, A, B);                 // T_lc_inv = CamPose[i] * inv(CamPose[i-1])
    minv44(cam_gt + 16 * (size_t)i, A);
    mmul44(cam_gt + 16 * (size_t)(i - 1), A, Cm);       // T_lc_gt = CamPose_gt[i-1] * inv(CamPose_gt[i])
    mmul44(B, Cm, E);
  } else {
    const int o = it - n_cam;
    const float* P = pose_pre + 16 * (size_t)o;
    minv44(P, A);
    mmul44(A, mot + 16 * (size_t)o, B);
    mmul44(B, P, Cm);                                   // RigMotBody = inv(ObjPosePre) * RigMot * ObjPosePre
    minv44(Cm, A);
    mmul44(A, mot_gt + 16 * (size_t)o, E);              // rpe = inv(RigMotBody) * RigMot_gt
  }
  pose_error(E, out + 2 * (size_t)it, out + 2 * (size_t)it + 1);
}

}  // namespace

int metric_error_host(vido_ctx* ctx, const float* cam, const float* cam_gt, int n, const float* mot, const float* pose_pre,
                      const float* mot_gt, int n_obj, vido_metric* out, float* per_item) {
  const int n_cam = n > 1 ? n - 1 : 0, items = n_cam + n_obj;
  out->cam_t = out->cam_r = out->obj_t = out->obj_r = 0.f;
  out->n_cam = n_cam; out->n_obj = n_obj;
  if (items == 0) return VIDO_OK;
  cudaStream_t s = ctx->stream;
  float *d_cam = nullptr, *d_gt = nullptr, *d_mot = nullptr, *d_pre = nullptr, *d_mgt = nullptr, *d_out = nullptr;
  const size_t cb = sizeof(float) * 16 * (size_t)(n > 0 ? n : 1), ob = sizeof(float) * 16 * (size_t)(n_obj > 0 ? n_obj : 1);
  VIDO_CUDA(cudaMallocAsync(&d_cam, cb, s)); VIDO_CUDA(cudaMallocAsync(&d_gt, cb, s));
  VIDO_CUDA(cudaMallocAsync(&d_mot, ob, s)); VIDO_CUDA(cudaMallocAsync(&d_pre, ob, s)); VIDO_CUDA(cudaMallocAsync(&d_mgt, ob, s));
  VIDO_CUDA(cudaMallocAsync(&d_out, sizeof(float) * 2 * (size_t)items, s));
  if (n_cam) {
    VIDO_CUDA(cudaMemcpyAsync(d_cam, cam, sizeof(float) * 16 * (size_t)n, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(d_gt, cam_gt, sizeof(float) * 16 * (size_t)n, cudaMemcpyHostToDevice, s));
  }
  if (n_obj) {
    VIDO_CUDA(cudaMemcpyAsync(d_mot, mot, sizeof(float) * 16 * (size_t)n_obj, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(d_pre, pose_pre, sizeof(float) * 16 * (size_t)n_obj, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(d_mgt, mot_gt, sizeof(float) * 16 * (size_t)n_obj, cudaMemcpyHostToDevice, s));
  }
  metric_kernel<<<(items + 127) / 128, 128, 0, s>>>(d_cam, d_gt, n_cam, d_mot, d_pre, d_mgt, n_obj, d_out);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  std::vector<float> h(2 * (size_t)items);
  VIDO_CUDA(cudaMemcpyAsync(h.data(), d_out, sizeof(float) * h.size(), cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(d_cam, s); cudaFreeAsync(d_gt, s); cudaFreeAsync(d_mot, s); cudaFreeAsync(d_pre, s); cudaFreeAsync(d_mgt, s); cudaFreeAsync(d_out, s);
  // averages in the reference's order: sequential float32 sums (src/Tracking.cc:3541-3577, 3592-3650)
  float ts = 0, rs = 0;
  for (int i = 0; i < n_cam; i++) { ts = ts + h[2 * (size_t)i]; rs = rs + h[2 * (size_t)i + 1]; }
  if (n_cam) { out->cam_t = ts / (float)n_cam; out->cam_r = rs / (float)n_cam; }
  ts = 0; rs = 0;
  for (int i = n_cam; i < items; i++) { ts = ts + h[2 * (size_t)i]; rs = rs + h[2 * (size_t)i + 1]; }
  if (n_obj) { out->obj_t = ts / (float)n_obj; out->obj_r = rs / (float)n_obj; }
  if (per_item) memcpy(per_item, h.data(), sizeof(float) * h.size());
  return VIDO_OK;
}
