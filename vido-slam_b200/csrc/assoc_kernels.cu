// assoc_kernels.cu -- per-frame gather stages: depth pre-scale, flow-guided static association, stride-4 object
// sampling, generic (mask, depth, flow) gathers for the host-side bookkeeping.
//
// Replaces (paths under /root/reference/vido_slam/src):
//   depth pre-scale loop of Tracking::GrabImageRGBD      Tracking.cc:299-322
//   static association of Frame::Frame                   Frame.cc:72-100 (+ depth lookup :164-177)
//   semi-dense object sampling of Frame::Frame           Frame.cc:184-211
//   per-point map lookups of GrabImageRGBD / RenewFrameInfo  Tracking.cc:369-421, 2976-3010
// Output order is the reference's (keypoint order / raster order): ordered compaction with warp ballots.
// The depth conversion can be applied on the fly at the gather points (raw = 1), so the 3.7 MB read+write pass over
// the depth map is only needed when the caller wants the reference's in-place side effect.
#include <cstring>

#include <algorithm>
#include <vector>

#include "ctx.h"

struct DepthConv {
  int mode;  // 0: already converted, 1 OMD, 2 KITTI, 3 KAIST
  float factor, bf, mscale;
};

__device__ __forceinline__ float conv_depth(float d, const DepthConv& c) {
  if (c.mode == 0) return d;
  if (d < 0) return 0.f;
  if (c.mode == 1) return __fdiv_rn(d, c.factor);
  if (c.mode == 2) return __fdiv_rn(c.bf, __fdiv_rn(d, c.factor));
  return __fdiv_rn(__fmul_rn(c.mscale, c.bf), __fdiv_rn(d, c.factor));
}

__global__ void depth_prep_kernel(float* __restrict__ depth, int w, int h, int stride, size_t fs, DepthConv c) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  float* p = depth + b * fs + (size_t)y * stride + x;
  *p = conv_depth(*p, c);
}

#define ASSOC_THREADS 256

// block-wide ordered append: every thread passes a flag; returns the output slot (or -1) and advances *base
__device__ __forceinline__ int ordered_slot(bool flag, int* s_warp, int* s_base) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  int off = *s_base;
  for (int w = 0; w < warp; w++) off += s_warp[w];
  const int slot = flag ? off + __popc(m & ((1u << lane) - 1u)) : -1;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < ASSOC_THREADS / 32; w++) t += s_warp[w];
    *s_base += t;
  }
  __syncthreads();
  return slot;
}

struct AssocArgs {
  int w, h;
  const float* depth; const float* flow; const int32_t* mask;  // [B] images, tight rows
  size_t img_fs;                                               // elements per frame (w*h)
  DepthConv conv;
  float th_bg, th_obj;
};

// one CTA per frame: keypoints in order -> compacted static features
__global__ void __launch_bounds__(ASSOC_THREADS) frame_associate_kernel(AssocArgs a, const vido_keypoint* __restrict__ kps,
                                                                        const int32_t* __restrict__ nkp, int kp_cap,
                                                                        int32_t* __restrict__ out_idx, float* __restrict__ corres,
                                                                        float* __restrict__ oflow, float* __restrict__ odepth,
                                                                        int32_t* __restrict__ out_n, int out_cap) {
  __shared__ int s_warp[ASSOC_THREADS / 32], s_base;
  const int b = blockIdx.x;
  const int n = nkp[b];
  const vido_keypoint* kp = kps + (size_t)b * kp_cap;
  const float* depth = a.depth + b * a.img_fs;
  const float* flow = a.flow + 2 * b * a.img_fs;
  const int32_t* mask = a.mask + b * a.img_fs;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += ASSOC_THREADS) {
    const int i = i0 + threadIdx.x;
    bool ok = false;
    float px = 0, py = 0, fx = 0, fy = 0, d = 0;
    if (i < n) {
      px = kp[i].x; py = kp[i].y;
      const int x = (int)px, y = (int)py;
      const size_t k = (size_t)y * a.w + x;
      if (mask[k] == 0) {
        d = conv_depth(depth[k], a.conv);
        if (!(d > a.th_bg || d <= 0)) {
          fx = flow[2 * k]; fy = flow[2 * k + 1];
          if (fx != 0 && fy != 0)
            ok = (__fadd_rn(px, fx) < (float)a.w) && (__fadd_rn(py, fy) < (float)a.h) && (px < (float)a.w) && (py < (float)a.h);
        }
      }
    }
    const int slot = ordered_slot(ok, s_warp, &s_base);
    if (slot >= 0 && slot < out_cap) {
      const size_t o = (size_t)b * out_cap + slot;
      out_idx[o] = i;
      corres[2 * o] = __fadd_rn(px, fx); corres[2 * o + 1] = __fadd_rn(py, fy);
      oflow[2 * o] = fx; oflow[2 * o + 1] = fy;
      odepth[o] = d;
    }
  }
  if (threadIdx.x == 0) out_n[b] = s_base;
}

// one CTA per frame: raster scan with stride 4 over the mask
__global__ void __launch_bounds__(ASSOC_THREADS) sample_objects_kernel(AssocArgs a, float* __restrict__ keys, float* __restrict__ corres,
                                                                       float* __restrict__ oflow, float* __restrict__ odepth,
                                                                       int32_t* __restrict__ label, int32_t* __restrict__ out_n,
                                                                       int out_cap) {
  __shared__ int s_warp[ASSOC_THREADS / 32], s_base;
  const int b = blockIdx.x;
  const float* depth = a.depth + b * a.img_fs;
  const float* flow = a.flow + 2 * b * a.img_fs;
  const int32_t* mask = a.mask + b * a.img_fs;
  const int gw = (a.w + 3) / 4, gh = (a.h + 3) / 4, total = gw * gh;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int t0 = 0; t0 < total; t0 += ASSOC_THREADS) {
    const int t = t0 + threadIdx.x;
    bool ok = false;
    int i = 0, j = 0, lab = 0;
    float fx = 0, fy = 0, d = 0;
    if (t < total) {
      i = (t / gw) * 4; j = (t % gw) * 4;
      const size_t k = (size_t)i * a.w + j;
      lab = mask[k];
      if (lab != 0) {
        d = conv_depth(depth[k], a.conv);
        if (d < a.th_obj && d > 0) {
          fx = flow[2 * k]; fy = flow[2 * k + 1];
          const float cx = __fadd_rn((float)j, fx), cy = __fadd_rn((float)i, fy);
          ok = cx < (float)a.w && cx > 0 && cy < (float)a.h && cy > 0;
        }
      }
    }
    const int slot = ordered_slot(ok, s_warp, &s_base);
    if (slot >= 0 && slot < out_cap) {
      const size_t o = (size_t)b * out_cap + slot;
      keys[2 * o] = (float)j; keys[2 * o + 1] = (float)i;
      corres[2 * o] = __fadd_rn((float)j, fx); corres[2 * o + 1] = __fadd_rn((float)i, fy);
      oflow[2 * o] = fx; oflow[2 * o + 1] = fy;
      odepth[o] = d;
      label[o] = lab;
    }
  }
  if (threadIdx.x == 0) out_n[b] = s_base;
}

// (mask, depth, flow) at truncated coordinates; out-of-image queries return mask = -1, depth = 0, flow = 0
__global__ void gather_kernel(AssocArgs a, int frame, const float* __restrict__ xy, int n, int32_t* __restrict__ omask,
                              float* __restrict__ odepth, float* __restrict__ oflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)xy[2 * i], y = (int)xy[2 * i + 1];
  if (x < 0 || y < 0 || x >= a.w || y >= a.h) { omask[i] = -1; odepth[i] = 0.f; oflow[2 * i] = 0.f; oflow[2 * i + 1] = 0.f; return; }
  const size_t k = (size_t)frame * a.img_fs + (size_t)y * a.w + x;
  omask[i] = a.mask[k];
  odepth[i] = conv_depth(a.depth[k], a.conv);
  oflow[2 * i] = a.flow[2 * k];
  oflow[2 * i + 1] = a.flow[2 * k + 1];
}

// =========================================================================================================
static DepthConv make_conv(const vido_ctx* ctx, int raw) {
  DepthConv c;
  c.mode = raw ? ctx->cfg.choose_data : 0;
  c.factor = ctx->cfg.depth_map_factor;
  c.bf = ctx->cfg.bf;
  c.mscale = ctx->mscale;
  return c;
}

int assoc_depth_prep(vido_ctx* ctx, float* d_depth, int nframes, size_t frame_stride, int stride) {
  const vido_config& c = ctx->cfg;
  dim3 grid((c.width + 255) / 256, c.height, nframes);
  depth_prep_kernel<<<grid, 256, 0, ctx->stream>>>(d_depth, c.width, c.height, stride, frame_stride, make_conv(ctx, 1));
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

static AssocArgs make_args(const vido_ctx* ctx, const float* depth, const float* flow, const int32_t* mask, int raw) {
  AssocArgs a;
  a.w = ctx->cfg.width; a.h = ctx->cfg.height;
  a.depth = depth; a.flow = flow; a.mask = mask;
  a.img_fs = (size_t)a.w * a.h;
  a.conv = make_conv(ctx, raw);
  a.th_bg = ctx->cfg.th_depth_bg; a.th_obj = ctx->cfg.th_depth_obj;
  return a;
}

int assoc_frame_associate(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int kp_cap, const float* d_depth,
                          const float* d_flow, const int32_t* d_mask, int nframes, int raw, int32_t* d_idx, float* d_corres,
                          float* d_oflow, float* d_odepth, int32_t* d_n, int out_cap) {
  frame_associate_kernel<<<nframes, ASSOC_THREADS, 0, ctx->stream>>>(make_args(ctx, d_depth, d_flow, d_mask, raw), d_kps, d_nkp,
                                                                     kp_cap, d_idx, d_corres, d_oflow, d_odepth, d_n, out_cap);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

int assoc_sample_objects(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int nframes, int raw,
                         float* d_keys, float* d_corres, float* d_oflow, float* d_odepth, int32_t* d_label, int32_t* d_n, int out_cap) {
  sample_objects_kernel<<<nframes, ASSOC_THREADS, 0, ctx->stream>>>(make_args(ctx, d_depth, d_flow, d_mask, raw), d_keys, d_corres,
                                                                    d_oflow, d_odepth, d_label, d_n, out_cap);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

int assoc_gather(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int frame, int raw,
                 const float* d_xy, int n, int32_t* d_omask, float* d_odepth, float* d_oflow) {
  if (n <= 0) return VIDO_OK;
  gather_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(make_args(ctx, d_depth, d_flow, d_mask, raw), frame, d_xy, n, d_omask,
                                                          d_odepth, d_oflow);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

// =========================================================================================================
// Tracking::UpdateMask (src/Tracking.cc:3291-3357): per semantic label (ascending, one after the other on the mask as
// modified so far) a vote of the current mask at the predicted object-feature positions; a lost mask (label 0 wins with
// >= 100 votes; ties go to the smaller label) is forward-warped from the last frame through its flow.
// Three small kernels per label, all on the context stream; the decision stays on the device.
// =========================================================================================================
#define UM_BINS 4096  // labels must be in [0, UM_BINS)

__global__ void um_vote_kernel(const int32_t* __restrict__ sem, const float* __restrict__ cor, int n, int32_t label,
                               const int32_t* __restrict__ mask_cur, int W, int H, int32_t* __restrict__ hist /* [UM_BINS + 2] */,
                               int32_t* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || sem[i] != label) return;
  const int u = (int)cor[2 * i], v = (int)cor[2 * i + 1];
  if (u < W && u > 0 && v < H && v > 0) {
    const int32_t l = mask_cur[(size_t)v * W + u];
    if (l < 0 || l >= UM_BINS) { atomicExch(err, 2); return; }
    atomicAdd(&hist[l], 1);
    atomicAdd(&hist[UM_BINS], 1);  // number of votes
  }
}

// winner of the vote -> hist[UM_BINS + 1] = 1 when the mask has to be recovered; clears the histogram for the next label
__global__ void um_decide_kernel(int32_t* __restrict__ hist, int32_t* __restrict__ recovered_k) {
  __shared__ int s_cnt[256], s_lab[256];
  const int tid = threadIdx.x;
  int best = -1, lab = 0;
  for (int l = tid; l < UM_BINS; l += blockDim.x) {
    const int c = hist[l];
    if (c > best) { best = c; lab = l; }  // ascending l per thread: the smaller label is kept among equals
  }
  s_cnt[tid] = best; s_lab[tid] = lab;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      const int c2 = s_cnt[tid + o], l2 = s_lab[tid + o];
      if (c2 > s_cnt[tid] || (c2 == s_cnt[tid] && l2 < s_lab[tid])) { s_cnt[tid] = c2; s_lab[tid] = l2; }
    }
    __syncthreads();
  }
  const int votes = hist[UM_BINS];
  __syncthreads();
  for (int l = tid; l < UM_BINS; l += blockDim.x) hist[l] = 0;
  if (tid == 0) {
    const int rec = (votes >= 100 && s_lab[0] == 0 && s_cnt[0] > 0) ? 1 : 0;
    hist[UM_BINS] = 0;
    hist[UM_BINS + 1] = rec;
    *recovered_k = rec;
  }
}

__global__ void um_warp_kernel(const int32_t* __restrict__ flag, int32_t label, const int32_t* __restrict__ mask_last,
                               const float* __restrict__ flow_last, int32_t* __restrict__ mask_cur, int W, int H) {
  if (*flag == 0) return;
  const size_t px = (size_t)W * H;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < px; k += (size_t)gridDim.x * blockDim.x) {
    if (mask_last[k] != label) continue;
    const int j = (int)(k / W), kx = (int)(k - (size_t)j * W);
    const int fx = (int)flow_last[2 * k], fy = (int)flow_last[2 * k + 1];  // truncation, like the reference's int conversion
    if (kx + fx < W && kx + fx > 0 && j + fy < H && j + fy > 0) mask_cur[(size_t)(j + fy) * W + kx + fx] = label;  // same value from every writer
  }
}

int assoc_update_mask(vido_ctx* ctx, const int32_t* sem_label, const float* corres_xy, int n, const int32_t* d_mask_last,
                      const float* d_flow_last, int32_t* d_mask_cur, int32_t* uniq_out, int32_t* recovered, int cap) {
  if (n <= 0) return 0;
  const int W = ctx->cfg.width, H = ctx->cfg.height;
  cudaStream_t s = ctx->stream;
  std::vector<int32_t> uni(sem_label, sem_label + n);
  std::sort(uni.begin(), uni.end());
  uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
  const int nl = (int)uni.size();
  int32_t *d_sem = nullptr, *d_hist = nullptr, *d_rec = nullptr;
  float* d_cor = nullptr;
  {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_cor = al(sizeof(int32_t) * n), o_hist = o_cor + al(sizeof(float) * 2 * n),
                 o_rec = o_hist + al(sizeof(int32_t) * (UM_BINS + 2)), need = o_rec + al(sizeof(int32_t) * nl);
    if (need > ctx->um_ws_bytes) {
      VIDO_CUDA(cudaStreamSynchronize(s));
      if (ctx->um_ws) cudaFree(ctx->um_ws);
      ctx->um_ws = nullptr; ctx->um_ws_bytes = 0;
      VIDO_CUDA(cudaMalloc(&ctx->um_ws, 2 * need));
      ctx->um_ws_bytes = 2 * need;
    }
    d_sem = (int32_t*)ctx->um_ws; d_cor = (float*)(ctx->um_ws + o_cor); d_hist = (int32_t*)(ctx->um_ws + o_hist);
    d_rec = (int32_t*)(ctx->um_ws + o_rec);
  }
  VIDO_CUDA(cudaMemcpyAsync(d_sem, sem_label, sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(d_cor, corres_xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * (UM_BINS + 2), s));
  for (int k = 0; k < nl; k++) {
    um_vote_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_sem, d_cor, n, uni[k], d_mask_cur, W, H, d_hist, ctx->d_err);
    um_decide_kernel<<<1, 256, 0, s>>>(d_hist, d_rec + k);
    um_warp_kernel<<<148 * 4, 256, 0, s>>>(d_hist + UM_BINS + 1, uni[k], d_mask_last, d_flow_last, d_mask_cur, W, H);
    ctx->launches += 3;
  }
  VIDO_CUDA(cudaGetLastError());
  std::vector<int32_t> rec(nl);
  int32_t flag = 0;
  VIDO_CUDA(cudaMemcpyAsync(rec.data(), d_rec, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaMemcpyAsync(&flag, ctx->d_err, 4, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  if (flag) { cudaMemsetAsync(ctx->d_err, 0, 4, s); ctx->err = "UpdateMask: mask label outside [0, 4096)"; return VIDO_ERR_ARG; }
  for (int k = 0; k < nl && k < cap; k++) { if (uniq_out) uniq_out[k] = uni[k]; if (recovered) recovered[k] = rec[k]; }
  return nl;
}
