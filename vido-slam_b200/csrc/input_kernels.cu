// Pixel conversions of the demo's input staging (vido_slam/demo/run_vido_slam.cc:114-122) on the device:
//   cv::cvtColor(raw, bgr, COLOR_BayerRG2BGR)   8-bit bilinear demosaic (OpenCV modules/imgproc/src/demosaicing.cpp, Bayer2RGB_)
//   depth.convertTo(CV_32F)                     u16 -> f32
//   mask.convertTo(CV_32SC1)                    u8  -> i32
// HBM-bound byte work: every thread produces 4 output pixels from 32-bit / 64-bit loads, stores are 128-bit where the type allows.
// OpenCV is not part of /root/reference; the demosaic was restated from its published bilinear scheme and pinned against
// cv2 4.13 (tests/golden/make_input_golden.py): at the site (y, x) of the mosaic, with c the sample itself,
//   (even, even): B = c, G = cross4, R = diag4        (even, odd): B = horizontal2, G = c, R = vertical2
//   (odd,  even): B = vertical2, G = c, R = horizontal2    (odd, odd): B = diag4, G = cross4, R = c
// cross4 / diag4 = (sum of the four neighbours + 2) >> 2, horizontal2 / vertical2 = (sum of the two + 1) >> 1; then the first and
// last column copy their inner neighbour and the first and last row copy theirs.
#include "ctx.h"

namespace {
__device__ __forceinline__ void demosaic_site(const uint8_t* __restrict__ src, int pitch, int y, int x, uint8_t* bgr) {
  const uint8_t* p = src + (size_t)y * pitch + x;
  const int c = p[0], up = p[-pitch], dn = p[pitch], lf = p[-1], rt = p[1];
  const int cross = (up + dn + lf + rt + 2) >> 2, hor = (lf + rt + 1) >> 1, ver = (up + dn + 1) >> 1;
  const int diag = (p[-pitch - 1] + p[-pitch + 1] + p[pitch - 1] + p[pitch + 1] + 2) >> 2;
  const int py = y & 1, px = x & 1;
  int b, g, r;
  if (!py && !px) { b = c; g = cross; r = diag; }
  else if (!py && px) { b = hor; g = c; r = ver; }
  else if (py && !px) { b = ver; g = c; r = hor; }
  else { b = diag; g = cross; r = c; }
  bgr[0] = (uint8_t)b; bgr[1] = (uint8_t)g; bgr[2] = (uint8_t)r;
}

// one thread = 4 consecutive output pixels (12 bytes) of one row; border pixels take the value of the clamped interior site
__global__ void __launch_bounds__(256) bayer_rg2bgr_kernel(const uint8_t* __restrict__ src, int spitch, size_t sfs, int W, int H,
                                                           uint8_t* __restrict__ dst, int dpitch, size_t dfs) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, f = blockIdx.z;
  if (x4 >= W) return;
  const uint8_t* s = src + (size_t)f * sfs;
  uint8_t* d = dst + (size_t)f * dfs + (size_t)y * dpitch + 3 * (size_t)x4;
  const int ys = min(max(y, 1), H - 2);
  uint8_t out[12];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int xs = min(max(x4 + i, 1), W - 2);
    demosaic_site(s, spitch, ys, xs, out + 3 * i);
  }
  if (x4 + 4 <= W && ((size_t)d & 3) == 0) {
    uint32_t* d32 = (uint32_t*)d;
    d32[0] = out[0] | (out[1] << 8) | (out[2] << 16) | ((uint32_t)out[3] << 24);
    d32[1] = out[4] | (out[5] << 8) | (out[6] << 16) | ((uint32_t)out[7] << 24);
    d32[2] = out[8] | (out[9] << 8) | (out[10] << 16) | ((uint32_t)out[11] << 24);
  } else {
    for (int i = 0; i < 12 && x4 + i / 3 < W; i++) d[i] = out[i];
  }
}

__global__ void __launch_bounds__(256) u16_to_f32_kernel(const uint16_t* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 4 <= n && ((size_t)(src + i4) & 7) == 0 && ((size_t)(dst + i4) & 15) == 0) {
    const uint2 v = *(const uint2*)(src + i4);
    *(float4*)(dst + i4) = make_float4((float)(v.x & 0xffffu), (float)(v.x >> 16), (float)(v.y & 0xffffu), (float)(v.y >> 16));
  } else {
    for (size_t i = i4; i < n && i < i4 + 4; i++) dst[i] = (float)src[i];
  }
}

__global__ void __launch_bounds__(256) u8_to_i32_kernel(const uint8_t* __restrict__ src, int32_t* __restrict__ dst, size_t n) {
  const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 4 <= n && ((size_t)(src + i4) & 3) == 0 && ((size_t)(dst + i4) & 15) == 0) {
    const uint32_t v = *(const uint32_t*)(src + i4);
    *(int4*)(dst + i4) = make_int4((int)(v & 0xffu), (int)((v >> 8) & 0xffu), (int)((v >> 16) & 0xffu), (int)(v >> 24));
  } else {
    for (size_t i = i4; i < n && i < i4 + 4; i++) dst[i] = (int32_t)src[i];
  }
}
}  // namespace

// raw inputs: host (any) or device pointers, tightly packed W x H per frame; outputs: device pointers (tightly packed)
int input_convert_raw(vido_ctx* ctx, const uint8_t* bayer, const uint16_t* depth16, const uint8_t* mask8, int nframes, uint8_t* d_bgr,
                      float* d_depth, int32_t* d_mask) {
  const int W = ctx->cfg.width, H = ctx->cfg.height;
  if (W < 4 || H < 4) { ctx->err = "vido_convert_raw: image too small"; return VIDO_ERR_ARG; }
  const size_t px = (size_t)W * H, n = px * (size_t)nframes;
  cudaStream_t s = ctx->stream;
  auto on_device = [](const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
  };
  // staging for host sources (grown on demand, kept by the context)
  auto stage = [&](const void* src, size_t bytes, void** slot, size_t* cap) -> const void* {
    if (on_device(src)) return src;
    if (*cap < bytes) {
      if (*slot) cudaFree(*slot);
      *slot = nullptr; *cap = 0;
      if (cudaMalloc(slot, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      *cap = bytes;
    }
    if (cudaMemcpyAsync(*slot, src, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return *slot;
  };
  if (bayer && d_bgr) {
    const uint8_t* src = (const uint8_t*)stage(bayer, n, &ctx->raw_stage[0], &ctx->raw_cap[0]);
    if (!src) { ctx->err = "vido_convert_raw: staging the raw image failed"; return VIDO_ERR_CUDA; }
    dim3 grid(((W + 3) / 4 + 255) / 256, H, nframes);
    bayer_rg2bgr_kernel<<<grid, 256, 0, s>>>(src, W, px, W, H, d_bgr, 3 * W, 3 * px);
    ctx->launches++;
  }
  if (depth16 && d_depth) {
    const uint16_t* src = (const uint16_t*)stage(depth16, 2 * n, &ctx->raw_stage[1], &ctx->raw_cap[1]);
    if (!src) { ctx->err = "vido_convert_raw: staging the depth image failed"; return VIDO_ERR_CUDA; }
    u16_to_f32_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, s>>>(src, d_depth, n);
    ctx->launches++;
  }
  if (mask8 && d_mask) {
    const uint8_t* src = (const uint8_t*)stage(mask8, n, &ctx->raw_stage[2], &ctx->raw_cap[2]);
    if (!src) { ctx->err = "vido_convert_raw: staging the mask failed"; return VIDO_ERR_CUDA; }
    u8_to_i32_kernel<<<(unsigned)((n / 4 + 256) / 256), 256, 0, s>>>(src, d_mask, n);
    ctx->launches++;
  }
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}
