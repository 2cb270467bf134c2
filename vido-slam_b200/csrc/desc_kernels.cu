// desc_kernels.cu -- descriptor stage of the ORB front-end: the per-level 7x7 Gaussian the reference runs in
// ORBextractor::operator() (src/ORBextractor.cc:1078-1079), rBRIEF (computeOrbDescriptor, :98-137, whose call at :1086 the
// reference has commented out: its descriptor matrix stays uninitialised) and a brute-force Hamming matcher (the
// reference has none).  BASELINE.json's north_star names all three, so they exist here as optional entry points
// beside the extraction; the tracking path does not call them -- like the reference's, it associates by optical flow.
// The kernels are HBM / L1-bound byte and bit work; their per-thread bodies live in desc_device.h (shared with the CPU
// emulation of the test-suite), this file holds the launch geometry and the workspace.
#include <stdlib.h>
#include <string.h>

#include "ctx.h"
#include "desc_device.h"
#include "../../include/vido_orb_pattern.h"

namespace {

struct DescWorkspace {
  uint8_t* d_blur = nullptr;      // blurred pyramid, same layout as ctx->d_pyr
  int8_t* d_pattern = nullptr;    // 512 (x, y) pairs
  uint8_t* d_desc = nullptr;      // [max_batch][kp_cap][32] staging of the host-pointer entry point
  uint8_t* d_ham = nullptr;       // grow-only scratch of the matcher (partials, and the staging of the host-pointer entry point)
  size_t ham_bytes = 0;
  BlurParams blur;
  DescParams desc;
  int blur_version = 1;           // VIDO_BLUR=v2 selects the sliding-window variant (desc_device.h), read when the workspace is created
};

__global__ void __launch_bounds__(BLUR_THREADS) blur7_kernel(BlurParams P, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
  blur7_thread(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.z, P, src, dst);
}

__global__ void __launch_bounds__(BLUR_THREADS) blur7_v2_kernel(BlurParams P, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst) {
  blur7_thread_v2(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.z, P, src, dst);
}

__global__ void __launch_bounds__(RBRIEF_THREADS) rbrief_kernel(DescParams P, const uint8_t* __restrict__ blurred, const DescKeyPoint* __restrict__ kps,
                                                     const int32_t* __restrict__ nkp, const int8_t* __restrict__ pattern,
                                                     uint8_t* __restrict__ desc) {
  rbrief_thread(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.z, P, blurred, kps, nkp, pattern, desc);
}

__global__ void __launch_bounds__(HAM_THREADS) hamming_partial_kernel(HamParams P, const uint8_t* __restrict__ q, const uint8_t* __restrict__ t,
                                                              const int32_t* __restrict__ nq, const int32_t* __restrict__ nt,
                                                              int32_t* __restrict__ part) {
  hamming_partial_thread(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, blockIdx.z, P, q, t, nq, nt, part);
}

__global__ void __launch_bounds__(HAM_THREADS) hamming_merge_kernel(HamParams P, const int32_t* __restrict__ nq, const int32_t* __restrict__ part,
                                                            int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist,
                                                            int32_t* __restrict__ second_dist) {
  hamming_merge_thread(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.z, P, nq, part, best_idx, best_dist, second_dist);
}

static_assert(sizeof(DescKeyPoint) == sizeof(vido_keypoint), "key point layout");

}  // namespace

// the matcher's scratch: grow-only, doubling (a cudaMalloc per call would synchronise the device, see vido_scratch)
static uint8_t* ham_scratch(vido_ctx* ctx, DescWorkspace* ws, size_t bytes) {
  if (ws->ham_bytes < bytes) {
    if (ws->d_ham) { cudaStreamSynchronize(ctx->stream); cudaFree(ws->d_ham); }
    ws->d_ham = nullptr; ws->ham_bytes = 0;
    if (cudaMalloc(&ws->d_ham, bytes * 2) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ws->ham_bytes = bytes * 2;
  }
  return ws->d_ham;
}

// created on first use: a context that only tracks never pays for the blurred pyramid
static int desc_workspace_build(vido_ctx* ctx, DescWorkspace* ws) {
  const int B = ctx->cfg.max_batch;
  VIDO_CUDA(cudaMalloc(&ws->d_blur, ctx->pyr_bytes));
  VIDO_CUDA(cudaMalloc(&ws->d_pattern, sizeof vido_orb_pattern_31));
  VIDO_CUDA(cudaMalloc(&ws->d_desc, (size_t)B * ctx->kp_cap * 32));
  VIDO_CUDA(cudaMemcpy(ws->d_pattern, vido_orb_pattern_31, sizeof vido_orb_pattern_31, cudaMemcpyHostToDevice));
  memset(&ws->blur, 0, sizeof ws->blur);
  memset(&ws->desc, 0, sizeof ws->desc);
  {
    const char* v = getenv("VIDO_BLUR");
    ws->blur_version = (v && !strcmp(v, "v2")) ? 2 : 1;
  }
  for (int l = 0; l < ctx->nlevels; l++) {
    const OrbLevel& L = ctx->lv[l];
    if (L.w < 8 || L.h < 8 || (L.pitch & 3) || (L.base & 3) || (L.frame_stride & 3)) {
      ctx->err = "descriptor stage: pyramid level too small or misaligned";
      return VIDO_ERR_ARG;
    }
    if (ws->blur_version == 2) blur2_params_add_level(ws->blur, l, L.w, L.h, L.pitch, (long long)L.base, (long long)L.frame_stride);
    else blur_params_add_level(ws->blur, l, L.w, L.h, L.pitch, (long long)L.base, (long long)L.frame_stride);
    desc_params_add_level(ws->desc, l, L.w, L.h, L.pitch, (long long)L.base, (long long)L.frame_stride, L.scale);
  }
  return VIDO_OK;
}

static int desc_workspace(vido_ctx* ctx, DescWorkspace** out) {
  if (!ctx->desc) {
    DescWorkspace* ws = new DescWorkspace();
    const int rc = desc_workspace_build(ctx, ws);
    if (rc != VIDO_OK) {
      cudaFree(ws->d_blur); cudaFree(ws->d_pattern); cudaFree(ws->d_desc);
      cudaGetLastError();
      delete ws;
      return rc;
    }
    ctx->desc = ws;
  }
  *out = (DescWorkspace*)ctx->desc;
  return VIDO_OK;
}

void desc_teardown(vido_ctx* ctx) {
  DescWorkspace* ws = (DescWorkspace*)ctx->desc;
  if (!ws) return;
  cudaFree(ws->d_blur); cudaFree(ws->d_pattern); cudaFree(ws->d_desc); cudaFree(ws->d_ham);
  delete ws;
  ctx->desc = nullptr;
}

// blur + describe the key points of the first `nframes` batch slots of the last extraction (device pointers, context stream)
int desc_run(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int nframes, int cap_per_frame, uint8_t* d_desc) {
  if (nframes < 1 || nframes > ctx->last_batch) { ctx->err = "describe: nframes exceeds the last extraction's batch"; return VIDO_ERR_ARG; }
  DescWorkspace* ws = nullptr;
  { int rc = desc_workspace(ctx, &ws); if (rc) return rc; }
  cudaStream_t st = ctx->stream;
  ws->blur.nframes = nframes;
  {
    dim3 grid(blur_grid_x(ws->blur), 1, nframes);
    if (ws->blur_version == 2) blur7_v2_kernel<<<grid, BLUR_THREADS, 0, st>>>(ws->blur, ctx->d_pyr, ws->d_blur);
    else blur7_kernel<<<grid, BLUR_THREADS, 0, st>>>(ws->blur, ctx->d_pyr, ws->d_blur);
    ctx->launches++;
  }
  ws->desc.nframes = nframes;
  ws->desc.cap_per_frame = cap_per_frame;
  {
    dim3 grid(rbrief_grid_x(cap_per_frame), 1, nframes);
    rbrief_kernel<<<grid, RBRIEF_THREADS, 0, st>>>(ws->desc, ws->d_blur, (const DescKeyPoint*)d_kps, d_nkp, ws->d_pattern, d_desc);
    ctx->launches++;
  }
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

uint8_t* desc_staging(vido_ctx* ctx) {
  DescWorkspace* ws = nullptr;
  if (desc_workspace(ctx, &ws) != VIDO_OK) return nullptr;
  return ws->d_desc;
}

int desc_get_blurred_level(vido_ctx* ctx, int frame, int level, uint8_t* out) {
  DescWorkspace* ws = (DescWorkspace*)ctx->desc;
  if (!ws) { ctx->err = "no descriptor pass has run"; return VIDO_ERR_STATE; }
  const OrbLevel& L = ctx->lv[level];
  VIDO_CUDA(cudaMemcpy2DAsync(out, L.w, ws->d_blur + L.base + (size_t)frame * L.frame_stride, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost,
                              ctx->stream));
  VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIDO_OK;
}

int desc_match_dev(vido_ctx* ctx, const uint8_t* d_q, size_t q_stride, const int32_t* d_nq, const uint8_t* d_t, size_t t_stride,
                   const int32_t* d_nt, int npairs, int qcap, int32_t* d_best_idx, int32_t* d_best_dist, int32_t* d_second_dist,
                   uint8_t* d_part) {
  HamParams P;
  P.npairs = npairs; P.qcap = qcap;
  P.q_stride = (long long)q_stride; P.t_stride = (long long)t_stride;
  cudaStream_t st = ctx->stream;
  {
    dim3 grid(hamming_grid_x(qcap), HAM_CHUNKS, npairs);
    hamming_partial_kernel<<<grid, HAM_THREADS, 0, st>>>(P, d_q, d_t, d_nq, d_nt, (int32_t*)d_part);
    ctx->launches++;
  }
  {
    dim3 grid(hamming_grid_x(qcap), 1, npairs);
    hamming_merge_kernel<<<grid, HAM_THREADS, 0, st>>>(P, d_nq, (const int32_t*)d_part, d_best_idx, d_best_dist, d_second_dist);
    ctx->launches++;
  }
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

int desc_match_device_api(vido_ctx* ctx, const uint8_t* d_q, size_t q_stride, const int32_t* d_nq, const uint8_t* d_t, size_t t_stride,
                          const int32_t* d_nt, int npairs, int qcap, int32_t* d_best_idx, int32_t* d_best_dist, int32_t* d_second_dist) {
  DescWorkspace* ws = nullptr;
  { int rc = desc_workspace(ctx, &ws); if (rc) return rc; }
  uint8_t* part = ham_scratch(ctx, ws, hamming_part_bytes(npairs, qcap));
  if (!part) { ctx->err = "matcher scratch allocation failed"; return VIDO_ERR_CUDA; }
  return desc_match_dev(ctx, d_q, q_stride, d_nq, d_t, t_stride, d_nt, npairs, qcap, d_best_idx, d_best_dist, d_second_dist, part);
}

// host pointers, one pair
int desc_match_host(vido_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* best_idx, int32_t* best_dist,
                    int32_t* second_dist) {
  DescWorkspace* ws = nullptr;
  { int rc = desc_workspace(ctx, &ws); if (rc) return rc; }
  if (nq == 0) return VIDO_OK;
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t oq = 0, ot = oq + up((size_t)nq * 32), on = ot + up((size_t)nt * 32 + 4), oo = on + 256, op = oo + up((size_t)nq * 12),
               total = op + hamming_part_bytes(1, nq);
  uint8_t* base = ham_scratch(ctx, ws, total);
  if (!base) { ctx->err = "matcher scratch allocation failed"; return VIDO_ERR_CUDA; }
  cudaStream_t st = ctx->stream;
  const int32_t counts[2] = {nq, nt};
  VIDO_CUDA(cudaMemcpyAsync(base + oq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, st));
  if (nt > 0) VIDO_CUDA(cudaMemcpyAsync(base + ot, t, (size_t)nt * 32, cudaMemcpyHostToDevice, st));
  VIDO_CUDA(cudaMemcpyAsync(base + on, counts, sizeof counts, cudaMemcpyHostToDevice, st));
  int32_t* d_out = (int32_t*)(base + oo);
  int rc = desc_match_dev(ctx, base + oq, 0, (const int32_t*)(base + on), base + ot, 0, (const int32_t*)(base + on) + 1, 1, nq, d_out,
                          d_out + nq, d_out + 2 * (size_t)nq, base + op);
  if (rc) return rc;
  VIDO_CUDA(cudaMemcpyAsync(best_idx, d_out, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, st));
  VIDO_CUDA(cudaMemcpyAsync(best_dist, d_out + nq, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, st));
  VIDO_CUDA(cudaMemcpyAsync(second_dist, d_out + 2 * (size_t)nq, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost, st));
  VIDO_CUDA(cudaStreamSynchronize(st));
  return VIDO_OK;
}
