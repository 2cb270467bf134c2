// track.cu -- per-sequence driver of the hot path (VO, static scene): the host-side state machine of
// Tracking::GrabImageRGBD / Track around the CUDA stages, with the Map kept as flat per-frame arrays and
// tracklets maintained incrementally (O(features) per frame instead of the reference's rebuild from frame 0).
//
// Replaces (paths under /root/reference/vido_slam/src):
//   Tracking::GrabImageRGBD    Tracking.cc:283-456    Tracking::Track            Tracking.cc:1081-1509
//   Tracking::Initialization   Tracking.cc:1512-1580  Tracking::RenewFrameInfo   Tracking.cc:2959-3135 (static part)
//   Tracking::GetStaticTrack   Tracking.cc:2514-2613  (incremental form, same tracklets)
//   Optimizer::PartialBatchOptimization graph construction / write-back  Optimizer.cc:43-362, 1056-1142
// The front-end (gray conversion, pyramid, FAST, quad-tree, orientation, association, per-keypoint map lookups) is run
// for a whole chunk of frames at once -- it has no inter-frame dependency -- the back-end is sequential per frame.
// Scope of this version: sensor RGBD, bJoint = true, UseSampleFeature = 0, all-zero object mask (no dynamic objects),
// no IMU.  float 4x4 products use double accumulation + one rounding like cv::Mat CV_32F gemm.
#include <algorithm>
#include <chrono>
#include <cstring>

#include "ctx.h"

namespace {

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void mul44(const float* A, const float* B, float* C) {
  float o[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += (double)A[4 * r + k] * (double)B[4 * k + c];
      o[4 * r + c] = (float)s;
    }
  memcpy(C, o, sizeof o);
}
void inv44(const float* T, float* Ti) {
  float o[16] = {0};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[4 * r + c] = T[4 * c + r];
  for (int r = 0; r < 3; r++) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)(-o[4 * r + k]) * (double)T[4 * k + 3];
    o[4 * r + 3] = (float)s;
  }
  o[15] = 1.f;
  memcpy(Ti, o, sizeof o);
}
void eye44(float* T) { memset(T, 0, sizeof(float) * 16); T[0] = T[5] = T[10] = T[15] = 1.f; }

struct MapFrame {           // Map::vpFeatSta / vfDepSta / vp3DPointSta / vnAssoSta of one frame + its pose
  std::vector<float> xy, depth, p3;
  std::vector<int> asso, track, pos;
  float Twc[16];            // vmCameraPose
  float rel[16];            // vmRigidMotion[f-1][0]
};
struct TrackInfo { int first_frame, len, pid, epoch; };  // pid valid for the window graph built in `epoch`

struct FrontFrame {         // front-end results of one frame, on the host
  std::vector<vido_keypoint> kps;
  std::vector<int32_t> kp_mask;          // per keypoint: mask / depth / flow at its truncated position
  std::vector<float> kp_depth, kp_flow;
  std::vector<int32_t> as_idx;           // Frame-ctor association
  std::vector<float> as_corres, as_flow, as_depth;
};

}  // namespace

struct TrackState {
  // sequence state
  std::vector<MapFrame> map;
  std::vector<TrackInfo> tracks;
  bool initialised = false, has_velocity = false;
  float mVelocity[16];
  float lastTcw[16];
  std::vector<float> last_keys, last_depth, last_corres, last_flow;  // mpLastFrame mvStatKeys / mvStatDepth / mvCorres / mvFlowNext
  int f_id = 0;
  int ba_epoch = 0;
  // window BA jobs: [fly] is being solved on the BA stream while the next frame is tracked and its job [fly ^ 1] is staged
  struct BaJob {
    int start = 0, end = 0;
    vido_track_stats* st = nullptr;
    vido_ba_problem pr;
    std::vector<float> poses, rel, pts, oxyz;
    std::vector<int> op, ol;
    std::vector<int> ofeat;          // per observation: feature index inside its frame, for the write-back
    std::vector<int> pframe, pfeat;  // per point: frame / feature of its first observation (initial position)
  } job[2];
  bool ba_pending = false, ba_staged = false;
  int ba_fly = 0, ba_stage_slot = 0, ba_rest = -1;
  // device buffers of one chunk
  int capB = 0;
  // Two pipeline slots.  A slot holds the inputs of one batch on the device, its front-end outputs on the device and
  // their pinned host mirror.  While the back-end walks through the batch of slot s, the next batch (or the first batch
  // of the next call, see vido_track_prefetch) is copied (copy stream) and run through the front-end (front-end stream)
  // in slot s ^ 1.
  struct FeSlot {
    uint8_t* d_img = nullptr;   // [B][H][W*3] or gray   } own input buffers (host / scattered inputs are copied here)
    float* d_depth = nullptr;   // [B][H][W]             }
    float* d_flow = nullptr;    // [B][H][W][2]          }
    int32_t* d_mask = nullptr;  // [B][H][W]             }
    char* d_out = nullptr;      // front-end outputs, one block: kp | kpmask | kpdepth | kpflow | asidx | ascor | asflow | asdepth | nkp | asn | err
    char* h_out = nullptr;      // pinned mirror
    vido_keypoint* d_kp; int32_t* d_kpmask; float* d_kpdepth; float* d_kpflow;
    int32_t* d_asidx; float* d_ascor; float* d_asflow; float* d_asdepth; int32_t* d_nkp; int32_t* d_asn; int32_t* d_flag;
    cudaEvent_t copied = nullptr, done = nullptr, ev0 = nullptr, ev1 = nullptr;
    bool launched = false;
    int B = 0, channels = 0;
    const void* key = nullptr;  // identity of the batch: image pointer of its first frame
    const uint8_t* in_img = nullptr; const float* in_depth = nullptr; const float* in_flow = nullptr; const int32_t* in_mask = nullptr;  // inputs the kernels read
  } fe[2];
  size_t fe_out_bytes = 0;
  int fe_cur = 0;
  uint8_t* d_gray = nullptr;     // [B][H][W] (front-end stream only)
  cudaStream_t copy_stream = nullptr, fe_stream = nullptr;
  std::vector<vido_frame_inputs> hint;  // frames announced by vido_track_prefetch
  float* d_q = nullptr; int32_t* d_qmask = nullptr; float* d_qdepth = nullptr; float* d_qflow = nullptr;  // per-frame queries
  float* d_check = nullptr; uint8_t* d_used = nullptr;
  // pinned host mirrors
  char* h_pin = nullptr; size_t h_pin_bytes = 0;
  int kp_cap = 0, q_cap = 8192;
};

// ---- per-keypoint map lookups for the renewal top-up (Tracking.cc:3044-3062): mask, depth, flow at (int)pt
__global__ void kp_lookup_kernel(const vido_keypoint* __restrict__ kps, const int32_t* __restrict__ nkp, int kp_cap, int w, int h,
                                 const float* __restrict__ depth, const float* __restrict__ flow, const int32_t* __restrict__ mask,
                                 size_t img_fs, int mode, float factor, float bf, float mscale, int32_t* __restrict__ omask,
                                 float* __restrict__ odepth, float* __restrict__ oflow) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nkp[b]) return;
  const size_t o = (size_t)b * kp_cap + i;
  const int x = (int)kps[o].x, y = (int)kps[o].y;
  if (x < 0 || y < 0 || x >= w || y >= h) { omask[o] = -1; odepth[o] = 0; oflow[2 * o] = 0; oflow[2 * o + 1] = 0; return; }
  const size_t k = b * img_fs + (size_t)y * w + x;
  float d = depth[k];
  if (mode) {
    if (d < 0) d = 0.f;
    else if (mode == 1) d = __fdiv_rn(d, factor);
    else if (mode == 2) d = __fdiv_rn(bf, __fdiv_rn(d, factor));
    else d = __fdiv_rn(__fmul_rn(mscale, bf), __fdiv_rn(d, factor));
  }
  omask[o] = mask[k];
  odepth[o] = d;
  oflow[2 * o] = flow[2 * k];
  oflow[2 * o + 1] = flow[2 * k + 1];
}

// used[i] = 1 if keypoint i lies within 1 px (Euclidean, float) of any already selected feature (Tracking.cc:3030-3040).
// One warp per keypoint: the lanes stride over the selected features, any hit ends the search (2.5 k x 1 k distance tests
// per frame; a thread per keypoint walking the whole list serially was the longest kernel of the tracking path).
__global__ void __launch_bounds__(256) topup_used_kernel(const vido_keypoint* __restrict__ kps, int n, const float* __restrict__ check, int m,
                                                         uint8_t* __restrict__ used) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float sx = kps[i].x, sy = kps[i].y;
  bool u = false;
  for (int j0 = 0; j0 < m; j0 += 32) {
    const int j = j0 + lane;
    bool hit = false;
    if (j < m) {
      const float dx = __fsub_rn(check[2 * j], sx), dy = __fsub_rn(check[2 * j + 1], sy);
      hit = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) < 1.0f;
    }
    if (__any_sync(0xffffffffu, hit)) { u = true; break; }
  }
  if (lane == 0) used[i] = u ? 1 : 0;
}

// latency-critical streams (per-frame back-end, window BA) get the highest priority, the run-ahead streams (input copies,
// front-end of the next batch) the lowest: when both have blocks ready the back-end's short kernels are scheduled first
cudaError_t vido_create_stream(cudaStream_t* s, bool high) {
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);  // lo = least priority (numerically greatest)
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, high ? hi : lo);
}

int trk_setup(vido_ctx* ctx) {
  TrackState* ts = new TrackState();
  ctx->trk = ts;
  const vido_config& c = ctx->cfg;
  const int B = c.max_batch;
  ts->capB = B;
  ts->kp_cap = ctx->kp_cap;
  const size_t px = (size_t)c.width * c.height;
  const size_t K = (size_t)ts->kp_cap * B;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_kp = 0, o_kpmask = o_kp + al(sizeof(vido_keypoint) * K), o_kpdepth = o_kpmask + al(4 * K), o_kpflow = o_kpdepth + al(4 * K),
               o_asidx = o_kpflow + al(8 * K), o_ascor = o_asidx + al(4 * K), o_asflow = o_ascor + al(8 * K), o_asdepth = o_asflow + al(8 * K),
               o_nkp = o_asdepth + al(4 * K), o_asn = o_nkp + al(4 * (size_t)B), o_flag = o_asn + al(4 * (size_t)B);
  ts->fe_out_bytes = o_flag + 256;
  for (int k = 0; k < 2; k++) {
    TrackState::FeSlot& F = ts->fe[k];
    VIDO_CUDA(cudaMalloc(&F.d_img, px * 3 * B));
    VIDO_CUDA(cudaMalloc(&F.d_depth, px * 4 * B));
    VIDO_CUDA(cudaMalloc(&F.d_flow, px * 8 * B));
    VIDO_CUDA(cudaMalloc(&F.d_mask, px * 4 * B));
    VIDO_CUDA(cudaMalloc(&F.d_out, ts->fe_out_bytes));
    VIDO_CUDA(cudaMemset(F.d_out, 0, ts->fe_out_bytes));
    VIDO_CUDA(cudaMallocHost(&F.h_out, ts->fe_out_bytes));
    F.d_kp = (vido_keypoint*)(F.d_out + o_kp); F.d_kpmask = (int32_t*)(F.d_out + o_kpmask); F.d_kpdepth = (float*)(F.d_out + o_kpdepth);
    F.d_kpflow = (float*)(F.d_out + o_kpflow); F.d_asidx = (int32_t*)(F.d_out + o_asidx); F.d_ascor = (float*)(F.d_out + o_ascor);
    F.d_asflow = (float*)(F.d_out + o_asflow); F.d_asdepth = (float*)(F.d_out + o_asdepth); F.d_nkp = (int32_t*)(F.d_out + o_nkp);
    F.d_asn = (int32_t*)(F.d_out + o_asn); F.d_flag = (int32_t*)(F.d_out + o_flag);
    VIDO_CUDA(cudaEventCreateWithFlags(&F.copied, cudaEventDisableTiming));
    VIDO_CUDA(cudaEventCreateWithFlags(&F.done, cudaEventDisableTiming));
    VIDO_CUDA(cudaEventCreate(&F.ev0));
    VIDO_CUDA(cudaEventCreate(&F.ev1));
  }
  VIDO_CUDA(cudaMalloc(&ts->d_gray, px * B));
  VIDO_CUDA(vido_create_stream(&ts->copy_stream, false));
  VIDO_CUDA(vido_create_stream(&ts->fe_stream, false));
  VIDO_CUDA(cudaMalloc(&ts->d_q, 8 * ts->q_cap)); VIDO_CUDA(cudaMalloc(&ts->d_qmask, 4 * ts->q_cap));
  VIDO_CUDA(cudaMalloc(&ts->d_qdepth, 4 * ts->q_cap)); VIDO_CUDA(cudaMalloc(&ts->d_qflow, 8 * ts->q_cap));
  VIDO_CUDA(cudaMalloc(&ts->d_check, 8 * ts->q_cap)); VIDO_CUDA(cudaMalloc(&ts->d_used, ts->kp_cap));
  return VIDO_OK;
}

void trk_teardown(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts) return;
  if (ts->copy_stream) { cudaStreamSynchronize(ts->copy_stream); cudaStreamDestroy(ts->copy_stream); }
  if (ts->fe_stream) { cudaStreamSynchronize(ts->fe_stream); cudaStreamDestroy(ts->fe_stream); }
  for (int k = 0; k < 2; k++) {
    TrackState::FeSlot& F = ts->fe[k];
    cudaFree(F.d_img); cudaFree(F.d_depth); cudaFree(F.d_flow); cudaFree(F.d_mask); cudaFree(F.d_out); cudaFreeHost(F.h_out);
    if (F.copied) cudaEventDestroy(F.copied);
    if (F.done) cudaEventDestroy(F.done);
    if (F.ev0) cudaEventDestroy(F.ev0);
    if (F.ev1) cudaEventDestroy(F.ev1);
  }
  cudaFree(ts->d_gray);
  cudaFree(ts->d_q); cudaFree(ts->d_qmask); cudaFree(ts->d_qdepth); cudaFree(ts->d_qflow); cudaFree(ts->d_check); cudaFree(ts->d_used);
  delete ts;
  ctx->trk = nullptr;
}

int trk_reset(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (ts->ba_pending) { vido_lm_stats ls; ba_collect(ctx, &ts->job[ts->ba_fly].pr, &ls); ts->ba_pending = false; }
  ts->map.clear(); ts->tracks.clear();
  ts->initialised = false; ts->has_velocity = false; ts->f_id = 0; ts->ba_epoch = 0;
  cudaStreamSynchronize(ts->copy_stream); cudaStreamSynchronize(ts->fe_stream);
  ts->fe[0].launched = ts->fe[1].launched = false; ts->hint.clear(); ts->ba_rest = -1; ts->ba_staged = false;
  ts->last_keys.clear(); ts->last_depth.clear(); ts->last_corres.clear(); ts->last_flow.clear();
  return VIDO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// front-end of a batch, asynchronous: fe_launch enqueues input copies (copy stream), gray conversion, ORB, association,
// per-keypoint lookups and the device->host copy of the results (front-end stream); fe_collect waits and unpacks.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) h2d_stream_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes);
static int prefetch_array(vido_ctx* ctx, cudaStream_t cs, void* dst, const void* src, size_t bytes);

static int fe_launch(vido_ctx* ctx, TrackState::FeSlot& F, const vido_frame_inputs* in, int B) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const size_t px = (size_t)c.width * c.height;
  const vido_frame_inputs* f0 = in;
  const int channels = f0->channels;
  bool contiguous_dev = f0->on_device != 0;
  for (int b = 0; b < B; b++) {
    const vido_frame_inputs& f = in[b];
    if (f.channels != channels || (f.on_device != 0) != (f0->on_device != 0) || (channels != 1 && channels != 3)) { ctx->err = "inconsistent frame inputs"; return VIDO_ERR_ARG; }
    if (f.on_device) {
      // device-resident frames are used in place when they are consecutive slices of one allocation
      if (f.image != f0->image + (size_t)b * px * channels || f.depth != f0->depth + (size_t)b * px ||
          f.flow != f0->flow + (size_t)b * px * 2 || f.mask != f0->mask + (size_t)b * px) contiguous_dev = false;
    }
  }
  cudaStream_t fs = ts->fe_stream;
  if (f0->on_device && contiguous_dev) {
    F.in_img = f0->image; F.in_depth = f0->depth; F.in_flow = f0->flow; F.in_mask = f0->mask;
  } else {
    cudaStream_t cs = ts->copy_stream;
    for (int b = 0; b < B; b++) {
      const vido_frame_inputs& f = in[b];
      if (f.on_device) {
        VIDO_CUDA(cudaMemcpyAsync(F.d_img + (size_t)b * px * channels, f.image, px * channels, cudaMemcpyDeviceToDevice, cs));
        VIDO_CUDA(cudaMemcpyAsync(F.d_depth + (size_t)b * px, f.depth, px * 4, cudaMemcpyDeviceToDevice, cs));
        VIDO_CUDA(cudaMemcpyAsync(F.d_flow + (size_t)b * px * 2, f.flow, px * 8, cudaMemcpyDeviceToDevice, cs));
        VIDO_CUDA(cudaMemcpyAsync(F.d_mask + (size_t)b * px, f.mask, px * 4, cudaMemcpyDeviceToDevice, cs));
      } else {
        int rc = prefetch_array(ctx, cs, F.d_img + (size_t)b * px * channels, f.image, px * channels);
        if (!rc) rc = prefetch_array(ctx, cs, F.d_depth + (size_t)b * px, f.depth, px * 4);
        if (!rc) rc = prefetch_array(ctx, cs, F.d_flow + (size_t)b * px * 2, f.flow, px * 8);
        if (!rc) rc = prefetch_array(ctx, cs, F.d_mask + (size_t)b * px, f.mask, px * 4);
        if (rc) return rc;
      }
    }
    VIDO_CUDA(cudaEventRecord(F.copied, cs));
    VIDO_CUDA(cudaStreamWaitEvent(fs, F.copied, 0));
    F.in_img = F.d_img; F.in_depth = F.d_depth; F.in_flow = F.d_flow; F.in_mask = F.d_mask;
  }
  // the ORB / association entry points launch on ctx->stream: point it at the front-end stream for this section
  cudaStream_t saved = ctx->stream;
  ctx->stream = fs;
  int rc = VIDO_OK;
  do {
    const uint8_t* d_gray = F.in_img;
    cudaEventRecord(F.ev0, fs);
    if (channels == 3) {
      rc = orb_bgr_to_gray(ctx, F.in_img, B, px * 3, c.width * 3, ts->d_gray, px, c.width);
      if (rc) break;
      d_gray = ts->d_gray;
    }
    rc = orb_run(ctx, d_gray, B, px, c.width, F.d_kp, ts->kp_cap, F.d_nkp);
    if (rc) break;
    rc = assoc_frame_associate(ctx, F.d_kp, F.d_nkp, ts->kp_cap, F.in_depth, F.in_flow, F.in_mask, B, 1, F.d_asidx, F.d_ascor,
                               F.d_asflow, F.d_asdepth, F.d_asn, ts->kp_cap);
    if (rc) break;
    dim3 grid((ts->kp_cap + 255) / 256, B);
    kp_lookup_kernel<<<grid, 256, 0, fs>>>(F.d_kp, F.d_nkp, ts->kp_cap, c.width, c.height, F.in_depth, F.in_flow, F.in_mask, px,
                                           c.choose_data, c.depth_map_factor, c.bf, ctx->mscale, F.d_kpmask, F.d_kpdepth, F.d_kpflow);
    ctx->launches++;
    cudaEventRecord(F.ev1, fs);
    if (cudaGetLastError() != cudaSuccess) { ctx->err = "front-end launch failed"; rc = VIDO_ERR_CUDA; break; }
    if (cudaMemcpyAsync(F.d_flag, ctx->d_err, 4, cudaMemcpyDeviceToDevice, fs) != cudaSuccess ||
        cudaMemcpyAsync(F.h_out, F.d_out, ts->fe_out_bytes, cudaMemcpyDeviceToHost, fs) != cudaSuccess ||
        cudaEventRecord(F.done, fs) != cudaSuccess) { ctx->err = "front-end copy failed"; rc = VIDO_ERR_CUDA; break; }
  } while (0);
  ctx->stream = saved;
  if (rc) return rc;
  F.launched = true; F.B = B; F.channels = channels; F.key = (const void*)f0->image;
  return VIDO_OK;
}

static int fe_collect(vido_ctx* ctx, TrackState::FeSlot& F, std::vector<FrontFrame>& out) {
  TrackState* ts = (TrackState*)ctx->trk;
  VIDO_CUDA(cudaEventSynchronize(F.done));
  F.launched = false;
  const int B = F.B;
  const size_t K = ts->kp_cap;
  const char* h = F.h_out;
  auto hp = [&](const void* dptr) { return h + ((const char*)dptr - F.d_out); };
  if (*(const int32_t*)hp(F.d_flag)) {
    ctx->err = "ORB front-end capacity flag set";
    cudaMemsetAsync(ctx->d_err, 0, 4, ts->fe_stream);
    return VIDO_ERR_CAPACITY;
  }
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, F.ev0, F.ev1) == cudaSuccess) { ctx->t_ms[0] += ms; ctx->t_n[0] += B; }
  }
  const int32_t* nkp = (const int32_t*)hp(F.d_nkp);
  const int32_t* asn = (const int32_t*)hp(F.d_asn);
  const vido_keypoint* kp = (const vido_keypoint*)hp(F.d_kp);
  const int32_t* kpmask = (const int32_t*)hp(F.d_kpmask); const float* kpdepth = (const float*)hp(F.d_kpdepth);
  const float* kpflow = (const float*)hp(F.d_kpflow); const int32_t* asidx = (const int32_t*)hp(F.d_asidx);
  const float* ascor = (const float*)hp(F.d_ascor); const float* asflow = (const float*)hp(F.d_asflow); const float* asdepth = (const float*)hp(F.d_asdepth);
  out.resize(B);
  for (int b = 0; b < B; b++) {
    FrontFrame& f = out[b];
    const size_t n = (size_t)std::min(std::max(nkp[b], 0), ts->kp_cap), m = (size_t)std::min(std::max(asn[b], 0), ts->kp_cap);
    const size_t o = (size_t)b * K;
    f.kps.assign(kp + o, kp + o + n); f.kp_mask.assign(kpmask + o, kpmask + o + n); f.kp_depth.assign(kpdepth + o, kpdepth + o + n);
    f.kp_flow.assign(kpflow + 2 * o, kpflow + 2 * (o + n));
    f.as_idx.assign(asidx + o, asidx + o + m); f.as_corres.assign(ascor + 2 * o, ascor + 2 * (o + m));
    f.as_flow.assign(asflow + 2 * o, asflow + 2 * (o + m)); f.as_depth.assign(asdepth + o, asdepth + o + m);
  }
  return VIDO_OK;
}

// (mask, depth, flow) of frame `slot` of the resident chunk at n query positions
static int query_maps(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int slot, const float* xy,
                      int n, int32_t* omask, float* odepth, float* oflow) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (n <= 0) return VIDO_OK;
  if (n > ts->q_cap) { ctx->err = "too many map queries"; return VIDO_ERR_CAPACITY; }
  cudaStream_t s = ctx->stream;
  VIDO_CUDA(cudaMemcpyAsync(ts->d_q, xy, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  int rc = assoc_gather(ctx, d_depth, d_flow, d_mask, slot, 1, ts->d_q, n, ts->d_qmask, ts->d_qdepth, ts->d_qflow);
  if (rc) return rc;
  VIDO_CUDA(cudaMemcpyAsync(omask, ts->d_qmask, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaMemcpyAsync(odepth, ts->d_qdepth, 4 * (size_t)n, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaMemcpyAsync(oflow, ts->d_qflow, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  return VIDO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// window graph from the flat map (Optimizer.cc:220-362) + solve + write-back (:1056-1142)
// ---------------------------------------------------------------------------------------------------------
// The solve is asynchronous and its host work is split so that only the solve itself is serial:
//   ba_stage(k)  -- graph STRUCTURE and observations of frame k's window (needs only the tracker's state) -- runs while the
//                   window of frame k-1 is still being solved;
//   ba_finish    -- waits for the window in flight and writes poses / points back into the map;
//   ba_go(k)     -- reads the (just refined) poses and point positions and launches.
// Nothing the tracker reads between launch and finish (last-frame keys / depth / pose, velocity) is touched by the BA
// (Tracking.cc:1320-1500 uses mpLastFrame / mVelocity only), so the result equals the reference's sequential order.
static int ba_finish(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->ba_pending) return VIDO_OK;
  ts->ba_pending = false;
  TrackState::BaJob& J = ts->job[ts->ba_fly];
  vido_lm_stats ls;
  int rc = ba_collect(ctx, &J.pr, &ls);
  if (rc) return rc;
  if (J.st) { J.st->ba_iterations = ls.iterations; J.st->ba_trials = ls.total_trials; }
  const int start = J.start, N = J.end;
  for (int i = start; i < N; i++) {
    memcpy(ts->map[i].Twc, &J.poses[16 * (size_t)(i - start)], sizeof(float) * 16);
    if (i > start) memcpy(ts->map[i].rel, &J.rel[16 * (size_t)(i - start - 1)], sizeof(float) * 16);
  }
  // Every observation of an optimised point receives the optimised position (Optimizer.cc:1107-1122).  Only the FIRST
  // observation of each point is read again before the next solve is launched (its initial position): write those now,
  // the rest in ba_writeback_rest, after the launch.
  const size_t np = J.pframe.size();
  const float* pts = J.pts.data();
  for (size_t l = 0; l < np; l++) {
    float* d = &ts->map[J.pframe[l]].p3[3 * (size_t)J.pfeat[l]];
    d[0] = pts[3 * l]; d[1] = pts[3 * l + 1]; d[2] = pts[3 * l + 2];
  }
  ts->ba_rest = (int)(&J - ts->job);
  return VIDO_OK;
}

static void ba_writeback_rest(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (ts->ba_rest < 0) return;
  TrackState::BaJob& J = ts->job[ts->ba_rest];
  ts->ba_rest = -1;
  const size_t nobs = J.op.size();
  const float* pts = J.pts.data();
  for (size_t o = 0; o < nobs; o++) {
    float* d = &ts->map[J.start + J.op[o]].p3[3 * (size_t)J.ofeat[o]];
    const float* q = pts + 3 * (size_t)J.ol[o];
    d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
  }
}

static int ba_stage(vido_ctx* ctx, int WINDOW, vido_track_stats* st) {
  TrackState* ts = (TrackState*)ctx->trk;
  ts->ba_staged = false;
  const int N = (int)ts->map.size();
  if (st) { st->ba_iterations = -1; st->ba_points = 0; st->ba_obs = 0; st->ba_trials = 0; }
  if (WINDOW <= 0) return VIDO_OK;
  ts->ba_stage_slot = ts->ba_pending ? (ts->ba_fly ^ 1) : ts->ba_fly;
  TrackState::BaJob& J = ts->job[ts->ba_stage_slot];
  const int start = N - WINDOW;
  J.start = start; J.end = N; J.st = st;
  J.poses.resize(16 * (size_t)WINDOW); J.rel.resize(16 * (size_t)std::max(WINDOW - 1, 0));
  J.oxyz.clear(); J.op.clear(); J.ol.clear(); J.ofeat.clear(); J.pframe.clear(); J.pfeat.clear();
  const int epoch = ++ts->ba_epoch;
  const float invfx = 1.0f / ctx->cfg.fx, invfy = 1.0f / ctx->cfg.fy, cx = ctx->cfg.cx, cy = ctx->cfg.cy;
  int npts = 0;
  // a track enters the window graph iff it is at least 3 long and was born inside the window
  for (int i = start; i < N; i++) {
    MapFrame& F = ts->map[i];
    const int n = (int)F.depth.size();
    for (int j = 0; j < n; j++) {
      const int t = F.track[j];
      if (t < 0) continue;
      TrackInfo& T = ts->tracks[t];
      if (T.len < 3 || T.first_frame < start) continue;
      int pid = (T.epoch == epoch) ? T.pid : -1;
      if (F.pos[j] == 0) {
        pid = npts++;
        T.pid = pid; T.epoch = epoch;
        J.pframe.push_back(i); J.pfeat.push_back(j);
      }
      if (pid < 0) continue;
      const float z = F.depth[j], u = F.xy[2 * j], v = F.xy[2 * j + 1];
      J.op.push_back(i - start); J.ol.push_back(pid); J.ofeat.push_back(j);
      J.oxyz.push_back((u - cx) * z * invfx); J.oxyz.push_back((v - cy) * z * invfy); J.oxyz.push_back(z);
    }
  }
  J.pts.resize(3 * (size_t)npts);
  vido_ba_problem& pr = J.pr;
  memset(&pr, 0, sizeof pr);
  vido_ba_default_params(&pr);
  pr.n_poses = WINDOW; pr.n_points = npts; pr.n_obs = (int)J.op.size();
  pr.poses = J.poses.data(); pr.rel_motion = J.rel.data(); pr.points = J.pts.data();
  pr.obs_pose = J.op.data(); pr.obs_point = J.ol.data(); pr.obs_xyz = J.oxyz.data();
  if (st) { st->ba_points = pr.n_points; st->ba_obs = pr.n_obs; }
  int rc = ba_prepare(ctx, &pr);
  if (rc) return rc;
  ts->ba_staged = true;
  return VIDO_OK;
}

static int ba_go(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->ba_staged) return VIDO_OK;
  ts->ba_staged = false;
  const int slot = ts->ba_stage_slot;
  TrackState::BaJob& J = ts->job[slot];
  const int start = J.start, N = J.end;
  for (int i = start; i < N; i++) {
    const MapFrame& F = ts->map[i];
    memcpy(&J.poses[16 * (size_t)(i - start)], F.Twc, sizeof(float) * 16);
    if (i != start) memcpy(&J.rel[16 * (size_t)(i - start - 1)], F.rel, sizeof(float) * 16);
  }
  const size_t np = J.pframe.size();
  for (size_t l = 0; l < np; l++) {
    const float* q = &ts->map[J.pframe[l]].p3[3 * (size_t)J.pfeat[l]];
    J.pts[3 * l] = q[0]; J.pts[3 * l + 1] = q[1]; J.pts[3 * l + 2] = q[2];
  }
  int rc = ba_launch(ctx, &J.pr, false);
  if (rc) return rc;
  ts->ba_pending = true;
  ts->ba_fly = slot;
  return VIDO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// back-end of one frame (sequential)
// ---------------------------------------------------------------------------------------------------------
static int back_end(vido_ctx* ctx, const FrontFrame& ff, int slot, const vido_keypoint* d_kp, const float* d_depth, const float* d_flow,
                    const int32_t* d_mask, float* Tcw_out, vido_track_stats* st) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const int W = c.width, H = c.height;
  cudaStream_t s = ctx->stream;
  const float invfx = 1.0f / c.fx, invfy = 1.0f / c.fy;
  float curTcw[16];
  eye44(curTcw);
  if (st) { memset(st, 0, sizeof *st); st->n_keypoints = (int)ff.kps.size(); }
  int skipped = 0;
  double t0 = now_ms();
  if (!ts->initialised) {
    // ---- Tracking::Initialization: features leaving frame 0 are the associated detections
    MapFrame F;
    const int m = (int)ff.as_idx.size();
    F.xy.resize(2 * (size_t)m); F.depth = ff.as_depth; F.p3.resize(3 * (size_t)m);
    F.asso.assign(m, -1); F.track.assign(m, -1); F.pos.assign(m, 0);
    for (int i = 0; i < m; i++) {
      const vido_keypoint& kp = ff.kps[ff.as_idx[i]];
      F.xy[2 * i] = kp.x; F.xy[2 * i + 1] = kp.y;
      const float z = F.depth[i];
      F.p3[3 * i] = (kp.x - c.cx) * z * invfx; F.p3[3 * i + 1] = (kp.y - c.cy) * z * invfy; F.p3[3 * i + 2] = z;
    }
    eye44(F.Twc); eye44(F.rel);
    ts->last_keys = F.xy; ts->last_depth = F.depth; ts->last_corres = ff.as_corres; ts->last_flow = ff.as_flow;
    memcpy(ts->lastTcw, curTcw, sizeof curTcw);
    ts->map.push_back(std::move(F));
    ts->initialised = true;
  } else {
    const int Ns = (int)(ts->last_corres.size() / 2);
    if (Ns < 2) skipped = 1;
    else {
      // ---- mvStatKeys = last mvCorres.  The reference also samples the new depth map at these positions into
      //      mvStatDepthTmp (Tracking.cc:369-389); in the static-only pipeline nothing reads that vector (RenewFrameInfo
      //      re-samples at the refined positions, Tracking.cc:2970-3010), so the lookup round trip is not issued.
      std::vector<float> keys = ts->last_corres;
      // ---- GetInitModelCam: 3-D points of the last frame, constant-velocity model, PnP-RANSAC
      std::vector<float> p3d(3 * (size_t)Ns, 0.f);
      std::vector<int32_t> valid(Ns, 1), ids(Ns);
      float Twl[16];
      inv44(ts->lastTcw, Twl);
      for (int i = 0; i < Ns; i++) {
        const float z = ts->last_depth[i];
        if (z < 0) { valid[i] = 0; continue; }
        const float xc[3] = {(ts->last_keys[2 * i] - c.cx) * z * invfx, (ts->last_keys[2 * i + 1] - c.cy) * z * invfy, z};
        for (int r = 0; r < 3; r++)
          p3d[3 * i + r] = (float)((double)Twl[4 * r] * xc[0] + (double)Twl[4 * r + 1] * xc[1] + (double)Twl[4 * r + 2] * xc[2]) + Twl[4 * r + 3];
      }
      vido_pnp_problem pp;
      memset(&pp, 0, sizeof pp);
      vido_pnp_default_params(&pp);
      pp.n = Ns; pp.cur_xy = keys.data(); pp.pts3d = p3d.data(); pp.valid = valid.data(); pp.inlier_ids = ids.data();
      if (ts->has_velocity) mul44(ts->mVelocity, ts->lastTcw, pp.Tcw_motion);
      else memcpy(pp.Tcw_motion, ts->lastTcw, sizeof(float) * 16);
      pp.fx = c.fx; pp.fy = c.fy; pp.cx = c.cx; pp.cy = c.cy;
      int rc = pnp_init_model_host(ctx, &pp);
      if (rc) return rc;
      std::vector<int> TM(ids.begin(), ids.begin() + pp.n_inliers);
      memcpy(curTcw, pp.Tcw_out, sizeof curTcw);
      double t1 = now_ms();
      // ---- PoseOptimizationFlow2Cam
      const int n = (int)TM.size();
      std::vector<float> obs(2 * (size_t)n), fl(2 * (size_t)n), dep(n), fo(2 * (size_t)n);
      std::vector<int32_t> inl(n);
      for (int i = 0; i < n; i++) {
        const int k = TM[i];
        obs[2 * i] = ts->last_keys[2 * k]; obs[2 * i + 1] = ts->last_keys[2 * k + 1];
        fl[2 * i] = ts->last_flow[2 * k]; fl[2 * i + 1] = ts->last_flow[2 * k + 1];
        dep[i] = ts->last_depth[k];
      }
      vido_poseopt_problem po;
      memset(&po, 0, sizeof po);
      vido_poseopt_default_params(&po);
      po.n = n; po.obs_xy = obs.data(); po.flow_xy = fl.data(); po.depth = dep.data();
      memcpy(po.Tcw_init, curTcw, sizeof curTcw);
      memcpy(po.Tcw_last, ts->lastTcw, sizeof curTcw);
      po.fx = c.fx; po.fy = c.fy; po.cx = c.cx; po.cy = c.cy;
      po.flow_out = fo.data(); po.inlier = inl.data();
      rc = po_flow2_host(ctx, &po, 1, nullptr);
      if (rc) return rc;
      memcpy(curTcw, po.Tcw_out, sizeof curTcw);
      if (n >= 3) {
        for (int i = 0; i < n; i++) {
          if (inl[i]) {
            const int k = TM[i];
            keys[2 * k] = (float)((double)ts->last_keys[2 * k] + (double)fo[2 * i]);
            keys[2 * k + 1] = (float)((double)ts->last_keys[2 * k + 1] + (double)fo[2 * i + 1]);
          } else TM[i] = -1;
        }
      }
      double t2 = now_ms();
      // ---- motion model: mVelocity = Tcw * LastTwc
      float LastTwc[16];
      inv44(ts->lastTcw, LastTwc);
      mul44(curTcw, LastTwc, ts->mVelocity);
      ts->has_velocity = true;
      // ---- RenewFrameInfo (static part).  (1) surviving inliers at their refined positions
      MapFrame F;
      std::vector<float> ncorres, nflow;
      const int maxn = c.max_track_bg;
      {
        std::vector<float> qxy;
        std::vector<int> qk;
        for (int i = 0; i < n; i++)
          if (TM[i] != -1) { qxy.push_back(keys[2 * TM[i]]); qxy.push_back(keys[2 * TM[i] + 1]); qk.push_back(TM[i]); }
        const int nq = (int)qk.size();
        std::vector<int32_t> m2(nq);
        std::vector<float> d2(nq), f2(2 * (size_t)nq);
        rc = query_maps(ctx, d_depth, d_flow, d_mask, slot, qxy.data(), nq, m2.data(), d2.data(), f2.data());
        if (rc) return rc;
        for (int i = 0; i < nq; i++) {
          const float px = qxy[2 * i], py = qxy[2 * i + 1];
          const int x = (int)px, y = (int)py;
          bool ok = !(x >= W || y >= H || x <= 0 || y <= 0) && m2[i] == 0 && !(d2[i] > 40 || d2[i] <= 0);
          const float fx = f2[2 * i], fy = f2[2 * i + 1];
          if (ok && fx != 0 && fy != 0 && px + fx < W && py + fy < H && px + fx > 0 && py + fy > 0) {
            F.xy.push_back(px); F.xy.push_back(py);
            ncorres.push_back(px + fx); ncorres.push_back(py + fy);
            nflow.push_back(fx); nflow.push_back(fy);
            F.depth.push_back(d2[i] > 0 ? d2[i] : -1.f);
            F.asso.push_back(qk[i]);
          }
          if ((int)F.asso.size() > maxn) break;
        }
      }
      // (2) top-up from the detected keypoints in 20 interleaved passes, skipping those within 1 px of a kept feature
      int tot = (int)F.asso.size();
      if (tot < maxn && !ff.kps.empty()) {
        const int nk = (int)ff.kps.size(), mcheck = tot;
        std::vector<uint8_t> used(nk, 0);
        if (mcheck > 0) {
          if (mcheck > ts->q_cap) { ctx->err = "renewal check list too long"; return VIDO_ERR_CAPACITY; }
          VIDO_CUDA(cudaMemcpyAsync(ts->d_check, F.xy.data(), 8 * (size_t)mcheck, cudaMemcpyHostToDevice, s));
          topup_used_kernel<<<(nk + 7) / 8, 256, 0, s>>>(d_kp + (size_t)slot * ts->kp_cap, nk, ts->d_check, mcheck, ts->d_used);
          ctx->launches++;
          VIDO_CUDA(cudaMemcpyAsync(used.data(), ts->d_used, nk, cudaMemcpyDeviceToHost, s));
          VIDO_CUDA(cudaStreamSynchronize(s));
        }
        int start_id = 0;
        const int step = 20;
        while (tot < maxn) {
          if (start_id == step) break;
          for (int i = start_id; i < nk; i += step) {
            if (used[i]) continue;
            const float px = ff.kps[i].x, py = ff.kps[i].y;
            const int x = (int)px, y = (int)py;
            if (x >= W || y >= H || x <= 0 || y <= 0) continue;
            if (ff.kp_mask[i] != 0) continue;
            const float d = ff.kp_depth[i];
            if (d > 40 || d <= 0) continue;
            const float fx = ff.kp_flow[2 * i], fy = ff.kp_flow[2 * i + 1];
            if (fx != 0 && fy != 0 && px + fx < W && py + fy < H && px + fx > 0 && py + fy > 0) {
              F.xy.push_back(px); F.xy.push_back(py);
              ncorres.push_back(px + fx); ncorres.push_back(py + fy);
              nflow.push_back(fx); nflow.push_back(fy);
              F.depth.push_back(d);
              F.asso.push_back(-1);
              tot++;
            }
            if (tot >= maxn) break;
          }
          start_id++;
        }
      }
      // (3)(4) world points through the current pose (Optimizer::Get3DinWorld)
      const int nf = (int)F.asso.size();
      float Twc[16];
      inv44(curTcw, Twc);
      F.p3.resize(3 * (size_t)nf);
      for (int i = 0; i < nf; i++) {
        const float z = F.depth[i];
        const float xc[3] = {(F.xy[2 * i] - c.cx) * z * invfx, (F.xy[2 * i + 1] - c.cy) * z * invfy, z};
        for (int r = 0; r < 3; r++)
          F.p3[3 * i + r] = (float)((double)Twc[4 * r] * xc[0] + (double)Twc[4 * r + 1] * xc[1] + (double)Twc[4 * r + 2] * xc[2]) + Twc[4 * r + 3];
      }
      // ---- tracklets, incrementally (same chains as Tracking::GetStaticTrack)
      F.track.assign(nf, -1); F.pos.assign(nf, 0);
      MapFrame& P = ts->map.back();
      const int fcur = (int)ts->map.size();
      for (int j = 0; j < nf; j++) {
        const int p = F.asso[j];
        if (p < 0) continue;
        if (P.track[p] >= 0) {
          TrackInfo& T = ts->tracks[P.track[p]];
          F.track[j] = P.track[p];
          F.pos[j] = T.len;
          T.len++;
        } else {
          const int t = (int)ts->tracks.size();
          ts->tracks.push_back({fcur - 1, 2, -1, -1});
          P.track[p] = t; P.pos[p] = 0;
          F.track[j] = t; F.pos[j] = 1;
        }
      }
      memcpy(F.Twc, Twc, sizeof Twc);
      inv44(ts->mVelocity, F.rel);
      ts->last_keys = F.xy; ts->last_depth = F.depth; ts->last_corres = ncorres; ts->last_flow = nflow;
      memcpy(ts->lastTcw, curTcw, sizeof curTcw);
      ts->map.push_back(std::move(F));
      double t3 = now_ms();
      if (st) {
        st->ms_init = t1 - t0; st->ms_poseopt = t2 - t1; st->ms_renew = t3 - t2;
        st->n_matches = Ns; st->n_init_inliers = pp.n_inliers; st->init_winner = pp.winner; st->n_pose_inliers = po.n_inliers;
        st->n_static = nf;
      }
    }
  }
  memcpy(Tcw_out, curTcw, sizeof(float) * 16);
  double t4 = now_ms();
  const int window = ts->f_id < c.window_size ? ts->f_id : c.window_size;
  int rc = skipped ? VIDO_OK : ba_stage(ctx, window, st);  // structure of this frame's window, staged while the previous
  if (rc) return rc;                                        // frame's window is still being solved
  rc = ba_finish(ctx);
  if (rc) return rc;
  rc = ba_go(ctx);
  ba_writeback_rest(ctx);  // off the critical path: the next solve is already running
  if (st) st->ms_ba = now_ms() - t4;
  ts->f_id++;
  if (rc) return rc;
  return skipped ? 1 : 0;
}


// Host->device copy done by SMs (the source is pinned host memory, directly addressable under UVA).  The prefetch of a
// whole chunk is ~140 MB; issued as cudaMemcpyAsync it would sit in the copy-engine queue in front of the small,
// latency-critical copies of the back-end (PnP / pose-opt / BA inputs) and delay each of them by up to one DMA
// descriptor.  A few CTAs streaming it keep the DMA queues free.  src and dst must be congruent modulo 16.
__global__ void __launch_bounds__(256) h2d_stream_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes) {
  const size_t head = (16 - ((size_t)src & 15)) & 15;
  const size_t h = head < bytes ? head : bytes;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gn = (size_t)gridDim.x * blockDim.x;
  if (gtid < h) dst[gtid] = src[gtid];
  const size_t body = (bytes - h) / 16;
  const uint4* s4 = (const uint4*)(src + h);
  uint4* d4 = (uint4*)(dst + h);
  size_t i = gtid;
  for (; i + 3 * gn < body; i += 4 * gn) {  // four independent 16-byte reads in flight per thread
    const uint4 a = s4[i], b = s4[i + gn], c = s4[i + 2 * gn], d = s4[i + 3 * gn];
    d4[i] = a; d4[i + gn] = b; d4[i + 2 * gn] = c; d4[i + 3 * gn] = d;
  }
  for (; i < body; i += gn) d4[i] = s4[i];
  const size_t tail0 = h + body * 16;
  if (tail0 + gtid < bytes) dst[tail0 + gtid] = src[tail0 + gtid];
}

static bool is_pinned_host(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

// one array of one frame onto the copy stream: SM copy when possible, DMA otherwise
static int prefetch_array(vido_ctx* ctx, cudaStream_t cs, void* dst, const void* src, size_t bytes) {
  if ((((size_t)dst ^ (size_t)src) & 15) == 0 && is_pinned_host(src)) {
    h2d_stream_kernel<<<8, 256, 0, cs>>>((uint8_t*)dst, (const uint8_t*)src, bytes);
    ctx->launches++;
    VIDO_CUDA(cudaGetLastError());
  } else {
    VIDO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs));
  }
  return VIDO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// vido_track_prefetch: remember the frames the NEXT vido_track_frames call will start with.  While the back-end of the
// current call works through its last batch, their copy and front-end already run in the idle pipeline slot.
int trk_prefetch(vido_ctx* ctx, const vido_frame_inputs* in, int nframes) {
  TrackState* ts = (TrackState*)ctx->trk;
  ts->hint.clear();
  if (nframes <= 0) return VIDO_OK;
  const int B = std::min(ts->capB, nframes);
  ts->hint.assign(in, in + B);
  return VIDO_OK;
}

int trk_track_chunk(vido_ctx* ctx, const vido_frame_inputs* in, int nframes, float* Tcw_out, vido_track_stats* stats) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  cudaStream_t s = ctx->stream;
  const size_t px = (size_t)c.width * c.height;
  if (ts->ba_pending) { vido_lm_stats ls; ba_collect(ctx, &ts->job[ts->ba_fly].pr, &ls); ts->ba_pending = false; }  // left by a failed call
  int done = 0;
  while (done < nframes) {
    const int B = std::min(ts->capB, nframes - done);
    double tf0 = now_ms();
    TrackState::FeSlot& F = ts->fe[ts->fe_cur];
    if (!(F.launched && F.key == (const void*)in[done].image && F.B == B && F.channels == in[done].channels)) {
      int rc = fe_launch(ctx, F, in + done, B);  // not announced (first batch of a sequence, or the hint did not match)
      if (rc) return rc;
    }
    std::vector<FrontFrame> ff;
    int rc = fe_collect(ctx, F, ff);
    if (rc) return rc;
    // look-ahead: the next batch of this call, or the announced first batch of the next call, goes through copy and
    // front-end in the other slot while this batch's back-end runs
    {
      TrackState::FeSlot& N = ts->fe[ts->fe_cur ^ 1];
      N.launched = false;
      if (done + B < nframes) rc = fe_launch(ctx, N, in + done + B, std::min(ts->capB, nframes - done - B));
      else if (!ts->hint.empty()) { rc = fe_launch(ctx, N, ts->hint.data(), (int)ts->hint.size()); ts->hint.clear(); }
      if (rc) return rc;
    }
    const float* d_depth = F.in_depth; const float* d_flow = F.in_flow; const int32_t* d_mask = F.in_mask;
    const double front_ms = (now_ms() - tf0) / B;
    for (int b = 0; b < B; b++) {
      vido_track_stats* st = stats ? stats + done + b : nullptr;
      rc = back_end(ctx, ff[b], b, F.d_kp, d_depth, d_flow, d_mask, Tcw_out + 16 * (size_t)(done + b), st);
      if (rc < 0) return rc;
      if (st) { st->ms_orb = front_ms; st->ms_assoc = 0; }
      // the reference pre-scales the caller's depth map in place (Tracking.cc:299-322): reproduce on request
      const vido_frame_inputs& f = in[done + b];
      if (f.write_back_depth) {
        rc = assoc_depth_prep(ctx, (float*)d_depth + (size_t)b * px, 1, px, c.width);
        if (rc) return rc;
        if (!f.on_device) VIDO_CUDA(cudaMemcpyAsync((void*)f.depth, d_depth + (size_t)b * px, px * 4, cudaMemcpyDeviceToHost, s));
        else if (d_depth != f.depth) VIDO_CUDA(cudaMemcpyAsync((void*)f.depth, d_depth + (size_t)b * px, px * 4, cudaMemcpyDeviceToDevice, s));
        VIDO_CUDA(cudaStreamSynchronize(s));
      }
    }
    ts->fe_cur ^= 1;
    done += B;
  }
  {
    int rc = ba_finish(ctx);  // drain: stats and map are final when the call returns
    ba_writeback_rest(ctx);
    return rc;
  }
}

// the ORB workspace is shared with the stand-alone extraction entry points: wait for a front-end running ahead
void trk_quiesce(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (ts && ts->fe_stream) cudaStreamSynchronize(ts->fe_stream);
}

int trk_num_frames(vido_ctx* ctx) { return (int)((TrackState*)ctx->trk)->map.size(); }

int trk_get_map_poses(vido_ctx* ctx, float* poses, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  const int n = (int)ts->map.size();
  for (int i = 0; i < n && i < cap; i++) memcpy(poses + 16 * (size_t)i, ts->map[i].Twc, sizeof(float) * 16);
  return n;
}

int trk_get_static(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (frame < 0 || frame >= (int)ts->map.size()) return -1;
  const MapFrame& F = ts->map[frame];
  const int n = (int)F.depth.size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = F.xy[2 * i]; xy[2 * i + 1] = F.xy[2 * i + 1];
    depth[i] = F.depth[i];
    p3[3 * i] = F.p3[3 * i]; p3[3 * i + 1] = F.p3[3 * i + 1]; p3[3 * i + 2] = F.p3[3 * i + 2];
    asso[i] = F.asso[i];
  }
  return n;
}
