// chain_kernels.cu -- the static (VO) back-end of one frame as a chain of kernels on the context stream, with the tracker
// state resident on the device: nothing between the front-end and the Map record needs the host, so the kernels of frame
// k+1 are queued behind those of frame k and the host only consumes finished records (Map bookkeeping, window-BA staging).
//
// Reference code replaced (src/Tracking.cc, static part of Tracking::Track):
//   GetInitModelCam prologue + cv::solvePnPRansac + motion-model test   :1914-2028   pnp_chain_enqueue (pnp_kernels.cu)
//   PoseOptimizationFlow2Cam and its inputs                              :1137-1160   po_chain_enqueue (poseopt_kernels.cu)
//   refined key points of the inliers, outliers dropped                  :1160-1180   chain_survive_kernel
//   mVelocity = Tcw * LastTwc                                            :1320-1330   chain_finish_kernel
//   RenewFrameInfo, static part (survivors; top-up from the detections in 20 interleaved passes, skipping detections within
//   1 px of a survivor; world points)                                    :2959-3110   chain_survive / chain_used / chain_finish
// Every float operation is spelled with the rounding the host path (csrc/track.cu back_end) and the oracle use: the chain
// produces the same bits as the host-driven path (tests/test_track_gpu.py compares both with the oracle).
#include <algorithm>
#include <cstring>
#include <vector>

#include "ctx.h"

struct ChainConv { int mode; float factor, bf, mscale; };

__device__ __forceinline__ float chain_conv_depth(float d, const ChainConv& c) {   // Tracking.cc:299-322 (see assoc_kernels.cu)
  if (c.mode == 0) return d;
  if (d < 0) return 0.f;
  if (c.mode == 1) return __fdiv_rn(d, c.factor);
  if (c.mode == 2) return __fdiv_rn(c.bf, __fdiv_rn(d, c.factor));
  return __fdiv_rn(__fmul_rn(c.mscale, c.bf), __fdiv_rn(d, c.factor));
}

// record of one frame as the host consumes it (device block, copied to a pinned mirror)
#define CHAIN_HDR 16   // status (0 tracked, 1 skipped), Ns, init inliers, winner, ransac inliers, mm inliers, pose inliers, nf, survivors, nkp
struct ChainRecPtr {
  int32_t* hdr; float *Tcw, *Twc, *rel, *vel, *xy, *depth, *p3, *corres, *flow; int32_t* asso;
};
static size_t chain_rec_bytes(int cap) { return 4 * (CHAIN_HDR + 64) + (size_t)cap * 4 * (2 + 1 + 3 + 2 + 2 + 1); }
__host__ __device__ static inline ChainRecPtr chain_rec(char* base, int cap) {
  ChainRecPtr r;
  r.hdr = (int32_t*)base; r.Tcw = (float*)(r.hdr + CHAIN_HDR); r.Twc = r.Tcw + 16; r.rel = r.Twc + 16; r.vel = r.rel + 16;
  r.xy = r.vel + 16; r.depth = r.xy + 2 * (size_t)cap; r.p3 = r.depth + cap; r.corres = r.p3 + 3 * (size_t)cap;
  r.flow = r.corres + 2 * (size_t)cap; r.asso = (int32_t*)(r.flow + 2 * (size_t)cap);
  return r;
}

struct ChainArgs {
  int W, H, maxn, cap, kp_cap;
  float fx, fy, cx, cy;
  ChainConv conv;
  // state
  ChainStateDev cur, nxt;
  // init model / pose optimisation outputs
  const int32_t* pnp_res; const int32_t* pnp_ids; const float* pnp_T;
  const int32_t* po_n; const int32_t* po_inl; const int32_t* po_ninl; const float* po_T; const float* po_flow;
  // frame inputs
  const vido_keypoint* kp; const int32_t* nkp; const int32_t* kpmask; const float* kpdepth; const float* kpflow;
  const float* depth; const float* flow; const int32_t* mask;
  // scratch / outputs
  int32_t* tot;      // survivors kept
  uint8_t* used;     // [kp_cap]
  char* rec;
};

// block-wide ordered append (1024 threads): returns the slot of a flagged thread, -1 otherwise; advances *s_base
__device__ __forceinline__ int chain_slot(bool flag, int* s_warp, int* s_base) {
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
    for (int w = 0; w < 32; w++) t += s_warp[w];
    *s_base += t;
  }
  __syncthreads();
  return slot;
}

// ---- survivors: inliers of the pose optimisation at their refined positions, re-sampled in the new maps
__global__ void __launch_bounds__(1024) chain_survive_kernel(ChainArgs a) {
  __shared__ int s_warp[32], s_base;
  const int tid = threadIdx.x;
  ChainRecPtr R = chain_rec(a.rec, a.cap);
  if (tid == 0) s_base = 0;
  __syncthreads();
  const bool skip = a.cur.hdr[2] != 0;
  const int n = skip ? 0 : a.po_n[0];
  const bool n3 = n >= 3;
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    bool keep = false;
    float px = 0, py = 0, fx = 0, fy = 0, d2 = 0;
    int k = -1;
    if (i < n && (!n3 || a.po_inl[i])) {
      k = a.pnp_ids[i];
      if (n3) {
        px = (float)((double)a.cur.keys[2 * k] + (double)a.po_flow[2 * i]);
        py = (float)((double)a.cur.keys[2 * k + 1] + (double)a.po_flow[2 * i + 1]);
      } else { px = a.cur.corres[2 * k]; py = a.cur.corres[2 * k + 1]; }
      const int x = (int)px, y = (int)py;
      int m2 = -1;
      if (!(x < 0 || y < 0 || x >= a.W || y >= a.H)) {
        const size_t q = (size_t)y * a.W + x;
        m2 = a.mask[q];
        d2 = chain_conv_depth(a.depth[q], a.conv);
        fx = a.flow[2 * q]; fy = a.flow[2 * q + 1];
      }
      const bool ok = !(x >= a.W || y >= a.H || x <= 0 || y <= 0) && m2 == 0 && !(d2 > 40 || d2 <= 0);
      keep = ok && fx != 0 && fy != 0 && __fadd_rn(px, fx) < (float)a.W && __fadd_rn(py, fy) < (float)a.H && __fadd_rn(px, fx) > 0 &&
             __fadd_rn(py, fy) > 0;
    }
    const int slot = chain_slot(keep, s_warp, &s_base);
    if (slot >= 0 && slot <= a.maxn) {   // the reference stops once the list holds maxn + 1 features
      R.xy[2 * slot] = px; R.xy[2 * slot + 1] = py;
      R.corres[2 * slot] = __fadd_rn(px, fx); R.corres[2 * slot + 1] = __fadd_rn(py, fy);
      R.flow[2 * slot] = fx; R.flow[2 * slot + 1] = fy;
      R.depth[slot] = d2 > 0 ? d2 : -1.f;
      R.asso[slot] = k;
    }
  }
  if (tid == 0) *a.tot = min(s_base, a.maxn + 1);
}

// ---- used[i] = detection i lies within 1 px of a survivor (Tracking.cc:3030-3040); one warp per detection
__global__ void __launch_bounds__(256) chain_used_kernel(ChainArgs a) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const int n = min(a.nkp[0], a.kp_cap), m = *a.tot;
  if (i >= n) return;
  ChainRecPtr R = chain_rec(a.rec, a.cap);
  const float sx = a.kp[i].x, sy = a.kp[i].y;
  bool u = false;
  if (m < a.maxn) {
    for (int j0 = 0; j0 < m; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      if (j < m) {
        const float dx = __fsub_rn(R.xy[2 * j], sx), dy = __fsub_rn(R.xy[2 * j + 1], sy);
        hit = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) < 1.0f;
      }
      if (__any_sync(0xffffffffu, hit)) { u = true; break; }
    }
  }
  if (lane == 0) a.used[i] = u ? 1 : 0;
}

// ---- top-up, world points, motion model, next state, record
__device__ void chain_inv44(const float* T, float* o) {   // track.cu inv44
  for (int k = 0; k < 16; k++) o[k] = 0.f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[4 * r + c] = T[4 * c + r];
  for (int r = 0; r < 3; r++) {
    double s = 0;
    for (int k = 0; k < 3; k++) s += (double)(-o[4 * r + k]) * (double)T[4 * k + 3];
    o[4 * r + 3] = (float)s;
  }
  o[15] = 1.f;
}
__device__ void chain_mul44(const float* A, const float* B, float* C) {   // track.cu mul44
  float o[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += (double)A[4 * r + k] * (double)B[4 * k + c];
      o[4 * r + c] = (float)s;
    }
  for (int k = 0; k < 16; k++) C[k] = o[k];
}

__global__ void __launch_bounds__(1024) chain_finish_kernel(ChainArgs a) {
  __shared__ int s_warp[32], s_base, s_cnt[20], s_off[21];
  __shared__ float curT[16], Twc[16], vel[16], rel[16];
  const int tid = threadIdx.x;
  ChainRecPtr R = chain_rec(a.rec, a.cap);
  const bool skip = a.cur.hdr[2] != 0;
  if (skip) {
    // lost tracking (fewer than two features): nothing is processed, the state is carried over, TrackRGBD returns identity
    const int n = a.cur.hdr[0];
    for (int i = tid; i < n; i += 1024) {
      a.nxt.keys[2 * i] = a.cur.keys[2 * i]; a.nxt.keys[2 * i + 1] = a.cur.keys[2 * i + 1];
      a.nxt.corres[2 * i] = a.cur.corres[2 * i]; a.nxt.corres[2 * i + 1] = a.cur.corres[2 * i + 1];
      a.nxt.flow[2 * i] = a.cur.flow[2 * i]; a.nxt.flow[2 * i + 1] = a.cur.flow[2 * i + 1];
      a.nxt.depth[i] = a.cur.depth[i];
    }
    if (tid < 16) {
      a.nxt.Tcw[tid] = a.cur.Tcw[tid]; a.nxt.vel[tid] = a.cur.vel[tid];
      R.Tcw[tid] = (tid % 5 == 0) ? 1.f : 0.f;
    }
    if (tid == 0) {
      a.nxt.hdr[0] = n; a.nxt.hdr[1] = a.cur.hdr[1]; a.nxt.hdr[2] = 0;
      for (int k = 0; k < CHAIN_HDR; k++) R.hdr[k] = 0;
      R.hdr[0] = 1; R.hdr[1] = n;
    }
    return;
  }
  const int tot = *a.tot, nk = min(a.nkp[0], a.kp_cap);
  if (tid == 0) {
    for (int k = 0; k < 16; k++) curT[k] = a.po_T[k];
    chain_inv44(curT, Twc);
    float LastTwc[16];
    chain_inv44(a.cur.Tcw, LastTwc);
    chain_mul44(curT, LastTwc, vel);
    chain_inv44(vel, rel);
    s_base = 0;
    int run = 0;
    for (int s = 0; s < 20; s++) { s_off[s] = run; s_cnt[s] = (nk > s) ? (nk - s + 19) / 20 : 0; run += s_cnt[s]; }
    s_off[20] = run;
  }
  __syncthreads();
  int nf = tot;
  if (tot < a.maxn && nk > 0) {
    const int need = a.maxn - tot;
    // detections in the reference's visiting order: pass s = 0..19 takes i = s, s + 20, s + 40, ...
    for (int p0 = 0; p0 < nk; p0 += 1024) {
      const int p = p0 + tid;
      bool pass = false;
      int i = 0;
      float px = 0, py = 0, fx = 0, fy = 0, d = 0;
      if (p < nk) {
        int s = 0;
        while (s < 19 && s_off[s + 1] <= p) s++;
        i = s + 20 * (p - s_off[s]);
        px = a.kp[i].x; py = a.kp[i].y;
        const int x = (int)px, y = (int)py;
        d = a.kpdepth[i];
        fx = a.kpflow[2 * i]; fy = a.kpflow[2 * i + 1];
        pass = !a.used[i] && !(x >= a.W || y >= a.H || x <= 0 || y <= 0) && a.kpmask[i] == 0 && !(d > 40 || d <= 0) && fx != 0 && fy != 0 &&
               __fadd_rn(px, fx) < (float)a.W && __fadd_rn(py, fy) < (float)a.H && __fadd_rn(px, fx) > 0 && __fadd_rn(py, fy) > 0;
      }
      const int rank = chain_slot(pass, s_warp, &s_base);
      if (rank >= 0 && rank < need) {
        const int slot = tot + rank;
        R.xy[2 * slot] = px; R.xy[2 * slot + 1] = py;
        R.corres[2 * slot] = __fadd_rn(px, fx); R.corres[2 * slot + 1] = __fadd_rn(py, fy);
        R.flow[2 * slot] = fx; R.flow[2 * slot + 1] = fy;
        R.depth[slot] = d;
        R.asso[slot] = -1;
      }
      if (s_base >= need) break;   // uniform: s_base is shared
    }
    nf = tot + min(s_base, need);
  }
  __syncthreads();
  // world points through the current pose (Optimizer::Get3DinWorld), next state
  const float invfx = __fdiv_rn(1.0f, a.fx), invfy = __fdiv_rn(1.0f, a.fy);
  for (int i = tid; i < nf; i += 1024) {
    const float z = R.depth[i], u = R.xy[2 * i], v = R.xy[2 * i + 1];
    const float xc[3] = {__fmul_rn(__fmul_rn(__fsub_rn(u, a.cx), z), invfx), __fmul_rn(__fmul_rn(__fsub_rn(v, a.cy), z), invfy), z};
    for (int r = 0; r < 3; r++)
      R.p3[3 * i + r] = __fadd_rn((float)((double)Twc[4 * r] * xc[0] + (double)Twc[4 * r + 1] * xc[1] + (double)Twc[4 * r + 2] * xc[2]), Twc[4 * r + 3]);
    a.nxt.keys[2 * i] = u; a.nxt.keys[2 * i + 1] = v;
    a.nxt.depth[i] = z;
    a.nxt.corres[2 * i] = R.corres[2 * i]; a.nxt.corres[2 * i + 1] = R.corres[2 * i + 1];
    a.nxt.flow[2 * i] = R.flow[2 * i]; a.nxt.flow[2 * i + 1] = R.flow[2 * i + 1];
  }
  if (tid < 16) {
    a.nxt.Tcw[tid] = curT[tid]; a.nxt.vel[tid] = vel[tid];
    R.Tcw[tid] = curT[tid]; R.Twc[tid] = Twc[tid]; R.rel[tid] = rel[tid]; R.vel[tid] = vel[tid];
  }
  if (tid == 0) {
    a.nxt.hdr[0] = nf; a.nxt.hdr[1] = 1; a.nxt.hdr[2] = 0;
    R.hdr[0] = 0; R.hdr[1] = a.cur.hdr[0]; R.hdr[2] = a.pnp_res[0]; R.hdr[3] = a.pnp_res[1]; R.hdr[4] = a.pnp_res[2]; R.hdr[5] = a.pnp_res[3];
    R.hdr[6] = a.po_ninl[0]; R.hdr[7] = nf; R.hdr[8] = tot; R.hdr[9] = nk;
  }
}

// =========================================================================================================
// host side
// =========================================================================================================
struct ChainWorkspace {
  int cap = 0, nslots = 0;
  char* d_state = nullptr;
  ChainStateDev st[2];
  int cur = 0;                    // state buffer that holds the last processed frame
  char* d_rec = nullptr; char* h_rec = nullptr;
  size_t rec_bytes = 0;
  int32_t* d_tot = nullptr;
  uint8_t* d_used = nullptr;
  std::vector<cudaEvent_t> ev;    // per slot: start, after init model, after pose optimisation, record copied
};

static ChainWorkspace* g_chain(vido_ctx* ctx) { return (ChainWorkspace*)ctx->chain; }

int chain_setup(vido_ctx* ctx, int nslots) {
  ChainWorkspace* ws = new ChainWorkspace();
  ctx->chain = ws;
  const int cap = ((ctx->cfg.max_track_bg + 8) + 7) & ~7;
  ws->cap = cap; ws->nslots = nslots;
  int rc = pnp_chain_setup(ctx, cap);
  if (rc) return rc;
  rc = po_chain_setup(ctx, cap);
  if (rc) return rc;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_hdr = 0, o_T = 256, o_vel = 512, o_keys = 768, o_depth = o_keys + al(8 * (size_t)cap), o_cor = o_depth + al(4 * (size_t)cap),
               o_flow = o_cor + al(8 * (size_t)cap), one = o_flow + al(8 * (size_t)cap);
  VIDO_CUDA(cudaMalloc(&ws->d_state, 2 * one));
  VIDO_CUDA(cudaMemset(ws->d_state, 0, 2 * one));
  for (int k = 0; k < 2; k++) {
    char* b = ws->d_state + k * one;
    ws->st[k].hdr = (int32_t*)(b + o_hdr); ws->st[k].Tcw = (float*)(b + o_T); ws->st[k].vel = (float*)(b + o_vel);
    ws->st[k].keys = (float*)(b + o_keys); ws->st[k].depth = (float*)(b + o_depth); ws->st[k].corres = (float*)(b + o_cor);
    ws->st[k].flow = (float*)(b + o_flow);
  }
  ws->rec_bytes = al(chain_rec_bytes(cap));
  VIDO_CUDA(cudaMalloc(&ws->d_rec, ws->rec_bytes * nslots));
  VIDO_CUDA(cudaMallocHost(&ws->h_rec, ws->rec_bytes * nslots));
  VIDO_CUDA(cudaMalloc(&ws->d_tot, 256));
  VIDO_CUDA(cudaMalloc(&ws->d_used, (size_t)ctx->kp_cap + 64));
  ws->ev.resize(4 * (size_t)nslots);
  for (auto& e : ws->ev) VIDO_CUDA(cudaEventCreate(&e));
  return VIDO_OK;
}

void chain_teardown(vido_ctx* ctx) {
  ChainWorkspace* ws = g_chain(ctx);
  if (!ws) return;
  for (auto& e : ws->ev) if (e) cudaEventDestroy(e);
  cudaFree(ws->d_state); cudaFree(ws->d_rec); cudaFreeHost(ws->h_rec); cudaFree(ws->d_tot); cudaFree(ws->d_used);
  delete ws;
  ctx->chain = nullptr;
}

int chain_capacity(vido_ctx* ctx) { return g_chain(ctx)->cap; }

// host state -> device (entering the chained mode): mpLastFrame's static features and the motion model
int chain_upload_state(vido_ctx* ctx, int n, const float* keys, const float* depth, const float* corres, const float* flow, const float* Tcw,
                       const float* vel, int has_velocity) {
  ChainWorkspace* ws = g_chain(ctx);
  if (n > ws->cap) { ctx->err = "tracker state exceeds the chain capacity"; return VIDO_ERR_CAPACITY; }
  cudaStream_t s = ctx->stream;
  const ChainStateDev& d = ws->st[ws->cur];
  const int32_t hdr[4] = {n, has_velocity, 0, 0};
  VIDO_CUDA(cudaMemcpyAsync(d.hdr, hdr, sizeof hdr, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(d.Tcw, Tcw, 64, cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync(d.vel, vel, 64, cudaMemcpyHostToDevice, s));
  if (n) {
    VIDO_CUDA(cudaMemcpyAsync(d.keys, keys, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(d.depth, depth, 4 * (size_t)n, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(d.corres, corres, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(d.flow, flow, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
  }
  VIDO_CUDA(cudaStreamSynchronize(s));   // the sources are pageable host vectors of the caller
  return VIDO_OK;
}

// queue the back-end of one frame (no synchronisation).  The frame's front-end results and maps are device pointers already
// offset to the frame; `slot` selects the record / event set.
int chain_enqueue_frame(vido_ctx* ctx, const vido_keypoint* kp, const int32_t* nkp, const int32_t* kpmask, const float* kpdepth,
                        const float* kpflow, const float* depth, const float* flow, const int32_t* mask, int slot) {
  ChainWorkspace* ws = g_chain(ctx);
  cudaStream_t s = ctx->stream;
  const vido_config& c = ctx->cfg;
  cudaEvent_t* ev = &ws->ev[4 * (size_t)slot];
  const ChainStateDev& cur = ws->st[ws->cur];
  const ChainStateDev& nxt = ws->st[ws->cur ^ 1];
  cudaEventRecord(ev[0], s);
  ChainPnpOut pn;
  int rc = pnp_chain_enqueue(ctx, cur, ws->cur, &pn);
  if (rc) return rc;
  cudaEventRecord(ev[1], s);
  ChainPoOut po;
  rc = po_chain_enqueue(ctx, cur, ws->cur, pn, &po);
  if (rc) return rc;
  cudaEventRecord(ev[2], s);
  ChainArgs a;
  memset(&a, 0, sizeof a);
  a.W = c.width; a.H = c.height; a.maxn = c.max_track_bg; a.cap = ws->cap; a.kp_cap = ctx->kp_cap;
  a.fx = c.fx; a.fy = c.fy; a.cx = c.cx; a.cy = c.cy;
  a.conv.mode = c.choose_data; a.conv.factor = c.depth_map_factor; a.conv.bf = c.bf; a.conv.mscale = ctx->mscale;
  a.cur = cur; a.nxt = nxt;
  a.pnp_res = pn.res; a.pnp_ids = pn.ids; a.pnp_T = pn.T;
  a.po_n = po.n; a.po_inl = po.inl; a.po_ninl = po.ninl; a.po_T = po.T; a.po_flow = po.flow;
  a.kp = kp; a.nkp = nkp; a.kpmask = kpmask; a.kpdepth = kpdepth; a.kpflow = kpflow;
  a.depth = depth; a.flow = flow; a.mask = mask;
  a.tot = ws->d_tot; a.used = ws->d_used;
  a.rec = ws->d_rec + ws->rec_bytes * slot;
  chain_survive_kernel<<<1, 1024, 0, s>>>(a);
  chain_used_kernel<<<(ctx->kp_cap + 7) / 8, 256, 0, s>>>(a);
  chain_finish_kernel<<<1, 1024, 0, s>>>(a);
  ctx->launches += 3;
  VIDO_CUDA(cudaGetLastError());
  VIDO_CUDA(cudaMemcpyAsync(ws->h_rec + ws->rec_bytes * slot, a.rec, chain_rec_bytes(ws->cap), cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaEventRecord(ev[3], s));
  ws->cur ^= 1;
  return VIDO_OK;
}

// wait for the record of `slot`; pointers into the pinned mirror (valid until the slot is reused)
int chain_wait_record(vido_ctx* ctx, int slot, const int32_t** hdr, const float** Tcw, const float** Twc, const float** rel, const float** vel, const float** xy,
                      const float** depth, const float** p3, const float** corres, const float** flow, const int32_t** asso) {
  ChainWorkspace* ws = g_chain(ctx);
  cudaEvent_t* ev = &ws->ev[4 * (size_t)slot];
  VIDO_CUDA(cudaEventSynchronize(ev[3]));
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) { ctx->t_ms[1] += ms; ctx->t_n[1]++; }
  if (cudaEventElapsedTime(&ms, ev[1], ev[2]) == cudaSuccess) { ctx->t_ms[2] += ms; ctx->t_n[2]++; }
  ChainRecPtr R = chain_rec(ws->h_rec + ws->rec_bytes * slot, ws->cap);
  *hdr = R.hdr; *Tcw = R.Tcw; *Twc = R.Twc; *rel = R.rel; *vel = R.vel; *xy = R.xy; *depth = R.depth; *p3 = R.p3; *corres = R.corres; *flow = R.flow; *asso = R.asso;
  return VIDO_OK;
}
