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
// Dynamic objects (non-zero mask): Tracking::UpdateMask Tracking.cc:3291-3357, object carry-over :391-421,
//   GetSceneFlowObj :1582-1668, DynObjTracking :1670-1912, GetInitModelObj :2030-2162 + Optimizer::PoseOptimizationFlow2
//   Optimizer.cc:3037-3253 (one launch for all objects of a frame), object part of RenewFrameInfo :3112-3289,
//   GetDynamicTrackNew :2615-2720 (incremental).  A static sequence (all-zero mask) never enters these branches.
// VIO mode (sensor = IMU_RGBD, after vido_track_set_imu): Tracking::ParseIMUParamFile / GrabImuData / PreintegrateIMU (the
//   preintegration kernel, one launch per front-end batch) / InitializeIMU + Optimizer::InertialOptimization (the inertial
//   kernel) / ScaleRefinement / UpdateFrameIMU  Tracking.cc:174-281, 784-1077, 1115-1119, 1452-1480, 1555-1561; the IMU state of
//   Frame  Frame.cc:437-521, Frame.h:44-110; Map::ApplyScaledRotation  Map.cc:55-119.
// Scope of this version: sensor RGBD / IMU_RGBD, UseSampleFeature = 0.
// float 4x4 products use double accumulation + one rounding like cv::Mat CV_32F gemm.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <thread>

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

struct ObjEntry { int label, sem; float motion[16]; float centre[3]; float motion_rf[16]; };  // vnRMLabel / vnSMLabel / vmRigidMotion[f-1][j>=1] / vmRigidCentre
struct DynTrack { int first_frame, first_feat, len, obj_id; };          // TrackletDyn / nObjID
struct MapFrame {           // Map::vpFeatSta / vfDepSta / vp3DPointSta / vnAssoSta of one frame + its pose
  std::vector<float> xy, depth, p3;
  std::vector<int> asso, track, pos;
  float Twc[16];            // vmCameraPose
  float rel[16];            // vmRigidMotion[f-1][0]
  float Twc_rf[16];         // vmCameraPose_RF (refined by the full-sequence optimisation)
  // Map::vpFeatDyn / vfDepDyn / vp3DPointDyn / vnAssoDyn / vnFeatLabel + the objects with an estimated motion
  std::vector<float> dxy, ddepth, dp3;
  std::vector<int> dasso, dlabel, dtrack;
  std::vector<ObjEntry> objects;
};

// ---- float32 cv::Mat helpers of the VIO glue (gemm: double accumulation, one rounding per expression)
void mm3(const float* A, const float* B, float* C) {
  float o[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = (float)((double)A[3 * r] * B[c] + (double)A[3 * r + 1] * B[3 + c] + (double)A[3 * r + 2] * B[6 + c]);
  memcpy(C, o, sizeof o);
}
void mv3(const float* A, const float* v, float* o, double alpha = 1.0, const float* w = nullptr) {   // alpha * A * v + w
  float t[3];
  for (int r = 0; r < 3; r++)
    t[r] = (float)(alpha * ((double)A[3 * r] * v[0] + (double)A[3 * r + 1] * v[1] + (double)A[3 * r + 2] * v[2]) + (w ? (double)w[r] : 0.0));
  memcpy(o, t, sizeof t);
}
void inv33d(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double id = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// IMU::ExpSO3(float) (ImuTypes.cc:38-50), entries rounded to float32
void exp_so3_host(const float* w, float* R) {
  const float x = w[0], y = w[1], z = w[2];
  const float d2 = x * x + y * y + z * z;
  const float d = std::sqrt(d2);
  const double W[9] = {0, -z, y, z, 0, -x, -y, x, 0};
  for (int k = 0; k < 9; k++) {
    const int i = k / 3, j = k % 3;
    const double w2 = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
    const double I = (i == j) ? 1.0 : 0.0;
    const double v = d < 1e-4f ? I + W[k] + 0.5 * (double)(float)w2
                               : I + W[k] * std::sin((double)d) / d + (double)(float)w2 * (1.0 - std::cos((double)d)) / d2;
    R[k] = (float)v;
  }
}
// IMU::Preintegrated::GetUpdatedDeltaRotation / Velocity / Position (ImuTypes.cc:370-386): db = (gyro, acc) bias deltas;
// NormalizeRotation (cv::SVDecomp U * Vt) = orthogonal polar factor
void updated_deltas(const vido_imu_preint& p, const float* db, float* dR, float* dV, float* dP) {
  const float* dbg = db; const float* dba = db + 3;
  float rj[3], E[9];
  mv3(p.JRg, dbg, rj);
  exp_so3_host(rj, E);
  double X[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      X[3 * i + j] = (double)(float)((double)p.dR[3 * i] * E[j] + (double)p.dR[3 * i + 1] * E[3 + j] + (double)p.dR[3 * i + 2] * E[6 + j]);
  for (int it = 0; it < 8; it++) {
    double Xi[9];
    inv33d(X, Xi);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) X[3 * i + j] = 0.5 * (X[3 * i + j] + Xi[3 * j + i]);
  }
  for (int k = 0; k < 9; k++) dR[k] = (float)X[k];
  float t1[3], t2[3], u1[3], u2[3];
  mv3(p.JVg, dbg, t1); mv3(p.JVa, dba, t2); mv3(p.JPg, dbg, u1); mv3(p.JPa, dba, u2);
  for (int i = 0; i < 3; i++) { dV[i] = (p.dV[i] + t1[i]) + t2[i]; dP[i] = (p.dP[i] + u1[i]) + u2[i]; }
}

struct ImuFrame {              // IMU members of Frame (Frame.h): mTcw, mVw, mImuBias, mpImuPreintegrated, mTimeStamp
  float Tcw[16];
  float vel[3] = {0.f, 0.f, 0.f};
  float bias[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // bax, bay, baz, bwx, bwy, bwz
  bool has_pre = false;
  vido_imu_preint pre;         // integrated with pre_b; db = bu - b: gyro (0..2), acc (3..5) (IMU::Preintegrated::db)
  float pre_b[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, db[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  size_t q1 = 0;               // queue samples delivered when the frame was preintegrated (Reintegrate)
  double t = 0, t_prev = 0;
};
struct TrackInfo { int first_frame, len, pid, epoch; };  // pid valid for the window graph built in `epoch`

struct FrontFrame {         // front-end results of one frame, on the host
  std::vector<vido_keypoint> kps;
  std::vector<int32_t> kp_mask;          // per keypoint: mask / depth / flow at its truncated position
  std::vector<float> kp_depth, kp_flow;
  std::vector<int32_t> as_idx;           // Frame-ctor association
  std::vector<float> as_corres, as_flow, as_depth;
  std::vector<float> ob_keys, ob_corres, ob_flow, ob_depth;  // Frame-ctor object samples (mvObjKeys ... of Frame.cc:184-211)
  std::vector<int32_t> ob_sem;
};

}  // namespace

struct TrackState {
  // ---- VIO state (Tracking: mpImuCalib, mlQueueImuData, mbImuInitialized, mScale, mRwg, mbg, mba, mTinit)
  bool vio = false;
  float Tbc[16], Tcb[16], imu_noise[4];
  std::vector<vido_imu_sample> imu_q;    // every delivered sample (kept: Reintegrate re-reads its range)
  std::vector<int64_t> imu_q_frame;      // frame counter before which each sample was delivered (nondecreasing)
  size_t imu_head = 0;                   // queue front (mlQueueImuData.front())
  int64_t frames_seen = 0;               // frames handed to the back-end (tracked or skipped)
  std::vector<ImuFrame> fr;              // by frame id; fr[0] (the initial frame) is not in Map::vpFrames
  struct PreCache { bool valid = false; double t_prev = 0, t = 0; float bias[6]; size_t q1 = 0; vido_imu_preint pre; };
  std::vector<PreCache> pre_cache;       // preintegrations of the current front-end batch (one launch)
  double cur_t = 0;                      // timestamp of the frame in the back-end
  vido_imu_state ist;
  // sequence state
  std::vector<MapFrame> map;
  std::vector<TrackInfo> tracks;
  bool initialised = false, has_velocity = false;
  float mVelocity[16];
  float lastTcw[16];
  std::vector<float> last_keys, last_depth, last_corres, last_flow;  // mpLastFrame mvStatKeys / mvStatDepth / mvCorres / mvFlowNext
  int f_id = 0;
  int ba_epoch = 0;
  double ht[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // debug (VIDO_HOST_TIMING): host ms in record wait / consume / BA stage+go / BA retire / enqueue / front-end
  long hn = 0;
  double dt[12] = {0};   // debug (VIDO_HOST_TIMING): host-path ms in mask update / PnP / pose-opt / carry-over / object tracking / object motions / static renewal / object renewal / last-map copies
  long dn = 0;
  // ---- window-solver host thread (device-chained path): tracklet linking, staging, launch and retirement of the window solves
  //      run beside the tracker thread.  It owns `tracks`, MapFrame::track / pos, the solver queue and (while jobs are pending)
  //      the Twc / rel / p3 write-back of Map frames; the tracker thread only appends frames (capacity reserved up front, so the
  //      elements never move) and joins (ba_async_join) before anything else touches that state.
  struct BaAsyncJob { int N, window; vido_track_stats* st; };
  std::thread ba_thread;
  std::mutex ba_mu;
  std::condition_variable ba_cv_job, ba_cv_idle;
  std::deque<BaAsyncJob> ba_jobs;
  bool ba_thread_on = false, ba_thread_stop = false, ba_thread_busy = false;
  int ba_async_rc = VIDO_OK;          // sticky: first error of a job
  std::string ba_async_err;
  int ba_N_override = 0;              // Map size the job being staged was posted for (0: the current size)
  double bt[2] = {0, 0};              // debug (VIDO_HOST_TIMING): solver-thread ms in stage+launch / retirement
  bool call_failed = false;    // the last vido_track_frames call returned an error: whatever it left queued is discarded
  bool chain_active = false;   // the static tracker state lives on the device (chain_kernels.cu); the host vectors mirror it
  // object state of mpLastFrame: mvObjKeys / mvObjDepth / mvObjCorres / mvObjFlowNext / vSemObjLabel and nModLabel /
  // nSemPosition / bObjStat / vObjMod
  std::vector<float> lo_keys, lo_depth, lo_corres, lo_flow;
  std::vector<int32_t> lo_sem;
  std::vector<int> l_mod_label, l_sem_pos;
  std::vector<char> l_obj_stat;
  std::vector<std::array<float, 16>> l_obj_mod;
  std::vector<DynTrack> dyn_tracks;
  int max_id = 1;
  int obj_cap = 0;                 // stride-4 sampling grid = upper bound of the object samples of a frame
  bool dyn_seen = false;           // a frame with object samples was seen: their device->host copy rides with the front-end
  int32_t* d_last_mask = nullptr;  // mSegMapLast / mFlowMapLast (Tracking.cc:777-780), kept while object features are alive
  float* d_last_flow = nullptr;
  bool have_last_maps = false;
  // window BA jobs: [fly] is being solved on the BA stream while the next frame is tracked and its job [fly ^ 1] is staged
  struct BaJob {
    int start = 0, end = 0;
    vido_track_stats* st = nullptr;
    vido_ba_problem pr;
    std::vector<float> poses, rel, pts, oxyz;
    std::vector<int> op, ol;
    std::vector<int> ofeat;          // per observation: feature index inside its frame, for the write-back
    std::vector<int> pframe, pfeat;  // per point: frame / feature of its first observation (initial position)
    std::vector<int> prev_pose, prev_point;  // index of every pose / point in the previous window's problem (-1: not in it)
    int epoch = 0;
  } job[3];
  // Up to two windows are queued on the BA stream: window k+1 is launched BEFORE window k has finished -- the values they
  // share (19 of 20 poses, the odometry between them, every point that stays in the window) are gathered from window k's
  // output on the device -- and window k's results are written back into the Map while window k+1 is already running.
  bool ba_staged = false;
  int ba_queue[3] = {0, 0, 0}, ba_nq = 0;   // job slots in flight, oldest first
  // the window of frame k is staged and queued while the camera PnP of frame k+1 runs on the GPU (ctx->idle_work)
  struct { bool valid = false; int window = 0; vido_track_stats* st = nullptr; } ba_deferred;
  int ba_deferred_rc = 0;
  int ba_stage_slot = 0, ba_rest = -1;
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
    int32_t* d_obn;             // object samples per frame (inside d_out)
    char* d_obj = nullptr;      // object samples, one block: keys | corres | flow | depth | label, each [B][obj_cap]
    char* h_obj = nullptr;      // pinned mirror
    float* d_obkeys; float* d_obcorres; float* d_obflow; float* d_obdepth; int32_t* d_obsem;
    bool obj_copied = false;
    cudaEvent_t copied = nullptr, done = nullptr, ev0 = nullptr, ev1 = nullptr;
    bool launched = false;
    int B = 0, channels = 0;
    const void* key = nullptr;  // identity of the batch: image pointer of its first frame
    const uint8_t* in_img = nullptr; const float* in_depth = nullptr; const float* in_flow = nullptr; const int32_t* in_mask = nullptr;  // inputs the kernels read
  } fe[2];
  size_t fe_out_bytes = 0, fe_obj_bytes = 0;
  int fe_cur = 0;
  uint8_t* d_gray = nullptr;     // [B][H][W] (front-end stream only)
  cudaStream_t copy_stream = nullptr, fe_stream = nullptr;
  cudaStream_t obj_stream = nullptr;   // object part of a frame whose static part runs on the device chain (hybrid_consume)
  bool obj_stream_pending = false;     // copies queued on obj_stream that nobody has waited for yet
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

static int ba_async_join(vido_ctx* ctx);
static void ba_async_shutdown(vido_ctx* ctx);

int trk_setup(vido_ctx* ctx) {
  TrackState* ts = new TrackState();
  ts->map.reserve(4096);   // the solver thread indexes Map frames while the tracker appends: no reallocation under it (chain_consume)
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
               o_nkp = o_asdepth + al(4 * K), o_asn = o_nkp + al(4 * (size_t)B), o_obn = o_asn + al(4 * (size_t)B),
               o_flag = o_obn + al(4 * (size_t)B);
  ts->fe_out_bytes = o_flag + 256;
  ts->obj_cap = ((c.width + 3) / 4) * ((c.height + 3) / 4);
  ts->q_cap = std::max(ts->q_cap, ts->obj_cap + 64);
  const size_t OC = (size_t)ts->obj_cap * B;
  const size_t q_keys = 0, q_corres = q_keys + al(8 * OC), q_flow = q_corres + al(8 * OC), q_depth = q_flow + al(8 * OC), q_sem = q_depth + al(4 * OC);
  ts->fe_obj_bytes = q_sem + al(4 * OC);
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
    F.d_asn = (int32_t*)(F.d_out + o_asn); F.d_flag = (int32_t*)(F.d_out + o_flag); F.d_obn = (int32_t*)(F.d_out + o_obn);
    VIDO_CUDA(cudaMalloc(&F.d_obj, ts->fe_obj_bytes));
    VIDO_CUDA(cudaMallocHost(&F.h_obj, ts->fe_obj_bytes));
    F.d_obkeys = (float*)(F.d_obj + q_keys); F.d_obcorres = (float*)(F.d_obj + q_corres); F.d_obflow = (float*)(F.d_obj + q_flow);
    F.d_obdepth = (float*)(F.d_obj + q_depth); F.d_obsem = (int32_t*)(F.d_obj + q_sem);
    VIDO_CUDA(cudaEventCreateWithFlags(&F.copied, cudaEventDisableTiming));
    VIDO_CUDA(cudaEventCreateWithFlags(&F.done, cudaEventDisableTiming));
    VIDO_CUDA(cudaEventCreate(&F.ev0));
    VIDO_CUDA(cudaEventCreate(&F.ev1));
  }
  VIDO_CUDA(cudaMalloc(&ts->d_gray, px * B));
  VIDO_CUDA(vido_create_stream(&ts->copy_stream, false));
  VIDO_CUDA(vido_create_stream(&ts->fe_stream, false));
  VIDO_CUDA(vido_create_stream(&ts->obj_stream, true));
  VIDO_CUDA(cudaMalloc(&ts->d_q, 8 * ts->q_cap)); VIDO_CUDA(cudaMalloc(&ts->d_qmask, 4 * ts->q_cap));
  VIDO_CUDA(cudaMalloc(&ts->d_qdepth, 4 * ts->q_cap)); VIDO_CUDA(cudaMalloc(&ts->d_qflow, 8 * ts->q_cap));
  VIDO_CUDA(cudaMalloc(&ts->d_check, 8 * ts->q_cap)); VIDO_CUDA(cudaMalloc(&ts->d_used, std::max(ts->kp_cap, ts->obj_cap)));
  VIDO_CUDA(cudaMalloc(&ts->d_last_mask, px * 4));
  VIDO_CUDA(cudaMalloc(&ts->d_last_flow, px * 8));
  return VIDO_OK;
}

void trk_teardown(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (ts) ba_async_shutdown(ctx);
  if (!ts) return;
  if (ts->copy_stream) { cudaStreamSynchronize(ts->copy_stream); cudaStreamDestroy(ts->copy_stream); }
  if (ts->fe_stream) { cudaStreamSynchronize(ts->fe_stream); cudaStreamDestroy(ts->fe_stream); }
  if (ts->obj_stream) { cudaStreamSynchronize(ts->obj_stream); cudaStreamDestroy(ts->obj_stream); }
  for (int k = 0; k < 2; k++) {
    TrackState::FeSlot& F = ts->fe[k];
    cudaFree(F.d_img); cudaFree(F.d_depth); cudaFree(F.d_flow); cudaFree(F.d_mask); cudaFree(F.d_out); cudaFreeHost(F.h_out);
    cudaFree(F.d_obj); cudaFreeHost(F.h_obj);
    if (F.copied) cudaEventDestroy(F.copied);
    if (F.done) cudaEventDestroy(F.done);
    if (F.ev0) cudaEventDestroy(F.ev0);
    if (F.ev1) cudaEventDestroy(F.ev1);
  }
  cudaFree(ts->d_gray); cudaFree(ts->d_last_mask); cudaFree(ts->d_last_flow);
  cudaFree(ts->d_q); cudaFree(ts->d_qmask); cudaFree(ts->d_qdepth); cudaFree(ts->d_qflow); cudaFree(ts->d_check); cudaFree(ts->d_used);
  delete ts;
  ctx->trk = nullptr;
}

int trk_reset(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  ba_async_join(ctx);   // (an error of the abandoned sequence is dropped with it)
  if (ts->obj_stream) cudaStreamSynchronize(ts->obj_stream);
  ts->obj_stream_pending = false;
  while (ts->ba_nq > 0) { vido_lm_stats ls; ba_collect(ctx, &ts->job[ts->ba_queue[0]].pr, &ls); ts->ba_queue[0] = ts->ba_queue[1]; ts->ba_queue[1] = ts->ba_queue[2]; ts->ba_nq--; }
  ts->ba_deferred.valid = false;
  ts->chain_active = false;
  ts->map.clear(); ts->tracks.clear();
  ts->initialised = false; ts->has_velocity = false; ts->f_id = 0; ts->ba_epoch = 0;
  cudaStreamSynchronize(ts->copy_stream); cudaStreamSynchronize(ts->fe_stream);
  ts->fe[0].launched = ts->fe[1].launched = false; ts->hint.clear(); ts->ba_rest = -1; ts->ba_staged = false;
  ts->last_keys.clear(); ts->last_depth.clear(); ts->last_corres.clear(); ts->last_flow.clear();
  ts->lo_keys.clear(); ts->lo_depth.clear(); ts->lo_corres.clear(); ts->lo_flow.clear(); ts->lo_sem.clear();
  ts->l_mod_label.clear(); ts->l_sem_pos.clear(); ts->l_obj_stat.clear(); ts->l_obj_mod.clear();
  ts->dyn_tracks.clear(); ts->max_id = 1; ts->have_last_maps = false;
  ts->fr.clear(); ts->imu_q.clear(); ts->imu_q_frame.clear(); ts->imu_head = 0; ts->frames_seen = 0; ts->pre_cache.clear();
  if (ts->vio) {
    memset(&ts->ist, 0, sizeof ts->ist);
    ts->ist.scale = 1.0; ts->ist.status = -1;
    ts->ist.Rwg[0] = ts->ist.Rwg[4] = ts->ist.Rwg[8] = 1.0;
  }
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
    rc = assoc_sample_objects(ctx, F.in_depth, F.in_flow, F.in_mask, B, 1, F.d_obkeys, F.d_obcorres, F.d_obflow, F.d_obdepth, F.d_obsem,
                              F.d_obn, ts->obj_cap);
    if (rc) break;
    dim3 grid((ts->kp_cap + 255) / 256, B);
    kp_lookup_kernel<<<grid, 256, 0, fs>>>(F.d_kp, F.d_nkp, ts->kp_cap, c.width, c.height, F.in_depth, F.in_flow, F.in_mask, px,
                                           c.choose_data, c.depth_map_factor, c.bf, ctx->mscale, F.d_kpmask, F.d_kpdepth, F.d_kpflow);
    ctx->launches++;
    cudaEventRecord(F.ev1, fs);
    if (cudaGetLastError() != cudaSuccess) { ctx->err = "front-end launch failed"; rc = VIDO_ERR_CUDA; break; }
    if (cudaMemcpyAsync(F.d_flag, ctx->d_err, 4, cudaMemcpyDeviceToDevice, fs) != cudaSuccess ||
        cudaMemcpyAsync(F.h_out, F.d_out, ts->fe_out_bytes, cudaMemcpyDeviceToHost, fs) != cudaSuccess ||
        (ts->dyn_seen && cudaMemcpyAsync(F.h_obj, F.d_obj, ts->fe_obj_bytes, cudaMemcpyDeviceToHost, fs) != cudaSuccess) ||
        cudaEventRecord(F.done, fs) != cudaSuccess) { ctx->err = "front-end copy failed"; rc = VIDO_ERR_CUDA; break; }
    F.obj_copied = ts->dyn_seen;
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
  const int32_t* obn = (const int32_t*)hp(F.d_obn);
  bool any_obj = false;
  for (int b = 0; b < B; b++) {
    if (obn[b] > ts->obj_cap) { ctx->err = "object samples exceed the sampling grid"; return VIDO_ERR_CAPACITY; }
    any_obj |= obn[b] > 0;
  }
  if (any_obj && !F.obj_copied) {  // first batch with objects: fetch the samples now; later batches carry them along
    ts->dyn_seen = true;
    VIDO_CUDA(cudaMemcpyAsync(F.h_obj, F.d_obj, ts->fe_obj_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
    F.obj_copied = true;
  }
  auto ho = [&](const void* dptr) { return F.h_obj + ((const char*)dptr - F.d_obj); };
  const float* obkeys = (const float*)ho(F.d_obkeys); const float* obcorres = (const float*)ho(F.d_obcorres);
  const float* obflow = (const float*)ho(F.d_obflow); const float* obdepth = (const float*)ho(F.d_obdepth);
  const int32_t* obsem = (const int32_t*)ho(F.d_obsem);
  out.resize(B);
  for (int b = 0; b < B; b++) {
    FrontFrame& f = out[b];
    {
      const size_t no = (size_t)std::max(obn[b], 0), oo = (size_t)b * ts->obj_cap;
      f.ob_keys.assign(obkeys + 2 * oo, obkeys + 2 * (oo + no)); f.ob_corres.assign(obcorres + 2 * oo, obcorres + 2 * (oo + no));
      f.ob_flow.assign(obflow + 2 * oo, obflow + 2 * (oo + no)); f.ob_depth.assign(obdepth + oo, obdepth + oo + no);
      f.ob_sem.assign(obsem + oo, obsem + oo + no);
    }
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
// dynamic objects
// ---------------------------------------------------------------------------------------------------------
// used[i] = 1 if object sample i (xy) lies within 1 px of an already kept object feature (Tracking.cc:3186-3200); the kept
// list is the same for every object of the frame, so one pass over all samples answers every later query
__global__ void __launch_bounds__(256) topup_used_xy_kernel(const float* __restrict__ xy, int n, const float* __restrict__ check, int m,
                                                            uint8_t* __restrict__ used) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  const float sx = xy[2 * i], sy = xy[2 * i + 1];
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

// Frame::UnprojectStereoObject / Optimizer::Get3DinWorld: pixel + depth through Twc (float, cv::Mat gemm rounding)
static inline void px_to_world(const vido_config& c, float u, float v, float z, const float* Twc, float* o) {
  const float invfx = 1.0f / c.fx, invfy = 1.0f / c.fy;
  const float xc[3] = {(u - c.cx) * z * invfx, (v - c.cy) * z * invfy, z};
  for (int r = 0; r < 3; r++)
    o[r] = (float)((double)Twc[4 * r] * xc[0] + (double)Twc[4 * r + 1] * xc[1] + (double)Twc[4 * r + 2] * xc[2]) + Twc[4 * r + 3];
}

// The object mask of frame b was re-warped by UpdateMask: the Frame-ctor stages that read the mask (static association,
// object sampling, per-keypoint lookups) are repeated for that frame on the back-end stream (rare path).
static int fe_redo_frame(vido_ctx* ctx, TrackState::FeSlot& F, int b, FrontFrame& f) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const size_t px = (size_t)c.width * c.height, K = ts->kp_cap, OC = ts->obj_cap;
  cudaStream_t s = ctx->stream;
  const float* dd = F.in_depth + b * px; const float* df = F.in_flow + 2 * b * px; const int32_t* dm = F.in_mask + b * px;
  int rc = assoc_frame_associate(ctx, F.d_kp + b * K, F.d_nkp + b, ts->kp_cap, dd, df, dm, 1, 1, F.d_asidx + b * K, F.d_ascor + 2 * b * K,
                                 F.d_asflow + 2 * b * K, F.d_asdepth + b * K, F.d_asn + b, ts->kp_cap);
  if (rc) return rc;
  rc = assoc_sample_objects(ctx, dd, df, dm, 1, 1, F.d_obkeys + 2 * b * OC, F.d_obcorres + 2 * b * OC, F.d_obflow + 2 * b * OC,
                            F.d_obdepth + b * OC, F.d_obsem + b * OC, F.d_obn + b, ts->obj_cap);
  if (rc) return rc;
  dim3 grid((ts->kp_cap + 255) / 256, 1);
  kp_lookup_kernel<<<grid, 256, 0, s>>>(F.d_kp + b * K, F.d_nkp + b, ts->kp_cap, c.width, c.height, dd, df, dm, px, c.choose_data,
                                        c.depth_map_factor, c.bf, ctx->mscale, F.d_kpmask + b * K, F.d_kpdepth + b * K, F.d_kpflow + 2 * b * K);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  int32_t cnt[2] = {0, 0};
  VIDO_CUDA(cudaMemcpyAsync(&cnt[0], F.d_asn + b, 4, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaMemcpyAsync(&cnt[1], F.d_obn + b, 4, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  const size_t n = f.kps.size(), m = (size_t)std::min(std::max(cnt[0], 0), ts->kp_cap), no = (size_t)std::min(std::max(cnt[1], 0), ts->obj_cap);
  f.as_idx.resize(m); f.as_corres.resize(2 * m); f.as_flow.resize(2 * m); f.as_depth.resize(m);
  f.ob_keys.resize(2 * no); f.ob_corres.resize(2 * no); f.ob_flow.resize(2 * no); f.ob_depth.resize(no); f.ob_sem.resize(no);
  auto back = [&](void* dst, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s) : cudaSuccess; };
  VIDO_CUDA(back(f.kp_mask.data(), F.d_kpmask + b * K, 4 * n));
  VIDO_CUDA(back(f.kp_depth.data(), F.d_kpdepth + b * K, 4 * n));
  VIDO_CUDA(back(f.kp_flow.data(), F.d_kpflow + 2 * b * K, 8 * n));
  VIDO_CUDA(back(f.as_idx.data(), F.d_asidx + b * K, 4 * m));
  VIDO_CUDA(back(f.as_corres.data(), F.d_ascor + 2 * b * K, 8 * m));
  VIDO_CUDA(back(f.as_flow.data(), F.d_asflow + 2 * b * K, 8 * m));
  VIDO_CUDA(back(f.as_depth.data(), F.d_asdepth + b * K, 4 * m));
  VIDO_CUDA(back(f.ob_keys.data(), F.d_obkeys + 2 * b * OC, 8 * no));
  VIDO_CUDA(back(f.ob_corres.data(), F.d_obcorres + 2 * b * OC, 8 * no));
  VIDO_CUDA(back(f.ob_flow.data(), F.d_obflow + 2 * b * OC, 8 * no));
  VIDO_CUDA(back(f.ob_depth.data(), F.d_obdepth + b * OC, 4 * no));
  VIDO_CUDA(back(f.ob_sem.data(), F.d_obsem + b * OC, 4 * no));
  VIDO_CUDA(cudaStreamSynchronize(s));
  return VIDO_OK;
}

// per-frame object state while the frame is tracked (mpCurrentFrame's object members)
struct DynFrame {
  std::vector<float> keys, depth;       // mvObjKeys (refined in place by the object pose optimisation), mvObjDepth
  std::vector<int32_t> sem;             // vSemObjLabel
  std::vector<int> label;               // vObjLabel
  std::vector<float> flow3;             // vFlow_3d
  std::vector<int> mod_label, sem_pos;  // nModLabel, nSemPosition
  std::vector<char> stat;               // bObjStat
  std::vector<std::array<float, 16>> mod;  // vObjMod
  std::vector<std::array<float, 3>> centre;
  std::vector<std::vector<int>> obj_id, inlier_id;  // vnObjID, vnObjInlierID
};

static int majority_label(const std::vector<int>& v) {  // std::map order + stable descending count (Tracking.cc:1853-1863)
  std::map<int, int> dups;
  for (int k : v) ++dups[k];
  int best = 0, cnt = -1;
  for (auto& k : dups)
    if (k.second > cnt) { cnt = k.second; best = k.first; }
  return best;
}

// GrabImageRGBD :391-421 -- object features of the new frame = last frame's correspondences + fresh depth / label lookups
static int dyn_carry_over(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int slot, DynFrame& D) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const int W = c.width, H = c.height;
  const int n = (int)(ts->lo_corres.size() / 2);
  D.keys = ts->lo_corres;
  D.depth.assign(n, -1.f); D.sem.assign(n, -1); D.label.assign(n, -2);
  if (n == 0) return VIDO_OK;
  std::vector<int32_t> m2(n);
  std::vector<float> d2(n), f2(2 * (size_t)n);
  int rc = query_maps(ctx, d_depth, d_flow, d_mask, slot, D.keys.data(), n, m2.data(), d2.data(), f2.data());
  if (rc) return rc;
  for (int i = 0; i < n; i++) {
    const int u = (int)D.keys[2 * i], v = (int)D.keys[2 * i + 1];
    if (u < (W - 1) && u > 0 && v < (H - 1) && v > 0 && d2[i] < c.th_depth_obj && d2[i] > 0) { D.depth[i] = d2[i]; D.sem[i] = m2[i]; }
    else { D.depth[i] = 0.1f; D.sem[i] = 0; }
  }
  return VIDO_OK;
}

// GetSceneFlowObj (Tracking.cc:1582-1668) + DynObjTracking (:1670-1912); returns the feature index lists of the objects
static std::vector<std::vector<int>> dyn_track_objects(vido_ctx* ctx, const float* curTcw, DynFrame& D) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const int W = c.width, H = c.height;
  const int N = (int)D.sem.size();
  float Twl[16], Twc[16];
  inv44(ts->lastTcw, Twl);
  inv44(curTcw, Twc);
  D.flow3.assign(3 * (size_t)N, 0.f);
  for (int i = 0; i < N; i++) {
    if (D.sem[i] <= 0 || ts->lo_sem[i] <= 0) { D.label[i] = -1; continue; }
    float xp[3], xc[3];
    px_to_world(c, ts->lo_keys[2 * i], ts->lo_keys[2 * i + 1], ts->lo_depth[i], Twl, xp);
    px_to_world(c, D.keys[2 * i], D.keys[2 * i + 1], D.depth[i], Twc, xc);
    for (int r = 0; r < 3; r++) D.flow3[3 * i + r] = xc[r] - xp[r];
  }
  std::vector<int> UniLab(D.sem.begin(), D.sem.end());
  std::sort(UniLab.begin(), UniLab.end());
  UniLab.erase(std::unique(UniLab.begin(), UniLab.end()), UniLab.end());
  std::vector<std::vector<int>> Posi(UniLab.size());
  for (int i = 0; i < N; i++) {
    if (D.label[i] == -1) continue;
    const size_t j = std::lower_bound(UniLab.begin(), UniLab.end(), (int)D.sem[i]) - UniLab.begin();
    Posi[j].push_back(i);
  }
  std::vector<std::vector<int>> ObjId;
  std::vector<int> sem_posi;
  const int shrin_thr_row = 10, shrin_thr_col = 20;
  for (size_t i = 0; i < Posi.size(); i++) {
    float count = 0;
    for (int id : Posi[i]) {
      const float u = D.keys[2 * id], v = D.keys[2 * id + 1];
      if (v < shrin_thr_row || v > (H - shrin_thr_row) || u < shrin_thr_col || u > (W - shrin_thr_col)) count = count + 1;
    }
    if (count / Posi[i].size() > 0.5f) {
      for (int id : Posi[i]) D.label[id] = -1;
      continue;
    }
    ObjId.push_back(Posi[i]);
    sem_posi.push_back(UniLab[i]);
  }
  std::vector<std::vector<int>> ObjIdNew;
  std::vector<int> SemPosNew;
  for (size_t i = 0; i < ObjId.size(); i++) {
    float obj_center_depth = 0, sf_count = 0;
    for (int id : ObjId[i]) {
      obj_center_depth = obj_center_depth + D.depth[id];
      const float fx = D.flow3[3 * id], fz = D.flow3[3 * id + 2];
      const float sf_norm = std::sqrt(fx * fx + fz * fz);
      if (sf_norm < c.sf_mg_thres) sf_count = sf_count + 1;
    }
    if (sf_count / ObjId[i].size() > c.sf_ds_thres) {
      for (int id : ObjId[i]) D.label[id] = 0;
      continue;
    } else if (obj_center_depth / ObjId[i].size() > c.th_depth_obj || ObjId[i].size() < 150) {
      for (int id : ObjId[i]) D.label[id] = -1;
      continue;
    }
    ObjIdNew.push_back(ObjId[i]);
    SemPosNew.push_back(sem_posi[i]);
  }
  if (ts->f_id == 1) ts->max_id = 1;
  std::vector<int> LabId(ObjIdNew.size());
  for (size_t i = 0; i < ObjIdNew.size(); i++) {
    std::vector<int> Lb_last;
    for (int id : ObjIdNew[i]) Lb_last.push_back(ts->lo_sem[id]);
    const int New_lab = majority_label(Lb_last);
    bool exist = false;
    if (ts->max_id != 1) {
      for (size_t k = 0; k < ts->l_sem_pos.size(); k++)
        if (ts->l_sem_pos[k] == New_lab && ts->l_obj_stat[k]) { LabId[i] = ts->l_mod_label[k]; exist = true; break; }
    }
    if (!exist) { LabId[i] = ts->max_id; ts->max_id = ts->max_id + 1; }
    for (int id : ObjIdNew[i]) D.label[id] = LabId[i];
  }
  D.mod_label = LabId;
  D.sem_pos = SemPosNew;
  return ObjIdNew;
}

// object loop of Tracking::Track (Tracking.cc:1179-1308): GetInitModelObj per object (PnP kernels), then ONE pose-optimisation
// launch for all objects that kept >= 50 inliers (the objects use disjoint feature sets, so the order does not matter)
static int dyn_object_motions(vido_ctx* ctx, const float* curTcw, const std::vector<std::vector<int>>& ObjIdNew, DynFrame& D) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const size_t no = ObjIdNew.size();
  D.stat.assign(no, 1); D.mod.resize(no); D.centre.assign(no, std::array<float, 3>{0.f, 0.f, 0.f});
  D.obj_id = ObjIdNew; D.inlier_id.assign(no, std::vector<int>());
  float Twl[16], Twc[16];
  inv44(ts->lastTcw, Twl);
  inv44(curTcw, Twc);
  struct Job { size_t obj; std::vector<int> ids; std::vector<float> obs, fl, dep, fo; std::vector<int32_t> inl; float init[16]; };
  std::vector<Job> jobs;
  // ---- GetInitModelObj for all objects of the frame: the PnP problems are independent (disjoint feature sets), so they go
  //      through the RANSAC kernels together (one H2D block, one launch pair, one D2H block)
  std::vector<vido_pnp_problem> pps(no);
  std::vector<std::vector<float>> cur2d(no), p3d(no);
  std::vector<std::vector<int32_t>> idsv(no);
  for (size_t i = 0; i < no; i++) {
    const std::vector<int>& ObjId = ObjIdNew[i];
    const int N = (int)ObjId.size();
    cur2d[i].resize(2 * (size_t)N); p3d[i].resize(3 * (size_t)N); idsv[i].resize(N);
    float cs[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < N; j++) {
      const int id = ObjId[j];
      float xp[3];
      px_to_world(c, ts->lo_keys[2 * id], ts->lo_keys[2 * id + 1], ts->lo_depth[id], Twl, xp);
      for (int r = 0; r < 3; r++) { cs[r] += xp[r]; p3d[i][3 * j + r] = xp[r]; }
      cur2d[i][2 * j] = D.keys[2 * id]; cur2d[i][2 * j + 1] = D.keys[2 * id + 1];
    }
    const float invn = (float)(1.0 / (double)N);
    D.centre[i] = {cs[0] * invn, cs[1] * invn, cs[2] * invn};
    int PreObjID = -1;
    for (size_t k = 0; k < ts->l_mod_label.size(); k++)
      if (ts->l_mod_label[k] == D.mod_label[i]) { PreObjID = (int)k; break; }
    vido_pnp_problem& pp = pps[i];
    memset(&pp, 0, sizeof pp);
    vido_pnp_default_params(&pp);
    pp.n = N; pp.cur_xy = cur2d[i].data(); pp.pts3d = p3d[i].data(); pp.valid = nullptr; pp.inlier_ids = idsv[i].data();
    if (PreObjID != -1) mul44(curTcw, ts->l_obj_mod[PreObjID].data(), pp.Tcw_motion);
    else { memcpy(pp.Tcw_motion, curTcw, sizeof(float) * 16); pp.no_motion_model = 1; }
    pp.fx = c.fx; pp.fy = c.fy; pp.cx = c.cx; pp.cy = c.cy;
  }
  if (no) {
    const int rc = pnp_init_model_batch(ctx, pps.data(), (int)no);
    if (rc) return rc;
  }
  for (size_t i = 0; i < no; i++) {
    const std::vector<int>& ObjId = ObjIdNew[i];
    const int N = (int)ObjId.size();
    const vido_pnp_problem& pp = pps[i];
    const std::vector<int32_t>& ids = idsv[i];
    std::vector<int> in_ids(pp.n_inliers);
    std::vector<char> keep(N, 0);
    for (int k = 0; k < pp.n_inliers; k++) { in_ids[k] = ObjId[ids[k]]; keep[ids[k]] = 1; }
    for (int j = 0; j < N; j++)
      if (!keep[j]) D.label[ObjId[j]] = -1;
    if (in_ids.size() < 50) {
      D.stat[i] = 0;
      eye44(D.mod[i].data());
      D.centre[i] = {0.f, 0.f, 0.f};
      D.inlier_id[i] = in_ids;
      continue;
    }
    Job J;
    J.obj = i; J.ids = in_ids;
    const size_t n = in_ids.size();
    J.obs.resize(2 * n); J.fl.resize(2 * n); J.dep.resize(n); J.fo.resize(2 * n); J.inl.resize(n);
    for (size_t k = 0; k < n; k++) {
      const int id = in_ids[k];
      J.obs[2 * k] = ts->lo_keys[2 * id]; J.obs[2 * k + 1] = ts->lo_keys[2 * id + 1];
      J.fl[2 * k] = ts->lo_flow[2 * id]; J.fl[2 * k + 1] = ts->lo_flow[2 * id + 1];
      J.dep[k] = ts->lo_depth[id];
    }
    memcpy(J.init, pp.Tcw_out, sizeof J.init);
    jobs.push_back(std::move(J));
  }
  if (!c.b_joint) {
    // ---- PoseOptimizationObjMot (Optimizer.cc:2826-3035) for all objects of the frame in one launch: world-frame motion H
    //      (initialised with inv(Tcw) * mInitModel), reprojection through P = K * Tcw, no refinement of the keypoints
    std::vector<vido_projopt_problem> pj(jobs.size());
    std::vector<std::vector<float>> obs(jobs.size()), p3(jobs.size());
    double P[12];
    const double KK[12] = {c.fx, 0, c.cx, 0, 0, c.fy, c.cy, 0, 0, 0, 1, 0};
    for (int r = 0; r < 3; r++)
      for (int q = 0; q < 4; q++) {
        double v = 0;
        for (int m = 0; m < 4; m++) v += KK[4 * r + m] * (double)curTcw[4 * m + q];
        P[4 * r + q] = v;
      }
    for (size_t k = 0; k < jobs.size(); k++) {
      Job& J = jobs[k];
      const size_t n = J.ids.size();
      obs[k].resize(2 * n); p3[k].resize(3 * n);
      for (size_t q = 0; q < n; q++) {
        const int id = J.ids[q];
        obs[k][2 * q] = D.keys[2 * id]; obs[k][2 * q + 1] = D.keys[2 * id + 1];
        px_to_world(c, ts->lo_keys[2 * id], ts->lo_keys[2 * id + 1], ts->lo_depth[id], Twl, &p3[k][3 * q]);
      }
      vido_projopt_problem& pr = pj[k];
      memset(&pr, 0, sizeof pr);
      vido_projopt_default_params(&pr, 1);
      pr.n = (int)n; pr.obs_xy = obs[k].data(); pr.pts3d = p3[k].data(); pr.inlier = J.inl.data();
      mul44(Twc, J.init, pr.T_init);
      memcpy(pr.P, P, sizeof P);
    }
    if (!jobs.empty()) {
      const int rc = projopt_host(ctx, pj.data(), (int)jobs.size(), nullptr);
      if (rc) return rc;
    }
    for (size_t k = 0; k < jobs.size(); k++) {
      Job& J = jobs[k];
      memcpy(D.mod[J.obj].data(), pj[k].T_out, sizeof(float) * 16);
      std::vector<int> InlierID;
      for (size_t q = 0; q < J.ids.size(); q++) {
        if (J.inl[q]) InlierID.push_back(J.ids[q]);
        else D.label[J.ids[q]] = -1;
      }
      D.inlier_id[J.obj] = InlierID;
    }
    return VIDO_OK;
  }
  // ---- PoseOptimizationFlow2 for all objects of the frame: one CTA per object, 16 objects per launch
  for (size_t j0 = 0; j0 < jobs.size(); j0 += 16) {
    const int nb = (int)std::min<size_t>(16, jobs.size() - j0);
    std::vector<vido_poseopt_problem> pr(nb);
    for (int k = 0; k < nb; k++) {
      Job& J = jobs[j0 + k];
      vido_poseopt_problem& po = pr[k];
      memset(&po, 0, sizeof po);
      vido_poseopt_default_params(&po);
      po.n = (int)J.ids.size(); po.obs_xy = J.obs.data(); po.flow_xy = J.fl.data(); po.depth = J.dep.data();
      memcpy(po.Tcw_init, J.init, sizeof J.init);
      memcpy(po.Tcw_last, ts->lastTcw, sizeof(float) * 16);
      po.fx = c.fx; po.fy = c.fy; po.cx = c.cx; po.cy = c.cy;
      po.flow_out = J.fo.data(); po.inlier = J.inl.data();
      po.info_prior = 0.5f; po.rounds = 1; po.its = 200;
    }
    int rc = po_flow2_host(ctx, pr.data(), nb, nullptr);
    if (rc) return rc;
    for (int k = 0; k < nb; k++) {
      Job& J = jobs[j0 + k];
      mul44(Twc, pr[k].Tcw_out, D.mod[J.obj].data());  // vObjMod = inv(Tcw) * Obj_X  (Tracking.cc:1271)
      std::vector<int> InlierID;
      for (size_t q = 0; q < J.ids.size(); q++) {
        const int id = J.ids[q];
        if (J.inl[q]) {
          D.keys[2 * id] = (float)((double)ts->lo_keys[2 * id] + (double)J.fo[2 * q]);
          D.keys[2 * id + 1] = (float)((double)ts->lo_keys[2 * id + 1] + (double)J.fo[2 * q + 1]);
          InlierID.push_back(id);
        } else D.label[id] = -1;
      }
      D.inlier_id[J.obj] = InlierID;
    }
  }
  return VIDO_OK;
}

// object part of Tracking::RenewFrameInfo (Tracking.cc:3112-3289) + Map bookkeeping (:1351-1355, 1390-1422) + dynamic
// tracklets (GetDynamicTrackNew, incremental); F is the MapFrame of the current frame
static int dyn_renew(vido_ctx* ctx, const FrontFrame& ff, const float* curTcw, const float* d_obkeys, const float* d_depth,
                     const float* d_flow, const int32_t* d_mask, int slot, DynFrame& D, MapFrame& F) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const int W = c.width, H = c.height, max_num_obj = c.max_track_obj;
  cudaStream_t s = ctx->stream;
  std::vector<float> keys, corres, fl, dep;
  std::vector<int32_t> sem;
  std::vector<int> inl, lab;
  const size_t no = D.inlier_id.size();
  std::vector<int> ObjFeaCount(no, -1);
  {  // (1) inliers of the tracked objects at their refined (then truncated) positions
    std::vector<float> qxy;
    for (size_t i = 0; i < no; i++)
      if (D.stat[i]) for (int id : D.inlier_id[i]) { qxy.push_back(D.keys[2 * id]); qxy.push_back(D.keys[2 * id + 1]); }
    const int nq = (int)(qxy.size() / 2);
    std::vector<int32_t> m2(nq);
    std::vector<float> d2(nq), f2(2 * (size_t)nq);
    int rc = query_maps(ctx, d_depth, d_flow, d_mask, slot, qxy.data(), nq, m2.data(), d2.data(), f2.data());
    if (rc) return rc;
    int q = 0;
    for (size_t i = 0; i < no; i++) {
      if (!D.stat[i]) continue;
      int count = 0;
      for (int id : D.inlier_id[i]) {
        const int k = q++;
        const int x = (int)D.keys[2 * id], y = (int)D.keys[2 * id + 1];
        if (x >= W || y >= H || x <= 0 || y <= 0) continue;
        if (m2[k] != 0 && d2[k] < 25 && d2[k] > 0) {
          const float fx = f2[2 * k], fy = f2[2 * k + 1];
          if (x + fx < W && y + fy < H && x + fx > 0 && y + fy > 0) {
            keys.push_back((float)x); keys.push_back((float)y);
            dep.push_back(d2[k]); sem.push_back(m2[k]);
            fl.push_back(fx); fl.push_back(fy);
            corres.push_back(x + fx); corres.push_back(y + fy);
            inl.push_back(id); lab.push_back(D.label[id]);
            count = count + 1;
          }
        }
      }
      ObjFeaCount[i] = count;
    }
  }
  // (2) top-up per tracked object from this frame's samples (15 interleaved passes), >= 1 px from every kept inlier
  const int nt = (int)ff.ob_sem.size(), mcheck = (int)(keys.size() / 2);
  bool need_topup = false;
  for (size_t i = 0; i < no; i++) need_topup |= D.stat[i] && ObjFeaCount[i] < max_num_obj;
  std::vector<uint8_t> used(nt, 0);
  if (need_topup && nt > 0 && mcheck > 0) {
    if (mcheck > ts->q_cap) { ctx->err = "object renewal check list too long"; return VIDO_ERR_CAPACITY; }
    VIDO_CUDA(cudaMemcpyAsync(ts->d_check, keys.data(), 8 * (size_t)mcheck, cudaMemcpyHostToDevice, s));
    topup_used_xy_kernel<<<(nt + 7) / 8, 256, 0, s>>>(d_obkeys, nt, ts->d_check, mcheck, ts->d_used);
    ctx->launches++;
    VIDO_CUDA(cudaGetLastError());
    VIDO_CUDA(cudaMemcpyAsync(used.data(), ts->d_used, nt, cudaMemcpyDeviceToHost, s));
    VIDO_CUDA(cudaStreamSynchronize(s));
  }
  auto push_sample = [&](int j, int label) {
    keys.push_back(ff.ob_keys[2 * j]); keys.push_back(ff.ob_keys[2 * j + 1]);
    dep.push_back(ff.ob_depth[j]); sem.push_back(ff.ob_sem[j]);
    fl.push_back(ff.ob_flow[2 * j]); fl.push_back(ff.ob_flow[2 * j + 1]);
    corres.push_back(ff.ob_corres[2 * j]); corres.push_back(ff.ob_corres[2 * j + 1]);
    inl.push_back(-1); lab.push_back(label);
  };
  for (size_t i = 0; i < no; i++) {
    if (!D.stat[i]) continue;
    const int SemLabel = D.sem_pos[i];
    int tot_num = ObjFeaCount[i], start_id = 0;
    const int step = 15;
    while (tot_num < max_num_obj) {
      if (start_id == step) break;
      for (int j = start_id; j < nt; j += step) {
        if (ff.ob_sem[j] != SemLabel) continue;
        if (used[j]) continue;
        push_sample(j, D.mod_label[i]);
        tot_num = tot_num + 1;
        if (tot_num >= max_num_obj) break;
      }
      start_id = start_id + 1;
    }
  }
  // (3) semantic labels without a tracked object: all their samples, tracking label -2
  std::vector<int> UniLab(ff.ob_sem.begin(), ff.ob_sem.end());
  std::sort(UniLab.begin(), UniLab.end());
  UniLab.erase(std::unique(UniLab.begin(), UniLab.end()), UniLab.end());
  std::vector<char> NewLab(UniLab.size(), 0);
  for (size_t i = 0; i < D.sem_pos.size(); i++)
    for (size_t j = 0; j < UniLab.size(); j++)
      if (UniLab[j] == D.sem_pos[i] && D.stat[i]) { NewLab[j] = 1; break; }
  for (size_t i = 0; i < NewLab.size(); i++) {
    if (NewLab[i]) continue;
    for (int j = 0; j < nt; j++)
      if (UniLab[i] == ff.ob_sem[j]) push_sample(j, -2);
  }
  // (4) world points, Map arrays
  const size_t nf = dep.size();
  float Twc[16];
  inv44(curTcw, Twc);
  F.dxy = keys; F.ddepth = dep; F.dasso = inl; F.dlabel = lab;
  F.dp3.resize(3 * nf);
  for (size_t i = 0; i < nf; i++) px_to_world(c, keys[2 * i], keys[2 * i + 1], dep[i], Twc, &F.dp3[3 * i]);
  for (size_t i = 0; i < D.mod.size(); i++) {
    if (!D.stat[i]) continue;
    ObjEntry e;
    e.label = D.mod_label[i]; e.sem = D.sem_pos[i];
    memcpy(e.motion, D.mod[i].data(), sizeof e.motion);
    memcpy(e.motion_rf, e.motion, sizeof e.motion);
    memcpy(e.centre, D.centre[i].data(), sizeof e.centre);
    F.objects.push_back(e);
  }
  // (5) dynamic tracklets, incrementally (same chains as Tracking::GetDynamicTrackNew)
  F.dtrack.assign(nf, -1);
  MapFrame& P = ts->map.back();
  const int fcur = (int)ts->map.size();
  for (size_t j = 0; j < nf; j++) {
    const int p = inl[j];
    if (p < 0) continue;
    if (P.dtrack[p] >= 0) {
      F.dtrack[j] = P.dtrack[p];
      ts->dyn_tracks[P.dtrack[p]].len++;
    } else {
      const int t = (int)ts->dyn_tracks.size();
      ts->dyn_tracks.push_back({fcur - 1, p, 2, lab[j]});
      F.dtrack[j] = t;
      P.dtrack[p] = t;   // first element of the chain (read by the full-sequence graph builder)
    }
  }
  // (6) becomes mpLastFrame
  ts->lo_keys = keys; ts->lo_depth = dep; ts->lo_corres = corres; ts->lo_flow = fl; ts->lo_sem = sem;
  ts->l_mod_label = D.mod_label; ts->l_sem_pos = D.sem_pos; ts->l_obj_stat = D.stat; ts->l_obj_mod = D.mod;
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
  if (ts->ba_nq == 0) return VIDO_OK;
  TrackState::BaJob& J = ts->job[ts->ba_queue[0]];
  ts->ba_queue[0] = ts->ba_queue[1]; ts->ba_queue[1] = ts->ba_queue[2];
  ts->ba_nq--;
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
  const int N = ts->ba_N_override > 0 ? ts->ba_N_override : (int)ts->map.size();
  if (st) { st->ba_iterations = -1; st->ba_points = 0; st->ba_obs = 0; st->ba_trials = 0; }
  if (WINDOW <= 0) return VIDO_OK;
  while (ts->ba_nq == 3 || (ts->ba_nq > 0 && ba_oldest_done(ctx))) {   // retire finished solves; the oldest one when every slot is taken
    int rc0 = ba_finish(ctx);
    if (rc0) return rc0;
    ba_writeback_rest(ctx);
  }
  while (ts->ba_nq >= 1 && ts->job[ts->ba_queue[ts->ba_nq - 1]].epoch != ts->ba_epoch) {  // cannot chain to the newest: drain first
    int rc0 = ba_finish(ctx);
    if (rc0) return rc0;
    ba_writeback_rest(ctx);
  }
  ts->ba_stage_slot = ts->ba_nq ? (ts->ba_queue[ts->ba_nq - 1] + 1) % 3 : ts->ba_stage_slot;
  TrackState::BaJob& J = ts->job[ts->ba_stage_slot];
  const TrackState::BaJob* Jp = ts->ba_nq ? &ts->job[ts->ba_queue[ts->ba_nq - 1]] : nullptr;   // the newest window in the queue: chain to it
  const int start = N - WINDOW;
  J.start = start; J.end = N; J.st = st;
  J.poses.resize(16 * (size_t)WINDOW); J.rel.resize(16 * (size_t)std::max(WINDOW - 1, 0));
  J.oxyz.clear(); J.op.clear(); J.ol.clear(); J.ofeat.clear(); J.pframe.clear(); J.pfeat.clear();
  J.prev_pose.assign(WINDOW, -1); J.prev_point.clear();
  if (Jp)
    for (int i = start; i < N; i++)
      if (i >= Jp->start && i < Jp->end) J.prev_pose[i - start] = i - Jp->start;
  const int prev_epoch = ts->ba_epoch;
  const int epoch = ++ts->ba_epoch;
  J.epoch = epoch;
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
        J.prev_point.push_back((Jp && T.epoch == prev_epoch) ? T.pid : -1);   // same track, point of the window in flight
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
  int rc = Jp ? ba_prepare_chained(ctx, &pr, J.prev_pose.data(), J.prev_point.data()) : ba_prepare(ctx, &pr);
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
  ts->ba_queue[ts->ba_nq++] = slot;
  return VIDO_OK;
}

// stage + queue the window solve that was put off at the end of the previous frame
static int ba_flush_deferred(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->ba_deferred.valid) return VIDO_OK;
  ts->ba_deferred.valid = false;
  double t0 = now_ms();
  int rc = ba_stage(ctx, ts->ba_deferred.window, ts->ba_deferred.st);
  if (!rc) rc = ba_go(ctx);
  if (ts->ba_deferred.st) ts->ba_deferred.st->ms_ba += now_ms() - t0;
  return rc;
}

// ---- the window-solver host thread (see TrackState::ba_thread).  One job per tracked frame, in frame order: link the frame's
//      static tracklets, stage its window (ba_stage retires finished solves / the oldest one when the three slots are taken),
//      launch, then retire whatever has finished meanwhile.
static void link_static_tracks_at(TrackState* ts, MapFrame& F, MapFrame& P, int fcur);
static void ba_async_main(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  cudaSetDevice(ctx->device);
  for (;;) {
    TrackState::BaAsyncJob job;
    {
      std::unique_lock<std::mutex> lk(ts->ba_mu);
      ts->ba_thread_busy = false;
      if (ts->ba_jobs.empty()) ts->ba_cv_idle.notify_all();
      ts->ba_cv_job.wait(lk, [ts] { return ts->ba_thread_stop || !ts->ba_jobs.empty(); });
      if (ts->ba_jobs.empty()) return;   // stop requested and nothing left
      job = ts->ba_jobs.front();
      ts->ba_jobs.pop_front();
      ts->ba_thread_busy = true;
      if (ts->ba_async_rc != VIDO_OK) continue;   // after an error the remaining jobs are dropped (the call fails at its next join)
    }
    const double t0 = now_ms();
    link_static_tracks_at(ts, ts->map[job.N - 1], ts->map[job.N - 2], job.N - 1);
    ts->ba_N_override = job.N;
    int rc = ba_stage(ctx, job.window, job.st);
    if (!rc) rc = ba_go(ctx);
    ts->ba_N_override = 0;
    const double t1 = now_ms();
    if (job.st) job.st->ms_ba += t1 - t0;
    while (ts->ba_nq > 0 && rc == VIDO_OK && ba_oldest_done(ctx)) { rc = ba_finish(ctx); ba_writeback_rest(ctx); }
    ts->bt[0] += t1 - t0; ts->bt[1] += now_ms() - t1;
    if (rc != VIDO_OK) {
      std::lock_guard<std::mutex> lk(ts->ba_mu);
      if (ts->ba_async_rc == VIDO_OK) { ts->ba_async_rc = rc; ts->ba_async_err = ctx->err; }
    }
  }
}

// wait until the solver thread has worked off every posted job (solves may still be queued on the DEVICE); returns the sticky error
static int ba_async_join(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->ba_thread_on) return VIDO_OK;
  std::unique_lock<std::mutex> lk(ts->ba_mu);
  ts->ba_cv_idle.wait(lk, [ts] { return ts->ba_jobs.empty() && !ts->ba_thread_busy; });
  const int rc = ts->ba_async_rc;
  if (rc != VIDO_OK) { ctx->err = ts->ba_async_err; ts->ba_async_rc = VIDO_OK; }
  return rc;
}

static int ba_async_post(vido_ctx* ctx, int N, int window, vido_track_stats* st) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->ba_thread_on) {
    ts->ba_thread_stop = false; ts->ba_thread_busy = false;
    ts->ba_thread = std::thread(ba_async_main, ctx);
    ts->ba_thread_on = true;
  }
  std::lock_guard<std::mutex> lk(ts->ba_mu);
  if (ts->ba_async_rc != VIDO_OK) { const int rc = ts->ba_async_rc; ctx->err = ts->ba_async_err; ts->ba_async_rc = VIDO_OK; return rc; }
  ts->ba_jobs.push_back({N, window, st});
  ts->ba_cv_job.notify_one();
  return VIDO_OK;
}

static void ba_async_shutdown(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->ba_thread_on) return;
  { std::lock_guard<std::mutex> lk(ts->ba_mu); ts->ba_thread_stop = true; }
  ts->ba_cv_job.notify_all();
  ts->ba_thread.join();
  ts->ba_thread_on = false;
}


// ---------------------------------------------------------------------------------------------------------
// VIO glue (sensor = IMU_RGBD): the host side of Tracking's IMU path around the preintegration / inertial kernels
// ---------------------------------------------------------------------------------------------------------
static int trk_apply_scaled_rotation_impl(vido_ctx* ctx, const float* R, float s);

static void vio_imu_rotation(const TrackState* ts, const ImuFrame& f, float* Rwb) {   // Frame::GetImuRotation: mRwc * Tcb.R
  float Rwc[9], Rcb[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { Rwc[3 * r + c] = f.Tcw[4 * c + r]; Rcb[3 * r + c] = ts->Tcb[4 * r + c]; }
  mm3(Rwc, Rcb, Rwb);
}
static void vio_imu_position(const TrackState* ts, const ImuFrame& f, float* twb) {   // mOwb = mRwc * tcb + mOw, mOw = -mRcw.t() * mtcw
  float Rwc[9], tcw[3] = {f.Tcw[3], f.Tcw[7], f.Tcw[11]}, tcb[3] = {ts->Tcb[3], ts->Tcb[7], ts->Tcb[11]}, Ow[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Rwc[3 * r + c] = f.Tcw[4 * c + r];
  mv3(Rwc, tcw, Ow, -1.0);
  mv3(Rwc, tcb, twb, 1.0, Ow);
}
static void vio_set_imu_pose_velocity(const TrackState* ts, ImuFrame& f, const float* Rwb, const float* twb, const float* Vwb) {  // Frame.cc:510-521
  memcpy(f.vel, Vwb, sizeof f.vel);
  float Rbw[9], tbw[3], Tbw[16];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Rbw[3 * r + c] = Rwb[3 * c + r];
  mv3(Rbw, twb, tbw, -1.0);
  eye44(Tbw);
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) Tbw[4 * r + c] = Rbw[3 * r + c];
    Tbw[4 * r + 3] = tbw[r];
  }
  mul44(ts->Tcb, Tbw, f.Tcw);
}
static void vio_pre_set_new_bias(ImuFrame& f, const float* b) {   // IMU::Preintegrated::SetNewBias (ImuTypes.cc:328-339)
  if (f.has_pre)
    for (int k = 0; k < 3; k++) { f.db[k] = b[3 + k] - f.pre_b[3 + k]; f.db[3 + k] = b[k] - f.pre_b[k]; }
}
static void vio_set_new_bias(ImuFrame& f, const float* b) {       // Frame::SetNewBias (Frame.cc:464-469)
  memcpy(f.bias, b, sizeof f.bias);
  vio_pre_set_new_bias(f, b);
}
// queue window handed to the kernel: everything from the current queue front on (older samples lie before every interval)
static int vio_run_jobs(vido_ctx* ctx, size_t base, int njobs, const double* tp, const double* tc, const float* bias, const size_t* q1,
                        vido_imu_preint* out) {
  TrackState* ts = (TrackState*)ctx->trk;
  size_t hi = base;
  std::vector<int32_t> nvis(njobs);
  for (int j = 0; j < njobs; j++) { hi = std::max(hi, q1[j]); nvis[j] = (int32_t)(q1[j] > base ? q1[j] - base : 0); }
  int rc = imu_preintegrate_host(ctx, ts->imu_q.data() + base, (int)(hi - base), tp, tc, njobs, bias, ts->imu_noise, out, nvis.data());
  if (rc) return rc;
  for (int j = 0; j < njobs; j++) out[j].n_consumed += (int32_t)base;   // absolute queue index
  return VIDO_OK;
}
static size_t vio_visible(const TrackState* ts, int64_t frame_counter) {
  return (size_t)(std::upper_bound(ts->imu_q_frame.begin(), ts->imu_q_frame.end(), frame_counter) - ts->imu_q_frame.begin());
}
// Tracking::PreintegrateIMU for the B frames of a front-end batch in ONE launch: they all integrate with the bias of the last
// tracked frame (it only changes at the initialisation / a reintegration, which invalidates the rest of the cache)
static int vio_preintegrate_batch(vido_ctx* ctx, const vido_frame_inputs* in, int B) {
  TrackState* ts = (TrackState*)ctx->trk;
  ts->pre_cache.assign(B, TrackState::PreCache());
  if (!ts->vio || ts->fr.empty()) {   // frame 0 of the sequence is in this batch: its successors are cached from b = 1 on
    if (!ts->vio || B < 2) return VIDO_OK;
  }
  const int b0 = ts->fr.empty() ? 1 : 0;
  const int nj = B - b0;
  if (nj <= 0) return VIDO_OK;
  std::vector<double> tp(nj), tc(nj);
  std::vector<float> bias(6 * (size_t)nj, 0.f);
  std::vector<size_t> q1(nj);
  std::vector<vido_imu_preint> out(nj);
  for (int j = 0; j < nj; j++) {
    const int b = b0 + j;
    tp[j] = b == 0 ? ts->fr.back().t : in[b - 1].timestamp;
    tc[j] = in[b].timestamp;
    if (!ts->fr.empty()) memcpy(&bias[6 * (size_t)j], ts->fr.back().bias, sizeof(float) * 6);
    q1[j] = vio_visible(ts, ts->frames_seen + b);
  }
  int rc = vio_run_jobs(ctx, ts->imu_head, nj, tp.data(), tc.data(), bias.data(), q1.data(), out.data());
  if (rc) return rc;
  for (int j = 0; j < nj; j++) {
    TrackState::PreCache& C = ts->pre_cache[b0 + j];
    C.valid = true; C.t_prev = tp[j]; C.t = tc[j]; C.q1 = q1[j]; C.pre = out[j];
    memcpy(C.bias, &bias[6 * (size_t)j], sizeof C.bias);
  }
  return VIDO_OK;
}
// the new frame's IMU state (Frame ctor Frame.cc:437-447, Tracking.cc:1115-1119): velocity / bias of the last frame, preintegration
static int vio_new_frame(vido_ctx* ctx, int slot) {
  TrackState* ts = (TrackState*)ctx->trk;
  ImuFrame f;
  const ImuFrame& pf = ts->fr.back();
  f.t = ts->cur_t; f.t_prev = pf.t;
  memcpy(f.vel, pf.vel, sizeof f.vel);
  memcpy(f.bias, pf.bias, sizeof f.bias);
  f.q1 = vio_visible(ts, ts->frames_seen);
  f.has_pre = false;
  if (ts->imu_head < f.q1) {   // else: "Not IMU data in mlQueueImuData"
    const TrackState::PreCache* C = (slot >= 0 && slot < (int)ts->pre_cache.size()) ? &ts->pre_cache[slot] : nullptr;
    if (C && C->valid && C->t_prev == f.t_prev && C->t == f.t && C->q1 == f.q1 && memcmp(C->bias, pf.bias, sizeof C->bias) == 0) {
      f.pre = C->pre;
    } else {
      int rc = vio_run_jobs(ctx, ts->imu_head, 1, &f.t_prev, &f.t, pf.bias, &f.q1, &f.pre);
      if (rc) return rc;
    }
    memcpy(f.pre_b, pf.bias, sizeof f.pre_b);
    f.has_pre = true;
    ts->imu_head = (size_t)f.pre.n_consumed;
  }
  ts->fr.push_back(f);
  return VIDO_OK;
}
// Tracking::UpdateFrameIMU (Tracking.cc:889-923); mpLastFrame == mpCurrentFrame == fr.back()
static void vio_update_frame_imu(TrackState* ts, const float* b) {
  ImuFrame& c = ts->fr.back();
  const ImuFrame& p = ts->fr[ts->fr.size() - 2];
  vio_set_new_bias(c, b);
  if (!c.has_pre) return;
  const float Gz[3] = {0.f, 0.f, -9.79f};
  float twb1[3], Rwb1[9], dR[9], dV[3], dP[3], Rwb[9], twb[3], Vwb[3], rp[3], rv[3];
  vio_imu_position(ts, p, twb1);
  vio_imu_rotation(ts, p, Rwb1);
  updated_deltas(c.pre, c.db, dR, dV, dP);
  const float t12 = c.pre.dT;
  mm3(Rwb1, dR, Rwb);
  mv3(Rwb1, dP, rp);
  mv3(Rwb1, dV, rv);
  const float ht2 = 0.5f * t12 * t12;
  for (int r = 0; r < 3; r++) {
    twb[r] = ((twb1[r] + (float)((double)p.vel[r] * (double)t12)) + (float)((double)ht2 * (double)Gz[r])) + rp[r];
    Vwb[r] = (p.vel[r] + (float)((double)Gz[r] * (double)t12)) + rv[r];
  }
  vio_set_imu_pose_velocity(ts, c, Rwb, twb, Vwb);
  memcpy(ts->lastTcw, c.Tcw, sizeof c.Tcw);
}
// the inertial problem over Map::vpFrames = fr[1..N] (Optimizer.cc:2441-2560 / 2336-2425) on the inertial kernel, and the
// write-back of mode 0 (Optimizer.cc:2589-2619).  *ok = false: a frame without preintegration breaks the chain.
static int vio_run_inertial(vido_ctx* ctx, int mode, float priorG, float priorA, bool* ok) {
  TrackState* ts = (TrackState*)ctx->trk;
  std::vector<ImuFrame>& fr = ts->fr;
  const int N = (int)fr.size() - 1;
  *ok = false;
  std::vector<float> Rwb(9 * (size_t)N), twb(3 * (size_t)N), vel(3 * (size_t)N), blin(6 * (size_t)(N - 1));
  std::vector<vido_imu_preint> pre(N - 1);
  for (int i = 1; i <= N; i++) {
    vio_imu_rotation(ts, fr[i], &Rwb[9 * (size_t)(i - 1)]);
    vio_imu_position(ts, fr[i], &twb[3 * (size_t)(i - 1)]);
    memcpy(&vel[3 * (size_t)(i - 1)], fr[i].vel, sizeof(float) * 3);
    vio_pre_set_new_bias(fr[i], fr[i - 1].bias);
    if (i >= 2) {
      if (!fr[i].has_pre) return VIDO_OK;
      pre[i - 2] = fr[i].pre;
      memcpy(&blin[6 * (size_t)(i - 2)], fr[i].pre_b, sizeof(float) * 6);
    }
  }
  vido_inertial_problem p;
  memset(&p, 0, sizeof p);
  inertial_default_params(&p);
  p.n_frames = N; p.Rwb = Rwb.data(); p.twb = twb.data(); p.velocity = vel.data(); p.preint = pre.data(); p.bias_lin = blin.data();
  p.mode = mode; p.its = mode == 0 ? 200 : 10;
  p.prior_g = priorG; p.prior_a = priorA;
  for (int k = 0; k < 9; k++) p.Rwg[k] = ts->ist.Rwg[k];
  p.scale = ts->ist.scale;
  for (int k = 0; k < 3; k++) { p.bg[k] = (double)fr[1].bias[3 + k]; p.ba[k] = (double)fr[1].bias[k]; }   // VertexGyroBias(vpFs.front())
  vido_lm_stats lm;
  int rc = inertial_opt_host(ctx, &p, &lm);
  if (rc) return rc;
  if (mode == 0) { ts->ist.lm_iterations = lm.iterations; ts->ist.lm_trials = lm.total_trials; }
  for (int k = 0; k < 9; k++) ts->ist.Rwg[k] = p.Rwg[k];
  ts->ist.scale = p.scale;
  if (mode == 0) {
    for (int k = 0; k < 3; k++) { ts->ist.bg[k] = p.bg[k]; ts->ist.ba[k] = p.ba[k]; }
    const float b[6] = {(float)p.ba[0], (float)p.ba[1], (float)p.ba[2], (float)p.bg[0], (float)p.bg[1], (float)p.bg[2]};
    std::vector<int> redo;
    for (int i = 1; i <= N; i++) {
      memcpy(fr[i].vel, &vel[3 * (size_t)(i - 1)], sizeof(float) * 3);
      double d2 = 0;
      for (int k = 0; k < 3; k++) { const float d = fr[i].bias[3 + k] - b[3 + k]; d2 += (double)d * d; }
      const bool re = std::sqrt(d2) > 0.01;
      vio_set_new_bias(fr[i], b);
      if (re && fr[i].has_pre) redo.push_back(i);
    }
    if (!redo.empty()) {   // IMU::Preintegrated::Reintegrate (ImuTypes.cc:236-243): the same samples with the new bias, one launch
      const int nj = (int)redo.size();
      std::vector<double> tp(nj), tc(nj);
      std::vector<float> bias(6 * (size_t)nj);
      std::vector<size_t> q1(nj);
      std::vector<vido_imu_preint> out(nj);
      for (int j = 0; j < nj; j++) {
        tp[j] = fr[redo[j]].t_prev; tc[j] = fr[redo[j]].t; q1[j] = fr[redo[j]].q1;
        memcpy(&bias[6 * (size_t)j], b, sizeof b);
      }
      rc = vio_run_jobs(ctx, 0, nj, tp.data(), tc.data(), bias.data(), q1.data(), out.data());
      if (rc) return rc;
      for (int j = 0; j < nj; j++) {
        ImuFrame& f = fr[redo[j]];
        f.pre = out[j];
        memcpy(f.pre_b, b, sizeof b);
        for (int k = 0; k < 6; k++) f.db[k] = 0.f;
        ts->ist.n_reintegrated++;
      }
    }
    ts->pre_cache.clear();   // the bias of the last frame changed: the batch's remaining preintegrations are redone
  }
  *ok = true;
  return VIDO_OK;
}
static int vio_apply_and_update(vido_ctx* ctx, const float* b) {   // tail shared by InitializeIMU / ScaleRefinement
  TrackState* ts = (TrackState*)ctx->trk;
  if (std::fabs(ts->ist.scale - 1.0) > 0.00001) {
    float Rgw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Rgw[3 * r + c] = (float)ts->ist.Rwg[3 * c + r];   // Converter::toCvMat(mRwg).t()
    int rc = trk_apply_scaled_rotation_impl(ctx, Rgw, (float)ts->ist.scale);
    if (rc) return rc;
    vio_update_frame_imu(ts, b);
  }
  return VIDO_OK;
}
// Tracking::InitializeIMU(1e2, 1e9) (Tracking.cc:937-1044)
static int vio_initialize_imu(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  std::vector<ImuFrame>& fr = ts->fr;
  vido_imu_state& ist = ts->ist;
  const int N = (int)fr.size() - 1;
  if (N < 10) { ist.status = 1; return VIDO_OK; }
  const double first_ts = fr[1].t;
  if (fr.back().t - first_ts < 2.0) { ist.status = 1; return VIDO_OK; }
  float dirG[3] = {0.f, 0.f, 0.f};
  for (int i = 1; i <= N; i++) {
    if (!fr[i].has_pre) continue;
    float R[9], dR[9], dV[3], dP[3], p1[3], p0[3], rv[3];
    vio_imu_rotation(ts, fr[i - 1], R);
    updated_deltas(fr[i].pre, fr[i].db, dR, dV, dP);
    vio_imu_position(ts, fr[i], p1);
    vio_imu_position(ts, fr[i - 1], p0);
    mv3(R, dV, rv);
    for (int r = 0; r < 3; r++) {
      dirG[r] = dirG[r] - rv[r];
      const float v = (float)((double)(p1[r] - p0[r]) * (1.0 / (double)fr[i].pre.dT));
      fr[i].vel[r] = v;
      fr[i - 1].vel[r] = v;
    }
  }
  const double nrm = std::sqrt((double)dirG[0] * dirG[0] + (double)dirG[1] * dirG[1] + (double)dirG[2] * dirG[2]);
  for (int r = 0; r < 3; r++) dirG[r] = (float)((double)dirG[r] * (1.0 / nrm));
  const float gI[3] = {0.f, 0.f, -1.f};
  const float v[3] = {gI[1] * dirG[2] - gI[2] * dirG[1], gI[2] * dirG[0] - gI[0] * dirG[2], gI[0] * dirG[1] - gI[1] * dirG[0]};
  const float nv = (float)std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
  const float cosg = (float)((double)gI[0] * dirG[0] + (double)gI[1] * dirG[1] + (double)gI[2] * dirG[2]);
  const float ang = (float)std::acos((double)cosg);
  float vzg[3], Rwg[9];
  for (int r = 0; r < 3; r++) vzg[r] = (float)((double)(float)((double)v[r] * (double)ang) * (1.0 / (double)nv));
  exp_so3_host(vzg, Rwg);
  for (int k = 0; k < 9; k++) ist.Rwg[k] = Rwg[k];
  ist.t_init = (float)(fr.back().t - first_ts);
  ist.scale = 1.0;
  bool ok = false;
  int rc = vio_run_inertial(ctx, 0, 1e2f, 1e9f, &ok);
  if (rc) return rc;
  if (!ok) { ist.status = 3; return VIDO_OK; }
  if (ist.scale < 1e-1) { ist.status = 2; return VIDO_OK; }
  float b[6];
  memcpy(b, fr[1].bias, sizeof b);   // vpF[0]->GetImuBias()
  rc = vio_apply_and_update(ctx, b);
  if (rc) return rc;
  ist.initialized = 1; ist.init_frame = N; ist.status = 0;
  return VIDO_OK;
}
// Tracking::ScaleRefinement (Tracking.cc:1046-1077)
static int vio_scale_refinement(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  for (int k = 0; k < 9; k++) ts->ist.Rwg[k] = (k % 4 == 0) ? 1.0 : 0.0;
  ts->ist.scale = 1.0;
  bool ok = false;
  int rc = vio_run_inertial(ctx, 1, 0.f, 0.f, &ok);
  if (rc || !ok) return rc;
  ts->ist.n_refinements++;
  if (ts->ist.scale < 1e-1) return VIDO_OK;
  float b[6];
  memcpy(b, ts->fr.back().bias, sizeof b);
  return vio_apply_and_update(ctx, b);
}
// the IMU part of Tracking::Track after the window optimisation (Tracking.cc:1452-1480)
static int vio_after_ba(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  vido_imu_state& ist = ts->ist;
  int rc = VIDO_OK;
  if (!ist.initialized) rc = vio_initialize_imu(ctx);
  if (rc) return rc;
  if (ist.initialized && ist.t_init < 100.0f) {
    ist.t_init = (float)((double)ist.t_init + (ts->fr.back().t - ts->fr[ts->fr.size() - 2].t));
    const float T = ist.t_init;
    const bool win = (T > 15.0f && T < 15.5f) || (T > 25.0f && T < 25.5f) || (T > 35.0f && T < 35.5f) || (T > 45.0f && T < 45.5f) ||
                     (T > 55.0f && T < 55.5f) || (T > 65.0f && T < 65.5f) || (T > 75.0f && T < 75.5f);
    if ((int)ts->fr.size() - 1 <= 1000 && win) rc = vio_scale_refinement(ctx);
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------------------
// back-end of one frame (sequential)
// ---------------------------------------------------------------------------------------------------------
// tracklets of the static features, incrementally (same chains as Tracking::GetStaticTrack, src/Tracking.cc:2514-2613): every
// feature of the new frame F continues the track of its predecessor in the last Map frame, or starts one with it
static void link_static_tracks_at(TrackState* ts, MapFrame& F, MapFrame& P, int fcur);
static void link_static_tracks(TrackState* ts, MapFrame& F) { link_static_tracks_at(ts, F, ts->map.back(), (int)ts->map.size()); }
// F = the frame with Map index fcur (not necessarily pushed yet), P = the frame in front of it
static void link_static_tracks_at(TrackState* ts, MapFrame& F, MapFrame& P, int fcur) {
  const int nf = (int)F.asso.size();
  F.track.assign(nf, -1); F.pos.assign(nf, 0);
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
}

// premask_nrec >= 0: UpdateMask (and the front-end redo it asked for) has been done by the caller, with that many recovered labels
static int back_end(vido_ctx* ctx, TrackState::FeSlot& FS, FrontFrame& ff, int slot, float* Tcw_out, vido_track_stats* st, int premask_nrec = -1) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const int W = c.width, H = c.height;
  cudaStream_t s = ctx->stream;
  const vido_keypoint* d_kp = FS.d_kp;
  const float* d_depth = FS.in_depth; const float* d_flow = FS.in_flow; const int32_t* d_mask = FS.in_mask;
  const size_t px = (size_t)W * H;
  const float invfx = 1.0f / c.fx, invfy = 1.0f / c.fy;
  float curTcw[16];
  eye44(curTcw);
  { int rcj = ba_async_join(ctx); if (rcj) return rcj; }   // the host-driven path stages its window solves on this thread
  if (ts->obj_stream_pending) { VIDO_CUDA(cudaStreamSynchronize(ts->obj_stream)); ts->obj_stream_pending = false; }
  if (st) { memset(st, 0, sizeof *st); st->n_keypoints = (int)ff.kps.size(); }
  int skipped = 0;
  double t0 = now_ms();
  // ---- Tracking::UpdateMask (Tracking.cc:353-364): votes of the last frame's object features in the new mask; a lost mask is
  //      re-warped in place and the mask-dependent Frame-ctor stages of this frame are repeated
  const bool htime = getenv("VIDO_HOST_TIMING") != nullptr;
  double tm = htime ? now_ms() : 0;
  auto lap = [&](int k) { if (htime) { const double t_ = now_ms(); ts->dt[k] += t_ - tm; tm = t_; } };
  if (premask_nrec >= 0) {
    if (st) st->n_masks_recovered = premask_nrec;
  } else if (ts->initialised && !ts->lo_sem.empty() && ts->have_last_maps) {
    const int nl = (int)ts->lo_sem.size();
    std::vector<int32_t> uniq(nl), rec(nl);
    const int nu = assoc_update_mask(ctx, ts->lo_sem.data(), ts->lo_corres.data(), nl, ts->d_last_mask, ts->d_last_flow,
                                     (int32_t*)d_mask + (size_t)slot * px, uniq.data(), rec.data(), nl);
    if (nu < 0) return nu;
    int nrec = 0;
    for (int k = 0; k < nu; k++) nrec += rec[k] ? 1 : 0;
    if (st) st->n_masks_recovered = nrec;
    if (nrec) { int rc = fe_redo_frame(ctx, FS, slot, ff); if (rc) return rc; }
  }
  if (!ts->initialised) {
    { int rc0 = ba_flush_deferred(ctx); if (rc0) return rc0; }
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
    eye44(F.Twc); eye44(F.rel); eye44(F.Twc_rf);
    if (ts->vio) {   // Tracking.cc:1555-1561: IMU pose from Tcb, zero velocity, empty preintegration
      ImuFrame f0;
      float Rwb0[9], twb0[3], V0[3] = {0.f, 0.f, 0.f};
      for (int r = 0; r < 3; r++) {
        for (int cc = 0; cc < 3; cc++) Rwb0[3 * r + cc] = ts->Tcb[4 * r + cc];
        twb0[r] = ts->Tcb[4 * r + 3];
      }
      vio_set_imu_pose_velocity(ts, f0, Rwb0, twb0, V0);
      memset(&f0.pre, 0, sizeof f0.pre);
      f0.pre.dR[0] = f0.pre.dR[4] = f0.pre.dR[8] = 1.f;
      f0.has_pre = true;
      f0.t = f0.t_prev = ts->cur_t;
      memcpy(curTcw, f0.Tcw, sizeof f0.Tcw);
      ts->fr.clear();
      ts->fr.push_back(f0);
    }
    ts->last_keys = F.xy; ts->last_depth = F.depth; ts->last_corres = ff.as_corres; ts->last_flow = ff.as_flow;
    memcpy(ts->lastTcw, curTcw, sizeof curTcw);
    {  // object samples of frame 0 (Frame.cc:184-211, Tracking.cc:1524-1541)
      const size_t no = ff.ob_sem.size();
      F.dxy = ff.ob_keys; F.ddepth = ff.ob_depth; F.dp3.resize(3 * no);
      F.dasso.assign(no, -1); F.dlabel.assign(no, -2); F.dtrack.assign(no, -1);
      for (size_t i = 0; i < no; i++) {
        const float z = F.ddepth[i], u = F.dxy[2 * i], v = F.dxy[2 * i + 1];
        F.dp3[3 * i] = (u - c.cx) * z * invfx; F.dp3[3 * i + 1] = (v - c.cy) * z * invfy; F.dp3[3 * i + 2] = z;
      }
      ts->lo_keys = ff.ob_keys; ts->lo_depth = ff.ob_depth; ts->lo_corres = ff.ob_corres; ts->lo_flow = ff.ob_flow; ts->lo_sem = ff.ob_sem;
      ts->l_mod_label.clear(); ts->l_sem_pos.clear(); ts->l_obj_stat.clear(); ts->l_obj_mod.clear();
      if (st) st->n_dyn_features = (int)no;
    }
    ts->map.push_back(std::move(F));
    ts->initialised = true;
  } else {
    const int Ns = (int)(ts->last_corres.size() / 2);
    if (Ns < 2) { skipped = 1; int rc0 = ba_flush_deferred(ctx); if (rc0) return rc0; }
    else {
      // ---- mvStatKeys = last mvCorres.  The reference also samples the new depth map at these positions into
      //      mvStatDepthTmp (Tracking.cc:369-389); in the static-only pipeline nothing reads that vector (RenewFrameInfo
      //      re-samples at the refined positions, Tracking.cc:2970-3010), so the lookup round trip is not issued.
      std::vector<float> keys = ts->last_corres;
      if (ts->vio) { int rcv = vio_new_frame(ctx, slot); if (rcv) return rcv; }
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
      lap(0);
      vido_pnp_problem pp;
      memset(&pp, 0, sizeof pp);
      vido_pnp_default_params(&pp);
      pp.n = Ns; pp.cur_xy = keys.data(); pp.pts3d = p3d.data(); pp.valid = valid.data(); pp.inlier_ids = ids.data();
      if (ts->has_velocity) mul44(ts->mVelocity, ts->lastTcw, pp.Tcw_motion);
      else memcpy(pp.Tcw_motion, ts->lastTcw, sizeof(float) * 16);
      pp.fx = c.fx; pp.fy = c.fy; pp.cx = c.cx; pp.cy = c.cy;
      // while the RANSAC kernels run: stage and queue the window solve of the previous frame
      ts->ba_deferred_rc = VIDO_OK;
      ctx->idle_work = [ctx, ts]() { ts->ba_deferred_rc = ba_flush_deferred(ctx); };
      int rc = pnp_init_model_host(ctx, &pp);
      if (ctx->idle_work) { ctx->idle_work = nullptr; ts->ba_deferred_rc = ba_flush_deferred(ctx); }
      if (rc) return rc;
      if (ts->ba_deferred_rc) return ts->ba_deferred_rc;
      std::vector<int> TM(ids.begin(), ids.begin() + pp.n_inliers);
      memcpy(curTcw, pp.Tcw_out, sizeof curTcw);
      double t1 = now_ms();
      lap(1);
      int n_pose_inliers = 0;
      const int n = (int)TM.size();
      if (c.b_joint) {
        // ---- PoseOptimizationFlow2Cam
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
        // while the pose optimisation runs: retire the older queued window solve (its results go back into the Map)
        ctx->idle_work = [ctx, ts]() {
          while (ts->ba_nq > 1 && ts->ba_deferred_rc == VIDO_OK) { ts->ba_deferred_rc = ba_finish(ctx); ba_writeback_rest(ctx); }
        };
        rc = po_flow2_host(ctx, &po, 1, nullptr);
        ctx->idle_work = nullptr;
        if (rc) return rc;
        if (ts->ba_deferred_rc) return ts->ba_deferred_rc;
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
        n_pose_inliers = po.n_inliers;
      } else {
        // ---- PoseOptimizationNew (Optimizer.cc:2180-2334): reprojection of the last frame's world points; the retirement
        //      of the older window solve still hides behind the kernel
        std::vector<float> obs(2 * (size_t)n), p3(3 * (size_t)n);
        std::vector<int32_t> inl(n);
        for (int i = 0; i < n; i++) {
          const int k = TM[i];
          obs[2 * i] = keys[2 * k]; obs[2 * i + 1] = keys[2 * k + 1];
          p3[3 * i] = p3d[3 * k]; p3[3 * i + 1] = p3d[3 * k + 1]; p3[3 * i + 2] = p3d[3 * k + 2];
        }
        vido_projopt_problem pj;
        memset(&pj, 0, sizeof pj);
        vido_projopt_default_params(&pj, 0);
        pj.n = n; pj.obs_xy = obs.data(); pj.pts3d = p3.data(); pj.inlier = inl.data();
        memcpy(pj.T_init, curTcw, sizeof curTcw);
        pj.fx = c.fx; pj.fy = c.fy; pj.cx = c.cx; pj.cy = c.cy;
        rc = projopt_host(ctx, &pj, 1, nullptr);
        if (rc) return rc;
        while (ts->ba_nq > 1) { rc = ba_finish(ctx); if (rc) return rc; ba_writeback_rest(ctx); }
        memcpy(curTcw, pj.T_out, sizeof curTcw);
        n_pose_inliers = pj.n_inliers;
        if (n >= 3)
          for (int i = 0; i < n; i++)
            if (!inl[i]) TM[i] = -1;
      }
      double t2 = now_ms();
      lap(2);
      // ---- motion model: mVelocity = Tcw * LastTwc
      float LastTwc[16];
      inv44(ts->lastTcw, LastTwc);
      mul44(curTcw, LastTwc, ts->mVelocity);
      ts->has_velocity = true;
      // ---- dynamic objects: carry-over, scene flow, object tracking, object motions (Tracking.cc:391-421, 1160-1308)
      DynFrame D;
      const bool dyn = !ts->lo_corres.empty() || !ff.ob_sem.empty();
      if (!ts->lo_corres.empty()) {
        rc = dyn_carry_over(ctx, d_depth, d_flow, d_mask, slot, D);
        if (rc) return rc;
        lap(3);
        const std::vector<std::vector<int>> ObjIdNew = dyn_track_objects(ctx, curTcw, D);
        lap(4);
        rc = dyn_object_motions(ctx, curTcw, ObjIdNew, D);
        if (rc) return rc;
        lap(5);
        if (st) { st->n_objects = (int)ObjIdNew.size(); for (char b : D.stat) st->n_objects_ok += b ? 1 : 0; }
      }
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
      link_static_tracks(ts, F);
      memcpy(F.Twc, Twc, sizeof Twc);
      memcpy(F.Twc_rf, Twc, sizeof Twc);
      if (ts->vio) memcpy(ts->fr.back().Tcw, curTcw, sizeof curTcw);
      inv44(ts->mVelocity, F.rel);
      lap(6);
      if (dyn) {  // RenewFrameInfo (object part), Map bookkeeping, dynamic tracklets
        rc = dyn_renew(ctx, ff, curTcw, FS.d_obkeys + 2 * (size_t)slot * ts->obj_cap, d_depth, d_flow, d_mask, slot, D, F);
        if (rc) return rc;
        if (st) st->n_dyn_features = (int)F.ddepth.size();
      }
      ts->last_keys = F.xy; ts->last_depth = F.depth; ts->last_corres = ncorres; ts->last_flow = nflow;
      memcpy(ts->lastTcw, curTcw, sizeof curTcw);
      ts->map.push_back(std::move(F));
      double t3 = now_ms();
      lap(7);
      if (st) {
        st->ms_init = t1 - t0; st->ms_poseopt = t2 - t1; st->ms_renew = t3 - t2;
        st->n_matches = Ns; st->n_init_inliers = pp.n_inliers; st->init_winner = pp.winner; st->n_pose_inliers = n_pose_inliers;
        st->n_static = nf;
      }
    }
  }
  // mSegMapLast / mFlowMapLast (Tracking.cc:777-780): the slot buffers are recycled by the look-ahead front-end, so the maps
  // of the frame that becomes mpLastFrame are kept in private copies -- only while it carries object features
  if (!skipped) {
    ts->have_last_maps = false;
    if (!ts->lo_sem.empty()) {
      VIDO_CUDA(cudaMemcpyAsync(ts->d_last_mask, d_mask + (size_t)slot * px, px * 4, cudaMemcpyDeviceToDevice, s));
      VIDO_CUDA(cudaMemcpyAsync(ts->d_last_flow, d_flow + 2 * (size_t)slot * px, px * 8, cudaMemcpyDeviceToDevice, s));
      // the slot (or, with zero-copy device inputs, the caller's buffer) is recycled by the run-ahead front-end on other
      // streams and may be reused by the caller once the call returns: the private copies must be complete before either
      VIDO_CUDA(cudaStreamSynchronize(s));
      ts->have_last_maps = true;
    }
  }
  memcpy(Tcw_out, curTcw, sizeof(float) * 16);
  double t4 = now_ms();
  lap(8);
  if (htime) ts->dn++;
  const int window = ts->f_id < c.window_size ? ts->f_id : c.window_size;
  // The window of this frame is staged and queued during the next frame's camera PnP (ba_flush_deferred): it chains to the
  // solve in flight on the device, and the host work disappears behind kernels that had to be waited for anyway.
  int rc = VIDO_OK;
  if (st) { st->ba_iterations = -1; st->ba_points = 0; st->ba_obs = 0; st->ba_trials = 0; }
  if (!skipped) { ts->ba_deferred.valid = true; ts->ba_deferred.window = window; ts->ba_deferred.st = st; }
  if (ts->vio && !skipped && ts->f_id > 0) {   // InitializeIMU / ScaleRefinement (Tracking.cc:1452-1480); TrackRGBD returns mTcw after them
    rc = vio_after_ba(ctx);
    memcpy(Tcw_out, ts->lastTcw, sizeof(float) * 16);
  }
  if (st) st->ms_ba = now_ms() - t4;
  ts->f_id++;
  ts->frames_seen++;
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
// ---------------------------------------------------------------------------------------------------------
// device-chained static back-end (chain_kernels.cu): eligibility, queueing, consumption of the finished records
// ---------------------------------------------------------------------------------------------------------
// A frame can go through the chain when nothing of the object / inertial machinery is involved and the tracker state fits the
// chain's buffers (always true from the second tracked frame on: RenewFrameInfo caps the static features at MaxTrackPointBG+1).
// VIO mode on the device-chained path: only frames whose Tracking::Track tail is bookkeeping -- the IMU is initialised and the
// frame does not fall into one of the ScaleRefinement windows of mTinit (vio_after_ba); T is advanced like there (float sum).
static bool vio_frame_is_plain(const TrackState* ts, float* T, double t_prev, double t, int ahead /* frames queued in front of this one */) {
  const vido_imu_state& ist = ts->ist;
  if (!ist.initialized) return false;
  if (*T >= 100.0f) return true;
  const float Tn = (float)((double)*T + (t - t_prev));
  const bool win = (Tn > 15.0f && Tn < 15.5f) || (Tn > 25.0f && Tn < 25.5f) || (Tn > 35.0f && Tn < 35.5f) || (Tn > 45.0f && Tn < 45.5f) ||
                   (Tn > 55.0f && Tn < 55.5f) || (Tn > 65.0f && Tn < 65.5f) || (Tn > 75.0f && Tn < 75.5f);
  if (win && (int)ts->fr.size() + ahead <= 1000) return false;   // (fr.size() - 1 <= 1000 once this frame has been appended)
  *T = Tn;
  return true;
}

static bool chain_eligible(vido_ctx* ctx, const FrontFrame& ff, const vido_frame_inputs& in) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  if (getenv("VIDO_NO_CHAIN")) return false;   // debug: host-driven path only
  if (!ts->initialised || !c.b_joint || in.write_back_depth) return false;
  const bool objects = !ts->lo_corres.empty() || !ff.ob_sem.empty() || !ts->lo_sem.empty();
  if (ts->vio) {
    if (objects || ts->fr.empty() || getenv("VIDO_NO_VIO_CHAIN")) return false;
    float T = ts->ist.t_init;
    if (!vio_frame_is_plain(ts, &T, ts->fr.back().t, in.timestamp, 0)) return false;
  }
  // frames with object features: the static part on the chain, the object part on the host beside it (hybrid_consume)
  if (objects && getenv("VIDO_NO_HYBRID")) return false;
  return (int)(ts->last_corres.size() / 2) <= chain_capacity(ctx);
}

// the record of frame `slot` -> Map frame, host mirrors of the tracker state, per-frame outputs; then the window solve of the
// frame is staged and queued at once (the host-driven path defers it to the next frame's kernels to hide it; here the device
// is already busy with the following frames)
static int chain_consume(vido_ctx* ctx, int slot, const FrontFrame& ff, float* Tcw_out, vido_track_stats* st) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const int32_t* hdr; const float *Tcw, *Twc, *rel, *vel, *xy, *depth, *p3, *corres, *flow; const int32_t* asso;
  const double th0 = now_ms();
  int rc = chain_wait_record(ctx, slot, &hdr, &Tcw, &Twc, &rel, &vel, &xy, &depth, &p3, &corres, &flow, &asso);
  if (rc) return rc;
  const double th1 = now_ms();
  if (st) { memset(st, 0, sizeof *st); st->n_keypoints = (int)ff.kps.size(); st->ba_iterations = -1; }
  memcpy(Tcw_out, Tcw, sizeof(float) * 16);
  if (hdr[0] == 1) {   // lost tracking: nothing was processed (see back_end)
    rc = ba_async_join(ctx);
    if (!rc) rc = ba_flush_deferred(ctx);
    ts->f_id++; ts->frames_seen++;
    return rc ? rc : 1;
  }
  if (ts->vio) {   // the frame's IMU members (back_end: vio_new_frame before the PnP, mTcw after the pose optimisation, mTinit in vio_after_ba)
    rc = vio_new_frame(ctx, slot);
    if (rc) return rc;
    memcpy(ts->fr.back().Tcw, Tcw, sizeof(float) * 16);
    vido_imu_state& ist = ts->ist;
    if (ist.initialized && ist.t_init < 100.0f) ist.t_init = (float)((double)ist.t_init + (ts->fr.back().t - ts->fr[ts->fr.size() - 2].t));
  }
  const int nf = hdr[7];
  MapFrame F;
  F.xy.assign(xy, xy + 2 * (size_t)nf); F.depth.assign(depth, depth + nf); F.p3.assign(p3, p3 + 3 * (size_t)nf);
  F.asso.assign(asso, asso + nf);
  const bool async = !getenv("VIDO_BA_INLINE");   // debug: VIDO_BA_INLINE=1 stages the window solves on this thread
  if (!async) link_static_tracks(ts, F);
  else { F.track.assign(nf, -1); F.pos.assign(nf, 0); }   // linked by the solver thread (left like this if its job is dropped after an error)
  memcpy(F.Twc, Twc, sizeof(float) * 16); memcpy(F.Twc_rf, Twc, sizeof(float) * 16); memcpy(F.rel, rel, sizeof(float) * 16);
  ts->last_keys = F.xy; ts->last_depth = F.depth;
  ts->last_corres.assign(corres, corres + 2 * (size_t)nf); ts->last_flow.assign(flow, flow + 2 * (size_t)nf);
  memcpy(ts->lastTcw, Tcw, sizeof(float) * 16);
  memcpy(ts->mVelocity, vel, sizeof(float) * 16);
  ts->has_velocity = true;
  if (ts->map.size() == ts->map.capacity()) {   // the solver thread indexes Map frames: they must not move under it
    rc = ba_async_join(ctx);
    if (rc) return rc;
    ts->map.reserve(std::max<size_t>(4096, 2 * ts->map.capacity()));
  }
  ts->map.push_back(std::move(F));
  ts->have_last_maps = false;
  if (st) {
    st->n_matches = hdr[1]; st->n_init_inliers = hdr[2]; st->init_winner = hdr[3]; st->n_pose_inliers = hdr[6]; st->n_static = nf;
  }
  const int window = ts->f_id < c.window_size ? ts->f_id : c.window_size;
  ts->f_id++; ts->frames_seen++;
  const double th2 = now_ms();
  double th3 = th2;
  if (async) {
    // tracklet linking, staging, launch and retirement of this frame's window solve: the solver thread's job
    rc = ba_async_post(ctx, (int)ts->map.size(), window, st);
  } else {
    ts->ba_deferred.valid = true; ts->ba_deferred.window = window; ts->ba_deferred.st = st;
    rc = ba_flush_deferred(ctx);
    th3 = now_ms();
    // finished solves are retired without blocking (ba_stage blocks only when all three slots are taken): the host stays ahead
    // of the solver stream instead of waiting for the solve before last after every frame
    while (ts->ba_nq > 0 && rc == VIDO_OK && ba_oldest_done(ctx)) { rc = ba_finish(ctx); ba_writeback_rest(ctx); }
  }
  const double th4 = now_ms();
  ts->ht[0] += th1 - th0; ts->ht[1] += th2 - th1; ts->ht[2] += th3 - th2; ts->ht[3] += th4 - th3; ts->hn++;
  return rc;
}

// A frame with object features on the device-chained path.  Its static part (camera PnP, pose optimisation, static renewal)
// was queued with the tracker chain on the context stream -- possibly several frames ago; the object part (Tracking::UpdateMask,
// the carry-over of the object features, scene flow + DynObjTracking, the per-object PnP and motion optimisation, the object
// part of RenewFrameInfo, the Map bookkeeping) runs here, with its kernels on a second stream and scratch the chain does not
// share, in the order of the host-driven path (back_end).  The static part does not depend on the object part of the same or
// of earlier frames except through UpdateMask: when that re-warps a lost label into this frame's mask the front-end products
// the chain used are stale -- *fallback is set and the caller redoes the frame on the host-driven path (rare).
static int hybrid_consume(vido_ctx* ctx, TrackState::FeSlot& FS, FrontFrame& ff, int slot, float* Tcw_out, vido_track_stats* st, int* fallback) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  const size_t px = (size_t)c.width * c.height;
  const float* d_depth = FS.in_depth; const float* d_flow = FS.in_flow; const int32_t* d_mask = FS.in_mask;
  *fallback = -1;
  struct StreamSwap {   // the host-side helpers of the object path launch on ctx->stream
    vido_ctx* c; cudaStream_t keep;
    ~StreamSwap() { c->stream = keep; }
  } swap{ctx, ctx->stream};
  ctx->stream = ts->obj_stream;
  const double th0 = now_ms();
  if (st) { memset(st, 0, sizeof *st); st->n_keypoints = (int)ff.kps.size(); st->ba_iterations = -1; }
  int nrec = 0;
  if (!ts->lo_sem.empty() && ts->have_last_maps) {   // ---- Tracking::UpdateMask (see back_end)
    const int nl = (int)ts->lo_sem.size();
    std::vector<int32_t> uniq(nl), rec(nl);
    const int nu = assoc_update_mask(ctx, ts->lo_sem.data(), ts->lo_corres.data(), nl, ts->d_last_mask, ts->d_last_flow,
                                     (int32_t*)d_mask + (size_t)slot * px, uniq.data(), rec.data(), nl);
    if (nu < 0) return nu;
    for (int k = 0; k < nu; k++) nrec += rec[k] ? 1 : 0;
    if (nrec) { *fallback = nrec; return VIDO_OK; }
  }
  DynFrame D;
  int rc = VIDO_OK;
  if (!ts->lo_corres.empty()) {
    rc = dyn_carry_over(ctx, d_depth, d_flow, d_mask, slot, D);
    if (rc) return rc;
  }
  // ---- the static part of the frame: the chain's record
  const int32_t* hdr; const float *Tcw, *Twc, *rel, *vel, *xy, *depth, *p3, *corres, *flow; const int32_t* asso;
  const double th1 = now_ms();
  rc = chain_wait_record(ctx, slot, &hdr, &Tcw, &Twc, &rel, &vel, &xy, &depth, &p3, &corres, &flow, &asso);
  if (rc) return rc;
  const double th2 = now_ms();
  memcpy(Tcw_out, Tcw, sizeof(float) * 16);
  if (hdr[0] == 1) {   // lost tracking: nothing was processed (see back_end)
    ts->f_id++; ts->frames_seen++;
    return 1;
  }
  float curTcw[16];
  memcpy(curTcw, Tcw, sizeof curTcw);
  // ---- object tracking and motions against the LAST frame's pose (ts->lastTcw is updated below)
  if (!ts->lo_corres.empty()) {
    const std::vector<std::vector<int>> ObjIdNew = dyn_track_objects(ctx, curTcw, D);
    rc = dyn_object_motions(ctx, curTcw, ObjIdNew, D);
    if (rc) return rc;
    if (st) { st->n_objects = (int)ObjIdNew.size(); for (char b : D.stat) st->n_objects_ok += b ? 1 : 0; }
  }
  const int nf = hdr[7];
  MapFrame F;
  F.xy.assign(xy, xy + 2 * (size_t)nf); F.depth.assign(depth, depth + nf); F.p3.assign(p3, p3 + 3 * (size_t)nf);
  F.asso.assign(asso, asso + nf);
  F.track.assign(nf, -1); F.pos.assign(nf, 0);   // linked by the solver thread
  memcpy(F.Twc, Twc, sizeof(float) * 16); memcpy(F.Twc_rf, Twc, sizeof(float) * 16); memcpy(F.rel, rel, sizeof(float) * 16);
  if (!ts->lo_corres.empty() || !ff.ob_sem.empty()) {   // RenewFrameInfo (object part), Map bookkeeping, dynamic tracklets
    rc = dyn_renew(ctx, ff, curTcw, FS.d_obkeys + 2 * (size_t)slot * ts->obj_cap, d_depth, d_flow, d_mask, slot, D, F);
    if (rc) return rc;
    if (st) st->n_dyn_features = (int)F.ddepth.size();
  }
  ts->last_keys = F.xy; ts->last_depth = F.depth;
  ts->last_corres.assign(corres, corres + 2 * (size_t)nf); ts->last_flow.assign(flow, flow + 2 * (size_t)nf);
  memcpy(ts->lastTcw, Tcw, sizeof(float) * 16);
  memcpy(ts->mVelocity, vel, sizeof(float) * 16);
  ts->has_velocity = true;
  if (ts->map.size() == ts->map.capacity()) {
    rc = ba_async_join(ctx);
    if (rc) return rc;
    ts->map.reserve(std::max<size_t>(4096, 2 * ts->map.capacity()));
  }
  ts->map.push_back(std::move(F));
  if (st) {
    st->n_matches = hdr[1]; st->n_init_inliers = hdr[2]; st->init_winner = hdr[3]; st->n_pose_inliers = hdr[6]; st->n_static = nf;
    st->n_masks_recovered = 0;
  }
  const int window = ts->f_id < c.window_size ? ts->f_id : c.window_size;
  ts->f_id++; ts->frames_seen++;
  rc = ba_async_post(ctx, (int)ts->map.size(), window, st);
  if (rc) return rc;
  // mSegMapLast / mFlowMapLast (see back_end): private copies while the frame carries object features
  ts->have_last_maps = false;
  if (!ts->lo_sem.empty()) {
    cudaStream_t s = ctx->stream;
    VIDO_CUDA(cudaMemcpyAsync(ts->d_last_mask, d_mask + (size_t)slot * px, px * 4, cudaMemcpyDeviceToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync(ts->d_last_flow, d_flow + 2 * (size_t)slot * px, px * 8, cudaMemcpyDeviceToDevice, s));
    // not waited for here: the only reader is the next frame's UpdateMask on this same stream; the slot is not recycled before
    // the next batch, and the host-driven path and the end of the call synchronise this stream first (obj_stream_pending)
    ts->obj_stream_pending = true;
    ts->have_last_maps = true;
  }
  const double th3 = now_ms();
  ts->ht[0] += th2 - th1; ts->ht[1] += (th1 - th0) + (th3 - th2); ts->hn++;
  return VIDO_OK;
}

int trk_prefetch(vido_ctx* ctx, const vido_frame_inputs* in, int nframes) {
  TrackState* ts = (TrackState*)ctx->trk;
  ts->hint.clear();
  if (nframes <= 0) return VIDO_OK;
  const int B = std::min(ts->capB, nframes);
  ts->hint.assign(in, in + B);
  return VIDO_OK;
}

static int trk_track_chunk_impl(vido_ctx* ctx, const vido_frame_inputs* in, int nframes, float* Tcw_out, vido_track_stats* stats) {
  TrackState* ts = (TrackState*)ctx->trk;
  const vido_config& c = ctx->cfg;
  cudaStream_t s = ctx->stream;
  const size_t px = (size_t)c.width * c.height;
  if (ts->call_failed) {   // left by a failed call: discarded.  (Solves a successful call without statistics left queued keep running.)
    ba_async_join(ctx);
    while (ts->ba_nq > 0) { vido_lm_stats ls; ba_collect(ctx, &ts->job[ts->ba_queue[0]].pr, &ls); ts->ba_queue[0] = ts->ba_queue[1]; ts->ba_queue[1] = ts->ba_queue[2]; ts->ba_nq--; }
    ts->call_failed = false;
  }
  ts->ba_deferred.valid = false;   // (a successful call never leaves one behind)
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
    // (On the device-chained path the launch waits until this batch's tracker kernels are queued: they feed the window solver,
    // whose queue of three solves must not run dry at the batch boundary; the look-ahead has the whole batch to finish.)
    bool ahead_done = false;
    auto launch_ahead = [&]() -> int {
      if (ahead_done) return VIDO_OK;
      ahead_done = true;
      const double ta0 = now_ms();
      TrackState::FeSlot& N = ts->fe[ts->fe_cur ^ 1];
      N.launched = false;
      int r = VIDO_OK;
      if (done + B < nframes) r = fe_launch(ctx, N, in + done + B, std::min(ts->capB, nframes - done - B));
      else if (!ts->hint.empty()) { r = fe_launch(ctx, N, ts->hint.data(), (int)ts->hint.size()); ts->hint.clear(); }
      ts->ht[5] += now_ms() - ta0;
      return r;
    };
    const float* d_depth = F.in_depth;
    const double front_ms = (now_ms() - tf0) / B;
    ts->ht[5] += now_ms() - tf0;
    if (ts->vio) { rc = vio_preintegrate_batch(ctx, in + done, B); if (rc) return rc; }
    for (int b = 0; b < B; b++) {
      vido_track_stats* st = stats ? stats + done + b : nullptr;
      ts->cur_t = in[done + b].timestamp;
      if (chain_eligible(ctx, ff[b], in[done + b])) {
        // a run of consecutive eligible frames: queue the kernels of all of them, then consume the records in order
        int e = b;
        if (ts->ba_deferred.valid) {   // a window the host-driven path put off: queue it before the solver thread takes over
          rc = ba_async_join(ctx);
          if (!rc) rc = ba_flush_deferred(ctx);
          if (rc) return rc;
        }
        if (!ts->chain_active) {
          rc = chain_upload_state(ctx, (int)(ts->last_corres.size() / 2), ts->last_keys.data(), ts->last_depth.data(), ts->last_corres.data(),
                                  ts->last_flow.data(), ts->lastTcw, ts->mVelocity, ts->has_velocity ? 1 : 0);
          if (rc) return rc;
          ts->chain_active = true;
        }
        const size_t K = ts->kp_cap;
        const bool hybrid_ok = !getenv("VIDO_NO_HYBRID");
        {
          float Tv = ts->ist.t_init;
          double tp = ts->vio ? ts->fr.back().t : 0.0;
          while (e < B && !in[done + e].write_back_depth && (hybrid_ok || ff[e].ob_sem.empty())) {   // the run [b, e)
            if (ts->vio) {
              if (!ff[e].ob_sem.empty() || !vio_frame_is_plain(ts, &Tv, tp, in[done + e].timestamp, e - b)) break;
              tp = in[done + e].timestamp;
            }
            e++;
          }
          if (e == b) e = b + 1;   // (chain_eligible accepted frame b under the same rule)
        }
        // The tracker kernels are queued a few frames ahead of the record being consumed, not the whole run at once: the first
        // record of a batch is then ~0.5 ms away instead of ~1.3 ms (16 enqueues + the look-ahead launch), short enough for
        // the window solver's queue of three to bridge the batch boundary.
        const int AHEAD = 6;
        int enq = b;
        auto enqueue_upto = [&](int upto) -> int {
          const double te0 = now_ms();
          for (; enq < upto && enq < e; enq++) {
            int r = chain_enqueue_frame(ctx, F.d_kp + enq * K, F.d_nkp + enq, F.d_kpmask + enq * K, F.d_kpdepth + enq * K, F.d_kpflow + 2 * enq * K,
                                        F.in_depth + (size_t)enq * px, F.in_flow + 2 * (size_t)enq * px, F.in_mask + (size_t)enq * px, enq);
            if (r) return r;
          }
          ts->ht[4] += now_ms() - te0;
          return VIDO_OK;
        };
        rc = enqueue_upto(b + 3);
        if (rc) return rc;
        rc = launch_ahead();
        if (rc) return rc;
        for (int k = b; k < e; k++) {
          rc = enqueue_upto(k + 1 + AHEAD);
          if (rc) return rc;
          vido_track_stats* sk = stats ? stats + done + k : nullptr;
          ts->cur_t = in[done + k].timestamp;
          if (!ts->lo_corres.empty() || !ts->lo_sem.empty() || !ff[k].ob_sem.empty()) {
            int nrec = -1;
            rc = hybrid_consume(ctx, F, ff[k], k, Tcw_out + 16 * (size_t)(done + k), sk, &nrec);
            if (rc < 0) return rc;
            if (nrec >= 0) {
              // UpdateMask re-warped a label into this frame's mask: what the chain computed for this and the queued frames
              // used stale front-end products.  Drop it, repeat the mask-dependent front-end stages and the whole frame on the
              // host-driven path; the next frame starts a new run from the host mirrors (unchanged since the last consume).
              const int32_t* hdr; const float *a0, *a1, *a2, *a3, *a4, *a5, *a6, *a7, *a8; const int32_t* a9;
              for (int j = k; j < enq; j++) { rc = chain_wait_record(ctx, j, &hdr, &a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9); if (rc) return rc; }
              ts->chain_active = false;
              rc = fe_redo_frame(ctx, F, k, ff[k]);
              if (rc) return rc;
              rc = back_end(ctx, F, ff[k], k, Tcw_out + 16 * (size_t)(done + k), sk, nrec);
              if (rc < 0) return rc;
              if (sk) { sk->ms_orb = front_ms; sk->ms_assoc = 0; }
              e = k + 1;
              break;
            }
          } else {
            rc = chain_consume(ctx, k, ff[k], Tcw_out + 16 * (size_t)(done + k), sk);
            if (rc < 0) return rc;
          }
          if (sk) { sk->ms_orb = front_ms; sk->ms_assoc = 0; }
        }
        b = e - 1;
        continue;
      }
      ts->chain_active = false;   // the host-driven path owns the state again (its vectors mirror the device state)
      rc = launch_ahead();
      if (rc) return rc;
      rc = back_end(ctx, F, ff[b], b, Tcw_out + 16 * (size_t)(done + b), st);
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
    rc = launch_ahead();
    if (rc) return rc;
    if (ts->obj_stream_pending) { VIDO_CUDA(cudaStreamSynchronize(ts->obj_stream)); ts->obj_stream_pending = false; }   // before the slot (or the caller's buffer) is reused
    ts->fe_cur ^= 1;
    done += B;
  }
  if (getenv("VIDO_HOST_TIMING") && ts->hn > 0) {
    ba_async_join(ctx);   // (debug output only: the solver thread's counters are read below)
    fprintf(stderr, "[host] per frame ms over %ld frames: record wait %.3f, consume %.3f, BA stage+launch %.3f, BA retire wait %.3f, chain enqueue %.3f, front-end collect+launch %.3f | solver thread: link+stage+launch %.3f, retirement %.3f\n",
            ts->hn, ts->ht[0] / ts->hn, ts->ht[1] / ts->hn, ts->ht[2] / ts->hn, ts->ht[3] / ts->hn, ts->ht[4] / ts->hn, ts->ht[5] / ts->hn,
            ts->bt[0] / ts->hn, ts->bt[1] / ts->hn);
    for (int k = 0; k < 8; k++) ts->ht[k] = 0;
    ts->bt[0] = ts->bt[1] = 0;
    ts->hn = 0;
  }
  if (getenv("VIDO_HOST_TIMING") && ts->dn > 0) {
    fprintf(stderr, "[host-path] per frame ms over %ld frames: mask update %.3f, PnP %.3f, pose-opt %.3f, carry-over %.3f, object tracking %.3f, object motions %.3f, static renewal %.3f, object renewal %.3f, last-map copies %.3f\n",
            ts->dn, ts->dt[0] / ts->dn, ts->dt[1] / ts->dn, ts->dt[2] / ts->dn, ts->dt[3] / ts->dn, ts->dt[4] / ts->dn, ts->dt[5] / ts->dn, ts->dt[6] / ts->dn, ts->dt[7] / ts->dn, ts->dt[8] / ts->dn);
    for (int k = 0; k < 12; k++) ts->dt[k] = 0;
    ts->dn = 0;
  }
  {
    // With a statistics array the call drains the solver queue: stats and Map are final when it returns.  Without one the last
    // (up to three) window solves stay queued on the solver stream and the next call continues behind them -- a stream of calls
    // then never empties the pipeline; every Map accessor, vido_sync, FullBatch and the stand-alone BA entry drain first.
    if (!(stats || ts->vio)) {
      if (!ts->ba_deferred.valid) return VIDO_OK;   // (only the host-driven path leaves a deferred window; the solver thread is idle then)
      int rc = ba_async_join(ctx);
      return rc ? rc : ba_flush_deferred(ctx);
    }
    int rc = ba_async_join(ctx);
    if (!rc) rc = ba_flush_deferred(ctx);
    while (ts->ba_nq > 0 && rc == VIDO_OK) { rc = ba_finish(ctx); ba_writeback_rest(ctx); }
    return rc;
  }
}

int trk_drain(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts) return VIDO_OK;
  if (ts->call_failed) { ba_async_join(ctx); return VIDO_OK; }   // the next call discards what the failed one left
  int rc = ba_async_join(ctx);
  if (!rc) rc = ba_flush_deferred(ctx);
  while (ts->ba_nq > 0 && rc == VIDO_OK) { rc = ba_finish(ctx); ba_writeback_rest(ctx); }
  return rc;
}

// The caller's per-frame statistics array only lives for the duration of the call: whatever a failed call leaves queued
// (a deferred window, solves in flight) must not keep pointers into it -- later entry points (the next chunk, FullBatch,
// ApplyScaledRotation) retire that work and would otherwise write through them.
int trk_track_chunk(vido_ctx* ctx, const vido_frame_inputs* in, int nframes, float* Tcw_out, vido_track_stats* stats) {
  TrackState* ts = (TrackState*)ctx->trk;
  const int rc = trk_track_chunk_impl(ctx, in, nframes, Tcw_out, stats);
  if (rc < 0) {   // nothing of a failed call may still read the caller's buffers or write its statistics
    { std::lock_guard<std::mutex> lk(ts->ba_mu); ts->ba_jobs.clear(); }
    ba_async_join(ctx);
    cudaStreamSynchronize(ctx->stream);
    if (ts->obj_stream) cudaStreamSynchronize(ts->obj_stream);
    ts->obj_stream_pending = false;
    ts->call_failed = true;
  }
  if (stats || rc < 0) {   // (without a statistics array every queued pointer is null already and the solver thread may be running)
    ts->ba_deferred.st = nullptr;
    ts->job[0].st = nullptr;
    ts->job[1].st = nullptr;
    ts->job[2].st = nullptr;
  }
  return rc;
}

// the ORB workspace is shared with the stand-alone extraction entry points: wait for a front-end running ahead
void trk_quiesce(vido_ctx* ctx) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (ts && ts->fe_stream) cudaStreamSynchronize(ts->fe_stream);
}

int trk_num_frames(vido_ctx* ctx) { return (int)((TrackState*)ctx->trk)->map.size(); }

int trk_get_map_poses(vido_ctx* ctx, float* poses, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  trk_drain(ctx);
  const int n = (int)ts->map.size();
  for (int i = 0; i < n && i < cap; i++) memcpy(poses + 16 * (size_t)i, ts->map[i].Twc, sizeof(float) * 16);
  return n;
}

int trk_get_static(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  trk_drain(ctx);
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

int trk_get_dynamic(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int32_t* label, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  trk_drain(ctx);
  if (frame < 0 || frame >= (int)ts->map.size()) return -1;
  const MapFrame& F = ts->map[frame];
  const int n = (int)F.ddepth.size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = F.dxy[2 * i]; xy[2 * i + 1] = F.dxy[2 * i + 1];
    depth[i] = F.ddepth[i];
    p3[3 * i] = F.dp3[3 * i]; p3[3 * i + 1] = F.dp3[3 * i + 1]; p3[3 * i + 2] = F.dp3[3 * i + 2];
    asso[i] = F.dasso[i];
    label[i] = F.dlabel[i];
  }
  return n;
}

int trk_get_objects(vido_ctx* ctx, int frame, int32_t* label, int32_t* sem_label, float* motion, float* centre, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  trk_drain(ctx);
  if (frame < 1 || frame >= (int)ts->map.size()) return -1;
  const MapFrame& F = ts->map[frame];
  const int n = (int)F.objects.size();
  for (int i = 0; i < n && i < cap; i++) {
    label[i] = F.objects[i].label; sem_label[i] = F.objects[i].sem;
    memcpy(motion + 16 * (size_t)i, F.objects[i].motion, sizeof(float) * 16);
    memcpy(centre + 3 * (size_t)i, F.objects[i].centre, sizeof(float) * 3);
  }
  return n;
}

int trk_get_dyn_tracks(vido_ctx* ctx, int32_t* len, int32_t* obj_id, int32_t* first_frame, int32_t* first_feat, int cap) {
  trk_drain(ctx);
  TrackState* ts = (TrackState*)ctx->trk;
  const int n = (int)ts->dyn_tracks.size();
  for (int i = 0; i < n && i < cap; i++) {
    len[i] = ts->dyn_tracks[i].len; obj_id[i] = ts->dyn_tracks[i].obj_id;
    first_frame[i] = ts->dyn_tracks[i].first_frame; first_feat[i] = ts->dyn_tracks[i].first_feat;
  }
  return n;
}

// ---------------------------------------------------------------------------------------------------------
// Optimizer::FullBatchOptimization on the Map: flat graph (Optimizer.cc:1235-1745), device solve (fba_kernels.cu),
// write-back (Optimizer.cc:2090-2176)
// ---------------------------------------------------------------------------------------------------------
namespace {
struct FullGraph {
  std::vector<float> se3, points, e6_meas, obs_xyz;
  std::vector<int32_t> e6_i, e6_j, e6_kind, obs_se3, obs_point, obs_kind, tern_p1, tern_p2, tern_h;
  int n_poses = 0, n_motions = 0;
  std::vector<std::vector<int>> VertexID;        // [frame][object entry] -> SE3 vertex
  std::vector<std::vector<int>> makSta, makDyn;  // per frame, per feature: point vertex or -1
};

void build_full_graph(TrackState* ts, const vido_config& c, FullGraph& G) {
  const int N = (int)ts->map.size();
  const float invfx = 1.0f / c.fx, invfy = 1.0f / c.fy;
  G.n_poses = N;
  G.VertexID.assign(N, std::vector<int>());
  G.makSta.resize(N); G.makDyn.resize(N);
  for (int i = 0; i < N; i++) G.se3.insert(G.se3.end(), ts->map[i].Twc, ts->map[i].Twc + 16);
  int next_se3 = N;
  float I16[16];
  eye44(I16);
  auto add_point = [&](const float* Xw) { G.points.push_back(Xw[0]); G.points.push_back(Xw[1]); G.points.push_back(Xw[2]); return (int)(G.points.size() / 3) - 1; };
  auto add_obs = [&](int se3v, int pt, int kind, float u, float v, float z) {
    G.obs_se3.push_back(se3v); G.obs_point.push_back(pt); G.obs_kind.push_back(kind);
    G.obs_xyz.push_back((u - c.cx) * z * invfx); G.obs_xyz.push_back((v - c.cy) * z * invfy); G.obs_xyz.push_back(z);
  };
  for (int i = 0; i < N; i++) {
    const MapFrame& F = ts->map[i];
    if (i != 0) {
      G.e6_i.push_back(i - 1); G.e6_j.push_back(i); G.e6_kind.push_back(0);
      G.e6_meas.insert(G.e6_meas.end(), F.rel, F.rel + 16);
    }
    // static tracklets of length >= 3: one point per tracklet, one observation per element
    const int ns = (int)F.depth.size();
    G.makSta[i].assign(ns, -1);
    for (int j = 0; j < ns; j++) {
      const int t = F.track[j];
      if (t < 0 || ts->tracks[t].len < 3) continue;
      int pid;
      if (F.pos[j] == 0) pid = add_point(&F.p3[3 * (size_t)j]);
      else pid = G.makSta[i - 1][F.asso[j]];
      if (pid == -1) continue;
      add_obs(i, pid, 0, F.xy[2 * j], F.xy[2 * j + 1], F.depth[j]);
      G.makSta[i][j] = pid;
    }
    // dynamic tracklets: every element is its own point; consecutive elements are tied by the object's motion
    const int nd = (int)F.ddepth.size();
    G.makDyn[i].assign(nd, -1);
    if (i == 0) {
      for (int j = 0; j < nd; j++) {
        const int t = F.dtrack[j];
        if (t < 0 || ts->dyn_tracks[t].len < 3) continue;
        const int pid = add_point(&F.dp3[3 * (size_t)j]);
        add_obs(i, pid, 1, F.dxy[2 * j], F.dxy[2 * j + 1], F.ddepth[j]);
        G.makDyn[i][j] = pid;
      }
      continue;
    }
    const size_t nobj = F.objects.size();
    G.VertexID[i].assign(nobj, -1);
    for (size_t j = 0; j < nobj; j++) {
      G.se3.insert(G.se3.end(), I16, I16 + 16);   // object motions start from identity (Optimizer.cc:1597)
      const int vid = next_se3++;
      if (i > 2) {  // SMOOTH_CONSTRAINT && i>2 (Optimizer.cc:1611-1638)
        const MapFrame& Pf = ts->map[i - 1];
        int TraceID = -1;
        for (size_t k = 0; k < Pf.objects.size(); k++)
          if (Pf.objects[k].label == F.objects[j].label) { TraceID = (int)k; break; }
        if (TraceID != -1) {
          G.e6_i.push_back(G.VertexID[i - 1][TraceID]); G.e6_j.push_back(vid); G.e6_kind.push_back(1);
          G.e6_meas.insert(G.e6_meas.end(), I16, I16 + 16);
        }
      }
      G.VertexID[i][j] = vid;
    }
    for (int j = 0; j < nd; j++) {
      const int t = F.dtrack[j];
      if (t < 0 || ts->dyn_tracks[t].len < 3) continue;
      const int pos = i - ts->dyn_tracks[t].first_frame;
      int ObjPositionID = -1;
      for (size_t k = 0; k < nobj; k++)
        if (F.objects[k].label == ts->dyn_tracks[t].obj_id) { ObjPositionID = G.VertexID[i][k]; break; }
      if (ObjPositionID == -1 && pos != 0) continue;
      int prev = -1;
      if (pos != 0) {
        prev = G.makDyn[i - 1][F.dasso[j]];
        if (prev == -1) continue;
      }
      const int pid = add_point(&F.dp3[3 * (size_t)j]);
      add_obs(i, pid, 1, F.dxy[2 * j], F.dxy[2 * j + 1], F.ddepth[j]);
      if (pos != 0) { G.tern_p1.push_back(prev); G.tern_p2.push_back(pid); G.tern_h.push_back(ObjPositionID); }
      G.makDyn[i][j] = pid;
    }
  }
  G.n_motions = next_se3 - N;
}

void fill_problem(FullGraph& G, vido_fba_problem& pr) {
  memset(&pr, 0, sizeof pr);
  vido_fba_default_params_impl(&pr);
  pr.n_poses = G.n_poses; pr.n_motions = G.n_motions; pr.n_points = (int)(G.points.size() / 3);
  pr.n_obs = (int)G.obs_se3.size(); pr.n_e6 = (int)G.e6_i.size(); pr.n_tern = (int)G.tern_p1.size();
  pr.se3 = G.se3.data(); pr.points = G.points.data();
  pr.e6_i = G.e6_i.data(); pr.e6_j = G.e6_j.data(); pr.e6_kind = G.e6_kind.data(); pr.e6_meas = G.e6_meas.data();
  pr.obs_se3 = G.obs_se3.data(); pr.obs_point = G.obs_point.data(); pr.obs_kind = G.obs_kind.data(); pr.obs_xyz = G.obs_xyz.data();
  pr.tern_p1 = G.tern_p1.data(); pr.tern_p2 = G.tern_p2.data(); pr.tern_h = G.tern_h.data();
}
}  // namespace

int trk_full_batch(vido_ctx* ctx, vido_lm_stats* stats, int32_t* sizes) {
  TrackState* ts = (TrackState*)ctx->trk;
  int rc = ba_async_join(ctx);
  if (!rc) rc = ba_flush_deferred(ctx);   // window solves left by a failed call
  if (rc) return rc;
  while (ts->ba_nq > 0) { rc = ba_finish(ctx); if (rc) return rc; ba_writeback_rest(ctx); }
  FullGraph G;
  build_full_graph(ts, ctx->cfg, G);
  vido_fba_problem pr;
  fill_problem(G, pr);
  if (sizes) { sizes[0] = pr.n_poses; sizes[1] = pr.n_motions; sizes[2] = pr.n_points; sizes[3] = pr.n_obs; sizes[4] = pr.n_e6; sizes[5] = pr.n_tern; }
  const char* save = getenv("VIDO_SAVE_G2O");   // the two graph files of src/Optimizer.cc:1937,1939, on request
  const bool save_files = save && save[0] == '1';
  if (save_files) fba_save_g2o(&pr, "dynamic_slam_graph_before_opt.g2o", 0);
  rc = fba_solve_host(ctx, &pr, stats);
  if (rc) return rc;
  if (save_files) fba_save_g2o(&pr, "dynamic_slam_graph_after_opt.g2o", 0);
  const int N = G.n_poses;
  for (int i = 1; i < N; i++) memcpy(ts->map[i].Twc_rf, &G.se3[16 * (size_t)i], sizeof(float) * 16);
  for (int i = 1; i < N; i++)
    for (size_t j = 0; j < G.VertexID[i].size(); j++) memcpy(ts->map[i].objects[j].motion_rf, &G.se3[16 * (size_t)G.VertexID[i][j]], sizeof(float) * 16);
  for (int i = 0; i < N; i++) {
    MapFrame& F = ts->map[i];
    for (size_t j = 0; j < G.makSta[i].size(); j++)
      if (G.makSta[i][j] != -1) memcpy(&F.p3[3 * j], &G.points[3 * (size_t)G.makSta[i][j]], sizeof(float) * 3);
    for (size_t j = 0; j < G.makDyn[i].size(); j++)
      if (G.makDyn[i][j] != -1) memcpy(&F.dp3[3 * j], &G.points[3 * (size_t)G.makDyn[i][j]], sizeof(float) * 3);
  }
  return VIDO_OK;
}

int trk_get_map_poses_rf(vido_ctx* ctx, float* poses, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  const int n = (int)ts->map.size();
  for (int i = 0; i < n && i < cap; i++) memcpy(poses + 16 * (size_t)i, ts->map[i].Twc_rf, sizeof(float) * 16);
  return n;
}

int trk_get_objects_rf(vido_ctx* ctx, int frame, float* motion, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (frame < 1 || frame >= (int)ts->map.size()) return -1;
  const MapFrame& F = ts->map[frame];
  for (int i = 0; i < (int)F.objects.size() && i < cap; i++) memcpy(motion + 16 * (size_t)i, F.objects[i].motion_rf, sizeof(float) * 16);
  return (int)F.objects.size();
}

int trk_export_full_graph(vido_ctx* ctx, int32_t* sizes, float* se3, float* points, int32_t* e6_i, int32_t* e6_j, int32_t* e6_kind,
                          float* e6_meas, int32_t* obs_se3, int32_t* obs_point, int32_t* obs_kind, float* obs_xyz, int32_t* tern_p1,
                          int32_t* tern_p2, int32_t* tern_h) {
  TrackState* ts = (TrackState*)ctx->trk;
  { int rc = trk_drain(ctx); if (rc) return rc; }
  FullGraph G;
  build_full_graph(ts, ctx->cfg, G);
  sizes[0] = G.n_poses; sizes[1] = G.n_motions; sizes[2] = (int)(G.points.size() / 3); sizes[3] = (int)G.obs_se3.size();
  sizes[4] = (int)G.e6_i.size(); sizes[5] = (int)G.tern_p1.size();
  if (!se3) return VIDO_OK;
  auto cp = [](auto* dst, const auto& v) { if (dst && !v.empty()) memcpy(dst, v.data(), sizeof(v[0]) * v.size()); };
  cp(se3, G.se3); cp(points, G.points); cp(e6_i, G.e6_i); cp(e6_j, G.e6_j); cp(e6_kind, G.e6_kind); cp(e6_meas, G.e6_meas);
  cp(obs_se3, G.obs_se3); cp(obs_point, G.obs_point); cp(obs_kind, G.obs_kind); cp(obs_xyz, G.obs_xyz);
  cp(tern_p1, G.tern_p1); cp(tern_p2, G.tern_p2); cp(tern_h, G.tern_h);
  return VIDO_OK;
}

// Map::ApplyScaledRotation(R, s, bScaledVel = true, t = 0) (src/Map.cc:55-119): camera poses, rigid motions and points of the
// Map, pose and velocity of the frames of Map::vpFrames.  Every queued window solve is retired first: the reference calls this
// after the PartialBatchOptimization of the frame.
static int trk_apply_scaled_rotation_impl(vido_ctx* ctx, const float* R, float s) {
  TrackState* ts = (TrackState*)ctx->trk;
  ts->chain_active = false;   // the last-frame pose changes below: a chained run re-uploads the state
  int rc = ba_async_join(ctx);
  if (!rc) rc = ba_flush_deferred(ctx);
  while (ts->ba_nq > 0 && rc == VIDO_OK) { rc = ba_finish(ctx); ba_writeback_rest(ctx); }
  if (rc) return rc;
  float Tyw[16];
  eye44(Tyw);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Tyw[4 * r + c] = R[3 * r + c];
  auto rot_pts = [&](std::vector<float>& P) {   // s * Ryw * p + tyw: one gemm with alpha = s, beta = 1 (tyw = 0)
    for (size_t i = 0; i + 2 < P.size(); i += 3) {
      const float x = P[i], y = P[i + 1], z = P[i + 2];
      for (int r = 0; r < 3; r++)
        P[i + r] = (float)((double)s * ((double)R[3 * r] * x + (double)R[3 * r + 1] * y + (double)R[3 * r + 2] * z) + 0.0);
    }
  };
  auto scale_pose = [&](float* pose) {   // pose.t *= s ; pose = Tyw * pose
    pose[3] *= s; pose[7] *= s; pose[11] *= s;
    mul44(Tyw, pose, pose);
  };
  auto frame_pose = [&](float* Tcw) {   // Twc.t *= s ; Tyc = Tyw * Twc ; SetPose(Tyc^-1)
    float Twc[16], Tyc[16];
    inv44(Tcw, Twc);
    Twc[3] *= s; Twc[7] *= s; Twc[11] *= s;
    mul44(Tyw, Twc, Tyc);
    inv44(Tyc, Tcw);
  };
  if (ts->vio) {
    for (size_t i = 1; i < ts->fr.size(); i++) {
      frame_pose(ts->fr[i].Tcw);
      mv3(R, ts->fr[i].vel, ts->fr[i].vel, (double)s);   // Ryw * Vw * s
    }
    if (ts->fr.size() > 1) memcpy(ts->lastTcw, ts->fr.back().Tcw, sizeof(float) * 16);
  } else if (ts->initialised) {
    frame_pose(ts->lastTcw);
  }
  for (size_t f = 0; f < ts->map.size(); f++) {
    MapFrame& F = ts->map[f];
    rot_pts(F.p3); rot_pts(F.dp3);
    scale_pose(F.Twc);
    if (f > 0) scale_pose(F.rel);
    for (ObjEntry& o : F.objects) scale_pose(o.motion);
  }
  return VIDO_OK;
}
int trk_apply_scaled_rotation(vido_ctx* ctx, const float* R, float s) { return trk_apply_scaled_rotation_impl(ctx, R, s); }

// Tracking::ParseIMUParamFile (Tracking.cc:174-275) -> IMU::Calib::Set (ImuTypes.cc:476-500): switches the driver to IMU_RGBD
int trk_set_imu(vido_ctx* ctx, const float* Tbc, const float* noise) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (ts->initialised) { ctx->err = "set_imu: the sequence has already started (call vido_track_reset first)"; return VIDO_ERR_STATE; }
  if (!Tbc) { ts->vio = false; return VIDO_OK; }   // back to sensor = RGBD
  ts->vio = true;
  memcpy(ts->Tbc, Tbc, sizeof ts->Tbc);
  memcpy(ts->imu_noise, noise, sizeof ts->imu_noise);
  eye44(ts->Tcb);
  float Rt[9], tb[3] = {Tbc[3], Tbc[7], Tbc[11]}, tc[3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { Rt[3 * r + c] = Tbc[4 * c + r]; ts->Tcb[4 * r + c] = Tbc[4 * c + r]; }
  mv3(Rt, tb, tc, -1.0);
  for (int r = 0; r < 3; r++) ts->Tcb[4 * r + 3] = tc[r];
  memset(&ts->ist, 0, sizeof ts->ist);
  ts->ist.scale = 1.0; ts->ist.status = -1;
  ts->ist.Rwg[0] = ts->ist.Rwg[4] = ts->ist.Rwg[8] = 1.0;
  return VIDO_OK;
}
// Tracking::GrabImuData (Tracking.cc:277-281)
int trk_grab_imu(vido_ctx* ctx, const vido_imu_sample* smp, int n, int frames_ahead) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->vio) { ctx->err = "grab_imu: not in IMU mode (vido_track_set_imu)"; return VIDO_ERR_STATE; }
  const int64_t target = ts->frames_seen + frames_ahead;
  if (!ts->imu_q_frame.empty() && target < ts->imu_q_frame.back()) { ctx->err = "grab_imu: deliveries must be in frame order"; return VIDO_ERR_ARG; }
  ts->imu_q.insert(ts->imu_q.end(), smp, smp + n);
  ts->imu_q_frame.insert(ts->imu_q_frame.end(), (size_t)n, target);
  return VIDO_OK;
}
int trk_get_imu_state(vido_ctx* ctx, vido_imu_state* out) {
  TrackState* ts = (TrackState*)ctx->trk;
  if (!ts->vio) { ctx->err = "not in IMU mode"; return VIDO_ERR_STATE; }
  *out = ts->ist;
  return VIDO_OK;
}
int trk_get_imu_frames(vido_ctx* ctx, float* Tcw, float* vel, float* bias, int cap) {
  TrackState* ts = (TrackState*)ctx->trk;
  const int n = (int)ts->fr.size();
  for (int i = 0; i < n && i < cap; i++) {
    if (Tcw) memcpy(Tcw + 16 * (size_t)i, ts->fr[i].Tcw, sizeof(float) * 16);
    if (vel) memcpy(vel + 3 * (size_t)i, ts->fr[i].vel, sizeof(float) * 3);
    if (bias) memcpy(bias + 6 * (size_t)i, ts->fr[i].bias, sizeof(float) * 6);
  }
  return n;
}

// Tracking::GetMetricError on the Map (metric_kernels.cu): camera poses and the estimated object motions in Map order
int trk_metric_error(vido_ctx* ctx, const float* cam_gt, int n_gt, int refined, const float* pose_pre, const float* mot_gt, int n_obj,
                     vido_metric* out, float* per_item) {
  TrackState* ts = (TrackState*)ctx->trk;
  { int rc = trk_drain(ctx); if (rc) return rc; }
  const int n = (int)ts->map.size();
  if (n_gt < n) { ctx->err = "metric: fewer ground-truth poses than map frames"; return VIDO_ERR_ARG; }
  std::vector<float> cam(16 * (size_t)std::max(n, 1)), mot;
  for (int i = 0; i < n; i++) memcpy(&cam[16 * (size_t)i], refined ? ts->map[i].Twc_rf : ts->map[i].Twc, sizeof(float) * 16);
  for (int i = 1; i < n; i++)
    for (const ObjEntry& o : ts->map[i].objects) mot.insert(mot.end(), refined ? o.motion_rf : o.motion, (refined ? o.motion_rf : o.motion) + 16);
  const int have = (int)(mot.size() / 16);
  if (n_obj > 0 && n_obj != have) { ctx->err = "metric: object ground truth does not match the number of estimated object motions"; return VIDO_ERR_ARG; }
  return metric_error_host(ctx, cam.data(), cam_gt, n, mot.data(), pose_pre, mot_gt, n_obj, out, per_item);
}
