// ctx.h -- device context shared by the C-ABI translation units of libvido_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <functional>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/vido_b200.h"

#define VIDO_MAX_LEVELS 8
#define VIDO_EDGE 19       // EDGE_THRESHOLD   (src/ORBextractor.cc:63)
#define VIDO_MINB 16       // EDGE_THRESHOLD-3 (src/ORBextractor.cc:763)

struct OrbLevel {
  int w, h, pitch;            // level size and row pitch (bytes, multiple of 64)
  size_t frame_stride;        // bytes between batch slots of this level
  size_t base;                // byte offset of the level inside the pyramid allocation
  int quota;                  // mnFeaturesPerLevel
  float scale;                // mvScaleFactor
  int nCols, nRows, wCell, hCell;
  int cell_begin, ncells;     // range in the cell table
  int maxBX, maxBY;           // cols-16, rows-16
  int nIni;                   // DistributeOctTree root count
  float hX;
  int boxW, boxH;             // TMA box (bytes x rows) of the FAST tile
  int cand_cap;               // upper bound on FAST candidates of this level (sum of cell slot caps)
  size_t cand_base;           // element offset of this level inside a frame's octree scratch
  int out_base;               // slot offset of this level in the per-frame level-output array
  int out_cap;                // quota + 4
};

struct OrbCell {  // one cv::FAST call of ComputeKeyPointsOctTree (src/ORBextractor.cc:779-819)
  int x0, y0;      // ROI origin in level coordinates
  int rw, rh;      // ROI size (detection zone is its interior minus 3 px)
  int offx, offy;  // j*wCell, i*hCell added to ROI-relative coordinates
  int slot_base;   // first candidate slot of this cell inside a frame's slot array
  int slot_cap;
};

struct vido_ctx {
  vido_config cfg;
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  std::atomic<int64_t> launches{0};   // kernels launched (the window-solver host thread counts too)
  float mscale = 1.f;  // Tracking::mScale (KAIST depth scale)
  // device-time accounting (CUDA events on `stream`), see vido_get_kernel_times
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double t_ms[4] = {0, 0, 0, 0};   // 0 ORB front-end, 1 init model, 2 pose optimisation, 3 window BA
  int64_t t_n[4] = {0, 0, 0, 0};
  double ba_alg_bytes = 0;         // algorithmic bytes of the BA launches so far (SURVEY.md 8d accounting)
  void* raw_stage[3] = {nullptr, nullptr, nullptr};   // vido_convert_raw: device staging of host raw image / depth / mask
  size_t raw_cap[3] = {0, 0, 0};
  void* scratch[3] = {nullptr, nullptr, nullptr};   // grow-only device scratch of the per-call entry points (0 inertial, 1 projection-only, 2 IMU preintegration): vido_scratch()
  size_t scratch_bytes[3] = {0, 0, 0};
  void* fba_arena = nullptr;       // FullBatch device arena, grow-only (fba_kernels.cu)
  size_t fba_arena_bytes = 0;
  void* raw_dev[4] = {nullptr, nullptr, nullptr, nullptr};   // vido_track_raw_frames: converted BGR / depth / flow / mask of one batch
  int raw_dev_frames = 0;

  // ---- ORB front-end ----
  int nlevels = 0;
  OrbLevel lv[VIDO_MAX_LEVELS];
  std::vector<OrbCell> cells;  // host copy
  int cells_per_frame = 0;
  int slots_per_frame = 0;     // candidate slots per frame (u32 each)
  size_t octree_keys_per_frame = 0;
  int out_slots_per_frame = 0; // sum of out_cap
  uint8_t* d_pyr = nullptr;    // all levels, all batch slots
  size_t pyr_bytes = 0;
  OrbCell* d_cells = nullptr;
  uint32_t* d_slots = nullptr;      // [B][slots_per_frame] packed x|y|score
  int32_t* d_cell_count = nullptr;  // [B][cells_per_frame]
  uint32_t* d_oct_keys = nullptr;   // [B][octree_keys_per_frame] ordered candidates per level (global scratch)
  uint16_t* d_oct_perm = nullptr;   // [B][2*octree_keys_per_frame] (only used when a level overflows shared memory)
  uint32_t* d_level_out = nullptr;  // [B][out_slots_per_frame] packed x|y|score of kept keypoints (list order)
  int32_t* d_level_cnt = nullptr;   // [B][nlevels] kept count, and [B][nlevels] candidate count after it
  int32_t* d_xofs[VIDO_MAX_LEVELS] = {};   // resize tables (level l built from l-1)
  int16_t* d_xa[VIDO_MAX_LEVELS] = {};
  int32_t* d_yofs[VIDO_MAX_LEVELS] = {};
  int16_t* d_ya[VIDO_MAX_LEVELS] = {};
  CUtensorMap tmap[VIDO_MAX_LEVELS];
  CUtensorMap* d_tmap = nullptr;     // device copy of the descriptors
  vido_keypoint* d_kp = nullptr;    // [B][kp_cap] staging for the host-pointer API
  int32_t* d_nkp = nullptr;         // [B]
  int kp_cap = 0;
  uint8_t* d_in = nullptr;          // staging for host inputs [B][H][in_pitch]
  int in_pitch = 0;
  int32_t* d_err = nullptr;         // device error flag
  int last_batch = 0;
  int oct_smem_keys = 0;            // shared-memory key capacity of the octree kernel
  size_t oct_smem_bytes = 0;

  // ---- graph optimisation ----
  void* ba = nullptr;  // BaWorkspace (ba_kernels.cu)
  void* po = nullptr;  // PoWorkspace (poseopt_kernels.cu)
  void* pnp = nullptr; // PnpWorkspace (pnp_kernels.cu)
  void* trk = nullptr; // TrackState (track.cu)
  void* chain = nullptr; // ChainWorkspace (chain_kernels.cu)
  void* desc = nullptr;  // DescWorkspace (desc_kernels.cu), created by the first descriptor / matcher call
  // One-shot hook run by the synchronous PnP / pose-optimisation wrappers after their launches and before they wait for the
  // results: the per-frame driver parks host work there (staging and queueing the previous frame's window solve, retiring
  // the one before) so that it overlaps the kernels instead of extending the frame's serial path.
  std::function<void()> idle_work;
  char* um_ws = nullptr;   // UpdateMask workspace (assoc_kernels.cu), grown on demand: cudaMalloc is expensive once peer
  size_t um_ws_bytes = 0;  // access is enabled (NCCL), so nothing on the per-frame path allocates
};

// Grow-only device scratch kept by the context (freed by vido_destroy): a cudaMalloc / cudaFree pair per call costs anything from
// a millisecond to several hundred (measured, driver-side) and synchronises the device.  Doubles on growth.
inline void* vido_scratch(vido_ctx* ctx, int slot, size_t bytes) {
  if (ctx->scratch_bytes[slot] < bytes) {
    if (ctx->scratch[slot]) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->scratch[slot]); }
    ctx->scratch[slot] = nullptr; ctx->scratch_bytes[slot] = 0;
    const size_t want = bytes * 2;
    if (cudaMalloc(&ctx->scratch[slot], want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ctx->scratch_bytes[slot] = want;
  }
  return ctx->scratch[slot];
}


#define VIDO_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      char buf_[512];                                                                       \
      snprintf(buf_, sizeof buf_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      ctx->err = buf_;                                                                      \
      return VIDO_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

// ---- device-resident tracker state of the static VO path (chain_kernels.cu): what Tracking keeps of mpLastFrame
//      (mvStatKeys / mvStatDepth / mvCorres / mvFlowNext, mTcw) and the motion model (mVelocity), as device arrays, so that
//      the per-frame kernels of frame k+1 can be queued behind those of frame k without a host round trip
struct ChainStateDev {
  int32_t* hdr;   // [0] number of features, [1] has_velocity, [2] frame was skipped (lost tracking)
  float *Tcw, *vel, *keys, *depth, *corres, *flow;
};
struct ChainPnpOut { const float* T; const int32_t* ids; const int32_t* res; };   // init model: pose, inlier ids, (n, winner, nr, nm)
struct ChainPoOut { const float* T; const float* flow; const int32_t* inl; const int32_t* ninl; const int32_t* n; };
int pnp_chain_setup(vido_ctx* ctx, int cap);
int pnp_chain_enqueue(vido_ctx* ctx, const ChainStateDev& st, int which, ChainPnpOut* out);
int po_chain_setup(vido_ctx* ctx, int cap);
int po_chain_enqueue(vido_ctx* ctx, const ChainStateDev& st, int which, const ChainPnpOut& pnp, ChainPoOut* out);

// chain_kernels.cu
int chain_setup(vido_ctx* ctx, int nslots);
void chain_teardown(vido_ctx* ctx);
int chain_capacity(vido_ctx* ctx);
int chain_upload_state(vido_ctx* ctx, int n, const float* keys, const float* depth, const float* corres, const float* flow, const float* Tcw,
                       const float* vel, int has_velocity);
int chain_enqueue_frame(vido_ctx* ctx, const vido_keypoint* kp, const int32_t* nkp, const int32_t* kpmask, const float* kpdepth,
                        const float* kpflow, const float* depth, const float* flow, const int32_t* mask, int slot);
int chain_wait_record(vido_ctx* ctx, int slot, const int32_t** hdr, const float** Tcw, const float** Twc, const float** rel, const float** vel,
                      const float** xy, const float** depth, const float** p3, const float** corres, const float** flow, const int32_t** asso);

// desc_kernels.cu
void desc_teardown(vido_ctx* ctx);
int desc_run(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int nframes, int cap_per_frame, uint8_t* d_desc);
uint8_t* desc_staging(vido_ctx* ctx);   // [max_batch][kp_cap][32] bytes
int desc_get_blurred_level(vido_ctx* ctx, int frame, int level, uint8_t* out);
int desc_match_device_api(vido_ctx* ctx, const uint8_t* d_q, size_t q_stride, const int32_t* d_nq, const uint8_t* d_t, size_t t_stride,
                          const int32_t* d_nt, int npairs, int qcap, int32_t* d_best_idx, int32_t* d_best_dist, int32_t* d_second_dist);
int desc_match_host(vido_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* best_idx, int32_t* best_dist,
                    int32_t* second_dist);

// orb_kernels.cu
int orb_setup(vido_ctx* ctx);
void orb_teardown(vido_ctx* ctx);
int orb_run(vido_ctx* ctx, const uint8_t* d_gray, int nframes, size_t frame_stride, int stride,
            vido_keypoint* d_out, int cap_per_frame, int32_t* d_n_out);
int orb_bgr_to_gray(vido_ctx* ctx, const uint8_t* d_bgr, int nframes, size_t frame_stride, int stride,
                    uint8_t* d_gray, size_t gray_frame_stride, int gray_stride);

cudaError_t vido_create_stream(cudaStream_t* s, bool high_priority);  // track.cu

// ba_kernels.cu
int ba_setup(vido_ctx* ctx, int capW, int capP, int capM);
void ba_teardown(vido_ctx* ctx);
int ba_partial_host(vido_ctx* ctx, vido_ba_problem* pr, vido_lm_stats* st);
int ba_submit(vido_ctx* ctx, const vido_ba_problem* pr, bool want_records);
int ba_prepare(vido_ctx* ctx, const vido_ba_problem* pr);
// same, for a problem that will be queued behind the solve in flight: prev_pose[i] / prev_point[l] = index of pose i / point l
// in that solve's problem (or -1); those values are then taken from its output block on the device
int ba_prepare_chained(vido_ctx* ctx, const vido_ba_problem* pr, const int* prev_pose, const int* prev_point);
int ba_launch(vido_ctx* ctx, const vido_ba_problem* pr, bool want_records);
int ba_collect(vido_ctx* ctx, vido_ba_problem* pr, vido_lm_stats* st);
bool ba_oldest_done(vido_ctx* ctx);
int fba_save_g2o(const vido_fba_problem* p, const char* path, int precision);
int input_convert_raw(vido_ctx* ctx, const uint8_t* bayer, const uint16_t* depth16, const uint8_t* mask8, int nframes, uint8_t* d_bgr,
                      float* d_depth, int32_t* d_mask);

// poseopt_kernels.cu
int po_setup(vido_ctx* ctx, int capN, int capProblems);
void po_teardown(vido_ctx* ctx);
int po_flow2_host(vido_ctx* ctx, vido_poseopt_problem* prs, int nproblems, vido_lm_stats* stats);

// pnp_kernels.cu
int pnp_setup(vido_ctx* ctx, int capN, int capIters);
void pnp_teardown(vido_ctx* ctx);
int pnp_init_model_host(vido_ctx* ctx, vido_pnp_problem* p);
int pnp_init_model_batch(vido_ctx* ctx, vido_pnp_problem* ps, int nproblems);  // all problems of a frame in one launch pair

// assoc_kernels.cu
int assoc_depth_prep(vido_ctx* ctx, float* d_depth, int nframes, size_t frame_stride, int stride);
int assoc_update_mask(vido_ctx* ctx, const int32_t* sem_label, const float* corres_xy, int n, const int32_t* d_mask_last,
                      const float* d_flow_last, int32_t* d_mask_cur, int32_t* uniq_out, int32_t* recovered, int cap);
int assoc_frame_associate(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int kp_cap, const float* d_depth,
                          const float* d_flow, const int32_t* d_mask, int nframes, int raw, int32_t* d_idx, float* d_corres,
                          float* d_oflow, float* d_odepth, int32_t* d_n, int out_cap);
int assoc_sample_objects(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int nframes, int raw,
                         float* d_keys, float* d_corres, float* d_oflow, float* d_odepth, int32_t* d_label, int32_t* d_n, int out_cap);
int assoc_gather(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int frame, int raw,
                 const float* d_xy, int n, int32_t* d_omask, float* d_odepth, float* d_oflow);

// track.cu
int trk_setup(vido_ctx* ctx);
void trk_teardown(vido_ctx* ctx);
int trk_reset(vido_ctx* ctx);
int trk_track_chunk(vido_ctx* ctx, const vido_frame_inputs* in, int nframes, float* Tcw_out, vido_track_stats* stats);
int trk_prefetch(vido_ctx* ctx, const vido_frame_inputs* in, int nframes);
void trk_quiesce(vido_ctx* ctx);
int trk_drain(vido_ctx* ctx);   // retire every queued window solve (Map and solver workspace final afterwards)
int trk_num_frames(vido_ctx* ctx);
int trk_get_map_poses(vido_ctx* ctx, float* poses, int cap);
int trk_get_static(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap);
int trk_get_dynamic(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int32_t* label, int cap);
int trk_get_objects(vido_ctx* ctx, int frame, int32_t* label, int32_t* sem_label, float* motion, float* centre, int cap);
int trk_get_dyn_tracks(vido_ctx* ctx, int32_t* len, int32_t* obj_id, int32_t* first_frame, int32_t* first_feat, int cap);

int trk_apply_scaled_rotation(vido_ctx* ctx, const float* R, float s);
int trk_set_imu(vido_ctx* ctx, const float* Tbc, const float* noise);
int trk_grab_imu(vido_ctx* ctx, const vido_imu_sample* smp, int n, int frames_ahead);
int trk_get_imu_state(vido_ctx* ctx, vido_imu_state* out);
int trk_get_imu_frames(vido_ctx* ctx, float* Tcw, float* vel, float* bias, int cap);
int trk_full_batch(vido_ctx* ctx, vido_lm_stats* stats, int32_t* sizes);
int trk_get_map_poses_rf(vido_ctx* ctx, float* poses, int cap);
int trk_get_objects_rf(vido_ctx* ctx, int frame, float* motion, int cap);
int trk_export_full_graph(vido_ctx* ctx, int32_t* sizes, float* se3, float* points, int32_t* e6_i, int32_t* e6_j, int32_t* e6_kind,
                          float* e6_meas, int32_t* obs_se3, int32_t* obs_point, int32_t* obs_kind, float* obs_xyz, int32_t* tern_p1,
                          int32_t* tern_p2, int32_t* tern_h);

// fba_kernels.cu
void vido_fba_default_params_impl(vido_fba_problem* p);
int fba_solve_host(vido_ctx* ctx, vido_fba_problem* p, vido_lm_stats* st);

// projopt_kernels.cu
void projopt_default_params(vido_projopt_problem* p, int kind);
int projopt_host(vido_ctx* ctx, vido_projopt_problem* prs, int nproblems, vido_lm_stats* stats);

// inertial_kernels.cu
void inertial_default_params(vido_inertial_problem* p);
int inertial_opt_host(vido_ctx* ctx, vido_inertial_problem* p, vido_lm_stats* st);

// metric_kernels.cu
int metric_error_host(vido_ctx* ctx, const float* cam, const float* cam_gt, int n, const float* mot, const float* pose_pre,
                      const float* mot_gt, int n_obj, vido_metric* out, float* per_item);
int trk_metric_error(vido_ctx* ctx, const float* cam_gt, int n_gt, int refined, const float* pose_pre, const float* mot_gt, int n_obj,
                     vido_metric* out, float* per_item);   // track.cu

// imu_kernels.cu
int imu_preintegrate_host(vido_ctx* ctx, const vido_imu_sample* samples, int n, const double* t_prev, const double* t_cur,
                          int njobs, const float* bias, const float* noise, vido_imu_preint* out, const int32_t* nvis = nullptr);
