/*
 * vido_b200.h -- C-ABI of libvido_b200.so: the B200 (sm_100a) implementation of VIDO-SLAM's
 * per-frame tracking-and-optimisation hot path.
 *
 * The reference (bxh1/VIDO-SLAM) has no FFI: its hot path lives behind C++ classes of
 * libvido_slam.so.  Each entry point below replaces one of those C++ interfaces (cited as
 * file:line under /root/reference/vido_slam/); the C++ host facade (vido-slam_b200/host/System.h)
 * keeps the reference's System/Tracking/Optimizer call surface and forwards to these functions.
 *
 * Conventions: plain C structs, caller-owned buffers, int status (0 ok, <0 error; text via
 * vido_last_error), no exceptions across the boundary, one context per CUDA device, thread-
 * compatible (not thread-safe), no CPU fallback: every call fails with VIDO_ERR_CUDA when no
 * sm_100 device / driver is present.  "_dev" variants take device pointers (inputs already in HBM).
 */
#ifndef VIDO_B200_H
#define VIDO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIDO_OK 0
#define VIDO_ERR_ARG (-1)
#define VIDO_ERR_CUDA (-2)
#define VIDO_ERR_CAPACITY (-3)
#define VIDO_ERR_STATE (-4)

typedef struct vido_ctx vido_ctx;

/* cv::KeyPoint mirror: pt.x, pt.y, size, angle, response, octave (include/ORBextractor.h, cv::KeyPoint) */
typedef struct vido_keypoint {
  float x, y, size, angle, response;
  int32_t octave;
} vido_keypoint;

/* Settings the reference reads from its YAML in Tracking::Tracking (src/Tracking.cc:45-171). */
typedef struct vido_config {
  int32_t width, height;          /* Camera.width / Camera.height */
  float fx, fy, cx, cy, bf;       /* Camera.fx .. Camera.bf */
  int32_t choose_data;            /* ChooseData: 1 OMD, 2 KITTI, 3 KAIST */
  float depth_map_factor;         /* DepthMapFactor */
  float th_depth_bg, th_depth_obj;/* ThDepthBG / ThDepthOBJ */
  int32_t max_track_bg, max_track_obj; /* MaxTrackPointBG / MaxTrackPointOBJ */
  int32_t window_size;            /* WINDOW_SIZE */
  int32_t nfeatures;              /* ORBextractor.nFeatures */
  float scale_factor;             /* ORBextractor.scaleFactor */
  int32_t nlevels;                /* ORBextractor.nLevels (<= 8) */
  int32_t ini_th_fast, min_th_fast; /* ORBextractor.iniThFAST / minThFAST */
  int32_t rgb;                    /* Camera.RGB: 1 = RGB order, 0 = BGR (3-channel input only) */
  int32_t max_batch;              /* frames the ORB front-end processes per launch (>=1) */
  int32_t device;                 /* CUDA device ordinal */
  float sf_mg_thres, sf_ds_thres; /* SFMgThres / SFDsThres: scene-flow magnitude and distribution thresholds of
                                     Tracking::DynObjTracking (src/Tracking.cc:159-160, 1746-1783) */
  int32_t b_joint;                /* Tracking::bJoint (never initialised in the reference, include/Tracking.h:184; default 1):
                                     1 = PoseOptimizationFlow2Cam / Flow2, 0 = PoseOptimizationNew / ObjMot (src/Tracking.cc:
                                     1133-1136, 1268-1274) without the time-seeded depth noise of that branch */
} vido_config;

void vido_default_config(vido_config* cfg);  /* reference's kitti_config.yaml values at 1242x375 */

/* lifetime: replaces System::Init -> new Tracking / new ORBextractor (src/System.cc:23-48, src/Tracking.cc:39-172) */
vido_ctx* vido_create(const vido_config* cfg);
void vido_destroy(vido_ctx* ctx);
const char* vido_last_error(vido_ctx* ctx); /* ctx may be NULL: error of the failed vido_create */
int vido_version(void);
/* number of this library's kernels launched since creation (bench.py's gpu_launches) */
int64_t vido_kernel_launches(vido_ctx* ctx);
/* pyramid geometry (ORBextractor::ComputePyramid, src/ORBextractor.cc:1107-1132) */
int vido_orb_level_info(vido_ctx* ctx, int32_t* w, int32_t* h, int32_t* quota, float* scale);

/*
 * ORB front-end: replaces ORBextractor::operator() (src/ORBextractor.cc:1034-1105), i.e.
 * ComputePyramid + ComputeKeyPointsOctTree (:755-843) + DistributeOctTree (:529-753) + IC_Angle (:67-94).
 * gray: nframes images of height x width CV_8UC1, row stride `stride` bytes, frame stride `frame_stride`.
 * out: nframes * cap_per_frame keypoints (frame f at out + f*cap_per_frame), n_out[f] = count,
 * keypoints bit-identical to the reference order (level 0..7, quad-tree list order).
 */
int vido_orb_extract(vido_ctx* ctx, const uint8_t* gray, int nframes, size_t frame_stride, int stride,
                     vido_keypoint* out, int cap_per_frame, int32_t* n_out);
/* same with gray / out / n_out in device memory; asynchronous on the context stream unless sync!=0 */
int vido_orb_extract_dev(vido_ctx* ctx, const uint8_t* d_gray, int nframes, size_t frame_stride, int stride,
                         vido_keypoint* d_out, int cap_per_frame, int32_t* d_n_out, int sync);
/* 3-channel 8-bit -> gray with OpenCV 4.x fixed-point coefficients (cvtColor in Tracking::GrabImageRGBD,
 * src/Tracking.cc:327-340); rgb order from the config.  Device pointers. */
int vido_bgr_to_gray_dev(vido_ctx* ctx, const uint8_t* d_bgr, int nframes, size_t frame_stride, int stride,
                         uint8_t* d_gray, size_t gray_frame_stride, int gray_stride);
/* debug/inspection: copy pyramid level `level` of batch slot `frame` of the last extraction to host (tight rows) */
int vido_orb_get_level(vido_ctx* ctx, int frame, int level, uint8_t* out);
/* debug/inspection: FAST candidates (x,y relative to minBorder, score) of (frame, level) in reference order */
int vido_orb_get_candidates(vido_ctx* ctx, int frame, int level, int32_t* xs, int32_t* ys, int32_t* scores,
                            int cap, int32_t* n);


/*
 * Descriptor stage of ORBextractor::operator() -- optional: the reference blurs every level (src/ORBextractor.cc:1078-1079) but
 * has the descriptor call commented out (:1086; its Frame::mDescriptors stays uninitialised) and contains no matcher; the
 * tracking entry points below never call these.
 * vido_orb_describe_dev: 7x7 sigma-2 GaussianBlur (BORDER_REFLECT_101, OpenCV's 8-bit fixed-point arithmetic) of every pyramid
 * level of the LAST extraction, then computeOrbDescriptor (:98-137, pattern :140-398 = include/vido_orb_pattern.h) for the key
 * points that extraction wrote: d_kps / d_nkp as filled by vido_orb_extract_dev (frame f at d_kps + f*cap_per_frame),
 * d_desc[(f*cap_per_frame + k)*32 .. +32) = descriptor of key point k of frame f, row order of the reference's _descriptors.
 * Rotated test points that leave the level's buffer (key points closer than 19 px to the top / bottom edge; undefined in the
 * reference) read 0.
 */
int vido_orb_describe_dev(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int nframes, int cap_per_frame,
                          uint8_t* d_desc, int sync);
/* ORBextractor::operator()(image, mask, keypoints, descriptors) with host pointers: vido_orb_extract plus desc
 * [nframes][cap_per_frame][32] */
int vido_orb_extract_describe(vido_ctx* ctx, const uint8_t* gray, int nframes, size_t frame_stride, int stride,
                              vido_keypoint* out, int cap_per_frame, int32_t* n_out, uint8_t* desc);
/* debug/inspection: blurred pyramid level of batch slot `frame` of the last descriptor pass (tight rows) */
int vido_orb_get_blurred_level(vido_ctx* ctx, int frame, int level, uint8_t* out);
/*
 * Brute-force Hamming matching of 256-bit descriptors (north_star's "Hamming match"; distance = ORB-SLAM's
 * ORBmatcher::DescriptorDistance, i.e. the popcount of the XOR; result = cv::BFMatcher(NORM_HAMMING) with k = 2): for every query
 * descriptor the train index of the smallest distance (the lowest index among equals), that distance and the second smallest
 * distance.  VIDO_HAMMING_NONE (0x7fffffff) / index -1 where the train set has fewer than two / no entries.
 * _dev: npairs independent (query set, train set) pairs, set p at d_query + p*query_stride with d_nq[p] descriptors (<= qcap),
 * outputs at [p*qcap + q]; all pointers device memory, 4-byte aligned.
 */
#define VIDO_HAMMING_NONE 0x7fffffff
int vido_hamming_match(vido_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* best_idx,
                       int32_t* best_dist, int32_t* second_dist);
int vido_hamming_match_dev(vido_ctx* ctx, const uint8_t* d_query, size_t query_stride, const int32_t* d_nq, const uint8_t* d_train,
                           size_t train_stride, const int32_t* d_nt, int npairs, int qcap, int32_t* d_best_idx,
                           int32_t* d_best_dist, int32_t* d_second_dist, int sync);

/* ---- Levenberg-Marquardt statistics (g2o G2OBatchStatistics-like, used for parity checks) ---- */
#define VIDO_LM_MAX_RECORDS 320
typedef struct vido_lm_record { double chi2; double lambda; int32_t trials; int32_t pad; } vido_lm_record;
typedef struct vido_lm_stats {
  int32_t iterations;   /* what SparseOptimizer::optimize returns (g2o/core/sparse_optimizer.cpp:354-427); -1: empty graph */
  int32_t n_records, total_trials, pad;
  vido_lm_record rec[VIDO_LM_MAX_RECORDS];  /* robust chi2 / lambda / #trials after each outer iteration */
} vido_lm_stats;

/*
 * Sliding-window graph optimisation: replaces Optimizer::PartialBatchOptimization (src/Optimizer.cc:43-1228) for the
 * flat graph the reference builds at :220-362: n_poses VertexSE3 (estimate = Map::vmCameraPose, float 4x4 Twc),
 * n_poses-1 EdgeSE3 (measurement = Map::vmRigidMotion[i-1][0]), n_points VertexPointXYZ (Map::vp3DPointSta of the
 * first observation) and one EdgeSE3PointXYZ per observation (measurement = Optimizer::Get3DinCamera, :3277-3294).
 * In/out arrays are float32 like the Map; results are written back as at :1056-1142.
 * Host pointers.  Returns VIDO_OK; the iteration count is in stats->iterations.
 */
typedef struct vido_ba_problem {
  int32_t n_poses, n_points, n_obs, pad;
  float* poses;             /* [n_poses][16] in/out */
  float* rel_motion;        /* [n_poses-1][16] in: EdgeSE3 measurement; out: inv(pose[i-1])*pose[i] */
  float* points;            /* [n_points][3] in/out (world) */
  const int32_t* obs_pose;  /* [n_obs] window-relative pose index */
  const int32_t* obs_point; /* [n_obs] point index */
  const float* obs_xyz;     /* [n_obs][3] camera-frame measurement */
  int32_t max_iterations;   /* 100 (:806) */
  float sigma2_cam, sigma2_3d, huber_cam, huber_3d; /* 0.0001, 16, 0.01, 0.01 (:192-216) */
  float gain_threshold;     /* SparseOptimizerTerminateAction gain 1e-3 (:183); <0 disables */
  int32_t fix_first;        /* reserved (the reference's prior edge at :228-237 never fires in its own calls) */
} vido_ba_problem;
void vido_ba_default_params(vido_ba_problem* p);
int vido_ba_partial(vido_ctx* ctx, vido_ba_problem* p, vido_lm_stats* stats);


/*
 * Full-sequence graph optimisation: replaces Optimizer::FullBatchOptimization (src/Optimizer.cc:1235-2178) for the flat graph
 * the reference builds at :1318-1745.  SE3 vertices: n_poses camera poses (estimate = Map::vmCameraPose, Twc) followed by
 * n_motions object motions (identity at :1597).  Point vertices: one per static tracklet, one per element of a dynamic
 * tracklet.  Edges: EdgeSE3Prior on SE3 vertex 0 (information prior_info, no kernel, :1336-1345); EdgeSE3 kind 0 = camera
 * odometry (Map::vmRigidMotion[i-1][0]), kind 1 = object smoothness (identity, :1611-1638); EdgeSE3PointXYZ kind 0 = static,
 * kind 1 = dynamic (measurement = Optimizer::Get3DinCamera); LandmarkMotionTernaryEdge (p1, p2, H).  Every point may be the
 * p2 of at most one and the p1 of at most one ternary edge (the elements of a tracklet form a chain).
 * This flat graph is also the wire format of the keyframe-factor all-gather of the multi-GPU configuration (one sequence per
 * GPU; see vido_fba_pack / INTEGRATION.md).  Host pointers; se3 / points are float32 in/out like the Map
 * (:2090-2176).  Returns VIDO_OK; the iteration count is in stats->iterations.
 */
typedef struct vido_fba_problem {
  int32_t n_poses, n_motions, n_points, n_obs, n_e6, n_tern;
  float* se3;               /* [n_poses + n_motions][16] in/out */
  float* points;            /* [n_points][3] in/out */
  const int32_t* e6_i;      /* [n_e6] SE3 vertex indices */
  const int32_t* e6_j;
  const int32_t* e6_kind;   /* 0 odometry, 1 smoothness */
  const float* e6_meas;     /* [n_e6][16] */
  const int32_t* obs_se3;   /* [n_obs] */
  const int32_t* obs_point;
  const int32_t* obs_kind;  /* 0 static, 1 dynamic */
  const float* obs_xyz;     /* [n_obs][3] */
  const int32_t* tern_p1;   /* [n_tern] point vertex of the previous frame */
  const int32_t* tern_p2;   /* point vertex of the current frame */
  const int32_t* tern_h;    /* SE3 vertex of the object motion */
  int32_t max_iterations;   /* 300 (:1941) */
  float sigma2_cam, sigma2_3d_sta, sigma2_3d_dyn, sigma2_obj, sigma2_smooth; /* 0.0001, 80, 80, 100, 0.001 (:1290-1295) */
  float huber_cam, huber_obj, huber_3d;  /* 0.01 each (:1312) */
  float gain_threshold;     /* SparseOptimizerTerminateAction gain 1e-4 (:1283) */
  float prior_info;         /* 100000 (:1341) */
  int32_t solver;           /* linear solver of (H + lambda I) x = b: 0 automatic, 1 direct (explicit Schur complement +
                               banded tiled Cholesky), 2 matrix-free (implicit Schur complement, preconditioned CG; for long
                               dynamic tracklets / thousands of SE3 vertices).  With solver 2 stats->pad = CG iterations. */
} vido_fba_problem;
void vido_fba_default_params(vido_fba_problem* p);
int vido_ba_full(vido_ctx* ctx, vido_fba_problem* p, vido_lm_stats* stats);
/* The graph as g2o text -- replaces optimizer.save("dynamic_slam_graph_before_opt.g2o" / "..._after_opt.g2o"),
 * src/Optimizer.cc:1937,1939 (g2o/core/optimizable_graph.cpp:589-622; tags g2o/types/types_slam3d.cpp:37-45).  Host only, no
 * context.  precision <= 0: 6 digits like the reference's default std::ostream.  vido_full_batch writes both files into the
 * working directory like the reference when the environment has VIDO_SAVE_G2O=1. */
int vido_fba_save_g2o(const vido_fba_problem* p, const char* path, int precision);


/*
 * Per-frame joint optical-flow + pose optimisation: replaces Optimizer::PoseOptimizationFlow2Cam
 * (src/Optimizer.cc:2622-2824; camera) and Optimizer::PoseOptimizationFlow2 (:3037-3253; one call per object, with
 * Tcw_init = mInitModel, info_prior = 0.5, rounds = 1, its = 200).  One problem = one SE3 vertex + n flow vertices.
 * Several problems are solved by one launch (one CTA each).  Host pointers.
 */
typedef struct vido_poseopt_problem {
  int32_t n, n_inliers;   /* in: matches; out: n - nBad (0 when n < 3, nothing optimised) */
  const float* obs_xy;    /* [n][2] pLastFrame keypoint of the match */
  const float* flow_xy;   /* [n][2] pLastFrame->mvFlowNext */
  const float* depth;     /* [n]    pLastFrame->mvStatDepth */
  float Tcw_init[16];     /* initial transform (pCurFrame->mTcw) */
  float Tcw_last[16];     /* pLastFrame->mTcw */
  float fx, fy, cx, cy;
  float Tcw_out[16];      /* optimised transform, float32 like Converter::toCvMat */
  float* flow_out;        /* [n][2] refined flow of every match (may be NULL) */
  int32_t* inlier;        /* [n] 1 inlier / 0 outlier (may be NULL) */
  float info_flow, info_prior, rp_thres, chi2_th; /* 0.1, 0.3, 0.04, 5.991 (:2680,2706,2624,2725) */
  int32_t rounds, its;    /* 4, 100 (:2725-2726) */
} vido_poseopt_problem;
void vido_poseopt_default_params(vido_poseopt_problem* p);
/* stats: NULL or 4*nproblems entries (entry 4*k + r = round r of problem k) */
int vido_pose_opt_flow2(vido_ctx* ctx, vido_poseopt_problem* problems, int nproblems, vido_lm_stats* stats);


/*
 * Reprojection-only optimisers of the bJoint == false branch (src/Tracking.cc:1133-1136, 1268-1274):
 *   kind 0 replaces Optimizer::PoseOptimizationNew    (src/Optimizer.cc:2180-2334): camera pose, Huber sqrt(rp_thres), 100 its
 *   kind 1 replaces Optimizer::PoseOptimizationObjMot (src/Optimizer.cc:2826-3035): object motion with P = K * Tcw, 200 its
 * One problem = one SE3 vertex + n reprojection edges, one CTA each, several problems per launch.  pts3d are the back-projected
 * points of the last frame as the caller computed them (the reference's kind-0 call adds time-seeded depth noise there; that is
 * the caller's business).  Outliers: chi2 > rp_thres after the optimisation.  Host pointers.
 */
typedef struct vido_projopt_problem {
  int32_t n, kind;
  const float* obs_xy;   /* [n][2] current keypoints */
  const float* pts3d;    /* [n][3] */
  float T_init[16];      /* kind 0: pCurFrame->mTcw; kind 1: inv(Tcw) * mInitModel */
  float fx, fy, cx, cy;  /* kind 0 */
  double P[12];          /* kind 1: 3x4 projection, row-major */
  float rp_thres;        /* 0.01 */
  int32_t its;           /* 100 / 200 */
  float T_out[16];
  int32_t* inlier;       /* [n] (may be NULL) */
  int32_t n_inliers;
} vido_projopt_problem;
void vido_projopt_default_params(vido_projopt_problem* p, int kind);
/* stats: NULL or nproblems entries */
int vido_pose_opt_proj(vido_ctx* ctx, vido_projopt_problem* problems, int nproblems, vido_lm_stats* stats);


/*
 * Initial camera / object model: replaces Tracking::GetInitModelCam (src/Tracking.cc:1914-2028) and GetInitModelObj
 * (:2030-2162): PnP-RANSAC (500 iterations, 0.4 px, confidence 0.98) against the constant-velocity model, the model
 * with more inliers wins.  The RANSAC is the deterministic variant documented in oracle/vido_oracle.h (OpenCV's
 * cv::solvePnPRansac internals are un-vendored and not reproducible bit for bit).  Host pointers.
 */
typedef struct vido_pnp_problem {
  int32_t n;
  int32_t no_motion_model; /* 1: GetInitModelObj for an object without a previous motion (src/Tracking.cc:2143-2151): the RANSAC
                              model is returned whatever its support; Tcw_motion only seeds the minimal solver */
  const float* cur_xy;     /* [n][2] current keypoints */
  const float* pts3d;      /* [n][3] world points of the last frame (Frame::UnprojectStereoStat) */
  const int32_t* valid;    /* [n] 0 where the depth was negative (excluded from RANSAC); may be NULL */
  float Tcw_motion[16];    /* mVelocity * mpLastFrame->mTcw */
  float fx, fy, cx, cy;
  int32_t iters;           /* 500 */
  float reproj_err, confidence; /* 0.4, 0.98 */
  float Tcw_out[16];
  int32_t* inlier_ids;     /* [n] out: indices of the winning model's inliers, ascending */
  int32_t n_inliers, winner /* 0 RANSAC, 1 motion model */, ransac_inliers, mm_inliers;
} vido_pnp_problem;
void vido_pnp_default_params(vido_pnp_problem* p);
int vido_init_model(vido_ctx* ctx, vido_pnp_problem* p);


/*
 * Per-frame gather stages.  Device pointers ("_dev"): depth f32 [B][H][W], flow f32 [B][H][W][2], mask i32 [B][H][W],
 * tight rows.  raw_depth != 0: the depth map still holds the caller's raw values and the conversion of
 * Tracking::GrabImageRGBD (src/Tracking.cc:299-322) is applied on the fly at the gather points (bit-identical).
 */
/* in-place depth pre-scale (src/Tracking.cc:299-322) */
int vido_depth_prep_dev(vido_ctx* ctx, float* d_depth, int nframes, size_t frame_stride_elems, int stride_elems);
/* static association of Frame::Frame (src/Frame.cc:72-100,164-177) for nframes frames; keypoints as produced by
 * vido_orb_extract_dev.  Outputs per frame f at offset f*out_cap: index into the frame's keypoints, corres xy,
 * flow xy, depth; d_n[f] = count */
int vido_frame_associate_dev(vido_ctx* ctx, const vido_keypoint* d_kps, const int32_t* d_nkp, int kp_cap,
                             const float* d_depth, const float* d_flow, const int32_t* d_mask, int nframes, int raw_depth,
                             int32_t* d_idx, float* d_corres_xy, float* d_flow_xy, float* d_depth_out, int32_t* d_n, int out_cap);
/* stride-4 object sampling of Frame::Frame (src/Frame.cc:184-211) */
int vido_frame_sample_objects_dev(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int nframes,
                                  int raw_depth, float* d_keys_xy, float* d_corres_xy, float* d_flow_xy, float* d_depth_out,
                                  int32_t* d_label, int32_t* d_n, int out_cap);
/* (mask, depth, flow) at truncated query coordinates of frame `frame` (lookups of src/Tracking.cc:369-421,2976-3010);
 * queries outside the image return mask -1 */
int vido_gather_dev(vido_ctx* ctx, const float* d_depth, const float* d_flow, const int32_t* d_mask, int frame, int raw_depth,
                    const float* d_xy, int n, int32_t* d_mask_out, float* d_depth_out, float* d_flow_out);
/* Tracking::UpdateMask (src/Tracking.cc:3291-3357).  sem_label / corres_xy (HOST, n object features of the last frame:
 * mpLastFrame->vSemObjLabel, mvObjCorres); d_mask_last / d_flow_last: last frame's mask and flow, d_mask_cur: the new
 * frame's mask, updated in place (device, tight rows of cfg.width).  For every semantic label in ascending order the
 * current mask votes at the predicted positions; a lost mask (majority 0 with >= 100 votes) is forward-warped from the
 * last frame.  Labels must lie in [0, 4096).  Returns the number of unique labels (uniq_out / recovered, optional,
 * receive them and whether each was warped) or a negative VIDO_ERR_*. */
int vido_update_mask_dev(vido_ctx* ctx, const int32_t* sem_label, const float* corres_xy, int n, const int32_t* d_mask_last,
                         const float* d_flow_last, int32_t* d_mask_cur, int32_t* uniq_out, int32_t* recovered, int cap);
/* KAIST depth scale mScale (src/Tracking.cc:318), 1 by default */
int vido_set_depth_scale(vido_ctx* ctx, float mscale);


/*
 * Per-frame driver: replaces System::TrackRGBD -> Tracking::GrabImageRGBD -> Tracking::Track
 * (src/System.cc:51-63, src/Tracking.cc:283-456, 1081-1509) including the PartialBatchOptimization of every frame.
 * This version covers sensor = RGBD (VO), bJoint = true, UseSampleFeature = 0; static scenes and scenes with dynamic
 * objects (UpdateMask, GetSceneFlowObj, DynObjTracking, GetInitModelObj, PoseOptimizationFlow2, object part of
 * RenewFrameInfo, GetDynamicTrackNew: src/Tracking.cc:1160-1308, 1582-1912, 2030-2162, 2615-2720, 3112-3357).  Like the
 * reference (mSegMap is a shallow copy of the caller's mask), a lost object mask is re-warped IN the frame's mask buffer.
 * A chunk of frames is passed at once so that the frame-independent front-end (gray conversion, ORB, association)
 * runs batched; results are identical to calling it frame by frame.
 */
typedef struct vido_frame_inputs {
  const uint8_t* image;   /* height x width x channels, tight rows (CV_8UC1 or CV_8UC3 in Camera.RGB order) */
  int32_t channels;       /* 1 or 3 */
  int32_t on_device;      /* 0: host pointers (copied H2D inside the call), 1: device pointers */
  const float* depth;     /* height x width CV_32F, raw (pre-scaled on the fly as in src/Tracking.cc:299-322) */
  const float* flow;      /* height x width x 2 CV_32FC2 */
  const int32_t* mask;    /* height x width CV_32SC1 */
  int32_t write_back_depth; /* 1: also write the pre-scaled depth back into `depth` like the reference does */
  int32_t pad;
  double timestamp;
} vido_frame_inputs;
typedef struct vido_track_stats {
  double ms_orb, ms_assoc, ms_init, ms_poseopt, ms_renew, ms_ba; /* host wall time per stage (ms_orb: front-end share) */
  int32_t n_keypoints, n_matches, n_init_inliers, init_winner, n_pose_inliers, n_static;
  int32_t ba_iterations, ba_trials, ba_points, ba_obs;
  int32_t n_dyn_features, n_objects, n_objects_ok, n_masks_recovered; /* object features leaving the frame; objects found by
                                DynObjTracking; objects with an estimated motion (bObjStat); labels re-warped by UpdateMask */
} vido_track_stats;
/* Tcw_out: nframes x 16 floats (what TrackRGBD returns per frame); stats may be NULL.  Returns VIDO_OK or <0.
 * With a stats array every window optimisation of the call has finished when it returns.  With stats == NULL the last (up to
 * three) window solves may still be queued on the solver stream: the next call continues behind them, so a stream of calls never
 * empties the pipeline; vido_sync, every vido_map_* accessor, vido_full_batch, vido_metric_error and vido_ba_partial retire them
 * first, so no caller can observe a Map that is not final. */
int vido_track_frames(vido_ctx* ctx, const vido_frame_inputs* frames, int nframes, float* Tcw_out, vido_track_stats* stats);
int vido_track_reset(vido_ctx* ctx);
/* Input staging of the demo (vido_slam/demo/run_vido_slam.cc:114-122) on the device, for nframes tightly packed width x height
 * images each: raw Bayer RG mosaic (8 bit) -> BGR like cv::cvtColor(COLOR_BayerRG2BGR); 16-bit depth -> float like
 * convertTo(CV_32F); 8-bit mask -> int32 like convertTo(CV_32SC1).  Sources: host or device pointers (NULL = skip that
 * conversion); destinations: device pointers (3 * width * height bytes / width * height floats / ints per frame), usable as
 * vido_frame_inputs with on_device = 1 (channels = 3 for the BGR image).  Asynchronous on the context's stream; host sources
 * must stay unchanged until vido_sync or the next synchronising call. */
int vido_convert_raw(vido_ctx* ctx, const uint8_t* bayer, const uint16_t* depth16, const uint8_t* mask8, int nframes, uint8_t* d_bgr,
                     float* d_depth, int32_t* d_mask);
/* The demo's per-frame sequence in one call (vido_slam/demo/run_vido_slam.cc:112-137): raw inputs as the files hold them (HOST
 * pointers: Bayer RG mosaic, 16-bit depth, float32 flow, 8-bit mask) -> vido_convert_raw into buffers the context owns ->
 * vido_track_frames on the device copies (image = the demosaiced BGR, 3 channels).  Same results as converting with OpenCV and
 * calling vido_track_frames; the host->device traffic is 5.6 instead of 8.9 MB per 1280 x 560 frame. */
typedef struct vido_raw_inputs {
  const uint8_t* bayer;      /* height x width */
  const uint16_t* depth16;   /* height x width */
  const float* flow;         /* height x width x 2 */
  const uint8_t* mask8;      /* height x width */
  double timestamp;
} vido_raw_inputs;
int vido_track_raw_frames(vido_ctx* ctx, const vido_raw_inputs* frames, int nframes, float* Tcw_out, vido_track_stats* stats);
/* Optional hint: start copying the HOST frames that the next vido_track_frames call will pass (at most max_batch of them are
 * taken) while the context is busy with the current call.  The copy runs on its own stream into a second set of input
 * buffers; vido_track_frames recognises the frames by the image pointer of the first one.  The host buffers must stay
 * unchanged until that call returns.  No reference counterpart (the reference reads cv::Mat inputs synchronously,
 * src/System.cc:51-63); results are identical with or without the hint. */
int vido_track_prefetch(vido_ctx* ctx, const vido_frame_inputs* frames, int nframes);
/* Map accessors (include/Map.h:44-97): number of frames, vmCameraPose (Twc, refined by the window optimisation),
 * per-frame static features vpFeatSta / vfDepSta / vp3DPointSta / vnAssoSta */
int vido_map_num_frames(vido_ctx* ctx);
int vido_map_get_poses(vido_ctx* ctx, float* poses, int cap);
int vido_map_get_static(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap);
/* Map::vpFeatDyn / vfDepDyn / vp3DPointDyn / vnAssoDyn[frame-1] / vnFeatLabel[frame-1]; returns the count */
int vido_map_get_dynamic(vido_ctx* ctx, int frame, float* xy, float* depth, float* p3, int32_t* asso, int32_t* label, int cap);
/* objects with an estimated motion in frame >= 1: Map::vnRMLabel / vnSMLabel / vmRigidMotion / vmRigidCentre [frame-1][1..]
 * (the camera entry 0 is vido_map_get_poses); motion = world-frame rigid motion H (4x4 float), centre = last-frame centroid */
int vido_map_get_objects(vido_ctx* ctx, int frame, int32_t* label, int32_t* sem_label, float* motion, float* centre, int cap);
/* Map::TrackletDyn / nObjID (Tracking::GetDynamicTrackNew, src/Tracking.cc:2615-2720): per track its length, object id and the
 * (frame, feature) of its first element; returns the number of tracks */
int vido_map_get_dyn_tracks(vido_ctx* ctx, int32_t* len, int32_t* obj_id, int32_t* first_frame, int32_t* first_feat, int cap);


/*
 * Optimizer::FullBatchOptimization on the context's Map (what Tracking::Track does at the stop frame, src/Tracking.cc:
 * 1490-1498): builds the flat graph (vido_fba_problem), solves it on the device and writes the results back like
 * src/Optimizer.cc:2090-2176: refined camera poses (Map::vmCameraPose_RF, vido_map_get_poses_rf), refined object motions
 * (Map::vmRigidMotion_RF, vido_map_get_objects_rf), static and dynamic 3-D points in place (vido_map_get_static / _dynamic).
 * sizes (optional, 6 ints): poses, motions, points, observations, SE3 edges, ternary edges.
 */
int vido_full_batch(vido_ctx* ctx, vido_lm_stats* stats, int32_t* sizes);
int vido_map_get_poses_rf(vido_ctx* ctx, float* poses, int cap);
int vido_map_get_objects_rf(vido_ctx* ctx, int frame, float* motion, int cap);
/* The flat FullBatch graph of the context's Map ("keyframe factors").  Call with se3 == NULL to get the sizes, then with
 * arrays of those sizes.  The index arrays are relative to this sequence; vido_fba_problem documents every field. */
int vido_map_export_full_graph(vido_ctx* ctx, int32_t* sizes, float* se3, float* points, int32_t* e6_i, int32_t* e6_j,
                               int32_t* e6_kind, float* e6_meas, int32_t* obs_se3, int32_t* obs_point, int32_t* obs_kind,
                               float* obs_xyz, int32_t* tern_p1, int32_t* tern_p2, int32_t* tern_h);


/*
 * IMU preintegration: replaces Tracking::PreintegrateIMU (src/Tracking.cc:784-887) + IMU::Preintegrated::
 * IntegrateNewMeasurement (src/ImuTypes.cc:245-300) for njobs frame intervals at once.  samples = the IMU queue
 * (ascending t, System::TrackRGBD's vImuMeas / IMU::Point); job j integrates (t_prev[j], t_cur[j]]; bias[j] =
 * (bax,bay,baz,bwx,bwy,bwz) of the previous frame; noise = (ng, na, ngw, naw) as given to IMU::Calib::Set.
 * Fields mirror IMU::Preintegrated (include/ImuTypes.h:120-233), float32.  Host pointers.
 */
typedef struct vido_imu_sample { double t; float ax, ay, az, wx, wy, wz; } vido_imu_sample;
typedef struct vido_imu_preint {
  float dT;
  float dR[9], dV[3], dP[3];
  float JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
  float C[225];
  float avgA[3], avgW[3];
  int32_t n_steps, n_consumed;
} vido_imu_preint;
int vido_imu_preintegrate(vido_ctx* ctx, const vido_imu_sample* samples, int n, const double* t_prev, const double* t_cur,
                          int njobs, const float* bias, const float* noise, vido_imu_preint* out);

/*
 * Inertial-only optimisation of the VIO initialisation: replaces Optimizer::InertialOptimization (src/Optimizer.cc:2441-2620,
 * called from Tracking::InitializeIMU, src/Tracking.cc:937-1044) with EdgeInertialGS (src/G2oTypes.cc:357-482).  Poses are
 * fixed; unknowns: one velocity per frame, gyro / acc bias (priors prior_g / prior_a), gravity direction Rwg, scale.
 * preint[i] = preintegration from frame i to i+1 (frame i+1's mpImuPreintegrated, e.g. from vido_imu_preintegrate),
 * bias_lin[i] = the bias it was integrated with; Rwb / twb = Frame::GetImuRotation / GetImuPosition (float32).  Host pointers.
 */
typedef struct vido_inertial_problem {
  int32_t n_frames, its;          /* its = 200 (:2444) */
  const float* Rwb;               /* [n][9] */
  const float* twb;               /* [n][3] */
  float* velocity;                /* [n][3] in/out (Frame::mVw) */
  const vido_imu_preint* preint;  /* [n-1] */
  const float* bias_lin;          /* [n-1][6] bax,bay,baz,bwx,bwy,bwz */
  double Rwg[9];                  /* in/out (Eigen::Matrix3d mRwg) */
  double scale;                   /* in/out (mScale) */
  double bg[3], ba[3];            /* in/out (mbg, mba) */
  float prior_g, prior_a;         /* 1e2, 1e9 (src/Tracking.cc:1453) */
  int32_t mode;                   /* 0: the initialisation above (its 200, lambda 1e3); 1: Optimizer::InertialOptimization(Map*,
                                     Rwg, scale) of Tracking::ScaleRefinement (src/Optimizer.cc:2336-2439, src/Tracking.cc:1046-
                                     1077): velocities and biases fixed, no priors, only gravity direction and scale; its = 10 */
} vido_inertial_problem;
void vido_inertial_default_params(vido_inertial_problem* p);
int vido_inertial_opt(vido_ctx* ctx, vido_inertial_problem* p, vido_lm_stats* stats);

/*
 * VIO mode of the per-frame driver (sensor = IMU_RGBD).  vido_track_set_imu replaces Tracking::ParseIMUParamFile (src/Tracking.cc:
 * 174-275): Tbc = camera-to-body transform (row-major 4x4, "Tbc" of the YAML), noise = (ng, na, ngw, naw) as handed to IMU::Calib
 * (already scaled by sqrt(IMU.Frequency), :258-262); must be called before the first frame (or after vido_track_reset);
 * Tbc = noise = NULL switches back to sensor = RGBD.
 * vido_track_grab_imu replaces Tracking::GrabImuData (:277-281) / the vImuMeas argument of System::TrackRGBD (src/System.cc:64-76):
 * the samples System::TrackRGBD would receive together with the frame that lies `frames_ahead` frames after the next one to be
 * tracked (0 = the next frame), so that a whole chunk can be announced before one vido_track_frames call; deliveries in frame order,
 * ascending t.  vido_frame_inputs.timestamp is the frame's time stamp.
 * With IMU data the driver then does what Tracking::Track does (src/Tracking.cc:1115-1119, 1452-1480, 1555-1561): preintegration of
 * every frame against the last frame's bias (one kernel launch per front-end batch), InitializeIMU(1e2, 1e9) after the window
 * optimisation of every frame until it succeeds (>= 10 map frames, >= 2 s; gravity direction from the summed delta-velocities,
 * velocities from finite differences, the inertial-only optimisation above, bias write-back with Reintegrate when the gyro bias
 * moved by > 0.01, Map::ApplyScaledRotation, Tracking::UpdateFrameIMU), and ScaleRefinement (mode 1) in the mTinit windows.
 * The pose returned for a frame is mpCurrentFrame->mTcw after these steps, like System::TrackRGBD.
 */
typedef struct vido_imu_state {
  int32_t initialized;     /* Tracking::mbImuInitialized */
  int32_t status;          /* last InitializeIMU attempt: -1 none, 0 done, 1 too few frames / too little time, 2 scale < 0.1,
                              3 a frame without preintegration breaks the chain (not initialised; the reference would skip the edge) */
  int32_t init_frame;      /* frame id at which the initialisation succeeded */
  int32_t n_refinements;   /* ScaleRefinement calls */
  int32_t n_reintegrated;  /* IMU::Preintegrated::Reintegrate calls */
  int32_t lm_iterations, lm_trials; /* of the initialisation's inertial-only optimisation */
  float t_init;            /* mTinit */
  double scale;            /* mScale of the last inertial-only optimisation */
  double Rwg[9], bg[3], ba[3];
} vido_imu_state;
int vido_track_set_imu(vido_ctx* ctx, const float* Tbc, const float* noise);
int vido_track_grab_imu(vido_ctx* ctx, const vido_imu_sample* samples, int n, int frames_ahead);
int vido_track_get_imu_state(vido_ctx* ctx, vido_imu_state* out);
/* per frame id (the initial frame included): Frame::mTcw (after ApplyScaledRotation / UpdateFrameIMU), mVw, mImuBias (bax..bwz);
 * returns the number of frames */
int vido_map_get_imu_frames(vido_ctx* ctx, float* Tcw, float* vel, float* bias, int cap);
/* Map::ApplyScaledRotation(R, s, bScaledVel = true, t = 0) (src/Map.cc:55-119): every camera pose, rigid motion and 3-D point of
 * the context's Map, the frame poses / velocities (VIO mode) or the last-frame pose, are scaled by s and rotated by R (row-major
 * 3x3); pending window solves are retired first */
int vido_map_apply_scaled_rotation(vido_ctx* ctx, const float* R, float s);

/*
 * Accuracy against ground truth: replaces Tracking::GetMetricError (src/Tracking.cc:3531-3674, bRMSError = false) on the context's
 * Map.  Camera: for every frame i >= 1 the relative pose error  CamPose[i] CamPose[i-1]^-1 * CamPose_gt[i-1] CamPose_gt[i]^-1
 * (translation norm, rotation angle in degrees with the reference's clamped-trace rule), averaged; cam_pose_gt = Map::
 * vmCameraPose_GT (Twc, n_gt >= map frames), refined != 0 evaluates vmCameraPose_RF instead of vmCameraPose.  Objects: for the
 * n_obj estimated object motions in the order vido_map_get_objects returns them over frames 1, 2, ... (refined: the _RF motions),
 * RigMotBody = ObjPosePre^-1 * RigMot * ObjPosePre against RigMot_gt (Map::vmObjPosePre / vmRigidMotion_GT, supplied by the
 * caller; NULL / 0 skips the object part).  per_item (optional): (t, r) of every camera pair, then of every object.
 */
typedef struct vido_metric { float cam_t, cam_r, obj_t, obj_r; int32_t n_cam, n_obj; } vido_metric;
int vido_metric_error(vido_ctx* ctx, const float* cam_pose_gt, int n_gt, int refined, const float* obj_pose_pre,
                      const float* obj_motion_gt, int n_obj, vido_metric* out, float* per_item);

/* accumulated device time (CUDA events on the context stream) of the kernel groups: ms[0] ORB front-end launches,
 * ms[1] init-model kernels, ms[2] pose-optimisation kernel, ms[3] window-BA kernel; launches[k] = timed regions;
 * ba_alg_bytes = algorithmic bytes of the BA launches (296 B per edge per linearisation + 152 B per edge per
 * residual pass) */
int vido_get_kernel_times(vido_ctx* ctx, double* ms, int64_t* launches, double* ba_alg_bytes);

/* stream handle (cudaStream_t) the context launches on, for event timing in bench.py */
void* vido_stream(vido_ctx* ctx);
int vido_sync(vido_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
