/*
 * vido_oracle.h -- CPU restatement of the VIDO-SLAM tracking/optimisation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  The product (vido-slam_b200/) never links or calls it.
 *
 * Parity status: the reference ships no tests/golden vectors for this path and cannot be
 * built here (no OpenCV/Eigen/CSparse C++), so the oracle is pinned as follows:
 *   - integer image stage (resize, FAST, fastAtan2): bit-exact against cv2 4.13.0 run in the
 *     authoring container (tests/golden/make_orb_golden.py -> tests/golden/orb_*.npz);
 *   - float/double stages (g2o edge math, LM, IMU preintegration): "parity unpinned" by the
 *     reference itself; analytic Jacobians are checked against central differences and the
 *     restatement cites the reference file:line it follows.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/vido_slam/, g2o/ = 3rdparty/g2o/g2o/).
 */
#ifndef VIDO_ORACLE_H
#define VIDO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint mirror (pt.x, pt.y, size, angle, response, octave) */
typedef struct vo_keypoint {
  float x, y, size, angle, response;
  int32_t octave;
} vo_keypoint;

typedef struct vo_orb_params {
  int32_t nfeatures;    /* ORBextractor.nFeatures   (2500) */
  float scale_factor;   /* ORBextractor.scaleFactor (1.2)  */
  int32_t nlevels;      /* ORBextractor.nLevels     (8)    */
  int32_t ini_th_fast;  /* ORBextractor.iniThFAST   (20)   */
  int32_t min_th_fast;  /* ORBextractor.minThFAST   (7)    */
} vo_orb_params;

/* ---- ORB front-end (src/ORBextractor.cc) ---- */
int vo_orb_level_sizes(int W, int H, const vo_orb_params* p, int* w, int* h, float* scale);
int vo_orb_level_quotas(const vo_orb_params* p, int* quota);
void vo_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                         uint8_t* dst, int dw, int dh, int dstride);
/* cv::FAST(img, kps, thr, nonmax=true), TYPE_9_16 on a whole ROI; returns count; xy/score arrays */
int vo_fast_roi(const uint8_t* img, int w, int h, int stride, int thr,
                int* xs, int* ys, int* scores, int cap);
float vo_fast_atan2(float y, float x);
/* candidates of one level in reference order (cell-row-major then FAST scan order);
   coordinates relative to minBorder (as fed to DistributeOctTree) */
int vo_orb_level_candidates(const uint8_t* img, int w, int h, int stride, const vo_orb_params* p,
                            int* xs, int* ys, int* scores, int cap);
/* full ORBextractor::operator(): returns number of keypoints (<= cap); pyramid optional out */
int vo_orb_extract(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p,
                   vo_keypoint* out, int cap);
/* operator() with the descriptor call of src/ORBextractor.cc:1086 enabled: desc[k*32 .. +32) for key point k */
int vo_orb_extract_describe(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p,
                            vo_keypoint* out, int cap, uint8_t* desc);
/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1 (src/ORBextractor.cc:1079), OpenCV's fixed-point path */
void vo_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
/* computeDescriptors (src/ORBextractor.cc:1023-1031) on a blurred level with tight rows: key points in level coordinates */
void vo_orb_describe_level(const uint8_t* blurred, int w, int h, const float* xs, const float* ys, const float* angles, int n,
                           uint8_t* desc);
/* brute-force Hamming matching (cv::BFMatcher(NORM_HAMMING), k = 2): best train index (first minimum), its distance, second distance;
 * 0x7fffffff / -1 where absent */
void vo_hamming_match(const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* best_idx, int32_t* best_dist,
                      int32_t* second_dist);
/* pyramid only: writes level l at out + offsets[l] (tight rows of width w[l]) */
int vo_orb_pyramid(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p,
                   uint8_t* out, int64_t* offsets);


/* ---- Levenberg-Marquardt bookkeeping shared by all graph oracles (g2o/core/optimization_algorithm_levenberg.cpp) ---- */
#define VO_LM_MAX_RECORDS 320
typedef struct vo_lm_record { double chi2; double lambda; int32_t trials; int32_t pad; } vo_lm_record;
typedef struct vo_lm_stats {
  int32_t iterations;      /* return value of SparseOptimizer::optimize */
  int32_t n_records;
  int32_t total_trials;
  int32_t pad;
  vo_lm_record rec[VO_LM_MAX_RECORDS]; /* robust chi2 / lambda after each outer iteration */
} vo_lm_stats;

/*
 * Sliding-window graph of Optimizer::PartialBatchOptimization (src/Optimizer.cc:43-1228) in flat form:
 * W camera poses (VertexSE3, estimate = Twc float 4x4 as stored in Map::vmCameraPose), W-1 odometry edges
 * (EdgeSE3, measurement = Map::vmRigidMotion[i-1][0]), P static points (VertexPointXYZ) and one EdgeSE3PointXYZ
 * per observation (measurement = Optimizer::Get3DinCamera of the feature).  All float32 in / out like the Map.
 */
typedef struct vo_ba_problem {
  int32_t n_poses, n_points, n_obs, pad;
  float* poses;            /* [n_poses][16]  in: vmCameraPose, out: optimised (float32 round trip of :1058-1069) */
  float* rel_motion;       /* [n_poses-1][16] in: EdgeSE3 measurements; out: inv(pose[i-1])*pose[i] (:1072-1075) */
  float* points;           /* [n_points][3] in: vp3DPointSta of the first observation; out: optimised */
  const int32_t* obs_pose; /* [n_obs] window-relative pose index */
  const int32_t* obs_point;/* [n_obs] point index */
  const float* obs_xyz;    /* [n_obs][3] camera-frame measurement */
  int32_t max_iterations;  /* 100 (:806) */
  float sigma2_cam, sigma2_3d, huber_cam, huber_3d; /* 0.0001, 16, 0.01, 0.01 (:192-216) */
  float gain_threshold;    /* 1e-3 (:183) */
  int32_t fix_first;       /* add the EdgeSE3Prior of :228-237 on pose 0 (never true in the reference's own calls) */
} vo_ba_problem;
void vo_ba_default_params(vo_ba_problem* p);
int vo_ba_partial(vo_ba_problem* p, vo_lm_stats* stats);
/* edge math exposed for Jacobian tests: EdgeSE3 (g2o/types/edge_se3.cpp:77-105) */
void vo_edge_se3(const double* Xi /*R row-major 9 + t 3*/, const double* Xj, const double* Z, double err[6],
                 double Ji[36], double Jj[36]);
void vo_se3_oplus(const double* X, const double* upd6, double* Xout); /* VertexSE3::oplusImpl */
void vo_edge_se3_pointxyz(const double* X, const double* p, const double* z, double err[3], double Ji[18], double Jj[9]);

/*
 * Full-sequence graph of Optimizer::FullBatchOptimization (src/Optimizer.cc:1235-2178) in flat form.
 * SE3 vertices: n_poses camera poses (VertexSE3, estimate = Map::vmCameraPose, Twc) followed by n_motions object motions
 * (VertexSE3, initialised to identity, :1594-1598).  Point vertices: one per static track, one per element of a dynamic
 * track.  Edges: EdgeSE3Prior on SE3 vertex 0 (information 1e5, no kernel, :1336-1345); EdgeSE3 kind 0 = camera odometry
 * (sigma2_cam), kind 1 = object smoothness (identity measurement, sigma2_smooth, :1611-1638); EdgeSE3PointXYZ kind 0 =
 * static (sigma2_3d_sta), kind 1 = dynamic (sigma2_3d_dyn); LandmarkMotionTernaryEdge (p1, p2, H), sigma2_obj.
 * Every point is the p2 of at most one and the p1 of at most one ternary edge (the elements of a tracklet form a chain).
 */
typedef struct vo_fba_problem {
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
  const float* obs_xyz;     /* [n_obs][3] camera-frame measurement (Optimizer::Get3DinCamera) */
  const int32_t* tern_p1;   /* [n_tern] point of the previous frame */
  const int32_t* tern_p2;   /* point of the current frame */
  const int32_t* tern_h;    /* SE3 vertex index of the object motion */
  int32_t max_iterations;   /* 300 (:1941) */
  float sigma2_cam, sigma2_3d_sta, sigma2_3d_dyn, sigma2_obj, sigma2_smooth; /* 0.0001, 80, 80, 100, 0.001 (:1290-1295) */
  float huber_cam, huber_obj, huber_3d;  /* 0.01 each (:1312) */
  float gain_threshold;     /* 1e-4 (:1283) */
  float prior_info;         /* 100000 (:1341) */
} vo_fba_problem;
void vo_fba_default_params(vo_fba_problem* p);
int vo_ba_full(vo_fba_problem* p, vo_lm_stats* stats);
/* LandmarkMotionTernaryEdge (g2o/types/types_dyn_slam3d.cpp:53-85): err = p1 - H^-1 p2; J1 = I, J2 (3x3), JH (3x6) */
void vo_edge_landmark_motion(const double* H /*R 9 + t 3*/, const double* p1, const double* p2, double err[3], double J2[9],
                             double JH[18]);
/* EdgeSE3Prior with identity offset (g2o/types/edge_se3_prior.cpp:89-102, isometry3d_gradients.h:265-325) */
void vo_edge_se3_prior(const double* X, const double* Z, double err[6], double J[36]);

/* ---- per-frame stages of Tracking::GrabImageRGBD / Frame::Frame ---- */
/* depth pre-scale, in place (src/Tracking.cc:299-322): negatives -> 0; OMD d/f; KITTI bf/(d/f); KAIST mScale*bf/(d/f) */
void vo_depth_prep(float* depth, int W, int H, int stride_elems, int choose_data, float depth_map_factor, float bf, float mscale);
/* static association of detected keypoints (src/Frame.cc:72-100,164-177): returns count; out_idx = index into kps */
int vo_frame_associate(const vo_keypoint* kps, int n, const float* depth, const float* flow, const int32_t* mask,
                       int W, int H, float th_depth_bg, int32_t* out_idx, float* out_corres_xy, float* out_flow_xy,
                       float* out_depth, int cap);
/* stride-4 object sampling (src/Frame.cc:184-211): returns count */
int vo_frame_sample_objects(const float* depth, const float* flow, const int32_t* mask, int W, int H, float th_depth_obj,
                            float* keys_xy, float* corres_xy, float* flow_xy, float* depth_out, int32_t* label, int cap);

/* Tracking::UpdateMask (src/Tracking.cc:3291-3357): majority vote of the current mask at the predicted object-feature positions
 * per semantic label; a label whose mask was lost (majority 0, >= 100 votes) is forward-warped from the last mask through
 * the last flow.  mask_cur is modified in place; returns the number of unique labels (ascending in uniq_out). */
int vo_update_mask(const int32_t* sem_label, const float* corres_xy, int n, const int32_t* mask_last, const float* flow_last,
                   int32_t* mask_cur, int W, int H, int32_t* uniq_out, int32_t* recovered, int cap);

/*
 * Per-frame camera pose optimisation Optimizer::PoseOptimizationFlow2Cam (src/Optimizer.cc:2622-2824):
 * one VertexSE3Expmap + n marginalised VertexSBAFlow, EdgeSE3ProjectFlow2 + EdgeFlowPrior per match, 4 rounds.
 */
typedef struct vo_poseopt_problem {
  int32_t n, pad;
  const float* obs_xy;    /* [n][2] pLastFrame->mvStatKeys[TM[i]].pt */
  const float* flow_xy;   /* [n][2] pLastFrame->mvFlowNext[TM[i]] */
  const float* depth;     /* [n]    pLastFrame->mvStatDepth[TM[i]] */
  float Tcw_init[16];     /* pCurFrame->mTcw on entry */
  float Tcw_last[16];     /* pLastFrame->mTcw */
  float fx, fy, cx, cy;
  float Tcw_out[16];      /* optimised pose (float32 as Converter::toCvMat) */
  float* flow_out;        /* [n][2] refined flow of every vertex */
  int32_t* inlier;        /* [n] 1 inlier / 0 outlier after the 4th round */
  float info_flow, info_prior, rp_thres, chi2_th; /* 0.1, 0.3, 0.04, 5.991 */
  int32_t rounds, its;    /* 4, 100 */
} vo_poseopt_problem;
void vo_poseopt_default_params(vo_poseopt_problem* p);
int vo_poseopt_flow2cam(vo_poseopt_problem* p, vo_lm_stats* stats /* [rounds] or NULL */);

/*
 * Reprojection-only optimisers of the bJoint == false branch (src/Tracking.cc:1133-1136, 1268-1274):
 *   kind 0  Optimizer::PoseOptimizationNew     (src/Optimizer.cc:2180-2334): camera pose, EdgeSE3ProjectXYZOnlyPose
 *           (g2o/types/types_six_dof_expmap.cpp:266-296), information I, Huber delta = sqrt(rp_thres), optimize(100)
 *   kind 1  Optimizer::PoseOptimizationObjMot  (src/Optimizer.cc:2826-3035): object motion, EdgeSE3ProjectXYZOnlyObjMotion
 *           (types_six_dof_expmap.cpp:394-441) with P = K * Tcw, no kernel, optimize(200)
 * One VertexSE3Expmap, n edges, one round; an edge whose chi2 exceeds rp_thres (0.01) afterwards is an outlier.  The
 * reference feeds kind 0 with depth-noised points (Frame::UnprojectStereoStat(i, 1), time-seeded RNG): pts3d is whatever
 * the caller back-projected, the noise is not part of this function.
 */
typedef struct vo_projopt_problem {
  int32_t n, kind;
  const float* obs_xy;   /* [n][2] current keypoints */
  const float* pts3d;    /* [n][3] Xw (float, as stored in the edge from a float cv::Mat) */
  float T_init[16];      /* vertex estimate at the start (kind 0: pCurFrame->mTcw; kind 1: inv(Tcw) * mInitModel) */
  float fx, fy, cx, cy;  /* kind 0 */
  double P[12];          /* kind 1: 3x4 projection K * Tcw (row-major) */
  float rp_thres;        /* 0.01 */
  int32_t its;           /* 100 / 200 */
  float T_out[16];
  int32_t* inlier;       /* [n] */
  int32_t n_inliers;
} vo_projopt_problem;
void vo_projopt_default_params(vo_projopt_problem* p, int kind);
int vo_pose_opt_proj(vo_projopt_problem* p, vo_lm_stats* stats);

/*
 * Initial camera model of Tracking::GetInitModelCam (src/Tracking.cc:1914-2028): PnP-RANSAC over the previous frame's
 * 3-D points and the current 2-D points versus the constant-velocity model, the one with more inliers wins.
 * cv::solvePnPRansac(500, 0.4 px, 0.98, SOLVEPNP_P3P) lives in un-vendored OpenCV and is not reproducible bit for
 * bit (internal RNG, minimal solver, EPnP refit: "parity unpinned"); the restatement fixes a deterministic variant:
 * counter-based sampling of 4 points, minimal solve by Gauss-Newton from the motion-model pose, OpenCV's adaptive
 * iteration count, refit on the consensus set.  Cross-checked loosely against cv2.solvePnPRansac in the tests.
 */
typedef struct vo_pnp_problem {
  int32_t n;
  int32_t no_motion_model; /* 1: GetInitModelObj without a previous motion of the object (src/Tracking.cc:2143-2151): the RANSAC
                              model is returned whatever its support; Tcw_motion only seeds the minimal solver */
  const float* cur_xy;     /* [n][2] current keypoints */
  const float* pts3d;      /* [n][3] world points of the last frame (UnprojectStereoStat, float) */
  const int32_t* valid;    /* [n] 0 where the depth was negative (excluded from RANSAC) */
  float Tcw_motion[16];    /* mVelocity * mpLastFrame->mTcw (float) */
  float fx, fy, cx, cy;
  int32_t iters;           /* 500 */
  float reproj_err, confidence; /* 0.4, 0.98 */
  float Tcw_out[16];
  int32_t* inlier_ids;     /* [n] indices of the winning model's inliers, ascending */
  int32_t n_inliers, winner /* 0 RANSAC, 1 motion model */, ransac_inliers, mm_inliers;
} vo_pnp_problem;
void vo_pnp_default_params(vo_pnp_problem* p);
int vo_init_model_cam(vo_pnp_problem* p);

/* ---- per-frame driver (VO, static + dynamic objects): System::TrackRGBD -> Tracking::GrabImageRGBD -> Track ---- */
typedef struct vo_track_config {
  int32_t width, height;
  float fx, fy, cx, cy, bf;
  int32_t choose_data;
  float depth_map_factor, th_depth_bg, th_depth_obj;
  int32_t max_track_bg, window_size;
  vo_orb_params orb;
  int32_t rebuild_tracklets; /* 1: rebuild all tracklets from frame 0 every frame like the reference (O(T)/frame) */
  int32_t max_track_obj;     /* MaxTrackPointOBJ (500) */
  float sf_mg_thres, sf_ds_thres; /* SFMgThres 0.12 / SFDsThres 0.3 (src/Tracking.cc:159-160) */
  int32_t b_joint;           /* Tracking::bJoint (uninitialised in the reference): 1 = joint flow + pose optimisers
                                (PoseOptimizationFlow2Cam / Flow2), 0 = reprojection-only (PoseOptimizationNew / ObjMot, without
                                the time-seeded depth noise the reference adds in that branch) */
} vo_track_config;
typedef struct vo_track_stats {
  double ms_orb, ms_assoc, ms_init, ms_poseopt, ms_renew, ms_ba;
  int32_t n_keypoints, n_matches, n_init_inliers, init_winner, n_pose_inliers, n_static;
  int32_t ba_iterations, ba_trials, ba_points, ba_obs;
  int32_t n_dyn_features, n_objects, n_objects_ok, n_masks_recovered; /* object features leaving the frame; objects found by
                                DynObjTracking; objects with an estimated motion (bObjStat); labels re-warped by UpdateMask */
} vo_track_stats;
void* vo_tracker_create(const vo_track_config* cfg);
void vo_tracker_destroy(void* h);
/* depth is modified in place (pre-scale) like the reference; returns 0 ok, 1 frame skipped (< 2 matches) */
int vo_tracker_track(void* h, const uint8_t* gray, float* depth, const float* flow, const int32_t* mask, float* Tcw_out,
                     vo_track_stats* st);
int vo_tracker_num_frames(void* h);
int vo_tracker_get_map_poses(void* h, float* poses /* [n][16] Map::vmCameraPose (Twc, BA-refined) */, int cap);
int vo_tracker_get_static(void* h, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap);
/* Map::vpFeatDyn / vfDepDyn / vp3DPointDyn / vnAssoDyn[frame-1] / vnFeatLabel[frame-1] (include/Map.h:44-97) */
int vo_tracker_get_dynamic(void* h, int frame, float* xy, float* depth, float* p3, int32_t* asso, int32_t* label, int cap);
/* objects of frame >= 1: Map::vnRMLabel / vnSMLabel / vmRigidMotion / vmRigidCentre [frame-1][1..] (camera entry 0 skipped) */
int vo_tracker_get_objects(void* h, int frame, int32_t* label, int32_t* sem_label, float* motion /* [n][16] */,
                           float* centre /* [n][3] */, int cap);
/* Map::TrackletDyn / nObjID (Tracking::GetDynamicTrackNew, src/Tracking.cc:2615-2720): per track its length, object id and
 * first (frame, feature) */
int vo_tracker_get_dyn_tracks(void* h, int32_t* len, int32_t* obj_id, int32_t* first_frame, int32_t* first_feat, int cap);
/* Optimizer::FullBatchOptimization on the tracker's Map (src/Tracking.cc:1490-1498): results in vmCameraPose_RF /
 * vmRigidMotion_RF, points refined in place.  sizes (optional, 6 ints): poses, motions, points, obs, e6, tern.  Returns the
 * iteration count. */
int vo_tracker_full_batch(void* h, vo_lm_stats* stats, int32_t* sizes);
int vo_tracker_get_map_poses_rf(void* h, float* poses, int cap);
int vo_tracker_get_objects_rf(void* h, int frame, float* motion, int cap);
/* the flat FullBatch graph the tracker would solve (arrays sized by a first call with NULL pointers -> sizes) */
int vo_tracker_export_full_graph(void* h, int32_t* sizes, float* se3, float* points, int32_t* e6_i, int32_t* e6_j, int32_t* e6_kind,
                                 float* e6_meas, int32_t* obs_se3, int32_t* obs_point, int32_t* obs_kind, float* obs_xyz,
                                 int32_t* tern_p1, int32_t* tern_p2, int32_t* tern_h);

/* ---- IMU preintegration: Tracking::PreintegrateIMU (src/Tracking.cc:784-887) + IMU::Preintegrated (src/ImuTypes.cc:143-300) ---- */
typedef struct vo_imu_sample { double t; float ax, ay, az, wx, wy, wz; } vo_imu_sample;
typedef struct vo_imu_preint {
  float dT;
  float dR[9], dV[3], dP[3];
  float JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
  float C[225];
  float avgA[3], avgW[3];
  int32_t n_steps, n_consumed; /* integrated steps; samples popped from the queue (the last used sample stays) */
} vo_imu_preint;
/* samples: the IMU queue (ascending t); bias = (bax,bay,baz,bwx,bwy,bwz); noise = (ng, na, ngw, naw) as passed to
 * IMU::Calib::Set.  All arithmetic float32 like the reference's cv::Mat code (products accumulated in double and
 * rounded once, as cv::gemm does for CV_32F); cv::SVDecomp-based NormalizeRotation is replaced by the polar factor. */
int vo_imu_preintegrate(const vo_imu_sample* samples, int n, double t_prev, double t_cur, const float* bias,
                        const float* noise, vo_imu_preint* out);

/*
 * Inertial-only optimisation of the VIO initialisation: Optimizer::InertialOptimization (src/Optimizer.cc:2441-2620) with
 * EdgeInertialGS (src/G2oTypes.cc:357-482).  Poses are fixed; unknowns: one velocity per frame, gyro / acc bias, gravity
 * direction Rwg (2 dof), scale.  preint[i] is the preintegration from frame i to frame i+1 (frame i+1's mpImuPreintegrated),
 * bias_lin[i] the bias it was integrated with.  Rwb / twb: body poses (Frame::GetImuRotation / GetImuPosition, float32).
 */
typedef struct vo_inertial_problem {
  int32_t n_frames, its;       /* its = 200 (:2444) */
  const float* Rwb;            /* [n][9] */
  const float* twb;            /* [n][3] */
  float* velocity;             /* [n][3] in/out (Frame::mVw) */
  const vo_imu_preint* preint; /* [n-1] */
  const float* bias_lin;       /* [n-1][6] bax,bay,baz,bwx,bwy,bwz */
  double Rwg[9];               /* in/out */
  double scale;                /* in/out */
  double bg[3], ba[3];         /* in/out */
  float prior_g, prior_a;      /* 1e2, 1e9 (src/Tracking.cc:1453) */
  int32_t mode;                /* 0: the initialisation above; 1: Optimizer::InertialOptimization(Map*, Rwg, scale) of
                                  Tracking::ScaleRefinement (src/Optimizer.cc:2336-2439): velocities and biases fixed, no priors,
                                  only gravity direction and scale, 10 iterations, default lambda */
} vo_inertial_problem;
void vo_inertial_default_params(vo_inertial_problem* p);
int vo_inertial_optimization(vo_inertial_problem* p, vo_lm_stats* stats);
/* IMU::Preintegrated::GetUpdatedDeltaRotation / Velocity / Position (src/ImuTypes.cc:370-386) for db = (dbg, dba) */
void vo_imu_updated_deltas(const vo_imu_preint* p, const float* dbg, const float* dba, float* dR9, float* dV3, float* dP3);
void vo_imu_exp_so3_f(const float* w3, float* R9); /* IMU::ExpSO3(float) (src/ImuTypes.cc:38-50) */
/* Map::ApplyScaledRotation(R, s, bScaledVel = true, t = 0) (src/Map.cc:55-119) on the tracker's Map and last frame */
int vo_tracker_apply_scaled_rotation(void* h, const float* R9, float s);
/* ---- VIO mode of the tracker (sensor = IMU_RGBD): Tracking::ParseIMUParamFile / GrabImuData / PreintegrateIMU / InitializeIMU /
 * ScaleRefinement / UpdateFrameIMU (src/Tracking.cc:174-281, 784-1077, 1452-1480).  set_imu: Tbc row-major 4x4, noise = (ng, na,
 * ngw, naw) as given to IMU::Calib::Set.  grab_imu before each track call = System::TrackRGBD's vImuMeas; set_timestamp = its
 * timestamp argument. */
typedef struct vo_imu_state {
  int32_t initialized;     /* Tracking::mbImuInitialized */
  int32_t status;          /* last InitializeIMU attempt: -1 none, 0 done, 1 too few frames / too little time, 2 scale < 0.1,
                              3 a frame without preintegration */
  int32_t init_frame;      /* frame id at which the initialisation succeeded */
  int32_t n_refinements;   /* ScaleRefinement calls */
  int32_t n_reintegrated;  /* IMU::Preintegrated::Reintegrate calls (gyro bias moved by more than 0.01) */
  int32_t lm_iterations, lm_trials; /* of the initialisation's inertial-only optimisation */
  float t_init;            /* mTinit */
  double scale;            /* mScale of the last inertial-only optimisation */
  double Rwg[9], bg[3], ba[3];
} vo_imu_state;
/* Tracking::GetMetricError (src/Tracking.cc:3531-3674, bRMSError = false) on the tracker's Map; see vido_metric_error */
typedef struct vo_metric { float cam_t, cam_r, obj_t, obj_r; int32_t n_cam, n_obj; } vo_metric;
int vo_tracker_metric_error(void* h, const float* cam_pose_gt, int n_gt, int refined, const float* obj_pose_pre,
                            const float* obj_motion_gt, int n_obj, vo_metric* out, float* per_item);
int vo_tracker_set_imu(void* h, const float* Tbc16, const float* noise4);
int vo_tracker_grab_imu(void* h, const vo_imu_sample* samples, int n);
void vo_tracker_set_timestamp(void* h, double t);
int vo_tracker_get_imu_state(void* h, vo_imu_state* out);
/* per frame id (frame 0 included): Frame::mTcw, mVw, mImuBias (bax..bwz); returns the number of frames */
int vo_tracker_get_imu_frames(void* h, float* Tcw, float* vel, float* bias, int cap);
/* information matrix of an EdgeInertialGS from the 15x15 float32 preintegration covariance (src/G2oTypes.cc:363-375) */
void vo_inertial_edge_information(const float* C15, double* info81);

#ifdef __cplusplus
}
#endif
#endif
