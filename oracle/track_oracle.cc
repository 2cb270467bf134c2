// track_oracle.cc -- CPU restatement of the per-frame driver (VO, static + dynamic objects).  TEST INFRASTRUCTURE ONLY.
//
// Follows, for sensor = RGBD, bJoint = true, UseSampleFeature = 0:
//   Tracking::GrabImageRGBD          src/Tracking.cc:283-456   (depth pre-scale, Frame, carry-over of correspondences)
//   Frame::Frame                     src/Frame.cc:36-241       (ORB, static association)
//   Tracking::Track                  src/Tracking.cc:1081-1509 (init model, pose optimisation, motion model, renewal,
//                                                                map bookkeeping, PartialBatchOptimization every frame)
//   Tracking::Initialization         src/Tracking.cc:1512-1580
//   Tracking::RenewFrameInfo         src/Tracking.cc:2959-3135 (static part)
//   Tracking::GetStaticTrack         src/Tracking.cc:2514-2613 (rebuilt from frame 0 every frame, like the reference)
//   Optimizer::PartialBatchOptimization graph construction   src/Optimizer.cc:43-362, write-back :1056-1142
//   Tracking::UpdateMask             src/Tracking.cc:3291-3357 ; object carry-over src/Tracking.cc:391-421
//   Tracking::GetSceneFlowObj        src/Tracking.cc:1582-1668 ; Tracking::DynObjTracking src/Tracking.cc:1670-1912
//   Tracking::GetInitModelObj        src/Tracking.cc:2030-2162 ; Optimizer::PoseOptimizationFlow2 src/Optimizer.cc:3037-3253
//   Tracking::RenewFrameInfo         src/Tracking.cc:3112-3289 (object part) ; GetDynamicTrackNew src/Tracking.cc:2615-2720
// VIO glue (sensor = IMU_RGBD): Tracking::ParseIMUParamFile / GrabImuData / PreintegrateIMU / UpdateFrameIMU / InitializeIMU /
//   ScaleRefinement  src/Tracking.cc:174-281, 784-1077, 1115-1119, 1452-1480, 1555-1561 ; Frame IMU state src/Frame.cc:437-521,
//   include/Frame.h:44-110 ; Map::ApplyScaledRotation src/Map.cc:55-119 ; the write-back of Optimizer::InertialOptimization
//   src/Optimizer.cc:2589-2619
// float 4x4 products follow cv::Mat CV_32F gemm (double accumulation, one rounding).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "vido_oracle.h"

namespace {

struct P2 { float x, y; };
struct P3 { float x, y, z; };

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

void inv44(const float* T, float* Ti) {  // Converter::toInvMatrix (src/Converter.cc:155-170)
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


// ---- float32 cv::Mat helpers of the VIO glue (gemm: double accumulation, one rounding per expression)
void mm3(const float* A, const float* B, float* C) {
  float o[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = (float)((double)A[3 * r] * B[c] + (double)A[3 * r + 1] * B[3 + c] + (double)A[3 * r + 2] * B[6 + c]);
  memcpy(C, o, sizeof o);
}
// alpha * A * v + beta * w
void mv3(const float* A, const float* v, float* o, double alpha = 1.0, const float* w = nullptr) {
  float t[3];
  for (int r = 0; r < 3; r++)
    t[r] = (float)(alpha * ((double)A[3 * r] * v[0] + (double)A[3 * r + 1] * v[1] + (double)A[3 * r + 2] * v[2]) + (w ? (double)w[r] : 0.0));
  memcpy(o, t, sizeof t);
}

struct ImuFrame {              // IMU members of Frame (include/Frame.h): mTcw, mVw, mImuBias, mpImuPreintegrated, mTimeStamp
  float Tcw[16];
  float vel[3] = {0.f, 0.f, 0.f};
  float bias[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // bax, bay, baz, bwx, bwy, bwz
  bool has_pre = false;
  vo_imu_preint pre;           // integrated with pre_b; db = bu - b: gyro (0..2), acc (3..5) (IMU::Preintegrated::db)
  float pre_b[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, db[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  size_t q0 = 0, q1 = 0;       // queue range the preintegration saw (Reintegrate)
  double t = 0, t_prev = 0;
};

struct Frame {
  std::vector<vo_keypoint> mvKeys;
  std::vector<P2> mvStatKeysTmp, mvCorres, mvFlowNext, mvStatKeys;
  std::vector<float> mvStatDepthTmp, mvStatDepth;
  std::vector<P3> mvStat3DPointTmp;
  std::vector<int> nStaInlierID;
  float Tcw[16];
  // object features (Frame.h: mvObjKeys, mvObjDepth, mvObjCorres, mvObjFlowNext, vSemObjLabel, vObjLabel, nDynInlierID,
  // mvObj3DPoint) and per-object results (nModLabel, nSemPosition, bObjStat, vObjMod, vObjCentre3D, vnObjID, vnObjInlierID)
  std::vector<P2> mvObjKeys, mvObjCorres, mvObjFlowNext;
  std::vector<float> mvObjDepth;
  std::vector<int> vSemObjLabel, vObjLabel, nDynInlierID;
  std::vector<P3> mvObj3DPoint, vFlow_3d;
  std::vector<int> nModLabel, nSemPosition;
  std::vector<char> bObjStat;
  std::vector<std::array<float, 16>> vObjMod;
  std::vector<P3> vObjCentre3D;
  std::vector<std::vector<int>> vnObjID, vnObjInlierID;
};

struct Map {
  std::vector<std::vector<P2>> vpFeatSta;
  std::vector<std::vector<float>> vfDepSta;
  std::vector<std::vector<P3>> vp3DPointSta;
  std::vector<std::vector<int>> vnAssoSta;
  std::vector<std::vector<std::pair<int, int>>> TrackletSta;
  std::vector<std::vector<float>> vmCameraPose;   // 16 floats each (Twc)
  std::vector<std::vector<float>> vmRigidMotion;  // camera motion [f-1][0]
  // dynamic part
  std::vector<std::vector<P2>> vpFeatDyn;
  std::vector<std::vector<float>> vfDepDyn;
  std::vector<std::vector<P3>> vp3DPointDyn;
  std::vector<std::vector<int>> vnAssoDyn, vnFeatLabel;
  std::vector<std::vector<std::pair<int, int>>> TrackletDyn;
  std::vector<int> nObjID;
  std::vector<std::vector<std::array<float, 16>>> vmObjMotion;  // vmRigidMotion[f-1][1..]
  std::vector<std::vector<int>> vnRMLabel, vnSMLabel;            // without the camera entry
  std::vector<std::vector<P3>> vmRigidCentre;
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// inputs and outputs of one Tracking::DynObjTracking call, kept when the test-suite asks for them (vo_tracker_dyn_log_*): an
// independent restatement of src/Tracking.cc:1670-1912 in tests/test_dynobj_independent.py is run on the inputs
struct DynRecord {
  int f_id = 0, max_id_before = 0, max_id_after = 0;
  std::vector<int> sem, lab_before, lab_after, last_sem, last_sem_pos, last_stat, last_mod, out_mod, out_sem_pos, out_len, out_ids;
  std::vector<float> key_xy, depth, flow3;
};

// inputs and outputs of the gravity-direction / velocity initialisation of Tracking::InitializeIMU (src/Tracking.cc:955-988), kept for
// tests/test_vio_oracle.py when the test-suite asks for them (same switch as the DynObjTracking recorder)
struct ImuInitRecord {
  std::vector<float> Tcw, dV, dT, vel_out;   // per Map frame: 16 / 3 (updated delta velocity) / 1 / 3 floats
  std::vector<int> has_pre;
  float Tbc[16], Rwg[9];
};

// inputs and outputs of Tracking::UpdateFrameIMU (src/Tracking.cc:889-925) for tests/test_vio_oracle.py (same switch)
struct ImuUpdateRecord { float Rwb1[9], twb1[3], Vwb1[3], dR[9], dV[3], dP[3], t12, Rwb[9], twb[3], Vwb[3]; };
static_assert(sizeof(ImuUpdateRecord) == 46 * sizeof(float), "record layout");

struct Tracker {
  std::vector<ImuUpdateRecord> imu_update_log;
  bool dyn_log_on = false;
  std::vector<DynRecord> dyn_log;
  std::vector<ImuInitRecord> imu_init_log;
  vo_track_config cfg;
  Map map;
  Frame* last = nullptr;
  Frame* cur = nullptr;
  bool has_velocity = false;
  float mVelocity[16];
  int f_id = 0;
  bool initialised = false;
  // incremental tracklet state (rebuild_tracklets == 0)
  std::vector<int> trackOfPrev, dynTrackOfPrev;
  int max_id = 1;
  // mSegMapLast / mFlowMapLast (src/Tracking.cc:777-780), kept only while the last frame carries object features
  std::vector<int32_t> segLast, segCur;
  std::vector<float> flowLast;
  // object samples of the current Frame ctor (mvTmpObjKeys ... of src/Tracking.cc:395-399)
  std::vector<P2> tmpKeys, tmpCorres, tmpFlow;
  std::vector<float> tmpDepth;
  std::vector<int> tmpSem;


  // ---- VIO state (Tracking: mpImuCalib, mlQueueImuData, mbImuInitialized, mScale, mRwg, mbg, mba, mTinit, mFirstTs)
  bool vio = false;
  float Tbc[16], Tcb[16], imu_noise[4];
  std::vector<vo_imu_sample> imu_q;
  size_t imu_head = 0;                 // queue front
  std::vector<ImuFrame> fr;            // by frame id; fr[0] (the initial frame) is not in Map::vpFrames
  double next_t = 0;
  vo_imu_state ist;
  vo_lm_stats imu_lm;

  void set_imu(const float* Tbc_, const float* noise) {   // IMU::Calib::Set (src/ImuTypes.cc:476-500)
    vio = true;
    memcpy(Tbc, Tbc_, sizeof Tbc);
    memcpy(imu_noise, noise, sizeof imu_noise);
    eye44(Tcb);
    float Rt[9], tb[3] = {Tbc[3], Tbc[7], Tbc[11]}, tc[3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) { Rt[3 * r + c] = Tbc[4 * c + r]; Tcb[4 * r + c] = Tbc[4 * c + r]; }
    mv3(Rt, tb, tc, -1.0);
    for (int r = 0; r < 3; r++) Tcb[4 * r + 3] = tc[r];
    memset(&ist, 0, sizeof ist);
    ist.scale = 1.0; ist.status = -1;
    ist.Rwg[0] = ist.Rwg[4] = ist.Rwg[8] = 1.0;
  }
  void imu_rotation(const ImuFrame& f, float* Rwb) const {   // Frame::GetImuRotation: mRwc * Tcb.R
    float Rwc[9], Rcb[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) { Rwc[3 * r + c] = f.Tcw[4 * c + r]; Rcb[3 * r + c] = Tcb[4 * r + c]; }
    mm3(Rwc, Rcb, Rwb);
  }
  void imu_position(const ImuFrame& f, float* twb) const {   // Frame::GetImuPosition: mOwb = mRwc * tcb + mOw, mOw = -mRcw.t() * mtcw
    float Rwc[9], tcw[3] = {f.Tcw[3], f.Tcw[7], f.Tcw[11]}, tcb[3] = {Tcb[3], Tcb[7], Tcb[11]}, Ow[3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Rwc[3 * r + c] = f.Tcw[4 * c + r];
    mv3(Rwc, tcw, Ow, -1.0);
    mv3(Rwc, tcb, twb, 1.0, Ow);
  }
  void set_imu_pose_velocity(ImuFrame& f, const float* Rwb, const float* twb, const float* Vwb) const {   // src/Frame.cc:510-521
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
    mul44(Tcb, Tbw, f.Tcw);
  }
  static void set_new_bias(ImuFrame& f, const float* b) {   // Frame::SetNewBias + IMU::Preintegrated::SetNewBias (src/ImuTypes.cc:328-339)
    memcpy(f.bias, b, sizeof f.bias);
    if (f.has_pre)
      for (int k = 0; k < 3; k++) { f.db[k] = b[3 + k] - f.pre_b[3 + k]; f.db[3 + k] = b[k] - f.pre_b[k]; }
  }
  static void pre_set_new_bias(ImuFrame& f, const float* b) {   // mpImuPreintegrated->SetNewBias only
    if (f.has_pre)
      for (int k = 0; k < 3; k++) { f.db[k] = b[3 + k] - f.pre_b[3 + k]; f.db[3 + k] = b[k] - f.pre_b[k]; }
  }
  void integrate(ImuFrame& f, const float* b) {   // new IMU::Preintegrated(b, calib) over the frame's queue range
    vo_imu_preintegrate(imu_q.data() + f.q0, (int)(f.q1 - f.q0), f.t_prev, f.t, b, imu_noise, &f.pre);
    memcpy(f.pre_b, b, sizeof f.pre_b);
    for (int k = 0; k < 6; k++) f.db[k] = 0.f;
    f.has_pre = true;
  }
  // Tracking::PreintegrateIMU (src/Tracking.cc:784-887) for the new frame, bias of the last frame
  void preintegrate(ImuFrame& f, const ImuFrame& prev) {
    f.has_pre = false;
    if (imu_head >= imu_q.size()) return;   // "Not IMU data in mlQueueImuData"
    f.q0 = imu_head; f.q1 = imu_q.size();
    integrate(f, prev.bias);
    imu_head += (size_t)f.pre.n_consumed;
  }
  // Tracking::UpdateFrameIMU (src/Tracking.cc:889-923); mpLastFrame == mpCurrentFrame == fr.back()
  void update_frame_imu(const float* b) {
    ImuFrame& c = fr.back();
    const ImuFrame& p = fr[fr.size() - 2];
    set_new_bias(c, b);
    if (!c.has_pre) return;
    const float Gz[3] = {0.f, 0.f, -9.79f};
    float twb1[3], Rwb1[9], dR[9], dV[3], dP[3], Rwb[9], twb[3], Vwb[3];
    imu_position(p, twb1);
    imu_rotation(p, Rwb1);
    vo_imu_updated_deltas(&c.pre, c.db, c.db + 3, dR, dV, dP);
    const float t12 = c.pre.dT;
    mm3(Rwb1, dR, Rwb);
    const float ht2 = 0.5f * t12 * t12;
    for (int r = 0; r < 3; r++) {
      // twb1 + Vwb1*t12 + 0.5f*t12*t12*Gz + Rwb1*dP ; Vwb1 + Gz*t12 + Rwb1*dV  (left to right, one rounding per operator)
      const float rp = (float)((double)Rwb1[3 * r] * dP[0] + (double)Rwb1[3 * r + 1] * dP[1] + (double)Rwb1[3 * r + 2] * dP[2]);
      const float rv = (float)((double)Rwb1[3 * r] * dV[0] + (double)Rwb1[3 * r + 1] * dV[1] + (double)Rwb1[3 * r + 2] * dV[2]);
      twb[r] = ((twb1[r] + (float)((double)p.vel[r] * (double)t12)) + (float)((double)ht2 * (double)Gz[r])) + rp;
      Vwb[r] = (p.vel[r] + (float)((double)Gz[r] * (double)t12)) + rv;
    }
    if (dyn_log_on) {
      ImuUpdateRecord u;
      memcpy(u.Rwb1, Rwb1, sizeof u.Rwb1); memcpy(u.twb1, twb1, sizeof u.twb1); memcpy(u.Vwb1, p.vel, sizeof u.Vwb1);
      memcpy(u.dR, dR, sizeof u.dR); memcpy(u.dV, dV, sizeof u.dV); memcpy(u.dP, dP, sizeof u.dP); u.t12 = t12;
      memcpy(u.Rwb, Rwb, sizeof u.Rwb); memcpy(u.twb, twb, sizeof u.twb); memcpy(u.Vwb, Vwb, sizeof u.Vwb);
      imu_update_log.push_back(u);
    }
    set_imu_pose_velocity(c, Rwb, twb, Vwb);
    memcpy(cur->Tcw, c.Tcw, sizeof c.Tcw);   // mpLastFrame = mpCurrentFrame
  }
  // Map::ApplyScaledRotation (src/Map.cc:55-119) with the frames of Map::vpFrames (fr[1..])
  void apply_scaled_rotation(const float* R, float s);
  // the inertial problem over Map::vpFrames = fr[1..N] (src/Optimizer.cc:2441-2560 / 2336-2425); false: a frame without preintegration
  bool run_inertial(int mode, float priorG, float priorA) {
    const int N = (int)fr.size() - 1;
    std::vector<float> Rwb(9 * (size_t)N), twb(3 * (size_t)N), vel(3 * (size_t)N), blin(6 * (size_t)(N - 1));
    std::vector<vo_imu_preint> pre(N - 1);
    for (int i = 1; i <= N; i++) {
      imu_rotation(fr[i], &Rwb[9 * (size_t)(i - 1)]);
      imu_position(fr[i], &twb[3 * (size_t)(i - 1)]);
      memcpy(&vel[3 * (size_t)(i - 1)], fr[i].vel, sizeof(float) * 3);
      pre_set_new_bias(fr[i], fr[i - 1].bias);
      if (i >= 2) {
        if (!fr[i].has_pre) return false;
        pre[i - 2] = fr[i].pre;
        memcpy(&blin[6 * (size_t)(i - 2)], fr[i].pre_b, sizeof(float) * 6);
      }
    }
    vo_inertial_problem p;
    memset(&p, 0, sizeof p);
    vo_inertial_default_params(&p);
    p.n_frames = N; p.Rwb = Rwb.data(); p.twb = twb.data(); p.velocity = vel.data(); p.preint = pre.data(); p.bias_lin = blin.data();
    p.mode = mode; p.its = mode == 0 ? 200 : 10;
    p.prior_g = priorG; p.prior_a = priorA;
    for (int k = 0; k < 9; k++) p.Rwg[k] = ist.Rwg[k];
    p.scale = ist.scale;
    for (int k = 0; k < 3; k++) { p.bg[k] = (double)fr[1].bias[3 + k]; p.ba[k] = (double)fr[1].bias[k]; }   // VertexGyroBias(vpFs.front())
    vo_inertial_optimization(&p, &imu_lm);
    for (int k = 0; k < 9; k++) ist.Rwg[k] = p.Rwg[k];
    ist.scale = p.scale;
    if (mode == 0) {   // write-back, src/Optimizer.cc:2589-2619
      for (int k = 0; k < 3; k++) { ist.bg[k] = p.bg[k]; ist.ba[k] = p.ba[k]; }
      const float b[6] = {(float)p.ba[0], (float)p.ba[1], (float)p.ba[2], (float)p.bg[0], (float)p.bg[1], (float)p.bg[2]};
      for (int i = 1; i <= N; i++) {
        memcpy(fr[i].vel, &vel[3 * (size_t)(i - 1)], sizeof(float) * 3);
        double d2 = 0;
        for (int k = 0; k < 3; k++) { const float d = fr[i].bias[3 + k] - b[3 + k]; d2 += (double)d * d; }
        const bool re = std::sqrt(d2) > 0.01;
        set_new_bias(fr[i], b);
        if (re && fr[i].has_pre) { integrate(fr[i], b); ist.n_reintegrated++; }
      }
    }
    return true;
  }
  void apply_and_update(const float* b) {   // tail shared by InitializeIMU / ScaleRefinement
    if (std::fabs(ist.scale - 1.0) > 0.00001) {
      float Rgw[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Rgw[3 * r + c] = (float)ist.Rwg[3 * c + r];   // Converter::toCvMat(mRwg).t()
      apply_scaled_rotation(Rgw, (float)ist.scale);
      update_frame_imu(b);
    }
  }
  // Tracking::InitializeIMU(1e2, 1e9) (src/Tracking.cc:937-1044)
  void initialize_imu() {
    const int N = (int)fr.size() - 1;
    if (N < 10) { ist.status = 1; return; }
    const double first_ts = fr[1].t;
    if (fr.back().t - first_ts < 2.0) { ist.status = 1; return; }
    float dirG[3] = {0.f, 0.f, 0.f};
    ImuInitRecord irec;
    if (dyn_log_on) {
      memcpy(irec.Tbc, Tbc, sizeof irec.Tbc);
      for (int i = 0; i <= N; i++) {
        irec.Tcw.insert(irec.Tcw.end(), fr[i].Tcw, fr[i].Tcw + 16);
        irec.has_pre.push_back(fr[i].has_pre ? 1 : 0);
        float dR[9], dV[3] = {0.f, 0.f, 0.f}, dP[3];
        if (fr[i].has_pre) vo_imu_updated_deltas(&fr[i].pre, fr[i].db, fr[i].db + 3, dR, dV, dP);
        irec.dV.insert(irec.dV.end(), dV, dV + 3);
        irec.dT.push_back(fr[i].has_pre ? fr[i].pre.dT : 0.f);
      }
    }
    for (int i = 1; i <= N; i++) {
      if (!fr[i].has_pre) continue;
      float R[9], dR[9], dV[3], dP[3], p1[3], p0[3];
      imu_rotation(fr[i - 1], R);
      vo_imu_updated_deltas(&fr[i].pre, fr[i].db, fr[i].db + 3, dR, dV, dP);
      imu_position(fr[i], p1);
      imu_position(fr[i - 1], p0);
      for (int r = 0; r < 3; r++) {
        const float rv = (float)((double)R[3 * r] * dV[0] + (double)R[3 * r + 1] * dV[1] + (double)R[3 * r + 2] * dV[2]);
        dirG[r] = dirG[r] - rv;
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
    vo_imu_exp_so3_f(vzg, Rwg);
    for (int k = 0; k < 9; k++) ist.Rwg[k] = Rwg[k];
    if (dyn_log_on) {
      memcpy(irec.Rwg, Rwg, sizeof irec.Rwg);
      for (int i = 0; i <= N; i++) irec.vel_out.insert(irec.vel_out.end(), fr[i].vel, fr[i].vel + 3);
      imu_init_log.push_back(irec);
    }
    ist.t_init = (float)(fr.back().t - first_ts);
    ist.scale = 1.0;
    if (!run_inertial(0, 1e2f, 1e9f)) { ist.status = 3; return; }
    ist.lm_iterations = imu_lm.iterations; ist.lm_trials = imu_lm.total_trials;
    if (ist.scale < 1e-1) { ist.status = 2; return; }
    float b[6];
    memcpy(b, fr[1].bias, sizeof b);   // vpF[0]->GetImuBias()
    apply_and_update(b);
    ist.initialized = 1; ist.init_frame = N; ist.status = 0;
  }
  // Tracking::ScaleRefinement (src/Tracking.cc:1046-1077)
  void scale_refinement() {
    for (int k = 0; k < 9; k++) ist.Rwg[k] = (k % 4 == 0) ? 1.0 : 0.0;
    ist.scale = 1.0;
    if (!run_inertial(1, 0.f, 0.f)) return;
    ist.n_refinements++;
    if (ist.scale < 1e-1) return;
    float b[6];
    memcpy(b, fr.back().bias, sizeof b);
    apply_and_update(b);
  }
  // the IMU part of Tracking::Track after the window optimisation (src/Tracking.cc:1452-1480)
  void vio_after_ba() {
    if (!ist.initialized) initialize_imu();
    if (ist.initialized && ist.t_init < 100.0f) {
      ist.t_init = (float)((double)ist.t_init + (fr.back().t - fr[fr.size() - 2].t));
      const float T = ist.t_init;
      const bool win = (T > 15.0f && T < 15.5f) || (T > 25.0f && T < 25.5f) || (T > 35.0f && T < 35.5f) || (T > 45.0f && T < 45.5f) ||
                       (T > 55.0f && T < 55.5f) || (T > 65.0f && T < 65.5f) || (T > 75.0f && T < 75.5f);
      if ((int)fr.size() - 1 <= 1000 && win) scale_refinement();
    }
  }

  ~Tracker() { delete last; if (cur != last) delete cur; }

  P3 unproject_cam(const P2& kp, float z) const {  // Optimizer::Get3DinCamera
    const float invfx = 1.0f / cfg.fx, invfy = 1.0f / cfg.fy;
    return {(kp.x - cfg.cx) * z * invfx, (kp.y - cfg.cy) * z * invfy, z};
  }
  P3 to_world(const P3& xc, const float* Twc) const {  // cv::Mat  R*x + t
    P3 o;
    float* po = &o.x;
    for (int r = 0; r < 3; r++)
      po[r] = (float)((double)Twc[4 * r] * xc.x + (double)Twc[4 * r + 1] * xc.y + (double)Twc[4 * r + 2] * xc.z) + Twc[4 * r + 3];
    return o;
  }

  // ---- Tracking::GetStaticTrack
  void rebuild_tracklets() {
    auto& TM = map.vnAssoSta;
    const int N = (int)TM.size();
    std::vector<int> pre;
    std::vector<std::vector<std::pair<int, int>>> T;
    for (int i = 0; i < N; i++) {
      std::vector<int> curc(TM[i].size(), -1);
      for (size_t j = 0; j < TM[i].size(); j++) {
        if (TM[i][j] == -1) continue;
        if (i > 0 && pre[TM[i][j]] != -1) {
          T[pre[TM[i][j]]].push_back({i + 1, (int)j});
          curc[j] = pre[TM[i][j]];
        } else {
          T.push_back({{i, TM[i][j]}, {i + 1, (int)j}});
          curc[j] = (int)T.size() - 1;
        }
      }
      pre = curc;
    }
    map.TrackletSta.swap(T);
  }
  void extend_tracklets() {  // same result, O(features) per frame ("fixed bookkeeping" baseline)
    auto& TM = map.vnAssoSta;
    const int i = (int)TM.size() - 1;
    std::vector<int> curc(TM[i].size(), -1);
    for (size_t j = 0; j < TM[i].size(); j++) {
      if (TM[i][j] == -1) continue;
      if (i > 0 && trackOfPrev[TM[i][j]] != -1) {
        map.TrackletSta[trackOfPrev[TM[i][j]]].push_back({i + 1, (int)j});
        curc[j] = trackOfPrev[TM[i][j]];
      } else {
        map.TrackletSta.push_back({{i, TM[i][j]}, {i + 1, (int)j}});
        curc[j] = (int)map.TrackletSta.size() - 1;
      }
    }
    trackOfPrev = curc;
  }


  // ---- Frame::UnprojectStereoObject(i, 0) (src/Frame.cc:735-769): world point of object feature i of frame f
  P3 unproject_obj(const Frame& f, int i) const {
    float Twf[16];
    inv44(f.Tcw, Twf);
    return to_world(unproject_cam(f.mvObjKeys[i], f.mvObjDepth[i]), Twf);
  }

  // ---- Tracking::GetSceneFlowObj (src/Tracking.cc:1582-1668)
  void scene_flow() {
    const int N = (int)cur->mvObjKeys.size();
    cur->vFlow_3d.assign(N, P3{0.f, 0.f, 0.f});
    float Twl[16], Twc[16];
    inv44(last->Tcw, Twl);
    inv44(cur->Tcw, Twc);
    for (int i = 0; i < N; i++) {
      if (cur->vSemObjLabel[i] <= 0 || last->vSemObjLabel[i] <= 0) { cur->vObjLabel[i] = -1; continue; }
      const P3 xp = to_world(unproject_cam(last->mvObjKeys[i], last->mvObjDepth[i]), Twl);
      const P3 xc = to_world(unproject_cam(cur->mvObjKeys[i], cur->mvObjDepth[i]), Twc);
      cur->vFlow_3d[i] = {xc.x - xp.x, xc.y - xp.y, xc.z - xp.z};
    }
  }

  // most frequent value; ties -> smallest value (std::map order + insertion sort of SortPairInt, src/Tracking.cc:1853-1863)
  static int majority(const std::vector<int>& v) {
    std::map<int, int> dups;
    for (int k : v) ++dups[k];
    int best = 0, cnt = -1;
    for (auto& k : dups)
      if (k.second > cnt) { cnt = k.second; best = k.first; }
    return best;
  }

  // ---- Tracking::DynObjTracking (src/Tracking.cc:1670-1912)
  std::vector<std::vector<int>> dyn_obj_tracking() {
    const int W = cfg.width, H = cfg.height;
    std::vector<int> UniLab = cur->vSemObjLabel;
    std::sort(UniLab.begin(), UniLab.end());
    UniLab.erase(std::unique(UniLab.begin(), UniLab.end()), UniLab.end());
    std::vector<std::vector<int>> Posi(UniLab.size());
    for (size_t i = 0; i < cur->vSemObjLabel.size(); i++) {
      if (cur->vObjLabel[i] == -1) continue;
      for (size_t j = 0; j < UniLab.size(); j++)
        if (cur->vSemObjLabel[i] == UniLab[j]) { Posi[j].push_back((int)i); break; }
    }
    // objects mostly on the image boundary are dropped
    std::vector<std::vector<int>> ObjId;
    std::vector<int> sem_posi;
    const int shrin_thr_row = 10, shrin_thr_col = 20;
    for (size_t i = 0; i < Posi.size(); i++) {
      float count = 0;
      const float count_thres = 0.5f;
      for (int id : Posi[i]) {
        const float u = cur->mvObjKeys[id].x, v = cur->mvObjKeys[id].y;
        if (v < shrin_thr_row || v > (H - shrin_thr_row) || u < shrin_thr_col || u > (W - shrin_thr_col)) count = count + 1;
      }
      if (count / Posi[i].size() > count_thres) {
        for (int id : Posi[i]) cur->vObjLabel[id] = -1;
        continue;
      }
      ObjId.push_back(Posi[i]);
      sem_posi.push_back(UniLab[i]);
    }
    // static / far / small objects
    std::vector<std::vector<int>> ObjIdNew;
    std::vector<int> SemPosNew;
    for (size_t i = 0; i < ObjId.size(); i++) {
      float obj_center_depth = 0, sf_count = 0;
      for (int id : ObjId[i]) {
        obj_center_depth = obj_center_depth + cur->mvObjDepth[id];
        const P3& f = cur->vFlow_3d[id];
        const float sf_norm = std::sqrt(f.x * f.x + f.z * f.z);
        if (sf_norm < cfg.sf_mg_thres) sf_count = sf_count + 1;
      }
      if (sf_count / ObjId[i].size() > cfg.sf_ds_thres) {
        for (int id : ObjId[i]) cur->vObjLabel[id] = 0;
        continue;
      } else if (obj_center_depth / ObjId[i].size() > cfg.th_depth_obj || ObjId[i].size() < 150) {
        for (int id : ObjId[i]) cur->vObjLabel[id] = -1;
        continue;
      }
      ObjIdNew.push_back(ObjId[i]);
      SemPosNew.push_back(sem_posi[i]);
    }
    // tracking ids: same semantic label as an object with a valid motion in the last frame -> same id, else a new one
    if (f_id == 1) max_id = 1;
    std::vector<int> LabId(ObjIdNew.size());
    for (size_t i = 0; i < ObjIdNew.size(); i++) {
      std::vector<int> Lb_last;
      for (int id : ObjIdNew[i]) Lb_last.push_back(last->vSemObjLabel[id]);
      const int New_lab = majority(Lb_last);
      bool exist = false;
      if (max_id != 1) {
        for (size_t k = 0; k < last->nSemPosition.size(); k++)
          if (last->nSemPosition[k] == New_lab && last->bObjStat[k]) { LabId[i] = last->nModLabel[k]; exist = true; break; }
      }
      if (!exist) { LabId[i] = max_id; max_id = max_id + 1; }
      for (int id : ObjIdNew[i]) cur->vObjLabel[id] = LabId[i];
    }
    cur->nModLabel = LabId;
    cur->nSemPosition = SemPosNew;
    return ObjIdNew;
  }

  // ---- object loop of Tracking::Track (src/Tracking.cc:1179-1308): GetInitModelObj + PoseOptimizationFlow2 per object
  void object_motions(const std::vector<std::vector<int>>& ObjIdNew) {
    const size_t no = ObjIdNew.size();
    cur->bObjStat.assign(no, 1);
    cur->vObjMod.resize(no);
    cur->vObjCentre3D.assign(no, P3{0.f, 0.f, 0.f});
    cur->vnObjID.resize(no);
    cur->vnObjInlierID.resize(no);
    float Twl[16], Twc[16];
    inv44(last->Tcw, Twl);
    inv44(cur->Tcw, Twc);
    for (size_t i = 0; i < no; i++) {
      const std::vector<int>& ObjId = ObjIdNew[i];
      const int N = (int)ObjId.size();
      cur->vnObjID[i] = ObjId;
      // centroid of the object's points in the last frame (cv::Mat float sums)
      std::vector<float> cur2d(2 * (size_t)N), p3d(3 * (size_t)N);
      P3 c = {0.f, 0.f, 0.f};
      for (int j = 0; j < N; j++) {
        const P3 xp = to_world(unproject_cam(last->mvObjKeys[ObjId[j]], last->mvObjDepth[ObjId[j]]), Twl);
        c.x += xp.x; c.y += xp.y; c.z += xp.z;
        p3d[3 * j] = xp.x; p3d[3 * j + 1] = xp.y; p3d[3 * j + 2] = xp.z;
        cur2d[2 * j] = cur->mvObjKeys[ObjId[j]].x; cur2d[2 * j + 1] = cur->mvObjKeys[ObjId[j]].y;
      }
      const float invn = (float)(1.0 / (double)N);
      cur->vObjCentre3D[i] = {c.x * invn, c.y * invn, c.z * invn};
      // ---- GetInitModelObj
      int PreObjID = -1;
      for (size_t k = 0; k < last->nModLabel.size(); k++)
        if (last->nModLabel[k] == cur->nModLabel[i]) { PreObjID = (int)k; break; }
      vo_pnp_problem pp;
      memset(&pp, 0, sizeof pp);
      vo_pnp_default_params(&pp);
      std::vector<int> ids(N);
      pp.n = N; pp.cur_xy = cur2d.data(); pp.pts3d = p3d.data(); pp.valid = nullptr; pp.inlier_ids = ids.data();
      if (PreObjID != -1) mul44(cur->Tcw, last->vObjMod[PreObjID].data(), pp.Tcw_motion);
      else { memcpy(pp.Tcw_motion, cur->Tcw, sizeof(float) * 16); pp.no_motion_model = 1; }
      pp.fx = cfg.fx; pp.fy = cfg.fy; pp.cx = cfg.cx; pp.cy = cfg.cy;
      vo_init_model_cam(&pp);
      std::vector<int> in_ids(pp.n_inliers);
      std::vector<char> keep(N, 0);
      for (int k = 0; k < pp.n_inliers; k++) { in_ids[k] = ObjId[ids[k]]; keep[ids[k]] = 1; }
      for (int j = 0; j < N; j++)
        if (!keep[j]) cur->vObjLabel[ObjId[j]] = -1;
      if (in_ids.size() < 50) {
        cur->bObjStat[i] = 0;
        eye44(cur->vObjMod[i].data());
        cur->vObjCentre3D[i] = {0.f, 0.f, 0.f};
        cur->vnObjInlierID[i] = in_ids;
        continue;
      }
      std::vector<int> InlierID;
      if (cfg.b_joint) {
        // ---- PoseOptimizationFlow2
        const int n = (int)in_ids.size();
        std::vector<float> obs(2 * (size_t)n), fl(2 * (size_t)n), dep(n), fo(2 * (size_t)n);
        std::vector<int> inl(n);
        for (int k = 0; k < n; k++) {
          const int id = in_ids[k];
          obs[2 * k] = last->mvObjKeys[id].x; obs[2 * k + 1] = last->mvObjKeys[id].y;
          fl[2 * k] = last->mvObjFlowNext[id].x; fl[2 * k + 1] = last->mvObjFlowNext[id].y;
          dep[k] = last->mvObjDepth[id];
        }
        vo_poseopt_problem po;
        memset(&po, 0, sizeof po);
        vo_poseopt_default_params(&po);
        po.n = n; po.obs_xy = obs.data(); po.flow_xy = fl.data(); po.depth = dep.data();
        memcpy(po.Tcw_init, pp.Tcw_out, sizeof(float) * 16);
        memcpy(po.Tcw_last, last->Tcw, sizeof(float) * 16);
        po.fx = cfg.fx; po.fy = cfg.fy; po.cx = cfg.cx; po.cy = cfg.cy;
        po.flow_out = fo.data(); po.inlier = inl.data();
        po.info_prior = 0.5f; po.rounds = 1; po.its = 200;
        vo_poseopt_flow2cam(&po, nullptr);
        mul44(Twc, po.Tcw_out, cur->vObjMod[i].data());  // vObjMod = inv(Tcw) * Obj_X
        for (int k = 0; k < n; k++) {
          const int id = in_ids[k];
          if (inl[k]) {
            cur->mvObjKeys[id].x = (float)((double)last->mvObjKeys[id].x + (double)fo[2 * k]);
            cur->mvObjKeys[id].y = (float)((double)last->mvObjKeys[id].y + (double)fo[2 * k + 1]);
            InlierID.push_back(id);
          } else cur->vObjLabel[id] = -1;
        }
      } else {
        // ---- PoseOptimizationObjMot (src/Optimizer.cc:2826-3035): world-frame motion H, projection P = K * Tcw
        const int n = (int)in_ids.size();
        std::vector<float> obs(2 * (size_t)n), p3(3 * (size_t)n);
        std::vector<int> inl(n);
        for (int k = 0; k < n; k++) {
          const int id = in_ids[k];
          const P3 xp = to_world(unproject_cam(last->mvObjKeys[id], last->mvObjDepth[id]), Twl);
          obs[2 * k] = cur->mvObjKeys[id].x; obs[2 * k + 1] = cur->mvObjKeys[id].y;
          p3[3 * k] = xp.x; p3[3 * k + 1] = xp.y; p3[3 * k + 2] = xp.z;
        }
        vo_projopt_problem pj;
        memset(&pj, 0, sizeof pj);
        vo_projopt_default_params(&pj, 1);
        pj.n = n; pj.obs_xy = obs.data(); pj.pts3d = p3.data(); pj.inlier = inl.data();
        mul44(Twc, pp.Tcw_out, pj.T_init);   // Init = inv(Tcw) * mInitModel
        const double KK[12] = {cfg.fx, 0, cfg.cx, 0, 0, cfg.fy, cfg.cy, 0, 0, 0, 1, 0};
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 4; c++) {
            double v = 0;
            for (int m = 0; m < 4; m++) v += KK[4 * r + m] * (double)cur->Tcw[4 * m + c];
            pj.P[4 * r + c] = v;
          }
        vo_pose_opt_proj(&pj, nullptr);
        memcpy(cur->vObjMod[i].data(), pj.T_out, sizeof(float) * 16);
        for (int k = 0; k < n; k++) {
          if (inl[k]) InlierID.push_back(in_ids[k]);
          else cur->vObjLabel[in_ids[k]] = -1;
        }
      }
      cur->vnObjInlierID[i] = InlierID;
    }
  }

  // ---- Tracking::RenewFrameInfo, object part (src/Tracking.cc:3112-3289)
  void renew_objects(const float* depth, const float* flow, const int32_t* mask) {
    const int W = cfg.width, H = cfg.height, max_num_obj = cfg.max_track_obj;
    std::vector<P2> keys, corres, fl;
    std::vector<float> dep;
    std::vector<int> sem, inl, lab;
    const size_t no = cur->vnObjInlierID.size();
    std::vector<int> ObjFeaCount(no);
    for (size_t i = 0; i < no; i++) {
      if (!cur->bObjStat[i]) { ObjFeaCount[i] = -1; continue; }
      int count = 0;
      for (int id : cur->vnObjInlierID[i]) {
        const int x = (int)cur->mvObjKeys[id].x, y = (int)cur->mvObjKeys[id].y;
        if (x >= W || y >= H || x <= 0 || y <= 0) continue;
        const size_t k = (size_t)y * W + x;
        if (mask[k] != 0 && depth[k] < 25 && depth[k] > 0) {
          const float fx = flow[2 * k], fy = flow[2 * k + 1];
          if (x + fx < W && y + fy < H && x + fx > 0 && y + fy > 0) {
            keys.push_back({(float)x, (float)y});
            dep.push_back(depth[k]);
            sem.push_back(mask[k]);
            fl.push_back({fx, fy});
            corres.push_back({x + fx, y + fy});
            inl.push_back(id);
            lab.push_back(cur->vObjLabel[id]);
            count = count + 1;
          }
        }
      }
      ObjFeaCount[i] = count;
    }
    // top-up per tracked object from this frame's samples, 15 interleaved passes, >= 1 px away from the kept inliers
    const std::vector<P2> check = keys;
    for (size_t i = 0; i < no; i++) {
      if (!cur->bObjStat[i]) continue;
      const int SemLabel = cur->nSemPosition[i];
      int tot_num = ObjFeaCount[i], start_id = 0;
      const int step = 15;
      while (tot_num < max_num_obj) {
        if (start_id == step) break;
        for (size_t j = start_id; j < tmpSem.size(); j += step) {
          if (tmpSem[j] != SemLabel) continue;
          float min_dist = 100;
          bool used = false;
          for (size_t k = 0; k < check.size(); k++) {
            const float d = std::sqrt((check[k].x - tmpKeys[j].x) * (check[k].x - tmpKeys[j].x) +
                                      (check[k].y - tmpKeys[j].y) * (check[k].y - tmpKeys[j].y));
            if (d < min_dist) min_dist = d;
            if (min_dist < 1.0) { used = true; break; }
          }
          if (used) continue;
          keys.push_back(tmpKeys[j]); dep.push_back(tmpDepth[j]); sem.push_back(tmpSem[j]); fl.push_back(tmpFlow[j]);
          corres.push_back(tmpCorres[j]); inl.push_back(-1); lab.push_back(cur->nModLabel[i]);
          tot_num = tot_num + 1;
          if (tot_num >= max_num_obj) break;
        }
        start_id = start_id + 1;
      }
    }
    // semantic labels without a tracked object: all their samples, label -2
    std::vector<int> UniLab = tmpSem;
    std::sort(UniLab.begin(), UniLab.end());
    UniLab.erase(std::unique(UniLab.begin(), UniLab.end()), UniLab.end());
    std::vector<char> NewLab(UniLab.size(), 0);
    for (size_t i = 0; i < cur->nSemPosition.size(); i++)
      for (size_t j = 0; j < UniLab.size(); j++)
        if (UniLab[j] == cur->nSemPosition[i] && cur->bObjStat[i]) { NewLab[j] = 1; break; }
    for (size_t i = 0; i < NewLab.size(); i++) {
      if (NewLab[i]) continue;
      for (size_t j = 0; j < tmpSem.size(); j++) {
        if (UniLab[i] != tmpSem[j]) continue;
        keys.push_back(tmpKeys[j]); dep.push_back(tmpDepth[j]); sem.push_back(tmpSem[j]); fl.push_back(tmpFlow[j]);
        corres.push_back(tmpCorres[j]); inl.push_back(-1); lab.push_back(-2);
      }
    }
    float Twc[16];
    inv44(cur->Tcw, Twc);
    std::vector<P3> p3(keys.size());
    for (size_t i = 0; i < keys.size(); i++) p3[i] = to_world(unproject_cam(keys[i], dep[i]), Twc);
    cur->mvObjKeys = keys; cur->mvObjDepth = dep; cur->mvObj3DPoint = p3; cur->mvObjCorres = corres; cur->mvObjFlowNext = fl;
    cur->vSemObjLabel = sem; cur->nDynInlierID = inl; cur->vObjLabel = lab;
  }

  // ---- Tracking::GetDynamicTrackNew (src/Tracking.cc:2615-2720), from frame 0 or incrementally (same chains)
  void rebuild_dyn_tracklets() {
    const auto& TM = map.vnAssoDyn;
    const auto& OL = map.vnFeatLabel;
    std::vector<int> pre;
    std::vector<std::vector<std::pair<int, int>>> T;
    std::vector<int> oid;
    for (int i = 0; i < (int)TM.size(); i++) {
      std::vector<int> curc(TM[i].size(), -1);
      for (size_t j = 0; j < TM[i].size(); j++) {
        if (TM[i][j] == -1) continue;
        if (i > 0 && pre[TM[i][j]] != -1) {
          T[pre[TM[i][j]]].push_back({i + 1, (int)j});
          curc[j] = pre[TM[i][j]];
        } else {
          T.push_back({{i, TM[i][j]}, {i + 1, (int)j}});
          oid.push_back(OL[i][j]);
          curc[j] = (int)T.size() - 1;
        }
      }
      pre = curc;
    }
    map.TrackletDyn.swap(T);
    map.nObjID.swap(oid);
  }
  void extend_dyn_tracklets() {
    const auto& TM = map.vnAssoDyn;
    const int i = (int)TM.size() - 1;
    std::vector<int> curc(TM[i].size(), -1);
    for (size_t j = 0; j < TM[i].size(); j++) {
      if (TM[i][j] == -1) continue;
      if (i > 0 && dynTrackOfPrev[TM[i][j]] != -1) {
        map.TrackletDyn[dynTrackOfPrev[TM[i][j]]].push_back({i + 1, (int)j});
        curc[j] = dynTrackOfPrev[TM[i][j]];
      } else {
        map.TrackletDyn.push_back({{i, TM[i][j]}, {i + 1, (int)j}});
        map.nObjID.push_back(map.vnFeatLabel[i][j]);
        curc[j] = (int)map.TrackletDyn.size() - 1;
      }
    }
    dynTrackOfPrev = curc;
  }

  // ---- Optimizer::PartialBatchOptimization
  void partial_batch(int WINDOW, vo_track_stats* st) {
    const int N = (int)map.vpFeatSta.size();
    if (st) { st->ba_iterations = -1; st->ba_points = 0; st->ba_obs = 0; }
    if (WINDOW <= 0) return;
    const auto& Tr = map.TrackletSta;  // (the reference copies it, :46)
    std::vector<std::vector<int>> lab(N), mak(N);
    for (int i = 0; i < N; i++) { lab[i].assign(map.vpFeatSta[i].size(), -1); mak[i] = lab[i]; }
    for (size_t t = 0; t < Tr.size(); t++) {
      if (Tr[t].size() < 3) continue;
      for (auto& e : Tr[t]) lab[e.first][e.second] = (int)t;
    }
    const int start = N - WINDOW;
    std::vector<float> poses, rel, pts, oxyz;
    std::vector<int> op, ol;
    std::vector<std::pair<int, int>> ptOwner;  // first (frame, feat) of every point
    for (int i = start; i < N; i++) {
      poses.insert(poses.end(), map.vmCameraPose[i].begin(), map.vmCameraPose[i].end());
      if (i != start) rel.insert(rel.end(), map.vmRigidMotion[i - 1].begin(), map.vmRigidMotion[i - 1].end());
      for (size_t j = 0; j < lab[i].size(); j++) {
        const int t = lab[i][j];
        if (t == -1) continue;
        int pos = -1;
        for (size_t k = 0; k < Tr[t].size(); k++)
          if (Tr[t][k].first == i && Tr[t][k].second == (int)j) { pos = (int)k; break; }
        if (pos == -1) continue;
        int pid;
        if (pos == 0) {
          pid = (int)ptOwner.size();
          ptOwner.push_back({i, (int)j});
          const P3& Xw = map.vp3DPointSta[i][j];
          pts.push_back(Xw.x); pts.push_back(Xw.y); pts.push_back(Xw.z);
        } else {
          pid = mak[Tr[t][pos - 1].first][Tr[t][pos - 1].second];
          if (pid == -1) continue;
        }
        mak[i][j] = pid;
        const P3 xc = unproject_cam(map.vpFeatSta[i][j], map.vfDepSta[i][j]);
        op.push_back(i - start); ol.push_back(pid);
        oxyz.push_back(xc.x); oxyz.push_back(xc.y); oxyz.push_back(xc.z);
      }
    }
    vo_ba_problem pr;
    memset(&pr, 0, sizeof pr);
    vo_ba_default_params(&pr);
    pr.n_poses = WINDOW; pr.n_points = (int)ptOwner.size(); pr.n_obs = (int)op.size();
    pr.poses = poses.data(); pr.rel_motion = rel.data(); pr.points = pts.data();
    pr.obs_pose = op.data(); pr.obs_point = ol.data(); pr.obs_xyz = oxyz.data();
    vo_lm_stats ls;
    const int its = vo_ba_partial(&pr, &ls);
    if (st) { st->ba_iterations = its; st->ba_points = pr.n_points; st->ba_obs = pr.n_obs; st->ba_trials = ls.total_trials; }
    for (int i = start; i < N; i++) {
      std::copy(poses.begin() + 16 * (i - start), poses.begin() + 16 * (i - start + 1), map.vmCameraPose[i].begin());
      if (i > start) std::copy(rel.begin() + 16 * (i - start - 1), rel.begin() + 16 * (i - start), map.vmRigidMotion[i - 1].begin());
    }
    for (int i = start; i < N; i++)
      for (size_t j = 0; j < mak[i].size(); j++)
        if (mak[i][j] != -1) map.vp3DPointSta[i][j] = {pts[3 * mak[i][j]], pts[3 * mak[i][j] + 1], pts[3 * mak[i][j] + 2]};
  }

  // ---- Tracking::RenewFrameInfo (static part)
  void renew(const std::vector<int>& TM_sta, const float* depth, const float* flow, const int32_t* mask) {
    const int W = cfg.width, H = cfg.height, maxn = cfg.max_track_bg;
    std::vector<P2> keys, corres, fl;
    std::vector<int> inl;
    auto try_add = [&](const P2& kp, int inlier_id) -> bool {
      const int x = (int)kp.x, y = (int)kp.y;
      if (x >= W || y >= H || x <= 0 || y <= 0) return false;
      const size_t k = (size_t)y * W + x;
      if (mask[k] != 0) return false;
      if (depth[k] > 40 || depth[k] <= 0) return false;
      const float fx = flow[2 * k], fy = flow[2 * k + 1];
      if (fx != 0 && fy != 0) {
        if (kp.x + fx < W && kp.y + fy < H && kp.x + fx > 0 && kp.y + fy > 0) {
          keys.push_back(kp);
          corres.push_back({kp.x + fx, kp.y + fy});
          fl.push_back({fx, fy});
          inl.push_back(inlier_id);
          return true;
        }
      }
      return false;
    };
    for (size_t i = 0; i < TM_sta.size(); i++) {
      if (TM_sta[i] == -1) continue;
      try_add(cur->mvStatKeys[TM_sta[i]], TM_sta[i]);
      if ((int)keys.size() > maxn) break;
    }
    int tot = (int)keys.size(), start_id = 0;
    const int step = 20;
    const std::vector<P2> check = keys;
    while (tot < maxn) {
      if (start_id == step) break;
      for (size_t i = start_id; i < cur->mvKeys.size(); i += step) {
        const P2 s = {cur->mvKeys[i].x, cur->mvKeys[i].y};
        float min_dist = 100;
        bool used = false;
        for (size_t j = 0; j < check.size(); j++) {
          const float d = std::sqrt((check[j].x - s.x) * (check[j].x - s.x) + (check[j].y - s.y) * (check[j].y - s.y));
          if (d < min_dist) min_dist = d;
          if (min_dist < 1.0) { used = true; break; }
        }
        if (used) continue;
        if (try_add(s, -1)) tot++;
        if (tot >= maxn) break;
      }
      start_id++;
    }
    const size_t n = keys.size();
    std::vector<float> dep(n, -1.f);
    std::vector<P3> p3(n);
    float Twc[16];
    inv44(cur->Tcw, Twc);
    for (size_t i = 0; i < n; i++) {
      const float d = depth[(size_t)(int)keys[i].y * W + (int)keys[i].x];
      if (d > 0) dep[i] = d;
      p3[i] = to_world(unproject_cam(keys[i], dep[i]), Twc);
    }
    cur->nStaInlierID = inl;
    cur->mvStatKeysTmp = keys;
    cur->mvStatDepthTmp = dep;
    cur->mvStat3DPointTmp = p3;
    cur->mvFlowNext = fl;
    cur->mvCorres = corres;
  }


  // ---- Optimizer::FullBatchOptimization, graph construction (src/Optimizer.cc:1235-1745) in flat form
  struct FullGraph {
    std::vector<float> se3, points, e6_meas, obs_xyz;
    std::vector<int> e6_i, e6_j, e6_kind, obs_se3, obs_point, obs_kind, tern_p1, tern_p2, tern_h;
    int n_poses = 0, n_motions = 0;
    std::vector<std::vector<int>> VertexID;        // [frame-1][object entry] -> SE3 vertex (camera entry excluded)
    std::vector<std::vector<int>> makSta, makDyn;  // per frame, per feature: point vertex or -1
  };
  FullGraph build_full_graph() {
    FullGraph G;
    const int N = (int)map.vpFeatSta.size();
    const auto& StaTracks = map.TrackletSta;
    const auto& DynTracks = map.TrackletDyn;
    std::vector<std::vector<int>> labSta(N), labDyn(N), posSta(N), posDyn(N);
    G.makSta.resize(N); G.makDyn.resize(N);
    for (int i = 0; i < N; i++) {
      labSta[i].assign(map.vpFeatSta[i].size(), -1); posSta[i] = labSta[i]; G.makSta[i] = labSta[i];
      labDyn[i].assign(map.vpFeatDyn[i].size(), -1); posDyn[i] = labDyn[i]; G.makDyn[i] = labDyn[i];
    }
    for (size_t t = 0; t < StaTracks.size(); t++) {
      if (StaTracks[t].size() < 3) continue;
      for (size_t k = 0; k < StaTracks[t].size(); k++) { labSta[StaTracks[t][k].first][StaTracks[t][k].second] = (int)t; posSta[StaTracks[t][k].first][StaTracks[t][k].second] = (int)k; }
    }
    for (size_t t = 0; t < DynTracks.size(); t++) {
      if (DynTracks[t].size() < 3) continue;
      for (size_t k = 0; k < DynTracks[t].size(); k++) { labDyn[DynTracks[t][k].first][DynTracks[t][k].second] = (int)t; posDyn[DynTracks[t][k].first][DynTracks[t][k].second] = (int)k; }
    }
    G.n_poses = N;
    for (int i = 0; i < N; i++) G.se3.insert(G.se3.end(), map.vmCameraPose[i].begin(), map.vmCameraPose[i].end());
    G.VertexID.resize(std::max(N - 1, 0));
    int next_se3 = N;
    float I16[16];
    eye44(I16);
    auto add_point = [&](const P3& Xw) { G.points.push_back(Xw.x); G.points.push_back(Xw.y); G.points.push_back(Xw.z); return (int)(G.points.size() / 3) - 1; };
    auto add_obs = [&](int se3v, int pt, int kind, const P3& xc) {
      G.obs_se3.push_back(se3v); G.obs_point.push_back(pt); G.obs_kind.push_back(kind);
      G.obs_xyz.push_back(xc.x); G.obs_xyz.push_back(xc.y); G.obs_xyz.push_back(xc.z);
    };
    for (int i = 0; i < N; i++) {
      if (i != 0) {
        G.e6_i.push_back(i - 1); G.e6_j.push_back(i); G.e6_kind.push_back(0);
        G.e6_meas.insert(G.e6_meas.end(), map.vmRigidMotion[i - 1].begin(), map.vmRigidMotion[i - 1].end());
      }
      // static features
      for (size_t j = 0; j < labSta[i].size(); j++) {
        const int t = labSta[i][j];
        if (t == -1) continue;
        const int pos = i == 0 ? 0 : posSta[i][j];
        int pid;
        if (pos == 0) pid = add_point(map.vp3DPointSta[i][j]);
        else pid = G.makSta[StaTracks[t][pos - 1].first][StaTracks[t][pos - 1].second];
        if (pid == -1) continue;
        add_obs(i, pid, 0, unproject_cam(map.vpFeatSta[i][j], map.vfDepSta[i][j]));
        G.makSta[i][j] = pid;
      }
      // dynamic part
      if (i == 0) {
        for (size_t j = 0; j < labDyn[i].size(); j++) {
          if (labDyn[i][j] == -1) continue;
          const int pid = add_point(map.vp3DPointDyn[i][j]);
          add_obs(i, pid, 1, unproject_cam(map.vpFeatDyn[i][j], map.vfDepDyn[i][j]));
          G.makDyn[i][j] = pid;
        }
        continue;
      }
      const auto& labels = map.vnRMLabel[i - 1];
      std::vector<int> ObjUniqueID(labels.size(), -1);
      G.VertexID[i - 1].assign(labels.size(), -1);
      for (size_t j = 0; j < labels.size(); j++) {
        G.se3.insert(G.se3.end(), I16, I16 + 16);   // object motions start from identity (:1597)
        const int vid = next_se3++;
        if (i > 2) {  // smoothness w.r.t. the same object's motion in the previous frame (SMOOTH_CONSTRAINT && i>2)
          int TraceID = -1;
          for (size_t k = 0; k < map.vnRMLabel[i - 2].size(); k++)
            if (map.vnRMLabel[i - 2][k] == labels[j]) { TraceID = (int)k; break; }
          if (TraceID != -1) {
            G.e6_i.push_back(G.VertexID[i - 2][TraceID]); G.e6_j.push_back(vid); G.e6_kind.push_back(1);
            G.e6_meas.insert(G.e6_meas.end(), I16, I16 + 16);
          }
        }
        ObjUniqueID[j] = vid;
        G.VertexID[i - 1][j] = vid;
      }
      for (size_t j = 0; j < labDyn[i].size(); j++) {
        const int t = labDyn[i][j];
        if (t == -1) continue;
        const int pos = posDyn[i][j];
        int ObjPositionID = -1;
        for (size_t k = 0; k < labels.size(); k++)
          if (labels[k] == map.nObjID[t]) { ObjPositionID = ObjUniqueID[k]; break; }
        if (ObjPositionID == -1 && pos != 0) continue;
        int prev = -1;
        if (pos != 0) {
          prev = G.makDyn[DynTracks[t][pos - 1].first][DynTracks[t][pos - 1].second];
          if (prev == -1) continue;  // (the reference would dereference a null vertex here; cannot happen for chains built by Track())
        }
        const int pid = add_point(map.vp3DPointDyn[i][j]);
        add_obs(i, pid, 1, unproject_cam(map.vpFeatDyn[i][j], map.vfDepDyn[i][j]));
        if (pos != 0) { G.tern_p1.push_back(prev); G.tern_p2.push_back(pid); G.tern_h.push_back(ObjPositionID); }
        G.makDyn[i][j] = pid;
      }
    }
    G.n_motions = next_se3 - N;
    return G;
  }
  void fill_problem(FullGraph& G, vo_fba_problem& pr) {
    memset(&pr, 0, sizeof pr);
    vo_fba_default_params(&pr);
    pr.n_poses = G.n_poses; pr.n_motions = G.n_motions; pr.n_points = (int)(G.points.size() / 3);
    pr.n_obs = (int)G.obs_se3.size(); pr.n_e6 = (int)G.e6_i.size(); pr.n_tern = (int)G.tern_p1.size();
    pr.se3 = G.se3.data(); pr.points = G.points.data();
    pr.e6_i = G.e6_i.data(); pr.e6_j = G.e6_j.data(); pr.e6_kind = G.e6_kind.data(); pr.e6_meas = G.e6_meas.data();
    pr.obs_se3 = G.obs_se3.data(); pr.obs_point = G.obs_point.data(); pr.obs_kind = G.obs_kind.data(); pr.obs_xyz = G.obs_xyz.data();
    pr.tern_p1 = G.tern_p1.data(); pr.tern_p2 = G.tern_p2.data(); pr.tern_h = G.tern_h.data();
  }
  std::vector<std::vector<float>> vmCameraPose_RF;
  std::vector<std::vector<std::array<float, 16>>> vmObjMotion_RF;
  int full_batch(vo_lm_stats* stats, int32_t* sizes) {
    if (cfg.rebuild_tracklets == 0) { rebuild_tracklets(); rebuild_dyn_tracklets(); }
    FullGraph G = build_full_graph();
    vo_fba_problem pr;
    fill_problem(G, pr);
    if (sizes) { sizes[0] = pr.n_poses; sizes[1] = pr.n_motions; sizes[2] = pr.n_points; sizes[3] = pr.n_obs; sizes[4] = pr.n_e6; sizes[5] = pr.n_tern; }
    const int its = vo_ba_full(&pr, stats);
    // write-back (src/Optimizer.cc:2090-2176): refined poses of frames >= 1, refined object motions, points in place
    const int N = G.n_poses;
    vmCameraPose_RF = map.vmCameraPose;
    vmObjMotion_RF = map.vmObjMotion;
    for (int i = 1; i < N; i++) std::copy(G.se3.begin() + 16 * (size_t)i, G.se3.begin() + 16 * (size_t)(i + 1), vmCameraPose_RF[i].begin());
    for (int i = 0; i + 1 < N; i++)
      for (size_t j = 0; j < G.VertexID[i].size(); j++) memcpy(vmObjMotion_RF[i][j].data(), &G.se3[16 * (size_t)G.VertexID[i][j]], sizeof(float) * 16);
    for (int i = 0; i < N; i++) {
      for (size_t j = 0; j < G.makSta[i].size(); j++)
        if (G.makSta[i][j] != -1) map.vp3DPointSta[i][j] = {G.points[3 * (size_t)G.makSta[i][j]], G.points[3 * (size_t)G.makSta[i][j] + 1], G.points[3 * (size_t)G.makSta[i][j] + 2]};
      for (size_t j = 0; j < G.makDyn[i].size(); j++)
        if (G.makDyn[i][j] != -1) map.vp3DPointDyn[i][j] = {G.points[3 * (size_t)G.makDyn[i][j]], G.points[3 * (size_t)G.makDyn[i][j] + 1], G.points[3 * (size_t)G.makDyn[i][j] + 2]};
    }
    return its;
  }

  int track(const uint8_t* gray, float* depth, const float* flow, const int32_t* mask, float* Tcw_out, vo_track_stats* st) {
    const int W = cfg.width, H = cfg.height;
    if (st) memset(st, 0, sizeof *st);
    double t0 = now_ms();
    vo_depth_prep(depth, W, H, W, cfg.choose_data, cfg.depth_map_factor, cfg.bf, 1.0f);
    // ---- UpdateMask (src/Tracking.cc:353-364): the reference edits the caller's mask in place; the oracle edits a copy
    if (initialised && !last->vSemObjLabel.empty() && !segLast.empty()) {
      segCur.assign(mask, mask + (size_t)W * H);
      std::vector<float> cor(2 * last->mvObjCorres.size());
      for (size_t i = 0; i < last->mvObjCorres.size(); i++) { cor[2 * i] = last->mvObjCorres[i].x; cor[2 * i + 1] = last->mvObjCorres[i].y; }
      std::vector<int32_t> uniq(last->vSemObjLabel.size() + 1), rec(last->vSemObjLabel.size() + 1);
      const int nu = vo_update_mask(last->vSemObjLabel.data(), cor.data(), (int)last->vSemObjLabel.size(), segLast.data(), flowLast.data(),
                                    segCur.data(), W, H, uniq.data(), rec.data(), (int)uniq.size());
      if (st) for (int i = 0; i < nu; i++) st->n_masks_recovered += rec[i] ? 1 : 0;
      mask = segCur.data();
    }
    cur = new Frame();
    eye44(cur->Tcw);
    cur->mvKeys.resize(cfg.orb.nfeatures + 64);
    int nk = vo_orb_extract(gray, W, H, W, &cfg.orb, cur->mvKeys.data(), (int)cur->mvKeys.size());
    cur->mvKeys.resize(nk < 0 ? 0 : nk);
    double t1 = now_ms();
    {  // Frame ctor association
      const int n = (int)cur->mvKeys.size();
      std::vector<int> idx(n);
      std::vector<float> cor(2 * (size_t)n), fl(2 * (size_t)n), dep(n);
      const int m = vo_frame_associate(cur->mvKeys.data(), n, depth, flow, mask, W, H, cfg.th_depth_bg, idx.data(), cor.data(),
                                       fl.data(), dep.data(), n);
      for (int i = 0; i < m; i++) {
        cur->mvStatKeysTmp.push_back({cur->mvKeys[idx[i]].x, cur->mvKeys[idx[i]].y});
        cur->mvCorres.push_back({cor[2 * i], cor[2 * i + 1]});
        cur->mvFlowNext.push_back({fl[2 * i], fl[2 * i + 1]});
        cur->mvStatDepthTmp.push_back(dep[i]);
      }
    }
    {  // Frame ctor: stride-4 object sampling (src/Frame.cc:184-211)
      const int cap = ((W + 3) / 4) * ((H + 3) / 4);
      std::vector<float> k2(2 * (size_t)cap), c2(2 * (size_t)cap), f2(2 * (size_t)cap), d1(cap);
      std::vector<int32_t> l1(cap);
      const int m = vo_frame_sample_objects(depth, flow, mask, W, H, cfg.th_depth_obj, k2.data(), c2.data(), f2.data(), d1.data(), l1.data(), cap);
      tmpKeys.resize(m); tmpCorres.resize(m); tmpFlow.resize(m); tmpDepth.assign(d1.begin(), d1.begin() + m); tmpSem.assign(l1.begin(), l1.begin() + m);
      for (int i = 0; i < m; i++) { tmpKeys[i] = {k2[2 * i], k2[2 * i + 1]}; tmpCorres[i] = {c2[2 * i], c2[2 * i + 1]}; tmpFlow[i] = {f2[2 * i], f2[2 * i + 1]}; }
    }
    if (initialised) {  // GrabImageRGBD :391-421: object features = last frame's correspondences, fresh depth / label lookups
      cur->mvObjKeys = last->mvObjCorres;
      const size_t no = cur->mvObjKeys.size();
      cur->mvObjDepth.assign(no, -1.f);
      cur->vSemObjLabel.assign(no, -1);
      for (size_t i = 0; i < no; i++) {
        const int u = (int)cur->mvObjKeys[i].x, v = (int)cur->mvObjKeys[i].y;
        if (u < (W - 1) && u > 0 && v < (H - 1) && v > 0 && depth[(size_t)v * W + u] < cfg.th_depth_obj && depth[(size_t)v * W + u] > 0) {
          cur->mvObjDepth[i] = depth[(size_t)v * W + u];
          cur->vSemObjLabel[i] = mask[(size_t)v * W + u];
        } else {
          cur->mvObjDepth[i] = 0.1f;
          cur->vSemObjLabel[i] = 0;
        }
      }
      cur->vObjLabel.assign(no, -2);
    } else {
      cur->mvObjKeys = tmpKeys; cur->mvObjCorres = tmpCorres; cur->mvObjFlowNext = tmpFlow; cur->mvObjDepth = tmpDepth;
      cur->vSemObjLabel = tmpSem;
      cur->vObjLabel.assign(tmpKeys.size(), -2);
    }
    if (initialised) {  // GrabImageRGBD :369-389
      cur->mvStatKeys = last->mvCorres;
      cur->mvStatDepth.assign(cur->mvStatKeys.size(), -1.f);
      for (size_t i = 0; i < cur->mvStatKeys.size(); i++) {
        const int v = (int)cur->mvStatKeys[i].y, u = (int)cur->mvStatKeys[i].x;
        if (u < (W - 1) && u > 0 && v < (H - 1) && v > 0) {
          const float d = depth[(size_t)v * W + u];
          if (d > 0) cur->mvStatDepth[i] = d;
        }
      }
    }
    double t2 = now_ms();
    if (st) { st->ms_orb = t1 - t0; st->ms_assoc = t2 - t1; st->n_keypoints = (int)cur->mvKeys.size(); }
    int rc = 0;
    if (!initialised) {
      // ---- Tracking::Initialization
      for (size_t i = 0; i < cur->mvStatKeysTmp.size(); i++)
        cur->mvStat3DPointTmp.push_back(unproject_cam(cur->mvStatKeysTmp[i], cur->mvStatDepthTmp[i]));
      map.vpFeatSta.push_back(cur->mvStatKeysTmp);
      map.vfDepSta.push_back(cur->mvStatDepthTmp);
      map.vp3DPointSta.push_back(cur->mvStat3DPointTmp);
      for (size_t i = 0; i < cur->mvObjKeys.size(); i++) cur->mvObj3DPoint.push_back(unproject_cam(cur->mvObjKeys[i], cur->mvObjDepth[i]));
      map.vpFeatDyn.push_back(cur->mvObjKeys);
      map.vfDepDyn.push_back(cur->mvObjDepth);
      map.vp3DPointDyn.push_back(cur->mvObj3DPoint);
      if (st) st->n_dyn_features = (int)cur->mvObjKeys.size();
      std::vector<float> I(16, 0.f);
      I[0] = I[5] = I[10] = I[15] = 1.f;
      map.vmCameraPose.push_back(I);
      eye44(cur->Tcw);
      if (vio) {   // src/Tracking.cc:1555-1561: IMU pose from Tcb, zero velocity, empty preintegration
        ImuFrame f0;
        float Rwb0[9], twb0[3], V0[3] = {0.f, 0.f, 0.f};
        for (int r = 0; r < 3; r++) {
          for (int c = 0; c < 3; c++) Rwb0[3 * r + c] = Tcb[4 * r + c];
          twb0[r] = Tcb[4 * r + 3];
        }
        set_imu_pose_velocity(f0, Rwb0, twb0, V0);
        memset(&f0.pre, 0, sizeof f0.pre);
        f0.pre.dR[0] = f0.pre.dR[4] = f0.pre.dR[8] = 1.f;
        f0.has_pre = true;
        f0.t = f0.t_prev = next_t;
        memcpy(cur->Tcw, f0.Tcw, sizeof f0.Tcw);
        fr.clear();
        fr.push_back(f0);
      }
      last = cur;
      last->mvStatKeys = cur->mvStatKeysTmp;
      last->mvStatDepth = cur->mvStatDepthTmp;
      initialised = true;
    } else {
      const int Ns = (int)cur->mvStatKeys.size();
      if (Ns < 2) { rc = 1; }
      else {
        if (vio) {   // src/Tracking.cc:1115-1119 (Frame ctor: mVw of the previous frame, src/Frame.cc:437-447)
          ImuFrame f;
          const ImuFrame& pf = fr.back();
          f.t = next_t; f.t_prev = pf.t;
          memcpy(f.vel, pf.vel, sizeof f.vel);
          memcpy(f.bias, pf.bias, sizeof f.bias);
          preintegrate(f, pf);
          fr.push_back(f);
        }
        // ---- GetInitModelCam
        std::vector<float> cur2d(2 * (size_t)Ns), p3d(3 * (size_t)Ns, 0.f);
        std::vector<int> valid(Ns, 1), ids(Ns);
        float Twl[16];
        inv44(last->Tcw, Twl);
        for (int i = 0; i < Ns; i++) {
          cur2d[2 * i] = cur->mvStatKeys[i].x; cur2d[2 * i + 1] = cur->mvStatKeys[i].y;
          const float z = last->mvStatDepth[i];
          if (z < 0) { valid[i] = 0; continue; }
          const P3 xw = to_world(unproject_cam(last->mvStatKeys[i], z), Twl);
          p3d[3 * i] = xw.x; p3d[3 * i + 1] = xw.y; p3d[3 * i + 2] = xw.z;
        }
        vo_pnp_problem pp;
        memset(&pp, 0, sizeof pp);
        vo_pnp_default_params(&pp);
        pp.n = Ns; pp.cur_xy = cur2d.data(); pp.pts3d = p3d.data(); pp.valid = valid.data(); pp.inlier_ids = ids.data();
        if (has_velocity) mul44(mVelocity, last->Tcw, pp.Tcw_motion);
        else memcpy(pp.Tcw_motion, last->Tcw, sizeof(float) * 16);
        pp.fx = cfg.fx; pp.fy = cfg.fy; pp.cx = cfg.cx; pp.cy = cfg.cy;
        vo_init_model_cam(&pp);
        std::vector<int> TM_sub(ids.begin(), ids.begin() + pp.n_inliers);
        memcpy(cur->Tcw, pp.Tcw_out, sizeof(float) * 16);
        double t3 = now_ms();
        int ninl = 0;
        if (cfg.b_joint) {
          // ---- PoseOptimizationFlow2Cam
          const int n = (int)TM_sub.size();
          std::vector<float> obs(2 * (size_t)n), fl(2 * (size_t)n), dep(n), fo(2 * (size_t)n);
          std::vector<int> inl(n);
          for (int i = 0; i < n; i++) {
            const int k = TM_sub[i];
            obs[2 * i] = last->mvStatKeys[k].x; obs[2 * i + 1] = last->mvStatKeys[k].y;
            fl[2 * i] = last->mvFlowNext[k].x; fl[2 * i + 1] = last->mvFlowNext[k].y;
            dep[i] = last->mvStatDepth[k];
          }
          vo_poseopt_problem po;
          memset(&po, 0, sizeof po);
          vo_poseopt_default_params(&po);
          po.n = n; po.obs_xy = obs.data(); po.flow_xy = fl.data(); po.depth = dep.data();
          memcpy(po.Tcw_init, cur->Tcw, sizeof(float) * 16);
          memcpy(po.Tcw_last, last->Tcw, sizeof(float) * 16);
          po.fx = cfg.fx; po.fy = cfg.fy; po.cx = cfg.cx; po.cy = cfg.cy;
          po.flow_out = fo.data(); po.inlier = inl.data();
          ninl = vo_poseopt_flow2cam(&po, nullptr);
          memcpy(cur->Tcw, po.Tcw_out, sizeof(float) * 16);
          if (n >= 3) {
            for (int i = 0; i < n; i++) {
              if (inl[i]) {
                const int k = TM_sub[i];
                cur->mvStatKeys[k].x = (float)((double)last->mvStatKeys[k].x + (double)fo[2 * i]);
                cur->mvStatKeys[k].y = (float)((double)last->mvStatKeys[k].y + (double)fo[2 * i + 1]);
              } else TM_sub[i] = -1;
            }
          }
        } else {
          // ---- PoseOptimizationNew (src/Optimizer.cc:2180-2334): reprojection of the last frame's world points (no depth noise)
          const int n = (int)TM_sub.size();
          std::vector<float> obs(2 * (size_t)n), p3(3 * (size_t)n);
          std::vector<int> inl(n);
          for (int i = 0; i < n; i++) {
            const int k = TM_sub[i];
            obs[2 * i] = cur->mvStatKeys[k].x; obs[2 * i + 1] = cur->mvStatKeys[k].y;
            p3[3 * i] = p3d[3 * k]; p3[3 * i + 1] = p3d[3 * k + 1]; p3[3 * i + 2] = p3d[3 * k + 2];
          }
          vo_projopt_problem pj;
          memset(&pj, 0, sizeof pj);
          vo_projopt_default_params(&pj, 0);
          pj.n = n; pj.obs_xy = obs.data(); pj.pts3d = p3.data(); pj.inlier = inl.data();
          memcpy(pj.T_init, cur->Tcw, sizeof(float) * 16);
          pj.fx = cfg.fx; pj.fy = cfg.fy; pj.cx = cfg.cx; pj.cy = cfg.cy;
          ninl = vo_pose_opt_proj(&pj, nullptr);
          memcpy(cur->Tcw, pj.T_out, sizeof(float) * 16);
          if (n >= 3)
            for (int i = 0; i < n; i++)
              if (!inl[i]) TM_sub[i] = -1;
        }
        double t4 = now_ms();
        // ---- motion model (src/Tracking.cc:1142-1148)
        float LastTwc[16];
        inv44(last->Tcw, LastTwc);
        mul44(cur->Tcw, LastTwc, mVelocity);
        has_velocity = true;
        // ---- scene flow, object tracking, object motions (src/Tracking.cc:1160-1308)
        std::vector<std::vector<int>> ObjIdNew;
        if (!cur->mvObjKeys.empty()) {
          scene_flow();
          DynRecord rec;
          if (dyn_log_on) {
            rec.f_id = f_id; rec.max_id_before = max_id;
            rec.sem = cur->vSemObjLabel; rec.lab_before = cur->vObjLabel; rec.last_sem = last->vSemObjLabel;
            rec.last_sem_pos = last->nSemPosition; rec.last_mod = last->nModLabel;
            for (char b : last->bObjStat) rec.last_stat.push_back(b ? 1 : 0);
            for (size_t i = 0; i < cur->mvObjKeys.size(); i++) {
              rec.key_xy.push_back(cur->mvObjKeys[i].x); rec.key_xy.push_back(cur->mvObjKeys[i].y);
              rec.depth.push_back(cur->mvObjDepth[i]);
              rec.flow3.push_back(cur->vFlow_3d[i].x); rec.flow3.push_back(cur->vFlow_3d[i].y); rec.flow3.push_back(cur->vFlow_3d[i].z);
            }
          }
          ObjIdNew = dyn_obj_tracking();
          if (dyn_log_on) {
            rec.max_id_after = max_id; rec.lab_after = cur->vObjLabel; rec.out_mod = cur->nModLabel; rec.out_sem_pos = cur->nSemPosition;
            for (auto& o : ObjIdNew) { rec.out_len.push_back((int)o.size()); rec.out_ids.insert(rec.out_ids.end(), o.begin(), o.end()); }
            dyn_log.push_back(std::move(rec));
          }
          object_motions(ObjIdNew);
        }
        double t4b = now_ms();
        // ---- RenewFrameInfo + map bookkeeping
        renew(TM_sub, depth, flow, mask);
        renew_objects(depth, flow, mask);
        Frame* old = last;
        last = cur;
        last->mvStatKeys = cur->mvStatKeysTmp;
        last->mvStatDepth = cur->mvStatDepthTmp;
        delete old;
        map.vpFeatSta.push_back(cur->mvStatKeysTmp);
        map.vfDepSta.push_back(cur->mvStatDepthTmp);
        map.vp3DPointSta.push_back(cur->mvStat3DPointTmp);
        map.vnAssoSta.push_back(cur->nStaInlierID);
        map.vpFeatDyn.push_back(cur->mvObjKeys);
        map.vfDepDyn.push_back(cur->mvObjDepth);
        map.vp3DPointDyn.push_back(cur->mvObj3DPoint);
        map.vnAssoDyn.push_back(cur->nDynInlierID);
        map.vnFeatLabel.push_back(cur->vObjLabel);
        if (cfg.rebuild_tracklets) { rebuild_tracklets(); rebuild_dyn_tracklets(); }
        else { extend_tracklets(); extend_dyn_tracklets(); }
        {  // (6.2) object motions and labels, (10) centres (src/Tracking.cc:1390-1422)
          std::vector<std::array<float, 16>> mot_o;
          std::vector<int> rml, sml;
          std::vector<P3> cen;
          for (size_t i = 0; i < cur->vObjMod.size(); i++) {
            if (!cur->bObjStat[i]) continue;
            mot_o.push_back(cur->vObjMod[i]); rml.push_back(cur->nModLabel[i]); sml.push_back(cur->nSemPosition[i]);
            cen.push_back(cur->vObjCentre3D[i]);
          }
          map.vmObjMotion.push_back(mot_o); map.vnRMLabel.push_back(rml); map.vnSMLabel.push_back(sml); map.vmRigidCentre.push_back(cen);
        }
        std::vector<float> Twc(16), mot(16);
        if (vio) memcpy(fr.back().Tcw, cur->Tcw, sizeof(float) * 16);
        inv44(cur->Tcw, Twc.data());
        inv44(mVelocity, mot.data());
        map.vmCameraPose.push_back(Twc);
        map.vmRigidMotion.push_back(mot);
        double t5 = now_ms();
        if (st) {
          st->ms_init = t3 - t2; st->ms_poseopt = t4 - t3; st->ms_renew = t5 - t4; (void)t4b;
          st->n_dyn_features = (int)cur->mvObjKeys.size(); st->n_objects = (int)cur->vObjMod.size();
          for (char b : cur->bObjStat) st->n_objects_ok += b ? 1 : 0;
          st->n_matches = Ns; st->n_init_inliers = pp.n_inliers; st->init_winner = pp.winner; st->n_pose_inliers = ninl;
          st->n_static = (int)cur->mvStatKeysTmp.size();
        }
      }
    }
    memcpy(Tcw_out, cur->Tcw, sizeof(float) * 16);
    // mSegMapLast / mFlowMapLast (src/Tracking.cc:777-780); only needed while object features are alive
    if (rc == 0 && !last->vSemObjLabel.empty()) {
      segLast.assign(mask, mask + (size_t)W * H);
      flowLast.assign(flow, flow + 2 * (size_t)W * H);
    } else if (rc == 0) { segLast.clear(); flowLast.clear(); }
    // ---- PartialBatchOptimization, every frame (src/Tracking.cc:1428-1451)
    double t6 = now_ms();
    const int window = f_id < cfg.window_size ? f_id : cfg.window_size;
    if (rc == 0) partial_batch(window, st);
    if (st) st->ms_ba = now_ms() - t6;
    if (vio && rc == 0 && f_id > 0) {   // InitializeIMU / ScaleRefinement (src/Tracking.cc:1452-1480); TrackRGBD returns mTcw after them
      vio_after_ba();
      memcpy(Tcw_out, cur->Tcw, sizeof(float) * 16);
    }
    f_id++;
    if (rc != 0 && cur != last) { delete cur; cur = last; }
    return rc;
  }
};


// Map::ApplyScaledRotation(R, s, bScaledVel = true, t = 0) (src/Map.cc:55-119)
void Tracker::apply_scaled_rotation(const float* R, float s) {
  float Tyw[16];
  eye44(Tyw);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) Tyw[4 * r + c] = R[3 * r + c];
  auto rot_pt = [&](P3& p) {   // s * Ryw * p + tyw: one gemm with alpha = s, beta = 1 (tyw = 0)
    const float x = p.x, y = p.y, z = p.z;
    float o[3];
    for (int r = 0; r < 3; r++)
      o[r] = (float)((double)s * ((double)R[3 * r] * x + (double)R[3 * r + 1] * y + (double)R[3 * r + 2] * z) + 0.0);
    p = {o[0], o[1], o[2]};
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
  if (vio) {
    for (size_t i = 1; i < fr.size(); i++) {   // Map::vpFrames
      frame_pose(fr[i].Tcw);
      mv3(R, fr[i].vel, fr[i].vel, (double)s);   // Ryw * Vw * s
    }
    if (last && fr.size() > 1) memcpy(last->Tcw, fr.back().Tcw, sizeof(float) * 16);
  } else if (last) {
    frame_pose(last->Tcw);
  }
  if (last) {   // the tracker only keeps the last frame's feature lists alive
    for (auto& p : last->mvStat3DPointTmp) rot_pt(p);
    for (auto& p : last->mvObj3DPoint) rot_pt(p);
  }
  for (auto& f : map.vp3DPointSta) for (auto& p : f) rot_pt(p);
  for (auto& f : map.vp3DPointDyn) for (auto& p : f) rot_pt(p);
  for (auto& pose : map.vmCameraPose) scale_pose(pose.data());
  for (auto& pose : map.vmRigidMotion) scale_pose(pose.data());
  for (auto& f : map.vmObjMotion) for (auto& m : f) scale_pose(m.data());
}

}  // namespace

extern "C" {

void* vo_tracker_create(const vo_track_config* cfg) {
  Tracker* t = new Tracker();
  t->cfg = *cfg;
  return t;
}
void vo_tracker_destroy(void* h) { delete (Tracker*)h; }
int vo_tracker_track(void* h, const uint8_t* gray, float* depth, const float* flow, const int32_t* mask, float* Tcw_out,
                     vo_track_stats* st) {
  return ((Tracker*)h)->track(gray, depth, flow, mask, Tcw_out, st);
}
// ---- test hooks: the recorded gravity initialisations (n = frames 0..N); returns the number of frames of record k, or the count for k < 0
int vo_tracker_imu_init_log(void* h, int k, float* Tcw, int32_t* has_pre, float* dV, float* dT, float* Tbc, float* Rwg, float* vel_out) {
  Tracker* t = (Tracker*)h;
  if (k < 0) return (int)t->imu_init_log.size();
  const ImuInitRecord& r = t->imu_init_log[k];
  const int n = (int)r.has_pre.size();
  if (!Tcw) return n;
  auto cp = [](auto* dst, const auto& v) { for (size_t i = 0; i < v.size(); i++) dst[i] = v[i]; };
  cp(Tcw, r.Tcw); cp(has_pre, r.has_pre); cp(dV, r.dV); cp(dT, r.dT); cp(vel_out, r.vel_out);
  memcpy(Tbc, r.Tbc, sizeof r.Tbc); memcpy(Rwg, r.Rwg, sizeof r.Rwg);
  return n;
}

// ---- test hooks: the recorded UpdateFrameIMU calls, 46 floats each in the order of ImuUpdateRecord; returns the count
int vo_tracker_imu_update_log(void* h, float* out, int cap) {
  Tracker* t = (Tracker*)h;
  const int n = (int)t->imu_update_log.size();
  for (int i = 0; i < n && i < cap; i++) memcpy(out + 46 * (size_t)i, &t->imu_update_log[i], sizeof(float) * 46);
  return n;
}

// ---- test hooks: the recorded DynObjTracking calls
void vo_tracker_dyn_log_enable(void* h, int on) { ((Tracker*)h)->dyn_log_on = on != 0; }
int vo_tracker_dyn_log_count(void* h) { return (int)((Tracker*)h)->dyn_log.size(); }
int vo_tracker_dyn_log_sizes(void* h, int k, int32_t* sizes) {
  const DynRecord& r = ((Tracker*)h)->dyn_log[k];
  const int v[8] = {(int)r.sem.size(), (int)r.last_sem.size(), (int)r.last_mod.size(), (int)r.out_mod.size(), (int)r.out_ids.size(),
                    r.f_id, r.max_id_before, r.max_id_after};
  for (int i = 0; i < 8; i++) sizes[i] = v[i];
  return 0;
}
int vo_tracker_dyn_log_get(void* h, int k, int32_t* sem, int32_t* lab_before, int32_t* lab_after, int32_t* last_sem, float* key_xy,
                           float* depth, float* flow3, int32_t* last_sem_pos, int32_t* last_stat, int32_t* last_mod, int32_t* out_mod,
                           int32_t* out_sem_pos, int32_t* out_len, int32_t* out_ids) {
  const DynRecord& r = ((Tracker*)h)->dyn_log[k];
  auto cp = [](auto* dst, const auto& v) { for (size_t i = 0; i < v.size(); i++) dst[i] = v[i]; };
  cp(sem, r.sem); cp(lab_before, r.lab_before); cp(lab_after, r.lab_after); cp(last_sem, r.last_sem); cp(key_xy, r.key_xy);
  cp(depth, r.depth); cp(flow3, r.flow3); cp(last_sem_pos, r.last_sem_pos); cp(last_stat, r.last_stat); cp(last_mod, r.last_mod);
  cp(out_mod, r.out_mod); cp(out_sem_pos, r.out_sem_pos); cp(out_len, r.out_len); cp(out_ids, r.out_ids);
  return 0;
}

// static half of Tracking::RenewFrameInfo (src/Tracking.cc:2959-3110) on caller-supplied frame state (test hook for
// tests/test_renew_independent.py); outputs sized for max_track_bg + 1 features; returns the count
int vo_renew_static(const vo_track_config* cfg, const int32_t* TM_sta, int n_tm, const float* stat_keys, int n_stat, const vo_keypoint* kps,
                    int n_kps, const float* depth, const float* flow, const int32_t* mask, const float* Tcw, float* keys, float* corres,
                    float* flow_next, int32_t* inlier_id, float* out_depth, float* p3) {
  Tracker t;
  t.cfg = *cfg;
  t.cur = new Frame();
  t.last = t.cur;
  for (int i = 0; i < n_stat; i++) t.cur->mvStatKeys.push_back({stat_keys[2 * i], stat_keys[2 * i + 1]});
  t.cur->mvKeys.assign(kps, kps + n_kps);
  memcpy(t.cur->Tcw, Tcw, sizeof(float) * 16);
  t.renew(std::vector<int>(TM_sta, TM_sta + n_tm), depth, flow, mask);
  const Frame& c = *t.cur;
  const int n = (int)c.mvStatKeysTmp.size();
  for (int i = 0; i < n; i++) {
    keys[2 * i] = c.mvStatKeysTmp[i].x; keys[2 * i + 1] = c.mvStatKeysTmp[i].y;
    corres[2 * i] = c.mvCorres[i].x; corres[2 * i + 1] = c.mvCorres[i].y;
    flow_next[2 * i] = c.mvFlowNext[i].x; flow_next[2 * i + 1] = c.mvFlowNext[i].y;
    inlier_id[i] = c.nStaInlierID[i];
    out_depth[i] = c.mvStatDepthTmp[i];
    p3[3 * i] = c.mvStat3DPointTmp[i].x; p3[3 * i + 1] = c.mvStat3DPointTmp[i].y; p3[3 * i + 2] = c.mvStat3DPointTmp[i].z;
  }
  return n;
}

// object half of Tracking::RenewFrameInfo (src/Tracking.cc:3112-3289) on caller-supplied frame state (test hook for
// tests/test_renew_independent.py); inlier_ids = the objects' inlier lists back to back (inlier_len[o] each); returns the count
int vo_renew_objects(const vo_track_config* cfg, int n_feat, const float* obj_keys, const int32_t* obj_label, int n_obj,
                     const int32_t* inlier_len, const int32_t* inlier_ids, const int32_t* obj_stat, const int32_t* sem_pos,
                     const int32_t* mod_label, int n_tmp, const float* tmp_keys, const float* tmp_depth, const int32_t* tmp_sem,
                     const float* tmp_flow, const float* tmp_corres, const float* depth, const float* flow, const int32_t* mask,
                     const float* Tcw, int cap, float* keys, float* out_depth, float* corres, float* flow_next, int32_t* sem,
                     int32_t* inlier_id, int32_t* label, float* p3) {
  Tracker t;
  t.cfg = *cfg;
  t.cur = new Frame();
  t.last = t.cur;
  Frame& c = *t.cur;
  for (int i = 0; i < n_feat; i++) { c.mvObjKeys.push_back({obj_keys[2 * i], obj_keys[2 * i + 1]}); c.vObjLabel.push_back(obj_label[i]); }
  int at = 0;
  for (int o = 0; o < n_obj; o++) {
    c.vnObjInlierID.push_back(std::vector<int>(inlier_ids + at, inlier_ids + at + inlier_len[o]));
    at += inlier_len[o];
    c.vnObjID.push_back({});
    c.bObjStat.push_back((char)(obj_stat[o] != 0)); c.nSemPosition.push_back(sem_pos[o]); c.nModLabel.push_back(mod_label[o]);
  }
  for (int j = 0; j < n_tmp; j++) {
    t.tmpKeys.push_back({tmp_keys[2 * j], tmp_keys[2 * j + 1]}); t.tmpDepth.push_back(tmp_depth[j]); t.tmpSem.push_back(tmp_sem[j]);
    t.tmpFlow.push_back({tmp_flow[2 * j], tmp_flow[2 * j + 1]}); t.tmpCorres.push_back({tmp_corres[2 * j], tmp_corres[2 * j + 1]});
  }
  memcpy(c.Tcw, Tcw, sizeof(float) * 16);
  t.renew_objects(depth, flow, mask);
  const int n = (int)c.mvObjKeys.size();
  if (n > cap) return -n;
  for (int i = 0; i < n; i++) {
    keys[2 * i] = c.mvObjKeys[i].x; keys[2 * i + 1] = c.mvObjKeys[i].y; out_depth[i] = c.mvObjDepth[i];
    corres[2 * i] = c.mvObjCorres[i].x; corres[2 * i + 1] = c.mvObjCorres[i].y;
    flow_next[2 * i] = c.mvObjFlowNext[i].x; flow_next[2 * i + 1] = c.mvObjFlowNext[i].y;
    sem[i] = c.vSemObjLabel[i]; inlier_id[i] = c.nDynInlierID[i]; label[i] = c.vObjLabel[i];
    p3[3 * i] = c.mvObj3DPoint[i].x; p3[3 * i + 1] = c.mvObj3DPoint[i].y; p3[3 * i + 2] = c.mvObj3DPoint[i].z;
  }
  return n;
}

// Tracking::DynObjTracking on caller-supplied frame state (test hook: randomised inputs reach the branches the synthetic sequences
// never take -- objects on the image boundary, static / far / small objects, lost ids, ties of the majority vote)
int vo_dyn_obj_tracking(const vo_track_config* cfg, int n, const int32_t* sem, int32_t* lab, const float* key_xy, const float* depth,
                        const float* flow3, const int32_t* last_sem, int n_last_obj, const int32_t* last_sem_pos, const int32_t* last_stat,
                        const int32_t* last_mod, int f_id, int32_t* max_id, int32_t* out_mod, int32_t* out_sem_pos, int32_t* out_len,
                        int32_t* out_ids) {
  Tracker t;
  t.cfg = *cfg;
  t.cur = new Frame();
  t.last = new Frame();
  t.f_id = f_id;
  t.max_id = *max_id;
  for (int i = 0; i < n; i++) {
    t.cur->vSemObjLabel.push_back(sem[i]); t.cur->vObjLabel.push_back(lab[i]);
    t.cur->mvObjKeys.push_back({key_xy[2 * i], key_xy[2 * i + 1]});
    t.cur->mvObjDepth.push_back(depth[i]);
    t.cur->vFlow_3d.push_back({flow3[3 * i], flow3[3 * i + 1], flow3[3 * i + 2]});
    t.last->vSemObjLabel.push_back(last_sem[i]);
  }
  for (int k = 0; k < n_last_obj; k++) {
    t.last->nSemPosition.push_back(last_sem_pos[k]); t.last->bObjStat.push_back((char)(last_stat[k] != 0)); t.last->nModLabel.push_back(last_mod[k]);
  }
  const std::vector<std::vector<int>> ids = t.dyn_obj_tracking();
  for (int i = 0; i < n; i++) lab[i] = t.cur->vObjLabel[i];
  *max_id = t.max_id;
  int at = 0;
  for (size_t o = 0; o < ids.size(); o++) {
    out_mod[o] = t.cur->nModLabel[o]; out_sem_pos[o] = t.cur->nSemPosition[o]; out_len[o] = (int)ids[o].size();
    for (int id : ids[o]) out_ids[at++] = id;
  }
  const int no = (int)ids.size();
  if (t.cur != t.last) delete t.cur;
  delete t.last;
  t.cur = t.last = nullptr;
  return no;
}

int vo_tracker_num_frames(void* h) { return (int)((Tracker*)h)->map.vmCameraPose.size(); }
int vo_tracker_get_map_poses(void* h, float* poses, int cap) {
  Tracker* t = (Tracker*)h;
  int n = (int)t->map.vmCameraPose.size();
  for (int i = 0; i < n && i < cap; i++) memcpy(poses + 16 * i, t->map.vmCameraPose[i].data(), sizeof(float) * 16);
  return n;
}
int vo_tracker_get_static(void* h, int frame, float* xy, float* depth, float* p3, int32_t* asso, int cap) {
  Tracker* t = (Tracker*)h;
  if (frame < 0 || frame >= (int)t->map.vpFeatSta.size()) return -1;
  const int n = (int)t->map.vpFeatSta[frame].size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = t->map.vpFeatSta[frame][i].x; xy[2 * i + 1] = t->map.vpFeatSta[frame][i].y;
    depth[i] = t->map.vfDepSta[frame][i];
    p3[3 * i] = t->map.vp3DPointSta[frame][i].x; p3[3 * i + 1] = t->map.vp3DPointSta[frame][i].y; p3[3 * i + 2] = t->map.vp3DPointSta[frame][i].z;
    asso[i] = frame > 0 ? t->map.vnAssoSta[frame - 1][i] : -1;
  }
  return n;
}

int vo_tracker_get_dynamic(void* h, int frame, float* xy, float* depth, float* p3, int32_t* asso, int32_t* label, int cap) {
  Tracker* t = (Tracker*)h;
  if (frame < 0 || frame >= (int)t->map.vpFeatDyn.size()) return -1;
  const int n = (int)t->map.vpFeatDyn[frame].size();
  for (int i = 0; i < n && i < cap; i++) {
    xy[2 * i] = t->map.vpFeatDyn[frame][i].x; xy[2 * i + 1] = t->map.vpFeatDyn[frame][i].y;
    depth[i] = t->map.vfDepDyn[frame][i];
    p3[3 * i] = t->map.vp3DPointDyn[frame][i].x; p3[3 * i + 1] = t->map.vp3DPointDyn[frame][i].y; p3[3 * i + 2] = t->map.vp3DPointDyn[frame][i].z;
    asso[i] = frame > 0 ? t->map.vnAssoDyn[frame - 1][i] : -1;
    label[i] = frame > 0 ? t->map.vnFeatLabel[frame - 1][i] : -2;
  }
  return n;
}
int vo_tracker_get_objects(void* h, int frame, int32_t* label, int32_t* sem_label, float* motion, float* centre, int cap) {
  Tracker* t = (Tracker*)h;
  if (frame < 1 || frame > (int)t->map.vmObjMotion.size()) return -1;
  const auto& M = t->map.vmObjMotion[frame - 1];
  const int n = (int)M.size();
  for (int i = 0; i < n && i < cap; i++) {
    label[i] = t->map.vnRMLabel[frame - 1][i]; sem_label[i] = t->map.vnSMLabel[frame - 1][i];
    memcpy(motion + 16 * i, M[i].data(), sizeof(float) * 16);
    centre[3 * i] = t->map.vmRigidCentre[frame - 1][i].x; centre[3 * i + 1] = t->map.vmRigidCentre[frame - 1][i].y;
    centre[3 * i + 2] = t->map.vmRigidCentre[frame - 1][i].z;
  }
  return n;
}
int vo_tracker_get_dyn_tracks(void* h, int32_t* len, int32_t* obj_id, int32_t* first_frame, int32_t* first_feat, int cap) {
  Tracker* t = (Tracker*)h;
  const int n = (int)t->map.TrackletDyn.size();
  for (int i = 0; i < n && i < cap; i++) {
    len[i] = (int)t->map.TrackletDyn[i].size(); obj_id[i] = t->map.nObjID[i];
    first_frame[i] = t->map.TrackletDyn[i][0].first; first_feat[i] = t->map.TrackletDyn[i][0].second;
  }
  return n;
}

int vo_tracker_apply_scaled_rotation(void* h, const float* R, float s) { ((Tracker*)h)->apply_scaled_rotation(R, s); return 0; }
int vo_tracker_set_imu(void* h, const float* Tbc, const float* noise) { ((Tracker*)h)->set_imu(Tbc, noise); return 0; }
int vo_tracker_grab_imu(void* h, const vo_imu_sample* smp, int n) {
  Tracker* t = (Tracker*)h;
  t->imu_q.insert(t->imu_q.end(), smp, smp + n);
  return 0;
}
void vo_tracker_set_timestamp(void* h, double ts) { ((Tracker*)h)->next_t = ts; }
int vo_tracker_get_imu_state(void* h, vo_imu_state* out) { *out = ((Tracker*)h)->ist; return 0; }
int vo_tracker_get_imu_frames(void* h, float* Tcw, float* vel, float* bias, int cap) {
  Tracker* t = (Tracker*)h;
  const int n = (int)t->fr.size();
  for (int i = 0; i < n && i < cap; i++) {
    if (Tcw) memcpy(Tcw + 16 * (size_t)i, t->fr[i].Tcw, sizeof(float) * 16);
    if (vel) memcpy(vel + 3 * (size_t)i, t->fr[i].vel, sizeof(float) * 3);
    if (bias) memcpy(bias + 6 * (size_t)i, t->fr[i].bias, sizeof(float) * 6);
  }
  return n;
}

int vo_tracker_full_batch(void* h, vo_lm_stats* stats, int32_t* sizes) { return ((Tracker*)h)->full_batch(stats, sizes); }
int vo_tracker_get_map_poses_rf(void* h, float* poses, int cap) {
  Tracker* t = (Tracker*)h;
  const int n = (int)t->vmCameraPose_RF.size();
  for (int i = 0; i < n && i < cap; i++) memcpy(poses + 16 * i, t->vmCameraPose_RF[i].data(), sizeof(float) * 16);
  return n;
}
int vo_tracker_get_objects_rf(void* h, int frame, float* motion, int cap) {
  Tracker* t = (Tracker*)h;
  if (frame < 1 || frame > (int)t->vmObjMotion_RF.size()) return -1;
  const auto& M = t->vmObjMotion_RF[frame - 1];
  for (int i = 0; i < (int)M.size() && i < cap; i++) memcpy(motion + 16 * i, M[i].data(), sizeof(float) * 16);
  return (int)M.size();
}
int vo_tracker_export_full_graph(void* h, int32_t* sizes, float* se3, float* points, int32_t* e6_i, int32_t* e6_j, int32_t* e6_kind,
                                 float* e6_meas, int32_t* obs_se3, int32_t* obs_point, int32_t* obs_kind, float* obs_xyz,
                                 int32_t* tern_p1, int32_t* tern_p2, int32_t* tern_h) {
  Tracker* t = (Tracker*)h;
  if (t->cfg.rebuild_tracklets == 0) { t->rebuild_tracklets(); t->rebuild_dyn_tracklets(); }
  auto G = t->build_full_graph();
  sizes[0] = G.n_poses; sizes[1] = G.n_motions; sizes[2] = (int)(G.points.size() / 3); sizes[3] = (int)G.obs_se3.size();
  sizes[4] = (int)G.e6_i.size(); sizes[5] = (int)G.tern_p1.size();
  if (!se3) return 0;
  auto cp = [](auto* dst, const auto& v) { if (!v.empty()) memcpy(dst, v.data(), sizeof(v[0]) * v.size()); };
  cp(se3, G.se3); cp(points, G.points); cp(e6_i, G.e6_i); cp(e6_j, G.e6_j); cp(e6_kind, G.e6_kind); cp(e6_meas, G.e6_meas);
  cp(obs_se3, G.obs_se3); cp(obs_point, G.obs_point); cp(obs_kind, G.obs_kind); cp(obs_xyz, G.obs_xyz);
  cp(tern_p1, G.tern_p1); cp(tern_p2, G.tern_p2); cp(tern_h, G.tern_h);
  return 0;
}

// Tracking::GetMetricError (src/Tracking.cc:3531-3674), bRMSError = false
static void metric_pair(const float* E, float* t_err, float* r_err) {
  *t_err = std::sqrt(E[3] * E[3] + E[7] * E[7] + E[11] * E[11]);
  float tr = 0;
  for (int j = 0; j < 3; j++) {
    const float d = E[5 * j];
    if (d > 1.0) tr = (float)((double)tr + 1.0 - ((double)d - 1.0));
    else tr = tr + d;
  }
  *r_err = (float)(std::acos(((double)tr - 1.0) / 2.0) * 180.0 / 3.1415926);
}
int vo_tracker_metric_error(void* h, const float* cam_gt, int n_gt, int refined, const float* pose_pre, const float* mot_gt, int n_obj,
                            vo_metric* out, float* per_item) {
  Tracker* t = (Tracker*)h;
  const auto& cam = (refined && !t->vmCameraPose_RF.empty()) ? t->vmCameraPose_RF : t->map.vmCameraPose;
  const int n = (int)cam.size();
  if (n_gt < n) return -1;
  memset(out, 0, sizeof *out);
  float ts = 0, rs = 0;
  int k = 0;
  for (int i = 1; i < n; i++) {
    float A[16], B[16], C[16], E[16], te, re;
    inv44(cam[i - 1].data(), A);
    mul44(cam[i].data(), A, B);
    inv44(cam_gt + 16 * (size_t)i, A);
    mul44(cam_gt + 16 * (size_t)(i - 1), A, C);
    mul44(B, C, E);
    metric_pair(E, &te, &re);
    ts = ts + te; rs = rs + re;
    if (per_item) { per_item[2 * k] = te; per_item[2 * k + 1] = re; }
    k++;
  }
  out->n_cam = n > 1 ? n - 1 : 0;
  if (out->n_cam) { out->cam_t = ts / (float)out->n_cam; out->cam_r = rs / (float)out->n_cam; }
  ts = 0; rs = 0;
  int o = 0;
  if (n_obj > 0) {
    for (size_t f = 0; f < t->map.vmObjMotion.size(); f++)
      for (size_t j = 0; j < t->map.vmObjMotion[f].size(); j++) {
        if (o >= n_obj) return -2;
        const float* M = (refined && !t->vmObjMotion_RF.empty()) ? t->vmObjMotion_RF[f][j].data() : t->map.vmObjMotion[f][j].data();
        const float* P = pose_pre + 16 * (size_t)o;
        float A[16], B[16], C[16], E[16], te, re;
        inv44(P, A);
        mul44(A, M, B);
        mul44(B, P, C);
        inv44(C, A);
        mul44(A, mot_gt + 16 * (size_t)o, E);
        metric_pair(E, &te, &re);
        ts = ts + te; rs = rs + re;
        if (per_item) { per_item[2 * k] = te; per_item[2 * k + 1] = re; }
        k++; o++;
      }
    if (o != n_obj) return -2;
    out->n_obj = n_obj;
    out->obj_t = ts / (float)n_obj; out->obj_r = rs / (float)n_obj;
  }
  return 0;
}

}  // extern "C"
