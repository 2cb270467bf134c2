// System.cc -- out-of-line members of the VIDO_SLAM::System facade (host/System.h), built into libvido_slam.so.
// Reference: vido_slam/src/System.cc:23-233 (Init, the two TrackRGBD overloads, SaveResultsIJRR2020 with its timing table).
#include "System.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>

namespace VIDO_SLAM {

static inline int mat_channels(const cv::Mat& m) { return m.channels(); }
static inline cv::Mat make_pose_mat(const float* T) {
#ifdef VIDO_HAVE_OPENCV
  cv::Mat m(4, 4, CV_32F);
#else
  cv::Mat m = cv::Mat::create(4, 4, CV_32FC1);
#endif
  memcpy(m.data, T, sizeof(float) * 16);
  return m;
}

System::~System() { if (ctx_) vido_destroy(ctx_); }

bool System::isImuInitialized() { vido_imu_state st; return vido_track_get_imu_state(ctx_, &st) == VIDO_OK && st.initialized; }

void System::Init(const std::string& strSettingsFile, const eSensor sensor) {
  sensor_ = sensor;
  std::ifstream f(strSettingsFile.c_str());
  if (!f.is_open()) {  // src/System.cc:33-37
    std::cerr << "Failed to open settings file at: " << strSettingsFile << std::endl;
    exit(-1);
  }
  std::map<std::string, std::string> kv;
  std::string line;
  std::vector<float> tbc;   // "Tbc: !!opencv-matrix ... data: [ 16 values ]" (Tracking::ParseIMUParamFile, src/Tracking.cc:174-196)
  int in_tbc = 0;           // 1: inside the Tbc node, 2: inside its data list
  while (std::getline(f, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line = line.substr(0, h);
    if (line.compare(0, 4, "Tbc:") == 0) { in_tbc = 1; continue; }
    if (in_tbc) {
      size_t pos = 0;
      if (in_tbc == 1) {
        const size_t d = line.find("data:");
        if (d == std::string::npos) { if (!line.empty() && line[0] != ' ') in_tbc = 0; else continue; }
        else { in_tbc = 2; pos = line.find('[', d); pos = (pos == std::string::npos) ? line.size() : pos + 1; }
      }
      if (in_tbc == 2) {
        std::string body = line.substr(pos);
        const size_t e = body.find(']');
        const bool last = e != std::string::npos;
        if (last) body = body.substr(0, e);
        for (char& ch : body) if (ch == ',') ch = ' ';
        std::istringstream is(body);
        float v;
        while (is >> v) tbc.push_back(v);
        if (last) in_tbc = 0;
        continue;
      }
    }
    const size_t c = line.find(':');
    if (c == std::string::npos || line[0] == '%') continue;
    std::string k = line.substr(0, c), v = line.substr(c + 1);
    auto trim = [](std::string& s) { size_t a = s.find_first_not_of(" \t\r\""), b = s.find_last_not_of(" \t\r\""); s = (a == std::string::npos) ? "" : s.substr(a, b - a + 1); };
    trim(k); trim(v);
    if (!k.empty() && !v.empty()) kv[k] = v;
  }
  auto num = [&](const char* k, double dflt) { auto it = kv.find(k); return it == kv.end() ? dflt : atof(it->second.c_str()); };
  vido_config c;
  vido_default_config(&c);
  c.width = (int)num("Camera.width", c.width); c.height = (int)num("Camera.height", c.height);
  c.fx = (float)num("Camera.fx", c.fx); c.fy = (float)num("Camera.fy", c.fy);
  c.cx = (float)num("Camera.cx", c.cx); c.cy = (float)num("Camera.cy", c.cy); c.bf = (float)num("Camera.bf", c.bf);
  c.rgb = (int)num("Camera.RGB", c.rgb);
  c.choose_data = (int)num("ChooseData", c.choose_data);
  c.depth_map_factor = (float)num("DepthMapFactor", c.depth_map_factor);
  c.th_depth_bg = (float)num("ThDepthBG", c.th_depth_bg); c.th_depth_obj = (float)num("ThDepthOBJ", c.th_depth_obj);
  c.max_track_bg = (int)num("MaxTrackPointBG", c.max_track_bg); c.max_track_obj = (int)num("MaxTrackPointOBJ", c.max_track_obj);
  c.window_size = (int)num("WINDOW_SIZE", c.window_size);
  c.nfeatures = (int)num("ORBextractor.nFeatures", c.nfeatures);
  c.scale_factor = (float)num("ORBextractor.scaleFactor", c.scale_factor);
  c.nlevels = (int)num("ORBextractor.nLevels", c.nlevels);
  c.ini_th_fast = (int)num("ORBextractor.iniThFAST", c.ini_th_fast); c.min_th_fast = (int)num("ORBextractor.minThFAST", c.min_th_fast);
  c.sf_mg_thres = (float)num("SFMgThres", c.sf_mg_thres); c.sf_ds_thres = (float)num("SFDsThres", c.sf_ds_thres);
  imu_noise_[0] = (float)num("IMU.NoiseGyro", 1.7e-4); imu_noise_[1] = (float)num("IMU.NoiseAcc", 2.0e-3);
  imu_noise_[2] = (float)num("IMU.GyroWalk", 1.9393e-05); imu_noise_[3] = (float)num("IMU.AccWalk", 3.0e-03);
  const float freq = (float)num("IMU.Frequency", 200.0), sf = sqrtf(freq);  // Tracking::ParseIMUParamFile (Tracking.cc:174-275)
  imu_noise_[0] *= sf; imu_noise_[1] *= sf; imu_noise_[2] /= sf; imu_noise_[3] /= sf;
  if ((int)num("UseSampleFeature", 0) != 0) std::cerr << "vido_b200: UseSampleFeature=1 is time-seeded in the reference; detected features are used" << std::endl;
  c.max_batch = 1;  // frame-by-frame facade; use vido_track_frames with chunks for throughput
  cfg_ = c;
  ctx_ = vido_create(&c);
  if (!ctx_) {
    std::cerr << "vido_b200: " << vido_last_error(nullptr) << std::endl;
    exit(-1);
  }
  if (sensor == IMU_RGBD) {   // src/Tracking.cc:111-121
    if (tbc.size() != 16) {
      std::cerr << "*Tbc matrix have to be a 4x4 transformation matrix*" << std::endl;
      std::cout << "*Error with the IMU parameters in the config file*" << std::endl;
    } else if (vido_track_set_imu(ctx_, tbc.data(), imu_noise_) < 0) {
      std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;
    }
  }
}

// one frame through the driver (Tracking::GrabImageRGBD, src/Tracking.cc:283-456)
cv::Mat System::track(const cv::Mat& im, cv::Mat& depthmap, const cv::Mat& flowmap, const cv::Mat& masksem, const cv::Mat& mTcw_gt,
                      const double& timestamp, const int& nImage) {
  (void)mTcw_gt;   // see after_track
  vido_frame_inputs in;
  memset(&in, 0, sizeof in);
  in.image = im.data; in.channels = mat_channels(im); in.on_device = 0;
  in.depth = (const float*)depthmap.data; in.flow = (const float*)flowmap.data; in.mask = (const int32_t*)masksem.data;
  in.write_back_depth = 1;
  in.timestamp = timestamp;
  float Tcw[16];
  vido_track_stats st;
  const int rc = vido_track_frames(ctx_, &in, 1, Tcw, &st);
  return after_track(rc, Tcw, st, timestamp, nImage);
}

// the demo's per-frame sequence (demo/run_vido_slam.cc:112-137) with the pixel conversions on the device
cv::Mat System::TrackRaw(const uint8_t* bayer, const uint16_t* depth16, const float* flow, const uint8_t* mask8,
                         const std::vector<IMU::Point>* vImuMeas, const double& timestamp, const int& nImage) {
  if ((sensor_ == IMU_RGBD) != (vImuMeas != nullptr)) {
    std::cerr << "ERROR: TrackRaw needs IMU measurements exactly when the sensor is IMU_RGBD." << std::endl;
    exit(-1);
  }
  if (vImuMeas) {
    std::vector<vido_imu_sample> q(vImuMeas->size());
    for (size_t i = 0; i < q.size(); i++) {
      const IMU::Point& m = (*vImuMeas)[i];
      q[i].t = m.t; q[i].ax = m.ax; q[i].ay = m.ay; q[i].az = m.az; q[i].wx = m.wx; q[i].wy = m.wy; q[i].wz = m.wz;
    }
    if (vido_track_grab_imu(ctx_, q.data(), (int)q.size(), 0) < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;
  }
  vido_raw_inputs in;
  in.bayer = bayer; in.depth16 = depth16; in.flow = flow; in.mask8 = mask8; in.timestamp = timestamp;
  float Tcw[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  vido_track_stats st;
  memset(&st, 0, sizeof st);
  const int rc = vido_track_raw_frames(ctx_, &in, 1, Tcw, &st);
  return after_track(rc, Tcw, st, timestamp, nImage);
}

cv::Mat System::after_track(int rc, const float* Tcw, const vido_track_stats& st, double timestamp, int nImage) {
  if (rc < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;  // the reference prints and continues
  else {
    // the reference's timing table (src/Tracking.cc:348-359,1121-1139,1177-1330,1451): mask update, camera pose estimation,
    // object tracking, object motion estimation, map update; local bundle adjustment.  The driver reports host wall time of
    // its stages: init model + pose optimisation = camera pose; renewal (incl. the object part) = map update.
    const double row[6] = {0.0, st.ms_init + st.ms_poseopt, 0.0, 0.0, st.ms_renew, st.ms_ba};
    if (frame_id_ > 0) stage_ms_.insert(stage_ms_.end(), row, row + 6);
  }
  trajectory_.insert(trajectory_.end(), Tcw, Tcw + 16);
  // Map::vmCameraPose_GT: the reference only ever stores the identity of the first frame (src/Tracking.cc:1546; the per-frame
  // assignment at :436-444 is commented out), so does this facade -- mTcw_gt is accepted and otherwise unused, like there
  if (frame_id_ == 0) { const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; cam_gt_.assign(I, I + 16); }
  last_t_ = timestamp;
  // f_id == StopFrame (= nImage - 1): the joint optimisation over the whole sequence, KITTI-style data only
  // (src/Tracking.cc:288, 1490-1498); results are read by SaveResultsIJRR2020
  if (frame_id_ == nImage - 1 && cfg_.choose_data == 2) {
    vido_lm_stats ls;
    if (vido_full_batch(ctx_, &ls, nullptr) < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;
    else full_batch_done_ = true;
  }
  frame_id_++;
  return make_pose_mat(Tcw);
}

cv::Mat System::TrackRGBD(const cv::Mat& im, cv::Mat& depthmap, const cv::Mat& flowmap, const cv::Mat& masksem, const cv::Mat& mTcw_gt,
                          const std::vector<std::vector<float> >& /*vObjPose_gt*/, const double& timestamp, cv::Mat& /*imTraj*/,
                          const int& nImage) {
  if (sensor_ != RGBD) {  // src/System.cc:55-59
    std::cerr << "ERROR: you called TrackRGBD but input sensor was not set to RGBD." << std::endl;
    exit(-1);
  }
  return track(im, depthmap, flowmap, masksem, mTcw_gt, timestamp, nImage);
}

cv::Mat System::TrackRGBD(const cv::Mat& im, cv::Mat& depthmap, const cv::Mat& flowmap, const cv::Mat& masksem,
                          const std::vector<IMU::Point>& vImuMeas, const cv::Mat& mTcw_gt,
                          const std::vector<std::vector<float> >& /*vObjPose_gt*/, const double& timestamp, cv::Mat& /*imTraj*/,
                          const int& nImage) {
  if (sensor_ != IMU_RGBD) {  // src/System.cc:69-73
    std::cerr << "ERROR: you called TrackRGBD_VIO but input sensor was not set to IMU_RGBD." << std::endl;
    exit(-1);
  }
  std::vector<vido_imu_sample> q(vImuMeas.size());
  for (size_t i = 0; i < vImuMeas.size(); i++) {   // Tracking::GrabImuData per measurement (src/System.cc:74-75)
    vido_imu_sample& s = q[i];
    s.t = vImuMeas[i].t; s.ax = vImuMeas[i].ax; s.ay = vImuMeas[i].ay; s.az = vImuMeas[i].az;
    s.wx = vImuMeas[i].wx; s.wy = vImuMeas[i].wy; s.wz = vImuMeas[i].wz;
  }
  if (vido_track_grab_imu(ctx_, q.data(), (int)q.size(), 0) < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;
  return track(im, depthmap, flowmap, masksem, mTcw_gt, timestamp, nImage);
}

// Files and layout of src/System.cc:80-198 ("frame [label] m00 .. m23 0 0 0 1", fixed, 9 decimals):
//   obj_mot_rgbd_new.txt  Map::vmRigidMotion[f][j >= 1] with Map::vnRMLabel;   obj_mot_gt.txt  Map::vmRigidMotion_GT -- never
//   filled by the reference (it would index an empty vector when an object motion exists): created empty here;
//   initial_rgbd_new.txt  Map::vmCameraPose (Twc, refined by the window optimisation);   refined_rgbd_new.txt
//   Map::vmCameraPose_RF (after FullBatchOptimization; equal to the initial one if it never ran);   cam_pose_gt.txt
//   Map::vmCameraPose_GT (the identity of the first frame, see track()).  Then the timing table of :200-233 on stdout.
void System::SaveResultsIJRR2020(const std::string& filename) {
  std::cout << std::endl << "Saving Results into TXT File..." << std::endl;
  auto row16 = [](std::ofstream& f, const float* M) {
    f << std::fixed << std::setprecision(9);
    for (int k = 0; k < 12; k++) f << M[k] << " ";
    f << 0.0 << " " << 0.0 << " " << 0.0 << " " << 1.0 << std::endl;
  };
  const int n = vido_map_num_frames(ctx_);
  {
    std::ofstream f((filename + "obj_mot_rgbd_new.txt").c_str(), std::ios::trunc);
    std::ofstream g((filename + "obj_mot_gt.txt").c_str(), std::ios::trunc);
    for (int i = 1; i < n; i++) {
      int32_t label[64], sem[64];
      float motion[64 * 16], centre[64 * 3];
      const int m = vido_map_get_objects(ctx_, i, label, sem, motion, centre, 64);
      for (int j = 0; j < m && j < 64; j++) { f << i << " " << label[j] << " "; row16(f, motion + 16 * j); }
    }
  }
  std::vector<float> P(16 * (size_t)(n > 0 ? n : 1));
  {
    std::ofstream f((filename + "initial_rgbd_new.txt").c_str(), std::ios::trunc);
    if (n > 0) vido_map_get_poses(ctx_, P.data(), n);
    for (int i = 0; i < n; i++) { f << i << " "; row16(f, P.data() + 16 * (size_t)i); }
  }
  {
    std::ofstream f((filename + "refined_rgbd_new.txt").c_str(), std::ios::trunc);
    if (n > 0) vido_map_get_poses_rf(ctx_, P.data(), n);
    for (int i = 0; i < n; i++) { f << i << " "; row16(f, P.data() + 16 * (size_t)i); }
  }
  {
    std::ofstream f((filename + "cam_pose_gt.txt").c_str(), std::ios::trunc);
    for (size_t i = 0; i + 16 <= cam_gt_.size(); i += 16) { f << i / 16 << " "; row16(f, cam_gt_.data() + i); }
  }
  // ---- time analysis (src/System.cc:200-233): component 3 is averaged over the frames in which it is non-zero
  const size_t nf = stage_ms_.size() / 6;
  double avg[6] = {0, 0, 0, 0, 0, 0};
  size_t obj_cnt = 0;
  for (size_t i = 0; i < nf; i++)
    for (int j = 0; j < 6; j++) {
      avg[j] += stage_ms_[6 * i + j];
      if (j == 3 && stage_ms_[6 * i + j] != 0) obj_cnt++;
    }
  std::cout << "Time of all components: " << std::endl;
  for (int j = 0; j < 5; j++) {
    const double d = (j == 3) ? (double)obj_cnt : (double)nf;
    std::cout << "(" << j << "): " << (d > 0 ? avg[j] / d : 0.0) << " ";
  }
  std::cout << std::endl;
  std::cout << "Time of local bundle adjustment: " << (nf ? avg[5] / nf : 0.0) << std::endl;
}

}  // namespace VIDO_SLAM
