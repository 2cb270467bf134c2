// System.h -- C++ host facade with the reference's call surface (vido_slam/include/System.h:72-114) on top of the
// C-ABI of libvido_b200.so.  Header-only.  Same names, argument meaning and error behaviour:
//   VIDO_SLAM::System::Init(yaml, sensor)                         src/System.cc:23-48
//   cv::Mat System::TrackRGBD(im, depth, flow, mask, Tcw_gt, objPose_gt, t, imTraj, nImage)   src/System.cc:51-63
//   cv::Mat System::TrackRGBD(..., vImuMeas, ...)                 src/System.cc:65-78: GrabImuData for every measurement, then the
//                                                                 frame; sensor IMU_RGBD runs the VIO mode of the driver
//                                                                 (preintegration, InitializeIMU, ScaleRefinement)
//   void System::SaveResultsIJRR2020(dir)                         src/System.cc:80-198 (object motions, camera trajectory after
//                                                                 the window optimisation and after FullBatch)
// The call with index nImage-1 also runs Optimizer::FullBatchOptimization when ChooseData == 2 (src/Tracking.cc:1490-1498).
// With OpenCV available define VIDO_HAVE_OPENCV before including: the signatures then use cv::Mat exactly like the
// reference.  Without OpenCV (this image has no OpenCV C++), VIDO_SLAM::Mat below is a minimal view type with the same
// fields the path touches (rows, cols, type, data, step).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/vido_b200.h"

#ifdef VIDO_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#endif

namespace VIDO_SLAM {

#ifdef VIDO_HAVE_OPENCV
typedef cv::Mat Mat;
inline int mat_channels(const Mat& m) { return m.channels(); }
inline Mat make_pose_mat(const float* T) { Mat m(4, 4, CV_32F); memcpy(m.data, T, sizeof(float) * 16); return m; }
#else
enum { CV_8U = 0, CV_32S = 4, CV_32F = 5 };
#define VIDO_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
enum { CV_8UC1 = VIDO_MAKETYPE(CV_8U, 1), CV_8UC3 = VIDO_MAKETYPE(CV_8U, 3), CV_32FC1 = VIDO_MAKETYPE(CV_32F, 1),
       CV_32FC2 = VIDO_MAKETYPE(CV_32F, 2), CV_32SC1 = VIDO_MAKETYPE(CV_32S, 1) };
struct Mat {  // non-owning view unless created through create()
  int rows = 0, cols = 0, flags = 0;
  unsigned char* data = nullptr;
  size_t step = 0;
  std::vector<unsigned char> storage;
  Mat() {}
  Mat(int r, int c, int type, void* d, size_t st = 0) : rows(r), cols(c), flags(type), data((unsigned char*)d) {
    step = st ? st : (size_t)c * elemSize();
  }
  static Mat create(int r, int c, int type) {
    Mat m;
    m.rows = r; m.cols = c; m.flags = type;
    m.step = (size_t)c * m.elemSize();
    m.storage.assign(m.step * r, 0);
    m.data = m.storage.data();
    return m;
  }
  int type() const { return flags; }
  int channels() const { return (flags >> 3) + 1; }
  size_t elemSize() const { const int d = flags & 7; return (size_t)channels() * (d == CV_8U ? 1 : 4); }
  bool empty() const { return data == nullptr; }
  template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
  template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
};
inline int mat_channels(const Mat& m) { return m.channels(); }
inline Mat make_pose_mat(const float* T) { Mat m = Mat::create(4, 4, CV_32FC1); memcpy(m.data, T, sizeof(float) * 16); return m; }
#endif

namespace IMU {
struct Point {  // include/ImuTypes.h:32-46
  float ax, ay, az, wx, wy, wz;
  double t;
  Point(float ax_, float ay_, float az_, float wx_, float wy_, float wz_, double t_) : ax(ax_), ay(ay_), az(az_), wx(wx_), wy(wy_), wz(wz_), t(t_) {}
};
}  // namespace IMU

class System {
 public:
  enum eSensor { MONOCULAR = 0, STEREO = 1, RGBD = 2, IMU_RGBD = 3 };

  System() {}
  ~System() { if (ctx_) vido_destroy(ctx_); }

  // Reads the OpenCV-FileStorage style YAML the reference uses (src/config/*.yaml) -- flat "Key: value" lines.
  void Init(const std::string& strSettingsFile, const eSensor sensor) {
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

  // depthmap is modified in place (pre-scaled) exactly like the reference (src/Tracking.cc:299-322)
  Mat TrackRGBD(const Mat& im, Mat& depthmap, const Mat& flowmap, const Mat& maskmap, const Mat& /*mTcw_gt*/,
                const std::vector<std::vector<float> >& /*vObjPose_gt*/, const double& timestamp, Mat& /*imTraj*/,
                const int& nImage) {
    if (sensor_ != RGBD && sensor_ != IMU_RGBD) {  // src/System.cc:55-59
      std::cerr << "ERROR: you called TrackRGBD but input sensor was not set to RGBD." << std::endl;
      exit(-1);
    }
    vido_frame_inputs in;
    memset(&in, 0, sizeof in);
    in.image = im.data; in.channels = mat_channels(im); in.on_device = 0;
    in.depth = (const float*)depthmap.data; in.flow = (const float*)flowmap.data; in.mask = (const int32_t*)maskmap.data;
    in.write_back_depth = 1;
    in.timestamp = timestamp;
    float Tcw[16];
    const int rc = vido_track_frames(ctx_, &in, 1, Tcw, nullptr);
    if (rc < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;  // the reference prints and continues
    trajectory_.insert(trajectory_.end(), Tcw, Tcw + 16);
    last_t_ = timestamp;
    // f_id == StopFrame (= nImage - 1): the joint optimisation over the whole sequence, KITTI-style data only
    // (src/Tracking.cc:288, 1490-1498); results are read by SaveResultsIJRR2020
    if (frame_id_ == nImage - 1 && cfg_.choose_data == 2) {
      vido_lm_stats st;
      if (vido_full_batch(ctx_, &st, nullptr) < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;
      else full_batch_done_ = true;
    }
    frame_id_++;
    return make_pose_mat(Tcw);
  }

  Mat TrackRGBD(const Mat& im, Mat& depthmap, const Mat& flowmap, const Mat& maskmap, const std::vector<IMU::Point>& vImuMeas,
                const Mat& mTcw_gt, const std::vector<std::vector<float> >& vObjPose_gt, const double& timestamp, Mat& imTraj,
                const int& nImage) {
    if (sensor_ != IMU_RGBD) {  // src/System.cc:69-73
      std::cerr << "ERROR: you called TrackRGBD(IMU) but input sensor was not set to IMU_RGBD." << std::endl;
      exit(-1);
    }
    std::vector<vido_imu_sample> q(vImuMeas.size());
    for (size_t i = 0; i < vImuMeas.size(); i++) {
      vido_imu_sample& s = q[i];
      s.t = vImuMeas[i].t; s.ax = vImuMeas[i].ax; s.ay = vImuMeas[i].ay; s.az = vImuMeas[i].az;
      s.wx = vImuMeas[i].wx; s.wy = vImuMeas[i].wy; s.wz = vImuMeas[i].wz;
    }
    if (vido_track_grab_imu(ctx_, q.data(), (int)q.size(), 0) < 0) std::cerr << "vido_b200: " << vido_last_error(ctx_) << std::endl;
    return TrackRGBD(im, depthmap, flowmap, maskmap, mTcw_gt, vObjPose_gt, timestamp, imTraj, nImage);
  }

  // obj_mot_rgbd_new.txt: "frame label m00 .. m33" per estimated object motion (Map::vmRigidMotion[f][j>=1]);
  // initial_rgbd_new.txt: "frame m00 .. m33" of Map::vmCameraPose (Twc, refined by the window optimisation);
  // refined_rgbd_new.txt: Map::vmCameraPose_RF (after FullBatchOptimization; equal to the initial one if it never ran).
  // Layout and precision of src/System.cc:80-160; the ground-truth files of the reference are not written.
  void SaveResultsIJRR2020(const std::string& filename) {
    auto row16 = [](std::ofstream& f, const float* M) {
      f << std::fixed << std::setprecision(9);
      for (int k = 0; k < 12; k++) f << M[k] << " ";
      f << 0.0 << " " << 0.0 << " " << 0.0 << " " << 1.0 << std::endl;
    };
    const int n = vido_map_num_frames(ctx_);
    {
      std::ofstream f((filename + "obj_mot_rgbd_new.txt").c_str(), std::ios::trunc);
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
  }

  vido_ctx* context() { return ctx_; }
  // Tracking::isImuInitialized / mScale
  bool isImuInitialized() { vido_imu_state st; return vido_track_get_imu_state(ctx_, &st) == VIDO_OK && st.initialized; }

 private:
  vido_ctx* ctx_ = nullptr;
  vido_config cfg_;
  eSensor sensor_ = RGBD;
  std::vector<float> trajectory_;
  float imu_noise_[4] = {0, 0, 0, 0};
  double last_t_ = 0;
  bool full_batch_done_ = false;
  int frame_id_ = 0;
};

}  // namespace VIDO_SLAM
