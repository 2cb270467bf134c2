// System.h -- C++ host facade with the reference's call surface (vido_slam/include/System.h:72-114) on top of the
// C-ABI of libvido_b200.so.  The members are defined out of line in host/System.cc and built into libvido_slam.so (the name of
// the reference's library, vido_slam/CMakeLists.txt:70-72), so that a caller links -lvido_slam exactly as before.  Same names,
// argument meaning and error behaviour:
//   VIDO_SLAM::System::Init(yaml, sensor)                         src/System.cc:23-48
//   cv::Mat System::TrackRGBD(im, depth, flow, mask, Tcw_gt, objPose_gt, t, imTraj, nImage)   src/System.cc:51-63 (sensor must
//                                                                 be RGBD: IMU_RGBD is rejected here like the reference does)
//   cv::Mat System::TrackRGBD(..., vImuMeas, ...)                 src/System.cc:65-78: GrabImuData for every measurement, then the
//                                                                 frame; sensor IMU_RGBD runs the VIO mode of the driver
//                                                                 (preintegration, InitializeIMU, ScaleRefinement)
//   void System::SaveResultsIJRR2020(dir)                         src/System.cc:80-233 (object motions, camera trajectory after
//                                                                 the window optimisation and after FullBatch, the two ground-
//                                                                 truth files, the timing table)
// The call with index nImage-1 also runs Optimizer::FullBatchOptimization when ChooseData == 2 (src/Tracking.cc:1490-1498).
// With OpenCV available define VIDO_HAVE_OPENCV before including: the signatures then use the real cv::Mat.  Without OpenCV
// (this image has no OpenCV C++) a stand-in class of the same NAME (cv::Mat) with the fields the path touches (rows, cols,
// type, data, step) is declared below: the mangled names of the members equal the reference's, the object layout does not --
// a binary compiled against the real OpenCV needs the shim built with -DVIDO_HAVE_OPENCV.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vido_b200.h"

#ifdef VIDO_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#else
namespace cv {
class Mat {  // non-owning view unless created through create()
 public:
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
  size_t elemSize() const { return (size_t)channels() * ((flags & 7) == 0 ? 1 : 4); }
  bool empty() const { return data == nullptr; }
  template <class T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
  template <class T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
};
}  // namespace cv
#endif

namespace VIDO_SLAM {

typedef cv::Mat Mat;
#ifndef VIDO_HAVE_OPENCV
enum { CV_8U = 0, CV_32S = 4, CV_32F = 5 };
#define VIDO_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
enum { CV_8UC1 = VIDO_MAKETYPE(CV_8U, 1), CV_8UC3 = VIDO_MAKETYPE(CV_8U, 3), CV_32FC1 = VIDO_MAKETYPE(CV_32F, 1),
       CV_32FC2 = VIDO_MAKETYPE(CV_32F, 2), CV_32SC1 = VIDO_MAKETYPE(CV_32S, 1) };
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
  ~System();

  // Reads the OpenCV-FileStorage style YAML the reference uses (src/config/*.yaml) -- flat "Key: value" lines and the Tbc node.
  void Init(const std::string& strSettingsFile, const eSensor sensor);

  // depthmap is modified in place (pre-scaled) exactly like the reference (src/Tracking.cc:299-322)
  cv::Mat TrackRGBD(const cv::Mat& im, cv::Mat& depthmap, const cv::Mat& flowmap, const cv::Mat& masksem, const cv::Mat& mTcw_gt,
                    const std::vector<std::vector<float> >& vObjPose_gt, const double& timestamp, cv::Mat& imTraj, const int& nImage);
  cv::Mat TrackRGBD(const cv::Mat& im, cv::Mat& depthmap, const cv::Mat& flowmap, const cv::Mat& masksem,
                    const std::vector<IMU::Point>& vImuMeas, const cv::Mat& mTcw_gt, const std::vector<std::vector<float> >& vObjPose_gt,
                    const double& timestamp, cv::Mat& imTraj, const int& nImage);
  void SaveResultsIJRR2020(const std::string& filename);

  // ---- additions of this library (no reference counterpart)
  vido_ctx* context() { return ctx_; }
  // What the demo does per frame (demo/run_vido_slam.cc:112-137) with the pixel conversions on the device: the raw Bayer RG
  // image, the 16-bit depth, the flow and the 8-bit mask exactly as the files hold them (host/InputDecode.h reads them).
  // vImuMeas: NULL for an RGBD system, the frame's IMU samples for an IMU_RGBD one.
  cv::Mat TrackRaw(const uint8_t* bayer, const uint16_t* depth16, const float* flow, const uint8_t* mask8,
                   const std::vector<IMU::Point>* vImuMeas, const double& timestamp, const int& nImage);
  bool isImuInitialized();   // Tracking::isImuInitialized

 private:
  cv::Mat track(const cv::Mat& im, cv::Mat& depthmap, const cv::Mat& flowmap, const cv::Mat& masksem, const cv::Mat& mTcw_gt,
                const double& timestamp, const int& nImage);
  cv::Mat after_track(int rc, const float* Tcw, const vido_track_stats& st, double timestamp, int nImage);
  vido_ctx* ctx_ = nullptr;
  vido_config cfg_;
  eSensor sensor_ = RGBD;
  std::vector<float> trajectory_;
  std::vector<float> cam_gt_;          // Map::vmCameraPose_GT (src/Tracking.cc:1546: the identity of the first frame only)
  std::vector<double> stage_ms_;       // per frame: mask update, camera pose, object tracking, object motion, map update, local BA
  float imu_noise_[4] = {0, 0, 0, 0};
  double last_t_ = 0;
  bool full_batch_done_ = false;
  int frame_id_ = 0;
};

}  // namespace VIDO_SLAM
