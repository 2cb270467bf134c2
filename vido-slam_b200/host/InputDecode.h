// InputDecode.h -- the on-disk formats the reference's demo reads before every TrackRGBD call
// (vido_slam/demo/run_vido_slam.cc:14-62, 114-122), without OpenCV:
//   raw camera image   8-bit PNG, Bayer RG mosaic      cv::imread(IMREAD_UNCHANGED) + cvtColor(COLOR_BayerRG2BGR)   :114-117
//   depth              16-bit PNG                      cv::imread(CV_LOAD_IMAGE_ANYDEPTH) + convertTo(CV_32F)       :119-120
//   semantic mask      8-bit PNG                       cv::imread(CV_LOAD_IMAGE_UNCHANGED) + convertTo(CV_32SC1)    :121-122
//   optical flow       Middlebury .flo                 cv::optflow::readOpticalFlow                                 :118
//   KAIST IMU csv      columns 0 (ns), 8-10 gyro, 11-13 acc                                                         :14-45
//   image time stamps  vTimestampsImage.txt (first line skipped, ns)                                                :47-66
// The files are parsed on the host (zlib inflate + PNG filters are sequential byte work); the pixel conversions
// (Bayer -> BGR, u16 -> f32, u8 -> i32) run on the device: vido_convert_raw in include/vido_b200.h.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace VIDO_SLAM {
namespace io {

struct Image {           // decoded PNG: rows top to bottom, samples in host byte order
  int width = 0, height = 0, channels = 0, bit_depth = 0;
  std::vector<uint8_t> data;   // width * height * channels * (bit_depth / 8) bytes
};

// Non-interlaced PNG of colour type 0 (grey), 2 (RGB), 4 (grey + alpha) or 6 (RGBA), 8 or 16 bits per sample.  Returns false and
// sets `err` otherwise (missing file, bad signature / CRC-less truncation, palette or interlaced images).
bool read_png(const std::string& path, Image& out, std::string* err = nullptr);
// Middlebury .flo: "PIEH", int32 width, int32 height, then width * height (u, v) float32 pairs
bool read_flo(const std::string& path, int& width, int& height, std::vector<float>& uv, std::string* err = nullptr);

struct ImuSample { double t; float ax, ay, az, wx, wy, wz; };
// LoadIMU (run_vido_slam.cc:14-45): lines starting with '#' skipped, comma separated, t = col0 / 1e9, gyro = cols 8-10, acc = cols 11-13
bool load_kaist_imu(const std::string& path, std::vector<ImuSample>& out, std::string* err = nullptr);
// LoadKaistImg (:47-66): image_dir/../vTimestampsImage.txt, header line skipped; name = first 19 characters of the printed number + ".png"
bool load_kaist_timestamps(const std::string& image_dir, std::vector<std::string>& names, std::vector<double>& times, std::string* err = nullptr);
// the samples of frame idx: lastImageTime <= t <= imageTime (:89-101); frame 0 has none
std::vector<ImuSample> imu_between(const std::vector<ImuSample>& all, double t_last, double t_cur);

}  // namespace io
}  // namespace VIDO_SLAM

// C wrappers for bindings / tests (same library)
extern "C" {
int vido_io_read_png(const char* path, int32_t* width, int32_t* height, int32_t* channels, int32_t* bit_depth, uint8_t* buf, size_t cap);
int vido_io_read_flo(const char* path, int32_t* width, int32_t* height, float* uv, size_t cap_floats);
int vido_io_load_kaist_imu(const char* path, double* rows7, int cap_rows);   /* rows of (t, ax, ay, az, wx, wy, wz); returns the count */
int vido_io_load_kaist_timestamps(const char* image_dir, double* times, char* names20, int cap);   /* names: 20 bytes each, NUL padded */
}
