// run_vido_slam -- the reference's demo (vido_slam/demo/run_vido_slam.cc:68-140) on top of libvido_slam.so / libvido_b200.so:
//   ./run_vido_slam path_to_config [result_prefix]
// The config is the demo's OpenCV-FileStorage YAML (demo/utils.h:17-34: image_path, imu_path, start_index, slam_mode 0 = VO /
// otherwise VIO) which also holds the camera / tracker settings System::Init reads.  Per frame: the raw Bayer image, the .flo flow,
// the 16-bit depth and the 8-bit mask are read from image_path, ../flow_image, ../depth_image, ../mask_image (host/InputDecode.h)
// and handed to System::TrackRaw, which demosaics / converts on the device.  With a result_prefix the five result files of
// System::SaveResultsIJRR2020 are written at the end (the reference's demo stops without saving).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <string>

#include "InputDecode.h"
#include "System.h"

namespace {
std::string yaml_value(const std::string& file, const std::string& key) {
  std::ifstream in(file);
  std::string line;
  while (std::getline(in, line)) {
    const size_t c = line.find(':');
    if (c == std::string::npos) continue;
    std::string k = line.substr(0, c);
    k.erase(0, k.find_first_not_of(" \t"));
    if (k != key) continue;
    std::string v = line.substr(c + 1);
    v.erase(0, v.find_first_not_of(" \t\""));
    v.erase(v.find_last_not_of(" \t\r\"") + 1);
    return v;
  }
  return "";
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2 || argc > 3) {
    std::cerr << std::endl << "Usage: ./run_vido_slam path_to_config [result_prefix]" << std::endl;
    return 1;
  }
  using namespace VIDO_SLAM;
  const std::string config_file = argv[1];
  const std::string img_path = yaml_value(config_file, "image_path"), imu_path = yaml_value(config_file, "imu_path");
  const int start_index = atoi(yaml_value(config_file, "start_index").c_str());
  const bool vio = atoi(yaml_value(config_file, "slam_mode").c_str()) != 0;
  if (img_path.empty()) { std::cerr << "ERROR: Wrong path to settings" << std::endl; return 1; }
  std::vector<std::string> names;
  std::vector<double> times;
  std::string err;
  if (!io::load_kaist_timestamps(img_path, names, times, &err)) { std::cout << "vTimestampsImage file open failed: " << err << std::endl; return 1; }
  std::map<int, std::vector<IMU::Point> > imu_of_frame;
  if (vio) {
    std::cout << "load imu data, waiting........." << std::endl;
    std::vector<io::ImuSample> all;
    if (!io::load_kaist_imu(imu_path, all, &err)) { std::cerr << err << std::endl; return 1; }
    for (size_t idx = 1; idx < times.size(); idx++) {
      std::vector<IMU::Point> v;
      for (const io::ImuSample& s : io::imu_between(all, times[idx - 1], times[idx])) v.push_back(IMU::Point(s.ax, s.ay, s.az, s.wx, s.wy, s.wz, s.t));
      imu_of_frame[(int)idx] = v;
    }
    std::cout << "load imu data done." << std::endl;
  }
  System sys;
  sys.Init(config_file, vio ? System::IMU_RGBD : System::RGBD);
  for (size_t idx = (size_t)start_index; idx < times.size(); idx++) {
    std::cout << "\nprocessing image idx --> " << idx << std::endl;
    const std::string stem = names[idx].substr(0, 19);
    io::Image raw, depth, mask;
    int fw = 0, fh = 0;
    std::vector<float> flow;
    if (!io::read_png(img_path + "/" + names[idx], raw, &err) || !io::read_flo(img_path + "/../flow_image/" + stem + ".flo", fw, fh, flow, &err) ||
        !io::read_png(img_path + "/../depth_image/" + stem + ".png", depth, &err) || !io::read_png(img_path + "/../mask_image/" + stem + ".png", mask, &err)) {
      std::cerr << "ERROR: " << err << std::endl;
      return 2;
    }
    const bool ok = raw.channels == 1 && raw.bit_depth == 8 && depth.channels == 1 && depth.bit_depth == 16 && mask.channels == 1 && mask.bit_depth == 8 &&
                    depth.width == raw.width && depth.height == raw.height && mask.width == raw.width && mask.height == raw.height && fw == raw.width && fh == raw.height;
    if (!ok) { std::cerr << "ERROR: frame " << idx << ": unexpected image type or size" << std::endl; return 2; }
    const std::vector<IMU::Point>* meas = vio ? &imu_of_frame[(int)idx] : nullptr;
    sys.TrackRaw(raw.data.data(), (const uint16_t*)depth.data.data(), flow.data(), mask.data.data(), meas, times[idx], 10000);
  }
  if (argc == 3) sys.SaveResultsIJRR2020(argv[2]);
  return 0;
}
