// assoc_oracle.cc -- CPU restatement of the per-frame gather stages.  TEST INFRASTRUCTURE ONLY.
//   depth pre-scale           src/Tracking.cc:299-322
//   static association        src/Frame.cc:72-100 (+ depth lookup :164-177)
//   stride-4 object sampling  src/Frame.cc:184-211
#include <cstddef>
#include "vido_oracle.h"

extern "C" {

void vo_depth_prep(float* depth, int W, int H, int stride, int choose_data, float factor, float bf, float mscale) {
  for (int i = 0; i < H; i++)
    for (int j = 0; j < W; j++) {
      float& d = depth[(size_t)i * stride + j];
      if (d < 0) d = 0;
      else {
        if (choose_data == 1) d = d / factor;                   // OMD
        else if (choose_data == 2) d = bf / (d / factor);       // KITTI
        else if (choose_data == 3) d = mscale * bf / (d / factor);  // KAIST
      }
    }
}

int vo_frame_associate(const vo_keypoint* kps, int n, const float* depth, const float* flow, const int32_t* mask, int W, int H,
                       float th, int32_t* out_idx, float* corres, float* oflow, float* odepth, int cap) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const int x = (int)kps[i].x, y = (int)kps[i].y;
    if (mask[(size_t)y * W + x] != 0) continue;
    const float d = depth[(size_t)y * W + x];
    if (d > th || d <= 0) continue;
    const float fx = flow[2 * ((size_t)y * W + x)], fy = flow[2 * ((size_t)y * W + x) + 1];
    if (fx != 0 && fy != 0) {
      if (kps[i].x + fx < W && kps[i].y + fy < H && kps[i].x < W && kps[i].y < H) {
        if (m < cap) {
          out_idx[m] = i;
          corres[2 * m] = kps[i].x + fx;
          corres[2 * m + 1] = kps[i].y + fy;
          oflow[2 * m] = fx;
          oflow[2 * m + 1] = fy;
          odepth[m] = d > 0 ? d : -1.f;
        }
        m++;
      }
    }
  }
  return m;
}

int vo_frame_sample_objects(const float* depth, const float* flow, const int32_t* mask, int W, int H, float th, float* keys,
                            float* corres, float* oflow, float* odepth, int32_t* label, int cap) {
  int m = 0;
  const int step = 4;
  for (int i = 0; i < H; i += step)
    for (int j = 0; j < W; j += step) {
      const size_t k = (size_t)i * W + j;
      if (mask[k] != 0 && depth[k] < th && depth[k] > 0) {
        const float fx = flow[2 * k], fy = flow[2 * k + 1];
        if (j + fx < W && j + fx > 0 && i + fy < H && i + fy > 0) {
          if (m < cap) {
            oflow[2 * m] = fx; oflow[2 * m + 1] = fy;
            corres[2 * m] = j + fx; corres[2 * m + 1] = i + fy;
            keys[2 * m] = (float)j; keys[2 * m + 1] = (float)i;
            odepth[m] = depth[k];
            label[m] = mask[k];
          }
          m++;
        }
      }
    }
  return m;
}

}  // extern "C"
