// assoc_oracle.cc -- CPU restatement of the per-frame gather stages.  TEST INFRASTRUCTURE ONLY.
//   depth pre-scale           src/Tracking.cc:299-322
//   static association        src/Frame.cc:72-100 (+ depth lookup :164-177)
//   stride-4 object sampling  src/Frame.cc:184-211
#include <algorithm>
#include <map>
#include <vector>
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

// Tracking::UpdateMask (src/Tracking.cc:3291-3357).  For every semantic label of the last frame's object features, in
// ascending label order: the labels of the CURRENT mask at the features' predicted positions (int truncation, strictly
// inside) vote; with >= 100 votes and label 0 winning (mask lost; ties go to the smaller label -- std::map order kept
// by the insertion sort std::sort uses for so few entries) the last frame's pixels of that label are forward-warped
// through the last flow (int truncation of the flow, strictly inside) into the current mask.  Labels are processed one
// after the other on the mask as modified so far.  recovered[k] = 1 if the k-th unique label was warped.
int vo_update_mask(const int32_t* sem_label, const float* corres_xy, int n, const int32_t* mask_last, const float* flow_last,
                   int32_t* mask_cur, int W, int H, int32_t* uniq_out, int32_t* recovered, int cap) {
  std::vector<int32_t> uni(sem_label, sem_label + n);
  std::sort(uni.begin(), uni.end());
  uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
  for (size_t k = 0; k < uni.size(); k++) {
    const int32_t L = uni[k];
    std::map<int, int> dups;
    int votes = 0;
    for (int i = 0; i < n; i++) {
      if (sem_label[i] != L) continue;
      const int u = (int)corres_xy[2 * i], v = (int)corres_xy[2 * i + 1];
      if (u < W && u > 0 && v < H && v > 0) { ++dups[mask_cur[(size_t)v * W + u]]; votes++; }
    }
    int rec = 0;
    if (votes >= 100) {
      int best = 0, best_cnt = -1;
      for (auto& kv : dups)
        if (kv.second > best_cnt) { best_cnt = kv.second; best = kv.first; }  // first (smallest) label among equals
      if (best == 0) {
        rec = 1;
        for (int j = 0; j < H; j++)
          for (int kx = 0; kx < W; kx++)
            if (mask_last[(size_t)j * W + kx] == L) {
              const int fx = (int)flow_last[2 * ((size_t)j * W + kx)], fy = (int)flow_last[2 * ((size_t)j * W + kx) + 1];
              if (kx + fx < W && kx + fx > 0 && j + fy < H && j + fy > 0) mask_cur[(size_t)(j + fy) * W + kx + fx] = L;
            }
      }
    }
    if ((int)k < cap) { if (uniq_out) uniq_out[k] = L; if (recovered) recovered[k] = rec; }
  }
  return (int)uni.size();
}

}  // extern "C"
