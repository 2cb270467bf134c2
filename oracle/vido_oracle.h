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
/* pyramid only: writes level l at out + offsets[l] (tight rows of width w[l]) */
int vo_orb_pyramid(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p,
                   uint8_t* out, int64_t* offsets);

#ifdef __cplusplus
}
#endif
#endif
