// desc_oracle.cc -- CPU restatement of the descriptor stage (TEST INFRASTRUCTURE ONLY: tests/, smoke(), bench.py's CPU legs).
//   vo_gauss7_u8            GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101)  src/ORBextractor.cc:1079
//   vo_orb_describe_level   computeDescriptors / computeOrbDescriptor                                  src/ORBextractor.cc:98-137, 1023-1031
//   vo_hamming_match        no counterpart in the reference: cv::BFMatcher(NORM_HAMMING) semantics
// OpenCV is an un-vendored dependency of the reference (3.4.x; absent from /root/reference).  GaussianBlur on CV_8U is OpenCV's
// fixed-point smoother (modules/imgproc/src/smooth.simd.hpp + fixedpoint.inl.hpp): restated below from the published algorithm and
// pinned against cv2 4.13 -- blurred pyramid bytes, descriptors of cv2.ORB.compute on the same key points, BFMatcher results --
// in tests/golden/desc_golden.npz (tests/golden/make_desc_golden.py, tests/test_desc_oracle.py).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../include/vido_orb_pattern.h"
#include "vido_oracle.h"

namespace {

inline int reflect101(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// getGaussianKernel(7, 2) in 8.8 fixed point with the rounding error diffused (getGaussianKernelFixedPoint_ED): the taps sum to 256
const unsigned kTap[7] = {18, 34, 48, 56, 48, 34, 18};

}  // namespace

extern "C" {

void vo_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  // horizontal pass: ufixedpoint16 (8 fractional bits) per pixel; vertical pass: ufixedpoint32, rounded half up to 8 bits
  std::vector<uint16_t> hrow((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t* s = src + (size_t)y * sstride;
    for (int x = 0; x < w; x++) {
      unsigned acc = 0;
      for (int k = 0; k < 7; k++) acc += kTap[k] * s[reflect101(x + k - 3, w)];
      hrow[(size_t)y * w + x] = (uint16_t)acc;   // <= 255 * 256: never saturates
    }
  }
  for (int y = 0; y < h; y++) {
    uint8_t* d = dst + (size_t)y * dstride;
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int k = 0; k < 7; k++) acc += kTap[k] * (uint32_t)hrow[(size_t)reflect101(y + k - 3, h) * w + x];
      d[x] = (uint8_t)((acc + 0x8000u) >> 16);
    }
  }
}

void vo_orb_describe_level(const uint8_t* blurred, int w, int h, const float* xs, const float* ys, const float* angles, int n,
                           uint8_t* desc) {
  const float factorPI = (float)(M_PI / 180.f);   // src/ORBextractor.cc:97
  const long long size = (long long)w * h;
  for (int i = 0; i < n; i++) {
    const float angle = angles[i] * factorPI;
    const float a = (float)cos((double)angle), b = (float)sin((double)angle);
    // center = &img.at<uchar>(cvRound(kpt.pt.y), cvRound(kpt.pt.x)); step = img.step (the clone is continuous: step = w)
    const long long center = (long long)lrintf(ys[i]) * w + lrintf(xs[i]);
    const int8_t* pat = vido_orb_pattern_31;
    for (int byte = 0; byte < 32; byte++, pat += 32) {
      int val = 0;
      for (int t = 0; t < 8; t++) {
        int v[2];
        for (int e = 0; e < 2; e++) {
          const float px = (float)pat[4 * t + 2 * e], py = (float)pat[4 * t + 2 * e + 1];
          const long long at = center + (long long)lrintf(px * b + py * a) * w + lrintf(px * a - py * b);   // GET_VALUE
          // inside the clone this is the reference's read (a horizontal overshoot lands in the neighbouring row, like there);
          // outside the buffer the reference reads whatever precedes / follows its allocation: defined as 0 here
          v[e] = (at >= 0 && at < size) ? blurred[at] : 0;
        }
        val |= (v[0] < v[1]) << t;
      }
      desc[(size_t)i * 32 + byte] = (uint8_t)val;
    }
  }
}

void vo_hamming_match(const uint8_t* query, int nq, const uint8_t* train, int nt, int32_t* best_idx, int32_t* best_dist,
                      int32_t* second_dist) {
  for (int i = 0; i < nq; i++) {
    int best = 0x7fffffff, second = 0x7fffffff, idx = -1;
    for (int j = 0; j < nt; j++) {
      int d = 0;
      for (int k = 0; k < 32; k++) d += __builtin_popcount((unsigned)(query[(size_t)i * 32 + k] ^ train[(size_t)j * 32 + k]));
      if (d < best) { second = best; best = d; idx = j; }
      else if (d < second) second = d;
    }
    best_idx[i] = idx; best_dist[i] = best; second_dist[i] = second;
  }
}

}  // extern "C"
