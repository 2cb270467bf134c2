// orb_oracle.cc -- CPU restatement of the ORB/FAST front-end.  TEST INFRASTRUCTURE ONLY (see vido_oracle.h).
//
// Follows src/ORBextractor.cc of the reference:
//   ctor (quotas, umax)            ORBextractor.cc:400-460
//   ComputePyramid                 ORBextractor.cc:1107-1132   (cv::resize INTER_LINEAR 8U, model pinned vs cv2)
//   ComputeKeyPointsOctTree        ORBextractor.cc:755-843     (per-cell cv::FAST, thr 20 -> 7 fallback)
//   DistributeOctTree / DivideNode ORBextractor.cc:529-753, 471-527
//   IC_Angle                       ORBextractor.cc:67-94       (cv::fastAtan2 model pinned vs cv2)
//   operator() tail                ORBextractor.cc:1068-1104   (scale coords, concatenate levels)
// OpenCV (un-vendored, reference links 3.4) arithmetic is restated from the models that were
// verified bit-exact against cv2 4.13.0 (tests/golden/make_orb_golden.py regenerates the vectors).
//
// Deterministic tie-break (reference sorts by heap pointer, ORBextractor.cc:671-675): among nodes
// of equal key count the LATER-created node is expanded first (what ascending heap addresses give).
#include "vido_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <list>
#include <vector>

namespace {

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

inline int cv_round_f(float v) { return (int)lrintf(v); }  // cvRound: round-half-even (default FP mode)

// ---------------------------------------------------------------- level geometry
struct Levels {
  std::vector<float> scale, inv;
  std::vector<int> w, h, quota;
};

Levels make_levels(int W, int H, const vo_orb_params& p) {
  Levels L;
  int n = p.nlevels;
  L.scale.resize(n); L.inv.resize(n); L.w.resize(n); L.h.resize(n); L.quota.resize(n);
  L.scale[0] = 1.0f;
  for (int i = 1; i < n; i++) L.scale[i] = L.scale[i - 1] * p.scale_factor;  // ORBextractor.cc:409-413
  for (int i = 0; i < n; i++) {
    L.inv[i] = 1.0f / L.scale[i];                                            // :418-422
    L.w[i] = cv_round_f((float)W * L.inv[i]);                                // :1111-1112
    L.h[i] = cv_round_f((float)H * L.inv[i]);
  }
  float factor = 1.0f / p.scale_factor;                                      // :427-437
  float nDesired = p.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
  int sum = 0;
  for (int l = 0; l < n - 1; l++) {
    L.quota[l] = cv_round_f(nDesired);
    sum += L.quota[l];
    nDesired *= factor;
  }
  L.quota[n - 1] = std::max(p.nfeatures - sum, 0);
  return L;
}

// umax table, ORBextractor.cc:444-459
void make_umax(int* umax) {
  int v, v0, vmax = (int)floor(kHalfPatch * sqrt(2.f) / 2 + 1);
  int vmin = (int)ceil(kHalfPatch * sqrt(2.f) / 2);
  const double hp2 = kHalfPatch * kHalfPatch;
  for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
  for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

// ---------------------------------------------------------------- cv::resize 8UC1 INTER_LINEAR
void axis_tab(int srcDim, int dstDim, std::vector<int>& ofs, std::vector<short>& a0, std::vector<short>& a1) {
  ofs.resize(dstDim); a0.resize(dstDim); a1.resize(dstDim);
  double scale = (double)srcDim / dstDim;
  for (int d = 0; d < dstDim; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= srcDim - 1) { s = srcDim - 1; f = 0.f; }
    ofs[d] = s;
    a0[d] = (short)lrintf((1.f - f) * 2048.f);
    a1[d] = (short)lrintf(f * 2048.f);
  }
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  std::vector<int> xo, yo;
  std::vector<short> xa0, xa1, ya0, ya1;
  axis_tab(sw, dw, xo, xa0, xa1);
  axis_tab(sh, dh, yo, ya0, ya1);
  std::vector<int> r0(dw), r1(dw);
  for (int y = 0; y < dh; y++) {
    const uint8_t* s0 = src + (size_t)yo[y] * sstride;
    const uint8_t* s1 = src + (size_t)std::min(yo[y] + 1, sh - 1) * sstride;
    for (int x = 0; x < dw; x++) {
      int sx = xo[x], sx1 = std::min(sx + 1, sw - 1);
      r0[x] = s0[sx] * xa0[x] + s0[sx1] * xa1[x];
      r1[x] = s1[sx] * xa0[x] + s1[sx1] * xa1[x];
    }
    int b0 = ya0[y], b1 = ya1[y];
    for (int x = 0; x < dw; x++) {
      int v = (((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2;
      dst[(size_t)y * dstride + x] = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
}

// ---------------------------------------------------------------- cv::FAST TYPE_9_16
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// threshold-independent corner score: max over 9-arcs of min(d) (bright) / min(-d) (dark), minus 1.
// A pixel is a corner at threshold t  <=>  score >= t.
int fast_score_at(const uint8_t* p, int stride) {
  int d[25];
  int c = p[0];
  for (int k = 0; k < 16; k++) d[k] = (int)p[kRingDy[k] * stride + kRingDx[k]] - c;
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -1000;
  for (int k = 0; k < 16; k++) {
    int mn = d[k], mx = d[k];
    for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
    best = std::max(best, std::max(mn, -mx));
  }
  return best - 1;
}

// FAST + strict 3x3 NMS over one ROI, results in row-major scan order (coords relative to ROI)
int fast_roi(const uint8_t* img, int w, int h, int stride, int thr, std::vector<int>& xs, std::vector<int>& ys,
             std::vector<int>& sc) {
  xs.clear(); ys.clear(); sc.clear();
  if (w < 7 || h < 7) return 0;
  std::vector<int> S((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      // exact early rejection (the usual FAST high-speed test): 9 contiguous ring pixels always contain at least two of
      // the compass pixels 0, 4, 8, 12, so a corner at threshold thr needs two of them brighter than c + thr or two darker
      // than c - thr; everything else has score < thr and never reaches the full score
      const uint8_t* pc = img + (size_t)y * stride + x;
      const int c = pc[0], hi = c + thr, lo = c - thr;
      const int p0 = pc[3 * stride], p4 = pc[3], p8 = pc[-3 * stride], p12 = pc[-3];
      const int nb = (p0 > hi) + (p4 > hi) + (p8 > hi) + (p12 > hi), nd = (p0 < lo) + (p4 < lo) + (p8 < lo) + (p12 < lo);
      if (nb < 2 && nd < 2) continue;
      int s = fast_score_at(pc, stride);
      S[(size_t)y * w + x] = (s >= thr) ? s : 0;
    }
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int s = S[(size_t)y * w + x];
      if (s <= 0) continue;
      bool keep = true;
      for (int dy = -1; dy <= 1 && keep; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          if (!dx && !dy) continue;
          if (S[(size_t)(y + dy) * w + x + dx] >= s) { keep = false; break; }
        }
      if (keep) { xs.push_back(x); ys.push_back(y); sc.push_back(s); }
    }
  return (int)xs.size();
}

// ---------------------------------------------------------------- cv::fastAtan2
float fast_atan2(float y, float x) {
  const float k = (float)(180.0 / M_PI);
  const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
  const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
  float ax = fabsf(x), ay = fabsf(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------- per-level candidates (cell loop)
struct Cand { float x, y; int score; };

void level_candidates(const uint8_t* img, int cols, int rows, int stride, const vo_orb_params& p,
                      std::vector<Cand>& out) {
  out.clear();
  const float W = 30;
  const int minBX = kEdge - 3, minBY = minBX;
  const int maxBX = cols - kEdge + 3, maxBY = rows - kEdge + 3;
  const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
  const int nCols = (int)(width / W), nRows = (int)(height / W);
  if (nCols <= 0 || nRows <= 0) return;
  const int wCell = (int)ceil(width / nCols), hCell = (int)ceil(height / nRows);
  std::vector<int> xs, ys, sc;
  for (int i = 0; i < nRows; i++) {
    const float iniY = (float)(minBY + i * hCell);
    float maxY = iniY + hCell + 6;
    if (iniY >= maxBY - 3) continue;
    if (maxY > maxBY) maxY = (float)maxBY;
    for (int j = 0; j < nCols; j++) {
      const float iniX = (float)(minBX + j * wCell);
      float maxX = iniX + wCell + 6;
      if (iniX >= maxBX - 6) continue;
      if (maxX > maxBX) maxX = (float)maxBX;
      int x0 = (int)iniX, y0 = (int)iniY, rw = (int)maxX - x0, rh = (int)maxY - y0;
      const uint8_t* roi = img + (size_t)y0 * stride + x0;
      int n = fast_roi(roi, rw, rh, stride, p.ini_th_fast, xs, ys, sc);
      if (n == 0) n = fast_roi(roi, rw, rh, stride, p.min_th_fast, xs, ys, sc);
      for (int k = 0; k < n; k++) out.push_back({(float)(xs[k] + j * wCell), (float)(ys[k] + i * hCell), sc[k]});
    }
  }
}

// ---------------------------------------------------------------- quad-tree culling
struct Node {
  int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
  std::vector<int> keys;  // indices into candidate array, original relative order preserved
  bool noMore = false;
  long serial = 0;        // creation order (deterministic stand-in for the heap address)
  std::list<Node>::iterator self;
};

void divide(const Node& n, const std::vector<Cand>& C, Node& n1, Node& n2, Node& n3, Node& n4) {
  const int halfX = (int)ceil((float)(n.URx - n.ULx) / 2);
  const int halfY = (int)ceil((float)(n.BRy - n.ULy) / 2);
  n1.ULx = n.ULx; n1.ULy = n.ULy; n1.URx = n.ULx + halfX; n1.URy = n.ULy;
  n1.BLx = n.ULx; n1.BLy = n.ULy + halfY; n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
  n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = n.URx; n2.URy = n.URy;
  n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
  n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
  n3.BLx = n.BLx; n3.BLy = n.BLy; n3.BRx = n1.BRx; n3.BRy = n.BLy;
  n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
  n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = n.BRx; n4.BRy = n.BRy;
  for (int id : n.keys) {
    const Cand& kp = C[id];
    if (kp.x < n1.URx) {
      if (kp.y < n1.BRy) n1.keys.push_back(id); else n3.keys.push_back(id);
    } else if (kp.y < n1.BRy) n2.keys.push_back(id);
    else n4.keys.push_back(id);
  }
  if (n1.keys.size() == 1) n1.noMore = true;
  if (n2.keys.size() == 1) n2.noMore = true;
  if (n3.keys.size() == 1) n3.noMore = true;
  if (n4.keys.size() == 1) n4.noMore = true;
}

struct SizePtr { int size; long serial; Node* node; };

std::vector<int> distribute_octtree(const std::vector<Cand>& C, int minX, int maxX, int minY, int maxY, int N) {
  std::vector<int> result;
  if (C.empty()) return result;
  const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
  if (nIni <= 0) return result;
  const float hX = (float)(maxX - minX) / nIni;
  std::list<Node> nodes;
  std::vector<Node*> ini(nIni);
  long serial = 0;
  for (int i = 0; i < nIni; i++) {
    Node ni;
    ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
    ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
    ni.BLx = ni.ULx; ni.BLy = maxY - minY;
    ni.BRx = ni.URx; ni.BRy = maxY - minY;
    ni.serial = serial++;
    nodes.push_back(ni);
    ini[i] = &nodes.back();
  }
  for (size_t i = 0; i < C.size(); i++) {
    int idx = (int)(C[i].x / hX);
    if (idx >= nIni) idx = nIni - 1;  // (reference would index out of range; cannot happen for x < width)
    ini[idx]->keys.push_back((int)i);
  }
  for (auto it = nodes.begin(); it != nodes.end();) {
    if (it->keys.size() == 1) { it->noMore = true; ++it; }
    else if (it->keys.empty()) it = nodes.erase(it);
    else ++it;
  }
  bool finish = false;
  std::vector<SizePtr> vsp;
  auto push_child = [&](Node& c, bool count, int& nToExpand) {
    if (c.keys.empty()) return;
    c.serial = serial++;
    nodes.push_front(c);
    if (c.keys.size() > 1) {
      if (count) nToExpand++;
      vsp.push_back({(int)c.keys.size(), nodes.front().serial, &nodes.front()});
      nodes.front().self = nodes.begin();
    }
  };
  while (!finish) {
    int prevSize = (int)nodes.size();
    auto lit = nodes.begin();
    int nToExpand = 0;
    vsp.clear();
    while (lit != nodes.end()) {
      if (lit->noMore) { ++lit; continue; }
      Node n1, n2, n3, n4;
      divide(*lit, C, n1, n2, n3, n4);
      push_child(n1, true, nToExpand);
      push_child(n2, true, nToExpand);
      push_child(n3, true, nToExpand);
      push_child(n4, true, nToExpand);
      lit = nodes.erase(lit);
    }
    if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
      finish = true;
    } else if ((int)nodes.size() + nToExpand * 3 > N) {
      while (!finish) {
        prevSize = (int)nodes.size();
        std::vector<SizePtr> prev = vsp;
        vsp.clear();
        std::sort(prev.begin(), prev.end(), [](const SizePtr& a, const SizePtr& b) {
          return a.size != b.size ? a.size < b.size : a.serial < b.serial;
        });
        for (int j = (int)prev.size() - 1; j >= 0; j--) {
          Node n1, n2, n3, n4;
          divide(*prev[j].node, C, n1, n2, n3, n4);
          int dummy = 0;
          push_child(n1, false, dummy);
          push_child(n2, false, dummy);
          push_child(n3, false, dummy);
          push_child(n4, false, dummy);
          nodes.erase(prev[j].node->self);
          if ((int)nodes.size() >= N) break;
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
      }
    }
  }
  for (auto& n : nodes) {
    int best = n.keys[0];
    int maxR = C[best].score;
    for (size_t k = 1; k < n.keys.size(); k++)
      if (C[n.keys[k]].score > maxR) { best = n.keys[k]; maxR = C[best].score; }
    result.push_back(best);
  }
  return result;
}

float ic_angle(const uint8_t* img, int stride, int px, int py, const int* umax) {
  int m01 = 0, m10 = 0;
  const uint8_t* c = img + (size_t)py * stride + px;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vs = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int vp = c[u + v * stride], vm = c[u - v * stride];
      vs += (vp - vm);
      m10 += u * (vp + vm);
    }
    m01 += v * vs;
  }
  return fast_atan2((float)m01, (float)m10);
}

void build_pyramid(const uint8_t* gray, int W, int H, int stride, const Levels& L, std::vector<std::vector<uint8_t>>& pyr) {
  int n = (int)L.w.size();
  pyr.resize(n);
  pyr[0].resize((size_t)W * H);
  for (int y = 0; y < H; y++) memcpy(&pyr[0][(size_t)y * W], gray + (size_t)y * stride, W);
  for (int l = 1; l < n; l++) {
    pyr[l].resize((size_t)L.w[l] * L.h[l]);
    resize_linear_u8(pyr[l - 1].data(), L.w[l - 1], L.h[l - 1], L.w[l - 1], pyr[l].data(), L.w[l], L.h[l], L.w[l]);
  }
}

}  // namespace

extern "C" {

int vo_orb_level_sizes(int W, int H, const vo_orb_params* p, int* w, int* h, float* scale) {
  Levels L = make_levels(W, H, *p);
  for (int i = 0; i < p->nlevels; i++) { w[i] = L.w[i]; h[i] = L.h[i]; if (scale) scale[i] = L.scale[i]; }
  return p->nlevels;
}

int vo_orb_level_quotas(const vo_orb_params* p, int* quota) {
  Levels L = make_levels(64, 64, *p);
  for (int i = 0; i < p->nlevels; i++) quota[i] = L.quota[i];
  return p->nlevels;
}

void vo_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  resize_linear_u8(src, sw, sh, sstride, dst, dw, dh, dstride);
}

int vo_fast_roi(const uint8_t* img, int w, int h, int stride, int thr, int* xs, int* ys, int* scores, int cap) {
  std::vector<int> x, y, s;
  int n = fast_roi(img, w, h, stride, thr, x, y, s);
  for (int i = 0; i < n && i < cap; i++) { xs[i] = x[i]; ys[i] = y[i]; scores[i] = s[i]; }
  return n;
}

float vo_fast_atan2(float y, float x) { return fast_atan2(y, x); }

int vo_orb_level_candidates(const uint8_t* img, int w, int h, int stride, const vo_orb_params* p, int* xs, int* ys,
                            int* scores, int cap) {
  std::vector<Cand> c;
  level_candidates(img, w, h, stride, *p, c);
  for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xs[i] = (int)c[i].x; ys[i] = (int)c[i].y; scores[i] = c[i].score; }
  return (int)c.size();
}

int vo_orb_pyramid(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p, uint8_t* out, int64_t* offsets) {
  Levels L = make_levels(W, H, *p);
  std::vector<std::vector<uint8_t>> pyr;
  build_pyramid(gray, W, H, stride, L, pyr);
  int64_t off = 0;
  for (int l = 0; l < p->nlevels; l++) {
    offsets[l] = off;
    memcpy(out + off, pyr[l].data(), pyr[l].size());
    off += (int64_t)pyr[l].size();
  }
  offsets[p->nlevels] = off;
  return p->nlevels;
}

// ORBextractor::operator() (src/ORBextractor.cc:1034-1105).  desc != nullptr additionally does what the reference's loop at
// :1067-1101 was written to do: blur a clone of the level (:1078-1079) and describe the level's key points at their LEVEL
// coordinates (:1086, commented out in the reference) before they are scaled to level 0 (:1094-1100).
static int extract_impl(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p, vo_keypoint* out, int cap, uint8_t* desc) {
  Levels L = make_levels(W, H, *p);
  int umax[kHalfPatch + 2];
  make_umax(umax);
  std::vector<std::vector<uint8_t>> pyr;
  build_pyramid(gray, W, H, stride, L, pyr);
  int total = 0;
  std::vector<Cand> cand;
  std::vector<uint8_t> blurred;
  for (int l = 0; l < p->nlevels; l++) {
    int cols = L.w[l], rows = L.h[l];
    level_candidates(pyr[l].data(), cols, rows, cols, *p, cand);
    const int minBX = kEdge - 3, minBY = minBX, maxBX = cols - kEdge + 3, maxBY = rows - kEdge + 3;
    std::vector<int> keep = distribute_octtree(cand, minBX, maxBX, minBY, maxBY, L.quota[l]);
    const int scaledPatch = (int)(kPatch * L.scale[l]);
    if (desc && !keep.empty()) {
      blurred.resize((size_t)cols * rows);
      vo_gauss7_u8(pyr[l].data(), cols, rows, cols, blurred.data(), cols);
    }
    for (int id : keep) {
      if (total >= cap) return -1;
      vo_keypoint kp;
      float x = cand[id].x + minBX, y = cand[id].y + minBY;  // ORBextractor.cc:832-833
      kp.octave = l;
      kp.size = (float)scaledPatch;
      kp.response = (float)cand[id].score;
      kp.angle = ic_angle(pyr[l].data(), cols, cv_round_f(x), cv_round_f(y), umax);
      if (desc) vo_orb_describe_level(blurred.data(), cols, rows, &x, &y, &kp.angle, 1, desc + (size_t)total * 32);
      if (l != 0) { x *= L.scale[l]; y *= L.scale[l]; }       // ORBextractor.cc:1096-1099
      kp.x = x; kp.y = y;
      out[total++] = kp;
    }
  }
  return total;
}

int vo_orb_extract(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p, vo_keypoint* out, int cap) {
  return extract_impl(gray, W, H, stride, p, out, cap, nullptr);
}

int vo_orb_extract_describe(const uint8_t* gray, int W, int H, int stride, const vo_orb_params* p, vo_keypoint* out, int cap,
                            uint8_t* desc) {
  return extract_impl(gray, W, H, stride, p, out, cap, desc);
}

}  // extern "C"
