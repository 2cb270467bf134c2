// desc_device.h -- per-thread bodies of the descriptor stage (rows K5 / K6 of SURVEY.md section 3; x1 in DESIGN.md):
//   blur7_thread     cv::GaussianBlur(level, level, Size(7,7), 2, 2, BORDER_REFLECT_101)   src/ORBextractor.cc:1078-1079
//   rbrief_thread    computeOrbDescriptor                                                  src/ORBextractor.cc:98-137
//   hamming_partial_thread / hamming_merge_thread    brute-force Hamming match (256-bit descriptors); the reference has no
//                    matcher -- the distance is ORB-SLAM's DescriptorDistance = popcount of the XOR, the result is
//                    what cv::BFMatcher(NORM_HAMMING) returns for k = 2 (best train index, its distance, second distance)
// Every body is a pure function of its global thread coordinates and its parameter block: no shared memory, no barriers, no
// atomics.  The __global__ wrappers in desc_kernels.cu only form the coordinates.  That makes the same source compilable for
// the host: tests/desc_emul.cc walks the launch grids thread by thread on the CPU and the CPU suite compares the result with
// the oracle bit for bit -- the GPU box is not needed to know that the arithmetic, the indexing and the border handling agree.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define VIDO_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <stdlib.h>
#define VIDO_HD inline
#endif

// ---- small portability layer (device intrinsic / host equivalent with the same IEEE result) ----
VIDO_HD int vd_round_f(float v) {   // cvRound: round to nearest even
#ifdef __CUDA_ARCH__
  return __float2int_rn(v);
#else
  return (int)lrintf(v);
#endif
}
VIDO_HD float vd_mul(float a, float b) {   // no FMA contraction: the reference's float code is not contracted
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
VIDO_HD float vd_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
VIDO_HD float vd_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
VIDO_HD int vd_popc(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
VIDO_HD uint32_t vd_load_u32(const uint8_t* p) {   // p is 4-byte aligned
#ifdef __CUDA_ARCH__
  return *(const uint32_t*)p;
#else
#ifdef VIDO_EMUL_CHECK_ALIGN   // the CPU emulation under AddressSanitizer also checks what the device requires of a vector access
  if ((uintptr_t)p & 3) abort();
#endif
  uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
VIDO_HD void vd_store_u32(uint8_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
  *(uint32_t*)p = v;
#else
#ifdef VIDO_EMUL_CHECK_ALIGN
  if ((uintptr_t)p & 3) abort();
#endif
  memcpy(p, &v, 4);
#endif
}
VIDO_HD int vd_reflect101(int i, int n) {   // BORDER_REFLECT_101 for -n < i < 2n-1; indices further out (only ever needed for
  if (i < 0) i = -i;                        // results that are not written) are clamped into the row / column range
  if (i >= n) i = 2 * n - 2 - i;
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

#define VIDO_DESC_MAX_LEVELS 8

// ================================================================================================================
// 7x7 Gaussian, sigma 2, 8-bit.  OpenCV smooths CV_8U in fixed point (modules/imgproc/src/smooth.simd.hpp: ufixedpoint16 rows,
// ufixedpoint32 columns): the kernel is {18, 34, 48, 56, 48, 34, 18} / 256 (the bit-exact kernel of getGaussianKernel(7, 2) with
// its rounding error diffused so that the taps add up to one), the horizontal pass keeps 8 fractional bits, the vertical pass
// 16, and the result is (sum + 0x8000) >> 16.  No saturation can occur (255 * 256 < 65536).  Pinned against cv2 in
// tests/golden/desc_golden.npz.
// One thread = 4 consecutive pixels x BLUR_ROWS rows of one level of one frame.  It forms the horizontal sums of the
// BLUR_ROWS + 6 source rows it needs (three aligned 32-bit loads per row away from the left / right border, reflected byte loads
// at the border) and keeps them in registers; vertically adjacent threads re-read 6 of those rows through L1/L2, HBM sees every
// source byte once and every result byte once.
// ================================================================================================================
#define BLUR_ROWS 4
struct BlurLevel {
  int w, h, pitch;          // level size, row pitch in bytes (multiple of 4, rows 4-byte aligned)
  int quads;                // (w + 3) / 4
  int strips;               // (h + BLUR_ROWS - 1) / BLUR_ROWS
  int first;                // first work item of this level in a frame's item list
  long long base;           // byte offset of the level inside the pyramid allocation
  long long frame_stride;   // bytes between the batch slots of this level
};
struct BlurParams {
  int nlevels, items_per_frame, nframes, pad;
  BlurLevel lv[VIDO_DESC_MAX_LEVELS];
};

// launch geometry, shared by the launcher (desc_kernels.cu) and the CPU emulation: level l of the pyramid joins the work list
inline void blur_params_add_level(BlurParams& P, int l, int w, int h, int pitch, long long base, long long frame_stride) {
  BlurLevel& b = P.lv[l];
  b.w = w; b.h = h; b.pitch = pitch;
  b.quads = (w + 3) / 4;
  b.strips = (h + BLUR_ROWS - 1) / BLUR_ROWS;
  b.first = l ? P.lv[l - 1].first + P.lv[l - 1].quads * P.lv[l - 1].strips : 0;
  b.base = base; b.frame_stride = frame_stride;
  P.items_per_frame = b.first + b.quads * b.strips;
  if (P.nlevels < l + 1) P.nlevels = l + 1;
}
#define BLUR_THREADS 128
inline unsigned blur_grid_x(const BlurParams& P) { return (unsigned)((P.items_per_frame + BLUR_THREADS - 1) / BLUR_THREADS); }

VIDO_HD void blur7_hrow(const uint8_t* row, int x0, int w, uint32_t* hs) {
  uint32_t p[10];   // pixels x0-3 .. x0+6
  if (x0 >= 4 && x0 + 8 <= w) {
    const uint32_t a = vd_load_u32(row + x0 - 4), b = vd_load_u32(row + x0), c = vd_load_u32(row + x0 + 4);
    p[0] = (a >> 8) & 0xff; p[1] = (a >> 16) & 0xff; p[2] = a >> 24;
    p[3] = b & 0xff; p[4] = (b >> 8) & 0xff; p[5] = (b >> 16) & 0xff; p[6] = b >> 24;
    p[7] = c & 0xff; p[8] = (c >> 8) & 0xff; p[9] = (c >> 16) & 0xff;
  } else {
#pragma unroll
    for (int i = 0; i < 10; i++) p[i] = row[vd_reflect101(x0 - 3 + i, w)];
  }
#pragma unroll
  for (int c = 0; c < 4; c++)
    hs[c] = 18u * (p[c] + p[c + 6]) + 34u * (p[c + 1] + p[c + 5]) + 48u * (p[c + 2] + p[c + 4]) + 56u * p[c + 3];
}

// gx = work item inside the frame (level-major, then strip, then 4-pixel group), gz = frame
VIDO_HD void blur7_thread(int gx, int gz, const BlurParams& P, const uint8_t* src, uint8_t* dst) {
  if (gx >= P.items_per_frame || gz >= P.nframes) return;
  int l = 0;
#pragma unroll
  for (int k = 1; k < VIDO_DESC_MAX_LEVELS; k++)
    if (k < P.nlevels && gx >= P.lv[k].first) l = k;
  const BlurLevel& L = P.lv[l];
  const int item = gx - L.first;
  const int strip = item / L.quads, q = item - strip * L.quads;
  const int x0 = 4 * q, y0 = strip * BLUR_ROWS;
  if (L.w < 4 || L.h < 4) return;   // reflection would leave the image; the launcher rejects such levels
  const uint8_t* s = src + L.base + (long long)gz * L.frame_stride;
  uint8_t* d = dst + L.base + (long long)gz * L.frame_stride;
  uint32_t hs[BLUR_ROWS + 6][4];
#pragma unroll
  for (int r = 0; r < BLUR_ROWS + 6; r++) {
    const int yy = vd_reflect101(y0 - 3 + r, L.h);
    blur7_hrow(s + (long long)yy * L.pitch, x0, L.w, hs[r]);
  }
#pragma unroll
  for (int r = 0; r < BLUR_ROWS; r++) {
    const int y = y0 + r;
    if (y >= L.h) break;
    uint32_t o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const uint32_t v = 18u * (hs[r][c] + hs[r + 6][c]) + 34u * (hs[r + 1][c] + hs[r + 5][c]) + 48u * (hs[r + 2][c] + hs[r + 4][c]) +
                         56u * hs[r + 3][c];
      o[c] = (v + 0x8000u) >> 16;
    }
    uint8_t* out = d + (long long)y * L.pitch + x0;
    if (x0 + 4 <= L.w) {
      vd_store_u32(out, o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24));
    } else {
      for (int c = 0; c < 4 && x0 + c < L.w; c++) out[c] = (uint8_t)o[c];
    }
  }
}

// ---- variant 2 (opt-in, VIDO_BLUR=v2; measured beside variant 1 by bench.py, not yet the default: it was written after the round's
// GPU budget was spent and has only run in the CPU emulation).  One thread = 4 consecutive pixels x BLUR2_ROWS rows, walking down its
// column with the horizontal sums of the last seven source rows in a register ring: every source row is loaded and summed
// (BLUR2_ROWS + 6) / BLUR2_ROWS = 1.4 times instead of 2.5, and the per-thread set-up (level search, division) is paid once per 16 rows.
#define BLUR2_ROWS 16
inline void blur2_params_add_level(BlurParams& P, int l, int w, int h, int pitch, long long base, long long frame_stride) {
  BlurLevel& b = P.lv[l];
  b.w = w; b.h = h; b.pitch = pitch;
  b.quads = (w + 3) / 4;
  b.strips = (h + BLUR2_ROWS - 1) / BLUR2_ROWS;
  b.first = l ? P.lv[l - 1].first + P.lv[l - 1].quads * P.lv[l - 1].strips : 0;
  b.base = base; b.frame_stride = frame_stride;
  P.items_per_frame = b.first + b.quads * b.strips;
  if (P.nlevels < l + 1) P.nlevels = l + 1;
}

VIDO_HD void blur7_thread_v2(int gx, int gz, const BlurParams& P, const uint8_t* src, uint8_t* dst) {
  if (gx >= P.items_per_frame || gz >= P.nframes) return;
  int l = 0;
#pragma unroll
  for (int k = 1; k < VIDO_DESC_MAX_LEVELS; k++)
    if (k < P.nlevels && gx >= P.lv[k].first) l = k;
  const BlurLevel& L = P.lv[l];
  const int item = gx - L.first;
  const int strip = item / L.quads, q = item - strip * L.quads;
  const int x0 = 4 * q, y0 = strip * BLUR2_ROWS;
  if (L.w < 4 || L.h < 4) return;
  const uint8_t* s = src + L.base + (long long)gz * L.frame_stride;
  uint8_t* d = dst + L.base + (long long)gz * L.frame_stride;
  uint32_t ring[7][4];
#pragma unroll
  for (int r = 0; r < 6; r++) blur7_hrow(s + (long long)vd_reflect101(y0 - 3 + r, L.h) * L.pitch, x0, L.w, ring[r]);
#pragma unroll
  for (int r = 0; r < BLUR2_ROWS; r++) {
    const int y = y0 + r;
    if (y >= L.h) break;
    blur7_hrow(s + (long long)vd_reflect101(y + 3, L.h) * L.pitch, x0, L.w, ring[(r + 6) % 7]);
    uint32_t o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const uint32_t v = 18u * (ring[r % 7][c] + ring[(r + 6) % 7][c]) + 34u * (ring[(r + 1) % 7][c] + ring[(r + 5) % 7][c]) +
                         48u * (ring[(r + 2) % 7][c] + ring[(r + 4) % 7][c]) + 56u * ring[(r + 3) % 7][c];
      o[c] = (v + 0x8000u) >> 16;
    }
    uint8_t* out = d + (long long)y * L.pitch + x0;
    if (x0 + 4 <= L.w) {
      vd_store_u32(out, o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24));
    } else {
      for (int c = 0; c < 4 && x0 + c < L.w; c++) out[c] = (uint8_t)o[c];
    }
  }
}

// ================================================================================================================
// rBRIEF.  One thread = one byte (8 tests) of one key point's descriptor; the 32 lanes of a warp write the 32 bytes of one
// descriptor.  Key points arrive as the extraction wrote them (level-0 coordinates, src/ORBextractor.cc:1094-1100); the level
// coordinates the reference describes at (integers, :98-105) are recovered exactly by rounding x / scale.
// Reads follow the reference's addressing on its tight clone of the level (center[iy * step + ix], step = level width): a
// rotated test point that leaves the row horizontally lands in the neighbouring row, like there.  A key point sits >= 16 pixels
// from the border (minBorder = EDGE_THRESHOLD - 3) while rotated pattern points reach 19 pixels, so for key points in the outer
// three rows of the detection zone the reference reads before / after its buffer (undefined there): those reads return 0 here
// and in the oracle.
// ================================================================================================================
struct DescLevel {
  int w, h, pitch, pad;
  long long base, frame_stride;
  float scale, pad2;
};
struct DescParams {
  int nlevels, nframes, cap_per_frame, pad;
  DescLevel lv[VIDO_DESC_MAX_LEVELS];
};
struct DescKeyPoint { float x, y, size, angle, response; int32_t octave; };   // = vido_keypoint
inline void desc_params_add_level(DescParams& P, int l, int w, int h, int pitch, long long base, long long frame_stride, float scale) {
  DescLevel& d = P.lv[l];
  d.w = w; d.h = h; d.pitch = pitch; d.base = base; d.frame_stride = frame_stride; d.scale = scale;
  if (P.nlevels < l + 1) P.nlevels = l + 1;
}
#define RBRIEF_THREADS 256
inline unsigned rbrief_grid_x(int cap_per_frame) { return (unsigned)((cap_per_frame * 32 + RBRIEF_THREADS - 1) / RBRIEF_THREADS); }

VIDO_HD uint32_t rbrief_fetch(const uint8_t* img, const DescLevel& L, int cx, int cy, int px, int py, float a, float b) {
  const int iy = vd_round_f(vd_add(vd_mul((float)px, b), vd_mul((float)py, a)));
  const int ix = vd_round_f(vd_add(vd_mul((float)px, a), -vd_mul((float)py, b)));
  int xx = cx + ix, yy = cy + iy;
  if (xx < 0) { xx += L.w; yy -= 1; }
  else if (xx >= L.w) { xx -= L.w; yy += 1; }
  if (yy < 0 || yy >= L.h) return 0u;
  return img[(long long)yy * L.pitch + xx];
}

// gx = key point * 32 + byte, gz = frame; pattern = the 512 (x, y) int8 pairs of include/vido_orb_pattern.h in device memory
VIDO_HD void rbrief_thread(int gx, int gz, const DescParams& P, const uint8_t* blurred, const DescKeyPoint* kps, const int32_t* nkp,
                           const int8_t* pattern, uint8_t* desc) {
  if (gz >= P.nframes) return;
  const int k = gx >> 5, byte = gx & 31;
  if (k >= P.cap_per_frame || k >= nkp[gz]) return;
  const DescKeyPoint kp = kps[(long long)gz * P.cap_per_frame + k];
  int l = kp.octave;
  if (l < 0 || l >= P.nlevels) l = 0;
  const DescLevel& L = P.lv[l];
  const int cx = vd_round_f(l ? vd_div(kp.x, L.scale) : kp.x), cy = vd_round_f(l ? vd_div(kp.y, L.scale) : kp.y);
  // float angle = kpt.angle * factorPI; a = (float)cos(angle), b = (float)sin(angle): evaluated in double and rounded once (the
  // float result of a correctly rounded cosine; CUDA's double cos / sin are within 2 ulp of it, far inside the float rounding)
  const float factorPI = (float)(3.14159265358979323846 / 180.0);
  const float ang = vd_mul(kp.angle, factorPI);
  const float a = (float)cos((double)ang), b = (float)sin((double)ang);
  const uint8_t* img = blurred + L.base + (long long)gz * L.frame_stride;
  const int8_t* pt = pattern + byte * 32;   // 16 points (x, y) per descriptor byte; test t compares point 2t with point 2t + 1
  uint32_t val = 0;
#pragma unroll
  for (int t = 0; t < 8; t++) {
    const uint32_t t0 = rbrief_fetch(img, L, cx, cy, pt[4 * t], pt[4 * t + 1], a, b);
    const uint32_t t1 = rbrief_fetch(img, L, cx, cy, pt[4 * t + 2], pt[4 * t + 3], a, b);
    val |= (uint32_t)(t0 < t1) << t;
  }
  desc[((long long)gz * P.cap_per_frame + k) * 32 + byte] = (uint8_t)val;
}

// ================================================================================================================
// Hamming match, brute force.  Pair p matches the nq[p] query descriptors at q + p * q_stride against the nt[p] train descriptors
// at t + p * t_stride (32 bytes each, 4-byte aligned).  The train set is cut into HAM_CHUNKS contiguous ranges; thread (query,
// chunk, pair) scans its range with the query in 8 registers (every lane of a warp reads the same train descriptor: one L1
// broadcast per 32 queries) and writes (best distance, best index, second distance); the merge thread of a query folds the
// chunks in ascending order, so among equal distances the lowest train index wins -- the first minimum, like cv::BFMatcher.
// ================================================================================================================
#define HAM_CHUNKS 8
#define HAM_NONE 0x7fffffff
struct HamParams {
  int npairs, qcap;               // qcap = queries per pair the partial / output arrays are laid out for
  long long q_stride, t_stride;   // bytes between pairs
};

#define HAM_THREADS 128
inline unsigned hamming_grid_x(int qcap) { return (unsigned)((qcap + HAM_THREADS - 1) / HAM_THREADS); }
inline size_t hamming_part_bytes(int npairs, int qcap) { return (size_t)npairs * HAM_CHUNKS * qcap * 3 * sizeof(int32_t); }

VIDO_HD void hamming_partial_thread(int gx, int gy, int gz, const HamParams& P, const uint8_t* q, const uint8_t* t, const int32_t* nq,
                                    const int32_t* nt, int32_t* part) {
  if (gz >= P.npairs || gy >= HAM_CHUNKS || gx >= P.qcap || gx >= nq[gz]) return;
  const int n = nt[gz];
  const int len = (n + HAM_CHUNKS - 1) / HAM_CHUNKS;
  const int j0 = gy * len, j1 = (j0 + len < n) ? j0 + len : n;
  const uint8_t* qp = q + (long long)gz * P.q_stride + (long long)gx * 32;
  uint32_t qw[8];
#pragma unroll
  for (int i = 0; i < 8; i++) qw[i] = vd_load_u32(qp + 4 * i);
  int best = HAM_NONE, bidx = -1, second = HAM_NONE;
  const uint8_t* tp = t + (long long)gz * P.t_stride;
  for (int j = j0; j < j1; j++) {
    const uint8_t* d = tp + (long long)j * 32;
    int dist = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) dist += vd_popc(qw[i] ^ vd_load_u32(d + 4 * i));
    if (dist < best) { second = best; best = dist; bidx = j; }
    else if (dist < second) second = dist;
  }
  int32_t* o = part + (((long long)gz * HAM_CHUNKS + gy) * P.qcap + gx) * 3;
  o[0] = best; o[1] = bidx; o[2] = second;
}

VIDO_HD void hamming_merge_thread(int gx, int gz, const HamParams& P, const int32_t* nq, const int32_t* part, int32_t* best_idx,
                                  int32_t* best_dist, int32_t* second_dist) {
  if (gz >= P.npairs || gx >= P.qcap || gx >= nq[gz]) return;
  int best = HAM_NONE, bidx = -1, second = HAM_NONE;
  for (int c = 0; c < HAM_CHUNKS; c++) {
    const int32_t* o = part + (((long long)gz * HAM_CHUNKS + c) * P.qcap + gx) * 3;
    const int b = o[0], i = o[1], s = o[2];
    if (b < best) { if (best < second) second = best; best = b; bidx = i; }   // the old best is now a candidate for second
    else if (b < second) second = b;
    if (s < second) second = s;
  }
  const long long at = (long long)gz * P.qcap + gx;
  best_idx[at] = bidx; best_dist[at] = best; second_dist[at] = second;
}
