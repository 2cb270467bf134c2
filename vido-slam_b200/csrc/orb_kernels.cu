// orb_kernels.cu -- ORB/FAST front-end of the VIDO-SLAM hot path, hand-written for sm_100a.
//
// Replaces ORBextractor::operator() of the reference (src/ORBextractor.cc:1034-1105):
//   K1 pyr_down_kernel      <- ComputePyramid            (:1107-1132; cv::resize INTER_LINEAR 8U fixed point)
//   K2 fast_cells_kernel    <- per-cell cv::FAST calls   (:779-819; thr 20 with fallback 7, cell-local NMS)
//   K3 octree_kernel        <- DistributeOctTree         (:529-753) + DivideNode (:471-527)
//   K4 finalize_kernel      <- IC_Angle (:67-94), cv::fastAtan2, coordinate scaling (:1096-1099), level concat
// All kernels are batched over frames (the front-end has no inter-frame dependency), FAST tiles are
// staged into shared memory by TMA (cp.async.bulk.tensor), and every stage is integer/bit exact.
#include <cfloat>
#include <cmath>
#include <cstring>

#include "ctx.h"

// =====================================================================================================
// K0: 3-channel -> gray (cv::cvtColor 8U, OpenCV 4.x coefficients: (B*3735 + G*19235 + R*9798 + 16384) >> 15)
// =====================================================================================================
__global__ void bgr2gray_kernel(const uint8_t* __restrict__ src, size_t sfs, int sstride, uint8_t* __restrict__ dst,
                                size_t dfs, int dstride, int w, int h, int c0, int c1, int c2) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  int b = blockIdx.z;
  if (x >= w) return;
  const uint8_t* p = src + b * sfs + (size_t)y * sstride + 3 * x;
  int v = (p[0] * c0 + p[1] * c1 + p[2] * c2 + 16384) >> 15;
  dst[b * dfs + (size_t)y * dstride + x] = (uint8_t)v;
}

// =====================================================================================================
// K1: one pyramid level from the previous one (cv::resize INTER_LINEAR, fixed point: horizontal pass with 11-bit
// coefficients, vertical pass (by * (r >> 4)) >> 16, + 2 >> 2).  A thread owns 4 adjacent output columns and walks down a
// strip of PYR_ROWS output rows: the column tables (source offset, coefficients) stay in registers for the whole strip, and
// the horizontal interpolation of a source row is reused when the next output row's upper source row is this row's lower
// one (5 rows out of 6 at scale 1.2) -- 2 byte loads and ~10 integer instructions per output pixel instead of 4 loads + 2
// table loads and ~50 instructions when every pixel was computed from scratch (round 1: issue-bound at 16 % of HBM peak).
// =====================================================================================================
#define PYR_ROWS 8   // strip height at large batches; smaller when the launch would not fill the machine (pyr_rows below)
__global__ void __launch_bounds__(128) pyr_down_kernel(const uint8_t* __restrict__ src, int spitch, size_t sfs, int sw,
                                                       int sh, uint8_t* __restrict__ dst, int dpitch, size_t dfs, int dw,
                                                       int dh, const int32_t* __restrict__ xofs,
                                                       const short2* __restrict__ xa, const int32_t* __restrict__ yofs,
                                                       const short2* __restrict__ ya, int rows) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y0 = blockIdx.y * rows;
  const int b = blockIdx.z;
  if (x4 >= dw) return;
  int sx0[4], sx1[4], a0[4], a1[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int x = min(x4 + i, dw - 1);
    const int sx = xofs[x];
    const short2 ax = xa[x];
    sx0[i] = sx; sx1[i] = min(sx + 1, sw - 1);
    a0[i] = ax.x; a1[i] = ax.y;
  }
  const uint8_t* const sb = src + b * sfs;
  uint8_t* const db = dst + b * dfs;
  auto hrow = [&](int sy, int* r) {   // horizontal pass of source row sy for the 4 columns
    const uint8_t* p = sb + (size_t)sy * spitch;
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = (p[sx0[i]] * a0[i] + p[sx1[i]] * a1[i]) >> 4;
  };
  int lo[4] = {0, 0, 0, 0}, lo_row = -1;   // horizontal result of the lower source row of the previous output row
  const int yend = min(y0 + rows, dh);
  for (int y = y0; y < yend; y++) {
    const int sy = yofs[y], sy1 = min(sy + 1, sh - 1);
    const short2 by = ya[y];
    int r0[4], r1[4];
    if (sy == lo_row) {
#pragma unroll
      for (int i = 0; i < 4; i++) r0[i] = lo[i];
    } else hrow(sy, r0);
    if (sy1 == sy) {
#pragma unroll
      for (int i = 0; i < 4; i++) r1[i] = r0[i];
    } else hrow(sy1, r1);
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int v = (((by.x * r0[i]) >> 16) + ((by.y * r1[i]) >> 16) + 2) >> 2;
      v = min(max(v, 0), 255);
      packed |= (x4 + i < dw ? (uint32_t)v : 0u) << (8 * i);
      lo[i] = r1[i];
    }
    lo_row = sy1;
    *reinterpret_cast<uint32_t*>(db + (size_t)y * dpitch + x4) = packed;  // pitch is a multiple of 64
  }
}

// =====================================================================================================
// K2: FAST-9/16 score + cell-local strict NMS + ordered compaction, one CTA per cv::FAST cell.
// =====================================================================================================
struct TmapPack {
  CUtensorMap m[VIDO_MAX_LEVELS];
};

struct OrbLevelView {
  const uint8_t* base;
  size_t frame_stride;
  int w, h, pitch;
};

struct FastParams {
  OrbLevelView view[VIDO_MAX_LEVELS];
  int ini_thr, min_thr;
  int cells_per_frame, slots_per_frame;
  int level_cell_begin[VIDO_MAX_LEVELS + 1];
  int boxW[VIDO_MAX_LEVELS], boxH[VIDO_MAX_LEVELS];
  int nlevels;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Threshold-independent FAST score: pixel is a corner at threshold t  <=>  score >= t.
// score = max( max_k min_{j<9} d[k+j], max_k min_{j<9} -d[k+j] ) - 1, d = ring - centre.  Returns 0 below thrMin.
__device__ __forceinline__ int fast_score(const uint8_t* t, int bw, int thrMin) {
  const int c = t[0];
  const int hi = c + thrMin, lo = c - thrMin;
  int r[16];
  r[0] = t[3 * bw];
  r[8] = t[-3 * bw];
  r[4] = t[3];
  r[12] = t[-3];
  // a 9-arc contains at least one pixel of every opposite pair
  bool pb = (r[0] > hi || r[8] > hi) && (r[4] > hi || r[12] > hi);
  bool pd = (r[0] < lo || r[8] < lo) && (r[4] < lo || r[12] < lo);
  if (!pb && !pd) return 0;
  r[1] = t[3 * bw + 1];
  r[2] = t[2 * bw + 2];
  r[3] = t[bw + 3];
  r[5] = t[-bw + 3];
  r[6] = t[-2 * bw + 2];
  r[7] = t[-3 * bw + 1];
  r[9] = t[-3 * bw - 1];
  r[10] = t[-2 * bw - 2];
  r[11] = t[-bw - 3];
  r[13] = t[bw - 3];
  r[14] = t[2 * bw - 2];
  r[15] = t[3 * bw - 1];
  uint32_t mb = 0, md = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    mb |= (uint32_t)(r[k] > hi) << k;
    md |= (uint32_t)(r[k] < lo) << k;
  }
  auto has9 = [](uint32_t m) {
    uint32_t m32 = m | (m << 16);
    uint32_t a = m32 & (m32 >> 1);
    a &= a >> 2;
    a &= a >> 4;
    a &= m32 >> 8;
    return (a & 0xffffu) != 0;
  };
  if (!has9(mb) && !has9(md)) return 0;
  // exact score (rare path)
  int d[16];
#pragma unroll
  for (int k = 0; k < 16; k++) d[k] = r[k] - c;
  int mn2[16], mx2[16], mn4[16], mx4[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    mn2[k] = min(d[k], d[(k + 1) & 15]);
    mx2[k] = max(d[k], d[(k + 1) & 15]);
  }
#pragma unroll
  for (int k = 0; k < 16; k++) {
    mn4[k] = min(mn2[k], mn2[(k + 2) & 15]);
    mx4[k] = max(mx2[k], mx2[(k + 2) & 15]);
  }
  int best_b = -512, best_d = 512;
#pragma unroll
  for (int k = 0; k < 16; k++) {
    int mn9 = min(min(mn4[k], mn4[(k + 4) & 15]), d[(k + 8) & 15]);
    int mx9 = max(max(mx4[k], mx4[(k + 4) & 15]), d[(k + 8) & 15]);
    best_b = max(best_b, mn9);
    best_d = min(best_d, mx9);
  }
  return max(best_b, -best_d) - 1;
}

// The same score for TWO horizontally adjacent pixels at once, packed as signed 16-bit pairs: sm_100 has single-instruction
// 2 x 16-bit min / max with three inputs (SASS VIMNMX.S16x2 / VIMNMX3.S16x2; the byte-wise __vminu4 family is emulated, ~6
// instructions each).  d = ring - centre fits 16 bits; min over a 9-arc = min3 of three min3 (32 instructions for the 16
// arcs), max over the arcs = a max3 tree.  Returns score(pixel 0) | score(pixel 1) << 16, each 0 below thrMin -- the same
// value as fast_score (corner at threshold t  <=>  score >= t).  ~85 instructions per pixel instead of ~400 when most
// pixels pass the 4-point pre-test (textured images); the pre-test still rejects flat pixel pairs after 10 byte loads.
__device__ __forceinline__ uint32_t fast_score_pair(const uint8_t* t, int bw, int thrMin) {
  auto pk = [](uint32_t a, uint32_t b) { return a | (b << 16); };
  const uint32_t c = pk(t[0], t[1]);
  const uint32_t nc = __vneg2(c);
  // compass points first: a 9-arc contains at least one pixel of every opposite pair
  const uint32_t e0 = t[3 * bw], e1 = t[3 * bw + 1], w0 = t[-3 * bw], w1 = t[-3 * bw + 1];
  const uint32_t r3 = t[3], r4 = t[4], l3 = t[-3], l2 = t[-2];
  uint32_t d[16];
  d[0] = __vadd2(pk(e0, e1), nc);
  d[8] = __vadd2(pk(w0, w1), nc);
  d[4] = __vadd2(pk(r3, r4), nc);
  d[12] = __vadd2(pk(l3, l2), nc);
  {
    const uint32_t mb = __vmins2(__vmaxs2(d[0], d[8]), __vmaxs2(d[4], d[12]));   // > thr  <=> a bright arc is possible
    const uint32_t md = __vmaxs2(__vmins2(d[0], d[8]), __vmins2(d[4], d[12]));   // < -thr <=> a dark arc is possible
    const int b0 = (short)(mb & 0xffffu), b1 = (short)(mb >> 16), k0 = (short)(md & 0xffffu), k1 = (short)(md >> 16);
    if (b0 <= thrMin && b1 <= thrMin && k0 >= -thrMin && k1 >= -thrMin) return 0u;
  }
  {
    const uint32_t a = t[3 * bw - 1], b = t[3 * bw + 2];                 // row +3: dx -1 .. 2
    d[15] = __vadd2(pk(a, e0), nc); d[1] = __vadd2(pk(e1, b), nc);
    const uint32_t g = t[-3 * bw - 1], h = t[-3 * bw + 2];               // row -3
    d[9] = __vadd2(pk(g, w0), nc); d[7] = __vadd2(pk(w1, h), nc);
    const uint32_t p2 = t[2 * bw + 2], p3 = t[2 * bw + 3], q2 = t[2 * bw - 2], q1 = t[2 * bw - 1];   // row +2
    d[2] = __vadd2(pk(p2, p3), nc); d[14] = __vadd2(pk(q2, q1), nc);
    const uint32_t u2 = t[-2 * bw + 2], u3 = t[-2 * bw + 3], v2 = t[-2 * bw - 2], v1 = t[-2 * bw - 1];   // row -2
    d[6] = __vadd2(pk(u2, u3), nc); d[10] = __vadd2(pk(v2, v1), nc);
    const uint32_t i3 = t[bw + 3], i4 = t[bw + 4], j3 = t[bw - 3], j2 = t[bw - 2];                     // row +1
    d[3] = __vadd2(pk(i3, i4), nc); d[13] = __vadd2(pk(j3, j2), nc);
    const uint32_t m3 = t[-bw + 3], m4 = t[-bw + 4], n3 = t[-bw - 3], n2 = t[-bw - 2];                 // row -1
    d[5] = __vadd2(pk(m3, m4), nc); d[11] = __vadd2(pk(n3, n2), nc);
  }
  uint32_t lo3[16], hi3[16];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    lo3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    hi3[k] = __vimax3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
  }
  uint32_t bb = 0x80008000u, bd = 0x7fff7fffu;   // max over the arcs of their min / min over the arcs of their max
#pragma unroll
  for (int k = 0; k < 16; k += 2) {
    const uint32_t mnA = __vimin3_s16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
    const uint32_t mnB = __vimin3_s16x2(lo3[k + 1], lo3[(k + 4) & 15], lo3[(k + 7) & 15]);
    bb = __vimax3_s16x2(bb, mnA, mnB);
    const uint32_t mxA = __vimax3_s16x2(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
    const uint32_t mxB = __vimax3_s16x2(hi3[k + 1], hi3[(k + 4) & 15], hi3[(k + 7) & 15]);
    bd = __vimin3_s16x2(bd, mxA, mxB);
  }
  const int sb0 = (short)(bb & 0xffffu), sb1 = (short)(bb >> 16), sd0 = (short)(bd & 0xffffu), sd1 = (short)(bd >> 16);
  int s0 = max(sb0, -sd0) - 1, s1 = max(sb1, -sd1) - 1;
  s0 = s0 >= thrMin ? s0 : 0;
  s1 = s1 >= thrMin ? s1 : 0;
  return (uint32_t)s0 | ((uint32_t)s1 << 16);
}

#define FAST_MAXDIM 80  // max ROI edge (wCell+6); cells are 30..59 px by construction (src/ORBextractor.cc:773-776)
#define FAST_THREADS 256

__global__ void __launch_bounds__(FAST_THREADS) fast_cells_kernel(const CUtensorMap* __restrict__ maps,
                                                                  const OrbCell* __restrict__ cells, FastParams P,
                                                                  uint32_t* __restrict__ slots,
                                                                  int32_t* __restrict__ cell_count) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t rowmask[FAST_MAXDIM][3][2];  // [row][32-px chunk][0: survivors >= min thr, 1: >= ini thr]
  __shared__ int rowoff[FAST_MAXDIM][2];
  __shared__ int total[2];

  const int cell_id = blockIdx.x;
  const int b = blockIdx.y;
  const OrbCell cell = cells[cell_id];
  int level = 0;
#pragma unroll
  for (int l = 1; l < VIDO_MAX_LEVELS; l++)
    if (l < P.nlevels && cell_id >= P.level_cell_begin[l]) level = l;
  const int bw = P.boxW[level], bh = P.boxH[level];
  // TMA needs the innermost start coordinate 16-byte aligned (measured: unaligned x0 -> illegal instruction),
  // so the box starts at x0 & ~15 and the ROI sits xoff bytes into each tile row.
  const int xoff = cell.x0 & 15, x0a = cell.x0 - xoff;
  uint8_t* tile0 = smem;                              // bw x bh, filled by TMA
  const uint8_t* tile = tile0 + xoff;
  uint8_t* score = smem + ((bw * bh + 127) & ~127);   // (rh) x (rw) scores, 0 outside the detection zone
  const int rw = cell.rw, rh = cell.rh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

#ifndef VIDO_NO_TMA
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bw * bh) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(smem_u32(tile0)),
        "l"(maps + level), "r"(x0a), "r"(cell.y0), "r"(b), "r"(smem_u32(&bar))
        : "memory");
  }
  // (no zero fill of the score array: the score pass below writes the one-pixel ring around the detection zone, which is all
  //  the non-maximum suppression reads outside it)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
          : "=r"(done)
          : "r"(smem_u32(&bar))
          : "memory");
    }
  }
  __syncthreads();

#else
  {  // debug build without TMA: plain cooperative tile load (zero fill outside the level)
    const OrbLevelView lvw = P.view[level];
    const uint8_t* img = lvw.base + (size_t)b * lvw.frame_stride;
    for (int i = tid; i < bw * bh; i += FAST_THREADS) {
      int ty = i / bw, tx = i - ty * bw;
      int gx = x0a + tx, gy = cell.y0 + ty;
      tile0[i] = (gx < lvw.w && gy < lvw.h) ? img[(size_t)gy * lvw.pitch + gx] : 0;
    }
    for (int i = tid; i < rw * rh; i += FAST_THREADS) score[i] = 0;
    (void)bar;
  }
  __syncthreads();
#endif
  // ---- scores of the detection zone [3, rw-3) x [3, rh-3), zeros on the one-pixel ring around it
  const int iw = rw - 6, ih = rh - 6;
  if (iw > 0 && ih > 0) {
    const int npair = (iw + 1) >> 1;
    for (int y = warp - 1; y <= ih; y += FAST_THREADS / 32) {
      uint8_t* srow = score + (y + 3) * rw + 3;
      if (y < 0 || y >= ih) {
        for (int x = lane - 1; x <= iw; x += 32) srow[x] = 0;
        continue;
      }
      if (lane == 0) { srow[-1] = 0; srow[iw] = 0; }
      const uint8_t* trow = tile + (y + 3) * bw + 3;
      for (int p = lane; p < npair; p += 32) {   // two pixels per lane (the odd one out of an odd-width zone is computed, not stored)
        const int x = 2 * p;
        const uint32_t s2 = fast_score_pair(trow + x, bw, P.min_thr);
        srow[x] = (uint8_t)(s2 & 0xffu);
        if (x + 1 < iw) srow[x + 1] = (uint8_t)(s2 >> 16);
      }
    }
  }
  __syncthreads();

  // ---- strict 3x3 NMS; survivors as per-row bit masks (row-major order is the reference's output order)
  for (int y = warp; y < max(ih, 0); y += FAST_THREADS / 32) {
    int c7 = 0, c20 = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      int x = lane + 32 * ch;
      bool keep = false;
      int s = 0;
      if (x < iw) {
        const uint8_t* sp = score + (y + 3) * rw + (x + 3);
        s = sp[0];
        if (s > 0) {
          int m = max(max(max(sp[-1], sp[1]), max(sp[-rw], sp[rw])),
                      max(max(sp[-rw - 1], sp[-rw + 1]), max(sp[rw - 1], sp[rw + 1])));
          keep = s > m;
        }
      }
      uint32_t m7 = __ballot_sync(0xffffffffu, keep);
      uint32_t m20 = __ballot_sync(0xffffffffu, keep && s >= P.ini_thr);
      if (lane == 0) {
        rowmask[y][ch][0] = m7;
        rowmask[y][ch][1] = m20;
      }
      c7 += __popc(m7);
      c20 += __popc(m20);
    }
    if (lane == 0) {
      rowoff[y][0] = c7;
      rowoff[y][1] = c20;
    }
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the row counts (ih <= 74)
    int base7 = 0, base20 = 0;
    for (int y0 = 0; y0 < max(ih, 0); y0 += 32) {
      int y = y0 + lane;
      int v7 = (y < ih) ? rowoff[y][0] : 0, v20 = (y < ih) ? rowoff[y][1] : 0;
      int s7 = v7, s20 = v20;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t7 = __shfl_up_sync(0xffffffffu, s7, o), t20 = __shfl_up_sync(0xffffffffu, s20, o);
        if (lane >= o) { s7 += t7; s20 += t20; }
      }
      if (y < ih) {
        rowoff[y][0] = base7 + s7 - v7;
        rowoff[y][1] = base20 + s20 - v20;
      }
      base7 += __shfl_sync(0xffffffffu, s7, 31);
      base20 += __shfl_sync(0xffffffffu, s20, 31);
    }
    if (lane == 0) { total[0] = base7; total[1] = base20; }
  }
  __syncthreads();
  // cv::FAST(thr=iniThFAST) first; only if it found nothing, the whole cell is redone at minThFAST (:802-806)
  const int sel = (total[1] > 0) ? 1 : 0;
  const int n = min(total[sel], cell.slot_cap);
  uint32_t* out = slots + (size_t)b * P.slots_per_frame + cell.slot_base;
  for (int y = warp; y < max(ih, 0); y += FAST_THREADS / 32) {
    int off = rowoff[y][sel];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      uint32_t m = rowmask[y][ch][sel];
      int x = lane + 32 * ch;
      if ((m >> lane) & 1u) {
        int pos = off + __popc(m & ((1u << lane) - 1u));
        if (pos < cell.slot_cap) {
          int s = score[(y + 3) * rw + (x + 3)];
          out[pos] = ((uint32_t)(x + 3 + cell.offx) << 20) | ((uint32_t)(y + 3 + cell.offy) << 8) | (uint32_t)s;
        }
      }
      off += __popc(m);
    }
  }
  if (tid == 0) cell_count[(size_t)b * P.cells_per_frame + cell_id] = n;
}

// =====================================================================================================
// K3: quad-tree culling (DistributeOctTree), one CTA per (level, frame).
// The node list is an array kept in list order.  A node owns a contiguous segment of a permutation
// array (its keys, in the reference's relative order); splitting = stable 4-way partition of the segment
// into the other permutation buffer.  Full passes split all multi-key nodes in parallel (one warp per
// node); the final "largest first" phase is sequential like the reference (warp 0).
// Tie-break among equal-size nodes (reference: heap address, src/ORBextractor.cc:671-675): later-created first.
// =====================================================================================================
struct ONode {
  short x0, x1, y0, y1;
  int start;
  int cnt_buf;  // (count << 1) | permutation buffer index
};

struct SortEnt {
  int cnt, serial, id;
};

struct OctParams {
  int nlevels;
  int cells_per_frame, slots_per_frame;
  int level_cell_begin[VIDO_MAX_LEVELS + 1];
  int quota[VIDO_MAX_LEVELS];
  int nIni[VIDO_MAX_LEVELS];
  float hX[VIDO_MAX_LEVELS];
  int ymax[VIDO_MAX_LEVELS];  // maxBorderY - minBorderY
  int out_base[VIDO_MAX_LEVELS], out_cap[VIDO_MAX_LEVELS];
  int cand_cap[VIDO_MAX_LEVELS];
  size_t cand_base[VIDO_MAX_LEVELS];
  int out_slots_per_frame;
  size_t keys_per_frame;
  int smem_keys;               // key capacity of the shared-memory key/perm arrays
  int lcap, tcap, qcap, ccap;  // list / child / sequence / cell-offset array capacities
};

#define OCT_THREADS 256
#define OCT_DEAD 0xffffu

__device__ __forceinline__ int key_x(uint32_t k) { return (int)(k >> 20); }
__device__ __forceinline__ int key_y(uint32_t k) { return (int)((k >> 8) & 0xfffu); }
__device__ __forceinline__ int key_s(uint32_t k) { return (int)(k & 0xffu); }

// stable 4-way partition of a node's segment by one warp; child counts returned in all lanes
__device__ __forceinline__ void warp_split(const uint32_t* keys, const uint16_t* src, uint16_t* dst, int start, int count,
                                           int mx, int my, int c[4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  c[0] = c[1] = c[2] = c[3] = 0;
  for (int i0 = 0; i0 < count; i0 += 32) {
    int i = i0 + lane;
    int cls = -1;
    if (i < count) {
      uint32_t k = keys[src[start + i]];
      cls = (key_x(k) < mx) ? ((key_y(k) < my) ? 0 : 2) : ((key_y(k) < my) ? 1 : 3);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) c[q] += __popc(__ballot_sync(0xffffffffu, cls == q));
  }
  int o[4];
  o[0] = start;
  o[1] = o[0] + c[0];
  o[2] = o[1] + c[1];
  o[3] = o[2] + c[2];
  for (int i0 = 0; i0 < count; i0 += 32) {
    int i = i0 + lane;
    int cls = -1;
    uint16_t id = 0;
    if (i < count) {
      id = src[start + i];
      uint32_t k = keys[id];
      cls = (key_x(k) < mx) ? ((key_y(k) < my) ? 0 : 2) : ((key_y(k) < my) ? 1 : 3);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      uint32_t m = __ballot_sync(0xffffffffu, cls == q);
      if (cls == q) dst[o[q] + __popc(m & lt)] = id;
      o[q] += __popc(m);
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void split_geometry(const ONode& n, int& mx, int& my) {
  mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2)   (DivideNode, :473-474)
  my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
}

__device__ __forceinline__ ONode child_of(const ONode& n, int k, int mx, int my, int start, int cnt) {
  ONode c;
  c.x0 = (k & 1) ? (short)mx : n.x0;
  c.x1 = (k & 1) ? n.x1 : (short)mx;
  c.y0 = (k & 2) ? (short)my : n.y0;
  c.y1 = (k & 2) ? n.y1 : (short)my;
  c.start = start;
  c.cnt_buf = (cnt << 1) | ((n.cnt_buf & 1) ^ 1);
  return c;
}

__global__ void __launch_bounds__(OCT_THREADS) octree_kernel(OctParams P, const OrbCell* __restrict__ cells,
                                                             const uint32_t* __restrict__ slots,
                                                             const int32_t* __restrict__ cell_count,
                                                             uint32_t* __restrict__ g_keys, uint16_t* __restrict__ g_perm,
                                                             uint32_t* __restrict__ level_out,
                                                             int32_t* __restrict__ level_cnt, int32_t* __restrict__ err) {
  extern __shared__ __align__(16) uint8_t osm[];
  const int level = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = OCT_THREADS / 32;
  const uint32_t lt = (1u << lane) - 1u;
  const int N = P.quota[level];

  // ---- carve shared memory
  uint8_t* sp = osm;
  ONode* cur = (ONode*)sp;            sp += sizeof(ONode) * P.lcap;
  ONode* nxt = (ONode*)sp;            sp += sizeof(ONode) * P.lcap;
  ONode* tmpc = (ONode*)sp;           sp += sizeof(ONode) * P.tcap;
  SortEnt* vspA = (SortEnt*)sp;       sp += sizeof(SortEnt) * P.lcap;
  SortEnt* vspB = (SortEnt*)sp;       sp += sizeof(SortEnt) * P.lcap;
  int* cell_off = (int*)sp;           sp += sizeof(int) * P.ccap;
  uint16_t* explist = (uint16_t*)sp;  sp += sizeof(uint16_t) * P.lcap;
  uint16_t* seqpos = (uint16_t*)sp;   sp += sizeof(uint16_t) * P.lcap;
  uint16_t* freest = (uint16_t*)sp;   sp += sizeof(uint16_t) * P.lcap;
  uint16_t* fin = (uint16_t*)sp;      sp += sizeof(uint16_t) * P.lcap;
  uint16_t* seq = (uint16_t*)sp;      sp += sizeof(uint16_t) * P.qcap;
  sp = (uint8_t*)(((uintptr_t)sp + 15) & ~(uintptr_t)15);
  uint32_t* s_keys = (uint32_t*)sp;   sp += sizeof(uint32_t) * P.smem_keys;
  uint16_t* s_perm = (uint16_t*)sp;

  __shared__ int sh_n, sh_S, sh_E, sh_state, sh_nv, sh_fin;

  int32_t* lc = level_cnt + (size_t)b * 2 * P.nlevels;
  // ---- 1. exclusive scan of this level's cell counts
  const int cb = P.level_cell_begin[level], ncell = P.level_cell_begin[level + 1] - cb;
  const int32_t* cc = cell_count + (size_t)b * P.cells_per_frame + cb;
  if (warp == 0) {
    int base = 0;
    for (int c0 = 0; c0 < ncell; c0 += 32) {
      int c = c0 + lane;
      int v = (c < ncell) ? cc[c] : 0, s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      if (c < ncell) cell_off[c] = base + s - v;
      base += __shfl_sync(0xffffffffu, s, 31);
    }
    if (lane == 0) sh_n = base;
  }
  __syncthreads();
  const int n = sh_n;
  if (n == 0 || n > 65535) {
    if (tid == 0) {
      lc[level] = 0;
      lc[P.nlevels + level] = n;
      if (n > 65535) atomicOr(err, 1);
    }
    return;
  }
  const bool in_smem = n <= P.smem_keys;
  uint32_t* keys = in_smem ? s_keys : g_keys + (size_t)b * P.keys_per_frame + P.cand_base[level];
  uint16_t* perm[2];
  perm[0] = in_smem ? s_perm : g_perm + 2 * ((size_t)b * P.keys_per_frame + P.cand_base[level]);
  perm[1] = perm[0] + (in_smem ? P.smem_keys : P.cand_cap[level]);

  // ---- 2. gather candidates in reference order (cell-row-major, then FAST scan order)
  for (int i = warp; i < ncell; i += NW) {
    const int cnt = cc[i], off = cell_off[i];
    const uint32_t* src = slots + (size_t)b * P.slots_per_frame + cells[cb + i].slot_base;
    for (int j = lane; j < cnt; j += 32) keys[off + j] = src[j];
  }
  __syncthreads();

  // ---- 3. root nodes: vpIniNodes[kp.pt.x / hX] (:557-561); nIni <= 32 (checked on the host)
  if (warp == 0) {
    const int nIni = P.nIni[level];
    const float hX = P.hX[level];
    int mycnt = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane, cls = -1;
      if (i < n) cls = min((int)__fdiv_rn((float)key_x(keys[i]), hX), nIni - 1);
      for (int c = 0; c < nIni; c++) {
        uint32_t m = __ballot_sync(0xffffffffu, cls == c);
        if (lane == c) mycnt += __popc(m);
      }
    }
    int incl = mycnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int mystart = incl - mycnt;
    int myoff = mystart;
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane, cls = -1;
      if (i < n) cls = min((int)__fdiv_rn((float)key_x(keys[i]), hX), nIni - 1);
      for (int c = 0; c < nIni; c++) {
        uint32_t m = __ballot_sync(0xffffffffu, cls == c);
        int oc = __shfl_sync(0xffffffffu, myoff, c);
        if (cls == c) perm[0][oc + __popc(m & lt)] = (uint16_t)i;
        if (lane == c) myoff += __popc(m);
      }
    }
    // list = non-empty roots in order (empty ones are erased, :567-577)
    bool ne = lane < nIni && mycnt > 0;
    uint32_t m = __ballot_sync(0xffffffffu, ne);
    if (ne) {
      ONode r;
      r.x0 = (short)(int)__fmul_rn(hX, (float)lane);
      r.x1 = (short)(int)__fmul_rn(hX, (float)(lane + 1));
      r.y0 = 0;
      r.y1 = (short)P.ymax[level];
      r.start = mystart;
      r.cnt_buf = (mycnt << 1) | 0;
      cur[__popc(m & lt)] = r;
    }
    if (lane == 0) {
      sh_S = __popc(m);
      sh_state = 0;
    }
  }
  __syncthreads();

  // ---- 4. full passes (:583-649)
  int state = 0;
  while (state == 0) {
    const int S = sh_S;
    if (warp == 0) {
      int E = 0;
      for (int i0 = 0; i0 < S; i0 += 32) {
        int i = i0 + lane;
        bool ex = i < S && (cur[i].cnt_buf >> 1) > 1;
        uint32_t m = __ballot_sync(0xffffffffu, ex);
        if (ex) explist[E + __popc(m & lt)] = (uint16_t)i;
        E += __popc(m);
      }
      if (lane == 0) sh_E = E;
    }
    __syncthreads();
    const int E = sh_E;
    if (4 * E > P.tcap) {  // cannot happen by construction (E <= N/3 after the first pass); guard anyway
      if (tid == 0) atomicOr(err, 2);
      break;
    }
    for (int r = warp; r < E; r += NW) {
      const ONode nd = cur[explist[r]];
      int mx, my, c[4];
      split_geometry(nd, mx, my);
      const int bi = nd.cnt_buf & 1;
      warp_split(keys, perm[bi], perm[bi ^ 1], nd.start, nd.cnt_buf >> 1, mx, my, c);
      if (lane < 4) {
        int st = nd.start;
        for (int k = 0; k < lane; k++) st += c[k];
        tmpc[4 * r + lane] = child_of(nd, lane, mx, my, st, c[lane]);
      }
    }
    __syncthreads();
    if (warp == 0) {
      // children in reverse creation order go to the front (push_front), single-key nodes keep their order behind
      int Mch = 0, nExp = 0;
      for (int q0 = 0; q0 < 4 * E; q0 += 32) {
        int q = q0 + lane;
        int cnt = (q < 4 * E) ? (tmpc[q].cnt_buf >> 1) : 0;
        Mch += __popc(__ballot_sync(0xffffffffu, cnt > 0));
        nExp += __popc(__ballot_sync(0xffffffffu, cnt > 1));
      }
      const int newS = Mch + (S - E);
      if (newS > P.lcap) {
        if (lane == 0) { atomicOr(err, 4); sh_state = 3; }
      } else {
        int rk = 0, rv = 0;
        for (int q0 = 0; q0 < 4 * E; q0 += 32) {
          int q = q0 + lane;
          int cnt = (q < 4 * E) ? (tmpc[q].cnt_buf >> 1) : 0;
          uint32_t m1 = __ballot_sync(0xffffffffu, cnt > 0);
          uint32_t m2 = __ballot_sync(0xffffffffu, cnt > 1);
          if (cnt > 0) {
            int dest = Mch - 1 - (rk + __popc(m1 & lt));
            nxt[dest] = tmpc[q];
            if (cnt > 1) vspA[rv + __popc(m2 & lt)] = {cnt, q, dest};
          }
          rk += __popc(m1);
          rv += __popc(m2);
        }
        int rn = 0;
        for (int i0 = 0; i0 < S; i0 += 32) {
          int i = i0 + lane;
          bool nm = i < S && (cur[i].cnt_buf >> 1) == 1;
          uint32_t m = __ballot_sync(0xffffffffu, nm);
          if (nm) nxt[Mch + rn + __popc(m & lt)] = cur[i];
          rn += __popc(m);
        }
        if (lane == 0) {
          sh_S = newS;
          sh_nv = nExp;
          if (newS >= N || newS == S) sh_state = 2;
          else if (newS + 3 * nExp > N) sh_state = 1;
          else sh_state = 0;
        }
      }
    }
    __syncthreads();
    state = sh_state;
    if (state == 3) break;
    ONode* t = cur; cur = nxt; nxt = t;
    __syncthreads();
  }
  if (state == 3 || state == 0) {  // capacity guard tripped
    if (tid == 0) { lc[level] = 0; lc[P.nlevels + level] = n; }
    return;
  }

  // ---- 5. sequential "largest node first" phase (:650-717), warp 0
  if (state == 1 && warp == 0) {
    int S = sh_S, nv = sh_nv;
    int seqLen = S, size = S, freeTop = 0, poolTop = S, serial = 1 << 28;
    for (int i = lane; i < S; i += 32) {
      seq[S - 1 - i] = (uint16_t)i;  // seq end = list front
      seqpos[i] = (uint16_t)(S - 1 - i);
    }
    __syncwarp();
    bool done = false, bad = false;
    while (!done) {
      const int prevSize = size;
      // rank sort ascending by (count, creation serial); processed from the back
      for (int e = lane; e < nv; e += 32) {
        const SortEnt a = vspA[e];
        int rank = 0;
        for (int f = 0; f < nv; f++) {
          const SortEnt o = vspA[f];
          rank += (o.cnt < a.cnt) || (o.cnt == a.cnt && o.serial < a.serial);
        }
        vspB[rank] = a;
      }
      __syncwarp();
      int nvNew = 0;
      for (int j = nv - 1; j >= 0; j--) {
        const SortEnt ent = vspB[j];
        const ONode nd = cur[ent.id];
        int mx, my, c[4];
        split_geometry(nd, mx, my);
        const int bi = nd.cnt_buf & 1;
        warp_split(keys, perm[bi], perm[bi ^ 1], nd.start, nd.cnt_buf >> 1, mx, my, c);
        if (seqLen + 4 > P.qcap) {  // compact the sequence array (drop erased entries)
          if (lane == 0) {
            int w = 0;
            for (int r = 0; r < seqLen; r++)
              if (seq[r] != OCT_DEAD) { seq[w] = seq[r]; seqpos[seq[w]] = (uint16_t)w; w++; }
            seqLen = w;
          }
          seqLen = __shfl_sync(0xffffffffu, seqLen, 0);
        }
        int st = nd.start;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (c[k] > 0) {
            int nid;
            if (freeTop > 0) nid = freest[--freeTop];
            else nid = poolTop++;
            if (nid >= P.lcap) { bad = true; nid = P.lcap - 1; }
            if (lane == 0) {
              cur[nid] = child_of(nd, k, mx, my, st, c[k]);
              seq[seqLen] = (uint16_t)nid;
              seqpos[nid] = (uint16_t)seqLen;
              if (c[k] > 1) vspA[nvNew] = {c[k], serial, nid};
            }
            seqLen++;
            size++;
            if (c[k] > 1) { nvNew++; serial++; }
          }
          st += c[k];
          __syncwarp();
        }
        if (lane == 0) {
          seq[seqpos[ent.id]] = OCT_DEAD;
          freest[freeTop] = (uint16_t)ent.id;
        }
        freeTop++;
        size--;
        __syncwarp();
        if (size >= N || bad) break;
      }
      if (size >= N || size == prevSize || bad) done = true;
      nv = nvNew;
    }
    if (bad && lane == 0) atomicOr(err, 8);
    // final list order: sequence back to front, erased entries skipped
    int cntf = 0;
    for (int j0 = seqLen - 1; j0 >= 0; j0 -= 32) {
      int j = j0 - lane;
      bool alive = j >= 0 && seq[j] != OCT_DEAD;
      uint32_t m = __ballot_sync(0xffffffffu, alive);
      if (alive && cntf + __popc(m & lt) < P.lcap) fin[cntf + __popc(m & lt)] = seq[j];
      cntf += __popc(m);
    }
    if (lane == 0) sh_fin = min(cntf, P.lcap);
  } else if (state == 2) {
    const int S = sh_S;
    for (int i = tid; i < S; i += OCT_THREADS) fin[i] = (uint16_t)i;
    if (tid == 0) sh_fin = S;
  }
  __syncthreads();

  // ---- 6. best key of every node (:720-737): maximum response, first one wins
  int nfin = sh_fin;
  if (nfin > P.out_cap[level]) {
    if (tid == 0) atomicOr(err, 16);
    nfin = P.out_cap[level];
  }
  uint32_t* out = level_out + (size_t)b * P.out_slots_per_frame + P.out_base[level];
  for (int t = tid; t < nfin; t += OCT_THREADS) {
    const ONode nd = cur[fin[t]];
    const uint16_t* pp = perm[nd.cnt_buf & 1] + nd.start;
    const int cnt = nd.cnt_buf >> 1;
    uint32_t best = keys[pp[0]];
    for (int k = 1; k < cnt; k++) {
      uint32_t kk = keys[pp[k]];
      if (key_s(kk) > key_s(best)) best = kk;
    }
    out[t] = ((uint32_t)(key_x(best) + VIDO_MINB) << 20) | ((uint32_t)(key_y(best) + VIDO_MINB) << 8) | (uint32_t)key_s(best);
  }
  if (tid == 0) {
    lc[level] = nfin;
    lc[P.nlevels + level] = n;
  }
}

// =====================================================================================================
// K4: orientation (IC_Angle + cv::fastAtan2), coordinate scaling and level concatenation.  One warp per keypoint.
// =====================================================================================================
struct FinParams {
  int nlevels;
  int out_base[VIDO_MAX_LEVELS + 1];
  int out_slots_per_frame;
  int pitch[VIDO_MAX_LEVELS];
  size_t frame_stride[VIDO_MAX_LEVELS], base[VIDO_MAX_LEVELS];
  float scale[VIDO_MAX_LEVELS];
  int cap_per_frame;
};

__constant__ int c_umax[16];

__device__ __forceinline__ float dev_fast_atan2(float y, float x) {
  const float k = (float)(180.0 / 3.14159265358979323846);
  const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
  const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

__global__ void __launch_bounds__(256) finalize_kernel(FinParams P, const uint8_t* __restrict__ pyr,
                                                       const uint32_t* __restrict__ level_out,
                                                       const int32_t* __restrict__ level_cnt,
                                                       vido_keypoint* __restrict__ out, int32_t* __restrict__ n_out,
                                                       int32_t* __restrict__ err) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // slot in the per-frame level-output array
  const int32_t* lc = level_cnt + (size_t)b * 2 * P.nlevels;
  int total = 0, level = -1, idx = 0, before = 0;
  for (int l = 0; l < P.nlevels; l++) {
    int c = lc[l];
    if (g >= P.out_base[l] && g < P.out_base[l + 1]) { level = l; idx = g - P.out_base[l]; before = total; }
    total += c;
  }
  if (g == 0 && lane == 0) {
    n_out[b] = min(total, P.cap_per_frame);
    if (total > P.cap_per_frame) atomicOr(err, 32);
  }
  if (level < 0 || idx >= lc[level]) return;
  const int pos = before + idx;
  if (pos >= P.cap_per_frame) return;
  const uint32_t k = level_out[(size_t)b * P.out_slots_per_frame + g];
  const int x = key_x(k), y = key_y(k), s = key_s(k);
  const uint8_t* img = pyr + P.base[level] + (size_t)b * P.frame_stride[level];
  const int pitch = P.pitch[level];
  // m10 = sum u*I, m01 = sum v*I over the circular patch of radius 15 (IC_Angle)
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int v = lane - 15;
    const int d = c_umax[abs(v)];
    const uint8_t* row = img + (size_t)(y + v) * pitch + x;
    int sum = 0;
    for (int u = -d; u <= d; u++) {
      int val = row[u];
      sum += val;
      m10 += u * val;
    }
    m01 = v * sum;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    m01 += __shfl_xor_sync(0xffffffffu, m01, o);
  }
  if (lane == 0) {
    vido_keypoint kp;
    const float sc = P.scale[level];
    float fx = (float)x, fy = (float)y;
    if (level != 0) { fx = __fmul_rn(fx, sc); fy = __fmul_rn(fy, sc); }
    kp.x = fx;
    kp.y = fy;
    kp.size = (float)(int)__fmul_rn(31.f, sc);  // PATCH_SIZE*mvScaleFactor[level] truncated (:826)
    kp.angle = dev_fast_atan2((float)m01, (float)m10);
    kp.response = (float)s;
    kp.octave = level;
    out[(size_t)b * P.cap_per_frame + pos] = kp;
  }
}

// =====================================================================================================
// host side
// =====================================================================================================
static inline int cv_round_f(float v) { return (int)lrintf(v); }

static void axis_table(int srcDim, int dstDim, std::vector<int32_t>& ofs, std::vector<int16_t>& a) {
  ofs.resize(dstDim);
  a.resize(2 * (size_t)dstDim);
  const double scale = (double)srcDim / dstDim;
  for (int d = 0; d < dstDim; d++) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= srcDim - 1) { s = srcDim - 1; f = 0.f; }
    ofs[d] = s;
    a[2 * d] = (int16_t)lrintf((1.f - f) * 2048.f);
    a[2 * d + 1] = (int16_t)lrintf(f * 2048.f);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int orb_setup(vido_ctx* ctx) {
  const vido_config& c = ctx->cfg;
  const int B = c.max_batch;
  ctx->nlevels = c.nlevels;
  if (c.nlevels < 1 || c.nlevels > VIDO_MAX_LEVELS) { ctx->err = "nlevels must be 1..8"; return VIDO_ERR_ARG; }
  if (c.width > 4000 || c.height > 4000 || c.width < 64 || c.height < 64) { ctx->err = "image size unsupported"; return VIDO_ERR_ARG; }

  // ---- level geometry (ORBextractor ctor :400-437, ComputePyramid :1111-1112)
  float scale = 1.f;
  float factor = 1.0f / c.scale_factor;
  float nDesired = c.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)c.nlevels));
  int sumF = 0;
  size_t pyr_off = 0;
  int cell_total = 0, slot_total = 0, out_total = 0;
  size_t key_total = 0;
  ctx->cells.clear();
  for (int l = 0; l < c.nlevels; l++) {
    OrbLevel& L = ctx->lv[l];
    if (l > 0) scale = scale * c.scale_factor;
    L.scale = scale;
    float inv = 1.0f / scale;
    L.w = cv_round_f((float)c.width * inv);
    L.h = cv_round_f((float)c.height * inv);
    L.pitch = (int)align_up(L.w, 64);
    L.frame_stride = align_up((size_t)L.pitch * L.h, 256);
    L.base = pyr_off;
    pyr_off += L.frame_stride * B;
    if (l < c.nlevels - 1) {
      L.quota = cv_round_f(nDesired);
      sumF += L.quota;
      nDesired *= factor;
    } else {
      L.quota = std::max(c.nfeatures - sumF, 0);
    }
    // cell grid (ComputeKeyPointsOctTree :763-776)
    L.maxBX = L.w - VIDO_EDGE + 3;
    L.maxBY = L.h - VIDO_EDGE + 3;
    const float width = (float)(L.maxBX - VIDO_MINB), height = (float)(L.maxBY - VIDO_MINB);
    L.nCols = (int)(width / 30.f);
    L.nRows = (int)(height / 30.f);
    L.cell_begin = cell_total;
    L.ncells = 0;
    L.cand_cap = 0;
    L.boxW = 16;
    L.boxH = 8;
    if (L.nCols > 0 && L.nRows > 0) {
      L.wCell = (int)ceil(width / L.nCols);
      L.hCell = (int)ceil(height / L.nRows);
      for (int i = 0; i < L.nRows; i++) {
        const float iniY = (float)(VIDO_MINB + i * L.hCell);
        float maxY = iniY + L.hCell + 6;
        if (iniY >= L.maxBY - 3) continue;
        if (maxY > L.maxBY) maxY = (float)L.maxBY;
        for (int j = 0; j < L.nCols; j++) {
          const float iniX = (float)(VIDO_MINB + j * L.wCell);
          float maxX = iniX + L.wCell + 6;
          if (iniX >= L.maxBX - 6) continue;
          if (maxX > L.maxBX) maxX = (float)L.maxBX;
          OrbCell cell;
          cell.x0 = (int)iniX;
          cell.y0 = (int)iniY;
          cell.rw = (int)maxX - cell.x0;
          cell.rh = (int)maxY - cell.y0;
          cell.offx = j * L.wCell;
          cell.offy = i * L.hCell;
          const int iw = std::max(cell.rw - 6, 0), ih = std::max(cell.rh - 6, 0);
          cell.slot_cap = std::max(((iw + 1) / 2) * ((ih + 1) / 2), 1);  // strict NMS: no two 8-adjacent survivors
          cell.slot_base = slot_total;
          slot_total += cell.slot_cap;
          L.cand_cap += cell.slot_cap;
          if (cell.rw > FAST_MAXDIM || cell.rh > FAST_MAXDIM) { ctx->err = "FAST cell larger than 80 px"; return VIDO_ERR_ARG; }
          L.boxW = std::max(L.boxW, (int)align_up(cell.rw + 15, 16));  // room for the 16-byte start alignment
          L.boxH = std::max(L.boxH, cell.rh);
          ctx->cells.push_back(cell);
          L.ncells++;
        }
      }
    } else {
      L.wCell = L.hCell = 0;
    }
    cell_total += L.ncells;
    // quad-tree roots (DistributeOctTree :533-535)
    const int dX = L.maxBX - VIDO_MINB, dY = L.maxBY - VIDO_MINB;
    L.nIni = (dY > 0) ? (int)roundf((float)dX / dY) : 0;
    if (L.nIni > 32) { ctx->err = "aspect ratio too wide for the quad-tree root partition (nIni > 32)"; return VIDO_ERR_ARG; }
    if (L.nIni < 1 && L.ncells > 0) { ctx->err = "aspect ratio too tall (nIni == 0)"; return VIDO_ERR_ARG; }
    L.hX = L.nIni > 0 ? (float)dX / L.nIni : 1.f;
    L.cand_base = key_total;
    key_total += align_up(L.cand_cap, 8);
    L.out_base = out_total;
    L.out_cap = L.quota + 4 * std::max(L.nIni, 1) + 4;
    out_total += L.out_cap;
  }
  ctx->cells_per_frame = cell_total;
  ctx->slots_per_frame = (int)align_up(slot_total, 4);
  ctx->octree_keys_per_frame = key_total;
  ctx->out_slots_per_frame = out_total;
  ctx->pyr_bytes = pyr_off;
  ctx->kp_cap = c.nfeatures + 64;

  // ---- allocations
  VIDO_CUDA(cudaMalloc(&ctx->d_pyr, ctx->pyr_bytes));
  VIDO_CUDA(cudaMemsetAsync(ctx->d_pyr, 0, ctx->pyr_bytes, ctx->stream));
  VIDO_CUDA(cudaMalloc(&ctx->d_cells, sizeof(OrbCell) * std::max<size_t>(ctx->cells.size(), 1)));
  VIDO_CUDA(cudaMemcpyAsync(ctx->d_cells, ctx->cells.data(), sizeof(OrbCell) * ctx->cells.size(), cudaMemcpyHostToDevice, ctx->stream));
  VIDO_CUDA(cudaMalloc(&ctx->d_slots, sizeof(uint32_t) * (size_t)ctx->slots_per_frame * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_cell_count, sizeof(int32_t) * (size_t)std::max(cell_total, 1) * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_oct_keys, sizeof(uint32_t) * std::max<size_t>(key_total, 1) * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_oct_perm, sizeof(uint16_t) * 2 * std::max<size_t>(key_total, 1) * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_level_out, sizeof(uint32_t) * (size_t)out_total * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_level_cnt, sizeof(int32_t) * 2 * c.nlevels * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_kp, sizeof(vido_keypoint) * (size_t)ctx->kp_cap * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_nkp, sizeof(int32_t) * B));
  ctx->in_pitch = (int)align_up(c.width * 3, 64);
  VIDO_CUDA(cudaMalloc(&ctx->d_in, (size_t)ctx->in_pitch * c.height * B));
  VIDO_CUDA(cudaMalloc(&ctx->d_err, sizeof(int32_t)));
  VIDO_CUDA(cudaMemsetAsync(ctx->d_err, 0, sizeof(int32_t), ctx->stream));

  // ---- resize tables
  for (int l = 1; l < c.nlevels; l++) {
    std::vector<int32_t> xo, yo;
    std::vector<int16_t> xa, ya;
    axis_table(ctx->lv[l - 1].w, ctx->lv[l].w, xo, xa);
    axis_table(ctx->lv[l - 1].h, ctx->lv[l].h, yo, ya);
    VIDO_CUDA(cudaMalloc(&ctx->d_xofs[l], xo.size() * 4));
    VIDO_CUDA(cudaMalloc(&ctx->d_xa[l], xa.size() * 2));
    VIDO_CUDA(cudaMalloc(&ctx->d_yofs[l], yo.size() * 4));
    VIDO_CUDA(cudaMalloc(&ctx->d_ya[l], ya.size() * 2));
    VIDO_CUDA(cudaMemcpy(ctx->d_xofs[l], xo.data(), xo.size() * 4, cudaMemcpyHostToDevice));
    VIDO_CUDA(cudaMemcpy(ctx->d_xa[l], xa.data(), xa.size() * 2, cudaMemcpyHostToDevice));
    VIDO_CUDA(cudaMemcpy(ctx->d_yofs[l], yo.data(), yo.size() * 4, cudaMemcpyHostToDevice));
    VIDO_CUDA(cudaMemcpy(ctx->d_ya[l], ya.data(), ya.size() * 2, cudaMemcpyHostToDevice));
  }
  // ---- umax (ORBextractor ctor :444-459)
  {
    int umax[17] = {0};
    const int HP = 15;
    int v, v0, vmax = (int)floor(HP * sqrt(2.f) / 2 + 1), vmin = (int)ceil(HP * sqrt(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
    VIDO_CUDA(cudaMemcpyToSymbol(c_umax, umax, sizeof(int) * 16));
  }
  // ---- TMA descriptors of the pyramid levels: dims {w, h, B}, box {boxW, boxH, 1}, u8, OOB -> 0
  {
    PFN_encodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VIDO_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { ctx->err = "cuTensorMapEncodeTiled unavailable"; return VIDO_ERR_CUDA; }
    for (int l = 0; l < c.nlevels; l++) {
      const OrbLevel& L = ctx->lv[l];
      cuuint64_t gdim[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)B};
      cuuint64_t gstr[2] = {(cuuint64_t)L.pitch, (cuuint64_t)L.frame_stride};
      cuuint32_t box[3] = {(cuuint32_t)L.boxW, (cuuint32_t)L.boxH, 1};
      cuuint32_t estr[3] = {1, 1, 1};
      CUresult r = encode(&ctx->tmap[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ctx->d_pyr + L.base, gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(level %d) failed: %d", l, (int)r);
        ctx->err = buf;
        return VIDO_ERR_CUDA;
      }
    }
    // descriptors live in global memory (64-byte aligned); the FAST kernel indexes them by level
    VIDO_CUDA(cudaMalloc(&ctx->d_tmap, sizeof(CUtensorMap) * VIDO_MAX_LEVELS));
    VIDO_CUDA(cudaMemcpy(ctx->d_tmap, ctx->tmap, sizeof(CUtensorMap) * c.nlevels, cudaMemcpyHostToDevice));
  }
  // ---- kernel attributes
  {
    int maxBox = 0;
    for (int l = 0; l < c.nlevels; l++) maxBox = std::max(maxBox, (int)align_up((size_t)ctx->lv[l].boxW * ctx->lv[l].boxH, 128));
    size_t fast_smem = maxBox + FAST_MAXDIM * FAST_MAXDIM + 256;
    VIDO_CUDA(cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
  }
  {
    int N0 = 0, nIni0 = 1, ccap = 32;
    for (int l = 0; l < c.nlevels; l++) {
      N0 = std::max(N0, ctx->lv[l].quota);
      nIni0 = std::max(nIni0, ctx->lv[l].nIni);
      ccap = std::max(ccap, ctx->lv[l].ncells);
    }
    int lcap = N0 + 4 * nIni0 + 16;
    int tcap = (4 * N0) / 3 + 4 * nIni0 + 32;
    int qcap = 4 * lcap;
    size_t fixed = sizeof(ONode) * (2 * (size_t)lcap + tcap) + sizeof(SortEnt) * 2 * lcap + sizeof(int) * ccap +
                   sizeof(uint16_t) * (4 * (size_t)lcap + qcap) + 64;
    int smem_keys = 6144;
    while (fixed + (size_t)smem_keys * 8 > 100 * 1024 && smem_keys > 1024) smem_keys -= 512;
    if (lcap > 65000 || fixed + (size_t)smem_keys * 8 > 220 * 1024) { ctx->err = "nFeatures too large for the quad-tree kernel"; return VIDO_ERR_ARG; }
    ctx->oct_smem_keys = smem_keys;
    ctx->oct_smem_bytes = fixed + (size_t)smem_keys * 8;
    VIDO_CUDA(cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->oct_smem_bytes));
  }
  VIDO_CUDA(cudaStreamSynchronize(ctx->stream));
  return VIDO_OK;
}

void orb_teardown(vido_ctx* ctx) {
  cudaFree(ctx->d_pyr); cudaFree(ctx->d_cells); cudaFree(ctx->d_slots); cudaFree(ctx->d_cell_count);
  cudaFree(ctx->d_oct_keys); cudaFree(ctx->d_oct_perm); cudaFree(ctx->d_level_out); cudaFree(ctx->d_level_cnt);
  cudaFree(ctx->d_kp); cudaFree(ctx->d_nkp); cudaFree(ctx->d_in); cudaFree(ctx->d_err); cudaFree(ctx->d_tmap);
  for (int l = 0; l < VIDO_MAX_LEVELS; l++) {
    cudaFree(ctx->d_xofs[l]); cudaFree(ctx->d_xa[l]); cudaFree(ctx->d_yofs[l]); cudaFree(ctx->d_ya[l]);
  }
}

int orb_bgr_to_gray(vido_ctx* ctx, const uint8_t* d_bgr, int nframes, size_t frame_stride, int stride, uint8_t* d_gray,
                    size_t gray_frame_stride, int gray_stride) {
  const vido_config& c = ctx->cfg;
  // cv::cvtColor BGR2GRAY: B*3735 + G*19235 + R*9798; RGB2GRAY swaps the outer coefficients
  const int cB = 3735, cG = 19235, cR = 9798;
  dim3 grid((c.width + 255) / 256, c.height, nframes);
  bgr2gray_kernel<<<grid, 256, 0, ctx->stream>>>(d_bgr, frame_stride, stride, d_gray, gray_frame_stride, gray_stride,
                                                 c.width, c.height, c.rgb ? cR : cB, cG, c.rgb ? cB : cR);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}

int orb_run(vido_ctx* ctx, const uint8_t* d_gray, int nframes, size_t frame_stride, int stride, vido_keypoint* d_out,
            int cap_per_frame, int32_t* d_n_out) {
  const vido_config& c = ctx->cfg;
  if (nframes < 1 || nframes > c.max_batch) { ctx->err = "nframes exceeds max_batch"; return VIDO_ERR_ARG; }
  cudaStream_t st = ctx->stream;
  const int B = nframes;
  ctx->last_batch = B;
  // level 0 = input copied into the pitched pyramid buffer (ComputePyramid level 0; the reflect border of the
  // reference is never read by FAST cells or IC_Angle, so it is not materialised)
  {
    const OrbLevel& L0 = ctx->lv[0];
    cudaMemcpy3DParms p = {};
    p.srcPtr = make_cudaPitchedPtr((void*)d_gray, stride, c.width, frame_stride / stride);
    p.dstPtr = make_cudaPitchedPtr(ctx->d_pyr + L0.base, L0.pitch, c.width, L0.frame_stride / L0.pitch);
    p.extent = make_cudaExtent(c.width, c.height, B);
    p.kind = cudaMemcpyDeviceToDevice;
    if (frame_stride % stride == 0 && L0.frame_stride % L0.pitch == 0) {
      VIDO_CUDA(cudaMemcpy3DAsync(&p, st));
    } else {
      for (int b = 0; b < B; b++)
        VIDO_CUDA(cudaMemcpy2DAsync(ctx->d_pyr + L0.base + b * L0.frame_stride, L0.pitch, d_gray + b * frame_stride, stride,
                                    c.width, c.height, cudaMemcpyDeviceToDevice, st));
    }
  }
  for (int l = 1; l < c.nlevels; l++) {
    const OrbLevel& S = ctx->lv[l - 1];
    const OrbLevel& D = ctx->lv[l];
    // strip height: as tall as possible (table and row reuse) while the launch still has ~4 resident waves of threads
    const long long quads = (long long)((D.w + 3) / 4) * D.h * B;   // 4-pixel groups of the level
    int rows = (int)(quads / ((long long)ctx->num_sms * 2048 * 2));
    rows = rows < 1 ? 1 : (rows > PYR_ROWS ? PYR_ROWS : rows);
    dim3 grid((D.w + 511) / 512, (D.h + rows - 1) / rows, B);
    pyr_down_kernel<<<grid, 128, 0, st>>>(ctx->d_pyr + S.base, S.pitch, S.frame_stride, S.w, S.h, ctx->d_pyr + D.base,
                                          D.pitch, D.frame_stride, D.w, D.h, ctx->d_xofs[l], (const short2*)ctx->d_xa[l],
                                          ctx->d_yofs[l], (const short2*)ctx->d_ya[l], rows);
    ctx->launches++;
  }
  if (ctx->cells_per_frame > 0) {
    FastParams P;
    memset(&P, 0, sizeof P);
    for (int l = 0; l < c.nlevels; l++) {
      P.level_cell_begin[l] = ctx->lv[l].cell_begin;
      P.boxW[l] = ctx->lv[l].boxW;
      P.boxH[l] = ctx->lv[l].boxH;
      P.view[l] = {ctx->d_pyr + ctx->lv[l].base, ctx->lv[l].frame_stride, ctx->lv[l].w, ctx->lv[l].h, ctx->lv[l].pitch};
    }
    P.level_cell_begin[c.nlevels] = ctx->cells_per_frame;
    P.ini_thr = c.ini_th_fast;
    P.min_thr = c.min_th_fast;
    P.cells_per_frame = ctx->cells_per_frame;
    P.slots_per_frame = ctx->slots_per_frame;
    P.nlevels = c.nlevels;
    int maxBox = 0;
    for (int l = 0; l < c.nlevels; l++) maxBox = std::max(maxBox, (int)align_up((size_t)ctx->lv[l].boxW * ctx->lv[l].boxH, 128));
    size_t smem = maxBox + FAST_MAXDIM * FAST_MAXDIM + 256;
    dim3 grid(ctx->cells_per_frame, B);
    fast_cells_kernel<<<grid, FAST_THREADS, smem, st>>>(ctx->d_tmap, ctx->d_cells, P, ctx->d_slots, ctx->d_cell_count);
    ctx->launches++;
  }
  {
    OctParams P;
    memset(&P, 0, sizeof P);
    P.nlevels = c.nlevels;
    P.cells_per_frame = ctx->cells_per_frame;
    P.slots_per_frame = ctx->slots_per_frame;
    int N0 = 0, nIni0 = 1, ccap = 32;
    for (int l = 0; l < c.nlevels; l++) {
      const OrbLevel& L = ctx->lv[l];
      P.level_cell_begin[l] = L.cell_begin;
      P.quota[l] = L.quota;
      P.nIni[l] = L.nIni;
      P.hX[l] = L.hX;
      P.ymax[l] = L.maxBY - VIDO_MINB;
      P.out_base[l] = L.out_base;
      P.out_cap[l] = L.out_cap;
      P.cand_cap[l] = L.cand_cap;
      P.cand_base[l] = L.cand_base;
      N0 = std::max(N0, L.quota);
      nIni0 = std::max(nIni0, L.nIni);
      ccap = std::max(ccap, L.ncells);
    }
    P.level_cell_begin[c.nlevels] = ctx->cells_per_frame;
    P.out_slots_per_frame = ctx->out_slots_per_frame;
    P.keys_per_frame = ctx->octree_keys_per_frame;
    P.smem_keys = ctx->oct_smem_keys;
    P.lcap = N0 + 4 * nIni0 + 16;
    P.tcap = (4 * N0) / 3 + 4 * nIni0 + 32;
    P.qcap = 4 * P.lcap;
    P.ccap = ccap;
    dim3 grid(c.nlevels, B);
    octree_kernel<<<grid, OCT_THREADS, ctx->oct_smem_bytes, st>>>(P, ctx->d_cells, ctx->d_slots, ctx->d_cell_count,
                                                                  ctx->d_oct_keys, ctx->d_oct_perm, ctx->d_level_out,
                                                                  ctx->d_level_cnt, ctx->d_err);
    ctx->launches++;
  }
  {
    FinParams P;
    memset(&P, 0, sizeof P);
    P.nlevels = c.nlevels;
    for (int l = 0; l < c.nlevels; l++) {
      const OrbLevel& L = ctx->lv[l];
      P.out_base[l] = L.out_base;
      P.pitch[l] = L.pitch;
      P.frame_stride[l] = L.frame_stride;
      P.base[l] = L.base;
      P.scale[l] = L.scale;
    }
    P.out_base[c.nlevels] = ctx->out_slots_per_frame;
    P.out_slots_per_frame = ctx->out_slots_per_frame;
    P.cap_per_frame = cap_per_frame;
    dim3 grid((ctx->out_slots_per_frame + 7) / 8, B);
    finalize_kernel<<<grid, 256, 0, st>>>(P, ctx->d_pyr, ctx->d_level_out, ctx->d_level_cnt, d_out, d_n_out, ctx->d_err);
    ctx->launches++;
  }
  VIDO_CUDA(cudaGetLastError());
  return VIDO_OK;
}
