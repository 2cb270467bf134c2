// ba_common.h -- declarations shared by the two window-BA kernels (ba_window.cu: shared-memory resident solver for the
// windows the tracker produces; ba_kernels.cu: L2-resident solver for oversized windows, and the host side of both).
#pragma once
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "ba_math.h"
#include "ctx.h"
#include "lm_device.h"

using namespace vb;

// a / x for a positive normal x without the library's special-case subroutine (see rsqrt_pos)
__device__ __forceinline__ double div_pos(double a, double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  const double q = a * y;
  return fma(fma(-x, q, a), y, q);
}


// Huber rho(e) and weight rho'(e) like vb::huber, with sqrt and the division replaced by one reciprocal square root
// (MUFU.RSQ64H + one third-order correction, ~1 ulp): s = e rsqrt(e), delta / s = delta rsqrt(e)
__device__ __forceinline__ void huber_fast(double e, double delta, double& rho0, double& w) {
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho0 = e; w = 1.0; }
  else {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(e));
    const double c = fma(e, -(y * y), 1.0);
    y = fma(fma(c, 0.375, 0.5), y * c, y);
    rho0 = 2 * (e * y) * delta - dsqr;
    w = delta * y;
  }
}

#define BA_THREADS 256
#define BA_MAX_W 24
#define BA_MAX_CLUSTER 16
#define BA_PCHUNK 8   // max chunks per pose in the pose-block reduction
#define BA_MAX_JOBS (BA_MAX_W * (BA_MAX_W + 1) / 2 + BA_MAX_W)   // Schur jobs: pose pairs + gradient jobs
#define BA_MAX_SLOTS (BA_MAX_CLUSTER * BA_THREADS / 32)            // partial sums per job: at most one per warp
#define BA_JOB_PAD 2   // fixed cost of entering a job (decode, prefix scan, flush), in units, for the load balance

struct BaArgs {
  int W, P, M;
  int max_iterations;
  int t_detail;   // debug: extra phase time stamps (VIDO_BA_TIMING=2)
  double info_cam, info_3d, d_cam, d_3d, gain_threshold;
  // graph (device)
  const float* poses_f32;   // [W][16]
  const float* rel_f32;     // [W-1][16]
  const float* points_f32;  // [P][3]  (sorted order)
  const int* obs_pose;      // [M]  pose-major observation index -> pose
  const int* obs_point;     // [M]  -> point
  const float* obs_xyz;     // [3][M] struct-of-arrays
  const int* pt_len;        // [P]
  const int* pt_first;      // [P]
  const int* grp_start;     // [W+1]
  const int* cnt_gt;        // [W][W+1]: #points of group f with track length > L
  const int* off;           // [W][W+1]: offset of group f inside pose p's range
  const int* pose_base;     // [W+1]
  // state
  Pose* X;        // [2][W]
  Pose* Zinv;     // [W-1]
  double* pts;    // [2][P][3]
  // system
  double* hl;     // [P]     point block = hl * I3
  double* bl;     // [P][3]
  // linearisation buffers, one per state buffer (index = state index): the trial state's observation pass fills the
  // other one, an accepted trial makes it current
  double* ow;     // [2][M]    robust weight * information of every observation   } the 6x3 block Hpl = ow * [-I | Q(zc)]^T R^T
  double* ozc;    // [2][3][M] point in the camera frame (struct-of-arrays)         } is never formed (see phase_schur_units)
  double* og;     // [2][3][M] ow * R * error: the point gradient is -sum og
  double* ohb;    // [4][M] per observation: its point's hl and bl (written by phase_blocks, read coalesced by the gradient jobs)
  double* Hpp;    // [W][36] diagonal blocks (points + odometry)
  double* Hoff;   // [W-1][36] blocks (i, i+1)
  double* bp;     // [W][6]
  double* ppart;  // [2][W][BA_PCHUNK][28] partial pose blocks (27 sums)
  double* spart;  // [jobs][BA_MAX_SLOTS][16] partial Schur moment sums
  double* pmax;   // [W] max |diagonal| of every pose block
  double* xp;     // [6W]
  double* part;   // [BA_MAX_CLUSTER][4]: chi2, scale, max point diag, max pose diag
  double* cinfo;  // [4]: pose part of the scale, solver failure flag
  double* seJ;    // [2][W][72]
  double* seE;    // [2][W][8]
  LmCtl* ctl_out;
  LmRec* rec;
  unsigned long long* t_phase;  // [24]
  float* out_poses; float* out_rel; float* out_points;
  // shared-memory resident kernel (ba_window.cu)
  double* wmom;   // [workers][jobs][16] Schur moment sums of every worker CTA
  double* wpsum;  // [2][workers][W][28] pose-block sums of every worker CTA, per linearisation buffer
  double* eH;     // [2][W][120] odometry edges: w Ji^T Ji | w Ji^T Jj | w Jj^T Jj | -w Ji^T e | -w Jj^T e, per linearisation buffer
  double* Sg;     // [(np+1) x (np+1)] reduced camera system (lower triangle, row np = right-hand side), assembled by all CTAs
  int capO, capPt, capT;  // observation / point / Schur-term capacity of a worker CTA (shared-memory carving)
  // chaining inside the kernel: values this window shares with the solve queued in front of it come from that solve's output
  // block (chain[i] = index of pose i / sorted point W + n in the previous problem, -1 = value from the input block)
  const int* chain;
  const float *prev_poses, *prev_rel, *prev_points;
  // the output block is also written to its pinned host mirror (h_delta = host block - device block, in bytes) followed by a
  // completion word: the host polls it, so that NOTHING but solver kernels sits on the solver's stream and consecutive
  // solves can be launched with programmatic dependent launch (prologue of solve k+1 behind the tail of solve k)
  const char* out_base;   // device output block
  char* h_out;            // its pinned host mirror (nullptr: no mirror)
  int out_bytes;
  volatile int* h_flag;
  int seq;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum 36 per-lane values over the warp with 54 shuffles instead of 180: two halving steps (lane pairs trade halves of
// their value sets), then a butterfly inside the 8-lane groups.  Afterwards v[j], j < 9, of lane L holds the sum of
// value 18*((L>>4)&1) + 9*((L>>3)&1) + j.  Fixed order, deterministic.
__device__ __forceinline__ void warp_sum36(double* v) {
  const int lane = threadIdx.x & 31;
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
#pragma unroll
  for (int k = 0; k < 18; k++) {
    const double send = hi16 ? v[k] : v[k + 18], keep = hi16 ? v[k + 18] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const double send = hi8 ? v[k] : v[k + 9], keep = hi8 ? v[k + 9] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 9; k++) {
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 4);
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 2);
    v[k] += __shfl_xor_sync(0xffffffffu, v[k], 1);
  }
}

// Same idea for 16 values with 16 shuffles: afterwards v[0] of lane L holds the sum of value 8*b4 + 4*b3 + 2*b2 + b1
// (b_i = bit i of L).
__device__ __forceinline__ void warp_sum16(double* v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 8, m = 16; h >= 1; h >>= 1, m >>= 1) {
    const bool hi = (lane & m) != 0;
#pragma unroll
    for (int k = 0; k < h; k++) {
      const double send = hi ? v[k] : v[k + h], keep = hi ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, m);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reduction of NV values per thread; results in sm[0..NV) (valid after the call for every thread)
template <int NV, bool MAX>
__device__ __forceinline__ void block_reduce(double* v, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = MAX ? warp_max(v[k]) : warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < NV; k++) sm[warp * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = sm[threadIdx.x];
    for (int w = 1; w < nw; w++) s = MAX ? fmax(s, sm[w * NV + threadIdx.x]) : s + sm[w * NV + threadIdx.x];
    sm[threadIdx.x] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned long long gtime();
__device__ __forceinline__ void ba_tick(unsigned long long* tp, int slot);
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ba_tick(unsigned long long* tp, int slot) {
  if (tp) { const unsigned long long t = gtime(); tp[slot] += t - tp[15]; tp[15] = t; }
}

