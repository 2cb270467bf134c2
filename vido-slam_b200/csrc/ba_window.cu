// ba_window.cu -- sliding-window graph optimisation (Optimizer::PartialBatchOptimization) with the whole problem resident
// in the SHARED MEMORY of one thread-block cluster.  Same mathematics and the same LM decision sequence as the L2-resident
// kernel of ba_kernels.cu (which stays for windows that do not fit); what changes is where the data lives and how the
// work is cut, because that kernel was bound by the latency of its phases, not by arithmetic or bandwidth:
//
//   * POINT partition.  CTA 0 is the solver, CTAs 1..C-1 are workers.  The points (sorted by first pose, length
//     descending) are dealt round-robin to the workers; a worker keeps ALL observations of its points -- measurement,
//     robust weight, camera-frame point, for the current and the trial linearisation -- in its shared memory (~105 B per
//     observation), so every term of the Schur complement  sum_l Hpl(p1,l) Hpl(p2,l)^T / (hl + lambda)  is local to one CTA
//     and is read at shared-memory latency instead of L2 latency.
//   * Per trial four cluster barriers (the old kernel: six, around L2 round trips):
//       workers: Schur moment sums of their points, 4 lanes per pose pair, no atomics            -> global partials
//       B1 | all CTAs: sum the partials of "their" pose pairs over the workers, assemble the 6x6 blocks, store them into
//            CTA 0's shared memory (DSMEM)
//       B2 | CTA 0: dense Cholesky (ba_chol.h: register-resident trailing matrix, FP64 tensor-core tiles), pose increments
//       B3 | workers: point back-substitution, trial points, observation pass at the trial state (robust chi2 + the
//            linearisation the next iteration needs if the trial is accepted); CTA 0: odometry edges at the trial state
//       B4 | every CTA: identical LM bookkeeping from the same partial sums (lm_device.h)
//
// Reference code replaced: see ba_kernels.cu (src/Optimizer.cc:220-362,806,1056-1142; g2o sparse_optimizer.cpp:354-427,
// optimization_algorithm_levenberg.cpp:61-189, block_solver.hpp:502-560, linear_solver_csparse.h:108-141).
#include <cooperative_groups.h>

#include "ba_chol.h"
#include "ba_common.h"

namespace cg = cooperative_groups;

#define BW_THREADS BC_THREADS
#define BW_NPAIRS_MAX (BA_MAX_W * (BA_MAX_W + 1) / 2)
#define BW_TAB (BA_MAX_W * (BA_MAX_W + 1))
#define BW_RSLOTS ((BW_NPAIRS_MAX + 7) / 8)   // pair jobs per reducer CTA at the smallest cluster (8)
#define BW_PSLOTS ((BA_MAX_W + 7) / 8)

struct BwShared {
  LmCtl ctl;
  int lgrp[BA_MAX_W + 1];     // first local point of group f
  int lcnt[BW_TAB];           // [f][k]: local points of group f with track length > k
  int lbase[BW_TAB];          // [f][k]: first local observation of segment (f, k); pose of the segment = f + k
  int pbase[BA_MAX_W + 1];    // pose-major observation list: first entry of pose p
  unsigned char jp1[BW_NPAIRS_MAX], jp2[BW_NPAIRS_MAX];
  int tstart[BW_NPAIRS_MAX + 1];             // first Schur term of every pair job in this worker's term list
  unsigned short jorder[BW_NPAIRS_MAX];      // pair jobs by descending term count (lane = job: similar trip counts per warp)
  int Mc, Pc;
  Pose X[2][BA_MAX_W];        // poses of both state buffers (every CTA keeps a copy)
  Pose Zinv[BA_MAX_W];        // inverse odometry measurements (CTA 0)
  double xp[6 * BA_MAX_W];    // pose increments of the trial
  double red[64];
  double rmom[BW_RSLOTS][16]; // reducer: summed moments of the owned pair jobs
  double rR12[BW_RSLOTS][9];
  double rps[BW_PSLOTS][28];  // reducer: summed pose-block sums of the owned poses
  double rgr[BW_PSLOTS][6];   // reducer: summed Schur gradient of the owned poses
  float pf32[BA_MAX_W][16];   // output poses (CTA 0)
  int s_bad;
};

// worker-side view of the dynamic shared memory.  Only scalar members: a pointer ARRAY indexed by the run-time buffer index
// would put the whole struct into local memory and turn every access into LDL + generic LD (measured in the first version).
struct BwObs {
  double* lin;      // [2][4][capO]: w | zx | zy | zz per linearisation buffer (robust weight * information, camera-frame point)
  double *gx, *gy, *gz;                   // transient per-observation 3-vectors (gradient terms)
  double* pts;      // [2][7][capPt]: px | py | pz | hl | bx | by | bz per state / linearisation buffer
  double* inv;      // 1 / (hl + lambda) of the trial
  float *mx, *my, *mz;                    // measurement
  unsigned int* t12;                      // Schur terms: observation in the job's first pose | in its second pose << 16
  unsigned short* tj;                     // Schur terms: local point
  unsigned short *opt, *plist;            // local point of an observation; pose-major observation list
  unsigned char* opo;                     // pose of an observation
  unsigned char *pf, *plen;
  int capO, capPt;
  __device__ __forceinline__ double* w(int b) const { return lin + (size_t)b * 4 * capO; }
  __device__ __forceinline__ double* zx(int b) const { return lin + ((size_t)b * 4 + 1) * capO; }
  __device__ __forceinline__ double* zy(int b) const { return lin + ((size_t)b * 4 + 2) * capO; }
  __device__ __forceinline__ double* zz(int b) const { return lin + ((size_t)b * 4 + 3) * capO; }
  __device__ __forceinline__ double* px(int b) const { return pts + (size_t)b * 7 * capPt; }
  __device__ __forceinline__ double* py(int b) const { return pts + ((size_t)b * 7 + 1) * capPt; }
  __device__ __forceinline__ double* pz(int b) const { return pts + ((size_t)b * 7 + 2) * capPt; }
  __device__ __forceinline__ double* hl(int b) const { return pts + ((size_t)b * 7 + 3) * capPt; }
  __device__ __forceinline__ double* bx(int b) const { return pts + ((size_t)b * 7 + 4) * capPt; }
  __device__ __forceinline__ double* by(int b) const { return pts + ((size_t)b * 7 + 5) * capPt; }
  __device__ __forceinline__ double* bz(int b) const { return pts + ((size_t)b * 7 + 6) * capPt; }
};

__host__ __device__ inline size_t bw_worker_bytes(int capO, int capPt, int capT) {
  return (size_t)capO * (11 * 8 + 3 * 4 + 2 * 2 + 1) + (size_t)capPt * (15 * 8 + 2) + (size_t)capT * 6 + 64;
}

__device__ __forceinline__ void bw_carve(char* base, int capO, int capPt, int capT, BwObs& o) {
  double* d = (double*)base;
  o.capO = capO; o.capPt = capPt;
  o.lin = d; d += 8 * (size_t)capO;
  o.gx = d; d += capO; o.gy = d; d += capO; o.gz = d; d += capO;
  o.pts = d; d += 14 * (size_t)capPt;
  o.inv = d; d += capPt;
  float* f = (float*)d;
  o.mx = f; f += capO; o.my = f; f += capO; o.mz = f; f += capO;
  o.t12 = (unsigned int*)f; f += capT;
  unsigned short* u = (unsigned short*)f;
  o.tj = u; u += capT;
  o.opt = u; u += capO; o.plist = u; u += capO;
  unsigned char* c = (unsigned char*)u;
  o.opo = c; c += capO; o.pf = c; c += capPt; o.plen = c;
}

// block-wide sum / max of one value per thread; result valid for every thread
template <bool MAX>
__device__ __forceinline__ double bw_block_reduce(double v, double* sm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double s = sm[0];
  for (int w = 1; w < BW_THREADS / 32; w++) s = MAX ? fmax(s, sm[w]) : s + sm[w];
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// worker: observation pass at state st (robust chi2 + linearisation), see phase_obs / phase_blocks of ba_kernels.cu
//   thread per observation: weight, camera-frame point, gradient term g = w R e;
//   thread per point:       point block h = sum w, gradient b = -sum g;
//   8 lanes per pose:       the 27 sums of the pose block over this worker's observations of the pose.
// Returns the chi2 partial (valid in every thread); hmax = max point block.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bw_obs_pass(const BaArgs& a, BwShared& sh, const BwObs& ob, int st, int wk, int nwk, double& chi_out,
                                            double& hmax_out) {
  const int tid = threadIdx.x, W = a.W, Mc = sh.Mc, Pc = sh.Pc;
  double* const ow = ob.w(st);
  double* const zx = ob.zx(st);
  double* const zy = ob.zy(st);
  double* const zz = ob.zz(st);
  const double* px = ob.px(st);
  const double* py = ob.py(st);
  const double* pz = ob.pz(st);
  double chi = 0;
  for (int o = tid; o < Mc; o += BW_THREADS) {
    const int j = ob.opt[o], p = ob.opo[o];
    const Pose& Xp = sh.X[st][p];
    const double d0 = px[j] - Xp.t[0], d1 = py[j] - Xp.t[1], d2 = pz[j] - Xp.t[2];
    const double z0 = Xp.R[0] * d0 + Xp.R[3] * d1 + Xp.R[6] * d2;
    const double z1 = Xp.R[1] * d0 + Xp.R[4] * d1 + Xp.R[7] * d2;
    const double z2 = Xp.R[2] * d0 + Xp.R[5] * d1 + Xp.R[8] * d2;
    const double e0 = z0 - (double)ob.mx[o], e1 = z1 - (double)ob.my[o], e2 = z2 - (double)ob.mz[o];
    double r0, w;
    huber_fast((e0 * e0 + e1 * e1 + e2 * e2) * a.info_3d, a.d_3d, r0, w);
    chi += r0;
    w *= a.info_3d;
    ow[o] = w; zx[o] = z0; zy[o] = z1; zz[o] = z2;
    ob.gx[o] = w * (Xp.R[0] * e0 + Xp.R[1] * e1 + Xp.R[2] * e2);
    ob.gy[o] = w * (Xp.R[3] * e0 + Xp.R[4] * e1 + Xp.R[5] * e2);
    ob.gz[o] = w * (Xp.R[6] * e0 + Xp.R[7] * e1 + Xp.R[8] * e2);
  }
  __syncthreads();
  double hmax = 0;
  for (int j = tid; j < Pc; j += BW_THREADS) {
    const int f = ob.pf[j], len = ob.plen[j], i = j - sh.lgrp[f];
    const int* lb = sh.lbase + f * (W + 1);
    double h = 0, b0 = 0, b1 = 0, b2 = 0;
    for (int k = 0; k < len; k++) {
      const int o = lb[k] + i;
      h += ow[o]; b0 -= ob.gx[o]; b1 -= ob.gy[o]; b2 -= ob.gz[o];
    }
    ob.hl(st)[j] = h; ob.bx(st)[j] = b0; ob.by(st)[j] = b1; ob.bz(st)[j] = b2;
    hmax = fmax(hmax, h);
  }
  // pose-block sums: the last 8 W threads, 8 lanes per pose (the point loop above occupies the first warps).  Whole warps
  // enter (the shuffles below need every lane); lanes in front of the first pose thread run with an empty range.
  if ((tid | 31) >= BW_THREADS - 8 * W) {
    const int q = tid - (BW_THREADS - 8 * W);
    const bool valid = q >= 0;
    const int p = valid ? q >> 3 : 0, c = q & 7;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    const int t1 = valid ? sh.pbase[p + 1] : 0;
    for (int t = sh.pbase[p] + c; t < t1; t += 8) {
      const int o = ob.plist[t];
      const double w = ow[o], z0 = zx[o], z1 = zy[o], z2 = zz[o];
      const double e[3] = {z0 - (double)ob.mx[o], z1 - (double)ob.my[o], z2 - (double)ob.mz[o]};
      const double qx = 2 * z0, qy = 2 * z1, qz = 2 * z2;
      const double J[3][6] = {{-1, 0, 0, 0, -qz, qy}, {0, -1, 0, qz, 0, -qx}, {0, 0, -1, -qy, qx, 0}};
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        acc[21 + r] -= w * (J[0][r] * e[0] + J[1][r] * e[1] + J[2][r] * e[2]);
#pragma unroll
        for (int cc = r; cc < 6; cc++) acc[idx++] += w * (J[0][r] * J[0][cc] + J[1][r] * J[1][cc] + J[2][r] * J[2][cc]);
      }
    }
    // 27 (padded to 28) sums over the 8 lanes of a pose with 14 + 7 + 4 = 25 shuffles instead of 81: lane pairs trade halves
    // of their value sets at every step.  Afterwards lane c holds the sums 14 b2 + 7 b1 + 4 b0' ... of its slice:
    //   step 1 (xor 4): b2 = c >> 2 keeps values [14 b2, 14 b2 + 14);  step 2 (xor 2): b1 keeps 7 of them;  step 3 (xor 1):
    //   b0 = 0 keeps 4, b0 = 1 keeps 3 (+1 padding)
    double v14[14], v7[7], v4[4];
    {
      const bool h2 = (c & 4) != 0, h1 = (c & 2) != 0, h0 = (c & 1) != 0;
#pragma unroll
      for (int k = 0; k < 14; k++) {
        const double lo = acc[k], hi = (k + 14 < 27) ? acc[k + 14] : 0.0;
        const double send = h2 ? lo : hi, keep = h2 ? hi : lo;
        v14[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int k = 0; k < 7; k++) {
        const double send = h1 ? v14[k] : v14[k + 7], keep = h1 ? v14[k + 7] : v14[k];
        v7[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double lo = v7[k], hi = (k + 4 < 7) ? v7[k + 4] : 0.0;
        const double send = h0 ? lo : hi, keep = h0 ? hi : lo;
        v4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
    }
    if (valid) {
      // lane c holds sums base .. base + n - 1 with base = 14 (c >> 2) + 7 ((c >> 1) & 1) + 4 (c & 1), n = 4 or 3
      double* out = a.wpsum + (((size_t)st * nwk + wk) * W + p) * 28 + 14 * (c >> 2) + 7 * ((c >> 1) & 1) + 4 * (c & 1);
      out[0] = v4[0]; out[1] = v4[1]; out[2] = v4[2];
      if (!(c & 1)) out[3] = v4[3];
    }
  }
  chi_out = bw_block_reduce<false>(chi, sh.red);
  hmax_out = bw_block_reduce<true>(hmax, sh.red + 32);
}

// ---------------------------------------------------------------------------------------------------------
// worker: back-substitution of the points at the current linearisation, trial points, point part of the scale
//   x_l = (bl - sum_o Hpl(o)^T xp) / (hl + lambda),  Hpl^T xp = w R (-x_t + 2 zc x x_r)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double bw_update(const BaArgs& a, BwShared& sh, const BwObs& ob, int cur, double lambda, int failed) {
  const int tid = threadIdx.x, W = a.W, Mc = sh.Mc, Pc = sh.Pc, trial = cur ^ 1;
  if (!failed) {
    const double* ow = ob.w(cur);
    const double* zx = ob.zx(cur);
    const double* zy = ob.zy(cur);
    const double* zz = ob.zz(cur);
    for (int o = tid; o < Mc; o += BW_THREADS) {
      const int p = ob.opo[o];
      const double* R = sh.X[cur][p].R;
      const double* x = sh.xp + 6 * p;
      const double w = ow[o], qx = 2 * zx[o], qy = 2 * zy[o], qz = 2 * zz[o];
      const double g0 = w * (qy * x[5] - qz * x[4] - x[0]), g1 = w * (qz * x[3] - qx * x[5] - x[1]), g2 = w * (qx * x[4] - qy * x[3] - x[2]);
      ob.gx[o] = R[0] * g0 + R[1] * g1 + R[2] * g2;
      ob.gy[o] = R[3] * g0 + R[4] * g1 + R[5] * g2;
      ob.gz[o] = R[6] * g0 + R[7] * g1 + R[8] * g2;
    }
  }
  __syncthreads();
  double scale = 0;
  for (int j = tid; j < Pc; j += BW_THREADS) {
    const double b0 = ob.bx(cur)[j], b1 = ob.by(cur)[j], b2 = ob.bz(cur)[j];
    double x0 = b0, x1 = b1, x2 = b2;
    if (!failed) {
      const int f = ob.pf[j], len = ob.plen[j], i = j - sh.lgrp[f];
      const int* lb = sh.lbase + f * (W + 1);
      double c0 = b0, c1 = b1, c2 = b2;
      for (int k = 0; k < len; k++) {
        const int o = lb[k] + i;
        c0 -= ob.gx[o]; c1 -= ob.gy[o]; c2 -= ob.gz[o];
      }
      const double s = ob.inv[j];
      x0 = s * c0; x1 = s * c1; x2 = s * c2;
    }
    ob.px(trial)[j] = ob.px(cur)[j] + x0;
    ob.py(trial)[j] = ob.py(cur)[j] + x1;
    ob.pz(trial)[j] = ob.pz(cur)[j] + x2;
    scale += x0 * (lambda * x0 + b0) + x1 * (lambda * x1 + b1) + x2 * (lambda * x2 + b2);
  }
  __syncthreads();
  return scale;   // per-thread partial
}

// ---------------------------------------------------------------------------------------------------------
// worker: Schur moment sums of this worker's points (see phase_schur_units of ba_kernels.cu for the algebra).
//   4 adjacent lanes per job stride over the job's FLAT term list (built once per launch; pair jobs sorted by term count so
//   that the 8 jobs of a warp have similar trip counts) and combine their 16 moment sums with a fixed shuffle tree -- no
//   atomics, deterministic.  Measured alternatives: walking the (group, prefix) structure per job (segments hold ~4
//   observations once the points are dealt over 15 CTAs: the lanes idle in the walk), and one lane per job (32 unrelated
//   gathers per load instruction: ~5 shared-memory wavefronts each, the phase became LSU-bound).  With 4 lanes per job the
//   lanes of a job read 4 consecutive terms, i.e. mostly one 32-byte segment per array.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bw_schur(const BaArgs& a, BwShared& sh, const BwObs& ob, int cur, double lambda, int wk) {
  const int tid = threadIdx.x, W = a.W, Pc = sh.Pc;
  const int npairs = W * (W + 1) / 2, njobs = npairs + W;
  for (int j = tid; j < Pc; j += BW_THREADS) ob.inv[j] = div_pos(1.0, ob.hl(cur)[j] + lambda);
  __syncthreads();
  const double* ow = ob.w(cur);
  const double* zx = ob.zx(cur);
  const double* zy = ob.zy(cur);
  const double* zz = ob.zz(cur);
  double* const wm = a.wmom + (size_t)wk * njobs * 16;
  const int s = tid & 3;
  const bool hi2 = (s & 2) != 0, hi1 = (s & 1) != 0;
  for (int r0 = 0; r0 < njobs; r0 += BW_THREADS / 4) {
    const int rk = r0 + (tid >> 2);   // rank in the sorted job order; the gradient jobs follow the pair jobs
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0;
    int job = -1;
    if (rk < npairs) {
      job = sh.jorder[rk];
      const int k1 = sh.tstart[job + 1];
      for (int k = sh.tstart[job] + s; k < k1; k += 4) {
        const unsigned int t = ob.t12[k];
        const int o1 = (int)(t & 0xffffu), o2 = (int)(t >> 16);
        const double c = ow[o1] * ow[o2] * ob.inv[ob.tj[k]];
        const double x1 = zx[o1], y1 = zy[o1], z1 = zz[o1], x2 = zx[o2], y2 = zy[o2], z2 = zz[o2];
        const double cx = c * x1, cy = c * y1, cz = c * z1;
        acc[0] += c;
        acc[1] += cx; acc[2] += cy; acc[3] += cz;
        acc[4] += c * x2; acc[5] += c * y2; acc[6] += c * z2;
        acc[7] += cx * x2; acc[8] += cx * y2; acc[9] += cx * z2;
        acc[10] += cy * x2; acc[11] += cy * y2; acc[12] += cy * z2;
        acc[13] += cz * x2; acc[14] += cz * y2; acc[15] += cz * z2;
      }
    } else if (rk < njobs) {
      // gradient job: sum_o Hpl(o) v,  v = bl / (hl + lambda);  Hpl v = w [-u ; u x 2 zc],  u = R^T v
      job = rk;
      const int p = rk - npairs;
      const double* R = sh.X[cur][p].R;
      for (int t = sh.pbase[p] + s; t < sh.pbase[p + 1]; t += 4) {
        const int o = ob.plist[t], j = ob.opt[o];
        const double s1 = ow[o] * ob.inv[j];
        const double v0 = s1 * ob.bx(cur)[j], v1 = s1 * ob.by(cur)[j], v2 = s1 * ob.bz(cur)[j];
        const double u0 = R[0] * v0 + R[3] * v1 + R[6] * v2, u1 = R[1] * v0 + R[4] * v1 + R[7] * v2, u2 = R[2] * v0 + R[5] * v1 + R[8] * v2;
        const double ax = 2 * zx[o], ay = 2 * zy[o], az = 2 * zz[o];
        acc[0] -= u0; acc[1] -= u1; acc[2] -= u2;
        acc[3] += u1 * az - u2 * ay; acc[4] += u2 * ax - u0 * az; acc[5] += u0 * ay - u1 * ax;
      }
    }
    // 16 sums over the 4 lanes with 12 shuffles: lane pairs trade halves of their value sets, twice.  Afterwards lane s
    // holds the complete sums 8 (s >> 1) + 4 (s & 1) + k, k < 4.
    double v[8], u[4];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double send = hi2 ? acc[k] : acc[k + 8], keep = hi2 ? acc[k + 8] : acc[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double send = hi1 ? v[k] : v[k + 4], keep = hi1 ? v[k + 4] : v[k];
      u[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    if (job >= 0) {
      double* out = wm + (size_t)job * 16 + 8 * (s >> 1) + 4 * (s & 1);
      out[0] = u[0]; out[1] = u[1]; out[2] = u[2]; out[3] = u[3];
    }
  }
}

__device__ __forceinline__ int bw_eps3(int x, int y) { return ((y - x + 3) % 3 == 1) ? 1 : -1; }  // eps_{x y (3-x-y)}, x != y

// ---------------------------------------------------------------------------------------------------------
// every CTA (reducer r): sum the worker partials of the pair jobs j == r (mod C) and of the poses p == r (mod C), assemble
// the blocks of the reduced system and store them into the global copy of the system (L2).  (Storing straight into CTA 0's
// shared memory through DSMEM was the first version: 16 CTAs x 4 KB into ONE receiver at ~20 B/clk cost 3.5 k cycles.)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bw_reduce_assemble(const BaArgs& a, BwShared& sh, int cur, double lambda, int rank, int nranks, int nwk,
                                                   double* S0 /* the system in global memory */, int ld, int np) {
  const int tid = threadIdx.x, W = a.W;
  const int npairs = W * (W + 1) / 2, njobs = npairs + W;
  const int nslots = (npairs > rank) ? (npairs - rank + nranks - 1) / nranks : 0;
  const int npslots = (W > rank) ? (W - rank + nranks - 1) / nranks : 0;
  const size_t wstride = (size_t)njobs * 16;
  for (int idx = tid; idx < nslots * 16; idx += BW_THREADS) {
    const int slot = idx >> 4, v = idx & 15, job = rank + slot * nranks;
    const double* src = a.wmom + (size_t)job * 16 + v;
    double t[BA_MAX_CLUSTER - 1];
#pragma unroll
    for (int c = 0; c < BA_MAX_CLUSTER - 1; c++) t[c] = (c < nwk) ? src[c * wstride] : 0.0;   // all loads in flight
    double s = 0;
#pragma unroll
    for (int c = 0; c < BA_MAX_CLUSTER - 1; c++) s += t[c];
    sh.rmom[slot][v] = s;
  }
  for (int idx = BW_THREADS - 1 - tid; idx < npslots * 33; idx += BW_THREADS) {   // from the top: overlaps the loop above
    const int ps = idx / 33, v = idx - 33 * ps, p = rank + ps * nranks;
    double s = 0;
    if (v < 27) {
      const double* src = a.wpsum + ((size_t)cur * nwk * W + p) * 28 + v;
      double t[BA_MAX_CLUSTER - 1];
#pragma unroll
      for (int c = 0; c < BA_MAX_CLUSTER - 1; c++) t[c] = (c < nwk) ? src[(size_t)c * W * 28] : 0.0;
#pragma unroll
      for (int c = 0; c < BA_MAX_CLUSTER - 1; c++) s += t[c];
      sh.rps[ps][v] = s;
    } else {
      const double* src = a.wmom + (size_t)(npairs + p) * 16 + (v - 27);
      double t[BA_MAX_CLUSTER - 1];
#pragma unroll
      for (int c = 0; c < BA_MAX_CLUSTER - 1; c++) t[c] = (c < nwk) ? src[c * wstride] : 0.0;
#pragma unroll
      for (int c = 0; c < BA_MAX_CLUSTER - 1; c++) s += t[c];
      sh.rgr[ps][v - 27] = s;
    }
  }
  for (int idx = tid; idx < nslots * 9; idx += BW_THREADS) {
    const int slot = idx / 9, e = idx - 9 * slot, job = rank + slot * nranks;
    const double* R1 = sh.X[cur][sh.jp1[job]].R;
    const double* R2 = sh.X[cur][sh.jp2[job]].R;
    const int i = e / 3, j = e - 3 * i;
    sh.rR12[slot][e] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
  }
  __syncthreads();
  const double* eH = a.eH + (size_t)cur * W * 120;
  for (int idx = tid; idx < nslots * 36; idx += BW_THREADS) {
    const int slot = idx / 36, e = idx - 36 * slot, job = rank + slot * nranks;
    const int p1 = sh.jp1[job], p2 = sh.jp2[job];
    const double* m = sh.rmom[slot];     // m[0] = C0, m[1..3] = A1, m[4..6] = A2, m[7..15] = Mz (row = zc1 component)
    const double* R12 = sh.rR12[slot];
    const int r = e / 6, c = e - 6 * r;
    double t = 0;
    if (r < 3 && c < 3) t = m[0] * R12[3 * r + c];
    else if (r < 3) {
      const int j = c - 3;
      for (int b = 0; b < 3; b++)
        if (b != j) { const int k = 3 - b - j; t -= R12[3 * r + b] * (double)bw_eps3(b, k) * 2.0 * m[4 + k]; }
    } else if (c < 3) {
      const int i = r - 3;
      for (int b = 0; b < 3; b++)
        if (b != i) { const int k = 3 - i - b; t += (double)bw_eps3(i, k) * 2.0 * m[1 + k] * R12[3 * b + c]; }
    } else {
      const int i = r - 3, j = c - 3;
      for (int k = 0; k < 3; k++) {
        if (k == i) continue;
        const int aa = 3 - i - k;
        for (int mm = 0; mm < 3; mm++) {
          if (mm == j) continue;
          const int bb = 3 - mm - j;
          t -= 4.0 * (double)(bw_eps3(i, k) * bw_eps3(bb, mm)) * m[7 + 3 * k + mm] * R12[3 * aa + bb];
        }
      }
    }
    double h = 0;
    if (p2 == p1) {
      const int lo = r < c ? r : c, hi = r < c ? c : r;
      h = sh.rps[slot][lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo)] + ((r == c) ? lambda : 0.0);
      if (p1 < W - 1) h += eH[p1 * 120 + e];            // w Ji^T Ji of edge p1
      if (p1 > 0) h += eH[(p1 - 1) * 120 + 72 + e];     // w Jj^T Jj of edge p1-1
    } else if (p2 == p1 + 1) h = eH[p1 * 120 + 36 + e]; // w Ji^T Jj of edge p1
    // block (p1,p2), p1 <= p2, transposed into the lower triangle
    if (p1 != p2 || r <= c) S0[(6 * p2 + c) * ld + 6 * p1 + r] = h - t;
  }
  for (int idx = tid; idx < npslots * 6; idx += BW_THREADS) {
    const int ps = idx / 6, k = idx - 6 * ps, p = rank + ps * nranks;
    double b = sh.rps[ps][21 + k];
    if (p < W - 1) b += eH[p * 120 + 108 + k];
    if (p > 0) b += eH[(p - 1) * 120 + 114 + k];
    a.bp[6 * p + k] = b;
    S0[np * ld + 6 * p + k] = b - sh.rgr[ps][k];   // right-hand side = row np
  }
  // identity padding up to a multiple of the factorisation block (written once per trial: the factorisation works in place)
  if (rank == nranks - 1) {
    const int n = 6 * W;
    for (int idx = tid; idx < (np - n) * (np + 1); idx += BW_THREADS) {
      const int r = n + idx / (np + 1), c = idx - (r - n) * (np + 1);
      if (c < np) { if (c <= r) S0[r * ld + c] = (r == c) ? 1.0 : 0.0; }
      else S0[np * ld + r] = 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA 0: odometry edges at state st -> products for the assembly, robust chi2 (returned as per-thread partial)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double bw_edges(const BaArgs& a, BwShared& sh, int st, double* scratch /* [W][80] dynamic smem */) {
  const int tid = threadIdx.x, W = a.W;
  double chi = 0;
  if (tid < W - 1) {
    double* J = scratch + 80 * tid;
    double e[6], r0, w;
    edge_se3(sh.X[st][tid], sh.X[st][tid + 1], sh.Zinv[tid], e, J, J + 36);
    double c = 0;
    for (int k = 0; k < 6; k++) c += e[k] * e[k];
    huber(c * a.info_cam, a.d_cam, r0, w);
    chi = r0;
    for (int k = 0; k < 6; k++) J[72 + k] = e[k];
    J[78] = w * a.info_cam;
  }
  __syncthreads();
  double* out = a.eH + (size_t)st * W * 120;
  for (int idx = tid; idx < (W - 1) * 120; idx += BW_THREADS) {
    const int i = idx / 120, v = idx - 120 * i;
    const double* Ji = scratch + 80 * i;
    const double* Jj = Ji + 36;
    const double* E = Ji + 72;
    const double w = Ji[78];
    double t = 0;
    if (v < 108) {
      const int blk = v / 36, e = v - 36 * blk, r = e / 6, c = e - 6 * r;
      const double* A = (blk == 2) ? Jj : Ji;
      const double* B = (blk == 0) ? Ji : Jj;
      for (int k = 0; k < 6; k++) t += A[6 * k + r] * B[6 * k + c];
      t *= w;
    } else {
      const int r = (v - 108) % 6;
      const double* A = (v < 114) ? Ji : Jj;
      for (int k = 0; k < 6; k++) t += A[6 * k + r] * E[k];
      t *= -w;
    }
    out[idx] = t;
  }
  __syncthreads();
  return chi;
}

// ---------------------------------------------------------------------------------------------------------
// the kernel: grid = one cluster of C CTAs (16 when the device allows it, else 8), BW_THREADS threads each
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BW_THREADS, 1) ba_window_kernel(BaArgs a) {
  extern __shared__ __align__(16) char dsm_raw[];
  __shared__ BwShared sh;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
  const int nwk = nranks - 1, wk = rank - 1;   // workers are ranks 1..nranks-1
  const int tid = threadIdx.x, W = a.W, P = a.P;
  const int n = 6 * W, np = bc_np(n), ld = np + 1;
  CholSm cs;
  bc_carve((double*)dsm_raw, n, cs);
  BwObs ob;
  bw_carve(dsm_raw, a.capO, a.capPt, a.capT, ob);
  __shared__ unsigned long long tph[8], tq[8], t_mark, t_start, t_mark2;   // phase timers (thread 0 of CTA 0 / of worker 0)
  const bool timer = (rank == 0 && tid == 0);
  if (tid < 8) { tph[tid] = 0; tq[tid] = 0; }
  __syncthreads();
  if (timer) { t_start = gtime(); t_mark = t_start; }
#define BW_TOC(slot) do { if (timer) { const unsigned long long t_ = gtime(); tph[slot] += t_ - t_mark; t_mark = t_; } } while (0)
  // second timer on worker 0 (rank 1): its own time inside the phases, slots 8..15 of t_phase
  const bool timer2 = (rank == 1 && tid == 0);
#define BW_MARK2() do { if (timer2) t_mark2 = gtime(); } while (0)
#define BW_TOC2(slot) do { if (timer2) { const unsigned long long t_ = gtime(); tq[slot] += t_ - t_mark2; t_mark2 = t_; } } while (0)

  // programmatic dependent launch: let the solve queued behind this one start its prologue (tables, term lists) right away;
  // it blocks at griddepcontrol.wait below until this grid has completed and flushed its results
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // ---- tables
  const int npairs = W * (W + 1) / 2;
  if (tid == 0) lm_reset(&sh.ctl);
  for (int job = tid; job < npairs; job += BW_THREADS) {
    int d = 0, rem = job;
    while (rem >= W - d) { rem -= W - d; d++; }
    sh.jp1[job] = (unsigned char)rem; sh.jp2[job] = (unsigned char)(rem + d);
  }
  if (rank == 0) {
    if (tid == 0) { sh.Mc = 0; sh.Pc = 0; }
  } else {
    // local point c + j * nwk <-> sorted point index; counts of the round-robin deal in closed form
    auto cntmod = [&](int x) { return (x + nwk - 1 - wk) / nwk; };   // #{ n in [0, x) : n % nwk == wk }
    for (int f = tid; f <= W; f += BW_THREADS) sh.lgrp[f] = cntmod(a.grp_start[f]);
    for (int e = tid; e < W * (W + 1); e += BW_THREADS) {
      const int f = e / (W + 1), g = a.grp_start[f];
      sh.lcnt[e] = cntmod(g + a.cnt_gt[e]) - cntmod(g);
    }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of lcnt -> lbase (one warp, consecutive runs per lane)
      const int E = W * (W + 1), per = (E + 31) / 32, e0 = tid * per, e1 = min(e0 + per, E);
      int s = 0;
      for (int e = e0; e < e1; e++) s += sh.lcnt[e];
      int incl = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      int run = incl - s;
      for (int e = e0; e < e1; e++) { sh.lbase[e] = run; run += sh.lcnt[e]; }
      if (tid == 31) sh.Mc = incl;
    } else if (tid == 32) {
      sh.Pc = cntmod(P);
    } else if (tid == 64) {
      int run = 0;
      for (int p = 0; p < W; p++) {
        sh.pbase[p] = run;
        for (int f = 0; f <= p; f++) run += sh.lcnt[f * (W + 1) + (p - f)];
      }
      sh.pbase[W] = run;
    }
    __syncthreads();
    const int Pc = sh.Pc;
    for (int j = tid; j < Pc; j += BW_THREADS) {
      const int nidx = wk + j * nwk, f = a.pt_first[nidx], len = a.pt_len[nidx];
      ob.pf[j] = (unsigned char)f; ob.plen[j] = (unsigned char)len;
      const int i = j - sh.lgrp[f], ig = nidx - a.grp_start[f];
      for (int k = 0; k < len; k++) {
        const int o = sh.lbase[f * (W + 1) + k] + i, p = f + k;
        const size_t q = (size_t)a.pose_base[p] + a.off[p * (W + 1) + f] + ig;
        ob.opt[o] = (unsigned short)j; ob.opo[o] = (unsigned char)p;
        ob.mx[o] = a.obs_xyz[q]; ob.my[o] = a.obs_xyz[(size_t)a.M + q]; ob.mz[o] = a.obs_xyz[2 * (size_t)a.M + q];
      }
    }
    for (int p = tid; p < W; p += BW_THREADS) {
      int pos = sh.pbase[p];
      for (int f = 0; f <= p; f++) {
        const int e = f * (W + 1) + (p - f), b = sh.lbase[e], c = sh.lcnt[e];
        for (int i = 0; i < c; i++) ob.plist[pos++] = (unsigned short)(b + i);
      }
    }
    // Schur term lists: the points common to poses p1 <= p2 are, for every group f <= p1, the first lcnt[f][p2-f] local points
    for (int job = tid; job < npairs; job += BW_THREADS) {
      const int p1 = sh.jp1[job], p2 = sh.jp2[job];
      int c = 0;
      for (int f = 0; f <= p1; f++) c += sh.lcnt[f * (W + 1) + (p2 - f)];
      sh.tstart[job + 1] = c;
    }
    __syncthreads();
    if (tid < 32) {   // inclusive scan of the counts -> tstart (one warp, consecutive runs per lane)
      const int per = (npairs + 31) / 32, e0 = tid * per, e1 = min(e0 + per, npairs);
      int sum = 0;
      for (int e = e0; e < e1; e++) sum += sh.tstart[e + 1];
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      int run = incl - sum;
      for (int e = e0; e < e1; e++) { const int c = sh.tstart[e + 1]; run += c; sh.tstart[e + 1] = run; }
      if (tid == 0) sh.tstart[0] = 0;
    }
    __syncthreads();
    for (int job = tid; job < npairs; job += BW_THREADS) {
      const int p1 = sh.jp1[job], p2 = sh.jp2[job];
      int k = sh.tstart[job];
      const int cnt = sh.tstart[job + 1] - k;
      for (int f = 0; f <= p1; f++) {
        const int n = sh.lcnt[f * (W + 1) + (p2 - f)];
        const int b1 = sh.lbase[f * (W + 1) + (p1 - f)], b2 = sh.lbase[f * (W + 1) + (p2 - f)], g0 = sh.lgrp[f];
        for (int i = 0; i < n; i++, k++) { ob.t12[k] = (unsigned int)(b1 + i) | ((unsigned int)(b2 + i) << 16); ob.tj[k] = (unsigned short)(g0 + i); }
      }
      // rank of this job by descending term count (ties by job index)
      int rk = 0;
      for (int j2 = 0; j2 < npairs; j2++) {
        const int c2 = sh.tstart[j2 + 1] - sh.tstart[j2];
        rk += (c2 > cnt || (c2 == cnt && j2 < job)) ? 1 : 0;
      }
      sh.jorder[rk] = (unsigned short)job;
    }
  }
  // ---- state values.  Everything above depends on the STRUCTURE of the window only; the values may be outputs of the solve in
  //      front of this one (same float32 bits as a round trip through the host Map): wait for it here.
  unsigned long long t_w0 = 0;
  if (timer) t_w0 = gtime();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (timer) { const unsigned long long t_ = gtime(); tq[4] = t_start; tq[5] = t_w0; t_start += t_ - t_w0; t_mark = t_; tq[7] = t_ - t_w0; tq[6] = t_; }   // the span excludes the time blocked behind the solve in front
  for (int p = tid; p < W; p += BW_THREADS) {
    const int src = a.chain ? a.chain[p] : -1;
    Pose X;
    pose_from_f32(src >= 0 ? a.prev_poses + 16 * src : a.poses_f32 + 16 * p, X);
    sh.X[0][p] = X;
    sh.X[1][p] = X;
  }
  if (rank == 0) {
    for (int i = tid; i < W - 1; i += BW_THREADS) {
      const int s0 = a.chain ? a.chain[i] : -1, s1 = a.chain ? a.chain[i + 1] : -1;
      Pose Z, I, Zi;
      pose_from_f32((s0 >= 0 && s1 == s0 + 1) ? a.prev_rel + 16 * s0 : a.rel_f32 + 16 * i, Z);
      for (int k = 0; k < 9; k++) I.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
      I.t[0] = I.t[1] = I.t[2] = 0;
      pose_inv_mul(Z, I, Zi);
      sh.Zinv[i] = Zi;
    }
  } else {
    const int Pc = sh.Pc;
    for (int j = tid; j < Pc; j += BW_THREADS) {
      const int nidx = wk + j * nwk;
      const int src = a.chain ? a.chain[W + nidx] : -1;
      const float* q = src >= 0 ? a.prev_points + 3 * (size_t)src : a.points_f32 + 3 * (size_t)nidx;
      ob.px(0)[j] = (double)q[0]; ob.py(0)[j] = (double)q[1]; ob.pz(0)[j] = (double)q[2];
    }
  }
  __syncthreads();
  if (W + P == 0 || a.max_iterations <= 0) {
    if (rank == 0 && tid == 0) { sh.ctl.iterations = (W + P == 0) ? -1 : 0; *a.ctl_out = sh.ctl; }
  }
  const bool run = !(W + P == 0 || a.max_iterations <= 0);

  if (run) {
    // ---- initial state: chi2 + linearisation
    {
      double chi = 0, hmax = 0;
      if (rank == 0) chi = bw_block_reduce<false>(bw_edges(a, sh, 0, (double*)dsm_raw), sh.red);
      else bw_obs_pass(a, sh, ob, 0, wk, nwk, chi, hmax);
      if (tid == 0) { a.part[rank * 4 + 0] = chi; a.part[rank * 4 + 2] = hmax; }
    }
    cluster.sync();
    BW_TOC(5);
    {  // initial chi2; max |H_jj| for the initial damping: point blocks (worker maxima) and pose-block diagonals
      double m = 0;
      for (int i = tid; i < n; i += BW_THREADS) {
        const int p = i / 6, k = i - 6 * p, idx = k * 6 - (k * (k - 1)) / 2;
        double h = 0;
        for (int c = 0; c < nwk; c++) h += a.wpsum[((size_t)c * W + p) * 28 + idx];
        if (p < W - 1) h += a.eH[p * 120 + 7 * k];
        if (p > 0) h += a.eH[(p - 1) * 120 + 72 + 7 * k];
        m = fmax(m, fabs(h));
      }
      m = bw_block_reduce<true>(m, sh.red);
      if (tid == 0) {
        double c = 0;
        for (int r = 0; r < nranks; r++) { c += a.part[r * 4]; m = fmax(m, a.part[r * 4 + 2]); }
        sh.ctl.currentChi = c;
        sh.red[63] = m;
      }
      __syncthreads();
    }
    const double maxdiag = sh.red[63];
    for (int it = 0; it < a.max_iterations; it++) {
      if (sh.ctl.stop_flag || !sh.ctl.ok) break;
      const int cur = sh.ctl.cur;  // state buffer == linearisation buffer
      if (tid == 0) lm_begin_iteration(&sh.ctl, it, maxdiag, -1.0);
      __syncthreads();
      while (true) {
        const double lambda = sh.ctl.lambda;
        BW_MARK2();
        if (rank > 0) bw_schur(a, sh, ob, cur, lambda, wk);
        BW_TOC2(0);
        cluster.sync();
        BW_TOC(0);
        BW_MARK2();
        bw_reduce_assemble(a, sh, cur, lambda, rank, nranks, nwk, a.Sg, ld, np);
        BW_TOC2(1);
        cluster.sync();
        BW_TOC(1);
        if (rank == 0) {
          bc_stage(cs, a.Sg);
          BW_TOC(6);
          bc_factor(cs, &sh.s_bad, nullptr, a.Sg);
          const int failed = sh.s_bad;
          if (!failed) bc_backsolve(cs);
          __syncthreads();
          BW_TOC(2);
          // increments (x = b when the solver failed, like LinearSolverCSparse), trial poses, pose part of the scale
          const double* ys = cs.S + (size_t)np * ld;
          double sc = 0;
          for (int i = tid; i < n; i += BW_THREADS) {
            const double b = a.bp[i], x = failed ? b : ys[i];
            a.xp[i] = x;
            sh.xp[i] = x;
            sc += x * (lambda * x + b);
          }
          __syncthreads();
          for (int p = tid; p < W; p += BW_THREADS) {
            Pose o;
            pose_oplus(sh.X[cur][p], sh.xp + 6 * p, o);
            sh.X[cur ^ 1][p] = o;
            a.X[(size_t)(cur ^ 1) * W + p] = o;
          }
          sc = bw_block_reduce<false>(sc, sh.red);
          if (tid == 0) { a.part[1] = sc; a.part[3] = failed ? 1.0 : 0.0; }
        }
        cluster.sync();
        BW_TOC(2);
        const int failed = a.part[3] != 0.0;
        if (rank == 0) {
          const double chi = bw_block_reduce<false>(bw_edges(a, sh, cur ^ 1, (double*)dsm_raw), sh.red);
          if (tid == 0) a.part[0] = chi;
        } else {
          BW_MARK2();
          for (int i = tid; i < n; i += BW_THREADS) sh.xp[i] = a.xp[i];
          for (int p = tid; p < W; p += BW_THREADS) sh.X[cur ^ 1][p] = a.X[(size_t)(cur ^ 1) * W + p];
          __syncthreads();
          BW_TOC2(2);
          double scale = bw_update(a, sh, ob, cur, lambda, failed);
          scale = bw_block_reduce<false>(scale, sh.red);
          BW_TOC2(3);
          double chi, hmax;
          bw_obs_pass(a, sh, ob, cur ^ 1, wk, nwk, chi, hmax);
          if (tid == 0) { a.part[rank * 4 + 0] = chi; a.part[rank * 4 + 1] = scale; }
          BW_TOC2(4);
        }
        cluster.sync();
        BW_TOC(3);
        if (tid == 0) {
          double chi = 0, scale = 0;
          for (int r = 0; r < nranks; r++) { chi += a.part[r * 4]; scale += a.part[r * 4 + 1]; }
          lm_trial(&sh.ctl, chi, scale, failed);
        }
        __syncthreads();
        // part[] is next written two cluster barriers later: no race with slower CTAs
        if (!lm_more_trials(&sh.ctl)) break;
      }
      if (tid == 0) lm_end_iteration(&sh.ctl, it, a.gain_threshold, rank == 0 ? a.rec : nullptr);
      __syncthreads();
      BW_TOC(4);
    }
  }
  // ---- results: poses, relative motions (Converter::toInvMatrix(pose[i-1]) * pose[i] on the float32 poses, cv::Mat CV_32F
  //      semantics: double accumulation, one rounding; src/Optimizer.cc:1072-1075), points
  const int fin = run ? sh.ctl.cur : 0;
  if (rank == 0) {
    for (int p = tid; p < W; p += BW_THREADS) {
      pose_to_f32(sh.X[fin][p], sh.pf32[p]);
      for (int k = 0; k < 16; k++) a.out_poses[16 * p + k] = sh.pf32[p][k];
    }
    __syncthreads();
    for (int i = 1 + tid; i < W; i += BW_THREADS) {
      const float* A = sh.pf32[i - 1];
      const float* B = sh.pf32[i];
      float Ai[16];
      for (int k = 0; k < 16; k++) Ai[k] = 0.f;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) Ai[4 * r + c] = A[4 * c + r];
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (double)(-Ai[4 * r + k]) * (double)A[4 * k + 3];
        Ai[4 * r + 3] = (float)s;
      }
      Ai[15] = 1.f;
      float* out = a.out_rel + 16 * (i - 1);
      for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
          double s = 0;
          for (int k = 0; k < 4; k++) s += (double)Ai[4 * r + k] * (double)B[4 * k + c];
          out[4 * r + c] = (float)s;
        }
    }
    if (tid == 0) {
      if (run) *a.ctl_out = sh.ctl;
      const unsigned long long t_end = gtime();
      tph[7] = t_end - t_start;
      for (int k = 0; k < 8; k++) a.t_phase[k] = tph[k];
      a.t_phase[16] = tq[6]; a.t_phase[17] = t_end; a.t_phase[18] = tq[7]; a.t_phase[21] = tq[4]; a.t_phase[22] = tq[5];   // absolute: released behind the solve in front, end; time blocked
    }
  } else {
    if (timer2) for (int k = 0; k < 8; k++) a.t_phase[8 + k] = tq[k];
    const int Pc = sh.Pc;
    for (int j = tid; j < Pc; j += BW_THREADS) {
      const size_t nidx = (size_t)wk + (size_t)j * nwk;
      a.out_points[3 * nidx] = (float)ob.px(fin)[j];
      a.out_points[3 * nidx + 1] = (float)ob.py(fin)[j];
      a.out_points[3 * nidx + 2] = (float)ob.pz(fin)[j];
    }
  }
  cluster.sync();   // every CTA's results are in the device block (and no CTA exits while others may address its shared memory)
  const unsigned long long t_sync = timer ? gtime() : 0;
  if (rank == 0 && a.h_out) {
    // mirror the output block into pinned host memory (posted writes over PCIe, ~30 KB) and publish the completion word: the
    // host polls it instead of waiting on a stream event
    const int n16 = a.out_bytes >> 4;
    const uint4* srcv = (const uint4*)a.out_base;
    uint4* dstv = (uint4*)a.h_out;
    for (int i = tid; i < n16; i += BW_THREADS) dstv[i] = srcv[i];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      // debug stamps straight into the mirror (t_phase[19], [20]): after the closing cluster barrier, after the mirror copy
      unsigned long long* ht = (unsigned long long*)((char*)a.h_out + ((const char*)a.t_phase - (const char*)a.out_base));
      ht[19] = t_sync; ht[20] = gtime();
      __threadfence_system();
      *a.h_flag = a.seq; __threadfence_system();
    }
  }
#undef BW_TOC
}

// =========================================================================================================
// host side
// =========================================================================================================
// shared memory a launch needs (dynamic part), or 0 when the problem does not fit the shared-memory resident kernel
size_t ba_window_smem(int W, int capO, int capPt, int capT) {
  const size_t solver = sizeof(double) * bc_smem_doubles(6 * W);
  const size_t edges = sizeof(double) * 80 * (size_t)W;
  const size_t worker = bw_worker_bytes(capO, capPt, capT);
  size_t need = solver > worker ? solver : worker;
  if (edges > need) need = edges;
  return need;
}

size_t ba_window_smem_limit() {
  cudaFuncAttributes at;
  if (cudaFuncGetAttributes(&at, ba_window_kernel) != cudaSuccess) { cudaGetLastError(); return 0; }
  return (size_t)227 * 1024 - at.sharedSizeBytes;
}

int ba_window_configure(size_t max_smem, int* cluster_out) {
  if (cudaFuncSetAttribute(ba_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem) != cudaSuccess) return -1;
  int cluster = 8;
  if (cudaFuncSetAttribute(ba_window_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(BW_THREADS); cfg.dynamicSmemBytes = max_smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, ba_window_kernel, &cfg) == cudaSuccess && nclusters >= 1) cluster = 16;
  }
  cudaGetLastError();
  *cluster_out = cluster;
  return 0;
}

// programmatic: the previous operation on the stream is a solve of this kernel that may still be running -- launch with
// programmatic stream serialisation (the kernel orders itself behind it with griddepcontrol.wait)
cudaError_t ba_window_launch(const BaArgs& a, int cluster, size_t smem, cudaStream_t s, bool programmatic) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster); cfg.blockDim = dim3(BW_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (programmatic && !getenv("VIDO_BA_NO_PDL")) {
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, ba_window_kernel, a);
}
