// ba_kernels.cu -- sliding-window graph optimisation (Optimizer::PartialBatchOptimization) on one thread-block
// cluster: the whole Levenberg-Marquardt loop (linearise, damp, Schur-reduce, dense Cholesky, back-substitute,
// update, robust chi2, accept/reject, every stop rule) runs inside ONE kernel launch; CTAs of the cluster exchange
// data through L2 and synchronise with cluster barriers, the host never intervenes.
//
// Reference code replaced (paths under /root/reference/vido_slam/):
//   graph + solve + write-back     src/Optimizer.cc:220-362, 806, 1056-1142
//   SparseOptimizer::optimize      3rdparty/g2o/g2o/core/sparse_optimizer.cpp:354-427 (chi2_check patch :393-396)
//   Levenberg::solve               3rdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
//   terminate action               3rdparty/g2o/g2o/core/sparse_optimizer_terminate_action.cpp:49-92
//   BlockSolver::buildSystem       3rdparty/g2o/g2o/core/block_solver.hpp:502-560 (+ base_binary_edge.hpp:55-120)
//   LinearSolverCSparse::solve     3rdparty/g2o/g2o/solvers/linear_solver_csparse.h:108-141 -- the reference factors the
//     full (poses+points) H; here the point blocks are eliminated first, the reduced camera system (6W x 6W) is
//     Cholesky-factored in shared memory.  Same linear system, same positive-definiteness test.
// Everything is FP64 (g2o is built in double); inputs/outputs are the float32 Map fields.
//
// Data layout (HBM/L2 resident, a few MB): points are sorted by (first observing pose f, track length descending) and
// every track observes consecutive poses (true for the reference's graph by construction).  Group f = points born
// at pose f; the points of group f seen by pose p are the PREFIX of the group with length > p - f.  Observations are
// stored POSE-MAJOR: pose p owns the contiguous range [pose_base[p], pose_base[p+1]) holding, group after group, the
// prefixes it sees; per-observation arrays (measurement, the 6x3 block Hpl) are struct-of-arrays over that index.
//   - observation of point (f, i) in pose p:  q = pose_base[p] + off[p][f] + i            (i < cnt[f][p-f])
//   - points common to poses p1 <= p2: for every f <= p1 the first cnt[f][p2-f] points of group f, found at the SAME
//     i in both poses' ranges -> every warp access is coalesced, no index lists, no searches, no atomics.
// Since J_point = R^T, the point block is (sum_o w_o) * I3: its inverse is a scalar and is applied on the fly.
#include <cooperative_groups.h>

#include <atomic>
#include <string>

#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "ba_common.h"

namespace cg = cooperative_groups;

// ba_window.cu
size_t ba_window_smem(int W, int capO, int capPt, int capT);
size_t ba_window_smem_limit();
int ba_window_configure(size_t max_smem, int* cluster_out);
cudaError_t ba_window_launch(const BaArgs& a, int cluster, size_t smem, cudaStream_t s, bool programmatic);

__device__ void phase_init(const BaArgs& a, int G, int GT) {
  for (int i = G; i < a.W; i += GT) {
    Pose X;
    pose_from_f32(a.poses_f32 + 16 * i, X);
    a.X[i] = X;
    a.X[a.W + i] = X;
  }
  for (int i = G; i < a.W - 1; i += GT) {
    Pose Z, I, Zi;
    pose_from_f32(a.rel_f32 + 16 * i, Z);
    for (int k = 0; k < 9; k++) I.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    I.t[0] = I.t[1] = I.t[2] = 0;
    pose_inv_mul(Z, I, Zi);
    a.Zinv[i] = Zi;
  }
  for (int i = G; i < 3 * a.P; i += GT) {
    const double v = (double)a.points_f32[i];
    a.pts[i] = v;
    a.pts[3 * (size_t)a.P + i] = v;
  }
}

__device__ __forceinline__ double obs_chi(const BaArgs& a, const Pose& Xp, const double* p, int o, double* zc, double* e, double& w) {
  const double z[3] = {(double)a.obs_xyz[o], (double)a.obs_xyz[a.M + o], (double)a.obs_xyz[2 * (size_t)a.M + o]};
  edge_xyz(Xp, p, z, zc, e);
  double r0;
  huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * a.info_3d, a.d_3d, r0, w);
  return r0;
}

// shared-memory copies of the small layout tables
struct BaTab {
  int grp[BA_MAX_W + 1], cnt[BA_MAX_W * (BA_MAX_W + 1)], off[BA_MAX_W * (BA_MAX_W + 1)], base[BA_MAX_W + 1];
  int ustart[BA_MAX_JOBS + 1];  // Schur work units (32 point-pairs each) + BA_JOB_PAD per non-empty job: prefix over the jobs
  unsigned char jp1[BA_MAX_JOBS], jp2[BA_MAX_JOBS];  // pose pair of every pair job
  int nch;                      // chunks per pose in the pose-block reduction
  int uq;                       // Schur units per warp
};

// The observation pass: ONE coalesced sweep over the observations (pose-major) at state `st`, written into the
// linearisation buffer of the same index.  It serves two purposes at once: the robust chi2 of the state (what the LM
// driver needs to accept or reject a trial) and the linearisation at that state (what the next iteration needs if the
// trial is accepted -- the reference recomputes both, g2o/core/sparse_optimizer.cpp:377-380 + block_solver.hpp:502).
//   warp per (pose, chunk): robust weight w, camera-frame point zc and gradient term g = w R e of every observation,
//                           the 27 sums of the pose block, the chi2;
//   thread per odometry edge: error, Jacobians, chi2.
// Output: ow/ozc/og[st], ppart[st], seJ/seE[st], part[rank][0] = chi2 partial.
__device__ void phase_obs(const BaArgs& a, int st, int G, int GT, int rank, int nranks, double* red, const BaTab& tb) {
  const Pose* X = a.X + (size_t)st * a.W;
  const double* pts = a.pts + (size_t)st * 3 * a.P;
  const int W = a.W;
  const size_t M = a.M;
  double* ow = a.ow + (size_t)st * M;
  double* ozc = a.ozc + (size_t)st * 3 * M;
  double* og = a.og + (size_t)st * 3 * M;
  const int lane = threadIdx.x & 31;
  const int gw = rank * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = nranks * (blockDim.x >> 5);
  const int nch = tb.nch;
  double chi = 0;
  for (int job = gw; job < W * nch; job += nw) {
    const int p = job / nch, ch = job - p * nch;
    const Pose Xp = X[p];
    const double* R = Xp.R;
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    const int q0 = tb.base[p], q1 = tb.base[p + 1];
    for (int o = q0 + ch * 32 + lane; o < q1; o += 32 * nch) {
      double zc[3], e[3], w;
      chi += obs_chi(a, Xp, pts + 3 * (size_t)a.obs_point[o], o, zc, e, w);
      w *= a.info_3d;
      // Hpl = w * J_pose^T J_point with J_pose = [-I | Q(zc)], J_point = R^T is a function of (w, zc) and the pose's
      // rotation only: keep those four numbers, every consumer rebuilds what it needs.  g = w R e: bl = -sum g.
      ow[o] = w;
      ozc[o] = zc[0]; ozc[M + o] = zc[1]; ozc[2 * M + o] = zc[2];
#pragma unroll
      for (int r = 0; r < 3; r++) og[r * M + o] = w * (R[3 * r] * e[0] + R[3 * r + 1] * e[1] + R[3 * r + 2] * e[2]);
      const double qx = 2 * zc[0], qy = 2 * zc[1], qz = 2 * zc[2];
      const double J[3][6] = {{-1, 0, 0, 0, -qz, qy}, {0, -1, 0, qz, 0, -qx}, {0, 0, -1, -qy, qx, 0}};
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        acc[21 + r] -= w * (J[0][r] * e[0] + J[1][r] * e[1] + J[2][r] * e[2]);
#pragma unroll
        for (int c = r; c < 6; c++) acc[idx++] += w * (J[0][r] * J[0][c] + J[1][r] * J[1][c] + J[2][r] * J[2][c]);
      }
    }
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
      double* o = a.ppart + (((size_t)st * W + p) * BA_PCHUNK + ch) * 28;
      for (int k = 0; k < 27; k++) o[k] = acc[k];
    }
  }
  // odometry edges: the lanes of the LAST warp of the cluster (normally without a pose job, so the long Jacobian
  // evaluation overlaps the observation sweep instead of following it)
  for (int i = (gw == nw - 1) ? lane : W; i < W - 1; i += 32) {
    double e[6], Ji[36], Jj[36], r0, w;
    edge_se3(X[i], X[i + 1], a.Zinv[i], e, Ji, Jj);
    double c = 0;
    for (int k = 0; k < 6; k++) c += e[k] * e[k];
    huber(c * a.info_cam, a.d_cam, r0, w);
    chi += r0;
    w *= a.info_cam;
    double* J = a.seJ + 72 * ((size_t)st * W + i);
    for (int k = 0; k < 36; k++) { J[k] = Ji[k]; J[36 + k] = Jj[k]; }
    double* E = a.seE + 8 * ((size_t)st * W + i);
    for (int k = 0; k < 6; k++) E[k] = e[k];
    E[6] = w;
  }
  double v[1] = {chi};
  block_reduce<1, false>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 0] = red[0];
}

// System blocks from the linearisation buffer `st` (after a barrier):
//   warp per pose:    sum the partial pose blocks in fixed order, add the odometry edges; lanes own the 36 entries of
//                     the block (two passes) and the 6 gradient entries;
//   thread per point: the point block (a scalar: J_point^T J_point = R R^T = I) and its gradient are plain sums of
//                     the stored per-observation terms; the loads of four observations are in flight together.
__device__ void phase_blocks(const BaArgs& a, int st, int G, int GT, int rank, int nranks, double* red, const BaTab& tb) {
  const int lane = threadIdx.x & 31;
  const int gw = rank * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = nranks * (blockDim.x >> 5);
  const int W = a.W, nch = tb.nch;
  const size_t M = a.M;
  for (int p = gw; p < W; p += nw) {
    const double* Ji = a.seJ + 72 * ((size_t)st * W + p);          // edge p ("from" vertex), valid when p < W-1
    const double* Jj = Ji + 36;
    const double* Jt = a.seJ + 72 * ((size_t)st * W + p - 1) + 36;  // edge p-1 ("to" vertex), valid when p > 0
    const double* Ee = a.seE + 8 * ((size_t)st * W + p);
    const double* Et = a.seE + 8 * ((size_t)st * W + p - 1);
    const double we = (p < W - 1) ? Ee[6] : 0.0, wt = (p > 0) ? Et[6] : 0.0;
    const double* pp = a.ppart + ((size_t)st * W + p) * BA_PCHUNK * 28;
    double dmax = 0;
    for (int e = lane; e < 36; e += 32) {
      const int r = e / 6, c = e - 6 * r;
      const int lo = r < c ? r : c, hi = r < c ? c : r;
      const int idx = lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
      double h = 0;
      for (int ch = 0; ch < nch; ch++) h += pp[ch * 28 + idx];
      if (p < W - 1) {
        double t = 0, to = 0;
        for (int k = 0; k < 6; k++) { t += Ji[6 * k + r] * Ji[6 * k + c]; to += Ji[6 * k + r] * Jj[6 * k + c]; }
        h += we * t;
        a.Hoff[36 * (size_t)p + e] = we * to;
      }
      if (p > 0) {
        double t = 0;
        for (int k = 0; k < 6; k++) t += Jt[6 * k + r] * Jt[6 * k + c];
        h += wt * t;
      }
      a.Hpp[36 * (size_t)p + e] = h;
      if (r == c) dmax = fmax(dmax, fabs(h));
    }
    if (lane < 6) {
      double b = 0;
      for (int ch = 0; ch < nch; ch++) b += pp[ch * 28 + 21 + lane];
      if (p < W - 1) {
        double t = 0;
        for (int k = 0; k < 6; k++) t += Ji[6 * k + lane] * Ee[k];
        b -= we * t;
      }
      if (p > 0) {
        double t = 0;
        for (int k = 0; k < 6; k++) t += Jt[6 * k + lane] * Et[k];
        b -= wt * t;
      }
      a.bp[6 * p + lane] = b;
    }
    dmax = warp_max(dmax);
    if (lane == 0) a.pmax[p] = dmax;
  }
  const double* ow = a.ow + (size_t)st * M;
  const double* og = a.og + (size_t)st * 3 * M;
  double mx = 0;
  for (int l = G; l < a.P; l += GT) {
    const int f = a.pt_first[l], len = a.pt_len[l], i = l - tb.grp[f];
    double h = 0, b0 = 0, b1 = 0, b2 = 0;
    for (int k0 = 0; k0 < len; k0 += 4) {
      double w[4], g0[4], g1[4], g2[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const bool on = k0 + j < len;
        const int pp = on ? f + k0 + j : f;
        const size_t o = (size_t)tb.base[pp] + tb.off[pp * (W + 1) + f] + i;
        w[j] = on ? ow[o] : 0.0;
        g0[j] = on ? og[o] : 0.0; g1[j] = on ? og[M + o] : 0.0; g2[j] = on ? og[2 * M + o] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) { h += w[j]; b0 -= g0[j]; b1 -= g1[j]; b2 -= g2[j]; }
    }
    a.hl[l] = h;
    a.bl[3 * (size_t)l] = b0; a.bl[3 * (size_t)l + 1] = b1; a.bl[3 * (size_t)l + 2] = b2;
    for (int k = 0; k < len; k++) {
      const int pp = f + k;
      const size_t o = (size_t)tb.base[pp] + tb.off[pp * (W + 1) + f] + i;
      a.ohb[o] = h; a.ohb[M + o] = b0; a.ohb[2 * M + o] = b1; a.ohb[3 * M + o] = b2;
    }
    mx = fmax(mx, h);
  }
  double v[1] = {mx};
  block_reduce<1, true>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 2] = red[0];
}

// reduced camera system: S(p1,p2) = Hpp(p1,p2) + lambda I - sum_l Hpl(p1,l) Hpl(p2,l)^T / (hl + lambda), and the
// reduced gradient b(p) = bp(p) - sum_o Hpl(o) bl / (hl + lambda).
//
// Structure used: Hpl(p,l) = w [-I | Q]^T R_p^T with Q = [2 zc]x, so for a pose pair
//   Hpl(p1,l) Hpl(p2,l)^T = w1 w2 [ R12, -R12 Q2 ; -Q1^T R12, Q1^T R12 Q2 ],  R12 = R_p1^T R_p2 (constant over the pair).
// Every entry is linear in (1, zc1, zc2, zc1 zc2^T): a pair needs only the 16 moment sums
//   C0 = sum c, A1 = sum c zc1, A2 = sum c zc2, Mz = sum c zc1 zc2^T,  c = w1 w2 / (hl + lambda),
// i.e. 9 loads and ~20 FMAs per common point instead of 37 loads and 108 FMAs, and a 16-value reduction.  The 6x6 block
// is assembled from the moments once per pair (phase_schur_reduce).
//
// Jobs: the W(W+1)/2 pose pairs (ordered by distance) followed by the W gradient jobs.  A job's terms are cut into
// UNITS of 32 (one per lane); the flat unit list is dealt in equal contiguous runs to the warps of the whole cluster,
// so the load is balanced whatever the track-length distribution.  A warp accumulates while it stays inside a job and
// flushes a partial when the job changes; phase_schur_reduce adds the partials of every job in slot order
// (deterministic, no atomics) and writes the system straight into CTA 0's shared memory (DSMEM).
__device__ __forceinline__ void job_pair(int job, int W, int& p1, int& p2) {
  int d = 0, rem = job;
  while (rem >= W - d) { rem -= W - d; d++; }
  p1 = rem; p2 = rem + d;
}

__device__ void phase_schur_units(const BaArgs& a, double lambda, int st, int rank, int nranks, const BaTab& tb) {
  const int W = a.W;
  const size_t M = a.M;
  const int lane = threadIdx.x & 31;
  const int gw = rank * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int npairs = W * (W + 1) / 2, njobs = npairs + W;
  const int U = tb.ustart[njobs], q = tb.uq;
  const double* ow = a.ow + (size_t)st * M;
  const double* ozc = a.ozc + (size_t)st * 3 * M;
  int u = gw * q;
  const int ue = min(u + q, U);
  if (u >= ue) return;
  int job;
  {  // last job whose first unit is <= u (empty jobs share their start with the next one and are skipped below)
    int lo = 0, hi = njobs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tb.ustart[mid] <= u) lo = mid; else hi = mid;
    }
    job = lo;
  }
  while (u < ue) {
    const int js = tb.ustart[job], je = tb.ustart[job + 1];
    if (je <= u) { job++; continue; }
    const int uend = min(ue, je - BA_JOB_PAD);  // the last BA_JOB_PAD units of a job are padding (no terms)
    const int unext = min(ue, je);
    const int slot = gw - js / q;
    double* out = a.spart + ((size_t)job * BA_MAX_SLOTS + slot) * 16;
    if (job < npairs) {
      const int p1 = tb.jp1[job], p2 = tb.jp2[job];
      // the pair's common points are the first cnt[f][p2-f] points of every group f <= p1: lane f holds the group's
      // count and inclusive prefix, the owner group of a flat index is found with p1 shuffles
      const int cf = (lane <= p1) ? tb.cnt[lane * (W + 1) + (p2 - lane)] : 0;
      int incl = cf;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      double acc[16];
#pragma unroll
      for (int k = 0; k < 16; k++) acc[k] = 0;
      // four units per trip: the group search is shared, and the 36 loads of the four terms are in flight together
      // (the warp's unit run is the critical path of this phase, so latency per unit is what counts)
      for (; u < uend; u += 4) {
        int t[4], f[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; j++) t[j] = (u + j < uend) ? (u + j - js) * 32 + lane : total;
        // f = number of groups k < p1 with incl_k <= t (incl is non-decreasing): binary search with per-lane shuffles
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int cand = f[j] + step;
            const int v = __shfl_sync(0xffffffffu, incl, (cand - 1) & 31);
            if (cand <= p1 && t[j] >= v) f[j] = cand;
          }
        }
        double c[4], x1[4], y1[4], z1[4], x2[4], y2[4], z2[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int excl = __shfl_sync(0xffffffffu, incl - cf, f[j]);
          const bool on = t[j] < total;
          const int fj = on ? f[j] : 0, i = on ? t[j] - excl : 0;
          const size_t q1 = tb.base[p1] + tb.off[p1 * (W + 1) + fj] + i, q2 = tb.base[p2] + tb.off[p2 * (W + 1) + fj] + i;
          const double hl = on ? a.hl[tb.grp[fj] + i] : 1.0;
          const double w12 = on ? ow[q1] * ow[q2] : 0.0;
          x1[j] = on ? ozc[q1] : 0.0; y1[j] = on ? ozc[M + q1] : 0.0; z1[j] = on ? ozc[2 * M + q1] : 0.0;
          x2[j] = on ? ozc[q2] : 0.0; y2[j] = on ? ozc[M + q2] : 0.0; z2[j] = on ? ozc[2 * M + q2] : 0.0;
          c[j] = div_pos(w12, hl + lambda);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double cx = c[j] * x1[j], cy = c[j] * y1[j], cz = c[j] * z1[j];
          acc[0] += c[j];
          acc[1] += cx; acc[2] += cy; acc[3] += cz;
          acc[4] += c[j] * x2[j]; acc[5] += c[j] * y2[j]; acc[6] += c[j] * z2[j];
          acc[7] += cx * x2[j]; acc[8] += cx * y2[j]; acc[9] += cx * z2[j];
          acc[10] += cy * x2[j]; acc[11] += cy * y2[j]; acc[12] += cy * z2[j];
          acc[13] += cz * x2[j]; acc[14] += cz * y2[j]; acc[15] += cz * z2[j];
        }
      }
      u = unext;
      warp_sum16(acc);
      if ((lane & 1) == 0) out[lane >> 1] = acc[0];  // lane bits 4..1 = moment index
    } else {
      const int p = job - npairs;
      const double* R = a.X[(size_t)st * W + p].R;
      double acc[6] = {0, 0, 0, 0, 0, 0};
      const int qe = tb.base[p + 1];
      for (; u < uend; u += 4) {
        // Hpl v = w [-u ; u x 2 zc],  u = R^T v,  v = bl / (hl + lambda); four units per trip (see above)
        double sc[4], b0[4], b1[4], b2[4], w[4], qx[4], qy[4], qz[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const size_t o = (size_t)tb.base[p] + (size_t)(u + j - js) * 32 + lane;
          const bool on = (u + j < uend) && o < (size_t)qe;
          sc[j] = on ? a.ohb[o] : 1.0;
          b0[j] = on ? a.ohb[M + o] : 0.0; b1[j] = on ? a.ohb[2 * M + o] : 0.0; b2[j] = on ? a.ohb[3 * M + o] : 0.0;
          w[j] = on ? ow[o] : 0.0;
          qx[j] = on ? ozc[o] : 0.0; qy[j] = on ? ozc[M + o] : 0.0; qz[j] = on ? ozc[2 * M + o] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double s1 = div_pos(w[j], sc[j] + lambda);
          const double v0 = s1 * b0[j], v1 = s1 * b1[j], v2 = s1 * b2[j];
          const double u0 = R[0] * v0 + R[3] * v1 + R[6] * v2, u1 = R[1] * v0 + R[4] * v1 + R[7] * v2, u2 = R[2] * v0 + R[5] * v1 + R[8] * v2;
          const double ax = 2 * qx[j], ay = 2 * qy[j], az = 2 * qz[j];
          acc[0] -= u0; acc[1] -= u1; acc[2] -= u2;
          acc[3] += u1 * az - u2 * ay; acc[4] += u2 * ax - u0 * az; acc[5] += u0 * ay - u1 * ax;
        }
      }
      u = unext;
#pragma unroll
      for (int r = 0; r < 6; r++) acc[r] = warp_sum(acc[r]);
      if (lane == 0)
        for (int r = 0; r < 6; r++) out[r] = acc[r];
    }
    job++;
  }
}

__device__ __forceinline__ int eps3(int x, int y) { return ((y - x + 3) % 3 == 1) ? 1 : -1; }  // eps_{x y (3-x-y)}, x != y

// warp per job: lanes 0..15 add the job's partial moment sums in slot order, the block is assembled from the moments
// and written (with the pose-pose Hessian block and the damping) into the solver's buffer
__device__ void phase_schur_reduce(const BaArgs& a, double lambda, int st, int rank, int nranks, const BaTab& tb, double* Ls0,
                                   double* smr /* [warps][32] shared scratch */) {
  const int W = a.W, n = 6 * W, ld = n + 1;
  const int npairs = W * (W + 1) / 2, njobs = npairs + W, q = tb.uq;
  const int lane = threadIdx.x & 31;
  const int gw = rank * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = nranks * (blockDim.x >> 5);
  for (int job = gw; job < njobs; job += nw) {
    const int js = tb.ustart[job], je = tb.ustart[job + 1];
    const int nslots = (je > js) ? (je - 1) / q - js / q + 1 : 0;
    double mom = 0;
    if (lane < 16) {
      const double* sp = a.spart + (size_t)job * BA_MAX_SLOTS * 16 + lane;
      for (int sl = 0; sl < nslots; sl++) mom += sp[16 * sl];
    }
    if (job < npairs) {
      const int p1 = tb.jp1[job], p2 = tb.jp2[job];
      // moments (lanes 0..15) and R12 = R_p1^T R_p2 (lanes 16..24) go through a per-warp shared scratch: they are
      // indexed by run-time (r, c) below, which as register arrays would live in local memory
      double* m = smr + 32 * (threadIdx.x >> 5);
      double* R12 = m + 16;
      __syncwarp();
      if (lane < 16) m[lane] = mom;
      else if (lane < 25) {
        const double* R1 = a.X[(size_t)st * W + p1].R;
        const double* R2 = a.X[(size_t)st * W + p2].R;
        const int i = (lane - 16) / 3, j = (lane - 16) - 3 * i;
        R12[lane - 16] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
      }
      __syncwarp();
      // m[0] = C0, m[1..3] = A1, m[4..6] = A2, m[7..15] = Mz (row = zc1 component)
      for (int e = lane; e < 36; e += 32) {
        const int r = e / 6, c = e - 6 * r;
        double t = 0;
        if (r < 3 && c < 3) t = m[0] * R12[3 * r + c];
        else if (r < 3) {
          const int j = c - 3;
          for (int b = 0; b < 3; b++)
            if (b != j) { const int k = 3 - b - j; t -= R12[3 * r + b] * (double)eps3(b, k) * 2.0 * m[4 + k]; }
        } else if (c < 3) {
          const int i = r - 3;
          for (int b = 0; b < 3; b++)
            if (b != i) { const int k = 3 - i - b; t += (double)eps3(i, k) * 2.0 * m[1 + k] * R12[3 * b + c]; }
        } else {
          const int i = r - 3, j = c - 3;
          for (int k = 0; k < 3; k++) {
            if (k == i) continue;
            const int aa = 3 - i - k;
            for (int mm = 0; mm < 3; mm++) {
              if (mm == j) continue;
              const int bb = 3 - mm - j;
              t -= 4.0 * (double)(eps3(i, k) * eps3(bb, mm)) * m[7 + 3 * k + mm] * R12[3 * aa + bb];
            }
          }
        }
        double h = 0;
        if (p2 == p1) h = a.Hpp[36 * (size_t)p1 + e] + ((r == c) ? lambda : 0.0);
        else if (p2 == p1 + 1) h = a.Hoff[36 * (size_t)p1 + e];
        // block (p1,p2), p1 <= p2, transposed into the lower triangle of the solver's buffer in CTA 0's shared memory
        if (p1 != p2 || r <= c) Ls0[(6 * p2 + c) * ld + 6 * p1 + r] = h - t;
      }
    } else {
      const int p = job - npairs;
      if (lane < 6) Ls0[n * ld + 6 * p + lane] = a.bp[6 * p + lane] - mom;  // right-hand side = row n
    }
  }
}

// blocked (6x6) LL^T of the reduced system in shared memory; one CTA.  ld is odd to avoid bank conflicts on column
// accesses.  Layout: rows 0..n-1 = lower triangle of S, row n = the right-hand side: carrying b as an extra row of the
// matrix turns the forward substitution into part of the panel/trailing steps (row n of L is y = L^-1 b).
// The factorisation is latency-bound (one rsqrt + a short FMA chain per pivot), so the critical path is kept short:
// the 6x6 diagonal block is factored in registers by one lane (reciprocal diagonal kept in Li), panels and the back
// substitution are 6-step triangular solves against it, and while warps 1.. apply column jb's trailing update, warp 0
// already updates and factors the next diagonal block (look-ahead).
// Also applies the pose increments (trial poses) and their part of the scale.
// 1/sqrt(d) for a positive normal d: the library fast path (MUFU.RSQ64H + one third-order correction) without its
// special-case subroutine -- a CALL inside a latency-critical chain makes the compiler park live values in local memory
__device__ __forceinline__ double rsqrt_pos(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(d, -(y * y), 1.0);
  return fma(fma(e, 0.375, 0.5), y * e, y);
}

// 6x6 Cholesky of a diagonal block by one lane, straight-line scalar code (no local arrays, no subroutine calls: either
// would put local-memory round trips into this latency-critical chain).  Writes L in place and 1/L_jj to dinv.
__device__ __forceinline__ bool chol_diag6(double* Ls, int ld, int j0, double* dinv) {
  double* r0 = Ls + (size_t)j0 * ld + j0;
  double *r1 = r0 + ld, *r2 = r1 + ld, *r3 = r2 + ld, *r4 = r3 + ld, *r5 = r4 + ld;
  const double a00 = r0[0];
  const double a10 = r1[0], a11 = r1[1];
  const double a20 = r2[0], a21 = r2[1], a22 = r2[2];
  const double a30 = r3[0], a31 = r3[1], a32 = r3[2], a33 = r3[3];
  const double a40 = r4[0], a41 = r4[1], a42 = r4[2], a43 = r4[3], a44 = r4[4];
  const double a50 = r5[0], a51 = r5[1], a52 = r5[2], a53 = r5[3], a54 = r5[4], a55 = r5[5];
  const double i0 = rsqrt_pos(a00), l00 = a00 * i0;
  const double l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0, l40 = a40 * i0, l50 = a50 * i0;
  const double d1 = a11 - l10 * l10;
  const double i1 = rsqrt_pos(d1), l11 = d1 * i1;
  const double l21 = (a21 - l20 * l10) * i1, l31 = (a31 - l30 * l10) * i1, l41 = (a41 - l40 * l10) * i1, l51 = (a51 - l50 * l10) * i1;
  const double d2 = a22 - l20 * l20 - l21 * l21;
  const double i2 = rsqrt_pos(d2), l22 = d2 * i2;
  const double l32 = (a32 - l30 * l20 - l31 * l21) * i2, l42 = (a42 - l40 * l20 - l41 * l21) * i2, l52 = (a52 - l50 * l20 - l51 * l21) * i2;
  const double d3 = a33 - l30 * l30 - l31 * l31 - l32 * l32;
  const double i3 = rsqrt_pos(d3), l33 = d3 * i3;
  const double l43 = (a43 - l40 * l30 - l41 * l31 - l42 * l32) * i3, l53 = (a53 - l50 * l30 - l51 * l31 - l52 * l32) * i3;
  const double d4 = a44 - l40 * l40 - l41 * l41 - l42 * l42 - l43 * l43;
  const double i4 = rsqrt_pos(d4), l44 = d4 * i4;
  const double l54 = (a54 - l50 * l40 - l51 * l41 - l52 * l42 - l53 * l43) * i4;
  const double d5 = a55 - l50 * l50 - l51 * l51 - l52 * l52 - l53 * l53 - l54 * l54;
  const double i5 = rsqrt_pos(d5), l55 = d5 * i5;
  if (!(a00 > 0 && d1 > 0 && d2 > 0 && d3 > 0 && d4 > 0 && d5 > 0)) return false;
  r0[0] = l00;
  r1[0] = l10; r1[1] = l11;
  r2[0] = l20; r2[1] = l21; r2[2] = l22;
  r3[0] = l30; r3[1] = l31; r3[2] = l32; r3[3] = l33;
  r4[0] = l40; r4[1] = l41; r4[2] = l42; r4[3] = l43; r4[4] = l44;
  r5[0] = l50; r5[1] = l51; r5[2] = l52; r5[3] = l53; r5[4] = l54; r5[5] = l55;
  dinv[0] = i0; dinv[1] = i1; dinv[2] = i2; dinv[3] = i3; dinv[4] = i4; dinv[5] = i5;
  return true;
}

__device__ void phase_chol(const BaArgs& a, double lambda, int cur, double* Ls, double* Li, double* red, unsigned long long* tp) {
  const int n = 6 * a.W, ld = n + 1, tid = threadIdx.x, nt = blockDim.x, nb = a.W;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  double* ys = Ls + (size_t)n * ld;  // row n: right-hand side -> y -> x (phase_schur_reduce wrote S and b here)
  __syncthreads();
  if (tid == 0 && !chol_diag6(Ls, ld, 0, Li)) s_bad = 1;
  __syncthreads();
  ba_tick(tp, 11);
  for (int jb = 0; jb < nb; jb++) {
    if (s_bad) break;
    const int j0 = 6 * jb;
    // (1) panel: rows below the block (and the rhs row): solve row' L11^T = row
    {
      const double* D = Ls + j0 * ld + j0;
      const double* dv = Li + 6 * jb;
      const double l10 = D[ld], l20 = D[2 * ld], l21 = D[2 * ld + 1], l30 = D[3 * ld], l31 = D[3 * ld + 1], l32 = D[3 * ld + 2];
      const double l40 = D[4 * ld], l41 = D[4 * ld + 1], l42 = D[4 * ld + 2], l43 = D[4 * ld + 3];
      const double l50 = D[5 * ld], l51 = D[5 * ld + 1], l52 = D[5 * ld + 2], l53 = D[5 * ld + 3], l54 = D[5 * ld + 4];
      const double d0 = dv[0], d1 = dv[1], d2 = dv[2], d3 = dv[3], d4 = dv[4], d5 = dv[5];
      for (int i = j0 + 6 + tid; i <= n; i += nt) {
        double* row = Ls + i * ld + j0;
        const double o0 = row[0] * d0;
        const double o1 = (row[1] - o0 * l10) * d1;
        const double o2 = (row[2] - o0 * l20 - o1 * l21) * d2;
        const double o3 = (row[3] - o0 * l30 - o1 * l31 - o2 * l32) * d3;
        const double o4 = (row[4] - o0 * l40 - o1 * l41 - o2 * l42 - o3 * l43) * d4;
        const double o5 = (row[5] - o0 * l50 - o1 * l51 - o2 * l52 - o3 * l53 - o4 * l54) * d5;
        row[0] = o0; row[1] = o1; row[2] = o2; row[3] = o3; row[4] = o4; row[5] = o5;
      }
    }
    __syncthreads();
    ba_tick(tp, 16);
    // (2) trailing update A22 -= L21 L21^T (lower triangle, plus the rhs row).  Warp 0: the next diagonal block and
    //     its factorisation; warps 1..: one row each, lanes over the columns <= row.
    const int m = n - j0 - 6;  // trailing matrix rows; local row m is the rhs
    if (warp == 0) {
      if (jb + 1 < nb) {
        if (lane < 21) {
          // row of lane l in the packed lower triangle (3 bits per lane in one constant: a decode loop costs ~150 cycles of
          // this critical chain); the 6-term dot product as a tree (depth 4 instead of 7)
          const int r = (int)((0x5b6db2491b6d2448ull >> (3 * lane)) & 7ull), c = lane - ((r * (r + 1)) >> 1);
          const double* lr = Ls + (j0 + 6 + r) * ld + j0;
          const double* lc = Ls + (j0 + 6 + c) * ld + j0;
          double* dst = Ls + (j0 + 6 + r) * ld + j0 + 6 + c;
          const double s01 = fma(lr[1], lc[1], lr[0] * lc[0]), s23 = fma(lr[3], lc[3], lr[2] * lc[2]), s45 = fma(lr[5], lc[5], lr[4] * lc[4]);
          *dst = *dst - ((s01 + s23) + s45);
        }
        __syncwarp();
        if (lane == 0 && !chol_diag6(Ls, ld, j0 + 6, Li + 6 * (jb + 1))) s_bad = 1;
        ba_tick(tp, 17);
      }
    } else {
      // groups of 8 rows (8 independent FMA chains per lane, the column operand is loaded once per group)
      // The cost of a group grows with its row index (columns <= row: 1..4 chunks of 32).  Groups are dealt heaviest first and
      // in alternating direction over the workers, so that the busiest warp has 6 chunks instead of 8 at the first step (the
      // FP64 pipe of this one SM bounds the early steps, the diagonal chain of warp 0 the late ones).
      const int ngroups = (m - 5 + 7) >> 3;  // local rows 6..m
      const int nwork = nwarp - 1, widx = warp - 1;
      for (int pass = 0; pass * nwork < ngroups; pass++) {
        const int gp = pass * nwork + ((pass & 1) ? nwork - 1 - widx : widx);
        if (gp >= ngroups) continue;
        const int g = ngroups - 1 - gp;
        const int r0 = 6 + 8 * g;
        double rv[8][6];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int r = min(r0 + i, m);
          const double* lr = Ls + (j0 + 6 + r) * ld + j0;
#pragma unroll
          for (int k = 0; k < 6; k++) rv[i][k] = lr[k];
        }
        const int rmax = min(r0 + 7, m), cmax = min(rmax, m - 1);
        for (int c = lane; c <= cmax; c += 32) {
          const double* lc = Ls + (j0 + 6 + c) * ld + j0;
          const double l0 = lc[0], l1 = lc[1], l2 = lc[2], l3 = lc[3], l4 = lc[4], l5 = lc[5];
          // load the 8 destinations, update, store: 8 independent chains (a read-modify-write per row would serialise
          // on the possible aliasing of consecutive rows)
          double cv[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int r = r0 + i;
            cv[i] = (r <= m && c <= r) ? Ls[(j0 + 6 + r) * ld + j0 + 6 + c] : 0.0;
          }
#pragma unroll
          for (int i = 0; i < 8; i++)
            cv[i] -= rv[i][0] * l0 + rv[i][1] * l1 + rv[i][2] * l2 + rv[i][3] * l3 + rv[i][4] * l4 + rv[i][5] * l5;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int r = r0 + i;
            if (r <= m && c <= r) Ls[(j0 + 6 + r) * ld + j0 + 6 + c] = cv[i];
          }
        }
      }
    }
    __syncthreads();
    ba_tick(tp, 18);
  }
  __syncthreads();
  ba_tick(tp, 12);
  const int failed = s_bad;
  if (!failed && warp == 0) {
    // back substitution with L^T by one warp (no block barriers).  Per block step every lane solves L11^T x_b = y_b
    // redundantly (broadcast loads, 6 dependent steps) and updates its rows above.  All loads of a step -- the block, y_b and
    // the lane's rows of the panel -- are issued BEFORE the solve: the stores of the previous step forbid the compiler to hoist
    // them itself, and a chunk-by-chunk load / update / store sequence serialises (measured 920 cycles per block step, 420
    // like this).  The updates are independent chains that end with the unknown that is known last.
    for (int jb = nb - 1; jb >= 0; jb--) {
      const int j0 = 6 * jb;
      const double* D = Ls + j0 * ld + j0;
      const double* dv = Li + 6 * jb;
      const double y0 = ys[j0], y1 = ys[j0 + 1], y2 = ys[j0 + 2], y3 = ys[j0 + 3], y4 = ys[j0 + 4], y5 = ys[j0 + 5];
      const double d0 = dv[0], d1 = dv[1], d2 = dv[2], d3 = dv[3], d4 = dv[4], d5 = dv[5];
      const double l10 = D[ld], l20 = D[2 * ld], l21 = D[2 * ld + 1], l30 = D[3 * ld], l31 = D[3 * ld + 1], l32 = D[3 * ld + 2];
      const double l40 = D[4 * ld], l41 = D[4 * ld + 1], l42 = D[4 * ld + 2], l43 = D[4 * ld + 3];
      const double l50 = D[5 * ld], l51 = D[5 * ld + 1], l52 = D[5 * ld + 2], l53 = D[5 * ld + 3], l54 = D[5 * ld + 4];
      constexpr int NCH = (6 * BA_MAX_W + 31) / 32;   // rows above the block, 32 per chunk
      double dl[NCH][6], yv[NCH];
#pragma unroll
      for (int u = 0; u < NCH; u++) {
        const int i = lane + 32 * u;
        const bool on = i < j0;
        yv[u] = on ? ys[i] : 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) dl[u][k] = on ? D[k * ld + i - j0] : 0.0;
      }
      const double x5 = y5 * d5;
      const double x4 = (y4 - l54 * x5) * d4;
      const double x3 = ((y3 - l53 * x5) - l43 * x4) * d3;
      const double x2 = (((y2 - l52 * x5) - l42 * x4) - l32 * x3) * d2;
      const double x1 = ((((y1 - l51 * x5) - l41 * x4) - l31 * x3) - l21 * x2) * d1;
      const double x0 = (((((y0 - l50 * x5) - l40 * x4) - l30 * x3) - l20 * x2) - l10 * x1) * d0;
#pragma unroll
      for (int u = 0; u < NCH; u++) {
        const int i = lane + 32 * u;
        double t = yv[u];
        t = fma(-dl[u][5], x5, t); t = fma(-dl[u][4], x4, t); t = fma(-dl[u][3], x3, t);
        t = fma(-dl[u][2], x2, t); t = fma(-dl[u][1], x1, t); t = fma(-dl[u][0], x0, t);
        if (i < j0) ys[i] = t;
      }
      if (lane < 6) {
        double xv = x0;
        xv = lane == 1 ? x1 : xv; xv = lane == 2 ? x2 : xv; xv = lane == 3 ? x3 : xv; xv = lane == 4 ? x4 : xv; xv = lane == 5 ? x5 : xv;
        ys[j0 + lane] = xv;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  ba_tick(tp, 13);
  // increments, trial poses, pose part of the scale (x = b when the solver failed, like LinearSolverCSparse)
  const int trial = cur ^ 1;
  double sc = 0;
  for (int i = tid; i < n; i += nt) {
    const double x = failed ? a.bp[i] : ys[i];
    a.xp[i] = x;
    sc += x * (lambda * x + a.bp[i]);
  }
  __syncthreads();
  for (int p = tid; p < a.W; p += nt) {
    Pose o;
    pose_oplus(a.X[(size_t)cur * a.W + p], a.xp + 6 * p, o);
    a.X[(size_t)trial * a.W + p] = o;
  }
  double v[1] = {sc};
  block_reduce<1, false>(v, red);
  if (tid == 0) { a.cinfo[0] = red[0]; a.cinfo[1] = failed ? 1.0 : 0.0; }
}

// back-substitution of the points (thread per point): x_l = (bl - sum_o Hpl(o)^T xp) / (hl + lambda) with
// Hpl^T xp = w R (-x_t + 2 zc x x_r); trial points, point part of the scale.  The chi2 of the trial state comes from
// the observation pass that follows.
__device__ void phase_update(const BaArgs& a, double lambda, int cur, int failed, int G, int GT, int rank, double* red,
                             const BaTab& tb, double* spose /* [W][16] shared scratch */) {
  const int trial = cur ^ 1, W = a.W;
  const size_t M = a.M;
  const double* pts = a.pts + (size_t)cur * 3 * a.P;
  double* ptt = a.pts + (size_t)trial * 3 * a.P;
  const double* ow = a.ow + (size_t)cur * M;
  const double* ozc = a.ozc + (size_t)cur * 3 * M;
  // per pose: rotation (9), -x_t (3), x_r (3) in shared memory
  for (int i = threadIdx.x; i < W * 16; i += blockDim.x) {
    const int p = i >> 4, k = i & 15;
    double v = 0;
    if (k < 9) v = a.X[(size_t)cur * W + p].R[k];
    else if (k < 12) v = -a.xp[6 * p + (k - 9)];
    else if (k < 15) v = a.xp[6 * p + 3 + (k - 12)];
    spose[i] = v;
  }
  __syncthreads();
  double scale = 0;
  for (int l = G; l < a.P; l += GT) {
    const double* b = a.bl + 3 * (size_t)l;
    const int f = a.pt_first[l], len = a.pt_len[l], i = l - tb.grp[f];
    double x[3];
    if (failed) { x[0] = b[0]; x[1] = b[1]; x[2] = b[2]; }
    else {
      double c0 = b[0], c1 = b[1], c2 = b[2];
      for (int k0 = 0; k0 < len; k0 += 4) {
        double w[4], zx[4], zy[4], zz[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const bool on = k0 + j < len;
          const int pp = on ? f + k0 + j : f;
          const size_t o = (size_t)tb.base[pp] + tb.off[pp * (W + 1) + f] + i;
          w[j] = on ? ow[o] : 0.0;
          zx[j] = on ? ozc[o] : 0.0; zy[j] = on ? ozc[M + o] : 0.0; zz[j] = on ? ozc[2 * M + o] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int pp = (k0 + j < len) ? f + k0 + j : f;
          const double* sp = spose + 16 * pp;
          const double qx = 2 * zx[j], qy = 2 * zy[j], qz = 2 * zz[j];
          const double g0 = w[j] * (qy * sp[14] - qz * sp[13] + sp[9]), g1 = w[j] * (qz * sp[12] - qx * sp[14] + sp[10]),
                       g2 = w[j] * (qx * sp[13] - qy * sp[12] + sp[11]);
          c0 -= sp[0] * g0 + sp[1] * g1 + sp[2] * g2;
          c1 -= sp[3] * g0 + sp[4] * g1 + sp[5] * g2;
          c2 -= sp[6] * g0 + sp[7] * g1 + sp[8] * g2;
        }
      }
      const double s = div_pos(1.0, a.hl[l] + lambda);
      x[0] = s * c0; x[1] = s * c1; x[2] = s * c2;
    }
    for (int k = 0; k < 3; k++) {
      ptt[3 * (size_t)l + k] = pts[3 * (size_t)l + k] + x[k];
      scale += x[k] * (lambda * x[k] + b[k]);
    }
  }
  double v[1] = {scale};
  block_reduce<1, false>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 1] = red[0];
}

__device__ void phase_output(const BaArgs& a, int cur, int G, int GT) {
  const Pose* X = a.X + (size_t)cur * a.W;
  const double* pts = a.pts + (size_t)cur * 3 * a.P;
  for (int i = G; i < a.W; i += GT) pose_to_f32(X[i], a.out_poses + 16 * i);
  for (int i = G; i < 3 * a.P; i += GT) a.out_points[i] = (float)pts[i];
}

// relative motions from the float32 poses: Converter::toInvMatrix(pose[i-1]) * pose[i], cv::Mat CV_32F semantics
// (double accumulation, one rounding) (src/Optimizer.cc:1072-1075)
__device__ void phase_output_rel(const BaArgs& a, int G, int GT) {
  for (int i = 1 + G; i < a.W; i += GT) {
    const float* A = a.out_poses + 16 * (i - 1);
    const float* B = a.out_poses + 16 * i;
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
}

// ---------------------------------------------------------------------------------------------------------
// the cluster kernel (cluster size set at launch: 16 CTAs when the device allows it, else 8)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BA_THREADS, 1) ba_window_big_kernel(BaArgs a) {
  extern __shared__ __align__(16) double dsm[];  // (6W+1)^2 + 36W doubles for the dense factorisation (CTA 0)
  __shared__ double red[16 * 2 + 32];
  __shared__ LmCtl ctl;  // every CTA keeps an identical copy: decisions are recomputed from the same partial sums
  __shared__ BaTab tb;
  __shared__ double spose[BA_MAX_W * 16];
  __shared__ double smr[(BA_THREADS / 32) * 32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
  double* const Ls0 = cluster.map_shared_rank(dsm, 0);
  const int G = rank * blockDim.x + threadIdx.x, GT = nranks * blockDim.x;
  const int tid = threadIdx.x;
  // phase timers (thread 0 of CTA 0 only): slots 0 linearise, 2 schur, 3 solve, 4 update, 5 init, 6 LM bookkeeping, 7 total,
  // 8-10 linearise sub-steps, 11-14 solve sub-steps; slot 15 is the running time stamp
  __shared__ unsigned long long tph[24];
  if (G == 0) { for (int k = 0; k < 24; k++) tph[k] = 0; tph[15] = gtime(); }
  const unsigned long long t_start = (G == 0) ? tph[15] : 0;
  unsigned long long* const tp = (G == 0) ? tph : nullptr;
#define TOC(slot) ba_tick(tp, slot)

  for (int i = tid; i <= a.W; i += blockDim.x) { tb.grp[i] = a.grp_start[i]; tb.base[i] = a.pose_base[i]; }
  for (int i = tid; i < a.W * (a.W + 1); i += blockDim.x) { tb.cnt[i] = a.cnt_gt[i]; tb.off[i] = a.off[i]; }
  if (tid == 0) lm_reset(&ctl);
  __syncthreads();
  {  // Schur work units per job (the structure is fixed for the whole solve), then their prefix
    const int W = a.W, npairs = W * (W + 1) / 2, njobs = npairs + W;
    for (int job = tid; job < njobs; job += blockDim.x) {
      int terms = 0;
      if (job < npairs) {
        int p1, p2;
        job_pair(job, W, p1, p2);
        for (int f = 0; f <= p1; f++) terms += tb.cnt[f * (W + 1) + (p2 - f)];
      } else terms = tb.base[job - npairs + 1] - tb.base[job - npairs];
      tb.ustart[job + 1] = terms > 0 ? ((terms + 31) >> 5) + BA_JOB_PAD : 0;
      if (job < npairs) { int p1, p2; job_pair(job, W, p1, p2); tb.jp1[job] = (unsigned char)p1; tb.jp2[job] = (unsigned char)p2; }
    }
    __syncthreads();
    if (tid == 0) {
      tb.ustart[0] = 0;
      for (int j = 0; j < njobs; j++) tb.ustart[j + 1] += tb.ustart[j];
      const int nw = nranks * (blockDim.x >> 5);
      const int q = (tb.ustart[njobs] + nw - 1) / nw;
      tb.uq = q > 0 ? q : 1;
      int nch = W > 0 ? nw / W : 1;
      tb.nch = nch < 1 ? 1 : (nch > BA_PCHUNK ? BA_PCHUNK : nch);
    }
    __syncthreads();
  }
  phase_init(a, G, GT);
  cluster.sync();
  if (a.W + a.P == 0 || a.max_iterations <= 0) {
    if (G == 0) { ctl.iterations = (a.W + a.P == 0) ? -1 : 0; *a.ctl_out = ctl; }
    phase_output(a, 0, G, GT);
    cluster.sync();
    phase_output_rel(a, G, GT);
    return;
  }
  phase_obs(a, 0, G, GT, rank, nranks, red, tb);  // chi2 and linearisation of the initial state
  cluster.sync();
  if (tid == 0) {
    double c = 0;
    for (int r = 0; r < nranks; r++) c += a.part[r * 4];
    ctl.currentChi = c;
  }
  __syncthreads();
  TOC(5);

  for (int it = 0; it < a.max_iterations; it++) {
    if (ctl.stop_flag || !ctl.ok) break;
    const int cur = ctl.cur;  // state buffer == linearisation buffer
    phase_blocks(a, cur, G, GT, rank, nranks, red, tb);
    cluster.sync();
    TOC(10);
    if (tid == 0) {
      double m = 0;
      for (int r = 0; r < nranks; r++) m = fmax(m, a.part[r * 4 + 2]);
      for (int p = 0; p < a.W; p++) m = fmax(m, a.pmax[p]);
      lm_begin_iteration(&ctl, it, m, -1.0);
    }
    __syncthreads();
    TOC(0);
    while (true) {
      const double lambda = ctl.lambda;
      phase_schur_units(a, lambda, cur, rank, nranks, tb);
      if (a.t_detail) { __syncthreads(); TOC(19); }
      cluster.sync();
      TOC(1);
      phase_schur_reduce(a, lambda, cur, rank, nranks, tb, Ls0, smr);
      cluster.sync();
      TOC(2);
      if (rank == 0) {
        phase_chol(a, lambda, cur, dsm, dsm + (size_t)(6 * a.W + 1) * (6 * a.W + 1), red, tp);
      }
      cluster.sync();
      TOC(3);
      const int failed = a.cinfo[1] != 0.0;
      phase_update(a, lambda, cur, failed, G, GT, rank, red, tb, spose);
      if (a.t_detail) { __syncthreads(); TOC(21); }
      cluster.sync();
      TOC(4);
      phase_obs(a, cur ^ 1, G, GT, rank, nranks, red, tb);  // chi2 of the trial state + its linearisation
      if (a.t_detail) { __syncthreads(); TOC(20); }
      cluster.sync();
      TOC(9);
      if (tid == 0) {
        double chi = 0, scale = a.cinfo[0];
        for (int r = 0; r < nranks; r++) { chi += a.part[r * 4]; scale += a.part[r * 4 + 1]; }
        lm_trial(&ctl, chi, scale, failed);
      }
      __syncthreads();
      // part[] is next written two cluster barriers later (phase_update / phase_obs): no race with slower CTAs
      if (!lm_more_trials(&ctl)) break;
    }
    if (tid == 0) lm_end_iteration(&ctl, it, a.gain_threshold, rank == 0 ? a.rec : nullptr);
    __syncthreads();
    TOC(6);
  }
  phase_output(a, ctl.cur, G, GT);
  cluster.sync();
  phase_output_rel(a, G, GT);
  if (G == 0) {
    *a.ctl_out = ctl;
    tph[7] = gtime() - t_start;
    for (int k = 0; k < 24; k++) a.t_phase[k] = tph[k];
  }
#undef TOC
}

// =========================================================================================================
// host side
// =========================================================================================================
#define BA_DEPTH 3   // window solves that may be queued at once (staging slots)
struct BaWorkspace {
  int capW = 0, capP = 0, capM = 0;
  int cluster = 8;
  // shared-memory resident kernel (ba_window.cu): cluster size, dynamic shared memory it may use; per staging slot whether the
  // problem fits and what it needs.  Oversized windows fall back to the L2-resident kernel of this file.
  int cluster_sm = 8;
  size_t smem_limit = 0;
  bool use_sm2[BA_DEPTH] = {};
  size_t smem2[BA_DEPTH] = {};
  std::vector<int> wk_obs;
  int seq = 0;                       // launch counter of the shared-memory kernel (completion word in the pinned output mirror)
  int seq2[BA_DEPTH] = {};
  char* d_base = nullptr;            // solver workspace
  char* d_in2[BA_DEPTH] = {};  // input blocks (two: the next problem is staged while the current one is solved)
  char* d_out2[BA_DEPTH] = {}; // output blocks (two: a solve may be queued behind the one in flight)
  BaArgs args;
  char* h_in2[BA_DEPTH] = {};  // pinned mirrors
  char* h_out2[BA_DEPTH] = {};
  // chaining: a window queued behind the previous one takes the poses / odometry / points they share straight from the
  // previous solve's output block on the device (gather kernel on the BA stream), so consecutive solves run back to back
  int* d_chain2[BA_DEPTH] = {};   // [capW + capP]: source index in the previous output per pose / sorted point, -1 = host value
  int* h_chain2[BA_DEPTH] = {};
  bool chained2[BA_DEPTH] = {};
  cudaEvent_t out_done[BA_DEPTH] = {};
  struct Flight { int slot, W, P, M; bool want_records; bool sm; };
  Flight flight[BA_DEPTH];
  int nflight = 0;
  int slot = 0;                      // staging slot of the problem being prepared / in flight
  bool prepared = false;
  BaArgs a_prep;                     // kernel arguments of the prepared problem
  size_t in_bytes = 0, out_bytes = 0;
  // the solve runs on its own stream so that a caller may overlap it with other work (ba_submit ... ba_collect)
  cudaStream_t stream = nullptr;
  cudaStream_t up_stream = nullptr;  // uploads the structure part of a staged problem while the previous one is solved
  cudaEvent_t up_done = nullptr;
  size_t values_bytes = 0;           // leading part of the input block holding poses | odometry | points
  cudaEvent_t ev0[BA_DEPTH] = {}, ev1[BA_DEPTH] = {};
  std::vector<int> newid2[BA_DEPTH], oldid, first, len, last, keycnt;  // newid per staging slot (needed again at collect)
};

static size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

template <class T>
static T* carve(char*& p, size_t n) {
  T* r = (T*)p;
  p += al(sizeof(T) * n);
  return r;
}

// input block: carved identically on the pinned host staging buffer and on the device, so one H2D copy moves it all
static void carve_inputs(char*& p, BaArgs& a, int W, int P, int M) {
  a.poses_f32 = carve<float>(p, 16 * W); a.rel_f32 = carve<float>(p, 16 * W);
  a.points_f32 = carve<float>(p, 3 * (size_t)P);
  a.obs_pose = carve<int>(p, M); a.obs_point = carve<int>(p, M); a.obs_xyz = carve<float>(p, 3 * (size_t)M);
  a.pt_len = carve<int>(p, P); a.pt_first = carve<int>(p, P);
  a.grp_start = carve<int>(p, W + 1); a.cnt_gt = carve<int>(p, W * (W + 1));
  a.off = carve<int>(p, W * (W + 1)); a.pose_base = carve<int>(p, W + 1);
}
// output block: one D2H copy
static void carve_outputs(char*& p, BaArgs& a, int W, int P) {
  a.ctl_out = carve<LmCtl>(p, 1); a.t_phase = carve<unsigned long long>(p, 24);
  a.out_poses = carve<float>(p, 16 * W); a.out_rel = carve<float>(p, 16 * W); a.out_points = carve<float>(p, 3 * (size_t)P);
  a.rec = carve<LmRec>(p, VIDO_LM_REC);
}

static void carve_all(char*& p, BaArgs& a, int capW, int capP, int capM) {
  a.X = carve<Pose>(p, 2 * capW); a.Zinv = carve<Pose>(p, capW); a.pts = carve<double>(p, 6 * (size_t)capP);
  a.hl = carve<double>(p, capP); a.bl = carve<double>(p, 3 * (size_t)capP); a.ow = carve<double>(p, 2 * (size_t)capM); a.ozc = carve<double>(p, 6 * (size_t)capM); a.og = carve<double>(p, 6 * (size_t)capM); a.ohb = carve<double>(p, 4 * (size_t)capM);
  a.Hpp = carve<double>(p, 36 * capW); a.Hoff = carve<double>(p, 36 * capW); a.bp = carve<double>(p, 6 * capW);
  a.ppart = carve<double>(p, 2 * (size_t)capW * BA_PCHUNK * 28);
  a.spart = carve<double>(p, (size_t)BA_MAX_JOBS * BA_MAX_SLOTS * 16); a.pmax = carve<double>(p, capW); a.xp = carve<double>(p, 6 * capW);
  a.part = carve<double>(p, 4 * BA_MAX_CLUSTER); a.cinfo = carve<double>(p, 4);
  a.seJ = carve<double>(p, 2 * 72 * capW); a.seE = carve<double>(p, 2 * 8 * capW);
  a.wmom = carve<double>(p, (size_t)(BA_MAX_CLUSTER - 1) * BA_MAX_JOBS * 16);
  a.wpsum = carve<double>(p, (size_t)2 * (BA_MAX_CLUSTER - 1) * capW * 28);
  a.eH = carve<double>(p, (size_t)2 * capW * 120);
  a.Sg = carve<double>(p, (size_t)(6 * capW + 9) * (6 * capW + 9));
}

int ba_setup(vido_ctx* ctx, int capW, int capP, int capM) {
  BaWorkspace* ws = new BaWorkspace();
  ctx->ba = ws;
  if (capW > BA_MAX_W) capW = BA_MAX_W;
  ws->capW = capW; ws->capP = capP; ws->capM = capM;
  BaArgs tmp;
  char* p = nullptr;
  carve_all(p, tmp, capW, capP, capM);
  const size_t need = (size_t)p;
  p = nullptr; carve_inputs(p, tmp, capW, capP, capM); ws->in_bytes = (size_t)p;
  p = nullptr; carve_outputs(p, tmp, capW, capP); ws->out_bytes = (size_t)p;
  VIDO_CUDA(cudaMalloc(&ws->d_base, need));
  VIDO_CUDA(cudaMemset(ws->d_base, 0, need));
  for (int k = 0; k < BA_DEPTH; k++) {
    VIDO_CUDA(cudaMalloc(&ws->d_in2[k], ws->in_bytes));
    VIDO_CUDA(cudaMallocHost(&ws->h_in2[k], ws->in_bytes));
  }
  for (int k = 0; k < BA_DEPTH; k++) {
    VIDO_CUDA(cudaMalloc(&ws->d_out2[k], ws->out_bytes));
    VIDO_CUDA(cudaMallocHost(&ws->h_out2[k], ws->out_bytes + 64));
    memset(ws->h_out2[k], 0, ws->out_bytes + 64);
    VIDO_CUDA(cudaMalloc(&ws->d_chain2[k], sizeof(int) * (size_t)(capW + capP)));
    VIDO_CUDA(cudaMallocHost(&ws->h_chain2[k], sizeof(int) * (size_t)(capW + capP)));
    VIDO_CUDA(cudaEventCreateWithFlags(&ws->out_done[k], cudaEventDisableTiming));
    VIDO_CUDA(cudaEventCreate(&ws->ev0[k]));
    VIDO_CUDA(cudaEventCreate(&ws->ev1[k]));
  }
  memset(&ws->args, 0, sizeof ws->args);
  p = ws->d_base;
  carve_all(p, ws->args, capW, capP, capM);
  const size_t smem = sizeof(double) * ((size_t)(6 * capW + 1) * (6 * capW + 1) + 36 * capW);
  if (smem > 200 * 1024) { ctx->err = "BA window too large for the shared-memory Cholesky"; return VIDO_ERR_ARG; }
  VIDO_CUDA(cudaFuncSetAttribute(ba_window_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // 16-CTA clusters are a non-portable size: opt in, and verify that one fits
  ws->cluster = 8;
  if (cudaFuncSetAttribute(ba_window_big_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(BA_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, ba_window_big_kernel, &cfg) == cudaSuccess && nclusters >= 1) ws->cluster = 16;
  }
  cudaGetLastError();
  if (getenv("VIDO_BA_CLUSTER") && atoi(getenv("VIDO_BA_CLUSTER")) == 8) ws->cluster = 8;
  ws->smem_limit = ba_window_smem_limit();
  if (ws->smem_limit == 0 || ba_window_configure(ws->smem_limit, &ws->cluster_sm) != 0) { ctx->err = "cannot configure the window-BA kernel"; return VIDO_ERR_CUDA; }
  if (getenv("VIDO_BA_CLUSTER") && atoi(getenv("VIDO_BA_CLUSTER")) == 8) ws->cluster_sm = 8;
  VIDO_CUDA(vido_create_stream(&ws->stream, true));
  VIDO_CUDA(vido_create_stream(&ws->up_stream, true));
  VIDO_CUDA(cudaEventCreateWithFlags(&ws->up_done, cudaEventDisableTiming));
  return VIDO_OK;
}

void ba_teardown(vido_ctx* ctx) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  if (!ws) return;
  if (ws->stream) { cudaStreamSynchronize(ws->stream); cudaStreamDestroy(ws->stream); }
  if (ws->up_stream) { cudaStreamSynchronize(ws->up_stream); cudaStreamDestroy(ws->up_stream); }
  if (ws->up_done) cudaEventDestroy(ws->up_done);
  for (int k = 0; k < BA_DEPTH; k++) {
    if (ws->ev0[k]) cudaEventDestroy(ws->ev0[k]);
    if (ws->ev1[k]) cudaEventDestroy(ws->ev1[k]);
    if (ws->out_done[k]) cudaEventDestroy(ws->out_done[k]);
    cudaFree(ws->d_out2[k]); cudaFreeHost(ws->h_out2[k]); cudaFree(ws->d_chain2[k]); cudaFreeHost(ws->h_chain2[k]);
  }
  cudaFree(ws->d_base);
  for (int k = 0; k < BA_DEPTH; k++) { cudaFree(ws->d_in2[k]); cudaFreeHost(ws->h_in2[k]); }
  delete ws;
  ctx->ba = nullptr;
}

// The solve is split in three host steps so that a caller can hide everything but the solve itself:
//   ba_prepare: lay the problem's STRUCTURE and observations out in the free staging slot (may run while the previous
//               problem is still being solved; pr->poses / rel_motion / points are not read);
//   ba_launch:  add the state values (poses, odometry, points), copy the block and launch on the BA stream;
//   ba_collect: wait and write the results back into the problem's arrays, which must stay alive in between.
// values of window k+1 that are outputs of window k, copied on the device (same float32 bits as the host round trip)
__global__ void __launch_bounds__(256) ba_chain_kernel(const int* __restrict__ chain, int W, int P, const float* __restrict__ prev_poses,
                                                       const float* __restrict__ prev_rel, const float* __restrict__ prev_points,
                                                       float* __restrict__ poses, float* __restrict__ rel, float* __restrict__ points) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int np = 16 * W, nr = 16 * (W - 1);
  if (i < np) {
    const int src = chain[i >> 4];
    if (src >= 0) poses[i] = prev_poses[16 * src + (i & 15)];
  } else if (i < np + nr) {
    const int k = i - np, j = k >> 4;
    const int a = chain[j], b = chain[j + 1];
    if (a >= 0 && b == a + 1) rel[k] = prev_rel[16 * a + (k & 15)];
  } else if (i < np + nr + 3 * P) {
    const int k = i - np - nr, n = k / 3;
    const int src = chain[W + n];
    if (src >= 0) points[k] = prev_points[3 * src + (k - 3 * n)];
  }
}

int ba_prepare_chained(vido_ctx* ctx, const vido_ba_problem* pr, const int* prev_pose, const int* prev_point);
int ba_prepare(vido_ctx* ctx, const vido_ba_problem* pr) { return ba_prepare_chained(ctx, pr, nullptr, nullptr); }

// prev_pose[i] / prev_point[l]: index of pose i / point l (caller numbering) in the problem that is IN FLIGHT right now, or -1;
// nullptr: no chaining (every value comes from pr at launch time)
int ba_prepare_chained(vido_ctx* ctx, const vido_ba_problem* pr, const int* prev_pose, const int* prev_point) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  ws->prepared = false;
  const int W = pr->n_poses, P = pr->n_points, M = pr->n_obs;
  if (W < 0 || P < 0 || M < 0) return VIDO_ERR_ARG;
  if (W > ws->capW || P > ws->capP || M > ws->capM) { ctx->err = "BA problem exceeds the context capacity"; return VIDO_ERR_CAPACITY; }
  if (ws->nflight >= BA_DEPTH) { ctx->err = "too many window solves are queued"; return VIDO_ERR_STATE; }
  const int slot = ws->nflight ? (ws->flight[ws->nflight - 1].slot + 1) % BA_DEPTH : ws->slot;
  ws->slot = slot;
  char* const h_in = ws->h_in2[slot];
  char* const d_in = ws->d_in2[slot];
  BaArgs a = ws->args;
  a.W = W; a.P = P; a.M = M;
  a.max_iterations = pr->max_iterations;
  a.t_detail = (getenv("VIDO_BA_TIMING") && atoi(getenv("VIDO_BA_TIMING")) > 1) ? 1 : 0;
  a.info_cam = 1.0 / (double)pr->sigma2_cam;
  a.info_3d = 1.0 / (double)pr->sigma2_3d;
  a.d_cam = (double)pr->huber_cam;
  a.d_3d = (double)pr->huber_3d;
  a.gain_threshold = (double)pr->gain_threshold;
  // ---- host-side layout: tracks sorted by (first pose, length descending); observations pose-major (see the header)
  BaArgs h;  // host view of the input block (same carving as the device view)
  {
    char* hp = h_in; carve_inputs(hp, h, W, P, M);
    char* dp = d_in; carve_inputs(dp, a, W, P, M);
    char* dq = ws->d_out2[slot]; carve_outputs(dq, a, W, P);
  }
  int* h_obs_pose = (int*)h.obs_pose; int* h_obs_point = (int*)h.obs_point; float* h_xyz = (float*)h.obs_xyz;
  int* h_pt_len = (int*)h.pt_len; int* h_pt_first = (int*)h.pt_first; float* h_pts = (float*)h.points_f32;
  int* h_grp = (int*)h.grp_start; int* h_cnt = (int*)h.cnt_gt; int* h_off = (int*)h.off; int* h_base = (int*)h.pose_base;
  std::vector<int>&first = ws->first, &len = ws->len, &last = ws->last, &keycnt = ws->keycnt, &newid = ws->newid2[slot], &oldid = ws->oldid;
  first.assign(P, 1 << 30); len.assign(P, 0); last.assign(P, -1);
  for (int o = 0; o < M; o++) {
    const int l = pr->obs_point[o], p = pr->obs_pose[o];
    if (l < 0 || l >= P || p < 0 || p >= W) { ctx->err = "BA observation index out of range"; return VIDO_ERR_ARG; }
    first[l] = std::min(first[l], p);
    last[l] = std::max(last[l], p);
    len[l]++;
  }
  for (int l = 0; l < P; l++)
    if (len[l] == 0 || last[l] - first[l] + 1 != len[l]) {
      ctx->err = "BA tracks must observe consecutive poses (one observation per pose); the reference's graph always does";
      return VIDO_ERR_ARG;
    }
  // counting sort of the points by key = first * (W+1) + (W - len)
  keycnt.assign((size_t)W * (W + 1) + 2, 0); newid.resize(P); oldid.resize(P);
  for (int l = 0; l < P; l++) keycnt[(size_t)first[l] * (W + 1) + (W - len[l]) + 1]++;
  for (size_t k = 1; k < keycnt.size(); k++) keycnt[k] += keycnt[k - 1];
  for (int l = 0; l < P; l++) {
    const int n = keycnt[(size_t)first[l] * (W + 1) + (W - len[l])]++;
    newid[l] = n;
    oldid[n] = l;
  }
  for (int f = 0; f <= W; f++) h_grp[f] = 0;
  for (int i = 0; i < W * (W + 1); i++) { h_cnt[i] = 0; h_off[i] = 0; }
  for (int l = 0; l < P; l++) {
    h_grp[first[l] + 1]++;
    for (int L = 0; L < len[l] && L <= W; L++) h_cnt[first[l] * (W + 1) + L]++;  // length > L
  }
  for (int f = 0; f < W; f++) h_grp[f + 1] += h_grp[f];
  h_base[0] = 0;
  for (int p = 0; p < W; p++) {
    int acc = 0;
    for (int f = 0; f <= p; f++) { h_off[p * (W + 1) + f] = acc; acc += h_cnt[f * (W + 1) + (p - f)]; }
    h_off[p * (W + 1) + p + 1] = acc;
    h_base[p + 1] = h_base[p] + acc;
  }
  for (int n = 0; n < P; n++) {
    const int l = oldid[n];
    h_pt_len[n] = len[l];
    h_pt_first[n] = first[l];
  }
  for (int o = 0; o < M; o++) {
    const int l = pr->obs_point[o], p = pr->obs_pose[o], n = newid[l], f = first[l];
    const int q = h_base[p] + h_off[p * (W + 1) + f] + (n - h_grp[f]);
    h_obs_pose[q] = p;
    h_obs_point[q] = n;
    h_xyz[q] = pr->obs_xyz[3 * o]; h_xyz[(size_t)M + q] = pr->obs_xyz[3 * o + 1]; h_xyz[2 * (size_t)M + q] = pr->obs_xyz[3 * o + 2];
  }
  {  // everything behind the state values is final: upload it now (own stream, other staging slot than the solve in flight)
    char* hp = h_in; BaArgs t2; carve_inputs(hp, t2, W, P, M);
    const size_t used = (size_t)(hp - h_in);
    const size_t vb = (size_t)((const char*)h.obs_pose - h_in);  // poses_f32 | rel_f32 | points_f32 come first
    ws->values_bytes = vb;
    VIDO_CUDA(cudaMemcpyAsync(d_in + vb, h_in + vb, used - vb, cudaMemcpyHostToDevice, ws->up_stream));
    ws->chained2[slot] = false;
    if (prev_pose && (prev_point || P == 0) && ws->nflight >= 1) {
      const BaWorkspace::Flight& Fp = ws->flight[ws->nflight - 1];   // the newest solve in the queue
      const std::vector<int>& pnew = ws->newid2[Fp.slot];
      int* hc = ws->h_chain2[slot];
      for (int i = 0; i < W; i++) hc[i] = (prev_pose[i] >= 0 && prev_pose[i] < Fp.W) ? prev_pose[i] : -1;
      for (int l = 0; l < P; l++) hc[W + newid[l]] = (prev_point[l] >= 0 && prev_point[l] < Fp.P) ? pnew[prev_point[l]] : -1;
      VIDO_CUDA(cudaMemcpyAsync(ws->d_chain2[slot], hc, sizeof(int) * (size_t)(W + P), cudaMemcpyHostToDevice, ws->up_stream));
      ws->chained2[slot] = true;
    }
    VIDO_CUDA(cudaEventRecord(ws->up_done, ws->up_stream));
  }
  {  // does the window fit the shared-memory resident kernel?  Points are dealt round-robin (sorted order) to the workers.
    const int nwk = ws->cluster_sm - 1;
    ws->wk_obs.assign(2 * (size_t)nwk, 0);
    for (int n = 0; n < P; n++) {
      const int L = h_pt_len[n];
      ws->wk_obs[n % nwk] += L;
      ws->wk_obs[nwk + n % nwk] += L * (L + 1) / 2;   // Schur terms: pose pairs of the track
    }
    int capO = 0, capT = 0;
    for (int c = 0; c < nwk; c++) { capO = std::max(capO, ws->wk_obs[c]); capT = std::max(capT, ws->wk_obs[nwk + c]); }
    const int capPt = (P + nwk - 1) / nwk;
    a.capO = (capO + 7) & ~7; a.capPt = (capPt + 7) & ~7; a.capT = (capT + 7) & ~7;
    const size_t need = ba_window_smem(W, a.capO, a.capPt, a.capT);
    const char* force = getenv("VIDO_BA_KERNEL");   // debug: "l2" forces the L2-resident kernel
    ws->use_sm2[slot] = need <= ws->smem_limit && capO < 65536 && !(force && !strcmp(force, "l2"));
    ws->smem2[slot] = need;
  }
  ws->a_prep = a;
  ws->prepared = true;
  return VIDO_OK;
}

int ba_launch(vido_ctx* ctx, const vido_ba_problem* pr, bool want_records) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  if (!ws->prepared) { ctx->err = "ba_launch without ba_prepare"; return VIDO_ERR_ARG; }
  if (ws->nflight >= BA_DEPTH) { ctx->err = "too many window solves are queued"; return VIDO_ERR_STATE; }
  if (ws->nflight >= 1 && !ws->chained2[ws->slot]) { ctx->err = "a window BA is already in flight"; return VIDO_ERR_ARG; }
  ws->prepared = false;
  const int W = pr->n_poses, P = pr->n_points, M = pr->n_obs;
  const int slot = ws->slot;
  cudaStream_t s = ws->stream;
  const BaArgs a = ws->a_prep;
  {  // state values
    BaArgs h;
    char* hp = ws->h_in2[slot]; carve_inputs(hp, h, W, P, M);
    float* h_poses = (float*)h.poses_f32; float* h_rel = (float*)h.rel_f32; float* h_pts = (float*)h.points_f32;
    memcpy(h_poses, pr->poses, sizeof(float) * 16 * W);
    if (W > 1) memcpy(h_rel, pr->rel_motion, sizeof(float) * 16 * (W - 1));
    const std::vector<int>& newid = ws->newid2[slot];
    for (int l = 0; l < P; l++) {
      const int n = newid[l];
      h_pts[3 * n] = pr->points[3 * l]; h_pts[3 * n + 1] = pr->points[3 * l + 1]; h_pts[3 * n + 2] = pr->points[3 * l + 2];
    }
  }
  // the state values travel on the upload stream too (behind the structure part): with a solve in flight the BA stream
  // then only has to run the gather and the kernel
  VIDO_CUDA(cudaMemcpyAsync(ws->d_in2[slot], ws->h_in2[slot], ws->values_bytes, cudaMemcpyHostToDevice, ws->up_stream));
  VIDO_CUDA(cudaEventRecord(ws->up_done, ws->up_stream));
  if (ws->use_sm2[slot]) {
    // Shared-memory resident kernel: nothing but solver kernels on the solver's stream.  The uploads are awaited by the host
    // (a few tens of microseconds, the host is about to launch anyway), the values shared with the solve in front are gathered
    // inside the kernel, the results come back through the pinned mirror + completion word.  Consecutive solves are launched
    // with programmatic stream serialisation: the prologue of this one overlaps the tail of the previous one.
    VIDO_CUDA(cudaEventSynchronize(ws->up_done));
    BaArgs b = a;
    const bool chained = ws->nflight >= 1 && ws->chained2[slot];
    if (chained) {
      const BaWorkspace::Flight& Fp = ws->flight[ws->nflight - 1];
      BaArgs po;
      { char* q = ws->d_out2[Fp.slot]; carve_outputs(q, po, Fp.W, Fp.P); }
      b.chain = ws->d_chain2[slot]; b.prev_poses = po.out_poses; b.prev_rel = po.out_rel; b.prev_points = po.out_points;
    } else { b.chain = nullptr; b.prev_poses = b.prev_rel = b.prev_points = nullptr; }
    BaArgs ho;
    char* hq = ws->h_out2[slot]; carve_outputs(hq, ho, W, P);
    const size_t out_used = want_records ? (size_t)(hq - ws->h_out2[slot]) : (size_t)((char*)ho.rec - ws->h_out2[slot]);
    b.out_base = ws->d_out2[slot]; b.h_out = ws->h_out2[slot]; b.out_bytes = (int)((out_used + 15) & ~(size_t)15);
    b.h_flag = (volatile int*)(ws->h_out2[slot] + ws->out_bytes);
    b.seq = ++ws->seq;
    ws->seq2[slot] = b.seq;
    VIDO_CUDA(ba_window_launch(b, ws->cluster_sm, ws->smem2[slot], s, ws->nflight >= 1));
    ctx->launches++;
    ws->flight[ws->nflight++] = {slot, W, P, M, want_records, true};
    return VIDO_OK;
  }
  VIDO_CUDA(cudaStreamWaitEvent(s, ws->up_done, 0));
  if (ws->nflight >= 1 && ws->chained2[slot]) {
    // queued behind the solve in flight (same stream, so it runs after it): shared values come from its output block
    const BaWorkspace::Flight& Fp = ws->flight[ws->nflight - 1];
    BaArgs po;
    { char* q = ws->d_out2[Fp.slot]; carve_outputs(q, po, Fp.W, Fp.P); }
    const int total = 16 * W + 16 * std::max(W - 1, 0) + 3 * P;
    ba_chain_kernel<<<(total + 255) / 256, 256, 0, s>>>(ws->d_chain2[slot], W, P, po.out_poses, po.out_rel, po.out_points,
                                                        (float*)a.poses_f32, (float*)a.rel_f32, (float*)a.points_f32);
    ctx->launches++;
  }
  const size_t smem = sizeof(double) * ((size_t)(6 * W + 1) * (6 * W + 1) + 36 * W);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ws->cluster); cfg.blockDim = dim3(BA_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ws->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEventRecord(ws->ev0[slot], s);
    VIDO_CUDA(cudaLaunchKernelEx(&cfg, ba_window_big_kernel, a));
    cudaEventRecord(ws->ev1[slot], s);
    ctx->launches++;
  }
  {
    BaArgs ho;
    char* hq = ws->h_out2[slot]; carve_outputs(hq, ho, W, P);
    // the LM records sit at the end of the block: copy them only when asked for
    const size_t out_used = want_records ? (size_t)(hq - ws->h_out2[slot]) : (size_t)((char*)ho.rec - ws->h_out2[slot]);
    VIDO_CUDA(cudaMemcpyAsync(ws->h_out2[slot], ws->d_out2[slot], out_used, cudaMemcpyDeviceToHost, s));
    VIDO_CUDA(cudaEventRecord(ws->out_done[slot], s));
  }
  ws->flight[ws->nflight++] = {slot, W, P, M, want_records, ws->use_sm2[slot]};
  return VIDO_OK;
}

int ba_submit(vido_ctx* ctx, const vido_ba_problem* pr, bool want_records) {
  int rc = ba_prepare(ctx, pr);
  if (rc) return rc;
  return ba_launch(ctx, pr, want_records);
}

int ba_collect(vido_ctx* ctx, vido_ba_problem* pr, vido_lm_stats* st) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  if (ws->nflight == 0) { ctx->err = "no window BA in flight"; return VIDO_ERR_ARG; }
  const BaWorkspace::Flight F = ws->flight[0];   // the oldest
  for (int k = 0; k + 1 < ws->nflight; k++) ws->flight[k] = ws->flight[k + 1];
  ws->nflight--;
  const int W = F.W, P = F.P, M = F.M;
  const std::vector<int>& newid = ws->newid2[F.slot];
  if (F.sm) {
    // completion word of the pinned output mirror (written by the kernel after the mirror itself)
    volatile int* flag = (volatile int*)(ws->h_out2[F.slot] + ws->out_bytes);
    const int want = ws->seq2[F.slot];
    long spins = 0;
    while (*flag != want) {
      if ((++spins & 0xfff) == 0) {
        const cudaError_t q = cudaStreamQuery(ws->stream);
        if (q != cudaSuccess && q != cudaErrorNotReady) { ctx->err = std::string("window BA kernel failed: ") + cudaGetErrorString(q); return VIDO_ERR_CUDA; }
        if (q == cudaSuccess && *flag != want) { ctx->err = "window BA finished without publishing its results"; return VIDO_ERR_CUDA; }
      }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
  } else {
    VIDO_CUDA(cudaEventSynchronize(ws->out_done[F.slot]));
  }
  BaArgs ho;
  { char* hq = ws->h_out2[F.slot]; carve_outputs(hq, ho, W, P); }
  const LmCtl ctl = *ho.ctl_out;
  const unsigned long long* tph = ho.t_phase;
  const LmRec* recs = ho.rec;
  memcpy(pr->poses, ho.out_poses, sizeof(float) * 16 * W);
  if (W > 1) memcpy(pr->rel_motion, ho.out_rel, sizeof(float) * 16 * (W - 1));
  const float* opts = ho.out_points;
  for (int l = 0; l < P; l++) {
    const int n = newid[l];
    pr->points[3 * l] = opts[3 * n]; pr->points[3 * l + 1] = opts[3 * n + 1]; pr->points[3 * l + 2] = opts[3 * n + 2];
  }
  {
    float ms = 0;
    if (F.sm) { ctx->t_ms[3] += (double)tph[7] * 1e-6; ctx->t_n[3]++; }   // the kernel's own %globaltimer span (no events on its stream)
    else if (cudaEventElapsedTime(&ms, ws->ev0[F.slot], ws->ev1[F.slot]) == cudaSuccess) { ctx->t_ms[3] += ms; ctx->t_n[3]++; }
    const double edges = (double)M + (double)std::max(W - 1, 0);
    ctx->ba_alg_bytes += edges * (296.0 * std::max(ctl.iterations, 0) + 152.0 * (ctl.total_trials + 1));
  }
  if (getenv("VIDO_BA_TIMING") && F.sm) {   // gap on the solver stream between the end of the solve in front and this one's release
    static unsigned long long last_end = 0, gap_sum = 0, blocked_sum = 0, sync_sum = 0, mirror_sum = 0; static int gap_n = 0;
    if (last_end && tph[16] > last_end && tph[16] - last_end < 5000000ull) {
      gap_sum += tph[16] - last_end; blocked_sum += tph[18]; gap_n++;
      sync_sum += tph[19] - tph[17]; mirror_sum += tph[20] - tph[19];
    }
    last_end = tph[17];
    if (atoi(getenv("VIDO_BA_TIMING")) > 1) {
      static unsigned long long t0 = 0;
      if (!t0) t0 = tph[21];
      fprintf(stderr, "[ba-line] seq %d its %d: begin %.1f wait-enter %.1f release %.1f end %.1f mirrored %.1f us\n", ws->seq2[F.slot], ctl.iterations,
              1e-3 * (double)(tph[21] - t0), 1e-3 * (double)(tph[22] - t0), 1e-3 * (double)(tph[16] - t0), 1e-3 * (double)(tph[17] - t0), 1e-3 * (double)(tph[20] - t0));
    }
    if (gap_n && gap_n % 64 == 0)
      fprintf(stderr, "[ba-gap] mean gap end->release %.1f us (own tail: closing barrier %.1f us, mirror copy %.1f us), blocked in griddepcontrol.wait %.1f us over %d solves\n",
              1e-3 * gap_sum / gap_n, 1e-3 * sync_sum / gap_n, 1e-3 * mirror_sum / gap_n, 1e-3 * blocked_sum / gap_n, gap_n);
  }
  if (getenv("VIDO_BA_TIMING") && F.sm)
    fprintf(stderr, "[ba-sm] cluster=%d W=%d P=%d M=%d its=%d trials=%d ns: init=%llu schur=%llu reduce=%llu stage=%llu solve=%llu update+obs=%llu lm=%llu total=%llu | worker 0 own: schur=%llu reduce=%llu load=%llu update=%llu obs=%llu\n",
            ws->cluster_sm, W, P, M, ctl.iterations, ctl.total_trials, tph[5], tph[0], tph[1], tph[6], tph[2], tph[3], tph[4], tph[7], tph[8], tph[9], tph[10], tph[11], tph[12]);
  else if (getenv("VIDO_BA_TIMING"))
    fprintf(stderr, "[ba] cluster=%d W=%d P=%d M=%d its=%d trials=%d ns: linearize=%llu schur=%llu chol=%llu update=%llu init=%llu end=%llu total=%llu\n",
            ws->cluster, W, P, M, ctl.iterations, ctl.total_trials, tph[0] + tph[8] + tph[9] + tph[10] + tph[20], tph[1] + tph[2] + tph[19], tph[3] + tph[11] + tph[12] + tph[13] + tph[16] + tph[17] + tph[18], tph[4] + tph[21], tph[5], tph[6], tph[7]);
  if (!F.sm && getenv("VIDO_BA_TIMING") && atoi(getenv("VIDO_BA_TIMING")) > 1)
    fprintf(stderr, "[ba]   cta0 own time: obs_pass=%llu units=%llu update=%llu\n", tph[20], tph[19], tph[21]);
  if (!F.sm && getenv("VIDO_BA_TIMING") && atoi(getenv("VIDO_BA_TIMING")) > 1)
    fprintf(stderr, "[ba]   obs_wait=%llu blocks=%llu lm_begin=%llu | schur: units=%llu reduce=%llu | chol: diag0=%llu panel=%llu diag=%llu trail_wait=%llu backsub=%llu epilogue=%llu\n",
            tph[9], tph[10], tph[0], tph[1], tph[2], tph[11], tph[16], tph[17], tph[18] + tph[12], tph[13], tph[3]);
  if (st) {
    st->iterations = ctl.iterations;
    st->n_records = F.want_records ? ctl.n_records : 0;
    st->total_trials = ctl.total_trials;
    for (int i = 0; i < st->n_records && i < VIDO_LM_MAX_RECORDS; i++) {
      st->rec[i].chi2 = recs[i].chi2;
      st->rec[i].lambda = recs[i].lambda;
      st->rec[i].trials = recs[i].trials;
    }
  }
  return VIDO_OK;
}

// true when the oldest queued solve has finished (ba_collect will not block)
bool ba_oldest_done(vido_ctx* ctx) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  if (ws->nflight == 0) return false;
  const BaWorkspace::Flight& F = ws->flight[0];
  if (F.sm) return *(volatile int*)(ws->h_out2[F.slot] + ws->out_bytes) == ws->seq2[F.slot];
  return cudaEventQuery(ws->out_done[F.slot]) == cudaSuccess;
}

int ba_partial_host(vido_ctx* ctx, vido_ba_problem* pr, vido_lm_stats* st) {
  int rc = ba_submit(ctx, pr, st != nullptr);
  if (rc) return rc;
  return ba_collect(ctx, pr, st);
}
