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

#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "ba_math.h"
#include "ctx.h"
#include "lm_device.h"

namespace cg = cooperative_groups;
using namespace vb;

#define BA_THREADS 256
#define BA_MAX_W 24
#define BA_MAX_CLUSTER 16
#define BA_PCHUNK 8   // chunks per pose in the pose-block reduction

struct BaArgs {
  int W, P, M;
  int max_iterations;
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
  double* Hpl;    // [18][M] (6x3 block per observation, struct-of-arrays)
  double* Hpp;    // [W][36] diagonal blocks (points + odometry)
  double* Hoff;   // [W-1][36] blocks (i, i+1)
  double* bp;     // [W][6]
  double* ppart;  // [W][BA_PCHUNK][28] partial pose blocks (27 sums)
  double* S;      // [6W][6W] reduced camera system (lower triangle)
  double* bred;   // [6W]
  double* xp;     // [6W]
  double* part;   // [BA_MAX_CLUSTER][4]: chi2, scale, max point diag, max pose diag
  double* cinfo;  // [4]: pose part of the scale, solver failure flag
  double* seJ;    // [W-1][72]
  double* seE;    // [W-1][8]
  LmCtl* ctl_out;
  LmRec* rec;
  unsigned long long* t_phase;  // [8]
  float* out_poses; float* out_rel; float* out_points;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
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

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------------------
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

__device__ __forceinline__ double se3_chi(const BaArgs& a, const Pose* X, int i) {
  double e[6], r0, w;
  edge_se3(X[i], X[i + 1], a.Zinv[i], e, nullptr, nullptr);
  double c = 0;
  for (int k = 0; k < 6; k++) c += e[k] * e[k];
  huber(c * a.info_cam, a.d_cam, r0, w);
  return r0;
}

// robust chi2 of state `st` (thread per observation) -> part[rank][0]
__device__ void phase_errors(const BaArgs& a, int st, int G, int GT, int rank, double* red) {
  const Pose* X = a.X + (size_t)st * a.W;
  const double* pts = a.pts + (size_t)st * 3 * a.P;
  double chi = 0;
  for (int o = G; o < a.M; o += GT) {
    double zc[3], e[3], w;
    chi += obs_chi(a, X[a.obs_pose[o]], pts + 3 * (size_t)a.obs_point[o], o, zc, e, w);
  }
  for (int i = G; i < a.W - 1; i += GT) chi += se3_chi(a, X, i);
  double v[1] = {chi};
  block_reduce<1, false>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 0] = red[0];
}

// linearisation, step 1: thread per observation -> Hpl; thread per odometry edge -> Jacobians
__device__ void phase_lin_obs(const BaArgs& a, int st, int G, int GT) {
  const Pose* X = a.X + (size_t)st * a.W;
  const double* pts = a.pts + (size_t)st * 3 * a.P;
  for (int o = G; o < a.M; o += GT) {
    const Pose& Xp = X[a.obs_pose[o]];
    double zc[3], e[3], w;
    obs_chi(a, Xp, pts + 3 * (size_t)a.obs_point[o], o, zc, e, w);
    w *= a.info_3d;
    // Hpl = w * J_pose^T J_point, J_pose = [-I | Q(zc)], J_point = R^T
    double* hp = a.Hpl + o;
    const size_t M = a.M;
    const double* R = Xp.R;
    const double qx = 2 * zc[0], qy = 2 * zc[1], qz = 2 * zc[2];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const double r0c = R[3 * c], r1c = R[3 * c + 1], r2c = R[3 * c + 2];
      hp[(c)*M] = -w * r0c;
      hp[(3 + c) * M] = -w * r1c;
      hp[(6 + c) * M] = -w * r2c;
      hp[(9 + c) * M] = w * (qz * r1c - qy * r2c);
      hp[(12 + c) * M] = w * (-qz * r0c + qx * r2c);
      hp[(15 + c) * M] = w * (qy * r0c - qx * r1c);
    }
  }
  for (int i = G; i < a.W - 1; i += GT) {
    double e[6], Ji[36], Jj[36], r0, w;
    edge_se3(X[i], X[i + 1], a.Zinv[i], e, Ji, Jj);
    double c = 0;
    for (int k = 0; k < 6; k++) c += e[k] * e[k];
    huber(c * a.info_cam, a.d_cam, r0, w);
    w *= a.info_cam;
    double* J = a.seJ + 72 * (size_t)i;
    for (int k = 0; k < 36; k++) { J[k] = Ji[k]; J[36 + k] = Jj[k]; }
    double* E = a.seE + 8 * (size_t)i;
    for (int k = 0; k < 6; k++) E[k] = e[k];
    E[6] = w;
  }
}

// linearisation, step 2: point blocks (thread per point) and partial pose blocks (warp per (pose, chunk))
struct BaTab {  // shared-memory copies of the small layout tables
  int grp[BA_MAX_W + 1], cnt[BA_MAX_W * (BA_MAX_W + 1)], off[BA_MAX_W * (BA_MAX_W + 1)], base[BA_MAX_W + 1];
};

__device__ void phase_lin_blocks(const BaArgs& a, int st, int G, int GT, int rank, int nranks, double* red, const BaTab& tb) {
  const Pose* X = a.X + (size_t)st * a.W;
  const double* pts = a.pts + (size_t)st * 3 * a.P;
  const int W = a.W;
  double mx = 0;
  for (int l = G; l < a.P; l += GT) {
    double h = 0, b[3] = {0, 0, 0};
    const double* p = pts + 3 * (size_t)l;
    const int f = a.pt_first[l], len = a.pt_len[l], i = l - tb.grp[f];
    for (int k = 0; k < len; k++) {
      const int pp = f + k;
      const int o = tb.base[pp] + tb.off[pp * (W + 1) + f] + i;
      const Pose& Xp = X[pp];
      double zc[3], e[3], w;
      obs_chi(a, Xp, p, o, zc, e, w);
      w *= a.info_3d;
      h += w;  // J_point^T J_point = R R^T = I
      const double* R = Xp.R;
      for (int r = 0; r < 3; r++) b[r] -= w * (R[3 * r] * e[0] + R[3 * r + 1] * e[1] + R[3 * r + 2] * e[2]);
    }
    a.hl[l] = h;
    for (int k = 0; k < 3; k++) a.bl[3 * (size_t)l + k] = b[k];
    mx = fmax(mx, h);
  }
  {
    double v[1] = {mx};
    block_reduce<1, true>(v, red);
    if (threadIdx.x == 0) a.part[rank * 4 + 2] = red[0];
  }
  // pose blocks: job = (pose p, chunk c) over the pose's contiguous observation range; 27 sums per lane, shuffles
  const int lane = threadIdx.x & 31;
  const int gw = rank * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = nranks * (blockDim.x >> 5);
  for (int job = gw; job < W * BA_PCHUNK; job += nw) {
    const int p = job / BA_PCHUNK, ch = job % BA_PCHUNK;
    const Pose Xp = X[p];
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    const int q0 = tb.base[p], q1 = tb.base[p + 1];
    for (int o = q0 + ch * 32 + lane; o < q1; o += 32 * BA_PCHUNK) {
      double zc[3], e[3], w;
      obs_chi(a, Xp, pts + 3 * (size_t)a.obs_point[o], o, zc, e, w);
      w *= a.info_3d;
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
      double* o = a.ppart + ((size_t)p * BA_PCHUNK + ch) * 28;
      for (int k = 0; k < 27; k++) o[k] = acc[k];
    }
  }
}

// linearisation, step 3 (after a barrier): sum the partial pose blocks in fixed order, add the odometry edges
__device__ void phase_lin_poses(const BaArgs& a, int G, int GT, int rank, double* red) {
  double mx = 0;
  for (int p = G; p < a.W; p += GT) {
    double H[36], b[6], s[27];
    for (int k = 0; k < 27; k++) s[k] = 0;
    for (int ch = 0; ch < BA_PCHUNK; ch++) {
      const double* o = a.ppart + ((size_t)p * BA_PCHUNK + ch) * 28;
      for (int k = 0; k < 27; k++) s[k] += o[k];
    }
    int idx = 0;
    for (int r = 0; r < 6; r++) {
      b[r] = s[21 + r];
      for (int c = r; c < 6; c++) { H[6 * r + c] = s[idx]; H[6 * c + r] = s[idx]; idx++; }
    }
    if (p < a.W - 1) {  // edge p: this pose is the "from" vertex
      const double* Ji = a.seJ + 72 * (size_t)p;
      const double* Jj = Ji + 36;
      const double* E = a.seE + 8 * (size_t)p;
      const double w = E[6];
      double* off = a.Hoff + 36 * (size_t)p;
      for (int r = 0; r < 6; r++) {
        double t = 0;
        for (int k = 0; k < 6; k++) t += Ji[6 * k + r] * E[k];
        b[r] -= w * t;
        for (int c = 0; c < 6; c++) {
          double h = 0, ho = 0;
          for (int k = 0; k < 6; k++) { h += Ji[6 * k + r] * Ji[6 * k + c]; ho += Ji[6 * k + r] * Jj[6 * k + c]; }
          H[6 * r + c] += w * h;
          off[6 * r + c] = w * ho;
        }
      }
    }
    if (p > 0) {  // edge p-1: "to" vertex
      const double* Jj = a.seJ + 72 * (size_t)(p - 1) + 36;
      const double* E = a.seE + 8 * (size_t)(p - 1);
      const double w = E[6];
      for (int r = 0; r < 6; r++) {
        double t = 0;
        for (int k = 0; k < 6; k++) t += Jj[6 * k + r] * E[k];
        b[r] -= w * t;
        for (int c = 0; c < 6; c++) {
          double h = 0;
          for (int k = 0; k < 6; k++) h += Jj[6 * k + r] * Jj[6 * k + c];
          H[6 * r + c] += w * h;
        }
      }
    }
    for (int k = 0; k < 36; k++) a.Hpp[36 * (size_t)p + k] = H[k];
    for (int r = 0; r < 6; r++) { a.bp[6 * p + r] = b[r]; mx = fmax(mx, fabs(H[7 * r])); }
  }
  double v[1] = {mx};
  block_reduce<1, true>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 3] = red[0];
}

// reduced camera system: S(p1,p2) = Hpp(p1,p2) + lambda I - sum_l Hpl(p1,l) Hpl(p2,l)^T / (hl + lambda).
// Job = pose pair, dealt round-robin to the CTAs (ordered by distance so that every CTA gets the same mix of big and
// small pairs); all threads of the CTA share the pair's common points (coalesced struct-of-arrays loads), 36-value
// warp-shuffle reduction, then a fixed-order sum over the warps in shared memory (deterministic, no atomics).
__device__ void phase_schur(const BaArgs& a, double lambda, int rank, int nranks, const BaTab& tb, double* sred /* [warps][36] */) {
  const int W = a.W, n = 6 * W;
  const size_t M = a.M;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int npairs = W * (W + 1) / 2;
  for (int job = rank; job < npairs + W; job += nranks) {
    if (job < npairs) {
      int d = 0, rem = job;
      while (rem >= W - d) { rem -= W - d; d++; }
      const int p1 = rem, p2 = rem + d;
      double acc[36];
#pragma unroll
      for (int k = 0; k < 36; k++) acc[k] = 0;
      for (int f = 0; f <= p1; f++) {
        const int cntf = tb.cnt[f * (W + 1) + (p2 - f)];
        const int q1 = tb.base[p1] + tb.off[p1 * (W + 1) + f], q2 = tb.base[p2] + tb.off[p2 * (W + 1) + f], l0 = tb.grp[f];
        for (int i = tid; i < cntf; i += blockDim.x) {
          const double s = 1.0 / (a.hl[l0 + i] + lambda);
          const double* h1 = a.Hpl + (q1 + i);
          const double* h2 = a.Hpl + (q2 + i);
          double t[18], g[18];
#pragma unroll
          for (int k = 0; k < 18; k++) { t[k] = h1[k * M] * s; g[k] = h2[k * M]; }
#pragma unroll
          for (int r = 0; r < 6; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) acc[6 * r + c] += t[3 * r] * g[3 * c] + t[3 * r + 1] * g[3 * c + 1] + t[3 * r + 2] * g[3 * c + 2];
        }
      }
#pragma unroll
      for (int k = 0; k < 36; k++) acc[k] = warp_sum(acc[k]);
      __syncthreads();  // sred free (previous job consumed)
      if (lane == 0)
        for (int k = 0; k < 36; k++) sred[warp * 36 + k] = acc[k];
      __syncthreads();
      if (tid < 36) {
        double t = 0;
        for (int w = 0; w < nwarp; w++) t += sred[w * 36 + tid];
        const int r = tid / 6, c = tid - 6 * r;
        double h = 0;
        if (p2 == p1) h = a.Hpp[36 * (size_t)p1 + tid] + ((r == c) ? lambda : 0.0);
        else if (p2 == p1 + 1) h = a.Hoff[36 * (size_t)p1 + tid];
        a.S[(size_t)(6 * p2 + c) * n + 6 * p1 + r] = h - t;  // block (p1,p2), p1 <= p2, transposed into the lower triangle
      }
    } else {
      const int p = job - npairs;
      double acc[6] = {0, 0, 0, 0, 0, 0};
      for (int o = tb.base[p] + tid; o < tb.base[p + 1]; o += blockDim.x) {
        const int l = a.obs_point[o];
        const double s = 1.0 / (a.hl[l] + lambda);
        const double* b = a.bl + 3 * (size_t)l;
        const double c0 = s * b[0], c1 = s * b[1], c2 = s * b[2];
        const double* h = a.Hpl + o;
#pragma unroll
        for (int r = 0; r < 6; r++) acc[r] += h[(3 * r) * M] * c0 + h[(3 * r + 1) * M] * c1 + h[(3 * r + 2) * M] * c2;
      }
#pragma unroll
      for (int r = 0; r < 6; r++) acc[r] = warp_sum(acc[r]);
      __syncthreads();
      if (lane == 0)
        for (int r = 0; r < 6; r++) sred[warp * 36 + r] = acc[r];
      __syncthreads();
      if (tid < 6) {
        double t = 0;
        for (int w = 0; w < nwarp; w++) t += sred[w * 36 + tid];
        a.bred[6 * p + tid] = a.bp[6 * p + tid] - t;
      }
    }
  }
}

// blocked (6x6) LL^T of the reduced system in shared memory + block substitutions; one CTA.  ld is odd to avoid
// bank conflicts on column accesses.  The inverse of every diagonal block is kept (Li), so panels and substitutions
// are plain products.  Also applies the pose increments (trial poses) and their part of the scale.
__device__ void phase_chol(const BaArgs& a, double lambda, int cur, double* Ls, double* ys, double* Li, double* red) {
  const int n = 6 * a.W, ld = n + 1, tid = threadIdx.x, nt = blockDim.x, nb = a.W;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int i = tid; i < n * n; i += nt) {
    const int r = i / n, c = i - r * n;
    if (c <= r) Ls[r * ld + c] = a.S[i];
  }
  for (int i = tid; i < n; i += nt) ys[i] = a.bred[i];
  __syncthreads();
  for (int jb = 0; jb < nb; jb++) {
    const int j0 = 6 * jb;
    // (1) diagonal block: L11 and its inverse, fully unrolled so that everything stays in registers (thread 0)
    if (tid == 0) {
      double L[36], Iv[36];
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) { L[6 * r + c] = (c <= r) ? Ls[(j0 + r) * ld + j0 + c] : 0.0; Iv[6 * r + c] = 0.0; }
      bool ok = true;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        double d = L[7 * j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k];
        ok = ok && (d > 0);
        const double ljj = sqrt(d), inv = 1.0 / ljj;
        L[7 * j] = ljj;
        Iv[7 * j] = inv;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
          double t = L[6 * i + j];
#pragma unroll
          for (int k = 0; k < j; k++) t -= L[6 * i + k] * L[6 * j + k];
          L[6 * i + j] = t * inv;
        }
      }
      if (!ok) s_bad = 1;
      else {
#pragma unroll
        for (int c = 0; c < 6; c++)  // Iv = L^-1 (lower), column by column
#pragma unroll
          for (int r = c + 1; r < 6; r++) {
            double t = 0;
#pragma unroll
            for (int k = c; k < r; k++) t += L[6 * r + k] * Iv[6 * k + c];
            Iv[6 * r + c] = -t * Iv[7 * r];
          }
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int c = 0; c <= r; c++) Ls[(j0 + r) * ld + j0 + c] = L[6 * r + c];
#pragma unroll
        for (int k = 0; k < 36; k++) Li[36 * jb + k] = Iv[k];
      }
    }
    __syncthreads();
    if (s_bad) break;
    // (2) panel: rows below the block, L21 = A21 * L11^-T  =>  row'[j] = sum_{k<=j} row[k] * Iv[j][k]
    for (int i = j0 + 6 + tid; i < n; i += nt) {
      double row[6], o[6];
#pragma unroll
      for (int k = 0; k < 6; k++) row[k] = Ls[i * ld + j0 + k];
      const double* Iv = Li + 36 * jb;
#pragma unroll
      for (int j = 0; j < 6; j++) {
        double t = 0;
#pragma unroll
        for (int k = 0; k <= j; k++) t += row[k] * Iv[6 * j + k];
        o[j] = t;
      }
#pragma unroll
      for (int j = 0; j < 6; j++) Ls[i * ld + j0 + j] = o[j];
    }
    __syncthreads();
    // (3) trailing update A22 -= L21 L21^T (lower triangle): warp per row, lanes over the columns <= row
    const int m = n - j0 - 6;
    for (int r = warp; r < m; r += nwarp) {
      const double* lr = Ls + (j0 + 6 + r) * ld + j0;
      const double r0 = lr[0], r1 = lr[1], r2 = lr[2], r3 = lr[3], r4 = lr[4], r5 = lr[5];
      for (int c = lane; c <= r; c += 32) {
        const double* lc = Ls + (j0 + 6 + c) * ld + j0;
        Ls[(j0 + 6 + r) * ld + j0 + 6 + c] -= r0 * lc[0] + r1 * lc[1] + r2 * lc[2] + r3 * lc[3] + r4 * lc[4] + r5 * lc[5];
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int failed = s_bad;
  if (!failed) {
    // forward substitution: y_b = Iv_b * rhs_b, then rhs_below -= L21 * y_b
    for (int jb = 0; jb < nb; jb++) {
      const int j0 = 6 * jb;
      double yb = 0;
      if (tid < 6) {
        const double* Iv = Li + 36 * jb;
        for (int k = 0; k <= tid; k++) yb += Iv[6 * tid + k] * ys[j0 + k];
      }
      __syncthreads();
      if (tid < 6) ys[j0 + tid] = yb;
      __syncthreads();
      for (int i = j0 + 6 + tid; i < n; i += nt) {
        const double* li = Ls + i * ld + j0;
        ys[i] -= li[0] * ys[j0] + li[1] * ys[j0 + 1] + li[2] * ys[j0 + 2] + li[3] * ys[j0 + 3] + li[4] * ys[j0 + 4] + li[5] * ys[j0 + 5];
      }
      __syncthreads();
    }
    // backward substitution with L^T: x_b = Iv_b^T * rhs_b, then rhs_above -= L(b, above)^T x_b
    for (int jb = nb - 1; jb >= 0; jb--) {
      const int j0 = 6 * jb;
      double xb = 0;
      if (tid < 6) {
        const double* Iv = Li + 36 * jb;
        for (int k = tid; k < 6; k++) xb += Iv[6 * k + tid] * ys[j0 + k];
      }
      __syncthreads();
      if (tid < 6) ys[j0 + tid] = xb;
      __syncthreads();
      for (int i = tid; i < j0; i += nt) {
        double t = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) t += Ls[(j0 + k) * ld + i] * ys[j0 + k];
        ys[i] -= t;
      }
      __syncthreads();
    }
  }
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

// back-substitution of the points (thread per point), trial points, scale, robust chi2 of the trial state
__device__ void phase_update(const BaArgs& a, double lambda, int cur, int failed, int G, int GT, int rank, double* red,
                             const BaTab& tb) {
  const int trial = cur ^ 1, W = a.W;
  const size_t M = a.M;
  const Pose* Xt = a.X + (size_t)trial * a.W;
  const double* pts = a.pts + (size_t)cur * 3 * a.P;
  double* ptt = a.pts + (size_t)trial * 3 * a.P;
  double scale = 0, chi = 0;
  for (int l = G; l < a.P; l += GT) {
    const double* b = a.bl + 3 * (size_t)l;
    const int f = a.pt_first[l], len = a.pt_len[l], i = l - tb.grp[f];
    double x[3];
    if (failed) { x[0] = b[0]; x[1] = b[1]; x[2] = b[2]; }
    else {
      double c0 = b[0], c1 = b[1], c2 = b[2];
      for (int k = 0; k < len; k++) {
        const int pp = f + k;
        const double* h = a.Hpl + (tb.base[pp] + tb.off[pp * (W + 1) + f] + i);
        const double* xp = a.xp + 6 * pp;
#pragma unroll
        for (int r = 0; r < 6; r++) { c0 -= h[(3 * r) * M] * xp[r]; c1 -= h[(3 * r + 1) * M] * xp[r]; c2 -= h[(3 * r + 2) * M] * xp[r]; }
      }
      const double s = 1.0 / (a.hl[l] + lambda);
      x[0] = s * c0; x[1] = s * c1; x[2] = s * c2;
    }
    double pn[3];
    for (int k = 0; k < 3; k++) {
      pn[k] = pts[3 * (size_t)l + k] + x[k];
      ptt[3 * (size_t)l + k] = pn[k];
      scale += x[k] * (lambda * x[k] + b[k]);
    }
    for (int k = 0; k < len; k++) {
      const int pp = f + k;
      double zc[3], e[3], w;
      chi += obs_chi(a, Xt[pp], pn, tb.base[pp] + tb.off[pp * (W + 1) + f] + i, zc, e, w);
    }
  }
  for (int i = G; i < a.W - 1; i += GT) chi += se3_chi(a, Xt, i);
  double v[2] = {chi, scale};
  block_reduce<2, false>(v, red);
  if (threadIdx.x == 0) { a.part[rank * 4 + 0] = red[0]; a.part[rank * 4 + 1] = red[1]; }
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
__global__ void __launch_bounds__(BA_THREADS, 1) ba_window_kernel(BaArgs a) {
  extern __shared__ __align__(16) double dsm[];  // (6W)(6W+1) + 6W + 36W doubles for the dense factorisation (CTA 0)
  __shared__ double red[16 * 2 + 32];
  __shared__ double sred[(BA_THREADS / 32) * 36];
  __shared__ LmCtl ctl;  // every CTA keeps an identical copy: decisions are recomputed from the same partial sums
  __shared__ BaTab tb;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
  const int G = rank * blockDim.x + threadIdx.x, GT = nranks * blockDim.x;
  const int tid = threadIdx.x;
  unsigned long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long t0 = (G == 0) ? gtime() : 0;
  const unsigned long long t_start = t0;
#define TOC(slot) do { if (G == 0) { unsigned long long t1_ = gtime(); tph[slot] += t1_ - t0; t0 = t1_; } } while (0)

  for (int i = tid; i <= a.W; i += blockDim.x) { tb.grp[i] = a.grp_start[i]; tb.base[i] = a.pose_base[i]; }
  for (int i = tid; i < a.W * (a.W + 1); i += blockDim.x) { tb.cnt[i] = a.cnt_gt[i]; tb.off[i] = a.off[i]; }
  if (tid == 0) lm_reset(&ctl);
  phase_init(a, G, GT);
  cluster.sync();
  if (a.W + a.P == 0 || a.max_iterations <= 0) {
    if (G == 0) { ctl.iterations = (a.W + a.P == 0) ? -1 : 0; *a.ctl_out = ctl; }
    phase_output(a, 0, G, GT);
    cluster.sync();
    phase_output_rel(a, G, GT);
    return;
  }
  phase_errors(a, 0, G, GT, rank, red);
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
    const int cur = ctl.cur;
    phase_lin_obs(a, cur, G, GT);
    cluster.sync();
    phase_lin_blocks(a, cur, G, GT, rank, nranks, red, tb);
    cluster.sync();
    phase_lin_poses(a, G, GT, rank, red);
    cluster.sync();
    if (tid == 0) {
      double m = 0;
      for (int r = 0; r < nranks; r++) m = fmax(m, fmax(a.part[r * 4 + 2], a.part[r * 4 + 3]));
      lm_begin_iteration(&ctl, it, m, -1.0);
    }
    __syncthreads();
    TOC(0);
    while (true) {
      const double lambda = ctl.lambda;
      phase_schur(a, lambda, rank, nranks, tb, sred);
      cluster.sync();
      TOC(2);
      if (rank == 0) {
        double* ysm = dsm + (size_t)(6 * a.W) * (6 * a.W + 1);
        phase_chol(a, lambda, cur, dsm, ysm, ysm + 6 * a.W, red);
      }
      cluster.sync();
      TOC(3);
      const int failed = a.cinfo[1] != 0.0;
      phase_update(a, lambda, cur, failed, G, GT, rank, red, tb);
      cluster.sync();
      TOC(4);
      if (tid == 0) {
        double chi = 0, scale = a.cinfo[0];
        for (int r = 0; r < nranks; r++) { chi += a.part[r * 4]; scale += a.part[r * 4 + 1]; }
        lm_trial(&ctl, chi, scale, failed);
      }
      __syncthreads();
      // the partial sums are next overwritten in phase_update, two cluster barriers later: no race with slower CTAs
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
    for (int k = 0; k < 8; k++) a.t_phase[k] = tph[k];
  }
#undef TOC
}

// =========================================================================================================
// host side
// =========================================================================================================
struct BaWorkspace {
  int capW = 0, capP = 0, capM = 0;
  int cluster = 8;
  char* d_base = nullptr;            // solver workspace
  char* d_in = nullptr;              // input block
  char* d_out = nullptr;             // output block
  BaArgs args;
  char* h_in = nullptr;              // pinned mirrors
  char* h_out = nullptr;
  size_t in_bytes = 0, out_bytes = 0;
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
  a.ctl_out = carve<LmCtl>(p, 1); a.t_phase = carve<unsigned long long>(p, 8);
  a.out_poses = carve<float>(p, 16 * W); a.out_rel = carve<float>(p, 16 * W); a.out_points = carve<float>(p, 3 * (size_t)P);
  a.rec = carve<LmRec>(p, VIDO_LM_REC);
}

static void carve_all(char*& p, BaArgs& a, int capW, int capP, int capM) {
  a.X = carve<Pose>(p, 2 * capW); a.Zinv = carve<Pose>(p, capW); a.pts = carve<double>(p, 6 * (size_t)capP);
  a.hl = carve<double>(p, capP); a.bl = carve<double>(p, 3 * (size_t)capP); a.Hpl = carve<double>(p, 18 * (size_t)capM);
  a.Hpp = carve<double>(p, 36 * capW); a.Hoff = carve<double>(p, 36 * capW); a.bp = carve<double>(p, 6 * capW);
  a.ppart = carve<double>(p, (size_t)capW * BA_PCHUNK * 28);
  a.S = carve<double>(p, 36 * (size_t)capW * capW); a.bred = carve<double>(p, 6 * capW); a.xp = carve<double>(p, 6 * capW);
  a.part = carve<double>(p, 4 * BA_MAX_CLUSTER); a.cinfo = carve<double>(p, 4);
  a.seJ = carve<double>(p, 72 * capW); a.seE = carve<double>(p, 8 * capW);
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
  VIDO_CUDA(cudaMalloc(&ws->d_in, ws->in_bytes));
  VIDO_CUDA(cudaMalloc(&ws->d_out, ws->out_bytes));
  VIDO_CUDA(cudaMallocHost(&ws->h_in, ws->in_bytes));
  VIDO_CUDA(cudaMallocHost(&ws->h_out, ws->out_bytes));
  memset(&ws->args, 0, sizeof ws->args);
  p = ws->d_base;
  carve_all(p, ws->args, capW, capP, capM);
  const size_t smem = sizeof(double) * ((size_t)(6 * capW) * (6 * capW + 1) + 6 * capW + 36 * capW);
  if (smem > 200 * 1024) { ctx->err = "BA window too large for the shared-memory Cholesky"; return VIDO_ERR_ARG; }
  VIDO_CUDA(cudaFuncSetAttribute(ba_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // 16-CTA clusters are a non-portable size: opt in, and verify that one fits
  ws->cluster = 8;
  if (cudaFuncSetAttribute(ba_window_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(BA_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, ba_window_kernel, &cfg) == cudaSuccess && nclusters >= 1) ws->cluster = 16;
  }
  cudaGetLastError();
  if (getenv("VIDO_BA_CLUSTER") && atoi(getenv("VIDO_BA_CLUSTER")) == 8) ws->cluster = 8;
  return VIDO_OK;
}

void ba_teardown(vido_ctx* ctx) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  if (!ws) return;
  cudaFree(ws->d_base); cudaFree(ws->d_in); cudaFree(ws->d_out);
  cudaFreeHost(ws->h_in); cudaFreeHost(ws->h_out);
  delete ws;
  ctx->ba = nullptr;
}

int ba_partial_host(vido_ctx* ctx, vido_ba_problem* pr, vido_lm_stats* st) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  const int W = pr->n_poses, P = pr->n_points, M = pr->n_obs;
  if (W < 0 || P < 0 || M < 0) return VIDO_ERR_ARG;
  if (W > ws->capW || P > ws->capP || M > ws->capM) { ctx->err = "BA problem exceeds the context capacity"; return VIDO_ERR_CAPACITY; }
  cudaStream_t s = ctx->stream;
  BaArgs a = ws->args;
  a.W = W; a.P = P; a.M = M;
  a.max_iterations = pr->max_iterations;
  a.info_cam = 1.0 / (double)pr->sigma2_cam;
  a.info_3d = 1.0 / (double)pr->sigma2_3d;
  a.d_cam = (double)pr->huber_cam;
  a.d_3d = (double)pr->huber_3d;
  a.gain_threshold = (double)pr->gain_threshold;
  // ---- host-side layout: tracks sorted by (first pose, length descending); observations pose-major (see the header)
  BaArgs h;  // host view of the input block (same carving as the device view)
  {
    char* hp = ws->h_in; carve_inputs(hp, h, W, P, M);
    char* dp = ws->d_in; carve_inputs(dp, a, W, P, M);
    char* dq = ws->d_out; carve_outputs(dq, a, W, P);
  }
  float* h_poses = (float*)h.poses_f32; float* h_rel = (float*)h.rel_f32;
  int* h_obs_pose = (int*)h.obs_pose; int* h_obs_point = (int*)h.obs_point; float* h_xyz = (float*)h.obs_xyz;
  int* h_pt_len = (int*)h.pt_len; int* h_pt_first = (int*)h.pt_first; float* h_pts = (float*)h.points_f32;
  int* h_grp = (int*)h.grp_start; int* h_cnt = (int*)h.cnt_gt; int* h_off = (int*)h.off; int* h_base = (int*)h.pose_base;
  memcpy(h_poses, pr->poses, sizeof(float) * 16 * W);
  if (W > 1) memcpy(h_rel, pr->rel_motion, sizeof(float) * 16 * (W - 1));
  std::vector<int> first(P, 1 << 30), len(P, 0), last(P, -1);
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
  std::vector<int> keycnt((size_t)W * (W + 1) + 2, 0), newid(P), oldid(P);
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
    h_pts[3 * n] = pr->points[3 * l]; h_pts[3 * n + 1] = pr->points[3 * l + 1]; h_pts[3 * n + 2] = pr->points[3 * l + 2];
  }
  for (int o = 0; o < M; o++) {
    const int l = pr->obs_point[o], p = pr->obs_pose[o], n = newid[l], f = first[l];
    const int q = h_base[p] + h_off[p * (W + 1) + f] + (n - h_grp[f]);
    h_obs_pose[q] = p;
    h_obs_point[q] = n;
    h_xyz[q] = pr->obs_xyz[3 * o]; h_xyz[(size_t)M + q] = pr->obs_xyz[3 * o + 1]; h_xyz[2 * (size_t)M + q] = pr->obs_xyz[3 * o + 2];
  }
  {
    char* hp = ws->h_in; BaArgs t2; carve_inputs(hp, t2, W, P, M);
    VIDO_CUDA(cudaMemcpyAsync(ws->d_in, ws->h_in, (size_t)(hp - ws->h_in), cudaMemcpyHostToDevice, s));
  }
  const size_t smem = sizeof(double) * ((size_t)(6 * W) * (6 * W + 1) + 6 * W + 36 * W);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ws->cluster); cfg.blockDim = dim3(BA_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ws->cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEventRecord(ctx->ev0, s);
    VIDO_CUDA(cudaLaunchKernelEx(&cfg, ba_window_kernel, a));
    cudaEventRecord(ctx->ev1, s);
    ctx->launches++;
  }
  BaArgs ho;
  size_t out_used;
  {
    char* hq = ws->h_out; carve_outputs(hq, ho, W, P);
    // the LM records sit at the end of the block: copy them only when asked for
    out_used = st ? (size_t)(hq - ws->h_out) : (size_t)((char*)ho.rec - ws->h_out);
  }
  VIDO_CUDA(cudaMemcpyAsync(ws->h_out, ws->d_out, out_used, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
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
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) { ctx->t_ms[3] += ms; ctx->t_n[3]++; }
    const double edges = (double)M + (double)std::max(W - 1, 0);
    ctx->ba_alg_bytes += edges * (296.0 * std::max(ctl.iterations, 0) + 152.0 * (ctl.total_trials + 1));
  }
  if (getenv("VIDO_BA_TIMING"))
    fprintf(stderr, "[ba] cluster=%d W=%d P=%d M=%d its=%d trials=%d ns: linearize=%llu schur=%llu chol=%llu update=%llu init=%llu end=%llu total=%llu\n",
            ws->cluster, W, P, M, ctl.iterations, ctl.total_trials, tph[0], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7]);
  if (st) {
    st->iterations = ctl.iterations;
    st->n_records = ctl.n_records;
    st->total_trials = ctl.total_trials;
    for (int i = 0; i < ctl.n_records && i < VIDO_LM_MAX_RECORDS; i++) {
      st->rec[i].chi2 = recs[i].chi2;
      st->rec[i].lambda = recs[i].lambda;
      st->rec[i].trials = recs[i].trials;
    }
  }
  return VIDO_OK;
}
