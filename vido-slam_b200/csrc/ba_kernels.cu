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
//     full (poses+points) H; here the point blocks are eliminated first (3x3 pivots), the reduced camera system
//     (6W x 6W) is Cholesky-factored in shared memory.  Same linear system, same positive-definiteness test.
// Everything is FP64 (g2o is built in double); inputs/outputs are the float32 Map fields.
#include <cooperative_groups.h>

#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "ba_math.h"
#include "ctx.h"

namespace cg = cooperative_groups;
using namespace vb;

#define BA_THREADS 512
#define BA_CLUSTER 8
#define BA_MAX_W 24
#define BA_MAX_REC 320

struct BaCtl {  // LM state, written by one thread, read by all after a cluster barrier
  double lambda, ni, currentChi, iniChi, tempChi, lastTrialChi, chi2_check, lastChi, rho;
  int it, qmax, nBad, stop_flag, ok, accepted, fail, phase_done;
  int cur;  // index of the accepted state buffer
  int iterations, n_records, total_trials;
  unsigned long long t_phase[8];  // ns spent in: 0 linearize, 1 prepare, 2 schur, 3 chol, 4 update, 5 errors, 6 decide, 7 total
};

struct BaRec {
  double chi2, lambda;
  int trials, pad;
};

struct BaArgs {
  int W, P, M;
  int max_iterations;
  double info_cam, info_3d, d_cam, d_3d, gain_threshold;
  // graph (device)
  const float* poses_f32;   // [W][16]
  const float* rel_f32;     // [W-1][16]
  const float* points_f32;  // [P][3]
  const int* obs_pose;      // [M] sorted by (point, pose)
  const float* obs_xyz;     // [M][3]
  const int* pt_start;      // [P+1]
  const int* pose_start;    // [W+1]
  const int* pose_obs;      // [M] observation ids grouped by pose
  const int* obs_point;     // [M]
  // state
  Pose* X;        // [2][W]
  Pose* Zinv;     // [W-1]
  double* pts;    // [2][P][3]
  // system
  double* Hll;    // [P][6]  (xx,xy,xz,yy,yz,zz)
  double* bl;     // [P][3]
  double* Dinv;   // [P][6]
  double* cl;     // [P][3]
  double* Hpl;    // [M][18] (6x3)
  double* Tpl;    // [M][18] Hpl * Dinv
  double* Hpp;    // [W][36] diagonal blocks (point + odometry edges)
  double* Hoff;   // [W-1][36] blocks (i, i+1)
  double* bp;     // [W][6]
  double* S;      // [6W][6W] reduced camera system (lower triangle used)
  double* bred;   // [6W]
  double* xp;     // [6W]
  double* xl;     // [P][3]
  double* part;   // [BA_CLUSTER][4] per-CTA partial sums (chi2, scale, maxdiag)
  double* seJ;    // [W-1][2][36] odometry Jacobians, [W-1] weights appended
  double* seE;    // [W-1][8]  error(6), weight, chi
  BaCtl* ctl;
  BaRec* rec;     // [BA_MAX_REC]
  // outputs
  float* out_poses;   // [W][16]
  float* out_rel;     // [W-1][16]
  float* out_points;  // [P][3]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV values per thread; result valid in thread 0 .. NV-1 of warp 0 (returned in out[] of thread 0)
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* smem /* [warps][NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0)
    for (int k = 0; k < NV; k++) smem[warp * NV + k] = v[k];
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
    for (int w = 0; w < nw; w++) s += smem[w * NV + threadIdx.x];
    smem[threadIdx.x] = s;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// phases.  G / GT: cluster-wide thread index / count; rank: CTA rank in the cluster
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
    pose_inv_mul(Z, I, Zi);  // Z^-1
    a.Zinv[i] = Zi;
  }
  for (int i = G; i < 3 * a.P; i += GT) {
    double v = (double)a.points_f32[i];
    a.pts[i] = v;
    a.pts[3 * (size_t)a.P + i] = v;
  }
}

// robust chi2 of state `st` -> per-CTA partial in part[rank][0]
__device__ void phase_errors(const BaArgs& a, int st, int G, int GT, int rank, double* red) {
  const Pose* X = a.X + (size_t)st * a.W;
  const double* pts = a.pts + (size_t)st * 3 * a.P;
  double chi = 0;
  for (int o = G; o < a.M; o += GT) {
    const int p = a.obs_pose[o], l = a.obs_point[o];
    const double z[3] = {(double)a.obs_xyz[3 * o], (double)a.obs_xyz[3 * o + 1], (double)a.obs_xyz[3 * o + 2]};
    double zc[3], e[3], r0, w;
    edge_xyz(X[p], pts + 3 * (size_t)l, z, zc, e);
    huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * a.info_3d, a.d_3d, r0, w);
    chi += r0;
  }
  for (int i = G; i < a.W - 1; i += GT) {
    double e[6], r0, w;
    edge_se3(X[i], X[i + 1], a.Zinv[i], e, nullptr, nullptr);
    double c = 0;
    for (int k = 0; k < 6; k++) c += e[k] * e[k];
    huber(c * a.info_cam, a.d_cam, r0, w);
    chi += r0;
  }
  double v[1] = {chi};
  block_sum<1>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 0] = red[0];
}

// linearisation at the accepted state: point blocks, pose-point blocks, pose blocks, odometry edges
__device__ void phase_linearize(const BaArgs& a, int st, int G, int GT, int rank, int nranks, double* red) {
  const Pose* X = a.X + (size_t)st * a.W;
  const double* pts = a.pts + (size_t)st * 3 * a.P;
  // (a) thread per point: Hll, bl, Hpl
  double mx = 0;
  for (int l = G; l < a.P; l += GT) {
    double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    const double* p = pts + 3 * (size_t)l;
    for (int o = a.pt_start[l]; o < a.pt_start[l + 1]; o++) {
      const Pose& Xp = X[a.obs_pose[o]];
      const double z[3] = {(double)a.obs_xyz[3 * o], (double)a.obs_xyz[3 * o + 1], (double)a.obs_xyz[3 * o + 2]};
      double zc[3], e[3], r0, w;
      edge_xyz(Xp, p, z, zc, e);
      huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * a.info_3d, a.d_3d, r0, w);
      w *= a.info_3d;
      const double* R = Xp.R;  // J_point = R^T  =>  J^T J = R R^T, J^T e = R e
      H[0] += w * (R[0] * R[0] + R[1] * R[1] + R[2] * R[2]);
      H[1] += w * (R[0] * R[3] + R[1] * R[4] + R[2] * R[5]);
      H[2] += w * (R[0] * R[6] + R[1] * R[7] + R[2] * R[8]);
      H[3] += w * (R[3] * R[3] + R[4] * R[4] + R[5] * R[5]);
      H[4] += w * (R[3] * R[6] + R[4] * R[7] + R[5] * R[8]);
      H[5] += w * (R[6] * R[6] + R[7] * R[7] + R[8] * R[8]);
      for (int i = 0; i < 3; i++) b[i] -= w * (R[3 * i] * e[0] + R[3 * i + 1] * e[1] + R[3 * i + 2] * e[2]);
      // Hpl = w * J_pose^T J_point, J_pose = [-I | Q], Q = 2[[0,-z,y],[z,0,-x],[-y,x,0]] (zc), J_point = R^T
      double* hp = a.Hpl + 18 * (size_t)o;
      const double qx = 2 * zc[0], qy = 2 * zc[1], qz = 2 * zc[2];
      for (int c = 0; c < 3; c++) {
        // column c of R^T = row c of R
        const double r0c = R[3 * c], r1c = R[3 * c + 1], r2c = R[3 * c + 2];  // (R^T)(0..2, c) = R(c, 0..2)
        hp[0 * 3 + c] = -w * r0c;
        hp[1 * 3 + c] = -w * r1c;
        hp[2 * 3 + c] = -w * r2c;
        // Q^T rows: Q^T = 2[[0,z,-y],[-z,0,x],[y,-x,0]]
        hp[3 * 3 + c] = w * (qz * r1c - qy * r2c);
        hp[4 * 3 + c] = w * (-qz * r0c + qx * r2c);
        hp[5 * 3 + c] = w * (qy * r0c - qx * r1c);
      }
    }
    for (int k = 0; k < 6; k++) a.Hll[6 * (size_t)l + k] = H[k];
    for (int k = 0; k < 3; k++) a.bl[3 * (size_t)l + k] = b[k];
    mx = fmax(mx, fmax(fabs(H[0]), fmax(fabs(H[3]), fabs(H[5]))));
  }
  // (b) odometry edges: thread per edge
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
  {  // per-CTA max of the point diagonals
    double v[1] = {mx};
    // max-reduce through the sum helper is wrong; do a dedicated max
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[0] = fmax(v[0], __shfl_xor_sync(0xffffffffu, v[0], o));
    __syncthreads();
    if (lane == 0) red[warp] = v[0];
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = 0;
      for (int w2 = 0; w2 < nw; w2++) m = fmax(m, red[w2]);
      a.part[rank * 4 + 2] = m;
    }
    __syncthreads();
  }
  cg::this_cluster().sync();  // seJ/seE visible
  // (c) pose blocks: CTA per pose (round robin), all threads over the pose's observations
  for (int p = rank; p < a.W; p += nranks) {
    double acc[27];
#pragma unroll
    for (int k = 0; k < 27; k++) acc[k] = 0;
    const Pose& Xp = X[p];
    for (int k = a.pose_start[p] + threadIdx.x; k < a.pose_start[p + 1]; k += blockDim.x) {
      const int o = a.pose_obs[k], l = a.obs_point[o];
      const double z[3] = {(double)a.obs_xyz[3 * o], (double)a.obs_xyz[3 * o + 1], (double)a.obs_xyz[3 * o + 2]};
      double zc[3], e[3], r0, w;
      edge_xyz(Xp, pts + 3 * (size_t)l, z, zc, e);
      huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * a.info_3d, a.d_3d, r0, w);
      w *= a.info_3d;
      const double qx = 2 * zc[0], qy = 2 * zc[1], qz = 2 * zc[2];
      // J = [-I | Q]; rows of Q: (0,-qz,qy), (qz,0,-qx), (-qy,qx,0)
      const double J[3][6] = {{-1, 0, 0, 0, -qz, qy}, {0, -1, 0, qz, 0, -qx}, {0, 0, -1, -qy, qx, 0}};
      int idx = 0;
#pragma unroll
      for (int r = 0; r < 6; r++) {
        acc[21 + r] -= w * (J[0][r] * e[0] + J[1][r] * e[1] + J[2][r] * e[2]);
#pragma unroll
        for (int c = r; c < 6; c++) acc[idx++] += w * (J[0][r] * J[0][c] + J[1][r] * J[1][c] + J[2][r] * J[2][c]);
      }
    }
    block_sum<27>(acc, red);
    if (threadIdx.x == 0) {
      double H[36], b[6];
      int idx = 0;
      for (int r = 0; r < 6; r++) {
        b[r] = red[21 + r];
        for (int c = r; c < 6; c++) { H[6 * r + c] = red[idx]; H[6 * c + r] = red[idx]; idx++; }
      }
      // odometry edges touching this pose: edge p (as "from", Ji) and edge p-1 (as "to", Jj)
      if (p < a.W - 1) {
        const double* Ji = a.seJ + 72 * (size_t)p;
        const double* Jj = Ji + 36;
        const double* E = a.seE + 8 * (size_t)p;
        const double w = E[6];
        double* off = a.Hoff + 36 * (size_t)p;
        for (int r = 0; r < 6; r++) {
          double s = 0;
          for (int k = 0; k < 6; k++) s += Ji[6 * k + r] * E[k];
          b[r] -= w * s;
          for (int c = 0; c < 6; c++) {
            double h = 0, ho = 0;
            for (int k = 0; k < 6; k++) { h += Ji[6 * k + r] * Ji[6 * k + c]; ho += Ji[6 * k + r] * Jj[6 * k + c]; }
            H[6 * r + c] += w * h;
            off[6 * r + c] = w * ho;
          }
        }
      }
      if (p > 0) {
        const double* Jj = a.seJ + 72 * (size_t)(p - 1) + 36;
        const double* E = a.seE + 8 * (size_t)(p - 1);
        const double w = E[6];
        for (int r = 0; r < 6; r++) {
          double s = 0;
          for (int k = 0; k < 6; k++) s += Jj[6 * k + r] * E[k];
          b[r] -= w * s;
          for (int c = 0; c < 6; c++) {
            double h = 0;
            for (int k = 0; k < 6; k++) h += Jj[6 * k + r] * Jj[6 * k + c];
            H[6 * r + c] += w * h;
          }
        }
      }
      double m = 0;
      for (int k = 0; k < 36; k++) a.Hpp[36 * (size_t)p + k] = H[k];
      for (int r = 0; r < 6; r++) { a.bp[6 * p + r] = b[r]; m = fmax(m, fabs(H[7 * r])); }
      a.part[rank * 4 + 3] = (p == rank) ? m : fmax(a.part[rank * 4 + 3], m);
    }
    __syncthreads();
  }
}

// (H_ll + lambda I)^-1, c_l, T = Hpl * Dinv; fail flag when a point block is not positive definite
__device__ void phase_prepare(const BaArgs& a, double lambda, int G, int GT, int* fail) {
  for (int l = G; l < a.P; l += GT) {
    const double* H = a.Hll + 6 * (size_t)l;
    const double a00 = H[0] + lambda, a01 = H[1], a02 = H[2], a11 = H[3] + lambda, a12 = H[4], a22 = H[5] + lambda;
    // Cholesky pivots (positive-definiteness test identical to the sparse LL^T with this elimination order)
    bool ok = a00 > 0;
    const double l00 = sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
    const double d1 = a11 - l10 * l10;
    ok = ok && d1 > 0;
    const double l11 = sqrt(d1), l21 = (a12 - l20 * l10) / l11;
    const double d2 = a22 - l20 * l20 - l21 * l21;
    ok = ok && d2 > 0;
    if (!ok) { atomicOr(fail, 1); continue; }
    const double l22 = sqrt(d2);
    // inverse of L (lower), Dinv = L^-T L^-1
    const double i00 = 1 / l00, i11 = 1 / l11, i22 = 1 / l22;
    const double i10 = -l10 * i00 * i11, i21 = -l21 * i11 * i22, i20 = -(l20 * i00 + l21 * i10) * i22;
    double D[6];
    D[0] = i00 * i00 + i10 * i10 + i20 * i20;
    D[1] = i10 * i11 + i20 * i21;
    D[2] = i20 * i22;
    D[3] = i11 * i11 + i21 * i21;
    D[4] = i21 * i22;
    D[5] = i22 * i22;
    for (int k = 0; k < 6; k++) a.Dinv[6 * (size_t)l + k] = D[k];
    const double* b = a.bl + 3 * (size_t)l;
    a.cl[3 * (size_t)l + 0] = D[0] * b[0] + D[1] * b[1] + D[2] * b[2];
    a.cl[3 * (size_t)l + 1] = D[1] * b[0] + D[3] * b[1] + D[4] * b[2];
    a.cl[3 * (size_t)l + 2] = D[2] * b[0] + D[4] * b[1] + D[5] * b[2];
    for (int o = a.pt_start[l]; o < a.pt_start[l + 1]; o++) {
      const double* h = a.Hpl + 18 * (size_t)o;
      double* t = a.Tpl + 18 * (size_t)o;
      for (int r = 0; r < 6; r++) {
        t[3 * r + 0] = h[3 * r] * D[0] + h[3 * r + 1] * D[1] + h[3 * r + 2] * D[2];
        t[3 * r + 1] = h[3 * r] * D[1] + h[3 * r + 1] * D[3] + h[3 * r + 2] * D[4];
        t[3 * r + 2] = h[3 * r] * D[2] + h[3 * r + 1] * D[4] + h[3 * r + 2] * D[5];
      }
    }
  }
}

// reduced camera system: S(p1,p2) = Hpp(p1,p2) + lambda I - sum_l T(p1,l) Hpl(p2,l)^T ; warp per block pair,
// lanes over the observations of pose p1, 36-value warp-shuffle reduction.  bred(p) likewise.
__device__ void phase_schur(const BaArgs& a, double lambda, int rank, int nranks) {
  const int W = a.W, n = 6 * W;
  const int lane = threadIdx.x & 31;
  const int gw = rank * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = nranks * (blockDim.x >> 5);
  const int npairs = W * (W + 1) / 2;
  for (int job = gw; job < npairs + W; job += nw) {
    if (job < npairs) {
      // unrank (p1 <= p2)
      int p1 = 0, rem = job;
      while (rem >= W - p1) { rem -= W - p1; p1++; }
      const int p2 = p1 + rem;
      double acc[36];
#pragma unroll
      for (int k = 0; k < 36; k++) acc[k] = 0;
      for (int k = a.pose_start[p1] + lane; k < a.pose_start[p1 + 1]; k += 32) {
        const int o1 = a.pose_obs[k], l = a.obs_point[o1];
        int o2 = -1;
        if (p2 == p1) o2 = o1;
        else
          for (int q = a.pt_start[l]; q < a.pt_start[l + 1]; q++)
            if (a.obs_pose[q] == p2) { o2 = q; break; }
        if (o2 < 0) continue;
        const double* t = a.Tpl + 18 * (size_t)o1;
        const double* h = a.Hpl + 18 * (size_t)o2;
#pragma unroll
        for (int r = 0; r < 6; r++)
#pragma unroll
          for (int c = 0; c < 6; c++) acc[6 * r + c] += t[3 * r] * h[3 * c] + t[3 * r + 1] * h[3 * c + 1] + t[3 * r + 2] * h[3 * c + 2];
      }
#pragma unroll
      for (int k = 0; k < 36; k++) acc[k] = warp_sum(acc[k]);
      if (lane == 0) {
        for (int r = 0; r < 6; r++)
          for (int c = 0; c < 6; c++) {
            double h = 0;
            if (p2 == p1) h = a.Hpp[36 * (size_t)p1 + 6 * r + c] + ((r == c) ? lambda : 0.0);
            else if (p2 == p1 + 1) h = a.Hoff[36 * (size_t)p1 + 6 * r + c];
            const double v = h - acc[6 * r + c];
            // lower triangle: row index >= column index; block (p1,p2) with p1<=p2 is stored transposed
            a.S[(size_t)(6 * p2 + c) * n + 6 * p1 + r] = v;
            if (p1 == p2) a.S[(size_t)(6 * p1 + r) * n + 6 * p1 + c] = v;
          }
      }
    } else {
      const int p = job - npairs;
      double acc[6] = {0, 0, 0, 0, 0, 0};
      for (int k = a.pose_start[p] + lane; k < a.pose_start[p + 1]; k += 32) {
        const int o = a.pose_obs[k], l = a.obs_point[o];
        const double* h = a.Hpl + 18 * (size_t)o;
        const double* c = a.cl + 3 * (size_t)l;
#pragma unroll
        for (int r = 0; r < 6; r++) acc[r] += h[3 * r] * c[0] + h[3 * r + 1] * c[1] + h[3 * r + 2] * c[2];
      }
#pragma unroll
      for (int r = 0; r < 6; r++) acc[r] = warp_sum(acc[r]);
      if (lane == 0)
        for (int r = 0; r < 6; r++) a.bred[6 * p + r] = a.bp[6 * p + r] - acc[r];
    }
  }
}

// dense LL^T + solve of the reduced system by one CTA in shared memory (n <= 6*BA_MAX_W)
__device__ void phase_chol(const BaArgs& a, double* Ls /* n*n */, double* ys /* n */, int* fail) {
  const int n = 6 * a.W, tid = threadIdx.x, nt = blockDim.x;
  __shared__ int s_bad;
  if (tid == 0) s_bad = 0;
  for (int i = tid; i < n * n; i += nt) {
    const int r = i / n, c = i - r * n;
    Ls[i] = (c <= r) ? a.S[i] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < n; j++) {
    const double d = Ls[j * n + j];
    if (!(d > 0)) {  // uniform: every thread reads the same value
      if (tid == 0) s_bad = 1;
      break;
    }
    const double ljj = sqrt(d);
    __syncthreads();
    for (int i = j + tid; i < n; i += nt) Ls[i * n + j] = (i == j) ? ljj : Ls[i * n + j] / ljj;
    __syncthreads();
    // trailing update of the lower triangle
    const int m = n - j - 1;
    for (int e = tid; e < m * (m + 1) / 2; e += nt) {
      // unrank (r >= c) within the m x m trailing block
      int r = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
      while ((r + 1) * (r + 2) / 2 <= e) r++;
      while (r * (r + 1) / 2 > e) r--;
      const int c = e - r * (r + 1) / 2;
      const int gi = j + 1 + r, gk = j + 1 + c;
      Ls[gi * n + gk] -= Ls[gi * n + j] * Ls[gk * n + j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (s_bad) {
    if (tid == 0) atomicOr(fail, 2);
    return;
  }
  if (tid < 32) {  // forward / backward substitution by one warp
    const int lane = tid;
    for (int i = 0; i < n; i++) {
      double s = 0;
      for (int k = lane; k < i; k += 32) s += Ls[i * n + k] * ys[k];
      s = warp_sum(s);
      if (lane == 0) ys[i] = (a.bred[i] - s) / Ls[i * n + i];
      __syncwarp();
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = 0;
      for (int k = i + 1 + lane; k < n; k += 32) s += Ls[k * n + i] * ys[k];
      s = warp_sum(s);
      if (lane == 0) ys[i] = (ys[i] - s) / Ls[i * n + i];
      __syncwarp();
    }
    for (int i = lane; i < n; i += 32) a.xp[i] = ys[i];
  }
  __syncthreads();
}

// back-substitution of the points, trial state = state (+) x, scale = x^T (lambda x + b)
__device__ void phase_update(const BaArgs& a, double lambda, int cur, int failed, int G, int GT, int rank, double* red) {
  const int trial = cur ^ 1;
  const Pose* X = a.X + (size_t)cur * a.W;
  Pose* Xt = a.X + (size_t)trial * a.W;
  const double* pts = a.pts + (size_t)cur * 3 * a.P;
  double* ptt = a.pts + (size_t)trial * 3 * a.P;
  double scale = 0;
  for (int l = G; l < a.P; l += GT) {
    double x[3];
    const double* b = a.bl + 3 * (size_t)l;
    if (failed) { x[0] = b[0]; x[1] = b[1]; x[2] = b[2]; }  // linear solver leaves x = b on failure
    else {
      x[0] = a.cl[3 * (size_t)l]; x[1] = a.cl[3 * (size_t)l + 1]; x[2] = a.cl[3 * (size_t)l + 2];
      for (int o = a.pt_start[l]; o < a.pt_start[l + 1]; o++) {
        const double* t = a.Tpl + 18 * (size_t)o;
        const double* xp = a.xp + 6 * a.obs_pose[o];
        for (int r = 0; r < 6; r++) {
          x[0] -= t[3 * r] * xp[r];
          x[1] -= t[3 * r + 1] * xp[r];
          x[2] -= t[3 * r + 2] * xp[r];
        }
      }
    }
    for (int k = 0; k < 3; k++) {
      ptt[3 * (size_t)l + k] = pts[3 * (size_t)l + k] + x[k];
      scale += x[k] * (lambda * x[k] + b[k]);
    }
  }
  for (int p = G; p < a.W; p += GT) {
    double u[6];
    for (int r = 0; r < 6; r++) {
      u[r] = failed ? a.bp[6 * p + r] : a.xp[6 * p + r];
      scale += u[r] * (lambda * u[r] + a.bp[6 * p + r]);
    }
    Pose o;
    pose_oplus(X[p], u, o);
    Xt[p] = o;
  }
  double v[1] = {scale};
  block_sum<1>(v, red);
  if (threadIdx.x == 0) a.part[rank * 4 + 1] = red[0];
}

__device__ void phase_output(const BaArgs& a, int cur, int G, int GT) {
  const Pose* X = a.X + (size_t)cur * a.W;
  const double* pts = a.pts + (size_t)cur * 3 * a.P;
  for (int i = G; i < a.W; i += GT) pose_to_f32(X[i], a.out_poses + 16 * i);
  for (int i = G; i < 3 * a.P; i += GT) a.out_points[i] = (float)pts[i];
}

// relative motions from the float32 poses, float32 arithmetic like cv::Mat (src/Optimizer.cc:1072-1075)
__device__ void phase_output_rel(const BaArgs& a, int G, int GT) {
  for (int i = 1 + G; i < a.W; i += GT) {
    const float* A = a.out_poses + 16 * (i - 1);
    const float* B = a.out_poses + 16 * i;
    float Ai[16];
    for (int k = 0; k < 16; k++) Ai[k] = 0.f;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Ai[4 * r + c] = A[4 * c + r];
    for (int r = 0; r < 3; r++)
      Ai[4 * r + 3] = -__fadd_rn(__fadd_rn(__fmul_rn(Ai[4 * r], A[3]), __fmul_rn(Ai[4 * r + 1], A[7])), __fmul_rn(Ai[4 * r + 2], A[11]));
    Ai[15] = 1.f;
    float* out = a.out_rel + 16 * (i - 1);
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        float s = 0.f;
        for (int k = 0; k < 4; k++) s = __fadd_rn(s, __fmul_rn(Ai[4 * r + k], B[4 * k + c]));
        out[4 * r + c] = s;
      }
  }
}

// ---------------------------------------------------------------------------------------------------------
// the cluster kernel
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define BA_TIC() unsigned long long t0_ = (G == 0) ? gtime() : 0
#define BA_TOC(slot) do { if (G == 0) { unsigned long long t1_ = gtime(); ctl->t_phase[slot] += t1_ - t0_; t0_ = t1_; } } while (0)

__global__ void __cluster_dims__(BA_CLUSTER, 1, 1) __launch_bounds__(BA_THREADS, 1) ba_window_kernel(BaArgs a) {
  extern __shared__ __align__(16) double dsm[];  // CTA 0: (6W)^2 + 6W doubles for the dense factorisation
  __shared__ double red[16 * 27 + 32];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), nranks = (int)cluster.num_blocks();
  const int G = rank * blockDim.x + threadIdx.x, GT = nranks * blockDim.x;
  BaCtl* ctl = a.ctl;
  int* failp = &ctl->fail;

  phase_init(a, G, GT);
  if (G == 0) {
    ctl->cur = 0; ctl->it = 0; ctl->stop_flag = 0; ctl->ok = 1; ctl->fail = 0; ctl->nBad = 0;
    ctl->chi2_check = 0; ctl->lastChi = 0; ctl->iterations = 0; ctl->n_records = 0; ctl->total_trials = 0;
    ctl->lambda = -1; ctl->ni = 2;
    for (int k = 0; k < 8; k++) ctl->t_phase[k] = 0;
  }
  cluster.sync();
  const unsigned long long t_start = (G == 0) ? gtime() : 0;
  BA_TIC();
  if (a.W + a.P == 0 || a.max_iterations <= 0) {
    if (G == 0) ctl->iterations = (a.W + a.P == 0) ? -1 : 0;
    phase_output(a, 0, G, GT);
    cluster.sync();
    phase_output_rel(a, G, GT);
    return;
  }
  // robust chi2 of the initial state
  phase_errors(a, 0, G, GT, rank, red);
  cluster.sync();
  if (G == 0) {
    double c = 0;
    for (int r = 0; r < nranks; r++) c += a.part[r * 4];
    ctl->currentChi = c;
  }
  cluster.sync();

  const double tau = 1e-5, upper = 2. / 3., lower = 1. / 3.;
  for (int it = 0; it < a.max_iterations; it++) {
    if (ctl->stop_flag || !ctl->ok) break;  // uniform: written before the last cluster barrier
    const int cur = ctl->cur;
    BA_TOC(6);
    phase_linearize(a, cur, G, GT, rank, nranks, red);
    cluster.sync();
    BA_TOC(0);
    if (G == 0) {
      ctl->iniChi = ctl->currentChi;
      ctl->tempChi = ctl->currentChi;
      if (it == 0) {
        double m = 0;
        for (int r = 0; r < nranks; r++) m = fmax(m, fmax(a.part[r * 4 + 2], (r < a.W) ? a.part[r * 4 + 3] : 0.0));
        ctl->lambda = tau * m;
        ctl->ni = 2;
        ctl->nBad = 0;
      }
      ctl->qmax = 0;
      ctl->rho = 0;
    }
    cluster.sync();
    // ---- trials
    while (true) {
      const double lambda = ctl->lambda;
      BA_TOC(6);
      phase_prepare(a, lambda, G, GT, failp);
      cluster.sync();
      BA_TOC(1);
      int failed = ctl->fail;
      if (!failed) {
        phase_schur(a, lambda, rank, nranks);
        cluster.sync();
        BA_TOC(2);
        if (rank == 0) phase_chol(a, dsm, dsm + (size_t)36 * a.W * a.W, failp);
        cluster.sync();
        BA_TOC(3);
        failed = ctl->fail;
      }
      phase_update(a, lambda, cur, failed, G, GT, rank, red);
      cluster.sync();
      BA_TOC(4);
      phase_errors(a, cur ^ 1, G, GT, rank, red);
      cluster.sync();
      BA_TOC(5);
      if (G == 0) {
        double chi = 0, scale = 0;
        for (int r = 0; r < nranks; r++) { chi += a.part[r * 4]; scale += a.part[r * 4 + 1]; }
        ctl->lastTrialChi = chi;
        double tempChi = failed ? DBL_MAX : chi;
        double rho = (ctl->currentChi - tempChi);
        scale += 1e-3;
        rho /= scale;
        int accepted = 0;
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - pow((2 * rho - 1), 3);
          alpha = fmin(alpha, upper);
          const double sf = fmax(lower, alpha);
          ctl->lambda *= sf;
          ctl->ni = 2;
          ctl->currentChi = tempChi;
          ctl->cur = cur ^ 1;
          accepted = 1;
        } else {
          ctl->lambda *= ctl->ni;
          ctl->ni *= 2;
        }
        ctl->rho = rho;
        ctl->tempChi = tempChi;
        ctl->accepted = accepted;
        ctl->qmax += 1;
        ctl->fail = 0;
      }
      cluster.sync();
      if (!(ctl->rho < 0 && ctl->qmax < 10)) break;
    }
    // ---- end of the LM step: stop rules, statistics, terminate action
    if (G == 0) {
      int result_ok;
      if (ctl->qmax == 10 || ctl->rho == 0) result_ok = 0;
      else {
        if ((ctl->iniChi - ctl->currentChi) * 1e3 < ctl->iniChi) ctl->nBad++;
        else ctl->nBad = 0;
        result_ok = ctl->nBad < 3;
      }
      int ok = result_ok;
      const double arc = ctl->lastTrialChi;  // activeRobustChi2() sees the errors of the last trial
      if (ctl->chi2_check < arc && it > 0) ok = 0;
      ctl->chi2_check = arc;
      ctl->total_trials += ctl->qmax;
      if (ctl->n_records < BA_MAX_REC) {
        BaRec& r = a.rec[ctl->n_records++];
        r.chi2 = ctl->currentChi;
        r.lambda = ctl->lambda;
        r.trials = ctl->qmax;
      }
      ctl->iterations = it + 1;
      if (a.gain_threshold >= 0) {
        const double chi = ctl->currentChi;  // errors recomputed at the accepted state
        if (it == 0) ctl->lastChi = chi;
        else {
          const double gain = (ctl->lastChi - chi) / chi;
          ctl->lastChi = chi;
          if (gain >= 0 && gain < a.gain_threshold) ctl->stop_flag = 1;
        }
      }
      ctl->ok = ok;
      ctl->it = it + 1;
    }
    cluster.sync();
  }
  phase_output(a, ctl->cur, G, GT);
  cluster.sync();
  phase_output_rel(a, G, GT);
  if (G == 0) ctl->t_phase[7] = gtime() - t_start;
}

// =========================================================================================================
// host side
// =========================================================================================================
struct BaWorkspace {
  int capW = 0, capP = 0, capM = 0;
  char* d_base = nullptr;
  size_t bytes = 0;
  BaArgs args;
  // host staging (pinned)
  char* h_base = nullptr;
  size_t h_bytes = 0;
};

static size_t al(size_t v) { return (v + 255) & ~(size_t)255; }

template <class T>
static T* carve(char*& p, size_t n) {
  T* r = (T*)p;
  p += al(sizeof(T) * n);
  return r;
}

int ba_setup(vido_ctx* ctx, int capW, int capP, int capM) {
  BaWorkspace* ws = new BaWorkspace();
  ctx->ba = ws;
  ws->capW = capW; ws->capP = capP; ws->capM = capM;
  size_t need = 0;
  {
    char* p = nullptr;
    carve<float>(p, 16 * capW); carve<float>(p, 16 * capW); carve<float>(p, 3 * (size_t)capP);
    carve<int>(p, capM); carve<float>(p, 3 * (size_t)capM); carve<int>(p, capP + 1); carve<int>(p, capW + 1);
    carve<int>(p, capM); carve<int>(p, capM);
    carve<Pose>(p, 2 * capW); carve<Pose>(p, capW); carve<double>(p, 6 * (size_t)capP);
    carve<double>(p, 6 * (size_t)capP); carve<double>(p, 3 * (size_t)capP); carve<double>(p, 6 * (size_t)capP);
    carve<double>(p, 3 * (size_t)capP); carve<double>(p, 18 * (size_t)capM); carve<double>(p, 18 * (size_t)capM);
    carve<double>(p, 36 * capW); carve<double>(p, 36 * capW); carve<double>(p, 6 * capW);
    carve<double>(p, 36 * (size_t)capW * capW); carve<double>(p, 6 * capW); carve<double>(p, 6 * capW);
    carve<double>(p, 3 * (size_t)capP); carve<double>(p, 4 * BA_CLUSTER); carve<double>(p, 72 * capW);
    carve<double>(p, 8 * capW); carve<BaCtl>(p, 1); carve<BaRec>(p, BA_MAX_REC);
    carve<float>(p, 16 * capW); carve<float>(p, 16 * capW); carve<float>(p, 3 * (size_t)capP);
    need = (size_t)p;
  }
  ws->bytes = need;
  VIDO_CUDA(cudaMalloc(&ws->d_base, need));
  VIDO_CUDA(cudaMemset(ws->d_base, 0, need));
  char* p = ws->d_base;
  BaArgs& a = ws->args;
  memset(&a, 0, sizeof a);
  a.poses_f32 = carve<float>(p, 16 * capW); a.rel_f32 = carve<float>(p, 16 * capW);
  a.points_f32 = carve<float>(p, 3 * (size_t)capP);
  a.obs_pose = carve<int>(p, capM); a.obs_xyz = carve<float>(p, 3 * (size_t)capM);
  a.pt_start = carve<int>(p, capP + 1); a.pose_start = carve<int>(p, capW + 1);
  a.pose_obs = carve<int>(p, capM); a.obs_point = carve<int>(p, capM);
  a.X = carve<Pose>(p, 2 * capW); a.Zinv = carve<Pose>(p, capW); a.pts = carve<double>(p, 6 * (size_t)capP);
  a.Hll = carve<double>(p, 6 * (size_t)capP); a.bl = carve<double>(p, 3 * (size_t)capP);
  a.Dinv = carve<double>(p, 6 * (size_t)capP); a.cl = carve<double>(p, 3 * (size_t)capP);
  a.Hpl = carve<double>(p, 18 * (size_t)capM); a.Tpl = carve<double>(p, 18 * (size_t)capM);
  a.Hpp = carve<double>(p, 36 * capW); a.Hoff = carve<double>(p, 36 * capW); a.bp = carve<double>(p, 6 * capW);
  a.S = carve<double>(p, 36 * (size_t)capW * capW); a.bred = carve<double>(p, 6 * capW); a.xp = carve<double>(p, 6 * capW);
  a.xl = carve<double>(p, 3 * (size_t)capP); a.part = carve<double>(p, 4 * BA_CLUSTER);
  a.seJ = carve<double>(p, 72 * capW); a.seE = carve<double>(p, 8 * capW);
  a.ctl = carve<BaCtl>(p, 1); a.rec = carve<BaRec>(p, BA_MAX_REC);
  a.out_poses = carve<float>(p, 16 * capW); a.out_rel = carve<float>(p, 16 * capW);
  a.out_points = carve<float>(p, 3 * (size_t)capP);
  // pinned host staging for the host-pointer API
  {
    char* hp = nullptr;
    carve<int>(hp, capM); carve<int>(hp, capM); carve<int>(hp, capM); carve<int>(hp, capM);
    carve<int>(hp, capP + 1); carve<int>(hp, capW + 1); carve<float>(hp, 3 * (size_t)capM);
    ws->h_bytes = (size_t)hp + 4096;
  }
  VIDO_CUDA(cudaMallocHost(&ws->h_base, ws->h_bytes));
  const size_t smem = sizeof(double) * ((size_t)36 * capW * capW + 6 * capW);
  if (smem > 200 * 1024) { ctx->err = "BA window too large for the shared-memory Cholesky"; return VIDO_ERR_ARG; }
  VIDO_CUDA(cudaFuncSetAttribute(ba_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return VIDO_OK;
}

void ba_teardown(vido_ctx* ctx) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  if (!ws) return;
  cudaFree(ws->d_base);
  cudaFreeHost(ws->h_base);
  delete ws;
  ctx->ba = nullptr;
}

int ba_partial_host(vido_ctx* ctx, vido_ba_problem* pr, vido_lm_stats* st) {
  BaWorkspace* ws = (BaWorkspace*)ctx->ba;
  const int W = pr->n_poses, P = pr->n_points, M = pr->n_obs;
  if (W > ws->capW || P > ws->capP || M > ws->capM || W > BA_MAX_W) { ctx->err = "BA problem exceeds the context capacity"; return VIDO_ERR_CAPACITY; }
  if (W < 0 || P < 0 || M < 0) return VIDO_ERR_ARG;
  cudaStream_t s = ctx->stream;
  BaArgs a = ws->args;
  a.W = W; a.P = P; a.M = M;
  a.max_iterations = pr->max_iterations;
  a.info_cam = 1.0 / (double)pr->sigma2_cam;
  a.info_3d = 1.0 / (double)pr->sigma2_3d;
  a.d_cam = (double)pr->huber_cam;
  a.d_3d = (double)pr->huber_3d;
  a.gain_threshold = (double)pr->gain_threshold;
  // ---- host-side index structures: observations sorted by (point, pose); CSR by point and by pose
  char* hp = ws->h_base;
  int* h_order = carve<int>(hp, M);
  int* h_obs_pose = carve<int>(hp, M);
  int* h_obs_point = carve<int>(hp, M);
  int* h_pose_obs = carve<int>(hp, M);
  int* h_pt_start = carve<int>(hp, P + 1);
  int* h_pose_start = carve<int>(hp, W + 1);
  float* h_xyz = carve<float>(hp, 3 * (size_t)M);
  for (int l = 0; l <= P; l++) h_pt_start[l] = 0;
  for (int p = 0; p <= W; p++) h_pose_start[p] = 0;
  for (int o = 0; o < M; o++) {
    const int l = pr->obs_point[o], p = pr->obs_pose[o];
    if (l < 0 || l >= P || p < 0 || p >= W) { ctx->err = "BA observation index out of range"; return VIDO_ERR_ARG; }
    h_pt_start[l + 1]++;
  }
  for (int l = 0; l < P; l++) h_pt_start[l + 1] += h_pt_start[l];
  {
    // stable counting sort by point; within a point the caller's order is frame order (frame-major input)
    std::vector<int> fill(h_pt_start, h_pt_start + P);
    for (int o = 0; o < M; o++) h_order[fill[pr->obs_point[o]]++] = o;
  }
  for (int k = 0; k < M; k++) {
    const int o = h_order[k];
    h_obs_pose[k] = pr->obs_pose[o];
    h_obs_point[k] = pr->obs_point[o];
    h_xyz[3 * k] = pr->obs_xyz[3 * o]; h_xyz[3 * k + 1] = pr->obs_xyz[3 * o + 1]; h_xyz[3 * k + 2] = pr->obs_xyz[3 * o + 2];
    h_pose_start[h_obs_pose[k] + 1]++;
  }
  for (int p = 0; p < W; p++) h_pose_start[p + 1] += h_pose_start[p];
  {
    std::vector<int> fill(h_pose_start, h_pose_start + W);
    for (int k = 0; k < M; k++) h_pose_obs[fill[h_obs_pose[k]]++] = k;
  }
  VIDO_CUDA(cudaMemcpyAsync((void*)a.poses_f32, pr->poses, sizeof(float) * 16 * W, cudaMemcpyHostToDevice, s));
  if (W > 1) VIDO_CUDA(cudaMemcpyAsync((void*)a.rel_f32, pr->rel_motion, sizeof(float) * 16 * (W - 1), cudaMemcpyHostToDevice, s));
  if (P) VIDO_CUDA(cudaMemcpyAsync((void*)a.points_f32, pr->points, sizeof(float) * 3 * P, cudaMemcpyHostToDevice, s));
  if (M) {
    VIDO_CUDA(cudaMemcpyAsync((void*)a.obs_pose, h_obs_pose, sizeof(int) * M, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync((void*)a.obs_point, h_obs_point, sizeof(int) * M, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync((void*)a.pose_obs, h_pose_obs, sizeof(int) * M, cudaMemcpyHostToDevice, s));
    VIDO_CUDA(cudaMemcpyAsync((void*)a.obs_xyz, h_xyz, sizeof(float) * 3 * M, cudaMemcpyHostToDevice, s));
  }
  VIDO_CUDA(cudaMemcpyAsync((void*)a.pt_start, h_pt_start, sizeof(int) * (P + 1), cudaMemcpyHostToDevice, s));
  VIDO_CUDA(cudaMemcpyAsync((void*)a.pose_start, h_pose_start, sizeof(int) * (W + 1), cudaMemcpyHostToDevice, s));
  const size_t smem = sizeof(double) * ((size_t)36 * W * W + 6 * W);
  cudaEventRecord(ctx->ev0, s);
  ba_window_kernel<<<BA_CLUSTER, BA_THREADS, smem, s>>>(a);
  cudaEventRecord(ctx->ev1, s);
  ctx->launches++;
  VIDO_CUDA(cudaGetLastError());
  BaCtl ctl;
  VIDO_CUDA(cudaMemcpyAsync(&ctl, a.ctl, sizeof ctl, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaMemcpyAsync(pr->poses, a.out_poses, sizeof(float) * 16 * W, cudaMemcpyDeviceToHost, s));
  if (W > 1) VIDO_CUDA(cudaMemcpyAsync(pr->rel_motion, a.out_rel, sizeof(float) * 16 * (W - 1), cudaMemcpyDeviceToHost, s));
  if (P) VIDO_CUDA(cudaMemcpyAsync(pr->points, a.out_points, sizeof(float) * 3 * P, cudaMemcpyDeviceToHost, s));
  static BaRec recs[BA_MAX_REC];
  if (st) VIDO_CUDA(cudaMemcpyAsync(recs, a.rec, sizeof(BaRec) * BA_MAX_REC, cudaMemcpyDeviceToHost, s));
  VIDO_CUDA(cudaStreamSynchronize(s));
  {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) { ctx->t_ms[3] += ms; ctx->t_n[3]++; }
    const double edges = (double)M + (double)std::max(W - 1, 0);
    ctx->ba_alg_bytes += edges * (296.0 * std::max(ctl.iterations, 0) + 152.0 * (ctl.total_trials + 1));
  }
  if (getenv("VIDO_BA_TIMING"))
    fprintf(stderr, "[ba] its=%d trials=%d ns: linearize=%llu prepare=%llu schur=%llu chol=%llu update=%llu errors=%llu decide=%llu total=%llu\n",
            ctl.iterations, ctl.total_trials, ctl.t_phase[0], ctl.t_phase[1], ctl.t_phase[2], ctl.t_phase[3], ctl.t_phase[4],
            ctl.t_phase[5], ctl.t_phase[6], ctl.t_phase[7]);
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
