// fba_kernels.cu -- full-sequence graph optimisation on the GPU.
//
// Replaces Optimizer::FullBatchOptimization (src/Optimizer.cc:1235-2178) + the embedded g2o it drives:
//   EdgeSE3Prior (identity offset)      g2o/types/edge_se3_prior.cpp:89-102  (= EdgeSE3 with Xi = I fixed, see ba_math.h)
//   EdgeSE3 (odometry, smoothness)      g2o/types/edge_se3.cpp:77-105
//   EdgeSE3PointXYZ (static, dynamic)   g2o/types/edge_se3_pointxyz.cpp:99-140
//   LandmarkMotionTernaryEdge           g2o/types/types_dyn_slam3d.cpp:53-85
//   BaseMultiEdge::constructQuadraticForm g2o/core/base_multi_edge.hpp:36-48,171-225 ; Huber robust_kernel_impl.cpp:78-91
//   Levenberg-Marquardt driver          lm_device.h (the same state machine as the window solver, stepped by the host here:
//                                       the full graph is solved once per sequence and one trial is a chain of launches)
//
// Linear system.  The reference factors the whole H with CSparse.  Here the points are eliminated first -- exactly, no
// approximation: a static point is a 3x3 pivot; the points of one dynamic tracklet are coupled to each other by the ternary
// edges and form a block-tridiagonal pivot that is factored along the chain (one warp per chain; the running 6x3 blocks
// Y_v = (B L^-T)_{v,k} of the SE3 vertices met so far live in shared memory).  The Schur complement on the SE3 vertices
// (camera poses + object motions, 6 x 6 blocks) is assembled densely in HBM (FP64) and factored by a tiled right-looking
// Cholesky (diagonal tile / panel solve / trailing update launches).  Point diagonal blocks are w*I because the point
// Jacobians are rotations (J^T J = I), so they are stored as scalars.
//
// Data layout: SE3 states as 12 doubles (R row-major, t), double-buffered (current / trial); edges as structure-of-arrays;
// per-point coupling lists (CSR) to the SE3 vertices; chains as CSR over points.
#include <algorithm>
#include <cstring>
#include <vector>

#include "ba_math.h"
#include <chrono>

#include "ctx.h"
#include "lm_device.h"

namespace {

using vb::Pose;

constexpr int FBA_MAX_ACTIVE = 256;   // SE3 vertices coupled to one chain (tracklet length * 2 + 1 for dynamic chains)
constexpr int FBA_CHAIN_WARPS = 1;    // chains per CTA of the elimination kernel (36 KB of running blocks per chain)
constexpr int FBA_TB = 48;            // Cholesky tile (8 SE3 blocks)

struct FbaDev {
  int NS, NP, NO, NE, NT, NCH, n;     // n = 6 NS
  Pose* X[2];                          // SE3 estimates: current / trial (index = LmCtl::cur)
  double* P[2];                        // points [NP][3]
  Pose Zprior_inv;
  const Pose* Z6inv;                   // [NE] inverse measurements
  const int *e6i, *e6j, *e6k;
  const int *os, *op, *ok;
  const double* meas;                  // [NO][3]
  const int *t1, *t2, *th;
  double info6[2], d6[2], info3[2], d3, infoT, dT, infoPrior;
  // linearisation
  double *Hss, *bs;                    // [n][n] row-major (lower triangle + diagonal blocks used), [n]
  double *hl, *bl;                     // [NP], [NP][3]
  double *Bo, *B1, *B2, *Ot;           // [NO][18], [NT][18], [NT][18], [NT][9]
  // trial
  double *S, *bp, *x;                  // working copy of Hss (+lambda, - Schur terms), rhs, solution [n + 3 NP]
  double *Lkk, *Lk1;                   // [NP][6] lower 3x3 factor, [NP][9] sub-diagonal block L_{k,k-1}
  const int *chain_start, *chain_pts, *chain_link;
  const int *cpl_start, *cpl_v, *cpl_kind, *cpl_e;
  double* partial;                     // reduction scratch
  int* fail;
  // ---- matrix-free path (pcg != 0): the SE3 block is kept as diagonal blocks + one block per SE3 edge, the Schur
  //      complement is applied implicitly (B T^-1 B^T through the factored chains), solved by preconditioned CG
  int pcg;
  double *Hd, *He6;                    // [NS][36], [NE][36] (block row = e6i, column = e6j)
  const int *cpl_pt;                   // [ncpl] point of a coupling
  const int *vc_start, *vc_ci;         // SE3 vertex -> its couplings
  const int *ve_start, *ve_e, *ve_side;// SE3 vertex -> its SE3 edges (side 0: vertex is e6i, 1: e6j)
  double *cg_r, *cg_z, *cg_p, *cg_q;   // [n]
  double *cg_t;                        // [NP][3]
  double *cg_M;                        // [NS][21] lower Cholesky factors of the preconditioner's diagonal pivots
  double *cg_A;                        // [NS][36] preconditioner diagonal blocks before factorisation
  double *cg_W;                        // [NS][36] sub-diagonal factor blocks along the SE3-edge paths
  const int *path_start, *path_v, *path_e;  // paths of the SE3-edge graph (odometry chain, one smoothness chain per object)
  int NPATH;
  double *cg_part;                     // [2][NS] per-vertex partial dot products
  double *cg_s;                        // scalars: 0 rz, 1 pq, 2 rr, 3 alpha, 4 beta, 5 rz_new, 6 bnorm2
};

// ---------------------------------------------------------------------------------------------------------
// edge math
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tern_error(const Pose& H, const double* p1, const double* p2, double* q, double* e) {
  const double d0 = p2[0] - H.t[0], d1 = p2[1] - H.t[1], d2 = p2[2] - H.t[2];
  for (int i = 0; i < 3; i++) {
    q[i] = H.R[i] * d0 + H.R[3 + i] * d1 + H.R[6 + i] * d2;   // H^-1 p2
    e[i] = p1[i] - q[i];
  }
}

__device__ __forceinline__ double huber_rho(double e, double delta, double* w) {
  double rho0, ww;
  vb::huber(e, delta, rho0, ww);
  if (w) *w = ww;
  return rho0;
}

// deterministic block sum -> partial[blockIdx.x]
__device__ __forceinline__ void block_sum_store(double v, double* partial) {
  __shared__ double sm[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sm[w];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) fba_final_sum_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  // fixed-order sum of the block partials (one warp): the chi2 / scale values that drive the LM decisions do not depend
  // on the launch's scheduling
  double s = 0;
  for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out = s;
}

// robust chi2 of every edge at state `sel` (activeRobustChi2)
__global__ void __launch_bounds__(256) fba_chi2_kernel(FbaDev d, int sel) {
  const Pose* X = d.X[sel];
  const double* P = d.P[sel];
  const int total = 1 + d.NE + d.NO + d.NT;
  double acc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i == 0) {
      Pose I;
      for (int k = 0; k < 9; k++) I.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
      I.t[0] = I.t[1] = I.t[2] = 0;
      double e[6];
      vb::edge_se3(I, X[0], d.Zprior_inv, e, nullptr, nullptr);
      double s = 0;
      for (int k = 0; k < 6; k++) s += e[k] * e[k];
      acc += s * d.infoPrior;
    } else if (i <= d.NE) {
      const int ed = i - 1, k6 = d.e6k[ed];
      double e[6];
      vb::edge_se3(X[d.e6i[ed]], X[d.e6j[ed]], d.Z6inv[ed], e, nullptr, nullptr);
      double s = 0;
      for (int k = 0; k < 6; k++) s += e[k] * e[k];
      acc += huber_rho(s * d.info6[k6], d.d6[k6], nullptr);
    } else if (i <= d.NE + d.NO) {
      const int o = i - 1 - d.NE;
      double zc[3], e[3];
      vb::edge_xyz(X[d.os[o]], P + 3 * (size_t)d.op[o], d.meas + 3 * (size_t)o, zc, e);
      acc += huber_rho((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * d.info3[d.ok[o]], d.d3, nullptr);
    } else {
      const int t = i - 1 - d.NE - d.NO;
      double q[3], e[3];
      tern_error(X[d.th[t]], P + 3 * (size_t)d.t1[t], P + 3 * (size_t)d.t2[t], q, e);
      acc += huber_rho((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * d.infoT, d.dT, nullptr);
    }
  }
  block_sum_store(acc, d.partial);
}

// H[a][c] += w Ja^T Jc for 6-column Jacobians with `rows` rows; only blocks on or below the diagonal are kept
__device__ __forceinline__ void add_block(const FbaDev& d, int a, int c, const double* Ja, const double* Jc, int rows, double w) {
  if (d.pcg) {
    if (a != c) return;   // off-diagonal SE3 blocks only come from SE3 edges and are stored per edge (He6)
    for (int r = 0; r < 6; r++)
      for (int q = 0; q < 6; q++) {
        double s = 0;
        for (int k = 0; k < rows; k++) s += Ja[6 * k + r] * Jc[6 * k + q];
        atomicAdd(&d.Hd[36 * (size_t)a + 6 * r + q], w * s);
      }
    return;
  }
  if (a < c) return;
  const int n = d.n;
  for (int r = 0; r < 6; r++)
    for (int q = 0; q < 6; q++) {
      double s = 0;
      for (int k = 0; k < rows; k++) s += Ja[6 * k + r] * Jc[6 * k + q];
      atomicAdd(&d.Hss[(size_t)(6 * a + r) * n + 6 * c + q], w * s);
    }
}

// Diagonal-block and gradient contribution of a 3-row edge (J: 3 x 6, residual e) to SE3 vertex a, aggregated over the warp
// when all 32 lanes hold edges of the SAME vertex: the edges are stored vertex by vertex (observations frame by frame, motion
// edges object by object), so that is the common case, and 32 x 42 FP64 atomics on the same 42 addresses (the round-1
// kernel: ~20 % of its samples were L2 atomic serialisation) become one butterfly reduction of the 27 unique sums and 42
// single-lane atomics.  Returns false when the warp is not uniform: the caller then adds per thread.
__device__ __forceinline__ bool add_diag_warp(const FbaDev& d, int a, const double* J, const double* e, double w) {
  const unsigned act = __activemask();
  if (act != 0xffffffffu) return false;
  const int a0 = __shfl_sync(0xffffffffu, a, 0);
  if (!__all_sync(0xffffffffu, a == a0)) return false;
  const int lane = threadIdx.x & 31;
  double hc[21], bc[6];
  {
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
      for (int c = r; c < 6; c++) hc[idx++] = w * (J[r] * J[c] + J[6 + r] * J[6 + c] + J[12 + r] * J[12 + c]);
      bc[r] = -w * (J[r] * e[0] + J[6 + r] * e[1] + J[12 + r] * e[2]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 21; k++) hc[k] += __shfl_xor_sync(0xffffffffu, hc[k], o);
#pragma unroll
    for (int k = 0; k < 6; k++) bc[k] += __shfl_xor_sync(0xffffffffu, bc[k], o);
  }
  double* const Hb = d.pcg ? d.Hd + 36 * (size_t)a : d.Hss + (size_t)(6 * a) * d.n + 6 * a;
  const int ldh = d.pcg ? 6 : d.n;
  {
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
      for (int c = r; c < 6; c++) {
        if (lane == idx) {
          atomicAdd(&Hb[(size_t)r * ldh + c], hc[idx]);
          if (r != c) atomicAdd(&Hb[(size_t)c * ldh + r], hc[idx]);
        }
        idx++;
      }
      if (lane == 21 + r) atomicAdd(&d.bs[6 * a + r], bc[r]);
    }
  }
  return true;
}

// buildSystem: one thread per edge.  Contributions to the SE3 block / rhs and to the point diagonals / rhs are summed with
// FP64 atomics; the pose-point and point-point blocks belong to exactly one edge and are plain stores.
__global__ void __launch_bounds__(128) fba_linearize_kernel(FbaDev d, int sel) {
  const Pose* X = d.X[sel];
  const double* P = d.P[sel];
  const int total = 1 + d.NE + d.NO + d.NT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = d.n;
  if (i == 0) {
    Pose I;
    for (int k = 0; k < 9; k++) I.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
    I.t[0] = I.t[1] = I.t[2] = 0;
    double e[6], Ji[36], Jj[36];
    vb::edge_se3(I, X[0], d.Zprior_inv, e, Ji, Jj);
    const double w = d.infoPrior;
    for (int r = 0; r < 6; r++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += Jj[6 * k + r] * e[k];
      atomicAdd(&d.bs[r], -w * s);
    }
    add_block(d, 0, 0, Jj, Jj, 6, w);
  } else if (i <= d.NE) {
    const int ed = i - 1, k6 = d.e6k[ed], a = d.e6i[ed], c = d.e6j[ed];
    double e[6], Ji[36], Jj[36];
    vb::edge_se3(X[a], X[c], d.Z6inv[ed], e, Ji, Jj);
    double s2 = 0;
    for (int k = 0; k < 6; k++) s2 += e[k] * e[k];
    double hw;
    huber_rho(s2 * d.info6[k6], d.d6[k6], &hw);
    const double w = hw * d.info6[k6];
    for (int r = 0; r < 6; r++) {
      double si = 0, sj = 0;
      for (int k = 0; k < 6; k++) { si += Ji[6 * k + r] * e[k]; sj += Jj[6 * k + r] * e[k]; }
      atomicAdd(&d.bs[6 * a + r], -w * si);
      atomicAdd(&d.bs[6 * c + r], -w * sj);
    }
    add_block(d, a, a, Ji, Ji, 6, w);
    add_block(d, c, c, Jj, Jj, 6, w);
    add_block(d, a, c, Ji, Jj, 6, w);
    add_block(d, c, a, Jj, Ji, 6, w);
    if (d.pcg) {
      double* Hb = d.He6 + 36 * (size_t)ed;
      for (int r = 0; r < 6; r++)
        for (int q = 0; q < 6; q++) {
          double s3 = 0;
          for (int k = 0; k < 6; k++) s3 += Ji[6 * k + r] * Jj[6 * k + q];
          Hb[6 * r + q] = w * s3;
        }
    }
  } else if (i <= d.NE + d.NO) {
    const int o = i - 1 - d.NE, a = d.os[o], l = d.op[o];
    const Pose Xa = X[a];   // a private copy: the stores below may alias X for all the compiler knows, and it re-read R per use
    double zc[3], e[3];
    vb::edge_xyz(Xa, P + 3 * (size_t)l, d.meas + 3 * (size_t)o, zc, e);
    double hw;
    huber_rho((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * d.info3[d.ok[o]], d.d3, &hw);
    const double w = hw * d.info3[d.ok[o]];
    // J_pose = [-I | Q(zc)], Q = 2 [[0,-z,y],[z,0,-x],[-y,x,0]];  J_point = R^T
    double Ji[18];
    for (int k = 0; k < 18; k++) Ji[k] = 0;
    Ji[0] = -1; Ji[7] = -1; Ji[14] = -1;
    Ji[4] = -2 * zc[2]; Ji[5] = 2 * zc[1];
    Ji[9] = 2 * zc[2];  Ji[11] = -2 * zc[0];
    Ji[15] = -2 * zc[1]; Ji[16] = 2 * zc[0];
    double* B = d.Bo + 18 * (size_t)o;
    const bool agg = add_diag_warp(d, a, Ji, e, w);
    for (int r = 0; r < 6; r++) {
      if (!agg) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Ji[6 * k + r] * e[k];
        atomicAdd(&d.bs[6 * a + r], -w * s);
      }
      for (int q = 0; q < 3; q++) {
        double h = 0;
        for (int k = 0; k < 3; k++) h += Ji[6 * k + r] * Xa.R[3 * q + k];   // J_point[k][q] = R[q][k]
        B[3 * r + q] = w * h;
      }
    }
    if (!agg) add_block(d, a, a, Ji, Ji, 3, w);
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += Xa.R[3 * r + k] * e[k];
      atomicAdd(&d.bl[3 * (size_t)l + r], -w * s);
    }
    atomicAdd(&d.hl[l], w);
  } else {
    const int t = i - 1 - d.NE - d.NO, a = d.t1[t], c = d.t2[t], hv = d.th[t];
    const Pose H = X[hv];   // private copy, see above
    double q[3], e[3];
    tern_error(H, P + 3 * (size_t)a, P + 3 * (size_t)c, q, e);
    double hw;
    huber_rho((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * d.infoT, d.dT, &hw);
    const double w = hw * d.infoT;
    // J1 = I ; J2 = -R_H^T (J2[k][r] = -H.R[r][k]) ; JH = [I | -[q]x]
    double JH[18];
    for (int k = 0; k < 18; k++) JH[k] = 0;
    JH[0] = 1; JH[7] = 1; JH[14] = 1;
    JH[4] = q[2];  JH[5] = -q[1];
    JH[9] = -q[2]; JH[11] = q[0];
    JH[15] = q[1]; JH[16] = -q[0];
    for (int r = 0; r < 3; r++) {
      atomicAdd(&d.bl[3 * (size_t)a + r], -w * e[r]);
      double s = 0;
      for (int k = 0; k < 3; k++) s += -H.R[3 * r + k] * e[k];
      atomicAdd(&d.bl[3 * (size_t)c + r], -w * s);
    }
    atomicAdd(&d.hl[a], w);
    atomicAdd(&d.hl[c], w);
    double* b1 = d.B1 + 18 * (size_t)t;
    double* b2 = d.B2 + 18 * (size_t)t;
    const bool agg = add_diag_warp(d, hv, JH, e, w);
    for (int r = 0; r < 6; r++) {
      if (!agg) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += JH[6 * k + r] * e[k];
        atomicAdd(&d.bs[6 * hv + r], -w * s);
      }
      for (int qq = 0; qq < 3; qq++) {
        b1[3 * r + qq] = w * JH[6 * qq + r];
        double h = 0;
        for (int k = 0; k < 3; k++) h += JH[6 * k + r] * (-H.R[3 * qq + k]);
        b2[3 * r + qq] = w * h;
      }
    }
    if (!agg) add_block(d, hv, hv, JH, JH, 3, w);
    double* O = d.Ot + 9 * (size_t)t;
    for (int r = 0; r < 3; r++)
      for (int qq = 0; qq < 3; qq++) O[3 * r + qq] = w * (-H.R[3 * qq + r]);   // J1^T J2 = J2
  }
}

__global__ void __launch_bounds__(256) fba_maxdiag_kernel(FbaDev d) {
  double m = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n + d.NP; i += gridDim.x * blockDim.x)
    m = fmax(m, fabs(i < d.n ? (d.pcg ? d.Hd[36 * (size_t)(i / 6) + 7 * (i % 6)] : d.Hss[(size_t)i * d.n + i]) : d.hl[i - d.n]));
  __shared__ double sm[32];
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = fmax(m, sm[w]);
    d.partial[blockIdx.x] = fmax(m, sm[0]);
  }
}
__global__ void fba_final_max_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
  double m = 0;
  for (int i = threadIdx.x; i < nblocks; i += 32) m = fmax(m, partial[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (threadIdx.x == 0) *out = m;
}

// S = Hss (lower blocks) + lambda I ; bp = bs
__global__ void __launch_bounds__(256) fba_prepare_trial_kernel(FbaDev d, double lambda) {
  const size_t nn = (size_t)d.n * d.n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / d.n, c = i % d.n;
    d.S[i] = d.Hss[i] + (r == c ? lambda : 0.0);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) d.bp[i] = d.bs[i];
}

// ---------------------------------------------------------------------------------------------------------
// point elimination: one warp per chain
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const double* cpl_block(const FbaDev& d, int ci) {
  const int k = d.cpl_kind[ci], e = d.cpl_e[ci];
  return k == 0 ? d.Bo + 18 * (size_t)e : (k == 1 ? d.B1 + 18 * (size_t)e : d.B2 + 18 * (size_t)e);
}

__global__ void __launch_bounds__(32 * FBA_CHAIN_WARPS) fba_chain_kernel(FbaDev d, double lambda) {
  __shared__ double sY[FBA_CHAIN_WARPS][FBA_MAX_ACTIVE * 18];
  __shared__ int sV[FBA_CHAIN_WARPS][FBA_MAX_ACTIVE];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * FBA_CHAIN_WARPS + warp;
  if (c >= d.NCH) return;
  double* Y = sY[warp];
  int* V = sV[warp];
  const int k0 = d.chain_start[c], k1 = d.chain_start[c + 1];
  const int n = d.n;
  int nact = 0;
  double Lp[6] = {1, 0, 1, 0, 0, 1};   // previous diagonal factor (l00, l10, l11, l20, l21, l22)
  double cprev[3] = {0, 0, 0};
  for (int k = k0; k < k1; k++) {
    const int p = d.chain_pts[k];
    // ---- every lane computes the 3x3 pivot redundantly (a handful of flops, no divergence)
    const double hd = d.hl[p] + lambda;
    double D[9] = {hd, 0, 0, 0, hd, 0, 0, 0, hd};
    double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // L_{k,k-1} = T_{k,k-1} Lp^-T,  T_{k,k-1} = O^T
    if (k > k0) {
      const double* O = d.Ot + 9 * (size_t)d.chain_link[k];
      for (int r = 0; r < 3; r++) {
        const double a0 = O[r], a1 = O[3 + r], a2 = O[6 + r];
        const double y0 = a0 / Lp[0];
        const double y1 = (a1 - y0 * Lp[1]) / Lp[2];
        const double y2 = (a2 - y0 * Lp[3] - y1 * Lp[4]) / Lp[5];
        M[3 * r] = y0; M[3 * r + 1] = y1; M[3 * r + 2] = y2;
      }
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) D[3 * r + q] -= M[3 * r] * M[3 * q] + M[3 * r + 1] * M[3 * q + 1] + M[3 * r + 2] * M[3 * q + 2];
    }
    double L[6];
    bool bad = false;
    {
      double v = D[0];
      if (!(v > 0)) bad = true;
      L[0] = sqrt(v);
      L[1] = D[3] / L[0];
      L[3] = D[6] / L[0];
      v = D[4] - L[1] * L[1];
      if (!(v > 0)) bad = true;
      L[2] = sqrt(v);
      L[4] = (D[7] - L[3] * L[1]) / L[2];
      v = D[8] - L[3] * L[3] - L[4] * L[4];
      if (!(v > 0)) bad = true;
      L[5] = sqrt(v);
    }
    if (bad) { if (lane == 0) *d.fail = 1; return; }
    if (lane == 0) {
      for (int q = 0; q < 6; q++) d.Lkk[6 * (size_t)p + q] = L[q];
      for (int q = 0; q < 9; q++) d.Lk1[9 * (size_t)p + q] = M[q];
    }
    // ---- Y_v <- -(Y_v M^T) for the active vertices (lanes over (vertex, row))
    for (int idx = lane; idx < nact * 6; idx += 32) {
      double* y = Y + 3 * idx;
      const double y0 = y[0], y1 = y[1], y2 = y[2];
      y[0] = -(y0 * M[0] + y1 * M[1] + y2 * M[2]);
      y[1] = -(y0 * M[3] + y1 * M[4] + y2 * M[5]);
      y[2] = -(y0 * M[6] + y1 * M[7] + y2 * M[8]);
    }
    __syncwarp();
    // ---- + B for the vertices coupled to this point (new vertices join the active list)
    const int c0 = d.cpl_start[p], c1 = d.cpl_start[p + 1];
    for (int ci = c0; ci < c1; ci++) {
      const int v = d.cpl_v[ci];
      int slot = -1;
      if (k > k0) {   // the first point of a chain only meets new vertices
        for (int a0 = 0; a0 < nact; a0 += 32) {
          const int a = a0 + lane;
          const unsigned hit = __ballot_sync(0xffffffffu, a < nact && V[a] == v);
          if (hit) { slot = a0 + __ffs(hit) - 1; break; }
        }
      }
      if (slot < 0) {
        if (nact >= FBA_MAX_ACTIVE) { if (lane == 0) *d.fail = 2; return; }
        slot = nact++;
        if (lane == 0) V[slot] = v;
        if (lane < 18) Y[18 * slot + lane] = 0.0;
        __syncwarp();
      }
      const double* B = cpl_block(d, ci);
      if (lane < 18) Y[18 * slot + lane] += B[lane];
      __syncwarp();
    }
    // ---- Y_v <- Y_v L^-T ; c_k = L^-1 (b_k - M c_{k-1})
    for (int idx = lane; idx < nact * 6; idx += 32) {
      double* y = Y + 3 * idx;
      const double y0 = y[0] / L[0];
      const double y1 = (y[1] - y0 * L[1]) / L[2];
      const double y2 = (y[2] - y0 * L[3] - y1 * L[4]) / L[5];
      y[0] = y0; y[1] = y1; y[2] = y2;
    }
    __syncwarp();
    double ck[3];
    {
      const double* b = d.bl + 3 * (size_t)p;
      const double r0 = b[0] - (M[0] * cprev[0] + M[1] * cprev[1] + M[2] * cprev[2]);
      const double r1 = b[1] - (M[3] * cprev[0] + M[4] * cprev[1] + M[5] * cprev[2]);
      const double r2 = b[2] - (M[6] * cprev[0] + M[7] * cprev[1] + M[8] * cprev[2]);
      ck[0] = r0 / L[0];
      ck[1] = (r1 - L[1] * ck[0]) / L[2];
      ck[2] = (r2 - L[3] * ck[0] - L[4] * ck[1]) / L[5];
    }
    // ---- Schur terms: bp[v] -= Y_v c_k ; S[u][v] -= Y_u Y_v^T for the active pairs with u >= v (lower blocks)
    for (int idx = lane; idx < nact * 6; idx += 32) {
      const double* y = Y + 3 * idx;
      atomicAdd(&d.bp[6 * V[idx / 6] + idx % 6], -(y[0] * ck[0] + y[1] * ck[1] + y[2] * ck[2]));
    }
    // one lane per vertex pair: the two 6x3 blocks go to registers once, then 36 accumulations (no per-entry index math)
    const int npair = nact * nact;
    for (int pr = lane; pr < npair; pr += 32) {
      const int a = pr / nact, b = pr - a * nact;
      const int va = V[a], vb2 = V[b];
      if (va < vb2) continue;
      double ya[18], yb[18];
#pragma unroll
      for (int q = 0; q < 18; q++) { ya[q] = Y[18 * a + q]; yb[q] = Y[18 * b + q]; }
      double* Sb = d.S + (size_t)(6 * va) * n + 6 * vb2;
#pragma unroll
      for (int r = 0; r < 6; r++)
#pragma unroll
        for (int q = 0; q < 6; q++)
          atomicAdd(Sb + (size_t)r * n + q, -(ya[3 * r] * yb[3 * q] + ya[3 * r + 1] * yb[3 * q + 1] + ya[3 * r + 2] * yb[3 * q + 2]));
    }
    __syncwarp();
    for (int q = 0; q < 6; q++) Lp[q] = L[q];
    cprev[0] = ck[0]; cprev[1] = ck[1]; cprev[2] = ck[2];
  }
}

// ---------------------------------------------------------------------------------------------------------
// dense Cholesky of the SE3 block (lower, in place, row-major), tiled right-looking
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int n, int j0, int jb, int* __restrict__ fail) {
  __shared__ double T[FBA_TB][FBA_TB + 1];
  for (int i = threadIdx.x; i < jb * jb; i += blockDim.x) T[i / jb][i % jb] = A[(size_t)(j0 + i / jb) * n + j0 + i % jb];
  __syncthreads();
  for (int j = 0; j < jb; j++) {
    if (threadIdx.x == 0) {
      const double v = T[j][j];
      if (!(v > 0)) { *fail = 3; T[j][j] = 1.0; }
      else T[j][j] = sqrt(v);
    }
    __syncthreads();
    const double ljj = T[j][j];
    for (int i = j + 1 + threadIdx.x; i < jb; i += blockDim.x) T[i][j] /= ljj;
    __syncthreads();
    for (int idx = threadIdx.x; idx < (jb - j - 1) * (jb - j - 1); idx += blockDim.x) {
      const int r = j + 1 + idx / (jb - j - 1), c = j + 1 + idx % (jb - j - 1);
      if (c <= r) T[r][c] -= T[r][j] * T[c][j];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < jb * jb; i += blockDim.x) {
    const int r = i / jb, c = i % jb;
    if (c <= r) A[(size_t)(j0 + r) * n + j0 + c] = T[r][c];
  }
}

// rows below the diagonal tile: A[i][j0..j0+jb) <- A[i][..] L^-T  (one thread per row)
// (`rows` = rows below the panel that can be non-zero: the factor keeps the envelope of the matrix, see fba_run)
__global__ void __launch_bounds__(128) chol_panel_kernel(double* __restrict__ A, int n, int j0, int jb, int rows) {
  __shared__ double T[FBA_TB][FBA_TB + 1];
  for (int i = threadIdx.x; i < jb * jb; i += blockDim.x) T[i / jb][i % jb] = A[(size_t)(j0 + i / jb) * n + j0 + i % jb];
  __syncthreads();
  const int i = j0 + jb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j0 + jb + rows) return;
  double* row = A + (size_t)i * n + j0;
  double y[FBA_TB];
  for (int c = 0; c < jb; c++) {
    double s = row[c];
    for (int k = 0; k < c; k++) s -= y[k] * T[c][k];
    y[c] = s / T[c][c];
  }
  for (int c = 0; c < jb; c++) row[c] = y[c];
}

// trailing update: A[i][k] -= sum_m P[i][m] P[k][m] for i >= k >= j0+jb, 32x32 tiles (lower tiles only)
__global__ void __launch_bounds__(256) chol_update_kernel(double* __restrict__ A, int n_full, int j0, int jb, int rows) {
  if (blockIdx.x > blockIdx.y) return;   // tile column <= tile row
  __shared__ double Pi[32][FBA_TB + 1], Pk[32][FBA_TB + 1];
  const int base = j0 + jb, i0 = base + blockIdx.y * 32, k0 = base + blockIdx.x * 32;
  const size_t n = (size_t)n_full;
  const int lim = base + rows;            // rows / columns beyond the envelope are untouched
  for (int idx = threadIdx.x; idx < 32 * jb; idx += blockDim.x) {
    const int r = idx / jb, m = idx % jb;
    Pi[r][m] = (i0 + r < lim) ? A[(size_t)(i0 + r) * n + j0 + m] : 0.0;
    Pk[r][m] = (k0 + r < lim) ? A[(size_t)(k0 + r) * n + j0 + m] : 0.0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * 32; idx += blockDim.x) {
    const int r = idx / 32, c = idx % 32;
    const int i = i0 + r, k = k0 + c;
    if (i >= lim || k >= lim || k > i) continue;
    double s = 0;
    for (int m = 0; m < jb; m++) s += Pi[r][m] * Pk[c][m];
    A[(size_t)i * n + k] -= s;
  }
}

// L y = b (column-oriented), L^T x = y (row-oriented); one CTA, the vector stays in global memory
// (`band`: L[i][j] = 0 for i - j > band)
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double* __restrict__ L, int n, double* __restrict__ b, int band) {
  for (int j = 0; j < n; j++) {
    __shared__ double xj;
    if (threadIdx.x == 0) { xj = b[j] / L[(size_t)j * n + j]; b[j] = xj; }
    __syncthreads();
    const double v = xj;
    const int hi = min(n, j + band + 1);
    for (int i = j + 1 + threadIdx.x; i < hi; i += blockDim.x) b[i] -= L[(size_t)i * n + j] * v;
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; j--) {
    __shared__ double xj2;
    if (threadIdx.x == 0) { xj2 = b[j] / L[(size_t)j * n + j]; b[j] = xj2; }
    __syncthreads();
    const double v = xj2;
    for (int k = max(0, j - band) + threadIdx.x; k < j; k += blockDim.x) b[k] -= L[(size_t)j * n + k] * v;
    __syncthreads();
  }
}

// points: T x_l = b_l - B^T x_s, chain by chain (one thread per chain: forward with L, backward with L^T)
__global__ void __launch_bounds__(128) fba_backsub_kernel(FbaDev d) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.NCH) return;
  const int k0 = d.chain_start[c], k1 = d.chain_start[c + 1];
  double* xl = d.x + d.n;
  double yp[3] = {0, 0, 0};
  for (int k = k0; k < k1; k++) {
    const int p = d.chain_pts[k];
    double r[3] = {d.bl[3 * (size_t)p], d.bl[3 * (size_t)p + 1], d.bl[3 * (size_t)p + 2]};
    for (int ci = d.cpl_start[p]; ci < d.cpl_start[p + 1]; ci++) {
      const double* B = cpl_block(d, ci);
      const double* xv = d.x + 6 * d.cpl_v[ci];
      for (int rr = 0; rr < 6; rr++) {
        const double xr = xv[rr];
        r[0] -= B[3 * rr] * xr; r[1] -= B[3 * rr + 1] * xr; r[2] -= B[3 * rr + 2] * xr;
      }
    }
    const double* M = d.Lk1 + 9 * (size_t)p;
    const double* L = d.Lkk + 6 * (size_t)p;
    if (k > k0)
      for (int q = 0; q < 3; q++) r[q] -= M[3 * q] * yp[0] + M[3 * q + 1] * yp[1] + M[3 * q + 2] * yp[2];
    yp[0] = r[0] / L[0];
    yp[1] = (r[1] - L[1] * yp[0]) / L[2];
    yp[2] = (r[2] - L[3] * yp[0] - L[4] * yp[1]) / L[5];
    xl[3 * (size_t)p] = yp[0]; xl[3 * (size_t)p + 1] = yp[1]; xl[3 * (size_t)p + 2] = yp[2];
  }
  double xn[3] = {0, 0, 0};
  for (int k = k1 - 1; k >= k0; k--) {
    const int p = d.chain_pts[k];
    const double* L = d.Lkk + 6 * (size_t)p;
    double r[3] = {xl[3 * (size_t)p], xl[3 * (size_t)p + 1], xl[3 * (size_t)p + 2]};
    if (k + 1 < k1) {
      const double* Mn = d.Lk1 + 9 * (size_t)d.chain_pts[k + 1];
      for (int q = 0; q < 3; q++) r[q] -= Mn[q] * xn[0] + Mn[3 + q] * xn[1] + Mn[6 + q] * xn[2];
    }
    xn[2] = r[2] / L[5];
    xn[1] = (r[1] - L[4] * xn[2]) / L[2];
    xn[0] = (r[0] - L[1] * xn[1] - L[3] * xn[2]) / L[0];
    xl[3 * (size_t)p] = xn[0]; xl[3 * (size_t)p + 1] = xn[1]; xl[3 * (size_t)p + 2] = xn[2];
  }
}


// ---------------------------------------------------------------------------------------------------------
// matrix-free path: (Hss + lambda I - B (T + lambda I)^-1 B^T) x = bs - B (T + lambda I)^-1 bl by preconditioned CG
// ---------------------------------------------------------------------------------------------------------
// Cholesky of every chain's block-tridiagonal pivot (thread per chain): Lkk, Lk1
__global__ void __launch_bounds__(128) pcg_chain_factor_kernel(FbaDev d, double lambda) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.NCH) return;
  double Lp[6] = {1, 0, 1, 0, 0, 1};
  const int k0 = d.chain_start[c], k1 = d.chain_start[c + 1];
  for (int k = k0; k < k1; k++) {
    const int p = d.chain_pts[k];
    const double hd = d.hl[p] + lambda;
    double D[9] = {hd, 0, 0, 0, hd, 0, 0, 0, hd};
    double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (k > k0) {
      const double* O = d.Ot + 9 * (size_t)d.chain_link[k];
      for (int r = 0; r < 3; r++) {
        const double a0 = O[r], a1 = O[3 + r], a2 = O[6 + r];
        const double y0 = a0 / Lp[0];
        const double y1 = (a1 - y0 * Lp[1]) / Lp[2];
        const double y2 = (a2 - y0 * Lp[3] - y1 * Lp[4]) / Lp[5];
        M[3 * r] = y0; M[3 * r + 1] = y1; M[3 * r + 2] = y2;
      }
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) D[3 * r + q] -= M[3 * r] * M[3 * q] + M[3 * r + 1] * M[3 * q + 1] + M[3 * r + 2] * M[3 * q + 2];
    }
    double L[6];
    double v = D[0];
    bool bad = !(v > 0);
    L[0] = sqrt(v); L[1] = D[3] / L[0]; L[3] = D[6] / L[0];
    v = D[4] - L[1] * L[1];
    bad |= !(v > 0);
    L[2] = sqrt(v); L[4] = (D[7] - L[3] * L[1]) / L[2];
    v = D[8] - L[3] * L[3] - L[4] * L[4];
    bad |= !(v > 0);
    L[5] = sqrt(v);
    if (bad) { *d.fail = 1; return; }
    for (int q = 0; q < 6; q++) { d.Lkk[6 * (size_t)p + q] = L[q]; Lp[q] = L[q]; }
    for (int q = 0; q < 9; q++) d.Lk1[9 * (size_t)p + q] = M[q];
  }
}

// in-place (T + lambda I) u = t for every chain (thread per chain), t / u = [NP][3]
__global__ void __launch_bounds__(128) pcg_chain_solve_kernel(FbaDev d, double* __restrict__ t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.NCH) return;
  const int k0 = d.chain_start[c], k1 = d.chain_start[c + 1];
  double yp[3] = {0, 0, 0};
  for (int k = k0; k < k1; k++) {
    const int p = d.chain_pts[k];
    const double* M = d.Lk1 + 9 * (size_t)p;
    const double* L = d.Lkk + 6 * (size_t)p;
    double r[3] = {t[3 * (size_t)p], t[3 * (size_t)p + 1], t[3 * (size_t)p + 2]};
    if (k > k0)
      for (int q = 0; q < 3; q++) r[q] -= M[3 * q] * yp[0] + M[3 * q + 1] * yp[1] + M[3 * q + 2] * yp[2];
    yp[0] = r[0] / L[0];
    yp[1] = (r[1] - L[1] * yp[0]) / L[2];
    yp[2] = (r[2] - L[3] * yp[0] - L[4] * yp[1]) / L[5];
    t[3 * (size_t)p] = yp[0]; t[3 * (size_t)p + 1] = yp[1]; t[3 * (size_t)p + 2] = yp[2];
  }
  double xn[3] = {0, 0, 0};
  for (int k = k1 - 1; k >= k0; k--) {
    const int p = d.chain_pts[k];
    const double* L = d.Lkk + 6 * (size_t)p;
    double r[3] = {t[3 * (size_t)p], t[3 * (size_t)p + 1], t[3 * (size_t)p + 2]};
    if (k + 1 < k1) {
      const double* Mn = d.Lk1 + 9 * (size_t)d.chain_pts[k + 1];
      for (int q = 0; q < 3; q++) r[q] -= Mn[q] * xn[0] + Mn[3 + q] * xn[1] + Mn[6 + q] * xn[2];
    }
    xn[2] = r[2] / L[5];
    xn[1] = (r[1] - L[4] * xn[2]) / L[2];
    xn[0] = (r[0] - L[1] * xn[1] - L[3] * xn[2]) / L[0];
    t[3 * (size_t)p] = xn[0]; t[3 * (size_t)p + 1] = xn[1]; t[3 * (size_t)p + 2] = xn[2];
  }
}

// t_k = sum over the couplings of point k of B^T v_vertex (thread per point); v == nullptr: t = bl
__global__ void __launch_bounds__(256) pcg_point_gather_kernel(FbaDev d, const double* __restrict__ v, double* __restrict__ t) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.NP) return;
  double r[3] = {0, 0, 0};
  if (!v) { r[0] = d.bl[3 * (size_t)p]; r[1] = d.bl[3 * (size_t)p + 1]; r[2] = d.bl[3 * (size_t)p + 2]; }
  else
    for (int ci = d.cpl_start[p]; ci < d.cpl_start[p + 1]; ci++) {
      const double* B = cpl_block(d, ci);
      const double* xv = v + 6 * d.cpl_v[ci];
      for (int rr = 0; rr < 6; rr++) {
        const double xr = xv[rr];
        r[0] += B[3 * rr] * xr; r[1] += B[3 * rr + 1] * xr; r[2] += B[3 * rr + 2] * xr;
      }
    }
  t[3 * (size_t)p] = r[0]; t[3 * (size_t)p + 1] = r[1]; t[3 * (size_t)p + 2] = r[2];
}

// warp per SE3 vertex.  mode 0: out_v = (Hd_v + lambda) in_v + sum_e6 He in_other - sum_cpl B u_pt, part[v] = in_v . out_v
//                       mode 1: out_v = bs_v - sum_cpl B u_pt   (reduced right-hand side)
__global__ void __launch_bounds__(128) pcg_vertex_kernel(FbaDev d, double lambda, int mode, const double* __restrict__ in,
                                                          const double* __restrict__ u, double* __restrict__ out, double* __restrict__ part) {
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (v >= d.NS) return;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int k = d.vc_start[v] + lane; k < d.vc_start[v + 1]; k += 32) {
    const int ci = d.vc_ci[k];
    const double* B = cpl_block(d, ci);
    const double* up = u + 3 * (size_t)d.cpl_pt[ci];
    const double u0 = up[0], u1 = up[1], u2 = up[2];
#pragma unroll
    for (int r = 0; r < 6; r++) acc[r] -= B[3 * r] * u0 + B[3 * r + 1] * u1 + B[3 * r + 2] * u2;
  }
#pragma unroll
  for (int r = 0; r < 6; r++)
    for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  if (lane != 0) return;
  if (mode == 1) {
    for (int r = 0; r < 6; r++) out[6 * (size_t)v + r] = d.bs[6 * (size_t)v + r] + acc[r];
    return;
  }
  const double* xv = in + 6 * (size_t)v;
  const double* Hv = d.Hd + 36 * (size_t)v;
  for (int r = 0; r < 6; r++) {
    double s = lambda * xv[r];
    for (int q = 0; q < 6; q++) s += Hv[6 * r + q] * xv[q];
    acc[r] += s;
  }
  for (int k = d.ve_start[v]; k < d.ve_start[v + 1]; k++) {
    const int e = d.ve_e[k];
    const double* Hb = d.He6 + 36 * (size_t)e;
    if (d.ve_side[k] == 0) {
      const double* xo = in + 6 * (size_t)d.e6j[e];
      for (int r = 0; r < 6; r++)
        for (int q = 0; q < 6; q++) acc[r] += Hb[6 * r + q] * xo[q];
    } else {
      const double* xo = in + 6 * (size_t)d.e6i[e];
      for (int r = 0; r < 6; r++)
        for (int q = 0; q < 6; q++) acc[r] += Hb[6 * q + r] * xo[q];
    }
  }
  double dot = 0;
  for (int r = 0; r < 6; r++) { out[6 * (size_t)v + r] = acc[r]; dot += xv[r] * acc[r]; }
  part[v] = dot;
}

// block-Jacobi preconditioner: M_v = Hd_v + lambda I - sum_cpl B B^T / (hl_pt + lambda) (exact diagonal block of the Schur
// complement for static points, point-diagonal approximation along dynamic chains); stored as its lower Cholesky factor
__global__ void __launch_bounds__(128) pcg_precond_kernel(FbaDev d, double lambda) {
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (v >= d.NS) return;
  double acc[21];
#pragma unroll
  for (int k = 0; k < 21; k++) acc[k] = 0;
  for (int k = d.vc_start[v] + lane; k < d.vc_start[v + 1]; k += 32) {
    const int ci = d.vc_ci[k];
    const double* B = cpl_block(d, ci);
    const double w = 1.0 / (d.hl[d.cpl_pt[ci]] + lambda);
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int q = 0; q <= r; q++) acc[idx++] -= w * (B[3 * r] * B[3 * q] + B[3 * r + 1] * B[3 * q + 1] + B[3 * r + 2] * B[3 * q + 2]);
  }
#pragma unroll
  for (int k = 0; k < 21; k++)
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
  if (lane != 0) return;
  double* A = d.cg_A + 36 * (size_t)v;
  int idx = 0;
  for (int r = 0; r < 6; r++)
    for (int q = 0; q <= r; q++) {
      const double a = d.Hd[36 * (size_t)v + 6 * r + q] + acc[idx++] + (r == q ? lambda : 0.0);
      A[6 * r + q] = a; A[6 * q + r] = a;
    }
}

// The preconditioner M = blockdiag(A_v) + the SE3-edge blocks.  Every SE3 vertex has at most one incoming and one outgoing
// SE3 edge (odometry: pose k-1 -> k; smoothness: an object's motion in consecutive frames), so M is block-tridiagonal along
// disjoint paths and is factored exactly without fill (thread per path).  It carries the stiff couplings of the system
// (information 1e4 / 1e3 against 1/80 for the points), which block-Jacobi alone leaves to CG.
__global__ void __launch_bounds__(64) pcg_path_factor_kernel(FbaDev d) {
  const int pth = blockIdx.x * blockDim.x + threadIdx.x;
  if (pth >= d.NPATH) return;
  double Lp[36];
  for (int k = d.path_start[pth]; k < d.path_start[pth + 1]; k++) {
    const int v = d.path_v[k], e = d.path_e[k];
    double D[36], W[36];
    for (int q = 0; q < 36; q++) { D[q] = d.cg_A[36 * (size_t)v + q]; W[q] = 0; }
    if (e >= 0) {
      // T_{k,k-1} = He6^T (He6: rows = e6i = previous vertex, columns = e6j = this vertex);  W = T_{k,k-1} Lp^-T
      const double* Hb = d.He6 + 36 * (size_t)e;
      for (int r = 0; r < 6; r++) {
        double y[6];
        for (int c = 0; c < 6; c++) {
          double s2 = Hb[6 * c + r];
          for (int m = 0; m < c; m++) s2 -= y[m] * Lp[6 * c + m];
          y[c] = s2 / Lp[6 * c + c];
        }
        for (int c = 0; c < 6; c++) W[6 * r + c] = y[c];
      }
      for (int r = 0; r < 6; r++)
        for (int c = 0; c <= r; c++) {
          double s2 = 0;
          for (int m = 0; m < 6; m++) s2 += W[6 * r + m] * W[6 * c + m];
          D[6 * r + c] -= s2;
        }
    }
    double Lf[36];
    bool bad = false;
    for (int j = 0; j < 6; j++) {
      double dd = D[6 * j + j];
      for (int m = 0; m < j; m++) dd -= Lf[6 * j + m] * Lf[6 * j + m];
      if (!(dd > 0)) { bad = true; dd = 1.0; }
      Lf[6 * j + j] = sqrt(dd);
      for (int i = j + 1; i < 6; i++) {
        double s2 = D[6 * i + j];
        for (int m = 0; m < j; m++) s2 -= Lf[6 * i + m] * Lf[6 * j + m];
        Lf[6 * i + j] = s2 / Lf[6 * j + j];
      }
      for (int i = 0; i < j; i++) Lf[6 * i + j] = 0;
    }
    if (bad) *d.fail = 1;
    int idx = 0;
    for (int r = 0; r < 6; r++)
      for (int c = 0; c <= r; c++) d.cg_M[21 * (size_t)v + idx++] = Lf[6 * r + c];
    for (int q = 0; q < 36; q++) { d.cg_W[36 * (size_t)v + q] = W[q]; Lp[q] = Lf[q]; }
  }
}

// z = M^-1 r along every path (forward with L / W, backward with their transposes) + the per-vertex partials r.z and r.r
__global__ void __launch_bounds__(64) pcg_path_solve_kernel(FbaDev d) {
  const int pth = blockIdx.x * blockDim.x + threadIdx.x;
  if (pth >= d.NPATH) return;
  const int k0 = d.path_start[pth], k1 = d.path_start[pth + 1];
  double yp[6] = {0, 0, 0, 0, 0, 0};
  for (int k = k0; k < k1; k++) {
    const int v = d.path_v[k];
    const double* L = d.cg_M + 21 * (size_t)v;
    const double* W = d.cg_W + 36 * (size_t)v;
    double r[6];
    for (int q = 0; q < 6; q++) r[q] = d.cg_r[6 * (size_t)v + q];
    if (k > k0)
      for (int q = 0; q < 6; q++)
        for (int m = 0; m < 6; m++) r[q] -= W[6 * q + m] * yp[m];
    int idx = 0;
    for (int i = 0; i < 6; i++) {
      double s2 = r[i];
      for (int m = 0; m < i; m++) s2 -= L[idx + m] * yp[m];
      yp[i] = s2 / L[idx + i];
      idx += i + 1;
    }
    for (int q = 0; q < 6; q++) d.cg_z[6 * (size_t)v + q] = yp[q];
  }
  double zn[6] = {0, 0, 0, 0, 0, 0};
  for (int k = k1 - 1; k >= k0; k--) {
    const int v = d.path_v[k];
    const double* L = d.cg_M + 21 * (size_t)v;
    double y[6];
    for (int q = 0; q < 6; q++) y[q] = d.cg_z[6 * (size_t)v + q];
    if (k + 1 < k1) {
      const double* Wn = d.cg_W + 36 * (size_t)d.path_v[k + 1];
      for (int q = 0; q < 6; q++)
        for (int m = 0; m < 6; m++) y[q] -= Wn[6 * m + q] * zn[m];
    }
    for (int i = 5; i >= 0; i--) {
      double s2 = y[i];
      for (int m = i + 1; m < 6; m++) s2 -= L[m * (m + 1) / 2 + i] * zn[m];
      zn[i] = s2 / L[i * (i + 1) / 2 + i];
    }
    double rz = 0, rr = 0;
    for (int q = 0; q < 6; q++) {
      const double rq = d.cg_r[6 * (size_t)v + q];
      d.cg_z[6 * (size_t)v + q] = zn[q];
      rz += rq * zn[q]; rr += rq * rq;
    }
    d.cg_part[v] = rz; d.cg_part[d.NS + v] = rr;
  }
}

// fixed-order sums of the per-vertex partials (one block); what = 0: pq -> alpha ; 1: rz_new, rr -> beta ; 2: init (rz, rr, bnorm2)
__global__ void __launch_bounds__(256) pcg_scalar_kernel(FbaDev d, int what) {
  __shared__ double sm[2][8];
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < d.NS; i += 256) { a += d.cg_part[i]; b += d.cg_part[d.NS + i]; }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = a; sm[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  a = 0; b = 0;
  for (int w = 0; w < 8; w++) { a += sm[0][w]; b += sm[1][w]; }
  double* sc = d.cg_s;
  if (what == 0) {
    sc[1] = a;
    if (!(a > 0)) { *d.fail = 1; sc[3] = 0; } else sc[3] = sc[0] / a;
  } else if (what == 1) {
    sc[5] = a; sc[2] = b;
    sc[4] = sc[0] != 0 ? a / sc[0] : 0.0;
    sc[0] = a;
  } else {
    sc[0] = a; sc[2] = b; sc[6] = b;
  }
}

// x += alpha p ; r -= alpha q   (the preconditioner solve and the dot products follow in pcg_path_solve_kernel)
__global__ void __launch_bounds__(256) pcg_step1_kernel(FbaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n) return;
  const double alpha = d.cg_s[3];
  d.x[i] += alpha * d.cg_p[i];
  d.cg_r[i] -= alpha * d.cg_q[i];
}
// p = z + beta p
__global__ void __launch_bounds__(256) pcg_step2_kernel(FbaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.n) return;
  d.cg_p[i] = d.cg_z[i] + d.cg_s[4] * d.cg_p[i];
}
// start: x = 0, p = 0 (then p = z + 0 * p), beta = 0
__global__ void __launch_bounds__(256) pcg_init_kernel(FbaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) d.cg_s[4] = 0.0;
  if (i >= d.n) return;
  d.x[i] = 0.0;
  d.cg_p[i] = 0.0;
}

// a failed factorisation leaves x = b (linear_solver_csparse.h:126-133)
__global__ void __launch_bounds__(256) fba_x_from_b_kernel(FbaDev d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n + 3 * d.NP; i += gridDim.x * blockDim.x)
    d.x[i] = i < d.n ? d.bs[i] : d.bl[i - d.n];
}
__global__ void __launch_bounds__(256) fba_copy_bp_kernel(FbaDev d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += gridDim.x * blockDim.x) d.x[i] = d.bp[i];
}

// trial state = current (+) x ; partial sums of x^T (lambda x + b)
__global__ void __launch_bounds__(256) fba_update_kernel(FbaDev d, int cur, double lambda) {
  double acc = 0;
  const int total = d.NS + d.NP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (i < d.NS) {
      const double* u = d.x + 6 * (size_t)i;
      vb::pose_oplus(d.X[cur][i], u, d.X[cur ^ 1][i]);
      for (int k = 0; k < 6; k++) acc += u[k] * (lambda * u[k] + d.bs[6 * (size_t)i + k]);
    } else {
      const int l = i - d.NS;
      const double* u = d.x + d.n + 3 * (size_t)l;
      for (int k = 0; k < 3; k++) {
        d.P[cur ^ 1][3 * (size_t)l + k] = d.P[cur][3 * (size_t)l + k] + u[k];
        acc += u[k] * (lambda * u[k] + d.bl[3 * (size_t)l + k]);
      }
    }
  }
  block_sum_store(acc, d.partial);
}

// One device allocation per solve, carved by a bump allocator: cudaMalloc / cudaFree are expensive (and synchronise peers)
// once NCCL has enabled peer access, so the solver makes exactly one of each.  Pass 1 (base == nullptr) only adds up sizes.
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <class T>
  T* take(size_t count) {
    const size_t bytes = (std::max<size_t>(sizeof(T) * count, 16) + 255) & ~(size_t)255;
    T* p = base ? (T*)(base + off) : nullptr;
    off += bytes;
    return p;
  }
};
template <class T>
int dev_upload(vido_ctx* ctx, Arena& A, const T* src, size_t count, T** out) {
  *out = A.take<T>(count);
  if (A.base && count) VIDO_CUDA(cudaMemcpyAsync(*out, src, sizeof(T) * count, cudaMemcpyHostToDevice, ctx->stream));
  return VIDO_OK;
}
template <class T>
int dev_alloc(vido_ctx* ctx, Arena& A, size_t count, T** out) {
  *out = A.take<T>(count);
  if (A.base) VIDO_CUDA(cudaMemsetAsync(*out, 0, std::max<size_t>(sizeof(T) * count, 16), ctx->stream));
  return VIDO_OK;
}

}  // namespace

void vido_fba_default_params_impl(vido_fba_problem* p) {
  p->max_iterations = 300;
  p->sigma2_cam = 0.0001f; p->sigma2_3d_sta = 80.f; p->sigma2_3d_dyn = 80.f; p->sigma2_obj = 100.f; p->sigma2_smooth = 0.001f;
  p->huber_cam = 0.01f; p->huber_obj = 0.01f; p->huber_3d = 0.01f;
  p->gain_threshold = 1e-4f;
  p->prior_info = 100000.f;
  p->solver = 0;
}

static int fba_run(vido_ctx* ctx, vido_fba_problem* p, vido_lm_stats* st, char** arena_base) {
  const bool tdbg = getenv("VIDO_FBA_TIMING") != nullptr;   // debug: host wall time of the phases of a solve
  auto tnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tt0 = tnow();
  double tt_alloc0 = 0, tt_alloc1 = 0, tt_loop0 = 0;
  cudaStream_t s = ctx->stream;
  FbaDev d;
  memset(&d, 0, sizeof d);
  d.NS = p->n_poses + p->n_motions; d.NP = p->n_points; d.NO = p->n_obs; d.NE = p->n_e6; d.NT = p->n_tern;
  d.n = 6 * d.NS;
  if (st) { st->iterations = -1; st->n_records = 0; st->total_trials = 0; }
  if (d.NS == 0) return VIDO_OK;
  for (int o = 0; o < d.NO; o++)
    if (p->obs_se3[o] < 0 || p->obs_se3[o] >= d.NS || p->obs_point[o] < 0 || p->obs_point[o] >= d.NP) { ctx->err = "FullBatch: observation index out of range"; return VIDO_ERR_ARG; }
  for (int t = 0; t < d.NT; t++)
    if (p->tern_p1[t] < 0 || p->tern_p1[t] >= d.NP || p->tern_p2[t] < 0 || p->tern_p2[t] >= d.NP || p->tern_h[t] < 0 || p->tern_h[t] >= d.NS) { ctx->err = "FullBatch: ternary edge index out of range"; return VIDO_ERR_ARG; }
  for (int e = 0; e < d.NE; e++)
    if (p->e6_i[e] < 0 || p->e6_i[e] >= d.NS || p->e6_j[e] < 0 || p->e6_j[e] >= d.NS) { ctx->err = "FullBatch: SE3 edge index out of range"; return VIDO_ERR_ARG; }
  d.info6[0] = 1.0 / (double)p->sigma2_cam; d.info6[1] = 1.0 / (double)p->sigma2_smooth;
  d.d6[0] = (double)p->huber_cam; d.d6[1] = (double)p->huber_cam;
  d.info3[0] = 1.0 / (double)p->sigma2_3d_sta; d.info3[1] = 1.0 / (double)p->sigma2_3d_dyn;
  d.d3 = (double)p->huber_3d; d.infoT = 1.0 / (double)p->sigma2_obj; d.dT = (double)p->huber_obj; d.infoPrior = (double)p->prior_info;
  // ---- states and measurements (float -> double exactly like Converter::toSE3Quat)
  std::vector<Pose> X(d.NS), Zi(d.NE);
  for (int i = 0; i < d.NS; i++) vb::pose_from_f32(p->se3 + 16 * (size_t)i, X[i]);
  Pose I;
  for (int k = 0; k < 9; k++) I.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
  I.t[0] = I.t[1] = I.t[2] = 0;
  vb::pose_inv_mul(X[0], I, d.Zprior_inv);
  for (int e = 0; e < d.NE; e++) { Pose Z; vb::pose_from_f32(p->e6_meas + 16 * (size_t)e, Z); vb::pose_inv_mul(Z, I, Zi[e]); }
  std::vector<double> P(3 * (size_t)d.NP), meas(3 * (size_t)d.NO);
  for (size_t i = 0; i < P.size(); i++) P[i] = p->points[i];
  for (size_t i = 0; i < meas.size(); i++) meas[i] = p->obs_xyz[i];
  // ---- structure: chains (every point is p2 of at most one and p1 of at most one ternary edge), coupling lists
  std::vector<int> prev_t(d.NP, -1), next_t(d.NP, -1);
  for (int t = 0; t < d.NT; t++) {
    if (next_t[p->tern_p1[t]] != -1 || prev_t[p->tern_p2[t]] != -1) { ctx->err = "FullBatch: ternary edges do not form chains"; return VIDO_ERR_ARG; }
    next_t[p->tern_p1[t]] = t; prev_t[p->tern_p2[t]] = t;
  }
  std::vector<int> chain_start(1, 0), chain_pts, chain_link;
  chain_pts.reserve(d.NP); chain_link.reserve(d.NP);
  for (int q0 = 0; q0 < d.NP; q0++) {
    if (prev_t[q0] != -1) continue;
    int q = q0, lk = -1;
    while (true) {
      chain_pts.push_back(q); chain_link.push_back(lk);
      if (next_t[q] == -1) break;
      lk = next_t[q]; q = p->tern_p2[lk];
      if ((int)chain_pts.size() > d.NP) { ctx->err = "FullBatch: cyclic ternary chain"; return VIDO_ERR_ARG; }
    }
    chain_start.push_back((int)chain_pts.size());
  }
  if ((int)chain_pts.size() != d.NP) { ctx->err = "FullBatch: cyclic ternary chain"; return VIDO_ERR_ARG; }
  d.NCH = (int)chain_start.size() - 1;
  std::vector<int> cpl_start(d.NP + 1, 0);
  for (int o = 0; o < d.NO; o++) cpl_start[p->obs_point[o] + 1]++;
  for (int t = 0; t < d.NT; t++) { cpl_start[p->tern_p1[t] + 1]++; cpl_start[p->tern_p2[t] + 1]++; }
  for (int q = 0; q < d.NP; q++) cpl_start[q + 1] += cpl_start[q];
  const int ncpl = cpl_start[d.NP];
  std::vector<int> cpl_v(ncpl), cpl_kind(ncpl), cpl_e(ncpl), fill(cpl_start.begin(), cpl_start.end() - 1);
  for (int o = 0; o < d.NO; o++) { const int k = fill[p->obs_point[o]]++; cpl_v[k] = p->obs_se3[o]; cpl_kind[k] = 0; cpl_e[k] = o; }
  for (int t = 0; t < d.NT; t++) {
    int k = fill[p->tern_p1[t]]++; cpl_v[k] = p->tern_h[t]; cpl_kind[k] = 1; cpl_e[k] = t;
    k = fill[p->tern_p2[t]]++; cpl_v[k] = p->tern_h[t]; cpl_kind[k] = 2; cpl_e[k] = t;
  }
  // ---- frame order of the SE3 vertices.  The reduced system couples a vertex only with vertices of nearby frames (odometry,
  //      smoothness, the frames a tracklet spans), so with the object motions interleaved with the camera poses of their
  //      frame it is banded; the Cholesky factor keeps the envelope, and the tiled factorisation / triangular solves below
  //      only touch `band` rows below each panel.  perm maps the caller's vertex index to the solver's.
  std::vector<int> perm(d.NS), inv_perm(d.NS);
  {
    std::vector<int> key(d.NS, -1), pt_pose(d.NP, -1);
    for (int i = 0; i < p->n_poses; i++) key[i] = 2 * i;
    for (int o = 0; o < d.NO; o++) pt_pose[p->obs_point[o]] = std::max(pt_pose[p->obs_point[o]], (int)p->obs_se3[o]);
    for (int t = 0; t < d.NT; t++) {
      const int h = p->tern_h[t], pp = pt_pose[p->tern_p2[t]];
      if (h >= p->n_poses && pp >= 0 && pp < p->n_poses) key[h] = std::max(key[h], 2 * pp + 1);
    }
    for (int pass = 0; pass < 4; pass++)
      for (int e = 0; e < d.NE; e++) {
        const int a = p->e6_i[e], b2 = p->e6_j[e];
        if (key[a] >= 0 && key[b2] < 0) key[b2] = key[a] + 2;
        else if (key[b2] >= 0 && key[a] < 0) key[a] = std::max(key[b2] - 2, 1);
      }
    for (int v = 0; v < d.NS; v++) { if (key[v] < 0) key[v] = 2 * p->n_poses + 1; inv_perm[v] = v; }
    std::stable_sort(inv_perm.begin(), inv_perm.end(), [&](int x, int y) { return key[x] < key[y]; });
    for (int v = 0; v < d.NS; v++) perm[inv_perm[v]] = v;
  }
  std::vector<int> e6i(d.NE), e6j(d.NE), obs_se3(d.NO), tern_h(d.NT);
  for (int e = 0; e < d.NE; e++) { e6i[e] = perm[p->e6_i[e]]; e6j[e] = perm[p->e6_j[e]]; }
  for (int o = 0; o < d.NO; o++) obs_se3[o] = perm[p->obs_se3[o]];
  for (int t = 0; t < d.NT; t++) tern_h[t] = perm[p->tern_h[t]];
  for (int k = 0; k < ncpl; k++) cpl_v[k] = perm[cpl_v[k]];
  {
    std::vector<Pose> Xp(d.NS);
    for (int v = 0; v < d.NS; v++) Xp[perm[v]] = X[v];
    X.swap(Xp);
  }
  int band_v = 0;   // largest index distance between two coupled SE3 vertices
  for (int e = 0; e < d.NE; e++) band_v = std::max(band_v, std::abs(e6i[e] - e6j[e]));
  for (int ch = 0; ch < d.NCH; ch++) {
    int lo = d.NS, hi = -1;
    for (int k = chain_start[ch]; k < chain_start[ch + 1]; k++)
      for (int ci = cpl_start[chain_pts[k]]; ci < cpl_start[chain_pts[k] + 1]; ci++) { lo = std::min(lo, cpl_v[ci]); hi = std::max(hi, cpl_v[ci]); }
    if (hi >= 0) band_v = std::max(band_v, hi - lo);
  }
  const int band = std::min(d.n, 6 * (band_v + 1));
  // ---- solver: the explicit Schur complement of a chain couples all the SE3 vertices it meets (A^2 blocks, A^3 work), fine
  //      for tracklets of a few dozen frames; long dynamic tracklets and very large systems take the matrix-free path
  int max_active = 0;
  {
    std::vector<int> seen(d.NS, -1);
    for (int ch = 0; ch < d.NCH; ch++) {
      int cnt = 0;
      for (int k = chain_start[ch]; k < chain_start[ch + 1]; k++)
        for (int ci = cpl_start[chain_pts[k]]; ci < cpl_start[chain_pts[k] + 1]; ci++)
          if (seen[cpl_v[ci]] != ch) { seen[cpl_v[ci]] = ch; cnt++; }
      max_active = std::max(max_active, cnt);
    }
  }
  d.pcg = p->solver == 2 || (p->solver != 1 && (max_active > 96 || d.n > 16384)) ? 1 : 0;
  if (!d.pcg && max_active > FBA_MAX_ACTIVE) { ctx->err = "FullBatch: a tracklet couples more SE3 vertices than the chain kernel holds (use the matrix-free solver)"; return VIDO_ERR_CAPACITY; }
  std::vector<int> cpl_pt(ncpl), vc_start(d.NS + 1, 0), vc_ci(ncpl), ve_start(d.NS + 1, 0), ve_e(2 * (size_t)d.NE), ve_side(2 * (size_t)d.NE);
  for (int q = 0; q < d.NP; q++)
    for (int ci = cpl_start[q]; ci < cpl_start[q + 1]; ci++) { cpl_pt[ci] = q; vc_start[cpl_v[ci] + 1]++; }
  for (int v = 0; v < d.NS; v++) vc_start[v + 1] += vc_start[v];
  {
    std::vector<int> f2(vc_start.begin(), vc_start.end() - 1);
    for (int ci = 0; ci < ncpl; ci++) vc_ci[f2[cpl_v[ci]]++] = ci;
  }
  // paths of the SE3-edge graph for the preconditioner (every vertex: at most one incoming, one outgoing edge, no cycle;
  // a graph that violates this falls back to single-vertex paths = block-Jacobi)
  std::vector<int> path_start(1, 0), path_v, path_e;
  {
    std::vector<int> e_in(d.NS, -1), e_out(d.NS, -1);
    bool ok = true;
    for (int e = 0; e < d.NE && ok; e++) {
      if (e6i[e] == e6j[e] || e_out[e6i[e]] != -1 || e_in[e6j[e]] != -1) ok = false;
      else { e_out[e6i[e]] = e; e_in[e6j[e]] = e; }
    }
    if (ok) {
      for (int v = 0; v < d.NS; v++) {
        if (e_in[v] != -1) continue;
        int u = v, le = -1;
        while (true) {
          path_v.push_back(u); path_e.push_back(le);
          if (e_out[u] == -1) break;
          le = e_out[u]; u = e6j[le];
        }
        path_start.push_back((int)path_v.size());
      }
      if ((int)path_v.size() != d.NS) ok = false;   // a cycle
    }
    if (!ok) {
      path_start.assign(1, 0); path_v.clear(); path_e.clear();
      for (int v = 0; v < d.NS; v++) { path_v.push_back(v); path_e.push_back(-1); path_start.push_back(v + 1); }
    }
  }
  d.NPATH = (int)path_start.size() - 1;
  for (int e = 0; e < d.NE; e++) { ve_start[e6i[e] + 1]++; ve_start[e6j[e] + 1]++; }
  for (int v = 0; v < d.NS; v++) ve_start[v + 1] += ve_start[v];
  {
    std::vector<int> f3(ve_start.begin(), ve_start.end() - 1);
    for (int e = 0; e < d.NE; e++) {
      int k = f3[e6i[e]]++; ve_e[k] = e; ve_side[k] = 0;
      k = f3[e6j[e]]++; ve_e[k] = e; ve_side[k] = 1;
    }
  }
  // ---- device buffers: pass 0 sizes the arena, pass 1 carves and uploads (pageable host memory: the async copies are
  //      staged by the runtime before the call returns, so the host vectors may go out of scope afterwards)
  int rc;
  Arena pool;
  const size_t nn = d.pcg ? 16 : (size_t)d.n * d.n;   // the dense system only exists on the direct path
  const int RB = 1024;   // reduction blocks
  for (int pass = 0; pass < 2; pass++) {
  if (pass == 1) {
    void* base = nullptr;
    tt_alloc0 = tnow();
    // grow-only arena kept by the context: a fresh cudaMalloc of ~100 MB costs anything from 1 to several hundred ms
    // (measured: bimodal, driver-side), a second solve on the same context none
    const size_t need = pool.off + 256;
    if (ctx->fba_arena_bytes < need) {
      if (ctx->fba_arena) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->fba_arena); ctx->fba_arena = nullptr; ctx->fba_arena_bytes = 0; }
      VIDO_CUDA(cudaMalloc(&ctx->fba_arena, need));
      ctx->fba_arena_bytes = need;
    }
    base = ctx->fba_arena;
    tt_alloc1 = tnow();
    *arena_base = (char*)base;
    pool.base = (char*)base; pool.off = 0;
  }
#define UP(vec, field) if ((rc = dev_upload(ctx, pool, (vec).data(), (vec).size(), &field))) return rc
#define UPC(ptr, count, field) if ((rc = dev_upload(ctx, pool, ptr, (size_t)(count), &field))) return rc
  Pose* dX0; Pose* dX1; double* dP0; double* dP1;
  UP(X, dX0); UP(X, dX1); UP(P, dP0); UP(P, dP1);
  d.X[0] = dX0; d.X[1] = dX1; d.P[0] = dP0; d.P[1] = dP1;
  Pose* dZ; UP(Zi, dZ); d.Z6inv = dZ;
  int* tmp;
  UP(e6i, tmp); d.e6i = tmp; UP(e6j, tmp); d.e6j = tmp; UPC(p->e6_kind, d.NE, tmp); d.e6k = tmp;
  UP(obs_se3, tmp); d.os = tmp; UPC(p->obs_point, d.NO, tmp); d.op = tmp; UPC(p->obs_kind, d.NO, tmp); d.ok = tmp;
  double* dm; UP(meas, dm); d.meas = dm;
  UPC(p->tern_p1, d.NT, tmp); d.t1 = tmp; UPC(p->tern_p2, d.NT, tmp); d.t2 = tmp; UP(tern_h, tmp); d.th = tmp;
  UP(chain_start, tmp); d.chain_start = tmp; UP(chain_pts, tmp); d.chain_pts = tmp; UP(chain_link, tmp); d.chain_link = tmp;
  UP(cpl_start, tmp); d.cpl_start = tmp; UP(cpl_v, tmp); d.cpl_v = tmp; UP(cpl_kind, tmp); d.cpl_kind = tmp; UP(cpl_e, tmp); d.cpl_e = tmp;
  UP(cpl_pt, tmp); d.cpl_pt = tmp; UP(vc_start, tmp); d.vc_start = tmp; UP(vc_ci, tmp); d.vc_ci = tmp;
  UP(ve_start, tmp); d.ve_start = tmp; UP(ve_e, tmp); d.ve_e = tmp; UP(ve_side, tmp); d.ve_side = tmp;
  UP(path_start, tmp); d.path_start = tmp; UP(path_v, tmp); d.path_v = tmp; UP(path_e, tmp); d.path_e = tmp;
#undef UP
#undef UPC
#define AL(field, count) if ((rc = dev_alloc(ctx, pool, (size_t)(count), &field))) return rc
  AL(d.Hss, nn); AL(d.bs, d.n); AL(d.hl, d.NP); AL(d.bl, 3 * (size_t)d.NP);
  AL(d.Bo, 18 * (size_t)d.NO); AL(d.B1, 18 * (size_t)d.NT); AL(d.B2, 18 * (size_t)d.NT); AL(d.Ot, 9 * (size_t)d.NT);
  AL(d.S, nn); AL(d.bp, d.n); AL(d.x, d.n + 3 * (size_t)d.NP);
  AL(d.Lkk, 6 * (size_t)d.NP); AL(d.Lk1, 9 * (size_t)d.NP);
  AL(d.partial, RB + 8); AL(d.fail, 4);
  if (d.pcg) {
    AL(d.Hd, 36 * (size_t)d.NS); AL(d.He6, 36 * (size_t)d.NE);
    AL(d.cg_r, d.n); AL(d.cg_z, d.n); AL(d.cg_p, d.n); AL(d.cg_q, d.n); AL(d.cg_t, 3 * (size_t)d.NP);
    AL(d.cg_M, 21 * (size_t)d.NS); AL(d.cg_A, 36 * (size_t)d.NS); AL(d.cg_W, 36 * (size_t)d.NS);
    AL(d.cg_part, 2 * (size_t)d.NS); AL(d.cg_s, 8);
  }
#undef AL
  }  // pass
  double* d_scalar = d.partial + RB;   // [0] chi2, [1] scale, [2] maxdiag
  const int n_edges = 1 + d.NE + d.NO + d.NT;
  const int red_blocks = std::min(RB, (std::max(n_edges, d.NS + d.NP) + 255) / 256);

  auto launch_ok = [&]() -> int { VIDO_CUDA(cudaGetLastError()); return VIDO_OK; };
  auto chi2_of = [&](int sel, double* out) -> int {
    fba_chi2_kernel<<<red_blocks, 256, 0, s>>>(d, sel);
    fba_final_sum_kernel<<<1, 32, 0, s>>>(d.partial, red_blocks, d_scalar);
    ctx->launches += 2;
    if ((rc = launch_ok())) return rc;
    VIDO_CUDA(cudaMemcpyAsync(out, d_scalar, sizeof(double), cudaMemcpyDeviceToHost, s));
    VIDO_CUDA(cudaStreamSynchronize(s));
    return VIDO_OK;
  };

  LmCtl c;
  lm_reset(&c);
  long cg_total = 0;
  std::vector<LmRec> rec(VIDO_LM_REC);
  const double gain = (double)p->gain_threshold;
  tt_loop0 = tnow();
  for (int it = 0; it < p->max_iterations && !c.stop_flag && c.ok; it++) {
    if (it == 0) { if ((rc = chi2_of(c.cur, &c.currentChi))) return rc; }
    // ---- buildSystem at the current state
    if (d.pcg) VIDO_CUDA(cudaMemsetAsync(d.Hd, 0, sizeof(double) * 36 * (size_t)d.NS, s));
    else VIDO_CUDA(cudaMemsetAsync(d.Hss, 0, sizeof(double) * nn, s));
    VIDO_CUDA(cudaMemsetAsync(d.bs, 0, sizeof(double) * d.n, s));
    VIDO_CUDA(cudaMemsetAsync(d.hl, 0, sizeof(double) * std::max(d.NP, 1), s));
    VIDO_CUDA(cudaMemsetAsync(d.bl, 0, sizeof(double) * 3 * std::max<size_t>(d.NP, 1), s));
    fba_linearize_kernel<<<(n_edges + 127) / 128, 128, 0, s>>>(d, c.cur);
    ctx->launches++;
    double maxdiag = 0;
    if (it == 0) {
      fba_maxdiag_kernel<<<red_blocks, 256, 0, s>>>(d);
      fba_final_max_kernel<<<1, 32, 0, s>>>(d.partial, red_blocks, d_scalar + 2);
      ctx->launches += 2;
      if ((rc = launch_ok())) return rc;
      VIDO_CUDA(cudaMemcpyAsync(&maxdiag, d_scalar + 2, sizeof(double), cudaMemcpyDeviceToHost, s));
      VIDO_CUDA(cudaStreamSynchronize(s));
    }
    lm_begin_iteration(&c, it, maxdiag, -1.0);
    do {
      const double lambda = c.lambda;
      VIDO_CUDA(cudaMemsetAsync(d.fail, 0, sizeof(int) * 4, s));
      int fail = 0;
      if (!d.pcg) {
        fba_prepare_trial_kernel<<<148 * 4, 256, 0, s>>>(d, lambda);
        if (d.NCH > 0) fba_chain_kernel<<<(d.NCH + FBA_CHAIN_WARPS - 1) / FBA_CHAIN_WARPS, 32 * FBA_CHAIN_WARPS, 0, s>>>(d, lambda);
        ctx->launches += 2;
        for (int j0 = 0; j0 < d.n; j0 += FBA_TB) {
          const int jb = std::min(FBA_TB, d.n - j0), rest = std::min(d.n - j0 - jb, band);
          chol_diag_kernel<<<1, 256, 0, s>>>(d.S, d.n, j0, jb, d.fail);
          ctx->launches++;
          if (rest > 0) {
            chol_panel_kernel<<<(rest + 127) / 128, 128, 0, s>>>(d.S, d.n, j0, jb, rest);
            const int tiles = (rest + 31) / 32;
            chol_update_kernel<<<dim3(tiles, tiles), 256, 0, s>>>(d.S, d.n, j0, jb, rest);
            ctx->launches += 2;
          }
        }
        if ((rc = launch_ok())) return rc;
        VIDO_CUDA(cudaMemcpyAsync(&fail, d.fail, sizeof(int), cudaMemcpyDeviceToHost, s));
        VIDO_CUDA(cudaStreamSynchronize(s));
        if (fail == 2) { ctx->err = "FullBatch: a tracklet couples more SE3 vertices than the chain kernel holds"; return VIDO_ERR_CAPACITY; }
        if (fail) {
          fba_x_from_b_kernel<<<148, 256, 0, s>>>(d);
          ctx->launches++;
        } else {
          fba_copy_bp_kernel<<<148, 256, 0, s>>>(d);
          chol_solve_kernel<<<1, 1024, 0, s>>>(d.S, d.n, d.x, band);
          if (d.NCH > 0) fba_backsub_kernel<<<(d.NCH + 127) / 128, 128, 0, s>>>(d);
          ctx->launches += 3;
        }
      } else {
        // ---- matrix-free: factor the chains, preconditioner, reduced right-hand side, then CG on the SE3 unknowns
        const int vb = (d.NS + 3) / 4, cb = (d.NCH + 127) / 128, pb = (d.NP + 255) / 256;
        if (d.NCH > 0) pcg_chain_factor_kernel<<<cb, 128, 0, s>>>(d, lambda);
        const int ptb = (d.NPATH + 63) / 64, nb256 = (d.n + 255) / 256;
        pcg_precond_kernel<<<vb, 128, 0, s>>>(d, lambda);
        pcg_path_factor_kernel<<<ptb, 64, 0, s>>>(d);
        if (d.NP > 0) pcg_point_gather_kernel<<<pb, 256, 0, s>>>(d, nullptr, d.cg_t);
        if (d.NCH > 0) pcg_chain_solve_kernel<<<cb, 128, 0, s>>>(d, d.cg_t);
        pcg_vertex_kernel<<<vb, 128, 0, s>>>(d, lambda, 1, nullptr, d.cg_t, d.cg_r, d.cg_part);
        pcg_init_kernel<<<nb256, 256, 0, s>>>(d);
        pcg_path_solve_kernel<<<ptb, 64, 0, s>>>(d);
        pcg_scalar_kernel<<<1, 256, 0, s>>>(d, 2);
        pcg_step2_kernel<<<nb256, 256, 0, s>>>(d);
        ctx->launches += 9;
        if ((rc = launch_ok())) return rc;
        double sc[8];
        VIDO_CUDA(cudaMemcpyAsync(&fail, d.fail, sizeof(int), cudaMemcpyDeviceToHost, s));
        VIDO_CUDA(cudaMemcpyAsync(sc, d.cg_s, sizeof sc, cudaMemcpyDeviceToHost, s));
        VIDO_CUDA(cudaStreamSynchronize(s));
        const double bnorm2 = sc[6];
        const int max_cg = std::max(400, 6 * d.n);
        int cg_it = 0;
        while (!fail && bnorm2 > 0 && cg_it < max_cg) {
          for (int k = 0; k < 16; k++, cg_it++) {
            if (d.NP > 0) pcg_point_gather_kernel<<<pb, 256, 0, s>>>(d, d.cg_p, d.cg_t);
            if (d.NCH > 0) pcg_chain_solve_kernel<<<cb, 128, 0, s>>>(d, d.cg_t);
            pcg_vertex_kernel<<<vb, 128, 0, s>>>(d, lambda, 0, d.cg_p, d.cg_t, d.cg_q, d.cg_part);
            pcg_scalar_kernel<<<1, 256, 0, s>>>(d, 0);
            pcg_step1_kernel<<<nb256, 256, 0, s>>>(d);
            pcg_path_solve_kernel<<<ptb, 64, 0, s>>>(d);
            pcg_scalar_kernel<<<1, 256, 0, s>>>(d, 1);
            pcg_step2_kernel<<<nb256, 256, 0, s>>>(d);
            ctx->launches += 8;
          }
          if ((rc = launch_ok())) return rc;
          VIDO_CUDA(cudaMemcpyAsync(&fail, d.fail, sizeof(int), cudaMemcpyDeviceToHost, s));
          VIDO_CUDA(cudaMemcpyAsync(sc, d.cg_s, sizeof sc, cudaMemcpyDeviceToHost, s));
          VIDO_CUDA(cudaStreamSynchronize(s));
          if (!(sc[2] > 1e-26 * bnorm2)) break;   // |r| <= 1e-13 |b|
        }
        cg_total += cg_it;
        if (fail) {
          fba_x_from_b_kernel<<<148, 256, 0, s>>>(d);
          ctx->launches++;
        } else {
          if (d.NCH > 0) fba_backsub_kernel<<<(d.NCH + 127) / 128, 128, 0, s>>>(d);
          ctx->launches++;
        }
      }
      fba_update_kernel<<<red_blocks, 256, 0, s>>>(d, c.cur, lambda);
      fba_final_sum_kernel<<<1, 32, 0, s>>>(d.partial, red_blocks, d_scalar + 1);
      ctx->launches += 2;
      if ((rc = launch_ok())) return rc;
      double scale = 0, chi = 0;
      VIDO_CUDA(cudaMemcpyAsync(&scale, d_scalar + 1, sizeof(double), cudaMemcpyDeviceToHost, s));
      if ((rc = chi2_of(c.cur ^ 1, &chi))) return rc;
      lm_trial(&c, chi, scale, fail);
    } while (lm_more_trials(&c));
    lm_end_iteration(&c, it, gain, rec.data());
  }
  // ---- results: float32 like Converter::toCvSE3 / toCvMat (src/Optimizer.cc:2090-2176)
  const double tt_loop1 = tnow();
  VIDO_CUDA(cudaMemcpy(X.data(), d.X[c.cur], sizeof(Pose) * d.NS, cudaMemcpyDeviceToHost));
  VIDO_CUDA(cudaMemcpy(P.data(), d.P[c.cur], sizeof(double) * 3 * d.NP, cudaMemcpyDeviceToHost));
  if (tdbg)
    fprintf(stderr, "[fba] host ms: graph preparation %.1f, cudaMalloc %.1f (%.1f MB), upload + set-up %.1f, LM loop %.1f, download %.1f\n",
            tt_alloc0 - tt0, tt_alloc1 - tt_alloc0, (double)pool.off / 1e6, tt_loop0 - tt_alloc1, tt_loop1 - tt_loop0, tnow() - tt_loop1);
  for (int i = 0; i < d.NS; i++) vb::pose_to_f32(X[perm[i]], p->se3 + 16 * (size_t)i);
  for (size_t i = 0; i < P.size(); i++) p->points[i] = (float)P[i];
  if (st) {
    st->iterations = c.iterations; st->n_records = c.n_records; st->total_trials = c.total_trials;
    st->pad = d.pcg ? (int32_t)std::min<long>(cg_total, 0x7fffffff) : 0;   // CG iterations of the matrix-free path
    for (int k = 0; k < c.n_records && k < VIDO_LM_MAX_RECORDS; k++) {
      st->rec[k].chi2 = rec[k].chi2; st->rec[k].lambda = rec[k].lambda; st->rec[k].trials = rec[k].trials;
    }
  }
  return VIDO_OK;
}

int fba_solve_host(vido_ctx* ctx, vido_fba_problem* p, vido_lm_stats* st) {
  char* arena = nullptr;
  const int rc = fba_run(ctx, p, st, &arena);
  cudaStreamSynchronize(ctx->stream);
  (void)arena;   // the arena belongs to the context (freed by vido_destroy)
  return rc;
}
