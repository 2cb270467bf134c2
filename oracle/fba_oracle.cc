// fba_oracle.cc -- CPU restatement of the full-sequence graph optimisation.  TEST INFRASTRUCTURE ONLY.
//
//   Optimizer::FullBatchOptimization          src/Optimizer.cc:1235-2178 (graph :1318-1745, optimize(300) :1941, write-back :2090-2176)
//   LandmarkMotionTernaryEdge                 g2o/types/types_dyn_slam3d.cpp:53-85 (err = p1 - H^-1 p2; the Jacobian w.r.t. H is the
//                                             reference's own [I | -[H^-1 p2]x], kept as written)
//   EdgeSE3Prior (identity offset)            g2o/types/edge_se3_prior.cpp:89-102, isometry3d_gradients.h:265-325: with P = I the
//                                             error and Jacobian equal EdgeSE3's (Xi = I fixed, Xj = X)
//   BaseMultiEdge::constructQuadraticForm     g2o/core/base_multi_edge.hpp:36-48,171-225 (robust weight rho[1] on Omega and on the rhs)
//   EdgeSE3 / EdgeSE3PointXYZ / VertexSE3 / LM shell: see ba_oracle.cc, lm_oracle.h
// Linear solver: the reference factors the whole H (no marginalised vertex) with LinearSolverCSparse; CSparse/AMD is
// un-vendored, so the restatement fixes the elimination order "points first (each tracklet chain as one block-tridiagonal
// pivot), then the SE3 vertices (dense LL^T)".  Any elimination order solves the same system; a non-positive pivot is
// reported as a solver failure (x = b) like csparse_extension.cpp:112.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "g2o_math.h"
#include "lm_oracle.h"
#include "vido_oracle.h"

namespace vo {

void edge_se3(const Iso& Xi, const Iso& Xj, const Iso& Z, double err[6], double Ji[6][6], double Jj[6][6]);
void edge_se3_pointxyz(const Iso& X, const V3& p, const V3& z, double err[3], double Ji[3][6], double Jj[3][3]);
Iso se3_oplus(const Iso& X, const double u[6]);

// LandmarkMotionTernaryEdge::computeError / linearizeOplus
void edge_landmark_motion(const Iso& H, const V3& p1, const V3& p2, double err[3], double J2[3][3], double JH[3][6]) {
  const Iso Hi = inverse(H);
  const V3 q = apply(Hi, p2);
  err[0] = p1.x - q.x; err[1] = p1.y - q.y; err[2] = p1.z - q.z;
  if (!J2) return;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) J2[i][j] = -Hi.R.m[3 * i + j];
  memset(JH, 0, sizeof(double) * 18);
  JH[0][0] = JH[1][1] = JH[2][2] = 1;
  JH[0][4] = q.z;  JH[0][5] = -q.y;
  JH[1][3] = -q.z; JH[1][5] = q.x;
  JH[2][3] = q.y;  JH[2][4] = -q.x;
}

struct Cpl { int v; int kind; int e; };  // coupling of a point to an SE3 vertex: kind 0 obs edge, 1 ternary as p1, 2 ternary as p2

struct FBASystem {
  int NS = 0, NP = 0, NO = 0, NE = 0, NT = 0;
  std::vector<Iso> X, Z6;        // SE3 estimates, EdgeSE3 measurements
  Iso Zprior;
  std::vector<V3> pts, meas;
  std::vector<int> e6i, e6j, e6k, os, op, ok, t1, t2, th;
  double info6[2], d6[2], info3[2], d3, infoT, dT, infoPrior;
  // linearisation
  std::vector<double> e6err, oerr, terr;
  double perr[6];
  std::vector<double> Hss, bs;            // dense (6 NS)^2, rhs
  std::vector<double> hl, bl;             // point diagonal (scalar * I), rhs [NP][3]
  std::vector<double> Bo, B1, B2, Ot;     // [NO][18] pose-point; [NT][18] H-p1, H-p2; [NT][9] p1-p2
  std::vector<double> x;                  // [6 NS + 3 NP]
  // structure
  std::vector<int> chain_start, chain_pts, chain_link;  // points of every chain in order; link[k] = ternary edge between k-1 and k
  std::vector<int> cpl_start;
  std::vector<Cpl> cpl;
  std::vector<std::vector<Iso>> bkX;
  std::vector<std::vector<V3>> bkP;

  int num_vertices() const { return NS + NP; }
  int dim() const { return 6 * NS + 3 * NP; }

  void build_structure() {
    std::vector<int> prev_t(NP, -1), next_t(NP, -1);
    for (int t = 0; t < NT; t++) { next_t[t1[t]] = t; prev_t[t2[t]] = t; }
    chain_start.assign(1, 0);
    for (int p = 0; p < NP; p++) {
      if (prev_t[p] != -1) continue;  // not a chain head
      int q = p, lk = -1;
      while (true) {
        chain_pts.push_back(q); chain_link.push_back(lk);
        if (next_t[q] == -1) break;
        lk = next_t[q]; q = t2[lk];
      }
      chain_start.push_back((int)chain_pts.size());
    }
    std::vector<std::vector<Cpl>> L(NP);
    for (int o = 0; o < NO; o++) L[op[o]].push_back({os[o], 0, o});
    for (int t = 0; t < NT; t++) { L[t1[t]].push_back({th[t], 1, t}); L[t2[t]].push_back({th[t], 2, t}); }
    cpl_start.assign(NP + 1, 0);
    for (int p = 0; p < NP; p++) { cpl_start[p + 1] = cpl_start[p] + (int)L[p].size(); cpl.insert(cpl.end(), L[p].begin(), L[p].end()); }
  }

  void compute_errors() {
    edge_se3(iso_identity(), X[0], Zprior, perr, nullptr, nullptr);
    for (int e = 0; e < NE; e++) edge_se3(X[e6i[e]], X[e6j[e]], Z6[e], &e6err[6 * (size_t)e], nullptr, nullptr);
    for (int o = 0; o < NO; o++) edge_se3_pointxyz(X[os[o]], pts[op[o]], meas[o], &oerr[3 * (size_t)o], nullptr, nullptr);
    for (int t = 0; t < NT; t++) edge_landmark_motion(X[th[t]], pts[t1[t]], pts[t2[t]], &terr[3 * (size_t)t], nullptr, nullptr);
  }
  double robust_chi2() const {
    double chi = 0, rho[3];
    {
      double e = 0;
      for (int k = 0; k < 6; k++) e += perr[k] * perr[k];
      chi += e * infoPrior;  // no kernel
    }
    for (int e = 0; e < NE; e++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += e6err[6 * (size_t)e + k] * e6err[6 * (size_t)e + k];
      huber(s * info6[e6k[e]], d6[e6k[e]], rho);
      chi += rho[0];
    }
    for (int o = 0; o < NO; o++) {
      const double* e = &oerr[3 * (size_t)o];
      huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * info3[ok[o]], d3, rho);
      chi += rho[0];
    }
    for (int t = 0; t < NT; t++) {
      const double* e = &terr[3 * (size_t)t];
      huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * infoT, dT, rho);
      chi += rho[0];
    }
    return chi;
  }
  void add_se3_block(int a, int c, const double* Ja, const double* Jc, int rows, double w) {  // Hss[a][c] += w Ja^T Jc (J: rows x 6)
    const int n = 6 * NS;
    for (int r = 0; r < 6; r++)
      for (int q = 0; q < 6; q++) {
        double s = 0;
        for (int k = 0; k < rows; k++) s += Ja[6 * k + r] * Jc[6 * k + q];
        Hss[(size_t)(6 * a + r) * n + 6 * c + q] += w * s;
      }
  }
  void build_system() {
    std::fill(Hss.begin(), Hss.end(), 0.0);
    std::fill(bs.begin(), bs.end(), 0.0);
    std::fill(hl.begin(), hl.end(), 0.0);
    std::fill(bl.begin(), bl.end(), 0.0);
    double rho[3];
    {  // prior on SE3 vertex 0
      double e[6], Ji[6][6], Jj[6][6];
      edge_se3(iso_identity(), X[0], Zprior, e, Ji, Jj);
      const double w = infoPrior;
      for (int r = 0; r < 6; r++) {
        double s = 0;
        for (int k = 0; k < 6; k++) s += Jj[k][r] * e[k];
        bs[r] += -w * s;
      }
      add_se3_block(0, 0, &Jj[0][0], &Jj[0][0], 6, w);
    }
    for (int ed = 0; ed < NE; ed++) {
      double e[6], Ji[6][6], Jj[6][6];
      const int a = e6i[ed], c = e6j[ed], k6 = e6k[ed];
      edge_se3(X[a], X[c], Z6[ed], e, Ji, Jj);
      double chi = 0;
      for (int k = 0; k < 6; k++) chi += e[k] * e[k];
      huber(chi * info6[k6], d6[k6], rho);
      const double w = rho[1] * info6[k6];
      const int v[2] = {a, c};
      const double* J[2] = {&Ji[0][0], &Jj[0][0]};
      for (int s2 = 0; s2 < 2; s2++) {
        for (int r = 0; r < 6; r++) {
          double s = 0;
          for (int k = 0; k < 6; k++) s += J[s2][6 * k + r] * e[k];
          bs[6 * v[s2] + r] += -w * s;
        }
        for (int s3 = 0; s3 < 2; s3++) add_se3_block(v[s2], v[s3], J[s2], J[s3], 6, w);
      }
    }
    for (int o = 0; o < NO; o++) {
      double e[3], Ji[3][6], Jj[3][3];
      const int a = os[o], l = op[o];
      edge_se3_pointxyz(X[a], pts[l], meas[o], e, Ji, Jj);
      huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * info3[ok[o]], d3, rho);
      const double w = rho[1] * info3[ok[o]];
      for (int r = 0; r < 6; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Ji[k][r] * e[k];
        bs[6 * a + r] += -w * s;
        for (int q = 0; q < 3; q++) {
          double h = 0;
          for (int k = 0; k < 3; k++) h += Ji[k][r] * Jj[k][q];
          Bo[18 * (size_t)o + 3 * r + q] = w * h;
        }
      }
      add_se3_block(a, a, &Ji[0][0], &Ji[0][0], 3, w);
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Jj[k][r] * e[k];
        bl[3 * (size_t)l + r] += -w * s;
      }
      hl[l] += w;  // Jj^T Jj = R R^T = I
    }
    for (int t = 0; t < NT; t++) {
      double e[3], J2[3][3], JH[3][6];
      const int a = t1[t], c = t2[t], hv = th[t];
      edge_landmark_motion(X[hv], pts[a], pts[c], e, J2, JH);
      huber((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * infoT, dT, rho);
      const double w = rho[1] * infoT;
      for (int r = 0; r < 3; r++) {
        bl[3 * (size_t)a + r] += -w * e[r];  // J1 = I
        double s = 0;
        for (int k = 0; k < 3; k++) s += J2[k][r] * e[k];
        bl[3 * (size_t)c + r] += -w * s;
      }
      hl[a] += w; hl[c] += w;            // J2^T J2 = I
      for (int r = 0; r < 6; r++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += JH[k][r] * e[k];
        bs[6 * hv + r] += -w * s;
        for (int q = 0; q < 3; q++) {
          B1[18 * (size_t)t + 3 * r + q] = w * JH[q][r];          // JH^T J1
          double h = 0;
          for (int k = 0; k < 3; k++) h += JH[k][r] * J2[k][q];
          B2[18 * (size_t)t + 3 * r + q] = w * h;                  // JH^T J2
        }
      }
      add_se3_block(hv, hv, &JH[0][0], &JH[0][0], 3, w);
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) Ot[9 * (size_t)t + 3 * r + q] = w * J2[r][q];  // J1^T J2
    }
  }
  double max_diag() const {
    const int n = 6 * NS;
    double m = 0;
    for (int i = 0; i < n; i++) m = std::max(m, std::fabs(Hss[(size_t)i * n + i]));
    for (int l = 0; l < NP; l++) m = std::max(m, std::fabs(hl[l]));
    return m;
  }
  const double* cpl_block(const Cpl& c) const {
    return c.kind == 0 ? &Bo[18 * (size_t)c.e] : (c.kind == 1 ? &B1[18 * (size_t)c.e] : &B2[18 * (size_t)c.e]);
  }
  bool solve(double lambda) {
    const int n = 6 * NS;
    std::vector<double> S(Hss), bp(bs);
    for (int i = 0; i < n; i++) S[(size_t)i * n + i] += lambda;
    std::vector<double> Lkk((size_t)NP * 9, 0.0), Lk1((size_t)NP * 9, 0.0);  // Cholesky of every chain: diagonal and sub-diagonal blocks
    bool ok2 = true;
    const int nch = (int)chain_start.size() - 1;
    for (int c = 0; c < nch && ok2; c++) {
      std::vector<int> act;                 // active SE3 vertices
      std::vector<double> Y;                // [act][18] current column of B L^-T
      double cprev[3] = {0, 0, 0};
      for (int k = chain_start[c]; k < chain_start[c + 1]; k++) {
        const int p = chain_pts[k];
        double D[9] = {hl[p] + lambda, 0, 0, 0, hl[p] + lambda, 0, 0, 0, hl[p] + lambda};
        double* M = &Lk1[9 * (size_t)p];    // L_{k,k-1}
        if (k > chain_start[c]) {
          const int pp = chain_pts[k - 1];
          const double* O = &Ot[9 * (size_t)chain_link[k]];   // T_{k-1,k}; T_{k,k-1} = O^T
          const double* Lp = &Lkk[9 * (size_t)pp];
          // M = O^T Lp^-T : rows of O^T solved against Lp^T
          for (int r = 0; r < 3; r++) {
            const double a0 = O[r], a1 = O[3 + r], a2 = O[6 + r];
            const double y0 = a0 / Lp[0];
            const double y1 = (a1 - y0 * Lp[3]) / Lp[4];
            const double y2 = (a2 - y0 * Lp[6] - y1 * Lp[7]) / Lp[8];
            M[3 * r] = y0; M[3 * r + 1] = y1; M[3 * r + 2] = y2;
          }
          for (int r = 0; r < 3; r++)
            for (int q = 0; q < 3; q++) D[3 * r + q] -= M[3 * r] * M[3 * q] + M[3 * r + 1] * M[3 * q + 1] + M[3 * r + 2] * M[3 * q + 2];
        }
        double* L = &Lkk[9 * (size_t)p];
        for (int j = 0; j < 3 && ok2; j++) {
          double d = D[3 * j + j];
          for (int q = 0; q < j; q++) d -= L[3 * j + q] * L[3 * j + q];
          if (d <= 0) { ok2 = false; break; }
          L[3 * j + j] = std::sqrt(d);
          for (int i = j + 1; i < 3; i++) {
            double s = D[3 * i + j];
            for (int q = 0; q < j; q++) s -= L[3 * i + q] * L[3 * j + q];
            L[3 * i + j] = s / L[3 * j + j];
          }
        }
        if (!ok2) break;
        // Y_v <- -(Y_v M^T) for the active vertices, + B for the vertices coupled to p, then * L^-T
        for (size_t a = 0; a < act.size(); a++) {
          double* y = &Y[18 * a];
          for (int r = 0; r < 6; r++) {
            const double y0 = y[3 * r], y1 = y[3 * r + 1], y2 = y[3 * r + 2];
            for (int q = 0; q < 3; q++) y[3 * r + q] = -(y0 * M[3 * q] + y1 * M[3 * q + 1] + y2 * M[3 * q + 2]);
          }
        }
        for (int ci = cpl_start[p]; ci < cpl_start[p + 1]; ci++) {
          const Cpl& cp = cpl[ci];
          size_t a = std::find(act.begin(), act.end(), cp.v) - act.begin();
          if (a == act.size()) { act.push_back(cp.v); Y.resize(18 * act.size(), 0.0); }
          const double* B = cpl_block(cp);
          for (int q = 0; q < 18; q++) Y[18 * a + q] += B[q];
        }
        for (size_t a = 0; a < act.size(); a++) {
          double* y = &Y[18 * a];
          for (int r = 0; r < 6; r++) {
            y[3 * r] = y[3 * r] / L[0];
            y[3 * r + 1] = (y[3 * r + 1] - y[3 * r] * L[3]) / L[4];
            y[3 * r + 2] = (y[3 * r + 2] - y[3 * r] * L[6] - y[3 * r + 1] * L[7]) / L[8];
          }
        }
        // c_k = L^-1 (b_k - M c_{k-1})
        double rb[3], ck[3];
        for (int r = 0; r < 3; r++) rb[r] = bl[3 * (size_t)p + r] - (M[3 * r] * cprev[0] + M[3 * r + 1] * cprev[1] + M[3 * r + 2] * cprev[2]);
        ck[0] = rb[0] / L[0];
        ck[1] = (rb[1] - L[3] * ck[0]) / L[4];
        ck[2] = (rb[2] - L[6] * ck[0] - L[7] * ck[1]) / L[8];
        for (size_t a = 0; a < act.size(); a++) {
          const double* ya = &Y[18 * a];
          for (int r = 0; r < 6; r++) bp[6 * act[a] + r] -= ya[3 * r] * ck[0] + ya[3 * r + 1] * ck[1] + ya[3 * r + 2] * ck[2];
          for (size_t b2 = 0; b2 < act.size(); b2++) {
            const double* yb = &Y[18 * b2];
            for (int r = 0; r < 6; r++)
              for (int q = 0; q < 6; q++)
                S[(size_t)(6 * act[a] + r) * n + 6 * act[b2] + q] -= ya[3 * r] * yb[3 * q] + ya[3 * r + 1] * yb[3 * q + 1] + ya[3 * r + 2] * yb[3 * q + 2];
          }
        }
        memcpy(cprev, ck, sizeof ck);
      }
    }
    if (ok2) {  // dense LL^T of the SE3 block
      for (int j = 0; j < n && ok2; j++) {
        double d = S[(size_t)j * n + j];
        for (int k = 0; k < j; k++) d -= S[(size_t)j * n + k] * S[(size_t)j * n + k];
        if (d <= 0) { ok2 = false; break; }
        const double ljj = std::sqrt(d);
        S[(size_t)j * n + j] = ljj;
        for (int i = j + 1; i < n; i++) {
          double s = S[(size_t)i * n + j];
          const double* ri = &S[(size_t)i * n];
          const double* rj = &S[(size_t)j * n];
          for (int k = 0; k < j; k++) s -= ri[k] * rj[k];
          S[(size_t)i * n + j] = s / ljj;
        }
      }
    }
    if (!ok2) {
      for (int i = 0; i < n; i++) x[i] = bs[i];
      for (int i = 0; i < 3 * NP; i++) x[n + i] = bl[i];
      return false;
    }
    for (int i = 0; i < n; i++) {
      double s = bp[i];
      for (int k = 0; k < i; k++) s -= S[(size_t)i * n + k] * x[k];
      x[i] = s / S[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = x[i];
      for (int k = i + 1; k < n; k++) s -= S[(size_t)k * n + i] * x[k];
      x[i] = s / S[(size_t)i * n + i];
    }
    // points: T x_l = b_l - B^T x_s, chain by chain (forward with L, backward with L^T)
    for (int c = 0; c < nch; c++) {
      const int k0 = chain_start[c], k1 = chain_start[c + 1];
      std::vector<double> y(3 * (size_t)(k1 - k0));
      for (int k = k0; k < k1; k++) {
        const int p = chain_pts[k];
        double r[3] = {bl[3 * (size_t)p], bl[3 * (size_t)p + 1], bl[3 * (size_t)p + 2]};
        for (int ci = cpl_start[p]; ci < cpl_start[p + 1]; ci++) {
          const double* B = cpl_block(cpl[ci]);
          const double* xv = &x[6 * cpl[ci].v];
          for (int q = 0; q < 3; q++)
            for (int rr = 0; rr < 6; rr++) r[q] -= B[3 * rr + q] * xv[rr];
        }
        const double* M = &Lk1[9 * (size_t)p];
        const double* L = &Lkk[9 * (size_t)p];
        if (k > k0) {
          const double* yp = &y[3 * (size_t)(k - 1 - k0)];
          for (int q = 0; q < 3; q++) r[q] -= M[3 * q] * yp[0] + M[3 * q + 1] * yp[1] + M[3 * q + 2] * yp[2];
        }
        double* yk = &y[3 * (size_t)(k - k0)];
        yk[0] = r[0] / L[0];
        yk[1] = (r[1] - L[3] * yk[0]) / L[4];
        yk[2] = (r[2] - L[6] * yk[0] - L[7] * yk[1]) / L[8];
      }
      for (int k = k1 - 1; k >= k0; k--) {
        const int p = chain_pts[k];
        const double* L = &Lkk[9 * (size_t)p];
        double r[3] = {y[3 * (size_t)(k - k0)], y[3 * (size_t)(k - k0) + 1], y[3 * (size_t)(k - k0) + 2]};
        if (k + 1 < k1) {
          const int pn = chain_pts[k + 1];
          const double* Mn = &Lk1[9 * (size_t)pn];  // L_{k+1,k}
          const double* xn = &x[n + 3 * (size_t)pn];
          for (int q = 0; q < 3; q++) r[q] -= Mn[q] * xn[0] + Mn[3 + q] * xn[1] + Mn[6 + q] * xn[2];  // M^T x_{k+1}
        }
        double* xl = &x[n + 3 * (size_t)p];
        xl[2] = r[2] / L[8];
        xl[1] = (r[1] - L[7] * xl[2]) / L[4];
        xl[0] = (r[0] - L[3] * xl[1] - L[6] * xl[2]) / L[0];
      }
    }
    return true;
  }
  void update() {
    for (int i = 0; i < NS; i++) X[i] = se3_oplus(X[i], &x[6 * (size_t)i]);
    for (int l = 0; l < NP; l++) {
      const double* u = &x[6 * (size_t)NS + 3 * (size_t)l];
      pts[l] = pts[l] + V3{u[0], u[1], u[2]};
    }
  }
  void push() { bkX.push_back(X); bkP.push_back(pts); }
  void pop() { X = bkX.back(); pts = bkP.back(); bkX.pop_back(); bkP.pop_back(); }
  void discard_top() { bkX.pop_back(); bkP.pop_back(); }
  double compute_scale(double lambda) const {
    double s = 0;
    const int n = 6 * NS;
    for (int j = 0; j < n; j++) s += x[j] * (lambda * x[j] + bs[j]);
    for (int j = 0; j < 3 * NP; j++) s += x[n + j] * (lambda * x[n + j] + bl[j]);
    return s;
  }
};

}  // namespace vo

using namespace vo;

extern "C" {

void vo_fba_default_params(vo_fba_problem* p) {
  p->max_iterations = 300;
  p->sigma2_cam = 0.0001f; p->sigma2_3d_sta = 80.f; p->sigma2_3d_dyn = 80.f; p->sigma2_obj = 100.f; p->sigma2_smooth = 0.001f;
  p->huber_cam = 0.01f; p->huber_obj = 0.01f; p->huber_3d = 0.01f;
  p->gain_threshold = 1e-4f;
  p->prior_info = 100000.f;
}

int vo_ba_full(vo_fba_problem* p, vo_lm_stats* stats) {
  FBASystem S;
  S.NS = p->n_poses + p->n_motions; S.NP = p->n_points; S.NO = p->n_obs; S.NE = p->n_e6; S.NT = p->n_tern;
  if (S.NS == 0) { if (stats) { stats->iterations = -1; stats->n_records = 0; stats->total_trials = 0; } return -1; }
  S.info6[0] = 1.0 / (double)p->sigma2_cam; S.info6[1] = 1.0 / (double)p->sigma2_smooth;
  S.d6[0] = (double)p->huber_cam; S.d6[1] = (double)p->huber_cam;   // the smoothness edges use deltaHuberCamMot too (:1631)
  S.info3[0] = 1.0 / (double)p->sigma2_3d_sta; S.info3[1] = 1.0 / (double)p->sigma2_3d_dyn;
  S.d3 = (double)p->huber_3d;
  S.infoT = 1.0 / (double)p->sigma2_obj; S.dT = (double)p->huber_obj;
  S.infoPrior = (double)p->prior_info;
  S.X.resize(S.NS);
  for (int i = 0; i < S.NS; i++) S.X[i] = iso_from_f32(p->se3 + 16 * (size_t)i);
  S.Zprior = S.X[0];
  S.Z6.resize(S.NE);
  for (int e = 0; e < S.NE; e++) S.Z6[e] = iso_from_f32(p->e6_meas + 16 * (size_t)e);
  S.e6i.assign(p->e6_i, p->e6_i + S.NE); S.e6j.assign(p->e6_j, p->e6_j + S.NE); S.e6k.assign(p->e6_kind, p->e6_kind + S.NE);
  S.pts.resize(S.NP);
  for (int l = 0; l < S.NP; l++) S.pts[l] = {p->points[3 * (size_t)l], p->points[3 * (size_t)l + 1], p->points[3 * (size_t)l + 2]};
  S.meas.resize(S.NO);
  for (int o = 0; o < S.NO; o++) S.meas[o] = {p->obs_xyz[3 * (size_t)o], p->obs_xyz[3 * (size_t)o + 1], p->obs_xyz[3 * (size_t)o + 2]};
  S.os.assign(p->obs_se3, p->obs_se3 + S.NO); S.op.assign(p->obs_point, p->obs_point + S.NO); S.ok.assign(p->obs_kind, p->obs_kind + S.NO);
  S.t1.assign(p->tern_p1, p->tern_p1 + S.NT); S.t2.assign(p->tern_p2, p->tern_p2 + S.NT); S.th.assign(p->tern_h, p->tern_h + S.NT);
  S.e6err.assign(6 * (size_t)S.NE, 0.0); S.oerr.assign(3 * (size_t)S.NO, 0.0); S.terr.assign(3 * (size_t)S.NT, 0.0);
  const size_t n = 6 * (size_t)S.NS;
  S.Hss.assign(n * n, 0.0); S.bs.assign(n, 0.0);
  S.hl.assign(S.NP, 0.0); S.bl.assign(3 * (size_t)S.NP, 0.0);
  S.Bo.assign(18 * (size_t)S.NO, 0.0); S.B1.assign(18 * (size_t)S.NT, 0.0); S.B2.assign(18 * (size_t)S.NT, 0.0); S.Ot.assign(9 * (size_t)S.NT, 0.0);
  S.x.assign(S.dim(), 0.0);
  S.build_structure();
  const int its = lm_optimize(S, p->max_iterations, (double)p->gain_threshold, -1.0, stats);
  for (int i = 0; i < S.NS; i++) iso_to_f32(S.X[i], p->se3 + 16 * (size_t)i);
  for (int l = 0; l < S.NP; l++) {
    p->points[3 * (size_t)l] = (float)S.pts[l].x; p->points[3 * (size_t)l + 1] = (float)S.pts[l].y; p->points[3 * (size_t)l + 2] = (float)S.pts[l].z;
  }
  return its;
}

void vo_edge_landmark_motion(const double* H, const double* p1, const double* p2, double err[3], double J2[9], double JH[18]) {
  Iso X;
  memcpy(X.R.m, H, sizeof(double) * 9);
  X.t = {H[9], H[10], H[11]};
  double j2[3][3], jh[3][6];
  edge_landmark_motion(X, {p1[0], p1[1], p1[2]}, {p2[0], p2[1], p2[2]}, err, j2, jh);
  memcpy(J2, j2, sizeof j2);
  memcpy(JH, jh, sizeof jh);
}

void vo_edge_se3_prior(const double* Xd, const double* Zd, double err[6], double J[36]) {
  Iso X, Z;
  memcpy(X.R.m, Xd, sizeof(double) * 9); X.t = {Xd[9], Xd[10], Xd[11]};
  memcpy(Z.R.m, Zd, sizeof(double) * 9); Z.t = {Zd[9], Zd[10], Zd[11]};
  double Ji[6][6], Jj[6][6];
  edge_se3(iso_identity(), X, Z, err, Ji, Jj);
  memcpy(J, Jj, sizeof Jj);
}

}  // extern "C"
