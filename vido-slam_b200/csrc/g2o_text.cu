// g2o text writer for the flat FullBatch graph: what optimizer.save("dynamic_slam_graph_before_opt.g2o") and
// ("..._after_opt.g2o") print in Optimizer::FullBatchOptimization (src/Optimizer.cc:1937,1939).  Host code only.
// Layout: g2o/core/optimizable_graph.cpp:589-622 (parameters, vertices `TAG id estimate`, edges `TAG ids payload`);
// payloads: vertex_se3.cpp:58-64 (x y z qx qy qz qw), vertex_pointxyz.cpp:47-53, edge_se3.cpp:67-75 (7 + the upper triangle of
// the 6x6 information), edge_se3_prior.cpp:77-86 and edge_se3_pointxyz.cpp:88-96 (parameter id first),
// types_dyn_slam3d.cpp:44-51; tags: types_slam3d.cpp:37-45.  Numbers like a default std::ostream (%g, 6 digits) unless asked
// otherwise.  Vertex ids start at 1 (the reference's counter, :1353): SE3 vertices first, then the points.  The Python twin is
// vido-slam_b200/g2o_text.py; tests/test_g2o_golden.py checks both against text saved by the reference's own g2o build.
#include <cmath>
#include <cstdio>

#include "ctx.h"

namespace {
// unit quaternion (x, y, z, w), w >= 0, of the rotation part of a row-major float 4x4
void quat_of(const float* T, double q[4]) {
  const double R[3][3] = {{T[0], T[1], T[2]}, {T[4], T[5], T[6]}, {T[8], T[9], T[10]}};
  const double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    q[0] = (R[2][1] - R[1][2]) / s; q[1] = (R[0][2] - R[2][0]) / s; q[2] = (R[1][0] - R[0][1]) / s; q[3] = 0.25 * s;
  } else {
    int i = 0;
    if (R[1][1] > R[i][i]) i = 1;
    if (R[2][2] > R[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    const double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0) * 2;
    q[i] = 0.25 * s; q[j] = (R[j][i] + R[i][j]) / s; q[k] = (R[k][i] + R[i][k]) / s; q[3] = (R[k][j] - R[j][k]) / s;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double sg = q[3] < 0 ? -1.0 : 1.0;
  for (int c = 0; c < 4; c++) q[c] = sg * q[c] / n;
}
struct Printer {
  FILE* fh;
  int prec;
  void num(double v) { fprintf(fh, "%.*g ", prec, v); }
  void qt(const float* T) {
    double q[4];
    quat_of(T, q);
    num(T[3]); num(T[7]); num(T[11]);
    for (int c = 0; c < 4; c++) num(q[c]);
  }
  void info(double w, int n) {
    for (int i = 0; i < n; i++)
      for (int j = i; j < n; j++) num(i == j ? w : 0.0);
  }
};
}  // namespace

int fba_save_g2o(const vido_fba_problem* p, const char* path, int precision) {
  FILE* fh = fopen(path, "w");
  if (!fh) return VIDO_ERR_ARG;
  Printer P{fh, precision > 0 ? precision : 6};
  const int n_se3 = p->n_poses + p->n_motions, first = 1;
  auto sid = [&](int k) { return first + k; };
  auto pid = [&](int k) { return first + n_se3 + k; };
  fprintf(fh, "PARAMS_SE3OFFSET 0 ");
  for (int c = 0; c < 7; c++) P.num(c == 6 ? 1.0 : 0.0);
  fputc('\n', fh);
  for (int k = 0; k < n_se3; k++) { fprintf(fh, "VERTEX_SE3:QUAT %d ", sid(k)); P.qt(p->se3 + 16 * (size_t)k); fputc('\n', fh); }
  for (int k = 0; k < p->n_points; k++) {
    fprintf(fh, "VERTEX_TRACKXYZ %d ", pid(k));
    for (int c = 0; c < 3; c++) P.num(p->points[3 * (size_t)k + c]);
    fputc('\n', fh);
  }
  if (p->n_poses > 0) {   // the prior on the first camera pose, measurement = its estimate (:1369-1376)
    fprintf(fh, "EDGE_SE3_PRIOR %d 0 ", sid(0)); P.qt(p->se3); P.info((double)p->prior_info, 6); fputc('\n', fh);
  }
  for (int e = 0; e < p->n_e6; e++) {
    fprintf(fh, "EDGE_SE3:QUAT %d %d ", sid(p->e6_i[e]), sid(p->e6_j[e]));
    P.qt(p->e6_meas + 16 * (size_t)e);
    P.info(1.0 / (double)(p->e6_kind[e] == 0 ? p->sigma2_cam : p->sigma2_smooth), 6);
    fputc('\n', fh);
  }
  for (int e = 0; e < p->n_obs; e++) {
    fprintf(fh, "EDGE_SE3_TRACKXYZ %d %d 0 ", sid(p->obs_se3[e]), pid(p->obs_point[e]));
    for (int c = 0; c < 3; c++) P.num(p->obs_xyz[3 * (size_t)e + c]);
    P.info(1.0 / (double)(p->obs_kind[e] == 0 ? p->sigma2_3d_sta : p->sigma2_3d_dyn), 3);
    fputc('\n', fh);
  }
  for (int e = 0; e < p->n_tern; e++) {
    fprintf(fh, "EDGE_SE3_MOTION %d %d %d ", pid(p->tern_p1[e]), pid(p->tern_p2[e]), sid(p->tern_h[e]));
    for (int c = 0; c < 3; c++) P.num(0.0);
    P.info(1.0 / (double)p->sigma2_obj, 3);
    fputc('\n', fh);
  }
  const bool ok = !ferror(fh);
  fclose(fh);
  return ok ? VIDO_OK : VIDO_ERR_ARG;
}
