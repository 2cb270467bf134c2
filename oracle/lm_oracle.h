// lm_oracle.h -- the reference's optimisation driver restated once, shared by every oracle graph.
// TEST INFRASTRUCTURE ONLY (see vido_oracle.h).
//   SparseOptimizer::optimize           g2o/core/sparse_optimizer.cpp:354-427 (incl. the local chi2_check patch :393-396)
//   OptimizationAlgorithmLevenberg      g2o/core/optimization_algorithm_levenberg.cpp:61-189 (incl. the nBad patch :154-161)
//   SparseOptimizerTerminateAction      g2o/core/sparse_optimizer_terminate_action.cpp:49-92
#pragma once
#include <cfloat>
#include <cmath>

#include "vido_oracle.h"

namespace vo {

// System concept:
//   int  num_vertices();  void compute_errors();  double robust_chi2();  void build_system();
//   double max_diag();  bool solve(double lambda);   // (H + lambda I) x = b; on failure x = b (linear_solver_csparse.h:126-133)
//   void update();  void push();  void pop();  void discard_top();  double compute_scale(double lambda);
template <class Sys>
int lm_optimize(Sys& S, int max_iterations, double gain_threshold /* <0: no terminate action */,
                double user_lambda_init /* <=0: tau*maxdiag */, vo_lm_stats* st) {
  if (st) { st->iterations = 0; st->n_records = 0; st->total_trials = 0; }
  if (S.num_vertices() == 0) return -1;  // "0 vertices to optimize"
  const double tau = 1e-5, upper = 2. / 3., lower = 1. / 3.;
  const int max_trials = 10;
  double lambda = -1, ni = 2;
  int nBad = 0;
  double chi2_check = 0.0, lastChi = 0.0;
  bool stop_flag = false, ok = true;
  int cj = 0;
  for (int i = 0; i < max_iterations && !stop_flag && ok; i++) {
    // ---- OptimizationAlgorithmLevenberg::solve(i)
    S.compute_errors();
    double currentChi = S.robust_chi2();
    double tempChi = currentChi;
    const double iniChi = currentChi;
    S.build_system();
    if (i == 0) {
      lambda = user_lambda_init > 0 ? user_lambda_init : tau * S.max_diag();
      ni = 2;
      nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      S.push();
      bool ok2 = S.solve(lambda);
      S.update();
      S.compute_errors();
      tempChi = S.robust_chi2();
      if (!ok2) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      double scale = S.compute_scale(lambda);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, upper);
        double scaleFactor = std::max(lower, alpha);
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        S.discard_top();
      } else {
        lambda *= ni;
        ni *= 2;
        S.pop();
      }
      qmax++;
    } while (rho < 0 && qmax < max_trials && !stop_flag);
    bool result_ok;
    if (qmax == max_trials || rho == 0) result_ok = false;
    else {
      if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
      else nBad = 0;
      result_ok = nBad < 3;
    }
    ok = result_ok;
    // ---- local patch of the reference: stop when the robust chi2 went up w.r.t. the previous iteration
    const double arc = S.robust_chi2();  // errors as left by the last trial
    if (chi2_check < arc && i > 0) ok = false;
    chi2_check = arc;
    if (st) {
      st->total_trials += qmax;
      if (st->n_records < VO_LM_MAX_RECORDS) {
        vo_lm_record& r = st->rec[st->n_records++];
        r.chi2 = currentChi;
        r.lambda = lambda;
        r.trials = qmax;
      }
    }
    ++cj;
    // ---- postIteration: SparseOptimizerTerminateAction
    if (gain_threshold >= 0) {
      S.compute_errors();
      const double chi = S.robust_chi2();
      if (i == 0) lastChi = chi;
      else {
        const double gain = (lastChi - chi) / chi;
        lastChi = chi;
        if (gain >= 0 && gain < gain_threshold) stop_flag = true;
      }
    }
  }
  if (st) st->iterations = cj;
  return cj;
}

}  // namespace vo
