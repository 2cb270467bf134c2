// lm_device.h -- the reference's Levenberg-Marquardt driver as a device-side state machine, executed by ONE thread
// of the solver kernel between barriers (no host round trips), or by the host between launches (full-sequence graph).
//
//   SparseOptimizer::optimize           g2o/core/sparse_optimizer.cpp:354-427 (local chi2_check patch :393-396)
//   OptimizationAlgorithmLevenberg      g2o/core/optimization_algorithm_levenberg.cpp:61-189 (nBad patch :154-161)
//   SparseOptimizerTerminateAction      g2o/core/sparse_optimizer_terminate_action.cpp:49-92
#pragma once
#include <cfloat>
#include <math.h>

#define VIDO_LM_REC 320
// the same state machine drives the in-kernel solvers (window BA, pose optimisation) and the host-side loop of the
// full-sequence optimisation (fba_kernels.cu)
#if defined(__CUDACC__)
#define LMHD __host__ __device__ __forceinline__
#else
#define LMHD inline
#endif

struct LmRec {
  double chi2, lambda;
  int trials, pad;
};

struct LmCtl {
  double lambda, ni, currentChi, iniChi, tempChi, lastTrialChi, chi2_check, lastChi, rho;
  int it, qmax, nBad, stop_flag, ok, accepted, fail, cur;
  int iterations, n_records, total_trials, pad;
};

LMHD void lm_reset(LmCtl* c) {
  c->lambda = -1; c->ni = 2; c->currentChi = 0; c->iniChi = 0; c->tempChi = 0; c->lastTrialChi = 0;
  c->chi2_check = 0; c->lastChi = 0; c->rho = 0;
  c->it = 0; c->qmax = 0; c->nBad = 0; c->stop_flag = 0; c->ok = 1; c->accepted = 0; c->fail = 0; c->cur = 0;
  c->iterations = 0; c->n_records = 0; c->total_trials = 0;
}

// start of Levenberg::solve(it): system has just been built; maxdiag only needed at it == 0
LMHD void lm_begin_iteration(LmCtl* c, int it, double maxdiag, double user_lambda) {
  c->iniChi = c->currentChi;
  c->tempChi = c->currentChi;
  if (it == 0) {
    c->lambda = user_lambda > 0 ? user_lambda : 1e-5 * maxdiag;  // _tau * max |H_jj|
    c->ni = 2;
    c->nBad = 0;
  }
  c->qmax = 0;
  c->rho = 0;
}

// after one trial: chi = robust chi2 at the trial state, scale = x^T(lambda x + b), failed = linear solver failed.
// Sets c->accepted and flips c->cur on acceptance.  Continue trying while (rho < 0 && qmax < 10).
LMHD void lm_trial(LmCtl* c, double chi, double scale, int failed) {
  c->lastTrialChi = chi;
  const double tempChi = failed ? DBL_MAX : chi;
  double rho = c->currentChi - tempChi;
  scale += 1e-3;
  rho /= scale;
  int accepted = 0;
  if (rho > 0 && isfinite(tempChi)) {
    double alpha = 1. - pow((2 * rho - 1), 3);
    alpha = fmin(alpha, 2. / 3.);
    const double sf = fmax(1. / 3., alpha);
    c->lambda *= sf;
    c->ni = 2;
    c->currentChi = tempChi;
    c->cur ^= 1;
    accepted = 1;
  } else {
    c->lambda *= c->ni;
    c->ni *= 2;
  }
  c->rho = rho;
  c->tempChi = tempChi;
  c->accepted = accepted;
  c->qmax += 1;
  c->fail = 0;
}

LMHD bool lm_more_trials(const LmCtl* c) { return c->rho < 0 && c->qmax < 10; }

// end of the iteration: stop rules of solve(), the chi2_check patch of optimize(), statistics, terminate action.
// gain_threshold < 0: no terminate action registered.
LMHD void lm_end_iteration(LmCtl* c, int it, double gain_threshold, LmRec* rec) {
  int result_ok;
  if (c->qmax == 10 || c->rho == 0) result_ok = 0;
  else {
    if ((c->iniChi - c->currentChi) * 1e3 < c->iniChi) c->nBad++;
    else c->nBad = 0;
    result_ok = c->nBad < 3;
  }
  int ok = result_ok;
  const double arc = c->lastTrialChi;  // activeRobustChi2() after solve(): errors of the last trial
  if (c->chi2_check < arc && it > 0) ok = 0;
  c->chi2_check = arc;
  c->total_trials += c->qmax;
  if (rec && c->n_records < VIDO_LM_REC) {
    LmRec& r = rec[c->n_records++];
    r.chi2 = c->currentChi;
    r.lambda = c->lambda;
    r.trials = c->qmax;
  }
  c->iterations = it + 1;
  if (gain_threshold >= 0) {
    const double chi = c->currentChi;  // the action recomputes the errors at the accepted state
    if (it == 0) c->lastChi = chi;
    else {
      const double gain = (c->lastChi - chi) / chi;
      c->lastChi = chi;
      if (gain >= 0 && gain < gain_threshold) c->stop_flag = 1;
    }
  }
  c->ok = ok;
  c->it = it + 1;
}
