// The ONE implementation of the optimiser's scalar control flow, shared by
//   * the device-resident loop (kalibr_b200/csrc/kb_kernels.cu: lm_pre_solve / lm_post_solve / lm_post_eval kernels run these
//     functions on the control block in device memory), and
//   * the host mirror of Optimizer2 (include/kalibr_b200/optimizer.hpp), which drives any LinearSystemSolver through the same
//     transitions, and the reference-interface adapter test (tests/cpp/reference_adapter_main.cpp).
// It restates, as a small event-driven state machine, what the reference spreads over
//   Optimizer2::optimize                                   BE/src/Optimizer2.cpp:183-273
//   TrustRegionPolicy::solveSystem / get_dJ                BE/src/TrustRegionPolicy.cpp:28-57
//   LevenbergMarquardtTrustRegionPolicy                    BE/src/LevenbergMarquardtTrustRegionPolicy.cpp:37-113
//   GaussNewtonTrustRegionPolicy                           BE/src/GaussNewtonTrustRegionPolicy.cpp:18-40
//   BlockCholeskyLinearSystemSolver's lambda^2 / lambda residual   BE/src/BlockCholeskyLinearSystemSolver.cpp:77-97
// (BE = aslam_optimizer/aslam_backend).  One iteration is three events:
//   lm_before_solve  -> decides: rebuild the normal equations or not, the new lambda, the damping of this solve
//   lm_after_solve   <- (dx^T (lambda dx + rhs), max|dx|, positive definite?)  -> decides: apply + evaluate, or a failed iteration
//   lm_after_eval    <- cost of the trial state                               -> decides: accept or revert, and whether to go on
// Plain scalars only, no allocation, callable from host and device code.
#ifndef KALIBR_B200_LM_STATE_MACHINE_H_
#define KALIBR_B200_LM_STATE_MACHINE_H_

#include <math.h>

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ inline
#else
#define KB_HD inline
#endif

namespace kalibr_b200 {

enum { KB_POLICY_LEVENBERG_MARQUARDT = 0, KB_POLICY_GAUSS_NEWTON = 1 };

struct LmState {
  // ---- decisions of the current iteration (the device kernels are predicated on these) ----
  int done;        // the loop has ended
  int need_build;  // lm_before_solve: linearise + assemble before this solve (first iteration, or the last step was accepted with rho > 0)
  int skip_eval;   // lm_after_solve: the solve failed (not positive definite) - no update, no evaluation this iteration
  int revert;      // lm_after_eval: the step was a regression - restore the design variables (stays set until the next lm_after_eval:
                   // the device loop restores lazily, right before the next update or when the loop ends)
  double damping;  // lm_before_solve: what this solve adds to every diagonal entry of H (residual + lambda^2)
  double lambda;   // the conditioner (squared when applied: LinearSystemSolver.hpp:34-38)
  double cost_new; // device loop: where the evaluation of the trial state puts its cost
  // ---- optimiser state ----
  double J, pJ, deltaX, deltaJ, JStart;
  int iterations, failed, prev_failed, solver_failure;
  // ---- trust-region policy state ----
  double mu, gamma, beta, polJ, pol_pJ, pol_lastJ, rho_den, max_dx;
  int p_exp, first;
  // ---- solver state / options ----
  double diag_residual, conv_dx, conv_dj;
  int semantic;  // 0: BlockCholesky (un-augments with lambda: a residual accumulates between builds), 1: SparseCholesky (no residual)
  int max_iterations;
  int policy;    // KB_POLICY_*
};

// the while-condition of Optimizer2::optimize
KB_HD bool lm_should_continue(const LmState* c) {
  return c->iterations < c->max_iterations && c->failed < c->max_iterations &&
         ((c->deltaX > c->conv_dx && fabs(c->deltaJ) > c->conv_dj) || c->solver_failure);
}

// after the first evaluation (cost J0): optimizationStarting + the locals of Optimizer2::optimize
KB_HD void lm_start(LmState* c, int policy, double J0, double lambda_init, double conv_dx, double conv_dj, int max_iterations, int semantic) {
  c->done = 0; c->need_build = 1; c->skip_eval = 0; c->revert = 0;
  c->damping = 0.0;
  c->lambda = policy == KB_POLICY_LEVENBERG_MARQUARDT ? lambda_init : 0.0;
  c->cost_new = J0;
  c->J = c->pJ = c->JStart = J0;
  c->deltaX = conv_dx + 1.0;
  c->deltaJ = conv_dj + 1.0;
  c->iterations = c->failed = c->prev_failed = c->solver_failure = 0;
  c->mu = 2.0; c->gamma = 3.0; c->beta = 2.0; c->p_exp = 3;  // muInit, gammaInit, betaInit, pInit
  c->polJ = c->pol_pJ = c->pol_lastJ = J0;
  c->rho_den = 1.0;
  c->max_dx = 0.0;
  c->first = 1;
  c->diag_residual = 0.0;
  c->conv_dx = conv_dx; c->conv_dj = conv_dj;
  c->semantic = semantic;
  c->max_iterations = max_iterations < 0 ? 0 : max_iterations;
  c->policy = policy;
  if (!lm_should_continue(c)) c->done = 1;
}

// Before the solve.  Needs c->rho_den = dx^T (lambda dx + rhs) of the PREVIOUS solve (set by lm_after_solve, or by the caller when
// it evaluates getLmRho the reference's way: from rhs() and dx on the host, before any rebuild).
KB_HD void lm_before_solve(LmState* c) {
  // TrustRegionPolicy::solveSystem: a failed iteration keeps the cost the gain ratio is measured from
  if (!c->prev_failed) {
    c->pol_pJ = c->pol_lastJ;
    c->pol_lastJ = c->J;
  }
  c->polJ = c->J;
  int build = 0;
  double lambda = c->lambda;
  if (c->policy == KB_POLICY_GAUSS_NEWTON) {
    build = 1;  // rebuild and solve every iteration, no conditioner
    lambda = 0.0;
  } else if (c->first) {
    build = 1;
  } else {
    const double rho = (c->pol_pJ - c->polJ) / c->rho_den;  // getLmRho
    if (c->prev_failed) {  // the last step was a regression (or the solve failed): same system, heavier damping
      c->mu *= 2.0;
      lambda *= c->mu;
    } else if (rho <= 0.0) {
      c->mu *= 10.0;
      lambda *= c->mu;
    } else {  // accepted with a positive gain ratio: relinearise, Nielsen's update of lambda
      build = 1;
      if (lambda > 1e-16) {
        const double u1 = 1.0 / c->gamma;
        const double u2 = 1.0 - (c->beta - 1.0) * pow(2.0 * rho - 1.0, (double)c->p_exp);
        lambda *= (u1 > u2) ? u1 : u2;
        c->mu = c->beta;
      } else {
        lambda = 1e-15;
      }
    }
  }
  c->first = 0;
  c->lambda = lambda;
  c->need_build = build;
  if (build) c->diag_residual = 0.0;  // buildSystem clears H
  c->damping = c->diag_residual + lambda * lambda;
}

// After the solve: its scalars, the lambda^2 / lambda residual of the BlockCholesky semantic, and the failed-solve branch.
KB_HD void lm_after_solve(LmState* c, double rho_den, double max_dx, int pos_def) {
  c->rho_den = rho_den;
  c->max_dx = max_dx;
  if (c->semantic == 0) c->diag_residual += c->lambda * c->lambda - c->lambda;
  if (!pos_def) {
    c->prev_failed = 1;
    c->solver_failure = 1;  // sticky: the loop then only ends on the iteration limits
    c->failed += 1;
    c->skip_eval = 1;
    if (!lm_should_continue(c)) c->done = 1;
  } else {
    c->skip_eval = 0;
  }
}

// After the update was applied and the trial state evaluated (cost_new).  Sets c->revert when the step must be undone.
KB_HD void lm_after_eval(LmState* c, double cost_new) {
  c->deltaX = c->max_dx;
  c->J = cost_new;
  c->deltaJ = c->pJ - c->J;
  c->revert = 0;
  if (c->policy == KB_POLICY_LEVENBERG_MARQUARDT) {  // revertOnFailure()
    if (c->deltaJ < 0.0) {
      c->revert = 1;
      c->failed += 1;
      c->prev_failed = 1;
    } else {
      c->pJ = c->J;
      c->prev_failed = 0;
    }
  } else {
    c->pJ = c->J;  // Gauss-Newton never reverts (and never clears a failure flag either)
  }
  c->iterations += 1;
  if (!lm_should_continue(c)) c->done = 1;
}

}  // namespace kalibr_b200

#undef KB_HD
#endif  // KALIBR_B200_LM_STATE_MACHINE_H_
