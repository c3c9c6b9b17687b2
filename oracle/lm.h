// oracle/lm.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product.
//
// Restatement of the Levenberg-Marquardt trust-region loop that `ceres::Solve` runs for the
// reference (src/internal/icp-ceres.cpp:66-95: getOptionsMedium -> SPARSE_NORMAL_CHOLESKY,
// max_num_iterations = 50, everything else Ceres defaults).  Ceres itself is NOT under
// /root/reference and cannot be built in this image (needs Eigen/glog), and its version is
// unpinned (README.md:52 suggests libceres-dev 1.13) => PARITY UNPINNED for this file: the loop
// below follows Ceres 1.13's TrustRegionMinimizer + LevenbergMarquardtStrategy from knowledge of
// that source [ext-knowledge] (SURVEY.md section 8(a) row A9 lists the semantics restated here).
//
// The model is abstract: anything that can (a) evaluate cost [+ the unscaled local normal
// equations H = J^T J, g = J^T r after robust correction] and (b) apply Plus(x, delta).
// Because the linear solver is NORMAL-equation based and all other uses of J are through
// J^T J, J^T r and column norms, a dense H is an exact stand-in for Ceres' sparse Jacobian.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <vector>

namespace orc {

struct LmOptions {
  int max_num_iterations = 50;              // icp-ceres.cpp:81
  double initial_trust_region_radius = 1e4; // Ceres defaults [ext-knowledge] ...
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6;
  double max_lm_diagonal = 1e32;
  int max_num_consecutive_invalid_steps = 5;
  double function_tolerance = 1e-6;
  double gradient_tolerance = 1e-10;
  double parameter_tolerance = 1e-8;
  int jacobi_scaling = 1;
};

enum LmTermination {
  LM_CONVERGENCE_FUNCTION = 0,
  LM_CONVERGENCE_GRADIENT = 1,
  LM_CONVERGENCE_PARAMETER = 2,
  LM_NO_CONVERGENCE_MAX_ITER = 3,
  LM_CONVERGENCE_MIN_RADIUS = 4,
  LM_FAILURE_INVALID_STEPS = 5,
  LM_FAILURE_EVAL = 6,
};

struct LmIterationRecord {   // one row of the trace (golden fixtures compare these)
  int iteration; int step_valid; int step_accepted;
  double cost; double candidate_cost; double model_cost_change; double relative_decrease;
  double radius; double step_norm; double gradient_max_norm;
};

struct LmSummary {
  int termination = -1;
  int num_iterations = 0;          // number of step attempts (Ceres iteration index of last one)
  int num_successful_steps = 0;
  int num_jacobian_evals = 0;
  int num_cost_evals = 0;
  double initial_cost = 0, final_cost = 0;
  std::vector<LmIterationRecord> trace;
};

struct LmModel {
  int n_local = 0;     // columns of the reduced Jacobian
  int n_ambient = 0;   // size of x (non-constant blocks only)
  // evaluate: returns false on failure. H (n_local^2 row-major, full symmetric) and g may be null.
  std::function<bool(const double* x, double* cost, double* H, double* g)> evaluate;
  std::function<void(const double* x, const double* delta, double* x_plus)> plus;
};

// In-place dense Cholesky A = L L^T on the lower triangle; returns false if not SPD.
inline bool cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

inline double vec_norm(const std::vector<double>& v) {
  double s = 0; for (double e : v) s += e * e; return std::sqrt(s);
}

// Runs the loop; x is updated in place.
inline LmSummary lm_minimize(const LmModel& model, const LmOptions& opt, std::vector<double>& x) {
  LmSummary sum;
  const int n = model.n_local, na = model.n_ambient;
  std::vector<double> H((size_t)n * n), g(n), scale(n, 1.0), diagonal(n), lm_diag(n);
  std::vector<double> A((size_t)n * n), rhs(n), step(n), delta(n), cand(na), tmp(na), negg(n);
  double x_cost = 0, radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_invalid = 0;
  double x_norm = vec_norm(x);
  double gradient_max_norm = 0;

  // EvaluateGradientAndJacobian (+ scaling at iteration 0) + gradient-tolerance quantities.
  auto eval_grad_jac = [&](bool first) -> bool {
    if (!model.evaluate(x.data(), &x_cost, H.data(), g.data())) return false;
    ++sum.num_jacobian_evals;
    if (opt.jacobi_scaling && first)
      for (int j = 0; j < n; ++j) scale[j] = 1.0 / (1.0 + std::sqrt(H[(size_t)j * n + j]));
    for (int j = 0; j < n; ++j) negg[j] = -g[j];
    model.plus(x.data(), negg.data(), tmp.data());
    gradient_max_norm = 0;
    for (int i = 0; i < na; ++i) gradient_max_norm = std::fmax(gradient_max_norm, std::fabs(x[i] - tmp[i]));
    return true;
  };

  if (!eval_grad_jac(true)) { sum.termination = LM_FAILURE_EVAL; return sum; }
  sum.initial_cost = x_cost;
  sum.trace.push_back({0, 1, 1, x_cost, x_cost, 0, 0, radius, 0, gradient_max_norm});
  if (gradient_max_norm <= opt.gradient_tolerance) {
    sum.termination = LM_CONVERGENCE_GRADIENT; sum.final_cost = x_cost; return sum;
  }

  int iteration = 0;
  while (true) {
    if (iteration >= opt.max_num_iterations) { sum.termination = LM_NO_CONVERGENCE_MAX_ITER; break; }
    ++iteration;
    sum.num_iterations = iteration;
    LmIterationRecord rec{iteration, 0, 0, x_cost, 0, 0, 0, radius, 0, gradient_max_norm};

    // ---- LevenbergMarquardtStrategy::ComputeStep on the column-scaled Jacobian
    if (!reuse_diagonal)
      for (int j = 0; j < n; ++j) {
        double d = scale[j] * scale[j] * H[(size_t)j * n + j];
        diagonal[j] = std::fmin(std::fmax(d, opt.min_lm_diagonal), opt.max_lm_diagonal);
      }
    for (int j = 0; j < n; ++j) lm_diag[j] = std::sqrt(diagonal[j] / radius);
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = scale[i] * H[(size_t)i * n + j] * scale[j];
      A[(size_t)i * n + i] += lm_diag[i] * lm_diag[i];
      rhs[i] = scale[i] * g[i];
    }
    bool solved = cholesky_solve(A, n, rhs);
    for (int i = 0; i < n && solved; ++i) if (!std::isfinite(rhs[i])) solved = false;
    reuse_diagonal = true;

    double model_cost_change = 0;
    bool valid = solved;
    if (solved) {
      for (int i = 0; i < n; ++i) step[i] = -rhs[i];
      // -(J s).(r + J s / 2) = -s^T g~ - 1/2 s^T H~ s
      double sg = 0, sHs = 0;
      for (int i = 0; i < n; ++i) {
        sg += step[i] * scale[i] * g[i];
        double row = 0;
        for (int j = 0; j < n; ++j) row += scale[i] * H[(size_t)i * n + j] * scale[j] * step[j];
        sHs += step[i] * row;
      }
      model_cost_change = -sg - 0.5 * sHs;
      valid = (model_cost_change > 0.0);
    }
    rec.model_cost_change = model_cost_change;
    rec.step_valid = valid ? 1 : 0;

    if (!valid) {  // HandleInvalidStep
      if (++num_invalid >= opt.max_num_consecutive_invalid_steps) {
        sum.trace.push_back(rec); sum.termination = LM_FAILURE_INVALID_STEPS; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      sum.trace.push_back(rec);
      continue;
    }
    num_invalid = 0;
    for (int i = 0; i < n; ++i) delta[i] = step[i] * scale[i];
    model.plus(x.data(), delta.data(), cand.data());
    double cand_cost = 0;
    bool ok = model.evaluate(cand.data(), &cand_cost, nullptr, nullptr);
    ++sum.num_cost_evals;
    if (!ok || !std::isfinite(cand_cost)) {  // treated as an invalid step by Ceres
      if (++num_invalid >= opt.max_num_consecutive_invalid_steps) {
        sum.trace.push_back(rec); sum.termination = LM_FAILURE_INVALID_STEPS; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0;
      sum.trace.push_back(rec);
      continue;
    }
    rec.candidate_cost = cand_cost;

    // ParameterToleranceReached
    double sn = 0; for (int i = 0; i < na; ++i) sn += (x[i] - cand[i]) * (x[i] - cand[i]);
    rec.step_norm = std::sqrt(sn);
    if (rec.step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) {
      sum.trace.push_back(rec); sum.termination = LM_CONVERGENCE_PARAMETER; break;
    }
    // FunctionToleranceReached (checked before acceptance: Ceres >= 1.13)
    if (std::fabs(x_cost - cand_cost) <= opt.function_tolerance * x_cost) {
      sum.trace.push_back(rec); sum.termination = LM_CONVERGENCE_FUNCTION; break;
    }
    rec.relative_decrease = (x_cost - cand_cost) / model_cost_change;
    if (rec.relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
      x = cand; x_norm = vec_norm(x);
      if (!eval_grad_jac(false)) { sum.termination = LM_FAILURE_EVAL; break; }
      ++sum.num_successful_steps;
      rec.step_accepted = 1;
      const double q = 2.0 * rec.relative_decrease - 1.0;
      radius = radius / std::fmax(1.0 / 3.0, 1.0 - q * q * q);
      radius = std::fmin(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      sum.trace.push_back(rec);
      if (gradient_max_norm <= opt.gradient_tolerance) { sum.termination = LM_CONVERGENCE_GRADIENT; break; }
    } else {  // HandleUnsuccessfulStep
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      sum.trace.push_back(rec);
      if (radius < opt.min_trust_region_radius) { sum.termination = LM_CONVERGENCE_MIN_RADIUS; break; }
    }
  }
  sum.final_cost = x_cost;
  return sum;
}

}  // namespace orc
