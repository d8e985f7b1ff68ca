// ref_stubs.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only). Not part of the product.
//
// The reference's hot path links against three things this container does not have
// (SURVEY.md section 8c): nlopt (behind kel::Optimize, kel_math/kel_optimize.cpp:8,330-442),
// Boost (behind ParseVCFCigar) and the genome-annotation library (behind GenomeReference).
// This file supplies exactly the six missing symbols so that the reference's own translation
// units can be compiled where they lie and linked into oracle/_ref/kgl_ref_harness.
//
//  * kel::Optimize::{boundingHypercube, stoppingCriteria, returnDescription, run_optimize}
//    nlopt is an un-vendored dependency with NO pinned version (CMakeLists.txt:9,51,665 only name
//    -lnlopt). run_optimize below restates nlopt's published LN_NELDERMEAD (Nelder & Mead 1965 with
//    nlopt's bound handling: reflected points are clamped into the hypercube; default initial step
//    (ub-lb)/4; alpha=1, beta=0.5, gamma=2, delta=0.5; stop on xtol_abs over the simplex extent or on
//    maxeval; the best evaluated point is returned). "parity unpinned" for the optimiser itself: the
//    tests compare against the converged optimum of the reference's own objective instead.
//  * It additionally records, per calling thread, the start point the reference chose and the
//    objective on a fixed f-grid (the objective is reachable only through Optimize::objectiveCallback,
//    kel_math/kel_optimize.h:299), so the harness can export logLikelihood(f) golden vectors.
#include "kel_optimize.h"
#include "kgl_variant_factory_vcf_parse_cigar.h"
#include "kgl_genome_genome.h"
#include "kgl_pf7_fws_parser.h"
#include "kgl_Pf7_physical_distance.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <limits>
#include <random>
#include <dlfcn.h>

namespace kel = kellerberrin;
namespace kgl = kellerberrin::genome;

// ---- capture channel read by ref_harness.cpp -------------------------------------------------
namespace kglref {
thread_local std::vector<double> tl_start_points;   // f0 of every run_optimize call on this thread
thread_local std::vector<double> tl_end_points;     // optimiser result of every call
thread_local std::vector<double> tl_grid_values;    // objective on g_grid, captured on the first call
thread_local std::vector<size_t> tl_evals;
std::vector<double> g_grid;                          // set once by the harness before threads start
std::atomic<bool> g_fixed_seed_enabled{false};       // --seed: make the reference's start points reproducible
unsigned g_fixed_seed{0};
}  // namespace kglref

// The reference seeds every per-genome mt19937_64 from std::random_device (kel_math/kel_distribution.h:26-44, Q2), which
// makes HallME / Loglikelihood differ run to run. Without touching the reference sources, the harness executable
// interposes libstdc++'s std::random_device::_M_getval: with --seed S every draw returns S, so each genome's start
// sequence is mt19937_64(S) through uniform_real_distribution -- reproducible and restatable in the oracle.
std::random_device::result_type std::random_device::_M_getval() {
  if (kglref::g_fixed_seed_enabled.load(std::memory_order_relaxed)) return kglref::g_fixed_seed;
  using Fn = result_type (*)(std::random_device*);
  static Fn real = reinterpret_cast<Fn>(dlsym(RTLD_NEXT, "_ZNSt13random_device9_M_getvalEv"));
  return real ? real(this) : 0x5eedu;
}

void kel::Optimize::boundingHypercube(const std::vector<double>& upper_bound, const std::vector<double>& lower_bound) {
  upper_bound_ = upper_bound;
  lower_bound_ = lower_bound;
}

void kel::Optimize::stoppingCriteria(OptimizeStoppingType stopping_type, const std::vector<double>& stopping_value) {
  stopping_vector_.push_back(OptimalStopping{stopping_type, stopping_value});
}

std::string kel::Optimize::returnDescription(OptimizationResult result) {
  switch (result) {
    case OptimizationResult::FAILURE: return "FAILURE";
    case OptimizationResult::INVALID_ARGS: return "INVALID_ARGS";
    case OptimizationResult::OUT_OF_MEMORY: return "OUT_OF_MEMORY";
    case OptimizationResult::ROUNDOFF_LIMITED: return "ROUNDOFF_LIMITED";
    case OptimizationResult::FORCED_STOP: return "FORCED_STOP";
    case OptimizationResult::SUCCESS: return "SUCCESS";
    case OptimizationResult::STOPVAL_REACHED: return "STOPVAL_REACHED";
    case OptimizationResult::FTOL_REACHED: return "FTOL_REACHED";
    case OptimizationResult::XTOL_REACHED: return "XTOL_REACHED";
    case OptimizationResult::MAXEVAL_REACHED: return "MAXEVAL_REACHED";
    case OptimizationResult::MAXTIME_REACHED: return "MAXTIME_REACHED";
  }
  return "UNKNOWN";
}

namespace {

bool closeTo(double a, double b) { return std::fabs(a - b) <= 1e-13 * (std::fabs(a) + std::fabs(b)); }

// xnew = c + scale*(c - xold), clamped to [lb, ub]; false if the move degenerated.
bool reflectPoint(size_t n, std::vector<double>& xnew, const std::vector<double>& c, double scale,
                  const std::vector<double>& xold, const std::vector<double>& lb, const std::vector<double>& ub) {
  bool equalc = true, equalold = true;
  for (size_t i = 0; i < n; ++i) {
    double newx = c[i] + scale * (c[i] - xold[i]);
    newx = std::min(std::max(newx, lb[i]), ub[i]);
    equalc = equalc && closeTo(newx, c[i]);
    equalold = equalold && closeTo(newx, xold[i]);
    xnew[i] = newx;
  }
  return !(equalc || equalold);
}

}  // namespace

kel::OptResultTuple kel::Optimize::run_optimize(std::vector<double>& x, void* data) {
  const size_t n = dimension_;
  if (x.size() != n) return {OptimizationResult::INVALID_ARGS, 0.0, 0};
  const double sign = (opt_type_ == OptimizationType::MAXIMIZE) ? -1.0 : 1.0;  // minimise sign*f
  std::vector<double> lb(n, -std::numeric_limits<double>::infinity()), ub(n, std::numeric_limits<double>::infinity());
  if (!lower_bound_.empty()) lb = lower_bound_;
  if (!upper_bound_.empty()) ub = upper_bound_;
  std::vector<double> xtol_abs(n, 0.0);
  size_t maxeval = 0;
  for (auto const& [type, value] : stopping_vector_) {
    if (type == OptimizeStoppingType::ABSOLUTE_PARAMETER_THRESHOLD) {
      for (size_t i = 0; i < n; ++i) xtol_abs[i] = value.size() == 1 ? value.front() : value[i];
    } else if (type == OptimizeStoppingType::MAXIMUM_EVALUATIONS) {
      maxeval = static_cast<size_t>(value.front());
    }
  }

  std::vector<double> grad;  // derivative-free
  size_t evals = 0;
  auto rawEval = [&](const std::vector<double>& p) { return objectiveCallback(p, grad, data); };

  // Golden-vector capture: objective on the harness grid, before the optimiser moves anything.
  if (n == 1 && !kglref::g_grid.empty() && kglref::tl_grid_values.empty()) {
    for (double g : kglref::g_grid) kglref::tl_grid_values.push_back(rawEval(std::vector<double>{g}));
  }
  if (n == 1) kglref::tl_start_points.push_back(x[0]);

  std::vector<double> best_x = x;
  double best_f = std::numeric_limits<double>::infinity();
  bool hit_maxeval = false;
  auto eval = [&](const std::vector<double>& p) {
    const double f = sign * rawEval(p);
    ++evals;
    if (f < best_f) { best_f = f; best_x = p; }
    if (maxeval > 0 && evals >= maxeval) hit_maxeval = true;
    return f;
  };

  // nlopt default initial step.
  std::vector<double> step(n);
  for (size_t i = 0; i < n; ++i) {
    double s = std::numeric_limits<double>::infinity();
    if (std::isfinite(ub[i]) && std::isfinite(lb[i]) && (ub[i] - lb[i]) * 0.25 < s && ub[i] > lb[i]) s = (ub[i] - lb[i]) * 0.25;
    if (std::isfinite(ub[i]) && ub[i] - x[i] < s && ub[i] > x[i]) s = (ub[i] - x[i]) * 1.1;
    if (std::isfinite(lb[i]) && x[i] - lb[i] < s && lb[i] < x[i]) s = (x[i] - lb[i]) * 1.1;
    if (!std::isfinite(s)) s = std::fabs(x[i]);
    if (!std::isfinite(s) || s == 0.0) s = 1.0;
    step[i] = s;
  }

  // Simplex of n+1 points.
  std::vector<std::vector<double>> pts(n + 1, x);
  std::vector<double> fv(n + 1);
  OptimizationResult ret = OptimizationResult::SUCCESS;
  fv[0] = eval(pts[0]);
  for (size_t i = 0; i < n && !hit_maxeval; ++i) {
    auto& pt = pts[i + 1];
    pt[i] += step[i];
    if (pt[i] > ub[i]) {
      pt[i] = (ub[i] - x[i] > std::fabs(step[i]) * 0.1) ? ub[i] : x[i] - std::fabs(step[i]);
    }
    if (pt[i] < lb[i]) {
      if (x[i] - lb[i] > std::fabs(step[i]) * 0.1) pt[i] = lb[i];
      else {
        pt[i] = x[i] + std::fabs(step[i]);
        if (pt[i] > ub[i]) pt[i] = 0.5 * ((ub[i] - x[i] > x[i] - lb[i] ? ub[i] : lb[i]) + x[i]);
      }
    }
    if (closeTo(pt[i], x[i])) { ret = OptimizationResult::FAILURE; break; }
    fv[i + 1] = eval(pt);
  }

  const double alpha = 1.0, beta = 0.5, gamm = 2.0, delta = 0.5;
  std::vector<double> c(n), xcur(n);
  while (ret == OptimizationResult::SUCCESS && !hit_maxeval) {
    size_t lo = 0, hi = 0;
    for (size_t i = 1; i <= n; ++i) { if (fv[i] < fv[lo]) lo = i; if (fv[i] > fv[hi]) hi = i; }
    if (lo == hi) hi = (lo + 1) % (n + 1);
    double second_hi = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i <= n; ++i) if (i != hi) second_hi = std::max(second_hi, fv[i]);
    const double fl = fv[lo];
    double fh = fv[hi];
    auto& xh = pts[hi];

    std::fill(c.begin(), c.end(), 0.0);
    for (size_t i = 0; i <= n; ++i) if (i != hi) for (size_t j = 0; j < n; ++j) c[j] += pts[i][j];
    for (size_t j = 0; j < n; ++j) c[j] /= double(n);

    // x convergence: maximum extent of the simplex about the centroid.
    bool xstop = true;
    for (size_t j = 0; j < n; ++j) {
      double r = 0.0;
      for (size_t i = 0; i <= n; ++i) r = std::max(r, std::fabs(pts[i][j] - c[j]));
      if (!(r < xtol_abs[j])) xstop = false;
    }
    if (xstop) { ret = OptimizationResult::XTOL_REACHED; break; }

    if (!reflectPoint(n, xcur, c, alpha, xh, lb, ub)) { ret = OptimizationResult::XTOL_REACHED; break; }
    const double fr = eval(xcur);
    if (hit_maxeval) break;

    if (fr < fl) {  // new best: try to expand
      std::vector<double> xe(n);
      if (!reflectPoint(n, xe, c, gamm, xh, lb, ub)) { ret = OptimizationResult::XTOL_REACHED; break; }
      const double fe = eval(xe);
      if (fe >= fr) { xh = xcur; fh = fr; } else { xh = xe; fh = fe; }
    } else if (fr < second_hi) {  // accept
      xh = xcur; fh = fr;
    } else {  // contract
      std::vector<double> xc(n);
      if (!reflectPoint(n, xc, c, fh <= fr ? -beta : beta, xh, lb, ub)) { ret = OptimizationResult::XTOL_REACHED; break; }
      const double fc = eval(xc);
      if (fc < fr && fc < fh) { xh = xc; fh = fc; }
      else {  // shrink toward the best vertex
        const std::vector<double> xl = pts[lo];
        bool moved = true;
        for (size_t i = 0; i <= n && !hit_maxeval; ++i) {
          if (i == lo) continue;
          std::vector<double> xs(n);
          if (!reflectPoint(n, xs, xl, -delta, pts[i], lb, ub)) { moved = false; break; }
          pts[i] = xs;
          fv[i] = eval(xs);
        }
        if (!moved) { ret = OptimizationResult::XTOL_REACHED; break; }
        continue;
      }
    }
    fv[hi] = fh;
  }
  if (hit_maxeval && ret == OptimizationResult::SUCCESS) ret = OptimizationResult::MAXEVAL_REACHED;

  x = best_x;
  if (n == 1) { kglref::tl_end_points.push_back(x[0]); kglref::tl_evals.push_back(evals); }
  return {ret, sign * best_f, evals};
}

// ---- non-numeric link stubs (never executed on the hot path) ----------------------------------
std::string kgl::ParseVCFCigar::generateCigar(const std::string& reference, const std::string& alternate) {
  return std::to_string(reference.size()) + "R" + std::to_string(alternate.size()) + "A";
}

std::optional<std::shared_ptr<const kgl::ContigReference>> kgl::GenomeReference::getContigSequence(const ContigId_t&) const {
  return std::nullopt;
}

// HeteroHomoZygous::location_summary (kga_PfEMP/kga_analysis_PfEMP_heterozygous.cpp:266-358) asks the Pf7 physical-distance
// resource for the samples around a location; its parser needs the Boost-based file IO. Never reached: the harnesses assemble
// the location summaries themselves (plugin_harness.cpp, runPfEMP). The two Pf7FwsResource members that kgl_ref_harness does not
// link for real are in ref_stubs_pf7.cpp.
std::vector<std::string> kgl::Pf7SampleLocation::sampleRadius(const std::string&, double, bool) const { return {}; }
