// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement of the WALNUTS hot path of flatironinstitute/walnuts
// ("walnutpie"), Eigen-free, plain fp64, sequential left-to-right sums.
// Every function cites the reference file:line (relative to /root/reference)
// it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// Pinning: (i) the reference's own known-answer tests (tests/util_test.cpp,
// tests/config_test.cpp) are replayed in tests/test_oracle_kat.py; (ii) the
// UNMODIFIED reference headers are compiled against a local Eigen API shim
// (oracle/eigen_shim, oracle/ref_capi.cpp -> oracle/_ref/) and this
// restatement must reproduce their transitions, warm-up and initialisation
// bit for bit on the same std::mt19937_64 stream (tests/test_oracle_vs_ref.py,
// golden copies in tests/golden/).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <exception>
#include <functional>
#include <limits>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "philox.hpp"
#include "targets.hpp"

namespace oracle {

// ---------------------------------------------------------------------------
// Random<RNG> over a stateful engine (util.hpp:78-162): libstdc++
// distributions, one object per sampler so the cached polar-method spare of
// std::normal_distribution lives and dies exactly as in the reference.
template <class RNG>
class StdRand {
 public:
  explicit StdRand(RNG& rng)
      : rng_(rng), unif_(0.0, 1.0), binary_(0.5), normal_(0.0, 1.0) {}
  void begin_transition(uint32_t) {}
  double uniform_real_01() { return unif_(rng_); }     // util.hpp:102
  bool uniform_binary() { return binary_(rng_); }      // util.hpp:112
  Vec standard_normal(std::size_t n) {                 // util.hpp:124-127
    Vec z(n);
    for (std::size_t i = 0; i < n; ++i) z[i] = normal_(rng_);
    return z;
  }
  RNG& rng() { return rng_; }

 private:
  RNG& rng_;
  std::uniform_real_distribution<double> unif_;
  std::bernoulli_distribution binary_;
  std::normal_distribution<double> normal_;
};

// ---------------------------------------------------------------------------
// util.hpp:174-183
inline double log_sum_exp(double x1, double x2) {
  double m = std::fmax(x1, x2);
  if (std::isnan(x1) || std::isnan(x2)) {
    return std::numeric_limits<double>::quiet_NaN();
  }
  if (std::isinf(m) || std::isnan(x1 + x2)) {
    return std::fmax(x1, x2);
  }
  return m + std::log(std::exp(x1 - m) + std::exp(x2 - m));
}

// util.hpp:195-205
inline double log_sum_exp(const Vec& x) {
  if (x.empty()) return -std::numeric_limits<double>::infinity();
  double m = x[0];
  bool any_nan = std::isnan(x[0]);
  for (std::size_t i = 1; i < x.size(); ++i) {
    if (std::isnan(x[i])) any_nan = true;
    if (x[i] > m) m = x[i];
  }
  if (any_nan) return std::numeric_limits<double>::quiet_NaN();
  if (std::isinf(m)) return m;
  double s = 0.0;
  for (double xi : x) s += std::exp(xi - m);
  return m + std::log(s);
}

// util.hpp:220-223
inline double logp_momentum(const Vec& rho, const Vec& inv_mass) {
  double s = 0.0;
  for (std::size_t i = 0; i < rho.size(); ++i) {
    s = madd(inv_mass[i], rho[i] * rho[i], s);
  }
  return -0.5 * s;
}

// util.hpp:379-382
inline double l2_rel_diff(const Vec& a, const Vec& b) {
  double s = 0.0;
  for (std::size_t i = 0; i < a.size(); ++i) {
    double r = (a[i] - b[i]) / b[i];
    s += r * r;
  }
  return std::sqrt(s);
}

// util.hpp:401-404
inline double variance(const Vec& xs) {
  double m = 0.0;
  for (double x : xs) m += x;
  m /= static_cast<double>(xs.size());
  double s = 0.0;
  for (double x : xs) s += (x - m) * (x - m);
  // 0/0 -> NaN for one element, exactly like the reference's size_t(n-1) cast
  return s / static_cast<double>(xs.size() - 1);
}

// util.hpp:242-259
template <class F>
double leapfrog_error(const F& logp_grad, const Vec& theta, const Vec& rho,
                      const Vec& inv_M, double step) {
  const std::size_t D = theta.size();
  Vec grad;
  double logp;
  logp_grad(theta, logp, grad);
  logp += logp_momentum(rho, inv_M);
  const double hs = 0.5 * step;
  Vec rho_star(D), theta_star(D);
  for (std::size_t i = 0; i < D; ++i) rho_star[i] = rho[i] + hs * grad[i];
  for (std::size_t i = 0; i < D; ++i) {
    theta_star[i] = theta[i] + step * (inv_M[i] * rho_star[i]);
  }
  double logp_star;
  logp_grad(theta_star, logp_star, grad);
  for (std::size_t i = 0; i < D; ++i) rho_star[i] = rho_star[i] + hs * grad[i];
  logp_star += logp_momentum(rho_star, inv_M);
  return logp_star - logp;
}

// util.hpp:285-303.  `rand` is a fresh Random over the caller's engine.
template <class Rand, class F>
double adapt_step(Rand& rand, const F& logp_grad, const Vec& theta,
                  const Vec& M, double step, std::size_t D) {
  Vec inv_M(D), rho(D);
  for (std::size_t i = 0; i < D; ++i) inv_M[i] = 1.0 / M[i];
  Vec z = rand.standard_normal(D);
  for (std::size_t i = 0; i < D; ++i) rho[i] = z[i] * std::sqrt(M[i]);
  while (leapfrog_error(logp_grad, theta, rho, inv_M, step) > std::log(0.9)) {
    step *= 2;
  }
  while (leapfrog_error(logp_grad, theta, rho, inv_M, step) < std::log(0.6)) {
    step *= std::sqrt(0.5);
  }
  return step;
}

// config.hpp:360-370 (per-chain part)
template <class F>
Vec init_mass_from_gradient(const F& logp_grad, const Vec& position,
                            double mass_smoothing) {
  if (!(mass_smoothing > 0 && mass_smoothing < 1)) {
    throw std::invalid_argument("mass_smoothing must be in (0, 1)");
  }
  Vec grad;
  double lp;
  logp_grad(position, lp, grad);
  Vec mass(position.size());
  for (std::size_t i = 0; i < mass.size(); ++i) {
    mass[i] = (1 - mass_smoothing) * std::fabs(grad[i]) + mass_smoothing;
  }
  return mass;
}

// ---------------------------------------------------------------------------
// util.hpp:311-351: exceptions from the density become logp=-inf, grad=0.
template <class F>
class NoExceptLogpGrad {
 public:
  using OnError = std::function<void(const Vec&, const std::exception&)>;
  NoExceptLogpGrad(const F& f, OnError on_error, uint64_t* counter = nullptr)
      : f_(&f), on_error_(std::move(on_error)), counter_(counter) {}
  void operator()(const Vec& x, double& logp, Vec& grad) const noexcept {
    if (counter_) ++*counter_;
    try {
      (*f_)(x, logp, grad);
    } catch (const std::exception& e) {
      if (on_error_) on_error_(x, e);
      logp = -std::numeric_limits<double>::infinity();
      grad.assign(x.size(), 0.0);
    }
  }
  const F& base() const { return *f_; }

 private:
  const F* f_;
  OnError on_error_;
  uint64_t* counter_;
};

// ---------------------------------------------------------------------------
// adam.hpp:35-109
class Adam {
 public:
  Adam(double step_size_init, double accept_rate_target, double learning_rate,
       double gradient_decay, double sq_gradient_decay, double stabilization,
       double learn_rate_decay)
      : theta_(std::log(step_size_init)), m_(0), v_(0), t_(0), b1p_(1),
        b2p_(1), target_(accept_rate_target), lr_(learning_rate),
        b1_(gradient_decay), b2_(sq_gradient_decay), eps_(stabilization),
        lr_decay_(learn_rate_decay) {}

  void operator()(double alpha) noexcept {  // adam.hpp:70-86
    ++t_;
    b1p_ *= b1_;
    b2p_ *= b2_;
    double grad = target_ - alpha;
    m_ = b1_ * m_ + (1 - b1_) * grad;
    v_ = b2_ * v_ + (1 - b2_) * grad * grad;
    double m_hat = m_ / (1 - b1p_);
    double v_hat = v_ / (1 - b2p_);
    double decayed = lr_ / std::pow(t_, lr_decay_);
    double denom = std::sqrt(v_hat) + eps_;
    theta_ -= decayed * m_hat / denom;
  }
  double step_size() const noexcept { return std::exp(theta_); }  // :93
  double log_step() const noexcept { return theta_; }

 private:
  double theta_, m_, v_, t_, b1p_, b2p_;
  const double target_, lr_, b1_, b2_, eps_, lr_decay_;
};

// walnuts.hpp:572-587
struct NoOpStepSizeAdapter {
  void operator()(double) const noexcept {}
};

// online_moments.hpp:22-86
class WelfordAccumulator {
 public:
  void observe(double x) {
    ++n_;
    const double delta = x - mean_;
    mean_ += delta / static_cast<double>(n_);
    const double delta2 = x - mean_;
    M2_ += delta * delta2;
  }
  std::size_t count() const { return n_; }
  double mean() const { return mean_; }
  double sample_variance() const {
    return n_ > 1 ? (M2_ / static_cast<double>(n_ - 1))
                  : std::numeric_limits<double>::quiet_NaN();
  }

 private:
  std::size_t n_ = 0;
  double mean_ = 0.0, M2_ = 0.0;
};

// online_moments.hpp:125-246.  NB `auto delta = y - mean_` in the reference is
// a lazy Eigen expression, so after `mean_ += delta / weight_` BOTH factors of
// the cross term see the updated mean: S = gamma S + (y - mu_new)^2
// (reproduced by the shim build of the unmodified header).
class OnlineMoments {
 public:
  OnlineMoments() : weight_(0) {}
  OnlineMoments(double init_weight, const Vec& init_mean,
                const Vec& init_variance)
      : weight_(init_weight), mean_(init_mean), ssd_(init_variance.size()) {
    if (!(init_weight > 0 && !std::isinf(init_weight))) {
      throw std::invalid_argument("init_weight must be in (0, inf).");
    }
    if (init_mean.size() != init_variance.size()) {
      throw std::invalid_argument(
          "init_mean and init_variance must be the same size.");
    }
    for (std::size_t i = 0; i < ssd_.size(); ++i) {
      ssd_[i] = init_weight * init_variance[i];
    }
  }
  void discount_observe(double gamma, const Vec& y) {  // :185-207
    if (!(gamma >= 0 && gamma <= 1)) {
      throw std::invalid_argument("discount_factor must be in [0, 1]");
    }
    weight_ = gamma * weight_ + 1;
    if (fused_arith()) {  // one reciprocal of the scalar weight, then multiply-adds
      const double r_w = 1.0 / weight_;
      for (std::size_t i = 0; i < y.size(); ++i) {
        mean_[i] = std::fma(y[i] - mean_[i], r_w, mean_[i]);
      }
    } else {
      for (std::size_t i = 0; i < y.size(); ++i) {
        mean_[i] += (y[i] - mean_[i]) / weight_;
      }
    }
    for (std::size_t i = 0; i < y.size(); ++i) {
      double d = y[i] - mean_[i];
      ssd_[i] = madd(gamma, ssd_[i], d * d);
    }
  }
  const Vec& mean() const { return mean_; }
  const Vec& sum_sq_dev() const { return ssd_; }
  double weight() const { return weight_; }
  Vec variance() const {  // :225-230
    Vec v(mean_.size());
    if (!(weight_ > 0)) {
      std::fill(v.begin(), v.end(), 1.0);
      return v;
    }
    for (std::size_t i = 0; i < v.size(); ++i) v[i] = ssd_[i] / weight_;
    return v;
  }

 private:
  double weight_;
  Vec mean_, ssd_;
};

// ---------------------------------------------------------------------------
// config.hpp:513-641 / :885-954 (values + the validation that can fire from
// the C-ABI; the builders themselves are host logic of the product)
struct WarmupConfig {
  std::size_t min_iter = 50, max_iter = 1000;
  double step_size_converge_tol = 0.1, mass_converge_tol = 1.0;
  double mass_init_count = 4.0, mass_additive_smoothing = 1e-5;
  double max_macro_steps_target = 15.0;
  double step_accept_rate_target = 0.8, step_learning_rate = 0.05;
  double step_gradient_decay = 0.8, step_sq_gradient_decay = 0.9;
  double step_stabilization = 1e-4, step_learn_rate_decay = 0.5;
  std::size_t publish_stride = 5, yield_period = 32;
};

struct SamplingConfig {
  std::size_t min_iter = 50, max_iter = 1000;
  std::size_t max_trajectory_doublings = 5, max_step_halvings = 5;
  double max_hamiltonian_error = 0.5;
  std::size_t min_micro_steps = 1;
  double rhat_converge_tol = 1.01;
};

struct InitChainConfig {
  double step_size;
  Vec position;
  Vec mass;
};

// ---------------------------------------------------------------------------
// walnuts.hpp:34-131
struct SpanW {
  Vec theta_bk, rho_bk, grad_bk;
  double logp_bk;
  Vec theta_fw, rho_fw, grad_fw;
  double logp_fw;
  Vec theta_select, grad_select;
  double logp_pos_select;
  double logp;  // log sum of joint densities over the span

  static SpanW from_initial_point(Vec&& theta, Vec&& rho, Vec&& grad,
                                  double logp_pos, double logp_joint) {
    SpanW s;
    s.theta_bk = theta; s.rho_bk = rho; s.grad_bk = grad;
    s.logp_bk = logp_joint;
    s.theta_fw = theta; s.rho_fw = std::move(rho); s.grad_fw = grad;
    s.logp_fw = logp_joint;
    s.theta_select = std::move(theta); s.grad_select = std::move(grad);
    s.logp_pos_select = logp_pos;
    s.logp = logp_joint;
    return s;
  }
  static SpanW from_subspans(SpanW&& span1, SpanW&& span2, Vec&& theta_select,
                             Vec&& grad_select, double logp_pos_select,
                             double logp_total) {
    SpanW s;
    s.theta_bk = std::move(span1.theta_bk);
    s.rho_bk = std::move(span1.rho_bk);
    s.grad_bk = std::move(span1.grad_bk);
    s.logp_bk = span1.logp_bk;
    s.theta_fw = std::move(span2.theta_fw);
    s.rho_fw = std::move(span2.rho_fw);
    s.grad_fw = std::move(span2.grad_fw);
    s.logp_fw = span2.logp_fw;
    s.theta_select = std::move(theta_select);
    s.grad_select = std::move(grad_select);
    s.logp_pos_select = logp_pos_select;
    s.logp = logp_total;
    return s;
  }
};

enum class Update { Barker, Metropolis };
enum class Direction { Backward, Forward };

// walnuts.hpp:192-201
template <Direction D>
bool uturn(const SpanW& span1, const SpanW& span2, const Vec& inv_mass) {
  const SpanW& span_bk = (D == Direction::Forward) ? span1 : span2;
  const SpanW& span_fw = (D == Direction::Forward) ? span2 : span1;
  const std::size_t n = inv_mass.size();
  double dot_fw = 0.0, dot_bk = 0.0;
  for (std::size_t i = 0; i < n; ++i) {
    double sd = inv_mass[i] * (span_fw.theta_fw[i] - span_bk.theta_bk[i]);
    dot_fw = madd(span_fw.rho_fw[i], sd, dot_fw);
    dot_bk = madd(span_bk.rho_bk[i], sd, dot_bk);
  }
  return dot_fw < 0 || dot_bk < 0;
}

// one leapfrog micro-step, walnuts.hpp:329-332 (== :228-231)
template <class F>
inline void leapfrog(const F& logp_grad, const Vec& inv_mass, double step,
                     double half_step, Vec& theta, Vec& rho, Vec& grad,
                     double& logp_pos) {
  const std::size_t D = theta.size();
  for (std::size_t i = 0; i < D; ++i) rho[i] = madd(half_step, grad[i], rho[i]);
  for (std::size_t i = 0; i < D; ++i) theta[i] = madd(step * inv_mass[i], rho[i], theta[i]);
  logp_grad(theta, logp_pos, grad);
  for (std::size_t i = 0; i < D; ++i) rho[i] = madd(half_step, grad[i], rho[i]);
}

// walnuts.hpp:218-235
template <class F>
bool within_tolerance(const F& logp_grad, const Vec& inv_mass, double step,
                      std::size_t num_steps, double max_error,
                      double logp_next, Vec& theta_next, Vec& rho_next,
                      Vec& grad_next) {
  double half_step = 0.5 * step;
  double logp = logp_next;
  for (std::size_t n = 0; n < num_steps; ++n) {
    leapfrog(logp_grad, inv_mass, step, half_step, theta_next, rho_next,
             grad_next, logp_next);
  }
  logp_next += logp_momentum(rho_next, inv_mass);
  return std::abs(logp_next - logp) <= max_error;
}

// walnuts.hpp:254-279
template <class F>
bool reversible(const F& logp_grad, const Vec& inv_mass, double step,
                std::size_t num_steps, std::size_t min_micro_steps,
                double max_error, double logp_next, const Vec& theta,
                const Vec& rho, const Vec& grad) {
  if (num_steps == 1) return true;
  Vec theta_next(theta.size()), rho_next(theta.size()), grad_next(theta.size());
  while (num_steps >= 2 * min_micro_steps) {
    theta_next = theta;
    for (std::size_t i = 0; i < rho.size(); ++i) rho_next[i] = -rho[i];
    grad_next = grad;
    num_steps /= 2;
    step *= 2;
    if (within_tolerance(logp_grad, inv_mass, step, num_steps, max_error,
                         logp_next, theta_next, rho_next, grad_next)) {
      return false;
    }
  }
  return true;
}

// walnuts.hpp:307-345
template <Direction D, class F, class A>
bool macro_step(const F& logp_grad, const Vec& inv_mass, double step,
                std::size_t max_step_halvings, std::size_t min_micro_steps,
                double max_error, const SpanW& span, Vec& theta_next,
                Vec& rho_next, Vec& grad_next, double& logp_pos_next,
                double& logp_next, A& adapt_handler) {
  constexpr bool is_forward = (D == Direction::Forward);
  const Vec& theta = is_forward ? span.theta_fw : span.theta_bk;
  const Vec& rho = is_forward ? span.rho_fw : span.rho_bk;
  const Vec& grad = is_forward ? span.grad_fw : span.grad_bk;
  double logp = is_forward ? span.logp_fw : span.logp_bk;
  step = is_forward ? step : -step;
  for (std::size_t num_steps = min_micro_steps, halvings = 0;
       halvings < max_step_halvings; ++halvings, num_steps *= 2, step *= 0.5) {
    theta_next = theta;
    rho_next = rho;
    grad_next = grad;
    double half_step = 0.5 * step;
    for (std::size_t n = 0; n < num_steps; ++n) {
      leapfrog(logp_grad, inv_mass, step, half_step, theta_next, rho_next,
               grad_next, logp_pos_next);
    }
    logp_next = logp_pos_next + logp_momentum(rho_next, inv_mass);
    if (num_steps == min_micro_steps) {
      double min_accept = std::exp(-std::fabs(logp - logp_next));
      adapt_handler(min_accept);
    }
    if (std::fabs(logp - logp_next) <= max_error) {
      return reversible(logp_grad, inv_mass, step, num_steps, min_micro_steps,
                        max_error, logp_next, theta_next, rho_next, grad_next);
    }
  }
  return false;
}

// walnuts.hpp:368-387
template <Update U, Direction D, class Rand>
SpanW combine(Rand& rng, SpanW&& span_old, SpanW&& span_new) {
  double logp_total = log_sum_exp(span_old.logp, span_new.logp);
  double log_denominator =
      (U == Update::Metropolis) ? span_old.logp : logp_total;
  double update_logprob = span_new.logp - log_denominator;
  bool update = std::log(rng.uniform_real_01()) < update_logprob;
  Vec& selected = update ? span_new.theta_select : span_old.theta_select;
  Vec& grad_selected = update ? span_new.grad_select : span_old.grad_select;
  double logp_pos_select =
      update ? span_new.logp_pos_select : span_old.logp_pos_select;
  SpanW& span_bk = (D == Direction::Forward) ? span_old : span_new;
  SpanW& span_fw = (D == Direction::Forward) ? span_new : span_old;
  return SpanW::from_subspans(std::move(span_bk), std::move(span_fw),
                              std::move(selected), std::move(grad_selected),
                              logp_pos_select, logp_total);
}

// walnuts.hpp:420-442
template <Direction D, class F, class A>
std::optional<SpanW> build_leaf(const F& logp_grad, const SpanW& span,
                                const Vec& inv_mass, double step,
                                std::size_t max_step_halvings,
                                std::size_t min_micro_steps, double max_error,
                                A& adapt_handler) {
  Vec theta_next, rho_next, grad_next;
  double logp_pos_next = -std::numeric_limits<double>::infinity();
  double logp_next = -std::numeric_limits<double>::infinity();
  if (!macro_step<D>(logp_grad, inv_mass, step, max_step_halvings,
                     min_micro_steps, max_error, span, theta_next, rho_next,
                     grad_next, logp_pos_next, logp_next, adapt_handler)) {
    return std::nullopt;
  }
  return SpanW::from_initial_point(std::move(theta_next), std::move(rho_next),
                                   std::move(grad_next), logp_pos_next,
                                   logp_next);
}

// walnuts.hpp:464-495
template <Direction D, class F, class Rand, class A>
std::optional<SpanW> build_span(Rand& rng, const F& logp_grad,
                                const Vec& inv_mass, double step,
                                std::size_t depth,
                                std::size_t max_step_halvings,
                                std::size_t min_micro_steps, double max_error,
                                const SpanW& last_span, A& adapt_handler) {
  if (depth == 0) {
    return build_leaf<D>(logp_grad, last_span, inv_mass, step,
                         max_step_halvings, min_micro_steps, max_error,
                         adapt_handler);
  }
  auto sub1 = build_span<D>(rng, logp_grad, inv_mass, step, depth - 1,
                            max_step_halvings, min_micro_steps, max_error,
                            last_span, adapt_handler);
  if (!sub1) return std::nullopt;
  auto sub2 = build_span<D>(rng, logp_grad, inv_mass, step, depth - 1,
                            max_step_halvings, min_micro_steps, max_error,
                            *sub1, adapt_handler);
  if (!sub2) return std::nullopt;
  if (uturn<D>(*sub1, *sub2, inv_mass)) return std::nullopt;
  return combine<Update::Barker, D>(rng, std::move(*sub1), std::move(*sub2));
}

// walnuts.hpp:520-563
template <class F, class Rand, class A>
Vec transition_w(Rand& rand, const F& logp_grad, const Vec& inv_mass,
                 const Vec& chol_mass, double step, std::size_t max_depth,
                 std::size_t max_step_halvings, std::size_t min_micro_steps,
                 double max_error, Vec&& theta, std::size_t& depth,
                 Vec& theta_grad, double& logp_pos_select,
                 A& step_size_adapter) {
  const std::size_t n = chol_mass.size();
  Vec z = rand.standard_normal(n);
  Vec rho(n);
  for (std::size_t i = 0; i < n; ++i) rho[i] = chol_mass[i] * z[i];
  Vec grad(theta.size());
  double logp_pos;
  logp_grad(theta, logp_pos, grad);
  double logp_joint = logp_pos + logp_momentum(rho, inv_mass);
  SpanW span_accum = SpanW::from_initial_point(
      std::move(theta), std::move(rho), std::move(grad), logp_pos, logp_joint);
  for (depth = 1; depth <= max_depth; ++depth) {
    bool go_forward = rand.uniform_binary();
    bool made_uturn;
    auto expand = [&](auto dir_tag) -> bool {
      constexpr Direction D = decltype(dir_tag)::value;
      auto next = build_span<D>(rand, logp_grad, inv_mass, step, depth - 1,
                                max_step_halvings, min_micro_steps, max_error,
                                span_accum, step_size_adapter);
      if (!next) return true;
      bool combined_uturn = uturn<D>(span_accum, *next, inv_mass);
      span_accum = combine<Update::Metropolis, D>(rand, std::move(span_accum),
                                                  std::move(*next));
      return combined_uturn;
    };
    made_uturn =
        go_forward
            ? expand(std::integral_constant<Direction, Direction::Forward>{})
            : expand(std::integral_constant<Direction, Direction::Backward>{});
    if (made_uturn) break;
  }
  theta_grad = span_accum.grad_select;
  logp_pos_select = span_accum.logp_pos_select;
  return std::move(span_accum.theta_select);
}

// ---------------------------------------------------------------------------
// Handler callbacks (concepts.hpp:173-245) as std::function slots.
struct ChainHandler {
  std::function<void(const Vec&, double, double, const Vec&)> on_warmup;
  std::function<void(double, const Vec&)> on_warmup_complete;
  std::function<void(const Vec&, double)> on_sample;
  std::function<void(const Vec&, const std::exception&)> on_logp_exception;
};

// walnuts.hpp:605-766
template <class F, class Rand>
class WalnutsSampler {
 public:
  WalnutsSampler(Rand rand, ChainHandler& handler, const F& logp_grad,
                 const Vec& theta, const Vec& inv_mass, double macro_time,
                 std::size_t max_nuts_depth, std::size_t max_step_halvings,
                 std::size_t min_micro_steps, double max_error,
                 uint64_t* grad_counter = nullptr)
      : rand_(std::move(rand)), handler_(&handler),
        logp_grad_(logp_grad, handler.on_logp_exception, grad_counter),
        theta_(theta), inv_mass_(inv_mass), chol_mass_(inv_mass.size()),
        macro_time_(macro_time), max_nuts_depth_(max_nuts_depth),
        max_step_halvings_(max_step_halvings),
        min_micro_steps_(min_micro_steps), max_error_(max_error) {
    for (std::size_t i = 0; i < inv_mass.size(); ++i) {
      chol_mass_[i] = 1.0 / std::sqrt(inv_mass[i]);  // :647 sqrt().inverse()
    }
    for (double v : inv_mass) {
      if (!(v > 0.0) || !std::isfinite(v)) {
        throw std::invalid_argument("inv_mass must be in (0, inf).");
      }
    }
    if (!(macro_time > 0 && !std::isinf(macro_time))) {
      throw std::invalid_argument("macro_time must be in (0, inf).");
    }
    if (max_nuts_depth == 0) {
      throw std::invalid_argument("max_nuts_depth must be in {1, 2, ... }");
    }
    if (max_step_halvings == 0) {
      throw std::invalid_argument("max_step_halvings must be in {1, 2, ... }");
    }
    if (min_micro_steps == 0) {
      throw std::invalid_argument("min_micro_steps must be in {1, 2, ... }");
    }
    if (!(max_error > 0 && !std::isinf(max_error))) {
      throw std::invalid_argument("max_error must be in (0, inf).");
    }
  }

  double operator()() {  // :682-692
    std::size_t depth;
    Vec grad_next;
    double logp_pos;
    rand_.begin_transition(iteration_);
    NoOpStepSizeAdapter noop;
    theta_ = transition_w(rand_, logp_grad_, inv_mass_, chol_mass_,
                          macro_time_, max_nuts_depth_, max_step_halvings_,
                          min_micro_steps_, max_error_, std::move(theta_),
                          depth, grad_next, logp_pos, noop);
    last_depth_ = depth;
    ++iteration_;
    if (handler_->on_sample) handler_->on_sample(theta_, logp_pos);
    return logp_pos;
  }
  std::size_t dim() const { return theta_.size(); }
  const Vec& theta() const { return theta_; }
  const Vec& inv_mass() const { return inv_mass_; }
  double macro_time() const { return macro_time_; }
  std::size_t last_depth() const { return last_depth_; }
  void set_iteration(uint32_t it) { iteration_ = it; }

 private:
  Rand rand_;
  ChainHandler* handler_;
  NoExceptLogpGrad<F> logp_grad_;
  Vec theta_, inv_mass_, chol_mass_;
  double macro_time_;
  std::size_t max_nuts_depth_, max_step_halvings_, min_micro_steps_;
  double max_error_;
  std::size_t last_depth_ = 0;
  uint32_t iteration_ = 0;  // Philox addressing only
};

// adaptive_walnuts.hpp:25-105
class MassEstimator {
 public:
  MassEstimator(const WarmupConfig& cfg, const InitChainConfig& init)
      : init_count_(cfg.mass_init_count) {
    Vec zero(init.position.size(), 0.0);
    Vec inv(init.mass.size());
    for (std::size_t i = 0; i < inv.size(); ++i) inv[i] = 1.0 / init.mass[i];
    score_ = OnlineMoments(cfg.mass_init_count, zero, init.mass);
    draw_ = OnlineMoments(cfg.mass_init_count, zero, inv);
  }
  void observe(const Vec& theta, const Vec& grad, std::size_t iteration) {
    double gamma = 1.0 - 1.0 / (init_count_ + static_cast<double>(iteration));
    draw_.discount_observe(gamma, theta);
    score_.discount_observe(gamma, grad);
  }
  Vec inv_mass_estimate() const {  // :89-94
    if (fused_arith() && draw_.weight() > 0) {
      // both estimators carry the same weight (:54-80): it cancels in the ratio
      const Vec& ds = draw_.sum_sq_dev();
      const Vec& ss = score_.sum_sq_dev();
      Vec r(ds.size());
      for (std::size_t i = 0; i < r.size(); ++i) r[i] = std::sqrt(ds[i] / ss[i]);
      return r;
    }
    Vec dv = draw_.variance(), sv = score_.variance();
    Vec r(dv.size());
    for (std::size_t i = 0; i < r.size(); ++i) r[i] = std::sqrt(dv[i] / sv[i]);
    return r;
  }

 private:
  double init_count_;
  OnlineMoments draw_, score_;
};

// adaptive_walnuts.hpp:119-164
class MinMicroStepsAdaptHandler {
 public:
  MinMicroStepsAdaptHandler(double target, std::size_t min_micro)
      : target_(target), min_micro_(min_micro), total_(2.0), count_(1.0) {}
  void observe(std::size_t macro_steps) {
    total_ += static_cast<double>(macro_steps);
    ++count_;
  }
  std::size_t min_micro_steps() const {
    double mean_micro = total_ / count_;
    double mm = mean_micro / target_;
    return std::max(min_micro_, static_cast<std::size_t>(std::lround(mm)));
  }

 private:
  const double target_;
  const std::size_t min_micro_;
  double total_, count_;
};

// adaptive_walnuts.hpp:182-363
template <class F, class Rand>
class AdaptiveWalnuts {
 public:
  AdaptiveWalnuts(Rand rand, ChainHandler& handler, const F& logp_grad,
                  const InitChainConfig& init, const WarmupConfig& warmup,
                  const SamplingConfig& sampling,
                  uint64_t* grad_counter = nullptr)
      : warmup_(warmup), sampling_(sampling), rand_(std::move(rand)),
        handler_(&handler),
        logp_grad_(logp_grad, handler.on_logp_exception, grad_counter),
        grad_counter_(grad_counter), theta_(init.position), iteration_(0),
        adam_(init.step_size, warmup.step_accept_rate_target,
              warmup.step_learning_rate, warmup.step_gradient_decay,
              warmup.step_sq_gradient_decay, warmup.step_stabilization,
              warmup.step_learn_rate_decay),
        mass_estimator_(warmup, init),
        min_micro_(warmup.max_macro_steps_target, sampling.min_micro_steps) {}

  void operator()() {  // :234-251
    Vec inv_mass = mass_estimator_.inv_mass_estimate();
    Vec chol_mass(inv_mass.size());
    for (std::size_t i = 0; i < inv_mass.size(); ++i) {
      chol_mass[i] = std::sqrt(1.0 / inv_mass[i]);  // :236 inverse().sqrt()
    }
    Vec grad_select;
    double logp_select;
    std::size_t depth;
    rand_.begin_transition(static_cast<uint32_t>(iteration_));
    theta_ = transition_w(rand_, logp_grad_, inv_mass, chol_mass,
                          adam_.step_size(), sampling_.max_trajectory_doublings,
                          sampling_.max_step_halvings,
                          min_micro_.min_micro_steps(),
                          sampling_.max_hamiltonian_error, std::move(theta_),
                          depth, grad_select, logp_select, adam_);
    mass_estimator_.observe(theta_, grad_select, iteration_);
    min_micro_.observe(static_cast<std::size_t>(1) << depth);
    last_depth_ = depth;
    last_logp_ = logp_select;
    if (handler_->on_warmup) {
      handler_->on_warmup(theta_, logp_select, step_size(), inv_mass);
    }
    ++iteration_;
  }

  // :263-271.  `make_rand` builds the sampler's Random (a fresh distribution
  // set over the same engine for StdRand; the same stream for Philox).
  template <class MakeRand>
  WalnutsSampler<F, Rand> sampler(MakeRand&& make_rand) {
    if (handler_->on_warmup_complete) {
      handler_->on_warmup_complete(step_size(), inv_mass());
    }
    return WalnutsSampler<F, Rand>(
        make_rand(rand_), *handler_, logp_grad_.base(), theta_, inv_mass(),
        step_size(), sampling_.max_trajectory_doublings,
        sampling_.max_step_halvings, min_micro_.min_micro_steps(),
        sampling_.max_hamiltonian_error, grad_counter_);
  }

  Vec inv_mass() const { return mass_estimator_.inv_mass_estimate(); }
  double step_size() const { return adam_.step_size(); }
  std::size_t min_micro_steps() const { return min_micro_.min_micro_steps(); }
  std::size_t dim() const { return theta_.size(); }
  double log_step_size() const { return std::log(step_size()); }  // :312
  Vec log_mass() const {                                          // :320-323
    Vec im = inv_mass();
    for (double& v : im) v = -std::log(v);
    return im;
  }
  std::size_t iter() const { return iteration_; }
  const Vec& theta() const { return theta_; }
  std::size_t last_depth() const { return last_depth_; }
  double last_logp() const { return last_logp_; }

 private:
  const WarmupConfig& warmup_;
  const SamplingConfig& sampling_;
  Rand rand_;
  ChainHandler* handler_;
  NoExceptLogpGrad<F> logp_grad_;
  uint64_t* grad_counter_;
  Vec theta_;
  std::size_t iteration_;
  Adam adam_;
  MassEstimator mass_estimator_;
  MinMicroStepsAdaptHandler min_micro_;
  std::size_t last_depth_ = 0;
  double last_logp_ = 0;
};

// ---------------------------------------------------------------------------
// Cross-chain controllers, as pure functions of the published snapshots.

struct AdaptSnapshot {  // adapt.hpp:26-54
  std::size_t iter = 0;
  double log_step = std::numeric_limits<double>::quiet_NaN();
  Vec log_mass, mass;
};

template <class A>
AdaptSnapshot make_snapshot(const A& adapter, std::size_t iter) {
  AdaptSnapshot s;  // adapt.hpp:132-139
  s.iter = iter;
  s.log_step = adapter.log_step_size();
  s.log_mass = adapter.log_mass();
  s.mass.resize(s.log_mass.size());
  for (std::size_t i = 0; i < s.mass.size(); ++i) {
    s.mass[i] = std::exp(s.log_mass[i]);
  }
  return s;
}

// one evaluation of the body of adapt.hpp:186-224; true = stop warm-up
inline bool warmup_should_stop(const std::vector<AdaptSnapshot>& latest,
                               const WarmupConfig& cfg,
                               double* max_rel_mass = nullptr,
                               double* max_rel_step = nullptr) {
  const std::size_t M = latest.size();
  const std::size_t D = latest.empty() ? 0 : latest[0].log_mass.size();
  std::size_t num_draws = 0;
  Vec mean_log_mass(D, 0.0);
  double mean_log_step = 0.0;
  for (std::size_t m = 0; m < M; ++m) {
    if (latest[m].iter < cfg.min_iter) return false;
    num_draws += latest[m].iter;
    mean_log_step += latest[m].log_step;
    for (std::size_t d = 0; d < D; ++d) mean_log_mass[d] += latest[m].log_mass[d];
  }
  mean_log_step /= static_cast<double>(M);
  Vec geom(D);
  for (std::size_t d = 0; d < D; ++d) {
    mean_log_mass[d] /= static_cast<double>(M);
    geom[d] = std::exp(mean_log_mass[d]);
  }
  double max_mass = 0.0, max_step = 0.0;
  double geom_step = std::exp(mean_log_step);
  for (std::size_t m = 0; m < M; ++m) {
    max_mass = std::fmax(max_mass, l2_rel_diff(latest[m].mass, geom));
    double s = std::exp(latest[m].log_step);
    max_step = std::fmax(max_step, (s - geom_step) / geom_step);
  }
  if (max_rel_mass) *max_rel_mass = max_mass;
  if (max_rel_step) *max_rel_step = max_step;
  bool converged = max_mass <= cfg.mass_converge_tol &&
                   max_step <= cfg.step_size_converge_tol;
  bool hit_max = num_draws == M * cfg.max_iter;
  return converged || hit_max;
}

struct ChainStats {  // sampler.hpp:30-39
  double sample_mean, sample_var;
  std::size_t count;
};

// one evaluation of the body of sampler.hpp:128-152; r_hat is NaN-safe as in
// the reference (single chain -> NaN -> never converges)
inline bool sampling_should_stop(const std::vector<ChainStats>& stats,
                                 const SamplingConfig& cfg, double* r_hat_out,
                                 bool* evaluated) {
  const std::size_t M = stats.size();
  std::size_t num_draws = 0;
  Vec means(M), vars(M);
  if (evaluated) *evaluated = false;
  for (std::size_t m = 0; m < M; ++m) {
    if (stats[m].count < cfg.min_iter) return false;
    num_draws += stats[m].count;
    means[m] = stats[m].sample_mean;
    vars[m] = stats[m].sample_var;
  }
  double var_of_means = variance(means);
  double mean_of_vars = 0.0;
  for (double v : vars) mean_of_vars += v;
  mean_of_vars /= static_cast<double>(M);
  double r_hat = std::sqrt(1 + var_of_means / mean_of_vars);
  if (r_hat_out) *r_hat_out = r_hat;
  if (evaluated) *evaluated = true;
  bool converged = r_hat <= cfg.rhat_converge_tol;
  bool hit_max = num_draws == M * cfg.max_iter;
  return converged || hit_max;
}

}  // namespace oracle
