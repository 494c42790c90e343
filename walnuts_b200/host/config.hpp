// Host-side configuration objects of the B200 sampler.
//
// Mirrors the reference's plugin-facing configuration API so that callers and
// tests read the same: include/walnutpie/config.hpp — InitConfig(:74-185) /
// InitConfigBuilder(:195-484), WarmupConfig(:513-641) / WarmupConfigBuilder
// (:646-850), SamplingConfig(:885-954) / SamplingConfigBuilder(:967-1059),
// WalnutsConfig(:1089-1137) — and the argument validators of validate.hpp
// (error strings are part of the contract: python/tests/test_pyfunc.py:68).
// Same names, same defaults, same messages; vectors are std::vector<double>
// (batched [C][D] on the device side), there is no Eigen.
#pragma once
#include <cmath>
#include <cstddef>
#include <ostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace walnuts_b200 {

using Vector = std::vector<double>;

namespace validate {
// validate.hpp:27-285
inline void finite_positive(double x, const std::string& var) {
  if (!(std::isfinite(x) && x > 0)) {
    throw std::invalid_argument(var + " must be finite and > 0");
  }
}
inline void finite_positive(const Vector& xs, const std::string& var) {
  for (double x : xs) finite_positive(x, var);
}
inline void finite_positive(const std::vector<Vector>& xs, const std::string& var) {
  for (const auto& x : xs) finite_positive(x, var);
}
inline void finite(const Vector& xs, const std::string& var) {
  for (double x : xs) {
    if (!std::isfinite(x)) throw std::invalid_argument(var + " must be finite");
  }
}
inline void finite(const std::vector<Vector>& xs, const std::string& var) {
  for (const auto& x : xs) finite(x, var);
}
inline void finite_gt1(double x, const std::string& var) {
  if (!(std::isfinite(x) && x > 1)) {
    throw std::invalid_argument(var + " must be finite and > 1");
  }
}
inline void positive(double x, const std::string& name) {
  if (!(x > 0 && !std::isinf(x))) {
    throw std::invalid_argument(name + " must be in (0, inf).");
  }
}
inline void positive(std::size_t x, const std::string& name) {
  if (!(x > 0)) throw std::invalid_argument(name + " must be in {1, 2, ... }");
}
inline void probability(double x, const std::string& name) {
  if (!(x > 0 && x < 1)) throw std::invalid_argument(name + " must be in (0, 1)");
}
inline void probability_inclusive(double x, const std::string& name) {
  if (!(x >= 0 && x <= 1)) throw std::invalid_argument(name + " must be in [0, 1]");
}
template <class T>
inline void size(const std::vector<T>& x, std::size_t n, const std::string& var,
                 const std::string& target) {
  if (x.size() != n) throw std::invalid_argument(var + " size must match " + target);
}
}  // namespace validate

// ---------------------------------------------------------------- init ----
struct InitChainConfig {  // config.hpp:25-63
  double step_size;
  Vector position;
  Vector mass;
};

class InitConfig {  // config.hpp:74-185
 public:
  std::size_t num_chains() const noexcept { return step_sizes_.size(); }
  std::size_t dims() const noexcept {
    return positions_.empty() ? 0u : positions_.front().size();
  }
  const std::vector<double>& step_sizes() const noexcept { return step_sizes_; }
  double step_size(std::size_t n) const noexcept { return step_sizes_[n]; }
  const std::vector<Vector>& positions() const noexcept { return positions_; }
  const Vector& position(std::size_t n) const noexcept { return positions_[n]; }
  const std::vector<Vector>& masses() const noexcept { return masses_; }
  const Vector& mass(std::size_t n) const noexcept { return masses_[n]; }
  InitChainConfig init_chain_config(std::size_t n) const {
    return {step_size(n), position(n), mass(n)};
  }

 private:
  friend class InitConfigBuilder;
  InitConfig() = default;
  std::vector<double> step_sizes_;
  std::vector<Vector> positions_;
  std::vector<Vector> masses_;
};

class InitConfigBuilder {  // config.hpp:195-484
 public:
  InitConfigBuilder(std::size_t num_chains, std::size_t dims)
      : num_chains_(num_chains), dims_(dims) {
    cfg_.step_sizes_.assign(num_chains, 0.1);            // :206
    cfg_.positions_.assign(num_chains, Vector(dims, 0.0));
    cfg_.masses_.assign(num_chains, Vector(dims, 1.0));
  }
  InitConfigBuilder& step_sizes(double v) {
    validate::finite_positive(v, "step size");
    cfg_.step_sizes_.assign(num_chains_, v);
    return *this;
  }
  InitConfigBuilder& step_sizes(const std::vector<double>& v) {
    validate::size(v, num_chains_, "step_sizes", "num_chains");
    validate::finite_positive(v, "step_size");
    cfg_.step_sizes_ = v;
    return *this;
  }
  InitConfigBuilder& positions(const Vector& v) {
    validate::size(v, dims_, "position", "dims");
    validate::finite(v, "position");
    cfg_.positions_.assign(num_chains_, v);
    return *this;
  }
  InitConfigBuilder& positions(const std::vector<Vector>& vs) {
    validate::size(vs, num_chains_, "positions", "num_chains");
    validate::finite(vs, "positions");
    for (const auto& v : vs) validate::size(v, dims_, "position", "dims");
    cfg_.positions_ = vs;
    return *this;
  }
  InitConfigBuilder& masses(const Vector& v) {
    validate::size(v, dims_, "masses", "dims");
    validate::finite_positive(v, "masses");
    cfg_.masses_.assign(num_chains_, v);
    return *this;
  }
  InitConfigBuilder& masses(const std::vector<Vector>& vs) {
    validate::size(vs, num_chains_, "masses", "num_chains");
    validate::finite_positive(vs, "masses");
    for (const auto& v : vs) validate::size(v, dims_, "all masses", "dims");
    cfg_.masses_ = vs;
    return *this;
  }
  // positions(rng, scale), masses(logp_grad, s) and adapt_step_build(rng, F)
  // (config.hpp:259-268, :360-382, :469-476) run on the device for a batch:
  // see wb200_session_init in include/walnuts_b200.h.
  InitConfig build() { return std::move(cfg_); }

 private:
  std::size_t num_chains_, dims_;
  InitConfig cfg_;
};

// -------------------------------------------------------------- warm-up ----
class WarmupConfig {  // config.hpp:513-641, defaults :626-640
 public:
  std::size_t min_iter() const { return min_iter_; }
  std::size_t max_iter() const { return max_iter_; }
  double step_size_converge_tol() const { return step_size_converge_tol_; }
  double mass_converge_tol() const { return mass_converge_tol_; }
  double mass_init_count() const { return mass_init_count_; }
  double mass_additive_smoothing() const { return mass_additive_smoothing_; }
  double max_macro_steps_target() const { return max_macro_steps_target_; }
  double step_accept_rate_target() const { return step_accept_rate_target_; }
  double step_learning_rate() const { return step_learning_rate_; }
  double step_gradient_decay() const { return step_gradient_decay_; }
  double step_sq_gradient_decay() const { return step_sq_gradient_decay_; }
  double step_stabilization() const { return step_stabilization_; }
  double step_learn_rate_decay() const { return step_learn_rate_decay_; }
  std::size_t publish_stride() const { return publish_stride_; }
  std::size_t yield_period() const { return yield_period_; }

 private:
  friend class WarmupConfigBuilder;
  WarmupConfig() = default;
  std::size_t min_iter_ = 50, max_iter_ = 1000;
  double step_size_converge_tol_ = 0.1, mass_converge_tol_ = 1.0;
  double mass_init_count_ = 4.0, mass_additive_smoothing_ = 1e-5;
  double max_macro_steps_target_ = 15.0;
  double step_accept_rate_target_ = 0.8, step_learning_rate_ = 0.05;
  double step_gradient_decay_ = 0.8, step_sq_gradient_decay_ = 0.9;
  double step_stabilization_ = 1e-4, step_learn_rate_decay_ = 0.5;
  std::size_t publish_stride_ = 5, yield_period_ = 32;
};

class WarmupConfigBuilder {  // config.hpp:646-850
 public:
  WarmupConfigBuilder& min_max_iter(std::size_t min_iter, std::size_t max_iter) {
    if (min_iter > max_iter) {
      throw std::invalid_argument("min_iter cannot be greater than than max_iter");
    }
    cfg_.min_iter_ = min_iter;
    cfg_.max_iter_ = max_iter;
    return *this;
  }
#define WB200_SETTER(NAME, CHECK, LABEL)                 \
  WarmupConfigBuilder& NAME(double v) {                  \
    validate::CHECK(v, LABEL);                           \
    cfg_.NAME##_ = v;                                    \
    return *this;                                        \
  }
  WB200_SETTER(step_size_converge_tol, finite_positive, "step_size_converge_tol")
  WB200_SETTER(mass_converge_tol, finite_positive, "mass_converge_tol")
  WB200_SETTER(mass_init_count, finite_positive, "mass_init_count")
  WB200_SETTER(mass_additive_smoothing, finite_positive, "mass_additive_smoothing")
  WB200_SETTER(max_macro_steps_target, finite_positive, "max_macro_steps_target")
  WB200_SETTER(step_accept_rate_target, probability, "step_accept_rate_target")
  WB200_SETTER(step_learning_rate, finite_positive, "step_learning_rate")
  WB200_SETTER(step_gradient_decay, probability, "step_gradient_decay")
  WB200_SETTER(step_sq_gradient_decay, probability, "step_sq_gradient_decay")
  WB200_SETTER(step_stabilization, finite_positive, "step_stabilization")
  WB200_SETTER(step_learn_rate_decay, probability, "step_learn_rate_decay")
#undef WB200_SETTER
  WarmupConfigBuilder& publish_stride(std::size_t v) {
    validate::positive(v, "publish_stride");
    cfg_.publish_stride_ = v;
    return *this;
  }
  WarmupConfigBuilder& yield_period(std::size_t v) {
    validate::positive(v, "yield_period");
    cfg_.yield_period_ = v;
    return *this;
  }
  WarmupConfig build() { return cfg_; }

 private:
  WarmupConfig cfg_;
};

// ------------------------------------------------------------- sampling ----
class SamplingConfig {  // config.hpp:885-954, defaults :947-953
 public:
  std::size_t min_iter() const noexcept { return min_iter_; }
  std::size_t max_iter() const noexcept { return max_iter_; }
  std::size_t max_trajectory_doublings() const noexcept { return max_trajectory_doublings_; }
  std::size_t max_step_halvings() const noexcept { return max_step_halvings_; }
  double max_hamiltonian_error() const noexcept { return max_hamiltonian_error_; }
  std::size_t min_micro_steps() const noexcept { return min_micro_steps_; }
  double rhat_converge_tol() const noexcept { return rhat_converge_tol_; }

 private:
  friend class SamplingConfigBuilder;
  SamplingConfig() = default;
  std::size_t min_iter_ = 50, max_iter_ = 1000;
  std::size_t max_trajectory_doublings_ = 5, max_step_halvings_ = 5;
  double max_hamiltonian_error_ = 0.5;
  std::size_t min_micro_steps_ = 1;
  double rhat_converge_tol_ = 1.01;
};

class SamplingConfigBuilder {  // config.hpp:967-1059
 public:
  SamplingConfigBuilder& min_max_iter(std::size_t min_iter, std::size_t max_iter) {
    if (min_iter > max_iter) {
      throw std::invalid_argument("min_iter must be <= max_iter");
    }
    cfg_.min_iter_ = min_iter;
    cfg_.max_iter_ = max_iter;
    return *this;
  }
  SamplingConfigBuilder& max_trajectory_doublings(std::size_t v) noexcept {
    cfg_.max_trajectory_doublings_ = v;
    return *this;
  }
  SamplingConfigBuilder& max_step_halvings(std::size_t v) noexcept {
    cfg_.max_step_halvings_ = v;
    return *this;
  }
  SamplingConfigBuilder& max_hamiltonian_error(double v) {
    validate::finite_positive(v, "max_hamiltonian_error");
    cfg_.max_hamiltonian_error_ = v;
    return *this;
  }
  SamplingConfigBuilder& min_micro_steps(std::size_t v) {
    validate::positive(v, "min_micro_steps");
    cfg_.min_micro_steps_ = v;
    return *this;
  }
  SamplingConfigBuilder& rhat_converge_tol(double v) {
    validate::finite_gt1(v, "rhat_convergence_tol");
    cfg_.rhat_converge_tol_ = v;
    return *this;
  }
  SamplingConfig build() { return cfg_; }

 private:
  SamplingConfig cfg_;
};

class WalnutsConfig {  // config.hpp:1089-1137
 public:
  WalnutsConfig(InitConfig init, WarmupConfig warmup, SamplingConfig sampling)
      : init_(std::move(init)), warmup_(std::move(warmup)),
        sampling_(std::move(sampling)) {}
  const InitConfig& init() const noexcept { return init_; }
  const WarmupConfig& warmup() const noexcept { return warmup_; }
  const SamplingConfig& sampling() const noexcept { return sampling_; }

 private:
  InitConfig init_;
  WarmupConfig warmup_;
  SamplingConfig sampling_;
};

inline std::ostream& operator<<(std::ostream& out, const WarmupConfig& c) {
  return out << "WarmupConfig\n"
             << "  min_iter                 = " << c.min_iter() << "\n"
             << "  max_iter                 = " << c.max_iter() << "\n"
             << "  step_size_converge_tol   = " << c.step_size_converge_tol() << "\n"
             << "  mass_converge_tol        = " << c.mass_converge_tol() << "\n"
             << "  mass_init_count          = " << c.mass_init_count() << "\n"
             << "  mass_additive_smoothing  = " << c.mass_additive_smoothing() << "\n"
             << "  max_macro_steps_target   = " << c.max_macro_steps_target() << "\n"
             << "  step_accept_rate_target  = " << c.step_accept_rate_target() << "\n"
             << "  step_learning_rate       = " << c.step_learning_rate() << "\n"
             << "  step_gradient_decay      = " << c.step_gradient_decay() << "\n"
             << "  step_sq_gradient_decay   = " << c.step_sq_gradient_decay() << "\n"
             << "  step_stabilization       = " << c.step_stabilization() << "\n"
             << "  step_learn_rate_decay    = " << c.step_learn_rate_decay() << "\n"
             << "  publish_stride           = " << c.publish_stride() << "\n"
             << "  yield_period             = " << c.yield_period() << "\n";
}

inline std::ostream& operator<<(std::ostream& out, const SamplingConfig& c) {
  return out << "SamplingConfig\n"
             << "  min_iter                   = " << c.min_iter() << "\n"
             << "  max_iter                   = " << c.max_iter() << "\n"
             << "  max_trajectory_doublings   = " << c.max_trajectory_doublings() << "\n"
             << "  max_step_halvings          = " << c.max_step_halvings() << "\n"
             << "  max_hamiltonian_error      = " << c.max_hamiltonian_error() << "\n"
             << "  min_micro_steps            = " << c.min_micro_steps() << "\n"
             << "  rhat_converge_tol          = " << c.rhat_converge_tol() << "\n";
}

}  // namespace walnuts_b200
