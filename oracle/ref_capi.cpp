// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// The `ref_*` entry points of oracle_capi.h over the UNMODIFIED reference
// headers (-I/root/reference/include), compiled against oracle/eigen_shim.
// Nothing here re-implements the algorithm: it only drives
// walnutpie::AdaptiveWalnuts / WalnutsSampler / InitConfigBuilder /
// walnutpie::walnuts exactly as python/src/walnutpie/walnutpy.cpp:20-84 and
// examples/walnutpie_api.cpp do.  Output: oracle/_ref/libwalnuts_ref.so.
#include "oracle_capi.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include <Eigen/Dense>

#include <walnutpie/adaptive_walnuts.hpp>
#include <walnutpie/api.hpp>
#include <walnutpie/config.hpp>
#include <walnutpie/util.hpp>
#include <walnutpie/walnuts.hpp>

#include "targets.hpp"

namespace {
using Eigen::VectorXd;

thread_local std::string g_err;

template <class Fn>
int guarded(Fn&& fn) {
  try {
    fn();
    return 0;
  } catch (const std::invalid_argument& e) {
    g_err = e.what();
    return -2;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Adapts the std::vector targets of oracle/targets.hpp (not part of the
// reference) to the reference's LogpGrad shape, so both checkers evaluate the
// very same density arithmetic; counts evaluations.
template <class T>
struct EigenTarget {
  T base;
  std::atomic<uint64_t>* counter = nullptr;
  void operator()(const VectorXd& x, double& lp, VectorXd& grad) const {
    if (counter) counter->fetch_add(1, std::memory_order_relaxed);
    grad.resize(x.size());
    base.eval(x.data(), static_cast<std::size_t>(x.size()), lp, grad.data());
  }
};

template <class Fn>
void with_target(const OracleTarget* t, std::atomic<uint64_t>* counter, Fn&& fn) {
  switch (t->kind) {
    case 0: fn(EigenTarget<oracle::StdNormal>{{}, counter}); break;
    case 1: {
      oracle::DiagGaussian g;
      const double* p = static_cast<const double*>(t->data0);
      g.prec.assign(p, p + t->D);
      fn(EigenTarget<oracle::DiagGaussian>{g, counter});
      break;
    }
    case 2: fn(EigenTarget<oracle::Funnel>{{}, counter}); break;
    case 3: {
      oracle::Logistic l;
      l.N = t->N; l.D = t->D;
      l.X = static_cast<const double*>(t->data0);
      l.y = static_cast<const double*>(t->data1);
      fn(EigenTarget<oracle::Logistic>{l, counter});
      break;
    }
    case 4: {
      oracle::CFuncTarget c;
      c.fn = reinterpret_cast<oracle::LOGP_CFUNC>(const_cast<void*>(t->data0));
      c.data = const_cast<void*>(t->data1);
      fn(EigenTarget<oracle::CFuncTarget>{c, counter});
      break;
    }
    default: throw std::invalid_argument("unknown target kind");
  }
}

VectorXd to_eigen(const double* p, size_t n) {
  VectorXd v(static_cast<Eigen::Index>(n));
  std::memcpy(v.data(), p, n * 8);
  return v;
}

walnutpie::WarmupConfig make_warmup(const OracleConfig& c) {
  auto b = walnutpie::WarmupConfigBuilder()
      .min_max_iter(c.min_warmup_iter, c.max_warmup_iter)
      .step_size_converge_tol(c.step_size_converge_tol)
      .mass_converge_tol(c.mass_converge_tol)
      .mass_init_count(c.mass_init_count)
      .mass_additive_smoothing(c.mass_additive_smoothing)
      .max_macro_steps_target(c.max_macro_steps_target)
      .step_accept_rate_target(c.step_accept_rate_target)
      .step_learning_rate(c.step_learning_rate)
      .step_gradient_decay(c.step_gradient_decay)
      .step_sq_gradient_decay(c.step_sq_gradient_decay)
      .step_stabilization(c.step_stabilization)
      .step_learn_rate_decay(c.step_learn_rate_decay);
  if (c.publish_stride > 0) b.publish_stride(c.publish_stride);
  return b.build();
}

walnutpie::SamplingConfig make_sampling(const OracleConfig& c) {
  return walnutpie::SamplingConfigBuilder()
      .min_max_iter(c.min_sampling_iter, c.max_sampling_iter)
      .rhat_converge_tol(c.rhat_converge_tol)
      .max_trajectory_doublings(c.max_trajectory_doublings)
      .max_step_halvings(c.max_step_halvings)
      .max_hamiltonian_error(c.max_hamiltonian_error)
      .min_micro_steps(c.min_micro_steps)
      .build();
}

// ChainHandler (concepts.hpp:236-245) writing into caller buffers
struct RecordingHandler {
  size_t D = 0;
  double* warmup_draws = nullptr; double* warmup_lp = nullptr;
  double* warmup_step = nullptr; double* warmup_inv_mass = nullptr;
  double* draws = nullptr; double* lp = nullptr;
  double* inv_mass_out = nullptr; double* step_out = nullptr;
  size_t wi = 0, si = 0;
  std::chrono::steady_clock::time_point warmup_done{};

  void on_sample(const VectorXd& th, double l) {
    if (draws) std::memcpy(draws + si * D, th.data(), D * 8);
    if (lp) lp[si] = l;
    ++si;
  }
  void on_warmup(const VectorXd& th, double l, double step, const VectorXd& im) {
    if (warmup_draws) std::memcpy(warmup_draws + wi * D, th.data(), D * 8);
    if (warmup_lp) warmup_lp[wi] = l;
    if (warmup_step) warmup_step[wi] = step;
    if (warmup_inv_mass) std::memcpy(warmup_inv_mass + wi * D, im.data(), D * 8);
    ++wi;
  }
  void on_warmup_complete(double step, const VectorXd& im) {
    if (step_out) *step_out = step;
    if (inv_mass_out) std::memcpy(inv_mass_out, im.data(), D * 8);
    warmup_done = std::chrono::steady_clock::now();
  }
  void on_logp_exception(const VectorXd&, const std::exception&) noexcept {}
};

struct NoGlobal { void on_r_hat(double) {} };
struct NoInterrupt { void throw_if_interrupted() const {} };

}  // namespace

extern "C" {

const char* ref_last_error(void) { return g_err.c_str(); }

int ref_run_chain(const OracleTarget* target, const OracleConfig* cfg,
                  uint32_t seed, uint32_t chain, int rng_policy,
                  const double* theta0, const double* mass0, double step0,
                  int n_warmup, int n_sampling, double* warmup_draws,
                  double* warmup_lp, double* warmup_step,
                  double* warmup_inv_mass, int* warmup_depth, double* draws,
                  double* lp, int* depth, double* inv_mass_out,
                  double* step_out, int* min_micro_out, uint64_t* grad_evals) {
  return guarded([&] {
    if (rng_policy != 0) throw std::invalid_argument("ref: mt19937_64 only");
    const size_t D = target->D;
    auto w = make_warmup(*cfg);
    auto s = make_sampling(*cfg);
    std::atomic<uint64_t> counter{0};
    with_target(target, &counter, [&](const auto& f) {
      using F = std::decay_t<decltype(f)>;
      std::seed_seq ss{static_cast<std::size_t>(seed),
                       static_cast<std::size_t>(chain) + 1u};  // api.hpp:49
      std::mt19937_64 rng(ss);
      RecordingHandler h;
      h.D = D;
      h.warmup_draws = warmup_draws; h.warmup_lp = warmup_lp;
      h.warmup_step = warmup_step; h.warmup_inv_mass = warmup_inv_mass;
      h.draws = draws; h.lp = lp;
      h.inv_mass_out = inv_mass_out; h.step_out = step_out;
      walnutpie::InitChainConfig init(step0, to_eigen(theta0, D),
                                      to_eigen(mass0, D));
      walnutpie::AdaptiveWalnuts<F, std::mt19937_64, RecordingHandler> adapter(
          rng, h, f, init, w, s);
      for (int n = 0; n < n_warmup; ++n) adapter();
      auto sampler = adapter.sampler();
      if (min_micro_out) *min_micro_out = static_cast<int>(adapter.min_micro_steps());
      for (int n = 0; n < n_sampling; ++n) sampler();
      (void)warmup_depth; (void)depth;  // not observable through the public API
    });
    if (grad_evals) *grad_evals = counter.load();
  });
}

int ref_run_sampler(const OracleTarget* target, uint32_t seed, uint32_t chain,
                    int rng_policy, uint32_t first_iter, const double* theta0,
                    const double* inv_mass, double step, int max_depth,
                    int max_halvings, int min_micro, double max_error,
                    int n_iter, double* draws, double* lp, int* depth,
                    uint64_t* grad_evals) {
  return guarded([&] {
    if (rng_policy != 0) throw std::invalid_argument("ref: mt19937_64 only");
    (void)first_iter; (void)depth;
    const size_t D = target->D;
    std::atomic<uint64_t> counter{0};
    with_target(target, &counter, [&](const auto& f) {
      using F = std::decay_t<decltype(f)>;
      std::seed_seq ss{static_cast<std::size_t>(seed),
                       static_cast<std::size_t>(chain) + 1u};
      std::mt19937_64 rng(ss);
      RecordingHandler h;
      h.D = D; h.draws = draws; h.lp = lp;
      walnutpie::WalnutsSampler<F, std::mt19937_64, RecordingHandler> sampler(
          rng, h, f, to_eigen(theta0, D), to_eigen(inv_mass, D), step,
          max_depth, max_halvings, min_micro, max_error);
      for (int n = 0; n < n_iter; ++n) sampler();
    });
    if (grad_evals) *grad_evals = counter.load();
  });
}

int ref_init_positions(size_t num_chains, size_t D, uint32_t seed,
                       double radius, double* positions) {
  return guarded([&] {
    std::seed_seq ss{seed, 1u};  // walnutpy.cpp:187-189
    std::mt19937_64 rng(ss);
    auto cfg = walnutpie::InitConfigBuilder{num_chains, D}
                   .positions(rng, radius)
                   .build();
    for (size_t c = 0; c < num_chains; ++c) {
      std::memcpy(positions + c * D, cfg.position(c).data(), D * 8);
    }
  });
}

int ref_init_mass_step(const OracleTarget* target, size_t num_chains,
                       uint32_t seed, const double* positions,
                       const double* mass_in, double smoothing,
                       double step_init, double* mass_out, double* step_out) {
  return guarded([&] {
    const size_t D = target->D;
    with_target(target, nullptr, [&](const auto& f) {
      std::vector<VectorXd> pos(num_chains);
      for (size_t c = 0; c < num_chains; ++c) pos[c] = to_eigen(positions + c * D, D);
      auto builder = walnutpie::InitConfigBuilder{num_chains, D}
                         .step_sizes(step_init)
                         .positions(pos);
      if (mass_in) {  // walnutpy.cpp:64-70
        std::vector<VectorXd> m(num_chains);
        for (size_t c = 0; c < num_chains; ++c) m[c] = to_eigen(mass_in + c * D, D);
        builder.masses(m);
      } else {        // walnutpy.cpp:72
        builder.masses(f, smoothing);
      }
      std::seed_seq ss{seed, 2u};  // walnutpy.cpp:75-76
      std::mt19937_64 init_rng(ss);
      auto cfg = builder.adapt_step_build(init_rng, f);
      for (size_t c = 0; c < num_chains; ++c) {
        std::memcpy(mass_out + c * D, cfg.mass(c).data(), D * 8);
        step_out[c] = cfg.step_size(c);
      }
    });
  });
}

double ref_leapfrog_error(const OracleTarget* target, const double* theta,
                          const double* rho, const double* inv_mass,
                          double step) {
  double out = 0;
  const size_t D = target->D;
  with_target(target, nullptr, [&](const auto& f) {
    out = walnutpie::detail::leapfrog_error(f, to_eigen(theta, D),
                                            to_eigen(rho, D),
                                            to_eigen(inv_mass, D), step);
  });
  return out;
}

double ref_log_sum_exp(double a, double b) {
  return walnutpie::detail::log_sum_exp(a, b);
}

double ref_logp_momentum(const double* rho, const double* inv_mass, size_t D) {
  return walnutpie::detail::logp_momentum(to_eigen(rho, D), to_eigen(inv_mass, D));
}

int ref_walnuts(const OracleTarget* target, const OracleConfig* cfg,
                size_t num_chains, uint32_t seed, const double* positions,
                const double* mass, const double* steps, int save_warmup,
                double* out, int* final_lengths, double* stepsize_out,
                double* inv_metric_out, uint64_t* grad_evals,
                double* seconds_warmup, double* seconds_sampling) {
  // the reference's own multi-threaded driver, api.hpp:33-69
  return guarded([&] {
    const size_t D = target->D, C = num_chains;
    auto w = make_warmup(*cfg);
    auto s = make_sampling(*cfg);
    const size_t stride_out = D * (s.max_iter() + (save_warmup ? w.max_iter() : 0));
    std::atomic<uint64_t> counter{0};
    with_target(target, &counter, [&](const auto& f) {
      std::vector<VectorXd> pos(C), m(C);
      std::vector<double> st(steps, steps + C);
      for (size_t c = 0; c < C; ++c) {
        pos[c] = to_eigen(positions + c * D, D);
        m[c] = to_eigen(mass + c * D, D);
      }
      walnutpie::WalnutsConfig config{
          walnutpie::InitConfigBuilder{C, D}.step_sizes(st).positions(pos).masses(m).build(),
          w, s};
      // BufferHandler semantics (python/src/walnutpie/handlers.hpp:63-116)
      struct Buf : RecordingHandler {
        bool save_warmup = false;
        double* base = nullptr;
        size_t written = 0, written_warmup = 0;
        void on_sample(const VectorXd& th, double) {
          if (base) std::memcpy(base + written * D, th.data(), D * 8);
          ++written;
        }
        void on_warmup(const VectorXd& th, double, double, const VectorXd&) {
          if (save_warmup) {
            if (base) std::memcpy(base + written * D, th.data(), D * 8);
            ++written; ++written_warmup;
          }
        }
      };
      std::vector<Buf> handlers(C);
      for (size_t c = 0; c < C; ++c) {
        handlers[c].D = D;
        handlers[c].save_warmup = save_warmup != 0;
        handlers[c].base = out ? out + stride_out * c : nullptr;
        handlers[c].step_out = stepsize_out ? stepsize_out + c : nullptr;
        handlers[c].inv_mass_out = inv_metric_out ? inv_metric_out + c * D : nullptr;
      }
      NoGlobal global;
      NoInterrupt interrupt;
      auto t0 = std::chrono::steady_clock::now();
      walnutpie::walnuts<std::mt19937_64>(seed, handlers, global, interrupt, f,
                                          config);
      auto t2 = std::chrono::steady_clock::now();
      auto t1 = t0;
      for (auto& h : handlers) t1 = std::max(t1, h.warmup_done);
      if (seconds_warmup) *seconds_warmup = std::chrono::duration<double>(t1 - t0).count();
      if (seconds_sampling) *seconds_sampling = std::chrono::duration<double>(t2 - t1).count();
      for (size_t c = 0; c < C; ++c) {
        if (final_lengths) {
          final_lengths[c] = static_cast<int>(handlers[c].written_warmup);
          final_lengths[c + C] =
              static_cast<int>(handlers[c].written - handlers[c].written_warmup);
        }
      }
    });
    if (grad_evals) *grad_evals = counter.load();
  });
}

}  // extern "C"
