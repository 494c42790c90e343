// walnuts_b200::walnuts — the C++ entry point of the B200 sampler.
//
// Mirrors the reference's top-level driver (include/walnutpie/api.hpp:33-69 with
// detail::adapt, adapt.hpp:173-259, and detail::sample, sampler.hpp:118-192):
//
//   walnutpie::walnuts<RNG>(seed, chain_handlers, global_handler, interrupt, logp_grad, config)
//
// with the density named by a WalnutModelDesc (include/walnuts_b200.h) instead of a host
// callable, and the per-chain threads replaced by one device-resident batch.  Handlers
// keep the reference's interface (concepts.hpp:174-245) with std::vector<double> in place
// of Eigen::VectorXd:
//
//   chain handler    on_warmup(position, lp, step_size, diag_inv_mass)
//                    on_warmup_complete(step_size, diag_inv_mass)
//                    on_sample(position, lp)
//                    on_logp_exception(position, exception) noexcept   [optional]
//   global handler   on_r_hat(r_hat)
//   interrupt        throw_if_interrupted()
//
// on_logp_exception (concepts.hpp:196-201, util.hpp:336-346): a batched device density
// (model kind 4) that returns non-zero inside a transition gives every chain of that
// evaluation logp = -inf and a zero gradient, and sampling goes on; each chain handler
// that has the method is told once per failed evaluation, with the chain's latest draw.
//
// The batch runs in blocks of WarmupConfig::publish_stride iterations; after each block
// the handlers of every chain receive that block's iterations in order, then the
// controllers (warm-up convergence, R-hat of lp) decide for all chains at once.
// Header-only over the C ABI: link with libwalnuts_b200.so.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/walnuts_b200.h"
#include "config.hpp"

namespace walnuts_b200 {

// How the session is initialised.  The reference's builder methods that need the density
// or a generator (InitConfigBuilder::positions(rng, scale), masses(logp_grad, s),
// adapt_step_build(rng, F); config.hpp:259-268, :360-382, :469-476) run on the device:
// set the flag and the corresponding part of InitConfig is ignored.
struct DeviceInit {
  bool random_positions = false;  // N(0, init_scale^2) instead of InitConfig::positions()
  double init_scale = 2.0;
  bool gradient_masses = false;   // (1 - s)|grad| + s instead of InitConfig::masses()
  bool adapt_steps = false;       // initial step search instead of InitConfig::step_sizes()
  unsigned int chain_offset = 0;  // global id of chain 0 (multi-GPU sharding)
  int device = 0;
};

namespace detail {

// `e` by reference: check(f(..., &e), e) must read the error after the call has set it
inline void check(int rc, WalnutpyError*& e) {  // errors.hpp:42-72 in reverse
  if (rc == 0) return;
  std::string msg = e ? walnutpie_get_error_message(e) : "unknown failure";
  const int type = e ? static_cast<int>(walnutpie_get_error_type(e)) : wb200_generic;
  if (e) walnutpie_destroy_error(e);
  e = nullptr;
  if (type == wb200_config) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

struct SessionGuard {
  wb200_session* s = nullptr;
  ~SessionGuard() {
    if (s) wb200_session_destroy(s);
  }
};

inline WalnutTuning tuning_from(const WalnutsConfig& c) {
  WalnutTuning t;
  walnuts_b200_default_tuning(&t);
  const WarmupConfig& w = c.warmup();
  const SamplingConfig& s = c.sampling();
  t.min_warmup_iter = static_cast<int>(w.min_iter());
  t.max_warmup_iter = static_cast<int>(w.max_iter());
  t.step_size_converge_tol = w.step_size_converge_tol();
  t.mass_converge_tol = w.mass_converge_tol();
  t.mass_init_count = w.mass_init_count();
  t.mass_additive_smoothing = w.mass_additive_smoothing();
  t.max_macro_steps_target = w.max_macro_steps_target();
  t.step_accept_rate_target = w.step_accept_rate_target();
  t.step_learning_rate = w.step_learning_rate();
  t.step_gradient_decay = w.step_gradient_decay();
  t.step_sq_gradient_decay = w.step_sq_gradient_decay();
  t.step_stabilization = w.step_stabilization();
  t.step_learn_rate_decay = w.step_learn_rate_decay();
  t.publish_stride = static_cast<int>(w.publish_stride());
  t.min_sampling_iter = static_cast<int>(s.min_iter());
  t.max_sampling_iter = static_cast<int>(s.max_iter());
  t.rhat_converge_tol = s.rhat_converge_tol();
  t.max_trajectory_doublings = static_cast<int>(s.max_trajectory_doublings());
  t.max_step_halvings = static_cast<int>(s.max_step_halvings());
  t.max_hamiltonian_error = s.max_hamiltonian_error();
  t.min_micro_steps = static_cast<int>(s.min_micro_steps());
  return t;
}

// ErrorCallback is optional here (the built-in device densities cannot throw)
template <class H, class = void>
struct has_on_logp_exception : std::false_type {};
template <class H>
struct has_on_logp_exception<
    H, std::void_t<decltype(std::declval<H&>().on_logp_exception(
           std::declval<const Vector&>(), std::declval<const std::exception&>()))>>
    : std::true_type {};

inline std::vector<double> flatten(const std::vector<Vector>& rows, std::size_t dims) {
  std::vector<double> out(rows.size() * dims);
  for (std::size_t m = 0; m < rows.size(); ++m) {
    std::copy(rows[m].begin(), rows[m].end(), out.begin() + m * dims);
  }
  return out;
}

}  // namespace detail

template <class H, class GH, class IC>
inline void walnuts(std::size_t seed, std::vector<H>& chain_handlers, GH& global_handler,
                    const IC& interrupt_callback, const WalnutModelDesc& model,
                    const WalnutsConfig& config, const DeviceInit& dev = DeviceInit()) {
  const std::size_t C = config.init().num_chains();
  const std::size_t D = config.init().dims();
  if (chain_handlers.size() != C) {  // api.hpp:41-44
    throw std::invalid_argument(
        "chain_handlers.size() must be equal to config.init().num_chains()");
  }
  if (static_cast<std::size_t>(model.D) != D) {
    throw std::invalid_argument("model dimension must be equal to config.init().dims()");
  }
  const WalnutTuning tuning = detail::tuning_from(config);
  const int stride = std::max(tuning.publish_stride, 1);
  const int max_warm = tuning.max_warmup_iter, max_samp = tuning.max_sampling_iter;

  WalnutpyError* e = nullptr;
  detail::SessionGuard guard;
  detail::check(wb200_session_create(&model, C, static_cast<unsigned int>(seed),
                                     dev.chain_offset, &tuning, dev.device, &guard.s, &e), e);
  wb200_session* s = guard.s;
  {
    const std::vector<double> pos = detail::flatten(config.init().positions(), D);
    const std::vector<double> mass = detail::flatten(config.init().masses(), D);
    detail::check(wb200_session_init(
                      s, dev.random_positions ? nullptr : pos.data(), dev.init_scale,
                      dev.gradient_masses ? nullptr : mass.data(),
                      dev.adapt_steps ? nullptr : config.init().step_sizes().data(), &e), e);
  }
  detail::check(wb200_session_reserve_draws(s, static_cast<long long>(max_warm) + max_samp, 1,
                                            &e), e);

  std::vector<double> draws, lp, step, inv_mass;
  Vector position(D), diag(D);
  // hand rows [first, first + n) of every chain to its handler, in iteration order
  auto publish = [&](long long first, int n, bool warm) {
    if (n <= 0) return;
    draws.resize(C * n * D);
    lp.resize(C * n);
    detail::check(wb200_session_get_draws(s, first, n, draws.data(), &e), e);
    if (warm) {
      step.resize(C * n);
      inv_mass.resize(C * n * D);
    }
    detail::check(wb200_session_get_trace(s, first, n, lp.data(), nullptr,
                                          warm ? step.data() : nullptr,
                                          warm ? inv_mass.data() : nullptr, &e), e);
    for (std::size_t m = 0; m < C; ++m) {
      for (int i = 0; i < n; ++i) {
        const std::size_t row = m * n + i;
        position.assign(draws.begin() + row * D, draws.begin() + (row + 1) * D);
        if (warm) {
          diag.assign(inv_mass.begin() + row * D, inv_mass.begin() + (row + 1) * D);
          chain_handlers[m].on_warmup(position, lp[row], step[row], diag);
        } else {
          chain_handlers[m].on_sample(position, lp[row]);
        }
      }
    }
  };

  // failed evaluations of a batched density since the last call (util.hpp:336-346)
  unsigned long long exceptions_seen = 0;
  auto report_exceptions = [&](long long last_row) {
    unsigned long long total = 0;
    if (wb200_session_logp_exceptions(s, &total) != 0 || total == exceptions_seen) return;
    if constexpr (detail::has_on_logp_exception<H>::value) {
      const std::runtime_error exn("logp failed: the batched density returned non-zero");
      draws.resize(C * D);
      detail::check(wb200_session_get_draws(s, last_row, 1, draws.data(), &e), e);
      for (unsigned long long k = exceptions_seen; k < total; ++k) {
        for (std::size_t m = 0; m < C; ++m) {
          position.assign(draws.begin() + m * D, draws.begin() + (m + 1) * D);
          chain_handlers[m].on_logp_exception(position, exn);
        }
      }
    }
    exceptions_seen = total;
  };

  // ---- detail::adapt (adapt.hpp:173-259): blocks of publish_stride, controller between
  int warm_done = 0;
  while (warm_done < max_warm) {
    const int n = std::min(stride, max_warm - warm_done);
    detail::check(wb200_session_warmup(s, n, 1, &e), e);
    publish(warm_done, n, true);
    warm_done += n;
    report_exceptions(warm_done - 1);
    interrupt_callback.throw_if_interrupted();  // adapt.hpp:227
    if (warm_done >= tuning.min_warmup_iter && warm_done < max_warm) {
      double deviation[2];
      detail::check(wb200_session_warmup_sums(s, nullptr, &e), e);
      detail::check(wb200_session_warmup_deviation(s, nullptr, deviation, &e), e);
      if (deviation[0] <= tuning.mass_converge_tol &&
          deviation[1] <= tuning.step_size_converge_tol) {
        break;
      }
    }
  }
  // ---- AdaptiveWalnuts::sampler() (adaptive_walnuts.hpp:263-271)
  detail::check(wb200_session_freeze(s, &e), e);
  {
    std::vector<double> im(C * D), st(C);
    detail::check(wb200_session_get_state(s, nullptr, im.data(), st.data(), nullptr, nullptr,
                                          &e), e);
    for (std::size_t m = 0; m < C; ++m) {
      diag.assign(im.begin() + m * D, im.begin() + (m + 1) * D);
      chain_handlers[m].on_warmup_complete(st[m], diag);
    }
  }
  // ---- detail::sample (sampler.hpp:118-192): R-hat of lp between blocks
  int samp_done = 0;
  while (samp_done < max_samp) {
    const int n = std::min(stride, max_samp - samp_done);
    detail::check(wb200_session_sample(s, n, 1, &e), e);
    publish(static_cast<long long>(warm_done) + samp_done, n, false);
    samp_done += n;
    report_exceptions(static_cast<long long>(warm_done) + samp_done - 1);
    interrupt_callback.throw_if_interrupted();  // sampler.hpp:154
    if (samp_done >= tuning.min_sampling_iter && samp_done < max_samp) {
      // util.hpp:401-404 is a two-pass variance: the mean of the chain means first
      double m0[4], m4[4];
      detail::check(wb200_session_lp_moments(s, m0, &e), e);
      detail::check(wb200_session_lp_moments_centered(s, m0[0] / m0[3], m4, &e), e);
      const double M = m4[3];
      const double var_of_means = (m4[1] - m4[0] * m4[0] / M) / (M - 1.0);
      const double mean_of_vars = m4[2] / M;
      const double r_hat = std::sqrt(1 + var_of_means / mean_of_vars);  // sampler.hpp:132-151
      global_handler.on_r_hat(r_hat);
      if (r_hat <= tuning.rhat_converge_tol) break;
    }
  }
}

}  // namespace walnuts_b200
