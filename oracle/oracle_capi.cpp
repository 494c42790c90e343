// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
// C entry points over oracle/walnuts_oracle.hpp (see oracle_capi.h).
#include "oracle_capi.h"

#include <atomic>
#include <barrier>
#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <thread>

#include "summary_oracle.hpp"
#include "walnuts_oracle.hpp"

namespace {
using namespace oracle;

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

template <class Fn>
void with_target(const OracleTarget* t, Fn&& fn) {
  switch (t->kind) {
    case kStdNormal: fn(StdNormal{}); break;
    case kDiagGaussian: {
      DiagGaussian g;
      const double* p = static_cast<const double*>(t->data0);
      g.prec.assign(p, p + t->D);
      fn(g);
      break;
    }
    case kFunnel: fn(Funnel{}); break;
    case kLogistic: {
      Logistic l;
      l.N = t->N; l.D = t->D;
      l.X = static_cast<const double*>(t->data0);
      l.y = static_cast<const double*>(t->data1);
      fn(l);
      break;
    }
    case 4: {
      CFuncTarget c;
      c.fn = reinterpret_cast<LOGP_CFUNC>(const_cast<void*>(t->data0));
      c.data = const_cast<void*>(t->data1);
      fn(c);
      break;
    }
    default: throw std::invalid_argument("unknown target kind");
  }
}

void split_config(const OracleConfig& c, WarmupConfig& w, SamplingConfig& s) {
  if (c.min_warmup_iter > c.max_warmup_iter) {
    throw std::invalid_argument("min_iter cannot be greater than than max_iter");
  }
  if (c.min_sampling_iter > c.max_sampling_iter) {
    throw std::invalid_argument("min_iter must be <= max_iter");
  }
  w.min_iter = c.min_warmup_iter; w.max_iter = c.max_warmup_iter;
  w.step_size_converge_tol = c.step_size_converge_tol;
  w.mass_converge_tol = c.mass_converge_tol;
  w.mass_init_count = c.mass_init_count;
  w.mass_additive_smoothing = c.mass_additive_smoothing;
  w.max_macro_steps_target = c.max_macro_steps_target;
  w.step_accept_rate_target = c.step_accept_rate_target;
  w.step_learning_rate = c.step_learning_rate;
  w.step_gradient_decay = c.step_gradient_decay;
  w.step_sq_gradient_decay = c.step_sq_gradient_decay;
  w.step_stabilization = c.step_stabilization;
  w.step_learn_rate_decay = c.step_learn_rate_decay;
  w.publish_stride = c.publish_stride > 0 ? c.publish_stride : 5;
  s.min_iter = c.min_sampling_iter; s.max_iter = c.max_sampling_iter;
  s.max_trajectory_doublings = c.max_trajectory_doublings;
  s.max_step_halvings = c.max_step_halvings;
  s.max_hamiltonian_error = c.max_hamiltonian_error;
  s.min_micro_steps = c.min_micro_steps;
  s.rhat_converge_tol = c.rhat_converge_tol;
}

struct ChainOut {
  double* warmup_draws = nullptr; double* warmup_lp = nullptr;
  double* warmup_step = nullptr; double* warmup_inv_mass = nullptr;
  int* warmup_depth = nullptr;
  double* draws = nullptr; double* lp = nullptr; int* depth = nullptr;
  double* inv_mass_out = nullptr; double* step_out = nullptr;
  int* min_micro_out = nullptr;
};

// warm-up then sampling for one chain with a given Random policy
template <class F, class Rand, class MakeRand>
void run_chain_impl(const F& f, Rand rand, MakeRand make_rand,
                    const WarmupConfig& w, const SamplingConfig& s,
                    const InitChainConfig& init, int n_warmup, int n_sampling,
                    const ChainOut& o, uint64_t* grad_evals) {
  const std::size_t D = init.position.size();
  ChainHandler h;
  int wi = 0, si = 0;
  h.on_warmup = [&](const Vec& th, double lp, double step, const Vec& im) {
    if (o.warmup_draws) std::memcpy(o.warmup_draws + wi * D, th.data(), D * 8);
    if (o.warmup_lp) o.warmup_lp[wi] = lp;
    if (o.warmup_step) o.warmup_step[wi] = step;
    if (o.warmup_inv_mass) std::memcpy(o.warmup_inv_mass + wi * D, im.data(), D * 8);
    ++wi;
  };
  h.on_warmup_complete = [&](double step, const Vec& im) {
    if (o.step_out) *o.step_out = step;
    if (o.inv_mass_out) std::memcpy(o.inv_mass_out, im.data(), D * 8);
  };
  h.on_sample = [&](const Vec& th, double lp) {
    if (o.draws) std::memcpy(o.draws + si * D, th.data(), D * 8);
    if (o.lp) o.lp[si] = lp;
    ++si;
  };
  uint64_t counter = 0;
  AdaptiveWalnuts<F, Rand> adapter(std::move(rand), h, f, init, w, s, &counter);
  for (int n = 0; n < n_warmup; ++n) {
    adapter();
    if (o.warmup_depth) o.warmup_depth[n] = static_cast<int>(adapter.last_depth());
  }
  auto sampler = adapter.sampler(make_rand);
  sampler.set_iteration(static_cast<uint32_t>(adapter.iter()));
  if (o.min_micro_out) *o.min_micro_out = static_cast<int>(adapter.min_micro_steps());
  for (int n = 0; n < n_sampling; ++n) {
    sampler();
    if (o.depth) o.depth[n] = static_cast<int>(sampler.last_depth());
  }
  if (grad_evals) *grad_evals = counter;
}

template <class F>
void run_chain_policy(const F& f, uint32_t seed, uint32_t chain, int policy,
                      const WarmupConfig& w, const SamplingConfig& s,
                      const InitChainConfig& init, int n_warmup,
                      int n_sampling, const ChainOut& o, uint64_t* grad_evals) {
  if (policy == 0) {
    std::seed_seq ss{static_cast<std::size_t>(seed),
                     static_cast<std::size_t>(chain) + 1u};
    std::mt19937_64 rng(ss);
    using R = StdRand<std::mt19937_64>;
    run_chain_impl(f, R(rng), [](R& r) { return R(r.rng()); }, w, s, init,
                   n_warmup, n_sampling, o, grad_evals);
  } else {
    using R = PhiloxRand;
    run_chain_impl(f, R(seed, chain), [](R& r) { return r; }, w, s, init,
                   n_warmup, n_sampling, o, grad_evals);
  }
}

std::vector<std::size_t> to_lengths(const int* lengths, int num_chains) {
  std::vector<std::size_t> l(num_chains);
  for (int i = 0; i < num_chains; ++i) l[i] = static_cast<std::size_t>(lengths[i]);
  return l;
}
}  // namespace

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

int oracle_run_chain(const OracleTarget* target, const OracleConfig* cfg,
                     uint32_t seed, uint32_t chain, int rng_policy,
                     const double* theta0, const double* mass0, double step0,
                     int n_warmup, int n_sampling, double* warmup_draws,
                     double* warmup_lp, double* warmup_step,
                     double* warmup_inv_mass, int* warmup_depth, double* draws,
                     double* lp, int* depth, double* inv_mass_out,
                     double* step_out, int* min_micro_out,
                     uint64_t* grad_evals) {
  return guarded([&] {
    WarmupConfig w; SamplingConfig s;
    split_config(*cfg, w, s);
    const std::size_t D = target->D;
    InitChainConfig init{step0, Vec(theta0, theta0 + D), Vec(mass0, mass0 + D)};
    ChainOut o{warmup_draws, warmup_lp, warmup_step, warmup_inv_mass,
               warmup_depth, draws, lp, depth, inv_mass_out, step_out,
               min_micro_out};
    with_target(target, [&](const auto& f) {
      run_chain_policy(f, seed, chain, rng_policy, w, s, init, n_warmup,
                       n_sampling, o, grad_evals);
    });
  });
}

int oracle_run_sampler(const OracleTarget* target, uint32_t seed,
                       uint32_t chain, int rng_policy, uint32_t first_iter,
                       const double* theta0, const double* inv_mass,
                       double step, int max_depth, int max_halvings,
                       int min_micro, double max_error, int n_iter,
                       double* draws, double* lp, int* depth,
                       uint64_t* grad_evals) {
  return guarded([&] {
    const std::size_t D = target->D;
    Vec th(theta0, theta0 + D), im(inv_mass, inv_mass + D);
    ChainHandler h;
    uint64_t counter = 0;
    auto loop = [&](auto& sampler) {
      sampler.set_iteration(first_iter);
      for (int n = 0; n < n_iter; ++n) {
        double l = sampler();
        if (draws) std::memcpy(draws + n * D, sampler.theta().data(), D * 8);
        if (lp) lp[n] = l;
        if (depth) depth[n] = static_cast<int>(sampler.last_depth());
      }
    };
    with_target(target, [&](const auto& f) {
      using F = std::decay_t<decltype(f)>;
      if (rng_policy == 0) {
        std::seed_seq ss{static_cast<std::size_t>(seed),
                         static_cast<std::size_t>(chain) + 1u};
        std::mt19937_64 rng(ss);
        using R = StdRand<std::mt19937_64>;
        WalnutsSampler<F, R> sm(R(rng), h, f, th, im, step, max_depth,
                                max_halvings, min_micro, max_error, &counter);
        loop(sm);
      } else {
        WalnutsSampler<F, PhiloxRand> sm(PhiloxRand(seed, chain), h, f, th, im,
                                         step, max_depth, max_halvings,
                                         min_micro, max_error, &counter);
        loop(sm);
      }
    });
    if (grad_evals) *grad_evals = counter;
  });
}

int oracle_init_positions(size_t num_chains, size_t D, uint32_t seed,
                          double radius, double* positions) {
  return guarded([&] {
    if (!(std::isfinite(radius) && radius > 0)) {
      throw std::invalid_argument("init_scale must be finite and > 0");
    }
    std::seed_seq ss{seed, 1u};
    std::mt19937_64 rng(ss);
    StdRand<std::mt19937_64> rand(rng);  // one Random for all chains
    for (size_t c = 0; c < num_chains; ++c) {
      Vec z = rand.standard_normal(D);
      for (size_t i = 0; i < D; ++i) positions[c * D + i] = z[i] * radius;
    }
  });
}

int oracle_init_mass_step(const OracleTarget* target, size_t num_chains,
                          uint32_t seed, const double* positions,
                          const double* mass_in, double smoothing,
                          double step_init, double* mass_out,
                          double* step_out) {
  return guarded([&] {
    const size_t D = target->D;
    with_target(target, [&](const auto& f) {
      for (size_t c = 0; c < num_chains; ++c) {
        Vec pos(positions + c * D, positions + (c + 1) * D);
        Vec mass = mass_in ? Vec(mass_in + c * D, mass_in + (c + 1) * D)
                           : init_mass_from_gradient(f, pos, smoothing);
        std::memcpy(mass_out + c * D, mass.data(), D * 8);
      }
      std::seed_seq ss{seed, 2u};
      std::mt19937_64 rng(ss);
      for (size_t c = 0; c < num_chains; ++c) {
        Vec pos(positions + c * D, positions + (c + 1) * D);
        Vec mass(mass_out + c * D, mass_out + (c + 1) * D);
        StdRand<std::mt19937_64> rand(rng);  // fresh per chain, util.hpp:288
        step_out[c] = adapt_step(rand, f, pos, mass, step_init, D);
      }
    });
  });
}

namespace {
// Random<RNG>::standard_normal (util.hpp:124-127) over one Philox kind
struct PhiloxKindRand {
  oracle::PhiloxRand base;
  uint32_t kind;
  std::vector<double> standard_normal(std::size_t n) { return base.normals(n, kind); }
};
}  // namespace

int oracle_init_positions_philox(size_t num_chains, size_t D, uint32_t seed,
                                 uint32_t chain_offset, double radius,
                                 double* positions) {
  return guarded([&] {
    if (!(std::isfinite(radius) && radius > 0)) {
      throw std::invalid_argument("init_scale must be finite and > 0");
    }
    for (size_t c = 0; c < num_chains; ++c) {
      PhiloxRand rand(seed, chain_offset + static_cast<uint32_t>(c));
      Vec z = rand.normals(D, kKindInit);  // config.hpp:259-268 per chain
      for (size_t i = 0; i < D; ++i) positions[c * D + i] = z[i] * radius;
    }
  });
}

int oracle_init_mass_step_philox(const OracleTarget* target, size_t num_chains,
                                 uint32_t seed, uint32_t chain_offset,
                                 const double* positions, const double* mass_in,
                                 double smoothing, double step_init,
                                 double* mass_out, double* step_out) {
  return guarded([&] {
    const size_t D = target->D;
    with_target(target, [&](const auto& f) {
      for (size_t c = 0; c < num_chains; ++c) {
        Vec pos(positions + c * D, positions + (c + 1) * D);
        Vec mass = mass_in ? Vec(mass_in + c * D, mass_in + (c + 1) * D)
                           : init_mass_from_gradient(f, pos, smoothing);
        std::memcpy(mass_out + c * D, mass.data(), D * 8);
        PhiloxKindRand rand{PhiloxRand(seed, chain_offset + static_cast<uint32_t>(c)),
                            kKindStepInit};
        step_out[c] = adapt_step(rand, f, pos, mass, step_init, D);
      }
    });
  });
}

int oracle_warmup_controller(size_t num_chains, size_t D, const double* log_step,
                             const double* log_mass, double* max_rel_mass,
                             double* max_rel_step) {
  return guarded([&] {
    std::vector<AdaptSnapshot> latest(num_chains);
    for (size_t m = 0; m < num_chains; ++m) {
      latest[m].iter = 1;
      latest[m].log_step = log_step[m];
      latest[m].log_mass.assign(log_mass + m * D, log_mass + (m + 1) * D);
      latest[m].mass.resize(D);
      for (size_t d = 0; d < D; ++d) latest[m].mass[d] = std::exp(latest[m].log_mass[d]);
    }
    WarmupConfig w;
    w.min_iter = 0;
    warmup_should_stop(latest, w, max_rel_mass, max_rel_step);
  });
}

int oracle_sampling_rhat(size_t num_chains, const double* mean, const double* var,
                         double* r_hat) {
  return guarded([&] {
    std::vector<ChainStats> stats(num_chains);
    for (size_t m = 0; m < num_chains; ++m) stats[m] = {mean[m], var[m], 1};
    SamplingConfig s;
    s.min_iter = 0;
    bool evaluated = false;
    sampling_should_stop(stats, s, r_hat, &evaluated);
  });
}

namespace {
inline float maddf(float a, float b, float c) {
  return oracle::fused_arith() ? std::fmaf(a, b, c) : a * b + c;
}
// log density and gradient of kinds 0-2 in float (funnel scalars in double, as the device)
void logp_grad_f32(const OracleTarget* t, const std::vector<float>& x, double& lp,
                   std::vector<float>& g) {
  const size_t D = t->D;
  if (t->kind == 0 || t->kind == 1) {
    const double* prec = static_cast<const double*>(t->data0);
    float s = 0.0f;
    for (size_t i = 0; i < D; ++i) {
      const float tt = t->kind == 0 ? x[i] : x[i] * static_cast<float>(prec[i]);
      s = maddf(x[i], tt, s);
      g[i] = -tt;
    }
    lp = static_cast<double>(-0.5f * s);
  } else if (t->kind == 2) {
    float ss = 0.0f;
    for (size_t i = 1; i < D; ++i) ss = maddf(x[i], x[i], ss);
    const double v = x[0], ev = std::exp(-v), hd = 0.5 * static_cast<double>(D - 1);
    const double q = 0.5 * ev * static_cast<double>(ss);
    const float evf = static_cast<float>(ev);
    for (size_t i = 1; i < D; ++i) g[i] = -(x[i] * evf);
    g[0] = static_cast<float>(-(v * oracle::Funnel::kInv9) - hd + q);
    lp = static_cast<double>(static_cast<float>(-((v * v) * oracle::Funnel::kInv18) - hd * v - q));
  } else {
    throw std::invalid_argument("fp32 mode covers the element-wise targets (kinds 0-2)");
  }
}
}  // namespace

int oracle_orbit_f32(const OracleTarget* target, const double* theta,
                     const double* rho, const double* inv_mass, double step,
                     int num_steps, double* theta_out, double* rho_out,
                     double* grad_out, double* logp_out, double* joint_out) {
  return guarded([&] {
    const size_t D = target->D;
    std::vector<float> th(D), rh(D), im(D), g(D);
    for (size_t i = 0; i < D; ++i) {
      th[i] = static_cast<float>(theta[i]);
      rh[i] = static_cast<float>(rho[i]);
      im[i] = static_cast<float>(inv_mass[i]);
    }
    double lp;
    logp_grad_f32(target, th, lp, g);
    const float h = static_cast<float>(step), hh = static_cast<float>(0.5 * step);
    for (int n = 0; n < num_steps; ++n) {   // walnuts.hpp:329-332 in float
      for (size_t i = 0; i < D; ++i) rh[i] = maddf(hh, g[i], rh[i]);
      for (size_t i = 0; i < D; ++i) th[i] = maddf(h * im[i], rh[i], th[i]);
      logp_grad_f32(target, th, lp, g);
      for (size_t i = 0; i < D; ++i) rh[i] = maddf(hh, g[i], rh[i]);
    }
    float kin = 0.0f;
    for (size_t i = 0; i < D; ++i) kin = maddf(im[i], rh[i] * rh[i], kin);
    for (size_t i = 0; i < D; ++i) {
      theta_out[i] = th[i]; rho_out[i] = rh[i]; grad_out[i] = g[i];
    }
    *logp_out = lp;
    *joint_out = lp + (-0.5 * static_cast<double>(kin));
  });
}

int oracle_set_fused_arith(int fused) {
  const int before = oracle::fused_arith() ? 1 : 0;
  oracle::fused_arith() = fused != 0;
  return before;
}

double oracle_leapfrog_error(const OracleTarget* target, const double* theta,
                             const double* rho, const double* inv_mass,
                             double step) {
  double out = std::numeric_limits<double>::quiet_NaN();
  const size_t D = target->D;
  with_target(target, [&](const auto& f) {
    out = leapfrog_error(f, Vec(theta, theta + D), Vec(rho, rho + D),
                         Vec(inv_mass, inv_mass + D), step);
  });
  return out;
}

double oracle_log_sum_exp(double a, double b) { return log_sum_exp(a, b); }

double oracle_logp_momentum(const double* rho, const double* inv_mass,
                            size_t D) {
  return logp_momentum(Vec(rho, rho + D), Vec(inv_mass, inv_mass + D));
}

int oracle_walnuts(const OracleTarget* target, const OracleConfig* cfg,
                   size_t num_chains, uint32_t seed, const double* positions,
                   const double* mass, const double* steps, int save_warmup,
                   double* out, int* final_lengths, double* stepsize_out,
                   double* inv_metric_out, uint64_t* grad_evals,
                   double* seconds_warmup, double* seconds_sampling) {
  // Deterministic restatement of api.hpp:33-69 + adapt.hpp + sampler.hpp:
  // one thread per chain; the controllers' convergence tests run at barriers
  // (every publish_stride warm-up iterations / every sampling iteration)
  // instead of on a free-running poll, which removes the reference's
  // thread-schedule dependence (docs/py.rst:13-20) and nothing else.
  return guarded([&] {
    WarmupConfig w; SamplingConfig s;
    split_config(*cfg, w, s);
    const size_t D = target->D, C = num_chains;
    const size_t stride_out = D * (s.max_iter + (save_warmup ? w.max_iter : 0));
    with_target(target, [&](const auto& f) {
      using F = std::decay_t<decltype(f)>;
      using R = StdRand<std::mt19937_64>;
      std::vector<std::mt19937_64> rngs;
      for (size_t m = 0; m < C; ++m) {
        std::seed_seq ss{static_cast<std::size_t>(seed), m + 1u};
        rngs.emplace_back(ss);
      }
      std::vector<ChainHandler> handlers(C);
      std::vector<size_t> written(C, 0), written_warmup(C, 0);
      std::vector<uint64_t> counters(C, 0);
      for (size_t m = 0; m < C; ++m) {
        double* base = out ? out + stride_out * m : nullptr;
        handlers[m].on_warmup = [&, m, base](const Vec& th, double, double,
                                             const Vec&) {
          if (save_warmup && base) {
            std::memcpy(base + written[m] * D, th.data(), D * 8);
            ++written[m]; ++written_warmup[m];
          }
        };
        handlers[m].on_sample = [&, m, base](const Vec& th, double) {
          if (base) std::memcpy(base + written[m] * D, th.data(), D * 8);
          ++written[m];
        };
        handlers[m].on_warmup_complete = [&, m](double step, const Vec& im) {
          if (stepsize_out) stepsize_out[m] = step;
          if (inv_metric_out) std::memcpy(inv_metric_out + m * D, im.data(), D * 8);
        };
      }
      std::vector<std::unique_ptr<AdaptiveWalnuts<F, R>>> adapters;
      for (size_t m = 0; m < C; ++m) {
        InitChainConfig init{steps[m], Vec(positions + m * D, positions + (m + 1) * D),
                             Vec(mass + m * D, mass + (m + 1) * D)};
        adapters.push_back(std::make_unique<AdaptiveWalnuts<F, R>>(
            R(rngs[m]), handlers[m], f, init, w, s, &counters[m]));
      }
      auto t0 = std::chrono::steady_clock::now();
      // ---- warm-up (adapt.hpp:110-129 + :173-229)
      {
        std::atomic<bool> stop{false};
        std::vector<AdaptSnapshot> latest(C);
        auto on_round = [&]() noexcept {
          if (warmup_should_stop(latest, w)) stop.store(true);
        };
        std::barrier bar(static_cast<std::ptrdiff_t>(C), on_round);
        std::vector<std::thread> threads;
        for (size_t m = 0; m < C; ++m) {
          threads.emplace_back([&, m] {
            auto& a = *adapters[m];
            latest[m] = make_snapshot(a, 0);
            size_t iter = 1;
            for (; iter <= w.max_iter; ++iter) {
              a();
              if (iter % w.publish_stride == 0 || iter == w.max_iter) {
                latest[m] = make_snapshot(a, iter);
                bar.arrive_and_wait();
                if (stop.load()) break;
              }
            }
          });
        }
        for (auto& t : threads) t.join();
      }
      auto t1 = std::chrono::steady_clock::now();
      // ---- sampling (sampler.hpp:79-94 + :118-158)
      std::vector<std::unique_ptr<WalnutsSampler<F, R>>> samplers;
      for (size_t m = 0; m < C; ++m) {
        samplers.push_back(std::make_unique<WalnutsSampler<F, R>>(
            adapters[m]->sampler([](R& r) { return R(r.rng()); })));
      }
      {
        std::atomic<bool> stop{false};
        std::vector<ChainStats> stats(C);
        const bool fixed = (s.min_iter == s.max_iter);
        auto on_round = [&]() noexcept {
          double rh; bool ev;
          if (sampling_should_stop(stats, s, &rh, &ev)) stop.store(true);
        };
        std::barrier bar(static_cast<std::ptrdiff_t>(C), on_round);
        std::vector<std::thread> threads;
        for (size_t m = 0; m < C; ++m) {
          threads.emplace_back([&, m] {
            WelfordAccumulator acc;
            for (size_t iter = 1; iter <= s.max_iter; ++iter) {
              double lpv = (*samplers[m])();
              acc.observe(lpv);
              if (fixed) continue;
              stats[m] = {acc.mean(), acc.sample_variance(), acc.count()};
              bar.arrive_and_wait();
              if (stop.load()) break;
            }
          });
        }
        for (auto& t : threads) t.join();
      }
      auto t2 = std::chrono::steady_clock::now();
      if (seconds_warmup) *seconds_warmup = std::chrono::duration<double>(t1 - t0).count();
      if (seconds_sampling) *seconds_sampling = std::chrono::duration<double>(t2 - t1).count();
      uint64_t total = 0;
      for (size_t m = 0; m < C; ++m) {
        total += counters[m];
        if (final_lengths) {
          final_lengths[m] = static_cast<int>(written_warmup[m]);
          final_lengths[m + C] = static_cast<int>(written[m] - written_warmup[m]);
        }
      }
      if (grad_evals) *grad_evals = total;
    });
  });
}

int oracle_orbit(const OracleTarget* target, const double* theta,
                 const double* rho, const double* inv_mass, double step,
                 int num_steps, double* theta_out, double* rho_out,
                 double* grad_out, double* logp_out, double* joint_out) {
  return guarded([&] {
    const size_t D = target->D;
    with_target(target, [&](const auto& f) {
      Vec th(theta, theta + D), rh(rho, rho + D), im(inv_mass, inv_mass + D), g;
      double lp;
      f(th, lp, g);
      double half = 0.5 * step;
      for (int n = 0; n < num_steps; ++n) leapfrog(f, im, step, half, th, rh, g, lp);
      std::memcpy(theta_out, th.data(), D * 8);
      std::memcpy(rho_out, rh.data(), D * 8);
      std::memcpy(grad_out, g.data(), D * 8);
      *logp_out = lp;
      *joint_out = lp + logp_momentum(rh, im);
    });
  });
}

int oracle_logp_grad(const OracleTarget* target, const double* theta,
                     double* logp, double* grad) {
  return guarded([&] {
    const size_t D = target->D;
    with_target(target, [&](const auto& f) {
      Vec g;
      f(Vec(theta, theta + D), *logp, g);
      std::memcpy(grad, g.data(), D * 8);
    });
  });
}

void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                   uint32_t k0, uint32_t k1, uint32_t* out4) {
  uint32_t ctr[4] = {c0, c1, c2, c3}, key[2] = {k0, k1};
  Philox4x32::block(ctr, key, out4);
}

void oracle_philox_normals(uint32_t seed, uint32_t chain, uint32_t iter,
                           uint32_t kind, size_t n, double* out) {
  PhiloxRand r(seed, chain);
  r.begin_transition(iter);
  Vec z = r.normals(n, kind);
  std::memcpy(out, z.data(), n * 8);
}

double oracle_philox_uniform(uint32_t seed, uint32_t chain, uint32_t iter,
                             uint32_t index) {
  uint32_t w[4];
  philox_draw(seed, chain, iter, kKindScalar, index, w);
  return u01_from_words(w[0], w[1]);
}

int oracle_ess(const double* draws, int num_draws, int num_params,
               const int* lengths, int num_chains, double* out) {
  return guarded([&] {
    Chains c(draws, num_draws, num_params, to_lengths(lengths, num_chains));
    auto r = effective_sample_size(c);
    std::memcpy(out, r.data(), r.size() * 8);
  });
}
int oracle_r_hat(const double* draws, int num_draws, int num_params,
                 const int* lengths, int num_chains, double* out) {
  return guarded([&] {
    Chains c(draws, num_draws, num_params, to_lengths(lengths, num_chains));
    auto r = r_hat(c);
    std::memcpy(out, r.data(), r.size() * 8);
  });
}
int oracle_mcse(const double* draws, int num_draws, int num_params,
                const int* lengths, int num_chains, double* out) {
  return guarded([&] {
    Chains c(draws, num_draws, num_params, to_lengths(lengths, num_chains));
    auto r = mcse(c);
    std::memcpy(out, r.data(), r.size() * 8);
  });
}
int oracle_autocovariance(const double* draws, int num_draws, int num_params,
                          const int* lengths, int num_chains, double* out) {
  return guarded([&] {
    Chains c(draws, num_draws, num_params, to_lengths(lengths, num_chains));
    for (std::size_t m = 0; m < c.num_chains(); ++m) {
      for (std::size_t d = 0; d < c.D; ++d) {
        auto ac = autocovariance(c, m, d);
        for (std::size_t t = 0; t < ac.size(); ++t) {
          out[(c.start[m] + t) * c.D + d] = ac[t];
        }
      }
    }
  });
}

}  // extern "C"
