// extern "C" surface of libwalnuts_b200.so (see include/walnuts_b200.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../host/config.hpp"
#include "engine.cuh"
#include "user_density.cuh"
#include "user_density.cuh"
#include "stream.cuh"

namespace wb200 {
// walnutpy.cpp:36-62: the builders validate exactly like the reference
static void validate_tuning(const WalnutTuning& t) {
  using namespace walnuts_b200;
  if (t.min_warmup_iter < 0 || t.max_warmup_iter < 0 || t.min_sampling_iter < 0 ||
      t.max_sampling_iter < 0) {
    throw std::invalid_argument("iteration counts must be non-negative");
  }
  WarmupConfigBuilder()
      .min_max_iter(t.min_warmup_iter, t.max_warmup_iter)
      .step_size_converge_tol(t.step_size_converge_tol)
      .mass_converge_tol(t.mass_converge_tol)
      .mass_init_count(t.mass_init_count)
      .mass_additive_smoothing(t.mass_additive_smoothing)
      .max_macro_steps_target(t.max_macro_steps_target)
      .step_accept_rate_target(t.step_accept_rate_target)
      .step_learning_rate(t.step_learning_rate)
      .step_gradient_decay(t.step_gradient_decay)
      .step_sq_gradient_decay(t.step_sq_gradient_decay)
      .step_stabilization(t.step_stabilization)
      .step_learn_rate_decay(t.step_learn_rate_decay)
      .build();
  SamplingConfigBuilder()
      .min_max_iter(t.min_sampling_iter, t.max_sampling_iter)
      .rhat_converge_tol(t.rhat_converge_tol)
      .max_trajectory_doublings(t.max_trajectory_doublings)
      .max_step_halvings(t.max_step_halvings)
      .max_hamiltonian_error(t.max_hamiltonian_error)
      .min_micro_steps(static_cast<std::size_t>(std::max(t.min_micro_steps, 0)))
      .build();
  // WalnutsSampler's own checks (walnuts.hpp:654-659)
  validate::positive(static_cast<std::size_t>(std::max(t.max_trajectory_doublings, 0)),
                     "max_nuts_depth");
  validate::positive(static_cast<std::size_t>(std::max(t.max_step_halvings, 0)),
                     "max_step_halvings");
  if (t.max_trajectory_doublings > kMaxDepth) {
    throw std::invalid_argument("max_trajectory_doublings above " +
                                std::to_string(kMaxDepth) +
                                " is not supported on the device");
  }
  validate::finite_positive(t.step_size_init, "step size");
}

}  // namespace wb200

using namespace wb200;

extern "C" {

const char* walnuts_b200_version(void) { return "walnuts_b200 0.1.0 (sm_100a)"; }

const char* walnutpie_get_error_message(const WalnutpyError* err) {
  if (err == nullptr) return "Something went wrong: No error found";
  return err->msg.c_str();
}
int walnutpie_get_error_type(const WalnutpyError* err) {
  if (err == nullptr) return wb200_generic;
  return err->type;
}
void walnutpie_destroy_error(WalnutpyError* err) { delete err; }
char walnutpie_separator_char(void) { return '\x1C'; }

void walnuts_b200_default_tuning(WalnutTuning* t) {
  t->min_warmup_iter = 50; t->max_warmup_iter = 1000;
  t->min_sampling_iter = 50; t->max_sampling_iter = 1000;
  t->max_trajectory_doublings = 5; t->max_step_halvings = 5; t->min_micro_steps = 1;
  t->max_hamiltonian_error = 0.5;
  t->step_size_converge_tol = 0.1; t->mass_converge_tol = 1.0;
  t->rhat_converge_tol = 1.01;
  t->mass_init_count = 4.0; t->mass_additive_smoothing = 1e-5;
  t->max_macro_steps_target = 15.0;
  t->step_size_init = 1.0;
  t->step_accept_rate_target = 0.8; t->step_learning_rate = 0.05;
  t->step_gradient_decay = 0.8; t->step_sq_gradient_decay = 0.9;
  t->step_stabilization = 1e-4; t->step_learn_rate_decay = 0.5;
  t->publish_stride = 5;
  t->precision = 0;
}

// ------------------------------------------------------------- session ----
int wb200_session_create(const WalnutModelDesc* model, size_t num_chains,
                         unsigned int seed, unsigned int chain_offset,
                         const WalnutTuning* tuning, int device,
                         wb200_session** out, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    if (!model || !tuning || !out) throw std::invalid_argument("null argument");
    if (num_chains < 1) throw std::invalid_argument("num_chains must be at least 1");
    if (model->D < 1) throw std::invalid_argument("num_params must be at least 1");
    validate_tuning(*tuning);
    require_gpu();
    WB200_CUDA(cudaSetDevice(device));
    auto s = std::make_unique<wb200_session>();
    s->device = device;
    s->kind = model->kind;
    s->D = model->D;
    s->N = model->N;
    s->C = static_cast<int>(num_chains);
    s->seed = seed;
    s->chain_offset = chain_offset;
    s->tuning = *tuning;
    if (s->tuning.publish_stride <= 0) s->tuning.publish_stride = 5;
    if (tuning->precision != 0 && tuning->precision != 1) {
      throw std::invalid_argument("precision must be 0 (fp64) or 1 (fp32)");
    }
    if (model->precision != 0 && model->precision != 1) {
      throw std::invalid_argument("precision must be 0 (fp64) or 1 (fp32)");
    }
    s->precision = (tuning->precision || model->precision) ? 1 : 0;
    s->shape = shape_for_dim(s->D);
    // rows are padded to the 2*T*K element slots of a chain's group: vector loads and
    // stores need no bounds checks (padding: theta = rho = grad = 0, unit metric)
    s->ld = 2 * s->shape.T * s->shape.K;
    WB200_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    WB200_CUDA(cudaEventCreate(&s->ev0));
    WB200_CUDA(cudaEventCreate(&s->ev1));
    WB200_CUDA(cudaEventCreate(&s->tm0));
    WB200_CUDA(cudaEventCreate(&s->tm1));
    const size_t CL = static_cast<size_t>(s->C) * s->ld;
    s->theta.alloc(CL);
    s->inv_mass.alloc(CL);
    s->est.alloc(4 * CL);
    s->sc.alloc(s->C);
    s->ticket.alloc(1);
    s->red.alloc(std::max<size_t>(s->C, s->D + 8));
    WB200_CUDA(cudaMemsetAsync(s->theta.ptr, 0, CL * 8, s->stream));
    WB200_CUDA(cudaMemsetAsync(s->inv_mass.ptr, 0, CL * 8, s->stream));
    WB200_CUDA(cudaMemsetAsync(s->sc.ptr, 0, s->C * sizeof(ChainScalars), s->stream));
    // target parameters
    if (s->kind == kDiagGaussian) {
      if (!model->data0) throw std::invalid_argument("diag_gaussian needs data0 = precision[D]");
      const double* prec = static_cast<const double*>(model->data0);
      for (int d = 0; d < s->D; ++d) {
        if (!(std::isfinite(prec[d]) && prec[d] > 0)) {
          throw std::invalid_argument("precision must be finite and > 0");
        }
      }
      s->tparam.alloc(s->ld);
      WB200_CUDA(cudaMemsetAsync(s->tparam.ptr, 0, s->ld * 8, s->stream));
      WB200_CUDA(cudaMemcpyAsync(s->tparam.ptr, prec, s->D * 8,
                                 cudaMemcpyHostToDevice, s->stream));
    } else if (s->kind == kDeviceSource) {
      // the caller's density as CUDA source, compiled into the chain-resident kernel
      if (!model->data0) throw std::invalid_argument("kind 5 needs data0 = CUDA source text");
      if (s->precision != 0) {
        throw std::invalid_argument("a run-time compiled density runs in fp64");
      }
      if (model->N > 0 && !model->data1) {
        throw std::invalid_argument("kind 5 with N > 0 needs data1 = N parameter doubles");
      }
      s->tparam.alloc(std::max<size_t>(model->N, 1));
      WB200_CUDA(cudaMemsetAsync(s->tparam.ptr, 0, std::max<size_t>(model->N, 1) * 8, s->stream));
      if (model->N > 0) {
        WB200_CUDA(cudaMemcpyAsync(s->tparam.ptr, model->data1, model->N * 8,
                                   cudaMemcpyHostToDevice, s->stream));
      }
      s->user = user_module(static_cast<const char*>(model->data0), s->shape, device);
    } else if (s->kind != kStdNormal && s->kind != kFunnel && s->kind != kLogistic &&
               s->kind != kBatchCallback) {
      throw std::invalid_argument("unsupported model kind for the device sampler");
    }
    if (s->kind == kFunnel && s->D < 2) {
      throw std::invalid_argument("funnel needs num_params >= 2");
    }
    // engine: chain-resident kernel for element-wise targets, lock-step ticks where the
    // gradient is a cross-chain batched contraction (or when forced, for testing)
    const char* eng = std::getenv("WB200_ENGINE");
    const bool use_tick = s->kind == kLogistic || s->kind == kBatchCallback ||
                          (eng && std::string(eng) == "tick");
    if (use_tick) {
      // WB200_TICK_SHAPE=64x4: two warps per chain with eight elements per thread for
      // 256 < D <= 512 on the lock-step engine (experiments; default four warps x four)
      const char* tshape = std::getenv("WB200_TICK_SHAPE");
      if (tshape && std::string(tshape) == "64x4" && s->D > 256 && s->D <= 512) {
        s->shape = LaunchShape{64, 4, 64, 1};
      }
      if (s->kind == kDeviceSource) {
        throw std::invalid_argument("a run-time compiled density runs on the chain-resident "
                                    "engine (WB200_ENGINE=tick is set)");
      }
      if (s->precision == 1) {
        throw std::invalid_argument("fp32 mode is implemented by the chain-resident kernel "
                                    "(element-wise targets); the lock-step engine is fp64");
      }
      tick_create(*s, *model);
      WB200_CUDA(cudaStreamSynchronize(s->stream));
      *out = s.release();
      return;
    }
    // slots: resident groups of the chain kernel
    int occ_adapt = 1, occ_sample = 1;
    if (s->kind == kDeviceSource) {
      const size_t dyn = static_cast<size_t>(s->shape.chains_per_cta) *
                         chain_smem_doubles(s->ld) * sizeof(double);
      occ_adapt = std::min(user_blocks_per_sm(s->user->adapt, s->shape.cta, dyn),
                           user_blocks_per_sm(s->user->adapt_free, s->shape.cta, dyn));
      occ_sample = std::min(user_blocks_per_sm(s->user->sample, s->shape.cta, dyn),
                            user_blocks_per_sm(s->user->sample_free, s->shape.cta, dyn));
    } else {
      occupancy_for(s->kind, s->shape, s->ld, s->precision, &occ_adapt, &occ_sample);
    }
    int sms = 0;
    WB200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int need = (s->C + s->shape.chains_per_cta - 1) / s->shape.chains_per_cta;
    s->grid = std::min(occ_sample * sms, need);
    s->grid_adapt = std::min(occ_adapt * sms, need);
    s->slots = std::max(s->grid, s->grid_adapt) * s->shape.chains_per_cta;
    const size_t stride =
        static_cast<size_t>(scratch_doubles(s->tuning.max_trajectory_doublings, s->ld));
    s->scratch.alloc(stride * s->slots);
    WB200_CUDA(cudaMemsetAsync(s->scratch.ptr, 0, stride * s->slots * 8, s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    *out = s.release();
  });
}

void wb200_session_destroy(wb200_session* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->tick) tick_destroy(*s);
  if (s->acc) stream_end(*s);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->tm0) cudaEventDestroy(s->tm0);
  if (s->tm1) cudaEventDestroy(s->tm1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

int wb200_session_init(wb200_session* s, const double* positions,
                       double init_radius, const double* mass,
                       const double* steps, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    const size_t C = s->C;
    const int D = s->D;
    if (positions) {
      for (size_t i = 0; i < C * D; ++i) {
        if (!std::isfinite(positions[i])) throw std::invalid_argument("positions must be finite");
      }
      upload_rows(s->theta.ptr, s->ld, positions, D, C, s->stream);
    } else {
      walnuts_b200::validate::finite_positive(init_radius, "init_scale");
    }
    if (mass) {
      for (size_t i = 0; i < C * D; ++i) {
        walnuts_b200::validate::finite_positive(mass[i], "masses");
      }
      upload_rows(s->inv_mass.ptr, s->ld, mass, D, C, s->stream);
    }
    if (steps) {
      for (size_t i = 0; i < C; ++i) {
        walnuts_b200::validate::finite_positive(steps[i], "step_size");
      }
      WB200_CUDA(cudaMemcpyAsync(s->red.ptr, steps, C * 8, cudaMemcpyHostToDevice,
                                 s->stream));
    }
    if (s->tick) {
      tick_init(*s, mass != nullptr, steps != nullptr, positions != nullptr, init_radius);
    } else {
      launch_init(*s, mass != nullptr, steps != nullptr, positions != nullptr,
                  init_radius);
    }
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    s->initialised = true;
    s->frozen = false;
  });
}

int wb200_session_reserve_draws(wb200_session* s, long long capacity, int trace,
                                WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (capacity < 0) throw std::invalid_argument("capacity must be non-negative");
    const size_t rows = static_cast<size_t>(s->C) * capacity;
    s->draws.alloc(rows * s->ld);
    s->trace = trace != 0;
    if (s->trace) {
      s->lp_out.alloc(rows);
      s->depth_out.alloc(rows);
      s->step_out.alloc(rows);
      s->im_out.alloc(rows * s->ld);
    }
    s->draw_cap = capacity;
    s->rows_written = 0;
    if (!s->tick) s->ragged = false;
  });
}

static void check_room(wb200_session* s, int n_iter, int store) {
  if (!s->initialised) throw std::runtime_error("session is not initialised");
  if (n_iter < 0) throw std::invalid_argument("n_iter must be non-negative");
  if (store && s->ragged && !s->tick) {
    throw std::runtime_error("chains hold different numbers of rows after a free-running "
                             "phase: store further draws with free-running launches");
  }
  if (store && s->rows_written + n_iter > s->draw_cap) {
    throw std::runtime_error("draw buffer too small: reserve more capacity");
  }
}

int wb200_session_warmup(wb200_session* s, int n_iter, int store, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    check_room(s, n_iter, store);
    if (s->frozen) throw std::runtime_error("warm-up after freeze");
    if (n_iter > 0) {
      if (s->tick) tick_run(*s, n_iter, 1, store != 0);
      else launch_chains(*s, n_iter, 1, store != 0);
    }
  });
}

// Free-running launch of the chain-resident engine: every chain gets `budget` gradient
// evaluations and completes the transitions that fit (chain_kernel.cuh, eval_budget);
// draws go to per-chain row counters.
static void chain_free_run(wb200_session* s, long long budget, long long iter_cap, int adapt,
                           bool store) {
  constexpr int kNoIterLimit = 0x7fffffff;
  if (iter_cap <= 0) iter_cap = 0x7fffffffffffffffll;
  if (!store) {
    launch_chains(*s, kNoIterLimit, adapt, false, budget, iter_cap, nullptr);
    return;
  }
  if (s->draw_cap == 0) throw std::runtime_error("reserve draws first");
  if (s->acc && !adapt) {
    // streaming summaries: every chain refills the staging block from row 0 and
    // stream_update folds each chain's rows into its running sums
    stream_flush(*s);
    WB200_CUDA(cudaMemsetAsync(s->acc_rows(), 0, s->C * sizeof(long long), s->stream));
    launch_chains(*s, kNoIterLimit, 0, true, budget, iter_cap, s->acc_rows());
    stream_update(*s, s->acc_rows(), 0);
    return;
  }
  if (!s->ragged) {  // the rows stored so far are uniform: start every counter there
    std::vector<long long> rows(s->C, s->rows_written);
    s->chain_rows.alloc(s->C);
    WB200_CUDA(cudaMemcpyAsync(s->chain_rows.ptr, rows.data(), s->C * sizeof(long long),
                               cudaMemcpyHostToDevice, s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    s->ragged = true;
  }
  launch_chains(*s, kNoIterLimit, adapt, true, budget, iter_cap, s->chain_rows.ptr);
}

int wb200_session_warmup_ticks(wb200_session* s, int n_ticks, int store,
                               WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->initialised) throw std::runtime_error("session is not initialised");
    if (s->frozen) throw std::runtime_error("warm-up after freeze");
    if (n_ticks < 0) throw std::invalid_argument("n_ticks must be non-negative");
    if (store && s->draw_cap == 0) throw std::runtime_error("reserve draws first");
    if (!s->tick) {
      if (n_ticks > 0) chain_free_run(s, n_ticks, 0, 1, store != 0);
      return;
    }
    tick_run_ticks(*s, n_ticks, 1, store != 0);
    if (store) s->ragged = true;
  });
}

int wb200_session_run_evals(wb200_session* s, int sampling, long long eval_budget,
                            long long iter_cap, int store, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->initialised) throw std::runtime_error("session is not initialised");
    if (sampling && !s->frozen) throw std::runtime_error("sample before freeze");
    if (!sampling && s->frozen) throw std::runtime_error("warm-up after freeze");
    if (eval_budget <= 0) throw std::invalid_argument("eval_budget must be positive");
    if (s->tick) {
      // lock-step engine: a tick is one gradient evaluation of every chain
      if (eval_budget > 0x7fffffffll) throw std::invalid_argument("eval_budget too large");
      const int n_ticks = static_cast<int>(eval_budget);
      if (store && s->draw_cap == 0) throw std::runtime_error("reserve draws first");
      if (sampling && s->acc && store) {
        stream_flush(*s);
        tick_run_ticks(*s, n_ticks, 0, true, iter_cap);
        tick_take_rows(*s, s->acc_rows());
        stream_update(*s, s->acc_rows(), 0);
        return;
      }
      tick_run_ticks(*s, n_ticks, sampling ? 0 : 1, store != 0, iter_cap);
      if (store) s->ragged = true;
      return;
    }
    chain_free_run(s, eval_budget, iter_cap, sampling ? 0 : 1, store != 0);
  });
}

int wb200_session_iter_stats(wb200_session* s, int sampling, long long* stats4,
                             WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->initialised) throw std::runtime_error("session is not initialised");
    chain_iter_stats(*s, sampling != 0, stats4);
  });
}

int wb200_session_freeze(wb200_session* s, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->initialised) throw std::runtime_error("session is not initialised");
    if (s->tick) tick_abort_inflight(*s);
    launch_freeze(*s);
    s->frozen = true;
  });
}

int wb200_session_sample(wb200_session* s, int n_iter, int store, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->initialised) throw std::runtime_error("session is not initialised");
    if (!s->frozen) throw std::runtime_error("sample before freeze");
    if (s->acc && store) {
      // streaming summaries: the draw buffer is a staging block that every launch
      // refills from row 0 and stream_update folds into the running sums
      if (n_iter < 0) throw std::invalid_argument("n_iter must be non-negative");
      for (int done = 0; done < n_iter;) {
        const int n = static_cast<int>(std::min<long long>(n_iter - done, s->draw_cap));
        if (s->rows_written + n > s->draw_cap) stream_flush(*s);  // block full: fold it in
        if (s->tick) tick_run(*s, n, 0, true);
        else launch_chains(*s, n, 0, true);
        done += n;
      }
      return;
    }
    check_room(s, n_iter, store);
    if (n_iter > 0) {
      if (s->tick) tick_run(*s, n_iter, 0, store != 0);
      else launch_chains(*s, n_iter, 0, store != 0);
    }
  });
}

int wb200_session_sample_ticks(wb200_session* s, int n_ticks, int store,
                               WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->initialised) throw std::runtime_error("session is not initialised");
    if (!s->frozen) throw std::runtime_error("sample before freeze");
    if (n_ticks < 0) throw std::invalid_argument("n_ticks must be non-negative");
    if (store && s->draw_cap == 0) throw std::runtime_error("reserve draws first");
    if (!s->tick) {
      if (n_ticks > 0) chain_free_run(s, n_ticks, 0, 0, store != 0);
      return;
    }
    if (s->acc && store) {
      // streaming: chains restart their staging rows at 0; a chain that fills the block
      // before the ticks are over idles until the next call (reserve generously)
      stream_flush(*s);
      tick_run_ticks(*s, n_ticks, 0, true);
      tick_take_rows(*s, s->acc_rows());
      stream_update(*s, s->acc_rows(), 0);
      return;
    }
    tick_run_ticks(*s, n_ticks, 0, store != 0);
    if (store) s->ragged = true;
  });
}

int wb200_session_chain_rows(wb200_session* s, long long* rows, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (s->tick) {
      tick_chain_rows(*s, rows);
      for (int c = 0; c < s->C; ++c) rows[c] = std::max(rows[c], s->rows_written);
    } else if (s->ragged) {
      WB200_CUDA(cudaMemcpyAsync(rows, s->chain_rows.ptr, s->C * sizeof(long long),
                                 cudaMemcpyDeviceToHost, s->stream));
      WB200_CUDA(cudaStreamSynchronize(s->stream));
    } else {
      for (int c = 0; c < s->C; ++c) rows[c] = s->rows_written;
    }
  });
}

int wb200_session_summary(wb200_session* s, long long first, double* rhat, double* ess,
                          double* mcse, double* mean, double* var, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    std::vector<long long> rows(s->C), start(s->C), len(s->C);
    WalnutpyError* e = nullptr;
    if (wb200_session_chain_rows(s, rows.data(), &e) != 0) {
      std::string msg = e ? e->msg : "chain_rows failed";
      delete e;
      throw std::runtime_error(msg);
    }
    // chains that have not yet stored 3 draws in the range (long first transitions in
    // free-running mode) are left out, as summary.hpp:595-603 would reject them
    int kept = 0;
    for (int c = 0; c < s->C; ++c) {
      const long long l = rows[c] - first;
      if (l < 3) continue;
      start[kept] = static_cast<long long>(c) * s->draw_cap + first;
      len[kept] = l;
      ++kept;
    }
    if (kept < 2) throw std::invalid_argument("fewer than two chains have 3 draws in the range");
    start.resize(kept);
    len.resize(kept);
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    device_summary(s->draws.ptr, s->ld, s->D, start, len, rhat, ess, mcse, mean, var,
                   s->stream);
  });
}

int wb200_session_rhat_moments(wb200_session* s, long long first, double* moments,
                               WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    std::vector<long long> rows(s->C), start, len;
    WalnutpyError* e = nullptr;
    if (wb200_session_chain_rows(s, rows.data(), &e) != 0) {
      std::string msg = e ? e->msg : "chain_rows failed";
      delete e;
      throw std::runtime_error(msg);
    }
    for (int c = 0; c < s->C; ++c) {
      const long long l = rows[c] - first;
      if (l < 3) continue;
      start.push_back(static_cast<long long>(c) * s->draw_cap + first);
      len.push_back(l);
    }
    if (start.empty()) throw std::invalid_argument("no chain has 3 draws in the range");
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    device_rhat_moments(s->draws.ptr, s->ld, s->D, start, len, moments, s->stream);
  });
}

int wb200_session_sync(wb200_session* s, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
  });
}

int wb200_session_get_draws(wb200_session* s, long long first, long long count,
                            double* out, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    // after a free-running phase chains hold different numbers of rows
    // (wb200_session_chain_rows); any range inside the capacity may be read
    const long long limit = s->ragged ? s->draw_cap : s->rows_written;
    if (first < 0 || count < 0 || first + count > limit) {
      throw std::invalid_argument("draw range out of bounds");
    }
    for (int c = 0; c < s->C; ++c) {
      download_rows(out + static_cast<size_t>(c) * count * s->D, s->D,
                    s->draws.ptr + (static_cast<size_t>(c) * s->draw_cap + first) * s->ld,
                    s->ld, count, s->stream);
    }
    WB200_CUDA(cudaStreamSynchronize(s->stream));
  });
}

int wb200_session_get_trace(wb200_session* s, long long first, long long count,
                            double* lp, int* depth, double* step, double* inv_mass,
                            WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (!s->trace) throw std::runtime_error("session was reserved without trace");
    if (first < 0 || count < 0 || first + count > s->rows_written) {
      throw std::invalid_argument("trace range out of bounds");
    }
    for (int c = 0; c < s->C; ++c) {
      const size_t src = static_cast<size_t>(c) * s->draw_cap + first;
      const size_t dst = static_cast<size_t>(c) * count;
      if (lp) WB200_CUDA(cudaMemcpyAsync(lp + dst, s->lp_out.ptr + src, count * 8,
                                         cudaMemcpyDeviceToHost, s->stream));
      if (depth) WB200_CUDA(cudaMemcpyAsync(depth + dst, s->depth_out.ptr + src, count * 4,
                                            cudaMemcpyDeviceToHost, s->stream));
      if (step) WB200_CUDA(cudaMemcpyAsync(step + dst, s->step_out.ptr + src, count * 8,
                                           cudaMemcpyDeviceToHost, s->stream));
      if (inv_mass) download_rows(inv_mass + dst * s->D, s->D,
                                  s->im_out.ptr + src * s->ld, s->ld, count, s->stream);
    }
    WB200_CUDA(cudaStreamSynchronize(s->stream));
  });
}

int wb200_session_get_state(wb200_session* s, double* theta, double* inv_mass,
                            double* step, int* min_micro,
                            unsigned long long* grad_evals, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    if (theta) download_rows(theta, s->D, s->theta.ptr, s->ld, s->C, s->stream);
    if (inv_mass) download_rows(inv_mass, s->D, s->inv_mass.ptr, s->ld, s->C, s->stream);
    std::vector<ChainScalars> h(s->C);
    WB200_CUDA(cudaMemcpyAsync(h.data(), s->sc.ptr, s->C * sizeof(ChainScalars),
                               cudaMemcpyDeviceToHost, s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    for (int c = 0; c < s->C; ++c) {
      if (step) step[c] = s->frozen ? h[c].step : std::exp(h[c].adam_x);
      if (min_micro) min_micro[c] = h[c].min_micro;
      if (grad_evals) grad_evals[c] = h[c].grad_evals;
    }
  });
}

int wb200_session_device_draws(wb200_session* s, double** draws, long long* capacity,
                               int* ld, long long* rows_written) {
  if (!s) return -1;
  if (draws) *draws = s->draws.ptr;
  if (capacity) *capacity = s->draw_cap;
  if (ld) *ld = s->ld;
  if (rows_written) *rows_written = s->rows_written;
  return 0;
}

int wb200_session_counters(wb200_session* s, unsigned long long* grad_evals,
                           unsigned long long* macro_steps,
                           unsigned long long* kernel_launches, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    std::vector<ChainScalars> h(s->C);
    WB200_CUDA(cudaMemcpyAsync(h.data(), s->sc.ptr, s->C * sizeof(ChainScalars),
                               cudaMemcpyDeviceToHost, s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    unsigned long long g = 0, m = 0;
    for (const auto& c : h) { g += c.grad_evals; m += c.macro_steps; }
    if (grad_evals) *grad_evals = g;
    if (macro_steps) *macro_steps = m;
    if (kernel_launches) *kernel_launches = s->launches;
  });
}

int wb200_session_last_kernel_ms(wb200_session* s, float* ms) {
  if (!s || !ms) return -1;
  cudaSetDevice(s->device);
  if (cudaEventSynchronize(s->ev1) != cudaSuccess) return -1;
  return cudaEventElapsedTime(ms, s->ev0, s->ev1) == cudaSuccess ? 0 : -1;
}

int wb200_session_timer_record(wb200_session* s, int which, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    WB200_CUDA(cudaEventRecord(which == 0 ? s->tm0 : s->tm1, s->stream));
  });
}

int wb200_session_timer_elapsed_ms(wb200_session* s, float* ms, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    WB200_CUDA(cudaEventSynchronize(s->tm1));
    WB200_CUDA(cudaEventElapsedTime(ms, s->tm0, s->tm1));
  });
}

// ---------------------------------------------------------------- orbit ----
int wb200_orbit(const WalnutModelDesc* model, size_t num_chains,
                const double* theta, const double* rho, const double* inv_mass,
                double step, int num_steps, double* theta_out, double* rho_out,
                double* grad_out, double* logp_out, double* joint_out,
                WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    require_gpu();
    const int D = model->D;
    const LaunchShape shape = shape_for_dim(D);
    const int ld = 2 * shape.T * shape.K;
    const size_t C = num_chains, CL = C * ld;
    DeviceBuffer<double> th, rh, im, g, lp, jt, tp;
    th.alloc(CL); rh.alloc(CL); im.alloc(CL); g.alloc(CL); lp.alloc(C); jt.alloc(C);
    cudaStream_t st = nullptr;
    WB200_CUDA(cudaMemset(th.ptr, 0, CL * 8));
    WB200_CUDA(cudaMemset(rh.ptr, 0, CL * 8));
    WB200_CUDA(cudaMemset(im.ptr, 0, CL * 8));
    upload_rows(th.ptr, ld, theta, D, C, st);
    upload_rows(rh.ptr, ld, rho, D, C, st);
    upload_rows(im.ptr, ld, inv_mass, D, C, st);
    if (model->kind == kDiagGaussian) {
      tp.alloc(ld);
      WB200_CUDA(cudaMemset(tp.ptr, 0, ld * 8));
      WB200_CUDA(cudaMemcpy(tp.ptr, model->data0, D * 8, cudaMemcpyHostToDevice));
    }
    std::shared_ptr<UserModule> user;
    if (model->kind == kDeviceSource) {
      if (!model->data0) throw std::invalid_argument("kind 5 needs data0 = CUDA source text");
      int device = 0;
      WB200_CUDA(cudaGetDevice(&device));
      user = user_module(static_cast<const char*>(model->data0), shape, device);
      tp.alloc(std::max<size_t>(model->N, 1));
      if (model->N > 0) {
        WB200_CUDA(cudaMemcpy(tp.ptr, model->data1, model->N * 8, cudaMemcpyHostToDevice));
      }
    }
    launch_orbit(model->kind, model->precision, D, ld, static_cast<int>(C), tp.ptr, th.ptr, rh.ptr,
                 im.ptr, g.ptr, lp.ptr, jt.ptr, step, num_steps, st,
                 user ? user->orbit : nullptr);
    download_rows(theta_out, D, th.ptr, ld, C, st);
    download_rows(rho_out, D, rh.ptr, ld, C, st);
    download_rows(grad_out, D, g.ptr, ld, C, st);
    WB200_CUDA(cudaMemcpy(logp_out, lp.ptr, C * 8, cudaMemcpyDeviceToHost));
    WB200_CUDA(cudaMemcpy(joint_out, jt.ptr, C * 8, cudaMemcpyDeviceToHost));
    WB200_CUDA(cudaDeviceSynchronize());
  });
}

int walnutpie_sample_cfunc(
    LOGP_CFUNC, void*, int, const double*, size_t, unsigned int, unsigned int,
    double, const double*, int, int, int, int, int, int, int, double, double,
    double, double, double, double, double, double, double, double, double,
    double, double, double, bool, double*, size_t, int*, double*, double*, int,
    PRINT_CALLBACK, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    throw std::runtime_error(
        "walnuts_b200 runs device-resident models only: a host LOGP_CFUNC cannot "
        "feed a GPU batch. Describe the model with WalnutModelDesc (a built-in kind, or "
        "kind 4 with a WB200_BATCH_LOGP_GRAD that evaluates all chains on the device) "
        "and call walnutpie_sample_device (same trailing arguments).");
  });
}

// walnutpy.cpp:227-243.  BridgeStan models stay CPU-reference-only (BASELINE.json
// north_star); the symbol exists so that the reference's _ffi.py, which binds it
// unconditionally at import (_ffi.py:235), loads this library unchanged.
int walnutpie_sample_bridgestan(
    const char*, const char*, STREAM_CALLBACK, unsigned int, const char*, size_t,
    unsigned int, unsigned int, double, const double*, int, int, int, int, int, int, int,
    double, double, double, double, double, double, double, double, double, double,
    double, double, double, double, bool, double*, size_t, int*, double*, double*, int,
    PRINT_CALLBACK, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    throw std::runtime_error(
        "walnuts_b200 runs device-resident models only: BridgeStan models are evaluated "
        "by a host library and stay with the CPU reference (walnutpie). Describe the "
        "model with WalnutModelDesc and call walnutpie_sample_device.");
  });
}

}  // extern "C"
