// walnutpie_sample_device_multi: the one-shot call over several GPUs of one process.
//
// The reference's walnutpie::walnuts (api.hpp:33-69) is ONE call that runs all chains and
// both controllers.  Here the chains are sharded over the given devices (contiguous blocks
// of global chain ids, which select the Philox streams: the chains of a G-GPU run are
// exactly those of a 1-GPU run), one host thread drives each device's session, and the
// only exchange is the cross-chain summaries, combined by NCCL all-reduce over NVLink:
//   * warm-up controller (adapt.hpp:186-224): sum_c log M_c[d], sum_c log eps_c, count
//     (SUM), then the two maxima of the deviations from the global geometric means (MAX);
//   * sampling controller (sampler.hpp:132-151): {sum mu, count} (SUM), then the sums
//     about the global mean (SUM) -- the two-pass variance of util.hpp:401-404;
//   * posterior summaries (summary.hpp:594-769): the two phases of the streaming
//     accumulators (stream.cu), (2D + 3) and (3 + T) D doubles (SUM; min_len MIN).
// Every thread sees the same reduced numbers, so all devices take the same decisions.
// NCCL is loaded at run time (libnccl.so.2; the copy already in the process -- e.g.
// PyTorch's -- is reused), so the library has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <exception>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../host/config.hpp"
#include "engine.cuh"

namespace {

// ---- the slice of the NCCL API used (nccl.h:169,181,215,260-286,392) -------------------
using ncclComm_t = struct ncclComm*;
enum : int { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3, kNcclFloat64 = 8 };
struct Nccl {
  void* lib = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* name : names) {
      n.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process?
      if (n.lib) break;
    }
    for (const char* name : names) {
      if (!n.lib) n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    }
    if (!n.lib) return;
    n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(dlsym(n.lib, "ncclCommInitAll"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.lib, "ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.lib, "ncclAllReduce"));
    n.GetErrorString =
        reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.lib, "ncclGetErrorString"));
  });
  if (!n.lib || !n.CommInitAll || !n.CommDestroy || !n.AllReduce) {
    throw std::runtime_error("walnuts_b200: NCCL (libnccl.so.2) could not be loaded; "
                             "walnutpie_sample_device_multi needs it for more than one GPU");
  }
  return n;
}

void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    Nccl& n = nccl();
    throw std::runtime_error(std::string("NCCL error in ") + what + ": " +
                             (n.GetErrorString ? n.GetErrorString(rc) : "?"));
  }
}

// all threads meet here before every collective: if one of them has failed, none enters it
class Rendezvous {
 public:
  explicit Rendezvous(int n) : n_(n) {}
  void fail() {
    std::lock_guard<std::mutex> lk(m_);
    failed_ = true;
    cv_.notify_all();
  }
  // returns false if some rank failed
  bool arrive() {
    std::unique_lock<std::mutex> lk(m_);
    if (failed_) return false;
    const long long gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen_ != gen || failed_; });
    }
    return !failed_;
  }

 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  long long gen_ = 0;
  bool failed_ = false;
};

struct PeerFailed {};

struct Rank {
  int index = 0, device = 0;
  size_t offset = 0, count = 0;
  ncclComm_t comm = nullptr;
  wb200_session* s = nullptr;
  wb200::DeviceBuffer<double> buf;  // reduction scratch on this rank's device
  std::exception_ptr error;
};

struct Shared {
  int world = 1;
  Rendezvous* meet = nullptr;
  std::atomic<int> interrupted{0};
};

// in-place all-reduce of n host doubles through this rank's device buffer
void all_reduce_host(Rank& r, Shared& sh, double* v, size_t n, int op) {
  if (sh.world == 1) return;
  if (!sh.meet->arrive()) throw PeerFailed{};
  if (r.buf.count < n) r.buf.alloc(n);
  WB200_CUDA(cudaMemcpyAsync(r.buf.ptr, v, n * 8, cudaMemcpyHostToDevice, r.s->stream));
  nccl_check(nccl().AllReduce(r.buf.ptr, r.buf.ptr, n, kNcclFloat64, op, r.comm, r.s->stream),
             "ncclAllReduce");
  WB200_CUDA(cudaMemcpyAsync(v, r.buf.ptr, n * 8, cudaMemcpyDeviceToHost, r.s->stream));
  WB200_CUDA(cudaStreamSynchronize(r.s->stream));
}

void all_reduce_device(Rank& r, Shared& sh, double* dev, size_t n, int op) {
  if (sh.world == 1) return;
  if (!sh.meet->arrive()) throw PeerFailed{};
  nccl_check(nccl().AllReduce(dev, dev, n, kNcclFloat64, op, r.comm, r.s->stream),
             "ncclAllReduce");
}

struct Args {
  const WalnutModelDesc* model;
  int num_params;
  const double* inits;
  size_t num_chains;
  unsigned int run_seed;
  double init_radius;
  const double* init_inv_metric;
  WalnutTuning t;
  int max_lags;
  double *mean, *var, *rhat, *ess, *mcse;
  int* truncated;
  int* final_lengths;
  double *stepsize_out, *inv_metric_out;
  int refresh;
  PRINT_CALLBACK print;
};

void check(int rc, WalnutpyError*& e) {
  if (rc != 0) {
    std::string msg = e ? e->msg : "unknown failure";
    const WalnutpyErrorType ty = e ? e->type : wb200_generic;
    delete e;
    e = nullptr;
    if (ty == wb200_config) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
  }
}

void run_rank(Rank& r, Shared& sh, const Args& a, int* warm_done_out, int* samp_done_out) {
  WB200_CUDA(cudaSetDevice(r.device));
  WalnutpyError* e = nullptr;
  const int D = a.num_params;
  const size_t C = r.count;
  check(wb200_session_create(a.model, C, a.run_seed, static_cast<unsigned int>(r.offset), &a.t,
                             r.device, &r.s, &e), e);
  check(wb200_session_init(r.s, a.inits ? a.inits + r.offset * D : nullptr, a.init_radius,
                           a.init_inv_metric ? a.init_inv_metric + r.offset * D : nullptr,
                           nullptr, &e), e);
  const long long stage = std::min<long long>(std::max(a.t.max_sampling_iter, 1), 50);
  check(wb200_session_reserve_draws(r.s, stage, 0, &e), e);
  auto say = [&](const std::string& m) {
    if (a.print && r.index == 0) a.print(m.c_str(), m.size(), false);
  };
  const int stride = a.t.publish_stride > 0 ? a.t.publish_stride : 5;
  auto block = [&](int done, int min_iter, int max_iter) {
    if (r.s->tick && done < min_iter) return min_iter - done;
    if (min_iter == max_iter) return std::min(std::max(stride, 20), max_iter - done);
    return std::min(stride, max_iter - done);
  };
  // Free-running phases (min < max on the chain-resident engine), exactly as in the
  // single-device call (driver.cu): every chain of every device gets the same budget of
  // gradient evaluations per block -- worked out from the all-reduced totals, so all
  // devices compute the same budgets a single device holding all chains would -- and the
  // controllers read the all-reduced {min, max, sum} of the per-chain iteration counts.
  const char* blocks_env = std::getenv("WB200_BLOCKS");
  const bool allow_free = !(blocks_env && std::string(blocks_env) == "uniform");
  double evals_seen = 0, iters_seen = 0;
  auto free_budget = [&] {
    const double per_iter = iters_seen > 0 ? evals_seen / iters_seen : 16.0;
    return std::max<long long>(1, std::llround(per_iter * stride));
  };
  auto global_stats = [&](bool sampling, double (&st)[4]) {
    long long local[4];
    check(wb200_session_iter_stats(r.s, sampling ? 1 : 0, local, &e), e);
    double mn = static_cast<double>(local[0]), mx = static_cast<double>(local[1]);
    double sums2[2] = {static_cast<double>(local[2]), static_cast<double>(local[3])};
    all_reduce_host(r, sh, &mn, 1, kNcclMin);
    all_reduce_host(r, sh, &mx, 1, kNcclMax);
    all_reduce_host(r, sh, sums2, 2, kNcclSum);
    st[0] = mn; st[1] = mx; st[2] = sums2[0]; st[3] = sums2[1];
  };
  auto free_phase = [&](bool sampling, int min_iter, int max_iter, auto&& converged) {
    double st[4];
    global_stats(sampling, st);
    double evals0 = st[3];
    const double iters_before = iters_seen;
    const double all_at_max = static_cast<double>(a.num_chains) * max_iter;
    while (st[2] < all_at_max) {
      check(wb200_session_run_evals(r.s, sampling ? 1 : 0, free_budget(), max_iter,
                                    sampling ? 1 : 0, &e), e);
      if (sh.interrupted.load()) throw wb200::InterruptException();
      global_stats(sampling, st);
      evals_seen += st[3] - evals0;
      evals0 = st[3];
      iters_seen = iters_before + st[2];
      if (st[0] >= min_iter && st[2] < all_at_max && converged()) break;
    }
    return static_cast<int>(st[1]);
  };
  auto chain_counts = [&](std::vector<long long>& out) {  // sampling iterations per chain
    std::vector<wb200::ChainScalars> h(C);
    WB200_CUDA(cudaMemcpyAsync(h.data(), r.s->sc.ptr, C * sizeof(wb200::ChainScalars),
                               cudaMemcpyDeviceToHost, r.s->stream));
    WB200_CUDA(cudaStreamSynchronize(r.s->stream));
    out.resize(C);
    for (size_t c = 0; c < C; ++c) out[c] = static_cast<long long>(h[c].lp_n);
  };
  // ---- warm-up (adapt.hpp:173-229 over all devices' chains)
  int warm_done = 0;
  wb200::DeviceBuffer<double> sums;
  sums.alloc(static_cast<size_t>(D) + 2);
  auto warmup_converged = [&] {
    double dev[2];
    check(wb200_session_warmup_sums(r.s, sums.ptr, &e), e);
    all_reduce_device(r, sh, sums.ptr, static_cast<size_t>(D) + 2, kNcclSum);
    check(wb200_session_warmup_deviation(r.s, sums.ptr, dev, &e), e);
    all_reduce_host(r, sh, dev, 2, kNcclMax);
    return dev[0] <= a.t.mass_converge_tol && dev[1] <= a.t.step_size_converge_tol;
  };
  const bool free_warm = allow_free && a.t.min_warmup_iter < a.t.max_warmup_iter;
  if (free_warm) {
    warm_done = free_phase(false, a.t.min_warmup_iter, a.t.max_warmup_iter, warmup_converged);
  }
  while (!free_warm && warm_done < a.t.max_warmup_iter) {
    const int n = block(warm_done, a.t.min_warmup_iter, a.t.max_warmup_iter);
    check(wb200_session_warmup(r.s, n, 0, &e), e);
    warm_done += n;
    if (sh.interrupted.load()) throw wb200::InterruptException();
    if (warm_done >= a.t.min_warmup_iter && warm_done < a.t.max_warmup_iter) {
      if (warmup_converged()) break;
    }
  }
  check(wb200_session_freeze(r.s, &e), e);
  check(wb200_session_stream_begin(r.s, a.max_lags, &e), e);
  // ---- sampling (sampler.hpp:118-158 over all devices' chains)
  int samp_done = 0;
  auto sampling_converged = [&] {
    double m0[4], m[4];
    check(wb200_session_lp_moments(r.s, m0, &e), e);
    all_reduce_host(r, sh, m0, 4, kNcclSum);
    check(wb200_session_lp_moments_centered(r.s, m0[0] / m0[3], m, &e), e);
    all_reduce_host(r, sh, m, 4, kNcclSum);
    const double M = m[3];
    const double r_hat =
        std::sqrt(1 + ((m[1] - m[0] * m[0] / M) / (M - 1.0)) / (m[2] / M));
    if (a.refresh != 0) {  // handlers.hpp:164-172
      std::stringstream ss;
      ss.precision(10);
      ss << "Controller: R-hat at " << r_hat << std::endl;
      say(ss.str());
    }
    return r_hat <= a.t.rhat_converge_tol;
  };
  const bool free_samp =
      allow_free && (a.t.min_sampling_iter < a.t.max_sampling_iter || free_warm);
  std::vector<long long> samp_rows;
  if (free_samp) {
    samp_done = free_phase(true, a.t.min_sampling_iter, a.t.max_sampling_iter,
                           sampling_converged);
    chain_counts(samp_rows);
  }
  while (!free_samp && samp_done < a.t.max_sampling_iter) {
    const int n = block(samp_done, a.t.min_sampling_iter, a.t.max_sampling_iter);
    check(wb200_session_sample(r.s, n, 1, &e), e);
    samp_done += n;
    if (sh.interrupted.load()) throw wb200::InterruptException();
    if (samp_done >= a.t.min_sampling_iter && samp_done < a.t.max_sampling_iter) {
      if (sampling_converged()) break;
    }
  }
  // ---- posterior summaries over ALL chains (stream.cu phases, two all-reduces)
  std::vector<double> r1(2 * static_cast<size_t>(D) + 3),
      r2(static_cast<size_t>(3 + a.max_lags) * D);
  check(wb200_session_stream_phase1(r.s, r1.data(), &e), e);
  double min_len = r1[2 * D + 2];
  if (min_len <= 0) min_len = 1e300;  // a rank without usable chains must not win the MIN
  all_reduce_host(r, sh, r1.data(), 2 * static_cast<size_t>(D) + 2, kNcclSum);
  all_reduce_host(r, sh, &min_len, 1, kNcclMin);
  r1[2 * D + 2] = min_len;
  check(wb200_session_stream_phase2(r.s, r1.data(), r2.data(), &e), e);
  all_reduce_host(r, sh, r2.data(), r2.size(), kNcclSum);
  if (r.index == 0) {
    check(wb200_stream_finish(D, a.max_lags, r1.data(), r2.data(),
                              a.num_chains > 1 ? a.rhat : nullptr, a.ess, a.mcse, a.mean,
                              a.var, a.truncated, &e), e);
  }
  // ---- per-chain outputs of this shard (walnutpy.cpp:196-221; handlers.hpp:91-100)
  for (size_t c = 0; c < C; ++c) {
    a.final_lengths[r.offset + c] = 0;  // warm-up draws are not saved in this form
    a.final_lengths[a.num_chains + r.offset + c] =
        free_samp ? static_cast<int>(samp_rows[c]) : samp_done;
  }
  check(wb200_session_get_state(
            r.s, nullptr, a.inv_metric_out ? a.inv_metric_out + r.offset * D : nullptr,
            a.stepsize_out ? a.stepsize_out + r.offset : nullptr, nullptr, nullptr, &e), e);
  *warm_done_out = warm_done;
  *samp_done_out = samp_done;
}

}  // namespace

extern "C" int walnutpie_sample_device_multi(
    const int* devices, int num_devices, const WalnutModelDesc* model, int num_params,
    const double* inits, size_t num_chains, unsigned int seed, unsigned int id,
    double init_radius, const double* init_inv_metric, int min_warmup_iter,
    int max_warmup_iter, int min_sampling_iter, int max_sampling_iter,
    int max_trajectory_doublings, int max_step_halvings, int min_micro_steps,
    double max_hamiltonian_error, double step_size_converge_tol, double mass_converge_tol,
    double rhat_converge_tol, double mass_init_count, double mass_additive_smoothing,
    double max_macro_steps_target, double step_size_init, double step_accept_rate_target,
    double step_learning_rate, double step_gradient_decay, double step_sq_gradient_decay,
    double step_stabilization, double step_learn_rate_decay, int max_lags, double* mean_out,
    double* var_out, double* rhat_out, double* ess_out, double* mcse_out, int* truncated_out,
    int* final_lengths, double* stepsize_out, double* inv_metric_out, int refresh,
    PRINT_CALLBACK print, WalnutpyError** err) {
  using namespace wb200;
  return catch_exceptions(err, [&] {
    if (refresh < 0) {
      std::stringstream msg;
      msg << "refresh must be non-negative, was " << refresh;
      throw std::invalid_argument(msg.str());
    }
    if (!model) throw std::invalid_argument("model descriptor is null");
    if (model->D != num_params) {
      throw std::invalid_argument("model dimension and num_params differ");
    }
    if (!devices || num_devices < 1) throw std::invalid_argument("no devices given");
    if (static_cast<size_t>(num_devices) > num_chains) {
      throw std::invalid_argument("more devices than chains");
    }
    require_gpu();
    Args a{};
    a.model = model; a.num_params = num_params; a.inits = inits; a.num_chains = num_chains;
    a.run_seed = seed + id + static_cast<unsigned int>(num_chains);  // walnutpy.cpp:82
    a.init_radius = init_radius; a.init_inv_metric = init_inv_metric;
    walnuts_b200_default_tuning(&a.t);
    a.t.min_warmup_iter = min_warmup_iter; a.t.max_warmup_iter = max_warmup_iter;
    a.t.min_sampling_iter = min_sampling_iter; a.t.max_sampling_iter = max_sampling_iter;
    a.t.max_trajectory_doublings = max_trajectory_doublings;
    a.t.max_step_halvings = max_step_halvings; a.t.min_micro_steps = min_micro_steps;
    a.t.max_hamiltonian_error = max_hamiltonian_error;
    a.t.step_size_converge_tol = step_size_converge_tol;
    a.t.mass_converge_tol = mass_converge_tol; a.t.rhat_converge_tol = rhat_converge_tol;
    a.t.mass_init_count = mass_init_count;
    a.t.mass_additive_smoothing = mass_additive_smoothing;
    a.t.max_macro_steps_target = max_macro_steps_target; a.t.step_size_init = step_size_init;
    a.t.step_accept_rate_target = step_accept_rate_target;
    a.t.step_learning_rate = step_learning_rate; a.t.step_gradient_decay = step_gradient_decay;
    a.t.step_sq_gradient_decay = step_sq_gradient_decay;
    a.t.step_stabilization = step_stabilization;
    a.t.step_learn_rate_decay = step_learn_rate_decay;
    a.max_lags = max_lags;
    a.mean = mean_out; a.var = var_out; a.rhat = rhat_out; a.ess = ess_out; a.mcse = mcse_out;
    a.truncated = truncated_out; a.final_lengths = final_lengths;
    a.stepsize_out = stepsize_out; a.inv_metric_out = inv_metric_out;
    a.refresh = refresh; a.print = print;

    const int G = num_devices;
    std::vector<Rank> ranks(G);
    const size_t base = num_chains / G, rem = num_chains % G;
    for (int g = 0; g < G; ++g) {
      ranks[g].index = g;
      ranks[g].device = devices[g];
      ranks[g].count = base + (static_cast<size_t>(g) < rem ? 1 : 0);
      ranks[g].offset = g * base + std::min<size_t>(g, rem);
    }
    std::vector<ncclComm_t> comms(G, nullptr);
    if (G > 1) {
      nccl_check(nccl().CommInitAll(comms.data(), G, devices), "ncclCommInitAll");
      for (int g = 0; g < G; ++g) ranks[g].comm = comms[g];
    }
    Rendezvous meet(G);
    Shared sh;
    sh.world = G;
    sh.meet = &meet;
    std::vector<int> warm(G, 0), samp(G, 0);
    std::vector<std::thread> threads;
    for (int g = 0; g < G; ++g) {
      threads.emplace_back([&, g] {
        try {
          run_rank(ranks[g], sh, a, &warm[g], &samp[g]);
        } catch (const PeerFailed&) {
        } catch (...) {
          ranks[g].error = std::current_exception();
          meet.fail();
        }
        if (ranks[g].s) {
          cudaSetDevice(ranks[g].device);
          ranks[g].buf.release();
          wb200_session_destroy(ranks[g].s);
          ranks[g].s = nullptr;
        }
      });
    }
    for (auto& t : threads) t.join();
    for (int g = 0; g < G; ++g) {
      if (comms[g]) nccl().CommDestroy(comms[g]);
    }
    for (int g = 0; g < G; ++g) {
      if (ranks[g].error) std::rethrow_exception(ranks[g].error);
    }
  });
}
