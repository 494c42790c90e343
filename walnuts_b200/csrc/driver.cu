// Cross-chain controllers and the one-shot entry point.
//
// Reference: the warm-up controller adapt.hpp:173-229 and the sampling
// controller sampler.hpp:118-158 poll lock-free per-chain snapshots from a host
// thread.  Here the per-chain statistics are reduced on the device and only a
// few scalars cross PCIe per check; with several GPUs the per-GPU sums are the
// payload of one NCCL all-reduce (done by the caller between the two phases).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <chrono>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>

#include "../host/config.hpp"
#include "engine.cuh"

namespace wb200 {

// sums[d] = sum_c log M_c[d]; sums[D] = sum_c log eps_c; sums[D+1] = C.
// One thread per dimension walks the chains in ascending order, so a single-GPU
// result equals the reference's sequential accumulation (adapt.hpp:195-206).
__global__ void warmup_sums_kernel(ChainParams p, double* sums) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < p.D) {
    double acc = 0.0;
    for (int c = 0; c < p.C; ++c) {
      const double* est_row = p.est + static_cast<long long>(c) * 4 * p.ld;
      const double w = p.sc[c].est_w;
      const double im = metric_from_sums(est_row[1 * p.ld + d], est_row[3 * p.ld + d], w);
      acc += -log(im);  // AdaptiveWalnuts::log_mass, adaptive_walnuts.hpp:320-323
    }
    sums[d] = acc;
  }
  if (d == 0) {
    double acc = 0.0;
    for (int c = 0; c < p.C; ++c) acc += log(exp(p.sc[c].adam_x));  // :312
    sums[p.D] = acc;
    sums[p.D + 1] = static_cast<double>(p.C);
  }
}

__device__ __forceinline__ void atomic_fmax_nonneg(double* addr, double v) {
  // std::fmax against a running maximum that starts at 0.0 (adapt.hpp:209-217):
  // NaN and negative candidates never win
  if (v > 0.0) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
  }
}

// out[0] = max_c ||(M_c - gm)/gm||_2, out[1] = max_c (eps_c - gs)/gs
__global__ void warmup_deviation_kernel(ChainParams p, const double* sums,
                                        double* out) {
  __shared__ double red[32];
  const int c = blockIdx.x;
  const double count = sums[p.D + 1];
  const double* est_row = p.est + static_cast<long long>(c) * 4 * p.ld;
  const double w = p.sc[c].est_w;
  double acc = 0.0;
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    const double gm = exp(sums[d] / count);
    const double im = metric_from_sums(est_row[1 * p.ld + d], est_row[3 * p.ld + d], w);
    const double mass = exp(-log(im));  // snap.mass = exp(log_mass), adapt.hpp:137
    const double r = (mass - gm) / gm;
    acc += r * r;
  }
  for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) s += red[i];
    atomic_fmax_nonneg(out + 0, sqrt(s));
    const double gs = exp(sums[p.D] / count);
    const double step = exp(log(exp(p.sc[c].adam_x)));
    atomic_fmax_nonneg(out + 1, (step - gs) / gs);
  }
}

// per-chain (mean, unbiased variance, count) of lp -> pinned-size arrays
__global__ void lp_stats_kernel(ChainParams p, double* mean, double* var, double* cnt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const ChainScalars& sc = p.sc[c];
  mean[c] = sc.lp_mean;
  var[c] = sc.lp_n > 1 ? sc.lp_m2 / static_cast<double>(sc.lp_n - 1) : nan("");
  cnt[c] = static_cast<double>(sc.lp_n);
}

__global__ void philox_kernel(const uint32_t* in6, size_t n, uint32_t* out4) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t* q = in6 + 6 * i;
  Philox4 r = philox4x32_10(q[0], q[1], q[2], q[3], q[4], q[5]);
  out4[4 * i + 0] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
}

__global__ void philox_normals_kernel(uint32_t seed, uint32_t chain, uint32_t iter,
                                      uint32_t kind, size_t n, double* out) {
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (2 * j >= n) return;
  double z0, z1;
  philox_normal_pair(seed, chain, iter, kind, static_cast<uint32_t>(j), z0, z1);
  out[2 * j] = z0;
  if (2 * j + 1 < n) out[2 * j + 1] = z1;
}

}  // namespace wb200

using namespace wb200;

namespace {
struct LastRun {
  unsigned long long grad_evals = 0, macro_steps = 0, launches = 0;
  int warmup_iters = 0, sampling_iters = 0;
} g_last_run;
}  // namespace

// interrupts.hpp:20-104: while a one-shot run is in flight SIGINT only raises a flag; the
// controller checks it between blocks of iterations (adapt.hpp:227, sampler.hpp:154) and
// the run ends with an `interrupt` error (KeyboardInterrupt in Python).  RAII restores
// the previous handler.
static volatile std::sig_atomic_t g_interrupted = 0;
class InterruptHandler {
 public:
  InterruptHandler() {
    g_interrupted = 0;
    std::memset(&custom_, 0, sizeof(custom_));
    sigemptyset(&custom_.sa_mask);
    sigaddset(&custom_.sa_mask, SIGINT);
    custom_.sa_flags = SA_RESETHAND;
    custom_.sa_handler = [](int) { g_interrupted = 1; };
    installed_ = sigaction(SIGINT, &custom_, &before_) == 0;
  }
  ~InterruptHandler() {
    if (installed_) sigaction(SIGINT, &before_, nullptr);
  }
  InterruptHandler(const InterruptHandler&) = delete;
  InterruptHandler& operator=(const InterruptHandler&) = delete;
  void throw_if_interrupted() const {
    if (g_interrupted) throw wb200::InterruptException();
  }

 private:
  struct sigaction before_, custom_;
  bool installed_ = false;
};

extern "C" {

int wb200_last_run_stats(unsigned long long* grad_evals, unsigned long long* macro_steps,
                         unsigned long long* kernel_launches, int* warmup_iters,
                         int* sampling_iters) {
  if (grad_evals) *grad_evals = g_last_run.grad_evals;
  if (macro_steps) *macro_steps = g_last_run.macro_steps;
  if (kernel_launches) *kernel_launches = g_last_run.launches;
  if (warmup_iters) *warmup_iters = g_last_run.warmup_iters;
  if (sampling_iters) *sampling_iters = g_last_run.sampling_iters;
  return 0;
}

int wb200_session_warmup_sums(wb200_session* s, double* sums_device,
                              WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    ChainParams p = s->params(0, 1, false);
    if (!sums_device) {  // single-GPU callers: the session's own buffer
      if (s->sums.count == 0) s->sums.alloc(static_cast<size_t>(s->D) + 2);
      sums_device = s->sums.ptr;
    }
    warmup_sums_kernel<<<(s->D + 127) / 128, 128, 0, s->stream>>>(p, sums_device);
    WB200_CUDA(cudaGetLastError());
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 1;
  });
}

int wb200_session_warmup_deviation(wb200_session* s, const double* sums_device,
                                   double* out_host2, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    ChainParams p = s->params(0, 1, false);
    if (!sums_device) {
      if (s->sums.count == 0) throw std::runtime_error("warmup_sums has not been called");
      sums_device = s->sums.ptr;
    }
    double* out = s->red.ptr;  // 2 doubles of scratch
    WB200_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(double), s->stream));
    warmup_deviation_kernel<<<s->C, 256, 0, s->stream>>>(p, sums_device, out);
    WB200_CUDA(cudaGetLastError());
    WB200_CUDA(cudaMemcpyAsync(out_host2, out, 2 * sizeof(double),
                               cudaMemcpyDeviceToHost, s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 1;
  });
}

// {sum mu, sum mu^2, sum var, count of chains, min lp count} over local chains:
// the payload of the sampling R-hat (sampler.hpp:132-151)
int wb200_session_lp_moments(wb200_session* s, double* moments_host4,
                             WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    ChainParams p = s->params(0, 0, false);
    DeviceBuffer<double> buf;
    buf.alloc(3 * static_cast<size_t>(s->C));
    lp_stats_kernel<<<(s->C + 255) / 256, 256, 0, s->stream>>>(
        p, buf.ptr, buf.ptr + s->C, buf.ptr + 2 * s->C);
    WB200_CUDA(cudaGetLastError());
    std::vector<double> h(3 * static_cast<size_t>(s->C));
    WB200_CUDA(cudaMemcpyAsync(h.data(), buf.ptr, h.size() * 8, cudaMemcpyDeviceToHost,
                               s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 1;
    double sm = 0, sm2 = 0, sv = 0;
    for (int c = 0; c < s->C; ++c) {
      sm += h[c];
      sm2 += h[c] * h[c];
      sv += h[s->C + c];
    }
    moments_host4[0] = sm; moments_host4[1] = sm2; moments_host4[2] = sv;
    moments_host4[3] = static_cast<double>(s->C);
  });
}

// The same payload about a caller-chosen centre: {sum (mu - c), sum (mu - c)^2, sum var,
// count}.  util.hpp:401-404 computes the variance of the chain means in two passes; with
// |lp| ~ 1e5 (logistic N = 100k) the one-pass form sum mu^2 - (sum mu)^2 / M cancels
// heavily, the centred one does not.  Multi-GPU callers all-reduce the plain sums first
// and pass the global mean of the means as the centre.
int wb200_session_lp_moments_centered(wb200_session* s, double center,
                                      double* moments_host4, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    WB200_CUDA(cudaSetDevice(s->device));
    ChainParams p = s->params(0, 0, false);
    DeviceBuffer<double> buf;
    buf.alloc(3 * static_cast<size_t>(s->C));
    lp_stats_kernel<<<(s->C + 255) / 256, 256, 0, s->stream>>>(
        p, buf.ptr, buf.ptr + s->C, buf.ptr + 2 * s->C);
    WB200_CUDA(cudaGetLastError());
    std::vector<double> h(3 * static_cast<size_t>(s->C));
    WB200_CUDA(cudaMemcpyAsync(h.data(), buf.ptr, h.size() * 8, cudaMemcpyDeviceToHost,
                               s->stream));
    WB200_CUDA(cudaStreamSynchronize(s->stream));
    s->launches += 1;
    double sm = 0, sm2 = 0, sv = 0;
    for (int c = 0; c < s->C; ++c) {
      const double d = h[c] - center;
      sm += d;
      sm2 += d * d;
      sv += h[s->C + c];
    }
    moments_host4[0] = sm; moments_host4[1] = sm2; moments_host4[2] = sv;
    moments_host4[3] = static_cast<double>(s->C);
  });
}

int wb200_session_logp_exceptions(wb200_session* s, unsigned long long* count) {
  if (!s || !count) return -1;
  *count = s->logp_exceptions;
  return 0;
}

int wb200_philox(const uint32_t* ctr_key6, size_t n, uint32_t* out4, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    DeviceBuffer<uint32_t> in, out;
    in.alloc(6 * n); out.alloc(4 * n);
    WB200_CUDA(cudaMemcpy(in.ptr, ctr_key6, 6 * n * 4, cudaMemcpyHostToDevice));
    philox_kernel<<<static_cast<unsigned>((n + 127) / 128), 128>>>(in.ptr, n, out.ptr);
    WB200_CUDA(cudaGetLastError());
    WB200_CUDA(cudaMemcpy(out4, out.ptr, 4 * n * 4, cudaMemcpyDeviceToHost));
  });
}

int wb200_philox_normals(unsigned int seed, unsigned int chain, unsigned int iter,
                         unsigned int kind, size_t n, double* out, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    DeviceBuffer<double> buf;
    buf.alloc(n);
    const size_t pairs = (n + 1) / 2;
    philox_normals_kernel<<<static_cast<unsigned>((pairs + 127) / 128), 128>>>(
        seed, chain, iter, kind, n, buf.ptr);
    WB200_CUDA(cudaGetLastError());
    WB200_CUDA(cudaMemcpy(out, buf.ptr, n * 8, cudaMemcpyDeviceToHost));
  });
}

// ---------------------------------------------------------------------------
// walnutpie_sample_device: the whole of walnutpy.cpp:134-222 + run_sampler
// (:20-84) + walnutpie::walnuts (api.hpp:33-69) for a device model.
namespace {
// outputs of the summaries-only form of the one-shot call (any pointer may be null)
struct StreamOutputs {
  int max_lags;
  double *mean, *var, *rhat, *ess, *mcse;
  int* truncated;
};
}  // namespace

// so == nullptr: the reference-shaped call, every draw copied to `out`;
// so != nullptr: no draw leaves the device, streaming summaries are returned instead
static int sample_device_impl(
    const StreamOutputs* so,
    const WalnutModelDesc* model, int num_params, const double* inits,
    size_t num_chains, unsigned int seed, unsigned int id, double init_radius,
    const double* init_inv_metric, int min_warmup_iter, int max_warmup_iter,
    int min_sampling_iter, int max_sampling_iter, int max_trajectory_doublings,
    int max_step_halvings, int min_micro_steps, double max_hamiltonian_error,
    double step_size_converge_tol, double mass_converge_tol,
    double rhat_converge_tol, double mass_init_count,
    double mass_additive_smoothing, double max_macro_steps_target,
    double step_size_init, double step_accept_rate_target,
    double step_learning_rate, double step_gradient_decay,
    double step_sq_gradient_decay, double step_stabilization,
    double step_learn_rate_decay, bool save_warmup, double* out,
    size_t out_size, int* final_lengths, double* stepsize_out,
    double* inv_metric_out, int refresh, PRINT_CALLBACK print,
    WalnutpyError** err) {
  wb200_session* s = nullptr;
  InterruptHandler interrupt;  // walnutpy.cpp:34
  int rc = catch_exceptions(err, [&] {
    if (refresh < 0) {  // errors.hpp:74-81
      std::stringstream msg;
      msg << "refresh must be non-negative, was " << refresh;
      throw std::invalid_argument(msg.str());
    }
    if (!model) throw std::invalid_argument("model descriptor is null");
    if (model->D != num_params) {
      throw std::invalid_argument("model dimension and num_params differ");
    }
    const size_t C = num_chains;
    const size_t rows_per_chain =
        static_cast<size_t>(max_sampling_iter) +
        (save_warmup ? static_cast<size_t>(max_warmup_iter) : 0);
    const size_t draws_offset = static_cast<size_t>(num_params) * rows_per_chain;
    if (so && save_warmup) {
      throw std::invalid_argument("save_warmup needs the draw buffer of "
                                  "walnutpie_sample_device");
    }
    if (!so && out_size < C * draws_offset) {  // walnutpy.cpp:153-160
      std::stringstream ss;
      ss << "Output buffer too small. Expected at least " << C << " chains of "
         << draws_offset << " doubles, got " << out_size;
      throw std::runtime_error(ss.str());
    }
    WalnutTuning t;
    walnuts_b200_default_tuning(&t);
    t.min_warmup_iter = min_warmup_iter; t.max_warmup_iter = max_warmup_iter;
    t.min_sampling_iter = min_sampling_iter; t.max_sampling_iter = max_sampling_iter;
    t.max_trajectory_doublings = max_trajectory_doublings;
    t.max_step_halvings = max_step_halvings;
    t.min_micro_steps = min_micro_steps;
    t.max_hamiltonian_error = max_hamiltonian_error;
    t.step_size_converge_tol = step_size_converge_tol;
    t.mass_converge_tol = mass_converge_tol;
    t.rhat_converge_tol = rhat_converge_tol;
    t.mass_init_count = mass_init_count;
    t.mass_additive_smoothing = mass_additive_smoothing;
    t.max_macro_steps_target = max_macro_steps_target;
    t.step_size_init = step_size_init;
    t.step_accept_rate_target = step_accept_rate_target;
    t.step_learning_rate = step_learning_rate;
    t.step_gradient_decay = step_gradient_decay;
    t.step_sq_gradient_decay = step_sq_gradient_decay;
    t.step_stabilization = step_stabilization;
    t.step_learn_rate_decay = step_learn_rate_decay;

    auto check = [&](int r, WalnutpyError*& e) {
      if (r != 0) {
        std::string msg = e ? e->msg : "unknown failure";
        WalnutpyErrorType ty = e ? e->type : wb200_generic;
        delete e;
        e = nullptr;
        if (ty == wb200_config) throw std::invalid_argument(msg);
        throw std::runtime_error(msg);
      }
    };
    WalnutpyError* e = nullptr;
    const bool trace_phases = std::getenv("WB200_TRACE_PHASES") != nullptr;
    auto t_phase = std::chrono::steady_clock::now();
    auto phase = [&](const char* name) {  // host-side phase timing, diagnostics only
      if (!trace_phases) return;
      if (s) cudaStreamSynchronize(s->stream);
      auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[wb200] %-12s %8.2f ms\n", name,
                   std::chrono::duration<double, std::milli>(now - t_phase).count());
      t_phase = now;
    };
    // per-chain streams are keyed by (seed + id + num_chains, chain): the same
    // mixing of seed and id as walnutpy.cpp:82
    const unsigned int run_seed = seed + id + static_cast<unsigned int>(num_chains);
    // like any CUDA library call, the run uses the calling thread's current device
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    check(wb200_session_create(model, C, run_seed, 0, &t, device, &s, &e), e);
    check(wb200_session_init(s, inits, init_radius, init_inv_metric, nullptr, &e), e);
    // summaries-only: the draw buffer is a staging block folded into running sums
    // (as many rows as fit in 6 GB, at least 50: every fold of the staging block reads and
    // writes the chains' lag accumulators -- 4 GB at c2 -- so fewer, larger folds are cheaper)
    const long long row_bytes = static_cast<long long>(C) * s->ld * 8;
    const long long stage_rows = std::min<long long>(
        std::max(max_sampling_iter, 1),
        std::max<long long>(50, std::min<long long>(1000, (6ll << 30) / row_bytes)));
    check(wb200_session_reserve_draws(
              s, so ? stage_rows : static_cast<long long>(rows_per_chain), 0, &e), e);

    phase("setup+init");
    auto say = [&](const std::string& m) {
      if (print) print(m.c_str(), m.size(), false);
    };
    auto progress = [&](int from, int to, bool warm) {  // handlers.hpp:38-48
      if (refresh == 0 || !print) return;
      for (int it = from + 1; it <= to; ++it) {
        if (it % refresh != 0) continue;
        for (size_t c = 0; c < C; ++c) {
          std::stringstream ss;
          ss << "Chain [" << (c + 1) << "]: Iteration " << it << "\t"
             << (warm ? "(Warmup)" : "(Sampling)") << std::endl;
          say(ss.str());
        }
      }
    };

    // handlers.hpp:30-36 (print_exception), for a batched density: one line per failed
    // batch evaluation; the chains went on with logp = -inf (util.hpp:336-346)
    size_t exceptions_reported = 0;
    auto report_exceptions = [&] {
      for (; exceptions_reported < s->exception_log.size(); ++exceptions_reported) {
        const auto& ex = s->exception_log[exceptions_reported];
        std::stringstream ss;
        ss << "Chains [1-" << C << "]:Error evaluating the log density during tick "
           << ex.first + 1 << ": logp failed with code " << ex.second << std::endl;
        say(ss.str());
      }
    };

    // Draw read-back overlaps sampling: the rows a block of iterations stored are
    // copied on a second stream while the next block runs (one strided 3-D copy
    // per block: D doubles x rows x chains).  The copy of block k is issued
    // after block k+1 is launched so that it overlaps even when `out` is
    // pageable and the copy call blocks the host; with a pinned `out` (see
    // wb200_host_alloc) it is a plain DMA at PCIe rate.
    WB200_CUDA(cudaSetDevice(s->device));
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t block_done = nullptr;
    WB200_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    WB200_CUDA(cudaEventCreateWithFlags(&block_done, cudaEventDisableTiming));
    struct Cleanup {
      cudaStream_t& st; cudaEvent_t& ev;
      ~Cleanup() {
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        if (ev) cudaEventDestroy(ev);
      }
    } cleanup{copy_stream, block_done};
    // at most two blocks of iterations are in flight, so that the host loop (progress
    // lines, interrupt checks) stays in step with the device
    cudaEvent_t ring[2] = {nullptr, nullptr};
    WB200_CUDA(cudaEventCreateWithFlags(&ring[0], cudaEventDisableTiming));
    WB200_CUDA(cudaEventCreateWithFlags(&ring[1], cudaEventDisableTiming));
    struct RingCleanup {
      cudaEvent_t* r;
      ~RingCleanup() { cudaEventDestroy(r[0]); cudaEventDestroy(r[1]); }
    } ring_cleanup{ring};
    long long blocks_launched = 0;
    auto throttle = [&] {  // call before launching a block
      if (blocks_launched >= 2) WB200_CUDA(cudaEventSynchronize(ring[blocks_launched % 2]));
    };
    auto launched = [&] {  // call after launching a block
      WB200_CUDA(cudaEventRecord(ring[blocks_launched % 2], s->stream));
      ++blocks_launched;
    };
    long long pending_from = 0, pending_to = 0;  // rows [from, to) stored, not yet copied
    bool pending_marked = false;
    auto mark_block = [&](long long rows_now) {  // after launching a storing block
      pending_to = rows_now;
      WB200_CUDA(cudaEventRecord(block_done, s->stream));
      pending_marked = true;
    };
    auto flush_pending = [&] {  // copy every row whose block has been marked
      if (!pending_marked || pending_to <= pending_from) return;
      WB200_CUDA(cudaStreamWaitEvent(copy_stream, block_done, 0));
      const long long n_rows = pending_to - pending_from;
      const size_t src_pitch = static_cast<size_t>(s->draw_cap) * s->ld * sizeof(double);
      const size_t dst_pitch = draws_offset * sizeof(double);
      const size_t max_pitch = (size_t{1} << 31) - 1;  // cudaDeviceProp::memPitch
      if (s->ld == num_params && src_pitch <= max_pitch && dst_pitch <= max_pitch) {
        // rows of a chain are contiguous on both sides: one 2-D copy whose
        // "row" is the chain's whole block
        WB200_CUDA(cudaMemcpy2DAsync(
            out + static_cast<size_t>(pending_from) * num_params,
            dst_pitch,
            s->draws.ptr + static_cast<size_t>(pending_from) * s->ld,
            src_pitch,
            static_cast<size_t>(n_rows) * num_params * sizeof(double), C,
            cudaMemcpyDeviceToHost, copy_stream));
      } else {
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(s->draws.ptr, s->ld * sizeof(double),
                                        num_params * sizeof(double), s->draw_cap);
        cp.srcPos = make_cudaPos(0, pending_from, 0);
        cp.dstPtr = make_cudaPitchedPtr(out, num_params * sizeof(double),
                                        num_params * sizeof(double), rows_per_chain);
        cp.dstPos = make_cudaPos(0, pending_from, 0);
        cp.extent = make_cudaExtent(num_params * sizeof(double), n_rows, C);
        cp.kind = cudaMemcpyDeviceToHost;
        WB200_CUDA(cudaMemcpy3DAsync(&cp, copy_stream));
      }
      pending_from = pending_to;
      pending_marked = false;
    };

    // ---- warm-up: blocks of publish_stride iterations, controller between
    const int stride = t.publish_stride > 0 ? t.publish_stride : 5;
    int warm_done = 0;
    DeviceBuffer<double> sums;
    sums.alloc(static_cast<size_t>(num_params) + 2);
    // The lock-step engine idles finished chains until the slowest one has done its
    // quota, so it gets the longest block that needs no controller decision.
    // With min == max there is no controller decision to take: longer launches (less tail
    // imbalance per iteration), still short enough for progress lines and Ctrl+C.
    // Fixed-length phases of the chain engine grow their blocks while a block takes less
    // than 25 ms (the time between two throttle() returns once two blocks are in flight):
    // every launch ends with the tail of its slowest chain, which long launches amortise
    // (c3: 15 blocks of 20 warm-up iterations 180-210 ms, one launch of 300 141 ms),
    // while Ctrl+C and progress lines stay responsive.
    int fixed_block = std::max(stride, 20);
    auto last_throttle = std::chrono::steady_clock::now();
    long long throttles = 0;
    auto block = [&](int done, int min_iter, int max_iter, bool copying) {
      if (s->tick && done < min_iter) return min_iter - done;
      if (min_iter == max_iter) {
        // (blocks whose draws are copied out stay short: the copy of the last block is
        // the one that cannot overlap)
        if (!s->tick && !copying) {
          const auto now = std::chrono::steady_clock::now();
          const double ms = std::chrono::duration<double, std::milli>(now - last_throttle).count();
          last_throttle = now;
          if (++throttles > 2 && ms < 25.0 && fixed_block < 1280) fixed_block *= 2;
        }
        return std::min(copying ? std::max(stride, 20) : fixed_block, max_iter - done);
      }
      return std::min(stride, max_iter - done);
    };
    // Free-running phases (min < max, both engines; on the lock-step engine a budget of
    // gradient evaluations is that many ticks): as in the reference, where
    // every chain is a thread that runs at its own pace until the controller stops it
    // (adapt.hpp:110-129, sampler.hpp:79-94), chains advance by equal WORK per block -- a
    // budget of gradient evaluations worth `stride` average iterations -- and not by equal
    // iteration counts, so the final lengths differ from chain to chain and no block waits
    // for the chain with the longest orbits.  The controllers keep the reference's rules:
    // no decision before every chain has done min_iter (adapt.hpp:196-203), stop on
    // convergence or when every chain has reached max_iter (:219-221).
    // WB200_BLOCKS=uniform restores blocks of equal iteration counts.
    const char* blocks_env = std::getenv("WB200_BLOCKS");
    const bool allow_free = !(blocks_env && std::string(blocks_env) == "uniform");
    long long evals_seen = 0, iters_seen = 0;  // all chains, both phases: the mean cost
    long long warm_max = 0;
    std::vector<long long> warm_rows;  // per chain, when the warm-up ran free
    auto free_budget = [&] {
      const double per_iter = iters_seen > 0 ? static_cast<double>(evals_seen) /
                                                   static_cast<double>(iters_seen)
                                             : 16.0;
      return std::max<long long>(1, std::llround(per_iter * stride));
    };
    // per-chain progress lines (handlers.hpp:38-48) of a free-running block
    std::vector<long long> count_before, count_after;
    auto chain_counts = [&](bool sampling, std::vector<long long>& out) {
      std::vector<wb200::ChainScalars> h(C);
      WB200_CUDA(cudaMemcpyAsync(h.data(), s->sc.ptr, C * sizeof(wb200::ChainScalars),
                                 cudaMemcpyDeviceToHost, s->stream));
      WB200_CUDA(cudaStreamSynchronize(s->stream));
      out.resize(C);
      for (size_t c = 0; c < C; ++c) {
        out[c] = sampling ? static_cast<long long>(h[c].lp_n)
                          : static_cast<long long>(h[c].warm_iter);
      }
    };
    auto progress_ragged = [&](bool sampling) {
      if (refresh == 0 || !print) return;
      chain_counts(sampling, count_after);
      if (count_before.size() != C) count_before.assign(C, 0);
      for (size_t c = 0; c < C; ++c) {
        // sampling iterations are numbered on from the chain's own warm-up
        const long long offset = !sampling ? 0 : warm_rows.size() == C ? warm_rows[c] : warm_max;
        for (long long it = count_before[c] + 1; it <= count_after[c]; ++it) {
          if ((offset + it) % refresh != 0) continue;
          std::stringstream ss;
          ss << "Chain [" << (c + 1) << "]: Iteration " << (offset + it) << "\t"
             << (sampling ? "(Sampling)" : "(Warmup)") << std::endl;
          say(ss.str());
        }
      }
      count_before = count_after;
    };
    // one free-running phase; returns the largest per-chain iteration count
    auto free_phase = [&](bool sampling, int min_iter, int max_iter, bool store,
                          auto&& converged) -> long long {
      long long st[4] = {0, 0, 0, 0};
      long long evals0 = 0;
      check(wb200_session_iter_stats(s, sampling ? 1 : 0, st, &e), e);
      evals0 = st[3];
      const long long iters_before = iters_seen;
      count_before.assign(C, 0);
      while (st[2] < static_cast<long long>(C) * max_iter) {
        throttle();
        check(wb200_session_run_evals(s, sampling ? 1 : 0, free_budget(), max_iter,
                                      store ? 1 : 0, &e), e);
        launched();
        progress_ragged(sampling);
        report_exceptions();
        interrupt.throw_if_interrupted();  // adapt.hpp:227, sampler.hpp:154
        check(wb200_session_iter_stats(s, sampling ? 1 : 0, st, &e), e);
        evals_seen += st[3] - evals0;
        evals0 = st[3];
        iters_seen = iters_before + st[2];
        if (st[0] >= min_iter && st[2] < static_cast<long long>(C) * max_iter &&
            converged()) {
          break;
        }
      }
      return st[1];
    };
    const bool free_warm = allow_free && min_warmup_iter < max_warmup_iter;
    if (free_warm) {
      warm_max = free_phase(false, min_warmup_iter, max_warmup_iter, save_warmup, [&] {
        double dev[2];
        check(wb200_session_warmup_sums(s, sums.ptr, &e), e);
        check(wb200_session_warmup_deviation(s, sums.ptr, dev, &e), e);
        return dev[0] <= mass_converge_tol && dev[1] <= step_size_converge_tol;
      });
      warm_done = static_cast<int>(warm_max);
      chain_counts(false, warm_rows);
    }
    while (!free_warm && warm_done < max_warmup_iter) {
      const int n = block(warm_done, min_warmup_iter, max_warmup_iter, save_warmup);
      throttle();
      check(wb200_session_warmup(s, n, save_warmup ? 1 : 0, &e), e);
      launched();
      if (save_warmup) {
        flush_pending();
        mark_block(warm_done + n);
      }
      progress(warm_done, warm_done + n, true);
      report_exceptions();
      warm_done += n;
      interrupt.throw_if_interrupted();  // adapt.hpp:227
      if (warm_done >= min_warmup_iter && warm_done < max_warmup_iter) {
        double dev[2];
        check(wb200_session_warmup_sums(s, sums.ptr, &e), e);
        check(wb200_session_warmup_deviation(s, sums.ptr, dev, &e), e);
        if (dev[0] <= mass_converge_tol && dev[1] <= step_size_converge_tol) break;
      }
    }
    check(wb200_session_freeze(s, &e), e);
    if (so) check(wb200_session_stream_begin(s, so->max_lags, &e), e);
    phase("warmup");
    const int saved_warm = save_warmup ? warm_done : 0;
    // ---- sampling: R-hat of lp between blocks (sampler.hpp:132-151)
    int samp_done = 0;
    if (!free_warm) warm_max = warm_done;
    const bool free_samp = allow_free && (min_sampling_iter < max_sampling_iter || free_warm);
    if (free_samp) {
      const long long samp_max =
          free_phase(true, min_sampling_iter, max_sampling_iter, true, [&] {
            double m0[4], m[4];  // util.hpp:401-404, as in the uniform loop below
            check(wb200_session_lp_moments(s, m0, &e), e);
            check(wb200_session_lp_moments_centered(s, m0[0] / m0[3], m, &e), e);
            const double M = m[3];
            const double var_of_means = (m[1] - m[0] * m[0] / M) / (M - 1.0);
            const double mean_of_vars = m[2] / M;
            const double r_hat = std::sqrt(1 + var_of_means / mean_of_vars);
            if (refresh != 0 && print) {  // handlers.hpp:164-172
              std::stringstream ss;
              ss.precision(10);
              ss << "Controller: R-hat at " << r_hat << std::endl;
              say(ss.str());
            }
            return r_hat <= rhat_converge_tol;
          });
      samp_done = static_cast<int>(samp_max);
    }
    while (!free_samp && samp_done < max_sampling_iter) {
      const int n = block(samp_done, min_sampling_iter, max_sampling_iter, !so);
      throttle();
      check(wb200_session_sample(s, n, 1, &e), e);
      launched();
      if (!so) {
        flush_pending();
        mark_block(saved_warm + samp_done + n);
      }
      progress(warm_done + samp_done, warm_done + samp_done + n, false);
      report_exceptions();
      samp_done += n;
      interrupt.throw_if_interrupted();  // sampler.hpp:154
      if (samp_done >= min_sampling_iter && samp_done < max_sampling_iter) {
        // util.hpp:401-404 is a two-pass variance: first the mean of the chain means,
        // then the squared deviations about it
        double m0[4], m[4];
        check(wb200_session_lp_moments(s, m0, &e), e);
        check(wb200_session_lp_moments_centered(s, m0[0] / m0[3], m, &e), e);
        const double M = m[3];
        const double var_of_means = (m[1] - m[0] * m[0] / M) / (M - 1.0);
        const double mean_of_vars = m[2] / M;
        const double r_hat = std::sqrt(1 + var_of_means / mean_of_vars);
        if (refresh != 0 && print) {  // handlers.hpp:164-172
          std::stringstream ss;
          ss.precision(10);
          ss << "Controller: R-hat at " << r_hat << std::endl;
          say(ss.str());
        }
        if (r_hat <= rhat_converge_tol) break;
      }
    }
    std::vector<long long> samp_rows;  // per chain, when sampling ran free
    if (free_samp) {
      chain_counts(true, samp_rows);
      if (!so) {  // every row any chain holds, in one strided copy
        long long rows_max = 0;
        for (size_t c = 0; c < C; ++c) {
          rows_max = std::max(rows_max, samp_rows[c] + (save_warmup && free_warm
                                                            ? warm_rows[c] : saved_warm));
        }
        pending_from = std::min<long long>(pending_from, rows_max);
        mark_block(rows_max);
      }
    }
    if (!so) flush_pending();
    check(wb200_session_sync(s, &e), e);
    phase("sampling");
    if (so) {
      check(wb200_session_stream_summary(s, C > 1 ? so->rhat : nullptr, so->ess, so->mcse,
                                         so->mean, so->var, so->truncated, &e), e);
      phase("summaries");
    }
    WB200_CUDA(cudaStreamSynchronize(copy_stream));
    phase("copy tail");
    // ---- outputs (walnutpy.cpp:196-221; handlers.hpp:73-100)
    for (size_t c = 0; c < C; ++c) {
      final_lengths[c] = !save_warmup ? 0
                         : free_warm  ? static_cast<int>(warm_rows[c])
                                      : saved_warm;
      final_lengths[c + C] = free_samp ? static_cast<int>(samp_rows[c]) : samp_done;
    }
    check(wb200_session_get_state(s, nullptr, inv_metric_out, stepsize_out, nullptr,
                                  nullptr, &e), e);
    check(wb200_session_counters(s, &g_last_run.grad_evals, &g_last_run.macro_steps,
                                 &g_last_run.launches, &e), e);
    g_last_run.warmup_iters = warm_done;
    g_last_run.sampling_iters = samp_done;
  });
  const auto t_destroy = std::chrono::steady_clock::now();
  wb200_session_destroy(s);
  if (std::getenv("WB200_TRACE_PHASES")) {
    std::fprintf(stderr, "[wb200] %-12s %8.2f ms\n", "destroy",
                 std::chrono::duration<double, std::milli>(
                     std::chrono::steady_clock::now() - t_destroy).count());
  }
  return rc;
}

int walnutpie_sample_device(
    const WalnutModelDesc* model, int num_params, const double* inits,
    size_t num_chains, unsigned int seed, unsigned int id, double init_radius,
    const double* init_inv_metric, int min_warmup_iter, int max_warmup_iter,
    int min_sampling_iter, int max_sampling_iter, int max_trajectory_doublings,
    int max_step_halvings, int min_micro_steps, double max_hamiltonian_error,
    double step_size_converge_tol, double mass_converge_tol,
    double rhat_converge_tol, double mass_init_count,
    double mass_additive_smoothing, double max_macro_steps_target,
    double step_size_init, double step_accept_rate_target,
    double step_learning_rate, double step_gradient_decay,
    double step_sq_gradient_decay, double step_stabilization,
    double step_learn_rate_decay, bool save_warmup, double* out,
    size_t out_size, int* final_lengths, double* stepsize_out,
    double* inv_metric_out, int refresh, PRINT_CALLBACK print,
    WalnutpyError** err) {
  return sample_device_impl(
      nullptr, model, num_params, inits, num_chains, seed, id, init_radius, init_inv_metric,
      min_warmup_iter, max_warmup_iter, min_sampling_iter, max_sampling_iter,
      max_trajectory_doublings, max_step_halvings, min_micro_steps, max_hamiltonian_error,
      step_size_converge_tol, mass_converge_tol, rhat_converge_tol, mass_init_count,
      mass_additive_smoothing, max_macro_steps_target, step_size_init,
      step_accept_rate_target, step_learning_rate, step_gradient_decay,
      step_sq_gradient_decay, step_stabilization, step_learn_rate_decay, save_warmup, out,
      out_size, final_lengths, stepsize_out, inv_metric_out, refresh, print, err);
}

// The same call for runs whose draws are too many to keep or ship (SURVEY.md section
// 8(f)-1: 65 536 chains x 1000 draws x 512 parameters are 268 GB): every argument of
// walnutpie_sample_device up to save_warmup, then instead of the draw buffer the posterior
// summaries of summary.hpp:371-405,594-769 computed by the streaming accumulators
// (stream.cu).  Host arrays of num_params, any may be NULL.
int walnutpie_sample_device_summary(
    const WalnutModelDesc* model, int num_params, const double* inits,
    size_t num_chains, unsigned int seed, unsigned int id, double init_radius,
    const double* init_inv_metric, int min_warmup_iter, int max_warmup_iter,
    int min_sampling_iter, int max_sampling_iter, int max_trajectory_doublings,
    int max_step_halvings, int min_micro_steps, double max_hamiltonian_error,
    double step_size_converge_tol, double mass_converge_tol,
    double rhat_converge_tol, double mass_init_count,
    double mass_additive_smoothing, double max_macro_steps_target,
    double step_size_init, double step_accept_rate_target,
    double step_learning_rate, double step_gradient_decay,
    double step_sq_gradient_decay, double step_stabilization,
    double step_learn_rate_decay, int max_lags, double* mean_out, double* var_out,
    double* rhat_out, double* ess_out, double* mcse_out, int* truncated_out,
    int* final_lengths, double* stepsize_out, double* inv_metric_out, int refresh,
    PRINT_CALLBACK print, WalnutpyError** err) {
  const StreamOutputs so{max_lags, mean_out, var_out, rhat_out, ess_out, mcse_out,
                         truncated_out};
  return sample_device_impl(
      &so, model, num_params, inits, num_chains, seed, id, init_radius, init_inv_metric,
      min_warmup_iter, max_warmup_iter, min_sampling_iter, max_sampling_iter,
      max_trajectory_doublings, max_step_halvings, min_micro_steps, max_hamiltonian_error,
      step_size_converge_tol, mass_converge_tol, rhat_converge_tol, mass_init_count,
      mass_additive_smoothing, max_macro_steps_target, step_size_init,
      step_accept_rate_target, step_learning_rate, step_gradient_decay,
      step_sq_gradient_decay, step_stabilization, step_learn_rate_decay, false, nullptr, 0,
      final_lengths, stepsize_out, inv_metric_out, refresh, print, err);
}

// Page-locked host memory for `out` / `inits`: makes the read-back of
// walnutpie_sample_device a direct DMA that overlaps sampling.
int wb200_host_alloc(size_t bytes, void** ptr, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    require_gpu();
    if (!ptr) throw std::invalid_argument("ptr is null");
    *ptr = nullptr;
    WB200_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  });
}
void wb200_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}

// Return the cached device memory of destroyed sessions to the driver.
int wb200_trim_memory(int device, WalnutpyError** err) {
  return catch_exceptions(err, [&] {
    require_gpu();
    cudaMemPool_t pool;
    WB200_CUDA(cudaDeviceSynchronize());
    WB200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    WB200_CUDA(cudaMemPoolTrimTo(pool, 0));
  });
}

}  // extern "C"
