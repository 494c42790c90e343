/* walnuts_b200 — C-ABI of the B200-native WALNUTS sampler (libwalnuts_b200.so).
 *
 * Drop-in boundary for the hot path of flatironinstitute/walnuts: the same
 * `extern "C"` surface as the reference's libwalnutpy
 * (python/src/walnutpie/walnutpy.cpp), plus the device-model entry point and a
 * session API that keeps thousands of chains resident in HBM.  Plain pointers
 * and sizes only.  File:line citations are relative to /root/reference.
 *
 * Error convention (python/src/walnutpie/errors.hpp:10-72): return 0 on
 * success, -1 on failure with *err set; free with walnutpie_destroy_error.
 */
#ifndef WALNUTS_B200_H
#define WALNUTS_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error objects: errors.hpp:10-24 ------------------------------------ */
typedef enum { wb200_generic = 0, wb200_config = 1, wb200_interrupt = 2 } WalnutpyErrorType;
typedef struct WalnutpyError WalnutpyError;

/* walnutpy.cpp:371-389 */
const char* walnutpie_get_error_message(const WalnutpyError* err);
int walnutpie_get_error_type(const WalnutpyError* err);
void walnutpie_destroy_error(WalnutpyError* err);
/* walnutpy.cpp:224-225 */
char walnutpie_separator_char(void);

/* ---- callbacks: walnutpy.cpp:127-132, handlers.hpp:15 -------------------- */
typedef int (*LOGP_CFUNC)(size_t theta_size, const double* theta, double* grad,
                          double* lp, void* data);
typedef void (*PRINT_CALLBACK)(const char* msg, size_t len, bool bad);

/* ---- device model descriptor (replaces `LOGP_CFUNC logp_c, void* data`) ---
 * kind 0  standard normal                  examples/walnutpie_api.cpp:39-43
 * kind 1  diagonal Gaussian, data0 = precision[D] (1/sigma_d^2), host fp64
 * kind 2  Neal's funnel (x0 ~ N(0,9), x_i | x0 ~ N(0, exp(x0)))
 * kind 3  Bayesian logistic regression, data0 = X[N][D] row-major host fp64,
 *         data1 = y[N] host fp64 in {0,1}, prior N(0, I)
 * kind 4  the caller's own density, batched on the device: data0 = a
 *         WB200_BATCH_LOGP_GRAD function, data1 = its `data` argument.  This is
 *         LOGP_CFUNC (walnutpy.cpp:127-130) for a whole batch of chains: once per
 *         lock-step tick the sampler posts the positions of all chains and the
 *         function ENQUEUES, on the CUDA stream it is given, work that fills
 *         grad and lp.  Everything is device memory: theta, grad [num_chains][ld]
 *         fp64 row-major (columns >= num_params are padding), lp [num_chains].
 *         Return 0.  A non-zero return is what a throwing density is to the reference
 *         (walnutpy.cpp:162-170): during initialisation it ends the run with the
 *         error "logp failed with code N"; inside a transition every chain of that
 *         tick continues with logp = -inf and a zero gradient (NoExceptLogpGrad,
 *         util.hpp:336-346), the failure is counted (wb200_session_logp_exceptions)
 *         and reported through the print callback.
 * kind 5  the caller's own density as CUDA source text, compiled at run time (NVRTC) into
 *         the chain-resident kernel: data0 = const char* source (zero-terminated), data1 =
 *         N parameter doubles (host; N may be 0), fp64.  What LogpGrad (concepts.hpp:25-60)
 *         is to the reference's host threads: the density becomes the `Target` of the
 *         transition kernel, so the chain stays in registers across an orbit like the
 *         built-in targets' (kind 4 streams the state through HBM once per gradient).
 *         Element-wise (separable) densities define
 *             __device__ void wb200_logp_grad(int d, double x, const double* par,
 *                                             double& lp, double& g);
 *         -- term d of the log density and its derivative at x.  Densities that need the
 *         chain's all-reduce (like kind 2) define a Target template and name it:
 *             template <int T, int K, class Real> struct MyTarget { init(...); grad(...); };
 *             #define WB200_USER_TARGET MyTarget
 *         (interface: walnuts_b200/csrc/chain_kernel.cuh, "Targets").  A source that does
 *         not compile is a config error carrying the compiler's log.  Compiled kernels are
 *         cached per (device, launch shape, source).
 */
typedef int (*WB200_BATCH_LOGP_GRAD)(size_t num_chains, size_t num_params, size_t ld,
                                     const double* theta, double* grad, double* lp,
                                     void* cuda_stream, void* data);
typedef struct {
  int kind;
  int D;
  size_t N;
  const void* data0;
  const void* data1;
  int precision;  /* 0 = fp64, 1 = fp32 mode (see WalnutTuning::precision); kinds 0-2 */
} WalnutModelDesc;

/* The 26 tuning arguments of walnutpie_sample_cfunc (walnutpy.cpp:134-149),
 * same names, same meaning, same defaults as pyfunc.py:51-82. */
typedef struct {
  int min_warmup_iter, max_warmup_iter;
  int min_sampling_iter, max_sampling_iter;
  int max_trajectory_doublings, max_step_halvings, min_micro_steps;
  double max_hamiltonian_error;
  double step_size_converge_tol, mass_converge_tol, rhat_converge_tol;
  double mass_init_count, mass_additive_smoothing, max_macro_steps_target;
  double step_size_init;
  double step_accept_rate_target, step_learning_rate, step_gradient_decay;
  double step_sq_gradient_decay, step_stabilization, step_learn_rate_decay;
  int publish_stride;   /* WarmupConfig::publish_stride, config.hpp:639 (0 -> 5) */
  /* 0 = fp64 (the reference's arithmetic).  1 = fp32 mode (element-wise targets): the
   * integrator state theta / rho / grad / M^-1 and the leapfrog, gradient, kinetic and
   * U-turn products are single precision inside a transition; energies and U-turn dots
   * are accumulated across threads in fp64, every scalar decision (energy error, merge
   * weights, Adam) and the adaptation statistics are fp64, stored draws and all
   * host-visible state stay fp64.  Tolerance: fixed-step orbits within
   * 2e-6 * sqrt(steps) relative of the fp64 integrator (tests/test_gpu_fp32.py). */
  int precision;
} WalnutTuning;

/* Reference defaults (config.hpp:626-640, :947-953; step_size_init as pyfunc.py:74) */
void walnuts_b200_default_tuning(WalnutTuning* t);

/* ---- one-shot entry point: replaces walnutpie_sample_cfunc (walnutpy.cpp:134)
 * for a device model.  Everything after (logp_c, data) keeps the reference's
 * order and meaning; host buffers in and out:
 *   inits            nullable [C][D]; else N(0, init_radius^2) (walnutpy.cpp:186-190)
 *   init_inv_metric  nullable [C][D]; used as the initial MASS, bug-compatible
 *                    with walnutpy.cpp:64-70
 *   out              [C][max_sampling_iter + save_warmup*max_warmup_iter][D]; a chain's
 *                    block holds its saved warm-up draws followed directly by its sampling
 *                    draws (walnutpy.cpp:196-203)
 *   final_lengths    [2C] warm-up lengths then sampling lengths, PER CHAIN: with
 *                    min_iter < max_iter the chains run free like the reference's threads
 *                    (adapt.hpp:110-129, sampler.hpp:79-94) -- by equal work, a budget of
 *                    gradient evaluations per block -- and stop on their own; the
 *                    controllers keep the reference's rules (no decision before every
 *                    chain has min_iter, stop on convergence or when all are at max_iter).
 *                    Environment WB200_BLOCKS=uniform: blocks of publish_stride
 *                    iterations, all chains stop together.
 */
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
    WalnutpyError** err);

/* The one-shot call for runs whose draws are too many to keep or ship (65 536 chains x
 * 1000 draws x 512 parameters are 268 GB): the arguments of walnutpie_sample_device up to
 * step_learn_rate_decay, then, instead of save_warmup and the draw buffer, the number of
 * autocovariance lags to keep (8, 16 or 32) and HOST arrays of num_params (any may be
 * NULL) that receive the posterior summaries of summary.hpp:371-405,594-769 -- pooled
 * mean and variance, R-hat, ESS, MCSE and a flag per dimension whose Geyer sequence
 * was cut at max_lags -- computed on the device by the streaming accumulators (see
 * "streaming summaries" below).  Nothing but these 5 x num_params doubles, the lengths,
 * step sizes and metrics crosses PCIe. */
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
    PRINT_CALLBACK print, WalnutpyError** err);

/* The same call over several GPUs of this process -- walnutpie::walnuts (api.hpp:33-69) is
 * ONE call that runs every chain and both controllers, and so is this.  Chains are sharded
 * over `devices` in contiguous blocks of global chain ids (which select the Philox streams:
 * the chains of a G-GPU run are exactly those of a 1-GPU run), one host thread drives each
 * device, and the only exchange is the cross-chain summaries, all-reduced with NCCL over
 * NVLink inside the library (ncclCommInitAll; libnccl.so.2 is loaded at run time): the
 * warm-up controller's sums and maxima (adapt.hpp:186-224), the sampling controller's lp
 * moments (sampler.hpp:132-151) and the two phases of the streaming summaries.  inits,
 * init_inv_metric, final_lengths, stepsize_out, inv_metric_out cover ALL chains, in global
 * chain order.  With one device no NCCL is needed. */
int walnutpie_sample_device_multi(
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
    PRINT_CALLBACK print, WalnutpyError** err);

/* walnutpie_sample_cfunc (walnutpy.cpp:134): exported for link compatibility; a
 * host callback cannot feed a device batch, so it fails with a `generic` error that
 * names walnutpie_sample_device and the batched device callback (WalnutModelDesc
 * kind 4), its device-side counterpart. */
int walnutpie_sample_cfunc(
    LOGP_CFUNC logp_c, void* data, int num_params, const double* inits,
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
    WalnutpyError** err);

/* walnutpy.cpp:227-243 (STREAM_CALLBACK: thirdparty/bridgestan/bridgestan.h:400).
 * Always fails with a `generic` error: BridgeStan models stay CPU-reference-only. */
typedef void (*STREAM_CALLBACK)(const char* data, size_t size);
int walnutpie_sample_bridgestan(
    const char* bs_dll, const char* json_data, STREAM_CALLBACK callback,
    unsigned int model_seed, const char* inits, size_t num_chains,
    unsigned int seed, unsigned int id, double init_radius,
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
    WalnutpyError** err);

/* ---- summaries: walnutpy.cpp:333-369.  `draws` is HOST, ROW-MAJOR
 * [num_draws][num_params] with chains stacked (what summary.py:30 passes);
 * computed on the device. */
int walnutpie_ess(const double* draws, int num_draws, int num_params,
                  const int* lengths, int num_chains, double* out,
                  WalnutpyError** err);
int walnutpie_r_hat(const double* draws, int num_draws, int num_params,
                    const int* lengths, int num_chains, double* out,
                    WalnutpyError** err);
int walnutpie_mcse(const double* draws, int num_draws, int num_params,
                   const int* lengths, int num_chains, double* out,
                   WalnutpyError** err);

/* ---- session API: chains stay resident in HBM -----------------------------
 * The pieces walnutpie::walnuts<RNG>() (api.hpp:33-69) is made of, batched:
 * create = per-chain RNG streams + AdaptiveWalnuts state (api.hpp:46-58),
 * warmup = AdaptiveWalnuts::operator() x n (adaptive_walnuts.hpp:234-251),
 * freeze = AdaptiveWalnuts::sampler() (adaptive_walnuts.hpp:263-271),
 * sample = WalnutsSampler::operator() x n (walnuts.hpp:682-692).
 * Chain `c` of a session uses the Philox stream of global chain id
 * chain_offset + c, so results do not depend on how chains are sharded. */
typedef struct wb200_session wb200_session;

int wb200_session_create(const WalnutModelDesc* model, size_t num_chains,
                         unsigned int seed, unsigned int chain_offset,
                         const WalnutTuning* tuning, int device,
                         wb200_session** out, WalnutpyError** err);
void wb200_session_destroy(wb200_session* s);

/* positions/mass/steps are HOST [C][D] / [C][D] / [C]; any may be NULL:
 * positions -> N(0, init_radius^2) per chain (Philox kind 2);
 * mass      -> (1-s)|grad logp(theta0)| + s, s = mass_additive_smoothing
 *              (config.hpp:360-370);
 * steps     -> step_size_init refined by the doubling / sqrt(1/2) search of
 *              util.hpp:285-303 on the device (Philox kind 3). */
int wb200_session_init(wb200_session* s, const double* positions,
                       double init_radius, const double* mass,
                       const double* steps, WalnutpyError** err);

/* Reserve device storage for `capacity` draws per chain; trace != 0 also
 * records lp, depth, step size and (warm-up) the metric used per iteration. */
int wb200_session_reserve_draws(wb200_session* s, long long capacity, int trace,
                                WalnutpyError** err);

/* n warm-up iterations for every chain; store != 0 appends the draws */
int wb200_session_warmup(wb200_session* s, int n_iter, int store,
                         WalnutpyError** err);
/* freeze step / inverse mass / min_micro (adaptive_walnuts.hpp:263-271) */
int wb200_session_freeze(wb200_session* s, WalnutpyError** err);
int wb200_session_sample(wb200_session* s, int n_iter, int store,
                         WalnutpyError** err);
/* Free-running sampling: the reference's chains are threads that run at their own pace
 * until the controller stops them (sampler.hpp:79-94), so draw counts are ragged.  The
 * device analogue of equal time is equal work -- n_ticks gradient evaluations per chain:
 *   lock-step sessions     exactly n_ticks ticks (one batched gradient each); transitions
 *                          in flight carry over to the next call;
 *   chain-resident sessions  every chain completes the transitions that fit into n_ticks
 *                          evaluations (the one that exhausts the budget is finished and
 *                          its excess comes off the chain's next budget).
 * No launch waits for the chain with the longest orbits.  wb200_session_chain_rows returns
 * the rows stored per chain; wb200_session_summary summarises rows [first, rows_c) of
 * every chain.  Chain c's rows are the first rows_c draws of the very chain a fixed-length
 * run produces (the Philox streams are keyed by chain and iteration). */
int wb200_session_sample_ticks(wb200_session* s, int n_ticks, int store,
                               WalnutpyError** err);
/* The same for the adaptive phase (AdaptiveWalnuts::operator(), adaptive_walnuts.hpp:234):
 * every chain adapts on its own transition count, like the reference's per-thread
 * chains.  wb200_session_freeze abandons the transitions still in flight. */
int wb200_session_warmup_ticks(wb200_session* s, int n_ticks, int store,
                               WalnutpyError** err);
/* The free-running launch with the phase's iteration limit (lock-step sessions: eval_budget
 * ticks) --
 * sampling = 0 warm-up / 1 sampling; no chain goes beyond iter_cap iterations of the phase
 * in total (<= 0: no limit), as each reference chain stops at max_iter (adapt.hpp:116,
 * sampler.hpp:82).  wb200_session_iter_stats returns stats4 = {min, max, sum} over the
 * chains of the phase's per-chain iteration counts -- what the reference's controllers
 * read from the chains' snapshots (adapt.hpp:196-203,219; sampler.hpp:134-141,149) -- and
 * the gradient evaluations of all chains since initialisation. */
int wb200_session_run_evals(wb200_session* s, int sampling, long long eval_budget,
                            long long iter_cap, int store, WalnutpyError** err);
int wb200_session_iter_stats(wb200_session* s, int sampling, long long* stats4,
                             WalnutpyError** err);
int wb200_session_chain_rows(wb200_session* s, long long* rows, WalnutpyError** err);
int wb200_session_summary(wb200_session* s, long long first, double* rhat, double* ess,
                          double* mcse, double* mean, double* var, WalnutpyError** err);
/* moments [3*D + 1] = {sum_c mean_c[D], sum_c mean_c^2[D], sum_c var_c[D], chains}: the
 * payload a multi-GPU caller all-reduces (NCCL, SUM) to get the R-hat over all ranks'
 * chains (summary.hpp:594-619) */
int wb200_session_rhat_moments(wb200_session* s, long long first, double* moments,
                               WalnutpyError** err);
/* block until the session's stream is idle */
int wb200_session_sync(wb200_session* s, WalnutpyError** err);

/* Cross-chain convergence statistics, evaluated on the device.
 * warm-up (adapt.hpp:186-224): out[0] = max_m ||(M_m - gm)/gm||_2,
 *   out[1] = max_m (eps_m - gs)/gs;  sums_io: if non-NULL, DEVICE buffer of
 *   (D + 2) doubles {sum log M[D], sum log eps, count} that the caller may
 *   all-reduce (NCCL) between the two phases: phase 0 fills it from the local
 *   chains, phase 1 consumes the (reduced) sums and writes out[0..1].
 * sampling (sampler.hpp:132-151): moments = {sum mu, sum mu^2, sum var, count}. */
int wb200_session_warmup_sums(wb200_session* s, double* sums_device,
                              WalnutpyError** err);
int wb200_session_warmup_deviation(wb200_session* s, const double* sums_device,
                                   double* out_host2, WalnutpyError** err);
int wb200_session_lp_moments(wb200_session* s, double* moments_host4,
                             WalnutpyError** err);

/* sampling moments about `center`: {sum (mu - c), sum (mu - c)^2, sum var, count}.
 * util.hpp:401-404 is a two-pass variance; a multi-GPU caller all-reduces the plain sums
 * first and passes the global mean of the chain means as the centre, which keeps the
 * between-chain variance free of cancellation when |lp| is large. */
int wb200_session_lp_moments_centered(wb200_session* s, double center,
                                      double* moments_host4, WalnutpyError** err);
/* kind 4: batched density evaluations that returned non-zero inside a transition.  As
 * NoExceptLogpGrad does for one chain (util.hpp:336-346), every chain of that tick went
 * on with logp = -inf and a zero gradient.  (A failure during initialisation ends the
 * run with "logp failed with code N", walnutpy.cpp:162-170.) */
int wb200_session_logp_exceptions(wb200_session* s, unsigned long long* count);

/* ---- streaming summaries (summary.hpp:594-769 without keeping the draws) ---------
 * After wb200_session_stream_begin the draw buffer reserved with
 * wb200_session_reserve_draws is a STAGING block: every storing wb200_session_sample /
 * _sample_ticks call refills it from row 0 and folds its rows into per-chain running
 * sums -- count, sum, lag products sum_i y_i y_{i+t} for t < max_lags (8, 16 or 32) and
 * the first / last max_lags values, all about the chain's first streamed draw.  From
 * those the reference's R-hat, ESS, MCSE, pooled mean and variance follow exactly,
 * provided the Geyer sequence of a dimension ends before lag max_lags (else that
 * dimension is flagged in `truncated` and its ESS is an upper bound).  Chains with
 * fewer than 3 streamed draws are left out (summary.hpp:595-603).
 *
 * Multi-GPU (chains sharded over ranks): the cross-chain part is two reductions a
 * caller performs with NCCL (or any all-reduce) between the calls:
 *   phase1 -> out1 [2D + 3] = {sum_k mu_k [D], sum_k n_k mu_k [D], K, N, min_len}
 *             all-reduce: SUM over [0, 2D + 2), MIN over element 2D + 2;
 *   phase2(reduced1) -> out2 [(3 + max_lags) D] = {sum_k (mu_k - mbar)^2, sum_k s_k^2,
 *             pooled sum of squares, sum_k acov_k(t) for t < max_lags}, centred on the
 *             GLOBAL means of phase 1;  all-reduce: SUM;
 *   wb200_stream_finish(reduced1, reduced2) -> the summaries, identical on every rank
 *             and equal to those of one session holding all the chains.
 * wb200_session_stream_summary runs the three steps for a single session.
 * Outputs are HOST arrays of num_params (any may be NULL). */
int wb200_session_stream_begin(wb200_session* s, int max_lags, WalnutpyError** err);
int wb200_session_stream_phase1(wb200_session* s, double* out1, WalnutpyError** err);
int wb200_session_stream_phase2(wb200_session* s, const double* reduced1, double* out2,
                                WalnutpyError** err);
int wb200_stream_finish(int num_params, int max_lags, const double* reduced1,
                        const double* reduced2, double* rhat, double* ess, double* mcse,
                        double* mean, double* var, int* truncated, WalnutpyError** err);
int wb200_session_stream_summary(wb200_session* s, double* rhat, double* ess, double* mcse,
                                 double* mean, double* var, int* truncated,
                                 WalnutpyError** err);
/* draws folded in so far, per chain (host [C]) */
int wb200_session_stream_counts(wb200_session* s, long long* counts, WalnutpyError** err);

/* Read-back (host buffers).  draws: [C][count][D] from row `first`. */
int wb200_session_get_draws(wb200_session* s, long long first, long long count,
                            double* out, WalnutpyError** err);
int wb200_session_get_trace(wb200_session* s, long long first, long long count,
                            double* lp, int* depth, double* step,
                            double* inv_mass, WalnutpyError** err);
int wb200_session_get_state(wb200_session* s, double* theta, double* inv_mass,
                            double* step, int* min_micro,
                            unsigned long long* grad_evals, WalnutpyError** err);
/* raw device pointers for zero-copy consumers (torch / NCCL): draws are
 * [C][capacity][ld] with ld = row stride in doubles */
int wb200_session_device_draws(wb200_session* s, double** draws, long long* capacity,
                               int* ld, long long* rows_written);
/* total gradient evaluations so far (sum over chains) and launch count */
int wb200_session_counters(wb200_session* s, unsigned long long* grad_evals,
                           unsigned long long* macro_steps,
                           unsigned long long* kernel_launches,
                           WalnutpyError** err);
/* last kernel's device time in ms (CUDA events on the session stream) */
int wb200_session_last_kernel_ms(wb200_session* s, float* ms);
/* user timer on the session stream: record event `which` (0 = start, 1 = stop)
 * where the stream currently is; elapsed waits for the stop event */
int wb200_session_timer_record(wb200_session* s, int which, WalnutpyError** err);
int wb200_session_timer_elapsed_ms(wb200_session* s, float* ms, WalnutpyError** err);
/* Statistics of the most recent walnutpie_sample_device call of this process:
 * gradient evaluations, macro steps, kernels launched, warm-up and sampling
 * iterations actually run. */
int wb200_last_run_stats(unsigned long long* grad_evals, unsigned long long* macro_steps,
                         unsigned long long* kernel_launches, int* warmup_iters,
                         int* sampling_iters);

/* Page-locked host memory for the `out` / `inits` buffers of
 * walnutpie_sample_device (the reference's callers allocate these with numpy,
 * util.py:16-32): with a pinned `out` the draw read-back is a direct DMA that
 * overlaps sampling.  Pageable buffers work too, only slower. */
int wb200_host_alloc(size_t bytes, void** ptr, WalnutpyError** err);
void wb200_host_free(void* ptr);
/* Device memory of destroyed sessions stays cached in the device's memory pool for
 * the next session; this returns it to the driver (e.g. before handing the GPU to
 * another library in the same process). */
int wb200_trim_memory(int device, WalnutpyError** err);

/* Fixed-step leapfrog orbit (walnuts.hpp:329-332) for parity checks:
 * theta/rho/inv_mass HOST [C][D]; advances `num_steps` micro-steps. */
int wb200_orbit(const WalnutModelDesc* model, size_t num_chains,
                const double* theta, const double* rho, const double* inv_mass,
                double step, int num_steps, double* theta_out, double* rho_out,
                double* grad_out, double* logp_out, double* joint_out,
                WalnutpyError** err);

/* device-side Philox (for the generator known-answer tests) */
int wb200_philox(const uint32_t* ctr_key6, size_t n, uint32_t* out4,
                 WalnutpyError** err);
int wb200_philox_normals(unsigned int seed, unsigned int chain, unsigned int iter,
                         unsigned int kind, size_t n, double* out,
                         WalnutpyError** err);

/* device summaries on DEVICE draws [C][capacity][ld] (equal lengths):
 * per-dimension R-hat, ESS, MCSE, pooled mean and variance (summary.hpp) */
int wb200_device_summary(const double* draws_device, size_t num_chains,
                         long long capacity, long long first, long long count,
                         int D, int ld, double* rhat, double* ess, double* mcse,
                         double* mean, double* var, WalnutpyError** err);

/* Batched logistic-regression log density and gradient on the tensor cores
 * (tcgen05): X host fp64 [N][D] row-major (rounded to bf16), y host fp64 [N],
 * theta host fp64 [C][D]; outputs logp [C], grad [C][D].  repeats > 0 also times
 * `repeats` further evaluations with CUDA events (ms per batched evaluation). */
int wb200_logistic_logp_grad(const double* X, const double* y, size_t N, int D,
                             const double* theta, size_t C, double* logp, double* grad,
                             int repeats, float* ms_per_eval, WalnutpyError** err);

/* kind 5: compile `source` for a model of num_params dimensions without touching a GPU
 * (NVRTC only); the compiler's output (warnings) goes to log[log_size] if non-null */
int wb200_compile_device_source(const char* source, int num_params, char* log,
                                size_t log_size, WalnutpyError** err);

const char* walnuts_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WALNUTS_B200_H */
