/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
 *
 * C entry points shared by the two CPU checkers:
 *   - prefix `oracle_` : the Eigen-free restatement (oracle/walnuts_oracle.hpp)
 *                        -> oracle/liboracle.so
 *   - prefix `ref_`    : the UNMODIFIED reference headers compiled against the
 *                        local Eigen API shim -> oracle/_ref/libwalnuts_ref.so
 * Both export the same signatures so tests can diff them call by call.
 */
#ifndef WALNUTS_ORACLE_CAPI_H
#define WALNUTS_ORACLE_CAPI_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kind: 0 std normal, 1 diagonal Gaussian (data0 = precision[D]),
 * 2 Neal's funnel, 3 logistic regression (data0 = X[N][D] row-major fp64,
 * data1 = y[N]), 4 C callback (data0 = LOGP_CFUNC, data1 = user data) */
typedef struct {
  int kind;
  size_t D;
  size_t N;
  const void* data0;
  const void* data1;
} OracleTarget;

/* WarmupConfig (config.hpp:626-640) + SamplingConfig (:947-953) */
typedef struct {
  int min_warmup_iter, max_warmup_iter;
  int min_sampling_iter, max_sampling_iter;
  int max_trajectory_doublings, max_step_halvings, min_micro_steps;
  int publish_stride;
  double max_hamiltonian_error;
  double step_size_converge_tol, mass_converge_tol, rhat_converge_tol;
  double mass_init_count, mass_additive_smoothing, max_macro_steps_target;
  double step_accept_rate_target, step_learning_rate, step_gradient_decay;
  double step_sq_gradient_decay, step_stabilization, step_learn_rate_decay;
} OracleConfig;

/* rng_policy: 0 = std::mt19937_64(seed_seq{seed, chain+1}) through libstdc++
 * distributions (the reference's stream, api.hpp:46-51);
 * 1 = stateless Philox4x32-10 keyed (seed, chain) (oracle only). */

#define ORACLE_DECLARE(P)                                                      \
  /* one chain: n_warmup AdaptiveWalnuts iterations, then sampler(), then     \
   * n_sampling WalnutsSampler iterations.  All outputs nullable. */           \
  int P##_run_chain(const OracleTarget* target, const OracleConfig* cfg,       \
                    uint32_t seed, uint32_t chain, int rng_policy,             \
                    const double* theta0, const double* mass0, double step0,   \
                    int n_warmup, int n_sampling, double* warmup_draws,        \
                    double* warmup_lp, double* warmup_step,                    \
                    double* warmup_inv_mass, int* warmup_depth, double* draws, \
                    double* lp, int* depth, double* inv_mass_out,              \
                    double* step_out, int* min_micro_out,                      \
                    uint64_t* grad_evals);                                     \
  /* fixed-parameter sampler only (WalnutsSampler, walnuts.hpp:605-766) */     \
  int P##_run_sampler(const OracleTarget* target, uint32_t seed,               \
                      uint32_t chain, int rng_policy, uint32_t first_iter,     \
                      const double* theta0, const double* inv_mass,            \
                      double step, int max_depth, int max_halvings,            \
                      int min_micro, double max_error, int n_iter,             \
                      double* draws, double* lp, int* depth,                   \
                      uint64_t* grad_evals);                                   \
  /* positions: N(0, radius^2) from mt19937_64(seed_seq{seed,1})              \
   * (walnutpy.cpp:186-190, config.hpp:259-268) */                            \
  int P##_init_positions(size_t num_chains, size_t D, uint32_t seed,           \
                         double radius, double* positions);                    \
  /* run_sampler()'s initialisation (walnutpy.cpp:64-80): masses (given, or   \
   * from |grad|), then adapt_step_build with mt19937_64(seed_seq{seed,2}) */  \
  int P##_init_mass_step(const OracleTarget* target, size_t num_chains,        \
                         uint32_t seed, const double* positions,               \
                         const double* mass_in, double smoothing,              \
                         double step_init, double* mass_out, double* step_out);\
  double P##_leapfrog_error(const OracleTarget* target, const double* theta,   \
                            const double* rho, const double* inv_mass,         \
                            double step);                                      \
  double P##_log_sum_exp(double a, double b);                                  \
  double P##_logp_momentum(const double* rho, const double* inv_mass,          \
                           size_t D);                                          \
  /* the whole multi-chain run (api.hpp:33-69) with chains on threads;        \
   * out is [C][max_sampling + save_warmup*max_warmup][D] as in walnutpy.cpp */\
  int P##_walnuts(const OracleTarget* target, const OracleConfig* cfg,         \
                  size_t num_chains, uint32_t seed, const double* positions,   \
                  const double* mass, const double* steps, int save_warmup,    \
                  double* out, int* final_lengths, double* stepsize_out,       \
                  double* inv_metric_out, uint64_t* grad_evals,                \
                  double* seconds_warmup, double* seconds_sampling);

ORACLE_DECLARE(oracle)
ORACLE_DECLARE(ref)

/* oracle only ------------------------------------------------------------ */
/* L leapfrog micro-steps (walnuts.hpp:329-332) from (theta, rho); outputs the
 * end state, its gradient, logp and joint log density */
int oracle_orbit(const OracleTarget* target, const double* theta,
                 const double* rho, const double* inv_mass, double step,
                 int num_steps, double* theta_out, double* rho_out,
                 double* grad_out, double* logp_out, double* joint_out);
/* The device's initialisation streams (include/walnuts_b200.h, wb200_session_init):
 * positions = radius * N(0,1) from Philox (seed, chain_offset + c, iter 0, kind 2);
 * masses as oracle_init_mass_step; the momentum of adapt_step (util.hpp:285-303) of
 * chain c from Philox (seed, chain_offset + c, iter 0, kind 3) -- per-chain
 * streams instead of the reference's one sequential engine (config.hpp:470-474). */
int oracle_init_positions_philox(size_t num_chains, size_t D, uint32_t seed,
                                 uint32_t chain_offset, double radius,
                                 double* positions);
int oracle_init_mass_step_philox(const OracleTarget* target, size_t num_chains,
                                 uint32_t seed, uint32_t chain_offset,
                                 const double* positions, const double* mass_in,
                                 double smoothing, double step_init,
                                 double* mass_out, double* step_out);
/* One evaluation of the warm-up controller's convergence statistics (adapt.hpp:186-224)
 * from per-chain snapshots {log_step[C], log_mass[C][D]}: max_m ||(M_m - gm)/gm||_2 and
 * max(0, max_m (eps_m - gs)/gs); and of the sampling controller's R-hat of lp
 * (sampler.hpp:132-151, util.hpp:401-404) from per-chain {mean, unbiased variance}. */
int oracle_warmup_controller(size_t num_chains, size_t D, const double* log_step,
                             const double* log_mass, double* max_rel_mass,
                             double* max_rel_step);
int oracle_sampling_rhat(size_t num_chains, const double* mean, const double* var,
                         double* r_hat);
/* Arithmetic policy of the accumulate sites (oracle/targets.hpp): 0 = the reference's
 * baseline build, separate roundings (default; the policy pinned to the reference);
 * 1 = fused multiply-add, what the device kernels ship.  Returns the previous value. */
int oracle_set_fused_arith(int fused);
/* The fixed-step orbit in the device's fp32 mode (include/walnuts_b200.h,
 * WalnutTuning::precision = 1): state and element-wise arithmetic in float, energies
 * in double; kinds 0-2.  Inputs and outputs are fp64 arrays (rounded to float inside). */
int oracle_orbit_f32(const OracleTarget* target, const double* theta,
                     const double* rho, const double* inv_mass, double step,
                     int num_steps, double* theta_out, double* rho_out,
                     double* grad_out, double* logp_out, double* joint_out);
int oracle_logp_grad(const OracleTarget* target, const double* theta,
                     double* logp, double* grad);
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                   uint32_t k0, uint32_t k1, uint32_t* out4);
void oracle_philox_normals(uint32_t seed, uint32_t chain, uint32_t iter,
                           uint32_t kind, size_t n, double* out);
double oracle_philox_uniform(uint32_t seed, uint32_t chain, uint32_t iter,
                             uint32_t index);
/* summaries on ROW-MAJOR [num_draws][D] ragged chains (summary.hpp) */
int oracle_ess(const double* draws, int num_draws, int num_params,
               const int* lengths, int num_chains, double* out);
int oracle_r_hat(const double* draws, int num_draws, int num_params,
                 const int* lengths, int num_chains, double* out);
int oracle_mcse(const double* draws, int num_draws, int num_params,
                const int* lengths, int num_chains, double* out);
int oracle_autocovariance(const double* draws, int num_draws, int num_params,
                          const int* lengths, int num_chains, double* out);
const char* oracle_last_error(void);
const char* ref_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
