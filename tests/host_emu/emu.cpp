// TEST INFRASTRUCTURE: the device transition state machine
// (walnuts_b200/csrc/chain_kernel.cuh) compiled for the host with one emulated
// thread per chain (T = 1, K = kEmuK register chunks).  Same source, same
// -ffp-contract=off arithmetic; sums run in element order, i.e. exactly the
// oracle's order, so tests can demand bit equality of whole warm-up + sampling
// runs against the oracle's Philox policy on CPU.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../walnuts_b200/csrc/chain_kernel.cuh"

using namespace wb200;

constexpr int kEmuK = 64;  // up to 128 dimensions

struct EmuTuning {
  int max_depth, max_halvings, min_micro;
  double max_error, mass_init_count, macro_target;
  double adam_target, adam_lr, adam_b1, adam_b2, adam_eps, adam_decay;
};

template <template <int, int> class TargetT>
static void run(const EmuTuning& t, int D, const double* tparam, uint32_t seed,
                uint32_t chain, const double* theta0, const double* mass0, double step0,
                int n_warmup, int n_sampling, double* draws, double* lp, int* depth,
                double* step_trace, double* im_trace, double* inv_mass_out,
                double* step_out, int* min_micro_out, unsigned long long* evals,
                long long budget = 0, int* launches = nullptr) {
  // budget > 0: both phases as free-running launches of `budget` gradient evaluations
  // (ChainParams::eval_budget) until the phase's iteration count is reached
  const int ld = 2 * kEmuK;  // rows padded to the group's element slots
  const int total = n_warmup + n_sampling;
  std::vector<double> theta(ld, 0.0), inv_mass(ld, 0.0), est(4 * ld, 0.0), tp(ld, 0.0);
  std::vector<double> scratch(static_cast<size_t>(scratch_doubles(t.max_depth, ld)), 0.0);
  std::vector<double> d_draws(static_cast<size_t>(total) * ld), d_im(static_cast<size_t>(total) * ld);
  std::memcpy(theta.data(), theta0, D * 8);
  if (tparam) std::memcpy(tp.data(), tparam, D * 8);
  // what init_kernel does for given positions / masses / steps
  for (int i = 0; i < ld; ++i) {
    double m = i < D ? mass0[i] : 1.0;
    est[1 * ld + i] = t.mass_init_count * (1.0 / m);
    est[3 * ld + i] = t.mass_init_count * m;
  }
  ChainScalars sc{};
  sc.adam_x = std::log(step0);
  sc.adam_b1p = 1.0; sc.adam_b2p = 1.0;
  sc.mm_total = 2.0; sc.mm_count = 1.0;
  sc.est_w = t.mass_init_count;
  sc.step = step0;
  sc.min_micro = t.min_micro;
  unsigned int ticket = 0;
  ChainParams p{};
  p.C = 1; p.D = D; p.ld = ld;
  p.max_depth = t.max_depth; p.max_halvings = t.max_halvings; p.min_micro_cfg = t.min_micro;
  p.max_error = t.max_error; p.mass_init_count = t.mass_init_count;
  p.macro_target = t.macro_target;
  p.adam_target = t.adam_target; p.adam_lr = t.adam_lr; p.adam_b1 = t.adam_b1;
  p.adam_b2 = t.adam_b2; p.adam_eps = t.adam_eps; p.adam_decay = t.adam_decay;
  p.seed = seed; p.chain_offset = chain;
  p.theta = theta.data(); p.inv_mass = inv_mass.data(); p.est = est.data(); p.sc = &sc;
  p.draws = d_draws.data(); p.draw_cap = total;
  p.lp_out = lp; p.depth_out = depth; p.step_out = step_trace; p.im_out = d_im.data();
  p.scratch = scratch.data(); p.scratch_stride = scratch.size();
  p.ticket = &ticket; p.tparam = tp.data();
  Group<1> grp{};
  ChainScalars sc_shared{};  // stands in for the per-chain record in shared memory
  DecisionCache decision_cache{};
  AdamQueue adam_queue{};
  RunLimits run_limits{};
  std::vector<double> chain_smem(static_cast<size_t>(chain_smem_doubles(ld)), 0.0);
  using Target = TargetT<1, kEmuK>;
  long long rows = 0;
  int n_launch = 0;
  if (budget > 0) {
    p.eval_budget = budget; p.rows = &rows;
  }
  // warm-up launch
  p.n_iter = n_warmup; p.adapt = 1; p.draw_base = 0;
  {
    ChainRunner<Target, 1, kEmuK, true> r(p, grp, scratch.data(), sc_shared,
                                           chain_smem.data());
    ChainRunner<Target, 1, kEmuK, true, double, true> rf(p, grp, scratch.data(), sc_shared,
                                                         chain_smem.data());
    r.dc = rf.dc = &decision_cache;
    r.aq = rf.aq = &adam_queue;
    rf.rl = &run_limits;
    if (budget > 0) {
      p.n_iter = 1; p.free_cap = 0x7fffffff; p.iter_cap = n_warmup;
      while (static_cast<int>(sc.warm_iter) < n_warmup) { rf.advance(0); ++n_launch; }
    } else if (n_warmup > 0) {
      r.run(0);
    }
  }
  // freeze_kernel
  for (int i = 0; i < ld; ++i) {
    inv_mass[i] = metric_from_sums(est[1 * ld + i], est[3 * ld + i], sc.est_w);
  }
  sc.step = std::exp(sc.adam_x);
  sc.min_micro = min_micro_steps(sc, p);
  *step_out = sc.step;
  *min_micro_out = sc.min_micro;
  std::memcpy(inv_mass_out, inv_mass.data(), D * 8);
  p.n_iter = n_sampling; p.adapt = 0; p.draw_base = n_warmup; p.im_out = nullptr;
  {
    ChainRunner<Target, 1, kEmuK, false> r(p, grp, scratch.data(), sc_shared,
                                            chain_smem.data());
    ChainRunner<Target, 1, kEmuK, false, double, true> rf(p, grp, scratch.data(), sc_shared,
                                                          chain_smem.data());
    r.dc = rf.dc = &decision_cache;
    rf.rl = &run_limits;
    if (budget > 0) {
      p.n_iter = 1; p.free_cap = 0x7fffffff; p.iter_cap = n_sampling;
      sc.eval_debt = 0;  // freeze_kernel
      while (static_cast<int>(sc.lp_n) < n_sampling) { rf.advance(0); ++n_launch; }
    } else if (n_sampling > 0) {
      r.run(0);
    }
  }
  if (launches) *launches = n_launch;
  for (int i = 0; i < total; ++i) {
    std::memcpy(draws + static_cast<size_t>(i) * D, d_draws.data() + static_cast<size_t>(i) * ld, D * 8);
    if (i < n_warmup && im_trace) {
      std::memcpy(im_trace + static_cast<size_t>(i) * D, d_im.data() + static_cast<size_t>(i) * ld, D * 8);
    }
  }
  *evals = sc.grad_evals;
}

extern "C" int emu_run_chain(int kind, int D, const double* tparam, const EmuTuning* t,
                             uint32_t seed, uint32_t chain, const double* theta0,
                             const double* mass0, double step0, int n_warmup,
                             int n_sampling, double* draws, double* lp, int* depth,
                             double* step_trace, double* im_trace, double* inv_mass_out,
                             double* step_out, int* min_micro_out,
                             unsigned long long* evals) {
  if (D > 2 * kEmuK || t->max_depth > kMaxDepth) return -1;
  switch (kind) {
    case 0: run<StdNormalTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0, n_warmup,
                                 n_sampling, draws, lp, depth, step_trace, im_trace,
                                 inv_mass_out, step_out, min_micro_out, evals); break;
    case 1: run<DiagGaussianTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0,
                                    n_warmup, n_sampling, draws, lp, depth, step_trace,
                                    im_trace, inv_mass_out, step_out, min_micro_out, evals); break;
    case 2: run<FunnelTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0, n_warmup,
                              n_sampling, draws, lp, depth, step_trace, im_trace,
                              inv_mass_out, step_out, min_micro_out, evals); break;
    default: return -2;
  }
  return 0;
}

// the same chain advanced by free-running launches of `budget` evaluations each
extern "C" int emu_run_chain_free(int kind, int D, const double* tparam, const EmuTuning* t,
                                  uint32_t seed, uint32_t chain, const double* theta0,
                                  const double* mass0, double step0, int n_warmup,
                                  int n_sampling, long long budget, double* draws,
                                  double* lp, int* depth, double* step_trace,
                                  double* im_trace, double* inv_mass_out, double* step_out,
                                  int* min_micro_out, unsigned long long* evals,
                                  int* launches) {
  if (D > 2 * kEmuK || t->max_depth > kMaxDepth || budget <= 0) return -1;
  switch (kind) {
    case 0: run<StdNormalTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0, n_warmup,
                                 n_sampling, draws, lp, depth, step_trace, im_trace,
                                 inv_mass_out, step_out, min_micro_out, evals, budget,
                                 launches); break;
    case 1: run<DiagGaussianTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0,
                                    n_warmup, n_sampling, draws, lp, depth, step_trace,
                                    im_trace, inv_mass_out, step_out, min_micro_out, evals,
                                    budget, launches); break;
    case 2: run<FunnelTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0, n_warmup,
                              n_sampling, draws, lp, depth, step_trace, im_trace,
                              inv_mass_out, step_out, min_micro_out, evals, budget,
                              launches); break;
    default: return -2;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// The lock-step tick engine (walnuts_b200/csrc/tick_kernel.cuh), same harness:
// the gradient stage between ticks is the element-wise target evaluated on the
// posted position, so the run must again equal the oracle bit for bit.
#include "../../walnuts_b200/csrc/tick_kernel.cuh"

template <template <int, int> class TargetT>
static void run_tick(const EmuTuning& t, int D, const double* tparam, uint32_t seed,
                     uint32_t chain, const double* theta0, const double* mass0,
                     double step0, int n_warmup, int n_sampling, double* draws, double* lp,
                     int* depth, double* step_trace, double* im_trace, double* inv_mass_out,
                     double* step_out, int* min_micro_out, unsigned long long* evals,
                     int warm_ticks = -1, int samp_ticks = -1, long long* rows_out = nullptr) {
  // warm_ticks / samp_ticks >= 0: that phase runs free (exactly that many ticks, as
  // tick_run_ticks does) instead of an iteration quota; n_warmup + n_sampling is then
  // only the draw capacity and rows_out = {rows after warm-up, rows at the end}.
  const int ld = 2 * kEmuK;  // rows padded to the group's element slots
  const int total = n_warmup + n_sampling;
  const int nvec = tick_vectors(t.max_depth);
  std::vector<double> theta(ld, 0.0), inv_mass(ld, 0.0), est(4 * ld, 0.0), tp_(ld, 0.0);
  std::vector<double> TH(ld, 0.0), G(ld, 0.0), vecs(static_cast<size_t>(nvec) * ld, 0.0);
  std::vector<double> d_draws(static_cast<size_t>(total) * ld), d_im(static_cast<size_t>(total) * ld);
  double LP = 0.0;
  std::memcpy(theta.data(), theta0, D * 8);
  if (tparam) std::memcpy(tp_.data(), tparam, D * 8);
  for (int i = 0; i < ld; ++i) {
    double m = i < D ? mass0[i] : 1.0;
    est[1 * ld + i] = t.mass_init_count * (1.0 / m);
    est[3 * ld + i] = t.mass_init_count * m;
  }
  ChainScalars sc{};
  sc.adam_x = std::log(step0);
  sc.adam_b1p = 1.0; sc.adam_b2p = 1.0;
  sc.mm_total = 2.0; sc.mm_count = 1.0;
  sc.est_w = t.mass_init_count;
  sc.step = step0;
  sc.min_micro = t.min_micro;
  TickState ts{};
  int active = 0;
  TickParams tp{};
  ChainParams& p = tp.cp;
  p.C = 1; p.D = D; p.ld = ld;
  p.max_depth = t.max_depth; p.max_halvings = t.max_halvings; p.min_micro_cfg = t.min_micro;
  p.max_error = t.max_error; p.mass_init_count = t.mass_init_count;
  p.macro_target = t.macro_target;
  p.adam_target = t.adam_target; p.adam_lr = t.adam_lr; p.adam_b1 = t.adam_b1;
  p.adam_b2 = t.adam_b2; p.adam_eps = t.adam_eps; p.adam_decay = t.adam_decay;
  p.seed = seed; p.chain_offset = chain;
  p.theta = theta.data(); p.inv_mass = inv_mass.data(); p.est = est.data(); p.sc = &sc;
  p.draws = d_draws.data(); p.draw_cap = total;
  p.lp_out = lp; p.depth_out = depth; p.step_out = step_trace; p.im_out = d_im.data();
  p.tparam = tp_.data();
  tp.TH = TH.data(); tp.G = G.data(); tp.LP = &LP; tp.vecs = vecs.data();
  std::vector<double> stk_logW(kMaxDepth, 0.0), stk_lp(kMaxDepth, 0.0);
  tp.st_logW = stk_logW.data(); tp.st_lp = stk_lp.data();
  tp.vec_stride = vecs.size(); tp.ts = &ts; tp.active_count = &active;
  Group<1> grp{};
  using Target = TargetT<1, kEmuK>;
  using V = Vec<1, kEmuK>;
  Target tgt;
  tgt.init(p, 0);
  auto gradient_stage = [&](const double* pos, double* grad_out, double& lp_out) {
    double x[kEmuK][2], gg[kEmuK][2], part;
    V::load(pos, ld, 0, x);
    tgt.grad(x, gg, part, grp);
    V::store(grad_out, ld, 0, gg);
    lp_out = part;
  };
  // initial gradient at theta0 (what session init provides)
  gradient_stage(theta.data(), vecs.data() + static_cast<size_t>(TV_CUR_G) * ld, ts.lp_cur);
  std::memcpy(vecs.data() + static_cast<size_t>(TV_CUR) * ld, theta.data(), ld * 8);
  unsigned long long logical = 1;
  auto run_batch = [&](int n_iter, int adapt, long long base) {
    if (n_iter == 0) return;
    p.n_iter = n_iter; p.adapt = adapt; p.draw_base = base;
    p.im_out = adapt ? d_im.data() : nullptr;
    ts.pc = PC_START_TRANSITION; ts.done_iters = 0;
    while (true) {
      active = 0;
      TickRunner<1, kEmuK> r(tp, grp);
      r.tick(0);
      if (active == 0) break;
      gradient_stage(TH.data(), G.data(), LP);
    }
  };
  auto run_free = [&](int n_ticks, int adapt) {
    p.n_iter = -1; p.adapt = adapt; p.draw_base = 0;
    p.im_out = adapt ? d_im.data() : nullptr;
    if (ts.pc == PC_DONE) ts.pc = PC_START_TRANSITION;  // tick_resume_kernel
    for (int k = 0; k < n_ticks; ++k) {
      active = 0;
      TickRunner<1, kEmuK> r(tp, grp);
      r.tick(0);
      gradient_stage(TH.data(), G.data(), LP);
    }
  };
  if (warm_ticks >= 0) {
    run_free(warm_ticks, 1);
    ts.pc = PC_DONE;  // freeze abandons the transition in flight (tick_abort_kernel)
  } else {
    run_batch(n_warmup, 1, 0);
  }
  if (rows_out) rows_out[0] = warm_ticks >= 0 ? ts.rows : n_warmup;
  for (int i = 0; i < ld; ++i) {
    inv_mass[i] = metric_from_sums(est[1 * ld + i], est[3 * ld + i], sc.est_w);
  }
  sc.step = std::exp(sc.adam_x);
  sc.min_micro = min_micro_steps(sc, p);
  *step_out = sc.step;
  *min_micro_out = sc.min_micro;
  std::memcpy(inv_mass_out, inv_mass.data(), D * 8);
  if (samp_ticks >= 0) {
    if (warm_ticks < 0 && ts.rows < n_warmup) ts.rows = n_warmup;  // rows_floor
    run_free(samp_ticks, 0);
  } else {
    run_batch(n_sampling, 0, warm_ticks >= 0 ? ts.rows : n_warmup);
  }
  if (rows_out) rows_out[1] = samp_ticks >= 0 ? ts.rows : rows_out[0] + n_sampling;
  for (int i = 0; i < total; ++i) {
    std::memcpy(draws + static_cast<size_t>(i) * D, d_draws.data() + static_cast<size_t>(i) * ld, D * 8);
    if (i < n_warmup && im_trace) {
      std::memcpy(im_trace + static_cast<size_t>(i) * D, d_im.data() + static_cast<size_t>(i) * ld, D * 8);
    }
  }
  // the reference re-evaluates the gradient at the start of every transition
  *evals = sc.grad_evals + logical + static_cast<unsigned long long>(total > 0 ? total - 1 : 0);
}

extern "C" int emu_run_chain_tick(int kind, int D, const double* tparam, const EmuTuning* t,
                                  uint32_t seed, uint32_t chain, const double* theta0,
                                  const double* mass0, double step0, int n_warmup,
                                  int n_sampling, double* draws, double* lp, int* depth,
                                  double* step_trace, double* im_trace,
                                  double* inv_mass_out, double* step_out,
                                  int* min_micro_out, unsigned long long* evals) {
  if (D > 2 * kEmuK || t->max_depth > kMaxDepth) return -1;
  switch (kind) {
    case 0: run_tick<StdNormalTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0,
                                      n_warmup, n_sampling, draws, lp, depth, step_trace,
                                      im_trace, inv_mass_out, step_out, min_micro_out, evals); break;
    case 1: run_tick<DiagGaussianTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0,
                                         n_warmup, n_sampling, draws, lp, depth, step_trace,
                                         im_trace, inv_mass_out, step_out, min_micro_out, evals); break;
    case 2: run_tick<FunnelTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0, n_warmup,
                                   n_sampling, draws, lp, depth, step_trace, im_trace,
                                   inv_mass_out, step_out, min_micro_out, evals); break;
    default: return -2;
  }
  return 0;
}

// free-running variant: see run_tick
extern "C" int emu_run_chain_tick_free(int kind, int D, const double* tparam,
                                       const EmuTuning* t, uint32_t seed, uint32_t chain,
                                       const double* theta0, const double* mass0,
                                       double step0, int n_warmup, int n_sampling,
                                       int warm_ticks, int samp_ticks, double* draws,
                                       double* lp, int* depth, double* step_trace,
                                       double* im_trace, double* inv_mass_out,
                                       double* step_out, int* min_micro_out,
                                       unsigned long long* evals, long long* rows_out) {
  if (D > 2 * kEmuK || t->max_depth > kMaxDepth) return -1;
  switch (kind) {
    case 0: run_tick<StdNormalTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0,
                                      n_warmup, n_sampling, draws, lp, depth, step_trace,
                                      im_trace, inv_mass_out, step_out, min_micro_out, evals,
                                      warm_ticks, samp_ticks, rows_out); break;
    case 1: run_tick<DiagGaussianTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0,
                                         n_warmup, n_sampling, draws, lp, depth, step_trace,
                                         im_trace, inv_mass_out, step_out, min_micro_out,
                                         evals, warm_ticks, samp_ticks, rows_out); break;
    case 2: run_tick<FunnelTarget>(*t, D, tparam, seed, chain, theta0, mass0, step0, n_warmup,
                                   n_sampling, draws, lp, depth, step_trace, im_trace,
                                   inv_mass_out, step_out, min_micro_out, evals, warm_ticks,
                                   samp_ticks, rows_out); break;
    default: return -2;
  }
  return 0;
}
