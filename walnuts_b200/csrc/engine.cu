// Kernel instantiations and launchers of the device sampler.
#include "engine.cuh"
#include "engine_shapes.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

namespace wb200 {

// ---------------------------------------------------------------------------
// (threads per chain, register chunks) by dimension: 2*T*K element slots
LaunchShape shape_for_dim(int D) {
  if (D <= 64) return {32, 1, 128, 4};
  if (D <= 128) return {32, 2, 128, 4};
  if (D <= 256) return {64, 2, 64, 1};
  if (D <= 512) return {128, 2, 128, 1};
  if (D <= 1024) {
    // WB200_SHAPE_1024=256x2 selects the 8-warp variant (experiments)
    const char* e = std::getenv("WB200_SHAPE_1024");
    if (e && std::string(e) == "256x2") return {256, 2, 256, 1};
    return {128, 4, 128, 1};
  }
  if (D <= 2048) return {256, 4, 256, 1};
  if (D <= 4096) return {512, 4, 512, 1};
  throw std::invalid_argument("num_params above 4096 is not supported by the "
                              "chain-resident kernel");
}

// ---------------------------------------------------------------------------
// Batched initialisation: InitConfigBuilder::positions(rng, scale)
// (config.hpp:259-268), ::masses(F, s) (:360-370), adapt_step (util.hpp:285-303)
// and the constructors of MassEstimator / Adam / MinMicroStepsAdaptHandler
// (adaptive_walnuts.hpp:54-62, adam.hpp:48-66, adaptive_walnuts.hpp:127-132).
struct InitParams {
  ChainParams cp;
  int have_positions, have_mass, have_steps;
  double init_radius, smoothing, step_init;
  double* mass;   // [C][ld] in: given masses (if have_mass); out: masses used
  double* steps;  // [C] in/out
};

template <template <int, int> class TargetT, int T, int K, int CTA>
__global__ void __launch_bounds__(CTA) init_kernel(const InitParams ip) {
  using Target = TargetT<T, K>;
  using V = Vec<T, K>;
  __shared__ double red_smem[group_smem_doubles<T>()];
  const ChainParams& p = ip.cp;
  Group<T> grp;
  grp.lane = threadIdx.x & 31;
  grp.red = red_smem;
  grp.parity = 0;
  int chain;
  if constexpr (T == 32) {
    grp.tid = grp.lane; grp.warp = 0;
    chain = blockIdx.x * (CTA / 32) + (threadIdx.x >> 5);
  } else {
    grp.tid = threadIdx.x; grp.warp = threadIdx.x >> 5;
    chain = blockIdx.x;
  }
  if (chain >= p.C) return;  // whole group exits together
  const int tid = grp.tid, ld = p.ld;
  const uint32_t gchain = p.chain_offset + chain;
  Target tgt;
  tgt.init(p, tid);
  double th[K][2], g[K][2], mass[K][2];
  double* theta_row = p.theta + static_cast<long long>(chain) * ld;
  if (ip.have_positions) {
    V::load(theta_row, ld, tid, th);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int j = tid + k * T;
      double z0 = 0.0, z1 = 0.0;
      if (2 * j < p.D) {
        philox_normal_pair(p.seed, gchain, 0u, kKindInit, j, z0, z1);
        if (2 * j + 1 >= p.D) z1 = 0.0;
      }
      th[k][0] = z0 * ip.init_radius;
      th[k][1] = z1 * ip.init_radius;
    }
    V::store(theta_row, ld, tid, th);
  }
  double lp_part;
  tgt.grad(th, g, lp_part, grp);
  double* mass_row = ip.mass + static_cast<long long>(chain) * ld;
  if (ip.have_mass) {
    V::load(mass_row, ld, tid, mass);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        mass[k][v] = (1 - ip.smoothing) * fabs(g[k][v]) + ip.smoothing;
      }
    }
  }
  // padding lanes keep mass 1 so that every later quotient stays finite
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if (2 * (tid + k * T) + v >= p.D) mass[k][v] = 1.0;
    }
  }
  V::store(mass_row, ld, tid, mass);
  // estimators: mean 0, S = w0 * var0 (online_moments.hpp:151-159)
  double* est_row = p.est + static_cast<long long>(chain) * 4 * ld;
  {
    double zero[K][2], sd[K][2], ss[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        zero[k][v] = 0.0;
        sd[k][v] = p.mass_init_count * (1.0 / mass[k][v]);
        ss[k][v] = p.mass_init_count * mass[k][v];
      }
    }
    V::store(est_row + 0 * ld, ld, tid, zero);
    V::store(est_row + 1 * ld, ld, tid, sd);
    V::store(est_row + 2 * ld, ld, tid, zero);
    V::store(est_row + 3 * ld, ld, tid, ss);
  }
  double step = ip.have_steps ? ip.steps[chain] : ip.step_init;
  if (!ip.have_steps) {
    // adapt_step, util.hpp:285-303, with leapfrog_error :242-259
    double invM[K][2], rho[K][2];
    double kin0 = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int j = tid + k * T;
      double z0 = 0.0, z1 = 0.0;
      if (2 * j < p.D) {
        philox_normal_pair(p.seed, gchain, 0u, kKindStepInit, j, z0, z1);
        if (2 * j + 1 >= p.D) z1 = 0.0;
      }
      invM[k][0] = 1.0 / mass[k][0]; invM[k][1] = 1.0 / mass[k][1];
      rho[k][0] = z0 * sqrt(mass[k][0]); rho[k][1] = z1 * sqrt(mass[k][1]);
      kin0 = madd(invM[k][0], rho[k][0] * rho[k][0], kin0);
      kin0 = madd(invM[k][1], rho[k][1] * rho[k][1], kin0);
    }
    double r0[2] = {lp_part, kin0};
    grp.sum(r0);
    const double H0 = r0[0] + (-0.5 * r0[1]);
    auto lf_error = [&](double s) -> double {
      double rs[K][2], ts[K][2], g2[K][2];
      const double hs = 0.5 * s;
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          rs[k][v] = rho[k][v] + hs * g[k][v];
          ts[k][v] = th[k][v] + s * (invM[k][v] * rs[k][v]);
        }
      }
      double lp2;
      tgt.grad(ts, g2, lp2, grp);
      double kin = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          rs[k][v] = rs[k][v] + hs * g2[k][v];
          kin = madd(invM[k][v], rs[k][v] * rs[k][v], kin);
        }
      }
      double r[2] = {lp2, kin};
      grp.sum(r);
      return (r[0] + (-0.5 * r[1])) - H0;
    };
    const double log09 = log(0.9), log06 = log(0.6), rt = sqrt(0.5);
    for (int guard = 0; guard < 2000 && lf_error(step) > log09; ++guard) step *= 2;
    for (int guard = 0; guard < 2000 && lf_error(step) < log06; ++guard) step *= rt;
  }
  if (tid == 0) {
    ip.steps[chain] = step;
    ChainScalars sc{};
    sc.adam_x = log(step);
    sc.adam_b1p = 1.0; sc.adam_b2p = 1.0;
    sc.mm_total = 2.0; sc.mm_count = 1.0;
    sc.est_w = p.mass_init_count;
    sc.step = step;
    sc.min_micro = p.min_micro_cfg;
    p.sc[chain] = sc;
  }
}

// AdaptiveWalnuts::sampler(), adaptive_walnuts.hpp:263-271
__global__ void freeze_kernel(ChainParams p) {
  const long long n = static_cast<long long>(p.C) * p.ld;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
       i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i / p.ld), e = static_cast<int>(i % p.ld);
    const double* est_row = p.est + static_cast<long long>(c) * 4 * p.ld;
    const double w = p.sc[c].est_w;
    p.inv_mass[i] = metric_from_sums(est_row[1 * p.ld + e], est_row[3 * p.ld + e], w);
    if (e == 0) {
      ChainScalars& sc = p.sc[c];
      sc.step = exp(sc.adam_x);
      sc.min_micro = min_micro_steps(sc, p);
      sc.eval_debt = 0;  // sampling starts every chain level
    }
  }
}

// fixed-step orbit for parity checks (walnuts.hpp:329-332 repeated)
struct OrbitParams {
  ChainParams cp;
  double* rho; double* grad; double* logp; double* joint;
  double step; int num_steps;
};

template <template <int, int> class TargetT, int T, int K, int CTA, class Real>
__global__ void __launch_bounds__(CTA) orbit_kernel(const OrbitParams op) {
  using Target = TargetT<T, K>;
  using V = VecT<T, K, Real>;
  __shared__ double red_smem[group_smem_doubles<T>()];
  const ChainParams& p = op.cp;
  Group<T> grp;
  grp.lane = threadIdx.x & 31; grp.red = red_smem; grp.parity = 0;
  int chain;
  if constexpr (T == 32) {
    grp.tid = grp.lane; grp.warp = 0;
    chain = blockIdx.x * (CTA / 32) + (threadIdx.x >> 5);
  } else {
    grp.tid = threadIdx.x; grp.warp = threadIdx.x >> 5;
    chain = blockIdx.x;
  }
  if (chain >= p.C) return;
  ChainScalars unused_sc{};
  ChainRunner<Target, T, K, false, Real> r(p, grp, nullptr, unused_sc, nullptr);
  r.tgt.init(p, grp.tid);
  const long long off = static_cast<long long>(chain) * p.ld;
  V::load64(p.theta + off, grp.tid, r.th);
  V::load64(op.rho + off, grp.tid, r.rho);
  V::load64(p.inv_mass + off, grp.tid, r.im);
  Real lp_part;
  r.tgt.grad(r.th, r.g, lp_part, grp);
  r.evals = 0;
  double lp, H;
  if (op.num_steps > 0) {
    double d0, d1;
    r.integrate(op.num_steps, op.step, lp, H, false, d0, d1);
  } else {
    Real kin = 0;
    for (int k = 0; k < K; ++k)
      for (int v = 0; v < 2; ++v) kin = madd(r.im[k][v], r.rho[k][v] * r.rho[k][v], kin);
    double s[2] = {static_cast<double>(lp_part), static_cast<double>(kin)};
    grp.sum(s);
    lp = s[0]; H = s[0] + (-0.5 * s[1]);
  }
  V::store64(p.theta + off, grp.tid, r.th);
  V::store64(op.rho + off, grp.tid, r.rho);
  V::store64(op.grad + off, grp.tid, r.g);
  if (grp.tid == 0) { op.logp[chain] = lp; op.joint[chain] = H; }
}

static int sm_count(int device) {
  int n = 0;
  WB200_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
  return n;
}

#define WB200_OCC(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)                      \
  do {                                                                         \
    occ_adapt = blocks_per_sm(                                                 \
        walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_A_, true>, CTA_, dyn_smem); \
    occ_sample = blocks_per_sm(                                                \
        walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_S_, false>, CTA_, dyn_smem); \
  } while (0)
#define WB200_LAUNCH_CHAIN(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)             \
  do {                                                                         \
    if (p.adapt) {                                                             \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_A_, true>        \
          <<<s.grid_adapt, CTA_, dyn_smem, s.stream>>>(p);                     \
    } else {                                                                   \
      walnuts_chain_kernel<TARGET<T_, K_>, T_, K_, CTA_, MINB_S_, false>       \
          <<<s.grid, CTA_, dyn_smem, s.stream>>>(p);                           \
    }                                                                          \
  } while (0)
#define WB200_LAUNCH_INIT(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)              \
  init_kernel<TARGET, T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(ip)
#define WB200_LAUNCH_ORBIT(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)             \
  orbit_kernel<TARGET, T_, K_, CTA_, double><<<grid, CTA_, 0, stream>>>(op)
#define WB200_LAUNCH_ORBIT_F32(TARGET, T_, K_, CTA_, MINB_A_, MINB_S_)         \
  orbit_kernel<TARGET, T_, K_, CTA_, float><<<grid, CTA_, 0, stream>>>(op)

// resident CTAs per SM of the adaptive and of the sampling instance
void occupancy_for(int kind, const LaunchShape& shape, int ld, int precision, int* adapt,
                   int* sample) {
  int occ_adapt = 1, occ_sample = 1;
  const size_t dyn_smem = chain_dyn_smem(shape, ld);
  if (precision == 1) {
    occupancy_f32(kind, shape, dyn_smem, &occ_adapt, &occ_sample);
  } else {
    WB200_FOR_TARGET(kind, shape, WB200_OCC);
  }
  *adapt = occ_adapt;
  *sample = occ_sample;
}

// Longest-processing-time-first hand-out.  A launch ends when its slowest ticket does, and
// orbit lengths differ by orders of magnitude where the geometry varies (the funnel:
// 2^1 .. 2^10 leaves of 1 .. 2^8 micro-steps) but are strongly autocorrelated along a
// chain -- so the gradient evaluations a chain needed in its previous launch predict the
// next one.  Chains are bucketed by half-octaves of that count (counting sort, one CTA)
// and tickets walk the buckets from the most expensive down; light chains fill the tail.
// Scheduling only: every chain computes exactly what it would in any other order.
constexpr int kOrderBuckets = 96;
__global__ void __launch_bounds__(1024)
lpt_order_kernel(const ChainScalars* sc, unsigned long long* prev, int C, int* order) {
  __shared__ int count[kOrderBuckets], start[kOrderBuckets];
  for (int b = threadIdx.x; b < kOrderBuckets; b += blockDim.x) count[b] = 0;
  __syncthreads();
  auto bucket_of = [](unsigned long long cost) {
    // descending cost: bucket 0 = the most expensive; two buckets per octave
    if (cost == 0) return kOrderBuckets - 1;
    const int lg = 63 - __clzll(static_cast<long long>(cost));
    const int half = (cost >> (lg > 0 ? lg - 1 : 0)) & 1;
    const int key = 2 * lg + (lg > 0 ? half : 0);
    return max(0, kOrderBuckets - 2 - key);
  };
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&count[bucket_of(sc[c].grad_evals - prev[c])], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < kOrderBuckets; ++b) { start[b] = acc; acc += count[b]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const unsigned long long now = sc[c].grad_evals;
    order[atomicAdd(&start[bucket_of(now - prev[c])], 1)] = c;
    prev[c] = now;
  }
}

void launch_chains(wb200_session& s, int n_iter, int adapt, bool store,
                   long long eval_budget, long long iter_cap, long long* rows) {
  ChainParams p = s.params(eval_budget > 0 ? 1 : n_iter, adapt, store);
  p.eval_budget = eval_budget;  // > 0: free-running launch (chain_kernel.cuh), one
  p.free_cap = n_iter;          // transition per ChainRunner::run, at most n_iter of them
  p.iter_cap = iter_cap;
  p.rows = rows;
  const size_t dyn_smem = chain_dyn_smem(s.shape, s.ld);
  if (s.C > s.slots) {  // more chains than resident groups: the hand-out order matters
    if (s.order.count == 0) {
      s.order.alloc(s.C);
      s.prev_evals.alloc(s.C);
      WB200_CUDA(cudaMemsetAsync(s.prev_evals.ptr, 0, s.C * sizeof(unsigned long long),
                                 s.stream));
    }
    lpt_order_kernel<<<1, 1024, 0, s.stream>>>(s.sc.ptr, s.prev_evals.ptr, s.C, s.order.ptr);
    WB200_CUDA(cudaGetLastError());
    s.launches += 1;
    p.order = s.order.ptr;
  }
  WB200_CUDA(cudaMemsetAsync(s.ticket.ptr, 0, sizeof(unsigned int), s.stream));
  WB200_CUDA(cudaEventRecord(s.ev0, s.stream));
  if (p.eval_budget > 0) {  // the free-running kernels (engine_free*.cu)
    if (s.precision == 1) launch_chain_free_f32(s, p, dyn_smem);
    else launch_chain_free(s, p, dyn_smem);
  } else if (s.precision == 1) {
    launch_chain_f32(s, p, dyn_smem);
  } else {
    WB200_FOR_TARGET(s.kind, s.shape, WB200_LAUNCH_CHAIN);
  }
  WB200_CUDA(cudaGetLastError());
  WB200_CUDA(cudaEventRecord(s.ev1, s.stream));
  s.launches += 1;
  if (store && eval_budget == 0) s.rows_written += n_iter;
}

// per-chain iteration counts of the current phase (warm-up: AdaptiveWalnuts::iteration_,
// sampling: draws observed by the lp accumulator): {min, max, sum}, then the gradient
// evaluations of all chains so far -- what the reference's
// controllers read from the chains' snapshots (adapt.hpp:196-203, sampler.hpp:134-141)
__global__ void __launch_bounds__(1024)
iter_stats_kernel(const ChainScalars* sc, int C, int sampling, long long* out4) {
  __shared__ long long smin[32], smax[32], ssum[32], sev[32];
  long long mn = 0x7fffffffffffffffll, mx = 0, sm = 0, ev = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long long n = sampling ? static_cast<long long>(sc[c].lp_n)
                                 : static_cast<long long>(sc[c].warm_iter);
    mn = n < mn ? n : mn; mx = n > mx ? n : mx; sm += n;
    ev += static_cast<long long>(sc[c].grad_evals);
  }
  for (int m = 16; m >= 1; m >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, mn, m);
    const long long b = __shfl_xor_sync(0xffffffffu, mx, m);
    sm += __shfl_xor_sync(0xffffffffu, sm, m);
    ev += __shfl_xor_sync(0xffffffffu, ev, m);
    mn = a < mn ? a : mn; mx = b > mx ? b : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; ssum[threadIdx.x >> 5] = sm;
    sev[threadIdx.x >> 5] = ev;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) {
      mn = smin[w] < mn ? smin[w] : mn; mx = smax[w] > mx ? smax[w] : mx; sm += ssum[w];
      ev += sev[w];
    }
    out4[0] = mn; out4[1] = mx; out4[2] = sm; out4[3] = ev;
  }
}

void chain_iter_stats(wb200_session& s, bool sampling, long long* out4_host) {
  if (s.iter_stats.count == 0) s.iter_stats.alloc(4);
  iter_stats_kernel<<<1, 1024, 0, s.stream>>>(s.sc.ptr, s.C, sampling ? 1 : 0,
                                              s.iter_stats.ptr);
  WB200_CUDA(cudaGetLastError());
  s.launches += 1;
  WB200_CUDA(cudaMemcpyAsync(out4_host, s.iter_stats.ptr, 4 * sizeof(long long),
                             cudaMemcpyDeviceToHost, s.stream));
  WB200_CUDA(cudaStreamSynchronize(s.stream));
}

void launch_init(wb200_session& s, bool have_mass, bool have_steps,
                 bool have_positions, double init_radius) {
  InitParams ip{};
  ip.cp = s.params(0, 1, false);
  ip.have_positions = have_positions;
  ip.have_mass = have_mass;
  ip.have_steps = have_steps;
  ip.init_radius = init_radius;
  ip.smoothing = s.tuning.mass_additive_smoothing;
  ip.step_init = s.tuning.step_size_init;
  ip.mass = s.inv_mass.ptr;  // the mass rows are staged in the inv_mass buffer
  ip.steps = s.red.ptr;      // [C] staged by the caller
  const int grid = (s.C + s.shape.chains_per_cta - 1) / s.shape.chains_per_cta;
  WB200_FOR_TARGET(s.kind, s.shape, WB200_LAUNCH_INIT);
  WB200_CUDA(cudaGetLastError());
  s.launches += 1;
}

void launch_freeze(wb200_session& s) {
  ChainParams p = s.params(0, 0, false);
  const long long n = static_cast<long long>(s.C) * s.ld;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  freeze_kernel<<<grid, 256, 0, s.stream>>>(p);
  WB200_CUDA(cudaGetLastError());
  s.launches += 1;
}

void launch_orbit(int kind, int precision, int D, int ld, int C, const double* tparam,
                  double* theta, double* rho, const double* inv_mass, double* grad,
                  double* logp, double* joint, double step, int num_steps,
                  cudaStream_t stream) {
  OrbitParams op{};
  op.cp.C = C; op.cp.D = D; op.cp.ld = ld;
  op.cp.theta = theta;
  op.cp.inv_mass = const_cast<double*>(inv_mass);
  op.cp.tparam = tparam;
  op.cp.max_halvings = 1;
  op.rho = rho; op.grad = grad; op.logp = logp; op.joint = joint;
  op.step = step; op.num_steps = num_steps;
  LaunchShape shape = shape_for_dim(D);
  const int grid = (C + shape.chains_per_cta - 1) / shape.chains_per_cta;
  if (precision == 1) {
    WB200_FOR_TARGET_F32(kind, shape, WB200_LAUNCH_ORBIT_F32);
  } else {
    WB200_FOR_TARGET(kind, shape, WB200_LAUNCH_ORBIT);
  }
  WB200_CUDA(cudaGetLastError());
}

}  // namespace wb200

namespace wb200 {
// adam.hpp:82: learning_rate / pow(t, decay), the same expression adam_update evaluates
__global__ void adam_table_kernel(double* tab, int n, double lr, double decay) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tab[i] = lr / pow(static_cast<double>(i + 1), decay);
}
}  // namespace wb200

wb200::ChainParams wb200_session::params(int n_iter, int adapt, bool store) {
  wb200::ChainParams p{};
  if (adapt && adam_tab.count == 0) {
    constexpr int kAdamTable = 1 << 16;  // updates per chain covered; beyond: pow on the spot
    adam_tab.alloc(kAdamTable);
    wb200::adam_table_kernel<<<(kAdamTable + 255) / 256, 256, 0, stream>>>(
        adam_tab.ptr, kAdamTable, tuning.step_learning_rate, tuning.step_learn_rate_decay);
    WB200_CUDA(cudaGetLastError());
  }
  p.adam_tab = adapt ? adam_tab.ptr : nullptr;
  p.adam_tab_n = adapt ? static_cast<int>(adam_tab.count) : 0;
  p.C = C; p.D = D; p.ld = ld;
  p.n_iter = n_iter;
  p.adapt = adapt;
  p.max_depth = tuning.max_trajectory_doublings;
  p.max_halvings = tuning.max_step_halvings;
  p.min_micro_cfg = tuning.min_micro_steps;
  p.max_error = tuning.max_hamiltonian_error;
  p.mass_init_count = tuning.mass_init_count;
  p.macro_target = tuning.max_macro_steps_target;
  p.adam_target = tuning.step_accept_rate_target;
  p.adam_lr = tuning.step_learning_rate;
  p.adam_b1 = tuning.step_gradient_decay;
  p.adam_b2 = tuning.step_sq_gradient_decay;
  p.adam_eps = tuning.step_stabilization;
  p.adam_decay = tuning.step_learn_rate_decay;
  p.seed = seed;
  p.chain_offset = chain_offset;
  p.theta = theta.ptr;
  p.inv_mass = inv_mass.ptr;
  p.est = est.ptr;
  p.sc = sc.ptr;
  p.draw_cap = draw_cap;
  p.draw_base = rows_written;
  if (store) {
    p.draws = draws.ptr;
    if (trace) {
      p.lp_out = lp_out.ptr;
      p.depth_out = depth_out.ptr;
      p.step_out = step_out.ptr;
      p.im_out = adapt ? im_out.ptr : nullptr;
    }
  }
  p.scratch = scratch.ptr;
  p.scratch_stride =
      wb200::scratch_doubles(tuning.max_trajectory_doublings, ld);
  p.ticket = ticket.ptr;
  p.tparam = tparam.ptr;
  return p;
}
