// Kernel instantiations and launchers of the device sampler.
#include "engine.cuh"
#include "engine_shapes.cuh"
#include "engine_kernels.cuh"
#include "user_density.cuh"

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
    if (e && std::string(e) == "64x8") return {64, 8, 64, 1};
    return {128, 4, 128, 1};
  }
  if (D <= 2048) return {256, 4, 256, 1};
  if (D <= 4096) return {512, 4, 512, 1};
  throw std::invalid_argument("num_params above 4096 is not supported by the "
                              "chain-resident kernel");
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
  if (s.kind == kDeviceSource) {  // run-time compiled density (user_density.cu)
    void* f = p.adapt ? (p.eval_budget > 0 ? s.user->adapt_free : s.user->adapt)
                      : (p.eval_budget > 0 ? s.user->sample_free : s.user->sample);
    user_launch(f, p.adapt ? s.grid_adapt : s.grid, s.shape.cta, dyn_smem, s.stream, &p);
  } else if (p.eval_budget > 0) {  // the free-running kernels (engine_free*.cu)
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
  if (s.kind == kDeviceSource) {
    user_launch(s.user->init, grid, s.shape.cta, 0, s.stream, &ip);
  } else {
    WB200_FOR_TARGET(s.kind, s.shape, WB200_LAUNCH_INIT);
  }
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
                  cudaStream_t stream, void* user_orbit) {
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
  if (user_orbit) {
    user_launch(user_orbit, grid, shape.cta, 0, stream, &op);
  } else if (precision == 1) {
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
