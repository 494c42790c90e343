// Host side of the lock-step tick engine (tick_kernel.cuh): buffers, batched
// initialisation, the tick / gradient loop.  Used for targets whose gradient is a
// cross-chain batched contraction (logistic regression, logistic.cu); element-wise
// targets can be routed through it too (WB200_ENGINE=tick) to test the plumbing.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "engine.cuh"
#include "logistic.cuh"
#include "tick_kernel.cuh"

namespace wb200 {

struct TickEngine {
  DeviceBuffer<double> TH, G, LP, vecs, st_logW, st_lp;
  DeviceBuffer<TickState> ts;
  DeviceBuffer<int> active;
  int* active_host = nullptr;  // pinned
  long long vec_stride = 0;
  // The batch is evaluated as two halves: in free-running mode each half runs its own
  // tick -> pack -> GEMM 1 -> GEMM 2 -> finalize chain on its own stream, so the
  // element-wise kernels of one half overlap the GEMMs of the other (the tick kernel
  // co-resides with GEMM 2's CTAs; GEMM 1 fills the register file).  logistic[1] is null
  // for small batches.
  LogisticGrad* logistic[2] = {nullptr, nullptr};
  int half_begin[2] = {0, 0}, half_count[2] = {0, 0};
  cudaStream_t half_stream[2] = {nullptr, nullptr};
  cudaEvent_t fork_event = nullptr, join_event[2] = {nullptr, nullptr};
  WB200_BATCH_LOGP_GRAD batch_fn = nullptr;  // kind 4: the caller's batched density
  void* batch_data = nullptr;
  unsigned long long ticks = 0, grad_batches = 0;
  ~TickEngine() {
    for (int h = 0; h < 2; ++h) {
      delete logistic[h];
      if (half_stream[h]) cudaStreamDestroy(half_stream[h]);
      if (join_event[h]) cudaEventDestroy(join_event[h]);
    }
    if (fork_event) cudaEventDestroy(fork_event);
    if (active_host) cudaFreeHost(active_host);
  }
};

// (T, K, CTA) dispatch shared with the chain kernel's shapes
#define WB200_TICK_SHAPE(S, MACRO)                                             \
  do {                                                                         \
    if ((S).T == 32 && (S).K == 1) { MACRO(32, 1, 128); }                      \
    else if ((S).T == 32 && (S).K == 2) { MACRO(32, 2, 128); }                 \
    else if ((S).T == 64 && (S).K == 2) { MACRO(64, 2, 64); }                  \
    else if ((S).T == 64) { MACRO(64, 4, 64); }                                \
    else if ((S).T == 128 && (S).K == 2) { MACRO(128, 2, 128); }               \
    else if ((S).T == 128 && (S).K == 4) { MACRO(128, 4, 128); }               \
    else if ((S).T == 256 && (S).K == 2) { MACRO(256, 2, 256); }               \
    else if ((S).T == 256 && (S).K == 4) { MACRO(256, 4, 256); }               \
    else { MACRO(512, 4, 512); }                                               \
  } while (0)

template <int T, int CTA>
__device__ __forceinline__ int group_setup(Group<T>& grp, double* red_smem) {
  grp.lane = threadIdx.x & 31;
  grp.red = red_smem;
  grp.parity = 0;
  if constexpr (T == 32) {
    grp.tid = grp.lane;
    grp.warp = 0;
    return blockIdx.x * (CTA / 32) + (threadIdx.x >> 5);
  } else {
    grp.tid = threadIdx.x;
    grp.warp = threadIdx.x >> 5;
    return blockIdx.x;
  }
}

// register slots beyond D (V::load fills them with 0) carry a unit mass so that the
// quotients of the step search stay finite
template <int T, int K>
__device__ __forceinline__ void unit_padding(double (&m)[K][2], int tid, int D) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if (2 * (tid + k * T) + v >= D) m[k][v] = 1.0;
    }
  }
}

// gradient stage for element-wise targets: G, LP at the posted positions TH
template <template <int, int> class TargetT, int T, int K, int CTA>
__global__ void __launch_bounds__(CTA) elementwise_grad_kernel(const TickParams tp) {
  __shared__ double red_smem[group_smem_doubles<T>()];
  Group<T> grp;
  const int chain = group_setup<T, CTA>(grp, red_smem);
  if (chain >= tp.cp.C) return;
  using V = Vec<T, K>;
  TargetT<T, K> tgt;
  tgt.init(tp.cp, grp.tid);
  double x[K][2], g[K][2], part;
  const long long off = static_cast<long long>(chain) * tp.cp.ld;
  V::load(tp.TH + off, tp.cp.ld, grp.tid, x);
  tgt.grad(x, g, part, grp);
  double r[1] = {part};
  grp.sum(r);
  V::store(tp.G + off, tp.cp.ld, grp.tid, g);
  if (grp.tid == 0) tp.LP[chain] = r[0];
}

__global__ void tick_begin_kernel(TickState* ts, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    ts[c].pc = PC_START_TRANSITION;
    ts[c].done_iters = 0;
  }
}

// ---- batched initialisation (config.hpp:259-268, :360-370; util.hpp:285-303) ----
struct TickInitParams {
  TickParams tp;
  int have_positions, have_mass, have_steps;
  double init_radius, smoothing, step_init;
  double* mass;   // [C][ld] staged masses (in/out)
  double* steps;  // [C] staged steps (in/out)
};

// positions -> TH (request) and TV_CUR
template <int T, int K, int CTA>
__global__ void __launch_bounds__(CTA) tick_init_positions_kernel(const TickInitParams ip) {
  __shared__ double red_smem[group_smem_doubles<T>()];
  Group<T> grp;
  const int chain = group_setup<T, CTA>(grp, red_smem);
  const ChainParams& p = ip.tp.cp;
  if (chain >= p.C) return;
  using V = Vec<T, K>;
  const int tid = grp.tid, ld = p.ld;
  const uint32_t gchain = p.chain_offset + chain;
  double th[K][2];
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
  V::store(ip.tp.TH + static_cast<long long>(chain) * ld, ld, tid, th);
  double* vb = ip.tp.vecs + static_cast<long long>(chain) * ip.tp.vec_stride;
  V::store(vb + static_cast<long long>(TV_CUR) * ld, ld, tid, th);
}

// after the first gradient: mass, estimators, scalars, search momentum
template <int T, int K, int CTA>
__global__ void __launch_bounds__(CTA) tick_init_state_kernel(const TickInitParams ip) {
  __shared__ double red_smem[group_smem_doubles<T>()];
  Group<T> grp;
  const int chain = group_setup<T, CTA>(grp, red_smem);
  const ChainParams& p = ip.tp.cp;
  if (chain >= p.C) return;
  using V = Vec<T, K>;
  const int tid = grp.tid, ld = p.ld;
  const uint32_t gchain = p.chain_offset + chain;
  const long long off = static_cast<long long>(chain) * ld;
  double* vb = ip.tp.vecs + static_cast<long long>(chain) * ip.tp.vec_stride;
  double g[K][2], mass[K][2];
  V::load(ip.tp.G + off, ld, tid, g);
  V::store(vb + static_cast<long long>(TV_CUR_G) * ld, ld, tid, g);
  double* mass_row = ip.mass + off;
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
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      if (2 * (tid + k * T) + v >= p.D) mass[k][v] = 1.0;
    }
  }
  V::store(mass_row, ld, tid, mass);
  double* est_row = p.est + static_cast<long long>(chain) * 4 * ld;
  double zero[K][2], sd[K][2], ss[K][2], rho[K][2];
  double kin = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int j = tid + k * T;
    double z0 = 0.0, z1 = 0.0;
    if (2 * j < p.D) {
      philox_normal_pair(p.seed, gchain, 0u, kKindStepInit, j, z0, z1);
      if (2 * j + 1 >= p.D) z1 = 0.0;
    }
    const double z[2] = {z0, z1};
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      zero[k][v] = 0.0;
      sd[k][v] = p.mass_init_count * (1.0 / mass[k][v]);
      ss[k][v] = p.mass_init_count * mass[k][v];
      rho[k][v] = z[v] * sqrt(mass[k][v]);
      kin = madd(1.0 / mass[k][v], rho[k][v] * rho[k][v], kin);
    }
  }
  V::store(est_row + 0 * ld, ld, tid, zero);
  V::store(est_row + 1 * ld, ld, tid, sd);
  V::store(est_row + 2 * ld, ld, tid, zero);
  V::store(est_row + 3 * ld, ld, tid, ss);
  V::store(vb + static_cast<long long>(TV_RHO) * ld, ld, tid, rho);
  double r[1] = {kin};
  grp.sum(r);
  if (tid == 0) {
    const double step = ip.have_steps ? ip.steps[chain] : ip.step_init;
    ChainScalars sc{};
    sc.adam_x = log(step);
    sc.adam_b1p = 1.0; sc.adam_b2p = 1.0;
    sc.mm_total = 2.0; sc.mm_count = 1.0;
    sc.est_w = p.mass_init_count;
    sc.step = step;
    sc.min_micro = p.min_micro_cfg;
    p.sc[chain] = sc;
    TickState st{};
    st.pc = PC_START_TRANSITION;
    st.lp_cur = ip.tp.LP[chain];
    st.Hs = st.lp_cur + (-0.5 * r[0]);   // joint at the start of the step search
    st.step = step;
    st.rung = 0;                          // search phase: 0 doubling, 1 shrinking
    st.reversing = ip.have_steps ? 1 : 0; // search finished?
    ip.tp.ts[chain] = st;
  }
}

// step search, post: theta* = theta + s (M^-1 (rho + s/2 g))  (util.hpp:250-252)
template <int T, int K, int CTA>
__global__ void __launch_bounds__(CTA) tick_search_post_kernel(const TickInitParams ip) {
  __shared__ double red_smem[group_smem_doubles<T>()];
  Group<T> grp;
  const int chain = group_setup<T, CTA>(grp, red_smem);
  const ChainParams& p = ip.tp.cp;
  if (chain >= p.C) return;
  using V = Vec<T, K>;
  const int tid = grp.tid, ld = p.ld;
  const TickState& st = ip.tp.ts[chain];
  if (st.reversing) return;
  const long long off = static_cast<long long>(chain) * ld;
  double* vb = ip.tp.vecs + static_cast<long long>(chain) * ip.tp.vec_stride;
  double th[K][2], g[K][2], rho[K][2], mass[K][2];
  V::load(vb + static_cast<long long>(TV_CUR) * ld, ld, tid, th);
  V::load(vb + static_cast<long long>(TV_CUR_G) * ld, ld, tid, g);
  V::load(vb + static_cast<long long>(TV_RHO) * ld, ld, tid, rho);
  V::load(ip.mass + off, ld, tid, mass);
  unit_padding<T, K>(mass, tid, p.D);
  const double s = st.step, hs = 0.5 * s;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const double rs = rho[k][v] + hs * g[k][v];
      th[k][v] = th[k][v] + s * ((1.0 / mass[k][v]) * rs);
    }
  }
  V::store(ip.tp.TH + off, ld, tid, th);
}

// step search, update: leapfrog_error and the two while loops of util.hpp:294-301
template <int T, int K, int CTA>
__global__ void __launch_bounds__(CTA) tick_search_update_kernel(const TickInitParams ip) {
  __shared__ double red_smem[group_smem_doubles<T>()];
  Group<T> grp;
  const int chain = group_setup<T, CTA>(grp, red_smem);
  const ChainParams& p = ip.tp.cp;
  if (chain >= p.C) return;
  using V = Vec<T, K>;
  const int tid = grp.tid, ld = p.ld;
  TickState st = ip.tp.ts[chain];
  if (st.reversing) return;
  const long long off = static_cast<long long>(chain) * ld;
  double* vb = ip.tp.vecs + static_cast<long long>(chain) * ip.tp.vec_stride;
  double g0[K][2], g1[K][2], rho[K][2], mass[K][2];
  V::load(vb + static_cast<long long>(TV_CUR_G) * ld, ld, tid, g0);
  V::load(ip.tp.G + off, ld, tid, g1);
  V::load(vb + static_cast<long long>(TV_RHO) * ld, ld, tid, rho);
  V::load(ip.mass + off, ld, tid, mass);
  unit_padding<T, K>(mass, tid, p.D);
  const double hs = 0.5 * st.step;
  double kin = 0.0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const double rs = (rho[k][v] + hs * g0[k][v]) + hs * g1[k][v];
      kin = madd(1.0 / mass[k][v], rs * rs, kin);
    }
  }
  double r[1] = {kin};
  grp.sum(r);
  const double err = (ip.tp.LP[chain] + (-0.5 * r[0])) - st.Hs;
  const double log09 = log(0.9), log06 = log(0.6);
  if (st.rung == 0) {
    if (err > log09 && st.sctr < 200) {
      st.step *= 2;
      st.sctr += 1;
    } else {
      st.rung = 1;
      st.sctr = 0;
    }
  }
  if (st.rung == 1) {
    if (err < log06 && st.sctr < 400) {
      st.step *= sqrt(0.5);
      st.sctr += 1;
    } else {
      st.reversing = 1;
    }
  }
  if (tid == 0) {
    ip.tp.ts[chain] = st;
    if (st.reversing) {
      ChainScalars& sc = p.sc[chain];
      sc.adam_x = log(st.step);
      sc.step = st.step;
      ip.steps[chain] = st.step;
    } else {
      atomicAdd(ip.tp.active_count, 1);
    }
  }
}

// ---------------------------------------------------------------------------
static TickParams tick_params(wb200_session& s, int n_iter, int adapt, bool store) {
  TickEngine& e = *s.tick;
  TickParams tp{};
  tp.cp = s.params(n_iter, adapt, store);
  tp.TH = e.TH.ptr; tp.G = e.G.ptr; tp.LP = e.LP.ptr;
  tp.vecs = e.vecs.ptr; tp.vec_stride = e.vec_stride;
  tp.ts = e.ts.ptr; tp.active_count = e.active.ptr;
  tp.st_logW = e.st_logW.ptr; tp.st_lp = e.st_lp.ptr;
  tp.chain_begin = 0; tp.chain_count = s.C;
  return tp;
}

#define WB200_EW_GRAD(T_, K_, CTA_)                                                     \
  do {                                                                                  \
    if (s.kind == kStdNormal)                                                           \
      elementwise_grad_kernel<StdNormalTarget, T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(tp); \
    else if (s.kind == kDiagGaussian)                                                   \
      elementwise_grad_kernel<DiagGaussianTarget, T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(tp); \
    else                                                                                \
      elementwise_grad_kernel<FunnelTarget, T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(tp); \
  } while (0)

// columns D .. ld of the gradient rows are padding the kernels expect to be zero
__global__ void zero_padding_kernel(double* G, int C, int D, int ld) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    for (int d = D; d < ld; ++d) G[static_cast<long long>(c) * ld + d] = 0.0;
  }
}

// NoExceptLogpGrad (util.hpp:336-346) for a whole batch: every chain of the tick sees
// logp = -inf and a zero gradient
__global__ void failed_batch_kernel(double* G, double* LP, int C, int ld) {
  const long long n = static_cast<long long>(C) * ld;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    G[i] = 0.0;
    if (i < C) LP[i] = -INFINITY;
  }
}

// `sampling` = inside a transition, where the reference wraps the density in
// NoExceptLogpGrad (adaptive_walnuts.hpp:205-215, walnuts.hpp:637-650); the
// initialisation (config.hpp:360-382, :469-476) calls it bare, so a failure there ends
// the run (walnutpy.cpp:162-170: "logp failed with code N").
static void tick_gradient(wb200_session& s, const TickParams& tp, bool sampling) {
  TickEngine& e = *s.tick;
  if (s.kind == kLogistic) {
    for (int h = 0; h < 2 && e.logistic[h]; ++h) {
      const long long off = static_cast<long long>(e.half_begin[h]) * s.ld;
      e.logistic[h]->evaluate(e.TH.ptr + off, e.G.ptr + off, e.LP.ptr + e.half_begin[h],
                              s.stream);
      s.launches += e.logistic[h]->kernels_per_eval();
    }
  } else if (s.kind == kBatchCallback) {
    const int rc = e.batch_fn(static_cast<size_t>(s.C), static_cast<size_t>(s.D),
                              static_cast<size_t>(s.ld), e.TH.ptr, e.G.ptr, e.LP.ptr,
                              static_cast<void*>(s.stream), e.batch_data);
    if (rc != 0) {
      if (!sampling) {
        throw std::runtime_error("logp failed with code " + std::to_string(rc));
      }
      const long long n = static_cast<long long>(s.C) * s.ld;
      failed_batch_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8)),
                            256, 0, s.stream>>>(e.G.ptr, e.LP.ptr, s.C, s.ld);
      WB200_CUDA(cudaGetLastError());
      s.logp_exceptions += 1;
      if (s.exception_log.size() < 64) s.exception_log.push_back({e.ticks, rc});
    }
    if (s.ld != s.D) {
      zero_padding_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(e.G.ptr, s.C, s.D, s.ld);
      WB200_CUDA(cudaGetLastError());
    }
  } else {
    const int grid = (s.C + s.shape.chains_per_cta - 1) / s.shape.chains_per_cta;
    WB200_TICK_SHAPE(s.shape, WB200_EW_GRAD);
    WB200_CUDA(cudaGetLastError());
    s.launches += 1;
  }
  e.grad_batches += 1;
}

static int read_active(wb200_session& s) {
  TickEngine& e = *s.tick;
  WB200_CUDA(cudaMemcpyAsync(e.active_host, e.active.ptr, sizeof(int), cudaMemcpyDeviceToHost,
                             s.stream));
  WB200_CUDA(cudaStreamSynchronize(s.stream));
  return *e.active_host;
}

void tick_create(wb200_session& s, const WalnutModelDesc& model) {
  s.tick = new TickEngine();
  TickEngine& e = *s.tick;
  const size_t CL = static_cast<size_t>(s.C) * s.ld;
  const int nvec = tick_vectors(s.tuning.max_trajectory_doublings);
  e.vec_stride = static_cast<long long>(nvec) * s.ld;
  e.TH.alloc(CL); e.G.alloc(CL); e.LP.alloc(s.C);
  e.vecs.alloc(static_cast<size_t>(e.vec_stride) * s.C);
  e.ts.alloc(s.C);
  e.st_logW.alloc(static_cast<size_t>(s.C) * kMaxDepth);
  e.st_lp.alloc(static_cast<size_t>(s.C) * kMaxDepth);
  e.active.alloc(1);
  WB200_CUDA(cudaMallocHost(&e.active_host, sizeof(int)));
  WB200_CUDA(cudaMemsetAsync(e.TH.ptr, 0, CL * 8, s.stream));
  WB200_CUDA(cudaMemsetAsync(e.G.ptr, 0, CL * 8, s.stream));
  WB200_CUDA(cudaMemsetAsync(e.vecs.ptr, 0, e.vecs.count * 8, s.stream));
  WB200_CUDA(cudaMemsetAsync(e.ts.ptr, 0, s.C * sizeof(TickState), s.stream));
  if (s.kind == kBatchCallback) {
    if (!model.data0) throw std::invalid_argument("kind 4 needs data0 = the density function");
    e.batch_fn = reinterpret_cast<WB200_BATCH_LOGP_GRAD>(const_cast<void*>(model.data0));
    e.batch_data = const_cast<void*>(model.data1);
  }
  if (s.kind == kLogistic) {
    if (!model.data0 || !model.data1 || model.N < 1) {
      throw std::invalid_argument("logistic needs data0 = X[N][D], data1 = y[N], N >= 1");
    }
    // WB200_TICK_PIPELINE=1: two halves (multiples of the 128-chain GEMM tile), each on
    // its own stream.  Measured at c4: no gain -- the step is power
    // capped, GEMM 1 fills the register file so only GEMM 2 can share an SM with the tick
    // kernel, and the half batches need split-K in GEMM 2 -- so one batch is the default.
    const char* env = std::getenv("WB200_TICK_PIPELINE");
    const bool split = s.C >= 2048 && env && std::string(env) == "1";
    const int first = split ? std::min(s.C, ((s.C / 2 + 127) / 128) * 128) : s.C;
    e.half_begin[0] = 0; e.half_count[0] = first;
    e.half_begin[1] = first; e.half_count[1] = s.C - first;
    e.logistic[0] = new LogisticGrad(static_cast<const double*>(model.data0),
                                     static_cast<const double*>(model.data1), model.N, s.D,
                                     first, s.ld, s.stream);
    if (e.half_count[1] > 0) {
      e.logistic[1] = new LogisticGrad(*e.logistic[0], e.half_count[1], s.stream);
      for (int h = 0; h < 2; ++h) {
        WB200_CUDA(cudaStreamCreateWithFlags(&e.half_stream[h], cudaStreamNonBlocking));
        WB200_CUDA(cudaEventCreateWithFlags(&e.join_event[h], cudaEventDisableTiming));
      }
      WB200_CUDA(cudaEventCreateWithFlags(&e.fork_event, cudaEventDisableTiming));
    }
  }
}

void tick_destroy(wb200_session& s) {
  delete s.tick;
  s.tick = nullptr;
}

#define WB200_TICK_INIT_POS(T_, K_, CTA_) \
  tick_init_positions_kernel<T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(ip)
#define WB200_TICK_INIT_STATE(T_, K_, CTA_) \
  tick_init_state_kernel<T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(ip)
#define WB200_TICK_SEARCH_POST(T_, K_, CTA_) \
  tick_search_post_kernel<T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(ip)
#define WB200_TICK_SEARCH_UPDATE(T_, K_, CTA_) \
  tick_search_update_kernel<T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(ip)
#define WB200_TICK_RUN(T_, K_, CTA_) \
  walnuts_tick_kernel<T_, K_, CTA_><<<grid, CTA_, 0, s.stream>>>(tp)

void tick_init(wb200_session& s, bool have_mass, bool have_steps, bool have_positions,
               double init_radius) {
  TickEngine& e = *s.tick;
  TickInitParams ip{};
  ip.tp = tick_params(s, 0, 1, false);
  ip.have_positions = have_positions;
  ip.have_mass = have_mass;
  ip.have_steps = have_steps;
  ip.init_radius = init_radius;
  ip.smoothing = s.tuning.mass_additive_smoothing;
  ip.step_init = s.tuning.step_size_init;
  ip.mass = s.inv_mass.ptr;  // masses staged in the inv_mass buffer until freeze
  ip.steps = s.red.ptr;
  const int grid = (s.C + s.shape.chains_per_cta - 1) / s.shape.chains_per_cta;
  WB200_TICK_SHAPE(s.shape, WB200_TICK_INIT_POS);
  WB200_CUDA(cudaGetLastError());
  tick_gradient(s, ip.tp, false);
  WB200_TICK_SHAPE(s.shape, WB200_TICK_INIT_STATE);
  WB200_CUDA(cudaGetLastError());
  s.launches += 2;
  if (!have_steps) {
    for (int it = 0; it < 700; ++it) {
      WB200_CUDA(cudaMemsetAsync(e.active.ptr, 0, sizeof(int), s.stream));
      WB200_TICK_SHAPE(s.shape, WB200_TICK_SEARCH_POST);
      tick_gradient(s, ip.tp, false);
      WB200_TICK_SHAPE(s.shape, WB200_TICK_SEARCH_UPDATE);
      WB200_CUDA(cudaGetLastError());
      s.launches += 2;
      if (read_active(s) == 0) break;
    }
  }
}

void tick_run(wb200_session& s, int n_iter, int adapt, bool store) {
  TickEngine& e = *s.tick;
  TickParams tp = tick_params(s, n_iter, adapt, store);
  const int grid = (s.C + s.shape.chains_per_cta - 1) / s.shape.chains_per_cta;
  tick_begin_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(e.ts.ptr, s.C);
  WB200_CUDA(cudaEventRecord(s.ev0, s.stream));
  while (true) {
    WB200_CUDA(cudaMemsetAsync(e.active.ptr, 0, sizeof(int), s.stream));
    WB200_TICK_SHAPE(s.shape, WB200_TICK_RUN);
    WB200_CUDA(cudaGetLastError());
    s.launches += 1;
    e.ticks += 1;
    if (read_active(s) == 0) break;
    tick_gradient(s, tp, true);
  }
  WB200_CUDA(cudaEventRecord(s.ev1, s.stream));
  if (store) s.rows_written += n_iter;
}

// free-running: chains that finished an iteration quota start over; nothing is reset
// (a chain that has done the phase's iter_cap iterations stays stopped)
__global__ void tick_resume_kernel(TickState* ts, const ChainScalars* sc, int C,
                                   long long rows_floor, int adapt, long long iter_cap) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const long long done = adapt ? static_cast<long long>(sc[c].warm_iter)
                                 : static_cast<long long>(sc[c].lp_n);
    if (ts[c].pc == PC_DONE && !(iter_cap > 0 && done >= iter_cap)) {
      ts[c].pc = PC_START_TRANSITION;
    }
    if (ts[c].rows < rows_floor) ts[c].rows = rows_floor;
  }
}

__global__ void tick_rows_kernel(const TickState* ts, int C, long long* rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) rows[c] = ts[c].rows;
}

// exactly n_ticks lock-step ticks; every chain does as many transitions as fit
// (ragged draw counts, like the reference's adaptive runs)
void tick_run_ticks(wb200_session& s, int n_ticks, int adapt, bool store,
                    long long iter_cap) {
  TickEngine& e = *s.tick;
  TickParams tp = tick_params(s, -1, adapt, store);
  tp.cp.iter_cap = iter_cap;
  const int grid = (s.C + s.shape.chains_per_cta - 1) / s.shape.chains_per_cta;
  tick_resume_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(
      e.ts.ptr, s.sc.ptr, s.C, s.rows_written, adapt, iter_cap);
  WB200_CUDA(cudaEventRecord(s.ev0, s.stream));
  if (s.kind == kLogistic && e.logistic[1]) {
    // two half batches, each a self-contained tick / gradient chain on its own stream
    WB200_CUDA(cudaEventRecord(e.fork_event, s.stream));
    for (int h = 0; h < 2; ++h) {
      WB200_CUDA(cudaStreamWaitEvent(e.half_stream[h], e.fork_event, 0));
    }
    for (int t = 0; t < n_ticks; ++t) {
      for (int h = 0; h < 2; ++h) {
        TickParams th = tp;
        th.chain_begin = e.half_begin[h];
        th.chain_count = e.half_count[h];
        const int hgrid = (th.chain_count + s.shape.chains_per_cta - 1) / s.shape.chains_per_cta;
#define WB200_TICK_RUN_HALF(T_, K_, CTA_) \
  walnuts_tick_kernel<T_, K_, CTA_><<<hgrid, CTA_, 0, e.half_stream[h]>>>(th)
        WB200_TICK_SHAPE(s.shape, WB200_TICK_RUN_HALF);
        WB200_CUDA(cudaGetLastError());
        const long long off = static_cast<long long>(th.chain_begin) * s.ld;
        e.logistic[h]->evaluate(e.TH.ptr + off, e.G.ptr + off, e.LP.ptr + th.chain_begin,
                                e.half_stream[h]);
        s.launches += 1 + e.logistic[h]->kernels_per_eval();
      }
      e.ticks += 1;
      e.grad_batches += 1;
    }
    for (int h = 0; h < 2; ++h) {
      WB200_CUDA(cudaEventRecord(e.join_event[h], e.half_stream[h]));
      WB200_CUDA(cudaStreamWaitEvent(s.stream, e.join_event[h], 0));
    }
  } else {
    for (int t = 0; t < n_ticks; ++t) {
      WB200_TICK_SHAPE(s.shape, WB200_TICK_RUN);
      WB200_CUDA(cudaGetLastError());
      tick_gradient(s, tp, true);
      s.launches += 1;
      e.ticks += 1;
    }
  }
  WB200_CUDA(cudaEventRecord(s.ev1, s.stream));
}

// freeze after a free-running warm-up: the transition a chain has in flight was
// started under the adapting step / metric and is abandoned; the chain restarts from
// its last completed draw (TV_CUR, its gradient and log density are only written at
// the end of a transition)
__global__ void tick_abort_kernel(TickState* ts, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) ts[c].pc = PC_DONE;
}

void tick_abort_inflight(wb200_session& s) {
  tick_abort_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(s.tick->ts.ptr, s.C);
  WB200_CUDA(cudaGetLastError());
}

void tick_chain_rows(wb200_session& s, long long* rows_host) {
  TickEngine& e = *s.tick;
  DeviceBuffer<long long> d;
  d.alloc(s.C);
  tick_rows_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(e.ts.ptr, s.C, d.ptr);
  WB200_CUDA(cudaMemcpyAsync(rows_host, d.ptr, s.C * sizeof(long long),
                             cudaMemcpyDeviceToHost, s.stream));
  WB200_CUDA(cudaStreamSynchronize(s.stream));
}

// rows[c] = draws chain c has staged since the last call; its counter restarts at 0
__global__ void tick_take_rows_kernel(TickState* ts, int C, long long* rows) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    rows[c] = ts[c].rows;
    ts[c].rows = 0;
  }
}

void tick_take_rows(wb200_session& s, long long* rows_dev) {
  tick_take_rows_kernel<<<(s.C + 255) / 256, 256, 0, s.stream>>>(s.tick->ts.ptr, s.C, rows_dev);
  WB200_CUDA(cudaGetLastError());
}

unsigned long long tick_count(const wb200_session& s) { return s.tick ? s.tick->ticks : 0; }

}  // namespace wb200
