// Chain-resident WALNUTS transition kernel for sm_100a.
//
// One group of T threads (a warp for D <= 128, a whole CTA above) owns one
// Markov chain for a whole launch: the live integrator state (theta, rho,
// grad), the macro-step start state, the inverse mass and the target's
// parameters stay in REGISTERS across every leapfrog micro-step, the
// within-orbit step-halving ladder, the reversibility ladder and the iterative
// orbit doubling.  Only the tree bookkeeping of sub-trees with two or more
// leaves (span ends, sub-tree first state, selected state) goes through a
// per-slot scratch area that is sized to live in the 126 MB L2; single leaves
// are merged straight from registers.  Chains are handed out by an atomic
// ticket so long and short orbits balance across the 148 SMs; there is no lane
// divergence between chains because a chain never shares a warp.
//
// Scalar decision arithmetic (Philox draws, log, log-sum-exp, Adam) runs on the
// group's first warp only and is broadcast through shared memory, so the other
// warps' issue slots go to the chains that share the SM.
//
// Reference semantics (all file:line relative to /root/reference/include/walnutpie):
//   transition_w walnuts.hpp:520-563 | build_span (recursion -> binary-counter
//   carry loop) :464-495 | build_leaf :420-442 | macro_step
//   :307-345 | reversible :254-279 | within_tolerance :218-235 | uturn :192-201
//   | combine :368-387 | AdaptiveWalnuts::operator() adaptive_walnuts.hpp:234-251
//   | MassEstimator :25-105 | MinMicroStepsAdaptHandler :119-164 | Adam
//   adam.hpp:70-93 | OnlineMoments::observe online_moments.hpp:185-191 |
//   WelfordAccumulator :34-70 (sampler.hpp:87-92).
//
// Arithmetic is written with explicit __dmul_rn/__dadd_rn in the order of the
// reference expressions, so element-wise results equal the CPU oracle's bit for
// bit; only cross-element sums (logp, kinetic energy, U-turn dots) differ, by
// summation order.
#pragma once
#if defined(__CUDACC_RTC__)
// run-time compilation of a user density (user_density.cu): no host headers
using uint32_t = unsigned int;
using uint64_t = unsigned long long;
#else
#include <cstdint>
#include <cuda_runtime.h>
#endif

#if !defined(__CUDACC__)
// g++ build of the very same state machine for tests/host_emu (one emulated
// thread owns the whole vector: T = 1); never part of the shipped library
#include "host_shims.hpp"
#endif

#include "philox.cuh"

namespace wb200 {

constexpr int kMaxDepth = 12;  // max_trajectory_doublings supported on device

// Arithmetic policy of the element-wise hot loops.  The reference is plain C++ whose
// products and sums round separately on its x86-64 baseline build and contract to fused
// multiply-adds where the compiler may (-march=native, aarch64).  "Fused" (the shipped
// default) contracts exactly the accumulate sites below -- a*b + c with one rounding:
// the two momentum kicks and the position drift of a leapfrog step (walnuts.hpp:329-332),
// the target's sum of squares, the kinetic energy (util.hpp:220-223), the U-turn dots
// (walnuts.hpp:192-201) and the discounted sum of squares of OnlineMoments
// (online_moments.hpp:189) -- 6 instead of 10 fp64 instructions per element and leapfrog
// step.  -DWB200_EXACT_ARITH keeps every rounding separate (bit-identical element-wise
// results to the reference's baseline build); the oracle implements both policies and
// tests/host_emu pins this source to it bit for bit under each.
#ifdef WB200_EXACT_ARITH
constexpr bool kFusedArith = false;
#else
constexpr bool kFusedArith = true;
#endif
__device__ __forceinline__ double madd(double a, double b, double c) {
  if constexpr (kFusedArith) return __fma_rn(a, b, c);
  return __dadd_rn(__dmul_rn(a, b), c);
}
// The fused policy also spares the estimators two of their fp64 divisions per element
// (35-60 instructions each): the two discounted Welford estimators of MassEstimator share
// one weight (adaptive_walnuts.hpp:54-80), so it cancels in var_draws / var_scores
// (:89-94), and the mean update divides by that scalar weight (online_moments.hpp:187),
// i.e. multiplies by a reciprocal taken once per transition.
// fp32 mode (Real = float): the integrator state and its element-wise arithmetic are
// single precision; energies, U-turn dots (beyond a thread's own <= 16 terms), every
// scalar decision and the adaptation statistics stay fp64.
__device__ __forceinline__ float madd(float a, float b, float c) {
  if constexpr (kFusedArith) return __fmaf_rn(a, b, c);
  return __fadd_rn(__fmul_rn(a, b), c);
}
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ double metric_from_sums(double S_draw, double S_score, double w) {
  if constexpr (kFusedArith) return sqrt(S_draw / S_score);
  return sqrt((S_draw / w) / (S_score / w));
}

enum TargetKind : int { kStdNormal = 0, kDiagGaussian = 1, kFunnel = 2,
                        kLogistic = 3, kBatchCallback = 4, kDeviceSource = 5 };

// persistent per-chain scalars (SoA would not help: one group reads one record)
struct ChainScalars {
  // Adam on log step (adam.hpp:48-66)
  double adam_x, adam_m, adam_v, adam_t, adam_b1p, adam_b2p;
  // min-micro controller (adaptive_walnuts.hpp:127-132)
  double mm_total, mm_count;
  // shared weight of the two discounted Welford estimators
  double est_w;
  // Welford of lp for the sampling R-hat (online_moments.hpp:34-70)
  double lp_mean, lp_m2;
  // frozen tuning used by sampling (adaptive_walnuts.hpp:263-271)
  double step;
  double last_lp;
  unsigned long long lp_n;
  unsigned long long grad_evals;
  unsigned long long macro_steps;   // leaves attempted
  unsigned long long rung_sum;      // sum of accepted rung indices (diagnostic)
  // free-running launches (ChainParams::eval_budget): evaluations spent beyond the budgets
  // handed out so far -- the transition that exhausts a budget is finished, and its excess
  // is taken off the next budget, so every chain's long-run evaluation rate is the same
  long long eval_debt;
  unsigned int iter;       // global transition index (Philox addressing)
  unsigned int warm_iter;  // AdaptiveWalnuts::iteration_
  int min_micro;
  int last_depth;
};

// A launch's limits for the chain a slot is working on: worked out once per chain by thread 0
// and read from shared memory where they are needed (registers are what the hot loop is
// short of).
struct RunLimits {  // free-running launches only (ChainRunner::advance)
  long long rows;    // draw row of the next transition
  long long budget;  // gradient evaluations within which transitions may start
  long long cap;     // transitions at most in this launch
  long long done;    // transitions completed in this launch
  long long spent;   // gradient evaluations used in this launch
  unsigned long long evals0;  // the chain's evaluation count when the launch took it up
  int keep_metric;   // run() is re-entered for the same chain: its Cholesky factor stands
};

struct ChainParams {
  int C, D, ld;
  int n_iter;
  int adapt;
  // Free-running mode (0: every chain does exactly n_iter transitions).  The reference's
  // chains are threads that each run at their own pace until the controller stops them
  // (adapt.hpp:110-129, sampler.hpp:79-94), so a chain with long orbits completes fewer
  // iterations than one with short orbits.  The device analogue of "equal time" is equal
  // WORK: every chain gets eval_budget gradient evaluations per launch and completes the
  // transitions that fit (at most free_cap of them, and never beyond iter_cap transitions
  // of the current phase in total); per-chain draw rows are counted in `rows`.  A launch
  // no longer waits for the chain with the longest orbits.  n_iter is 1 in this mode.
  long long eval_budget;
  long long free_cap;
  long long iter_cap;
  long long* rows;      // nullable [C]: next draw row of each chain (else draw_base + it)
  int max_depth, max_halvings, min_micro_cfg;
  double max_error;
  double mass_init_count, macro_target;
  double adam_target, adam_lr, adam_b1, adam_b2, adam_eps, adam_decay;
  // adam_lr / pow(t, adam_decay) for t = 1 .. adam_tab_n: the decayed learning rate
  // depends on the update count only, and pow is the longest link of the scalar chain
  // every macro step of the adaptive phase waits for (null: computed on the spot)
  const double* adam_tab;
  int adam_tab_n;
  uint32_t seed, chain_offset;
  double* theta;        // [C][ld]
  double* inv_mass;     // [C][ld]  fixed metric used when adapt == 0
  double* est;          // [C][4][ld]  mu_draw, S_draw, mu_score, S_score
  ChainScalars* sc;     // [C]
  double* draws;        // nullable [C][draw_cap][ld]
  long long draw_cap;
  long long draw_base;  // first row written by this launch
  double* lp_out;       // nullable [C][draw_cap]
  int* depth_out;       // nullable [C][draw_cap]
  double* step_out;     // nullable [C][draw_cap]  step AFTER the iteration (on_warmup)
  double* im_out;       // nullable [C][draw_cap][ld]  metric used (on_warmup)
  double* scratch;      // [slots][nvec][ld]
  long long scratch_stride;
  unsigned int* ticket;
  // ticket -> chain: chains in descending order of the work their previous launch took
  // (longest-processing-time-first, engine.cu); null: identity
  const int* order;
  const double* tparam; // target parameters (e.g. precision[ld])
};

// scratch vector indices
enum : int { A_TH_BK = 0, A_RHO_BK, A_G_BK, A_TH_FW, A_RHO_FW, A_G_FW, A_SEL,
             E_TH, E_RHO, E_G, A_IM, ST_BASE };
enum : int { ST_THF = 0, ST_RHOF = 1, ST_SEL = 2 };
__host__ __device__ inline int scratch_vectors(int max_depth) {
  return ST_BASE + 3 * max_depth;
}
// doubles of scratch per resident group: the vectors, then the Adam queue (one |dH| per
// macro step of a transition, at most 2^max_depth of them)
__host__ __device__ inline long long scratch_doubles(int max_depth, int ld) {
  return static_cast<long long>(scratch_vectors(max_depth)) * ld + (1ll << max_depth);
}

// Shared memory of one resident chain: the macro-step start state (theta, rho, grad) --
// the previous leaf, which the halving ladder restarts from and odd leaves take their
// U-turn dots against -- and the (logW, lp) stack of finished sub-trees.  Parking the
// start state here instead of in a second set of registers removes the register copies
// between the two roles (a third of the executed instructions of the round-1 kernel
// were MOVs) and a third of the register footprint.
__host__ __device__ inline int chain_smem_doubles(int ld) { return 3 * ld + 2 * kMaxDepth; }

// ---------------------------------------------------------------------------
// T cooperating threads.  sum(): all-reduce, bitwise identical in every thread.
// bcast(): values computed by the control warp reach every thread.
constexpr int kRedStride = 4;  // doubles per warp row / broadcast row

#if defined(__CUDACC__)
// One-warp groups (D <= 128: four independent chains per CTA, sixteen instruction streams
// per SM) keep ONE out-of-line copy of each butterfly: the kernel's executed footprint is
// what its warps fight over in the 32 KB instruction cache (ncu: 57 % of the D = 100
// kernel's stall samples were instruction fetches with the butterflies inlined at every
// reduction site).  Same operations in the same order as the inline code below.
struct Sum2 { double a, b; };
struct Sum4 { double a, b, c, d; };
__device__ __noinline__ inline Sum2 warp_allsum2(double v0, double v1, int lane) {
  const bool hi16 = (lane & 16) != 0;
  double a = (hi16 ? v1 : v0) + __shfl_xor_sync(0xffffffffu, hi16 ? v0 : v1, 16);
  a += __shfl_xor_sync(0xffffffffu, a, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  return Sum2{__shfl_sync(0xffffffffu, a, 0), __shfl_sync(0xffffffffu, a, 16)};
}
__device__ __noinline__ inline Sum4 warp_allsum4(double v0, double v1, double v2, double v3,
                                                 int lane) {
  const bool hi16 = (lane & 16) != 0;
  double a = hi16 ? v2 : v0;
  double b = hi16 ? v3 : v1;
  a += __shfl_xor_sync(0xffffffffu, hi16 ? v0 : v2, 16);
  b += __shfl_xor_sync(0xffffffffu, hi16 ? v1 : v3, 16);
  const bool hi8 = (lane & 8) != 0;
  const double keep = hi8 ? b : a, give = hi8 ? a : b;
  a = keep + __shfl_xor_sync(0xffffffffu, give, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 4);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  return Sum4{__shfl_sync(0xffffffffu, a, 0), __shfl_sync(0xffffffffu, a, 8),
              __shfl_sync(0xffffffffu, a, 16), __shfl_sync(0xffffffffu, a, 24)};
}
#endif

template <int T>
struct Group {
  static constexpr int W = T / 32;
  int tid, lane, warp;
  double* red;  // shared, [2][W + 1][kRedStride] when W > 1
  int parity;

  __device__ __forceinline__ bool ctl() const { return W <= 1 || warp == 0; }

  __device__ __forceinline__ void sync() const {
    if constexpr (W > 1) {
      __syncthreads();
    } else {
      __syncwarp();
    }
  }

  // All-reduce of N partial sums.  The warp stage halves the number of VALUES a lane
  // carries at every butterfly step where it can (lanes with the xor bit set keep the upper
  // half of the values, the others the lower half), so N = 4 costs 7 shuffled doubles
  // instead of 20: after the steps, lane group l / (32 / N) holds the warp total of value
  // l / (32 / N).  Warps then meet in shared memory (W > 1) or the group leaders'
  // totals are broadcast by shuffle (one warp).
  template <int N>
  __device__ __forceinline__ void sum(double (&v)[N]) {
    static_assert(N <= kRedStride, "too many values");
#if defined(__CUDACC__)
    if constexpr (T == 32 && N == 2) {
      const Sum2 r = warp_allsum2(v[0], v[1], lane);
      v[0] = r.a; v[1] = r.b;
      return;
    }
    if constexpr (T == 32 && N == 4) {
      const Sum4 r = warp_allsum4(v[0], v[1], v[2], v[3], lane);
      v[0] = r.a; v[1] = r.b; v[2] = r.c; v[3] = r.d;
      return;
    }
#endif
    if constexpr (T >= 32 && (N == 2 || N == 4)) {
      double a = v[0], b = v[N > 2 ? 1 : 0];
      if constexpr (N == 4) {
        const bool hi16 = (lane & 16) != 0;
        a = hi16 ? v[2] : v[0];
        b = hi16 ? v[3] : v[1];
        a += __shfl_xor_sync(0xffffffffu, hi16 ? v[0] : v[2], 16);
        b += __shfl_xor_sync(0xffffffffu, hi16 ? v[1] : v[3], 16);
        const bool hi8 = (lane & 8) != 0;
        const double keep = hi8 ? b : a, give = hi8 ? a : b;
        a = keep + __shfl_xor_sync(0xffffffffu, give, 8);
      } else {
        const bool hi16 = (lane & 16) != 0;
        a = (hi16 ? v[1] : v[0]) + __shfl_xor_sync(0xffffffffu, hi16 ? v[0] : v[1], 16);
        a += __shfl_xor_sync(0xffffffffu, a, 8);
      }
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      constexpr int kGroup = 32 / N;  // lanes holding the total of one value
      if constexpr (W > 1) {
        double* buf = red + parity * ((W + 1) * kRedStride);
        if ((lane & (kGroup - 1)) == 0) buf[warp * kRedStride + lane / kGroup] = a;
        __syncthreads();
#pragma unroll
        for (int n = 0; n < N; ++n) {
          double s = buf[n];
#pragma unroll
          for (int w = 1; w < W; ++w) s += buf[w * kRedStride + n];
          v[n] = s;
        }
        parity ^= 1;
      } else {
#pragma unroll
        for (int n = 0; n < N; ++n) v[n] = __shfl_sync(0xffffffffu, a, n * kGroup);
      }
      return;
    }
    if constexpr (T >= 32) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
          v[n] += __shfl_xor_sync(0xffffffffu, v[n], m);
        }
      }
    }
    if constexpr (W > 1) {
      double* buf = red + parity * ((W + 1) * kRedStride);
      if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) buf[warp * kRedStride + n] = v[n];
      }
      __syncthreads();
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double s = buf[n];
#pragma unroll
        for (int w = 1; w < W; ++w) s += buf[w * kRedStride + n];
        v[n] = s;
      }
      parity ^= 1;
    }
  }

  // v[] as computed by the control warp becomes visible to every thread
  template <int N>
  __device__ __forceinline__ void bcast(double (&v)[N]) {
    static_assert(N <= kRedStride, "too many values");
    if constexpr (W > 1) {
      double* buf = red + parity * ((W + 1) * kRedStride) + W * kRedStride;
      if (warp == 0 && lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) buf[n] = v[n];
      }
      __syncthreads();
#pragma unroll
      for (int n = 0; n < N; ++n) v[n] = buf[n];
      parity ^= 1;
    }
  }
};

template <int T>
constexpr int group_smem_doubles() {
  return (T / 32 > 1) ? 2 * (T / 32 + 1) * kRedStride : 1;
}

// ---------------------------------------------------------------------------
// Rows are padded to the 2*T*K element slots of their group (row_stride()), so loads and
// stores need no bounds checks; padding holds theta = rho = grad = 0 and a unit metric.
template <class Real> struct PairOf;
template <> struct PairOf<double> { using type = double2; };
template <> struct PairOf<float> { using type = float2; };

template <int T, int K, class Real>
struct VecT {
  using Pair = typename PairOf<Real>::type;
  // element pair owned by this thread in chunk k: 2*(tid + k*T), +1
  __device__ __forceinline__ static void load(const Real* row, int, int tid,
                                              Real (&x)[K][2]) {
    const Pair* r2 = reinterpret_cast<const Pair*>(row) + tid;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const Pair v = r2[k * T];
      x[k][0] = v.x; x[k][1] = v.y;
    }
  }
  __device__ __forceinline__ static void store(Real* row, int, int tid,
                                               const Real (&x)[K][2]) {
    Pair* r2 = reinterpret_cast<Pair*>(row) + tid;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      Pair v;
      v.x = x[k][0]; v.y = x[k][1];
      r2[k * T] = v;
    }
  }
  // the same against fp64 rows (host-visible state is always fp64)
  __device__ __forceinline__ static void load64(const double* row, int tid, Real (&x)[K][2]) {
    const double2* r2 = reinterpret_cast<const double2*>(row) + tid;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 v = r2[k * T];
      x[k][0] = static_cast<Real>(v.x); x[k][1] = static_cast<Real>(v.y);
    }
  }
  __device__ __forceinline__ static void store64(double* row, int tid, const Real (&x)[K][2]) {
    double2* r2 = reinterpret_cast<double2*>(row) + tid;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      r2[k * T] = make_double2(static_cast<double>(x[k][0]), static_cast<double>(x[k][1]));
    }
  }
  __device__ __forceinline__ static void copy(Real (&dst)[K][2], const Real (&src)[K][2]) {
#pragma unroll
    for (int k = 0; k < K; ++k) { dst[k][0] = src[k][0]; dst[k][1] = src[k][1]; }
  }
};
template <int T, int K>
using Vec = VecT<T, K, double>;

// ---------------------------------------------------------------------------
// Targets.  grad() fills g and a per-thread partial `lp_part` whose group sum
// is logp (so Gaussians need no reduction inside a micro-step).
template <int T, int K, class Real>
struct StdNormalTargetT {  // examples/walnutpie_api.cpp:39-43
  __device__ __forceinline__ void init(const ChainParams&, int) {}
  __device__ __forceinline__ void grad(const Real (&th)[K][2], Real (&g)[K][2],
                                       Real& lp_part, Group<T>&) const {
    Real s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        s = madd(th[k][v], th[k][v], s);
        g[k][v] = -th[k][v];
      }
    }
    lp_part = static_cast<Real>(-0.5) * s;
  }
};

template <int T, int K, class Real>
struct DiagGaussianTargetT {  // generalises examples/examples.cpp:20-31
  Real prec[K][2];
  __device__ __forceinline__ void init(const ChainParams& p, int tid) {
    VecT<T, K, Real>::load64(p.tparam, tid, prec);
  }
  __device__ __forceinline__ void grad(const Real (&th)[K][2], Real (&g)[K][2],
                                       Real& lp_part, Group<T>&) const {
    Real s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        Real t = mul_rn(th[k][v], prec[k][v]);
        s = madd(th[k][v], t, s);
        g[k][v] = -t;
      }
    }
    lp_part = static_cast<Real>(-0.5) * s;
  }
};

template <int T, int K, class Real>
struct FunnelTargetT {
  double half_dm1;
  bool owner;  // owns element 0 (v)
  int room;    // D - 2 * tid: element slot 2 * k * T + v of this thread is real iff < room
  __device__ __forceinline__ void init(const ChainParams& p, int tid) {
    half_dm1 = 0.5 * static_cast<double>(p.D - 1);
    owner = (tid == 0);
    room = p.D - 2 * tid;
  }
  __device__ __forceinline__ void grad(const Real (&th)[K][2], Real (&g)[K][2],
                                       Real& lp_part, Group<T>& grp) const {
    Real ss = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        bool is_v = owner && k == 0 && v == 0;
        ss = is_v ? ss : madd(th[k][v], th[k][v], ss);
      }
    }
    double r[2] = {static_cast<double>(ss), owner ? static_cast<double>(th[0][0]) : 0.0};
    grp.sum(r);
    const double v0 = r[1];
    const double ev = exp(-v0);
    const double q = __dmul_rn(__dmul_rn(0.5, ev), r[0]);
    const Real ev_r = static_cast<Real>(ev);
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        // padding slots hold theta = 0: 0 * exp(-v0) must stay 0 when exp overflows
        g[k][v] = (2 * k * T + v < room) ? -mul_rn(th[k][v], ev_r) : static_cast<Real>(0);
      }
    }
    if (owner) {
      // the prior's 1/18 and 1/9 as rounded constants (this target is specified by the
      // north star, not the reference): two fp64 divisions were a tenth of the executed
      // instructions of the D = 100 kernel
      constexpr double kInv18 = 1.0 / 18.0, kInv9 = 1.0 / 9.0;
      double lp = __dadd_rn(__dadd_rn(-__dmul_rn(__dmul_rn(v0, v0), kInv18),
                                      -__dmul_rn(half_dm1, v0)), -q);
      g[0][0] = static_cast<Real>(__dadd_rn(__dadd_rn(-__dmul_rn(v0, kInv9), -half_dm1), q));
      lp_part = static_cast<Real>(lp);
    } else {
      lp_part = 0;
    }
  }
};
template <int T, int K> using StdNormalTarget = StdNormalTargetT<T, K, double>;
template <int T, int K> using DiagGaussianTarget = DiagGaussianTargetT<T, K, double>;
template <int T, int K> using FunnelTarget = FunnelTargetT<T, K, double>;
template <int T, int K> using StdNormalTargetF = StdNormalTargetT<T, K, float>;
template <int T, int K> using DiagGaussianTargetF = DiagGaussianTargetT<T, K, float>;
template <int T, int K> using FunnelTargetF = FunnelTargetT<T, K, float>;

// ---------------------------------------------------------------------------
__device__ __forceinline__ double log_sum_exp2(double x1, double x2) {
  // util.hpp:174-183.  exp(max - max) is exactly 1, so the reference's
  // exp(x1-m) + exp(x2-m) equals 1 + exp(min - max) bit for bit: one exp fewer.
  double m = fmax(x1, x2);
  if (isnan(x1) || isnan(x2)) return nan("");
  if (isinf(m) || isnan(x1 + x2)) return fmax(x1, x2);
  return m + log(1.0 + exp(fmin(x1, x2) - m));
}

// Scalar helpers are kept out of line: the transition kernel is large and the
// instruction cache is what its hot loop competes for.
__device__ __noinline__ inline void adam_update(ChainScalars& sc, const ChainParams& p,
                                         double alpha) {
  // adam.hpp:70-86
  sc.adam_t += 1.0;
  sc.adam_b1p *= p.adam_b1;
  sc.adam_b2p *= p.adam_b2;
  double grad = p.adam_target - alpha;
  sc.adam_m = p.adam_b1 * sc.adam_m + (1 - p.adam_b1) * grad;
  sc.adam_v = p.adam_b2 * sc.adam_v + (1 - p.adam_b2) * grad * grad;
  double m_hat = sc.adam_m / (1 - sc.adam_b1p);
  double v_hat = sc.adam_v / (1 - sc.adam_b2p);
  double decayed = (p.adam_tab != nullptr && sc.adam_t <= static_cast<double>(p.adam_tab_n))
                       ? p.adam_tab[static_cast<int>(sc.adam_t) - 1]
                       : p.adam_lr / pow(sc.adam_t, p.adam_decay);
  double denom = sqrt(v_hat) + p.adam_eps;
  sc.adam_x -= decayed * m_hat / denom;
}

// Adam feeds on one acceptance probability per macro step, but its state is only read at
// the start of the next transition.  The chain kernel therefore queues |dH| and lets the
// control warp work the queue off in one go: the expensive, mutually independent parts of
// every update (exp, three divisions, a square root, the decayed rate) run one update
// per lane, while the cheap linear recurrences (t, beta powers, m, v, x) run in order --
// the same IEEE operations on the same operands as adam_update, so the same bits.
struct AdamQueue {
  int n;  // entries waiting in the group's scratch queue
};

template <int LANES>
__device__ __noinline__ inline void adam_flush(ChainScalars& sc, const ChainParams& p,
                                        AdamQueue& q, const double* dH, int lane) {
  __syncwarp();
  const int n = q.n;
  if (n == 0) return;
  double t = sc.adam_t, b1p = sc.adam_b1p, b2p = sc.adam_b2p, m = sc.adam_m, v = sc.adam_v;
  double x = sc.adam_x;
  for (int base = 0; base < n; base += LANES) {
    const int cnt = n - base < LANES ? n - base : LANES;
    double alpha = 0.0;
    if (lane < cnt) alpha = exp(-dH[base + lane]);
    double my_t = 0.0, my_b1p = 0.0, my_b2p = 0.0, my_m = 0.0, my_v = 0.0;
    for (int j = 0; j < cnt; ++j) {
      const double a_j = LANES > 1 ? __shfl_sync(0xffffffffu, alpha, j) : alpha;
      t += 1.0;
      b1p *= p.adam_b1;
      b2p *= p.adam_b2;
      const double grad = p.adam_target - a_j;
      m = p.adam_b1 * m + (1 - p.adam_b1) * grad;
      v = p.adam_b2 * v + (1 - p.adam_b2) * grad * grad;
      if (j == lane) { my_t = t; my_b1p = b1p; my_b2p = b2p; my_m = m; my_v = v; }
    }
    double term = 0.0;
    if (lane < cnt) {
      const double m_hat = my_m / (1 - my_b1p);
      const double v_hat = my_v / (1 - my_b2p);
      const double decayed =
          (p.adam_tab != nullptr && my_t <= static_cast<double>(p.adam_tab_n))
              ? p.adam_tab[static_cast<int>(my_t) - 1]
              : p.adam_lr / pow(my_t, p.adam_decay);
      const double denom = sqrt(v_hat) + p.adam_eps;
      term = decayed * m_hat / denom;
    }
    for (int j = 0; j < cnt; ++j) {
      x -= LANES > 1 ? __shfl_sync(0xffffffffu, term, j) : term;
    }
  }
  __syncwarp();
  if (lane == 0) {
    sc.adam_t = t; sc.adam_b1p = b1p; sc.adam_b2p = b2p;
    sc.adam_m = m; sc.adam_v = v; sc.adam_x = x;
    q.n = 0;
  }
  __syncwarp();
}

__device__ __forceinline__ int min_micro_steps(double mm_total, double mm_count,
                                               const ChainParams& p) {
  // adaptive_walnuts.hpp:152-157
  double mean_micro = mm_total / mm_count;
  long long r = llround(mean_micro / p.macro_target);
  long long c = p.min_micro_cfg;
  return static_cast<int>(r > c ? r : c);
}
__device__ __forceinline__ int min_micro_steps(const ChainScalars& sc,
                                               const ChainParams& p) {
  return min_micro_steps(sc.mm_total, sc.mm_count, p);
}

// The scalar stream is stateless (Philox keyed by (chain, iteration, index)), so the
// control warp can evaluate log(u) for the next 32 decision indices at once, one per
// lane, instead of one Philox block + one log() on the critical path of every merge:
// every other warp of the group waits at a barrier for that scalar chain.
struct DecisionCache {
  double logu[32];
  uint32_t base, iter;  // covers indices [base, base + lanes) of iteration `iter`
  uint32_t valid;
};

// combine<U> (walnuts.hpp:368-387): logW of the union and whether the new span's
// selection wins; one uniform from the chain's scalar stream
struct MergeResult { double logW; bool take_new; };
__device__ __noinline__ inline MergeResult merge_scalar(bool metropolis, double logW_old,
                                                 double logW_new, uint32_t seed,
                                                 uint32_t gchain, uint32_t iter,
                                                 uint32_t index) {
  const double lw = log_sum_exp2(logW_old, logW_new);
  const double u = philox_uniform(seed, gchain, iter, index);
  const double denom = metropolis ? logW_old : lw;
  return MergeResult{lw, log(u) < logW_new - denom};
}

// the same with log(u) already known
__device__ __noinline__ inline MergeResult merge_scalar_logu(bool metropolis, double logW_old,
                                                      double logW_new, double log_u) {
  const double lw = log_sum_exp2(logW_old, logW_new);
  const double denom = metropolis ? logW_old : lw;
  return MergeResult{lw, log_u < logW_new - denom};
}

__device__ __noinline__ inline double log_uniform(uint32_t seed, uint32_t gchain, uint32_t iter,
                                           uint32_t index) {
  return log(philox_uniform(seed, gchain, iter, index));
}

__device__ __noinline__ inline bool direction_bit(uint32_t seed, uint32_t gchain, uint32_t iter,
                                           uint32_t index) {
  return philox_bit(seed, gchain, iter, index);
}

__device__ __noinline__ inline double2 momentum_normals(uint32_t seed, uint32_t gchain,
                                                 uint32_t iter, uint32_t j) {
  double z0, z1;
  philox_normal_pair(seed, gchain, iter, kKindNormal, j, z0, z1);
  return make_double2(z0, z1);
}

__device__ __noinline__ inline double exp_noinline(double x) { return exp(x); }
// fp64 division and square root expand to 35-60 instructions each; the per-transition
// code uses them 8 elements x several times per thread, so one shared body each keeps
// the kernel inside the instruction cache (same IEEE operations)
__device__ __noinline__ inline double div_noinline(double a, double b) { return a / b; }
__device__ __noinline__ inline double sqrt_noinline(double a) { return sqrt(a); }

// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// The per-transition adaptation blocks of the chain kernel, out of line and free of the
// runner's register arrays: inside the function that holds the macro-step loop they
// raised the register pressure enough to spill whole state arrays (ncu: 5 % of the
// executed instructions were local loads).  They communicate through the group's
// scratch row (the metric), the chain's estimator rows and the shared scalars.
//
// Start of a transition (adaptive_walnuts.hpp:235-245): M^-1 from the estimators ->
// scratch row `im_row`; thread 0 publishes step = exp(adam_x) and min-micro in `sc`.
// Returns the group's barrier parity.
template <int T, int K, class Real>
__device__ __noinline__ int adapt_begin(const ChainParams& p, Group<T> grp, ChainScalars& sc,
                                        const double* est_row, Real* im_row) {
  using V = Vec<T, K>;
  const int ld = p.ld, tid = grp.tid;
  grp.sync();  // thread 0's updates at the end of the previous transition are visible
  const double est_w = sc.est_w;
  double Sd[K][2], Ss[K][2], im[K][2];
  V::load(est_row + 1 * ld, ld, tid, Sd);
  V::load(est_row + 3 * ld, ld, tid, Ss);
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      // MassEstimator::inv_mass_estimate, adaptive_walnuts.hpp:89-94
      im[k][v] = kFusedArith
                     ? sqrt_noinline(div_noinline(Sd[k][v], Ss[k][v]))
                     : sqrt_noinline(div_noinline(div_noinline(Sd[k][v], est_w),
                                                  div_noinline(Ss[k][v], est_w)));
    }
  }
  if constexpr (sizeof(Real) == 8) {
    V::store(im_row, ld, tid, im);
  } else {
    Real im_r[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      im_r[k][0] = static_cast<Real>(im[k][0]);
      im_r[k][1] = static_cast<Real>(im[k][1]);
    }
    VecT<T, K, Real>::store(im_row, ld, tid, im_r);
  }
  if (tid == 0) {
    sc.step = exp(sc.adam_x);
    sc.min_micro = min_micro_steps(sc.mm_total, sc.mm_count, p);
  }
  grp.sync();
  return grp.parity;
}

// End of a transition (adaptive_walnuts.hpp:247-250): gradient at the selected draw,
// discounted Welford updates of draws and scores, estimator weight, min-micro controller.
template <class Target, int T, int K, class Real>
__device__ __noinline__ int adapt_end(const ChainParams& p, Group<T> grp, ChainScalars& sc,
                                      double* est_row, const Real* sel_row, int depth) {
  using V = Vec<T, K>;
  const int ld = p.ld, tid = grp.tid;
  Target tgt;
  tgt.init(p, tid);
  Real cur[K][2], gsel[K][2], lp_dummy;
  VecT<T, K, Real>::load(sel_row, ld, tid, cur);
  tgt.grad(cur, gsel, lp_dummy, grp);  // grad_select (cached by the reference)
  const double gamma = 1.0 - 1.0 / (p.mass_init_count + static_cast<double>(sc.warm_iter));
  const double est_w = gamma * sc.est_w + 1.0;
  const double r_w = 1.0 / est_w;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    double mu[K][2], S[K][2];
    V::load(est_row + (2 * e) * ld, ld, tid, mu);
    V::load(est_row + (2 * e + 1) * ld, ld, tid, S);
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const double y = static_cast<double>(e == 0 ? cur[k][v] : gsel[k][v]);
        // online_moments.hpp:185-191 (both factors see the updated mean)
        mu[k][v] = kFusedArith
                       ? madd(__dadd_rn(y, -mu[k][v]), r_w, mu[k][v])
                       : __dadd_rn(mu[k][v], div_noinline(__dadd_rn(y, -mu[k][v]), est_w));
        const double d = __dadd_rn(y, -mu[k][v]);
        S[k][v] = madd(gamma, S[k][v], __dmul_rn(d, d));
      }
    }
    V::store(est_row + (2 * e) * ld, ld, tid, mu);
    V::store(est_row + (2 * e + 1) * ld, ld, tid, S);
  }
  grp.sync();  // every thread has read the scalars thread 0 now advances
  if (tid == 0) {
    sc.est_w = est_w;
    sc.mm_total += static_cast<double>(1ull << depth);
    sc.mm_count += 1.0;
    sc.warm_iter += 1;
  }
  return grp.parity;
}

// ADAPT is a compile-time copy of ChainParams::adapt: the sampling instance carries none
// of the adaptation code (the kernel competes for the instruction cache)
// FREE: the free-running mode (advance(), ChainParams::eval_budget) -- one transition per
// run(), draw rows and the re-use of the metric's factor taken from RunLimits
template <class Target, int T, int K, bool ADAPT = true, class Real = double,
          bool FREE = false>
struct ChainRunner {
  using V = VecT<T, K, Real>;
  const ChainParams& p;
  Group<T>& grp;
  Target tgt;
  double* scr;   // this slot's scratch (rows of Real, then the Adam queue in doubles)
  int ld, tid;
  // registers: the live integrator state / newest leaf, and the metric
  Real th[K][2], rho[K][2], g[K][2];
  Real im[K][2];
  // shared memory of this chain (chain_smem_doubles): the macro-step start state =
  // previous leaf, and the stack of finished sub-trees
  Real* s_ths; Real* s_rhos; Real* s_gs;
  double* st_logW; double* st_lp;
  // per-chain scalars live in shared memory; read-modify-write only by thread 0,
  // read by others only after a barrier.
  ChainScalars& sc;
  DecisionCache* dc = nullptr;  // shared memory, control warp only (null: no look-ahead)
  AdamQueue* aq = nullptr;      // shared memory; the entries live behind the scratch vectors
  RunLimits* rl = nullptr;      // shared memory
  static constexpr int kAdamLanes = T >= 32 ? 32 : T;
  __device__ __forceinline__ double* adam_dH() const {
    return scr + static_cast<long long>(scratch_vectors(p.max_depth)) * ld;
  }
  uint32_t u_iter;
  unsigned long long evals;

  __device__ ChainRunner(const ChainParams& p_, Group<T>& grp_, double* scr_,
                         ChainScalars& sc_, double* chain_smem)
      : p(p_), grp(grp_), scr(scr_), ld(p_.ld), tid(grp_.tid), sc(sc_) {
    st_logW = chain_smem;
    st_lp = st_logW + kMaxDepth;
    s_ths = reinterpret_cast<Real*>(chain_smem + 2 * kMaxDepth);
    s_rhos = s_ths + p_.ld;
    s_gs = s_ths + 2 * p_.ld;
  }

  __device__ __forceinline__ Real* sv(int v) const {
    return reinterpret_cast<Real*>(scr) + static_cast<long long>(v) * ld;
  }

  // the live state becomes / is restored from the macro-step start state
  __device__ __forceinline__ void park_start() {
    V::store(s_ths, ld, tid, th);
    V::store(s_rhos, ld, tid, rho);
    V::store(s_gs, ld, tid, g);
  }
  __device__ __forceinline__ void load_start() {
    V::load(s_ths, ld, tid, th);
    V::load(s_rhos, ld, tid, rho);
    V::load(s_gs, ld, tid, g);
  }

  // one leapfrog micro-step, walnuts.hpp:329-332
  __device__ __forceinline__ void leapfrog(Real h, Real hh, Real& lp_part) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        rho[k][v] = madd(hh, g[k][v], rho[k][v]);
        th[k][v] = madd(mul_rn(h, im[k][v]), rho[k][v], th[k][v]);
      }
    }
    tgt.grad(th, g, lp_part, grp);
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        rho[k][v] = madd(hh, g[k][v], rho[k][v]);
      }
    }
  }

  // n micro-steps from the live state; logp and joint (util.hpp:220-223).  With
  // `with_dots` the U-turn dots (:192-201) of the end state against the start
  // state (shared memory) ride in the same reduction.
  __device__ __forceinline__ void integrate(int n, double h, double& lp, double& H,
                                            bool with_dots, double& dot_new,
                                            double& dot_old) {
    const Real h_r = static_cast<Real>(h), hh = static_cast<Real>(0.5 * h);
    Real lp_part = 0;
    for (int j = 0; j < n; ++j) leapfrog(h_r, hh, lp_part);
    evals += static_cast<unsigned long long>(n);
    Real kin = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        kin = madd(im[k][v], mul_rn(rho[k][v], rho[k][v]), kin);
      }
    }
    if (with_dots) {
      Real a = 0, b = 0;
      Real ths[K][2], rhos[K][2];
      V::load(s_ths, ld, tid, ths);
      V::load(s_rhos, ld, tid, rhos);
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          Real sd = mul_rn(im[k][v], add_rn(th[k][v], -ths[k][v]));
          a = madd(rho[k][v], sd, a);
          b = madd(rhos[k][v], sd, b);
        }
      }
      double r[4] = {static_cast<double>(lp_part), static_cast<double>(kin),
                     static_cast<double>(a), static_cast<double>(b)};
      grp.sum(r);
      lp = r[0];
      H = r[0] + (-0.5 * r[1]);
      dot_new = r[2];
      dot_old = r[3];
    } else {
      double r[2] = {static_cast<double>(lp_part), static_cast<double>(kin)};
      grp.sum(r);
      lp = r[0];
      H = r[0] + (-0.5 * r[1]);
    }
  }

  // macro_step (:307-345) + reversible (:254-279).  On entry the live registers AND the
  // parked start state hold the D-end of the span; on success the new leaf is in
  // (th, rho, g) with (lpn, Hn) and the parked state still is the start.
  __device__ __forceinline__ bool macro_step(int dir, double step, int min_micro,
                                             double Hs, double& lpn, double& Hn,
                                             bool with_dots, double& dot_new,
                                             double& dot_old) {
    // One loop, one integrate() site (code size): forward rungs of the halving
    // ladder, then -- once a rung conserves the Hamiltonian -- the reversibility
    // ladder, which re-integrates from the accepted end state with the momentum
    // flipped at every coarser rung and must find none acceptable.
    double h = dir > 0 ? step : -step;
    int n = min_micro, rung = 0;
    bool reversing = false, first_rev = true;
    int cur_n = n;
    double cur_h = h;
    while (true) {
      double lp2, H2, d_new = 0.0, d_old = 0.0;
      integrate(cur_n, cur_h, lp2, H2, with_dots && !reversing, d_new, d_old);
      if (!reversing) {
        lpn = lp2; Hn = H2; dot_new = d_new; dot_old = d_old;
        if (rung == 0 && ADAPT && tid == 0) {  // coarsest attempt only (:335-338)
          adam_dH()[aq->n++] = fabs(Hs - Hn);  // worked off at the end of the transition
        }
        if (!(fabs(Hs - Hn) <= p.max_error)) {
          ++rung;
          if (rung >= p.max_halvings) return false;
          n *= 2; h *= 0.5;
          cur_n = n; cur_h = h;
          load_start();
          continue;
        }
        if (tid == 0) sc.rung_sum += rung;
        if (n == 1 || n < 2 * min_micro) return true;
        V::store(sv(E_TH), ld, tid, th);
        V::store(sv(E_RHO), ld, tid, rho);
        V::store(sv(E_G), ld, tid, g);
        reversing = true;
      } else if (fabs(H2 - Hn) <= p.max_error) {
        return false;  // a coarser rung is acceptable from the end: irreversible
      }
      if (cur_n < 2 * min_micro) {  // ladder exhausted: restore the accepted leaf
        V::load(sv(E_TH), ld, tid, th);
        V::load(sv(E_RHO), ld, tid, rho);
        V::load(sv(E_G), ld, tid, g);
        return true;
      }
      cur_n /= 2;
      cur_h *= 2;
      if (!first_rev) {
        V::load(sv(E_TH), ld, tid, th);
        V::load(sv(E_RHO), ld, tid, rho);
        V::load(sv(E_G), ld, tid, g);
      }
      first_rev = false;
#pragma unroll
      for (int k = 0; k < K; ++k) { rho[k][0] = -rho[k][0]; rho[k][1] = -rho[k][1]; }
    }
  }

  // uturn (:192-201) between a far state F (in scratch) and the newest leaf
  // L = (th, rho); dir gives the time order.
  __device__ __forceinline__ bool uturn(const Real* thF_row, const Real* rhoF_row,
                                        int dir) {
    Real thF[K][2], rhoF[K][2];
    V::load(thF_row, ld, tid, thF);
    V::load(rhoF_row, ld, tid, rhoF);
    Real a = 0, b = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        Real sd = mul_rn(im[k][v], add_rn(th[k][v], -thF[k][v]));
        a = madd(rho[k][v], sd, a);
        b = madd(rhoF[k][v], sd, b);
      }
    }
    double r[2] = {static_cast<double>(a), static_cast<double>(b)};
    grp.sum(r);
    if (dir < 0) { r[0] = -r[0]; r[1] = -r[1]; }
    return r[0] < 0 || r[1] < 0;
  }

  // combine<U> (:368-387) on the control warp: {take_new, logW of the union}
  __device__ __forceinline__ void merge_decision(bool metropolis, double logW_old,
                                                 double logW_new, uint32_t gchain,
                                                 uint32_t iter, uint32_t index,
                                                 bool& take_new, double& logW) {
    double r[2] = {0.0, 0.0};
    if (grp.ctl()) {
      MergeResult m;
      if (dc != nullptr) {
        constexpr uint32_t kLanes = T >= 32 ? 32u : static_cast<uint32_t>(T);
        if (!(dc->valid && dc->iter == iter && index - dc->base < kLanes)) {
          __syncwarp();
          if (static_cast<uint32_t>(grp.lane) < kLanes) {
            dc->logu[grp.lane] = log_uniform(p.seed, gchain, iter, index + grp.lane);
          }
          __syncwarp();
          if (grp.lane == 0) { dc->base = index; dc->iter = iter; dc->valid = 1; }
          __syncwarp();
        }
        m = merge_scalar_logu(metropolis, logW_old, logW_new, dc->logu[index - dc->base]);
      } else {
        m = merge_scalar(metropolis, logW_old, logW_new, p.seed, gchain, iter, index);
      }
      r[0] = m.take_new ? 1.0 : 0.0;
      r[1] = m.logW;
    }
    grp.bcast(r);
    take_new = r[0] != 0.0;
    logW = r[1];
  }

  // copy one vector between rows (parked state / scratch), through registers
  __device__ __forceinline__ void copy_row(Real* dst, const Real* src) {
    Real t[K][2];
    V::load(src, ld, tid, t);
    V::store(dst, ld, tid, t);
  }
  // fp64 rows of the session (theta, stored draws) <-> rows of the working precision
  __device__ __forceinline__ void row_from64(Real* dst, const double* src) {
    Real t[K][2];
    V::load64(src, tid, t);
    V::store(dst, ld, tid, t);
  }
  __device__ __forceinline__ void row_to64(double* dst, const Real* src) {
    Real t[K][2];
    V::load(src, ld, tid, t);
    V::store64(dst, tid, t);
  }

  __device__ __forceinline__ void run(int chain) {
    const uint32_t gchain = p.chain_offset + static_cast<uint32_t>(chain);
    if (tid == 0) {
      sc = p.sc[chain];
      // the look-ahead cache is keyed by (iteration, index) only: a slot that takes a
      // second chain in the same launch must not see the first chain's uniforms
      if (dc != nullptr) dc->valid = 0;
    }
    grp.sync();
    u_iter = sc.iter;
    evals = 0;
    tgt.init(p, tid);
    double* theta_row = p.theta + static_cast<long long>(chain) * ld;
    // the adaptation scalars (estimator weight, min-micro controller, warm-up count) stay
    // in shared memory and are read where they are used
    auto est_row = [&]() { return p.est + static_cast<long long>(chain) * 4 * ld; };
    // the chain's current position lives in scratch row A_SEL between transitions
    row_from64(sv(A_SEL), theta_row);
    bool have_factor = false;
    if constexpr (FREE) have_factor = rl->keep_metric != 0;
    if (!ADAPT && !have_factor) {
      // Cholesky factor of the fixed metric, sqrt().inverse() (walnuts.hpp:647): constant
      // over the launch, so its square roots and divisions are paid once per chain
      double im64[K][2];
      Vec<T, K>::load(p.inv_mass + static_cast<long long>(chain) * ld, ld, tid, im64);
      Real c[K][2];
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const bool pad = 2 * (tid + k * T) + v >= p.D;
          im[k][v] = static_cast<Real>(im64[k][v]);
          c[k][v] = static_cast<Real>(
              div_noinline(1.0, sqrt_noinline(pad ? 1.0 : im64[k][v])));
        }
      }
      V::store(sv(A_IM), ld, tid, c);
    }

    // draw row of transition `it`; evaluated where it is used (limits and parameters come
    // from the constant bank and shared memory: nothing is kept live across a transition)
    auto row_of = [&](int it) -> long long {
      if constexpr (FREE) return (p.rows ? rl->rows : p.draw_base) + it;
      return p.draw_base + it;
    };
    for (int it = 0; it < p.n_iter; ++it) {
      const uint32_t iter = u_iter;
      uint32_t sctr = 0;
      double step;
      int min_micro;
      // ---- metric, step, min-micro for this transition
      if (ADAPT) {
        grp.parity = adapt_begin<T, K, Real>(p, grp, sc, est_row(), sv(A_IM));
        V::load(sv(A_IM), ld, tid, im);
      }
      step = sc.step;
      min_micro = sc.min_micro;
      // lanes beyond D (padding of the 2*T*K register slots) carry a unit metric so
      // that every quotient below stays finite; their theta / rho / grad stay 0
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (2 * (tid + k * T) + v >= p.D) im[k][v] = 1;
        }
      }
      if (p.im_out) {
        V::store64(p.im_out + (static_cast<long long>(chain) * p.draw_cap + row_of(it)) * ld,
                   tid, im);
      }
      // ---- momentum refresh rho = chol_mass * z  (walnuts.hpp:528-529)
      V::load(sv(A_SEL), ld, tid, th);
      {
        Real c[K][2];
        if (!ADAPT) V::load(sv(A_IM), ld, tid, c);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = tid + k * T;
          double z0 = 0.0, z1 = 0.0;
          if (2 * j < p.D) {
            const double2 z = momentum_normals(p.seed, gchain, iter, j);
            z0 = z.x;
            z1 = (2 * j + 1 >= p.D) ? 0.0 : z.y;
          }
          // adaptive: inverse().sqrt() (adaptive_walnuts.hpp:236);
          // fixed:    sqrt().inverse() (walnuts.hpp:647), precomputed above
          const Real c0 = ADAPT ? static_cast<Real>(sqrt_noinline(div_noinline(
                                      1.0, static_cast<double>(im[k][0])))) : c[k][0];
          const Real c1 = ADAPT ? static_cast<Real>(sqrt_noinline(div_noinline(
                                      1.0, static_cast<double>(im[k][1])))) : c[k][1];
          rho[k][0] = mul_rn(c0, static_cast<Real>(z0));
          rho[k][1] = mul_rn(c1, static_cast<Real>(z1));
        }
      }
      // ---- initial point (walnuts.hpp:532-535)
      double lp0, H0;
      {
        Real lp_part;
        tgt.grad(th, g, lp_part, grp);
        evals += 1;
        Real kin = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            kin = madd(im[k][v], mul_rn(rho[k][v], rho[k][v]), kin);
          }
        }
        double r[2] = {static_cast<double>(lp_part), static_cast<double>(kin)};
        grp.sum(r);
        lp0 = r[0];
        H0 = r[0] + (-0.5 * r[1]);
      }
      V::store(sv(A_TH_BK), ld, tid, th);  V::store(sv(A_TH_FW), ld, tid, th);
      V::store(sv(A_RHO_BK), ld, tid, rho); V::store(sv(A_RHO_FW), ld, tid, rho);
      V::store(sv(A_G_BK), ld, tid, g);    V::store(sv(A_G_FW), ld, tid, g);
      double H_bk = H0, H_fw = H0, logW = H0, lp_sel = lp0;
      park_start();
      int regs_dir = 0;  // +-1: the live / parked state equals that end of the span
      bool first_ext = true;

      int depth;
      for (depth = 1; depth <= p.max_depth; ++depth) {
        const bool fwd = direction_bit(p.seed, gchain, iter, sctr++);  // :552
        const int dir = fwd ? 1 : -1;
        if (!first_ext && regs_dir != dir) {
          const int b = fwd ? A_TH_FW : A_TH_BK;
          V::load(sv(b), ld, tid, th);
          V::load(sv(b + 1), ld, tid, rho);
          V::load(sv(b + 2), ld, tid, g);
          park_start();
        }
        first_ext = false;
        double Hs = fwd ? H_fw : H_bk;
        double lps = 0.0;  // logp of the previous leaf (the level-0 entry)
        // ---- build_span(depth-1): 2^(depth-1) leaves, binary-counter merges.
        // The level-0 entry (a single leaf) is never written out: it is the
        // macro-step start state, still parked when leaf i+1 arrives.
        const int nleaf = 1 << (depth - 1);
        int sp = 0;
        bool ok = true;
        double sub_logW = 0.0, sub_lp = 0.0;
        for (int i = 0; i < nleaf; ++i) {
          double lpn, Hn, dot_new = 0.0, dot_old = 0.0;
          const bool odd = (i & 1) != 0;
          if (tid == 0) sc.macro_steps += 1;
          ok = macro_step(dir, step, min_micro, Hs, lpn, Hn, odd, dot_new, dot_old);
          if (!ok) break;
          double cur_logW = Hn, cur_lp = lpn;
          // selection of the span being assembled: -1 newest leaf (live),
          // -2 previous leaf (parked), >= 0 already in that stack slot
          int cur_sel = -1;
          const int nm = __ffs(~i) - 1;  // trailing one bits of i = merges
          if (odd) {
            // merge with the previous leaf: uturn (:490) from the fused dots
            if (dir < 0) { dot_new = -dot_new; dot_old = -dot_old; }
            if (dot_new < 0 || dot_old < 0) { ok = false; break; }
            bool take_new;
            double lw;
            merge_decision(false, Hs, cur_logW, gchain, iter, sctr++, take_new, lw);
            if (!take_new) { cur_sel = -2; cur_lp = lps; }
            cur_logW = lw;
            for (int m = 1; m < nm; ++m) {
              const int s = sp - 1;
              const int sb = ST_BASE + 3 * s;
              if (uturn(sv(sb + ST_THF), sv(sb + ST_RHOF), dir)) { ok = false; break; }
              // entry s is read before the barrier inside merge_decision and rewritten
              // (below) only after it
              const double logW_s = st_logW[s], lp_s = st_lp[s];
              merge_decision(false, logW_s, cur_logW, gchain, iter, sctr++, take_new, lw);
              if (take_new) {
                if (cur_sel >= 0) {
                  copy_row(sv(sb + ST_SEL), sv(ST_BASE + 3 * cur_sel + ST_SEL));
                  cur_sel = s;
                }
              } else {
                cur_sel = s;
                cur_lp = lp_s;
              }
              cur_logW = lw;
              sp = s;
            }
            if (!ok) break;
            const int sb = ST_BASE + 3 * sp;
            if (nm == 1) {  // new two-leaf entry: its first state is the previous leaf
              copy_row(sv(sb + ST_THF), s_ths);
              copy_row(sv(sb + ST_RHOF), s_rhos);
            }
            if (cur_sel == -1) V::store(sv(sb + ST_SEL), ld, tid, th);
            if (cur_sel == -2) copy_row(sv(sb + ST_SEL), s_ths);
            // one writer; a lane that has not yet read the old entry is waited for (above
            // one warp the barrier inside merge_decision has done that, and the next read
            // of the entry lies behind the next reduction's barrier)
            if constexpr (T <= 32) grp.sync();
            if (tid == 0) {
              st_logW[sp] = cur_logW;
              st_lp[sp] = cur_lp;
            }
            if constexpr (T <= 32) grp.sync();
            ++sp;
          }
          sub_logW = cur_logW;
          sub_lp = cur_lp;
          // the newest leaf becomes the start of the next macro step
          park_start();
          Hs = Hn;
          lps = lpn;
        }
        if (!ok) break;  // extension rejected, span unchanged (:543-545)
        // ---- top level: U-turn across the whole span, then Metropolis merge
        const int farb = fwd ? A_TH_BK : A_TH_FW;
        const bool ut = uturn(sv(farb), sv(farb + 1), dir);  // :546 (live == parked here)
        bool take;
        double lw;
        merge_decision(true, logW, sub_logW, gchain, iter, sctr++, take, lw);
        if (take) {
          if (nleaf == 1) {
            V::store(sv(A_SEL), ld, tid, th);
          } else {
            copy_row(sv(A_SEL), sv(ST_BASE + ST_SEL));
          }
          lp_sel = sub_lp;
        }
        const int nb = fwd ? A_TH_FW : A_TH_BK;
        V::store(sv(nb), ld, tid, th);
        V::store(sv(nb + 1), ld, tid, rho);
        V::store(sv(nb + 2), ld, tid, g);
        if (fwd) H_fw = Hs; else H_bk = Hs;
        logW = lw;
        regs_dir = dir;
        if (ut) break;  // :556-558
      }
      // ---- the draw (scratch row A_SEL)
      if (ADAPT) {
        grp.parity = adapt_end<Target, T, K, Real>(p, grp, sc, est_row(), sv(A_SEL), depth);
      } else if (tid == 0) {
        // WelfordAccumulator::observe (sampler.hpp:87-88)
        sc.lp_n += 1;
        const double delta = lp_sel - sc.lp_mean;
        sc.lp_mean += delta / static_cast<double>(sc.lp_n);
        sc.lp_m2 += delta * (lp_sel - sc.lp_mean);
      }
      u_iter += 1;
      if (ADAPT && grp.ctl()) adam_flush<kAdamLanes>(sc, p, *aq, adam_dH(), grp.lane);
      if (tid == 0) {
        sc.last_depth = depth;
        sc.last_lp = lp_sel;
      }
      if (p.draws) {
        row_to64(p.draws + (static_cast<long long>(chain) * p.draw_cap + row_of(it)) * ld,
                 sv(A_SEL));
      }
      if (tid == 0) {
        const long long o = static_cast<long long>(chain) * p.draw_cap + row_of(it);
        if (p.lp_out) p.lp_out[o] = lp_sel;
        if (p.depth_out) p.depth_out[o] = depth;
        if (p.step_out) p.step_out[o] = ADAPT ? exp_noinline(sc.adam_x) : sc.step;
      }
    }
    row_to64(theta_row, sv(A_SEL));
    // (a one-leaf transition of a one-warp group meets no barrier of its own -- shuffles
    // only -- between the lanes' reads of the record above and this write)
    if constexpr (T <= 32) grp.sync();
    if (tid == 0) {
      sc.grad_evals += evals;
      sc.iter = u_iter;
      p.sc[chain] = sc;
    }
    grp.sync();
  }

  // One ticket.  Quota launch: run() once, n_iter transitions.  Free-running launch
  // (ChainParams::eval_budget): the chain's transitions one run() at a time while its
  // budget -- less what earlier launches overspent -- is not used up.  run() has this one
  // call site and its hot loop is untouched; per transition the free-running mode adds the
  // chain record's round trip and one copy of the position row.
  __device__ __forceinline__ void advance(int chain) {
    // nothing of this bookkeeping lives in registers across run(): the limits sit in shared
    // memory, and whether the launch runs free is a kernel parameter (constant bank)
    if (!FREE) { run(chain); return; }
    if (p.eval_budget > 0) {
      if (tid == 0) {
        const ChainScalars& g = p.sc[chain];
        const long long done0 = ADAPT ? static_cast<long long>(g.warm_iter)
                                      : static_cast<long long>(g.lp_n);
        rl->budget = p.eval_budget - g.eval_debt;
        rl->evals0 = g.grad_evals;
        rl->spent = 0;
        rl->rows = p.rows ? p.rows[chain] : p.draw_base;
        long long cap = p.free_cap;
        if (p.iter_cap - done0 < cap) cap = p.iter_cap > done0 ? p.iter_cap - done0 : 0;
        if (p.rows && p.draws && p.draw_cap - rl->rows < cap) {  // the chain's rows are full
          cap = p.draw_cap > rl->rows ? p.draw_cap - rl->rows : 0;
        }
        rl->cap = cap;
        rl->done = 0;
        rl->keep_metric = 0;
      }
      grp.sync();
    }
    while (true) {
      if (p.eval_budget > 0 && !(rl->done < rl->cap && rl->spent < rl->budget)) break;
      run(chain);  // ends with a barrier: every thread is past its reads of the limits
      if (p.eval_budget <= 0) return;
      if (tid == 0) {
        rl->spent = static_cast<long long>(p.sc[chain].grad_evals - rl->evals0);
        rl->done += 1;
        rl->rows += 1;
        rl->keep_metric = 1;
      }
      grp.sync();
    }
    if (tid == 0) {
      // a chain that stopped at a cap (block full, last iteration of the phase) idles and
      // owes nothing; otherwise the excess of its last transition is carried over
      const long long over = rl->spent - rl->budget;
      p.sc[chain].eval_debt = (rl->done < rl->cap || over > 0) ? over : 0;
      if (p.rows && p.draws) p.rows[chain] = rl->rows;
    }
    grp.sync();
  }
};

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// FREE selects the free-running mode (ChainRunner::advance).  It is a separate kernel so
// that the quota launch's code -- registers, layout, instruction-cache footprint -- is
// exactly what it is without the mode: these kernels sit at their register caps and the
// one-warp shapes are instruction-fetch bound (wrapping run() in the free-running loop
// cost the D = 100 warm-up 14 % and the D = 1000 sampling kernel 200 bytes of spills).
template <class Target, int T, int K, int CTA, int MINB, bool ADAPT, class Real = double,
          bool FREE = false>
__global__ void __launch_bounds__(CTA, MINB)
walnuts_chain_kernel(const ChainParams p) {
  extern __shared__ double chain_smem[];  // [CTA / T][chain_smem_doubles(ld)]
  __shared__ double red_smem[group_smem_doubles<T>()];
  __shared__ ChainScalars sc_smem[CTA / T];
  __shared__ DecisionCache dc_smem[CTA / T];
  __shared__ AdamQueue aq_smem[ADAPT ? CTA / T : 1];
  __shared__ RunLimits rl_smem[FREE ? CTA / T : 1];
  __shared__ int next_chain;
  Group<T> grp;
  grp.lane = threadIdx.x & 31;
  grp.red = red_smem;
  grp.parity = 0;
  int slot;
  if constexpr (T == 32) {
    grp.tid = grp.lane;
    grp.warp = 0;
    slot = blockIdx.x * (CTA / 32) + (threadIdx.x >> 5);
  } else {
    static_assert(T == 32 || CTA == T, "one chain per CTA above one warp");
    grp.tid = threadIdx.x;
    grp.warp = threadIdx.x >> 5;
    slot = blockIdx.x;
  }
  double* scr = p.scratch + static_cast<long long>(slot) * p.scratch_stride;
  ChainRunner<Target, T, K, ADAPT, Real, FREE> runner(
      p, grp, scr, sc_smem[threadIdx.x / T],
      chain_smem + static_cast<int>(threadIdx.x / T) * chain_smem_doubles(p.ld));
  if (grp.tid == 0) {
    dc_smem[threadIdx.x / T].valid = 0;
    if (ADAPT) aq_smem[threadIdx.x / T].n = 0;
  }
  grp.sync();
  runner.dc = &dc_smem[threadIdx.x / T];
  if (FREE) runner.rl = &rl_smem[threadIdx.x / T];
  if (ADAPT) runner.aq = &aq_smem[threadIdx.x / T];
  while (true) {
    int chain;
    if constexpr (T == 32) {
      chain = 0;
      if (grp.lane == 0) chain = static_cast<int>(atomicAdd(p.ticket, 1u));
      chain = __shfl_sync(0xffffffffu, chain, 0);
    } else {
      if (threadIdx.x == 0) next_chain = static_cast<int>(atomicAdd(p.ticket, 1u));
      __syncthreads();
      chain = next_chain;
      __syncthreads();
    }
    if (chain >= p.C) break;
    if constexpr (FREE) {
      runner.advance(p.order ? p.order[chain] : chain);
    } else {
      runner.run(p.order ? p.order[chain] : chain);
    }
  }
}
#endif  // __CUDACC__

}  // namespace wb200
