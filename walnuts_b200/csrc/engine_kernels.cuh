// The kernels around the transition that also depend on the target: batched initialisation
// and the fixed-step orbit.  Device code only (no host headers): this file, chain_kernel.cuh
// and philox.cuh are what a run-time compiled user density is built against
// (user_density.cu, NVRTC).
#pragma once
#include "chain_kernel.cuh"

namespace wb200 {

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

}  // namespace wb200
