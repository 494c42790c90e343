// Lock-step ("tick") form of the WALNUTS transition for targets whose gradient is
// a cross-chain batched contraction (Bayesian logistic regression on the tensor
// cores): every chain is a masked, resumable state machine that is advanced from
// one gradient request to the next per tick.
//
//   tick t:   tick kernel  : consume (logp, grad) of the requested position,
//                            finish the micro-step, run all the control logic that
//                            needs no gradient (energy test, halving / reversibility
//                            ladders, U-turns, Barker / Metropolis merges, end of
//                            transition, adaptation, momentum refresh), start the
//                            next micro-step and post the next position
//             gradient     : one batched evaluation for ALL chains (GEMMs)
//
// State lives in HBM, structure-of-arrays: TH / G / LP are [C][ld] / [C][ld] / [C]
// arrays shared with the gradient stage; everything else of a chain sits in its
// `vecs` block.  Semantics, order of gradient evaluations, U-turn tests and random
// draws are those of chain_kernel.cuh (and hence of the reference, walnuts.hpp:
// 192-563, adaptive_walnuts.hpp:234-251); the same helpers are reused.  One
// difference in COST only: the gradient of the selected draw is carried with the
// selection, so the reference's re-evaluation of logp_grad(theta) at the start of
// every transition (walnuts.hpp:532, same point, same value) is not repeated.
#pragma once
#include "chain_kernel.cuh"

namespace wb200 {

// per-chain vectors in the `vecs` block
enum : int {
  TV_RHO = 0, TV_THS, TV_RHOS, TV_GS, TV_IM, TV_CUR, TV_CUR_G,
  TV_A_TH_BK, TV_A_RHO_BK, TV_A_G_BK, TV_A_TH_FW, TV_A_RHO_FW, TV_A_G_FW,
  TV_A_SEL, TV_A_SEL_G, TV_E_TH, TV_E_RHO, TV_E_G, TV_ST_BASE
};
enum : int { TS_THF = 0, TS_RHOF = 1, TS_SEL = 2, TS_SEL_G = 3 };
__host__ __device__ inline int tick_vectors(int max_depth) {
  return TV_ST_BASE + 4 * max_depth;
}

enum : int { PC_START_TRANSITION = 0, PC_IN_INTEGRATE = 1, PC_DONE = 2 };

struct TickState {
  int pc;
  int done_iters;
  long long rows;  // draws this chain has stored so far (ragged / free-running mode)
  // transition
  uint32_t sctr;
  int depth, dir, nleaf, leaf_i, sp, regs_dir, first_ext;
  double step;
  int min_micro;
  double H_bk, H_fw, logW, lp_sel, lp_cur;
  double sub_logW, sub_lp;
  // macro step
  int rung, n, cur_n, micro_done, reversing, first_rev, with_dots;
  double h, cur_h, Hs, lps, lpn, Hn, dot_new, dot_old;
};
// The (logW, lp) stack of finished sub-trees lives in its own arrays (TickParams::st_logW /
// st_lp, [C][kMaxDepth]) and is accessed in place: inside TickState its dynamic indexing
// forced every thread's copy of the whole record into local memory (904 B of stack per
// thread, two thirds of the kernel's DRAM writes were evictions of those copies).

struct TickParams {
  ChainParams cp;      // tuning, seeds, est / sc / draws as in the chain kernel
  double* TH;          // [C][ld] requested positions (in: consumed, out: next request)
  double* G;           // [C][ld] gradient at TH   (written by the gradient stage)
  double* LP;          // [C]     logp at TH
  double* vecs;        // [C][nvec][ld]
  long long vec_stride;
  TickState* ts;       // [C]
  double* st_logW;     // [C][kMaxDepth] sub-tree stack: log weights
  double* st_lp;       // [C][kMaxDepth]                 log densities of the selections
  int* active_count;   // chains that still have transitions to run after this tick
  int chain_begin, chain_count;  // the launch advances chains [begin, begin + count)
};

template <int T, int K>
struct TickRunner {
  using V = Vec<T, K>;
  const TickParams& tp;
  const ChainParams& p;
  Group<T>& grp;
  int ld, tid;
  double* vb;  // this chain's vecs block
  double th[K][2], rho[K][2], g[K][2], im[K][2];

  __device__ TickRunner(const TickParams& tp_, Group<T>& grp_)
      : tp(tp_), p(tp_.cp), grp(grp_), ld(tp_.cp.ld), tid(grp_.tid) {}

  __device__ __forceinline__ double* vv(int v) const {
    return vb + static_cast<long long>(v) * ld;
  }
  __device__ __forceinline__ void copy_vec(int dst, int src) {
    double t[K][2];
    V::load(vv(src), ld, tid, t);
    V::store(vv(dst), ld, tid, t);
  }
  __device__ __forceinline__ double kinetic_part() const {
    double kin = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        kin = madd(im[k][v], __dmul_rn(rho[k][v], rho[k][v]), kin);
      }
    }
    return kin;
  }
  // U-turn dots (walnuts.hpp:192-201) of the live state L = (th, rho) against F
  __device__ __forceinline__ void dots_against(int vthF, int vrhoF, double& a, double& b) {
    double thF[K][2], rhoF[K][2];
    V::load(vv(vthF), ld, tid, thF);
    V::load(vv(vrhoF), ld, tid, rhoF);
    a = 0.0; b = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        double sd = __dmul_rn(im[k][v], __dadd_rn(th[k][v], -thF[k][v]));
        a = madd(rho[k][v], sd, a);
        b = madd(rhoF[k][v], sd, b);
      }
    }
  }
  __device__ __forceinline__ bool uturn(int vthF, int vrhoF, int dir) {
    double r[2];
    dots_against(vthF, vrhoF, r[0], r[1]);
    grp.sum(r);
    if (dir < 0) { r[0] = -r[0]; r[1] = -r[1]; }
    return r[0] < 0 || r[1] < 0;
  }
  __device__ __forceinline__ void merge_decision(bool metropolis, double logW_old,
                                                 double logW_new, uint32_t gchain,
                                                 uint32_t iter, uint32_t index,
                                                 bool& take_new, double& logW) {
    double r[2] = {0.0, 0.0};
    if (grp.ctl()) {
      const MergeResult m =
          merge_scalar(metropolis, logW_old, logW_new, p.seed, gchain, iter, index);
      r[0] = m.take_new ? 1.0 : 0.0;
      r[1] = m.logW;
    }
    grp.bcast(r);
    take_new = r[0] != 0.0;
    logW = r[1];
  }

  // first half kick + drift of a micro-step from the live registers; posts TH
  __device__ __forceinline__ void begin_micro_step(double h, double* th_row) {
    const double hh = 0.5 * h;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        rho[k][v] = madd(hh, g[k][v], rho[k][v]);
        th[k][v] = madd(__dmul_rn(h, im[k][v]), rho[k][v], th[k][v]);
      }
    }
    V::store(th_row, ld, tid, th);
    V::store(vv(TV_RHO), ld, tid, rho);
  }

  __device__ __forceinline__ void load_metric(ChainScalars& sc, int chain) {
    if (p.adapt) {
      const double* est_row = p.est + static_cast<long long>(chain) * 4 * ld;
      double Sd[K][2], Ss[K][2];
      V::load(est_row + 1 * ld, ld, tid, Sd);
      V::load(est_row + 3 * ld, ld, tid, Ss);
#pragma unroll
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          im[k][v] = metric_from_sums(Sd[k][v], Ss[k][v], sc.est_w);
        }
      }
    } else {
      V::load(p.inv_mass + static_cast<long long>(chain) * ld, ld, tid, im);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        if (2 * (tid + k * T) + v >= p.D) im[k][v] = 1.0;
      }
    }
  }

  // One tick of one chain.  `sc` / `st` are the chain's records (every thread holds
  // an identical copy; thread 0 writes them back).
  __device__ void tick(int chain) {
    const uint32_t gchain = p.chain_offset + static_cast<uint32_t>(chain);
    ChainScalars sc = p.sc[chain];
    TickState st = tp.ts[chain];
    // every warp of the group holds its copy of the records before thread 0 may write
    // them back at the end of the tick (the mid-integration path has no other barrier)
    grp.sync();
    if (st.pc == PC_DONE) return;
    vb = tp.vecs + static_cast<long long>(chain) * tp.vec_stride;
    double* stk_logW = tp.st_logW + static_cast<long long>(chain) * kMaxDepth;
    double* stk_lp = tp.st_lp + static_cast<long long>(chain) * kMaxDepth;
    double* th_row = tp.TH + static_cast<long long>(chain) * ld;
    double* g_row = tp.G + static_cast<long long>(chain) * ld;
    unsigned long long evals = 0;
    bool request = false;  // a new position has been posted
    // the live registers (th, rho, g) equal the stored macro-step start state (THS, RHOS,
    // GS) and im is loaded: the next macro step can start without reading them back
    bool live_is_start = false;

    while (!request) {
      if (st.pc == PC_START_TRANSITION) {
        // ---- metric, step, min-micro (adaptive_walnuts.hpp:235-245)
        if (p.adapt) {
          st.step = exp_noinline(sc.adam_x);
          st.min_micro = min_micro_steps(sc, p);
        } else {
          st.step = sc.step;
          st.min_micro = sc.min_micro;
        }
        load_metric(sc, chain);
        if (p.adapt) V::store(vv(TV_IM), ld, tid, im);
        const long long row = p.n_iter < 0 ? st.rows : p.draw_base + st.done_iters;
        if (p.im_out) {
          V::store(p.im_out + (static_cast<long long>(chain) * p.draw_cap + row) * ld,
                   ld, tid, im);
        }
        // ---- momentum refresh (walnuts.hpp:528-529); theta, grad, logp carried
        V::load(vv(TV_CUR), ld, tid, th);
        V::load(vv(TV_CUR_G), ld, tid, g);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int j = tid + k * T;
          double z0 = 0.0, z1 = 0.0;
          if (2 * j < p.D) {
            const double2 z = momentum_normals(p.seed, gchain, sc.iter, j);
            z0 = z.x;
            z1 = (2 * j + 1 >= p.D) ? 0.0 : z.y;
          }
          double c0 = p.adapt ? sqrt(1.0 / im[k][0]) : 1.0 / sqrt(im[k][0]);
          double c1 = p.adapt ? sqrt(1.0 / im[k][1]) : 1.0 / sqrt(im[k][1]);
          rho[k][0] = __dmul_rn(c0, z0);
          rho[k][1] = __dmul_rn(c1, z1);
        }
        double r[1] = {kinetic_part()};
        grp.sum(r);
        const double lp0 = st.lp_cur;
        const double H0 = lp0 + (-0.5 * r[0]);
        V::store(vv(TV_A_TH_BK), ld, tid, th);  V::store(vv(TV_A_TH_FW), ld, tid, th);
        V::store(vv(TV_A_RHO_BK), ld, tid, rho); V::store(vv(TV_A_RHO_FW), ld, tid, rho);
        V::store(vv(TV_A_G_BK), ld, tid, g);    V::store(vv(TV_A_G_FW), ld, tid, g);
        V::store(vv(TV_A_SEL), ld, tid, th);    V::store(vv(TV_A_SEL_G), ld, tid, g);
        V::store(vv(TV_THS), ld, tid, th); V::store(vv(TV_RHOS), ld, tid, rho);
        V::store(vv(TV_GS), ld, tid, g);
        live_is_start = true;
        st.H_bk = H0; st.H_fw = H0; st.logW = H0; st.lp_sel = lp0;
        st.regs_dir = 0; st.first_ext = 1; st.sctr = 0;
        st.depth = 0;  // incremented by the doubling step below
        st.pc = 100;   // -> next doubling
      }

      if (st.pc == PC_IN_INTEGRATE) {
        // ---- the gradient at TH has arrived: finish the micro-step (:332)
        V::load(th_row, ld, tid, th);
        V::load(vv(TV_RHO), ld, tid, rho);
        V::load(g_row, ld, tid, g);
        V::load(p.adapt ? vv(TV_IM) : p.inv_mass + static_cast<long long>(chain) * ld,
                ld, tid, im);
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            if (2 * (tid + k * T) + v >= p.D) im[k][v] = 1.0;
          }
        }
        const double hh = 0.5 * st.cur_h;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            rho[k][v] = madd(hh, g[k][v], rho[k][v]);
          }
        }
        evals += 1;
        st.micro_done += 1;
        if (st.micro_done < st.cur_n) {
          begin_micro_step(st.cur_h, th_row);
          request = true;
          break;
        }
        // ---- attempt finished: energies (+ fused U-turn dots for odd leaves)
        const double lp2 = tp.LP[chain];
        double r[3] = {kinetic_part(), 0.0, 0.0};
        const bool dots = st.with_dots && !st.reversing;
        if (dots) dots_against(TV_THS, TV_RHOS, r[1], r[2]);
        grp.sum(r);
        const double H2 = lp2 + (-0.5 * r[0]);
        // ---- body of the macro-step loop (walnuts.hpp:322-343, :254-279)
        bool macro_done = false, macro_ok = false, again = false;
        if (!st.reversing) {
          st.lpn = lp2; st.Hn = H2; st.dot_new = r[1]; st.dot_old = r[2];
          if (st.rung == 0 && p.adapt) {
            adam_update(sc, p, exp_noinline(-fabs(st.Hs - st.Hn)));
          }
          if (!(fabs(st.Hs - st.Hn) <= p.max_error)) {
            st.rung += 1;
            if (st.rung >= p.max_halvings) {
              macro_done = true; macro_ok = false;
            } else {
              st.n *= 2; st.h *= 0.5;
              st.cur_n = st.n; st.cur_h = st.h;
              V::load(vv(TV_THS), ld, tid, th);
              V::load(vv(TV_RHOS), ld, tid, rho);
              V::load(vv(TV_GS), ld, tid, g);
              again = true;
            }
          } else {
            sc.rung_sum += st.rung;
            if (st.n == 1 || st.n < 2 * st.min_micro) {
              macro_done = true; macro_ok = true;
            } else {
              V::store(vv(TV_E_TH), ld, tid, th);
              V::store(vv(TV_E_RHO), ld, tid, rho);
              V::store(vv(TV_E_G), ld, tid, g);
              st.reversing = 1;
            }
          }
        } else if (fabs(H2 - st.Hn) <= p.max_error) {
          macro_done = true; macro_ok = false;  // irreversible
        }
        if (!macro_done && !again) {
          // next rung of the reversibility ladder, or its end
          if (st.cur_n < 2 * st.min_micro) {
            V::load(vv(TV_E_TH), ld, tid, th);
            V::load(vv(TV_E_RHO), ld, tid, rho);
            V::load(vv(TV_E_G), ld, tid, g);
            macro_done = true; macro_ok = true;
          } else {
            st.cur_n /= 2;
            st.cur_h *= 2;
            if (!st.first_rev) {
              V::load(vv(TV_E_TH), ld, tid, th);
              V::load(vv(TV_E_RHO), ld, tid, rho);
              V::load(vv(TV_E_G), ld, tid, g);
            }
            st.first_rev = 0;
#pragma unroll
            for (int k = 0; k < K; ++k) { rho[k][0] = -rho[k][0]; rho[k][1] = -rho[k][1]; }
            again = true;
          }
        }
        if (again) {
          st.micro_done = 0;
          begin_micro_step(st.cur_h, th_row);
          request = true;
          break;
        }
        // ---- the macro step is over: leaf logic (build_span's carry loop)
        bool ok = macro_ok;
        if (ok) {
          const bool odd = (st.leaf_i & 1) != 0;
          double cur_logW = st.Hn, cur_lp = st.lpn;
          int cur_sel = -1;
          const int nm = __ffs(~st.leaf_i) - 1;
          if (odd) {
            double dn = st.dot_new, dol = st.dot_old;
            if (st.dir < 0) { dn = -dn; dol = -dol; }
            if (dn < 0 || dol < 0) ok = false;
            if (ok) {
              bool take_new;
              double lw;
              merge_decision(false, st.Hs, cur_logW, gchain, sc.iter, st.sctr++, take_new, lw);
              if (!take_new) { cur_sel = -2; cur_lp = st.lps; }
              cur_logW = lw;
              for (int m = 1; m < nm && ok; ++m) {
                const int s = st.sp - 1;
                const int sb = TV_ST_BASE + 4 * s;
                if (uturn(sb + TS_THF, sb + TS_RHOF, st.dir)) { ok = false; break; }
                // entry s is read before the barrier inside merge_decision; thread 0
                // rewrites it (below) only after it
                const double logW_s = stk_logW[s], lp_s = stk_lp[s];
                merge_decision(false, logW_s, cur_logW, gchain, sc.iter, st.sctr++,
                               take_new, lw);
                if (take_new) {
                  if (cur_sel >= 0) {
                    copy_vec(sb + TS_SEL, TV_ST_BASE + 4 * cur_sel + TS_SEL);
                    copy_vec(sb + TS_SEL_G, TV_ST_BASE + 4 * cur_sel + TS_SEL_G);
                    cur_sel = s;
                  }
                } else {
                  cur_sel = s;
                  cur_lp = lp_s;
                }
                cur_logW = lw;
                st.sp = s;
              }
            }
            if (ok) {
              const int sb = TV_ST_BASE + 4 * st.sp;
              if (nm == 1) {
                copy_vec(sb + TS_THF, TV_THS);
                copy_vec(sb + TS_RHOF, TV_RHOS);
              }
              if (cur_sel == -1) {
                V::store(vv(sb + TS_SEL), ld, tid, th);
                V::store(vv(sb + TS_SEL_G), ld, tid, g);
              }
              if (cur_sel == -2) {
                copy_vec(sb + TS_SEL, TV_THS);
                copy_vec(sb + TS_SEL_G, TV_GS);
              }
              if constexpr (T <= 32) grp.sync();  // no lane still reads the old entry
              if (tid == 0) {
                stk_logW[st.sp] = cur_logW;
                stk_lp[st.sp] = cur_lp;
              }
              st.sp += 1;
            }
          }
          if (ok) {
            st.sub_logW = cur_logW;
            st.sub_lp = cur_lp;
            V::store(vv(TV_THS), ld, tid, th);
            V::store(vv(TV_RHOS), ld, tid, rho);
            V::store(vv(TV_GS), ld, tid, g);
            live_is_start = true;
            st.Hs = st.Hn;
            st.lps = st.lpn;
            st.leaf_i += 1;
          }
        }
        if (!ok) {
          st.pc = 200;  // extension rejected -> transition ends (walnuts.hpp:543-545)
        } else if (st.leaf_i < st.nleaf) {
          st.pc = 110;  // next leaf of the same sub-tree
        } else {
          // ---- sub-tree complete: top-level U-turn, Metropolis merge (:546-549)
          const bool fwd = st.dir > 0;
          const int farb = fwd ? TV_A_TH_BK : TV_A_TH_FW;
          const bool ut = uturn(farb, farb + 1, st.dir);
          bool take;
          double lw;
          merge_decision(true, st.logW, st.sub_logW, gchain, sc.iter, st.sctr++, take, lw);
          if (take) {
            if (st.nleaf == 1) {
              V::store(vv(TV_A_SEL), ld, tid, th);
              V::store(vv(TV_A_SEL_G), ld, tid, g);
            } else {
              copy_vec(TV_A_SEL, TV_ST_BASE + TS_SEL);
              copy_vec(TV_A_SEL_G, TV_ST_BASE + TS_SEL_G);
            }
            st.lp_sel = st.sub_lp;
          }
          const int nb = fwd ? TV_A_TH_FW : TV_A_TH_BK;
          V::store(vv(nb), ld, tid, th);
          V::store(vv(nb + 1), ld, tid, rho);
          V::store(vv(nb + 2), ld, tid, g);
          if (fwd) st.H_fw = st.Hs; else st.H_bk = st.Hs;
          st.logW = lw;
          st.regs_dir = st.dir;
          st.pc = ut ? 200 : 100;
        }
      }

      if (st.pc == 100) {
        // ---- next doubling (walnuts.hpp:536-552)
        st.depth += 1;
        if (st.depth > p.max_depth) {
          st.pc = 200;
        } else {
          const bool fwd = direction_bit(p.seed, gchain, sc.iter, st.sctr++);
          st.dir = fwd ? 1 : -1;
          if (!st.first_ext && st.regs_dir != st.dir) {
            const int b = fwd ? TV_A_TH_FW : TV_A_TH_BK;
            copy_vec(TV_THS, b);
            copy_vec(TV_RHOS, b + 1);
            copy_vec(TV_GS, b + 2);
            live_is_start = false;
          }
          st.first_ext = 0;
          st.Hs = fwd ? st.H_fw : st.H_bk;
          st.lps = 0.0;
          st.nleaf = 1 << (st.depth - 1);
          st.leaf_i = 0;
          st.sp = 0;
          st.pc = 110;
        }
      }

      if (st.pc == 110) {
        // ---- start a macro step from (THS, RHOS, GS)   (walnuts.hpp:316-327)
        sc.macro_steps += 1;
        st.with_dots = (st.leaf_i & 1);
        st.h = st.dir > 0 ? st.step : -st.step;
        st.n = st.min_micro;
        st.rung = 0;
        st.reversing = 0;
        st.first_rev = 1;
        st.cur_n = st.n;
        st.cur_h = st.h;
        st.micro_done = 0;
        if (!live_is_start) {
          V::load(vv(TV_THS), ld, tid, th);
          V::load(vv(TV_RHOS), ld, tid, rho);
          V::load(vv(TV_GS), ld, tid, g);
          V::load(p.adapt ? vv(TV_IM) : p.inv_mass + static_cast<long long>(chain) * ld,
                  ld, tid, im);
#pragma unroll
          for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
              if (2 * (tid + k * T) + v >= p.D) im[k][v] = 1.0;
            }
          }
        }
        begin_micro_step(st.cur_h, th_row);
        st.pc = PC_IN_INTEGRATE;
        request = true;
        break;
      }

      if (st.pc == 200) {
        // ---- end of the transition: the draw, adaptation, outputs
        const int depth = st.depth;  // loop variable at exit (max_depth + 1 if exhausted)
        double cur[K][2], gsel[K][2];
        V::load(vv(TV_A_SEL), ld, tid, cur);
        V::load(vv(TV_A_SEL_G), ld, tid, gsel);
        V::store(vv(TV_CUR), ld, tid, cur);
        V::store(vv(TV_CUR_G), ld, tid, gsel);
        st.lp_cur = st.lp_sel;
        const bool ragged = p.n_iter < 0;  // free-running: no iteration quota
        const long long row = ragged ? st.rows : p.draw_base + st.done_iters;
        if (p.adapt) {
          double* est_row = p.est + static_cast<long long>(chain) * 4 * ld;
          const double gamma =
              1.0 - 1.0 / (p.mass_init_count + static_cast<double>(sc.warm_iter));
          sc.est_w = gamma * sc.est_w + 1.0;
          const double r_w = 1.0 / sc.est_w;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            double mu[K][2], S[K][2];
            V::load(est_row + (2 * e) * ld, ld, tid, mu);
            V::load(est_row + (2 * e + 1) * ld, ld, tid, S);
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
              for (int v = 0; v < 2; ++v) {
                const double y = e == 0 ? cur[k][v] : gsel[k][v];
                mu[k][v] = kFusedArith
                               ? madd(__dadd_rn(y, -mu[k][v]), r_w, mu[k][v])
                               : __dadd_rn(mu[k][v], __dadd_rn(y, -mu[k][v]) / sc.est_w);
                const double d = __dadd_rn(y, -mu[k][v]);
                S[k][v] = madd(gamma, S[k][v], __dmul_rn(d, d));
              }
            }
            V::store(est_row + (2 * e) * ld, ld, tid, mu);
            V::store(est_row + (2 * e + 1) * ld, ld, tid, S);
          }
          sc.mm_total += static_cast<double>(1ull << depth);
          sc.mm_count += 1.0;
          sc.warm_iter += 1;
        } else {
          sc.lp_n += 1;
          const double delta = st.lp_sel - sc.lp_mean;
          sc.lp_mean += delta / static_cast<double>(sc.lp_n);
          sc.lp_m2 += delta * (st.lp_sel - sc.lp_mean);
        }
        sc.iter += 1;
        sc.last_depth = depth;
        sc.last_lp = st.lp_sel;
        if (p.draws) {
          V::store(p.draws + (static_cast<long long>(chain) * p.draw_cap + row) * ld, ld,
                   tid, cur);
        }
        if (tid == 0) {
          const long long o = static_cast<long long>(chain) * p.draw_cap + row;
          if (p.lp_out) p.lp_out[o] = st.lp_sel;
          if (p.depth_out) p.depth_out[o] = depth;
          if (p.step_out) p.step_out[o] = p.adapt ? exp_noinline(sc.adam_x) : sc.step;
        }
        st.done_iters += 1;
        if (p.draws) st.rows = row + 1;
        // free-running: a chain stops when its rows are full or it has done the phase's
        // iter_cap iterations (a reference chain stops at max_iter, sampler.hpp:82)
        const long long phase_iters = p.adapt ? static_cast<long long>(sc.warm_iter)
                                              : static_cast<long long>(sc.lp_n);
        if (ragged ? ((p.draws && st.rows >= p.draw_cap) ||
                      (p.iter_cap > 0 && phase_iters >= p.iter_cap))
                   : (st.done_iters >= p.n_iter)) {
          st.pc = PC_DONE;
          // keep theta in sync with the chain kernel's convention
          V::store(p.theta + static_cast<long long>(chain) * ld, ld, tid, cur);
          break;
        }
        st.pc = PC_START_TRANSITION;
      }
    }
    sc.grad_evals += evals;
    if (tid == 0) {
      p.sc[chain] = sc;
      tp.ts[chain] = st;
      if (st.pc != PC_DONE) atomicAdd(tp.active_count, 1);
    }
  }
};

#if defined(__CUDACC__)
template <int T, int K, int CTA>
#ifndef WB200_TICK_THREADS_PER_SM
#define WB200_TICK_THREADS_PER_SM 512  // resident threads asked for (register cap 128)
#endif
__global__ void __launch_bounds__(CTA, (CTA <= 128 ? WB200_TICK_THREADS_PER_SM / CTA : 1))
walnuts_tick_kernel(const TickParams tp) {
  __shared__ double red_smem[group_smem_doubles<T>()];
  Group<T> grp;
  grp.lane = threadIdx.x & 31;
  grp.red = red_smem;
  grp.parity = 0;
  int chain;
  if constexpr (T == 32) {
    grp.tid = grp.lane;
    grp.warp = 0;
    chain = blockIdx.x * (CTA / 32) + (threadIdx.x >> 5);
  } else {
    grp.tid = threadIdx.x;
    grp.warp = threadIdx.x >> 5;
    chain = blockIdx.x;
  }
  if (chain >= tp.chain_count) return;
  TickRunner<T, K> runner(tp, grp);
  runner.tick(tp.chain_begin + chain);
}
#endif

}  // namespace wb200
