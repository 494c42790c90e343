"""TEST INFRASTRUCTURE: g++ build of the device transition state machine (T = 1)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SO = HERE / "libemu.so"
SRC = HERE / "emu.cpp"
CSRC = HERE.parents[1] / "walnuts_b200" / "csrc"
KIND = {"std_normal": 0, "diag_gaussian": 1, "funnel": 2}


class EmuTuning(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("max_depth", "max_halvings", "min_micro")] + [
        (n, C.c_double) for n in ("max_error", "mass_init_count", "macro_target",
                                  "adam_target", "adam_lr", "adam_b1", "adam_b2",
                                  "adam_eps", "adam_decay")]


def build(exact_arith: bool = False):
    """exact_arith: -DWB200_EXACT_ARITH, every rounding separate (the reference's baseline
    build); default: the fused policy the shipped kernels use."""
    so = HERE / ("libemu_exact.so" if exact_arith else "libemu.so")
    deps = [SRC, CSRC / "chain_kernel.cuh", CSRC / "tick_kernel.cuh", CSRC / "philox.cuh",
            CSRC / "host_shims.hpp"]
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared",
                        "-I/usr/local/cuda/include"] +
                       (["-DWB200_EXACT_ARITH"] if exact_arith else []) +
                       ["-o", str(so), str(SRC)], check=True)
    return C.CDLL(str(so))


def run_chain(lib, kind, D, tparam, cfg, seed, chain, th0, m0, step0, nw, ns, engine="chain",
              warm_ticks=-1, samp_ticks=-1, eval_budget=0):
    """cfg is an oracle.binding.OracleConfig; returns the device-code results.
    engine: "chain" (chain_kernel.cuh) or "tick" (tick_kernel.cuh).
    engine "tick" with warm_ticks / samp_ticks >= 0: that phase runs free for exactly
    that many ticks (nw + ns is then the draw capacity); result["rows"] = (rows after
    warm-up, rows at the end).
    engine "chain" with eval_budget > 0: both phases as free-running launches of that many
    gradient evaluations each (ChainParams::eval_budget) until nw / ns iterations are done;
    result["launches"] counts them."""
    t = EmuTuning(cfg.max_trajectory_doublings, cfg.max_step_halvings, cfg.min_micro_steps,
                  cfg.max_hamiltonian_error, cfg.mass_init_count,
                  cfg.max_macro_steps_target, cfg.step_accept_rate_target,
                  cfg.step_learning_rate, cfg.step_gradient_decay,
                  cfg.step_sq_gradient_decay, cfg.step_stabilization,
                  cfg.step_learn_rate_decay)
    n = nw + ns
    draws, lp = np.zeros((n, D)), np.zeros(n)
    depth, st = np.zeros(n, np.int32), np.zeros(n)
    im, imo = np.zeros((max(nw, 1), D)), np.zeros(D)
    so, mm, ev = C.c_double(0), C.c_int(0), C.c_ulonglong(0)

    def dp(a):
        return a.ctypes.data_as(C.POINTER(C.c_double))

    tp = None if tparam is None else np.ascontiguousarray(tparam, np.float64)
    th0 = np.ascontiguousarray(th0, np.float64)
    m0 = np.ascontiguousarray(m0, np.float64)
    rows = (C.c_longlong * 2)(nw, nw + ns)
    if warm_ticks >= 0 or samp_ticks >= 0:
        assert engine == "tick"
        rc = lib.emu_run_chain_tick_free(
            KIND[kind], D, None if tp is None else dp(tp), C.byref(t), C.c_uint32(seed),
            C.c_uint32(chain), dp(th0), dp(m0), C.c_double(step0), nw, ns, warm_ticks,
            samp_ticks, dp(draws), dp(lp), depth.ctypes.data_as(C.POINTER(C.c_int)), dp(st),
            dp(im), dp(imo), C.byref(so), C.byref(mm), C.byref(ev), rows)
    elif eval_budget > 0:
        assert engine == "chain"
        launches = C.c_int(0)
        rc = lib.emu_run_chain_free(
            KIND[kind], D, None if tp is None else dp(tp), C.byref(t), C.c_uint32(seed),
            C.c_uint32(chain), dp(th0), dp(m0), C.c_double(step0), nw, ns,
            C.c_longlong(eval_budget), dp(draws), dp(lp),
            depth.ctypes.data_as(C.POINTER(C.c_int)), dp(st), dp(im), dp(imo),
            C.byref(so), C.byref(mm), C.byref(ev), C.byref(launches))
        if rc == 0:
            return dict(rows=(nw, nw + ns), draws=draws, lp=lp, depth=depth, step_trace=st,
                        warmup_inv_mass=im[:nw], inv_mass=imo, step=so.value,
                        min_micro=mm.value, grad_evals=ev.value, launches=launches.value)
    else:
        fn = lib.emu_run_chain if engine == "chain" else lib.emu_run_chain_tick
        rc = fn(KIND[kind], D, None if tp is None else dp(tp), C.byref(t),
                C.c_uint32(seed), C.c_uint32(chain), dp(th0), dp(m0),
                C.c_double(step0), nw, ns, dp(draws), dp(lp),
                depth.ctypes.data_as(C.POINTER(C.c_int)), dp(st), dp(im),
                dp(imo), C.byref(so), C.byref(mm), C.byref(ev))
    if rc != 0:
        raise RuntimeError(f"emu_run_chain rc={rc}")
    return dict(rows=(rows[0], rows[1]), draws=draws, lp=lp, depth=depth, step_trace=st, warmup_inv_mass=im[:nw],
                inv_mass=imo, step=so.value, min_micro=mm.value, grad_evals=ev.value)
