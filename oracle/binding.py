"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.

ctypes binding over oracle/oracle_capi.h for the two CPU checkers:

* ``load_oracle()``  -> oracle/liboracle.so          (Eigen-free restatement)
* ``load_ref()``     -> oracle/_ref/libwalnuts_ref.so (unmodified reference
  headers + local Eigen shim; built by ``make -C oracle ref`` where
  /root/reference exists, shipped prebuilt to the GPU box)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs import this module.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libwalnuts_ref.so"
REFERENCE = Path(os.environ.get("WALNUTS_REFERENCE", "/root/reference"))


class OracleTarget(C.Structure):
    _fields_ = [("kind", C.c_int), ("D", C.c_size_t), ("N", C.c_size_t),
                ("data0", C.c_void_p), ("data1", C.c_void_p)]


class OracleConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "min_warmup_iter", "max_warmup_iter", "min_sampling_iter",
        "max_sampling_iter", "max_trajectory_doublings", "max_step_halvings",
        "min_micro_steps", "publish_stride")] + [(n, C.c_double) for n in (
            "max_hamiltonian_error", "step_size_converge_tol",
            "mass_converge_tol", "rhat_converge_tol", "mass_init_count",
            "mass_additive_smoothing", "max_macro_steps_target",
            "step_accept_rate_target", "step_learning_rate",
            "step_gradient_decay", "step_sq_gradient_decay",
            "step_stabilization", "step_learn_rate_decay")]


def default_config(**kw) -> OracleConfig:
    """Reference defaults: config.hpp:626-640 and :947-953."""
    d = dict(min_warmup_iter=50, max_warmup_iter=1000, min_sampling_iter=50,
             max_sampling_iter=1000, max_trajectory_doublings=5,
             max_step_halvings=5, min_micro_steps=1, publish_stride=5,
             max_hamiltonian_error=0.5, step_size_converge_tol=0.1,
             mass_converge_tol=1.0, rhat_converge_tol=1.01,
             mass_init_count=4.0, mass_additive_smoothing=1e-5,
             max_macro_steps_target=15.0, step_accept_rate_target=0.8,
             step_learning_rate=0.05, step_gradient_decay=0.8,
             step_sq_gradient_decay=0.9, step_stabilization=1e-4,
             step_learn_rate_decay=0.5)
    d.update(kw)
    return OracleConfig(**d)


KIND = {"std_normal": 0, "diag_gaussian": 1, "funnel": 2, "logistic": 3,
        "cfunc": 4}


@dataclass
class Target:
    kind: str
    D: int
    prec: Optional[np.ndarray] = None      # diag_gaussian
    X: Optional[np.ndarray] = None         # logistic [N][D] fp64
    y: Optional[np.ndarray] = None         # logistic [N]
    cfunc: Optional[object] = None         # ctypes function pointer
    _keep: list = field(default_factory=list)

    def c(self) -> OracleTarget:
        t = OracleTarget(kind=KIND[self.kind], D=self.D, N=0, data0=None,
                         data1=None)
        if self.kind == "diag_gaussian":
            p = np.ascontiguousarray(self.prec, dtype=np.float64)
            self._keep.append(p)
            t.data0 = p.ctypes.data
        elif self.kind == "logistic":
            X = np.ascontiguousarray(self.X, dtype=np.float64)
            y = np.ascontiguousarray(self.y, dtype=np.float64)
            self._keep += [X, y]
            t.N = X.shape[0]
            t.data0 = X.ctypes.data
            t.data1 = y.ctypes.data
        elif self.kind == "cfunc":
            t.data0 = C.cast(self.cfunc, C.c_void_p)
        return t


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int))


class Checker:
    """One of the two CPU implementations behind the shared C signatures."""

    def __init__(self, path: Path, prefix: str):
        self.lib = C.CDLL(str(path))
        self.prefix = prefix
        self.path = path
        f = self._fn
        f("leapfrog_error").restype = C.c_double
        f("log_sum_exp").restype = C.c_double
        f("log_sum_exp").argtypes = [C.c_double, C.c_double]
        f("logp_momentum").restype = C.c_double
        getattr(self.lib, prefix + "_last_error").restype = C.c_char_p

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def _check(self, rc):
        if rc != 0:
            msg = getattr(self.lib, self.prefix + "_last_error")().decode()
            raise (ValueError if rc == -2 else RuntimeError)(msg)

    # -- scalars ----------------------------------------------------------
    def log_sum_exp(self, a, b):
        return self._fn("log_sum_exp")(a, b)

    def logp_momentum(self, rho, inv_mass):
        rho = np.ascontiguousarray(rho, np.float64)
        im = np.ascontiguousarray(inv_mass, np.float64)
        return self._fn("logp_momentum")(_dp(rho), _dp(im), C.c_size_t(rho.size))

    def leapfrog_error(self, target: Target, theta, rho, inv_mass, step):
        t = target.c()
        a = [np.ascontiguousarray(v, np.float64) for v in (theta, rho, inv_mass)]
        return self._fn("leapfrog_error")(C.byref(t), _dp(a[0]), _dp(a[1]),
                                          _dp(a[2]), C.c_double(step))

    # -- chains -----------------------------------------------------------
    def run_chain(self, target: Target, cfg: OracleConfig, seed, chain,
                  theta0, mass0, step0, n_warmup, n_sampling, rng_policy=0):
        D = target.D
        t = target.c()
        theta0 = np.ascontiguousarray(theta0, np.float64)
        mass0 = np.ascontiguousarray(mass0, np.float64)
        out = dict(
            warmup_draws=np.zeros((n_warmup, D)), warmup_lp=np.zeros(n_warmup),
            warmup_step=np.zeros(n_warmup), warmup_inv_mass=np.zeros((n_warmup, D)),
            warmup_depth=np.zeros(n_warmup, np.int32),
            draws=np.zeros((n_sampling, D)), lp=np.zeros(n_sampling),
            depth=np.zeros(n_sampling, np.int32), inv_mass=np.zeros(D))
        step_out = C.c_double(0)
        mm = C.c_int(0)
        ge = C.c_uint64(0)
        rc = self._fn("run_chain")(
            C.byref(t), C.byref(cfg), C.c_uint32(seed), C.c_uint32(chain),
            C.c_int(rng_policy), _dp(theta0), _dp(mass0), C.c_double(step0),
            C.c_int(n_warmup), C.c_int(n_sampling), _dp(out["warmup_draws"]),
            _dp(out["warmup_lp"]), _dp(out["warmup_step"]),
            _dp(out["warmup_inv_mass"]), _ip(out["warmup_depth"]),
            _dp(out["draws"]), _dp(out["lp"]), _ip(out["depth"]),
            _dp(out["inv_mass"]), C.byref(step_out), C.byref(mm), C.byref(ge))
        self._check(rc)
        out.update(step=step_out.value, min_micro=mm.value, grad_evals=ge.value)
        return out

    def run_sampler(self, target: Target, seed, chain, theta0, inv_mass, step,
                    max_depth, max_halvings, min_micro, max_error, n_iter,
                    rng_policy=0, first_iter=0):
        D = target.D
        t = target.c()
        theta0 = np.ascontiguousarray(theta0, np.float64)
        inv_mass = np.ascontiguousarray(inv_mass, np.float64)
        draws = np.zeros((n_iter, D))
        lp = np.zeros(n_iter)
        depth = np.zeros(n_iter, np.int32)
        ge = C.c_uint64(0)
        rc = self._fn("run_sampler")(
            C.byref(t), C.c_uint32(seed), C.c_uint32(chain), C.c_int(rng_policy),
            C.c_uint32(first_iter), _dp(theta0), _dp(inv_mass), C.c_double(step),
            C.c_int(max_depth), C.c_int(max_halvings), C.c_int(min_micro),
            C.c_double(max_error), C.c_int(n_iter), _dp(draws), _dp(lp),
            _ip(depth), C.byref(ge))
        self._check(rc)
        return dict(draws=draws, lp=lp, depth=depth, grad_evals=ge.value)

    def init_positions(self, num_chains, D, seed, radius):
        pos = np.zeros((num_chains, D))
        self._check(self._fn("init_positions")(
            C.c_size_t(num_chains), C.c_size_t(D), C.c_uint32(seed),
            C.c_double(radius), _dp(pos)))
        return pos

    def init_mass_step(self, target: Target, positions, seed, step_init,
                       mass_in=None, smoothing=1e-5):
        positions = np.ascontiguousarray(positions, np.float64)
        Cn, D = positions.shape
        t = target.c()
        if mass_in is not None:
            mass_in = np.ascontiguousarray(mass_in, np.float64)
        mass = np.zeros((Cn, D))
        steps = np.zeros(Cn)
        self._check(self._fn("init_mass_step")(
            C.byref(t), C.c_size_t(Cn), C.c_uint32(seed), _dp(positions),
            _dp(mass_in), C.c_double(smoothing), C.c_double(step_init),
            _dp(mass), _dp(steps)))
        return mass, steps

    def init_positions_philox(self, num_chains, D, seed, radius, chain_offset=0):
        """The device's position stream (Philox kind 2); oracle only."""
        pos = np.zeros((num_chains, D))
        self._check(self.lib.oracle_init_positions_philox(
            C.c_size_t(num_chains), C.c_size_t(D), C.c_uint32(seed),
            C.c_uint32(chain_offset), C.c_double(radius), _dp(pos)))
        return pos

    def init_mass_step_philox(self, target: Target, positions, seed, step_init,
                              mass_in=None, smoothing=1e-5, chain_offset=0):
        """masses(F, s) + adapt_step with the device's per-chain Philox kind-3
        momentum; oracle only."""
        positions = np.ascontiguousarray(positions, np.float64)
        Cn, D = positions.shape
        t = target.c()
        if mass_in is not None:
            mass_in = np.ascontiguousarray(mass_in, np.float64)
        mass = np.zeros((Cn, D))
        steps = np.zeros(Cn)
        self._check(self.lib.oracle_init_mass_step_philox(
            C.byref(t), C.c_size_t(Cn), C.c_uint32(seed), C.c_uint32(chain_offset),
            _dp(positions), _dp(mass_in), C.c_double(smoothing), C.c_double(step_init),
            _dp(mass), _dp(steps)))
        return mass, steps

    def set_fused_arith(self, fused: bool) -> bool:
        """Arithmetic policy of the accumulate sites (oracle/targets.hpp); oracle only.
        Returns the previous setting."""
        return bool(self.lib.oracle_set_fused_arith(C.c_int(int(fused))))

    @contextlib.contextmanager
    def fused_arith(self, fused: bool = True):
        before = self.set_fused_arith(fused)
        try:
            yield self
        finally:
            self.set_fused_arith(before)

    def warmup_controller(self, log_step, log_mass):
        """(max_m ||(M_m - gm)/gm||_2, max(0, max_m (eps_m - gs)/gs)) of adapt.hpp:186-224
        from per-chain log step sizes [C] and log masses [C][D]; oracle only."""
        log_step = np.ascontiguousarray(log_step, np.float64)
        log_mass = np.ascontiguousarray(log_mass, np.float64)
        Cn, D = log_mass.shape
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.oracle_warmup_controller(
            C.c_size_t(Cn), C.c_size_t(D), _dp(log_step), _dp(log_mass), C.byref(a),
            C.byref(b)))
        return a.value, b.value

    def sampling_rhat(self, mean, var):
        """R-hat of lp from per-chain means and unbiased variances (sampler.hpp:132-151);
        oracle only."""
        mean = np.ascontiguousarray(mean, np.float64)
        var = np.ascontiguousarray(var, np.float64)
        r = C.c_double(0)
        self._check(self.lib.oracle_sampling_rhat(C.c_size_t(mean.size), _dp(mean), _dp(var),
                                                  C.byref(r)))
        return r.value

    def walnuts(self, target: Target, cfg: OracleConfig, seed, positions, mass,
                steps, save_warmup=False):
        positions = np.ascontiguousarray(positions, np.float64)
        mass = np.ascontiguousarray(mass, np.float64)
        steps = np.ascontiguousarray(steps, np.float64)
        Cn, D = positions.shape
        t = target.c()
        ndraw = cfg.max_sampling_iter + (cfg.max_warmup_iter if save_warmup else 0)
        out = np.zeros((Cn, ndraw, D))
        lengths = np.zeros(2 * Cn, np.int32)
        stepsize = np.zeros(Cn)
        inv_metric = np.zeros((Cn, D))
        ge = C.c_uint64(0)
        sw, ss = C.c_double(0), C.c_double(0)
        self._check(self._fn("walnuts")(
            C.byref(t), C.byref(cfg), C.c_size_t(Cn), C.c_uint32(seed),
            _dp(positions), _dp(mass), _dp(steps), C.c_int(int(save_warmup)),
            _dp(out), _ip(lengths), _dp(stepsize), _dp(inv_metric),
            C.byref(ge), C.byref(sw), C.byref(ss)))
        return dict(out=out, warmup_lengths=lengths[:Cn].copy(),
                    sampling_lengths=lengths[Cn:].copy(), stepsize=stepsize,
                    inv_metric=inv_metric, grad_evals=ge.value,
                    seconds_warmup=sw.value, seconds_sampling=ss.value)

    # -- oracle-only ------------------------------------------------------
    def orbit(self, target: Target, theta, rho, inv_mass, step, num_steps, f32=False):
        """f32: the orbit in the device's fp32 mode (float state, double energies)."""
        D = target.D
        t = target.c()
        a = [np.ascontiguousarray(v, np.float64) for v in (theta, rho, inv_mass)]
        th, rh, g = np.zeros(D), np.zeros(D), np.zeros(D)
        lp, jt = C.c_double(0), C.c_double(0)
        self._check((self.lib.oracle_orbit_f32 if f32 else self.lib.oracle_orbit)(
            C.byref(t), _dp(a[0]), _dp(a[1]), _dp(a[2]), C.c_double(step),
            C.c_int(num_steps), _dp(th), _dp(rh), _dp(g), C.byref(lp),
            C.byref(jt)))
        return th, rh, g, lp.value, jt.value

    def logp_grad(self, target: Target, theta):
        t = target.c()
        theta = np.ascontiguousarray(theta, np.float64)
        g = np.zeros(target.D)
        lp = C.c_double(0)
        self._check(self.lib.oracle_logp_grad(C.byref(t), _dp(theta),
                                              C.byref(lp), _dp(g)))
        return lp.value, g

    def philox(self, ctr, key):
        out = (C.c_uint32 * 4)()
        self.lib.oracle_philox(*[C.c_uint32(int(c)) for c in ctr],
                               *[C.c_uint32(int(k)) for k in key], out)
        return [int(x) for x in out]

    def philox_normals(self, seed, chain, it, kind, n):
        out = np.zeros(n)
        self.lib.oracle_philox_normals(C.c_uint32(seed), C.c_uint32(chain),
                                       C.c_uint32(it), C.c_uint32(kind),
                                       C.c_size_t(n), _dp(out))
        return out

    def philox_uniform(self, seed, chain, it, index):
        self.lib.oracle_philox_uniform.restype = C.c_double
        return self.lib.oracle_philox_uniform(C.c_uint32(seed), C.c_uint32(chain),
                                              C.c_uint32(it), C.c_uint32(index))

    def _summary(self, name, chains):
        lengths = np.array([c.shape[0] for c in chains], np.int32)
        stacked = np.ascontiguousarray(np.concatenate(chains, axis=0), np.float64)
        n, d = stacked.shape
        out = np.zeros(stacked.shape if name == "autocovariance" else d)
        self._check(getattr(self.lib, "oracle_" + name)(
            _dp(stacked), C.c_int(n), C.c_int(d), _ip(lengths),
            C.c_int(len(chains)), _dp(out)))
        return out

    def ess(self, chains):
        return self._summary("ess", chains)

    def r_hat(self, chains):
        return self._summary("r_hat", chains)

    def mcse(self, chains):
        return self._summary("mcse", chains)

    def autocovariance(self, chains):
        return self._summary("autocovariance", chains)


def build_oracle(force: bool = False) -> Path:
    """Compile oracle/liboracle.so (building the checker is not using it)."""
    if force or not ORACLE_SO.exists():
        subprocess.run(["make", "-C", str(HERE)] + (["-B"] if force else []),
                       check=True, capture_output=True)
    return ORACLE_SO


def build_ref() -> Optional[Path]:
    """Compile oracle/_ref from the reference sources where they lie; returns
    None where /root/reference is absent and no prebuilt copy travelled."""
    if (REFERENCE / "include" / "walnutpie" / "walnuts.hpp").exists():
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True,
                       capture_output=True)
    return REF_SO if REF_SO.exists() else None


def load_oracle() -> Checker:
    return Checker(build_oracle(), "oracle")


def load_ref() -> Optional[Checker]:
    p = build_ref()
    return Checker(p, "ref") if p else None
