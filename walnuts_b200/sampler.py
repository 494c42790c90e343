"""Python entry points of the device sampler.

``walnuts_device`` keeps the keyword signature and return type of the
reference's ``walnuts_pyfunc`` (python/src/walnutpie/pyfunc.py:45-83, :270-286);
the only change is that ``logp`` is a :class:`DeviceModel` instead of a host
callback.  :class:`Session` exposes the device-resident batch (thousands of
chains kept in HBM) for callers that do not want the draws copied back.
"""
from __future__ import annotations

import contextlib
import ctypes
from typing import Optional

import numpy as np

from . import _ffi
from .models import DeviceModel
from .buffers import (WarmupInfo, prepare_inv_metric, prepare_output_buffer,
                   prepare_seed)


class WalnutsOutputArray(np.ndarray):
    """ndarray of draws with a ``warmup`` attribute (pyfunc.py:10-29)."""

    warmup: WarmupInfo

    def __new__(cls, input_array, warmup: WarmupInfo):
        obj = np.asarray(input_array).view(cls)
        obj.warmup = warmup
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.warmup = getattr(obj, "warmup", None)


def make_tuning(**kw) -> _ffi.WalnutTuning:
    t = _ffi.WalnutTuning()
    _ffi._lib.walnuts_b200_default_tuning(ctypes.byref(t))
    for k, v in kw.items():
        if not hasattr(t, k):
            raise TypeError(f"unknown tuning argument {k!r}")
        setattr(t, k, v)
    return t


@contextlib.contextmanager
def _reraise_callback_errors(model):
    errors = getattr(model, "errors", None)
    if errors:
        errors.clear()
    try:
        yield
    except RuntimeError as failure:
        # only a failure during initialisation ends a run (walnutpy.cpp:162-170); inside a
        # transition the chains go on with logp = -inf and the exception is only printed
        if errors and str(failure).startswith("logp failed with code"):
            exc = errors[0]
            errors.clear()
            raise exc from failure
        raise


def walnuts_device(
    logp: DeviceModel,
    *,
    num_params: Optional[int] = None,
    inits: Optional[np.ndarray] = None,
    num_chains: int = 4,
    seed: Optional[int] = None,
    id: int = 1,
    init_radius: float = 2.0,
    init_inv_metric: Optional[np.ndarray] = None,
    save_inv_metric: bool = False,
    min_warmup_iter: int = 50,
    max_warmup_iter: int = 1000,
    min_sampling_iter: int = 50,
    max_sampling_iter: int = 1000,
    max_trajectory_doublings: int = 5,
    max_step_halvings: int = 5,
    min_micro_steps: int = 1,
    max_hamiltonian_error: float = 0.5,
    step_size_converge_tol: float = 0.1,
    mass_converge_tol: float = 1.0,
    rhat_converge_tol: float = 1.01,
    mass_init_count: float = 4.0,
    mass_additive_smoothing: float = 1e-5,
    max_macro_steps_target: float = 15.0,
    step_size_init: float = 1.0,
    step_accept_rate_target: float = 0.8,
    step_learning_rate: float = 0.05,
    step_gradient_decay: float = 0.8,
    step_sq_gradient_decay: float = 0.9,
    step_stabilization: float = 1e-4,
    step_learn_rate_decay: float = 0.5,
    save_warmup: bool = False,
    refresh: int = 0,
) -> list[WalnutsOutputArray]:
    """Sample ``logp`` (a device model) with ``num_chains`` chains on the GPU.

    Arguments, defaults, validation errors and the returned list of
    :class:`WalnutsOutputArray` follow ``walnuts_pyfunc`` (pyfunc.py:45-286).
    Differences: (i) ``logp`` is a :class:`DeviceModel`; (ii) with ``min < max`` the
    chains run free like the reference's threads -- by equal
    work (gradient evaluations) per block, so final lengths differ by chain as in the
    reference (docs/py.rst:13-20) but are reproducible; ``WB200_BLOCKS=uniform`` stops all
    chains at the same iteration, decided every ``publish_stride`` (5) iterations;
    (iii) random numbers come
    from a counter-based Philox stream keyed by ``(seed + id + num_chains,
    chain)``, not from ``std::mt19937_64``.
    """
    if not isinstance(logp, DeviceModel):
        raise TypeError(
            "walnuts_b200 samples device models only; pass a walnuts_b200.models."
            "DeviceModel (a host callback cannot feed a GPU batch)")
    if num_params is None:
        num_params = logp.num_params
    if num_params != logp.num_params:
        raise ValueError("num_params does not match the model")

    seed = prepare_seed(seed)
    out = prepare_output_buffer(num_chains=num_chains, num_params=num_params,
                                max_sampling_iter=max_sampling_iter,
                                max_warmup_iter=max_warmup_iter,
                                save_warmup=save_warmup, pinned=True)
    if inits is not None:
        inits = np.ascontiguousarray(inits, dtype=np.float64)
        if inits.shape == (num_params,):
            inits = np.ascontiguousarray(
                np.repeat(inits[np.newaxis], num_chains, axis=0))
        elif inits.shape == (num_chains, num_params):
            pass
        else:
            raise ValueError(
                f"Invalid inits size. Expected a {(num_params,)} "
                f"or {(num_chains, num_params)} matrix.")
    init_inv_metric = prepare_inv_metric(init_inv_metric, (num_params,), num_chains)

    lengths_out = np.zeros((num_chains * 2,), dtype=np.int32)
    stepsize_out = np.zeros(num_chains, dtype=np.float64)
    inv_metric_out = None
    if save_inv_metric:
        inv_metric_out = np.zeros((num_chains, num_params), dtype=np.float64)

    desc = logp.desc()
    with _reraise_callback_errors(logp):
        _ffi._ffi_sample_device(
            ctypes.byref(desc), num_params, inits, num_chains, seed, id, init_radius,
            init_inv_metric, min_warmup_iter, max_warmup_iter, min_sampling_iter,
            max_sampling_iter, max_trajectory_doublings, max_step_halvings,
            min_micro_steps, max_hamiltonian_error, step_size_converge_tol,
            mass_converge_tol, rhat_converge_tol, mass_init_count,
            mass_additive_smoothing, max_macro_steps_target, step_size_init,
            step_accept_rate_target, step_learning_rate, step_gradient_decay,
            step_sq_gradient_decay, step_stabilization, step_learn_rate_decay,
            save_warmup, out, out.size, lengths_out, stepsize_out, inv_metric_out,
            refresh, _ffi.print_callback)

    outputs = []
    for i in range(num_chains):
        warmup_written = lengths_out[i]
        samples_written = lengths_out[i + num_chains]
        warmup_info = WarmupInfo(
            stepsize=stepsize_out[i],
            inv_metric=inv_metric_out[i] if inv_metric_out is not None else None,
            warmup_draws=(out[i, 0:warmup_written, :] if save_warmup else None))
        outputs.append(WalnutsOutputArray(
            out[i, warmup_written:warmup_written + samples_written, :], warmup_info))
    return outputs


def walnuts_device_summary(logp: DeviceModel, *, num_chains: int = 4, seed: Optional[int] = None,
                           id: int = 1, inits: Optional[np.ndarray] = None,
                           init_radius: float = 2.0,
                           init_inv_metric: Optional[np.ndarray] = None, max_lags: int = 32,
                           refresh: int = 0, devices=None, **tuning):
    """``walnuts_device`` for runs whose draws are too many to keep: the same sampler run
    (``walnutpie_sample_device_summary``), returning the posterior summaries computed on
    the device by streaming accumulators -- ``mean``, ``variance``, ``r_hat``, ``ess``,
    ``mcse`` per parameter (summary.hpp:371-405,594-769), ``truncated`` flags, per-chain
    ``stepsize`` / ``inv_metric`` and the iteration counts.  ``tuning`` takes the tuning
    keywords of ``walnuts_device`` (reference defaults).  ``devices``: a list of CUDA device
    ordinals -- the chains are sharded over them inside the one call
    (``walnutpie_sample_device_multi``: one host thread per GPU, controllers and summaries
    all-reduced with NCCL); the results do not depend on the number of devices."""
    if not isinstance(logp, DeviceModel):
        raise TypeError("walnuts_b200 samples device models only")
    D = logp.num_params
    t = make_tuning(**tuning)
    seed = prepare_seed(seed)
    if inits is not None:
        inits = np.ascontiguousarray(inits, dtype=np.float64)
        if inits.shape == (D,):
            inits = np.ascontiguousarray(np.repeat(inits[np.newaxis], num_chains, axis=0))
        elif inits.shape != (num_chains, D):
            raise ValueError(f"Invalid inits size. Expected a {(D,)} or "
                             f"{(num_chains, D)} matrix.")
    init_inv_metric = prepare_inv_metric(init_inv_metric, (D,), num_chains)
    out = {k: np.zeros(D) for k in ("mean", "variance", "r_hat", "ess", "mcse")}
    cut = np.zeros(D, np.int32)
    lengths = np.zeros(2 * num_chains, np.int32)
    stepsize = np.zeros(num_chains)
    inv_metric = np.zeros((num_chains, D))
    desc = logp.desc()
    if devices is not None:
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        call = lambda *args: _ffi._ffi_sample_device_multi(dev, len(dev), *args)  # noqa: E731
    else:
        call = _ffi._ffi_sample_device_summary
    with _reraise_callback_errors(logp):
        call(
            ctypes.byref(desc), D, inits, num_chains, seed, id, init_radius, init_inv_metric,
            t.min_warmup_iter, t.max_warmup_iter, t.min_sampling_iter, t.max_sampling_iter,
            t.max_trajectory_doublings, t.max_step_halvings, t.min_micro_steps,
            t.max_hamiltonian_error, t.step_size_converge_tol, t.mass_converge_tol,
            t.rhat_converge_tol, t.mass_init_count, t.mass_additive_smoothing,
            t.max_macro_steps_target, t.step_size_init, t.step_accept_rate_target,
            t.step_learning_rate, t.step_gradient_decay, t.step_sq_gradient_decay,
            t.step_stabilization, t.step_learn_rate_decay, int(max_lags), out["mean"],
            out["variance"], out["r_hat"] if num_chains > 1 else None, out["ess"],
            out["mcse"], cut, lengths, stepsize, inv_metric, refresh, _ffi.print_callback)
    # final_lengths reports SAVED warm-up draws (0 here); the iterations run are in the stats
    out.update(truncated=cut, stepsize=stepsize, inv_metric=inv_metric,
               warmup_iters=(_ffi.last_run_stats()["warmup_iters"] if devices is None
                             else None),
               sampling_iters=int(lengths[num_chains:].max()),
               sampling_lengths=lengths[num_chains:].astype(np.int64))
    return out


class Session:
    """A device-resident batch of chains (include/walnuts_b200.h, session API).

    The stages of ``walnutpie::walnuts`` (api.hpp:33-69), batched:
    ``init`` -> ``warmup(n)`` ... -> ``freeze()`` -> ``sample(n)`` ...
    """

    def __init__(self, model: DeviceModel, num_chains: int, seed: int = 0,
                 chain_offset: int = 0, device: int = 0, **tuning):
        self.model = model
        self.num_chains = int(num_chains)
        self.num_params = int(model.num_params)
        self._desc = model.desc()
        self.tuning = make_tuning(**tuning)
        self._h = _ffi.session_p()
        _ffi.session_create(ctypes.byref(self._desc), self.num_chains, seed,
                            chain_offset, ctypes.byref(self.tuning), device,
                            ctypes.byref(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _ffi.session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def _c(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.float64)

    def _run(self, fn, *args):
        """Calls that evaluate the model: an exception raised inside a batch-callback
        density comes back as the library's runtime error and is re-raised as itself."""
        with _reraise_callback_errors(self.model):
            fn(self._h, *args)
        return self

    def init(self, positions=None, init_radius: float = 2.0, mass=None, steps=None):
        return self._run(_ffi.session_init, self._c(positions), float(init_radius),
                         self._c(mass), self._c(steps))

    def reserve(self, capacity: int, trace: bool = False):
        _ffi.session_reserve(self._h, int(capacity), int(trace))
        return self

    def warmup(self, n_iter: int, store: bool = False):
        return self._run(_ffi.session_warmup, int(n_iter), int(store))

    def freeze(self):
        _ffi.session_freeze(self._h)
        return self

    def sample(self, n_iter: int, store: bool = True):
        return self._run(_ffi.session_sample, int(n_iter), int(store))

    def sync(self):
        _ffi.session_sync(self._h)
        return self

    def sample_ticks(self, n_ticks: int, store: bool = True):
        """Free-running sampling: ``n_ticks`` gradient evaluations per chain; chains
        complete as many transitions as fit (ragged draw counts, like the reference's
        per-thread chains, sampler.hpp:79-94).  Lock-step sessions: exactly ``n_ticks``
        ticks, transitions in flight carry over.  Chain-resident sessions: the transition
        that exhausts the budget is finished and its excess comes off the next budget."""
        return self._run(_ffi.session_sample_ticks, int(n_ticks), int(store))

    def warmup_ticks(self, n_ticks: int, store: bool = False):
        """Free-running adaptive warm-up: ``n_ticks`` gradient evaluations per chain; every
        chain adapts over as many transitions as fit.  On lock-step sessions ``freeze``
        abandons the transitions in flight."""
        return self._run(_ffi.session_warmup_ticks, int(n_ticks), int(store))

    def run_evals(self, eval_budget: int, *, sampling: bool, iter_cap: int = 0,
                  store: bool = True):
        """One free-running launch of ``eval_budget`` gradient evaluations per chain (lock-step
        sessions: that many ticks) in which no chain exceeds ``iter_cap`` iterations of the
        phase in total (0: no limit) -- a reference chain stops at ``max_iter``."""
        _ffi.session_run_evals(self._h, int(sampling), int(eval_budget), int(iter_cap),
                               int(store))
        return self

    def iter_stats(self, *, sampling: bool) -> tuple:
        """(min, max, sum) over the chains of the phase's per-chain iteration counts, and
        the gradient evaluations of all chains since initialisation."""
        out = np.zeros(4, np.int64)
        _ffi.session_iter_stats(self._h, int(sampling), out)
        return int(out[0]), int(out[1]), int(out[2]), int(out[3])

    def chain_rows(self) -> np.ndarray:
        rows = np.zeros(self.num_chains, np.int64)
        _ffi.session_chain_rows(self._h, rows)
        return rows

    def rhat_moments(self, first: int = 0) -> np.ndarray:
        """[3*D + 1] per-dimension chain-moment sums + chain count: all-reduce (SUM) over
        ranks, then walnuts_b200.distributed.rhat_from_dimension_moments."""
        out = np.zeros(3 * self.num_params + 1)
        _ffi.session_rhat_moments(self._h, int(first), out)
        return out

    def summary_ragged(self, first: int = 0):
        """R-hat / ESS / MCSE / mean / variance over rows [first, rows_c) of every chain."""
        D = self.num_params
        out = {k: np.zeros(D) for k in ("r_hat", "ess", "mcse", "mean", "variance")}
        _ffi.session_summary(self._h, int(first),
                             out["r_hat"] if self.num_chains > 1 else None, out["ess"],
                             out["mcse"], out["mean"], out["variance"])
        return out

    def draws(self, first: int, count: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Stored draws [chains][count][D] on the host (``out``: a caller-owned, e.g.
        pinned, buffer).  After a free-running phase use ``chain_rows`` for the number of
        valid rows of each chain."""
        if out is None:
            out = np.zeros((self.num_chains, count, self.num_params))
        elif out.shape != (self.num_chains, count, self.num_params):
            raise ValueError("out has the wrong shape")
        _ffi.session_get_draws(self._h, first, count, out)
        return out

    def trace(self, first: int, count: int):
        C, D = self.num_chains, self.num_params
        lp = np.zeros((C, count))
        depth = np.zeros((C, count), np.int32)
        step = np.zeros((C, count))
        im = np.zeros((C, count, D))
        _ffi.session_get_trace(self._h, first, count, lp, depth, step, im)
        return dict(lp=lp, depth=depth, step=step, inv_mass=im)

    def state(self):
        C, D = self.num_chains, self.num_params
        theta = np.zeros((C, D))
        im = np.zeros((C, D))
        step = np.zeros(C)
        mm = np.zeros(C, np.int32)
        ge = np.zeros(C, np.uint64)
        _ffi.session_get_state(self._h, theta, im, step, mm, ge.ctypes.data)
        return dict(theta=theta, inv_mass=im, step=step, min_micro=mm, grad_evals=ge)

    def counters(self):
        g, m, l = (ctypes.c_ulonglong(0) for _ in range(3))
        _ffi.session_counters(self._h, ctypes.byref(g), ctypes.byref(m), ctypes.byref(l))
        return dict(grad_evals=g.value, macro_steps=m.value, kernel_launches=l.value)

    def last_kernel_ms(self) -> float:
        ms = ctypes.c_float(0)
        if _ffi.session_last_kernel_ms(self._h, ctypes.byref(ms)) != 0:
            raise RuntimeError("no kernel has been timed on this session")
        return ms.value

    def timer_start(self):
        _ffi.session_timer_record(self._h, 0)

    def timer_stop_ms(self) -> float:
        _ffi.session_timer_record(self._h, 1)
        ms = ctypes.c_float(0)
        _ffi.session_timer_elapsed_ms(self._h, ctypes.byref(ms))
        return ms.value

    def device_draws(self):
        """(device pointer, capacity, ld, rows_written) of the draw buffer."""
        p = ctypes.c_void_p()
        cap, rows = ctypes.c_longlong(0), ctypes.c_longlong(0)
        ld = ctypes.c_int(0)
        _ffi.session_device_draws(self._h, ctypes.byref(p), ctypes.byref(cap),
                                  ctypes.byref(ld), ctypes.byref(rows))
        return p.value, cap.value, ld.value, rows.value

    def summary(self, first: int, count: int):
        """Per-dimension R-hat, ESS, MCSE, mean, variance of stored draws,
        computed on the device (summary.hpp:594-769)."""
        ptr, cap, ld, _ = self.device_draws()
        D = self.num_params
        out = {k: np.zeros(D) for k in ("r_hat", "ess", "mcse", "mean", "variance")}
        self.sync()
        _ffi.device_summary(ptr, self.num_chains, cap, first, count, D, ld,
                            out["r_hat"] if self.num_chains > 1 else None,
                            out["ess"], out["mcse"], out["mean"], out["variance"])
        return out

    # -- streaming summaries (include/walnuts_b200.h) ---------------------------------
    def stream_begin(self, max_lags: int = 32):
        """From now on storing ``sample`` / ``sample_ticks`` calls fold their draws into
        per-chain running sums instead of keeping them: the reserved draw buffer becomes
        a staging block (``reserve(block)`` first)."""
        _ffi.session_stream_begin(self._h, int(max_lags))
        self._stream_lags = int(max_lags)
        return self

    def stream_counts(self) -> np.ndarray:
        n = np.zeros(self.num_chains, np.int64)
        _ffi.session_stream_counts(self._h, n)
        return n

    def stream_phase1(self) -> np.ndarray:
        out = np.zeros(2 * self.num_params + 3)
        _ffi.session_stream_phase1(self._h, out)
        return out

    def stream_phase2(self, reduced1: np.ndarray) -> np.ndarray:
        out = np.zeros((3 + self._stream_lags) * self.num_params)
        _ffi.session_stream_phase2(self._h, np.ascontiguousarray(reduced1, np.float64), out)
        return out

    def stream_summary(self):
        """R-hat, ESS, MCSE, pooled mean and variance of everything streamed so far
        (summary.hpp:594-769), plus ``truncated`` flags per dimension."""
        D = self.num_params
        out = {k: np.zeros(D) for k in ("r_hat", "ess", "mcse", "mean", "variance")}
        cut = np.zeros(D, np.int32)
        _ffi.session_stream_summary(self._h, out["r_hat"] if self.num_chains > 1 else None,
                                    out["ess"], out["mcse"], out["mean"], out["variance"], cut)
        out["truncated"] = cut
        return out

    def warmup_deviation(self, sums_device_ptr: int):
        out = np.zeros(2)
        _ffi.session_warmup_deviation(self._h, sums_device_ptr, out)
        return out

    def warmup_sums(self, sums_device_ptr: int):
        _ffi.session_warmup_sums(self._h, sums_device_ptr)

    def lp_moments(self, center: Optional[float] = None):
        """{sum mu, sum mu^2, sum var, chains} of the per-chain Welford moments of lp
        (sampler.hpp:132-151); with ``center`` the first two are taken about it (the
        second pass of util.hpp:401-404)."""
        out = np.zeros(4)
        if center is None:
            _ffi.session_lp_moments(self._h, out)
        else:
            _ffi.session_lp_moments_centered(self._h, float(center), out)
        return out

    def logp_exceptions(self) -> int:
        """Batched density evaluations that failed inside a transition and were replaced
        by logp = -inf, grad = 0 (util.hpp:336-346)."""
        n = ctypes.c_ulonglong(0)
        _ffi.session_logp_exceptions(self._h, ctypes.byref(n))
        return n.value


def stream_finish(num_params: int, max_lags: int, reduced1, reduced2, want_rhat: bool = True):
    """The summaries from the (all-reduced) payloads of ``Session.stream_phase1`` /
    ``stream_phase2``; identical on every rank."""
    D = int(num_params)
    out = {k: np.zeros(D) for k in ("r_hat", "ess", "mcse", "mean", "variance")}
    cut = np.zeros(D, np.int32)
    _ffi.stream_finish(D, int(max_lags), np.ascontiguousarray(reduced1, np.float64),
                       np.ascontiguousarray(reduced2, np.float64),
                       out["r_hat"] if want_rhat else None, out["ess"], out["mcse"],
                       out["mean"], out["variance"], cut)
    out["truncated"] = cut
    return out


def orbit(model: DeviceModel, theta, rho, inv_mass, step: float, num_steps: int):
    """``num_steps`` leapfrog micro-steps (walnuts.hpp:329-332) for a batch."""
    theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    rho = np.ascontiguousarray(np.atleast_2d(rho), dtype=np.float64)
    im = np.ascontiguousarray(np.atleast_2d(inv_mass), dtype=np.float64)
    C, D = theta.shape
    desc = model.desc()
    th, rh, g = np.zeros((C, D)), np.zeros((C, D)), np.zeros((C, D))
    lp, jt = np.zeros(C), np.zeros(C)
    _ffi.orbit(ctypes.byref(desc), C, theta, rho, im, float(step), int(num_steps),
               th, rh, g, lp, jt)
    return th, rh, g, lp, jt


def logistic_logp_grad(X, y, theta, repeats: int = 0):
    """Batched logistic-regression log density and gradient for C parameter vectors
    on the tensor cores; returns (logp [C], grad [C][D], ms per evaluation or None)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    N, D = X.shape
    C = theta.shape[0]
    lp, g = np.zeros(C), np.zeros((C, D))
    ms = ctypes.c_float(0)
    _ffi.logistic_logp_grad(X, y, N, D, theta, C, lp, g, int(repeats), ctypes.byref(ms))
    return lp, g, (ms.value if repeats > 0 else None)
