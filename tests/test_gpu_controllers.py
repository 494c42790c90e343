"""Value-level parity of what surrounds the transition on the device: the batched
initialisation, the two cross-chain controllers, the kind-4 failure semantics and the
per-slot look-ahead cache -- all through the C-ABI, against the oracle.

Reference: config.hpp:259-268,360-382,469-476 and util.hpp:285-303 (initialisation),
adapt.hpp:186-224 (warm-up controller), sampler.hpp:132-151 (sampling controller),
util.hpp:336-346 (NoExceptLogpGrad), summary.hpp:594-619 (per-dimension R-hat).
"""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle.binding import Target, default_config

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _device_arithmetic_policy(oracle):
    """The shipped kernels use the fused arithmetic policy (chain_kernel.cuh, kFusedArith);
    the oracle is switched to the same policy for every comparison in this module."""
    with oracle.fused_arith(True):
        yield
ROOT = Path(__file__).resolve().parent.parent


def make(wb, kind, D):
    if kind == "diag_gaussian":
        var = 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
        return wb.models.diag_gaussian(var), Target(kind, D, prec=1 / var)
    return getattr(wb.models, kind)(D), Target(kind, D)


# ---- device initialisation -------------------------------
@pytest.mark.parametrize("engine", ["chain", "tick"])
@pytest.mark.parametrize("kind,D,C,radius,step_init", [
    ("std_normal", 100, 48, 2.0, 1.0),
    ("diag_gaussian", 1000, 24, 2.0, 1.0),     # c2 shape: 128-thread groups
    ("diag_gaussian", 37, 40, 0.7, 1e-4),      # odd D, search has to double 14 times
    ("funnel", 11, 64, 1.0, 100.0),            # search has to shrink
    ("funnel", 100, 32, 1.0, 1.0),             # c3 shape
])
def test_device_init_equals_oracle_philox_policy(wb, oracle, monkeypatch, engine, kind, D, C,
                                                 radius, step_init):
    """Session.init() with positions = mass = steps = None: positions from Philox kind 2,
    mass = (1-s)|grad| + s, step from the doubling / sqrt(1/2) search on Philox kind-3
    momenta -- against the oracle's restatement fed the same streams."""
    if engine == "tick":
        monkeypatch.setenv("WB200_ENGINE", "tick")
    model, target = make(wb, kind, D)
    seed, off, smooth = 4242, 7, 1e-5
    with wb.Session(model, C, seed=seed, chain_offset=off, step_size_init=step_init,
                    mass_additive_smoothing=smooth) as s:
        s.init(init_radius=radius)
        st = s.state()   # before freeze(): theta0, the initial masses, exp(log step)
    pos = oracle.init_positions_philox(C, D, seed, radius, chain_offset=off)
    # same Philox words; the Box-Muller log / sincos differ between CUDA and glibc by ulps
    np.testing.assert_allclose(st["theta"], pos, rtol=1e-13, atol=1e-15)
    # masses and steps from the DEVICE's positions: element-wise arithmetic is identical
    mass, steps = oracle.init_mass_step_philox(target, st["theta"], seed, step_init,
                                               smoothing=smooth, chain_offset=off)
    if kind == "funnel":   # d/dv sums over the other coordinates (summation order)
        np.testing.assert_allclose(st["inv_mass"], mass, rtol=1e-12)
    else:
        np.testing.assert_array_equal(st["inv_mass"], mass)
    # the search multiplies by 2 or sqrt(1/2): a different decision is a factor, not an ulp
    np.testing.assert_allclose(st["step"], steps, rtol=1e-12)
    assert len(np.unique(np.round(np.log2(steps) * 2))) > 1 or kind == "std_normal"


def test_device_init_given_mass_skips_the_gradient_rule(wb, oracle):
    D, C = 20, 16
    model, target = make(wb, "diag_gaussian", D)
    rng = np.random.default_rng(0)
    mass_in = rng.uniform(0.5, 3.0, (C, D))
    with wb.Session(model, C, seed=5) as s:
        s.init(init_radius=1.5, mass=mass_in)
        st = s.state()
    np.testing.assert_array_equal(st["inv_mass"], mass_in)
    _, steps = oracle.init_mass_step_philox(target, st["theta"], 5, 1.0, mass_in=mass_in)
    np.testing.assert_allclose(st["step"], steps, rtol=1e-12)


# ---- controllers ------------------------------------------------
@pytest.mark.parametrize("engine", ["chain", "tick"])
@pytest.mark.parametrize("kind,D,C", [("diag_gaussian", 40, 96), ("funnel", 11, 64),
                                      ("diag_gaussian", 300, 12)])
def test_warmup_controller_statistics_equal_the_reference_math(wb, oracle, monkeypatch,
                                                               engine, kind, D, C):
    """warmup_sums / warmup_deviation after a real warm-up against adapt.hpp:190-223
    evaluated by the oracle on the very chains' (log step, log mass)."""
    import torch
    if engine == "tick":
        monkeypatch.setenv("WB200_ENGINE", "tick")
    model, _ = make(wb, kind, D)
    with wb.Session(model, C, seed=99, max_step_halvings=7) as s:
        s.init(init_radius=1.0)
        s.warmup(23)
        sums = torch.zeros(D + 2, dtype=torch.float64, device="cuda")
        s.warmup_sums(sums.data_ptr())
        dev = s.warmup_deviation(sums.data_ptr())
        s.freeze()       # inv_mass <- sqrt(var_draws / var_scores), step <- exp(adam_x)
        st = s.state()
    sums = sums.cpu().numpy()
    log_mass = -np.log(st["inv_mass"])          # adaptive_walnuts.hpp:320-323
    log_step = np.log(st["step"])               # :312
    np.testing.assert_allclose(sums[:D], log_mass.sum(0), rtol=1e-12, atol=1e-12)
    assert sums[D] == pytest.approx(log_step.sum(), rel=1e-12)
    assert sums[D + 1] == C
    max_mass, max_step = oracle.warmup_controller(log_step, log_mass)
    assert dev[0] == pytest.approx(max_mass, rel=1e-12)
    assert dev[1] == pytest.approx(max_step, rel=1e-10, abs=1e-14)
    assert max_mass > 0 and np.isfinite(max_mass)


@pytest.mark.parametrize("engine", ["chain", "tick"])
def test_sampling_controller_statistics_equal_the_reference_math(wb, oracle, monkeypatch,
                                                                 engine):
    """lp_moments (plain and centred) and rhat_moments after real sampling against the
    per-chain Welford statistics of the lp trace and sampler.hpp:132-151 /
    summary.hpp:594-619 evaluated by the oracle."""
    if engine == "tick":
        monkeypatch.setenv("WB200_ENGINE", "tick")
    D, C, n = 30, 40, 37
    model, _ = make(wb, "diag_gaussian", D)
    with wb.Session(model, C, seed=3) as s:
        s.init(init_radius=2.0)
        s.reserve(n, trace=True)
        s.warmup(40).freeze().sample(n).sync()
        m = s.lp_moments()
        lp = s.trace(0, n)["lp"]
        center = m[0] / m[3]
        mc = s.lp_moments(center)
        dm = s.rhat_moments(0)
        draws = s.draws(0, n)
    mu, var = lp.mean(1), lp.var(1, ddof=1)
    assert m[3] == C and mc[3] == C
    assert m[0] == pytest.approx(mu.sum(), rel=1e-12)
    assert m[1] == pytest.approx((mu ** 2).sum(), rel=1e-12)
    assert m[2] == pytest.approx(var.sum(), rel=1e-10)
    assert mc[1] == pytest.approx(((mu - center) ** 2).sum(), rel=1e-9)
    rhat = np.sqrt(1 + (mc[1] - mc[0] ** 2 / C) / (C - 1) / (mc[2] / C))
    assert rhat == pytest.approx(oracle.sampling_rhat(mu, var), rel=1e-10)
    # per-dimension payload of the NCCL R-hat
    cm, cv = draws.mean(1), draws.var(1, ddof=1)
    np.testing.assert_allclose(dm[:D], cm.sum(0), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dm[D:2 * D], (cm ** 2).sum(0), rtol=1e-10)
    np.testing.assert_allclose(dm[2 * D:3 * D], cv.sum(0), rtol=1e-10)
    assert dm[3 * D] == C
    from walnuts_b200.distributed import rhat_from_dimension_moments
    np.testing.assert_allclose(rhat_from_dimension_moments(dm),
                               oracle.r_hat([draws[c] for c in range(C)]), rtol=1e-9)


def test_one_shot_stop_decisions_equal_a_replay_of_the_controllers(wb, oracle, monkeypatch):
    """walnutpie_sample_device with min < max and blocks of equal iteration counts
    (WB200_BLOCKS=uniform; the free-running default is replayed in test_gpu_ragged.py):
    the iteration at which the library stops warm-up and sampling equals the first
    publish_stride boundary at which the reference's tests (adapt.hpp:218-219,
    sampler.hpp:147-148) pass on the same chains, replayed block by block on a Session
    with the oracle evaluating the statistics."""
    monkeypatch.setenv("WB200_BLOCKS", "uniform")
    D, C = 16, 32
    model, _ = make(wb, "diag_gaussian", D)
    kw = dict(min_warmup_iter=20, max_warmup_iter=400, min_sampling_iter=20,
              max_sampling_iter=300, mass_converge_tol=0.9, step_size_converge_tol=0.35,
              rhat_converge_tol=1.01)
    seed, ident = 17, 1
    fit = wb.walnuts_device(model, num_chains=C, seed=seed, id=ident, save_warmup=True,
                            save_inv_metric=True, **kw)
    warm_len = len(fit[0].warmup.warmup_draws)
    samp_len = len(fit[0])
    assert all(len(f.warmup.warmup_draws) == warm_len and len(f) == samp_len for f in fit)
    # replay: the one-shot call keys its streams by seed + id + num_chains (walnutpy.cpp:82)
    tune = {k: v for k, v in kw.items() if "iter" not in k}
    with wb.Session(model, C, seed=seed + ident + C, **tune, **{k: v for k, v in kw.items()
                                                               if "iter" in k}) as s:
        s.init(init_radius=2.0)
        s.reserve(kw["max_sampling_iter"], trace=True)
        done = 0
        while done < kw["max_warmup_iter"]:
            s.warmup(5)
            done += 5
            if done >= kw["min_warmup_iter"] and done < kw["max_warmup_iter"]:
                # the statistics of the adapting chains: estimator state through a frozen copy
                import torch
                sums = torch.zeros(D + 2, dtype=torch.float64, device="cuda")
                s.warmup_sums(sums.data_ptr())
                dm, ds = s.warmup_deviation(sums.data_ptr())
                if dm <= kw["mass_converge_tol"] and ds <= kw["step_size_converge_tol"]:
                    break
        assert done == warm_len
        assert kw["min_warmup_iter"] < done < kw["max_warmup_iter"], "case must stop early"
        s.freeze()
        st = s.state()
        n = 0
        while n < kw["max_sampling_iter"]:
            s.sample(5)
            n += 5
            if n >= kw["min_sampling_iter"] and n < kw["max_sampling_iter"]:
                lp = s.trace(0, n)["lp"]
                if oracle.sampling_rhat(lp.mean(1), lp.var(1, ddof=1)) <= kw["rhat_converge_tol"]:
                    break
        assert n == samp_len
        assert kw["min_sampling_iter"] <= n < kw["max_sampling_iter"], "case must stop early"
        draws = s.draws(0, n)
    for c in range(C):
        np.testing.assert_array_equal(np.asarray(fit[c]), draws[c])
        assert fit[c].warmup.stepsize == st["step"][c]
        np.testing.assert_array_equal(fit[c].warmup.inv_metric, st["inv_mass"][c])


# ---- the per-slot look-ahead cache of log(u) ---------------------
def test_single_iteration_launches_equal_one_long_launch_with_more_chains_than_slots(
        wb, oracle):
    """20 000 chains on a few thousand resident slots: every slot takes several chains per
    launch.  k launches of 1 iteration must equal one launch of k iterations and the
    oracle -- the look-ahead cache of merge uniforms is per chain, not per slot."""
    D, C, k = 4, 20000, 6
    model, target = make(wb, "std_normal", D)
    rng = np.random.default_rng(1)
    pos = rng.normal(size=(C, D))
    mass = np.ones((C, D))
    steps = np.full(C, 0.6)
    out = []
    for blocks in ([k], [1] * k, [2, 1, 3]):
        with wb.Session(model, C, seed=8, max_trajectory_doublings=6) as s:
            s.init(positions=pos, mass=mass, steps=steps)
            s.reserve(2 * k)
            for b in blocks:
                s.warmup(b, store=True)
            s.freeze()
            for b in blocks:
                s.sample(b)
            s.sync()
            out.append(s.draws(0, 2 * k))
    np.testing.assert_array_equal(out[1], out[0])
    np.testing.assert_array_equal(out[2], out[0])
    cfg = default_config(max_trajectory_doublings=6)
    for c in (0, 1, 4097, 9999, C - 1):
        o = oracle.run_chain(target, cfg, 8, c, pos[c], mass[c], steps[c], k, k, rng_policy=1)
        ref = np.concatenate([o["warmup_draws"], o["draws"]])
        np.testing.assert_allclose(out[0][c], ref, rtol=1e-9, atol=1e-12)


# ---- kind 4: a failing batched density ---------------------------------
def test_failing_batched_density_continues_with_minus_infinity(wb, capsys):
    """util.hpp:336-346 for a batch: a non-zero return inside a transition gives every
    chain logp = -inf and a zero gradient for that evaluation, the run goes on, the
    failure is counted and printed; a failure during initialisation ends the run
    (walnutpy.cpp:162-170)."""
    import torch
    D, C = 3, 8
    calls = {"n": 0, "fail_from": 10 ** 9, "fail_to": 10 ** 9}

    def density(Cn, Dn, ld, theta, grad, lp, stream):
        calls["n"] += 1
        if calls["fail_from"] <= calls["n"] < calls["fail_to"]:
            raise ArithmeticError("density blew up")
        with torch.cuda.stream(torch.cuda.ExternalStream(int(stream or 0))):
            t = torch.as_tensor(wb.models._DevicePointer(theta, (Cn, ld)), device="cuda")
            g = torch.as_tensor(wb.models._DevicePointer(grad, (Cn, ld)), device="cuda")
            l = torch.as_tensor(wb.models._DevicePointer(lp, (Cn,)), device="cuda")
            g.copy_(-t)
            l.copy_(-0.5 * (t[:, :Dn] * t[:, :Dn]).sum(dim=1))

    model = wb.models.batch_callback(D, density)
    with wb.Session(model, C, seed=2) as s:
        s.init(init_radius=1.0)
        s.reserve(40)
        s.warmup(10).freeze()
        assert s.logp_exceptions() == 0
        s.sample(10).sync()
        good = s.draws(0, 10).copy()
        # every evaluation of the next 5 transitions fails: each transition's first leaf has
        # an infinite energy error at every rung, the extension is rejected
        # (walnuts.hpp:339-345, :543-545) and the chain stays where it is
        calls["fail_from"], calls["fail_to"] = calls["n"] + 1, 10 ** 9
        s.sample(5).sync()
        stuck = s.draws(10, 5)
        n_fail = s.logp_exceptions()
        assert n_fail > 0
        for c in range(C):
            np.testing.assert_array_equal(stuck[c], np.tile(good[c, -1], (5, 1)))
        # the density recovers: the chains move again
        calls["fail_to"] = calls["n"] + 1
        s.sample(10).sync()
        after = s.draws(15, 10)
        assert s.logp_exceptions() == n_fail
        assert np.all(np.isfinite(after))
        assert np.all(np.any(after[:, -1] != good[:, -1], axis=1))
    assert "density blew up" in capsys.readouterr().out      # pyfunc.py:38-40 prints it
    # initialisation is not wrapped by NoExceptLogpGrad: the run ends with the exception
    calls.update(n=0, fail_from=2, fail_to=10 ** 9)
    with wb.Session(wb.models.batch_callback(D, density), C, seed=2) as s:
        with pytest.raises(ArithmeticError, match="blew up"):
            s.init(init_radius=1.0)


def test_failing_batched_density_is_reported_by_the_one_shot_call(wb):
    """The C-ABI one-shot call prints one line per failed batch evaluation through the
    PRINT_CALLBACK (handlers.hpp:30-36) and still returns draws."""
    import ctypes
    from walnuts_b200 import _ffi
    D, C = 2, 4
    state = {"n": 0}

    @wb.models.BATCH_LOGP_GRAD
    def density(Cn, Dn, ld, theta, grad, lp, stream, data):
        import torch
        state["n"] += 1
        if state["n"] == 40:
            return 7
        with torch.cuda.stream(torch.cuda.ExternalStream(int(stream or 0))):
            t = torch.as_tensor(wb.models._DevicePointer(theta, (Cn, ld)), device="cuda")
            g = torch.as_tensor(wb.models._DevicePointer(grad, (Cn, ld)), device="cuda")
            l = torch.as_tensor(wb.models._DevicePointer(lp, (Cn,)), device="cuda")
            g.copy_(-t)
            l.copy_(-0.5 * (t[:, :Dn] * t[:, :Dn]).sum(dim=1))
        return 0

    lines = []

    @_ffi.print_callback_type
    def printer(msg, n, bad):
        lines.append(ctypes.string_at(msg, n).decode())

    desc = _ffi.WalnutModelDesc(kind=4, D=D, N=0,
                                data0=ctypes.cast(density, ctypes.c_void_p).value, data1=None)
    out = np.zeros((C, 20, D))
    lengths = np.zeros(2 * C, np.int32)
    _ffi._ffi_sample_device(
        ctypes.byref(desc), D, None, C, 1, 1, 1.0, None, 20, 20, 20, 20, 5, 5, 1, 0.5, 0.1,
        1.0, 1.01, 4.0, 1e-5, 15.0, 1.0, 0.8, 0.05, 0.8, 0.9, 1e-4, 0.5, False, out, out.size,
        lengths, None, None, 0, printer)
    assert lengths[C:].tolist() == [20] * C
    assert np.all(np.isfinite(out))
    hits = [ln for ln in lines if "Error evaluating the log density" in ln]
    assert len(hits) == 1 and "logp failed with code 7" in hits[0]


# ---- DistributedController over REAL sessions (e) -----------------
WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["WB200_ROOT"])
import walnuts_b200 as wb
from walnuts_b200.distributed import DistributedController, SessionAdapter, shard

backend = os.environ["WB200_BACKEND"]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
dev = rank % ndev if backend == "gloo" else rank
torch.cuda.set_device(dev)
dist.init_process_group(backend, rank=rank, world_size=world,
                        **({"device_id": torch.device("cuda", dev)} if backend == "nccl" else {}))
D, TOTAL = 12, 48
var = np.linspace(0.5, 6.0, D)
off, cnt = shard(TOTAL, world, rank)
cfg = dict(min_warmup_iter=20, max_warmup_iter=300, min_sampling_iter=20,
           max_sampling_iter=300)
with wb.Session(wb.models.diag_gaussian(var), cnt, seed=21, chain_offset=off, device=dev,
                **cfg) as s:
    s.init(init_radius=2.0)
    s.reserve(300)
    ctl = DistributedController(SessionAdapter(s, torch.device("cuda", dev)),
                                torch.device("cuda", dev))
    warm = ctl.run_warmup(20, 300, 5, mass_tol=0.9, step_tol=0.35)
    n, rhat = ctl.run_sampling(20, 300, 5, rhat_tol=1.01)
    s.sync()
    draws = s.draws(0, n)
    out = dict(rank=rank, warm=warm, n=n, rhat=rhat, off=off, cnt=cnt,
               draws=draws.tolist())
with open(os.environ["WB200_OUT"] + f".{rank}", "w") as f:
    json.dump(out, f)
dist.barrier()
dist.destroy_process_group()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_ranks(world, backend, tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WB200_ROOT=str(ROOT), WB200_BACKEND=backend,
               WB200_OUT=str(tmp_path / f"out_{backend}_{world}"), MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(_free_port()), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return [json.loads(Path(env["WB200_OUT"] + f".{r}").read_text()) for r in range(world)]


@pytest.mark.timeout(900)
def test_distributed_controller_on_real_sessions_is_invariant_to_sharding(tmp_path):
    """DistributedController(SessionAdapter(Session)) with the chains sharded over 2 ranks
    (NCCL when the box has two GPUs; otherwise both ranks share GPU 0 and meet over gloo)
    stops warm-up and sampling at the same iterations, reports the same R-hat and produces
    the same draws as one rank holding all the chains."""
    import torch
    one = _run_ranks(1, "gloo", tmp_path)[0]
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    two = _run_ranks(2, backend, tmp_path)
    print(f"\nbackend {backend}: warm-up stopped at {one['warm']}, sampling at {one['n']}, "
          f"R-hat {one['rhat']:.6f}")
    assert 20 < one["warm"] < 300 and 20 <= one["n"] < 300, "case must stop early"
    whole = np.asarray(one["draws"])
    for r in two:
        assert r["warm"] == one["warm"] and r["n"] == one["n"]
        assert r["rhat"] == pytest.approx(one["rhat"], rel=1e-10)
        np.testing.assert_array_equal(np.asarray(r["draws"]),
                                      whole[r["off"]:r["off"] + r["cnt"]])
