"""The DEVICE transition state machine (walnuts_b200/csrc/chain_kernel.cuh) compiled
with g++ for one emulated thread per chain, against the oracle's Philox policy.

Same source as the CUDA build, same fp64 arithmetic, sums in element order — so
whole warm-up + sampling runs must equal the oracle's bit for bit: iterative doubling
vs the reference's recursion (walnuts.hpp:464-495), the halving and reversibility
ladders (:254-345), Barker / Metropolis selection (:368-387), Adam, the discounted
Welford estimators and the min-micro controller (adaptive_walnuts.hpp:234-251).
Both arithmetic policies are pinned: "exact" (-DWB200_EXACT_ARITH, every rounding
separate) against the oracle's reference policy -- the one that reproduces the
unmodified reference headers bit for bit -- and "fused" (what ships: fused multiply-add
at the accumulate sites) against the oracle's fused policy.  Runs without a GPU.
"""
import numpy as np
import pytest

from oracle.binding import Target, default_config
from tests import host_emu

CASES = [
    ("std_normal", 10, dict(), 0.4, 60, 60),
    ("std_normal", 100, dict(), 0.5, 30, 30),
    ("diag_gaussian", 12, dict(max_trajectory_doublings=8), 0.7, 80, 80),
    ("funnel", 11, dict(max_step_halvings=8, max_trajectory_doublings=7), 0.5, 120, 120),
    ("std_normal", 5, dict(min_micro_steps=2, max_macro_steps_target=3.0), 1.5, 80, 80),
    ("diag_gaussian", 7, dict(max_step_halvings=1, max_trajectory_doublings=3), 0.3, 40, 40),
    ("funnel", 2, dict(max_step_halvings=10, max_trajectory_doublings=10), 1.0, 60, 60),
    ("std_normal", 1, dict(max_trajectory_doublings=12), 0.2, 30, 30),
    ("std_normal", 37, dict(max_macro_steps_target=1.0), 0.05, 25, 25),  # min_micro grows
]


@pytest.fixture(scope="module", params=["fused", "exact"])
def emu(request, oracle):
    exact = request.param == "exact"
    lib = host_emu.build(exact_arith=exact)
    with oracle.fused_arith(not exact):
        yield lib


@pytest.mark.parametrize("engine", ["chain", "tick"])
@pytest.mark.parametrize("kind,D,over,step0,nw,ns", CASES)
def test_device_state_machine_equals_oracle_bitwise(emu, oracle, engine, kind, D, over, step0,
                                                    nw, ns):
    rng = np.random.default_rng(100 * D + nw)
    prec = rng.uniform(0.05, 20.0, D) if kind == "diag_gaussian" else None
    target = Target(kind, D, prec=prec)
    cfg = default_config(**over)
    for chain in (0, 5):
        th0 = rng.normal(size=D)
        m0 = rng.uniform(0.3, 3.0, D)
        e = host_emu.run_chain(emu, kind, D, prec, cfg, 4242, chain, th0, m0, step0, nw, ns,
                               engine=engine)
        o = oracle.run_chain(target, cfg, 4242, chain, th0, m0, step0, nw, ns, rng_policy=1)
        np.testing.assert_array_equal(e["draws"], np.concatenate([o["warmup_draws"], o["draws"]]))
        np.testing.assert_array_equal(e["lp"], np.concatenate([o["warmup_lp"], o["lp"]]))
        np.testing.assert_array_equal(e["depth"], np.concatenate([o["warmup_depth"], o["depth"]]))
        np.testing.assert_array_equal(e["step_trace"][:nw], o["warmup_step"])
        np.testing.assert_array_equal(e["warmup_inv_mass"], o["warmup_inv_mass"])
        np.testing.assert_array_equal(e["inv_mass"], o["inv_mass"])
        assert e["step"] == o["step"]
        assert e["min_micro"] == o["min_micro"]
        assert e["grad_evals"] == o["grad_evals"]


# Extreme starting points and step sizes: energies of 1e200, rejected extensions at every
# doubling, halving ladders that run out.  The device logic must make the same decisions
# as the reference's (the chain stays put, same gradient count, same adaptation state).
EDGE_CASES = [
    ("std_normal", 6, 1e100, 0.5),
    ("diag_gaussian", 4, 1e140, 1.0),
    ("funnel", 5, 30.0, 2.0),
    ("funnel", 5, 300.0, 2.0),
    ("std_normal", 6, 50.0, 1e6),
]


@pytest.mark.parametrize("budget", [1, 37, 400])
@pytest.mark.parametrize("kind,D,over,step0,nw,ns", CASES[:5])
def test_chain_engine_free_running_launches_equal_oracle_bitwise(emu, oracle, budget, kind, D,
                                                                 over, step0, nw, ns):
    """ChainParams::eval_budget: a chain advanced by launches of `budget` gradient
    evaluations each -- the transition that exhausts a budget is finished, its excess is
    taken off the next budget -- is the very chain of the fixed-length run (the reference's
    threads, adapt.hpp:110-129 / sampler.hpp:79-94, stop between iterations, too), and the
    long-run cost of a launch is the budget."""
    rng = np.random.default_rng(100 * D + nw)
    prec = rng.uniform(0.05, 20.0, D) if kind == "diag_gaussian" else None
    target = Target(kind, D, prec=prec)
    cfg = default_config(**over)
    th0 = rng.normal(size=D)
    m0 = rng.uniform(0.3, 3.0, D)
    e = host_emu.run_chain(emu, kind, D, prec, cfg, 4242, 3, th0, m0, step0, nw, ns,
                           eval_budget=budget)
    o = oracle.run_chain(target, cfg, 4242, 3, th0, m0, step0, nw, ns, rng_policy=1)
    np.testing.assert_array_equal(e["draws"], np.concatenate([o["warmup_draws"], o["draws"]]))
    np.testing.assert_array_equal(e["lp"], np.concatenate([o["warmup_lp"], o["lp"]]))
    np.testing.assert_array_equal(e["inv_mass"], o["inv_mass"])
    assert e["step"] == o["step"] and e["grad_evals"] == o["grad_evals"]
    # debts carry over: launches * budget brackets the work (each phase ends inside a launch)
    assert (e["launches"] - 2) * budget <= e["grad_evals"]
    if budget == 1:  # at least one launch per iteration, idle launches pay the debts off
        assert e["launches"] >= nw + ns


@pytest.mark.parametrize("engine", ["chain", "tick"])
@pytest.mark.parametrize("kind,D,scale,step0", EDGE_CASES)
def test_extreme_inputs_take_the_reference_decisions(emu, oracle, engine, kind, D, scale, step0):
    rng = np.random.default_rng(1)
    prec = rng.uniform(0.5, 2.0, D) if kind == "diag_gaussian" else None
    target = Target(kind, D, prec=prec)
    cfg = default_config(max_trajectory_doublings=6, max_step_halvings=4)
    th0, m0 = rng.normal(size=D) * scale, np.ones(D)
    e = host_emu.run_chain(emu, kind, D, prec, cfg, 5, 0, th0, m0, step0, 8, 8, engine=engine)
    o = oracle.run_chain(target, cfg, 5, 0, th0, m0, step0, 8, 8, rng_policy=1)
    np.testing.assert_array_equal(e["draws"], np.concatenate([o["warmup_draws"], o["draws"]]))
    np.testing.assert_array_equal(e["lp"], np.concatenate([o["warmup_lp"], o["lp"]]))
    np.testing.assert_array_equal(e["depth"], np.concatenate([o["warmup_depth"], o["depth"]]))
    np.testing.assert_array_equal(e["inv_mass"], o["inv_mass"])
    assert e["step"] == o["step"] and e["grad_evals"] == o["grad_evals"]
    assert np.all(np.isfinite(e["draws"]))


FREE_CASES = [
    ("std_normal", 10, dict(), 0.4),
    ("diag_gaussian", 12, dict(max_trajectory_doublings=8), 0.7),
    ("funnel", 11, dict(max_step_halvings=8, max_trajectory_doublings=7), 0.5),
]


@pytest.mark.parametrize("kind,D,over,step0", FREE_CASES)
def test_free_running_sampling_equals_oracle_bitwise(emu, oracle, kind, D, over, step0):
    """free-running lock-step mode (wb200_session_sample_ticks): after a tick budget the
    chain has completed some k transitions, and those are the oracle's first k draws"""
    rng = np.random.default_rng(7 * D)
    prec = rng.uniform(0.05, 20.0, D) if kind == "diag_gaussian" else None
    target, cfg = Target(kind, D, prec=prec), default_config(**over)
    th0, m0 = rng.normal(size=D), rng.uniform(0.3, 3.0, D)
    nw, cap = 40, 400
    e = host_emu.run_chain(emu, kind, D, prec, cfg, 99, 3, th0, m0, step0, nw, cap,
                           engine="tick", samp_ticks=1500)
    k = e["rows"][1] - nw
    assert 10 < k < cap
    o = oracle.run_chain(target, cfg, 99, 3, th0, m0, step0, nw, k, rng_policy=1)
    np.testing.assert_array_equal(e["draws"][:nw + k],
                                  np.concatenate([o["warmup_draws"], o["draws"]]))
    np.testing.assert_array_equal(e["lp"][nw:nw + k], o["lp"])
    np.testing.assert_array_equal(e["depth"][nw:nw + k], o["depth"])


@pytest.mark.parametrize("kind,D,over,step0", FREE_CASES)
def test_free_running_warmup_equals_oracle_bitwise(emu, oracle, kind, D, over, step0):
    """free-running adaptive warm-up (wb200_session_warmup_ticks): the k warm-up
    transitions completed within the tick budget are the oracle's first k"""
    rng = np.random.default_rng(11 * D)
    prec = rng.uniform(0.05, 20.0, D) if kind == "diag_gaussian" else None
    target, cfg = Target(kind, D, prec=prec), default_config(**over)
    th0, m0 = rng.normal(size=D), rng.uniform(0.3, 3.0, D)
    cap = 400
    e = host_emu.run_chain(emu, kind, D, prec, cfg, 17, 1, th0, m0, step0, cap, 0,
                           engine="tick", warm_ticks=1200, samp_ticks=0)
    k = e["rows"][0]
    assert 10 < k < cap
    o = oracle.run_chain(target, cfg, 17, 1, th0, m0, step0, k, 0, rng_policy=1)
    np.testing.assert_array_equal(e["draws"][:k], o["warmup_draws"])
    np.testing.assert_array_equal(e["lp"][:k], o["warmup_lp"])
    np.testing.assert_array_equal(e["step_trace"][:k], o["warmup_step"])
    np.testing.assert_array_equal(e["warmup_inv_mass"][:k], o["warmup_inv_mass"])


def test_trajectories_exercise_every_branch(emu, oracle):
    """the cases above must actually reach halvings, reversibility ladders, deep
    trees and rejected extensions, or the bitwise agreement proves little"""
    cfg = default_config(max_step_halvings=8, max_trajectory_doublings=7)
    rng = np.random.default_rng(5)
    th0, m0 = rng.normal(size=11), np.ones(11)
    e = host_emu.run_chain(emu, "funnel", 11, None, cfg, 1, 0, th0, m0, 0.5, 200, 200)
    depths = np.bincount(e["depth"], minlength=9)
    assert (depths > 0).sum() >= 5          # many different tree depths
    assert e["grad_evals"] > 3 * 400 * 2    # ladders well beyond one step per leaf
