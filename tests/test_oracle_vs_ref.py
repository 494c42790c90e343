"""The oracle restatement against the UNMODIFIED reference headers.

* golden: tests/golden/ref_chains.npz was produced by tests/golden/make_golden.py
  from oracle/_ref (reference headers + Eigen shim); the oracle must reproduce
  every array bit for bit (same std::mt19937_64 stream, same libstdc++).
* live: where oracle/_ref is present, extra randomly drawn cases are diffed
  directly, including the multi-chain controllers.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle.binding import Target, default_config
from tests.golden.make_golden import CASES

GOLD = np.load(Path(__file__).parent / "golden" / "ref_chains.npz")


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_chain_matches_reference_golden(oracle, case):
    name, kind, D, extra, over, seed, chain, step0, nw, ns = case
    t = Target(kind, D, **extra)
    cfg = default_config(**over)
    th0, m0 = GOLD[f"{name}/theta0"], GOLD[f"{name}/mass0"]
    r = oracle.run_chain(t, cfg, seed, chain, th0, m0, step0, nw, ns)
    for k in ("warmup_draws", "warmup_lp", "warmup_step", "warmup_inv_mass", "draws",
              "lp", "inv_mass"):
        np.testing.assert_array_equal(r[k], GOLD[f"{name}/{k}"], err_msg=k)
    step, mm, evals = GOLD[f"{name}/scalars"]
    assert r["step"] == step and r["min_micro"] == mm and r["grad_evals"] == evals
    s = oracle.run_sampler(t, seed, chain, th0, 1 / m0, step0, 6, 6, 2, 0.5, 200)
    np.testing.assert_array_equal(s["draws"], GOLD[f"{name}/fixed_draws"])
    np.testing.assert_array_equal(s["lp"], GOLD[f"{name}/fixed_lp"])
    assert s["grad_evals"] == GOLD[f"{name}/fixed_evals"][0]


def test_initialisation_matches_reference_golden(oracle):
    pos = oracle.init_positions(4, 6, 42, 2.0)
    np.testing.assert_array_equal(pos, GOLD["init/positions"])
    t = Target("diag_gaussian", 6, prec=np.array([1, 2, 3, 4, 5, 6.0]))
    mass, steps = oracle.init_mass_step(t, pos, 42, 1.0)
    np.testing.assert_array_equal(mass, GOLD["init/mass"])
    np.testing.assert_array_equal(steps, GOLD["init/steps"])
    _, steps2 = oracle.init_mass_step(t, pos, 42, 100.2, mass_in=np.ones((4, 6)))
    np.testing.assert_array_equal(steps2, GOLD["init/steps_given_mass"])


@pytest.mark.parametrize("kind,D", [("std_normal", 4), ("diag_gaussian", 9), ("funnel", 6)])
def test_chain_matches_reference_live(oracle, ref, kind, D):
    rng = np.random.default_rng(D)
    extra = dict(prec=rng.uniform(0.1, 10, D)) if kind == "diag_gaussian" else {}
    t = Target(kind, D, **extra)
    cfg = default_config(max_step_halvings=7, max_trajectory_doublings=6,
                         min_micro_steps=1 + D % 2)
    th0, m0 = rng.normal(size=D), rng.uniform(0.2, 3, D)
    a = oracle.run_chain(t, cfg, 4321, 2, th0, m0, 0.4, 80, 80)
    b = ref.run_chain(t, cfg, 4321, 2, th0, m0, 0.4, 80, 80)
    for k in ("warmup_draws", "warmup_step", "warmup_inv_mass", "draws", "lp", "inv_mass"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert a["grad_evals"] == b["grad_evals"] and a["step"] == b["step"]
    assert ref.leapfrog_error(t, th0, m0, 1 / m0, 0.3) == oracle.leapfrog_error(
        t, th0, m0, 1 / m0, 0.3)


def test_multichain_fixed_length_run_matches_reference(oracle, ref):
    """api.hpp:33-69 with min == max iterations (deterministic, docs/py.rst:13-20):
    the reference's threaded driver and the oracle's give identical draws."""
    D, C = 5, 3
    t = Target("diag_gaussian", D, prec=np.array([1.0, 4.0, 0.25, 9.0, 1.0]))
    cfg = default_config(min_warmup_iter=60, max_warmup_iter=60, min_sampling_iter=40,
                         max_sampling_iter=40)
    pos = oracle.init_positions(C, D, 11, 2.0)
    mass, steps = oracle.init_mass_step(t, pos, 11, 1.0)
    a = oracle.walnuts(t, cfg, 48, pos, mass, steps, save_warmup=True)
    b = ref.walnuts(t, cfg, 48, pos, mass, steps, save_warmup=True)
    np.testing.assert_array_equal(a["out"], b["out"])
    np.testing.assert_array_equal(a["stepsize"], b["stepsize"])
    np.testing.assert_array_equal(a["inv_metric"], b["inv_metric"])
    assert a["grad_evals"] == b["grad_evals"]
    assert list(a["warmup_lengths"]) == [60] * C and list(a["sampling_lengths"]) == [40] * C


def test_adaptive_stopping_lengths_within_bounds(oracle, ref):
    """python/tests/test_pyfunc.py:38-64 for both CPU drivers."""
    t = Target("std_normal", 2)
    cfg = default_config(min_warmup_iter=10, max_warmup_iter=30, min_sampling_iter=10,
                         max_sampling_iter=30)
    pos = oracle.init_positions(4, 2, 3, 2.0)
    mass, steps = oracle.init_mass_step(t, pos, 3, 1.0)
    for impl in (oracle, ref):
        r = impl.walnuts(t, cfg, 5, pos, mass, steps, save_warmup=True)
        assert all(10 <= n <= 30 for n in r["warmup_lengths"])
        assert all(10 <= n <= 30 for n in r["sampling_lengths"])


def test_invalid_iteration_bounds_message(oracle):
    """python/tests/test_pyfunc.py:67-71"""
    t = Target("std_normal", 2)
    cfg = default_config(min_sampling_iter=100, max_sampling_iter=99)
    with pytest.raises(ValueError, match="min_iter must be"):
        oracle.run_chain(t, cfg, 1, 0, np.zeros(2), np.ones(2), 0.5, 1, 1)
