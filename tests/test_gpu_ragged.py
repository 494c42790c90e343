"""Free-running launches of the chain-resident engine and per-chain (ragged) early
stopping in the one-shot call.

Reference: every chain is a thread that iterates at its own pace until the controller
stops it (AdaptWorker adapt.hpp:110-129, ChainWorker sampler.hpp:79-94), so the final
lengths differ from chain to chain (docs/py.rst:13-20), the controllers wait until every
chain has done min_iter (adapt.hpp:196-203, sampler.hpp:134-141) and stop on convergence
or when every chain is at max_iter (adapt.hpp:219-221, sampler.hpp:148-150).  On the
device "equal time" is equal work: a budget of gradient evaluations per chain and launch
(chain_kernel.cuh, ChainParams::eval_budget).  A chain's draws are the first draws of the
very chain a fixed-length run produces (oracle parity of those: test_gpu_parity.py).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def funnel_session(wb, D, C, seed, **tuning):
    rng = np.random.default_rng(3)
    pos = rng.normal(size=(C, D))
    pos[:, 0] = rng.normal(scale=2.0, size=C)          # spread over the funnel: orbit
    pos[:, 1:] *= np.exp(0.5 * pos[:, :1])             # lengths differ by chain
    s = wb.Session(wb.models.funnel(D), C, seed=seed, max_step_halvings=8,
                   max_trajectory_doublings=8, **tuning)
    s.init(positions=pos, mass=np.ones((C, D)), steps=np.full(C, 0.4))
    return s


@pytest.mark.parametrize("D,C", [(11, 96), (100, 300)])
def test_free_running_rows_are_prefixes_of_the_fixed_length_chains(wb, D, C):
    cap = 400
    with funnel_session(wb, D, C, 77) as s:
        s.reserve(cap)
        s.warmup(0).freeze()
        budgets = [150, 400, 90, 1000, 250]
        for b in budgets:
            s.sample_ticks(b)
        s.sync()
        rows = s.chain_rows()
        evals = s.state()["grad_evals"]
        free = s.draws(0, int(rows.max()))
        mn, mx, total, total_evals = s.iter_stats(sampling=True)
    assert (mn, mx, total) == (rows.min(), rows.max(), rows.sum())
    assert total_evals == evals.sum()
    assert rows.min() >= 1 and rows.max() < cap
    assert rows.max() >= 2 * rows.min(), "chains in the neck and in the mouth differ in cost"
    # equal work: every chain has spent the budgets, plus at most its last transition
    assert evals.min() >= sum(budgets)
    with funnel_session(wb, D, C, 77) as s:
        s.reserve(int(rows.max()))
        s.warmup(0).freeze().sample(int(rows.max())).sync()
        fixed = s.draws(0, int(rows.max()))
    for c in range(C):
        np.testing.assert_array_equal(free[c, :rows[c]], fixed[c, :rows[c]])


def test_free_running_iteration_cap_and_full_rows(wb):
    """iter_cap stops a chain at the phase's max_iter (adapt.hpp:116, sampler.hpp:82); a
    chain whose rows are full idles instead of overwriting."""
    D, C = 11, 64
    with funnel_session(wb, D, C, 5) as s:
        s.reserve(30)
        s.warmup(0).freeze()
        for _ in range(40):
            s.run_evals(300, sampling=True, iter_cap=12)
        s.sync()
        assert set(s.chain_rows()) == {12}
        assert s.iter_stats(sampling=True)[:3] == (12, 12, 12 * C)
        for _ in range(60):
            s.sample_ticks(500)
        s.sync()
        assert set(s.chain_rows()) == {30}
        with pytest.raises(RuntimeError, match="free-running"):
            s.sample(1)


def test_free_running_warmup_adapts_every_chain_on_its_own_count(wb):
    D, C = 11, 128
    with funnel_session(wb, D, C, 9) as s:
        s.reserve(4000)
        for _ in range(10):
            s.warmup_ticks(400, store=True)
        s.sync()
        rows = s.chain_rows()
        mn, mx, total, _ = s.iter_stats(sampling=False)
        assert (mn, mx, total) == (rows.min(), rows.max(), rows.sum())
        assert rows.max() > rows.min() and rows.max() < 3000
        s.freeze()
        st = s.state()
        assert np.all(np.isfinite(st["step"])) and np.all(st["step"] > 0)
        assert np.all(np.isfinite(st["inv_mass"])) and np.all(st["inv_mass"] > 0)
        # sampling rows follow each chain's own warm-up rows
        s.sample_ticks(2000).sync()
        rows2 = s.chain_rows()
        assert np.all(rows2 > rows)


def test_streamed_summaries_of_free_running_chain_engine(wb, oracle):
    """the streaming accumulators fold each chain's own number of staged rows: the
    summaries equal those of the kept ragged draws (summary.hpp:594-769 on ragged chains)"""
    D, C = 11, 48
    with funnel_session(wb, D, C, 21) as s:
        s.reserve(2000)
        s.warmup(0).freeze()
        for _ in range(12):
            s.sample_ticks(1500)
        s.sync()
        rows = s.chain_rows()
        kept = s.draws(0, int(rows.max()))
    assert rows.min() >= 3
    with funnel_session(wb, D, C, 21) as s:
        s.reserve(64)                      # a staging block, refilled by every launch
        s.warmup(0).freeze()
        s.stream_begin(max_lags=32)
        for _ in range(12):
            s.sample_ticks(1500)
        got = s.stream_summary()
        counts = s.stream_counts()
    # a staging block of 64 rows can cut a launch short: same chains, possibly fewer rows
    assert np.all(counts <= rows) and np.all(counts >= 3)
    chains = [kept[c, :counts[c]] for c in range(C)]
    np.testing.assert_allclose(got["r_hat"], oracle.r_hat(chains), rtol=1e-9)
    np.testing.assert_allclose(got["mean"], np.concatenate(chains).mean(0), rtol=1e-9,
                               atol=1e-12)
    ok = got["truncated"] == 0
    assert ok.any()
    np.testing.assert_allclose(got["ess"][ok], oracle.ess(chains)[ok], rtol=1e-7)


def _replay_budget(evals_seen, iters_seen, stride=5):
    per_iter = evals_seen / iters_seen if iters_seen > 0 else 16.0
    return max(1, int(math.floor(per_iter * stride + 0.5)))


def test_one_shot_call_stops_every_chain_on_its_own(wb, oracle, monkeypatch):
    """walnutpie_sample_device with min < max on the chain-resident engine: final lengths
    differ by chain, lie in [min, max], and equal a replay of the reference's controller
    rules on a Session driven by the same budgets; every chain's draws are the first draws
    of the fixed-length run's chain.  WB200_BLOCKS=uniform: equal lengths."""
    D, C = 16, 32
    model = wb.models.diag_gaussian(10.0 ** (4 * np.arange(D) / (D - 1)))
    fixed_warm = dict(min_warmup_iter=100, max_warmup_iter=100)
    kw = dict(min_sampling_iter=20, max_sampling_iter=300, rhat_converge_tol=1.01)
    seed, ident = 17, 1
    fit = wb.walnuts_device(model, num_chains=C, seed=seed, id=ident, **fixed_warm, **kw)
    lens = np.array([len(f) for f in fit])
    assert lens.min() >= kw["min_sampling_iter"] and lens.max() <= kw["max_sampling_iter"]
    assert lens.max() > lens.min(), "orbit lengths differ, so must the final lengths"
    assert lens.max() < kw["max_sampling_iter"], "case must stop on R-hat"
    # the same chains at fixed length
    full = wb.walnuts_device(model, num_chains=C, seed=seed, id=ident, **fixed_warm,
                             **{**kw, "min_sampling_iter": int(lens.max()),
                                "max_sampling_iter": int(lens.max())})
    for c in range(C):
        np.testing.assert_array_equal(np.asarray(fit[c]), np.asarray(full[c])[:lens[c]])
    # replay of the controller on a session: budgets of 5 average iterations
    with wb.Session(model, C, seed=seed + ident + C, **fixed_warm, **kw) as s:
        s.init(init_radius=2.0)
        s.reserve(kw["max_sampling_iter"], trace=True)
        s.warmup(100).freeze()
        st = s.iter_stats(sampling=True)
        evals0, evals_seen, iters_seen = st[3], 0, 0
        # the library has seen the warm-up's cost only if the warm-up ran free
        while st[2] < C * kw["max_sampling_iter"]:
            s.run_evals(_replay_budget(evals_seen, iters_seen), sampling=True,
                        iter_cap=kw["max_sampling_iter"])
            st = s.iter_stats(sampling=True)
            evals_seen += st[3] - evals0
            evals0 = st[3]
            iters_seen = st[2]
            if st[0] >= kw["min_sampling_iter"] and st[2] < C * kw["max_sampling_iter"]:
                m0 = s.lp_moments()
                m = s.lp_moments(center=m0[0] / m0[3])
                var_of_means = (m[1] - m[0] * m[0] / m[3]) / (m[3] - 1.0)
                if math.sqrt(1 + var_of_means / (m[2] / m[3])) <= kw["rhat_converge_tol"]:
                    break
        np.testing.assert_array_equal(s.chain_rows(), lens)
    monkeypatch.setenv("WB200_BLOCKS", "uniform")
    uni = wb.walnuts_device(model, num_chains=C, seed=seed, id=ident, **fixed_warm, **kw)
    assert len({len(f) for f in uni}) == 1


def test_one_shot_call_free_running_warmup_and_saved_warmup_layout(wb):
    """min_warmup < max_warmup: the warm-up runs free, too; with save_warmup every chain's
    block holds its own warm-up draws followed directly by its sampling draws
    (walnutpy.cpp:196-203, handlers.hpp:73-89)."""
    D, C = 16, 32
    model = wb.models.diag_gaussian(10.0 ** (4 * np.arange(D) / (D - 1)))
    kw = dict(min_warmup_iter=20, max_warmup_iter=400, min_sampling_iter=25,
              max_sampling_iter=25, mass_converge_tol=0.9, step_size_converge_tol=0.35)
    fit = wb.walnuts_device(model, num_chains=C, seed=8, save_warmup=True,
                            save_inv_metric=True, **kw)
    wl = np.array([len(f.warmup.warmup_draws) for f in fit])
    sl = np.array([len(f) for f in fit])
    assert wl.min() >= kw["min_warmup_iter"] and wl.max() <= kw["max_warmup_iter"]
    assert wl.max() > wl.min()
    assert set(sl) == {25}, "min == max: every chain reaches the last sampling iteration"
    moved = 0
    for f in fit:
        assert np.all(np.isfinite(np.asarray(f))) and np.all(np.isfinite(f.warmup.warmup_draws))
        assert f.warmup.stepsize > 0
        # no stale or skipped rows at the seam: consecutive states of a chain differ (a
        # transition may stay where it is, so not every one of them)
        seam = np.concatenate([f.warmup.warmup_draws[-2:], np.asarray(f)[:2]])
        moved += int(np.sum(np.abs(np.diff(seam, axis=0)).sum(1) > 0))
        assert np.abs(np.asarray(f)).sum(1).min() > 0 and \
            np.abs(f.warmup.warmup_draws).sum(1).min() > 0, "an unwritten row"
    assert moved >= 0.9 * 3 * C


def test_one_shot_summary_call_reports_ragged_lengths(wb):
    D, C = 16, 32
    out = wb.walnuts_device_summary(
        wb.models.diag_gaussian(10.0 ** (4 * np.arange(D) / (D - 1))), num_chains=C, seed=17,
        min_warmup_iter=100, max_warmup_iter=100, min_sampling_iter=20,
        max_sampling_iter=300, rhat_converge_tol=1.01)
    lens = out["sampling_lengths"]
    assert lens.shape == (C,) and lens.min() >= 20 and lens.max() < 300
    assert lens.max() > lens.min()
    assert out["sampling_iters"] == lens.max()
    assert np.all(np.isfinite(out["mean"])) and np.all(out["ess"] > 0)


def test_one_shot_call_runs_free_on_the_lock_step_engine(wb, monkeypatch):
    """A logistic model (tensor-core gradient, lock-step ticks) with min < max: a block is
    a budget of ticks, every chain stops on its own, and a chain that has reached max_iter
    stays stopped; the posterior agrees with the run in blocks of equal iteration counts."""
    from tests.test_gpu_parity import make_logistic
    N, D, C = 400, 8, 96
    X, y = make_logistic(N, D, 3)
    model = wb.models.logistic(X, y)
    kw = dict(num_chains=C, seed=6, min_warmup_iter=40, max_warmup_iter=120,
              min_sampling_iter=30, max_sampling_iter=150, rhat_converge_tol=1.02,
              mass_converge_tol=0.9, step_size_converge_tol=0.35, max_trajectory_doublings=8)
    fit = wb.walnuts_device(model, save_warmup=True, **kw)
    wl = np.array([len(f.warmup.warmup_draws) for f in fit])
    sl = np.array([len(f) for f in fit])
    assert wl.min() >= 40 and wl.max() <= 120 and sl.min() >= 30 and sl.max() <= 150
    assert sl.max() > sl.min() or sl.max() == 150
    assert wl.max() > wl.min() or wl.max() == 120
    free = np.concatenate([np.asarray(f) for f in fit])
    assert np.all(np.isfinite(free))
    out = wb.walnuts_device_summary(model, **kw)        # the same run, draws streamed
    np.testing.assert_array_equal(out["sampling_lengths"], sl)
    np.testing.assert_allclose(out["mean"], free.mean(0), rtol=1e-9, atol=1e-12)
    monkeypatch.setenv("WB200_BLOCKS", "uniform")
    uni = wb.walnuts_device(model, **kw)
    assert len({len(f) for f in uni}) == 1
    pooled = np.concatenate([np.asarray(f) for f in uni])
    se = np.sqrt(free.var(0) / 200 + pooled.var(0) / 200)    # generous: ESS >= 200 each
    assert np.max(np.abs(free.mean(0) - pooled.mean(0)) / se) < 5.0
    # all chains reach max_iter and stop there
    fit = wb.walnuts_device(model, num_chains=C, seed=6, min_warmup_iter=20,
                            max_warmup_iter=20, min_sampling_iter=10, max_sampling_iter=40,
                            rhat_converge_tol=1.0 + 1e-9)
    assert {len(f) for f in fit} == {40}


@pytest.mark.parametrize("C,D", [(1, 1), (2, 3), (5, 129)])
def test_free_running_one_shot_edge_sizes(wb, C, D):
    """one chain (no R-hat: runs to max_iter, as the reference's NaN comparison does), one
    dimension, a dimension just above a shape boundary"""
    fit = wb.walnuts_device(wb.models.std_normal(D), num_chains=C, seed=2, min_warmup_iter=10,
                            max_warmup_iter=40, min_sampling_iter=10, max_sampling_iter=45,
                            save_warmup=True)
    for f in fit:
        assert 10 <= len(f) <= 45 and 10 <= len(f.warmup.warmup_draws) <= 40
        assert np.all(np.isfinite(np.asarray(f))) and np.asarray(f).shape[1] == D
    if C == 1:
        assert len(fit[0]) == 45
