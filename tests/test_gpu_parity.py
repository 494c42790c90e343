"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.

Bars (BASELINE.json north_star):
 * fixed-step leapfrog orbits: <= 1e-12 relative, fp64;
 * identically seeded (Philox) trajectories agree until the first
   rounding-induced branch flip -- asserted as: every chain agrees over a
   stated prefix, and the large majority over the whole run;
 * posterior moments within Monte Carlo standard error.
Element-wise arithmetic on the device mirrors the oracle operation by operation;
only cross-element sums differ (summation order), so tolerances are tight.
"""
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

ORBIT_RTOL = 1e-12


def rel_err(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


# ---- generator ---------------------------------------------------------------
def test_device_philox_known_answers(wb):
    from walnuts_b200 import _ffi
    q = np.array([[0, 0, 0, 0, 0, 0], [0xFFFFFFFF] * 6,
                  [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0]],
                 dtype=np.uint32)
    out = np.zeros((3, 4), np.uint32)
    _ffi.philox(q, 3, out)
    assert out.tolist() == [[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8],
                            [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD],
                            [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]]


def test_device_normals_match_oracle_stream(wb, oracle):
    from walnuts_b200 import _ffi
    for n in (1, 2, 7, 1000):
        dev = np.zeros(n)
        _ffi.philox_normals(123, 45, 6, 0, n, dev)
        ora = oracle.philox_normals(123, 45, 6, 0, n)
        # same Philox words; log / sincos differ between libm and CUDA by <= a few ulp
        np.testing.assert_allclose(dev, ora, rtol=1e-13, atol=1e-15)


# ---- fixed-step orbits -------------------------------------------------------
ORBIT_CASES = [
    ("std_normal", 100, 0.37, 64),
    ("diag_gaussian", 1000, 0.2, 50),
    ("funnel", 100, 0.05, 40),
    ("std_normal", 1, 0.5, 10),
    ("diag_gaussian", 513, 0.1, 33),   # odd dimension: padded lane
    ("funnel", 2, 0.1, 20),
    ("diag_gaussian", 2000, 0.1, 16),
    ("std_normal", 4096, 0.1, 8),      # maximum supported dimension
]


def make_model(wb, kind, D, rng):
    if kind == "diag_gaussian":
        var = 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
        return wb.models.diag_gaussian(var), Target(kind, D, prec=1 / var)
    return getattr(wb.models, kind)(D), Target(kind, D)


@pytest.mark.parametrize("kind,D,step,nsteps", ORBIT_CASES)
def test_fixed_step_orbit_matches_oracle(wb, oracle, kind, D, step, nsteps):
    rng = np.random.default_rng(D + nsteps)
    model, target = make_model(wb, kind, D, rng)
    C = 5
    inv_mass = rng.uniform(0.5, 2.0, (C, D))
    if kind == "diag_gaussian":
        inv_mass *= 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
    theta = rng.normal(size=(C, D)) * (0.3 if kind == "funnel" else 1.0)
    rho = rng.normal(size=(C, D)) / np.sqrt(inv_mass)
    for sign in (+1, -1):
        th, rh, g, lp, jt = wb.orbit(model, theta, rho, inv_mass, sign * step, nsteps)
        for c in range(C):
            o = oracle.orbit(target, theta[c], rho[c], inv_mass[c], sign * step, nsteps)
            scale = max(np.max(np.abs(o[0])), 1e-30)
            assert np.max(np.abs(th[c] - o[0])) / scale <= ORBIT_RTOL
            assert np.max(np.abs(rh[c] - o[1])) / max(np.max(np.abs(o[1])), 1e-30) <= ORBIT_RTOL
            assert np.max(np.abs(g[c] - o[2])) / max(np.max(np.abs(o[2])), 1e-30) <= ORBIT_RTOL
            assert abs(lp[c] - o[3]) <= ORBIT_RTOL * max(abs(o[3]), 1.0)
            assert abs(jt[c] - o[4]) <= ORBIT_RTOL * max(abs(o[4]), 1.0)
        if kind != "funnel":
            # element-wise dynamics: not merely close, identical
            o = oracle.orbit(target, theta[0], rho[0], inv_mass[0], sign * step, nsteps)
            np.testing.assert_array_equal(th[0], o[0])
            np.testing.assert_array_equal(rh[0], o[1])


def test_orbit_reversibility_and_zero_steps(wb):
    """size-independent properties at the full c2 size: zero steps is the identity,
    forward then backward returns to the start (leapfrog is time-reversible)."""
    D, C = 1000, 64
    rng = np.random.default_rng(3)
    model = wb.models.ill_conditioned_gaussian(D)
    var = 1e4 ** (np.arange(D) / (D - 1))
    theta = rng.normal(size=(C, D)) * np.sqrt(var)
    inv_mass = np.tile(var, (C, 1))
    rho = rng.normal(size=(C, D)) / np.sqrt(inv_mass)
    th0, rh0, *_ = wb.orbit(model, theta, rho, inv_mass, 0.3, 0)
    np.testing.assert_array_equal(th0, theta)
    np.testing.assert_array_equal(rh0, rho)
    th1, rh1, *_ = wb.orbit(model, theta, rho, inv_mass, 0.3, 25)
    th2, rh2, *_ = wb.orbit(model, th1, rh1, inv_mass, -0.3, 25)
    assert np.max(np.abs(th2 - theta) / np.sqrt(var)) < 1e-11
    assert np.max(np.abs(rh2 - rho) * np.sqrt(var)) < 1e-11


# ---- identically seeded trajectories ----------------------------------------
TRAJ_CASES = [
    # kind, D, chains, tuning overrides, n_warmup, n_sampling
    ("std_normal", 100, 6, dict(), 60, 60),
    ("diag_gaussian", 10, 8, dict(max_trajectory_doublings=8), 80, 80),
    ("diag_gaussian", 1000, 3, dict(max_trajectory_doublings=10), 40, 30),
    ("funnel", 11, 8, dict(max_step_halvings=8, max_trajectory_doublings=7), 60, 60),
    ("std_normal", 5, 8, dict(min_micro_steps=2, max_macro_steps_target=3.0), 60, 60),
    ("diag_gaussian", 300, 4, dict(), 30, 30),          # 128-thread groups
    ("std_normal", 37, 5, dict(max_step_halvings=1), 40, 40),  # odd D, no halving
]


def first_divergence(a, b, rtol):
    """index of the first row where a and b differ by more than rtol (or len)."""
    scale = np.maximum(np.max(np.abs(b), axis=-1), 1e-300)
    bad = np.max(np.abs(a - b), axis=-1) / scale > rtol
    idx = np.flatnonzero(bad)
    return int(idx[0]) if idx.size else len(a)


@pytest.mark.parametrize("kind,D,C,over,nw,ns", TRAJ_CASES)
def test_seeded_trajectories_match_oracle(wb, oracle, kind, D, C, over, nw, ns):
    rng = np.random.default_rng(1000 + D)
    model, target = make_model(wb, kind, D, rng)
    seed = 777
    positions = rng.normal(size=(C, D))
    mass = rng.uniform(0.5, 2.0, (C, D))
    steps = rng.uniform(0.2, 0.6, C)
    cfg = default_config(**over)
    with wb.Session(model, C, seed=seed, **over) as s:
        s.init(positions=positions, mass=mass, steps=steps)
        s.reserve(nw + ns, trace=True)
        s.warmup(nw, store=True).freeze().sample(ns, store=True).sync()
        draws = s.draws(0, nw + ns)
        tr = s.trace(0, nw + ns)
        st = s.state()
    full, prefix = 0, []
    rtol = 1e-9
    for c in range(C):
        o = oracle.run_chain(target, cfg, seed, c, positions[c], mass[c], steps[c], nw, ns,
                             rng_policy=1)
        ref_draws = np.concatenate([o["warmup_draws"], o["draws"]])
        k = first_divergence(draws[c], ref_draws, rtol)
        prefix.append(k)
        if k == nw + ns:
            full += 1
            ref_depth = np.concatenate([o["warmup_depth"], o["depth"]])
            np.testing.assert_array_equal(tr["depth"][c], ref_depth)
            np.testing.assert_allclose(tr["lp"][c], np.concatenate([o["warmup_lp"], o["lp"]]),
                                       rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(tr["step"][c, :nw], o["warmup_step"], rtol=1e-10)
            np.testing.assert_allclose(tr["inv_mass"][c, :nw], o["warmup_inv_mass"], rtol=1e-9)
            np.testing.assert_allclose(st["inv_mass"][c], o["inv_mass"], rtol=1e-9)
            assert st["step"][c] == pytest.approx(o["step"], rel=1e-10)
            assert st["min_micro"][c] == o["min_micro"]
            assert int(st["grad_evals"][c]) == o["grad_evals"]
    print(f"\n[{kind} D={D}] first-divergence iteration per chain: {prefix} "
          f"(run length {nw + ns})")
    # every chain agrees over the first iterations; most agree to the end
    assert min(prefix) >= 5
    assert full >= (C + 1) // 2


@pytest.mark.parametrize("engine", ["chain", "tick"])
@pytest.mark.parametrize("kind,D,scale,step0", [
    ("std_normal", 6, 1e100, 0.5), ("diag_gaussian", 4, 1e140, 1.0),
    ("funnel", 5, 30.0, 2.0), ("funnel", 5, 300.0, 2.0), ("std_normal", 6, 50.0, 1e6)])
def test_extreme_inputs_take_the_reference_decisions(wb, oracle, monkeypatch, engine, kind, D,
                                                     scale, step0):
    """Energies of 1e200, rejected extensions at every doubling, exhausted halving ladders:
    both engines make the reference's decisions (the same cases run bit for bit against
    the oracle in tests/test_host_emulation.py)."""
    if engine == "tick":
        monkeypatch.setenv("WB200_ENGINE", "tick")
    rng = np.random.default_rng(1)
    prec = rng.uniform(0.5, 2.0, D) if kind == "diag_gaussian" else None
    model = {"std_normal": lambda: wb.models.std_normal(D),
             "diag_gaussian": lambda: wb.models.diag_gaussian(1.0 / prec),
             "funnel": lambda: wb.models.funnel(D)}[kind]()
    target = Target(kind, D, prec=prec)
    over = dict(max_trajectory_doublings=6, max_step_halvings=4)
    cfg = default_config(**over)
    th0, m0 = rng.normal(size=D) * scale, np.ones(D)
    C = 3
    with wb.Session(model, C, seed=5, **over) as s:
        s.init(positions=np.tile(th0, (C, 1)), mass=np.tile(m0, (C, 1)),
               steps=np.full(C, step0))
        s.reserve(16, trace=True)
        s.warmup(8, store=True).freeze().sample(8).sync()
        draws, tr, st = s.draws(0, 16), s.trace(0, 16), s.state()
    assert np.all(np.isfinite(draws))
    for c in range(C):
        o = oracle.run_chain(target, cfg, 5, c, th0, m0, step0, 8, 8, rng_policy=1)
        np.testing.assert_allclose(draws[c], np.concatenate([o["warmup_draws"], o["draws"]]),
                                   rtol=1e-9)
        np.testing.assert_array_equal(tr["depth"][c],
                                      np.concatenate([o["warmup_depth"], o["depth"]]))
        # the lock-step engine carries the gradient of the selected draw instead of
        # re-evaluating it at the start of each of the 16 transitions
        assert int(st["grad_evals"][c]) == o["grad_evals"] - (16 if engine == "tick" else 0)
        assert st["step"][c] == pytest.approx(o["step"], rel=1e-9)


def test_fixed_parameter_sampler_matches_oracle_to_rounding(wb, oracle):
    """With frozen tuning, Gaussian dynamics are purely element-wise; the only
    differences from the oracle are the last bits of the Box-Muller normals (CUDA
    vs glibc log / sincos) and of the reduced energies, so whole chains agree to
    ~1e-12 as long as no accept / U-turn decision flips.  (Bit equality of the
    same source is asserted on CPU by tests/test_host_emulation.py.)"""
    D, C, n = 50, 6, 40
    rng = np.random.default_rng(8)
    var = rng.uniform(0.5, 4.0, D)
    model, target = wb.models.diag_gaussian(var), Target("diag_gaussian", D, prec=1 / var)
    positions = rng.normal(size=(C, D))
    inv_mass = rng.uniform(0.5, 2.0, (C, D))
    with wb.Session(model, C, seed=5, max_trajectory_doublings=6) as s:
        s.init(positions=positions, mass=1 / inv_mass, steps=np.full(C, 0.45))
        s.reserve(n, trace=True)
        s.freeze().sample(n, store=True).sync()
        draws = s.draws(0, n)
        st = s.state()
    exact = 0
    for c in range(C):
        # the frozen step is exp(log(0.45)), one ulp away from 0.45: take the device's
        o = oracle.run_sampler(target, 5, c, positions[c], st["inv_mass"][c], st["step"][c],
                               6, 5, 1, 0.5, n, rng_policy=1)
        scale = np.max(np.abs(o["draws"]), axis=1, keepdims=True)
        exact += int(np.max(np.abs(draws[c] - o["draws"]) / scale) < 1e-11)
    assert exact >= C - 1


# ---- posterior moments -------------------------------------------------------
@pytest.mark.parametrize("kind,D,C", [("std_normal", 100, 512), ("diag_gaussian", 64, 512)])
def test_posterior_moments_match_truth_and_oracle(wb, oracle, kind, D, C):
    rng = np.random.default_rng(D)
    model, target = make_model(wb, kind, D, rng)
    var = 1.0 / target.prec if kind == "diag_gaussian" else np.ones(D)
    nw, ns = 150, 100
    with wb.Session(model, C, seed=2024, max_trajectory_doublings=8) as s:
        s.init(init_radius=2.0)
        s.reserve(ns)
        s.warmup(nw).freeze().sample(ns).sync()
        summ = s.summary(0, ns)
        draws = s.draws(0, ns)
    # analytic truth within MCSE (z < 5 over D dims x 2 moments)
    z_mean = np.abs(summ["mean"]) / summ["mcse"]
    assert np.max(z_mean) < 5.0, np.max(z_mean)
    ess = summ["ess"]
    var_se = var * np.sqrt(2.0 / np.minimum(ess, C * ns))
    assert np.max(np.abs(summ["variance"] - var) / var_se) < 6.0
    assert np.max(summ["r_hat"]) < 1.05
    # the CPU oracle run (8 chains, same settings) agrees within combined MCSE
    cfg = default_config(min_warmup_iter=nw, max_warmup_iter=nw, min_sampling_iter=ns,
                         max_sampling_iter=ns, max_trajectory_doublings=8)
    pos = oracle.init_positions(8, D, 9, 2.0)
    mass, steps = oracle.init_mass_step(target, pos, 9, 1.0)
    cpu = oracle.walnuts(target, cfg, 9, pos, mass, steps)
    chains = [cpu["out"][c, :ns] for c in range(8)]
    cpu_mean = np.mean(np.concatenate(chains), axis=0)
    cpu_mcse = oracle.mcse(chains)
    z = np.abs(summ["mean"] - cpu_mean) / np.sqrt(summ["mcse"] ** 2 + cpu_mcse ** 2)
    assert np.max(z) < 5.0, np.max(z)
    # device summaries equal the oracle's on the same draws
    sub = [draws[c] for c in range(16)]
    np.testing.assert_allclose(wb.ess(sub), oracle.ess(sub), rtol=1e-8)
    np.testing.assert_allclose(wb.r_hat(sub), oracle.r_hat(sub), rtol=1e-10)
    np.testing.assert_allclose(wb.mcse(sub), oracle.mcse(sub), rtol=1e-8)


def test_funnel_moments_match_oracle_within_mcse(wb, oracle):
    D, C, nw, ns = 11, 1024, 200, 200
    over = dict(max_step_halvings=8, max_trajectory_doublings=8)
    with wb.Session(wb.models.funnel(D), C, seed=31, **over) as s:
        s.init(init_radius=1.0)
        s.reserve(ns)
        s.warmup(nw).freeze().sample(ns).sync()
        summ = s.summary(0, ns)
    target = Target("funnel", D)
    cfg = default_config(min_warmup_iter=nw, max_warmup_iter=nw, min_sampling_iter=4 * ns,
                         max_sampling_iter=4 * ns, **over)
    pos = oracle.init_positions(8, D, 4, 1.0)
    mass, steps = oracle.init_mass_step(target, pos, 4, 1.0)
    cpu = oracle.walnuts(target, cfg, 4, pos, mass, steps)
    chains = [cpu["out"][c, :4 * ns] for c in range(8)]
    cpu_mean = np.mean(np.concatenate(chains), axis=0)
    cpu_mcse = oracle.mcse(chains)
    z = np.abs(summ["mean"] - cpu_mean) / np.sqrt(summ["mcse"] ** 2 + cpu_mcse ** 2)
    assert np.max(z) < 5.0, z


# ---- summaries on the reference's golden data -------------------------------
def test_device_summaries_on_reference_golden_vectors(wb):
    """tests/summary_test.cpp:1073-1083, :1182-1192, :866-879 through the C-ABI."""
    from tests.ar1_data import AR1_CHAINS
    ess = wb.ess(AR1_CHAINS)
    assert ess[0] == pytest.approx(96.256789181, abs=1e-5)
    assert ess[1] == pytest.approx(7.315045989, abs=1e-5)
    m = wb.mcse(AR1_CHAINS)
    assert m[0] == pytest.approx(0.096327220756986, abs=1e-7)
    assert m[1] == pytest.approx(0.250085871061602, abs=1e-7)
    ragged = [np.array([[1, 5], [3, 3], [2, 4]], float),
              np.array([[4, 2], [6, 4], [5, 3], [7, 5]], float)]
    np.testing.assert_allclose(wb.r_hat(ragged),
                               [np.sqrt(1 + 147 / 32), np.sqrt(1 + 3 / 32)], rtol=1e-14)
    with pytest.raises(ValueError, match="at least two chains"):
        wb.r_hat([np.arange(10.0).reshape(5, 2)])
    with pytest.raises(ValueError, match="at least 3 draws"):
        wb.ess([np.array([[1.0, 2.0], [3.0, 4.0]])])


def test_device_summaries_on_long_and_ragged_chains(wb, oracle):
    """Chains beyond the shared-memory tile of the autocovariance kernel (3200 draws) and
    of different lengths, against the oracle (summary.hpp:594-769)."""
    rng = np.random.default_rng(12)
    chains = []
    for n, phi in ((5000, 0.9), (4100, 0.5), (3600, -0.3)):
        x = np.zeros((n, 3))
        e = rng.normal(size=(n, 3))
        for i in range(1, n):
            x[i] = phi * x[i - 1] + e[i]
        chains.append(x)
    np.testing.assert_allclose(wb.ess(chains), oracle.ess(chains), rtol=1e-8)
    np.testing.assert_allclose(wb.r_hat(chains), oracle.r_hat(chains), rtol=1e-10)
    np.testing.assert_allclose(wb.mcse(chains), oracle.mcse(chains), rtol=1e-8)


def test_device_summary_errors_match_the_reference(wb):
    """summary_test.cpp:780-806, :1019-1035 through the host-buffer entry points."""
    with pytest.raises(ValueError, match="at least two chains"):
        wb.r_hat([np.arange(10.0).reshape(5, 2)])
    with pytest.raises(ValueError, match="at least 3 draws"):
        wb.r_hat([np.ones((2, 2)), np.ones((3, 2))])
    with pytest.raises(ValueError, match="at least 3 draws"):
        wb.ess([np.array([[1.0, 2.0], [3.0, 4.0]])])


# ---- the reference's behavioural tests through the drop-in entry point -------
@pytest.mark.parametrize("MIN,MAX", [(10, 12), (77, 77), (10, 30)])
def test_warmup_requested_iter(wb, MIN, MAX):  # python/tests/test_pyfunc.py:38-50
    fit = wb.walnuts_device(wb.models.std_normal(2), min_warmup_iter=MIN, max_warmup_iter=MAX,
                            min_sampling_iter=1, max_sampling_iter=1, save_warmup=True)
    for chain in fit:
        assert MIN <= len(chain.warmup.warmup_draws) <= MAX


@pytest.mark.parametrize("MIN,MAX", [(10, 12), (77, 77), (10, 30)])
def test_sampling_requested_iter(wb, MIN, MAX):  # test_pyfunc.py:53-64
    fit = wb.walnuts_device(wb.models.std_normal(2), min_sampling_iter=MIN,
                            max_sampling_iter=MAX, min_warmup_iter=100, max_warmup_iter=100)
    for chain in fit:
        assert MIN <= len(chain) <= MAX


def test_seed_works(wb):  # test_pyfunc.py:89-125
    kw = dict(min_warmup_iter=400, max_warmup_iter=400, save_warmup=True, save_inv_metric=True)
    m = wb.models.std_normal(4)
    fit1 = wb.walnuts_device(m, seed=1234, **kw)
    fit2 = wb.walnuts_device(m, seed=1234, **kw)
    fit3 = wb.walnuts_device(m, seed=452, **kw)
    for c1, c2 in zip(fit1, fit2, strict=True):
        n = min(c1.shape[0], c2.shape[0])
        np.testing.assert_array_equal(c1[:n], c2[:n])
        assert c1.warmup.stepsize == c2.warmup.stepsize
        np.testing.assert_array_equal(c1.warmup.inv_metric, c2.warmup.inv_metric)
        np.testing.assert_array_equal(c1.warmup.warmup_draws, c2.warmup.warmup_draws)
    assert not np.array_equal(fit1[0][:10], fit3[0][:10])


def test_results_do_not_depend_on_sharding(wb):
    """Chains are keyed by global id: 8 chains in one session == 2 sessions of 4
    with chain_offset 0 and 4 (what each rank of a multi-GPU run holds)."""
    D, n = 20, 30
    m = wb.models.ill_conditioned_gaussian(D, 100.0)

    def run(C, off):
        with wb.Session(m, C, seed=9, chain_offset=off) as s:
            s.init(init_radius=2.0)
            s.reserve(n)
            s.warmup(40).freeze().sample(n).sync()
            return s.draws(0, n)

    whole = run(8, 0)
    np.testing.assert_array_equal(whole[:4], run(4, 0))
    np.testing.assert_array_equal(whole[4:], run(4, 4))


# ---- lock-step tick engine and the tensor-core logistic gradient --------------
@pytest.fixture
def tick_engine_env(monkeypatch):
    monkeypatch.setenv("WB200_ENGINE", "tick")


@pytest.mark.parametrize("kind,D,C,over,nw,ns", [
    ("std_normal", 100, 4, dict(), 40, 40),
    ("diag_gaussian", 10, 6, dict(max_trajectory_doublings=8), 60, 60),
    ("funnel", 11, 6, dict(max_step_halvings=8, max_trajectory_doublings=7), 50, 50),
    ("diag_gaussian", 300, 3, dict(), 25, 25),
])
def test_tick_engine_trajectories_match_oracle(wb, oracle, tick_engine_env, kind, D, C, over,
                                               nw, ns):
    """The lock-step engine (one gradient request per tick per chain) through the same
    session API, element-wise targets: identical-seed trajectories vs the oracle."""
    rng = np.random.default_rng(2000 + D)
    model, target = make_model(wb, kind, D, rng)
    positions = rng.normal(size=(C, D))
    mass = rng.uniform(0.5, 2.0, (C, D))
    steps = rng.uniform(0.2, 0.6, C)
    cfg = default_config(**over)
    with wb.Session(model, C, seed=99, **over) as s:
        s.init(positions=positions, mass=mass, steps=steps)
        s.reserve(nw + ns, trace=True)
        s.warmup(nw, store=True).freeze().sample(ns, store=True).sync()
        draws = s.draws(0, nw + ns)
        tr = s.trace(0, nw + ns)
        st = s.state()
    full = 0
    for c in range(C):
        o = oracle.run_chain(target, cfg, 99, c, positions[c], mass[c], steps[c], nw, ns,
                             rng_policy=1)
        ref_draws = np.concatenate([o["warmup_draws"], o["draws"]])
        k = first_divergence(draws[c], ref_draws, 1e-9)
        assert k >= 5
        if k == nw + ns:
            full += 1
            np.testing.assert_array_equal(tr["depth"][c],
                                          np.concatenate([o["warmup_depth"], o["depth"]]))
            assert st["step"][c] == pytest.approx(o["step"], rel=1e-10)
            # the carried gradient saves the reference's re-evaluation per transition
            assert int(st["grad_evals"][c]) == o["grad_evals"] - (nw + ns)
    assert full >= (C + 1) // 2


def bf16_round(a):
    import torch
    return torch.tensor(a, dtype=torch.float64).to(torch.bfloat16).to(torch.float64).numpy()


def make_logistic(N, D, seed):
    rng = np.random.default_rng(seed)
    X = bf16_round(rng.normal(size=(N, D)))
    tstar = rng.normal(size=D) / np.sqrt(D)
    y = (rng.uniform(size=N) < 1 / (1 + np.exp(-X @ tstar))).astype(np.float64)
    return X, y


@pytest.mark.parametrize("N,D,C", [(256, 64, 128), (300, 16, 5), (1000, 200, 130),
                                   (4096, 512, 256), (1, 1, 1)])
def test_logistic_gradient_operator_matches_oracle(wb, oracle, N, D, C):
    """tcgen05 GEMM path vs the fp64 CPU density.  Tolerances: logp is accumulated in
    fp32 tiles / fp64 totals from a bf16 hi+lo split of theta (1e-5 relative); the
    residual r = y - sigmoid(z) enters the second GEMM in bf16, i.e. 2^-9 relative per
    element (3e-3 of the largest gradient component)."""
    from walnuts_b200.sampler import logistic_logp_grad
    X, y = make_logistic(N, D, N + D)
    theta = np.random.default_rng(C).normal(size=(C, D)) * 0.3
    lp, g, _ = logistic_logp_grad(X, y, theta)
    t = Target("logistic", D, X=X, y=y)
    for c in range(min(C, 8)):
        olp, og = oracle.logp_grad(t, theta[c])
        assert abs(lp[c] - olp) <= 1e-5 * max(1.0, abs(olp))
        assert np.max(np.abs(g[c] - og)) <= 3e-3 * max(np.max(np.abs(og)), 1.0)


def test_logistic_sampler_moments_match_oracle_within_mcse(wb, oracle):
    N, D, C, nw, ns = 500, 8, 256, 150, 100
    X, y = make_logistic(N, D, 77)
    over = dict(max_trajectory_doublings=8)
    with wb.Session(wb.models.logistic(X, y), C, seed=5, **over) as s:
        s.init(init_radius=1.0)
        s.reserve(ns)
        s.warmup(nw).freeze().sample(ns).sync()
        summ = s.summary(0, ns)
        evals = s.counters()["grad_evals"]
    assert evals > C * (nw + ns)
    assert np.max(summ["r_hat"]) < 1.05
    target = Target("logistic", D, X=X, y=y)
    cfg = default_config(min_warmup_iter=nw, max_warmup_iter=nw, min_sampling_iter=4 * ns,
                         max_sampling_iter=4 * ns, **over)
    pos = oracle.init_positions(8, D, 4, 1.0)
    mass, steps = oracle.init_mass_step(target, pos, 4, 1.0)
    cpu = oracle.walnuts(target, cfg, 4, pos, mass, steps)
    chains = [cpu["out"][c, :4 * ns] for c in range(8)]
    cpu_mean = np.mean(np.concatenate(chains), axis=0)
    cpu_var = np.var(np.concatenate(chains), axis=0, ddof=1)
    cpu_mcse = oracle.mcse(chains)
    z = np.abs(summ["mean"] - cpu_mean) / np.sqrt(summ["mcse"] ** 2 + cpu_mcse ** 2)
    assert np.max(z) < 5.0, z
    assert np.max(np.abs(summ["variance"] / cpu_var - 1)) < 0.25


def test_free_running_ticks_give_ragged_draws_with_the_same_posterior(wb, oracle):
    """Lock-step budget mode: exactly n ticks, chains roll straight into their next
    transition, so draw counts are ragged (as the reference's adaptive runs are) while
    every lane stays busy; the posterior is the same as with iteration quotas."""
    N, D, C = 400, 6, 192
    X, y = make_logistic(N, D, 3)
    with wb.Session(wb.models.logistic(X, y), C, seed=8, max_trajectory_doublings=7) as s:
        s.init(init_radius=0.5)
        s.reserve(400)
        s.warmup(120).freeze()
        before = s.counters()["grad_evals"]
        s.sample_ticks(600).sync()
        evals = s.counters()["grad_evals"] - before
        rows = s.chain_rows()
        ragged = s.summary_ragged(0)
    # every chain consumes one gradient per tick (the very first tick only posts a request)
    assert 599 * C <= evals <= 600 * C
    assert rows.min() >= 3 and rows.max() > rows.min()   # ragged
    target = Target("logistic", D, X=X, y=y)
    cfg = default_config(min_warmup_iter=120, max_warmup_iter=120, min_sampling_iter=400,
                         max_sampling_iter=400, max_trajectory_doublings=7)
    pos = oracle.init_positions(8, D, 4, 0.5)
    mass, steps = oracle.init_mass_step(target, pos, 4, 1.0)
    cpu = oracle.walnuts(target, cfg, 4, pos, mass, steps)
    chains = [cpu["out"][c, :400] for c in range(8)]
    cpu_mean = np.mean(np.concatenate(chains), axis=0)
    cpu_mcse = oracle.mcse(chains)
    z = np.abs(ragged["mean"] - cpu_mean) / np.sqrt(ragged["mcse"] ** 2 + cpu_mcse ** 2)
    assert np.max(z) < 5.0, z
    assert np.max(ragged["r_hat"]) < 1.05


def test_free_running_warmup_adapts_every_chain(wb, oracle):
    """wb200_session_warmup_ticks: adaptive warm-up on a tick budget; chains adapt over
    ragged transition counts, freeze abandons the transitions in flight, and the frozen
    step sizes / metrics give the same posterior as an iteration-quota warm-up."""
    N, D, C = 400, 6, 192
    X, y = make_logistic(N, D, 3)
    with wb.Session(wb.models.logistic(X, y), C, seed=21, max_trajectory_doublings=7) as s:
        s.init(init_radius=0.5)
        s.reserve(400)
        s.warmup_ticks(2500).freeze()
        st = s.state()
        s.sample_ticks(800).sync()
        ragged = s.summary_ragged(0)
    with wb.Session(wb.models.logistic(X, y), C, seed=21, max_trajectory_doublings=7) as s:
        s.init(init_radius=0.5)
        s.warmup(150).freeze()
        quota = s.state()
    # the adapted step sizes and metrics agree in distribution with the quota warm-up's
    assert np.all(np.isfinite(st["step"])) and np.all(st["step"] > 0)
    assert abs(np.median(st["step"]) / np.median(quota["step"]) - 1) < 0.15
    assert np.max(np.abs(np.median(st["inv_mass"], axis=0) /
                         np.median(quota["inv_mass"], axis=0) - 1)) < 0.25
    target = Target("logistic", D, X=X, y=y)
    cfg = default_config(min_warmup_iter=120, max_warmup_iter=120, min_sampling_iter=400,
                         max_sampling_iter=400, max_trajectory_doublings=7)
    pos = oracle.init_positions(8, D, 4, 0.5)
    mass, steps = oracle.init_mass_step(target, pos, 4, 1.0)
    cpu = oracle.walnuts(target, cfg, 4, pos, mass, steps)
    chains = [cpu["out"][c, :400] for c in range(8)]
    cpu_mean = np.mean(np.concatenate(chains), axis=0)
    cpu_mcse = oracle.mcse(chains)
    z = np.abs(ragged["mean"] - cpu_mean) / np.sqrt(ragged["mcse"] ** 2 + cpu_mcse ** 2)
    assert np.max(z) < 5.0, z
    assert np.max(ragged["r_hat"]) < 1.05


def test_one_shot_readback_is_identical_for_pinned_and_pageable_buffers(wb):
    """walnutpie_sample_device copies the draws back block by block while sampling runs
    (2-D copies when D is even, 3-D strided copies when rows are padded): the result
    must not depend on the kind of host buffer, and must equal the session's draws."""
    import ctypes
    from walnuts_b200 import _ffi
    for D in (6, 7):
        C, nw, ns = 12, 23, 37
        model = wb.models.diag_gaussian(np.linspace(0.5, 3.0, D))
        desc = model.desc()
        outs = []
        for pinned in (False, True):
            out = _ffi.pinned_empty((C, nw + ns, D)) if pinned else np.zeros((C, nw + ns, D))
            out[...] = -7.0
            lengths, stepsize = np.zeros(2 * C, np.int32), np.zeros(C)
            _ffi._ffi_sample_device(
                ctypes.byref(desc), D, None, C, 31, 2, 2.0, None, nw, nw, ns, ns, 10, 5, 1,
                0.5, 0.1, 1.0, 1.01, 4.0, 1e-5, 15.0, 1.0, 0.8, 0.05, 0.8, 0.9, 1e-4, 0.5,
                True, out, out.size, lengths, stepsize, None, 0, _ffi.print_callback)
            assert np.all(lengths[:C] == nw) and np.all(lengths[C:] == ns)
            outs.append(np.array(out))
        np.testing.assert_array_equal(outs[0], outs[1])
        assert not np.any(outs[0] == -7.0)
        with wb.Session(model, C, seed=31 + 2 + C) as s:
            s.init(init_radius=2.0)
            s.reserve(nw + ns)
            s.warmup(nw, store=True).freeze().sample(ns).sync()
            np.testing.assert_array_equal(s.draws(0, nw + ns), outs[0])


# ---- the caller's own density, batched on the device (WalnutModelDesc kind 4) ---------
def test_torch_density_reproduces_the_builtin_target(wb, tick_engine_env):
    """A diagonal Gaussian written in PyTorch and plugged in through the batched-density
    callback starts on the same trajectories as the built-in target on the lock-step engine
    (same Philox streams).  The two densities round differently (order of the products and
    of the logp sum), so the runs separate at rounding level after a few transitions; from
    then on they must agree in distribution (adapted step sizes)."""
    import torch
    D, C, nw, ns = 24, 6, 50, 50
    var = np.linspace(0.3, 4.0, D)
    prec = torch.tensor(1.0 / var, dtype=torch.float64, device="cuda")

    def grad_fn(theta):
        return -0.5 * (theta * theta * prec).sum(dim=1), -(theta * prec)

    runs = []
    for model in (wb.models.diag_gaussian(var),
                  wb.models.torch_density(D, None, grad_fn=grad_fn),
                  wb.models.torch_density(D, lambda t: -0.5 * (t * t * prec).sum(dim=1))):
        with wb.Session(model, C, seed=77) as s:
            s.init(init_radius=1.5)
            s.reserve(nw + ns)
            s.warmup(nw, store=True).freeze().sample(ns).sync()
            runs.append((s.draws(0, nw + ns), s.state()))
    ref, ref_state = runs[0]
    for draws, state in runs[1:]:
        np.testing.assert_allclose(draws[:, 0], ref[:, 0], rtol=1e-12, atol=1e-13)
        it = first_divergence(draws, ref, 1e-9)
        assert it is None or it >= 3, f"diverged at iteration {it}"
        assert abs(np.median(state["step"]) / np.median(ref_state["step"]) - 1) < 0.2


def test_torch_density_samples_a_correlated_gaussian(wb):
    """A target none of the built-in kinds can express: N(mu, Sigma) with a dense
    covariance, defined in PyTorch; posterior moments within MCSE of the truth, through
    the reference-shaped one-shot call."""
    import torch
    D, C = 5, 64
    rng = np.random.default_rng(2)
    A = rng.normal(size=(D, D))
    Sigma = A @ A.T / D + 0.5 * np.eye(D)
    mu = rng.normal(size=D)
    P = torch.tensor(np.linalg.inv(Sigma), dtype=torch.float64, device="cuda")
    m = torch.tensor(mu, dtype=torch.float64, device="cuda")

    def logp(theta):
        d = theta - m
        return -0.5 * ((d @ P) * d).sum(dim=1)

    fit = wb.walnuts_device(wb.models.torch_density(D, logp), num_chains=C, seed=3,
                            min_warmup_iter=150, max_warmup_iter=150,
                            min_sampling_iter=150, max_sampling_iter=150)
    chains = [np.asarray(c) for c in fit]
    allx = np.concatenate(chains)
    mcse = wb.mcse(chains)
    assert np.max(np.abs(allx.mean(axis=0) - mu) / mcse) < 5.0
    cov = np.cov(allx.T)
    assert np.max(np.abs(cov - Sigma)) < 0.15 * np.max(np.abs(Sigma))
    assert np.max(wb.r_hat(chains)) < 1.05


def test_exceptions_inside_a_batched_density_surface_as_themselves(wb):
    calls = {"n": 0}

    def bad(C, D, ld, theta, grad, lp, stream):
        calls["n"] += 1
        if calls["n"] >= 2:   # the first evaluation of the step-size search: initialisation
            raise KeyError("boom")
        import torch
        with torch.cuda.stream(torch.cuda.ExternalStream(int(stream or 0))):
            wb.models.torch_density  # noqa: B018 (the adapter is exercised elsewhere)
            g = torch.as_tensor(wb.models._DevicePointer(grad, (C, ld)), device="cuda")
            t = torch.as_tensor(wb.models._DevicePointer(theta, (C, ld)), device="cuda")
            l = torch.as_tensor(wb.models._DevicePointer(lp, (C,)), device="cuda")
            g.copy_(-t)
            l.copy_(-0.5 * (t * t).sum(dim=1))

    with wb.Session(wb.models.batch_callback(3, bad), 4, seed=1) as s:
        with pytest.raises(KeyError, match="boom"):
            s.init(init_radius=1.0)
            s.warmup(5)


# ---- BASELINE.json sizes ---------------------------------------------------------------
def test_c4_size_gradient_operator_matches_oracle(wb, oracle):
    """The tensor-core gradient at the full c4 shape (N = 100 000 data rows: 391 column
    tiles with 96 padded rows, D = 512, 8192 chains: 64 row tiles) against the fp64 CPU
    density for chains spread over the batch."""
    import sys
    sys.path.insert(0, ".")
    import bench
    from walnuts_b200.sampler import logistic_logp_grad
    N, D, C = 100_000, 512, 8192
    X, y = bench.logistic_data(N, D)
    theta = np.random.default_rng(4).normal(size=(C, D)) / np.sqrt(D)
    lp, g, _ = logistic_logp_grad(X, y, theta)
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(g))
    t = Target("logistic", D, X=X, y=y)
    worst = 0.0
    for c in (0, 127, 128, 4095, 4096, C - 1):
        olp, og = oracle.logp_grad(t, theta[c])
        worst = max(worst, abs(lp[c] - olp))
        assert abs(lp[c] - olp) <= 2e-6 * abs(olp)     # |logp| ~ 7e4: 0.14 absolute
        assert np.max(np.abs(g[c] - og)) <= 3e-3 * np.max(np.abs(og))
    # What the sampler acts on is the energy DIFFERENCE along an orbit (compared with
    # max_hamiltonian_error = 0.5): the error of logp is a smooth function of theta (bf16
    # split of theta, fp32 accumulation), so it cancels between nearby points.
    theta2 = theta.copy()
    theta2[1::2] = theta[0::2] + 0.02 * np.random.default_rng(5).normal(size=(C // 2, D))
    lp2, _, _ = logistic_logp_grad(X, y, theta2)
    worst_diff = 0.0
    for c in (0, 126, 2048, 4094, C - 2):
        o0, _ = oracle.logp_grad(t, theta2[c])
        o1, _ = oracle.logp_grad(t, theta2[c + 1])
        worst_diff = max(worst_diff, abs((lp2[c + 1] - lp2[c]) - (o1 - o0)))
    print(f"\nlogp at c4 size: worst |device - fp64| = {worst:.4f}; worst error of the "
          f"logp difference between points one step apart = {worst_diff:.5f}")
    assert worst_diff <= 0.02


def test_c2_size_posterior_and_sharding(wb):
    """c2 at full size (D = 1000, cond 1e4, 4096 chains): variances within 10 % after a
    short run, and the last 64 chains equal those of a 64-chain session with the same
    global chain ids (what another rank would hold)."""
    D, C = 1000, 4096
    model = wb.models.ill_conditioned_gaussian(D, 1e4)
    tune = dict(max_trajectory_doublings=10, max_step_halvings=5)
    with wb.Session(model, C, seed=11, **tune) as s:
        s.init(init_radius=2.0)
        s.reserve(40)
        s.warmup(150).freeze().sample(40).sync()
        summ = s.summary(0, 40)
        tail = s.draws(0, 40)[C - 64:]
    var = 1e4 ** (np.arange(D) / (D - 1))
    assert np.max(np.abs(summ["variance"] / var - 1)) < 0.10
    assert np.max(np.abs(summ["mean"]) / np.sqrt(var)) < 0.05
    with wb.Session(model, 64, seed=11, chain_offset=C - 64, **tune) as s:
        s.init(init_radius=2.0)
        s.reserve(40)
        s.warmup(150).freeze().sample(40).sync()
        np.testing.assert_array_equal(s.draws(0, 40), tail)


def test_ctrl_c_interrupts_a_one_shot_run(wb):
    """interrupts.hpp / errors.hpp:30-48: SIGINT during sampling ends the call with the
    `interrupt` error, which the Python layer raises as KeyboardInterrupt; the previous
    handler is back afterwards."""
    import os
    import signal
    import threading
    import time
    before = signal.getsignal(signal.SIGINT)
    timer = threading.Timer(0.7, lambda: os.kill(os.getpid(), signal.SIGINT))
    t0 = time.perf_counter()
    timer.start()
    try:
        with pytest.raises(KeyboardInterrupt):
            wb.walnuts_device(wb.models.ill_conditioned_gaussian(1000, 1e4), num_chains=4096,
                              seed=1, min_warmup_iter=200000, max_warmup_iter=200000,
                              max_trajectory_doublings=10, max_sampling_iter=5,
                              min_sampling_iter=5)
    finally:
        timer.cancel()
    assert time.perf_counter() - t0 < 20.0
    assert signal.getsignal(signal.SIGINT) is before
