"""fp32 mode of the chain-resident kernel (BASELINE.json north_star: "fp32 mode within a
stated tolerance"; config c3 "fp64 and fp32 variants"): WalnutTuning::precision = 1 /
models.*(dtype="f32").  The integrator state and its element-wise arithmetic are single
precision inside a transition; energies and U-turn dots are accumulated across threads in
fp64, scalar decisions and the adaptation statistics are fp64, draws are stored as fp64.

Stated tolerances (asserted below):
 * fixed-step orbits: max-norm relative error vs the reference's fp64 integrator (oracle)
   <= 2e-6 * sqrt(steps); vs the oracle's fp32 restatement of the same orbit <= 1e-6 for
   the Gaussians (summation order and per-thread partial sums only; half the fp64
   tolerance for the funnel, whose exp(-v) amplifies them);
 * trajectories: the fp32 chain follows the fp64 chain of the same seed to 1e-3 relative
   over the first transitions, then separates (rounding-induced branch flips come sooner);
 * posterior moments within Monte Carlo standard error, as in fp64."""
import numpy as np
import pytest

from oracle.binding import Target

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _device_arithmetic_policy(oracle):
    with oracle.fused_arith(True):
        yield


def make(wb, kind, D, dtype):
    if kind == "diag_gaussian":
        var = 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
        return wb.models.diag_gaussian(var, dtype=dtype), Target(kind, D, prec=1 / var)
    return getattr(wb.models, kind)(D, dtype=dtype), Target(kind, D)


@pytest.mark.parametrize("kind,D,step,nsteps", [
    ("std_normal", 100, 0.37, 64), ("diag_gaussian", 1000, 0.2, 50), ("funnel", 100, 0.05, 40),
    ("diag_gaussian", 513, 0.1, 33), ("funnel", 2, 0.1, 20), ("std_normal", 4096, 0.1, 16)])
def test_fp32_orbits_stay_within_the_stated_tolerance(wb, oracle, kind, D, step, nsteps):
    rng = np.random.default_rng(D + nsteps)
    model, target = make(wb, kind, D, "f32")
    C = 4
    inv_mass = rng.uniform(0.5, 2.0, (C, D))
    if kind == "diag_gaussian":
        inv_mass *= 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
    theta = rng.normal(size=(C, D)) * (0.3 if kind == "funnel" else 1.0)
    rho = rng.normal(size=(C, D)) / np.sqrt(inv_mass)
    th, rh, g, lp, jt = wb.orbit(model, theta, rho, inv_mass, step, nsteps)
    tol64 = 2e-6 * np.sqrt(nsteps)
    worst64 = worst32 = 0.0
    for c in range(C):
        o64 = oracle.orbit(target, theta[c], rho[c], inv_mass[c], step, nsteps)
        o32 = oracle.orbit(target, theta[c], rho[c], inv_mass[c], step, nsteps, f32=True)
        for dev, a64, a32 in ((th[c], o64[0], o32[0]), (rh[c], o64[1], o32[1]),
                              (g[c], o64[2], o32[2])):
            # per-coordinate scale of an orbit: relative to the largest component
            worst64 = max(worst64, np.max(np.abs(dev - a64)) / np.max(np.abs(a64)))
            worst32 = max(worst32, np.max(np.abs(dev - a32)) / np.max(np.abs(a32)))
        assert abs(jt[c] - o64[4]) <= 1e-5 * max(abs(o64[4]), 1.0) + 1e-4
        assert abs(lp[c] - o32[3]) <= 1e-5 * max(abs(o32[3]), 1.0)
    print(f"\n[{kind} D={D}, {nsteps} steps] fp32 vs fp64 oracle {worst64:.2e} (tolerance "
          f"{tol64:.1e}); vs fp32 oracle {worst32:.2e}")
    assert worst64 <= tol64
    # against the same orbit restated in float: summation order / partial sums only (the
    # funnel's exp(-v) amplifies them)
    assert worst32 <= (tol64 / 2 if kind == "funnel" else 1e-6 * max(1.0, np.sqrt(nsteps) / 4))


def test_fp32_chain_follows_the_fp64_chain_then_separates(wb):
    D, C, n = 50, 8, 30
    rng = np.random.default_rng(4)
    var = rng.uniform(0.5, 4.0, D)
    pos = rng.normal(size=(C, D))
    mass = rng.uniform(0.5, 2.0, (C, D))
    out = {}
    for dtype in ("f64", "f32"):
        with wb.Session(wb.models.diag_gaussian(var, dtype=dtype), C, seed=5) as s:
            s.init(positions=pos, mass=mass, steps=np.full(C, 0.4))
            s.reserve(n)
            s.freeze().sample(n).sync()
            out[dtype] = s.draws(0, n)
    rel = np.max(np.abs(out["f32"] - out["f64"]), axis=2) / np.max(np.abs(out["f64"]), axis=2)
    first = [int(np.argmax(r > 1e-3)) if np.any(r > 1e-3) else n for r in rel]
    print(f"\nfirst iteration at which the fp32 chain leaves the fp64 chain (1e-3): {first}")
    assert np.max(rel[:, 0]) < 1e-4
    assert np.median(first) >= 3


@pytest.mark.parametrize("kind,D,C", [("diag_gaussian", 64, 512), ("std_normal", 100, 512)])
def test_fp32_posterior_moments_within_mcse(wb, kind, D, C):
    model, target = make(wb, kind, D, "f32")
    var = 1.0 / target.prec if kind == "diag_gaussian" else np.ones(D)
    nw, ns = 150, 100
    with wb.Session(model, C, seed=2024, max_trajectory_doublings=8) as s:
        s.init(init_radius=2.0)
        s.reserve(ns)
        s.warmup(nw).freeze().sample(ns).sync()
        summ = s.summary(0, ns)
        st = s.state()
    assert np.max(np.abs(summ["mean"]) / summ["mcse"]) < 5.0
    var_se = var * np.sqrt(2.0 / np.minimum(summ["ess"], C * ns))
    assert np.max(np.abs(summ["variance"] - var) / var_se) < 6.0
    assert np.max(summ["r_hat"]) < 1.05
    assert np.all(np.isfinite(st["step"])) and np.all(st["step"] > 0)


def test_fp32_funnel_keeps_the_exact_posterior_at_c3_size(wb):
    """c3's fp32 variant at full size: started in the exact funnel with fixed tuning, the
    fp32 transition keeps it (as the fp64 one does in tests/test_gpu_c3.py)."""
    D, C = 100, 16384
    rng = np.random.default_rng(1)
    v = rng.normal(0.0, 3.0, C)
    inits = np.concatenate([v[:, None], rng.normal(size=(C, D - 1)) * np.exp(0.5 * v)[:, None]],
                           axis=1)
    with wb.Session(wb.models.funnel(D, dtype="f32"), C, seed=3, max_step_halvings=8,
                    max_trajectory_doublings=10) as s:
        s.init(positions=inits, mass=np.ones((C, D)), steps=np.full(C, 0.4))
        s.reserve(1)
        s.freeze()
        s.sample(40, store=False)
        s.sample(1).sync()
        last = s.draws(0, 1)[:, 0]
    vv = last[:, 0]
    zm = vv.mean() / np.sqrt(9.0 / C)
    zv = (vv.var(ddof=1) - 9.0) / (9.0 * np.sqrt(2.0 / C))
    z1 = last[:, 1] * np.exp(-0.5 * vv)
    print(f"\nfp32 funnel after 41 transitions: E[v] z = {zm:+.2f}, Var[v] z = {zv:+.2f}, "
          f"standardised x_1 mean z = {z1.mean() * np.sqrt(C):+.2f}")
    assert abs(zm) < 4.5 and abs(zv) < 4.5 and abs(z1.mean() * np.sqrt(C)) < 4.5


def test_fp32_is_refused_where_it_is_not_implemented(wb):
    X = np.random.default_rng(0).normal(size=(300, 8))
    y = (np.random.default_rng(1).uniform(size=300) < 0.5).astype(float)
    with pytest.raises(ValueError, match="fp32 mode"):
        wb.Session(wb.models.logistic(X, y), 4, seed=1, precision=1)
