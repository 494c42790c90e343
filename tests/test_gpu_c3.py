"""BASELINE.json config c3 at spec: Neal's funnel, D = 100, 16 384 chains, fp64.

The funnel's neck makes v = theta[0] mix slowly under a diagonal metric: measured here on
both arms, the chain means of v relax with a time constant of about 1200 iterations
after the adaptive phase (which holds E[v] near +1.3), so the spec'd 300 + 200 schedule
is a transient for the reference and for the device alike.  Three checks:
 (1) started in the exact posterior with fixed tuning, the device transition keeps it
     (invariance of the kernel, at full size);
 (2) on the spec'd schedule the device's moments agree with the CPU reference arm
     (unmodified reference headers, mt19937_64) within combined Monte Carlo error;
 (3) run long enough (tau reported) the device reaches the true moments.
Monte Carlo errors are cross-chain (chains are independent): sd of the per-chain
statistic / sqrt(chains)."""
import numpy as np
import pytest

from oracle.binding import Target, default_config

pytestmark = pytest.mark.gpu

D, C = 100, 16384
TUNE = dict(max_step_halvings=8, max_trajectory_doublings=10)


def exact_funnel(rng, n, d):
    v = rng.normal(0.0, 3.0, n)
    x = rng.normal(size=(n, d - 1)) * np.exp(0.5 * v)[:, None]
    return np.concatenate([v[:, None], x], axis=1)


def z_scores(sample, mean0, var0):
    """z of the sample mean and of the sample variance of i.i.d. values"""
    n = len(sample)
    zm = (sample.mean() - mean0) / np.sqrt(var0 / n)
    zv = (sample.var(ddof=1) - var0) / (var0 * np.sqrt(2.0 / n))
    return zm, zv


def test_c3_transition_keeps_the_exact_funnel_posterior(wb):
    rng = np.random.default_rng(1)
    inits = exact_funnel(rng, C, D)
    with wb.Session(wb.models.funnel(D), C, seed=3, **TUNE) as s:
        s.init(positions=inits, mass=np.ones((C, D)), steps=np.full(C, 0.4))
        s.reserve(1)
        s.freeze()
        s.sample(40, store=False)
        s.sample(1).sync()
        last = s.draws(0, 1)[:, 0]
        evals = s.counters()["grad_evals"] / (41 * C)
    v = last[:, 0]
    zm, zv = z_scores(v, 0.0, 9.0)
    z1 = last[:, 1] * np.exp(-0.5 * v)          # x_1 e^{-v/2} ~ N(0, 1)
    zm1, zv1 = z_scores(z1, 0.0, 1.0)
    print(f"\nafter 41 transitions from the exact posterior: E[v] z = {zm:+.2f}, Var[v] z = "
          f"{zv:+.2f}, standardised x_1: {zm1:+.2f} / {zv1:+.2f}; {evals:.1f} gradients per "
          f"transition; moved: {np.mean(v != inits[:, 0]):.3f}")
    assert np.mean(v != inits[:, 0]) > 0.95
    assert max(abs(zm), abs(zv), abs(zm1), abs(zv1)) < 4.5


def test_c3_spec_schedule_agrees_with_the_cpu_reference_arm(wb, oracle):
    """300 adaptive + 200 sampling iterations from N(0, 1) starts, as c3 is specified: the
    device's 16 384 chains and 512 chains of the CPU reference sampler (the unmodified
    headers where built, else the oracle port) agree on the moments of v at the end of the
    run within combined Monte Carlo error -- transient bias included, it is the same
    algorithm."""
    from oracle.binding import load_ref
    cpu = load_ref() or oracle
    nw, ns, Cc = 300, 200, 512
    with wb.Session(wb.models.funnel(D), C, seed=20250, **TUNE) as s:
        s.init(init_radius=1.0)
        s.reserve(ns)
        s.warmup(nw).freeze().sample(ns).sync()
        gv = s.draws(0, ns)[:, :, 0]
    target = Target("funnel", D)
    cfg = default_config(min_warmup_iter=nw, max_warmup_iter=nw, min_sampling_iter=ns,
                         max_sampling_iter=ns, **TUNE)
    pos = cpu.init_positions(Cc, D, 11, 1.0)
    mass, steps = cpu.init_mass_step(target, pos, 11, 1.0)
    cv = cpu.walnuts(target, cfg, 11, pos, mass, steps)["out"][:, :ns, 0]
    lines = []
    for name, stat in (("mean", lambda a: a.mean(1)), ("variance", lambda a: a.var(1, ddof=1))):
        g, c = stat(gv[:, -50:]), stat(cv[:, -50:])
        se = np.sqrt(g.var(ddof=1) / len(g) + c.var(ddof=1) / len(c))
        z = (g.mean() - c.mean()) / se
        lines.append(f"{name} of v over the last 50 draws: device {g.mean():+.3f}, CPU "
                     f"{c.mean():+.3f}, z = {z:+.2f}")
        assert abs(z) < 4.5, lines
    print("\n" + "\n".join(lines))
    # both arms are still far from E[v] = 0 at this point: the agreement is not trivial
    assert gv[:, -50:].mean() > 0.5


def test_c3_long_run_reaches_the_true_moments(wb):
    """enough iterations for the transient to die (tau reported): E[v] = 0, Var[v] = 9,
    E[x_i] = 0 within cross-chain Monte Carlo error at 16 384 chains"""
    nw, burn, ns = 300, 7000, 60
    with wb.Session(wb.models.funnel(D), C, seed=20250, **TUNE) as s:
        s.init(init_radius=1.0)
        s.reserve(ns)
        s.warmup(nw).freeze()
        s.sample(burn, store=False)
        s.stream_begin(32)
        s.sample(ns).sync()
        draws = s.draws(0, ns)
        summ = s.stream_summary()
    v = draws[:, -1, 0]
    zm, zv = z_scores(v, 0.0, 9.0)
    tau = C * ns / summ["ess"][0]
    zx = np.abs(draws[:, -1, 1:].mean(0)) / (draws[:, -1, 1:].std(0, ddof=1) / np.sqrt(C))
    print(f"\nafter {nw} + {burn} iterations: E[v] = {v.mean():+.4f} (z {zm:+.2f}), Var[v] = "
          f"{v.var(ddof=1):.3f} (z {zv:+.2f}), max z of E[x_i] = {zx.max():.2f}; "
          f"integrated autocorrelation time of v over {ns} draws >= {tau:.0f} iterations "
          f"(lag window cut: {bool(summ['truncated'][0])})")
    assert abs(zm) < 4.5 and abs(zv) < 4.5
    assert zx.max() < 5.0
