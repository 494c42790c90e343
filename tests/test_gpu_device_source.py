"""Model kind 5 on the GPU: a density given as CUDA source runs as the Target of the
chain-resident transition kernel (the reference's contract for a
density is LogpGrad, concepts.hpp:25-60, exercised by examples/walnutpie_api.cpp:39-43).

Parity: (i) a source that restates the built-in diagonal Gaussian target produces the
built-in target's chains bit for bit -- and those are pinned to the oracle in
test_gpu_parity.py; (ii) the element-wise form against the oracle's density on fixed-step
orbits (<= 1e-12 relative) and along identically seeded trajectories; (iii) a density that
is not built in against a numpy restatement of the leapfrog (walnuts.hpp:329-332) and its
known posterior moments."""
import numpy as np
import pytest

from oracle.binding import Target, default_config

pytestmark = pytest.mark.gpu

ELEMENTWISE_GAUSS = r'''
__device__ void wb200_logp_grad(int d, double x, const double* par, double& lp, double& g) {
  const double t = x * par[d];
  lp = -0.5 * x * t;
  g = -t;
}
'''

# the built-in DiagGaussianTargetT, restated through the full Target interface
FULL_GAUSS = r'''
template <int T, int K, class Real>
struct MyGauss {
  Real prec[K][2];
  __device__ void init(const wb200::ChainParams& p, int tid) {
    for (int k = 0; k < K; ++k)
      for (int v = 0; v < 2; ++v) {
        const int d = 2 * (tid + k * T) + v;
        prec[k][v] = d < p.D ? static_cast<Real>(p.tparam[d]) : static_cast<Real>(0);
      }
  }
  __device__ void grad(const Real (&th)[K][2], Real (&g)[K][2], Real& lp_part,
                       wb200::Group<T>&) const {
    Real s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        Real t = wb200::mul_rn(th[k][v], prec[k][v]);
        s = wb200::madd(th[k][v], t, s);
        g[k][v] = -t;
      }
    }
    lp_part = static_cast<Real>(-0.5) * s;
  }
};
#define WB200_USER_TARGET MyGauss
'''

# logistic distribution with location par[d]: lp = -|z| - 2 log(1 + exp(-|z|)), z = x - par[d]
LOGISTIC_DIST = r'''
__device__ void wb200_logp_grad(int d, double x, const double* par, double& lp, double& g) {
  const double z = x - par[d];
  const double e = exp(-fabs(z));
  lp = -fabs(z) - 2.0 * log1p(e);
  const double t = (1.0 - e) / (1.0 + e);
  g = z >= 0 ? -t : t;
}
'''


@pytest.fixture(autouse=True, scope="module")
def _device_arithmetic_policy(oracle):
    with oracle.fused_arith(True):
        yield


@pytest.mark.parametrize("D,C", [(10, 12), (100, 40), (1000, 24)])
def test_full_target_source_equals_the_built_in_target_bitwise(wb, D, C):
    var = 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
    rng = np.random.default_rng(D)
    pos, mass, steps = rng.normal(size=(C, D)), rng.uniform(0.5, 2, (C, D)), np.full(C, 0.3)
    out = []
    for model in (wb.models.diag_gaussian(var),
                  wb.models.device_source(FULL_GAUSS, D, params=1.0 / var)):
        with wb.Session(model, C, seed=31, max_trajectory_doublings=8) as s:
            s.init(positions=pos, mass=mass, steps=steps)
            s.reserve(60, trace=True)
            s.warmup(30, store=True).freeze().sample(30).sync()
            st = s.state()
            out.append((s.draws(0, 60), s.trace(0, 60)["lp"], st["step"], st["inv_mass"],
                        st["grad_evals"]))
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("D", [7, 100, 1000])
def test_elementwise_source_orbits_match_the_oracle(wb, oracle, D):
    var = 10.0 ** (4 * np.arange(D) / max(D - 1, 1))
    C = 6
    rng = np.random.default_rng(5)
    th, rho = rng.normal(size=(C, D)) * np.sqrt(var), rng.normal(size=(C, D))
    im = rng.uniform(0.5, 2.0, (C, D))
    model = wb.models.device_source(ELEMENTWISE_GAUSS, D, params=1.0 / var)
    got = wb.orbit(model, th, rho, im, 0.05, 40)
    target = Target("diag_gaussian", D, prec=1.0 / var)
    for c in range(C):
        want = oracle.orbit(target, th[c], rho[c], im[c], 0.05, 40)
        for k, name in enumerate(("theta", "rho", "grad")):
            np.testing.assert_allclose(got[k][c], want[k], rtol=1e-12, atol=1e-13,
                                       err_msg=name)
        np.testing.assert_allclose(got[3][c], want[3], rtol=1e-12)   # logp
        np.testing.assert_allclose(got[4][c], want[4], rtol=1e-12)   # joint


def test_elementwise_source_trajectories_follow_the_oracle(wb, oracle):
    """same Philox streams, same decisions: draws agree with the oracle's chain to rounding
    until the first rounding-induced branch flip (logp is summed term by term here)"""
    D, C, nw, ns = 12, 8, 40, 40
    var = np.linspace(0.5, 6.0, D)
    rng = np.random.default_rng(1)
    pos, mass, steps = rng.normal(size=(C, D)), rng.uniform(0.5, 2, (C, D)), np.full(C, 0.4)
    model = wb.models.device_source(ELEMENTWISE_GAUSS, D, params=1.0 / var)
    with wb.Session(model, C, seed=11) as s:
        s.init(positions=pos, mass=mass, steps=steps)
        s.reserve(nw + ns)
        s.warmup(nw, store=True).freeze().sample(ns).sync()
        draws = s.draws(0, nw + ns)
    same = 0
    for c in range(C):
        o = oracle.run_chain(Target("diag_gaussian", D, prec=1.0 / var), default_config(), 11, c,
                             pos[c], mass[c], steps[c], nw, ns, rng_policy=1)
        ref = np.concatenate([o["warmup_draws"], o["draws"]])
        close = np.all(np.abs(draws[c] - ref) <= 1e-9 * (1 + np.abs(ref)), axis=1)
        first = int(np.argmin(close)) if not close.all() else nw + ns
        same += first
        assert first >= 10, f"chain {c} leaves the oracle's chain at iteration {first}"
    print(f"\nmean first-divergence iteration {same / C:.1f} of {nw + ns}")


def numpy_orbit(grad, th, rho, im, h, n):
    g = grad(th)
    for _ in range(n):
        rho = rho + 0.5 * h * g
        th = th + h * (im * rho)
        g = grad(th)
        rho = rho + 0.5 * h * g
    return th, rho, g


def test_a_density_that_is_not_built_in(wb):
    D, C = 37, 256
    loc = np.linspace(-3.0, 3.0, D)
    model = wb.models.device_source(LOGISTIC_DIST, D, params=loc)
    rng = np.random.default_rng(2)
    th, rho, im = rng.normal(size=(4, D)) * 2, rng.normal(size=(4, D)), rng.uniform(0.5, 2, (4, D))

    def grad(x):
        return -np.tanh((x - loc) / 2)

    got = wb.orbit(model, th, rho, im, 0.1, 25)
    t, r, g = numpy_orbit(grad, th, rho, im, 0.1, 25)
    np.testing.assert_allclose(got[0], t, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(got[1], r, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(got[2], g, rtol=1e-11, atol=1e-12)
    z = t - loc
    np.testing.assert_allclose(got[3], (-np.abs(z) - 2 * np.log1p(np.exp(-np.abs(z)))).sum(1),
                               rtol=1e-12)
    # posterior: mean loc, variance pi^2 / 3, through the one-shot call (summaries on device)
    out = wb.walnuts_device_summary(model, num_chains=C, seed=3, min_warmup_iter=150,
                                    max_warmup_iter=150, min_sampling_iter=400,
                                    max_sampling_iter=400)
    z_mean = (out["mean"] - loc) / out["mcse"]
    assert np.max(np.abs(z_mean)) < 4.5, z_mean
    rel = out["variance"] / (np.pi ** 2 / 3) - 1
    assert np.max(np.abs(rel)) < 0.06, rel
    assert np.max(out["r_hat"]) < 1.02


def test_device_source_through_the_drop_in_call_and_free_running(wb):
    D, C = 20, 32
    var = np.linspace(0.5, 6.0, D)
    model = wb.models.device_source(ELEMENTWISE_GAUSS, D, params=1.0 / var)
    fit = wb.walnuts_device(model, num_chains=C, seed=9, min_warmup_iter=100,
                            max_warmup_iter=100, min_sampling_iter=20, max_sampling_iter=300,
                            rhat_converge_tol=1.01)
    lens = np.array([len(f) for f in fit])
    assert lens.min() >= 20 and lens.max() < 300 and lens.max() > lens.min()
    pooled = np.concatenate([np.asarray(f) for f in fit])
    assert np.all(np.abs(pooled.var(0) / var - 1) < 0.35)


def test_device_source_errors(wb):
    with pytest.raises(ValueError, match="does not compile"):
        wb.Session(wb.models.device_source("this is not CUDA", 5), 4)
    with pytest.raises(ValueError, match="fp64"):
        m = wb.models.device_source(ELEMENTWISE_GAUSS, 5, params=np.ones(5))
        m.dtype = "f32"
        wb.Session(m, 4)
