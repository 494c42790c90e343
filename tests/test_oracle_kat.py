"""The oracle against the reference's own known-answer tests.

Every case cites the reference test it replays (tests/*.cpp under
/root/reference); values and tolerances are the reference's.
"""
import math

import numpy as np
import pytest

from oracle.binding import Target, default_config

INF = float("inf")
NAN = float("nan")


# ---- tests/util_test.cpp:102-160 ------------------------------------------
def test_log_sum_exp_known_values(oracle):
    lse = oracle.log_sum_exp
    assert lse(0.0, 0.0) == pytest.approx(math.log(2.0), abs=1e-15)
    assert lse(-3.0, -3.0) == pytest.approx(-3.0 + math.log(2.0), abs=1e-15)
    assert lse(1000.0, 1000.0) == pytest.approx(1000.0 + math.log(2.0), abs=1e-10)
    assert lse(1.0, 2.0) == pytest.approx(math.log(math.exp(1) + math.exp(2)), abs=1e-15)
    assert lse(-1.0, -2.0) == pytest.approx(math.log(math.exp(-1) + math.exp(-2)), abs=1e-15)
    assert lse(1.0, 2.0) == lse(2.0, 1.0)
    assert lse(1000.0, 0.0) == pytest.approx(1000.0, abs=1e-10)
    assert lse(-1.0, -1000.0) == pytest.approx(-1.0, abs=1e-10)
    r = lse(1e308, 1e308)
    assert not math.isinf(r) and r == pytest.approx(1e308 + math.log(2.0), abs=1e295)


def test_log_sum_exp_inf_nan(oracle):  # util_test.cpp:142-165
    lse = oracle.log_sum_exp
    assert lse(-INF, 5.0) == 5.0 and lse(5.0, -INF) == 5.0
    assert lse(-INF, -INF) == -INF
    assert lse(INF, 0.0) == INF and lse(INF, INF) == INF
    assert math.isinf(lse(INF, -INF))
    assert math.isnan(lse(NAN, 1.0)) and math.isnan(lse(1.0, NAN))


def test_logp_momentum(oracle):  # util_test.cpp:236-266
    assert oracle.logp_momentum([2.0], [1.0]) == -2.0
    assert oracle.logp_momentum(np.zeros(4), np.full(4, 2.5)) == 0.0
    assert oracle.logp_momentum([1.0, 2.0, 3.0], [1.0, 1.0, 1.0]) == -0.5 * 14.0
    assert oracle.logp_momentum([2.0, 3.0], [0.5, 2.0]) == -10.0


def _solution(step, inv_m, rho):  # util_test.cpp:385-387
    return -1.0 / 8.0 * step**4 * inv_m**3 * rho**2


def test_leapfrog_error_closed_forms(oracle):  # util_test.cpp:391-476
    sn1, sn2, sn3 = (Target("std_normal", d) for d in (1, 2, 3))
    le = oracle.leapfrog_error
    assert le(sn3, np.zeros(3), np.zeros(3), np.ones(3), 1.0) == 0.0
    assert le(sn1, [0.0], [2.5], [0.3], 0.75) == pytest.approx(_solution(0.75, 0.3, 2.5), abs=1e-12)
    assert le(sn2, np.zeros(2), np.ones(2), np.ones(2), 1.0) == pytest.approx(2 * _solution(1, 1, 1), abs=1e-12)
    assert le(sn1, [0.0], [1.0], [0.25], 1.0) == pytest.approx(_solution(1.0, 0.25, 1.0), abs=1e-12)
    assert le(sn1, [0.0], [1.0], [1.0], 0.5) == pytest.approx(_solution(1, 1, 1) / 16, abs=1e-12)
    assert le(sn1, [1.0], [1.0], [1.0], 1.0) == pytest.approx(-5.0 / 32.0, abs=1e-12)
    assert le(sn1, [1.0], [0.0], [1.0], 1.0) == pytest.approx(3.0 / 32.0, abs=1e-12)
    assert le(sn2, [1.0, -2.0], [0.5, 1.0], [1.0, 1.0], 1e-4) == pytest.approx(0.0, abs=1e-12)


def test_masses_hand_calculation(oracle):  # config_test.cpp:383-398
    t = Target("std_normal", 2)
    mass, _ = oracle.init_mass_step(t, np.array([[1.0, 2.0]]), 1, 0.5, smoothing=0.5)
    np.testing.assert_allclose(mass[0], [0.5 * 1 + 0.5, 0.5 * 2 + 0.5], atol=1e-10)


def test_adapt_step_converges_from_low_and_high(oracle):  # config_test.cpp:483-497
    t = Target("std_normal", 3)
    pos = np.zeros((1, 3))
    _, lo = oracle.init_mass_step(t, pos, 287456, 1e-4, mass_in=np.ones((1, 3)))
    _, hi = oracle.init_mass_step(t, pos, 287456, 100.0, mass_in=np.ones((1, 3)))
    assert abs(math.log2(lo[0]) - math.log2(hi[0])) <= 1.01


def test_adapt_step_scales_with_inverse_mass(oracle):  # config_test.cpp:527-537
    D = 10
    t = Target("std_normal", D)

    def geo(mass, n):
        logs = []
        for i in range(n):
            _, s = oracle.init_mass_step(t, np.zeros((1, D)), 285222 + i, 0.1,
                                         mass_in=np.full((1, D), mass))
            logs.append(math.log(s[0]))
        return math.exp(np.mean(logs))

    h_unit, h_heavy, h_light = geo(1.0, 256), geo(100.0, 256), geo(0.01, 256)
    # config_test passes the value as the MASS to masses(): heavier mass, longer step
    assert h_heavy / h_unit == pytest.approx(10, abs=1.5)
    assert h_unit / h_light == pytest.approx(10, abs=1.5)


# ---- tests/summary_test.cpp -------------------------------------------------
def _acov_chains():
    return [np.array([[1, 2], [4, 6]], float), np.array([[3, 8], [7, 1], [2, 9]], float),
            np.array([[6, 4], [1, 7], [8, 2]], float)]


def test_autocovariance_golden(oracle):  # summary_test.cpp:661-677
    expected = np.array([[9 / 4, 4], [-9 / 8, -2], [14 / 3, 38 / 3], [-3, -25 / 3],
                         [2 / 3, 2], [26 / 3, 38 / 9], [-16 / 3, -64 / 27], [1, 7 / 27]])
    np.testing.assert_allclose(oracle.autocovariance(_acov_chains()), expected, atol=1e-10)


def test_r_hat_exact(oracle):  # summary_test.cpp:825-879
    perm = [np.array([[1, 2], [3, 4], [2, 3]], float), np.array([[2, 3], [1, 2], [3, 4]], float),
            np.array([[3, 4], [2, 3], [1, 2]], float)]
    np.testing.assert_array_equal(oracle.r_hat(perm), [1.0, 1.0])
    eq = [np.array([[1, 10], [2, 8], [3, 9]], float), np.array([[4, 5], [6, 7], [5, 6]], float),
          np.array([[7, 2], [9, 4], [8, 3]], float)]
    np.testing.assert_allclose(oracle.r_hat(eq), [math.sqrt(10.0)] * 2, rtol=1e-15)
    ragged = [np.array([[1, 5], [3, 3], [2, 4]], float),
              np.array([[4, 2], [6, 4], [5, 3], [7, 5]], float)]
    np.testing.assert_allclose(oracle.r_hat(ragged),
                               [math.sqrt(1 + 147 / 32), math.sqrt(1 + 3 / 32)], rtol=1e-15)


def test_r_hat_throws(oracle):  # summary_test.cpp:780-806
    with pytest.raises(ValueError, match="at least two chains"):
        oracle.r_hat([np.arange(10.0).reshape(5, 2)])
    with pytest.raises(ValueError, match="at least 3 draws"):
        oracle.r_hat([np.ones((2, 2)), np.ones((3, 2))])


def ar1_chains():
    from tests.ar1_data import AR1_CHAINS
    return [c.copy() for c in AR1_CHAINS]


def test_ess_golden(oracle):  # summary_test.cpp:1073-1083
    ess = oracle.ess(ar1_chains())
    assert ess[0] == pytest.approx(96.256789181, abs=1e-5)
    assert ess[1] == pytest.approx(7.315045989, abs=1e-5)


def test_ess_floor(oracle):  # summary_test.cpp:1117-1133
    ess = oracle.ess([np.array([[10.0], [10.1], [9.9]]), np.array([[10.0], [9.9], [10.1]])])
    assert 0 < ess[0] <= 6.0 * math.log10(6.0) + 1e-10


def test_ess_throws(oracle):  # summary_test.cpp:1019-1035
    with pytest.raises(ValueError, match="at least 3 draws"):
        oracle.ess([np.array([[1.0, 2.0], [3.0, 4.0]])])


def test_mcse_golden(oracle):  # summary_test.cpp:1182-1192
    m = oracle.mcse(ar1_chains())
    assert m[0] == pytest.approx(0.096327220756986, abs=1e-7)
    assert m[1] == pytest.approx(0.250085871061602, abs=1e-7)


# ---- Philox: Random123 known answers ---------------------------------------
def test_philox_known_answers(oracle):
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                         [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_normals_are_standard(oracle):
    z = oracle.philox_normals(7, 3, 11, 0, 200001)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.01
    assert abs((z**4).mean() - 3) < 0.1


# ---- the device's initialisation streams (oracle Philox policy) ---------------------
def test_philox_init_positions_are_the_kind2_normals(oracle):
    """config.hpp:259-268 per chain on the stateless stream: chain c of a session with
    chain_offset o draws radius * N(0,1) from (seed, o + c, iteration 0, kind 2)."""
    pos = oracle.init_positions_philox(3, 7, 99, 2.5, chain_offset=10)
    for c in range(3):
        np.testing.assert_array_equal(pos[c], 2.5 * oracle.philox_normals(99, 10 + c, 0, 2, 7))
    with pytest.raises(ValueError, match="init_scale"):
        oracle.init_positions_philox(1, 2, 1, -1.0)


def test_philox_init_mass_and_step_follow_the_reference_rules(oracle):
    """config.hpp:360-370 (mass = (1-s)|grad| + s) and util.hpp:285-303 (doubling while the
    one-step energy error exceeds log 0.9, then sqrt(1/2) while below log 0.6) with the
    momentum of chain c from Philox kind 3; the properties of tests/config_test.cpp:483-537."""
    from oracle.binding import Target
    D = 16
    t = Target("std_normal", D)
    pos = oracle.init_positions_philox(4, D, 5, 1.0)
    mass, steps = oracle.init_mass_step_philox(t, pos, 5, 1.0, smoothing=1e-3)
    np.testing.assert_allclose(mass, (1 - 1e-3) * np.abs(pos) + 1e-3, rtol=1e-15)
    # converges from both sides to within a factor 2 (config_test.cpp:483-497)
    _, lo = oracle.init_mass_step_philox(t, pos, 5, 1e-4, mass_in=np.ones((4, D)))
    _, hi = oracle.init_mass_step_philox(t, pos, 5, 100.0, mass_in=np.ones((4, D)))
    assert np.all(lo / hi < 2.0 + 1e-12) and np.all(hi / lo < 2.0 + 1e-12)
    # the accepted step brackets the energy-error window with that chain's own momentum
    for c in range(4):
        rho = oracle.philox_normals(5, c, 0, 3, D)       # sqrt(M) = 1
        err = oracle.leapfrog_error(t, pos[c], rho, np.ones(D), lo[c])
        assert err >= np.log(0.6)


# ---- the two arithmetic policies -----------------------------------------------------
@pytest.mark.parametrize("kind,D,step,nsteps", [("std_normal", 100, 0.37, 64),
                                                ("diag_gaussian", 1000, 0.2, 50),
                                                ("funnel", 100, 0.05, 40)])
def test_fused_policy_orbits_agree_with_the_reference_policy_to_1e12(oracle, kind, D, step,
                                                                     nsteps):
    """BASELINE.json north_star: fixed-step leapfrog orbits within 1e-12 relative of the
    reference's integrator.  The fused policy (one rounding per a*b + c at the accumulate
    sites, what the kernels ship) differs from the reference policy (separate roundings,
    pinned bit for bit to the reference's headers) by rounding only."""
    rng = np.random.default_rng(D + nsteps)
    prec = 1.0 / (10.0 ** (4 * np.arange(D) / (D - 1))) if kind == "diag_gaussian" else None
    t = Target(kind, D, prec=prec)
    inv_mass = rng.uniform(0.5, 2.0, D) * (1.0 / prec if prec is not None else 1.0)
    theta = rng.normal(size=D) * (0.3 if kind == "funnel" else 1.0)
    rho = rng.normal(size=D) / np.sqrt(inv_mass)
    ref = oracle.orbit(t, theta, rho, inv_mass, step, nsteps)
    with oracle.fused_arith(True):
        fused = oracle.orbit(t, theta, rho, inv_mass, step, nsteps)
    assert any(np.any(a != b) for a, b in zip(ref[:3], fused[:3])), "policies must differ"
    for a, b in zip(ref[:3], fused[:3]):
        assert np.max(np.abs(a - b)) / np.max(np.abs(a)) <= 1e-12
    assert abs(ref[3] - fused[3]) <= 1e-12 * max(abs(ref[3]), 1.0)
    assert abs(ref[4] - fused[4]) <= 1e-12 * max(abs(ref[4]), 1.0)
    assert oracle.set_fused_arith(False) is False      # the context manager restored it


def test_fused_policy_trajectories_follow_the_reference_policy_until_a_branch_flips(oracle):
    """Identically seeded chains under the two policies agree to rounding until the first
    rounding-induced accept / U-turn flip (north_star); the prefix is reported."""
    D = 20
    t = Target("diag_gaussian", D, prec=np.linspace(0.5, 3.0, D))
    cfg = default_config()
    rng = np.random.default_rng(3)
    th0, m0 = rng.normal(size=D), rng.uniform(0.5, 2.0, D)
    ref = oracle.run_chain(t, cfg, 7, 0, th0, m0, 0.4, 40, 40, rng_policy=1)
    with oracle.fused_arith(True):
        fused = oracle.run_chain(t, cfg, 7, 0, th0, m0, 0.4, 40, 40, rng_policy=1)
    a = np.concatenate([ref["warmup_draws"], ref["draws"]])
    b = np.concatenate([fused["warmup_draws"], fused["draws"]])
    bad = np.flatnonzero(np.max(np.abs(a - b), axis=1) > 1e-9 * np.max(np.abs(a), axis=1))
    first = int(bad[0]) if bad.size else len(a)
    print(f"\nfirst divergence between the arithmetic policies: iteration {first} of {len(a)}")
    assert first >= 10
