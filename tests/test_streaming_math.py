"""The algebra of the streaming summaries (walnuts_b200/csrc/stream.cu) restated in NumPy
and checked against the oracle's summary.hpp restatement -- no GPU needed.  What is pinned
here is the claim the device code rests on: per-chain running sums {n, S1, P_t, first /
last T values} about the chain's first draw, folded in block by block, reproduce the
reference's R-hat / ESS / MCSE (summary.hpp:594-769) exactly, for ragged chains, and the
two-phase cross-rank combination is invariant to how chains are split over ranks."""
import numpy as np
import pytest


class ChainSums:
    def __init__(self, D, T):
        self.T, self.n = T, 0
        self.ref = None
        self.S1 = np.zeros(D)
        self.P = np.zeros((T, D))
        self.head = np.zeros((T, D))
        self.tail = np.zeros((T, D))

    def update(self, block):      # stream_update_kernel
        T = self.T
        for x in block:
            if self.n == 0:
                self.ref = x.copy()
            y = x - self.ref
            i = self.n
            self.S1 += y
            self.P[0] += y * y
            for t in range(1, min(T - 1, i) + 1):
                self.P[t] += y * self.tail[(i - t) % T]
            self.tail[i % T] = y
            if i < T:
                self.head[i] = y
            self.n += 1

    def stats(self):              # stream_chain_stats_kernel
        my = self.S1 / self.n
        return self.ref + my, (self.P[0] - self.n * my * my) / (self.n - 1)

    def acov(self):               # stream_acov_kernel, one chain
        T, n = self.T, self.n
        my = self.S1 / n
        out = np.zeros((T, len(my)))
        hsum = np.zeros_like(my)
        lsum = np.zeros_like(my)
        for t in range(min(T, n)):
            a_t, b_t = self.S1 - lsum, self.S1 - hsum
            out[t] = (self.P[t] - my * (a_t + b_t) + (n - t) * my * my) / n
            hsum = hsum + self.head[t]
            lsum = lsum + self.tail[(n - 1 - t) % T]
        return out


def phase1(chains):
    ok = [c for c in chains if c.n >= 3]
    mu = np.array([c.stats()[0] for c in ok])
    n = np.array([c.n for c in ok], float)
    return dict(sum_mu=mu.sum(0), sum_nmu=(n[:, None] * mu).sum(0), K=len(ok), N=n.sum(),
                min_len=n.min())


def phase2(chains, mbar, pm):
    ok = [c for c in chains if c.n >= 3]
    q = w = ss = ac = 0
    for c in ok:
        m, v = c.stats()
        q = q + (m - mbar) ** 2
        w = w + v
        ss = ss + (c.n - 1) * v + c.n * (m - pm) ** 2
        ac = ac + c.acov()
    return dict(q=q, w=w, ss=ss, acov=ac)


def finish(r1, r2, T):            # stream_across_kernel + stream_geyer_kernel
    K, N, min_len = r1["K"], r1["N"], int(r1["min_len"])
    W, B = r2["w"] / K, r2["q"] / (K - 1)
    pvar = r2["ss"] / (N - 1)
    D = len(W)
    ess, cut = np.zeros(D), np.zeros(D, int)
    for d in range(D):
        vp = W[d] + B[d] if K > 1 else W[d]
        rho = np.zeros(T + 4)
        acov = lambda t: r2["acov"][t, d] / K  # noqa: E731
        even, odd = 1.0, 1.0 - (W[d] - acov(1)) / vp
        rho[0], rho[1] = even, odd
        t = 1
        while t < min_len - 4 and even + odd > 0:
            if t + 2 >= T:
                cut[d] = 1
                break
            even = 1.0 - (W[d] - acov(t + 1)) / vp
            odd = 1.0 - (W[d] - acov(t + 2)) / vp
            if even + odd >= 0:
                rho[t + 1], rho[t + 2] = even, odd
            if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
                rho[t + 1] = (rho[t - 1] + rho[t]) / 2
                rho[t + 2] = rho[t + 1]
            t += 2
        if even > 0:
            rho[t + 1] = even
        tau = max(-1 + 2 * rho[:t].sum() + rho[t + 1], 1 / np.log10(N))
        ess[d] = N / tau
    return dict(r_hat=np.sqrt(1 + B / W), ess=ess, mcse=np.sqrt(pvar) / np.sqrt(ess),
                mean=r1["sum_nmu"] / N, variance=pvar, truncated=cut)


def combine(shards, T):
    """what walnuts_b200.distributed.stream_summary_all_ranks does with its all-reduces"""
    p1 = [phase1(s) for s in shards]
    r1 = dict(sum_mu=sum(p["sum_mu"] for p in p1), sum_nmu=sum(p["sum_nmu"] for p in p1),
              K=sum(p["K"] for p in p1), N=sum(p["N"] for p in p1),
              min_len=min(p["min_len"] for p in p1))
    p2 = [phase2(s, r1["sum_mu"] / r1["K"], r1["sum_nmu"] / r1["N"]) for s in shards]
    r2 = {k: sum(p[k] for p in p2) for k in ("q", "w", "ss", "acov")}
    return finish(r1, r2, T)


def ar1(rng, n, phi, D, offset):
    x = np.zeros((n, D))
    e = rng.normal(size=(n, D))
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    return x + offset


@pytest.mark.parametrize("T", [8, 16, 32])
def test_streamed_sums_reproduce_the_reference_summaries(oracle, T):
    rng = np.random.default_rng(T)
    D = 3
    lens, phis = (200, 157, 230, 180, 211), (0.3, -0.2, 0.1, 0.4, 0.0)
    data = [ar1(rng, n, p, D, offset=np.array([1e3, -5.0, 0.0])) for n, p in zip(lens, phis)]
    chains = []
    for x in data:
        c = ChainSums(D, T)
        for lo in range(0, len(x), 37):          # blocks of uneven size
            c.update(x[lo:lo + 37])
        chains.append(c)
    whole = combine([chains], T)
    # the noise of the autocorrelation estimates keeps some Geyer sequences positive past
    # a short lag window: those dimensions are flagged, the others are exact
    ok = whole["truncated"] == 0
    assert ok.all() if T == 32 else ok.any()
    np.testing.assert_allclose(whole["ess"][ok], oracle.ess(data)[ok], rtol=1e-9)
    np.testing.assert_allclose(whole["r_hat"], oracle.r_hat(data), rtol=1e-10)
    np.testing.assert_allclose(whole["mcse"][ok], oracle.mcse(data)[ok], rtol=1e-9)
    np.testing.assert_allclose(whole["mean"], np.concatenate(data).mean(0), rtol=1e-12)
    np.testing.assert_allclose(whole["variance"], np.concatenate(data).var(0, ddof=1),
                               rtol=1e-10)
    # sharding over ranks changes nothing
    split = combine([chains[:2], chains[2:3], chains[3:]], T)
    for k in ("ess", "r_hat", "mcse", "mean", "variance"):
        np.testing.assert_allclose(split[k], whole[k], rtol=1e-12)


def test_slowly_mixing_chains_are_flagged_when_the_lag_window_is_too_short(oracle):
    rng = np.random.default_rng(1)
    data = [ar1(rng, 400, 0.95, 2, 0.0) for _ in range(4)]
    chains = []
    for x in data:
        c = ChainSums(2, 8)
        c.update(x)
        chains.append(c)
    out = combine([chains], 8)
    assert out["truncated"].all()
    assert np.all(out["ess"] >= oracle.ess(data))     # a cut sequence over-estimates ESS


def test_short_chains_are_left_out_like_the_reference_rejects_them(oracle):
    rng = np.random.default_rng(2)
    data = [ar1(rng, n, 0.2, 2, 0.0) for n in (50, 2, 60, 1)]
    chains = []
    for x in data:
        c = ChainSums(2, 16)
        c.update(x)
        chains.append(c)
    out = combine([chains], 16)
    kept = [data[0], data[2]]
    np.testing.assert_allclose(out["ess"], oracle.ess(kept), rtol=1e-9)
    np.testing.assert_allclose(out["r_hat"], oracle.r_hat(kept), rtol=1e-10)
