"""N > 1 host logic on CPU: world_size 2, gloo backend (127.0.0.1 rendezvous).

The device sessions are replaced by a stand-in that holds each rank's share of a
fixed synthetic population of chains, so what is tested is exactly what runs
between the kernels on a multi-GPU box: the sharding arithmetic, the all-reduced
controller statistics and the stop decisions — against the same statistics
computed by the oracle / NumPy on ALL chains in one process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from walnuts_b200.distributed import (DistributedController, combine_dimension_moments,
                                      rhat_from_moments, shard)


def test_shard_partitions_contiguously():
    for total in (1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (o0, c0), (o1, _) in zip(blocks, blocks[1:]):
                assert o0 + c0 == o1
            assert blocks[-1][0] + blocks[-1][1] == total
            sizes = [c for _, c in blocks]
            assert max(sizes) - min(sizes) <= 1


def population(C=12, D=5, seed=3):
    rng = np.random.default_rng(seed)
    log_mass = rng.normal(0.0, 0.3, (C, D))
    log_step = rng.normal(-1.0, 0.05, C)
    lp_mean = rng.normal(-50.0, 0.4, C)
    lp_var = rng.uniform(20.0, 30.0, C)
    dim_mean = rng.normal(0.0, 0.1, (C, D))
    dim_var = rng.uniform(0.8, 1.2, (C, D))
    return log_mass, log_step, lp_mean, lp_var, dim_mean, dim_var


def expected_warmup_deviation(log_mass, log_step):
    """adapt.hpp:190-223 on all chains (the oracle's warmup_should_stop math)."""
    gm = np.exp(log_mass.mean(0))
    gs = np.exp(log_step.mean())
    mass_dev = max(np.sqrt((((np.exp(log_mass[c]) - gm) / gm) ** 2).sum())
                   for c in range(len(log_step)))
    step_dev = max(0.0, max((np.exp(s) - gs) / gs for s in log_step))
    return mass_dev, step_dev


class FakeSession:
    """Holds chains [offset, offset+count) of the population; mimics what
    wb200_session_warmup_sums / _deviation / _lp_moments return."""

    def __init__(self, offset, count, tighten_after=None):
        lm, ls, mu, var, *_ = population()
        sl = slice(offset, offset + count)
        self.lm, self.ls, self.mu, self.var = lm[sl], ls[sl], mu[sl], var[sl]
        self.warm = self.samp = 0
        self.frozen = False
        self.tighten_after = tighten_after

    def warmup(self, n, store):
        self.warm += n

    def sample(self, n, store):
        self.samp += n

    def freeze(self):
        self.frozen = True

    def _scale(self):
        # chains agree once `tighten_after` warm-up iterations have run
        if self.tighten_after is not None and self.warm >= self.tighten_after:
            return 0.01
        return 1.0

    def local_warmup_sums(self):
        k = self._scale()
        D = self.lm.shape[1]
        t = torch.zeros(D + 2, dtype=torch.float64)
        t[:D] = torch.as_tensor((k * self.lm).sum(0))
        t[D] = float((k * self.ls).sum())
        t[D + 1] = len(self.ls)
        return t

    def local_warmup_deviation(self, sums):
        k = self._scale()
        D = self.lm.shape[1]
        s = sums.numpy()
        gm, gs = np.exp(s[:D] / s[D + 1]), np.exp(s[D] / s[D + 1])
        md = max(np.sqrt((((np.exp(k * self.lm[c]) - gm) / gm) ** 2).sum())
                 for c in range(len(self.ls)))
        sd = max(0.0, max((np.exp(k * x) - gs) / gs for x in self.ls))
        return torch.tensor([md, sd], dtype=torch.float64)

    def local_lp_moments(self, center=None):
        mu = self.mu - (center or 0.0)
        return torch.tensor([mu.sum(), (mu ** 2).sum(), self.var.sum(),
                             len(self.mu)], dtype=torch.float64)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lm, ls, mu, var, dmean, dvar = population()
        C = len(ls)
        off, cnt = shard(C, world, rank)
        ctl = DistributedController(FakeSession(off, cnt))
        md, sd = ctl.warmup_deviation()
        emd, esd = expected_warmup_deviation(lm, ls)
        assert md == pytest.approx(emd, rel=1e-12) and sd == pytest.approx(esd, rel=1e-12)
        rhat = ctl.lp_rhat()
        expect = np.sqrt(1 + mu.var(ddof=1) / var.mean())
        assert rhat == pytest.approx(expect, rel=1e-13)
        # stop decisions: identical on every rank, at the first check that passes
        ctl2 = DistributedController(FakeSession(off, cnt, tighten_after=20))
        done = ctl2.run_warmup(min_iter=10, max_iter=60, stride=5, mass_tol=0.05,
                               step_tol=0.05)
        assert done == 20 and ctl2.s.frozen
        ctl3 = DistributedController(FakeSession(off, cnt))
        done, _ = ctl3.run_warmup(10, 30, 5, 1e-9, 1e-9), None
        assert ctl3.s.warm == 30                       # never converges -> max_iter
        n, r = DistributedController(FakeSession(off, cnt)).run_sampling(10, 40, 5, 10.0)
        assert n == 10 and r == pytest.approx(expect, rel=1e-9)
        n, _ = DistributedController(FakeSession(off, cnt)).run_sampling(10, 40, 5, 1.0000001)
        assert n == 40
        # per-dimension R-hat over all ranks' chains == single-process formula
        rh = combine_dimension_moments(dmean[off:off + cnt], dvar[off:off + cnt],
                                       torch.device("cpu"))
        np.testing.assert_allclose(rh, np.sqrt(1 + dmean.var(0, ddof=1) / dvar.mean(0)),
                                   rtol=1e-9)
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_controllers_agree_across_two_ranks():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}


def test_rhat_single_chain_is_nan():
    assert np.isnan(rhat_from_moments(torch.tensor([1.0, 1.0, 1.0, 1.0])))
