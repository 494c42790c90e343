"""Multi-GPU driver: chains shard over ranks, moments meet in one all-reduce.

One process per GPU (torchrun).  Rank r owns the contiguous block of global chain
ids [r*C, (r+1)*C) — the ids select the Philox streams, so the chains of a G-GPU
run are exactly those of a 1-GPU run of G*C chains.  There is no collective in the
data path; ranks exchange only the cross-chain summaries the reference's
controllers compute from per-chain snapshots:

* warm-up (adapt.hpp:186-224): sum_c log M_c[d], sum_c log eps_c, chain count
  -> all-reduce(SUM) -> every rank evaluates its chains' deviations from the
  global geometric means -> all-reduce(MAX) of the two maxima;
* sampling (sampler.hpp:132-151): {sum mu_c, sum mu_c^2, sum var_c, count}
  -> all-reduce(SUM) -> R-hat of lp.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard(total_chains: int, world: int, rank: int) -> Tuple[int, int]:
    """(chain_offset, count) of rank's contiguous block; sizes differ by <= 1."""
    base, rem = divmod(total_chains, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def _all_reduce(t: torch.Tensor, op) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if t.is_cuda and dist.get_backend() == "gloo":
            # CPU rendezvous over device sessions (tests on a single-GPU box)
            h = t.cpu()
            dist.all_reduce(h, op=op)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=op)
    return t


def rhat_from_moments(m: torch.Tensor) -> float:
    """R-hat = sqrt(1 + var_{M-1}(mu) / mean(var)) from all-reduced
    {sum mu, sum mu^2, sum var, M} (sampler.hpp:143-146, util.hpp:401-404)."""
    s1, s2, sv, M = (float(x) for x in m)
    if M < 2:
        return float("nan")  # single chain: NaN, never converges
    var_of_means = (s2 - s1 * s1 / M) / (M - 1.0)
    return math.sqrt(1.0 + var_of_means / (sv / M))


class DistributedController:
    """The two controller loops over any session-like object (a
    walnuts_b200.Session on a GPU; a stand-in in the gloo tests)."""

    def __init__(self, session, device: Optional[torch.device] = None):
        self.s = session
        self.device = device or torch.device("cpu")

    # -- warm-up --------------------------------------------------------------
    def warmup_deviation(self) -> Tuple[float, float]:
        sums = self.s.local_warmup_sums()            # tensor [D + 2] on self.device
        _all_reduce(sums, dist.ReduceOp.SUM)
        dev = self.s.local_warmup_deviation(sums)    # tensor [2]
        _all_reduce(dev, dist.ReduceOp.MAX)
        return float(dev[0]), float(dev[1])

    def run_warmup(self, min_iter: int, max_iter: int, stride: int, mass_tol: float,
                   step_tol: float, store: bool = False) -> int:
        done = 0
        while done < max_iter:
            n = min(stride, max_iter - done)
            self.s.warmup(n, store)
            done += n
            if min_iter <= done < max_iter:
                dm, ds = self.warmup_deviation()
                if dm <= mass_tol and ds <= step_tol:
                    break
        self.s.freeze()
        return done

    # -- sampling -------------------------------------------------------------
    def lp_rhat(self) -> float:
        # util.hpp:401-404 is a two-pass variance: first the global mean of the chain
        # means, then the squared deviations about it (no cancellation for large |lp|)
        m0 = self.s.local_lp_moments()               # tensor [4]
        _all_reduce(m0, dist.ReduceOp.SUM)
        if float(m0[3]) < 2:
            return float("nan")
        m = self.s.local_lp_moments(float(m0[0]) / float(m0[3]))
        _all_reduce(m, dist.ReduceOp.SUM)
        return rhat_from_moments(m)

    def run_sampling(self, min_iter: int, max_iter: int, stride: int, rhat_tol: float,
                     store: bool = True) -> Tuple[int, float]:
        done, rhat = 0, float("nan")
        while done < max_iter:
            n = min(stride, max_iter - done)
            self.s.sample(n, store)
            done += n
            if min_iter <= done < max_iter:
                rhat = self.lp_rhat()
                if rhat <= rhat_tol:
                    break
        return done, rhat


class SessionAdapter:
    """walnuts_b200.Session -> the tensor interface DistributedController wants;
    buffers live on the session's GPU so NCCL reduces them in place."""

    def __init__(self, session, device: torch.device):
        self.sess = session
        self.device = device
        self._sums = torch.zeros(session.num_params + 2, dtype=torch.float64, device=device)

    def warmup(self, n, store):
        self.sess.warmup(n, store)

    def sample(self, n, store):
        self.sess.sample(n, store)

    def freeze(self):
        self.sess.freeze()

    def local_warmup_sums(self):
        self.sess.warmup_sums(self._sums.data_ptr())
        return self._sums

    def local_warmup_deviation(self, sums):
        torch.cuda.synchronize(self.device)
        out = self.sess.warmup_deviation(sums.data_ptr())
        return torch.tensor(out, dtype=torch.float64, device=self.device)

    def local_lp_moments(self, center=None):
        return torch.tensor(self.sess.lp_moments(center), dtype=torch.float64,
                            device=self.device)


def combine_dimension_moments(mean_c: np.ndarray, var_c: np.ndarray,
                              device: torch.device) -> np.ndarray:
    """Per-dimension R-hat over ALL ranks' chains from local per-chain means and
    variances [C_local, D] (summary.hpp:594-619): all-reduce of
    {sum mu, sum mu^2, sum var, count} per dimension."""
    D = mean_c.shape[1]
    pack = torch.zeros(3 * D + 1, dtype=torch.float64, device=device)
    pack[:D] = torch.as_tensor(mean_c.sum(0))
    pack[D:2 * D] = torch.as_tensor((mean_c ** 2).sum(0))
    pack[2 * D:3 * D] = torch.as_tensor(var_c.sum(0))
    pack[3 * D] = mean_c.shape[0]
    _all_reduce(pack, dist.ReduceOp.SUM)
    p = pack.cpu().numpy()
    M = p[3 * D]
    between = (p[D:2 * D] - p[:D] ** 2 / M) / (M - 1.0)
    return np.sqrt(1.0 + between / (p[2 * D:3 * D] / M))


def rhat_from_dimension_moments(m) -> np.ndarray:
    """Per-dimension R-hat from the (all-reduced) payload of
    ``Session.rhat_moments``: {sum mu, sum mu^2, sum var}[D] + chain count."""
    m = np.asarray(m, dtype=np.float64)
    D = (m.size - 1) // 3
    M = m[3 * D]
    between = (m[D:2 * D] - m[:D] ** 2 / M) / (M - 1.0)
    return np.sqrt(1.0 + between / (m[2 * D:3 * D] / M))


def stream_summary_all_ranks(session, device: Optional[torch.device] = None):
    """R-hat / ESS / MCSE / mean / variance over the streamed draws of ALL ranks' chains
    (summary.hpp:594-769): the two small all-reduces of the streaming summaries
    (include/walnuts_b200.h) -- phase 1 {sum mu, sum n mu}[D] + {K, N} (SUM) and min_len
    (MIN); phase 2, centred on the global means, {between, within, pooled SS,
    sum_k acov_k(t)}[D] (SUM) -- then the Geyer loop on the combined sums.  The result is
    identical on every rank and equal to that of one session holding all the chains."""
    from .sampler import stream_finish
    device = device or torch.device("cpu")
    D = session.num_params
    r1 = torch.as_tensor(session.stream_phase1(), device=device)
    mn = r1[2 * D + 2:].clone()
    _all_reduce(r1, dist.ReduceOp.SUM)
    _all_reduce(mn, dist.ReduceOp.MIN)
    r1[2 * D + 2] = mn[0]
    r1h = r1.cpu().numpy()
    r2 = torch.as_tensor(session.stream_phase2(r1h), device=device)
    _all_reduce(r2, dist.ReduceOp.SUM)
    return stream_finish(D, session._stream_lags, r1h, r2.cpu().numpy(),
                         want_rhat=r1h[2 * D] >= 2)
