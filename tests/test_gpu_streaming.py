"""Streaming summaries on the device (walnuts_b200/csrc/stream.cu; SURVEY.md section
8(f)-1): R-hat / ESS / MCSE / mean / variance from running sums must equal the
reference's formulas (summary.hpp:371-405,594-769) evaluated on the very same draws --
by the oracle and by the stored-draw device path -- for equal-length and ragged chains,
and the two-phase cross-rank combination must equal one session holding all chains."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def run_pair(wb, model, C, seed, n, block, lags, calls, **tune):
    """the same chains twice: draws kept / draws streamed through a staging block"""
    with wb.Session(model, C, seed=seed, **tune) as s:
        s.init(init_radius=1.5)
        s.reserve(n)
        s.warmup(60).freeze()
        for k in calls:
            s.sample(k)
        s.sync()
        kept = s.summary(0, n)
        draws = s.draws(0, n)
    with wb.Session(model, C, seed=seed, **tune) as s:
        s.init(init_radius=1.5)
        s.reserve(block)
        s.warmup(60).freeze()
        s.stream_begin(lags)
        for k in calls:
            s.sample(k)
        streamed = s.stream_summary()
        counts = s.stream_counts()
    return kept, draws, streamed, counts


@pytest.mark.parametrize("engine", ["chain", "tick"])
@pytest.mark.parametrize("D,C,lags,block,calls", [
    (16, 64, 32, 10, [10, 10, 10, 7]),        # block-sized launches and a short tail
    (100, 48, 16, 16, [5, 40, 3, 25]),        # launches larger than the staging block
    (300, 12, 32, 50, [23, 23]),              # several launches share one block
])
def test_streamed_summaries_equal_the_summaries_of_the_kept_draws(wb, oracle, monkeypatch,
                                                                  engine, D, C, lags, block,
                                                                  calls):
    if engine == "tick":
        monkeypatch.setenv("WB200_ENGINE", "tick")
    var = np.linspace(0.5, 50.0, D)
    n = sum(calls)
    kept, draws, st, counts = run_pair(wb, wb.models.diag_gaussian(var), C, 5, n, block, lags,
                                       calls)
    assert counts.tolist() == [n] * C
    ok = st["truncated"] == 0
    assert ok.mean() > 0.8
    chains = [draws[c] for c in range(C)]
    np.testing.assert_allclose(st["mean"], kept["mean"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(st["variance"], kept["variance"], rtol=1e-9)
    np.testing.assert_allclose(st["r_hat"], kept["r_hat"], rtol=1e-9)
    np.testing.assert_allclose(st["ess"][ok], kept["ess"][ok], rtol=1e-8)
    np.testing.assert_allclose(st["mcse"][ok], kept["mcse"][ok], rtol=1e-8)
    # and the reference's formulas on the same draws (oracle restatement)
    np.testing.assert_allclose(st["ess"][ok], oracle.ess(chains)[ok], rtol=1e-8)
    np.testing.assert_allclose(st["r_hat"], oracle.r_hat(chains), rtol=1e-9)
    np.testing.assert_allclose(st["mcse"][ok], oracle.mcse(chains)[ok], rtol=1e-8)


def test_streamed_summaries_of_ragged_free_running_chains(wb, oracle, monkeypatch):
    """free-running lock-step sampling: every chain streams as many draws as it completes;
    the result equals the ragged summaries of an identical run that kept its draws."""
    monkeypatch.setenv("WB200_ENGINE", "tick")
    D, C = 12, 40
    model = wb.models.diag_gaussian(np.linspace(0.5, 8.0, D))
    ticks = [150, 150, 200]
    with wb.Session(model, C, seed=9) as s:
        s.init(init_radius=1.5)
        s.reserve(400)
        s.warmup(40).freeze()
        for t in ticks:
            s.sample_ticks(t)
        s.sync()
        rows = s.chain_rows()
        kept = s.summary_ragged(0)
        draws = s.draws(0, 400)
    with wb.Session(model, C, seed=9) as s:
        s.init(init_radius=1.5)
        s.reserve(200)
        s.warmup(40).freeze()
        s.stream_begin(32)
        for t in ticks:
            s.sample_ticks(t)
        st = s.stream_summary()
        counts = s.stream_counts()
    assert rows.min() < rows.max(), "chains should be ragged"
    np.testing.assert_array_equal(counts, rows)
    ok = st["truncated"] == 0
    chains = [draws[c, :rows[c]] for c in range(C)]
    np.testing.assert_allclose(st["mean"], np.concatenate(chains).mean(0), rtol=1e-10,
                               atol=1e-12)
    np.testing.assert_allclose(st["r_hat"], oracle.r_hat(chains), rtol=1e-9)
    np.testing.assert_allclose(st["ess"][ok], oracle.ess(chains)[ok], rtol=1e-8)
    np.testing.assert_allclose(st["ess"][ok], kept["ess"][ok], rtol=1e-8)


def test_one_shot_summary_call_equals_the_one_shot_draws(wb, oracle):
    """walnutpie_sample_device_summary runs the very chains of walnutpie_sample_device and
    returns the reference summaries of their draws without shipping them."""
    D, C = 24, 32
    model = wb.models.diag_gaussian(np.linspace(0.2, 9.0, D))
    kw = dict(min_warmup_iter=80, max_warmup_iter=80, min_sampling_iter=120,
              max_sampling_iter=120)
    fit = wb.walnuts_device(model, num_chains=C, seed=77, save_inv_metric=True, **kw)
    summ = wb.walnuts_device_summary(model, num_chains=C, seed=77, max_lags=32, **kw)
    chains = [np.asarray(f) for f in fit]
    assert summ["warmup_iters"] == 80 and summ["sampling_iters"] == 120
    np.testing.assert_array_equal(summ["stepsize"], [f.warmup.stepsize for f in fit])
    np.testing.assert_array_equal(summ["inv_metric"], [f.warmup.inv_metric for f in fit])
    ok = summ["truncated"] == 0
    assert ok.mean() > 0.8
    np.testing.assert_allclose(summ["mean"], np.concatenate(chains).mean(0), rtol=1e-10,
                               atol=1e-12)
    np.testing.assert_allclose(summ["variance"], np.concatenate(chains).var(0, ddof=1),
                               rtol=1e-9)
    np.testing.assert_allclose(summ["r_hat"], oracle.r_hat(chains), rtol=1e-9)
    np.testing.assert_allclose(summ["ess"][ok], oracle.ess(chains)[ok], rtol=1e-8)
    np.testing.assert_allclose(summ["mcse"][ok], oracle.mcse(chains)[ok], rtol=1e-8)


# ---- cross-rank combination (NCCL when the box has 2 GPUs, else gloo on GPU 0) ---------
WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["WB200_ROOT"])
import walnuts_b200 as wb
from walnuts_b200.distributed import shard, stream_summary_all_ranks

backend = os.environ["WB200_BACKEND"]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
dev = rank % ndev if backend == "gloo" else rank
torch.cuda.set_device(dev)
dist.init_process_group(backend, rank=rank, world_size=world,
                        **({"device_id": torch.device("cuda", dev)} if backend == "nccl" else {}))
D, TOTAL = 20, 50
off, cnt = shard(TOTAL, world, rank)
with wb.Session(wb.models.diag_gaussian(np.linspace(0.5, 6.0, D)), cnt, seed=4,
                chain_offset=off, device=dev) as s:
    s.init(init_radius=2.0)
    s.reserve(25)
    s.warmup(60).freeze()
    s.stream_begin(32)
    for _ in range(6):
        s.sample(20)
    out = stream_summary_all_ranks(
        s, torch.device("cuda", dev) if backend == "nccl" else torch.device("cpu"))
with open(os.environ["WB200_OUT"] + f".{rank}", "w") as f:
    json.dump({k: np.asarray(v).tolist() for k, v in out.items()}, f)
dist.barrier()
dist.destroy_process_group()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_ranks(world, backend, tmp_path):
    import json
    script = tmp_path / "stream_worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WB200_ROOT=str(ROOT), WB200_BACKEND=backend,
               WB200_OUT=str(tmp_path / f"s_{backend}_{world}"), MASTER_ADDR="127.0.0.1",
               MASTER_PORT=str(_free_port()), WORLD_SIZE=str(world))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(world)]
    logs = [p.communicate(timeout=600)[0] for p in procs]
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return [json.loads(Path(env["WB200_OUT"] + f".{r}").read_text()) for r in range(world)]


@pytest.mark.timeout(900)
def test_cross_rank_streamed_summaries_equal_one_session_with_all_chains(tmp_path):
    """ESS / R-hat / MCSE combined over 2 ranks (one all-reduce of
    (3 + T) D sums after a (2 D + 3)-element one) equal those of a single session holding
    all 50 chains to 1e-8 -- the reference's formula over ALL chains, not a sum of
    per-rank ESS."""
    import torch
    one = _run_ranks(1, "gloo", tmp_path)[0]
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    two = _run_ranks(2, backend, tmp_path)
    print(f"\nbackend {backend}: min ESS {min(one['ess']):.1f}, max R-hat {max(one['r_hat']):.5f}")
    for r in two:
        for k in ("mean", "variance", "r_hat", "ess", "mcse"):
            np.testing.assert_allclose(r[k], one[k], rtol=1e-8, atol=1e-12, err_msg=k)
        assert r["truncated"] == one["truncated"]
    assert two[0] == two[1]       # identical on every rank


@pytest.mark.parametrize("blocks", ["free-running", "uniform"])
def test_multi_device_one_shot_call_equals_the_single_device_call(wb, oracle, monkeypatch,
                                                                  blocks):
    """walnutpie_sample_device_multi (one call, one host thread per GPU, controllers and
    summaries all-reduced by NCCL inside the library) against walnutpie_sample_device_summary
    on one device: same stop iterations, same per-chain step sizes and metrics, summaries
    equal to 1e-8 -- with early stopping enabled, so the all-reduced controllers decide.
    On a single-GPU box the multi-device call runs with one device (no NCCL).  Both block
    modes: free-running (default; budgets from the all-reduced totals, per-chain lengths)
    and blocks of equal iteration counts."""
    import torch
    if blocks == "uniform":
        monkeypatch.setenv("WB200_BLOCKS", "uniform")
    D, C = 16, 48
    model = wb.models.diag_gaussian(np.linspace(0.3, 5.0, D))
    kw = dict(min_warmup_iter=20, max_warmup_iter=300, min_sampling_iter=40,
              max_sampling_iter=300, mass_converge_tol=0.9, step_size_converge_tol=0.35,
              rhat_converge_tol=1.01)
    one = wb.walnuts_device_summary(model, num_chains=C, seed=5, **kw)
    ndev = min(torch.cuda.device_count(), 2)
    many = wb.walnuts_device_summary(model, num_chains=C, seed=5, devices=list(range(ndev)),
                                     **kw)
    print(f"\n{ndev} device(s): sampling stopped at {many['sampling_iters']} iterations, "
          f"min ESS {many['ess'].min():.1f}")
    assert many["sampling_iters"] == one["sampling_iters"]
    np.testing.assert_array_equal(many["sampling_lengths"], one["sampling_lengths"])
    if blocks == "uniform":
        assert len(set(one["sampling_lengths"])) == 1
    else:
        assert one["sampling_lengths"].max() > one["sampling_lengths"].min()
    np.testing.assert_array_equal(many["stepsize"], one["stepsize"])
    np.testing.assert_array_equal(many["inv_metric"], one["inv_metric"])
    for k in ("mean", "variance", "r_hat", "ess", "mcse"):
        np.testing.assert_allclose(many[k], one[k], rtol=1e-8, atol=1e-12, err_msg=k)
    with pytest.raises(ValueError, match="more devices than chains"):
        wb.walnuts_device_summary(model, num_chains=1, devices=[0, 0], **kw)


def test_multi_device_call_on_the_lock_step_engine(wb):
    """the same for a logistic model (lock-step ticks, tensor-core gradient): free-running
    blocks are budgets of ticks, the stop decisions and per-chain lengths of the
    multi-device call equal the single-device call's"""
    import torch
    from tests.test_gpu_parity import make_logistic
    X, y = make_logistic(300, 8, 11)
    model = wb.models.logistic(X, y)
    kw = dict(num_chains=48, seed=3, min_warmup_iter=30, max_warmup_iter=100,
              min_sampling_iter=30, max_sampling_iter=120, rhat_converge_tol=1.03,
              mass_converge_tol=0.9, step_size_converge_tol=0.35, max_trajectory_doublings=8)
    one = wb.walnuts_device_summary(model, **kw)
    ndev = min(torch.cuda.device_count(), 2)
    many = wb.walnuts_device_summary(model, devices=list(range(ndev)), **kw)
    np.testing.assert_array_equal(many["sampling_lengths"], one["sampling_lengths"])
    np.testing.assert_array_equal(many["stepsize"], one["stepsize"])
    for k in ("mean", "variance", "r_hat", "ess", "mcse"):
        np.testing.assert_allclose(many[k], one[k], rtol=1e-8, atol=1e-12, err_msg=k)
