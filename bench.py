#!/usr/bin/env python
"""Benchmark of the WALNUTS hot path (BASELINE.json metric: gradient evaluations
per second and min-ESS per second vs the CPU reference sampler).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4|c5] [--no-extra-workloads]

The line the driver reads (config.workload "c2"): BASELINE.json configs[1] -- the
1000-dimensional ill-conditioned diagonal Gaussian (condition number 1e4), 4096 chains per
GPU, fp64, max_trajectory_doublings 10, max_step_halvings 5, 300 adaptive warm-up
iterations (untimed set-up) then sampling.  One STEP = `--iters-per-step` (10) WALNUTS
transitions of every chain on the GPU.

 * value      gradient evaluations / s over exactly K steps, state resident in HBM,
              timed with CUDA events on the launching stream (max over ranks)
 * e2e        the same metric through the reference-facing C-ABI call with pinned HOST
              buffers: walnutpie_sample_device_summary -- session set-up, initial positions
              uploaded, adaptive warm-up + K steps of sampling, and the posterior summaries
              (mean, variance, R-hat, ESS, MCSE per parameter, computed on the device by the
              streaming accumulators) read back; `e2e_all_draws` is the same run through
              walnutpie_sample_device, every draw copied back to the host as the reference
              returns them (PCIe-bound: 3.3 GB per step at c2)
 * roofline   7*D*8 algorithmic bytes per gradient evaluation
              against the measured HBM copy bandwidth; the chain-resident kernel keeps the
              state on chip, so `roofline_binding` restates the kernel against what does
              bind it -- fp64 instruction issue (per-evaluation instruction counts from the
              committed ncu capture x the live evaluation rate)
 * cpu_baseline  the reference's own sampler (oracle/_ref: unmodified headers on the
              Eigen shim; falls back to the oracle port) one chain per host core
 * workloads  short runs of the other BASELINE.json configs in the same process, each with
              its own roofline, clocks and posterior / R-hat check: c3 (Neal's funnel D=100,
              16384 chains), c4 (Bayesian logistic regression N=100k, D=512, 8192 chains:
              lock-step engine + tcgen05 gradient) and c5 (65,536 logistic chains sharded
              over the ranks -- strong scaling -- with the R-hat AND ESS moments combined by
              NCCL all-reduce inside a timed summary phase)
With --impl reference the reference arm alone is timed (rank 0 only).
--workload c3|c4|c5 runs that workload alone as the main line (full length, CPU leg).
stdout carries exactly one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D = 1000
CHAINS_PER_GPU = 4096
COND = 1e4
WARMUP_ITERS = 300
MAX_DOUBLINGS = 10
MAX_HALVINGS = 5
SEED = int(os.environ.get("WB200_BENCH_SEED", "20250"))  # (env: A/B experiments)
ALG_BYTES_PER_EVAL = 7 * D * 8  # read theta, rho, grad, M^-1; write theta, rho, grad

# element-wise workloads of run_ours / the reference arm.  c2 is the default line; c3
# (BASELINE.json configs[2]) exercises the within-orbit halving ladder.
ELEMENTWISE = {
    "c2": dict(label="c2: 1000-dim ill-conditioned diagonal Gaussian (cond 1e4), "
                     "4096 chains per GPU, fp64",
               kind="diag_gaussian", D=D, chains=CHAINS_PER_GPU, init_radius=2.0,
               warmup_iters=WARMUP_ITERS, burn_iters=0, max_doublings=MAX_DOUBLINGS,
               max_halvings=MAX_HALVINGS, cpu_warm=300, cpu_samp=1000,
               kernel="walnuts_chain_kernel<DiagGaussianTarget<128,4>, ADAPT=false>"),
    # the funnel's v relaxes with a time constant of ~1200 iterations after the adaptive
    # phase (tests/test_gpu_c3.py): 6000 unstored sampling iterations precede the timed
    # steps so that the posterior check is against the true moments
    "c3": dict(label="c3: Neal's funnel D=100, 16384 chains per GPU, fp64",
               kind="funnel", D=100, chains=16384, init_radius=1.0, warmup_iters=300,
               burn_iters=6000, max_doublings=10, max_halvings=8, cpu_warm=300, cpu_samp=300,
               kernel="walnuts_chain_kernel<FunnelTarget<32,2>, ADAPT=false>"),
}

# Bayesian logistic regression (BASELINE.json configs[3], [4]): the tensor-core roofline
C4 = dict(N=100_000, D=512, chains=8192, warmup_ticks=3000, ticks_per_step=100,
          max_doublings=8, max_halvings=5)
# fp64 instruction issue ceiling of one B200: 148 SMs x 4 schedulers x 1 warp instruction
# per clock; the fp64 pipe takes one warp instruction every other clock per scheduler
SM_COUNT, SCHEDULERS_PER_SM = 148, 4


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly one JSON line: keep a private copy of fd 1 for it and point
    fd 1 at stderr, so that banners of native libraries (NCCL prints its version to stdout
    at NCCL_DEBUG=VERSION) cannot get in front of it."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        # NVML in-process where available (a query takes ~0.1 ms, so short timed regions
        # still get tens of samples); the nvidia-smi query of the profiling recipe otherwise
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40),
                    ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = get_reasons(h)
                self.rows.append([str(sm), str(mx), "0"] +
                                 ["Active" if r & b else "Not Active" for _, b in bits])
                self._stop.wait(0.005)
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                     "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def variances():
    return COND ** (np.arange(D, dtype=np.float64) / (D - 1))


def bf16_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return (float(j["bf16_tflops_sustained"]), float(j["bf16_tflops"]),
                "measured (MEASURED_PEAKS.json bf16_tflops_sustained; burst in peak_burst)")
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


_LOGISTIC_CACHE = {}


def logistic_data(N, Dm):
    if (N, Dm) not in _LOGISTIC_CACHE:
        _LOGISTIC_CACHE[(N, Dm)] = _logistic_data(N, Dm)
    return _LOGISTIC_CACHE[(N, Dm)]


def _logistic_data(N, Dm):
    """X iid N(0,1) rounded to bf16, theta* ~ N(0, I/D),
    y ~ Bernoulli(sigmoid(X theta*)); host generator seeded 20250."""
    import torch
    rng = np.random.default_rng(SEED)
    X = torch.tensor(rng.standard_normal((N, Dm), dtype=np.float32)).to(torch.bfloat16)
    X = X.to(torch.float64).numpy()
    tstar = rng.standard_normal(Dm) / np.sqrt(Dm)
    y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-X @ tstar))).astype(np.float64)
    return X, y


class Comm:
    """torch.distributed over NCCL, one rank per GPU (launched by torchrun); a no-op at N=1."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: walnuts_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.on = self.world > 1
        if self.on:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.on:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op="sum"):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.device)
        if self.on:
            ops = {"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX,
                   "min": self.dist.ReduceOp.MIN}
            self.dist.all_reduce(t, op=ops[op])
        return [float(x) for x in t.cpu()]

    def close(self):
        if self.on:
            self.dist.barrier()
            self.dist.destroy_process_group()


def instr_per_eval(kernel_key):
    """Executed warp instructions per gradient evaluation of a chain-kernel instance, from
    the committed ncu capture (profiles/r2_chain_kernel_instr_per_eval.json, written by
    tools/ncu_instr_per_eval.py).  None if the file is absent."""
    p = ROOT / "profiles" / "r2_chain_kernel_instr_per_eval.json"
    if not p.exists():
        return None
    return json.loads(p.read_text()).get(kernel_key)


def binding_roofline(kernel_key, evals_per_s_per_gpu, sm_mhz):
    """What actually bounds the chain-resident kernel: instruction issue, with the fp64 pipe
    as its narrowest part.  achieved = fp64 warp instructions / s (per-evaluation count from
    the ncu capture x live evaluation rate); peak = SMs x schedulers x clock / 2."""
    ipe = instr_per_eval(kernel_key)
    if not ipe or not sm_mhz:
        return None
    clock = sm_mhz * 1e6
    fp64_peak = SM_COUNT * SCHEDULERS_PER_SM * clock / 2.0
    issue_peak = SM_COUNT * SCHEDULERS_PER_SM * clock
    fp64 = evals_per_s_per_gpu * ipe["fp64_warp_instr_per_eval"]
    total = evals_per_s_per_gpu * ipe["warp_instr_per_eval"]
    return {"bound": "fp64-issue", "achieved": fp64 / 1e9, "peak": fp64_peak / 1e9,
            "unit": "G fp64 warp instructions/s", "frac": fp64 / fp64_peak,
            "issue_slots_frac": total / issue_peak,
            "fp64_warp_instr_per_eval": ipe["fp64_warp_instr_per_eval"],
            "warp_instr_per_eval": ipe["warp_instr_per_eval"],
            "ncu": {k: ipe.get(k) for k in ("sm__inst_executed_pipe_fp64_pct",
                                            "smsp__issue_active_pct", "registers",
                                            "resident_warps_per_sm", "source")},
            "clock_mhz": sm_mhz,
            "note": "ideal = 6 fp64 instructions per element and leapfrog step (fused "
                    "policy) = D * 6 / 32 warp instructions per evaluation; the rest is the "
                    "per-leaf work of the tree (energy and U-turn reductions, selection, "
                    "bookkeeping) and the per-transition momentum refresh"}


# ---------------------------------------------------------------------------
def bench_logistic(args, comm, strong, K, W, warmup_ticks, chains=None, cpu_leg=True):
    """Logistic regression on the lock-step tick engine + tcgen05 gradient.  strong: the
    chains (65,536 by default) are sharded over the ranks (c5); else 8192 per GPU (c4).
    Sampling streams into the summary accumulators; the cross-rank R-hat / ESS / MCSE
    combination (two NCCL all-reduces) is timed as its own phase."""
    import torch

    import walnuts_b200 as wb
    from oracle.binding import Target, default_config
    from walnuts_b200 import _ffi
    from walnuts_b200.distributed import shard, stream_summary_all_ranks

    world, rank, local_rank = comm.world, comm.rank, comm.local_rank
    cfgw = dict(C4)
    if strong:
        total = chains if chains else 65536
        chain_offset, C = shard(total, world, rank)
    else:
        C = chains if chains else cfgw["chains"]
        chain_offset, total = rank * C, C * world
    N, Dm = cfgw["N"], cfgw["D"]
    tps = cfgw["ticks_per_step"]
    X, y = logistic_data(N, Dm)
    tune = dict(max_trajectory_doublings=cfgw["max_doublings"],
                max_step_halvings=cfgw["max_halvings"])
    stage = max(16, tps // 4)   # staging block: rows a chain may complete per step
    comm.barrier()
    # e2e: everything a user of the C-ABI session pays, from host X / y to host summaries
    t_e2e = time.perf_counter()
    sess = wb.Session(wb.models.logistic(X, y), C, seed=SEED, chain_offset=chain_offset,
                      device=local_rank, **tune)
    sess.init(init_radius=0.1)
    sess.reserve(stage)
    c0 = sess.counters()
    t0 = time.perf_counter()
    # free-running adaptive warm-up: every chain adapts over the transitions that fit; an
    # iteration quota would idle the batch on its slowest chain
    sess.warmup_ticks(warmup_ticks)
    sess.freeze().sync()
    warm_s = time.perf_counter() - t0
    c1 = sess.counters()
    sess.stream_begin(32)
    # a step = `tps` lock-step ticks: one batched gradient evaluation for every chain per
    # tick; chains roll straight into their next transition (free-running, ragged draws)
    for _ in range(W):
        sess.sample_ticks(tps)
    sess.sync()
    c2 = sess.counters()
    comm.barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        sess.timer_start()
        for _ in range(K):
            sess.sample_ticks(tps)
        total_ms = sess.timer_stop_ms()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    c3 = sess.counters()
    # ---- summary phase: the path's only collective.  {sum mu, sum n mu}[D] + counts, then
    # {between, within, pooled SS, sum_k acov_k(t), t < 32}[D]: two all-reduces over NVLink,
    # then the reference's R-hat / Geyer ESS / MCSE on the combined sums -- over ALL chains
    comm.barrier()
    t0 = time.perf_counter()
    summ = stream_summary_all_ranks(sess, comm.device if comm.on else None)
    torch.cuda.synchronize()
    summary_s = time.perf_counter() - t0
    rows = sess.stream_counts()
    e2e_s = time.perf_counter() - t_e2e
    evals_local = c3["grad_evals"] - c2["grad_evals"]
    launches = c3["kernel_launches"] - c2["kernel_launches"]
    total_ms, e2e_s, summary_s = comm.reduce([total_ms, e2e_s, summary_s], "max")
    evals, e2e_evals, draws_total = comm.reduce(
        [float(evals_local), float(c3["grad_evals"]), float(rows.sum())], "sum")
    value = evals / (total_ms * 1e-3)
    active_lane_fraction = evals_local / float(K * tps * C)
    clock_summary = clocks.summary()
    sess.close()
    if rank != 0:
        return None
    sustained, burst, src = bf16_peaks()
    flops_alg = 4.0 * N * Dm
    achieved = value / world * flops_alg / 1e12   # per GPU
    e2e_steps = (warmup_ticks + (W + K) * tps) / tps
    line = {
        "metric": "grad_evals_per_sec", "value": value, "unit": "grad_evals/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
        "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None,
        "dtype": "bf16 tensor cores (hi+lo split), fp32 accumulate, fp64 state",
        "data": "synthetic",
        "config": {"workload": ("c5: Bayesian logistic regression N=100k, D=512, "
                                f"{total} chains sharded over the GPUs, "
                                if strong else
                                "c4: Bayesian logistic regression N=100k, D=512, 8192 chains "
                                "per GPU, ") +
                               "lock-step tick engine + tcgen05 batched gradient",
                   "N": N, "dims": Dm, "chains_per_gpu": C, "chains_total": total,
                   "ticks_per_step": tps,
                   "parallelism": f"chains sharded over {world} GPU(s); X replicated; the "
                                  "only collective is the summary phase: two NCCL "
                                  "all-reduces of R-hat / ESS moment sums",
                   "adaptive_warmup_ticks": warmup_ticks,
                   "max_trajectory_doublings": cfgw["max_doublings"],
                   "l2": "operands (X 102 MB, R^T 1.6 GB) exceed the 126 MB L2"},
        "min_ess": float(np.min(summ["ess"])),
        "min_ess_per_sec": float(np.min(summ["ess"])) / (total_ms * 1e-3 * (W + K) / K),
        "max_r_hat": float(np.max(summ["r_hat"])),
        "median_r_hat": float(np.median(summ["r_hat"])),
        "r_hat_quantiles_50_90_99": [float(q) for q in np.quantile(summ["r_hat"],
                                                                   [0.5, 0.9, 0.99])],
        "summary_phase": {"seconds": summary_s, "chains": total, "draws": draws_total,
                          "payload_doubles": (2 * Dm + 3) + (3 + 32) * Dm,
                          "collective": ("NCCL all-reduce x2 (SUM) + MIN" if comm.on
                                         else "none (1 GPU)"),
                          "ess_truncated_dims": int(np.sum(summ["truncated"])),
                          "what": "reference R-hat / ESS / MCSE (summary.hpp:594-769) over "
                                  "ALL chains from streamed per-chain sums"},
        "active_lane_fraction": active_lane_fraction,
        "draws_per_chain": {"min": int(rows.min()), "mean": float(rows.mean()),
                            "max": int(rows.max())},
        "wall_ms": wall_ms, "gpu_launches": int(launches),
        "warmup_phase": {"ticks": warmup_ticks, "seconds": warm_s,
                         "grad_evals_per_sec": (c1["grad_evals"] - c0["grad_evals"]) / warm_s},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained,
                     "unit": "TFLOP/s", "frac": achieved / sustained,
                     # DRAM bytes of one batched evaluation at 8192 chains (GEMM 1 1.72 GB:
                     # R^T written once; GEMM 2 1.75 GB: R^T read once), ncu --set full
                     "traffic": 3.47e9 if C == 8192 else None,
                     "traffic_source": "profiles/r1_ncu_gemm_logistic_final_c4.csv",
                     "peak_burst": burst, "peak_source": src,
                     "algorithmic_flops_per_eval": flops_alg,
                     "kernel": "gemm_kmajor_kernel<256,1> + gemm_kmajor_kernel<256,2>"},
        "e2e": {"value": e2e_evals / e2e_s, "unit": "grad_evals/s",
                "h2d_bytes_per_step": int((X.nbytes + y.nbytes) * world / e2e_steps),
                "d2h_bytes_per_step": int(5 * Dm * 8 * world / e2e_steps),
                "seconds": e2e_s, "grad_evals": e2e_evals, "steps": e2e_steps,
                "api": "C-ABI session (wb200_session_create with host X / y, init, "
                       "warmup_ticks, freeze, stream_begin, sample_ticks, stream summary): "
                       "upload, adaptive warm-up, every sampling step of this run and the "
                       "summaries read back, all inside the timed region"},
        "clocks": clock_summary,
    }
    if cpu_leg:
        # stand-alone timing of the batched gradient (the dominant kernels)
        from walnuts_b200.sampler import logistic_logp_grad
        theta = np.random.default_rng(1).normal(size=(min(C, 8192), Dm)) * 0.05
        _, _, grad_ms = logistic_logp_grad(X, y, theta, repeats=5)
        line["roofline"]["gradient_only_ms_per_batched_eval"] = grad_ms
        line["roofline"]["gradient_only_tflops"] = (
            len(theta) * flops_alg / (grad_ms * 1e-3) / 1e12)
        # CPU baseline: the reference's sampler on a bounded sample of the same data
        checker, kind = load_cpu_checker()
        cores = os.cpu_count() or 1
        target = Target("logistic", Dm, X=X, y=y)
        ccfg = default_config(min_warmup_iter=2, max_warmup_iter=2, min_sampling_iter=2,
                              max_sampling_iter=2,
                              max_trajectory_doublings=cfgw["max_doublings"],
                              max_step_halvings=cfgw["max_halvings"])
        pos = checker.init_positions(cores, Dm, SEED, 0.5)
        t0 = time.perf_counter()
        r = checker.walnuts(target, ccfg, SEED, pos, np.ones((cores, Dm)),
                            np.full(cores, 0.02))
        cpu_s = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": r["grad_evals"] / cpu_s, "unit": "grad_evals/s", "cores": cores,
            "kind": kind, "seconds": cpu_s,
            "sample": f"{cores} chains x (2 warm-up + 2 sampling) iterations, same data"}
    return line


# ---------------------------------------------------------------------------
def reference_step(checker, kind, chains, warm, samp, seed, wl=None):
    """One bounded sample of the workload on the host cores: `chains` chains of
    (warm + samp) fixed iterations through the reference's own threaded driver
    (api.hpp:33-69).  Returns (gradient evaluations, seconds, draws)."""
    from oracle.binding import Target, default_config
    wl = wl or ELEMENTWISE["c2"]
    Dw = wl["D"]
    target = (Target("diag_gaussian", Dw, prec=1.0 / variances())
              if wl["kind"] == "diag_gaussian" else Target(wl["kind"], Dw))
    cfg = default_config(min_warmup_iter=warm, max_warmup_iter=warm, min_sampling_iter=samp,
                         max_sampling_iter=samp,
                         max_trajectory_doublings=wl["max_doublings"],
                         max_step_halvings=wl["max_halvings"])
    pos = checker.init_positions(chains, Dw, seed, wl["init_radius"])
    mass, steps = checker.init_mass_step(target, pos, seed, 1.0)
    t0 = time.perf_counter()
    r = checker.walnuts(target, cfg, seed, pos, mass, steps)
    dt = time.perf_counter() - t0
    return r["grad_evals"], dt, r["out"][:, :samp]


def load_cpu_checker():
    from oracle.binding import load_oracle, load_ref
    ref = load_ref()
    if ref is not None:
        return ref, "reference"
    return load_oracle(), "port"


CPU_ARM_NOTE = ("the reference's headers are compiled unmodified against a local stand-in for "
                "Eigen (oracle/eigen_shim: scalar left-to-right loops, not Eigen's packet "
                "code), so a real-Eigen build may be somewhat faster than this baseline")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    checker, kind = load_cpu_checker()
    cores = os.cpu_count() or 1
    wl = ELEMENTWISE[args.workload if args.workload in ELEMENTWISE else "c2"]
    warm, samp = 100, 50 * args.iters_per_step
    for i in range(args.warmup):
        reference_step(checker, kind, cores, warm, samp, SEED + i, wl)
    evals, secs = 0, 0.0
    for i in range(args.steps):
        e, dt, _ = reference_step(checker, kind, cores, warm, samp, SEED + 100 + i, wl)
        evals += e
        secs += dt
    value = evals / secs
    sample = (f"{cores} chains (one per core) x ({warm} warm-up + {samp} sampling) fixed "
              f"iterations per step, {wl['label']}")
    line = {
        "impl": "reference", "metric": "grad_evals_per_sec", "value": value,
        "unit": "grad_evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["label"] + " -- reference: CPU threads, one chain per core",
                   "dims": wl["D"], "chains": cores,
                   "max_trajectory_doublings": wl["max_doublings"],
                   "max_step_halvings": wl["max_halvings"]},
        "cpu_baseline": {"value": value, "unit": "grad_evals/s", "cores": cores,
                         "kind": kind, "sample": sample, "note": CPU_ARM_NOTE},
        "e2e": {"value": value, "unit": "grad_evals/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------
def sess_ld(D):
    """an upper bound of the padded row length (doubles) of a D-dimensional session"""
    ld = 64
    while ld < D:
        ld *= 2
    return ld


def one_shot(model, wl, C, rank, n_warm, n_samp, summaries, inits):
    """The C-ABI one-shot call with pinned host buffers on this rank's GPU; returns
    (seconds, gradient evaluations, h2d bytes, d2h bytes, result)."""
    import ctypes

    from walnuts_b200 import _ffi
    Dw = wl["D"]
    desc = model.desc()
    lengths = np.zeros(2 * C, np.int32)
    stepsize = np.zeros(C)
    tail = (wl["max_doublings"], wl["max_halvings"], 1, 0.5, 0.1, 1.0, 1.01, 4.0, 1e-5, 15.0,
            1.0, 0.8, 0.05, 0.8, 0.9, 1e-4, 0.5)
    head = (ctypes.byref(desc), Dw, inits, C, SEED, 1 + rank * C, wl["init_radius"], None,
            n_warm, n_warm, n_samp, n_samp)
    if summaries:
        out = {k: _ffi.pinned_empty((Dw,)) for k in ("mean", "variance", "r_hat", "ess", "mcse")}
        cut = np.zeros(Dw, np.int32)
        t0 = time.perf_counter()
        _ffi._ffi_sample_device_summary(
            *head, *tail, 32, out["mean"], out["variance"], out["r_hat"], out["ess"],
            out["mcse"], cut, lengths, stepsize, None, 0, _ffi.print_callback)
        dt = time.perf_counter() - t0
        d2h = 5 * Dw * 8 + cut.nbytes + stepsize.nbytes + lengths.nbytes
        result = {k: np.array(v) for k, v in out.items()}
        result["truncated"] = cut
    else:
        out = _ffi.pinned_empty((C, n_samp, Dw))
        t0 = time.perf_counter()
        _ffi._ffi_sample_device(*head, *tail, False, out, out.size, lengths, stepsize, None, 0,
                                _ffi.print_callback)
        dt = time.perf_counter() - t0
        d2h = out.nbytes + stepsize.nbytes + lengths.nbytes
        result = None
        del out
    return dt, float(_ffi.last_run_stats()["grad_evals"]), inits.nbytes, d2h, result


def bench_elementwise(args, comm, name, K, W, ips, cpu_leg=True, all_draws_leg=True,
                      dtype="f64"):
    """c2 / c3 on the chain-resident kernel: device-timed sampling steps, the one-shot
    C-ABI calls, the posterior check and (cpu_leg) the reference on the host cores."""
    import torch

    import walnuts_b200 as wb
    from walnuts_b200 import _ffi
    from walnuts_b200.distributed import stream_summary_all_ranks

    world, rank, local_rank = comm.world, comm.rank, comm.local_rank
    wl = ELEMENTWISE[name]
    Dw, n_warm, burn = wl["D"], wl["warmup_iters"], wl["burn_iters"]
    width = 4 if dtype == "f32" else 8
    alg_bytes = 7 * Dw * width
    C = args.chains if (args.chains and name == args.workload) else wl["chains"]
    model = (wb.models.diag_gaussian(variances(), dtype=dtype)
             if wl["kind"] == "diag_gaussian" else wb.models.funnel(Dw, dtype=dtype))
    tune = dict(max_trajectory_doublings=wl["max_doublings"],
                max_step_halvings=wl["max_halvings"])
    sess = wb.Session(model, C, seed=SEED, chain_offset=rank * C, device=local_rank, **tune)
    sess.init(init_radius=wl["init_radius"])
    # rows of the quota steps, then room for the free-running steps (ragged: chains with
    # short orbits complete several times the average number of transitions)
    free_factor = 8 if wl["kind"] == "funnel" else 2
    Kf = min(K, 20)                       # timed free-running steps (bounded for large K)
    row_bytes = C * sess_ld(Dw) * 8       # one draw row of every chain
    free_rows = (W + Kf) * ips * free_factor
    free_mem = torch.cuda.mem_get_info(local_rank)[0]
    if ((W + K) * ips + free_rows) * row_bytes > 0.6 * free_mem:
        raise SystemExit(f"--steps {K}: the stored draws of the timed steps do not fit the GPU")
    sess.reserve((W + K) * ips + free_rows)
    # ---- untimed set-up: adaptive warm-up (device time reported separately)
    sess.sync()
    c0 = sess.counters()
    sess.timer_start()
    sess.warmup(n_warm)
    warm_ms = sess.timer_stop_ms()
    sess.freeze()
    c1 = sess.counters()
    warm_evals = c1["grad_evals"] - c0["grad_evals"]
    if burn:
        for _ in range(0, burn, 50):
            sess.sample(50, store=False)
    for _ in range(W):
        sess.sample(ips)
    sess.sync()
    # ---- timed region: exactly K steps
    c2 = sess.counters()
    comm.barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        sess.timer_start()
        for _ in range(K):
            sess.sample(ips)
        total_ms = sess.timer_stop_ms()
        torch.cuda.synchronize()
        wall_ms = 1e3 * (time.perf_counter() - t0)
    comm.barrier()
    c3 = sess.counters()
    evals = c3["grad_evals"] - c2["grad_evals"]
    launches = c3["kernel_launches"] - c2["kernel_launches"]
    (max_ms,) = comm.reduce([total_ms], "max")
    (total_evals,) = comm.reduce([float(evals)], "sum")
    value = total_evals / (max_ms * 1e-3)
    clock_summary = clocks.summary()

    # ---- posterior summary of the timed draws over ALL ranks' chains: the stored draws of
    # the timed steps are folded into the streaming accumulators, then two all-reduces
    first = W * ips
    local = sess.summary(first, K * ips)
    draws_last = sess.draws(first + K * ips - 1, 1)[:, 0]       # one draw per chain

    # ---- free-running steps: the same mean work per chain and step as above, handed out
    # as a budget of gradient evaluations (wb200_session_sample_ticks) instead of an
    # iteration count, so no launch waits for the chain with the longest orbits; every
    # chain stores its draws in its own rows (ragged counts per chain)
    budget = max(1, int(round(evals / (C * K))))
    for _ in range(W):
        sess.sample_ticks(budget)
    sess.sync()
    f0 = sess.counters()
    comm.barrier()
    sess.timer_start()
    for _ in range(Kf):
        sess.sample_ticks(budget)
    free_ms = sess.timer_stop_ms()
    torch.cuda.synchronize()
    comm.barrier()
    f1 = sess.counters()
    free_counts = sess.chain_rows() - (W + K) * ips
    free_summary = sess.summary_ragged((W + K) * ips)
    (free_max_ms,) = comm.reduce([free_ms], "max")
    (free_evals,) = comm.reduce([float(f1["grad_evals"] - f0["grad_evals"])], "sum")
    (free_min_ess,) = comm.reduce([float(np.min(free_summary["ess"]))], "sum")
    free_running = {
        "value": free_evals / (free_max_ms * 1e-3), "unit": "grad_evals/s",
        "ms_per_step": free_max_ms / Kf, "steps": Kf,
        "eval_budget_per_chain_and_step": budget,
        "gpu_launches": int(f1["kernel_launches"] - f0["kernel_launches"]),
        "draws_per_chain_min_max": [int(free_counts.min()), int(free_counts.max())],
        "min_ess": free_min_ess, "max_r_hat": float(np.max(free_summary["r_hat"])),
        "roofline_frac": None,
        "what": "K steps of wb200_session_sample_ticks: every chain gets the same number of "
                "gradient evaluations per launch and completes the transitions that fit "
                "(ragged draw counts, like the reference's per-thread chains, "
                "sampler.hpp:79-94); summaries over every chain's own rows, warm-up "
                "steps included"}
    sess.close()
    comm.barrier()
    # cross-rank moments of the timed draws (chains independent: pooled over ranks)
    mom = comm.reduce(list(local["mean"]) + list(local["variance"]), "sum")
    post_mean = np.array(mom[:Dw]) / world
    post_var = np.array(mom[Dw:]) / world
    if wl["kind"] == "diag_gaussian":
        true_var = variances()
        posterior = {
            "max_abs_mean_over_sd": float(np.max(np.abs(post_mean) / np.sqrt(true_var))),
            "max_rel_var_error": float(np.max(np.abs(post_var / true_var - 1.0))),
            "max_abs_z_mean": float(np.max(np.abs(local["mean"]) / local["mcse"])),
            "truth": "N(0, diag(variances)); z of the pooled mean against the device MCSE"}
    else:
        v = draws_last[:, 0]
        n = len(v)
        posterior = {
            "E_v": float(v.mean()), "Var_v": float(v.var(ddof=1)),
            "z_E_v": float(v.mean() / np.sqrt(9.0 / n)),
            "z_Var_v": float((v.var(ddof=1) - 9.0) / (9.0 * np.sqrt(2.0 / n))),
            "max_z_E_x": float(np.max(np.abs(draws_last[:, 1:].mean(0)) /
                                      (draws_last[:, 1:].std(0, ddof=1) / np.sqrt(n)))),
            "truth": "v ~ N(0, 9), E x_i = 0: cross-chain z scores of the last timed draw "
                     f"of this rank's {n} chains, after {n_warm} adaptive + {burn} + "
                     f"{(W + K) * ips} sampling iterations",
            "tau_v_iterations_lower_bound": float(C * K * ips / local["ess"][0])}
    min_ess = comm.reduce([float(np.min(local["ess"]))], "sum")[0]  # independent shards add

    # ---- e2e: the C-ABI one-shot calls with pinned host buffers, every rank on its GPU
    torch.cuda.synchronize()
    e2e_samp = K * ips
    inits = _ffi.pinned_empty((C, Dw))
    inits[...] = np.random.default_rng(SEED).normal(size=(C, Dw)) * wl["init_radius"]
    # one short untimed call first (W warm-up steps of the e2e path): the GPU has idled
    # through the host-side summaries above and its clocks ramp up over the first ~50 ms
    one_shot(model, wl, C, rank, 40, 2 * ips, True, inits)
    comm.barrier()
    dt, ev, h2d, d2h, e2e_summary = one_shot(model, wl, C, rank, n_warm, e2e_samp, True, inits)
    (e2e_s,) = comm.reduce([dt], "max")
    (e2e_evals,) = comm.reduce([ev], "sum")
    e2e_steps = (n_warm + e2e_samp) / ips
    e2e = {"value": e2e_evals / e2e_s, "unit": "grad_evals/s",
           "h2d_bytes_per_step": int(h2d * world / e2e_steps),
           "d2h_bytes_per_step": int(d2h * world / e2e_steps),
           "seconds": e2e_s, "grad_evals": e2e_evals, "steps": e2e_steps,
           "api": "walnutpie_sample_device_summary (C-ABI, pinned host buffers): session "
                  "set-up, initial positions uploaded, adaptive warm-up, sampling and the "
                  "posterior summaries (mean, variance, R-hat, ESS, MCSE per parameter, "
                  "streamed on the device) read back, all inside the timed call",
           "min_ess": float(np.min(e2e_summary["ess"])),
           "max_r_hat": float(np.max(e2e_summary["r_hat"]))}
    e2e_all = None
    import psutil
    if all_draws_leg and C * e2e_samp * Dw * 8 * (world if world <= 8 else 8) > \
            0.5 * psutil.virtual_memory().available:
        all_draws_leg = False   # (every rank of the node pins its own buffer)
    if all_draws_leg:
        comm.barrier()
        dt, ev, h2d, d2h, _ = one_shot(model, wl, C, rank, n_warm, e2e_samp, False, inits)
        (all_s,) = comm.reduce([dt], "max")
        (all_evals,) = comm.reduce([ev], "sum")
        e2e_all = {"value": all_evals / all_s, "unit": "grad_evals/s",
                   "h2d_bytes_per_step": int(h2d * world / e2e_steps),
                   "d2h_bytes_per_step": int(d2h * world / e2e_steps), "seconds": all_s,
                   "api": "walnutpie_sample_device: the same run with every draw copied "
                          "back to the host (read-back overlapped with sampling), as the "
                          "reference returns them"}
    del inits
    if rank != 0:
        return None

    hbm_peak, peak_src = measured_peaks()
    free_running["roofline_frac"] = (free_running["value"] / world) * alg_bytes / 1e9 / hbm_peak
    achieved = (total_evals / world) * alg_bytes / (max_ms * 1e-3) / 1e9
    line = {
        "metric": "grad_evals_per_sec", "value": value, "unit": "grad_evals/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": max_ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if dtype == "f64" else
        "f32 integrator state and element-wise arithmetic, f64 energies / decisions / "
        "adaptation / stored draws",
        "data": "synthetic",
        "config": {
            "workload": wl["label"] if dtype == "f64" else wl["label"].replace("fp64", "fp32 mode"),
            "dims": Dw, "chains_per_gpu": C, "chains_total": C * world,
            "iters_per_step": ips, "adaptive_warmup_iters": n_warm,
            "unstored_sampling_iters_before_timing": burn,
            "max_trajectory_doublings": wl["max_doublings"],
            "max_step_halvings": wl["max_halvings"],
            "arithmetic": "fused policy (FMA at the accumulate sites; DESIGN.md section 3.1)",
            "parallelism": f"chains sharded over {world} GPU(s), no data-path collective",
            "l2": "working set per step (chain state + scratch + stored draws, "
                  f"{(C * Dw * 8 * (2 + ips)) / 1e6:.0f} MB) exceeds the 126 MB L2",
        },
        "min_ess_per_sec": min_ess / (max_ms * 1e-3),
        "min_ess": min_ess,
        "grad_evals_per_transition": total_evals / (C * world * K * ips),
        "wall_ms": wall_ms,
        "posterior_check": posterior,
        "warmup_phase": {"iters": n_warm, "ms": warm_ms,
                         "grad_evals_per_sec": warm_evals / (warm_ms * 1e-3)},
        "e2e": e2e,
        "free_running": free_running,
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of one timed launch (10
            # transitions x 4096 chains) from the committed ncu --set full capture
            "traffic": (instr_per_eval("c2_sampling") or {}).get("dram_bytes_per_launch")
            if name == "c2" and ips == 10 and C == 4096 else None,
            "traffic_source": "profiles/r2_ncu_chain_sampling_c2.csv",
            "peak_source": peak_src,
            "kernel": wl["kernel"],
            "algorithmic_bytes_per_eval": alg_bytes,
            "ms_per_launch": max_ms / max(launches, 1),
            "streaming_equivalent": True,
            "note": "SURVEY.md section 8(d)'s streaming model: 7*D*w bytes per evaluation. The "
                    "chain-resident kernel keeps theta/rho/grad/M^-1 on chip across "
                    "micro-steps, so its real DRAM traffic (`traffic`) is a few percent of "
                    "that and frac > 1 only says it is faster than a lock-step HBM-streaming "
                    "kernel could be; what binds it is in roofline_binding",
        },
        "roofline_binding": binding_roofline(f"{name}_sampling", total_evals / world /
                                             (max_ms * 1e-3), clock_summary.get("sm_mhz"))
        if dtype == "f64" else None,
        "clocks": clock_summary,
    }
    if e2e_all:
        line["e2e_all_draws"] = e2e_all
    if cpu_leg:
        checker, kind = load_cpu_checker()
        cores = os.cpu_count() or 1
        cpu_warm, cpu_samp = wl["cpu_warm"], wl["cpu_samp"]
        cpu_evals, cpu_s, cpu_draws = reference_step(checker, kind, cores, cpu_warm, cpu_samp,
                                                     SEED, wl)
        oracle_checker = checker if kind == "port" else __import__(
            "oracle.binding", fromlist=["load_oracle"]).load_oracle()
        cpu_min_ess = float(np.min(oracle_checker.ess([cpu_draws[c] for c in range(cores)])))
        line["cpu_baseline"] = {
            "value": cpu_evals / cpu_s, "unit": "grad_evals/s", "cores": cores, "kind": kind,
            "sample": f"{cores} chains (one per core) x ({cpu_warm} warm-up + {cpu_samp} "
                      f"sampling) fixed iterations, same target and limits",
            "min_ess_per_sec": cpu_min_ess / cpu_s, "seconds": cpu_s, "note": CPU_ARM_NOTE}
        if wl["kind"] == "funnel":
            v = cpu_draws[:, -50:, 0]
            line["cpu_baseline"]["E_v_last_50_draws"] = float(v.mean())
    return line


def compact(line):
    """the keys of a full line that a `workloads` entry keeps"""
    if line is None:
        return None
    keep = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling",
            "dtype", "config", "min_ess", "min_ess_per_sec", "max_r_hat", "posterior_check",
            "summary_phase", "active_lane_fraction", "grad_evals_per_transition",
            "warmup_phase", "roofline", "roofline_binding", "e2e", "free_running",
            "gpu_launches", "clocks")
    return {k: line[k] for k in keep if k in line}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=None,
                    help="chains per GPU (c2, c3, c4) or in total (c5); default per workload")
    ap.add_argument("--iters-per-step", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the CPU leg (scaling sweeps)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="c2 / c3: fp32 mode of the chain-resident kernel")
    ap.add_argument("--no-extra-workloads", action="store_true",
                    help="c2 line only: skip the c3 / c4 / c5 blocks")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
        return
    comm = Comm()
    K, W, cpu = args.steps, args.warmup, not args.no_cpu_baseline
    try:
        if args.workload in ("c4", "c5"):
            line = bench_logistic(args, comm, args.workload == "c5", K, W,
                                  C4["warmup_ticks"], chains=args.chains, cpu_leg=cpu)
        else:
            line = bench_elementwise(args, comm, args.workload, K, W, args.iters_per_step,
                                     cpu_leg=cpu, dtype=args.dtype)
        if args.workload == "c2" and args.dtype == "f64" and not args.no_extra_workloads:
            # the other BASELINE.json configs, short, in the same process
            extra = {}
            extra["c3"] = compact(bench_elementwise(args, comm, "c3", 5, 3, 10, cpu_leg=False,
                                                    all_draws_leg=False))
            extra["c2_f32"] = compact(bench_elementwise(args, comm, "c2", 5, 3, 10,
                                                        cpu_leg=False, all_draws_leg=False,
                                                        dtype="f32"))
            extra["c3_f32"] = compact(bench_elementwise(args, comm, "c3", 5, 3, 10,
                                                        cpu_leg=False, all_draws_leg=False,
                                                        dtype="f32"))
            extra["c4"] = compact(bench_logistic(args, comm, False, 3, 3, 2000, cpu_leg=False))
            extra["c5"] = compact(bench_logistic(args, comm, True, 2, 3, 1200, cpu_leg=False))
            if line is not None:
                line["workloads"] = extra
        if line is not None:
            emit(line)
    finally:
        comm.close()


if __name__ == "__main__":
    main()
