"""c3 (Neal's funnel D=100, 16384 chains): how long until the posterior of v = theta[0] is
stationary?  Prints window statistics of the stored draws (device summaries)."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import walnuts_b200 as wb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=16384)
ap.add_argument("--dims", type=int, default=100)
ap.add_argument("--warmup", type=int, default=300)
ap.add_argument("--samples", type=int, default=2000)
ap.add_argument("--window", type=int, default=250)
ap.add_argument("--halvings", type=int, default=8)
ap.add_argument("--doublings", type=int, default=10)
ap.add_argument("--radius", type=float, default=1.0)
ap.add_argument("--out", default=None)
a = ap.parse_args()

rows = []
with wb.Session(wb.models.funnel(a.dims), a.chains, seed=20250, max_step_halvings=a.halvings,
                max_trajectory_doublings=a.doublings) as s:
    s.init(init_radius=a.radius)
    s.reserve(a.samples)
    t0 = time.perf_counter()
    s.warmup(a.warmup).freeze().sync()
    tw = time.perf_counter() - t0
    c0 = s.counters()["grad_evals"]
    st = s.state()
    print(f"warm-up {a.warmup} iters: {tw:.2f} s, {c0 / tw / 1e6:.1f} M evals/s; step median "
          f"{np.median(st['step']):.4f} [{st['step'].min():.4f}, {st['step'].max():.4f}], "
          f"min_micro max {st['min_micro'].max()}, inv_mass[v] median "
          f"{np.median(st['inv_mass'][:, 0]):.3f}, inv_mass[x] median "
          f"{np.median(st['inv_mass'][:, 1:]):.3f}", flush=True)
    done = 0
    while done < a.samples:
        n = min(a.window, a.samples - done)
        t0 = time.perf_counter()
        for _ in range(0, n, 10):
            s.sample(min(10, n))
        s.sync()
        dt = time.perf_counter() - t0
        c1 = s.counters()["grad_evals"]
        summ = s.summary(done, n)
        row = dict(first=done, count=n, mean_v=float(summ["mean"][0]),
                   var_v=float(summ["variance"][0]), rhat_v=float(summ["r_hat"][0]),
                   ess_v=float(summ["ess"][0]), mcse_v=float(summ["mcse"][0]),
                   max_abs_mean_x=float(np.max(np.abs(summ["mean"][1:]))),
                   mean_var_x=float(np.mean(summ["variance"][1:])),
                   evals_per_iter=(c1 - c0) / (a.chains * n), evals_per_s=(c1 - c0) / dt)
        c0 = c1
        rows.append(row)
        print(json.dumps(row), flush=True)
        done += n
    # whole second half
    half = a.samples // 2
    summ = s.summary(half, a.samples - half)
    print("second half:", json.dumps(dict(mean_v=float(summ["mean"][0]),
          var_v=float(summ["variance"][0]), ess_v=float(summ["ess"][0]),
          mcse_v=float(summ["mcse"][0]), rhat_v=float(summ["r_hat"][0]))), flush=True)
if a.out:
    Path(a.out).write_text(json.dumps(rows, indent=1))
