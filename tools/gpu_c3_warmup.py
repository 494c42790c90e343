#!/usr/bin/env python
"""c3 shape (funnel D=100, 16384 chains): device time of 300 adaptive iterations in one
launch, and of 20 quota launches of 10 sampling iterations (select a build with WB200_LIB)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import walnuts_b200 as wb  # noqa: E402

C, D = 16384, 100
with wb.Session(wb.models.funnel(D), C, seed=1, max_trajectory_doublings=10,
                max_step_halvings=8) as s:
    s.init(init_radius=1.0)
    s.reserve(10)
    s.sync()
    c0 = s.counters()
    s.timer_start()
    s.warmup(300)
    ms = s.timer_stop_ms()
    c1 = s.counters()
    ev = c1["grad_evals"] - c0["grad_evals"]
    print(f"warm-up 300 iterations: {ms:.1f} ms, {ev / ms / 1e3:.1f} M evals/s")
    s.freeze()
    s.timer_start()
    for _ in range(20):
        s.sample(10, store=False)
    ms = s.timer_stop_ms()
    c2 = s.counters()
    ev = c2["grad_evals"] - c1["grad_evals"]
    print(f"sampling 20 x 10 iterations: {ms:.1f} ms, {ev / ms / 1e3:.1f} M evals/s")
    budget = int(ev / C / 20)
    for _ in range(3):
        s.sample_ticks(budget, store=False)
    c3 = s.counters()
    s.timer_start()
    for _ in range(20):
        s.sample_ticks(budget, store=False)
    ms = s.timer_stop_ms()
    c4 = s.counters()
    ev = c4["grad_evals"] - c3["grad_evals"]
    print(f"free-running 20 x {budget} evaluations per chain: {ms:.1f} ms, "
          f"{ev / ms / 1e3:.1f} M evals/s")
