#!/usr/bin/env python
"""c2 shape (diagonal Gaussian D=1000, 4096 chains): warm-up and sampling rates of the launch
shape selected by WB200_SHAPE_1024 (default 128x4)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import walnuts_b200 as wb  # noqa: E402

D, C = 1000, 4096
var = 10.0 ** (4 * np.arange(D) / (D - 1))
model = (wb.models.std_normal(D) if os.environ.get("WB200_TEST_MODEL") == "std_normal"
         else wb.models.diag_gaussian(var))
with wb.Session(model, C, seed=1, max_trajectory_doublings=10) as s:
    s.init(init_radius=2.0)
    s.reserve(10)
    c0 = s.counters()
    s.timer_start()
    s.warmup(300)
    wms = s.timer_stop_ms()
    c1 = s.counters()
    s.freeze()
    s.sample(30, store=False)
    c2 = s.counters()
    s.timer_start()
    for _ in range(20):
        s.sample(10, store=False)
    sms = s.timer_stop_ms()
    c3 = s.counters()
    st = s.state()
print(f"shape {os.environ.get('WB200_SHAPE_1024', '128x4'):6s} warm-up "
      f"{(c1['grad_evals'] - c0['grad_evals']) / wms / 1e3:6.1f} M evals/s | sampling "
      f"{(c3['grad_evals'] - c2['grad_evals']) / sms / 1e3:6.1f} M evals/s | evals "
      f"{c3['grad_evals']} | mean step {st['step'].mean():.6f}")
