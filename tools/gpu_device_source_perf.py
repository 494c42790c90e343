#!/usr/bin/env python
"""c2 shape (diagonal Gaussian D=1000, 4096 chains): the built-in target against the same
density given as CUDA source (model kind 5: full-Target form and element-wise form)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import walnuts_b200 as wb  # noqa: E402
from tests.test_gpu_device_source import ELEMENTWISE_GAUSS, FULL_GAUSS  # noqa: E402

D, C = 1000, 4096
var = 10.0 ** (4 * np.arange(D) / (D - 1))
models = {"built-in": wb.models.diag_gaussian(var),
          "source, full Target": wb.models.device_source(FULL_GAUSS, D, params=1 / var),
          "source, element-wise": wb.models.device_source(ELEMENTWISE_GAUSS, D, params=1 / var)}
for name, model in models.items():
    t0 = time.perf_counter()
    with wb.Session(model, C, seed=1, max_trajectory_doublings=10) as s:
        t_create = time.perf_counter() - t0
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
    print(f"{name:22s} create {t_create:5.2f} s | warm-up "
          f"{(c1['grad_evals'] - c0['grad_evals']) / wms / 1e3:6.1f} M evals/s | sampling "
          f"{(c3['grad_evals'] - c2['grad_evals']) / sms / 1e3:6.1f} M evals/s")
