"""Phase breakdown of the one-shot C-ABI call at the c2 size (WB200_TRACE_PHASES)
next to the raw pinned D2H bandwidth of the box."""
import ctypes
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
os.environ["WB200_TRACE_PHASES"] = "1"
import torch
import bench
from walnuts_b200 import _ffi, models

C, D = 4096, 1000
samp = int(sys.argv[1]) if len(sys.argv) > 1 else 200
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 300
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    h.copy_(x, non_blocking=True); torch.cuda.synchronize()
    print(f"raw pinned D2H: {1.0737 / (time.perf_counter() - t0):.1f} GB/s")
del x, h
model = models.ill_conditioned_gaussian(D, bench.COND)
inits = _ffi.pinned_empty((C, D))
inits[...] = np.random.default_rng(1).normal(size=(C, D)) * 2.0
out = _ffi.pinned_empty((C, samp, D))
lengths = np.zeros(2 * C, np.int32)
stepsize = np.zeros(C)
desc = model.desc()
for rep in range(2):
    t0 = time.perf_counter()
    _ffi._ffi_sample_device(
        ctypes.byref(desc), D, inits, C, 7, 1, 2.0, None, warm, warm, samp, samp,
        bench.MAX_DOUBLINGS, bench.MAX_HALVINGS, 1, 0.5, 0.1, 1.0, 1.01, 4.0, 1e-5, 15.0,
        1.0, 0.8, 0.05, 0.8, 0.9, 1e-4, 0.5, False, out, out.size, lengths, stepsize, None,
        0, _ffi.print_callback)
    dt = time.perf_counter() - t0
    st = _ffi.last_run_stats()
    print(f"rep {rep}: {dt * 1e3:.1f} ms, {st['grad_evals'] / dt / 1e6:.1f} M evals/s, "
          f"out {out.nbytes / 1e9:.2f} GB")
