"""Time the batched logistic gradient alone at the c4 size (run under ncu for the
per-kernel split)."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
sys.path.insert(0, ".")
import bench
from walnuts_b200.sampler import logistic_logp_grad

C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
X, y = bench.logistic_data(100_000, 512)
theta = np.random.default_rng(1).normal(size=(C, 512)) * 0.05
lp, g, ms = logistic_logp_grad(X, y, theta, repeats=reps)
fl = 4.0 * 100_000 * 512 * C
print(f"C={C}: {ms:.3f} ms per batched eval -> {fl / ms / 1e9:.1f} algorithmic TFLOP/s")
