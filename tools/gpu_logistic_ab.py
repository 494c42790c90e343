#!/usr/bin/env python
"""Batched logistic gradient at the c4 shape: save logp / grad of seeded inputs (select a
build with WB200_LIB), or compare two saved results.  usage: gpu_logistic_ab.py save OUT.npz
| compare A.npz B.npz"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

if sys.argv[1] == "save":
    import bench
    from walnuts_b200.sampler import logistic_logp_grad
    N, D, C = 100_000, 512, 8192
    X, y = bench.logistic_data(N, D)
    rng = np.random.default_rng(4)
    # a posterior-like cloud: a common centre plus small per-chain spread
    centre = rng.normal(size=D) / np.sqrt(D)
    theta = centre + 0.01 * rng.normal(size=(C, D))
    lp1, g1, _ = logistic_logp_grad(X, y, theta)
    lp2, g2, _ = logistic_logp_grad(X, y, theta)
    print("run-to-run: max |dlp|", np.max(np.abs(lp1 - lp2)), "max |dg|", np.max(np.abs(g1 - g2)))
    np.savez(sys.argv[2], lp=lp1, g=g1, theta=theta)
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    dlp = a["lp"] - b["lp"]
    dg = a["g"] - b["g"]
    gs = np.abs(b["g"]).max()
    print("logp: mean diff %.4g, sd of diff %.4g, max |diff| %.4g (|logp| ~ %.3g)"
          % (dlp.mean(), dlp.std(), np.abs(dlp).max(), np.abs(b["lp"]).mean()))
    print("grad: rms diff %.4g, max |diff| %.4g, largest component %.4g"
          % (np.sqrt((dg ** 2).mean()), np.abs(dg).max(), gs))
    # the sampler acts on logp DIFFERENCES between nearby chains' positions
    d_a = np.diff(a["lp"]); d_b = np.diff(b["lp"])
    print("differences between neighbouring chains: sd of (A - B) %.4g, sd of B %.4g"
          % ((d_a - d_b).std(), d_b.std()))
