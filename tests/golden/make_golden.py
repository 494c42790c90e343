"""Generates tests/golden/ref_chains.npz from the UNMODIFIED reference headers
(oracle/_ref/libwalnuts_ref.so = /root/reference/include compiled against
oracle/eigen_shim).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures pin the oracle restatement on boxes where the reference is absent.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle.binding import Target, default_config, load_ref  # noqa: E402

CASES = [
    # name, kind, D, extra, cfg overrides, seed, chain, step0, n_warmup, n_sampling
    ("std_normal_d7", "std_normal", 7, {}, dict(), 99, 3, 0.3, 120, 150),
    ("ill_gauss_d10", "diag_gaussian", 10, dict(prec=1 / np.logspace(0, 4, 10)),
     dict(max_trajectory_doublings=8), 17, 0, 0.7, 150, 150),
    ("funnel_d11", "funnel", 11, {},
     dict(max_step_halvings=8, max_trajectory_doublings=7), 5, 1, 0.5, 150, 200),
    ("std_normal_minmicro", "std_normal", 5, {},
     dict(min_micro_steps=2, max_macro_steps_target=3.0), 7, 2, 1.5, 100, 100),
]


def main():
    ref = load_ref()
    assert ref is not None, "needs /root/reference"
    out = {}
    rng = np.random.default_rng(20250)
    for name, kind, D, extra, over, seed, chain, step0, nw, ns in CASES:
        t = Target(kind, D, **extra)
        cfg = default_config(**over)
        th0 = rng.normal(size=D)
        m0 = np.abs(rng.normal(size=D)) + 0.1
        r = ref.run_chain(t, cfg, seed, chain, th0, m0, step0, nw, ns)
        for k in ("warmup_draws", "warmup_lp", "warmup_step", "warmup_inv_mass",
                  "draws", "lp", "inv_mass"):
            out[f"{name}/{k}"] = r[k]
        out[f"{name}/scalars"] = np.array([r["step"], r["min_micro"], r["grad_evals"]])
        out[f"{name}/theta0"] = th0
        out[f"{name}/mass0"] = m0
        s = ref.run_sampler(t, seed, chain, th0, 1 / m0, step0, 6, 6, 2, 0.5, 200)
        out[f"{name}/fixed_draws"] = s["draws"]
        out[f"{name}/fixed_lp"] = s["lp"]
        out[f"{name}/fixed_evals"] = np.array([s["grad_evals"]])
    # initialisation (walnutpy.cpp:64-80, :186-190)
    pos = ref.init_positions(4, 6, 42, 2.0)
    out["init/positions"] = pos
    t = Target("diag_gaussian", 6, prec=np.array([1, 2, 3, 4, 5, 6.0]))
    mass, steps = ref.init_mass_step(t, pos, 42, 1.0)
    out["init/mass"] = mass
    out["init/steps"] = steps
    mass2, steps2 = ref.init_mass_step(t, pos, 42, 100.2, mass_in=np.ones((4, 6)))
    out["init/steps_given_mass"] = steps2
    np.savez_compressed(Path(__file__).parent / "ref_chains.npz", **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
