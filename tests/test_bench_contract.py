"""bench.py's contract, the part that runs without a GPU: the reference arm
(`--impl reference`) times the reference's own CPU implementation -- the unmodified headers
compiled into oracle/_ref where present, else the oracle port -- and prints one JSON line
with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    assert line["metric"] == "grad_evals_per_sec" and line["unit"] == "grad_evals/s"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"],
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["gpu_launches"] == 0


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1",
                        "--no-extra-workloads", "--no-cpu-baseline"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)
