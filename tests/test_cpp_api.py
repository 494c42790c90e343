"""The C++ entry point (walnuts_b200/host/api.hpp), mirror of walnutpie::walnuts
(api.hpp:33-69): compiled with g++ against the C ABI, driven like
examples/walnutpie_api.cpp."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "cpp" / "test_api.cpp"
EXE = ROOT / "tests" / "cpp" / "test_api"
LIBDIR = ROOT / "walnuts_b200"


def build():
    deps = [SRC, ROOT / "walnuts_b200" / "host" / "api.hpp",
            ROOT / "walnuts_b200" / "host" / "config.hpp", ROOT / "include" / "walnuts_b200.h"]
    if not EXE.exists() or any(d.stat().st_mtime > EXE.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-Wall", "-o", str(EXE), str(SRC),
                        f"-L{LIBDIR}", "-lwalnuts_b200", f"-Wl,-rpath,{LIBDIR}",
                        "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       check=True)
    return EXE


def test_cpp_api_host_side(wb):
    """Handler-count and configuration errors carry the reference's messages; without a
    GPU the run fails loudly (no CPU path)."""
    exe = build()
    r = subprocess.run([str(exe), "cpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cpu checks passed" in r.stdout


@pytest.mark.gpu
def test_cpp_api_matches_the_session_api(wb, tmp_path):
    """walnuts_b200::walnuts delivers, through the reference's handler interface, exactly
    the draws the session API produces for the same seed and configuration."""
    exe = build()
    out = tmp_path / "draws.bin"
    r = subprocess.run([str(exe), "gpu", str(out)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    C, D, warm, samp = 6, 5, 40, 30
    got = np.fromfile(out).reshape(C, warm + samp, D)
    with wb.Session(wb.models.std_normal(D), C, seed=1234) as s:
        s.init(init_radius=1.5, mass=np.ones((C, D)), steps=np.full(C, 0.4))
        s.reserve(warm + samp)
        s.warmup(warm, store=True).freeze().sample(samp).sync()
        np.testing.assert_array_equal(s.draws(0, warm + samp), got)
