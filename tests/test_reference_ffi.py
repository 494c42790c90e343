"""The reference's own ctypes layer, unchanged, over libwalnuts_b200.so.

python/src/walnutpie/_ffi.py loads `libwalnutpy` from its package directory and binds nine
symbols at import (walnutpy.cpp:134,225,227,333,346,358,371,379,387).  Here a scratch
package `walnutpie` is assembled from links to the reference's Python files (read where
they lie under /root/reference -- nothing is copied into the repository) with
`libwalnutpy.so` pointing at this repository's library.  No GPU is needed: every call made
here ends in the library's error path, which is exactly what is being checked -- the
error objects, their types and the reference's `ErrorHandledCFunc` translation of them.
Skipped where /root/reference does not exist (the GPU box)."""
import importlib
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_PKG = Path(os.environ.get("WALNUTS_REFERENCE", "/root/reference")) / "python" / "src" / "walnutpie"


@pytest.fixture(scope="module")
def ref_ffi(tmp_path_factory):
    if not (REF_PKG / "_ffi.py").exists():
        pytest.skip("reference sources not present")
    lib = ROOT / "walnuts_b200" / "libwalnuts_b200.so"
    if not lib.exists():
        pytest.skip("libwalnuts_b200.so not built")
    base = tmp_path_factory.mktemp("refpkg")
    pkg = base / "walnutpie"
    pkg.mkdir()
    for name in ("_ffi.py", "summary.py", "util.py", "pyfunc.py"):
        (pkg / name).symlink_to(REF_PKG / name)
    (pkg / "__init__.py").write_text("")       # the reference's imports stan.py (BridgeStan)
    (pkg / "libwalnutpy.so").symlink_to(lib)
    sys.path.insert(0, str(base))
    try:
        for m in [m for m in sys.modules if m == "walnutpie" or m.startswith("walnutpie.")]:
            del sys.modules[m]
        yield importlib.import_module("walnutpie._ffi")
    finally:
        sys.path.remove(str(base))
        for m in [m for m in sys.modules if m == "walnutpie" or m.startswith("walnutpie.")]:
            del sys.modules[m]


def test_reference_ffi_imports_against_this_library(ref_ffi):
    """import = np.ctypeslib.load_library + getattr of all nine symbols (_ffi.py:148-257)."""
    assert Path(ref_ffi._lib._name).resolve() == (ROOT / "walnuts_b200" /
                                                  "libwalnuts_b200.so").resolve()
    assert ref_ffi.WALNUTPY_SEP == b"\x1c"
    for f in ("_ffi_sample_cfunc", "_ffi_sample_bridgestan", "_ffi_ess", "_ffi_r_hat",
              "_ffi_mcse"):
        assert callable(getattr(ref_ffi, f))


def _sampling_tail(ref_ffi, C, D, n):
    out = np.zeros((C, n, D))
    return [C, 1, 1, 2.0, None, 5, 5, n, n, 5, 5, 1, 0.5, 0.1, 1.0, 1.01, 4.0, 1e-5, 15.0,
            1.0, 0.8, 0.05, 0.8, 0.9, 1e-4, 0.5, False, out, out.size,
            np.zeros(2 * C, np.int32), None, None, 0, ref_ffi.print_callback]


def test_reference_error_translation_works_on_this_library(ref_ffi):
    """errors.hpp:10-24 through the reference's ErrorHandledCFunc (_ffi.py:161-215): the
    host-callback and BridgeStan entry points refuse loudly with a `generic` error."""
    @ref_ffi.logp_cfunc_type
    def logp(n, theta, grad, lp, data):
        return 0

    with pytest.raises(RuntimeError, match="walnutpie_sample_device"):
        ref_ffi._ffi_sample_cfunc(logp, None, 2, None, *_sampling_tail(ref_ffi, 2, 2, 5))
    with pytest.raises(RuntimeError, match="BridgeStan"):
        ref_ffi._ffi_sample_bridgestan(b"model.so", b"{}", ref_ffi.bs_print_callback_type(
            lambda m, n, bad: None), 1, None, *_sampling_tail(ref_ffi, 2, 2, 5))


def test_reference_summarizer_reaches_this_library(ref_ffi):
    """summary.py's Summarizer over _ffi_ess / _ffi_r_hat / _ffi_mcse: with a GPU the values
    are the device summaries (pinned on the GPU by tests/test_gpu_parity.py); without one
    the library's loud no-CPU-path error arrives as the reference's RuntimeError, and the
    config errors of summary.hpp:595-603 as ValueError either way."""
    summary = importlib.import_module("walnutpie.summary")
    rng = np.random.default_rng(0)
    chains = [rng.normal(size=(50, 2)), rng.normal(size=(40, 2))]
    import torch
    if torch.cuda.is_available():
        s = summary.Summarizer(chains)
        assert s.ess().shape == (2,) and np.all(s.r_hat() < 1.2) and np.all(s.mcse() > 0)
        with pytest.raises(ValueError, match="at least two chains"):
            summary.Summarizer(chains[:1]).r_hat()
    else:
        with pytest.raises(RuntimeError, match="no CUDA device"):
            summary.Summarizer(chains).ess()
