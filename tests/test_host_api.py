"""Host-side logic that needs no GPU: the C-ABI library loads and exports what
include/walnuts_b200.h declares, config validation mirrors the reference's
messages, and the Python surface mirrors walnutpie's."""
import ctypes
import inspect
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "walnuts_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b((?:walnutpie|wb200|walnuts_b200)_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(wb):
    from walnuts_b200 import _ffi
    lib = ctypes.CDLL(str(_ffi.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_ffi.EXPORTED_SYMBOLS) <= set(syms)


def test_reference_entry_points_are_present(wb):
    """The nine symbols python/src/walnutpie/_ffi.py:148-257 binds."""
    from walnuts_b200 import _ffi
    for s in ["walnutpie_sample_cfunc", "walnutpie_sample_bridgestan",
              "walnutpie_separator_char", "walnutpie_ess",
              "walnutpie_r_hat", "walnutpie_mcse", "walnutpie_get_error_message",
              "walnutpie_get_error_type", "walnutpie_destroy_error"]:
        assert hasattr(_ffi._lib, s)
    assert _ffi.WALNUTPY_SEP == b"\x1c"  # walnutpy.cpp:224


def test_default_tuning_matches_reference_defaults(wb):
    """config.hpp:626-640, :947-953; step_size_init as pyfunc.py:74."""
    from walnuts_b200.sampler import make_tuning
    t = make_tuning()
    expect = dict(min_warmup_iter=50, max_warmup_iter=1000, min_sampling_iter=50,
                  max_sampling_iter=1000, max_trajectory_doublings=5,
                  max_step_halvings=5, min_micro_steps=1, max_hamiltonian_error=0.5,
                  step_size_converge_tol=0.1, mass_converge_tol=1.0,
                  rhat_converge_tol=1.01, mass_init_count=4.0,
                  mass_additive_smoothing=1e-5, max_macro_steps_target=15.0,
                  step_size_init=1.0, step_accept_rate_target=0.8,
                  step_learning_rate=0.05, step_gradient_decay=0.8,
                  step_sq_gradient_decay=0.9, step_stabilization=1e-4,
                  step_learn_rate_decay=0.5, publish_stride=5)
    for k, v in expect.items():
        assert getattr(t, k) == v, k


def test_signature_matches_walnuts_pyfunc(wb):
    """python/src/walnutpie/pyfunc.py:45-83: same keywords, same defaults."""
    sig = inspect.signature(wb.walnuts_device)
    expect = dict(num_params=None, inits=None, num_chains=4, seed=None, id=1,
                  init_radius=2.0, init_inv_metric=None, save_inv_metric=False,
                  min_warmup_iter=50, max_warmup_iter=1000, min_sampling_iter=50,
                  max_sampling_iter=1000, max_trajectory_doublings=5,
                  max_step_halvings=5, min_micro_steps=1, max_hamiltonian_error=0.5,
                  step_size_converge_tol=0.1, mass_converge_tol=1.0,
                  rhat_converge_tol=1.01, mass_init_count=4.0,
                  mass_additive_smoothing=1e-5, max_macro_steps_target=15.0,
                  step_size_init=1.0, step_accept_rate_target=0.8,
                  step_learning_rate=0.05, step_gradient_decay=0.8,
                  step_sq_gradient_decay=0.9, step_stabilization=1e-4,
                  step_learn_rate_decay=0.5, save_warmup=False, refresh=0)
    params = list(sig.parameters.values())
    assert params[0].name == "logp"
    got = {p.name: p.default for p in params[1:]}
    assert got == expect
    assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in params[1:])


@pytest.mark.parametrize("kw,msg", [
    (dict(min_sampling_iter=100, max_sampling_iter=99), "min_iter must be"),       # test_pyfunc.py:67-71
    (dict(min_warmup_iter=5, max_warmup_iter=2), "min_iter cannot be greater"),    # config.hpp:650-656
    (dict(refresh=-1), "refresh must be non-negative"),                            # errors.hpp:74-81
    (dict(rhat_converge_tol=1.0), "rhat_convergence_tol must be finite and > 1"),  # config.hpp:1050
    (dict(step_accept_rate_target=1.0), "step_accept_rate_target must be in \\(0, 1\\)"),
    (dict(max_hamiltonian_error=0.0), "max_hamiltonian_error must be finite and > 0"),
    (dict(min_micro_steps=0), "min_micro_steps must be in"),
    (dict(mass_init_count=float("inf")), "mass_init_count must be finite and > 0"),
    (dict(step_size_init=-1.0), "step size must be finite and > 0"),
    # config.hpp:675-808 (validate.hpp:91-285), one per warm-up knob
    (dict(step_size_converge_tol=0.0), "step_size_converge_tol must be finite and > 0"),
    (dict(mass_converge_tol=float("nan")), "mass_converge_tol must be finite and > 0"),
    (dict(mass_additive_smoothing=-1e-3), "mass_additive_smoothing must be finite and > 0"),
    (dict(max_macro_steps_target=0.0), "max_macro_steps_target must be finite and > 0"),
    (dict(step_learning_rate=float("inf")), "step_learning_rate must be finite and > 0"),
    (dict(step_gradient_decay=1.0), "step_gradient_decay must be in \\(0, 1\\)"),
    (dict(step_sq_gradient_decay=-0.1), "step_sq_gradient_decay must be in \\(0, 1\\)"),
    (dict(step_stabilization=0.0), "step_stabilization must be finite and > 0"),
    (dict(step_learn_rate_decay=1.5), "step_learn_rate_decay must be in \\(0, 1\\)"),
    (dict(max_trajectory_doublings=0), "max_nuts_depth must be in"),
    (dict(max_step_halvings=-1), "max_step_halvings must be in"),
    (dict(min_warmup_iter=-1), "iteration counts must be non-negative"),
])
def test_config_errors_are_value_errors_with_reference_text(wb, kw, msg):
    """Validation fires before any device work, so this runs without a GPU."""
    with pytest.raises(ValueError, match=msg):
        wb.walnuts_device(wb.models.std_normal(2), **kw)


def test_python_side_argument_checks(wb):
    m = wb.models.std_normal(3)
    with pytest.raises(ValueError, match="num_chains must be at least 1"):   # util.py:24-25
        wb.walnuts_device(m, num_chains=0)
    with pytest.raises(ValueError, match="max_sampling_iter must be at least 1"):
        wb.walnuts_device(m, max_sampling_iter=0, min_sampling_iter=0)
    with pytest.raises(ValueError, match="Invalid inits size"):              # pyfunc.py:193-203
        wb.walnuts_device(m, inits=np.zeros((2, 2)))
    with pytest.raises(ValueError, match="Invalid initial metric size"):     # util.py:35-47
        wb.walnuts_device(m, init_inv_metric=np.ones(7))
    with pytest.raises(TypeError, match="device models only"):
        wb.walnuts_device(lambda x: (0.0, -x), num_params=3)


def test_host_callback_entry_point_refuses_loudly(wb):
    """No CPU fallback: the reference's callback entry point is exported for link
    compatibility only and says what to call instead."""
    from walnuts_b200 import _ffi

    @_ffi.logp_cfunc_type
    def cb(n, theta, grad, lp, data):
        return 0

    out = np.zeros(8)
    with pytest.raises(RuntimeError, match="walnutpie_sample_device"):
        _ffi._ffi_sample_cfunc(cb, None, 2, None, 1, 1, 1, 2.0, None, 1, 1, 1, 1, 5, 5, 1,
                               0.5, 0.1, 1.0, 1.01, 4.0, 1e-5, 15.0, 1.0, 0.8, 0.05, 0.8,
                               0.9, 1e-4, 0.5, False, out, out.size,
                               np.zeros(2, np.int32), None, None, 0, _ffi.print_callback)


def test_no_gpu_means_loud_failure_not_fallback(wb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        wb.Session(wb.models.std_normal(4), 8)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under walnuts_b200/ may use it."""
    uses = re.compile(r"#\s*include\s*[\"<][^\">]*oracle|^\s*(from|import)\s+oracle\b"
                      r"|liboracle|libwalnuts_ref|oracle/", re.M)
    for p in (ROOT / "walnuts_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".hpp", ".cpp", ".h"} or p.name == "Makefile":
            assert not uses.search(p.read_text()), p


def test_batch_callback_descriptor():
    """WalnutModelDesc kind 4 carries the function pointer of the batched density; the
    Python trampoline turns exceptions into a non-zero return and keeps them for the
    caller (no GPU needed to check the plumbing)."""
    import ctypes
    from walnuts_b200 import models

    seen = []

    def fn(C, D, ld, theta, grad, lp, stream):
        seen.append((C, D, ld))
        if C == 13:
            raise ValueError("unlucky")

    m = models.batch_callback(7, fn)
    d = m.desc()
    assert d.kind == 4 and d.D == 7 and d.data0
    f = ctypes.cast(d.data0, models.BATCH_LOGP_GRAD)
    assert f(3, 7, 8, None, None, None, None, None) == 0
    assert f(13, 7, 8, None, None, None, None, None) == 1
    assert seen == [(3, 7, 8), (13, 7, 8)]
    assert isinstance(m.errors[0], ValueError)
    iface = models._DevicePointer(4096, (2, 8)).__cuda_array_interface__
    assert iface["data"] == (4096, False) and iface["typestr"] == "<f8"
