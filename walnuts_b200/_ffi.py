"""ctypes binding of libwalnuts_b200.so (include/walnuts_b200.h).

Mirrors python/src/walnutpie/_ffi.py of the reference: same loader shape, same
``ErrorHandledCFunc`` convention (rc != 0 -> the C error object becomes
RuntimeError / ValueError / KeyboardInterrupt, _ffi.py:161-215), same argument
lists for the entry points the reference binds.  There is no fallback: if the
CUDA library is missing the import fails.
"""
from __future__ import annotations

import ctypes
import os
import functools
import sys
from pathlib import Path

import numpy as np
from numpy.ctypeslib import ndpointer

_HERE = Path(__file__).resolve().parent
# WB200_LIB: a build variant of the same library (kernel experiments, csrc/Makefile VARIANT=)
LIB_PATH = Path(os.environ["WB200_LIB"]) if os.environ.get("WB200_LIB") else \
    _HERE / "libwalnuts_b200.so"

if not LIB_PATH.exists():
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C walnuts_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "walnuts_b200 has no CPU fallback.")

_lib = ctypes.CDLL(str(LIB_PATH))


def wrapped_ndptr(*args, **kwargs):
    """ndpointer that also accepts None (passed as NULL), as _ffi.py:55-68."""
    base = ndpointer(*args, **kwargs)

    def from_param(_cls, obj):
        if obj is None:
            return obj
        return base.from_param(obj)

    return type(base.__name__, (base,), {"from_param": classmethod(from_param)})


double_array = ndpointer(dtype=ctypes.c_double, flags=("C_CONTIGUOUS"))
int_array = ndpointer(dtype=ctypes.c_int, flags=("C_CONTIGUOUS"))
nullable_double_array = wrapped_ndptr(dtype=ctypes.c_double, flags=("C_CONTIGUOUS"))
nullable_int_array = wrapped_ndptr(dtype=ctypes.c_int, flags=("C_CONTIGUOUS"))
err_ptr = ctypes.POINTER(ctypes.c_void_p)

logp_cfunc_type = ctypes.CFUNCTYPE(
    ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_double),
    ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
    ctypes.c_void_p)

print_callback_type = ctypes.CFUNCTYPE(
    None, ctypes.POINTER(ctypes.c_char), ctypes.c_size_t, ctypes.c_bool)


@print_callback_type
def print_callback(msg, size, is_error):
    print(ctypes.string_at(msg, size).decode("utf-8"),
          file=sys.stderr if is_error else sys.stdout, end="")


class WalnutModelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("D", ctypes.c_int), ("N", ctypes.c_size_t),
                ("data0", ctypes.c_void_p), ("data1", ctypes.c_void_p),
                ("precision", ctypes.c_int)]


class WalnutTuning(ctypes.Structure):
    _fields_ = [
        ("min_warmup_iter", ctypes.c_int), ("max_warmup_iter", ctypes.c_int),
        ("min_sampling_iter", ctypes.c_int), ("max_sampling_iter", ctypes.c_int),
        ("max_trajectory_doublings", ctypes.c_int),
        ("max_step_halvings", ctypes.c_int), ("min_micro_steps", ctypes.c_int),
        ("max_hamiltonian_error", ctypes.c_double),
        ("step_size_converge_tol", ctypes.c_double),
        ("mass_converge_tol", ctypes.c_double),
        ("rhat_converge_tol", ctypes.c_double),
        ("mass_init_count", ctypes.c_double),
        ("mass_additive_smoothing", ctypes.c_double),
        ("max_macro_steps_target", ctypes.c_double),
        ("step_size_init", ctypes.c_double),
        ("step_accept_rate_target", ctypes.c_double),
        ("step_learning_rate", ctypes.c_double),
        ("step_gradient_decay", ctypes.c_double),
        ("step_sq_gradient_decay", ctypes.c_double),
        ("step_stabilization", ctypes.c_double),
        ("step_learn_rate_decay", ctypes.c_double),
        ("publish_stride", ctypes.c_int),
        ("precision", ctypes.c_int),
    ]


_common_sampling_argtypes = [
    ctypes.c_size_t,  # num_chains
    ctypes.c_uint,  # seed
    ctypes.c_uint,  # id
    ctypes.c_double,  # init_radius
    nullable_double_array,  # metric init in
    ctypes.c_int, ctypes.c_int,  # min/max warmup iter
    ctypes.c_int, ctypes.c_int,  # min/max sampling iter
    ctypes.c_int,  # max_trajectory_doublings
    ctypes.c_int,  # max_step_halvings
    ctypes.c_int,  # min_micro_steps
    ctypes.c_double,  # max_hamiltonian_error
    ctypes.c_double,  # step_size_converge_tol
    ctypes.c_double,  # mass_converge_tol
    ctypes.c_double,  # rhat_converge_tol
    ctypes.c_double,  # mass_init_count
    ctypes.c_double,  # mass_additive_smoothing
    ctypes.c_double,  # max_macro_steps_target
    ctypes.c_double,  # step_size_init
    ctypes.c_double,  # step_accept_rate_target
    ctypes.c_double,  # step_learning_rate
    ctypes.c_double,  # step_gradient_decay
    ctypes.c_double,  # step_sq_gradient_decay
    ctypes.c_double,  # step_stabilization
    ctypes.c_double,  # step_learn_rate_decay
    ctypes.c_bool,  # save_warmup
    double_array,  # out
    ctypes.c_size_t,  # buffer size
    int_array,  # final lengths
    nullable_double_array,  # stepsize out
    nullable_double_array,  # metric out
    ctypes.c_int,  # refresh
    print_callback_type,
    err_ptr,
]

_common_summary_argtypes = [
    double_array, ctypes.c_int, ctypes.c_int, int_array, ctypes.c_int,
    double_array, err_ptr,
]

_get_error_msg = _lib.walnutpie_get_error_message
_get_error_msg.restype = ctypes.c_char_p
_get_error_msg.argtypes = [ctypes.c_void_p]
_get_error_type = _lib.walnutpie_get_error_type
_get_error_type.restype = ctypes.c_int
_get_error_type.argtypes = [ctypes.c_void_p]
_free_error = _lib.walnutpie_destroy_error
_free_error.restype = None
_free_error.argtypes = [ctypes.c_void_p]


class ErrorHandledCFunc:
    """Fallible C functions: int rc + trailing error pointer -> exceptions."""

    _exception_types = [RuntimeError, ValueError, KeyboardInterrupt]

    def __init__(self, f):
        f.restype = ctypes.c_int
        f.errcheck = ErrorHandledCFunc._check_rc
        functools.update_wrapper(self, f)

    def __call__(self, *args):
        ptr = ctypes.pointer(ctypes.c_void_p())
        self.__wrapped__(*args, ptr)

    def __setattr__(self, name, value):
        if name == "argtypes":
            if value[-1] != err_ptr:
                raise AttributeError("Last entry of 'argtypes' must be err_ptr")
            self.__wrapped__.__setattr__(name, value)
        super().__setattr__(name, value)

    @classmethod
    def _check_rc(cls, rc, f, args):
        if rc == 0:
            return
        ptr = args[-1]
        if ptr.contents:
            msg = _get_error_msg(ptr.contents).decode("utf-8")
            exception_type = _get_error_type(ptr.contents)
            _free_error(ptr.contents)
            CPlusPlusError = cls._exception_types[exception_type]
            raise CPlusPlusError(msg)
        raise RuntimeError(f"Unknown error, function returned code {rc}")


def erroring(f):
    return ErrorHandledCFunc(f)


# ---- the reference's entry points ------------------------------------------
_ffi_sample_device = erroring(_lib.walnutpie_sample_device)
_ffi_sample_device.argtypes = [
    ctypes.POINTER(WalnutModelDesc),  # model (replaces callback + data)
    ctypes.c_int,  # num_params
    nullable_double_array,  # inits
] + _common_sampling_argtypes

# the same call without the draw buffer: save_warmup, out, out_size -> max_lags and the
# five summary arrays + truncation flags (include/walnuts_b200.h)
_ffi_sample_device_summary = erroring(_lib.walnutpie_sample_device_summary)
_ffi_sample_device_summary.argtypes = [
    ctypes.POINTER(WalnutModelDesc), ctypes.c_int, nullable_double_array,
] + _common_sampling_argtypes[:26] + [
    ctypes.c_int,  # max_lags
    nullable_double_array, nullable_double_array, nullable_double_array,
    nullable_double_array, nullable_double_array, nullable_int_array,
] + _common_sampling_argtypes[29:]

_ffi_sample_device_multi = erroring(_lib.walnutpie_sample_device_multi)
_ffi_sample_device_multi.argtypes = [
    int_array, ctypes.c_int,  # devices
] + _ffi_sample_device_summary.__wrapped__.argtypes

_ffi_sample_cfunc = erroring(_lib.walnutpie_sample_cfunc)
_ffi_sample_cfunc.argtypes = [
    logp_cfunc_type, ctypes.c_void_p, ctypes.c_int, nullable_double_array,
] + _common_sampling_argtypes

_ffi_ess = erroring(_lib.walnutpie_ess)
_ffi_ess.argtypes = _common_summary_argtypes
_ffi_r_hat = erroring(_lib.walnutpie_r_hat)
_ffi_r_hat.argtypes = _common_summary_argtypes
_ffi_mcse = erroring(_lib.walnutpie_mcse)
_ffi_mcse.argtypes = _common_summary_argtypes

_get_separator = _lib.walnutpie_separator_char
_get_separator.restype = ctypes.c_char
_get_separator.argtypes = []
WALNUTPY_SEP = _get_separator()

_lib.walnuts_b200_default_tuning.restype = None
_lib.walnuts_b200_default_tuning.argtypes = [ctypes.POINTER(WalnutTuning)]
_lib.walnuts_b200_version.restype = ctypes.c_char_p

# ---- session API -------------------------------------------------------------
session_p = ctypes.c_void_p


def _sess(name, argtypes):
    f = erroring(getattr(_lib, name))
    f.argtypes = argtypes + [err_ptr]
    return f


session_create = _sess("wb200_session_create", [
    ctypes.POINTER(WalnutModelDesc), ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint,
    ctypes.POINTER(WalnutTuning), ctypes.c_int, ctypes.POINTER(session_p)])
_lib.wb200_session_destroy.restype = None
_lib.wb200_session_destroy.argtypes = [session_p]
session_destroy = _lib.wb200_session_destroy
session_init = _sess("wb200_session_init", [
    session_p, nullable_double_array, ctypes.c_double, nullable_double_array,
    nullable_double_array])
session_reserve = _sess("wb200_session_reserve_draws",
                        [session_p, ctypes.c_longlong, ctypes.c_int])
session_warmup = _sess("wb200_session_warmup", [session_p, ctypes.c_int, ctypes.c_int])
session_freeze = _sess("wb200_session_freeze", [session_p])
session_sample = _sess("wb200_session_sample", [session_p, ctypes.c_int, ctypes.c_int])
session_sync = _sess("wb200_session_sync", [session_p])
session_sample_ticks = _sess("wb200_session_sample_ticks",
                             [session_p, ctypes.c_int, ctypes.c_int])
session_warmup_ticks = _sess("wb200_session_warmup_ticks",
                             [session_p, ctypes.c_int, ctypes.c_int])
compile_device_source = _sess("wb200_compile_device_source", [
    ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t])
session_run_evals = _sess("wb200_session_run_evals", [
    session_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int])
session_iter_stats = _sess("wb200_session_iter_stats", [
    session_p, ctypes.c_int, ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")])
session_chain_rows = _sess("wb200_session_chain_rows", [
    session_p, ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")])
session_rhat_moments = _sess("wb200_session_rhat_moments", [
    session_p, ctypes.c_longlong, double_array])
session_summary = _sess("wb200_session_summary", [
    session_p, ctypes.c_longlong, nullable_double_array, nullable_double_array,
    nullable_double_array, nullable_double_array, nullable_double_array])
session_warmup_sums = _sess("wb200_session_warmup_sums", [session_p, ctypes.c_void_p])
session_warmup_deviation = _sess("wb200_session_warmup_deviation",
                                 [session_p, ctypes.c_void_p, double_array])
session_lp_moments = _sess("wb200_session_lp_moments", [session_p, double_array])
session_lp_moments_centered = _sess("wb200_session_lp_moments_centered",
                                    [session_p, ctypes.c_double, double_array])
_lib.wb200_session_logp_exceptions.restype = ctypes.c_int
_lib.wb200_session_logp_exceptions.argtypes = [session_p,
                                               ctypes.POINTER(ctypes.c_ulonglong)]
session_logp_exceptions = _lib.wb200_session_logp_exceptions
session_stream_begin = _sess("wb200_session_stream_begin", [session_p, ctypes.c_int])
session_stream_phase1 = _sess("wb200_session_stream_phase1", [session_p, double_array])
session_stream_phase2 = _sess("wb200_session_stream_phase2",
                              [session_p, double_array, double_array])
stream_finish = _sess("wb200_stream_finish", [
    ctypes.c_int, ctypes.c_int, double_array, double_array, nullable_double_array,
    nullable_double_array, nullable_double_array, nullable_double_array,
    nullable_double_array, nullable_int_array])
session_stream_summary = _sess("wb200_session_stream_summary", [
    session_p, nullable_double_array, nullable_double_array, nullable_double_array,
    nullable_double_array, nullable_double_array, nullable_int_array])
session_stream_counts = _sess("wb200_session_stream_counts", [
    session_p, ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")])
session_get_draws = _sess("wb200_session_get_draws", [
    session_p, ctypes.c_longlong, ctypes.c_longlong, double_array])
session_get_trace = _sess("wb200_session_get_trace", [
    session_p, ctypes.c_longlong, ctypes.c_longlong, nullable_double_array,
    nullable_int_array, nullable_double_array, nullable_double_array])
session_get_state = _sess("wb200_session_get_state", [
    session_p, nullable_double_array, nullable_double_array, nullable_double_array,
    nullable_int_array, ctypes.c_void_p])
session_counters = _sess("wb200_session_counters", [
    session_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong),
    ctypes.POINTER(ctypes.c_ulonglong)])
_lib.wb200_session_device_draws.restype = ctypes.c_int
_lib.wb200_session_device_draws.argtypes = [
    session_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_longlong),
    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)]
session_device_draws = _lib.wb200_session_device_draws
_lib.wb200_session_last_kernel_ms.restype = ctypes.c_int
_lib.wb200_session_last_kernel_ms.argtypes = [session_p, ctypes.POINTER(ctypes.c_float)]
session_last_kernel_ms = _lib.wb200_session_last_kernel_ms

session_timer_record = _sess("wb200_session_timer_record", [session_p, ctypes.c_int])
session_timer_elapsed_ms = _sess("wb200_session_timer_elapsed_ms",
                                 [session_p, ctypes.POINTER(ctypes.c_float)])
_lib.wb200_last_run_stats.restype = ctypes.c_int
_lib.wb200_last_run_stats.argtypes = [
    ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong),
    ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_int),
    ctypes.POINTER(ctypes.c_int)]


def last_run_stats():
    g, m, l = (ctypes.c_ulonglong(0) for _ in range(3))
    w, s_ = ctypes.c_int(0), ctypes.c_int(0)
    _lib.wb200_last_run_stats(ctypes.byref(g), ctypes.byref(m), ctypes.byref(l),
                              ctypes.byref(w), ctypes.byref(s_))
    return dict(grad_evals=g.value, macro_steps=m.value, kernel_launches=l.value,
                warmup_iters=w.value, sampling_iters=s_.value)


_host_alloc = _sess("wb200_host_alloc", [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)])
trim_memory = _sess("wb200_trim_memory", [ctypes.c_int])
_lib.wb200_host_free.restype = None
_lib.wb200_host_free.argtypes = [ctypes.c_void_p]


def pinned_empty(shape, dtype=np.float64):
    """A numpy array over page-locked host memory (wb200_host_alloc); freed with the
    array.  Contents are uninitialised."""
    import weakref
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64))
    p = ctypes.c_void_p()
    _host_alloc(max(n, 1) * dtype.itemsize, ctypes.byref(p))
    buf = (ctypes.c_char * (max(n, 1) * dtype.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    weakref.finalize(buf, _lib.wb200_host_free, p.value)
    return arr


orbit = _sess("wb200_orbit", [
    ctypes.POINTER(WalnutModelDesc), ctypes.c_size_t, double_array, double_array,
    double_array, ctypes.c_double, ctypes.c_int, double_array, double_array,
    double_array, double_array, double_array])
philox = _sess("wb200_philox", [
    ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS"), ctypes.c_size_t,
    ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")])
philox_normals = _sess("wb200_philox_normals", [
    ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_size_t,
    double_array])
device_summary = _sess("wb200_device_summary", [
    ctypes.c_void_p, ctypes.c_size_t, ctypes.c_longlong, ctypes.c_longlong,
    ctypes.c_longlong, ctypes.c_int, ctypes.c_int, nullable_double_array,
    nullable_double_array, nullable_double_array, nullable_double_array,
    nullable_double_array])

logistic_logp_grad = _sess("wb200_logistic_logp_grad", [
    double_array, double_array, ctypes.c_size_t, ctypes.c_int, double_array,
    ctypes.c_size_t, double_array, double_array, ctypes.c_int,
    ctypes.POINTER(ctypes.c_float)])

EXPORTED_SYMBOLS = [
    "walnutpie_sample_device", "walnutpie_sample_cfunc", "walnutpie_separator_char",
    "walnutpie_ess", "walnutpie_r_hat", "walnutpie_mcse",
    "walnutpie_get_error_message", "walnutpie_get_error_type",
    "walnutpie_destroy_error", "walnuts_b200_default_tuning", "walnuts_b200_version",
    "wb200_session_create", "wb200_session_destroy", "wb200_session_init",
    "wb200_session_reserve_draws", "wb200_session_warmup", "wb200_session_freeze",
    "wb200_session_sample", "wb200_session_sync", "wb200_session_sample_ticks",
    "wb200_session_warmup_ticks",
    "wb200_session_run_evals", "wb200_session_iter_stats", "wb200_compile_device_source",
    "wb200_session_chain_rows", "wb200_session_summary", "wb200_session_rhat_moments", "wb200_session_warmup_sums",
    "wb200_session_warmup_deviation", "wb200_session_lp_moments",
    "wb200_session_lp_moments_centered", "wb200_session_logp_exceptions",
    "walnutpie_sample_bridgestan", "walnutpie_sample_device_summary",
    "walnutpie_sample_device_multi",
    "wb200_session_stream_begin",
    "wb200_session_stream_phase1", "wb200_session_stream_phase2", "wb200_stream_finish",
    "wb200_session_stream_summary", "wb200_session_stream_counts",
    "wb200_session_get_draws", "wb200_session_get_trace", "wb200_session_get_state",
    "wb200_session_device_draws", "wb200_session_counters",
    "wb200_session_last_kernel_ms", "wb200_session_timer_record",
    "wb200_session_timer_elapsed_ms", "wb200_host_alloc", "wb200_host_free", "wb200_trim_memory",
    "wb200_last_run_stats", "wb200_orbit", "wb200_philox",
    "wb200_philox_normals", "wb200_device_summary", "wb200_logistic_logp_grad",
]
