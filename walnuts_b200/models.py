"""Device model descriptors: what replaces the ``logp`` callback.

The reference hands a host function pointer (LOGP_CFUNC, walnutpy.cpp:127-132)
to C threads; a device batch needs the density on the GPU, so a model is named
by a small descriptor (``WalnutModelDesc`` in include/walnuts_b200.h).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from ._ffi import WalnutModelDesc

KINDS = {"std_normal": 0, "diag_gaussian": 1, "funnel": 2, "logistic": 3, "batch_callback": 4,
         "device_source": 5}

# WB200_BATCH_LOGP_GRAD (include/walnuts_b200.h): num_chains, num_params, ld, theta, grad,
# lp (device pointers), cuda_stream, data -> 0 on success
BATCH_LOGP_GRAD = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t,
                                   ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)


@dataclass
class DeviceModel:
    kind: str
    num_params: int
    precision: Optional[np.ndarray] = None   # diag_gaussian: 1 / sigma_d^2
    X: Optional[np.ndarray] = None           # logistic: [N][D]
    y: Optional[np.ndarray] = None           # logistic: [N]
    callback: Optional[object] = None        # batch_callback: a BATCH_LOGP_GRAD instance
    source: Optional[str] = None             # device_source: CUDA source text
    params: Optional[np.ndarray] = None      # device_source: the density's parameters
    dtype: str = "f64"                       # "f32": fp32 mode (element-wise targets)
    errors: list = field(default_factory=list, repr=False)  # exceptions raised inside it
    _keep: list = field(default_factory=list, repr=False)

    def desc(self) -> WalnutModelDesc:
        if self.dtype not in ("f64", "f32"):
            raise ValueError("dtype must be 'f64' or 'f32'")
        d = WalnutModelDesc(kind=KINDS[self.kind], D=int(self.num_params), N=0,
                            data0=None, data1=None,
                            precision=1 if self.dtype == "f32" else 0)
        if self.kind == "diag_gaussian":
            p = np.ascontiguousarray(self.precision, dtype=np.float64)
            if p.shape != (self.num_params,):
                raise ValueError("precision must have shape (num_params,)")
            self._keep.append(p)
            d.data0 = p.ctypes.data
        elif self.kind == "logistic":
            X = np.ascontiguousarray(self.X, dtype=np.float64)
            y = np.ascontiguousarray(self.y, dtype=np.float64)
            self._keep += [X, y]
            d.N = X.shape[0]
            d.data0 = X.ctypes.data
            d.data1 = y.ctypes.data
        elif self.kind == "batch_callback":
            d.data0 = ctypes.cast(self.callback, ctypes.c_void_p).value
        elif self.kind == "device_source":
            if self.dtype != "f64":
                raise ValueError("a run-time compiled density runs in fp64")
            src = ctypes.create_string_buffer(self.source.encode())
            par = np.ascontiguousarray(
                self.params if self.params is not None else np.zeros(0), dtype=np.float64)
            self._keep += [src, par]
            d.N = par.size
            d.data0 = ctypes.cast(src, ctypes.c_void_p).value
            d.data1 = par.ctypes.data if par.size else None
        return d


def std_normal(num_params: int, dtype: str = "f64") -> DeviceModel:
    """p(x) = N(0, I) (examples/walnutpie_api.cpp:39-43).  dtype "f32": fp32 mode
    (include/walnuts_b200.h, WalnutTuning::precision)."""
    return DeviceModel("std_normal", num_params, dtype=dtype)


def diag_gaussian(variances, dtype: str = "f64") -> DeviceModel:
    """p(x) = N(0, diag(variances)); generalises ``ill_normal``
    (examples/examples.cpp:20-31)."""
    v = np.asarray(variances, dtype=np.float64)
    return DeviceModel("diag_gaussian", v.size, precision=1.0 / v, dtype=dtype)


def ill_conditioned_gaussian(num_params: int, condition: float = 1e4,
                             dtype: str = "f64") -> DeviceModel:
    """BASELINE config c2: variances log-spaced over `condition`."""
    d = np.arange(num_params, dtype=np.float64)
    var = condition ** (d / max(num_params - 1, 1))
    return diag_gaussian(var, dtype=dtype)


def funnel(num_params: int, dtype: str = "f64") -> DeviceModel:
    """Neal's funnel: x0 ~ N(0, 9), x_i | x0 ~ N(0, exp(x0))."""
    return DeviceModel("funnel", num_params, dtype=dtype)


def logistic(X, y) -> DeviceModel:
    """Bayesian logistic regression with a N(0, I) prior: logp = sum_n [y_n z_n -
    softplus(z_n)] - |theta|^2 / 2, z = X theta.  X is rounded to bf16 on the device
    (the gradient is two tensor-core GEMMs batched over all chains)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be [N][D] and y [N]")
    return DeviceModel("logistic", X.shape[1], X=X, y=y)


def batch_callback(num_params: int, fn) -> DeviceModel:
    """The caller's own density, batched on the device (WalnutModelDesc kind 4).

    ``fn(num_chains, num_params, ld, theta_ptr, grad_ptr, lp_ptr, stream_ptr)`` is called
    once per lock-step tick with raw device pointers (fp64; ``theta`` / ``grad`` are
    ``[num_chains][ld]`` row-major, ``lp`` is ``[num_chains]``) and the CUDA stream on which
    it must enqueue the work that fills ``grad`` and ``lp``.  An exception raised inside
    ``fn`` is handled like the reference's trampoline handles one (pyfunc.py:32-42,
    util.hpp:336-346): it is printed and the call returns 1; during initialisation that
    ends the run (the sampler call re-raises the exception), inside a transition every
    chain of the tick continues with ``logp = -inf`` and a zero gradient.  The last
    exceptions are kept in ``model.errors``."""
    model = DeviceModel("batch_callback", int(num_params))

    def trampoline(C, D, ld, theta, grad, lp, stream, _data):
        try:
            fn(C, D, ld, theta, grad, lp, stream)
            return 0
        except BaseException as exc:  # must not propagate through the C frames
            print(exc)
            if len(model.errors) < 16:
                model.errors.append(exc)
            return 1

    model.callback = BATCH_LOGP_GRAD(trampoline)
    return model


class _DevicePointer:
    """A raw device allocation as ``__cuda_array_interface__`` (fp64, C order)."""

    def __init__(self, ptr: int, shape: tuple):
        self.__cuda_array_interface__ = {
            "shape": shape, "typestr": "<f8", "data": (int(ptr), False), "version": 3,
            "strides": None}


def torch_density(num_params: int, logp_fn, grad_fn=None) -> DeviceModel:
    """A density written with PyTorch, evaluated for all chains at once on the GPU.

    ``logp_fn(theta)`` maps a float64 CUDA tensor ``[num_chains, num_params]`` to the log
    densities ``[num_chains]``; the gradient comes from autograd unless ``grad_fn(theta)``
    -> ``(logp, grad)`` is given.  This is what ``walnuts_pyfunc``'s Python ``logp``
    callback (pyfunc.py:45-83) becomes when the chains live on the device."""
    import torch

    def fn(C, D, ld, theta_ptr, grad_ptr, lp_ptr, stream_ptr):
        stream = torch.cuda.ExternalStream(int(stream_ptr or 0))
        with torch.cuda.stream(stream):
            theta = torch.as_tensor(_DevicePointer(theta_ptr, (C, ld)), device="cuda")[:, :D]
            grad = torch.as_tensor(_DevicePointer(grad_ptr, (C, ld)), device="cuda")
            lp = torch.as_tensor(_DevicePointer(lp_ptr, (C,)), device="cuda")
            if grad_fn is not None:
                with torch.no_grad():
                    value, g = grad_fn(theta)
            else:
                x = theta.detach().clone().requires_grad_(True)
                with torch.enable_grad():
                    value = logp_fn(x)
                    (g,) = torch.autograd.grad(value.sum(), x)
            grad[:, :D].copy_(g)
            lp.copy_(value.detach().reshape(C))

    return batch_callback(num_params, fn)


def device_source(source: str, num_params: int, params=None) -> DeviceModel:
    """The caller's own density as CUDA source, compiled at run time into the
    chain-resident kernel (model kind 5, include/walnuts_b200.h) -- the device counterpart
    of handing the reference a ``logp`` callable (pyfunc.py:45-83).  Element-wise densities
    define ``__device__ void wb200_logp_grad(int d, double x, const double* par, double& lp,
    double& g)``; ``params`` reaches it as ``par``."""
    return DeviceModel("device_source", int(num_params), source=str(source),
                       params=None if params is None else np.asarray(params, np.float64))


def compile_device_source(source: str, num_params: int) -> str:
    """Compile only (no GPU needed); returns the compiler's output, raises ``ValueError``
    with its log if the source does not compile."""
    from . import _ffi
    log = ctypes.create_string_buffer(1 << 16)
    _ffi.compile_device_source(source.encode(), int(num_params), log, len(log))
    return log.value.decode()
