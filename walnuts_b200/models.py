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

KINDS = {"std_normal": 0, "diag_gaussian": 1, "funnel": 2, "logistic": 3}


@dataclass
class DeviceModel:
    kind: str
    num_params: int
    precision: Optional[np.ndarray] = None   # diag_gaussian: 1 / sigma_d^2
    X: Optional[np.ndarray] = None           # logistic: [N][D]
    y: Optional[np.ndarray] = None           # logistic: [N]
    _keep: list = field(default_factory=list, repr=False)

    def desc(self) -> WalnutModelDesc:
        d = WalnutModelDesc(kind=KINDS[self.kind], D=int(self.num_params), N=0,
                            data0=None, data1=None)
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
        return d


def std_normal(num_params: int) -> DeviceModel:
    """p(x) = N(0, I) (examples/walnutpie_api.cpp:39-43)."""
    return DeviceModel("std_normal", num_params)


def diag_gaussian(variances) -> DeviceModel:
    """p(x) = N(0, diag(variances)); generalises ``ill_normal``
    (examples/examples.cpp:20-31)."""
    v = np.asarray(variances, dtype=np.float64)
    return DeviceModel("diag_gaussian", v.size, precision=1.0 / v)


def ill_conditioned_gaussian(num_params: int, condition: float = 1e4) -> DeviceModel:
    """BASELINE config c2: variances log-spaced over `condition`."""
    d = np.arange(num_params, dtype=np.float64)
    var = condition ** (d / max(num_params - 1, 1))
    return diag_gaussian(var)


def funnel(num_params: int) -> DeviceModel:
    """Neal's funnel: x0 ~ N(0, 9), x_i | x0 ~ N(0, exp(x0))."""
    return DeviceModel("funnel", num_params)


def logistic(X, y) -> DeviceModel:
    """Bayesian logistic regression with a N(0, I) prior: logp = sum_n [y_n z_n -
    softplus(z_n)] - |theta|^2 / 2, z = X theta.  X is rounded to bf16 on the device
    (the gradient is two tensor-core GEMMs batched over all chains)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be [N][D] and y [N]")
    return DeviceModel("logistic", X.shape[1], X=X, y=y)
