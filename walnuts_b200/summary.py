"""Posterior summaries, same surface as python/src/walnutpie/summary.py.

``ess`` / ``r_hat`` / ``mcse`` run on the GPU through the library's
``walnutpie_ess`` / ``_r_hat`` / ``_mcse`` (include/walnuts_b200.h), which take
the stacked ``(draws, params)`` matrix ROW-MAJOR — the layout summary.py:30
really passes.  (The reference's C shim maps that buffer column-major,
walnutpy.cpp:86-95, so for num_params > 1 its Python numbers are computed on a
scrambled view; the C++ functions pinned by tests/summary_test.cpp are what is
reproduced here.)
"""
from typing import List

import numpy as np

from ._ffi import _ffi_ess, _ffi_mcse, _ffi_r_hat


class Summarizer:
    """Holds multivariate MCMC draws and summarises them (summary.py:11-145)."""

    def __init__(self, draws: List[np.ndarray]):
        self._stacked = np.ascontiguousarray(np.concatenate(draws), dtype=np.float64)
        self._num_draws, self._num_params = self._stacked.shape
        self._lengths = np.array([c.shape[0] for c in draws], dtype=np.int32)
        self._num_chains = len(draws)

    def mean(self):
        return np.mean(self._stacked, axis=0)

    def variance(self):
        return np.var(self._stacked, axis=0, ddof=1)

    def standard_deviation(self):
        return np.std(self._stacked, axis=0, ddof=1)

    def _call(self, f):
        out = np.zeros((self._num_params,))
        f(self._stacked, self._num_draws, self._num_params, self._lengths,
          self._num_chains, out)
        return out

    def ess(self) -> np.ndarray:
        return self._call(_ffi_ess)

    def r_hat(self) -> np.ndarray:
        return self._call(_ffi_r_hat)

    def mcse(self) -> np.ndarray:
        return self._call(_ffi_mcse)


def ess(draws):
    return Summarizer(draws).ess()


def r_hat(draws):
    return Summarizer(draws).r_hat()


def mcse(draws):
    return Summarizer(draws).mcse()


def mean(draws):
    return Summarizer(draws).mean()


def variance(draws):
    return Summarizer(draws).variance()


def standard_deviation(draws):
    return Summarizer(draws).standard_deviation()
