"""Host buffers and small argument helpers of the Python front end.

Behaviour follows the reference's helpers (python/src/walnutpie/util.py: seeds :7-13,
output buffer :16-32, initial metric :35-47, WarmupInfo :53-70); what is new is that the
output buffer can live in page-locked memory, so that the library's read-back of the draws
is a DMA that overlaps sampling.
"""
from __future__ import annotations

import secrets
from dataclasses import dataclass
from typing import Generic, Optional, TypeVar

import numpy as np

Draws = TypeVar("Draws")


@dataclass
class WarmupInfo(Generic[Draws]):
    """What adaptation leaves behind for one chain."""

    stepsize: float
    inv_metric: Optional[np.ndarray]
    warmup_draws: Optional[Draws]


def prepare_seed(seed: Optional[int]) -> int:
    """The caller's seed, or a fresh 32-bit one."""
    if seed is None:
        return secrets.randbelow(2**32 - 1)
    return seed


def _check_run_lengths(num_chains: int, max_sampling_iter: int, max_warmup_iter: int) -> None:
    problems = {
        "num_chains must be at least 1": num_chains < 1,
        "max_warmup_iter must be non-negative": max_warmup_iter < 0,
        "max_sampling_iter must be at least 1": max_sampling_iter < 1,
    }
    for message, failed in problems.items():
        if failed:
            raise ValueError(message)


def prepare_output_buffer(*, num_chains: int, num_params: int, max_sampling_iter: int,
                          max_warmup_iter: int, save_warmup: bool,
                          pinned: bool = False) -> np.ndarray:
    """``[chains][warm-up rows (if saved) + sampling rows][params]`` float64.

    ``pinned=True`` asks the library for page-locked memory (rows beyond the lengths a run
    returns are then uninitialised rather than zero; callers slice by length).  If
    page-locking is refused the buffer is ordinary zeroed memory, which works the same,
    only with a slower read-back."""
    _check_run_lengths(num_chains, max_sampling_iter, max_warmup_iter)
    rows = max_sampling_iter + (max_warmup_iter if save_warmup else 0)
    shape = (num_chains, rows, num_params)
    if pinned:
        from . import _ffi
        try:
            return _ffi.pinned_empty(shape)
        except (RuntimeError, MemoryError):
            pass
    return np.zeros(shape, dtype=np.float64)


def prepare_inv_metric(init_inv_metric: Optional[np.ndarray], metric_size: tuple,
                       num_chains: int) -> Optional[np.ndarray]:
    """One metric for all chains, or one per chain, as a contiguous ``[chains, ...]`` array."""
    if init_inv_metric is None:
        return None
    metric = np.ascontiguousarray(init_inv_metric, dtype=np.float64)
    per_chain = (num_chains, *metric_size)
    if metric.shape == per_chain:
        return metric
    if metric.shape == metric_size:
        return np.ascontiguousarray(np.broadcast_to(metric, per_chain))
    raise ValueError(f"Invalid initial metric size. Expected a {metric_size} "
                     f"or {per_chain} matrix.")
