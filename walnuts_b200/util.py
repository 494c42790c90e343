"""Output-buffer helpers; same behaviour as python/src/walnutpie/util.py."""
from dataclasses import dataclass
from typing import Generic, Optional, TypeVar

import numpy as np


def rand_u32():
    """Generate a random 32-bit unsigned integer (util.py:7-9)."""
    return int(np.random.randint(0, 2**32 - 1, dtype=np.uint32))


def prepare_seed(seed: Optional[int]) -> int:
    return seed if seed is not None else rand_u32()


def prepare_output_buffer(*, num_chains: int, num_params: int, max_sampling_iter: int,
                          max_warmup_iter: int, save_warmup: bool,
                          pinned: bool = False) -> np.ndarray:
    # util.py:16-32; pinned=True places the buffer in page-locked memory so that the
    # library's draw read-back is a DMA overlapping sampling (rows beyond the returned
    # lengths are then uninitialised instead of zero; callers slice by length)
    if num_chains < 1:
        raise ValueError("num_chains must be at least 1")
    if max_warmup_iter < 0:
        raise ValueError("max_warmup_iter must be non-negative")
    if max_sampling_iter < 1:
        raise ValueError("max_sampling_iter must be at least 1")
    num_draws = max_sampling_iter + max_warmup_iter * save_warmup
    shape = (num_chains, num_draws, num_params)
    if pinned:
        from . import _ffi
        try:
            return _ffi.pinned_empty(shape)
        except (RuntimeError, MemoryError):
            pass  # page-locking refused (size limits): a pageable buffer still works
    return np.zeros(shape, dtype=np.float64)


def prepare_inv_metric(init_inv_metric: Optional[np.ndarray], metric_size: tuple,
                       num_chains: int) -> Optional[np.ndarray]:
    # util.py:35-47
    if init_inv_metric is not None:
        init_inv_metric = np.ascontiguousarray(init_inv_metric, dtype=np.float64)
        if init_inv_metric.shape == metric_size:
            return np.ascontiguousarray(
                np.repeat(init_inv_metric[np.newaxis], num_chains, axis=0))
        elif init_inv_metric.shape == (num_chains, *metric_size):
            return init_inv_metric
        else:
            raise ValueError(
                f"Invalid initial metric size. Expected a {metric_size} "
                f"or {(num_chains, *metric_size)} matrix.")
    return None


T = TypeVar("T")


@dataclass
class WarmupInfo(Generic[T]):
    """Warm-up output of one chain (util.py:53-70)."""

    stepsize: float
    inv_metric: Optional[np.ndarray]
    warmup_draws: Optional[T]
