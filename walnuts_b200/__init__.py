"""walnuts_b200: a B200-native WALNUTS sampler behind walnutpie's API.

Public names mirror python/src/walnutpie/__init__.py; ``walnuts_pyfunc`` is
replaced by ``walnuts_device`` (same keywords, a device model instead of a host
callback).  Importing this package loads libwalnuts_b200.so and fails if it has
not been built — there is no CPU path.
"""
from . import models
from ._ffi import logp_cfunc_type
from .models import DeviceModel
from .sampler import (Session, WalnutsOutputArray, orbit, walnuts_device,
                      walnuts_device_summary)
from .summary import (Summarizer, ess, mcse, mean, r_hat, standard_deviation,
                      variance)

__all__ = [
    "walnuts_device", "walnuts_device_summary", "Session", "DeviceModel", "models", "orbit",
    "WalnutsOutputArray", "logp_cfunc_type", "r_hat", "ess", "mcse", "mean",
    "variance", "standard_deviation", "Summarizer",
]
__version__ = "0.1.0"
