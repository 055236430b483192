"""psqrt -- B200-native parallel square-root Kalman filtering / RTS smoothing.

Drop-in for the ``sqrt=True, parallel=True`` path of EEA-sensors/sqrt-parallel-smoothers
(``parsmooth``): same API (``psqrt.methods``), same containers, same linearization protocol;
the numerics run in hand-written sm_100a CUDA kernels (libpsqrt.so) behind a C ABI.
"""
from ._base import MVNStandard, MVNSqrt, FunctionalModel, ConditionalMomentsModel, are_inputs_compatible
from .methods import filtering, smoothing, filter_smoother, iterated_smoothing, sampling
from . import fp32, grad, linearization, methods, models

__all__ = ["MVNStandard", "MVNSqrt", "FunctionalModel", "ConditionalMomentsModel", "are_inputs_compatible",
           "filtering", "smoothing", "filter_smoother", "iterated_smoothing", "sampling", "linearization", "methods", "models", "grad", "fp32"]
