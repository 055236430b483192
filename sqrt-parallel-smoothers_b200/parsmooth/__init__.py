"""Drop-in alias: the reference's package name (`parsmooth`) over the B200 implementation (`psqrt`).

    from parsmooth.methods import filtering, smoothing, iterated_smoothing
    from parsmooth.linearization import extended, cubature, gauss_hermite, unscented
    from parsmooth._base import MVNSqrt, FunctionalModel

Only the parallel square-root path exists (see psqrt.methods); everything else raises NotImplementedError.
"""
import sys

import psqrt
from psqrt import (MVNStandard, MVNSqrt, FunctionalModel, ConditionalMomentsModel,  # noqa: F401
                   filtering, smoothing, iterated_smoothing, filter_smoother, sampling)
from psqrt import _base, linearization, methods, models  # noqa: F401

__version__ = "1.0.0+psqrt"

for _name, _mod in (("_base", _base), ("linearization", linearization), ("methods", methods), ("models", models)):
    sys.modules[f"{__name__}.{_name}"] = _mod
