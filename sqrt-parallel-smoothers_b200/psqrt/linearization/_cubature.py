"""Spherical cubature rule (reference: parsmooth/linearization/_cubature.py:10-85)."""
from __future__ import annotations

import numpy as np
import torch

from .._base import FunctionalModel
from ._common import require_sqrt
from ._sigma_points import linearize_conditional, linearize_functional


def _cubature_weights(n_dim: int):
    """2n points xi = +-sqrt(n) e_i, weights 1/(2n)   (_cubature.py:63-85)"""
    wm = np.ones(shape=(2 * n_dim,)) / (2 * n_dim)
    xi = np.concatenate([np.eye(n_dim), -np.eye(n_dim)], axis=0) * np.sqrt(n_dim)
    return wm, wm, xi


def linearize(model, x):
    require_sqrt(x)
    if isinstance(model, FunctionalModel):
        builtin = getattr(model.function, "_psqrt_builtin", None)
    else:
        builtin = getattr(model[0], "_psqrt_builtin", None)
    n = x.mean.shape[-1]
    wm, wc, xi = _cubature_weights(n)
    if builtin is not None and hasattr(builtin, "slr"):
        out = builtin.slr(model, x, xi, wm, wc)
        if out is not None:
            return out
    dev = x.mean.device
    xi_t, wm_t, wc_t = (torch.as_tensor(a, dtype=torch.float64, device=dev) for a in (xi, wm, wc))
    if isinstance(model, FunctionalModel):
        f, q = model
        return linearize_functional(f, x, q, xi_t, wm_t, wc_t)
    return linearize_conditional(model[0], model[1], x, xi_t, wm_t, wc_t)


# tags read by psqrt.grad (tangent of the linearization): the rule and its unit sigma points (wm, wc, xi)
linearize._psqrt_kind = "slr"
linearize._psqrt_points = lambda n: _cubature_weights(n)
