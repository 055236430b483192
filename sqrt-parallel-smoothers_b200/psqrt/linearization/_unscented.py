"""Unscented transform (reference: parsmooth/linearization/_unscented.py:10-101).

2n + 1 points: the mean and +-sqrt(n + lamda) times the columns of the factor; lamda = alpha^2 (n + kappa) - n
with the reference's defaults alpha = 1, beta = 0, kappa = 3 + n (its line 66)."""
from __future__ import annotations

import numpy as np
import torch

from .._base import FunctionalModel
from ._common import require_sqrt
from ._sigma_points import linearize_conditional, linearize_functional


def _unscented_weights(n_dim: int, alpha: float, beta: float, kappa):
    """(_unscented.py:42-58, 73-101) -> wm, wc, unit points xi [2n + 1, n]"""
    if kappa is None:
        kappa = 3.0 + n_dim
    lamda = alpha ** 2 * (n_dim + kappa) - n_dim
    wm = np.full(2 * n_dim + 1, 1 / (2 * (n_dim + lamda)))
    wm[0] = lamda / (n_dim + lamda)
    wc = wm.copy()
    wc[0] = lamda / (n_dim + lamda) + (1 - alpha ** 2 + beta)
    xi = np.concatenate([np.zeros((1, n_dim)), np.eye(n_dim), -np.eye(n_dim)], axis=0) * np.sqrt(n_dim + lamda)
    return wm, wc, xi


def linearize(model, x, alpha: float = 1.0, beta: float = 0.0, kappa=None):
    require_sqrt(x)
    if isinstance(model, FunctionalModel):
        builtin = getattr(model.function, "_psqrt_builtin", None)
    else:
        builtin = getattr(model[0], "_psqrt_builtin", None)
    n = x.mean.shape[-1]
    wm, wc, xi = _unscented_weights(n, alpha, beta, kappa)
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
linearize._psqrt_points = lambda n, alpha=1.0, beta=0.0, kappa=None: _unscented_weights(n, alpha, beta, kappa)
