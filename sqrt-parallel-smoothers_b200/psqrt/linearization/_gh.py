"""Gauss-Hermite product rule (reference: parsmooth/linearization/_gh.py:13-151)."""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np
import torch

from .._base import FunctionalModel
from ._common import require_sqrt
from ._sigma_points import linearize_conditional, linearize_functional


def _hermite_coeff(order: int):
    """physicists' Hermite polynomials H_0..H_order, highest power first (_gh.py:129-151)"""
    H = [np.array([1]), np.array([2, 0])]
    for i in range(2, order + 1):
        H.append(2 * np.append(H[i - 1], 0) - 2 * (i - 1) * np.pad(H[i - 2], (2, 0), "constant", constant_values=0))
    return H


@lru_cache(maxsize=None)
def _gauss_hermite_weights(n_dim: int, order: int = 3):
    """order**n points (_gh.py:73-126).  The 1-D nodes come from np.roots exactly as upstream
    (they are asymmetric at 1e-16 and the parity contract is 1e-9, so the procedure is kept);
    dimension r of point j uses node (j // order**r) % order."""
    hc = _hermite_coeff(order)
    roots = np.flip(np.roots(hc[-1]))
    w_1d = np.array([2 ** (order - 1) * math.factorial(order) * np.sqrt(np.pi) /
                     (order ** 2 * (np.polyval(hc[order - 1], roots[i])) ** 2) for i in range(order)])
    j = np.arange(order ** n_dim)
    table = np.stack([(j // (order ** r)) % order for r in range(n_dim)], axis=0)
    s = 1 / (np.sqrt(np.pi) ** n_dim)
    wm = s * np.prod(w_1d[table], axis=0)
    xi = (np.sqrt(2) * roots[table]).T.copy()        # [P, n]
    return wm, wm, xi


def linearize(model, x, order: int = 3):
    require_sqrt(x)
    if isinstance(model, FunctionalModel):
        builtin = getattr(model.function, "_psqrt_builtin", None)
    else:
        builtin = getattr(model[0], "_psqrt_builtin", None)
    n = x.mean.shape[-1]
    wm, wc, xi = _gauss_hermite_weights(n, order)
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
linearize._psqrt_points = lambda n, order=3: _gauss_hermite_weights(n, order)
