"""Dispatch object attached to the built-in model functions (attribute `_psqrt_builtin`): on CUDA
tensors the linearization methods call the device kernels of csrc/psqrt_models.cu through
psqrt_linearize_builtin; on CPU tensors (a user calling a linearization method directly on host data)
they use the function's analytic torch Jacobian like for any other callable."""
from __future__ import annotations

import numpy as np

from .. import _lib


class Builtin:
    def __init__(self, model_id, params, n_in, n_out, conditional=False):
        self.model_id, self.params, self.n_in, self.n_out, self.conditional = model_id, list(params), n_in, n_out, conditional

    # extended, functional model:  (F, chol_q, f(m) - F m + m_q)          linearization/_extended.py:68-70
    def extended(self, x, q):
        if not x.mean.is_cuda:
            return None
        F, _, b = _lib.linearize_builtin(self.model_id, self.params, _lib.LIN_EXTENDED, self.n_in, self.n_out, False,
                                         x.mean, m_q=q.mean)
        return F, q.chol, b

    # extended, conditional-moments model: (F, c_chol(m), c_m(m) - F m)   linearization/_extended.py:51-56
    def extended_conditional(self, x):
        if not x.mean.is_cuda:
            return None
        return _lib.linearize_builtin(self.model_id, self.params, _lib.LIN_EXTENDED, self.n_in, self.n_out, True, x.mean)

    # statistical linear regression from unit points xi [P, n] and weights   linearization/_sigma_points.py:25-100
    def slr(self, model, x, xi, wm, wc):
        if not x.mean.is_cuda:
            return None
        pts = (np.ascontiguousarray(xi, dtype=np.float64), np.ascontiguousarray(wm, dtype=np.float64),
               np.ascontiguousarray(wc, dtype=np.float64))
        if self.conditional:
            return _lib.linearize_builtin(self.model_id, self.params, _lib.LIN_SLR, self.n_in, self.n_out, True,
                                          x.mean, x.chol, points=pts)
        q = model.mvn
        return _lib.linearize_builtin(self.model_id, self.params, _lib.LIN_SLR, self.n_in, self.n_out, False,
                                      x.mean, x.chol, q.mean, q.chol, points=pts)
