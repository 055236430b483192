"""First-order Taylor linearization (reference: parsmooth/linearization/_extended.py:9-70)."""
from __future__ import annotations

from .._base import FunctionalModel, MVNSqrt, are_inputs_compatible
from ._common import apply_fn, mv, require_sqrt, value_and_jac


def linearize(model, x):
    """F = df/dx(m), b = f(m) - F m + m_q, noise factor passed through (_extended.py:68-70);
    conditional-moments models: F = d c_m/dx, b = c_m(m) - F m, chol = c_chol(m) (51-56).
    Uses only the nominal mean."""
    require_sqrt(x)
    if isinstance(model, FunctionalModel):
        f, q = model
        are_inputs_compatible(x, q)
        require_sqrt(q)
        builtin = getattr(f, "_psqrt_builtin", None)
        if builtin is not None:
            out = builtin.extended(x, q)
            if out is not None:
                return out
        m_x = x.mean
        res, F_x = value_and_jac(f, m_x)
        return F_x, q.chol, res - mv(F_x, m_x) + q.mean
    c_m, c_chol = model
    builtin = getattr(c_m, "_psqrt_builtin", None)
    if builtin is not None:
        out = builtin.extended_conditional(x)
        if out is not None:
            return out
    m = x.mean
    res, F = value_and_jac(c_m, m)
    return F, apply_fn(c_chol, m), res - mv(F, m)


# tags read by psqrt.grad (tangent of the linearization): the rule, and its unit sigma points if it has any
linearize._psqrt_kind = "extended"
linearize._psqrt_points = None
