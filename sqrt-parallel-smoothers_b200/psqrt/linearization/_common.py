"""Shared helpers of the linearization methods (host side, torch)."""
from __future__ import annotations

import torch

from .._base import ConditionalMomentsModel, FunctionalModel, MVNSqrt, MVNStandard


def require_sqrt(x):
    if isinstance(x, MVNStandard):
        raise NotImplementedError("psqrt implements the square-root path only: pass MVNSqrt, not MVNStandard")
    if not isinstance(x, MVNSqrt):
        raise TypeError(f"expected MVNSqrt, got {type(x)}")


def as_f64(t, device=None):
    t = torch.as_tensor(t, dtype=torch.float64)
    if device is not None and t.device != device:
        t = t.to(device)
    return t


def apply_fn(f, x: torch.Tensor) -> torch.Tensor:
    """Evaluate a user function of a 1-D state on [..., n]."""
    if getattr(f, "_psqrt_batched", False):
        return f(x)
    if x.dim() == 1:
        return f(x)
    flat = x.reshape(-1, x.shape[-1])
    out = torch.func.vmap(f)(flat)
    return out.reshape(*x.shape[:-1], *out.shape[1:])


def value_and_jac(f, x: torch.Tensor):
    """(f(x), df/dx) on [..., n] via forward-mode AD, like jax.jacfwd in
    parsmooth/linearization/_extended.py:59-60."""
    vj = getattr(f, "_psqrt_value_and_jac", None)
    if vj is not None:
        return vj(x)
    if x.dim() == 1:
        return f(x), torch.func.jacfwd(f)(x)
    flat = x.reshape(-1, x.shape[-1])
    val = torch.func.vmap(f)(flat)
    jac = torch.func.vmap(torch.func.jacfwd(f))(flat)
    return val.reshape(*x.shape[:-1], *val.shape[1:]), jac.reshape(*x.shape[:-1], *jac.shape[1:])


def mv(M, v):
    return torch.einsum("...ij,...j->...i", M, v)


class _PartialInX:
    """q_ -> f(x, q_) for a BATCH of states x [..., n], as the linearization methods want their functions:
    callable on points [..., n] or [..., P, n] (`_psqrt_batched`), with its own value-and-Jacobian in q_."""
    _psqrt_batched = True

    def __init__(self, f, x):
        self.f, self.x = f, x

    def _pairs(self, fn, qp):
        x = self.x
        extra = qp.dim() - x.dim()
        xe = x.reshape(x.shape[:-1] + (1,) * extra + x.shape[-1:]).expand(qp.shape[:-1] + x.shape[-1:])
        out = torch.func.vmap(fn)(xe.reshape(-1, x.shape[-1]), qp.reshape(-1, qp.shape[-1]))
        return out.reshape(*qp.shape[:-1], *out.shape[1:])

    def __call__(self, qp):
        return self._pairs(self.f, qp)

    def _psqrt_value_and_jac(self, qm):
        return self._pairs(self.f, qm), self._pairs(torch.func.jacfwd(self.f, argnums=1), qm)


def get_conditional_model(f, q, linearization_method):
    """parsmooth.linearization.get_conditional_model (linearization/_common.py:17-66), square-root branch:
    the conditional-moments model of a function f(x, q_) that is non-linear in the noise too -- for every x the
    map q_ -> f(x, q_) is linearised at the noise Gaussian with `linearization_method`; the conditional mean is
    F q.mean + bias and the conditional factor tria([F q.chol | chol]).

    f is a torch function of two 1-D tensors.  The returned functions take a batch of states [..., n] at once
    (`_psqrt_batched`), which is how the sigma-point methods evaluate them.  As upstream, x and q must have
    the same dimension (NotImplementedError otherwise).  An OUTER `extended` linearization differentiates the
    conditional mean with torch.func, which works when the inner method is `extended` too (a sigma-point inner
    method calls the CUDA triangularisation, which torch.func cannot trace)."""
    from .. import _lib
    require_sqrt(q)
    q = MVNSqrt(as_f64(q.mean), as_f64(q.chol))
    try:
        f(q.mean, q.mean)
    except Exception:  # noqa: BLE001  (upstream: bare except, _common.py:43-46)
        raise NotImplementedError("`x` and `q` with different dimensions are not supported yet.") from None
    n = q.mean.shape[-1]

    def _lin(x):
        x = as_f64(x)
        lead = x.shape[:-1]
        qm = q.mean.to(x.device).expand(*lead, n)
        qc = q.chol.to(x.device).expand(*lead, n, n)
        zero = MVNSqrt(torch.zeros(n, dtype=torch.float64, device=x.device),
                       torch.zeros(n, n, dtype=torch.float64, device=x.device))
        F, chol_val, bias = linearization_method(FunctionalModel(_PartialInX(f, x), zero), MVNSqrt(qm, qc))
        return F, chol_val, bias, qm, qc

    def conditional_mean(x):
        F, _, bias, qm, _ = _lin(x)
        return mv(F, qm) + bias

    def conditional_chol(x):
        F, chol_val, _, _, qc = _lin(x)
        cv = chol_val.expand(F.shape[:-2] + chol_val.shape[-2:])
        return _lib.tria(torch.cat([F @ qc, cv], -1))

    conditional_mean._psqrt_batched = True
    conditional_chol._psqrt_batched = True
    return ConditionalMomentsModel(conditional_mean, conditional_chol)
