"""Shared helpers of the linearization methods (host side, torch)."""
from __future__ import annotations

import torch

from .._base import MVNSqrt, MVNStandard


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
