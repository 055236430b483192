"""Statistical linear regression from sigma points, square-root form
(reference: parsmooth/linearization/_sigma_points.py:19-100).  The triangularisations and the
Cholesky downdates run in libpsqrt.so (psqrt_tria_batched / psqrt_chol_update_batched)."""
from __future__ import annotations

import torch

from .. import _lib
from .._base import are_inputs_compatible
from ._common import apply_fn, mv, require_sqrt


def _points(x, xi):
    """points[..., p, :] = m + chol @ xi[p]   (_cubature.py:56, _gh.py:68)"""
    m_x, chol_x = x
    return m_x[..., None, :] + torch.einsum("...ij,pj->...pi", chol_x, xi)


def _common(f, x, xi, wm, wc):
    """_sigma_points.py:88-100"""
    m_x, chol_x = x
    pts = _points(x, xi)
    f_pts = apply_fn(f, pts)
    m_f = torch.einsum("p,...pi->...i", wm, f_pts)
    Psi = torch.einsum("...pi,p,...pj->...ij", pts - m_x[..., None, :], wc, f_pts - m_f[..., None, :])
    # cho_solve((chol_x, True), Psi): lower triangle only, any diagonal sign
    F_x = torch.cholesky_solve(Psi, torch.tril(chol_x), upper=False).transpose(-1, -2)
    return F_x, pts, f_pts, m_f


def linearize_functional(f, x, q, xi, wm, wc):
    """_sigma_points.py:63-85 (sqrt branch)."""
    are_inputs_compatible(x, q)
    require_sqrt(x)
    require_sqrt(q)
    F_x, pts, f_pts, m_f = _common(f, x, xi, wm, wc)
    m_x, chol_x = x
    m_q, chol_q = q
    sqrt_Phi = torch.sqrt(wc)[:, None] * (f_pts - m_f[..., None, :])        # [..., P, d]
    n_pts, dim_out = sqrt_Phi.shape[-2:]
    sqrt_Phi = sqrt_Phi.transpose(-1, -2)
    if n_pts >= dim_out:
        sqrt_Phi = _lib.tria(sqrt_Phi)
    else:
        pad = sqrt_Phi.new_zeros(sqrt_Phi.shape[:-1] + (dim_out - n_pts,))
        sqrt_Phi = torch.cat([sqrt_Phi, pad], -1)
    cq = chol_q.expand(sqrt_Phi.shape[:-2] + chol_q.shape[-2:])
    chol_L = _lib.tria(torch.cat([sqrt_Phi, cq], -1))
    chol_L = _lib.chol_update_many(chol_L, (F_x @ chol_x).transpose(-1, -2), -1.0)
    return F_x, chol_L, m_f - mv(F_x, m_x) + m_q


def linearize_conditional(c_m, c_chol, x, xi, wm, wc):
    """_sigma_points.py:25-49 (sqrt branch)."""
    require_sqrt(x)
    m_x, chol_x = x
    F_x, pts, f_pts, m_f = _common(c_m, x, xi, wm, wc)
    sqrt_Phi = torch.sqrt(wc)[:, None] * (f_pts - m_f[..., None, :])
    sqrt_Phi = _lib.tria(sqrt_Phi.transpose(-1, -2))
    chol_pts = apply_fn(c_chol, pts)                                       # [..., P, d, d]
    temp = torch.sqrt(wc)[:, None, None] * chol_pts
    temp = temp.transpose(-3, -2)                                          # [..., d, P, d]
    temp = temp.reshape(*temp.shape[:-2], -1)                              # [..., d, P*d]
    chol_L = _lib.tria(torch.cat([sqrt_Phi, temp], -1))
    chol_L = _lib.chol_update_many(chol_L, (F_x @ chol_x).transpose(-1, -2), -1.0)
    return F_x, chol_L, m_f - mv(F_x, m_x)
