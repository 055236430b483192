"""NumPy stand-in for the five staged kernels (TEST INFRASTRUCTURE, built on the oracle): lets the
host-side sharding logic of psqrt/dist.py run on CPU tensors with the gloo backend."""
import numpy as np
import torch

import parsmooth_np as O


def _np(t):
    return t.detach().cpu().numpy()


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)


def _lin(ssm, T, b):
    out = []
    for name, core in (("F", 2), ("cholQ", 2), ("b", 1), ("H", 2), ("cholR", 2), ("c", 1)):
        t = getattr(ssm, name)
        if t is None:
            out.append(None)
            continue
        a = _np(t)
        lead = a.ndim - core
        if lead == 2:
            a = a[b]
        if a.ndim == core:
            a = np.broadcast_to(a, (T,) + a.shape)
        out.append(a)
    return out


def _pack_f(e):
    return np.concatenate([np.asarray(x).reshape(-1) for x in e])


def _unpack_f(v, n):
    o, out = 0, []
    for sz, shp in ((n * n, (n, n)), (n, (n,)), (n * n, (n, n)), (n, (n,)), (n * n, (n, n))):
        out.append(v[o:o + sz].reshape(shp))
        o += sz
    return tuple(out)


def _unpack_s(v, n):
    return v[:n], v[n:n + n * n].reshape(n, n), v[n + n * n:].reshape(n, n)


def tria(L):
    return _t(O.tria(_np(L)))


def filter_reduce(ssm, y, nx, chunk_len=0):
    B, T, ny = y.shape
    outs = []
    for b in range(B):
        lin = _lin(ssm, T, b)
        el = O.sqrt_filtering_elements(*lin, np.zeros((T, nx)), np.zeros((T, nx, nx)), _np(y[b]))
        pref = O.sequential_fold_scan(O.sqrt_filtering_operator, el)
        outs.append(_pack_f([p[-1] for p in pref]))
    return _t(np.stack(outs))


def carry_filter(totals, rank, m0, L0):
    B, nx = m0.shape
    cm, cL = _np(m0).copy(), _np(L0).copy()
    for b in range(B):
        e1 = (np.zeros((1, nx, nx)), cm[b][None], cL[b][None], np.zeros((1, nx)), np.zeros((1, nx, nx)))
        for r in range(rank):
            e2 = tuple(x[None] for x in _unpack_f(_np(totals[r, b]), nx))
            res = O.sqrt_filtering_operator(e1, e2)
            e1 = (np.zeros((1, nx, nx)), res[1], res[2], np.zeros((1, nx)), np.zeros((1, nx, nx)))
        cm[b], cL[b] = e1[1][0], e1[2][0]
    return _t(cm), _t(cL)


def filter_apply(ssm, y, carry_m, carry_L, smooth=True, loglik=False, chunk_len=0):
    B, T, ny = y.shape
    nx = carry_m.shape[-1]
    fm, fL = np.zeros((B, T + 1, nx)), np.zeros((B, T + 1, nx, nx))
    ell, stot = np.zeros(B), []
    for b in range(B):
        lin = _lin(ssm, T, b)
        ms = np.concatenate([_np(carry_m[b])[None], np.zeros((T - 1, nx))])
        Ls = np.concatenate([_np(carry_L[b])[None], np.zeros((T - 1, nx, nx))])
        el = O.sqrt_filtering_elements(*lin, ms, Ls, _np(y[b]))
        _, m, c, _, _ = O.associative_scan(O.sqrt_filtering_operator, el)
        fm[b] = np.concatenate([_np(carry_m[b])[None], m])
        fL[b] = np.concatenate([_np(carry_L[b])[None], c])
        ell[b] = np.sum(O.sqrt_loglikelihood_terms(*lin, fm[b, :-1], fL[b, :-1], _np(y[b])))
        if smooth:
            se = O.sqrt_smoothing_elements(lin[0], lin[1], lin[2], fm[b, :-1], fL[b, :-1])
            suf = O.sequential_fold_scan(O.sqrt_smoothing_operator, se, reverse=True)
            stot.append(np.concatenate([suf[0][0].reshape(-1), suf[1][0].reshape(-1), suf[2][0].reshape(-1)]))
    return _t(fm), _t(fL), (_t(ell) if loglik else None), (_t(np.stack(stot)) if smooth else None)


def carry_smoother(stotals, rank, n_ranks, mT, LT):
    B, nx = mT.shape
    cm, cL = _np(mT).copy(), _np(LT).copy()
    for b in range(B):
        for r in range(n_ranks - 1, rank, -1):
            g, E, D = _unpack_s(_np(stotals[r, b]), nx)
            cm[b], cL[b] = E @ cm[b] + g, O.tria(np.concatenate([E @ cL[b], D], 1))
    return _t(cm), _t(cL)


def smoother_apply(ssm, fm, fL, carry_m, carry_L, write_terminal=True, chunk_len=0):
    B, T1, nx = fm.shape
    T = T1 - 1
    sm, sL = np.zeros((B, T1, nx)), np.zeros((B, T1, nx, nx))
    for b in range(B):
        lin = _lin(ssm, T, b)
        g, E, D = O.sqrt_smoothing_elements(lin[0], lin[1], lin[2], _np(fm[b, :-1]), _np(fL[b, :-1]))
        g = np.concatenate([g, _np(carry_m[b])[None]])
        E = np.concatenate([E, np.zeros((1, nx, nx))])
        D = np.concatenate([D, _np(carry_L[b])[None]])
        m, _, c = O.associative_scan(O.sqrt_smoothing_operator, (g, E, D), reverse=True)
        sm[b], sL[b] = m, c
    return _t(sm), _t(sL)
