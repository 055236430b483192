"""Gradient path: the log-likelihood of an iterated smoother and its derivative with respect to model parameters.

Reference: ``jax.value_and_grad`` through ``parsmooth.methods.iterated_smoothing(..., return_loglikelihood=True)``
(methods.py:54-76), i.e. the implicit differentiation of the fixed point (``parsmooth/_utils.py:103-133``) followed by
the derivative of the filter's log-likelihood (``parallel/_filtering.py:53-60,149-154``); protocol
``notebooks/experiment_bearing_only_param_estimation_run_time.ipynb`` (``prec_r -> -ell``, L-BFGS-B on top).

The reference differentiates in reverse mode: with x* the nominal trajectory the loop stopped at, n_iter the number of
applications of f = filter_smoother(.; theta) it made, J = df/dx and B = df/dtheta at (x*, theta), its custom VJP
(_utils.py:118-133) returns   theta_bar = B^T sum_{j=0}^{n_iter+1} (J^T)^j x*_bar   (and a zero cotangent for the
initial trajectory), to which the direct derivative of the final log-likelihood is added.  Here the same number is
computed in FORWARD mode, one parameter direction at a time (the protocol's parameter is a scalar):

    xdot_0 = B thetadot,   xdot_{j+1} = B thetadot + J xdot_j   (n_iter + 1 times)  =>  xdot = sum_{j=0}^{n_iter+1} J^j B thetadot
    elldot = d ell/d theta . thetadot + d ell/d x* . xdot

Each J / B product is one tangent pass on the device at the SAME primal solution (psqrt_filter_smoother_tangent:
the tangent of a Kalman filter / RTS smoother is an affine recursion, evaluated as an associative scan of matrix
products, csrc/psqrt_tangent.cu), fed by the tangent of the linearization (psqrt_linearize_builtin_tangent for the
built-in models, torch.func.jvp for a user-supplied torch function with the extended method).
"""
from __future__ import annotations

import functools
from typing import Callable, NamedTuple, Optional, Sequence

import torch

from . import _lib, methods
from ._base import ConditionalMomentsModel, FunctionalModel, MVNSqrt

__all__ = ["Tangents", "TangentPass", "loglikelihood_jvp", "value_and_grad", "Cotangents", "loglikelihood_vjp",
           "value_and_grad_reverse"]


class Tangents(NamedTuple):
    """One direction of change of the inputs of a run; None entries are zero.  MVNSqrt entries hold the derivative
    of (mean, chol) -- factors, not covariances -- exactly like differentiating the reference's inputs."""
    x0: Optional[MVNSqrt] = None
    transition_noise: Optional[MVNSqrt] = None        # FunctionalModel.mvn of the transition model
    observation_noise: Optional[MVNSqrt] = None
    transition_params: Optional[Sequence[float]] = None   # parameters of a built-in model function (psqrt.models)
    observation_params: Optional[Sequence[float]] = None


def _cov_tangent(chol, dchol):
    """d(chol chol^T)"""
    if dchol is None:
        return None
    t = dchol @ chol.transpose(-1, -2)
    return t + t.transpose(-1, -2)


def _method_info(lin):
    kw = {}
    while isinstance(lin, functools.partial):
        kw = {**lin.keywords, **kw}
        lin = lin.func
    kind = getattr(lin, "_psqrt_kind", None)
    if kind is None:
        raise NotImplementedError("psqrt.grad: linearization_method must be one of psqrt.linearization."
                                  "{extended, cubature, gauss_hermite, unscented} (optionally functools.partial-ed)")
    return kind, (None if kind == "extended" else (lambda n: lin._psqrt_points(n, **kw)))


def _linearization_tangent(lin, model, x: MVNSqrt, dx_mean, dx_chol, dnoise: Optional[MVNSqrt], dparams, dev):
    """Tangent of ``lin(model, x)`` -> (dF, dQ, db) with dQ the COVARIANCE tangent of the returned factor; entries may
    be None (zero) or time-invariant."""
    kind, points = _method_info(lin)
    n = x.mean.shape[-1]
    if isinstance(model, FunctionalModel):
        f, q = model
        builtin = getattr(f, "_psqrt_builtin", None)
        dmq = _t(dnoise.mean, dev) if dnoise is not None and dnoise.mean is not None else None
        dQq = _cov_tangent(q.chol, _t(dnoise.chol, dev)) if dnoise is not None and dnoise.chol is not None else None
    elif isinstance(model, ConditionalMomentsModel):
        f, q = model[0], None
        builtin = getattr(f, "_psqrt_builtin", None)
        dmq = dQq = None
        if dnoise is not None:
            raise ValueError("a ConditionalMomentsModel has no noise tangent")
    else:
        raise TypeError(f"expected FunctionalModel or ConditionalMomentsModel, got {type(model)}")
    if builtin is not None and not hasattr(builtin, "model_id"):
        # linear function x -> M x (psqrt.models.lgssm): every method returns (M, chol_q, m_q), independent of the nominal
        if dparams is not None:
            raise NotImplementedError("a linear model function carries no differentiable parameters")
        return None, dQq, dmq
    if builtin is not None:
        conditional = bool(builtin.conditional)
        if kind == "extended":
            dF, dQ, db = _lib.linearize_builtin_tangent(builtin.model_id, builtin.params, dparams, _lib.LIN_EXTENDED,
                                                        builtin.n_in, builtin.n_out, conditional, x.mean,
                                                        dnom_m=dx_mean, dm_q=dmq)
            if not conditional:
                dQ = dQq
            return dF, dQ, db
        wm, wc, xi = points(n)
        import numpy as np
        pts = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in (xi, wm, wc))
        return _lib.linearize_builtin_tangent(builtin.model_id, builtin.params, dparams, _lib.LIN_SLR, builtin.n_in,
                                              builtin.n_out, conditional, x.mean, torch.tril(x.chol), dx_mean,
                                              dx_chol, dmq, dQq, points=pts)
    if dparams is not None:
        raise NotImplementedError("parameter tangents are available for the built-in models of psqrt.models only")
    if kind != "extended" or not isinstance(model, FunctionalModel):
        raise NotImplementedError("psqrt.grad: a user-supplied torch model is differentiated with the extended "
                                  "linearization only (the sigma-point rules run in CUDA kernels torch.func cannot trace)")
    # extended, torch function: F = jac f(m), b = f(m) - F m + m_q  (linearization/_extended.py:68-70)
    m = x.mean

    def lin1(v):
        J = torch.func.jacfwd(f)(v)
        return J, f(v) - J @ v

    if dx_mean is None:
        dF = db = None
    else:
        _, (dF, db) = torch.func.jvp(torch.func.vmap(lin1), (m,), (dx_mean,))
    if dmq is not None:
        db = dmq if db is None else db + dmq
    return dF, dQq, db


def _t(x, dev):
    return None if x is None else torch.as_tensor(x, dtype=torch.float64).to(dev)


class TangentPass:
    """The primal pass at a nominal trajectory (linearization, filter, smoother, log-likelihood) and tangent passes
    at that solution.  `jvp(dnominal)` is one application of  xdot -> J xdot + B thetadot."""

    def __init__(self, observations, x0, transition_model, observation_model, linearization_method, nominal,
                 tangents: Tangents):
        dev = methods._device()
        self.dev = dev
        self.lin = linearization_method
        self.ys = methods._t(observations, dev)
        self.x0 = methods._mvn(x0, dev)
        self.tm = methods._model(transition_model, dev)
        self.om = methods._model(observation_model, dev)
        self.nominal = methods._mvn(nominal, dev)
        self.tangents = tangents
        T = self.ys.shape[0]
        self.T = T
        self.ssm = methods._linearize(self.lin, self.tm, self.om, self.nominal)
        L0 = methods._prior_factor(self.x0.chol)
        fm, fL, sm, sL, ell = _lib.filter_smoother(self.ssm, self.ys, self.x0.mean, L0, smooth=True, loglik=True)
        self.fm, self.fL, self.sm, self.sL, self.ell = fm, fL, sm, sL, ell
        tx0 = tangents.x0
        self.dm0 = _t(tx0.mean, dev) if tx0 is not None and tx0.mean is not None else None
        self.dP0 = _cov_tangent(self.x0.chol, _t(tx0.chol, dev)) if tx0 is not None and tx0.chol is not None else None

    @property
    def smoothed(self):
        return MVNSqrt(self.sm, self.sL)

    def _model_tangent(self, dnominal):
        nom = self.nominal
        dm = dL = None
        if dnominal is not None:
            dm, dL = dnominal
        sl0, sl1 = slice(None, -1), slice(1, None)
        cut = lambda t, sl: None if t is None else t[sl].contiguous()
        tg = self.tangents
        dF, dQ, db = _linearization_tangent(self.lin, self.tm, MVNSqrt(nom.mean[sl0], nom.chol[sl0]), cut(dm, sl0),
                                            cut(dL, sl0), tg.transition_noise, tg.transition_params, self.dev)
        dH, dR, dc = _linearization_tangent(self.lin, self.om, MVNSqrt(nom.mean[sl1], nom.chol[sl1]), cut(dm, sl1),
                                            cut(dL, sl1), tg.observation_noise, tg.observation_params, self.dev)
        return {"dF": dF, "dQ": dQ, "db": db, "dH": dH, "dR": dR, "dc": dc}

    def tangent(self, dnominal=None, *, smooth=True, loglik=True):
        """-> (dfm, dfP, dsm, dsP, dell): covariance-form tangents of the filtered / smoothed moments and of ell along
        (self.tangents, dnominal); dnominal = (dmean [T+1,nx], dchol [T+1,nx,nx]) or None."""
        dssm = self._model_tangent(dnominal)
        return _lib.filter_smoother_tangent(self.ssm, dssm, self.ys, self.fm, self.fL, self.sm, self.sL, self.dm0,
                                            self.dP0, smooth=smooth, loglik=loglik)

    def jvp(self, dnominal=None):
        """One application of xdot -> d filter_smoother(x*; theta) [xdot, thetadot]: (dmean, dchol) of the smoothed
        trajectory, the factor tangent being that of the lower-triangular factor the pass returns."""
        _, _, dsm, dsP, _ = self.tangent(dnominal, smooth=True, loglik=False)
        dL = _lib.cov_tangent_to_chol(self.sL, dsP)
        # in the gauge of the nominal factor it perturbs next: column signs of a triangularisation are arbitrary, and
        # the sigma-point rules are invariant to flipping a column of (chol, dchol) together, not of one alone
        flip = torch.sign(torch.diagonal(self.sL, dim1=-2, dim2=-1)) * \
            torch.sign(torch.diagonal(self.nominal.chol, dim1=-2, dim2=-1))
        flip = torch.where(flip == 0, torch.ones_like(flip), flip)
        return dsm, dL * flip[..., None, :]

    def ell_jvp(self, dnominal=None):
        _, _, _, _, dell = self.tangent(dnominal, smooth=False, loglik=True)
        return dell


def _fixed_point_counted(f, x0, criterion):
    """parsmooth/_utils.py:136-146 with the iteration count the custom VJP keeps (118-120)."""
    i, x_prev, x = 1, x0, f(x0)
    while bool(criterion(i, x_prev, x)):
        i, x_prev, x = i + 1, x, f(x)
    return x, i


def loglikelihood_jvp(observations, x0, transition_model, observation_model, linearization_method: Callable,
                      tangents: Tangents, init_nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True,
                      criterion: Callable = methods._default_criterion, implicit_terms: Optional[int] = None):
    """(nominal*, ell, d ell) of ``iterated_smoothing(..., return_loglikelihood=True)`` along `tangents`.

    `implicit_terms`: number of Neumann terms of the implicit fixed-point derivative; default n_iter + 2, what the
    reference's custom VJP accumulates (_utils.py:122-125)."""
    methods._check_parallel(parallel)
    dev = methods._device()
    observations = methods._t(observations, dev)
    x0 = methods._mvn(x0, dev)
    transition_model = methods._model(transition_model, dev)
    observation_model = methods._model(observation_model, dev)
    if init_nominal_trajectory is None:
        init_nominal_trajectory = methods.filter_smoother(observations, x0, transition_model, observation_model,
                                                          linearization_method, None, parallel)

    def fun_to_iter(nominal):
        return methods.filter_smoother(observations, x0, transition_model, observation_model, linearization_method,
                                       nominal, parallel)

    nominal, n_iter = _fixed_point_counted(fun_to_iter, init_nominal_trajectory, criterion)
    tp = TangentPass(observations, x0, transition_model, observation_model, linearization_method, nominal, tangents)
    terms = n_iter + 2 if implicit_terms is None else int(implicit_terms)
    dx = None
    for _ in range(terms):
        dx = tp.jvp(dx)
    return nominal, tp.ell, tp.ell_jvp(dx)


def value_and_grad(build: Callable, theta, observations, linearization_method: Callable,
                   init_nominal_trajectory: Optional[MVNSqrt] = None, criterion: Callable = methods._default_criterion,
                   implicit_terms: Optional[int] = None):
    """(ell, d ell / d theta) for ``build(theta) -> (x0, transition_model, observation_model)``, the analogue of
    ``jax.value_and_grad(lambda theta: iterated_smoothing(...)[1])``.  `theta` is a 1-D fp64 tensor; `build` must
    compute the means / factors of x0 and of the two noise MVNSqrt with torch operations on `theta` (their tangents
    come from torch.func.jvp); the model functions themselves may not depend on theta."""
    dev = methods._device()
    theta = torch.as_tensor(theta, dtype=torch.float64).reshape(-1).to(dev)
    funcs = {}

    def tensors(th):
        x0, tm, om = build(th)
        funcs["tm"], funcs["om"] = tm, om
        out = [x0.mean, x0.chol]
        for m in (tm, om):
            if isinstance(m, FunctionalModel):
                out += [m.mvn.mean, m.mvn.chol]
        return tuple(torch.as_tensor(v, dtype=torch.float64).to(dev) for v in out)

    grad = torch.zeros_like(theta)
    ell = None
    for i in range(theta.numel()):
        e = torch.zeros_like(theta)
        e[i] = 1.0
        prim, tang = torch.func.jvp(tensors, (theta,), (e,))
        prim = [p.detach() for p in prim]
        tang = [t.detach() for t in tang]
        tm, om = funcs["tm"], funcs["om"]
        x0 = MVNSqrt(prim[0], prim[1])
        k = 2
        tn = on = None
        if isinstance(tm, FunctionalModel):
            tm = FunctionalModel(tm.function, MVNSqrt(prim[k], prim[k + 1]))
            tn = MVNSqrt(tang[k], tang[k + 1])
            k += 2
        if isinstance(om, FunctionalModel):
            om = FunctionalModel(om.function, MVNSqrt(prim[k], prim[k + 1]))
            on = MVNSqrt(tang[k], tang[k + 1])
        tg = Tangents(x0=MVNSqrt(tang[0], tang[1]), transition_noise=tn, observation_noise=on)
        _, ell, dell = loglikelihood_jvp(observations, x0, tm, om, linearization_method, tg, init_nominal_trajectory,
                                         True, criterion, implicit_terms)
        grad[i] = dell
    return ell, grad


# ---- reverse mode: d ell / d (everything) at a fixed linearised model in ONE adjoint pass --------------------------
class Cotangents(NamedTuple):
    """Gradient of the log-likelihood of one filtering pass with respect to its inputs (what ``jax.grad`` returns for
    the same arguments of ``parsmooth.methods.filtering(..., return_loglikelihood=True)`` at a FIXED nominal
    trajectory).  `x0`, `transition_noise`, `observation_noise`: MVNSqrt of gradients w.r.t. (mean, chol) -- factors,
    like differentiating the reference's inputs (noise entries None for a ConditionalMomentsModel).  `model`: the
    per-step gradients {"gF","gQ","gb","gH","gR","gc"} w.r.t. the linearised model (Q, R in covariance form) and the
    costates {"lam","Lam"} = d ell / d (filtered mean, covariance) of every step."""
    x0: MVNSqrt
    transition_noise: Optional[MVNSqrt]
    observation_noise: Optional[MVNSqrt]
    model: dict


def loglikelihood_vjp(observations, x0, transition_model, observation_model, linearization_method: Callable,
                      nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True):
    """(filtered, ell, Cotangents): ``filtering(..., return_loglikelihood=True)`` (methods.py:14-24) and the gradient
    of ell w.r.t. the prior, the additive noises and every entry of the linearised model of every step, by ONE adjoint
    pass (psqrt_loglik_adjoint: the costate recursion of the Kalman filter as a reverse associative scan).  The cost
    does not depend on the number of parameters; for a linear-Gaussian model (psqrt.models.lgssm) this is the complete
    gradient.  For a nonlinear model it is the derivative at a fixed nominal trajectory -- the term ``d ell/d theta``
    of the module docstring; the dependence through the nominal trajectory stays in forward mode
    (loglikelihood_jvp)."""
    methods._check_parallel(parallel)
    dev = methods._device()
    ys = methods._t(observations, dev)
    x0 = methods._mvn(x0, dev)
    tm = methods._model(transition_model, dev)
    om = methods._model(observation_model, dev)
    T = ys.shape[0]
    if nominal_trajectory is None:
        nominal_trajectory = methods._default_nominal(T + 1, x0.mean.shape[-1], dev)
    nominal = methods._mvn(nominal_trajectory, dev)
    ssm = methods._linearize(linearization_method, tm, om, nominal)
    L0 = methods._prior_factor(x0.chol)
    fm, fL, _, _, ell = _lib.filter_smoother(ssm, ys, x0.mean, L0, smooth=False, loglik=True)
    g = _lib.loglik_adjoint(ssm, ys, fm, fL)

    def noise(model, gm, gcov):
        if not isinstance(model, FunctionalModel):
            return None
        chol = model.mvn.chol
        S = gcov.sum(0)
        return MVNSqrt(gm.sum(0), (S + S.transpose(-1, -2)) @ chol)

    Lam0 = g["Lam"][0]
    gx0 = MVNSqrt(g["lam"][0], (Lam0 + Lam0.transpose(-1, -2)) @ x0.chol)
    return MVNSqrt(fm, fL), ell, Cotangents(gx0, noise(tm, g["gb"], g["gQ"]), noise(om, g["gc"], g["gR"]), g)


def value_and_grad_reverse(build: Callable, theta, observations, linearization_method: Callable,
                           nominal_trajectory: Optional[MVNSqrt] = None):
    """(ell, d ell / d theta) for ``build(theta) -> (x0, transition_model, observation_model)`` at a fixed nominal
    trajectory, ALL components of theta from one adjoint pass: the cotangents of loglikelihood_vjp are pulled back
    through `build` with torch.autograd (means / factors of x0 and of the two noises must be torch functions of
    theta; the model functions themselves may not depend on it).  The analogue of
    ``jax.value_and_grad(lambda theta: filtering(ys, *build(theta), lin, nominal, True, True)[1])``."""
    dev = methods._device()
    theta = torch.as_tensor(theta, dtype=torch.float64).reshape(-1).to(dev).detach().requires_grad_(True)
    with torch.enable_grad():
        x0, tm, om = build(theta)
        leaves = [x0.mean, x0.chol]
        for m in (tm, om):
            if isinstance(m, FunctionalModel):
                leaves += [m.mvn.mean, m.mvn.chol]
        leaves = [torch.as_tensor(v, dtype=torch.float64).to(dev) for v in leaves]
    det = [v.detach() for v in leaves]
    k = 2
    tm_d, om_d = tm, om
    if isinstance(tm, FunctionalModel):
        tm_d = FunctionalModel(tm.function, MVNSqrt(det[k], det[k + 1]))
        k += 2
    if isinstance(om, FunctionalModel):
        om_d = FunctionalModel(om.function, MVNSqrt(det[k], det[k + 1]))
    _, ell, ct = loglikelihood_vjp(observations, MVNSqrt(det[0], det[1]), tm_d, om_d, linearization_method,
                                   nominal_trajectory)
    cots = [ct.x0.mean, ct.x0.chol]
    for c in (ct.transition_noise, ct.observation_noise):
        if c is not None:
            cots += [c.mean, c.chol]
    pairs = [(v, c) for v, c in zip(leaves, cots) if v.requires_grad]
    if not pairs:
        return ell, torch.zeros_like(theta)
    (grad,) = torch.autograd.grad([v for v, _ in pairs], [theta], [c.reshape(v.shape) for v, c in pairs],
                                  allow_unused=True)
    return ell, (torch.zeros_like(theta) if grad is None else grad).detach()
