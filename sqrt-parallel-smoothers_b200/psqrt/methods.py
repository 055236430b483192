"""User-facing API: a drop-in for parsmooth.methods on the parallel square-root path
(reference: parsmooth/methods.py:14-76 -> parallel/_filtering.py:13-61, parallel/_smoothing.py:14-57).

Same names, positional order and return conventions as the reference.  What differs:
* arrays are fp64 torch CUDA tensors (NumPy / CPU inputs are moved to the current CUDA device);
* ``parallel=False`` and ``MVNStandard`` inputs raise NotImplementedError -- this package is the
  sqrt=True, parallel=True path only and has no CPU or covariance-form fallback;
* the scan is the chunked three-sweep scheme of libpsqrt.so instead of jax.lax.associative_scan.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib
from ._base import MVNSqrt, MVNStandard, FunctionalModel, ConditionalMomentsModel, are_inputs_compatible
from ._lib import LinearizedSSM

__all__ = ["filtering", "smoothing", "filter_smoother", "iterated_smoothing", "sampling"]


def _device():
    if not torch.cuda.is_available():
        raise _lib.PsqrtError("psqrt needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _t(x, dev):
    """-> fp64 tensor on `dev`.  Inputs that live on the host keep a NumPy mirror (`_psqrt_host`) so that
    small time-invariant model matrices can travel to the kernels by value (psqrt._lib.LinearizedSSM)."""
    if torch.is_tensor(x) and x.is_cuda:
        return x.to(dev, dtype=torch.float64)
    host = np.asarray(x.detach().cpu().numpy() if torch.is_tensor(x) else x, dtype=np.float64)
    t = torch.as_tensor(host).to(dev)
    if host.size <= 4096:
        t._psqrt_host = np.array(host, dtype=np.float64, copy=True)   # private copy: the caller may reuse its array
        t._psqrt_host_version = t._version                            # stale after any in-place update of t
    return t


def _mvn(x, dev):
    if isinstance(x, MVNStandard):
        raise NotImplementedError("psqrt implements the square-root path only (MVNSqrt); "
                                  "covariance-form inputs are out of scope")
    if not isinstance(x, MVNSqrt):
        raise TypeError(f"expected MVNSqrt, got {type(x)}")
    return MVNSqrt(_t(x.mean, dev), _t(x.chol, dev))


def _model(model, dev):
    if isinstance(model, FunctionalModel):
        return FunctionalModel(model.function, _mvn(model.mvn, dev))
    if isinstance(model, ConditionalMomentsModel):
        return model
    raise TypeError(f"expected FunctionalModel or ConditionalMomentsModel, got {type(model)}")


def _check_parallel(parallel):
    if not parallel:
        raise NotImplementedError("psqrt is the parallel=True path; the sequential algorithms are not part of it")


def _default_nominal(T1, nx, dev):
    """zeros mean, identity chol, length T+1 (parallel/_filtering.py:25-28, _smoothing.py:22-28);
    the identity is a stride-0 view, nothing of size T is allocated for it."""
    mean = torch.zeros((T1, nx), dtype=torch.float64, device=dev)
    chol = torch.eye(nx, dtype=torch.float64, device=dev).expand(T1, nx, nx)
    return MVNSqrt(mean, chol)


def _slice(nominal, sl):
    return MVNSqrt(nominal.mean[sl], nominal.chol[sl])


def _lower(chol):
    """The kernels read only the lower triangle of cholQ: any other square root of the process noise is
    triangularised first (tria, parsmooth/_utils.py:22-24; L L^T is unchanged)."""
    if getattr(chol, "_psqrt_lower", False):
        return chol
    host = getattr(chol, "_psqrt_host", None)
    if host is not None and getattr(chol, "_psqrt_host_version", None) != chol._version:
        host = None
    if host is not None:
        if not np.any(np.triu(host, 1)):
            return chol
    elif chol.dim() == 2 and not bool((torch.triu(chol, 1) != 0).any()):
        return chol
    out = _lib.tria(chol)
    out._psqrt_lower = True
    return out


_FUSE_LIN_DEFAULT = "1"   # bearings T = 1e5 x 10 iterations: 2.8-3.2 ms fused, 3.45-3.6 ms unfused


def _host_mirror(t):
    """The private host copy _t attached to a tensor it built from host data, if still valid."""
    h = getattr(t, "_psqrt_host", None)
    if h is None or getattr(t, "_psqrt_host_version", None) != t._version:
        return None
    return h


def _fused_ssm(lin, transition_model, observation_model, nominal):
    """The built-in bearings-only model with the extended linearization is linearised INSIDE the sweeps
    (csrc/psqrt_fused.cuh): no per-step (F, b, H, c) arrays are formed.  Needs the time-invariant noise on the host
    (mirrors kept by _t for inputs that came from host data) and a lower-triangular cholQ; None = not applicable.
    PSQRT_FUSE_LIN=0 turns it off."""
    if os.environ.get("PSQRT_FUSE_LIN", _FUSE_LIN_DEFAULT) == "0" or getattr(lin, "_psqrt_kind", None) != "extended":
        return None
    if not (isinstance(transition_model, FunctionalModel) and isinstance(observation_model, FunctionalModel)):
        return None
    bt = getattr(transition_model.function, "_psqrt_builtin", None)
    bo = getattr(observation_model.function, "_psqrt_builtin", None)
    if getattr(bt, "model_id", None) != _lib.MODEL_CT_TRANSITION or \
            getattr(bo, "model_id", None) != _lib.MODEL_BEARINGS_OBSERVATION:
        return None
    if not nominal.mean.is_cuda or nominal.mean.dim() != 2:
        return None
    hq = [_host_mirror(t) for t in (transition_model.mvn.chol, transition_model.mvn.mean, observation_model.mvn.chol,
                                    observation_model.mvn.mean)]
    if any(h is None for h in hq) or np.any(np.triu(hq[0], 1)):
        return None
    return LinearizedSSM.fused_ct_bearings(nominal.mean, list(bt.params) + list(bo.params), hq[0], hq[1], hq[2], hq[3])


def _linearize(lin, transition_model, observation_model, nominal, fused_ok=False):
    """transition at nominal[:-1], observation at nominal[1:] (parallel/_filtering.py:103-104,117-119)."""
    if fused_ok and observation_model is not None:
        ssm = _fused_ssm(lin, transition_model, observation_model, nominal)
        if ssm is not None:
            return ssm
    F, cholQ, b = lin(transition_model, _slice(nominal, slice(None, -1)))
    cholQ = _lower(cholQ)
    if observation_model is None:
        return LinearizedSSM(F, cholQ, b)
    H, cholR, c = lin(observation_model, _slice(nominal, slice(1, None)))
    return LinearizedSSM(F, cholQ, b, H, cholR, c)


def _prior_factor(L0):
    """The kernels read the lower triangle of the carry-in factor; any other square root of the
    prior covariance is first triangularised (tria, parsmooth/_utils.py:22-24).  Memoised on the tensor
    (with its version counter): an iterated smoother passes the same x0 every iteration."""
    memo = getattr(L0, "_psqrt_prior", None)
    if memo is not None and memo[0] == L0._version:
        return memo[1]
    out = _lib.tria(L0)
    try:
        L0._psqrt_prior = (L0._version, out)
    except AttributeError:
        pass
    return out


def _run(observations, x0, transition_model, observation_model, lin, nominal, smooth, loglik):
    dev = _device()
    ys = _t(observations, dev)
    x0 = _mvn(x0, dev)
    transition_model = _model(transition_model, dev)
    observation_model = _model(observation_model, dev)
    T = ys.shape[0]
    nx = x0.mean.shape[-1]
    if nominal is not None:
        are_inputs_compatible(x0, nominal)
        nominal = _mvn(nominal, dev)
    else:
        nominal = _default_nominal(T + 1, nx, dev)
    ssm = _linearize(lin, transition_model, observation_model, nominal, fused_ok=True)
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, ys, x0.mean, _prior_factor(x0.chol), smooth=smooth,
                                               loglik=loglik)
    # index 0 of the filtered trajectory is x0 itself (parallel/_filtering.py:45-46)
    fm[0].copy_(x0.mean)
    fL[0].copy_(x0.chol)
    return MVNSqrt(fm, fL), (MVNSqrt(sm, sL) if smooth else None), ell


def filtering(observations, x0, transition_model, observation_model, linearization_method: Callable,
              nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True,
              return_loglikelihood: bool = False):
    """parsmooth.methods.filtering (methods.py:14-26) for MVNSqrt x0, parallel=True."""
    _check_parallel(parallel)
    filt, _, ell = _run(observations, x0, transition_model, observation_model, linearization_method,
                        nominal_trajectory, smooth=False, loglik=return_loglikelihood)
    if return_loglikelihood:
        return filt, ell
    return filt


def smoothing(transition_model, filter_trajectory, linearization_method: Callable,
              nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True):
    """parsmooth.methods.smoothing (methods.py:29-35; parallel/_smoothing.py:14-57)."""
    _check_parallel(parallel)
    dev = _device()
    ft = _mvn(filter_trajectory, dev)
    transition_model = _model(transition_model, dev)
    T1, nx = ft.mean.shape
    if nominal_trajectory is not None:
        are_inputs_compatible(filter_trajectory, nominal_trajectory)
        nominal = _mvn(nominal_trajectory, dev)
    else:
        nominal = _default_nominal(T1, nx, dev)
    ssm = _linearize(linearization_method, transition_model, None, nominal)
    fL = ft.chol
    # kernels read lower triangles only; factors this library wrote are tagged (no device round trip to find out)
    if not getattr(fL, "_psqrt_lower", False) and bool((torch.triu(fL, 1) != 0).any()):
        fL = _lib.tria(fL)
    sm, sL = _lib.smoother(ssm, ft.mean, fL)
    return MVNSqrt(sm, sL)


def filter_smoother(observations, x0, transition_model, observation_model, linearization_method: Callable,
                    nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True):
    """parsmooth.methods.filter_smoother (methods.py:38-47), fused into one pass: the transition
    model is linearised once and the smoothing elements come out of the filter's own
    triangularisation."""
    _check_parallel(parallel)
    _, smoothed, _ = _run(observations, x0, transition_model, observation_model, linearization_method,
                          nominal_trajectory, smooth=True, loglik=False)
    return smoothed


def _default_criterion(_i, nominal_traj_prev, curr_nominal_traj):
    """methods.py:50-51"""
    return torch.mean((nominal_traj_prev.mean - curr_nominal_traj.mean) ** 2) > 1e-6


def fixed_point(f, x0, criterion):
    """parsmooth/_utils.py:103-105,136-146: carry (1, x0, f(x0)); iterate while criterion(i, prev, x)."""
    i, x_prev, x = 1, x0, f(x0)
    while bool(criterion(i, x_prev, x)):
        i, x_prev, x = i + 1, x, f(x)
    return x


def sampling(key, n_samples: int, transition_model, filter_trajectory, linearization_method: Callable,
             nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True):
    """parsmooth.methods.sampling (methods.py:79-91; _pathwise_sampler.py:13-38, 61-81, 109-124): joint samples
    [T + 1, n_samples, nx] of the smoothing distribution given the filtered trajectory.

    `key`: an int seed or a torch.Generator for the standard normal draws (upstream: a jax PRNGKey), or the draws
    themselves as an array [T + 1, n_samples, nx] (draw 0 makes the last state, draw t + 1 the increment of step
    t, as upstream).  The transition model is linearised at `nominal_trajectory` (default: the smoothed
    trajectory, methods.py:86-87).  Each step's factor enters with a non-negative diagonal, so a given draw maps
    to the same sample whatever sign convention the triangularisation used."""
    _check_parallel(parallel)
    dev = _device()
    ft = _mvn(filter_trajectory, dev)
    transition_model = _model(transition_model, dev)
    T1, nx = ft.mean.shape
    if nominal_trajectory is None:
        nominal_trajectory = smoothing(transition_model, ft, linearization_method, None, parallel)
    are_inputs_compatible(filter_trajectory, nominal_trajectory)
    nominal = _mvn(nominal_trajectory, dev)
    ssm = _linearize(linearization_method, transition_model, None, nominal)
    fL = ft.chol
    # kernels read lower triangles only; factors this library wrote are tagged (no device round trip to find out)
    if not getattr(fL, "_psqrt_lower", False) and bool((torch.triu(fL, 1) != 0).any()):
        fL = _lib.tria(fL)
    shape = (T1, int(n_samples), nx)
    if isinstance(key, (np.ndarray, torch.Tensor)) and tuple(key.shape) == shape:
        eps = _t(key, dev)
    else:
        gen = key
        if not isinstance(gen, torch.Generator):
            gen = torch.Generator(device=dev)
            gen.manual_seed(int(np.asarray(key).ravel()[-1]))
        eps = torch.randn(shape, dtype=torch.float64, device=dev, generator=gen)
    g, E, D = _lib.smoother_elements(ssm, ft.mean, fL)
    return _lib.sample_paths(g, E, D, eps)


def iterated_smoothing(observations, x0, transition_model, observation_model, linearization_method: Callable,
                       init_nominal_trajectory: Optional[MVNSqrt] = None, parallel: bool = True,
                       criterion: Callable = _default_criterion, return_loglikelihood: bool = False):
    """parsmooth.methods.iterated_smoothing (methods.py:54-76)."""
    _check_parallel(parallel)
    # host inputs move to the device once, not once per iteration (each conversion is a blocking H2D copy)
    dev = _device()
    observations = _t(observations, dev)
    x0 = _mvn(x0, dev)
    transition_model = _model(transition_model, dev)
    observation_model = _model(observation_model, dev)
    if init_nominal_trajectory is None:
        init_nominal_trajectory = filter_smoother(observations, x0, transition_model, observation_model,
                                                  linearization_method, None, parallel)

    def fun_to_iter(curr_nominal_traj):
        return filter_smoother(observations, x0, transition_model, observation_model, linearization_method,
                               curr_nominal_traj, parallel)

    nominal_traj = fixed_point(fun_to_iter, init_nominal_trajectory, criterion)
    if return_loglikelihood:
        _, ell = filtering(observations, x0, transition_model, observation_model, linearization_method,
                           nominal_traj, parallel, return_loglikelihood=True)
        return nominal_traj, ell
    return nominal_traj
