"""float32 mode of the parallel square-root smoother (notebooks/robustness_100runs.py:7,41-77 run the
square-root and the covariance-form smoothers in float32 and count the runs that end in NaN).

The scan -- elements, both associative operators, every triangularisation -- runs in float32 on the device through the
generic path (psqrt_filter_smoother_f32, csrc/psqrt_generic.cu).  The model is linearised by the ordinary fp64 code at
the (float32-rounded) nominal trajectory and rounded to float32 before the pass, so what is exercised is the
robustness of the square-root recursion itself, which is the point of the reference's experiment.  Same argument order
as psqrt.methods; trajectories come back as float32 MVNSqrt, the log-likelihood as a float64 scalar.
`psqrt.dist.nonfinite_runs` / `psqrt._lib.count_nonfinite` give the NaN count."""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib, methods
from ._base import MVNSqrt, are_inputs_compatible

__all__ = ["filter_smoother", "iterated_smoothing"]


def _pass(ys, x0, transition_model, observation_model, lin, nominal, loglik):
    ssm = methods._linearize(lin, transition_model, observation_model, nominal)      # fp64, unfused
    fm, fL, sm, sL, ell = _lib.filter_smoother_f32(ssm.F, ssm.cholQ, ssm.b, ssm.H, ssm.cholR, ssm.c, ys, x0.mean,
                                                   x0.chol, smooth=True, loglik=loglik)
    return MVNSqrt(fm, fL), MVNSqrt(sm, sL), ell


def _prepare(observations, x0, transition_model, observation_model, nominal):
    dev = methods._device()
    ys = methods._t(observations, dev)
    x0 = methods._mvn(x0, dev)
    tm = methods._model(transition_model, dev)
    om = methods._model(observation_model, dev)
    T, nx = ys.shape[0], x0.mean.shape[-1]
    if nominal is not None:
        are_inputs_compatible(x0, nominal)
        nominal = methods._mvn(MVNSqrt(nominal.mean.double() if torch.is_tensor(nominal.mean) else nominal.mean,
                                       nominal.chol.double() if torch.is_tensor(nominal.chol) else nominal.chol), dev)
    else:
        nominal = methods._default_nominal(T + 1, nx, dev)
    return ys, x0, tm, om, nominal


def filter_smoother(observations, x0, transition_model, observation_model, linearization_method: Callable,
                    nominal_trajectory: Optional[MVNSqrt] = None, return_loglikelihood: bool = False):
    """psqrt.methods.filter_smoother in float32 (optionally with the log-likelihood of the filtering pass)."""
    ys, x0, tm, om, nominal = _prepare(observations, x0, transition_model, observation_model, nominal_trajectory)
    _, smoothed, ell = _pass(ys, x0, tm, om, linearization_method, nominal, return_loglikelihood)
    return (smoothed, ell) if return_loglikelihood else smoothed


def iterated_smoothing(observations, x0, transition_model, observation_model, linearization_method: Callable,
                       init_nominal_trajectory: Optional[MVNSqrt] = None,
                       criterion: Callable = methods._default_criterion, return_loglikelihood: bool = False):
    """psqrt.methods.iterated_smoothing (methods.py:54-76) with every pass in float32."""
    ys, x0, tm, om, nominal = _prepare(observations, x0, transition_model, observation_model, init_nominal_trajectory)
    up = lambda t: MVNSqrt(t.mean.double(), t.chol.double())
    if init_nominal_trajectory is None:
        nominal = up(_pass(ys, x0, tm, om, linearization_method, nominal, False)[1])

    def f(curr):
        return up(_pass(ys, x0, tm, om, linearization_method, curr, False)[1])

    nominal = methods.fixed_point(f, nominal, criterion)
    out = MVNSqrt(nominal.mean.float(), nominal.chol.float())
    if return_loglikelihood:
        _, _, ell = _pass(ys, x0, tm, om, linearization_method, nominal, True)
        return out, ell
    return out
