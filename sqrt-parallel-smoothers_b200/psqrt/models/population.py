"""Stochastic Ricker map with Poisson observations as a conditional-moments model
(reference: notebooks/population_model.py:23-34,51-62,84-129)."""
from __future__ import annotations

import numpy as np
import torch

from .._base import ConditionalMomentsModel

_LOG_447 = float(np.log(44.7))


def make_parameters(lam: float, Q):
    """-> (transition ConditionalMomentsModel, observation ConditionalMomentsModel), sqrt form:
    E[x'|x] = log 44.7 + x - exp x, chol = sqrt(Q);  E[y|x] = lam exp x, chol = sqrt(lam exp x)."""
    Q_t = torch.as_tensor(np.asarray(Q), dtype=torch.float64).reshape(1, 1)

    def mean_t(x):
        return _LOG_447 + x - torch.exp(x)

    mean_t._psqrt_batched = True
    mean_t._psqrt_value_and_jac = lambda x: (mean_t(x), (1.0 - torch.exp(x))[..., None])

    def chol_t(x):
        return torch.sqrt(Q_t).to(x.device).expand(*x.shape[:-1], 1, 1)

    chol_t._psqrt_batched = True

    def mean_o(x):
        return lam * torch.exp(x)

    mean_o._psqrt_batched = True
    mean_o._psqrt_value_and_jac = lambda x: (mean_o(x), (lam * torch.exp(x))[..., None])

    def chol_o(x):
        return torch.sqrt(lam * torch.exp(x))[..., None]

    chol_o._psqrt_batched = True
    from .. import _lib
    from ._builtin import Builtin
    mean_t._psqrt_builtin = Builtin(_lib.MODEL_RICKER_TRANSITION, [float(torch.sqrt(Q_t).item())], 1, 1, conditional=True)
    mean_o._psqrt_builtin = Builtin(_lib.MODEL_POISSON_OBSERVATION, [float(lam)], 1, 1, conditional=True)
    return ConditionalMomentsModel(mean_t, chol_t), ConditionalMomentsModel(mean_o, chol_o)


def get_data(x0, T, Q, lam, random_state=None):
    """Simulate the Ricker/Poisson model (population_model.py:170-208; NumPy RNG instead of jax.random)."""
    if random_state is None or isinstance(random_state, int):
        random_state = np.random.RandomState(random_state)
    x = float(np.asarray(x0).reshape(-1)[0])
    sq = float(np.sqrt(np.asarray(Q).reshape(-1)[0]))
    xs = np.empty((T + 1, 1))
    ys = np.empty((T, 1))
    xs[0] = x
    for k in range(T):
        x = _LOG_447 + x - np.exp(x) + sq * random_state.randn()
        xs[k + 1] = x
        ys[k] = random_state.poisson(lam * np.exp(x))
    return xs, ys
