"""Linear-Gaussian state-space model x -> M x (reference: tests/_lgssm.py:5-39)."""
from __future__ import annotations

import numpy as np
import torch

from .._base import MVNSqrt


class _LinearBuiltin:
    """Recognised by the linearization methods: every method returns (M, chol_q, m_q) for a linear
    function (tests/test_linearization.py:71-107), and it is time-invariant, so nothing of size T
    is materialised."""

    def __init__(self, M):
        self.M = M

    def _M(self, like):
        t = torch.as_tensor(self.M, dtype=torch.float64).to(like.device)
        t._psqrt_host = np.array(self.M.detach().cpu().numpy(), dtype=np.float64, copy=True)   # host mirror (by-value path)
        t._psqrt_host_version = t._version
        return t

    def extended(self, x: MVNSqrt, q: MVNSqrt):
        return self._M(x.mean), q.chol, q.mean


def linear_function(M):
    """f(x) = M x as a recognised built-in (works on [..., n])."""
    M_t = torch.as_tensor(np.asarray(M) if not torch.is_tensor(M) else M, dtype=torch.float64)

    def f(x):
        return torch.einsum("ij,...j->...i", M_t.to(x.device), x)

    f._psqrt_batched = True
    f._psqrt_value_and_jac = lambda x: (f(x), M_t.to(x.device).expand(*x.shape[:-1], *M_t.shape))
    f._psqrt_builtin = _LinearBuiltin(M_t)
    return f


transition_function = linear_function
observation_function = linear_function


def get_data(x0, A, H, R, Q, b, c, T, random_state=None, chol_R=None, dtype=np.float64):
    """Simulate an LGSSM (procedure of tests/_lgssm.py:42-95; fp64 by default here)."""
    if random_state is None or isinstance(random_state, int):
        random_state = np.random.RandomState(random_state)
    nr, nq = R.shape[0], Q.shape[0]
    normals = random_state.randn(T, nq + nr).astype(dtype)
    if chol_R is None:
        chol_R = np.linalg.cholesky(R)
    chol_Q = np.linalg.cholesky(Q)
    x = np.copy(x0).astype(dtype)
    observations = np.empty((T, nr), dtype=dtype)
    true_states = np.empty((T + 1, nq), dtype=dtype)
    true_states[0] = x
    for i in range(T):
        x = A @ x + chol_Q @ normals[i, :nq] + b
        true_states[i + 1] = x
        observations[i] = H @ x + chol_R @ normals[i, nq:] + c
    return true_states, observations
