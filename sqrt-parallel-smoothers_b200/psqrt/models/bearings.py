"""Coordinated-turn dynamics observed by two bearings-only sensors
(reference: tests/bearings/bearings_utils.py:7-115 == notebooks/bearing_data.py:11-134;
parameter-estimation variant notebooks/bearing_data_pe.py:123-134)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def make_transition_function(dt: float):
    """x = (px, py, vx, vy, w).  For |w| < 1e-6 the reference's lax.cond takes sin(wt)/w -> dt and
    (cos(wt)-1)/w -> 0 as CONSTANTS (bearings_utils.py:24-37), so their w-derivative is zero."""

    def _parts(x):
        w = x[..., 4]
        small = torch.abs(w) < 1e-6
        ws = torch.where(small, torch.ones_like(w), w)
        cw, sw = torch.cos(w * dt), torch.sin(w * dt)
        a = torch.where(small, torch.full_like(w, dt), sw / ws)
        bq = torch.where(small, torch.zeros_like(w), (cw - 1) / ws)
        return w, small, ws, cw, sw, a, bq

    def f(x):
        w, small, ws, cw, sw, a, bq = _parts(x)
        px, py, vx, vy = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
        return torch.stack([px + a * vx - bq * vy, py + bq * vx + a * vy, cw * vx + sw * vy,
                            -sw * vx + cw * vy, w], -1)

    def value_and_jac(x):
        w, small, ws, cw, sw, a, bq = _parts(x)
        vx, vy = x[..., 2], x[..., 3]
        zero = torch.zeros_like(w)
        da = torch.where(small, zero, (dt * cw * ws - sw) / (ws * ws))
        db = torch.where(small, zero, (-dt * sw * ws - (cw - 1)) / (ws * ws))
        dcw, dsw = -dt * sw, dt * cw
        J = x.new_zeros(x.shape[:-1] + (5, 5))
        J[..., 0, 0] = 1
        J[..., 0, 2] = a
        J[..., 0, 3] = -bq
        J[..., 0, 4] = da * vx - db * vy
        J[..., 1, 1] = 1
        J[..., 1, 2] = bq
        J[..., 1, 3] = a
        J[..., 1, 4] = db * vx + da * vy
        J[..., 2, 2] = cw
        J[..., 2, 3] = sw
        J[..., 2, 4] = dcw * vx + dsw * vy
        J[..., 3, 2] = -sw
        J[..., 3, 3] = cw
        J[..., 3, 4] = -dsw * vx + dcw * vy
        J[..., 4, 4] = 1
        return f(x), J

    f._psqrt_batched = True
    f._psqrt_value_and_jac = value_and_jac
    from ._builtin import Builtin
    f._psqrt_builtin = Builtin(_lib.MODEL_CT_TRANSITION, [float(dt)], 5, 5)
    return f


def make_observation_function(s1, s2):
    s1 = [float(v) for v in np.asarray(s1).reshape(-1)]
    s2 = [float(v) for v in np.asarray(s2).reshape(-1)]

    def f(x):
        return torch.stack([torch.atan2(x[..., 1] - s1[1], x[..., 0] - s1[0]),
                            torch.atan2(x[..., 1] - s2[1], x[..., 0] - s2[0])], -1)

    def value_and_jac(x):
        J = x.new_zeros(x.shape[:-1] + (2, x.shape[-1]))
        for i, s in enumerate((s1, s2)):
            dx, dy = x[..., 0] - s[0], x[..., 1] - s[1]
            r2 = dx * dx + dy * dy
            J[..., i, 0] = -dy / r2
            J[..., i, 1] = dx / r2
        return f(x), J

    f._psqrt_batched = True
    f._psqrt_value_and_jac = value_and_jac
    from ._builtin import Builtin
    f._psqrt_builtin = Builtin(_lib.MODEL_BEARINGS_OBSERVATION, [s1[0], s1[1], s2[0], s2[1]], 5, 2)
    return f


def make_parameters(qc, qw, r, dt, s1, s2, r2=None):
    """-> Q, R, observation_function, transition_function (bearings_utils.py:72-115);
    r2 = 0.1 gives the parameter-estimation variant R = diag(r^2, r2^2) (bearing_data_pe.py:129)."""
    Q = np.array([[qc * dt ** 3 / 3, 0, qc * dt ** 2 / 2, 0, 0],
                  [0, qc * dt ** 3 / 3, 0, qc * dt ** 2 / 2, 0],
                  [qc * dt ** 2 / 2, 0, qc * dt, 0, 0],
                  [0, qc * dt ** 2 / 2, 0, qc * dt, 0],
                  [0, 0, 0, 0, dt * qw]])
    R = r ** 2 * np.eye(2) if r2 is None else np.diag([r ** 2, r2 ** 2])
    return Q, R, make_observation_function(s1, s2), make_transition_function(dt)


def inverse_bearings(observations, s1, s2):
    """Positions as if the two bearings were noise-free (notebooks/bearing_data_pe.py:68-90): the initial linearization
    points of the parameter-estimation experiments.  observations [..., 2] -> [..., 2]."""
    ys = np.asarray(observations, dtype=np.float64)
    t1, t2 = np.tan(ys[..., 0]), np.tan(ys[..., 1])
    b1, b2 = s1[0] * t1 - s1[1], s2[0] * t2 - s2[1]
    # [[t1, -1], [t2, -1]] p = [b1, b2]
    px = (b1 - b2) / (t1 - t2)
    return np.stack([px, t1 * px - b1], -1)


def get_data_pe(x0, dt, r, T, s1, s2, q=10.0, random_state=None, r2=0.1):
    """Parameter-estimation variant (notebooks/bearing_data_pe.py:137-198): first sensor's noise std r, second
    sensor's fixed at r2 = 0.1."""
    return get_data(x0, dt, r, T, s1, s2, q=q, random_state=random_state, r2=r2)


def get_data(x0, dt, r, T, s1, s2, q=10.0, random_state=None, r2=None):
    """Simulated trajectory + bearings (procedure of notebooks/bearing_data.py:137-198, float32
    like the reference; the closed-form matrix exponential of the turn replaces scipy.linalg.expm)."""
    if random_state is None or isinstance(random_state, int):
        random_state = np.random.RandomState(random_state)
    a_s = (1 + q * dt * np.cumsum(random_state.randn(T))).astype(np.float32)
    s1 = np.asarray(s1, dtype=np.float32)
    s2 = np.asarray(s2, dtype=np.float32)
    x = np.copy(x0).astype(np.float32)
    observations = np.empty((T, 2), dtype=np.float32)
    true_states = np.zeros((T + 1, 5), dtype=np.float32)
    ts = np.linspace(dt, (T + 1) * dt, T).astype(np.float32)
    true_states[0, :4] = x
    normals = random_state.randn(T, 2).astype(np.float32)
    for i, a in enumerate(a_s):
        a = float(a)
        if abs(a) < 1e-12:
            sa, ca1 = dt, 0.0
        else:
            sa, ca1 = np.sin(a * dt) / a, (1 - np.cos(a * dt)) / a
        c, s = np.cos(a * dt), np.sin(a * dt)
        E = np.array([[1, 0, sa, ca1], [0, 1, -ca1, sa], [0, 0, c, s], [0, 0, -s, c]], dtype=np.float32)
        x = E @ x
        y1 = np.arctan2(x[1] - s1[1], x[0] - s1[0]) + r * normals[i, 0]
        y2 = np.arctan2(x[1] - s2[1], x[0] - s2[0]) + (r if r2 is None else r2) * normals[i, 1]
        observations[i] = [y1, y2]
        true_states[i + 1] = np.concatenate((x, np.array([a], dtype=np.float32)))
    return ts, true_states, observations
