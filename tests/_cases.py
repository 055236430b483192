"""Seeded synthetic cases shared by the CPU and GPU tests (SURVEY.md section 8d recipes)."""
import numpy as np

import parsmooth_np as O


def lgssm_case(n, ny, T, seed, triangular_prior=True):
    """C1 recipe: stable F = 0.99 * orth, cholQ = 0.1 (tril(U) + I), H ~ N, cholR = 0.5 (tril(U) + I)."""
    rng = np.random.RandomState(seed)
    Qr, _ = np.linalg.qr(rng.randn(n, n))
    F = 0.99 * Qr
    cholQ = 0.1 * (np.tril(rng.rand(n, n)) + np.eye(n))
    H = rng.randn(ny, n)
    cholR = 0.5 * (np.tril(rng.rand(ny, ny)) + np.eye(ny))
    b = 0.1 * rng.randn(n)
    c = 0.1 * rng.randn(ny)
    m0 = rng.randn(n)
    L0 = np.tril(rng.rand(n, n)) + np.eye(n) if triangular_prior else rng.randn(n, n)
    _, ys = O.lgssm_get_data(m0, F, H, cholR @ cholR.T, cholQ @ cholQ.T, b, c, T, random_state=rng, dtype=np.float64)
    return dict(F=F, cholQ=cholQ, b=b, H=H, cholR=cholR, c=c, ys=ys, m0=m0, L0=L0)


def oracle_lgssm_models(case):
    tm = O.FunctionalModel(O.lgssm_function(case["F"]), O.MVNSqrt(case["b"], case["cholQ"]))
    om = O.FunctionalModel(O.lgssm_function(case["H"]), O.MVNSqrt(case["c"], case["cholR"]))
    return tm, om


def time_varying_case(n, ny, T, seed):
    """Random per-step linearised SSM (what a nonlinear model produces)."""
    rng = np.random.RandomState(seed)
    F = np.stack([0.95 * np.linalg.qr(rng.randn(n, n))[0] for _ in range(T)]) + 0.02 * rng.randn(T, n, n)
    cholQ = 0.2 * (np.tril(rng.rand(T, n, n)) + np.eye(n))
    b = 0.1 * rng.randn(T, n)
    H = rng.randn(T, ny, n)
    cholR = 0.3 * (np.tril(rng.rand(T, ny, ny)) + np.eye(ny))
    c = 0.1 * rng.randn(T, ny)
    ys = rng.randn(T, ny)
    m0 = rng.randn(n)
    L0 = np.tril(rng.rand(n, n)) + np.eye(n)
    return dict(F=F, cholQ=cholQ, b=b, H=H, cholR=cholR, c=c, ys=ys, m0=m0, L0=L0)


def oracle_from_ssm(case, scan=None):
    """Oracle parallel filter + smoother straight from a linearised SSM."""
    scan = scan or O.associative_scan
    T = case["ys"].shape[0]
    n = case["m0"].shape[0]
    bc = lambda a, core: np.broadcast_to(a, (T,) + a.shape[-core:]) if a.ndim == core else a
    F, cQ, b = bc(case["F"], 2), bc(case["cholQ"], 2), bc(case["b"], 1)
    H, cR, c = bc(case["H"], 2), bc(case["cholR"], 2), bc(case["c"], 1)
    ms = np.concatenate([case["m0"][None], np.zeros((T - 1, n))])
    Ls = np.concatenate([case["L0"][None], np.zeros((T - 1, n, n))])
    el = O.sqrt_filtering_elements(F, cQ, b, H, cR, c, ms, Ls, case["ys"])
    _, fm, fc, _, _ = scan(O.sqrt_filtering_operator, el)
    fm = np.concatenate([case["m0"][None], fm])
    fc = np.concatenate([case["L0"][None], fc])
    ell = np.sum(O.sqrt_loglikelihood_terms(F, cQ, b, H, cR, c, fm[:-1], fc[:-1], case["ys"]))
    g, E, D = O.sqrt_smoothing_elements(F, cQ, b, fm[:-1], fc[:-1])
    g = np.concatenate([g, fm[-1:]])
    E = np.concatenate([E, np.zeros((1, n, n))])
    D = np.concatenate([D, fc[-1:]])
    sm, _, sc = scan(O.sqrt_smoothing_operator, (g, E, D), reverse=True)
    return fm, fc, sm, sc, ell


def LLt(L):
    return L @ np.swapaxes(L, -1, -2)


def rel_err(a, b):
    """max |a - b| / max |b|: error relative to the scale of the reference quantity."""
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def bearings_pe_case(T, seed, r_true=0.05, dt=0.01):
    """Parameter-estimation variant of the bearings-only problem (notebooks/bearing_data_pe.py:123-146,
    experiment_bearing_only_param_estimation_run_time.ipynb): first sensor's noise std r = 1 / prec is the
    parameter, second sensor fixed at 0.1.  Data simulated with the coordinated-turn dynamics in float64; the
    prior is centred near the truth so that the iterated smoothers converge in a few iterations."""
    rng = np.random.RandomState(seed)
    s1, s2 = np.array([-1.0, 0.5]), np.array([1.0, 1.0])
    qc = qw = 0.1
    x = np.array([0.1, 0.2, 1.0, 0.0, 1.0])
    f = O.ct_transition_function(dt)
    Q, _, _, _ = O.bearings_make_parameters(qc, qw, r_true, dt, s1, s2, r2=0.1)
    cQ = np.linalg.cholesky(Q)
    ys = np.zeros((T, 2))
    for t in range(T):
        x = f(x) + cQ @ rng.randn(5)
        ys[t, 0] = np.arctan2(x[1] - s1[1], x[0] - s1[0]) + r_true * rng.randn()
        ys[t, 1] = np.arctan2(x[1] - s2[1], x[0] - s2[0]) + 0.1 * rng.randn()
    m0 = np.array([0.1, 0.2, 1.0, 0.0, 1.0]) + 0.05 * rng.randn(5)
    L0 = np.diag([0.2, 0.2, 0.3, 0.3, 0.5])
    return dict(ys=ys, m0=m0, L0=L0, s1=s1, s2=s2, qc=qc, qw=qw, dt=dt)


def oracle_bearings_pe_models(case, prec):
    """(transition_model, observation_model) at precision prec = 1 / r of the first sensor."""
    Q, R, obs, trans = O.bearings_make_parameters(case["qc"], case["qw"], 1.0 / prec, case["dt"], case["s1"],
                                                  case["s2"], r2=0.1)
    tm = O.FunctionalModel(trans, O.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    om = O.FunctionalModel(obs, O.MVNSqrt(np.zeros(2), np.linalg.cholesky(R)))
    return tm, om
