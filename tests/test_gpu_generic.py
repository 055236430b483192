"""GPU tests of the generic path (csrc/psqrt_generic.cu): dimensions the tuned kernels do not cover (SURVEY 8b
"generic <= 16": nx in {7, 9..16}, ny > 4, ny > nx), the same path forced on tuned dimensions, batches, the standalone
smoother, psqrt_tria_batched / psqrt_chol_update_batched beyond 8 rows, a user-supplied torch model with nx = 7 through
the public API, and the float32 mode.  Checker: the NumPy oracle; fp64 tolerances are BASELINE.json's
(1e-9 means / L L^T, 1e-8 log-likelihood), the fp32 tolerance is stated in its test."""
import numpy as np
import pytest
import torch

import parsmooth_np as O
from _cases import LLt, lgssm_case, oracle_from_ssm, rel_err, time_varying_case

pytestmark = pytest.mark.gpu

TOL, TOL_ELL = 1e-9, 1e-8


def _dev():
    return torch.device("cuda", 0)


def _g(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=_dev())


def _ssm(case):
    from psqrt._lib import LinearizedSSM
    return LinearizedSSM(*[_g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])


def _check(name, m, L, om, oL, tol=TOL):
    em, eL = rel_err(m.cpu().numpy(), om), rel_err(LLt(L.cpu().numpy()), LLt(oL))
    assert em < tol and eL < tol, f"{name}: mean err {em:.3e}, LL^T err {eL:.3e}"


@pytest.mark.parametrize("n,ny,T", [(7, 3, 200), (9, 5, 129), (12, 6, 64), (16, 16, 33), (3, 5, 100), (10, 2, 1),
                                    (7, 7, 2), (11, 1, 1000), (6, 6, 50)])
def test_unsupported_dims_fall_back_to_generic(n, ny, T):
    """psqrt_filter_smoother on dimensions without tuned kernels (incl. ny > nx, where Z = tria(Z),
    parallel/_filtering.py:141-146)."""
    from psqrt import _lib
    assert not _lib.supported(n, ny) and _lib.supported_generic(n, ny)
    case = lgssm_case(n, ny, T, seed=10 * n + ny, triangular_prior=False)
    fm, fL, sm, sL, ell = _lib.filter_smoother(_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]),
                                               smooth=True, loglik=True)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    _check("filtered", fm, fL, ofm, ofc)
    _check("smoothed", sm, sL, osm, osc)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)


@pytest.mark.parametrize("n,ny,T", [(4, 2, 500), (5, 2, 77), (8, 4, 130), (1, 1, 40)])
def test_generic_matches_oracle_and_tuned(n, ny, T):
    """The generic path forced on tuned dimensions, time-varying model: against the oracle and the tuned kernels."""
    from psqrt import _lib
    case = time_varying_case(n, ny, T, seed=n + ny)
    args = (_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]))
    gen = _lib.filter_smoother(*args, smooth=True, loglik=True, generic=True)
    tun = _lib.filter_smoother(*args, smooth=True, loglik=True)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    _check("generic filtered", gen[0], gen[1], ofm, ofc)
    _check("generic smoothed", gen[2], gen[3], osm, osc)
    _check("generic vs tuned", gen[2], gen[3], tun[2].cpu().numpy(), tun[3].cpu().numpy())
    assert abs(gen[4].item() - oell) <= TOL_ELL * abs(oell)


def test_generic_batch_and_smoother_only():
    from psqrt import _lib
    n, ny, T, B = 7, 3, 90, 3
    cases = [time_varying_case(n, ny, T, seed=40 + k) for k in range(B)]
    stack = lambda key: _g(np.stack([c[key] for c in cases]))
    ssm = _lib.LinearizedSSM(*[stack(k) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, stack("ys"), stack("m0"), stack("L0"), smooth=True, loglik=True)
    for k, case in enumerate(cases):
        ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
        _check(f"seq {k} filtered", fm[k], fL[k], ofm, ofc)
        _check(f"seq {k} smoothed", sm[k], sL[k], osm, osc)
        assert abs(ell[k].item() - oell) <= TOL_ELL * abs(oell)
    # smoothing(...) on an existing filtered trajectory (psqrt_smoother -> generic)
    one = _lib.LinearizedSSM(*[_g(cases[0][k]) for k in ("F", "cholQ", "b")])
    sm1, sL1 = _lib.smoother(one, fm[0].contiguous(), fL[0].contiguous())
    _check("smoother only", sm1, sL1, sm[0].cpu().numpy(), sL[0].cpu().numpy())


def test_generic_tria_and_chol_update():
    from psqrt import _lib
    rng = np.random.RandomState(4)
    for rows, cols in ((7, 30), (9, 9), (12, 5), (16, 243), (10, 1)):
        A = rng.randn(11, rows, cols)
        L = _lib.tria(_g(A)).cpu().numpy()
        assert np.all(np.triu(L, 1) == 0)
        assert rel_err(LLt(L), A @ np.swapaxes(A, -1, -2)) < 1e-12
    for n, k in ((7, 3), (9, 1), (16, 5)):
        L0 = np.tril(rng.rand(6, n, n)) + 2 * np.eye(n)
        V = 0.2 * rng.randn(6, k, n)
        for alpha in (1.0, -1.0):
            got = _lib.chol_update_many(_g(L0), _g(V), alpha).cpu().numpy()
            ref = O.cholesky_update_many(L0, V, alpha)
            assert rel_err(got, ref) < 1e-10, (n, k, alpha)


def test_user_model_nx7_through_public_api():
    """A user-supplied torch model with nx = 7, ny = 3 (no tuned kernels): iterated extended smoother + log-likelihood
    through psqrt.methods, against the oracle on the same functions."""
    import psqrt
    n, ny, T = 7, 3, 120
    rng = np.random.RandomState(8)
    A = 0.9 * np.linalg.qr(rng.randn(n, n))[0]
    Hm = rng.randn(ny, n)
    cQ, cR = 0.1 * (np.tril(rng.rand(n, n)) + np.eye(n)), 0.3 * (np.tril(rng.rand(ny, ny)) + np.eye(ny))
    f_np = lambda x: np.einsum("ij,...j->...i", A, x) + 0.1 * np.sin(x)
    h_np = lambda x: np.einsum("ij,...j->...i", Hm, x) + 0.05 * np.cos(np.einsum("ij,...j->...i", Hm, x))
    f_np.jac = lambda x: A + 0.1 * (np.cos(x)[..., :, None] * np.eye(n))
    h_np.jac = lambda x: Hm - 0.05 * np.sin(np.einsum("ij,...j->...i", Hm, x))[..., :, None] * Hm
    At, Ht = _g(A), _g(Hm)
    f_t = lambda x: At @ x + 0.1 * torch.sin(x)
    h_t = lambda x: Ht @ x + 0.05 * torch.cos(Ht @ x)
    x = rng.randn(n)
    ys = np.zeros((T, ny))
    for t in range(T):
        x = f_np(x) + cQ @ rng.randn(n)
        ys[t] = h_np(x) + cR @ rng.randn(ny)
    m0, L0 = rng.randn(n), np.eye(n)
    res, ell = psqrt.iterated_smoothing(ys, psqrt.MVNSqrt(m0, L0),
                                        psqrt.FunctionalModel(f_t, psqrt.MVNSqrt(np.zeros(n), cQ)),
                                        psqrt.FunctionalModel(h_t, psqrt.MVNSqrt(np.zeros(ny), cR)),
                                        psqrt.linearization.extended, None, True, criterion=lambda i, *_: i < 4,
                                        return_loglikelihood=True)
    ores, oell = O.iterated_smoothing(ys, O.MVNSqrt(m0, L0), O.FunctionalModel(f_np, O.MVNSqrt(np.zeros(n), cQ)),
                                      O.FunctionalModel(h_np, O.MVNSqrt(np.zeros(ny), cR)), O.extended, None, True,
                                      lambda i, *_: i < 4, True)
    _check("nx = 7 iterated", res.mean, res.chol, ores.mean, ores.chol, tol=1e-7)
    assert abs(ell.item() - oell) <= 1e-7 * abs(oell)


def test_fp32_mode():
    """psqrt.fp32 (psqrt_filter_smoother_f32): the square-root pass in float32 against the fp64 oracle.  Tolerance:
    1e-3 relative on means and covariances for this well-conditioned LGSSM at T = 2000 (float32 has 2^-24 = 6e-8
    unit round-off; the scan depth and the conditioning of the problem amplify it), 1e-4 on the log-likelihood; and
    the robustness experiment's criterion -- no NaN, factors finite -- on an iterated bearings-only run."""
    import psqrt
    from psqrt import _lib, fp32
    from psqrt.models import bearings, lgssm
    case = lgssm_case(4, 2, 2000, seed=5)
    x0 = psqrt.MVNSqrt(case["m0"], case["L0"])
    tm = psqrt.FunctionalModel(lgssm.transition_function(case["F"]), psqrt.MVNSqrt(case["b"], case["cholQ"]))
    om = psqrt.FunctionalModel(lgssm.observation_function(case["H"]), psqrt.MVNSqrt(case["c"], case["cholR"]))
    smo, ell = fp32.filter_smoother(case["ys"], x0, tm, om, psqrt.linearization.extended, None, True)
    assert smo.mean.dtype == torch.float32 and smo.chol.dtype == torch.float32
    _, _, osm, osc, oell = oracle_from_ssm(case)
    assert rel_err(smo.mean.double().cpu().numpy(), osm) < 1e-3
    assert rel_err(LLt(smo.chol.double().cpu().numpy()), LLt(osc)) < 1e-3
    assert abs(ell.item() - oell) <= 1e-4 * abs(oell)
    # bearings-only, iterated, float32
    T = 1000
    s1, s2 = np.array([-1.5, 0.5]), np.array([1.0, 1.0])
    _, _, ys = bearings.get_data(np.array([0.1, 0.2, 1.0, 0.0]), 0.01, 0.5, T, s1, s2, random_state=3)
    Q, R, obs_f, trans_f = bearings.make_parameters(0.01, 0.1, 0.5, 0.01, s1, s2)
    x0 = psqrt.MVNSqrt(np.array([-1.0, -1.0, 0.0, 0.0, 0.0]), np.eye(5))
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(np.zeros(2), np.linalg.cholesky(R)))
    res32 = fp32.iterated_smoothing(ys.astype(np.float64), x0, tm, om, psqrt.linearization.extended, None,
                                    criterion=lambda i, *_: i < 5)
    res64 = psqrt.iterated_smoothing(ys.astype(np.float64), x0, tm, om, psqrt.linearization.extended, None, True,
                                     criterion=lambda i, *_: i < 5)
    assert int(_lib.count_nonfinite(res32.mean.double()[None]).item()) == 0
    assert int(_lib.count_nonfinite(res32.chol.double()[None]).item()) == 0
    assert rel_err(res32.mean.double().cpu().numpy(), res64.mean.cpu().numpy()) < 5e-2
