"""Pins the oracle (oracle/parsmooth_np.py) before anything is compared with it (CPU only).

1. The reference's stored goldens -- tests/bearings/{ys.npy, ieks.npz, icks.npz} upstream, copied as DATA
   into tests/golden/ (upstream produced them with a float32 run and compares at 3 decimals,
   tests/test_bearings_only.py:16,59-72).
2. Vectors produced by executing the reference's OWN source files on a NumPy shim of the JAX surface they
   use (tests/golden/make_golden.py -> tests/golden/reference_vectors.npz).
3. The reference's cross-implementation invariants re-run on the restatement: sqrt == standard operators
   (tests/test_parallel_operators.py), parallel == sequential (tests/test_parallel_filter.py:80-122,
   tests/test_parallel_smoother.py:61-99), association-order independence of the scan.
"""
import os

import numpy as np
import pytest

import parsmooth_np as O
from _cases import LLt, lgssm_case, oracle_lgssm_models, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bearings():
    ys = np.load(os.path.join(GOLD, "bearings_ys.npy")).astype(np.float64)
    Q, R, obs, trans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    x0 = O.MVNSqrt(np.array([-1.0, -1.0, 0.0, 0.0, 0.0]), np.eye(5))
    tm = O.FunctionalModel(trans, O.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    om = O.FunctionalModel(obs, O.MVNSqrt(np.zeros(2), np.linalg.cholesky(R)))
    return ys, x0, tm, om


def test_reference_golden_ieks():
    """100-iteration IEKS, the reference's own call (tests/test_bearings_only.py:57-59)."""
    ys, x0, tm, om = _bearings()
    with np.load(os.path.join(GOLD, "bearings_ieks.npz")) as z:
        exp_m, exp_P = z["arr_0"], z["arr_1"]
    res = O.iterated_smoothing(ys, x0, tm, om, O.extended, None, True, criterion=lambda i, *_: i < 100)
    np.testing.assert_array_almost_equal(res.mean[1:], exp_m, decimal=3)
    np.testing.assert_array_almost_equal(LLt(res.chol)[1:], exp_P, decimal=3)
    assert np.max(np.abs(res.mean[1:] - exp_m)) < 1e-4      # observed 1.1e-5 (float32 golden)


def test_reference_golden_icks():
    """100-pass ICKS.  On this data the cubature iteration does not converge: it settles on a
    period-4 limit cycle (successive iterates differ by ~0.17), so the iterate COUNT matters.  The
    stored golden equals 100 filter+smoother passes in total, i.e. the initial pass plus 99
    fixed-point applications; the current upstream loop (methods.py:63-71 with _utils.py:136-146)
    performs the initial pass plus 100 -- the upstream test is skipped (test_bearings_only.py:21),
    which is how the one-pass drift went unnoticed.  Pinned here with 99 applications (5.5e-5 from
    the golden); with 100 the restatement sits 0.18 away, the size of the cycle."""
    ys, x0, tm, om = _bearings()
    with np.load(os.path.join(GOLD, "bearings_icks.npz")) as z:
        exp_m, exp_P = z["arr_0"], z["arr_1"]
    res = O.iterated_smoothing(ys, x0, tm, om, O.cubature, None, True, criterion=lambda i, *_: i < 99)
    np.testing.assert_array_almost_equal(res.mean[1:], exp_m, decimal=3)
    np.testing.assert_array_almost_equal(LLt(res.chol)[1:], exp_P, decimal=3)
    nxt = O.filter_smoother(ys, x0, tm, om, O.cubature, res, True)
    assert np.max(np.abs(nxt.mean - res.mean)) > 0.1        # the limit cycle, not convergence


@pytest.mark.parametrize("dim_x", [1, 2, 3, 5])
@pytest.mark.parametrize("seed", [0, 42])
def test_sqrt_vs_standard_operators(dim_x, seed):
    """tests/test_parallel_operators.py:17-89 on the restatement."""
    np.random.seed(seed)
    tri = lambda: np.tril(np.random.rand(dim_x, dim_x))
    A1, A2 = np.random.randn(dim_x, dim_x), np.random.randn(dim_x, dim_x)
    b1, b2 = np.random.randn(dim_x), np.random.randn(dim_x)
    U1, U2 = tri(), tri()
    eta1, eta2 = np.random.randn(dim_x), np.random.randn(dim_x)
    Z1, Z2 = tri(), tri()
    As, bs, C, etas, J = O.standard_filtering_operator((A1, b1, U1 @ U1.T, eta1, Z1 @ Z1.T),
                                                       (A2, b2, U2 @ U2.T, eta2, Z2 @ Z2.T))
    Aq, bq, U, etaq, Z = O.sqrt_filtering_operator((A1, b1, U1, eta1, Z1), (A2, b2, U2, eta2, Z2))
    for a, b in ((As, Aq), (bs, bq), (etas, etaq), (C, LLt(U)), (J, LLt(Z))):
        np.testing.assert_allclose(a, b, atol=1e-6, rtol=1e-6)
    g, E, L = O.standard_smoothing_operator((b1, A1, U1 @ U1.T), (b2, A2, U2 @ U2.T))
    gq, Eq, D = O.sqrt_smoothing_operator((b1, A1, U1), (b2, A2, U2))
    np.testing.assert_allclose(g, gq, atol=1e-9)
    np.testing.assert_allclose(E, Eq, atol=1e-9)
    np.testing.assert_allclose(L, LLt(D), atol=1e-9)


@pytest.mark.parametrize("dim_x,dim_y", [(1, 1), (2, 1), (3, 2), (1, 3), (2, 3), (4, 2)])
@pytest.mark.parametrize("lin", [O.extended, O.cubature, O.gauss_hermite])
def test_parallel_vs_sequential(dim_x, dim_y, lin):
    """parallel sqrt == sequential sqrt (filter, ell, smoother) with an explicit nominal trajectory."""
    T = 30
    case = lgssm_case(dim_x, dim_y, T, seed=dim_x * 7 + dim_y)
    tm, om = oracle_lgssm_models(case)
    rng = np.random.RandomState(3)
    nominal = O.MVNSqrt(rng.randn(T + 1, dim_x), np.repeat(np.eye(dim_x)[None], T + 1, 0))
    x0 = O.MVNSqrt(case["m0"], case["L0"])
    fp, ellp = O.par_filtering(case["ys"], x0, tm, om, lin, nominal, True)
    fs, ells = O.seq_filtering(case["ys"], x0, tm, om, lin, nominal, True)
    assert rel_err(fp.mean, fs.mean) < 1e-10 and rel_err(LLt(fp.chol), LLt(fs.chol)) < 1e-10
    assert abs(ellp - ells) < 1e-9 * abs(ells)
    sp = O.par_smoothing(tm, fp, lin, nominal)
    ss = O.seq_smoothing(tm, fs, lin, nominal)
    assert rel_err(sp.mean, ss.mean) < 1e-10 and rel_err(LLt(sp.chol), LLt(ss.chol)) < 1e-10


def test_scan_association_order():
    """tree order (associative_scan) == left fold to rounding."""
    case = lgssm_case(4, 2, 300, seed=1)
    tm, om = oracle_lgssm_models(case)
    x0 = O.MVNSqrt(case["m0"], case["L0"])
    a = O.par_filtering(case["ys"], x0, tm, om, O.extended, None, scan=O.associative_scan)
    b = O.par_filtering(case["ys"], x0, tm, om, O.extended, None, scan=O.sequential_fold_scan)
    assert rel_err(a.mean, b.mean) < 1e-12 and rel_err(LLt(a.chol), LLt(b.chol)) < 1e-12
    sa = O.par_smoothing(tm, a, O.extended, scan=O.associative_scan)
    sb = O.par_smoothing(tm, a, O.extended, scan=O.sequential_fold_scan)
    assert rel_err(sa.mean, sb.mean) < 1e-12 and rel_err(LLt(sa.chol), LLt(sb.chol)) < 1e-12


def test_scan_edge_lengths():
    for n in (1, 2, 3, 4, 5, 8, 9):
        x = (np.arange(1.0, n + 1)[:, None],)
        add = lambda a, b: (a[0] + b[0],)
        np.testing.assert_allclose(O.associative_scan(add, x)[0][:, 0], np.cumsum(np.arange(1.0, n + 1)))
        np.testing.assert_allclose(O.associative_scan(add, x, reverse=True)[0][:, 0],
                                   np.cumsum(np.arange(1.0, n + 1)[::-1])[::-1])


@pytest.mark.parametrize("multiplier", [1.0, -0.1])
@pytest.mark.parametrize("seed", [0, 42, 666])
@pytest.mark.parametrize("dim_x", [2, 3, 10, 11])
def test_cholesky_update(multiplier, seed, dim_x):
    """tests/test_math_utils.py:16-54."""
    np.random.seed(seed)
    cholQ = np.tril(np.random.rand(dim_x, dim_x))
    v = np.random.randn(dim_x)
    expected = cholQ @ cholQ.T + multiplier * v[:, None] @ v[None, :]
    if min(np.linalg.eigvals(expected).real) <= 1e-6:
        pytest.skip("random vectors do not result in a positive definite matrix.")
    res = O.cholesky_update(cholQ, v, multiplier)
    np.testing.assert_allclose(res @ res.T, expected, rtol=1e-4)
    np.testing.assert_allclose(res, np.linalg.cholesky(expected), atol=1e-6, rtol=1e-4)


def test_cholesky_update_nonfinite_guard():
    """_utils.py:80: a downdate past positive definiteness yields zeros, never NaN."""
    L = 0.1 * np.eye(3)
    out = O.cholesky_update(L, np.array([1.0, 2.0, 3.0]), -1.0)
    assert np.all(np.isfinite(out)) and np.all(np.triu(out, 1) == 0)


@pytest.mark.parametrize("lin", [O.extended, O.cubature, O.gauss_hermite])
@pytest.mark.parametrize("dim_x", [1, 3])
def test_linear_functional(lin, dim_x):
    """tests/test_linearization.py:67-107: every method recovers (a, c + m_q, chol_q chol_q^T)."""
    np.random.seed(0)
    a, c = np.random.randn(dim_x, dim_x), np.random.randn(dim_x)
    m_x, m_q = np.random.randn(dim_x), np.random.randn(dim_x)
    chol_x, chol_q = np.tril(np.random.rand(dim_x, dim_x)), np.tril(np.random.rand(dim_x, dim_x))

    def fun(x):
        return np.einsum("ij,...j->...i", a, x) + c

    fun.jac = lambda x: np.broadcast_to(a, x.shape[:-1] + a.shape)
    F, Ql, rem = lin(O.FunctionalModel(fun, O.MVNSqrt(m_q, chol_q)), O.MVNSqrt(m_x[None], chol_x[None]))
    xp = np.random.randn(dim_x)
    np.testing.assert_allclose(F[0], a, atol=1e-7)
    np.testing.assert_allclose(F[0] @ xp + rem[0], fun(xp) + m_q, atol=1e-7)
    np.testing.assert_allclose(LLt(Ql[0]), chol_q @ chol_q.T, atol=1e-7)


def test_gauss_hermite_constants():
    """_gh.py:73-126 evaluated here: order 3 nodes {0, +sqrt3, -sqrt3} (np.roots order), weights
    {2/3, 1/6, 1/6}; first state dimension varies fastest."""
    wm, wc, xi = O.gauss_hermite_weights(2, 3)
    assert xi.shape == (2, 9) and abs(wm.sum() - 1) < 1e-14
    np.testing.assert_allclose(sorted(set(np.round(xi[0], 12))), [-np.sqrt(3), 0.0, np.sqrt(3)], atol=1e-12)
    assert abs(xi[0, 0]) < 1e-12 and xi[0, 1] > 0 and xi[0, 2] < 0       # 0, +sqrt3, -sqrt3
    np.testing.assert_allclose(xi[1, :3], xi[1, 0])                       # dim 1 constant over the first 3
    np.testing.assert_allclose(wm[0], 4 / 9, rtol=1e-13)
    wmc, _, xic = O.cubature_weights(3)
    np.testing.assert_allclose(xic[:3], np.sqrt(3) * np.eye(3))
    np.testing.assert_allclose(wmc, 1 / 6)


def test_reference_source_vectors():
    """Outputs of the reference's own source (run on the NumPy JAX-shim by tests/golden/make_golden.py)."""
    path = os.path.join(GOLD, "reference_vectors.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/reference_vectors.npz not generated")
    from golden.check_vectors import check_all
    check_all(path)


# ---- SURVEY 8(f) components: unscented linearization, pathwise sampler --------------------------------
def _next_vectors():
    return np.load(os.path.join(GOLD, "reference_vectors_next.npz"))


def test_unscented_vs_reference_source():
    """oracle unscented == the reference's own parsmooth/linearization/_unscented.py run on the NumPy shim
    (tests/golden/make_golden_next.py): weights, functional (bearings) and conditional (population) models,
    and a whole LGSSM pass."""
    z = _next_vectors()
    for n in (1, 2, 5):
        wm, wc, _ = O.unscented_weights(n)
        np.testing.assert_allclose(wm, z[f"ut{n}_wm"], rtol=0, atol=1e-16)
        np.testing.assert_allclose(wc, z[f"ut{n}_wc"], rtol=0, atol=1e-16)
    Q, R, obs, trans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    tm = O.FunctionalModel(trans, O.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    om = O.FunctionalModel(obs, O.MVNSqrt(np.zeros(2), np.linalg.cholesky(R)))
    x = O.MVNSqrt(z["bear_pts_m"], z["bear_pts_L"])
    for name, model in (("t", tm), ("o", om)):
        F, ch, b = O.unscented(model, x)
        assert rel_err(F, z[f"bear_ut_{name}_F"]) < 1e-12 and rel_err(b, z[f"bear_ut_{name}_b"]) < 1e-12
        # Phi + Q - F P F^T is a difference of nearly equal matrices (_sigma_points.py:77-78): its error is
        # measured on the scale of the terms
        FL = F @ x.chol
        scale = max(float(np.abs(LLt(ch)).max()), float(np.abs(LLt(FL)).max()))
        assert float(np.abs(LLt(ch) - LLt(z[f"bear_ut_{name}_chol"])).max()) / scale < 1e-12
    tmod, omod = O.population_model(10.0, np.array([[0.09]]))
    xp = O.MVNSqrt(z["pop_pts_m"], z["pop_pts_L"])
    for name, model in (("t", tmod), ("o", omod)):
        F, ch, b = O.unscented(model, xp)
        assert rel_err(F, z[f"pop_ut_{name}_F"]) < 1e-12 and rel_err(b, z[f"pop_ut_{name}_b"]) < 1e-12
        assert rel_err(LLt(ch), LLt(z[f"pop_ut_{name}_chol"])) < 1e-10
    ltm = O.FunctionalModel(O.lgssm_function(z["lg_F"]), O.MVNSqrt(z["lg_b"], z["lg_cQ"]))
    lom = O.FunctionalModel(O.lgssm_function(z["lg_H"]), O.MVNSqrt(z["lg_c"], z["lg_cR"]))
    nom = O.MVNSqrt(z["lg_nom_m"], np.repeat(np.eye(3)[None], 15, 0))
    f, ell = O.filtering(z["lg_ys"], O.MVNSqrt(z["lg_m0"], z["lg_L0"]), ltm, lom, O.unscented, nom, True, True)
    s = O.smoothing(ltm, f, O.unscented, nom, True)
    assert rel_err(f.mean, z["lg_ut_fm"]) < 1e-10 and rel_err(LLt(f.chol), LLt(z["lg_ut_fc"])) < 1e-10
    assert rel_err(s.mean, z["lg_ut_sm"]) < 1e-10 and rel_err(LLt(s.chol), LLt(z["lg_ut_sc"])) < 1e-10
    assert abs(ell - z["lg_ut_ell"]) <= 1e-10 * abs(z["lg_ut_ell"])


def test_sampler_vs_reference_source():
    """oracle pathwise sampler == parsmooth/_pathwise_sampler.py (parallel and sequential) on the same normal
    draws; parallel == sequential (the reference's tests/test_sampler.py compares both with the smoother)."""
    z = _next_vectors()
    ltm = O.FunctionalModel(O.lgssm_function(z["lg_F"]), O.MVNSqrt(z["lg_b"], z["lg_cQ"]))
    f, s = O.MVNSqrt(z["smp_fm"], z["smp_fc"]), O.MVNSqrt(z["smp_sm"], z["smp_sc"])
    for lname, lin in (("ext", O.extended), ("cub", O.cubature)):
        a = O.par_sampling(z["smp_eps"], ltm, f, lin, s)
        b = O.seq_sampling(z["smp_eps"], ltm, f, lin, s)
        assert rel_err(a, z[f"smp_{lname}_par"]) < 1e-12 and rel_err(b, z[f"smp_{lname}_seq"]) < 1e-12
        assert rel_err(a, b) < 1e-12
    Q, _, _, trans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    tm = O.FunctionalModel(trans, O.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    fb, sb = O.MVNSqrt(z["smpb_fm"], z["smpb_fc"]), O.MVNSqrt(z["smpb_sm"], z["smpb_sc"])
    a = O.par_sampling(z["smpb_eps"], tm, fb, O.extended, sb)
    assert rel_err(a, z["smpb_ext_par"]) < 1e-12


def _gcm_function(a, b, c):
    """f(x, q) = a x + sin(x) q + b q + 0.3 q^2 + c (non-linear in both arguments), NumPy-batched, with partials."""
    eye = np.eye(len(a))

    def f(x, q):
        return np.einsum("ij,...j->...i", a, x) + np.sin(x) * q + np.einsum("ij,...j->...i", b, q) + 0.3 * q * q + c

    f.jac_q = lambda x, q: np.einsum("...i,ij->...ij", np.sin(x) + 0.6 * q, eye) + b
    f.jac_x = lambda x, q: a + np.einsum("...i,ij->...ij", np.cos(x) * q, eye)
    return f


def test_get_conditional_model_vs_reference_source():
    """oracle get_conditional_model == parsmooth/linearization/_common.py:17-66 run on the NumPy shim, for the
    sigma-point methods (outer == inner, as tests/test_linearization.py:243-284) and inner extended under an outer
    cubature; dimension mismatch raises NotImplementedError like upstream."""
    z = _next_vectors()
    for n in (1, 2):
        f = _gcm_function(z[f"gcm{n}_a"], z[f"gcm{n}_b"], z[f"gcm{n}_c"])
        q = O.MVNSqrt(z[f"gcm{n}_qm"], z[f"gcm{n}_qL"])
        x = O.MVNSqrt(z[f"gcm{n}_xm"], z[f"gcm{n}_xL"])
        for tag, inner, outer in (("cub", O.cubature, O.cubature), ("gh", O.gauss_hermite, O.gauss_hermite),
                                  ("ut", O.unscented, O.unscented), ("extcub", O.extended, O.cubature)):
            F, ch, b = outer(O.get_conditional_model(f, q, inner), x)
            assert rel_err(F, z[f"gcm{n}_{tag}_F"]) < 1e-12 and rel_err(b, z[f"gcm{n}_{tag}_b"]) < 1e-12
            assert rel_err(LLt(ch), LLt(z[f"gcm{n}_{tag}_chol"])) < 1e-12
    # the reference's own check on a linear f (tests/test_linearization.py:243-284), extended in extended
    rng = np.random.RandomState(0)
    a, b, c = rng.randn(2, 2), rng.randn(2, 2), rng.randn(2)
    f = lambda x, q: np.einsum("ij,...j->...i", a, x) + np.einsum("ij,...j->...i", b, q) + c      # noqa: E731
    f.jac_q = lambda x, q: np.broadcast_to(b, x.shape[:-1] + b.shape)
    f.jac_x = lambda x, q: np.broadcast_to(a, x.shape[:-1] + a.shape)
    q = O.MVNSqrt(rng.randn(2), np.tril(rng.rand(2, 2)))
    x = O.MVNSqrt(rng.randn(2), np.tril(rng.rand(2, 2)))
    for lin in (O.extended, O.cubature):
        F, ch, rem = lin(O.get_conditional_model(f, q, lin), x)
        np.testing.assert_allclose(F, a, atol=1e-10)
        np.testing.assert_allclose(rem, b @ q.mean + c, atol=1e-10)
        np.testing.assert_allclose(LLt(ch), LLt(b @ q.chol), atol=1e-10)
    with pytest.raises(NotImplementedError):
        O.get_conditional_model(lambda x, q: a @ x + np.ones((2, 3)) @ q, O.MVNSqrt(np.zeros(3), np.eye(3)), O.extended)


def test_loglikelihood_tangent_vs_finite_differences():
    """Groundwork for the gradient row: the oracle's forward-mode tangent of the log-likelihood
    equals central finite differences of the (reference-pinned) parallel filter's ell, for a parameter in the
    observation noise factor (the `r` of the bearings experiment) and one in the transition matrix."""
    case = lgssm_case(3, 2, 60, seed=5)
    T = case["ys"].shape[0]
    tile = lambda a: np.broadcast_to(a, (T,) + a.shape).copy()
    dirs = {
        "cholR": np.array([[1.0, 0.0], [0.3, 0.5]]),
        "F": np.array([[0.2, -0.1, 0.0], [0.0, 0.1, 0.3], [0.05, 0.0, -0.2]]),
    }

    def ell_at(name, eps):
        c2 = dict(case)
        c2[name] = case[name] + eps * dirs[name]
        tm, om = oracle_lgssm_models(c2)
        _, ell = O.par_filtering(c2["ys"], O.MVNSqrt(c2["m0"], c2["L0"]), tm, om, O.extended, None, True)
        return ell

    names = ("F", "cholQ", "b", "H", "cholR", "c")
    ssm = tuple(tile(case[k]) for k in names)
    for name, direction in dirs.items():
        dssm = tuple(tile(direction) if k == name else np.zeros_like(s_) for k, s_ in zip(names, ssm))
        ell, dell = O.seq_loglikelihood_jvp(ssm, dssm, case["m0"], case["L0"], case["ys"])
        assert abs(ell - ell_at(name, 0.0)) <= 1e-10 * abs(ell)
        h = 1e-5
        fd = (ell_at(name, h) - ell_at(name, -h)) / (2 * h)
        assert abs(dell - fd) <= 1e-6 * max(1.0, abs(fd)), (name, dell, fd)
