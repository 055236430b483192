"""The product's per-thread fp64 algebra (csrc/psqrt_math.cuh, compiled for the CPU by
tests/hostcheck) and the chunked three-sweep algorithm, against the oracle.  CPU only."""
import ctypes

import numpy as np
import pytest

import parsmooth_np as O
from _cases import LLt, lgssm_case, oracle_from_ssm, rel_err, time_varying_case

P = ctypes.POINTER(ctypes.c_double)


def dp(a):
    return a.ctypes.data_as(P)


def _strides(arrs):
    ts = lambda a, base: 0 if a.ndim == base else int(np.prod(a.shape[1:]))
    return np.array([ts(arrs[0], 2), ts(arrs[1], 2), ts(arrs[2], 1), ts(arrs[3], 2), ts(arrs[4], 2), ts(arrs[5], 1)],
                    dtype=np.int64)


def run_pass(lib, case, K):
    T, ny = case["ys"].shape
    n = case["m0"].shape[0]
    arrs = [np.ascontiguousarray(case[k], dtype=np.float64) for k in ("F", "cholQ", "b", "H", "cholR", "c", "ys")]
    st = _strides(arrs)
    fm, fL = np.zeros((T + 1, n)), np.zeros((T + 1, n, n))
    sm, sL, ell = np.zeros_like(fm), np.zeros_like(fL), np.zeros(1)
    m0, L0 = np.ascontiguousarray(case["m0"]), np.ascontiguousarray(case["L0"])
    rc = lib.hc_pass(n, ny, ctypes.c_longlong(T), K, *[dp(a) for a in arrs],
                     st.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), dp(m0), dp(L0), dp(fm), dp(fL), dp(sm), dp(sL),
                     dp(ell))
    assert rc == 0
    return fm, fL, sm, sL, ell[0]


@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 1000, 7), (4, 2, 300, 1), (5, 2, 333, 4), (1, 1, 100, 3), (1, 3, 50, 2),
                                      (2, 3, 77, 5), (3, 3, 500, 16), (4, 2, 5, 2), (2, 1, 1, 1), (3, 1, 4097, 1)])
def test_pass_lgssm(hostcheck, n, ny, T, K):
    case = lgssm_case(n, ny, T, seed=100 * n + ny)
    fm, fL, sm, sL, ell = run_pass(hostcheck, case, K)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    assert rel_err(fm, ofm) < 1e-11 and rel_err(LLt(fL), LLt(ofc)) < 1e-11
    assert rel_err(sm, osm) < 1e-11 and rel_err(LLt(sL), LLt(osc)) < 1e-11
    assert abs(ell - oell) < 1e-11 * abs(oell)


@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 300, 6), (5, 2, 200, 5), (2, 2, 64, 1), (3, 3, 130, 3)])
def test_pass_time_varying(hostcheck, n, ny, T, K):
    case = time_varying_case(n, ny, T, seed=7 * n + ny)
    fm, fL, sm, sL, ell = run_pass(hostcheck, case, K)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    assert rel_err(fm, ofm) < 1e-11 and rel_err(LLt(fL), LLt(ofc)) < 1e-11
    assert rel_err(sm, osm) < 1e-11 and rel_err(LLt(sL), LLt(osc)) < 1e-11
    assert abs(ell - oell) < 1e-11 * abs(oell)


@pytest.mark.parametrize("dim_x", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("seed", [0, 42])
def test_operators(hostcheck, dim_x, seed):
    """known-answer generator of tests/test_parallel_operators.py:20-46,64-79."""
    np.random.seed(seed)
    tri = lambda: np.tril(np.random.rand(dim_x, dim_x))
    e1 = [np.random.randn(dim_x, dim_x), np.random.randn(dim_x), tri(), np.random.randn(dim_x), tri()]
    e2 = [np.random.randn(dim_x, dim_x), np.random.randn(dim_x), tri(), np.random.randn(dim_x), tri()]
    outs = [np.zeros_like(a) for a in e1]
    ins = (P * 10)(*[dp(a) for a in e1 + e2])
    oo = (P * 5)(*[dp(a) for a in outs])
    assert hostcheck.hc_filter_combine(dim_x, ins, oo) == 0
    ref = O.sqrt_filtering_operator(tuple(a[None] for a in e1), tuple(a[None] for a in e2))
    for i, fac in enumerate((0, 0, 1, 0, 1)):
        a, b = (LLt(outs[i]), LLt(ref[i][0])) if fac else (outs[i], ref[i][0])
        assert rel_err(a, b) < 1e-11
    s1, s2 = [e1[1], e1[0], e1[2]], [e2[1], e2[0], e2[2]]
    souts = [np.zeros_like(a) for a in s1]
    assert hostcheck.hc_smoothing_combine(dim_x, (P * 6)(*[dp(a) for a in s1 + s2]), (P * 3)(*[dp(a) for a in souts])) == 0
    sref = O.sqrt_smoothing_operator(tuple(a[None] for a in s1), tuple(a[None] for a in s2))
    assert rel_err(souts[0], sref[0][0]) < 1e-12 and rel_err(souts[1], sref[1][0]) < 1e-12
    assert rel_err(LLt(souts[2]), LLt(sref[2][0])) < 1e-12


@pytest.mark.parametrize("rows,cols", [(1, 1), (2, 5), (3, 3), (4, 8), (5, 243), (5, 2)])
def test_tria(hostcheck, rows, cols):
    rng = np.random.RandomState(rows + cols)
    A = rng.randn(rows, cols)
    L = np.zeros((rows, rows))
    assert hostcheck.hc_tria(rows, cols, dp(A), dp(L)) == 0
    assert rel_err(L @ L.T, A @ A.T) < 1e-12 and np.all(np.triu(L, 1) == 0)


@pytest.mark.parametrize("alpha", [1.0, -0.1, -1.0])
@pytest.mark.parametrize("n", [2, 3, 5])
def test_chol_update(hostcheck, n, alpha):
    rng = np.random.RandomState(n)
    L = np.tril(rng.rand(n, n)) + np.eye(n)
    V = 0.3 * rng.rand(3, n)
    ref = O.cholesky_update_many(L, V, alpha)
    got = L.copy()
    assert hostcheck.hc_chol_update(n, dp(got), dp(np.ascontiguousarray(V)), 3, ctypes.c_double(alpha)) == 0
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)
    bad = 0.1 * np.eye(n)
    ref_bad = O.cholesky_update_many(bad, 10 * V, -1.0)
    assert hostcheck.hc_chol_update(n, dp(bad), dp(np.ascontiguousarray(10 * V)), 3, ctypes.c_double(-1.0)) == 0
    assert np.all(np.isfinite(bad))
    np.testing.assert_allclose(bad, ref_bad, rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("n,ny", [(1, 1), (3, 2), (4, 2), (5, 2), (1, 3), (2, 3), (3, 3)])
def test_elements_and_loglik(hostcheck, n, ny):
    """parallel/_filtering.py:115-154 (element with the prior folded in, Z padding / tria branch, ell terms)."""
    T = 40
    case = lgssm_case(n, ny, T, seed=n * 13 + ny, triangular_prior=False)
    arrs = [np.ascontiguousarray(case[k], dtype=np.float64) for k in ("F", "cholQ", "b", "H", "cholR", "c", "ys")]
    st = _strides(arrs)
    A, U, Z = np.zeros((T, n, n)), np.zeros((T, n, n)), np.zeros((T, n, n))
    b, eta, terms = np.zeros((T, n)), np.zeros((T, n)), np.zeros(T)
    rng = np.random.RandomState(0)
    fm, fL = rng.randn(T, n), rng.randn(T, n, n)
    m0, L0 = np.ascontiguousarray(case["m0"]), np.ascontiguousarray(case["L0"])
    rc = hostcheck.hc_filter_elements(n, ny, ctypes.c_longlong(T), *[dp(a) for a in arrs],
                                      st.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), dp(m0), dp(L0), dp(A), dp(b),
                                      dp(U), dp(eta), dp(Z), dp(fm), dp(fL), dp(terms))
    assert rc == 0
    bc = lambda a, core: np.broadcast_to(a, (T,) + a.shape[-core:])
    lin = (bc(case["F"], 2), bc(case["cholQ"], 2), bc(case["b"], 1), bc(case["H"], 2), bc(case["cholR"], 2),
           bc(case["c"], 1))
    ms = np.concatenate([case["m0"][None], np.zeros((T - 1, n))])
    Ls = np.concatenate([case["L0"][None], np.zeros((T - 1, n, n))])
    oel = O.sqrt_filtering_elements(*lin, ms, Ls, case["ys"])
    for got, exp, fac in zip((A, b, U, eta, Z), oel, (0, 0, 1, 0, 1)):
        assert (rel_err(LLt(got), LLt(exp)) if fac else rel_err(got, exp)) < 1e-11
    if n > ny:
        assert np.all(Z[:, :, ny:] == 0)                    # zero padding, _filtering.py:141-142
    oterms = O.sqrt_loglikelihood_terms(*lin, fm, fL, case["ys"])
    assert rel_err(terms, oterms) < 1e-11


def _limit_case(kind, n, ny, T, seed):
    """The reference's limit cases (tests/test_sequential_filter.py:105-184, test_sequential_smoother.py:71-120)
    on the C1 recipe: noise-free / very informative / uninformative observations, (almost) deterministic or
    very noisy dynamics."""
    case = lgssm_case(n, ny, T, seed)
    if kind == "cholR=0":            # noise-free observation: the filtered covariance is singular
        case["cholR"] = np.zeros_like(case["cholR"])
    elif kind == "R*1e-6":           # test_filter_infinite_info
        case["cholR"] = 1e-3 * case["cholR"]
    elif kind == "R*1e12":           # test_filter_no_info
        case["cholR"] = 1e6 * case["cholR"]
    elif kind == "Q*1e12":           # test_smooth_one_standard_vs_sqrt_infinite_noise
        case["cholQ"] = 1e6 * case["cholQ"]
    elif kind == "Q*1e-12":          # (almost) test_smooth_one_standard_vs_sqrt_no_noise
        case["cholQ"] = 1e-6 * case["cholQ"]
    return case


@pytest.mark.parametrize("kind", ["cholR=0", "R*1e-6", "R*1e12", "Q*1e12", "Q*1e-12"])
@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 200, 7), (3, 1, 64, 4), (5, 2, 90, 9), (2, 2, 40, 5)])
def test_pass_limit_cases(hostcheck, kind, n, ny, T, K):
    if kind == "cholR=0" and ny >= n:
        pytest.skip("noise-free observation of the whole state: the predicted covariance is Q, upstream divides too")
    case = _limit_case(kind, n, ny, T, seed=3 * n + ny)
    fm, fL, sm, sL, ell = run_pass(hostcheck, case, K)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case, scan=O.sequential_scan if hasattr(O, "sequential_scan") else None)
    tol = 1e-9 if kind != "Q*1e-12" else 1e-7
    assert np.all(np.isfinite(sm)) and np.all(np.isfinite(sL))
    assert rel_err(fm, ofm) < tol and rel_err(LLt(fL), LLt(ofc)) < tol
    assert rel_err(sm, osm) < tol and rel_err(LLt(sL), LLt(osc)) < tol
    assert abs(ell - oell) < 1e-8 * abs(oell)
