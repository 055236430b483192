"""GPU parity tests: the CUDA path (through the C ABI of libpsqrt.so, via psqrt._lib / psqrt.methods)
against the NumPy oracle on the same seeded inputs.

Tolerances are BASELINE.json's: means 1e-9 relative, covariances compared as L L^T 1e-9 relative
(QR sign ambiguity, parsmooth/_utils.py:22-24), log-likelihood 1e-8 relative.
"""
import numpy as np
import pytest
import torch

import parsmooth_np as O
from _cases import LLt, lgssm_case, oracle_from_ssm, oracle_lgssm_models, rel_err, time_varying_case

pytestmark = pytest.mark.gpu

TOL = 1e-9
TOL_ELL = 1e-8


def _dev():
    return torch.device("cuda", 0)


def _g(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=_dev())


def _ssm(case):
    from psqrt._lib import LinearizedSSM
    return LinearizedSSM(*[_g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])


def _check_traj(name, m, L, om, oL, tol=TOL):
    em, eL = rel_err(m.cpu().numpy(), om), rel_err(LLt(L.cpu().numpy()), LLt(oL))
    assert em < tol and eL < tol, f"{name}: mean err {em:.3e}, LL^T err {eL:.3e}"
    return em, eL


def test_library_loaded_from_tree():
    from psqrt import _lib
    lib = _lib.load()
    assert lib.psqrt_version() == 100
    assert "sqrt-parallel-smoothers_b200" in _lib.lib_path()


@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 1000, 0), (4, 2, 1000, 7), (4, 2, 1000, 1), (5, 2, 333, 4), (1, 1, 100, 3),
                                      (1, 3, 50, 2), (2, 3, 77, 5), (3, 3, 500, 16), (4, 2, 5, 2), (2, 1, 1, 1),
                                      (6, 4, 257, 3), (8, 4, 130, 0), (3, 1, 31, 1), (3, 1, 32, 1), (3, 1, 33, 1),
                                      (4, 2, 4097, 1), (4, 2, 12289, 3),
                                      # several groups in the mid scans (half-warp combines, psqrt_coop.cuh);
                                      # T = 40000, K = 1: 40 groups, i.e. two group totals per half-warp at level C
                                      (5, 2, 3000, 1), (3, 3, 2500, 1), (2, 1, 2100, 1), (1, 1, 1500, 1),
                                      (4, 2, 40000, 1), (5, 2, 35000, 1), (6, 4, 1100, 1), (8, 4, 1100, 1), (8, 4, 33000, 1),
                                      (6, 3, 34000, 1)])
def test_pass_vs_oracle_lgssm(n, ny, T, K):
    """Whole filter + smoother pass + ell on a time-invariant LGSSM, every chunk length regime
    (ragged tails, single chunk, > 1 CTA, > 32 warps in the mid scan)."""
    from psqrt import _lib
    case = lgssm_case(n, ny, T, seed=100 * n + ny)
    fm, fL, sm, sL, ell = _lib.filter_smoother(_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]),
                                               smooth=True, loglik=True, chunk_len=K)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    _check_traj("filtered", fm, fL, ofm, ofc)
    _check_traj("smoothed", sm, sL, osm, osc)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)


@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 300, 0), (5, 2, 1000, 5), (2, 2, 64, 1), (3, 4, 200, 3)])
def test_pass_vs_oracle_time_varying(n, ny, T, K):
    """Per-step (F, cholQ, b, H, cholR, c): what a nonlinear model's linearisation feeds the scan."""
    from psqrt import _lib
    case = time_varying_case(n, ny, T, seed=7 * n + ny)
    fm, fL, sm, sL, ell = _lib.filter_smoother(_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]),
                                               smooth=True, loglik=True, chunk_len=K)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    _check_traj("filtered", fm, fL, ofm, ofc)
    _check_traj("smoothed", sm, sL, osm, osc)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)


@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 1000, 0), (5, 2, 333, 4), (1, 1, 100, 3), (2, 3, 77, 5), (8, 4, 130, 0),
                                      (6, 4, 257, 3), (3, 1, 33, 1)])
def test_pass_by_value_model(n, ny, T, K):
    """Time-invariant model carried by value in the kernel parameters (host mirrors given): same
    results as the pointer path and as the oracle; also through the public API with NumPy inputs."""
    import psqrt
    from psqrt import _lib
    from psqrt._lib import LinearizedSSM
    from psqrt.models import lgssm
    case = lgssm_case(n, ny, T, seed=100 * n + ny)
    names = ("F", "cholQ", "b", "H", "cholR", "c")
    ssm = LinearizedSSM(*[_g(case[k]) for k in names], host={k: case[k] for k in names})
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, _g(case["ys"]), _g(case["m0"]), _g(case["L0"]), smooth=True,
                                               loglik=True, chunk_len=K)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    _check_traj("filtered", fm, fL, ofm, ofc)
    _check_traj("smoothed", sm, sL, osm, osc)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)
    sm2, sL2 = _lib.smoother(LinearizedSSM(ssm.F, ssm.cholQ, ssm.b, host=ssm.host), fm, fL, chunk_len=K)
    _check_traj("smoother-only", sm2, sL2, osm, osc)
    # public API, NumPy inputs: mirrors are attached automatically
    tm = psqrt.FunctionalModel(lgssm.transition_function(case["F"]), psqrt.MVNSqrt(case["b"], case["cholQ"]))
    om = psqrt.FunctionalModel(lgssm.observation_function(case["H"]), psqrt.MVNSqrt(case["c"], case["cholR"]))
    res = psqrt.filter_smoother(case["ys"], psqrt.MVNSqrt(case["m0"], case["L0"]), tm, om, psqrt.linearization.extended)
    _check_traj("api", res.mean, res.chol, osm, osc)


def test_pass_vs_sequential_oracle():
    """Second oracle: the reference's sequential sqrt filter / smoother (sequential/_filtering.py, _smoothing.py)."""
    from psqrt import _lib
    case = lgssm_case(4, 2, 400, seed=3)
    tm, om = oracle_lgssm_models(case)
    nominal = O._default_nominal(case["m0"], 401)
    fs, ells = O.seq_filtering(case["ys"], O.MVNSqrt(case["m0"], case["L0"]), tm, om, O.extended, nominal, True)
    ss = O.seq_smoothing(tm, fs, O.extended, nominal)
    fm, fL, sm, sL, ell = _lib.filter_smoother(_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]),
                                               smooth=True, loglik=True)
    _check_traj("filtered", fm, fL, fs.mean, fs.chol)
    _check_traj("smoothed", sm, sL, ss.mean, ss.chol)
    assert abs(ell.item() - ells) <= TOL_ELL * abs(ells)


@pytest.mark.parametrize("n,ny,R,T", [(4, 2, 3, 1000), (5, 2, 2, 301), (8, 4, 4, 260), (1, 1, 2, 64), (2, 3, 8, 50)])
def test_staged_calls_fake_ranks(n, ny, R, T):
    """The time-sharded path (psqrt/dist.py steps 1-5) with R fake ranks on one GPU: shard totals,
    carry folding, carry application, suffix carries -- equal to the unsharded oracle."""
    from psqrt import _lib
    from psqrt import dist as pdist
    case = lgssm_case(n, ny, T, seed=9 * n + ny)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    ssm = _ssm(case)
    ys = _g(case["ys"])
    m0, L0 = _g(case["m0"])[None], _g(case["L0"])[None]
    spans = [pdist.shard_bounds(T, R, r) for r in range(R)]
    try:
        totals = []
        for r, (t0, t1) in enumerate(spans):
            _lib.WS_SLOT = 100 + r
            totals.append(_lib.filter_reduce(ssm, ys[t0:t1][None].contiguous(), n))
        totals = torch.stack(totals)
        outs, payloads, ell_sum = [], [], 0.0
        for r, (t0, t1) in enumerate(spans):
            _lib.WS_SLOT = 100 + r
            cm, cL = _lib.carry_filter(totals, r, m0, L0)
            fm, fL, ell, stot = _lib.filter_apply(ssm, ys[t0:t1][None].contiguous(), cm, cL, smooth=True, loglik=True)
            ell_sum += ell.item()
            outs.append((fm, fL))
            payloads.append(stot)
            _check_traj(f"filtered shard {r}", fm[0], fL[0], ofm[t0:t1 + 1], ofc[t0:t1 + 1])
        assert abs(ell_sum - oell) <= TOL_ELL * abs(oell)
        stotals = torch.stack(payloads)
        mT, LT = outs[-1][0][:, -1].contiguous(), outs[-1][1][:, -1].contiguous()
        for r, (t0, t1) in enumerate(spans):
            _lib.WS_SLOT = 100 + r
            scm, scL = _lib.carry_smoother(stotals, r, R, mT, LT)
            sm, sL = _lib.smoother_apply(ssm, outs[r][0], outs[r][1], scm, scL, write_terminal=True)
            _check_traj(f"smoothed shard {r}", sm[0], sL[0], osm[t0:t1 + 1], osc[t0:t1 + 1])
    finally:
        _lib.WS_SLOT = 0


def test_batched_pass():
    """batch axis: independent sequences sharing the model (config 5 shape)."""
    from psqrt import _lib
    from psqrt._lib import LinearizedSSM
    n, ny, T, B = 4, 2, 257, 5
    cases = [lgssm_case(n, ny, T, seed=50) for _ in range(B)]
    rng = np.random.RandomState(0)
    ys = np.stack([c["ys"] + 0.1 * rng.randn(T, ny) for c in cases])
    m0 = np.stack([c["m0"] + 0.1 * i for i, c in enumerate(cases)])
    L0 = np.stack([c["L0"] for c in cases])
    ssm = LinearizedSSM(*[_g(cases[0][k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, _g(ys), _g(m0), _g(L0), smooth=True, loglik=True)
    for i in range(B):
        c = dict(cases[0], ys=ys[i], m0=m0[i], L0=L0[i])
        ofm, ofc, osm, osc, oell = oracle_from_ssm(c)
        _check_traj(f"filtered[{i}]", fm[i], fL[i], ofm, ofc)
        _check_traj(f"smoothed[{i}]", sm[i], sL[i], osm, osc)
        assert abs(ell[i].item() - oell) <= TOL_ELL * abs(oell)


@pytest.mark.parametrize("dim_x", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("seed", [0, 42])
def test_filtering_operator(dim_x, seed):
    """Known-answer generator of the reference's tests/test_parallel_operators.py:17-58, at 1e-9."""
    from psqrt import _lib
    np.random.seed(seed)
    tri = lambda: np.tril(np.random.rand(dim_x, dim_x))
    A1, A2 = np.random.randn(dim_x, dim_x), np.random.randn(dim_x, dim_x)
    b1, b2 = np.random.randn(dim_x), np.random.randn(dim_x)
    U1, U2 = tri(), tri()
    eta1, eta2 = np.random.randn(dim_x), np.random.randn(dim_x)
    Z1, Z2 = tri(), tri()
    e1, e2 = (A1, b1, U1, eta1, Z1), (A2, b2, U2, eta2, Z2)
    out = _lib.filter_combine([_g(a) for a in e1], [_g(a) for a in e2])
    A, b, U, eta, Z = [o.cpu().numpy() for o in out]
    oA, ob, oU, oeta, oZ = O.sqrt_filtering_operator(tuple(a[None] for a in e1), tuple(a[None] for a in e2))
    sA, sb, sC, seta, sJ = O.standard_filtering_operator((A1, b1, U1 @ U1.T, eta1, Z1 @ Z1.T),
                                                         (A2, b2, U2 @ U2.T, eta2, Z2 @ Z2.T))
    for got, exp in ((A, oA[0]), (b, ob[0]), (eta, oeta[0]), (LLt(U), LLt(oU[0])), (LLt(Z), LLt(oZ[0]))):
        assert rel_err(got, exp) < TOL
    for got, exp in ((A, sA), (b, sb), (eta, seta), (LLt(U), sC), (LLt(Z), sJ)):   # sqrt == standard
        assert rel_err(got, exp) < 1e-7


@pytest.mark.parametrize("dim_x", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("seed", [0, 42])
def test_smoothing_operator(dim_x, seed):
    """tests/test_parallel_operators.py:61-89."""
    from psqrt import _lib
    np.random.seed(seed)
    g1, g2 = np.random.randn(dim_x), np.random.randn(dim_x)
    E1, E2 = np.random.randn(dim_x, dim_x), np.random.randn(dim_x, dim_x)
    D1, D2 = np.tril(np.random.rand(dim_x, dim_x)), np.tril(np.random.rand(dim_x, dim_x))
    out = _lib.smoother_combine([_g(a) for a in (g1, E1, D1)], [_g(a) for a in (g2, E2, D2)])
    g, E, D = [o.cpu().numpy() for o in out]
    og, oE, oD = O.sqrt_smoothing_operator((g1[None], E1[None], D1[None]), (g2[None], E2[None], D2[None]))
    assert rel_err(g, og[0]) < TOL and rel_err(E, oE[0]) < TOL and rel_err(LLt(D), LLt(oD[0])) < TOL
    sg, sE, sL = O.standard_smoothing_operator((g1, E1, D1 @ D1.T), (g2, E2, D2 @ D2.T))
    assert rel_err(LLt(D), sL) < 1e-7


@pytest.mark.parametrize("n,ny", [(1, 1), (2, 1), (3, 3), (4, 2), (5, 2), (1, 3), (2, 3), (8, 4)])
def test_elements_and_scans(n, ny):
    """The reference's own seams: element construction (parallel/_filtering.py:100-146,
    _smoothing.py:47-85), the two associative scans, and the log-likelihood terms."""
    from psqrt import _lib
    T = 203
    case = lgssm_case(n, ny, T, seed=11 * n + ny, triangular_prior=False)   # dense prior factor allowed here
    ssm = _ssm(case)
    ys = _g(case["ys"])
    A, b, U, eta, Z = _lib.filter_elements(ssm, ys, _g(case["m0"]), _g(case["L0"]))
    bc = lambda a, core: np.broadcast_to(a, (T,) + a.shape[-core:])
    lin = (bc(case["F"], 2), bc(case["cholQ"], 2), bc(case["b"], 1), bc(case["H"], 2), bc(case["cholR"], 2),
           bc(case["c"], 1))
    ms = np.concatenate([case["m0"][None], np.zeros((T - 1, n))])
    Ls = np.concatenate([case["L0"][None], np.zeros((T - 1, n, n))])
    oel = O.sqrt_filtering_elements(*lin, ms, Ls, case["ys"])
    for got, exp, is_factor in zip((A, b, U, eta, Z), oel, (0, 0, 1, 0, 1)):
        got = got.cpu().numpy()
        assert rel_err(LLt(got), LLt(exp)) < TOL if is_factor else rel_err(got, exp) < TOL
    means, chols = _lib.filter_scan(A, b, U, eta, Z, chunk_len=3)
    _, ofm, ofc, _, _ = O.associative_scan(O.sqrt_filtering_operator, oel)
    _check_traj("filter_scan", means, chols, ofm, ofc)
    fm = torch.cat([_g(case["m0"])[None], means])
    fL = torch.cat([_lib.tria(_g(case["L0"]))[None], chols])
    terms = _lib.loglik_terms(ssm, ys, fm, fL).cpu().numpy()
    ofm1 = np.concatenate([case["m0"][None], ofm])
    ofc1 = np.concatenate([case["L0"][None], ofc])
    oterms = O.sqrt_loglikelihood_terms(*lin, ofm1[:-1], ofc1[:-1], case["ys"])
    assert rel_err(terms, oterms) < TOL_ELL
    g, E, D = _lib.smoother_elements(ssm, fm, fL)
    og, oE, oD = O.sqrt_smoothing_elements(lin[0], lin[1], lin[2], ofm1[:-1], ofc1[:-1])
    assert rel_err(g[:-1].cpu().numpy(), og) < TOL and rel_err(E[:-1].cpu().numpy(), oE) < TOL
    assert rel_err(LLt(D[:-1].cpu().numpy()), LLt(oD)) < TOL
    assert torch.equal(g[-1], fm[-1]) and float(E[-1].abs().max()) == 0.0
    sm, sL = _lib.smoother_scan(g, E, D, chunk_len=2)
    og = np.concatenate([og, ofm1[-1:]])
    oE = np.concatenate([oE, np.zeros((1, n, n))])
    oD = np.concatenate([oD, ofc1[-1:]])
    osm, _, osc = O.associative_scan(O.sqrt_smoothing_operator, (og, oE, oD), reverse=True)
    _check_traj("smoother_scan", sm, sL, osm, osc)


@pytest.mark.parametrize("rows,cols", [(1, 1), (2, 4), (3, 3), (4, 8), (5, 10), (5, 243), (6, 7), (8, 16), (2, 9)])
def test_tria(rows, cols):
    from psqrt import _lib
    rng = np.random.RandomState(rows * 100 + cols)
    A = rng.randn(37, rows, cols)
    L = _lib.tria(_g(A)).cpu().numpy()
    assert np.all(np.triu(L, 1) == 0)
    assert rel_err(LLt(L), A @ np.swapaxes(A, -1, -2)) < 1e-12
    assert rel_err(LLt(L), LLt(O.tria(A))) < 1e-12


@pytest.mark.parametrize("multiplier", [1.0, -0.1])
@pytest.mark.parametrize("dim_x", [2, 3, 5, 8])
@pytest.mark.parametrize("seed", [0, 42, 666])
def test_cholesky_update(multiplier, dim_x, seed):
    """tests/test_math_utils.py:19-54 (update and update_many), plus the non-finite -> 0 guard."""
    from psqrt import _lib
    np.random.seed(seed)
    B = 3
    cholQ = np.tril(np.random.rand(dim_x, dim_x))
    v = np.random.rand(B, dim_x)
    expected = cholQ @ cholQ.T + multiplier * sum(v[k, :, None] @ v[k, None, :] for k in range(B))
    got = _lib.chol_update_many(_g(cholQ), _g(v), multiplier).cpu().numpy()
    ref = O.cholesky_update_many(cholQ, v, multiplier)
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-300, equal_nan=True)
    if min(np.linalg.eigvals(expected).real) > 1e-6:
        np.testing.assert_allclose(got @ got.T, expected, rtol=1e-4)
    # downdate past positive-definiteness: NaNs are replaced by zeros exactly like _utils.py:80
    bad = _lib.chol_update_many(_g(0.1 * cholQ), _g(10 * v), -1.0).cpu().numpy()
    ref_bad = O.cholesky_update_many(0.1 * cholQ, 10 * v, -1.0)
    assert np.all(np.isfinite(bad))
    np.testing.assert_allclose(bad, ref_bad, rtol=1e-10, atol=1e-300)


def _torch_models(case):
    import psqrt
    from psqrt.models import lgssm
    tm = psqrt.FunctionalModel(lgssm.transition_function(case["F"]), psqrt.MVNSqrt(_g(case["b"]), _g(case["cholQ"])))
    om = psqrt.FunctionalModel(lgssm.observation_function(case["H"]), psqrt.MVNSqrt(_g(case["c"]), _g(case["cholR"])))
    return tm, om


@pytest.mark.parametrize("dim_x,dim_y", [(1, 1), (2, 1), (3, 2), (2, 3), (4, 2)])
@pytest.mark.parametrize("lin_name", ["extended", "cubature", "gauss_hermite", "unscented"])
def test_methods_api_lgssm(dim_x, dim_y, lin_name):
    """methods.filtering / smoothing / filter_smoother / iterated_smoothing through the public API,
    mirroring tests/test_parallel_filter.py:80-122, test_parallel_smoother.py:61-99 and
    test_iterated_smoother.py:31-69 (explicit random nominal trajectory, linear model)."""
    import psqrt
    T = 25
    case = lgssm_case(dim_x, dim_y, T, seed=dim_x * 10 + dim_y)
    rng = np.random.RandomState(5)
    nominal_np = O.MVNSqrt(rng.randn(T + 1, dim_x), np.repeat(np.eye(dim_x)[None], T + 1, 0))
    nominal = psqrt.MVNSqrt(_g(nominal_np.mean), _g(nominal_np.chol))
    lin = getattr(psqrt.linearization, lin_name)
    olin = getattr(O, lin_name)
    tm, om = _torch_models(case)
    otm, oom = oracle_lgssm_models(case)
    x0 = psqrt.MVNSqrt(_g(case["m0"]), _g(case["L0"]))
    ox0 = O.MVNSqrt(case["m0"], case["L0"])
    filt, ell = psqrt.filtering(case["ys"], x0, tm, om, lin, nominal, True, return_loglikelihood=True)
    ofilt, oell = O.par_filtering(case["ys"], ox0, otm, oom, olin, nominal_np, True)
    _check_traj("filtering", filt.mean, filt.chol, ofilt.mean, ofilt.chol)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)
    sfilt, sell = O.seq_filtering(case["ys"], ox0, otm, oom, olin, nominal_np, True)
    _check_traj("filtering-vs-seq", filt.mean, filt.chol, sfilt.mean, sfilt.chol)
    assert torch.equal(filt.mean[0], x0.mean) and torch.equal(filt.chol[0], x0.chol)
    smo = psqrt.smoothing(tm, filt, lin, nominal, True)
    osmo = O.par_smoothing(otm, ofilt, olin, nominal_np)
    _check_traj("smoothing", smo.mean, smo.chol, osmo.mean, osmo.chol)
    fs = psqrt.filter_smoother(case["ys"], x0, tm, om, lin, nominal, True)
    _check_traj("filter_smoother", fs.mean, fs.chol, osmo.mean, osmo.chol)
    it = psqrt.iterated_smoothing(case["ys"], x0, tm, om, lin, nominal, True, criterion=lambda i, *_: i < 5)
    oseq = O.seq_smoothing(otm, O.seq_filtering(case["ys"], ox0, otm, oom, olin, nominal_np), olin, nominal_np)
    _check_traj("iterated", it.mean, it.chol, oseq.mean, oseq.chol, tol=1e-8)
    # default nominal (None) and default criterion
    it2, ell2 = psqrt.iterated_smoothing(case["ys"], x0, tm, om, lin, None, True, return_loglikelihood=True)
    oit2, oell2 = O.iterated_smoothing(case["ys"], ox0, otm, oom, olin, None, True, return_loglikelihood=True)
    _check_traj("iterated-default", it2.mean, it2.chol, oit2.mean, oit2.chol, tol=1e-8)
    assert abs(ell2.item() - oell2) <= TOL_ELL * abs(oell2)


@pytest.mark.parametrize("lin_name", ["extended", "cubature", "gauss_hermite", "unscented"])
def test_builtin_linearization_kernels(lin_name):
    """psqrt_linearize_builtin (csrc/psqrt_models.cu) against the oracle's linearization
    (linearization/_extended.py, _sigma_points.py, _cubature.py, _gh.py) at random nominal points,
    including a |w| < 1e-6 turn rate (the lax.cond branch of bearings_utils.py:24-37)."""
    import psqrt
    from psqrt.models import bearings, population
    rng = np.random.RandomState(11)
    T = 301
    nm = rng.randn(T, 5) * np.array([2.0, 2.0, 3.0, 3.0, 1.0])
    nm[7, 4] = 1e-8
    nL = 0.3 * (np.tril(rng.rand(T, 5, 5)) + np.eye(5))
    Q, R, obs_f, trans_f = bearings.make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    _, _, oobs, otrans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    cQ, cR = np.linalg.cholesky(Q), np.linalg.cholesky(R)
    mq, mr = 0.1 * rng.randn(5), 0.1 * rng.randn(2)
    lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
    x = psqrt.MVNSqrt(_g(nm), _g(nL))
    ox = O.MVNSqrt(nm, nL)
    for f, of, m_q, c_q in ((trans_f, otrans, mq, cQ), (obs_f, oobs, mr, cR)):
        assert hasattr(f, "_psqrt_builtin")
        F, ch, b = lin(psqrt.FunctionalModel(f, psqrt.MVNSqrt(_g(m_q), _g(c_q))), x)
        oF, och, ob = olin(O.FunctionalModel(of, O.MVNSqrt(m_q, c_q)), ox)
        assert rel_err(F.cpu().numpy(), oF) < TOL and rel_err(b.cpu().numpy(), ob) < TOL
        chn = np.broadcast_to(ch.cpu().numpy(), och.shape)
        # the residual covariance Phi + Q - F P F^T is a difference of nearly equal matrices for SLR
        # (_sigma_points.py:77-78): the meaningful scale of its error is that of the terms, F P F^T
        FL = oF @ nL
        scale = max(float(np.abs(LLt(och)).max()), float(np.abs(FL @ np.swapaxes(FL, -1, -2)).max()))
        err = float(np.abs(LLt(chn) - LLt(och)).max()) / scale
        assert err < TOL, f"{lin_name}: residual covariance error {err:.3e} (scale {scale:.3e})"
    pm = np.log(7.0) + 0.5 * rng.randn(T, 1)
    pL = 0.05 + 0.5 * rng.rand(T, 1, 1)
    tmod, omod = population.make_parameters(10.0, np.array([[0.09]]))
    otmod, oomod = O.population_model(10.0, np.array([[0.09]]))
    for mod, omodel in ((tmod, otmod), (omod, oomod)):
        F, ch, b = lin(mod, psqrt.MVNSqrt(_g(pm), _g(pL)))
        oF, och, ob = olin(omodel, O.MVNSqrt(pm, pL))
        assert rel_err(F.cpu().numpy(), oF) < TOL and rel_err(b.cpu().numpy(), ob) < TOL
        assert rel_err(LLt(ch.cpu().numpy()), LLt(och)) < TOL


def test_non_triangular_noise_factors():
    """cholQ / cholR given as arbitrary (non-triangular) square roots, like the reference allows: the host
    layer triangularises cholQ before the kernels read its lower triangle; results depend on Q Q^T only."""
    import psqrt
    from psqrt.models import lgssm
    T = 150
    case = lgssm_case(4, 2, T, seed=21)
    rng = np.random.RandomState(2)
    Oq, _ = np.linalg.qr(rng.randn(4, 4))
    Or, _ = np.linalg.qr(rng.randn(2, 2))
    sqQ, sqR = case["cholQ"] @ Oq, case["cholR"] @ Or                 # same Q, R; dense factors
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    for dev_inputs in (False, True):
        conv = _g if dev_inputs else (lambda a: a)
        tm = psqrt.FunctionalModel(lgssm.transition_function(case["F"]), psqrt.MVNSqrt(conv(case["b"]), conv(sqQ)))
        om = psqrt.FunctionalModel(lgssm.observation_function(case["H"]), psqrt.MVNSqrt(conv(case["c"]), conv(sqR)))
        x0 = psqrt.MVNSqrt(conv(case["m0"]), conv(case["L0"]))
        filt, ell = psqrt.filtering(case["ys"], x0, tm, om, psqrt.linearization.extended, None, True, True)
        smo = psqrt.filter_smoother(case["ys"], x0, tm, om, psqrt.linearization.extended)
        _check_traj("filtered", filt.mean, filt.chol, ofm, ofc)
        _check_traj("smoothed", smo.mean, smo.chol, osm, osc)
        assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)


def _bearings_setup(T, seed=0):
    from psqrt.models import bearings
    s1, s2, r, dt, qc, qw = np.array([-1.5, 0.5]), np.array([1.0, 1.0]), 0.5, 0.01, 0.01, 0.1
    _, _, ys = bearings.get_data(np.array([0.1, 0.2, 1.0, 0.0]), dt, r, T, s1, s2, random_state=seed)
    ys = ys.astype(np.float64)
    Q, R, obs_f, trans_f = bearings.make_parameters(qc, qw, r, dt, s1, s2)
    oQ, oR, oobs, otrans = O.bearings_make_parameters(qc, qw, r, dt, s1, s2)
    cholQ, cholR = np.linalg.cholesky(Q), np.linalg.cholesky(R)
    m0 = np.array([-4.0, -1.0, 2.0, 7.0, 3.0])
    return ys, m0, cholQ, cholR, (obs_f, trans_f), (oobs, otrans)


@pytest.mark.parametrize("lin_name", ["extended", "cubature", "gauss_hermite", "unscented"])
def test_bearings_iterated_smoother(lin_name):
    """Config 2/3 shape at test size: coordinated-turn + 2 bearings, nx=5, iterated sqrt parallel
    smoother from the notebooks' initial nominal, 10 iterations, + log-likelihood."""
    import psqrt
    T = 500
    ys, m0, cholQ, cholR, (obs_f, trans_f), (oobs, otrans) = _bearings_setup(T)
    lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
    x0 = psqrt.MVNSqrt(_g(m0), _g(np.eye(5)))
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(_g(np.zeros(5)), _g(cholQ)))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(_g(np.zeros(2)), _g(cholR)))
    otm = O.FunctionalModel(otrans, O.MVNSqrt(np.zeros(5), cholQ))
    oom = O.FunctionalModel(oobs, O.MVNSqrt(np.zeros(2), cholR))
    nom_m = np.tile(np.array([-1.0, -1.0, 6.0, 4.0, 2.0]), (T + 1, 1))
    nom_L = np.repeat(np.eye(5)[None], T + 1, 0)
    n_iter = 10
    res, ell = psqrt.iterated_smoothing(ys, x0, tm, om, lin, psqrt.MVNSqrt(_g(nom_m), _g(nom_L)), True,
                                        criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)
    ores, oell = O.iterated_smoothing(ys, O.MVNSqrt(m0, np.eye(5)), otm, oom, olin, O.MVNSqrt(nom_m, nom_L), True,
                                      criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)
    # 10 nonlinear iterations amplify rounding differences; the contract is per pass, so also check one pass
    _check_traj("iterated", res.mean, res.chol, ores.mean, ores.chol, tol=1e-7)
    assert abs(ell.item() - oell) <= 1e-7 * abs(oell)
    one = psqrt.filter_smoother(ys, x0, tm, om, lin, psqrt.MVNSqrt(_g(ores.mean), _g(ores.chol)), True)
    oone = O.filter_smoother(ys, O.MVNSqrt(m0, np.eye(5)), otm, oom, olin, ores, True)
    _check_traj("one pass at the oracle's nominal", one.mean, one.chol, oone.mean, oone.chol)


@pytest.mark.parametrize("lin_name", ["extended", "cubature"])
def test_bearings_batched_runs(lin_name):
    """Config 5 shape at test size: independent bearings-only runs smoothed as ONE batch
    (psqrt.dist.iterated_smoothing_batched) give, run by run, what psqrt.iterated_smoothing gives for each
    run on its own and what the oracle gives for the first run.  (A single sequence runs its smoothing mid scan
    inside K3 with the per-thread combine, a batch runs it as its own kernel with the sub-warp combine: equal up to
    rounding, which four iterations of a nonlinear smoother amplify -- hence 1e-9, not bit equality.)"""
    import psqrt
    from psqrt.dist import iterated_smoothing_batched
    T, B, n_iter = 300, 5, 4
    sets = [_bearings_setup(T, seed=k) for k in range(B)]
    ys, m0, cholQ, cholR, (obs_f, trans_f), (oobs, otrans) = sets[0]
    lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
    x0 = psqrt.MVNSqrt(_g(m0), _g(np.eye(5)))
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(_g(np.zeros(5)), _g(cholQ)))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(_g(np.zeros(2)), _g(cholR)))
    nom_m = np.tile(np.array([-1.0, -1.0, 6.0, 4.0, 2.0]), (T + 1, 1))
    nom_L = np.repeat(np.eye(5)[None], T + 1, 0)
    nominal = psqrt.MVNSqrt(_g(nom_m), _g(nom_L))
    ys_b = np.stack([s_[0] for s_ in sets])
    res, ell = iterated_smoothing_batched(ys_b, x0, tm, om, lin, nominal, n_iter=n_iter, return_loglikelihood=True)
    assert res.mean.shape == (B, T + 1, 5) and ell.shape == (B,)
    for k in range(B):
        one, ell1 = psqrt.iterated_smoothing(sets[k][0], x0, tm, om, lin, nominal, True,
                                             criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)
        _check_traj(f"run {k}", res.mean[k], res.chol[k], one.mean.cpu().numpy(), one.chol.cpu().numpy(), tol=1e-9)
        assert abs(ell[k].item() - ell1.item()) <= 1e-9 * abs(ell1.item())
    otm = O.FunctionalModel(otrans, O.MVNSqrt(np.zeros(5), cholQ))
    oom = O.FunctionalModel(oobs, O.MVNSqrt(np.zeros(2), cholR))
    ores, oell = O.iterated_smoothing(ys, O.MVNSqrt(m0, np.eye(5)), otm, oom, olin, O.MVNSqrt(nom_m, nom_L), True,
                                      criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)
    _check_traj("run 0 vs oracle", res.mean[0], res.chol[0], ores.mean, ores.chol, tol=1e-8)
    assert abs(ell[0].item() - oell) <= 1e-8 * abs(oell)
    # default nominal (zeros / identity) and per-sequence priors
    x0b = psqrt.MVNSqrt(_g(np.tile(m0, (B, 1))), _g(np.repeat(np.eye(5)[None], B, 0)))
    res2 = iterated_smoothing_batched(ys_b, x0b, tm, om, lin, None, n_iter=2)
    one2 = psqrt.iterated_smoothing(sets[1][0], x0, tm, om, lin, None, True, criterion=lambda i, *_: i < 2)
    _check_traj("default nominal", res2.mean[1], res2.chol[1], one2.mean.cpu().numpy(), one2.chol.cpu().numpy(), tol=1e-9)


def test_population_model():
    """Config 5b: Ricker / Poisson conditional-moments model, nx = ny = 1, sqrt path
    (notebooks/population_model.py; experiment-poisson.ipynb: lam = 10, Q = 0.09, x0 = log 7)."""
    import psqrt
    from psqrt.models import population
    T = 129
    lam, Q = 10.0, np.array([[0.09]])
    _, ys = population.get_data(np.log(7.0), T, Q, lam, random_state=0)
    tmod, omod = population.make_parameters(lam, Q)
    otmod, oomod = O.population_model(lam, Q)
    x0 = psqrt.MVNSqrt(_g(np.array([np.log(7.0)])), _g(np.eye(1)))
    ox0 = O.MVNSqrt(np.array([np.log(7.0)]), np.eye(1))
    nom_m = np.full((T + 1, 1), np.log(7.0))
    nom_L = np.repeat(np.eye(1)[None], T + 1, 0)
    for lin_name in ("extended", "cubature", "gauss_hermite"):
        lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
        res = psqrt.iterated_smoothing(ys, x0, tmod, omod, lin, psqrt.MVNSqrt(_g(nom_m), _g(nom_L)), True,
                                       criterion=lambda i, *_: i < 5)
        ores = O.iterated_smoothing(ys, ox0, otmod, oomod, olin, O.MVNSqrt(nom_m, nom_L), True,
                                    criterion=lambda i, *_: i < 5)
        _check_traj(f"population/{lin_name}", res.mean, res.chol, ores.mean, ores.chol, tol=1e-8)


def test_user_supplied_torch_model():
    """A model the library does not know: plain torch callables differentiated with torch.func
    (the 'user-supplied models still linearise on the host side and feed the CUDA scan' seam)."""
    import psqrt
    T = 200
    rng = np.random.RandomState(1)
    ys = rng.randn(T, 1)

    def f(x):
        return torch.stack([x[0] + 0.1 * torch.sin(x[1]), 0.9 * x[1] + 0.05 * torch.cos(x[0])])

    def h(x):
        return torch.stack([torch.sqrt(1.0 + x[0] ** 2 + x[1] ** 2)])

    def f_np(x):
        return np.stack([x[..., 0] + 0.1 * np.sin(x[..., 1]), 0.9 * x[..., 1] + 0.05 * np.cos(x[..., 0])], -1)

    def f_jac(x):
        J = np.zeros(x.shape[:-1] + (2, 2))
        J[..., 0, 0], J[..., 0, 1] = 1.0, 0.1 * np.cos(x[..., 1])
        J[..., 1, 0], J[..., 1, 1] = -0.05 * np.sin(x[..., 0]), 0.9
        return J

    f_np.jac = f_jac

    def h_np(x):
        return np.sqrt(1.0 + x[..., 0] ** 2 + x[..., 1] ** 2)[..., None]

    h_np.jac = lambda x: (x / np.sqrt(1.0 + x[..., 0] ** 2 + x[..., 1] ** 2)[..., None])[..., None, :]
    cQ, cR = 0.3 * np.eye(2), 0.5 * np.eye(1)
    x0 = psqrt.MVNSqrt(_g(np.array([0.5, -0.2])), _g(np.eye(2)))
    tm = psqrt.FunctionalModel(f, psqrt.MVNSqrt(_g(np.zeros(2)), _g(cQ)))
    om = psqrt.FunctionalModel(h, psqrt.MVNSqrt(_g(np.zeros(1)), _g(cR)))
    otm = O.FunctionalModel(f_np, O.MVNSqrt(np.zeros(2), cQ))
    oom = O.FunctionalModel(h_np, O.MVNSqrt(np.zeros(1), cR))
    for lin_name in ("extended", "cubature"):
        lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
        res = psqrt.iterated_smoothing(ys, x0, tm, om, lin, None, True, criterion=lambda i, *_: i < 3)
        ores = O.iterated_smoothing(ys, O.MVNSqrt(np.array([0.5, -0.2]), np.eye(2)), otm, oom, olin, None, True,
                                    criterion=lambda i, *_: i < 3)
        _check_traj(f"user/{lin_name}", res.mean, res.chol, ores.mean, ores.chol, tol=1e-8)


def test_reference_golden_bearings():
    """The reference's only stored goldens (tests/test_bearings_only.py:23-72): 100-pass ICKS / IEKS on
    tests/bearings/ys.npy, produced upstream by a float32 run and compared there at 3 decimals.  The
    cubature iteration sits on a period-4 limit cycle, and the stored golden corresponds to 100 passes in
    total = the initial pass + 99 applications (see tests/test_oracle_golden.py::test_reference_golden_icks)."""
    import os
    import psqrt
    from psqrt.models import bearings
    gold = os.path.join(os.path.dirname(__file__), "golden")
    ys = np.load(os.path.join(gold, "bearings_ys.npy")).astype(np.float64)
    with np.load(os.path.join(gold, "bearings_icks.npz")) as z:
        exp_m, exp_P = z["arr_0"], z["arr_1"]
    Q, R, obs_f, trans_f = bearings.make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    x0 = psqrt.MVNSqrt(_g(np.array([-1.0, -1.0, 0.0, 0.0, 0.0])), _g(np.eye(5)))
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(_g(np.zeros(5)), _g(np.linalg.cholesky(Q))))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(_g(np.zeros(2)), _g(np.linalg.cholesky(R))))
    res = psqrt.iterated_smoothing(ys, x0, tm, om, psqrt.linearization.cubature, None, True,
                                   criterion=lambda i, *_: i < 99)
    m = res.mean.cpu().numpy()[1:]
    P = LLt(res.chol.cpu().numpy())[1:]
    np.testing.assert_array_almost_equal(m, exp_m, decimal=3)
    np.testing.assert_array_almost_equal(P, exp_P, decimal=3)
    with np.load(os.path.join(gold, "bearings_ieks.npz")) as z:
        exp_m, exp_P = z["arr_0"], z["arr_1"]
    res = psqrt.iterated_smoothing(ys, x0, tm, om, psqrt.linearization.extended, None, True,
                                   criterion=lambda i, *_: i < 100)
    np.testing.assert_array_almost_equal(res.mean.cpu().numpy()[1:], exp_m, decimal=3)
    np.testing.assert_array_almost_equal(LLt(res.chol.cpu().numpy())[1:], exp_P, decimal=3)


@pytest.mark.parametrize("n,ny,T,S", [(3, 2, 14, 6), (4, 2, 5000, 7), (1, 1, 300, 33), (5, 2, 2000, 130), (8, 4, 60, 5)])
def test_sampler_vs_oracle(n, ny, T, S):
    """psqrt.sampling (psqrt_smoother_elements + psqrt_sample_paths) against the oracle's restatement of
    parsmooth/_pathwise_sampler.py on the SAME normal draws, in the sign convention of the CUDA path
    (non-negative factor diagonals).  T = 5000 with 7 samples takes the time-chunked route, the others the
    one-thread-per-sample route."""
    import psqrt
    case = lgssm_case(n, ny, T, seed=11 * n + ny)
    tm, om = _torch_models(case)
    otm, oom = oracle_lgssm_models(case)
    x0 = psqrt.MVNSqrt(_g(case["m0"]), _g(case["L0"]))
    filt = psqrt.filtering(case["ys"], x0, tm, om, psqrt.linearization.extended)
    smo = psqrt.smoothing(tm, filt, psqrt.linearization.extended)
    eps = np.random.RandomState(5).randn(T + 1, S, n)
    got = psqrt.sampling(eps, S, tm, filt, psqrt.linearization.extended, smo)
    of = O.MVNSqrt(filt.mean.cpu().numpy(), filt.chol.cpu().numpy())
    os_ = O.MVNSqrt(smo.mean.cpu().numpy(), smo.chol.cpu().numpy())
    exp = O.seq_sampling(eps, otm, of, O.extended, os_, canonical=True)
    assert got.shape == (T + 1, S, n)
    assert rel_err(got.cpu().numpy(), exp) < TOL
    # nominal_trajectory=None linearises at the smoothed trajectory (methods.py:86-87): same result here
    got2 = psqrt.sampling(eps, S, tm, filt, psqrt.linearization.extended)
    assert rel_err(got2.cpu().numpy(), exp) < TOL


@pytest.mark.parametrize("dim_x,dim_y", [(1, 2), (3, 3)])
def test_sampler_marginals(dim_x, dim_y):
    """The reference's own sampler test (tests/test_sampler.py:28-66): the marginals of 100 000 joint samples
    reproduce the smoothing means and variances to 1e-2."""
    import psqrt
    T, N = 10, 100_000
    case = lgssm_case(dim_x, dim_y, T, seed=3)
    tm, om = _torch_models(case)
    x0 = psqrt.MVNSqrt(_g(case["m0"]), _g(case["L0"]))
    for lin in (psqrt.linearization.cubature, psqrt.linearization.extended):
        filt = psqrt.filtering(case["ys"], x0, tm, om, lin)
        smo = psqrt.smoothing(tm, filt, lin)
        samples = psqrt.sampling(123, N, tm, filt, lin, smo)
        assert samples.shape == (T + 1, N, dim_x)
        var = LLt(smo.chol.cpu().numpy()).diagonal(axis1=1, axis2=2)
        np.testing.assert_allclose(samples.mean(1).cpu().numpy(), smo.mean.cpu().numpy(), rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(samples.var(1).cpu().numpy(), var, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("n", [1, 2])
def test_get_conditional_model(n):
    """psqrt.linearization.get_conditional_model (linearization/_common.py:17-66) against the oracle on a function
    that is non-linear in the state AND the noise, inner == outer for the sigma-point methods, inner extended under
    an outer cubature, and the reference's linear check (tests/test_linearization.py:243-284) with extended in
    extended; dimension mismatch raises NotImplementedError."""
    import psqrt
    from psqrt.linearization import get_conditional_model
    from test_oracle_golden import _gcm_function
    rng = np.random.RandomState(40 + n)
    a, b, c = rng.randn(n, n), rng.randn(n, n), rng.randn(n)
    at, bt, ct = _g(a), _g(b), _g(c)
    f = lambda x, q: at @ x + torch.sin(x) * q + bt @ q + 0.3 * q * q + ct          # noqa: E731
    of = _gcm_function(a, b, c)
    qm, qL = rng.randn(n), 0.5 * (np.tril(rng.rand(n, n)) + np.eye(n))
    T = 17
    xm, xL = rng.randn(T, n), 0.4 * (np.tril(rng.rand(T, n, n)) + np.eye(n))
    x, ox = psqrt.MVNSqrt(_g(xm), _g(xL)), O.MVNSqrt(xm, xL)
    q, oq = psqrt.MVNSqrt(_g(qm), _g(qL)), O.MVNSqrt(qm, qL)
    L = psqrt.linearization
    for inner, outer in (("cubature", "cubature"), ("gauss_hermite", "gauss_hermite"), ("unscented", "unscented"),
                         ("extended", "cubature"), ("extended", "extended")):
        F, ch, rem = getattr(L, outer)(get_conditional_model(f, q, getattr(L, inner)), x)
        oF, och, orem = getattr(O, outer)(O.get_conditional_model(of, oq, getattr(O, inner)), ox)
        assert rel_err(F.cpu().numpy(), oF) < TOL and rel_err(rem.cpu().numpy(), orem) < TOL, (inner, outer)
        assert rel_err(LLt(ch.cpu().numpy()), LLt(och)) < TOL, (inner, outer)
    flin = lambda x, q: at @ x + bt @ q + ct                                        # noqa: E731
    for name in ("extended", "cubature", "gauss_hermite", "unscented"):
        F, ch, rem = getattr(L, name)(get_conditional_model(flin, q, getattr(L, name)), psqrt.MVNSqrt(_g(xm[0]), _g(xL[0])))
        np.testing.assert_allclose(F.cpu().numpy(), a, atol=1e-8)
        np.testing.assert_allclose(rem.cpu().numpy(), b @ qm + c, atol=1e-8)
        np.testing.assert_allclose(LLt(ch.cpu().numpy()), LLt(b @ qL), atol=1e-8)
    bad = _g(np.ones((n, n + 1)))
    with pytest.raises(NotImplementedError):
        get_conditional_model(lambda x, q: at @ x + bad @ q, psqrt.MVNSqrt(_g(np.zeros(n + 1)), _g(np.eye(n + 1))),
                              L.extended)


def test_full_size_properties():
    """BASELINE size (T = 1e6, nx = 4) through size-independent properties: (1) the pass is
    invariant to the chunking (two different chunk lengths = two different association orders);
    (2) its first 20 000 steps agree with the oracle filter; (3) the last smoothed state equals the
    last filtered state; (4) smoothed factors never exceed filtered ones in trace."""
    from psqrt import _lib
    T = 1_000_000
    case = lgssm_case(4, 2, 20_000, seed=0)
    rng = np.random.RandomState(1)
    ys = np.concatenate([case["ys"], rng.randn(T - 20_000, 2)])
    ssm = _ssm(case)
    a = _lib.filter_smoother(ssm, _g(ys), _g(case["m0"]), _g(case["L0"]), smooth=True, loglik=True)
    b = _lib.filter_smoother(ssm, _g(ys), _g(case["m0"]), _g(case["L0"]), smooth=True, loglik=True, chunk_len=61)
    for i, name in ((0, "fm"), (2, "sm")):
        assert rel_err(a[i].cpu().numpy(), b[i].cpu().numpy()) < TOL, name
    for i, name in ((1, "fL"), (3, "sL")):
        assert rel_err(LLt(a[i].cpu().numpy()), LLt(b[i].cpu().numpy())) < TOL, name
    assert abs(a[4].item() - b[4].item()) <= TOL_ELL * abs(b[4].item())
    ofm, ofc, _, _, _ = oracle_from_ssm(case)
    _check_traj("prefix", a[0][:20_001], a[1][:20_001], ofm, ofc)
    assert torch.equal(a[2][-1], a[0][-1])
    tr_f = (a[1] ** 2).sum((-1, -2))
    tr_s = (a[3] ** 2).sum((-1, -2))
    assert bool((tr_s <= tr_f * (1 + 1e-9)).all())


@pytest.mark.parametrize("kind", ["cholR=0", "R*1e-6", "R*1e12", "Q*1e12", "Q*1e-12"])
@pytest.mark.parametrize("n,ny,T,K", [(4, 2, 3000, 0), (4, 2, 200, 7), (3, 1, 64, 4), (5, 2, 1500, 9), (2, 2, 40, 5),
                                      (8, 4, 300, 3)])
def test_limit_cases(kind, n, ny, T, K):
    """The reference's limit cases (tests/test_sequential_filter.py:105-184: no information R*1e12, (almost)
    infinite information R*1e-6; tests/test_sequential_smoother.py:71-120: no / infinite process noise) plus an
    exactly noise-free observation (cholR = 0: the FILTERED covariance is singular, only predicted factors may be
    inverted, as parallel/_smoothing.py:76-85 does), through the whole CUDA pass against the oracle's parallel AND
    sequential square-root smoothers."""
    from psqrt import _lib
    if kind == "cholR=0" and ny >= n:
        pytest.skip("noise-free observation of the whole state")
    case = lgssm_case(n, ny, T, seed=3 * n + ny)
    if kind == "cholR=0":
        case["cholR"] = np.zeros_like(case["cholR"])
    elif kind == "R*1e-6":
        case["cholR"] = 1e-3 * case["cholR"]
    elif kind == "R*1e12":
        case["cholR"] = 1e6 * case["cholR"]
    elif kind == "Q*1e12":
        case["cholQ"] = 1e6 * case["cholQ"]
    elif kind == "Q*1e-12":
        case["cholQ"] = 1e-6 * case["cholQ"]
    fm, fL, sm, sL, ell = _lib.filter_smoother(_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]),
                                               smooth=True, loglik=True, chunk_len=K)
    assert bool(torch.isfinite(sm).all()) and bool(torch.isfinite(sL).all())
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    tol = TOL if kind != "Q*1e-12" else 1e-7      # nearly deterministic dynamics: cond(P_pred) ~ 1e12
    _check_traj("filtered", fm, fL, ofm, ofc, tol=tol)
    _check_traj("smoothed", sm, sL, osm, osc, tol=tol)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)
    if T <= 300:
        tm, om = oracle_lgssm_models(case)
        x0 = O.MVNSqrt(case["m0"], case["L0"])
        sf = O.seq_filtering(case["ys"], x0, tm, om, O.extended, None)
        ss = O.seq_smoothing(tm, sf, O.extended, None)
        _check_traj("smoothed-vs-seq", sm, sL, ss.mean, ss.chol, tol=max(tol, 1e-8))


@pytest.mark.parametrize("lin_name", ["extended", "cubature"])
def test_bearings_full_size_pass(lin_name):
    """BASELINE.json configs[1]/[2] at their own size (bearings-only coordinated turn, nx = 5, T = 1e5): one sqrt
    parallel filter + smoother pass + log-likelihood through the public API against the oracle at the same nominal
    trajectory (a smooth random walk around the notebooks' initial nominal, so every step linearises elsewhere)."""
    import psqrt
    T = 100_000
    ys, m0, cholQ, cholR, (obs_f, trans_f), (oobs, otrans) = _bearings_setup(T)
    lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
    rng = np.random.RandomState(5)
    nom_m = np.array([-1.0, -1.0, 6.0, 4.0, 2.0]) + 0.02 * np.cumsum(rng.randn(T + 1, 5), 0) / np.sqrt(np.arange(1, T + 2))[:, None]
    nom_L = 0.2 * (np.tril(rng.rand(T + 1, 5, 5)) + np.eye(5))
    x0 = psqrt.MVNSqrt(_g(m0), _g(np.eye(5)))
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(_g(np.zeros(5)), _g(cholQ)))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(_g(np.zeros(2)), _g(cholR)))
    nominal = psqrt.MVNSqrt(_g(nom_m), _g(nom_L))
    filt, ell = psqrt.filtering(ys, x0, tm, om, lin, nominal, True, True)
    smo = psqrt.filter_smoother(ys, x0, tm, om, lin, nominal, True)
    otm = O.FunctionalModel(otrans, O.MVNSqrt(np.zeros(5), cholQ))
    oom = O.FunctionalModel(oobs, O.MVNSqrt(np.zeros(2), cholR))
    ox0, onom = O.MVNSqrt(m0, np.eye(5)), O.MVNSqrt(nom_m, nom_L)
    ofilt, oell = O.filtering(ys, ox0, otm, oom, olin, onom, True, True)
    osmo = O.smoothing(otm, ofilt, olin, onom, True)
    _check_traj("filtered", filt.mean, filt.chol, ofilt.mean, ofilt.chol)
    _check_traj("smoothed", smo.mean, smo.chol, osmo.mean, osmo.chol)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)


def test_fused_builtin_linearization(monkeypatch):
    """The bearings-only model with the extended linearization is linearised INSIDE the sweeps (psqrt_ssm.fused_model,
    csrc/psqrt_fused.cuh; parallel/_filtering.py:110-119 evaluated per step in registers): same pass as the unfused path
    (psqrt_linearize_builtin + model arrays, PSQRT_FUSE_LIN=0) and as the oracle, including a |w| < 1e-6 nominal turn
    rate; and no linearization kernel is launched on the fused path."""
    import psqrt
    from psqrt import _lib
    T = 3000
    ys, m0, cholQ, cholR, (obs_f, trans_f), (oobs, otrans) = _bearings_setup(T)
    rng = np.random.RandomState(17)
    nom_m = np.array([-1.0, -1.0, 6.0, 4.0, 2.0]) + 0.05 * np.cumsum(rng.randn(T + 1, 5), 0) / np.sqrt(np.arange(1, T + 2))[:, None]
    nom_m[11, 4] = 1e-9
    nom_L = np.repeat(np.eye(5)[None], T + 1, 0)
    x0 = psqrt.MVNSqrt(m0, np.eye(5))
    mq, mr = 0.01 * rng.randn(5), 0.01 * rng.randn(2)
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(mq, cholQ))        # host inputs: mirrors for the fused path
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(mr, cholR))
    nominal = psqrt.MVNSqrt(_g(nom_m), _g(nom_L))
    calls = []
    real = _lib.linearize_builtin
    monkeypatch.setattr(_lib, "linearize_builtin", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    lin = psqrt.linearization.extended
    monkeypatch.setenv("PSQRT_FUSE_LIN", "1")
    filt, ell = psqrt.filtering(ys, x0, tm, om, lin, nominal, True, True)
    smo = psqrt.filter_smoother(ys, x0, tm, om, lin, nominal, True)
    assert not calls, "the fused path must not launch the linearization kernels"
    monkeypatch.setenv("PSQRT_FUSE_LIN", "0")
    filt0, ell0 = psqrt.filtering(ys, x0, tm, om, lin, nominal, True, True)
    smo0 = psqrt.filter_smoother(ys, x0, tm, om, lin, nominal, True)
    assert calls, "PSQRT_FUSE_LIN=0 must take the unfused path"
    monkeypatch.setenv("PSQRT_FUSE_LIN", "1")
    _check_traj("fused vs unfused, filtered", filt.mean, filt.chol, filt0.mean.cpu().numpy(), filt0.chol.cpu().numpy())
    _check_traj("fused vs unfused, smoothed", smo.mean, smo.chol, smo0.mean.cpu().numpy(), smo0.chol.cpu().numpy())
    assert abs(ell.item() - ell0.item()) <= TOL_ELL * abs(ell0.item())
    otm = O.FunctionalModel(otrans, O.MVNSqrt(mq, cholQ))
    oom = O.FunctionalModel(oobs, O.MVNSqrt(mr, cholR))
    onom = O.MVNSqrt(nom_m, nom_L)
    ofilt, oell = O.filtering(ys, O.MVNSqrt(m0, np.eye(5)), otm, oom, O.extended, onom, True, True)
    osmo = O.filter_smoother(ys, O.MVNSqrt(m0, np.eye(5)), otm, oom, O.extended, onom, True)
    _check_traj("fused vs oracle, filtered", filt.mean, filt.chol, ofilt.mean, ofilt.chol)
    _check_traj("fused vs oracle, smoothed", smo.mean, smo.chol, osmo.mean, osmo.chol)
    assert abs(ell.item() - oell) <= TOL_ELL * abs(oell)
    # iterated smoother through the fused path
    res = psqrt.iterated_smoothing(ys, x0, tm, om, lin, nominal, True, criterion=lambda i, *_: i < 5)
    ores = O.iterated_smoothing(ys, O.MVNSqrt(m0, np.eye(5)), otm, oom, O.extended, onom, True,
                                criterion=lambda i, *_: i < 5)
    _check_traj("fused iterated vs oracle", res.mean, res.chol, ores.mean, ores.chol, tol=1e-7)


def test_count_nonfinite():
    """psqrt_count_nonfinite: the NaN-rate report of the robustness sweeps (notebooks/robustness_100runs.py:41-77)."""
    from psqrt import _lib
    rng = np.random.RandomState(2)
    x = rng.randn(7, 1001, 5)
    x[2, 17, 3] = np.nan
    x[2, 900, 0] = np.inf
    x[5, 0, 0] = -np.inf
    counts = _lib.count_nonfinite(_g(x)).cpu().numpy()
    assert counts.tolist() == [0, 0, 2, 0, 0, 1, 0]
    big = torch.zeros(3, 300001, dtype=torch.float64, device=_dev())
    big[1, 299999] = float("nan")
    assert _lib.count_nonfinite(big).cpu().numpy().tolist() == [0, 1, 0]
