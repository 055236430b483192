"""GPU tests of the gradient path (BASELINE.json configs[2]: "log-likelihood + gradient path"): the forward-mode tangent
kernels of csrc/psqrt_tangent.cu through the C ABI, against
  * the oracle's direct differentiation of the sequential recursions (seq_filter_smoother_jvp; itself pinned by central
    differences of the reference-pinned parallel pass, tests/test_tangent_maps.py)                       -- 1e-9,
  * central differences of the oracle's linearizations (tangents of the built-in linearization kernels)  -- 1e-6,
  * central differences of the UNMODIFIED reference source (tests/golden/reference_vectors_grad.npz) for
    d ell / d prec_r of the iterated smoothers, the protocol of
    notebooks/experiment_bearing_only_param_estimation_run_time.ipynb                                    -- 1e-6.
"""
import os

import numpy as np
import pytest
import torch

import parsmooth_np as O
from _cases import bearings_pe_case, lgssm_case, oracle_bearings_pe_models, rel_err, time_varying_case

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda", 0)


def _g(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=_dev())


def _sym(dc, c):
    t = dc @ np.swapaxes(c, -1, -2)
    return t + np.swapaxes(t, -1, -2)


@pytest.mark.parametrize("n,ny,T", [(4, 2, 300), (5, 2, 1000), (1, 1, 40), (2, 1, 7), (3, 3, 257), (6, 4, 90),
                                    (8, 4, 70), (4, 2, 1), (5, 2, 9000), (4, 2, 70000)])
def test_tangent_pass_vs_oracle(n, ny, T):
    """psqrt_filter_smoother_tangent on a time-varying model with every tangent non-zero."""
    from psqrt import _lib
    case = time_varying_case(n, ny, T, seed=3 * n + ny)
    rng = np.random.RandomState(n + T)
    d = dict(dF=0.3 * rng.randn(T, n, n), dcQ=np.tril(rng.randn(T, n, n)) * 0.1, db=rng.randn(T, n),
             dH=rng.randn(T, ny, n), dcR=0.1 * rng.randn(T, ny, ny), dc=rng.randn(T, ny))
    dm0, dL0 = rng.randn(n), np.tril(rng.randn(n, n))
    dQ, dR, dP0 = _sym(d["dcQ"], case["cholQ"]), _sym(d["dcR"], case["cholR"]), _sym(dL0, case["L0"])
    ssm = _lib.LinearizedSSM(*[_g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    ys = _g(case["ys"])
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, ys, _g(case["m0"]), _g(case["L0"]), smooth=True, loglik=True)
    dssm = {"dF": _g(d["dF"]), "dQ": _g(dQ), "db": _g(d["db"]), "dH": _g(d["dH"]), "dR": _g(dR), "dc": _g(d["dc"])}
    dfm, dfP, dsm, dsP, dell = _lib.filter_smoother_tangent(ssm, dssm, ys, fm, fL, sm, sL, _g(dm0), _g(dP0))
    if T > 20000:   # the NumPy recursion is slow: compare a prefix of the filter tangent, and ell on the short cases
        Tc = 3000
        ref = O.seq_filter_smoother_jvp(tuple(case[k][:Tc] for k in ("F", "cholQ", "b", "H", "cholR", "c")),
                                        tuple(a[:Tc] for a in (d["dF"], dQ, d["db"], d["dH"], dR, d["dc"])),
                                        case["m0"], case["L0"], dm0, dP0, case["ys"][:Tc])
        assert rel_err(dfm[:Tc + 1].cpu().numpy(), ref["dfm"]) < 1e-9
        assert rel_err(dfP[:Tc + 1].cpu().numpy(), ref["dfP"]) < 1e-9
        assert torch.isfinite(dsm).all() and torch.isfinite(dsP).all() and torch.isfinite(dell)
        assert rel_err(dsm[-1].cpu().numpy(), dfm[-1].cpu().numpy()) == 0.0
        return
    ref = O.seq_filter_smoother_jvp(tuple(case[k] for k in ("F", "cholQ", "b", "H", "cholR", "c")),
                                    (d["dF"], dQ, d["db"], d["dH"], dR, d["dc"]), case["m0"], case["L0"], dm0, dP0,
                                    case["ys"])
    assert abs(ell.item() - ref["ell"]) < 1e-8 * abs(ref["ell"])
    for name, got in (("dfm", dfm), ("dfP", dfP), ("dsm", dsm), ("dsP", dsP)):
        err = rel_err(got.cpu().numpy(), ref[name])
        assert err < 1e-9, f"{name}: {err:.3e}"
    assert abs(dell.item() - ref["dell"]) < 1e-9 * max(1.0, abs(ref["dell"]))
    # factor tangent: d(L L^T) = dP for the lower-triangular factors the pass wrote
    dsL = _lib.cov_tangent_to_chol(sL[1:], dsP[1:])
    back = dsL @ sL[1:].transpose(-1, -2)
    back = back + back.transpose(-1, -2)
    assert rel_err(back.cpu().numpy(), dsP[1:].cpu().numpy()) < 1e-9
    assert float(torch.triu(dsL, 1).abs().max()) == 0.0
    oracle_dL = O.chol_tangent_from_cov(sL[1:].cpu().numpy(), dsP[1:].cpu().numpy())
    assert rel_err(dsL.cpu().numpy(), oracle_dL) < 1e-9


def test_tangent_pass_time_invariant_and_zero_entries():
    """Time-invariant model and tangents (stride 0), NULL tangents: the parameter-estimation shape (only dR)."""
    from psqrt import _lib
    n, ny, T = 4, 2, 500
    case = lgssm_case(n, ny, T, seed=42)
    dcR = np.array([[-0.3, 0.0], [0.0, 0.0]])
    dR = _sym(dcR, case["cholR"])
    ssm = _lib.LinearizedSSM(*[_g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    ys = _g(case["ys"])
    fm, fL, sm, sL, _ = _lib.filter_smoother(ssm, ys, _g(case["m0"]), _g(case["L0"]), smooth=True, loglik=True)
    out = _lib.filter_smoother_tangent(ssm, {"dR": _g(dR)}, ys, fm, fL, sm, sL, None, None)
    rep = lambda a: np.repeat(a[None], T, 0)
    z = lambda *s: np.zeros((T,) + s)
    ref = O.seq_filter_smoother_jvp(tuple(rep(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")),
                                    (z(n, n), z(n, n), z(n), z(ny, n), rep(dR), z(ny)), case["m0"], case["L0"],
                                    np.zeros(n), np.zeros((n, n)), case["ys"])
    for name, got in zip(("dfm", "dfP", "dsm", "dsP"), out[:4]):
        assert rel_err(got.cpu().numpy(), ref[name]) < 1e-9, name
    assert abs(out[4].item() - ref["dell"]) < 1e-9 * max(1.0, abs(ref["dell"]))


@pytest.mark.parametrize("lin_name", ["extended", "cubature", "gauss_hermite", "unscented"])
def test_builtin_linearization_tangents(lin_name):
    """psqrt_linearize_builtin_tangent against central differences of the oracle's linearization along a random
    direction of (nominal mean, nominal factor, noise mean, noise factor); coordinated turn (incl. the |w| < 1e-6
    branch), bearings, and the conditional-moments population model."""
    import psqrt
    from psqrt import grad
    from psqrt.models import bearings, population
    rng = np.random.RandomState(21)
    T = 64
    nm = rng.randn(T, 5) * np.array([2.0, 2.0, 3.0, 3.0, 1.0])
    nm[5, 4] = 1e-8
    nL = 0.3 * (np.tril(rng.rand(T, 5, 5)) + np.eye(5))
    dnm, dnL = rng.randn(T, 5), np.tril(rng.randn(T, 5, 5))
    dnm[5, 4] = 0.0      # stay inside the branch
    s1, s2 = np.array([-1.5, 0.5]), np.array([1.0, 1.0])
    Q, R, obs_f, trans_f = bearings.make_parameters(0.01, 0.1, 0.5, 0.01, s1, s2)
    _, _, oobs, otrans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, s1, s2)
    cQ, cR = np.linalg.cholesky(Q), np.linalg.cholesky(R)
    lin, olin = getattr(psqrt.linearization, lin_name), getattr(O, lin_name)
    h = 1e-6
    x = psqrt.MVNSqrt(_g(nm), _g(nL))
    for f, of, c_q in ((trans_f, otrans, cQ), (obs_f, oobs, cR)):
        d = c_q.shape[0]
        m_q, dm_q, dc_q = 0.1 * rng.randn(d), rng.randn(d), 0.1 * np.tril(rng.randn(d, d))
        model = psqrt.FunctionalModel(f, psqrt.MVNSqrt(_g(m_q), _g(c_q)))
        dF, dQ, db = grad._linearization_tangent(lin, model, x, _g(dnm), _g(dnL),
                                                 psqrt.MVNSqrt(_g(dm_q), _g(dc_q)), None, _dev())

        def at(eps):
            F, ch, b = olin(O.FunctionalModel(of, O.MVNSqrt(m_q + eps * dm_q, c_q + eps * dc_q)),
                            O.MVNSqrt(nm + eps * dnm, nL + eps * dnL))
            ch = np.broadcast_to(ch, (T, d, d))
            return F, ch @ np.swapaxes(ch, -1, -2), b

        fd = [(p - m) / (2 * h) for p, m in zip(at(h), at(-h))]
        dQ = np.broadcast_to(dQ.cpu().numpy(), fd[1].shape)
        for name, got, ref in (("dF", dF.cpu().numpy(), fd[0]), ("dQ", dQ, fd[1]), ("db", db.cpu().numpy(), fd[2])):
            err = np.max(np.abs(got - ref)) / max(1.0, np.max(np.abs(ref)))
            assert err < 2e-6, f"{lin_name} {of.__name__ if hasattr(of, '__name__') else ''} {name}: {err:.3e}"
    # conditional-moments population model (no noise tangent)
    pm, pL = np.log(7.0) + 0.3 * rng.randn(T, 1), 0.2 + 0.1 * rng.rand(T, 1, 1)
    dpm, dpL = rng.randn(T, 1), rng.randn(T, 1, 1)
    tmod, omod = population.make_parameters(10.0, np.array([[0.09]]))
    otmod, oomod = O.population_model(10.0, np.array([[0.09]]))
    for mod, omodel in ((tmod, otmod), (omod, oomod)):
        dF, dQ, db = grad._linearization_tangent(lin, mod, psqrt.MVNSqrt(_g(pm), _g(pL)), _g(dpm), _g(dpL), None, None,
                                                 _dev())

        def at(eps):
            F, ch, b = olin(omodel, O.MVNSqrt(pm + eps * dpm, pL + eps * dpL))
            return F, ch @ np.swapaxes(ch, -1, -2), b

        fd = [(p - m) / (2 * h) for p, m in zip(at(h), at(-h))]
        for name, got, ref in (("dF", dF, fd[0]), ("dQ", dQ, fd[1]), ("db", db, fd[2])):
            err = np.max(np.abs(got.cpu().numpy() - ref)) / max(1.0, np.max(np.abs(ref)))
            assert err < 2e-6, f"population {lin_name} {name}: {err:.3e}"


def _pe_models(case, prec):
    import psqrt
    from psqrt.models import bearings
    Q, R, obs_f, trans_f = bearings.make_parameters(case["qc"], case["qw"], 1.0 / prec, case["dt"], case["s1"],
                                                    case["s2"], r2=0.1)
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(_g(np.zeros(5)), _g(np.linalg.cholesky(Q))))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(_g(np.zeros(2)), _g(np.linalg.cholesky(R))))
    return tm, om


@pytest.mark.parametrize("T,seed,lname", [(60, 0, "ext"), (60, 0, "cub"), (60, 0, "gh"), (120, 1, "ext"),
                                          (120, 1, "cub")])
def test_loglikelihood_gradient_vs_reference_source(T, seed, lname):
    """d ell / d prec_r of iterated_smoothing(..., return_loglikelihood=True) (psqrt.grad.loglikelihood_jvp: implicit
    fixed-point tangent, n_iter + 2 Neumann terms like the reference's custom VJP) against central differences of the
    unmodified reference source on the same data."""
    import psqrt
    from psqrt import grad
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_grad.npz"))
    key = f"pe_T{T}_{lname}"
    prec, iters = float(g[key + "_prec"]), int(g[key + "_iters"])
    lin = {"ext": psqrt.linearization.extended, "cub": psqrt.linearization.cubature,
           "gh": psqrt.linearization.gauss_hermite}[lname]
    case = bearings_pe_case(T, seed)
    x0 = psqrt.MVNSqrt(_g(case["m0"]), _g(case["L0"]))
    tm, om = _pe_models(case, prec)
    # chol R = diag(1 / prec, 0.1): d / d prec = diag(-1 / prec^2, 0)
    tg = grad.Tangents(observation_noise=psqrt.MVNSqrt(None, _g(np.diag([-1.0 / prec ** 2, 0.0]))))
    nominal, ell, dell = grad.loglikelihood_jvp(_g(case["ys"]), x0, tm, om, lin, tg, None, True,
                                                criterion=lambda i, *_: i < iters)
    assert abs(ell.item() - float(g[key + "_ell"])) < 1e-8 * abs(float(g[key + "_ell"]))
    assert rel_err(nominal.mean.cpu().numpy(), g[key + "_m"]) < 1e-7
    ref = float(g[key + "_dell"])
    assert abs(dell.item() - ref) < 1e-6 * max(1.0, abs(ref)), (dell.item(), ref)
    # the same through the jax.value_and_grad-like front end (tangents of the inputs from torch.func.jvp)
    from psqrt.models import bearings
    Q, _, obs_f, trans_f = bearings.make_parameters(case["qc"], case["qw"], 1.0, case["dt"], case["s1"], case["s2"],
                                                    r2=0.1)
    cQ = _g(np.linalg.cholesky(Q))

    def build(theta):
        cR = torch.diag(torch.stack([1.0 / theta[0], torch.as_tensor(0.1, dtype=torch.float64, device=theta.device)]))
        zero5, zero2 = torch.zeros(5, dtype=torch.float64, device=_dev()), torch.zeros(2, dtype=torch.float64, device=_dev())
        return (psqrt.MVNSqrt(_g(case["m0"]), _g(case["L0"])), psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(zero5, cQ)),
                psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(zero2, cR)))

    ell2, grad2 = grad.value_and_grad(build, [prec], _g(case["ys"]), lin, None, criterion=lambda i, *_: i < iters)
    assert abs(ell2.item() - ell.item()) < 1e-10 * abs(ell.item())
    assert abs(grad2[0].item() - dell.item()) < 1e-9 * max(1.0, abs(dell.item()))


def test_loglikelihood_gradient_lgssm_all_parameters():
    """Linear model: d ell along random directions of every input (x0, both noises) against central differences of the
    oracle's log-likelihood; one tangent pass (no fixed point to differentiate: the linearization is exact)."""
    import psqrt
    from psqrt import grad
    from psqrt.models import lgssm
    n, ny, T = 4, 2, 400
    case = lgssm_case(n, ny, T, seed=9)
    rng = np.random.RandomState(3)
    d = dict(m0=rng.randn(n), L0=np.tril(rng.randn(n, n)), b=rng.randn(n), cQ=0.1 * np.tril(rng.randn(n, n)),
             c=rng.randn(ny), cR=0.1 * np.tril(rng.randn(ny, ny)))

    def oracle_ell(eps):
        tm = O.FunctionalModel(O.lgssm_function(case["F"]), O.MVNSqrt(case["b"] + eps * d["b"], case["cholQ"] + eps * d["cQ"]))
        om = O.FunctionalModel(O.lgssm_function(case["H"]), O.MVNSqrt(case["c"] + eps * d["c"], case["cholR"] + eps * d["cR"]))
        x0 = O.MVNSqrt(case["m0"] + eps * d["m0"], case["L0"] + eps * d["L0"])
        return O.filtering(case["ys"], x0, tm, om, O.extended, None, True, True)[1]

    h = 1e-5
    ref = (oracle_ell(h) - oracle_ell(-h)) / (2 * h)
    x0 = psqrt.MVNSqrt(_g(case["m0"]), _g(case["L0"]))
    tm = psqrt.FunctionalModel(lgssm.transition_function(case["F"]), psqrt.MVNSqrt(_g(case["b"]), _g(case["cholQ"])))
    om = psqrt.FunctionalModel(lgssm.observation_function(case["H"]), psqrt.MVNSqrt(_g(case["c"]), _g(case["cholR"])))
    tg = grad.Tangents(x0=psqrt.MVNSqrt(_g(d["m0"]), _g(d["L0"])),
                       transition_noise=psqrt.MVNSqrt(_g(d["b"]), _g(d["cQ"])),
                       observation_noise=psqrt.MVNSqrt(_g(d["c"]), _g(d["cR"])))
    nominal = psqrt.MVNSqrt(torch.zeros(T + 1, n, dtype=torch.float64, device=_dev()),
                            torch.eye(n, dtype=torch.float64, device=_dev()).expand(T + 1, n, n).contiguous())
    tp = grad.TangentPass(_g(case["ys"]), x0, tm, om, psqrt.linearization.extended, nominal, tg)
    assert abs(tp.ell.item() - oracle_ell(0.0)) < 1e-8 * abs(tp.ell.item())
    dell = tp.ell_jvp(None)
    assert abs(dell.item() - ref) < 1e-6 * max(1.0, abs(ref)), (dell.item(), ref)


@pytest.mark.parametrize("n,ny,T", [(4, 2, 300), (5, 2, 1000), (1, 1, 40), (2, 1, 7), (3, 3, 257), (6, 4, 90),
                                    (8, 4, 70), (4, 2, 1), (5, 2, 9000), (4, 2, 70000)])
def test_adjoint_pass(n, ny, T):
    """psqrt_loglik_adjoint (reverse mode) on a time-varying model:
      * the dot-product test against the forward-mode tangent pass for a random direction of every model entry and of
        the prior (<gradient, direction> == d ell) -- 1e-9,
      * costates and per-step gradients against the NumPy mirror of the adjoint algebra (tests/_tangent_maps.py, itself
        checked against the oracle's direct differentiation on the CPU) on short cases -- 1e-9."""
    from psqrt import _lib
    import _tangent_maps as TM
    case = time_varying_case(n, ny, T, seed=5 * n + ny)
    rng = np.random.RandomState(n + T + 1)
    d = dict(dF=0.3 * rng.randn(T, n, n), dcQ=np.tril(rng.randn(T, n, n)) * 0.1, db=rng.randn(T, n),
             dH=rng.randn(T, ny, n), dcR=0.1 * rng.randn(T, ny, ny), dc=rng.randn(T, ny))
    dm0, dL0 = rng.randn(n), np.tril(rng.randn(n, n))
    dQ, dR, dP0 = _sym(d["dcQ"], case["cholQ"]), _sym(d["dcR"], case["cholR"]), _sym(dL0, case["L0"])
    ssm = _lib.LinearizedSSM(*[_g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    ys = _g(case["ys"])
    fm, fL, _, _, _ = _lib.filter_smoother(ssm, ys, _g(case["m0"]), _g(case["L0"]), smooth=False, loglik=True)
    g = _lib.loglik_adjoint(ssm, ys, fm, fL)
    dssm = {"dF": _g(d["dF"]), "dQ": _g(dQ), "db": _g(d["db"]), "dH": _g(d["dH"]), "dR": _g(dR), "dc": _g(d["dc"])}
    _, _, _, _, dell = _lib.filter_smoother_tangent(ssm, dssm, ys, fm, fL, None, None, _g(dm0), _g(dP0), smooth=False)
    dot = (g["lam"][0] @ _g(dm0) + (g["Lam"][0] * _g(dP0)).sum() + (g["gF"] * dssm["dF"]).sum() +
           (g["gQ"] * dssm["dQ"]).sum() + (g["gb"] * dssm["db"]).sum() + (g["gH"] * dssm["dH"]).sum() +
           (g["gR"] * dssm["dR"]).sum() + (g["gc"] * dssm["dc"]).sum())
    scale = sum(float((g[k] * dssm["d" + k[1:]]).abs().sum()) for k in ("gF", "gQ", "gb", "gH", "gR", "gc"))
    assert abs(dot.item() - dell.item()) < 1e-9 * max(1.0, scale), (dot.item(), dell.item())
    assert float(g["lam"][T].abs().max()) == 0.0 and float(g["Lam"][T].abs().max()) == 0.0
    assert float((g["Lam"] - g["Lam"].transpose(-1, -2)).abs().max()) < 1e-12 * max(1.0, float(g["Lam"].abs().max()))
    if T > 2000:
        return
    fmn, fLn = fm.cpu().numpy(), fL.cpu().numpy()
    lam = np.zeros((T + 1, n))
    Lam = np.zeros((T + 1, n, n))
    z = np.zeros
    ref = {k: [] for k in ("gF", "gQ", "gb", "gH", "gR", "gc")}
    for t in range(T - 1, -1, -1):
        args = [case[k][t] for k in ("F", "cholQ", "b", "H", "cholR", "c")] + [case["ys"][t], fmn[t], fLn[t]]
        a, r = TM.felem(*args, z((n, n)), z((n, n)), z(n), z((ny, n)), z((ny, ny)), z(ny))
        for k, v in zip(("gF", "gQ", "gb", "gH", "gR", "gc"), TM.adj_grad(*args, lam[t + 1], Lam[t + 1])):
            ref[k].append(v)
        lam[t], Lam[t] = TM.apply_adj(TM.adj_map(a, r), lam[t + 1], Lam[t + 1])
    assert rel_err(g["lam"].cpu().numpy(), lam) < 1e-9
    assert rel_err(g["Lam"].cpu().numpy(), Lam) < 1e-9
    for k in ref:
        assert rel_err(g[k].cpu().numpy(), np.stack(ref[k][::-1])) < 1e-9, k


def test_reverse_mode_lgssm_gradient_vs_oracle_differences():
    """psqrt.grad.value_and_grad_reverse: all parameters of an LGSSM from ONE adjoint pass, against central differences
    of the oracle's parallel filter log-likelihood (reference-pinned) and against the forward-mode path."""
    import psqrt
    from psqrt import grad as G
    from psqrt.models import lgssm
    n, ny, T = 4, 2, 400
    case = lgssm_case(n, ny, T, seed=9)
    ys = _g(case["ys"])
    base = {k: _g(case[k]) for k in ("m0", "L0", "b", "cholQ", "c", "cholR")}
    theta0 = np.array([0.3, -0.2, 0.15, 0.4, -0.1, 0.25, 0.05])

    def build(th):
        x0 = psqrt.MVNSqrt(base["m0"] + th[0], base["L0"] * torch.exp(th[1]))
        q = psqrt.MVNSqrt(base["b"] * (1.0 + th[2]), base["cholQ"] * torch.exp(th[3]) +
                          th[6] * torch.tril(torch.ones_like(base["cholQ"]), -1))
        r = psqrt.MVNSqrt(base["c"] + th[4], base["cholR"] * torch.exp(th[5]))
        return (x0, psqrt.FunctionalModel(lgssm.transition_function(case["F"]), q),
                psqrt.FunctionalModel(lgssm.observation_function(case["H"]), r))

    ell, grad = G.value_and_grad_reverse(build, theta0, ys, psqrt.linearization.extended)

    def oracle_ell(th):
        th = torch.as_tensor(th, dtype=torch.float64, device=_dev())
        x0, tm, om = build(th)
        c = dict(case)
        c.update(m0=x0.mean.cpu().numpy(), L0=x0.chol.cpu().numpy(), b=tm.mvn.mean.cpu().numpy(),
                 cholQ=tm.mvn.chol.cpu().numpy(), c=om.mvn.mean.cpu().numpy(), cholR=om.mvn.chol.cpu().numpy())
        rep = lambda a: np.repeat(a[None], T, 0)
        ssm = tuple(rep(c[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c"))
        ms = np.concatenate([c["m0"][None], np.zeros((T - 1, n))])
        Ls = np.concatenate([c["L0"][None], np.zeros((T - 1, n, n))])
        _, fm, fc, _, _ = O.associative_scan(O.sqrt_filtering_operator, O.sqrt_filtering_elements(*ssm, ms, Ls, c["ys"]))
        fm = np.concatenate([c["m0"][None], fm])
        fc = np.concatenate([c["L0"][None], fc])
        return float(np.sum(O.sqrt_loglikelihood_terms(*ssm, fm[:-1], fc[:-1], c["ys"])))

    assert abs(ell.item() - oracle_ell(theta0)) < 1e-8 * abs(ell.item())
    h = 1e-5
    for i in range(theta0.size):
        e = np.zeros_like(theta0)
        e[i] = h
        fd = (oracle_ell(theta0 + e) - oracle_ell(theta0 - e)) / (2 * h)
        assert abs(grad[i].item() - fd) < 2e-6 * max(1.0, abs(fd)), (i, grad[i].item(), fd)
    # forward mode, one direction at a time, at the same (nominal-independent) model
    ell_f, grad_f = G.value_and_grad(build, theta0, ys, psqrt.linearization.extended,
                                     criterion=lambda i, *_: i < 1, implicit_terms=1)
    assert abs(ell_f.item() - ell.item()) < 1e-10 * abs(ell.item())
    assert rel_err(grad.cpu().numpy(), grad_f.cpu().numpy()) < 1e-8
