"""CPU checks of the gradient path's algebra (no GPU): the affine tangent maps of csrc/psqrt_tangent.cu (mirrored in
tests/_tangent_maps.py) against the oracle's direct differentiation of the sequential filter / smoother, their
associativity, the oracle tangent against central finite differences of the reference-pinned parallel pass, and the
factor tangent."""
import numpy as np

import parsmooth_np as O
import _tangent_maps as TM
from _cases import lgssm_case


def _problem(n, ny, T, seed):
    case = lgssm_case(n, ny, T, seed=seed)
    rng = np.random.default_rng(seed + 100)
    rep = lambda a: np.repeat(a[None], T, 0)
    F, cQ, b, H, cR, c = [rep(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")]
    F = F + 0.02 * rng.standard_normal(F.shape)
    H = H + 0.02 * rng.standard_normal(H.shape)
    d = dict(dF=rng.standard_normal(F.shape), dcQ=np.tril(rng.standard_normal(cQ.shape)),
             db=rng.standard_normal(b.shape), dH=rng.standard_normal(H.shape),
             dcR=rng.standard_normal(cR.shape), dc=rng.standard_normal(c.shape),
             dm0=rng.standard_normal(n), dL0=np.tril(rng.standard_normal((n, n))))
    sym = lambda dc_, c_: dc_ @ np.swapaxes(c_, -1, -2) + c_ @ np.swapaxes(dc_, -1, -2)
    d["dQ"], d["dR"] = sym(d["dcQ"], cQ), sym(d["dcR"], cR)
    d["dP0"] = sym(d["dL0"], case["L0"])
    return case, (F, cQ, b, H, cR, c), d


def _pass(ssm, m0, L0, ys):
    """reference-pinned parallel pass from given model arrays (oracle elements + associative scans)"""
    T, n = ys.shape[0], m0.shape[0]
    ms = np.concatenate([m0[None], np.zeros((T - 1, n))])
    Ls = np.concatenate([L0[None], np.zeros((T - 1, n, n))])
    _, fm, fc, _, _ = O.associative_scan(O.sqrt_filtering_operator, O.sqrt_filtering_elements(*ssm, ms, Ls, ys))
    fm = np.concatenate([m0[None], fm])
    fc = np.concatenate([L0[None], fc])
    ell = np.sum(O.sqrt_loglikelihood_terms(*ssm, fm[:-1], fc[:-1], ys))
    g, E, D = O.sqrt_smoothing_elements(ssm[0], ssm[1], ssm[2], fm[:-1], fc[:-1])
    g = np.concatenate([g, fm[-1:]])
    E = np.concatenate([E, np.zeros((1, n, n))])
    D = np.concatenate([D, fc[-1:]])
    sm, _, sc = O.associative_scan(O.sqrt_smoothing_operator, (g, E, D), reverse=True)
    return fm, fc, sm, sc, ell


def test_oracle_tangent_vs_finite_differences():
    for n, ny, T, seed in ((4, 2, 50, 1), (5, 2, 30, 2), (3, 3, 40, 3), (2, 1, 25, 4)):
        case, ssm, d = _problem(n, ny, T, seed)
        F, cQ, b, H, cR, c = ssm

        def run(eps):
            s = (F + eps * d["dF"], cQ + eps * d["dcQ"], b + eps * d["db"], H + eps * d["dH"], cR + eps * d["dcR"],
                 c + eps * d["dc"])
            fm, fc, sm, sc, ell = _pass(s, case["m0"] + eps * d["dm0"], case["L0"] + eps * d["dL0"], case["ys"])
            return fm, fc @ np.swapaxes(fc, -1, -2), sm, sc @ np.swapaxes(sc, -1, -2), ell

        h = 1e-5
        fd = [(x - y) / (2 * h) for x, y in zip(run(h), run(-h))]
        out = O.seq_filter_smoother_jvp(ssm, (d["dF"], d["dQ"], d["db"], d["dH"], d["dR"], d["dc"]), case["m0"],
                                        case["L0"], d["dm0"], d["dP0"], case["ys"])
        for name, ref in zip(("dfm", "dfP", "dsm", "dsP", "dell"), fd):
            err = np.max(np.abs(out[name] - ref)) / np.max(np.abs(ref))
            assert err < 2e-7, (n, ny, name, err)     # central differences, h = 1e-5
        assert abs(out["ell"] - run(0.0)[4]) < 1e-9 * abs(out["ell"])


def test_affine_maps_reproduce_the_direct_tangent():
    for n, ny, T, seed in ((4, 2, 40, 5), (5, 2, 33, 6), (1, 1, 20, 7), (3, 3, 17, 8)):
        case, ssm, d = _problem(n, ny, T, seed)
        F, cQ, b, H, cR, c = ssm
        ref = O.seq_filter_smoother_jvp(ssm, (d["dF"], d["dQ"], d["db"], d["dH"], d["dR"], d["dc"]), case["m0"],
                                        case["L0"], d["dm0"], d["dP0"], case["ys"])
        fm, fc, sm, sc, _ = _pass(ssm, case["m0"], case["L0"], case["ys"])
        maps, recs = [], []
        for t in range(T):
            a, r = TM.felem(F[t], cQ[t], b[t], H[t], cR[t], c[t], case["ys"][t], fm[t], fc[t], d["dF"][t], d["dQ"][t],
                            d["db"][t], d["dH"][t], d["dR"][t], d["dc"][t])
            maps.append(a)
            recs.append(r)
        # step by step, and through composed prefixes (what the scan does)
        dm, dP = d["dm0"], d["dP0"]
        dell = 0.0
        acc = None
        for t in range(T):
            B, e = recs[t]
            dell += maps[t][1] @ dm + np.sum(B * dP) + e
            dm, dP = TM.apply(maps[t], dm, dP)
            acc = maps[t] if acc is None else TM.compose(acc, maps[t])
            dm2, dP2 = TM.apply(acc, d["dm0"], d["dP0"])
            scale = max(1.0, np.max(np.abs(ref["dfP"][t + 1])))
            assert np.max(np.abs(dm - ref["dfm"][t + 1])) < 1e-9 * max(1.0, np.max(np.abs(ref["dfm"])))
            assert np.max(np.abs(dP - ref["dfP"][t + 1])) < 1e-9 * scale
            assert np.max(np.abs(dm2 - dm)) < 1e-9 * max(1.0, np.max(np.abs(dm)))
            assert np.max(np.abs(dP2 - dP)) < 1e-9 * scale
        assert abs(dell - ref["dell"]) < 1e-9 * max(1.0, abs(ref["dell"]))
        # smoother, backwards
        dm, dP = ref["dfm"][T], ref["dfP"][T]
        for t in range(T - 1, -1, -1):
            a = TM.selem(F[t], cQ[t], b[t], fm[t], fc[t], ref["dfm"][t], ref["dfP"][t], sm[t + 1], sc[t + 1],
                         d["dF"][t], d["dQ"][t], d["db"][t])
            dm, dP = TM.apply(a, dm, dP)
            assert np.max(np.abs(dm - ref["dsm"][t])) < 1e-9 * max(1.0, np.max(np.abs(ref["dsm"])))
            assert np.max(np.abs(dP - ref["dsP"][t])) < 1e-9 * max(1.0, np.max(np.abs(ref["dsP"])))


def test_affine_map_composition_is_associative():
    rng = np.random.default_rng(11)
    n = 4
    def rnd():
        C = rng.standard_normal((n, n))
        return rng.standard_normal((n, n)), rng.standard_normal(n), rng.standard_normal(n), C + C.T
    a, b, c = rnd(), rnd(), rnd()
    left = TM.compose(TM.compose(a, b), c)
    right = TM.compose(a, TM.compose(b, c))
    for x, y in zip(left, right):
        assert np.allclose(x, y, rtol=1e-12, atol=1e-12)


def test_chol_tangent_from_cov():
    rng = np.random.default_rng(12)
    for n in (1, 3, 5):
        L = np.tril(rng.standard_normal((7, n, n))) + 2 * np.eye(n)
        L[:, 0, 0] *= -1          # any diagonal signs
        dL = np.tril(rng.standard_normal((7, n, n)))
        dP = dL @ np.swapaxes(L, -1, -2) + L @ np.swapaxes(dL, -1, -2)
        assert np.allclose(O.chol_tangent_from_cov(L, dP), dL, rtol=1e-10, atol=1e-10)


def _grad_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors_grad.npz"))


def test_oracle_implicit_gradient_vs_reference_source_differences():
    """The forward-mode statement of the reference's implicit fixed-point gradient (oracle.implicit_loglikelihood_jvp)
    against central differences of the UNMODIFIED reference source (tests/golden/make_golden_grad.py) for the
    parameter-estimation protocol (prec_r -> ell), extended / cubature / Gauss-Hermite."""
    from _cases import bearings_pe_case, oracle_bearings_pe_models
    g = _grad_golden()
    for T, seed, lname, lin in ((60, 0, "ext", O.extended), (60, 0, "cub", O.cubature), (60, 0, "gh", O.gauss_hermite),
                                (120, 1, "ext", O.extended)):
        key = f"pe_T{T}_{lname}"
        prec, iters = float(g[key + "_prec"]), int(g[key + "_iters"])
        case = bearings_pe_case(T, seed)
        x0 = O.MVNSqrt(case["m0"], case["L0"])
        tm, om = oracle_bearings_pe_models(case, prec)
        nom, ell = O.iterated_smoothing(case["ys"], x0, tm, om, lin, None, True, lambda i, *_: i < iters, True)
        assert abs(ell - float(g[key + "_ell"])) < 1e-9 * abs(ell)
        assert np.max(np.abs(nom.mean - g[key + "_m"])) < 1e-8
        _, dell = O.implicit_loglikelihood_jvp(case["ys"], lambda e: x0,
                                               lambda e: oracle_bearings_pe_models(case, prec + e)[0],
                                               lambda e: oracle_bearings_pe_models(case, prec + e)[1], lin, nom,
                                               iters + 2)
        ref = float(g[key + "_dell"])
        assert abs(dell - ref) < 1e-6 * max(1.0, abs(ref)), (key, dell, ref)


def test_adjoint_maps_and_gradient_contraction():
    """Reverse mode (psqrt_loglik_adjoint, mirrored in _tangent_maps.adj_*): the costate recursion and the per-step
    gradient contraction reproduce the forward-mode d ell for random directions of EVERY model entry and of the prior
    (the dot-product test), step by step and through composed suffix maps (what the reverse scan does)."""
    for n, ny, T, seed in ((4, 2, 40, 21), (5, 2, 33, 22), (1, 1, 20, 23), (3, 3, 17, 24), (2, 4, 12, 25)):
        case, ssm, d = _problem(n, ny, T, seed)
        F, cQ, b, H, cR, c = ssm
        ref = O.seq_filter_smoother_jvp(ssm, (d["dF"], d["dQ"], d["db"], d["dH"], d["dR"], d["dc"]), case["m0"],
                                        case["L0"], d["dm0"], d["dP0"], case["ys"])
        fm, fc, _, _, _ = _pass(ssm, case["m0"], case["L0"], case["ys"])
        z = lambda a: np.zeros_like(a)
        amaps = []
        for t in range(T):
            a, r = TM.felem(F[t], cQ[t], b[t], H[t], cR[t], c[t], case["ys"][t], fm[t], fc[t], z(F[t]), z(F[t]),
                            z(b[t]), z(H[t]), z(cR[t]), z(c[t]))
            amaps.append(TM.adj_map(a, r))
        lam = np.zeros((T + 1, n))
        Lam = np.zeros((T + 1, n, n))
        acc = None
        for t in range(T - 1, -1, -1):
            lam[t], Lam[t] = TM.apply_adj(amaps[t], lam[t + 1], Lam[t + 1])
            acc = amaps[t] if acc is None else TM.compose_adj(acc, amaps[t])
            l2, L2 = TM.apply_adj(acc, lam[T], Lam[T])
            assert np.allclose(l2, lam[t], rtol=1e-10, atol=1e-10)
            assert np.allclose(L2, Lam[t], rtol=1e-10, atol=1e-10)
        dell = lam[0] @ d["dm0"] + np.sum(Lam[0] * d["dP0"])
        for t in range(T):
            gF, gQ, gb, gH, gR, gc = TM.adj_grad(F[t], cQ[t], b[t], H[t], cR[t], c[t], case["ys"][t], fm[t], fc[t],
                                                 lam[t + 1], Lam[t + 1])
            dell += (np.sum(gF * d["dF"][t]) + np.sum(gQ * d["dQ"][t]) + gb @ d["db"][t] + np.sum(gH * d["dH"][t]) +
                     np.sum(gR * d["dR"][t]) + gc @ d["dc"][t])
        assert abs(dell - ref["dell"]) < 1e-9 * max(1.0, abs(ref["dell"])), (n, ny, dell, ref["dell"])
