"""Compare the oracle (oracle/parsmooth_np.py) with tests/golden/reference_vectors.npz -- outputs of the
reference's own source files run on a NumPy shim of JAX (tests/golden/make_golden.py).  Needs neither
/root/reference nor the shim, so it runs on the GPU box too.  Factors are compared through L L^T."""
from __future__ import annotations

import numpy as np

import parsmooth_np as O

TOL = 1e-10


def _close(a, b, what, tol=TOL):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(float(np.max(np.abs(b))), 1e-300) if b.size else 1.0
    err = float(np.max(np.abs(a - b))) / scale if b.size else 0.0
    assert a.shape == b.shape and err < tol, f"{what}: relative error {err:.3e} (shapes {a.shape} vs {b.shape})"


def _LLt(L):
    return L @ np.swapaxes(L, -1, -2)


def check_all(path):
    v = dict(np.load(path))
    n_checked = 0
    # 1. operators
    for n in (1, 2, 3, 5):
        e1 = tuple(v[f"fop{n}_e1_{i}"][None] for i in range(5))
        e2 = tuple(v[f"fop{n}_e2_{i}"][None] for i in range(5))
        res = O.sqrt_filtering_operator(e1, e2)
        for i, fac in enumerate((0, 0, 1, 0, 1)):
            got, exp = res[i][0], v[f"fop{n}_out_{i}"]
            _close(_LLt(got) if fac else got, _LLt(exp) if fac else exp, f"filtering operator n={n} field {i}")
        s1 = tuple(v[f"sop{n}_e1_{i}"][None] for i in range(3))
        s2 = tuple(v[f"sop{n}_e2_{i}"][None] for i in range(3))
        sres = O.sqrt_smoothing_operator(s1, s2)
        _close(sres[0][0], v[f"sop{n}_out_0"], "smoothing operator g")
        _close(sres[1][0], v[f"sop{n}_out_1"], "smoothing operator E")
        _close(_LLt(sres[2][0]), _LLt(v[f"sop{n}_out_2"]), "smoothing operator D")
        n_checked += 8
    # 2. math utils
    for n in (2, 3, 5):
        L, vec, V = v[f"chol{n}_L"], v[f"chol{n}_v"], v[f"chol{n}_V"]
        _close(O.cholesky_update(L, vec, 1.0), v[f"chol{n}_up"], "cholesky update")
        _close(O.cholesky_update(L, vec, -0.1), v[f"chol{n}_down"], "cholesky downdate")
        _close(O.cholesky_update_many(L, V, -1.0), v[f"chol{n}_many"], "cholesky_update_many", 1e-9)
        bad = O.cholesky_update_many(0.1 * L, 10 * V, -1.0)
        assert np.all(np.isfinite(bad)) and np.all(np.isfinite(v[f"chol{n}_bad"]))
        _close(bad, v[f"chol{n}_bad"], "non-finite guard", 1e-9)
        _close(_LLt(O.tria(v[f"tria{n}_in"])), _LLt(v[f"tria{n}_out"]), "tria")
        _close(O.mvn_loglikelihood(v[f"mvn{n}_x"], L), v[f"mvn{n}_ll"], "mvn_loglikelihood")
        n_checked += 6
    for n, p in ((1, 3), (2, 3), (5, 3), (2, 5)):
        wm, _, xi = O.gauss_hermite_weights(n, p)
        _close(wm, v[f"gh{n}_{p}_wm"], "GH weights", 1e-15)
        _close(xi, v[f"gh{n}_{p}_xi"], "GH points", 1e-15)
        n_checked += 2
    wm, _, xi = O.cubature_weights(5)
    _close(wm, v["cub5_wm"], "cubature weights", 1e-15)
    _close(xi, v["cub5_xi"], "cubature points", 1e-15)
    # 3. LGSSM end to end
    lins = {"ext": O.extended, "cub": O.cubature, "gh": O.gauss_hermite}
    for (n, ny, T) in ((2, 1, 12), (3, 2, 15), (1, 3, 9), (4, 2, 20)):
        tag = f"lg{n}{ny}"
        g = lambda k: v[f"{tag}_{k}"]
        tm = O.FunctionalModel(O.lgssm_function(g("F")), O.MVNSqrt(g("b"), g("cQ")))
        om = O.FunctionalModel(O.lgssm_function(g("H")), O.MVNSqrt(g("c"), g("cR")))
        x0 = O.MVNSqrt(g("m0"), g("L0"))
        nom = O.MVNSqrt(g("nom_m"), np.repeat(np.eye(n)[None], T + 1, 0))
        for lname, lin in lins.items():
            if lname == "gh" and n > 3:
                continue
            for par in (True, False):
                p = "par" if par else "seq"
                f, ell = O.filtering(g("ys"), x0, tm, om, lin, nom, par, True)
                s = O.smoothing(tm, f, lin, nom, par)
                what = f"{tag}/{lname}/{p}"
                _close(f.mean, g(f"{lname}_{p}_fm"), what + " filtered mean")
                _close(_LLt(f.chol), _LLt(g(f"{lname}_{p}_fc")), what + " filtered cov")
                _close(s.mean, g(f"{lname}_{p}_sm"), what + " smoothed mean")
                _close(_LLt(s.chol), _LLt(g(f"{lname}_{p}_sc")), what + " smoothed cov")
                _close(ell, g(f"{lname}_{p}_ell"), what + " ell")
                n_checked += 5
        lin = O.linearize_ssm(O.extended, tm, om, O.MVNSqrt(nom.mean[:2], nom.chol[:2]))
        el = O.sqrt_filtering_elements(*lin, g("m0")[None], g("L0")[None], g("ys")[:1])
        for i, fac in enumerate((0, 0, 1, 0, 1)):
            got, exp = el[i][0], g(f"elem_{i}")
            _close(_LLt(got) if fac else got, _LLt(exp) if fac else exp, f"{tag} filtering element field {i}")
        _close(O.sqrt_loglikelihood_terms(*lin, g("m0")[None], g("L0")[None], g("ys")[:1])[0], g("ellterm"),
               f"{tag} ell term")
        sel = O.sqrt_smoothing_elements(lin[0], lin[1], lin[2], g("m0")[None], g("L0")[None])
        _close(sel[0][0], g("selem_0"), f"{tag} smoothing element g")
        _close(sel[1][0], g("selem_1"), f"{tag} smoothing element E")
        _close(_LLt(sel[2][0]), _LLt(g("selem_2")), f"{tag} smoothing element D")
        n_checked += 9
    # 4. bearings-only model
    Q, R, obs, trans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    cQ, cR = np.linalg.cholesky(Q), np.linalg.cholesky(R)
    tm = O.FunctionalModel(trans, O.MVNSqrt(np.zeros(5), cQ))
    om = O.FunctionalModel(obs, O.MVNSqrt(np.zeros(2), cR))
    pts = O.MVNSqrt(v["bear_pts_m"], v["bear_pts_L"])
    for lname, lin in lins.items():
        for mname, model in (("t", tm), ("o", om)):
            F, ch, b = lin(model, pts)
            _close(F, v[f"bear_{lname}_{mname}_F"], f"bearings {lname} {mname} F", 1e-9)
            _close(_LLt(ch), _LLt(v[f"bear_{lname}_{mname}_chol"]), f"bearings {lname} {mname} chol", 1e-9)
            _close(b, v[f"bear_{lname}_{mname}_b"], f"bearings {lname} {mname} b", 1e-9)
            n_checked += 3
    x0 = O.MVNSqrt(np.array([-1.0, -1.0, 0.0, 0.0, 0.0]), np.eye(5))
    for lname, lin, iters in (("ext", O.extended, 4), ("cub", O.cubature, 3)):
        res, ell = O.iterated_smoothing(v["bear_ys"], x0, tm, om, lin, None, True, criterion=lambda i, *_: i < iters,
                                        return_loglikelihood=True)
        _close(res.mean, v[f"bear_it_{lname}_m"], f"bearings iterated {lname} mean", 1e-8)
        _close(_LLt(res.chol), _LLt(v[f"bear_it_{lname}_c"]), f"bearings iterated {lname} cov", 1e-8)
        _close(ell, v[f"bear_it_{lname}_ell"], f"bearings iterated {lname} ell", 1e-8)
        res2 = O.filter_smoother(v["bear_ys"], x0, tm, om, lin, None, False)
        _close(res2.mean, v[f"bear_seq_{lname}_m"], f"bearings sequential {lname} mean", 1e-8)
        _close(_LLt(res2.chol), _LLt(v[f"bear_seq_{lname}_c"]), f"bearings sequential {lname} cov", 1e-8)
        n_checked += 5
    # 5. population model (conditional moments; upstream helpers use Q = 0.3^2, lam = 10)
    tmod, omod = O.population_model(10.0, np.array([[0.3 ** 2]]))
    pts = O.MVNSqrt(v["pop_pts_m"], v["pop_pts_L"])
    for lname, lin in lins.items():
        for mname, model in (("t", tmod), ("o", omod)):
            F, ch, b = lin(model, pts)
            _close(F, v[f"pop_{lname}_{mname}_F"], f"population {lname} {mname} F", 1e-9)
            _close(_LLt(ch), _LLt(v[f"pop_{lname}_{mname}_chol"]), f"population {lname} {mname} chol", 1e-9)
            _close(b, v[f"pop_{lname}_{mname}_b"], f"population {lname} {mname} b", 1e-9)
            n_checked += 3
    return n_checked


if __name__ == "__main__":
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(here)), "oracle"))
    print("checked", check_all(os.path.join(here, "reference_vectors.npz")), "groups")
