"""Golden vectors for the SURVEY section 8(f) components (unscented linearization, pathwise sampler),
produced like make_golden.py: the UNMODIFIED reference sources from /root/reference executed on the
NumPy shim of the JAX surface they use.  `jax.random.normal` is shimmed by a seeded NumPy generator and
the draws are stored next to the outputs, so the sampler comparison is deterministic.

    python tests/golden/make_golden_next.py      (build container only: needs /root/reference)

Writes tests/golden/reference_vectors_next.npz.
"""
from __future__ import annotations

import os
import sys
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (the shim lives there)

OUT = os.path.join(HERE, "reference_vectors_next.npz")


def main():
    if not os.path.isdir(mg.REFERENCE):
        raise SystemExit("needs the reference checkout at /root/reference")
    mg.install_shim()
    draws = []

    def normal(key, shape, dtype=None):
        rng = np.random.RandomState(int(np.asarray(key).ravel()[-1]))
        e = rng.randn(*shape)
        draws.append(e)
        return mg._wrap(e)

    sys.modules["jax.random"].normal = normal
    sys.modules["jax.random"].PRNGKey = lambda s: np.array([0, s], dtype=np.uint32)
    sys.path.insert(0, mg.REFERENCE)
    from parsmooth._base import MVNSqrt, FunctionalModel, ConditionalMomentsModel
    from parsmooth.linearization import unscented, cubature, extended
    from parsmooth.linearization._unscented import _unscented_weights
    from parsmooth.methods import filtering, smoothing, sampling
    from tests._lgssm import transition_function as lgssm_f, observation_function as lgssm_h
    from tests.bearings.bearings_utils import make_parameters as bearings_parameters
    import tests.test_linearization as tl

    A = lambda v: np.array(np.asarray(v), dtype=np.float64)
    rng = np.random.RandomState(2024)
    tril = lambda n: np.tril(rng.rand(n, n)) + np.eye(n)
    out = {}

    # 1. unscented weights and linearisations --------------------------------------------------------
    for n in (1, 2, 5):
        wm, wc, lam = _unscented_weights(n, 1.0, 0.0, 3.0 + n)
        out[f"ut{n}_wm"], out[f"ut{n}_wc"], out[f"ut{n}_lam"] = A(wm), A(wc), A(lam)
    s1, s2 = np.array([-1.5, 0.5]), np.array([1.0, 1.0])
    Q, R, obs_f, trans_f = bearings_parameters(0.01, 0.1, 0.5, 0.01, s1, s2)
    cQ, cR = np.linalg.cholesky(A(Q)), np.linalg.cholesky(A(R))
    tm = FunctionalModel(trans_f, MVNSqrt(np.zeros(5), cQ))
    om = FunctionalModel(obs_f, MVNSqrt(np.zeros(2), cR))
    pts_m = np.array([[-1.0, -1.0, 6.0, 4.0, 2.0], [0.3, -2.0, 1.0, 0.5, 1e-8], [2.0, 1.5, -3.0, 0.2, -0.7]])
    pts_L = np.stack([np.eye(5), 0.3 * tril(5), 0.1 * tril(5)])
    out["bear_pts_m"], out["bear_pts_L"] = pts_m, pts_L
    for mname, model in (("t", tm), ("o", om)):
        res = [unscented(model, MVNSqrt(pts_m[i], pts_L[i])) for i in range(3)]
        for j, nm in enumerate(("F", "chol", "b")):
            out[f"bear_ut_{mname}_{nm}"] = A(np.stack([A(r[j]) for r in res]))
    lam = 10.0
    tmod = ConditionalMomentsModel(tl.transition_mean, tl.transition_chol)
    omod = ConditionalMomentsModel(partial(tl.observation_mean, lam=lam), partial(tl.observation_chol, lam=lam))
    pm, pL = np.array([[np.log(7.0)], [0.5], [2.2]]), np.array([[[1.0]], [[0.3]], [[0.05]]])
    out["pop_pts_m"], out["pop_pts_L"] = pm, pL
    for mname, model in (("t", tmod), ("o", omod)):
        res = [unscented(model, MVNSqrt(pm[i], pL[i])) for i in range(3)]
        for j, nm in enumerate(("F", "chol", "b")):
            out[f"pop_ut_{mname}_{nm}"] = A(np.stack([A(r[j]) for r in res]))
    # an LGSSM pass with the unscented linearisation (exact for linear models)
    n, ny, T = 3, 2, 14
    Fm = 0.9 * np.linalg.qr(rng.randn(n, n))[0]
    Hm = rng.randn(ny, n)
    cq, cr = 0.3 * tril(n), 0.4 * tril(ny)
    b, c, m0, L0 = 0.1 * rng.randn(n), 0.1 * rng.randn(ny), rng.randn(n), tril(n)
    ys = rng.randn(T, ny)
    nom = MVNSqrt(rng.randn(T + 1, n), np.repeat(np.eye(n)[None], T + 1, 0))
    for k, v in dict(F=Fm, H=Hm, cQ=cq, cR=cr, b=b, c=c, m0=m0, L0=L0, ys=ys, nom_m=nom.mean).items():
        out[f"lg_{k}"] = A(v)
    ltm = FunctionalModel(partial(lgssm_f, A=Fm), MVNSqrt(b, cq))
    lom = FunctionalModel(partial(lgssm_h, H=Hm), MVNSqrt(c, cr))
    x0 = MVNSqrt(m0, L0)
    f, ell = filtering(ys, x0, ltm, lom, unscented, nom, True, return_loglikelihood=True)
    s = smoothing(ltm, f, unscented, nom, True)
    out["lg_ut_fm"], out["lg_ut_fc"], out["lg_ut_sm"], out["lg_ut_sc"], out["lg_ut_ell"] = (
        A(f.mean), A(f.chol), A(s.mean), A(s.chol), A(ell))

    # 2. pathwise sampler (parsmooth/_pathwise_sampler.py), parallel and sequential --------------------
    f, _ = filtering(ys, x0, ltm, lom, extended, nom, True, return_loglikelihood=True)
    s = smoothing(ltm, f, extended, nom, True)
    out["smp_fm"], out["smp_fc"], out["smp_sm"], out["smp_sc"] = A(f.mean), A(f.chol), A(s.mean), A(s.chol)
    key = np.array([0, 123], dtype=np.uint32)
    for lname, lin in (("ext", extended), ("cub", cubature)):
        for par in (True, False):
            del draws[:]
            smp = sampling(key, 6, ltm, f, lin, s, parallel=par)
            out[f"smp_{lname}_{'par' if par else 'seq'}"] = A(smp)
            out["smp_eps"] = A(draws[0])                      # same key -> same draws in every call
    # nonlinear transition (bearings CT model), nominal = smoothed trajectory of an extended pass
    Tb = 25
    ysb = np.load(os.path.join(mg.REFERENCE, "tests", "bearings", "ys.npy")).astype(np.float64)[:Tb]
    xb0 = MVNSqrt(np.array([-1.0, -1.0, 0.0, 0.0, 0.0]), np.eye(5))
    fb = filtering(ysb, xb0, tm, om, extended, None, True)
    sb = smoothing(tm, fb, extended, None, True)
    out["smpb_ys"], out["smpb_fm"], out["smpb_fc"], out["smpb_sm"], out["smpb_sc"] = (
        ysb, A(fb.mean), A(fb.chol), A(sb.mean), A(sb.chol))
    del draws[:]
    smp = sampling(key, 4, tm, fb, extended, sb, parallel=True)
    out["smpb_ext_par"], out["smpb_eps"] = A(smp), A(draws[0])

    # 3. get_conditional_model (linearization/_common.py:17-66) on a function non-linear in x AND q ----------
    from parsmooth.linearization import gauss_hermite, get_conditional_model
    jnp = sys.modules["jax.numpy"]
    for n in (1, 2):
        a, bq, cc = rng.randn(n, n), rng.randn(n, n), rng.randn(n)
        f = lambda x, q_, a=a, bq=bq, cc=cc: a @ x + jnp.sin(x) * q_ + bq @ q_ + 0.3 * q_ * q_ + cc   # noqa: E731
        qmvn = MVNSqrt(rng.randn(n), 0.5 * tril(n))
        xs_m, xs_L = rng.randn(3, n), np.stack([0.4 * tril(n) for _ in range(3)])
        for k, v in dict(a=a, b=bq, c=cc, qm=qmvn.mean, qL=qmvn.chol, xm=xs_m, xL=xs_L).items():
            out[f"gcm{n}_{k}"] = A(v)
        # outer == inner for the sigma-point methods (as tests/test_linearization.py:243-284 does); inner extended
        # under an outer cubature (the shim's complex-step jacfwd cannot be nested, so no extended-in-extended)
        for tag, inner, outer in (("cub", cubature, cubature), ("gh", gauss_hermite, gauss_hermite),
                                  ("ut", unscented, unscented), ("extcub", extended, cubature)):
            model = get_conditional_model(f, qmvn, inner)
            res = [outer(model, MVNSqrt(xs_m[i], xs_L[i])) for i in range(3)]
            for j, nm in enumerate(("F", "chol", "b")):
                out[f"gcm{n}_{tag}_{nm}"] = A(np.stack([A(r[j]) for r in res]))

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
