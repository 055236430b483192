"""Generate tests/golden/reference_vectors.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

JAX is not installed in this image (and there is no network), so `import parsmooth` fails as is.
This script installs a small NumPy/SciPy shim of exactly the JAX surface parsmooth touches
(jax.numpy, jax.scipy.linalg.{qr, solve_triangular, solve, cho_solve}, jax.vmap, jax.jacfwd,
jax.lax.{scan, cond, while_loop, associative_scan}, jax.tree_util.tree_map, jit, custom_vjp,
closure_convert) into sys.modules, imports the UNMODIFIED parsmooth package from /root/reference,
runs its functions on seeded inputs and stores inputs + outputs.  The formulas that produce the
vectors are therefore upstream's; only the array primitives underneath are NumPy/LAPACK
(jacfwd is a complex-step derivative, exact to rounding for the analytic test models).

    python tests/golden/make_golden.py      # needs /root/reference; run in the build container only

tests/golden/check_vectors.py compares the oracle (oracle/parsmooth_np.py) with the stored vectors and
runs everywhere (it needs neither /root/reference nor this shim).
"""
from __future__ import annotations

import os
import sys
import types
from functools import partial

import numpy as np
import scipy.linalg as sla

REFERENCE = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


# ------------------------------------------------------------------------------------------------
# the shim
# ------------------------------------------------------------------------------------------------
class JArr(np.ndarray):
    """ndarray with the .at[idx].set(value) functional update of jax arrays."""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        a = self.a

        class _Setter:
            def set(self, v):
                out = np.array(a, copy=True)
                out[idx] = v
                return out.view(JArr)

        return _Setter()


def _wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, JArr):
        return x.view(JArr)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(_wrap(v) for v in x)
    return x


def _wrapfn(f):
    def g(*a, **k):
        return _wrap(f(*a, **k))
    return g


def _arctan2(y, x):
    y, x = np.asarray(y), np.asarray(x)
    if np.iscomplexobj(y) or np.iscomplexobj(x):   # first-order analytic continuation (complex step)
        yr, xr = np.real(y), np.real(x)
        return np.arctan2(yr, xr) + 1j * (xr * np.imag(y) - yr * np.imag(x)) / (xr * xr + yr * yr)
    return np.arctan2(y, x)


class _NS(types.ModuleType):
    """module whose missing attributes fall through to a numpy namespace, outputs viewed as JArr"""

    def __init__(self, name, base):
        super().__init__(name)
        self._base = base

    def __getattr__(self, item):
        v = getattr(self._base, item)
        return _wrapfn(v) if callable(v) and not isinstance(v, type) else v


def _is_node(x):
    return x is None or isinstance(x, (tuple, list, dict))


def tree_leaves(t):
    if t is None:
        return []
    if isinstance(t, dict):
        return [l for k in t for l in tree_leaves(t[k])]
    if isinstance(t, (tuple, list)):
        return [l for v in t for l in tree_leaves(v)]
    return [t]


def tree_map(f, t, *rest):
    if t is None:
        return None
    if isinstance(t, dict):
        return {k: tree_map(f, t[k], *[r[k] for r in rest]) for k in t}
    if isinstance(t, tuple) and hasattr(t, "_fields"):
        return type(t)(*[tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t)])
    if isinstance(t, (tuple, list)):
        return type(t)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t))
    return f(t, *rest)


def _stack(trees):
    return tree_map(lambda *ls: _wrap(np.stack([np.asarray(l) for l in ls])), trees[0], *trees[1:])


def vmap(fn, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = list(in_axes) if isinstance(in_axes, (list, tuple)) else [in_axes] * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            ls = tree_leaves(a)
            if ls:
                n = np.asarray(ls[0]).shape[0]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else tree_map(lambda z: _wrap(np.asarray(z)[i]), a) for a, ax in zip(args, axes)]
            outs.append(fn(*call))
        return _stack(outs)
    return wrapped


def jacfwd(f, argnums=0):
    assert argnums == 0

    def jac(x, *a):
        x = np.asarray(x, dtype=float)
        cols = []
        for i in range(x.shape[0]):
            xc = x.astype(complex)
            xc[i] += 1e-30j
            cols.append(np.imag(np.asarray(f(xc.view(JArr), *a))) / 1e-30)
        return _wrap(np.stack(cols, -1))
    return jac


def scan(f, init, xs, length=None, reverse=False):
    ls = tree_leaves(xs)
    n = np.asarray(ls[0]).shape[0] if ls else length
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        carry, y = f(carry, tree_map(lambda z: _wrap(np.asarray(z)[i]), xs))
        ys[i] = y
    if ys and ys[0] is None:
        return carry, None
    return carry, _stack(ys)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def while_loop(cond_fun, body_fun, init):
    val = init
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def associative_scan(fn, elems, reverse=False, axis=0):
    """odd/even recursive doubling of jax.lax.associative_scan"""
    if reverse:
        elems = tree_map(lambda e: np.asarray(e)[::-1], elems)

    def sl(t, s):
        return tree_map(lambda e: _wrap(np.asarray(e)[s]), t)

    def _scan(es):
        n = np.asarray(tree_leaves(es)[0]).shape[0]
        if n < 2:
            return es
        reduced = fn(sl(es, slice(0, n - 1, 2)), sl(es, slice(1, None, 2)))
        odd = _scan(reduced)
        if n == 2:      # nothing beyond element 0 on the even side (an empty vmap batch in JAX)
            even = sl(es, slice(0, 1))
        else:
            if n % 2 == 0:
                even = fn(sl(odd, slice(None, -1)), sl(es, slice(2, None, 2)))
            else:
                even = fn(odd, sl(es, slice(2, None, 2)))
            even = tree_map(lambda e, ev: np.concatenate([np.asarray(e)[0:1], np.asarray(ev)]), es, even)

        def inter(ev, od):
            o = np.empty((n,) + np.asarray(ev).shape[1:], dtype=np.result_type(ev, od))
            o[0::2] = ev
            o[1::2] = od
            return _wrap(o)
        return tree_map(inter, even, odd)

    res = _scan(elems)
    if reverse:
        res = tree_map(lambda e: _wrap(np.asarray(e)[::-1]), res)
    return res


def _jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


class _CustomVjp:
    def __init__(self, f, nondiff_argnums=()):
        self.f = f

    def __call__(self, *a, **k):
        return self.f(*a, **k)

    def defvjp(self, fwd, bwd):
        pass


def install_shim():
    jnp = _NS("jax.numpy", np)
    jnp.ndarray = np.ndarray
    jnp.arctan2 = _arctan2
    jnp.linalg = _NS("jax.numpy.linalg", np.linalg)
    jsl = types.ModuleType("jax.scipy.linalg")
    jsl.qr = lambda a, mode="full": _wrap(tuple(sla.qr(np.asarray(a), mode=mode)))
    jsl.solve_triangular = lambda a, b, trans=0, lower=False: _wrap(
        sla.solve_triangular(np.asarray(a), np.asarray(b), trans=int(trans), lower=lower))
    jsl.solve = lambda a, b, assume_a="gen": _wrap(sla.solve(np.asarray(a), np.asarray(b), assume_a=assume_a))
    jsl.cho_solve = lambda c_and_lower, b: _wrap(sla.cho_solve((np.asarray(c_and_lower[0]), c_and_lower[1]),
                                                                 np.asarray(b)))
    jscipy = types.ModuleType("jax.scipy")
    jscipy.linalg = jsl
    lax = types.ModuleType("jax.lax")
    lax.scan, lax.cond, lax.while_loop, lax.associative_scan = scan, cond, while_loop, associative_scan
    tu = types.ModuleType("jax.tree_util")
    tu.tree_map = tree_map
    cd = types.ModuleType("jax.custom_derivatives")
    cd.closure_convert = lambda f, *ex: ((lambda x, *params: f(x)), ())
    fu = types.ModuleType("jax.flatten_util")
    fu.ravel_pytree = lambda t: (_ for _ in ()).throw(NotImplementedError("no reverse mode in the shim"))
    rnd = types.ModuleType("jax.random")
    tst = types.ModuleType("jax.test_util")
    tst.check_grads = lambda *a, **k: None
    jax = types.ModuleType("jax")
    jax.numpy, jax.scipy, jax.lax, jax.tree_util, jax.custom_derivatives = jnp, jscipy, lax, tu, cd
    jax.flatten_util, jax.random, jax.test_util = fu, rnd, tst
    jax.vmap, jax.jacfwd, jax.jit = vmap, jacfwd, _jit
    jax.custom_vjp = lambda f=None, nondiff_argnums=(): (_CustomVjp(f) if f is not None else
                                                        (lambda g: _CustomVjp(g)))
    jax.vjp = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError)
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)
    for name, mod in (("jax", jax), ("jax.numpy", jnp), ("jax.scipy", jscipy), ("jax.scipy.linalg", jsl),
                      ("jax.lax", lax), ("jax.tree_util", tu), ("jax.custom_derivatives", cd),
                      ("jax.flatten_util", fu), ("jax.random", rnd), ("jax.test_util", tst)):
        sys.modules[name] = mod


# ------------------------------------------------------------------------------------------------
# the cases
# ------------------------------------------------------------------------------------------------
def main():
    if not os.path.isdir(REFERENCE):
        raise SystemExit(f"{REFERENCE} not present: the vectors can only be regenerated in the build container")
    install_shim()
    sys.path.insert(0, REFERENCE)
    import parsmooth  # noqa: F401  (the unmodified upstream package)
    from parsmooth._base import MVNSqrt, FunctionalModel, ConditionalMomentsModel
    from parsmooth._utils import _cholesky_update, cholesky_update_many, tria, mvn_loglikelihood
    from parsmooth.linearization import extended, cubature, gauss_hermite
    from parsmooth.linearization._gh import _gauss_hermite_weights
    from parsmooth.linearization._cubature import _cubature_weights
    from parsmooth.methods import filtering, smoothing, filter_smoother, iterated_smoothing
    from parsmooth.parallel._operators import sqrt_filtering_operator, sqrt_smoothing_operator
    from parsmooth.parallel._filtering import _sqrt_associative_params_one, _sqrt_loglikelihood
    from parsmooth.parallel._smoothing import _sqrt_associative_params as _sqrt_smoothing_params
    from tests._lgssm import transition_function as lgssm_f, observation_function as lgssm_h
    from tests.bearings.bearings_utils import make_parameters as bearings_parameters
    import tests.test_linearization as tl

    out = {}
    A = lambda x: np.asarray(x, dtype=np.float64)
    rng = np.random.RandomState(2024)
    tril = lambda n: np.tril(rng.rand(n, n)) + 0.5 * np.eye(n)

    # 1. operators -------------------------------------------------------------------------------
    for n in (1, 2, 3, 5):
        e1 = (rng.randn(n, n), rng.randn(n), tril(n), rng.randn(n), tril(n))
        e2 = (rng.randn(n, n), rng.randn(n), tril(n), rng.randn(n), tril(n))
        res = sqrt_filtering_operator(e1, e2)
        for i, v in enumerate(e1):
            out[f"fop{n}_e1_{i}"] = A(v)
        for i, v in enumerate(e2):
            out[f"fop{n}_e2_{i}"] = A(v)
        for i, v in enumerate(res):
            out[f"fop{n}_out_{i}"] = A(v)
        s1, s2 = (rng.randn(n), rng.randn(n, n), tril(n)), (rng.randn(n), rng.randn(n, n), tril(n))
        sres = sqrt_smoothing_operator(s1, s2)
        for i in range(3):
            out[f"sop{n}_e1_{i}"], out[f"sop{n}_e2_{i}"], out[f"sop{n}_out_{i}"] = A(s1[i]), A(s2[i]), A(sres[i])

    # 2. math utils ------------------------------------------------------------------------------
    for n in (2, 3, 5):
        L, v, V = tril(n), 0.3 * rng.randn(n), 0.3 * rng.rand(3, n)
        out[f"chol{n}_L"], out[f"chol{n}_v"], out[f"chol{n}_V"] = L, v, V
        out[f"chol{n}_up"] = A(_cholesky_update(L, v, 1.0))
        out[f"chol{n}_down"] = A(_cholesky_update(L, v, -0.1))
        out[f"chol{n}_many"] = A(cholesky_update_many(L, V, -1.0))
        out[f"chol{n}_bad"] = A(cholesky_update_many(0.1 * L, 10 * V, -1.0))   # non-finite -> 0 guard
        M = rng.randn(n, 2 * n)
        out[f"tria{n}_in"], out[f"tria{n}_out"] = M, A(tria(M))
        x = rng.randn(n)
        out[f"mvn{n}_x"], out[f"mvn{n}_ll"] = x, A(mvn_loglikelihood(x, L))
    for n, p in ((1, 3), (2, 3), (5, 3), (2, 5)):
        wm, wc, xi = _gauss_hermite_weights(n, p)
        out[f"gh{n}_{p}_wm"], out[f"gh{n}_{p}_xi"] = A(wm), A(xi)
    wm, wc, xi = _cubature_weights(5)
    out["cub5_wm"], out["cub5_xi"] = A(wm), A(xi)

    # 3. LGSSM end to end (parallel / sequential, three linearisations) ------------------------------
    for (n, ny, T) in ((2, 1, 12), (3, 2, 15), (1, 3, 9), (4, 2, 20)):
        Fm = 0.9 * np.linalg.qr(rng.randn(n, n))[0]
        Hm = rng.randn(ny, n)
        cQ, cR = 0.3 * tril(n), 0.4 * tril(ny)
        b, c, m0, L0 = 0.1 * rng.randn(n), 0.1 * rng.randn(ny), rng.randn(n), tril(n)
        ys = rng.randn(T, ny)
        nom = MVNSqrt(rng.randn(T + 1, n), np.repeat(np.eye(n)[None], T + 1, 0))
        tag = f"lg{n}{ny}"
        for k, v in dict(F=Fm, H=Hm, cQ=cQ, cR=cR, b=b, c=c, m0=m0, L0=L0, ys=ys, nom_m=nom.mean).items():
            out[f"{tag}_{k}"] = A(v)
        tm = FunctionalModel(partial(lgssm_f, A=Fm), MVNSqrt(b, cQ))
        om = FunctionalModel(partial(lgssm_h, H=Hm), MVNSqrt(c, cR))
        x0 = MVNSqrt(m0, L0)
        for lname, lin in (("ext", extended), ("cub", cubature), ("gh", gauss_hermite)):
            if lname == "gh" and n > 3:
                continue
            for par in (True, False):
                f, ell = filtering(ys, x0, tm, om, lin, nom, par, return_loglikelihood=True)
                s = smoothing(tm, f, lin, nom, par)
                p = "par" if par else "seq"
                out[f"{tag}_{lname}_{p}_fm"], out[f"{tag}_{lname}_{p}_fc"] = A(f.mean), A(f.chol)
                out[f"{tag}_{lname}_{p}_sm"], out[f"{tag}_{lname}_{p}_sc"] = A(s.mean), A(s.chol)
                out[f"{tag}_{lname}_{p}_ell"] = A(ell)
        # element construction and the ell term on their own
        el, ssm = _sqrt_associative_params_one(extended, tm, om, MVNSqrt(nom.mean[0], nom.chol[0]),
                                               MVNSqrt(nom.mean[1], nom.chol[1]), m0, L0, ys[0])
        for i, v in enumerate(el):
            out[f"{tag}_elem_{i}"] = A(v)
        out[f"{tag}_ellterm"] = A(_sqrt_loglikelihood(*ssm, m0, L0, ys[0]))
        sel = _sqrt_smoothing_params(extended, tm, MVNSqrt(nom.mean[0], nom.chol[0]), m0, L0)
        for i, v in enumerate(sel):
            out[f"{tag}_selem_{i}"] = A(v)

    # 4. bearings-only model: linearisations and iterated smoothers ----------------------------------
    s1, s2 = np.array([-1.5, 0.5]), np.array([1.0, 1.0])
    Q, R, obs_f, trans_f = bearings_parameters(0.01, 0.1, 0.5, 0.01, s1, s2)
    Q, R = A(Q), A(R)
    cQ, cR = np.linalg.cholesky(Q), np.linalg.cholesky(R)
    tm = FunctionalModel(trans_f, MVNSqrt(np.zeros(5), cQ))
    om = FunctionalModel(obs_f, MVNSqrt(np.zeros(2), cR))
    pts_m = np.array([[-1.0, -1.0, 6.0, 4.0, 2.0], [0.3, -2.0, 1.0, 0.5, 1e-8], [2.0, 1.5, -3.0, 0.2, -0.7]])
    pts_L = np.stack([np.eye(5), 0.3 * tril(5), 0.1 * tril(5)])
    out["bear_pts_m"], out["bear_pts_L"] = pts_m, pts_L
    for lname, lin in (("ext", extended), ("cub", cubature), ("gh", gauss_hermite)):
        for mname, model in (("t", tm), ("o", om)):
            res = [lin(model, MVNSqrt(pts_m[i], pts_L[i])) for i in range(3)]
            for j, nm in enumerate(("F", "chol", "b")):
                out[f"bear_{lname}_{mname}_{nm}"] = A(np.stack([A(r[j]) for r in res]))
    T = 40
    ys = np.load(os.path.join(REFERENCE, "tests", "bearings", "ys.npy")).astype(np.float64)[:T]
    out["bear_ys"] = ys
    x0 = MVNSqrt(np.array([-1.0, -1.0, 0.0, 0.0, 0.0]), np.eye(5))
    for lname, lin, iters in (("ext", extended, 4), ("cub", cubature, 3)):
        res, ell = iterated_smoothing(ys, x0, tm, om, lin, None, True, criterion=lambda i, *_: i < iters,
                                      return_loglikelihood=True)
        out[f"bear_it_{lname}_m"], out[f"bear_it_{lname}_c"], out[f"bear_it_{lname}_ell"] = A(res.mean), A(res.chol), A(ell)
        res2 = filter_smoother(ys, x0, tm, om, lin, None, False)       # sequential, running-estimate nominal
        out[f"bear_seq_{lname}_m"], out[f"bear_seq_{lname}_c"] = A(res2.mean), A(res2.chol)

    # 5. population model (conditional moments), reference's own test helpers ------------------------
    lam = 10.0
    tmod = ConditionalMomentsModel(tl.transition_mean, tl.transition_chol)
    omod = ConditionalMomentsModel(partial(tl.observation_mean, lam=lam), partial(tl.observation_chol, lam=lam))
    pm, pL = np.array([[np.log(7.0)], [0.5], [2.2]]), np.array([[[1.0]], [[0.3]], [[0.05]]])
    out["pop_pts_m"], out["pop_pts_L"] = pm, pL
    for lname, lin in (("ext", extended), ("cub", cubature), ("gh", gauss_hermite)):
        for mname, model in (("t", tmod), ("o", omod)):
            res = [lin(model, MVNSqrt(pm[i], pL[i])) for i in range(3)]
            for j, nm in enumerate(("F", "chol", "b")):
                out[f"pop_{lname}_{mname}_{nm}"] = A(np.stack([A(r[j]) for r in res]))

    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
