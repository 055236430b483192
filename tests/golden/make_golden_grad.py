"""Golden vectors for the gradient path (BASELINE config 3): d ell / d prec_r of
``iterated_smoothing(..., return_loglikelihood=True)`` for the parameter-estimation protocol of
notebooks/experiment_bearing_only_param_estimation_run_time.ipynb (R = diag((1 / prec_r)^2, 0.1^2)).

JAX is not installable here, so the reference's own reverse-mode gradient cannot be produced.  What CAN be produced
is the derivative of the UNMODIFIED reference source itself: its `iterated_smoothing` is executed on the NumPy shim
of make_golden.py at prec_r +- h, +- 2h and the Richardson-extrapolated central difference is stored.  With a
fixed number of iterations well past convergence, the unrolled derivative (this file) and the implicit fixed-point
derivative (the reference's custom VJP, _utils.py:108-133, and psqrt.grad) agree to the convergence error.

    python tests/golden/make_golden_grad.py      (build container only: needs /root/reference)

Writes tests/golden/reference_vectors_grad.npz.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import make_golden as mg  # noqa: E402

OUT = os.path.join(HERE, "reference_vectors_grad.npz")


def main():
    if not os.path.isdir(mg.REFERENCE):
        raise SystemExit("needs the reference checkout at /root/reference")
    from _cases import bearings_pe_case     # data only (NumPy)
    mg.install_shim()
    sys.path.insert(0, mg.REFERENCE)
    from parsmooth._base import MVNSqrt, FunctionalModel
    from parsmooth.linearization import extended, cubature, gauss_hermite
    from parsmooth.methods import iterated_smoothing
    from tests.bearings.bearings_utils import make_parameters

    A = lambda v: np.array(np.asarray(v), dtype=np.float64)
    out = {}
    for T, seed, prec in ((60, 0, 12.0), (120, 1, 18.0)):
        case = bearings_pe_case(T, seed)
        Q, _, obs_f, trans_f = make_parameters(case["qc"], case["qw"], 1.0, case["dt"], case["s1"], case["s2"])
        cQ = np.linalg.cholesky(A(Q))
        x0 = MVNSqrt(case["m0"], case["L0"])
        tm = FunctionalModel(trans_f, MVNSqrt(np.zeros(5), cQ))
        for lname, lin, iters in (("ext", extended, 10), ("cub", cubature, 10), ("gh", gauss_hermite, 8)):
            if lname == "gh" and T > 60:
                continue

            def ell_of(p):
                cR = np.diag([1.0 / p, 0.1])          # chol of R = diag(r^2, 0.1^2), r = 1 / prec (the notebook)
                om = FunctionalModel(obs_f, MVNSqrt(np.zeros(2), cR))
                res, ell = iterated_smoothing(case["ys"], x0, tm, om, lin, None, True,
                                              criterion=lambda i, *_: i < iters, return_loglikelihood=True)
                return res, float(A(ell))

            h = 1e-3
            e = {k: ell_of(prec + k * h)[1] for k in (-2, -1, 1, 2)}
            fd1 = (e[1] - e[-1]) / (2 * h)
            fd2 = (e[2] - e[-2]) / (4 * h)
            res, ell = ell_of(prec)
            key = f"pe_T{T}_{lname}"
            out[key + "_prec"], out[key + "_iters"] = np.array(prec), np.array(iters)
            out[key + "_ell"] = np.array(ell)
            out[key + "_dell"] = np.array((4 * fd1 - fd2) / 3)
            out[key + "_dell_h"] = np.array(fd1)            # plain central difference: the two agree to O(h^2)
            out[key + "_m"], out[key + "_c"] = A(res.mean), A(res.chol)
            print(key, ell, out[key + "_dell"], fd1)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
