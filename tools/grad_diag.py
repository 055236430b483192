"""Diagnostic for the gradient bench scenario: convergence of the iterated smoother and of the Neumann terms of the
implicit derivative, device gradient against a central difference of the primal path.  Run on a GPU box."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sqrt-parallel-smoothers_b200"))
import psqrt
from psqrt import grad as pgrad
from psqrt.models import bearings

dev = torch.device("cuda", 0)
g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
lin = getattr(psqrt.linearization, sys.argv[1] if len(sys.argv) > 1 else "extended")
qdata = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
sx = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0     # sensors at (-sx, 0.5), (sx, 1.0)
init = sys.argv[4] if len(sys.argv) > 4 else "inverse"    # initial nominal: inverted bearings | truth + noise
for T in (1000, 10000, 100000):
    dt, r_true = 0.01, 0.05
    s1, s2 = np.array([-sx, 0.5]), np.array([sx, 1.0])
    _, xs, ys = bearings.get_data_pe(np.array([0.1, 0.2, 1.0, 0.0]), dt, r_true, T, s1, s2, q=qdata, random_state=0)
    ys = ys.astype(np.float64)
    Q, _, obs_f, trans_f = bearings.make_parameters(0.1, 0.1, r_true, dt, s1, s2, r2=0.1)
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    x0 = psqrt.MVNSqrt(np.array([2.0, 0.0, 0.0, 0.0, 0.0]), np.diag([0.5, 0.5, 0.5, 0.5, 1.0]))
    pos = bearings.inverse_bearings(ys, s1, s2)
    nom_m = np.concatenate([np.concatenate([np.zeros((1, 2)), pos], 0), np.zeros((T + 1, 3))], 1)
    if init == "truth":
        nom_m = xs.astype(np.float64) + 0.1 * np.random.RandomState(1).randn(T + 1, 5)
        x0 = psqrt.MVNSqrt(np.array([0.1, 0.2, 1.0, 0.0, 1.0]), np.diag([0.5, 0.5, 0.5, 0.5, 1.0]))
    nominal = psqrt.MVNSqrt(g(nom_m), (np.sqrt(0.1) * torch.eye(5, dtype=torch.float64, device=dev)).expand(T + 1, 5, 5).contiguous())
    prec = 10.0
    om_of = lambda p: psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(np.zeros(2), np.diag([1.0 / p, 0.1])))
    tg = pgrad.Tangents(observation_noise=psqrt.MVNSqrt(None, np.diag([-1.0 / prec ** 2, 0.0])))
    ys_d = g(ys)
    hist = []
    def crit(i, prev, cur, n=20):
        hist.append(float(torch.mean((prev.mean - cur.mean) ** 2)))
        return i < n
    for n_iter in (10, 20):
        hist.clear()
        c = lambda i, p, q_, n=n_iter: crit(i, p, q_, n)
        nomx, ell, dell = pgrad.loglikelihood_jvp(ys_d, x0, tm, om_of(prec), lin, tg, nominal, True, criterion=c)
        tp = pgrad.TangentPass(ys_d, x0, tm, om_of(prec), lin, nomx, tg)
        dx, norms = None, []
        for _ in range(n_iter + 2):
            new = tp.jvp(dx)
            norms.append(float(new[0].abs().max()) if dx is None else float((new[0] - dx[0]).abs().max()))
            dx = new
        h = 1e-4 * prec
        e = [float(psqrt.iterated_smoothing(ys_d, x0, tm, om_of(prec + s * h), lin, nominal, True,
                                            criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)[1]) for s in (1, -1)]
        fd = (e[0] - e[1]) / (2 * h)
        print(f"T={T} {sys.argv[1:]} n_iter={n_iter}: ell={float(ell):.6f} dell={float(dell):.6e} fd={fd:.6e} rel={abs(float(dell)-fd)/abs(fd):.2e}")
        print("   iterate msq changes:", " ".join(f"{v:.1e}" for v in hist[:n_iter]))
        print("   Neumann increments :", " ".join(f"{v:.1e}" for v in norms))
