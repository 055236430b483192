"""NumPy mirror of the affine tangent maps csrc/psqrt_tangent.cu builds and scans (k_felem, k_selem, compose, apply),
formula for formula: the CPU check of the algebra the kernels rely on (tests/test_tangent_maps.py compares it with
the oracle's direct differentiation, oracle/parsmooth_np.py seq_filter_smoother_jvp)."""
import numpy as np


def felem(F, cQ, b, H, cR, c, y, m, L, dF, dQ, db, dH, dR, dc):
    """-> (Phi, w, cvec, C), (B, e) of one filtering step (k_felem)."""
    n = F.shape[0]
    P = L @ L.T
    cQ = np.tril(cQ)
    Q, R = cQ @ cQ.T, cR @ cR.T
    mp = F @ m + b
    FP = F @ P
    Pp = FP @ F.T + Q
    HPp = H @ Pp
    S = HPp @ H.T + R
    r = y - H @ mp - c
    s = np.linalg.solve(S, r)
    Kt = np.linalg.solve(S, HPp)
    K = Kt.T
    Phiu = np.eye(n) - K @ H
    Phi = Phiu @ F
    h = H.T @ s
    w = F.T @ h
    a0 = dF @ m + db
    X = dF @ FP.T
    A0 = X + X.T + 0.5 * (dQ + dQ.T)
    dHPp = dH @ Pp
    dRs = 0.5 * (dR + dR.T)
    Y = K @ dHPp @ Phiu.T
    C = Phiu @ A0 @ Phiu.T + K @ dRs @ K.T - Y - Y.T
    cvec = Phiu @ (a0 + A0 @ h + dHPp.T @ s) - K @ (dH @ mp + dc + dHPp @ h + dRs @ s)
    HF = H @ F
    B = 0.5 * (np.outer(w, w) - HF.T @ np.linalg.solve(S, HF))
    S0 = H @ A0 @ H.T + dHPp @ H.T + (dHPp @ H.T).T + dRs
    e = s @ (H @ a0 + dH @ mp + dc) + 0.5 * s @ S0 @ s - 0.5 * np.trace(np.linalg.solve(S, S0))
    return (Phi, w, cvec, C), (B, e)


def selem(F, cQ, b, m, L, dm, dP, ms_next, Ls_next, dF, dQ, db):
    """-> (G, 0, g, C) of one smoothing step (k_selem)."""
    n = F.shape[0]
    P = L @ L.T
    cQ = np.tril(cQ)
    Q = cQ @ cQ.T
    Ps = Ls_next @ Ls_next.T
    mp = F @ m + b
    FP = F @ P
    Pp = FP @ F.T + Q
    Gt = np.linalg.solve(Pp, FP)
    dmp = F @ dm + dF @ m + db
    X = dF @ FP.T
    FdP = F @ dP
    dPp = FdP @ F.T + X + X.T + 0.5 * (dQ + dQ.T)
    dGt = np.linalg.solve(Pp, FdP + dF @ P - dPp @ Gt)
    g = dm + dGt.T @ (ms_next - mp) - Gt.T @ dmp
    D = Ps - Pp
    W1 = dGt.T @ (D @ Gt)
    C = dP + W1 + W1.T - Gt.T @ dPp @ Gt
    return Gt.T, np.zeros(n), g, C


def compose(a, b):
    """b o a (a acts first)"""
    P1, w1, c1, C1 = a
    P2, w2, c2, C2 = b
    return P2 @ P1, w1 + P1.T @ w2, P2 @ (c1 + C1 @ w2) + c2, P2 @ C1 @ P2.T + C2


def apply(a, dm, dP):
    P, w, c, C = a
    return P @ (dm + dP @ w) + c, P @ dP @ P.T + C
