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


# ---- reverse mode (psqrt_loglik_adjoint): compose_adj / apply_adj / k_adj_grad, formula for formula ------------------
def adj_map(felem_map, rec):
    """forward step record -> adjoint map (Psi, v, D):  lam = Psi^T lam' + v,  Lam = Psi^T Lam' Psi + sym((Psi^T lam') v^T) + D"""
    Phi, w, _, _ = felem_map
    B, _ = rec
    return Phi, w, 0.5 * (B + B.T)


def compose_adj(a, b):
    """G_b o G_a: a = the LATER steps (applied first going backwards), b = the earlier ones"""
    Pa, wa, Da = a
    Pb, wb, Db = b
    t = Pb.T @ wa
    return Pa @ Pb, wb + t, Db + 0.5 * (np.outer(t, wb) + np.outer(wb, t)) + Pb.T @ Da @ Pb


def apply_adj(a, lam, Lam):
    P, w, D = a
    t = P.T @ lam
    return t + w, P.T @ Lam @ P + 0.5 * (np.outer(t, w) + np.outer(w, t)) + D


def adj_grad(F, cQ, b, H, cR, c, y, m, L, l1, L1):
    """d ell / d (F, Q, b, H, R, c) of one step from the costates (l1, L1) of the NEXT filtered state (k_adj_grad);
    Q, R in covariance form."""
    n = F.shape[0]
    P = L @ L.T
    cQ = np.tril(cQ)
    Q, R = cQ @ cQ.T, cR @ cR.T
    mp = F @ m + b
    FP = F @ P
    Pp = FP @ F.T + Q
    HPp = H @ Pp
    S = HPp @ H.T + R
    S = 0.5 * (S + S.T)
    s = np.linalg.solve(S, y - H @ mp - c)
    Kt = np.linalg.solve(S, HPp)
    Phiu = np.eye(n) - Kt.T @ H
    h = H.T @ s
    u = Phiu.T @ l1
    kl = Kt @ l1
    Si = np.linalg.inv(S)
    sym = lambda X: 0.5 * (X + X.T)
    GA = Phiu.T @ L1 @ Phiu + sym(np.outer(u, h)) + 0.5 * np.outer(h, h) - 0.5 * H.T @ Si @ H
    gQ = sym(GA)
    gb = u + h
    gF = np.outer(gb, m) + 2.0 * gQ @ FP
    gc = s - kl
    KtL = Kt @ L1
    gR = sym(KtL @ Kt.T) - sym(np.outer(kl, s)) + 0.5 * (np.outer(s, s) - sym(Si))
    gH = -2.0 * KtL @ Phiu @ Pp + np.outer(s, Pp @ u + mp + Pp @ h) - np.outer(kl, mp + Pp @ h) - Kt
    return gF, gQ, gb, gH, gR, gc
