// psqrt_tangent.cu -- forward-mode tangent (JVP) of the filter + smoother pass and of its log-likelihood:
// the device side of the gradient path (jax.value_and_grad through parsmooth.methods.iterated_smoothing(...,
// return_loglikelihood=True): methods.py:54-76, the implicit fixed point of _utils.py:103-146, protocol
// notebooks/experiment_bearing_only_param_estimation_run_time.ipynb).
//
// The tangent of a Kalman filter at a fixed primal solution is an AFFINE recursion in (dm, dP): with the primal
// filtered states known (psqrt_filter_smoother), step k maps the incoming tangent to
//     dm' = Phi (dm + dP w) + c,        dP' = Phi dP Phi^T + C,
// where Phi = (I - K H) F is the closed-loop transition, w = F^T H^T S^-1 (y - H m^- - c) and (c, C) collect the model
// tangents (dF, dQ, db, dH, dR, dc).  Maps of this form are closed under composition,
//     (Phi, w, c, C)_2 o (Phi, w, c, C)_1 = (Phi2 Phi1,  w1 + Phi1^T w2,  Phi2 (c1 + C1 w2) + c2,  Phi2 C1 Phi2^T + C2),
// so the whole tangent trajectory is one associative scan over matrix products -- no triangularisation, nothing singular
// when the information factors are rank deficient (which differentiating through the square-root combine would hit).
// The log-likelihood tangent is the sum of w_k . dm_k + <B_k, dP_k> + e_k over the steps.  The RTS smoother's tangent is
// the same kind of recursion backwards, with Phi = G_k (the smoother gain) and w = 0.
//
// Kernels (one tangent direction, one sequence):
//   k_felem   one thread per step: primal (m_k, L_k) + model + model tangents -> map (Phi, w, c, C) and (B, e)
//   k_reduce  one thread per chunk of K steps: ordered composition of the chunk's maps
//   k_bscan   Hillis-Steele scan of up to 1024 maps per CTA (ping-pong in global memory), twice (chunks, CTAs)
//   k_apply   one thread per chunk: prior tangent through the two exclusive prefixes, then step by step through the
//             chunk, writing (dm, dP) and the chunk's share of d ell
//   k_selem   smoothing maps from the primal filtered / smoothed states and the filtered tangents; reverse scan with the
//             same three kernels
//   k_dp2dl   dP -> dL for a lower-triangular factor (L^-1 dP L^-T, lower half, half diagonal)
// plus the tangents of the built-in linearizations (extended and sigma-point SLR) in dual-number arithmetic.
// Covariance-form tangents throughout: dQ = d(cholQ cholQ^T), dR = d(cholR cholR^T), dP = d(L L^T).
#include "../../include/psqrt.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace psq {
namespace tangent {

constexpr int kScanBlock = 256;
constexpr int kElemBlock = 128;

template <int N>
struct Rec {
  static constexpr int TRI = N * (N + 1) / 2;
  static constexpr int NT = N * N + 2 * N + TRI;  // Phi, w, c, C (packed lower)
  static constexpr int NB = TRI + 1;              // B (packed lower), e
};

// ------------------------------------------------------------------------------------------------
// the affine tangent map
// ------------------------------------------------------------------------------------------------
template <int N>
struct Map {
  double P[N][N], w[N], c[N], C[N][N];
  __device__ void identity() {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      w[i] = 0.0; c[i] = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) { P[i][j] = (i == j) ? 1.0 : 0.0; C[i][j] = 0.0; }
    }
  }
  // field f of the packed record at base[f * stride]
  __device__ void load(const double* base, long long stride) {
    int f = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) P[i][j] = base[(f++) * stride];
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = base[(f++) * stride];
#pragma unroll
    for (int i = 0; i < N; ++i) c[i] = base[(f++) * stride];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) { C[i][j] = base[(f++) * stride]; C[j][i] = C[i][j]; }
  }
  __device__ void store(double* base, long long stride) const {
    int f = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) base[(f++) * stride] = P[i][j];
#pragma unroll
    for (int i = 0; i < N; ++i) base[(f++) * stride] = w[i];
#pragma unroll
    for (int i = 0; i < N; ++i) base[(f++) * stride] = c[i];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) base[(f++) * stride] = 0.5 * (C[i][j] + C[j][i]);
  }
};

// out = b o a   (a acts first)
template <int N>
__device__ void compose(const Map<N>& a, const Map<N>& b, Map<N>& out) {
  double t[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = a.c[i];
#pragma unroll
    for (int j = 0; j < N; ++j) s = fma(a.C[i][j], b.w[j], s);
    t[i] = s;
  }
  double BC[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0, p = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) { s = fma(b.P[i][k], a.C[k][j], s); p = fma(b.P[i][k], a.P[k][j], p); }
      BC[i][j] = s;
      out.P[i][j] = p;
    }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = b.c[i], ww = a.w[i];
#pragma unroll
    for (int j = 0; j < N; ++j) { s = fma(b.P[i][j], t[j], s); ww = fma(a.P[j][i], b.w[j], ww); }
    out.c[i] = s;
    out.w[i] = ww;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double q = b.C[i][j];
#pragma unroll
      for (int k = 0; k < N; ++k) q = fma(BC[i][k], b.P[j][k], q);
      out.C[i][j] = q;
    }
  }
}

// (dm, dP) <- map(dm, dP)
template <int N>
__device__ void apply(const Map<N>& a, double (&dm)[N], double (&dP)[N][N]) {
  double t[N], PS[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = dm[i];
#pragma unroll
    for (int j = 0; j < N; ++j) s = fma(dP[i][j], a.w[j], s);
    t[i] = s;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = a.c[i];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      s = fma(a.P[i][j], t[j], s);
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) q = fma(a.P[i][k], dP[k][j], q);
      PS[i][j] = q;
    }
    dm[i] = s;
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double q = 0.5 * (a.C[i][j] + a.C[j][i]);
#pragma unroll
      for (int k = 0; k < N; ++k) q = fma(PS[i][k], a.P[j][k], q);
      dP[i][j] = q;
      dP[j][i] = q;
    }
}

// ------------------------------------------------------------------------------------------------
// The ADJOINT family (reverse mode, psqrt_loglik_adjoint): costates (lam, Lam) = d ell / d (m_k, P_k) obey the transposed
// recursion backwards in time,
//     lam_k = Phi^T lam_{k+1} + w,     Lam_k = Phi^T Lam_{k+1} Phi + sym((Phi^T lam_{k+1}) w^T) + B,
// again closed under composition; a Map holds (P = Psi, w = v, C = D; c unused):
//     lam = Psi^T lam' + v,   Lam = Psi^T Lam' Psi + sym((Psi^T lam') v^T) + D.
// ------------------------------------------------------------------------------------------------
// out = G_b o G_a for a = LATER steps (applied first going backwards), b = EARLIER step(s)
template <int N>
__device__ void compose_adj(const Map<N>& a, const Map<N>& b, Map<N>& out) {
  double t[N];   // b.P^T a.w
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s = fma(b.P[j][i], a.w[j], s);
    t[i] = s;
    out.c[i] = 0.0;
  }
  double AC[N][N];   // a.C b.P
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0, p = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) { s = fma(a.C[i][k], b.P[k][j], s); p = fma(a.P[i][k], b.P[k][j], p); }
      AC[i][j] = s;
      out.P[i][j] = p;
    }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    out.w[i] = b.w[i] + t[i];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double q = b.C[i][j] + 0.5 * (t[i] * b.w[j] + t[j] * b.w[i]);
#pragma unroll
      for (int k = 0; k < N; ++k) q = fma(b.P[k][i], AC[k][j], q);
      out.C[i][j] = q;
    }
  }
}
// (lam, Lam) <- G(lam, Lam)
template <int N>
__device__ void apply_adj(const Map<N>& a, double (&lam)[N], double (&Lam)[N][N]) {
  double t[N], LP[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) s = fma(a.P[j][i], lam[j], s);
    t[i] = s;
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) q = fma(Lam[i][k], a.P[k][j], q);
      LP[i][j] = q;
    }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    lam[i] = t[i] + a.w[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double q = 0.5 * (a.C[i][j] + a.C[j][i]) + 0.5 * (t[i] * a.w[j] + t[j] * a.w[i]);
#pragma unroll
      for (int k = 0; k < N; ++k) q = fma(a.P[k][i], LP[k][j], q);
      Lam[i][j] = q;
      Lam[j][i] = q;
    }
  }
}
template <int N, bool ADJ>
__device__ __forceinline__ void compose_sel(const Map<N>& a, const Map<N>& b, Map<N>& out) {
  if constexpr (ADJ) compose_adj<N>(a, b, out); else compose<N>(a, b, out);
}
template <int N, bool ADJ>
__device__ __forceinline__ void apply_sel(const Map<N>& a, double (&x)[N], double (&X)[N][N]) {
  if constexpr (ADJ) apply_adj<N>(a, x, X); else apply<N>(a, x, X);
}

// ------------------------------------------------------------------------------------------------
// small dense helpers (row-major, compile-time shapes)
// ------------------------------------------------------------------------------------------------
template <int R, int C>
__device__ void load_mat(const double* p, double (&A)[R][C]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) A[i][j] = p ? p[i * C + j] : 0.0;
}
// the pass reads only the lower triangle of cholQ (include/psqrt.h, psqrt_ssm)
template <int D>
__device__ void zero_upper(double (&A)[D][D]) {
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i + 1; j < D; ++j) A[i][j] = 0.0;
}
template <int R>
__device__ void load_vec(const double* p, double (&a)[R]) {
#pragma unroll
  for (int i = 0; i < R; ++i) a[i] = p ? p[i] : 0.0;
}
// C = A B
template <int R, int K, int C>
__device__ void mm(const double (&A)[R][K], const double (&B)[K][C], double (&O)[R][C]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[i][k], B[k][j], s);
      O[i][j] = s;
    }
}
// C = A B^T
template <int R, int K, int C>
__device__ void mmt(const double (&A)[R][K], const double (&B)[C][K], double (&O)[R][C]) {
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s = fma(A[i][k], B[j][k], s);
      O[i][j] = s;
    }
}
template <int R, int C>
__device__ void mv(const double (&A)[R][C], const double (&x)[C], double (&y)[R]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < C; ++j) s = fma(A[i][j], x[j], s);
    y[i] = s;
  }
}
// y = A^T x
template <int R, int C>
__device__ void mtv(const double (&A)[R][C], const double (&x)[R], double (&y)[C]) {
#pragma unroll
  for (int j = 0; j < C; ++j) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < R; ++i) s = fma(A[i][j], x[i], s);
    y[j] = s;
  }
}
// in-place Cholesky of an SPD matrix, lower factor (entries above the diagonal left untouched)
template <int D>
__device__ void chol_lower(double (&A)[D][D]) {
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma(-A[j][k], A[j][k], d);
    d = sqrt(d);
    A[j][j] = d;
    const double inv = 1.0 / d;
#pragma unroll
    for (int i = j + 1; i < D; ++i) {
      double s = A[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-A[i][k], A[j][k], s);
      A[i][j] = s * inv;
    }
  }
}
// X <- (L L^T)^-1 X for a lower-triangular L, X [D][C]
template <int D, int C>
__device__ void cho_solve(const double (&L)[D][D], double (&X)[D][C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s = X[i][c];
#pragma unroll
      for (int k = 0; k < i; ++k) s = fma(-L[i][k], X[k][c], s);
      X[i][c] = s / L[i][i];
    }
#pragma unroll
    for (int i = D - 1; i >= 0; --i) {
      double s = X[i][c];
#pragma unroll
      for (int k = i + 1; k < D; ++k) s = fma(-L[k][i], X[k][c], s);
      X[i][c] = s / L[i][i];
    }
  }
}

struct ModelPtrs {  // one tangent direction of the linearised model (psqrt_ssm_tangent) next to the model itself
  const double *F, *cholQ, *b, *H, *cholR, *c;
  long long F_ts, cholQ_ts, b_ts, H_ts, cholR_ts, c_ts;
  const double *dF, *dQ, *db, *dH, *dR, *dc;
  long long dF_ts, dQ_ts, db_ts, dH_ts, dR_ts, dc_ts;
};

// record addressing: field f of step j of chunk c at ((f * K + j) * Cn + c)
struct Lay {
  long long T, Cn;
  int K;
};

// ------------------------------------------------------------------------------------------------
// k_felem: the filtering tangent map of every step
// ------------------------------------------------------------------------------------------------
template <int N, int NY>
__global__ void __launch_bounds__(kElemBlock)
k_felem(ModelPtrs mp, const double* __restrict__ y, const double* __restrict__ fm, const double* __restrict__ fL,
        Lay lay, double* __restrict__ rec, double* __restrict__ recB) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lay.Cn * lay.K) return;
  const long long cidx = t % lay.Cn;
  const int j = (int)(t / lay.Cn);
  const long long k = cidx * lay.K + j;
  double* out = rec + (long long)j * lay.Cn + cidx;
  double* outB = recB + (long long)j * lay.Cn + cidx;
  const long long fs = (long long)lay.K * lay.Cn;
  if (k >= lay.T) {   // past the end: identity map, no likelihood term
    Map<N> id;
    id.identity();
    id.store(out, fs);
#pragma unroll
    for (int f = 0; f < Rec<N>::NB; ++f) outB[f * fs] = 0.0;
    return;
  }
  double m[N], L[N][N], F[N][N], cQ[N][N], b[N], H[NY][N], cR[NY][NY], c[NY], yk[NY];
  double dF[N][N], dQ[N][N], db[N], dH[NY][N], dR[NY][NY], dc[NY];
  load_vec<N>(fm + k * N, m);
  load_mat<N, N>(fL + k * N * N, L);
  load_mat<N, N>(mp.F + k * mp.F_ts, F);
  load_mat<N, N>(mp.cholQ + k * mp.cholQ_ts, cQ);
  zero_upper<N>(cQ);
  load_vec<N>(mp.b + k * mp.b_ts, b);
  load_mat<NY, N>(mp.H + k * mp.H_ts, H);
  load_mat<NY, NY>(mp.cholR + k * mp.cholR_ts, cR);
  load_vec<NY>(mp.c + k * mp.c_ts, c);
  load_vec<NY>(y + k * NY, yk);
  load_mat<N, N>(mp.dF ? mp.dF + k * mp.dF_ts : nullptr, dF);
  load_mat<N, N>(mp.dQ ? mp.dQ + k * mp.dQ_ts : nullptr, dQ);
  load_vec<N>(mp.db ? mp.db + k * mp.db_ts : nullptr, db);
  load_mat<NY, N>(mp.dH ? mp.dH + k * mp.dH_ts : nullptr, dH);
  load_mat<NY, NY>(mp.dR ? mp.dR + k * mp.dR_ts : nullptr, dR);
  load_vec<NY>(mp.dc ? mp.dc + k * mp.dc_ts : nullptr, dc);

  double P[N][N], Q[N][N], R[NY][NY];
  mmt<N, N, N>(L, L, P);
  mmt<N, N, N>(cQ, cQ, Q);
  mmt<NY, NY, NY>(cR, cR, R);
  double mpred[N], FP[N][N], Pp[N][N];
  mv<N, N>(F, m, mpred);
#pragma unroll
  for (int i = 0; i < N; ++i) mpred[i] += b[i];
  mm<N, N, N>(F, P, FP);
  mmt<N, N, N>(FP, F, Pp);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Pp[i][q] += Q[i][q];
  double HPp[NY][N], S[NY][NY];
  mm<NY, N, N>(H, Pp, HPp);
  mmt<NY, N, NY>(HPp, H, S);
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < NY; ++q) S[i][q] += R[i][q];
  double Ls[NY][NY];
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < NY; ++q) Ls[i][q] = 0.5 * (S[i][q] + S[q][i]);
  chol_lower<NY>(Ls);
  double r[NY], s[NY][1];
  mv<NY, N>(H, mpred, r);
#pragma unroll
  for (int i = 0; i < NY; ++i) { r[i] = yk[i] - r[i] - c[i]; s[i][0] = r[i]; }
  cho_solve<NY, 1>(Ls, s);
  double sv[NY];
#pragma unroll
  for (int i = 0; i < NY; ++i) sv[i] = s[i][0];
  double Kt[NY][N];  // K^T = S^-1 H Pp
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Kt[i][q] = HPp[i][q];
  cho_solve<NY, N>(Ls, Kt);
  double Phiu[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double v = (i == q) ? 1.0 : 0.0;
#pragma unroll
      for (int a = 0; a < NY; ++a) v = fma(-Kt[a][i], H[a][q], v);
      Phiu[i][q] = v;
    }
  Map<N> e;
  mm<N, N, N>(Phiu, F, e.P);
  double h[N];
  mtv<NY, N>(H, sv, h);
  mtv<N, N>(F, h, e.w);
  // model-tangent terms
  double a0[N], X[N][N], A0[N][N];
  mv<N, N>(dF, m, a0);
#pragma unroll
  for (int i = 0; i < N; ++i) a0[i] += db[i];
  mmt<N, N, N>(dF, FP, X);  // dF P F^T
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) A0[i][q] = X[i][q] + X[q][i] + 0.5 * (dQ[i][q] + dQ[q][i]);
  double dHPp[NY][N];  // dH Pp
  mm<NY, N, N>(dH, Pp, dHPp);
  double dRs[NY][NY];
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < NY; ++q) dRs[i][q] = 0.5 * (dR[i][q] + dR[q][i]);
  // C = Phiu A0 Phiu^T + K dR K^T - Y - Y^T,  Y = K dH Pp Phiu^T
  {
    double T1[N][N], T2[N][N];
    mm<N, N, N>(Phiu, A0, T1);
    mmt<N, N, N>(T1, Phiu, T2);
    double KdR[N][NY];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < NY; ++q) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < NY; ++a) v = fma(Kt[a][i], dRs[a][q], v);
        KdR[i][q] = v;
      }
    double G1[NY][N];  // dH Pp Phiu^T
    mmt<NY, N, N>(dHPp, Phiu, G1);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) {
        double v = T2[i][q];
#pragma unroll
        for (int a = 0; a < NY; ++a) {
          v = fma(KdR[i][a], Kt[a][q], v);
          v = fma(-Kt[a][i], G1[a][q], v);
          v = fma(-Kt[a][q], G1[a][i], v);
        }
        e.C[i][q] = v;
      }
  }
  // c = Phiu (a0 + A0 h) - K (dH mpred + dc) + Phiu Pp dH^T s - K (dH Pp h + dR s)
  double u1[NY], u2[NY], u3[NY];
  mv<NY, N>(dH, mpred, u1);
  mv<NY, N>(dHPp, h, u2);
  mv<NY, NY>(dRs, sv, u3);
  {
    double v1[N], v2[N];
    mv<N, N>(A0, h, v1);
    mtv<NY, N>(dHPp, sv, v2);  // Pp dH^T s  (Pp symmetric)
#pragma unroll
    for (int i = 0; i < N; ++i) v1[i] += a0[i] + v2[i];
    mv<N, N>(Phiu, v1, e.c);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double v = e.c[i];
#pragma unroll
      for (int a = 0; a < NY; ++a) v = fma(-Kt[a][i], u1[a] + dc[a] + u2[a] + u3[a], v);
      e.c[i] = v;
    }
  }
  e.store(out, fs);
  // likelihood record: B = (w w^T - (HF)^T S^-1 (HF)) / 2; e = s.(H a0 + dH mpred + dc) + s^T S0 s / 2 - tr(S^-1 S0) / 2
  {
    double HF[NY][N], SiHF[NY][N];
    mm<NY, N, N>(H, F, HF);
#pragma unroll
    for (int i = 0; i < NY; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) SiHF[i][q] = HF[i][q];
    cho_solve<NY, N>(Ls, SiHF);
    int f = 0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q <= i; ++q) {
        double v = e.w[i] * e.w[q];
#pragma unroll
        for (int a = 0; a < NY; ++a) v = fma(-HF[a][i], SiHF[a][q], v);
        // off-diagonal entries count twice in <B, dP> with the packed lower triangle
        outB[(f++) * fs] = (i == q) ? 0.5 * v : v;
      }
    double S0[NY][NY], HA0[NY][N], T3[NY][NY];
    mm<NY, N, N>(H, A0, HA0);
    mmt<NY, N, NY>(HA0, H, S0);
    mmt<NY, N, NY>(dHPp, H, T3);
#pragma unroll
    for (int i = 0; i < NY; ++i)
#pragma unroll
      for (int q = 0; q < NY; ++q) S0[i][q] += T3[i][q] + T3[q][i] + dRs[i][q];
    double Ha0[NY];
    mv<NY, N>(H, a0, Ha0);
    double ev = 0.0;
#pragma unroll
    for (int i = 0; i < NY; ++i) {
      ev = fma(sv[i], Ha0[i] + u1[i] + dc[i], ev);
#pragma unroll
      for (int q = 0; q < NY; ++q) ev = fma(0.5 * sv[i], S0[i][q] * sv[q], ev);
    }
    cho_solve<NY, NY>(Ls, S0);
#pragma unroll
    for (int i = 0; i < NY; ++i) ev -= 0.5 * S0[i][i];
    outB[(long long)Rec<N>::TRI * fs] = ev;
  }
}

// ------------------------------------------------------------------------------------------------
// k_selem: the smoothing tangent map of every step (w = 0)
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(kElemBlock)
k_selem(ModelPtrs mp, const double* __restrict__ fm, const double* __restrict__ fL, const double* __restrict__ sm,
        const double* __restrict__ sL, const double* __restrict__ dfm, const double* __restrict__ dfP, Lay lay,
        double* __restrict__ rec) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lay.Cn * lay.K) return;
  const long long cidx = t % lay.Cn;
  const int j = (int)(t / lay.Cn);
  const long long k = cidx * lay.K + j;
  double* out = rec + (long long)j * lay.Cn + cidx;
  const long long fs = (long long)lay.K * lay.Cn;
  Map<N> e;
  if (k >= lay.T) {
    e.identity();
    e.store(out, fs);
    return;
  }
  double m[N], L[N][N], F[N][N], cQ[N][N], b[N], dF[N][N], dQ[N][N], db[N], dm[N], dP[N][N], ms[N], Ls[N][N];
  load_vec<N>(fm + k * N, m);
  load_mat<N, N>(fL + k * N * N, L);
  load_mat<N, N>(mp.F + k * mp.F_ts, F);
  load_mat<N, N>(mp.cholQ + k * mp.cholQ_ts, cQ);
  zero_upper<N>(cQ);
  load_vec<N>(mp.b + k * mp.b_ts, b);
  load_mat<N, N>(mp.dF ? mp.dF + k * mp.dF_ts : nullptr, dF);
  load_mat<N, N>(mp.dQ ? mp.dQ + k * mp.dQ_ts : nullptr, dQ);
  load_vec<N>(mp.db ? mp.db + k * mp.db_ts : nullptr, db);
  load_vec<N>(dfm + k * N, dm);
  load_mat<N, N>(dfP + k * N * N, dP);
  load_vec<N>(sm + (k + 1) * N, ms);
  load_mat<N, N>(sL + (k + 1) * N * N, Ls);
  double P[N][N], Q[N][N], Ps[N][N];
  mmt<N, N, N>(L, L, P);
  mmt<N, N, N>(cQ, cQ, Q);
  mmt<N, N, N>(Ls, Ls, Ps);
  double mpred[N], FP[N][N], Pp[N][N];
  mv<N, N>(F, m, mpred);
#pragma unroll
  for (int i = 0; i < N; ++i) mpred[i] += b[i];
  mm<N, N, N>(F, P, FP);
  mmt<N, N, N>(FP, F, Pp);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Pp[i][q] += Q[i][q];
  double Lp[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Lp[i][q] = 0.5 * (Pp[i][q] + Pp[q][i]);
  chol_lower<N>(Lp);
  double Gt[N][N];  // G^T = Pp^-1 F P
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Gt[i][q] = FP[i][q];
  cho_solve<N, N>(Lp, Gt);
  // tangents of the prediction
  double dmp[N], t1[N];
  mv<N, N>(F, dm, dmp);
  mv<N, N>(dF, m, t1);
#pragma unroll
  for (int i = 0; i < N; ++i) dmp[i] += t1[i] + db[i];
  double X[N][N], FdP[N][N], dPp[N][N];
  mmt<N, N, N>(dF, FP, X);  // dF P F^T
  mm<N, N, N>(F, dP, FdP);
  mmt<N, N, N>(FdP, F, dPp);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) dPp[i][q] += X[i][q] + X[q][i] + 0.5 * (dQ[i][q] + dQ[q][i]);
  // dG^T = Pp^-1 (F dP + dF P - dPp G^T)        (dP, P, Pp symmetric)
  double dGt[N][N];
  {
    double dFP[N][N], W[N][N];
    mm<N, N, N>(dF, P, dFP);
    mm<N, N, N>(dPp, Gt, W);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) dGt[i][q] = FdP[i][q] + dFP[i][q] - W[i][q];
    cho_solve<N, N>(Lp, dGt);
  }
  double dlt[N];
#pragma unroll
  for (int i = 0; i < N; ++i) dlt[i] = ms[i] - mpred[i];
  // g = dm + dG (ms+ - mpred) - G dmpred
  {
    double a1[N], a2[N];
    mtv<N, N>(dGt, dlt, a1);
    mtv<N, N>(Gt, dmp, a2);
#pragma unroll
    for (int i = 0; i < N; ++i) e.c[i] = dm[i] + a1[i] - a2[i];
  }
  // C = dP + dG D G^T + (dG D G^T)^T - G dPp G^T,   D = Ps+ - Pp
  {
    double D[N][N], DG[N][N], W1[N][N], W2[N][N], W3[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) D[i][q] = Ps[i][q] - Pp[i][q];
    mm<N, N, N>(D, Gt, DG);   // D G^T
    // dG (D G^T) = dGt^T DG
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a) v = fma(dGt[a][i], DG[a][q], v);
        W1[i][q] = v;
      }
    mm<N, N, N>(dPp, Gt, W2);  // dPp G^T
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < N; ++a) v = fma(Gt[a][i], W2[a][q], v);
        W3[i][q] = v;
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) e.C[i][q] = dP[i][q] + W1[i][q] + W1[q][i] - W3[i][q];
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    e.w[i] = 0.0;
#pragma unroll
    for (int q = 0; q < N; ++q) e.P[i][q] = Gt[q][i];
  }
  e.store(out, fs);
}

// ------------------------------------------------------------------------------------------------
// scan: chunk reduce, block scan (twice), apply
// ------------------------------------------------------------------------------------------------
// Scan order: REV = false walks the steps forwards; REV = true backwards (the smoother), chunk c has scan position
// Cn - 1 - c and its steps compose from the last to the first.
template <int N, bool REV, bool ADJ = false>
__global__ void __launch_bounds__(kElemBlock)
k_reduce(const double* __restrict__ rec, Lay lay, double* __restrict__ items) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= lay.Cn) return;
  const long long fs = (long long)lay.K * lay.Cn;
  Map<N> acc, e, o;
  acc.identity();
#pragma unroll 1
  for (int jj = 0; jj < lay.K; ++jj) {
    const int j = REV ? lay.K - 1 - jj : jj;
    e.load(rec + (long long)j * lay.Cn + c, fs);
    compose_sel<N, ADJ>(acc, e, o);
    acc = o;
  }
  const long long pos = REV ? lay.Cn - 1 - c : c;
  acc.store(items + pos, lay.Cn);
}

// Inclusive Hillis-Steele scan of the `count` maps of one CTA's segment (ping-pong between bufA and bufB, both
// [NT][stride]); writes the EXCLUSIVE prefix of every item to pref [NT][stride] and the segment total to
// tot [NT][gridDim.x].
template <int N, bool ADJ = false>
__global__ void __launch_bounds__(kScanBlock)
k_bscan(double* __restrict__ bufA, double* __restrict__ bufB, long long count, long long stride,
        double* __restrict__ pref, double* __restrict__ tot) {
  const long long i = (long long)blockIdx.x * kScanBlock + threadIdx.x;
  const bool have = i < count;
  double* cur = bufA;
  double* nxt = bufB;
  Map<N> a, b, o;
#pragma unroll 1
  for (int d = 1; d < kScanBlock; d <<= 1) {
    if (have) {
      b.load(cur + i, stride);
      if ((int)threadIdx.x >= d) {
        a.load(cur + i - d, stride);
        compose_sel<N, ADJ>(a, b, o);
        o.store(nxt + i, stride);
      } else {
        b.store(nxt + i, stride);
      }
    }
    __syncthreads();
    double* t = cur; cur = nxt; nxt = t;
  }
  if (have) {
    if (threadIdx.x == 0) a.identity(); else a.load(cur + i - 1, stride);
    a.store(pref + i, stride);
    const long long last = ((long long)(blockIdx.x + 1) * kScanBlock < count ? (long long)(blockIdx.x + 1) * kScanBlock : count) - 1;
    if (i == last) {
      b.load(cur + i, stride);
      b.store(tot + blockIdx.x, gridDim.x);
    }
  }
}

template <int N, bool REV, bool ADJ = false>
__global__ void __launch_bounds__(kElemBlock)
k_apply(const double* __restrict__ rec, const double* __restrict__ recB, Lay lay, const double* __restrict__ pref,
        const double* __restrict__ bpref, long long nblk, const double* __restrict__ dm0,
        const double* __restrict__ dP0, double* __restrict__ dm_out, double* __restrict__ dP_out,
        double* __restrict__ dell_part) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= lay.Cn) return;
  const long long fs = (long long)lay.K * lay.Cn;
  const long long pos = REV ? lay.Cn - 1 - c : c;
  double dm[N], dP[N][N];
  load_vec<N>(dm0, dm);
  load_mat<N, N>(dP0, dP);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < i; ++q) { const double v = 0.5 * (dP[i][q] + dP[q][i]); dP[i][q] = v; dP[q][i] = v; }
  Map<N> e;
  e.load(bpref + pos / kScanBlock, nblk);
  apply_sel<N, ADJ>(e, dm, dP);
  e.load(pref + pos, lay.Cn);
  apply_sel<N, ADJ>(e, dm, dP);
  double dell = 0.0;
#pragma unroll 1
  for (int jj = 0; jj < lay.K; ++jj) {
    const int j = REV ? lay.K - 1 - jj : jj;
    const long long k = c * lay.K + j;
    if (k >= lay.T) continue;   // identity maps past the end (REV meets them first, forwards last)
    if (!REV && recB) {
      const double* rb = recB + (long long)j * lay.Cn + c;
      int f = 0;
      double v = rb[(long long)Rec<N>::TRI * fs];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int q = 0; q <= i; ++q) v = fma(rb[(f++) * fs], dP[i][q], v);
      const double* rw = rec + (long long)j * lay.Cn + c + (long long)(N * N) * fs;
#pragma unroll
      for (int i = 0; i < N; ++i) v = fma(rw[i * fs], dm[i], v);
      dell += v;
    }
    e.load(rec + (long long)j * lay.Cn + c, fs);
    apply_sel<N, ADJ>(e, dm, dP);
    const long long o = REV ? k : k + 1;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      dm_out[o * N + i] = dm[i];
#pragma unroll
      for (int q = 0; q < N; ++q) dP_out[(o * N + i) * N + q] = dP[i][q];
    }
  }
  if (!REV && dell_part) dell_part[c] = dell;
}

// ------------------------------------------------------------------------------------------------
// reverse mode: records of the adjoint scan from the forward records, and the per-step gradient contraction
// ------------------------------------------------------------------------------------------------
// recA (Map layout: Phi, w, 0, B) from rec (Phi, w, ., .) and recB (B packed with doubled off-diagonals, e)
template <int N>
__global__ void __launch_bounds__(kElemBlock)
k_adj_prep(const double* rec, const double* __restrict__ recB, Lay lay, double* recA) {   // rec may alias recA
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= lay.Cn * lay.K) return;
  const long long fs = (long long)lay.K * lay.Cn;
  Map<N> e;
  e.load(rec + t, fs);
  int f = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    e.c[i] = 0.0;
#pragma unroll
    for (int q = 0; q <= i; ++q) {
      const double v = recB[t + (long long)(f++) * fs];
      e.C[i][q] = (i == q) ? v : 0.5 * v;
      e.C[q][i] = e.C[i][q];
    }
  }
  e.store(recA + t, fs);
}

// d ell / d (F, Q, b, H, R, c)_k from the costates (lam, Lam)_{k+1} and the primal step (covariance form for Q, R):
//   u = Phiu^T lam', kl = K^T lam', h = H^T s
//   gb = u + h                     gQ = Phiu^T Lam' Phiu + sym(u h^T) + (h h^T - H^T S^-1 H) / 2
//   gF = gb m^T + 2 gQ F P         gc = s - kl
//   gR = K^T Lam' K - sym(kl s^T) + (s s^T - S^-1) / 2
//   gH = -2 K^T Lam' Phiu Pp + s (Pp u + mp + Pp h)^T - kl (mp + Pp h)^T - S^-1 H Pp
template <int N, int NY>
__global__ void __launch_bounds__(kElemBlock)
k_adj_grad(ModelPtrs mp, const double* __restrict__ y, const double* __restrict__ fm, const double* __restrict__ fL,
           const double* __restrict__ lam, const double* __restrict__ Lam, long long T, double* __restrict__ gF,
           double* __restrict__ gQ, double* __restrict__ gb, double* __restrict__ gH, double* __restrict__ gR,
           double* __restrict__ gc) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= T) return;
  double m[N], L[N][N], F[N][N], cQ[N][N], b[N], H[NY][N], cR[NY][NY], c[NY], yk[NY], l1[N], L1[N][N];
  load_vec<N>(fm + k * N, m);
  load_mat<N, N>(fL + k * N * N, L);
  load_mat<N, N>(mp.F + k * mp.F_ts, F);
  load_mat<N, N>(mp.cholQ + k * mp.cholQ_ts, cQ);
  zero_upper<N>(cQ);
  load_vec<N>(mp.b + k * mp.b_ts, b);
  load_mat<NY, N>(mp.H + k * mp.H_ts, H);
  load_mat<NY, NY>(mp.cholR + k * mp.cholR_ts, cR);
  load_vec<NY>(mp.c + k * mp.c_ts, c);
  load_vec<NY>(y + k * NY, yk);
  load_vec<N>(lam + (k + 1) * N, l1);
  load_mat<N, N>(Lam + (k + 1) * N * N, L1);
  double P[N][N], Q[N][N], R[NY][NY];
  mmt<N, N, N>(L, L, P);
  mmt<N, N, N>(cQ, cQ, Q);
  mmt<NY, NY, NY>(cR, cR, R);
  double mpred[N], FP[N][N], Pp[N][N];
  mv<N, N>(F, m, mpred);
#pragma unroll
  for (int i = 0; i < N; ++i) mpred[i] += b[i];
  mm<N, N, N>(F, P, FP);
  mmt<N, N, N>(FP, F, Pp);
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Pp[i][q] += Q[i][q];
  double HPp[NY][N], S[NY][NY], Ls[NY][NY];
  mm<NY, N, N>(H, Pp, HPp);
  mmt<NY, N, NY>(HPp, H, S);
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < NY; ++q) S[i][q] += R[i][q];
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < NY; ++q) Ls[i][q] = 0.5 * (S[i][q] + S[q][i]);
  chol_lower<NY>(Ls);
  double s1[NY][1], sv[NY];
  {
    double r[NY];
    mv<NY, N>(H, mpred, r);
#pragma unroll
    for (int i = 0; i < NY; ++i) s1[i][0] = yk[i] - r[i] - c[i];
    cho_solve<NY, 1>(Ls, s1);
#pragma unroll
    for (int i = 0; i < NY; ++i) sv[i] = s1[i][0];
  }
  double Kt[NY][N];
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) Kt[i][q] = HPp[i][q];
  cho_solve<NY, N>(Ls, Kt);   // K^T = S^-1 H Pp
  double Phiu[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double v = (i == q) ? 1.0 : 0.0;
#pragma unroll
      for (int a = 0; a < NY; ++a) v = fma(-Kt[a][i], H[a][q], v);
      Phiu[i][q] = v;
    }
  double h[N], u[N], kl[NY];
  mtv<NY, N>(H, sv, h);
  mtv<N, N>(Phiu, l1, u);
  mv<NY, N>(Kt, l1, kl);
  double Si[NY][NY];   // S^-1
#pragma unroll
  for (int i = 0; i < NY; ++i)
#pragma unroll
    for (int q = 0; q < NY; ++q) Si[i][q] = (i == q) ? 1.0 : 0.0;
  cho_solve<NY, NY>(Ls, Si);
  // gQ
  double GA[N][N];
  {
    double T1[N][N], SiH[NY][N];
    mm<N, N, N>(L1, Phiu, T1);            // Lam' Phiu
    mm<NY, NY, N>(Si, H, SiH);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) {
        double v = 0.5 * (u[i] * h[q] + u[q] * h[i]) + 0.5 * h[i] * h[q];
#pragma unroll
        for (int a = 0; a < N; ++a) v = fma(Phiu[a][i], T1[a][q], v);
#pragma unroll
        for (int a = 0; a < NY; ++a) v = fma(-0.5 * H[a][i], SiH[a][q], v);
        GA[i][q] = v;
      }
  }
  double ga[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    ga[i] = u[i] + h[i];
    gb[k * N + i] = ga[i];
  }
  {
    double GAs[N][N], G2[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) {
        GAs[i][q] = 0.5 * (GA[i][q] + GA[q][i]);
        gQ[(k * N + i) * N + q] = GAs[i][q];
      }
    mm<N, N, N>(GAs, FP, G2);             // gQ F P
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int q = 0; q < N; ++q) gF[(k * N + i) * N + q] = fma(ga[i], m[q], 2.0 * G2[i][q]);
  }
  // gc, gR
#pragma unroll
  for (int a = 0; a < NY; ++a) gc[k * NY + a] = sv[a] - kl[a];
  double KtL[NY][N];                       // K^T Lam'
  mm<NY, N, N>(Kt, L1, KtL);
#pragma unroll
  for (int a = 0; a < NY; ++a)
#pragma unroll
    for (int q = 0; q < NY; ++q) {
      double v = -0.5 * (kl[a] * sv[q] + kl[q] * sv[a]) + 0.5 * (sv[a] * sv[q] - 0.5 * (Si[a][q] + Si[q][a]));
#pragma unroll
      for (int i = 0; i < N; ++i) v = fma(0.5 * KtL[a][i], Kt[q][i], v);
#pragma unroll
      for (int i = 0; i < N; ++i) v = fma(0.5 * KtL[q][i], Kt[a][i], v);
      gR[(k * NY + a) * NY + q] = v;
    }
  // gH
  {
    double Ppu[N], Pph[N], W[N][N], KLW[NY][N], SiHPp[NY][N];
    mv<N, N>(Pp, u, Ppu);
    mv<N, N>(Pp, h, Pph);
    mm<N, N, N>(Phiu, Pp, W);              // Phiu Pp
    mm<NY, N, N>(KtL, W, KLW);             // K^T Lam' Phiu Pp
#pragma unroll
    for (int a = 0; a < NY; ++a)
#pragma unroll
      for (int q = 0; q < N; ++q) SiHPp[a][q] = Kt[a][q];   // S^-1 H Pp = K^T
#pragma unroll
    for (int a = 0; a < NY; ++a)
#pragma unroll
      for (int q = 0; q < N; ++q)
        gH[(k * NY + a) * N + q] = -2.0 * KLW[a][q] + sv[a] * (Ppu[q] + mpred[q] + Pph[q]) -
                                   kl[a] * (mpred[q] + Pph[q]) - SiHPp[a][q];
  }
}

// fixed-order sum of n partials (one CTA)
__global__ void __launch_bounds__(1024) k_sum(const double* __restrict__ part, long long n, double* __restrict__ out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += part[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int d = 512; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

template <int N>
__global__ void k_copy_state(const double* __restrict__ sm, const double* __restrict__ sP, long long src,
                             double* __restrict__ dm, double* __restrict__ dP, long long dst) {
  const int i = threadIdx.x;
  if (i < N) dm[dst * N + i] = sm ? sm[src * N + i] : 0.0;
  if (i < N * N) dP[dst * N * N + i] = sP ? sP[src * N * N + i] : 0.0;
}

// dP -> dL for lower-triangular L:  dL = L Phi(L^-1 dP L^-T), Phi = strictly lower part + half the diagonal
template <int N>
__global__ void __launch_bounds__(kElemBlock)
k_dp2dl(const double* __restrict__ Lp, const double* __restrict__ dPp, double* __restrict__ dLp, long long count) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  double L[N][N], X[N][N];
  load_mat<N, N>(Lp + t * N * N, L);
  load_mat<N, N>(dPp + t * N * N, X);
  // X <- L^-1 X (columns), then X <- X L^-T (= (L^-1 X^T)^T)
#pragma unroll
  for (int c = 0; c < N; ++c)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = X[i][c];
#pragma unroll
      for (int k = 0; k < i; ++k) s = fma(-L[i][k], X[k][c], s);
      X[i][c] = s / L[i][i];
    }
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = X[r][i];
#pragma unroll
      for (int k = 0; k < i; ++k) s = fma(-L[i][k], X[r][k], s);
      X[r][i] = s / L[i][i];
    }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double v = 0.0;
      if (q <= i) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
          if (k < q || k > i) continue;   // Phi(M) is lower triangular, and so is L
          const double phi = (k == q) ? 0.5 * X[k][k] : 0.5 * (X[k][q] + X[q][k]);
          v = fma(L[i][k], phi, v);
        }
      }
      dLp[(t * N + i) * N + q] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// host side of the pass
// ------------------------------------------------------------------------------------------------
inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

inline Lay make_lay(long long T) {
  Lay l;
  l.T = T;
  long long K = ceil_div(T, (long long)kScanBlock * kScanBlock);
  if (K < 8) K = 8;
  if (K > T) K = T > 0 ? T : 1;
  l.K = (int)K;
  l.Cn = ceil_div(T, K);
  return l;
}

template <int N>
size_t ws_doubles(long long T) {
  const Lay l = make_lay(T);
  const long long nblk = ceil_div(l.Cn, kScanBlock);
  size_t n = 0;
  n += (size_t)Rec<N>::NT * l.K * l.Cn;          // rec
  n += (size_t)Rec<N>::NB * l.K * l.Cn;          // recB
  n += 3 * (size_t)Rec<N>::NT * l.Cn;            // items, ping-pong, pref
  n += 4 * (size_t)Rec<N>::NT * kScanBlock;      // block totals, ping-pong, bpref, (unused) grand total
  n += (size_t)l.Cn + 8;                          // dell partials
  (void)nblk;
  return n;
}

template <int N>
struct Bufs {
  double *rec, *recB, *items, *pp, *pref, *btot, *bpp, *bpref, *gtot, *dellp;
  Bufs(double* ws, const Lay& l) {
    double* p = ws;
    rec = p; p += (size_t)Rec<N>::NT * l.K * l.Cn;
    recB = p; p += (size_t)Rec<N>::NB * l.K * l.Cn;
    items = p; p += (size_t)Rec<N>::NT * l.Cn;
    pp = p; p += (size_t)Rec<N>::NT * l.Cn;
    pref = p; p += (size_t)Rec<N>::NT * l.Cn;
    btot = p; p += (size_t)Rec<N>::NT * kScanBlock;
    bpp = p; p += (size_t)Rec<N>::NT * kScanBlock;
    bpref = p; p += (size_t)Rec<N>::NT * kScanBlock;
    gtot = p; p += (size_t)Rec<N>::NT * kScanBlock;
    dellp = p;
  }
};

template <int N, bool REV, bool ADJ = false>
void run_scan(const Bufs<N>& B, const Lay& l, const double* dm0, const double* dP0, double* dm_out, double* dP_out,
              bool ell, cudaStream_t st) {
  const unsigned cb = (unsigned)ceil_div(l.Cn, kElemBlock);
  const long long nblk = ceil_div(l.Cn, kScanBlock);
  k_reduce<N, REV, ADJ><<<cb, kElemBlock, 0, st>>>(B.rec, l, B.items);
  k_bscan<N, ADJ><<<(unsigned)nblk, kScanBlock, 0, st>>>(B.items, B.pp, l.Cn, l.Cn, B.pref, B.btot);
  k_bscan<N, ADJ><<<1, kScanBlock, 0, st>>>(B.btot, B.bpp, nblk, nblk, B.bpref, B.gtot);
  k_apply<N, REV, ADJ><<<cb, kElemBlock, 0, st>>>(B.rec, ell ? B.recB : nullptr, l, B.pref, B.bpref, nblk, dm0, dP0, dm_out,
                                             dP_out, ell ? B.dellp : nullptr);
}

template <int N, int NY>
int run_pass(const ModelPtrs& mp, const double* y, long long T, const double* fm, const double* fL, const double* sm,
             const double* sL, const double* dm0, const double* dP0, double* dfm, double* dfP, double* dsm,
             double* dsP, double* dell, double* ws, cudaStream_t st) {
  const Lay l = make_lay(T);
  const Bufs<N> B(ws, l);
  const unsigned eb = (unsigned)ceil_div(l.Cn * l.K, kElemBlock);
  k_felem<N, NY><<<eb, kElemBlock, 0, st>>>(mp, y, fm, fL, l, B.rec, B.recB);
  k_copy_state<N><<<1, N * N < 32 ? 32 : N * N, 0, st>>>(dm0, dP0, 0, dfm, dfP, 0);
  run_scan<N, false>(B, l, dm0, dP0, dfm, dfP, dell != nullptr, st);
  if (dell) k_sum<<<1, 1024, 0, st>>>(B.dellp, l.Cn, dell);
  if (dsm && dsP) {
    if (!sm || !sL) return PSQRT_EINVAL;
    k_selem<N><<<eb, kElemBlock, 0, st>>>(mp, fm, fL, sm, sL, dfm, dfP, l, B.rec);
    k_copy_state<N><<<1, N * N < 32 ? 32 : N * N, 0, st>>>(dfm, dfP, T, dsm, dsP, T);
    run_scan<N, true>(B, l, dfm + T * N, dfP + T * N * N, dsm, dsP, false, st);
  }
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

template <int N, int NY>
int run_adjoint(const ModelPtrs& mp, const double* y, long long T, const double* fm, const double* fL, double* lam,
                double* Lam, double* gF, double* gQ, double* gb, double* gH, double* gR, double* gc, double* ws,
                cudaStream_t st) {
  const Lay l = make_lay(T);
  Bufs<N> B(ws, l);
  const unsigned eb = (unsigned)ceil_div(l.Cn * l.K, kElemBlock);
  // forward records with zero model tangents: (Phi, w) and (B, .) of every step
  k_felem<N, NY><<<eb, kElemBlock, 0, st>>>(mp, y, fm, fL, l, B.rec, B.recB);
  // adjoint records overwrite the forward ones (each thread reads its own record before it writes it)
  k_adj_prep<N><<<eb, kElemBlock, 0, st>>>(B.rec, B.recB, l, B.rec);
  k_copy_state<N><<<1, N * N < 32 ? 32 : N * N, 0, st>>>(nullptr, nullptr, 0, lam, Lam, T);
  run_scan<N, true, true>(B, l, nullptr, nullptr, lam, Lam, false, st);
  k_adj_grad<N, NY><<<(unsigned)ceil_div(T, kElemBlock), kElemBlock, 0, st>>>(mp, y, fm, fL, lam, Lam, T, gF, gQ, gb, gH,
                                                                               gR, gc);
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

template <int N>
int dispatch_adjoint(int ny, const ModelPtrs& mp, const double* y, long long T, const double* fm, const double* fL,
                     double* lam, double* Lam, double* gF, double* gQ, double* gb, double* gH, double* gR, double* gc,
                     double* ws, cudaStream_t st) {
  switch (ny) {
    case 1: return run_adjoint<N, 1>(mp, y, T, fm, fL, lam, Lam, gF, gQ, gb, gH, gR, gc, ws, st);
    case 2: return run_adjoint<N, 2>(mp, y, T, fm, fL, lam, Lam, gF, gQ, gb, gH, gR, gc, ws, st);
    case 3: return run_adjoint<N, 3>(mp, y, T, fm, fL, lam, Lam, gF, gQ, gb, gH, gR, gc, ws, st);
    case 4: return run_adjoint<N, 4>(mp, y, T, fm, fL, lam, Lam, gF, gQ, gb, gH, gR, gc, ws, st);
    default: return PSQRT_EUNSUPPORTED;
  }
}

template <int N>
int dispatch_ny(int ny, const ModelPtrs& mp, const double* y, long long T, const double* fm, const double* fL,
                const double* sm, const double* sL, const double* dm0, const double* dP0, double* dfm, double* dfP,
                double* dsm, double* dsP, double* dell, double* ws, cudaStream_t st) {
#define PSQ_TN(NYV)                                                                                                  \
  case NYV:                                                                                                          \
    return run_pass<N, NYV>(mp, y, T, fm, fL, sm, sL, dm0, dP0, dfm, dfP, dsm, dsP, dell, ws, st)
  switch (ny) {
    PSQ_TN(1);
    PSQ_TN(2);
    PSQ_TN(3);
    PSQ_TN(4);
    default: return PSQRT_EUNSUPPORTED;
  }
#undef PSQ_TN
}

// ================================================================================================
// tangents of the built-in linearizations, in dual-number arithmetic
// ================================================================================================
struct Dual {
  double v, d;
  __host__ __device__ Dual() : v(0.0), d(0.0) {}
  __host__ __device__ Dual(double a) : v(a), d(0.0) {}
  __host__ __device__ Dual(double a, double b) : v(a), d(b) {}
};
__device__ inline Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ inline Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ inline Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ inline Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, fma(a.v, b.d, a.d * b.v)); }
__device__ inline Dual operator/(Dual a, Dual b) {
  const double q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ inline Dual dsin(Dual a) { double s, c; sincos(a.v, &s, &c); return Dual(s, c * a.d); }
__device__ inline Dual dcos(Dual a) { double s, c; sincos(a.v, &s, &c); return Dual(c, -s * a.d); }
__device__ inline Dual dexp(Dual a) { const double e = exp(a.v); return Dual(e, e * a.d); }
__device__ inline Dual dsqrt(Dual a) { const double s = sqrt(a.v); return Dual(s, 0.5 * a.d / s); }
__device__ inline Dual datan2(Dual y, Dual x) {
  return Dual(atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / (x.v * x.v + y.v * y.v));
}

// model functors in dual arithmetic: f(x) and, for the extended method, the analytic Jacobian (whose dual part is the
// directional derivative of the Jacobian: the second-order term the tangent of a Taylor linearization needs)
struct CTd {  // tests/bearings/bearings_utils.py:7-46; params {dt}
  static constexpr int NIN = 5, NOUT = 5, NPAR = 1;
  static constexpr bool CONDITIONAL = false;
  Dual dt;
  __device__ void set(const double* p, const double* dp) { dt = Dual(p[0], dp ? dp[0] : 0.0); }
  __device__ void parts(const Dual* x, Dual& cw, Dual& sw, Dual& a, Dual& bq, Dual& da, Dual& db) const {
    const Dual w = x[4];
    const bool small = fabs(w.v) < 1e-6;   // lax.cond: the taken branch returns dt and 0 as constants in w
    sw = dsin(w * dt);
    cw = dcos(w * dt);
    if (small) {
      a = dt; bq = Dual(0.0); da = Dual(0.0); db = Dual(0.0);
    } else {
      const Dual iw = Dual(1.0) / w;
      a = sw * iw;
      bq = (cw - Dual(1.0)) * iw;
      da = (dt * cw * w - sw) * iw * iw;
      db = (-(dt * sw * w) - (cw - Dual(1.0))) * iw * iw;
    }
  }
  __device__ void f(const Dual* x, Dual* y) const {
    Dual cw, sw, a, bq, da, db;
    parts(x, cw, sw, a, bq, da, db);
    y[0] = x[0] + a * x[2] - bq * x[3];
    y[1] = x[1] + bq * x[2] + a * x[3];
    y[2] = cw * x[2] + sw * x[3];
    y[3] = -(sw * x[2]) + cw * x[3];
    y[4] = x[4];
  }
  __device__ void jac(const Dual* x, Dual* y, Dual (&J)[5][5]) const {
    Dual cw, sw, a, bq, da, db;
    parts(x, cw, sw, a, bq, da, db);
    f(x, y);
    const Dual vx = x[2], vy = x[3], dcw = -(dt * sw), dsw = dt * cw;
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 5; ++j) J[i][j] = Dual(0.0);
    J[0][0] = Dual(1.0); J[0][2] = a;  J[0][3] = -bq; J[0][4] = da * vx - db * vy;
    J[1][1] = Dual(1.0); J[1][2] = bq; J[1][3] = a;   J[1][4] = db * vx + da * vy;
    J[2][2] = cw;  J[2][3] = sw;  J[2][4] = dcw * vx + dsw * vy;
    J[3][2] = -sw; J[3][3] = cw;  J[3][4] = -(dsw * vx) + dcw * vy;
    J[4][4] = Dual(1.0);
  }
  __device__ void cholq(const Dual*, Dual (&)[5][5]) const {}
};

struct Bearingsd {  // tests/bearings/bearings_utils.py:49-69; params {s1x, s1y, s2x, s2y}
  static constexpr int NIN = 5, NOUT = 2, NPAR = 4;
  static constexpr bool CONDITIONAL = false;
  Dual s[4];
  __device__ void set(const double* p, const double* dp) {
    for (int i = 0; i < 4; ++i) s[i] = Dual(p[i], dp ? dp[i] : 0.0);
  }
  __device__ void f(const Dual* x, Dual* y) const {
    y[0] = datan2(x[1] - s[1], x[0] - s[0]);
    y[1] = datan2(x[1] - s[3], x[0] - s[2]);
  }
  __device__ void jac(const Dual* x, Dual* y, Dual (&J)[2][5]) const {
    f(x, y);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 5; ++j) J[i][j] = Dual(0.0);
    for (int i = 0; i < 2; ++i) {
      const Dual dx = x[0] - s[2 * i], dy = x[1] - s[2 * i + 1], r2 = dx * dx + dy * dy;
      J[i][0] = -(dy / r2);
      J[i][1] = dx / r2;
    }
  }
  __device__ void cholq(const Dual*, Dual (&)[2][2]) const {}
};

struct Rickerd {  // notebooks/population_model.py:23-62; params {sqrt(Q)}
  static constexpr int NIN = 1, NOUT = 1, NPAR = 1;
  static constexpr bool CONDITIONAL = true;
  Dual sq;
  __device__ void set(const double* p, const double* dp) { sq = Dual(p[0], dp ? dp[0] : 0.0); }
  __device__ void f(const Dual* x, Dual* y) const { y[0] = Dual(3.7999735016195233) + x[0] - dexp(x[0]); }
  __device__ void jac(const Dual* x, Dual* y, Dual (&J)[1][1]) const { f(x, y); J[0][0] = Dual(1.0) - dexp(x[0]); }
  __device__ void cholq(const Dual*, Dual (&C)[1][1]) const { C[0][0] = sq; }
};

struct Poissond {  // notebooks/population_model.py:84-129; params {lam}
  static constexpr int NIN = 1, NOUT = 1, NPAR = 1;
  static constexpr bool CONDITIONAL = true;
  Dual lam;
  __device__ void set(const double* p, const double* dp) { lam = Dual(p[0], dp ? dp[0] : 0.0); }
  __device__ void f(const Dual* x, Dual* y) const { y[0] = lam * dexp(x[0]); }
  __device__ void jac(const Dual* x, Dual* y, Dual (&J)[1][1]) const { f(x, y); J[0][0] = y[0]; }
  __device__ void cholq(const Dual* x, Dual (&C)[1][1]) const { C[0][0] = dsqrt(lam * dexp(x[0])); }
};

struct ParamPack { double p[4], dp[4]; };

// extended (linearization/_extended.py:51-70): dF = d/deps J(m + eps dm; p + eps dp), db = d/deps (f - J m + m_q);
// conditional-moments models also dQ = d/deps chol chol^T
template <class M>
__global__ void k_lin_ext_tan(ParamPack pk, const double* __restrict__ nom_m, const double* __restrict__ dnom_m,
                              long long count, const double* __restrict__ dm_q, double* __restrict__ dF,
                              double* __restrict__ dQ, double* __restrict__ db) {
  constexpr int N = M::NIN, D = M::NOUT;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  M model;
  model.set(pk.p, pk.dp);
  Dual x[N], y[D], J[D][N];
  for (int k = 0; k < N; ++k) x[k] = Dual(nom_m[i * N + k], dnom_m ? dnom_m[i * N + k] : 0.0);
  model.jac(x, y, J);
  for (int r = 0; r < D; ++r) {
    Dual s = y[r];
    for (int k = 0; k < N; ++k) {
      s = s - J[r][k] * x[k];
      dF[(i * D + r) * N + k] = J[r][k].d;
    }
    db[i * D + r] = s.d + (dm_q ? dm_q[r] : 0.0);
  }
  if (M::CONDITIONAL) {
    Dual C[D][D];
    model.cholq(x, C);
    for (int r = 0; r < D; ++r)
      for (int q = 0; q < D; ++q) {
        Dual s(0.0);
        for (int k = 0; k < D; ++k) s = s + C[r][k] * C[q][k];
        dQ[(i * D + r) * D + q] = s.d;
      }
  }
}

// sigma-point SLR (linearization/_sigma_points.py:25-100) in covariance form: with points m + L xi_i and
// Xi = sum_i wc_i xi_i (f_i - fbar)^T,   F = (L^-T Xi)^T,   Omega = sum_i wc_i (f_i - fbar)(f_i - fbar)^T - Xi^T Xi + Q
// (Q = chol_q chol_q^T, or sum_i wc_i chol(x_i) chol(x_i)^T for conditional-moments models), b = fbar - F m + m_q.
template <class M>
__global__ void k_lin_slr_tan(ParamPack pk, const double* __restrict__ xi, const double* __restrict__ wm,
                              const double* __restrict__ wc, int P, const double* __restrict__ nom_m,
                              const double* __restrict__ nom_L, const double* __restrict__ dnom_m,
                              const double* __restrict__ dnom_L, long long count, const double* __restrict__ dm_q,
                              const double* __restrict__ dQ_q, double* __restrict__ dF, double* __restrict__ dQ,
                              double* __restrict__ db) {
  constexpr int N = M::NIN, D = M::NOUT;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  M model;
  model.set(pk.p, pk.dp);
  Dual m[N], L[N][N];
  for (int k = 0; k < N; ++k) m[k] = Dual(nom_m[i * N + k], dnom_m ? dnom_m[i * N + k] : 0.0);
  for (int r = 0; r < N; ++r)
    for (int q = 0; q < N; ++q)
      L[r][q] = (q <= r) ? Dual(nom_L[(i * N + r) * N + q], dnom_L ? dnom_L[(i * N + r) * N + q] : 0.0) : Dual(0.0);
  Dual fbar[D];
  for (int r = 0; r < D; ++r) fbar[r] = Dual(0.0);
  Dual x[N], y[D];
  for (int p = 0; p < P; ++p) {
    for (int r = 0; r < N; ++r) {
      Dual s = m[r];
      for (int q = 0; q <= r; ++q) s = s + L[r][q] * Dual(xi[p * N + q]);
      x[r] = s;
    }
    model.f(x, y);
    for (int r = 0; r < D; ++r) fbar[r] = fbar[r] + Dual(wm[p]) * y[r];
  }
  Dual Xi[N][D], Om[D][D];
  for (int r = 0; r < N; ++r)
    for (int q = 0; q < D; ++q) Xi[r][q] = Dual(0.0);
  for (int r = 0; r < D; ++r)
    for (int q = 0; q < D; ++q) Om[r][q] = Dual(0.0);
  for (int p = 0; p < P; ++p) {
    for (int r = 0; r < N; ++r) {
      Dual s = m[r];
      for (int q = 0; q <= r; ++q) s = s + L[r][q] * Dual(xi[p * N + q]);
      x[r] = s;
    }
    model.f(x, y);
    const Dual w(wc[p]);
    for (int r = 0; r < D; ++r) y[r] = y[r] - fbar[r];
    for (int r = 0; r < N; ++r)
      for (int q = 0; q < D; ++q) Xi[r][q] = Xi[r][q] + w * Dual(xi[p * N + r]) * y[q];
    for (int r = 0; r < D; ++r)
      for (int q = 0; q < D; ++q) Om[r][q] = Om[r][q] + w * y[r] * y[q];
    if (M::CONDITIONAL) {
      Dual C[D][D];
      model.cholq(x, C);
      for (int r = 0; r < D; ++r)
        for (int q = 0; q < D; ++q) {
          Dual s(0.0);
          for (int k = 0; k < D; ++k) s = s + C[r][k] * C[q][k];
          Om[r][q] = Om[r][q] + w * s;
        }
    }
  }
  for (int r = 0; r < D; ++r)
    for (int q = 0; q < D; ++q) {
      Dual s = Om[r][q];
      for (int k = 0; k < N; ++k) s = s - Xi[k][r] * Xi[k][q];
      dQ[(i * D + r) * D + q] = s.d + ((!M::CONDITIONAL && dQ_q) ? dQ_q[r * D + q] : 0.0);
    }
  // F^T = L^-T Xi: back substitution with the upper-triangular L^T, column by column
  Dual Ft[N][D];
  for (int q = 0; q < D; ++q)
    for (int r = N - 1; r >= 0; --r) {
      Dual s = Xi[r][q];
      for (int k = r + 1; k < N; ++k) s = s - L[k][r] * Ft[k][q];
      Ft[r][q] = s / L[r][r];
    }
  for (int r = 0; r < D; ++r) {
    Dual s = fbar[r];
    for (int k = 0; k < N; ++k) {
      s = s - Ft[k][r] * m[k];
      dF[(i * D + r) * N + k] = Ft[k][r].d;
    }
    db[i * D + r] = s.d + ((!M::CONDITIONAL && dm_q) ? dm_q[r] : 0.0);
  }
}

template <class M>
int launch_lin_tan(int lin_id, const double* params, const double* dparams, const double* xi, const double* wm,
                   const double* wc, int P, const double* nom_m, const double* nom_L, const double* dnom_m,
                   const double* dnom_L, long long count, const double* dm_q, const double* dQ_q, double* dF,
                   double* dQ, double* db, cudaStream_t st) {
  ParamPack pk;
  memset(&pk, 0, sizeof(pk));
  for (int i = 0; i < M::NPAR; ++i) { pk.p[i] = params[i]; pk.dp[i] = dparams ? dparams[i] : 0.0; }
  const unsigned blocks = (unsigned)ceil_div(count, 64);
  if (lin_id == PSQRT_LIN_EXTENDED) {
    if (M::CONDITIONAL && !dQ) return PSQRT_EINVAL;
    k_lin_ext_tan<M><<<blocks, 64, 0, st>>>(pk, nom_m, dnom_m, count, dm_q, dF, dQ, db);
  } else if (lin_id == PSQRT_LIN_SLR) {
    if (!xi || !wm || !wc || P <= 0 || !nom_L || !dQ) return PSQRT_EINVAL;
    k_lin_slr_tan<M><<<blocks, 64, 0, st>>>(pk, xi, wm, wc, P, nom_m, nom_L, dnom_m, dnom_L, count, dm_q, dQ_q, dF, dQ,
                                            db);
  } else {
    return PSQRT_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

}  // namespace tangent
}  // namespace psq

using namespace psq::tangent;

extern "C" {

size_t psqrt_tangent_workspace_bytes(int nx, int ny, int64_t T) {
  (void)ny;
  if (T <= 0) return 0;
  switch (nx) {
    case 1: return 8 * ws_doubles<1>(T);
    case 2: return 8 * ws_doubles<2>(T);
    case 3: return 8 * ws_doubles<3>(T);
    case 4: return 8 * ws_doubles<4>(T);
    case 5: return 8 * ws_doubles<5>(T);
    case 6: return 8 * ws_doubles<6>(T);
    case 8: return 8 * ws_doubles<8>(T);
    default: return 0;
  }
}

int psqrt_filter_smoother_tangent(const psqrt_ssm* ssm, const psqrt_ssm_tangent* dssm, const double* y, int nx, int ny,
                                  int64_t T, const double* fm, const double* fL, const double* sm, const double* sL,
                                  const double* dm0, const double* dP0, double* dfm, double* dfP, double* dsm,
                                  double* dsP, double* dell, void* ws, size_t ws_bytes, void* stream) {
  if (!ssm || !dssm || !y || !fm || !fL || !dfm || !dfP || !ws || T <= 0) return PSQRT_EINVAL;
  if (!ssm->F || !ssm->cholQ || !ssm->b || !ssm->H || !ssm->cholR || !ssm->c) return PSQRT_EINVAL;
  if ((dsm == nullptr) != (dsP == nullptr)) return PSQRT_EINVAL;
  const size_t need = psqrt_tangent_workspace_bytes(nx, ny, T);
  if (need == 0) return PSQRT_EUNSUPPORTED;
  if (ws_bytes < need) return PSQRT_EWORKSPACE;
  ModelPtrs mp;
  mp.F = ssm->F; mp.cholQ = ssm->cholQ; mp.b = ssm->b; mp.H = ssm->H; mp.cholR = ssm->cholR; mp.c = ssm->c;
  mp.F_ts = ssm->F_ts; mp.cholQ_ts = ssm->cholQ_ts; mp.b_ts = ssm->b_ts;
  mp.H_ts = ssm->H_ts; mp.cholR_ts = ssm->cholR_ts; mp.c_ts = ssm->c_ts;
  mp.dF = dssm->dF; mp.dQ = dssm->dQ; mp.db = dssm->db; mp.dH = dssm->dH; mp.dR = dssm->dR; mp.dc = dssm->dc;
  mp.dF_ts = dssm->dF_ts; mp.dQ_ts = dssm->dQ_ts; mp.db_ts = dssm->db_ts;
  mp.dH_ts = dssm->dH_ts; mp.dR_ts = dssm->dR_ts; mp.dc_ts = dssm->dc_ts;
  cudaStream_t st = (cudaStream_t)stream;
  double* w = (double*)ws;
#define PSQ_TNX(NV)                                                                                                 \
  case NV:                                                                                                          \
    return dispatch_ny<NV>(ny, mp, y, T, fm, fL, sm, sL, dm0, dP0, dfm, dfP, dsm, dsP, dell, w, st)
  switch (nx) {
    PSQ_TNX(1);
    PSQ_TNX(2);
    PSQ_TNX(3);
    PSQ_TNX(4);
    PSQ_TNX(5);
    PSQ_TNX(6);
    PSQ_TNX(8);
    default: return PSQRT_EUNSUPPORTED;
  }
#undef PSQ_TNX
}

int psqrt_loglik_adjoint(const psqrt_ssm* ssm, const double* y, int nx, int ny, int64_t T, const double* fm,
                         const double* fL, double* lam, double* Lam, double* gF, double* gQ, double* gb, double* gH,
                         double* gR, double* gc, void* ws, size_t ws_bytes, void* stream) {
  if (!ssm || !y || !fm || !fL || !lam || !Lam || !gF || !gQ || !gb || !gH || !gR || !gc || !ws || T <= 0)
    return PSQRT_EINVAL;
  if (!ssm->F || !ssm->cholQ || !ssm->b || !ssm->H || !ssm->cholR || !ssm->c) return PSQRT_EINVAL;
  const size_t need = psqrt_tangent_workspace_bytes(nx, ny, T);
  if (need == 0) return PSQRT_EUNSUPPORTED;
  if (ws_bytes < need) return PSQRT_EWORKSPACE;
  ModelPtrs mp;
  memset(&mp, 0, sizeof(mp));
  mp.F = ssm->F; mp.cholQ = ssm->cholQ; mp.b = ssm->b; mp.H = ssm->H; mp.cholR = ssm->cholR; mp.c = ssm->c;
  mp.F_ts = ssm->F_ts; mp.cholQ_ts = ssm->cholQ_ts; mp.b_ts = ssm->b_ts;
  mp.H_ts = ssm->H_ts; mp.cholR_ts = ssm->cholR_ts; mp.c_ts = ssm->c_ts;
  cudaStream_t st = (cudaStream_t)stream;
  double* w = (double*)ws;
#define PSQ_ANX(NV)                                                                                                \
  case NV:                                                                                                         \
    return dispatch_adjoint<NV>(ny, mp, y, T, fm, fL, lam, Lam, gF, gQ, gb, gH, gR, gc, w, st)
  switch (nx) {
    PSQ_ANX(1);
    PSQ_ANX(2);
    PSQ_ANX(3);
    PSQ_ANX(4);
    PSQ_ANX(5);
    PSQ_ANX(6);
    PSQ_ANX(8);
    default: return PSQRT_EUNSUPPORTED;
  }
#undef PSQ_ANX
}

int psqrt_cov_tangent_to_chol(const double* L, const double* dP, double* dL, int n, int64_t count, void* stream) {
  if (!L || !dP || !dL || count <= 0) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)ceil_div(count, kElemBlock);
  switch (n) {
    case 1: k_dp2dl<1><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    case 2: k_dp2dl<2><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    case 3: k_dp2dl<3><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    case 4: k_dp2dl<4><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    case 5: k_dp2dl<5><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    case 6: k_dp2dl<6><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    case 8: k_dp2dl<8><<<blocks, kElemBlock, 0, st>>>(L, dP, dL, count); break;
    default: return PSQRT_EUNSUPPORTED;
  }
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

int psqrt_linearize_builtin_tangent(int model_id, const double* model_params, const double* dmodel_params, int lin_id,
                                    const double* xi, const double* wm, const double* wc, int n_points,
                                    const double* nom_m, const double* nom_L, const double* dnom_m,
                                    const double* dnom_L, int64_t count, const double* dm_q, const double* dQ_q,
                                    double* dF, double* dQ, double* db, void* stream) {
  if (!model_params || !nom_m || !dF || !db || count <= 0) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
#define PSQ_LT(M)                                                                                                  \
  return launch_lin_tan<M>(lin_id, model_params, dmodel_params, xi, wm, wc, n_points, nom_m, nom_L, dnom_m, dnom_L, \
                           count, dm_q, dQ_q, dF, dQ, db, st)
  switch (model_id) {
    case PSQRT_MODEL_CT_TRANSITION: PSQ_LT(CTd);
    case PSQRT_MODEL_BEARINGS_OBSERVATION: PSQ_LT(Bearingsd);
    case PSQRT_MODEL_RICKER_TRANSITION: PSQ_LT(Rickerd);
    case PSQRT_MODEL_POISSON_OBSERVATION: PSQ_LT(Poissond);
    default: return PSQRT_EINVAL;
  }
#undef PSQ_LT
}

}  // extern "C"
