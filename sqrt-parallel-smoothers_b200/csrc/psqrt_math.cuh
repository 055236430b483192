// psqrt_math.cuh -- register-resident fp64 small-matrix algebra for the square-root
// parallel Kalman filter / RTS smoother (sm_100a).  Everything here is a per-thread,
// fully unrolled template on the state dimension N and observation dimension NY so that
// every matrix entry lives in a register (no local arrays survive unrolling).
//
// The functions are __host__ __device__ so the very same arithmetic can be exercised by the
// CPU-only unit tests (tests/hostcheck) where no GPU exists; the product only ever calls
// them from the CUDA kernels in psqrt_kernels.cuh.
//
// Reference formulas (EEA-sensors/sqrt-parallel-smoothers):
//   tria                     parsmooth/_utils.py:22-24
//   filtering element        parsmooth/parallel/_filtering.py:115-146
//   filtering combine        parsmooth/parallel/_operators.py:43-77
//   smoothing element        parsmooth/parallel/_smoothing.py:72-85
//   smoothing combine        parsmooth/parallel/_operators.py:104-125
//   log-likelihood term      parsmooth/parallel/_filtering.py:149-154, _utils.py:152-160
//   rank-1 Cholesky update   parsmooth/_utils.py:39-81
#pragma once
#include <math.h>

#include <type_traits>

#if defined(__CUDACC__)
#define PSQ_HD __host__ __device__ __forceinline__
#define PSQ_UNROLL _Pragma("unroll")
#else
#define PSQ_HD inline __attribute__((always_inline))
#define PSQ_UNROLL
#endif

namespace psq {

constexpr double kHalfLog2Pi = 0.91893853320467274178;  // log(2*pi)/2

// Branch-free reciprocal / reciprocal square root: hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~2^-20
// relative error) + ONE cubically convergent step -> below 2^-57, no slow-path call.  The cubic step is
// one or two dependent operations shorter than two Newton steps, and these chains sit on the critical
// path of every reflector (-DPSQ_NEWTON_SEEDS restores the two Newton steps).  Inputs here are norms and
// diagonal entries of factors; 0 gives NaN/inf exactly where the reference's division does.
PSQ_HD double rcp_nr(double d) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#if !defined(PSQ_NEWTON_SEEDS)
  // one cubic step: y (1 + e + e^2), e = 1 - d y  (seed error 2^-20 -> 2^-60; 3 dependent operations)
  const double e = fma(-d, y, 1.0);
  return fma(y, fma(e, e, e), y);
#else
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  return fma(y, e, y);
#endif
#else
  return 1.0 / d;
#endif
}
PSQ_HD double rsqrt_nr(double d) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#if !defined(PSQ_NEWTON_SEEDS)
  // one Halley step: y (1 + r/2 + 3 r^2 / 8), r = 1 - d y^2  (seed error 2^-20 -> ~2^-58; 4 dependent operations)
  const double r = fma(-(d * y), y, 1.0);
  return fma(y * r, fma(0.375, r, 0.5), y);
#else
  const double h = 0.5 * d;
  double r = fma(-(h * y), y, 0.5);
  y = fma(y, r, y);
  r = fma(-(h * y), y, 0.5);
  return fma(y, r, y);
#endif
#else
  return 1.0 / sqrt(d);
#endif
}

PSQ_HD double ldg(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// ---------------------------------------------------------------------------------------
// Element containers (packed; lower-triangular factors keep only i >= j).
// ---------------------------------------------------------------------------------------
template <int N>
struct FElem {  // filtering element (A, b, U, eta, Z)      _operators.py:58-59
  static constexpr int TRI = N * (N + 1) / 2;
  static constexpr int NF = N * N + 2 * N + 2 * TRI;
  double v[NF];
  PSQ_HD double& A(int i, int j) { return v[i * N + j]; }
  PSQ_HD double& b(int i) { return v[N * N + i]; }
  PSQ_HD double& U(int i, int j) { return v[N * N + N + i * (i + 1) / 2 + j]; }
  PSQ_HD double& eta(int i) { return v[N * N + N + TRI + i]; }
  PSQ_HD double& Z(int i, int j) { return v[N * N + 2 * N + TRI + i * (i + 1) / 2 + j]; }
  PSQ_HD double A(int i, int j) const { return v[i * N + j]; }
  PSQ_HD double b(int i) const { return v[N * N + i]; }
  PSQ_HD double U(int i, int j) const { return v[N * N + N + i * (i + 1) / 2 + j]; }
  PSQ_HD double eta(int i) const { return v[N * N + N + TRI + i]; }
  PSQ_HD double Z(int i, int j) const { return v[N * N + 2 * N + TRI + i * (i + 1) / 2 + j]; }
  // identity of the filtering operator: (I, 0, 0, 0, 0)
  PSQ_HD void set_identity() {
    PSQ_UNROLL
    for (int f = 0; f < NF; ++f) v[f] = 0.0;
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) A(i, i) = 1.0;
  }
};

template <int N>
struct SElem {  // smoothing element (g, E, D)               _operators.py:118-119
  static constexpr int TRI = N * (N + 1) / 2;
  static constexpr int NF = N + N * N + TRI;
  double v[NF];
  PSQ_HD double& g(int i) { return v[i]; }
  PSQ_HD double& E(int i, int j) { return v[N + i * N + j]; }
  PSQ_HD double& D(int i, int j) { return v[N + N * N + i * (i + 1) / 2 + j]; }
  PSQ_HD double g(int i) const { return v[i]; }
  PSQ_HD double E(int i, int j) const { return v[N + i * N + j]; }
  PSQ_HD double D(int i, int j) const { return v[N + N * N + i * (i + 1) / 2 + j]; }
  // identity of the smoothing operator: (0, I, 0)
  PSQ_HD void set_identity() {
    PSQ_UNROLL
    for (int f = 0; f < NF; ++f) v[f] = 0.0;
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) E(i, i) = 1.0;
  }
};

template <int N>
struct Gauss {  // (mean, lower-triangular sqrt factor)
  static constexpr int TRI = N * (N + 1) / 2;
  double m[N];
  double L[TRI];
  PSQ_HD double& Lc(int i, int j) { return L[i * (i + 1) / 2 + j]; }
  PSQ_HD double Lc(int i, int j) const { return L[i * (i + 1) / 2 + j]; }
};

// ---------------------------------------------------------------------------------------
// tria: Householder reflections from the right (LAPACK dgeqr2 on the transpose).
// Rows [0, NREFL) of M (R x C) become lower-trapezoidal; the reflectors are applied to every
// row below.  After the call M[i][j], j <= i, holds the factor; entries right of the
// diagonal in rows < NREFL are scratch.  A zero tail gives tau = 0 like dlarfg.
// TRIBLK > 0 declares that row r has no non-zeros right of column TRIBLK + r (the second
// block is itself lower triangular), which shortens every reflector.
// ---------------------------------------------------------------------------------------
// Compile-time loop: `#pragma unroll` is only a request, and for the larger blocks (N >= 5) nvcc left the row
// loops of house_rows rolled, which turned the register matrices into local-memory arrays.
template <int I, int E, class F>
PSQ_HD void static_for(F&& f) {
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

template <int R, int C, int NREFL, int TRIBLK = 0>
PSQ_HD void house_rows(double (&M)[R][C]) {
  static_for<0, NREFL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int kend = (TRIBLK > 0) ? ((TRIBLK + j + 1 < C) ? TRIBLK + j + 1 : C) : C;
    if constexpr (j + 1 < kend) {
      const double alpha = M[j][j];
      // two partial sums halve the serial depth of the norm
      double sigma = 0.0, sigma2 = 0.0;
      PSQ_UNROLL
      for (int k = j + 1; k < kend; k += 2) {
        sigma = fma(M[j][k], M[j][k], sigma);
        if (k + 1 < kend) sigma2 = fma(M[j][k + 1], M[j][k + 1], sigma2);
      }
      sigma += sigma2;
      // Branch-free on purpose: a branch around the rsqrt / rcp chain stops the scheduler from
      // overlapping it with the independent dot products below.  Only an all-zero row needs H = I
      // (mask = 0, like dlarfg's tau = 0); a zero tail with alpha != 0 just flips the sign of column j.
      const double q = fma(alpha, alpha, sigma);
      const double mask = (q != 0.0) ? 1.0 : 0.0;
      const double qs = (q != 0.0) ? q : 1.0;
      const double norm = qs * rsqrt_nr(qs);
      const double beta = -copysign(norm, alpha) * mask;
      const double v0 = alpha - beta;  // = alpha + sign(alpha) * norm : no cancellation
      const double s = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;  // 1 / (norm (norm + |alpha|))
      // the tail dot products do not depend on the norm: issued first so they overlap the
      // rsqrt / rcp dependency chain; v0 and s enter last
      static_for<j + 1, R>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        double d = 0.0;
        PSQ_UNROLL
        for (int k = j + 1; k < kend; ++k) d = fma(M[i][k], M[j][k], d);
        d = fma(M[i][j], v0, d) * s;
        M[i][j] = fma(-d, v0, M[i][j]);
        PSQ_UNROLL
        for (int k = j + 1; k < kend; ++k) M[i][k] = fma(-d, M[j][k], M[i][k]);
      });
      M[j][j] = beta;
    }
  });
}

// Rank-K "update" tria([L | W]) with L (N x N) already lower triangular: row j's reflector
// touches only column j of L and the K columns of W.
template <int N, int K, class LAcc>
PSQ_HD void tria_append(LAcc&& Lij, double (&W)[N][K]) {
  PSQ_UNROLL
  for (int j = 0; j < N; ++j) {
    const double alpha = Lij(j, j);
    double sigma = 0.0;
    PSQ_UNROLL
    for (int k = 0; k < K; ++k) sigma = fma(W[j][k], W[j][k], sigma);
    const double q = fma(alpha, alpha, sigma);  // branch-free, see house_rows
    const double mask = (q != 0.0) ? 1.0 : 0.0;
    const double qs = (q != 0.0) ? q : 1.0;
    const double norm = qs * rsqrt_nr(qs);
    const double beta = -copysign(norm, alpha) * mask;
    const double v0 = alpha - beta;
    const double s = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
    PSQ_UNROLL
    for (int i = j + 1; i < N; ++i) {
      double d = 0.0;  // dot product first (independent of the norm), v0 and s last
      PSQ_UNROLL
      for (int k = 0; k < K; ++k) d = fma(W[i][k], W[j][k], d);
      d = fma(Lij(i, j), v0, d) * s;
      Lij(i, j) = fma(-d, v0, Lij(i, j));
      PSQ_UNROLL
      for (int k = 0; k < K; ++k) W[i][k] = fma(-d, W[j][k], W[i][k]);
    }
    Lij(j, j) = beta;
  }
}

// ---------------------------------------------------------------------------------------
// Linearised state-space model of one time step (all row-major, read through the
// read-only path; a stride of 0 upstream makes every thread read the same address).
//   x_{k+1} = F x_k + bq + N(0, Q Q^T),   y_k = H x_{k+1} + c + N(0, R R^T)
// ---------------------------------------------------------------------------------------
struct StepPtrs {
  const double* F;   // [N][N]
  const double* Q;   // [N][N]   any square-root factor of the process noise
  const double* bq;  // [N]
  const double* H;   // [NY][N]
  const double* R;   // [NY][NY] any square-root factor of the observation noise
  const double* c;   // [NY]
  const double* y;   // [NY]
  template <int N> PSQ_HD double fF(int i, int j) const { return ldg(F + i * N + j); }
  template <int N> PSQ_HD double fQ(int i, int j) const { return ldg(Q + i * N + j); }
  PSQ_HD double fb(int i) const { return ldg(bq + i); }
  template <int N> PSQ_HD double fH(int a, int k) const { return ldg(H + a * N + k); }
  template <int NY> PSQ_HD double fR(int a, int q) const { return ldg(R + a * NY + q); }
  PSQ_HD double fc(int a) const { return ldg(c + a); }
  PSQ_HD double fy(int a) const { return ldg(y + a); }
};

// A time-invariant model passed BY VALUE as a kernel parameter: its entries are then constant-bank
// operands of the FP64 instructions (no load instructions, no registers).  `y` stays a pointer.
template <int N, int NY>
struct ModelVals {
  double F[N * N], Q[N * N], bq[N], H[NY * N], R[NY * NY], c[NY];
};
template <int N, int NY>
struct StepVals {
  const ModelVals<N, NY>& m;
  const double* y;
  template <int N_> PSQ_HD double fF(int i, int j) const { return m.F[i * N + j]; }
  template <int N_> PSQ_HD double fQ(int i, int j) const { return m.Q[i * N + j]; }
  PSQ_HD double fb(int i) const { return m.bq[i]; }
  template <int N_> PSQ_HD double fH(int a, int k) const { return m.H[a * N + k]; }
  template <int NY_> PSQ_HD double fR(int a, int q) const { return m.R[a * NY + q]; }
  PSQ_HD double fc(int a) const { return m.c[a]; }
  PSQ_HD double fy(int a) const { return ldg(y + a); }
};

// Transition part only (backward sweep).
template <int N>
struct ModelValsT {
  double F[N * N], Q[N * N], bq[N];
};
template <int N>
struct StepValsT {
  const ModelValsT<N>& m;
  template <int N_> PSQ_HD double fF(int i, int j) const { return m.F[i * N + j]; }
  template <int N_> PSQ_HD double fQ(int i, int j) const { return m.Q[i * N + j]; }
  PSQ_HD double fb(int i) const { return m.bq[i]; }
};

// Shared by the Kalman update of both sweeps: given the predicted factor Np (lower, N x N,
// read through Npij) build Psi = tria([[H Np, R], [Np, 0]])            _filtering.py:126-131
// On return M2 holds Psi11 (rows/cols < NY), Psi21 (rows >= NY, cols < NY) and the posterior
// factor (rows >= NY, cols >= NY, lower).
template <int N, int NY, class P, class NAcc>
PSQ_HD void build_update(const P& p, NAcc&& Npij, double (&M2)[NY + N][N + NY]) {
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double h[N];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) h[k] = p.template fH<N>(a, k);
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) acc = fma(h[k], Npij(k, j), acc);
      M2[a][j] = acc;
    }
    PSQ_UNROLL
    for (int q = 0; q < NY; ++q) M2[a][N + q] = p.template fR<NY>(a, q);
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    PSQ_UNROLL
    for (int j = 0; j < N + NY; ++j) M2[NY + i][j] = (j <= i) ? Npij(i, j) : 0.0;
  }
  house_rows<NY + N, N + NY, NY + N - 1>(M2);
}

// Inverse diagonal of Psi11 (NY divisions shared by all forward substitutions).
template <int N, int NY>
PSQ_HD void psi11_inv_diag(const double (&M2)[NY + N][N + NY], double (&inv)[NY]) {
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) inv[a] = rcp_nr(M2[a][a]);
}

// ---------------------------------------------------------------------------------------
// Sweep 1 -- chunk-summary recursion.  acc <- acc (x) e_k where e_k is the zero-prior
// filtering element of step k (_filtering.py:115-146 with m0 = 0, L0 = 0), evaluated without
// ever forming e_k: the combine _operators.py:58-77 collapses, for a rank-NY information
// factor Z_k, to one square-root Kalman predict+update on (b, U), two N x N products on A and
// a rank-NY append on Z.  (Equality with the generic combine is tested against the oracle.)
//
// The running summary keeps ANY dense square root Y of U U^T (like kalman_step_dense): only the NY
// reflectors of the update that produce Psi11 / Psi21 are loop-carried; the N - 1 reflectors that would
// make the factor lower triangular are applied once, at the end of the chunk (FAcc::to_elem).
//
// `pre` is called once per step with the PREDICT-ONLY summary of the step (F A, F b + bq,
// tria([F Y | Q]), eta, Z -- i.e. acc (x) (F, bq, Q, 0, 0)).  Sweep 1 stores that of the LAST step of a
// chunk: the chunk's smoothing total is built from it (chunk_smoothing_total), so that -- like the
// per-step elements of _smoothing.py:76-85 -- only PREDICTED factors are ever inverted.
// ---------------------------------------------------------------------------------------
template <int N>
struct FAcc {
  static constexpr int TRI = N * (N + 1) / 2;
  double A[N][N], b[N], Y[N][N], eta[N], Z[TRI];
  PSQ_HD double& Zc(int i, int j) { return Z[i * (i + 1) / 2 + j]; }
  PSQ_HD void set_identity() {
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      b[i] = 0.0;
      eta[i] = 0.0;
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) {
        A[i][j] = (i == j) ? 1.0 : 0.0;
        Y[i][j] = 0.0;
      }
    }
    PSQ_UNROLL
    for (int f = 0; f < TRI; ++f) Z[f] = 0.0;
  }
  // packed element with U = tria(Y)
  PSQ_HD void to_elem(FElem<N>& e) const {
    double Yt[N][N];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      e.b(i) = b[i];
      e.eta(i) = eta[i];
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) {
        e.A(i, j) = A[i][j];
        Yt[i][j] = Y[i][j];
      }
    }
    house_rows<N, N, N - 1>(Yt);
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j <= i; ++j) {
        e.U(i, j) = Yt[i][j];
        e.Z(i, j) = Z[i * (i + 1) / 2 + j];
      }
  }
};

struct NoPre {
  template <class... Args>
  PSQ_HD void operator()(Args&&...) const {}
};

template <int N, int NY, class P, class PRE>
PSQ_HD void filter_reduce_step(FAcc<N>& acc, const P& p, PRE&& pre) {
  double F[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);

  double mp[N], FA[N][N], M1[N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(F[i][k], acc.b[k], s);
    mp[i] = s;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double a = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) a = fma(F[i][k], acc.A[k][j], a);
      FA[i][j] = a;
      double u = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) u = fma(F[i][k], acc.Y[k][j], u);
      M1[i][j] = u;
      M1[i][N + j] = (j <= i) ? p.template fQ<N>(i, j) : 0.0;  // cholQ is lower triangular (psqrt.h)
    }
  }
  house_rows<N, 2 * N, N, N>(M1);  // predicted factor = tria([F Y | Q]); Q's triangle shortens every reflector
  pre(FA, mp, M1, acc);

  // [[H Np, R], [Np, 0]]: only the NY reflectors that give Psi11 / Psi21; rows >= NY, columns >= NY then hold a
  // dense square root of the posterior covariance                                    _filtering.py:126-131
  double M2[NY + N][N + NY];
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) s = fma(p.template fH<N>(a, k), M1[k][j], s);
      M2[a][j] = s;
    }
    PSQ_UNROLL
    for (int q = 0; q < NY; ++q) M2[a][N + q] = p.template fR<NY>(a, q);
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    PSQ_UNROLL
    for (int j = 0; j < N + NY; ++j) M2[NY + i][j] = (j <= i) ? M1[i][j] : 0.0;
  }
  house_rows<NY + N, N + NY, NY>(M2);
  double inv[NY];
  psi11_inv_diag<N, NY>(M2, inv);

  // V = Psi11^{-1} (H F A)  (NY x N),  rr = Psi11^{-1} (y - H mp - c)
  double V[NY][N], rr[NY];
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double h[N];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) h[k] = p.template fH<N>(a, k);
    double r = p.fy(a) - p.fc(a);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) r = fma(-h[k], mp[k], r);
    PSQ_UNROLL
    for (int q = 0; q < a; ++q) r = fma(-M2[a][q], rr[q], r);
    rr[a] = r * inv[a];
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(h[k], FA[k][j], s);
      PSQ_UNROLL
      for (int q = 0; q < a; ++q) s = fma(-M2[a][q], V[q][j], s);
      V[a][j] = s * inv[a];
    }
  }
  // A <- FA - Psi21 V ; b <- mp + Psi21 rr ; eta <- eta + V^T rr ; Y <- posterior factor
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double bi = mp[i];
    PSQ_UNROLL
    for (int a = 0; a < NY; ++a) bi = fma(M2[NY + i][a], rr[a], bi);
    acc.b[i] = bi;
    double e = acc.eta[i];
    PSQ_UNROLL
    for (int a = 0; a < NY; ++a) e = fma(V[a][i], rr[a], e);
    acc.eta[i] = e;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = FA[i][j];
      PSQ_UNROLL
      for (int a = 0; a < NY; ++a) s = fma(-M2[NY + i][a], V[a][j], s);
      acc.A[i][j] = s;
      acc.Y[i][j] = M2[NY + i][NY + j];
    }
  }
  // Z <- tria([Z | V^T])
  double W[N][NY];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int a = 0; a < NY; ++a) W[i][a] = V[a][i];
  tria_append<N, NY>([&](int i, int j) -> double& { return acc.Zc(i, j); }, W);
}

// ---------------------------------------------------------------------------------------
// Sweep 2 -- one square-root Kalman step x <- filter(x, step k) that also (SMOOTH) emits the
// smoothing element of step k from the *same* triangularisation:
//   tria([[F L, Q], [L, 0]]) = [[Phi11, 0], [Phi21, D]]                _smoothing.py:76-81
// where Phi11 is the predicted factor the update needs (sequential/_filtering.py:80-86).
// Returns the log-likelihood increment (_filtering.py:149-154) computed from Psi11.
// ---------------------------------------------------------------------------------------
template <int N, int NY, bool SMOOTH, class P>
PSQ_HD double kalman_step(Gauss<N>& x, const P& p, SElem<N>* se) {
  double F[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);
  double mp[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(F[i][k], x.m[k], s);
    mp[i] = s;
  }
  constexpr int RR = SMOOTH ? 2 * N : N;
  double M1[RR][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double u = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) u = fma(F[i][k], x.Lc(k, j), u);
      M1[i][j] = u;
      M1[i][N + j] = (j <= i) ? p.template fQ<N>(i, j) : 0.0;  // cholQ is lower triangular (psqrt.h)
    }
  if (SMOOTH) {
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < 2 * N; ++j) M1[RR - N + i][j] = (j <= i) ? x.Lc(i, j) : 0.0;
  }
  house_rows<RR, 2 * N, N, N>(M1);
  if (SMOOTH) {
    // D = tria(bottom-right block)
    double B[N][N];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) B[i][j] = M1[RR - N + i][N + j];
    house_rows<N, N, N - 1>(B);
    // E = Phi21 Phi11^{-1}  (row-wise back substitution, _smoothing.py:83)
    double inv[N];
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) inv[j] = rcp_nr(M1[j][j]);
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      PSQ_UNROLL
      for (int j = N - 1; j >= 0; --j) {
        double s = M1[RR - N + i][j];
        PSQ_UNROLL
        for (int k = j + 1; k < N; ++k) s = fma(-se->E(i, k), M1[k][j], s);
        se->E(i, j) = s * inv[j];
      }
      double gi = x.m[i];
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) gi = fma(-se->E(i, k), mp[k], gi);
      se->g(i) = gi;                                                    // g = m - E (F m + b)
      PSQ_UNROLL
      for (int j = 0; j <= i; ++j) se->D(i, j) = B[i][j];
    }
  }
  double M2[NY + N][N + NY];
  build_update<N, NY>(p, [&](int i, int j) { return M1[i][j]; }, M2);
  double inv2[NY], rr[NY];
  psi11_inv_diag<N, NY>(M2, inv2);
  double quad = 0.0, detS = 1.0;
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double r = p.fy(a) - p.fc(a);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) r = fma(-p.template fH<N>(a, k), mp[k], r);
    PSQ_UNROLL
    for (int q = 0; q < a; ++q) r = fma(-M2[a][q], rr[q], r);
    rr[a] = r * inv2[a];
    quad = fma(rr[a], rr[a], quad);
    detS *= M2[a][a];
  }
  const double logdet = log(fabs(detS));  // sum_a log|Psi11_aa| with one log (NY <= 4 factors of O(1) size)
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double mi = mp[i];
    PSQ_UNROLL
    for (int a = 0; a < NY; ++a) mi = fma(M2[NY + i][a], rr[a], mi);
    x.m[i] = mi;
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) x.Lc(i, j) = M2[NY + i][NY + j];
  }
  return -0.5 * quad - logdet - NY * kHalfLog2Pi;
}

// ---------------------------------------------------------------------------------------
// The same Kalman step for the inside of a chunk, carrying ANY square root Y of the filtered
// covariance (dense N x N) instead of the lower-triangular one.  Only NY reflectors of the update are
// loop-carried: after them rows >= NY of tria's work array read [Psi21 | Y] with Y Y^T the posterior
// covariance, which is all the next prediction tria([F Y | Q]) needs.  The remaining N - 1 reflectors
// that make the OUTPUT factor lower triangular (_filtering.py:126-131 returns tria's triangle) are
// software-pipelined: each call triangularises the state it was GIVEN (`in_tri`, the output of the
// previous step) while it advances x by one step.  The two streams are independent, which is what lets
// the in-order issue of a warp overlap the reflector of one with the rsqrt / rcp chain of the other.
// LOGLIK = false skips the log-likelihood term (returns 0).
// ---------------------------------------------------------------------------------------
template <int N>
struct GaussD {  // (mean, any dense square-root factor)
  double m[N];
  double Y[N][N];
};

template <int N>
PSQ_HD void gaussd_tri(const GaussD<N>& x, Gauss<N>& out) {
  double Yt[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    out.m[i] = x.m[i];
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) Yt[i][j] = x.Y[i][j];
  }
  house_rows<N, N, N - 1>(Yt);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) out.Lc(i, j) = Yt[i][j];
}

template <int N, int NY, bool LOGLIK, class P>
PSQ_HD double kalman_step_dense(GaussD<N>& x, const P& p, Gauss<N>& in_tri) {
  gaussd_tri<N>(x, in_tri);
  double F[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);
  double mp[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(F[i][k], x.m[k], s);
    mp[i] = s;
  }
  double M1[N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double u = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) u = fma(F[i][k], x.Y[k][j], u);
      M1[i][j] = u;
      M1[i][N + j] = (j <= i) ? p.template fQ<N>(i, j) : 0.0;  // cholQ is lower triangular (psqrt.h)
    }
  house_rows<N, 2 * N, N, N>(M1);
  double M2[NY + N][N + NY];
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double h[N];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) h[k] = p.template fH<N>(a, k);
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) acc = fma(h[k], M1[k][j], acc);
      M2[a][j] = acc;
    }
    PSQ_UNROLL
    for (int q = 0; q < NY; ++q) M2[a][N + q] = p.template fR<NY>(a, q);
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    PSQ_UNROLL
    for (int j = 0; j < N + NY; ++j) M2[NY + i][j] = (j <= i) ? M1[i][j] : 0.0;
  }
  house_rows<NY + N, N + NY, NY>(M2);  // only Psi11 / Psi21; rows >= NY, columns >= NY: the factor Y
  double inv2[NY], rr[NY];
  psi11_inv_diag<N, NY>(M2, inv2);
  double quad = 0.0, detS = 1.0;
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double r = p.fy(a) - p.fc(a);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) r = fma(-p.template fH<N>(a, k), mp[k], r);
    PSQ_UNROLL
    for (int q = 0; q < a; ++q) r = fma(-M2[a][q], rr[q], r);
    rr[a] = r * inv2[a];
    quad = fma(rr[a], rr[a], quad);
    detS *= M2[a][a];
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double mi = mp[i];
    PSQ_UNROLL
    for (int a = 0; a < NY; ++a) mi = fma(M2[NY + i][a], rr[a], mi);
    x.m[i] = mi;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) x.Y[i][j] = M2[NY + i][NY + j];
  }
  if (!LOGLIK) return 0.0;
  return -0.5 * quad - log(fabs(detS)) - NY * kHalfLog2Pi;
}

// Smoothing element of step k from the filtered state alone (used by sweep 3, which
// recomputes instead of re-reading 8(2N^2+N) bytes per step).            _smoothing.py:72-85
template <int N, class P>
PSQ_HD void smoothing_element(const Gauss<N>& x, const P& p, SElem<N>& se) {
  double F[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);
  double mp[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(F[i][k], x.m[k], s);
    mp[i] = s;
  }
  double M1[2 * N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double u = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) u = fma(F[i][k], x.Lc(k, j), u);
      M1[i][j] = u;
      M1[i][N + j] = (j <= i) ? p.template fQ<N>(i, j) : 0.0;  // cholQ is lower triangular (psqrt.h)
    }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < 2 * N; ++j) M1[N + i][j] = (j <= i) ? x.Lc(i, j) : 0.0;
  house_rows<2 * N, 2 * N, N, N>(M1);
  double B[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) B[i][j] = M1[N + i][N + j];
  house_rows<N, N, N - 1>(B);
  double inv[N];
  PSQ_UNROLL
  for (int j = 0; j < N; ++j) inv[j] = rcp_nr(M1[j][j]);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    PSQ_UNROLL
    for (int j = N - 1; j >= 0; --j) {
      double s = M1[N + i][j];
      PSQ_UNROLL
      for (int k = j + 1; k < N; ++k) s = fma(-se.E(i, k), M1[k][j], s);
      se.E(i, j) = s * inv[j];
    }
    double gi = x.m[i];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) gi = fma(-se.E(i, k), mp[k], gi);
    se.g(i) = gi;
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) se.D(i, j) = B[i][j];
  }
}

// ---------------------------------------------------------------------------------------
// Smoothing combine (_operators.py:118-125): e1 = accumulated LATER side, e2 = earlier side.
//   g = E2 g1 + g2 ; E = E2 E1 ; D = tria([E2 D1 | D2])
// ---------------------------------------------------------------------------------------
template <int N>
PSQ_HD SElem<N> smoothing_combine(const SElem<N>& e1, const SElem<N>& e2) {
  SElem<N> o;
  double M[N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double gi = e2.g(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) gi = fma(e2.E(i, k), e1.g(k), gi);
    o.g(i) = gi;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0, d = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(e2.E(i, k), e1.E(k, j), s);
      o.E(i, j) = s;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) d = fma(e2.E(i, k), e1.D(k, j), d);
      M[i][j] = d;
      M[i][N + j] = (j <= i) ? e2.D(i, j) : 0.0;
    }
  }
  house_rows<N, 2 * N, N, N>(M);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) o.D(i, j) = M[i][j];
  return o;
}

// Smoothing "apply": the same combine when only (g1, D1) of the later side is known -- i.e.
// one RTS step x_s(k) from x_s(k+1):  m = E m_s + g ; L = tria([E L_s | D]).
template <int N>
PSQ_HD void smoothing_apply(Gauss<N>& xs, const SElem<N>& e2) {
  double M[N][2 * N], mn[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double gi = e2.g(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) gi = fma(e2.E(i, k), xs.m[k], gi);
    mn[i] = gi;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double d = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) d = fma(e2.E(i, k), xs.Lc(k, j), d);
      M[i][j] = d;
      M[i][N + j] = (j <= i) ? e2.D(i, j) : 0.0;
    }
  }
  house_rows<N, 2 * N, N, N>(M);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    xs.m[i] = mn[i];
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) xs.Lc(i, j) = M[i][j];
  }
}

// ---------------------------------------------------------------------------------------
// Generic filtering combine e1 (x) e2 (_operators.py:58-77), e1 = accumulated EARLIER side.
// FULL = false evaluates only (b, U) of the result, which need only (b1, U1) of e1
// ("apply": this is how a carry-in state is pushed through a chunk summary).
// ---------------------------------------------------------------------------------------
template <int N, bool FULL>
PSQ_HD void filtering_combine_core(const FElem<N>* e1full, const double (&b1)[N], const double (&U1)[N][N],
                                   const FElem<N>& e2, FElem<N>* out, double (&bo)[N], double (&Uo)[N][N]) {
  // Xi = [[U1^T Z2, I], [Z2, 0]]                                         _operators.py:63-64
  double Xi[2 * N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = (i > j ? i : j); k < N; ++k) s = fma(U1[k][i], e2.Z(k, j), s);
      Xi[i][j] = s;
      Xi[i][N + j] = (i == j) ? 1.0 : 0.0;
      Xi[N + i][j] = (j <= i) ? e2.Z(i, j) : 0.0;
      Xi[N + i][N + j] = 0.0;
    }
  // Xi11 = Xi[<N][<N] lower, Xi21 = Xi[N+.][<N], rest = pre-Xi22.  Row j of the top block [U1^T Z2 | I] has
  // nothing right of column N + j (earlier reflectors only fill columns N .. N + j - 1): short reflectors.
  house_rows<2 * N, 2 * N, N, N>(Xi);
  // T1 = Xi11^{-1} U1^T   (N x N; U1^T is upper triangular so T1[i][j] needs k <= ... dense in general)
  double inv[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) inv[i] = rcp_nr(Xi[i][i]);
  double T1[N][N];
  PSQ_UNROLL
  for (int j = 0; j < N; ++j)
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = (j >= i) ? U1[j][i] : 0.0;  // (U1^T)[i][j]
      PSQ_UNROLL
      for (int k = 0; k < i; ++k) s = fma(-Xi[i][k], T1[k][j], s);
      T1[i][j] = s * inv[i];
    }
  // W = A2 T1^T                                                      (S1 of the oracle)
  double W[N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(e2.A(i, k), T1[j][k], s);
      W[i][j] = s;
      W[i][N + j] = (j <= i) ? e2.U(i, j) : 0.0;
    }
  // b = A2 (t - T1^T Xi21^T t) + b2,  t = b1 + U1 U1^T eta2                  _operators.py:71
  double t[N], u[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    PSQ_UNROLL
    for (int k = i; k < N; ++k) s = fma(U1[k][i], e2.eta(k), s);
    u[i] = s;  // U1^T eta2
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = b1[i];
    PSQ_UNROLL
    for (int k = 0; k <= i; ++k) s = fma(U1[i][k], u[k], s);
    t[i] = s;
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(Xi[N + k][i], t[k], s);
    u[i] = s;  // Xi21^T t
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = t[i];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(-T1[k][i], u[k], s);
    t[i] = s;
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = e2.b(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(e2.A(i, k), t[k], s);
    bo[i] = s;
  }
  if (FULL) {
    const FElem<N>& e1 = *e1full;
    // A = A2 A1 - W Xi21^T A1 = (A2 - W Xi21^T) A1                            _operators.py:70
    double G[N][N];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) {
        double s = e2.A(i, j);
        PSQ_UNROLL
        for (int k = 0; k < N; ++k) s = fma(-W[i][k], Xi[N + j][k], s);
        G[i][j] = s;
      }
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
        PSQ_UNROLL
        for (int k = 0; k < N; ++k) s = fma(G[i][k], e1.A(k, j), s);
        out->A(i, j) = s;
      }
    // eta = A1^T (s - Xi21 T1 s) + eta1,  s = eta2 - Z2 Z2^T b1             _operators.py:73-74
    double sv[N], w[N];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = i; k < N; ++k) s = fma(e2.Z(k, i), b1[k], s);
      w[i] = s;  // Z2^T b1
    }
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = e2.eta(i);
      PSQ_UNROLL
      for (int k = 0; k <= i; ++k) s = fma(-e2.Z(i, k), w[k], s);
      sv[i] = s;
    }
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(T1[i][k], sv[k], s);
      w[i] = s;  // T1 s
    }
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = sv[i];
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(-Xi[N + i][k], w[k], s);
      sv[i] = s;
    }
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = e1.eta(i);
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(e1.A(k, i), sv[k], s);
      out->eta(i) = s;
    }
    // Z = tria([A1^T Xi22 | Z1]) with Xi22 = tria(bottom-right block)       _operators.py:75
    double B[N][N];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) B[i][j] = Xi[N + i][N + j];
    house_rows<N, N, N - 1>(B);
    double Zm[N][2 * N];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
        PSQ_UNROLL
        for (int k = j; k < N; ++k) s = fma(e1.A(k, i), B[k][j], s);
        Zm[i][j] = s;
        Zm[i][N + j] = (j <= i) ? e1.Z(i, j) : 0.0;
      }
    house_rows<N, 2 * N, N, N>(Zm);
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j <= i; ++j) out->Z(i, j) = Zm[i][j];
  }
  // U = tria([W | U2])                                                        _operators.py:72
  house_rows<N, 2 * N, N, N>(W);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) Uo[i][j] = (j <= i) ? W[i][j] : 0.0;
}

template <int N>
PSQ_HD FElem<N> filtering_combine(const FElem<N>& e1, const FElem<N>& e2) {
  FElem<N> o;
  double b1[N], U1[N][N], bo[N], Uo[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    b1[i] = e1.b(i);
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) U1[i][j] = (j <= i) ? e1.U(i, j) : 0.0;
  }
  filtering_combine_core<N, true>(&e1, b1, U1, e2, &o, bo, Uo);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    o.b(i) = bo[i];
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) o.U(i, j) = Uo[i][j];
  }
  return o;
}

// x <- x (x) e2 keeping only the state part (mean, factor).
template <int N>
PSQ_HD void filtering_apply(Gauss<N>& x, const FElem<N>& e2) {
  double b1[N], U1[N][N], bo[N], Uo[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    b1[i] = x.m[i];
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) U1[i][j] = (j <= i) ? x.Lc(i, j) : 0.0;
  }
  filtering_combine_core<N, false>(nullptr, b1, U1, e2, nullptr, bo, Uo);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    x.m[i] = bo[i];
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) x.Lc(i, j) = Uo[i][j];
  }
}

// ---------------------------------------------------------------------------------------
// Reference-layout filtering element (with the prior folded into step 0), for the
// element-level entry point psqrt_filter_elements.                     _filtering.py:115-146
// m0/L0 may be null (zero prior).  L0 is a dense N x N factor (not nec. triangular).
// ---------------------------------------------------------------------------------------
template <int N, int NY>
PSQ_HD void filtering_element(const StepPtrs& p, const double* m0, const double* L0,
                              double* A, double* bo, double* U, double* eta, double* Z) {
  double F[N][N], mz[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    mz[i] = m0 ? m0[i] : 0.0;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);
  }
  double m1[N], M1[N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(F[i][k], mz[k], s);
    m1[i] = s;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double u = 0.0;
      if (L0) {
        PSQ_UNROLL
        for (int k = 0; k < N; ++k) u = fma(F[i][k], L0[k * N + j], u);
      }
      M1[i][j] = u;
      M1[i][N + j] = p.template fQ<N>(i, j);
    }
  }
  house_rows<N, 2 * N, N>(M1);
  double M2[NY + N][N + NY];
  build_update<N, NY>(p, [&](int i, int j) { return M1[i][j]; }, M2);
  double inv[NY];
  psi11_inv_diag<N, NY>(M2, inv);
  // HF = H F ; V = Psi11^{-1} HF ; r1 = Psi11^{-1}(y - H m1 - c) ; r2 = Psi11^{-1}(y - H bq - c)
  double V[NY][N], r1[NY], r2[NY];
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double h[N];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) h[k] = p.template fH<N>(a, k);
    double ra = p.fy(a) - p.fc(a), rb = ra;
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) {
      ra = fma(-h[k], m1[k], ra);
      rb = fma(-h[k], p.fb(k), rb);
    }
    PSQ_UNROLL
    for (int q = 0; q < a; ++q) {
      ra = fma(-M2[a][q], r1[q], ra);
      rb = fma(-M2[a][q], r2[q], rb);
    }
    r1[a] = ra * inv[a];
    r2[a] = rb * inv[a];
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) s = fma(h[k], F[k][j], s);
      PSQ_UNROLL
      for (int q = 0; q < a; ++q) s = fma(-M2[a][q], V[q][j], s);
      V[a][j] = s * inv[a];
    }
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double bi = m1[i], e = 0.0;
    PSQ_UNROLL
    for (int a = 0; a < NY; ++a) {
      bi = fma(M2[NY + i][a], r1[a], bi);
      e = fma(V[a][i], r2[a], e);
    }
    bo[i] = bi;
    eta[i] = e;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = F[i][j];
      PSQ_UNROLL
      for (int a = 0; a < NY; ++a) s = fma(-M2[NY + i][a], V[a][j], s);
      A[i * N + j] = s;
      U[i * N + j] = (j <= i) ? M2[NY + i][NY + j] : 0.0;
    }
  }
  // Z = [V^T | 0] if N > NY else tria(V^T)                                _filtering.py:141-144
  if (N > NY) {
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) Z[i * N + j] = (j < NY) ? V[j][i] : 0.0;
  } else {
    double Zt[N][NY];
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int a = 0; a < NY; ++a) Zt[i][a] = V[a][i];
    house_rows<N, NY, N>(Zt);
    PSQ_UNROLL
    for (int i = 0; i < N; ++i)
      PSQ_UNROLL
      for (int j = 0; j < N; ++j) Z[i * N + j] = (j <= i) ? Zt[i][j] : 0.0;
  }
}

// Log-likelihood term from the filtered state at k-1 (dense factor allowed)  _filtering.py:149-154
template <int N, int NY>
PSQ_HD double loglik_term(const StepPtrs& p, const double* m, const double* L) {
  double F[N][N], mp[N], M1[N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(F[i][k], m[k], s);
    mp[i] = s;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double u = 0.0;
      PSQ_UNROLL
      for (int k = 0; k < N; ++k) u = fma(F[i][k], L[k * N + j], u);
      M1[i][j] = u;
      M1[i][N + j] = p.template fQ<N>(i, j);
    }
  }
  house_rows<N, 2 * N, N>(M1);
  double S[NY][N + NY];
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) s = fma(p.template fH<N>(a, k), M1[k][j], s);
      S[a][j] = s;
    }
    PSQ_UNROLL
    for (int q = 0; q < NY; ++q) S[a][N + q] = p.template fR<NY>(a, q);
  }
  house_rows<NY, N + NY, NY>(S);
  double rr[NY], quad = 0.0, logdet = 0.0;
  PSQ_UNROLL
  for (int a = 0; a < NY; ++a) {
    double r = p.fy(a) - p.fc(a);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) r = fma(-p.template fH<N>(a, k), mp[k], r);
    PSQ_UNROLL
    for (int q = 0; q < a; ++q) r = fma(-S[a][q], rr[q], r);
    rr[a] = r * rcp_nr(S[a][a]);
    quad = fma(rr[a], rr[a], quad);
    logdet += log(fabs(S[a][a]));
  }
  return -0.5 * quad - logdet - NY * kHalfLog2Pi;
}

// ---------------------------------------------------------------------------------------
// Chunk-level smoothing element.  For a run of steps k0 .. k1-1 whose filtering summary is
// e = (A, b, U, eta, Z) (the ordered combine of its elements, _operators.py:58-77) and the filtered
// state x = (m, L) at index k0, the ordered combine of the run's smoothing elements
// (_operators.py:118-125) is the conditional  p(x_k0 | x_k1, y_{<k1}) = N(E x_k1 + g, D D^T).
// It follows from the summary in two steps instead of k1 - k0 combines:
//   (1) condition x_k0 on the run's observations, which enter through (eta, Z Z^T) exactly as in the
//       filtering combine:  Xi = tria([[L^T Z, I], [Z, 0]]),  L' = L Xi11^{-T},
//       m' = (I - L' Xi21^T)(m + L L^T eta);
//   (2) one smoothing-element construction (_smoothing.py:72-85) for the "big step"
//       x_k1 = A x_k0 + b + N(0, U U^T) from (m', L').
// ---------------------------------------------------------------------------------------
template <int N>
struct ElemAsStep {  // presents a filtering summary as the linear model of one (big) step
  const FElem<N>& e;
  template <int N_> PSQ_HD double fF(int i, int j) const { return e.A(i, j); }
  template <int N_> PSQ_HD double fQ(int i, int j) const { return e.U(i, j); }  // only j <= i is read
  PSQ_HD double fb(int i) const { return e.b(i); }
};

template <int N>
PSQ_HD void chunk_smoothing_total(const Gauss<N>& x, const FElem<N>& e, SElem<N>& out) {
  double Xi[2 * N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
      PSQ_UNROLL
      for (int k = (i > j ? i : j); k < N; ++k) s = fma(x.Lc(k, i), e.Z(k, j), s);
      Xi[i][j] = s;
      Xi[i][N + j] = (i == j) ? 1.0 : 0.0;
      Xi[N + i][j] = (j <= i) ? e.Z(i, j) : 0.0;
      Xi[N + i][N + j] = 0.0;
    }
  house_rows<2 * N, 2 * N, N, N>(Xi);
  double inv[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) inv[i] = rcp_nr(Xi[i][i]);
  // T1 = Xi11^{-1} L^T, so that L' = T1^T
  double T1[N][N];
  PSQ_UNROLL
  for (int j = 0; j < N; ++j)
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      double s = (j >= i) ? x.Lc(j, i) : 0.0;
      PSQ_UNROLL
      for (int k = 0; k < i; ++k) s = fma(-Xi[i][k], T1[k][j], s);
      T1[i][j] = s * inv[i];
    }
  // m' = t - T1^T Xi21^T t,  t = m + L L^T eta
  double t[N], u[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    PSQ_UNROLL
    for (int k = i; k < N; ++k) s = fma(x.Lc(k, i), e.eta(k), s);
    u[i] = s;
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = x.m[i];
    PSQ_UNROLL
    for (int k = 0; k <= i; ++k) s = fma(x.Lc(i, k), u[k], s);
    t[i] = s;
  }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = 0.0;
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(Xi[N + k][i], t[k], s);
    u[i] = s;
  }
  Gauss<N> xp;
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = t[i];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(-T1[k][i], u[k], s);
    xp.m[i] = s;
  }
  // lower-triangular factor of L' L'^T (the element construction reads a packed triangle)
  double Lp[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) Lp[i][j] = T1[j][i];
  house_rows<N, N, N - 1>(Lp);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) xp.Lc(i, j) = Lp[i][j];
  smoothing_element<N>(xp, ElemAsStep<N>{e}, out);
}

// ---------------------------------------------------------------------------------------
// One RTS step of sweep 3: smoothed state at k from the smoothed state xs at k+1, the FILTERED
// state xf at k and the model of step k.  Fuses the smoothing-element construction
// (_smoothing.py:72-85) with the combine that applies it (_operators.py:118-125):
//   tria([[F L, Q], [L, 0]]) = [[Phi11, 0], [Phi21, Phi22]],  E = Phi21 Phi11^{-1},
//   m_s <- m + E (m_s - F m - b),   L_s <- tria([E L_s | Phi22]).
// Phi22 enters the last triangularisation as left by the first one (dense N x N): triangularising
// it on its own first, as the element-level seam does, would not change L_s L_s^T.
// ---------------------------------------------------------------------------------------
template <int N, class P>
PSQ_HD void rts_step(Gauss<N>& xs, const Gauss<N>& xf, const P& p) {
  double F[N][N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) F[i][j] = p.template fF<N>(i, j);
  double dm[N];  // m_s(k+1) - (F m + b)
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = xs.m[i] - p.fb(i);
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) s = fma(-F[i][k], xf.m[k], s);
    dm[i] = s;
  }
  double M1[2 * N][2 * N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double u = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) u = fma(F[i][k], xf.Lc(k, j), u);
      M1[i][j] = u;
      M1[i][N + j] = (j <= i) ? p.template fQ<N>(i, j) : 0.0;  // cholQ is lower triangular (psqrt.h)
    }
  PSQ_UNROLL
  for (int i = 0; i < N; ++i)
    PSQ_UNROLL
    for (int j = 0; j < 2 * N; ++j) M1[N + i][j] = (j <= i) ? xf.Lc(i, j) : 0.0;
  house_rows<2 * N, 2 * N, N, N>(M1);
  double inv[N];
  PSQ_UNROLL
  for (int j = 0; j < N; ++j) inv[j] = rcp_nr(M1[j][j]);
  double W[N][2 * N], mn[N];
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    double E[N];
    PSQ_UNROLL
    for (int j = N - 1; j >= 0; --j) {
      double s = M1[N + i][j];
      PSQ_UNROLL
      for (int k = j + 1; k < N; ++k) s = fma(-E[k], M1[k][j], s);
      E[j] = s * inv[j];
    }
    double mi = xf.m[i];
    PSQ_UNROLL
    for (int k = 0; k < N; ++k) mi = fma(E[k], dm[k], mi);
    mn[i] = mi;
    PSQ_UNROLL
    for (int j = 0; j < N; ++j) {
      double d = 0.0;
      PSQ_UNROLL
      for (int k = j; k < N; ++k) d = fma(E[k], xs.Lc(k, j), d);
      W[i][j] = d;
      W[i][N + j] = M1[N + i][N + j];
    }
  }
  house_rows<N, 2 * N, N>(W);
  PSQ_UNROLL
  for (int i = 0; i < N; ++i) {
    xs.m[i] = mn[i];
    PSQ_UNROLL
    for (int j = 0; j <= i; ++j) xs.Lc(i, j) = W[i][j];
  }
}

// Rank-one Cholesky update, literal column sweep of _utils.py:39-81 (incl. the
// non-finite -> 0 guard, line 80).  L is dense row-major N x N (lower part used).
template <int N>
PSQ_HD void chol_update(double (&L)[N][N], double (&w)[N], double mult) {
  double b = 1.0;
  PSQ_UNROLL
  for (int j = 0; j < N; ++j) {
    const double d = L[j][j];
    const double wj = w[j];
    const double nd = sqrt(fma(d, d, mult / b * (wj * wj)));
    const double gamma = fma(d * d, b, mult * (wj * wj));
    const double wd = wj / d;
    const double cg = mult * wj / gamma;
    PSQ_UNROLL
    for (int i = 0; i < N; ++i) {
      const double col = L[i][j];  // entries above the diagonal are 0 and stay irrelevant
      w[i] = fma(-wd, col, w[i]);
      double nc = nd * (col / d + cg * w[i]);
      if (i == j) nc = nd;
      if (i < j) nc = 0.0;
      L[i][j] = isfinite(nc) ? nc : 0.0;
    }
    b = fma(mult, wd * wd, b);
  }
}

}  // namespace psq
