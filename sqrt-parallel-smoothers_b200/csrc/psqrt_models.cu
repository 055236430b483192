// psqrt_models.cu -- linearization of the built-in models as device code (one thread per time step):
//   model 1  coordinated-turn transition          tests/bearings/bearings_utils.py:7-46   (nx 5 -> 5)
//   model 2  two-sensor bearings observation      tests/bearings/bearings_utils.py:49-69  (nx 5 -> 2)
//   model 3  Ricker map, conditional moments      notebooks/population_model.py:23-34,51-62   (1 -> 1)
//   model 4  Poisson observation, cond. moments   notebooks/population_model.py:84-129        (1 -> 1)
// with
//   lin 0    extended / first-order Taylor        parsmooth/linearization/_extended.py:51-70
//   lin 1    statistical linear regression from a sigma-point set (cubature, Gauss-Hermite, ...)
//            parsmooth/linearization/_sigma_points.py:25-100, incl. the Cholesky downdates of
//            parsmooth/_utils.py:13-19,39-81.
// Output per step: (F [d,n], chol [d,d], b [d]) -- the tuple linearization_method returns
// (parallel/_filtering.py:117-119), which the sweeps of psqrt_kernels.cuh consume.
//
// The SLR streams over the P sigma points twice (mean, then moments) so nothing of size P is stored:
// Psi accumulates in registers and the residual factor is built by Householder appends of the
// weighted deviations, 4 columns at a time.
#include "../../include/psqrt.h"

#include <cuda_runtime.h>

#include "psqrt_math.cuh"

namespace psq {
namespace {

// ---------------------------------------------------------------------------------------------------
// model functors: NIN, NOUT, f(x) and (for extended) the Jacobian df/dx
// ---------------------------------------------------------------------------------------------------
struct CTTransition {  // params: dt
  static constexpr int NIN = 5, NOUT = 5;
  static constexpr bool CONDITIONAL = false;
  double dt;
  __device__ void parts(const double* x, double& cw, double& sw, double& a, double& bq, double& da, double& db) const {
    const double w = x[4];
    const bool small = fabs(w) < 1e-6;  // lax.cond branch: sin(wt)/w -> dt, (cos(wt)-1)/w -> 0 as constants
    sincos(w * dt, &sw, &cw);
    if (small) {
      a = dt; bq = 0.0; da = 0.0; db = 0.0;
    } else {
      const double iw = 1.0 / w;
      a = sw * iw;
      bq = (cw - 1.0) * iw;
      da = (dt * cw * w - sw) * iw * iw;
      db = (-dt * sw * w - (cw - 1.0)) * iw * iw;
    }
  }
  __device__ void f(const double* x, double* y) const {
    double cw, sw, a, bq, da, db;
    parts(x, cw, sw, a, bq, da, db);
    y[0] = x[0] + a * x[2] - bq * x[3];
    y[1] = x[1] + bq * x[2] + a * x[3];
    y[2] = cw * x[2] + sw * x[3];
    y[3] = -sw * x[2] + cw * x[3];
    y[4] = x[4];
  }
  __device__ void jac(const double* x, double* y, double (&J)[5][5]) const {
    double cw, sw, a, bq, da, db;
    parts(x, cw, sw, a, bq, da, db);
    f(x, y);
    const double vx = x[2], vy = x[3], dcw = -dt * sw, dsw = dt * cw;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) J[i][j] = 0.0;
    J[0][0] = 1.0; J[0][2] = a;   J[0][3] = -bq; J[0][4] = da * vx - db * vy;
    J[1][1] = 1.0; J[1][2] = bq;  J[1][3] = a;   J[1][4] = db * vx + da * vy;
    J[2][2] = cw;  J[2][3] = sw;  J[2][4] = dcw * vx + dsw * vy;
    J[3][2] = -sw; J[3][3] = cw;  J[3][4] = -dsw * vx + dcw * vy;
    J[4][4] = 1.0;
  }
};

struct BearingsObservation {  // params: s1x, s1y, s2x, s2y
  static constexpr int NIN = 5, NOUT = 2;
  static constexpr bool CONDITIONAL = false;
  double s1x, s1y, s2x, s2y;
  __device__ void f(const double* x, double* y) const {
    y[0] = atan2(x[1] - s1y, x[0] - s1x);
    y[1] = atan2(x[1] - s2y, x[0] - s2x);
  }
  __device__ void jac(const double* x, double* y, double (&J)[2][5]) const {
    f(x, y);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) J[i][j] = 0.0;
    {
      const double dx = x[0] - s1x, dy = x[1] - s1y, r2 = dx * dx + dy * dy;
      J[0][0] = -dy / r2; J[0][1] = dx / r2;
    }
    {
      const double dx = x[0] - s2x, dy = x[1] - s2y, r2 = dx * dx + dy * dy;
      J[1][0] = -dy / r2; J[1][1] = dx / r2;
    }
  }
};

struct RickerTransition {  // params: sqrt(Q); E[x'|x] = log 44.7 + x - exp x, chol = sqrt Q
  static constexpr int NIN = 1, NOUT = 1;
  static constexpr bool CONDITIONAL = true;
  double sqrtQ;
  __device__ void f(const double* x, double* y) const { y[0] = 3.7999735016195233 + x[0] - exp(x[0]); }
  __device__ void jac(const double* x, double* y, double (&J)[1][1]) const { f(x, y); J[0][0] = 1.0 - exp(x[0]); }
  __device__ void chol(const double*, double (&C)[1][1]) const { C[0][0] = sqrtQ; }
};

struct PoissonObservation {  // params: lam; E[y|x] = lam exp x, chol = sqrt(lam exp x)
  static constexpr int NIN = 1, NOUT = 1;
  static constexpr bool CONDITIONAL = true;
  double lam;
  __device__ void f(const double* x, double* y) const { y[0] = lam * exp(x[0]); }
  __device__ void jac(const double* x, double* y, double (&J)[1][1]) const { f(x, y); J[0][0] = y[0]; }
  __device__ void chol(const double* x, double (&C)[1][1]) const { C[0][0] = sqrt(lam * exp(x[0])); }
};

template <class M, bool C = M::CONDITIONAL>
struct CondChol {
  template <int D>
  static __device__ void eval(const M&, const double*, double (&)[D][D]) {}
};
template <class M>
struct CondChol<M, true> {
  template <int D>
  static __device__ void eval(const M& m, const double* x, double (&Cm)[D][D]) { m.chol(x, Cm); }
};

// ---------------------------------------------------------------------------------------------------
// extended                                                                  _extended.py:51-70
// ---------------------------------------------------------------------------------------------------
template <class M>
__global__ void k_lin_extended(M model, const double* __restrict__ nom_m, long long count,
                               const double* __restrict__ m_q, double* __restrict__ F, double* __restrict__ chol,
                               double* __restrict__ b) {
  constexpr int N = M::NIN, D = M::NOUT;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double x[N], y[D], J[D][N];
#pragma unroll
  for (int k = 0; k < N; ++k) x[k] = nom_m[i * N + k];
  model.jac(x, y, J);
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double s = y[r];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      s = fma(-J[r][k], x[k], s);
      F[(i * D + r) * N + k] = J[r][k];
    }
    b[i * D + r] = s + (m_q ? m_q[r] : 0.0);  // res - F x + m_q
  }
  if (M::CONDITIONAL) {
    double Cm[D][D];
    CondChol<M>::eval(model, x, Cm);
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q < D; ++q) chol[(i * D + r) * D + q] = Cm[r][q];
  }
}

// ---------------------------------------------------------------------------------------------------
// statistical linear regression                                         _sigma_points.py:25-100
// ---------------------------------------------------------------------------------------------------
template <class M>
__global__ void k_lin_slr(M model, const double* __restrict__ xi, const double* __restrict__ wm,
                          const double* __restrict__ wc, int P, const double* __restrict__ nom_m,
                          const double* __restrict__ nom_L, long long count, const double* __restrict__ m_q,
                          const double* __restrict__ chol_q, double* __restrict__ F, double* __restrict__ chol,
                          double* __restrict__ b) {
  constexpr int N = M::NIN, D = M::NOUT;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double m[N], L[N][N];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    m[r] = nom_m[i * N + r];
#pragma unroll
    for (int q = 0; q < N; ++q) L[r][q] = nom_L[(i * N + r) * N + q];
  }
  auto point = [&](int p, double* x, double* dx) {  // x = m + L xi_p                   _cubature.py:56, _gh.py:68
#pragma unroll
    for (int r = 0; r < N; ++r) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < N; ++q) s = fma(L[r][q], __ldg(xi + (long long)p * N + q), s);
      dx[r] = s;
      x[r] = m[r] + s;
    }
  };
  // pass 1: m_f = sum_p wm_p f(x_p)
  double mf[D];
#pragma unroll
  for (int r = 0; r < D; ++r) mf[r] = 0.0;
#pragma unroll 1
  for (int p = 0; p < P; ++p) {
    double x[N], dx[N], y[D];
    point(p, x, dx);
    model.f(x, y);
    const double w = __ldg(wm + p);
#pragma unroll
    for (int r = 0; r < D; ++r) mf[r] = fma(w, y[r], mf[r]);
  }
  // pass 2: Psi = sum wc (x - m)(f - m_f)^T ; residual factor by appends of sqrt(wc) (f - m_f)
  double Psi[N][D], Lr[D][D];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int q = 0; q < D; ++q) Psi[r][q] = 0.0;
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int q = 0; q < D; ++q) Lr[r][q] = 0.0;
  constexpr int CW = M::CONDITIONAL ? 4 : 4;
  double W[D][CW];
  int fill = 0;
  auto flush = [&]() {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int q = 0; q < CW; ++q)
        if (q >= fill) W[r][q] = 0.0;
    tria_append<D, CW>([&](int r, int q) -> double& { return Lr[r][q]; }, W);
    fill = 0;
  };
  auto push_col = [&](const double* col) {
#pragma unroll
    for (int q = 0; q < CW; ++q)
      if (q == fill) {
#pragma unroll
        for (int r = 0; r < D; ++r) W[r][q] = col[r];
      }
    if (++fill == CW) flush();
  };
#pragma unroll 1
  for (int p = 0; p < P; ++p) {
    double x[N], dx[N], y[D], col[D];
    point(p, x, dx);
    model.f(x, y);
    const double w = __ldg(wc + p), sw = sqrt(w);
#pragma unroll
    for (int q = 0; q < D; ++q) {
      const double df = y[q] - mf[q];
      col[q] = sw * df;
#pragma unroll
      for (int r = 0; r < N; ++r) Psi[r][q] = fma(w * dx[r], df, Psi[r][q]);
    }
    push_col(col);
    if (M::CONDITIONAL) {  // + sqrt(wc) chol(x_p) blocks                         _sigma_points.py:40-47
      double Cm[D][D];
      CondChol<M>::eval(model, x, Cm);
#pragma unroll
      for (int c2 = 0; c2 < D; ++c2) {
        double cc[D];
#pragma unroll
        for (int r = 0; r < D; ++r) cc[r] = sw * Cm[r][c2];
        push_col(cc);
      }
    }
  }
  if (!M::CONDITIONAL) {  // + chol_q                                             _sigma_points.py:77
#pragma unroll
    for (int c2 = 0; c2 < D; ++c2) {
      double cc[D];
#pragma unroll
      for (int r = 0; r < D; ++r) cc[r] = chol_q[r * D + c2];
      push_col(cc);
    }
  }
  if (fill) flush();
  // F^T = cho_solve((L, lower), Psi): lower triangle of L only                   _sigma_points.py:98
  double X[N][D];
#pragma unroll
  for (int q = 0; q < D; ++q) {
#pragma unroll
    for (int r = 0; r < N; ++r) {  // forward: L z = Psi
      double s = Psi[r][q];
#pragma unroll
      for (int k = 0; k < r; ++k) s = fma(-L[r][k], X[k][q], s);
      X[r][q] = s / L[r][r];
    }
#pragma unroll
    for (int r = N - 1; r >= 0; --r) {  // backward: L^T x = z
      double s = X[r][q];
#pragma unroll
      for (int k = r + 1; k < N; ++k) s = fma(-L[k][r], X[k][q], s);
      X[r][q] = s / L[r][r];
    }
  }
  // downdates with the columns of F L  (rows of (F L)^T)                          _sigma_points.py:78
#pragma unroll 1
  for (int c2 = 0; c2 < N; ++c2) {
    double v[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(X[k][r], L[k][c2], s);  // F[r][k] = X[k][r]
      v[r] = s;
    }
    // chol_update wants compile-time indices: select the column with a switch-free unrolled copy
    chol_update<D>(Lr, v, -1.0);
  }
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double s = mf[r];
#pragma unroll
    for (int k = 0; k < N; ++k) {
      s = fma(-X[k][r], m[k], s);
      F[(i * D + r) * N + k] = X[k][r];
    }
    b[i * D + r] = s + ((!M::CONDITIONAL && m_q) ? m_q[r] : 0.0);
#pragma unroll
    for (int q = 0; q < D; ++q) chol[(i * D + r) * D + q] = (q <= r) ? Lr[r][q] : 0.0;
  }
}

template <class M>
int launch(const M& model, int lin_id, const double* xi, const double* wm, const double* wc, int P,
           const double* nom_m, const double* nom_L, long long count, const double* m_q, const double* chol_q,
           double* F, double* chol, double* b, cudaStream_t st) {
  const unsigned blocks = (unsigned)((count + 63) / 64);
  if (lin_id == 0) {
    if (M::CONDITIONAL && !chol) return PSQRT_EINVAL;
    k_lin_extended<M><<<blocks, 64, 0, st>>>(model, nom_m, count, m_q, F, chol, b);
  } else if (lin_id == 1) {
    if (!xi || !wm || !wc || P <= 0 || !nom_L || !chol || (!M::CONDITIONAL && !chol_q)) return PSQRT_EINVAL;
    k_lin_slr<M><<<blocks, 64, 0, st>>>(model, xi, wm, wc, P, nom_m, nom_L, count, m_q, chol_q, F, chol, b);
  } else {
    return PSQRT_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

}  // namespace
}  // namespace psq

extern "C" int psqrt_linearize_builtin(int model_id, const double* model_params, int lin_id, const double* xi,
                                       const double* wm, const double* wc, int n_points, const double* nom_m,
                                       const double* nom_L, int64_t count, const double* m_q, const double* chol_q,
                                       double* F, double* chol, double* b, void* stream) {
  if (!model_params || !nom_m || count <= 0 || !F || !b) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const double* p = model_params;  // HOST array of model parameters
  switch (model_id) {
    case PSQRT_MODEL_CT_TRANSITION:
      return psq::launch(psq::CTTransition{p[0]}, lin_id, xi, wm, wc, n_points, nom_m, nom_L, count, m_q, chol_q, F,
                         chol, b, st);
    case PSQRT_MODEL_BEARINGS_OBSERVATION:
      return psq::launch(psq::BearingsObservation{p[0], p[1], p[2], p[3]}, lin_id, xi, wm, wc, n_points, nom_m, nom_L,
                         count, m_q, chol_q, F, chol, b, st);
    case PSQRT_MODEL_RICKER_TRANSITION:
      return psq::launch(psq::RickerTransition{p[0]}, lin_id, xi, wm, wc, n_points, nom_m, nom_L, count, m_q, chol_q,
                         F, chol, b, st);
    case PSQRT_MODEL_POISSON_OBSERVATION:
      return psq::launch(psq::PoissonObservation{p[0]}, lin_id, xi, wm, wc, n_points, nom_m, nom_L, count, m_q,
                         chol_q, F, chol, b, st);
    default:
      return PSQRT_EUNSUPPORTED;
  }
}
