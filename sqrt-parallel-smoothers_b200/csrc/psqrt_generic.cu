// psqrt_generic.cu -- the generic path: any state dimension nx <= 16 and observation dimension ny <= 16, in fp64 or
// fp32.  It serves what the tuned kernels (psqrt_kernels.cuh, one fully unrolled instantiation per nx in {1..6, 8},
// ny <= 4, fp64) do not cover:
//   * other dimensions ("generic <= 16"): psqrt_filter_smoother / psqrt_smoother fall back to it;
//   * the float32 runs of the robustness experiments (notebooks/robustness_100runs.py:7,41-77 compares the
//     square-root and the covariance form in float32): psqrt_filter_smoother_f32.
// It is a literal CUDA statement of the reference's algorithm, one thread per time step and runtime dimensions:
//   elements      parallel/_filtering.py:100-146, parallel/_smoothing.py:47-57,72-85
//   operators     parallel/_operators.py:43-77 (filtering), 104-125 (smoothing)
//   scan          jax.lax.associative_scan as a Hillis-Steele scan: ceil(log2 T) levels, each one launch over all
//                 steps, ping-pong between two element arrays in HBM (O(T log T) combines -- the price of generality;
//                 the tuned path does O(T) and never materialises elements)
//   ell           parallel/_filtering.py:149-154
//   tria          parsmooth/_utils.py:22-24 (Householder from the right, row by row)
// Every matrix lives in per-thread local memory; nothing here is tuned.
#include "../../include/psqrt.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace psq {
namespace generic {

constexpr int kBlock = 64;

// ---- small dense helpers, runtime shapes, row-major with leading dimension ------------------------------------
// Lower-trapezoidal triangularisation from the right: A [R x C] (ld) <- A Q with the first min(R, C) rows made
// lower triangular (LAPACK dgeqr2 on the transpose; a zero tail gives tau = 0 like dlarfg).
template <class S>
__device__ void g_tria(S* A, int R, int C, int ld) {
  const int nr = R < C ? R : C;
  for (int j = 0; j < nr; ++j) {
    S* rj = A + j * ld;
    S sigma = S(0);
    for (int k = j + 1; k < C; ++k) sigma += rj[k] * rj[k];
    if (sigma == S(0)) continue;
    const S alpha = rj[j];
    const S norm = sqrt(alpha * alpha + sigma);
    const S beta = alpha >= S(0) ? -norm : norm;
    const S v0 = alpha - beta;
    const S scale = S(1) / (norm * (norm + fabs(alpha)));   // 2 / (v^T v), v = (v0, tail)
    for (int i = j + 1; i < R; ++i) {
      S* ri = A + i * ld;
      S d = ri[j] * v0;
      for (int k = j + 1; k < C; ++k) d += ri[k] * rj[k];
      d *= scale;
      ri[j] -= d * v0;
      for (int k = j + 1; k < C; ++k) ri[k] -= d * rj[k];
    }
    rj[j] = beta;
    for (int k = j + 1; k < C; ++k) rj[k] = S(0);
  }
}
// X [n x m] (ldx) <- L^-1 X, L [n x n] lower (ldl)
template <class S>
__device__ void g_solve_lower(const S* L, int ldl, S* X, int ldx, int n, int m) {
  for (int c = 0; c < m; ++c)
    for (int i = 0; i < n; ++i) {
      S s = X[i * ldx + c];
      for (int k = 0; k < i; ++k) s -= L[i * ldl + k] * X[k * ldx + c];
      X[i * ldx + c] = s / L[i * ldl + i];
    }
}
// X [n x m] <- L^-T X
template <class S>
__device__ void g_solve_lower_t(const S* L, int ldl, S* X, int ldx, int n, int m) {
  for (int c = 0; c < m; ++c)
    for (int i = n - 1; i >= 0; --i) {
      S s = X[i * ldx + c];
      for (int k = i + 1; k < n; ++k) s -= L[k * ldl + i] * X[k * ldx + c];
      X[i * ldx + c] = s / L[i * ldl + i];
    }
}

template <class S>
struct Model {   // one step of the linearised model + the strides that lead to it
  const S *F, *Q, *b, *H, *R, *c;
};
template <class S>
struct ModelArgs {
  const S *F, *Q, *b, *H, *R, *c;
  long long tF, tQ, tb, tH, tR, tc;
  long long sF, sQ, sb, sH, sR, sc;
  __device__ Model<S> at(long long seq, long long k) const {
    Model<S> m;
    m.F = F + seq * sF + k * tF;
    m.Q = Q + seq * sQ + k * tQ;
    m.b = b + seq * sb + k * tb;
    m.H = H ? H + seq * sH + k * tH : nullptr;
    m.R = R ? R + seq * sR + k * tR : nullptr;
    m.c = c ? c + seq * sc + k * tc : nullptr;
    return m;
  }
};

// filtering element layout per step: A [n,n] | b [n] | U [n,n] | eta [n] | Z [n,n]; smoothing: g [n] | E [n,n] | D [n,n]
__host__ __device__ inline int fe_size(int n) { return 3 * n * n + 2 * n; }
__host__ __device__ inline int se_size(int n) { return 2 * n * n + n; }

// ---- filtering elements                                                  parallel/_filtering.py:100-146
template <class S, int MN>
__global__ void __launch_bounds__(kBlock)
k_felems(ModelArgs<S> ma, const S* __restrict__ y, const S* __restrict__ m0, const S* __restrict__ L0, int n, int ny,
         long long T, S* __restrict__ elems) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long seq = blockIdx.y;
  if (k >= T) return;
  const Model<S> p = ma.at(seq, k);
  const S* yk = y + (seq * T + k) * ny;
  S* e = elems + (seq * T + k) * fe_size(n);
  S *eA = e, *eb = e + n * n, *eU = eb + n, *eeta = eU + n * n, *eZ = eeta + n;
  S N1[MN][2 * MN], m1[MN], Psi[2 * MN][2 * MN], HF[MN][MN], r[MN];
  // m1 = F m + b, N1 = tria([F L, Q]) with (m, L) = prior at step 0 and zero afterwards (lines 106-108, 121-122)
  for (int i = 0; i < n; ++i) {
    S s = p.b[i];
    for (int j = 0; j < n; ++j) {
      if (k == 0) s += p.F[i * n + j] * m0[seq * n + j];
      S v = S(0);
      if (k == 0)
        for (int q = 0; q < n; ++q) v += p.F[i * n + q] * L0[(seq * n + q) * n + j];
      N1[i][j] = v;
      N1[i][n + j] = p.Q[i * n + j];
    }
    m1[i] = s;
  }
  g_tria<S>(&N1[0][0], n, 2 * n, 2 * MN);
  // Psi = tria([[H N1, R], [N1, 0]])                                         lines 126-131
  const int W = n + ny;
  for (int a = 0; a < ny; ++a) {
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += p.H[a * n + q] * N1[q][j];
      Psi[a][j] = v;
    }
    for (int j = 0; j < ny; ++j) Psi[a][n + j] = p.R[a * ny + j];
  }
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) Psi[ny + i][j] = (j <= i) ? N1[i][j] : S(0);
    for (int j = 0; j < ny; ++j) Psi[ny + i][n + j] = S(0);
  }
  g_tria<S>(&Psi[0][0], n + ny, W, 2 * MN);
  // Psi11 = Psi[:ny,:ny], Psi21 = Psi[ny:,:ny], U = Psi[ny:,ny:]
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) eU[i * n + j] = (j <= i) ? Psi[ny + i][ny + j] : S(0);
  // HF = H F; Z^T = Psi11^-1 H F (ny x n)
  for (int a = 0; a < ny; ++a)
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += p.H[a * n + q] * p.F[q * n + j];
      HF[a][j] = v;
    }
  // K = Psi21 Psi11^-1  ->  A = F - K H F,  b = m1 + K (y - H m1 - c)             lines 133-138
  // K (H F) = Psi21 (Psi11^-1 H F) = Psi21 Zt
  S Zt[MN][MN];
  for (int a = 0; a < ny; ++a)
    for (int j = 0; j < n; ++j) Zt[a][j] = HF[a][j];
  g_solve_lower<S>(&Psi[0][0], 2 * MN, &Zt[0][0], MN, ny, n);
  for (int a = 0; a < ny; ++a) {
    S s = yk[a] - p.c[a];
    for (int q = 0; q < n; ++q) s -= p.H[a * n + q] * m1[q];
    r[a] = s;
  }
  g_solve_lower<S>(&Psi[0][0], 2 * MN, r, 1, ny, 1);   // Psi11^-1 (y - H m1 - c)
  for (int i = 0; i < n; ++i) {
    S s = m1[i];
    for (int a = 0; a < ny; ++a) s += Psi[ny + i][a] * r[a];
    eb[i] = s;
    for (int j = 0; j < n; ++j) {
      S v = p.F[i * n + j];
      for (int a = 0; a < ny; ++a) v -= Psi[ny + i][a] * Zt[a][j];
      eA[i * n + j] = v;
    }
  }
  // eta = Z Psi11^-1 (y - H b - c), Z = Zt^T                                     lines 140-144
  for (int a = 0; a < ny; ++a) {
    S s = yk[a] - p.c[a];
    for (int q = 0; q < n; ++q) s -= p.H[a * n + q] * p.b[q];
    r[a] = s;
  }
  g_solve_lower<S>(&Psi[0][0], 2 * MN, r, 1, ny, 1);
  for (int i = 0; i < n; ++i) {
    S s = S(0);
    for (int a = 0; a < ny; ++a) s += Zt[a][i] * r[a];
    eeta[i] = s;
  }
  if (n >= ny) {   // Z = [Zt^T | 0]   (for n == ny the reference triangularises a square matrix: same Z Z^T)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) eZ[i * n + j] = (j < ny) ? Zt[j][i] : S(0);
  } else {         // Z = tria(Zt^T)  (n x ny -> n x n)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < ny; ++j) Psi[i][j] = Zt[j][i];
    g_tria<S>(&Psi[0][0], n, ny, 2 * MN);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) eZ[i * n + j] = (j <= i) ? Psi[i][j] : S(0);
  }
}

// ---- sqrt filtering operator, e1 earlier                                  parallel/_operators.py:43-77
template <class S, int MN>
__device__ void f_combine(const S* e1, const S* e2, S* out, int n) {
  const S *A1 = e1, *b1 = e1 + n * n, *U1 = b1 + n, *eta1 = U1 + n * n, *Z1 = eta1 + n;
  const S *A2 = e2, *b2 = e2 + n * n, *U2 = b2 + n, *eta2 = U2 + n * n, *Z2 = eta2 + n;
  S *oA = out, *ob = out + n * n, *oU = ob + n, *oeta = oU + n * n, *oZ = oeta + n;
  S Xi[2 * MN][2 * MN];
  // Xi = [[U1^T Z2, I], [Z2, 0]]
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += U1[q * n + i] * Z2[q * n + j];
      Xi[i][j] = v;
      Xi[i][n + j] = (i == j) ? S(1) : S(0);
      Xi[n + i][j] = Z2[i * n + j];
      Xi[n + i][n + j] = S(0);
    }
  g_tria<S>(&Xi[0][0], 2 * n, 2 * n, 2 * MN);
  // T1 = Xi11^-1 U1^T (n x n);  S1 = A2 T1^T;  W = S1 Xi21^T
  S T1[MN][MN], S1[MN][2 * MN], tmp[MN][MN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) T1[i][j] = U1[j * n + i];
  g_solve_lower<S>(&Xi[0][0], 2 * MN, &T1[0][0], MN, n, n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += A2[i * n + q] * T1[j][q];
      S1[i][j] = v;                       // (Xi11^-1 U1^T A2^T)^T
    }
  // G = A2 - S1 Xi21^T  (so that A = G A1, b = G (b1 + U1 U1^T eta2) + b2)        lines 70-71
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S v = A2[i * n + j];
      for (int q = 0; q < n; ++q) v -= S1[i][q] * Xi[n + j][q];
      tmp[i][j] = v;
    }
  S tv[MN], sv[MN];
  for (int i = 0; i < n; ++i) {
    S u = S(0);
    for (int q = 0; q < n; ++q) u += U1[q * n + i] * eta2[q];   // U1^T eta2
    tv[i] = u;
  }
  for (int i = 0; i < n; ++i) {
    S s = b1[i];
    for (int q = 0; q < n; ++q) s += U1[i * n + q] * tv[q];
    sv[i] = s;                                                   // b1 + U1 U1^T eta2
  }
  for (int i = 0; i < n; ++i) {
    S s = b2[i];
    for (int j = 0; j < n; ++j) {
      s += tmp[i][j] * sv[j];
      S v = S(0);
      for (int q = 0; q < n; ++q) v += tmp[i][q] * A1[q * n + j];
      oA[i * n + j] = v;
    }
    ob[i] = s;
  }
  // eta = A1^T (I - Xi21 Xi11^-1 U1^T)(eta2 - Z2 Z2^T b1) + eta1                    lines 73-74
  for (int i = 0; i < n; ++i) {
    S u = S(0);
    for (int q = 0; q < n; ++q) u += Z2[q * n + i] * b1[q];     // Z2^T b1
    tv[i] = u;
  }
  for (int i = 0; i < n; ++i) {
    S s = eta2[i];
    for (int q = 0; q < n; ++q) s -= Z2[i * n + q] * tv[q];
    sv[i] = s;
  }
  for (int i = 0; i < n; ++i) {     // tv = T1 sv  (Xi11^-1 U1^T sv)
    S u = S(0);
    for (int q = 0; q < n; ++q) u += T1[i][q] * sv[q];
    tv[i] = u;
  }
  for (int i = 0; i < n; ++i) {     // sv -= Xi21 tv
    S s = sv[i];
    for (int q = 0; q < n; ++q) s -= Xi[n + i][q] * tv[q];
    sv[i] = s;
  }
  for (int i = 0; i < n; ++i) {
    S s = eta1[i];
    for (int q = 0; q < n; ++q) s += A1[q * n + i] * sv[q];
    oeta[i] = s;
  }
  // U = tria([S1 | U2])                                                             line 72
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) S1[i][n + j] = U2[i * n + j];
  g_tria<S>(&S1[0][0], n, 2 * n, 2 * MN);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) oU[i * n + j] = (j <= i) ? S1[i][j] : S(0);
  // Z = tria([A1^T Xi22 | Z1])                                                      line 75
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += A1[q * n + i] * ((j <= q) ? Xi[n + q][n + j] : S(0));
      S1[i][j] = v;
      S1[i][n + j] = Z1[i * n + j];
    }
  g_tria<S>(&S1[0][0], n, 2 * n, 2 * MN);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) oZ[i * n + j] = (j <= i) ? S1[i][j] : S(0);
}

// ---- sqrt smoothing operator, e1 = later (accumulated) side                parallel/_operators.py:104-125
template <class S, int MN>
__device__ void s_combine(const S* e1, const S* e2, S* out, int n) {
  const S *g1 = e1, *E1 = e1 + n, *D1 = E1 + n * n;
  const S *g2 = e2, *E2 = e2 + n, *D2 = E2 + n * n;
  S *og = out, *oE = out + n, *oD = oE + n * n;
  S M[MN][2 * MN];
  for (int i = 0; i < n; ++i) {
    S s = g2[i];
    for (int j = 0; j < n; ++j) {
      s += E2[i * n + j] * g1[j];
      S v = S(0), d = S(0);
      for (int q = 0; q < n; ++q) {
        v += E2[i * n + q] * E1[q * n + j];
        d += E2[i * n + q] * D1[q * n + j];
      }
      oE[i * n + j] = v;
      M[i][j] = d;
      M[i][n + j] = D2[i * n + j];
    }
    og[i] = s;
  }
  g_tria<S>(&M[0][0], n, 2 * n, 2 * MN);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) oD[i * n + j] = (j <= i) ? M[i][j] : S(0);
}

// one Hillis-Steele level over the T elements of every sequence: out[i] = in[i - d] (x) in[i]  (i >= d), else in[i].
// REV: scan position i counts from the end of the sequence (the smoother's reverse scan).
template <class S, int MN, bool SMOOTH>
__global__ void __launch_bounds__(kBlock)
k_level(const S* __restrict__ in, S* __restrict__ out, int n, long long T, long long d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long seq = blockIdx.y;
  if (i >= T) return;
  const int ne = SMOOTH ? se_size(n) : fe_size(n);
  const long long at = SMOOTH ? (T - 1 - i) : i;
  const S* e2 = in + (seq * T + at) * ne;
  S* o = out + (seq * T + at) * ne;
  if (i < d) {
    for (int f = 0; f < ne; ++f) o[f] = e2[f];
    return;
  }
  const long long prev = SMOOTH ? (T - 1 - (i - d)) : (i - d);
  const S* e1 = in + (seq * T + prev) * ne;
  if (SMOOTH) s_combine<S, MN>(e1, e2, o, n);
  else f_combine<S, MN>(e1, e2, o, n);
}

// filtered trajectory from the scanned elements: index 0 = prior, k + 1 = (b, U) of element k      lines 34-46
template <class S>
__global__ void k_filtered_out(const S* __restrict__ elems, const S* __restrict__ m0, const S* __restrict__ L0, int n,
                               long long T, S* __restrict__ fm, S* __restrict__ fL) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. T
  const long long seq = blockIdx.y;
  if (k > T) return;
  S* om = fm + (seq * (T + 1) + k) * n;
  S* oL = fL + (seq * (T + 1) + k) * n * n;
  if (k == 0) {
    for (int i = 0; i < n; ++i) om[i] = m0[seq * n + i];
    for (int i = 0; i < n * n; ++i) oL[i] = L0[seq * n * n + i];
    return;
  }
  const S* e = elems + (seq * T + k - 1) * fe_size(n);
  for (int i = 0; i < n; ++i) om[i] = e[n * n + i];
  for (int i = 0; i < n * n; ++i) oL[i] = e[n * n + n + i];
}

// per-step log-likelihood terms                                              parallel/_filtering.py:149-154
template <class S, int MN>
__global__ void __launch_bounds__(kBlock)
k_ell_terms(ModelArgs<S> ma, const S* __restrict__ y, const S* __restrict__ fm, const S* __restrict__ fL, int n,
            int ny, long long T, double* __restrict__ terms) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long seq = blockIdx.y;
  if (k >= T) return;
  const Model<S> p = ma.at(seq, k);
  const S* m = fm + (seq * (T + 1) + k) * n;
  const S* L = fL + (seq * (T + 1) + k) * n * n;
  S N1[MN][2 * MN], O[MN][2 * MN], r[MN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += p.F[i * n + q] * L[q * n + j];
      N1[i][j] = v;
      N1[i][n + j] = p.Q[i * n + j];
    }
  g_tria<S>(&N1[0][0], n, 2 * n, 2 * MN);
  for (int a = 0; a < ny; ++a) {
    S s = y[(seq * T + k) * ny + a] - p.c[a];
    for (int i = 0; i < n; ++i) {
      S pm = p.b[i];
      for (int q = 0; q < n; ++q) pm += p.F[i * n + q] * m[q];
      s -= p.H[a * n + i] * pm;
    }
    r[a] = s;
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = j; q < n; ++q) v += p.H[a * n + q] * N1[q][j];
      O[a][j] = v;
    }
    for (int j = 0; j < ny; ++j) O[a][n + j] = p.R[a * ny + j];
  }
  g_tria<S>(&O[0][0], ny, n + ny, 2 * MN);
  g_solve_lower<S>(&O[0][0], 2 * MN, r, 1, ny, 1);
  double q2 = 0.0, ld = 0.0;
  for (int a = 0; a < ny; ++a) {
    q2 += (double)r[a] * (double)r[a];
    ld += log(fabs((double)O[a][a]));
  }
  terms[seq * T + k] = -0.5 * q2 - ld - 0.5 * ny * 1.8378770664093453;   // log(2 pi)
}

__global__ void __launch_bounds__(256) k_sum_terms(const double* __restrict__ terms, long long T, double* __restrict__ out) {
  __shared__ double sh[256];
  const long long seq = blockIdx.x;
  double s = 0.0;
  for (long long i = threadIdx.x; i < T; i += 256) s += terms[seq * T + i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[seq] = sh[0];
}

// ---- smoothing elements (T + 1 of them, the last is (m_T, 0, L_T))         parallel/_smoothing.py:47-57,72-85
template <class S, int MN>
__global__ void __launch_bounds__(kBlock)
k_selems(ModelArgs<S> ma, const S* __restrict__ fm, const S* __restrict__ fL, int n, long long T,
         S* __restrict__ elems) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. T
  const long long seq = blockIdx.y;
  if (k > T) return;
  const S* m = fm + (seq * (T + 1) + k) * n;
  const S* L = fL + (seq * (T + 1) + k) * n * n;
  S* e = elems + (seq * (T + 1) + k) * se_size(n);
  S *eg = e, *eE = e + n, *eD = eE + n * n;
  if (k == T) {
    for (int i = 0; i < n; ++i) eg[i] = m[i];
    for (int i = 0; i < n * n; ++i) { eE[i] = S(0); eD[i] = L[i]; }
    return;
  }
  const Model<S> p = ma.at(seq, k);
  S Phi[2 * MN][2 * MN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      S v = S(0);
      for (int q = 0; q < n; ++q) v += p.F[i * n + q] * L[q * n + j];
      Phi[i][j] = v;
      Phi[i][n + j] = p.Q[i * n + j];
      Phi[n + i][j] = L[i * n + j];
      Phi[n + i][n + j] = S(0);
    }
  g_tria<S>(&Phi[0][0], 2 * n, 2 * n, 2 * MN);
  // E = Phi21 Phi11^-1: E^T = Phi11^-T Phi21^T
  S Et[MN][MN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Et[i][j] = Phi[n + j][i];
  g_solve_lower_t<S>(&Phi[0][0], 2 * MN, &Et[0][0], MN, n, n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      eE[i * n + j] = Et[j][i];
      eD[i * n + j] = (j <= i) ? Phi[n + i][n + j] : S(0);
    }
  }
  for (int i = 0; i < n; ++i) {
    S s = m[i];
    for (int j = 0; j < n; ++j) {
      S pm = p.b[j];
      for (int q = 0; q < n; ++q) pm += p.F[j * n + q] * m[q];
      s -= Et[j][i] * pm;
    }
    eg[i] = s;
  }
}

template <class S>
__global__ void k_smoothed_out(const S* __restrict__ elems, int n, long long T1, S* __restrict__ sm,
                               S* __restrict__ sL) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long seq = blockIdx.y;
  if (k >= T1) return;
  const S* e = elems + (seq * T1 + k) * se_size(n);
  for (int i = 0; i < n; ++i) sm[(seq * T1 + k) * n + i] = e[i];
  for (int i = 0; i < n * n; ++i) sL[(seq * T1 + k) * n * n + i] = e[n + n * n + i];
}

// generic tria for psqrt_tria_batched: A [batch, rows, cols] -> L [batch, rows, rows], rows <= 16, any cols.
// Streaming: L <- tria([L | next block of columns]) keeps L L^T = sum of the blocks' products, so only a
// [rows][2 rows] panel lives in local memory whatever cols is.
template <class S, int MN>
__global__ void __launch_bounds__(kBlock)
k_tria_stream(const S* __restrict__ A, S* __restrict__ L, int rows, int cols, long long batch) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch) return;
  const S* a = A + t * (long long)rows * cols;
  S W[MN][2 * MN];
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < rows; ++j) W[i][j] = S(0);
  for (int c0 = 0; c0 < cols; c0 += rows) {
    const int w = (cols - c0 < rows) ? cols - c0 : rows;
    for (int i = 0; i < rows; ++i)
      for (int j = 0; j < rows; ++j) W[i][rows + j] = (j < w) ? a[i * cols + c0 + j] : S(0);
    g_tria<S>(&W[0][0], rows, 2 * rows, 2 * MN);
  }
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < rows; ++j) L[(t * rows + i) * rows + j] = (j <= i) ? W[i][j] : S(0);
}

// generic rank-one up/downdates for psqrt_chol_update_batched (parsmooth/_utils.py:13-19,39-81, the column sweep of
// Krause & Igel with the non-finite -> 0 guard of line 80), n <= 16
template <class S, int MN>
__global__ void __launch_bounds__(kBlock)
k_chol_update_any(S* __restrict__ Lp, const S* __restrict__ V, int n, int kvec, S alpha, long long batch) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch) return;
  S L[MN][MN], om[MN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) L[i][j] = Lp[(t * n + i) * n + j];
  for (int v = 0; v < kvec; ++v) {
    for (int i = 0; i < n; ++i) om[i] = V[(t * kvec + v) * n + i];
    S b = S(1);
    for (int j = 0; j < n; ++j) {
      const S d = L[j][j], w = om[j];
      const S nd = sqrt(d * d + alpha / b * (w * w));
      const S gamma = d * d * b + alpha * (w * w);
      for (int i = 0; i < n; ++i) om[i] -= (w / d) * L[i][j];
      for (int i = 0; i < n; ++i) {
        const S nc = nd * (L[i][j] / d + (alpha * w / gamma) * om[i]);
        L[i][j] = nc;
      }
      b += alpha * (w / d) * (w / d);
      L[j][j] = nd;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        S x = (j <= i) ? L[i][j] : S(0);
        if (!isfinite(x)) x = S(0);
        L[i][j] = x;
      }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) Lp[(t * n + i) * n + j] = L[i][j];
}

// ---- host side ---------------------------------------------------------------------------------------------
inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

template <class S>
size_t ws_elems(int n, long long T, long long B) {   // two ping-pong element arrays sized for the larger kind + ell terms
  const size_t fe = (size_t)fe_size(n) * T, se = (size_t)se_size(n) * (T + 1);
  const size_t per = fe > se ? fe : se;
  return 2 * per * B + (sizeof(double) / sizeof(S)) * (size_t)T * B + 64;
}

template <class S, int MN>
int run(const ModelArgs<S>& ma, const S* y, const S* m0, const S* L0, int n, int ny, long long T, long long B, S* fm,
        S* fL, S* sm, S* sL, double* ell, S* ws, cudaStream_t st) {
  const size_t fe = (size_t)fe_size(n) * T, se = (size_t)se_size(n) * (T + 1);
  const size_t per = (fe > se ? fe : se) * (size_t)B;
  S* bufA = ws;
  S* bufB = ws + per;
  double* terms = reinterpret_cast<double*>(ws + 2 * per + (((uintptr_t)(ws + 2 * per) & 7) ? 1 : 0));
  const dim3 gT((unsigned)ceil_div(T, kBlock), (unsigned)B, 1), gT1((unsigned)ceil_div(T + 1, kBlock), (unsigned)B, 1);
  if (y) {
    k_felems<S, MN><<<gT, kBlock, 0, st>>>(ma, y, m0, L0, n, ny, T, bufA);
    S *cur = bufA, *nxt = bufB;
    for (long long d = 1; d < T; d <<= 1) {
      k_level<S, MN, false><<<gT, kBlock, 0, st>>>(cur, nxt, n, T, d);
      S* t = cur; cur = nxt; nxt = t;
    }
    k_filtered_out<S><<<gT1, kBlock, 0, st>>>(cur, m0, L0, n, T, fm, fL);
    if (ell) {
      k_ell_terms<S, MN><<<gT, kBlock, 0, st>>>(ma, y, fm, fL, n, ny, T, terms);
      k_sum_terms<<<(unsigned)B, 256, 0, st>>>(terms, T, ell);
    }
  }
  if (sm && sL) {
    k_selems<S, MN><<<gT1, kBlock, 0, st>>>(ma, fm, fL, n, T, bufA);
    S *cur = bufA, *nxt = bufB;
    for (long long d = 1; d < T + 1; d <<= 1) {
      k_level<S, MN, true><<<gT1, kBlock, 0, st>>>(cur, nxt, n, T + 1, d);
      S* t = cur; cur = nxt; nxt = t;
    }
    k_smoothed_out<S><<<gT1, kBlock, 0, st>>>(cur, n, T + 1, sm, sL);
  }
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

template <class S>
int dispatch(const ModelArgs<S>& ma, const S* y, const S* m0, const S* L0, int n, int ny, long long T, long long B,
             S* fm, S* fL, S* sm, S* sL, double* ell, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 1 || n > 16 || ny < 0 || ny > 16 || T <= 0 || B <= 0 || B > 65535) return PSQRT_EUNSUPPORTED;
  if (!ws || ws_bytes < ws_elems<S>(n, T, B) * sizeof(S)) return PSQRT_EWORKSPACE;
  const int mx = n > ny ? n : ny;
  if (mx <= 8) return run<S, 8>(ma, y, m0, L0, n, ny, T, B, fm, fL, sm, sL, ell, (S*)ws, st);
  return run<S, 16>(ma, y, m0, L0, n, ny, T, B, fm, fL, sm, sL, ell, (S*)ws, st);
}

template <class S>
ModelArgs<S> args_of(const S* F, const S* Q, const S* b, const S* H, const S* R, const S* c, const int64_t* ts,
                     const int64_t* bs) {
  ModelArgs<S> a;
  a.F = F; a.Q = Q; a.b = b; a.H = H; a.R = R; a.c = c;
  a.tF = ts[0]; a.tQ = ts[1]; a.tb = ts[2]; a.tH = ts[3]; a.tR = ts[4]; a.tc = ts[5];
  a.sF = bs[0]; a.sQ = bs[1]; a.sb = bs[2]; a.sH = bs[3]; a.sR = bs[4]; a.sc = bs[5];
  return a;
}

}  // namespace generic
}  // namespace psq

using namespace psq::generic;

extern "C" {

int psqrt_supported_generic(int nx, int ny) { return nx >= 1 && nx <= 16 && ny >= 0 && ny <= 16; }

size_t psqrt_generic_workspace_bytes(int nx, int64_t T, int64_t batch, int fp32) {
  if (nx < 1 || nx > 16 || T <= 0 || batch <= 0) return 0;
  return fp32 ? ws_elems<float>(nx, T, batch) * sizeof(float) : ws_elems<double>(nx, T, batch) * sizeof(double);
}

int psqrt_filter_smoother_generic(const psqrt_ssm* s, const double* y, const double* m0, const double* L0, int nx,
                                  int ny, int64_t T, int64_t batch, double* fm, double* fL, double* sm, double* sL,
                                  double* ell, void* ws, size_t ws_bytes, void* stream) {
  if (!s || !s->F || !s->cholQ || !s->b || !fm || !fL || ((sm == nullptr) != (sL == nullptr))) return PSQRT_EINVAL;
  if (s->fused_model != PSQRT_FUSED_NONE) return PSQRT_EINVAL;
  if (y && (!s->H || !s->cholR || !s->c || !m0 || !L0 || ny <= 0)) return PSQRT_EINVAL;
  const int64_t ts[6] = {s->F_ts, s->cholQ_ts, s->b_ts, s->H_ts, s->cholR_ts, s->c_ts};
  const int64_t bs[6] = {s->F_bs, s->cholQ_bs, s->b_bs, s->H_bs, s->cholR_bs, s->c_bs};
  return dispatch<double>(args_of<double>(s->F, s->cholQ, s->b, s->H, s->cholR, s->c, ts, bs), y, m0, L0, nx, ny, T,
                          batch, fm, fL, sm, sL, ell, ws, ws_bytes, (cudaStream_t)stream);
}

int psqrt_filter_smoother_f32(const float* F, const float* cholQ, const float* b, const float* H, const float* cholR,
                              const float* c, const int64_t* time_strides, const int64_t* batch_strides,
                              const float* y, const float* m0, const float* L0, int nx, int ny, int64_t T,
                              int64_t batch, float* fm, float* fL, float* sm, float* sL, double* ell, void* ws,
                              size_t ws_bytes, void* stream) {
  if (!F || !cholQ || !b || !H || !cholR || !c || !time_strides || !batch_strides || !y || !m0 || !L0 || !fm || !fL ||
      ny <= 0 || ((sm == nullptr) != (sL == nullptr)))
    return PSQRT_EINVAL;
  return dispatch<float>(args_of<float>(F, cholQ, b, H, cholR, c, time_strides, batch_strides), y, m0, L0, nx, ny, T,
                         batch, fm, fL, sm, sL, ell, ws, ws_bytes, (cudaStream_t)stream);
}

int psqrt_tria_generic(const double* A, double* L, int rows, int cols, int64_t batch, void* stream) {
  if (!A || !L || rows < 1 || rows > 16 || cols < 1 || batch <= 0) return PSQRT_EINVAL;
  k_tria_stream<double, 16><<<(unsigned)ceil_div(batch, kBlock), kBlock, 0, (cudaStream_t)stream>>>(A, L, rows, cols,
                                                                                                  batch);
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

int psqrt_chol_update_generic(double* L, const double* V, int n, int k, double alpha, int64_t batch, void* stream) {
  if (!L || !V || n < 1 || n > 16 || k < 0 || batch <= 0) return PSQRT_EINVAL;
  k_chol_update_any<double, 16><<<(unsigned)ceil_div(batch, kBlock), kBlock, 0, (cudaStream_t)stream>>>(L, V, n, k,
                                                                                                      alpha, batch);
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

}  // extern "C"
