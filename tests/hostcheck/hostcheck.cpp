// hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles the product's per-thread fp64 algebra
// (sqrt-parallel-smoothers_b200/csrc/psqrt_math.cuh, __host__ __device__ templates) with g++
// and runs it on the CPU so the arithmetic and the chunked three-sweep algorithm can be
// checked against the oracle in this GPU-less container.  The warp / CTA scans of the CUDA
// kernels are emulated lane by lane with the same Kogge-Stone schedule.  Never loaded by the
// product package.
#include <cstdint>
#include <cstring>
#include <vector>

#include "psqrt_math.cuh"

using namespace psq;

namespace {

struct HSsm {
  const double *F, *Q, *b, *H, *R, *c, *y;
  long long tF, tQ, tb, tH, tR, tc, ty;
};
inline StepPtrs sp(const HSsm& a, long long k) {
  StepPtrs p;
  p.F = a.F + k * a.tF; p.Q = a.Q + k * a.tQ; p.bq = a.b + k * a.tb;
  p.H = a.H ? a.H + k * a.tH : nullptr; p.R = a.R ? a.R + k * a.tR : nullptr;
  p.c = a.c ? a.c + k * a.tc : nullptr; p.y = a.y ? a.y + k * a.ty : nullptr;
  return p;
}

template <int N> FElem<N> comb(const FElem<N>& a, const FElem<N>& b) { return filtering_combine<N>(a, b); }
template <int N> SElem<N> comb(const SElem<N>& a, const SElem<N>& b) { return smoothing_combine<N>(a, b); }

// Kogge-Stone inclusive scan over 32 "lanes" exactly as warp_scan_inclusive does.
template <class E> void ks_scan(E* lanes, bool rev) {
  for (int d = 1; d < 32; d <<= 1) {
    E nxt[32];
    for (int l = 0; l < 32; ++l) {
      int src = rev ? l + d : l - d;
      bool valid = rev ? (l + d < 32) : (l >= d);
      nxt[l] = valid ? comb(lanes[src], lanes[l]) : lanes[l];
    }
    for (int l = 0; l < 32; ++l) lanes[l] = nxt[l];
  }
}

template <int N> void load_dense(const double* m, const double* L, Gauss<N>& x) {
  for (int i = 0; i < N; ++i) { x.m[i] = m[i]; for (int j = 0; j <= i; ++j) x.Lc(i, j) = L[i * N + j]; }
}
template <int N> void store_dense(double* m, double* L, const Gauss<N>& x) {
  for (int i = 0; i < N; ++i) { m[i] = x.m[i]; for (int j = 0; j < N; ++j) L[i * N + j] = j <= i ? x.Lc(i, j) : 0.0; }
}

// The whole pass with the kernels' structure: chunks of K steps, 32 chunks per warp,
// warp scans, sequential mid-level exclusive scan, apply sweeps.
template <int N, int NY>
int pass(const HSsm& a, long long T, int K, const double* m0, const double* L0, double* fm, double* fL, double* sm,
         double* sL, double* ell_out) {
  const long long P = (T + K - 1) / K, Ppad = (P + 127) / 128 * 128, M = Ppad / 32;
  std::vector<FElem<N>> chunk_pref(Ppad), chunk_own(Ppad), warp_tot(M);
  for (long long w = 0; w < M; ++w) {
    FElem<N> lanes[32];
    for (int l = 0; l < 32; ++l) {
      long long c = w * 32 + l, k0 = c * K, k1 = std::min<long long>(T, k0 + K);
      FAcc<N> acc;
      acc.set_identity();
      for (long long k = k0; k < k1; ++k) {
        const bool last = (k + 1 == k1);
        filter_reduce_step<N, NY>(acc, sp(a, k), [&](const double (&FA)[N][N], const double (&mp)[N],
                                                     const double (&Np)[N][2 * N], const FAcc<N>& s) {
          if (!last) return;   // the summary with its last step predict-only (what k_filter_reduce stores)
          FElem<N>& o = chunk_own[c];
          for (int i = 0; i < N; ++i) {
            o.b(i) = mp[i];
            o.eta(i) = s.eta[i];
            for (int j = 0; j < N; ++j) o.A(i, j) = FA[i][j];
            for (int j = 0; j <= i; ++j) { o.U(i, j) = Np[i][j]; o.Z(i, j) = s.Z[i * (i + 1) / 2 + j]; }
          }
        });
      }
      acc.to_elem(lanes[l]);
    }
    ks_scan(lanes, false);
    for (int l = 0; l < 32; ++l) {
      if (l == 0) chunk_pref[w * 32].set_identity(); else chunk_pref[w * 32 + l] = lanes[l - 1];
    }
    warp_tot[w] = lanes[31];
  }
  {  // mid-level exclusive scan (in place)
    FElem<N> run; run.set_identity();
    for (long long w = 0; w < M; ++w) { FElem<N> x = warp_tot[w]; warp_tot[w] = run; run = comb(run, x); }
  }
  std::vector<SElem<N>> chunk_suf(Ppad), warp_stot(M);
  double ell = 0.0;
  for (long long w = 0; w < M; ++w) {
    SElem<N> lanes[32];
    for (int l = 0; l < 32; ++l) {
      long long c = w * 32 + l, k0 = c * K, k1 = std::min<long long>(T, k0 + K);
      Gauss<N> x; load_dense<N>(m0, L0, x);
      filtering_apply<N>(x, warp_tot[w]);
      filtering_apply<N>(x, chunk_pref[c]);
      if (c == 0) store_dense<N>(fm, fL, x);
      // the chunk's smoothing total straight from its filtering summary (no per-step combines)
      if (k0 < k1) chunk_smoothing_total<N>(x, chunk_own[c], lanes[l]); else lanes[l].set_identity();
      // inside the chunk the recursion carries a dense factor; x receives the triangular outputs
      GaussD<N> xd;
      for (int i = 0; i < N; ++i) { xd.m[i] = x.m[i]; for (int j = 0; j < N; ++j) xd.Y[i][j] = j <= i ? x.Lc(i, j) : 0.0; }
      for (long long k = k0; k <= k1 && k0 < k1; ++k) {   // step k also triangularises the state at index k
        if (k < k1) ell += kalman_step_dense<N, NY, true>(xd, sp(a, k), x); else gaussd_tri<N>(xd, x);
        if (k > k0) store_dense<N>(fm + k * N, fL + k * N * N, x);
      }
    }
    ks_scan(lanes, true);
    for (int l = 0; l < 32; ++l) {
      if (l == 31) chunk_suf[w * 32 + 31].set_identity(); else chunk_suf[w * 32 + l] = lanes[l + 1];
    }
    warp_stot[w] = lanes[0];
  }
  *ell_out = ell;
  if (!sm) return 0;
  {
    SElem<N> run; run.set_identity();
    for (long long w = M - 1; w >= 0; --w) { SElem<N> x = warp_stot[w]; warp_stot[w] = run; run = comb(run, x); }
  }
  for (long long c = 0; c < Ppad; ++c) {
    long long k0 = c * K, k1 = std::min<long long>(T, k0 + K);
    if (k0 >= k1) continue;
    Gauss<N> xs; load_dense<N>(fm + T * N, fL + T * N * N, xs);
    if (k1 == T) store_dense<N>(sm + T * N, sL + T * N * N, xs);
    smoothing_apply<N>(xs, warp_stot[c / 32]);
    smoothing_apply<N>(xs, chunk_suf[c]);
    for (long long k = k1 - 1; k >= k0; --k) {
      Gauss<N> xf; load_dense<N>(fm + k * N, fL + k * N * N, xf);
      rts_step<N>(xs, xf, sp(a, k));
      store_dense<N>(sm + k * N, sL + k * N * N, xs);
    }
  }
  return 0;
}

template <int N> void fload(const double* A, const double* b, const double* U, const double* e, const double* Z, FElem<N>& x) {
  for (int r = 0; r < N; ++r) {
    x.b(r) = b[r]; x.eta(r) = e[r];
    for (int q = 0; q < N; ++q) x.A(r, q) = A[r * N + q];
    for (int q = 0; q <= r; ++q) { x.U(r, q) = U[r * N + q]; x.Z(r, q) = Z[r * N + q]; }
  }
}
template <int N> int fcombine(const double** in, double** out) {
  FElem<N> x, y; fload<N>(in[0], in[1], in[2], in[3], in[4], x); fload<N>(in[5], in[6], in[7], in[8], in[9], y);
  FElem<N> o = filtering_combine<N>(x, y);
  for (int r = 0; r < N; ++r) {
    out[1][r] = o.b(r); out[3][r] = o.eta(r);
    for (int q = 0; q < N; ++q) {
      out[0][r * N + q] = o.A(r, q);
      out[2][r * N + q] = q <= r ? o.U(r, q) : 0.0;
      out[4][r * N + q] = q <= r ? o.Z(r, q) : 0.0;
    }
  }
  return 0;
}
template <int N> int scombine(const double** in, double** out) {
  SElem<N> x, y;
  for (int s = 0; s < 2; ++s) {
    SElem<N>& e = s ? y : x;
    for (int r = 0; r < N; ++r) {
      e.g(r) = in[3 * s][r];
      for (int q = 0; q < N; ++q) e.E(r, q) = in[3 * s + 1][r * N + q];
      for (int q = 0; q <= r; ++q) e.D(r, q) = in[3 * s + 2][r * N + q];
    }
  }
  SElem<N> o = smoothing_combine<N>(x, y);
  for (int r = 0; r < N; ++r) {
    out[0][r] = o.g(r);
    for (int q = 0; q < N; ++q) { out[1][r * N + q] = o.E(r, q); out[2][r * N + q] = q <= r ? o.D(r, q) : 0.0; }
  }
  return 0;
}
template <int R> int tria_stream(const double* A, int C, double* L) {
  double Lt[R][R];
  for (int r = 0; r < R; ++r) for (int q = 0; q < R; ++q) Lt[r][q] = 0.0;
  for (int c0 = 0; c0 < C; c0 += 4) {
    double W[R][4];
    for (int r = 0; r < R; ++r) for (int q = 0; q < 4; ++q) W[r][q] = (c0 + q < C) ? A[r * C + c0 + q] : 0.0;
    tria_append<R, 4>([&](int r, int q) -> double& { return Lt[r][q]; }, W);
  }
  for (int r = 0; r < R; ++r) for (int q = 0; q < R; ++q) L[r * R + q] = q <= r ? Lt[r][q] : 0.0;
  return 0;
}
template <int N> int cholupd(double* L, const double* V, int k, double alpha) {
  double Lt[N][N], w[N];
  for (int r = 0; r < N; ++r) for (int q = 0; q < N; ++q) Lt[r][q] = L[r * N + q];
  for (int v = 0; v < k; ++v) { for (int r = 0; r < N; ++r) w[r] = V[v * N + r]; chol_update<N>(Lt, w, alpha); }
  for (int r = 0; r < N; ++r) for (int q = 0; q < N; ++q) L[r * N + q] = Lt[r][q];
  return 0;
}
template <int N, int NY> int felems(const HSsm& a, long long T, const double* m0, const double* L0, double* A, double* b,
                                    double* U, double* eta, double* Z, const double* fm, const double* fL, double* terms) {
  for (long long k = 0; k < T; ++k) {
    filtering_element<N, NY>(sp(a, k), k == 0 ? m0 : nullptr, k == 0 ? L0 : nullptr, A + k * N * N, b + k * N,
                             U + k * N * N, eta + k * N, Z + k * N * N);
    if (terms) terms[k] = loglik_term<N, NY>(sp(a, k), fm + k * N, fL + k * N * N);
  }
  return 0;
}

#define DISPATCH_N(n, CALL) \
  switch (n) { case 1: { constexpr int N = 1; CALL; } case 2: { constexpr int N = 2; CALL; } \
               case 3: { constexpr int N = 3; CALL; } case 4: { constexpr int N = 4; CALL; } \
               case 5: { constexpr int N = 5; CALL; } default: return -2; }
#define DISPATCH_NY(ny, CALL) \
  switch (ny) { case 1: { constexpr int NY = 1; CALL; } case 2: { constexpr int NY = 2; CALL; } \
                case 3: { constexpr int NY = 3; CALL; } default: return -2; }

}  // namespace

extern "C" {

int hc_pass(int n, int ny, long long T, int K, const double* F, const double* Q, const double* b, const double* H,
            const double* R, const double* c, const double* y, const long long* tstrides, const double* m0,
            const double* L0, double* fm, double* fL, double* sm, double* sL, double* ell) {
  HSsm a{F, Q, b, H, R, c, y, tstrides[0], tstrides[1], tstrides[2], tstrides[3], tstrides[4], tstrides[5], ny};
  DISPATCH_N(n, DISPATCH_NY(ny, return (pass<N, NY>(a, T, K, m0, L0, fm, fL, sm, sL, ell))));
}
int hc_filter_combine(int n, const double** in, double** out) { DISPATCH_N(n, return fcombine<N>(in, out)); }
int hc_smoothing_combine(int n, const double** in, double** out) { DISPATCH_N(n, return scombine<N>(in, out)); }
int hc_tria(int rows, int cols, const double* A, double* L) { DISPATCH_N(rows, return tria_stream<N>(A, cols, L)); }
int hc_chol_update(int n, double* L, const double* V, int k, double alpha) { DISPATCH_N(n, return cholupd<N>(L, V, k, alpha)); }
int hc_filter_elements(int n, int ny, long long T, const double* F, const double* Q, const double* b, const double* H,
                       const double* R, const double* c, const double* y, const long long* tstrides, const double* m0,
                       const double* L0, double* A, double* bo, double* U, double* eta, double* Z, const double* fm,
                       const double* fL, double* terms) {
  HSsm a{F, Q, b, H, R, c, y, tstrides[0], tstrides[1], tstrides[2], tstrides[3], tstrides[4], tstrides[5], ny};
  DISPATCH_N(n, DISPATCH_NY(ny, return (felems<N, NY>(a, T, m0, L0, A, bo, U, eta, Z, fm, fL, terms))));
}

}  // extern "C"
