// psqrt_sampler.cu -- pathwise (joint) samples of the smoothing distribution, square-root form.
//
// Reference: parsmooth/_pathwise_sampler.py:13-38 (parallel), 61-81, 109-124.  The per-step gain, offset
// and factor (G_t, inc_m, inc_L) of _sqrt_gain_and_inc are exactly the smoothing element (E_t, g_t, D_t)
// of parallel/_smoothing.py:72-85, and the last-state draw is its terminal element (0, m_T, L_T), so the
// caller builds them with psqrt_smoother_elements and this file runs the affine recursion
//     x_t = E_t x_{t+1} + g_t + D_t eps_t ,   t = n_el - 1 .. 0 ,   x_{n_el} = 0
// for S independent samples.  Upstream scans (G, e) pairs with e of shape [S, n] per element
// (associative_scan over 2 log2 T levels of [T, S, n] arrays).  Here samples are the parallel axis and
// time is cut into P chunks only as far as the GPU needs more threads than S:
//   k_sample_gprod   thread per chunk:            G_c = E_{k0} ... E_{k1-1}
//   k_sample_sweep<false>  thread per (chunk, sample):  e_c = the recursion through the chunk from x = 0
//   k_sample_mid           thread per sample:           chunk-boundary states, x_c = G_c x_{c+1} + e_c (P steps)
//   k_sample_sweep<true>   thread per (chunk, sample):  the recursion from the boundary state, samples written
// Consecutive threads are consecutive samples, so the draws eps[t, s, :] and the outputs are read and
// written as contiguous 32 n-double runs per warp; E_t, g_t, D_t are warp-uniform (broadcast) loads.
// The pass is HBM-bound: 8 n bytes of draws in (twice when P > 1) and 8 n bytes of samples out per
// (t, s) against 4 n^2 flops.
//
// Column signs of D_t are arbitrary (tria, _utils.py:22-24) and change WHICH sample a given draw maps to,
// not the distribution; D is used with its diagonal made non-negative so the result is a deterministic
// function of (inputs, draws) whatever Householder sign convention produced D.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <type_traits>

#include "psqrt.h"

namespace psq {
namespace {

template <int I, int E, class F>
__device__ __forceinline__ void static_for_s(F&& f) {   // compile-time loop: register arrays need static indices
  if constexpr (I < E) {
    f(std::integral_constant<int, I>{});
    static_for_s<I + 1, E>(f);
  }
}

// draws of element t: eps[(t + 1) % n_el]  (eps[0] makes the last state, _pathwise_sampler.py:71,79)
__device__ __forceinline__ long long eps_row(long long t, long long n_el) { return (t + 1 == n_el) ? 0 : t + 1; }

template <int N>
__global__ void k_sample_gprod(const double* __restrict__ E, long long n_el, int K, long long P, double* __restrict__ G) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P) return;
  const long long k0 = c * K, k1 = (k0 + K < n_el) ? k0 + K : n_el;
  double A[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) A[i][j] = (i == j) ? 1.0 : 0.0;
  for (long long t = k1 - 1; t >= k0; --t) {   // A <- E_t A
    double B[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(__ldg(E + t * N * N + i * N + k), A[k][j], s);
        B[i][j] = s;
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) A[i][j] = B[i][j];
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) G[c * N * N + i * N + j] = A[i][j];
}

// APPLY = false: e_c (from x = 0) -> ebuf[c][s];  APPLY = true: from the boundary state xb[c + 1][s], writing samples.
// All threads of a CTA work on the SAME chunk (consecutive samples), so the coefficients (g_t, E_t, sign-
// normalised D_t) of a tile of TILE steps are staged once per CTA in shared memory -- fetched into registers while
// the previous tile is being processed, so their latency is hidden -- and read back as broadcasts; each thread's
// draws come through a 4-deep register ring (loads issued 4 steps ahead of their use).  ncu on the first version
// (coefficients and draws loaded at the point of use): 18-24 warps per issue stalled on long_scoreboard.
// BS = threads per CTA, TILE = steps staged at a time: 128 / 16 normally, 32 / 4 when there are at most 32 samples
// (a 128-thread CTA would be three quarters idle and cap the number of resident chunks).
template <int N, bool APPLY, int BS, int TILE>
__global__ void __launch_bounds__(BS, BS == 32 ? 16 : 4)
k_sample_sweep(const double* __restrict__ g, const double* __restrict__ E, const double* __restrict__ D,
               const double* __restrict__ eps, long long n_el, long long S, int K, long long P,
               const double* __restrict__ xb, double* __restrict__ out) {
  constexpr int NN = N * N, NC = N + 2 * NN;          // per step: g [N], E [N][N], D [N][N]
  static_assert(TILE % 4 == 0, "the draw ring has 4 slots");
  constexpr int NV = (TILE * NC + BS - 1) / BS;       // staged values per thread and tile
  __shared__ double coef[TILE * NC];
  const int tid = threadIdx.x;
  const long long s = (long long)blockIdx.x * BS + tid;
  const bool live = s < S;                            // no early return: every thread stages and synchronises
  const long long c = blockIdx.y;
  const long long k0 = c * K, k1 = (k0 + K < n_el) ? k0 + K : n_el;
  double x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = (APPLY && live && c + 1 < P) ? xb[((c + 1) * S + s) * N + i] : 0.0;

  double sv[NV];
  auto fetch = [&](long long hi) {                     // coefficients of steps hi, hi - 1, ... -> registers
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int idx = tid + v * BS;
      const int u = idx / NC, f = idx % NC;
      const long long t = hi - u;
      double val = 0.0;
      if (idx < TILE * NC && t >= k0) {
        if (f < N) {
          val = __ldg(g + t * N + f);
        } else if (f < N + NN) {
          val = __ldg(E + t * NN + (f - N));
        } else {
          const int q = f - N - NN, i = q / N, j = q % N;
          const double d = __ldg(D + t * NN + q), dj = __ldg(D + t * NN + j * N + j);
          val = (j <= i) ? (dj < 0.0 ? -d : d) : 0.0;   // lower triangle, non-negative diagonal
        }
      }
      sv[v] = val;
    }
  };
  double er[4][N];
  auto eload = [&](auto slot, long long t) {
    if (live && t >= k0) {
      const double* ep = eps + (eps_row(t, n_el) * S + s) * N;
#pragma unroll
      for (int i = 0; i < N; ++i) er[decltype(slot)::value][i] = __ldcs(ep + i);   // streamed once per sweep
    }
  };
  long long hi = k1 - 1;
  fetch(hi);
  eload(std::integral_constant<int, 0>{}, hi);
  eload(std::integral_constant<int, 1>{}, hi - 1);
  eload(std::integral_constant<int, 2>{}, hi - 2);
  eload(std::integral_constant<int, 3>{}, hi - 3);
#pragma unroll 1
  while (hi >= k0) {
    __syncthreads();                                   // the previous tile has been read by everyone
#pragma unroll
    for (int v = 0; v < NV; ++v)
      if (tid + v * BS < TILE * NC) coef[tid + v * BS] = sv[v];
    __syncthreads();
    if (hi - TILE >= k0) fetch(hi - TILE);             // in flight during this tile's steps
#pragma unroll 1
    for (int u0 = 0; u0 < TILE; u0 += 4)               // rolled: registers stay low enough for ~16 warps per SM
    static_for_s<0, 4>([&](auto uc) {
      constexpr int slot = decltype(uc)::value;        // ring slot = step index mod 4 (TILE is a multiple of 4)
      const int u = u0 + slot;
      const long long t = hi - u;
      if (t >= k0) {                                   // uniform over the CTA
        double e[N];
#pragma unroll
        for (int i = 0; i < N; ++i) e[i] = er[slot][i];
        eload(std::integral_constant<int, slot>{}, t - 4);
        const double* cf = coef + u * NC;
        double y[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
          double a = cf[i];
#pragma unroll
          for (int j = 0; j < N; ++j) a = fma(cf[N + i * N + j], x[j], a);
#pragma unroll
          for (int j = 0; j <= i; ++j) a = fma(cf[N + NN + i * N + j], e[j], a);
          y[i] = a;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = y[i];
        if (APPLY && live) {
          double* o = out + (t * S + s) * N;
#pragma unroll
          for (int i = 0; i < N; ++i) __stcs(o + i, x[i]);
        }
      }
    });
    hi -= TILE;
  }
  if (!APPLY && live) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[(c * S + s) * N + i] = x[i];
  }
}

// boundary states: xb[c][s] = state at element index c K (start of chunk c), from the last chunk backwards
template <int N>
__global__ void k_sample_mid(const double* __restrict__ G, const double* __restrict__ ebuf, long long S, long long P,
                             double* __restrict__ xb) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double x[N];
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = 0.0;
  // the recursion is a chain of P dependent mat-vecs; G_c and e_c do not depend on it, so the next chunk's are
  // loaded while the current one is applied (otherwise every step pays an exposed L2 / DRAM round trip)
  double Gn[N][N], en[N];
  auto load = [&](long long c) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      en[i] = ebuf[(c * S + s) * N + i];
#pragma unroll
      for (int j = 0; j < N; ++j) Gn[i][j] = __ldg(G + c * N * N + i * N + j);
    }
  };
  if (P > 1) load(P - 1);
  for (long long c = P - 1; c >= 1; --c) {   // xb[0] is never read
    double Gc[N][N], ec[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      ec[i] = en[i];
#pragma unroll
      for (int j = 0; j < N; ++j) Gc[i][j] = Gn[i][j];
    }
    if (c > 1) load(c - 1);
    double y[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double a = ec[i];
#pragma unroll
      for (int j = 0; j < N; ++j) a = fma(Gc[i][j], x[j], a);
      y[i] = a;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      x[i] = y[i];
      xb[(c * S + s) * N + i] = y[i];
    }
  }
}

struct SamplerPlan {
  int K;
  long long P;
  size_t g_off, e_off, x_off, bytes;
};

SamplerPlan sampler_plan(int n, long long n_el, long long S) {
  SamplerPlan p;
  // enough (chunk, sample) threads for ~8 resident warps per scheduler on 148 SMs, chunks not shorter than 8;
  // the boundary recursion (k_sample_mid) is sequential over the P chunks, the sweeps over the K = n_el / P
  // steps of a chunk: with few samples P stays near sqrt(n_el) so that neither chain dominates
  const long long want = 148LL * 2048;
  long long P = (want + S - 1) / S;
  const long long maxP = n_el / 8 > 0 ? n_el / 8 : 1;
  if (P > maxP) P = maxP;
  long long root = 1;
  while ((root + 1) * (root + 1) <= n_el) ++root;
  if (P > 2 * root) P = 2 * root;
  if (P > 65535) P = 65535;
  if (P < 1) P = 1;
  p.K = (int)((n_el + P - 1) / P);
  p.P = (n_el + p.K - 1) / p.K;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  p.g_off = 0;
  p.e_off = al(sizeof(double) * (size_t)p.P * n * n);
  p.x_off = p.e_off + al(sizeof(double) * (size_t)p.P * S * n);
  p.bytes = p.x_off + al(sizeof(double) * (size_t)(p.P + 1) * S * n);
  if (p.P == 1) p.bytes = 256;
  return p;
}

template <int N>
int run(const double* g, const double* E, const double* D, const double* eps, double* samples, long long n_el,
        long long S, void* ws, size_t ws_bytes, cudaStream_t st) {
  const SamplerPlan p = sampler_plan(N, n_el, S);
  if (p.P > 1 && (!ws || ws_bytes < p.bytes)) return PSQRT_EWORKSPACE;
  const bool narrow = S <= 32;
  const unsigned bs = narrow ? 32 : 128;
  const dim3 grid((unsigned)((S + bs - 1) / bs), (unsigned)p.P, 1);
  double* Gc = nullptr;
  double* eb = nullptr;
  double* xb = nullptr;
  if (p.P > 1) {
    char* w = static_cast<char*>(ws);
    Gc = reinterpret_cast<double*>(w + p.g_off);
    eb = reinterpret_cast<double*>(w + p.e_off);
    xb = reinterpret_cast<double*>(w + p.x_off);
    k_sample_gprod<N><<<(unsigned)((p.P + 127) / 128), 128, 0, st>>>(E, n_el, p.K, p.P, Gc);
    if (narrow) k_sample_sweep<N, false, 32, 4><<<grid, bs, 0, st>>>(g, E, D, eps, n_el, S, p.K, p.P, nullptr, eb);
    else k_sample_sweep<N, false, 128, 16><<<grid, bs, 0, st>>>(g, E, D, eps, n_el, S, p.K, p.P, nullptr, eb);
    k_sample_mid<N><<<(unsigned)((S + 127) / 128), 128, 0, st>>>(Gc, eb, S, p.P, xb);
  }
  if (narrow) k_sample_sweep<N, true, 32, 4><<<grid, bs, 0, st>>>(g, E, D, eps, n_el, S, p.K, p.P, xb, samples);
  else k_sample_sweep<N, true, 128, 16><<<grid, bs, 0, st>>>(g, E, D, eps, n_el, S, p.K, p.P, xb, samples);
  return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA;
}

}  // namespace
}  // namespace psq

extern "C" size_t psqrt_sampler_workspace_bytes(int nx, int64_t n_elements, int64_t n_samples) {
  if (nx < 1 || nx > 8 || n_elements <= 0 || n_samples <= 0) return 0;
  return psq::sampler_plan(nx, n_elements, n_samples).bytes;
}

extern "C" int psqrt_sample_paths(const double* g, const double* E, const double* D, const double* eps,
                                  double* samples, int nx, int64_t n_elements, int64_t n_samples, void* ws,
                                  size_t ws_bytes, void* stream) {
  if (!g || !E || !D || !eps || !samples || n_elements <= 0 || n_samples <= 0) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  switch (nx) {
    case 1: return psq::run<1>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    case 2: return psq::run<2>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    case 3: return psq::run<3>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    case 4: return psq::run<4>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    case 5: return psq::run<5>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    case 6: return psq::run<6>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    case 8: return psq::run<8>(g, E, D, eps, samples, n_elements, n_samples, ws, ws_bytes, st);
    default: return PSQRT_EUNSUPPORTED;
  }
}
