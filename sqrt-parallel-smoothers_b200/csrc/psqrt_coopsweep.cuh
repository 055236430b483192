// psqrt_coopsweep.cuh -- the three sweeps (K1 / K3 / K5) in sub-warp form for the larger state dimensions.
//
// One thread per chunk (psqrt_kernels.cuh) keeps every matrix entry of a step in registers.  At nx = 8 one
// triangularisation alone (8 x 16, then 12 x 12) needs more than the 255 registers a thread can have: ptxas spills
// 3 - 6 KB per thread and the sweeps run at a tenth of the roofline.  Here a group of G = 8
// lanes shares ONE chunk and holds one matrix ROW per lane:
//   * every triangularisation is the in-register Householder tria of psqrt_coop2.cuh (coop_house_shfl: the pivot row
//     travels by warp shuffles, each lane updates its own row);
//   * products with the model (F Y, H N, E L_s ...) are one output row per lane, the other operand's rows gathered
//     through the group's shared-memory buffer (each lane publishes its row, every lane reads all of them with
//     16-byte broadcast loads);
//   * only the NY reflectors of the measurement update are loop-carried, exactly like kalman_step_dense.
// A CTA is 256 threads = 32 chunks, which takes the place of the WARP of the per-thread sweeps in the three-level
// prefix hierarchy (chunk -> 32 chunks -> group -> sequence): every scratch layout, the mid-level scans K2 / K4 and
// the carry kernels are shared with the per-thread path.  What runs once per chunk (pushing the carry through the
// prefixes, the chunk's smoothing total, the scan over 32 chunk summaries) stays per-thread code in three small
// kernels of their own (k_chunk_start, k_chunk_end; the scans over the 32 chunks of a CTA: k_unit_scan) that hand the per-chunk start states to the
// sub-warp step loops through scratch.
//
// Which form runs where: nx = 8 uses the step loops of psqrt_coopsweep2.cuh (G = 4 lanes per chunk, TWO rows per lane:
// 2.72 ms per pass at T = 1e6 against 3.49 ms for the one-row-per-lane loops below and 5.93 ms per thread); the loops
// in this file are compiled for nx = 6 (opt-in with PSQRT_COOP=7: its per-thread sweeps are faster) and carry the
// idle-lane handling for N < 8.  The buffer layout, the model loaders' helpers, coop_psi11_solve and the
// once-per-chunk kernels below serve both.
//
// Formulas: the same as psq::filter_reduce_step, kalman_step_dense, rts_step (psqrt_math.cuh), i.e.
//   parsmooth/parallel/_filtering.py:100-154, _operators.py:58-77 (K1), sequential/_filtering.py:80-108 (K3),
//   parallel/_smoothing.py:72-85 + _operators.py:118-125 (K5).
#pragma once
#include "psqrt_kernels.cuh"

namespace psq {

constexpr int kCG = 8;                     // lanes per chunk
constexpr int kCChunks = 32;               // chunks per CTA (= one scan unit of 32 chunks)
constexpr int kCBlock = kCG * kCChunks;    // 256 threads
// resident CTAs per SM requested from ptxas for the sub-warp step loops (2: 128 registers, 1: 255); measured A/B in
// DESIGN.md section 5
#ifndef PSQ_COOP_MINB_K1
#define PSQ_COOP_MINB_K1 1
#endif
#ifndef PSQ_COOP_MINB_K3
#define PSQ_COOP_MINB_K3 2
#endif
#ifndef PSQ_COOP_MINB_K5
#define PSQ_COOP_MINB_K5 2
#endif

template <int N>
struct CoopSweep {
  static_assert(N <= kCG, "one row per lane");
  // row stride of the gather regions: rows start 16-byte aligned for even N, and the 8 rows of a group fall into
  // different banks for the 16-byte stores of a quarter warp
  static constexpr int RS = (N % 4 == 0) ? N + 2 : N;
  static constexpr int R0 = 0;                 // [N][RS]
  static constexpr int R1 = N * RS;            // [N][RS]
  static constexpr int HB = 2 * N * RS;        // [4][RS]   H rows / V rows
  static constexpr int V0 = HB + 4 * RS;       // [8]
  static constexpr int V1 = V0 + 8;            // [8]
  static constexpr int SZ = spread_stride(V1 + 8);
  static constexpr size_t smem_bytes() { return sizeof(double) * (size_t)SZ * kCChunks; }
};

// ---------------------------------------------------------------------------------------------------------------
// Householder reflectors of the pivot rows `top` (one per lane, row index r, rows on consecutive lanes from warp
// lane src0) applied to `top` (rows below the pivot) AND to a second row `bot` per lane (always).  TRIBLK as in
// house_rows.  Rows [0, NREFL) of top become lower-trapezoidal.  Every lane of the warp must call.
// ---------------------------------------------------------------------------------------------------------------
template <int C, int NREFL, int TRIBLK, int R>
__device__ __forceinline__ void coop_house2(double (&top)[C], double (&bot)[C], const int r, const int src0) {
  static_for<0, NREFL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int kend = (TRIBLK > 0) ? ((TRIBLK + j + 1 < C) ? TRIBLK + j + 1 : C) : C;
    if constexpr (j + 1 < kend) {
      double p[C];
#pragma unroll
      for (int k = j; k < kend; ++k) p[k] = __shfl_sync(0xffffffffu, top[k], src0 + j);
      const double alpha = p[j];
      double sigma = 0.0, sigma2 = 0.0;
#pragma unroll
      for (int k = j + 1; k < kend; k += 2) {
        sigma = fma(p[k], p[k], sigma);
        if (k + 1 < kend) sigma2 = fma(p[k + 1], p[k + 1], sigma2);
      }
      sigma += sigma2;
      double dt = 0.0, dt2 = 0.0, db = 0.0, db2 = 0.0;
#pragma unroll
      for (int k = j + 1; k < kend; k += 2) {
        dt = fma(top[k], p[k], dt);
        db = fma(bot[k], p[k], db);
        if (k + 1 < kend) {
          dt2 = fma(top[k + 1], p[k + 1], dt2);
          db2 = fma(bot[k + 1], p[k + 1], db2);
        }
      }
      dt += dt2;
      db += db2;
      const double q = fma(alpha, alpha, sigma);
      const double mask = (q != 0.0) ? 1.0 : 0.0;
      const double qs = (q != 0.0) ? q : 1.0;
      const double norm = qs * rsqrt_nr(qs);
      const double beta = -copysign(norm, alpha) * mask;
      const double v0 = alpha - beta;
      const double s = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
      dt = fma(top[j], v0, dt) * s;
      db = fma(bot[j], v0, db) * s;
      const bool below = (r > j) && (r < R);
      const double ddt = below ? dt : 0.0;
      top[j] = (r == j) ? beta : fma(-ddt, v0, top[j]);
      bot[j] = fma(-db, v0, bot[j]);
#pragma unroll
      for (int k = j + 1; k < kend; ++k) {
        top[k] = fma(-ddt, p[k], top[k]);
        bot[k] = fma(-db, p[k], bot[k]);
      }
    }
  });
}

// Z <- tria([Z | W]) with Z (N x N) lower triangular, row r per lane in z[0..N) (zeros right of the diagonal), and
// W (N x K) dense, row r per lane: reflector j touches only column j of Z and the K columns of W (tria_append).
template <int N, int K>
__device__ __forceinline__ void coop_tria_append(double (&z)[N], double (&w)[K], const int r, const int src0) {
  static_for<0, N>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    const double alpha = __shfl_sync(0xffffffffu, z[j], src0 + j);
    double pw[K];
    double sigma = 0.0, d = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      pw[a] = __shfl_sync(0xffffffffu, w[a], src0 + j);
      sigma = fma(pw[a], pw[a], sigma);
      d = fma(w[a], pw[a], d);
    }
    const double q = fma(alpha, alpha, sigma);
    const double mask = (q != 0.0) ? 1.0 : 0.0;
    const double qs = (q != 0.0) ? q : 1.0;
    const double norm = qs * rsqrt_nr(qs);
    const double beta = -copysign(norm, alpha) * mask;
    const double v0 = alpha - beta;
    const double s = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
    d = fma(z[j], v0, d) * s;
    const bool below = (r > j) && (r < N);
    const double dd = below ? d : 0.0;
    z[j] = (r == j) ? beta : fma(-dd, v0, z[j]);
#pragma unroll
    for (int a = 0; a < K; ++a) w[a] = fma(-dd, pw[a], w[a]);
  });
}

// The NY reflectors of the measurement update on [[H Np, R], [Np, 0]] (_filtering.py:126-131): pivot rows `hrow`
// live on lanes 0 .. NY-1 of the group (N + NY entries), every lane's `mrow` is one of the N bottom rows.  On return
// hrow[0..a] of lane a is row a of Psi11, mrow[0..NY) is the lane's row of Psi21 and mrow[NY..NY+N) its row of a dense
// square root of the posterior covariance.
template <int N, int NY>
__device__ __forceinline__ void coop_update_reflectors(double (&hrow)[N + NY], double (&mrow)[N + NY], const int l,
                                                       const int gbase) {
  constexpr int C = N + NY;
  static_for<0, NY>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    double p[C];
#pragma unroll
    for (int k = a; k < C; ++k) p[k] = __shfl_sync(0xffffffffu, hrow[k], gbase + a);
    const double alpha = p[a];
    double sigma = 0.0, sigma2 = 0.0, dm = 0.0, dm2 = 0.0, dh = 0.0, dh2 = 0.0;
#pragma unroll
    for (int k = a + 1; k < C; k += 2) {
      sigma = fma(p[k], p[k], sigma);
      dm = fma(mrow[k], p[k], dm);
      dh = fma(hrow[k], p[k], dh);
      if (k + 1 < C) {
        sigma2 = fma(p[k + 1], p[k + 1], sigma2);
        dm2 = fma(mrow[k + 1], p[k + 1], dm2);
        dh2 = fma(hrow[k + 1], p[k + 1], dh2);
      }
    }
    sigma += sigma2;
    dm += dm2;
    dh += dh2;
    const double q = fma(alpha, alpha, sigma);
    const double mask = (q != 0.0) ? 1.0 : 0.0;
    const double qs = (q != 0.0) ? q : 1.0;
    const double norm = qs * rsqrt_nr(qs);
    const double beta = -copysign(norm, alpha) * mask;
    const double v0 = alpha - beta;
    const double s = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
    dm = fma(mrow[a], v0, dm) * s;
    dh = fma(hrow[a], v0, dh) * s;
    const double ddh = (l > a && l < NY) ? dh : 0.0;
    mrow[a] = fma(-dm, v0, mrow[a]);
    hrow[a] = (l == a) ? beta : fma(-ddh, v0, hrow[a]);
#pragma unroll
    for (int k = a + 1; k < C; ++k) {
      mrow[k] = fma(-dm, p[k], mrow[k]);
      hrow[k] = fma(-ddh, p[k], hrow[k]);
    }
  });
}

// Psi11 (rows on lanes 0 .. NY-1, see above) and the residuals of those lanes to EVERY lane, then
// rr = Psi11^{-1} res, inverse diagonal, |rr|^2 and det Psi11.
template <int N, int NY>
__device__ __forceinline__ void coop_psi11_solve(const double (&hrow)[N + NY], const double res, const int gbase,
                                                 double (&P11)[NY][NY], double (&inv)[NY], double (&rr)[NY],
                                                 double& quad, double& det) {
  quad = 0.0;
  det = 1.0;
#pragma unroll
  for (int a = 0; a < NY; ++a) {
    double r = __shfl_sync(0xffffffffu, res, gbase + a);
#pragma unroll
    for (int q = 0; q <= a; ++q) P11[a][q] = __shfl_sync(0xffffffffu, hrow[q], gbase + a);
    inv[a] = rcp_nr(P11[a][a]);
#pragma unroll
    for (int q = 0; q < a; ++q) r = fma(-P11[a][q], rr[q], r);
    rr[a] = r * inv[a];
    quad = fma(rr[a], rr[a], quad);
    det *= P11[a][a];
  }
}

// Global-memory rows: 16-byte accesses when the launcher found every base and stride 16-byte aligned (vec), 8-byte
// accesses otherwise (a uniform branch).
template <int N>
__device__ __forceinline__ void gld_row(const double* __restrict__ p, double (&r)[N], const bool vec) {
  if (vec) {
    ld_row<N>(p, r);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = p[k];
  }
}
template <int N>
__device__ __forceinline__ void gst_row(double* __restrict__ p, const double (&r)[N], const bool vec) {
  if (vec) {
    st_row<N>(p, r);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) p[k] = r[k];
  }
}

// Model rows of one step as a lane needs them: row li of F, Q; entry li of bq; row l of H, R and entry l of c, y on
// the lanes l < NY (zeros elsewhere).
template <int N, int NY>
struct CoopModel {
  double F[N], Q[N], bq, H[N], R[NY], c, y;
  __device__ __forceinline__ void load_transition(const SSMArgs& a, long long seq, long long k, int li, bool vec) {
    gld_row<N>(a.F + seq * a.sF + k * a.tF + li * N, F, vec);
    gld_row<N>(a.Q + seq * a.sQ + k * a.tQ + li * N, Q, vec);
    bq = a.bq[seq * a.sb + k * a.tb + li];
  }
  __device__ __forceinline__ void load_observation(const SSMArgs& a, long long seq, long long k, int l, bool vec) {
    const bool h = l < NY;
    const int la = h ? l : 0;
    gld_row<N>(a.H + seq * a.sH + k * a.tH + la * N, H, vec);
    const double* r = a.R + seq * a.sR + k * a.tR + la * NY;
#pragma unroll
    for (int q = 0; q < NY; ++q) R[q] = h ? r[q] : 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) H[j] = h ? H[j] : 0.0;
    c = h ? a.c[seq * a.sc + k * a.tc + la] : 0.0;
  }
};

// =================================================================================================================
// K3, step loop.  cstate: per-chunk start states (mean, packed lower factor: field f of chunk c of sequence seq at
// cstate[seq * cs_stride + f * Ppad + c]) written by k_chunk_start.
// Writes the filtered trajectory at indices k + 1, the log-likelihood partial of the CTA's 32 chunks and -- only when
// the backward sweep is the per-thread kernel -- the packed filtered states it reads (fpack).
// =================================================================================================================
template <int N, int NY, bool LOGLIK>
__global__ void __launch_bounds__(kCBlock, PSQ_COOP_MINB_K3)
k_coop_filter_apply(const SSMArgs a, long long T, int K, long long Ppad, const double* __restrict__ cstate,
                    long long cs_stride, double* __restrict__ fm, double* __restrict__ fL,
                    double* __restrict__ ell_part, double* __restrict__ fpack, const int vec_i) {
  const bool vec = vec_i != 0;
  using CS = CoopSweep<N>;
  constexpr int RS = CS::RS;
  constexpr int TRI = N * (N + 1) / 2;
  constexpr int NP = N + TRI;
  extern __shared__ __align__(16) double coop_sm[];
  __shared__ double s_ell[kCChunks];
  const int g = threadIdx.x / kCG, l = threadIdx.x % kCG;
  const int gbase = (threadIdx.x & 31) - l;
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kCChunks + g;
  double* const buf = coop_sm + g * CS::SZ;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const bool rowact = l < N;
  const int r = rowact ? l : N;
  const int li = rowact ? l : 0;

  const double* cs = cstate + seq * cs_stride + c;
  double m = cs[li * Ppad];
  double Y[N];
#pragma unroll
  for (int j = 0; j < N; ++j) Y[j] = (j <= li) ? cs[(N + li * (li + 1) / 2 + j) * Ppad] : 0.0;
  double* const fmS = fm + seq * (T + 1) * N;
  double* const fLS = fL + seq * (T + 1) * N * N;
  double* const fp = fpack ? fpack + (seq * K * NP) * Ppad + c : nullptr;
  if (fp && rowact && len > 0) {   // slot 0 = the state the chunk starts from
    fp[li * Ppad] = m;
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (j <= li) fp[(N + li * (li + 1) / 2 + j) * Ppad] = Y[j];
  }
  const bool tv_t = (a.tF | a.tQ | a.tb) != 0, tv_o = (a.tH | a.tR | a.tc) != 0;
  CoopModel<N, NY> md;
  double ell = 0.0;
#pragma unroll 1
  for (int j = 0; j < K; ++j) {
    const bool act = j < len;
    const long long k = act ? k0 + j : 0;
    if (j == 0 || tv_t) md.load_transition(a, seq, k, li, vec);
    if (j == 0 || tv_o) md.load_observation(a, seq, k, l, vec);
    const double yv = (l < NY) ? a.y[seq * a.sy + k * a.ty + l] : 0.0;
    // ---- predict: mp = F m + bq, Np = tria([F Y | Q])
    if (rowact) {
      st_row<N>(buf + CS::R0 + l * RS, Y);
      buf[CS::V0 + l] = m;
    }
    __syncwarp();
    double M1[2 * N];
    double mp = md.bq;
#pragma unroll
    for (int q = 0; q < N; ++q) M1[q] = 0.0;
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double f = md.F[kk];
#pragma unroll
      for (int q = 0; q < N; ++q) M1[q] = fma(f, t[q], M1[q]);
      mp = fma(f, buf[CS::V0 + kk], mp);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) M1[N + q] = (q <= li) ? md.Q[q] : 0.0;
    __syncwarp();
    coop_house_shfl<2 * N, N, N, N>(M1, r, gbase);
    double Np[N];
#pragma unroll
    for (int q = 0; q < N; ++q) Np[q] = (q <= li) ? M1[q] : 0.0;
    if (rowact) {
      st_row<N>(buf + CS::R0 + l * RS, Np);
      buf[CS::V0 + l] = mp;
    }
    __syncwarp();
    // ---- update: rows [H Np | R] on the lanes < NY, residual y - H mp - c
    double hrow[N + NY], mrow[N + NY];
    double res = yv - md.c;
#pragma unroll
    for (int q = 0; q < N; ++q) hrow[q] = 0.0;
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double h = md.H[kk];
#pragma unroll
      for (int q = 0; q < N; ++q) hrow[q] = fma(h, t[q], hrow[q]);
      res = fma(-h, buf[CS::V0 + kk], res);
    }
#pragma unroll
    for (int q = 0; q < NY; ++q) {
      hrow[N + q] = md.R[q];
      mrow[N + q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < N; ++q) mrow[q] = Np[q];
    __syncwarp();
    coop_update_reflectors<N, NY>(hrow, mrow, l, gbase);
    double P11[NY][NY], inv[NY], rr[NY], quad, det;
    coop_psi11_solve<N, NY>(hrow, res, gbase, P11, inv, rr, quad, det);
    double mn = mp;
#pragma unroll
    for (int q = 0; q < NY; ++q) mn = fma(mrow[q], rr[q], mn);
    m = mn;
#pragma unroll
    for (int q = 0; q < N; ++q) Y[q] = mrow[NY + q];
    if (LOGLIK && act) ell += -0.5 * quad - log(fabs(det)) - NY * kHalfLog2Pi;
    // ---- the filtered state at index k + 1 leaves with a lower-triangular factor (_filtering.py:126-131)
    double Lr[N];
#pragma unroll
    for (int q = 0; q < N; ++q) Lr[q] = Y[q];
    coop_house_shfl<N, N - 1, 0, N>(Lr, r, gbase);
#pragma unroll
    for (int q = 0; q < N; ++q) Lr[q] = (q <= li) ? Lr[q] : 0.0;
    if (act && rowact) {
      fmS[(k + 1) * N + l] = m;
      gst_row<N>(fLS + ((k + 1) * N + l) * N, Lr, vec);
      if (fp && j + 1 < K) {
        double* s = fp + (long long)(j + 1) * NP * Ppad;
        s[li * Ppad] = m;
#pragma unroll
        for (int q = 0; q < N; ++q)
          if (q <= li) s[(N + li * (li + 1) / 2 + q) * Ppad] = Lr[q];
      }
    }
  }
  if (LOGLIK) {
    if (l == 0) s_ell[g] = ell;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
#pragma unroll 1
      for (int q = 0; q < kCChunks; ++q) s += s_ell[q];
      ell_part[seq * (Ppad / 32) + blockIdx.x] = s;
    }
  }
}

// =================================================================================================================
// K5, step loop.  cstate: per-chunk smoothed states at the chunk's END [B][NP][Ppad] written by k_chunk_end.  Reads
// the filtered trajectory the forward sweep wrote (fm, fL: a group's 8 lanes read 8 consecutive rows, so the API
// layout is already coalesced for this mapping and no packed copy is needed).
// =================================================================================================================
template <int N>
__global__ void __launch_bounds__(kCBlock, PSQ_COOP_MINB_K5)
k_coop_smooth_apply(const SSMArgs a, long long T, int K, long long Ppad, const double* __restrict__ cstate,
                    long long cs_stride, const double* __restrict__ fm, const double* __restrict__ fL,
                    double* __restrict__ sm, double* __restrict__ sL, const int vec_i) {
  const bool vec = vec_i != 0;
  using CS = CoopSweep<N>;
  constexpr int RS = CS::RS;
  constexpr int TRI = N * (N + 1) / 2;
  constexpr int NP = N + TRI;
  extern __shared__ __align__(16) double coop_sm[];
  const int g = threadIdx.x / kCG, l = threadIdx.x % kCG;
  const int gbase = (threadIdx.x & 31) - l;
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kCChunks + g;
  double* const buf = coop_sm + g * CS::SZ;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const bool rowact = l < N;
  const int r = rowact ? l : N;
  const int li = rowact ? l : 0;

  const double* cs = cstate + seq * cs_stride + c;
  double ms = cs[li * Ppad];
  double Ls[N];
#pragma unroll
  for (int j = 0; j < N; ++j) Ls[j] = (j <= li) ? cs[(N + li * (li + 1) / 2 + j) * Ppad] : 0.0;
  const double* const fmS = fm + seq * (T + 1) * N;
  const double* const fLS = fL + seq * (T + 1) * N * N;
  double* const smS = sm + seq * (T + 1) * N;
  double* const sLS = sL + seq * (T + 1) * N * N;
  const bool tv_t = (a.tF | a.tQ | a.tb) != 0;
  CoopModel<N, 1> md;
#pragma unroll 1
  for (int jj = 0; jj < K; ++jj) {
    const int j = K - 1 - jj;
    const bool act = j < len;
    const long long k = act ? k0 + j : 0;
    if (jj == 0 || tv_t) md.load_transition(a, seq, k, li, vec);
    double Lf[N];
    gld_row<N>(fLS + (k * N + li) * N, Lf, vec);
    const double mf = fmS[k * N + li];
    if (rowact) {
      st_row<N>(buf + CS::R0 + l * RS, Lf);
      st_row<N>(buf + CS::R1 + l * RS, Ls);
      buf[CS::V0 + l] = mf;
    }
    __syncwarp();
    // ---- tria([[F L, Q], [L, 0]]): N reflectors from the top rows                     _smoothing.py:76-81
    double top[2 * N], bot[2 * N];
    double mpf = md.bq;
#pragma unroll
    for (int q = 0; q < N; ++q) top[q] = 0.0;
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double f = md.F[kk];
#pragma unroll
      for (int q = 0; q < N; ++q) top[q] = fma(f, t[q], top[q]);
      mpf = fma(f, buf[CS::V0 + kk], mpf);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) {
      top[N + q] = (q <= li) ? md.Q[q] : 0.0;
      bot[q] = (q <= li) ? Lf[q] : 0.0;
      bot[N + q] = 0.0;
    }
    const double dlt = ms - mpf;
    __syncwarp();
    coop_house2<2 * N, N, N, N>(top, bot, r, gbase);
    {
      double P[N];
      double dg = 1.0;
#pragma unroll
      for (int q = 0; q < N; ++q) {
        P[q] = (q <= li) ? top[q] : 0.0;
        dg = (q == li) ? top[q] : dg;
      }
      if (rowact) {
        st_row<N>(buf + CS::R0 + l * RS, P);     // Phi11
        buf[CS::V0 + l] = dlt;
        buf[CS::V1 + l] = rcp_nr(dg);
      }
    }
    __syncwarp();
    // ---- E = Phi21 Phi11^{-1} (row l), right-looking back substitution           _smoothing.py:83
    double E[N];
    static_for<0, N>([&](auto kc) {
      constexpr int kk = N - 1 - decltype(kc)::value;
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      E[kk] = bot[kk] * buf[CS::V1 + kk];
#pragma unroll
      for (int q = 0; q < kk; ++q) bot[q] = fma(-E[kk], t[q], bot[q]);
    });
    // ---- m_s <- m + E (m_s - F m - b),  L_s <- tria([E L_s | Phi22])             _operators.py:118-125
    double mn = mf;
    double W[2 * N];
#pragma unroll
    for (int q = 0; q < N; ++q) W[q] = 0.0;
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R1 + kk * RS, t);
      mn = fma(E[kk], buf[CS::V0 + kk], mn);
#pragma unroll
      for (int q = 0; q < N; ++q) W[q] = fma(E[kk], t[q], W[q]);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) W[N + q] = bot[N + q];
    __syncwarp();
    coop_house_shfl<2 * N, N, 0, N>(W, r, gbase);
    ms = act ? mn : ms;
#pragma unroll
    for (int q = 0; q < N; ++q) Ls[q] = act ? ((q <= li) ? W[q] : 0.0) : Ls[q];
    if (act && rowact) {
      smS[k * N + l] = ms;
      gst_row<N>(sLS + (k * N + l) * N, Ls, vec);
    }
  }
}

// =================================================================================================================
// K1, step loop: the chunk summary (A, b, U, eta, Z) by the collapsed combine (psq::filter_reduce_step), one row
// per lane.  Writes the chunk's summary with its LAST step predict-only to chunk_own (what chunk_smoothing_total
// reads) and the full summary to `summ` (FElem order, SoA [NF][Ppad]), which k_chunk_scan_f scans in place.
// =================================================================================================================
template <int N, int NY>
__global__ void __launch_bounds__(kCBlock, PSQ_COOP_MINB_K1)
k_coop_filter_reduce(const SSMArgs a, long long T, int K, long long Ppad, double* __restrict__ chunk_own,
                     double* __restrict__ summ, const int vec_i) {
  const bool vec = vec_i != 0;
  using CS = CoopSweep<N>;
  constexpr int RS = CS::RS;
  constexpr int TRI = N * (N + 1) / 2;
  constexpr int NF = FElem<N>::NF;
  constexpr int oA = 0, ob = N * N, oU = N * N + N, oe = N * N + N + TRI, oZ = N * N + 2 * N + TRI;
  extern __shared__ __align__(16) double coop_sm[];
  const int g = threadIdx.x / kCG, l = threadIdx.x % kCG;
  const int gbase = (threadIdx.x & 31) - l;
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kCChunks + g;
  double* const buf = coop_sm + g * CS::SZ;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  const bool rowact = l < N;
  const int r = rowact ? l : N;
  const int li = rowact ? l : 0;

  double A[N], Y[N], Z[N];
  double b = 0.0, eta = 0.0;
#pragma unroll
  for (int q = 0; q < N; ++q) {
    A[q] = (q == li) ? 1.0 : 0.0;
    Y[q] = 0.0;
    Z[q] = 0.0;
  }
  double* const own = chunk_own + seq * NF * Ppad + c;
  const bool tv_t = (a.tF | a.tQ | a.tb) != 0, tv_o = (a.tH | a.tR | a.tc) != 0;
  CoopModel<N, NY> md;
#pragma unroll 1
  for (int j = 0; j < K; ++j) {
    const bool act = j < len;
    const long long k = act ? k0 + j : 0;
    if (j == 0 || tv_t) md.load_transition(a, seq, k, li, vec);
    if (j == 0 || tv_o) md.load_observation(a, seq, k, l, vec);
    const double yv = (l < NY) ? a.y[seq * a.sy + k * a.ty + l] : 0.0;
    if (rowact) {
      st_row<N>(buf + CS::R0 + l * RS, Y);
      st_row<N>(buf + CS::R1 + l * RS, A);
      buf[CS::V0 + l] = b;
    }
    if (l < NY) st_row<N>(buf + CS::HB + l * RS, md.H);
    __syncwarp();
    // ---- predict-only summary: F A, F b + bq, tria([F Y | Q])
    double M1[2 * N], FA[N];
    double mp = md.bq;
#pragma unroll
    for (int q = 0; q < N; ++q) {
      M1[q] = 0.0;
      FA[q] = 0.0;
    }
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N], u[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      ld_row<N>(buf + CS::R1 + kk * RS, u);
      const double f = md.F[kk];
#pragma unroll
      for (int q = 0; q < N; ++q) {
        M1[q] = fma(f, t[q], M1[q]);
        FA[q] = fma(f, u[q], FA[q]);
      }
      mp = fma(f, buf[CS::V0 + kk], mp);
    }
#pragma unroll
    for (int q = 0; q < N; ++q) M1[N + q] = (q <= li) ? md.Q[q] : 0.0;
    __syncwarp();
    coop_house_shfl<2 * N, N, N, N>(M1, r, gbase);
    double Np[N];
#pragma unroll
    for (int q = 0; q < N; ++q) Np[q] = (q <= li) ? M1[q] : 0.0;
    if (act && j + 1 == len && rowact) {
      own[(ob + li) * Ppad] = mp;
      own[(oe + li) * Ppad] = eta;
#pragma unroll
      for (int q = 0; q < N; ++q) {
        own[(oA + li * N + q) * Ppad] = FA[q];
        if (q <= li) {
          own[(oU + li * (li + 1) / 2 + q) * Ppad] = Np[q];
          own[(oZ + li * (li + 1) / 2 + q) * Ppad] = Z[q];
        }
      }
    }
    if (rowact) {
      st_row<N>(buf + CS::R0 + l * RS, Np);
      st_row<N>(buf + CS::R1 + l * RS, FA);
      buf[CS::V0 + l] = mp;
    }
    __syncwarp();
    // ---- update rows and the lane's COLUMN of H F A
    double hrow[N + NY], mrow[N + NY], Vc[NY];
    double res = yv - md.c;
#pragma unroll
    for (int q = 0; q < N; ++q) hrow[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NY; ++q) Vc[q] = 0.0;
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double h = md.H[kk];
#pragma unroll
      for (int q = 0; q < N; ++q) hrow[q] = fma(h, t[q], hrow[q]);
      res = fma(-h, buf[CS::V0 + kk], res);
      const double fa = buf[CS::R1 + kk * RS + li];      // (F A)[kk][l]
#pragma unroll
      for (int q = 0; q < NY; ++q) Vc[q] = fma(buf[CS::HB + q * RS + kk], fa, Vc[q]);
    }
#pragma unroll
    for (int q = 0; q < NY; ++q) {
      hrow[N + q] = md.R[q];
      mrow[N + q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < N; ++q) mrow[q] = Np[q];
    __syncwarp();
    coop_update_reflectors<N, NY>(hrow, mrow, l, gbase);
    double P11[NY][NY], inv[NY], rr[NY], quad, det;
    coop_psi11_solve<N, NY>(hrow, res, gbase, P11, inv, rr, quad, det);
    // V = Psi11^{-1} (H F A): column l
#pragma unroll
    for (int q = 0; q < NY; ++q) {
      double s = Vc[q];
#pragma unroll
      for (int p = 0; p < q; ++p) s = fma(-P11[q][p], Vc[p], s);
      Vc[q] = s * inv[q];
    }
    if (rowact) {
#pragma unroll
      for (int q = 0; q < NY; ++q) buf[CS::R0 + q * RS + l] = Vc[q];
    }
    __syncwarp();
    // A <- F A - Psi21 V ; b <- mp + Psi21 rr ; eta <- eta + V^T rr ; Y <- posterior factor ; Z <- tria([Z | V^T])
    double An[N];
#pragma unroll
    for (int q = 0; q < N; ++q) An[q] = FA[q];
    double bn = mp, en = eta;
#pragma unroll
    for (int p = 0; p < NY; ++p) {
      double t[N];
      ld_row<N>(buf + CS::R0 + p * RS, t);
#pragma unroll
      for (int q = 0; q < N; ++q) An[q] = fma(-mrow[p], t[q], An[q]);
      bn = fma(mrow[p], rr[p], bn);
      en = fma(Vc[p], rr[p], en);
    }
    double Zn[N], Wv[NY];
#pragma unroll
    for (int q = 0; q < N; ++q) Zn[q] = Z[q];
#pragma unroll
    for (int q = 0; q < NY; ++q) Wv[q] = Vc[q];
    coop_tria_append<N, NY>(Zn, Wv, r, gbase);
    b = act ? bn : b;
    eta = act ? en : eta;
#pragma unroll
    for (int q = 0; q < N; ++q) {
      A[q] = act ? An[q] : A[q];
      Y[q] = act ? mrow[NY + q] : Y[q];
      Z[q] = act ? ((q <= li) ? Zn[q] : 0.0) : Z[q];
    }
    __syncwarp();
  }
  // U = tria(Y); the full summary in FElem order
  coop_house_shfl<N, N - 1, 0, N>(Y, r, gbase);
  if (rowact) {
    double* s = summ + seq * NF * Ppad + c;
    s[(ob + li) * Ppad] = b;
    s[(oe + li) * Ppad] = eta;
#pragma unroll
    for (int q = 0; q < N; ++q) {
      s[(oA + li * N + q) * Ppad] = A[q];
      if (q <= li) {
        s[(oU + li * (li + 1) / 2 + q) * Ppad] = Y[q];
        s[(oZ + li * (li + 1) / 2 + q) * Ppad] = Z[q];
      }
    }
  }
}

// =================================================================================================================
// Once-per-chunk stages around the sub-warp step loops: per-thread code (one thread per chunk, one warp per CTA so
// that the 32-chunk scan unit is the warp) -- the tails / prologues of k_filter_reduce, k_filter_apply and
// k_smooth_apply, reading and writing the same scratch.
// =================================================================================================================
// Exclusive scan over the 32 items of every scan unit (= the chunks of one CTA of the step loops), in place, unit
// totals to unit_tot[B][NF][Ppad / 32] (time order): the warp-level scan of the per-thread sweeps, done with the
// sub-warp combines of psqrt_coop2.cuh (one lane group per item, Kogge-Stone through shared memory, 5 dependent
// combines).  REV: suffix scan (item 31 of the unit comes first in scan order).  One CTA per unit.
template <class OP, int NF>
constexpr size_t unit_scan_smem_bytes() {
  return sizeof(double) * (size_t)(2 * 32 * OP::NFD + 32 * OP::WS) + sizeof(int) * NF;
}
template <class OP, int NF, bool REV>
__global__ void __launch_bounds__(32 * OP::G, 1)
k_unit_scan(double* __restrict__ items, long long Ppad, double* __restrict__ unit_tot,
            unsigned int* __restrict__ counter, unsigned int* __restrict__ fuse_ctr) {
  constexpr int G = OP::G;
  constexpr int NFD = OP::NFD;
  constexpr int IT = 32;
  extern __shared__ __align__(16) double unit_sm[];
  double* const slots = unit_sm;                        // [2][IT][NFD]
  double* const wsall = unit_sm + 2 * IT * NFD;         // [IT][WS]
  int* const dmap = reinterpret_cast<int*>(wsall + IT * OP::WS);
  const long long seq = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int l = threadIdx.x % G;
  const int gbase = lane - l;
  const int x = threadIdx.x / G;                        // scan position within the unit
  double* const ws = wsall + x * OP::WS;
  auto slot = [&](int s, int xx) { return slots + (s * IT + xx) * NFD; };
  if (blockIdx.x == 0 && threadIdx.x == 0 && counter) {   // arms the ticket of the mid-level scan that follows
    counter[seq] = 0u;
    if (fuse_ctr) {
      fuse_ctr[2 * seq] = 0u;
      fuse_ctr[2 * seq + 1] = 0u;
    }
  }
  for (int f = threadIdx.x; f < NF; f += blockDim.x) dmap[f] = OP::dense_of(f);
  for (int k = threadIdx.x; k < 2 * IT * NFD; k += blockDim.x) slots[k] = 0.0;   // upper triangles stay zero
  __syncthreads();
  double* const base = items + seq * NF * Ppad;
  const long long gi = (long long)blockIdx.x * IT + (REV ? IT - 1 - x : x);
  {
    constexpr int PER = (NF + G - 1) / G;
    double v[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int f = l + q * G;
      v[q] = (f < NF) ? __ldcg(base + f * Ppad + gi) : 0.0;
    }
    double* d = slot(0, x);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int f = l + q * G;
      if (f < NF) d[dmap[f]] = v[q];
    }
  }
  __syncthreads();
  int cur = 0;
#pragma unroll 1
  for (int lev = 0; lev < 5; ++lev) {
    const int d = 1 << lev;
    const bool keep = x < d;
    const double* e1 = slot(cur, keep ? x : x - d);
    const double* e2 = slot(cur, x);
    double* o = slot(cur ^ 1, x);
    OP::combine(e1, e2, o, ws, l, gbase);
    if (keep) copy_slot<NFD, G>(o, e2, l);
    __syncthreads();
    cur ^= 1;
  }
  {
    const double* sp = slot(cur, x > 0 ? x - 1 : 0);
    for (int f = l; f < NF; f += G) {
      const int off = dmap[f];
      base[f * Ppad + gi] = (x == 0) ? OP::ident(off) : sp[off];
    }
  }
  if (x == IT - 1) {
    const double* sp = slot(cur, x);
    double* t = unit_tot + seq * NF * (Ppad / IT) + blockIdx.x;
    for (int f = l; f < NF; f += G) t[f * (Ppad / IT)] = sp[dmap[f]];
  }
}

// Carry pushed through the three exclusive prefixes -> the state every chunk starts from (cstate, aliasing the
// first N + TRI fields of chunk_own column c, which this thread has consumed by then); SMOOTH: the chunk's smoothing
// total (prologue of k_filter_apply without its warp scan, which k_unit_scan does).
template <int N, bool SMOOTH>
__global__ void __launch_bounds__(32)
k_chunk_start(long long T, int K, long long Ppad, const double* __restrict__ carry_m,
              const double* __restrict__ carry_L, double* chunk_own, const double* __restrict__ chunk_pref,
              const double* __restrict__ warp_pref, const double* __restrict__ group_pref, double* __restrict__ fm,
              double* __restrict__ fL, double* __restrict__ chunk_suf, unsigned int* __restrict__ counter) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
  const long long Mw = Ppad / 32;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  if (SMOOTH && c == 0) counter[seq] = 0u;
  Gauss<N> x;
  load_gauss_dense<N>(carry_m + seq * N, carry_L + seq * N * N, x);
#pragma unroll 1
  for (int lvl = 0; lvl < 3; ++lvl) {
    constexpr int IT = MidCfg<N>::IT;
    const double* buf = (lvl == 0) ? group_pref : (lvl == 1) ? warp_pref : chunk_pref;
    const long long n_items = (lvl == 0) ? (Mw + IT - 1) / IT : (lvl == 1) ? Mw : Ppad;
    const long long idx = (lvl == 0) ? c / (32 * IT) : (lvl == 1) ? c / 32 : c;
    FElem<N> e;
    soa_load(buf, seq, n_items, idx, e);
    filtering_apply<N>(x, e);
  }
  if (c == 0) store_gauss_dense<N>(fm + seq * (T + 1) * N, fL + seq * (T + 1) * N * N, x);
  if (SMOOTH) {
    SElem<N> sacc;
    sacc.set_identity();
    if (k0 < k1) {
      FElem<N> e;
      soa_load(chunk_own, seq, Ppad, c, e);
      chunk_smoothing_total<N>(x, e, sacc);
    }
    soa_store(chunk_suf, seq, Ppad, c, sacc);   // k_unit_scan<CoopS2, ., true> turns the totals into suffixes
  }
  double* cs = chunk_own + seq * FElem<N>::NF * Ppad + c;   // cs_stride of the step loop = FElem<N>::NF * Ppad
#pragma unroll
  for (int f = 0; f < N; ++f) cs[f * Ppad] = x.m[f];
#pragma unroll
  for (int f = 0; f < Gauss<N>::TRI; ++f) cs[(N + f) * Ppad] = x.L[f];
}

// Terminal state pushed through the three exclusive suffixes -> the smoothed state at every chunk's end (cstate,
// aliasing the first N + TRI fields of chunk_suf column c) (prologue of k_smooth_apply).
template <int N>
__global__ void __launch_bounds__(32)
k_chunk_end(long long T, int K, long long Ppad, const double* __restrict__ carry_m,
            const double* __restrict__ carry_L, long long carry_mstride, long long carry_Lstride, double* chunk_suf,
            const double* __restrict__ warp_suf, const double* __restrict__ group_suf, double* __restrict__ sm,
            double* __restrict__ sL, int write_terminal) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * 32 + threadIdx.x;
  const long long Mw = Ppad / 32;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  if (k1 <= k0) return;
  Gauss<N> xs;
  load_gauss_dense<N>(carry_m + seq * carry_mstride, carry_L + seq * carry_Lstride, xs);
  if (write_terminal && k1 == T)
    store_gauss_dense<N>(sm + seq * (T + 1) * N + T * N, sL + seq * (T + 1) * N * N + T * N * N, xs);
#pragma unroll 1
  for (int lvl = 0; lvl < 3; ++lvl) {
    constexpr int IT = MidCfg<N>::IT;
    const double* buf = (lvl == 0) ? group_suf : (lvl == 1) ? warp_suf : chunk_suf;
    const long long n_items = (lvl == 0) ? (Mw + IT - 1) / IT : (lvl == 1) ? Mw : Ppad;
    const long long idx = (lvl == 0) ? (Mw - 1 - c / 32) / IT : (lvl == 1) ? c / 32 : c;
    SElem<N> e;
    soa_load(buf, seq, n_items, idx, e);
    smoothing_apply<N>(xs, e);
  }
  double* cs = chunk_suf + seq * SElem<N>::NF * Ppad + c;
#pragma unroll
  for (int f = 0; f < N; ++f) cs[f * Ppad] = xs.m[f];
#pragma unroll
  for (int f = 0; f < Gauss<N>::TRI; ++f) cs[(N + f) * Ppad] = xs.L[f];
}

}  // namespace psq
