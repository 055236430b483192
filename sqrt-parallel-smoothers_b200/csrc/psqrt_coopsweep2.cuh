// psqrt_coopsweep2.cuh -- the sub-warp sweeps of psqrt_coopsweep.cuh with R = N / G matrix rows per lane (G lanes per
// chunk; nx = 8: G = 4, R = 2).
//
// With one row per lane every reflector pays its scalar chain (norm, rsqrt, rcp) and its pivot-row broadcast for ONE
// row update per lane.  With R rows per lane the same shuffles and the same scalar chain serve R independent row
// updates: fewer instructions per chunk (a warp carries 32 / G chunks) and R independent DFMA streams per lane for the
// in-order issue to overlap.  Row r of a matrix lives on lane r % G, slot r / G, so that for a fixed slot consecutive
// lanes hold consecutive rows (the trajectory rows of a group stay contiguous in memory).
// Same formulas, same scratch, same once-per-chunk kernels as psqrt_coopsweep.cuh.
#pragma once
#include "psqrt_coopsweep.cuh"

namespace psq {

// Householder triangularisation from the right, rows distributed (lane r % G, slot r / G); rows [0, NREFL) become
// lower-trapezoidal, TRIBLK as in house_rows.  Every lane of the warp must call.
template <int C, int NREFL, int TRIBLK, int G, int R>
__device__ __forceinline__ void coop_house_rows(double (&row)[R][C], const int l, const int gbase) {
  static_for<0, NREFL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int kend = (TRIBLK > 0) ? ((TRIBLK + j + 1 < C) ? TRIBLK + j + 1 : C) : C;
    constexpr int js = j / G, jl = j % G;
    if constexpr (j + 1 < kend) {
      double p[C];
#pragma unroll
      for (int k = j; k < kend; ++k) p[k] = __shfl_sync(0xffffffffu, row[js][k], gbase + jl);
      const double alpha = p[j];
      double sigma = 0.0, sigma2 = 0.0;
#pragma unroll
      for (int k = j + 1; k < kend; k += 2) {
        sigma = fma(p[k], p[k], sigma);
        if (k + 1 < kend) sigma2 = fma(p[k + 1], p[k + 1], sigma2);
      }
      sigma += sigma2;
      double d[R];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        double d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int k = j + 1; k < kend; k += 2) {
          d1 = fma(row[s][k], p[k], d1);
          if (k + 1 < kend) d2 = fma(row[s][k + 1], p[k + 1], d2);
        }
        d[s] = d1 + d2;
      }
      const double q = fma(alpha, alpha, sigma);
      const double mask = (q != 0.0) ? 1.0 : 0.0;
      const double qs = (q != 0.0) ? q : 1.0;
      const double norm = qs * rsqrt_nr(qs);
      const double beta = -copysign(norm, alpha) * mask;
      const double v0 = alpha - beta;
      const double sc = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int rs = l + s * G;
        const double dv = fma(row[s][j], v0, d[s]) * sc;
        const double dd = (rs > j) ? dv : 0.0;
        row[s][j] = (rs == j) ? beta : fma(-dd, v0, row[s][j]);
#pragma unroll
        for (int k = j + 1; k < kend; ++k) row[s][k] = fma(-dd, p[k], row[s][k]);
      }
    }
  });
}

// the same with a second row set `bot` that every reflector updates (tria([[F L, Q], [L, 0]]) of the RTS step)
template <int C, int NREFL, int TRIBLK, int G, int R>
__device__ __forceinline__ void coop_house2_rows(double (&top)[R][C], double (&bot)[R][C], const int l,
                                                 const int gbase) {
  static_for<0, NREFL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int kend = (TRIBLK > 0) ? ((TRIBLK + j + 1 < C) ? TRIBLK + j + 1 : C) : C;
    constexpr int js = j / G, jl = j % G;
    if constexpr (j + 1 < kend) {
      double p[C];
#pragma unroll
      for (int k = j; k < kend; ++k) p[k] = __shfl_sync(0xffffffffu, top[js][k], gbase + jl);
      const double alpha = p[j];
      double sigma = 0.0, sigma2 = 0.0;
#pragma unroll
      for (int k = j + 1; k < kend; k += 2) {
        sigma = fma(p[k], p[k], sigma);
        if (k + 1 < kend) sigma2 = fma(p[k + 1], p[k + 1], sigma2);
      }
      sigma += sigma2;
      double dt[R], db[R];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        double t1 = 0.0, b1 = 0.0;
#pragma unroll
        for (int k = j + 1; k < kend; ++k) {
          t1 = fma(top[s][k], p[k], t1);
          b1 = fma(bot[s][k], p[k], b1);
        }
        dt[s] = t1;
        db[s] = b1;
      }
      const double q = fma(alpha, alpha, sigma);
      const double mask = (q != 0.0) ? 1.0 : 0.0;
      const double qs = (q != 0.0) ? q : 1.0;
      const double norm = qs * rsqrt_nr(qs);
      const double beta = -copysign(norm, alpha) * mask;
      const double v0 = alpha - beta;
      const double sc = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int rs = l + s * G;
        const double tv = fma(top[s][j], v0, dt[s]) * sc;
        const double bv = fma(bot[s][j], v0, db[s]) * sc;
        const double dd = (rs > j) ? tv : 0.0;
        top[s][j] = (rs == j) ? beta : fma(-dd, v0, top[s][j]);
        bot[s][j] = fma(-bv, v0, bot[s][j]);
#pragma unroll
        for (int k = j + 1; k < kend; ++k) {
          top[s][k] = fma(-dd, p[k], top[s][k]);
          bot[s][k] = fma(-bv, p[k], bot[s][k]);
        }
      }
    }
  });
}

// Z <- tria([Z | W]), rows distributed as above (coop_tria_append with R rows per lane)
template <int N, int K, int G, int R>
__device__ __forceinline__ void coop_tria_append_rows(double (&z)[R][N], double (&w)[R][K], const int l,
                                                      const int gbase) {
  static_for<0, N>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int js = j / G, jl = j % G;
    const double alpha = __shfl_sync(0xffffffffu, z[js][j], gbase + jl);
    double pw[K];
    double sigma = 0.0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      pw[a] = __shfl_sync(0xffffffffu, w[js][a], gbase + jl);
      sigma = fma(pw[a], pw[a], sigma);
    }
    const double q = fma(alpha, alpha, sigma);
    const double mask = (q != 0.0) ? 1.0 : 0.0;
    const double qs = (q != 0.0) ? q : 1.0;
    const double norm = qs * rsqrt_nr(qs);
    const double beta = -copysign(norm, alpha) * mask;
    const double v0 = alpha - beta;
    const double sc = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      double d = 0.0;
#pragma unroll
      for (int a = 0; a < K; ++a) d = fma(w[s][a], pw[a], d);
      d = fma(z[s][j], v0, d) * sc;
      const double dd = (rs > j) ? d : 0.0;
      z[s][j] = (rs == j) ? beta : fma(-dd, v0, z[s][j]);
#pragma unroll
      for (int a = 0; a < K; ++a) w[s][a] = fma(-dd, pw[a], w[s][a]);
    }
  });
}

// the NY reflectors of the measurement update: pivot rows hrow on lanes 0 .. NY-1 (NY <= G), R bottom rows per lane
template <int N, int NY, int G, int R>
__device__ __forceinline__ void coop_update_reflectors_rows(double (&hrow)[N + NY], double (&mrow)[R][N + NY],
                                                            const int l, const int gbase) {
  static_assert(NY <= G, "the pivot rows of the update sit one per lane");
  constexpr int C = N + NY;
  static_for<0, NY>([&](auto ac) {
    constexpr int a = decltype(ac)::value;
    double p[C];
#pragma unroll
    for (int k = a; k < C; ++k) p[k] = __shfl_sync(0xffffffffu, hrow[k], gbase + a);
    const double alpha = p[a];
    double sigma = 0.0, sigma2 = 0.0, dh = 0.0, dh2 = 0.0;
#pragma unroll
    for (int k = a + 1; k < C; k += 2) {
      sigma = fma(p[k], p[k], sigma);
      dh = fma(hrow[k], p[k], dh);
      if (k + 1 < C) {
        sigma2 = fma(p[k + 1], p[k + 1], sigma2);
        dh2 = fma(hrow[k + 1], p[k + 1], dh2);
      }
    }
    sigma += sigma2;
    dh += dh2;
    double dm[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      double d1 = 0.0, d2 = 0.0;
#pragma unroll
      for (int k = a + 1; k < C; k += 2) {
        d1 = fma(mrow[s][k], p[k], d1);
        if (k + 1 < C) d2 = fma(mrow[s][k + 1], p[k + 1], d2);
      }
      dm[s] = d1 + d2;
    }
    const double q = fma(alpha, alpha, sigma);
    const double mask = (q != 0.0) ? 1.0 : 0.0;
    const double qs = (q != 0.0) ? q : 1.0;
    const double norm = qs * rsqrt_nr(qs);
    const double beta = -copysign(norm, alpha) * mask;
    const double v0 = alpha - beta;
    const double sc = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
    dh = fma(hrow[a], v0, dh) * sc;
    const double ddh = (l > a && l < NY) ? dh : 0.0;
    hrow[a] = (l == a) ? beta : fma(-ddh, v0, hrow[a]);
#pragma unroll
    for (int k = a + 1; k < C; ++k) hrow[k] = fma(-ddh, p[k], hrow[k]);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const double dv = fma(mrow[s][a], v0, dm[s]) * sc;
      mrow[s][a] = fma(-dv, v0, mrow[s][a]);
#pragma unroll
      for (int k = a + 1; k < C; ++k) mrow[s][k] = fma(-dv, p[k], mrow[s][k]);
    }
  });
}

// Model rows of one step: rows l + s G of F, Q, bq; row l of H, R and entry l of c on the lanes l < NY
template <int N, int NY, int G, int R>
struct CoopModelR {
  double F[R][N], Q[R][N], bq[R], H[N], Rn[NY], c;
  __device__ __forceinline__ void load_transition(const SSMArgs& a, long long seq, long long k, int l, bool vec) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      gld_row<N>(a.F + seq * a.sF + k * a.tF + rs * N, F[s], vec);
      gld_row<N>(a.Q + seq * a.sQ + k * a.tQ + rs * N, Q[s], vec);
      bq[s] = a.bq[seq * a.sb + k * a.tb + rs];
    }
  }
  __device__ __forceinline__ void load_observation(const SSMArgs& a, long long seq, long long k, int l, bool vec) {
    const bool h = l < NY;
    const int la = h ? l : 0;
    gld_row<N>(a.H + seq * a.sH + k * a.tH + la * N, H, vec);
    const double* r = a.R + seq * a.sR + k * a.tR + la * NY;
#pragma unroll
    for (int q = 0; q < NY; ++q) Rn[q] = h ? r[q] : 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) H[j] = h ? H[j] : 0.0;
    c = h ? a.c[seq * a.sc + k * a.tc + la] : 0.0;
  }
};

// K1 only: where the model of a step is read from.  A time-invariant part is staged ONCE per CTA in a shared-memory
// bank; a time-varying part is read from global memory at its step.  Either way the rows are loaded where they are
// used -- nothing of the model stays in registers across the step loop, which K1 (five state matrices per lane) needs
// for its triangularisations: stack 600 -> 112 bytes, 1.12 -> 1.03 ms at T = 1e6.  K3 and K5 keep the model rows in
// registers (CoopModelR): loading at the point of use made them 7 - 10 % slower.
template <int N, int NY>
struct CoopBank {
  static constexpr int bF = 0, bQ = N * N, bb = 2 * N * N, bH = 2 * N * N + N, bR = bH + NY * N, bc = bR + NY * NY;
  static constexpr int SZ = (bc + NY + 1) & ~1;
  const double *F, *Q, *bq, *H, *R, *c;   // this step's model (generic addresses: bank or global)
  bool vt, vo;                            // 16-byte loads allowed for the transition / observation rows
  // all threads of the CTA; followed by __syncthreads() in the caller
  static __device__ __forceinline__ void fill(double* bank, const SSMArgs& a, long long seq, bool tv_t, bool tv_o,
                                              bool obs) {
    if (!tv_t) {
      for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
        bank[bF + i] = a.F[seq * a.sF + i];
        bank[bQ + i] = a.Q[seq * a.sQ + i];
      }
      for (int i = threadIdx.x; i < N; i += blockDim.x) bank[bb + i] = a.bq[seq * a.sb + i];
    }
    if (obs && !tv_o) {
      for (int i = threadIdx.x; i < NY * N; i += blockDim.x) bank[bH + i] = a.H[seq * a.sH + i];
      for (int i = threadIdx.x; i < NY * NY; i += blockDim.x) bank[bR + i] = a.R[seq * a.sR + i];
      for (int i = threadIdx.x; i < NY; i += blockDim.x) bank[bc + i] = a.c[seq * a.sc + i];
    }
  }
  __device__ __forceinline__ void at(const double* bank, const SSMArgs& a, long long seq, long long k, bool tv_t,
                                     bool tv_o, bool obs, bool vec) {
    F = tv_t ? a.F + seq * a.sF + k * a.tF : bank + bF;
    Q = tv_t ? a.Q + seq * a.sQ + k * a.tQ : bank + bQ;
    bq = tv_t ? a.bq + seq * a.sb + k * a.tb : bank + bb;
    vt = tv_t ? vec : true;
    if (obs) {
      H = tv_o ? a.H + seq * a.sH + k * a.tH : bank + bH;
      R = tv_o ? a.R + seq * a.sR + k * a.tR : bank + bR;
      c = tv_o ? a.c + seq * a.sc + k * a.tc : bank + bc;
      vo = tv_o ? vec : (NY * N % 2 == 0);
    }
  }
};

// dynamic shared memory of K1: the 32 group buffers + the model bank
template <int N, int NY>
constexpr size_t coopr_k1_smem_bytes() {
  return CoopSweep<N>::smem_bytes() + sizeof(double) * CoopBank<N, NY>::SZ;
}

#ifndef PSQ_COOPR_MINB_K1
#define PSQ_COOPR_MINB_K1 2
#endif
#ifndef PSQ_COOPR_MINB_K3
#define PSQ_COOPR_MINB_K3 2
#endif
#ifndef PSQ_COOPR_MINB_K5
#define PSQ_COOPR_MINB_K5 2
#endif

// =================================================================================================================
// K3, step loop (k_coop_filter_apply with R rows per lane).  CTA = 32 chunks x G lanes.
// =================================================================================================================
template <int N, int NY, int G, bool LOGLIK>
__global__ void __launch_bounds__(kCChunks * G, PSQ_COOPR_MINB_K3)
k_coopr_filter_apply(const SSMArgs a, long long T, int K, long long Ppad, const double* __restrict__ cstate,
                     long long cs_stride, double* __restrict__ fm, double* __restrict__ fL,
                     double* __restrict__ ell_part, double* __restrict__ fpack, const int vec_i) {
  static_assert(N % G == 0, "rows per lane");
  constexpr int R = N / G;
  const bool vec = vec_i != 0;
  using CS = CoopSweep<N>;
  constexpr int RS = CS::RS;
  constexpr int TRI = N * (N + 1) / 2;
  constexpr int NP = N + TRI;
  extern __shared__ __align__(16) double coop_sm[];
  __shared__ double s_ell[kCChunks];
  const int g = threadIdx.x / G, l = threadIdx.x % G;
  const int gbase = (threadIdx.x & 31) - l;
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kCChunks + g;
  double* const buf = coop_sm + g * CS::SZ;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;

  const double* cs = cstate + seq * cs_stride + c;
  double m[R], Y[R][N];
#pragma unroll
  for (int s = 0; s < R; ++s) {
    const int rs = l + s * G;
    m[s] = cs[rs * Ppad];
#pragma unroll
    for (int j = 0; j < N; ++j) Y[s][j] = (j <= rs) ? cs[(N + rs * (rs + 1) / 2 + j) * Ppad] : 0.0;
  }
  double* const fmS = fm + seq * (T + 1) * N;
  double* const fLS = fL + seq * (T + 1) * N * N;
  double* const fp = fpack ? fpack + (seq * K * NP) * Ppad + c : nullptr;
  if (fp && len > 0) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      fp[rs * Ppad] = m[s];
#pragma unroll
      for (int j = 0; j < N; ++j)
        if (j <= rs) fp[(N + rs * (rs + 1) / 2 + j) * Ppad] = Y[s][j];
    }
  }
  const bool tv_t = (a.tF | a.tQ | a.tb) != 0, tv_o = (a.tH | a.tR | a.tc) != 0;
  CoopModelR<N, NY, G, R> md;
  double ell = 0.0;
#pragma unroll 1
  for (int j = 0; j < K; ++j) {
    const bool act = j < len;
    const long long k = act ? k0 + j : 0;
    if (j == 0 || tv_t) md.load_transition(a, seq, k, l, vec);
    if (j == 0 || tv_o) md.load_observation(a, seq, k, l, vec);
    const double yv = (l < NY) ? a.y[seq * a.sy + k * a.ty + l] : 0.0;
    // ---- predict
#pragma unroll
    for (int s = 0; s < R; ++s) {
      st_row<N>(buf + CS::R0 + (l + s * G) * RS, Y[s]);
      buf[CS::V0 + l + s * G] = m[s];
    }
    __syncwarp();
    double M1[R][2 * N], mp[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mp[s] = md.bq[s];
#pragma unroll
      for (int q = 0; q < N; ++q) M1[s][q] = 0.0;
    }
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double mk = buf[CS::V0 + kk];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const double f = md.F[s][kk];
#pragma unroll
        for (int q = 0; q < N; ++q) M1[s][q] = fma(f, t[q], M1[s][q]);
        mp[s] = fma(f, mk, mp[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
      for (int q = 0; q < N; ++q) M1[s][N + q] = (q <= l + s * G) ? md.Q[s][q] : 0.0;
    __syncwarp();
    coop_house_rows<2 * N, N, N, G, R>(M1, l, gbase);
    double Np[R][N];
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int q = 0; q < N; ++q) Np[s][q] = (q <= l + s * G) ? M1[s][q] : 0.0;
      st_row<N>(buf + CS::R0 + (l + s * G) * RS, Np[s]);
      buf[CS::V0 + l + s * G] = mp[s];
    }
    __syncwarp();
    // ---- update
    double hrow[N + NY], mrow[R][N + NY];
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
    for (int q = 0; q < NY; ++q) hrow[N + q] = md.Rn[q];
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int q = 0; q < N; ++q) mrow[s][q] = Np[s][q];
#pragma unroll
      for (int q = 0; q < NY; ++q) mrow[s][N + q] = 0.0;
    }
    __syncwarp();
    coop_update_reflectors_rows<N, NY, G, R>(hrow, mrow, l, gbase);
    double P11[NY][NY], inv[NY], rr[NY], quad, det;
    coop_psi11_solve<N, NY>(hrow, res, gbase, P11, inv, rr, quad, det);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      double mn = mp[s];
#pragma unroll
      for (int q = 0; q < NY; ++q) mn = fma(mrow[s][q], rr[q], mn);
      m[s] = mn;
#pragma unroll
      for (int q = 0; q < N; ++q) Y[s][q] = mrow[s][NY + q];
    }
    if (LOGLIK && act) ell += -0.5 * quad - log(fabs(det)) - NY * kHalfLog2Pi;
    // ---- the filtered state at index k + 1 leaves with a lower-triangular factor
    double Lr[R][N];
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
      for (int q = 0; q < N; ++q) Lr[s][q] = Y[s][q];
    coop_house_rows<N, N - 1, 0, G, R>(Lr, l, gbase);
    if (act) {
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int rs = l + s * G;
#pragma unroll
        for (int q = 0; q < N; ++q) Lr[s][q] = (q <= rs) ? Lr[s][q] : 0.0;
        fmS[(k + 1) * N + rs] = m[s];
        gst_row<N>(fLS + ((k + 1) * N + rs) * N, Lr[s], vec);
        if (fp && j + 1 < K) {
          double* sp = fp + (long long)(j + 1) * NP * Ppad;
          sp[rs * Ppad] = m[s];
#pragma unroll
          for (int q = 0; q < N; ++q)
            if (q <= rs) sp[(N + rs * (rs + 1) / 2 + q) * Ppad] = Lr[s][q];
        }
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
// K5, step loop (k_coop_smooth_apply with R rows per lane)
// =================================================================================================================
template <int N, int G>
__global__ void __launch_bounds__(kCChunks * G, PSQ_COOPR_MINB_K5)
k_coopr_smooth_apply(const SSMArgs a, long long T, int K, long long Ppad, const double* __restrict__ cstate,
                     long long cs_stride, const double* __restrict__ fm, const double* __restrict__ fL,
                     double* __restrict__ sm, double* __restrict__ sL, const int vec_i) {
  static_assert(N % G == 0, "rows per lane");
  constexpr int R = N / G;
  const bool vec = vec_i != 0;
  using CS = CoopSweep<N>;
  constexpr int RS = CS::RS;
  extern __shared__ __align__(16) double coop_sm[];
  const int g = threadIdx.x / G, l = threadIdx.x % G;
  const int gbase = (threadIdx.x & 31) - l;
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kCChunks + g;
  double* const buf = coop_sm + g * CS::SZ;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;

  const double* cs = cstate + seq * cs_stride + c;
  double ms[R], Ls[R][N];
#pragma unroll
  for (int s = 0; s < R; ++s) {
    const int rs = l + s * G;
    ms[s] = cs[rs * Ppad];
#pragma unroll
    for (int j = 0; j < N; ++j) Ls[s][j] = (j <= rs) ? cs[(N + rs * (rs + 1) / 2 + j) * Ppad] : 0.0;
  }
  const double* const fmS = fm + seq * (T + 1) * N;
  const double* const fLS = fL + seq * (T + 1) * N * N;
  double* const smS = sm + seq * (T + 1) * N;
  double* const sLS = sL + seq * (T + 1) * N * N;
  const bool tv_t = (a.tF | a.tQ | a.tb) != 0;
  CoopModelR<N, 1, G, R> md;
#pragma unroll 1
  for (int jj = 0; jj < K; ++jj) {
    const int j = K - 1 - jj;
    const bool act = j < len;
    const long long k = act ? k0 + j : 0;
    if (jj == 0 || tv_t) md.load_transition(a, seq, k, l, vec);
    double Lf[R][N], mf[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      gld_row<N>(fLS + (k * N + rs) * N, Lf[s], vec);
      mf[s] = fmS[k * N + rs];
      st_row<N>(buf + CS::R0 + rs * RS, Lf[s]);
      st_row<N>(buf + CS::R1 + rs * RS, Ls[s]);
      buf[CS::V0 + rs] = mf[s];
    }
    __syncwarp();
    double top[R][2 * N], bot[R][2 * N], mpf[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mpf[s] = md.bq[s];
#pragma unroll
      for (int q = 0; q < N; ++q) top[s][q] = 0.0;
    }
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double mk = buf[CS::V0 + kk];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const double f = md.F[s][kk];
#pragma unroll
        for (int q = 0; q < N; ++q) top[s][q] = fma(f, t[q], top[s][q]);
        mpf[s] = fma(f, mk, mpf[s]);
      }
    }
    double dlt[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
#pragma unroll
      for (int q = 0; q < N; ++q) {
        top[s][N + q] = (q <= rs) ? md.Q[s][q] : 0.0;
        bot[s][q] = (q <= rs) ? Lf[s][q] : 0.0;
        bot[s][N + q] = 0.0;
      }
      dlt[s] = ms[s] - mpf[s];
    }
    __syncwarp();
    coop_house2_rows<2 * N, N, N, G, R>(top, bot, l, gbase);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      double P[N];
      double dg = 1.0;
#pragma unroll
      for (int q = 0; q < N; ++q) {
        P[q] = (q <= rs) ? top[s][q] : 0.0;
        dg = (q == rs) ? top[s][q] : dg;
      }
      st_row<N>(buf + CS::R0 + rs * RS, P);     // Phi11
      buf[CS::V0 + rs] = dlt[s];
      buf[CS::V1 + rs] = rcp_nr(dg);
    }
    __syncwarp();
    // ---- E = Phi21 Phi11^{-1}, right-looking back substitution (R rows per lane)
    double E[R][N];
    static_for<0, N>([&](auto kc) {
      constexpr int kk = N - 1 - decltype(kc)::value;
      double t[N];
      ld_row<N>(buf + CS::R0 + kk * RS, t);
      const double iv = buf[CS::V1 + kk];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        E[s][kk] = bot[s][kk] * iv;
#pragma unroll
        for (int q = 0; q < kk; ++q) bot[s][q] = fma(-E[s][kk], t[q], bot[s][q]);
      }
    });
    double mn[R], W[R][2 * N];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mn[s] = mf[s];
#pragma unroll
      for (int q = 0; q < N; ++q) W[s][q] = 0.0;
    }
#pragma unroll
    for (int kk = 0; kk < N; ++kk) {
      double t[N];
      ld_row<N>(buf + CS::R1 + kk * RS, t);
      const double dk = buf[CS::V0 + kk];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        mn[s] = fma(E[s][kk], dk, mn[s]);
#pragma unroll
        for (int q = 0; q < N; ++q) W[s][q] = fma(E[s][kk], t[q], W[s][q]);
      }
    }
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
      for (int q = 0; q < N; ++q) W[s][N + q] = bot[s][N + q];
    __syncwarp();
    coop_house_rows<2 * N, N, 0, G, R>(W, l, gbase);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      ms[s] = act ? mn[s] : ms[s];
#pragma unroll
      for (int q = 0; q < N; ++q) Ls[s][q] = act ? ((q <= rs) ? W[s][q] : 0.0) : Ls[s][q];
      if (act) {
        smS[k * N + rs] = ms[s];
        gst_row<N>(sLS + (k * N + rs) * N, Ls[s], vec);
      }
    }
  }
}

// =================================================================================================================
// K1, step loop (k_coop_filter_reduce with R rows per lane)
// =================================================================================================================
template <int N, int NY, int G>
__global__ void __launch_bounds__(kCChunks * G, PSQ_COOPR_MINB_K1)
k_coopr_filter_reduce(const SSMArgs a, long long T, int K, long long Ppad, double* __restrict__ chunk_own,
                      double* __restrict__ summ, const int vec_i) {
  static_assert(N % G == 0, "rows per lane");
  constexpr int R = N / G;
  const bool vec = vec_i != 0;
  using CS = CoopSweep<N>;
  constexpr int RS = CS::RS;
  constexpr int TRI = N * (N + 1) / 2;
  constexpr int NF = FElem<N>::NF;
  constexpr int oA = 0, ob = N * N, oU = N * N + N, oe = N * N + N + TRI, oZ = N * N + 2 * N + TRI;
  extern __shared__ __align__(16) double coop_sm[];
  const int g = threadIdx.x / G, l = threadIdx.x % G;
  const int gbase = (threadIdx.x & 31) - l;
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kCChunks + g;
  double* const buf = coop_sm + g * CS::SZ;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;

  double A[R][N], Y[R][N], Z[R][N], b[R], eta[R];
#pragma unroll
  for (int s = 0; s < R; ++s) {
    b[s] = 0.0;
    eta[s] = 0.0;
#pragma unroll
    for (int q = 0; q < N; ++q) {
      A[s][q] = (q == l + s * G) ? 1.0 : 0.0;
      Y[s][q] = 0.0;
      Z[s][q] = 0.0;
    }
  }
  double* const own = chunk_own + seq * NF * Ppad + c;
  const bool tv_t = (a.tF | a.tQ | a.tb) != 0, tv_o = (a.tH | a.tR | a.tc) != 0;
  double* const bank = coop_sm + kCChunks * CS::SZ;
  CoopBank<N, NY>::fill(bank, a, seq, tv_t, tv_o, true);
  __syncthreads();
  CoopBank<N, NY> md;
  const bool hl = l < NY;
  const int la = hl ? l : 0;
#pragma unroll 1
  for (int j = 0; j < K; ++j) {
    const bool act = j < len;
    const long long k = act ? k0 + j : 0;
    md.at(bank, a, seq, k, tv_t, tv_o, true, vec);
    const double yv = hl ? a.y[seq * a.sy + k * a.ty + l] : 0.0;
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      st_row<N>(buf + CS::R0 + rs * RS, Y[s]);
      st_row<N>(buf + CS::R1 + rs * RS, A[s]);
      buf[CS::V0 + rs] = b[s];
    }
    __syncwarp();
    double M1[R][2 * N], FA[R][N], mp[R];
    {
      double Fr[R][N];
#pragma unroll
      for (int s = 0; s < R; ++s) {
        gld_row<N>(md.F + (l + s * G) * N, Fr[s], md.vt);
        mp[s] = md.bq[l + s * G];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          M1[s][q] = 0.0;
          FA[s][q] = 0.0;
        }
      }
#pragma unroll
      for (int kk = 0; kk < N; ++kk) {
        double t[N], u[N];
        ld_row<N>(buf + CS::R0 + kk * RS, t);
        ld_row<N>(buf + CS::R1 + kk * RS, u);
        const double bk = buf[CS::V0 + kk];
#pragma unroll
        for (int s = 0; s < R; ++s) {
          const double f = Fr[s][kk];
#pragma unroll
          for (int q = 0; q < N; ++q) {
            M1[s][q] = fma(f, t[q], M1[s][q]);
            FA[s][q] = fma(f, u[q], FA[s][q]);
          }
          mp[s] = fma(f, bk, mp[s]);
        }
      }
#pragma unroll
      for (int s = 0; s < R; ++s) {
        double Qr[N];
        gld_row<N>(md.Q + (l + s * G) * N, Qr, md.vt);
#pragma unroll
        for (int q = 0; q < N; ++q) M1[s][N + q] = (q <= l + s * G) ? Qr[q] : 0.0;
      }
    }
    __syncwarp();
    coop_house_rows<2 * N, N, N, G, R>(M1, l, gbase);
    double Np[R][N];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
#pragma unroll
      for (int q = 0; q < N; ++q) Np[s][q] = (q <= rs) ? M1[s][q] : 0.0;
      if (act && j + 1 == len) {
        own[(ob + rs) * Ppad] = mp[s];
        own[(oe + rs) * Ppad] = eta[s];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          own[(oA + rs * N + q) * Ppad] = FA[s][q];
          if (q <= rs) {
            own[(oU + rs * (rs + 1) / 2 + q) * Ppad] = Np[s][q];
            own[(oZ + rs * (rs + 1) / 2 + q) * Ppad] = Z[s][q];
          }
        }
      }
      st_row<N>(buf + CS::R0 + rs * RS, Np[s]);
      st_row<N>(buf + CS::R1 + rs * RS, FA[s]);
      buf[CS::V0 + rs] = mp[s];
    }
    __syncwarp();
    double hrow[N + NY], mrow[R][N + NY], Vc[R][NY];
    double res = hl ? yv - md.c[la] : 0.0;
    {
      double Hr[N];
      gld_row<N>(md.H + la * N, Hr, md.vo);
#pragma unroll
      for (int q = 0; q < N; ++q) {
        Hr[q] = hl ? Hr[q] : 0.0;
        hrow[q] = 0.0;
      }
#pragma unroll
      for (int s = 0; s < R; ++s)
#pragma unroll
        for (int q = 0; q < NY; ++q) Vc[s][q] = 0.0;
#pragma unroll
      for (int kk = 0; kk < N; ++kk) {
        double t[N];
        ld_row<N>(buf + CS::R0 + kk * RS, t);
        const double h = Hr[kk];
#pragma unroll
        for (int q = 0; q < N; ++q) hrow[q] = fma(h, t[q], hrow[q]);
        res = fma(-h, buf[CS::V0 + kk], res);
#pragma unroll
        for (int s = 0; s < R; ++s) {
          const double fa = buf[CS::R1 + kk * RS + l + s * G];      // (F A)[kk][column l + s G]
#pragma unroll
          for (int q = 0; q < NY; ++q) Vc[s][q] = fma(md.H[q * N + kk], fa, Vc[s][q]);
        }
      }
#pragma unroll
      for (int q = 0; q < NY; ++q) hrow[N + q] = hl ? md.R[la * NY + q] : 0.0;
    }
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int q = 0; q < N; ++q) mrow[s][q] = Np[s][q];
#pragma unroll
      for (int q = 0; q < NY; ++q) mrow[s][N + q] = 0.0;
    }
    __syncwarp();
    coop_update_reflectors_rows<N, NY, G, R>(hrow, mrow, l, gbase);
    double P11[NY][NY], inv[NY], rr[NY], quad, det;
    coop_psi11_solve<N, NY>(hrow, res, gbase, P11, inv, rr, quad, det);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int q = 0; q < NY; ++q) {
        double v = Vc[s][q];
#pragma unroll
        for (int p = 0; p < q; ++p) v = fma(-P11[q][p], Vc[s][p], v);
        Vc[s][q] = v * inv[q];
      }
#pragma unroll
      for (int q = 0; q < NY; ++q) buf[CS::R0 + q * RS + l + s * G] = Vc[s][q];
    }
    __syncwarp();
    double An[R][N], bn[R], en[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      bn[s] = mp[s];
      en[s] = eta[s];
#pragma unroll
      for (int q = 0; q < N; ++q) An[s][q] = FA[s][q];
    }
#pragma unroll
    for (int p = 0; p < NY; ++p) {
      double t[N];
      ld_row<N>(buf + CS::R0 + p * RS, t);
#pragma unroll
      for (int s = 0; s < R; ++s) {
#pragma unroll
        for (int q = 0; q < N; ++q) An[s][q] = fma(-mrow[s][p], t[q], An[s][q]);
        bn[s] = fma(mrow[s][p], rr[p], bn[s]);
        en[s] = fma(Vc[s][p], rr[p], en[s]);
      }
    }
    double Zn[R][N], Wv[R][NY];
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int q = 0; q < N; ++q) Zn[s][q] = Z[s][q];
#pragma unroll
      for (int q = 0; q < NY; ++q) Wv[s][q] = Vc[s][q];
    }
    coop_tria_append_rows<N, NY, G, R>(Zn, Wv, l, gbase);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rs = l + s * G;
      b[s] = act ? bn[s] : b[s];
      eta[s] = act ? en[s] : eta[s];
#pragma unroll
      for (int q = 0; q < N; ++q) {
        A[s][q] = act ? An[s][q] : A[s][q];
        Y[s][q] = act ? mrow[s][NY + q] : Y[s][q];
        Z[s][q] = act ? ((q <= rs) ? Zn[s][q] : 0.0) : Z[s][q];
      }
    }
    __syncwarp();
  }
  coop_house_rows<N, N - 1, 0, G, R>(Y, l, gbase);
  double* sp = summ + seq * NF * Ppad + c;
#pragma unroll
  for (int s = 0; s < R; ++s) {
    const int rs = l + s * G;
    sp[(ob + rs) * Ppad] = b[s];
    sp[(oe + rs) * Ppad] = eta[s];
#pragma unroll
    for (int q = 0; q < N; ++q) {
      sp[(oA + rs * N + q) * Ppad] = A[s][q];
      if (q <= rs) {
        sp[(oU + rs * (rs + 1) / 2 + q) * Ppad] = Y[s][q];
        sp[(oZ + rs * (rs + 1) / 2 + q) * Ppad] = Z[s][q];
      }
    }
  }
}

}  // namespace psq
