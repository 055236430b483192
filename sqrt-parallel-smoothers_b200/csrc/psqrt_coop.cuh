// psqrt_coop.cuh -- the filtering mid-level scan (K2) with one HALF-WARP per combine.
//
// The mid-level scan is a dozen dependent combines over ~1e3 warp totals: pure latency.  With one
// thread per combine (k_mid_scan) a filtering combine is ~2000 FP64 instructions of straight-line code
// (4.8 us per Kogge-Stone level at nx = 4, a third of it instruction fetch).  Here the 16 lanes of a
// half-warp share ONE combine:
//   * the elements live in shared memory, un-packed (dense U and Z), so no index arithmetic on
//     triangular storage is left in the dependent path;
//   * every triangularisation keeps one matrix ROW per lane in registers; the pivot row of a
//     Householder reflector is published in shared memory and all rows below it are updated at once
//     (coop_house) -- the serial depth of tria([2n x 2n]) drops from ~n(2n)^2 to ~n(3n) operations;
//   * the small matrix products are one output ENTRY per lane.
// A warp carries two independent combines (lanes 0-15, 16-31), so a group of 32 items is a CTA of 16
// warps: 4 warps per scheduler, which keeps the FP64 pipe (one warp instruction per 2 cycles per
// scheduler, whatever the number of active lanes) below the dependency latency of a combine.
// The combine has ONE call site (a flat step loop serves both scan levels): cold straight-line code is
// what a latency-bound kernel pays for.
//
// Formulas: filtering combine  parsmooth/parallel/_operators.py:58-77
// Factors differ from the per-thread path by column signs only (tria is unique up to those,
// parsmooth/_utils.py:22-24); Xi22 enters Z = tria([A1^T Xi22 | Z1]) un-triangularised, which leaves
// Z Z^T unchanged.  The smoothing mid scan (K4) stays on k_mid_scan: its combine is one small tria and
// the per-thread form is already at the latency floor (measured: 15 us vs 28 us cooperative).
// Used for nx >= 5, where the per-thread combine of k_mid_scan spills (2 KB of stack per thread at nx = 5,
// 12 KB at nx = 8); at nx = 4 the two forms take the same time and the per-thread one stays.
#pragma once
#include "psqrt_math.cuh"

namespace psq {

constexpr int kCoopWarps = 16;   // warps per CTA of the cooperative mid scan (2 items each)

// Householder triangularisation from the right of a matrix held one ROW per lane.
// The lane with row index r (< R) holds row r in row[0..C).  Rows [0, NREFL) become lower-trapezoidal;
// TRIBLK as in house_rows (row r has nothing right of column TRIBLK + r).  pv: 2 C doubles of shared
// scratch shared by exactly the lanes that work on this matrix (double-buffered pivot row).
// Every lane of the warp must call (warp-level syncs); lanes with r >= R compute and discard.
// Branch-free: selects instead of divergent updates, one __syncwarp per reflector.
template <int C, int NREFL, int TRIBLK, int R>
__device__ __forceinline__ void coop_house(double (&row)[C], const int r, double* pv) {
  static_for<0, NREFL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int kend = (TRIBLK > 0) ? ((TRIBLK + j + 1 < C) ? TRIBLK + j + 1 : C) : C;
    if constexpr (j + 1 < kend) {
      double* buf = pv + (j & 1) * C;
      if (r == j) {
#pragma unroll
        for (int k = j; k < kend; ++k) buf[k] = row[k];
      }
      __syncwarp();
      double p[C];
#pragma unroll
      for (int k = j; k < kend; ++k) p[k] = buf[k];
      const double alpha = p[j];
      double sigma = 0.0, sigma2 = 0.0;
#pragma unroll
      for (int k = j + 1; k < kend; k += 2) {
        sigma = fma(p[k], p[k], sigma);
        if (k + 1 < kend) sigma2 = fma(p[k + 1], p[k + 1], sigma2);
      }
      sigma += sigma2;
      // the lane's own dot product does not depend on the norm: it overlaps the rsqrt / rcp chain
      double d = 0.0, d2 = 0.0;
#pragma unroll
      for (int k = j + 1; k < kend; k += 2) {
        d = fma(row[k], p[k], d);
        if (k + 1 < kend) d2 = fma(row[k + 1], p[k + 1], d2);
      }
      d += d2;
      const double q = fma(alpha, alpha, sigma);  // branch-free like house_rows (mask = 0: H = I)
      const double mask = (q != 0.0) ? 1.0 : 0.0;
      const double qs = (q != 0.0) ? q : 1.0;
      const double norm = qs * rsqrt_nr(qs);
      const double beta = -copysign(norm, alpha) * mask;
      const double v0 = alpha - beta;
      const double s = rcp_nr(fma(fabs(alpha), norm, qs)) * mask;
      d = fma(row[j], v0, d) * s;
      const bool below = (r > j) && (r < R);
      const double dd = below ? d : 0.0;   // rows that are not below the pivot stay as they are
      row[j] = (r == j) ? beta : fma(-dd, v0, row[j]);
#pragma unroll
      for (int k = j + 1; k < kend; ++k) row[k] = fma(-dd, p[k], row[k]);
    }
  });
}

// ---------------------------------------------------------------------------------------------
// Filtering combine e1 (x) e2, e1 = earlier / accumulated side.  Elements in the DENSE shared-memory
// layout below (out distinct from e1, e2); ws: WS doubles of scratch private to this half-warp;
// l = lane & 15.  Both halves of a warp must call together.
// ---------------------------------------------------------------------------------------------
template <int N>
struct CoopF {
  static constexpr int NN = N * N;
  static constexpr int TRI = N * (N + 1) / 2;
  // Slot layout.  Dense (N <= 6): A [N][N], b [N], U [N][N] (upper part zero), eta [N], Z [N][N] (upper part
  // zero) -- no index arithmetic on triangular storage.  Packed (N >= 7, where three dense slot sets and the
  // scratch would exceed the 227 KB of shared memory): FElem<N>'s own layout, lower triangles row by row.
  static constexpr bool PACKED = (N > 6);
  static constexpr int dA = 0, db = NN, dU = NN + N;
  static constexpr int de = PACKED ? NN + N + TRI : 2 * NN + N;
  static constexpr int dZ = PACKED ? NN + 2 * N + TRI : 2 * NN + 2 * N;
  static constexpr int NFD = PACKED ? FElem<N>::NF + (FElem<N>::NF & 1) : 3 * NN + 2 * N + (N & 1);
  static constexpr int WS = 5 * NN + 11 * N + (N & 1);
  // entry (i, j) of the lower-triangular factor stored at `base` (zero above the diagonal)
  static __device__ __forceinline__ double tril(const double* base, int i, int j) {
    if constexpr (PACKED) return (j <= i) ? base[i * (i + 1) / 2 + j] : 0.0;
    else return base[i * N + j];
  }
  // FElem<N>::v index -> slot offset
  static __device__ int dense_of(int f) {
    if constexpr (PACKED) return f;
    constexpr int pU = NN + N, pe = NN + N + TRI, pZ = NN + 2 * N + TRI;
    if (f < pU) return f;                       // A, b
    if (f >= pe && f < pZ) return de + (f - pe);
    const int t = (f < pe) ? f - pU : f - pZ;   // triangular index i (i + 1) / 2 + j
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    const int j = t - i * (i + 1) / 2;
    return ((f < pe) ? dU : dZ) + i * N + j;
  }
  static __device__ __forceinline__ double ident(int off) { return (off < NN && off / N == off % N) ? 1.0 : 0.0; }

  static __device__ __forceinline__ void combine(const double* __restrict__ e1, const double* __restrict__ e2,
                                                 double* __restrict__ out, double* __restrict__ ws, const int l) {
    static_assert(2 * N <= 16 && N <= 8, "a half-warp holds the 2N rows of Xi");
    double* X11 = ws;            // Xi11 (lower part meaningful); dead once T1 exists, then W lives here
    double* W = ws;              // A2 T1^T
    double* X21 = ws + NN;       // Xi21
    double* BB = ws + 2 * NN;    // bottom-right block of the triangularised Xi: a square root of Xi22 Xi22^T
    double* T1 = ws + 3 * NN;    // Xi11^{-1} U1^T, later G = A2 - W Xi21^T
    double* P = ws + 4 * NN;     // U1^T Z2, later Xi21 T1, later A1^T BB
    double* tv = ws + 5 * NN;    // b1 + U1 U1^T eta2
    double* sv = tv + N;         // eta2 - Z2 Z2^T b1
    double* sv2 = sv + N;        // (I - Xi21 T1) sv
    double* pv = sv2 + N;        // pivot rows: 2 matrices x 2 buffers x 2N
    // P <- U1^T Z2 (dense factors: the zeros of the upper triangles do the masking)   _operators.py:63
    for (int it = l; it < NN; it += 16) {
      const int i = it / N, j = it % N;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(tril(e1 + dU, k, i), tril(e2 + dZ, k, j), s);
      P[it] = s;
    }
    __syncwarp();
    // Xi = [[U1^T Z2, I], [Z2, 0]], one row per lane                                  _operators.py:63-64
    double row[2 * N];
    {
      const bool top = l < N, bot = (l >= N) && (l < 2 * N);
      const int zi = bot ? l - N : 0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double a = 0.0;
        if (top) a = P[l * N + j];
        if (bot) a = tril(e2 + dZ, zi, j);
        row[j] = a;
        row[N + j] = (l == j) ? 1.0 : 0.0;
      }
    }
    coop_house<2 * N, N, N, 2 * N>(row, l, pv);
    if (l < 2 * N) {
      double* dst = (l < N) ? X11 + l * N : X21 + (l - N) * N;
#pragma unroll
      for (int j = 0; j < N; ++j) dst[j] = row[j];
      if (l >= N) {
#pragma unroll
        for (int j = 0; j < N; ++j) BB[(l - N) * N + j] = row[N + j];
      }
    }
    __syncwarp();
    // T1 = Xi11^{-1} U1^T, one column per lane; two more lanes form tv and sv
    if (l < N) {
      double tc[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = tril(e1 + dU, l, i);  // (U1^T)[i][l]; zero for l < i
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-X11[i * N + k], tc[k], s);
        tc[i] = s * rcp_nr(X11[i * N + i]);
        T1[i * N + l] = tc[i];
      }
    } else if (l == 8 || l == 9) {
      // x = a + sgn L (L^T v):  l == 8: tv = b1 + U1 (U1^T eta2);  l == 9: sv = eta2 - Z2 (Z2^T b1)
      const bool first = (l == 8);
      const double* Lp = first ? e1 + dU : e2 + dZ;
      const double* vp = first ? e2 + de : e1 + db;
      const double* ap = first ? e1 + db : e2 + de;
      double* xo = first ? tv : sv;
      const double sgn = first ? 1.0 : -1.0;
      double u[N];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = i; k < N; ++k) s = fma(tril(Lp, k, i), vp[k], s);
        u[i] = s;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k <= i; ++k) s = fma(tril(Lp, i, k), u[k], s);
        xo[i] = fma(sgn, s, ap[i]);
      }
    }
    __syncwarp();
    // W = A2 T1^T  and  P = Xi21 T1
    for (int it = l; it < 2 * NN; it += 16) {
      const bool second = it >= NN;
      const int e = second ? it - NN : it;
      const int i = e / N, j = e % N;
      const double* X = second ? X21 + i * N : e2 + dA + i * N;
      const double* Y = second ? T1 + j : T1 + j * N;
      const int ys = second ? N : 1;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) s = fma(X[k], Y[k * ys], s);
      (second ? P : W)[e] = s;
    }
    __syncwarp();
    // G = A2 - W Xi21^T (over T1)  and  sv2 = sv - P sv
    for (int it = l; it < NN + N; it += 16) {
      if (it < NN) {
        const int i = it / N, j = it % N;
        double s = e2[dA + it];
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(-W[i * N + k], X21[j * N + k], s);
        T1[it] = s;
      } else {
        const int i = it - NN;
        double s = sv[i];
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(-P[i * N + k], sv[k], s);
        sv2[i] = s;
      }
    }
    __syncwarp();
    // A = G A1, b = G tv + b2, eta = A1^T sv2 + eta1 (_operators.py:70-74); P <- A1^T BB
    for (int it = l; it < 2 * NN + 2 * N; it += 16) {
      if (it < 2 * NN) {
        // two products with the same shape: out = X^T-or-X times Y
        const bool second = it >= NN;
        const int e = second ? it - NN : it;
        const int i = e / N, j = e % N;
        // first:  P[i][j]   = sum_k A1[k][i] BB[k][j]      second: A[i][j] = sum_k G[i][k] A1[k][j]
        const double* X = second ? T1 + i * N : e1 + dA + i;
        const int xs = second ? 1 : N;
        const double* Y = (second ? e1 + dA : BB) + j;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(X[k * xs], Y[k * N], s);
        (second ? out + dA : P)[e] = s;
      } else {
        const bool second = it >= 2 * NN + N;
        const int i = second ? it - 2 * NN - N : it - 2 * NN;
        // first: b[i] = b2[i] + sum_k G[i][k] tv[k]         second: eta[i] = eta1[i] + sum_k A1[k][i] sv2[k]
        const double* X = second ? e1 + dA + i : T1 + i * N;
        const int xs = second ? N : 1;
        const double* v = second ? sv2 : tv;
        double s = second ? e1[de + i] : e2[db + i];
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(X[k * xs], v[k], s);
        out[(second ? de : db) + i] = s;
      }
    }
    __syncwarp();
    // U = tria([W | U2]) on lanes 0..7, Z = tria([A1^T BB | Z1]) on lanes 8..15          _operators.py:72,75
    {
      const int r = l & 7;
      const int sub = l & 8;
      const bool act = r < N;
      const double* Mx = (sub ? P : W) + r * N;
      const double* Tp = sub ? e1 + dZ : e2 + dU;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        row[j] = act ? Mx[j] : 0.0;
        row[N + j] = act ? tril(Tp, r, j) : 0.0;
      }
      coop_house<2 * N, N, N, N>(row, r, pv + (sub ? 4 * N : 0));
      if (act) {
        if constexpr (PACKED) {
          double* o = out + (sub ? dZ : dU) + r * (r + 1) / 2;
#pragma unroll
          for (int j = 0; j < N; ++j)
            if (j <= r) o[j] = row[j];
        } else {
          double* o = out + (sub ? dZ : dU) + r * N;
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = (j <= r) ? row[j] : 0.0;
        }
      }
    }
    __syncwarp();
  }
};

// =========================================================================================
// K2, cooperative form.  Same contract as k_mid_scan<FElem<N>, false> (psqrt_kernels.cuh):
//   pass 0 (every CTA): exclusive scan of the CTA's group of 32 items in place, group total to groups[];
//   pass 1 (the CTA that takes the last ticket): exclusive scan of the G group totals in place, sequence
//           total to total_out, ticket counter re-armed.
// Both passes are the same flat step loop: half-warp x owns q consecutive items (q = 1 in pass 0,
// ceil(G / 32) in pass 1): q - 1 sequential combines, 5 Kogge-Stone steps across the half-warps, q - 1
// combines that run the exclusive prefix through the owned items.  Three element slots per half-warp.
// =========================================================================================
template <int N>
constexpr size_t coop_smem_bytes() {
  return sizeof(double) * (size_t)(3 * 32 * CoopF<N>::NFD + 32 * CoopF<N>::WS) + sizeof(int) * FElem<N>::NF;
}

template <int N>
__global__ void __launch_bounds__(32 * kCoopWarps, 1)
k_mid_scan_coop(double* __restrict__ items, long long M, double* __restrict__ groups, long long G,
                unsigned int* __restrict__ counter, double* __restrict__ total_out) {
  using Op = CoopF<N>;
  constexpr int NF = FElem<N>::NF;
  constexpr int NFD = Op::NFD;
  extern __shared__ __align__(16) double coop_sm[];
  __shared__ unsigned int s_ticket;
  double* const slots = coop_sm;                    // [3][32][NFD]
  double* const wsall = coop_sm + 3 * 32 * NFD;     // [32][WS]
  int* const dmap = reinterpret_cast<int*>(wsall + 32 * Op::WS);   // packed field -> dense offset
  const long long seq = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int l = lane & 15;
  const int x = (threadIdx.x >> 4);                 // half-warp index
  double* const ws = wsall + x * Op::WS;
  auto slot = [&](int s, int xx) { return slots + (s * 32 + xx) * NFD; };

  if (threadIdx.x < NF) dmap[threadIdx.x] = Op::dense_of(threadIdx.x);
  for (int k = threadIdx.x; k < 3 * 32 * NFD; k += blockDim.x) slots[k] = 0.0;   // upper triangles stay zero
  __syncthreads();

#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    double* const arr = pass ? groups : items;
    const long long Mp = pass ? G : M;
    const int q = pass ? (int)((G + 31) / 32) : 1;
    const long long s0 = pass ? (long long)x * q : (long long)blockIdx.x * 32 + x;
    double* const base = arr + seq * NF * Mp;
    // element at scan-order index idx -> dense slot (identity past the end)
    auto load = [&](long long idx, double* d) {
      for (int f = l; f < NF; f += 16) {
        const int off = dmap[f];
        d[off] = (idx < Mp) ? __ldcg(base + f * Mp + idx) : Op::ident(off);
      }
    };
    auto store = [&](long long idx, const double* s) {
      if (idx < Mp)
        for (int f = l; f < NF; f += 16) base[f * Mp + idx] = s[dmap[f]];
    };
    int a = 0, b = 1, c = 2;
    load(s0, slot(a, x));
    __syncwarp();
    const int ks0 = q - 1, run0 = q - 1 + 5, nsteps = run0 + q - 1;
#pragma unroll 1
    for (int st = 0; st < nsteps; ++st) {
      const double *e1, *e2;
      double* out;
      bool keep = false;
      if (st < ks0) {            // own items folded into one accumulator (slot a)
        load(s0 + st + 1, slot(b, x));
        __syncwarp();
        e1 = slot(a, x); e2 = slot(b, x); out = slot(c, x);
      } else if (st < run0) {    // Kogge-Stone across the 32 half-warps on slot a
        const int d = 1 << (st - ks0);
        if (st == ks0) __syncthreads();
        keep = x < d;            // combines with itself, result discarded: keeps the warp converged
        e1 = slot(a, keep ? x : x - d); e2 = slot(a, x); out = slot(c, x);
      } else {                   // the exclusive prefix (slot b) runs through the owned items
        const long long idx = s0 + (st - run0);
        load(idx, slot(c, x));
        store(idx, slot(b, x));
        __syncwarp();
        e1 = slot(b, x); e2 = slot(c, x); out = slot(a, x);
      }
      Op::combine(e1, e2, out, ws, l);
      if (keep) {
        for (int k = l; k < NFD; k += 16) out[k] = e2[k];
      }
      if (st < ks0) {
        const int t = a; a = c; c = t;
      } else if (st < run0) {
        __syncthreads();
        const int t = a; a = c; c = t;
        if (st == run0 - 1) {    // slot a now holds the inclusive scan over the half-warps' accumulators
          if (x == 31) {
            const double* s = slot(a, 31);
            if (pass == 0) {
              for (int f = l; f < NF; f += 16) groups[(seq * NF + f) * G + blockIdx.x] = s[dmap[f]];
            } else if (total_out) {
              for (int f = l; f < NF; f += 16) total_out[seq * NF + f] = s[dmap[f]];
            }
          }
          const double* s = slot(a, x > 0 ? x - 1 : 0);
          double* o = slot(b, x);
          for (int k = l; k < NFD; k += 16) o[k] = (x > 0) ? s[k] : Op::ident(k);
          __syncthreads();       // slots a and c of half-warp x are private again from here on
        }
      } else {
        const int t = b; b = a; a = t;
      }
    }
    __syncwarp();
    store(s0 + q - 1, slot(b, x));
    if (pass == 0) {
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_ticket = atomicAdd(counter + seq, 1u);
      __syncthreads();
      if (s_ticket != (unsigned int)(G - 1)) return;
      __threadfence();
    }
  }
  if (threadIdx.x == 0) counter[seq] = 0u;
}

}  // namespace psq
