// psqrt_coop2.cuh -- sub-warp ("one matrix row per lane") forms of the two associative combines and the
// mid-level scans K2 / K4 built on them.
//
// The mid-level scans are a dozen DEPENDENT combines over ~1e3 warp totals while the rest of the GPU idles:
// what counts is the latency of one combine, not throughput.  One thread per combine (k_mid_scan) needs
// ~1200 instructions of straight-line code per level and a lone warp issues them at ~7 cycles each
// (dependent-issue latency + instruction fetch: profiles/r01_ncu_full_v5_restructured.txt).  Here a group of
// G lanes (8 for nx <= 4, 16 for nx <= 8) shares ONE combine:
//   * every triangularisation keeps one matrix ROW per lane in registers; the pivot row of a Householder
//     reflector travels by warp shuffles (coop_house_shfl) -- the in-register Householder tria via shuffles of
//     the north star; the serial depth of tria([2n x 2n]) drops from ~n (2n)^2 to ~n (3n) operations;
//   * the small matrix products are one output ROW per lane, operands broadcast from shared memory;
//   * top lanes (rows of [U1^T Z2 | I]) and bottom lanes (rows of [Z2 | 0]) run the SAME instruction stream on
//     different pointers (selected once per combine): no divergent code, so a warp carries 32 / G combines at
//     the cost of one.
// Elements live in shared memory in a DENSE layout (U, Z with explicit zeros above the diagonal), so the KS
// partner of an item is a pointer, not a data movement.
//
// Formulas: filtering combine  parsmooth/parallel/_operators.py:58-77
//           smoothing combine  parsmooth/parallel/_operators.py:118-125
// Factors differ from the per-thread path by column signs only (tria is unique up to those,
// parsmooth/_utils.py:22-24); Xi22 enters Z = tria([A1^T Xi22 | Z1]) un-triangularised, which leaves Z Z^T
// unchanged.
#pragma once
#include "psqrt_math.cuh"

namespace psq {

// Items per group of the mid-level scans == warps per "group" of the sweeps' three-level prefix hierarchy
// (chunk -> warp -> group -> sequence).  64 for nx <= 4: at most 1184 warps per sequence (148 SMs x 256 threads)
// give <= 19 groups, so both levels are single Kogge-Stone passes (6 + 5 = the minimal 11 dependent combines);
// 32 for larger nx, where 64 dense elements would not fit the shared memory of one SM.
#ifndef PSQ_MID2
#define PSQ_MID2 1   // 0: the one-thread-per-combine mid scans of round 1 (k_mid_scan / k_mid_scan_coop), for A/B timing
#endif
template <int N>
struct MidCfg {
  static constexpr int G = (N <= 4) ? 8 : 16;      // lanes per filtering combine (>= 2 N rows of Xi)
  static constexpr int GS = (N <= 4) ? 4 : 8;      // lanes per smoothing combine (>= N rows)
#if PSQ_MID2
  static constexpr int IT = (N <= 4) ? 64 : 32;    // items per CTA
#else
  static constexpr int IT = 32;
#endif
};

// Strides (in doubles) of per-lane-group regions of shared memory: 2 (mod 4), so that the 16-byte accesses of
// the 2 - 4 lane groups of a warp (same offset, consecutive regions) fall into different banks.
constexpr int spread_stride(int n) { return n + ((2 - n % 4) + 4) % 4; }

// ---------------------------------------------------------------------------------------------
// Time-shard exchange over peer-mapped memory (NVLink), fused into the kernels that produce / consume the shard
// totals.  Every rank owns one exchange buffer mapped into all peers.  The CTA that finishes a mid-level scan
// stores the shard total (plus optional extra payload) straight into slot `rank` of EVERY rank's buffer and then
// publishes the pass number there (PushArgs, k_mid_scan2); the carry kernel of the consumer waits for the pass
// number of ALL ranks before it folds the totals it needs (k_carry_scan).  Pass numbers are counted on the device
// (one counter per sequence and phase), so a pass has no per-step host argument and replays from a CUDA graph.
// Slots are double-buffered by pass parity; because every consumer waits for all ranks, no rank can be more than
// one pass ahead of a reader of its slot, in filter-only passes too.
// ---------------------------------------------------------------------------------------------
struct PeerCtx {                      // one phase (filter or smoother) of the exchange
  double* const* bufs;                // [n_ranks] every rank's exchange buffer as mapped into THIS rank's address space
  int rank, n_ranks;
  long long batch;
  long long flags_off;                // [n_ranks][batch] 64-bit pass numbers published by each rank (in this rank's buffer)
  long long ctr_off;                  // [batch] this rank's own pass counter
  long long data_off;                 // [2 halves][n_ranks][slot] doubles (double-buffered by pass parity)
  long long slot;                     // doubles per rank: batch * payload
  long long payload;                  // doubles per (rank, sequence): packed element + extras
};
struct PushArgs {
  PeerCtx pc;
  int on;                             // 0: no exchange (single GPU, or NCCL all-gather by the host)
  const double *x1, *x2;              // extra payload after the packed total: n1 doubles at x1 + seq * s1, then n2 at x2 + seq * s2
  long long s1, s2;
  int n1, n2;
};

// -DPSQ_MID_TRACE: thread 0 of every CTA of the mid scans records %globaltimer at phase boundaries (development
// aid; read back with psqrt_debug_trace in psqrt_capi.cu).  Slot layout: [kernel kind 0/1][cta < 128][stamp < 16].
#if defined(PSQ_MID_TRACE)
__device__ unsigned long long g_mid_trace[2][128][16];
__device__ __forceinline__ void mid_trace(int kind, int stamp) {
  if (threadIdx.x == 0 && blockIdx.x < 128 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_mid_trace[kind][blockIdx.x][stamp] = t;
  }
}
#define PSQ_TRACE(kind, stamp) mid_trace(kind, stamp)
#else
#define PSQ_TRACE(kind, stamp)
#endif

template <int N>
__device__ __forceinline__ void ld_row(const double* __restrict__ p, double (&r)[N]) {
  if constexpr (N % 2 == 0) {
#pragma unroll
    for (int k = 0; k < N; k += 2) {
      const double2 v = *reinterpret_cast<const double2*>(p + k);
      r[k] = v.x;
      r[k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = p[k];
  }
}
template <int N>
__device__ __forceinline__ void st_row(double* __restrict__ p, const double (&r)[N]) {
  if constexpr (N % 2 == 0) {
#pragma unroll
    for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(p + k) = make_double2(r[k], r[k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < N; ++k) p[k] = r[k];
  }
}

// Copy one dense slot (NFD doubles) with the G lanes of a group, all loads in flight before the first store
// (a rolled load -> store loop serialises the shared-memory / DSMEM latencies).
template <int NFD, int G>
__device__ __forceinline__ void copy_slot(double* __restrict__ dst, const double* __restrict__ src, const int l) {
  constexpr int PER = (NFD + G - 1) / G;
  double v[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) v[q] = (l + q * G < NFD) ? src[l + q * G] : 0.0;
#pragma unroll
  for (int q = 0; q < PER; ++q)
    if (l + q * G < NFD) dst[l + q * G] = v[q];
}

// Householder triangularisation from the right of a matrix held one ROW per lane: the lane with row index r
// (0 <= r < R) holds row r in row[0..C); the lane holding row 0 is warp lane src0 (rows are on consecutive
// lanes).  Rows [0, NREFL) become lower-trapezoidal; TRIBLK as in house_rows (row r has nothing right of column
// TRIBLK + r).  Lanes with r outside [0, R) execute the same instructions and keep their registers.
// Every lane of the warp must call.  Branch-free; one shuffle round per reflector.
template <int C, int NREFL, int TRIBLK, int R>
__device__ __forceinline__ void coop_house_shfl(double (&row)[C], const int r, const int src0) {
  static_for<0, NREFL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    constexpr int kend = (TRIBLK > 0) ? ((TRIBLK + j + 1 < C) ? TRIBLK + j + 1 : C) : C;
    if constexpr (j + 1 < kend) {
      double p[C];
#pragma unroll
      for (int k = j; k < kend; ++k) p[k] = __shfl_sync(0xffffffffu, row[k], src0 + j);
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
// Filtering combine e1 (x) e2 (e1 = earlier / accumulated side) by G lanes.  Elements in the dense
// shared-memory layout below; `out` distinct from e1, e2; ws: WS doubles private to this lane group.
// l = lane index within the group (0 .. G-1), gbase = warp lane of the group's lane 0.
// All 32 lanes of the warp must call (groups of one warp run in lockstep).
// ---------------------------------------------------------------------------------------------
template <int N>
struct CoopF2 {
  static constexpr int NN = N * N;
  static constexpr int G = MidCfg<N>::G;
  static_assert(2 * N <= G && G <= 32, "a lane group holds the 2N rows of Xi");
  // dense slot: A [N][N], b [N], U [N][N] (zeros above the diagonal), eta [N], Z [N][N] (zeros above the diagonal)
  static constexpr int dA = 0, db = NN, dU = NN + N, de = 2 * NN + N, dZ = 2 * NN + 2 * N;
  static constexpr int NFD = spread_stride(3 * NN + 2 * N);
  // workspace: X11, X21, BB, T1, T1T [N][N] each; invd [N]; tvsv [2N]; sv2 [N]
  static constexpr int wX11 = 0, wX21 = NN, wBB = 2 * NN, wT1 = 3 * NN, wT1T = 4 * NN, wInv = 5 * NN, wTvSv = 5 * NN + N,
                       wSv2 = 5 * NN + 3 * N;
  static constexpr int WS = spread_stride(5 * NN + 4 * N);

  // FElem<N>::v index -> slot offset
  static __host__ __device__ int dense_of(int f) {
    constexpr int TRI = N * (N + 1) / 2;
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
                                                 double* __restrict__ out, double* __restrict__ ws, const int l,
                                                 const int gbase) {
    const bool top = l < N;
    const bool bot = (l >= N) && (l < 2 * N);
    const int i = top ? l : (bot ? l - N : 0);          // row index within the half (0 for idle lanes)
    const int hbase = gbase + (top ? 0 : N);            // warp lane of row 0 of this lane's half
    // ---- stage A: rows of Xi = [[U1^T Z2, I], [Z2, 0]] (_operators.py:63-64); tv = b1 + U1 U1^T eta2 (top),
    //      sv = eta2 - Z2 Z2^T b1 (bottom) with the column / row of the lane's own factor
    const double* own = top ? e1 + dU : e2 + dZ;         // top: U1, bottom: Z2
    const double* vin = top ? e2 + de : e1 + db;         // vector the factor's transpose is applied to
    const double* add = top ? e1 + db : e2 + de;         // vector the result is added to
    double row[2 * N];
    double ownrow[N];
    {
      double col[N], z[N];
      double dcol = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        col[k] = own[k * N + i];                         // column i of the own factor
        dcol = fma(col[k], vin[k], dcol);                // (U1^T eta2)_i  /  (Z2^T b1)_i
      }
      ld_row<N>(own + i * N, ownrow);                    // row i of the own factor
#pragma unroll
      for (int j = 0; j < N; ++j) row[j] = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        ld_row<N>(e2 + dZ + k * N, z);
#pragma unroll
        for (int j = 0; j < N; ++j) row[j] = fma(col[k], z[j], row[j]);   // top: (U1^T Z2)[i][:]
      }
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        const double dk = __shfl_sync(0xffffffffu, dcol, hbase + k);
        acc = fma(ownrow[k], dk, acc);
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        row[j] = top ? row[j] : ownrow[j];               // bottom: Z2[i][:]
        row[N + j] = (l == j) ? 1.0 : 0.0;
      }
      if (top || bot) ws[wTvSv + l] = top ? add[i] + acc : add[i] - acc;
    }
    coop_house_shfl<2 * N, N, N, 2 * N>(row, l, gbase);
    // ---- publish Xi11 (+ inverse diagonal) / Xi21, pre-Xi22
    {
      double lo[N], hi[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        lo[j] = row[j];
        hi[j] = row[N + j];
      }
      if (top) {
        st_row<N>(ws + wX11 + i * N, lo);
        double dg = 1.0;
#pragma unroll
        for (int j = 0; j < N; ++j) dg = (j == i) ? row[j] : dg;
        ws[wInv + i] = rcp_nr(dg);
      } else if (bot) {
        st_row<N>(ws + wX21 + i * N, lo);
        st_row<N>(ws + wBB + i * N, hi);
      }
    }
    __syncwarp();
    // ---- stage C: T1 = Xi11^{-1} U1^T, column c = i per (top) lane; (U1^T)[k][c] = U1[c][k] = ownrow[k]
    {
      double tc[N];
#pragma unroll
      for (int k = 0; k < N; ++k) {
        double s = ownrow[k];
#pragma unroll
        for (int m = 0; m < k; ++m) s = fma(-ws[wX11 + k * N + m], tc[m], s);
        tc[k] = s * ws[wInv + k];
      }
      if (top) {
#pragma unroll
        for (int k = 0; k < N; ++k) ws[wT1 + k * N + i] = tc[k];
        st_row<N>(ws + wT1T + i * N, tc);
      }
    }
    __syncwarp();
    // ---- stage D: top: W[i][:] = sum_k A2[i][k] T1T[k][:] (W = A2 T1^T);  bottom: Pm[i][:] = sum_k Xi21[i][k] T1[k][:]
    double prod[N];
    {
      double coef[N], a2[N], m[N];
      ld_row<N>(e2 + dA + i * N, a2);
#pragma unroll
      for (int k = 0; k < N; ++k) coef[k] = top ? a2[k] : row[k];
      const double* mat = ws + (top ? wT1T : wT1);
#pragma unroll
      for (int j = 0; j < N; ++j) prod[j] = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        ld_row<N>(mat + k * N, m);
#pragma unroll
        for (int j = 0; j < N; ++j) prod[j] = fma(coef[k], m[j], prod[j]);
      }
      // ---- stage E: top: G[i][j] = A2[i][j] - sum_k W[i][k] Xi21[j][k];  b_i = b2_i + G[i][:] tv
      //               bottom: sv2_i = sv_i - Pm[i][:] sv
      double g[N];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        ld_row<N>(ws + wX21 + j * N, m);
        double s = a2[j];
#pragma unroll
        for (int k = 0; k < N; ++k) s = fma(-prod[k], m[k], s);
        g[j] = s;
      }
      const double* vec = ws + wTvSv + (top ? 0 : N);
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) acc = fma(top ? g[k] : prod[k], vec[k], acc);
      if (top) out[db + i] = e2[db + i] + acc;
      else if (bot) ws[wSv2 + i] = vec[i] - acc;
      __syncwarp();
      // ---- stage F: top: A[i][:] = sum_k G[i][k] A1[k][:];  bottom: Zm[i][:] = sum_k A1[k][i] BB[k][:],
      //               eta_i = eta1_i + sum_k A1[k][i] sv2[k]                          _operators.py:70-75
#pragma unroll
      for (int k = 0; k < N; ++k) coef[k] = top ? g[k] : e1[dA + k * N + i];
      const double* mat2 = top ? e1 + dA : ws + wBB;
      double o[N];
#pragma unroll
      for (int j = 0; j < N; ++j) o[j] = 0.0;
      double ea = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        ld_row<N>(mat2 + k * N, m);
#pragma unroll
        for (int j = 0; j < N; ++j) o[j] = fma(coef[k], m[j], o[j]);
        ea = fma(coef[k], ws[wSv2 + k], ea);
      }
      if (top) st_row<N>(out + dA + i * N, o);
      else if (bot) out[de + i] = e1[de + i] + ea;
      // ---- stage G: U = tria([W | U2]) on the top lanes, Z = tria([A1^T BB | Z1]) on the bottom lanes
      double t2[N];
      ld_row<N>((top ? e2 + dU : e1 + dZ) + i * N, t2);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        row[j] = top ? prod[j] : o[j];
        row[N + j] = t2[j];
      }
    }
    coop_house_shfl<2 * N, N, N, N>(row, (top || bot) ? i : N, hbase);
    {
      double lo[N];
#pragma unroll
      for (int j = 0; j < N; ++j) lo[j] = (j <= i) ? row[j] : 0.0;
      if (top) st_row<N>(out + dU + i * N, lo);
      else if (bot) st_row<N>(out + dZ + i * N, lo);
    }
    __syncwarp();
  }
};

// ---------------------------------------------------------------------------------------------
// Smoothing combine (e1 = accumulated LATER side, e2 = earlier side) by GS lanes, one row each:
//   g = E2 g1 + g2 ; E = E2 E1 ; D = tria([E2 D1 | D2])                         _operators.py:118-125
// Dense slot: g [N], E [N][N], D [N][N] (zeros above the diagonal).
// ---------------------------------------------------------------------------------------------
template <int N>
struct CoopS2 {
  static constexpr int NN = N * N;
  static constexpr int G = MidCfg<N>::GS;
  static_assert(N <= G, "a lane group holds the N rows");
  static constexpr int dg = 0, dE = (N + 1) & ~1, dD = dE + NN;
  static constexpr int NFD = spread_stride(dD + NN);
  static constexpr int WS = 0;
  static __host__ __device__ int dense_of(int f) {   // SElem<N>::v index -> slot offset
    if (f < N) return dg + f;
    if (f < N + NN) return dE + (f - N);
    const int t = f - N - NN;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    return dD + i * N + (t - i * (i + 1) / 2);
  }
  static __device__ __forceinline__ double ident(int off) {
    return (off >= dE && off < dE + NN && (off - dE) / N == (off - dE) % N) ? 1.0 : 0.0;
  }
  static __device__ __forceinline__ void combine(const double* __restrict__ e1, const double* __restrict__ e2,
                                                 double* __restrict__ out, double* /*ws*/, const int l, const int gbase) {
    const bool act = l < N;
    const int i = act ? l : 0;
    double coef[N], m[N], eo[N], row[2 * N];
    ld_row<N>(e2 + dE + i * N, coef);
    double gi = e2[dg + i];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      eo[j] = 0.0;
      row[j] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
      gi = fma(coef[k], e1[dg + k], gi);
      ld_row<N>(e1 + dE + k * N, m);
#pragma unroll
      for (int j = 0; j < N; ++j) eo[j] = fma(coef[k], m[j], eo[j]);
      ld_row<N>(e1 + dD + k * N, m);
#pragma unroll
      for (int j = 0; j < N; ++j) row[j] = fma(coef[k], m[j], row[j]);
    }
    ld_row<N>(e2 + dD + i * N, m);
#pragma unroll
    for (int j = 0; j < N; ++j) row[N + j] = m[j];
    coop_house_shfl<2 * N, N, N, N>(row, act ? i : N, gbase);
    if (act) {
      out[dg + i] = gi;
      st_row<N>(out + dE + i * N, eo);
#pragma unroll
      for (int j = 0; j < N; ++j) m[j] = (j <= i) ? row[j] : 0.0;
      st_row<N>(out + dD + i * N, m);
    }
    __syncwarp();
  }
};

// =========================================================================================
// K2 / K4, sub-warp form.  Exclusive scan of the M warp totals of one sequence, two levels in one launch:
//   pass 0 (every CTA): Kogge-Stone over the CTA's group of IT items (lane group x owns item x): in-group
//           exclusive prefixes written back in place, group total to groups[];
//   pass 1 (the CTA that takes the last ticket): the same over the Gc = ceil(M / IT) group totals, in waves of
//           IT items chained by a carry element (ONE wave whenever Gc <= IT, which holds for nx <= 4 at every plan
//           psqrt_capi.cu makes), the sequence total to total_out, the fixed-order sum of the log-likelihood
//           partials, ticket re-armed.
// REV mirrors the item index (suffix scan); groups[] is indexed in scan order.
// Shared memory: two slot buffers [IT][NFD], one carry slot, [IT][WS] workspace, the packed -> dense index table.
// =========================================================================================
template <class OP, int NF, int IT>
constexpr size_t mid2_smem_bytes() {
  return sizeof(double) * (size_t)((2 * IT + 1) * OP::NFD + IT * OP::WS) + sizeof(int) * NF;
}

template <class OP, int NF, int IT, bool REV>
__global__ void __launch_bounds__(IT * OP::G, 1)
k_mid_scan2(double* __restrict__ items, long long M, double* __restrict__ groups, long long Gc,
            unsigned int* __restrict__ counter, double* __restrict__ total_out, const double* __restrict__ ell_part,
            double* __restrict__ ell_out, const PushArgs push) {
  constexpr int G = OP::G;
  constexpr int NFD = OP::NFD;
  extern __shared__ __align__(16) double mid2_sm[];
  __shared__ unsigned int s_ticket;
  double* const slots = mid2_sm;                       // [2][IT][NFD]
  double* const carry = mid2_sm + 2 * IT * NFD;        // inclusive total of the previous waves (pass 1, Gc > IT)
  double* const wsall = carry + NFD;                   // [IT][WS]
  int* const dmap = reinterpret_cast<int*>(wsall + IT * OP::WS);
  const long long seq = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int l = threadIdx.x % G;
  const int gbase = lane - l;
  const int x = threadIdx.x / G;                       // item (lane group) index within the CTA
  double* const ws = wsall + x * OP::WS;
  auto slot = [&](int s, int xx) { return slots + (s * IT + xx) * NFD; };

  PSQ_TRACE(REV, 0);
  for (int f = threadIdx.x; f < NF; f += blockDim.x) dmap[f] = OP::dense_of(f);
  for (int k = threadIdx.x; k < (2 * IT + 1) * NFD; k += blockDim.x) slots[k] = 0.0;   // upper triangles stay zero
  pdl_entry();                                         // shared-memory set-up overlaps the predecessor's tail
  __syncthreads();
  PSQ_TRACE(REV, 1);

#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    double* const arr = pass ? groups : items;
    const long long Mp = pass ? Gc : M;
    double* const base = arr + seq * NF * Mp;
    const long long first = pass ? 0 : (long long)blockIdx.x * IT;
    const long long n_here = pass ? Mp : ((Mp - first < IT) ? Mp - first : IT);   // items this CTA scans in this pass
#pragma unroll 1
    for (long long w0 = 0; w0 < n_here; w0 += IT) {
      const int nw = (int)((n_here - w0 < IT) ? n_here - w0 : IT);   // items of this wave
      const long long sidx = first + w0 + x;           // scan-order index of this lane group's item
      const long long gi = (REV && pass == 0) ? (Mp - 1 - sidx) : sidx;   // groups[] is in scan order already
      const bool have = x < nw;
      {
        // all loads of the lane in flight at once (a rolled load -> store loop would serialise the L2 latencies)
        constexpr int PER = (NF + G - 1) / G;
        double v[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int f = l + q * G;
          v[q] = (have && f < NF) ? __ldcg(base + f * Mp + gi) : 0.0;
        }
        double* d = slot(0, x);
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int f = l + q * G;
          if (f < NF) {
            const int off = dmap[f];
            d[off] = have ? v[q] : OP::ident(off);
          }
        }
      }
      __syncthreads();
      PSQ_TRACE(REV, 2 + 6 * pass);
      int nlev = 0;
      while ((1 << nlev) < nw) ++nlev;
      int cur = 0;
      // Level -1 (later waves only): item 0 <- carry (x) item 0, so that every inclusive prefix of the wave contains
      // the carry.  Same call site as the Kogge-Stone levels: the combine is inlined once and every lane group runs
      // it at every level (groups whose result is not wanted combine their item with itself and discard it).
#pragma unroll 1
      for (int lev = (w0 > 0) ? -1 : 0; lev < nlev; ++lev) {
        const int d = (lev < 0) ? 0 : (1 << lev);
        const bool keep = (lev < 0) ? (x != 0) : (x < d);
        const double* e1 = (lev < 0 && x == 0) ? carry : slot(cur, keep ? x : x - d);
        const double* e2 = slot(cur, x);
        double* o = slot(cur ^ 1, x);
        OP::combine(e1, e2, o, ws, l, gbase);
        if (keep) {
          copy_slot<NFD, G>(o, e2, l);
        }
        __syncthreads();
        cur ^= 1;
      }
      PSQ_TRACE(REV, 3 + 6 * pass);
      // slot(cur, x) = inclusive prefix (carry of earlier waves folded in); exclusive = the previous item's inclusive,
      // for item 0 the carry (identity in the first wave)
      if (have) {
        const double* s = (x > 0) ? slot(cur, x - 1) : carry;
        const bool ident0 = (x == 0) && (w0 == 0);
        for (int f = l; f < NF; f += G) {
          const int off = dmap[f];
          base[f * Mp + gi] = ident0 ? OP::ident(off) : s[off];
        }
      }
      __syncthreads();                                  // the carry has been read by item 0
      if (x == nw - 1) {
        const double* s = slot(cur, x);
        if (w0 + IT < n_here) {                         // more waves follow
          copy_slot<NFD, G>(carry, s, l);
        } else if (pass == 0) {                         // total of the group
          for (int f = l; f < NF; f += G) groups[(seq * NF + f) * Gc + blockIdx.x] = s[dmap[f]];
        } else {                                        // total of the sequence
          if (total_out)
            for (int f = l; f < NF; f += G) total_out[seq * NF + f] = s[dmap[f]];
          if (push.on) {
            // time-sharded run: publish the shard total (+ extras) in every rank's exchange buffer
            const PeerCtx& pc = push.pc;
            const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
            unsigned long long epoch = 0;
            if (l == 0) {
              volatile unsigned long long* c =
                  reinterpret_cast<volatile unsigned long long*>(pc.bufs[pc.rank] + pc.ctr_off + seq);
              epoch = *c + 1ull;
              *c = epoch;
            }
            epoch = __shfl_sync(gmask, epoch, gbase);
            const long long off = pc.data_off + (long long)(epoch & 1ull) * pc.n_ranks * pc.slot +
                                  (long long)pc.rank * pc.slot + seq * pc.payload;
            for (int r = 0; r < pc.n_ranks; ++r) {
              double* dst = pc.bufs[r] + off;
              for (int f = l; f < NF; f += G) dst[f] = s[dmap[f]];
              for (int k = l; k < push.n1; k += G) dst[NF + k] = push.x1[seq * push.s1 + k];
              for (int k = l; k < push.n2; k += G) dst[NF + push.n1 + k] = push.x2[seq * push.s2 + k];
            }
            __threadfence_system();
            __syncwarp(gmask);
            if (l == 0) {
              for (int r = 0; r < pc.n_ranks; ++r) {
                volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
                    pc.bufs[r] + pc.flags_off + (long long)pc.rank * pc.batch + seq);
                *f = epoch;
              }
            }
          }
        }
      }
      __syncthreads();
    }
    PSQ_TRACE(REV, 4 + 6 * pass);
    if (pass == 0) {
      __threadfence();
      __syncthreads();
      PSQ_TRACE(REV, 5);
      if (threadIdx.x == 0) s_ticket = atomicAdd(counter + seq, 1u);
      __syncthreads();
      PSQ_TRACE(REV, 6);
      if (s_ticket != (unsigned int)(Gc - 1)) return;
      __threadfence();
      PSQ_TRACE(REV, 7);
    }
  }
  if (ell_part && threadIdx.x < 32) {
    double sum = 0.0;
    for (long long i2 = lane; i2 < M; i2 += 32) sum += ell_part[seq * M + i2];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
    if (lane == 0) ell_out[seq] = sum;
  }
  if (threadIdx.x == 0) counter[seq] = 0u;
}

}  // namespace psq

namespace psq {

// =========================================================================================
// Time-shard carries (multi-GPU): the carry-in state of a rank from the shard totals of the other ranks.
//   filter  : x0 pushed through the totals of ranks 0 .. rank-1, i.e. the (b, U) part of
//             (0, m0, L0, 0, 0) (x) total_0 (x) ... (x) total_{rank-1}      (the prior as a filtering element)
//   smoother: the (g, D) part of (m_T, 0, L_T) (x) stotal_{R-1} (x) ... (x) stotal_{rank+1}  (the terminal element
//             of parallel/_smoothing.py:56-57 followed by the later shards, reverse scan order)
// One CTA per sequence: the synthetic first element and up to IT - 1 totals per wave are scanned with the sub-warp
// combines above (Kogge-Stone: ceil(log2(count + 1)) dependent combines instead of `count` sequential applies).
// Optionally (flags != nullptr) the kernel first waits until every rank has published its total of this pass:
// see PeerCtx.
// =========================================================================================

template <class OP>
struct CarrySeed;
template <int N>
struct CarrySeed<CoopF2<N>> {   // (0, m, L, 0, 0)
  static __device__ __forceinline__ void fill(double* d, const double* m, const double* L, int t, int nt) {
    using OP = CoopF2<N>;
    for (int k = t; k < OP::NFD; k += nt) d[k] = 0.0;
    __syncthreads();
    for (int k = t; k < N; k += nt) d[OP::db + k] = m[k];
    for (int k = t; k < N * N; k += nt) d[OP::dU + k] = (k % N <= k / N) ? L[k] : 0.0;
  }
  static constexpr int om = CoopF2<N>::db, oL = CoopF2<N>::dU;
};
template <int N>
struct CarrySeed<CoopS2<N>> {   // (m, 0, L)
  static __device__ __forceinline__ void fill(double* d, const double* m, const double* L, int t, int nt) {
    using OP = CoopS2<N>;
    for (int k = t; k < OP::NFD; k += nt) d[k] = 0.0;
    __syncthreads();
    for (int k = t; k < N; k += nt) d[OP::dg + k] = m[k];
    for (int k = t; k < N * N; k += nt) d[OP::dD + k] = (k % N <= k / N) ? L[k] : 0.0;
  }
  static constexpr int om = CoopS2<N>::dg, oL = CoopS2<N>::dD;
};

constexpr int kCarryIT = 16;   // items per wave of the carry scan (seed / running carry + 15 totals)

template <class OP, int NF>
constexpr size_t carry_smem_bytes() {
  return sizeof(double) * (size_t)(2 * kCarryIT * OP::NFD + kCarryIT * OP::WS) + sizeof(int) * NF;
}

// totals: [R][B][NF] packed elements; the `count` totals first, first + step, ... are folded after the seed in that
// order.  m, L: seed state of sequence seq at m + seq * ms, L + seq * Ls.  With a PeerCtx the totals are read from
// this rank's exchange buffer (half = pass parity) after waiting for all ranks, and `mext` / `Lext` (doubles past the
// packed element inside a rank's payload; negative = unused) locate a seed state published by rank `seed_rank`.
template <class OP, int NF, int N>
__global__ void __launch_bounds__(kCarryIT * OP::G, 1)
k_carry_scan(const double* __restrict__ totals, int first, int step, int count, long long B, long long payload,
             const double* __restrict__ m, const double* __restrict__ L, long long ms, long long Ls,
             double* __restrict__ cm, double* __restrict__ cL, const PeerCtx pc, int use_peer, int seed_rank,
             long long mext, long long Lext, const PushArgs push, const double* __restrict__ own_total) {
  constexpr int G = OP::G;
  constexpr int NFD = OP::NFD;
  constexpr int IT = kCarryIT;
  extern __shared__ __align__(16) double carry_sm[];
  double* const slots = carry_sm;                      // [2][IT][NFD]
  double* const wsall = carry_sm + 2 * IT * NFD;       // [IT][WS]
  int* const dmap = reinterpret_cast<int*>(wsall + IT * OP::WS);
  const long long seq = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int l = threadIdx.x % G;
  const int gbase = lane - l;
  const int x = threadIdx.x / G;
  double* const ws = wsall + x * OP::WS;
  auto slot = [&](int s, int xx) { return slots + (s * IT + xx) * NFD; };
  for (int f = threadIdx.x; f < NF; f += blockDim.x) dmap[f] = OP::dense_of(f);
  for (int k = threadIdx.x; k < 2 * IT * NFD; k += blockDim.x) slots[k] = 0.0;
  if (use_peer) {
    if (push.on) {
      // Deferred publication: this rank's own total (own_total [B][NF]) and extras go into every rank's buffer HERE,
      // by the consumer, instead of by the kernel that produced them -- which lets the producer be the smoothing mid
      // scan hidden inside K3 (fused_smooth_mid), whose extras (the shard's last filtered state) exist only when K3
      // has ended.  Same protocol as the push of k_mid_scan2 / k_mid_scan3.
      if (x == 0) {
        const PeerCtx& pp = push.pc;
        const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
        unsigned long long ep = 0;
        if (l == 0) {
          volatile unsigned long long* c =
              reinterpret_cast<volatile unsigned long long*>(pp.bufs[pp.rank] + pp.ctr_off + seq);
          ep = *c + 1ull;
          *c = ep;
        }
        ep = __shfl_sync(gmask, ep, gbase);
        const long long off = pp.data_off + (long long)(ep & 1ull) * pp.n_ranks * pp.slot + (long long)pp.rank * pp.slot +
                              seq * pp.payload;
        for (int r = 0; r < pp.n_ranks; ++r) {
          double* dst = pp.bufs[r] + off;
          for (int f = l; f < NF; f += G) dst[f] = own_total[seq * NF + f];
          for (int k = l; k < push.n1; k += G) dst[NF + k] = push.x1[seq * push.s1 + k];
          for (int k = l; k < push.n2; k += G) dst[NF + push.n1 + k] = push.x2[seq * push.s2 + k];
        }
        __threadfence_system();
        __syncwarp(gmask);
        if (l == 0) {
          for (int r = 0; r < pp.n_ranks; ++r) {
            volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
                pp.bufs[r] + pp.flags_off + (long long)pp.rank * pp.batch + seq);
            *f = ep;
          }
        }
      }
      __syncthreads();
    }
    // every rank has published pass number `epoch` for this sequence (all ranks, not only the ones whose totals
    // are folded: this is the back-edge that keeps any rank from running two passes ahead of a reader)
    const double* mine = pc.bufs[pc.rank];
    const unsigned long long epoch =
        *reinterpret_cast<const volatile unsigned long long*>(mine + pc.ctr_off + seq);
    if ((int)threadIdx.x < pc.n_ranks) {
      const volatile unsigned long long* f =
          reinterpret_cast<const volatile unsigned long long*>(mine + pc.flags_off + (long long)threadIdx.x * pc.batch + seq);
      while (*f < epoch) __nanosleep(20);
    }
    __threadfence_system();
    __syncthreads();
    totals = mine + pc.data_off + (long long)(epoch & 1ull) * pc.n_ranks * pc.slot;
    if (seed_rank >= 0) {
      const double* pr = totals + (long long)seed_rank * pc.slot + seq * payload;
      m = pr + mext - seq * ms;     // so that m + seq * ms below lands on this sequence's seed
      L = pr + Lext - seq * Ls;
    }
  }
  __syncthreads();
  int done = 0;
  int cur = 0;
  CarrySeed<OP>::fill(slot(0, 0), m + seq * ms, L + seq * Ls, threadIdx.x, blockDim.x);
  __syncthreads();
#pragma unroll 1
  while (true) {
    const int nw = (count - done < IT - 1) ? count - done : IT - 1;   // totals in this wave (item 0 = running carry)
    if (x >= 1 && x <= nw) {
      const double* src = totals + ((long long)(first + step * (done + x - 1)) * B + seq) * payload;   // slot == B * payload
      double* d = slot(cur, x);
      constexpr int PER = (NF + G - 1) / G;
      double v[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) v[q] = (l + q * G < NF) ? __ldcg(src + l + q * G) : 0.0;
#pragma unroll
      for (int q = 0; q < PER; ++q)
        if (l + q * G < NF) d[dmap[l + q * G]] = v[q];
    } else if (x > nw) {
      double* d = slot(cur, x);
      for (int k = l; k < NFD; k += G) d[k] = OP::ident(k);
    }
    __syncthreads();
    int nlev = 0;
    while ((1 << nlev) < nw + 1) ++nlev;
#pragma unroll 1
    for (int lev = 0; lev < nlev; ++lev) {
      const int d = 1 << lev;
      const bool keep = x < d;
      const double* e1 = slot(cur, keep ? x : x - d);
      const double* e2 = slot(cur, x);
      double* o = slot(cur ^ 1, x);
      OP::combine(e1, e2, o, ws, l, gbase);
      if (keep) {
        copy_slot<NFD, G>(o, e2, l);
      }
      __syncthreads();
      cur ^= 1;
    }
    done += nw;
    if (done >= count) {
      const double* s = slot(cur, nw);
      for (int k = threadIdx.x; k < N; k += blockDim.x) cm[seq * N + k] = s[CarrySeed<OP>::om + k];
      for (int k = threadIdx.x; k < N * N; k += blockDim.x) cL[seq * N * N + k] = s[CarrySeed<OP>::oL + k];
      break;
    }
    // next wave: the running carry moves to item 0
    {
      const double* s = slot(cur, nw);
      double* d = slot(cur, 0);
      __syncthreads();
      if (x == 0 && nw != 0)
        for (int k = l; k < NFD; k += G) d[k] = s[k];
      __syncthreads();
    }
  }
}

}  // namespace psq

#include <cooperative_groups.h>

namespace psq {

// =========================================================================================
// K2 in CLUSTER form (nx <= 4).  One combine level of k_mid_scan2 is bound by the shared-memory / shuffle
// pipe of ONE SM once more than ~4 warps of lane groups share it (tools/bench_combine.cu: 1.6 us per level
// with <= 2 warps, 1.7 with 4, 3.9 with the 16 warps that 64 items need) while 140 other SMs idle.  Here a group
// of IT = CS x IC items is scanned by a thread-block CLUSTER of CS CTAs, IC items each, on CS different SMs: the
// Kogge-Stone partner of an item may live in another CTA of the cluster and is then read through distributed
// shared memory (copied once into a local staging slot), and the levels are separated by cluster barriers.
// Same contract as k_mid_scan2 with Gc <= IT (one wave per pass); grid = Gc clusters.
// =========================================================================================
template <class OP, int NF, int CS, int IC>
constexpr size_t mid3_smem_bytes() {
  return sizeof(double) * (size_t)(3 * IC * OP::NFD + IC * OP::WS) + sizeof(int) * NF;
}

template <class OP, int NF, int CS, int IC, bool REV>
__global__ void __launch_bounds__(IC * OP::G, 1)
k_mid_scan3(double* __restrict__ items, long long M, double* __restrict__ groups, long long Gc,
            unsigned int* __restrict__ counter, double* __restrict__ total_out, const double* __restrict__ ell_part,
            double* __restrict__ ell_out, const PushArgs push) {
  namespace cg = cooperative_groups;
  constexpr int G = OP::G;
  constexpr int NFD = OP::NFD;
  constexpr int IT = CS * IC;
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();            // CTA within the cluster
  const long long grp = blockIdx.x / CS;               // cluster = group index (scan order)
  extern __shared__ __align__(16) double mid3_sm[];
  __shared__ unsigned int s_ticket;
  double* const slots = mid3_sm;                       // [2][IC][NFD]
  double* const stage = mid3_sm + 2 * IC * NFD;        // [IC][NFD] local copies of remote partners
  double* const wsall = stage + IC * NFD;              // [IC][WS]
  int* const dmap = reinterpret_cast<int*>(wsall + IC * OP::WS);
  const long long seq = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int l = threadIdx.x % G;
  const int gbase = lane - l;
  const int x = threadIdx.x / G;                       // item within the CTA
  const int X = cr * IC + x;                           // item within the cluster
  double* const ws = wsall + x * OP::WS;
  auto slot = [&](int s, int xx) { return slots + (s * IC + xx) * NFD; };

  PSQ_TRACE(REV, 0);
  for (int f = threadIdx.x; f < NF; f += blockDim.x) dmap[f] = OP::dense_of(f);
  for (int k = threadIdx.x; k < 3 * IC * NFD; k += blockDim.x) slots[k] = 0.0;   // upper triangles stay zero
  pdl_entry();                                         // shared-memory set-up overlaps the predecessor's tail
  __syncthreads();
  PSQ_TRACE(REV, 1);

#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    double* const arr = pass ? groups : items;
    const long long Mp = pass ? Gc : M;
    double* const base = arr + seq * NF * Mp;
    const long long first = pass ? 0 : grp * IT;
    const int nw = (int)((Mp - first < IT) ? Mp - first : IT);     // items this cluster scans in this pass
    const long long sidx = first + X;
    const long long gi = (REV && pass == 0) ? (Mp - 1 - sidx) : sidx;
    const bool have = X < nw;
    {
      constexpr int PER = (NF + G - 1) / G;
      double v[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int f = l + q * G;
        v[q] = (have && f < NF) ? __ldcg(base + f * Mp + gi) : 0.0;
      }
      double* d = slot(0, x);
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int f = l + q * G;
        if (f < NF) {
          const int off = dmap[f];
          d[off] = have ? v[q] : OP::ident(off);
        }
      }
    }
    cluster.sync();
    PSQ_TRACE(REV, 2 + 6 * pass);
    int nlev = 0;
    while ((1 << nlev) < nw) ++nlev;
    int cur = 0;
#pragma unroll 1
    for (int lev = 0; lev < nlev; ++lev) {
      const int d = 1 << lev;
      const bool keep = X < d;                          // combines with itself, result discarded
      const int Xp = keep ? X : X - d;
      const int rp = Xp / IC, xp = Xp % IC;
      const double* e1 = slot(cur, xp);
      if (rp != cr) {                                   // partner in another CTA of the cluster: DSMEM -> staging slot
        const double* r = cluster.map_shared_rank(slot(cur, xp), rp);
        double* st = stage + x * NFD;
        copy_slot<NFD, G>(st, r, l);
        e1 = st;
      }
      __syncwarp();
      const double* e2 = slot(cur, x);
      double* o = slot(cur ^ 1, x);
      OP::combine(e1, e2, o, ws, l, gbase);
      if (keep) {
        copy_slot<NFD, G>(o, e2, l);
      }
      cluster.sync();
      cur ^= 1;
    }
    PSQ_TRACE(REV, 3 + 6 * pass);
    // slot(cur, .) = inclusive prefixes; exclusive = the previous item's (identity for item 0), possibly remote
    if (have) {
      const int Xq = (X > 0) ? X - 1 : 0;
      const double* s = (Xq / IC == cr) ? slot(cur, Xq % IC) : cluster.map_shared_rank(slot(cur, Xq % IC), Xq / IC);
      constexpr int PER = (NF + G - 1) / G;
      double v[PER];
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int f = l + q * G;
        if (f < NF) v[q] = (X == 0) ? OP::ident(dmap[f]) : s[dmap[f]];
      }
#pragma unroll
      for (int q = 0; q < PER; ++q) {
        const int f = l + q * G;
        if (f < NF) base[f * Mp + gi] = v[q];
      }
    }
    if (X == nw - 1) {
      const double* s = slot(cur, x);
      if (pass == 0) {                                  // total of the group
        for (int f = l; f < NF; f += G) groups[(seq * NF + f) * Gc + grp] = s[dmap[f]];
      } else {                                          // total of the sequence
        if (total_out)
          for (int f = l; f < NF; f += G) total_out[seq * NF + f] = s[dmap[f]];
        if (push.on) {
          const PeerCtx& pc = push.pc;
          const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
          unsigned long long epoch = 0;
          if (l == 0) {
            volatile unsigned long long* c =
                reinterpret_cast<volatile unsigned long long*>(pc.bufs[pc.rank] + pc.ctr_off + seq);
            epoch = *c + 1ull;
            *c = epoch;
          }
          epoch = __shfl_sync(gmask, epoch, gbase);
          const long long off = pc.data_off + (long long)(epoch & 1ull) * pc.n_ranks * pc.slot +
                                (long long)pc.rank * pc.slot + seq * pc.payload;
          for (int r = 0; r < pc.n_ranks; ++r) {
            double* dst = pc.bufs[r] + off;
            for (int f = l; f < NF; f += G) dst[f] = s[dmap[f]];
            for (int k = l; k < push.n1; k += G) dst[NF + k] = push.x1[seq * push.s1 + k];
            for (int k = l; k < push.n2; k += G) dst[NF + push.n1 + k] = push.x2[seq * push.s2 + k];
          }
          __threadfence_system();
          __syncwarp(gmask);
          if (l == 0) {
            for (int r = 0; r < pc.n_ranks; ++r) {
              volatile unsigned long long* f = reinterpret_cast<volatile unsigned long long*>(
                  pc.bufs[r] + pc.flags_off + (long long)pc.rank * pc.batch + seq);
              *f = epoch;
            }
          }
        }
      }
    }
    PSQ_TRACE(REV, 4 + 6 * pass);
    if (pass == 0) {
      // the cluster that takes the last ticket goes on to scan the group totals
      __threadfence();
      cluster.sync();                                   // all remote reads of this pass are done, all stores fenced
      PSQ_TRACE(REV, 5);
      if (cr == 0 && threadIdx.x == 0) {
        const unsigned int t = atomicAdd(counter + seq, 1u);
        for (int r = 0; r < CS; ++r) *cluster.map_shared_rank(&s_ticket, r) = t;
      }
      cluster.sync();
      PSQ_TRACE(REV, 6);
      if (s_ticket != (unsigned int)(Gc - 1)) return;
      __threadfence();
      PSQ_TRACE(REV, 7);
    }
  }
  if (ell_part && cr == 0 && threadIdx.x < 32) {
    double sum = 0.0;
    for (long long i2 = lane; i2 < M; i2 += 32) sum += ell_part[seq * M + i2];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, d);
    if (lane == 0) ell_out[seq] = sum;
  }
  if (cr == 0 && threadIdx.x == 0) counter[seq] = 0u;
  cluster.sync();                                       // nobody leaves while its shared memory may still be read
}

}  // namespace psq
