// psqrt_kernels.cuh -- the five kernels of one square-root parallel filter + RTS smoother pass.
//
// Time is cut into P chunks of K consecutive steps, one chunk per thread, 32 chunks per warp.
//   K1 filter_reduce   : per chunk, the chunk summary (A,b,U,eta,Z) by the collapsed combine
//                        (psq::filter_reduce_step; also stores the summary with its last step predict-only,
//                        from which K3 builds the chunk's smoothing total); warp-level Kogge-Stone scan of the 32
//                        summaries (generic combine, parsmooth/parallel/_operators.py:43-77);
//                        writes the in-warp exclusive prefix per chunk and one total per warp.
//   K2 mid_scan<FElem> : one CTA per sequence scans the warp totals (exclusive), emits the
//                        sequence total (the element a time-sharded run all-gathers).
//   K3 filter_apply    : carry-in state pushed through the two exclusive prefixes, then a
//                        sequential sqrt Kalman filter inside the chunk; writes filtered
//                        (mean, chol), accumulates the log-likelihood, and (SMOOTH) builds the
//                        smoothing elements from the same triangularisation and reduces them.
//   K4 mid_scan<SElem> : reverse scan of the smoothing warp totals; sums the ell partials.
//   K5 smooth_apply    : carry-in (terminal) state pushed through the suffixes, then the RTS
//                        recursion backwards inside the chunk; writes smoothed (mean, chol).
// The associative-scan seam of the reference (jax.lax.associative_scan at
// parallel/_filtering.py:34-35 and parallel/_smoothing.py:33-34) is K1..K5 together.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "psqrt_math.cuh"
#include "psqrt_async.cuh"
#include "psqrt_coop2.cuh"

namespace psq {

constexpr int kBlock = 128;      // threads per CTA in the sweeps (4 warps)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kYDepth = 4;       // observations are fetched this many steps ahead (psqrt_async.cuh)
constexpr int kXDepth = 2;       // packed filtered states of the backward sweep likewise
// minimum resident CTAs per SM requested from ptxas for the three sweeps (register cap =
// 65536 / (128 * MINB)); tuned on B200, see DESIGN.md section 5
#ifndef PSQ_MINB_K1
#define PSQ_MINB_K1 2
#endif
#ifndef PSQ_MINB_K3
#define PSQ_MINB_K3 2
#endif
#ifndef PSQ_MINB_K5
#define PSQ_MINB_K5 2
#endif

// Linearised SSM as the kernels see it: base pointers + per-step and per-sequence strides in
// doubles (0 = shared by all steps / all sequences).
struct HostFused;   // psqrt_launch.h: fused built-in linearization (psqrt_fused.cuh), host side
struct SSMArgs {
  const double *F, *Q, *bq, *H, *R, *c, *y;
  long long tF, tQ, tb, tH, tR, tc, ty;  // time strides
  long long sF, sQ, sb, sH, sR, sc, sy;  // sequence (batch) strides
  const HostFused* fused;                // HOST pointer, read by the launch code only (nullptr: model arrays above)
};

__device__ __forceinline__ StepPtrs step_ptrs(const SSMArgs& a, long long seq, long long k) {
  StepPtrs p;
  p.F = a.F + seq * a.sF + k * a.tF;
  p.Q = a.Q + seq * a.sQ + k * a.tQ;
  p.bq = a.bq + seq * a.sb + k * a.tb;
  p.H = a.H ? a.H + seq * a.sH + k * a.tH : nullptr;
  p.R = a.R ? a.R + seq * a.sR + k * a.tR : nullptr;
  p.c = a.c ? a.c + seq * a.sc + k * a.tc : nullptr;
  p.y = a.y ? a.y + seq * a.sy + k * a.ty : nullptr;
  return p;
}

// Where a sweep gets the model of step k from: pointers + strides (any model), or the model
// itself carried by value in the kernel parameters (time-invariant models whose entries the
// host knows: they become constant-bank operands, see psq::ModelVals).
struct SrcPtr {
  SSMArgs a;
  __device__ __forceinline__ StepPtrs at(long long seq, long long k) const { return step_ptrs(a, seq, k); }
  __device__ __forceinline__ const double* yp(long long seq, long long k) const { return a.y + seq * a.sy + k * a.ty; }
};
template <int N, int NY>
struct SrcVal {
  ModelVals<N, NY> m;
  const double* y;
  long long ty, sy;
  __device__ __forceinline__ StepVals<N, NY> at(long long seq, long long k) const {
    return StepVals<N, NY>{m, y ? y + seq * sy + k * ty : nullptr};
  }
  __device__ __forceinline__ const double* yp(long long seq, long long k) const { return y + seq * sy + k * ty; }
};
template <int N>
struct SrcValT {  // transition model only (backward sweep)
  ModelValsT<N> m;
  __device__ __forceinline__ StepValsT<N> at(long long, long long) const { return StepValsT<N>{m}; }
};

// ---- element traits: combine(acc, x) with acc = everything earlier in SCAN order ------------
template <class Elem>
struct ScanOp;
template <int N>
struct ScanOp<FElem<N>> {  // forward in time: acc = earlier = elem1
  static __device__ __forceinline__ FElem<N> combine(const FElem<N>& acc, const FElem<N>& x) {
    return filtering_combine<N>(acc, x);
  }
};
template <int N>
struct ScanOp<SElem<N>> {  // backward in time: acc = later = elem1
  static __device__ __forceinline__ SElem<N> combine(const SElem<N>& acc, const SElem<N>& x) {
    return smoothing_combine<N>(acc, x);
  }
};

// SoA scratch: field f of item i of sequence seq lives at buf[(seq * NF + f) * n_items + i]
template <class Elem>
__device__ __forceinline__ void soa_store(double* buf, long long seq, long long n_items, long long i, const Elem& e) {
  double* p = buf + seq * Elem::NF * n_items + i;
#pragma unroll
  for (int f = 0; f < Elem::NF; ++f) p[f * n_items] = e.v[f];
}
template <class Elem>
__device__ __forceinline__ void soa_load(const double* buf, long long seq, long long n_items, long long i, Elem& e) {
  const double* p = buf + seq * Elem::NF * n_items + i;
#pragma unroll
  for (int f = 0; f < Elem::NF; ++f) e.v[f] = p[f * n_items];
}

template <class Elem, bool REV>
__device__ __forceinline__ Elem shfl_elem(const Elem& e, int delta) {
  Elem o;
#pragma unroll
  for (int f = 0; f < Elem::NF; ++f)
    o.v[f] = REV ? __shfl_down_sync(kFull, e.v[f], delta) : __shfl_up_sync(kFull, e.v[f], delta);
  return o;
}

// Inclusive Kogge-Stone scan across the 32 lanes.  REV = false: lane order is scan order;
// REV = true: lane 31 comes first in scan order (suffix scan).
template <class Elem, bool REV>
__device__ __forceinline__ Elem warp_scan_inclusive(Elem e, int lane) {
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Elem o = shfl_elem<Elem, REV>(e, d);
    const bool valid = REV ? (lane + d < 32) : (lane >= d);
    Elem c = ScanOp<Elem>::combine(o, e);
    if (valid) e = c;
  }
  return e;
}

// exclusive = inclusive of the previous lane in scan order (identity for the first lane)
template <class Elem, bool REV>
__device__ __forceinline__ Elem warp_exclusive_from_inclusive(const Elem& incl, int lane) {
  Elem x = shfl_elem<Elem, REV>(incl, 1);
  const bool first = REV ? (lane == 31) : (lane == 0);
  if (first) x.set_identity();
  return x;
}

template <int N>
__device__ __forceinline__ void load_gauss_dense(const double* m, const double* L, Gauss<N>& x) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    x.m[i] = m[i];
#pragma unroll
    for (int j = 0; j <= i; ++j) x.Lc(i, j) = L[i * N + j];
  }
}
template <int N>
__device__ __forceinline__ void store_gauss_dense(double* m, double* L, const Gauss<N>& x) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    m[i] = x.m[i];
#pragma unroll
    for (int j = 0; j < N; ++j) L[i * N + j] = (j <= i) ? x.Lc(i, j) : 0.0;
  }
}

// The filtered states travel from the forward sweep (K3) to the backward sweep (K5) a second time in a
// packed, warp-coalesced scratch array next to the API trajectory: K5 then needs no staging for its
// input.  Layout [B][K][NP][Ppad], NP = N + N(N+1)/2 (mean + lower triangle): slot j of chunk c (the
// filtered state at trajectory index c K + j, i.e. BEFORE step c K + j) keeps field f at
// ((j * NP + f) * Ppad + c), so the 32 lanes of a warp (consecutive chunks, same j) touch 32 consecutive
// doubles -- fully coalesced plain loads / stores.
template <int N>
__device__ __forceinline__ void fpack_store(double* fpack, long long seq, int K, long long Ppad, long long c, int j,
                                            const Gauss<N>& x) {
  constexpr int NP = N + Gauss<N>::TRI;
  double* p = fpack + ((seq * K + j) * NP) * Ppad + c;
#pragma unroll
  for (int f = 0; f < N; ++f) p[f * Ppad] = x.m[f];
#pragma unroll
  for (int f = 0; f < Gauss<N>::TRI; ++f) p[(N + f) * Ppad] = x.L[f];
}
// The observation of the step being processed comes out of the prefetch ring (psqrt_async.cuh); this
// adaptor hands it to the algebra in place of the pointer.
template <class P, int NY>
struct WithY {
  const P& p;
  const double (&yv)[NY];
  template <int N_> __device__ __forceinline__ double fF(int i, int j) const { return p.template fF<N_>(i, j); }
  template <int N_> __device__ __forceinline__ double fQ(int i, int j) const { return p.template fQ<N_>(i, j); }
  __device__ __forceinline__ double fb(int i) const { return p.fb(i); }
  template <int N_> __device__ __forceinline__ double fH(int a, int k) const { return p.template fH<N_>(a, k); }
  template <int NY_> __device__ __forceinline__ double fR(int a, int q) const { return p.template fR<NY_>(a, q); }
  __device__ __forceinline__ double fc(int a) const { return p.fc(a); }
  __device__ __forceinline__ double fy(int a) const { return yv[a]; }
};
// =========================================================================================
// WarpOut: how a sweep writes the API trajectories (mean [*, N], factor [*, N, N]).
// Every lane owns a run of consecutive records, K x REC x 8 bytes away from its neighbour's: written
// directly, one warp-wide store touches 32 different cache lines.  Instead each lane parks the record of
// the current step in the warp's shared-memory tile ([32][RECP], padded against bank conflicts) and the
// warp then copies the tile out together: consecutive lanes write consecutive 16-byte (VEC = 2; even N,
// 16-byte aligned bases) or 8-byte (VEC = 1) words of one record, so each store instruction covers whole
// 32-byte sectors -- for N = 4 a factor record is exactly one 128-byte line.
// (Round-1 history: per-lane TMA bulk copies did this job before.  cp.async.bulk is a warp-uniform
// instruction -- SASS UBLKCP behind a 12-instruction loop over the 32 lanes -- and cost ~380 issue slots
// per step; see DESIGN.md.)
// All 32 lanes of a warp must call put() in lockstep; `active` says whether this lane has a record.
// =========================================================================================
template <int N, int VEC>
struct WarpOut {
  static constexpr int REC = N + N * N;
  // lane stride in doubles: odd number of VEC-words
  static constexpr int RECP = (VEC == 2) ? (((REC / 2) % 2 == 1) ? REC : REC + 2) : ((REC % 2 == 1) ? REC : REC + 1);
  static constexpr size_t smem_bytes(int threads) { return (size_t)threads * RECP * sizeof(double); }
  double* tile;   // this warp's [32][RECP]
  unsigned tile_s;  // its shared-space address
  double *gm, *gL;
  long long c0;   // first chunk of the warp
  long long T;
  int K, voff, lane;
  // lane t writes trajectory index (c0 + t) K + j at step j, provided (c0 + t) K + j + voff < T
  __device__ __forceinline__ WarpOut(unsigned char* smem, double* m, double* L, long long c, int K_, long long T_, int voff_)
      : tile(reinterpret_cast<double*>(smem) + (size_t)(threadIdx.x & ~31) * RECP), gm(m), gL(L),
        c0(c - (threadIdx.x & 31)), T(T_), K(K_), voff(voff_), lane(threadIdx.x & 31) {
    tile_s = (unsigned)__cvta_generic_to_shared(tile);
  }

  // Copy one part (means or factors) of the tile out: W doubles per word, PER words per record.  Word e of the
  // part (e = 32 i + lane) belongs to lane t = e / PER; because the lanes' chunks are consecutive, its global
  // word address is linear in (e, t): B + e + t (K - 1) PER, and so is its address in the padded tile.
  template <int W, int PER>
  __device__ __forceinline__ void copy_part(int j, int nvalid, int src_off, double* g) const {
    const long long B = (c0 * K + j) * PER;
    const long long S = (long long)(K - 1) * PER;
#pragma unroll
    for (int i = 0; i < PER; ++i) {   // 32 * PER words in the tile, 32 per iteration
      const int e = i * 32 + lane;
      const int t = e / PER;
      // explicit PTX: an unconditional shared load (always inside the tile) and ONE predicated global store per
      // word; left to the compiler this became a branch around every word with the address arithmetic redone
      const unsigned sa = tile_s + (unsigned)((e + src_off / W + t * (RECP / W - PER)) * W * 8);
      double* gp = g + (B + e + t * S) * W;
      if (W == 2) {
        double v0, v1;
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(sa));
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %3, %4;\n\t@p st.global.v2.f64 [%0], {%1, %2};\n\t}"
                     ::"l"(gp), "d"(v0), "d"(v1), "r"(t), "r"(nvalid) : "memory");
      } else {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa));
        asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, %3;\n\t@p st.global.f64 [%0], %1;\n\t}"
                     ::"l"(gp), "d"(v), "r"(t), "r"(nvalid) : "memory");
      }
    }
  }
  __device__ __forceinline__ void put(int j, bool active, const Gauss<N>& x) {
    if (active) {
      double* r = tile + lane * RECP;
      if (VEC == 2) {
#pragma unroll
        for (int i = 0; i < N; i += 2) *reinterpret_cast<double2*>(r + i) = make_double2(x.m[i], x.m[i + 1]);
#pragma unroll
        for (int i = 0; i < N; ++i) {
#pragma unroll
          for (int q = 0; q < N; q += 2) {
            const double a = (q <= i) ? x.Lc(i, q) : 0.0;
            const double b = (q + 1 <= i) ? x.Lc(i, q + 1) : 0.0;
            *reinterpret_cast<double2*>(r + N + i * N + q) = make_double2(a, b);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = x.m[i];
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
          for (int q = 0; q < N; ++q) r[N + i * N + q] = (q <= i) ? x.Lc(i, q) : 0.0;
      }
    }
    __syncwarp();
    // lanes with a record at this step form a prefix of the warp (their chunks are consecutive in time)
    const int nvalid = __popc(__ballot_sync(kFull, (c0 + lane) * K + j + voff < T));
    copy_part<VEC, N / VEC>(j, nvalid, 0, gm);
    copy_part<VEC, N * N / VEC>(j, nvalid, N, gL);
    __syncwarp();
  }
};

// =========================================================================================
// K1
// =========================================================================================
template <int N, int NY, class SRC>
__global__ void __launch_bounds__(kBlock, PSQ_MINB_K1)
k_filter_reduce(const __grid_constant__ SRC src, long long T, int K, long long Ppad, double* __restrict__ chunk_own,
                double* __restrict__ chunk_pref, double* __restrict__ warp_tot, unsigned int* __restrict__ counter,
                unsigned int* __restrict__ fuse_ctr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  pdl_entry();
  if (c == 0) {
    counter[seq] = 0u;  // arms the ticket of the mid-level scan that follows
    if (fuse_ctr) {     // and the publish / ticket counters of the smoothing mid scan fused into K3
      fuse_ctr[2 * seq] = 0u;
      fuse_ctr[2 * seq + 1] = 0u;
    }
  }
  FAcc<N> acc;
  acc.set_identity();
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  // observations arrive through a per-lane cp.async ring, kYDepth steps ahead of their use
  LaneRing<NY, kYDepth> yring(smem_raw);
#pragma unroll
  for (int d = 0; d < kYDepth; ++d) yring.issue(d, k0 + d < k1, src.yp(seq, k0 + d), 1);
  int slot = 0;
  double* const own = chunk_own + seq * FElem<N>::NF * Ppad + c;
#pragma unroll 1
  for (long long k = k0; k < k1; ++k) {
    double ycur[NY];
    yring.wait_oldest();
#pragma unroll
    for (int a = 0; a < NY; ++a) ycur[a] = yring.get(slot, a);
    yring.issue(slot, k + kYDepth < k1, src.yp(seq, k + kYDepth), 1);
    slot = (slot + 1 == kYDepth) ? 0 : slot + 1;
    const auto p = src.at(seq, k);
    const bool last = (k + 1 == k1);
    filter_reduce_step<N, NY>(
        acc, WithY<decltype(p), NY>{p, ycur},
        [&](const double (&FA)[N][N], const double (&mp)[N], const double (&Np)[N][2 * N], const FAcc<N>& a) {
          // the chunk's summary with its LAST step predict-only, in FElem order (A, b, U, eta, Z): K3 derives the
          // chunk's smoothing total from it (psq::chunk_smoothing_total)
          if (last) {
            constexpr int TRI = FElem<N>::TRI;
#pragma unroll
            for (int i = 0; i < N; ++i) {
              own[(N * N + i) * Ppad] = mp[i];
              own[(N * N + N + TRI + i) * Ppad] = a.eta[i];
#pragma unroll
              for (int j = 0; j < N; ++j) own[(i * N + j) * Ppad] = FA[i][j];
#pragma unroll
              for (int j = 0; j <= i; ++j) {
                own[(N * N + N + i * (i + 1) / 2 + j) * Ppad] = Np[i][j];
                own[(N * N + 2 * N + TRI + i * (i + 1) / 2 + j) * Ppad] = a.Z[i * (i + 1) / 2 + j];
              }
            }
          }
        });
  }
  FElem<N> own_e;
  acc.to_elem(own_e);
  FElem<N> incl = warp_scan_inclusive<FElem<N>, false>(own_e, lane);
  FElem<N> excl = warp_exclusive_from_inclusive<FElem<N>, false>(incl, lane);
  soa_store(chunk_pref, seq, Ppad, c, excl);
  if (lane == 31) soa_store(warp_tot, seq, Ppad / 32, c / 32, incl);
}


// =========================================================================================
// K2 / K4: exclusive scan of the M warp totals of one sequence, two levels in one launch.
//   level B: one single-warp CTA per group of 32 items (spread over the SMs): Kogge-Stone scan,
//            in-group exclusive prefixes written in place, group total to groups[].
//   level C: the CTA that finishes last (atomic ticket) scans the G = ceil(M / 32) group totals in
//            place (exclusive), writes the sequence total and (optionally) the fixed-order sum of
//            the log-likelihood partials, and re-arms the ticket counter.
// REV mirrors the item index (suffix scan); groups[] is indexed in scan order.
// The counter must be zero on entry: the reduce kernel that produced the items zeroes it.
// =========================================================================================
template <class Elem>
__device__ __forceinline__ void soa_load_cg(const double* buf, long long seq, long long n_items, long long i, Elem& e) {
  const double* p = buf + seq * Elem::NF * n_items + i;
#pragma unroll
  for (int f = 0; f < Elem::NF; ++f) e.v[f] = __ldcg(p + f * n_items);
}

// The mid-level scans are pure latency: a dozen dependent combines by a handful of warps while the rest
// of the GPU idles.  Each combine is therefore split over the R warps of the CTA by OUTPUT field: every
// warp holds the same 32 elements, runs its own inlined copy of the combine from which the compiler
// removes everything its fields do not need (the three triangularisations of the filtering combine end up
// in three different warps), and the pieces are exchanged through shared memory.
template <class Elem>
struct Split;
template <int N>
struct Split<FElem<N>> {  // 0: (b, U)   1: Z   2: (A, eta)
  static constexpr int R = 3;
  static __host__ __device__ constexpr int owner(int f) {
    constexpr int TRI = N * (N + 1) / 2;
    return f < N * N ? 2 : (f < N * N + N + TRI ? 0 : (f < N * N + 2 * N + TRI ? 2 : 1));
  }
};
template <int N>
struct Split<SElem<N>> {  // the smoothing combine is small: one warp, inlined (measured: splitting it loses)
  static constexpr int R = 1;
  static __host__ __device__ constexpr int owner(int) { return 0; }
};

// out = combine(a, b) computed by the R warps together; every warp gets the full result.  Must be called
// by the whole CTA.  sm: [NF][32] doubles.  Deliberately not inlined (see k_mid_scan).
template <class Elem>
__device__ __noinline__ void combine_split_impl(const Elem& a, const Elem& b, Elem& out, double* sm, int role, int lane) {
#pragma unroll
  for (int r = 0; r < Split<Elem>::R; ++r) {
    if (role == r) {
      const Elem t = ScanOp<Elem>::combine(a, b);
#pragma unroll
      for (int f = 0; f < Elem::NF; ++f)
        if (Split<Elem>::owner(f) == r) sm[f * 32 + lane] = t.v[f];
    }
  }
  __syncthreads();
#pragma unroll
  for (int f = 0; f < Elem::NF; ++f) out.v[f] = sm[f * 32 + lane];
  __syncthreads();
}
template <class Elem>
__device__ __forceinline__ void combine_split(const Elem& a, const Elem& b, Elem& out, double* sm, int role, int lane) {
  if constexpr (Split<Elem>::R == 1) out = ScanOp<Elem>::combine(a, b);
  else combine_split_impl<Elem>(a, b, out, sm, role, lane);
}

// Straight-line code that runs once is expensive here: a cold instruction fetch costs ~0.2 us per KB
// (ncu: the last CTA's extra inlined copies of the combine took longer than the five warm Kogge-Stone
// steps before them).  combine_split is therefore NOT inlined: level B and the three phases of level C
// all run the same (per-role) copy of the combine, which is warm in the instruction cache after the first
// Kogge-Stone step.
template <class Elem>
__device__ __forceinline__ Elem scan_inclusive_split(Elem e, double* sm, int role, int lane) {
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    const Elem o = shfl_elem<Elem, false>(e, d);
    Elem c;
    combine_split<Elem>(o, e, c, sm, role, lane);
    if (lane >= d) e = c;
  }
  return e;
}

template <class Elem, bool REV>
__global__ void __launch_bounds__(32 * Split<Elem>::R, 1)
k_mid_scan(double* __restrict__ items, long long M, double* __restrict__ groups, long long G,
           unsigned int* __restrict__ counter, double* __restrict__ total_out,
           const double* __restrict__ ell_part, double* __restrict__ ell_out) {
  __shared__ double sm[Split<Elem>::R > 1 ? Elem::NF * 32 : 1];
  __shared__ unsigned int s_ticket;
  const long long seq = blockIdx.y;
  const long long g = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;
  {
    const long long sidx = g * 32 + lane;
    const long long i = REV ? (M - 1 - sidx) : sidx;
    Elem e;
    e.set_identity();
    if (sidx < M) soa_load(items, seq, M, i, e);
    Elem incl = scan_inclusive_split<Elem>(e, sm, role, lane);
    Elem excl = warp_exclusive_from_inclusive<Elem, false>(incl, lane);
    if (role == 0) {
      if (sidx < M) soa_store(items, seq, M, i, excl);
      if (lane == 31) soa_store(groups, seq, G, g, incl);
    }
  }
  if (role == 0) {
    __threadfence();
    if (lane == 0) s_ticket = atomicAdd(counter + seq, 1u);
  }
  __syncthreads();
  if (s_ticket != (unsigned int)(G - 1)) return;
  __threadfence();
  // level C: this CTA finished last; lane l owns the q consecutive group totals [l q, l q + q)
  const int q = (int)((G + 31) / 32);
  const long long s0 = (long long)lane * q;
  Elem acc;
  acc.set_identity();
#pragma unroll 1
  for (int i = 0; i < q; ++i) {
    Elem x;
    x.set_identity();
    if (s0 + i < G) soa_load_cg(groups, seq, G, s0 + i, x);
    if (i == 0) {
      acc = x;
    } else {
      Elem c;
      combine_split<Elem>(acc, x, c, sm, role, lane);
      if (s0 + i < G) acc = c;
    }
  }
  Elem incl = scan_inclusive_split<Elem>(acc, sm, role, lane);
  Elem run = warp_exclusive_from_inclusive<Elem, false>(incl, lane);
  if (role == 0 && lane == 31 && total_out) {
#pragma unroll
    for (int f = 0; f < Elem::NF; ++f) total_out[seq * Elem::NF + f] = incl.v[f];
  }
#pragma unroll 1
  for (int i = 0; i < q; ++i) {
    Elem x;
    x.set_identity();
    if (s0 + i < G) soa_load_cg(groups, seq, G, s0 + i, x);
    __syncthreads();   // every warp has read the total before warp 0 overwrites it with the prefix
    if (role == 0 && s0 + i < G) soa_store(groups, seq, G, s0 + i, run);
    if (i + 1 < q) {
      Elem c;
      combine_split<Elem>(run, x, c, sm, role, lane);
      run = c;
    }
  }
  if (ell_part && role == 0) {
    double sum = 0.0;
    for (long long i = lane; i < M; i += 32) sum += ell_part[seq * M + i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(kFull, sum, d);
    if (lane == 0) ell_out[seq] = sum;
  }
  if (threadIdx.x == 0) counter[seq] = 0u;
}

// =========================================================================================
// K4 fused into K3.  The smoothing warp totals exist ~20 us into K3 (right after the once-per-chunk prologue) while
// K3 goes on for another ~65 us filtering through its chunks, and the sweeps leave a few CTA slots of the GPU empty
// (290 CTAs on 148 x 2 slots at T = 1e6).  A handful of extra CTAs appended to K3's grid wait until every worker warp
// has published its total, then run the two-level suffix scan of K4 (one warp per group of IT totals, the warp that
// takes the last ticket scans the group totals) concurrently with the workers' step loops: the mid-scan latency
// leaves the critical path and K4's launch disappears.  Workers never wait for the scan CTAs, so the scheme cannot
// deadlock whatever the block scheduler does; if the extra CTAs only get a slot when workers retire, the scan simply
// starts late.  ctr[0]: published warp totals, ctr[1]: ticket; both zeroed by K1 of the same pass.
// =========================================================================================
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int N>
__device__ __noinline__ void fused_smooth_mid(double* __restrict__ items, long long M, double* __restrict__ groups,
                                              unsigned int* __restrict__ ctr, double* __restrict__ total_out,
                                              long long seq, int sw, int nsw) {
  using Elem = SElem<N>;
  constexpr int IT = MidCfg<N>::IT;
  constexpr int PER = IT / 32;   // consecutive items (scan order) per lane
  const int lane = threadIdx.x & 31;
  const long long Gc = (M + IT - 1) / IT;
  if (lane == 0) {
    while (ld_acquire_u32(ctr) < (unsigned int)M) __nanosleep(256);
  }
  __syncwarp();
#pragma unroll 1
  for (long long g = sw; g < Gc; g += nsw) {
    Elem it[PER];
    Elem acc;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const long long sidx = g * IT + (long long)lane * PER + q;
      it[q].set_identity();
      if (sidx < M) soa_load_cg(items, seq, M, M - 1 - sidx, it[q]);
      if (q == 0) {
        acc = it[0];
      } else {
        Elem c2 = ScanOp<Elem>::combine(acc, it[q]);
        if (sidx < M) acc = c2;
      }
    }
    Elem incl = warp_scan_inclusive<Elem, false>(acc, lane);
    Elem run = warp_exclusive_from_inclusive<Elem, false>(incl, lane);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const long long sidx = g * IT + (long long)lane * PER + q;
      if (sidx < M) soa_store(items, seq, M, M - 1 - sidx, run);
      if (q + 1 < PER) run = ScanOp<Elem>::combine(run, it[q]);
    }
    if (lane == 31) soa_store(groups, seq, Gc, g, incl);
  }
  __threadfence();
  unsigned int ticket = 0u;
  if (lane == 0) ticket = atomicAdd(ctr + 1, 1u);
  ticket = __shfl_sync(kFull, ticket, 0);
  if (ticket != (unsigned int)(nsw - 1)) return;
  __threadfence();
  // this warp finished last: exclusive scan of the Gc group totals, lane l owning q consecutive ones
  const int q = (int)((Gc + 31) / 32);
  const long long s0 = (long long)lane * q;
  Elem acc;
  acc.set_identity();
#pragma unroll 1
  for (int i = 0; i < q; ++i) {
    Elem x;
    x.set_identity();
    if (s0 + i < Gc) soa_load_cg(groups, seq, Gc, s0 + i, x);
    if (i == 0) {
      acc = x;
    } else {
      Elem c2 = ScanOp<Elem>::combine(acc, x);
      if (s0 + i < Gc) acc = c2;
    }
  }
  Elem incl = warp_scan_inclusive<Elem, false>(acc, lane);
  Elem run = warp_exclusive_from_inclusive<Elem, false>(incl, lane);
  if (lane == 31 && total_out) {
#pragma unroll
    for (int f = 0; f < Elem::NF; ++f) total_out[seq * Elem::NF + f] = incl.v[f];
  }
#pragma unroll 1
  for (int i = 0; i < q; ++i) {
    Elem x;
    x.set_identity();
    if (s0 + i < Gc) {
      soa_load_cg(groups, seq, Gc, s0 + i, x);
      soa_store(groups, seq, Gc, s0 + i, run);
    }
    if (i + 1 < q) run = ScanOp<Elem>::combine(run, x);
  }
}

// =========================================================================================
// K3
// =========================================================================================
template <int N, int NY, bool SMOOTH, bool LOGLIK, class SRC, class OUT>
__global__ void __launch_bounds__(kBlock, PSQ_MINB_K3)
k_filter_apply(const __grid_constant__ SRC src, long long T, int K, long long Ppad,
               const double* __restrict__ carry_m, const double* __restrict__ carry_L,  // [B][N], [B][N][N] lower
               const double* __restrict__ chunk_own, const double* __restrict__ chunk_pref,
               const double* __restrict__ warp_pref, const double* __restrict__ group_pref,
               double* __restrict__ fm, double* __restrict__ fL,  // [B][T+1][N], [B][T+1][N][N]; index k+1 written
               double* __restrict__ chunk_suf, double* __restrict__ warp_stot, double* __restrict__ ell_part,
               unsigned int* __restrict__ counter, double* __restrict__ fpack,
               // fused smoothing mid scan (fused_smooth_mid above): n_work = worker CTAs in x (0: not fused; the CTAs
               // beyond them run the scan), its group array, sequence total (or null) and counters
               int n_work, double* __restrict__ group_s, double* __restrict__ stotal,
               unsigned int* __restrict__ fuse_ctr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long long Mw = Ppad / 32;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;

  pdl_entry();
  if constexpr (SMOOTH) {
    if (n_work > 0 && (int)blockIdx.x >= n_work) {
      const int nsw = ((int)gridDim.x - n_work) * (kBlock / 32);
      const int sw = ((int)blockIdx.x - n_work) * (kBlock / 32) + (threadIdx.x >> 5);
      fused_smooth_mid<N>(warp_stot, Mw, group_s, fuse_ctr + 2 * seq, stotal, seq, sw, nsw);
      return;
    }
  }
  if (SMOOTH && c == 0) counter[seq] = 0u;
  LaneRing<NY, kYDepth> yring(smem_raw + OUT::smem_bytes(kBlock));
#pragma unroll
  for (int d = 0; d < kYDepth; ++d) yring.issue(d, k0 + d < k1, src.yp(seq, k0 + d), 1);
  Gauss<N> x;
  load_gauss_dense<N>(carry_m + seq * N, carry_L + seq * N * N, x);
  // carry pushed through the three exclusive prefixes (group, warp, chunk).  A loop, not three inlined
  // copies: code that runs once per kernel is paid for in cold instruction fetches (see k_mid_scan).
#pragma unroll 1
  for (int lvl = 0; lvl < 3; ++lvl) {
    constexpr int IT = MidCfg<N>::IT;   // warps per group of the mid-level scan
    const double* buf = (lvl == 0) ? group_pref : (lvl == 1) ? warp_pref : chunk_pref;
    const long long n_items = (lvl == 0) ? (Mw + IT - 1) / IT : (lvl == 1) ? Mw : Ppad;
    const long long idx = (lvl == 0) ? c / (32 * IT) : (lvl == 1) ? c / 32 : c;
    FElem<N> e;
    soa_load(buf, seq, n_items, idx, e);
    filtering_apply<N>(x, e);
  }
  double* fmS = fm + seq * (T + 1) * N;
  double* fLS = fL + seq * (T + 1) * N * N;
  if (c == 0) store_gauss_dense<N>(fmS, fLS, x);  // trajectory index 0 = carry-in state
  if (SMOOTH) {
    // the chunk's smoothing total from its filtering summary and the state it starts from, then the
    // suffix scan across the warp: nothing of the smoother stays live in the step loop below
    SElem<N> sacc;
    sacc.set_identity();
    if (k0 < k1) {
      FElem<N> e;
      soa_load(chunk_own, seq, Ppad, c, e);
      chunk_smoothing_total<N>(x, e, sacc);
    }
    SElem<N> incl = warp_scan_inclusive<SElem<N>, true>(sacc, lane);
    SElem<N> excl = warp_exclusive_from_inclusive<SElem<N>, true>(incl, lane);
    soa_store(chunk_suf, seq, Ppad, c, excl);
    if (lane == 0) {
      soa_store(warp_stot, seq, Mw, c / 32, incl);
      if (n_work > 0) {   // publish to the scan CTAs
        __threadfence();
        atomicAdd(fuse_ctr + 2 * seq, 1u);
      }
    }
  }
  // The step loop runs K + 1 times in every lane of a warp that has any work (the cooperative writes
  // need the whole warp); lanes past the end of the sequence idle through it.  Inside the chunk the
  // recursion carries a dense square-root factor and iteration j advances it by step k0 + j WHILE it
  // triangularises and writes out the state at index k0 + j (psq::kalman_step_dense).
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  OUT out(smem_raw, fmS, fLS, c, K, T, -1);
  double ell = 0.0;
  int slot = 0;
  if ((c - lane) * K < T) {
    GaussD<N> xd;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      xd.m[i] = x.m[i];
#pragma unroll
      for (int q = 0; q < N; ++q) xd.Y[i][q] = (q <= i) ? x.Lc(i, q) : 0.0;
    }
#pragma unroll 1
    for (int j = 0; j <= K; ++j) {
      if (j < len) {
        const long long k = k0 + j;
        double ycur[NY];
        yring.wait_oldest();
#pragma unroll
        for (int a = 0; a < NY; ++a) ycur[a] = yring.get(slot, a);
        yring.issue(slot, j + kYDepth < len, src.yp(seq, k + kYDepth), 1);
        slot = (slot + 1 == kYDepth) ? 0 : slot + 1;
        const auto p = src.at(seq, k);
        ell += kalman_step_dense<N, NY, LOGLIK>(xd, WithY<decltype(p), NY>{p, ycur}, x);
        if (SMOOTH) fpack_store<N>(fpack, seq, K, Ppad, c, j, x);  // filtered state at index k
      } else if (j == len) {
        gaussd_tri<N>(xd, x);
      }
      if (j > 0) out.put(j, j <= len, x);  // trajectory index k0 + j
    }
  }
  if (LOGLIK) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ell += __shfl_down_sync(kFull, ell, d);
    if (lane == 0) ell_part[seq * Mw + c / 32] = ell;
  }
}

// Standalone smoothing reduce (smoothing(...) called on an existing filter trajectory): no filtering
// summaries exist, so the chunk totals are the ordered combine of the per-step elements.
template <int N, class SRC>
__global__ void __launch_bounds__(kBlock)
k_smooth_reduce(const __grid_constant__ SRC src, long long T, int K, long long Ppad, const double* __restrict__ fm,
                const double* __restrict__ fL, double* __restrict__ chunk_suf, double* __restrict__ warp_stot,
                unsigned int* __restrict__ counter, double* __restrict__ fpack) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (c == 0) counter[seq] = 0u;
  const long long Mw = Ppad / 32;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  const double* fmS = fm + seq * (T + 1) * N;
  const double* fLS = fL + seq * (T + 1) * N * N;
  SElem<N> sacc;
  sacc.set_identity();
#pragma unroll 1
  for (long long k = k0; k < k1; ++k) {
    const auto p = src.at(seq, k);
    Gauss<N> xf;
    load_gauss_dense<N>(fmS + k * N, fLS + k * N * N, xf);
    fpack_store<N>(fpack, seq, K, Ppad, c, (int)(k - k0), xf);
    SElem<N> se;
    smoothing_element<N>(xf, p, se);
    sacc = (k == k0) ? se : smoothing_combine<N>(se, sacc);
  }
  SElem<N> incl = warp_scan_inclusive<SElem<N>, true>(sacc, lane);
  SElem<N> excl = warp_exclusive_from_inclusive<SElem<N>, true>(incl, lane);
  soa_store(chunk_suf, seq, Ppad, c, excl);
  if (lane == 0) soa_store(warp_stot, seq, Mw, c / 32, incl);
}

// =========================================================================================
// K5
// =========================================================================================
template <int N, class SRC, class OUT>
__global__ void __launch_bounds__(kBlock, PSQ_MINB_K5)
k_smooth_apply(const __grid_constant__ SRC src, long long T, int K, long long Ppad,
               const double* __restrict__ carry_m, const double* __restrict__ carry_L,  // smoothed state at index T
               long long carry_mstride, long long carry_Lstride,
               const double* __restrict__ chunk_suf, const double* __restrict__ warp_suf,
               const double* __restrict__ group_suf, const double* __restrict__ fpack,
               double* __restrict__ sm, double* __restrict__ sL, int write_terminal) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long long Mw = Ppad / 32;
  const long long k0 = c * K;
  const long long k1 = (k0 + K < T) ? k0 + K : T;
  pdl_entry();
  if ((c - lane) * K >= T) return;  // the whole warp lies past the end of the sequence
  const int len = (k1 > k0) ? (int)(k1 - k0) : 0;
  double* smS = sm + seq * (T + 1) * N;
  double* sLS = sL + seq * (T + 1) * N * N;
  constexpr int NP = N + Gauss<N>::TRI;
  LaneRing<NP, kXDepth> xring(smem_raw + OUT::smem_bytes(kBlock));
  const double* fp = fpack + (seq * K * NP) * Ppad + c;  // slot j of this chunk: fp + j NP Ppad, field stride Ppad
#pragma unroll
  for (int d = 0; d < kXDepth; ++d) xring.issue(d, len - 1 - d >= 0, fp + (long long)(len - 1 - d) * NP * Ppad, Ppad);
  Gauss<N> xs;
  if (len > 0) {
    load_gauss_dense<N>(carry_m + seq * carry_mstride, carry_L + seq * carry_Lstride, xs);
    if (write_terminal && k1 == T) store_gauss_dense<N>(smS + T * N, sLS + T * N * N, xs);
#pragma unroll 1
    for (int lvl = 0; lvl < 3; ++lvl) {   // terminal state pushed through the three exclusive suffixes
      constexpr int IT = MidCfg<N>::IT;
      const double* buf = (lvl == 0) ? group_suf : (lvl == 1) ? warp_suf : chunk_suf;
      const long long n_items = (lvl == 0) ? (Mw + IT - 1) / IT : (lvl == 1) ? Mw : Ppad;
      const long long idx = (lvl == 0) ? (Mw - 1 - c / 32) / IT : (lvl == 1) ? c / 32 : c;  // groups: scan (reverse) order
      SElem<N> e;
      soa_load(buf, seq, n_items, idx, e);
      smoothing_apply<N>(xs, e);
    }
  }
  OUT out(smem_raw, smS, sLS, c, K, T, 0);
  int slot = 0;
#pragma unroll 1
  for (int j = K - 1; j >= 0; --j) {
    const bool active = j < len;
    if (active) {
      Gauss<N> xf;
      xring.wait_oldest();
#pragma unroll
      for (int f = 0; f < N; ++f) xf.m[f] = xring.get(slot, f);
#pragma unroll
      for (int f = 0; f < Gauss<N>::TRI; ++f) xf.L[f] = xring.get(slot, N + f);
      xring.issue(slot, j - kXDepth >= 0, fp + (long long)(j - kXDepth) * NP * Ppad, Ppad);
      slot = (slot + 1 == kXDepth) ? 0 : slot + 1;
      rts_step<N>(xs, xf, src.at(seq, k0 + j));
    }
    out.put(j, active, xs);  // trajectory index k0 + j
  }
}

// =========================================================================================
// Time-shard carries (multi-GPU): fold the all-gathered shard totals of the ranks before
// (filter) / after (smoother) this one into the carry-in state.  One thread per sequence.
// =========================================================================================
template <int N>
__global__ void k_carry_filter(const double* __restrict__ totals /*[R][B][NF]*/, int rank, long long B,
                               const double* __restrict__ m0, const double* __restrict__ L0,
                               double* __restrict__ cm, double* __restrict__ cL) {
  const long long seq = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (seq >= B) return;
  Gauss<N> x;
  load_gauss_dense<N>(m0 + seq * N, L0 + seq * N * N, x);
#pragma unroll 1
  for (int r = 0; r < rank; ++r) {
    FElem<N> e;
    const double* p = totals + ((long long)r * B + seq) * FElem<N>::NF;
#pragma unroll
    for (int f = 0; f < FElem<N>::NF; ++f) e.v[f] = p[f];
    filtering_apply<N>(x, e);
  }
  store_gauss_dense<N>(cm + seq * N, cL + seq * N * N, x);
}

template <int N>
__global__ void k_carry_smoother(const double* __restrict__ totals /*[R][B][NF]*/, int rank, int R, long long B,
                                 const double* __restrict__ mT, const double* __restrict__ LT,
                                 double* __restrict__ cm, double* __restrict__ cL) {
  const long long seq = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (seq >= B) return;
  Gauss<N> x;
  load_gauss_dense<N>(mT + seq * N, LT + seq * N * N, x);
#pragma unroll 1
  for (int r = R - 1; r > rank; --r) {
    SElem<N> e;
    const double* p = totals + ((long long)r * B + seq) * SElem<N>::NF;
#pragma unroll
    for (int f = 0; f < SElem<N>::NF; ++f) e.v[f] = p[f];
    smoothing_apply<N>(x, e);
  }
  store_gauss_dense<N>(cm + seq * N, cL + seq * N * N, x);
}

// =========================================================================================
// Element-level kernels (the reference's own seams, one thread per time step)
// =========================================================================================
template <int N, int NY>
__global__ void k_filter_elements(SSMArgs a, long long T, long long B, const double* __restrict__ m0,
                                  const double* __restrict__ L0, double* __restrict__ A, double* __restrict__ b,
                                  double* __restrict__ U, double* __restrict__ eta, double* __restrict__ Z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * B) return;
  const long long seq = i / T, k = i % T;
  StepPtrs p = step_ptrs(a, seq, k);
  const bool first = (k == 0) && m0 != nullptr;
  filtering_element<N, NY>(p, first ? m0 + seq * N : nullptr, first ? L0 + seq * N * N : nullptr,
                           A + i * N * N, b + i * N, U + i * N * N, eta + i * N, Z + i * N * N);
}

template <int N>
__global__ void k_smoother_elements(SSMArgs a, long long T, long long B, const double* __restrict__ fm,
                                    const double* __restrict__ fL, double* __restrict__ g, double* __restrict__ E,
                                    double* __restrict__ D) {
  // T+1 elements per sequence; the last one is (m_T, 0, L_T)               _smoothing.py:56-57
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (T + 1) * B) return;
  const long long seq = i / (T + 1), k = i % (T + 1);
  Gauss<N> xf;
  load_gauss_dense<N>(fm + i * N, fL + i * N * N, xf);
  SElem<N> se;
  if (k < T) {
    StepPtrs p = step_ptrs(a, seq, k);
    smoothing_element<N>(xf, p, se);
  } else {
#pragma unroll
    for (int r = 0; r < N; ++r) {
      se.g(r) = xf.m[r];
#pragma unroll
      for (int q = 0; q < N; ++q) se.E(r, q) = 0.0;
#pragma unroll
      for (int q = 0; q <= r; ++q) se.D(r, q) = xf.Lc(r, q);
    }
  }
#pragma unroll
  for (int r = 0; r < N; ++r) {
    g[i * N + r] = se.g(r);
#pragma unroll
    for (int q = 0; q < N; ++q) {
      E[i * N * N + r * N + q] = se.E(r, q);
      D[i * N * N + r * N + q] = (q <= r) ? se.D(r, q) : 0.0;
    }
  }
}

template <int N, int NY>
__global__ void k_loglik_terms(SSMArgs a, long long T, long long B, const double* __restrict__ fm,
                               const double* __restrict__ fL, double* __restrict__ terms) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * B) return;
  const long long seq = i / T, k = i % T;
  StepPtrs p = step_ptrs(a, seq, k);
  const long long j = seq * (T + 1) + k;  // filtered state at k (i.e. before step k)
  terms[i] = loglik_term<N, NY>(p, fm + j * N, fL + j * N * N);
}

// Generic element scans: inputs are reference-layout dense arrays (AoS per step).
// Square-root factors arriving through the element-level seams may be ANY n x n factor (the
// reference's raw elements carry Z = [Z0 | 0], parallel/_filtering.py:141-142, which is not
// triangular): they are triangularised on load, which leaves F F^T unchanged.
template <int N>
__device__ __forceinline__ void load_factor_tria(const double* M, double (&Lw)[N][N]) {
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int q = 0; q < N; ++q) Lw[r][q] = M[r * N + q];
  house_rows<N, N, N - 1>(Lw);
}
template <int N>
__device__ __forceinline__ void load_felem_dense(const double* A, const double* b, const double* U, const double* eta,
                                                 const double* Z, long long i, FElem<N>& e) {
  double Uw[N][N], Zw[N][N];
  load_factor_tria<N>(U + i * N * N, Uw);
  load_factor_tria<N>(Z + i * N * N, Zw);
#pragma unroll
  for (int r = 0; r < N; ++r) {
    e.b(r) = b[i * N + r];
    e.eta(r) = eta[i * N + r];
#pragma unroll
    for (int q = 0; q < N; ++q) e.A(r, q) = A[i * N * N + r * N + q];
#pragma unroll
    for (int q = 0; q <= r; ++q) {
      e.U(r, q) = Uw[r][q];
      e.Z(r, q) = Zw[r][q];
    }
  }
}
template <int N>
__device__ __forceinline__ void load_selem_dense(const double* g, const double* E, const double* D, long long i,
                                                 SElem<N>& e) {
  double Dw[N][N];
  load_factor_tria<N>(D + i * N * N, Dw);
#pragma unroll
  for (int r = 0; r < N; ++r) {
    e.g(r) = g[i * N + r];
#pragma unroll
    for (int q = 0; q < N; ++q) e.E(r, q) = E[i * N * N + r * N + q];
#pragma unroll
    for (int q = 0; q <= r; ++q) e.D(r, q) = Dw[r][q];
  }
}

template <int N>
__global__ void __launch_bounds__(kBlock)
k_escan_filter_reduce(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                      long long T, int K, long long Ppad, double* chunk_pref, double* warp_tot,
                      unsigned int* counter) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (c == 0) counter[seq] = 0u;
  const long long k0 = c * K, k1 = (k0 + K < T) ? k0 + K : T;
  FElem<N> acc;
  acc.set_identity();
#pragma unroll 1
  for (long long k = k0; k < k1; ++k) {
    FElem<N> e;
    load_felem_dense<N>(A, b, U, eta, Z, seq * T + k, e);
    if (k == k0) acc = e; else acc = filtering_combine<N>(acc, e);
  }
  FElem<N> incl = warp_scan_inclusive<FElem<N>, false>(acc, lane);
  FElem<N> excl = warp_exclusive_from_inclusive<FElem<N>, false>(incl, lane);
  soa_store(chunk_pref, seq, Ppad, c, excl);
  if (lane == 31) soa_store(warp_tot, seq, Ppad / 32, c / 32, incl);
}

// has_carry = 0: plain inclusive scan of the given elements (outputs (b, U) of each prefix).
template <int N>
__global__ void __launch_bounds__(kBlock)
k_escan_filter_apply(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                     long long T, int K, long long Ppad, const double* carry_m, const double* carry_L,
                     const double* chunk_pref, const double* warp_pref, const double* group_pref, double* om,
                     double* oL) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const long long k0 = c * K, k1 = (k0 + K < T) ? k0 + K : T;
  if (k0 >= k1) return;
  Gauss<N> x;
  if (carry_m) {
    load_gauss_dense<N>(carry_m + seq * N, carry_L + seq * N * N, x);
  } else {
#pragma unroll
    for (int r = 0; r < N; ++r) {
      x.m[r] = 0.0;
#pragma unroll
      for (int q = 0; q <= r; ++q) x.Lc(r, q) = 0.0;
    }
  }
  FElem<N> e;
  soa_load(group_pref, seq, (Ppad / 32 + MidCfg<N>::IT - 1) / MidCfg<N>::IT, c / (32 * MidCfg<N>::IT), e);
  filtering_apply<N>(x, e);
  soa_load(warp_pref, seq, Ppad / 32, c / 32, e);
  filtering_apply<N>(x, e);
  soa_load(chunk_pref, seq, Ppad, c, e);
  filtering_apply<N>(x, e);
#pragma unroll 1
  for (long long k = k0; k < k1; ++k) {
    load_felem_dense<N>(A, b, U, eta, Z, seq * T + k, e);
    filtering_apply<N>(x, e);
    store_gauss_dense<N>(om + (seq * T + k) * N, oL + (seq * T + k) * N * N, x);
  }
}

template <int N>
__global__ void __launch_bounds__(kBlock)
k_escan_smooth_reduce(const double* g, const double* E, const double* D, long long T, int K, long long Ppad,
                      double* chunk_suf, double* warp_stot, unsigned int* counter) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (c == 0) counter[seq] = 0u;
  const long long k0 = c * K, k1 = (k0 + K < T) ? k0 + K : T;
  SElem<N> acc;
  acc.set_identity();
#pragma unroll 1
  for (long long k = k0; k < k1; ++k) {
    SElem<N> e;
    load_selem_dense<N>(g, E, D, seq * T + k, e);
    if (k == k0) acc = e; else acc = smoothing_combine<N>(e, acc);
  }
  SElem<N> incl = warp_scan_inclusive<SElem<N>, true>(acc, lane);
  SElem<N> excl = warp_exclusive_from_inclusive<SElem<N>, true>(incl, lane);
  soa_store(chunk_suf, seq, Ppad, c, excl);
  if (lane == 0) soa_store(warp_stot, seq, Ppad / 32, c / 32, incl);
}

// Suffix scan outputs (g, D).  Without a carry the last element seeds the state exactly
// (g_T, D_T), as the reference's inclusive reverse scan does.
template <int N>
__global__ void __launch_bounds__(kBlock)
k_escan_smooth_apply(const double* g, const double* E, const double* D, long long T, int K, long long Ppad,
                     const double* carry_m, const double* carry_L, const double* chunk_suf, const double* warp_suf,
                     const double* group_suf, double* om, double* oL) {
  const long long seq = blockIdx.y;
  const long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const long long k0 = c * K, k1 = (k0 + K < T) ? k0 + K : T;
  if (k0 >= k1) return;
  Gauss<N> x;
  if (carry_m) {
    load_gauss_dense<N>(carry_m + seq * N, carry_L + seq * N * N, x);
  } else {
#pragma unroll
    for (int r = 0; r < N; ++r) {
      x.m[r] = 0.0;
#pragma unroll
      for (int q = 0; q <= r; ++q) x.Lc(r, q) = 0.0;
    }
  }
  SElem<N> e;
  soa_load(group_suf, seq, (Ppad / 32 + MidCfg<N>::IT - 1) / MidCfg<N>::IT, (Ppad / 32 - 1 - c / 32) / MidCfg<N>::IT, e);
  smoothing_apply<N>(x, e);
  soa_load(warp_suf, seq, Ppad / 32, c / 32, e);
  smoothing_apply<N>(x, e);
  soa_load(chunk_suf, seq, Ppad, c, e);
  smoothing_apply<N>(x, e);
#pragma unroll 1
  for (long long k = k1 - 1; k >= k0; --k) {
    load_selem_dense<N>(g, E, D, seq * T + k, e);
    smoothing_apply<N>(x, e);
    store_gauss_dense<N>(om + (seq * T + k) * N, oL + (seq * T + k) * N * N, x);
  }
}

// One-off combines of explicit element pairs (unit-test seam for _operators.py).
template <int N>
__global__ void k_filter_combine_pairs(const double* A1, const double* b1, const double* U1, const double* e1,
                                       const double* Z1, const double* A2, const double* b2, const double* U2,
                                       const double* e2, const double* Z2, long long n, double* A, double* b, double* U,
                                       double* eta, double* Z) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FElem<N> x, y;
  load_felem_dense<N>(A1, b1, U1, e1, Z1, i, x);
  load_felem_dense<N>(A2, b2, U2, e2, Z2, i, y);
  FElem<N> o = filtering_combine<N>(x, y);
#pragma unroll
  for (int r = 0; r < N; ++r) {
    b[i * N + r] = o.b(r);
    eta[i * N + r] = o.eta(r);
#pragma unroll
    for (int q = 0; q < N; ++q) {
      A[i * N * N + r * N + q] = o.A(r, q);
      U[i * N * N + r * N + q] = (q <= r) ? o.U(r, q) : 0.0;
      Z[i * N * N + r * N + q] = (q <= r) ? o.Z(r, q) : 0.0;
    }
  }
}

template <int N>
__global__ void k_smooth_combine_pairs(const double* g1, const double* E1, const double* D1, const double* g2,
                                       const double* E2, const double* D2, long long n, double* g, double* E,
                                       double* D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SElem<N> x, y;
  load_selem_dense<N>(g1, E1, D1, i, x);
  load_selem_dense<N>(g2, E2, D2, i, y);
  SElem<N> o = smoothing_combine<N>(x, y);
#pragma unroll
  for (int r = 0; r < N; ++r) {
    g[i * N + r] = o.g(r);
#pragma unroll
    for (int q = 0; q < N; ++q) {
      E[i * N * N + r * N + q] = o.E(r, q);
      D[i * N * N + r * N + q] = (q <= r) ? o.D(r, q) : 0.0;
    }
  }
}

// Fixed-order sum of per-warp log-likelihood partials (filter-only passes; one CTA / sequence).
template <int UNUSED>
__global__ void __launch_bounds__(256)
k_ell_sum(const double* __restrict__ ell_part, long long M, double* __restrict__ ell_out) {
  __shared__ double sred[8];
  const long long seq = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0;
  for (long long i = tid; i < M; i += 256) s += ell_part[seq * M + i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(kFull, s, d);
  if (lane == 0) sred[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double t = (lane < 8) ? sred[lane] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(kFull, t, d);
    if (lane == 0) ell_out[seq] = t;
  }
}

// tria of a [R x C] matrix per thread, streaming over the columns in blocks of 4:
// L <- tria([L | next 4 columns]) keeps only R(R+1)/2 + 4R doubles live.   _utils.py:22-24
template <int R>
__global__ void k_tria_batched(const double* __restrict__ A, double* __restrict__ L, int C, long long batch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const double* a = A + i * R * C;
  double Lt[R][R];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int q = 0; q < R; ++q) Lt[r][q] = 0.0;
#pragma unroll 1
  for (int c0 = 0; c0 < C; c0 += 4) {
    double W[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) W[r][q] = (c0 + q < C) ? a[r * C + c0 + q] : 0.0;
    tria_append<R, 4>([&](int r, int q) -> double& { return Lt[r][q]; }, W);
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int q = 0; q < R; ++q) L[i * R * R + r * R + q] = (q <= r) ? Lt[r][q] : 0.0;
}

template <int N>
__global__ void k_chol_update_batched(double* __restrict__ L, const double* __restrict__ V, int k, double alpha,
                                      long long batch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch) return;
  double Lt[N][N], w[N];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int q = 0; q < N; ++q) Lt[r][q] = L[i * N * N + r * N + q];
#pragma unroll 1
  for (int v = 0; v < k; ++v) {
#pragma unroll
    for (int r = 0; r < N; ++r) w[r] = V[(i * k + v) * N + r];
    chol_update<N>(Lt, w, alpha);
  }
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int q = 0; q < N; ++q) L[i * N * N + r * N + q] = Lt[r][q];
}

}  // namespace psq
