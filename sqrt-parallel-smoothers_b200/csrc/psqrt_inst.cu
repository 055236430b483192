// psqrt_inst.cu -- instantiates every kernel for ONE state dimension (compile with
// -DPSQ_N=<nx>); observation dimensions 1..PSQ_MAX_NY are instantiated for that nx.
#include "psqrt_launch.h"

#include <stdint.h>

#ifndef PSQ_N
#error "compile with -DPSQ_N=<nx>"
#endif
#ifndef PSQ_MAX_NY
#define PSQ_MAX_NY 4
#endif

#define PSQ_CAT_(a, b) a##b
#define PSQ_CAT(a, b) PSQ_CAT_(a, b)

#ifdef PSQ_STUB
// development builds (PSQRT_NX_LIST): this state dimension is not compiled in
namespace psq {
const LaunchN* PSQ_CAT(launch_n, PSQ_N)() { return nullptr; }
}
#else
#include "psqrt_kernels.cuh"
#if PSQ_N == 5
#include "psqrt_fused.cuh"
#endif
#if !PSQ_MID2
#include "psqrt_coop.cuh"
#endif
// sub-warp sweeps (psqrt_coopsweep.cuh): compiled for the state dimensions whose per-thread sweeps spill
#ifndef PSQ_COOP_SWEEPS
#define PSQ_COOP_SWEEPS (PSQ_N == 6 || PSQ_N == 8)
#endif
#if PSQ_COOP_SWEEPS
#include "psqrt_coopsweep.cuh"
#endif
// two rows per lane (psqrt_coopsweep2.cuh): nx = 8 on groups of 4 lanes
#ifndef PSQ_COOP_ROWS2
#define PSQ_COOP_ROWS2 (PSQ_N == 8)
#endif
#if PSQ_COOP_SWEEPS && PSQ_COOP_ROWS2
#include "psqrt_coopsweep2.cuh"
#endif

#include <stdlib.h>
#include <string.h>

namespace psq {
namespace {

constexpr int N = PSQ_N;

// The by-value (constant-bank) model variants double the number of sweep instantiations; they are
// compiled for the small state dimensions where they pay (nx <= 5) to keep build times in check.
constexpr bool kByValue = (PSQ_N <= 5);

// Fill a by-value model (kernel parameter) from the host mirrors.
template <int NY>
SrcVal<N, NY> make_src_val(const HostModel& h, const SSMArgs& a) {
  SrcVal<N, NY> s;
  for (int i = 0; i < N * N; ++i) { s.m.F[i] = h.F[i]; s.m.Q[i] = h.Q[i]; }
  for (int i = 0; i < N; ++i) s.m.bq[i] = h.bq[i];
  for (int i = 0; i < NY * N; ++i) s.m.H[i] = h.H ? h.H[i] : 0.0;
  for (int i = 0; i < NY * NY; ++i) s.m.R[i] = h.R ? h.R[i] : 0.0;
  for (int i = 0; i < NY; ++i) s.m.c[i] = h.c ? h.c[i] : 0.0;
  s.y = a.y; s.ty = a.ty; s.sy = a.sy;
  return s;
}

template <int NN>
SrcValT<NN> make_src_valT(const HostModel& h) {
  SrcValT<NN> s;
  for (int i = 0; i < NN * NN; ++i) { s.m.F[i] = h.F[i]; s.m.Q[i] = h.Q[i]; }
  for (int i = 0; i < NN; ++i) s.m.bq[i] = h.bq[i];
  return s;
}

#if PSQ_N == 5
inline void fill_fused(const HostFused& h, FusedCTBParams& p) {
  for (int i = 0; i < 25; ++i) p.Q[i] = h.Q[i];
  for (int i = 0; i < 5; ++i) p.mq[i] = h.mq[i];
  for (int i = 0; i < 4; ++i) p.R[i] = h.R[i];
  for (int i = 0; i < 2; ++i) p.mr[i] = h.mr[i];
  p.dt = h.dt; p.s1x = h.s1x; p.s1y = h.s1y; p.s2x = h.s2x; p.s2y = h.s2y;
}
inline SrcFusedCTB make_src_fused(const SSMArgs& a) {
  SrcFusedCTB s;
  fill_fused(*a.fused, s.pr);
  s.nom = a.fused->nom; s.nbs = a.fused->nbs;
  s.trig = a.fused->trig; s.tbs = a.fused->tbs;
  s.y = a.y; s.ty = a.ty; s.sy = a.sy;
  return s;
}
inline SrcFusedCT make_src_fusedT(const SSMArgs& a) {
  SrcFusedCT s;
  fill_fused(*a.fused, s.pr);
  s.nom = a.fused->nom; s.nbs = a.fused->nbs;
  s.trig = a.fused->trig; s.tbs = a.fused->tbs;
  return s;
}
void fused_prepare(const SSMArgs& a, long long T, long long B, cudaStream_t st) {
  const HostFused& h = *a.fused;
  k_fused_trig<<<dim3((unsigned)((T + 1 + 127) / 128), (unsigned)B, 1), 128, 0, st>>>(h.dt, h.s1x, h.s1y, h.s2x, h.s2y,
                                                                                    h.nom, h.nbs, T + 1, h.trig);
}
#endif

inline dim3 sweep_grid(long long Ppad, long long B) { return dim3((unsigned)(Ppad / kBlock), (unsigned)B, 1); }
inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// How a sweep writes its trajectory (psqrt_kernels.cuh, WarpOut): 16-byte words when the records are
// multiples of 16 bytes (even N) and the bases are 16-byte aligned, 8-byte words otherwise.
template <int NN>
bool vec2_ok(const void* m, const void* L) {
  return NN % 2 == 0 && aligned16(m) && aligned16(L);
}
// PSQRT_PDL=1 turns programmatic dependent launch on (psqrt_async.cuh, pdl_entry).  Default off: measured on B200 the
// early-scheduled successor CTAs slow the sweeps' tails more than the hidden launch latency gains (T = 1e6, nx = 4:
// 0.283 ms per pass with it, 0.276 ms without; profiles/r02_pdl_ab.txt).
inline bool pdl_on() {
  static const bool on = [] {
    const char* e = getenv("PSQRT_PDL");
    return e ? atoi(e) != 0 : false;
  }();
  return on;
}
// Launch of a kernel that calls pdl_entry(): optional cluster dimension, programmatic stream serialization.
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                       Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_on()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
template <class KERN, class... Args>
void launch_sweep_x(KERN kern, size_t smem, long long Ppad, long long B, int extra_ctas, cudaStream_t st, Args... args) {
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid = sweep_grid(Ppad, B);
  grid.x += (unsigned)extra_ctas;
  launch_pdl(kern, grid, dim3(kBlock, 1, 1), smem, st, 1, args...);
}
template <class KERN, class... Args>
void launch_sweep(KERN kern, size_t smem, long long Ppad, long long B, cudaStream_t st, Args... args) {
  launch_sweep_x(kern, smem, Ppad, B, 0, st, args...);
}
// Extra CTAs appended to K3 for the fused smoothing mid scan (psqrt_kernels.cuh, fused_smooth_mid): what is left of the
// sweeps' resident slots (2 CTAs per SM), at least 1 and at most 8.
inline int fuse_extra_ctas(long long work_ctas) {
  static const int slots = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * PSQ_MINB_K3;
  }();
  long long e = slots - work_ctas;
  if (e < 1) e = 1;
  if (e > 8) e = 8;
  return (int)e;
}

// Which sweeps run in sub-warp form: PSQRT_COOP = bit mask (1 = K1, 2 = K3, 4 = K5), default all where compiled.
inline int coop_mask() {
#if PSQ_COOP_SWEEPS
  static const int m = [] {
    const char* e = getenv("PSQRT_COOP");
    // measured on B200 (T = 1e6): nx = 8 5.9 -> 2.7 ms per pass in sub-warp form (two rows per lane), nx = 6
    // 1.49 -> 2.25 ms (its per-thread sweeps spill little, and two of its eight lanes idle): on by default at
    // nx = 8 only
    return e ? (atoi(e) & 7) : (PSQ_N == 8 ? 7 : 0);
  }();
  return m;
#else
  return 0;
#endif
}
#if PSQ_COOP_SWEEPS
// lanes per chunk: 4 with two matrix rows per lane (nx = 8), 8 with one row per lane
constexpr int kCoopLanes = PSQ_COOP_ROWS2 ? 4 : kCG;
inline dim3 coop_grid(long long Ppad, long long B) { return dim3((unsigned)(Ppad / kCChunks), (unsigned)B, 1); }
inline bool even(long long v) { return (v & 1) == 0; }
// 16-byte global accesses of the sub-warp sweeps: even N, every base 16-byte aligned, every stride even
inline int coop_vec(const SSMArgs& a, bool obs, const void* p0, const void* p1, const void* p2, const void* p3) {
  if (N % 2) return 0;
  bool ok = aligned16(a.F) && aligned16(a.Q) && even(a.tF) && even(a.tQ) && even(a.sF) && even(a.sQ);
  if (obs) ok = ok && aligned16(a.H) && even(a.tH) && even(a.sH);
  ok = ok && aligned16(p0) && aligned16(p1) && aligned16(p2) && aligned16(p3);
  return ok ? 1 : 0;
}
template <class OP, int NF, bool REV>
void unit_scan(double* items, long long Ppad, long long B, double* unit_tot, unsigned int* counter,
               unsigned int* fuse_ctr, cudaStream_t st) {
  constexpr size_t smem = unit_scan_smem_bytes<OP, NF>();
  static_assert(smem <= 227 * 1024, "unit scan: shared memory budget");
  auto kern = k_unit_scan<OP, NF, REV>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<coop_grid(Ppad, B), 32 * OP::G, smem, st>>>(items, Ppad, unit_tot, counter, fuse_ctr);
}
template <class KERN>
void coop_smem(KERN kern) {
  constexpr size_t smem = CoopSweep<N>::smem_bytes();
  static_assert(smem <= 113 * 1024, "two CTAs of a sub-warp sweep per SM");
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
#endif

template <int NY>
struct NYImpl {
  static void filter_reduce(const SSMArgs& a, const HostModel* hm, long long T, int K, long long Ppad, long long B,
                            double* chunk_own, double* chunk_pref, double* warp_tot, unsigned int* counter,
                            unsigned int* fuse_ctr, cudaStream_t st) {
    const size_t ysmem = LaneRing<NY, kYDepth>::smem_bytes(kBlock);
#if PSQ_COOP_SWEEPS
    if ((coop_mask() & 1) && !a.fused) {
      {
#if PSQ_COOP_ROWS2
        auto kern = k_coopr_filter_reduce<N, NY, kCoopLanes>;   // two rows per lane (psqrt_coopsweep2.cuh)
#else
        auto kern = k_coop_filter_reduce<N, NY>;
#endif
#if PSQ_COOP_ROWS2
        constexpr size_t k1smem = coopr_k1_smem_bytes<N, NY>();   // + the model bank
#else
        constexpr size_t k1smem = CoopSweep<N>::smem_bytes();
#endif
        static_assert(k1smem <= 113 * 1024, "two CTAs of K1 per SM");
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1smem);
        kern<<<coop_grid(Ppad, B), kCChunks * kCoopLanes, k1smem, st>>>(
            a, T, K, Ppad, chunk_own, chunk_pref, coop_vec(a, true, 0, 0, 0, 0));
      }
      unit_scan<CoopF2<N>, FElem<N>::NF, false>(chunk_pref, Ppad, B, warp_tot, counter, fuse_ctr, st);
      return;
    }
#endif
#if PSQ_N == 5
    if constexpr (NY == 2) {
      if (a.fused) {
        launch_sweep(k_filter_reduce<N, NY, SrcFusedCTB>, ysmem, Ppad, B, st, make_src_fused(a), T, K, Ppad, chunk_own,
                     chunk_pref, warp_tot, counter, fuse_ctr);
        return;
      }
    }
#endif
    if constexpr (kByValue) {
      if (hm) {
        launch_sweep(k_filter_reduce<N, NY, SrcVal<N, NY>>, ysmem, Ppad, B, st, make_src_val<NY>(*hm, a), T, K, Ppad,
                     chunk_own, chunk_pref, warp_tot, counter, fuse_ctr);
        return;
      }
    }
    launch_sweep(k_filter_reduce<N, NY, SrcPtr>, ysmem, Ppad, B, st, SrcPtr{a}, T, K, Ppad, chunk_own, chunk_pref,
                 warp_tot, counter, fuse_ctr);
  }
  template <bool SMOOTH, bool LOGLIK, class SRC>
  static void filter_apply_t(const SRC& src, long long T, int K, long long Ppad, long long B, const double* cm,
                             const double* cL, const double* chunk_own, const double* chunk_pref,
                             const double* warp_pref, const double* group_pref, double* fm, double* fL,
                             double* chunk_suf, double* warp_stot, double* ell_part, unsigned int* counter_s,
                             double* fpack, const FuseArgs* fuse, cudaStream_t st) {
    // fused smoothing mid scan: extra CTAs behind the workers (one sequence only: a batch orders its CTAs by sequence)
    const bool fz = SMOOTH && fuse && fuse->ctr && B == 1;
    const int n_work = fz ? (int)(Ppad / kBlock) : 0;
    const int extra = fz ? fuse_extra_ctas(n_work) : 0;
    double* const g_s = fz ? fuse->group_s : nullptr;
    double* const tot_s = fz ? fuse->stotal : nullptr;
    unsigned int* const fctr = fz ? fuse->ctr : nullptr;
#define PSQ_K3(OUT)                                                                                                  \
  launch_sweep_x(k_filter_apply<N, NY, SMOOTH, LOGLIK, SRC, OUT>,                                                    \
                 OUT::smem_bytes(kBlock) + LaneRing<NY, kYDepth>::smem_bytes(kBlock), Ppad, B, extra, st, src, T, K, \
                 Ppad, cm, cL, chunk_own, chunk_pref, warp_pref, group_pref, fm, fL, chunk_suf, warp_stot, ell_part, \
                 counter_s, fpack, n_work, g_s, tot_s, fctr)
    if constexpr (N % 2 == 0) {
      if (vec2_ok<N>(fm, fL)) { using O = WarpOut<N, 2>; PSQ_K3(O); return; }
    }
    { using O = WarpOut<N, 1>; PSQ_K3(O); }
#undef PSQ_K3
  }
  static void filter_apply(int smooth, const SSMArgs& a, const HostModel* hm, long long T, int K, long long Ppad,
                           long long B, const double* cm, const double* cL, double* chunk_own,
                           const double* chunk_pref, const double* warp_pref, const double* group_pref, double* fm,
                           double* fL, double* chunk_suf, double* warp_stot, double* ell_part,
                           unsigned int* counter_s, double* fpack, const FuseArgs* fuse, cudaStream_t st) {
#if PSQ_COOP_SWEEPS
    if ((coop_mask() & 2) && !a.fused) {
      // once per chunk: start states (parked in chunk_own), smoothing totals; then the sub-warp step loop
      if (smooth) {
        k_chunk_start<N, true><<<coop_grid(Ppad, B), 32, 0, st>>>(T, K, Ppad, cm, cL, chunk_own, chunk_pref, warp_pref,
                                                                  group_pref, fm, fL, chunk_suf, counter_s);
        unit_scan<CoopS2<N>, SElem<N>::NF, true>(chunk_suf, Ppad, B, warp_stot, nullptr, nullptr, st);
      } else {
        k_chunk_start<N, false><<<coop_grid(Ppad, B), 32, 0, st>>>(T, K, Ppad, cm, cL, chunk_own, chunk_pref, warp_pref,
                                                                   group_pref, fm, fL, chunk_suf, counter_s);
      }
      double* const fp = (smooth && !(coop_mask() & 4)) ? fpack : nullptr;   // only the per-thread K5 reads it
      const int vec = coop_vec(a, true, fm, fL, 0, 0);
      const long long css = (long long)FElem<N>::NF * Ppad;
      if (ell_part) {
#if PSQ_COOP_ROWS2
        auto kern = k_coopr_filter_apply<N, NY, kCoopLanes, true>;
#else
        auto kern = k_coop_filter_apply<N, NY, true>;
#endif
        coop_smem(kern);
        kern<<<coop_grid(Ppad, B), kCChunks * kCoopLanes, CoopSweep<N>::smem_bytes(), st>>>(a, T, K, Ppad, chunk_own,
                                                                                          css, fm, fL, ell_part, fp, vec);
      } else {
#if PSQ_COOP_ROWS2
        auto kern = k_coopr_filter_apply<N, NY, kCoopLanes, false>;
#else
        auto kern = k_coop_filter_apply<N, NY, false>;
#endif
        coop_smem(kern);
        kern<<<coop_grid(Ppad, B), kCChunks * kCoopLanes, CoopSweep<N>::smem_bytes(), st>>>(a, T, K, Ppad, chunk_own,
                                                                                          css, fm, fL, ell_part, fp, vec);
      }
      return;
    }
#endif
#define PSQ_FA(SM, SRCV)                                                                                             \
  do {                                                                                                               \
    if (ell_part)                                                                                                    \
      filter_apply_t<SM, true>(SRCV, T, K, Ppad, B, cm, cL, chunk_own, chunk_pref, warp_pref, group_pref, fm, fL,   \
                               chunk_suf, warp_stot, ell_part, counter_s, fpack, fuse, st);                          \
    else                                                                                                             \
      filter_apply_t<SM, false>(SRCV, T, K, Ppad, B, cm, cL, chunk_own, chunk_pref, warp_pref, group_pref, fm, fL,  \
                                chunk_suf, warp_stot, ell_part, counter_s, fpack, fuse, st);                         \
  } while (0)
#if PSQ_N == 5
    if constexpr (NY == 2) {
      if (a.fused) {
        const SrcFusedCTB sf = make_src_fused(a);
        if (smooth) PSQ_FA(true, sf); else PSQ_FA(false, sf);
        return;
      }
    }
#endif
    if constexpr (kByValue) {
      if (hm) {
        const SrcVal<N, NY> sv = make_src_val<NY>(*hm, a);
        if (smooth) PSQ_FA(true, sv); else PSQ_FA(false, sv);
        return;
      }
    }
    const SrcPtr sp{a};
    if (smooth) PSQ_FA(true, sp); else PSQ_FA(false, sp);
#undef PSQ_FA
  }
  static void filter_elements(const SSMArgs& a, long long T, long long B, const double* m0, const double* L0,
                              double* A, double* b, double* U, double* eta, double* Z, cudaStream_t st) {
    k_filter_elements<N, NY><<<blocks_for(T * B, 128), 128, 0, st>>>(a, T, B, m0, L0, A, b, U, eta, Z);
  }
  static void loglik_terms(const SSMArgs& a, long long T, long long B, const double* fm, const double* fL,
                           double* terms, cudaStream_t st) {
    k_loglik_terms<N, NY><<<blocks_for(T * B, 128), 128, 0, st>>>(a, T, B, fm, fL, terms);
  }
  static const LaunchNY* table() {
    static const LaunchNY t = {&filter_reduce, &filter_apply, &filter_elements, &loglik_terms};
    return &t;
  }
};

const LaunchNY* for_ny(int ny) {
  switch (ny) {
    case 1: return NYImpl<1>::table();
#if PSQ_MAX_NY >= 2
    case 2: return NYImpl<2>::table();
#endif
#if PSQ_MAX_NY >= 3
    case 3: return NYImpl<3>::table();
#endif
#if PSQ_MAX_NY >= 4
    case 4: return NYImpl<4>::table();
#endif
    default: return nullptr;
  }
}

#if PSQ_MID2
// Mid-level scans in the sub-warp form (psqrt_coop2.cuh): one lane group per combine, dense elements in shared memory.
template <class OP, int NF, bool REV>
void launch_mid2(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                 const double* ell_part, double* ell_out, const PushArgs* push, cudaStream_t st) {
  constexpr int IT = MidCfg<N>::IT;
  constexpr size_t smem = mid2_smem_bytes<OP, NF, IT>();
  static_assert(smem <= 227 * 1024, "mid scan: shared memory budget");
  auto kern = k_mid_scan2<OP, NF, IT, REV>;
  if (smem > 48 * 1024)  // per device, so not cached per process
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long Gc = (M + IT - 1) / IT;
  PushArgs pa;
  if (push) pa = *push; else memset(&pa, 0, sizeof(pa));
  launch_pdl(kern, dim3((unsigned)Gc, (unsigned)B, 1), dim3(IT * OP::G, 1, 1), smem, st, 1, items, M, groups, Gc,
             counter, total, ell_part, ell_out, pa);
}
// Cluster form (psqrt_coop2.cuh, k_mid_scan3): the IT = 64 items of a group spread over a cluster of CS CTAs on CS SMs.
// PSQRT_MID_CLUSTER: bit 0 = filtering scan, bit 1 = smoothing scan (default below; 0 = single-CTA groups everywhere).
constexpr bool kClusterMid = (PSQ_N <= 8);
constexpr int kClusterDefault = 1;
inline int cluster_mask() {
  static const int m = [] {
    const char* e = getenv("PSQRT_MID_CLUSTER");
    return e ? atoi(e) : kClusterDefault;
  }();
  return m;
}
template <class OP, int NF, bool REV>
bool launch_mid3(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                 const double* ell_part, double* ell_out, const PushArgs* push, cudaStream_t st) {
  if constexpr (kClusterMid) {
    constexpr int IT = MidCfg<N>::IT, CS = 4, IC = IT / CS;
    const long long Gc = (M + IT - 1) / IT;
    if (Gc > IT) return false;                     // more than one wave at the top level: single-CTA kernel
    // one CTA per SM: the dynamic shared-memory request is padded beyond half an SM
    constexpr size_t need = mid3_smem_bytes<OP, NF, CS, IC>();
    constexpr size_t smem = need > 116 * 1024 ? need : 116 * 1024;
    auto kern = k_mid_scan3<OP, NF, CS, IC, REV>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    PushArgs pa;
    if (push) pa = *push; else memset(&pa, 0, sizeof(pa));
    return launch_pdl(kern, dim3((unsigned)(Gc * CS), (unsigned)B, 1), dim3(IC * OP::G, 1, 1), smem, st, CS, items, M,
                      groups, Gc, counter, total, ell_part, ell_out, pa) == cudaSuccess;
  }
  return false;
}
void mid_filter(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                const PushArgs* push, cudaStream_t st) {
  if ((cluster_mask() & 1) &&
      launch_mid3<CoopF2<N>, FElem<N>::NF, false>(items, M, B, groups, counter, total, nullptr, nullptr, push, st))
    return;
  launch_mid2<CoopF2<N>, FElem<N>::NF, false>(items, M, B, groups, counter, total, nullptr, nullptr, push, st);
}
void mid_smooth(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                const double* ell_part, double* ell_out, const PushArgs* push, cudaStream_t st) {
  if ((cluster_mask() & 2) &&
      launch_mid3<CoopS2<N>, SElem<N>::NF, true>(items, M, B, groups, counter, total, ell_part, ell_out, push, st))
    return;
  launch_mid2<CoopS2<N>, SElem<N>::NF, true>(items, M, B, groups, counter, total, ell_part, ell_out, push, st);
}
// Time-shard carries: log-depth scan of the shard totals after a seed element built from the prior / terminal state.
template <class OP, int NF>
void launch_carry(const double* totals, int first, int step, int count, long long B, const double* m, const double* L,
                  double* cm, double* cL, const PeerCtx* pc, int seed_rank, long long mext, long long Lext,
                  const PushArgs* push, const double* own_total, cudaStream_t st) {
  constexpr size_t smem = carry_smem_bytes<OP, NF>();
  static_assert(smem <= 227 * 1024, "carry scan: shared memory budget");
  auto kern = k_carry_scan<OP, NF, N>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  PeerCtx p;
  if (pc) p = *pc; else memset(&p, 0, sizeof(p));
  const long long payload = pc ? pc->payload : NF;
  PushArgs pa;
  if (push) pa = *push; else memset(&pa, 0, sizeof(pa));
  kern<<<(unsigned)B, kCarryIT * OP::G, smem, st>>>(totals, first, step, count, B, payload, m, L, (long long)N,
                                                    (long long)N * N, cm, cL, p, pc ? 1 : 0, seed_rank, mext, Lext, pa,
                                                    own_total);
}
void carry_filter(const double* totals, int rank, long long B, const double* m0, const double* L0, double* cm,
                  double* cL, const PeerCtx* pc, cudaStream_t st) {
  launch_carry<CoopF2<N>, FElem<N>::NF>(totals, 0, 1, rank, B, m0, L0, cm, cL, pc, -1, 0, 0, nullptr, nullptr, st);
}
void carry_smoother(const double* totals, int rank, int R, long long B, const double* mT, const double* LT,
                    double* cm, double* cL, const PeerCtx* pc, cudaStream_t st) {
  // With a peer exchange `totals` [B][NF] is this rank's OWN smoothing total and (mT, LT) its own last filtered
  // state: the kernel publishes them to every rank first, then waits for all ranks; the terminal state is the last
  // rank's published last filtered state.
  if (pc) {
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.pc = *pc; pa.on = 1;
    pa.x1 = mT; pa.s1 = N; pa.n1 = N;
    pa.x2 = LT; pa.s2 = (long long)N * N; pa.n2 = N * N;
    launch_carry<CoopS2<N>, SElem<N>::NF>(nullptr, R - 1, -1, R - 1 - rank, B, mT, LT, cm, cL, pc, R - 1, SElem<N>::NF,
                                          SElem<N>::NF + N, &pa, totals, st);
    return;
  }
  launch_carry<CoopS2<N>, SElem<N>::NF>(totals, R - 1, -1, R - 1 - rank, B, mT, LT, cm, cL, nullptr, -1, SElem<N>::NF,
                                        SElem<N>::NF + N, nullptr, nullptr, st);
}
#else
inline dim3 mid_grid(long long M, long long B) { return dim3((unsigned)((M + 31) / 32), (unsigned)B, 1); }

// round-1 mid scans: the filtering one runs one half-warp per combine (psqrt_coop.cuh) for nx >= 5, where the
// one-thread-per-combine kernel spills
constexpr bool kCoopMid = (PSQ_N >= 5);
void mid_filter(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                const PushArgs*, cudaStream_t st) {
  if constexpr (kCoopMid) {
    constexpr size_t smem = coop_smem_bytes<N>();
    static_assert(smem <= 227 * 1024, "cooperative mid scan: shared memory budget");
    if (smem > 48 * 1024)  // per device, so not cached per process
      cudaFuncSetAttribute(k_mid_scan_coop<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_mid_scan_coop<N><<<mid_grid(M, B), 32 * kCoopWarps, smem, st>>>(items, M, groups, (M + 31) / 32, counter, total);
    return;
  }
  k_mid_scan<FElem<N>, false><<<mid_grid(M, B), 32 * Split<FElem<N>>::R, 0, st>>>(items, M, groups, (M + 31) / 32,
                                                                                 counter, total, nullptr, nullptr);
}
void mid_smooth(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                const double* ell_part, double* ell_out, const PushArgs*, cudaStream_t st) {
  k_mid_scan<SElem<N>, true><<<mid_grid(M, B), 32 * Split<SElem<N>>::R, 0, st>>>(items, M, groups, (M + 31) / 32, counter, total, ell_part,
                                                          ell_out);
}
#endif
void smooth_reduce(const SSMArgs& a, const HostModel* hm, long long T, int K, long long Ppad, long long B,
                   const double* fm, const double* fL, double* chunk_suf, double* warp_stot, unsigned int* counter,
                   double* fpack, cudaStream_t st) {
  (void)hm;  // the standalone smoother is not a hot path: pointer model only
  k_smooth_reduce<N, SrcPtr><<<sweep_grid(Ppad, B), kBlock, 0, st>>>(SrcPtr{a}, T, K, Ppad, fm, fL, chunk_suf,
                                                                      warp_stot, counter, fpack);
}
template <class SRC>
void smooth_apply_t(const SRC& src, long long T, int K, long long Ppad, long long B, const double* cm,
                    const double* cL, long long cms, long long cLs, const double* chunk_suf, const double* warp_suf,
                    const double* group_suf, const double* fpack, double* sm, double* sL, int write_terminal,
                    cudaStream_t st) {
#define PSQ_K5(OUT)                                                                                                  \
  launch_sweep(k_smooth_apply<N, SRC, OUT>, OUT::smem_bytes(kBlock) + LaneRing<N + Gauss<N>::TRI, kXDepth>::smem_bytes(kBlock), Ppad, B, st, src, T, K, Ppad, cm, cL, cms, cLs, \
               chunk_suf, warp_suf, group_suf, fpack, sm, sL, write_terminal)
  if constexpr (N % 2 == 0) {
    if (vec2_ok<N>(sm, sL)) { using O = WarpOut<N, 2>; PSQ_K5(O); return; }
  }
  { using O = WarpOut<N, 1>; PSQ_K5(O); }
#undef PSQ_K5
}
void smooth_apply(const SSMArgs& a, const HostModel* hm, long long T, int K, long long Ppad, long long B,
                  const double* cm, const double* cL, long long cms, long long cLs, double* chunk_suf,
                  const double* warp_suf, const double* group_suf, const double* fpack, const double* fm,
                  const double* fL, double* sm, double* sL, int write_terminal, cudaStream_t st) {
#if PSQ_COOP_SWEEPS
  if ((coop_mask() & 4) && !a.fused && fm && fL) {
    k_chunk_end<N><<<coop_grid(Ppad, B), 32, 0, st>>>(T, K, Ppad, cm, cL, cms, cLs, chunk_suf, warp_suf, group_suf, sm,
                                                      sL, write_terminal);
#if PSQ_COOP_ROWS2
    auto kern = k_coopr_smooth_apply<N, kCoopLanes>;
#else
    auto kern = k_coop_smooth_apply<N>;
#endif
    coop_smem(kern);
    kern<<<coop_grid(Ppad, B), kCChunks * kCoopLanes, CoopSweep<N>::smem_bytes(), st>>>(
        a, T, K, Ppad, chunk_suf, (long long)SElem<N>::NF * Ppad, fm, fL, sm, sL, coop_vec(a, false, fm, fL, sm, sL));
    return;
  }
#endif
#if PSQ_N == 5
  if (a.fused) {
    smooth_apply_t(make_src_fusedT(a), T, K, Ppad, B, cm, cL, cms, cLs, chunk_suf, warp_suf, group_suf, fpack, sm, sL,
                   write_terminal, st);
    return;
  }
#endif
  if constexpr (kByValue) {
    if (hm) {
      smooth_apply_t(make_src_valT<N>(*hm), T, K, Ppad, B, cm, cL, cms, cLs, chunk_suf, warp_suf, group_suf, fpack, sm,
                     sL, write_terminal, st);
      return;
    }
  }
  smooth_apply_t(SrcPtr{a}, T, K, Ppad, B, cm, cL, cms, cLs, chunk_suf, warp_suf, group_suf, fpack, sm, sL,
                 write_terminal, st);
}
#if !PSQ_MID2
void carry_filter(const double* totals, int rank, long long B, const double* m0, const double* L0, double* cm,
                  double* cL, const PeerCtx*, cudaStream_t st) {
  k_carry_filter<N><<<blocks_for(B, 32), 32, 0, st>>>(totals, rank, B, m0, L0, cm, cL);
}
void carry_smoother(const double* totals, int rank, int R, long long B, const double* mT, const double* LT,
                    double* cm, double* cL, const PeerCtx*, cudaStream_t st) {
  k_carry_smoother<N><<<blocks_for(B, 32), 32, 0, st>>>(totals, rank, R, B, mT, LT, cm, cL);
}
#endif
void smoother_elements(const SSMArgs& a, long long T, long long B, const double* fm, const double* fL, double* g,
                       double* E, double* D, cudaStream_t st) {
  k_smoother_elements<N><<<blocks_for((T + 1) * B, 128), 128, 0, st>>>(a, T, B, fm, fL, g, E, D);
}
void escan_filter_reduce(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                         long long T, int K, long long Ppad, long long B, double* chunk_pref, double* warp_tot,
                         unsigned int* counter, cudaStream_t st) {
  k_escan_filter_reduce<N><<<sweep_grid(Ppad, B), kBlock, 0, st>>>(A, b, U, eta, Z, T, K, Ppad, chunk_pref, warp_tot,
                                                                  counter);
}
void escan_filter_apply(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                        long long T, int K, long long Ppad, long long B, const double* chunk_pref,
                        const double* warp_pref, const double* group_pref, double* om, double* oL, cudaStream_t st) {
  k_escan_filter_apply<N><<<sweep_grid(Ppad, B), kBlock, 0, st>>>(A, b, U, eta, Z, T, K, Ppad, nullptr, nullptr,
                                                                 chunk_pref, warp_pref, group_pref, om, oL);
}
void escan_smooth_reduce(const double* g, const double* E, const double* D, long long T, int K, long long Ppad,
                         long long B, double* chunk_suf, double* warp_stot, unsigned int* counter, cudaStream_t st) {
  k_escan_smooth_reduce<N><<<sweep_grid(Ppad, B), kBlock, 0, st>>>(g, E, D, T, K, Ppad, chunk_suf, warp_stot, counter);
}
void escan_smooth_apply(const double* g, const double* E, const double* D, long long T, int K, long long Ppad,
                        long long B, const double* chunk_suf, const double* warp_suf, const double* group_suf,
                        double* om, double* oL, cudaStream_t st) {
  k_escan_smooth_apply<N><<<sweep_grid(Ppad, B), kBlock, 0, st>>>(g, E, D, T, K, Ppad, nullptr, nullptr, chunk_suf,
                                                                 warp_suf, group_suf, om, oL);
}
void filter_combine(const double* A1, const double* b1, const double* U1, const double* e1, const double* Z1,
                    const double* A2, const double* b2, const double* U2, const double* e2, const double* Z2,
                    long long n, double* A, double* b, double* U, double* eta, double* Z, cudaStream_t st) {
  k_filter_combine_pairs<N><<<blocks_for(n, 64), 64, 0, st>>>(A1, b1, U1, e1, Z1, A2, b2, U2, e2, Z2, n, A, b, U, eta,
                                                             Z);
}
void smooth_combine(const double* g1, const double* E1, const double* D1, const double* g2, const double* E2,
                    const double* D2, long long n, double* g, double* E, double* D, cudaStream_t st) {
  k_smooth_combine_pairs<N><<<blocks_for(n, 64), 64, 0, st>>>(g1, E1, D1, g2, E2, D2, n, g, E, D);
}
void tria(const double* A, double* L, int cols, long long batch, cudaStream_t st) {
  k_tria_batched<N><<<blocks_for(batch, 128), 128, 0, st>>>(A, L, cols, batch);
}
void chol_update(double* L, const double* V, int k, double alpha, long long batch, cudaStream_t st) {
  k_chol_update_batched<N><<<blocks_for(batch, 128), 128, 0, st>>>(L, V, k, alpha, batch);
}

const LaunchN kTable = {N,
                        FElem<N>::NF,
                        SElem<N>::NF,
                        N + Gauss<N>::TRI,
                        &for_ny,
                        &mid_filter,
                        &mid_smooth,
                        &smooth_reduce,
                        &smooth_apply,
                        &carry_filter,
                        &carry_smoother,
                        &smoother_elements,
                        &escan_filter_reduce,
                        &escan_filter_apply,
                        &escan_smooth_reduce,
                        &escan_smooth_apply,
                        &filter_combine,
                        &smooth_combine,
#if PSQ_N == 5
                        &fused_prepare,
#else
                        nullptr,
#endif
                        &tria,
                        &chol_update,
                        &coop_mask};

}  // namespace

const LaunchN* PSQ_CAT(launch_n, PSQ_N)() { return &kTable; }

#if defined(PSQ_MID_TRACE) && PSQ_N == 4
void mid_trace_read(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_mid_trace, sizeof(g_mid_trace));
}
#endif


}  // namespace psq
#endif  // PSQ_STUB
