// psqrt_launch.h -- type-erased launch tables: one LaunchN per compiled state dimension,
// one LaunchNY per compiled (nx, ny) pair.  psqrt_inst.cu fills them (one translation unit per
// nx so the unrolled templates compile in parallel); psqrt_capi.cu dispatches through them.
#pragma once
#include <cuda_runtime.h>

namespace psq {

struct SSMArgs;
struct PushArgs;
struct PeerCtx;

// host mirrors of a time- and batch-invariant model (all six non-null) or nullptr entries
struct HostModel {
  const double *F, *Q, *bq, *H, *R, *c;
};

// Built-in model linearised inside the sweeps (psqrt_fused.cuh; psqrt_ssm.fused_* in include/psqrt.h): coordinated
// turn + two bearings, extended linearization, nx = 5, ny = 2.  All host data except `nom` (device).
struct HostFused {
  double Q[25], mq[5], R[4], mr[2];
  double dt, s1x, s1y, s2x, s2y;
  const double* nom;   // [B][T + 1][5] nominal means (device)
  long long nbs;       // batch stride of nom, 0 = shared
  double* trig;        // [B][T + 1][4] workspace: transcendental values per nominal point (k_fused_trig)
  long long tbs;       // (T + 1) * 4
};

// Smoothing mid scan fused into K3 (psqrt_kernels.cuh, fused_smooth_mid): its group array, the sequence total (or
// null) and the two counters per sequence that K1 zeroes.  ctr == nullptr: K4 runs as its own kernel.
struct FuseArgs {
  double* group_s;
  double* stotal;
  unsigned int* ctr;
};

struct LaunchNY {
  void (*filter_reduce)(const SSMArgs&, const HostModel* hm, long long T, int K, long long Ppad, long long B,
                        double* chunk_own, double* chunk_pref, double* warp_tot, unsigned int* counter,
                        unsigned int* fuse_ctr, cudaStream_t);
  void (*filter_apply)(int smooth, const SSMArgs&, const HostModel* hm, long long T, int K, long long Ppad, long long B,
                       const double* carry_m, const double* carry_L, double* chunk_own,
                       const double* chunk_pref, const double* warp_pref, const double* group_pref, double* fm,
                       double* fL, double* chunk_suf, double* warp_stot, double* ell_part, unsigned int* counter_s,
                       double* fpack, const FuseArgs* fuse, cudaStream_t);
  void (*filter_elements)(const SSMArgs&, long long T, long long B, const double* m0, const double* L0, double* A,
                          double* b, double* U, double* eta, double* Z, cudaStream_t);
  void (*loglik_terms)(const SSMArgs&, long long T, long long B, const double* fm, const double* fL, double* terms,
                       cudaStream_t);
};

struct LaunchN {
  int n;
  int nf_filter, nf_smoother, nf_state;
  const LaunchNY* (*for_ny)(int ny);
  // push != nullptr: the CTA that finishes the scan publishes the total in every rank's exchange buffer (psqrt_coop2.cuh)
  void (*mid_filter)(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                     const PushArgs* push, cudaStream_t);
  void (*mid_smooth)(double* items, long long M, long long B, double* groups, unsigned int* counter, double* total,
                     const double* ell_part, double* ell_out, const PushArgs* push, cudaStream_t);
  void (*smooth_reduce)(const SSMArgs&, const HostModel* hm, long long T, int K, long long Ppad, long long B, const double* fm,
                        const double* fL, double* chunk_suf, double* warp_stot, unsigned int* counter, double* fpack,
                        cudaStream_t);
  // chunk_suf is an in/out scratch (the sub-warp form parks the per-chunk start states in it); fm, fL: the filtered
  // trajectory of the pass (read by the sub-warp form instead of fpack)
  void (*smooth_apply)(const SSMArgs&, const HostModel* hm, long long T, int K, long long Ppad, long long B,
                       const double* carry_m, const double* carry_L, long long cms, long long cLs,
                       double* chunk_suf, const double* warp_suf, const double* group_suf, const double* fpack,
                       const double* fm, const double* fL, double* sm, double* sL, int write_terminal, cudaStream_t);
  // pc != nullptr: wait for all ranks' pass number, then fold the totals out of the local exchange buffer
  void (*carry_filter)(const double* totals, int rank, long long B, const double* m0, const double* L0, double* cm,
                       double* cL, const PeerCtx* pc, cudaStream_t);
  void (*carry_smoother)(const double* totals, int rank, int R, long long B, const double* mT, const double* LT,
                         double* cm, double* cL, const PeerCtx* pc, cudaStream_t);
  void (*smoother_elements)(const SSMArgs&, long long T, long long B, const double* fm, const double* fL, double* g,
                            double* E, double* D, cudaStream_t);
  void (*escan_filter_reduce)(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                              long long T, int K, long long Ppad, long long B, double* chunk_pref, double* warp_tot,
                              unsigned int* counter, cudaStream_t);
  void (*escan_filter_apply)(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                             long long T, int K, long long Ppad, long long B, const double* chunk_pref,
                             const double* warp_pref, const double* group_pref, double* om, double* oL, cudaStream_t);
  void (*escan_smooth_reduce)(const double* g, const double* E, const double* D, long long T, int K, long long Ppad,
                              long long B, double* chunk_suf, double* warp_stot, unsigned int* counter, cudaStream_t);
  void (*escan_smooth_apply)(const double* g, const double* E, const double* D, long long T, int K, long long Ppad,
                             long long B, const double* chunk_suf, const double* warp_suf, const double* group_suf,
                             double* om, double* oL, cudaStream_t);
  void (*filter_combine)(const double* A1, const double* b1, const double* U1, const double* e1, const double* Z1,
                         const double* A2, const double* b2, const double* U2, const double* e2, const double* Z2,
                         long long n, double* A, double* b, double* U, double* eta, double* Z, cudaStream_t);
  void (*smooth_combine)(const double* g1, const double* E1, const double* D1, const double* g2, const double* E2,
                         const double* D2, long long n, double* g, double* E, double* D, cudaStream_t);
  // fused built-in linearization (nx = 5 only, else nullptr): fills HostFused::trig for the pass
  void (*fused_prepare)(const SSMArgs&, long long T, long long B, cudaStream_t);
  void (*tria)(const double* A, double* L, int cols, long long batch, cudaStream_t);
  void (*chol_update)(double* L, const double* V, int k, double alpha, long long batch, cudaStream_t);
  // bit mask of the sweeps that run in sub-warp form (psqrt_coopsweep.cuh): 1 = K1, 2 = K3, 4 = K5; 0 = per-thread
  int (*coop_mask)();
};

// defined by the per-nx translation units (psqrt_inst.cu compiled with -DPSQ_N=<nx>)
const LaunchN* launch_n1();
const LaunchN* launch_n2();
const LaunchN* launch_n3();
const LaunchN* launch_n4();
const LaunchN* launch_n5();
const LaunchN* launch_n6();
const LaunchN* launch_n8();
void ell_sum(const double* ell_part, long long M, long long B, double* ell_out, cudaStream_t);

}  // namespace psq
