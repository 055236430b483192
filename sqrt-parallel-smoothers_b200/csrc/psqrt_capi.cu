// psqrt_capi.cu -- extern "C" entry points declared in include/psqrt.h.
// Argument checking, chunk planning, workspace carving and dispatch to the per-(nx, ny)
// launch tables.  No allocation, no synchronisation, no exceptions.
#include "../../include/psqrt.h"

#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "psqrt_kernels.cuh"
#include "psqrt_launch.h"

namespace psq {
// FP64 FMA throughput probe (psqrt_fp64_probe): 8 independent dependent-DFMA chains per thread
__global__ void __launch_bounds__(128) k_fp64_probe(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256)
k_count_nonfinite(const double* __restrict__ x, long long row_len, unsigned long long* __restrict__ counts) {
  const long long row = blockIdx.y;
  unsigned long long n = 0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < row_len; i += (long long)gridDim.x * 256)
    n += isfinite(x[row * row_len + i]) ? 0ull : 1ull;
  n = __reduce_add_sync(0xffffffffu, (unsigned)n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counts + row, n);
}

void ell_sum(const double* ell_part, long long M, long long B, double* ell_out, cudaStream_t st) {
  k_ell_sum<0><<<(unsigned)B, 256, 0, st>>>(ell_part, M, ell_out);
}
}  // namespace psq

namespace {

using psq::HostModel;
using psq::LaunchN;
using psq::LaunchNY;
using psq::SSMArgs;

const LaunchN* table_for(int nx) {
  switch (nx) {
    case 1: return psq::launch_n1();
    case 2: return psq::launch_n2();
    case 3: return psq::launch_n3();
    case 4: return psq::launch_n4();
    case 5: return psq::launch_n5();
    case 6: return psq::launch_n6();
    case 8: return psq::launch_n8();
    default: return nullptr;
  }
}

// Chunking: aim at a number of chunks over the whole batch that depends on the form of the sweeps (make_plan below;
// kTargetThreads = 148 SMs x 256 threads is the register-limited occupancy of the per-thread sweeps), never fewer than
// kMinChunk steps per thread so the per-chunk prologue (two applies + one warp scan) stays amortised.
constexpr long long kTargetThreads = 148LL * 256;
constexpr int kMinChunk = 4;

// Which chunk-count target a pass uses must be the same in every stage of the pass (they share the scratch layouts),
// so it may depend only on what every stage sees: the state dimension and the TRANSITION part of the model.
// one_cta: per-thread sweeps with ONE 128-thread CTA per SM, three SMs left over (145 x 128 chunks), instead of two CTAs
// per SM (148 x 256).  The step loops are bound by FP64 issue and run no faster with two CTAs, which only double the
// chunk prologues and the mid-scan input; a grid that fills all 148 SMs exactly is slower again.  Measured on B200
// (T = 1e6, nx = 4, model by value: K = 27 / 53 / 54 -> 0.2669 / 0.2676 / 0.2646 ms per pass; nx = 5: 0.545 -> 0.471 ms;
// T = 1e5: nx = 4 0.115 -> 0.098, nx = 5 0.202 -> 0.165 ms).  A model read from HBM per step (nx <= 4, T = 1e6:
// 0.292 ms with two CTAs, 0.336 with one) needs the second CTA to hide its loads; nx = 6 keeps two as well.
bool one_cta_plan(const LaunchN* ln, const psqrt_ssm* s) {
  if (ln->coop_mask()) return false;
  if (ln->n == 5) return true;
  if (ln->n > 5 || !s) return false;
  return s->fused_model == PSQRT_FUSED_NONE && s->hF && s->hcholQ && s->hb && !s->F_ts && !s->cholQ_ts && !s->b_ts &&
         !s->F_bs && !s->cholQ_bs && !s->b_bs;
}

int make_plan(const LaunchN* ln, int64_t T, int64_t batch, int chunk_len, bool one_cta, psqrt_plan* p) {
  if (T <= 0 || batch <= 0 || chunk_len < 0) return PSQRT_EINVAL;
  long long K = chunk_len;
  if (K == 0) {
    // sub-warp sweeps (psqrt_coopsweep.cuh): 64 chunks per SM (two CTAs of 32 chunks).  PSQRT_TARGET_CHUNKS overrides
    // the number of chunks aimed at over the whole batch (tuning runs).
    static const long long env_target = [] { const char* e = getenv("PSQRT_TARGET_CHUNKS"); return e ? atoll(e) : 0LL; }();
    const long long target = env_target > 0 ? env_target
                             : ln->coop_mask() ? kTargetThreads / 4
                             : one_cta         ? 145LL * 128
                                               : kTargetThreads;
    long long per_seq = target / batch;
    if (per_seq < 32) per_seq = 32;
    K = (T + per_seq - 1) / per_seq;
    if (K < kMinChunk) K = kMinChunk;
  }
  if (K > T) K = T;
  if (K > 0x7fffffff) return PSQRT_EINVAL;
  long long P = (T + K - 1) / K;
  long long Ppad = (P + psq::kBlock - 1) / psq::kBlock * psq::kBlock;
  p->chunk_len = (int32_t)K;
  p->n_chunks = P;
  p->n_chunks_pad = Ppad;
  p->n_warps = Ppad / 32;
  p->nf_filter = ln->nf_filter;
  p->nf_smoother = ln->nf_smoother;
  return PSQRT_OK;
}

// Workspace carving (doubles).  One layout serves the fused pass, the staged calls and the
// element scans, so a workspace sized for the op can be reused across stages.
struct Ws {
  double *chunk_own, *chunk_pref, *warp_tot, *group_f, *ftotal, *chunk_suf, *warp_stot, *group_s, *stotal, *ell_part,
      *ell_tmp;
  double* fpack;  // [B][K][nf_state][Ppad] packed filtered states handed from the forward to the backward sweep
  double* trig;   // [B][T + 1][4] transcendental values of a fused built-in linearization (psqrt_fused.cuh)
  unsigned int *counter_f, *counter_s;
  unsigned int* counter_x;  // [B][2] publish / ticket counters of the smoothing mid scan fused into K3
  size_t doubles;
};
Ws carve(void* base, const psqrt_plan& p, int nf_state, int64_t B, int64_t T) {
  Ws w;
  double* d = (double*)base;
  size_t off = 0;
  auto take = [&](size_t n) {
    double* r = d ? d + off : nullptr;
    off += (n + 1) & ~(size_t)1;  // keep 16-byte alignment
    return r;
  };
  w.chunk_own = take((size_t)B * p.nf_filter * p.n_chunks_pad);
  w.chunk_pref = take((size_t)B * p.nf_filter * p.n_chunks_pad);
  const size_t G = (size_t)(p.n_warps + 31) / 32;
  w.warp_tot = take((size_t)B * p.nf_filter * p.n_warps);
  w.group_f = take((size_t)B * p.nf_filter * G);
  w.ftotal = take((size_t)B * p.nf_filter);
  w.chunk_suf = take((size_t)B * p.nf_smoother * p.n_chunks_pad);
  w.warp_stot = take((size_t)B * p.nf_smoother * p.n_warps);
  w.group_s = take((size_t)B * p.nf_smoother * G);
  w.stotal = take((size_t)B * p.nf_smoother);
  w.ell_part = take((size_t)B * p.n_warps);
  w.ell_tmp = take((size_t)B);
  w.counter_f = (unsigned int*)take((size_t)B);   // one 8-byte slot per sequence, used as uint32
  w.counter_s = (unsigned int*)take((size_t)B);
  w.counter_x = (unsigned int*)take((size_t)B);   // two uint32 per sequence
  w.fpack = take((size_t)B * (size_t)p.chunk_len * (size_t)nf_state * (size_t)p.n_chunks_pad);
  w.trig = take(nf_state == 20 ? (size_t)B * (size_t)(T + 1) * 4 : 0);   // nx = 5 only (nf_state = 5 + 15)
  w.doubles = off;
  return w;
}

// The fused-linearization description lives until this thread's next make_args call: every launch function copies it
// into its kernel parameters before returning.
SSMArgs make_args(const psqrt_ssm* s, const double* y, int ny, int64_t T) {
  static thread_local psq::HostFused hf;
  SSMArgs a;
  a.fused = nullptr;
  if (s->fused_model == PSQRT_FUSED_CT_BEARINGS) {
    for (int i = 0; i < 25; ++i) hf.Q[i] = s->hcholQ[i];
    for (int i = 0; i < 5; ++i) hf.mq[i] = s->hb[i];
    for (int i = 0; i < 4; ++i) hf.R[i] = s->hcholR ? s->hcholR[i] : 0.0;
    for (int i = 0; i < 2; ++i) hf.mr[i] = s->hc ? s->hc[i] : 0.0;
    hf.dt = s->fused_params[0];
    hf.s1x = s->fused_params[1]; hf.s1y = s->fused_params[2];
    hf.s2x = s->fused_params[3]; hf.s2y = s->fused_params[4];
    hf.nom = s->nom_m; hf.nbs = s->nom_bs;
    a.fused = &hf;
  }
  a.F = s->F; a.Q = s->cholQ; a.bq = s->b; a.H = s->H; a.R = s->cholR; a.c = s->c; a.y = y;
  a.tF = s->F_ts; a.tQ = s->cholQ_ts; a.tb = s->b_ts; a.tH = s->H_ts; a.tR = s->cholR_ts; a.tc = s->c_ts;
  a.ty = ny;
  a.sF = s->F_bs; a.sQ = s->cholQ_bs; a.sb = s->b_bs; a.sH = s->H_bs; a.sR = s->cholR_bs; a.sc = s->c_bs;
  a.sy = (long long)T * ny;
  return a;
}

// Host mirrors usable?  Only for a model shared by every step and sequence.
const HostModel* host_model(const psqrt_ssm* s, bool need_obs, HostModel* out) {
  if (!s->hF || !s->hcholQ || !s->hb) return nullptr;
  if (s->F_ts || s->cholQ_ts || s->b_ts || s->F_bs || s->cholQ_bs || s->b_bs) return nullptr;
  if (need_obs) {
    if (!s->hH || !s->hcholR || !s->hc) return nullptr;
    if (s->H_ts || s->cholR_ts || s->c_ts || s->H_bs || s->cholR_bs || s->c_bs) return nullptr;
  }
  out->F = s->hF; out->Q = s->hcholQ; out->bq = s->hb;
  out->H = need_obs ? s->hH : nullptr; out->R = need_obs ? s->hcholR : nullptr; out->c = need_obs ? s->hc : nullptr;
  return out;
}

// workspace part of a fused built-in linearization (the HostFused behind a.fused is this thread's, see make_args)
void bind_fused(const SSMArgs& a, const Ws& w, int64_t T) {
  if (!a.fused) return;
  psq::HostFused* h = const_cast<psq::HostFused*>(a.fused);
  h->trig = w.trig;
  h->tbs = (long long)(T + 1) * 4;
}

// Dimensions the tuned kernels cover; everything else up to 16 goes to the generic path (psqrt_generic.cu).
// PSQRT_FORCE_GENERIC=1 sends the whole-pass entry points there for every dimension (tests, A/B).
bool force_generic() {
  static const bool f = [] { const char* e = getenv("PSQRT_FORCE_GENERIC"); return e && atoi(e) != 0; }();
  return f;
}
bool tuned_dims(int nx, int ny) {
  const LaunchN* ln = table_for(nx);
  if (!ln) return false;
  return ny == 0 || ln->for_ny(ny) != nullptr;
}

int check_launch() { return cudaGetLastError() == cudaSuccess ? PSQRT_OK : PSQRT_ECUDA; }

bool peer_ok(const psqrt_peer* p, int64_t batch) {
  return p->bufs && p->n_ranks > 0 && p->rank >= 0 && p->rank < p->n_ranks && p->batch == batch && p->slot > 0 &&
         p->payload > 0;
}
psq::PeerCtx peer_ctx(const psqrt_peer* p) {
  psq::PeerCtx c;
  c.bufs = reinterpret_cast<double* const*>(p->bufs);
  c.rank = p->rank; c.n_ranks = p->n_ranks; c.batch = p->batch;
  c.flags_off = p->flags_off; c.ctr_off = p->ctr_off; c.data_off = p->data_off; c.slot = p->slot; c.payload = p->payload;
  return c;
}

// fused_dims: (nx, ny) of an entry point that accepts a fused built-in linearization ({0, 0}: it does not)
bool ssm_ok(const psqrt_ssm* s, bool need_obs, int fused_nx = 0, int fused_ny = 0) {
  if (!s) return false;
  if (s->fused_model != PSQRT_FUSED_NONE) {
    if (s->fused_model != PSQRT_FUSED_CT_BEARINGS || fused_nx != 5 || (need_obs && fused_ny != 2)) return false;
    if (!s->nom_m || !s->fused_params || !s->hcholQ || !s->hb) return false;
    if (need_obs && (!s->hcholR || !s->hc)) return false;
    return true;
  }
  if (!s->F || !s->cholQ || !s->b) return false;
  if (need_obs && (!s->H || !s->cholR || !s->c)) return false;
  return true;
}

struct Ctx {
  const LaunchN* ln;
  const LaunchNY* lny;
  psqrt_plan plan;
  Ws ws;
};

int setup(Ctx& c, int nx, int ny, int64_t T, int64_t B, int chunk_len, void* ws, size_t ws_bytes,
          const psqrt_ssm* ssm) {
  c.ln = table_for(nx);
  if (!c.ln) return PSQRT_EUNSUPPORTED;
  c.lny = nullptr;
  if (ny > 0) {
    c.lny = c.ln->for_ny(ny);
    if (!c.lny) return PSQRT_EUNSUPPORTED;
  }
  if (B > 65535) return PSQRT_EINVAL;
  int rc = make_plan(c.ln, T, B, chunk_len, one_cta_plan(c.ln, ssm), &c.plan);
  if (rc) return rc;
  c.ws = carve(ws, c.plan, c.ln->nf_state, B, T);
  if (!ws || ws_bytes < c.ws.doubles * sizeof(double)) return PSQRT_EWORKSPACE;
  return PSQRT_OK;
}

}  // namespace

extern "C" {

int psqrt_version(void) { return PSQRT_VERSION; }

const char* psqrt_error_string(int code) {
  switch (code) {
    case PSQRT_OK: return "ok";
    case PSQRT_EINVAL: return "invalid argument";
    case PSQRT_EUNSUPPORTED: return "state/observation dimension not compiled in";
    case PSQRT_EWORKSPACE: return "workspace missing or too small";
    case PSQRT_ECUDA: return "CUDA launch failed";
    default: return "unknown error";
  }
}

int psqrt_supported(int nx, int ny) {
  const LaunchN* ln = table_for(nx);
  if (!ln) return 0;
  if (ny == 0) return 1;
  return ln->for_ny(ny) != nullptr;
}

int psqrt_get_plan(int nx, int ny, int64_t T, int64_t batch, int chunk_len, psqrt_plan* out) {
  (void)ny;
  const LaunchN* ln = table_for(nx);
  if (!ln) return PSQRT_EUNSUPPORTED;
  if (!out) return PSQRT_EINVAL;
  return make_plan(ln, T, batch, chunk_len, one_cta_plan(ln, nullptr), out);
}

int psqrt_get_plan_ssm(const psqrt_ssm* ssm, int nx, int ny, int64_t T, int64_t batch, int chunk_len, psqrt_plan* out) {
  const LaunchN* ln = table_for(nx);
  if (!ln || (ny > 0 && !ln->for_ny(ny))) return PSQRT_EUNSUPPORTED;
  if (!out) return PSQRT_EINVAL;
  return make_plan(ln, T, batch, chunk_len, one_cta_plan(ln, ssm), out);
}

size_t psqrt_workspace_bytes(int op, int nx, int ny, int64_t T, int64_t batch, int chunk_len) {
  const size_t gen = (op == PSQRT_OP_FILTER_SMOOTHER) ? psqrt_generic_workspace_bytes(nx, T, batch, 0) : 0;
  if (!tuned_dims(nx, ny)) return gen;   // served by the generic path (0 if it cannot either)
  const LaunchN* ln = table_for(nx);
  // the plan may depend on the form of the model (one_cta_plan): a workspace of this size serves either
  size_t tuned = 0;
  for (int v = 0; v < 2; ++v) {
    psqrt_plan p;
    if (make_plan(ln, T, batch, chunk_len, v != 0, &p)) return 0;
    const size_t b = carve(nullptr, p, ln->nf_state, batch, T).doubles * sizeof(double);
    tuned = b > tuned ? b : tuned;
  }
  return (force_generic() && gen > tuned) ? gen : tuned;
}

int psqrt_filter_reduce(const psqrt_ssm* ssm, const double* y, int nx, int ny, int64_t T, int64_t batch,
                        int chunk_len, double* ftotal, void* ws, size_t ws_bytes, const psqrt_peer* peer,
                        void* stream) {
  if (!ssm_ok(ssm, true, nx, ny) || !y || ny <= 0) return PSQRT_EINVAL;
  if (peer && !peer_ok(peer, batch)) return PSQRT_EINVAL;
  Ctx c;
  int rc = setup(c, nx, ny, T, batch, chunk_len, ws, ws_bytes, ssm);
  if (rc) return rc;
  if (peer && peer->payload != c.ln->nf_filter) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  SSMArgs a = make_args(ssm, y, ny, T);
  bind_fused(a, c.ws, T);
  if (a.fused) {   // the transcendental values of every nominal point, once per pass (the later stages reuse them)
    if (!c.ln->fused_prepare) return PSQRT_EUNSUPPORTED;
    c.ln->fused_prepare(a, T, batch, st);
  }
  HostModel hmv;
  c.lny->filter_reduce(a, host_model(ssm, true, &hmv), T, c.plan.chunk_len, c.plan.n_chunks_pad, batch, c.ws.chunk_own,
                       c.ws.chunk_pref, c.ws.warp_tot, c.ws.counter_f, c.ws.counter_x, st);
  psq::PushArgs pa;
  memset(&pa, 0, sizeof(pa));
  if (peer) { pa.pc = peer_ctx(peer); pa.on = 1; }
  c.ln->mid_filter(c.ws.warp_tot, c.plan.n_warps, batch, c.ws.group_f, c.ws.counter_f, ftotal ? ftotal : c.ws.ftotal,
                   peer ? &pa : nullptr, st);
  return check_launch();
}

int psqrt_carry_filter(const double* totals, int rank, int64_t batch, int nx, const double* m0, const double* L0,
                       double* carry_m, double* carry_L, const psqrt_peer* peer, void* stream) {
  const LaunchN* ln = table_for(nx);
  if (!ln) return PSQRT_EUNSUPPORTED;
  if (rank < 0 || batch <= 0 || batch > 65535 || !m0 || !L0 || !carry_m || !carry_L) return PSQRT_EINVAL;
  psq::PeerCtx pc;
  if (peer) {
    if (!peer_ok(peer, batch) || peer->rank != rank || peer->payload != ln->nf_filter) return PSQRT_EINVAL;
    pc = peer_ctx(peer);
  } else if (rank > 0 && !totals) {
    return PSQRT_EINVAL;
  }
  ln->carry_filter(totals, rank, batch, m0, L0, carry_m, carry_L, peer ? &pc : nullptr, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_filter_apply(const psqrt_ssm* ssm, const double* y, const double* carry_m, const double* carry_L, int nx,
                       int ny, int64_t T, int64_t batch, int chunk_len, double* fm, double* fL, double* ell,
                       double* stotal, void* ws, size_t ws_bytes, const psqrt_peer* peer, void* stream) {
  if (!ssm_ok(ssm, true, nx, ny) || !y || ny <= 0 || !carry_m || !carry_L || !fm || !fL) return PSQRT_EINVAL;
  if (peer && (!peer_ok(peer, batch) || !stotal)) return PSQRT_EINVAL;
  Ctx c;
  int rc = setup(c, nx, ny, T, batch, chunk_len, ws, ws_bytes, ssm);
  if (rc) return rc;
  if (peer && peer->payload != c.ln->nf_smoother + nx + (int64_t)nx * nx) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  SSMArgs a = make_args(ssm, y, ny, T);
  bind_fused(a, c.ws, T);
  const int smooth = stotal != nullptr;
  HostModel hmv;
  // Without a peer exchange and for a single sequence the smoothing mid scan (K4) runs inside K3 on a few extra
  // CTAs, concurrently with the workers' step loops (psqrt_kernels.cuh, fused_smooth_mid); PSQRT_FUSE_MID=0: own kernel.
  static const bool fuse_env = [] { const char* e = getenv("PSQRT_FUSE_MID"); return e ? atoi(e) != 0 : true; }();
  const bool fuse = smooth && batch == 1 && fuse_env && !(c.ln->coop_mask() & 2);
  psq::FuseArgs fa;
  fa.group_s = c.ws.group_s; fa.stotal = stotal; fa.ctr = fuse ? c.ws.counter_x : nullptr;
  c.lny->filter_apply(smooth, a, host_model(ssm, true, &hmv), T, c.plan.chunk_len, c.plan.n_chunks_pad, batch, carry_m,
                      carry_L, c.ws.chunk_own, c.ws.chunk_pref, c.ws.warp_tot, c.ws.group_f, fm, fL, c.ws.chunk_suf,
                      c.ws.warp_stot, ell ? c.ws.ell_part : nullptr, c.ws.counter_s, c.ws.fpack, &fa, st);
  if (fuse) {
    if (ell) psq::ell_sum(c.ws.ell_part, c.plan.n_warps, batch, ell, st);
  } else if (smooth) {
    // time-sharded pass: the smoothing total is published -- together with the shard's last filtered state -- by
    // psqrt_carry_smoother, so that it may come from the scan hidden inside K3 as well as from this kernel
    c.ln->mid_smooth(c.ws.warp_stot, c.plan.n_warps, batch, c.ws.group_s, c.ws.counter_s, stotal,
                     ell ? c.ws.ell_part : nullptr, ell, nullptr, st);
  } else if (ell) {
    psq::ell_sum(c.ws.ell_part, c.plan.n_warps, batch, ell, st);
  }
  return check_launch();
}

int psqrt_carry_smoother(const double* totals, int rank, int n_ranks, int64_t batch, int nx, const double* mT,
                         const double* LT, double* carry_m, double* carry_L, const psqrt_peer* peer, void* stream) {
  const LaunchN* ln = table_for(nx);
  if (!ln) return PSQRT_EUNSUPPORTED;
  if (rank < 0 || rank >= n_ranks || batch <= 0 || batch > 65535 || !carry_m || !carry_L) return PSQRT_EINVAL;
  psq::PeerCtx pc;
  if (peer) {
    if (!peer_ok(peer, batch) || peer->rank != rank || peer->n_ranks != n_ranks ||
        peer->payload != ln->nf_smoother + nx + (int64_t)nx * nx)
      return PSQRT_EINVAL;
    if (!totals || !mT || !LT) return PSQRT_EINVAL;   // this rank's own total and last filtered state: published here
    pc = peer_ctx(peer);
  } else if (!mT || !LT || (rank + 1 < n_ranks && !totals)) {
    return PSQRT_EINVAL;
  }
  ln->carry_smoother(totals, rank, n_ranks, batch, mT, LT, carry_m, carry_L, peer ? &pc : nullptr,
                     (cudaStream_t)stream);
  return check_launch();
}

int64_t psqrt_peer_layout(int nx, int n_ranks, int64_t batch, psqrt_peer* f, psqrt_peer* s) {
  const LaunchN* ln = table_for(nx);
  if (!ln || n_ranks <= 0 || batch <= 0 || !f || !s) return 0;
  int64_t off = 0;
  auto phase = [&](psqrt_peer* p, int64_t payload) {
    p->n_ranks = n_ranks; p->batch = batch; p->payload = payload; p->slot = batch * payload;
    p->flags_off = off; off += (int64_t)n_ranks * batch;
    p->ctr_off = off; off += batch;
    off = (off + 1) & ~(int64_t)1;
    p->data_off = off; off += 2 * (int64_t)n_ranks * p->slot;
    off = (off + 1) & ~(int64_t)1;
  };
  phase(f, ln->nf_filter);
  phase(s, ln->nf_smoother + nx + (int64_t)nx * nx);
  return off;
}

int psqrt_smoother_apply(const psqrt_ssm* ssm, const double* fm, const double* fL, const double* carry_m,
                         const double* carry_L, int write_terminal, int nx, int64_t T, int64_t batch, int chunk_len,
                         double* sm, double* sL, void* ws, size_t ws_bytes, void* stream) {
  // fm, fL: the per-thread backward sweep reads the packed copy psqrt_filter_apply left in the workspace (psqrt.h);
  // the sub-warp form (nx = 6, 8) reads the trajectory itself
  if (!ssm_ok(ssm, false, nx, 0) || !carry_m || !carry_L || !sm || !sL) return PSQRT_EINVAL;
  Ctx c;
  int rc = setup(c, nx, 0, T, batch, chunk_len, ws, ws_bytes, ssm);
  if (rc) return rc;
  SSMArgs a = make_args(ssm, nullptr, 0, T);
  bind_fused(a, c.ws, T);
  HostModel hmv;
  c.ln->smooth_apply(a, host_model(ssm, false, &hmv), T, c.plan.chunk_len, c.plan.n_chunks_pad, batch, carry_m, carry_L,
                     nx, (long long)nx * nx, c.ws.chunk_suf, c.ws.warp_stot, c.ws.group_s, c.ws.fpack, fm, fL, sm, sL,
                     write_terminal, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_filter_smoother(const psqrt_ssm* ssm, const double* y, const double* m0, const double* L0, int nx, int ny,
                          int64_t T, int64_t batch, int chunk_len, double* fm, double* fL, double* sm, double* sL,
                          double* ell, void* ws, size_t ws_bytes, void* stream) {
  if (!m0 || !L0 || !fm || !fL || ((sm == nullptr) != (sL == nullptr))) return PSQRT_EINVAL;
  if ((!tuned_dims(nx, ny) || force_generic()) && ssm && ssm->fused_model == PSQRT_FUSED_NONE)
    return psqrt_filter_smoother_generic(ssm, y, m0, L0, nx, ny, T, batch, fm, fL, sm, sL, ell, ws, ws_bytes, stream);
  Ctx c;
  int rc = setup(c, nx, ny, T, batch, chunk_len, ws, ws_bytes, ssm);
  if (rc) return rc;
  rc = psqrt_filter_reduce(ssm, y, nx, ny, T, batch, chunk_len, c.ws.ftotal, ws, ws_bytes, nullptr, stream);
  if (rc) return rc;
  const bool smooth = sm != nullptr;
  rc = psqrt_filter_apply(ssm, y, m0, L0, nx, ny, T, batch, chunk_len, fm, fL, ell, smooth ? c.ws.stotal : nullptr, ws,
                          ws_bytes, nullptr, stream);
  if (rc || !smooth) return rc;
  // terminal carry = filtered state at index T of every sequence
  SSMArgs a = make_args(ssm, nullptr, 0, T);
  bind_fused(a, c.ws, T);
  HostModel hmv;
  c.ln->smooth_apply(a, host_model(ssm, false, &hmv), T, c.plan.chunk_len, c.plan.n_chunks_pad, batch,
                     fm + (size_t)T * nx, fL + (size_t)T * nx * nx, (long long)(T + 1) * nx,
                     (long long)(T + 1) * nx * nx, c.ws.chunk_suf, c.ws.warp_stot, c.ws.group_s, c.ws.fpack, fm, fL, sm,
                     sL, 1, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_smoother(const psqrt_ssm* ssm, const double* fm, const double* fL, int nx, int64_t T, int64_t batch,
                   int chunk_len, double* sm, double* sL, void* ws, size_t ws_bytes, void* stream) {
  if (!ssm_ok(ssm, false) || !fm || !fL || !sm || !sL) return PSQRT_EINVAL;
  if (!tuned_dims(nx, 0) || force_generic())
    return psqrt_filter_smoother_generic(ssm, nullptr, nullptr, nullptr, nx, 0, T, batch, const_cast<double*>(fm),
                                         const_cast<double*>(fL), sm, sL, nullptr, ws, ws_bytes, stream);
  Ctx c;
  int rc = setup(c, nx, 0, T, batch, chunk_len, ws, ws_bytes, ssm);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SSMArgs a = make_args(ssm, nullptr, 0, T);
  HostModel hmv;
  c.ln->smooth_reduce(a, host_model(ssm, false, &hmv), T, c.plan.chunk_len, c.plan.n_chunks_pad, batch, fm, fL,
                      c.ws.chunk_suf, c.ws.warp_stot, c.ws.counter_s, c.ws.fpack, st);
  c.ln->mid_smooth(c.ws.warp_stot, c.plan.n_warps, batch, c.ws.group_s, c.ws.counter_s, c.ws.stotal, nullptr, nullptr,
                   nullptr, st);
  c.ln->smooth_apply(a, host_model(ssm, false, &hmv), T, c.plan.chunk_len, c.plan.n_chunks_pad, batch,
                     fm + (size_t)T * nx, fL + (size_t)T * nx * nx, (long long)(T + 1) * nx,
                     (long long)(T + 1) * nx * nx, c.ws.chunk_suf, c.ws.warp_stot, c.ws.group_s, c.ws.fpack, fm, fL, sm,
                     sL, 1, st);
  return check_launch();
}

int psqrt_filter_elements(const psqrt_ssm* ssm, const double* y, const double* m0, const double* L0, int nx, int ny,
                          int64_t T, int64_t batch, double* A, double* b, double* U, double* eta, double* Z,
                          void* stream) {
  if (!ssm_ok(ssm, true) || !y || T <= 0 || batch <= 0 || !A || !b || !U || !eta || !Z) return PSQRT_EINVAL;
  if ((m0 == nullptr) != (L0 == nullptr)) return PSQRT_EINVAL;
  const LaunchN* ln = table_for(nx);
  const LaunchNY* lny = ln ? ln->for_ny(ny) : nullptr;
  if (!lny) return PSQRT_EUNSUPPORTED;
  lny->filter_elements(make_args(ssm, y, ny, T), T, batch, m0, L0, A, b, U, eta, Z, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_filter_scan(const double* A, const double* b, const double* U, const double* eta, const double* Z, int nx,
                      int64_t T, int64_t batch, int chunk_len, double* means, double* chols, void* ws,
                      size_t ws_bytes, void* stream) {
  if (!A || !b || !U || !eta || !Z || !means || !chols) return PSQRT_EINVAL;
  Ctx c;
  int rc = setup(c, nx, 0, T, batch, chunk_len, ws, ws_bytes, nullptr);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  c.ln->escan_filter_reduce(A, b, U, eta, Z, T, c.plan.chunk_len, c.plan.n_chunks_pad, batch, c.ws.chunk_pref,
                            c.ws.warp_tot, c.ws.counter_f, st);
  c.ln->mid_filter(c.ws.warp_tot, c.plan.n_warps, batch, c.ws.group_f, c.ws.counter_f, c.ws.ftotal, nullptr, st);
  c.ln->escan_filter_apply(A, b, U, eta, Z, T, c.plan.chunk_len, c.plan.n_chunks_pad, batch, c.ws.chunk_pref,
                           c.ws.warp_tot, c.ws.group_f, means, chols, st);
  return check_launch();
}

int psqrt_smoother_elements(const psqrt_ssm* ssm, const double* fm, const double* fL, int nx, int64_t T,
                            int64_t batch, double* g, double* E, double* D, void* stream) {
  if (!ssm_ok(ssm, false) || !fm || !fL || T <= 0 || batch <= 0 || !g || !E || !D) return PSQRT_EINVAL;
  const LaunchN* ln = table_for(nx);
  if (!ln) return PSQRT_EUNSUPPORTED;
  ln->smoother_elements(make_args(ssm, nullptr, 0, T), T, batch, fm, fL, g, E, D, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_smoother_scan(const double* g, const double* E, const double* D, int nx, int64_t n, int64_t batch,
                        int chunk_len, double* means, double* chols, void* ws, size_t ws_bytes, void* stream) {
  if (!g || !E || !D || !means || !chols) return PSQRT_EINVAL;
  Ctx c;
  int rc = setup(c, nx, 0, n, batch, chunk_len, ws, ws_bytes, nullptr);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  c.ln->escan_smooth_reduce(g, E, D, n, c.plan.chunk_len, c.plan.n_chunks_pad, batch, c.ws.chunk_suf, c.ws.warp_stot,
                            c.ws.counter_s, st);
  c.ln->mid_smooth(c.ws.warp_stot, c.plan.n_warps, batch, c.ws.group_s, c.ws.counter_s, c.ws.stotal, nullptr, nullptr,
                   nullptr, st);
  c.ln->escan_smooth_apply(g, E, D, n, c.plan.chunk_len, c.plan.n_chunks_pad, batch, c.ws.chunk_suf, c.ws.warp_stot,
                           c.ws.group_s, means, chols, st);
  return check_launch();
}

int psqrt_loglik_terms(const psqrt_ssm* ssm, const double* y, const double* fm, const double* fL, int nx, int ny,
                       int64_t T, int64_t batch, double* terms, void* stream) {
  if (!ssm_ok(ssm, true) || !y || !fm || !fL || !terms || T <= 0 || batch <= 0) return PSQRT_EINVAL;
  const LaunchN* ln = table_for(nx);
  const LaunchNY* lny = ln ? ln->for_ny(ny) : nullptr;
  if (!lny) return PSQRT_EUNSUPPORTED;
  lny->loglik_terms(make_args(ssm, y, ny, T), T, batch, fm, fL, terms, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_filter_combine(const double* A1, const double* b1, const double* U1, const double* eta1, const double* Z1,
                         const double* A2, const double* b2, const double* U2, const double* eta2, const double* Z2,
                         int nx, int64_t n, double* A, double* b, double* U, double* eta, double* Z, void* stream) {
  const LaunchN* ln = table_for(nx);
  if (!ln) return PSQRT_EUNSUPPORTED;
  if (n <= 0 || !A1 || !b1 || !U1 || !eta1 || !Z1 || !A2 || !b2 || !U2 || !eta2 || !Z2 || !A || !b || !U || !eta || !Z)
    return PSQRT_EINVAL;
  ln->filter_combine(A1, b1, U1, eta1, Z1, A2, b2, U2, eta2, Z2, n, A, b, U, eta, Z, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_smoother_combine(const double* g1, const double* E1, const double* D1, const double* g2, const double* E2,
                           const double* D2, int nx, int64_t n, double* g, double* E, double* D, void* stream) {
  const LaunchN* ln = table_for(nx);
  if (!ln) return PSQRT_EUNSUPPORTED;
  if (n <= 0 || !g1 || !E1 || !D1 || !g2 || !E2 || !D2 || !g || !E || !D) return PSQRT_EINVAL;
  ln->smooth_combine(g1, E1, D1, g2, E2, D2, n, g, E, D, (cudaStream_t)stream);
  return check_launch();
}

int psqrt_tria_batched(const double* A, double* L, int rows, int cols, int64_t batch, void* stream) {
  const LaunchN* ln = table_for(rows);
  if (!ln) return (rows >= 1 && rows <= 16) ? psqrt_tria_generic(A, L, rows, cols, batch, stream) : PSQRT_EUNSUPPORTED;
  if (!A || !L || cols <= 0 || batch <= 0) return PSQRT_EINVAL;
  ln->tria(A, L, cols, batch, (cudaStream_t)stream);
  return check_launch();
}

#if defined(PSQ_MID_TRACE)
}  // extern "C"
namespace psq { void mid_trace_read(unsigned long long* out); }
extern "C" {
// development aid (-DPSQ_MID_TRACE builds only; not declared in psqrt.h): out[2][128][16] %globaltimer stamps
int psqrt_debug_trace(unsigned long long* out) { psq::mid_trace_read(out); return 0; }
#endif

int psqrt_count_nonfinite(const double* x, int64_t rows, int64_t row_len, int64_t* counts, void* stream) {
  if (!x || !counts || rows <= 0 || rows > 65535 || row_len <= 0) return PSQRT_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counts, 0, sizeof(int64_t) * (size_t)rows, st);
  long long bx = (row_len + 256 * 8 - 1) / (256 * 8);
  if (bx > 1024) bx = 1024;
  psq::k_count_nonfinite<<<dim3((unsigned)bx, (unsigned)rows, 1), 256, 0, st>>>(
      x, row_len, reinterpret_cast<unsigned long long*>(counts));
  return check_launch();
}

int psqrt_fp64_probe(double* out, int iters, double* flops_out, void* stream) {
  if (!out || iters <= 0) return PSQRT_EINVAL;
  const int ctas = 148 * 4, threads = 128;
  psq::k_fp64_probe<<<ctas, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999, 1e-3);
  if (flops_out) *flops_out = 2.0 * 8.0 * 16.0 * (double)iters * (double)ctas * (double)threads;
  return check_launch();
}

int psqrt_chol_update_batched(double* L, const double* V, int n, int k, double alpha, int64_t batch, void* stream) {
  const LaunchN* ln = table_for(n);
  if (!ln) return (n >= 1 && n <= 16) ? psqrt_chol_update_generic(L, V, n, k, alpha, batch, stream) : PSQRT_EUNSUPPORTED;
  if (!L || !V || k < 0 || batch <= 0) return PSQRT_EINVAL;
  ln->chol_update(L, V, k, alpha, batch, (cudaStream_t)stream);
  return check_launch();
}

}  // extern "C"
