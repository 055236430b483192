/* psqrt.h -- C ABI of libpsqrt.so: the B200 (sm_100a) square-root parallel Kalman filter /
 * RTS smoother path of EEA-sensors/sqrt-parallel-smoothers ("parsmooth").
 *
 * The reference has no FFI of its own (it is pure Python on JAX); its seams for this path are
 *   - the API facade                parsmooth/methods.py:14-76
 *   - the linearization protocol    parsmooth/parallel/_filtering.py:117-119, _smoothing.py:73
 *   - the associative-scan seam     parsmooth/parallel/_filtering.py:34-35, _smoothing.py:33-34
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the
 * ctypes binding (used by the PyTorch host layer in this repo) and the XLA-FFI shim a
 * parsmooth maintainer would add to call the same symbols from JAX.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to caller-owned, contiguous, 8-byte aligned fp64 data
 *     (row-major, batch outermost, then time); nothing is allocated or freed by the library;
 *   - scratch comes from the caller: size it with psqrt_workspace_bytes();
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never
 *     synchronises the device and never throws;
 *   - return value: PSQRT_OK or a negative PSQRT_E* code (psqrt_error_string() names it);
 *   - square-root factors are compared/defined up to a right orthogonal factor, exactly like
 *     the reference's tria() (parsmooth/_utils.py:22-24): only L L^T is meaningful;
 *   - trajectories have T+1 entries (index 0 = prior / carry-in state).
 */
#ifndef PSQRT_H_
#define PSQRT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSQRT_VERSION 100 /* 0.1.0 */

#define PSQRT_OK 0
#define PSQRT_EINVAL (-1)       /* null pointer / non-positive size / bad enum            */
#define PSQRT_EUNSUPPORTED (-2) /* (nx, ny) not in the compiled set (psqrt_supported())   */
#define PSQRT_EWORKSPACE (-3)   /* workspace too small                                    */
#define PSQRT_ECUDA (-4)        /* a launch failed (cudaGetLastError() != cudaSuccess)    */

/* op codes for psqrt_workspace_bytes */
#define PSQRT_OP_FILTER_SMOOTHER 0 /* also covers FILTER, SMOOTHER and the staged calls    */
#define PSQRT_OP_ELEMENT_SCAN 1    /* psqrt_filter_scan / psqrt_smoother_scan             */

/* Linearised state-space model, what linearization_method(model, nominal) returns for every
 * step (parsmooth/parallel/_filtering.py:117-119):
 *     x_{k+1} = F_k x_k + b_k + N(0, cholQ_k cholQ_k^T)        F [nx,nx] cholQ [nx,nx] b [nx]
 *     y_k     = H_k x_{k+1} + c_k + N(0, cholR_k cholR_k^T)    H [ny,nx] cholR [ny,ny] c [ny]
 * cholQ must be LOWER TRIANGULAR in the pass / staged / smoother entry points (entries above its diagonal
 * are ignored; any other square root of Q is first passed through psqrt_tria_batched by the caller) --
 * that triangle shortens every reflector of the predict step; cholR may be any square root.
 * *_ts is the stride in doubles between consecutive time steps (0 = time-invariant),
 * *_bs the stride between sequences of a batch (0 = shared).  H, cholR, c may be NULL for
 * smoother-only calls. */
typedef struct psqrt_ssm {
  const double *F, *cholQ, *b, *H, *cholR, *c;
  int64_t F_ts, cholQ_ts, b_ts, H_ts, cholR_ts, c_ts;
  int64_t F_bs, cholQ_bs, b_bs, H_bs, cholR_bs, c_bs;
  /* Optional HOST mirrors of the six arrays (NULL when unknown).  When the model is shared by all
   * steps and sequences (every stride 0) and its host mirrors are given, the sweeps carry the
   * model by value in their kernel parameters, i.e. as constant-bank operands. */
  const double *hF, *hcholQ, *hb, *hH, *hcholR, *hc;
  /* Optional: a built-in model linearised INSIDE the sweeps (the per-step arrays are then never formed in HBM; this is
   * the fused counterpart of psqrt_linearize_builtin + the pass, SURVEY 8b's psqrt_pass_fused).  fused_model =
   * PSQRT_FUSED_CT_BEARINGS: coordinated-turn transition + two-bearings observation with the extended (first-order
   * Taylor) linearization, nx = 5, ny = 2: step k uses F, b at nom_m[k] and H, c at nom_m[k + 1]
   * (parallel/_filtering.py:103-104,117-119).  Then F, b, H, c are ignored (may be NULL); the noise is time-invariant
   * and read from the HOST mirrors hcholQ [5,5] (lower), hb = m_q [5], hcholR [2,2], hc = m_r [2] (hcholR / hc may be
   * NULL in smoother-only calls); fused_params (HOST) = {dt, s1x, s1y, s2x, s2y}; nom_m (DEVICE) [B, T+1, 5] with
   * batch stride nom_bs (0 = shared).  Accepted by psqrt_filter_smoother and the staged calls
   * (psqrt_filter_reduce / psqrt_filter_apply / psqrt_smoother_apply); PSQRT_EINVAL elsewhere. */
  int32_t fused_model, fused_reserved;
  const double* nom_m;
  int64_t nom_bs;
  const double* fused_params;
} psqrt_ssm;
#define PSQRT_FUSED_NONE 0
#define PSQRT_FUSED_CT_BEARINGS 1

typedef struct psqrt_plan {
  int32_t chunk_len;     /* K: consecutive steps handled by one thread                    */
  int64_t n_chunks;      /* P = ceil(T / K)                                               */
  int64_t n_chunks_pad;  /* P rounded up to the CTA size                                  */
  int64_t n_warps;       /* n_chunks_pad / 32 (items of the mid-level scan)               */
  int32_t nf_filter;     /* doubles in one packed filtering element  2 nx^2 + 3 nx        */
  int32_t nf_smoother;   /* doubles in one packed smoothing element (3 nx^2 + 3 nx) / 2   */
} psqrt_plan;

int psqrt_version(void);
const char* psqrt_error_string(int code);
/* 1 if kernels for (nx, ny) are compiled in; ny = 0 asks about smoother-only support. */
int psqrt_supported(int nx, int ny);
/* Chunking used for a problem size; chunk_len = 0 lets the library choose.  The library's choice depends on the state
 * dimension and, for nx <= 4, on whether the TRANSITION part of the model is time-invariant with host mirrors (it then
 * travels by value and the sweeps want fewer, longer chunks): psqrt_get_plan answers for a model read from device
 * memory, psqrt_get_plan_ssm for the model given.  Every stage of a pass makes the same choice from the same model. */
int psqrt_get_plan(int nx, int ny, int64_t T, int64_t batch, int chunk_len, psqrt_plan* out);
int psqrt_get_plan_ssm(const psqrt_ssm* ssm, int nx, int ny, int64_t T, int64_t batch, int chunk_len, psqrt_plan* out);
size_t psqrt_workspace_bytes(int op, int nx, int ny, int64_t T, int64_t batch, int chunk_len);

/* ---- whole pass: filtering + smoothing (methods.py:38-47 filter_smoother, parallel=True,
 *      MVNSqrt inputs; parallel/_filtering.py:13-61 + parallel/_smoothing.py:14-44) ---------
 * y [B,T,ny]; m0 [B,nx]; L0 [B,nx,nx] LOWER-TRIANGULAR sqrt of the prior covariance;
 * fm [B,T+1,nx], fL [B,T+1,nx,nx]: filtered trajectory (index 0 = (m0, L0));
 * sm, sL: smoothed trajectory, same shapes (both NULL = filter only);
 * ell [B] log-likelihood (NULL = skip; parallel/_filtering.py:53-60). */
int psqrt_filter_smoother(const psqrt_ssm* ssm, const double* y, const double* m0, const double* L0,
                          int nx, int ny, int64_t T, int64_t batch, int chunk_len,
                          double* fm, double* fL, double* sm, double* sL, double* ell,
                          void* ws, size_t ws_bytes, void* stream);

/* smoothing(...) on an existing filtered trajectory (parallel/_smoothing.py:14-57).  fL must
 * be lower triangular (what psqrt_filter_smoother writes). */
int psqrt_smoother(const psqrt_ssm* ssm, const double* fm, const double* fL, int nx, int64_t T,
                   int64_t batch, int chunk_len, double* sm, double* sL, void* ws, size_t ws_bytes,
                   void* stream);

/* ---- staged calls for a time-sharded run (one shard per GPU) --------
 * 1. psqrt_filter_reduce  : local chunk summaries + shard total  -> ftotal [B, nf_filter]
 * 2. (exchange of the R totals)  psqrt_carry_filter: the totals of ranks < rank folded into x0
 * 3. psqrt_filter_apply   : filtered trajectory of the shard from the carry-in state, ell
 *                           partial; with stotal != NULL also the smoothing reduce -> stotal
 * 4. (exchange of the R smoothing totals + the last rank's terminal state)  psqrt_carry_smoother
 * 5. psqrt_smoother_apply : smoothed trajectory of the shard from the carry-in state; fm / fL must be the arrays
 *                           the psqrt_filter_apply of the same pass wrote (its workspace holds their packed copy).
 * The same workspace must be passed to 1, 3 and 5.  The carries are log-depth scans of the R totals.
 *
 * The exchange itself (steps 2 and 4) is either the caller's (peer = NULL: e.g. two NCCL all-gathers; totals
 * [R, B, nf] are then passed to the carry calls) or FUSED into these kernels over peer-mapped memory
 * (peer != NULL; NVLink / NVSwitch P2P, e.g. CUDA IPC or torch symmetric memory):
 *   - every rank owns ONE exchange buffer of psqrt_peer_layout(...) 8-byte words, zeroed once, mapped into all
 *     ranks; `bufs` is a DEVICE array of the n_ranks base pointers as seen from this rank;
 *   - filter phase: the CTA that finishes the mid-level scan of psqrt_filter_reduce stores the shard total straight
 *     into slot `rank` of EVERY rank's buffer, then publishes the pass number there; psqrt_carry_filter waits, on the
 *     GPU and in stream order, until ALL ranks have published that pass number and folds the totals it needs out of
 *     its local buffer (totals ignored);
 *   - smoother phase: psqrt_filter_apply only WRITES the shard's smoothing total (stotal; its `peer` argument is
 *     validated, nothing is sent -- the total may then come from the scan hidden inside the forward sweep);
 *     psqrt_carry_smoother is given this rank's OWN total (totals = stotal [B, nf_smoother]) and last filtered state
 *     (mT, LT = fm[:, T], fL[:, T]), publishes them in every rank's buffer, waits for all ranks and folds the totals of
 *     the later shards onto the last rank's published state.
 * Pass numbers are counted on the device (one counter per sequence and phase) and the slots are double-buffered by
 * pass parity: no per-pass host argument (a whole time-sharded pass replays from one CUDA graph), nothing is ever
 * reset, and -- because every carry waits for all ranks -- no rank can run more than one pass ahead of a reader of
 * its slot, also in filter-only passes. */
typedef struct psqrt_peer {
  void* const* bufs;        /* DEVICE array [n_ranks]: every rank's exchange buffer, mapped into this rank            */
  int32_t rank, n_ranks;
  int64_t batch;
  int64_t flags_off, ctr_off, data_off, slot, payload;   /* 8-byte words; filled by psqrt_peer_layout               */
} psqrt_peer;
/* Fills the offsets of the two phases for (nx, n_ranks, batch) (bufs / rank are left to the caller) and returns the
 * size of one rank's exchange buffer in 8-byte words (0: unsupported nx). */
int64_t psqrt_peer_layout(int nx, int n_ranks, int64_t batch, psqrt_peer* filter_phase, psqrt_peer* smoother_phase);

int psqrt_filter_reduce(const psqrt_ssm* ssm, const double* y, int nx, int ny, int64_t T, int64_t batch,
                        int chunk_len, double* ftotal, void* ws, size_t ws_bytes, const psqrt_peer* peer,
                        void* stream);
int psqrt_carry_filter(const double* totals /*[R,B,nf_filter]*/, int rank, int64_t batch, int nx,
                       const double* m0, const double* L0, double* carry_m, double* carry_L,
                       const psqrt_peer* peer, void* stream);
int psqrt_filter_apply(const psqrt_ssm* ssm, const double* y, const double* carry_m, const double* carry_L,
                       int nx, int ny, int64_t T, int64_t batch, int chunk_len, double* fm, double* fL,
                       double* ell, double* stotal, void* ws, size_t ws_bytes, const psqrt_peer* peer,
                       void* stream);
int psqrt_carry_smoother(const double* totals /*[R,B,nf_smoother]*/, int rank, int n_ranks, int64_t batch,
                         int nx, const double* mT, const double* LT, double* carry_m, double* carry_L,
                         const psqrt_peer* peer, void* stream);
int psqrt_smoother_apply(const psqrt_ssm* ssm, const double* fm, const double* fL, const double* carry_m,
                         const double* carry_L, int write_terminal, int nx, int64_t T, int64_t batch,
                         int chunk_len, double* sm, double* sL, void* ws, size_t ws_bytes, void* stream);

/* ---- element-level seams (for callers that build or consume the associative elements) -----
 * psqrt_filter_elements   parallel/_filtering.py:100-146  (prior folded into step 0 when m0 != NULL;
 *                          here L0 may be any dense square-root factor)
 * psqrt_filter_scan       jax.lax.associative_scan(vmap(sqrt_filtering_operator)) at _filtering.py:34-35:
 *                          inclusive prefixes, outputs (b, U) = filtered means / factors, [B,T,...]
 * psqrt_smoother_elements parallel/_smoothing.py:47-57,72-85 (T+1 elements, the last is (m_T, 0, L_T))
 * psqrt_smoother_scan     reverse associative_scan at _smoothing.py:33-34, outputs (g, D), n = number of elements
 * psqrt_loglik_terms      parallel/_filtering.py:149-154 per-step terms [B,T] (sum them for ell)
 * psqrt_filter_combine / psqrt_smoother_combine: one application of
 *                          parsmooth/parallel/_operators.py:43-77 / 104-125 to n element pairs. */
int psqrt_filter_elements(const psqrt_ssm* ssm, const double* y, const double* m0, const double* L0,
                          int nx, int ny, int64_t T, int64_t batch,
                          double* A, double* b, double* U, double* eta, double* Z, void* stream);
int psqrt_filter_scan(const double* A, const double* b, const double* U, const double* eta, const double* Z,
                      int nx, int64_t T, int64_t batch, int chunk_len, double* means, double* chols,
                      void* ws, size_t ws_bytes, void* stream);
int psqrt_smoother_elements(const psqrt_ssm* ssm, const double* fm, const double* fL, int nx, int64_t T,
                            int64_t batch, double* g, double* E, double* D, void* stream);
int psqrt_smoother_scan(const double* g, const double* E, const double* D, int nx, int64_t n, int64_t batch,
                        int chunk_len, double* means, double* chols, void* ws, size_t ws_bytes, void* stream);
int psqrt_loglik_terms(const psqrt_ssm* ssm, const double* y, const double* fm, const double* fL,
                       int nx, int ny, int64_t T, int64_t batch, double* terms, void* stream);
int psqrt_filter_combine(const double* A1, const double* b1, const double* U1, const double* eta1,
                         const double* Z1, const double* A2, const double* b2, const double* U2,
                         const double* eta2, const double* Z2, int nx, int64_t n,
                         double* A, double* b, double* U, double* eta, double* Z, void* stream);
int psqrt_smoother_combine(const double* g1, const double* E1, const double* D1, const double* g2,
                           const double* E2, const double* D2, int nx, int64_t n,
                           double* g, double* E, double* D, void* stream);

/* ---- math utilities (parsmooth/_utils.py) ------------------------------------------------
 * psqrt_tria_batched        _utils.py:22-24: A [batch, rows, cols] -> L [batch, rows, rows] lower
 *                           triangular with L L^T = A A^T (rows <= 16, any cols >= 1; rows in {1..6, 8} tuned)
 * psqrt_chol_update_batched _utils.py:13-19,39-81: L [batch,n,n] <- chol(L L^T + alpha sum_k v_k v_k^T),
 *                           V [batch,k,n], sequential over k, non-finite entries -> 0 */
int psqrt_tria_batched(const double* A, double* L, int rows, int cols, int64_t batch, void* stream);
int psqrt_chol_update_batched(double* L, const double* V, int n, int k, double alpha, int64_t batch,
                              void* stream);

/* ---- pathwise sampler (parsmooth/_pathwise_sampler.py:13-38,61-81,109-124) ----------------------
 * Joint samples of the smoothing distribution.  (g, E, D) are the n_elements = T + 1 smoothing elements
 * of psqrt_smoother_elements (which equal the sampler's (inc_m, gain, inc_L) and, last, (m_T, 0, L_T));
 * eps [n_elements, n_samples, nx] are standard normal draws, eps[0] makes the last state and eps[t + 1]
 * the increment of step t (same indexing as upstream, lines 71-79); samples [n_elements, n_samples, nx].
 * D enters with its diagonal made non-negative (column signs of tria are arbitrary).  Workspace:
 * psqrt_sampler_workspace_bytes. */
size_t psqrt_sampler_workspace_bytes(int nx, int64_t n_elements, int64_t n_samples);
int psqrt_sample_paths(const double* g, const double* E, const double* D, const double* eps, double* samples,
                       int nx, int64_t n_elements, int64_t n_samples, void* ws, size_t ws_bytes, void* stream);

/* ---- linearization of the built-in models as device code (one thread per time step) -----------
 * Replaces vmap(linearization_method(model, nominal)) of parallel/_filtering.py:110-119 and
 * _smoothing.py:50-55 for the models of the reference's tests / notebooks:
 *   PSQRT_MODEL_CT_TRANSITION         tests/bearings/bearings_utils.py:7-46    params {dt}                 5 -> 5
 *   PSQRT_MODEL_BEARINGS_OBSERVATION  tests/bearings/bearings_utils.py:49-69   params {s1x,s1y,s2x,s2y}    5 -> 2
 *   PSQRT_MODEL_RICKER_TRANSITION     notebooks/population_model.py:23-62      params {sqrt(Q)}            1 -> 1 (conditional moments)
 *   PSQRT_MODEL_POISSON_OBSERVATION   notebooks/population_model.py:84-129     params {lam}                1 -> 1 (conditional moments)
 * lin_id PSQRT_LIN_EXTENDED (linearization/_extended.py:51-70; xi/wm/wc/nom_L unused, chol written only for
 * conditional-moments models) or PSQRT_LIN_SLR (linearization/_sigma_points.py:25-100 with the unit sigma
 * points xi [n_points, n] and weights wm, wc [n_points] -- cubature: _cubature.py:63-85, Gauss-Hermite:
 * _gh.py:73-126 -- incl. the Cholesky downdates of _utils.py:13-19,39-81).
 * model_params is a HOST array; nom_m [count, n], nom_L [count, n, n], m_q [d], chol_q [d, d] (functional models),
 * outputs F [count, d, n], chol [count, d, d], b [count, d] are device arrays. */
#define PSQRT_MODEL_CT_TRANSITION 1
#define PSQRT_MODEL_BEARINGS_OBSERVATION 2
#define PSQRT_MODEL_RICKER_TRANSITION 3
#define PSQRT_MODEL_POISSON_OBSERVATION 4
#define PSQRT_LIN_EXTENDED 0
#define PSQRT_LIN_SLR 1
/* Non-finite entries per row of x [rows, row_len] -> counts [rows] (int64, device): the NaN-rate report of the robustness
 * sweeps (notebooks/robustness_100runs.py:41-77 counts runs whose result is NaN). */
int psqrt_count_nonfinite(const double* x, int64_t rows, int64_t row_len, int64_t* counts, void* stream);

int psqrt_linearize_builtin(int model_id, const double* model_params, int lin_id, const double* xi,
                            const double* wm, const double* wc, int n_points, const double* nom_m,
                            const double* nom_L, int64_t count, const double* m_q, const double* chol_q,
                            double* F, double* chol, double* b, void* stream);

/* ---- forward-mode tangent (JVP) of the pass and of its log-likelihood: the gradient path ---------------------
 * Replaces jax.jvp / jax.value_and_grad through parsmooth.methods.filter_smoother / filtering(..., return_loglikelihood=True)
 * (methods.py:14-47, 72-75; parallel/_filtering.py:13-61,149-154; parallel/_smoothing.py:14-57) and, iterated, the
 * implicit differentiation of the fixed point (parsmooth/_utils.py:103-146).  One tangent direction, one sequence.
 *
 * psqrt_ssm_tangent: the directional derivative of the linearised model, COVARIANCE form for the noises:
 *   dF [nx,nx], dQ [nx,nx] = d(cholQ cholQ^T) (symmetric), db [nx], dH [ny,nx], dR [ny,ny] = d(cholR cholR^T), dc [ny];
 *   *_ts = stride in doubles between time steps (0 = the same for every step); a NULL pointer is a zero tangent.
 * psqrt_filter_smoother_tangent: given the primal trajectories of psqrt_filter_smoother on the same inputs
 *   (fm, fL filtered with index 0 = the prior; sm, sL smoothed, NULL = filter only) and the tangent of the prior
 *   (dm0 [nx], dP0 [nx,nx] = d(L0 L0^T); NULL = 0), writes the tangents of the filtered moments dfm [T+1,nx],
 *   dfP [T+1,nx,nx] (= d(fL fL^T)), of the smoothed moments dsm, dsP (both NULL = skip) and of the log-likelihood
 *   dell [1] (NULL = skip).  The tangent recursions are affine maps (Phi, w, c, C) composed by an associative scan of
 *   matrix products (csrc/psqrt_tangent.cu); nothing is triangularised.
 * psqrt_cov_tangent_to_chol: dL [count,n,n] with d(L L^T) = dP for LOWER-triangular nonsingular L (what the pass writes).
 * psqrt_linearize_builtin_tangent: tangent of psqrt_linearize_builtin along (dnom_m, dnom_L [lower], dmodel_params,
 *   dm_q, dQ_q = d(chol_q chol_q^T)); any of them may be NULL (= 0).  Outputs dF [count,d,n], db [count,d] and dQ
 *   [count,d,d] (covariance form; for PSQRT_LIN_EXTENDED with a functional model dQ is not written -- it is dQ_q). */
typedef struct psqrt_ssm_tangent {
  const double *dF, *dQ, *db, *dH, *dR, *dc;
  int64_t dF_ts, dQ_ts, db_ts, dH_ts, dR_ts, dc_ts;
} psqrt_ssm_tangent;
size_t psqrt_tangent_workspace_bytes(int nx, int ny, int64_t T);
int psqrt_filter_smoother_tangent(const psqrt_ssm* ssm, const psqrt_ssm_tangent* dssm, const double* y, int nx, int ny,
                                  int64_t T, const double* fm, const double* fL, const double* sm, const double* sL,
                                  const double* dm0, const double* dP0, double* dfm, double* dfP, double* dsm,
                                  double* dsP, double* dell, void* ws, size_t ws_bytes, void* stream);
int psqrt_cov_tangent_to_chol(const double* L, const double* dP, double* dL, int n, int64_t count, void* stream);
/* Reverse mode for the log-likelihood at a fixed linearised model: d ell / d (every entry of the model), all steps at
 * once, for the cost of about one tangent pass.  The costates (lam_k, Lam_k) = d ell / d (m_k, P_k) obey the transposed
 * affine recursion backwards in time (again an associative scan of matrix products); they are written to lam [T+1,nx],
 * Lam [T+1,nx,nx] (index 0 = gradient w.r.t. the prior mean / covariance), and contracted per step into
 *   gF [T,nx,nx], gb [T,nx], gH [T,ny,nx], gc [T,ny]   and, in COVARIANCE form,   gQ [T,nx,nx], gR [T,ny,ny]
 * (d ell = <gQ, dQ> for symmetric dQ; w.r.t. a factor: 2 gQ cholQ).  fm, fL: the primal filtered trajectory of
 * psqrt_filter_smoother.  Workspace: psqrt_tangent_workspace_bytes.  One sequence. */
int psqrt_loglik_adjoint(const psqrt_ssm* ssm, const double* y, int nx, int ny, int64_t T, const double* fm,
                         const double* fL, double* lam, double* Lam, double* gF, double* gQ, double* gb, double* gH,
                         double* gR, double* gc, void* ws, size_t ws_bytes, void* stream);
int psqrt_linearize_builtin_tangent(int model_id, const double* model_params, const double* dmodel_params, int lin_id,
                                    const double* xi, const double* wm, const double* wc, int n_points,
                                    const double* nom_m, const double* nom_L, const double* dnom_m,
                                    const double* dnom_L, int64_t count, const double* dm_q, const double* dQ_q,
                                    double* dF, double* dQ, double* db, void* stream);

/* ---- generic path: any nx <= 16, ny <= 16, fp64 or fp32 (csrc/psqrt_generic.cu) -----------------------------
 * A literal statement of the reference's algorithm -- elements (parallel/_filtering.py:100-146, _smoothing.py:47-85),
 * operators (parallel/_operators.py:43-125) and the associative scan as a Hillis-Steele scan over element arrays in
 * HBM -- with runtime dimensions, one thread per time step; O(T log T) combines, nothing tuned.  It backs
 *   - psqrt_filter_smoother / psqrt_smoother / psqrt_workspace_bytes / psqrt_tria_batched / psqrt_chol_update_batched
 *     for the dimensions the tuned kernels do not cover (psqrt_supported() == 0): they fall back to it on their own;
 *   - the float32 mode of the robustness experiments (notebooks/robustness_100runs.py:7,41-77):
 *     psqrt_filter_smoother_f32, all arrays float, time_strides / batch_strides [6] in ELEMENTS for
 *     (F, cholQ, b, H, cholR, c), ell accumulated in double.
 * psqrt_filter_smoother_generic takes the arguments of psqrt_filter_smoother (y == NULL: smoother only, fm / fL are then
 * inputs; sm == sL == NULL: filter only); L0 and cholQ may be any square-root factors.  The staged, seam, sampler and
 * tangent entry points have no generic form.  Workspace: psqrt_generic_workspace_bytes(nx, T, batch, fp32). */
int psqrt_supported_generic(int nx, int ny);
size_t psqrt_generic_workspace_bytes(int nx, int64_t T, int64_t batch, int fp32);
int psqrt_filter_smoother_generic(const psqrt_ssm* ssm, const double* y, const double* m0, const double* L0, int nx,
                                  int ny, int64_t T, int64_t batch, double* fm, double* fL, double* sm, double* sL,
                                  double* ell, void* ws, size_t ws_bytes, void* stream);
int psqrt_filter_smoother_f32(const float* F, const float* cholQ, const float* b, const float* H, const float* cholR,
                              const float* c, const int64_t* time_strides, const int64_t* batch_strides,
                              const float* y, const float* m0, const float* L0, int nx, int ny, int64_t T,
                              int64_t batch, float* fm, float* fL, float* sm, float* sL, double* ell, void* ws,
                              size_t ws_bytes, void* stream);
int psqrt_tria_generic(const double* A, double* L, int rows, int cols, int64_t batch, void* stream);
int psqrt_chol_update_generic(double* L, const double* V, int n, int k, double alpha, int64_t batch, void* stream);

/* ---- measurement aid (bench.py): FP64 FMA throughput probe ------------------------------------
 * Launches 148 x 4 CTAs of 128 threads, each thread running 8 independent chains of `iters` x 16 dependent
 * DFMAs, and writes the flop count of the launch to *flops_out (host).  The caller times the launch with
 * CUDA events: flops / seconds is the FP64 roofline denominator SURVEY.md 8(d) asks to be measured.
 * out: device scratch of at least 148 * 512 doubles. */
int psqrt_fp64_probe(double* out, int iters, double* flops_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSQRT_H_ */
