"""Multi-GPU drivers: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Time sharding.  A sequence of R * T_local steps is cut into R contiguous
shards, one per rank.  A prefix scan decomposes into  local scan -> exchange of R shard totals ->
carry application, so one pass costs two latency-bound all-gathers of tiny payloads
(R x (2 nx^2 + 3 nx) and R x ((3 nx^2 + 3 nx)/2 + nx + nx^2) doubles) and, when the log-likelihood
is requested, one scalar all-reduce:

    1. psqrt_filter_reduce    local chunk summaries + shard total (A, b, U, eta, Z)
    2. all_gather(total)      every rank folds the totals of ranks < r into the prior: carry-in state
    3. psqrt_filter_apply     filtered trajectory of the shard (+ smoothing elements reduced to the
                              shard's smoothing total (g, E, D))
    4. all_gather(smoothing total ++ last filtered state)   carry = fold of ranks > r onto x_T
    5. psqrt_smoother_apply   smoothed trajectory of the shard

Local trajectories have T_local + 1 entries: entry 0 is the carry-in state (= the previous rank's last
entry) and entry T_local the last state of the shard, so neighbouring shards overlap by one entry and
an iterated smoother needs no halo exchange: the next nominal trajectory is the local smoothed
trajectory itself (transition linearised at entries [:-1], observation at entries [1:],
parallel/_filtering.py:103-104).

Batch sharding: independent sequences (Monte-Carlo runs) are dealt round-robin to the ranks; there
is no data-path collective, only a gather of scalars.

The collectives go through torch.distributed on the current stream's device; `ops` is the kernel
backend (psqrt._lib by default -- the CPU tests inject a NumPy stand-in to exercise this host logic
with the gloo backend).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from ._base import MVNSqrt


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather(x: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """[...]-> [world, ...] (rank-major)."""
    if world == 1:
        return x.unsqueeze(0)
    x = x.contiguous()
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)      # concatenation along dim 0, rank-major
    return out.view((world,) + tuple(x.shape))


class PeerExchange:
    """The two exchanges of a time-sharded pass over peer-mapped memory instead of NCCL, FUSED into the kernels that
    produce and consume the shard totals (include/psqrt.h, "staged calls"; csrc/psqrt_coop2.cuh).

    Every rank owns one exchange buffer (torch symmetric memory: CUDA IPC / fabric handles, mapped into every
    peer of the node).  The CTA that finishes the mid-level scan of a rank stores the shard total straight into slot
    `rank` of EVERY rank's buffer over NVLink and publishes the pass number there; the consumer's carry kernel waits
    for the pass number of ALL ranks, then folds the totals it needs (log-depth) out of local memory.  No
    collective call, no separate push / wait launches, no host synchronisation, no per-pass host argument (pass
    numbers are counted on the device), so a whole pass is captured in one CUDA graph.

    Slots are double-buffered by pass parity, and since every carry waits for all ranks no rank can be more than one
    pass ahead of a reader of its slot -- with or without the smoother phase."""

    def __init__(self, world, rank, B, nx, device, group=None, ops=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        if ops is None:
            from . import _lib as ops
        self.R, self.rank, self.B, self.nx = world, rank, B, nx
        n_words, self.filter_phase, self.smoother_phase = ops.peer_layout(nx, world, B)
        grp = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty(n_words, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.peer_bufs = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
        for ph in (self.filter_phase, self.smoother_phase):
            ph.bufs = ctypes.c_void_p(self.peer_bufs.data_ptr())
            ph.rank = rank
        self.device = device
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every buffer is zeroed before anyone publishes


class TimeShardedSmoother:
    """One filter + smoother pass over a sequence time-sharded across the ranks of `group`.

    exchange = "peer": shard totals travel by P2P stores into peer-mapped buffers (PeerExchange; falls back to
    NCCL if the symmetric-memory rendezvous is not available); "nccl": two all-gathers."""

    def __init__(self, nx: int, ny: int, T_local: int, device=None, group=None, ops=None, chunk_len: int = 0,
                 exchange: str = "nccl"):
        if ops is None:
            from . import _lib as ops
        self.ops, self.nx, self.ny, self.T, self.group, self.chunk_len = ops, nx, ny, T_local, group, chunk_len
        self.world, self.rank = _world(group)
        self.device = device
        self.exchange = exchange if self.world > 1 else "nccl"
        self._peer = None
        self.exchange_error = None

    def _peer_exchange(self, B):
        if self._peer is None or self._peer.B != B:
            err = None
            try:
                peer = PeerExchange(self.world, self.rank, B, self.nx, self.device, self.group, ops=self.ops)
            except Exception as e:            # no symmetric memory on this system
                peer, err = None, repr(e)
            # the ranks must agree: if the rendezvous failed anywhere, everybody uses the NCCL all-gathers
            # (a rank spinning on a flag while another sits in an all-gather would hang the job)
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                import warnings
                self.exchange, self.exchange_error = "nccl", err or "peer exchange unavailable on another rank"
                warnings.warn(f"psqrt: exchange='peer' requested but not available ({self.exchange_error}); "
                              f"using NCCL all-gathers")
                return None
            self._peer = peer
        return self._peer

    def filter_smoother(self, ssm, y, m0, L0, *, smooth: bool = True, loglik: bool = False):
        """ssm: psqrt._lib.LinearizedSSM of THIS shard; y [B, T_local, ny]; m0 [B, nx], L0 [B, nx, nx]
        (lower triangular; the prior of the whole sequence, identical on all ranks).
        Returns (fm, fL, sm, sL, ell): local trajectories [B, T_local + 1, ...]; ell is the
        log-likelihood of the WHOLE sequence (same value on every rank) or None."""
        ops, R, r = self.ops, self.world, self.rank
        B = y.shape[0]
        peer = self._peer_exchange(B) if self.exchange == "peer" else None
        m0, L0 = m0.contiguous(), L0.contiguous()
        if peer is not None:
            # exchange fused into the kernels: K2's last CTA publishes the total, the carry kernel waits + folds
            ops.filter_reduce(ssm, y, self.nx, chunk_len=self.chunk_len, peer=peer.filter_phase)
            cm, cL = ops.carry_filter(None, r, m0, L0, peer=peer.filter_phase)
        else:
            ftotal = ops.filter_reduce(ssm, y, self.nx, chunk_len=self.chunk_len)             # [B, nf_filter]
            totals = _all_gather(ftotal, R, self.group)                                      # [R, B, nf]
            cm, cL = ops.carry_filter(totals, r, m0, L0)
        fm, fL, ell, stotal = ops.filter_apply(ssm, y, cm, cL, smooth=smooth, loglik=loglik, chunk_len=self.chunk_len,
                                               **({"peer": peer.smoother_phase} if (peer is not None and smooth) else {}))
        if loglik and R > 1:
            dist.all_reduce(ell, op=dist.ReduceOp.SUM, group=self.group)
        if not smooth:
            return fm, fL, None, None, ell
        nfs = stotal.shape[-1]
        if peer is not None:
            # the carry kernel publishes this shard's smoothing total and last filtered state, waits for all ranks, folds
            sm_c, sL_c = ops.carry_smoother(stotal, r, R, fm[:, -1].contiguous(), fL[:, -1].contiguous(),
                                            peer=peer.smoother_phase)
        else:
            payload = torch.cat([stotal, fm[:, -1], fL[:, -1].reshape(B, -1)], dim=-1)       # [B, nfs + nx + nx^2]
            gathered = _all_gather(payload, R, self.group)
            stotals = gathered[:, :, :nfs].contiguous()
            mT = gathered[R - 1, :, nfs:nfs + self.nx].contiguous()
            LT = gathered[R - 1, :, nfs + self.nx:].reshape(B, self.nx, self.nx).contiguous()
            sm_c, sL_c = ops.carry_smoother(stotals, r, R, mT, LT)
        sm, sL = ops.smoother_apply(ssm, fm, fL, sm_c, sL_c, write_terminal=True, chunk_len=self.chunk_len)
        return fm, fL, sm, sL, ell


def shard_bounds(T: int, world: int, rank: int):
    """Contiguous time shard [t0, t1) of rank `rank`: the first T % world ranks get one extra step."""
    base, rem = divmod(T, world)
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def batch_indices(n_items: int, world: int, rank: int):
    """Round-robin deal of independent sequences to ranks (config 5: 100 runs over 8 GPUs -> 13/12 each)."""
    return list(range(rank, n_items, world))


def filter_smoother_sharded(observations_local, x0: MVNSqrt, transition_model, observation_model,
                            linearization_method: Callable, nominal_local: Optional[MVNSqrt] = None,
                            return_loglikelihood: bool = False, group=None, ops=None, linearize=None):
    """psqrt.methods.filter_smoother for a time-sharded sequence: every rank passes ITS observations
    [T_local, ny] and (optionally) its local nominal trajectory [T_local + 1, ...] (entry 0 = the state
    before the shard's first step).  Returns the local (filtered, smoothed[, ell]) trajectories."""
    from . import methods
    if linearize is None:
        linearize = methods._linearize
    dev = observations_local.device
    T, ny = observations_local.shape
    nx = x0.mean.shape[-1]
    if nominal_local is None:
        nominal_local = methods._default_nominal(T + 1, nx, dev)
    ssm = linearize(linearization_method, transition_model, observation_model, nominal_local)
    sharded = TimeShardedSmoother(nx, ny, T, device=dev, group=group, ops=ops)
    if ops is None:
        from . import _lib
        L0 = _lib.tria(x0.chol)
    else:
        L0 = ops.tria(x0.chol)
    fm, fL, sm, sL, ell = sharded.filter_smoother(ssm, observations_local[None].contiguous(), x0.mean[None], L0[None],
                                                  smooth=True, loglik=return_loglikelihood)
    if sharded.rank == 0:      # entry 0 of the whole trajectory is x0 itself (parallel/_filtering.py:45-46)
        fm[0, 0].copy_(x0.mean)
        fL[0, 0].copy_(x0.chol)
    filt, smo = MVNSqrt(fm[0], fL[0]), MVNSqrt(sm[0], sL[0])
    if return_loglikelihood:
        return filt, smo, ell[0]
    return filt, smo


def iterated_smoothing_sharded(observations_local, x0: MVNSqrt, transition_model, observation_model,
                               linearization_method: Callable, init_nominal_local: Optional[MVNSqrt] = None,
                               n_iter: int = 10, return_loglikelihood: bool = False, group=None, ops=None,
                               linearize=None):
    """Time-sharded iterated smoother with a fixed iteration count (criterion `lambda i, *_: i < n_iter`
    of parsmooth.methods.iterated_smoothing, methods.py:54-76): the nominal trajectory stays sharded
    between iterations, only the two all-gathers per pass cross the GPUs."""
    kw = dict(group=group, ops=ops, linearize=linearize)
    if init_nominal_local is None:
        _, init_nominal_local = filter_smoother_sharded(observations_local, x0, transition_model, observation_model,
                                                        linearization_method, None, **kw)
    nominal = init_nominal_local
    for _ in range(n_iter):
        _, nominal = filter_smoother_sharded(observations_local, x0, transition_model, observation_model,
                                             linearization_method, nominal, **kw)
    if return_loglikelihood:
        filt, _, ell = filter_smoother_sharded(observations_local, x0, transition_model, observation_model,
                                               linearization_method, nominal, True, **kw)
        del filt
        return nominal, ell
    return nominal


# ---------------------------------------------------------------------------------------------------
# Batch of independent sequences (BASELINE.json configs[4]: Monte-Carlo runs).  The batch is the outermost
# axis of every kernel (blockIdx.y), so B short sequences fill the GPU in ONE pass per iteration instead of B
# latency-bound passes; across GPUs the runs are dealt round-robin (batch_indices) with no collective.
# ---------------------------------------------------------------------------------------------------
def _linearize_batched(lin, transition_model, observation_model, nominal: MVNSqrt):
    """methods._linearize for a nominal trajectory with a leading batch axis [B, T + 1, ...]: transition at
    nominal[:, :-1], observation at nominal[:, 1:] (parallel/_filtering.py:103-104, 117-119)."""
    from . import methods
    from ._lib import LinearizedSSM
    F, cholQ, b = lin(transition_model, MVNSqrt(nominal.mean[:, :-1], nominal.chol[:, :-1]))
    cholQ = methods._lower(cholQ)
    H, cholR, c = lin(observation_model, MVNSqrt(nominal.mean[:, 1:], nominal.chol[:, 1:]))
    return LinearizedSSM(F, cholQ, b, H, cholR, c)


def filter_smoother_batched(observations, x0: MVNSqrt, transition_model, observation_model,
                            linearization_method: Callable, nominal: Optional[MVNSqrt] = None,
                            return_loglikelihood: bool = False):
    """psqrt.methods.filter_smoother for B independent sequences at once: observations [B, T, ny], x0 shared
    ([nx], [nx, nx]) or per sequence ([B, nx], [B, nx, nx]), nominal [B, T + 1, ...] (default: zeros / identity).
    Returns (filtered, smoothed[, ell [B]]) with a leading batch axis; every sequence gets exactly what the
    unbatched call returns for it."""
    from . import _lib, methods
    dev = methods._device()
    ys = methods._t(observations, dev)
    B, T, _ = ys.shape
    x0 = methods._mvn(x0, dev)
    nx = x0.mean.shape[-1]
    transition_model = methods._model(transition_model, dev)
    observation_model = methods._model(observation_model, dev)
    if nominal is None:
        mean = torch.zeros((B, T + 1, nx), dtype=torch.float64, device=dev)
        nominal = MVNSqrt(mean, torch.eye(nx, dtype=torch.float64, device=dev).expand(B, T + 1, nx, nx))
    else:
        nominal = methods._mvn(nominal, dev)
    ssm = _linearize_batched(linearization_method, transition_model, observation_model, nominal)
    m0 = x0.mean.expand(B, nx).contiguous()
    L0 = methods._prior_factor(x0.chol).expand(B, nx, nx).contiguous()
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, ys, m0, L0, smooth=True, loglik=return_loglikelihood)
    fm[:, 0].copy_(x0.mean.expand(B, nx))          # entry 0 is x0 itself (parallel/_filtering.py:45-46)
    fL[:, 0].copy_(x0.chol.expand(B, nx, nx))
    if return_loglikelihood:
        return MVNSqrt(fm, fL), MVNSqrt(sm, sL), ell
    return MVNSqrt(fm, fL), MVNSqrt(sm, sL)


def iterated_smoothing_batched(observations, x0: MVNSqrt, transition_model, observation_model,
                               linearization_method: Callable, init_nominal: Optional[MVNSqrt] = None,
                               n_iter: int = 10, return_loglikelihood: bool = False):
    """parsmooth.methods.iterated_smoothing (methods.py:54-76) with the fixed-count criterion
    `lambda i, *_: i < n_iter` for B independent sequences at once (notebooks/robustness_100runs.py runs them one
    after the other).  init_nominal: [B, T + 1, ...] or a single [T + 1, ...] trajectory shared by all sequences."""
    from . import methods
    dev = methods._device()
    ys = methods._t(observations, dev)
    B = ys.shape[0]
    x0 = methods._mvn(x0, dev)
    transition_model = methods._model(transition_model, dev)
    observation_model = methods._model(observation_model, dev)
    args = (ys, x0, transition_model, observation_model, linearization_method)
    if init_nominal is None:
        _, nominal = filter_smoother_batched(*args, None)
    else:
        nominal = methods._mvn(init_nominal, dev)
        if nominal.mean.dim() == 2:
            nominal = MVNSqrt(nominal.mean.expand(B, *nominal.mean.shape), nominal.chol.expand(B, *nominal.chol.shape))
    # methods.fixed_point: one application, then again while criterion(i, ...) holds, i = 1, 2, ... -> n_iter in all
    for _ in range(n_iter):
        _, nominal = filter_smoother_batched(*args, nominal)
    if return_loglikelihood:
        _, _, ell = filter_smoother_batched(*args, nominal, True)
        return nominal, ell
    return nominal


def iterated_smoothing_batch_sharded(observations, x0: MVNSqrt, transition_model, observation_model,
                                     linearization_method: Callable, init_nominal: Optional[MVNSqrt] = None,
                                     n_iter: int = 10, return_loglikelihood: bool = True, group=None, smoother=None):
    """BASELINE.json configs[4] across GPUs (notebooks/robustness_100runs.py:52-72 runs the 100 data sets one after
    the other): `observations` [n_runs, T, ny] is the SAME array on every rank (or any per-run indexable); run i goes
    to rank i % world (batch_indices), every rank smooths its share as ONE batch (iterated_smoothing_batched) and
    there is no data-path collective -- only the per-run log-likelihoods are gathered at the end.

    Returns (local_indices, local_nominal [n_local, T + 1, ...], ell_all [n_runs] or None); ell_all is identical on
    every rank, in run order.  `smoother` replaces iterated_smoothing_batched (the CPU tests inject a NumPy stand-in
    to exercise this host logic over gloo)."""
    world, rank = _world(group)
    n_runs = len(observations)
    idx = batch_indices(n_runs, world, rank)
    if smoother is None:
        smoother = iterated_smoothing_batched
    if len(idx) > 0:
        obs_local = observations[idx] if hasattr(observations, "shape") else torch.stack([observations[i] for i in idx])
        nominal = init_nominal
        if nominal is not None and nominal.mean.dim() == 3 and nominal.mean.shape[0] == n_runs:   # per-run nominals
            nominal = MVNSqrt(nominal.mean[idx], nominal.chol[idx])
        out = smoother(obs_local, x0, transition_model, observation_model, linearization_method, nominal,
                       n_iter=n_iter, return_loglikelihood=return_loglikelihood)
        local_nominal, ell_local = out if return_loglikelihood else (out, None)
    else:
        local_nominal, ell_local = None, None
    if not return_loglikelihood:
        return idx, local_nominal, None
    # gather of scalars: every rank contributes a fixed-size vector (ceil(n_runs / world) entries, NaN-padded)
    per = (n_runs + world - 1) // world
    if ell_local is not None:
        dev, dtype = ell_local.device, ell_local.dtype
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        dtype = torch.float64
    mine = torch.full((per,), float("nan"), dtype=dtype, device=dev)
    if ell_local is not None:
        mine[:len(idx)] = ell_local.reshape(-1)
    gathered = _all_gather(mine, world, group)                  # [world, per]
    ell_all = torch.empty((n_runs,), dtype=dtype, device=dev)
    for r in range(world):
        ids = batch_indices(n_runs, world, r)
        if ids:
            ell_all[ids] = gathered[r, :len(ids)]
    return idx, local_nominal, ell_all


def nonfinite_runs(trajectory: MVNSqrt) -> torch.Tensor:
    """[B] bool: the runs of a batch [B, T + 1, ...] whose trajectory contains a NaN or Inf -- the failure count the
    robustness sweeps report (notebooks/robustness_100runs.py:41-77); counted on the device (psqrt_count_nonfinite)."""
    from . import _lib
    return (_lib.count_nonfinite(trajectory.mean) + _lib.count_nonfinite(trajectory.chol)) > 0
