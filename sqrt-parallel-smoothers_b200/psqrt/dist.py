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


class TimeShardedSmoother:
    """One filter + smoother pass over a sequence time-sharded across the ranks of `group`."""

    def __init__(self, nx: int, ny: int, T_local: int, device=None, group=None, ops=None, chunk_len: int = 0):
        if ops is None:
            from . import _lib as ops
        self.ops, self.nx, self.ny, self.T, self.group, self.chunk_len = ops, nx, ny, T_local, group, chunk_len
        self.world, self.rank = _world(group)
        self.device = device

    def filter_smoother(self, ssm, y, m0, L0, *, smooth: bool = True, loglik: bool = False):
        """ssm: psqrt._lib.LinearizedSSM of THIS shard; y [B, T_local, ny]; m0 [B, nx], L0 [B, nx, nx]
        (lower triangular; the prior of the whole sequence, identical on all ranks).
        Returns (fm, fL, sm, sL, ell): local trajectories [B, T_local + 1, ...]; ell is the
        log-likelihood of the WHOLE sequence (same value on every rank) or None."""
        ops, R, r = self.ops, self.world, self.rank
        ftotal = ops.filter_reduce(ssm, y, self.nx, chunk_len=self.chunk_len)                 # [B, nf_filter]
        totals = _all_gather(ftotal, R, self.group)                                          # [R, B, nf]
        cm, cL = ops.carry_filter(totals, r, m0.contiguous(), L0.contiguous())
        fm, fL, ell, stotal = ops.filter_apply(ssm, y, cm, cL, smooth=smooth, loglik=loglik,
                                               chunk_len=self.chunk_len)
        if loglik and R > 1:
            dist.all_reduce(ell, op=dist.ReduceOp.SUM, group=self.group)
        if not smooth:
            return fm, fL, None, None, ell
        B = y.shape[0]
        nfs = stotal.shape[-1]
        payload = torch.cat([stotal, fm[:, -1], fL[:, -1].reshape(B, -1)], dim=-1)           # [B, nfs + nx + nx^2]
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
