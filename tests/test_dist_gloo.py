"""Host-side logic of the multi-GPU drivers (psqrt/dist.py) with world_size = 2 on CPU (gloo): time
sharding with two all-gathers per pass + scalar all-reduce, one-entry overlap of the local trajectories,
halo-free iterated smoothing, batch dealing.  The kernels are replaced by a NumPy stand-in built on the
oracle (tests/_np_backend.py); the CUDA kernels behind the same stage calls are checked on the GPU by
tests/test_gpu_parity.py::test_staged_calls_fake_ranks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parsmooth_np as O
from _cases import LLt, lgssm_case, oracle_from_ssm, rel_err

HERE = os.path.dirname(os.path.abspath(__file__))


def _setup(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_lgssm(rank, world, port, T_total, q):
    for p in (HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.join(os.path.dirname(HERE), "sqrt-parallel-smoothers_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        _setup(rank, world, port)
        import _np_backend as ops
        from psqrt import dist as pdist
        from psqrt._lib import LinearizedSSM
        case = lgssm_case(3, 2, T_total, seed=5)
        t0, t1 = pdist.shard_bounds(T_total, world, rank)
        g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
        ssm = LinearizedSSM(*[g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
        sh = pdist.TimeShardedSmoother(3, 2, t1 - t0, ops=ops)
        fm, fL, sm, sL, ell = sh.filter_smoother(ssm, g(case["ys"][t0:t1])[None], g(case["m0"])[None],
                                                 g(case["L0"])[None], smooth=True, loglik=True)
        q.put((rank, t0, t1, fm[0].numpy(), fL[0].numpy(), sm[0].numpy(), sL[0].numpy(), float(ell[0])))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


def _run(worker, world, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=worker, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] != "error", r
    return sorted(res, key=lambda r: r[0])


@pytest.mark.parametrize("T_total", [40, 41])
def test_time_sharded_pass_world2(T_total):
    res = _run(_worker_lgssm, 2, T_total)
    case = lgssm_case(3, 2, T_total, seed=5)
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    for rank, t0, t1, fm, fL, sm, sL, ell in res:
        assert fm.shape[0] == t1 - t0 + 1
        assert rel_err(fm, ofm[t0:t1 + 1]) < 1e-10 and rel_err(LLt(fL), LLt(ofc[t0:t1 + 1])) < 1e-10
        assert rel_err(sm, osm[t0:t1 + 1]) < 1e-10 and rel_err(LLt(sL), LLt(osc[t0:t1 + 1])) < 1e-10
        assert abs(ell - oell) < 1e-10 * abs(oell)            # all-reduced: whole-sequence value on every rank
    # neighbouring shards overlap by exactly one entry
    np.testing.assert_allclose(res[0][3][-1], res[1][3][0], rtol=1e-12)
    np.testing.assert_allclose(res[0][5][-1], res[1][5][0], rtol=1e-12)


def _worker_iterated(rank, world, port, T_total, q):
    for p in (HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.join(os.path.dirname(HERE), "sqrt-parallel-smoothers_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        _setup(rank, world, port)
        import _np_backend as ops
        import psqrt
        from psqrt import dist as pdist
        from psqrt._lib import LinearizedSSM
        ys, x0, tm, om = _bearings_oracle(T_total)
        t0, t1 = pdist.shard_bounds(T_total, world, rank)
        g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)

        def linearize(lin, _tm, _om, nominal):          # oracle linearisation of the local nominal trajectory
            nom = O.MVNSqrt(nominal.mean.numpy(), nominal.chol.numpy())
            return LinearizedSSM(*[g(a) for a in O.linearize_ssm(O.extended, tm, om, nom)])

        nominal, ell = pdist.iterated_smoothing_sharded(
            g(ys[t0:t1]), psqrt.MVNSqrt(g(x0.mean), g(x0.chol)), None, None, None, None, n_iter=3,
            return_loglikelihood=True, ops=ops, linearize=linearize)
        q.put((rank, t0, t1, nominal.mean.numpy(), nominal.chol.numpy(), float(ell)))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


def _bearings_oracle(T):
    gold = os.path.join(HERE, "golden")
    ys = np.load(os.path.join(gold, "bearings_ys.npy")).astype(np.float64)[:T]
    Q, R, obs, trans = O.bearings_make_parameters(0.01, 0.1, 0.5, 0.01, np.array([-1.5, 0.5]), np.array([1.0, 1.0]))
    x0 = O.MVNSqrt(np.array([-1.0, -1.0, 0.0, 0.0, 0.0]), np.eye(5))
    tm = O.FunctionalModel(trans, O.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    om = O.FunctionalModel(obs, O.MVNSqrt(np.zeros(2), np.linalg.cholesky(R)))
    return ys, x0, tm, om


def test_time_sharded_iterated_world2():
    """nonlinear model: the nominal trajectory stays sharded across iterations (no halo exchange)."""
    T_total = 30
    res = _run(_worker_iterated, 2, T_total)
    ys, x0, tm, om = _bearings_oracle(T_total)
    ores, oell = O.iterated_smoothing(ys, x0, tm, om, O.extended, None, True, criterion=lambda i, *_: i < 3,
                                      return_loglikelihood=True)
    for rank, t0, t1, m, L, ell in res:
        assert rel_err(m, ores.mean[t0:t1 + 1]) < 1e-9 and rel_err(LLt(L), LLt(ores.chol[t0:t1 + 1])) < 1e-9
        assert abs(ell - oell) < 1e-9 * abs(oell)


def test_partition_helpers():
    from psqrt import dist as pdist
    for T, W in ((10, 3), (7, 8), (1_000_000, 8), (5, 1)):
        spans = [pdist.shard_bounds(T, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == T
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    deal = [pdist.batch_indices(100, 8, r) for r in range(8)]
    assert sorted(sum(deal, [])) == list(range(100)) and {len(d) for d in deal} == {12, 13}


def _worker_batch(rank, world, port, n_runs, q):
    for p in (HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.join(os.path.dirname(HERE), "sqrt-parallel-smoothers_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        _setup(rank, world, port)
        import psqrt
        from psqrt import dist as pdist
        ys_all, x0, tm, om = _batch_runs(n_runs)

        def smoother(obs_local, _x0, _tm, _om, _lin, nominal, n_iter, return_loglikelihood):
            # NumPy stand-in for iterated_smoothing_batched: the oracle, run by run
            ms, Ls, ells = [], [], []
            for y in obs_local.numpy():
                res, ell = O.iterated_smoothing(y, x0, tm, om, O.extended, None, True,
                                                criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)
                ms.append(res.mean), Ls.append(res.chol), ells.append(ell)
            out = psqrt.MVNSqrt(torch.as_tensor(np.stack(ms)), torch.as_tensor(np.stack(Ls)))
            return (out, torch.as_tensor(np.array(ells))) if return_loglikelihood else out

        idx, nominal, ell_all = pdist.iterated_smoothing_batch_sharded(
            torch.as_tensor(ys_all), None, None, None, None, None, n_iter=2, return_loglikelihood=True,
            smoother=smoother)
        q.put((rank, idx, None if nominal is None else nominal.mean.numpy(), ell_all.numpy()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, "error", repr(e)))
        raise


def _batch_runs(n_runs, T=12):
    ys, x0, tm, om = _bearings_oracle(T * n_runs)
    return ys.reshape(n_runs, T, 2), x0, tm, om


@pytest.mark.parametrize("n_runs", [5, 1])
def test_batch_sharded_world2(n_runs):
    """BASELINE.json configs[4]: independent runs dealt round-robin, no data-path collective, the per-run
    log-likelihoods gathered in run order on every rank (also when a rank gets no run at all)."""
    res = _run(_worker_batch, 2, n_runs)
    ys_all, x0, tm, om = _batch_runs(n_runs)
    expect = []
    for y in ys_all:
        r, ell = O.iterated_smoothing(y, x0, tm, om, O.extended, None, True, criterion=lambda i, *_: i < 2,
                                      return_loglikelihood=True)
        expect.append((r.mean, ell))
    for rank, idx, means, ell_all in res:
        assert idx == list(range(rank, n_runs, 2))
        np.testing.assert_allclose(ell_all, [e for _, e in expect], rtol=1e-12)
        for j, i in enumerate(idx):
            np.testing.assert_allclose(means[j], expect[i][0], rtol=1e-12, atol=1e-14)
