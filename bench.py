#!/usr/bin/env python
"""bench.py -- smoothed time-steps / second of one sqrt parallel filter + RTS smoother pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl psqrt|reference]

Workload (BASELINE.json metric, SURVEY.md section 8d recipe C1 at the metric's length): random stable
LGSSM, nx = 4, ny = 2, T = 1e6 per GPU, fp64, one `filter_smoother` pass = one step.  With N > 1
(torchrun) the sequence is N x 1e6 steps long and time-sharded: local scans, two exchanges of the
shard totals (P2P stores into peer-mapped buffers, or NCCL all-gathers when peer mapping is unavailable),
carry application (weak scaling).

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM,
`e2e` = the same pass through the public API with pinned HOST buffers (H2D of the observations and
D2H of the smoothed trajectory inside the timed region), `roofline` = the dominant kernel against
the measured HBM peak, `cpu_baseline` = the NumPy restatement of the reference timed on this
box's cores.  `--impl reference` times that CPU restatement alone (JAX is not installable here:
no wheel in /opt/wheelhouse, no network), on the same workload definition.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "sqrt-parallel-smoothers_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

NX, NY = 4, 2
T_PER_GPU = 1_000_000
METRIC = "smoothed time-steps/sec, one sqrt parallel filter+RTS smoother pass, T=1e6 per GPU, nx=4, fp64"


def algorithmic_bytes_per_step(n):
    """read one filtering element + write filtered (m, L) + read one smoothing
    element + write smoothed (m, L) = 8 (7 n^2 + 5 n)."""
    return 8 * (7 * n * n + 5 * n)


def algorithmic_flops_per_step(n):
    """one filtering combine (28.33 n^3 + 22 n^2) + one smoothing combine (7.33 n^3 + 2 n^2)."""
    return 35.67 * n ** 3 + 24 * n ** 2


def measure_fp64_peak(dev):
    """FP64 FMA throughput of this GPU measured live with the library's probe kernel (independent DFMA
    chains, 16 warps per SM): TFLOP/s.  SURVEY.md 8(d): the FP64 peak is not in MEASURED_PEAKS.json."""
    import ctypes
    import torch
    from psqrt import _lib
    lib = _lib.load()
    if not hasattr(lib, "psqrt_fp64_probe"):
        return None
    out = torch.empty(148 * 512, dtype=torch.float64, device=dev)
    iters = 20000
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flops = ctypes.c_double(0.0)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.psqrt_fp64_probe(ctypes.c_void_p(out.data_ptr()), ctypes.c_int(iters), ctypes.byref(flops), st)
        e1.record()
        torch.cuda.synchronize()
        if rc != 0:
            return None
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def make_lgssm(n, ny, seed=0):
    rng = np.random.RandomState(seed)
    Qr, _ = np.linalg.qr(rng.randn(n, n))
    F = 0.99 * Qr
    cholQ = 0.1 * (np.tril(rng.rand(n, n)) + np.eye(n))
    H = rng.randn(ny, n)
    cholR = 0.5 * (np.tril(rng.rand(ny, ny)) + np.eye(ny))
    b = 0.1 * rng.randn(n)
    c = 0.1 * rng.randn(ny)
    m0 = rng.randn(n)
    return dict(F=F, cholQ=cholQ, b=b, H=H, cholR=cholR, c=c, m0=m0, L0=np.eye(n))


def simulate(model, T, seed):
    """fp64 simulation of the LGSSM (procedure of the reference's tests/_lgssm.py:89-93)."""
    rng = np.random.RandomState(seed)
    n, ny = model["F"].shape[0], model["H"].shape[0]
    wq = rng.randn(T, n) @ model["cholQ"].T + model["b"]
    wr = rng.randn(T, ny) @ model["cholR"].T + model["c"]
    F, H = model["F"], model["H"]
    x = model["m0"].copy()
    ys = np.empty((T, ny))
    Ft = np.ascontiguousarray(F.T)
    Ht = np.ascontiguousarray(H.T)
    for k in range(T):
        x = x @ Ft + wq[k]
        ys[k] = x @ Ht + wr[k]
    return ys


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU restatement arm (reference / cpu_baseline)
# --------------------------------------------------------------------------------------------------
def _threaded_tria(n_threads):
    """The oracle's tria with the batch split over host threads (LAPACK releases the GIL)."""
    import parsmooth_np as O
    from concurrent.futures import ThreadPoolExecutor
    base = O.tria
    pool = ThreadPoolExecutor(n_threads)

    def tria(A):
        A = np.asarray(A)
        if A.ndim < 3 or A.shape[0] < 4 * n_threads:
            return base(A)
        parts = np.array_split(np.arange(A.shape[0]), n_threads)
        outs = list(pool.map(lambda idx: base(A[idx[0]:idx[-1] + 1]), [p for p in parts if len(p)]))
        return np.concatenate(outs, 0)

    return base, tria


def cpu_pass_rate(model, T_sample, threads, seed=123):
    """steps/s of the NumPy restatement of the reference's PARALLEL sqrt filter + smoother
    (oracle/parsmooth_np.py: associative_scan over sqrt_filtering_operator / sqrt_smoothing_operator)."""
    import parsmooth_np as O
    ys = simulate(model, T_sample, seed)
    tm = O.FunctionalModel(O.lgssm_function(model["F"]), O.MVNSqrt(model["b"], model["cholQ"]))
    om = O.FunctionalModel(O.lgssm_function(model["H"]), O.MVNSqrt(model["c"], model["cholR"]))
    x0 = O.MVNSqrt(model["m0"], model["L0"])
    base, tt = _threaded_tria(threads)
    O.tria = tt
    try:
        t0 = time.perf_counter()
        O.filter_smoother(ys, x0, tm, om, O.extended, None, True)
        dt = time.perf_counter() - t0
    finally:
        O.tria = base
    return T_sample / dt, dt


def cpu_sequential_rate(model, T_sample, seed=123):
    """steps/s of the NumPy restatement of the reference's SEQUENTIAL sqrt filter + smoother (one core: a step
    loop of _sqrt_predict / _sqrt_update / _sqrt_smooth, sequential/_filtering.py:80-108, _smoothing.py:60-77)."""
    import parsmooth_np as O
    ys = simulate(model, T_sample, seed)
    tm = O.FunctionalModel(O.lgssm_function(model["F"]), O.MVNSqrt(model["b"], model["cholQ"]))
    om = O.FunctionalModel(O.lgssm_function(model["H"]), O.MVNSqrt(model["c"], model["cholR"]))
    x0 = O.MVNSqrt(model["m0"], model["L0"])
    t0 = time.perf_counter()
    O.filter_smoother(ys, x0, tm, om, O.extended, None, False)
    dt = time.perf_counter() - t0
    return T_sample / dt, dt


def _jax_reference():
    """The UNMODIFIED reference (parsmooth on JAX) if this image can import it: BASELINE.md section 3.  Returns a
    function T -> (seconds per jitted filter_smoother call, parallel=True) or None with the reason."""
    try:
        import jax  # noqa: F401
    except Exception as e:      # not installed in this image, no wheel in /opt/wheelhouse, no network
        return None, f"import jax failed: {type(e).__name__}: {e}"
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(cand, "parsmooth")) and cand not in sys.path:
            sys.path.insert(0, cand)
    try:
        import jax
        import jax.numpy as jnp
        jax.config.update("jax_enable_x64", True)
        jax.config.update("jax_platform_name", "cpu")
        from functools import partial
        from parsmooth import FunctionalModel as FM, MVNSqrt as MS
        from parsmooth.linearization import extended as jext
        from parsmooth.methods import filter_smoother as jfs
    except Exception as e:
        return None, f"import parsmooth failed: {type(e).__name__}: {e}"

    def run(model, ys, parallel):
        f = lambda x, A: jnp.dot(A, x)
        tm = FM(partial(f, A=jnp.asarray(model["F"])), MS(jnp.asarray(model["b"]), jnp.asarray(model["cholQ"])))
        om = FM(partial(f, A=jnp.asarray(model["H"])), MS(jnp.asarray(model["c"]), jnp.asarray(model["cholR"])))
        x0 = MS(jnp.asarray(model["m0"]), jnp.asarray(model["L0"]))
        fn = jax.jit(lambda y: jfs(y, x0, tm, om, jext, None, parallel))
        y = jnp.asarray(ys)
        jax.block_until_ready(fn(y))          # compile + warm-up, as the reference's notebooks do
        t0 = time.perf_counter()
        jax.block_until_ready(fn(y))
        return time.perf_counter() - t0
    return run, None


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores, same workload definition
    (C1 LGSSM nx=4 ny=2, one parallel sqrt filter_smoother pass over T=1e6 steps per step).  JAX when importable
    (BASELINE.md section 3); otherwise the NumPy restatement (oracle port), said so in the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    model = make_lgssm(NX, NY)
    T = args.T
    budget_s = float(os.environ.get("PSQRT_REF_BUDGET_S", "150"))
    jax_run, why = _jax_reference()
    times = []
    seq = None
    if jax_run is not None:
        kind, ys = "reference", simulate(model, T, 123)
        dt = jax_run(model, ys, True)
        times.append(dt)
        while len(times) < args.steps and sum(times) + dt < budget_s:
            times.append(jax_run(model, ys, True))
        Ts = min(T, 20_000)
        seq = {"value": Ts / jax_run(model, ys[:Ts], False), "unit": "steps/s", "cores": 1,
               "sample": f"jax.jit(filter_smoother, parallel=False), T={Ts} prefix"}
        sample = f"jax.jit(parsmooth.filter_smoother, parallel=True) on CPU, fp64, T={T} (the full workload)"
    else:
        kind = "port"
        cpu_pass_rate(model, 10_000, threads)                 # warm-up (thread pool, LAPACK)
        _, dt = cpu_pass_rate(model, T, threads)
        times.append(dt)
        while len(times) < args.steps and sum(times) + dt < budget_s:
            times.append(cpu_pass_rate(model, T, threads)[1])
        Ts = 20_000
        r_seq, dt_seq = cpu_sequential_rate(model, Ts)
        seq = {"value": r_seq, "unit": "steps/s", "cores": 1,
               "sample": f"NumPy restatement of the SEQUENTIAL sqrt filter+smoother, T={Ts} sample ({dt_seq:.1f} s)"}
        sample = (f"NumPy restatement of parsmooth's parallel sqrt filter+smoother (associative_scan, batched LAPACK QR "
                  f"split over {threads} threads), T={T} (the full workload), {len(times)} timed pass(es) of "
                  f"{args.steps} requested (bounded to ~{budget_s:.0f} s)")
    value = T * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C1 LGSSM nx={NX} ny={NY}, T={T:.0e}, one parallel sqrt filter_smoother pass per step "
                               "(CPU arm: the same workload on the host cores)",
                   "nx": NX, "ny": NY, "T_per_gpu": T, "same_config": True},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": threads, "kind": kind, "sample": sample,
                         "sequential": seq},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if jax_run is None:
        line["note"] = (f"unmodified reference not runnable here ({why}; no jax wheel in /opt/wheelhouse, no "
                        "network): this arm is the oracle port (see DESIGN.md)")
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# psqrt arm
# --------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """Best effort: run this rank (and therefore first-touch its pinned host buffers) on the NUMA node the GPU's
    PCIe root hangs off, so that N ranks do not all stage their copies through node 0."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def sharded_parity(world, rank, dev, exchange):
    """Correctness of the time-sharded pass, checked in the same run that times it (N > 1): one sequence of
    world x 5e4 steps smoothed (a) time-sharded over the ranks, three passes in a row (slot reuse / epochs of the peer
    exchange), smoother + log-likelihood and a filter-only pass, and (b) unsharded on every rank's own GPU;
    rank 0 also checks the filtered prefix against the oracle (CPU restatement of the reference, used as the
    checker only).  Returns the dict printed under "parity"."""
    import torch
    import torch.distributed as dist
    from psqrt import _lib, dist as pdist
    from psqrt._lib import LinearizedSSM
    Tl = 50_000
    model = make_lgssm(NX, NY)
    ys_full = simulate(model, Tl * world, seed=7)
    g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    names = ("F", "cholQ", "b", "H", "cholR", "c")
    ssm = LinearizedSSM(*[g(model[k]) for k in names], host={k: model[k] for k in names})
    m0, L0 = g(model["m0"]), g(model["L0"])
    fm_ref, fL_ref, sm_ref, sL_ref, ell_ref = _lib.filter_smoother(ssm, g(ys_full), m0, L0, smooth=True, loglik=True)
    y_loc = g(ys_full[rank * Tl:(rank + 1) * Tl])[None].contiguous()
    sh = pdist.TimeShardedSmoother(NX, NY, Tl, device=dev, exchange=exchange)
    for _ in range(3):
        fm, fL, sm, sL, ell = sh.filter_smoother(ssm, y_loc, m0[None], L0[None], smooth=True, loglik=True)
    for _ in range(2):
        ffm, ffL, _, _, _ = sh.filter_smoother(ssm, y_loc, m0[None], L0[None], smooth=False, loglik=False)
    torch.cuda.synchronize()
    sl = slice(rank * Tl, (rank + 1) * Tl + 1)
    LLt = lambda L: L @ L.transpose(-1, -2)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    errs = [rel(sm[0], sm_ref[sl]), rel(LLt(sL[0]), LLt(sL_ref[sl])), float((ell[0] - ell_ref).abs() / ell_ref.abs()),
            max(rel(ffm[0][1:], fm_ref[sl][1:]), rel(fm[0][1:], fm_ref[sl][1:])),
            max(rel(LLt(ffL[0][1:]), LLt(fL_ref[sl][1:])), rel(LLt(fL[0][1:]), LLt(fL_ref[sl][1:])))]
    t = torch.tensor(errs, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"T_local": Tl, "vs": "unsharded pass on one GPU (every rank) + oracle filtered prefix (rank 0)",
           "max_rel_mean": float(t[0]), "max_rel_LLt": float(t[1]), "rel_ell": float(t[2]),
           "max_rel_filtered_mean": float(t[3]), "max_rel_filtered_LLt": float(t[4]), "exchange": sh.exchange}
    ok = bool(t[0] < 1e-9 and t[1] < 1e-9 and t[2] < 1e-8 and t[3] < 1e-9 and t[4] < 1e-9)
    if rank == 0:
        import parsmooth_np as O
        n_or = 4000
        tm = O.FunctionalModel(O.lgssm_function(model["F"]), O.MVNSqrt(model["b"], model["cholQ"]))
        om = O.FunctionalModel(O.lgssm_function(model["H"]), O.MVNSqrt(model["c"], model["cholR"]))
        of = O.filtering(ys_full[:n_or], O.MVNSqrt(model["m0"], model["L0"]), tm, om, O.extended, None, True)
        gm, gL = fm[0][1:n_or + 1].cpu().numpy(), fL[0][1:n_or + 1].cpu().numpy()
        om_, oL_ = np.asarray(of.mean)[1:], np.asarray(of.chol)[1:]
        e1 = float(np.abs(gm - om_).max() / np.abs(om_).max())
        oP = oL_ @ np.swapaxes(oL_, -1, -2)
        e2 = float(np.abs(gL @ np.swapaxes(gL, -1, -2) - oP).max() / np.abs(oP).max())
        out["oracle_prefix"] = {"steps": n_or, "max_rel_mean": e1, "max_rel_LLt": e2}
        ok = ok and e1 < 1e-9 and e2 < 1e-9
    flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item() > 0.5)
    return out


def c4_strong(world, rank, dev, steps=5):
    """BASELINE.json configs[3]: LGSSM nx=8 ny=4, T = 1e7 steps IN TOTAL, time-sharded over the ranks (strong
    scaling); one filter_smoother pass per step, replayed from a CUDA graph.  Observations are iid N(0,1) generated on
    the device (no kernel branches on data: timing is data-independent)."""
    import torch
    import torch.distributed as dist
    from psqrt import _lib, dist as pdist
    from psqrt._lib import LinearizedSSM
    nx, ny, T_total = 8, 4, 10_000_000
    if not _lib.supported(nx, ny):
        return {"unavailable": "nx=8 not compiled in"}
    Tl = T_total // world
    model = make_lgssm(nx, ny, seed=1)
    g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    names = ("F", "cholQ", "b", "H", "cholR", "c")
    ssm = LinearizedSSM(*[g(model[k]) for k in names], host={k: model[k] for k in names})
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    y = torch.randn((1, Tl, ny), dtype=torch.float64, device=dev, generator=gen)
    m0, L0 = g(model["m0"])[None].contiguous(), g(model["L0"])[None].contiguous()
    if world > 1:
        sh = pdist.TimeShardedSmoother(nx, ny, Tl, device=dev, exchange=os.environ.get("PSQRT_EXCHANGE", "peer"))
        eager = lambda: sh.filter_smoother(ssm, y, m0, L0)
    else:
        eager = lambda: _lib.filter_smoother(ssm, y, m0, L0, smooth=True, loglik=False)
    one, launch = eager, "eager launches"
    try:
        for _ in range(2):
            eager()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            if sh.exchange != "peer":
                raise RuntimeError("exchange is NCCL")
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            eager()
        torch.cuda.synchronize()
        one, launch = graph.replay, "one CUDA graph per pass"
    except Exception as e:
        launch = f"eager launches ({e})"
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= steps
    hbm = algorithmic_bytes_per_step(nx)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    value = T_total / (ms * 1e-3)
    del y
    torch.cuda.empty_cache()
    return {"workload": f"LGSSM nx={nx} ny={ny}, T_total={T_total:.0e} time-sharded over {world} GPU(s) "
                        f"({Tl} steps each), one filter_smoother pass", "scaling": "strong", "ms_per_step": ms,
            "value": value, "unit": "steps/s", "steps": steps, "launch": launch,
            "frac_of_hbm_bound_per_gpu": value / world / (peak * 1e9 / hbm),
            "note": "strong-scaling efficiency = value(N) / (N x value(1)); computed by the reader from the N=1 line"}


def run_psqrt(args):
    import torch
    import torch.distributed as dist
    import psqrt
    from psqrt import _lib
    from psqrt._lib import LinearizedSSM
    from psqrt.models import lgssm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the psqrt arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity0 = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa(local_rank)   # pinned buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    # ---- N > 1: the sharded pass is CHECKED before it is timed ---------------------------------------
    parity = None
    if world > 1:
        parity = sharded_parity(world, rank, dev, os.environ.get("PSQRT_EXCHANGE", "peer"))
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "error": "sharded pass failed its parity check",
                                  "parity": parity}))
            dist.barrier()
            dist.destroy_process_group()
            return 1

    T = args.T
    model = make_lgssm(NX, NY)
    ys_np = simulate(model, T, seed=1000 + rank)
    g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    ssm = LinearizedSSM(*[g(model[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")],
                        host=None if args.no_host_model else {k: model[k] for k in ("F", "cholQ", "b", "H", "cholR", "c")})
    ys = g(ys_np)
    m0, L0 = g(model["m0"]), g(model["L0"])
    x0 = psqrt.MVNSqrt(m0, L0)
    tm = psqrt.FunctionalModel(lgssm.transition_function(model["F"]), psqrt.MVNSqrt(g(model["b"]), g(model["cholQ"])))
    om = psqrt.FunctionalModel(lgssm.observation_function(model["H"]), psqrt.MVNSqrt(g(model["c"]), g(model["cholR"])))

    if world > 1:
        from psqrt import dist as pdist
        sharded = pdist.TimeShardedSmoother(NX, NY, T, device=dev, exchange=os.environ.get("PSQRT_EXCHANGE", "peer"))

        yb_, m0b_, L0b_ = ys[None].contiguous(), m0[None].contiguous(), L0[None].contiguous()

        def eager_pass():
            return sharded.filter_smoother(ssm, yb_, m0b_, L0b_)

        one_pass = eager_pass
        graph_note = "eager launches"
        if os.environ.get("PSQRT_GRAPH", "1") != "0":
            # the host side of a sharded pass (11 launches + tensor bookkeeping through ctypes) costs more than the
            # GPU work it feeds: capture the pass once, replay it per step.  Exchanges are kernels (peer mode) or
            # NCCL collectives (capturable), epochs live on the device.
            try:
                for _ in range(3):
                    eager_pass()
                torch.cuda.synchronize()
                dist.barrier()
                if sharded.exchange != "peer":     # NCCL collectives inside a capture hung on this pool: stay eager
                    raise RuntimeError("exchange is NCCL")
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    graph_out = eager_pass()
                torch.cuda.synchronize()

                def one_pass():
                    graph.replay()
                    return graph_out
                graph_note = "one CUDA graph per pass"
            except Exception as e:       # capture not possible on this system: eager launches
                graph_note = f"eager launches ({e})"
                one_pass = eager_pass
    else:
        def eager_pass():
            return _lib.filter_smoother(ssm, ys, m0, L0, smooth=True, loglik=False, chunk_len=args.chunk)

        one_pass = eager_pass
        graph_note = "eager launches"
        if os.environ.get("PSQRT_GRAPH", "1") != "0":
            # the five kernels of a pass replayed from one CUDA graph: no launch gaps, no per-step host work
            try:
                for _ in range(3):
                    eager_pass()
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    graph_out = eager_pass()
                torch.cuda.synchronize()

                def one_pass():
                    graph.replay()
                    return graph_out
                graph_note = "one CUDA graph per pass"
            except Exception as e:       # capture not possible on this system: eager launches
                graph_note = f"eager launches ({e})"
                one_pass = eager_pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        one_pass()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_pass()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = T * world / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers ------------------------------------
    # Host buffers are pinned; the D2H of pass i runs on a copy stream into one of two result buffers while the H2D
    # and the kernels of pass i + 1 proceed on the compute stream (PCIe is full duplex).  Every pass's inputs and
    # results cross the bus inside the timed region, which ends when the last result has landed on the host.
    ys_host = torch.as_tensor(ys_np).pin_memory()
    out_m = [torch.empty((T + 1, NX), dtype=torch.float64).pin_memory() for _ in range(2)]
    out_L = [torch.empty((T + 1, NX, NX), dtype=torch.float64).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    e2e_count = [0]

    def e2e_pass():
        if world > 1:
            # the sharded pass replays from its CUDA graph on the static input buffer: H2D into that buffer, replay,
            # D2H of the smoothed shard (static result buffers: the copy must finish before the next replay)
            ys.copy_(ys_host, non_blocking=True)
            fm, fL, sm, sL, _ = one_pass()
            out_m[0].copy_(sm.reshape(out_m[0].shape), non_blocking=True)
            out_L[0].copy_(sL.reshape(out_L[0].shape), non_blocking=True)
            return
        y_dev = ys_host.to(dev, non_blocking=True)
        res = psqrt.filter_smoother(y_dev, x0, tm, om, psqrt.linearization.extended, None, True)
        sm, sL = res.mean, res.chol
        done = torch.cuda.Event()
        done.record()
        i = e2e_count[0] & 1
        e2e_count[0] += 1
        copy_stream.wait_event(done)
        with torch.cuda.stream(copy_stream):
            out_m[i].copy_(sm.reshape(out_m[i].shape), non_blocking=True)
            out_L[i].copy_(sL.reshape(out_L[i].shape), non_blocking=True)
        sm.record_stream(copy_stream)
        sL.record_stream(copy_stream)

    for _ in range(2):
        e2e_pass()
    copy_stream.synchronize()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(3, min(args.steps, 10))
    e2.record()
    for _ in range(n_e2e):
        e2e_pass()
    torch.cuda.current_stream().wait_stream(copy_stream)   # the timed region ends after the last D2H
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    ms_e2e /= n_e2e
    clocks = sampler.stop() if rank == 0 else None
    e2e = {"value": T * world / (ms_e2e * 1e-3), "unit": "steps/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": int(ys_host.numel() * 8) * world,
           "d2h_bytes_per_step": int((out_m[0].numel() + out_L[0].numel()) * 8) * world,
           "overlap": ("D2H of pass i on a copy stream (two pinned result buffers) under the H2D + kernels of pass i + 1"
                       if world == 1 else "none (static graph buffers)")}

    # ---- per-kernel-stage timing for the roofline (staged C-ABI calls, same kernels) -------------
    roofline = None
    stages = None
    if rank == 0:
        yb, m0b, L0b = ys[None].contiguous(), m0[None].contiguous(), L0[None].contiguous()
        names = ("filter_reduce(K1+K2)", "filter_apply+smooth_reduce(K3+K4)", "smooth_apply(K5)")
        acc = {n: [] for n in names}
        evs = []
        for it in range(3 + 10):   # back to back, no sync in between: launch latency stays hidden behind the GPU work
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            _lib.filter_reduce(ssm, yb, NX, chunk_len=args.chunk)
            ev[1].record()
            fm, fL, _, stot = _lib.filter_apply(ssm, yb, m0b, L0b, smooth=True, loglik=False, chunk_len=args.chunk)
            ev[2].record()
            _lib.smoother_apply(ssm, fm, fL, fm[:, -1].contiguous(), fL[:, -1].contiguous(), chunk_len=args.chunk)
            ev[3].record()
            evs.append(ev)
        torch.cuda.synchronize()
        for ev in evs[3:]:
            for i, n in enumerate(names):
                acc[n].append(ev[i].elapsed_time(ev[i + 1]))
        stages = {n: float(np.median(v)) for n, v in acc.items()}
        # canonical per-step bytes attributed to each stage
        share = {names[0]: 8 * (3 * NX * NX + 2 * NX), names[1]: 8 * (NX * NX + NX),
                 names[2]: 8 * (2 * NX * NX + NX) + 8 * (NX * NX + NX)}
        dom = max(stages, key=stages.get)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = share[dom] * T / (stages[dom] * 1e-3) / 1e9
        # whole pass, per GPU: B(n) bytes per step over the time of the complete pass
        pass_achieved = algorithmic_bytes_per_step(NX) * (T / (ms_per_step * 1e-3)) / 1e9
        # DRAM bytes of the dominant stage per launch, from the committed `ncu --set full` capture of this exact
        # configuration (profiles/r02_dram_traffic.json, keyed by nx/ny/T); null for any other configuration
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"nx{NX}_ny{NY}_T{T}", {}).get(dom)
        fp64_peak = measure_fp64_peak(dev)
        fp64_src = "measured live (psqrt_fp64_probe: independent DFMA chains, 16 warps/SM)"
        if fp64_peak is None:
            fp64_peak, fp64_src = 37.2, "nominal (148 SMs x 64 FMA/clk x 2 x 1.965 GHz)"
        steps_per_s = T / (ms_per_step * 1e-3)
        fp64_achieved = algorithmic_flops_per_step(NX) * steps_per_s / 1e12
        hbm_bound = peak * 1e9 / algorithmic_bytes_per_step(NX)
        fp64_bound = fp64_peak * 1e12 / algorithmic_flops_per_step(NX)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_step": share[dom], "stage_ms": stages,
                    "whole_pass": {"algorithmic_bytes_per_step": algorithmic_bytes_per_step(NX),
                                   "achieved": pass_achieved, "frac": pass_achieved / peak,
                                   "note": "per GPU; B(n) = 8(7n^2+5n) canonical bytes per step / time of the whole pass"},
                    "fp64": {"algorithmic_flops_per_step": algorithmic_flops_per_step(NX), "achieved": fp64_achieved,
                             "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_achieved / fp64_peak,
                             "peak_source": fp64_src},
                    "north_star": {"steps_per_s": steps_per_s, "hbm_bound_steps_per_s": hbm_bound,
                                   "fp64_bound_steps_per_s": fp64_bound,
                                   "frac_of_slower_bound": steps_per_s / min(hbm_bound, fp64_bound),
                                   "note": "BASELINE.json: fraction of min(HBM bound on canonical element bytes, FP64 "
                                           "bound on canonical combine flops)"}}

    # ---- CPU baseline beside it (rank 0, N = 1 only, bounded sample) -----------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, affinity0)    # the CPU baseline may use every host core again
        threads = os.cpu_count() or 1
        cpu_pass_rate(model, 5_000, threads)
        Tc = T if T <= 1_000_000 else 1_000_000
        r, dt = cpu_pass_rate(model, Tc, threads)
        r_seq, dt_seq = cpu_sequential_rate(model, 20_000)
        cpu_baseline = {"value": r, "unit": "steps/s", "cores": threads, "kind": "port",
                        "sample": f"NumPy restatement of the reference's PARALLEL sqrt filter+smoother, one pass over "
                                  f"T={Tc} steps of the same LGSSM ({dt:.1f} s), LAPACK QR batch split over {threads} "
                                  f"threads (JAX is not installable in this image)",
                        "sequential": {"value": r_seq, "unit": "steps/s", "cores": 1,
                                       "sample": f"NumPy restatement of the SEQUENTIAL sqrt filter+smoother, T=20000 "
                                                 f"sample ({dt_seq:.1f} s)"}}

    # ---- secondary: BASELINE.json configs[3], strong scaling (nx=8, ny=4, T_total=1e7) ------------
    secondary = {}
    if (NX, NY) == (4, 2) and T == T_PER_GPU and not args.no_secondary:
        try:
            res = c4_strong(world, rank, dev)
        except Exception as e:      # informational line: never fails the metric
            res = {"unavailable": repr(e)}
        secondary["c4_strong"] = res

    if rank == 0:
        plan = _lib.get_plan(NX, NY, T, 1, args.chunk, ssm=ssm)
        line = {
            "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"random stable LGSSM (SURVEY 8d recipe C1) nx={NX} ny={NY}, T={T:.0e} steps per GPU, "
                                   "one sqrt parallel filter_smoother pass per step"
                                   + (", time-sharded over the GPUs with two exchanges of the shard totals (see "
                                      "`exchange`)" if world > 1 else ""),
                       "nx": NX, "ny": NY, "T_per_gpu": T, "T_total": T * world, "chunk_len": plan.chunk_len,
                       "parallelism": f"time-shard x{world}" if world > 1 else "single GPU",
                       "launch": graph_note,
                       "exchange": (None if world == 1 else
                                    ("P2P stores into peer-mapped buffers, fused into the mid-scan / carry kernels (psqrt_peer, include/psqrt.h)"
                                     if sharded.exchange == "peer" else
                                     "NCCL all-gather x2" + (f" (peer exchange unavailable: {sharded.exchange_error})"
                                                             if sharded.exchange_error else ""))),
                       "l2": (f"working set per pass per GPU (y {8e-6 * NY * T:.0f} MB + filtered "
                              f"{8e-6 * (NX + NX * NX) * T:.0f} MB + smoothed {8e-6 * (NX + NX * NX) * T:.0f} MB) "
                              + ("exceeds" if 8e-6 * (NY + 2 * (NX + NX * NX)) * T > 126 else "does NOT exceed")
                              + " the 126 MB L2; no explicit flush")},
            "e2e": e2e,
            # K1, K2, K3 (the smoothing mid scan K4 runs inside it on spare CTAs unless PSQRT_FUSE_MID=0), K5;
            # time-sharded with the peer exchange: + K4 and the two carry scans
            # kernels of one pass: K1, K2, K3 (+ K4 inside unless PSQRT_FUSE_MID=0; the nx = 8 sub-warp form has its own
            # launch list), K5; a time-sharded pass adds the two carry kernels
            "gpu_launches": ((4 if os.environ.get("PSQRT_FUSE_MID", "1") != "0" else 5) + (0 if world == 1 else 2))
                            * args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if parity is not None:
            line["parity"] = parity
        if numa is not None:
            line["config"]["host_numa_node_rank0"] = numa   # the rank ran (and pinned its buffers) on this node
        if secondary:
            line["secondary"] = secondary
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_bearings(args):
    """Secondary workloads.  BASELINE.json configs[1]: bearings-only coordinated-turn tracking, nx = 5, 2 sensors,
    T = 1e5, iterated sqrt extended (or cubature / Gauss-Hermite) parallel smoother, 10 iterations, through the public
    API (psqrt.methods.iterated_smoothing).  configs[4] (--runs 100 --T 10000 --iters 20 --batched): independent
    Monte-Carlo runs, dealt round-robin over the ranks under torchrun (psqrt.dist.iterated_smoothing_batch_sharded:
    no data-path collective, the log-likelihoods gathered at the end).  One informational JSON line; not the driver's
    metric."""
    import torch
    import torch.distributed as dist
    import psqrt
    from psqrt.models import bearings
    from psqrt import dist as pdist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = args.T if args.T != T_PER_GPU else 100_000
    s1, s2, r, dt, qc, qw = np.array([-1.5, 0.5]), np.array([1.0, 1.0]), 0.5, 0.01, 0.01, 0.1
    runs = max(args.runs, 1)      # > 1: BASELINE.json configs[4], independent Monte-Carlo runs
    mine = pdist.batch_indices(runs, world, rank)
    ys_all = {k: bearings.get_data(np.array([0.1, 0.2, 1.0, 0.0]), dt, r, T, s1, s2, random_state=k)[2] for k in mine}
    Q, R, obs_f, trans_f = bearings.make_parameters(qc, qw, r, dt, s1, s2)
    g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    x0 = psqrt.MVNSqrt(np.array([-4.0, -1.0, 2.0, 7.0, 3.0]), np.eye(5))
    tm = psqrt.FunctionalModel(trans_f, psqrt.MVNSqrt(np.zeros(5), np.linalg.cholesky(Q)))
    om = psqrt.FunctionalModel(obs_f, psqrt.MVNSqrt(np.zeros(2), np.linalg.cholesky(R)))
    nominal = psqrt.MVNSqrt(g(np.tile(np.array([-1.0, -1.0, 6.0, 4.0, 2.0]), (T + 1, 1))),
                            torch.eye(5, dtype=torch.float64, device=dev).expand(T + 1, 5, 5))
    ys_ds = [g(ys_all[k].astype(np.float64)) for k in mine]
    lin = getattr(psqrt.linearization, args.lin)
    n_iter = args.iters
    ys_batch = torch.stack(ys_ds) if (args.batched and ys_ds) else None

    class _Runs:          # per-run observations, indexable by run number (only this rank's runs are resident)
        def __len__(self):
            return runs

        def __getitem__(self, i):
            return ys_ds[mine.index(i)]

    if args.grad:
        # BASELINE.json configs[2]: log-likelihood + gradient path.  Protocol of
        # notebooks/experiment_bearing_only_param_estimation_run_time.ipynb: parameter prec_r = 1 / r of the first
        # sensor, R = diag(r^2, 0.1^2), value_and_grad of -ell through n_iter iterations of the iterated smoother
        # (psqrt.grad.loglikelihood_jvp: n_iter primal passes + n_iter + 3 tangent passes on the device).
        from psqrt import grad as pgrad
        # the notebook's model (sensor noises, prior shape) on a track that stays observable for 1e5 steps: sensors at
        # (-5, 0.5), (5, 1), slowly drifting turn rate (q = 0.05), initial nominal = simulated states + N(0, 0.1^2).
        # (The notebook's own set-up -- sensors at (-1, 0.5), (1, 1), q = 10, nominal from the inverted bearings -- is
        # used up to T = 4096 there; at T = 1e5 its undamped iterated smoother does not converge)
        s1g, s2g, qcg, qwg, r_true = np.array([-5.0, 0.5]), np.array([5.0, 1.0]), 0.1, 0.1, 0.05
        _, xs_g, ys_g = bearings.get_data_pe(np.array([0.1, 0.2, 1.0, 0.0]), dt, r_true, T, s1g, s2g, q=0.05,
                                             random_state=0)
        ys_g = ys_g.astype(np.float64)
        Qg, _, obs_g, trans_g = bearings.make_parameters(qcg, qwg, r_true, dt, s1g, s2g, r2=0.1)
        tm = psqrt.FunctionalModel(trans_g, psqrt.MVNSqrt(np.zeros(5), np.linalg.cholesky(Qg)))
        x0 = psqrt.MVNSqrt(np.array([0.1, 0.2, 1.0, 0.0, 1.0]), np.diag([0.5, 0.5, 0.5, 0.5, 1.0]))
        nom_m = xs_g.astype(np.float64) + 0.1 * np.random.RandomState(1).randn(T + 1, 5)
        nominal = psqrt.MVNSqrt(g(nom_m), (np.sqrt(0.1) * torch.eye(5, dtype=torch.float64, device=dev)).expand(T + 1, 5, 5))
        prec = args.prec
        om_of = lambda p: psqrt.FunctionalModel(obs_g, psqrt.MVNSqrt(np.zeros(2), np.diag([1.0 / p, 0.1])))
        om_pe = om_of(prec)
        tg = pgrad.Tangents(observation_noise=psqrt.MVNSqrt(None, np.diag([-1.0 / prec ** 2, 0.0])))
        ys_d = g(ys_g)

        def run_grad():
            return pgrad.loglikelihood_jvp(ys_d, x0, tm, om_pe, lin, tg, nominal, True,
                                           criterion=lambda i, *_: i < n_iter)

        for _ in range(max(args.warmup, 2)):
            run_grad()
        torch.cuda.synchronize()
        def median_ms(fn):   # one event pair per call, median (see the note at the primal timing below)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            out = None
            for a, b in evs:
                a.record()
                out = fn()
                b.record()
            torch.cuda.synchronize()
            t = sorted(a.elapsed_time(b) for a, b in evs)
            return (t[len(t) // 2] if len(t) % 2 else 0.5 * (t[len(t) // 2 - 1] + t[len(t) // 2])), out

        ms, (nom, ell, dell) = median_ms(run_grad)
        # the primal alone, for the cost of the gradient relative to the value
        ms_val, _ = median_ms(lambda: psqrt.iterated_smoothing(ys_d, x0, tm, om_pe, lin, nominal, True,
                                                               criterion=lambda i, *_: i < n_iter,
                                                               return_loglikelihood=True))
        hfd = 1e-4 * prec   # check of the device gradient: central difference of the primal path itself
        ells = [float(psqrt.iterated_smoothing(ys_d, x0, tm, om_of(prec + sgn * hfd), lin, nominal, True,
                                               criterion=lambda i, *_: i < n_iter, return_loglikelihood=True)[1])
                for sgn in (1.0, -1.0)]
        fd = (ells[0] - ells[1]) / (2 * hfd)
        print(json.dumps({"workload": f"bearings-only CT nx=5 ny=2 T={T}, {args.lin} linearization, log-likelihood of the "
                                      f"iterated sqrt parallel smoother ({n_iter} iterations) + d ell / d prec_r "
                                      f"(BASELINE.json configs[2]; implicit fixed-point tangent, {n_iter + 2} Neumann terms)",
                          "n_gpus": 1, "ms_per_value_and_grad": ms, "ms_per_value": ms_val,
                          "tangent_passes": n_iter + 3, "ell": float(ell), "dell_dprec": float(dell),
                          "dell_dprec_central_difference": fd, "prec_r": prec,
                          "value": T * (2 * n_iter + 3) / (ms * 1e-3), "unit": "step-passes/s (primal + tangent)",
                          "finite": bool(np.isfinite(float(ell)) and np.isfinite(float(dell)))}))
        return 0

    def run():
        if args.batched and runs > 1:
            # config 5 through its driver: runs dealt round-robin, this rank's share smoothed as ONE batch per
            # iteration, log-likelihoods (robustness_100runs.py:60-66) gathered at the end
            _, nom, _ = pdist.iterated_smoothing_batch_sharded(_Runs(), x0, tm, om, lin, nominal, n_iter=n_iter,
                                                               return_loglikelihood=True)
            return nom
        if args.batched:
            return pdist.iterated_smoothing_batched(ys_batch, x0, tm, om, lin, nominal, n_iter=n_iter)
        res = None
        for ys_d in ys_ds:
            res = psqrt.iterated_smoothing(ys_d, x0, tm, om, lin, nominal, True, criterion=lambda i, *_: i < n_iter)
        return res

    for _ in range(max(args.warmup, 2)):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # one event pair per call, median reported: the host loop of an iterated smoother (~50 launches per iteration) is
    # exposed to scheduling hiccups of a busy host, and one 40 ms stall would otherwise set the mean of a few calls
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        a.record()
        res = run()
        b.record()
    torch.cuda.synchronize()
    per_call = sorted(a.elapsed_time(b) for a, b in evs)
    ms = per_call[len(per_call) // 2] if len(per_call) % 2 else 0.5 * (per_call[len(per_call) // 2 - 1] +
                                                                      per_call[len(per_call) // 2])
    ms_mean, ms_min = sum(per_call) / len(per_call), per_call[0]
    finite = True if res is None else bool(torch.isfinite(res.mean).all().item())
    if world > 1:
        t = torch.tensor([ms, 0.0 if finite else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, finite = float(t[0]), bool(t[1] == 0)
    if rank == 0:
        print(json.dumps({"workload": f"bearings-only CT nx=5 ny=2 T={T}, iterated sqrt {args.lin} parallel smoother, "
                                      f"{n_iter} iterations" + (f", {runs} independent runs "
                                                                + ("as one batch per GPU" if args.batched else "in sequence")
                                                                + " (BASELINE.json configs[4])" if runs > 1
                                                                else " (BASELINE.json configs[1])"),
                          "n_gpus": world, "parallelism": f"batch-shard x{world} (round-robin, no collective)" if world > 1 else "single GPU",
                          "ms_per_call": ms, "ms_per_call_mean": ms_mean, "ms_per_call_min": ms_min,
                          "timing": f"median of {args.steps} calls, CUDA events around each call",
                          "value": runs * T * n_iter / (ms * 1e-3), "unit": "step-passes/s", "finite": finite}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="psqrt", choices=["psqrt", "reference"])
    ap.add_argument("--T", type=int, default=T_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the informational nx=8 T=1e7 strong-scaling line")
    ap.add_argument("--no-host-model", action="store_true", help="load the model from HBM per step instead of by value")
    ap.add_argument("--chunk", type=int, default=0, help="chunk length override (0 = library default)")
    ap.add_argument("--workload", default="lgssm", choices=["lgssm", "bearings"],
                    help="lgssm = the driver's metric; bearings = informational configs[1] line")
    ap.add_argument("--lin", default="extended", choices=["extended", "cubature", "gauss_hermite", "unscented"])
    ap.add_argument("--runs", type=int, default=1, help="bearings workload: independent data sets smoothed in sequence")
    ap.add_argument("--batched", action="store_true", help="bearings workload: smooth the runs as one batch")
    ap.add_argument("--grad", action="store_true", help="bearings workload: log-likelihood + gradient (configs[2])")
    ap.add_argument("--prec", type=float, default=10.0, help="--grad: the parameter prec_r = 1 / r of the first sensor")
    ap.add_argument("--iters", type=int, default=10, help="bearings workload: iterations of the iterated smoother")
    ap.add_argument("--nx", type=int, default=4, help="state dimension of the LGSSM workload (informational runs; "
                                                      "the driver's metric is the default nx=4, ny=2)")
    ap.add_argument("--ny", type=int, default=2)
    args = ap.parse_args()
    global NX, NY, METRIC
    if (args.nx, args.ny) != (NX, NY):
        NX, NY = args.nx, args.ny
        METRIC = METRIC.replace("nx=4", f"nx={NX}")
    if args.workload == "bearings":
        return run_bearings(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_psqrt(args)


if __name__ == "__main__":
    sys.exit(main())
