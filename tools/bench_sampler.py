"""Informational: throughput of the pathwise sampler kernels (psqrt_sample_paths) against the HBM roofline.
Algorithmic bytes per (step, sample): 8 nx in (draws) + 8 nx out (samples).  Run on a GPU box."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sqrt-parallel-smoothers_b200"))
from psqrt import _lib  # noqa: E402

PEAK = 6451.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

dev = torch.device("cuda", 0)
for nx, T, S in ((4, 100_000, 128), (4, 1_000_000, 16), (5, 1000, 100_000), (4, 10, 1_000_000)):
    n_el = T + 1
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    g = torch.randn(n_el, nx, dtype=torch.float64, device=dev, generator=gen)
    E = 0.5 * torch.randn(n_el, nx, nx, dtype=torch.float64, device=dev, generator=gen) / nx
    D = torch.tril(torch.randn(n_el, nx, nx, dtype=torch.float64, device=dev, generator=gen))
    eps = torch.randn(n_el, S, nx, dtype=torch.float64, device=dev, generator=gen)
    for _ in range(3):
        out = _lib.sample_paths(g, E, D, eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        out = _lib.sample_paths(g, E, D, eps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = 16.0 * nx * n_el * S / (ms * 1e-3) / 1e9
    print(json.dumps({"workload": f"pathwise sampler nx={nx} T={T} S={S}", "ms": ms, "sample_steps_per_s": n_el * S / (ms * 1e-3),
                      "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / PEAK, "finite": bool(torch.isfinite(out).all().item())}))
