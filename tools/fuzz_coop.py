"""Randomised parity sweep of the nx = 8 sub-warp path (and the nx <= 6 per-thread path with the current chunk plans)
against the oracle: random ny, T, chunk length, batch, time-varying or not, smoother on / off.
    python tools/fuzz_coop.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("sqrt-parallel-smoothers_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

from _cases import LLt, lgssm_case, oracle_from_ssm, rel_err, time_varying_case   # noqa: E402
from psqrt import _lib   # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda", 0)
g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
worst = 0.0
for it in range(n_cases):
    n = int(rng.choice([8, 8, 8, 5, 4, 6]))
    ny = int(rng.randint(1, 5))
    T = int(rng.choice([1, 2, 7, 31, 32, 33, 100, 257, 1000, 2049, 4100]))
    K = int(rng.choice([0, 0, 1, 2, 3, 5, 9, 33]))
    tv = bool(rng.randint(2))
    smooth = bool(rng.randint(4))
    case = (time_varying_case if tv else lgssm_case)(n, ny, T, seed=int(rng.randint(1 << 30)))
    host = None if tv or rng.randint(2) else {k: case[k] for k in ("F", "cholQ", "b", "H", "cholR", "c")}
    ssm = _lib.LinearizedSSM(*[g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")], host=host)
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, g(case["ys"]), g(case["m0"]), g(case["L0"]), smooth=smooth, loglik=True,
                                               chunk_len=K)
    torch.cuda.synchronize()
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    errs = [rel_err(fm.cpu().numpy(), ofm), rel_err(LLt(fL.cpu().numpy()), LLt(ofc)), abs(ell.item() - oell) / abs(oell)]
    if smooth:
        errs += [rel_err(sm.cpu().numpy(), osm), rel_err(LLt(sL.cpu().numpy()), LLt(osc))]
    e = max(x if np.isfinite(x) else 1e9 for x in errs)
    worst = max(worst, e)
    flag = "" if e < 1e-9 else "   <-- FAIL"
    print(f"n={n} ny={ny} T={T} K={K} tv={tv} smooth={smooth} host={host is not None}: {e:.1e}{flag}", flush=True)
print("WORST", worst)
