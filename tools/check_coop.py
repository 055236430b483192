"""Development check of the sub-warp sweeps (csrc/psqrt_coopsweep.cuh): one filter + smoother pass per case against the
oracle, printing the error of every output.  Run once per PSQRT_COOP mask (read once per process):
    for m in 0 2 4 6 1 7; do PSQRT_COOP=$m python tools/check_coop.py; done"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("sqrt-parallel-smoothers_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))

from _cases import LLt, lgssm_case, oracle_from_ssm, rel_err, time_varying_case   # noqa: E402
from psqrt import _lib   # noqa: E402

dev = torch.device("cuda", 0)
g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
cases = [("lgssm", 8, 4, 130, 0), ("lgssm", 8, 4, 1100, 1), ("lgssm", 8, 2, 700, 5), ("tv", 8, 4, 500, 3),
         ("lgssm", 6, 4, 257, 3), ("tv", 6, 3, 300, 4), ("lgssm", 8, 4, 33000, 1), ("lgssm", 8, 1, 64, 2),
         ("lgssm", 8, 4, 20000, 0)]
worst = 0.0
for kind, n, ny, T, K in cases:
    case = (lgssm_case if kind == "lgssm" else time_varying_case)(n, ny, T, seed=100 * n + ny)
    ssm = _lib.LinearizedSSM(*[g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, g(case["ys"]), g(case["m0"]), g(case["L0"]), smooth=True,
                                               loglik=True, chunk_len=K)
    torch.cuda.synchronize()
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    errs = [rel_err(fm.cpu().numpy(), ofm), rel_err(LLt(fL.cpu().numpy()), LLt(ofc)), rel_err(sm.cpu().numpy(), osm),
            rel_err(LLt(sL.cpu().numpy()), LLt(osc)), abs(ell.item() - oell) / abs(oell)]
    up = float(torch.triu(fL, 1).abs().max()), float(torch.triu(sL, 1).abs().max())
    worst = max(worst, *[e if np.isfinite(e) else 1e9 for e in errs])
    print(f"mask={os.environ.get('PSQRT_COOP', 'default')} {kind} n={n} ny={ny} T={T} K={K}: fm {errs[0]:.1e} fLLt {errs[1]:.1e} "
          f"sm {errs[2]:.1e} sLLt {errs[3]:.1e} ell {errs[4]:.1e} upper {up}", flush=True)
print("WORST", worst)
