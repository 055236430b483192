"""torchrun --nproc-per-node R tools/check_sharded.py : time-sharded pass (both exchange modes) against the
single-GPU pass on the same sequence.  Prints the largest relative deviations; exits non-zero above 1e-9."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "sqrt-parallel-smoothers_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from bench import make_lgssm, simulate
from psqrt import _lib, dist as pdist
from psqrt._lib import LinearizedSSM


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    nx, ny, Tl = 4, 2, 50_000
    model = make_lgssm(nx, ny)
    ys_full = simulate(model, Tl * world, seed=7)
    g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    ssm = LinearizedSSM(*[g(model[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")],
                        host={k: model[k] for k in ("F", "cholQ", "b", "H", "cholR", "c")})
    m0, L0 = g(model["m0"]), g(model["L0"])
    y_loc = g(ys_full[rank * Tl:(rank + 1) * Tl])
    _, _, sm_ref, sL_ref, ell_ref = _lib.filter_smoother(ssm, g(ys_full), m0, L0, smooth=True, loglik=True)
    worst = 0.0
    for mode in ("nccl", "peer"):
        sh = pdist.TimeShardedSmoother(nx, ny, Tl, device=dev, exchange=mode)
        for it in range(3):      # several passes: epochs / parity of the peer exchange
            fm, fL, sm, sL, ell = sh.filter_smoother(ssm, y_loc[None], m0[None], L0[None], smooth=True, loglik=True)
        torch.cuda.synchronize()
        ref_m = sm_ref[rank * Tl:(rank + 1) * Tl + 1]
        ref_P = sL_ref[rank * Tl:(rank + 1) * Tl + 1]
        em = float((sm[0] - ref_m).abs().max() / ref_m.abs().max())
        P, Pr = sL[0] @ sL[0].transpose(-1, -2), ref_P @ ref_P.transpose(-1, -2)
        eP = float((P - Pr).abs().max() / Pr.abs().max())
        el = float((ell[0] - ell_ref).abs() / ell_ref.abs())
        t = torch.tensor([em, eP, el], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[check_sharded] world={world} exchange={sh.exchange} (requested {mode}) err={sh.exchange_error} "
                  f"max rel err: mean {t[0]:.2e}  LL^T {t[1]:.2e}  ell {t[2]:.2e}", flush=True)
        worst = max(worst, float(t.max()))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if worst < 1e-9 else 1)


if __name__ == "__main__":
    main()
