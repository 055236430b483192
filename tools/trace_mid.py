"""Development aid: phase timeline of the mid-level scans from a -DPSQ_MID_TRACE build (PSQRT_LIB=...)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "sqrt-parallel-smoothers_b200"), ROOT]
import numpy as np
import torch
from bench import make_lgssm, simulate
from psqrt import _lib
from psqrt._lib import LinearizedSSM

dev = torch.device("cuda", 0)
model = make_lgssm(4, 2)
T = 1_000_000
g = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
names = ("F", "cholQ", "b", "H", "cholR", "c")
ssm = LinearizedSSM(*[g(model[k]) for k in names], host={k: model[k] for k in names})
ys = g(np.random.RandomState(0).randn(T, 2))
for _ in range(5):
    _lib.filter_smoother(ssm, ys, g(model["m0"]), g(model["L0"]), smooth=True, loglik=False)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * (2 * 128 * 16))()
lib.psqrt_debug_trace(buf)
tr = np.frombuffer(buf, dtype=np.uint64).reshape(2, 128, 16).astype(np.int64)
for kind, name in ((0, "filter mid scan"), (1, "smoother mid scan")):
    t = tr[kind]
    used = t[:, 0] > 0
    t0 = t[used, 0].min()
    print(f"== {name}: {used.sum()} CTAs; stamps in ns relative to the first CTA start (min / max over CTAs)")
    labels = ["start", "init done", "p0 loaded", "p0 levels done", "p0 stored", "p0 fenced", "ticket known",
              "p1 fence", "p1 loaded", "p1 levels done", "p1 stored"]
    for s, lab in enumerate(labels):
        col = t[used, s]
        col = col[col >= t0]
        if len(col):
            print(f"  {lab:16s} {col.min() - t0:8d} {col.max() - t0:8d}   (n={len(col)})")
