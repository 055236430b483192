"""Time-sharded pass on >= 2 real GPUs (one process per GPU, torchrun): both exchange modes -- NCCL all-gathers and
P2P stores into peer-mapped buffers fused into the mid-scan / carry kernels (psqrt_peer) -- against the single-GPU pass on the same
sequence.  Skipped on boxes with fewer than two GPUs; the host logic of the sharding is covered on the CPU by
tests/test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_time_sharded_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    lines = [l for l in r.stdout.splitlines() if "[check_sharded]" in l]
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert len(lines) == 2 and "exchange=nccl" in lines[0] and "exchange=peer" in lines[1], lines
