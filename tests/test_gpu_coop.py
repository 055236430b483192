"""GPU tests of the sub-warp sweeps (csrc/psqrt_coopsweep.cuh: K1 / K3 / K5 with one chunk per 8-lane group, the form
nx = 8 runs in by default).  The same pass must come out of every mix of per-thread and sub-warp sweeps (they share
all scratch layouts), for batches, the standalone smoother, time-varying models and 8-byte-aligned inputs (scalar
global accesses instead of 16-byte ones).  Reference: parsmooth/parallel/_filtering.py:13-61, _smoothing.py:14-57."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from _cases import LLt, lgssm_case, oracle_from_ssm, rel_err, time_varying_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _g(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=torch.device("cuda", 0))


def _ssm(case, **kw):
    from psqrt import _lib
    return _lib.LinearizedSSM(*[_g(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")], **kw)


def _check(case, fm, fL, sm, sL, ell, tol=1e-9):
    ofm, ofc, osm, osc, oell = oracle_from_ssm(case)
    assert rel_err(fm.cpu().numpy(), ofm) < tol
    assert rel_err(LLt(fL.cpu().numpy()), LLt(ofc)) < tol
    assert float(torch.triu(fL, 1).abs().max()) == 0.0
    if sm is not None:
        assert rel_err(sm.cpu().numpy(), osm) < tol
        assert rel_err(LLt(sL.cpu().numpy()), LLt(osc)) < tol
        assert float(torch.triu(sL, 1).abs().max()) == 0.0
    if ell is not None:
        assert abs(ell.item() - oell) <= 1e-8 * abs(oell)


@pytest.mark.parametrize("mask", [0, 1, 2, 4, 7])
def test_every_mix_of_sweeps(mask):
    """PSQRT_COOP is read once per process: one child per mask runs tools/check_coop.py (nine nx = 8 / 6 cases: ragged
    chunks, several scan units and groups, time-varying models, ny = 1..4) against the oracle."""
    env = dict(os.environ, PSQRT_COOP=str(mask))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_coop.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = [ln for ln in r.stdout.splitlines() if ln.startswith("WORST")]
    assert worst and float(worst[-1].split()[1]) < 1e-9, r.stdout[-3000:]


@pytest.mark.parametrize("n,ny,T,B,K", [(8, 4, 900, 3, 0), (8, 2, 257, 5, 4), (8, 4, 40, 2, 1)])
def test_batched_and_filter_only(n, ny, T, B, K):
    from psqrt import _lib
    cases = [lgssm_case(n, ny, T, seed=31 * n + s) for s in range(B)]
    stack = lambda k: np.stack([c[k] for c in cases])
    ssm = _lib.LinearizedSSM(*[_g(stack(k)[:, None]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, _g(stack("ys")), _g(stack("m0")), _g(stack("L0")), smooth=True,
                                               loglik=True, chunk_len=K)
    for s, case in enumerate(cases):
        _check(case, fm[s], fL[s], sm[s], sL[s], ell[s])
    fm2, fL2, _, _, ell2 = _lib.filter_smoother(ssm, _g(stack("ys")), _g(stack("m0")), _g(stack("L0")), smooth=False,
                                                loglik=True, chunk_len=K)
    assert rel_err(fm2.cpu().numpy(), fm.cpu().numpy()) < 1e-12
    assert rel_err(ell2.cpu().numpy(), ell.cpu().numpy()) < 1e-12


def test_standalone_smoother_and_time_varying():
    from psqrt import _lib
    case = time_varying_case(8, 4, 1500, seed=77)
    ssm = _ssm(case)
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, _g(case["ys"]), _g(case["m0"]), _g(case["L0"]), smooth=True,
                                               loglik=True)
    _check(case, fm, fL, sm, sL, ell)
    sm2, sL2 = _lib.smoother(_lib.LinearizedSSM(_g(case["F"]), _g(case["cholQ"]), _g(case["b"])), fm, fL)
    assert rel_err(sm2.cpu().numpy(), sm.cpu().numpy()) < 1e-9
    assert rel_err(LLt(sL2.cpu().numpy()), LLt(sL.cpu().numpy())) < 1e-9


def test_eight_byte_aligned_inputs():
    """Model arrays that are only 8-byte aligned: the sweeps fall back from 16-byte to 8-byte global accesses."""
    from psqrt import _lib
    case = lgssm_case(8, 4, 700, seed=5)

    def off(a):   # same values, data pointer 8 bytes past a 16-byte boundary
        buf = torch.empty(a.size + 1, dtype=torch.float64, device=torch.device("cuda", 0))
        view = buf[1:].view(a.shape)
        view.copy_(_g(a))
        assert view.data_ptr() % 16 == 8
        return view

    ssm = _lib.LinearizedSSM(*[off(case[k]) for k in ("F", "cholQ", "b", "H", "cholR", "c")])
    fm, fL, sm, sL, ell = _lib.filter_smoother(ssm, _g(case["ys"]), _g(case["m0"]), _g(case["L0"]), smooth=True,
                                               loglik=True)
    _check(case, fm, fL, sm, sL, ell)


def test_full_size_nx8_properties():
    """BASELINE.json configs[3] size on one GPU (nx = 8, ny = 4, T = 1e6): prefix agreement with the oracle,
    invariance to the chunk length (association order), terminal smoothed = filtered."""
    from psqrt import _lib
    T = 1_000_000
    case = lgssm_case(8, 4, T, seed=11)
    args = (_ssm(case), _g(case["ys"]), _g(case["m0"]), _g(case["L0"]))
    fm, fL, sm, sL, ell = _lib.filter_smoother(*args, smooth=True, loglik=True)
    fm2, fL2, sm2, sL2, ell2 = _lib.filter_smoother(*args, smooth=True, loglik=True, chunk_len=61)
    assert torch.isfinite(sm).all() and torch.isfinite(sL).all()
    assert rel_err(fm2.cpu().numpy(), fm.cpu().numpy()) < 1e-9
    assert rel_err(sm2.cpu().numpy(), sm.cpu().numpy()) < 1e-9
    assert rel_err(LLt(sL2[::97].cpu().numpy()), LLt(sL[::97].cpu().numpy())) < 1e-9
    assert abs(ell.item() - ell2.item()) <= 1e-10 * abs(ell.item())
    assert rel_err(sm[-1].cpu().numpy(), fm[-1].cpu().numpy()) == 0.0
    Tc = 3000
    pre = {k: (v[:Tc] if k == "ys" else v) for k, v in case.items()}
    ofm, ofc, _, _, _ = oracle_from_ssm(pre)
    assert rel_err(fm[:Tc + 1].cpu().numpy(), ofm) < 1e-9
    assert rel_err(LLt(fL[:Tc + 1].cpu().numpy()), LLt(ofc)) < 1e-9


@pytest.mark.parametrize("n,ny,T", [(4, 2, 200000), (3, 1, 150000), (5, 2, 20000), (4, 2, 3000)])
def test_plan_follows_the_transition_model_only(n, ny, T):
    """The chunk plan of a pass depends on the state dimension and on the form of the TRANSITION model (by value or read
    from memory: psqrt_get_plan_ssm), never on the observation part, so that every stage of a pass -- the backward sweep
    sees no observation model -- cuts the sequence the same way.  Host mirrors for the transition part only: forward
    sweeps read the model from memory, the backward sweep takes it by value, same plan; all three forms of the model
    give the same pass (association order differs: 1e-10) and agree with the oracle on a prefix."""
    from psqrt import _lib
    case = lgssm_case(n, ny, T, seed=17 * n + ny)
    names = ("F", "cholQ", "b", "H", "cholR", "c")
    dev_arrays = [_g(case[k]) for k in names]
    full = _lib.LinearizedSSM(*dev_arrays, host={k: case[k] for k in names})
    part = _lib.LinearizedSSM(*dev_arrays, host={k: case[k] for k in names[:3]})
    none = _lib.LinearizedSSM(*dev_arrays)
    p_full, p_part, p_none = (_lib.get_plan(n, ny, T, 1, 0, ssm=s) for s in (full, part, none))
    assert p_full.chunk_len == p_part.chunk_len
    assert p_none.chunk_len <= p_full.chunk_len
    if n <= 4 and T >= 100000:
        assert p_none.chunk_len < p_full.chunk_len           # a model read from memory gets more, shorter chunks
    assert _lib.get_plan(n, ny, T).chunk_len == p_none.chunk_len
    outs = [_lib.filter_smoother(ssm, _g(case["ys"]), _g(case["m0"]), _g(case["L0"]), smooth=True, loglik=True)
            for ssm in (full, part, none)]
    for o in outs[1:]:
        assert rel_err(o[0].cpu().numpy(), outs[0][0].cpu().numpy()) < 1e-10
        assert rel_err(o[2].cpu().numpy(), outs[0][2].cpu().numpy()) < 1e-10
        assert rel_err(LLt(o[3][::53].cpu().numpy()), LLt(outs[0][3][::53].cpu().numpy())) < 1e-10
        assert abs(o[4].item() - outs[0][4].item()) <= 1e-10 * abs(outs[0][4].item())
    Tc = min(T, 3000)
    pre = {k: (v[:Tc] if k == "ys" else v) for k, v in case.items()}
    ofm, ofc, osm, osc, _ = oracle_from_ssm(pre)
    assert rel_err(outs[0][0][:Tc + 1].cpu().numpy(), ofm) < 1e-9
    assert rel_err(LLt(outs[0][1][:Tc + 1].cpu().numpy()), LLt(ofc)) < 1e-9
    if Tc == T:
        assert rel_err(outs[0][2].cpu().numpy(), osm) < 1e-9
        assert rel_err(LLt(outs[0][3].cpu().numpy()), LLt(osc)) < 1e-9


def test_random_configurations():
    """tools/fuzz_coop.py: 40 random (nx, ny, T, chunk length, time-varying / by-value model, smoother on / off) passes
    against the oracle (nx = 8 on the sub-warp sweeps, nx <= 6 per thread with the current chunk plans)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_coop.py"), "40", "7"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = [ln for ln in r.stdout.splitlines() if ln.startswith("WORST")]
    assert worst and float(worst[-1].split()[1]) < 1e-9, r.stdout[-3000:]
