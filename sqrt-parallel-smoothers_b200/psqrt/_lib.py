"""ctypes binding of libpsqrt.so (C ABI in include/psqrt.h) for torch CUDA tensors.

PyTorch is plumbing here: device memory, the current stream, and torch.distributed.  Every
numerical kernel runs inside libpsqrt.so.  There is NO fallback: if the library is missing or
a tensor is not an fp64 CUDA tensor the call raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("PSQRT_LIB", os.path.join(_HERE, "libpsqrt.so"))  # PSQRT_LIB: tuning builds only
_lib = None

c_double_p = ctypes.c_void_p  # raw device addresses


class PsqrtError(RuntimeError):
    pass


class _Ssm(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("F", "cholQ", "b", "H", "cholR", "c")] + \
               [(n, ctypes.c_int64) for n in ("F_ts", "cholQ_ts", "b_ts", "H_ts", "cholR_ts", "c_ts")] + \
               [(n, ctypes.c_int64) for n in ("F_bs", "cholQ_bs", "b_bs", "H_bs", "cholR_bs", "c_bs")] + \
               [(n, ctypes.c_void_p) for n in ("hF", "hcholQ", "hb", "hH", "hcholR", "hc")] + \
               [("fused_model", ctypes.c_int32), ("fused_reserved", ctypes.c_int32), ("nom_m", ctypes.c_void_p),
                ("nom_bs", ctypes.c_int64), ("fused_params", ctypes.c_void_p)]


FUSED_CT_BEARINGS = 1


class Peer(ctypes.Structure):
    """psqrt_peer (include/psqrt.h): one phase of the time-shard exchange over peer-mapped memory."""
    _fields_ = [("bufs", ctypes.c_void_p), ("rank", ctypes.c_int32), ("n_ranks", ctypes.c_int32),
                ("batch", ctypes.c_int64), ("flags_off", ctypes.c_int64), ("ctr_off", ctypes.c_int64),
                ("data_off", ctypes.c_int64), ("slot", ctypes.c_int64), ("payload", ctypes.c_int64)]


class _SsmTangent(ctypes.Structure):
    """psqrt_ssm_tangent (include/psqrt.h): one tangent direction of the linearised model, covariance form."""
    _fields_ = [(n, ctypes.c_void_p) for n in ("dF", "dQ", "db", "dH", "dR", "dc")] + \
               [(n + "_ts", ctypes.c_int64) for n in ("dF", "dQ", "db", "dH", "dR", "dc")]


class Plan(ctypes.Structure):
    _fields_ = [("chunk_len", ctypes.c_int32), ("n_chunks", ctypes.c_int64), ("n_chunks_pad", ctypes.c_int64),
                ("n_warps", ctypes.c_int64), ("nf_filter", ctypes.c_int32), ("nf_smoother", ctypes.c_int32)]


EXPORTS = (
    "psqrt_version", "psqrt_error_string", "psqrt_supported", "psqrt_get_plan", "psqrt_get_plan_ssm",
    "psqrt_workspace_bytes",
    "psqrt_filter_smoother", "psqrt_smoother", "psqrt_filter_reduce", "psqrt_carry_filter", "psqrt_filter_apply",
    "psqrt_carry_smoother", "psqrt_smoother_apply", "psqrt_filter_elements", "psqrt_filter_scan",
    "psqrt_smoother_elements", "psqrt_smoother_scan", "psqrt_loglik_terms", "psqrt_filter_combine",
    "psqrt_smoother_combine", "psqrt_tria_batched", "psqrt_chol_update_batched", "psqrt_linearize_builtin",
    "psqrt_fp64_probe", "psqrt_peer_layout", "psqrt_sampler_workspace_bytes", "psqrt_sample_paths",
    "psqrt_tangent_workspace_bytes", "psqrt_filter_smoother_tangent", "psqrt_loglik_adjoint",
    "psqrt_cov_tangent_to_chol",
    "psqrt_linearize_builtin_tangent", "psqrt_count_nonfinite", "psqrt_supported_generic",
    "psqrt_generic_workspace_bytes", "psqrt_filter_smoother_generic", "psqrt_filter_smoother_f32", "psqrt_tria_generic",
    "psqrt_chol_update_generic",
)

MODEL_CT_TRANSITION, MODEL_BEARINGS_OBSERVATION, MODEL_RICKER_TRANSITION, MODEL_POISSON_OBSERVATION = 1, 2, 3, 4
LIN_EXTENDED, LIN_SLR = 0, 1


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load libpsqrt.so (built in-tree by sqrt-parallel-smoothers_b200/build.py).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise PsqrtError(f"{_LIB_PATH} is missing: build it with `python sqrt-parallel-smoothers_b200/build.py` "
                         f"(or __graft_entry__.build()).  There is no CPU / eager fallback.")
    lib = ctypes.CDLL(_LIB_PATH)
    lib.psqrt_error_string.restype = ctypes.c_char_p
    lib.psqrt_sampler_workspace_bytes.restype = ctypes.c_size_t
    lib.psqrt_sampler_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64]
    lib.psqrt_workspace_bytes.restype = ctypes.c_size_t
    lib.psqrt_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                          ctypes.c_int]
    lib.psqrt_get_plan.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int,
                                   ctypes.POINTER(Plan)]
    lib.psqrt_generic_workspace_bytes.restype = ctypes.c_size_t
    lib.psqrt_generic_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    lib.psqrt_tangent_workspace_bytes.restype = ctypes.c_size_t
    lib.psqrt_tangent_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64]
    lib.psqrt_peer_layout.restype = ctypes.c_int64
    lib.psqrt_peer_layout.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.POINTER(Peer),
                                      ctypes.POINTER(Peer)]
    _lib = lib
    return lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().psqrt_error_string(rc).decode()
        raise PsqrtError(f"{what} failed: {msg} (code {rc})")


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return ctypes.c_void_p(0)
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise PsqrtError(f"expected a contiguous fp64 CUDA tensor, got {t.dtype} on {t.device} "
                         f"(contiguous={t.is_contiguous()})")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_ws_cache = {}
_ws_gen = {}
WS_SLOT = 0   # staged calls of one pass must share a workspace; tests emulating several ranks on one GPU switch slots


def _ws_key(device: torch.device, slot: Optional[int] = None):
    slot = WS_SLOT if slot is None else slot
    return (device.index if device.index is not None else torch.cuda.current_device(), slot)


def workspace(nbytes: int, device: torch.device, slot: Optional[int] = None, *, writer: bool = True) -> torch.Tensor:
    """Per-(device, slot) scratch buffer.  Every call that may overwrite it (`writer`) starts a new generation:
    staged calls that read what an earlier stage left there (smoother_apply after filter_apply) check that no
    other call used the buffer in between (see _stage_token / _check_stage)."""
    key = _ws_key(device, slot)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    if writer:
        _ws_gen[key] = _ws_gen.get(key, 0) + 1
    return buf


def _stage_token(device):
    key = _ws_key(device)
    return (key, _ws_gen.get(key, 0), _ws_cache[key].data_ptr())


def _check_stage(token, what):
    if token is None:
        raise PsqrtError(f"{what}: the filtered trajectory must come from psqrt._lib.filter_apply of the same pass "
                         f"(its workspace holds the packed filtered states this stage reads)")
    key, gen, ptr = token
    buf = _ws_cache.get(key)
    if buf is None or buf.data_ptr() != ptr or _ws_gen.get(key, 0) != gen:
        raise PsqrtError(f"{what}: the workspace of the matching filter_apply was reused or reallocated by another "
                         f"psqrt call in between; run the stages of one pass back to back (or on separate WS_SLOTs)")
    return buf


def supported(nx: int, ny: int = 0) -> bool:
    return bool(load().psqrt_supported(int(nx), int(ny)))


def get_plan(nx: int, ny: int, T: int, batch: int = 1, chunk_len: int = 0, ssm: "Optional[LinearizedSSM]" = None) -> Plan:
    """The chunking the library uses; with `ssm`, the one it uses for that model (psqrt_get_plan_ssm: a time-invariant
    transition model carried by value gets fewer, longer chunks at nx <= 4)."""
    p = Plan()
    lib = load()
    if ssm is None:
        _check(lib.psqrt_get_plan(nx, ny, T, batch, chunk_len, ctypes.byref(p)), "psqrt_get_plan")
    else:
        keep = []
        s = ssm.struct(T, batch, keep)
        _check(lib.psqrt_get_plan_ssm(ctypes.byref(s), nx, ny, ctypes.c_int64(T), ctypes.c_int64(batch), chunk_len,
                                      ctypes.byref(p)), "psqrt_get_plan_ssm")
    return p


def supported_generic(nx: int, ny: int = 0) -> bool:
    return bool(load().psqrt_supported_generic(int(nx), int(ny)))


def _require(nx: int, ny: int, generic_ok: bool = False):
    """Tuned kernels: nx in {1,2,3,4,5,6,8}, ny in 1..4.  The whole-pass calls (`generic_ok`) fall back to the generic
    path of csrc/psqrt_generic.cu for any nx, ny <= 16 inside the library."""
    if supported(nx, ny) or (generic_ok and supported_generic(nx, ny)):
        return
    raise PsqrtError(f"libpsqrt.so has no kernels for nx={nx}, ny={ny} (tuned: nx in 1,2,3,4,5,6,8 and ny in 1..4; "
                     f"generic whole-pass path: nx, ny <= 16)")


class LinearizedSSM:
    """Per-step linearised model (F, cholQ, b, H, cholR, c).  Leading dims of each entry:
    [] (shared by every step and sequence), [T] (per step, shared by the sequences), [B, T] (per sequence and
    step) or [B, 1] (per sequence, time-invariant).  A single leading dim is ALWAYS time: a per-sequence
    time-invariant entry must be given as [B, 1, ...] (a bare [B, ...] entry raises unless B == T, where it
    cannot be told apart from a per-step entry and is read as one).

    `host`: optional {name: numpy array} mirrors of time-invariant entries (also picked up from a
    `_psqrt_host` attribute that psqrt.methods attaches to tensors it created from host data, as long as the
    tensor has not been modified since).  With mirrors for every entry of a fully shared model the kernels take
    the model by value (constant-bank operands) instead of loading it per step."""

    def __init__(self, F, cholQ, b, H=None, cholR=None, c=None, host=None):
        self.F, self.cholQ, self.b, self.H, self.cholR, self.c = F, cholQ, b, H, cholR, c
        self.host = dict(host or {})
        self.fused = None

    @classmethod
    def fused_ct_bearings(cls, nom_mean: torch.Tensor, params, cholQ, m_q, cholR=None, m_r=None):
        """The built-in bearings-only model linearised (extended) INSIDE the sweeps (psqrt_ssm.fused_model,
        csrc/psqrt_fused.cuh): nom_mean [T+1, 5] or [B, T+1, 5] device tensor; params = (dt, s1x, s1y, s2x, s2y) and the
        time-invariant noise (cholQ lower [5,5], m_q [5], cholR [2,2], m_r [2]) as HOST arrays."""
        obj = cls(None, None, None)
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        obj.fused = dict(nom=nom_mean.contiguous(), params=f64(params), cholQ=f64(cholQ), m_q=f64(m_q), cholR=f64(cholR),
                         m_r=f64(m_r))
        return obj

    def struct(self, T: int, batch: int, keep: list) -> _Ssm:
        s = _Ssm()
        if self.fused is not None:
            fz = self.fused
            nom = fz["nom"]
            if nom.shape[-2:] != (T + 1, 5) or nom.dim() not in (2, 3) or (nom.dim() == 3 and nom.shape[0] != batch):
                raise PsqrtError(f"fused model: nominal means {tuple(nom.shape)} do not match T + 1 = {T + 1}, nx = 5")
            keep.extend([nom] + [v for k, v in fz.items() if k != "nom" and v is not None])
            s.fused_model = FUSED_CT_BEARINGS
            s.nom_m = _ptr(nom).value
            s.nom_bs = (T + 1) * 5 if nom.dim() == 3 else 0
            s.fused_params = fz["params"].ctypes.data
            s.hcholQ, s.hb = fz["cholQ"].ctypes.data, fz["m_q"].ctypes.data
            if fz["cholR"] is not None:
                s.hcholR, s.hc = fz["cholR"].ctypes.data, fz["m_r"].ctypes.data
            return s
        for name, core in (("F", 2), ("cholQ", 2), ("b", 1), ("H", 2), ("cholR", 2), ("c", 1)):
            t = getattr(self, name)
            if t is None:
                setattr(s, name, None)
                continue
            t = t.contiguous()
            keep.append(t)
            lead = t.shape[:t.dim() - core]
            size = 1
            for d in t.shape[t.dim() - core:]:
                size *= d
            ts = bs = 0
            if len(lead) == 0:
                pass
            elif len(lead) == 1:
                if lead[0] != T:
                    raise PsqrtError(f"{name}: leading dim {lead[0]} != T={T} (a single leading dim is time; "
                                     f"per-sequence time-invariant entries must be [B, 1, ...])")
                ts = size
            elif len(lead) == 2:
                if lead[0] != batch or lead[1] not in (1, T):
                    raise PsqrtError(f"{name}: leading dims {tuple(lead)} != (batch={batch}, T={T})")
                ts = size if lead[1] == T else 0
                bs = size * lead[1]
            else:
                raise PsqrtError(f"{name}: too many leading dims {tuple(lead)}")
            setattr(s, name, _ptr(t).value)
            setattr(s, name + "_ts", ts)
            setattr(s, name + "_bs", bs)
            h = self.host.get(name)
            if h is None:
                src = getattr(self, name)
                mirror = getattr(src, "_psqrt_host", None)
                # a mirror attached by psqrt.methods._t is a private copy stamped with the tensor's version
                # counter: an in-place update of the device tensor since then makes it stale, and it is ignored
                if mirror is not None and getattr(src, "_psqrt_host_version", None) == src._version:
                    h = mirror
            if h is not None and len(lead) == 0:
                h = np.ascontiguousarray(h, dtype=np.float64)
                if h.shape == tuple(t.shape):
                    keep.append(h)
                    setattr(s, "h" + name, h.ctypes.data)
        return s


def _batchify(x: torch.Tensor, core: int):
    """-> ([B, ...] tensor, had_batch)"""
    if x.dim() == core:
        return x.unsqueeze(0), False
    return x, True


def filter_smoother(ssm: LinearizedSSM, y: torch.Tensor, m0: torch.Tensor, L0: torch.Tensor, *, smooth: bool = True,
                    loglik: bool = False, chunk_len: int = 0, generic: bool = False):
    """One pass.  y [T, ny] or [B, T, ny]; m0 [nx] / [B, nx]; L0 lower-triangular [nx, nx] / [B, nx, nx].
    Returns (fm, fL, sm, sL, ell) with sm/sL/ell None when not requested.  Dimensions outside the tuned set run on
    the library's generic path (nx, ny <= 16); `generic=True` forces that path (psqrt_filter_smoother_generic)."""
    lib = load()
    yb, had_b = _batchify(y, 2)
    B, T, ny = yb.shape
    m0b, _ = _batchify(m0, 1)
    L0b, _ = _batchify(L0, 2)
    nx = m0b.shape[-1]
    _require(nx, ny, generic_ok=True)
    if m0b.shape[0] != B:
        m0b = m0b.expand(B, nx)
        L0b = L0b.expand(B, nx, nx)
    yb, m0b, L0b = yb.contiguous(), m0b.contiguous(), L0b.contiguous()
    dev = yb.device
    fm = torch.empty((B, T + 1, nx), dtype=torch.float64, device=dev)
    fL = torch.empty((B, T + 1, nx, nx), dtype=torch.float64, device=dev)
    sm = torch.empty_like(fm) if smooth else None
    sL = torch.empty_like(fL) if smooth else None
    ell = torch.empty((B,), dtype=torch.float64, device=dev) if loglik else None
    keep = []
    s = ssm.struct(T, B, keep)
    if generic:
        nbytes = int(lib.psqrt_generic_workspace_bytes(nx, ctypes.c_int64(T), ctypes.c_int64(B), 0))
        ws = workspace(nbytes, dev)
        with torch.cuda.device(dev):
            rc = lib.psqrt_filter_smoother_generic(ctypes.byref(s), _ptr(yb), _ptr(m0b), _ptr(L0b), nx, ny,
                                                   ctypes.c_int64(T), ctypes.c_int64(B), _ptr(fm), _ptr(fL), _ptr(sm),
                                                   _ptr(sL), _ptr(ell), ctypes.c_void_p(ws.data_ptr()),
                                                   ctypes.c_size_t(ws.numel()), _stream())
        _check(rc, "psqrt_filter_smoother_generic")
    else:
        nbytes = lib.psqrt_workspace_bytes(0, nx, ny, T, B, chunk_len)
        ws = workspace(nbytes, dev)
        with torch.cuda.device(dev):
            rc = lib.psqrt_filter_smoother(ctypes.byref(s), _ptr(yb), _ptr(m0b), _ptr(L0b), nx, ny,
                                           ctypes.c_int64(T), ctypes.c_int64(B), chunk_len, _ptr(fm), _ptr(fL), _ptr(sm),
                                           _ptr(sL), _ptr(ell), ctypes.c_void_p(ws.data_ptr()),
                                           ctypes.c_size_t(ws.numel()), _stream())
        _check(rc, "psqrt_filter_smoother")
    if not had_b:
        fm, fL = fm[0], fL[0]
        sm, sL = (sm[0], sL[0]) if smooth else (None, None)
        ell = ell[0] if loglik else None
    fL._psqrt_lower = True      # written by the kernels with exactly zero upper triangles
    if sL is not None:
        sL._psqrt_lower = True
    return fm, fL, sm, sL, ell


def smoother(ssm: LinearizedSSM, fm: torch.Tensor, fL: torch.Tensor, *, chunk_len: int = 0):
    lib = load()
    fmb, had_b = _batchify(fm, 2)
    fLb, _ = _batchify(fL, 3)
    fmb, fLb = fmb.contiguous(), fLb.contiguous()
    B, Tp1, nx = fmb.shape
    T = Tp1 - 1
    _require(nx, 0, generic_ok=True)
    sm, sL = torch.empty_like(fmb), torch.empty_like(fLb)
    if T == 0:
        sm.copy_(fmb)
        sL.copy_(fLb)
    else:
        nbytes = lib.psqrt_workspace_bytes(0, nx, 0, T, B, chunk_len)
        ws = workspace(nbytes, fmb.device)
        keep = []
        s = ssm.struct(T, B, keep)
        with torch.cuda.device(fmb.device):
            rc = lib.psqrt_smoother(ctypes.byref(s), _ptr(fmb), _ptr(fLb), nx, ctypes.c_int64(T), ctypes.c_int64(B),
                                    chunk_len, _ptr(sm), _ptr(sL), ctypes.c_void_p(ws.data_ptr()),
                                    ctypes.c_size_t(ws.numel()), _stream())
        _check(rc, "psqrt_smoother")
    return (sm, sL) if had_b else (sm[0], sL[0])


# ---- staged calls (time-sharded runs; see psqrt/dist.py) ---------------------------------------
def _peer_arg(peer):
    return ctypes.byref(peer) if peer is not None else None


def peer_layout(nx: int, n_ranks: int, batch: int):
    """-> (words per exchange buffer, filter-phase Peer, smoother-phase Peer) with the offsets filled in."""
    f, s = Peer(), Peer()
    n = int(load().psqrt_peer_layout(int(nx), int(n_ranks), ctypes.c_int64(batch), ctypes.byref(f), ctypes.byref(s)))
    if n <= 0:
        raise PsqrtError(f"psqrt_peer_layout: unsupported nx={nx}")
    return n, f, s


def filter_reduce(ssm, y, nx, chunk_len=0, peer=None):
    lib = load()
    B, T, ny = y.shape
    _require(nx, ny)
    plan = get_plan(nx, ny, T, B, chunk_len)
    total = torch.empty((B, plan.nf_filter), dtype=torch.float64, device=y.device)
    ws = workspace(lib.psqrt_workspace_bytes(0, nx, ny, T, B, chunk_len), y.device)
    keep = []
    s = ssm.struct(T, B, keep)
    with torch.cuda.device(y.device):
        rc = lib.psqrt_filter_reduce(ctypes.byref(s), _ptr(y), nx, ny, ctypes.c_int64(T), ctypes.c_int64(B), chunk_len,
                                     _ptr(total), ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()),
                                     _peer_arg(peer), _stream())
    _check(rc, "psqrt_filter_reduce")
    return total


def carry_filter(totals, rank, m0, L0, peer=None):
    lib = load()
    B, nx = m0.shape
    cm, cL = torch.empty_like(m0), torch.empty_like(L0)
    with torch.cuda.device(m0.device):
        rc = lib.psqrt_carry_filter(_ptr(totals), rank, ctypes.c_int64(B), nx, _ptr(m0), _ptr(L0), _ptr(cm), _ptr(cL),
                                    _peer_arg(peer), _stream())
    _check(rc, "psqrt_carry_filter")
    return cm, cL


def filter_apply(ssm, y, carry_m, carry_L, *, smooth=True, loglik=False, chunk_len=0, peer=None):
    lib = load()
    B, T, ny = y.shape
    nx = carry_m.shape[-1]
    plan = get_plan(nx, ny, T, B, chunk_len)
    dev = y.device
    fm = torch.empty((B, T + 1, nx), dtype=torch.float64, device=dev)
    fL = torch.empty((B, T + 1, nx, nx), dtype=torch.float64, device=dev)
    ell = torch.empty((B,), dtype=torch.float64, device=dev) if loglik else None
    stotal = torch.empty((B, plan.nf_smoother), dtype=torch.float64, device=dev) if smooth else None
    ws = workspace(lib.psqrt_workspace_bytes(0, nx, ny, T, B, chunk_len), dev, writer=False)   # continues filter_reduce
    keep = []
    s = ssm.struct(T, B, keep)
    with torch.cuda.device(dev):
        rc = lib.psqrt_filter_apply(ctypes.byref(s), _ptr(y), _ptr(carry_m), _ptr(carry_L), nx, ny, ctypes.c_int64(T),
                                    ctypes.c_int64(B), chunk_len, _ptr(fm), _ptr(fL), _ptr(ell), _ptr(stotal),
                                    ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()), _peer_arg(peer),
                                    _stream())
    _check(rc, "psqrt_filter_apply")
    fm._psqrt_stage = _stage_token(dev)    # smoother_apply reads the packed states this call left in the workspace
    return fm, fL, ell, stotal


def carry_smoother(totals, rank, n_ranks, mT, LT, peer=None):
    """Without `peer`: totals [R, B, nf] all-gathered smoothing totals, (mT, LT) the last rank's last filtered state.
    With `peer`: totals [B, nf] is THIS rank's smoothing total and (mT, LT) its own last filtered state; the kernel
    publishes them in every rank's exchange buffer, waits for all ranks and folds the later shards' totals."""
    lib = load()
    B, nx = mT.shape
    cm, cL = torch.empty_like(mT), torch.empty_like(LT)
    with torch.cuda.device(mT.device):
        rc = lib.psqrt_carry_smoother(_ptr(totals), rank, n_ranks, ctypes.c_int64(B), nx, _ptr(mT), _ptr(LT), _ptr(cm),
                                      _ptr(cL), _peer_arg(peer), _stream())
    _check(rc, "psqrt_carry_smoother")
    return cm, cL


def smoother_apply(ssm, fm, fL, carry_m, carry_L, *, write_terminal=True, chunk_len=0):
    """Stage 5 of a time-sharded pass.  `fm`, `fL` must be the tensors returned by filter_apply of the SAME pass:
    the kernel reads the packed copy of the filtered states that call left in the workspace (psqrt.h), and this
    wrapper refuses to run if any other psqrt call touched that workspace in between."""
    lib = load()
    B, Tp1, nx = fm.shape
    T = Tp1 - 1
    sm, sL = torch.empty_like(fm), torch.empty_like(fL)
    ws = _check_stage(getattr(fm, "_psqrt_stage", None), "smoother_apply")
    if ws.numel() < lib.psqrt_workspace_bytes(0, nx, 0, T, B, chunk_len):
        raise PsqrtError("smoother_apply: workspace smaller than this stage needs (different T / batch / chunk_len "
                         "than the matching filter_apply?)")
    keep = []
    s = ssm.struct(T, B, keep)
    with torch.cuda.device(fm.device):
        rc = lib.psqrt_smoother_apply(ctypes.byref(s), _ptr(fm), _ptr(fL), _ptr(carry_m), _ptr(carry_L),
                                      int(write_terminal), nx, ctypes.c_int64(T), ctypes.c_int64(B), chunk_len,
                                      _ptr(sm), _ptr(sL), ctypes.c_void_p(ws.data_ptr()),
                                      ctypes.c_size_t(ws.numel()), _stream())
    _check(rc, "psqrt_smoother_apply")
    return sm, sL


# ---- element-level seams ------------------------------------------------------------------------
def filter_elements(ssm, y, m0=None, L0=None):
    lib = load()
    yb, had_b = _batchify(y, 2)
    yb = yb.contiguous()
    B, T, ny = yb.shape
    keep = []
    s = ssm.struct(T, B, keep)
    nx = ssm.F.shape[-1]
    _require(nx, ny)
    dev = yb.device
    A = torch.empty((B, T, nx, nx), dtype=torch.float64, device=dev)
    U, Z = torch.empty_like(A), torch.empty_like(A)
    b = torch.empty((B, T, nx), dtype=torch.float64, device=dev)
    eta = torch.empty_like(b)
    if m0 is not None:
        m0 = _batchify(m0, 1)[0].expand(B, nx).contiguous()
        L0 = _batchify(L0, 2)[0].expand(B, nx, nx).contiguous()
    with torch.cuda.device(dev):
        rc = lib.psqrt_filter_elements(ctypes.byref(s), _ptr(yb), _ptr(m0), _ptr(L0), nx, ny, ctypes.c_int64(T),
                                       ctypes.c_int64(B), _ptr(A), _ptr(b), _ptr(U), _ptr(eta), _ptr(Z), _stream())
    _check(rc, "psqrt_filter_elements")
    out = (A, b, U, eta, Z)
    return out if had_b else tuple(o[0] for o in out)


def filter_scan(A, b, U, eta, Z, *, chunk_len=0):
    lib = load()
    Ab, had_b = _batchify(A, 3)
    args = [_batchify(t, c)[0].contiguous() for t, c in ((A, 3), (b, 2), (U, 3), (eta, 2), (Z, 3))]
    B, T, nx, _ = Ab.shape
    _require(nx, 0)
    means = torch.empty((B, T, nx), dtype=torch.float64, device=Ab.device)
    chols = torch.empty((B, T, nx, nx), dtype=torch.float64, device=Ab.device)
    ws = workspace(lib.psqrt_workspace_bytes(1, nx, 0, T, B, chunk_len), Ab.device)
    with torch.cuda.device(Ab.device):
        rc = lib.psqrt_filter_scan(*[_ptr(t) for t in args], nx, ctypes.c_int64(T), ctypes.c_int64(B), chunk_len,
                                   _ptr(means), _ptr(chols), ctypes.c_void_p(ws.data_ptr()),
                                   ctypes.c_size_t(ws.numel()), _stream())
    _check(rc, "psqrt_filter_scan")
    return (means, chols) if had_b else (means[0], chols[0])


def smoother_elements(ssm, fm, fL):
    lib = load()
    fmb, had_b = _batchify(fm, 2)
    fLb = _batchify(fL, 3)[0].contiguous()
    fmb = fmb.contiguous()
    B, Tp1, nx = fmb.shape
    T = Tp1 - 1
    _require(nx, 0)
    keep = []
    s = ssm.struct(T, B, keep)
    g = torch.empty_like(fmb)
    E, D = torch.empty_like(fLb), torch.empty_like(fLb)
    with torch.cuda.device(fmb.device):
        rc = lib.psqrt_smoother_elements(ctypes.byref(s), _ptr(fmb), _ptr(fLb), nx, ctypes.c_int64(T),
                                         ctypes.c_int64(B), _ptr(g), _ptr(E), _ptr(D), _stream())
    _check(rc, "psqrt_smoother_elements")
    out = (g, E, D)
    return out if had_b else tuple(o[0] for o in out)


def smoother_scan(g, E, D, *, chunk_len=0):
    lib = load()
    gb, had_b = _batchify(g, 2)
    args = [_batchify(t, c)[0].contiguous() for t, c in ((g, 2), (E, 3), (D, 3))]
    B, n, nx = gb.shape
    _require(nx, 0)
    means = torch.empty((B, n, nx), dtype=torch.float64, device=gb.device)
    chols = torch.empty((B, n, nx, nx), dtype=torch.float64, device=gb.device)
    ws = workspace(lib.psqrt_workspace_bytes(1, nx, 0, n, B, chunk_len), gb.device)
    with torch.cuda.device(gb.device):
        rc = lib.psqrt_smoother_scan(*[_ptr(t) for t in args], nx, ctypes.c_int64(n), ctypes.c_int64(B), chunk_len,
                                     _ptr(means), _ptr(chols), ctypes.c_void_p(ws.data_ptr()),
                                     ctypes.c_size_t(ws.numel()), _stream())
    _check(rc, "psqrt_smoother_scan")
    return (means, chols) if had_b else (means[0], chols[0])


def loglik_terms(ssm, y, fm, fL):
    lib = load()
    yb, had_b = _batchify(y, 2)
    yb = yb.contiguous()
    fmb = _batchify(fm, 2)[0].contiguous()
    fLb = _batchify(fL, 3)[0].contiguous()
    B, T, ny = yb.shape
    nx = fmb.shape[-1]
    _require(nx, ny)
    keep = []
    s = ssm.struct(T, B, keep)
    terms = torch.empty((B, T), dtype=torch.float64, device=yb.device)
    with torch.cuda.device(yb.device):
        rc = lib.psqrt_loglik_terms(ctypes.byref(s), _ptr(yb), _ptr(fmb), _ptr(fLb), nx, ny, ctypes.c_int64(T),
                                    ctypes.c_int64(B), _ptr(terms), _stream())
    _check(rc, "psqrt_loglik_terms")
    return terms if had_b else terms[0]


def filter_combine(e1: Sequence[torch.Tensor], e2: Sequence[torch.Tensor]):
    """sqrt_filtering_operator on n pairs (leading axis), or on a single pair."""
    lib = load()
    single = e1[0].dim() == 2
    cores = (3, 2, 3, 2, 3)
    a1 = [(t.unsqueeze(0) if single else t).contiguous() for t in e1]
    a2 = [(t.unsqueeze(0) if single else t).contiguous() for t in e2]
    n, nx = a1[1].shape
    _require(nx, 0)
    outs = [torch.empty_like(t) for t in a1]
    with torch.cuda.device(a1[0].device):
        rc = lib.psqrt_filter_combine(*[_ptr(t) for t in a1], *[_ptr(t) for t in a2], nx, ctypes.c_int64(n),
                                      *[_ptr(t) for t in outs], _stream())
    _check(rc, "psqrt_filter_combine")
    del cores
    return tuple(o[0] for o in outs) if single else tuple(outs)


def smoother_combine(e1, e2):
    lib = load()
    single = e1[0].dim() == 1
    a1 = [(t.unsqueeze(0) if single else t).contiguous() for t in e1]
    a2 = [(t.unsqueeze(0) if single else t).contiguous() for t in e2]
    n, nx = a1[0].shape
    _require(nx, 0)
    outs = [torch.empty_like(t) for t in a1]
    with torch.cuda.device(a1[0].device):
        rc = lib.psqrt_smoother_combine(*[_ptr(t) for t in a1], *[_ptr(t) for t in a2], nx, ctypes.c_int64(n),
                                        *[_ptr(t) for t in outs], _stream())
    _check(rc, "psqrt_smoother_combine")
    return tuple(o[0] for o in outs) if single else tuple(outs)


def tria(A: torch.Tensor) -> torch.Tensor:
    """tria(A) of parsmooth/_utils.py:22-24 for [..., rows, cols] (rows in the compiled set)."""
    lib = load()
    rows, cols = A.shape[-2:]
    lead = A.shape[:-2]
    Ab = A.reshape(-1, rows, cols).contiguous()
    L = torch.empty((Ab.shape[0], rows, rows), dtype=torch.float64, device=A.device)
    if Ab.shape[0] > 0:
        with torch.cuda.device(A.device):
            rc = lib.psqrt_tria_batched(_ptr(Ab), _ptr(L), rows, cols, ctypes.c_int64(Ab.shape[0]), _stream())
        _check(rc, "psqrt_tria_batched")
    L = L.reshape(*lead, rows, rows)
    L._psqrt_lower = True
    return L


def chol_update_many(L: torch.Tensor, V: torch.Tensor, alpha: float) -> torch.Tensor:
    """cholesky_update_many of parsmooth/_utils.py:13-19: L [..., n, n], V [..., k, n]."""
    lib = load()
    n = L.shape[-1]
    k = V.shape[-2]
    lead = L.shape[:-2]
    Lb = L.reshape(-1, n, n).contiguous().clone()
    Vb = V.reshape(-1, k, n).contiguous()
    if Lb.shape[0] > 0:
        with torch.cuda.device(L.device):
            rc = lib.psqrt_chol_update_batched(_ptr(Lb), _ptr(Vb), n, k, ctypes.c_double(alpha),
                                               ctypes.c_int64(Lb.shape[0]), _stream())
        _check(rc, "psqrt_chol_update_batched")
    Lb = Lb.reshape(*lead, n, n)
    Lb._psqrt_lower = True            # _utils.py:79: the update zeroes the upper triangle
    return Lb


def sample_paths(g: torch.Tensor, E: torch.Tensor, D: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """x_t = E_t x_{t+1} + g_t + D_t eps_t backwards through the n smoothing elements for every sample
    (parsmooth/_pathwise_sampler.py:13-38): g [n, nx], E, D [n, nx, nx], eps [n, S, nx] -> samples [n, S, nx]."""
    lib = load()
    n_el, nx = g.shape
    S = eps.shape[1]
    assert E.shape == (n_el, nx, nx) and D.shape == (n_el, nx, nx) and eps.shape == (n_el, S, nx)
    g, E, D, eps = (t.contiguous() for t in (g, E, D, eps))
    out = torch.empty_like(eps)
    nbytes = int(lib.psqrt_sampler_workspace_bytes(nx, ctypes.c_int64(n_el), ctypes.c_int64(S)))
    if nbytes == 0:
        raise PsqrtError(f"psqrt_sample_paths: unsupported sizes nx={nx}, n={n_el}, S={S}")
    ws = workspace(nbytes, g.device)
    with torch.cuda.device(g.device):
        rc = lib.psqrt_sample_paths(_ptr(g), _ptr(E), _ptr(D), _ptr(eps), _ptr(out), nx, ctypes.c_int64(n_el),
                                    ctypes.c_int64(S), ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()),
                                    _stream())
    _check(rc, "psqrt_sample_paths")
    return out


_points_cache = {}


def _device_points(xi, wm, wc, device):
    """Sigma-point tables are tiny constants per (rule, n): uploaded once per device."""
    key = (xi.tobytes(), wm.tobytes(), wc.tobytes(), str(device))
    hit = _points_cache.get(key)
    if hit is None:
        hit = tuple(torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device) for a in (xi, wm, wc))
        _points_cache[key] = hit
    return hit


def linearize_builtin(model_id: int, params, lin_id: int, n_in: int, n_out: int, conditional: bool,
                      nom_m: torch.Tensor, nom_L: Optional[torch.Tensor] = None, m_q: Optional[torch.Tensor] = None,
                      chol_q: Optional[torch.Tensor] = None, points=None):
    """psqrt_linearize_builtin on [..., n] nominal means (and [..., n, n] factors for SLR).
    Returns (F [..., d, n], chol [..., d, d] or None, b [..., d])."""
    lib = load()
    lead = nom_m.shape[:-1]
    dev = nom_m.device
    m = nom_m.reshape(-1, n_in).contiguous()
    count = m.shape[0]
    F = torch.empty((count, n_out, n_in), dtype=torch.float64, device=dev)
    b = torch.empty((count, n_out), dtype=torch.float64, device=dev)
    need_chol = conditional or lin_id == LIN_SLR
    chol = torch.empty((count, n_out, n_out), dtype=torch.float64, device=dev) if need_chol else None
    L = xi = wm = wc = None
    P = 0
    if lin_id == LIN_SLR:
        L = nom_L.expand(*lead, n_in, n_in).reshape(-1, n_in, n_in).contiguous()
        xi, wm, wc = _device_points(*points, dev)
        P = xi.shape[0]
    pars = (ctypes.c_double * len(params))(*[float(v) for v in params])
    if count > 0:
        with torch.cuda.device(dev):
            rc = lib.psqrt_linearize_builtin(int(model_id), pars, int(lin_id), _ptr(xi), _ptr(wm), _ptr(wc), int(P),
                                             _ptr(m), _ptr(L), ctypes.c_int64(count),
                                             _ptr(m_q.contiguous() if m_q is not None else None),
                                             _ptr(chol_q.contiguous() if chol_q is not None else None),
                                             _ptr(F), _ptr(chol), _ptr(b), _stream())
        _check(rc, "psqrt_linearize_builtin")
    if need_chol:
        chol = chol.reshape(*lead, n_out, n_out)
        chol._psqrt_lower = True      # written by the kernel with an exactly zero upper triangle
    return F.reshape(*lead, n_out, n_in), chol, b.reshape(*lead, n_out)


# ---- gradient path: forward-mode tangents (csrc/psqrt_tangent.cu; host driver psqrt/grad.py) ------------------
def filter_smoother_tangent(ssm: LinearizedSSM, dssm: dict, y: torch.Tensor, fm, fL, sm=None, sL=None, dm0=None,
                            dP0=None, *, smooth: bool = True, loglik: bool = True):
    """psqrt_filter_smoother_tangent for one sequence.  `dssm`: {"dF", "dQ", "db", "dH", "dR", "dc"} -> tensor with the
    entry's own shape (time-invariant) or a leading [T] axis, or None (zero); dQ / dR are COVARIANCE tangents.
    (fm, fL, sm, sL): the primal trajectories of filter_smoother on the same inputs.
    Returns (dfm [T+1,nx], dfP [T+1,nx,nx], dsm, dsP, dell) -- covariance tangents; dsm/dsP/dell None if not requested."""
    lib = load()
    T, ny = y.shape
    nx = fm.shape[-1]
    dev = y.device
    keep = []
    s = ssm.struct(T, 1, keep)
    d = _SsmTangent()
    for name, core in (("dF", (nx, nx)), ("dQ", (nx, nx)), ("db", (nx,)), ("dH", (ny, nx)), ("dR", (ny, ny)),
                       ("dc", (ny,))):
        t = dssm.get(name)
        if t is None:
            setattr(d, name, None)
            continue
        t = t.contiguous()
        keep.append(t)
        if tuple(t.shape) == core:
            ts = 0
        elif tuple(t.shape) == (T,) + core:
            ts = int(np.prod(core))
        else:
            raise PsqrtError(f"{name}: shape {tuple(t.shape)} is neither {core} nor {(T,) + core}")
        setattr(d, name, _ptr(t).value)
        setattr(d, name + "_ts", ts)
    nbytes = int(lib.psqrt_tangent_workspace_bytes(nx, ny, ctypes.c_int64(T)))
    if nbytes == 0:
        raise PsqrtError(f"psqrt_filter_smoother_tangent: unsupported nx={nx}, ny={ny}")
    ws = workspace(nbytes, dev, slot=-1)
    dfm = torch.empty((T + 1, nx), dtype=torch.float64, device=dev)
    dfP = torch.empty((T + 1, nx, nx), dtype=torch.float64, device=dev)
    dsm = torch.empty_like(dfm) if smooth else None
    dsP = torch.empty_like(dfP) if smooth else None
    dell = torch.empty((1,), dtype=torch.float64, device=dev) if loglik else None
    args = [t.contiguous() if t is not None else None for t in (y, fm, fL, sm if smooth else None,
                                                                  sL if smooth else None, dm0, dP0)]
    with torch.cuda.device(dev):
        rc = lib.psqrt_filter_smoother_tangent(ctypes.byref(s), ctypes.byref(d), _ptr(args[0]), nx, ny,
                                               ctypes.c_int64(T), _ptr(args[1]), _ptr(args[2]), _ptr(args[3]),
                                               _ptr(args[4]), _ptr(args[5]), _ptr(args[6]), _ptr(dfm), _ptr(dfP),
                                               _ptr(dsm), _ptr(dsP), _ptr(dell), ctypes.c_void_p(ws.data_ptr()),
                                               ctypes.c_size_t(ws.numel()), _stream())
    _check(rc, "psqrt_filter_smoother_tangent")
    return dfm, dfP, dsm, dsP, (dell[0] if loglik else None)


def loglik_adjoint(ssm: LinearizedSSM, y: torch.Tensor, fm: torch.Tensor, fL: torch.Tensor) -> dict:
    """psqrt_loglik_adjoint for one sequence: reverse mode of the log-likelihood at the linearised model `ssm`.
    (fm, fL): the filtered trajectory of filter_smoother on the same inputs.  Returns
    {"lam" [T+1,nx], "Lam" [T+1,nx,nx], "gF" [T,nx,nx], "gQ" [T,nx,nx], "gb" [T,nx], "gH" [T,ny,nx], "gR" [T,ny,ny],
    "gc" [T,ny]}: d ell / d (filtered mean, filtered covariance) of every step (index 0: the prior) and d ell / d (model
    entries of every step), Q and R in COVARIANCE form."""
    lib = load()
    T, ny = y.shape
    nx = fm.shape[-1]
    dev = y.device
    keep = []
    s = ssm.struct(T, 1, keep)
    nbytes = int(lib.psqrt_tangent_workspace_bytes(nx, ny, ctypes.c_int64(T)))
    if nbytes == 0:
        raise PsqrtError(f"psqrt_loglik_adjoint: unsupported nx={nx}, ny={ny}")
    ws = workspace(nbytes, dev, slot=-1)
    new = lambda *shape: torch.empty(shape, dtype=torch.float64, device=dev)
    out = {"lam": new(T + 1, nx), "Lam": new(T + 1, nx, nx), "gF": new(T, nx, nx), "gQ": new(T, nx, nx),
           "gb": new(T, nx), "gH": new(T, ny, nx), "gR": new(T, ny, ny), "gc": new(T, ny)}
    yc, fmc, fLc = y.contiguous(), fm.contiguous(), fL.contiguous()
    with torch.cuda.device(dev):
        rc = lib.psqrt_loglik_adjoint(ctypes.byref(s), _ptr(yc), nx, ny, ctypes.c_int64(T), _ptr(fmc), _ptr(fLc),
                                      *[_ptr(out[k]) for k in ("lam", "Lam", "gF", "gQ", "gb", "gH", "gR", "gc")],
                                      ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()), _stream())
    _check(rc, "psqrt_loglik_adjoint")
    return out


def cov_tangent_to_chol(L: torch.Tensor, dP: torch.Tensor) -> torch.Tensor:
    """dL with d(L L^T) = dP for lower-triangular L [..., n, n]."""
    lib = load()
    n = L.shape[-1]
    Lc, dPc = L.reshape(-1, n, n).contiguous(), dP.reshape(-1, n, n).contiguous()
    dL = torch.empty_like(Lc)
    with torch.cuda.device(L.device):
        rc = lib.psqrt_cov_tangent_to_chol(_ptr(Lc), _ptr(dPc), _ptr(dL), n, ctypes.c_int64(Lc.shape[0]), _stream())
    _check(rc, "psqrt_cov_tangent_to_chol")
    return dL.reshape(L.shape)


def linearize_builtin_tangent(model_id: int, params, dparams, lin_id: int, n_in: int, n_out: int, conditional: bool,
                              nom_m, nom_L=None, dnom_m=None, dnom_L=None, dm_q=None, dQ_q=None, points=None):
    """psqrt_linearize_builtin_tangent on [count, n] nominal means: -> (dF [count,d,n], dQ [count,d,d] or None, db)."""
    lib = load()
    dev = nom_m.device
    m = nom_m.reshape(-1, n_in).contiguous()
    count = m.shape[0]
    dF = torch.empty((count, n_out, n_in), dtype=torch.float64, device=dev)
    db = torch.empty((count, n_out), dtype=torch.float64, device=dev)
    need_q = conditional or lin_id == LIN_SLR
    dQ = torch.empty((count, n_out, n_out), dtype=torch.float64, device=dev) if need_q else None
    L = dL = xi = wm = wc = None
    P = 0
    if lin_id == LIN_SLR:
        L = nom_L.expand(count, n_in, n_in).contiguous()
        dL = dnom_L.expand(count, n_in, n_in).contiguous() if dnom_L is not None else None
        xi, wm, wc = _device_points(*points, dev)
        P = xi.shape[0]
    c = lambda t: t.contiguous() if t is not None else None
    pars = (ctypes.c_double * len(params))(*[float(v) for v in params])
    dpars = (ctypes.c_double * len(params))(*[float(v) for v in dparams]) if dparams is not None else None
    with torch.cuda.device(dev):
        rc = lib.psqrt_linearize_builtin_tangent(int(model_id), pars, dpars, int(lin_id), _ptr(xi), _ptr(wm), _ptr(wc),
                                                 int(P), _ptr(m), _ptr(L), _ptr(c(dnom_m)), _ptr(dL),
                                                 ctypes.c_int64(count), _ptr(c(dm_q)), _ptr(c(dQ_q)), _ptr(dF),
                                                 _ptr(dQ), _ptr(db), _stream())
    _check(rc, "psqrt_linearize_builtin_tangent")
    return dF, dQ, db


def count_nonfinite(x: torch.Tensor) -> torch.Tensor:
    """Non-finite entries per leading index of x [rows, ...] -> int64 [rows] (psqrt_count_nonfinite): the NaN-rate
    report of the robustness sweeps (notebooks/robustness_100runs.py:41-77)."""
    lib = load()
    rows = x.shape[0]
    xc = x.reshape(rows, -1).contiguous()
    counts = torch.empty((rows,), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.psqrt_count_nonfinite(_ptr(xc), ctypes.c_int64(rows), ctypes.c_int64(xc.shape[1]),
                                       ctypes.c_void_p(counts.data_ptr()), _stream())
    _check(rc, "psqrt_count_nonfinite")
    return counts


def filter_smoother_f32(F, cholQ, b, H, cholR, c, y, m0, L0, *, smooth: bool = True, loglik: bool = False):
    """psqrt_filter_smoother_f32: the pass in float32 (generic path; notebooks/robustness_100runs.py runs the square-root
    smoother in float32).  Model entries: float32 CUDA tensors with their own shape (time-invariant) or a leading [T]
    axis; y [T, ny]; one sequence.  Returns (fm, fL, sm, sL, ell) -- float32 trajectories, ell a float64 scalar."""
    lib = load()
    T, ny = y.shape
    nx = m0.shape[-1]
    dev = y.device
    core = {"F": (nx, nx), "cholQ": (nx, nx), "b": (nx,), "H": (ny, nx), "cholR": (ny, ny), "c": (ny,)}
    arrs, ts = [], []
    for name, t in (("F", F), ("cholQ", cholQ), ("b", b), ("H", H), ("cholR", cholR), ("c", c)):
        t = t.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(t.shape) == core[name]:
            ts.append(0)
        elif tuple(t.shape) == (T,) + core[name]:
            ts.append(int(np.prod(core[name])))
        else:
            raise PsqrtError(f"{name}: shape {tuple(t.shape)} is neither {core[name]} nor {(T,) + core[name]}")
        arrs.append(t)
    f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
    y32, m032, L032 = f32(y), f32(m0), f32(L0)
    fm = torch.empty((T + 1, nx), dtype=torch.float32, device=dev)
    fL = torch.empty((T + 1, nx, nx), dtype=torch.float32, device=dev)
    sm = torch.empty_like(fm) if smooth else None
    sL = torch.empty_like(fL) if smooth else None
    ell = torch.empty((1,), dtype=torch.float64, device=dev) if loglik else None
    nbytes = int(lib.psqrt_generic_workspace_bytes(nx, ctypes.c_int64(T), ctypes.c_int64(1), 1))
    if nbytes == 0:
        raise PsqrtError(f"psqrt_filter_smoother_f32: unsupported nx={nx}")
    ws = workspace(nbytes, dev, slot=-2)
    p32 = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    ts_arr = (ctypes.c_int64 * 6)(*ts)
    bs_arr = (ctypes.c_int64 * 6)(0, 0, 0, 0, 0, 0)
    with torch.cuda.device(dev):
        rc = lib.psqrt_filter_smoother_f32(*[p32(t) for t in arrs], ts_arr, bs_arr, p32(y32), p32(m032), p32(L032), nx,
                                           ny, ctypes.c_int64(T), ctypes.c_int64(1), p32(fm), p32(fL), p32(sm),
                                           p32(sL), _ptr(ell), ctypes.c_void_p(ws.data_ptr()),
                                           ctypes.c_size_t(ws.numel()), _stream())
    _check(rc, "psqrt_filter_smoother_f32")
    return fm, fL, sm, sL, (ell[0] if loglik else None)
