"""The C-ABI library loads and exports every symbol include/psqrt.h declares (no compute calls:
there is no GPU in the build container)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "psqrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(psqrt_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    from psqrt import _lib
    names = _declared()
    assert len(names) >= 20
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"libpsqrt.so does not export {n}"
    assert set(names) == set(_lib.EXPORTS)


def test_host_side_queries():
    """Entry points that never touch the device."""
    from psqrt import _lib
    lib = _lib.load()
    assert lib.psqrt_version() == 100
    assert lib.psqrt_error_string(-2).decode().startswith("state/observation")
    assert _lib.supported(4, 2) and _lib.supported(5, 2) and _lib.supported(8, 4) and _lib.supported(1, 1)
    assert not _lib.supported(7, 2) and not _lib.supported(4, 5) and not _lib.supported(9, 0)
    p = _lib.get_plan(4, 2, 1_000_000)
    assert p.chunk_len * p.n_chunks >= 1_000_000 and p.n_chunks_pad % 128 == 0 and p.n_warps * 32 == p.n_chunks_pad
    assert p.nf_filter == 2 * 16 + 3 * 4 and p.nf_smoother == (3 * 16 + 3 * 4) // 2
    p1 = _lib.get_plan(5, 2, 7, 1, 3)
    assert (p1.chunk_len, p1.n_chunks) == (3, 3)
    assert lib.psqrt_workspace_bytes(0, 4, 2, 1_000_000, 1, 0) > 0
    # nx = 7 has no tuned kernels: the whole-pass op is served by the generic path, the element scan is not
    assert lib.psqrt_workspace_bytes(0, 7, 2, 1000, 1, 0) == lib.psqrt_generic_workspace_bytes(7, 1000, 1, 0) > 0
    assert lib.psqrt_workspace_bytes(1, 7, 2, 1000, 1, 0) == 0
    assert lib.psqrt_workspace_bytes(0, 17, 2, 1000, 1, 0) == 0          # beyond the generic path too
    assert _lib.supported_generic(7, 2) and _lib.supported_generic(16, 16) and not _lib.supported_generic(17, 1)
    # argument validation happens before any launch
    s = _lib._Ssm()
    rc = lib.psqrt_filter_smoother(ctypes.byref(s), None, None, None, 4, 2, ctypes.c_int64(10), ctypes.c_int64(1), 0,
                                   None, None, None, None, None, None, ctypes.c_size_t(0), None)
    assert rc == -1
    rc = lib.psqrt_tria_batched(None, None, 7, 3, ctypes.c_int64(1), None)     # rows = 7: generic path, validates too
    assert rc == -1
    rc = lib.psqrt_tria_batched(None, None, 17, 3, ctypes.c_int64(1), None)    # beyond both paths
    assert rc == -2
