"""Build libpsqrt.so (sm_100a) in-tree with nvcc.

    python sqrt-parallel-smoothers_b200/build.py [--force] [--jobs J]

One translation unit per state dimension nx (psqrt_inst.cu with -DPSQ_N=nx) so the fully
unrolled templates compile in parallel; objects are cached under build/, each keyed by the hash of
its own source, the headers it includes and the flags, so a change rebuilds only what depends on it.  The result is sqrt-parallel-smoothers_b200/psqrt/libpsqrt.so, which is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.environ.get("PSQRT_BUILD_DIR", os.path.join(HERE, "build"))   # tuning builds: use a separate directory
OUT = os.path.join(HERE, "psqrt", "libpsqrt.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ALL_NX = (1, 2, 3, 4, 5, 6, 8)    # state dimensions the dispatcher knows (psqrt_capi.cu)
# compiled state dimensions; PSQRT_NX_LIST="4,5" builds a development subset (the others then
# report PSQRT_EUNSUPPORTED at run time)
NX_LIST = tuple(int(v) for v in os.environ.get("PSQRT_NX_LIST", "").split(",") if v) or ALL_NX
MAX_NY = int(os.environ.get("PSQRT_MAX_NY", "4"))   # observation dimensions 1..MAX_NY for each nx (dev builds: fewer)
_INC = re.compile(r'\s*#\s*include\s*"([^"]+)"')

EXTRA = [f for f in os.environ.get("PSQRT_NVCC_EXTRA", "").split() if f]
OUT = os.environ.get("PSQRT_OUT", OUT)
NVCC_FLAGS = EXTRA + [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libpsqrt.so cannot be built")
    return nvcc


def _closure(src: str, seen=None) -> list:
    """`src` and every file it #include "..."s, transitively (csrc/ and include/)."""
    seen = seen if seen is not None else []
    if src in seen or not os.path.isfile(src):
        return seen
    seen.append(src)
    with open(src, "r") as f:
        for line in f:
            m = _INC.match(line)
            if not m:
                continue
            name = m.group(1)
            for base in (os.path.dirname(src), INCLUDE):
                cand = os.path.normpath(os.path.join(base, name))
                if os.path.isfile(cand):
                    _closure(cand, seen)
                    break
    return seen


def _digest(src: str, extra: str) -> str:
    """Hash of one translation unit: its source, the headers it includes, the flags."""
    h = hashlib.sha256()
    for path in sorted(_closure(src)):
        with open(path, "rb") as f:
            h.update(os.path.basename(path).encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(extra.encode())
    return h.hexdigest()[:16]


def _compile(args):
    src, obj, defs = args
    cmd = [_nvcc(), *NVCC_FLAGS, *defs, "-I", INCLUDE, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return obj


def build_xla_shim(verbose: bool = True):
    """The XLA FFI shim (csrc/xla/xla_ffi_shim.cc: libpsqrt.so's whole-pass entry points as JAX custom calls) needs
    jaxlib's headers (xla/ffi/api/ffi.h), which this image does not have: it is compiled into libpsqrt_xla.so only when
    PSQRT_XLA_INCLUDE (or an importable jaxlib) provides them; otherwise the file is syntax-checked against psqrt.h."""
    src = os.path.join(CSRC, "xla", "xla_ffi_shim.cc")
    if not os.path.exists(src):
        return None
    inc = os.environ.get("PSQRT_XLA_INCLUDE")
    if not inc:
        try:
            import jaxlib
            cand = os.path.join(os.path.dirname(jaxlib.__file__), "include")
            inc = cand if os.path.exists(os.path.join(cand, "xla", "ffi", "api", "ffi.h")) else None
        except Exception:
            inc = None
    gxx = shutil.which("g++")
    if not gxx:
        return None
    if not inc:
        r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-I", INCLUDE, src], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"xla_ffi_shim.cc does not parse against psqrt.h:\n{r.stderr}")
        if verbose:
            print("[psqrt build] xla_ffi_shim.cc: no XLA FFI headers here (jaxlib absent); syntax-checked only")
        return None
    out = os.path.join(os.path.dirname(OUT), "libpsqrt_xla.so")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(_nvcc())), "include")
    cmd = [gxx, "-std=c++17", "-O2", "-shared", "-fPIC", "-DPSQRT_HAVE_XLA_FFI", "-I", inc, "-I", INCLUDE, "-I", cuda_inc,
           src, "-L", os.path.dirname(OUT), "-lpsqrt", "-Wl,-rpath,$ORIGIN", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"xla shim build failed: {' '.join(cmd)}\n{r.stderr}")
    if verbose:
        print(f"[psqrt build] wrote {out}")
    return out


def build(force: bool = False, jobs: int | None = None, verbose: bool = True) -> str:
    os.makedirs(BUILD, exist_ok=True)
    units = []
    inst = os.path.join(CSRC, "psqrt_inst.cu")
    for n in ALL_NX:
        defs = [f"-DPSQ_N={n}", f"-DPSQ_MAX_NY={MAX_NY}"]
        if n not in NX_LIST:
            defs.append("-DPSQ_STUB")       # launch table symbol only, no kernels
        units.append((inst, os.path.join(BUILD, f"inst_n{n}_{_digest(inst, ' '.join(defs))}.o"), defs))
    for name in ("psqrt_capi.cu", "psqrt_models.cu", "psqrt_sampler.cu", "psqrt_tangent.cu", "psqrt_generic.cu"):
        src = os.path.join(CSRC, name)
        if os.path.exists(src):
            units.append((src, os.path.join(BUILD, f"{name[6:-3]}_{_digest(src, '')}.o"), []))
    tag = hashlib.sha256(" ".join(os.path.basename(u[1]) for u in units).encode()).hexdigest()[:16]
    stamp = OUT + ".stamp"
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read().strip() == tag:
        if verbose:
            print(f"[psqrt build] up to date: {OUT}")
        return OUT
    todo = [u for u in units if force or not os.path.exists(u[1])]
    jobs = jobs or min(len(todo), os.cpu_count() or 1) or 1
    if verbose:
        print(f"[psqrt build] compiling {len(todo)} translation units with {jobs} jobs (sm_100a)")
    # biggest first
    todo.sort(key=lambda u: -int(u[2][0].split("=")[1]) if u[2] else 0)
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        list(ex.map(_compile, todo))
    objs = [u[1] for u in units]
    cmd = [_nvcc(), "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(tag)
    build_xla_shim(verbose)
    # drop stale objects
    keep = {os.path.basename(o) for o in objs}
    for name in os.listdir(BUILD):
        if name.endswith(".o") and name not in keep:
            os.remove(os.path.join(BUILD, name))
    if verbose:
        print(f"[psqrt build] wrote {OUT}")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    build(force=a.force, jobs=a.jobs)
    sys.exit(0)
