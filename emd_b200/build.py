"""In-tree build of ``libemd_b200.so`` (sm_100a only) with plain nvcc.

``python -m emd_b200.build`` or ``emd_b200.build.build()``.  The shared library
is written next to this file so it travels with the source tree; it is
git-ignored.  Rebuilds only when a source is newer than the library.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libemd_b200.so"
HOSTMATH_SRC = PKG.parent / "tests" / "hostmath" / "hostmath.cpp"      # test scaffolding, not product source
HOSTLIB = PKG.parent / "tests" / "hostmath" / "libemd_b200_hostmath.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",  # contraction allowed everywhere EXCEPT the c_* intrinsics (which ptxas never fuses)
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
] + [f"-D{k}={os.environ[k]}" for k in ("EMD_SEG_BATCHES",) if os.environ.get(k)]  # tuning experiments only


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: emd_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def sources():
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [Path(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    objs = []
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        if not force and not _stale(obj, [src] + list(CSRC.glob("*.cuh")) + [Path(__file__)]):
            continue
        cmd = [_nvcc(), "-c", str(src), "-o", str(obj)] + [f for f in NVCC_FLAGS if f != "-shared"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src.name}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [_nvcc(), "-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB


def build_hostmath(force: bool = False) -> Path:
    """Host-only compile of the kernels' math headers (``-ffp-contract=off``; source under ``tests/hostmath/``) so the
    CPU test-suite can check the product's canonical op order against the oracle
    without a GPU.  Test scaffolding: not used, and not loaded, by the product path."""
    src = HOSTMATH_SRC
    deps = [src] + list(CSRC.glob("*.cuh"))
    if not force and not _stale(HOSTLIB, deps):
        return HOSTLIB
    cmd = ["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
           f"-I{CSRC}", str(src), "-o", str(HOSTLIB)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on hostmath:\n{r.stdout}")
    return HOSTLIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
    print(build_hostmath(force="--force" in sys.argv))
