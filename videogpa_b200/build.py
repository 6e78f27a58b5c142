"""In-tree nvcc build of the C-ABI library (sm_100a only).

    python -m videogpa_b200.build [--force] [--verbose]

Every csrc/*.cu is compiled to an object with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3`` and linked into
``videogpa_b200/lib/libvideogpa_b200.so`` with the CUDA runtime linked statically, so the library
has no libcuda/libcudart load-time dependency and can be dlopen'ed on a CPU-only host (the CPU test
suite checks the exported symbols there). nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "lib" / "obj"
LIB = LIBDIR / "libvideogpa_b200.so"
INCLUDE = PKG.parent / "include"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-I", str(INCLUDE)]
# Files whose index / mask decisions must be reproducible bit-for-bit by the numpy oracle:
# no FMA contraction, IEEE division and sqrt.
STRICT_FP = {"mvcs.cu", "reproject.cu", "pointcloud.cu", "epipolar.cu", "consistency.cu"}
STRICT_FLAGS = ["--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the videogpa_b200 CUDA library cannot be built")
    return exe


def _sig(src: Path, flags: list[str]) -> str:
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, force: bool, verbose: bool) -> tuple[Path, bool, str]:
    flags = ARCH + COMMON + (STRICT_FLAGS if src.name in STRICT_FP else [])
    obj = OBJDIR / (src.stem + ".o")
    stamp = OBJDIR / (src.stem + ".sig")
    sig = _sig(src, flags)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == sig:
        return obj, False, ""
    cmd = [nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", str(src), "-o", str(obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{p.stdout}\n{p.stderr}")
    stamp.write_text(sig)
    return obj, True, p.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJDIR.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    if not srcs:
        raise RuntimeError("no CUDA sources under videogpa_b200/csrc")
    rebuilt = False
    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, did, log in ex.map(lambda s: _compile(s, force, verbose), srcs):
            objs.append(obj)
            rebuilt |= did
            if verbose and log:
                print(log, file=sys.stderr)
    if rebuilt or not LIB.exists():
        cmd = [nvcc()] + ARCH + ["-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", str(LIB)] + \
              [str(o) for o in objs]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
