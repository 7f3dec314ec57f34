"""In-tree build of the CUDA library: ``python -m differt_b200.build [--force]``.

nvcc cross-compiles for sm_100a without a GPU; the resulting ``libdiffert_b200.so`` sits next to this
file so that it travels to the GPU box with the repository snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libdiffert_b200.so"
OBJ_DIR = PKG / "csrc" / "_obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    # Parity contract: a*b+c must round twice like the CPU reference. FMAs are only issued through
    # explicit __fmaf_rn in code documented as conservative culling.
    "-fmad=false",
    "-prec-div=true",
    "-prec-sqrt=true",
    "-ftz=false",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-DDRT_BUILDING",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found; cannot build libdiffert_b200.so")
    return exe


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps() -> list[Path]:
    return [*sources(), *CSRC.glob("*.cuh"), *INCLUDE.glob("*.h"), Path(__file__)]


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    OBJ_DIR.mkdir(exist_ok=True)
    exe = nvcc()

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [exe, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [exe, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    tmp.replace(LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
