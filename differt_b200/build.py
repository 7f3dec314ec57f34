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


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None,
          out: Path | None = None) -> Path:
    """Compile every ``csrc/*.cu`` and link ``libdiffert_b200.so``.  ``extra_flags`` / ``out`` build a
    tuning variant elsewhere (``python -m differt_b200.build --variant NAME -DDRT_STAGES=3 ...``)."""
    global LIB, OBJ_DIR
    if out is not None:
        LIB, OBJ_DIR, force = out, out.parent / "_obj", True
        out.parent.mkdir(parents=True, exist_ok=True)
    if not force and not is_stale():
        return LIB
    OBJ_DIR.mkdir(exist_ok=True)
    exe = nvcc()
    flags = [*NVCC_FLAGS, *(extra_flags or [])]

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [exe, *flags, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
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
    # --no-undefined: an unresolved internal symbol must fail the build, not the first dlopen
    cmd = [exe, "-shared", "-o", str(tmp), *map(str, objs), "-cudart", "static", "-Xlinker", "--no-undefined"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    tmp.replace(LIB)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:  # tuning builds: variants/<name>/libdiffert_b200.so (+ DIFFERT_B200_LIB)
        name = sys.argv[sys.argv.index("--variant") + 1]
        defs = [a for a in sys.argv[1:] if a.startswith("-D")]
        print(build(extra_flags=defs, out=PKG.parent / "variants" / name / "libdiffert_b200.so",
                    verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
