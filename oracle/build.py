"""Build recipe for the C oracle (test infrastructure).  ``python -m oracle.build``."""

from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "oracle.c"
LIB = HERE / "liboracle.so"

# -ffp-contract=off: never fuse a*b+c (the oracle's contract); no -ffast-math, no -march=native
# (the library is built in the CPU container and travels to the GPU box).
CFLAGS = ["-O3", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-Wall",
          "-Wno-maybe-uninitialized"]


def build_oracle(force: bool = False) -> Path:
    if not force and LIB.exists() and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    subprocess.run(["gcc", *CFLAGS, str(SRC), "-o", str(LIB), "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build_oracle(force=True))
