"""Timings of the relaxed (smoothing_factor) trace step, forward and reverse mode,
with the NumPy oracle timed beside it on a bounded sample.  Not the contract bench (/bench.py).

    gpurun -- python tools/bench_relaxed.py [--json gpurun_out/relaxed.json]

Workload: BASELINE config 2's scene (street canyon, 986 triangles), 1 TX x 256 RX, 4096 sampled
order-2 candidates = 1.05e6 paths; the relaxed blockage is a SUM, so every one of the
P (k+1) T = 3.1e9 (segment, triangle) pairs is evaluated (8 sigmoids each): no early exit, no ordering.
Timing: CUDA events on the current stream, 3 warm-ups, median of 10.
"""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import differt_b200 as drt  # noqa: E402
from differt_b200 import scenes  # noqa: E402
from oracle import differt_oracle as orc  # noqa: E402  (CPU baseline leg only)


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    v, t = scenes.street_canyon(41)
    T, order, alpha = t.shape[0], 2, 30.0
    lo, hi = v.min(0), v.max(0)
    tx = np.array([[0.5 * (lo[0] + hi[0]), 0.0, 1.2 * hi[2]]], np.float32)
    rx = scenes.receivers_grid(v, 16, 16)
    cand = scenes.sampled_candidates(T, order, 4096)
    P = tx.shape[0] * rx.shape[0] * cand.shape[0]
    pairs = P * (order + 1) * T
    rows = []

    def row(name, ms, note=""):
        r = {"kernel": name, "ms": ms, "paths_per_s": P / (ms * 1e-3), "relaxed_tests_per_s": pairs / (ms * 1e-3), "note": note}
        rows.append(r)
        print(json.dumps(r), flush=True)

    mesh = drt.Mesh.from_numpy(v, t)
    txc, rxc, cc = (torch.from_numpy(x).to(dev) for x in (tx, rx, cand))
    ms = timed(lambda: drt.trace_path_candidates(mesh, txc, rxc, cc, smoothing_factor=alpha))
    row(f"N4 relaxed trace forward [986 tri, 1 x 256 RX x 4096 candidates, order 2, alpha {alpha}]", ms,
        "stage kernel + warp-per-path blockage sum")
    ms_hard = timed(lambda: drt.trace_path_candidates(mesh, txc, rxc, cc, dense_blockage=True))
    row("   same batch, hard trace with dense blockage (early exit + ordering) for scale", ms_hard)

    mg = drt.Mesh(mesh.vertices.clone().requires_grad_(True), mesh.triangles)
    txg, rxg = txc.clone().requires_grad_(True), rxc.clone().requires_grad_(True)
    paths = drt.trace_path_candidates(mg, txg, rxg, cc, smoothing_factor=alpha)
    gm = torch.ones_like(paths.mask)
    ms = timed(lambda: torch.autograd.grad(paths.mask, (mg.vertices, txg, rxg), gm, retain_graph=True))
    flagged = int(((paths.mask > 0) & (paths.mask < 1)).sum())
    row("N4 relaxed trace reverse mode (cotangent on the confidences) [same batch]", ms,
        f"{flagged} of {P} confidences strictly inside (0, 1); blockage adjoint runs only where 1 - blocked is the min and unclipped")

    # CPU: the NumPy restatement on a bounded sample of the same batch (4 RX x 256 candidates)
    s_rx, s_c = 4, 256
    t0 = time.perf_counter()
    orc.trace_path_candidates(v, t, tx, rx[:s_rx], cand[:s_c], smoothing_factor=alpha)
    dt = time.perf_counter() - t0
    sp = s_rx * s_c * (order + 1) * T
    r = {"kernel": "N4 NumPy oracle (CPU port), same scene", "ms": dt * 1e3, "relaxed_tests_per_s": sp / dt,
         "note": f"sample: 1 x {s_rx} RX x {s_c} candidates = {sp} pairs; single-threaded NumPy fp32"}
    rows.append(r)
    print(json.dumps(r), flush=True)
    if args.json:
        Path(args.json).write_text(json.dumps({"rows": rows}, indent=1))


if __name__ == "__main__":
    main()
