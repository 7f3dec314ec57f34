#!/usr/bin/env python
"""Where does the blockage time of the bench workload go?  Buckets the candidates by how close their
segments come to lying in a coordinate plane (exact zero direction component / |d^_j| < 1e-3 / rest),
then times the dense trace and reads the executed-test counter for each bucket separately."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import differt_b200 as drt  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


def main() -> None:
    wl = bench.build_workload(bench.DEFAULT_WORKLOAD, 0, 1)
    mesh = drt.Mesh.from_numpy(wl["vertices"], wl["triangles"])
    tx, rx = torch.from_numpy(wl["tx"]).cuda(), torch.from_numpy(wl["rx"]).cuda()
    cand = torch.from_numpy(wl["cand"]).cuda()
    p = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=True)
    v = p.vertices[0]  # [rx, cand, 5, 3]
    score_zero = torch.zeros(cand.shape[0], device="cuda")
    score_graze = torch.zeros(cand.shape[0], device="cuda")
    for r0 in range(0, v.shape[0], 256):
        d = v[r0:r0 + 256, :, 1:] - v[r0:r0 + 256, :, :-1]
        ln = d.norm(dim=-1, keepdim=True)
        dh = (d / ln.clamp_min(1e-30)).abs()
        some_zero = ((d == 0).any(-1) & ~(d == 0).all(-1)).any(-1)             # [rx, cand]
        graze = ((dh < 1e-3) & (d != 0)).any(-1).any(-1)
        score_zero += some_zero.float().sum(0)
        score_graze += graze.float().sum(0)
    score_zero /= v.shape[0]
    score_graze /= v.shape[0]
    del p, v
    buckets = {
        "all": torch.ones_like(score_zero, dtype=torch.bool),
        "exact_zero_component": score_zero > 0.5,
        "near_grazing_1e-3": (score_graze > 0.5) & ~(score_zero > 0.5),
        "rest": ~(score_graze > 0.5) & ~(score_zero > 0.5),
    }
    out = {}
    for name, sel in buckets.items():
        c = cand[sel].contiguous()
        if c.shape[0] == 0:
            continue
        ms, res = timed(lambda: drt.trace_path_candidates(mesh, tx, rx, c, dense_blockage=True, with_stats=True))
        pairs = c.shape[0] * rx.shape[0]
        out[name] = {"candidates": int(c.shape[0]), "ms": ms, "us_per_1k_pairs": ms * 1e3 / pairs * 1e3,
                     "tests_per_pair": res.stats["tests_done"] / pairs, "blocked_or_invalid": 1 - float(res.mask.float().mean())}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
