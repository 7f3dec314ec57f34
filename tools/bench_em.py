#!/usr/bin/env python
"""Kernel-only timing of the EM consumer (drt_em_path_coefficients, drt_em_fresnel_coefficients) on
synthetic compacted paths over the urban10k mesh: CUDA events on the launching stream, inputs larger
than L2.  Writes one JSON object (profiles/r2_em_kernel.json is a copy of a run on a B200)."""

from __future__ import annotations

import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import differt_b200 as drt  # noqa: E402
from differt_b200 import geometry, scenes  # noqa: E402
from differt_b200._lib import check, lib  # noqa: E402
from differt_b200._tensor import ptr, stream_ptr  # noqa: E402


def timed(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]))


def main() -> None:
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())
    hbm = float(peaks.get("hbm_gbs", peaks.get("hbm", {}).get("gbs", 6549.4))) if isinstance(peaks, dict) else 6549.4
    dev = torch.device("cuda", 0)
    v, t = scenes.urban_grid(29, 29)
    mesh = drt.Mesh.from_numpy(v, t)
    T = t.shape[0]
    pack = geometry.pack_mesh(mesh.vertices, mesh.triangles, None)
    g = torch.Generator(device=dev).manual_seed(0)
    out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": hbm, "triangles": T, "cases": []}
    n_r = torch.view_as_real((torch.rand(T, device=dev, generator=g) * 2 + 1.5).to(torch.complex64) - 0.2j).contiguous()
    thick = torch.where(torch.rand(T, device=dev, generator=g) < 0.5, 0.1, -1.0).float()
    for order, n in ((1, 1 << 23), (3, 1 << 22), (4, 1 << 22)):
        verts = (torch.rand((n, order + 2, 3), device=dev, generator=g) * 800).contiguous()
        objs = torch.randint(0, T, (n, order + 2), device=dev, generator=g, dtype=torch.int32)
        num_pairs = 4096
        pair = (torch.arange(n, device=dev) * num_pairs // n).to(torch.int64)
        a = torch.empty((n, 2), device=dev)
        length = torch.empty(n, device=dev)
        field = torch.zeros((num_pairs, 2), device=dev)
        power = torch.zeros(num_pairs, device=dev)

        def run(acc):
            check(lib.drt_em_path_coefficients(
                stream_ptr(), n, order, ptr(verts), ptr(objs), T, ptr(pack), ptr(n_r), ptr(thick), 2.4e9, 0, 0, ptr(a),
                ptr(length), ptr(pair) if acc else None, num_pairs if acc else 0, ptr(field) if acc else None,
                ptr(power) if acc else None))

        for acc in (False, True):
            ms = timed(lambda: run(acc))
            # algorithmic bytes per path: vertices + objects in, coefficient + length out (+ pair index);
            # the per-triangle normal / n_r / thickness gathers hit the 0.6 MB tables in L2
            nbytes = n * ((order + 2) * 16 + 12 + (8 if acc else 0))
            out["cases"].append({"order": order, "paths": n, "accumulate": acc, "ms": ms, "paths_per_s": n / ms * 1e3,
                                 "algorithmic_gbs": nbytes / ms * 1e-6, "frac_of_hbm": nbytes / ms * 1e-6 / hbm})
    n = 1 << 24
    nr = torch.view_as_real((torch.rand(n, device=dev, generator=g) * 2 + 1).to(torch.complex64) - 0.1j).contiguous()
    ct = torch.rand(n, device=dev, generator=g)
    outs = [torch.empty((n, 2), device=dev) for _ in range(4)]
    ms = timed(lambda: check(lib.drt_em_fresnel_coefficients(stream_ptr(), n, ptr(nr), 1, ptr(ct), 1, *(ptr(o) for o in outs))))
    nbytes = n * (8 + 4 + 32)
    out["fresnel"] = {"n": n, "ms": ms, "algorithmic_gbs": nbytes / ms * 1e-6, "frac_of_hbm": nbytes / ms * 1e-6 / hbm}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
