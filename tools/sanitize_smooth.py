"""compute-sanitizer driver for the relaxed trace alone (orders 0-3, quads, masks, ragged sizes).

    gpurun -- compute-sanitizer --tool memcheck python tools/sanitize_smooth.py
"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import differt_b200 as drt
from differt_b200 import scenes

rng = np.random.default_rng(0)
v, t = scenes.urban_grid(3, 3)
T = t.shape[0] - t.shape[0] % 2
t = t[:T]
tx = np.array([[15.0, 15.0, 48.0], [40.0, -5.0, 30.0]], np.float32)
rx = scenes.receivers_grid(v, 3, 2)[:5]
for quads in (False, True):
    for masked in (False, True):
        mask = rng.uniform(size=T) > 0.3 if masked else None
        if quads and masked:
            mask[1::2] = mask[::2]
        mesh = drt.Mesh.from_numpy(v, t, mask, assume_quads=quads)
        for order in (0, 1, 2, 3, 8):
            cand = rng.integers(0, T, size=(37, order)).astype(np.int32) if order else np.empty((1, 0), np.int32)
            if quads:
                cand -= cand % 2
            p = drt.trace_path_candidates(mesh, tx, rx, cand, smoothing_factor=7.0)
            print(quads, masked, order, float(torch.nan_to_num(p.mask).sum()), p.num_valid_paths, p.masked().vertices.shape[0])
            # reverse mode: cotangents on the confidences and on the path vertices
            mg = drt.Mesh(mesh.vertices.clone().requires_grad_(True), mesh.triangles, mesh.mask, assume_quads=quads)
            txg = torch.from_numpy(tx).cuda().requires_grad_(True)
            rxg = torch.from_numpy(rx).cuda().requires_grad_(True)
            for a in (0.3, 7.0):
                q = drt.trace_path_candidates(mg, txg, rxg, cand, smoothing_factor=a)
                (torch.nan_to_num(q.mask).sum() + torch.nan_to_num(q.vertices).sum()).backward()
            print("   grad", float(torch.nan_to_num(mg.vertices.grad).abs().sum()), float(torch.nan_to_num(txg.grad).abs().sum()))
torch.cuda.synchronize()
print("sanitize run complete")
