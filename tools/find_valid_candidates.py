"""Find path candidates that are VALID for the bench scene, with the GPU tracer in its default
(pruned) mode, and store them as a small fixture so that bench.py's workload contains real specular
paths (`valid paths/s` is part of BASELINE.json's metric) without any search at bench time.

    gpurun -- python tools/find_valid_candidates.py          (needs a GPU; writes tests/golden/)

tests/test_oracle_golden.py::test_bench_valid_candidates re-validates the fixture with the CPU oracle.
"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import bench
import differt_b200 as drt
from differt_b200 import scenes

wl = bench.build_workload("urban10k_1tx_4096rx_order3", 0, 1)
v, t, tx, rx = wl["vertices"], wl["triangles"], wl["tx"], wl["rx"]
T = t.shape[0]
mesh = drt.Mesh.from_numpy(v, t)
rx_sub = rx[:: max(1, rx.shape[0] // 256)][:256]


def valid_candidates(cand: np.ndarray, rx_pts: np.ndarray) -> np.ndarray:
    """Boolean [C]: candidate valid for at least one receiver."""
    out = np.zeros(cand.shape[0], bool)
    step = max(1, (1 << 24) // max(rx_pts.shape[0], 1))
    for s in range(0, cand.shape[0], step):
        p = drt.trace_path_candidates(mesh, tx, rx_pts, cand[s:s + step])
        out[s:s + step] = p.mask.any(dim=1)[0].cpu().numpy()
    return out


found = {}
c1 = scenes.complete_graph_candidates(T, 1)
ok1 = valid_candidates(c1, rx)
found[1] = c1[ok1]
print("order 1:", found[1].shape[0], "valid candidates")

# order 2: exhaustive over (triangles valid at order 1 ∪ their quad partners) x all triangles, both ways
seed = np.unique(np.concatenate([found[1][:, 0], found[1][:, 0] ^ 1]))
seed = seed[seed < T]
allt = np.arange(T, dtype=np.int32)
pairs = np.concatenate([
    np.stack(np.meshgrid(seed, allt, indexing="ij"), -1).reshape(-1, 2),
    np.stack(np.meshgrid(allt, seed, indexing="ij"), -1).reshape(-1, 2),
]).astype(np.int32)
pairs = np.unique(pairs[pairs[:, 0] != pairs[:, 1]], axis=0)
ok2 = valid_candidates(pairs, rx_sub)
found[2] = pairs[ok2]
print("order 2:", found[2].shape[0], "valid candidates of", pairs.shape[0])

# order 3: exhaustive over the triangles seen in valid order-1/2 candidates
S = np.unique(np.concatenate([found[1].ravel(), found[2].ravel()]))
if S.size > 400:
    S = S[:: int(np.ceil(S.size / 400))]
a, b, c = np.meshgrid(S, S, S, indexing="ij")
tr = np.stack([a.ravel(), b.ravel(), c.ravel()], -1).astype(np.int32)
tr = tr[(tr[:, 0] != tr[:, 1]) & (tr[:, 1] != tr[:, 2])]
ok3 = valid_candidates(tr, rx_sub)
found[3] = tr[ok3]
print("order 3:", found[3].shape[0], "valid candidates of", tr.shape[0], "from", S.size, "triangles")

np.savez_compressed(
    "tests/golden/urban10k_valid_candidates.npz",
    order1=found[1][:2048].astype(np.int32), order2=found[2][:2048].astype(np.int32),
    order3=found[3][:2048].astype(np.int32),
)
np.savez_compressed("gpurun_out/urban10k_valid_candidates.npz",
                    order1=found[1][:2048].astype(np.int32), order2=found[2][:2048].astype(np.int32),
                    order3=found[3][:2048].astype(np.int32))
