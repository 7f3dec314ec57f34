"""Small driver for `ncu` launch lists: default-mode trace, trace VJP, K1, K5 at bench sizes."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench, differt_b200 as drt
wl = bench.build_workload(bench.DEFAULT_WORKLOAD, 0, 1)
dev = torch.device("cuda", 0)
mesh = drt.Mesh(torch.from_numpy(wl["vertices"]).to(dev).requires_grad_(True), torch.from_numpy(wl["triangles"]).to(dev))
tx = torch.from_numpy(wl["tx"]).to(dev).requires_grad_(True)
rx = torch.from_numpy(wl["rx"]).to(dev).requires_grad_(True)
cand = torch.from_numpy(wl["cand"]).to(dev)
for _ in range(2):
    p = drt.trace_path_candidates(mesh, tx, rx, cand)
    torch.autograd.grad(p.vertices, (mesh.vertices, tx, rx), torch.ones_like(p.vertices))
rng = np.random.default_rng(0)
n = 1 << 24
o = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).to(dev)
d = torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)).to(dev)
tv = torch.from_numpy(rng.normal(size=(n, 3, 3)).astype(np.float32)).to(dev)
for _ in range(2):
    drt.ray_intersect_triangle(o, d, tv)
n5 = 1 << 22
mv = tv[:n5].contiguous(); mn = torch.nn.functional.normalize(tv[n5:2 * n5], dim=-1).contiguous()
for _ in range(2):
    drt.image_method(o[:n5], d[:n5], mv, mn)
torch.cuda.synchronize()
