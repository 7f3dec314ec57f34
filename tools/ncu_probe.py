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
# flat queries at 2^18 rays x 10 094 triangles: the all-pairs engine (C ABI) and the culled traversal
from differt_b200._lib import check, lib
from differt_b200._tensor import ptr, stream_ptr
from differt_b200.geometry import pack_mesh, sort_pack_by_area
R, T = 1 << 18, wl["triangles"].shape[0]
lo, hi = wl["vertices"].min(0), wl["vertices"].max(0)
ro = rng.uniform(lo, hi, size=(R, 3)).astype(np.float32); re = rng.uniform(lo, hi, size=(R, 3)).astype(np.float32)
ro[:, 2] = rng.uniform(0.5, 45.0, R); re[:, 2] = rng.uniform(0.5, 45.0, R)
ro_d, rd_d = torch.from_numpy(ro).to(dev), torch.from_numpy((re - ro).astype(np.float32)).to(dev)
pack = pack_mesh(mesh.vertices.detach(), mesh.triangles, None)
spack = sort_pack_by_area(pack, T)
hit = torch.empty(R, dtype=torch.uint8, device=dev); idx = torch.empty(R, dtype=torch.int32, device=dev); tt = torch.empty(R, device=dev)
ws = torch.empty(lib.drt_any_hit_workspace_bytes(T), dtype=torch.uint8, device=dev)
for _ in range(2):
    check(lib.drt_ray_intersect_any_triangle(stream_ptr(), R, ptr(ro_d), ptr(rd_d), ptr(spack), T, 1.19e-6, 1.19e-5, ptr(hit), None))
    check(lib.drt_first_triangle_hit_by_ray(stream_ptr(), R, ptr(ro_d), ptr(rd_d), ptr(pack), T, 1.19e-6, 512, ptr(idx), ptr(tt), None))
    check(lib.drt_ray_intersect_any_triangle_culled(stream_ptr(), R, ptr(ro_d), ptr(rd_d), ptr(spack), T, 1.19e-6, 1.19e-5, ptr(ws), ws.numel(), ptr(hit), None))
    check(lib.drt_first_triangle_hit_by_ray_culled(stream_ptr(), R, ptr(ro_d), ptr(rd_d), ptr(pack), T, 1.19e-6, 512, ptr(ws), ws.numel(), ptr(idx), ptr(tt), None))
torch.cuda.synchronize()
