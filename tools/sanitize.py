"""Small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).

    gpurun -- compute-sanitizer --tool memcheck python tools/sanitize.py
"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import differt_b200 as drt
from differt_b200 import scenes
from differt_b200.distributed import trace_path_candidates_sharded

rng = np.random.default_rng(0)
v, t = scenes.urban_grid(11, 11)           # 1454 triangles: 3 tiles → head + ring both exercised
mesh = drt.Mesh.from_numpy(v, t)
tri = mesh.triangle_vertices.contiguous()
lo, hi = v.min(0), v.max(0)
o = torch.from_numpy(rng.uniform(lo, hi + [0, 0, 20], (6000, 3)).astype(np.float32)).cuda()
e = torch.from_numpy(rng.uniform(lo, hi + [0, 0, 20], (6000, 3)).astype(np.float32)).cuda()
d = e - o
print("any", int(drt.ray_intersect_any_triangle(o, d, tri).sum()))
print("mesh any", int(mesh.ray_intersect_any_triangle(o, d).sum()))
idx, tt = drt.first_triangle_hit_by_ray(o, d, tri)
print("first", int((idx >= 0).sum()))
print("visible", int(mesh.triangles_visible_from_vertex(torch.tensor([[150.0, 150.0, 60.0]]).cuda(), num_rays=20000).sum()))
# the same flat queries behind the exact cull (walk.cuh), which the public API only takes for large batches
from differt_b200 import geometry
_min_work, geometry._CULL_MIN_WORK = geometry._CULL_MIN_WORK, 0
print("culled any", int(drt.ray_intersect_any_triangle(o, d, tri).sum()), int(mesh.ray_intersect_any_triangle(o, d).sum()))
print("culled first", int((drt.first_triangle_hit_by_ray(o, d, tri)[0] >= 0).sum()))
print("culled visible", int(mesh.triangles_visible_from_vertex(torch.tensor([[150.0, 150.0, 60.0]]).cuda(), num_rays=20000).sum()))
geometry._CULL_MIN_WORK = _min_work
tx = np.array([[150.0, 150.0, 48.0]], np.float32)
rx = scenes.receivers_grid(v, 8)
for order, n in ((1, 1454), (2, 3000), (3, 3000), (6, 500)):
    cand = scenes.complete_graph_candidates(t.shape[0], 1) if order == 1 else scenes.sampled_candidates(t.shape[0], order, n)
    for dense in (False, True):
        p = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense, with_stats=True)
        print("trace", order, dense, p.num_valid_paths, p.stats)
# a dense batch large enough (>= 8 * 32768 paths) to run the ordering pass with its greedy rounds
rx_big = scenes.receivers_grid(v, 16, 8)
big = drt.trace_path_candidates(mesh, tx, rx_big, scenes.sampled_candidates(t.shape[0], 2, 2048), dense_blockage=True, with_stats=True)
print("dense with ordering pass", big.num_valid_paths, big.stats)
# a mesh of <= 512 triangles keeps the ordering pass + cascade of resident passes: dense, >= 262 144 paths
vs, ts_ = scenes.street_canyon(20)
small = drt.Mesh.from_numpy(vs, ts_)
rx_s = scenes.receivers_grid(vs, 16, 16)
sp = drt.trace_path_candidates(small, np.array([[100.0, 0.0, 45.0]], np.float32), rx_s,
                               scenes.sampled_candidates(ts_.shape[0], 2, 1100), dense_blockage=True, with_stats=True)
print("small mesh, ordering pass + cascade", sp.num_valid_paths, sp.stats)
comp = drt.trace_valid_path_candidates(mesh, tx, rx, scenes.complete_graph_candidates(t.shape[0], 1), capacity=7)
print("compact", comp.num_valid_paths)
print("bvh", int(mesh.ray_intersect_any_triangle(o, d, accel="bvh").sum()), int((mesh.first_triangle_hit_by_ray(o, d, accel="bvh")[0] >= 0).sum()))
print("smooth", float(drt.ray_intersect_any_triangle(o[:500], d[:500], tri, smoothing_factor=5.0).sum()))
sm = drt.trace_path_candidates(mesh, tx, rx[:4], scenes.sampled_candidates(t.shape[0], 2, 37), smoothing_factor=5.0)
print("smooth trace", float(torch.nan_to_num(sm.mask).sum()))
lp = drt.launch_paths(mesh, tx, rx[:8], 1, num_rays=2000, max_dist=1.0)
print("sbr", int(lp.masks.sum()), "mlm", int((drt.compute_tx_mlm(mesh, tx, max_order=1, dim_x=4, dim_y=4, num_rays=2000, receiver_height=1.5, min_x=0.0, max_x=300.0, min_y=0.0, max_y=300.0) != 0).sum()))
paths, valid = trace_path_candidates_sharded(mesh, tx, rx, torch.from_numpy(scenes.complete_graph_candidates(t.shape[0], 1)).cuda())
print("sharded", valid.num_valid_paths)
mg = drt.Mesh(mesh.vertices.clone().requires_grad_(True), mesh.triangles)
txg = torch.from_numpy(tx).cuda().requires_grad_(True)
p = drt.trace_path_candidates(mg, txg, rx, scenes.sampled_candidates(t.shape[0], 2, 512))
p.vertices.sum().backward()
print("vjp", float(mg.vertices.grad.abs().sum()), float(txg.grad.abs().sum()))
f = torch.from_numpy(rng.normal(size=(1000, 3)).astype(np.float32)).cuda().requires_grad_(True)
mv = torch.from_numpy(rng.normal(size=(1000, 4, 3)).astype(np.float32)).cuda()
mn = torch.nn.functional.normalize(torch.from_numpy(rng.normal(size=(1000, 4, 3)).astype(np.float32)).cuda(), dim=-1)
drt.image_method(f, f.detach() + 1.0, mv, mn).sum().backward()
gen = drt.VisiblePathCandidates(50, 3, rng.uniform(size=50) < 0.5, rng.uniform(size=50) < 0.5, None)
print("digraph", len(gen), int(gen.chunk().sum()))
# EM consumer: Fresnel coefficients, s/p bases, per-path coefficients with the fused accumulation (ragged last warp)
ep = drt.trace_paths(small, np.array([[100.0, 0.0, 45.0]], np.float32), rx_s[:37], 1)
n_r = torch.full((ts_.shape[0],), 2.0 - 0.3j, dtype=torch.complex64)
a, length, field, power = drt.em.path_coefficients(ep, small, n_r, 2.4e9, thickness=torch.full((ts_.shape[0],), 0.1),
                                                  accumulate=True)
print("em", int(a.numel()), float(power.sum()), float(field.abs().sum()))
(r_s, _), _ = drt.em.fresnel_coefficients(np.linspace(1.1, 3.0, 1001).astype(np.float32), 0.5)
(e_s, _), _ = drt.em.sp_directions(d[:1001], -d[:1001], torch.nn.functional.normalize(o[:1001], dim=-1))
print("fresnel", float(r_s.abs().sum()), float(e_s.abs().sum()))
torch.cuda.synchronize()
print("sanitize run complete")
