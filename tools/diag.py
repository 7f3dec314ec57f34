"""Stage statistics of the bench workload (how many candidates each blockage pass sees)."""
import sys, json
sys.path.insert(0, ".")
import differt_b200 as drt
import bench
wl = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD, 0, 1)
mesh = drt.Mesh.from_numpy(wl["vertices"], wl["triangles"])
for dense in (True, False):
    p = drt.trace_path_candidates(mesh, wl["tx"], wl["rx"], wl["cand"], dense_blockage=dense, with_stats=True)
    P = p.mask.numel()
    v = p.vertices.reshape(P, -1)
    zero = int((v == 0).all(dim=1).sum().item())
    print(json.dumps({"dense": dense, "P": P, **p.stats, "valid": int(p.mask.sum().item()), "zeroed_nonfinite": zero}))
