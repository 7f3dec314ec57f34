import sys; sys.path.insert(0, ".")
import torch, bench, differt_b200 as drt
wl = bench.build_workload(bench.DEFAULT_WORKLOAD, 0, 1)
dev = torch.device("cuda", 0)
mesh = drt.Mesh.from_numpy(wl["vertices"], wl["triangles"])
tx, rx, cand = (torch.from_numpy(wl[k]).to(dev) for k in ("tx", "rx", "cand"))
for name, fn in (("default mode dense outputs", lambda: drt.trace_path_candidates(mesh, tx, rx, cand)),
                 ("compact", lambda: drt.trace_valid_path_candidates(mesh, tx, rx, cand))):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.3f ms" % (e0.elapsed_time(e1) / 50))
