"""Diagnostic: where does the host stall inside a resident bench step?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench, differt_b200 as drt
from differt_b200.distributed import GatherRecord, fill_record

wl = bench.build_workload("urban10k_1tx_4096rx_order3", 0, 1)
dev = torch.device("cuda", 0)
mesh = drt.Mesh.from_numpy(wl["vertices"], wl["triangles"])
tx, rx, cand = (torch.from_numpy(wl[n]).to(dev) for n in ("tx", "rx", "cand"))
mesh = drt.Mesh(mesh.vertices.requires_grad_(True), mesh.triangles)
tx.requires_grad_(True); rx.requires_grad_(True)
cot = torch.ones((1, rx.shape[0], cand.shape[0], 5, 3), device=dev)
record = GatherRecord(1024, 3, dev)

def step(vjp=True, fill=True):
    p = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=True)
    if vjp:
        torch.autograd.grad(p.vertices, (mesh.vertices, tx, rx), cot)
    if fill:
        fill_record(record, p, wl["cand_global"], 0)
    return p

def timed(n, **kw):
    torch.cuda.synchronize()
    s0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    p = None
    host = []
    for _ in range(n):
        h0 = time.perf_counter(); p = step(**kw); host.append(time.perf_counter() - h0)
    e1.record(); torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print(kw, "ms/step", e0.elapsed_time(e1) / n, "host enqueue ms/step", 1e3 * np.mean(host), "max", 1e3 * np.max(host),
          "cudaMalloc calls", s1["num_device_alloc"] - s0["num_device_alloc"], "frees", s1["num_device_free"] - s0["num_device_free"],
          "retries", s1["num_alloc_retries"] - s0["num_alloc_retries"])

p = None
for _ in range(6):
    p = step()
timed(10)
timed(10)
timed(10, vjp=False)
timed(10, vjp=False, fill=False)
