"""Multi-GPU exhaustive search check: every rank traces its range of the candidate indices, one
all-gather, identical merged result on every rank and equal to the single-GPU result.

    gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_search.py
"""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
import differt_b200 as drt
from differt_b200 import scenes
from differt_b200.distributed import trace_valid_paths_sharded

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
v, t = scenes.street_canyon(41)
mesh = drt.Mesh.from_numpy(v, t)
lo, hi = v.min(0), v.max(0)
tx = np.array([[0.5 * (lo[0] + hi[0]), 0.0, 1.2 * hi[2]]], np.float32)
rx = scenes.receivers_grid(v, 16, 16)
for order in (1, 2):
    trace_valid_paths_sharded(mesh, tx, rx, order)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = trace_valid_paths_sharded(mesh, tx, rx, order)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ref = drt.trace_valid_paths(mesh, tx, rx, order)
    ok = torch.equal(got.index, ref.index) and torch.equal(got.vertices, ref.vertices) and torch.equal(got.objects, ref.objects)
    print(f"rank {rank}/{world} order {order}: {got.num_valid_paths} valid paths, per-rank counts {got.counts}, "
          f"{dt * 1e3:.2f} ms, equals single-GPU result: {ok}", flush=True)
    assert ok
if world > 1:
    dist.destroy_process_group()
