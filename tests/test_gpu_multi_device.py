"""GPU tests that need MORE than one device (skipped on a single-GPU box):

* one process driving two devices — the mode JAX uses, the stated integration target: the kernels'
  dynamic shared-memory attributes, the SM count and the stream are per device;
* two processes over NCCL: ``trace_path_candidates_sharded`` (candidate shards + ONE all-gather) must
  return, on every rank, exactly the single-GPU ``masked()`` list, which in turn equals the oracle's.
"""

from __future__ import annotations

import os
import socket
import sys
import time
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def _scene():
    from differt_b200 import scenes

    v, t = scenes.street_canyon(5)
    tx = np.array([[20.0, 0.0, 30.0], [22.0, 1.0, 25.0]], np.float32)
    rx = np.array([[x, y, 1.5] for x in (2.0, 11.0, 19.0, 33.0) for y in (-6.0, 5.0)], np.float32)
    return v, t, tx, rx


def test_one_process_two_devices():
    """any-hit, first-hit and the trace on cuda:1 after cuda:0 in ONE process, each equal to the oracle
    (fails with a launch error if the dynamic shared-memory attribute is only set on the first device)."""
    _need_two_gpus()
    import differt_b200 as drt
    from differt_b200 import scenes
    from oracle import c_oracle as co
    from oracle import differt_oracle as orc

    v, t, tx, rx = _scene()
    cand = scenes.complete_graph_candidates(t.shape[0], 2)
    rng = np.random.default_rng(3)
    o = rng.uniform(-10, 60, size=(5000, 3)).astype(np.float32)
    d = rng.uniform(-30, 30, size=(5000, 3)).astype(np.float32)
    tri = orc.triangle_vertices(v, t)
    exp_any = co.ray_intersect_any_triangle(o, d, tri)
    exp_idx, exp_t = co.first_triangle_hit_by_ray(o, d, tri)
    ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True)
    assert em.sum() > 0
    torch.cuda.set_device(0)
    for index in (0, 1, 0):  # cuda:1 is never the current device
        dev = torch.device("cuda", index)
        oc, dc, tc = (torch.from_numpy(x).to(dev) for x in (o, d, tri))
        hit = drt.ray_intersect_any_triangle(oc, dc, tc)
        assert hit.device == dev
        np.testing.assert_array_equal(hit.cpu().numpy(), exp_any)
        idx, tt = drt.first_triangle_hit_by_ray(oc, dc, tc)
        np.testing.assert_array_equal(idx.cpu().numpy(), exp_idx)
        np.testing.assert_array_equal(tt.cpu().numpy().view(np.uint32), exp_t.view(np.uint32))
        mesh = drt.Mesh(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev))
        for dense in (False, True):
            paths = drt.trace_path_candidates(mesh, torch.from_numpy(tx).to(dev), torch.from_numpy(rx).to(dev),
                                              torch.from_numpy(cand).to(dev), dense_blockage=dense)
            assert paths.mask.device == dev
            np.testing.assert_array_equal(paths.mask.cpu().numpy(), em)
            np.testing.assert_array_equal(paths.vertices.cpu().numpy().view(np.uint32), ev.view(np.uint32))
        assert paths.masked().vertices.shape[0] == int(em.sum())
    with pytest.raises(ValueError):  # operands on two devices are rejected, not silently mis-launched
        drt.ray_intersect_any_triangle(oc.to("cuda:0"), dc.to("cuda:1"), tc.to("cuda:0"))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import differt_b200 as drt
        from differt_b200 import scenes
        from differt_b200.distributed import trace_path_candidates_sharded, trace_valid_paths_sharded

        v, t, tx, rx = _scene()
        cand = torch.from_numpy(scenes.complete_graph_candidates(t.shape[0], 2)).to(dev)
        mesh = drt.Mesh(torch.from_numpy(v).to(dev), torch.from_numpy(t).to(dev))
        out = {}
        for capacity in (1 << 12, 2):  # 2 forces the overflow → retry path on every rank together
            _, valid = trace_path_candidates_sharded(mesh, tx, rx, cand, capacity=capacity, dense_blockage=True)
            out[f"index{capacity}"] = valid.index.cpu().numpy()
            out[f"vertices{capacity}"] = valid.vertices.cpu().numpy()
            out[f"objects{capacity}"] = valid.objects.cpu().numpy()
            out[f"counts{capacity}"] = np.array(valid.counts)
        _, valid = trace_path_candidates_sharded(mesh, tx, rx, cand, shard="receivers", dense_blockage=True)
        out["rx_index"] = valid.index.cpu().numpy()
        out["rx_vertices"] = valid.vertices.cpu().numpy()
        out["rx_objects"] = valid.objects.cpu().numpy()
        single = drt.trace_path_candidates(mesh, tx, rx, cand).masked()  # this rank alone, all candidates
        out["single_vertices"] = single.vertices.cpu().numpy()
        out["single_objects"] = single.objects.cpu().numpy()
        search = trace_valid_paths_sharded(mesh, tx, rx, 2)  # exhaustive search, sharded by candidate index
        out["search_index"] = search.index.cpu().numpy()
        out["search_vertices"] = search.vertices.cpu().numpy()
        np.savez(Path(out_dir) / f"rank{rank}.npz", **out)
    finally:
        dist.destroy_process_group()


def test_nccl_world_size_two_sharded_trace_equals_single_gpu_and_oracle(tmp_path):
    _need_two_gpus()
    import torch.multiprocessing as mp

    from differt_b200 import scenes
    from oracle import c_oracle as co

    world = 2
    # a rank that dies or a fabric that never finishes the rendezvous must fail the test, not hang the suite
    ctx = mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=False)
    deadline = time.monotonic() + 240.0
    while not ctx.join(timeout=5.0):
        if time.monotonic() > deadline:
            for proc in ctx.processes:
                if proc.is_alive():
                    proc.kill()
            pytest.fail("the two NCCL ranks did not finish within 240 s")
    v, t, tx, rx = _scene()
    cand = scenes.complete_graph_candidates(t.shape[0], 2)
    ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True)
    exp = np.flatnonzero(em.reshape(-1))
    assert exp.size > 2
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        for capacity in (1 << 12, 2):
            np.testing.assert_array_equal(got[f"index{capacity}"], exp)
            np.testing.assert_array_equal(got[f"vertices{capacity}"].view(np.uint32),
                                          ev.reshape(-1, 4, 3)[exp].view(np.uint32))
            np.testing.assert_array_equal(got[f"objects{capacity}"], eo.reshape(-1, 4)[exp])
            assert int(got[f"counts{capacity}"].sum()) == exp.size and got[f"counts{capacity}"].size == world
        np.testing.assert_array_equal(got["single_vertices"].view(np.uint32), got["vertices4096"].view(np.uint32))
        np.testing.assert_array_equal(got["single_objects"], got["objects4096"])
        np.testing.assert_array_equal(got["rx_index"], exp)  # receivers dealt round-robin: same merged list
        np.testing.assert_array_equal(got["rx_vertices"].view(np.uint32), got["vertices4096"].view(np.uint32))
        np.testing.assert_array_equal(got["rx_objects"], got["objects4096"])
        np.testing.assert_array_equal(got["search_index"], exp)
        np.testing.assert_array_equal(got["search_vertices"].view(np.uint32), got["vertices4096"].view(np.uint32))
