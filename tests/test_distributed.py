"""Host-side logic of the multi-GPU path on CPU: world-size-2 ``gloo`` process group.

Each rank traces its contiguous candidate shard with the C oracle (standing in for the CUDA kernels,
which need a GPU), fills a ``GatherRecord`` exactly like ``fill_record`` does on the device, and the
product code ``gather_valid_paths`` all-gathers and merges.  The merged list must equal the
single-process oracle's ``masked()`` order (reference ``_paths.py:299-328``: row-major over
``[Ntx, Nrx, C]``).
"""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _scene():
    from differt_b200 import scenes

    v, t = scenes.street_canyon(3)
    tx = np.array([[10.0, 0.0, 30.0], [12.0, 1.0, 25.0]], np.float32)
    rx = np.array([[x, y, 1.5] for x in (2.0, 11.0, 19.0) for y in (-6.0, 5.0)], np.float32)  # in the street
    cand = scenes.complete_graph_candidates(t.shape[0], 1)  # every triangle once: order-1 paths exist
    return v, t, tx, rx, cand


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, capacity: int, out_dir: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from differt_b200.distributed import GatherRecord, gather_valid_paths, global_path_index, shard_bounds
        from oracle import c_oracle as co

        v, t, tx, rx, cand = _scene()
        C = cand.shape[0]
        start, stop = shard_bounds(C, world, rank)
        ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand[start:stop], early_exit=True)
        k = cand.shape[1]
        local_index = torch.from_numpy(np.flatnonzero(em.reshape(-1)))
        n = local_index.numel()
        while True:
            record = GatherRecord(capacity, k, "cpu")
            count, index, vertices, objects = record.fields()
            count[0] = n
            m = min(n, capacity)
            index[:m] = global_path_index(local_index, stop - start, C, start)[:m]
            vertices[:m] = torch.from_numpy(ev.reshape(-1, k + 2, 3))[local_index[:m]]
            objects[:m] = torch.from_numpy(eo.reshape(-1, k + 2))[local_index[:m]]
            valid = gather_valid_paths(record)
            if valid is not None:
                break
            capacity *= 4  # same retry rule as trace_path_candidates_sharded
        np.savez(
            Path(out_dir) / f"rank{rank}.npz", index=valid.index.numpy(), vertices=valid.vertices.numpy(),
            objects=valid.objects.numpy(), counts=np.array(valid.counts), capacity=capacity,
        )
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("capacity", [4096, 2])  # 2 forces the overflow → retry path
def test_sharded_gather_matches_single_process_order(tmp_path, capacity):
    from oracle import c_oracle as co

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), capacity, str(tmp_path)), nprocs=world, join=True)
    v, t, tx, rx, cand = _scene()
    ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True)
    exp_index = np.flatnonzero(em.reshape(-1))
    assert exp_index.size > 2, "scene must have valid paths for the test to mean anything"
    k = cand.shape[1]
    for rank in range(world):
        got = np.load(tmp_path / f"rank{rank}.npz")
        np.testing.assert_array_equal(got["index"], exp_index)
        np.testing.assert_array_equal(got["vertices"].view(np.uint32),
                                      ev.reshape(-1, k + 2, 3)[exp_index].view(np.uint32))
        np.testing.assert_array_equal(got["objects"], eo.reshape(-1, k + 2)[exp_index])
        assert int(got["counts"].sum()) == exp_index.size
        assert int(got["capacity"]) >= int(got["counts"].max())


def test_shard_bounds_partition():
    from differt_b200.distributed import shard_bounds

    for C in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(C, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_global_path_index_roundtrip():
    from differt_b200.distributed import global_path_index

    ntx, nrx, C, start, n_local = 2, 3, 11, 4, 5
    local = torch.arange(ntx * nrx * n_local)
    g = global_path_index(local, n_local, C, start)
    pair, c = np.divmod(local.numpy(), n_local)
    np.testing.assert_array_equal(g.numpy(), pair * C + c + start)


def test_receiver_sharding_index_roundtrip():
    from differt_b200.distributed import global_path_index_receivers, receiver_shard

    ntx, nrx, C = 3, 11, 5
    full = np.arange(ntx * nrx * C).reshape(ntx, nrx, C)
    seen = []
    for world, block in ((1, 16), (2, 1), (4, 1), (2, 3), (3, 4)):
        seen.clear()
        for rank in range(world):
            mine = receiver_shard(nrx, world, rank, block)
            local = torch.arange(ntx * len(mine) * C)
            g = global_path_index_receivers(local, C, mine, nrx).numpy()
            np.testing.assert_array_equal(g, full[:, mine.numpy(), :].reshape(-1))
            seen.append(g)
        np.testing.assert_array_equal(np.sort(np.concatenate(seen)), full.reshape(-1))  # a partition


def test_record_layout_is_aligned_and_disjoint():
    from differt_b200.distributed import GatherRecord

    r = GatherRecord(37, 3, "cpu")
    count, index, vertices, objects = r.fields()
    assert index.shape == (37,) and vertices.shape == (37, 5, 3) and objects.shape == (37, 5)
    for t in (count, index, vertices, objects):
        assert t.data_ptr() % 8 == 0
    index.fill_(-1)
    vertices.fill_(1.5)
    objects.fill_(7)
    count[0] = 3
    assert int(count[0]) == 3 and bool((index == -1).all()) and bool((vertices == 1.5).all()) and bool((objects == 7).all())
