"""Multi-GPU form of the trace-and-validate step: shard the path candidates, gather the survivors.

The reference has no distributed code at all; every ``(tx, rx, candidate)`` unit of
``_trace_path_candidates`` (reference ``differt/src/differt/geometry/_solvers.py:576-717``) is
independent given the mesh, so the candidate axis is split into contiguous shards, one per rank
(one process per GPU), with the mesh, ``tx`` and ``rx`` replicated.  The hot loop has no
communication.  The only exchange is ONE all-gather of a fixed-capacity record per rank holding the
count, flat global index, vertices and objects of the valid paths the rank found — what
``TracedPaths.masked()`` (reference ``_paths.py:299-328``) would return — after which every rank
sorts the union back into the reference's row-major ``(tx, rx, candidate)`` order.

The gather logic (record layout, merge, overflow handling) is device-agnostic so that it is covered
by ``gloo`` tests on CPU; the compaction that fills the record is the CUDA library's
``drt_compact_valid_paths``.
"""

from __future__ import annotations

import dataclasses

import torch
import torch.distributed as dist

__all__ = [
    "trace_valid_paths_sharded",
    "GatherRecord",
    "ValidPaths",
    "gather_valid_paths",
    "global_path_index",
    "global_path_index_receivers",
    "receiver_shard",
    "shard_bounds",
    "trace_path_candidates_sharded",
]


def shard_bounds(num_candidates: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous near-equal split of ``range(num_candidates)``: the first ``C % world`` ranks get one
    extra candidate.  Concatenating the shards in rank order gives back the original order."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    base, extra = divmod(int(num_candidates), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def global_path_index(
    local_index: torch.Tensor, num_local: int, num_global: int, start: int
) -> torch.Tensor:
    """Flat index into the global ``[Ntx, Nrx, C]`` array of a flat index into a rank's
    ``[Ntx, Nrx, C_local]`` array, when the rank holds candidates ``start .. start + C_local``."""
    if num_local == 0:
        return local_index
    pair = torch.div(local_index, num_local, rounding_mode="floor")
    return pair * num_global + (local_index - pair * num_local) + start


RX_BLOCK = 16  # receivers per block of the block-cyclic deal


def receiver_shard(num_rx: int, world_size: int, rank: int, block: int = RX_BLOCK) -> torch.Tensor:
    """Receivers dealt block-cyclically: blocks of ``block`` consecutive receivers go round-robin to the
    ranks, and every rank traces its receivers against EVERY candidate.  The cost of a path depends on
    where its receiver stands and on which candidate it follows: the deal gives every rank a sample of
    the whole receiver set and the full candidate list, so the shards cost the same to within a few
    percent — contiguous shards of a few hundred candidates differ by tens of percent (measured:
    4.76 ms vs 3.71 ms at 8 ranks) — while consecutive receivers of a rank stay neighbours, which the
    blockage traversal's cache locality needs (a stride-8 deal costs 15 % more per path)."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    idx = torch.arange(int(num_rx), dtype=torch.int64)
    return idx[(idx // max(int(block), 1)) % world_size == rank]


def global_path_index_receivers(
    local_index: torch.Tensor, num_candidates: int, receivers: torch.Tensor, num_rx_global: int
) -> torch.Tensor:
    """Flat index into the global ``[Ntx, Nrx, C]`` array of a flat index into a rank's
    ``[Ntx, Nrx_local, C]`` array, ``receivers`` being the global numbers of the rank's receivers."""
    num_rx_local = int(receivers.shape[0])
    if num_rx_local == 0 or num_candidates == 0:
        return local_index
    per_tx = num_rx_local * num_candidates
    itx = torch.div(local_index, per_tx, rounding_mode="floor")
    rem = local_index - itx * per_tx
    j = torch.div(rem, num_candidates, rounding_mode="floor")
    c = rem - j * num_candidates
    return (itx * num_rx_global + receivers.to(local_index.device)[j]) * num_candidates + c


@dataclasses.dataclass
class ValidPaths:
    """The valid paths of all ranks in the reference's compacted (row-major) order."""

    index: torch.Tensor     # [n] int64: flat index into the global [Ntx, Nrx, C] array
    vertices: torch.Tensor  # [n, k+2, 3] f32
    objects: torch.Tensor   # [n, k+2] i32
    counts: list[int]       # valid paths found by each rank

    @property
    def num_valid_paths(self) -> int:
        return int(self.index.shape[0])


class GatherRecord:
    """One rank's fixed-capacity record: ``[count i64 | index i64[cap] | vertices f32[cap,k+2,3] |
    objects i32[cap,k+2]]`` in a single byte buffer, so that a single all-gather moves everything."""

    def __init__(self, capacity: int, order: int, device) -> None:
        self.capacity, self.order = int(capacity), int(order)
        nv = order + 2
        self.off_index = 256
        self.off_vertices = self.off_index + _align(8 * self.capacity)
        self.off_objects = self.off_vertices + _align(12 * nv * self.capacity)
        self.nbytes = self.off_objects + _align(4 * nv * self.capacity)
        self.buffer = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)

    @staticmethod
    def views(buffer: torch.Tensor, capacity: int, order: int):
        """``(count [1] i64, index [cap] i64, vertices [cap,k+2,3] f32, objects [cap,k+2] i32)`` views
        of a record buffer (any device)."""
        nv = order + 2
        off_i = 256
        off_v = off_i + _align(8 * capacity)
        off_o = off_v + _align(12 * nv * capacity)
        count = buffer[0:8].view(torch.int64)
        index = buffer[off_i:off_i + 8 * capacity].view(torch.int64)
        vertices = buffer[off_v:off_v + 12 * nv * capacity].view(torch.float32).view(capacity, nv, 3)
        objects = buffer[off_o:off_o + 4 * nv * capacity].view(torch.int32).view(capacity, nv)
        return count, index, vertices, objects

    def fields(self):
        return self.views(self.buffer, self.capacity, self.order)


def _align(n: int) -> int:
    return (int(n) + 255) & ~255


def gather_valid_paths(record: GatherRecord, *, group=None) -> ValidPaths | None:
    """All-gather one record per rank (a single collective) and merge the survivors of all ranks in
    ascending global index, i.e. the reference's ``masked()`` order.

    ``record.fields()[1]`` must already hold GLOBAL flat indices.  Returns ``None`` if any rank found
    more valid paths than the record capacity (the caller retries with a larger record); the decision
    is taken from the gathered counts, so it is the same on every rank.
    """
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world > 1:
        gathered = torch.empty(world * record.nbytes, dtype=torch.uint8, device=record.buffer.device)
        dist.all_gather_into_tensor(gathered, record.buffer, group=group)
    else:
        gathered = record.buffer
    parts = [
        GatherRecord.views(gathered[r * record.nbytes:(r + 1) * record.nbytes], record.capacity, record.order)
        for r in range(world)
    ]
    counts = torch.cat([p[0] for p in parts]).cpu().tolist()  # the one device→host read
    if any(c > record.capacity for c in counts):
        return None
    index = torch.cat([p[1][:c] for p, c in zip(parts, counts)])
    vertices = torch.cat([p[2][:c] for p, c in zip(parts, counts)])
    objects = torch.cat([p[3][:c] for p, c in zip(parts, counts)])
    order = torch.argsort(index, stable=True)
    return ValidPaths(index[order], vertices[order], objects[order], [int(c) for c in counts])


def fill_record(record: GatherRecord, paths, num_global: int, start: int, *, receivers=None) -> None:
    """Compact the valid paths of a rank's dense ``TracedPaths`` into ``record`` on the device
    (``drt_compact_valid_paths``) and rewrite the indices as global ones.  No host synchronisation.
    ``receivers = (num_rx_global, global numbers of this rank's receivers)`` selects the receiver
    sharding (:func:`receiver_shard`) instead of contiguous candidate shards ``(num_global, start)``."""
    from ._lib import check, lib
    from ._tensor import numel, ptr, stream_ptr

    k = paths.order
    P = numel(paths.mask.shape)
    num_local = int(paths.mask.shape[-1])
    count, index, vertices, objects = record.fields()
    v = paths.vertices.detach().reshape(P, k + 2, 3)
    o = paths.objects.reshape(P, k + 2)
    m = paths.mask.reshape(P)
    if m.dtype.is_floating_point:  # relaxed trace: confidence >= threshold
        m = m >= paths.confidence_threshold
    m = m.view(torch.uint8) if m.dtype == torch.bool else m
    dev = v.device
    ws = torch.empty(max(lib.drt_compact_workspace_bytes(P), 1), dtype=torch.uint8, device=dev)
    check(
        lib.drt_compact_valid_paths(
            stream_ptr(), P, k, ptr(v), ptr(o), ptr(m), record.capacity, ptr(ws), ws.numel(),
            ptr(count), ptr(index), ptr(vertices), ptr(objects),
        )
    )
    if receivers is not None:
        nrx_global, mine = receivers
        if int(mine.shape[0]) != nrx_global:
            mine = mine.to(index.device)
            # entries beyond `count` are garbage (clamped here, never read)
            index.copy_(global_path_index_receivers(index.clamp_(0, max(P - 1, 0)), num_local, mine, nrx_global))
            # the receiver column holds local receiver numbers
            objects[:, -1] = mine[objects[:, -1].clamp_(0, max(int(mine.shape[0]) - 1, 0)).long()].to(torch.int32)
    elif num_local != num_global or start != 0:
        # entries beyond `count` are garbage and never read
        index.copy_(global_path_index(index, num_local, num_global, start))


def trace_path_candidates_sharded(
    mesh, tx_vertices, rx_vertices, path_candidates, *, group=None, capacity: int = 1 << 16,
    shard: str = "candidates", **kwargs
):
    """Trace this rank's shard of the ``(tx, rx, candidate)`` units and gather every rank's valid paths.

    ``shard="candidates"``: contiguous candidate shards — every rank passes the SAME full candidate
    array (or a tuple ``(num_global, start, local_shard)`` when the shards are generated per rank).
    ``shard="receivers"``: receivers dealt round-robin (:func:`receiver_shard`), every candidate on every
    rank — the balanced split when there are many receivers.  Returns ``(local TracedPaths, ValidPaths
    of all ranks)``; the merged list is the reference's ``masked()`` order either way.
    """
    from .solvers import trace_path_candidates

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if shard == "receivers":
        rx_all = torch.as_tensor(rx_vertices).reshape(-1, 3)
        mine = receiver_shard(rx_all.shape[0], world, rank).to(rx_all.device)
        paths = trace_path_candidates(mesh, tx_vertices, rx_all[mine], path_candidates, **kwargs)
        while True:
            record = GatherRecord(capacity, paths.order, paths.vertices.device)
            fill_record(record, paths, 0, 0, receivers=(int(rx_all.shape[0]), mine))
            valid = gather_valid_paths(record, group=group)
            if valid is not None:
                return paths, valid
            capacity *= 4
    if shard != "candidates":
        raise ValueError(f"shard must be 'candidates' or 'receivers', got {shard!r}")
    if isinstance(path_candidates, tuple):
        num_global, start, local = path_candidates
    else:
        num_global = int(path_candidates.shape[0])
        start, stop = shard_bounds(num_global, world, rank)
        local = path_candidates[start:stop]
    paths = trace_path_candidates(mesh, tx_vertices, rx_vertices, local, **kwargs)
    while True:
        record = GatherRecord(capacity, paths.order, paths.vertices.device)
        fill_record(record, paths, num_global, start)
        valid = gather_valid_paths(record, group=group)
        if valid is not None:
            return paths, valid
        capacity *= 4


def trace_valid_paths_sharded(mesh, tx_vertices, rx_vertices, order: int, *, group=None,
                              chunk_size: int = 1 << 20, solver: str = "exhaustive", **kwargs) -> ValidPaths:
    """Exhaustive (or hybrid) search of ``order`` sharded over the ranks: the candidate list is never
    materialised anywhere — every rank decodes and traces its own contiguous range of candidate
    indices with the compact kernel — and the valid paths of all ranks are exchanged with ONE
    all-gather and merged into the reference's ``masked()`` order on every rank."""
    from .solvers import _candidate_chunks, trace_valid_path_candidates

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    total, _ = _candidate_chunks(mesh, tx_vertices, rx_vertices, order, chunk_size, solver,
                                 kwargs.pop("num_rays", 1_000_000), kwargs.pop("accel", "brute"))
    start, stop = shard_bounds(total, world, rank)
    dev = mesh.vertices.device
    idx, verts, objs = [], [], []
    if solver == "exhaustive":
        from .solvers import generate_all_path_candidates

        decode = lambda s0, n: generate_all_path_candidates(  # noqa: E731
            mesh.num_primitives, order, assume_quads=mesh.assume_quads, start=s0, count=n, device=dev)
    else:
        from .solvers import generate_visible_path_candidates

        vis = generate_visible_path_candidates(mesh, tx_vertices, rx_vertices, order)
        decode = vis.chunk
    for s0 in range(start, stop, chunk_size):
        cand = decode(s0, min(chunk_size, stop - s0))
        part = trace_valid_path_candidates(mesh, tx_vertices, rx_vertices, cand, index_offset=(total, s0), **kwargs)
        idx.append(part.index), verts.append(part.vertices), objs.append(part.objects)
    k2 = order + 2
    index = torch.cat(idx) if idx else torch.zeros(0, dtype=torch.int64, device=dev)
    vertices = torch.cat(verts) if verts else torch.zeros((0, k2, 3), device=dev)
    objects = torch.cat(objs) if objs else torch.zeros((0, k2), dtype=torch.int32, device=dev)
    capacity = max(int(index.numel()), 1)
    if world > 1:  # agree on one record size: the largest local count
        cap = torch.tensor([capacity], dtype=torch.int64, device=dev)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        capacity = int(cap.item())
    record = GatherRecord(capacity, order, dev)
    count, r_index, r_vertices, r_objects = record.fields()
    n = int(index.numel())
    count[0] = n
    r_index[:n], r_vertices[:n], r_objects[:n] = index, vertices, objects
    return gather_valid_paths(record, group=group)
