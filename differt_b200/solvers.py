"""The fused trace-and-validate step behind ``Scene.trace_paths``.

Mirror of ``_trace_path_candidates`` (reference ``differt/src/differt/geometry/_solvers.py:499-770``)
with the same arguments and defaults; the candidate enumeration of ``ExhaustivePathTracer``
(``_solvers.py:803-848`` → ``differt-core/src/geometry/graph.rs:286-491``) is available as an on-device
decode (``generate_all_path_candidates``).
"""

from __future__ import annotations

from typing import Iterator

import torch

from . import _lib
from ._lib import check, lib
from ._tensor import F32_EPS, Placement, ptr, stream_ptr
from .mesh import Mesh, TracedPaths
from .scenes import num_complete_graph_candidates

__all__ = [
    "VisiblePathCandidates",
    "generate_visible_path_candidates",
    "generate_all_path_candidates",
    "generate_all_path_candidates_chunks_iter",
    "trace_path_candidates",
    "trace_valid_path_candidates",
    "trace_paths",
    "trace_paths_chunks_iter",
    "trace_valid_paths",
]


class _TraceVertices(torch.autograd.Function):
    """``vertices`` of the traced paths with the reference's gradient: through the image method to
    ``tx``, ``rx`` and ``mesh.vertices`` (mirror vertex gather + normals), none through the mask."""

    @staticmethod
    def forward(ctx, mesh_vertices, tx, rx, triangles, cand, out_vertices):
        ctx.save_for_backward(mesh_vertices, tx, rx, triangles, cand)
        return out_vertices

    @staticmethod
    def backward(ctx, g):
        mesh_vertices, tx, rx, triangles, cand = ctx.saved_tensors
        g = g.contiguous().to(torch.float32)
        g_v = torch.empty_like(mesh_vertices)
        g_tx = torch.empty_like(tx)
        g_rx = torch.empty_like(rx)
        check(
            lib.drt_trace_path_candidates_vjp(
                stream_ptr(), mesh_vertices.shape[0], triangles.shape[0], ptr(mesh_vertices),
                ptr(triangles), tx.shape[0], ptr(tx), rx.shape[0], ptr(rx), cand.shape[0], cand.shape[1],
                ptr(cand), ptr(g), ptr(g_tx), ptr(g_rx), ptr(g_v),
            )
        )
        return g_v, g_tx, g_rx, None, None, None


class _SmoothTrace(torch.autograd.Function):
    """Relaxed trace (``smoothing_factor``): ``(vertices, confidence, objects)`` with the gradient
    ``jax.grad`` gives on the reference's relaxed branch (``_solvers.py:576-713``) — through the
    confidence AND through the path vertices — to ``mesh.vertices``, ``tx`` and ``rx``."""

    @staticmethod
    def forward(ctx, mesh_vertices, tx, rx, triangles, cand, mask_u8, quads, eps, tol, min_len, alpha):
        dev = mesh_vertices.device
        ntx, nrx, (C, k) = tx.shape[0], rx.shape[0], cand.shape
        V, T = mesh_vertices.shape[0], triangles.shape[0]
        out_v = torch.empty((ntx, nrx, C, k + 2, 3), dtype=torch.float32, device=dev)
        out_o = torch.empty((ntx, nrx, C, k + 2), dtype=torch.int32, device=dev)
        out_f = torch.empty((ntx, nrx, C), dtype=torch.float32, device=dev)
        ws = torch.empty(max(lib.drt_trace_smooth_workspace_bytes(T, ntx, nrx, C), 1), dtype=torch.uint8, device=dev)
        check(
            lib.drt_trace_path_candidates_smooth(
                stream_ptr(), V, T, ptr(mesh_vertices), ptr(triangles), ptr(mask_u8), int(quads), ntx, ptr(tx),
                nrx, ptr(rx), C, k, ptr(cand), eps, tol, min_len, alpha, ptr(ws), ws.numel(), ptr(out_v),
                ptr(out_o), ptr(out_f),
            )
        )
        ctx.save_for_backward(mesh_vertices, tx, rx, triangles, cand, mask_u8, out_v, out_f)
        ctx.params = (quads, eps, tol, min_len, alpha)
        ctx.mark_non_differentiable(out_o)
        ctx.set_materialize_grads(False)
        return out_v, out_f, out_o

    @staticmethod
    def backward(ctx, g_v, g_f, _g_o):
        mesh_vertices, tx, rx, triangles, cand, mask_u8, out_v, out_f = ctx.saved_tensors
        quads, eps, tol, min_len, alpha = ctx.params
        ntx, nrx, (C, k) = tx.shape[0], rx.shape[0], cand.shape
        V, T = mesh_vertices.shape[0], triangles.shape[0]
        g_mv, g_tx, g_rx = torch.zeros_like(mesh_vertices), torch.zeros_like(tx), torch.zeros_like(rx)
        if g_v is None and g_f is None:
            return (g_mv, g_tx, g_rx) + (None,) * 8
        g_v = None if g_v is None else g_v.contiguous().to(torch.float32)
        g_f = None if g_f is None else g_f.contiguous().to(torch.float32)
        ws = torch.empty(max(lib.drt_trace_smooth_vjp_workspace_bytes(V, T, ntx, nrx, C, k), 1), dtype=torch.uint8,
                         device=mesh_vertices.device)
        if g_f is None:
            g_f = torch.zeros_like(out_f)
        check(
            lib.drt_trace_path_candidates_smooth_vjp(
                stream_ptr(), V, T, ptr(mesh_vertices), ptr(triangles), ptr(mask_u8), int(quads), ntx, ptr(tx),
                nrx, ptr(rx), C, k, ptr(cand), eps, tol, min_len, alpha, ptr(out_v), ptr(out_f), ptr(g_v),
                ptr(g_f), ptr(ws), ws.numel(), ptr(g_tx), ptr(g_rx), ptr(g_mv),
            )
        )
        return (g_mv, g_tx, g_rx) + (None,) * 8


def trace_path_candidates(
    mesh: Mesh,
    tx_vertices,
    rx_vertices,
    path_candidates,
    interaction_types=None,
    *,
    epsilon=None,
    hit_tol=None,
    min_len=None,
    smoothing_factor=None,
    confidence_threshold: float = 0.5,
    batch_size: int | None = 512,
    dense_blockage: bool = False,
    with_stats: bool = False,
    _stats_accumulate: torch.Tensor | None = None,
    _profile: bool = False,
) -> TracedPaths:
    """Trace every ``(tx, rx, candidate)`` with the image method and validate it.

    Returns ``TracedPaths`` with ``vertices [Ntx,Nrx,C,k+2,3]``, ``objects [Ntx,Nrx,C,k+2]``,
    ``mask [Ntx,Nrx,C]`` and ``interaction_types [Ntx,Nrx,C,k]``, exactly the reference's fields.
    ``dense_blockage=True`` makes the blockage stage test every candidate (the amount of work the
    reference does); the default skips candidates that already failed a cheaper test — identical
    outputs.  ``batch_size`` is accepted and ignored.  ``_stats_accumulate`` (device int64[4]) and
    ``_profile`` are measurement hooks for ``bench.py``: the call's counters are added to the tensor
    on the device (no host read), and an event pair is recorded around the blockage kernel
    (``DRT_TRACE_PROFILE``).
    """
    del batch_size
    pl = Placement()
    pl.device = mesh.vertices.device
    tx = pl.put(tx_vertices, torch.float32).reshape(-1, 3).contiguous()
    rx = pl.put(rx_vertices, torch.float32).reshape(-1, 3).contiguous()
    cand = pl.put(path_candidates, torch.int32).contiguous()
    if cand.ndim != 2:
        raise TypeError("path_candidates must have shape [num_path_candidates, order]")
    dev = mesh.vertices.device
    ntx, nrx, (C, k) = tx.shape[0], rx.shape[0], cand.shape
    if k > _lib.DRT_MAX_ORDER:
        raise NotImplementedError(f"order {k} > {_lib.DRT_MAX_ORDER} is not supported")
    T = mesh.num_triangles
    out_v = torch.empty((ntx, nrx, C, k + 2, 3), dtype=torch.float32, device=dev)
    out_o = torch.empty((ntx, nrx, C, k + 2), dtype=torch.int32, device=dev)
    if interaction_types is not None:
        it = pl.put(interaction_types, torch.int32).expand(ntx, nrx, C, k)
    else:
        it = torch.zeros((1, 1, 1, 1), dtype=torch.int32, device=dev).expand(ntx, nrx, C, k)
    if smoothing_factor is not None:
        # relaxed validation (_solvers.py:599-713): float confidence in [0, 1], differentiable
        out_v, out_f, out_o = _SmoothTrace.apply(
            mesh.vertices, tx, rx, mesh.triangles, cand, mesh._mask_u8(), bool(mesh.assume_quads),
            10.0 * F32_EPS if epsilon is None else float(epsilon),
            100.0 * F32_EPS if hit_tol is None else float(hit_tol),
            10.0 * F32_EPS if min_len is None else float(min_len),
            float(smoothing_factor),
        )
        return TracedPaths(vertices=out_v, objects=out_o, mask=out_f, interaction_types=it,
                           confidence_threshold=confidence_threshold)
    out_m = torch.empty((ntx, nrx, C), dtype=torch.uint8, device=dev)
    want_stats = with_stats or _stats_accumulate is not None
    stats = torch.zeros(4, dtype=torch.int64, device=dev) if want_stats else None
    flags = (_lib.DRT_TRACE_DENSE_BLOCKAGE if dense_blockage else 0) | (_lib.DRT_TRACE_PROFILE if _profile else 0)
    ws = torch.empty(max(lib.drt_trace_workspace_bytes(T, ntx, nrx, C), 1), dtype=torch.uint8, device=dev)
    prepared = mesh._trace_prepared()  # mesh-only work, cached per mesh state: one D2D copy per call
    ws[:prepared.numel()].copy_(prepared)
    flags |= _lib.DRT_TRACE_PREPARED
    mask_u8 = mesh._mask_u8()  # named: must outlive the call
    check(
        lib.drt_trace_path_candidates(
            stream_ptr(), mesh.vertices.shape[0], T, ptr(mesh.vertices.detach()), ptr(mesh.triangles),
            ptr(mask_u8), int(mesh.assume_quads), ntx, ptr(tx.detach()), nrx, ptr(rx.detach()),
            C, k, ptr(cand),
            10.0 * F32_EPS if epsilon is None else float(epsilon),
            100.0 * F32_EPS if hit_tol is None else float(hit_tol),
            10.0 * F32_EPS if min_len is None else float(min_len),
            flags, ptr(ws), ws.numel(), ptr(out_v), ptr(out_o), ptr(out_m), ptr(stats),
        )
    )
    if torch.is_grad_enabled() and any(x.requires_grad for x in (mesh.vertices, tx, rx)):
        out_v = _TraceVertices.apply(mesh.vertices, tx, rx, mesh.triangles, cand, out_v)
    paths = TracedPaths(
        vertices=out_v,
        objects=out_o,
        mask=out_m.view(torch.bool),  # the kernels write exactly 0 / 1: reinterpret, no copy
        interaction_types=it,
        confidence_threshold=confidence_threshold,
    )
    if _stats_accumulate is not None:
        _stats_accumulate.add_(stats)
    if with_stats:
        s = stats.cpu().tolist()
        paths.stats = {"tests_done": s[0], "candidates_blockage_tested": s[1], "head_pass_survivors": s[2],
                       "ordering_pass": s[3] & 255, "culled_pass": bool(s[3] & 256)}
    return paths


def trace_valid_path_candidates(
    mesh: Mesh, tx_vertices, rx_vertices, path_candidates, *, epsilon=None, hit_tol=None, min_len=None,
    capacity: int = 1 << 16, index_offset: tuple[int, int] | None = None,
):
    """Compact form of :func:`trace_path_candidates` for exhaustive searches
    (``drt_trace_valid_path_candidates``): same validation, but only the valid paths are produced —
    no dense ``[Ntx, Nrx, C, …]`` arrays.  Returns a :class:`differt_b200.distributed.ValidPaths` in
    the reference's row-major ``masked()`` order.  ``index_offset = (num_global, start)`` maps the
    indices of a chunk of candidates into the global candidate list."""
    from .distributed import ValidPaths, global_path_index

    pl = Placement()
    dev = pl.device = mesh.vertices.device
    tx = pl.put(tx_vertices, torch.float32).reshape(-1, 3).contiguous()
    rx = pl.put(rx_vertices, torch.float32).reshape(-1, 3).contiguous()
    cand = pl.put(path_candidates, torch.int32).contiguous()
    if cand.ndim != 2:
        raise TypeError("path_candidates must have shape [num_path_candidates, order]")
    ntx, nrx, (C, k) = tx.shape[0], rx.shape[0], cand.shape
    if k > 5:
        raise NotImplementedError("the compact trace supports orders up to 5; use trace_path_candidates")
    T = mesh.num_triangles
    mask_u8 = mesh._mask_u8()
    prepared = mesh._trace_prepared()
    while True:
        out_i = torch.empty(capacity, dtype=torch.int64, device=dev)
        out_v = torch.empty((capacity, k + 2, 3), dtype=torch.float32, device=dev)
        out_o = torch.empty((capacity, k + 2), dtype=torch.int32, device=dev)
        out_ok = torch.empty(capacity, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        ws = torch.empty(max(lib.drt_trace_valid_workspace_bytes(T, capacity), 1), dtype=torch.uint8, device=dev)
        ws[:prepared.numel()].copy_(prepared)
        check(
            lib.drt_trace_valid_path_candidates(
                stream_ptr(), mesh.vertices.shape[0], T, ptr(mesh.vertices.detach()), ptr(mesh.triangles),
                ptr(mask_u8), int(mesh.assume_quads), ntx, ptr(tx), nrx, ptr(rx), C, k, ptr(cand),
                10.0 * F32_EPS if epsilon is None else float(epsilon),
                100.0 * F32_EPS if hit_tol is None else float(hit_tol),
                10.0 * F32_EPS if min_len is None else float(min_len),
                _lib.DRT_TRACE_PREPARED, capacity, ptr(ws), ws.numel(), ptr(count), ptr(out_i), ptr(out_v), ptr(out_o),
                ptr(out_ok),
            )
        )
        n = int(count.item())  # the one host read
        if n <= capacity:
            break
        capacity = n
    keep = out_ok[:n].view(torch.bool)
    index = out_i[:n][keep]
    if index_offset is not None:
        index = global_path_index(index, C, index_offset[0], index_offset[1])
    perm = torch.argsort(index, stable=True)  # slots are filled in no particular order
    return ValidPaths(index[perm], out_v[:n][keep][perm], out_o[:n][keep][perm], [int(index.numel())])


def generate_all_path_candidates(
    num_primitives: int, order: int, *, assume_quads: bool = False, start: int = 0,
    count: int | None = None, device=None,
) -> torch.Tensor:
    """All ``n (n-1)^(k-1)`` candidates of the complete graph in the reference's order
    (``_solvers.py:803-848``), decoded on the device from the linear index; under ``assume_quads``
    the indices are the even triangles (``2 *`` primitive index, ``_solvers.py:836-838``)."""
    total = num_complete_graph_candidates(num_primitives, order)
    count = total - start if count is None else max(min(count, total - start), 0)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    out = torch.empty((count, order), dtype=torch.int32, device=dev)
    check(
        lib.drt_complete_graph_candidates(
            stream_ptr(), num_primitives, order, start, count, 2 if assume_quads else 1, ptr(out)
        )
    )
    return out


def generate_all_path_candidates_chunks_iter(
    num_primitives: int, order: int, chunk_size: int = 1000, *, assume_quads: bool = False
) -> Iterator[torch.Tensor]:
    """Chunked variant (reference ``_solvers.py:850-934``)."""
    total = num_complete_graph_candidates(num_primitives, order)
    for start in range(0, total, chunk_size):
        yield generate_all_path_candidates(
            num_primitives, order, assume_quads=assume_quads, start=start, count=chunk_size
        )


class VisiblePathCandidates:
    """Candidates of ``HybridPathTracer`` (reference ``_solvers.py:993-1058``): every tuple of
    primitives without consecutive repeats whose first element is visible from a transmitter, whose
    last element is visible from a receiver and whose elements are all active, in the reference's
    (lexicographic DFS, ``graph.rs:1063-1108``) order.  The completion counts live on the device
    (``drt_digraph_candidates_prepare``); ``chunk(start, count)`` decodes any slice independently."""

    def __init__(self, num_primitives: int, order: int, visible_from_tx, visible_from_rx, active=None,
                 *, assume_quads: bool = False, device=None) -> None:
        if num_primitives ** max(order, 1) >= 2 ** 62:
            raise OverflowError("num_primitives ** order must stay below 2**62")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        u8 = lambda m: None if m is None else torch.as_tensor(m).to(dev, torch.uint8).contiguous()
        self.n, self.order, self.mult, self.device = int(num_primitives), int(order), 2 if assume_quads else 1, dev
        self._ws = torch.empty(max(lib.drt_digraph_candidates_workspace_bytes(self.n, self.order), 1),
                               dtype=torch.uint8, device=dev)
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        # keep the three masks alive across the call: temporaries would share one allocator block
        m_from, m_to, m_active = u8(visible_from_tx), u8(visible_from_rx), u8(active)
        check(
            lib.drt_digraph_candidates_prepare(
                stream_ptr(), self.n, self.order, ptr(m_from), ptr(m_to), ptr(m_active), ptr(self._ws),
                self._ws.numel(), ptr(total),
            )
        )
        self.total = int(total.item())  # the one host read: sizes the candidate arrays

    def __len__(self) -> int:
        return self.total

    def chunk(self, start: int = 0, count: int | None = None) -> torch.Tensor:
        count = self.total - start if count is None else max(min(count, self.total - start), 0)
        # zero candidates come back as [0, 0] like the reference's empty iterator (graph.rs:40-54)
        out = torch.empty((count, self.order if self.total > 0 else 0), dtype=torch.int32, device=self.device)
        if count > 0 and self.order > 0:
            check(
                lib.drt_digraph_candidates(
                    stream_ptr(), self.n, self.order, ptr(self._ws), start, count, self.mult, ptr(out)
                )
            )
        return out

    def chunks_iter(self, chunk_size: int) -> Iterator[torch.Tensor]:
        for start in range(0, self.total, chunk_size):
            yield self.chunk(start, chunk_size)


def generate_visible_path_candidates(
    mesh: Mesh, tx_vertices, rx_vertices, order: int, *, num_rays: int = 1_000_000, accel: str = "brute"
) -> VisiblePathCandidates:
    """``HybridPathTracer.generate_path_candidates`` (reference ``_solvers.py:993-1058``): visibility
    of every triangle from the transmitters and from the receivers (K4), merged over quads and over
    the tx / rx batches, then the pruned candidate graph decoded on the device."""
    pl = Placement()
    pl.device = mesh.vertices.device
    tx = pl.put(tx_vertices, torch.float32).reshape(-1, 3)
    rx = pl.put(rx_vertices, torch.float32).reshape(-1, 3)
    vis_tx = mesh.triangles_visible_from_vertex(tx, num_rays=num_rays, accel=accel).any(dim=0)
    vis_rx = mesh.triangles_visible_from_vertex(rx, num_rays=num_rays, accel=accel).any(dim=0)
    active = mesh.mask
    if mesh.assume_quads:
        vis_tx = vis_tx.reshape(-1, 2).any(dim=-1)
        vis_rx = vis_rx.reshape(-1, 2).any(dim=-1)
        if active is not None:
            active = active[0::2] & active[1::2]
    return VisiblePathCandidates(
        mesh.num_primitives, order, vis_tx, vis_rx, active, assume_quads=mesh.assume_quads,
        device=mesh.vertices.device,
    )


def _candidate_chunks(mesh: Mesh, tx_vertices, rx_vertices, order: int, chunk_size: int, solver: str,
                      num_rays: int, accel: str):
    """``(total, iterator of (start, candidates))`` for the exhaustive or the hybrid candidate set."""
    if solver == "exhaustive":
        total = num_complete_graph_candidates(mesh.num_primitives, order)
        gen = lambda start: generate_all_path_candidates(  # noqa: E731
            mesh.num_primitives, order, assume_quads=mesh.assume_quads, start=start, count=chunk_size,
            device=mesh.vertices.device)
    elif solver == "hybrid":
        vis = generate_visible_path_candidates(mesh, tx_vertices, rx_vertices, order, num_rays=num_rays, accel=accel)
        total = len(vis)
        gen = lambda start: vis.chunk(start, chunk_size)  # noqa: E731
    else:
        raise ValueError(f"Unknown solver: {solver}")
    return total, ((start, gen(start)) for start in range(0, total, chunk_size))


def trace_paths_chunks_iter(mesh: Mesh, tx_vertices, rx_vertices, order: int, *, chunk_size: int,
                            solver: str = "exhaustive", num_rays: int = 1_000_000, accel: str = "brute",
                            **kwargs) -> Iterator[TracedPaths]:
    """``Scene.trace_paths(order, chunk_size=...)`` (reference ``_scene.py:738-751``): one dense
    ``TracedPaths`` per chunk of candidates; the candidates of every chunk are decoded on the device."""
    _, chunks = _candidate_chunks(mesh, tx_vertices, rx_vertices, order, chunk_size, solver, num_rays, accel)
    for _, cand in chunks:
        yield trace_path_candidates(mesh, tx_vertices, rx_vertices, cand, **kwargs)


def trace_valid_paths(mesh: Mesh, tx_vertices, rx_vertices, order: int, *, chunk_size: int = 1 << 20,
                      solver: str = "exhaustive", num_rays: int = 1_000_000, accel: str = "brute",
                      **kwargs):
    """Every valid path of ``order`` — what ``Scene.trace_paths(order).masked()`` returns — without
    the dense arrays: chunks of candidates are decoded, traced and validated on the device by the
    compact kernel (``trace_valid_path_candidates``; the dense kernel + compaction for orders > 5) and
    the survivors merged back into the reference's row-major ``(tx, rx, candidate)`` order.
    ``num_tx * num_rx * chunk_size`` must stay below 2**32.
    Returns a :class:`differt_b200.distributed.ValidPaths`."""
    from .distributed import GatherRecord, ValidPaths, fill_record

    total, chunks = _candidate_chunks(mesh, tx_vertices, rx_vertices, order, chunk_size, solver, num_rays, accel)
    idx, verts, objs = [], [], []
    for start, cand in chunks:
        if order <= 5 and not kwargs.get("dense_blockage", False):
            part = trace_valid_path_candidates(mesh, tx_vertices, rx_vertices, cand, index_offset=(total, start),
                                               **{k: v for k, v in kwargs.items() if k in ("epsilon", "hit_tol", "min_len")})
            idx.append(part.index), verts.append(part.vertices), objs.append(part.objects)
            continue
        paths = trace_path_candidates(mesh, tx_vertices, rx_vertices, cand, **kwargs)
        capacity = 1 << 12
        while True:
            record = GatherRecord(capacity, paths.order, paths.vertices.device)
            fill_record(record, paths, total, start)
            count, index, v, o = record.fields()
            n = int(count.item())
            if n <= capacity:
                break
            capacity = n
        idx.append(index[:n].clone()), verts.append(v[:n].clone()), objs.append(o[:n].clone())
    dev = mesh.vertices.device
    if not idx:
        return ValidPaths(torch.zeros(0, dtype=torch.int64, device=dev), torch.zeros((0, order + 2, 3), device=dev),
                          torch.zeros((0, order + 2), dtype=torch.int32, device=dev), [0])
    index, vertices, objects = torch.cat(idx), torch.cat(verts), torch.cat(objs)
    perm = torch.argsort(index, stable=True)
    return ValidPaths(index[perm], vertices[perm], objects[perm], [int(index.numel())])


def trace_paths(mesh: Mesh, tx_vertices, rx_vertices, order: int, *, solver: str = "exhaustive",
                num_rays: int = 1_000_000, **kwargs) -> TracedPaths:
    """``Scene.trace_paths(order, solver=...)`` (reference ``_scene.py:650-764``) for one mesh:
    ``"exhaustive"`` (every candidate of the complete graph) or ``"hybrid"`` (visibility-pruned)."""
    if solver == "exhaustive":
        cand = generate_all_path_candidates(
            mesh.num_primitives, order, assume_quads=mesh.assume_quads, device=mesh.vertices.device
        )
    elif solver == "hybrid":
        cand = generate_visible_path_candidates(mesh, tx_vertices, rx_vertices, order, num_rays=num_rays).chunk()
    else:
        raise ValueError(f"Unknown solver: {solver}")  # _scene.py:700-702
    return trace_path_candidates(mesh, tx_vertices, rx_vertices, cand, **kwargs)
