"""Public functions of the hot path, with the reference's names, argument meaning and defaults.

Mirror of ``differt.geometry`` for the functions BASELINE.json names (reference:
``differt/src/differt/geometry/_utils.py:1157-1960`` and ``_solver_image_method.py:11-454``; the
keyword defaults are those listed in SURVEY.md Appendix D).  Arguments are ``torch`` tensors (CUDA or
host) or anything ``numpy.asarray`` accepts; results live where the inputs lived.  All arithmetic is
done by the CUDA library behind the C ABI (``include/differt_b200.h``) — there is no fallback.

Differentiation surface (the reference's ``custom_vjp`` surface): ``image_method`` and
the hit distance ``t`` of ``first_triangle_hit_by_ray`` / ``ray_intersect_triangle`` carry gradients
(``torch.autograd``); boolean / index outputs never do.  ``smoothing_factor`` is supported by
``ray_intersect_triangle``, ``ray_intersect_any_triangle`` (float outputs, differentiable),
``consecutive_vertices_are_on_same_side_of_mirror`` (a function of signs: zero gradient, like the
reference's) and the trace (``solvers.trace_path_candidates(smoothing_factor=...)``, differentiable
end to end).
"""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch

from . import _lib
from ._lib import check, lib
from ._tensor import F32_EPS, Placement, batch_strides, numel, ptr, stream_ptr, sum_to_shape

__all__ = [
    "assemble_path",
    "cartesian_to_spherical",
    "consecutive_vertices_are_on_same_side_of_mirror",
    "fibonacci_lattice",
    "first_triangle_hit_by_ray",
    "image_method",
    "image_of_vertex_with_respect_to_mirror",
    "intersection_of_ray_with_plane",
    "normalize",
    "pack_mesh",
    "pack_triangle_vertices",
    "ray_intersect_any_triangle",
    "ray_intersect_triangle",
    "spherical_to_cartesian",
    "triangles_visible_from_vertex",
    "viewing_frustum",
]


# below this many rays an any-hit call is launch-bound and sorting the pack first does not pay
_SORT_MIN_RAYS = 4096
_CULL_MIN_TRIANGLES = 2048  # above this the flat queries can go through the exact cull (csrc/cull.cuh)
_CULL_MIN_WORK = 1 << 30    # ... when rays x triangles is large enough to pay for building its hierarchy
#                             (~0.3 ms: 10 000 rays x 14 206 triangles, the reference's harness, stay all-pairs)


def use_cull(num_rays: int, num_triangles: int) -> bool:
    return num_triangles > _CULL_MIN_TRIANGLES and num_rays * num_triangles >= _CULL_MIN_WORK


def _default(value, factor: float) -> float:
    return factor * F32_EPS if value is None else float(value)


# ------------------------------------------------------------------------------------------------
# packed meshes
# ------------------------------------------------------------------------------------------------


def pack_mesh(vertices: torch.Tensor, triangles: torch.Tensor, mask: torch.Tensor | None = None):
    """Pack ``Mesh.vertices`` / ``Mesh.triangles`` (CUDA tensors) for the intersection kernels."""
    T = int(triangles.shape[0])
    pack = torch.empty(lib.drt_mesh_pack_bytes(T), dtype=torch.uint8, device=vertices.device)
    check(
        lib.drt_mesh_pack(
            stream_ptr(), int(vertices.shape[0]), T, ptr(vertices), ptr(triangles), ptr(mask), ptr(pack)
        )
    )
    return pack


def pack_triangle_vertices(triangle_vertices: torch.Tensor, mask: torch.Tensor | None = None):
    T = int(triangle_vertices.shape[0])
    pack = torch.empty(lib.drt_mesh_pack_bytes(T), dtype=torch.uint8, device=triangle_vertices.device)
    check(
        lib.drt_mesh_pack_triangle_vertices(
            stream_ptr(), T, ptr(triangle_vertices), ptr(mask), ptr(pack)
        )
    )
    return pack


def sort_pack_by_area(pack: torch.Tensor, num_triangles: int) -> torch.Tensor:
    """Any-hit ordering of a pack (largest triangles first, ``drt_mesh_pack_sort_by_area``): the
    result of an any-hit query does not depend on the order, blocked rays just stop sooner."""
    out = torch.empty_like(pack)
    ws = torch.empty(max(lib.drt_mesh_pack_sort_workspace_bytes(num_triangles), 1), dtype=torch.uint8,
                     device=pack.device)
    check(lib.drt_mesh_pack_sort_by_area(stream_ptr(), num_triangles, ptr(pack), ptr(ws), ws.numel(), ptr(out)))
    return out


def first_hit_launch(pack: torch.Tensor, T: int, o: torch.Tensor, d: torch.Tensor, eps: float, batch_size: int,
                     idx: torch.Tensor, t: torch.Tensor) -> None:
    """Nearest hit of ``R`` flat rays against a packed mesh (in the mesh's own triangle order) into
    ``idx`` / ``t``: behind the exact conservative cull (``csrc/cull.cuh``, identical results) when the
    mesh and the batch are large enough to pay for its hierarchy, else the all-pairs engine."""
    R = int(o.shape[0])
    if use_cull(R, T):
        ws = torch.empty(lib.drt_any_hit_workspace_bytes(T), dtype=torch.uint8, device=o.device)
        check(lib.drt_first_triangle_hit_by_ray_culled(stream_ptr(), R, ptr(o), ptr(d), ptr(pack), T, eps, batch_size,
                                                       ptr(ws), ws.numel(), ptr(idx), ptr(t), None))
    else:
        check(lib.drt_first_triangle_hit_by_ray(stream_ptr(), R, ptr(o), ptr(d), ptr(pack), T, eps, batch_size,
                                                ptr(idx), ptr(t), None))


def pack_normals(pack: torch.Tensor, num_triangles: int) -> torch.Tensor:
    """Unit normals stored in the pack (``Mesh.normals``, reference ``_mesh.py:950-956``)."""
    return pack.view(torch.float32).view(-1, 12)[:num_triangles, 9:12]


# ------------------------------------------------------------------------------------------------
# small helpers of the reference API (a10)
# ------------------------------------------------------------------------------------------------


def normalize(vectors, keepdims: bool = False):
    """Reference ``_utils.py:29-72``: unit vectors and lengths; zero vectors stay zero."""
    pl = Placement()
    v = pl.put(vectors, torch.float32)
    lengths = torch.linalg.norm(v, dim=-1, keepdim=True)
    safe = torch.where(lengths == 0.0, torch.ones_like(lengths), lengths)
    return pl.out(v / safe), pl.out(lengths if keepdims else lengths.squeeze(-1))


def path_length(path):
    """Reference ``_utils.py:150-182``: sum of the segment lengths of ``[*batch, path_length, 3]``
    (host-side helper on ``torch`` ops: differentiable, the usual ``fun`` of ``TracedPaths.reduce``)."""
    pl = Placement()
    p = pl.put(path, torch.float32)
    return pl.out(torch.linalg.norm(p[..., 1:, :] - p[..., :-1, :], dim=-1).sum(dim=-1))


def assemble_path(from_vertex, intermediate_vertices, to_vertex=None):
    """Reference ``_utils.py:514-565``: concatenate ``[from, *intermediate, to]`` on axis -2."""
    pl = Placement()
    f = pl.put(from_vertex, torch.float32)
    m = pl.put(intermediate_vertices, torch.float32)
    if to_vertex is None:
        batch = torch.broadcast_shapes(f.shape[:-1], m.shape[:-1])
        return pl.out(
            torch.cat((f.expand(*batch, 3).unsqueeze(-2), m.expand(*batch, 3).unsqueeze(-2)), dim=-2)
        )
    t = pl.put(to_vertex, torch.float32)
    batch = torch.broadcast_shapes(f.shape[:-1], m.shape[:-2], t.shape[:-1])
    return pl.out(
        torch.cat(
            (
                f.expand(*batch, 3).unsqueeze(-2),
                m.expand(*batch, *m.shape[-2:]),
                t.expand(*batch, 3).unsqueeze(-2),
            ),
            dim=-2,
        )
    )


# ------------------------------------------------------------------------------------------------
# K1: ray_intersect_triangle
# ------------------------------------------------------------------------------------------------


class _FirstHitDistanceGrad(torch.autograd.Function):
    """Attach the reference's hit-distance VJP (``_mesh.py:226-255, 308-338``) to a computed ``t``."""

    @staticmethod
    def forward(ctx, t, vertices, triangles, origins, directions, faces):
        ctx.save_for_backward(vertices, triangles, origins, directions, faces)
        return t.clone()

    @staticmethod
    def backward(ctx, g_t):
        vertices, triangles, origins, directions, faces = ctx.saved_tensors
        R, V, T = origins.shape[0], vertices.shape[0], triangles.shape[0]
        g_t = g_t.contiguous().to(torch.float32)
        g_v = torch.empty_like(vertices)
        g_o = torch.empty_like(origins)
        g_d = torch.empty_like(directions)
        check(
            lib.drt_first_triangle_hit_by_ray_vjp(
                stream_ptr(), R, V, T, ptr(vertices), ptr(triangles), ptr(origins), ptr(directions),
                ptr(faces), ptr(g_t), ptr(g_v), ptr(g_o), ptr(g_d),
            )
        )
        return None, g_v, None, g_o, g_d, None


class _SmoothTriangle(torch.autograd.Function):
    """Relaxed Möller–Trumbore on flat ``[n, …]`` inputs with its reverse mode: both ``t`` and the
    relaxed ``hit`` carry gradients (``jax.grad`` of ``_utils.py:1263-1322`` with smoothing)."""

    @staticmethod
    def forward(ctx, o, d, tv, eps, alpha):
        n = o.shape[0]
        t = torch.empty(n, dtype=torch.float32, device=o.device)
        hit = torch.empty(n, dtype=torch.float32, device=o.device)
        ndim, shape, (so, sd, st), keep = batch_strides((n,), [(o, 1), (d, 1), (tv, 2)])
        check(
            lib.drt_ray_intersect_triangle_smooth(
                stream_ptr(), ndim, shape, ptr(keep[0]), so, ptr(keep[1]), sd, ptr(keep[2]), st, eps, alpha,
                ptr(t), ptr(hit),
            )
        )
        ctx.save_for_backward(o, d, tv)
        ctx.params = (eps, alpha)
        ctx.set_materialize_grads(False)
        return t, hit

    @staticmethod
    def backward(ctx, g_t, g_hit):
        o, d, tv = ctx.saved_tensors
        eps, alpha = ctx.params
        g_o, g_d, g_tv = torch.zeros_like(o), torch.zeros_like(d), torch.zeros_like(tv)
        if g_t is not None or g_hit is not None:
            g_t = None if g_t is None else g_t.contiguous().to(torch.float32)
            g_hit = None if g_hit is None else g_hit.contiguous().to(torch.float32)
            check(
                lib.drt_ray_intersect_triangle_smooth_vjp(
                    stream_ptr(), o.shape[0], ptr(o), ptr(d), ptr(tv), eps, alpha, ptr(g_t), ptr(g_hit),
                    ptr(g_o), ptr(g_d), ptr(g_tv),
                )
            )
        return g_o, g_d, g_tv, None, None


class _SmoothAnyTriangle(torch.autograd.Function):
    """Relaxed any-hit of flat rays against one mesh, with its reverse mode
    (``jax.grad`` of ``_utils.py:1452-1476``)."""

    @staticmethod
    def forward(ctx, o, d, tv, act, eps, tol, alpha):
        T = tv.shape[0]
        pack = pack_triangle_vertices(tv, act)
        res = torch.empty(o.shape[0], dtype=torch.float32, device=o.device)
        check(
            lib.drt_ray_intersect_any_triangle_smooth(
                stream_ptr(), o.shape[0], ptr(o), ptr(d), ptr(pack), T, eps, tol, alpha, ptr(res)
            )
        )
        ctx.save_for_backward(o, d, tv, act)
        ctx.params = (eps, tol, alpha)
        return res

    @staticmethod
    def backward(ctx, g):
        o, d, tv, act = ctx.saved_tensors
        eps, tol, alpha = ctx.params
        pack = pack_triangle_vertices(tv, act)
        g = g.contiguous().to(torch.float32)
        g_o, g_d, g_tv = torch.empty_like(o), torch.empty_like(d), torch.empty_like(tv)
        check(
            lib.drt_ray_intersect_any_triangle_smooth_vjp(
                stream_ptr(), o.shape[0], ptr(o), ptr(d), ptr(pack), tv.shape[0], eps, tol, alpha, ptr(g),
                ptr(g_o), ptr(g_d), ptr(g_tv),
            )
        )
        return g_o, g_d, g_tv, None, None, None, None


def ray_intersect_triangle(
    ray_origins, ray_directions, triangle_vertices, *, epsilon=None, smoothing_factor=None
):
    """Möller–Trumbore ``(t, hit)`` — reference ``_utils.py:1157-1322``.

    ``t`` is returned even where ``hit`` is false; ``epsilon`` defaults to ``10 * eps(float32)``.
    With ``smoothing_factor`` the hit is a float in [0, 1] (sigmoid relaxation, reference
    ``_utils.py:1279-1318``) and both outputs are differentiable.
    """
    pl = Placement()
    o = pl.put(ray_origins, torch.float32)
    d = pl.put(ray_directions, torch.float32)
    tv = pl.put(triangle_vertices, torch.float32)
    batch = torch.broadcast_shapes(o.shape[:-1], d.shape[:-1], tv.shape[:-2])
    n = numel(batch)
    t = torch.empty(batch, dtype=torch.float32, device=o.device)
    if smoothing_factor is not None:
        hit_f = torch.empty(batch, dtype=torch.float32, device=o.device)
        if n > 0 and torch.is_grad_enabled() and any(x.requires_grad for x in (o, d, tv)):
            # autograd of `expand` sums the per-element cotangents over the broadcast axes
            t, hit_f = _SmoothTriangle.apply(
                o.expand(*batch, 3).reshape(n, 3).contiguous(), d.expand(*batch, 3).reshape(n, 3).contiguous(),
                tv.expand(*batch, 3, 3).reshape(n, 3, 3).contiguous(), _default(epsilon, 10.0), float(smoothing_factor),
            )
            return pl.out(t.reshape(batch)), pl.out(hit_f.reshape(batch))
        if n > 0:
            ndim, shape, (so, sd, st), keep = batch_strides(batch, [(o.detach(), 1), (d.detach(), 1), (tv.detach(), 2)])
            check(
                lib.drt_ray_intersect_triangle_smooth(
                    stream_ptr(), ndim, shape, ptr(keep[0]), so, ptr(keep[1]), sd, ptr(keep[2]), st,
                    _default(epsilon, 10.0), float(smoothing_factor), ptr(t), ptr(hit_f),
                )
            )
        return pl.out(t), pl.out(hit_f)
    hit = torch.empty(batch, dtype=torch.uint8, device=o.device)
    needs_grad = torch.is_grad_enabled() and any(x.requires_grad for x in (o, d, tv))
    if n > 0:
        ndim, shape, (so, sd, st), keep = batch_strides(batch, [(o.detach(), 1), (d.detach(), 1), (tv.detach(), 2)])
        check(
            lib.drt_ray_intersect_triangle(
                stream_ptr(), ndim, shape, ptr(keep[0]), so, ptr(keep[1]), sd, ptr(keep[2]), st,
                _default(epsilon, 10.0), ptr(t), ptr(hit),
            )
        )
    if needs_grad and n > 0:
        oe = o.expand(*batch, 3).reshape(n, 3)
        de = d.expand(*batch, 3).reshape(n, 3)
        ve = tv.expand(*batch, 3, 3).reshape(3 * n, 3)
        tris = torch.arange(3 * n, dtype=torch.int32, device=o.device).view(n, 3)
        faces = torch.arange(n, dtype=torch.int32, device=o.device)
        t = _FirstHitDistanceGrad.apply(
            t.reshape(n), ve.contiguous(), tris, oe.contiguous(), de.contiguous(), faces
        ).reshape(batch)
    return pl.out(t), pl.out(hit.view(torch.bool))


# ------------------------------------------------------------------------------------------------
# K2 / K3 / K4 over a (possibly batched) set of meshes
# ------------------------------------------------------------------------------------------------


def _mesh_batches(batch, tv, active):
    """Yield ``(index tuple into the broadcast batch, tv [T,3,3], active [T] | None)`` per mesh."""
    mesh_batch = torch.broadcast_shapes(tv.shape[:-3], () if active is None else active.shape[:-1])
    if numel(mesh_batch) == 1:
        yield (
            (Ellipsis,),
            tv.reshape(-1, 3, 3),
            None if active is None else active.reshape(-1),
        )
        return
    lead = len(batch) - len(mesh_batch)
    tvb = tv.expand(*mesh_batch, *tv.shape[-3:])
    ab = None if active is None else active.expand(*mesh_batch, active.shape[-1])
    for idx in np.ndindex(*mesh_batch):
        sel = tuple(slice(None) for _ in range(lead)) + tuple(
            # a mesh-batch axis of size 1 broadcasts against the whole batch axis
            slice(i, i + 1) if mesh_batch[a] != 1 else slice(None) for a, i in enumerate(idx)
        )
        yield sel, tvb[idx], None if ab is None else ab[idx]


def ray_intersect_any_triangle(
    ray_origins,
    ray_directions,
    triangle_vertices,
    active_triangles=None,
    *,
    hit_tol=None,
    smoothing_factor=None,
    batch_size: int | None = 512,
    **kwargs: Any,
):
    """``any_j[(t < 1 - hit_tol) & hit & active_j]`` — reference ``_utils.py:1353-1537``.

    ``batch_size`` is accepted for signature compatibility and ignored (it only bounds the
    reference's memory use); ``epsilon`` may be passed through ``**kwargs`` like in the reference.
    """
    del batch_size
    _tests_done = kwargs.pop("_tests_done", None)  # measurement hook: device int64 counter of executed tests
    pl = Placement()
    o = pl.put(ray_origins, torch.float32)
    d = pl.put(ray_directions, torch.float32)
    tv = pl.put(triangle_vertices, torch.float32)
    act = None if active_triangles is None else pl.put(active_triangles, torch.uint8)
    batch = torch.broadcast_shapes(
        o.shape[:-1], d.shape[:-1], tv.shape[:-3], () if act is None else act.shape[:-1]
    )
    T = int(tv.shape[-3])
    if smoothing_factor is not None:  # sigmoid relaxation (_utils.py:1465-1476) → float
        outf = torch.zeros(batch, dtype=torch.float32, device=o.device)
        if T == 0 or numel(batch) == 0:
            return pl.out(outf)
        ob, db = o.expand(*batch, 3), d.expand(*batch, 3)
        eps, tol = _default(kwargs.get("epsilon"), 10.0), _default(hit_tol, 100.0)
        if torch.is_grad_enabled() and any(x.requires_grad for x in (o, d, tv)):
            mesh_batch = torch.broadcast_shapes(tv.shape[:-3], () if act is None else act.shape[:-1])
            if numel(mesh_batch) != 1:
                raise NotImplementedError("gradient of the relaxed any-hit over a BATCH of meshes is not built")
            res = _SmoothAnyTriangle.apply(
                ob.reshape(-1, 3).contiguous(), db.reshape(-1, 3).contiguous(), tv.reshape(-1, 3, 3).contiguous(),
                None if act is None else act.reshape(-1).contiguous(), eps, tol, float(smoothing_factor),
            )
            return pl.out(res.reshape(batch))
        for sel, tvi, acti in _mesh_batches(batch, tv, act):
            oi, di = ob[sel].reshape(-1, 3).contiguous(), db[sel].reshape(-1, 3).contiguous()
            pack = pack_triangle_vertices(tvi.contiguous(), None if acti is None else acti.contiguous())
            res = torch.empty(oi.shape[0], dtype=torch.float32, device=o.device)
            check(
                lib.drt_ray_intersect_any_triangle_smooth(
                    stream_ptr(), oi.shape[0], ptr(oi), ptr(di), ptr(pack), T, eps, tol, float(smoothing_factor), ptr(res)
                )
            )
            outf[sel] = res.view(outf[sel].shape)
        return pl.out(outf)
    out = torch.zeros(batch, dtype=torch.uint8, device=o.device)
    if T == 0 or numel(batch) == 0:
        return pl.out(out.view(torch.bool))
    ob, db = o.expand(*batch, 3), d.expand(*batch, 3)
    eps, tol = _default(kwargs.get("epsilon"), 10.0), _default(hit_tol, 100.0)
    for sel, tvi, acti in _mesh_batches(batch, tv, act):
        oi, di = ob[sel].reshape(-1, 3).contiguous(), db[sel].reshape(-1, 3).contiguous()
        pack = pack_triangle_vertices(tvi.contiguous(), None if acti is None else acti.contiguous())
        if oi.shape[0] >= _SORT_MIN_RAYS:
            pack = sort_pack_by_area(pack, T)
        res = torch.empty(oi.shape[0], dtype=torch.uint8, device=o.device)
        if use_cull(oi.shape[0], T):
            # same test, same results, behind the exact conservative cull (csrc/cull.cuh): O(log T) per ray
            ws = torch.empty(lib.drt_any_hit_workspace_bytes(T), dtype=torch.uint8, device=o.device)
            check(
                lib.drt_ray_intersect_any_triangle_culled(
                    stream_ptr(), oi.shape[0], ptr(oi), ptr(di), ptr(pack), T, eps, tol, ptr(ws), ws.numel(),
                    ptr(res), ptr(_tests_done),
                )
            )
        else:
            check(
                lib.drt_ray_intersect_any_triangle(
                    stream_ptr(), oi.shape[0], ptr(oi), ptr(di), ptr(pack), T, eps, tol, ptr(res), ptr(_tests_done)
                )
            )
        out[sel] = res.view(out[sel].shape)
    return pl.out(out.view(torch.bool))


def first_triangle_hit_by_ray(
    ray_origins,
    ray_directions,
    triangle_vertices,
    active_triangles=None,
    batch_size: int | None = 512,
    **kwargs: Any,
):
    """``(index, t)`` of the nearest hit, ``(-1, inf)`` on a miss — reference ``_utils.py:1775-1960``.

    ``batch_size`` only matters on exactly equal distances, where it reproduces the reference's tie
    rule (lowest index inside a batch of triangles, latest batch across batches).
    """
    pl = Placement()
    o = pl.put(ray_origins, torch.float32)
    d = pl.put(ray_directions, torch.float32)
    tv = pl.put(triangle_vertices, torch.float32)
    act = None if active_triangles is None else pl.put(active_triangles, torch.uint8)
    batch = torch.broadcast_shapes(
        o.shape[:-1], d.shape[:-1], tv.shape[:-3], () if act is None else act.shape[:-1]
    )
    idx = torch.full(batch, -1, dtype=torch.int32, device=o.device)
    t = torch.full(batch, math.inf, dtype=torch.float32, device=o.device)
    T = int(tv.shape[-3])
    if T == 0 or numel(batch) == 0:
        return pl.out(idx), pl.out(t)
    eps = _default(kwargs.get("epsilon"), 10.0)
    bs = 0 if batch_size is None else int(batch_size)
    ob, db = o.expand(*batch, 3), d.expand(*batch, 3)
    needs_grad = torch.is_grad_enabled() and any(x.requires_grad for x in (o, d, tv))
    single = numel(torch.broadcast_shapes(tv.shape[:-3], () if act is None else act.shape[:-1])) == 1
    if needs_grad and not single:
        # the reference's t is differentiable in every case; dropping the gradient silently is worse
        # than refusing (the relaxed any-hit does the same)
        raise NotImplementedError("gradient of first_triangle_hit_by_ray over a BATCH of meshes is not built; "
                                  "loop over the meshes or detach the inputs")
    for sel, tvi, acti in _mesh_batches(batch, tv, act):
        oi, di = ob[sel].reshape(-1, 3).contiguous(), db[sel].reshape(-1, 3).contiguous()
        tvc = tvi.contiguous()
        pack = pack_triangle_vertices(tvc.detach(), None if acti is None else acti.contiguous())
        ii = torch.empty(oi.shape[0], dtype=torch.int32, device=o.device)
        ti = torch.empty(oi.shape[0], dtype=torch.float32, device=o.device)
        first_hit_launch(pack, T, oi, di, eps, bs, ii, ti)
        if needs_grad and single:
            tris = torch.arange(3 * T, dtype=torch.int32, device=o.device).view(T, 3)
            ti = _FirstHitDistanceGrad.apply(ti, tvc.reshape(3 * T, 3), tris, oi, di, ii)
            idx = ii.view(batch)
            t = ti.view(batch)
        else:
            idx[sel] = ii.view(idx[sel].shape)
            t[sel] = ti.view(t[sel].shape)
    return pl.out(idx), pl.out(t)


# ------------------------------------------------------------------------------------------------
# ray generation for the visibility query (host-side in the reference too: _mesh.py:3216-3228)
# ------------------------------------------------------------------------------------------------


def cartesian_to_spherical(xyz):
    """Reference ``_utils.py:930-958`` → ``(r, polar, azimuth)``."""
    pl = Placement()
    v = pl.put(xyz, torch.float32)
    r = torch.linalg.norm(v, dim=-1)
    r = torch.where(r == 0.0, torch.ones_like(r), r)
    p = torch.acos(v[..., 2] / r)
    a = torch.atan2(v[..., 1], v[..., 0])
    return pl.out(torch.stack((r, p, a), dim=-1))


def spherical_to_cartesian(rpa):
    """Reference ``_utils.py:961-993`` (radius optional)."""
    pl = Placement()
    v = pl.put(rpa, torch.float32)
    p, a = v[..., -2], v[..., -1]
    sp = torch.sin(p)
    xyz = torch.stack((sp * torch.cos(a), sp * torch.sin(a), torch.cos(p)), dim=-1)
    if v.shape[-1] == 3:
        xyz = xyz * v[..., 0, None]
    return pl.out(xyz)


def viewing_frustum(viewing_vertex, world_vertices, active_vertices=None):
    """Reference ``_utils.py:639-927`` (``reduce=False`` form) → ``[*batch, 2, 3]``."""
    pl = Placement()
    vv = pl.put(viewing_vertex, torch.float32)
    wv = pl.put(world_vertices, torch.float32)
    av = None if active_vertices is None else pl.put(active_vertices, torch.bool)
    rpa = cartesian_to_spherical(wv - vv[..., None, :])
    r, p, a = rpa[..., 0], rpa[..., 1], rpa[..., 2]
    pi, two_pi = math.pi, 2.0 * math.pi

    def rmin(x, init):
        x = x if av is None else torch.where(av, x, torch.full_like(x, init))
        return torch.clamp(x.amin(dim=-1), max=init)

    def rmax(x, init):
        x = x if av is None else torch.where(av, x, torch.full_like(x, init))
        return torch.clamp(x.amax(dim=-1), min=init)

    r_min, r_max = rmin(r, math.inf), rmax(r, 0.0)
    p_min, p_max = rmin(p, float(np.float32(pi))), rmax(p, 0.0)
    a_min, a_max = rmin(a, float(np.float32(pi))), rmax(a, -float(np.float32(pi)))
    a0 = torch.remainder(a + np.float32(two_pi), np.float32(two_pi))
    a0_min, a0_max = rmin(a0, float(np.float32(two_pi))), rmax(a0, 0.0)
    a_width, a0_width = a_max - a_min, a0_max - a0_min
    swap = a_width > a0_width
    a_min, a_max = torch.where(swap, a0_min, a_min), torch.where(swap, a0_max, a_max)
    full = torch.minimum(a_width, a0_width) > float(np.float32(1.5) * np.float32(pi))
    a_min = torch.where(full, torch.full_like(a_min, -float(np.float32(pi))), a_min)
    a_max = torch.where(full, torch.full_like(a_max, float(np.float32(pi))), a_max)
    p0_min, p0_max = p_min, p_max
    degenerate = p_min == p_max
    p_min = torch.where(degenerate, torch.zeros_like(p_min), p_min)
    p0_max = torch.where(degenerate, torch.full_like(p0_max, float(np.float32(pi))), p0_max)
    wider = (p_max - p_min) > (p0_max - p0_min)
    p_min, p_max = torch.where(wider, p0_min, p_min), torch.where(wider, p0_max, p_max)
    lo = torch.stack((r_min, p_min, a_min), dim=-1)
    hi = torch.stack((r_max, p_max, a_max), dim=-1)
    return pl.out(torch.stack((lo, hi), dim=-2))


def fibonacci_lattice(n: int, dtype=None, *, frustum=None):
    """Reference ``_utils.py:351-490`` → ``[*frustum batch, n, 3]`` unit directions."""
    if n <= 0:
        raise ValueError(f"Invalid size {n!r}, must be strictly positive.")
    del dtype
    pl = Placement()
    fr = None if frustum is None else pl.put(frustum, torch.float32)
    device = fr.device if fr is not None else pl.put(np.zeros(1, np.float32), torch.float32).device
    if fr is None:
        pl.saw_cuda = True
    i = torch.arange(0, n, dtype=torch.float32, device=device)
    inv_phi = 0.6180339887498949
    m1, m2 = 262144.0, 512.0
    inv_phi_m1 = float(np.float32((inv_phi * m1) % 1.0))
    inv_phi_m2 = float(np.float32((inv_phi * m2) % 1.0))
    q1 = torch.floor(i / m1)
    rem = i - q1 * m1
    q2 = torch.floor(rem / m2)
    r = rem - q2 * m2
    frac = torch.remainder(q1 * inv_phi_m1 + q2 * inv_phi_m2 + r * float(np.float32(inv_phi)), 1.0)
    if fr is not None:
        p_min, a_min = fr[..., 0, 1, None], fr[..., 0, 2, None]
        p_max, a_max = fr[..., 1, 1, None], fr[..., 1, 2, None]
        cmin, cmax = torch.cos(p_min), torch.cos(p_max)
        denom = float(n - 1) if n > 1 else 1.0
        lat = torch.acos(cmin - (cmin - cmax) * (i / denom))
        lon = a_min + (a_max - a_min) * frac
    else:
        lat = torch.acos(1.0 - 2.0 * i / float(n))
        lon = float(np.float32(2.0 * math.pi)) * frac
    sp = torch.sin(lat)
    return pl.out(torch.stack((sp * torch.cos(lon), sp * torch.sin(lon), torch.cos(lat)), dim=-1))


def visibility_directions(vertices: torch.Tensor, tv: torch.Tensor, active, num_rays: int) -> torch.Tensor:
    """Ray directions of the visibility query (reference ``_utils.py:1668-1700``): frustum over the
    three vertices AND the centroid of every active triangle, Fibonacci lattice inside it."""
    centers = tv.mean(dim=-2, keepdim=True)
    world = torch.cat((tv, centers), dim=-2).reshape(-1, 3)
    av = None if active is None else active.bool().repeat_interleave(4)
    frustum = viewing_frustum(vertices, world, av)
    return fibonacci_lattice(num_rays, frustum=frustum)


def triangles_visible_from_vertex(
    vertex,
    triangle_vertices,
    active_triangles=None,
    num_rays: int = 1_000_000,
    batch_size: int | None = 512,
    *,
    ray_directions=None,
    **kwargs: Any,
):
    """Visibility mask ``[*batch, T]`` — reference ``_utils.py:1540-1772``.

    The ray directions (viewing frustum → Fibonacci lattice) are generated on the device with the
    reference's formulas and handed to the kernel, as the reference's own launcher does
    (``_mesh.py:3216-3250``); pass ``ray_directions [*batch, num_rays, 3]`` to supply them.
    """
    del batch_size
    pl = Placement()
    vx = pl.put(vertex, torch.float32)
    tv = pl.put(triangle_vertices, torch.float32)
    act = None if active_triangles is None else pl.put(active_triangles, torch.uint8)
    T = int(tv.shape[-3])
    # reference broadcasting (_utils.py:1540-1548): vertex [*#b,3], triangles [*#b,T,3,3], active [*#b,T]
    batch = tuple(torch.broadcast_shapes(vx.shape[:-1], tv.shape[:-3], () if act is None else act.shape[:-1]))
    B = numel(batch)
    out = torch.zeros((*batch, T), dtype=torch.uint8, device=vx.device)
    if T == 0 or B == 0:
        return pl.out(out.view(torch.bool))
    vb = vx.expand(*batch, 3)
    dirs_all = None
    if ray_directions is not None:
        dirs_all = pl.put(ray_directions, torch.float32)
        num_rays = int(dirs_all.shape[-2])
        dirs_all = dirs_all.expand(*batch, num_rays, 3)
    eps = _default(kwargs.get("epsilon"), 10.0)
    for sel, tvi, acti in _mesh_batches(batch, tv, act):  # one launch per distinct mesh of the batch
        vi = vb[sel].reshape(-1, 3).contiguous()
        Bi = int(vi.shape[0])
        if Bi == 0:
            continue
        tvi = tvi.contiguous()
        acti = None if acti is None else acti.contiguous()
        if dirs_all is None:
            dirs = visibility_directions(vi, tvi, acti, num_rays)
        else:
            dirs = dirs_all[sel]
        dirs = dirs.reshape(Bi, num_rays, 3).contiguous()
        pack = pack_triangle_vertices(tvi, acti)
        res = torch.zeros((Bi, T), dtype=torch.uint8, device=vx.device)
        if use_cull(Bi * num_rays, T):
            # nearest hits behind the exact cull (identical to the all-pairs reduction), then the scatter
            origins = vi[:, None, :].expand(Bi, num_rays, 3).reshape(-1, 3).contiguous()
            idx = torch.empty(Bi * num_rays, dtype=torch.int32, device=vx.device)
            tt = torch.empty(Bi * num_rays, dtype=torch.float32, device=vx.device)
            first_hit_launch(pack, T, origins, dirs.reshape(-1, 3), eps, 512, idx, tt)
            check(lib.drt_scatter_visible(stream_ptr(), Bi, num_rays, T, ptr(idx), ptr(res)))
        else:
            check(
                lib.drt_triangles_visible_from_vertex(
                    stream_ptr(), Bi, num_rays, ptr(vi), ptr(dirs), ptr(pack), T, eps, ptr(res), None,
                )
            )
        out[sel] = res.view(out[sel].shape)
    return pl.out(out.view(torch.bool))


# ------------------------------------------------------------------------------------------------
# K5: image method
# ------------------------------------------------------------------------------------------------


def image_of_vertex_with_respect_to_mirror(vertex, mirror_vertex, mirror_normal):
    """``p - 2((p - m)·n)n`` — reference ``_solver_image_method.py:11-79``."""
    pl = Placement()
    p = pl.put(vertex, torch.float32)
    m = pl.put(mirror_vertex, torch.float32)
    n = pl.put(mirror_normal, torch.float32)
    batch = torch.broadcast_shapes(p.shape[:-1], m.shape[:-1], n.shape[:-1])
    out = torch.empty((*batch, 3), dtype=torch.float32, device=p.device)
    if numel(batch) > 0:
        ndim, shape, (sp, sm, sn), keep = batch_strides(batch, [(p, 1), (m, 1), (n, 1)])
        check(
            lib.drt_image_of_vertex_with_respect_to_mirror(
                stream_ptr(), ndim, shape, ptr(keep[0]), sp, ptr(keep[1]), sm, ptr(keep[2]), sn, ptr(out)
            )
        )
    return pl.out(out)


def intersection_of_ray_with_plane(ray_origin, ray_direction, plane_vertex, plane_normal):
    """Reference ``_solver_image_method.py:82-135`` (parallel rays → ``inf`` unless on the plane)."""
    pl = Placement()
    o = pl.put(ray_origin, torch.float32)
    d = pl.put(ray_direction, torch.float32)
    v = pl.put(plane_vertex, torch.float32)
    n = pl.put(plane_normal, torch.float32)
    batch = torch.broadcast_shapes(o.shape[:-1], d.shape[:-1], v.shape[:-1], n.shape[:-1])
    out = torch.empty((*batch, 3), dtype=torch.float32, device=o.device)
    if numel(batch) > 0:
        ndim, shape, (so, sd, sv, sn), keep = batch_strides(batch, [(o, 1), (d, 1), (v, 1), (n, 1)])
        check(
            lib.drt_intersection_of_ray_with_plane(
                stream_ptr(), ndim, shape, ptr(keep[0]), so, ptr(keep[1]), sd, ptr(keep[2]), sv,
                ptr(keep[3]), sn, ptr(out),
            )
        )
    return pl.out(out)


def _image_method_launch(f, t, mv, mn, batch, k):
    out = torch.empty((*batch, k, 3), dtype=torch.float32, device=f.device)
    if numel(batch) > 0 and k > 0:
        ndim, shape, (sf, st, sv, sn), keep = batch_strides(batch, [(f, 1), (t, 1), (mv, 2), (mn, 2)])
        check(
            lib.drt_image_method(
                stream_ptr(), ndim, shape, k, ptr(keep[0]), sf, ptr(keep[1]), st, ptr(keep[2]), sv,
                ptr(keep[3]), sn, ptr(out),
            )
        )
    return out


class _ImageMethod(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, t, mv, mn):
        k = int(mv.shape[-2])
        batch = torch.broadcast_shapes(f.shape[:-1], t.shape[:-1], mv.shape[:-2], mn.shape[:-2])
        ctx.save_for_backward(f, t, mv, mn)
        ctx.batch, ctx.k = tuple(batch), k
        return _image_method_launch(f, t, mv, mn, batch, k)

    @staticmethod
    def backward(ctx, g):
        f, t, mv, mn = ctx.saved_tensors
        batch, k = ctx.batch, ctx.k
        n = numel(batch)
        if k > _lib.DRT_MAX_ORDER:
            raise NotImplementedError(f"image_method gradient supports at most {_lib.DRT_MAX_ORDER} mirrors")
        g = g.contiguous().to(torch.float32)
        gf = torch.zeros((*batch, 3), dtype=torch.float32, device=g.device)
        gt = torch.zeros_like(gf)
        gv = torch.zeros((*batch, k, 3), dtype=torch.float32, device=g.device)
        gn = torch.zeros_like(gv)
        if n > 0:
            ndim, shape, (sf, st, sv, sn), keep = batch_strides(batch, [(f, 1), (t, 1), (mv, 2), (mn, 2)])
            check(
                lib.drt_image_method_vjp(
                    stream_ptr(), ndim, shape, k, ptr(keep[0]), sf, ptr(keep[1]), st, ptr(keep[2]), sv,
                    ptr(keep[3]), sn, ptr(g), ptr(gf), ptr(gt), ptr(gv), ptr(gn),
                )
            )
        return (
            sum_to_shape(gf, f.shape),
            sum_to_shape(gt, t.shape),
            sum_to_shape(gv, mv.shape),
            sum_to_shape(gn, mn.shape),
        )


def image_method(from_vertex, to_vertex, mirror_vertices, mirror_normals):
    """Image-method path points ``[*batch, k, 3]`` — reference ``_solver_image_method.py:206-363``."""
    pl = Placement()
    f = pl.put(from_vertex, torch.float32)
    t = pl.put(to_vertex, torch.float32)
    mv = pl.put(mirror_vertices, torch.float32)
    mn = pl.put(mirror_normals, torch.float32)
    if mv.shape[-2] != mn.shape[-2]:
        raise TypeError("mirror_vertices and mirror_normals must hold the same number of mirrors")
    if torch.is_grad_enabled() and any(x.requires_grad for x in (f, t, mv, mn)):
        return pl.out(_ImageMethod.apply(f, t, mv, mn))
    batch = torch.broadcast_shapes(f.shape[:-1], t.shape[:-1], mv.shape[:-2], mn.shape[:-2])
    return pl.out(_image_method_launch(f, t, mv, mn, batch, int(mv.shape[-2])))


def consecutive_vertices_are_on_same_side_of_mirror(
    vertices, mirror_vertices, mirror_normals, *, smoothing_factor=None
):
    """Reference ``_solver_image_method.py:386-454`` → bool ``[*batch, k]`` (float in [0, 1] with
    ``smoothing_factor``, forward only)."""
    pl = Placement()
    v = pl.put(vertices, torch.float32)
    mv = pl.put(mirror_vertices, torch.float32)
    mn = pl.put(mirror_normals, torch.float32)
    k = int(mv.shape[-2])
    if v.shape[-2] != k + 2:
        raise TypeError(f"vertices must hold num_mirrors + 2 = {k + 2} points, got {v.shape[-2]}")
    batch = torch.broadcast_shapes(v.shape[:-2], mv.shape[:-2], mn.shape[:-2])
    if smoothing_factor is not None:
        outf = torch.empty((*batch, k), dtype=torch.float32, device=v.device)
        if numel(batch) > 0 and k > 0:
            ndim, shape, (sv, sm, sn), keep = batch_strides(batch, [(v, 2), (mv, 2), (mn, 2)])
            check(
                lib.drt_consecutive_vertices_are_on_same_side_of_mirror_smooth(
                    stream_ptr(), ndim, shape, k, ptr(keep[0]), sv, ptr(keep[1]), sm, ptr(keep[2]), sn,
                    float(smoothing_factor), ptr(outf),
                )
            )
        return pl.out(outf)
    out = torch.empty((*batch, k), dtype=torch.uint8, device=v.device)
    if numel(batch) > 0 and k > 0:
        ndim, shape, (sv, sm, sn), keep = batch_strides(batch, [(v, 2), (mv, 2), (mn, 2)])
        check(
            lib.drt_consecutive_vertices_are_on_same_side_of_mirror(
                stream_ptr(), ndim, shape, k, ptr(keep[0]), sv, ptr(keep[1]), sm, ptr(keep[2]), sn, ptr(out)
            )
        )
    return pl.out(out.view(torch.bool))
