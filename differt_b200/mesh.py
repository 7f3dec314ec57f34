"""``Mesh`` and ``TracedPaths``: the data the hot path reads and writes, in the reference's layout.

Only the fields and methods on the hot path are provided (reference
``differt/src/differt/geometry/_mesh.py:612-700, 899-956, 3018-3253`` and
``differt/src/differt/geometry/_paths.py:77-116, 299-328``); mesh editing, loading and plotting are out
of scope.
"""

from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np
import torch

from . import geometry, scenes
from ._lib import check, lib
from ._tensor import F32_EPS, Placement, numel, ptr, require_cuda, stream_ptr

__all__ = ["Mesh", "TracedPaths"]


def _check_accel(accel: str) -> str:
    if accel not in ("brute", "bvh"):
        raise ValueError(f"accel must be 'brute' or 'bvh', got {accel!r}")
    return accel


@dataclasses.dataclass
class Mesh:
    """Triangle mesh resident on the GPU: ``vertices [V,3] f32``, ``triangles [T,3] i32``,
    optional ``mask [T] bool`` of active triangles, ``assume_quads`` (even/odd triangles pair up)."""

    vertices: torch.Tensor
    triangles: torch.Tensor
    mask: torch.Tensor | None = None
    assume_quads: bool = False

    def __post_init__(self) -> None:
        dev = self.vertices.device if isinstance(self.vertices, torch.Tensor) and self.vertices.is_cuda else require_cuda()
        self.vertices = torch.as_tensor(self.vertices).to(dev, torch.float32).contiguous()
        self.triangles = torch.as_tensor(self.triangles).to(dev, torch.int32).contiguous()
        if self.mask is not None:
            self.mask = torch.as_tensor(self.mask).to(dev, torch.bool).contiguous()
        if self.vertices.ndim != 2 or self.vertices.shape[-1] != 3:
            raise TypeError("vertices must have shape [num_vertices, 3]")
        if self.triangles.ndim != 2 or self.triangles.shape[-1] != 3:
            raise TypeError("triangles must have shape [num_triangles, 3]")
        if self.assume_quads and self.triangles.shape[0] % 2 != 0:
            raise ValueError("assume_quads needs an even number of triangles")  # _mesh.py:650-660

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def box(cls, length=1.0, width=1.0, height=1.0, *, with_top=False, with_bottom=True) -> "Mesh":
        """Reference ``Mesh.box`` (``_mesh.py:2113-2217``): same vertex and triangle order."""
        v, t = scenes.box(length, width, height, with_top=with_top, with_bottom=with_bottom)
        return cls(torch.from_numpy(v), torch.from_numpy(t))

    @classmethod
    def from_numpy(cls, vertices: np.ndarray, triangles: np.ndarray, mask=None, assume_quads=False) -> "Mesh":
        return cls(
            torch.from_numpy(np.ascontiguousarray(vertices, np.float32)),
            torch.from_numpy(np.ascontiguousarray(triangles, np.int32)),
            None if mask is None else torch.from_numpy(np.ascontiguousarray(mask, bool)),
            assume_quads,
        )

    def set_assume_quads(self, flag: bool = True) -> "Mesh":
        return dataclasses.replace(self, assume_quads=flag)

    # -- properties -----------------------------------------------------------------------------
    @property
    def num_triangles(self) -> int:
        return int(self.triangles.shape[0])

    @property
    def num_primitives(self) -> int:
        return self.num_triangles // 2 if self.assume_quads else self.num_triangles

    @property
    def triangle_vertices(self) -> torch.Tensor:
        """``vertices[triangles]`` → ``[T,3,3]`` (reference ``_mesh.py:899-905``)."""
        if self.triangles.numel() == 0:
            return self.vertices.new_empty((0, 3, 3))
        return self.vertices[self.triangles.long()]

    @property
    def normals(self) -> torch.Tensor:
        """Unit normals ``normalize((v1-v0) × (v2-v1))`` (reference ``_mesh.py:950-956``), computed
        by the packing kernel."""
        if self.num_triangles == 0:
            return self.vertices.new_empty((0, 3))
        pack = geometry.pack_mesh(self.vertices, self.triangles, None)
        return geometry.pack_normals(pack, self.num_triangles).contiguous()

    def _mask_u8(self) -> torch.Tensor | None:
        return None if self.mask is None else self.mask.to(torch.uint8)

    def _trace_prepared(self) -> torch.Tensor:
        """The mesh-only part of the trace (packed mesh, area order, culled hierarchy:
        ``drt_trace_prepare``), built once per mesh STATE and copied to the front of every trace
        call's workspace.  The key is the storage and version counter of ``vertices`` / ``triangles`` /
        ``mask`` — an in-place update (an optimizer step) bumps the version and the next call rebuilds,
        so the bytes can never describe another mesh (the staleness the reference's process-global
        ``_WARP_MESHES_CACHE`` has, ``_mesh.py:48-55``).  The library itself keeps no state."""
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) if t is not None else None
                    for t in (self.vertices, self.triangles, self.mask))
        cached = self.__dict__.get("_prepared_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        T = self.num_triangles
        blob = torch.empty(max(lib.drt_trace_prepared_bytes(T), 1), dtype=torch.uint8, device=self.vertices.device)
        mask_u8 = self._mask_u8()
        check(lib.drt_trace_prepare(stream_ptr(), self.vertices.shape[0], T, ptr(self.vertices.detach()),
                                    ptr(self.triangles), ptr(mask_u8), ptr(blob), blob.numel()))
        self.__dict__["_prepared_cache"] = (key, blob)
        return blob

    def build_bvh(self, relative_pad: float = 1e-4) -> torch.Tensor:
        """Linear BVH over the (masked) triangles for the opt-in ``accel="bvh"`` queries
        (``drt_bvh_build``): a caller-owned blob, rebuilt in ~0.1 ms — nothing is cached, so it can
        never go stale the way the reference's ``_WARP_MESHES_CACHE`` can (``_mesh.py:48-55``)."""
        T = self.num_triangles
        dev = self.vertices.device
        pack = geometry.pack_mesh(self.vertices.detach(), self.triangles, self._mask_u8())
        bvh = torch.empty(max(lib.drt_bvh_bytes(T), 256), dtype=torch.uint8, device=dev)
        ws = torch.empty(max(lib.drt_bvh_workspace_bytes(T), 256), dtype=torch.uint8, device=dev)
        check(lib.drt_bvh_build(stream_ptr(), T, ptr(pack), float(relative_pad), ptr(ws), ws.numel(), ptr(bvh)))
        return bvh

    # -- the three accelerated queries (reference: Warp launchers, _mesh.py:3018-3253) ------------
    def ray_intersect_any_triangle(self, ray_origins, ray_directions, *, hit_tol=None, epsilon=None,
                                   accel: str = "brute"):
        """Reference ``Mesh.ray_intersect_any_triangle`` (``_mesh.py:3018-3094``); no gradient.
        ``accel="bvh"`` opts into the BVH traversal (see ``csrc/bvh.cu`` for the exactness caveat)."""
        pl = Placement()
        pl.device = self.vertices.device
        o = pl.put(ray_origins, torch.float32)
        d = pl.put(ray_directions, torch.float32)
        batch = torch.broadcast_shapes(o.shape[:-1], d.shape[:-1])
        out = torch.zeros(batch, dtype=torch.uint8, device=o.device)
        R = numel(batch)
        if self.num_triangles == 0 or R == 0:
            return pl.out(out.view(torch.bool))
        o = o.detach().expand(*batch, 3).reshape(R, 3).contiguous()
        d = d.detach().expand(*batch, 3).reshape(R, 3).contiguous()
        eps_ = 10.0 * F32_EPS if epsilon is None else float(epsilon)
        tol_ = 100.0 * F32_EPS if hit_tol is None else float(hit_tol)
        if _check_accel(accel) == "bvh":
            bvh = self.build_bvh()
            check(
                lib.drt_bvh_ray_intersect_any_triangle(
                    stream_ptr(), R, ptr(o), ptr(d), ptr(bvh), self.num_triangles, eps_, tol_, ptr(out)
                )
            )
            return pl.out(out.view(torch.bool))
        pack = geometry.pack_mesh(self.vertices.detach(), self.triangles, self._mask_u8())
        if R >= geometry._SORT_MIN_RAYS:
            pack = geometry.sort_pack_by_area(pack, self.num_triangles)
        if geometry.use_cull(R, self.num_triangles):
            # same test, same results, behind the exact conservative cull (csrc/cull.cuh)
            ws = torch.empty(lib.drt_any_hit_workspace_bytes(self.num_triangles), dtype=torch.uint8, device=o.device)
            check(
                lib.drt_ray_intersect_any_triangle_culled(
                    stream_ptr(), R, ptr(o), ptr(d), ptr(pack), self.num_triangles, eps_, tol_, ptr(ws), ws.numel(),
                    ptr(out), None,
                )
            )
            return pl.out(out.view(torch.bool))
        check(
            lib.drt_ray_intersect_any_triangle(
                stream_ptr(), R, ptr(o), ptr(d), ptr(pack), self.num_triangles, eps_, tol_, ptr(out), None,
            )
        )
        return pl.out(out.view(torch.bool))

    def first_triangle_hit_by_ray(self, ray_origins, ray_directions, *, epsilon=None, batch_size=512,
                                  accel: str = "brute"):
        """Reference ``Mesh.first_triangle_hit_by_ray`` (``_mesh.py:3096-3162``): ``(index, t)`` with
        ``t`` differentiable w.r.t. origins, directions and ``self.vertices`` (``custom_vjp``,
        ``_mesh.py:258-344``); the index carries no gradient."""
        pl = Placement()
        pl.device = self.vertices.device
        o = pl.put(ray_origins, torch.float32)
        d = pl.put(ray_directions, torch.float32)
        batch = torch.broadcast_shapes(o.shape[:-1], d.shape[:-1])
        R = numel(batch)
        idx = torch.full((R,), -1, dtype=torch.int32, device=o.device)
        t = torch.full((R,), float("inf"), dtype=torch.float32, device=o.device)
        if self.num_triangles == 0 or R == 0:
            return pl.out(idx.view(batch)), pl.out(t.view(batch))
        of = o.expand(*batch, 3).reshape(R, 3).contiguous()
        df = d.expand(*batch, 3).reshape(R, 3).contiguous()
        eps_ = 10.0 * F32_EPS if epsilon is None else float(epsilon)
        bs_ = 0 if batch_size is None else int(batch_size)
        if _check_accel(accel) == "bvh":
            bvh = self.build_bvh()
            check(
                lib.drt_bvh_first_triangle_hit_by_ray(
                    stream_ptr(), R, ptr(of), ptr(df), ptr(bvh), self.num_triangles, eps_, bs_, ptr(idx), ptr(t)
                )
            )
        else:
            pack = geometry.pack_mesh(self.vertices.detach(), self.triangles, self._mask_u8())
            geometry.first_hit_launch(pack, self.num_triangles, of.detach(), df.detach(), eps_, bs_, idx, t)
        if torch.is_grad_enabled() and any(x.requires_grad for x in (of, df, self.vertices)):
            t = geometry._FirstHitDistanceGrad.apply(t, self.vertices, self.triangles, of, df, idx)
        return pl.out(idx.view(batch)), pl.out(t.view(batch))

    def triangles_visible_from_vertex(self, vertex, num_rays: int = 1_000_000, *, accel: str = "brute",
                                      **kwargs: Any):
        """Reference ``Mesh.triangles_visible_from_vertex`` (``_mesh.py:3164-3253``); no gradient.
        ``accel="bvh"``: the same rays, nearest hits from the BVH, then the scatter."""
        if _check_accel(accel) == "brute" or self.num_triangles == 0:
            return geometry.triangles_visible_from_vertex(
                vertex, self.triangle_vertices.detach(), self.mask, num_rays=num_rays, **kwargs
            )
        pl = Placement()
        pl.device = self.vertices.device
        vx = pl.put(vertex, torch.float32)
        batch = tuple(vx.shape[:-1])
        B, T = numel(batch), self.num_triangles
        out = torch.zeros((*batch, T), dtype=torch.uint8, device=vx.device)
        if B == 0:
            return pl.out(out.view(torch.bool))
        dirs = kwargs.get("ray_directions")
        if dirs is None:
            dirs = geometry.visibility_directions(vx.reshape(B, 3), self.triangle_vertices.detach(), self.mask, num_rays)
        dirs = pl.put(dirs, torch.float32).reshape(B, -1, 3).contiguous()
        n = int(dirs.shape[1])
        origins = vx.reshape(B, 1, 3).expand(B, n, 3).contiguous()
        idx = torch.empty(B * n, dtype=torch.int32, device=vx.device)
        tt = torch.empty(B * n, dtype=torch.float32, device=vx.device)
        bvh = self.build_bvh()
        eps_ = 10.0 * F32_EPS if kwargs.get("epsilon") is None else float(kwargs["epsilon"])
        check(
            lib.drt_bvh_first_triangle_hit_by_ray(
                stream_ptr(), B * n, ptr(origins), ptr(dirs), ptr(bvh), T, eps_, 0, ptr(idx), ptr(tt)
            )
        )
        check(lib.drt_scatter_visible(stream_ptr(), B, n, T, ptr(idx), ptr(out)))
        return pl.out(out.view(torch.bool))


@dataclasses.dataclass
class TracedPaths:
    """Reference ``TracedPaths`` (``_paths.py:77-116``): dense per-candidate fields + validity mask."""

    vertices: torch.Tensor           # [*batch, k+2, 3] f32
    objects: torch.Tensor            # [*batch, k+2] i32
    mask: torch.Tensor               # [*batch] bool (float confidence with smoothing_factor)
    interaction_types: torch.Tensor  # [*batch, k] i32
    confidence_threshold: float = 0.5
    stats: dict | None = None

    @property
    def order(self) -> int:
        return int(self.objects.shape[-1]) - 2

    def reshape(self, *batch: int) -> "TracedPaths":
        k = self.order
        return dataclasses.replace(
            self,
            vertices=self.vertices.reshape(*batch, k + 2, 3),
            objects=self.objects.reshape(*batch, k + 2),
            mask=self.mask.reshape(*batch),
            interaction_types=self.interaction_types.reshape(*batch, k),
        )

    def _valid(self) -> torch.Tensor:
        """Boolean validity: a float mask (relaxed trace) is a confidence, valid when
        ``>= confidence_threshold`` (``_paths.py:101-103, 270-283``); NaN is never valid."""
        if self.mask.dtype == torch.bool:
            return self.mask
        return self.mask >= self.confidence_threshold

    @property
    def num_valid_paths(self) -> int:
        return int(self._valid().sum().item())

    def _valid_flat_indices(self) -> torch.Tensor:
        """Flat row-major indices of the valid paths, ascending (``drt_compact_valid_paths``: the
        stable compaction runs on the device; one 8-byte device→host read sizes the result)."""
        k = self.order
        P = numel(self.mask.shape)
        dev = self.vertices.device
        m = self._valid().reshape(P).to(torch.uint8).contiguous()
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        ws = torch.empty(max(lib.drt_compact_workspace_bytes(P), 1), dtype=torch.uint8, device=dev)
        index = torch.empty(P, dtype=torch.int64, device=dev)
        check(
            lib.drt_compact_valid_paths(
                stream_ptr(), P, k, None, None, ptr(m), P, ptr(ws), ws.numel(), ptr(count), ptr(index),
                None, None,
            )
        )
        return index[: int(count.item())]

    def masked(self, index: torch.Tensor | None = None) -> "TracedPaths":
        """Keep the valid paths only, flattened in row-major order (reference ``_paths.py:299-328``)."""
        k = self.order
        P = numel(self.mask.shape)
        dev = self.vertices.device
        if index is None:
            index = self._valid_flat_indices()
        n = int(index.shape[0])
        vertices = self.vertices.reshape(P, k + 2, 3)[index]  # keeps the autograd graph
        return TracedPaths(
            vertices=vertices,
            objects=self.objects.reshape(P, k + 2)[index],
            mask=torch.ones(n, dtype=torch.bool, device=dev),
            interaction_types=self.interaction_types.reshape(P, k)[index],
            confidence_threshold=self.confidence_threshold,
        )

    def reduce(self, fun, axis=None) -> torch.Tensor:
        """``sum(fun(vertices) * mask)`` over ``axis`` (all axes by default) — reference
        ``_paths.py:461-479``: the first consumer of traced paths (received power, delay spread …).
        With a float mask (relaxed trace) the confidences weight the sum and the result is
        differentiable w.r.t. the scene through both ``vertices`` and ``mask``; with a boolean mask
        invalid paths contribute exactly zero (``where=mask``: their ``fun`` value, possibly NaN or
        inf, is never added)."""
        values = fun(self.vertices)
        if self.mask.dtype != torch.bool:
            values = values * self.mask
        else:
            values = torch.where(self.mask, values, torch.zeros_like(values))
        return values.sum() if axis is None else values.sum(dim=axis)

    @property
    def masked_vertices(self) -> torch.Tensor:
        return self.masked().vertices

    @property
    def masked_objects(self) -> torch.Tensor:
        return self.masked().objects
