"""Shooting-and-bouncing rays: ``SBRPathLauncher.launch_paths`` and the multipath lifetime map.

Mirror of the reference's ray-launching callers of the hot path
(``differt/src/differt/geometry/_solvers.py:358-491, 1202-1226`` and
``differt/src/differt/geometry/_scene.py:62-302, 1250-1371``).  Every bounce is one launch of the
first-hit kernel (K3, the all-pairs engine) followed by one fused element-wise kernel
(``drt_sbr_bounce`` / ``drt_mlm_step``); ray generation (viewing frustum → Fibonacci lattice) runs on
the device with the reference's formulas, as for the visibility query.

Semantics delta: the nearest hit follows the pure-JAX ``first_triangle_hit_by_ray``
(``t > epsilon``); the reference asks Warp's BVH (``wp.mesh_query_ray``, third-party arithmetic).
"""

from __future__ import annotations

import dataclasses

import numpy as np
import torch

from . import geometry
from ._lib import check, lib
from ._tensor import F32_EPS, Placement, ptr, stream_ptr
from .mesh import Mesh, TracedPaths

__all__ = ["LaunchedPaths", "compute_tx_mlm", "launch_paths", "launch_rays"]


@dataclasses.dataclass
class LaunchedPaths:
    """Reference ``LaunchedPaths`` (``_paths.py:513-715``): paths of every order up to ``order`` from
    one ray launch.  The per-ray quantities are stored once (``ray_vertices [Ntx,R,order,3]``,
    ``ray_objects [Ntx,R,order]``); ``vertices`` / ``objects`` broadcast them over the receivers on
    demand, exactly the arrays the reference materialises (``_solvers.py:446-491``)."""

    tx_vertices: torch.Tensor   # [Ntx, 3]
    rx_vertices: torch.Tensor   # [Nrx, 3]
    ray_vertices: torch.Tensor  # [Ntx, R, order, 3]
    ray_objects: torch.Tensor   # [Ntx, R, order] i32
    masks: torch.Tensor         # [Ntx, Nrx, R, order + 1] bool
    confidence_threshold: float = 0.5

    @property
    def order(self) -> int:
        return int(self.ray_objects.shape[-1])

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(self.masks.shape[:-1])

    @property
    def mask(self) -> torch.Tensor:
        """Highest-order mask (``_paths.py:561-564``)."""
        return self.masks[..., -1]

    @property
    def vertices(self) -> torch.Tensor:
        """``[Ntx, Nrx, R, order + 2, 3]`` (``assemble_path``, ``_solvers.py:450-454``)."""
        ntx, nrx, r = self.shape
        k = self.order
        return torch.cat(
            (
                self.tx_vertices[:, None, None, None, :].expand(ntx, nrx, r, 1, 3),
                self.ray_vertices[:, None].expand(ntx, nrx, r, k, 3),
                self.rx_vertices[None, :, None, None, :].expand(ntx, nrx, r, 1, 3),
            ),
            dim=-2,
        )

    @property
    def objects(self) -> torch.Tensor:
        """``[Ntx, Nrx, R, order + 2]`` = tx index, hit triangles, rx index (``_solvers.py:456-481``)."""
        ntx, nrx, r = self.shape
        dev = self.ray_objects.device
        return torch.cat(
            (
                torch.arange(ntx, dtype=torch.int32, device=dev)[:, None, None, None].expand(ntx, nrx, r, 1),
                self.ray_objects[:, None].expand(ntx, nrx, r, self.order),
                torch.arange(nrx, dtype=torch.int32, device=dev)[None, :, None, None].expand(ntx, nrx, r, 1),
            ),
            dim=-1,
        )

    @property
    def interaction_types(self) -> torch.Tensor:
        ntx, nrx, r = self.shape
        return torch.zeros((1, 1, 1, 1), dtype=torch.int32, device=self.masks.device).expand(ntx, nrx, r, self.order)

    def get_paths(self, order: int) -> TracedPaths:
        """``LaunchedPaths.get_paths`` (``_paths.py:566-600``), compacted to the valid paths of that
        order so that the dense ``[Ntx,Nrx,R,…]`` arrays are never materialised."""
        if order < 0 or order > self.order:
            raise ValueError(
                f"Paths order must be strictly between 0 and {self.order} (incl.), but you provided {order}."
            )
        idx = self.masks[..., order].nonzero()  # row-major (tx, rx, ray), like masked()
        itx, irx, iray = idx[:, 0], idx[:, 1], idx[:, 2]
        verts = torch.cat(
            (self.tx_vertices[itx, None, :], self.ray_vertices[itx, iray, :order, :], self.rx_vertices[irx, None, :]),
            dim=-2,
        )
        objs = torch.cat(
            (itx[:, None].to(torch.int32), self.ray_objects[itx, iray, :order], irx[:, None].to(torch.int32)), dim=-1
        )
        n = idx.shape[0]
        return TracedPaths(
            vertices=verts, objects=objs, mask=torch.ones(n, dtype=torch.bool, device=verts.device),
            interaction_types=torch.zeros((n, order), dtype=torch.int32, device=verts.device),
            confidence_threshold=self.confidence_threshold,
        )

    @property
    def masked_vertices(self) -> torch.Tensor:
        return self.get_paths(self.order).vertices

    @property
    def masked_objects(self) -> torch.Tensor:
        return self.get_paths(self.order).objects


def launch_rays(mesh: Mesh, tx_vertices, rx_vertices, num_rays: int):
    """``SBRPathLauncher.launch_rays`` (``_solvers.py:1202-1226``) → origins, directions
    ``[Ntx, num_rays, 3]``."""
    pl = Placement()
    pl.device = mesh.vertices.device
    tx = pl.put(tx_vertices, torch.float32).reshape(-1, 3)
    rx = pl.put(rx_vertices, torch.float32).reshape(-1, 3)
    world = torch.cat((mesh.triangle_vertices.detach().reshape(-1, 3), rx), dim=0)
    frustums = geometry.viewing_frustum(tx, world)
    dirs = geometry.fibonacci_lattice(num_rays, frustum=frustums)
    return tx[:, None, :].expand(tx.shape[0], num_rays, 3).contiguous(), dirs.contiguous()


def _first_hit(pack, T, o, d, eps, faces, t, bvh=None):
    n = o.shape[0] * o.shape[1]
    if bvh is not None:  # opt-in accel="bvh"
        check(lib.drt_bvh_first_triangle_hit_by_ray(stream_ptr(), n, ptr(o), ptr(d), ptr(bvh), T, eps, 512, ptr(faces), ptr(t)))
    else:
        geometry.first_hit_launch(pack, T, o.reshape(n, 3), d.reshape(n, 3), eps, 512, faces, t)


def launch_paths(
    mesh: Mesh, tx_vertices, rx_vertices, order: int, *, num_rays: int = 1_000_000, epsilon=None,
    max_dist: float = 1e-3, ray_directions=None, accel: str = "brute",
) -> LaunchedPaths:
    """``SBRPathLauncher.launch_paths`` (``_solvers.py:358-491``): ``order + 1`` bounces of nearest hit
    → receivers in the vicinity of the segment (``filter_rays``) → specular bounce (``bounce_rays``).
    ``ray_directions [Ntx, R, 3]`` may be supplied instead of being generated."""
    if mesh.num_triangles == 0:
        raise NotImplementedError("launch_paths needs a non-empty mesh")
    pl = Placement()
    dev = pl.device = mesh.vertices.device
    tx = pl.put(tx_vertices, torch.float32).reshape(-1, 3).contiguous()
    rx = pl.put(rx_vertices, torch.float32).reshape(-1, 3).contiguous()
    if ray_directions is None:
        o, d = launch_rays(mesh, tx, rx, num_rays)
    else:
        d = pl.put(ray_directions, torch.float32).reshape(tx.shape[0], -1, 3).contiguous().clone()
        o = tx[:, None, :].expand_as(d).contiguous()
    ntx, nrays, nrx, T = d.shape[0], d.shape[1], rx.shape[0], mesh.num_triangles
    eps = 10.0 * F32_EPS if epsilon is None else float(epsilon)
    pack = geometry.pack_mesh(mesh.vertices.detach(), mesh.triangles, mesh._mask_u8())
    bvh = mesh.build_bvh() if accel == "bvh" else None
    valid = torch.ones((ntx, nrays), dtype=torch.uint8, device=dev)
    # bounce-major storage (what the reference's scan stacks); the public views move that axis last
    masks = torch.empty((order + 1, ntx, nrx, nrays), dtype=torch.uint8, device=dev)
    verts = torch.empty((order + 1, ntx, nrays, 3), dtype=torch.float32, device=dev)
    faces = torch.empty((order + 1, ntx, nrays), dtype=torch.int32, device=dev)
    t_hit = torch.empty((ntx, nrays), dtype=torch.float32, device=dev)
    for b in range(order + 1):
        _first_hit(pack, T, o, d, eps, faces[b], t_hit, bvh)
        check(
            lib.drt_sbr_bounce(
                stream_ptr(), ntx, nrays, nrx, T, ptr(pack), ptr(o), ptr(d), ptr(valid), ptr(faces[b]),
                ptr(t_hit), ptr(rx), float(max_dist), ptr(masks[b]), ptr(verts[b]),
            )
        )
    return LaunchedPaths(
        tx_vertices=tx, rx_vertices=rx,
        ray_vertices=verts[:order].permute(1, 2, 0, 3),
        ray_objects=faces[:order].permute(1, 2, 0),
        masks=masks.view(torch.bool).permute(1, 2, 3, 0),
    )


def compute_tx_mlm(
    mesh: Mesh, tx_vertices, *, max_order: int, min_order: int = 0, dim_x: int, dim_y: int,
    num_rays: int = 1_000_000, receiver_height: float, min_x: float, max_x: float, min_y: float, max_y: float,
    ray_directions=None, epsilon: float = 1e-4, accel: str = "brute",
) -> torch.Tensor:
    """Multipath lifetime map (reference ``_compute_tx_mlm``, ``_scene.py:226-302`` + kernel ``:81-171``):
    for every transmitter, every cell of the ``dim_x × dim_y`` receiver grid at ``z = receiver_height``
    receives the bitwise OR of the hashes of the ray paths (sequences of hit primitives) that cross
    it after ``min_order … max_order`` bounces → ``[Ntx, dim_x, dim_y] uint32`` (as ``int64`` holding
    the unsigned values, torch has no arithmetic uint32)."""
    pl = Placement()
    dev = pl.device = mesh.vertices.device
    tx = pl.put(tx_vertices, torch.float32).reshape(-1, 3).contiguous()
    ntx, T = tx.shape[0], mesh.num_triangles
    if ray_directions is None:
        world = mesh.triangle_vertices.detach().reshape(-1, 3)
        active = None if mesh.mask is None else mesh.mask.repeat_interleave(3)
        corners = torch.tensor(
            [[min_x, min_y, receiver_height], [max_x, min_y, receiver_height],
             [max_x, max_y, receiver_height], [min_x, max_y, receiver_height]], dtype=torch.float32, device=dev)
        world = torch.cat((world, corners), dim=0)
        if active is not None:
            active = torch.cat((active, torch.ones(4, dtype=torch.bool, device=dev)))
        f = geometry.viewing_frustum(tx, world, active).clone()
        f[..., 1, 1] = float(np.float32(np.pi))  # _scene.py:262 ("TODO: fixme" in the reference)
        d = geometry.fibonacci_lattice(num_rays, frustum=f).contiguous()
    else:
        d = pl.put(ray_directions, torch.float32).reshape(ntx, -1, 3).contiguous().clone()
    nrays = d.shape[1]
    o = tx[:, None, :].expand_as(d).contiguous()
    out = torch.zeros((ntx, dim_x, dim_y), dtype=torch.int32, device=dev)  # uint32 bit patterns
    hashes = torch.full((ntx, nrays), 0x811C9DC5 - (1 << 32), dtype=torch.int32, device=dev)
    alive = torch.ones((ntx, nrays), dtype=torch.uint8, device=dev)
    faces = torch.full((ntx, nrays), -1, dtype=torch.int32, device=dev)
    t_first = torch.full((ntx, nrays), float("inf"), dtype=torch.float32, device=dev)
    pack = bvh = None
    if T > 0:
        pack = geometry.pack_mesh(mesh.vertices.detach(), mesh.triangles, mesh._mask_u8())
        bvh = mesh.build_bvh() if accel == "bvh" else None
    for it in range(max_order + 1):
        if T > 0:
            _first_hit(pack, T, o, d, 10.0 * F32_EPS, faces, t_first, bvh)
        check(
            lib.drt_mlm_step(
                stream_ptr(), ntx, nrays, T, ptr(pack), ptr(o), ptr(d), ptr(hashes), ptr(alive), ptr(faces),
                ptr(t_first), it, min_order, int(mesh.assume_quads), float(receiver_height), float(min_x),
                float(max_x), float(min_y), float(max_y), dim_x, dim_y, float(epsilon), ptr(out),
            )
        )
    return out.to(torch.int64) & 0xFFFFFFFF
