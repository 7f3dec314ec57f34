"""Reference-side binding: DiffeRT's three accelerated ``Mesh`` methods on top of differt_b200.

What a DiffeRT maintainer would add next to ``differt/src/differt/geometry/_mesh.py`` in place of the
Warp launchers (``_mesh.py:142-401``) and their ``wp.jax_callable`` call sites (``:266-276, 3082-3092,
3239-3250``).  NOT RUN IN THIS IMAGE (no jax/jaxlib) — it documents the binding; the same entry
points are exercised here through ctypes + torch by ``differt_b200/`` and ``tests/``.

Build the FFI library first (see ``integration/xla_ffi.cc``), then::

    import differt_b200_jax as drt_jax
    drt_jax.register("/path/to/libdiffert_b200_xla.so")
    hit = drt_jax.ray_intersect_any_triangle(mesh, ray_origins, ray_directions)
"""

from __future__ import annotations

import ctypes
from functools import partial

import jax
import jax.numpy as jnp
import numpy as np

_TARGETS = {
    "drt_ray_intersect_any_triangle": "DrtRayIntersectAnyTriangle",
    "drt_first_triangle_hit_by_ray": "DrtFirstTriangleHitByRay",
    "drt_first_triangle_hit_by_ray_vjp": "DrtFirstTriangleHitByRayVjp",
    "drt_triangles_visible_from_vertex": "DrtTrianglesVisibleFromVertex",
    "drt_trace_path_candidates": "DrtTracePathCandidates",
    "drt_trace_valid_path_candidates": "DrtTraceValidPathCandidates",
    "drt_complete_graph_candidates": "DrtCompleteGraphCandidates",
    "drt_trace_path_candidates_smooth": "DrtTracePathCandidatesSmooth",
    "drt_trace_path_candidates_smooth_vjp": "DrtTracePathCandidatesSmoothVjp",
    "drt_em_fresnel_coefficients": "DrtEmFresnelCoefficients",
    "drt_em_path_coefficients": "DrtEmPathCoefficients",
}


def register(library_path: str) -> None:
    lib = ctypes.CDLL(library_path)
    for name, symbol in _TARGETS.items():
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, symbol)), platform="CUDA")


def _mask_u8(mesh):
    # an empty buffer means "no mask" to the handlers
    return jnp.zeros((0,), jnp.uint8) if mesh.mask is None else mesh.mask.astype(jnp.uint8)


def _eps(dtype, factor):
    return np.float32(factor * jnp.finfo(dtype).eps)


def ray_intersect_any_triangle(mesh, ray_origins, ray_directions, *, hit_tol=None):
    """Replacement body of ``Mesh.ray_intersect_any_triangle`` (``_mesh.py:3018-3094``)."""
    ray_origins, ray_directions = jnp.broadcast_arrays(ray_origins, ray_directions)
    batch = ray_origins.shape[:-1]
    if mesh.triangles.shape[0] == 0:
        return jnp.full(batch, False)
    o, d = ray_origins.reshape(-1, 3), ray_directions.reshape(-1, 3)
    out = jax.ffi.ffi_call(
        "drt_ray_intersect_any_triangle",
        jax.ShapeDtypeStruct((o.shape[0],), jnp.bool_),
        vmap_method="sequential",
    )(
        jax.lax.stop_gradient(mesh.vertices), mesh.triangles, _mask_u8(mesh),
        jax.lax.stop_gradient(o), jax.lax.stop_gradient(d),
        epsilon=_eps(o.dtype, 10.0),
        hit_tol=np.float32(hit_tol) if hit_tol is not None else _eps(o.dtype, 100.0),
    )
    return jax.lax.stop_gradient(out.reshape(batch))


@partial(jax.custom_vjp, nondiff_argnums=())
def _first_hit(vertices, triangles, mask, o, d):
    face, t = jax.ffi.ffi_call(
        "drt_first_triangle_hit_by_ray",
        (jax.ShapeDtypeStruct((o.shape[0],), jnp.int32), jax.ShapeDtypeStruct((o.shape[0],), jnp.float32)),
        vmap_method="sequential",
    )(vertices, triangles, mask, o, d, epsilon=_eps(o.dtype, 10.0), batch_size=np.int64(512))
    return face, t


def _first_hit_fwd(vertices, triangles, mask, o, d):
    face, t = _first_hit(vertices, triangles, mask, o, d)
    return (face, t), (vertices, triangles, o, d, face)  # same residuals as _mesh.py:280-305


def _first_hit_bwd(res, cot):
    vertices, triangles, o, d, face = res
    _, g_t = cot
    g_v, g_o, g_d = jax.ffi.ffi_call(
        "drt_first_triangle_hit_by_ray_vjp",
        (jax.ShapeDtypeStruct(vertices.shape, jnp.float32), jax.ShapeDtypeStruct(o.shape, jnp.float32),
         jax.ShapeDtypeStruct(d.shape, jnp.float32)),
    )(vertices, triangles, o, d, face, g_t)
    return g_v, None, None, g_o, g_d  # triangles / mask get None, like _mesh.py:338


_first_hit.defvjp(_first_hit_fwd, _first_hit_bwd)


def first_triangle_hit_by_ray(mesh, ray_origins, ray_directions):
    """Replacement body of ``Mesh.first_triangle_hit_by_ray`` (``_mesh.py:3096-3162``)."""
    ray_origins, ray_directions = jnp.broadcast_arrays(ray_origins, ray_directions)
    batch = ray_origins.shape[:-1]
    if mesh.triangles.shape[0] == 0:
        return jnp.full(batch, -1, jnp.int32), jnp.full(batch, jnp.inf, ray_origins.dtype)
    face, t = _first_hit(
        mesh.vertices, mesh.triangles, _mask_u8(mesh), ray_origins.reshape(-1, 3), ray_directions.reshape(-1, 3)
    )
    return jax.lax.stop_gradient(face.reshape(batch)), t.reshape(batch)


def triangles_visible_from_vertex(mesh, vertex, ray_directions):
    """Replacement of the FFI call inside ``Mesh.triangles_visible_from_vertex`` (``_mesh.py:3239-3253``);
    ``ray_directions [*batch, num_rays, 3]`` are generated by the caller exactly as today
    (``viewing_frustum`` → ``fibonacci_lattice``, ``_mesh.py:3216-3228``)."""
    batch = vertex.shape[:-1]
    T = mesh.triangles.shape[0]
    out = jax.ffi.ffi_call(
        "drt_triangles_visible_from_vertex",
        jax.ShapeDtypeStruct((int(np.prod(batch, dtype=int)), T), jnp.bool_),
    )(
        jax.lax.stop_gradient(mesh.vertices), mesh.triangles, _mask_u8(mesh),
        jax.lax.stop_gradient(vertex.reshape(-1, 3)),
        jax.lax.stop_gradient(ray_directions.reshape(-1, ray_directions.shape[-2], 3)),
        epsilon=_eps(vertex.dtype, 10.0),
    )
    return jax.lax.stop_gradient(out.reshape(*batch, T))


def trace_path_candidates(mesh, tx_vertices, rx_vertices, path_candidates, *, epsilon=None, hit_tol=None,
                          min_len=None):
    """Fused replacement of ``_trace_path_candidates`` (``_solvers.py:499-770``, non-smoothing branch)
    → ``(vertices, objects, mask)``, the fields of ``TracedPaths``."""
    ntx, nrx, (C, k) = tx_vertices.shape[0], rx_vertices.shape[0], path_candidates.shape
    dt = tx_vertices.dtype
    return jax.ffi.ffi_call(
        "drt_trace_path_candidates",
        (jax.ShapeDtypeStruct((ntx, nrx, C, k + 2, 3), jnp.float32),
         jax.ShapeDtypeStruct((ntx, nrx, C, k + 2), jnp.int32),
         jax.ShapeDtypeStruct((ntx, nrx, C), jnp.bool_)),
    )(
        mesh.vertices, mesh.triangles, _mask_u8(mesh), tx_vertices, rx_vertices,
        path_candidates.astype(jnp.int32), assume_quads=bool(mesh.assume_quads),
        epsilon=np.float32(epsilon) if epsilon is not None else _eps(dt, 10.0),
        hit_tol=np.float32(hit_tol) if hit_tol is not None else _eps(dt, 100.0),
        min_len=np.float32(min_len) if min_len is not None else _eps(dt, 10.0),
    )


def generate_all_path_candidates(num_primitives: int, order: int, *, assume_quads: bool = False, start: int = 0,
                                 count: int | None = None):
    """Device-side replacement of ``ExhaustivePathTracer.generate_path_candidates``
    (``_solvers.py:803-848``): no host DFS, no host→device copy; any chunk of the linear index."""
    total = num_primitives * max(num_primitives - 1, 1) ** max(order - 1, 0) if order > 0 else 1
    count = total - start if count is None else min(count, total - start)
    return jax.ffi.ffi_call(
        "drt_complete_graph_candidates", jax.ShapeDtypeStruct((count, order), jnp.int32)
    )(num_nodes=np.int64(num_primitives), start=np.int64(start), stride_multiplier=np.int64(2 if assume_quads else 1))


def trace_valid_path_candidates(mesh, tx_vertices, rx_vertices, path_candidates, *, capacity: int = 1 << 16):
    """Compact trace: ``(count, index, vertices, objects, valid)`` with a static ``capacity``; the
    caller re-traces with a larger capacity when ``count > capacity``."""
    k = path_candidates.shape[1]
    dt = tx_vertices.dtype
    return jax.ffi.ffi_call(
        "drt_trace_valid_path_candidates",
        (jax.ShapeDtypeStruct((1,), jnp.int64), jax.ShapeDtypeStruct((capacity,), jnp.int64),
         jax.ShapeDtypeStruct((capacity, k + 2, 3), jnp.float32), jax.ShapeDtypeStruct((capacity, k + 2), jnp.int32),
         jax.ShapeDtypeStruct((capacity,), jnp.bool_)),
    )(
        mesh.vertices, mesh.triangles, _mask_u8(mesh), tx_vertices, rx_vertices, path_candidates.astype(jnp.int32),
        assume_quads=bool(mesh.assume_quads), epsilon=_eps(dt, 10.0), hit_tol=_eps(dt, 100.0), min_len=_eps(dt, 10.0),
        capacity=np.int64(capacity),
    )


def trace_path_candidates_smooth(mesh, tx_vertices, rx_vertices, path_candidates, smoothing_factor):
    """The relaxed branch of ``_trace_path_candidates`` (``_solvers.py:599-713``) as ONE differentiable
    call: ``(vertices, objects, confidence)`` with a ``jax.custom_vjp`` whose residuals are the two
    float outputs and whose backward pass is ``drt_trace_path_candidates_smooth_vjp``."""
    cand = path_candidates.astype(jnp.int32)
    ntx, nrx, (C, k) = tx_vertices.shape[0], rx_vertices.shape[0], cand.shape
    dt = tx_vertices.dtype
    attrs = dict(assume_quads=bool(mesh.assume_quads), epsilon=_eps(dt, 10.0), hit_tol=_eps(dt, 100.0),
                 min_len=_eps(dt, 10.0), smoothing_factor=np.float32(smoothing_factor))
    triangles, mask = mesh.triangles, _mask_u8(mesh)

    def call(vertices, tx, rx):
        return jax.ffi.ffi_call(
            "drt_trace_path_candidates_smooth",
            (jax.ShapeDtypeStruct((ntx, nrx, C, k + 2, 3), jnp.float32),
             jax.ShapeDtypeStruct((ntx, nrx, C, k + 2), jnp.int32),
             jax.ShapeDtypeStruct((ntx, nrx, C), jnp.float32)),
        )(vertices, triangles, mask, tx, rx, cand, **attrs)

    @jax.custom_vjp
    def traced(vertices, tx, rx):
        out_v, _, out_m = call(vertices, tx, rx)
        return out_v, out_m

    def fwd(vertices, tx, rx):
        out_v, _, out_m = call(vertices, tx, rx)
        return (out_v, out_m), (vertices, tx, rx, out_v, out_m)

    def bwd(res, cts):
        vertices, tx, rx, out_v, out_m = res
        g_v, g_m = cts
        return jax.ffi.ffi_call(
            "drt_trace_path_candidates_smooth_vjp",
            (jax.ShapeDtypeStruct(vertices.shape, jnp.float32), jax.ShapeDtypeStruct(tx.shape, jnp.float32),
             jax.ShapeDtypeStruct(rx.shape, jnp.float32)),
        )(vertices, triangles, mask, tx, rx, cand, out_v, out_m, g_v, g_m, **attrs)

    traced.defvjp(fwd, bwd)
    out_v, out_m = traced(mesh.vertices, tx_vertices, rx_vertices)
    # objects = [tx index, candidate..., rx index] (_solvers.py:721-750): integer bookkeeping, no kernel needed
    shape = (ntx, nrx, C, 1)
    objects = jnp.concatenate(
        (jnp.broadcast_to(jnp.arange(ntx, dtype=jnp.int32)[:, None, None, None], shape),
         jnp.broadcast_to(cand[None, None], (ntx, nrx, C, k)),
         jnp.broadcast_to(jnp.arange(nrx, dtype=jnp.int32)[None, :, None, None], shape)), axis=-1)
    return out_v, objects, out_m


def fresnel_coefficients(n_r, cos_theta_i):
    """Replacement body of ``differt.em.fresnel_coefficients`` (``em/_fresnel.py:46-213``):
    ``((r_s, r_p), (t_s, t_p))``, broadcast like the reference."""
    n_r, cos_theta_i = jnp.broadcast_arrays(jnp.asarray(n_r, jnp.complex64), jnp.asarray(cos_theta_i, jnp.float32))
    out = jax.ShapeDtypeStruct((n_r.size,), jnp.complex64)
    r_s, r_p, t_s, t_p = jax.ffi.ffi_call("drt_em_fresnel_coefficients", (out, out, out, out))(
        n_r.reshape(-1), cos_theta_i.reshape(-1))
    shape = n_r.shape
    return (r_s.reshape(shape), r_p.reshape(shape)), (t_s.reshape(shape), t_p.reshape(shape))


def path_coefficients(paths, mesh, n_r, thickness, frequency: float, *, polarization=("V", "V")):
    """The field chain of ``differt.plugins.deepmimo.export`` (``plugins/deepmimo.py:516-665, 694-696``) for the
    valid paths of ``paths`` in one custom call: ``(a [n], length [n], field [*tx, *rx], power [*tx, *rx])``.
    ``n_r`` / ``thickness`` are per triangle (``n_complex[mesh.face_materials]``)."""
    masked = paths.masked()  # dynamic shape: outside jit, like the reference's own `paths.masked()` callers
    n = masked.vertices.shape[0]
    pair_shape = paths.mask.shape[:-1]
    pair_index = jnp.nonzero(paths.mask.reshape(-1))[0] // paths.mask.shape[-1]
    pairs = int(np.prod(pair_shape)) if pair_shape else 1
    pol = {"V": 0, "H": 1}
    a, length, field, power = jax.ffi.ffi_call(
        "drt_em_path_coefficients",
        (jax.ShapeDtypeStruct((n,), jnp.complex64), jax.ShapeDtypeStruct((n,), jnp.float32),
         jax.ShapeDtypeStruct((pairs,), jnp.complex64), jax.ShapeDtypeStruct((pairs,), jnp.float32)),
    )(
        mesh.vertices, mesh.triangles, masked.vertices, masked.objects.astype(jnp.int32),
        jnp.asarray(n_r, jnp.complex64), jnp.asarray(thickness, jnp.float32), pair_index.astype(jnp.int64),
        frequency=np.float64(frequency), tx_polarization=np.int64(pol[polarization[0]]),
        rx_polarization=np.int64(pol[polarization[1]]),
    )
    return a, length, field.reshape(pair_shape), power.reshape(pair_shape)
