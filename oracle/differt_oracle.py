"""CPU oracle for the DiffeRT geometric hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy float32 restatement of the reference's *pure-JAX* algorithms (the real reference needs
jax/equinox/warp, none of which exist in this image).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; nothing
under ``differt_b200/`` does.

Parity pin: every function below is checked against the reference's own known-answer tests
(``tests/test_oracle_golden.py`` ← ``tests/golden/reference_kats.json`` ← the reference test
files, see ``tests/golden/make_golden.py``), including the ``two_buildings.obj`` order 0-4 golden
paths of ``differt/tests/geometry/test_scene.py:116-160``.

Arithmetic contract (what the CUDA kernels must reproduce bit-for-bit on masks):
all values float32, no fused multiply-add, IEEE division / square root, three-term sums evaluated
left to right ``((x0*y0 + x1*y1) + x2*y2)``, comparisons with NaN are false.

File:line citations are relative to ``/root/reference/differt/src/differt/geometry/``.
"""

from __future__ import annotations

import numpy as np

F32 = np.float32
EPS = np.finfo(np.float32).eps  # 2**-23
INF = F32(np.inf)


def _f(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float32)


def dot3(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``jnp.sum(a * b, axis=-1)`` for 3-vectors, left-to-right, fp32, no FMA."""
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def cross3(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """``jnp.cross`` for 3-vectors: (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)."""
    a, b = np.broadcast_arrays(a, b)
    out = np.empty(a.shape, dtype=np.float32)
    out[..., 0] = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    out[..., 1] = a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2]
    out[..., 2] = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    return out


# --------------------------------------------------------------------------------------
# a10: normalize / normals / triangle_vertices / assemble_path
# --------------------------------------------------------------------------------------


def normalize(vectors) -> tuple[np.ndarray, np.ndarray]:
    """``_utils.py:66-72``: divide by the L2 norm, or by 1 where the norm is 0."""
    v = _f(vectors)
    with np.errstate(all="ignore"):
        lengths = np.sqrt(dot3(v, v))
        safe = np.where(lengths == 0.0, F32(1.0), lengths)
        return v / safe[..., None], lengths


def triangle_vertices(vertices, triangles) -> np.ndarray:
    """``_mesh.py:899-905``: gather ``vertices[triangles]`` → [T,3,3]."""
    return _f(vertices)[np.asarray(triangles, dtype=np.int64)]


def triangle_normals(tri) -> np.ndarray:
    """``_mesh.py:950-956``: ``normalize(cross(v1 - v0, v2 - v1))``."""
    tri = _f(tri)
    a = tri[..., 1, :] - tri[..., 0, :]
    b = tri[..., 2, :] - tri[..., 1, :]
    with np.errstate(all="ignore"):
        return normalize(cross3(a, b))[0]


def assemble_path(from_vertex, intermediate, to_vertex) -> np.ndarray:
    """``_utils.py:536-565``: concatenate [from, *intermediate, to] along axis -2."""
    from_vertex, intermediate, to_vertex = _f(from_vertex), _f(intermediate), _f(to_vertex)
    batch = np.broadcast_shapes(
        from_vertex.shape[:-1], intermediate.shape[:-2], to_vertex.shape[:-1]
    )
    return np.concatenate(
        (
            np.broadcast_to(from_vertex[..., None, :], (*batch, 1, 3)),
            np.broadcast_to(intermediate, (*batch, *intermediate.shape[-2:])),
            np.broadcast_to(to_vertex[..., None, :], (*batch, 1, 3)),
        ),
        axis=-2,
    )


# --------------------------------------------------------------------------------------
# a1: Möller–Trumbore
# --------------------------------------------------------------------------------------


def ray_intersect_triangle(ray_origins, ray_directions, tri, *, epsilon=None):
    """``_utils.py:1253-1322`` (non-smoothing branch) → ``(t, hit)``.

    ``t`` is returned even where ``hit`` is false; ``a == 0`` is replaced by ``+inf`` so that
    ``f = 0`` and ``t = 0`` for parallel rays.
    """
    o, d, tri = _f(ray_origins), _f(ray_directions), _f(tri)
    eps = F32(10.0 * EPS) if epsilon is None else F32(epsilon)
    v0, v1, v2 = tri[..., 0, :], tri[..., 1, :], tri[..., 2, :]
    with np.errstate(all="ignore"):
        e1 = v1 - v0
        e2 = v2 - v0
        h = cross3(d, e2)
        a = dot3(h, e1)
        a = np.where(a == 0.0, INF, a)
        hit = np.abs(a) > eps
        f = F32(1.0) / a
        s = o - v0
        u = f * dot3(s, h)
        hit = hit & (u >= 0.0) & (u <= 1.0)
        q = cross3(s, e1)
        v = f * dot3(q, d)
        hit = hit & (v >= 0.0) & (u + v <= 1.0)
        t = f * dot3(q, e2)
        hit = hit & (t > eps)
    return t.astype(np.float32), hit


# --------------------------------------------------------------------------------------
# a2: any-hit, a3: first-hit, a4: visibility
# --------------------------------------------------------------------------------------


def ray_intersect_any_triangle(
    ray_origins, ray_directions, tri, active=None, *, hit_tol=None, epsilon=None, chunk=4096
):
    """``_utils.py:1414-1537``: ``any_j[(t_ij < 1 - hit_tol) & hit_ij & active_j]``.

    ``tri`` is [T,3,3] shared by all rays (the only form the hot path uses); ``T == 0`` → False.
    The reference's ``batch_size`` only bounds memory (OR is order independent).
    """
    o, d, tri = _f(ray_origins), _f(ray_directions), _f(tri)
    o, d = np.broadcast_arrays(o, d)
    batch = o.shape[:-1]
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    tol = F32(100.0 * EPS) if hit_tol is None else F32(hit_tol)
    thr = F32(1.0) - tol
    T = tri.shape[0]
    out = np.zeros(o.shape[0], dtype=bool)
    if T == 0:
        return out.reshape(batch)
    act = None if active is None else np.asarray(active, dtype=bool)
    rchunk = max(1, (1 << 22) // max(T, 1))
    for r0 in range(0, o.shape[0], rchunk):
        t, hit = ray_intersect_triangle(
            o[r0 : r0 + rchunk, None, :], d[r0 : r0 + rchunk, None, :], tri[None], epsilon=epsilon
        )
        m = (t < thr) & hit
        if act is not None:
            m &= act[None, :]
        out[r0 : r0 + rchunk] = m.any(axis=-1)
    return out.reshape(batch)


def first_triangle_hit_by_ray(
    ray_origins, ray_directions, tri, active=None, *, batch_size=512, epsilon=None
):
    """``_utils.py:1821-1960`` → ``(idx i32, t f32)``; miss = ``(-1, +inf)``.

    Tie rule restated exactly: ``argmin`` (first minimum) inside each batch of ``batch_size``
    triangles (``:1886``), and across batches the carry survives only if ``carry_t < new_t``
    (``:1865-1868``) — i.e. the *latest* batch wins an exact tie.
    """
    o, d, tri = _f(ray_origins), _f(ray_directions), _f(tri)
    o, d = np.broadcast_arrays(o, d)
    batch = o.shape[:-1]
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    R, T = o.shape[0], tri.shape[0]
    idx = np.full(R, -1, dtype=np.int32)
    tmin = np.full(R, INF, dtype=np.float32)
    if T == 0:
        return idx.reshape(batch), tmin.reshape(batch)
    bs = T if batch_size is None else max(min(int(batch_size), T), 1)
    nb, rem = divmod(T, bs)
    act = None if active is None else np.asarray(active, dtype=bool)
    starts = [b * bs for b in range(nb)]
    sizes = [bs] * nb
    if rem > 0:
        starts.append(T - rem)
        sizes.append(rem)
    rchunk = max(1, (1 << 22) // bs)
    for r0 in range(0, R, rchunk):
        oo, dd = o[r0 : r0 + rchunk, None, :], d[r0 : r0 + rchunk, None, :]
        ci = idx[r0 : r0 + rchunk].copy()
        ct = tmin[r0 : r0 + rchunk].copy()
        for s0, n in zip(starts, sizes):
            t, hit = ray_intersect_triangle(oo, dd, tri[None, s0 : s0 + n], epsilon=epsilon)
            if act is not None:
                hit = hit & act[None, s0 : s0 + n]
            t = np.where(hit, t, INF)
            bi = np.argmin(t, axis=-1).astype(np.int32)
            bt = np.min(t, axis=-1)
            bi = np.where(np.isinf(bt), np.int32(-1), bi) + np.int32(s0)
            cond = ct < bt
            ct = np.where(cond, ct, bt)
            ci = np.where(cond, ci, bi)
        fin = np.isfinite(ct)
        idx[r0 : r0 + rchunk] = np.where(fin, ci, np.int32(-1))
        tmin[r0 : r0 + rchunk] = np.where(fin, ct, INF)
    return idx.reshape(batch), tmin.reshape(batch)


def cartesian_to_spherical(xyz) -> np.ndarray:
    """``_utils.py:930-958``."""
    xyz = _f(xyz)
    with np.errstate(all="ignore"):
        r = np.sqrt(dot3(xyz, xyz))
        r = np.where(r == 0.0, F32(1.0), r)
        p = np.arccos(xyz[..., 2] / r)
        a = np.arctan2(xyz[..., 1], xyz[..., 0])
    return np.stack((r, p, a), axis=-1).astype(np.float32)


def spherical_to_cartesian_pa(p, a) -> np.ndarray:
    """``_utils.py:961-993`` for unit radius."""
    p, a = _f(p), _f(a)
    sp = np.sin(p)
    return np.stack((sp * np.cos(a), sp * np.sin(a), np.cos(p)), axis=-1).astype(np.float32)


def viewing_frustum(viewing_vertex, world_vertices, active_vertices=None) -> np.ndarray:
    """``_utils.py:821-927`` for a single viewing vertex → [2,3] = [[r,p,a]_min,[r,p,a]_max]."""
    vv, wv = _f(viewing_vertex), _f(world_vertices)
    rpa = cartesian_to_spherical(wv - vv[None, :])
    r, p, a = rpa[:, 0], rpa[:, 1], rpa[:, 2]
    w = np.ones(r.shape, bool) if active_vertices is None else np.asarray(active_vertices, bool)
    pi = F32(np.pi)
    two_pi = F32(2 * np.pi)

    def rmin(x, init):
        return np.min(x, where=w, initial=F32(init)).astype(np.float32)

    def rmax(x, init):
        return np.max(x, where=w, initial=F32(init)).astype(np.float32)

    r_min, r_max = rmin(r, np.inf), rmax(r, 0)
    p_min, p_max = rmin(p, pi), rmax(p, 0)
    a_min, a_max = rmin(a, pi), rmax(a, -pi)
    a0 = np.mod(a + two_pi, two_pi).astype(np.float32)
    a0_min, a0_max = rmin(a0, two_pi), rmax(a0, 0)
    a_width, a0_width = a_max - a_min, a0_max - a0_min
    if a_width > a0_width:
        a_min, a_max = a0_min, a0_max
    if min(a_width, a0_width) > F32(1.5) * pi:
        a_min, a_max = -pi, pi
    p0_min, p0_max = p_min, p_max
    if p_min == p_max:
        p_min = F32(0.0)
        p0_max = pi
    if (p_max - p_min) > (p0_max - p0_min):
        p_min, p_max = p0_min, p0_max
    return np.array([[r_min, p_min, a_min], [r_max, p_max, a_max]], dtype=np.float32)


def fibonacci_lattice(n: int, frustum=None) -> np.ndarray:
    """``_utils.py:416-490`` → [n,3] unit directions (fp32)."""
    i = np.arange(0.0, n, dtype=np.float32)
    inv_phi = 0.6180339887498949
    m1, m2 = 262144.0, 512.0
    inv_phi_m1 = F32((inv_phi * m1) % 1.0)
    inv_phi_m2 = F32((inv_phi * m2) % 1.0)
    q1 = np.floor(i / F32(m1))
    rem = i - q1 * F32(m1)
    q2 = np.floor(rem / F32(m2))
    r = rem - q2 * F32(m2)
    frac = np.mod(q1 * inv_phi_m1 + q2 * inv_phi_m2 + r * F32(inv_phi), F32(1.0))
    if frustum is not None:
        fr = _f(frustum)
        p_min, a_min = fr[0, 1], fr[0, 2]
        p_max, a_max = fr[1, 1], fr[1, 2]
        cmin, cmax = np.cos(p_min), np.cos(p_max)
        denom = F32(n - 1) if n > 1 else F32(1.0)
        cos_lat = cmin - (cmin - cmax) * (i / denom)
        with np.errstate(all="ignore"):
            lat = np.arccos(cos_lat)
        lon = a_min + (a_max - a_min) * frac
    else:
        lat = np.arccos(F32(1.0) - F32(2.0) * i / F32(n))
        lon = F32(2 * np.pi) * frac
    return spherical_to_cartesian_pa(lat.astype(np.float32), lon.astype(np.float32))


def visibility_directions(vertex, tri, active=None, num_rays=1_000_000) -> np.ndarray:
    """Ray directions of ``triangles_visible_from_vertex`` (``_utils.py:1668-1700``)."""
    tri = _f(tri)
    centers = tri.mean(axis=-2, keepdims=True, dtype=np.float32)
    world = np.concatenate((tri, centers), axis=-2).reshape(-1, 3)
    av = None if active is None else np.repeat(np.asarray(active, bool), 4)
    return fibonacci_lattice(num_rays, viewing_frustum(vertex, world, av))


def triangles_visible_from_vertex_dirs(vertex, ray_directions, tri, active=None, *, epsilon=None):
    """``_utils.py:1702-1772`` with the directions given: first hit per ray, scatter True."""
    tri = _f(tri)
    d = _f(ray_directions)
    o = np.broadcast_to(_f(vertex), d.shape)
    idx, _ = first_triangle_hit_by_ray(o, d, tri, active, batch_size=None, epsilon=epsilon)
    vis = np.zeros(tri.shape[0], dtype=bool)
    vis[idx[idx >= 0]] = True
    return vis


def triangles_visible_from_vertex(vertex, tri, active=None, num_rays=1_000_000, *, epsilon=None):
    """``_utils.py:1540-1772`` for one viewing vertex."""
    dirs = visibility_directions(vertex, tri, active, num_rays)
    return triangles_visible_from_vertex_dirs(vertex, dirs, tri, active, epsilon=epsilon)


# --------------------------------------------------------------------------------------
# a5-a8: image method
# --------------------------------------------------------------------------------------


def image_of_vertex_with_respect_to_mirror(vertex, mirror_vertex, mirror_normal) -> np.ndarray:
    """``_solver_image_method.py:68-79``: ``p - 2 ((p - m)·n) n``."""
    p, m, n = _f(vertex), _f(mirror_vertex), _f(mirror_normal)
    with np.errstate(all="ignore"):
        inc = p - m
        return (p - F32(2.0) * dot3(inc, n)[..., None] * n).astype(np.float32)


def intersection_of_ray_with_plane(ray_origin, ray_direction, plane_vertex, plane_normal):
    """``_solver_image_method.py:110-135``."""
    o, u, pv, n = _f(ray_origin), _f(ray_direction), _f(plane_vertex), _f(plane_normal)
    with np.errstate(all="ignore"):
        v = pv - o
        un = dot3(u, n)[..., None]
        vn = dot3(v, n)[..., None]
        parallel = un == 0.0
        un = np.where(parallel, F32(1.0), un)
        t = vn / un
        res = o + u * t
        return np.where(parallel & (vn != 0.0), INF, res).astype(np.float32)


def image_method(from_vertex, to_vertex, mirror_vertices, mirror_normals) -> np.ndarray:
    """``_solver_image_method.py:138-203, 344-363`` → [*batch, k, 3]."""
    fv, tv = _f(from_vertex), _f(to_vertex)
    mv, mn = _f(mirror_vertices), _f(mirror_normals)
    k = mv.shape[-2]
    batch = np.broadcast_shapes(fv.shape[:-1], tv.shape[:-1], mv.shape[:-2], mn.shape[:-2])
    if k == 0:
        return np.empty((*batch, 0, 3), dtype=np.float32)
    fv = np.broadcast_to(fv, (*batch, 3))
    tv = np.broadcast_to(tv, (*batch, 3))
    mv = np.broadcast_to(mv, (*batch, k, 3))
    mn = np.broadcast_to(mn, (*batch, k, 3))
    images = np.empty((*batch, k, 3), dtype=np.float32)
    prev = fv
    for i in range(k):
        prev = image_of_vertex_with_respect_to_mirror(prev, mv[..., i, :], mn[..., i, :])
        images[..., i, :] = prev
    paths = np.empty((*batch, k, 3), dtype=np.float32)
    prev = tv
    with np.errstate(all="ignore"):
        for i in range(k - 1, -1, -1):
            no_prev = np.isinf(prev)
            prev0 = np.where(no_prev, F32(0.0), prev)
            inter = intersection_of_ray_with_plane(
                prev0, images[..., i, :] - prev0, mv[..., i, :], mn[..., i, :]
            )
            prev = np.where(no_prev, INF, inter).astype(np.float32)
            paths[..., i, :] = prev
    return paths


def consecutive_vertices_are_on_same_side_of_mirror(vertices, mirror_vertices, mirror_normals):
    """``_solver_image_method.py:418-454`` → bool [*batch, k]."""
    v, mv, mn = _f(vertices), _f(mirror_vertices), _f(mirror_normals)
    if v.shape[-2] != mv.shape[-2] + 2:
        raise TypeError("vertices must hold num_mirrors + 2 points")
    with np.errstate(all="ignore"):
        d_prev = v[..., :-2, :] - mv
        d_next = v[..., 2:, :] - mv
        return np.sign(dot3(d_prev, mn)) == np.sign(dot3(d_next, mn))


# --------------------------------------------------------------------------------------
# a9: fused trace + validate
# --------------------------------------------------------------------------------------


def trace_path_candidates(
    vertices,
    triangles,
    tx,
    rx,
    path_candidates,
    *,
    mask=None,
    assume_quads=False,
    epsilon=None,
    hit_tol=None,
    min_len=None,
    stages=False,
    smoothing_factor=None,
):
    """``_solvers.py:514-770``, blockage with the pure-JAX any-hit.  With ``smoothing_factor`` the
    relaxed branch (``:599-713``): ``mask`` is a float in [0, 1] (NaN where the reference's is).

    Returns ``(vertices [Ntx,Nrx,C,k+2,3] f32, objects [Ntx,Nrx,C,k+2] i32, mask [Ntx,Nrx,C] bool)``;
    with ``stages=True`` also a dict of the five intermediate masks.
    """
    V = _f(vertices)
    tris = np.asarray(triangles, dtype=np.int64)
    tx, rx = _f(tx).reshape(-1, 3), _f(rx).reshape(-1, 3)
    cand = np.asarray(path_candidates, dtype=np.int64)
    C, k = cand.shape
    ntx, nrx = tx.shape[0], rx.shape[0]
    ml = F32(10.0 * EPS) if min_len is None else F32(min_len)
    tri_all = V[tris] if tris.size else np.empty((0, 3, 3), np.float32)
    normals = triangle_normals(tri_all) if tris.size else np.empty((0, 3), np.float32)

    q = 2 if assume_quads else 1
    if assume_quads:
        cand_x = np.repeat(cand, 2, axis=-1)
        cand_x[:, 1::2] += 1
    else:
        cand_x = cand
    tv = tri_all[cand_x].reshape(C, q * k, 3, 3)
    active_rays = None if mask is None else np.asarray(mask, bool)[cand_x].all(axis=-1)
    mirror_v = tv[:, ::q, 0, :]
    mirror_n = normals[cand].reshape(C, k, 3)

    if C == 0:
        full = np.empty((ntx, nrx, 0, k + 2, 3), dtype=np.float32)
    else:
        paths = image_method(tx[:, None, None, :], rx[None, :, None, :], mirror_v, mirror_n)
        full = assemble_path(tx[:, None, None, :], paths, rx[None, :, None, :])

    if smoothing_factor is not None:
        return _trace_smooth_tail(full, tv, mirror_v, mirror_n, tri_all, mask, active_rays, cand, assume_quads,
                                  epsilon, hit_tol, ml, smoothing_factor)
    with np.errstate(all="ignore"):
        ro = full[..., :-1, :]
        rd = np.diff(full, axis=-2)
        if assume_quads:
            h = ray_intersect_triangle(
                np.repeat(ro[..., :-1, :], 2, axis=-2),
                np.repeat(rd[..., :-1, :], 2, axis=-2),
                tv,
                epsilon=epsilon,
            )[1]
            inside = h.reshape(ntx, nrx, C, k, 2).any(axis=-1).all(axis=-1)
        else:
            inside = ray_intersect_triangle(
                ro[..., :-1, :], rd[..., :-1, :], tv, epsilon=epsilon
            )[1].all(axis=-1)
        same_side = consecutive_vertices_are_on_same_side_of_mirror(full, mirror_v, mirror_n).all(
            axis=-1
        )
        blocked = ray_intersect_any_triangle(
            ro, rd, tri_all, mask, hit_tol=hit_tol, epsilon=epsilon
        ).any(axis=-1)
        too_small = (dot3(rd, rd) < ml).any(axis=-1)
        finite = np.isfinite(full).all(axis=(-1, -2))
    full = np.where(finite[..., None, None], full, F32(0.0)).astype(np.float32)
    valid = inside & same_side & ~blocked & ~too_small & finite
    if active_rays is not None:
        valid = valid & active_rays[None, None, :]

    objects = np.empty((ntx, nrx, C, k + 2), dtype=np.int32)
    objects[..., 0] = np.arange(ntx, dtype=np.int32)[:, None, None]
    objects[..., 1:-1] = cand[None, None].astype(np.int32)
    objects[..., -1] = np.arange(nrx, dtype=np.int32)[None, :, None]
    if stages:
        return full, objects, valid, {
            "inside": inside,
            "same_side": same_side,
            "blocked": blocked,
            "too_small": too_small,
            "finite": finite,
        }
    return full, objects, valid


# --------------------------------------------------------------------------------------
# VJPs (Appendix B of SURVEY.md): closed-form reverse mode of the formulas above
# --------------------------------------------------------------------------------------


def first_hit_distance(vertices, triangles, o, d, faces):
    """``_mesh.py:226-255`` (``_differentiable_distance``): t at the given faces; inf on miss."""
    V = _f(vertices)
    tris = np.asarray(triangles, np.int64)
    o, d = _f(o), _f(d)
    faces = np.asarray(faces, np.int64)
    tri = V[tris][np.maximum(faces, 0)]
    with np.errstate(all="ignore"):
        e1 = tri[:, 1] - tri[:, 0]
        e2 = tri[:, 2] - tri[:, 0]
        a = dot3(cross3(d, e2), e1)
        a = np.where(a == 0.0, INF, a)
        f = F32(1.0) / a
        q = cross3(o - tri[:, 0], e1)
        t = f * dot3(q, e2)
    return np.where(faces != -1, t, INF).astype(np.float32)


def first_hit_vjp(vertices, triangles, o, d, faces, g_t):
    """Reverse mode of :func:`first_hit_distance` → ``(g_vertices [V,3], g_o [R,3], g_d [R,3])``.

    Mirrors what ``jax.vjp`` produces for ``_mesh.py:308-338``: rays with ``face == -1`` and rays
    with ``a == 0`` contribute zero.
    """
    V = _f(vertices)
    tris = np.asarray(triangles, np.int64)
    o, d, g_t = _f(o), _f(d), _f(g_t)
    faces = np.asarray(faces, np.int64)
    fi = np.maximum(faces, 0)
    tri = V[tris][fi]
    v0, v1, v2 = tri[:, 0], tri[:, 1], tri[:, 2]
    with np.errstate(all="ignore"):
        e1, e2 = v1 - v0, v2 - v0
        h = cross3(d, e2)
        a = dot3(h, e1)
        live = (faces != -1) & (a != 0.0)
        a_s = np.where(a == 0.0, INF, a)
        f = F32(1.0) / a_s
        s = o - v0
        q = cross3(s, e1)
        qe2 = dot3(q, e2)
        g = np.where(live, g_t, F32(0.0))
        g_f = g * qe2                      # t = f * (q·e2)
        g_qe2 = g * f
        g_q = g_qe2[:, None] * e2
        g_e2 = g_qe2[:, None] * q
        g_a = -g_f * f * f                 # f = 1/a
        g_h = g_a[:, None] * e1            # a = h·e1
        g_e1 = g_a[:, None] * h
        g_d = cross3(e2, g_h)              # h = d × e2
        g_e2 = g_e2 + cross3(g_h, d)
        g_s = cross3(e1, g_q)              # q = s × e1
        g_e1 = g_e1 + cross3(g_q, s)
        g_o = g_s
        g_v0 = -g_s - g_e1 - g_e2
    gV = np.zeros_like(V)
    ti = tris[fi]
    for col, gv in ((0, g_v0), (1, g_e1), (2, g_e2)):
        np.add.at(gV, ti[:, col], np.where(live[:, None], gv, F32(0.0)).astype(np.float32))
    return gV, g_o.astype(np.float32), g_d.astype(np.float32)


def image_method_vjp(from_vertex, to_vertex, mirror_vertices, mirror_normals, g_paths):
    """Reverse sweep of :func:`image_method` for flat batches (no broadcasting).

    Inputs ``[N,3],[N,3],[N,k,3],[N,k,3]``, cotangent ``[N,k,3]`` →
    ``(g_from [N,3], g_to [N,3], g_mv [N,k,3], g_mn [N,k,3])``.  Branches selected by ``where``
    (parallel ray / infinite previous point) contribute zero, like JAX's ``where`` gradient.
    """
    fv, tv = _f(from_vertex), _f(to_vertex)
    mv, mn, g = _f(mirror_vertices), _f(mirror_normals), _f(g_paths)
    N, k = mv.shape[0], mv.shape[1]
    with np.errstate(all="ignore"):
        images = np.empty((N, k + 1, 3), np.float32)
        images[:, 0] = fv
        for i in range(k):
            images[:, i + 1] = image_of_vertex_with_respect_to_mirror(
                images[:, i], mv[:, i], mn[:, i]
            )
        P = np.empty((N, k + 2, 3), np.float32)
        P[:, k + 1] = tv
        rec = [None] * k
        for i in range(k - 1, -1, -1):
            prev = P[:, i + 2]
            no_prev = np.isinf(prev)
            prev0 = np.where(no_prev, F32(0.0), prev)
            dirv = images[:, i + 1] - prev0
            w = mv[:, i] - prev0
            un = dot3(dirv, mn[:, i])
            vn = dot3(w, mn[:, i])
            par = un == 0.0
            un_s = np.where(par, F32(1.0), un)
            t = vn / un_s
            res = prev0 + dirv * t[:, None]
            res = np.where((par & (vn != 0.0))[:, None], INF, res)
            P[:, i + 1] = np.where(no_prev, INF, res)
            rec[i] = (no_prev, prev0, dirv, w, un_s, par, vn, t)
        g_mv = np.zeros_like(mv)
        g_mn = np.zeros_like(mn)
        g_img = np.zeros((N, k + 1, 3), np.float32)
        g_next = np.zeros((N, 3), np.float32)  # cotangent of P_{i+1} flowing from P_i, i ascending
        # paths[i] = P[i+1]; reverse of the backward scan runs i = 0..k-1
        carry = np.zeros((N, 3), np.float32)
        for i in range(k):
            no_prev, prev0, dirv, w, un_s, par, vn, t = rec[i]
            gi = g[:, i] + carry
            gi = np.where(no_prev, F32(0.0), gi)                       # where(no_prev, inf, ·)
            gi = np.where((par & (vn != 0.0))[:, None], F32(0.0), gi)  # where(parallel&vn!=0, inf, ·)
            g_prev0 = gi.copy()
            g_t = dot3(gi, dirv)
            g_dir = gi * t[:, None]
            g_vn = g_t / un_s
            g_un = np.where(par, F32(0.0), -g_t * t / un_s)            # un replaced by 1 if parallel
            g_dir = g_dir + g_un[:, None] * mn[:, i]
            g_mn[:, i] += g_un[:, None] * dirv + g_vn[:, None] * w
            g_w = g_vn[:, None] * mn[:, i]
            g_mv[:, i] += g_w
            g_img[:, i + 1] += g_dir
            g_prev0 = g_prev0 - g_dir - g_w
            carry = np.where(no_prev, F32(0.0), g_prev0).astype(np.float32)
        g_to = carry
        for i in range(k - 1, -1, -1):
            gI = g_img[:, i + 1]
            inc = images[:, i] - mv[:, i]
            c = dot3(inc, mn[:, i])
            g_c = F32(-2.0) * dot3(gI, mn[:, i])
            g_img[:, i] += gI + g_c[:, None] * mn[:, i]
            g_mv[:, i] -= g_c[:, None] * mn[:, i]
            g_mn[:, i] += g_c[:, None] * inc - F32(2.0) * c[:, None] * gI
        del g_next
    return g_img[:, 0].copy(), g_to, g_mv, g_mn


def trace_vjp(vertices, triangles, tx, rx, path_candidates, g_out_vertices):
    """Reverse mode of :func:`trace_path_candidates`'s ``vertices`` output w.r.t. ``tx``, ``rx`` and
    the mesh ``vertices`` (``_solvers.py:535-586, 696-699`` under JAX autodiff): through the image
    method, the mirror-vertex gather and ``Mesh.normals``; non-finite paths get zero gradient.
    """
    V = _f(vertices)
    tris = np.asarray(triangles, np.int64)
    tx, rx = _f(tx).reshape(-1, 3), _f(rx).reshape(-1, 3)
    cand = np.asarray(path_candidates, np.int64)
    C, k = cand.shape
    ntx, nrx = tx.shape[0], rx.shape[0]
    g = _f(g_out_vertices).reshape(ntx, nrx, C, k + 2, 3)
    tri = V[tris[cand]]                                      # [C,k,3,3]
    v0, v1, v2 = tri[..., 0, :], tri[..., 1, :], tri[..., 2, :]
    with np.errstate(all="ignore"):
        A, B = v1 - v0, v2 - v1
        Nv = cross3(A, B)
        L = np.sqrt(dot3(Nv, Nv))
        Ls = np.where(L == 0.0, F32(1.0), L)
        n = Nv / Ls[..., None]
        N = ntx * nrx * C
        fv = np.broadcast_to(tx[:, None, None, :], (ntx, nrx, C, 3)).reshape(N, 3)
        tv = np.broadcast_to(rx[None, :, None, :], (ntx, nrx, C, 3)).reshape(N, 3)
        mv = np.broadcast_to(v0[None, None], (ntx, nrx, C, k, 3)).reshape(N, k, 3)
        mn = np.broadcast_to(n[None, None], (ntx, nrx, C, k, 3)).reshape(N, k, 3)
        paths = image_method(fv, tv, mv, mn) if k > 0 else np.empty((N, 0, 3), np.float32)
        finite = np.isfinite(paths).all(axis=(-1, -2)) & np.isfinite(fv).all(-1) & np.isfinite(tv).all(-1)
        gf = np.where(finite[:, None, None], g.reshape(N, k + 2, 3), F32(0.0)).astype(np.float32)
        if k > 0:
            g_from, g_to, g_mv, g_mn = image_method_vjp(fv, tv, mv, mn, gf[:, 1:-1])
            for arr in (g_from, g_to, g_mv, g_mn):
                arr[~finite] = 0.0
        else:
            g_from = np.zeros((N, 3), np.float32)
            g_to = np.zeros((N, 3), np.float32)
            g_mv = g_mn = np.zeros((N, 0, 3), np.float32)
        g_from = g_from + gf[:, 0]
        g_to = g_to + gf[:, -1]
        g_tx = g_from.reshape(ntx, nrx * C, 3).sum(axis=1)
        g_rx = g_to.reshape(ntx, nrx, C, 3).sum(axis=(0, 2))
        g_mv = g_mv.reshape(ntx * nrx, C, k, 3).sum(axis=0)
        g_mn = g_mn.reshape(ntx * nrx, C, k, 3).sum(axis=0)
        proj = dot3(n, g_mn)[..., None]
        gN = np.where((L == 0.0)[..., None], g_mn, (g_mn - n * proj) / Ls[..., None])
        gA, gB = cross3(B, gN), cross3(gN, A)
    gV = np.zeros_like(V)
    ti = tris[cand]                                           # [C,k,3]
    np.add.at(gV, ti[..., 0], (g_mv - gA).astype(np.float32))
    np.add.at(gV, ti[..., 1], (gA - gB).astype(np.float32))
    np.add.at(gV, ti[..., 2], gB.astype(np.float32))
    return g_tx.astype(np.float32), g_rx.astype(np.float32), gV


# ------------------------------------------------------------------------------------------------
# N3: shooting-and-bouncing rays (SBRPathLauncher) and the multipath lifetime map
# ------------------------------------------------------------------------------------------------


def sbr_launch_rays(tri, tx, rx, num_rays: int):
    """``SBRPathLauncher.launch_rays`` (``_solvers.py:1202-1226``): frustum over every triangle vertex
    and receiver, Fibonacci lattice inside it → ``(origins, directions) [num_tx, num_rays, 3]``."""
    tx = _f(tx).reshape(-1, 3)
    world = np.concatenate((_f(tri).reshape(-1, 3), _f(rx).reshape(-1, 3)), axis=0)
    dirs = np.stack([fibonacci_lattice(num_rays, frustum=viewing_frustum(t, world)) for t in tx])
    return np.broadcast_to(tx[:, None, :], dirs.shape).copy(), dirs.astype(np.float32)


def sbr_launch_paths(vertices, triangles, tx, rx, ray_directions, order: int, *, max_dist=1e-3,
                     epsilon=None, mask=None, first_hit=None):
    """``SBRPathLauncher.launch_paths`` (``_solvers.py:358-491``) with the ray directions given.

    ``lax.scan`` over ``order + 1`` bounces of: nearest hit → ``filter_rays`` (``:320-356``) →
    ``bounce_rays`` (``:279-318``).  The nearest hit uses the pure-JAX ``first_triangle_hit_by_ray``
    definition (``_utils.py:1775-1960``), like every other kernel here.  Returns
    ``(path_candidates [num_tx, num_rays, order] i32, vertices [num_tx, num_rays, order, 3] f32,
    masks [num_tx, num_rx, num_rays, order + 1] bool)`` — the quantities the reference assembles
    into ``LaunchedPaths`` (``:446-491``).
    """
    tri = triangle_vertices(vertices, triangles)
    nrm = triangle_normals(tri)
    tx, rx = _f(tx).reshape(-1, 3), _f(rx).reshape(-1, 3)
    d = _f(ray_directions).copy()
    ntx, nrays = d.shape[0], d.shape[1]
    o = np.broadcast_to(tx[:, None, :], d.shape).astype(np.float32).copy()
    valid = np.ones((ntx, nrays), bool)
    first_hit = first_hit or (lambda oo, dd: first_triangle_hit_by_ray(oo, dd, tri, mask, epsilon=epsilon))
    cands, verts, masks = [], [], []
    max_dist = F32(max_dist)
    for _ in range(order + 1):
        faces, t_hit = first_hit(o.reshape(-1, 3), d.reshape(-1, 3))
        faces, t_hit = faces.reshape(ntx, nrays), t_hit.reshape(ntx, nrays).astype(np.float32)
        # filter_rays
        v = rx[None, :, None, :] - o[:, None, :, :]
        c = cross3(d[:, None, :, :], v)
        dist2 = (c[..., 0] * c[..., 0] + c[..., 1] * c[..., 1]) + c[..., 2] * c[..., 2]
        t_rx = dot3(d[:, None, :, :], v)
        with np.errstate(invalid="ignore"):
            near = (t_rx > 0) & (t_rx < t_hit[:, None, :]) & valid[:, None, :] & (dist2 < max_dist)
        masks.append(near)
        # bounce_rays
        inside = np.isfinite(t_hit)
        valid = valid & inside
        t = np.where(inside, t_hit, F32(0))
        o = (o + t[..., None] * d).astype(np.float32)
        n = nrm[faces]  # jnp.take wraps the -1 of a miss to the last triangle
        k = F32(2.0) * dot3(d, n)
        d = (d - k[..., None] * n).astype(np.float32)
        cands.append(faces.astype(np.int32))
        verts.append(o.copy())
    path_candidates = np.moveaxis(np.stack(cands[:-1]), 0, -1) if order > 0 else np.zeros((ntx, nrays, 0), np.int32)
    vertices_out = np.moveaxis(np.stack(verts[:-1]), 0, -2) if order > 0 else np.zeros((ntx, nrays, 0, 3), np.float32)
    return path_candidates, vertices_out, np.moveaxis(np.stack(masks), 0, -1)


MLM_MAGIC_1, MLM_MAGIC_2, MLM_MAGIC_3 = np.uint32(0x9E3779B9), np.uint32(0x045D9F3B), np.uint32(0x811C9DC5)


def mlm_combine_hashes(h1, h2):
    """``combine_hashes`` (``_scene.py:65-69``), uint32 wrap-around arithmetic."""
    h1, h2 = np.asarray(h1, np.uint32), np.asarray(h2, np.uint32)
    with np.errstate(over="ignore"):
        return h1 ^ (h2 + MLM_MAGIC_1 + (h1 << np.uint32(6)) + (h1 >> np.uint32(2)))


def mlm_hash_int(x):
    """``hash_int`` (``_scene.py:72-78``)."""
    x = np.asarray(x, np.uint32)
    with np.errstate(over="ignore"):
        x = ((x >> np.uint32(16)) ^ x) * MLM_MAGIC_2
        x = ((x >> np.uint32(16)) ^ x) * MLM_MAGIC_2
    return (x >> np.uint32(16)) ^ x


def compute_tx_mlm(vertices, triangles, tx, ray_directions, *, max_order, min_order, assume_quads, dim_x,
                   dim_y, receiver_height, min_x, max_x, min_y, max_y, mask=None, epsilon=1e-4, first_hit=None):
    """The loop of ``_compute_tx_mlm_kernel`` (``_scene.py:81-171``) with the ray directions given and
    the nearest hit taken from ``first_triangle_hit_by_ray`` (the reference asks Warp's BVH,
    third-party arithmetic).  → ``[num_tx, dim_x, dim_y] uint32``."""
    tri = triangle_vertices(vertices, triangles)
    nrm = triangle_normals(tri)
    tx = _f(tx).reshape(-1, 3)
    d = _f(ray_directions).copy()
    ntx, nrays = d.shape[0], d.shape[1]
    qo = np.broadcast_to(tx[:, None, :], d.shape).astype(np.float32).copy()
    first_hit = first_hit or (lambda oo, dd: first_triangle_hit_by_ray(oo, dd, tri, mask))
    eps = F32(epsilon)
    h = np.full((ntx, nrays), MLM_MAGIC_3, np.uint32)
    alive = np.ones((ntx, nrays), bool)
    out = np.zeros((ntx, dim_x, dim_y), np.uint32)
    rh, x0, x1, y0, y1 = F32(receiver_height), F32(min_x), F32(max_x), F32(min_y), F32(max_y)
    dx, dy = (x1 - x0) / F32(dim_x), (y1 - y0) / F32(dim_y)
    itx = np.broadcast_to(np.arange(ntx)[:, None], (ntx, nrays))
    for it in range(max_order + 1):
        faces, res_t = first_hit(qo.reshape(-1, 3), d.reshape(-1, 3))
        faces, res_t = faces.reshape(ntx, nrays), res_t.reshape(ntx, nrays).astype(np.float32)
        hit = faces >= 0
        t_hit = np.where(hit, res_t + eps if it > 0 else res_t, INF).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            u = ((rh - qo[..., 2]) / d[..., 2]).astype(np.float32)
            P = (qo + d * u[..., None]).astype(np.float32)
            ok = alive & (np.abs(d[..., 2]) > F32(1e-6)) & (u > 0) & (u < t_hit) & (it >= min_order)
            ok &= (P[..., 0] >= x0) & (P[..., 0] <= x1) & (P[..., 1] >= y0) & (P[..., 1] <= y1)
            ix = np.clip(np.floor(((P[..., 0] - x0) / dx).astype(np.float32)), 0, dim_x - 1)
            iy = np.clip(np.floor(((P[..., 1] - y0) / dy).astype(np.float32)), 0, dim_y - 1)
        sel = np.nonzero(ok)
        np.bitwise_or.at(out, (itx[sel], ix[sel].astype(np.int64), iy[sel].astype(np.int64)), h[sel])
        alive = alive & hit
        o = (qo + d * res_t[..., None]).astype(np.float32)
        n = nrm[np.where(hit, faces, 0)]
        k = F32(2.0) * dot3(d, n)
        d_new = (d - k[..., None] * n).astype(np.float32)
        h = np.where(alive, mlm_combine_hashes(h, mlm_hash_int((faces // 2 if assume_quads else faces).astype(np.uint32))), h)
        qo = np.where(alive[..., None], (o + d_new * eps).astype(np.float32), qo)
        d = np.where(alive[..., None], d_new, d)
    return out


# ------------------------------------------------------------------------------------------------
# N4 (forward): smoothed primitives
# ------------------------------------------------------------------------------------------------


def smoothing_function(x, smoothing_factor=1.0):
    """``jax.nn.sigmoid(x * alpha)`` (``differt/src/differt/utils.py:70-89``), float32."""
    with np.errstate(over="ignore", invalid="ignore"):
        z = (_f(x) * F32(smoothing_factor)).astype(np.float32)
        return (F32(1.0) / (F32(1.0) + np.exp(-z).astype(np.float32))).astype(np.float32)


def ray_intersect_triangle_smooth(ray_origins, ray_directions, tri, *, epsilon=None, smoothing_factor=1.0):
    """``ray_intersect_triangle(..., smoothing_factor=alpha)`` (``_utils.py:1263-1322``) → ``(t, hit)``
    with ``hit`` a float in [0, 1]."""
    eps = F32(10 * EPS if epsilon is None else epsilon)
    o, d, tri = _f(ray_origins), _f(ray_directions), _f(tri)
    with np.errstate(all="ignore"):
        v0, v1, v2 = tri[..., 0, :], tri[..., 1, :], tri[..., 2, :]
        e1, e2 = v1 - v0, v2 - v0
        h = cross3(d, e2)
        a = dot3(h, e1)
        a = np.where(a == 0, INF, a).astype(np.float32)
        hit = smoothing_function(np.abs(a) - eps, smoothing_factor)
        f = (F32(1.0) / a).astype(np.float32)
        s = o - v0
        u = f * dot3(s, h)
        one = np.ones_like(hit)
        hit = np.minimum.reduce([hit, smoothing_function(u - F32(0), smoothing_factor),
                                 smoothing_function(F32(1) - u, smoothing_factor), one])
        q = cross3(s, e1)
        v = f * dot3(q, d)
        hit = np.minimum.reduce([hit, smoothing_function(v - F32(0), smoothing_factor),
                                 smoothing_function(F32(1) - (u + v), smoothing_factor), one])
        t = f * dot3(q, e2)
        hit = np.minimum(hit, smoothing_function(t - eps, smoothing_factor))
    return t.astype(np.float32), hit.astype(np.float32)


def ray_intersect_any_triangle_smooth(ray_origins, ray_directions, tri, active=None, *, epsilon=None,
                                      hit_tol=None, smoothing_factor=1.0):
    """``ray_intersect_any_triangle(..., smoothing_factor=alpha)`` (``_utils.py:1452-1476``): sum over the
    active triangles of ``min(hit, sigmoid((1 - hit_tol - t) alpha))``, clipped at 1."""
    thr = F32(1.0) - F32(100 * EPS if hit_tol is None else hit_tol)
    o, d, tri = _f(ray_origins), _f(ray_directions), _f(tri)
    t, hit = ray_intersect_triangle_smooth(o[..., None, :], d[..., None, :], tri, epsilon=epsilon,
                                           smoothing_factor=smoothing_factor)
    term = np.minimum(hit, smoothing_function(thr - t, smoothing_factor))
    if active is not None:
        term = np.where(np.asarray(active, bool), term, F32(0))
    return np.minimum(term.sum(axis=-1, dtype=np.float32), F32(1.0)).astype(np.float32)


def _trace_smooth_tail(full, tv, mirror_v, mirror_n, tri_all, mask, active_rays, cand, assume_quads, epsilon,
                       hit_tol, ml, alpha):
    """Steps 3.1-3.5 of ``_trace_path_candidates`` with smoothing (``_solvers.py:599-713``): AND → min,
    OR → max, NOT x → 1 - x; the hard ``is_finite`` enters the min as 0 / 1."""
    ntx, nrx, C = full.shape[:3]
    k = cand.shape[1]
    with np.errstate(all="ignore"):
        ro = full[..., :-1, :]
        rd = np.diff(full, axis=-2)
        nmin = lambda x, init: np.minimum(x.min(axis=-1, initial=np.inf), F32(init)) if x.shape[-1] else np.full(x.shape[:-1], F32(init))  # noqa: E731
        nmax = lambda x, init: np.maximum(x.max(axis=-1, initial=-np.inf), F32(init)) if x.shape[-1] else np.full(x.shape[:-1], F32(init))  # noqa: E731
        if assume_quads:
            h = ray_intersect_triangle_smooth(np.repeat(ro[..., :-1, :], 2, axis=-2),
                                              np.repeat(rd[..., :-1, :], 2, axis=-2), tv, epsilon=epsilon,
                                              smoothing_factor=alpha)[1]
            inside = nmin(nmax(h.reshape(ntx, nrx, C, k, 2), 0.0), 1.0)
        else:
            inside = nmin(ray_intersect_triangle_smooth(ro[..., :-1, :], rd[..., :-1, :], tv, epsilon=epsilon,
                                                        smoothing_factor=alpha)[1], 1.0)
        same = nmin(consecutive_vertices_are_on_same_side_of_mirror_smooth(full, mirror_v, mirror_n, alpha), 1.0)
        blocked = nmax(ray_intersect_any_triangle_smooth(ro, rd, tri_all, mask, epsilon=epsilon, hit_tol=hit_tol,
                                                         smoothing_factor=alpha), 0.0)
        too_small = nmax(smoothing_function(ml - dot3(rd, rd), alpha), 0.0)
        finite = np.isfinite(full).all(axis=(-1, -2))
        full = np.where(finite[..., None, None], full, F32(0.0)).astype(np.float32)
        soft = np.stack([inside, same, F32(1) - blocked, F32(1) - too_small, finite.astype(np.float32)], axis=-1)
        soft = soft.min(axis=-1)  # np.min propagates NaN like jnp.min
        if active_rays is not None:
            soft = soft * active_rays[None, None, :].astype(np.float32)
    objects = np.empty((ntx, nrx, C, k + 2), dtype=np.int32)
    objects[..., 0] = np.arange(ntx, dtype=np.int32)[:, None, None]
    objects[..., 1:-1] = cand[None, None].astype(np.int32)
    objects[..., -1] = np.arange(nrx, dtype=np.int32)[None, :, None]
    return full, objects, soft.astype(np.float32)


def consecutive_vertices_are_on_same_side_of_mirror_smooth(vertices, mirror_vertices, mirror_normals,
                                                           smoothing_factor=1.0):
    """``_solver_image_method.py:440-454`` with smoothing: ``sigmoid(sign(dot_prev) sign(dot_next) alpha)``."""
    v, mv, mn = _f(vertices), _f(mirror_vertices), _f(mirror_normals)
    dp = dot3(v[..., :-2, :] - mv, mn)
    dn = dot3(v[..., 2:, :] - mv, mn)
    return smoothing_function(np.sign(dp) * np.sign(dn), smoothing_factor)
