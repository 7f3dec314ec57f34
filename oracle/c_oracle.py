"""ctypes front end of ``oracle/oracle.c`` (TEST INFRASTRUCTURE — see the header of that file)."""

from __future__ import annotations

import ctypes as C

import numpy as np

from .build import build_oracle

_lib = None
EPS = float(np.finfo(np.float32).eps)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build_oracle()))
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a) -> np.ndarray | None:
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(C.c_int(n))


def ray_intersect_triangle(o, d, tri, *, epsilon=None):
    o, d, tri = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3), _f32(tri).reshape(-1, 3, 3)
    n = o.shape[0]
    t = np.empty(n, np.float32)
    hit = np.empty(n, np.uint8)
    eps = 10 * EPS if epsilon is None else float(epsilon)
    lib().orc_ray_intersect_triangle(
        C.c_int64(n), _p(o), _p(d), _p(tri), C.c_float(eps), _p(t), _p(hit)
    )
    return t, hit.astype(bool)


def ray_intersect_any_triangle(o, d, tri, active=None, *, hit_tol=None, epsilon=None, early_exit=False):
    o, d, tri = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3), _f32(tri).reshape(-1, 3, 3)
    act = _u8(active)
    out = np.empty(o.shape[0], np.uint8)
    eps = 10 * EPS if epsilon is None else float(epsilon)
    tol = 100 * EPS if hit_tol is None else float(hit_tol)
    lib().orc_ray_intersect_any_triangle(
        C.c_int64(o.shape[0]), C.c_int64(tri.shape[0]), _p(o), _p(d), _p(tri), _p(act),
        C.c_float(eps), C.c_float(tol), C.c_int(int(early_exit)), _p(out),
    )
    return out.astype(bool)


def first_triangle_hit_by_ray(o, d, tri, active=None, *, batch_size=512, epsilon=None):
    o, d, tri = _f32(o).reshape(-1, 3), _f32(d).reshape(-1, 3), _f32(tri).reshape(-1, 3, 3)
    act = _u8(active)
    idx = np.empty(o.shape[0], np.int32)
    t = np.empty(o.shape[0], np.float32)
    eps = 10 * EPS if epsilon is None else float(epsilon)
    lib().orc_first_triangle_hit_by_ray(
        C.c_int64(o.shape[0]), C.c_int64(tri.shape[0]), _p(o), _p(d), _p(tri), _p(act),
        C.c_float(eps), C.c_int64(0 if batch_size is None else int(batch_size)), _p(idx), _p(t),
    )
    return idx, t


def triangles_visible_from_vertex_dirs(vertex, dirs, tri, active=None, *, epsilon=None):
    vertex = _f32(vertex).reshape(-1, 3)
    B = vertex.shape[0]
    dirs = _f32(dirs).reshape(B, -1, 3)
    tri = _f32(tri).reshape(-1, 3, 3)
    act = _u8(active)
    out = np.empty((B, tri.shape[0]), np.uint8)
    eps = 10 * EPS if epsilon is None else float(epsilon)
    lib().orc_triangles_visible_from_vertex(
        C.c_int64(B), C.c_int64(dirs.shape[1]), C.c_int64(tri.shape[0]), _p(vertex), _p(dirs),
        _p(tri), _p(act), C.c_float(eps), _p(out),
    )
    return out.astype(bool)


def image_method(from_v, to_v, mv, mn):
    mv, mn = _f32(mv), _f32(mn)
    N, k = mv.shape[0], mv.shape[1]
    from_v, to_v = _f32(from_v).reshape(N, 3), _f32(to_v).reshape(N, 3)
    out = np.empty((N, k, 3), np.float32)
    lib().orc_image_method(C.c_int64(N), C.c_int(k), _p(from_v), _p(to_v), _p(mv), _p(mn), _p(out))
    return out


def trace_path_candidates(
    vertices, triangles, tx, rx, cand, *, mask=None, assume_quads=False, epsilon=None,
    hit_tol=None, min_len=None, early_exit=False, stages=False, count_tests=False,
):
    V = _f32(vertices)
    tris = np.ascontiguousarray(triangles, np.int32)
    tx, rx = _f32(tx).reshape(-1, 3), _f32(rx).reshape(-1, 3)
    cand = np.ascontiguousarray(cand, np.int32)
    Cn, k = cand.shape
    ntx, nrx = tx.shape[0], rx.shape[0]
    P = ntx * nrx * Cn
    ov = np.empty((ntx, nrx, Cn, k + 2, 3), np.float32)
    oo = np.empty((ntx, nrx, Cn, k + 2), np.int32)
    om = np.empty((ntx, nrx, Cn), np.uint8)
    st = np.empty((P, 5), np.uint8) if stages else None
    nt = C.c_int64(0)
    eps = 10 * EPS if epsilon is None else float(epsilon)
    tol = 100 * EPS if hit_tol is None else float(hit_tol)
    ml = 10 * EPS if min_len is None else float(min_len)
    lib().orc_trace_path_candidates(
        C.c_int64(V.shape[0]), C.c_int64(tris.shape[0]), _p(V), _p(tris), _p(_u8(mask)),
        C.c_int(int(assume_quads)), C.c_int64(ntx), _p(tx), C.c_int64(nrx), _p(rx), C.c_int64(Cn),
        C.c_int(k), _p(cand), C.c_float(eps), C.c_float(tol), C.c_float(ml),
        C.c_int(int(early_exit)), _p(ov), _p(oo), _p(om), _p(st), C.byref(nt),
    )
    res = [ov, oo, om.astype(bool)]
    if stages:
        s = st.reshape(ntx, nrx, Cn, 5).astype(bool)
        res.append({
            "inside": s[..., 0], "same_side": s[..., 1], "blocked": s[..., 2],
            "too_small": s[..., 3], "finite": s[..., 4],
        })
    if count_tests:
        res.append(int(nt.value))
    return tuple(res)
