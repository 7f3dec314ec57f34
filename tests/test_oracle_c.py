"""The C oracle must equal the (golden-pinned) NumPy oracle bit for bit (CPU only)."""

from __future__ import annotations

import numpy as np
import pytest

from differt_b200 import scenes
from oracle import c_oracle as co
from oracle import differt_oracle as orc


def _rays(rng, n, lo=-20.0, hi=120.0):
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    e = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    o[:, 2] = np.abs(o[:, 2]) * 0.4
    e[:, 2] = np.abs(e[:, 2]) * 0.4
    return o, e - o


def test_mt_elementwise(rng):
    o = rng.uniform(size=(1000, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, size=(1000, 3)).astype(np.float32)
    tri = rng.uniform(size=(1000, 3, 3)).astype(np.float32)
    t0, h0 = orc.ray_intersect_triangle(o, d, tri)
    t1, h1 = co.ray_intersect_triangle(o, d, tri)
    np.testing.assert_array_equal(t0.view(np.uint32), t1.view(np.uint32))
    np.testing.assert_array_equal(h0, h1)


@pytest.mark.parametrize("use_mask", [False, True])
@pytest.mark.parametrize("early_exit", [False, True])
def test_any_hit(rng, use_mask, early_exit):
    v, t = scenes.urban_grid(4, 4)
    tri = orc.triangle_vertices(v, t)
    o, d = _rays(rng, 3000)
    act = rng.uniform(size=tri.shape[0]) > 0.5 if use_mask else None
    a = orc.ray_intersect_any_triangle(o, d, tri, act)
    b = co.ray_intersect_any_triangle(o, d, tri, act, early_exit=early_exit)
    np.testing.assert_array_equal(a, b)
    assert 0.05 < a.mean() < 0.95


@pytest.mark.parametrize("batch_size", [None, 7, 512])
def test_first_hit(rng, batch_size):
    v, t = scenes.urban_grid(3, 3)
    tri = orc.triangle_vertices(v, t)
    o, d = _rays(rng, 2000, -20, 80)
    act = rng.uniform(size=tri.shape[0]) > 0.2
    i0, t0 = orc.first_triangle_hit_by_ray(o, d, tri, act, batch_size=batch_size)
    i1, t1 = co.first_triangle_hit_by_ray(o, d, tri, act, batch_size=batch_size)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(t0.view(np.uint32), t1.view(np.uint32))


def test_first_hit_tie_rule_cross_batch():
    # two coincident triangles: with batch_size=1 the later batch wins the exact tie,
    # with one batch the first index wins (reference _utils.py:1865-1868, 1886)
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]] * 2, np.float32)
    o = np.array([[0.2, 0.2, 1.0]], np.float32)
    d = np.array([[0.0, 0.0, -1.0]], np.float32)
    for mod in (orc, co):
        assert mod.first_triangle_hit_by_ray(o, d, tri, batch_size=1)[0][0] == 1
        assert mod.first_triangle_hit_by_ray(o, d, tri, batch_size=None)[0][0] == 0


def test_visibility(rng):
    v, t = scenes.box(with_top=True)
    tri = orc.triangle_vertices(v, t)
    vertex = np.array([2.0, 2.0, 0.0], np.float32)
    dirs = orc.visibility_directions(vertex, tri, None, 5000)
    a = orc.triangles_visible_from_vertex_dirs(vertex, dirs, tri)
    b = co.triangles_visible_from_vertex_dirs(vertex, dirs, tri)[0]
    np.testing.assert_array_equal(a, b)
    assert a.sum() == 4


def test_image_method(rng):
    N, k = 500, 4
    fv = rng.uniform(size=(N, 3)).astype(np.float32)
    tv = rng.uniform(size=(N, 3)).astype(np.float32)
    mv = rng.uniform(size=(N, k, 3)).astype(np.float32)
    mn = orc.normalize(rng.uniform(-1, 1, size=(N, k, 3)).astype(np.float32))[0]
    # rows ::7: all mirrors are the plane z=0 and from/to sit at z=1, so the k-th image is at z=1
    # too, the last ray is parallel to its mirror (un == 0, vn != 0) and inf propagates backwards
    mv[::7] = 0.0
    mn[::7] = np.array([0.0, 0.0, 1.0], np.float32)
    fv[::7, 2] = 1.0
    tv[::7, 2] = 1.0
    a = orc.image_method(fv, tv, mv, mn)
    b = co.image_method(fv, tv, mv, mn)
    np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.isinf(a).any()


@pytest.mark.parametrize("assume_quads", [False, True])
@pytest.mark.parametrize("use_mask", [False, True])
def test_trace_two_buildings(two_buildings, kats, assume_quads, use_mask):
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    rx2 = np.stack((rx, rx + np.float32(0.5)))
    n = t.shape[0] // 2 if assume_quads else t.shape[0]
    cand = scenes.complete_graph_candidates(n, 2) * (2 if assume_quads else 1)
    mask = (np.random.default_rng(3).uniform(size=t.shape[0]) > 0.2) if use_mask else None
    a = orc.trace_path_candidates(v, t, tx, rx2, cand, mask=mask, assume_quads=assume_quads, stages=True)
    b = co.trace_path_candidates(v, t, tx, rx2, cand, mask=mask, assume_quads=assume_quads, stages=True)
    np.testing.assert_array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    for key in a[3]:
        np.testing.assert_array_equal(a[3][key], b[3][key], err_msg=key)


def test_trace_urban_random_candidates(rng):
    v, t = scenes.urban_grid(3, 3)
    tx = np.array([[30.0, 30.0, 50.0]], np.float32)
    rx = scenes.receivers_grid(v, 3)
    cand = scenes.sampled_candidates(t.shape[0], 3, 200)
    a = orc.trace_path_candidates(v, t, tx, rx, cand, stages=True)
    b = co.trace_path_candidates(v, t, tx, rx, cand, stages=True, early_exit=True)
    np.testing.assert_array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    np.testing.assert_array_equal(a[2], b[2])
    for key in a[3]:
        np.testing.assert_array_equal(a[3][key], b[3][key], err_msg=key)
