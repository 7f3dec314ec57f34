"""Pin the NumPy oracle to the reference's own known-answer tests (CPU only).

Each test cites the reference test it restates (paths relative to
``/root/reference/differt/tests/geometry/``); the vectors come from
``tests/golden/reference_kats.json`` (written by ``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import numpy as np
import pytest

from differt_b200 import scenes
from oracle import differt_oracle as orc


def test_ray_intersect_triangle_hit_table(kats):  # test_utils.py:555-577
    k = kats["ray_intersect_triangle_hit_table"]
    tri = np.array([k["triangle"]], dtype=np.float32)
    for case in k["cases"]:
        o = np.array(case["orig"], np.float32)
        d = np.array(case["dest"], np.float32) - o
        t, hit = orc.ray_intersect_triangle(o, d, tri)
        assert bool(((t < 1.0) & hit)[0]) == case["expected"]


def test_ray_intersect_triangle_t_and_hit(kats):  # test_utils.py:580-606
    k = kats["ray_intersect_triangle_t_and_hit"]
    o = np.array(k["ray_origin"], np.float32)
    d = np.array(k["ray_directions"], np.float32)
    tri = np.array(k["triangles"], np.float32)
    t, hit = orc.ray_intersect_triangle(o[None, None, :], d[:, None, :], tri)
    np.testing.assert_array_equal(t, np.array(k["expected_t"], np.float32))
    np.testing.assert_array_equal(hit, np.array(k["expected_hit"]))


def test_ray_intersect_triangle_hit_implies_positive_t(rng):  # test_utils.py:609-628
    o = rng.uniform(size=(15, 5, 3)).astype(np.float32)
    d = rng.uniform(size=(15, 5, 3)).astype(np.float32)
    tri = rng.uniform(size=(5, 3, 3)).astype(np.float32)
    t, hit = orc.ray_intersect_triangle(o, d, tri)
    assert np.where(hit, t > 0.0, True).all()


@pytest.mark.parametrize("epsilon", [None, 1e-6, 1e-2])
@pytest.mark.parametrize("hit_tol", [None, 0.0, 1e-3, 0.5, -0.5])
@pytest.mark.parametrize("use_mask", [False, True])
def test_any_triangle_equals_any_of_triangle(rng, epsilon, hit_tol, use_mask):
    # test_utils.py:649-714
    o = rng.uniform(size=(20, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, size=(20, 3)).astype(np.float32)
    tri = rng.uniform(size=(30, 3, 3)).astype(np.float32)
    active = rng.uniform(size=30) > 0.5 if use_mask else None
    got = orc.ray_intersect_any_triangle(o, d, tri, active, hit_tol=hit_tol, epsilon=epsilon)
    t, hit = orc.ray_intersect_triangle(o[:, None], d[:, None], tri[None], epsilon=epsilon)
    tol = np.float32(100 * orc.EPS if hit_tol is None else hit_tol)
    exp = (t < np.float32(1.0) - tol) & hit
    if active is not None:
        exp &= active[None]
    np.testing.assert_array_equal(got, exp.any(axis=-1))


def test_first_hit_vs_bruteforce(rng):  # test_utils.py:910-962
    o = rng.uniform(size=(50, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, size=(50, 3)).astype(np.float32)
    tri = rng.uniform(size=(77, 3, 3)).astype(np.float32)
    idx, t = orc.first_triangle_hit_by_ray(o, d, tri, batch_size=11)
    tt, hit = orc.ray_intersect_triangle(o[:, None], d[:, None], tri[None])
    tt = np.where(hit, tt, np.inf)
    exp_t = tt.min(axis=-1)
    np.testing.assert_allclose(t, exp_t, rtol=1e-5)
    assert ((idx == -1) == np.isinf(exp_t)).all()
    sel = idx >= 0
    np.testing.assert_array_equal(tt[np.arange(50)[sel], idx[sel]], t[sel])


def test_first_hit_empty_mesh():  # _utils.py:1848-1857
    idx, t = orc.first_triangle_hit_by_ray(
        np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32), np.empty((0, 3, 3), np.float32)
    )
    assert (idx == -1).all() and np.isinf(t).all()


@pytest.mark.parametrize("num_rays", [20, 10_000])
def test_cube_visibility_counts(kats, num_rays):  # test_utils.py:717-767
    v, t = scenes.box(with_top=True)
    tri = orc.triangle_vertices(v, t)
    assert tri.shape == (12, 3, 3)
    for case in kats["cube_visibility"]["cases"]:
        vis = orc.triangles_visible_from_vertex(
            np.array(case["vertex"], np.float32), tri, num_rays=num_rays
        )
        assert int(vis.sum()) == case["expected_number"]


def test_box_in_box_masked_visibility(kats):  # test_utils.py:770-806
    k = kats["box_in_box_visibility"]
    vo, to = scenes.box(*k["outer"])
    vi, ti = scenes.box(*k["inner"])
    v = np.concatenate((vo, vi))
    t = np.concatenate((to, ti + 8))
    tri = orc.triangle_vertices(v, t)
    mask = np.concatenate((np.ones(len(to), bool), np.zeros(len(ti), bool)))
    vis_tx = orc.triangles_visible_from_vertex(np.array(k["tx"], np.float32), tri, mask, 100_000)
    vis_rx = orc.triangles_visible_from_vertex(np.array(k["rx"], np.float32), tri, mask, 100_000)
    np.testing.assert_array_equal(vis_tx, vis_rx)
    assert int(vis_tx.sum()) == k["expected_masked_count"]
    np.testing.assert_array_equal(vis_tx, mask)


def test_image_of_vertex(kats):  # test_image_method.py:19-29
    k = kats["image_of_vertex"]
    got = orc.image_of_vertex_with_respect_to_mirror(
        np.array(k["vertices"]), np.array(k["mirror_vertices"]), np.array(k["mirror_normals"])
    )
    np.testing.assert_allclose(got, np.array(k["expected"], np.float32))


def test_intersection_of_ray_with_plane(kats):  # test_image_method.py:70-91
    k = kats["intersection_of_ray_with_plane"]
    o = np.array(k["ray_origins"], np.float32)
    d = np.array(k["ray_end"], np.float32)[None] - o
    got = orc.intersection_of_ray_with_plane(
        o, d, np.array(k["plane_vertices"]), np.array(k["plane_normals"])
    )
    np.testing.assert_allclose(got, np.array(k["expected"], np.float32), atol=1e-7)


def test_intersection_of_ray_with_plane_parallel(kats):  # test_image_method.py:94-130
    k = kats["intersection_of_ray_with_plane_parallel"]
    o = np.array(k["ray_origins"], np.float32)
    d = np.array(k["ray_end"], np.float32)[None] - o
    n = np.array(k["plane_normals"], np.float32)
    got = orc.intersection_of_ray_with_plane(o, d, np.array(k["plane_vertices_off"]), n)
    assert np.isposinf(got).all()
    got = orc.intersection_of_ray_with_plane(o, d, np.array(k["plane_vertices_on"]), n)
    np.testing.assert_array_equal(got, o)


@pytest.mark.parametrize("batch", [(), (10,), (10, 20, 30)])
def test_corridor_image_method(kats, batch, rng):  # test_image_method.py:160-191
    k = kats["corridor"]
    mv = np.broadcast_to(np.array(k["mirror_vertices"], np.float32), (*batch, 4, 3)).copy()
    mn = np.broadcast_to(np.array(k["mirror_normals"], np.float32), (*batch, 4, 3)).copy()
    # "no-effect noise": in-plane shifts of the mirror vertex and normal sign flips
    shift = rng.normal(size=mv.shape).astype(np.float32) * np.float32(0.1)
    shift = shift - orc.dot3(shift, mn)[..., None] * mn
    sign = rng.choice(np.array([1.0, -1.0], np.float32), size=mv.shape[:-1])
    got = orc.image_method(
        np.array(k["from"], np.float32), np.array(k["to"], np.float32), mv + shift, mn * sign[..., None]
    )
    exp = np.broadcast_to(np.array(k["paths"], np.float32), got.shape)
    np.testing.assert_allclose(got, exp, atol=1e-6)


def test_image_method_points_on_mirrors(rng):  # test_image_method.py:194-222
    fv = rng.uniform(size=(10, 3)).astype(np.float32)
    tv = rng.uniform(size=(10, 3)).astype(np.float32)
    mv = rng.uniform(size=(10, 5, 3)).astype(np.float32)
    mn = orc.normalize(rng.uniform(-1, 1, size=(10, 5, 3)).astype(np.float32))[0]
    paths = orc.image_method(fv, tv, mv, mn)
    assert np.abs(orc.dot3(paths - mv, mn)).max() < 1e-3


def test_image_method_zero_mirrors():  # _solver_image_method.py:349-358
    out = orc.image_method(np.zeros((4, 3)), np.ones((4, 3)), np.zeros((4, 0, 3)), np.zeros((4, 0, 3)))
    assert out.shape == (4, 0, 3)


def test_same_side_shape_error():  # _solver_image_method.py:422-424
    with pytest.raises(TypeError):
        orc.consecutive_vertices_are_on_same_side_of_mirror(
            np.zeros((4, 3)), np.zeros((3, 3)), np.zeros((3, 3))
        )


@pytest.mark.parametrize("assume_quads", [False, True])
@pytest.mark.parametrize("use_mask", [False, True])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
def test_two_buildings_golden_paths(kats, two_buildings, order, assume_quads, use_mask):
    # test_scene.py:116-260 (exhaustive solver)
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    n = t.shape[0] // 2 if assume_quads else t.shape[0]
    cand = scenes.complete_graph_candidates(n, order)
    if assume_quads:
        cand = cand * 2
    mask = np.ones(t.shape[0], bool) if use_mask else None
    full, objects, valid = orc.trace_path_candidates(
        v, t, tx, rx, cand, mask=mask, assume_quads=assume_quads
    )
    exp = k["orders"][str(order)]
    exp_obj = np.array(exp["objects"], np.int32)
    if assume_quads:
        exp_obj = exp_obj - exp_obj % 2
    assert int(valid.sum()) == 1
    got_v = full[valid][0]
    got_o = objects[valid][0]
    exp_v = np.concatenate((tx[None], np.array(exp["vertices"], np.float32).reshape(-1, 3), rx[None]))
    np.testing.assert_allclose(got_v, exp_v, rtol=k["rtol"])
    np.testing.assert_array_equal(got_o, exp_obj)


def test_two_buildings_golden_paths_order4(kats, two_buildings):
    # test_scene.py:148-158 — 292 008 candidates, quads variant to keep the CPU suite short
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    cand = scenes.complete_graph_candidates(t.shape[0] // 2, 4) * 2
    full, objects, valid = orc.trace_path_candidates(v, t, tx, rx, cand, assume_quads=True)
    exp = k["orders"]["4"]
    exp_obj = np.array(exp["objects"], np.int32)
    exp_obj -= exp_obj % 2
    assert int(valid.sum()) == 1
    exp_v = np.concatenate((tx[None], np.array(exp["vertices"], np.float32), rx[None]))
    np.testing.assert_allclose(full[valid][0], exp_v, rtol=k["rtol"])
    np.testing.assert_array_equal(objects[valid][0], exp_obj)


def test_masked_mesh_equals_submesh(two_buildings, kats):  # test_scene.py:585-647
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    rng = np.random.default_rng(7)
    mask = rng.uniform(size=t.shape[0]) > 0.3
    keep = np.nonzero(mask)[0]
    cand_sub = scenes.complete_graph_candidates(len(keep), 2)
    full_s, _, valid_s = orc.trace_path_candidates(v, t[keep], tx, rx, cand_sub)
    cand_full = keep[cand_sub].astype(np.int32)
    full_m, _, valid_m = orc.trace_path_candidates(v, t, tx, rx, cand_full, mask=mask)
    np.testing.assert_array_equal(valid_s, valid_m)
    np.testing.assert_array_equal(full_s, full_m)


def test_first_hit_vjp_matches_finite_differences():
    # structure of test_mesh.py:2029-2073: 2x2x2 box, axis-aligned rays
    v, t = scenes.box(2.0, 2.0, 2.0, with_top=True, with_bottom=True)
    o = np.array([[0.1, 0.2, 3.0], [0.1, 3.0, 0.2], [3.0, 0.1, 0.2]], np.float32)
    d = np.array([[0.0, 0.0, -1.0], [0.0, -1.0, 0.0], [-1.0, 0.0, 0.0]], np.float32)
    tri = orc.triangle_vertices(v, t)
    faces, tt = orc.first_triangle_hit_by_ray(o, d, tri)
    assert (faces >= 0).all()
    np.testing.assert_allclose(tt, 2.0, rtol=1e-6)
    np.testing.assert_allclose(orc.first_hit_distance(v, t, o, d, faces), tt, rtol=1e-6)
    g_t = np.array([1.0, -2.0, 0.5], np.float32)
    gV, gO, gD = orc.first_hit_vjp(v, t, o, d, faces, g_t)

    def f64(vv, oo, dd):
        tri64 = vv[t][faces]
        e1, e2 = tri64[:, 1] - tri64[:, 0], tri64[:, 2] - tri64[:, 0]
        a = np.einsum("ij,ij->i", np.cross(dd, e2), e1)
        q = np.cross(oo - tri64[:, 0], e1)
        return float(np.dot(np.einsum("ij,ij->i", q, e2) / a, g_t.astype(np.float64)))

    def num_grad(fun, x):
        x = x.astype(np.float64)
        g = np.zeros_like(x)
        it = np.nditer(x, flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            xp, xm = x.copy(), x.copy()
            xp[i] += 1e-6
            xm[i] -= 1e-6
            g[i] = (fun(xp) - fun(xm)) / 2e-6
        return g

    v64, o64, d64 = v.astype(np.float64), o.astype(np.float64), d.astype(np.float64)
    np.testing.assert_allclose(gV, num_grad(lambda x: f64(x, o64, d64), v), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gO, num_grad(lambda x: f64(v64, x, d64), o), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gD, num_grad(lambda x: f64(v64, o64, x), d), rtol=1e-4, atol=1e-5)


def test_image_method_vjp_matches_finite_differences(rng):
    N, k = 6, 3
    fv = rng.uniform(size=(N, 3))
    tv = rng.uniform(size=(N, 3))
    mv = rng.uniform(size=(N, k, 3))
    mn = rng.uniform(-1, 1, size=(N, k, 3))
    mn /= np.linalg.norm(mn, axis=-1, keepdims=True)
    g = rng.normal(size=(N, k, 3))
    got = orc.image_method_vjp(fv, tv, mv, mn, g)

    def f64(fv, tv, mv, mn):
        imgs, prev = [], fv
        for i in range(k):
            prev = prev - 2.0 * np.sum((prev - mv[:, i]) * mn[:, i], -1, keepdims=True) * mn[:, i]
            imgs.append(prev)
        total, prev = 0.0, tv
        for i in range(k - 1, -1, -1):
            dirv = imgs[i] - prev
            un = np.sum(dirv * mn[:, i], -1, keepdims=True)
            vn = np.sum((mv[:, i] - prev) * mn[:, i], -1, keepdims=True)
            prev = prev + dirv * (vn / un)
            total += float(np.sum(prev * g[:, i]))
        return total

    args = [fv, tv, mv, mn]
    for ai, ga in enumerate(got):
        num = np.zeros_like(args[ai])
        it = np.nditer(args[ai], flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            ap = [a.copy() for a in args]
            am = [a.copy() for a in args]
            ap[ai][i] += 1e-6
            am[ai][i] -= 1e-6
            num[i] = (f64(*ap) - f64(*am)) / 2e-6
        scale = max(1.0, float(np.abs(num).max()))
        np.testing.assert_allclose(ga, num, rtol=2e-3, atol=2e-3 * scale)


def test_bench_valid_candidates_fixture_is_valid_per_oracle():
    """tests/golden/urban10k_valid_candidates.npz (found on the GPU by tools/find_valid_candidates.py)
    holds candidates the ORACLE also accepts for at least one receiver of the bench workload."""
    import sys
    from pathlib import Path

    GOLDEN = Path(__file__).resolve().parent / "golden"
    sys.path.insert(0, str(GOLDEN.parent.parent))
    import bench
    from oracle import c_oracle as co

    wl = bench.build_workload("urban10k_1tx_4096rx_order3", 0, 1)
    known = np.load(GOLDEN / "urban10k_valid_candidates.npz")
    rx = wl["rx"][:: wl["rx"].shape[0] // 256][:256]
    for order, limit in ((3, 64), (2, 16), (1, 32)):
        cand = known[f"order{order}"][:limit]
        assert cand.shape[0] > 0
        rx_o = wl["rx"] if order == 1 else rx  # order-1 candidates were searched over every receiver
        _, _, mask = co.trace_path_candidates(wl["vertices"], wl["triangles"], wl["tx"], rx_o, cand, early_exit=True)
        assert mask.any(axis=1)[0].all(), f"order-{order} fixture candidate rejected by the oracle"
    # and the bench workload holds them, spread evenly over the candidate list (so that every shard of a
    # multi-GPU run finds valid paths)
    n3 = min(known["order3"].shape[0], wl["cand"].shape[0] // 4)
    slots = np.arange(n3) * (wl["cand"].shape[0] // n3)
    np.testing.assert_array_equal(wl["cand"][slots], known["order3"][:n3])


def test_mlm_hash_functions_known_answers():
    """combine_hashes / hash_int (reference differt/src/differt/geometry/_scene.py:60-78) evaluated by
    hand with Python integers."""
    M1, M2, M3 = 0x9E3779B9, 0x045D9F3B, 0x811C9DC5

    def hash_int(x):
        x = (((x >> 16) ^ x) * M2) & 0xFFFFFFFF
        x = (((x >> 16) ^ x) * M2) & 0xFFFFFFFF
        return (x >> 16) ^ x

    def combine(h1, h2):
        return h1 ^ ((h2 + M1 + ((h1 << 6) & 0xFFFFFFFF) + (h1 >> 2)) & 0xFFFFFFFF)

    for x in (0, 1, 5, 12345, 0xFFFFFFFF):
        assert int(orc.mlm_hash_int(x)) == hash_int(x)
        assert int(orc.mlm_combine_hashes(M3, orc.mlm_hash_int(x))) == combine(M3, hash_int(x))
    assert int(orc.MLM_MAGIC_3) == 2166136261  # the LOS value differt/tests/geometry/test_scene.py:910 expects


def test_sbr_oracle_finds_the_order_one_reflection_of_a_wall():
    """One vertical wall, TX and RX on the same side: the specular point of the image method
    (x = 0 plane) is where SBR rays that pass near the RX bounce."""
    v = np.array([[0, -50, 0], [0, 50, 0], [0, 50, 50], [0, -50, 50]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    tx, rx = np.array([[10.0, -5.0, 10.0]], np.float32), np.array([[10.0, 5.0, 10.0]], np.float32)
    _, dirs = orc.sbr_launch_rays(orc.triangle_vertices(v, t), tx, rx, 40_000)
    cand, verts, masks = orc.sbr_launch_paths(v, t, tx, rx, dirs, 1, max_dist=1e-2)
    assert masks[0, 0, :, 1].any()
    hit = verts[0, masks[0, 0, :, 1], 0, :]
    np.testing.assert_allclose(hit.mean(0), [0.0, 0.0, 10.0], atol=0.2)


def test_smoothing_oracle_matches_hard_decisions_for_large_alpha():
    """Reference property (differt/tests/geometry/test_utils.py:636-646, 701-714 and
    test_image_method.py:225-255): with smoothing_factor = 1e8 the relaxed outputs, thresholded at 0.5,
    equal the hard ones; and sigmoid(0) = 0.5, smoothing_function is monotone."""
    rng = np.random.default_rng(3)
    tri = rng.normal(size=(40, 3, 3)).astype(np.float32)
    o = rng.normal(size=(300, 3)).astype(np.float32)
    d = (rng.normal(size=(300, 3)) * 3).astype(np.float32)
    t, hit = orc.ray_intersect_triangle(o[:, None], d[:, None], tri)
    ts, hs = orc.ray_intersect_triangle_smooth(o[:, None], d[:, None], tri, smoothing_factor=1e8)
    np.testing.assert_array_equal(hs > 0.5, hit)
    np.testing.assert_array_equal(ts.view(np.uint32), t.view(np.uint32))
    mask = rng.uniform(size=40) < 0.5
    for act in (None, mask):
        a = orc.ray_intersect_any_triangle(o, d, tri, act)
        a_s = orc.ray_intersect_any_triangle_smooth(o, d, tri, act, smoothing_factor=1e8)
        np.testing.assert_array_equal(a_s > 0.5, a)
        assert a_s.min() >= 0.0 and a_s.max() <= 1.0
    v = rng.normal(size=(50, 5, 3)).astype(np.float32)
    mv, mn = rng.normal(size=(50, 3, 3)).astype(np.float32), rng.normal(size=(50, 3, 3)).astype(np.float32)
    hard = orc.consecutive_vertices_are_on_same_side_of_mirror(v, mv, mn)
    soft = orc.consecutive_vertices_are_on_same_side_of_mirror_smooth(v, mv, mn, 1e8)
    np.testing.assert_array_equal(soft > 0.5, hard)
    assert float(orc.smoothing_function(0.0, 7.0)) == 0.5
    x = np.linspace(-3, 3, 50, dtype=np.float32)
    assert (np.diff(orc.smoothing_function(x, 2.0)) > 0).all()
