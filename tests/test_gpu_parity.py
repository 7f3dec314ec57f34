"""GPU parity tests: the CUDA path (through the public API → C ABI) against the CPU oracle.

Bar (BASELINE.json north_star): boolean / index outputs bit-exact; fp32 hit distances and path
vertices within 1e-5 relative — in practice they are bit-exact too, because the kernels follow the
oracle's operation order without FMA, and the tests assert that stronger property where it holds.
Gradients (float atomics, different summation order) are compared at rtol 1e-4.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from differt_b200 import scenes
from oracle import c_oracle as co
from oracle import differt_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north-star tolerance for fp32 hit points / path vertices


@pytest.fixture(scope="module")
def drt():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import differt_b200

    return differt_b200


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def scene_rays(rng, v, n):
    lo, hi = v.min(0), v.max(0)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    e = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    o[:, 2] = rng.uniform(0.5, 45.0, n)
    e[:, 2] = rng.uniform(0.5, 45.0, n)
    return o, (e - o).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------


def test_k1_reference_kats(drt, kats):
    k = kats["ray_intersect_triangle_hit_table"]  # test_utils.py:555-577
    tri = np.array([k["triangle"]], np.float32)
    for case in k["cases"]:
        o = np.array(case["orig"], np.float32)
        d = np.array(case["dest"], np.float32) - o
        t, hit = drt.ray_intersect_triangle(o, d, tri)
        assert bool(((t < 1.0) & hit)[0]) == case["expected"]
    k = kats["ray_intersect_triangle_t_and_hit"]  # test_utils.py:580-606
    o = np.array(k["ray_origin"], np.float32)
    d = np.array(k["ray_directions"], np.float32)
    tri = np.array(k["triangles"], np.float32)
    t, hit = drt.ray_intersect_triangle(o[None, None, :], d[:, None, :], tri)
    np.testing.assert_array_equal(t.numpy(), np.array(k["expected_t"], np.float32))
    np.testing.assert_array_equal(hit.numpy(), np.array(k["expected_hit"]))


@pytest.mark.parametrize(
    "shapes", [((3,), (3,), (3, 3)), ((15, 5, 3), (15, 5, 3), (5, 3, 3)), ((7, 1, 3), (1, 9, 3), (7, 9, 3, 3)),
               ((2, 3, 4, 5, 6, 3), (6, 3), (5, 1, 3, 3))]
)
@pytest.mark.parametrize("epsilon", [None, 1e-3])
def test_k1_bit_exact_vs_oracle(drt, rng, shapes, epsilon):
    o = rng.uniform(size=shapes[0]).astype(np.float32)
    d = rng.uniform(-1, 1, size=shapes[1]).astype(np.float32)
    tri = rng.uniform(size=shapes[2]).astype(np.float32)
    t0, h0 = orc.ray_intersect_triangle(o, d, tri, epsilon=epsilon)
    t1, h1 = drt.ray_intersect_triangle(o, d, tri, epsilon=epsilon)
    assert tuple(t1.shape) == t0.shape
    np.testing.assert_array_equal(bits(t1.numpy()), bits(t0))
    np.testing.assert_array_equal(h1.numpy(), h0)
    assert np.where(h0, t0 > 0, True).all()


def test_k1_degenerate_inputs(drt):
    # parallel ray (a == 0 → t = 0), zero-area triangle, NaN input: no hit, finite-or-NaN t like oracle
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 0], [0, 0, 0], [0, 0, 0]],
                    [[np.nan, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    o = np.array([[0.2, 0.2, 1.0]] * 3, np.float32)
    d = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 0.0, -1.0]], np.float32)
    t0, h0 = orc.ray_intersect_triangle(o, d, tri)
    t1, h1 = drt.ray_intersect_triangle(o, d, tri)
    np.testing.assert_array_equal(h1.numpy(), h0)
    assert not h0.any()
    np.testing.assert_array_equal(np.isnan(t1.numpy()), np.isnan(t0))
    np.testing.assert_array_equal(t1.numpy()[~np.isnan(t0)], t0[~np.isnan(t0)])


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("epsilon", [None, 1e-6, 1e-2])
@pytest.mark.parametrize("hit_tol", [None, 0.0, 1e-3, 0.5, -0.5])
@pytest.mark.parametrize("use_mask", [False, True])
def test_k2_any_equals_any_of_elementwise(drt, rng, epsilon, hit_tol, use_mask):
    # test_utils.py:649-714
    o = rng.uniform(size=(21, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, size=(21, 3)).astype(np.float32)
    tri = rng.uniform(size=(30, 3, 3)).astype(np.float32)
    active = rng.uniform(size=30) > 0.5 if use_mask else None
    kw = {} if epsilon is None else {"epsilon": epsilon}
    got = drt.ray_intersect_any_triangle(o, d, tri, active, hit_tol=hit_tol, batch_size=11, **kw)
    exp = orc.ray_intersect_any_triangle(o, d, tri, active, hit_tol=hit_tol, epsilon=epsilon)
    np.testing.assert_array_equal(got.numpy(), exp)


@pytest.mark.parametrize("grid,n_rays", [((1, 1), 37), ((3, 3), 4001), ((7, 7), 3000), ((29, 29), 2048)])
@pytest.mark.parametrize("use_mask", [False, True])
def test_k2_scene_bit_exact(drt, rng, grid, n_rays, use_mask):
    # T = 14 / 110 (one tile), 590 (two tiles, resident ring), 10 094 (20 tiles, streamed ring)
    v, t = scenes.urban_grid(*grid)
    tri = orc.triangle_vertices(v, t)
    o, d = scene_rays(rng, v, n_rays)
    active = rng.uniform(size=t.shape[0]) > 0.5 if use_mask else None
    exp = co.ray_intersect_any_triangle(o, d, tri, active)
    got = drt.ray_intersect_any_triangle(o, d, tri, active)
    np.testing.assert_array_equal(got.numpy(), exp)
    assert 0.02 < exp.mean() < 0.98
    mesh = drt.Mesh.from_numpy(v, t, active)
    got2 = mesh.ray_intersect_any_triangle(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    np.testing.assert_array_equal(got2.cpu().numpy(), exp)


def test_k2_empty_and_ragged(drt, rng):
    o = rng.uniform(size=(5, 3)).astype(np.float32)
    d = rng.uniform(size=(5, 3)).astype(np.float32)
    out = drt.ray_intersect_any_triangle(o, d, np.empty((0, 3, 3), np.float32))  # _utils.py:1441-1450
    assert out.shape == (5,) and not out.any()
    out = drt.ray_intersect_any_triangle(np.empty((0, 3), np.float32), np.empty((0, 3), np.float32),
                                         rng.uniform(size=(4, 3, 3)).astype(np.float32))
    assert out.shape == (0,)
    # broadcast: one origin, many directions; batched meshes
    tri = rng.uniform(size=(3, 6, 3, 3)).astype(np.float32)
    dd = rng.uniform(-1, 1, size=(4, 1, 3)).astype(np.float32)
    got = drt.ray_intersect_any_triangle(o[0], dd, tri)
    exp = np.stack([orc.ray_intersect_any_triangle(o[0], dd[:, 0], tri[j]) for j in range(3)], axis=-1)
    np.testing.assert_array_equal(got.numpy(), exp)


def test_k2_counts_tests(drt, rng):
    from differt_b200 import _lib
    from differt_b200._tensor import ptr, stream_ptr
    from differt_b200.geometry import pack_triangle_vertices

    v, t = scenes.urban_grid(5, 5)
    tri = torch.from_numpy(orc.triangle_vertices(v, t)).cuda()
    o, d = scene_rays(rng, v, 1000)
    # rays that cannot hit anything (pointing up from above the roofs): no early exit → R * T_pad tests
    o[:, 2] = 100.0
    d[:] = (0.0, 0.0, 1.0)
    oc, dc = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    pack = pack_triangle_vertices(tri)
    out = torch.empty(1000, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib.drt_ray_intersect_any_triangle(
        stream_ptr(), 1000, ptr(oc), ptr(dc), ptr(pack), t.shape[0], 1e-6, 1e-5, ptr(out), ptr(cnt)))
    assert not out.any()
    assert int(cnt.item()) == 1000 * 512


def test_k2_culled_any_hit_equals_all_pairs_and_oracle(drt, rng, bruxelles, monkeypatch):
    """The flat any-hit behind the exact cull: same bits as the all-pairs engine and the oracle, on the
    urban grid, on the reference's bruxelles.obj (a general, non axis-aligned mesh) and on the coplanar
    clusters that produce noise hits far from the triangles — with a fraction of the tests."""
    from differt_b200 import _lib, geometry

    monkeypatch.setattr(geometry, "_CULL_MIN_WORK", 0)  # the public API takes the culled path at any size
    from differt_b200._tensor import ptr, stream_ptr
    from differt_b200.geometry import pack_triangle_vertices, sort_pack_by_area

    cases = []
    v, t = scenes.urban_grid(29, 29)
    cases.append(("urban", orc.triangle_vertices(v, t), *scene_rays(rng, v, 60_000), None))
    bv, bt = bruxelles
    cases.append(("bruxelles", orc.triangle_vertices(bv, bt), *scene_rays(rng, bv, 40_000), rng.uniform(size=bt.shape[0]) < 0.5))
    cv, ct, planes = _coplanar_clusters(np.random.default_rng(9))
    c0, a, b, n, _ = planes[0]
    uv0, uv1 = rng.uniform(-2500, 2500, (30_000, 2)), rng.uniform(-2500, 2500, (30_000, 2))
    po = (c0 + uv0[:, 0:1] * a + uv0[:, 1:2] * b).astype(np.float32)
    pe = (c0 + uv1[:, 0:1] * a + uv1[:, 1:2] * b).astype(np.float32)
    cases.append(("coplanar", orc.triangle_vertices(cv, ct), po, (pe - po).astype(np.float32), None))
    for name, tri, o, d, active in cases:
        T = tri.shape[0]
        exp = co.ray_intersect_any_triangle(o, d, tri, active)
        got = drt.ray_intersect_any_triangle(o, d, tri, active)  # public API: culled for T > 2048
        np.testing.assert_array_equal(got.numpy(), exp, err_msg=name)
        assert exp.any() and not exp.all()
        # through the C ABI, both engines, with their test counters
        tc, oc, dc = (torch.from_numpy(x).cuda() for x in (tri, o, d))
        act = None if active is None else torch.from_numpy(active.astype(np.uint8)).cuda()
        pack = sort_pack_by_area(pack_triangle_vertices(tc, act), T)
        res = [torch.empty(o.shape[0], dtype=torch.uint8, device="cuda") for _ in range(2)]
        cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
        ws = torch.empty(_lib.lib.drt_any_hit_workspace_bytes(T), dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib.drt_ray_intersect_any_triangle(
            stream_ptr(), o.shape[0], ptr(oc), ptr(dc), ptr(pack), T, 10 * orc.EPS, 100 * orc.EPS, ptr(res[0]), ptr(cnt[0:1])))
        _lib.check(_lib.lib.drt_ray_intersect_any_triangle_culled(
            stream_ptr(), o.shape[0], ptr(oc), ptr(dc), ptr(pack), T, 10 * orc.EPS, 100 * orc.EPS, ptr(ws), ws.numel(),
            ptr(res[1]), ptr(cnt[1:2])))
        assert torch.equal(res[0], res[1]), name
        np.testing.assert_array_equal(res[1].cpu().numpy().astype(bool), exp, err_msg=name)
        brute, culled = cnt.cpu().tolist()
        print(f"{name}: all-pairs engine {brute} tests, culled {culled} tests")
        if name != "coplanar":  # (coplanar rays are exactly what the cull cannot prove anything about)
            assert culled < 0.5 * brute, (name, brute, culled)
    # parameters outside the proof fall back to the all-pairs engine (same results as the oracle)
    name, tri, o, d, active = cases[0]
    for kw in ({"hit_tol": -0.5}, {"epsilon": 0.0}, {"epsilon": -1.0}):
        exp = co.ray_intersect_any_triangle(o[:5000], d[:5000], tri, **kw)
        np.testing.assert_array_equal(drt.ray_intersect_any_triangle(o[:5000], d[:5000], tri, **kw).numpy(), exp)


# ------------------------------------------------------------------------------------------------
# K3 (+ K3b)
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("grid,n_rays", [((2, 2), 999), ((7, 7), 2000), ((29, 29), 1024)])
@pytest.mark.parametrize("batch_size", [None, 7, 512])
def test_k3_first_hit_bit_exact(drt, rng, grid, n_rays, batch_size):
    v, t = scenes.urban_grid(*grid)
    tri = orc.triangle_vertices(v, t)
    o, d = scene_rays(rng, v, n_rays)
    active = rng.uniform(size=t.shape[0]) > 0.2
    i0, t0 = co.first_triangle_hit_by_ray(o, d, tri, active, batch_size=batch_size)
    i1, t1 = drt.first_triangle_hit_by_ray(o, d, tri, active, batch_size=batch_size)
    np.testing.assert_array_equal(bits(t1.numpy()), bits(t0))
    np.testing.assert_array_equal(i1.numpy(), i0)  # includes the reference's tie rule on shared edges
    assert (i0 == -1).any() and (i0 >= 0).any()
    assert np.isinf(t0[i0 == -1]).all()


@pytest.mark.parametrize("batch_size", [512, 7, None])
def test_k3_culled_first_hit_ties_general_mesh_and_masks(drt, rng, bruxelles, batch_size, monkeypatch):
    """The nearest-hit query behind the exact cull (meshes > 2048 triangles): exact ties on an
    integer-lattice city (the tie rule picks the index), the reference's bruxelles.obj with half the
    triangles masked, and the visibility scatter built on it — index AND distance bits vs the oracle."""
    from differt_b200 import geometry

    monkeypatch.setattr(geometry, "_CULL_MIN_WORK", 0)  # the public API takes the culled path at any size
    parts = [scenes.box(4.0, 4.0, 6.0, with_top=True, center=(8.0 * i, 8.0 * j, 3.0)) for i in range(15) for j in range(15)]
    v, t = scenes._merge(parts)  # 2700 triangles on integer coordinates
    tri = orc.triangle_vertices(v, t)
    g = np.arange(-4, 122, 2, dtype=np.float32)
    pts = np.stack(np.meshgrid(g, g, np.array([0.0, 3.0, 6.0, 7.0], np.float32), indexing="ij"), -1).reshape(-1, 3)
    a, b = pts[rng.integers(0, pts.shape[0], 20000)], pts[rng.integers(0, pts.shape[0], 20000)]
    o, d = a, (b - a).astype(np.float32)
    ei, et = co.first_triangle_hit_by_ray(o, d, tri, batch_size=batch_size)
    gi, gt = drt.first_triangle_hit_by_ray(o, d, tri, batch_size=batch_size)
    np.testing.assert_array_equal(gi.numpy(), ei)
    np.testing.assert_array_equal(bits(gt.numpy()), bits(et))
    assert (ei >= 0).sum() > 5000
    bv, bt = bruxelles
    btri = orc.triangle_vertices(bv, bt)
    active = rng.uniform(size=bt.shape[0]) < 0.5
    o, d = scene_rays(rng, bv, 30_000)
    ei, et = co.first_triangle_hit_by_ray(o, d, btri, active, batch_size=batch_size)
    gi, gt = drt.first_triangle_hit_by_ray(o, d, btri, active, batch_size=batch_size)
    np.testing.assert_array_equal(gi.numpy(), ei)
    np.testing.assert_array_equal(bits(gt.numpy()), bits(et))
    mesh = drt.Mesh.from_numpy(bv, bt, mask=active)
    mi, mt_ = mesh.first_triangle_hit_by_ray(o, d, batch_size=batch_size)
    np.testing.assert_array_equal(mi.cpu().numpy(), ei)
    np.testing.assert_array_equal(bits(mt_.cpu().numpy()), bits(et))
    if batch_size == 512:  # visibility = nearest hits + scatter, with shared directions
        vertex = np.array([[60.0, 60.0, 30.0], [10.0, 100.0, 2.0]], np.float32)
        dirs = rng.normal(size=(2, 3000, 3)).astype(np.float32)
        exp = co.triangles_visible_from_vertex_dirs(vertex, dirs, tri)
        got = drt.triangles_visible_from_vertex(vertex, tri, ray_directions=dirs)
        np.testing.assert_array_equal(got.numpy(), exp)
        assert exp.any()


def test_k3_tie_rule_and_empty(drt):
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]] * 3, np.float32)
    o = np.array([[0.2, 0.2, 1.0]], np.float32)
    d = np.array([[0.0, 0.0, -1.0]], np.float32)
    assert int(drt.first_triangle_hit_by_ray(o, d, tri, batch_size=1)[0][0]) == 2   # latest batch wins
    assert int(drt.first_triangle_hit_by_ray(o, d, tri, batch_size=2)[0][0]) == 2   # remainder batch wins
    assert int(drt.first_triangle_hit_by_ray(o, d, tri, batch_size=None)[0][0]) == 0  # first in batch
    assert int(drt.first_triangle_hit_by_ray(o, d, tri)[0][0]) == 0
    i, t = drt.first_triangle_hit_by_ray(o, d, np.empty((0, 3, 3), np.float32))  # _utils.py:1848-1857
    assert int(i[0]) == -1 and np.isinf(float(t[0]))


def test_k3_mesh_method_matches_function(drt, rng):
    # test_mesh.py:2004-2027: Mesh method == function on triangle_vertices
    v, t = scenes.urban_grid(3, 3)
    o, d = scene_rays(rng, v, 500)
    mesh = drt.Mesh.from_numpy(v, t)
    i0, t0 = drt.first_triangle_hit_by_ray(o, d, orc.triangle_vertices(v, t))
    i1, t1 = mesh.first_triangle_hit_by_ray(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda())
    np.testing.assert_array_equal(i1.cpu().numpy(), i0.numpy())
    np.testing.assert_array_equal(bits(t1.cpu().numpy()), bits(t0.numpy()))


def test_k3b_vjp_vs_oracle(drt, rng):
    v, t = scenes.urban_grid(3, 3)
    o, d = scene_rays(rng, v, 800)
    mesh = drt.Mesh.from_numpy(v, t)
    mesh.vertices.requires_grad_(True)
    oc = torch.from_numpy(o).cuda().requires_grad_(True)
    dc = torch.from_numpy(d).cuda().requires_grad_(True)
    idx, tt = mesh.first_triangle_hit_by_ray(oc, dc)
    assert not idx.requires_grad
    g = rng.normal(size=800).astype(np.float32)
    hit = (idx >= 0)
    (torch.where(hit, tt, torch.zeros_like(tt)) * torch.from_numpy(g).cuda()).sum().backward()
    faces = idx.cpu().numpy()
    gV, gO, gD = orc.first_hit_vjp(v, t, o, d, faces, np.where(faces >= 0, g, 0).astype(np.float32))
    np.testing.assert_allclose(oc.grad.cpu().numpy(), gO, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(dc.grad.cpu().numpy(), gD, rtol=1e-4, atol=1e-6)
    scale = np.abs(gV).max()
    np.testing.assert_allclose(mesh.vertices.grad.cpu().numpy(), gV, rtol=1e-3, atol=1e-5 * scale)


def test_k3b_jacobian_structure_box(drt):
    # test_mesh.py:2029-2073: 2x2x2 box, axis-aligned rays; Jacobians of t vs the brute-force path
    v, t = scenes.box(2.0, 2.0, 2.0, with_top=True)
    o = np.array([[0.1, 0.2, 3.0], [0.1, 3.0, 0.2], [3.0, 0.1, 0.2]], np.float32)
    d = np.array([[0.0, 0.0, -1.0], [0.0, -1.0, 0.0], [-1.0, 0.0, 0.0]], np.float32)
    mesh = drt.Mesh.from_numpy(v, t)
    for r in range(3):
        mesh.vertices.grad = None
        mesh.vertices.requires_grad_(True)
        oc = torch.from_numpy(o).cuda().requires_grad_(True)
        dc = torch.from_numpy(d).cuda().requires_grad_(True)
        idx, tt = mesh.first_triangle_hit_by_ray(oc, dc)
        np.testing.assert_allclose(tt.detach().cpu().numpy(), 2.0, rtol=1e-6)
        tt[r].backward()
        e = np.zeros(3, np.float32)
        e[r] = 1.0
        gV, gO, gD = orc.first_hit_vjp(v, t, o, d, idx.cpu().numpy(), e)
        np.testing.assert_allclose(oc.grad.cpu().numpy(), gO, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(dc.grad.cpu().numpy(), gD, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(mesh.vertices.grad.cpu().numpy(), gV, rtol=1e-5, atol=1e-5)
        assert np.count_nonzero(gO[np.arange(3) != r]) == 0


def test_k1_t_gradient(drt, rng):
    o = rng.uniform(size=(40, 3)).astype(np.float32)
    d = rng.uniform(-1, 1, size=(40, 3)).astype(np.float32)
    tri = rng.uniform(size=(40, 3, 3)).astype(np.float32)
    oc = torch.from_numpy(o).cuda().requires_grad_(True)
    tc = torch.from_numpy(tri).cuda().requires_grad_(True)
    t, _ = drt.ray_intersect_triangle(oc, torch.from_numpy(d).cuda(), tc)
    g = rng.normal(size=40).astype(np.float32)
    (t * torch.from_numpy(g).cuda()).sum().backward()
    gV, gO, _ = orc.first_hit_vjp(tri.reshape(-1, 3), np.arange(120).reshape(40, 3), o, d, np.arange(40), g)
    np.testing.assert_allclose(oc.grad.cpu().numpy(), gO, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(tc.grad.cpu().numpy().reshape(-1, 3), gV, rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("num_rays", [20, 10_000])
def test_k4_cube_counts(drt, kats, num_rays):  # test_utils.py:717-767
    v, t = scenes.box(with_top=True)
    tri = orc.triangle_vertices(v, t)
    for case in kats["cube_visibility"]["cases"]:
        vis = drt.triangles_visible_from_vertex(np.array(case["vertex"], np.float32), tri, num_rays=num_rays)
        assert int(vis.sum()) == case["expected_number"]


def test_k4_box_in_box_masked(drt, kats):  # test_utils.py:770-806
    k = kats["box_in_box_visibility"]
    vo, to = scenes.box(*k["outer"])
    vi, ti = scenes.box(*k["inner"])
    v, t = np.concatenate((vo, vi)), np.concatenate((to, ti + 8))
    mask = np.concatenate((np.ones(len(to), bool), np.zeros(len(ti), bool)))
    mesh = drt.Mesh.from_numpy(v, t, mask)
    both = mesh.triangles_visible_from_vertex(np.array([k["tx"], k["rx"]], np.float32), num_rays=100_000)
    both = both.cpu().numpy()
    np.testing.assert_array_equal(both[0], both[1])
    assert int(both[0].sum()) == k["expected_masked_count"]
    np.testing.assert_array_equal(both[0], mask)


def test_k4_bit_exact_with_shared_directions(drt, rng):
    v, t = scenes.urban_grid(4, 4)
    tri = orc.triangle_vertices(v, t)
    active = rng.uniform(size=t.shape[0]) > 0.3
    vertices = np.array([[45.0, 45.0, 60.0], [15.0, 45.0, 1.5], [-10.0, -10.0, 5.0]], np.float32)
    dirs = np.stack([orc.visibility_directions(p, tri, active, 3001) for p in vertices])
    got = drt.triangles_visible_from_vertex(vertices, tri, active, ray_directions=dirs)
    exp = co.triangles_visible_from_vertex_dirs(vertices, dirs, tri, active)
    np.testing.assert_array_equal(got.numpy(), exp)
    assert exp.any() and not exp[:, ~active].any()


def test_k4_batched_meshes_and_masks_follow_the_reference_broadcasting(drt, rng):
    """`triangles_visible_from_vertex(vertex [*#b,3], triangles [*#b,T,3,3], active [*#b,T])`
    (reference `_utils.py:1540-1548`): a batch of meshes / masks broadcast against a batch of vertices,
    each batch element equal to the single-mesh call and to the oracle (shared ray directions)."""
    v, t = scenes.street_canyon(3)
    base = orc.triangle_vertices(v, t)
    T = base.shape[0]
    meshes = np.stack([base, base + np.float32([5.0, 0.0, 0.0]), base * np.float32(1.5)])      # [3,T,3,3]
    masks = np.stack([np.ones(T, bool), rng.uniform(size=T) < 0.6])                              # [2,T]
    vertices = rng.uniform([0, -8, 1], [25, 8, 30], size=(2, 1, 3)).astype(np.float32)          # [2,1,3]
    dirs = rng.normal(size=(400, 3)).astype(np.float32)
    # batch = broadcast([2,1], [3], [2,1]→masks as [2,1,T]) = [2,3]
    got = drt.triangles_visible_from_vertex(vertices, meshes[None], masks[:, None, :], ray_directions=dirs)
    assert tuple(got.shape) == (2, 3, T)
    for i in range(2):
        for j in range(3):
            exp = co.triangles_visible_from_vertex_dirs(vertices[i, 0], dirs[None], meshes[j], masks[i])
            np.testing.assert_array_equal(got[i, j].numpy(), exp[0])
            one = drt.triangles_visible_from_vertex(vertices[i, 0], meshes[j], masks[i], ray_directions=dirs)
            np.testing.assert_array_equal(one.numpy(), exp[0])
    assert got.any() and not got.all()
    # generated directions (frustum + Fibonacci lattice per mesh): batched call == per-mesh calls
    gen = drt.triangles_visible_from_vertex(vertices, meshes[None], num_rays=2000)
    for j in range(3):
        np.testing.assert_array_equal(gen[:, j].numpy(),
                                      drt.triangles_visible_from_vertex(vertices[:, 0], meshes[j], num_rays=2000).numpy())


def test_ray_generation_matches_oracle(drt, rng):
    v, t = scenes.urban_grid(2, 2)
    tri = orc.triangle_vertices(v, t)
    world = np.concatenate((tri, tri.mean(axis=-2, keepdims=True)), axis=-2).reshape(-1, 3)
    for p in ([15.0, 15.0, 60.0], [-30.0, 10.0, 1.5], [15.0, 15.0, 5.0]):
        p = np.array(p, np.float32)
        fr0 = orc.viewing_frustum(p, world)
        fr1 = drt.viewing_frustum(p, world).numpy()
        np.testing.assert_allclose(fr1, fr0, rtol=1e-6, atol=1e-6)
        d0 = orc.fibonacci_lattice(500, fr0)
        d1 = drt.fibonacci_lattice(500, frustum=fr0).numpy()
        np.testing.assert_allclose(d1, d0, atol=2e-6)
    np.testing.assert_allclose(drt.fibonacci_lattice(100).cpu().numpy(), orc.fibonacci_lattice(100), atol=2e-6)


# ------------------------------------------------------------------------------------------------
# K5 (+ K5b) and the small element-wise kernels
# ------------------------------------------------------------------------------------------------


def test_image_kats(drt, kats):
    k = kats["image_of_vertex"]  # test_image_method.py:19-29
    got = drt.image_of_vertex_with_respect_to_mirror(
        np.array(k["vertices"], np.float32), np.array(k["mirror_vertices"], np.float32),
        np.array(k["mirror_normals"], np.float32))
    np.testing.assert_array_equal(got.numpy(), np.array(k["expected"], np.float32))
    k = kats["intersection_of_ray_with_plane"]  # test_image_method.py:70-91
    o = np.array(k["ray_origins"], np.float32)
    d = np.array(k["ray_end"], np.float32)[None] - o
    got = drt.intersection_of_ray_with_plane(o, d, np.array(k["plane_vertices"], np.float32),
                                             np.array(k["plane_normals"], np.float32))
    np.testing.assert_allclose(got.numpy(), np.array(k["expected"], np.float32), atol=1e-7)
    k = kats["intersection_of_ray_with_plane_parallel"]  # test_image_method.py:94-130
    n = np.array(k["plane_normals"], np.float32)
    got = drt.intersection_of_ray_with_plane(o, d, np.array(k["plane_vertices_off"], np.float32), n)
    assert torch.isposinf(got).all()
    got = drt.intersection_of_ray_with_plane(o, d, np.array(k["plane_vertices_on"], np.float32), n)
    np.testing.assert_array_equal(got.numpy(), o)


@pytest.mark.parametrize("batch", [(), (10,), (10, 20, 30)])
def test_k5_corridor(drt, kats, batch, rng):  # test_image_method.py:160-191
    k = kats["corridor"]
    mv = np.broadcast_to(np.array(k["mirror_vertices"], np.float32), (*batch, 4, 3)).copy()
    mn = np.broadcast_to(np.array(k["mirror_normals"], np.float32), (*batch, 4, 3)).copy()
    shift = rng.normal(size=mv.shape).astype(np.float32) * np.float32(0.1)
    shift = shift - orc.dot3(shift, mn)[..., None] * mn
    sign = rng.choice(np.array([1.0, -1.0], np.float32), size=mv.shape[:-1])
    args = (np.array(k["from"], np.float32), np.array(k["to"], np.float32), mv + shift, mn * sign[..., None])
    got = drt.image_method(*args).numpy()
    np.testing.assert_allclose(got, np.broadcast_to(np.array(k["paths"], np.float32), got.shape), atol=1e-6)
    np.testing.assert_array_equal(bits(got), bits(orc.image_method(*args)))


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 12])
def test_k5_bit_exact_with_inf_cases(drt, rng, k):
    N = 300
    fv = rng.uniform(size=(N, 3)).astype(np.float32)
    tv = rng.uniform(size=(N, 3)).astype(np.float32)
    mv = rng.uniform(size=(N, k, 3)).astype(np.float32)
    mn = orc.normalize(rng.uniform(-1, 1, size=(N, k, 3)).astype(np.float32))[0]
    mv[::7] = 0.0
    mn[::7] = np.array([0.0, 0.0, 1.0], np.float32)
    fv[::7, 2] = 1.0
    tv[::7, 2] = 1.0 if k % 2 == 0 else -1.0  # k-th image sits at z = ±1 → last ray parallel
    exp = orc.image_method(fv, tv, mv, mn)
    got = drt.image_method(fv, tv, mv, mn).numpy()
    np.testing.assert_array_equal(bits(got), bits(exp))
    assert np.isinf(exp).any()
    # broadcast form used by the solver: [Ntx,1,1] x [1,Nrx,1] x [C,k]
    got = drt.image_method(fv[:3, None, None], tv[None, :4, None], mv[:50], mn[:50]).numpy()
    exp = orc.image_method(fv[:3, None, None], tv[None, :4, None], mv[:50], mn[:50])
    assert got.shape == (3, 4, 50, k, 3)
    np.testing.assert_array_equal(bits(got), bits(exp))


def test_k5_zero_mirrors_and_same_side(drt, rng):
    out = drt.image_method(np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32),
                           np.zeros((4, 0, 3), np.float32), np.zeros((4, 0, 3), np.float32))
    assert tuple(out.shape) == (4, 0, 3)  # _solver_image_method.py:349-358
    v = rng.uniform(-1, 1, size=(6, 7, 5, 3)).astype(np.float32)
    mv = rng.uniform(-1, 1, size=(7, 3, 3)).astype(np.float32)
    mn = rng.uniform(-1, 1, size=(6, 1, 3, 3)).astype(np.float32)
    got = drt.consecutive_vertices_are_on_same_side_of_mirror(v, mv, mn)
    exp = orc.consecutive_vertices_are_on_same_side_of_mirror(v, mv, mn)
    np.testing.assert_array_equal(got.numpy(), exp)
    with pytest.raises(TypeError):  # _solver_image_method.py:422-424
        drt.consecutive_vertices_are_on_same_side_of_mirror(v[..., :4, :], mv, mn)


@pytest.mark.parametrize("k", [1, 3, 8])
def test_k5b_vjp_vs_oracle(drt, rng, k):
    N = 64
    fv = rng.uniform(size=(N, 3)).astype(np.float32)
    tv = rng.uniform(size=(N, 3)).astype(np.float32)
    mv = rng.uniform(size=(N, k, 3)).astype(np.float32)
    mn = orc.normalize(rng.uniform(-1, 1, size=(N, k, 3)).astype(np.float32))[0]
    g = rng.normal(size=(N, k, 3)).astype(np.float32)
    ts = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (fv, tv, mv, mn)]
    (drt.image_method(*ts) * torch.from_numpy(g).cuda()).sum().backward()
    exp = orc.image_method_vjp(fv, tv, mv, mn, g)
    for tns, e in zip(ts, exp):
        got = tns.grad.cpu().numpy()
        scale = max(np.abs(e).max(), 1.0)
        np.testing.assert_allclose(got, e, rtol=1e-4, atol=1e-5 * scale)


def test_k5b_vjp_broadcast_reduction(drt, rng):
    k = 2
    fv = rng.uniform(size=(2, 1, 1, 3)).astype(np.float32)
    tv = rng.uniform(size=(1, 3, 1, 3)).astype(np.float32)
    mv = rng.uniform(size=(5, k, 3)).astype(np.float32)
    mn = orc.normalize(rng.uniform(-1, 1, size=(5, k, 3)).astype(np.float32))[0]
    g = rng.normal(size=(2, 3, 5, k, 3)).astype(np.float32)
    ts = [torch.from_numpy(a).cuda().requires_grad_(True) for a in (fv, tv, mv, mn)]
    (drt.image_method(*ts) * torch.from_numpy(g).cuda()).sum().backward()
    N = 30
    bf = np.broadcast_to(fv, (2, 3, 5, 3)).reshape(N, 3)
    bt = np.broadcast_to(tv, (2, 3, 5, 3)).reshape(N, 3)
    bv = np.broadcast_to(mv, (2, 3, 5, k, 3)).reshape(N, k, 3)
    bn = np.broadcast_to(mn, (2, 3, 5, k, 3)).reshape(N, k, 3)
    gf, gt, gv, gn = orc.image_method_vjp(bf, bt, bv, bn, g.reshape(N, k, 3))
    np.testing.assert_allclose(ts[0].grad.cpu().numpy(), gf.reshape(2, 3, 5, 3).sum((1, 2))[:, None, None], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ts[1].grad.cpu().numpy(), gt.reshape(2, 3, 5, 3).sum((0, 2))[None, :, None], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ts[2].grad.cpu().numpy(), gv.reshape(6, 5, k, 3).sum(0), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(ts[3].grad.cpu().numpy(), gn.reshape(6, 5, k, 3).sum(0), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------------------------------------
# K6 (+ K6b), compaction, candidates
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("assume_quads", [False, True])
@pytest.mark.parametrize("use_mask", [False, True])
@pytest.mark.parametrize("order", [0, 1, 2, 3, 4])
def test_k6_two_buildings_golden(drt, kats, two_buildings, order, assume_quads, use_mask):
    # test_scene.py:116-260 (exhaustive solver; order 4 = 292 008 candidates without quads)
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    mask = np.ones(t.shape[0], bool) if use_mask else None
    mesh = drt.Mesh.from_numpy(v, t, mask, assume_quads)
    paths = drt.trace_paths(mesh, tx, rx, order)
    exp = k["orders"][str(order)]
    exp_obj = np.array(exp["objects"], np.int32)
    if assume_quads:
        exp_obj = exp_obj - exp_obj % 2
    exp_v = np.concatenate((tx[None], np.array(exp["vertices"], np.float32).reshape(-1, 3), rx[None]))
    assert paths.mask.shape[:2] == (1, 1)
    assert paths.num_valid_paths == 1
    m = paths.masked()
    np.testing.assert_allclose(m.vertices.cpu().numpy()[0], exp_v, rtol=k["rtol"])
    np.testing.assert_array_equal(m.objects.cpu().numpy()[0], exp_obj)
    # reflection law (test_scene.py:248-260)
    if order > 0:
        nrm = mesh.normals.cpu().numpy()[m.objects.cpu().numpy()[0, 1:-1]]
        rays = orc.normalize(np.diff(m.vertices.cpu().numpy()[0], axis=0))[0]
        np.testing.assert_allclose(-orc.dot3(rays[:-1], nrm), orc.dot3(rays[1:], nrm), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("assume_quads,use_mask,dense", [(False, False, False), (False, False, True),
                                                         (True, False, False), (False, True, False),
                                                         (True, True, True)])
def test_k6_bit_exact_vs_oracle(drt, rng, order, assume_quads, use_mask, dense):
    v, t = scenes.urban_grid(3, 3)
    tx = np.array([[30.0, 30.0, 50.0], [0.0, 60.0, 45.0]], np.float32)
    rx = scenes.receivers_grid(v, 4, 3)
    n = t.shape[0] // 2 if assume_quads else t.shape[0]
    cand = scenes.sampled_candidates(n, order, 300) * (2 if assume_quads else 1)
    # make sure some valid paths exist: add the exhaustive order-1 set for the first receiver rows
    if order == 1:
        cand = np.concatenate((cand, (np.arange(n, dtype=np.int32) * (2 if assume_quads else 1))[:, None]))
    mask = rng.uniform(size=t.shape[0]) > 0.15 if use_mask else None
    if mask is not None and assume_quads:
        mask = np.repeat(mask[::2], 2)
    ev, eo, em, st = co.trace_path_candidates(v, t, tx, rx, cand, mask=mask, assume_quads=assume_quads, stages=True)
    mesh = drt.Mesh.from_numpy(v, t, mask, assume_quads)
    got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense, with_stats=True)
    np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
    np.testing.assert_array_equal(got.objects.cpu().numpy(), eo)
    np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
    assert got.interaction_types.shape == (2, 12, cand.shape[0], order) and not got.interaction_types.any()
    prevalid = st["inside"] & st["same_side"] & ~st["too_small"] & st["finite"]
    if dense:
        assert got.stats["tests_done"] > 0
    elif mask is None:
        assert got.stats["candidates_blockage_tested"] == int(prevalid.sum())
    if order == 1 and not use_mask:
        assert em.any()
    m = got.masked()
    np.testing.assert_array_equal(bits(m.vertices.cpu().numpy()), bits(ev[em]))
    np.testing.assert_array_equal(m.objects.cpu().numpy(), eo[em])


def test_k6_order_five_and_seven(drt, rng):
    # K+1 = 6 uses the path-per-warp kernel, K+1 = 8 the generic segment kernel
    v, t = scenes.urban_grid(2, 2)
    tx = np.array([[15.0, 15.0, 60.0]], np.float32)
    rx = scenes.receivers_grid(v, 2)
    for order in (5, 7):
        cand = scenes.sampled_candidates(t.shape[0], order, 64)
        ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand)
        for dense in (False, True):
            got = drt.trace_path_candidates(drt.Mesh.from_numpy(v, t), tx, rx, cand, dense_blockage=dense)
            np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
            np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))


def test_k6_edge_cases(drt):
    v, t = scenes.box(with_top=True)
    mesh = drt.Mesh.from_numpy(v, t)
    tx = np.array([[2.0, 0.0, 0.0]], np.float32)
    rx = np.array([[3.0, 0.5, 0.2], [-3.0, 0.0, 0.0]], np.float32)
    # zero candidates (_solvers.py:566-573)
    p = drt.trace_path_candidates(mesh, tx, rx, np.empty((0, 2), np.int32))
    assert p.vertices.shape == (1, 2, 0, 4, 3) and p.mask.shape == (1, 2, 0) and p.masked().vertices.shape[0] == 0
    # order 0 = line of sight: first rx visible, second behind the cube
    p = drt.trace_paths(mesh, tx, rx, 0)
    np.testing.assert_array_equal(p.mask.cpu().numpy().ravel(), [True, False])
    ev, eo, em = orc.trace_path_candidates(v, t, tx, rx, np.empty((1, 0), np.int32))
    np.testing.assert_array_equal(p.mask.cpu().numpy(), em)
    np.testing.assert_array_equal(p.objects.cpu().numpy(), eo)
    # all triangles masked out ⇒ nothing valid at order 1 (test_scene.py:649-678)
    mesh0 = drt.Mesh.from_numpy(v, t, np.zeros(12, bool))
    assert drt.trace_paths(mesh0, tx, rx, 1).num_valid_paths == 0
    # empty mesh, order 0: always line of sight
    empty = drt.Mesh.from_numpy(np.empty((0, 3), np.float32), np.empty((0, 3), np.int32))
    assert drt.trace_paths(empty, tx, rx, 0).mask.all()


def test_k6_masked_mesh_equals_submesh(drt, two_buildings, kats):  # test_scene.py:585-647
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    mask = np.random.default_rng(7).uniform(size=t.shape[0]) > 0.3
    keep = np.nonzero(mask)[0]
    sub = drt.trace_paths(drt.Mesh.from_numpy(v, t[keep]), tx, rx, 2)
    cand_full = keep[sub.objects[0, 0, :, 1:-1].cpu().numpy()].astype(np.int32)
    full = drt.trace_path_candidates(drt.Mesh.from_numpy(v, t, mask), tx, rx, cand_full)
    np.testing.assert_array_equal(full.mask.cpu().numpy(), sub.mask.cpu().numpy())
    np.testing.assert_array_equal(bits(full.vertices.cpu().numpy()), bits(sub.vertices.cpu().numpy()))


@pytest.mark.parametrize("order", [0, 1, 3])
def test_k6b_vjp_vs_oracle(drt, rng, order):
    v, t = scenes.urban_grid(2, 2)
    tx = np.array([[15.0, 15.0, 60.0], [40.0, -5.0, 55.0]], np.float32)
    rx = scenes.receivers_grid(v, 3, 2)
    cand = scenes.sampled_candidates(t.shape[0], order, 40) if order else np.empty((1, 0), np.int32)
    mesh = drt.Mesh.from_numpy(v, t)
    mesh.vertices.requires_grad_(True)
    txc = torch.from_numpy(tx).cuda().requires_grad_(True)
    rxc = torch.from_numpy(rx).cuda().requires_grad_(True)
    paths = drt.trace_path_candidates(mesh, txc, rxc, cand)
    g = rng.normal(size=tuple(paths.vertices.shape)).astype(np.float32)
    (paths.vertices * torch.from_numpy(g).cuda()).sum().backward()
    gtx, grx, gV = orc.trace_vjp(v, t, tx, rx, cand, g)
    for got, exp in ((txc.grad, gtx), (rxc.grad, grx), (mesh.vertices.grad, gV)):
        scale = max(np.abs(exp).max(), 1.0)
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=2e-3, atol=2e-5 * scale)


def test_k6b_gradient_only_through_valid_paths(drt):
    # the masked (valid) paths keep the autograd graph; a loss on them reaches tx
    v, t = scenes.urban_grid(2, 2)
    mesh = drt.Mesh.from_numpy(v, t)
    tx = torch.tensor([[15.0, 15.0, 60.0]], device="cuda", requires_grad=True)
    rx = torch.from_numpy(scenes.receivers_grid(v, 3)).cuda()
    paths = drt.trace_paths(mesh, tx, rx, 1)
    valid = paths.masked()
    assert valid.vertices.shape[0] > 0
    valid.vertices.sum().backward()
    assert tx.grad is not None and torch.isfinite(tx.grad).all() and tx.grad.abs().sum() > 0


def test_compaction_order_and_capacity(drt, rng):
    from differt_b200 import _lib
    from differt_b200._tensor import ptr, stream_ptr

    P, k = 70_001, 2
    mask = rng.uniform(size=P) < 0.03
    verts = rng.normal(size=(P, k + 2, 3)).astype(np.float32)
    objs = rng.integers(0, 1000, size=(P, k + 2)).astype(np.int32)
    paths = drt.TracedPaths(torch.from_numpy(verts).cuda(), torch.from_numpy(objs).cuda(),
                            torch.from_numpy(mask).cuda(), torch.zeros((P, k), dtype=torch.int32, device="cuda"))
    m = paths.masked()
    np.testing.assert_array_equal(m.vertices.cpu().numpy(), verts[mask])
    np.testing.assert_array_equal(m.objects.cpu().numpy(), objs[mask])
    # fused gather with a capacity smaller than the number of survivors
    cap = 100
    dv, do, dm = paths.vertices, paths.objects, paths.mask.to(torch.uint8)
    ws = torch.empty(_lib.lib.drt_compact_workspace_bytes(P), dtype=torch.uint8, device="cuda")
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    idx = torch.full((cap,), -1, dtype=torch.int64, device="cuda")
    ov = torch.zeros((cap, k + 2, 3), dtype=torch.float32, device="cuda")
    oo = torch.zeros((cap, k + 2), dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib.drt_compact_valid_paths(stream_ptr(), P, k, ptr(dv), ptr(do), ptr(dm), cap, ptr(ws),
                                                ws.numel(), ptr(count), ptr(idx), ptr(ov), ptr(oo)))
    assert int(count.item()) == int(mask.sum())
    np.testing.assert_array_equal(idx.cpu().numpy(), np.nonzero(mask)[0][:cap])
    np.testing.assert_array_equal(ov.cpu().numpy(), verts[mask][:cap])
    np.testing.assert_array_equal(oo.cpu().numpy(), objs[mask][:cap])


@pytest.mark.parametrize("n,order", [(5, 1), (5, 3), (24, 2), (12, 4), (1, 1), (1, 2)])
def test_candidate_decode_matches_host_enumeration(drt, n, order):
    got = drt.generate_all_path_candidates(n, order).cpu().numpy()
    exp = scenes.complete_graph_candidates(n, order)
    np.testing.assert_array_equal(got, exp)
    assert got.shape[0] == scenes.num_complete_graph_candidates(n, order)
    if got.shape[0] and order > 1:
        assert (np.diff(got, axis=1) != 0).all()
    part = drt.generate_all_path_candidates(n, order, start=3, count=7, assume_quads=True).cpu().numpy()
    np.testing.assert_array_equal(part, 2 * exp[3:10])


# ------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE sizes (the oracle cannot reach these in seconds)
# ------------------------------------------------------------------------------------------------


def test_properties_at_scale(drt, rng):
    v, t = scenes.urban_grid(29, 29)  # 10 094 triangles
    tri = torch.from_numpy(orc.triangle_vertices(v, t)).cuda()
    o, d = scene_rays(rng, v, 200_000)
    oc, dc = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    any_hit = drt.ray_intersect_any_triangle(oc, dc, tri)
    idx, tt = drt.first_triangle_hit_by_ray(oc, dc, tri)
    thr = np.float32(1.0) - np.float32(100 * orc.EPS)
    # any-hit ⇔ the nearest hit lies before the threshold
    np.testing.assert_array_equal(any_hit.cpu().numpy(), (tt.cpu().numpy() < thr))
    # permutation invariance of both reductions
    perm = torch.from_numpy(rng.permutation(t.shape[0])).cuda()
    np.testing.assert_array_equal(drt.ray_intersect_any_triangle(oc, dc, tri[perm]).cpu().numpy(),
                                  any_hit.cpu().numpy())
    idx2, tt2 = drt.first_triangle_hit_by_ray(oc, dc, tri[perm], batch_size=None)
    np.testing.assert_array_equal(bits(tt2.cpu().numpy()), bits(tt.cpu().numpy()))
    # the winning triangle reproduces t through the element-wise kernel
    sel = (idx >= 0).nonzero().squeeze(-1)[:50_000]
    t_el, hit_el = drt.ray_intersect_triangle(oc[sel], dc[sel], tri[idx[sel].long()])
    assert hit_el.all()
    np.testing.assert_array_equal(bits(t_el.cpu().numpy()), bits(tt[sel].cpu().numpy()))
    # spot-check 2 000 rays against the C oracle
    exp = co.ray_intersect_any_triangle(o[:2000], d[:2000], tri.cpu().numpy())
    np.testing.assert_array_equal(any_hit[:2000].cpu().numpy(), exp)


def test_trace_dense_equals_pruned_at_scale(drt):
    v, t = scenes.urban_grid(29, 29)
    mesh = drt.Mesh.from_numpy(v, t)
    tx = np.array([[420.0, 420.0, 48.0]], np.float32)
    rx = scenes.receivers_grid(v, 16)
    cand = scenes.sampled_candidates(t.shape[0], 3, 512)
    a = drt.trace_path_candidates(mesh, tx, rx, cand, with_stats=True)
    b = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=True, with_stats=True)
    assert torch.equal(a.mask, b.mask) and torch.equal(a.vertices, b.vertices)
    assert b.stats["tests_done"] >= a.stats["tests_done"]
    ev, eo, em = co.trace_path_candidates(v, t, tx, rx[:8], cand[:64], early_exit=True)
    np.testing.assert_array_equal(a.mask[:, :8, :64].cpu().numpy(), em)
    np.testing.assert_array_equal(bits(a.vertices[:, :8, :64].cpu().numpy()), bits(ev))


# ------------------------------------------------------------------------------------------------
# the benchmarked code path: dense batches of >= 262 144 paths take the batch-specific ordering pass
# (hit counts on a sample, greedy set-cover rounds with partial re-sorts) and the culled cascade —
# code that smaller batches never reach.  Compared against the C oracle like everything else.
# ------------------------------------------------------------------------------------------------


def _bench_like_case(grid, order, n_rx, n_cand, seed=1234):
    v, t = scenes.urban_grid(*grid)
    lo, hi = v.min(0), v.max(0)
    tx = np.array([[0.5 * (lo[0] + hi[0]) + 15.0, 0.5 * (lo[1] + hi[1]) + 15.0, 1.2 * hi[2]]], np.float32)
    rx = scenes.receivers_grid(v, *n_rx)
    cand = scenes.sampled_candidates(t.shape[0], order, n_cand, seed=seed)
    return v, t, tx, rx, cand


def test_trace_bench_scale_ordering_pass_bit_exact_vs_oracle(drt, golden_dir):
    """BASELINE config 3 geometry (10 094 triangles, order 3) at P = 512 x 1024 = 524 288 paths, incl.
    the candidates known to be valid: the FULL mask, vertices and objects against the C oracle."""
    v, t, tx, rx, cand = _bench_like_case((29, 29), 3, (32, 16), 1024)
    known = np.load(golden_dir / "urban10k_valid_candidates.npz")["order3"]
    cand[: known.shape[0]] = known
    mesh = drt.Mesh.from_numpy(v, t)
    got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=True, with_stats=True)
    assert got.stats["ordering_pass"] >= 2 or got.stats["culled_pass"], "the batch must take the bench's code path"
    ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True)
    np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
    np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
    np.testing.assert_array_equal(got.objects.cpu().numpy(), eo)
    # the default (pruned) mode of the API must agree as well
    pruned = drt.trace_path_candidates(mesh, tx, rx, cand)
    assert torch.equal(pruned.mask, got.mask) and torch.equal(pruned.vertices, got.vertices)


@pytest.mark.parametrize("grid,order,n_rx,n_cand,quads,use_mask", [
    ((29, 29), 3, (32, 16), 768, True, False),    # assume_quads
    ((29, 29), 2, (32, 16), 640, False, True),    # masked mesh, order 2
    ((64, 65), 4, (32, 16), 512, False, False),   # BASELINE config 5 geometry: 49 922 triangles, order 4
    ((29, 29), 1, (64, 64), 64, False, False),    # order 1: P = 262 144 exactly at the threshold
])
def test_trace_ordering_pass_variants_vs_oracle(drt, rng, grid, order, n_rx, n_cand, quads, use_mask):
    """Other shapes of the same branch; the oracle checks a strided block of receivers (the GPU traced
    all of them in one dense batch)."""
    v, t, tx, rx, cand = _bench_like_case(grid, order, n_rx, n_cand, seed=7)
    if quads:
        cand = (cand // 2 * 2).astype(np.int32)
        for j in range(1, order):  # re-draw consecutive repeats created by the rounding
            same = cand[:, j] == cand[:, j - 1]
            cand[same, j] = (cand[same, j] + 2) % (t.shape[0] // 2 * 2)
    mask = (rng.uniform(size=t.shape[0]) <= 0.7) if use_mask else None
    mesh = drt.Mesh.from_numpy(v, t, mask=mask, assume_quads=quads)
    got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=True, with_stats=True)
    assert got.stats["ordering_pass"] >= 1 or got.stats["culled_pass"]
    sel = np.arange(0, rx.shape[0], 32 if grid == (64, 65) else 8)
    ev, eo, em = co.trace_path_candidates(v, t, tx, rx[sel], cand, mask=mask, assume_quads=quads,
                                          early_exit=True)
    np.testing.assert_array_equal(got.mask[:, sel].cpu().numpy(), em)
    np.testing.assert_array_equal(bits(got.vertices[:, sel].cpu().numpy()), bits(ev))
    pruned = drt.trace_path_candidates(mesh, tx, rx, cand)
    assert torch.equal(pruned.mask, got.mask)


# ------------------------------------------------------------------------------------------------
# sharded trace + gather (world size 1 here; the world-size-2 merge logic is covered on CPU with gloo)
# ------------------------------------------------------------------------------------------------


def test_sharded_trace_gathers_valid_paths_in_reference_order(drt):
    from differt_b200.distributed import trace_path_candidates_sharded

    v, t = scenes.street_canyon(3)
    mesh = drt.Mesh.from_numpy(v, t)
    tx = np.array([[10.0, 0.0, 30.0], [12.0, 1.0, 25.0]], np.float32)
    rx = np.array([[x, y, 1.5] for x in (2.0, 11.0, 19.0) for y in (-6.0, 5.0)], np.float32)
    cand = scenes.complete_graph_candidates(t.shape[0], 2)
    for capacity in (1 << 12, 2):  # 2 forces the overflow → retry path
        paths, valid = trace_path_candidates_sharded(mesh, tx, rx, torch.from_numpy(cand).cuda(),
                                                     capacity=capacity, dense_blockage=True)
        ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True)
        idx = np.flatnonzero(em.reshape(-1))
        assert idx.size > 0
        np.testing.assert_array_equal(valid.index.cpu().numpy(), idx)
        np.testing.assert_array_equal(bits(valid.vertices.cpu().numpy()), bits(ev.reshape(-1, 4, 3)[idx]))
        np.testing.assert_array_equal(valid.objects.cpu().numpy(), eo.reshape(-1, 4)[idx])
        m = paths.masked()
        assert torch.equal(m.vertices, valid.vertices) and torch.equal(m.objects, valid.objects)


def test_area_sorted_pack_is_a_permutation_and_keeps_any_hit_results(drt, rng):
    from differt_b200._lib import check, lib
    from differt_b200._tensor import ptr, stream_ptr
    from differt_b200.geometry import pack_triangle_vertices, sort_pack_by_area

    v, t = scenes.urban_grid(7, 7)
    T = t.shape[0]
    mask = rng.uniform(size=T) < 0.7
    tri = torch.from_numpy(orc.triangle_vertices(v, t)).cuda()
    pack = pack_triangle_vertices(tri, torch.from_numpy(mask.astype(np.uint8)).cuda())
    srt = sort_pack_by_area(pack, T)
    a = pack.view(torch.float32).view(-1, 12).cpu().numpy()
    b = srt.view(torch.float32).view(-1, 12).cpu().numpy()
    # same multiset of records (NaN-origin never-hit records compare through their bit patterns)
    key = lambda x: sorted(map(bytes, np.ascontiguousarray(x).view(np.uint8).reshape(x.shape[0], -1)))
    assert key(a) == key(b)
    live = ~np.isnan(b[:, 0])
    assert live.sum() == mask.sum() and live[: live.sum()].all(), "never-hit records must sort last"
    e1, e2 = b[live, 3:6].astype(np.float64), b[live, 6:9].astype(np.float64)
    area2 = (np.cross(e1, e2) ** 2).sum(-1)
    assert (np.diff(area2) <= 1e-6 * area2[:-1]).all(), "areas must be non-increasing"
    # identical any-hit answers from both orders, and equal to the oracle
    o, d = scene_rays(rng, v, 5000)
    oc, dc = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    outs = []
    for pk in (pack, srt):
        out = torch.empty(5000, dtype=torch.uint8, device="cuda")
        check(lib.drt_ray_intersect_any_triangle(stream_ptr(), 5000, ptr(oc), ptr(dc), ptr(pk), T,
                                                 10 * orc.EPS, 100 * orc.EPS, ptr(out), None))
        outs.append(out.cpu().numpy().astype(bool))
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], co.ray_intersect_any_triangle(o, d, tri.cpu().numpy(), mask))


# ------------------------------------------------------------------------------------------------
# N1b: visibility-pruned (HybridPathTracer) candidates decoded on the device
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("n,order", [(1, 1), (2, 2), (6, 1), (6, 2), (7, 3), (9, 4), (5, 5)])
@pytest.mark.parametrize("use_masks", [False, True])
def test_digraph_candidates_match_reference_dfs(drt, rng, n, order, use_masks):
    from oracle.graph_oracle import hybrid_path_candidates

    if use_masks:
        a, b, m = rng.uniform(size=n) < 0.6, rng.uniform(size=n) < 0.6, rng.uniform(size=n) < 0.8
    else:
        a = b = np.ones(n, bool)
        m = None
    exp = hybrid_path_candidates(n, order, a, b, m)
    gen = drt.VisiblePathCandidates(n, order, None if not use_masks else a, None if not use_masks else b, m)
    assert len(gen) == exp.shape[0]
    got = gen.chunk().cpu().numpy()
    np.testing.assert_array_equal(got.reshape(exp.shape), exp)
    # chunks decode independently and concatenate to the full list
    if len(gen):
        parts = [c.cpu().numpy() for c in gen.chunks_iter(5)]
        np.testing.assert_array_equal(np.concatenate(parts), exp)
    # quads: indices are the even triangles
    gen2 = drt.VisiblePathCandidates(n, order, a if use_masks else None, b if use_masks else None, m,
                                     assume_quads=True)
    np.testing.assert_array_equal(gen2.chunk().cpu().numpy().reshape(exp.shape), 2 * exp)


def test_digraph_order_zero_and_empty(drt):
    assert len(drt.VisiblePathCandidates(5, 0, None, None)) == 0  # no direct path (graph.rs:1076-1092)
    assert drt.VisiblePathCandidates(5, 0, None, None).chunk().shape == (0, 0)
    assert len(drt.VisiblePathCandidates(4, 2, np.zeros(4, bool), None)) == 0
    assert len(drt.VisiblePathCandidates(0, 2, None, None)) == 0


@pytest.mark.parametrize("assume_quads", [False, True])
@pytest.mark.parametrize("order", [1, 2])
def test_hybrid_trace_finds_the_golden_paths(drt, kats, two_buildings, order, assume_quads):
    """test_scene.py:116-160 with method="hybrid": the visibility-pruned candidate set still
    contains the valid paths, and is (much) smaller than the exhaustive one."""
    v, t = two_buildings
    g = kats["two_buildings_scene"]
    mesh = drt.Mesh.from_numpy(v, t, assume_quads=assume_quads)
    tx, rx = np.array(g["tx"], np.float32), np.array(g["rx"], np.float32)
    hyb = drt.trace_paths(mesh, tx, rx, order, solver="hybrid", num_rays=100_000)
    exh = drt.trace_paths(mesh, tx, rx, order, solver="exhaustive")
    assert 0 < hyb.mask.numel() < exh.mask.numel()
    mh, me = hyb.masked(), exh.masked()
    assert torch.equal(mh.objects, me.objects) and torch.equal(mh.vertices, me.vertices)
    assert mh.vertices.shape[0] >= 1


# ------------------------------------------------------------------------------------------------
# N3: shooting-and-bouncing rays (SBRPathLauncher.launch_paths) and the multipath lifetime map
# ------------------------------------------------------------------------------------------------


def _c_first_hit(tri, mask=None):
    return lambda o, d: co.first_triangle_hit_by_ray(o, d, tri, mask, batch_size=512)


@pytest.mark.parametrize("order", [0, 1, 3])
def test_sbr_launch_paths_bit_exact(drt, order):
    v, t = scenes.street_canyon(4)
    mesh = drt.Mesh.from_numpy(v, t)
    tx = np.array([[15.0, 0.0, 25.0], [5.0, 2.0, 12.0]], np.float32)
    rx = np.array([[x, y, 1.5] for x in (2.0, 11.0, 19.0, 33.0) for y in (-6.0, 5.0)], np.float32)
    tri = orc.triangle_vertices(v, t)
    _, dirs = orc.sbr_launch_rays(tri, tx, rx, 3000)
    got = drt.launch_paths(mesh, tx, rx, order, ray_directions=dirs, max_dist=4.0)
    cand, verts, masks = orc.sbr_launch_paths(v, t, tx, rx, dirs, order, max_dist=4.0, first_hit=_c_first_hit(tri))
    np.testing.assert_array_equal(got.masks.cpu().numpy(), masks)
    np.testing.assert_array_equal(got.ray_objects.cpu().numpy(), cand)
    np.testing.assert_array_equal(bits(got.ray_vertices.cpu().numpy()), bits(verts))
    assert masks.any(), "the scene must produce receivers in the vicinity of some rays"
    # the dense views of the reference and the compacted per-order paths agree
    assert got.vertices.shape == (2, 8, 3000, order + 2, 3) and got.objects.shape == (2, 8, 3000, order + 2)
    for k in range(order + 1):
        p = got.get_paths(k)
        idx = np.argwhere(masks[..., k])
        assert p.vertices.shape[0] == idx.shape[0]
        if idx.shape[0]:
            np.testing.assert_array_equal(p.objects[:, 0].cpu().numpy(), idx[:, 0])
            np.testing.assert_array_equal(p.objects[:, -1].cpu().numpy(), idx[:, 1])
            np.testing.assert_array_equal(p.objects[:, 1:-1].cpu().numpy(), cand[idx[:, 0], idx[:, 2], :k])


def test_sbr_ray_generation_and_golden_scene(drt, kats, two_buildings):
    """launch_rays matches the oracle's frustum + lattice, and on the reference's two_buildings scene
    SBR finds the golden order-1 interaction (test_scene.py:116-160 with method="sbr", which the
    reference itself only checks at rtol=1)."""
    v, t = two_buildings
    g = kats["two_buildings_scene"]
    mesh = drt.Mesh.from_numpy(v, t)
    tx, rx = np.array(g["tx"], np.float32).reshape(1, 3), np.array(g["rx"], np.float32).reshape(1, 3)
    o, d = drt.launch_rays(mesh, tx, rx, 5000)
    eo, ed = orc.sbr_launch_rays(orc.triangle_vertices(v, t), tx, rx, 5000)
    np.testing.assert_allclose(d.cpu().numpy(), ed, rtol=0, atol=2e-6)  # libm vs CUDA sin/cos/acos
    got = drt.launch_paths(mesh, tx, rx, 1, num_rays=200_000, max_dist=1e-1)
    objs = got.get_paths(1).objects[:, 1].unique().cpu().numpy()
    expected = np.array(g["orders"]["1"]["objects"]).reshape(-1, 3)[:, 1]  # [tx, triangle, rx] rows
    assert np.isin(expected - expected % 2, objs - objs % 2).all()


def test_mlm_matches_oracle_and_hash_constants(drt):
    v, t = scenes.street_canyon(3)
    mesh = drt.Mesh.from_numpy(v, t)
    tx = np.array([[10.0, 0.0, 20.0], [14.0, 3.0, 8.0]], np.float32)
    tri = orc.triangle_vertices(v, t)
    rng = np.random.default_rng(5)
    dirs = rng.normal(size=(2, 4000, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    kw = dict(max_order=2, min_order=0, dim_x=9, dim_y=7, receiver_height=1.5, min_x=-10.0, max_x=30.0,
              min_y=-25.0, max_y=25.0)
    got = drt.compute_tx_mlm(mesh, tx, ray_directions=dirs, **kw).cpu().numpy().astype(np.uint32)
    exp = orc.compute_tx_mlm(v, t, tx, dirs, assume_quads=False, first_hit=_c_first_hit(tri), **kw)
    np.testing.assert_array_equal(got, exp)
    assert (got != 0).sum() > 10
    # line-of-sight cells carry the FNV offset basis 0x811C9DC5 = 2166136261 (test_scene.py:910)
    los = drt.compute_tx_mlm(mesh, tx, ray_directions=dirs, **{**kw, "max_order": 0}).cpu().numpy()
    assert set(np.unique(los)) <= {0, 2166136261} and (los == 2166136261).any()
    # min_order filters the direct paths out
    no_los = drt.compute_tx_mlm(mesh, tx, ray_directions=dirs, **{**kw, "min_order": 1}).cpu().numpy()
    exp2 = orc.compute_tx_mlm(v, t, tx, dirs, assume_quads=False, first_hit=_c_first_hit(tri), **{**kw, "min_order": 1})
    np.testing.assert_array_equal(no_los.astype(np.uint32), exp2)
    # generated rays: runs end to end and is deterministic
    a = drt.compute_tx_mlm(mesh, tx, num_rays=20000, **kw)
    b = drt.compute_tx_mlm(mesh, tx, num_rays=20000, **kw)
    assert torch.equal(a, b) and int((a != 0).sum()) > 0


# ------------------------------------------------------------------------------------------------
# N2: opt-in BVH — same answers as the brute-force kernels on the test scenes
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("scene", ["urban", "bruxelles", "box", "single"])
@pytest.mark.parametrize("use_mask", [False, True])
def test_bvh_queries_equal_brute_force(drt, rng, bruxelles, scene, use_mask):
    if scene == "urban":
        v, t = scenes.urban_grid(29, 29)
    elif scene == "bruxelles":
        v, t = bruxelles
    elif scene == "box":
        v, t = scenes.box(2.0, 3.0, 4.0, with_top=True)
    else:
        v, t = scenes.box(2.0, 3.0, 4.0, with_top=True)
        t = t[:1]
    mask = (rng.uniform(size=t.shape[0]) < 0.6) if use_mask else None
    mesh = drt.Mesh.from_numpy(v, t, mask=mask)
    n = 200_000 if t.shape[0] > 100 else 20_000
    lo, hi = v.min(0) - 1.0, v.max(0) + 1.0
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    e = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    oc, dc = torch.from_numpy(o).cuda(), torch.from_numpy(e - o).cuda()
    a0 = mesh.ray_intersect_any_triangle(oc, dc)
    a1 = mesh.ray_intersect_any_triangle(oc, dc, accel="bvh")
    assert torch.equal(a0, a1) and (scene == "single" or bool(a0.any()))
    for bs in (512, None):
        i0, t0 = mesh.first_triangle_hit_by_ray(oc, dc, batch_size=bs)
        i1, t1 = mesh.first_triangle_hit_by_ray(oc, dc, batch_size=bs, accel="bvh")
        assert torch.equal(t0.view(torch.int32), t1.view(torch.int32))
        assert torch.equal(i0, i1)
    # and against the oracle on a sub-sample
    tri = orc.triangle_vertices(v, t)
    ei, et = co.first_triangle_hit_by_ray(o[:3000], dc[:3000].cpu().numpy(), tri, mask, batch_size=None)
    np.testing.assert_array_equal(i1[:3000].cpu().numpy(), ei)  # i1: last loop iteration, batch_size=None
    np.testing.assert_array_equal(bits(t1[:3000].cpu().numpy()), bits(et))


def test_bvh_visibility_sbr_and_hybrid_match_brute_force(drt, kats, two_buildings):
    v, t = scenes.urban_grid(7, 7)
    mesh = drt.Mesh.from_numpy(v, t)
    vx = np.array([[90.0, 95.0, 50.0], [15.0, 15.0, 1.5]], np.float32)
    assert torch.equal(mesh.triangles_visible_from_vertex(vx, num_rays=50_000),
                       mesh.triangles_visible_from_vertex(vx, num_rays=50_000, accel="bvh"))
    tx, rx = vx[:1], np.array([[45.0, 15.0, 1.5], [105.0, 75.0, 1.5]], np.float32)
    a = drt.launch_paths(mesh, tx, rx, 2, num_rays=30_000, max_dist=1.0)
    b = drt.launch_paths(mesh, tx, rx, 2, num_rays=30_000, max_dist=1.0, accel="bvh")
    assert torch.equal(a.masks, b.masks) and torch.equal(a.ray_objects, b.ray_objects)
    assert torch.equal(a.ray_vertices.view(torch.int32), b.ray_vertices.view(torch.int32))
    kw = dict(max_order=2, dim_x=8, dim_y=8, num_rays=30_000, receiver_height=1.5, min_x=-15.0, max_x=195.0,
              min_y=-15.0, max_y=195.0)
    assert torch.equal(drt.compute_tx_mlm(mesh, tx, **kw), drt.compute_tx_mlm(mesh, tx, accel="bvh", **kw))
    # empty mesh and degenerate rays
    empty = drt.Mesh(torch.zeros((0, 3)), torch.zeros((0, 3), dtype=torch.int32))
    assert not bool(empty.ray_intersect_any_triangle(vx, vx, accel="bvh").any())
    z = torch.zeros((5, 3)).cuda()
    i1, t1 = mesh.first_triangle_hit_by_ray(z + 50.0, z, accel="bvh")
    assert bool((i1 == -1).all()) and bool(torch.isinf(t1).all())


# ------------------------------------------------------------------------------------------------
# the exactness fallbacks of the inlined reciprocal (|a| >= 2^126, epsilon below FLT_MIN, inf / NaN)
# ------------------------------------------------------------------------------------------------


def _extreme_scene(rng, n_tri=700, n_rays=3000):
    """Triangles and rays whose determinant `a` spans the whole float range: coordinates from 1e-12 to
    1e14 (|a| up to ~1e38+, overflowing to inf), plus rays with inf / NaN components."""
    scale_t = 10.0 ** rng.uniform(-12, 14, size=(n_tri, 1, 1))
    tri = (rng.normal(size=(n_tri, 3, 3)) * scale_t).astype(np.float32)
    scale_r = 10.0 ** rng.uniform(-12, 14, size=(n_rays, 1))
    o = (rng.normal(size=(n_rays, 3)) * scale_r).astype(np.float32)
    d = (rng.normal(size=(n_rays, 3)) * scale_r * 10.0).astype(np.float32)
    # aim a third of the rays at a triangle centroid so that real hits exist at every magnitude
    aim = rng.integers(0, n_tri, size=n_rays // 3)
    c = tri[aim].mean(axis=1)
    o[: aim.size] = (c - d[: aim.size] * np.float32(0.5)).astype(np.float32)
    d[-5:, 0] = np.inf
    o[-10:-5, 1] = np.nan
    d[-15:-10] = 0.0
    return tri, o, d


@pytest.mark.parametrize("epsilon", [None, 1e-42, 0.0, -1.0])
def test_any_and_first_hit_bit_exact_on_extreme_magnitudes(drt, rng, epsilon):
    tri, o, d = _extreme_scene(rng)
    with np.errstate(all="ignore"):
        exp_any = co.ray_intersect_any_triangle(o, d, tri, epsilon=epsilon)
        ei, et = co.first_triangle_hit_by_ray(o, d, tri, epsilon=epsilon)
    kw = {} if epsilon is None else {"epsilon": epsilon}
    got_any = drt.ray_intersect_any_triangle(o, d, tri, **kw)
    np.testing.assert_array_equal(got_any.numpy(), exp_any)
    gi, gt = drt.first_triangle_hit_by_ray(o, d, tri, **kw)
    np.testing.assert_array_equal(gi.numpy(), ei)
    np.testing.assert_array_equal(bits(gt.numpy()), bits(et))
    assert exp_any.any() and (ei >= 0).sum() > 100
    # the sorted-pack path (>= 4096 rays) as well
    o2, d2 = np.tile(o, (2, 1)), np.tile(d, (2, 1))
    np.testing.assert_array_equal(drt.ray_intersect_any_triangle(o2, d2, tri, **kw).numpy(), np.tile(exp_any, 2))


def test_trace_bit_exact_on_extreme_magnitudes(drt, rng):
    """Dense and pruned blockage on a mesh whose determinants overflow the fast reciprocal's range."""
    n_tri = 600
    # mostly small triangles scattered over +-1e3, a few astronomically large ones far away
    scale = np.where(rng.uniform(size=(n_tri, 1, 1)) < 0.97, 10.0 ** rng.uniform(-3, 1.5, size=(n_tri, 1, 1)),
                     10.0 ** rng.uniform(9, 13, size=(n_tri, 1, 1)))
    centre = rng.normal(size=(n_tri, 1, 3)) * np.where(scale < 1e3, 1e3, 1e15)
    tv = (centre + rng.normal(size=(n_tri, 3, 3)) * scale).astype(np.float32)
    v = tv.reshape(-1, 3)
    t = np.arange(3 * n_tri, dtype=np.int32).reshape(n_tri, 3)
    tx = (rng.normal(size=(2, 3)) * 1e3).astype(np.float32)
    rx = (rng.normal(size=(5, 3)) * 1e3).astype(np.float32)
    cand = scenes.sampled_candidates(n_tri, 2, 700)
    mesh = drt.Mesh.from_numpy(v, t)
    with np.errstate(all="ignore"):
        ev, eo, em, st = co.trace_path_candidates(v, t, tx, rx, cand, stages=True)
    for dense in (True, False):
        got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense)
        np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
        np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
    assert st["blocked"].any() and (~st["blocked"]).any()


def test_exact_zero_and_tie_cases_on_a_lattice_scene(drt, rng):
    """Integer-coordinate boxes and rays between lattice points: determinants, barycentrics and
    distances hit exact zeros, ones and ties (rays through edges and corners, rays inside the plane
    of a wall, zero-length rays), where signed zeros and the tie rule decide."""
    parts = [scenes.box(4.0, 4.0, 6.0, with_top=True, center=(8.0 * i, 8.0 * j, 3.0)) for i in range(4) for j in range(4)]
    v, t = scenes._merge(parts)
    tri = orc.triangle_vertices(v, t)
    g = np.arange(-4, 30, 2, dtype=np.float32)
    pts = np.stack(np.meshgrid(g, g, np.array([0.0, 3.0, 6.0, 7.0], np.float32), indexing="ij"), -1).reshape(-1, 3)
    a = pts[rng.integers(0, pts.shape[0], 6000)]
    b = pts[rng.integers(0, pts.shape[0], 6000)]
    o, d = a, (b - a).astype(np.float32)
    for bs in (512, 7, None):
        ei, et = co.first_triangle_hit_by_ray(o, d, tri, batch_size=bs)
        gi, gt = drt.first_triangle_hit_by_ray(o, d, tri, batch_size=bs)
        np.testing.assert_array_equal(gi.numpy(), ei)
        np.testing.assert_array_equal(bits(gt.numpy()), bits(et))
    np.testing.assert_array_equal(drt.ray_intersect_any_triangle(o, d, tri).numpy(),
                                  co.ray_intersect_any_triangle(o, d, tri))
    mesh = drt.Mesh.from_numpy(v, t)
    gi2, gt2 = mesh.first_triangle_hit_by_ray(o, d, accel="bvh")
    ei, et = co.first_triangle_hit_by_ray(o, d, tri, batch_size=512)
    # the BVH reaches the same triangles on this scene too (ties resolved by the same key)
    np.testing.assert_array_equal(gi2.cpu().numpy(), ei)
    np.testing.assert_array_equal(bits(gt2.cpu().numpy()), bits(et))
    # trace: lattice tx / rx, every order-2 candidate of a sub-mesh, with and without quads
    tx, rx = pts[[5, 77]], pts[rng.integers(0, pts.shape[0], 40)]
    sub_v, sub_t = scenes._merge(parts[:3])
    for quads in (False, True):
        cand = scenes.complete_graph_candidates(sub_t.shape[0] // (2 if quads else 1), 2) * (2 if quads else 1)
        m = drt.Mesh.from_numpy(sub_v, sub_t, assume_quads=quads)
        ev, eo, em = co.trace_path_candidates(sub_v, sub_t, tx, rx, cand, assume_quads=quads)
        for dense in (True, False):
            got = drt.trace_path_candidates(m, tx, rx, cand, dense_blockage=dense)
            np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
            np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))


# ------------------------------------------------------------------------------------------------
# the culled blockage pass (csrc/cull.cuh) on inputs built to defeat a bounding-volume cull: meshes
# with more than 2048 triangles take it for every candidate the head tiles leave undecided
# ------------------------------------------------------------------------------------------------


def _plane_basis(rng):
    n = rng.normal(size=3)
    n /= np.linalg.norm(n)
    a = np.cross(n, [1.0, 0.0, 0.0])
    a /= np.linalg.norm(a)
    return a, np.cross(n, a), n


def _coplanar_clusters(rng, n_planes=4, clusters_per_plane=3, tri_per_cluster=500):
    """Clusters of small triangles lying in a few tilted planes → (vertices, triangles, planes) with
    planes = [(origin, a, b, normal, [cluster centres in plane coordinates])]."""
    tris, planes = [], []
    for _ in range(n_planes):
        a, b, n = _plane_basis(rng)
        c0 = rng.uniform(-300, 300, 3)
        centres = rng.uniform(-400, 400, (clusters_per_plane, 2))
        for c in centres:
            uv = c + rng.uniform(-5, 5, (tri_per_cluster, 3, 2))
            tris.append(c0 + uv[..., 0:1] * a + uv[..., 1:2] * b)
        planes.append((c0, a, b, n, centres))
    tv = np.concatenate(tris).astype(np.float32)
    t = np.arange(3 * tv.shape[0], dtype=np.int32).reshape(-1, 3)
    return tv.reshape(-1, 3), t, planes


@pytest.mark.parametrize("lift", [0.0, 1e-3, 0.3])
def test_cull_reproduces_noise_hits_of_segments_in_the_plane_of_far_triangles(drt, lift):
    """Segments lying in (lift = 0) or at a tiny angle to (lift > 0, metres of out-of-plane offset over
    hundreds of metres) the plane of triangles that are hundreds of metres AWAY: the determinant is
    rounding noise and the reference's fp32 test reports hits there.  Order-0 paths (one segment
    tx → rx each), 64 x 4096 = 262 144 of them, against 6000 triangles."""
    rng = np.random.default_rng(5)
    v, t, planes = _coplanar_clusters(rng)
    tx, rx = [], []
    for c0, a, b, n, _ in planes:
        for pts, count in ((tx, 16), (rx, 1024)):
            uv = rng.uniform(-2500, 2500, (count, 2))
            off = rng.uniform(-lift, lift, (count, 1))
            pts.append(c0 + uv[:, 0:1] * a + uv[:, 1:2] * b + off * n)
    tx, rx = np.concatenate(tx).astype(np.float32), np.concatenate(rx).astype(np.float32)
    cand = np.zeros((1, 0), np.int32)
    mesh = drt.Mesh.from_numpy(v, t)
    ev, eo, em, st = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True, stages=True)
    blocked = st["blocked"][..., 0]
    # how far from every cluster centre of ITS plane does each same-plane segment pass?
    far_hits = 0
    for k, (c0, a, b, n, centres) in enumerate(planes):
        p_tx = np.stack(((tx[16 * k:16 * k + 16] - c0) @ a, (tx[16 * k:16 * k + 16] - c0) @ b), -1)
        p_rx = np.stack(((rx[1024 * k:1024 * k + 1024] - c0) @ a, (rx[1024 * k:1024 * k + 1024] - c0) @ b), -1)
        o2 = p_tx[:, None, :]
        d2 = p_rx[None, :, :] - o2
        dist = np.full(d2.shape[:2], np.inf)
        for c in centres:
            tt = np.clip(((c - o2) * d2).sum(-1) / np.maximum((d2 * d2).sum(-1), 1e-9), 0.0, 1.0)
            dist = np.minimum(dist, np.linalg.norm(o2 + tt[..., None] * d2 - c, axis=-1))
        far_hits += int((blocked[16 * k:16 * k + 16, 1024 * k:1024 * k + 1024] & (dist > 50.0)).sum())
    if lift == 0.0:
        assert far_hits > 20, "the case must contain hits far away from the triangles that cause them"
    for dense in (True, False):
        got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense, with_stats=True)
        np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
    # the cull did its job on everything it could prove: far fewer tests than the dense count
    assert got.stats["tests_done"] < 0.5 * tx.shape[0] * rx.shape[0] * t.shape[0]


def test_cull_with_degenerate_triangles_extreme_magnitudes_and_masks(drt, rng):
    """3000 triangles: slivers, zero-area and duplicated ones, a few astronomically large or far ones,
    tiny ones, a masked third — everything that makes a node un-cullable or its margin huge."""
    n_tri = 3000
    kind = rng.uniform(size=(n_tri, 1, 1))
    scale = np.where(kind < 0.9, 10.0 ** rng.uniform(-2, 1.5, size=(n_tri, 1, 1)),
                     np.where(kind < 0.95, 10.0 ** rng.uniform(-9, -5, size=(n_tri, 1, 1)),
                              10.0 ** rng.uniform(8, 13, size=(n_tri, 1, 1))))
    centre = rng.normal(size=(n_tri, 1, 3)) * np.where(scale < 1e3, 300.0, 1e14)
    tv = (centre + rng.normal(size=(n_tri, 3, 3)) * scale).astype(np.float32)
    tv[:40, 1] = tv[:40, 0]                                              # zero area
    tv[40:80, 2] = tv[40:80, 0] + (tv[40:80, 1] - tv[40:80, 0]) * 0.5    # collinear
    tv[80:160, 2] = tv[80:160, 1] + np.float32(1e-6)                     # slivers
    tv[160:200] = tv[200:240]                                            # duplicates
    v = tv.reshape(-1, 3)
    t = np.arange(3 * n_tri, dtype=np.int32).reshape(n_tri, 3)
    tx = (rng.normal(size=(4, 3)) * 300).astype(np.float32)
    rx = (rng.normal(size=(64, 3)) * 300).astype(np.float32)
    rx[:8] = v[rng.integers(0, v.shape[0], 8)]                           # receivers ON mesh vertices
    for order, n_cand in ((0, 1), (1, 1000), (2, 500)):
        cand = scenes.sampled_candidates(n_tri, order, n_cand) if order else np.zeros((1, 0), np.int32)
        for mask in (None, rng.uniform(size=n_tri) < 0.66):
            mesh = drt.Mesh.from_numpy(v, t, mask=mask)
            with np.errstate(all="ignore"):
                ev, eo, em, st = co.trace_path_candidates(v, t, tx, rx, cand, mask=mask, early_exit=True, stages=True)
            for dense in (True, False):
                got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense)
                np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
                np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
            comp = drt.trace_valid_path_candidates(mesh, tx, rx, cand)
            np.testing.assert_array_equal(comp.index.cpu().numpy(), np.flatnonzero(em.reshape(-1)))
        if order == 0:
            assert st["blocked"].any() and (~st["blocked"]).any()


@pytest.mark.parametrize("hit_tol,epsilon", [(None, None), (0.0, None), (0.25, 1e-4), (-0.5, None), (None, 0.0)])
def test_cull_parameter_ranges(drt, rng, hit_tol, epsilon):
    """hit_tol / epsilon inside and outside the range the cull's proof covers (outside, the plain cascade
    must run): 10 094 triangles, order-1 candidates through the street level."""
    v, t = scenes.urban_grid(29, 29)
    tx = np.array([[435.0, 435.0, 30.0]], np.float32)
    rx = scenes.receivers_grid(v, 12, 12)
    cand = scenes.sampled_candidates(t.shape[0], 1, 600)
    kw = {}
    if hit_tol is not None:
        kw["hit_tol"] = hit_tol
    if epsilon is not None:
        kw["epsilon"] = epsilon
    mesh = drt.Mesh.from_numpy(v, t)
    ev, eo, em, st = co.trace_path_candidates(v, t, tx, rx, cand, early_exit=True, stages=True, **kw)
    for dense in (True, False):
        got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense, **kw)
        np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
    assert st["blocked"].any() and (~st["blocked"]).any()


def test_prepared_mesh_cache_follows_in_place_updates(drt, rng):
    """The mesh-only part of the trace is cached per mesh STATE (storage + version counter): an
    in-place update of the vertices (what an optimizer step does) or of the mask must be seen by the
    next call, and a cached call must equal an uncached one."""
    v, t = scenes.urban_grid(16, 16)  # 3074 triangles: takes the culled traversal
    tx = np.array([[230.0, 230.0, 45.0]], np.float32)
    rx = scenes.receivers_grid(v, 6, 6)
    cand = scenes.sampled_candidates(t.shape[0], 1, 2000)
    mask = rng.uniform(size=t.shape[0]) < 0.8
    mesh = drt.Mesh.from_numpy(v, t, mask=mask)
    for step in range(3):
        vv = mesh.vertices.cpu().numpy()
        mm = mesh.mask.cpu().numpy()
        ev, eo, em = co.trace_path_candidates(vv, t, tx, rx, cand, mask=mm, early_exit=True)
        for dense in (True, False, True):  # the second and third calls hit the cache
            got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense)
            np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
            np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
        comp = drt.trace_valid_path_candidates(mesh, tx, rx, cand)
        np.testing.assert_array_equal(comp.index.cpu().numpy(), np.flatnonzero(em.reshape(-1)))
        assert em.any() and not em.all()
        if step == 0:  # raise every other building by 15 m, in place
            with torch.no_grad():
                sel = (mesh.vertices[:, 2] > 0) & ((mesh.vertices[:, 0] // 30).long() % 2 == 0)
                mesh.vertices[sel, 2] += 15.0
        else:          # flip a third of the mask, in place
            flip = torch.from_numpy(rng.uniform(size=t.shape[0]) < 0.33).to(mesh.mask.device)
            mesh.mask ^= flip


@pytest.mark.parametrize("solver", ["exhaustive", "hybrid"])
def test_chunked_trace_equals_one_shot_masked(drt, two_buildings, kats, solver):
    """Scene.trace_paths(chunk_size=...) semantics (_scene.py:738-751) and the merged valid paths:
    identical to tracing every candidate at once and calling masked()."""
    v, t = two_buildings
    g = kats["two_buildings_scene"]
    mesh = drt.Mesh.from_numpy(v, t)
    tx = np.array([g["tx"], [1.0, 3.0, 20.0]], np.float32)
    rx = np.array([g["rx"], [0.5, 9.0, 1.5], [-2.0, 12.0, 2.0]], np.float32)
    kw = dict(solver=solver, num_rays=50_000)
    one = drt.trace_paths(mesh, tx, rx, 2, **kw).masked()
    valid = drt.trace_valid_paths(mesh, tx, rx, 2, chunk_size=37, **kw)
    assert one.vertices.shape[0] > 0
    assert torch.equal(valid.vertices, one.vertices) and torch.equal(valid.objects, one.objects)
    chunks = list(drt.trace_paths_chunks_iter(mesh, tx, rx, 2, chunk_size=100, **kw))
    assert sum(int(c.mask.sum()) for c in chunks) == one.vertices.shape[0]
    assert all(c.mask.shape[:2] == (2, 3) and c.mask.shape[2] <= 100 for c in chunks)


# ------------------------------------------------------------------------------------------------
# Scene: the reference's caller-side signatures (test_scene.py:116-160, 334-364, 650-764 error cases)
# ------------------------------------------------------------------------------------------------


def test_scene_call_signatures_and_golden_paths(drt, kats, two_buildings):
    v, t = two_buildings
    g = kats["two_buildings_scene"]
    mesh = drt.Mesh.from_numpy(v, t)
    scene = drt.Scene(np.array(g["tx"], np.float32), np.array(g["rx"], np.float32), mesh)
    assert scene.num_transmitters == 1 and scene.num_receivers == 1
    for order in (0, 1, 2):
        exp = g["orders"][str(order)]
        got = scene.trace_paths(order)
        assert tuple(got.mask.shape) == (scenes.num_complete_graph_candidates(t.shape[0], order),)
        m = got.masked()
        n_valid = int(m.vertices.shape[0])
        exp_v = np.array(exp["vertices"], np.float32).reshape(n_valid, order, 3)
        np.testing.assert_allclose(m.vertices[:, 1:-1].cpu().numpy(), exp_v, rtol=g["rtol"])
        np.testing.assert_array_equal(m.objects.cpu().numpy(), np.array(exp["objects"]).reshape(n_valid, order + 2))
    # explicit candidates == exhaustive (test_scene.py:334-364), quads round the indices down
    cand = scenes.complete_graph_candidates(t.shape[0], 2)
    a, b = scene.trace_paths(2), scene.trace_paths(path_candidates=cand)
    assert torch.equal(a.mask, b.mask) and torch.equal(a.vertices, b.vertices)
    q = scene.set_assume_quads()
    bq = q.trace_paths(path_candidates=cand[:50] | 1)
    assert bool((bq.objects[..., 1:-1] % 2 == 0).all())
    # chunked iterator and hybrid
    chunks = list(scene.trace_paths(2, chunk_size=100))
    assert sum(int(c.mask.sum()) for c in chunks) == int(a.mask.sum()) and len(chunks) == 6
    h = scene.trace_paths(1, solver="hybrid", num_rays=50_000)
    assert torch.equal(h.masked().objects, scene.trace_paths(1).masked().objects)
    # error behaviour of the reference (_scene.py:692-717)
    with pytest.raises(ValueError, match="one of 'order' or `path_candidates`"):
        scene.trace_paths()
    with pytest.raises(ValueError, match="one of 'order' or `path_candidates`"):
        scene.trace_paths(1, path_candidates=cand)
    with pytest.raises(ValueError, match="Unknown solver"):
        scene.trace_paths(1, solver="nope")
    with pytest.raises(ValueError, match="required when using HybridPathTracer"):
        scene.trace_paths(path_candidates=cand, solver="hybrid")
    # batched transmitters / receivers, grids, SBR and MLM entry points
    grid = scene.with_receivers_grid(3, 2).with_transmitters_grid(2, 1, height=30.0)
    assert tuple(grid.receivers.shape) == (2, 3, 3) and tuple(grid.transmitters.shape) == (1, 2, 3)
    p = grid.trace_paths(1)
    assert tuple(p.mask.shape) == (1, 2, 2, 3, t.shape[0]) and tuple(p.vertices.shape[-2:]) == (3, 3)
    lp = grid.launch_paths(1, num_rays=2000, max_dist=1.0)
    assert tuple(lp.masks.shape) == (2, 6, 2000, 2)
    mlm = grid.compute_tx_mlm(1, 4, 5, num_rays=5000)
    assert tuple(mlm.shape) == (1, 2, 4, 5) and int((mlm != 0).sum()) > 0  # cells hold ORs of path hashes


@pytest.mark.parametrize("order", [0, 1, 2, 3, 5])
@pytest.mark.parametrize("assume_quads,use_mask", [(False, False), (True, False), (False, True)])
def test_compact_trace_equals_dense_masked(drt, rng, order, assume_quads, use_mask):
    """drt_trace_valid_path_candidates (no dense outputs) returns exactly masked() of the dense trace,
    including the capacity-overflow retry."""
    v, t = scenes.street_canyon(4)
    mask = (rng.uniform(size=t.shape[0]) < 0.8) if use_mask else None
    if use_mask and assume_quads is False:
        mask[-2:] = True  # keep the ground
    mesh = drt.Mesh.from_numpy(v, t, mask=mask, assume_quads=assume_quads)
    tx = np.array([[15.0, 0.0, 25.0], [5.0, 2.0, 12.0]], np.float32)
    rx = np.array([[x, y, 1.5] for x in (2.0, 11.0, 19.0, 33.0) for y in (-6.0, 5.0)], np.float32)
    n = mesh.num_primitives
    cand = (scenes.complete_graph_candidates(n, order) if order <= 2 else scenes.sampled_candidates(n, order, 20000))
    cand = cand * (2 if assume_quads else 1)
    dense = drt.trace_path_candidates(mesh, tx, rx, cand).masked()
    for capacity in (1 << 16, 3):
        got = drt.trace_valid_path_candidates(mesh, tx, rx, cand, capacity=capacity)
        assert torch.equal(got.vertices, dense.vertices) and torch.equal(got.objects, dense.objects)
    if order in (1, 2):
        assert dense.vertices.shape[0] > 0


def test_sharded_exhaustive_search_world_size_one(drt, two_buildings, kats):
    from differt_b200.distributed import trace_valid_paths_sharded

    v, t = two_buildings
    g = kats["two_buildings_scene"]
    mesh = drt.Mesh.from_numpy(v, t)
    tx, rx = np.array([g["tx"]], np.float32), np.array([g["rx"], [0.5, 9.0, 1.5]], np.float32)
    for order in (1, 2, 3):
        got = trace_valid_paths_sharded(mesh, tx, rx, order, chunk_size=1000)
        exp = drt.trace_paths(mesh, tx, rx, order).masked()
        assert torch.equal(got.vertices, exp.vertices) and torch.equal(got.objects, exp.objects)
        assert got.num_valid_paths >= 1


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fuzz_degenerate_meshes_and_candidates(drt, seed):
    """Zero-area and duplicated triangles, shared vertices, candidates that repeat a triangle or use a
    degenerate one, tx / rx lying on mesh vertices and in triangle planes: every output of the fused
    trace (dense and pruned), any-hit and first-hit still equals the oracle bit for bit."""
    rng = np.random.default_rng(seed)
    nv, nt = 40, 90
    v = rng.integers(-4, 5, size=(nv, 3)).astype(np.float32)  # lattice vertices: many coincidences
    t = rng.integers(0, nv, size=(nt, 3)).astype(np.int32)
    t[:8, 1] = t[:8, 0]                      # zero-area (two identical vertices)
    t[8:12] = t[12:16]                       # duplicated triangles
    v[5] = v[6]                              # coincident vertices
    tx = np.concatenate([v[[0, 3]], rng.uniform(-5, 5, (2, 3)).astype(np.float32)])
    rx = np.concatenate([v[[7]], rng.integers(-5, 6, (6, 3)).astype(np.float32)])
    tri = orc.triangle_vertices(v, t)
    for order in (1, 2, 3):
        cand = rng.integers(0, nt, size=(1500, order)).astype(np.int32)  # repeats allowed on purpose
        mesh = drt.Mesh.from_numpy(v, t)
        with np.errstate(all="ignore"):
            ev, eo, em = co.trace_path_candidates(v, t, tx, rx, cand)
        for dense in (True, False):
            got = drt.trace_path_candidates(mesh, tx, rx, cand, dense_blockage=dense)
            np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
            np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
            np.testing.assert_array_equal(got.objects.cpu().numpy(), eo)
        comp = drt.trace_valid_path_candidates(mesh, tx, rx, cand)
        idx = np.flatnonzero(em.reshape(-1))
        np.testing.assert_array_equal(comp.index.cpu().numpy(), idx)
        np.testing.assert_array_equal(bits(comp.vertices.cpu().numpy()), bits(ev.reshape(-1, order + 2, 3)[idx]))
    o = np.concatenate([v[rng.integers(0, nv, 2000)], rng.integers(-5, 6, (2000, 3)).astype(np.float32)])
    d = (np.concatenate([v[rng.integers(0, nv, 2000)], rng.integers(-5, 6, (2000, 3)).astype(np.float32)]) - o).astype(np.float32)
    with np.errstate(all="ignore"):
        np.testing.assert_array_equal(drt.ray_intersect_any_triangle(o, d, tri).numpy(),
                                      co.ray_intersect_any_triangle(o, d, tri))
        ei, et = co.first_triangle_hit_by_ray(o, d, tri)
    gi, gt = drt.first_triangle_hit_by_ray(o, d, tri)
    np.testing.assert_array_equal(gi.numpy(), ei)
    np.testing.assert_array_equal(bits(gt.numpy()), bits(et))


# ------------------------------------------------------------------------------------------------
# N4 (forward): smoothed primitives, parity to tolerance (the sigmoid is a transcendental)
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("alpha", [1.0, 10.0, 1e8])
def test_smoothed_primitives_match_oracle(drt, rng, alpha):
    tri = rng.normal(size=(300, 3, 3)).astype(np.float32)
    o = rng.normal(size=(500, 3)).astype(np.float32)
    d = (rng.normal(size=(500, 3)) * 3).astype(np.float32)
    t, hit = drt.ray_intersect_triangle(o[:, None], d[:, None], tri, smoothing_factor=alpha)
    et, eh = orc.ray_intersect_triangle_smooth(o[:, None], d[:, None], tri, smoothing_factor=alpha)
    assert hit.dtype == torch.float32 and tuple(hit.shape) == (500, 300)
    np.testing.assert_array_equal(bits(t.numpy()), bits(et))
    np.testing.assert_allclose(hit.numpy(), eh, rtol=1e-5, atol=1e-6)
    mask = rng.uniform(size=300) < 0.6
    for act in (None, mask):
        got = drt.ray_intersect_any_triangle(o, d, tri, act, smoothing_factor=alpha)
        exp = orc.ray_intersect_any_triangle_smooth(o, d, tri, act, smoothing_factor=alpha)
        np.testing.assert_allclose(got.numpy(), exp, rtol=2e-5, atol=1e-6)
    v = rng.normal(size=(64, 5, 3)).astype(np.float32)
    mv, mn = rng.normal(size=(64, 3, 3)).astype(np.float32), rng.normal(size=(64, 3, 3)).astype(np.float32)
    got = drt.consecutive_vertices_are_on_same_side_of_mirror(v, mv, mn, smoothing_factor=alpha)
    exp = orc.consecutive_vertices_are_on_same_side_of_mirror_smooth(v, mv, mn, alpha)
    np.testing.assert_allclose(got.numpy(), exp, rtol=1e-5, atol=1e-6)
    if alpha == 1e8:  # the reference's own test: a huge slope reproduces the hard decisions
        _, hard = drt.ray_intersect_triangle(o[:, None], d[:, None], tri)
        np.testing.assert_array_equal((hit > 0.5).numpy(), hard.numpy())
        np.testing.assert_array_equal(
            (drt.ray_intersect_any_triangle(o, d, tri, smoothing_factor=alpha) > 0.5).numpy(),
            drt.ray_intersect_any_triangle(o, d, tri).numpy())


def _smooth_case(two_buildings, kats, order, quads, masked, rng):
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx = np.stack([np.array(k["tx"], np.float32).reshape(3), np.array([2.0, -6.0, 9.0], np.float32)])
    rx = np.stack([np.array(k["rx"], np.float32).reshape(3), np.array([20.0, 3.0, 1.5], np.float32),
                   np.array([-4.0, 11.0, 4.0], np.float32)])
    T = t.shape[0]
    if order == 0:
        cand = np.empty((1, 0), np.int32)
    else:
        cand = rng.integers(0, T, size=(150, order)).astype(np.int32)
        if quads:
            cand -= cand % 2
    mask = (rng.uniform(size=T) > 0.25) if masked else None
    if quads and mask is not None:
        mask[1::2] = mask[::2]
    return v, t, tx, rx, cand, mask


@pytest.mark.parametrize("alpha", [0.05, 3.0, 1000.0])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("quads,masked", [(False, False), (True, False), (False, True), (True, True)])
def test_smoothed_trace_matches_oracle(drt, two_buildings, kats, rng, alpha, order, quads, masked):
    # relaxed _trace_path_candidates (_solvers.py:599-713) on the reference's 24-triangle scene, with
    # masks, quads and the NaNs of non-finite paths (most confidences are ~0 here: the relaxed blockage
    # sum saturates; test_smoothed_trace_ground_and_wall covers the spread-out regime)
    v, t, tx, rx, cand, mask = _smooth_case(two_buildings, kats, order, quads, masked, rng)
    mesh = drt.Mesh.from_numpy(v, t, mask, assume_quads=quads)
    got = drt.trace_path_candidates(mesh, tx, rx, cand, smoothing_factor=alpha)
    ev, eo, em = orc.trace_path_candidates(v, t, tx, rx, cand, mask=mask, assume_quads=quads, smoothing_factor=alpha)
    assert got.mask.dtype == torch.float32 and tuple(got.mask.shape) == em.shape
    np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
    np.testing.assert_array_equal(got.objects.cpu().numpy(), eo)
    gm = got.mask.cpu().numpy()
    np.testing.assert_array_equal(np.isnan(gm), np.isnan(em))
    np.testing.assert_allclose(gm, em, rtol=5e-5, atol=2e-6)
    # the hard trace returns the same dense vertices / objects
    hard = drt.trace_path_candidates(mesh, tx, rx, cand)
    np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(hard.vertices.cpu().numpy()))
    np.testing.assert_array_equal(got.objects.cpu().numpy(), hard.objects.cpu().numpy())


@pytest.mark.parametrize("alpha", [0.5, 4.0, 40.0])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("quads", [False, True])
def test_smoothed_trace_ground_and_wall(drt, alpha, order, quads):
    # 4 triangles (a ground quad and a wall quad): few enough that the relaxed blockage sum stays
    # below its clip, so the confidences spread over (0, 1) and every term of the min matters
    import itertools

    v = np.array([[-10, -10, 0], [10, -10, 0], [10, 10, 0], [-10, 10, 0],
                  [3, -4, 0], [3, 4, 0], [3, 4, 6], [3, -4, 6]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)
    r = np.random.default_rng(5)
    tx = r.uniform([-8, -8, 1], [1, 8, 8], size=(3, 3)).astype(np.float32)
    rx = r.uniform([-8, -8, 1], [9, 8, 8], size=(40, 3)).astype(np.float32)
    prim = range(0, 4, 2 if quads else 1)
    cand = (np.array(list(itertools.product(prim, repeat=order)), np.int32).reshape(-1, order)
            if order else np.empty((1, 0), np.int32))
    for mask in (None, np.array([True, True, False, False])):
        mesh = drt.Mesh.from_numpy(v, t, mask, assume_quads=quads)
        got = drt.trace_path_candidates(mesh, tx, rx, cand, smoothing_factor=alpha)
        ev, eo, em = orc.trace_path_candidates(v, t, tx, rx, cand, mask=mask, assume_quads=quads,
                                               smoothing_factor=alpha)
        gm = got.mask.cpu().numpy()
        np.testing.assert_array_equal(bits(got.vertices.cpu().numpy()), bits(ev))
        np.testing.assert_array_equal(got.objects.cpu().numpy(), eo)
        np.testing.assert_array_equal(np.isnan(gm), np.isnan(em))
        np.testing.assert_allclose(gm, em, rtol=5e-5, atol=2e-6)
        assert got.num_valid_paths == int((em >= 0.5).sum()) or np.any(np.abs(em - 0.5) < 1e-4)
        if mask is None:
            assert np.unique(em[np.isfinite(em)]).size > 15  # a non-trivial relaxation


@pytest.mark.parametrize("order", [0, 1, 2])
def test_smoothed_trace_huge_slope_is_the_hard_trace(drt, two_buildings, kats, order):
    # in the spirit of the reference's check (test_scene.py:383-440, test_utils.py:642): a huge slope
    # reproduces the hard decisions
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    mesh = drt.Mesh.from_numpy(v, t)
    hard = drt.trace_paths(mesh, tx, rx, order)
    for alpha in (1e8, 1e12):
        soft = drt.trace_paths(mesh, tx, rx, order, smoothing_factor=alpha)
        m = soft.mask.cpu().numpy()
        np.testing.assert_array_equal(m >= 0.5, hard.mask.cpu().numpy())
        # NaN (paths through parallel mirrors: non-finite vertices) counts as invalid, as in the reference
        np.testing.assert_allclose(np.nan_to_num(m, nan=0.0), hard.mask.cpu().numpy().astype(np.float32), atol=1e-6)
        assert soft.num_valid_paths == hard.num_valid_paths
        a, b = soft.masked(), hard.masked()
        np.testing.assert_array_equal(a.objects.cpu().numpy(), b.objects.cpu().numpy())
        np.testing.assert_array_equal(bits(a.vertices.cpu().numpy()), bits(b.vertices.cpu().numpy()))
    # Scene front end, chunked and not (ExhaustivePathTracer(chunk_size, smoothing_factor))
    scene = drt.Scene(mesh=mesh, transmitters=torch.from_numpy(tx).cuda(), receivers=torch.from_numpy(rx).cuda())
    whole = scene.trace_paths(order, smoothing_factor=1e8)
    assert whole.mask.dtype == torch.float32 and whole.num_valid_paths == hard.num_valid_paths
    if order > 0:
        n = sum(p.num_valid_paths for p in scene.trace_paths(order, chunk_size=100, smoothing_factor=1e8))
        assert n == hard.num_valid_paths


def test_smoothed_trace_edge_cases(drt, two_buildings, rng):
    v, t = two_buildings
    mesh = drt.Mesh.from_numpy(v, t)
    tx, rx = np.array([[0.0, 0.0, 30.0]], np.float32), np.array([[5.0, 5.0, 1.0], [np.inf, 0.0, 1.0]], np.float32)
    # zero candidates, zero receivers
    p = drt.trace_path_candidates(mesh, tx, rx, np.empty((0, 2), np.int32), smoothing_factor=2.0)
    assert tuple(p.mask.shape) == (1, 2, 0) and p.mask.dtype == torch.float32
    p = drt.trace_path_candidates(mesh, tx, rx[:0], np.zeros((3, 1), np.int32), smoothing_factor=2.0)
    assert tuple(p.vertices.shape) == (1, 0, 3, 3, 3)
    # a non-finite receiver: vertices zeroed like the hard trace, never valid (0 or NaN as in the reference)
    cand = rng.integers(0, t.shape[0], size=(20, 2)).astype(np.int32)
    p = drt.trace_path_candidates(mesh, tx, rx, cand, smoothing_factor=50.0)
    assert not (p.mask[0, 1] >= 0.5).any() and (p.vertices[0, 1] == 0).all()
    ev, eo, em = orc.trace_path_candidates(v, t, tx, rx, cand, smoothing_factor=50.0)
    np.testing.assert_array_equal(np.isnan(p.mask.cpu().numpy()), np.isnan(em))
    np.testing.assert_allclose(p.mask.cpu().numpy(), em, rtol=5e-5, atol=2e-6)
    # empty mesh at order 0: nothing blocks; line of sight only limited by the too-small term
    empty = drt.Mesh.from_numpy(np.empty((0, 3), np.float32), np.empty((0, 3), np.int32))
    p = drt.trace_paths(empty, tx, rx[:1], 0, smoothing_factor=10.0)
    e = orc.trace_path_candidates(np.empty((0, 3), np.float32), np.empty((0, 3), np.int32), tx, rx[:1],
                                  np.empty((1, 0), np.int32), smoothing_factor=10.0)[2]
    np.testing.assert_allclose(p.mask.cpu().numpy(), e, rtol=1e-5)


@pytest.mark.parametrize("alpha", [0.5, 4.0, 40.0])
def test_smoothed_primitives_gradients_vs_autograd_oracle(drt, alpha):
    # jax.grad of the relaxed primitives (_utils.py:1263-1322, 1452-1476): float64 torch oracle
    from oracle import smooth_grad_oracle as sg

    r = np.random.default_rng(21)
    tri = r.normal(size=(7, 3, 3)).astype(np.float32)
    o = r.normal(size=(60, 3)).astype(np.float32)
    d = (r.normal(size=(60, 3)) * 2).astype(np.float32)
    wt, wh = r.normal(size=(60, 7)).astype(np.float32), r.normal(size=(60, 7)).astype(np.float32)

    def leaves(dtype, dev):
        return [torch.tensor(x, dtype=dtype, device=dev, requires_grad=True) for x in (o, d, tri)]

    # elementwise, with broadcasting on both sides ([60,1] rays x [7] triangles): t and hit carry gradients
    co, cd, ct = leaves(torch.float32, "cuda")
    t, hit = drt.ray_intersect_triangle(co[:, None], cd[:, None], ct, smoothing_factor=alpha)
    ((t * torch.from_numpy(wt).cuda()).sum() + (hit * torch.from_numpy(wh).cuda()).sum()).backward()
    eo, ed, et = leaves(torch.float64, "cpu")
    t64, hit64 = sg.ray_intersect_triangle_smooth(eo[:, None], ed[:, None], et, smoothing_factor=alpha)
    np.testing.assert_allclose(hit.detach().cpu().numpy(), hit64.detach().numpy(), rtol=1e-4, atol=1e-5)
    ((t64 * torch.from_numpy(wt).double()).sum() + (hit64 * torch.from_numpy(wh).double()).sum()).backward()
    for name, got, exp in (("o", co.grad, eo.grad), ("d", cd.grad, ed.grad), ("tri", ct.grad, et.grad)):
        e = exp.numpy()
        np.testing.assert_allclose(got.cpu().numpy(), e, rtol=2e-3, atol=2e-3 * np.abs(e).max(), err_msg=name)

    # any-hit: few triangles so that some sums stay below the clip, with and without a mask
    w = r.normal(size=60).astype(np.float32)
    for act in (None, np.array([True, False, True, True, False, True, True])):
        co, cd, ct = leaves(torch.float32, "cuda")
        got = drt.ray_intersect_any_triangle(co, cd, ct, act, smoothing_factor=alpha)
        (got * torch.from_numpy(w).cuda()).sum().backward()
        eo, ed, et = leaves(torch.float64, "cpu")
        exp = sg.ray_intersect_any_triangle_smooth(eo, ed, et, act, smoothing_factor=alpha)
        np.testing.assert_allclose(got.detach().cpu().numpy(), exp.detach().numpy(), rtol=1e-4, atol=1e-5)
        assert (exp < 1).any()
        (exp * torch.from_numpy(w).double()).sum().backward()
        for name, g, e in (("o", co.grad, eo.grad), ("d", cd.grad, ed.grad), ("tri", ct.grad, et.grad)):
            e = e.numpy()
            np.testing.assert_allclose(g.cpu().numpy(), e, rtol=2e-3, atol=2e-3 * max(np.abs(e).max(), 1e-6), err_msg=name)
        if act is not None:
            assert (ct.grad[~torch.from_numpy(act).cuda()] == 0).all()  # masked-out triangles get no gradient
    # the same-side relaxation is a function of signs: no gradient, as in the reference
    v = torch.from_numpy(r.normal(size=(5, 4, 3)).astype(np.float32)).cuda().requires_grad_(True)
    mv = torch.from_numpy(r.normal(size=(5, 2, 3)).astype(np.float32)).cuda()
    assert not drt.consecutive_vertices_are_on_same_side_of_mirror(v, mv, mv, smoothing_factor=alpha).requires_grad


def test_reduce_and_gradient_ascent_on_relaxed_power(drt):
    # TracedPaths.reduce (_paths.py:461-479) + path_length (_utils.py:150-182): the consumer the relaxed
    # trace exists for.  J(tx) = sum over order-0/1 paths of confidence / length^2 is differentiable in
    # the transmitter position; check d J against a finite difference and climb it.
    v, t, _, rx, _ = _ground_and_wall(1, False)
    mesh = drt.Mesh.from_numpy(v, t)
    rxc = torch.from_numpy(rx).cuda()

    def power(tx, alpha=4.0):
        total = 0.0
        for order in (0, 1):
            paths = drt.trace_paths(mesh, tx, rxc, order, smoothing_factor=alpha)
            total = total + paths.reduce(lambda p: 1.0 / drt.path_length(p) ** 2)
        return total

    tx = torch.tensor([[-6.0, 1.0, 2.0]], device="cuda", requires_grad=True)
    j0 = power(tx)
    j0.backward()
    g = tx.grad.clone()
    assert torch.isfinite(g).all() and g.abs().max() > 0
    step = 1e-2 * g / g.norm()
    with torch.no_grad():
        fd = (power(tx + step) - power(tx - step)) / 2e-2
    assert abs(fd.item() - g.norm().item()) <= 0.05 * g.norm().item()
    cur = tx.detach().clone()
    for _ in range(25):
        cur.requires_grad_(True)
        j = power(cur)
        (gc,) = torch.autograd.grad(j, cur)
        cur = (cur + 0.3 * gc / gc.norm()).detach()
    assert power(cur).item() > 1.2 * j0.item()

    # boolean mask: reduce adds the valid paths only, whatever fun returns on the others
    hard = drt.trace_paths(mesh, tx.detach(), rxc, 1)
    lengths = drt.path_length(hard.vertices)
    exp = lengths[hard.mask].sum()
    got = hard.reduce(lambda p: torch.where(hard.mask, drt.path_length(p), torch.full_like(lengths, float("nan"))))
    torch.testing.assert_close(got, exp)
    per_rx = hard.reduce(drt.path_length, axis=-1)
    assert tuple(per_rx.shape) == (1, rx.shape[0])
    torch.testing.assert_close(per_rx.sum(), exp)
    np.testing.assert_allclose(drt.path_length(np.array([[1.0, 0, 0], [1, 1, 0], [1, 0, 0]], np.float32)), 2.0)


def _ground_and_wall(order, quads):
    import itertools

    # not axis-aligned on purpose: every component of every cotangent is exercised
    v = np.array([[-10, -10, 0.3], [10, -10, -0.2], [10, 10, 0.4], [-10, 10, 0.1],
                  [3, -4, 0], [3.5, 4, 0], [3.2, 4.3, 6], [2.8, -4, 6.2]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)
    r = np.random.default_rng(5)
    tx = r.uniform([-8, -8, 1], [1, 8, 8], size=(3, 3)).astype(np.float32)
    rx = r.uniform([-8, -8, 1], [9, 8, 8], size=(20, 3)).astype(np.float32)
    if order == 0:
        cand = np.empty((1, 0), np.int32)
    else:  # consecutive mirrors from different quads: no (near-)coplanar double reflections, which are
        # ill-conditioned in fp32 (the fp64 gradient oracle would take other branches)
        prim = range(0, 4, 2 if quads else 1)
        cand = np.array([c for c in itertools.product(prim, repeat=order)
                         if all(c[i] // 2 != c[i + 1] // 2 for i in range(order - 1))], np.int32).reshape(-1, order)
    return v, t, tx, rx, cand


@pytest.mark.parametrize("alpha", [0.5, 4.0, 40.0])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("quads", [False, True])
def test_smoothed_trace_gradient_vs_autograd_oracle(drt, alpha, order, quads):
    # what jax.grad gives on the relaxed branch (_solvers.py:576-713): cotangents on the confidences AND on
    # the path vertices, to mesh.vertices / tx / rx; the oracle is the float64 torch restatement
    from oracle import smooth_grad_oracle as sg

    v, t, tx, rx, cand = _ground_and_wall(order, quads)
    r = np.random.default_rng(11)
    w = r.normal(size=(tx.shape[0], rx.shape[0], cand.shape[0])).astype(np.float32)
    # (order 3 has far, ill-conditioned image points: fp32 vs fp64 differ by 1e-3 there; K6b's own tests cover them)
    gw = ((0.01 if order < 3 else 0.0) * r.normal(size=(*w.shape, order + 2, 3))).astype(np.float32)
    for mask in ((None, np.array([True, True, True, False])) if not quads else (None,)):
        mesh = drt.Mesh.from_numpy(v, t, mask, assume_quads=quads)
        mesh.vertices.requires_grad_(True)
        txc = torch.from_numpy(tx).cuda().requires_grad_(True)
        rxc = torch.from_numpy(rx).cuda().requires_grad_(True)
        paths = drt.trace_path_candidates(mesh, txc, rxc, cand, smoothing_factor=alpha)
        assert not torch.isnan(paths.mask).any()
        ((paths.mask * torch.from_numpy(w).cuda()).sum() + (paths.vertices * torch.from_numpy(gw).cuda()).sum()).backward()

        V = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        TX = torch.tensor(tx, dtype=torch.float64, requires_grad=True)
        RX = torch.tensor(rx, dtype=torch.float64, requires_grad=True)
        full, conf, idx = sg.relaxed_trace(V, t, TX, RX, cand, mask=mask, assume_quads=quads, smoothing_factor=alpha)
        assert idx.numel() == w.size  # every path of this scene is finite
        np.testing.assert_allclose(paths.mask.detach().cpu().numpy().reshape(-1), conf.detach().numpy(), rtol=1e-4, atol=1e-5)
        ((conf * torch.from_numpy(w.reshape(-1)).double()).sum()
         + (full * torch.from_numpy(gw.reshape(-1, order + 2, 3)).double()).sum()).backward()
        for name, got, exp in (("tx", txc.grad, TX.grad), ("rx", rxc.grad, RX.grad), ("vertices", mesh.vertices.grad, V.grad)):
            e = exp.numpy()
            np.testing.assert_allclose(got.cpu().numpy(), e, rtol=2e-3, atol=2e-3 * max(np.abs(e).max(), 1e-6),
                                       err_msg=f"{name} alpha={alpha} order={order} quads={quads} mask={mask is not None}")
        # a cotangent on the vertices alone gives the hard trace's gradient (same image-method VJP)
        if mask is None and order == 2:
            def grads(**kw):
                m = drt.Mesh(mesh.vertices.detach().clone().requires_grad_(True), mesh.triangles, assume_quads=quads)
                a = txc.detach().clone().requires_grad_(True)
                b = rxc.detach().clone().requires_grad_(True)
                (drt.trace_path_candidates(m, a, b, cand, **kw).vertices * torch.from_numpy(gw).cuda()).sum().backward()
                return [x.grad.cpu().numpy() for x in (m.vertices, a, b)]

            for got, exp in zip(grads(smoothing_factor=alpha), grads()):
                np.testing.assert_allclose(got, exp, rtol=1e-4, atol=1e-6)


def test_compute_tx_mlm_reference_properties(drt):
    """The reference's own MLM test restated (differt/tests/geometry/test_scene.py:761-875: same box,
    same transmitters, same grids, same assertions), and its masked-mesh test (:877-917) on a scene
    whose every triangle but the ground is masked (the reference masks all of simple_street_canyon but
    its two ground triangles; the assertion — order-0 cells all carry the empty path's hash
    2166136261, no path of two or more bounces exists — does not depend on the buildings)."""
    vertices = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                         [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32)
    triangles = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4],
                          [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6], [3, 0, 4], [3, 4, 7]], np.int32)
    mesh = drt.Mesh.from_numpy(vertices, triangles)
    scene = drt.Scene(np.array([0.0, 0.0, 5.0], np.float32), np.empty((0, 3), np.float32), mesh)
    scene = scene.with_receivers_grid(m=10, n=10, height=1.5)
    for kwargs in ({}, {"height": 2.0}, {"min_order": 1}):
        mlm = scene.compute_tx_mlm(max_order=1, dim_x=10, dim_y=10, num_rays=500, **kwargs)
        assert tuple(mlm.shape) == (10, 10)
        assert int(mlm.min()) >= 0 and int(mlm.max()) < (1 << 32)  # uint32 values
    assert torch.all(scene.compute_tx_mlm(max_order=1, min_order=2, dim_x=10, dim_y=10, num_rays=500) == 0)
    assert tuple(scene.set_assume_quads(True).compute_tx_mlm(max_order=1, dim_x=10, dim_y=10, num_rays=500).shape) == (10, 10)
    above = scene.compute_tx_mlm(max_order=0, min_order=0, dim_x=5, dim_y=5, num_rays=50000, height=1.5)
    assert torch.any(above > 0) and above[0, 0] > 0 and above[4, 4] > 0
    inside = scene.compute_tx_mlm(max_order=2, dim_x=5, dim_y=5, num_rays=1000, height=0.0)
    assert torch.all(inside == 0)
    below = scene.compute_tx_mlm(max_order=0, min_order=0, dim_x=5, dim_y=5, num_rays=1000, height=-2.0)
    assert torch.all(below == 0)
    multi = drt.Scene(np.array([[0.0, 0.0, 5.0], [0.0, 0.0, 6.0]], np.float32), np.empty((0, 3), np.float32), mesh)
    assert tuple(multi.compute_tx_mlm(max_order=1, dim_x=10, dim_y=10, num_rays=500).shape) == (2, 10, 10)
    tx2d = np.tile(np.array([[0.0, 0.0, 5.0], [0.0, 0.0, 6.0], [0.0, 0.0, 7.0]], np.float32), (2, 1, 1))
    multi = drt.Scene(tx2d, np.empty((0, 3), np.float32), mesh)
    assert tuple(multi.compute_tx_mlm(max_order=1, dim_x=10, dim_y=10, num_rays=500).shape) == (2, 3, 10, 10)

    v, t = scenes.street_canyon(3)
    ground = np.zeros(t.shape[0], bool)
    ground[-2:] = True  # the two ground triangles come last (scenes.street_canyon), like the reference scene's 72:74
    assert np.allclose(v[t[-2:]][..., 2], 0.0)
    canyon = drt.Scene(np.array([-33.0, 0.0, 32.0], np.float32), np.empty((0, 3), np.float32),
                       drt.Mesh.from_numpy(v, t, mask=ground))
    mlm_0 = canyon.compute_tx_mlm(max_order=0, min_order=0, dim_x=5, dim_y=5, num_rays=50000, height=1.5)
    assert torch.all(mlm_0 == 2166136261)
    assert torch.all(canyon.compute_tx_mlm(max_order=2, min_order=2, dim_x=5, dim_y=5, num_rays=50000, height=1.5) == 0)


@pytest.mark.parametrize("tilt_some", [False, True])
def test_cull_axis_aligned_planes_and_in_plane_segments(drt, tilt_some, monkeypatch):
    """Exactly axis-aligned rectangles on a few shared planes (the floors, the ground and the rows of walls
    of a city model) and segments lying exactly IN those planes (d_j == 0: what two consecutive reflections
    on one plane produce): cull.cuh proves that an aligned triangle cannot be hit by such a segment and
    drops the axis from the grazing guard.  Masks must equal the oracle's, with far fewer tests than the
    whole mesh per segment; `tilt_some` rotates a third of the rectangles by a fraction of a degree so that
    aligned and general nodes mix (and in-plane segments graze the tilted ones for real)."""
    rng = np.random.default_rng(11)
    quads = []
    for axis in range(3):
        for plane in rng.integers(-20, 20, 6) * 16.0:
            for _ in range(60):
                c = rng.integers(-300, 300, 3).astype(np.float64)
                c[axis] = plane
                a, b = np.zeros(3), np.zeros(3)
                a[(axis + 1) % 3], b[(axis + 2) % 3] = rng.integers(2, 30), rng.integers(2, 30)
                quads.append(np.stack((c, c + a, c + a + b, c + b)))
    quads = np.array(quads)
    if tilt_some:
        ang = 0.003
        rot = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        quads[::3] = quads[::3] @ rot.T
    v = quads.reshape(-1, 3).astype(np.float32)
    base = 4 * np.arange(quads.shape[0])[:, None]
    t = np.concatenate((base + [0, 1, 2], base + [0, 2, 3]), axis=1).reshape(-1, 3).astype(np.int32)
    mesh = drt.Mesh.from_numpy(v, t)
    # order 0: transmitters / receivers exactly on the shared planes (and a few generic ones)
    tx, rx = [], []
    for axis in range(3):
        planes = np.unique(quads[:, 0, axis])[:4]
        for pts, count in ((tx, 4), (rx, 192)):
            p = rng.uniform(-320, 320, (len(planes), count, 3))
            p[..., axis] = planes[:, None]
            pts.append(p.reshape(-1, 3))
    tx.append(rng.uniform(-320, 320, (4, 3))), rx.append(rng.uniform(-320, 320, (64, 3)))
    tx, rx = np.concatenate(tx).astype(np.float32), np.concatenate(rx).astype(np.float32)
    cand0 = np.zeros((1, 0), np.int32)
    # orders 2 and 3: consecutive mirrors on one plane (the two triangles of a rectangle, two rectangles of a plane)
    tri = rng.integers(0, t.shape[0] // 2, 400) * 2
    same_plane = np.array([rng.choice(np.nonzero((quads[:, 0, a] == quads[q // 2, 0, a]))[0]) * 2
                           for q in tri for a in [int(np.argmax(np.ptp(quads[q // 2], axis=0) == 0))]])
    cand2 = np.stack((tri, tri + 1), -1).astype(np.int32)
    cand3 = np.stack((tri, same_plane + 1, rng.integers(0, t.shape[0], 400)), -1).astype(np.int32)
    cand3 = cand3[cand3[:, 0] != cand3[:, 1]]
    dense_tests = 0
    for cand in (cand0, cand2, cand3):
        # reflected paths: a transmitter / receiver subset that still has points on every kind of plane
        tt_, r = (tx[::4], rx[::37]) if cand.shape[1] else (tx, rx)
        _, _, em = co.trace_path_candidates(v, t, tt_, r, cand)
        for dense in (True, False):
            got = drt.trace_path_candidates(mesh, tt_, r, cand, dense_blockage=dense, with_stats=True)
            np.testing.assert_array_equal(got.mask.cpu().numpy(), em)
            if dense and cand.shape[1] == 0:
                dense_tests = got.stats["tests_done"]
    if not tilt_some:  # in-plane segments no longer walk every triangle of their plane's axis
        assert dense_tests < 0.2 * tx.shape[0] * rx.shape[0] * t.shape[0]
    # the flat queries behind the same cull: rays with exactly zero direction components from points on the planes
    o = np.repeat(rx[:256], 8, axis=0)
    d = rng.uniform(-400, 400, o.shape).astype(np.float32)
    d[np.arange(d.shape[0]), rng.integers(0, 3, d.shape[0])] = 0.0
    from differt_b200 import geometry

    monkeypatch.setattr(geometry, "_CULL_MIN_WORK", 0)  # the public API takes the culled path at any size
    tvv = orc.triangle_vertices(v, t)
    np.testing.assert_array_equal(drt.ray_intersect_any_triangle(o, d, tvv).numpy(), co.ray_intersect_any_triangle(o, d, tvv))
    idx, tt = drt.first_triangle_hit_by_ray(o, d, tvv)
    i0, t0 = co.first_triangle_hit_by_ray(o, d, tvv)
    np.testing.assert_array_equal(bits(tt.numpy()), bits(t0))
    assert np.isfinite(t0).any()
