"""CPU checks of the relaxed (smoothing_factor) restatement of the trace step
(reference ``_solvers.py:599-713``): it has no golden vector in the reference's tests beyond "a huge
slope reproduces the hard decisions" (``test_utils.py:642, 710``; ``test_image_method.py:252``;
``test_scene.py:383-440``), so that is what pins it, together with structural properties."""

from __future__ import annotations

import itertools

import numpy as np
import pytest

from oracle import differt_oracle as orc


def _all_candidates(T, order, step=1):
    if order == 0:
        return np.empty((1, 0), np.int32)
    return np.array(list(itertools.product(range(0, T, step), repeat=order)), np.int32).reshape(-1, order)


@pytest.mark.parametrize("order", [0, 1, 2])
def test_huge_slope_reproduces_hard_trace(two_buildings, kats, order):
    v, t = two_buildings
    k = kats["two_buildings_scene"]
    tx, rx = np.array(k["tx"], np.float32), np.array(k["rx"], np.float32)
    cand = _all_candidates(t.shape[0], order)
    hv, ho, hm = orc.trace_path_candidates(v, t, tx, rx, cand)
    sv, so, sm = orc.trace_path_candidates(v, t, tx, rx, cand, smoothing_factor=1e8)
    assert sm.dtype == np.float32 and sm.shape == hm.shape
    np.testing.assert_array_equal(sm >= 0.5, hm)
    np.testing.assert_array_equal(sv.view(np.uint32), hv.view(np.uint32))
    np.testing.assert_array_equal(so, ho)


@pytest.mark.parametrize("quads", [False, True])
def test_relaxed_mask_properties(quads):
    v = np.array([[-10, -10, 0], [10, -10, 0], [10, 10, 0], [-10, 10, 0],
                  [3, -4, 0], [3, 4, 0], [3, 4, 6], [3, -4, 6]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)
    r = np.random.default_rng(5)
    tx = r.uniform([-8, -8, 1], [1, 8, 8], size=(2, 3)).astype(np.float32)
    rx = r.uniform([-8, -8, 1], [9, 8, 8], size=(10, 3)).astype(np.float32)
    cand = _all_candidates(4, 2, 2 if quads else 1)
    _, _, m = orc.trace_path_candidates(v, t, tx, rx, cand, assume_quads=quads, smoothing_factor=4.0)
    assert m.shape == (2, 10, cand.shape[0]) and np.all((m >= 0) & (m <= 1))
    assert np.unique(m).size > 10
    # masking out the wall: candidates through it get confidence exactly 0, the others can only gain
    mask = np.array([True, True, False, False])
    _, _, mm = orc.trace_path_candidates(v, t, tx, rx, cand, mask=mask, assume_quads=quads, smoothing_factor=4.0)
    uses_wall = (cand >= 2).any(axis=1)
    assert np.all(mm[..., uses_wall] == 0)
    assert np.all(mm[..., ~uses_wall] >= m[..., ~uses_wall] - 1e-7)
    # a masked mesh equals the sub-mesh for candidates that avoid the masked triangles
    _, _, ms = orc.trace_path_candidates(v[:4], t[:2], tx, rx, cand[~uses_wall], assume_quads=quads,
                                         smoothing_factor=4.0)
    np.testing.assert_allclose(mm[..., ~uses_wall], ms, rtol=1e-6)


# ------------------------------------------------------------------------------------------------
# the float64 torch restatement used as the gradient oracle of the relaxed trace
# ------------------------------------------------------------------------------------------------


def _scene():
    v = np.array([[-10, -10, 0.3], [10, -10, -0.2], [10, 10, 0.4], [-10, 10, 0.1],
                  [3, -4, 0], [3.5, 4, 0], [3.2, 4.3, 6], [2.8, -4, 6.2]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]], np.int32)
    r = np.random.default_rng(5)
    tx = r.uniform([-8, -8, 1], [1, 8, 8], size=(2, 3)).astype(np.float32)
    rx = r.uniform([-8, -8, 1], [9, 8, 8], size=(6, 3)).astype(np.float32)
    return v, t, tx, rx


@pytest.mark.parametrize("order,quads", [(0, False), (1, False), (2, False), (2, True), (3, False)])
def test_gradient_oracle_forward_matches_numpy_oracle(order, quads):
    import torch

    from oracle import smooth_grad_oracle as sg

    v, t, tx, rx = _scene()
    prim = range(0, 4, 2 if quads else 1)
    cand = (np.array([c for c in itertools.product(prim, repeat=order)
                      if all(c[i] // 2 != c[i + 1] // 2 for i in range(order - 1))], np.int32).reshape(-1, order)
            if order else np.empty((1, 0), np.int32))
    for alpha in (0.5, 4.0, 40.0):
        for mask in (None, np.array([True, True, True, False]) if not quads else None):
            ev, _, em = orc.trace_path_candidates(v, t, tx, rx, cand, mask=mask, assume_quads=quads, smoothing_factor=alpha)
            full, conf, idx = sg.relaxed_trace(torch.tensor(v, dtype=torch.float64), t, torch.tensor(tx, dtype=torch.float64),
                                               torch.tensor(rx, dtype=torch.float64), cand, mask=mask, assume_quads=quads,
                                               smoothing_factor=alpha)
            assert idx.numel() == em.size
            np.testing.assert_allclose(conf.numpy(), em.reshape(-1), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(full.numpy(), ev.reshape(-1, order + 2, 3), rtol=2e-3, atol=1e-3)  # far, ill-conditioned points


def test_gradient_oracle_agrees_with_finite_differences():
    import torch

    from oracle import smooth_grad_oracle as sg

    v, t, tx, rx = _scene()
    cand = np.array([[0, 2], [3, 1], [2, 0]], np.int32)
    w = torch.tensor(np.random.default_rng(3).normal(size=2 * 6 * 3))

    def loss(V, TX, RX):
        return (sg.relaxed_trace(V, t, TX, RX, cand, smoothing_factor=4.0)[1] * w).sum()

    args = [torch.tensor(x, dtype=torch.float64, requires_grad=True) for x in (v, tx, rx)]
    loss(*args).backward()
    r = np.random.default_rng(4)
    for k, a in enumerate(args):  # directional derivatives, central differences
        d = torch.tensor(r.normal(size=tuple(a.shape)))
        h = 1e-6
        plus = [x.detach() + (h * d if i == k else 0) for i, x in enumerate(args)]
        minus = [x.detach() - (h * d if i == k else 0) for i, x in enumerate(args)]
        fd = (loss(*plus) - loss(*minus)).item() / (2 * h)
        an = (a.grad * d).sum().item()
        assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (k, fd, an)


def test_gradient_oracle_primitives_match_numpy_oracle():
    import torch

    from oracle import smooth_grad_oracle as sg

    r = np.random.default_rng(21)
    tri = r.normal(size=(7, 3, 3)).astype(np.float32)
    o = r.normal(size=(60, 3)).astype(np.float32)
    d = (r.normal(size=(60, 3)) * 2).astype(np.float32)
    T64 = lambda x: torch.tensor(x, dtype=torch.float64)  # noqa: E731
    for alpha in (0.5, 4.0, 40.0):
        t, hit = sg.ray_intersect_triangle_smooth(T64(o)[:, None], T64(d)[:, None], T64(tri), smoothing_factor=alpha)
        et, eh = orc.ray_intersect_triangle_smooth(o[:, None], d[:, None], tri, smoothing_factor=alpha)
        np.testing.assert_allclose(hit.numpy(), eh, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(t.numpy(), et, rtol=1e-3, atol=1e-4)
        for act in (None, np.array([True, False, True, True, False, True, True])):
            got = sg.ray_intersect_any_triangle_smooth(T64(o), T64(d), T64(tri), act, smoothing_factor=alpha)
            exp = orc.ray_intersect_any_triangle_smooth(o, d, tri, act, smoothing_factor=alpha)
            np.testing.assert_allclose(got.numpy(), exp, rtol=1e-4, atol=1e-5)


def test_smoothing_function_reference_kats():
    # differt/tests/test_utils.py:58-82: limits of the sigmoid relaxation
    r = np.random.default_rng(0)
    alpha = r.uniform(0, 100, size=(20, 1)).astype(np.float32)
    got = orc.smoothing_function(np.array([-1e8, 0.0, 1e8], np.float32), alpha)
    np.testing.assert_allclose(got, np.broadcast_to(np.array([0.0, 0.5, 1.0], np.float32), got.shape), atol=1e-6)
    x = (r.normal(size=(40, 1, 10)) * 1000.0).astype(np.float32)
    x[0, 0, 0] = 0.0
    np.testing.assert_array_equal(orc.smoothing_function(x, 1e8), 0.5 * (np.sign(x) + 1))
    assert orc.smoothing_function(x, alpha).shape == (40, 20, 10)
