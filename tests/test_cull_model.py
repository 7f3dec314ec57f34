"""The exact cull's formulas and constants (NumPy restatement of csrc/cull.cu / cull.cuh in tests/_cull_model.py)
against the CPU oracle: a group of triangles may be culled for a segment only if the reference's fp32 test reports
NO hit of that segment on any of its triangles.  Searched where culling is hardest — the noise hits of in-plane
segments kilometres away from the triangles, exactly axis-aligned planes with in-plane segments (the case the
aligned flag exists for), rounding-residue segments — and required to be non-vacuous on both sides: the scenes
contain reported hits, and most groups ARE culled.

What this sample discriminates (checked by breaking the model on purpose): without the grazing guard (g > 0) the
noise-hit test fails at once; the margin's constant is NOT discriminated here — the reported hits the guard lets
through lie within a few percent of the proven distance, which tests/test_cull_bound.py measures directly."""

from __future__ import annotations

import numpy as np
import pytest

from _cull_model import Groups
from oracle import c_oracle as co

THR = np.float32(1.0 - 100 * float(np.finfo(np.float32).eps))


def _reported(o, d, tri):
    """bool [T]: the reference reports a blocking hit of segment (o, d) on triangle j."""
    n = tri.shape[0]
    t, hit = co.ray_intersect_triangle(np.broadcast_to(o, (n, 3)), np.broadcast_to(d, (n, 3)), tri)
    return hit & (t < THR)


def _check(groups, tri, segments):
    culled_total = hits_total = 0
    for o, d in segments:
        culled = groups.culled(o, d)
        rep = _reported(o, d, tri)
        bad = rep.reshape(-1, 8).any(-1) & culled
        assert not bad.any(), f"segment {o} + t {d}: group(s) {np.nonzero(bad)[0]} culled although the reference reports a hit"
        culled_total += int(culled.sum())
        hits_total += int(rep.sum())
    return culled_total, hits_total


@pytest.mark.parametrize("lift", [0.0, 1e-3])
def test_noise_hits_of_in_plane_segments_are_never_culled(lift):
    from test_gpu_parity import _coplanar_clusters

    rng = np.random.default_rng(5)
    v, t, planes = _coplanar_clusters(rng)
    tri = v[t]
    groups = Groups(tri)
    segments = []
    for c0, a, b, n, _ in planes:
        uvt, uvr = rng.uniform(-2500, 2500, (6, 2)), rng.uniform(-2500, 2500, (80, 2))
        tx = (c0 + uvt[:, 0:1] * a + uvt[:, 1:2] * b + rng.uniform(-lift, lift, (6, 1)) * n).astype(np.float32)
        rx = (c0 + uvr[:, 0:1] * a + uvr[:, 1:2] * b + rng.uniform(-lift, lift, (80, 1)) * n).astype(np.float32)
        segments += [(p, (q - p).astype(np.float32)) for p in tx for q in rx]
    culled, hits = _check(groups, tri, segments)
    assert hits > (500 if lift == 0.0 else 100)                  # the noise hits are in the sample …
    assert culled > 0.6 * len(segments) * tri.shape[0] // 8      # … and the cull still removes most of the work


@pytest.mark.parametrize("tilt_some", [False, True])
def test_axis_aligned_planes_with_in_plane_segments(tilt_some):
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
    quads = quads.astype(np.float32)
    tri = np.concatenate((quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]), axis=1).reshape(-1, 3, 3)
    # groups of one plane's triangles (what a spatial order produces): the flag needs every triangle aligned
    groups = Groups(tri)
    assert groups.flag.any() and (tilt_some or groups.flag.all())
    segments = []
    for axis in range(3):
        for plane in np.unique(quads[:, 0, axis])[:4]:
            p = rng.uniform(-320, 320, (40, 3)).astype(np.float32)
            q = rng.uniform(-320, 320, (40, 3)).astype(np.float32)
            p[:, axis] = q[:, axis] = plane                     # exactly in the plane: d_j == 0
            segments += [(pi, (qi - pi).astype(np.float32)) for pi, qi in zip(p, q)]
    segments += [(p, (q - p).astype(np.float32)) for p, q in zip(rng.uniform(-320, 320, (100, 3)).astype(np.float32),
                                                               rng.uniform(-320, 320, (100, 3)).astype(np.float32))]
    culled, hits = _check(groups, tri, segments)
    assert hits > 20
    if not tilt_some:  # an in-plane segment no longer makes every node of that plane's axis un-cullable
        in_plane = segments[:480]
        frac = np.mean([groups.culled(o, d).mean() for o, d in in_plane[::8]])
        assert frac > 0.8, frac


def test_rounding_residue_segments():
    """Segments of 2^-30 … 2^-14 on and next to the triangles of a city-sized scene (two path vertices that coincide
    up to rounding; below ~2^-20 the fp32 rounding of the origin alone keeps them off the plane and hits become
    very rare): never culled when the reference reports a hit, culled for most other groups."""
    rng = np.random.default_rng(3)
    tri = []
    for _ in range(40):  # 40 groups, each 8 large triangles of one (random) plane: coherent normals, as a spatial order gives
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        a = np.cross(n, [1.0, 0.0, 0.0] if abs(n[0]) < 0.9 else [0.0, 1.0, 0.0])
        a /= np.linalg.norm(a)
        b = np.cross(n, a)
        c0 = rng.uniform(-400, 400, 3)
        uv = rng.uniform(-60, 60, (8, 3, 2))
        tri.append(c0 + uv[..., 0:1] * a + uv[..., 1:2] * b)
    tri = np.concatenate(tri).astype(np.float32)
    groups = Groups(tri)
    segments = []
    for j in rng.integers(0, tri.shape[0], 600):
        r = rng.dirichlet((1, 1, 1))
        on_tri = (r[:, None] * tri[j].astype(np.float64)).sum(0)
        d = rng.normal(size=3)
        d = (d / np.linalg.norm(d) * 2.0 ** rng.uniform(-30, -14)).astype(np.float32)
        segments.append(((on_tri - 0.5 * d.astype(np.float64)).astype(np.float32), d))
    culled, hits = _check(groups, tri, segments)
    assert hits > 20
    assert culled > 0.8 * len(segments) * 40  # 40 groups: nearly all but the one the segment sits on are culled
