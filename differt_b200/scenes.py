"""Seeded synthetic scenes and path-candidate sets (host side, NumPy only).

The reference's benchmark scenes are either downloaded at test time (``simple_street_canyon``,
``differt/src/differt/geometry/_sionna.py:23-123``) or too small; these generators build the
scenes that SURVEY.md §8(d) and BASELINE.md §4 specify, from the same vertex/triangle layout the
reference's ``Mesh.box`` uses (``differt/src/differt/geometry/_mesh.py:2172-2217``), so that
even/odd triangles form coplanar quads and ``assume_quads`` can be toggled.
"""

from __future__ import annotations

import numpy as np

__all__ = [
    "box",
    "street_canyon",
    "urban_grid",
    "receivers_grid",
    "complete_graph_candidates",
    "num_complete_graph_candidates",
    "sampled_candidates",
]

# Triangle list of Mesh.box: 4 side faces, then bottom, then top (each face = 2 triangles).
_BOX_SIDES = np.array(
    [[0, 1, 2], [0, 2, 3], [3, 2, 4], [3, 4, 5], [5, 4, 6], [5, 6, 7], [7, 6, 1], [7, 1, 0]],
    dtype=np.int32,
)
_BOX_BOTTOM = np.array([[1, 4, 2], [1, 6, 4]], dtype=np.int32)
_BOX_TOP = np.array([[0, 3, 5], [0, 5, 7]], dtype=np.int32)


def box(
    length: float = 1.0,
    width: float = 1.0,
    height: float = 1.0,
    *,
    with_top: bool = False,
    with_bottom: bool = True,
    center=(0.0, 0.0, 0.0),
) -> tuple[np.ndarray, np.ndarray]:
    """Axis-aligned box with the reference's vertex order → ``(vertices [8,3], triangles [T,3])``."""
    dx = np.array([length * 0.5, 0.0, 0.0], dtype=np.float32)
    dy = np.array([0.0, width * 0.5, 0.0], dtype=np.float32)
    dz = np.array([0.0, 0.0, height * 0.5], dtype=np.float32)
    vertices = np.stack(
        (
            +dx + dy + dz,
            +dx + dy - dz,
            -dx + dy - dz,
            -dx + dy + dz,
            -dx - dy - dz,
            -dx - dy + dz,
            +dx - dy - dz,
            +dx - dy + dz,
        )
    ).astype(np.float32)
    vertices = vertices + np.asarray(center, dtype=np.float32)
    parts = [_BOX_SIDES]
    if with_bottom:
        parts.append(_BOX_BOTTOM)
    if with_top:
        parts.append(_BOX_TOP)
    return vertices, np.concatenate(parts, axis=0).astype(np.int32)


def _merge(parts: list[tuple[np.ndarray, np.ndarray]]) -> tuple[np.ndarray, np.ndarray]:
    vs, ts, off = [], [], 0
    for v, t in parts:
        vs.append(v)
        ts.append(t + off)
        off += v.shape[0]
    return np.concatenate(vs).astype(np.float32), np.concatenate(ts).astype(np.int32)


def _ground(xmin: float, xmax: float, ymin: float, ymax: float) -> tuple[np.ndarray, np.ndarray]:
    v = np.array(
        [[xmin, ymin, 0.0], [xmax, ymin, 0.0], [xmax, ymax, 0.0], [xmin, ymax, 0.0]],
        dtype=np.float32,
    )
    t = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    return v, t


def urban_grid(
    nx: int, ny: int, *, pitch: float = 30.0, footprint: float = 10.0, seed: int = 1234
) -> tuple[np.ndarray, np.ndarray]:
    """``nx × ny`` boxes (12 triangles each, heights U(10,40) m) on a 2-triangle ground.

    ``urban_grid(29, 29)`` → 10 094 triangles (S10k); ``urban_grid(64, 65)`` → 49 922 (S50k).
    """
    rng = np.random.default_rng(seed)
    heights = rng.uniform(10.0, 40.0, size=(nx, ny))
    parts = []
    for i in range(nx):
        for j in range(ny):
            h = float(heights[i, j])
            parts.append(
                box(
                    footprint,
                    footprint,
                    h,
                    with_top=True,
                    with_bottom=True,
                    center=(i * pitch, j * pitch, 0.5 * h),
                )
            )
    m = 0.5 * pitch
    parts.append(_ground(-m, (nx - 1) * pitch + m, -m, (ny - 1) * pitch + m))
    return _merge(parts)


def street_canyon(
    n_per_row: int = 41, *, street_width: float = 20.0, footprint: float = 10.0, seed: int = 1234
) -> tuple[np.ndarray, np.ndarray]:
    """Two rows of ``n_per_row`` boxes either side of a street + ground: 986 triangles (S1k)."""
    rng = np.random.default_rng(seed)
    heights = rng.uniform(10.0, 40.0, size=(2, n_per_row))
    parts = []
    y_off = 0.5 * (street_width + footprint)
    for row, y in enumerate((-y_off, +y_off)):
        for i in range(n_per_row):
            h = float(heights[row, i])
            parts.append(
                box(
                    footprint,
                    footprint,
                    h,
                    with_top=True,
                    with_bottom=True,
                    center=(i * footprint, y, 0.5 * h),
                )
            )
    m = footprint
    parts.append(_ground(-m, n_per_row * footprint, -y_off - m, y_off + m))
    return _merge(parts)


def receivers_grid(
    vertices: np.ndarray, m: int, n: int | None = None, *, height: float = 1.5
) -> np.ndarray:
    """``m × n`` receivers at ``z = height`` over the mesh's bounding box
    (cf. ``Scene.with_receivers_grid``, ``differt/src/differt/geometry/_scene.py:377-407``)."""
    n = m if n is None else n
    lo, hi = vertices.min(axis=0), vertices.max(axis=0)
    x = np.linspace(lo[0], hi[0], m, dtype=np.float32)
    y = np.linspace(lo[1], hi[1], n, dtype=np.float32)
    xx, yy = np.meshgrid(x, y, indexing="ij")
    return np.stack((xx, yy, np.full_like(xx, height)), axis=-1).reshape(-1, 3).astype(np.float32)


def num_complete_graph_candidates(num_nodes: int, order: int) -> int:
    """``n (n-1)^(k-1)`` (``differt-core/src/geometry/graph.rs:356-362``); 1 for order 0."""
    if order == 0:
        return 1
    if num_nodes == 0:
        return 0
    return num_nodes * (num_nodes - 1) ** (order - 1)


def complete_graph_candidates(
    num_nodes: int, order: int, start: int = 0, count: int | None = None
) -> np.ndarray:
    """All walks of ``order`` nodes without consecutive repeats, in the reference's DFS
    (lexicographic) order, as a closed-form decode of the linear index
    (``differt-core/src/geometry/graph.rs:400-470`` enumerates the same sequence).

    Candidate ``i`` ↦ digits of ``i`` in base ``(n-1)`` after the leading base-``n`` digit; each
    later digit ``g`` maps to node ``g + (g >= previous)``.
    """
    total = num_complete_graph_candidates(num_nodes, order)
    count = total - start if count is None else min(count, total - start)
    out = np.empty((max(count, 0), order), dtype=np.int32)
    if order == 0 or count <= 0:
        return out
    idx = np.arange(start, start + count, dtype=np.int64)
    base = max(num_nodes - 1, 1)
    digits = np.empty((count, order), dtype=np.int64)
    rem = idx
    for j in range(order - 1, 0, -1):
        digits[:, j] = rem % base
        rem = rem // base
    digits[:, 0] = rem
    out[:, 0] = digits[:, 0]
    for j in range(1, order):
        out[:, j] = digits[:, j] + (digits[:, j] >= out[:, j - 1])
    return out


def sampled_candidates(num_triangles: int, order: int, count: int, *, seed: int = 1234) -> np.ndarray:
    """``count`` random index tuples with consecutive repeats re-drawn."""
    rng = np.random.default_rng(seed)
    cand = rng.integers(0, num_triangles, size=(count, order))
    for j in range(1, order):
        same = cand[:, j] == cand[:, j - 1]
        while same.any():
            cand[same, j] = rng.integers(0, num_triangles, size=int(same.sum()))
            same = cand[:, j] == cand[:, j - 1]
    return cand.astype(np.int32)
