"""Deprecated plural names of the hot-path functions (``differt.rt`` before 0.10).

Reference: ``differt/src/differt/rt/__init__.py:1-45`` re-exports ``differt.geometry`` under the old
names with a ``DeprecationWarning``; ``CHANGELOG.md:48-51`` lists the renames.
"""

from __future__ import annotations

import warnings

from . import geometry as _g

_RENAMED = {
    "rays_intersect_triangles": "ray_intersect_triangle",
    "rays_intersect_any_triangle": "ray_intersect_any_triangle",
    "triangles_visible_from_vertices": "triangles_visible_from_vertex",
    "first_triangles_hit_by_rays": "first_triangle_hit_by_ray",
    "image_of_vertices_with_respect_to_mirrors": "image_of_vertex_with_respect_to_mirror",
    "intersection_of_rays_with_planes": "intersection_of_ray_with_plane",
    "consecutive_vertices_are_on_same_side_of_mirrors": "consecutive_vertices_are_on_same_side_of_mirror",
}
_SAME = ("image_method", "fibonacci_lattice", "viewing_frustum", "normalize", "assemble_path")

__all__ = [*_RENAMED, *_RENAMED.values(), *_SAME]


def __getattr__(name: str):
    if name in _RENAMED:
        new = _RENAMED[name]
        warnings.warn(
            f"differt_b200.rt.{name} is deprecated, use differt_b200.geometry.{new}",
            DeprecationWarning,
            stacklevel=2,
        )
        return getattr(_g, new)
    if name in _RENAMED.values() or name in _SAME:
        return getattr(_g, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
