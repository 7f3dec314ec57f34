"""differt_b200 — the DiffeRT geometric hot path as hand-written CUDA for NVIDIA B200 (sm_100a).

Drop-in for ``differt.geometry``'s ray–triangle, visibility and image-method functions and for the
fused trace-and-validate step behind ``Scene.trace_paths`` (see DESIGN.md / INTEGRATION.md).
Importing the package loads ``libdiffert_b200.so``; there is no CPU or PyTorch fallback.
"""

from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
from . import em, geometry, launch, rt, scenes, solvers
from .geometry import (
    consecutive_vertices_are_on_same_side_of_mirror,
    fibonacci_lattice,
    first_triangle_hit_by_ray,
    image_method,
    image_of_vertex_with_respect_to_mirror,
    intersection_of_ray_with_plane,
    path_length,
    ray_intersect_any_triangle,
    ray_intersect_triangle,
    triangles_visible_from_vertex,
    viewing_frustum,
)
from .launch import LaunchedPaths, compute_tx_mlm, launch_paths, launch_rays
from .mesh import Mesh, TracedPaths
from .scene import Scene
from .solvers import (
    VisiblePathCandidates,
    generate_all_path_candidates,
    generate_visible_path_candidates,
    trace_path_candidates,
    trace_paths,
    trace_paths_chunks_iter,
    trace_valid_path_candidates,
    trace_valid_paths,
)

__version__ = "0.1.0"

__all__ = [
    "LaunchedPaths",
    "Mesh",
    "Scene",
    "compute_tx_mlm",
    "launch_paths",
    "launch_rays",
    "path_length",
    "TracedPaths",
    "VisiblePathCandidates",
    "generate_visible_path_candidates",
    "consecutive_vertices_are_on_same_side_of_mirror",
    "em",
    "fibonacci_lattice",
    "first_triangle_hit_by_ray",
    "generate_all_path_candidates",
    "geometry",
    "image_method",
    "image_of_vertex_with_respect_to_mirror",
    "intersection_of_ray_with_plane",
    "ray_intersect_any_triangle",
    "ray_intersect_triangle",
    "rt",
    "scenes",
    "solvers",
    "trace_path_candidates",
    "trace_paths",
    "trace_paths_chunks_iter",
    "trace_valid_path_candidates",
    "trace_valid_paths",
    "triangles_visible_from_vertex",
    "viewing_frustum",
]
