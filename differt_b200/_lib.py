"""ctypes binding of ``libdiffert_b200.so`` (the C ABI declared in ``include/differt_b200.h``).

There is deliberately no fallback: if the CUDA library is missing, importing this module raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# DIFFERT_B200_LIB selects a tuning variant of the SAME library (differt_b200.build --variant)
LIB_PATH = Path(os.environ.get("DIFFERT_B200_LIB") or _PKG / "libdiffert_b200.so")

i32, i64, f32, u32 = C.c_int32, C.c_int64, C.c_float, C.c_uint32
ptr, size_t = C.c_void_p, C.c_size_t
p_i64 = C.POINTER(C.c_int64)

# name → (restype, argtypes); mirrors include/differt_b200.h one to one
PROTOTYPES: dict[str, tuple] = {
    "drt_abi_version": (C.c_int, []),
    "drt_error_string": (C.c_char_p, [C.c_int]),
    "drt_mesh_pack_bytes": (size_t, [i64]),
    "drt_mesh_pack": (C.c_int, [ptr, i64, i64, ptr, ptr, ptr, ptr]),
    "drt_mesh_pack_triangle_vertices": (C.c_int, [ptr, i64, ptr, ptr, ptr]),
    "drt_mesh_pack_sort_workspace_bytes": (size_t, [i64]),
    "drt_mesh_pack_sort_by_area": (C.c_int, [ptr, i64, ptr, ptr, size_t, ptr]),
    "drt_mesh_pack_sort_by_keys": (C.c_int, [ptr, i64, ptr, ptr, ptr, size_t, ptr]),
    "drt_ray_intersect_triangle": (
        C.c_int, [ptr, i32, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, f32, ptr, ptr]),
    "drt_ray_intersect_any_triangle": (C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, f32, ptr, ptr]),
    "drt_any_hit_workspace_bytes": (size_t, [i64]),
    "drt_ray_intersect_any_triangle_culled": (
        C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, f32, ptr, size_t, ptr, ptr]),
    "drt_first_triangle_hit_by_ray": (
        C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, i64, ptr, ptr, ptr]),
    "drt_first_triangle_hit_by_ray_culled": (
        C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, i64, ptr, size_t, ptr, ptr, ptr]),
    "drt_first_triangle_hit_by_ray_vjp": (
        C.c_int, [ptr, i64, i64, i64, ptr, ptr, ptr, ptr, ptr, ptr, ptr, ptr, ptr]),
    "drt_triangles_visible_from_vertex": (
        C.c_int, [ptr, i64, i64, ptr, ptr, ptr, i64, f32, ptr, ptr]),
    "drt_image_method": (
        C.c_int, [ptr, i32, p_i64, i32, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr]),
    "drt_image_method_vjp": (
        C.c_int,
        [ptr, i32, p_i64, i32, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr, ptr, ptr, ptr, ptr]),
    "drt_image_of_vertex_with_respect_to_mirror": (
        C.c_int, [ptr, i32, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr]),
    "drt_intersection_of_ray_with_plane": (
        C.c_int, [ptr, i32, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr]),
    "drt_consecutive_vertices_are_on_same_side_of_mirror": (
        C.c_int, [ptr, i32, p_i64, i32, ptr, p_i64, ptr, p_i64, ptr, p_i64, ptr]),
    "drt_trace_workspace_bytes": (size_t, [i64, i64, i64, i64]),
    "drt_trace_path_candidates": (
        C.c_int,
        [ptr, i64, i64, ptr, ptr, ptr, i32, i64, ptr, i64, ptr, i64, i32, ptr, f32, f32, f32, u32,
         ptr, size_t, ptr, ptr, ptr, ptr]),
    "drt_trace_prepared_bytes": (size_t, [i64]),
    "drt_trace_prepare": (C.c_int, [ptr, i64, i64, ptr, ptr, ptr, ptr, size_t]),
    "drt_trace_valid_workspace_bytes": (size_t, [i64, i64]),
    "drt_trace_valid_path_candidates": (
        C.c_int,
        [ptr, i64, i64, ptr, ptr, ptr, i32, i64, ptr, i64, ptr, i64, i32, ptr, f32, f32, f32, u32, i64, ptr, size_t,
         ptr, ptr, ptr, ptr, ptr]),
    "drt_trace_path_candidates_vjp": (
        C.c_int, [ptr, i64, i64, ptr, ptr, i64, ptr, i64, ptr, i64, i32, ptr, ptr, ptr, ptr, ptr]),
    "drt_profile_reset": (C.c_int, []),
    "drt_profile_count": (C.c_int, []),
    "drt_profile_elapsed_ms": (C.c_int, [i32, C.POINTER(C.c_float)]),
    "drt_compact_workspace_bytes": (size_t, [i64]),
    "drt_compact_valid_paths": (
        C.c_int, [ptr, i64, i32, ptr, ptr, ptr, i64, ptr, size_t, ptr, ptr, ptr, ptr]),
    "drt_complete_graph_candidates": (C.c_int, [ptr, i64, i32, i64, i64, i32, ptr]),
    "drt_ray_intersect_triangle_smooth": (
        C.c_int, [ptr, i32, p_i64, ptr, p_i64, ptr, p_i64, ptr, p_i64, f32, f32, ptr, ptr]),
    "drt_ray_intersect_any_triangle_smooth": (C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, f32, f32, ptr]),
    "drt_consecutive_vertices_are_on_same_side_of_mirror_smooth": (
        C.c_int, [ptr, i32, p_i64, i32, ptr, p_i64, ptr, p_i64, ptr, p_i64, f32, ptr]),
    "drt_ray_intersect_triangle_smooth_vjp": (
        C.c_int, [ptr, i64, ptr, ptr, ptr, f32, f32, ptr, ptr, ptr, ptr, ptr]),
    "drt_ray_intersect_any_triangle_smooth_vjp": (
        C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, f32, f32, ptr, ptr, ptr, ptr]),
    "drt_trace_smooth_workspace_bytes": (size_t, [i64, i64, i64, i64]),
    "drt_trace_path_candidates_smooth": (
        C.c_int,
        [ptr, i64, i64, ptr, ptr, ptr, i32, i64, ptr, i64, ptr, i64, i32, ptr, f32, f32, f32, f32, ptr, size_t,
         ptr, ptr, ptr]),
    "drt_trace_smooth_vjp_workspace_bytes": (size_t, [i64, i64, i64, i64, i64, i32]),
    "drt_trace_path_candidates_smooth_vjp": (
        C.c_int,
        [ptr, i64, i64, ptr, ptr, ptr, i32, i64, ptr, i64, ptr, i64, i32, ptr, f32, f32, f32, f32, ptr, ptr, ptr,
         ptr, ptr, size_t, ptr, ptr, ptr]),
    "drt_em_fresnel_coefficients": (C.c_int, [ptr, i64, ptr, i64, ptr, i64, ptr, ptr, ptr, ptr]),
    "drt_em_sp_directions": (C.c_int, [ptr, i64, ptr, ptr, ptr, ptr, ptr, ptr]),
    "drt_em_path_coefficients": (
        C.c_int, [ptr, i64, i32, ptr, ptr, i64, ptr, ptr, ptr, C.c_double, i32, i32, ptr, ptr, ptr, i64, ptr, ptr]),
    "drt_bvh_bytes": (size_t, [i64]),
    "drt_bvh_workspace_bytes": (size_t, [i64]),
    "drt_bvh_build": (C.c_int, [ptr, i64, ptr, f32, ptr, size_t, ptr]),
    "drt_bvh_ray_intersect_any_triangle": (C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, f32, ptr]),
    "drt_bvh_first_triangle_hit_by_ray": (C.c_int, [ptr, i64, ptr, ptr, ptr, i64, f32, i64, ptr, ptr]),
    "drt_scatter_visible": (C.c_int, [ptr, i64, i64, i64, ptr, ptr]),
    "drt_sbr_bounce": (C.c_int, [ptr, i64, i64, i64, i64, ptr, ptr, ptr, ptr, ptr, ptr, ptr, f32, ptr, ptr]),
    "drt_mlm_step": (
        C.c_int,
        [ptr, i64, i64, i64, ptr, ptr, ptr, ptr, ptr, ptr, ptr, i32, i32, i32, f32, f32, f32, f32, f32, i32,
         i32, f32, ptr]),
    "drt_digraph_candidates_workspace_bytes": (size_t, [i64, i32]),
    "drt_digraph_candidates_prepare": (C.c_int, [ptr, i64, i32, ptr, ptr, ptr, ptr, size_t, ptr]),
    "drt_digraph_candidates": (C.c_int, [ptr, i64, i32, ptr, i64, i64, i32, ptr]),
}

DRT_TRACE_DENSE_BLOCKAGE = 1
DRT_TRACE_PROFILE = 2
DRT_TRACE_PREPARED = 4
DRT_MAX_ORDER = 8
DRT_MAX_BATCH_DIMS = 4
DRT_TILE_TRIANGLES = 512


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m differt_b200.build` "
            "(differt_b200 has no CPU or PyTorch fallback)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the ABI
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.drt_abi_version() != 2:
        raise ImportError("libdiffert_b200.so ABI version mismatch; rebuild the library")
    return lib


class _CurrentStream:
    """Placeholder for the stream argument (``_tensor.stream_ptr()``): resolved by the call wrapper
    once the device of the operands is known."""


CURRENT_STREAM = _CurrentStream()
_tls = threading.local()


def begin_call() -> _CurrentStream:
    """Start collecting the operands of one library call (always evaluated first: it is the first
    argument of every stream-ordered entry point)."""
    _tls.device = None
    return CURRENT_STREAM


def note_device(index: int) -> None:
    """Called by ``_tensor.ptr`` for every device operand of the call being assembled."""
    seen = getattr(_tls, "device", None)
    if seen is None:
        _tls.device = index
    elif seen != index:
        raise ValueError(f"differt_b200: operands live on different devices (cuda:{seen} and cuda:{index})")


class _Library:
    """The C ABI with device-correct launches: a call whose first argument is ``stream_ptr()`` runs
    with the operands' device current and on THAT device's current stream — the kernels' dynamic
    shared-memory attributes, the SM count and the stream are all per device (a process may drive
    several GPUs, as JAX's single-process mode does)."""

    def __init__(self, cdll: C.CDLL) -> None:
        self._cdll = cdll

    def __getattr__(self, name: str):
        fn = getattr(self._cdll, name)

        def call(*args):
            if not args or args[0] is not CURRENT_STREAM:
                return fn(*args)
            import torch

            dev = getattr(_tls, "device", None)
            _tls.device = None
            if dev is None or dev == torch.cuda.current_device():
                return fn(C.c_void_p(torch.cuda.current_stream().cuda_stream), *args[1:])
            with torch.cuda.device(dev):
                return fn(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), *args[1:])

        call.__name__ = name
        setattr(self, name, call)  # cache: __getattr__ is only consulted on a miss
        return call


lib = _Library(_load())


class DrtError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise DrtError(f"differt_b200: {lib.drt_error_string(rc).decode()} (code {rc})")


def i64_array(values) -> C.Array:
    values = [int(v) for v in values]
    return (C.c_int64 * max(len(values), 1))(*values)
