"""CPU-only checks of the drop-in boundary: the library loads and exports every symbol that
``include/differt_b200.h`` declares; host-only helpers behave (no kernel is launched here)."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "differt_b200.h"


@pytest.fixture(scope="module")
def lib():
    from differt_b200 import build

    build.build()
    from differt_b200 import _lib

    return _lib


def declared_symbols() -> list[str]:
    text = HEADER.read_text()
    return sorted(set(re.findall(r"DRT_API\s+[\w\s\*]+?\b(drt_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in (
        "drt_ray_intersect_triangle",
        "drt_ray_intersect_any_triangle",
        "drt_first_triangle_hit_by_ray",
        "drt_first_triangle_hit_by_ray_vjp",
        "drt_triangles_visible_from_vertex",
        "drt_image_method",
        "drt_image_method_vjp",
        "drt_trace_path_candidates",
        "drt_trace_path_candidates_vjp",
        "drt_compact_valid_paths",
        "drt_complete_graph_candidates",
    ):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(str(lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_python_prototypes_cover_the_header(lib):
    assert sorted(lib.PROTOTYPES) == declared_symbols()


def test_host_only_helpers(lib):
    L = lib.lib
    assert L.drt_abi_version() == 2
    assert L.drt_error_string(0) == b"ok"
    assert b"NULL" in L.drt_error_string(-1)
    # 48 bytes per triangle, padded to 512-triangle tiles, at least one tile
    assert L.drt_mesh_pack_bytes(0) == 512 * 48
    assert L.drt_mesh_pack_bytes(512) == 512 * 48
    assert L.drt_mesh_pack_bytes(513) == 1024 * 48
    assert L.drt_mesh_pack_bytes(10094) == 20 * 512 * 48
    assert L.drt_trace_workspace_bytes(10094, 1, 64, 128) >= 2 * 20 * 512 * 48 + 4 * 64 * 128
    assert L.drt_compact_workspace_bytes(10_000_000) > 0


def test_argument_validation_needs_no_gpu(lib):
    L = lib.lib
    # negative extents and NULL outputs are rejected before anything touches the device
    assert L.drt_ray_intersect_any_triangle(None, -1, None, None, None, 0, 0.0, 0.0, None, None) == -2
    assert L.drt_ray_intersect_any_triangle(None, 4, None, None, None, 8, 0.0, 0.0, None, None) == -1
    assert L.drt_ray_intersect_any_triangle(None, 0, None, None, None, 8, 0.0, 0.0, None, None) == 0
    assert L.drt_first_triangle_hit_by_ray(None, 4, None, None, None, 8, 0.0, 0, None, None, None) == -1
    assert L.drt_image_method(None, 5, None, 1, None, None, None, None, None, None, None, None, None) == -3
    assert (
        L.drt_trace_path_candidates(
            None, 0, 0, None, None, None, 0, 1, None, 1, None, 1, 9, None, 0.0, 0.0, 0.0, 0, None, 0,
            None, None, None, None,
        )
        == -3
    )
    with pytest.raises(lib.DrtError):
        lib.check(-4)


def test_missing_cuda_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np

    import differt_b200 as drt

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        drt.ray_intersect_any_triangle(
            np.zeros((2, 3), np.float32), np.ones((2, 3), np.float32), np.zeros((1, 3, 3), np.float32)
        )


def test_deprecated_plural_names(lib):
    import differt_b200 as drt

    with pytest.warns(DeprecationWarning):
        fn = drt.rt.rays_intersect_triangles
    assert fn is drt.geometry.ray_intersect_triangle
    with pytest.warns(DeprecationWarning):
        assert drt.rt.rays_intersect_any_triangle is drt.geometry.ray_intersect_any_triangle
    with pytest.warns(DeprecationWarning):
        assert drt.rt.triangles_visible_from_vertices is drt.geometry.triangles_visible_from_vertex
    with pytest.warns(DeprecationWarning):
        assert drt.rt.first_triangles_hit_by_rays is drt.geometry.first_triangle_hit_by_ray
    assert drt.rt.image_method is drt.geometry.image_method


def _compile_c_example(tmp_path, lib):
    """gcc -std=c99 on integration/c_abi_example.c: the header must be valid C and every entry point
    the example uses must link against the shared library."""
    import shutil
    import subprocess

    cuda = Path("/usr/local/cuda")
    if shutil.which("gcc") is None or not (cuda / "include" / "cuda_runtime_api.h").exists():
        pytest.skip("needs gcc and the CUDA runtime headers")
    exe = tmp_path / "c_abi_example"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", f"-I{cuda / 'include'}",
           str(ROOT / "integration" / "c_abi_example.c"), f"-L{lib.LIB_PATH.parent}", "-ldiffert_b200",
           f"-L{cuda / 'lib64'}", "-lcudart", f"-Wl,-rpath,{lib.LIB_PATH.parent}", "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def test_header_is_valid_c_and_a_pure_c_consumer_links(lib, tmp_path):
    assert _compile_c_example(tmp_path, lib).exists()


@pytest.mark.gpu
def test_pure_c_consumer_runs_on_the_gpu(lib, tmp_path):
    import subprocess

    exe = _compile_c_example(tmp_path, lib)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "OK: 6 valid order-1 paths" in res.stdout
    assert res.stdout.count("blocked=1") == 6 and res.stdout.count("t=0.500") == 6


def test_xla_ffi_shim_type_checks(tmp_path):
    """integration/xla_ffi.cc against the REAL C header and a declaration-only stand-in for jaxlib's
    xla/ffi/api/ffi.h (integration/stub): every drt_* call must match include/differt_b200.h, every handler's
    parameter list must match the binding it is registered with, every target the Python side registers
    (integration/differt_b200_jax.py) must be defined.  A syntax / type check, not an execution: jaxlib is not
    installable in this image."""
    import re
    import shutil
    import subprocess

    cuda = Path("/usr/local/cuda")
    if shutil.which("g++") is None or not (cuda / "include" / "cuda_runtime_api.h").exists():
        pytest.skip("needs g++ and the CUDA runtime headers")
    shim = ROOT / "integration" / "xla_ffi.cc"
    base = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", f"-I{ROOT / 'integration' / 'stub'}", f"-I{ROOT / 'include'}",
            f"-I{cuda / 'include'}"]
    res = subprocess.run([*base, str(shim)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # the check has teeth: a binding that decodes another attribute type, or a call with one argument too many, fails
    text = shim.read_text()
    bad = tmp_path / "bad_binding.cc"
    bad.write_text(text.replace('.Attr<int64_t>("batch_size")', '.Attr<float>("batch_size")', 1))
    res = subprocess.run([*base, str(bad)], capture_output=True, text=True)
    assert res.returncode != 0 and "do not match" in res.stderr
    bad = tmp_path / "bad_call.cc"
    bad.write_text(text.replace("drt_em_fresnel_coefficients(\n        stream, n,", "drt_em_fresnel_coefficients(\n        stream, n, 0,", 1))
    assert bad.read_text() != text
    assert subprocess.run([*base, str(bad)], capture_output=True, text=True).returncode != 0
    # every FFI target the Python shim registers is a handler symbol of the C++ shim
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", text))
    py = (ROOT / "integration" / "differt_b200_jax.py").read_text()
    registered = set(re.findall(r'"drt_\w+": "(\w+)"', py))
    assert registered and registered <= defined, registered - defined
