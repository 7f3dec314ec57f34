#!/usr/bin/env python
"""Regenerate the committed golden fixtures from the read-only reference checkout.

Run in the BUILD container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (all small, committed):

* ``two_buildings.npz``   vertices f32 [56,3] / triangles i32 [24,3] parsed from
  ``differt/tests/geometry/two_buildings.obj`` with the reference loader's rule that
  non-triangle polygons are skipped (``differt-core/src/geometry/mesh.rs:411-428``;
  ``differt/tests/geometry/test_mesh.py:125-126`` pins 24 triangles).
* ``bruxelles.npz``       the reference's own "medium" benchmark mesh
  (``docs/source/notebooks/bruxelles.obj``, 14 206 triangles;
  ``differt/tests/benchmarks/fixtures.py:52-64``) stored as f32/i32 arrays.
* ``reference_kats.json`` known-answer tables transcribed from the reference's tests,
  each entry citing the file:line it was read from.

The script only *reads data files* and writes derived arrays; no reference source code
is copied.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
HERE = Path(__file__).resolve().parent


def parse_obj_triangles_only(path: Path) -> tuple[np.ndarray, np.ndarray]:
    """OBJ → (vertices, triangles); polygons that are not triangles are dropped."""
    vertices: list[list[float]] = []
    triangles: list[list[int]] = []
    for line in path.read_text().splitlines():
        parts = line.split()
        if not parts:
            continue
        if parts[0] == "v":
            vertices.append([float(x) for x in parts[1:4]])
        elif parts[0] == "f":
            idx = [int(tok.split("/")[0]) for tok in parts[1:]]
            if len(idx) == 3:
                triangles.append([i - 1 if i > 0 else len(vertices) + i for i in idx])
    return (
        np.asarray(vertices, dtype=np.float32),
        np.asarray(triangles, dtype=np.int32),
    )


def em_kats() -> dict:
    """Known answers of the EM utilities, transcribed from the reference's tests;
    paths relative to /root/reference/differt/."""
    return {
        "_comment": "Transcribed from differt/tests/em/*.py; exact-equality asserts of the reference "
        "that depend on XLA's libm / FMA contraction are kept with the tolerance given here.",
        "constants": {
            "source": "src/differt/em/_constants.py, tests/em/test_constants.py:7-25 (scipy.constants, 1e-6)",
            "c": 299792458.0, "mu_0": 1.25663706212e-06, "epsilon_0": 8.8541878128e-12, "z_0": 376.73031341259,
        },
        "refractive_index": {
            "source": "tests/em/test_fresnel.py:16-30; Glass at 1 GHz: eps_r = 6.27 f^0 "
            "(src/differt/em/_material.py:363-366, relative_permittivity :55-67)",
            "cases": [{"epsilon_r": 1.0, "expected": 1.0}, {"epsilon_r": 6.27, "expected": 2.503997}],
        },
        "fresnel_identities": {
            "source": "tests/em/test_fresnel.py:33-55",
            "n_1_n_2_range": [0.01, 2.0], "num": 100, "theta_i": "linspace(0, pi/2, 50)",
            "identities": ["t_s == r_s + 1", "n_r * t_p == r_p + 1"], "atol": 1e-6,
        },
        "reflection_coefficients": {
            "source": "tests/em/test_fresnel.py:58-95",
            "cases": [
                {"name": "normal incidence", "n_r": 1.5, "cos_theta_i": 1.0, "expect": "r_s == -r_p"},
                {"name": "grazing incidence", "n_r": 1.5, "cos_theta_i": "cos(pi/2)", "expect": "r_s**2 == -r_p"},
                {"name": "Brewster", "n_r": 1.5, "cos_theta_i": "cos(arctan(n_r))", "expect": "r_p == 0"},
                {"name": "total reflection", "n_r": "1/1.5", "cos_theta_i": "cos(arcsin(n_r))",
                 "expect": "r_s == r_p == 1"},
            ],
        },
        "sp_directions": {
            "source": "tests/em/test_utils.py:62-90",
            "k_i": [["cos30", "-sin30", 0.0], [0.0, -1.0, 0.0]],
            "k_r": [["cos30", "+sin30", 0.0], [0.0, 1.0, 0.0]],
            "normals": [[0.0, 1.0, 0.0], [0.0, 1.0, 0.0]],
            "e_i_s": [[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]],
            "e_i_p": [["+sin30", "cos30", 0.0], [0.0, 0.0, -1.0]],
            "e_r_p": [["-sin30", "cos30", 0.0], [0.0, 0.0, 1.0]],
        },
        "sp_rotation_matrix": {
            "source": "tests/em/test_utils.py:93-128; rotation_matrix_along_z_axis "
            "(src/differt/geometry/_utils.py:282-286) = [[cos, -sin], [sin, cos]]",
            "e_i_s": [1.0, 0.0, 0.0], "e_i_p": [0.0, 1.0, 0.0],
            "cases": [
                {"e_r_s": [0.0, 1.0, 0.0], "e_r_p": [-1.0, 0.0, 0.0], "angle": "-pi/2", "atol": 1e-7},
                {"e_r_s": ["s", "s", 0.0], "e_r_p": ["-s", "s", 0.0], "angle": "-pi/4", "s": "sqrt(2)/2"},
                {"e_r_s": [1.0, 0.0, 0.0], "e_r_p": [0.0, -1.0, 0.0], "expected": [[1.0, 0.0], [0.0, -1.0]]},
            ],
        },
        "fspl": {
            "source": "tests/em/test_utils.py:131-141",
            "d_range": [1.0, 100.0], "f_range": [0.1e9, 10e9],
            "identities": ["10 log10(fspl) == fspl(dB=True)", "fspl(dB=True) == 20 log10 d + 20 log10 f - 147.55 (rtol 2e-4)"],
        },
        "fspl_vs_los": {
            "source": "tests/em/test_utils.py:144-170 (the received power of a line-of-sight link equals "
            "1 / fspl at the direction of maximum radiation)",
            "frequencies": [0.1e9, 1e9, 10e9], "r_range": [10.0, 1000.0], "rtol": 2e-4,
        },
    }


def main() -> None:
    v, t = parse_obj_triangles_only(REF / "differt/tests/geometry/two_buildings.obj")
    assert v.shape == (56, 3) and t.shape == (24, 3), (v.shape, t.shape)
    np.savez_compressed(HERE / "two_buildings.npz", vertices=v, triangles=t)

    v, t = parse_obj_triangles_only(REF / "docs/source/notebooks/bruxelles.obj")
    assert t.shape == (14206, 3), t.shape
    np.savez_compressed(HERE / "bruxelles.npz", vertices=v, triangles=t)

    kats = {
        "_comment": "Known-answer tables transcribed from the reference tests; "
        "paths relative to /root/reference/differt/tests/geometry/.",
        "two_buildings_scene": {
            "source": "fixtures.py:64-71, test_scene.py:116-160",
            "tx": [0.0, 4.9352, 22.0],
            "rx": [0.0, 10.034, 1.50],
            "rtol": 1e-6,
            "orders": {
                "0": {"vertices": [], "objects": [0, 0]},
                "1": {
                    "vertices": [
                        [-0.06917738914489746, 14.946798324584961, 8.24851131439209]
                    ],
                    "objects": [0, 8, 0],
                },
                "2": {
                    "vertices": [
                        [-0.125960111618042, 14.946202278137207, 13.787875175476074],
                        [-0.04232808202505112, 5.0, 5.629261016845703],
                    ],
                    "objects": [0, 9, 22, 0],
                },
                "3": {
                    "vertices": [
                        [-0.17936798930168152, 14.945640563964844, 16.1051082611084],
                        [-0.14879928529262543, 5.0, 10.249288558959961],
                        [-0.11822860687971115, 14.946282386779785, 4.393090724945068],
                    ],
                    "objects": [0, 9, 22, 8, 0],
                },
                "4": {
                    "vertices": [
                        [-0.233406662940979, 14.945074081420898, 17.426870346069336],
                        [-0.25651583075523376, 5.0, 12.884565353393555],
                        [-0.2796238660812378, 14.944588661193848, 8.342482566833496],
                        [-0.09397590905427933, 5.0, 3.799619674682617],
                    ],
                    "objects": [0, 9, 23, 8, 22, 0],
                },
            },
        },
        "ray_intersect_triangle_hit_table": {
            "source": "test_utils.py:555-577",
            "triangle": [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]],
            "cases": [
                {"orig": [0.5, 0.5, 1.0], "dest": [0.5, 0.5, -1.0], "expected": True},
                {"orig": [0.0, 0.0, 1.0], "dest": [1.0, 1.0, -1.0], "expected": True},
                {"orig": [0.5, 0.5, 1.0], "dest": [0.5, 0.5, 0.5], "expected": False},
                {"orig": [0.5, 0.5, 1.0], "dest": [1.0, 1.0, 1.0], "expected": False},
                {"orig": [0.5, 0.5, 1.0], "dest": [1.0, 1.0, 1.5], "expected": False},
            ],
        },
        "ray_intersect_triangle_t_and_hit": {
            "source": "test_utils.py:580-606",
            "ray_origin": [0.5, 0.5, -1.0],
            "ray_directions": [
                [0.0, 0.0, 1.0],
                [0.0, 0.0, 0.5],
                [0.0, 0.0, -1.0],
                [1.0, 0.0, 0.0],
            ],
            "triangles": [
                [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]],
                [[0.0, 0.0, 1.0], [1.0, 0.0, 1.0], [0.0, 1.0, 1.0]],
            ],
            "expected_t": [[1.0, 2.0], [2.0, 4.0], [-1.0, -2.0], [0.0, 0.0]],
            "expected_hit": [
                [True, True],
                [True, True],
                [False, False],
                [False, False],
            ],
        },
        "cube_visibility": {
            "source": "test_utils.py:438-445, 717-767 (Mesh.box(with_top=True), "
            "_mesh.py:2172-2208)",
            "cases": [
                {"vertex": [2.0, 0.0, 0.0], "expected_number": 2},
                {"vertex": [2.0, 2.0, 0.0], "expected_number": 4},
                {"vertex": [2.0, 2.0, 2.0], "expected_number": 6},
            ],
            "num_rays": [20, 10000],
        },
        "box_in_box_visibility": {
            "source": "test_utils.py:770-806 (masked count only; the un-masked 11/12 "
            "counts graze exact edges and are libm-sensitive, SURVEY.md §8c)",
            "outer": [4.0, 4.0, 4.0],
            "inner": [1.0, 1.0, 1.0],
            "tx": [-1.0, 0.0, 0.0],
            "rx": [1.0, 0.0, 0.0],
            "expected_masked_count": 10,
        },
        "image_of_vertex": {
            "source": "test_image_method.py:19-29",
            "vertices": [[0.0, 0.0, 1.0], [1.0, 2.0, 3.0]],
            "mirror_vertices": [[0.0, 0.0, 0.0]],
            "mirror_normals": [[0.0, 0.0, 1.0]],
            "expected": [[0.0, 0.0, -1.0], [1.0, 2.0, -3.0]],
        },
        "intersection_of_ray_with_plane": {
            "source": "test_image_method.py:70-91",
            "ray_origins": [[-1.0, 1.0, 0.0], [-2.0, 1.0, 0.0], [-3.0, 1.0, 0.0]],
            "ray_end": [2.0, -1.0, 0.0],
            "plane_vertices": [[0.0, 0.0, 0.0]],
            "plane_normals": [[0.0, 1.0, 0.0]],
            "expected": [[0.5, 0.0, 0.0], [0.0, 0.0, 0.0], [-0.5, 0.0, 0.0]],
        },
        "intersection_of_ray_with_plane_parallel": {
            "source": "test_image_method.py:94-130",
            "ray_origins": [[-1.0, 1.0, 0.0], [-2.0, 1.0, 0.0], [-3.0, 1.0, 0.0]],
            "ray_end": [2.0, -1.0, 0.0],
            "plane_normals": [[0.0, 0.0, 1.0]],
            "plane_vertices_off": [[0.0, 0.0, -1.0]],
            "expected_off": "inf",
            "plane_vertices_on": [[0.0, 0.0, 0.0]],
            "expected_on": "ray_origins",
        },
        "corridor": {
            "source": "fixtures.py:82-117, test_image_method.py:160-191",
            "from": [0.0, 0.0, 0.0],
            "to": [1.0, 0.0, 0.0],
            "mirror_vertices": [
                [0.0, 1.0, 0.0],
                [0.0, -1.0, 0.0],
                [0.0, 1.0, 0.0],
                [0.0, -1.0, 0.0],
            ],
            "mirror_normals": [
                [0.0, -1.0, 0.0],
                [0.0, 1.0, 0.0],
                [0.0, -1.0, 0.0],
                [0.0, 1.0, 0.0],
            ],
            "paths": [
                [0.125, 1.0, 0.0],
                [0.375, -1.0, 0.0],
                [0.625, 1.0, 0.0],
                [0.875, -1.0, 0.0],
            ],
        },
        "first_hit_jacobian_box": {
            "source": "../geometry/test_mesh.py:2029-2073 (2x2x2 box, axis-aligned rays; "
            "Jacobians of t w.r.t. origins/directions/vertices match autodiff of the "
            "brute-force path at rtol=atol=1e-5)",
        },
    }
    (HERE / "reference_kats.json").write_text(json.dumps(kats, indent=1) + "\n")
    (HERE / "em_kats.json").write_text(json.dumps(em_kats(), indent=1) + "\n")
    print("wrote", sorted(p.name for p in HERE.iterdir()))


if __name__ == "__main__":
    main()
