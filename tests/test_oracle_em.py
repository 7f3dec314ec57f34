"""The EM oracle (oracle/em_oracle.py) against the reference's own known answers
(tests/golden/em_kats.json, transcribed from differt/tests/em/test_fresnel.py, test_utils.py and
test_constants.py) and against physical properties of the composed per-path field chain."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

from oracle import em_oracle as eo

KATS = json.loads((Path(__file__).parent / "golden" / "em_kats.json").read_text())
F = np.float32
COS30, SIN30 = float(np.cos(np.pi / 6)), float(np.sin(np.pi / 6))
SYMBOLS = {"cos30": COS30, "-sin30": -SIN30, "+sin30": SIN30, "s": np.sqrt(2) / 2, "-s": -np.sqrt(2) / 2}


def _vec(rows):
    return np.array([[SYMBOLS.get(x, x) for x in row] for row in rows], dtype=np.float32)


def test_constants():
    k = KATS["constants"]
    assert (eo.C0, eo.MU0, eo.EPS0, eo.Z0) == (k["c"], k["mu_0"], k["epsilon_0"], k["z_0"])


def test_refractive_index():
    for case in KATS["refractive_index"]["cases"]:
        np.testing.assert_allclose(eo.refractive_index(F(case["epsilon_r"])), case["expected"], rtol=1e-6)
    # complex only when an input is complex (em/_fresnel.py:33-36)
    assert not np.iscomplexobj(eo.refractive_index(F(4.0), F(1.0)))
    assert np.iscomplexobj(eo.refractive_index(np.complex64(4 - 1j)))


def test_fresnel_identities():
    k = KATS["fresnel_identities"]
    rng = np.random.default_rng(0)
    lo, hi = k["n_1_n_2_range"]
    n_1 = rng.uniform(lo, hi, k["num"]).astype(F)
    n_2 = rng.uniform(lo, hi, k["num"]).astype(F)
    n_r = (n_2 / n_1).astype(np.complex64)[:, None]
    theta_i = np.linspace(0, np.pi / 2, 50, dtype=F)
    cos_theta_i = np.cos(theta_i)[None, :]
    (r_s, r_p), (t_s, t_p) = eo.fresnel_coefficients(n_r, cos_theta_i)
    theta_c = np.arcsin(np.minimum(n_r.real, 1.0))
    for arr in (r_s, r_p, t_s, t_p):
        assert np.isfinite(np.where(theta_i <= theta_c, arr, 0.0)).all()
    assert all(np.array_equal(a, b) for a, b in zip((r_s, r_p), eo.reflection_coefficients(n_r, cos_theta_i)))
    assert all(np.array_equal(a, b) for a, b in zip((t_s, t_p), eo.refraction_coefficients(n_r, cos_theta_i)))
    np.testing.assert_allclose(t_s, r_s + 1, atol=k["atol"])
    np.testing.assert_allclose(n_r * t_p, r_p + 1, atol=k["atol"])


def test_reflection_coefficients_kats():
    n_r = F(1.5)
    r_s, r_p = eo.reflection_coefficients(n_r, F(1.0))
    np.testing.assert_allclose(r_s, -r_p, atol=1e-6)  # the reference asserts equality under XLA's division
    r_s, r_p = eo.reflection_coefficients(n_r, np.cos(F(np.pi / 2)))
    np.testing.assert_allclose(r_s**2, -r_p, atol=1e-6)
    _, r_p = eo.reflection_coefficients(n_r, np.cos(np.arctan(n_r)))
    assert abs(r_p) < 1e-6  # Brewster
    n_r = F(1) / F(1.5)
    r_s, r_p = eo.reflection_coefficients(n_r, np.cos(np.arcsin(n_r)))
    # at the critical angle sqrt(n^2 + cos^2 - 1) amplifies the last bit of the cosine to ~1e-3: the
    # modulus is pinned to 1e-6, the value to the square root of that
    np.testing.assert_allclose(np.abs([r_s, r_p]), 1.0, atol=1e-6)
    np.testing.assert_allclose([r_s, r_p], 1.0 + 0j, atol=3e-3)


def test_sp_directions_kat():
    k = KATS["sp_directions"]
    k_i, k_r, normals = _vec(k["k_i"]), _vec(k["k_r"]), _vec(k["normals"])
    (e_i_s, e_i_p), (e_r_s, e_r_p) = eo.sp_directions(k_i, k_r, normals)
    assert np.array_equal(e_i_s, e_r_s)
    for (s, p), kk in (((e_i_s, e_i_p), k_i), ((e_r_s, e_r_p), k_r)):
        np.testing.assert_allclose(np.cross(p, s), kk, atol=1e-6)
        np.testing.assert_allclose(np.cross(kk, p), s, atol=1e-6)
        np.testing.assert_allclose(np.cross(s, kk), p, atol=1e-6)
    np.testing.assert_allclose(e_i_s, _vec(k["e_i_s"]), atol=1e-6)
    np.testing.assert_allclose(e_i_p, _vec(k["e_i_p"]), atol=1e-6)
    np.testing.assert_allclose(e_r_p, _vec(k["e_r_p"]), atol=1e-6)


def test_sp_directions_normal_incidence():
    # em/_utils.py:248-254: k_i parallel to the normal → perpendicular_vector(k_i)
    k_i = np.array([[0.0, 0.0, -1.0], [1.0, 0.0, 0.0]], F)
    (e_i_s, e_i_p), (_, e_r_p) = eo.sp_directions(k_i, -k_i, -k_i)
    for v in (e_i_s, e_i_p, e_r_p):
        np.testing.assert_allclose(np.linalg.norm(v, axis=-1), 1.0, atol=1e-6)
    np.testing.assert_allclose(np.sum(e_i_s * k_i, -1), 0.0, atol=1e-6)


def _rot(angle):
    return np.array([[np.cos(angle), -np.sin(angle)], [np.sin(angle), np.cos(angle)]], F)


def test_sp_rotation_matrix_kat():
    k = KATS["sp_rotation_matrix"]
    e_i_s, e_i_p = np.array(k["e_i_s"], F), np.array(k["e_i_p"], F)
    for case in k["cases"]:
        e_r_s, e_r_p = _vec([case["e_r_s"]])[0], _vec([case["e_r_p"]])[0]
        got = eo.sp_rotation_matrix(e_i_s, e_i_p, e_r_s, e_r_p)
        if "angle" in case:
            expected = _rot({"-pi/2": -np.pi / 2, "-pi/4": -np.pi / 4}[case["angle"]])
        else:
            expected = np.array(case["expected"], F)
            np.testing.assert_allclose(np.linalg.det(got), -1.0, atol=1e-6)
        np.testing.assert_allclose(got, expected, atol=max(case.get("atol", 0.0), 1e-6))
        np.testing.assert_allclose(got @ got.T, np.eye(2), atol=1e-6)


def test_fspl_and_delay():
    k = KATS["fspl"]
    rng = np.random.default_rng(1)
    d = rng.uniform(*k["d_range"], (30, 1)).astype(F)
    f = rng.uniform(*k["f_range"], (1, 50)).astype(F)
    got, got_db = eo.fspl(d, f), eo.fspl(d, f, dB=True)
    np.testing.assert_allclose(10 * np.log10(got), got_db, rtol=1e-5)
    np.testing.assert_allclose(got_db, 20 * np.log10(d) + 20 * np.log10(f) - 147.55, rtol=2e-4)
    length = rng.uniform(1, 10, (20, 10)).astype(F)
    np.testing.assert_allclose(eo.length_to_delay(length, 2.0), length / 2.0, rtol=1e-6)
    path = rng.normal(size=(20, 10, 3)).astype(F)
    expected = np.sum(np.linalg.norm(np.diff(path, axis=-2), axis=-1), axis=-1) / eo.C0
    np.testing.assert_allclose(eo.path_delay(path), expected, rtol=1e-5)
    assert eo.path_delay(np.zeros((0, 3), F)) == 0.0  # test_utils.py:48 `(0, 3)` case


@pytest.mark.parametrize("frequency", KATS["fspl_vs_los"]["frequencies"])
def test_line_of_sight_coefficient_is_inverse_fspl(frequency):
    # test_utils.py:144-170 restated for the field chain: a horizontal line-of-sight link between
    # matched polarisations receives 1 / fspl; crossed polarisations receive nothing
    rng = np.random.default_rng(2)
    r = rng.uniform(*KATS["fspl_vs_los"]["r_range"], 1000).astype(F)
    azim = rng.uniform(0, 2 * np.pi, 1000)
    rx = np.stack((r * np.cos(azim), r * np.sin(azim), np.zeros_like(r)), -1).astype(F)
    vertices = np.stack((np.zeros_like(rx), rx), -2)
    objects = np.zeros((1000, 2), np.int32)
    for pol in ("V", "H"):
        a, s = eo.path_coefficients(vertices, objects, None, None, None, frequency, pol, pol)
        np.testing.assert_allclose(np.abs(a) ** 2, 1.0 / eo.fspl(s, F(frequency)), rtol=2e-4)
        np.testing.assert_allclose(s, r, rtol=1e-6)
    a, _ = eo.path_coefficients(vertices, objects, None, None, None, frequency, "V", "H")
    assert np.all(a == 0)


def test_ground_reflection_on_a_perfect_conductor():
    # one bounce on the plane z = 0 with |n_r| → ∞: r_s → -1, r_p → +1, so |a| = lambda / (4 pi s)
    # for both polarisations and the path length is that of the image source
    normals = np.array([[0.0, 0.0, 1.0]], F)
    n_r = np.array([1e6 - 1e6j], np.complex64)
    tx, rx = np.array([0.0, 0.0, 10.0], F), np.array([30.0, 40.0, 5.0], F)
    t = tx[2] / (tx[2] + rx[2])
    hit = np.array([rx[0] * t, rx[1] * t, 0.0], F)
    vertices = np.stack((tx, hit, rx))[None]
    objects = np.array([[0, 0, 0]], np.int32)
    frequency = 1e9
    expected_s = np.sqrt(30.0**2 + 40.0**2 + 15.0**2)
    for pol in ("V", "H"):
        a, s = eo.path_coefficients(vertices, objects, normals, n_r, np.array([-1.0], F), frequency, pol, pol)
        np.testing.assert_allclose(s, expected_s, rtol=1e-6)
        np.testing.assert_allclose(np.abs(a), (eo.C0 / frequency) / (4 * np.pi * expected_s), rtol=1e-4)
    # a slab of zero thickness reflects nothing (plugins/deepmimo.py:398-400: 1 - exp(0) = 0)
    a, _ = eo.path_coefficients(vertices, objects, normals, np.array([2.0 - 0.1j], np.complex64), np.array([0.0], F),
                                frequency)
    assert abs(a[0]) == 0


def test_accumulate():
    a = np.array([1 + 1j, 2 - 1j, -1 + 0j], np.complex64)
    field, power = eo.accumulate(a, np.array([1, 1, 0]), 3)
    np.testing.assert_allclose(field, [-1, 3 + 0j, 0])
    np.testing.assert_allclose(power, [1, 7, 0])
