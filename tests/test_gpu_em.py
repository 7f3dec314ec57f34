"""GPU parity of the EM consumer (csrc/em.cu through differt_b200.em → C ABI) against
oracle/em_oracle.py.

Floating-point bar: 1e-5 relative for the Fresnel coefficients and the polarisation bases (north-star
tolerance).  The per-path coefficient carries the phase ``2 pi f s / c`` (10^3–10^4 rad at GHz
frequencies), so one unit in the last place of the path length moves it by ``eps * phase``: its bound
is ``|a| (2e-5 + 8 eps |phase|)`` — 4e-5 at 10 MHz, about 1e-2 at 2.4 GHz over 100 m.
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest
import torch

from differt_b200 import scenes
from oracle import em_oracle as eo

pytestmark = pytest.mark.gpu
F = np.float32
EPS = float(np.finfo(np.float32).eps)
KATS = json.loads((Path(__file__).parent / "golden" / "em_kats.json").read_text())


@pytest.fixture(scope="module")
def drt():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import differt_b200

    return differt_b200


def close(got, expected, rtol=1e-5, atol=1e-6):
    got = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    np.testing.assert_allclose(got, expected, rtol=rtol, atol=atol)


def test_fresnel_coefficients_match_oracle(drt):
    rng = np.random.default_rng(0)
    cos_theta = np.cos(np.linspace(0, np.pi / 2, 50, dtype=F))[None, :]
    for n_r in (
        (rng.uniform(0.01, 2.0, 100) / rng.uniform(0.01, 2.0, 100)).astype(F)[:, None],           # real, both sides of 1
        (rng.uniform(1.0, 9.0, 100) - 1j * rng.uniform(0.0, 5.0, 100)).astype(np.complex64)[:, None],  # lossy
    ):
        (r_s, r_p), (t_s, t_p) = drt.em.fresnel_coefficients(n_r, cos_theta)
        (e_rs, e_rp), (e_ts, e_tp) = eo.fresnel_coefficients(n_r, cos_theta)
        assert r_s.shape == (100, 50) and r_s.dtype == torch.complex64
        # at the critical angle the square root amplifies the last bit (see test_oracle_em.py): 2e-3 there
        near_critical = np.abs(np.asarray(n_r) ** 2 + cos_theta**2 - 1) < 1e-5
        for got, exp in ((r_s, e_rs), (r_p, e_rp), (t_s, e_ts), (t_p, e_tp)):
            got = got.numpy()
            np.testing.assert_allclose(got[~near_critical], exp[~near_critical], rtol=1e-5, atol=2e-6)
            np.testing.assert_allclose(got[near_critical], exp[near_critical], atol=1e-2)
        close(t_s, r_s.numpy() + 1, atol=2e-6)
        close(np.asarray(n_r) * t_p.numpy(), r_p.numpy() + 1, atol=2e-6)
    # scalar broadcasting (stride 0) on both operands, reflection / refraction halves
    got = drt.em.reflection_coefficients(1.5, cos_theta[0])
    exp = eo.reflection_coefficients(F(1.5), cos_theta[0])
    for g, e in zip(got, exp):
        close(g, e)
    got = drt.em.refraction_coefficients(np.complex64(2 - 1j), 0.3)
    exp = eo.refraction_coefficients(np.complex64(2 - 1j), F(0.3))
    for g, e in zip(got, exp):
        close(g, e)
    assert drt.em.fresnel_coefficients(np.zeros((0,), F), 1.0)[0][0].shape == (0,)


def test_reflection_coefficient_kats(drt):
    # differt/tests/em/test_fresnel.py:58-95
    r_s, r_p = drt.em.reflection_coefficients(1.5, 1.0)
    close(r_s, -r_p.numpy())
    r_s, r_p = drt.em.reflection_coefficients(1.5, float(np.cos(F(np.pi / 2))))
    close(r_s.numpy() ** 2, -r_p.numpy())
    _, r_p = drt.em.reflection_coefficients(1.5, float(np.cos(np.arctan(F(1.5)))))
    assert abs(r_p.item()) < 1e-6
    n_r = F(1) / F(1.5)
    r_s, r_p = drt.em.reflection_coefficients(n_r, np.cos(np.arcsin(n_r)))
    close(np.abs([r_s.item(), r_p.item()]), 1.0)
    close(drt.em.refractive_index(6.27), 2.503997, rtol=1e-6)
    assert drt.em.refractive_index(torch.tensor(4 - 1j)).is_complex()


def test_sp_directions_and_rotation(drt):
    k = KATS["sp_directions"]
    c30, s30 = float(np.cos(np.pi / 6)), float(np.sin(np.pi / 6))
    sym = {"cos30": c30, "-sin30": -s30, "+sin30": s30}
    vec = lambda rows: np.array([[sym.get(x, x) for x in r] for r in rows], F)  # noqa: E731
    (e_i_s, e_i_p), (e_r_s, e_r_p) = drt.em.sp_directions(vec(k["k_i"]), vec(k["k_r"]), vec(k["normals"]))
    close(e_i_s, vec(k["e_i_s"])), close(e_i_p, vec(k["e_i_p"])), close(e_r_p, vec(k["e_r_p"]))
    assert e_r_s is e_i_s
    # random directions + normal incidence rows + broadcasting of a single normal
    rng = np.random.default_rng(3)
    k_i = eo.normalize(rng.normal(size=(4, 257, 3)))[0]
    nrm = eo.normalize(rng.normal(size=(257, 3)))[0]
    k_i[:, :5] = -nrm[:5]
    k_r = (k_i - 2 * np.sum(k_i * nrm, -1, keepdims=True) * nrm).astype(F)
    got = drt.em.sp_directions(k_i, k_r, nrm)
    exp = eo.sp_directions(k_i, k_r, np.broadcast_to(nrm, k_i.shape))
    for g, e in zip((got[0][0], got[0][1], got[1][1]), (exp[0][0], exp[0][1], exp[1][1])):
        assert g.shape == (4, 257, 3)
        close(g, e, atol=2e-6)
    r = drt.em.sp_rotation_matrix(got[0][0], got[0][1], got[1][0], got[1][1])
    close(r, eo.sp_rotation_matrix(*(x.numpy() for x in (got[0][0], got[0][1], got[1][0], got[1][1]))), atol=1e-6)
    kk = KATS["sp_rotation_matrix"]
    r = drt.em.sp_rotation_matrix(kk["e_i_s"], kk["e_i_p"], kk["cases"][0]["e_r_s"], kk["cases"][0]["e_r_p"])
    close(r, [[0.0, 1.0], [-1.0, 0.0]], atol=1e-7)


def test_fspl_and_delay(drt):
    rng = np.random.default_rng(1)
    d = rng.uniform(1, 100, (30, 1)).astype(F)
    f = rng.uniform(0.1e9, 10e9, (1, 50)).astype(F)
    close(drt.em.fspl(d, f), eo.fspl(d, f), rtol=1e-5)
    close(drt.em.fspl(d, f, dB=True), eo.fspl(d, f, dB=True), rtol=1e-5)
    path = rng.normal(size=(20, 10, 3)).astype(F)
    close(drt.em.path_delay(path), eo.path_delay(path), rtol=1e-5, atol=0)
    close(drt.em.length_to_delay(d, speed=2.0), d / 2)


def _materials(T, rng):
    n_r = (rng.uniform(1.5, 3.0, T) - 1j * rng.uniform(0.0, 0.5, T)).astype(np.complex64)
    thickness = np.where(rng.uniform(size=T) < 0.5, rng.uniform(0.05, 0.3, T), -1.0).astype(F)
    return n_r, thickness


@pytest.mark.parametrize("order", [0, 1, 2])
def test_path_coefficients_match_oracle(drt, order):
    v, t = scenes.street_canyon(6)
    mesh = drt.Mesh.from_numpy(v, t)
    T = t.shape[0]
    rng = np.random.default_rng(order)
    tx = np.array([[25.0, 0.0, 30.0], [40.0, 2.0, 25.0]], F)
    rx = np.array([[x, y, 1.5] for x in (5.0, 22.0, 41.0, 55.0) for y in (-6.0, 5.0)], F)
    paths = drt.trace_paths(mesh, tx, rx, order)
    valid = paths.mask.cpu().numpy()
    n_valid = int(valid.sum())
    assert n_valid > 0
    comp = paths.masked()
    pv, po = comp.vertices.cpu().numpy(), comp.objects.cpu().numpy()
    normals = mesh.normals.cpu().numpy()
    n_r, thickness = _materials(T, rng)
    pair_index = np.nonzero(valid.reshape(-1))[0] // valid.shape[-1]
    for frequency in (1e7, 2.4e9):
        for pol in (("V", "V"), ("H", "H"), ("V", "H"), ("H", "V")):
            for th in (thickness, None):
                a, length, field, power = drt.em.path_coefficients(
                    paths, mesh, n_r, frequency, thickness=th, polarization=pol, accumulate=True)
                e_a, e_len = eo.path_coefficients(pv, po, normals, n_r, np.full(T, -1.0, F) if th is None else th,
                                                  frequency, *pol)
                assert a.shape == (n_valid,) and field.shape == valid.shape[:-1]
                close(length, e_len, rtol=2 * EPS, atol=0)
                amplitude = (eo.C0 / frequency) / (4 * np.pi) / e_len
                phase = 2 * np.pi * frequency * e_len / eo.C0
                bound = amplitude * (2e-5 + 8 * EPS * phase)
                err = np.abs(a.cpu().numpy() - e_a)
                assert (err <= bound).all(), (frequency, pol, float((err / bound).max()))
                # accumulation is checked against the sum of the kernel's own per-path values
                e_field, e_power = eo.accumulate(a.cpu().numpy(), pair_index, valid[..., 0].size)
                scale = np.abs(a.cpu().numpy()).max() + 1e-30
                close(field.reshape(-1), e_field, rtol=1e-5, atol=1e-5 * scale)
                close(power.reshape(-1), e_power, rtol=1e-5, atol=1e-5 * scale**2)
    # already compacted paths: per-path values only
    a2, _ = drt.em.path_coefficients(comp, mesh, n_r, 2.4e9, thickness=thickness)
    a1, _ = drt.em.path_coefficients(paths, mesh, n_r, 2.4e9, thickness=thickness)
    assert torch.equal(a1, a2)
    with pytest.raises(ValueError):
        drt.em.path_coefficients(comp, mesh, n_r, 2.4e9, accumulate=True)


def test_line_of_sight_and_conductor_properties(drt):
    # the properties test_oracle_em.py pins on the oracle, on the kernel
    rng = np.random.default_rng(2)
    r = rng.uniform(10, 1000, 1000).astype(F)
    azim = rng.uniform(0, 2 * np.pi, 1000)
    rx = np.stack((r * np.cos(azim), r * np.sin(azim), np.zeros_like(r)), -1).astype(F)
    mesh = drt.Mesh.box()
    dev = mesh.vertices.device
    vertices = torch.from_numpy(np.stack((np.zeros_like(rx), rx), -2)).to(dev)
    paths = drt.TracedPaths(vertices, torch.zeros((1000, 2), dtype=torch.int32, device=dev),
                            torch.ones(1000, dtype=torch.bool, device=dev),
                            torch.zeros((1000, 0), dtype=torch.int32, device=dev))
    for frequency in KATS["fspl_vs_los"]["frequencies"]:
        a, s = drt.em.path_coefficients(paths, mesh, None, frequency)
        close(a.abs() ** 2, 1.0 / eo.fspl(s.cpu().numpy(), F(frequency)), rtol=2e-4, atol=0)
        a, _ = drt.em.path_coefficients(paths, mesh, None, frequency, polarization=("V", "H"))
        assert torch.all(a == 0)
    # one bounce on a perfectly conducting floor (bottom face of a large box)
    v, t = scenes.box(1000.0, 1000.0, 1.0, with_bottom=True)
    v = v.copy()
    v[:, 2] -= v[:, 2].min()
    mesh = drt.Mesh.from_numpy(v, t)
    tx, rx1 = np.array([[0.0, 0.0, 10.0]], F), np.array([[30.0, 40.0, 5.0]], F)
    paths = drt.trace_paths(mesh, tx, rx1, 1)
    assert paths.num_valid_paths == 1
    expected_s = np.sqrt(30.0**2 + 40.0**2 + 15.0**2)
    for pol in (("V", "V"), ("H", "H")):
        a, s = drt.em.path_coefficients(paths, mesh, np.complex64(1e6 - 1e6j), 1e9, polarization=pol)
        close(s, [expected_s], rtol=1e-6)
        close(a.abs(), [(eo.C0 / 1e9) / (4 * np.pi * expected_s)], rtol=1e-4, atol=0)
