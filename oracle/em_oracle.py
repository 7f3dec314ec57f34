"""CPU oracle of the first EM consumer of the traced paths — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy (complex64 / float32) restatement of the reference's Fresnel / polarisation utilities and of
the way its consumers combine them into one complex coefficient per path.  Citations are relative
to /root/reference/differt/src/differt/.

* refractive_index, fresnel_coefficients, reflection_coefficients, refraction_coefficients
  — em/_fresnel.py:9-43, 46-213, 216-487, 490-516
* sp_directions, sp_rotation_matrix, fspl, length_to_delay, path_delay — em/_utils.py:13-80, 84-262,
  267-302, 344-367; perpendicular_vector — geometry/_utils.py:100-108; normalize — :66-72
* path_coefficients: the per-path field chain of plugins/deepmimo.py:337-405 (spherical basis, slab
  reflection coefficients) and :516-665 (J = R_out diag(r_s, r_p) R_in per interaction, product along
  the path, projection on the receive polarisation, spreading 1/s and phase exp(-j 2 pi f s / c),
  lambda / 4 pi).  The reference's own `transition_matrix` (em/_utils.py:306-341) raises
  NotImplementedError, so the plugin is the only place where the reference composes these pieces.

Pinned by the reference's own known answers (tests/golden/em_kats.json, transcribed by
tests/golden/make_golden.py from differt/tests/em/test_fresnel.py:58-95 and test_utils.py:62-136).
The composed `path_coefficients` has no golden vector in the reference (its DeepMIMO tests need
Sionna scenes): parity of that function is UNPINNED beyond its pinned building blocks.
"""

from __future__ import annotations

import numpy as np

C0 = 299792458.0            # em/_constants.py
MU0 = 1.25663706212e-06
EPS0 = 8.8541878128e-12
Z0 = 376.73031341259


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def safe_divide(num, den):  # utils.py:60-67
    num, den = np.asarray(num), np.asarray(den)
    zero = den == 0
    den = np.where(zero, np.ones_like(den), den)
    return np.where(zero, np.zeros(np.broadcast(num, den).shape, dtype=np.result_type(num, den)), num / den)


def normalize(x, keepdims=False):  # geometry/_utils.py:66-72
    x = _f32(x)
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    out = x / np.where(n == 0, np.float32(1), n)
    return out.astype(np.float32), (n if keepdims else n[..., 0])


def perpendicular_vector(u):  # geometry/_utils.py:100-108
    u = _f32(u)
    z = np.zeros_like(u[..., 0])
    v = np.where((np.abs(u[..., 0]) > np.abs(u[..., 1]))[..., None],
                 np.stack((-u[..., 1], u[..., 0], z), -1), np.stack((z, -u[..., 2], u[..., 1]), -1))
    return normalize(np.cross(u, v).astype(np.float32))[0]


def refractive_index(epsilon_r, mu_r=None):  # em/_fresnel.py:43
    return np.sqrt(epsilon_r if mu_r is None else epsilon_r * mu_r)


def fresnel_coefficients(n_r, cos_theta_i):  # em/_fresnel.py:183-213
    n_r = np.asarray(n_r)
    n_r = n_r.astype(np.complex64) if np.iscomplexobj(n_r) else n_r.astype(np.float32)
    c = np.abs(_f32(cos_theta_i))
    n2 = n_r * n_r
    c2 = c * c
    n2c = n2 * c
    nct = np.sqrt((n2 + c2 - np.float32(1)).astype(np.complex64))
    two_c = np.float32(2) * c
    r_s = safe_divide(c - nct, c + nct)
    t_s = safe_divide(two_c, c + nct)
    r_p = safe_divide(n2c - nct, n2c + nct)
    t_p = safe_divide(n_r * two_c, n2c + nct)
    cast = lambda x: np.asarray(x, dtype=np.complex64)  # noqa: E731
    return (cast(r_s), cast(r_p)), (cast(t_s), cast(t_p))


def reflection_coefficients(n_r, cos_theta_i):  # em/_fresnel.py:487
    return fresnel_coefficients(n_r, cos_theta_i)[0]


def refraction_coefficients(n_r, cos_theta_i):  # em/_fresnel.py:516
    return fresnel_coefficients(n_r, cos_theta_i)[1]


def sp_directions(k_i, k_r, normals):  # em/_utils.py:243-262
    k_i, k_r, normals = _f32(k_i), _f32(k_r), _f32(normals)
    e_i_s, nrm = normalize(np.cross(k_i, normals).astype(np.float32), keepdims=True)
    e_i_s = np.where(nrm == 0, perpendicular_vector(k_i), e_i_s)
    e_i_p = normalize(np.cross(e_i_s, k_i).astype(np.float32))[0]
    e_r_s = e_i_s
    e_r_p = normalize(np.cross(e_r_s, k_r).astype(np.float32))[0]
    return (e_i_s, e_i_p), (e_r_s, e_r_p)


def sp_rotation_matrix(e_a_s, e_a_p, e_b_s, e_b_p):  # em/_utils.py:289-302
    e_a_s, e_a_p, e_b_s, e_b_p = np.broadcast_arrays(_f32(e_a_s), _f32(e_a_p), _f32(e_b_s), _f32(e_b_p))
    dot = lambda a, b: np.sum(a * b, axis=-1, dtype=np.float32)  # noqa: E731
    return np.stack((np.stack((dot(e_b_s, e_a_s), dot(e_b_s, e_a_p)), -1),
                     np.stack((dot(e_b_p, e_a_s), dot(e_b_p, e_a_p)), -1)), -2).astype(np.float32)


def fspl(d, f, dB=False):  # em/_utils.py:360-367
    d, f = np.asarray(d, np.float32), np.asarray(f, np.float32)
    if dB:
        return np.float32(20) * np.log10(d) + np.float32(20) * np.log10(f) - np.float32(147.55221677811662)
    x = np.float32(4 * np.pi) * d * f / np.float32(C0)
    return x * x


def length_to_delay(length, speed=C0):  # em/_utils.py:43
    return _f32(length) / np.float32(speed)


def path_delay(path, **kw):  # em/_utils.py:74-80
    path = _f32(path)
    seg = np.diff(path, axis=-2)
    lengths = np.sqrt(np.sum(seg * seg, axis=-1, dtype=np.float32))
    return length_to_delay(lengths.sum(axis=-1, dtype=np.float32), **kw)


def spherical_basis(k):  # plugins/deepmimo.py:348-363
    k = _f32(k)
    z = np.clip(k[..., 2], -1.0, 1.0)
    theta = np.arccos(z)
    phi = np.arctan2(k[..., 1], k[..., 0])
    st, ct, sp, cp = np.sin(theta), np.cos(theta), np.sin(phi), np.cos(phi)
    return (np.stack((ct * cp, ct * sp, -st), -1).astype(np.float32),
            np.stack((-sp, cp, np.zeros_like(phi)), -1).astype(np.float32))


def slab_reflection_coefficients(n_r, cos_theta_i, thickness, wavelength):  # plugins/deepmimo.py:390-405
    r_s_inf, r_p_inf = reflection_coefficients(n_r, cos_theta_i)
    with np.errstate(all="ignore"):  # the slab branch of a half space (thickness < 0) is computed, then discarded
        return _slab(r_s_inf, r_p_inf, n_r, cos_theta_i, thickness, wavelength)


def _slab(r_s_inf, r_p_inf, n_r, cos_theta_i, thickness, wavelength):
    n_r = np.asarray(n_r, np.complex64)
    c = _f32(cos_theta_i)
    a = np.sqrt(n_r * n_r - (np.float32(1) - c * c))
    q = (np.float32(2.0 * np.pi) * _f32(thickness) / np.float32(wavelength)) * a
    e = np.exp(np.complex64(-2j) * q)
    r_s_slab = safe_divide(r_s_inf * (1 - e), 1 - r_s_inf * r_s_inf * e)
    r_p_slab = safe_divide(r_p_inf * (1 - e), 1 - r_p_inf * r_p_inf * e)
    slab = _f32(thickness) >= 0
    return (np.where(slab, r_s_slab, r_s_inf).astype(np.complex64), np.where(slab, r_p_slab, r_p_inf).astype(np.complex64))


def path_coefficients(vertices, objects, normals, n_r, thickness, frequency, tx_pol="V", rx_pol="V"):
    """One complex coefficient and one length per path (plugins/deepmimo.py:516-665 + :694-696).

    vertices [n, k+2, 3]; objects [n, k+2] (columns 1..k = triangle indices); normals [T, 3] unit;
    n_r [T] complex relative refractive index and thickness [T] (negative = half space) per triangle.
    Returns (a [n] complex64, length [n] f32).
    """
    vertices = _f32(vertices)
    n, nv, _ = vertices.shape
    k_order = nv - 2
    seg = np.diff(vertices, axis=-2)
    k, s = normalize(seg, keepdims=True)            # [n, k+1, 3], [n, k+1, 1]
    theta_hat, phi_hat = spherical_basis(k)
    one, zero = np.ones(n, np.complex64), np.zeros(n, np.complex64)
    e = np.stack((one, zero), -1) if tx_pol == "V" else np.stack((zero, one), -1)
    wavelength = C0 / frequency
    if k_order > 0:
        tri = np.asarray(objects)[:, 1:-1]
        nn = _f32(normals)[tri]                     # [n, k, 3]
        k_in, k_out = k[:, :-1], k[:, 1:]
        (e_i_s, e_i_p), (e_r_s, e_r_p) = sp_directions(k_in, k_out, nn)
        cos_i = np.sum(nn * -k_in, axis=-1, dtype=np.float32)
        r_s, r_p = slab_reflection_coefficients(np.asarray(n_r, np.complex64)[tri], cos_i, _f32(thickness)[tri], wavelength)
        rin = sp_rotation_matrix(theta_hat[:, :-1], phi_hat[:, :-1], e_i_s, e_i_p)
        rout = sp_rotation_matrix(e_r_s, e_r_p, theta_hat[:, 1:], phi_hat[:, 1:])
        d = np.zeros((n, k_order, 2, 2), np.complex64)
        d[..., 0, 0], d[..., 1, 1] = r_s, r_p
        j = np.matmul(rout.astype(np.complex64), np.matmul(d, rin.astype(np.complex64)))
        for i in range(k_order):
            e = np.matmul(j[:, i], e[..., None])[..., 0]
    a_coeff = np.sum(theta_hat[:, -1] * spherical_basis(-k[:, -1])[0], axis=-1, dtype=np.float32)
    u = np.stack((a_coeff, np.zeros_like(a_coeff)), -1) if rx_pol == "V" else np.stack((np.zeros_like(a_coeff), -a_coeff), -1)
    a = np.sum(u * e, axis=-1)
    s_tot = s.sum(axis=-2, dtype=np.float32)[..., 0]
    phase = np.float32(-2.0 * np.pi * frequency) * s_tot / np.float32(C0)
    a = a * (safe_divide(np.float32(1), s_tot) * (np.cos(phase) + 1j * np.sin(phase)).astype(np.complex64))
    a = a * np.float32(wavelength / (4 * np.pi))
    return a.astype(np.complex64), s_tot.astype(np.float32)


def accumulate(a, pair_index, num_pairs):
    """Coherent sum of the coefficients and sum of |a|^2 per (tx, rx) pair."""
    field = np.zeros(num_pairs, np.complex128)
    power = np.zeros(num_pairs, np.float64)
    np.add.at(field, pair_index, a.astype(np.complex128))
    np.add.at(power, pair_index, np.abs(a.astype(np.complex128)) ** 2)
    return field.astype(np.complex64), power.astype(np.float32)
