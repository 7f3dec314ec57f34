"""First EM consumer of the traced paths (second half): the reference's Fresnel /
polarisation utilities with their names and argument meaning (``differt.em``), and the per-path field
chain its consumers build from them, fused with the accumulation per (transmitter, receiver) pair.

Complex results are ``torch.complex64`` tensors whose storage the kernels write as interleaved
``(re, im)`` floats.  Everything that is more than one arithmetic expression runs in
``csrc/em.cu``; the one-liners (``refractive_index``, ``fspl``, ``length_to_delay``,
``sp_rotation_matrix``) are ``torch`` expressions on the device, like ``geometry.normalize``.
"""

from __future__ import annotations

import math

import torch

from . import geometry
from ._lib import DRT_MAX_ORDER, check, lib
from ._tensor import Placement, numel, ptr, stream_ptr

# reference em/_constants.py
c: float = 299792458.0
mu_0: float = 1.25663706212e-06
epsilon_0: float = 8.8541878128e-12
z_0: float = 376.73031341259

__all__ = [
    "accumulate_fields", "c", "epsilon_0", "fresnel_coefficients", "fspl", "length_to_delay", "mu_0", "path_coefficients",
    "path_delay", "reflection_coefficients", "refraction_coefficients", "refractive_index", "sp_directions",
    "sp_rotation_matrix", "transition_matrix", "z_0",
]


def _put_inexact(pl: Placement, x) -> torch.Tensor:
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x)
    return pl.put(t, torch.complex64 if t.is_complex() else torch.float32)


def refractive_index(epsilon_r, mu_r=None):
    """``sqrt(epsilon_r mu_r)`` — reference ``em/_fresnel.py:9-43``; complex only if an input is."""
    pl = Placement()
    e = _put_inexact(pl, epsilon_r)
    if mu_r is not None:
        e = e * _put_inexact(pl, mu_r)
    return pl.out(torch.sqrt(e))


def fresnel_coefficients(n_r, cos_theta_i):
    """``((r_s, r_p), (t_s, t_p))`` — reference ``em/_fresnel.py:46-213``; ``n_r`` real or complex,
    broadcast against ``cos_theta_i``; angles outside [-pi/2, pi/2] are folded by ``abs``."""
    pl = Placement()
    n = pl.put(n_r if isinstance(n_r, torch.Tensor) else torch.as_tensor(n_r), torch.complex64)
    cth = pl.put(cos_theta_i, torch.float32)
    batch = torch.broadcast_shapes(n.shape, cth.shape)
    count = numel(batch)
    outs = [torch.empty(batch, dtype=torch.complex64, device=n.device) for _ in range(4)]
    if count > 0:
        # scalars broadcast through stride 0; anything else is materialised flat
        n_flat, n_stride = (n.reshape(1), 0) if n.numel() == 1 else (n.expand(batch).contiguous(), 1)
        c_flat, c_stride = (cth.reshape(1), 0) if cth.numel() == 1 else (cth.expand(batch).contiguous(), 1)
        check(lib.drt_em_fresnel_coefficients(
            stream_ptr(), count, ptr(torch.view_as_real(n_flat)), n_stride, ptr(c_flat), c_stride,
            *(ptr(torch.view_as_real(o)) for o in outs)))
    r_s, r_p, t_s, t_p = (pl.out(o) for o in outs)
    return (r_s, r_p), (t_s, t_p)


def reflection_coefficients(n_r, cos_theta_i):
    """Reference ``em/_fresnel.py:216-487``."""
    return fresnel_coefficients(n_r, cos_theta_i)[0]


def refraction_coefficients(n_r, cos_theta_i):
    """Reference ``em/_fresnel.py:490-516``."""
    return fresnel_coefficients(n_r, cos_theta_i)[1]


def length_to_delay(length, speed=c):
    """Reference ``em/_utils.py:13-43``."""
    pl = Placement()
    return pl.out(pl.put(length, torch.float32) / pl.put(speed, torch.float32))


def path_delay(path, **kwargs):
    """Reference ``em/_utils.py:46-80``."""
    return length_to_delay(geometry.path_length(path), **kwargs)


def sp_directions(k_i, k_r, normals):
    """``((e_i_s, e_i_p), (e_r_s, e_r_p))`` — reference ``em/_utils.py:84-262`` (normal incidence
    falls back to ``perpendicular_vector(k_i)``)."""
    pl = Placement()
    ki, kr, nn = (pl.put(x, torch.float32) for x in (k_i, k_r, normals))
    batch = torch.broadcast_shapes(ki.shape[:-1], kr.shape[:-1], nn.shape[:-1])
    e_i_s, e_i_p, e_r_p = (torch.empty((*batch, 3), dtype=torch.float32, device=ki.device) for _ in range(3))
    if numel(batch) > 0:
        flat = [x.expand(*batch, 3).contiguous() for x in (ki, kr, nn)]
        check(lib.drt_em_sp_directions(stream_ptr(), numel(batch), *(ptr(x) for x in flat), ptr(e_i_s), ptr(e_i_p),
                                       ptr(e_r_p)))
    e_i_s, e_i_p, e_r_p = pl.out(e_i_s), pl.out(e_i_p), pl.out(e_r_p)
    return (e_i_s, e_i_p), (e_i_s, e_r_p)


def sp_rotation_matrix(e_a_s, e_a_p, e_b_s, e_b_p):
    """Rotation from basis ``(e_a_s, e_a_p)`` to ``(e_b_s, e_b_p)`` — reference ``em/_utils.py:265-302``."""
    pl = Placement()
    a_s, a_p, b_s, b_p = (pl.put(x, torch.float32) for x in (e_a_s, e_a_p, e_b_s, e_b_p))
    rows = [(b_s * a_s).sum(-1), (b_s * a_p).sum(-1), (b_p * a_s).sum(-1), (b_p * a_p).sum(-1)]
    r = torch.stack(torch.broadcast_tensors(*rows), dim=-1)
    return pl.out(r.reshape(*r.shape[:-1], 2, 2))


def transition_matrix(vertices, objects, interaction_types, object_normals):
    """Reference ``em/_utils.py:305-341`` raises ``NotImplementedError`` itself; the composition its
    consumers use is :func:`path_coefficients`."""
    raise NotImplementedError


def fspl(d, f, *, dB: bool = False):  # noqa: N803
    """Free-space path loss — reference ``em/_utils.py:344-367``."""
    pl = Placement()
    dd, ff = pl.put(d, torch.float32), pl.put(f, torch.float32)
    if dB:
        return pl.out(20 * torch.log10(dd) + 20 * torch.log10(ff) - 147.55221677811662)
    x = 4 * math.pi * dd * ff / c
    return pl.out(x * x)


_POL = {"V": 0, "H": 1}


def path_coefficients(paths, mesh, n_r, frequency: float, *, thickness=None, polarization=("V", "V"),
                      accumulate: bool = False):
    """One complex coefficient per valid path: the field chain of the reference's consumer
    (``plugins/deepmimo.py:516-665``, ``:694-696``) on the compacted paths of ``paths``.

    ``n_r`` ``[T]`` complex refractive index per triangle (``refractive_index`` of the material's
    permittivity), ``thickness`` ``[T]`` slab thickness per triangle (negative or ``None``: half
    space).  Returns ``(a [n] complex64, length [n] f32)``; with ``accumulate=True`` also the coherent
    field ``[num_tx, num_rx] complex64`` and the incoherent power ``[num_tx, num_rx] f32`` summed in
    the same kernel (needs ``paths`` to carry its ``[num_tx, num_rx, C]`` mask).
    """
    if paths.order > DRT_MAX_ORDER:
        raise ValueError(f"order must be <= {DRT_MAX_ORDER}")
    tx_pol, rx_pol = (_POL[p] for p in polarization)  # KeyError on anything else, like the library's rc
    dev = mesh.vertices.device
    if accumulate and paths.mask.ndim < 2:
        raise ValueError("accumulate=True needs un-compacted paths (a mask with the [num_tx, num_rx, C] axes)")
    pair_shape = tuple(paths.mask.shape[:-1])
    index = paths._valid_flat_indices()
    comp = paths.masked(index)
    pair_index = index // int(paths.mask.shape[-1]) if accumulate else None
    v = comp.vertices.detach().to(dev, torch.float32).contiguous()
    o = comp.objects.to(dev, torch.int32).contiguous()
    n = int(v.shape[0])
    order = paths.order
    T = mesh.num_triangles
    nr = torch.as_tensor(n_r).to(dev, torch.complex64).expand(T).contiguous() if order > 0 else None
    th = None if thickness is None else torch.as_tensor(thickness).to(dev, torch.float32).expand(T).contiguous()
    pack = geometry.pack_mesh(mesh.vertices.detach(), mesh.triangles, None) if order > 0 else None
    a = torch.empty(n, dtype=torch.complex64, device=dev)
    length = torch.empty(n, dtype=torch.float32, device=dev)
    field = power = None
    num_pairs = 0
    if accumulate:
        num_pairs = numel(pair_shape)
        field = torch.zeros(pair_shape, dtype=torch.complex64, device=dev)
        power = torch.zeros(pair_shape, dtype=torch.float32, device=dev)
        pair_index = pair_index.to(dev, torch.int64).contiguous()
    if n > 0:
        check(lib.drt_em_path_coefficients(
            stream_ptr(), n, order, ptr(v), ptr(o), T, ptr(pack), ptr(None if nr is None else torch.view_as_real(nr)),
            ptr(th), float(frequency), tx_pol, rx_pol, ptr(torch.view_as_real(a)), ptr(length), ptr(pair_index), num_pairs,
            ptr(None if field is None else torch.view_as_real(field)), ptr(power)))
    return (a, length, field, power) if accumulate else (a, length)


def accumulate_fields(paths, mesh, n_r, frequency: float, **kwargs):
    """Coherent field and incoherent power per (transmitter, receiver) pair — the fused form of
    ``path_coefficients(..., accumulate=True)`` for callers that do not want the per-path values."""
    _, _, field, power = path_coefficients(paths, mesh, n_r, frequency, accumulate=True, **kwargs)
    return field, power
