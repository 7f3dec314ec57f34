"""TEST INFRASTRUCTURE ONLY — gradient oracle of the relaxed (``smoothing_factor``) trace step.

A float64 restatement of ``_trace_path_candidates``' relaxed branch (reference
``differt/src/differt/geometry/_solvers.py:576-713``; primitives ``_utils.py:1263-1322, 1452-1476``,
``_solver_image_method.py:68-79, 110-135, 138-203, 440-454``; ``differt/src/differt/utils.py:70-89``)
written with differentiable ``torch`` ops on the CPU, so that ``torch.autograd`` plays the role of
``jax.grad`` on the reference.  Reductions use ``amin`` / ``amax`` (ties share the gradient, like
JAX's ``min`` / ``max``).  Only ``tests/`` may import this module.

**Parity unpinned** beyond the forward: the reference holds no golden gradient for this branch and jax
is not installable here; the forward of this module is checked against ``differt_oracle``'s NumPy
restatement (``tests/test_oracle_smooth.py``) and its gradient against central finite differences.

Paths whose vertices are not finite are excluded up front (their confidence is 0 / NaN and JAX's
gradient through them is NaN-contaminated; the CUDA path gives them a zero gradient).
"""

from __future__ import annotations

import numpy as np
import torch

F32_EPS = float(np.finfo(np.float32).eps)


def _sig(x, alpha):
    return torch.sigmoid(x * alpha)


def _cross(a, b):  # broadcasting cross product (torch.linalg.cross wants equal ranks)
    return torch.stack((a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
                        a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                        a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]), -1)


def _mt(o, d, v0, e1, e2, eps, alpha):
    """relaxed Möller–Trumbore → (t, hit); broadcasting on the leading axes"""
    h = _cross(d, e2)
    a = (h * e1).sum(-1)
    a = torch.where(a == 0, torch.full_like(a, float("inf")), a)
    hit = _sig(a.abs() - eps, alpha)
    f = 1.0 / a
    s = o - v0
    u = f * (s * h).sum(-1)
    one = torch.ones_like(hit)
    hit = torch.stack((hit, _sig(u, alpha), _sig(1.0 - u, alpha), one), -1).amin(-1)
    q = _cross(s, e1)
    v = f * (q * d).sum(-1)
    hit = torch.stack((hit, _sig(v, alpha), _sig(1.0 - (u + v), alpha), one), -1).amin(-1)
    t = f * (q * e2).sum(-1)
    hit = torch.minimum(hit, _sig(t - eps, alpha))
    return t, hit


def relaxed_trace(vertices, triangles, tx, rx, cand, *, mask=None, assume_quads=False, epsilon=None,
                  hit_tol=None, min_len=None, smoothing_factor=1.0, paths=None):
    """``(full [n, k+2, 3], confidence [n], path_index [n])`` for the FINITE paths among
    ``paths`` (flat indices into ``[Ntx, Nrx, C]``; all of them when ``None``).  ``vertices``, ``tx``
    and ``rx`` may require grad (float64 tensors)."""
    eps = 10 * F32_EPS if epsilon is None else float(epsilon)
    thr = 1.0 - (100 * F32_EPS if hit_tol is None else float(hit_tol))
    ml = 10 * F32_EPS if min_len is None else float(min_len)
    alpha = float(smoothing_factor)
    tris = torch.as_tensor(np.asarray(triangles, np.int64))
    cand = torch.as_tensor(np.asarray(cand, np.int64))
    C, k = cand.shape
    ntx, nrx = tx.shape[0], rx.shape[0]
    P = ntx * nrx * C
    idx = torch.arange(P) if paths is None else torch.as_tensor(np.asarray(paths, np.int64))
    c = idx % C
    irx = (idx // C) % nrx
    itx = idx // (C * nrx)

    def forward(idx_sel):
        cs, fr, to = c[idx_sel], tx[itx[idx_sel]], rx[irx[idx_sel]]
        tv = vertices[tris]                                  # [T, 3, 3]
        e_a, e_b = tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 1]
        nrm = torch.linalg.cross(e_a, e_b)
        nrm = nrm / nrm.norm(dim=-1, keepdim=True)           # _mesh.py:950-956
        cc = cand[cs]                                        # [n, k]
        mv, mn = tv[cc][:, :, 0], nrm[cc]                    # [n, k, 3]
        imgs, prev = [], fr
        for i in range(k):
            prev = prev - 2.0 * ((prev - mv[:, i]) * mn[:, i]).sum(-1, keepdim=True) * mn[:, i]
            imgs.append(prev)
        pts, prev = [None] * k, to
        for i in range(k - 1, -1, -1):
            u = imgs[i] - prev
            un = (u * mn[:, i]).sum(-1, keepdim=True)
            vn = ((mv[:, i] - prev) * mn[:, i]).sum(-1, keepdim=True)
            prev = prev + u * (vn / un)
            pts[i] = prev
        full = torch.stack([fr, *pts, to], dim=1)            # [n, k+2, 3]
        ro, rd = full[:, :-1], full[:, 1:] - full[:, :-1]
        n = full.shape[0]
        one = torch.ones(n, dtype=full.dtype)
        if k > 0:
            q = 2 if assume_quads else 1
            hits = []
            for j in range(q):
                tj = tv[cc + j]                              # [n, k, 3, 3]
                hits.append(_mt(ro[:, :-1], rd[:, :-1], tj[..., 0, :], tj[..., 1, :] - tj[..., 0, :],
                                tj[..., 2, :] - tj[..., 0, :], eps, alpha)[1])
            h = torch.stack(hits, -1).amax(-1) if q == 2 else hits[0]
            inside = torch.minimum(h.amin(-1), one)
            dp = ((full[:, :-2] - mv) * mn).sum(-1)
            dn = ((full[:, 2:] - mv) * mn).sum(-1)
            same = torch.minimum(_sig(torch.sign(dp) * torch.sign(dn), alpha).amin(-1), one)
        else:
            inside, same = one, one
        act = torch.ones(tris.shape[0], dtype=torch.bool) if mask is None else torch.as_tensor(np.asarray(mask, bool))
        ta = tv[act]                                         # active triangles only
        if ta.shape[0] > 0:
            t, hit = _mt(ro[:, :, None], rd[:, :, None], ta[None, None, :, 0], (ta[:, 1] - ta[:, 0])[None, None],
                         (ta[:, 2] - ta[:, 0])[None, None], eps, alpha)
            blocked = torch.minimum(hit, _sig(thr - t, alpha)).sum(-1).clamp(max=1.0).amax(-1)
        else:
            blocked = torch.zeros(n, dtype=full.dtype)
        small = _sig(ml - (rd * rd).sum(-1), alpha).amax(-1)
        conf = torch.stack((inside, same, 1.0 - blocked, 1.0 - small), -1).amin(-1)
        if mask is not None and k > 0:
            am = act[cc].all(-1) if not assume_quads else (act[cc] & act[cc + 1]).all(-1)
            conf = conf * am.to(conf.dtype)
        return full, conf

    with torch.no_grad():
        full0, _ = forward(torch.arange(idx.numel()))
        finite = torch.isfinite(full0).all(-1).all(-1)
    sel = torch.nonzero(finite).reshape(-1)
    full, conf = forward(sel)
    return full, conf, idx[sel]


def ray_intersect_triangle_smooth(ray_origins, ray_directions, triangle_vertices, *, epsilon=None,
                                  smoothing_factor=1.0):
    """``_utils.py:1263-1322`` with smoothing, differentiable float64 tensors, broadcasting →
    ``(t, hit)``."""
    eps = 10 * F32_EPS if epsilon is None else float(epsilon)
    tv = triangle_vertices
    return _mt(ray_origins, ray_directions, tv[..., 0, :], tv[..., 1, :] - tv[..., 0, :],
               tv[..., 2, :] - tv[..., 0, :], eps, float(smoothing_factor))


def ray_intersect_any_triangle_smooth(ray_origins, ray_directions, triangle_vertices, active=None, *,
                                      epsilon=None, hit_tol=None, smoothing_factor=1.0):
    """``_utils.py:1452-1476`` with smoothing: clipped sum over the active triangles."""
    thr = 1.0 - (100 * F32_EPS if hit_tol is None else float(hit_tol))
    alpha = float(smoothing_factor)
    t, hit = ray_intersect_triangle_smooth(ray_origins[..., None, :], ray_directions[..., None, :],
                                           triangle_vertices, epsilon=epsilon, smoothing_factor=alpha)
    term = torch.minimum(hit, _sig(thr - t, alpha))
    if active is not None:
        term = term * torch.as_tensor(np.asarray(active, bool)).to(term.dtype)
    return term.sum(-1).clamp(max=1.0)
