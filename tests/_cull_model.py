"""NumPy restatement of the exact cull's leaf level (differt_b200/csrc/cull.cu: tri_info, AxisSet,
cull_group_nodes_kernel; cull.cuh: make_seg_cull, node_culled) — TEST INFRASTRUCTURE for
tests/test_cull_model.py, which checks on the CPU that the formulas and constants the kernels use never cull
a (segment, triangle) pair the reference's fp32 test reports as a hit.

fp32 throughout (NumPy float32 operations round once per operation, like the kernels' un-fused arithmetic);
the kernels' explicit FMAs are evaluated as round_f32(double(a) * double(b) + double(c)) — the product of two
floats is exact in double, so this differs from a true FMA only by a rare double rounding, far inside the
margins being tested.  The MUFU reciprocals of the kernel (relative error 2^-22) are exact divisions here.
"""

from __future__ import annotations

import numpy as np

F = np.float32


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def cross3(a, b):
    return np.stack((a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]), -1)


class Groups:
    """Leaf nodes over consecutive groups of 8 triangles of `tri` [T, 3, 3] f32 (T a multiple of 8)."""

    def __init__(self, tri):
        tri = np.asarray(tri, F)
        assert tri.shape[0] % 8 == 0
        v0, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]  # what drt_mesh_pack stores
        with np.errstate(all="ignore"):
            v1, v2 = v0 + e1, v0 + e2
            lo, hi = np.minimum(v0, np.minimum(v1, v2)), np.maximum(v0, np.maximum(v1, v2))
            r = np.maximum(np.abs(lo).max(-1), np.abs(hi).max(-1))
            l1, l2 = np.sqrt(dot3(e1, e1)), np.sqrt(dot3(e2, e2))
            e = (l1 + l2) * F(1.0001)
            n = cross3(e1, e2)
            ln = np.sqrt(dot3(n, n))
            st = ln / (l1 * l2)
            sin_theta = st * F(0.9999) - F(1e-6)
            good = (np.isfinite(r) & np.isfinite(l1) & np.isfinite(l2) & np.isfinite(ln) & (l1 >= F(2.0**-20))
                    & (l2 >= F(2.0**-20)) & (l1 * l2 >= F(1e-30)) & (st >= F(2.0**-10)) & (st <= F(1.001)))
            rn = np.where(good, F(1) / ln, F(0)).astype(F)
            nrm = n * rn[:, None]
            beta = np.where(good, F(2.4e-7) / st, F(0)).astype(F)
        zero = (e1 == 0) & (e2 == 0)                       # [T, 3]
        aligned = good & (zero.sum(-1) == 1)
        nrm = np.where(aligned[:, None], zero.astype(F), nrm)
        G = tri.shape[0] // 8
        self.ctr, self.half = np.zeros((G, 3), F), np.zeros((G, 3), F)
        self.axes = np.zeros((G, 3, 3), F)
        self.cos_a, self.sin_a, self.st_min = np.zeros(G, F), np.zeros(G, F), np.zeros(G, F)
        self.R, self.E = np.zeros(G, F), np.zeros(G, F)
        self.flag = np.zeros(G, bool)
        for g in range(G):
            idx = np.arange(8 * g, 8 * g + 8)
            cullable = bool(good[idx].all())
            axes, sa = [], F(0)
            for i in idx[good[idx]]:                       # AxisSet::add
                best = F(2)
                for c in axes:
                    x = cross3(nrm[i], c)
                    best = min(best, np.sqrt(dot3(x, x)))
                if not axes or (best > F(0.05) and len(axes) < 3):
                    axes.append(nrm[i].copy())
                    best = F(0)
                sa = max(sa, F(best + beta[i]))
            if cullable:
                glo, ghi = lo[idx].min(0), hi[idx].max(0)
            else:  # a degenerate triangle: the node is never culled
                glo = ghi = np.zeros(3, F)
            ctr = F(0.5) * glo + F(0.5) * ghi
            self.ctr[g] = ctr
            self.half[g] = np.maximum(ghi - ctr, ctr - glo) * F(1.000001)
            s = min(F(sa * F(1.001) + F(2e-6)), F(1))
            self.cos_a[g] = np.sqrt(max(F(1) - s * s, F(0))) * F(0.9999)
            self.sin_a[g] = s
            self.st_min[g] = max(sin_theta[idx].min(), F(0)) if (cullable and axes) else F(0)
            a0 = axes[0] if axes else np.array([1, 0, 0], F)
            self.axes[g] = [a0, axes[1] if len(axes) > 1 else a0, axes[2] if len(axes) > 2 else a0]
            self.R[g], self.E[g] = r[idx].max(), e[idx].max()
            self.flag[g] = cullable and bool(aligned[idx].all())

    def culled(self, o, d):
        """node_culled for ONE segment (o, d) against every group → bool [G]."""
        o, d = np.asarray(o, F), np.asarray(d, F)
        with np.errstate(all="ignore"):
            length = np.sqrt(dot3(d, d))
            rl = F(1) / length if length >= F(2.0**-30) else F(0)
            dhat = (d * rl).astype(F)
            dhat = np.where((d != 0) & (dhat == 0), F(1e-37), dhat)
            inv = np.where(np.abs(d) >= F(2.0**-100), F(1) / np.where(d == 0, F(1), d), np.copysign(F(3.402823466e38), d)).astype(F)
            rseg = max(np.abs(o).max(), np.abs(o + d).max())
            p = np.abs(fma(dhat[0], self.axes[:, :, 0], fma(dhat[1], self.axes[:, :, 1], dhat[2] * self.axes[:, :, 2])))
            p = np.where(self.flag[:, None] & (p == 0), F(2), p)
            pmin = p.min(-1)
            g = fma(self.st_min, fma(pmin, self.cos_a, -self.sin_a), F(-2e-5))
            sx = np.abs(o - self.ctr) + self.half
            rsum = rseg + self.R
            S = fma(F(1.7320509), sx.max(-1), F(1e-7) * rsum)
            m = fma(F(6e-6) * S, (F(1) / g) * F(1.0001), fma(F(1e-6), self.E, F(4e-6) * rsum))
            h = self.half + m[:, None]
            a = ((self.ctr - h) - o) * inv
            b = ((self.ctr + h) - o) * inv
            tmin = np.maximum(np.minimum(a, b).max(-1), F(0))
            tmax = np.minimum(np.maximum(a, b).min(-1), F(1))
            miss = tmin > fma(tmax, F(1e-5), tmax) + F(1e-30)
        return (g > 0) & (rsum >= F(1e-20)) & (rsum <= F(1e9)) & miss
