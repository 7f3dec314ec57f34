#!/usr/bin/env python
"""CPU model of the culled traversal (csrc/walk.cuh) on a sample of the bench workload: counts node
steps, group steps and Möller–Trumbore tests per candidate for a given grouping of the Morton-ordered
triangles and a given stack discipline, so that hierarchy / ordering ideas can be ranked before any
GPU time is spent on them.  A development tool, not product code: the path vertices it walks come from the CPU
oracle (oracle/c_oracle.py).  The node test is the plain slab test (the margin m of cull.cuh is ~1e-4 m
here and the grazing guard is rare); the triangle test is Möller–Trumbore in float64.

    python tools/sim_walk.py [--cand 192] [--rx 48] [--grouping fixed|greedy] [--order lifo|near]
"""

from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from differt_b200 import scenes  # noqa: E402  (NumPy only)
from oracle import c_oracle as co  # noqa: E402

EPS = 10 * np.finfo(np.float32).eps
THR = 1 - 100 * np.finfo(np.float32).eps


def spread10(v):
    v = v.astype(np.uint64)
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton_order(tv):
    c = tv[:, 0] + ((tv[:, 1] - tv[:, 0]) + (tv[:, 2] - tv[:, 0])) / 3
    lo, hi = c.min(0), c.max(0)
    ext = (hi - lo).max()
    q = np.clip((c - lo) / ext * 1024, 0, 1023).astype(np.uint32)
    key = 2 + ((spread10(q[:, 0]) << 2) | (spread10(q[:, 1]) << 1) | spread10(q[:, 2]))
    return np.argsort(-key.astype(np.int64), kind="stable")


def sa(lo, hi):
    d = np.maximum(hi - lo, 0)
    return 2 * (d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 0] * d[..., 2])


def group_fixed(tlo, thi, n):
    return [list(range(i, min(i + 8, n))) for i in range(0, n, 8)]


def group_greedy(tlo, thi, n, window=64, alpha=1.3):
    """Groups of <= 8 consecutive triangles; a group is closed early when the next triangle would
    inflate its box: SA(union) > alpha * (SA(group) + SA(triangle))  (boxes of flat triangles have the
    area of their two faces, so the sum is the natural scale)."""
    groups = []
    for w0 in range(0, n, window):
        cur, lo, hi = [], None, None
        for i in range(w0, min(w0 + window, n)):
            if cur:
                ulo, uhi = np.minimum(lo, tlo[i]), np.maximum(hi, thi[i])
                if len(cur) == 8 or sa(ulo, uhi) > alpha * (sa(lo, hi) + sa(tlo[i], thi[i])):
                    groups.append(cur)
                    cur = []
            if not cur:
                cur, lo, hi = [i], tlo[i].copy(), thi[i].copy()
            else:
                cur.append(i)
                lo, hi = ulo, uhi
        if cur:
            groups.append(cur)
    return groups


def build_levels(groups, tlo, thi):
    glo = np.array([tlo[g].min(0) for g in groups])
    ghi = np.array([thi[g].max(0) for g in groups])
    levels = [(glo, ghi)]
    while levels[0][0].shape[0] > 8:
        lo, hi = levels[0]
        n = lo.shape[0]
        m = (n + 7) // 8
        plo = np.array([lo[8 * i:8 * i + 8].min(0) for i in range(m)])
        phi = np.array([hi[8 * i:8 * i + 8].max(0) for i in range(m)])
        levels.insert(0, (plo, phi))
    return levels


class Cull:
    """cull.cuh:node_culled in float64 for meshes whose normals are the coordinate axes (urban grids):
    per node the set of axes present, min sin(theta), E and R."""

    def __init__(self, tv, groups, levels):
        e1, e2 = tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]
        nrm = np.cross(e1, e2)
        ln = np.linalg.norm(nrm, axis=-1)
        l1, l2 = np.linalg.norm(e1, axis=-1), np.linalg.norm(e2, axis=-1)
        st = ln / (l1 * l2) * 0.9999 - 1e-6
        ax = np.abs(nrm / ln[:, None]) > 0.5  # which axis the normal is
        r = np.abs(tv).max((1, 2))
        e = (l1 + l2) * 1.0001
        cur = (np.array([ax[g].any(0) for g in groups]), np.array([st[g].min() for g in groups]),
               np.array([e[g].max() for g in groups]), np.array([r[g].max() for g in groups]))
        self.info = [cur]
        for L in range(len(levels) - 2, -1, -1):
            a, s_, e_, r_ = self.info[0]
            m = levels[L][0].shape[0]
            self.info.insert(0, (np.array([a[8 * i:8 * i + 8].any(0) for i in range(m)]),
                                 np.array([s_[8 * i:8 * i + 8].min() for i in range(m)]),
                                 np.array([e_[8 * i:8 * i + 8].max() for i in range(m)]),
                                 np.array([r_[8 * i:8 * i + 8].max() for i in range(m)])))

    stats: list = []
    exact_axes = False
    kconst = 6e-6

    def seg(self, o, d):
        self.o, self.dhat = o, d / np.linalg.norm(d)
        self.rseg = max(np.abs(o).max(), np.abs(o + d).max())

    def keep(self, L, idx, lo, hi, inv):
        a, st, e, r = (x[idx] for x in self.info[L])
        aligned_zero = (self.dhat == 0)[None, :] & self.exact_axes  # a == 0 exactly for those triangles: no hit
        p = np.where(a & ~aligned_zero, np.abs(self.dhat)[None, :], np.inf).min(-1)
        g = st * (p * 0.9999 - 2.2e-6) - 2e-5
        ctr, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
        S = 1.7320509 * (np.abs(self.o - ctr) + half).max(-1) + 1e-9 * (self.rseg + r)
        with np.errstate(divide="ignore", invalid="ignore"):
            m = self.kconst * S / g * 1.0001 + 1e-6 * e + 4e-6 * (self.rseg + r)
        m = m[:, None]
        hit = slab(self.o, inv, lo - m, hi + m)
        plain = slab(self.o, inv, lo, hi)
        extra = ((g <= 0) | hit) & ~plain
        self.stats.append((L, int(plain.sum()), int(extra.sum()), int((extra & (g <= 0)).sum()), g[extra & (g > 0)], m[extra & (g > 0), 0], self.dhat))
        return (g <= 0) | hit


def slab(o, inv, lo, hi):
    with np.errstate(invalid="ignore", over="ignore"):
        a, b = (lo - o) * inv, (hi - o) * inv
    tmin = np.maximum(np.minimum(a, b).max(-1), 0.0)
    tmax = np.minimum(np.maximum(a, b).min(-1), 1.0)
    return tmin <= tmax


def mt_any(o, d, v0, e1, e2):
    h = np.cross(d, e2)
    a = np.einsum("ij,ij->i", h, e1)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(a == 0, 0.0, 1.0 / a)
    s = o - v0
    u = f * np.einsum("ij,ij->i", s, h)
    q = np.cross(s, e1)
    v = f * (q @ d)
    t = f * np.einsum("ij,ij->i", q, e2)
    return bool(((np.abs(a) > EPS) & (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > EPS) & (t < THR)).any())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--cand", type=int, default=192)
    ap.add_argument("--rx", type=int, default=48)
    ap.add_argument("--grouping", default="fixed")
    ap.add_argument("--alpha", type=float, default=1.3)
    ap.add_argument("--window", type=int, default=64)
    ap.add_argument("--order", default="lifo", help="lifo (kernel) | near (children nearest the segment end on top)")
    ap.add_argument("--group-at", type=int, default=4, help="run a group step as soon as this many groups are pending")
    ap.add_argument("--head", type=int, default=1)
    ap.add_argument("--kconst", type=float, default=6e-6)
    ap.add_argument("--aligned", action="store_true", help="exclude axes along which the segment is exactly 0 (axis-aligned meshes)")
    ap.add_argument("--exact-cull", action="store_true", help="model the margin and the grazing guard of node_culled")
    args = ap.parse_args()

    v, t = scenes.urban_grid(29, 29)
    lo, hi = v.min(0), v.max(0)
    tx = np.array([[0.5 * (lo[0] + hi[0]) + 15.0, 0.5 * (lo[1] + hi[1]) + 15.0, 1.2 * hi[2]]], np.float32)
    rx_all = scenes.receivers_grid(v, 64, 64)
    cand_all = scenes.sampled_candidates(t.shape[0], 3, 4096, seed=1234)
    rng = np.random.default_rng(7)
    rx = rx_all[rng.choice(rx_all.shape[0], args.rx, replace=False)]
    cand = cand_all[rng.choice(cand_all.shape[0], args.cand, replace=False)]
    pv, _, _ = co.trace_path_candidates(v, t, tx, rx, cand)
    paths = pv.reshape(-1, 5, 3).astype(np.float64)

    tv = v[t].astype(np.float64)
    area = 0.5 * np.linalg.norm(np.cross(tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]), axis=-1)
    head = np.argsort(-area, kind="stable")[: 32 * args.head]
    order = morton_order(tv.astype(np.float32))
    tv = tv[order]
    n = tv.shape[0]
    tlo, thi = tv.min(1), tv.max(1)
    groups = group_fixed(tlo, thi, n) if args.grouping == "fixed" else group_greedy(tlo, thi, n, args.window, args.alpha)
    levels = build_levels(groups, tlo, thi)
    leaf = len(levels) - 1
    start = 0
    while start < leaf and levels[start + 1][0].shape[0] <= 32:
        start += 1
    cull = Cull(tv, groups, levels) if args.exact_cull else None
    if cull:
        cull.exact_axes = args.aligned
        cull.kconst = args.kconst
    V0, E1, E2 = tv[:, 0], tv[:, 1] - tv[:, 0], tv[:, 2] - tv[:, 0]
    hv = v[t][head].astype(np.float64)
    H0, H1, H2 = hv[:, 0], hv[:, 1] - hv[:, 0], hv[:, 2] - hv[:, 0]
    gsize = np.array([len(g) for g in groups])
    print(f"groups {len(groups)} (mean size {gsize.mean():.2f}), levels {[l[0].shape[0] for l in levels]}, start level {start}, "
          f"leaf box area sum {sa(*levels[-1]).sum():.3e}")

    tot = dict(node_steps=0, group_steps=0, tests=0, node_lanes=0, group_lanes=0, blocked=0, head_blocked=0, first=0)
    per_unblocked = []
    for p in paths:
        segs = [(p[i], p[i + 1] - p[i]) for i in range(4)]
        blocked = False
        if args.head:
            blocked = any(mt_any(o, d, H0, H1, H2) for o, d in segs)
            tot["head_blocked"] += blocked
        ns = gs = 0
        for o, d in reversed(segs):
            if blocked:
                break
            if not np.isfinite(o).all() or not np.isfinite(d).all() or not d.any():
                continue
            with np.errstate(divide="ignore"):
                inv = np.where(np.abs(d) >= 1e-30, 1.0 / np.where(d == 0, 1, d), np.copysign(3.4e38, d))
            end = o + d
            tot["first"] += 1
            if cull:
                cull.seg(o, d)
                allidx = np.arange(levels[start][0].shape[0])
                keep = np.nonzero(cull.keep(start, allidx, *levels[start], inv))[0]
            else:
                keep = np.nonzero(slab(o, inv, *levels[start]))[0]
            nodes, grp = [], []
            (grp if start == leaf else nodes).extend((start + 1, int(i)) for i in keep)
            while nodes or grp:
                if len(grp) >= args.group_at or not nodes:
                    take = grp[-4:]
                    del grp[-4:]
                    gs += 1
                    tris = [i for _, g in take for i in groups[g]]
                    tot["group_lanes"] += len(tris)
                    tot["tests"] += len(tris)
                    if mt_any(o, d, V0[tris], E1[tris], E2[tris]):
                        blocked = True
                        break
                else:
                    take = nodes[-4:]
                    del nodes[-4:]
                    ns += 1
                    new_nodes, new_grp = [], []
                    for L, idx in take:
                        llo, lhi = levels[L]
                        ch = np.arange(8 * idx, min(8 * idx + 8, llo.shape[0]))
                        tot["node_lanes"] += len(ch)
                        k = ch[cull.keep(L, ch, llo[ch], lhi[ch], inv) if cull else slab(o, inv, llo[ch], lhi[ch])]
                        (new_grp if L == leaf else new_nodes).extend((L + 1, int(i)) for i in k)
                    if args.order == "near":  # the child nearest the segment's end is popped first
                        for lst in (new_nodes, new_grp):
                            lst.sort(key=lambda e: -np.abs(0.5 * (levels[e[0] - 1][0][e[1]] + levels[e[0] - 1][1][e[1]]) - end).sum())
                    nodes.extend(new_nodes)
                    grp.extend(new_grp)
        tot["node_steps"] += ns
        tot["group_steps"] += gs
        tot["blocked"] += blocked
        if not blocked:
            per_unblocked.append((ns, gs))
    P = paths.shape[0]
    print(f"paths {P}: blocked {tot['blocked'] / P:.3f} (head row {tot['head_blocked'] / P:.3f}); per path: "
          f"node steps {tot['node_steps'] / P:.2f}, group steps {tot['group_steps'] / P:.2f}, walk tests {tot['tests'] / P:.1f}, "
          f"first steps {tot['first'] / P:.2f}")
    print(f"lanes per node step {tot['node_lanes'] / max(tot['node_steps'], 1):.1f}, per group step "
          f"{tot['group_lanes'] / max(tot['group_steps'], 1):.1f}")
    if per_unblocked:
        u = np.array(per_unblocked)
        print(f"unblocked paths ({len(u)}): node steps {u[:, 0].mean():.1f}, group steps {u[:, 1].mean():.1f}")
    if cull:
        for L in range(len(levels)):
            st = [x for x in cull.stats if x[0] == L]
            if not st:
                continue
            plain, extra, graze = (sum(x[k] for x in st) for k in (1, 2, 3))
            gs_ = np.concatenate([x[4] for x in st]) if st else np.zeros(0)
            ms_ = np.concatenate([x[5] for x in st]) if st else np.zeros(0)
            print(f"level {L}: kept by plain slab {plain}, extra kept {extra} (g <= 0: {graze}); extra with g > 0: "
                  f"g quantiles {np.quantile(gs_, [0.1, 0.5, 0.9]) if gs_.size else None}, m quantiles {np.quantile(ms_, [0.1, 0.5, 0.9]) if ms_.size else None}")
        dh = np.array([np.sort(np.abs(x[6])) for x in cull.stats if x[0] == len(levels) - 1 and x[3] > 0])
        if dh.size:
            print("dhat (sorted abs components) of segments with un-cullable leaf nodes, quantiles of the smallest:",
                  np.quantile(dh[:, 0], [0.1, 0.5, 0.9]))
    # instruction model (ncu, SASS of path_walk_kernel): head row 260, first step ~150 (segment setup + one node test),
    # node step ~140, group step ~80, loop control ~8 per step
    model = 260 * bool(args.head) + (150 * tot["first"] + 148 * tot["node_steps"] + 88 * tot["group_steps"]) / P
    print(f"modelled warp instructions per candidate: {model:.0f}")


if __name__ == "__main__":
    main()
