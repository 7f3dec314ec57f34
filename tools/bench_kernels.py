"""Per-kernel timings of every C-ABI entry point on the hot path, against the
roofline that bounds each one.  Not the contract bench (that is /bench.py): this is the
table behind BASELINE.md §6.

    gpurun -- python tools/bench_kernels.py [--json gpurun_out/kernels.json]

Timing: CUDA events on the current stream, 3 warm-ups, median of 10; every problem is larger than the
126 MB L2 unless noted.  `gbs` = algorithmic bytes / time (per-unit figures of SURVEY.md §8d).
"""

from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import differt_b200 as drt  # noqa: E402
from differt_b200 import scenes  # noqa: E402
from differt_b200._lib import check, lib  # noqa: E402
from differt_b200._tensor import ptr, stream_ptr  # noqa: E402

PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timed(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1234)
    rows = []

    def row(name, ms, units, unit_name, bytes_total=None, note=""):
        r = {"kernel": name, "ms": ms, f"{unit_name}_per_s": units / (ms * 1e-3), "note": note}
        if bytes_total is not None:
            r["gbs"] = bytes_total / (ms * 1e-3) / 1e9
            r["frac_of_hbm_peak"] = r["gbs"] / PEAK
        rows.append(r)
        print(json.dumps(r), flush=True)

    v, t = scenes.urban_grid(29, 29)
    T = t.shape[0]
    mesh = drt.Mesh.from_numpy(v, t)
    tri = mesh.triangle_vertices.contiguous()
    lo, hi = v.min(0), v.max(0)

    def rays(n):
        o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
        e = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
        o[:, 2] = rng.uniform(0.5, 45.0, n)
        e[:, 2] = rng.uniform(0.5, 45.0, n)
        return torch.from_numpy(o).to(dev), torch.from_numpy(e - o).to(dev)

    # ---- K1: element-wise Möller–Trumbore, 65 B / test ------------------------------------------------
    n = 1 << 24
    o, d = rays(n)
    tv = tri[torch.randint(0, T, (n,), device=dev)].contiguous()
    ms = timed(lambda: drt.ray_intersect_triangle(o, d, tv))
    row("K1 ray_intersect_triangle [2^24 pairs]", ms, n, "tests", 65 * n)
    del tv

    # ---- pack (+ area sort) ---------------------------------------------------------------------------
    ms = timed(lambda: drt.geometry.pack_mesh(mesh.vertices, mesh.triangles))
    row("drt_mesh_pack [10 094 tri]", ms, T, "triangles", 84 * T, "launch-bound (one small kernel)")
    pk = drt.geometry.pack_mesh(mesh.vertices, mesh.triangles)
    ms = timed(lambda: drt.geometry.sort_pack_by_area(pk, T))
    row("drt_mesh_pack_sort_by_area [10 094 tri]", ms, T, "triangles", None, "launch-bound (keys + CUB radix sort + gather)")

    # ---- K2 / K3 / K4 all-pairs: FP32-issue bound; 36 B / executed test under the streamed model ------
    R = 1 << 20
    o1, d1 = rays(R)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    pack_sorted = drt.geometry.sort_pack_by_area(pk, T)
    out8 = torch.empty(R, dtype=torch.uint8, device=dev)

    def k2():
        check(lib.drt_ray_intersect_any_triangle(stream_ptr(), R, ptr(o1), ptr(d1), ptr(pack_sorted), T,
                                                 1.1920929e-6, 1.1920929e-5, ptr(out8), ptr(cnt)))
    cnt.zero_(); k2(); executed = int(cnt.item())
    ms = timed(k2)
    row("K2 ray_intersect_any_triangle [2^20 rays x 10 094 tri, kernel only]", ms, R * T, "tests", 36 * executed,
        f"executed {executed / (R * T):.3f} of rays x triangles; blocked fraction {float(out8.float().mean()):.3f}")
    ms = timed(lambda: mesh.ray_intersect_any_triangle(o1, d1))
    row("K2 Mesh.ray_intersect_any_triangle [same, incl. pack + sort]", ms, R * T, "tests")

    idx = torch.empty(R, dtype=torch.int32, device=dev)
    tt = torch.empty(R, dtype=torch.float32, device=dev)

    def k3():
        check(lib.drt_first_triangle_hit_by_ray(stream_ptr(), R, ptr(o1), ptr(d1), ptr(pk), T, 1.1920929e-6, 512,
                                                ptr(idx), ptr(tt), None))
    ms = timed(k3)
    row("K3 first_triangle_hit_by_ray [2^20 rays x 10 094 tri]", ms, R * T, "tests", 36 * R * T, "no early exit by definition")

    # the reference's own harness size: 10 000 rays x bruxelles (14 206 triangles)
    bx = np.load(ROOT / "tests" / "golden" / "bruxelles.npz")
    bmesh = drt.Mesh.from_numpy(bx["vertices"], bx["triangles"])
    fo = torch.from_numpy(np.tile(bx["vertices"].mean(0) + np.array([0, 0, 50.0], np.float32), (10_000, 1)).astype(np.float32)).to(dev)
    fd = drt.fibonacci_lattice(10_000).to(dev) * 500.0
    ms = timed(lambda: bmesh.ray_intersect_any_triangle(fo, fd))
    row("K2 reference harness: 10 000 rays x bruxelles.obj (14 206 tri), incl. pack", ms, 10_000 * 14_206, "tests", None,
        "differt/tests/benchmarks/test_rt.py:77-100 size; L2-resident, latency-bound")
    ms = timed(lambda: bmesh.first_triangle_hit_by_ray(fo, fd))
    row("K3 reference harness: 10 000 rays x bruxelles.obj, incl. pack", ms, 10_000 * 14_206, "tests", None,
        "differt/tests/benchmarks/test_rt.py:125-148 size")
    ms = timed(lambda: bmesh.triangles_visible_from_vertex(fo[:1], num_rays=10_000))
    row("K4 reference harness: visibility, 10 000 rays x bruxelles.obj, incl. ray generation", ms, 10_000 * 14_206, "tests", None,
        "differt/tests/benchmarks/test_rt.py:103-122 size")
    txp = torch.tensor([[420.0, 420.0, 48.0]], device=dev)
    ms = timed(lambda: mesh.triangles_visible_from_vertex(txp, num_rays=1_000_000), iters=5)
    row("K4 triangles_visible_from_vertex [1 vertex, 10^6 rays (reference default) x 10 094 tri]", ms, 1_000_000 * T, "tests",
        36 * 1_000_000 * T, "includes frustum + Fibonacci lattice generation")

    # ---- N2: the opt-in BVH on the same queries ---------------------------------------------------------
    ms = timed(lambda: mesh.build_bvh())
    row("N2 drt_bvh_build [10 094 tri] incl. pack", ms, T, "triangles", None, "Morton keys, radix sort, Karras tree, refit")
    ms = timed(lambda: mesh.ray_intersect_any_triangle(o1, d1, accel="bvh"))
    row("N2 BVH any-hit [2^20 rays x 10 094 tri] incl. build", ms, R * T, "tests", None, "opt-in; brute-force-equivalent tests/s")
    ms = timed(lambda: mesh.first_triangle_hit_by_ray(o1, d1, accel="bvh"))
    row("N2 BVH first-hit [2^20 rays x 10 094 tri] incl. build", ms, R * T, "tests", None, "opt-in; brute-force-equivalent tests/s")
    ms = timed(lambda: mesh.triangles_visible_from_vertex(txp, num_rays=1_000_000, accel="bvh"), iters=5)
    row("N2 BVH visibility [1 vertex, 10^6 rays x 10 094 tri] incl. ray generation + build", ms, 1_000_000 * T, "tests", None, "opt-in")
    ms = timed(lambda: bmesh.first_triangle_hit_by_ray(fo, fd, accel="bvh"))
    row("N2 BVH first-hit, reference harness: 10 000 rays x bruxelles.obj incl. build", ms, 10_000 * 14_206, "tests", None, "opt-in")

    # ---- K5 image method, 36 k B / path ----------------------------------------------------------------
    for n5, k in ((1 << 22, 3), (10_000, 8)):
        f = torch.from_numpy(rng.uniform(-100, 100, (n5, 3)).astype(np.float32)).to(dev)
        g = torch.from_numpy(rng.uniform(-100, 100, (n5, 3)).astype(np.float32)).to(dev)
        mv = torch.from_numpy(rng.uniform(-100, 100, (n5, k, 3)).astype(np.float32)).to(dev)
        mn = torch.nn.functional.normalize(torch.from_numpy(rng.normal(size=(n5, k, 3)).astype(np.float32)).to(dev), dim=-1)
        ms = timed(lambda: drt.image_method(f, g, mv, mn))
        row(f"K5 image_method [{n5} paths x {k} mirrors]", ms, n5, "paths", (24 + 36 * k) * n5,
            "reference harness size (test_rt.py:35-53); launch-bound" if n5 == 10_000 else "")
        if n5 > 10_000:
            f.requires_grad_(True); mv.requires_grad_(True); mn.requires_grad_(True); g.requires_grad_(True)
            outp = drt.image_method(f, g, mv, mn)
            go = torch.ones_like(outp)
            ms = timed(lambda: torch.autograd.grad(outp, (f, g, mv, mn), go, retain_graph=True))
            row(f"K5b image_method VJP [{n5} paths x {k} mirrors]", ms, n5, "paths", (24 + 36 * k + 12 * k + 24 + 24 * k) * n5)

    # ---- K6 stage A alone (the API's default mode on the bench workload), K6b, compaction ---------------
    import bench

    wl = bench.build_workload(bench.DEFAULT_WORKLOAD, 0, 1)
    tx = torch.from_numpy(wl["tx"]).to(dev)
    rx = torch.from_numpy(wl["rx"]).to(dev)
    cand = torch.from_numpy(wl["cand"]).to(dev)
    P = tx.shape[0] * rx.shape[0] * cand.shape[0]
    k = wl["order"]
    ms = timed(lambda: drt.trace_path_candidates(mesh, tx, rx, cand))
    row("K6 trace_path_candidates, default mode [1 x 4096 x 4096, order 3]", ms, P, "candidate_pairs",
        (12 * (k + 2) + 4 * (k + 2) + 1) * P, "stage A writes the dense TracedPaths fields; pack + sort + blockage of survivors included")
    paths = drt.trace_path_candidates(mesh, tx, rx, cand)
    ms = timed(lambda: paths.masked())
    row("TracedPaths.masked() [16.8 M paths]", ms, P, "paths", P, "count + scan + scatter; reads the 1-byte mask twice")
    mesh_g = drt.Mesh(mesh.vertices.clone().requires_grad_(True), mesh.triangles)
    txg, rxg = tx.clone().requires_grad_(True), rx.clone().requires_grad_(True)
    pg = drt.trace_path_candidates(mesh_g, txg, rxg, cand)
    gov = torch.ones_like(pg.vertices)
    ms = timed(lambda: torch.autograd.grad(pg.vertices, (mesh_g.vertices, txg, rxg), gov, retain_graph=True), iters=5)
    row("K6b trace VJP of vertices.sum() wrt tx, rx, mesh.vertices [16.8 M paths]", ms, P, "candidate_pairs", 12 * (k + 2) * P,
        "reads the cotangent; float atomics into 3 small gradient arrays")

    # ---- BASELINE config 2 end to end: exhaustive order 2 on the street canyon, chunked --------------------
    cv, ct = scenes.street_canyon(41)
    cmesh = drt.Mesh.from_numpy(cv, ct)
    clo, chi = cv.min(0), cv.max(0)
    ctx = torch.tensor([[0.5 * (clo[0] + chi[0]), 0.0, 1.2 * chi[2]]], device=dev)
    crx = torch.from_numpy(scenes.receivers_grid(cv, 16, 16)).to(dev)
    ncand = scenes.num_complete_graph_candidates(ct.shape[0], 2)
    res = {}

    def cfg2():
        res["v"] = drt.trace_valid_paths(cmesh, ctx, crx, 2)  # one chunk: 971 210 x 256 < 2^32 paths
    ms = timed(cfg2, warmup=1, iters=3)
    row(f"config 2: exhaustive order 2, 986 tri, 1 x 256 RX, {ncand} candidates → valid paths (compact kernel)",
        ms, ncand * 256, "candidate_pairs", None, f"{res['v'].num_valid_paths} valid paths; default (pruned) mode, candidates decoded on the device")

    # the reference's city-scale notebook scene: ALL 2.02e8 order-2 candidates of bruxelles.obj, 1 TX, 1 RX
    # (docs/source/notebooks/ray_tracing_at_city_scale.ipynb cell 3: "You probably don't want to try
    #  order > 1 (too slow if testing all paths)")
    bc = bx["vertices"].mean(0)
    btx = torch.tensor([[bc[0], bc[1], float(bx["vertices"][:, 2].max()) + 10.0]], device=dev)
    brx = torch.tensor([[bc[0] + 60.0, bc[1] - 40.0, 1.5]], device=dev)
    nb = scenes.num_complete_graph_candidates(14_206, 2)

    def city():
        res["c"] = drt.trace_valid_paths(bmesh, btx, brx, 2, chunk_size=1 << 26)
    ms = timed(city, warmup=1, iters=3)
    row(f"city scale: exhaustive order 2 on bruxelles.obj (14 206 tri), 1 TX x 1 RX, {nb} candidates → valid paths",
        ms, nb, "candidate_pairs", None, f"{res['c'].num_valid_paths} valid paths; the reference's notebook calls this too slow to try")

    # ---- N1 candidate generators ------------------------------------------------------------------------
    ms = timed(lambda: drt.generate_all_path_candidates(T, 2, start=0, count=1 << 24))
    row("N1 complete-graph candidates [2^24 x order 2 of 10 094 nodes]", ms, 1 << 24, "candidates", 8 * (1 << 24))
    vis = torch.from_numpy(rng.uniform(size=T) < 0.3).to(dev)
    gen = drt.VisiblePathCandidates(T, 2, vis, vis, None)
    cnt2 = min(len(gen), 1 << 24)
    ms = timed(lambda: gen.chunk(0, cnt2))
    row(f"N1b visibility-pruned candidates [{cnt2} x order 2, 30 % visible]", ms, cnt2, "candidates", 8 * cnt2)

    if args.json:
        Path(args.json).write_text(json.dumps({"hbm_peak_gbs": PEAK, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
