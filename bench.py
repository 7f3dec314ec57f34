#!/usr/bin/env python
"""bench.py — throughput of the DiffeRT geometric hot path (trace + validate path candidates).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic input: trace-and-validate
(`_trace_path_candidates`, reference differt/src/differt/geometry/_solvers.py:499-770) of every
(tx, rx, candidate) of the workload, with the blockage test decided for every candidate like the
reference does ("dense"), the reverse mode of `vertices.sum()` w.r.t. tx, rx and the mesh vertices
(BASELINE config 3 is "with VJP"), followed by the compaction of the valid paths (`TracedPaths.masked()`)
and — for N > 1, where the FIXED workload's candidates are sharded over the ranks (strong scaling) —
ONE all-gather of the survivors.  The metric is BASELINE.json's: ray–triangle tests per second,
counted as SURVEY.md §8(d) defines it — every (ray, triangle) pair the step DECIDES, rays x
triangles ("algorithmic", what the reference's dense evaluation executes) — in BOTH arms: the GPU
arm and the CPU arm (`--impl reference`, `cpu_baseline`) both stop a candidate at its first blocker,
both report decided pairs per second, so their ratio is a ratio of times for the same job.  The
Möller–Trumbore evaluations actually executed are reported next to it in both arms
(`executed_tests_per_s`), and the roofline is computed from measured instruction counts only.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of `value`, `e2e`,
`roofline` and `cpu_baseline`.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "ray_triangle_tests_per_s"
UNIT = "tests/s"
BYTES_PER_TEST = 36  # streamed-operand model: one 9-float triangle operand per test

# name → (scene builder args, rx grid, order, candidates per GPU)
WORKLOADS = {
    # BASELINE.json configs[2]: urban scene (~10k tris), 1 TX × 4096 RX, order-3 — the configuration
    # the north star quotes the metric on ("on a 10k-triangle mesh").
    "urban10k_1tx_4096rx_order3": dict(scene=("urban", 29, 29), rx=(64, 64), order=3, cand=4096),
    # BASELINE.json configs[1], one chunk of the exhaustive candidate list
    "canyon1k_1tx_256rx_order2": dict(scene=("canyon", 41), rx=(16, 16), order=2, cand=65536),
    # BASELINE.json configs[3]: 16 TX x 4096 RX, order 3, 4096 candidates sharded over the GPUs
    "urban10k_16tx_4096rx_order3": dict(scene=("urban", 29, 29), rx=(64, 64), order=3, cand=4096, ntx=16),
    # BASELINE.json configs[4]: 50k-triangle mesh, 1 TX x 16 384 RX, order 4, 2048 candidates (sweep: --cand)
    "urban50k_1tx_16384rx_order4": dict(scene=("urban", 64, 65), rx=(128, 128), order=4, cand=2048),
    # small variant for quick checks (not a bench line)
    "urban10k_small": dict(scene=("urban", 29, 29), rx=(16, 16), order=3, cand=1024),
}
DEFAULT_WORKLOAD = "urban10k_1tx_4096rx_order3"


_SCENES = None
RX_BLOCK = int(os.environ.get("DRT_RX_BLOCK", "16"))  # = differt_b200.distributed.RX_BLOCK


def load_scenes():
    """`differt_b200/scenes.py` (NumPy only) loaded BY PATH: building a workload must not import the
    product package — importing it dlopens libdiffert_b200.so, which the reference arm never touches."""
    global _SCENES
    if _SCENES is None:
        import importlib.util

        spec = importlib.util.spec_from_file_location("_bench_scenes", ROOT / "differt_b200" / "scenes.py")
        _SCENES = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_SCENES)
    return _SCENES


def build_workload(name: str, rank: int, world: int, *, weak: bool = False):
    """Seeded synthetic inputs (host, NumPy).  The workload is FIXED (strong scaling): its receivers
    are dealt round-robin to the ranks (differt_b200.distributed.receiver_shard) and rank r traces
    receivers r, r + world, ... against EVERY candidate — every shard then costs the same to within a
    percent, which contiguous shards of a few hundred candidates do not.  `weak=True`
    gives every rank the full receiver set and `cand` candidates of its own instead (the extra
    weak-scaling leg)."""
    scenes = load_scenes()

    w = WORKLOADS[name]
    if w["scene"][0] == "urban":
        v, t = scenes.urban_grid(w["scene"][1], w["scene"][2])
    else:
        v, t = scenes.street_canyon(w["scene"][1])
    lo, hi = v.min(0), v.max(0)
    tx = np.array([[0.5 * (lo[0] + hi[0]) + 15.0, 0.5 * (lo[1] + hi[1]) + 15.0, 1.2 * hi[2]]], np.float32)
    if w.get("ntx", 1) > 1:  # 4 x 4 grid of transmitters at the same height
        n = int(round(w["ntx"] ** 0.5))
        gx = np.linspace(lo[0] + 100.0, hi[0] - 100.0, n, dtype=np.float32) + 15.0
        gy = np.linspace(lo[1] + 100.0, hi[1] - 100.0, n, dtype=np.float32) + 15.0
        xx, yy = np.meshgrid(gx, gy, indexing="ij")
        tx = np.stack((xx, yy, np.full_like(xx, 1.2 * hi[2])), -1).reshape(-1, 3).astype(np.float32)
    rx = scenes.receivers_grid(v, *w["rx"])
    total = w["cand"] * (world if weak else 1)
    cand_all = scenes.sampled_candidates(t.shape[0], w["order"], total, seed=1234)
    # candidates known to be valid for some receivers (found by tools/find_valid_candidates.py and
    # re-validated by the CPU oracle in tests/) are spread over the list, so that every shard of the
    # step produces real paths
    fixture = ROOT / "tests" / "golden" / "urban10k_valid_candidates.npz"
    if w["scene"] == ("urban", 29, 29) and fixture.exists():
        known = np.load(fixture)[f"order{w['order']}"]
        n = min(known.shape[0], total // 4)
        slots = (np.arange(n) * (total // max(n, 1))).astype(np.int64)
        cand_all[slots] = known[:n]
    if weak:
        start, stop = rank * w["cand"], (rank + 1) * w["cand"]
        return dict(name=name, vertices=v, triangles=t, tx=tx, rx=rx, order=w["order"],
                    cand=np.ascontiguousarray(cand_all[start:stop]), cand_global=total, cand_start=start,
                    rx_global=int(rx.shape[0]), receivers=None)
    idx = np.arange(rx.shape[0])
    mine = idx[(idx // RX_BLOCK) % world == rank]  # differt_b200.distributed.receiver_shard, restated (see load_scenes)
    return dict(
        name=name, vertices=v, triangles=t, tx=tx, rx=np.ascontiguousarray(rx[mine]), order=w["order"],
        cand=cand_all, cand_global=total, cand_start=0, rx_global=int(rx.shape[0]), rx_index=mine,
        receivers=(int(rx.shape[0]), mine),
    )


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

_SMI_FIELDS = (
    "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
    "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
    "clocks_event_reasons.sw_power_cap"
)
_REASONS = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")


class ClockSampler:
    """`nvidia-smi -lms 200` on this rank's GPU.  It is started at the very beginning of the run and
    the warm-up waits for its first sample (its NVML start-up contends with kernel launches for a few
    hundred ms — seconds on an 8-GPU box); it keeps polling through the timed region and only the
    samples stamped inside [mark_start(), stop()] are reported."""

    def __init__(self, device_index: int) -> None:
        import torch

        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            sel = uuid if uuid.startswith("GPU-") else f"GPU-{uuid}"
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", sel, f"--query-gpu={_SMI_FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"],
                stdout=self.file, stderr=subprocess.DEVNULL,
            )
        except Exception:  # nvidia-smi missing: report no clocks rather than fail the bench
            self.proc = None
        self.t0 = time.time()
        import atexit

        atexit.register(self._kill)  # never leave the poller behind if the bench dies early

    def _kill(self) -> None:
        if self.proc is not None and self.proc.poll() is None:
            self.proc.kill()

    def wait_ready(self, timeout: float = 8.0) -> None:
        """Block until nvidia-smi has printed its first sample (NVML initialised, steady polling) or
        `timeout` elapses — so that its start-up can never fall inside the timed region, however few
        warm-up and timed steps the caller asks for."""
        if self.proc is None:
            return
        deadline = time.time() + timeout
        while time.time() < deadline and self.proc.poll() is None:
            try:
                if os.path.getsize(self.file.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.05)

    def mark_start(self) -> None:
        self.t0 = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        t1 = time.time()
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        import datetime

        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                stamp = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if not (self.t0 - 0.2 <= stamp <= t1 + 0.2):
                    continue
            except ValueError:
                pass  # unknown timestamp format: keep the sample
            parts = parts[1:]
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(_REASONS, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.file.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {
            "sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
            "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# CPU legs (the oracle port of the reference algorithm, all host threads)
# ------------------------------------------------------------------------------------------------


def cpu_sample(wl: dict, target_tests: float):
    """A bounded sub-problem of the workload: the first c candidates × an evenly strided subset of r
    receivers, sized to about `target_tests` ray–triangle tests."""
    T = wl["triangles"].shape[0]
    per_pair = (wl["order"] + 1) * T
    pairs = max(int(target_tests / per_pair), 64)
    c = int(min(wl["cand"].shape[0], max(8, round(pairs ** 0.5))))
    r = int(min(wl["rx"].shape[0], max(1, pairs // c)))
    rx_idx = np.linspace(0, wl["rx"].shape[0] - 1, r).astype(np.int64)
    return wl["cand"][:c], wl["rx"][rx_idx], f"first {c} candidates x {r} strided receivers of {wl['name']}"


def cpu_threads() -> int:
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it)."""
    from oracle import c_oracle as co

    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    co.set_num_threads(n)
    return co.num_threads()


def cpu_step(wl: dict, cand, rx, early_exit: int = 2):
    """One trace + validate on the host cores → (decided pairs, executed tests, seconds, valid).
    `early_exit=2`: a candidate stops at its first blocker — the same short cut the GPU arm takes,
    identical outputs; 0 = every (ray, triangle) pair evaluated like the reference's fori_loop."""
    from oracle import c_oracle as co

    t0 = time.perf_counter()
    _, _, mask, tests = co.trace_path_candidates(
        wl["vertices"], wl["triangles"], wl["tx"], rx, cand, early_exit=early_exit, count_tests=True
    )
    dt = time.perf_counter() - t0
    decided = int(wl["tx"].shape[0]) * int(rx.shape[0]) * int(cand.shape[0]) * (wl["order"] + 1) * int(wl["triangles"].shape[0])
    return decided, tests, dt, int(mask.sum())


def cpu_calibrate(wl: dict, seconds: float):
    """Pick a sample that takes about `seconds` on this host (early-exit evaluation)."""
    cand, rx, _ = cpu_sample(wl, 2e8)
    cpu_step(wl, cand[:8], rx[:1])  # load + thread pool warm-up
    decided, _, dt, _ = cpu_step(wl, cand, rx)
    rate = decided / max(dt, 1e-6)
    return cpu_sample(wl, rate * seconds)


def cpu_baseline_block(wl: dict, seconds: float) -> dict:
    """The `cpu_baseline` object: the oracle port on all host threads on a bounded sample of the
    workload — decided pairs/s with the early exit (comparable with the GPU arm's `value`), the
    executed rate, and the dense (no early exit, the reference's literal evaluation) rate on a
    quarter of the sample."""
    cores = cpu_threads()
    cand, rx, sample = cpu_calibrate(wl, seconds)
    decided, tests, dt, _ = cpu_step(wl, cand, rx)
    q = max(1, rx.shape[0] // 4)
    d_decided, d_tests, d_dt, _ = cpu_step(wl, cand, rx[:q], early_exit=0)
    return {
        "value": decided / dt, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "seconds": dt,
        "host_cpus": os.cpu_count(),
        "what": "oracle/oracle.c (C/OpenMP/AVX2 restatement of the reference algorithm; JAX/Warp not installable "
                "here), a candidate stops at its first blocker like the GPU arm: decided pairs per second",
        "executed_tests_per_s": tests / dt, "executed_fraction_of_algorithmic": tests / max(decided, 1),
        "dense_no_early_exit_tests_per_s": d_tests / d_dt,
    }


def run_reference(args, rank: int) -> None:
    """`--impl reference`: the reference's CPU algorithm (oracle port; JAX/Warp are not installable
    here — DESIGN.md) on all host threads, bounded sample per step, the same early exit and the same
    unit (decided pairs per second) as the GPU arm.  Rank 0 only.  Loads nothing of the product."""
    if rank != 0:
        return
    wl = build_workload(args.workload, 0, 1)
    cores = cpu_threads()
    # bounded sample per step, sized so that the whole --steps K --warmup W run stays near two minutes
    per_step = max(0.5, min(args.cpu_seconds, 120.0 / max(args.steps + args.warmup, 1)))
    cand, rx, sample = cpu_calibrate(wl, per_step)
    for _ in range(args.warmup):
        cpu_step(wl, cand, rx)
    decided = tests = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n, e, _, valid = cpu_step(wl, cand, rx)
        decided += n
        tests += e
    dt = time.perf_counter() - t0
    value = decided / dt
    q = max(1, rx.shape[0] // 4)
    _, d_tests, d_dt, _ = cpu_step(wl, cand, rx[:q], early_exit=0)
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(wl, 1, sample=sample),
        "executed_tests_per_s": tests / dt, "executed_fraction_of_algorithmic": tests / max(decided, 1),
        "dense_no_early_exit_tests_per_s": d_tests / d_dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(),
                         "what": "oracle/oracle.c, a candidate stops at its first blocker: decided pairs per second"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def launches_per_step(wl: dict, with_vjp: bool) -> int:
    """Kernels of OURS launched per steady-state step when profiles/inst_counts.json (the ncu-counted
    figure) has no entry for the workload: stage A, the culled traversal (meshes > 512 triangles; the
    ordering pass + cascade of resident passes below that), the stats marker, 3 compaction kernels, the
    VJP.  The mesh-only kernels (packs, sorts, hierarchy: ~14 launches) run once per mesh state."""
    tiles = -(-int(wl["triangles"].shape[0]) // 512)
    blockage = 2 if tiles > 1 else 1 + 3 + 2
    return 1 + blockage + 3 + (1 if with_vjp else 0)


def workload_config(wl: dict, world: int, **extra) -> dict:
    T = int(wl["triangles"].shape[0])
    pairs = int(wl["tx"].shape[0] * wl["rx_global"] * wl["cand_global"])
    cfg = {
        "workload": wl["name"], "triangles": T, "num_tx": int(wl["tx"].shape[0]),
        "num_rx": int(wl["rx_global"]), "order": int(wl["order"]),
        "candidates": int(wl["cand_global"]), "candidate_pairs": pairs,
        "algorithmic_tests_per_step": pairs * (wl["order"] + 1) * T,
        "blockage": "dense: every segment of every candidate is decided by the any-hit test, like the "
                    "reference; value counts rays x triangles decided, executed_tests_per_s the "
                    "Moller-Trumbore evaluations actually run (a candidate stops at its first blocker; "
                    "pairs the exact cull proves to be misses are skipped)",
        "parallelism": f"the workload's receivers dealt round-robin to {world} rank(s), every candidate on every "
                       "rank, one all-gather of valid paths",
        "l2_policy": "no flush: each step writes >1.3 GB of path vertices/objects (L2 is 126 MB); "
                     "the 0.5 MB packed mesh is L2/shared-memory resident by design",
    }
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


def parity_check(wl: dict, paths, max_pairs: int = 16384) -> dict:
    """Compare a strided sub-block of the LAST timed step's outputs (this rank's shard) with the C
    oracle: the validity mask, the vertices (bits) and the objects of `max_pairs` (rx, candidate) pairs."""
    from oracle import c_oracle as co

    import torch

    nrx, nc = wl["rx"].shape[0], wl["cand"].shape[0]
    n_r = max(1, min(nrx, 32))
    n_c = max(1, min(nc, max_pairs // n_r))
    sel_r = np.linspace(0, nrx - 1, n_r).astype(np.int64)
    sel_c = np.linspace(0, nc - 1, n_c).astype(np.int64)
    # plus the receivers / candidates of (up to 8) paths the GPU reports valid, so that the oracle
    # confirms valid paths too and not only rejections
    valid = paths.mask[0].nonzero()[:8].cpu().numpy()
    sel_r = np.unique(np.concatenate([sel_r, valid[:, 0]]))
    sel_c = np.unique(np.concatenate([sel_c, valid[:, 1]]))
    tx0 = wl["tx"][:1]  # first transmitter
    ev, eo, em = co.trace_path_candidates(wl["vertices"], wl["triangles"], tx0, wl["rx"][sel_r],
                                          wl["cand"][sel_c], early_exit=2)
    eo[..., -1] = sel_r[eo[..., -1]].astype(np.int32)  # the oracle numbered the sub-sampled receivers
    ir = torch.from_numpy(sel_r).to(paths.mask.device)
    ic = torch.from_numpy(sel_c).to(paths.mask.device)
    gm = paths.mask[:1].index_select(1, ir).index_select(2, ic).cpu().numpy()
    gv = paths.vertices.detach()[:1].index_select(1, ir).index_select(2, ic).cpu().numpy()
    go = paths.objects[:1].index_select(1, ir).index_select(2, ic).cpu().numpy()
    bad = int((gm != em).sum())
    bad_v = int((gv.view(np.uint32) != ev.view(np.uint32)).any(axis=(-1, -2)).sum())
    bad_o = int((go != eo).any(axis=-1).sum())
    return {"checked_pairs": int(em.size), "mismatches": bad + bad_v + bad_o, "mask_mismatches": bad,
            "vertex_bit_mismatches": bad_v, "object_mismatches": bad_o, "oracle_valid": int(em.sum()),
            "against": "oracle/oracle.c on a strided block of the last timed step (rank 0's shard)"}


def run_ours(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    import differt_b200 as drt
    from differt_b200 import _lib
    from differt_b200.distributed import GatherRecord, fill_record, gather_valid_paths

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # started first, seconds before anything is timed (see ClockSampler)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if world > 1:
        # a rank that dies or a fabric that never completes a collective ends the run with an error after 5
        # minutes instead of NCCL's default 10 (every leg between two collectives is seconds long)
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=5))

    if args.cand is not None:  # BASELINE config 5's sweep over the number of candidates
        WORKLOADS[args.workload]["cand"] = args.cand
    wl = build_workload(args.workload, rank, world)
    if args.emulate_shard:  # one process traces the shard rank r of w would get (no collective)
        r_, w_ = (int(x) for x in args.emulate_shard.split("/"))
        wl = build_workload(args.workload, r_, w_)
        wl["receivers"] = None
    k = wl["order"]
    capacity = 1 << 10  # valid paths per rank in the gather record (overflow → the API's retry path)

    # host buffers (pinned) for the e2e leg; device-resident copies for the kernel-level leg
    host = {n: torch.from_numpy(wl[n]).pin_memory() for n in ("vertices", "triangles", "tx", "rx", "cand")}
    mesh = drt.Mesh.from_numpy(wl["vertices"], wl["triangles"])
    tx_d, rx_d, cand_d = (host[n].to(dev) for n in ("tx", "rx", "cand"))
    # two records / receive buffers (preallocated): the all-gather of step i runs on NCCL's own stream
    # while step i + 1 computes, and is only waited for when its buffers are about to be reused
    records = [GatherRecord(capacity, k, dev) for _ in range(2)]
    record = records[0]
    gathered_bufs = [torch.empty(world * record.nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    in_flight: list = [None, None]
    step_no = [0]
    stats_acc = torch.zeros(4, dtype=torch.int64, device=dev)
    rx_shard = None if wl["receivers"] is None else (wl["receivers"][0], torch.from_numpy(wl["receivers"][1]).to(dev))

    # reverse mode every step (BASELINE config 3 is "with VJP"; SURVEY §8d: VJP of vertices.sum() w.r.t.
    # tx, rx and mesh.vertices): an all-ones cotangent, resident like the other inputs
    with_vjp = not args.no_vjp
    if with_vjp:
        mesh = drt.Mesh(mesh.vertices.requires_grad_(True), mesh.triangles)
        tx_d.requires_grad_(True)
        rx_d.requires_grad_(True)
        cot = torch.ones((tx_d.shape[0], rx_d.shape[0], cand_d.shape[0], k + 2, 3), dtype=torch.float32, device=dev)

    def step_resident(profile: bool):
        paths = drt.trace_path_candidates(
            mesh, tx_d, rx_d, cand_d, dense_blockage=True, _stats_accumulate=stats_acc, _profile=profile
        )
        if with_vjp:
            torch.autograd.grad(paths.vertices, (mesh.vertices, tx_d, rx_d), cot)
        slot = step_no[0] & 1
        step_no[0] += 1
        if in_flight[slot] is not None:  # the gather that used these buffers two steps ago
            in_flight[slot].wait()
        fill_record(records[slot], paths, wl["cand_global"], wl["cand_start"], receivers=rx_shard)
        if world > 1:
            in_flight[slot] = dist.all_gather_into_tensor(gathered_bufs[slot], records[slot].buffer, async_op=True)
        return paths

    def drain():
        """Every gather issued so far has completed on the current stream."""
        for w_ in in_flight:
            if w_ is not None:
                w_.wait()

    mask_host = torch.empty((tx_d.shape[0], rx_d.shape[0], cand_d.shape[0]), dtype=torch.bool, pin_memory=True)
    grad_host: list = []

    def step_e2e():
        """Host buffers in, host results out, through the public API."""
        m = drt.Mesh(host["vertices"].to(dev, non_blocking=True).requires_grad_(with_vjp),
                     host["triangles"].to(dev, non_blocking=True))
        tx_e = host["tx"].to(dev, non_blocking=True).requires_grad_(with_vjp)
        rx_e = host["rx"].to(dev, non_blocking=True).requires_grad_(with_vjp)
        paths = drt.trace_path_candidates(
            m, tx_e, rx_e, host["cand"], dense_blockage=True, _stats_accumulate=stats_acc
        )
        grads = ()
        if with_vjp:
            grads = torch.autograd.grad(paths.vertices, (m.vertices, tx_e, rx_e), cot)
        fill_record(records[0], paths, wl["cand_global"], wl["cand_start"], receivers=rx_shard)
        # device → host: the mask and the gradients into pinned buffers (asynchronous), then the gather,
        # whose count read is the synchronisation point
        mask_host.copy_(paths.mask, non_blocking=True)
        grads_h = []
        for i, g in enumerate(grads):
            if i >= len(grad_host):
                grad_host.append(torch.empty(g.shape, dtype=g.dtype, pin_memory=True))
            grad_host[i].copy_(g, non_blocking=True)
            grads_h.append(grad_host[i])
        valid = gather_valid_paths(records[0])  # all-gather (N>1) + counts to host
        assert valid is not None, "gather record overflow: raise `capacity` in bench.py"
        out = (valid.index.cpu(), valid.vertices.cpu(), valid.objects.cpu(), mask_host, *grads_h)
        torch.cuda.current_stream().synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        r = None
        for i in range(n):
            r = fn(i)
        drain()  # the last gathers are part of the work
        e1.record()
        barrier()
        return e0.elapsed_time(e1), r

    if args.profile_only:
        # for `ncu`: exactly K steps of the timed loop's body and nothing else, so that per-step kernel
        # and instruction counts are the capture's totals divided by K (tools/inst_counts.py)
        keep = None
        for _ in range(args.steps):
            keep = step_resident(False)
        barrier()
        if rank == 0:
            emit({"profile_only": True, "steps": args.steps, "workload": wl["name"], "n_gpus": world,
                  "valid_paths": int(keep.mask.sum().item())})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- kernel-level leg: inputs resident in HBM ------------------------------------------------
    if sampler:
        sampler.wait_ready()
    # The results are bound to `paths` exactly like in the timed loop: the previous step's outputs stay
    # alive while the next step allocates, so the caching allocator reaches its two-generation steady
    # state HERE and the timed region never calls cudaMalloc (it would, once, in its second step).
    paths = None
    for _ in range(6):  # set-up: allocator steady state and cold-start effects of a fresh box, before the
        paths = step_resident(True)  # W warm-ups (same flags as the timed steps; the ring is reset below)
    for _ in range(args.warmup):
        paths = step_resident(True)
    barrier()
    stats_acc.zero_()
    _lib.check(_lib.lib.drt_profile_reset())
    paths = None  # timed() binds its own results: with this one still alive a THIRD generation of the 1.4 GB
    #               outputs would be allocated (cudaMalloc, a 0.25 s host stall) inside the timed region
    if sampler:
        sampler.mark_start()
    diag = os.environ.get("DRT_BENCH_DIAG")
    if diag:
        s0 = torch.cuda.memory_stats()
        host_ms = []

        def probe(i):
            h0 = time.perf_counter()
            r = step_resident(i < 64)
            host_ms.append(1e3 * (time.perf_counter() - h0))
            return r

        ms, paths = timed(probe, args.steps)
        s1 = torch.cuda.memory_stats()
        sys.stderr.write(f"[diag] ms/step {ms / args.steps:.3f} host enqueue ms {host_ms} cudaMalloc "
                         f"{s1['num_device_alloc'] - s0['num_device_alloc']} cudaFree "
                         f"{s1['num_device_free'] - s0['num_device_free']}\n")
    else:
        ms, paths = timed(lambda i: step_resident(i < 64), args.steps)
    clocks = sampler.stop() if sampler else None
    tests_local = int(stats_acc[0].item())
    valid_local = int(paths.mask.sum().item())
    import ctypes as C

    kern_ms = []
    for slot in range(_lib.lib.drt_profile_count()):
        f = C.c_float()
        _lib.check(_lib.lib.drt_profile_elapsed_ms(slot, C.byref(f)))
        kern_ms.append(f.value)

    # ---- parity: the last timed step against the oracle; the gathered records against the ranks' masks --
    parity = parity_check(wl, paths) if rank == 0 and not args.no_cpu else None
    torch.cuda.synchronize()
    last = (step_no[0] - 1) & 1
    gathered, record = gathered_bufs[last], records[last]
    parts = [GatherRecord.views(gathered[r * record.nbytes:(r + 1) * record.nbytes], capacity, k) for r in range(world)] \
        if world > 1 else [record.fields()]
    counts = [int(p_[0].item()) for p_ in parts]
    gather_ok = counts[rank] == valid_local and all(c <= capacity for c in counts)
    if gather_ok and valid_local > 0:  # this rank's own record inside the gathered buffer: right paths, right bits
        mine = parts[rank]
        from differt_b200.distributed import global_path_index_receivers

        idx_local = paths.mask.reshape(-1).nonzero().squeeze(-1)
        idx_global = global_path_index_receivers(idx_local, cand_d.shape[0], rx_shard[1], wl["rx_global"]) \
            if rx_shard is not None else idx_local
        gather_ok = bool(torch.equal(mine[1][:valid_local], idx_global)) and bool(
            torch.equal(mine[2][:valid_local], paths.vertices.detach().reshape(-1, k + 2, 3)[idx_local]))
    del paths

    # ---- sustained: the same step for at least 5 s (the K timed steps above can be a short burst) ---------
    n_sus = max(args.steps, int(np.ceil(args.sustain_seconds * 1e3 / max(ms / args.steps, 1e-3)))) if args.sustain_seconds > 0 else 0
    ms_sus = None
    if n_sus:
        if world > 1:  # every rank must run the same number of steps
            t = torch.tensor([n_sus], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_sus = int(t.item())
        ms_sus, _ = timed(lambda i: step_resident(False), n_sus)

    # ---- the API's default (pruned) mode: blockage only for candidates that pass the cheap tests -----
    for _ in range(2):
        drt.trace_path_candidates(mesh, tx_d, rx_d, cand_d)
    ms_pruned, _ = timed(lambda i: drt.trace_path_candidates(mesh, tx_d, rx_d, cand_d), 10)
    ms_pruned /= 10

    # ---- end-to-end leg: host buffers through the public API, as many steps as the timed region ---------
    e2e_steps = args.steps if args.e2e_steps is None else args.e2e_steps
    step_e2e()
    barrier()
    stats_acc.zero_()
    ms_e2e, out = timed(lambda i: step_e2e(), e2e_steps)
    tests_e2e_local = int(stats_acc[0].item())
    h2d = sum(int(h.numel() * h.element_size()) for h in host.values())
    d2h = sum(int(o.numel() * o.element_size()) for o in out) + 8 * world
    num_valid_global = int(out[0].shape[0])

    # ---- weak-scaling leg (N > 1): every rank traces a full-size shard of its own ------------------------
    weak = None
    if world > 1 and args.weak_steps > 0:
        wlw = build_workload(args.workload, rank, world, weak=True)
        cand_w = torch.from_numpy(wlw["cand"]).to(dev)
        rx_w = torch.from_numpy(wlw["rx"]).to(dev).requires_grad_(with_vjp)  # ALL receivers on every rank
        cot_w = torch.ones((tx_d.shape[0], rx_w.shape[0], cand_w.shape[0], k + 2, 3), dtype=torch.float32, device=dev) \
            if with_vjp else None

        def step_weak(_i):
            p_ = drt.trace_path_candidates(mesh, tx_d, rx_w, cand_w, dense_blockage=True)
            if with_vjp:
                torch.autograd.grad(p_.vertices, (mesh.vertices, tx_d, rx_w), cot_w)
            fill_record(records[0], p_, wlw["cand_global"], wlw["cand_start"])
            dist.all_gather_into_tensor(gathered_bufs[0], records[0].buffer)
            return p_

        keep = [step_weak(0), step_weak(1)]
        ms_weak, _ = timed(step_weak, args.weak_steps)
        del keep
        weak = ms_weak

    # ---- reduce over ranks: max time, summed work ----------------------------------------------------
    kms_local = float(np.mean(kern_ms)) if kern_ms else 0.0
    red = torch.tensor([ms, ms_e2e, ms_sus or 0.0, weak or 0.0, kms_local, -kms_local], dtype=torch.float64, device=dev)
    tot = torch.tensor([tests_local, tests_e2e_local, valid_local, int(gather_ok)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e, ms_sus, ms_weak, kms_max, kms_min_neg = red.tolist()
    tests, tests_e2e, valid_total, gather_ok_ranks = tot.tolist()

    if rank == 0:
        cfg = workload_config(wl, world)
        algo_step = cfg["algorithmic_tests_per_step"]
        pairs_step = cfg["candidate_pairs"]
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        kms = float(np.mean(kern_ms)) if kern_ms else None
        tests_per_launch = tests_local / max(args.steps, 1)
        # what bounds the blockage kernels: instruction issue (no tensor cores, operands on chip).  The
        # warp-instruction count per step comes from the committed ncu launch list of this very command
        # (profiles/inst_counts.json, written by tools/inst_counts.py), never from a literal.
        inst = None
        inst_file = ROOT / "profiles" / "inst_counts.json"
        if inst_file.exists():
            inst = json.loads(inst_file.read_text()).get(f"{wl['name']}@{world}")
        issue_peak = 148 * 4 * sm_mhz * 1e6
        achieved = inst["blockage_warp_instructions_per_step"] / (kms * 1e-3) if inst and kms else None
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        hbm_peak = float(json.loads(peaks_file.read_text())["hbm_gbs"]) if peaks_file.exists() else 6650.0
        # DRAM bytes of the blockage kernels per step, from the same ncu launch list as the instruction counts
        traffic = (inst or {}).get("blockage_dram_bytes_per_step") or None
        roofline = {
            "kernel": "blockage pass = drt::path_walk_kernel<order+1> (one warp per candidate: head test against the "
                      "largest triangles, then the exact culled traversal of the 8-ary hierarchy, csrc/walk.cuh); on meshes "
                      "of <= 512 triangles drt::path_head_kernel + drt::hit_count_kernel (resident all-pairs passes)",
            "bound": "fp32_issue", "achieved": achieved, "peak": issue_peak, "unit": "warp-inst/s",
            "frac": achieved / issue_peak if achieved else None, "traffic": traffic,
            "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
            "instructions_source": (inst or {}).get("source", "profiles/inst_counts.json has no entry for this "
                                                              "workload: run tools/inst_counts.py on the ncu launch list"),
            "warp_instructions_per_executed_test": (inst["blockage_warp_instructions_per_step"] * 32 / tests_per_launch
                                                     if inst and tests_per_launch else None),
            "kernel_ms": kms, "kernel_share_of_step": kms * args.steps / ms if kms else None,
            "kernel_ms_over_ranks": {"max": kms_max, "min": -kms_min_neg,
                                     "note": "blockage kernel time of the slowest / fastest rank: the shards hold "
                                             "different candidates, the step ends with the slowest"},
            "executed_tests_per_launch": tests_per_launch,
            "hbm_model": {
                "note": "INAPPLICABLE as a roofline (kept because BASELINE.json asks for %HBM): 36 B of triangle "
                        "operand per executed test against the measured copy bandwidth; the packed mesh is resident "
                        "on chip, so the real DRAM traffic (`traffic`, ncu) is the path vertices only",
                "bytes_per_test": BYTES_PER_TEST,
                "model_gbs": BYTES_PER_TEST * tests_per_launch / (kms * 1e-3) / 1e9 if kms else None,
                "hbm_peak_gbs": hbm_peak,
                "dram_gbs_measured": traffic / (ms / args.steps * 1e-3) / 1e9 if traffic else None,
            },
        }
        line = {
            "metric": METRIC, "value": algo_step * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(
                wl, world,
                reverse_mode=("every step also runs the VJP of vertices.sum() w.r.t. tx, rx and mesh.vertices "
                              "(all-ones cotangent); not counted in value's tests") if with_vjp else "off"),
            "candidate_pairs_per_s": pairs_step * args.steps / (ms * 1e-3),
            "valid_paths_per_s": valid_total * args.steps / (ms * 1e-3),
            "valid_paths_per_step": valid_total,
            "executed_tests_per_s": tests / (ms * 1e-3),
            "executed_fraction_of_algorithmic": tests / max(args.steps * algo_step, 1),
            "sustained": {"steps": n_sus, "ms_per_step": ms_sus / n_sus, "seconds": ms_sus * 1e-3,
                          "value": algo_step * n_sus / (ms_sus * 1e-3)} if n_sus else None,
            "default_mode": {"what": "trace_path_candidates() as the API runs it by default: identical outputs, "
                                     "blockage only for candidates that pass the cheap tests (rank 0's shard, "
                                     "not part of value)",
                             "ms_per_step": ms_pruned,
                             "candidate_pairs_per_s": wl["tx"].shape[0] * wl["rx"].shape[0] * wl["cand"].shape[0]
                             / (ms_pruned * 1e-3)},
            "shard": {"receivers_per_rank": int(wl["rx"].shape[0]), "candidates_per_rank": int(wl["cand"].shape[0])},
            "e2e": {"value": algo_step * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "executed_tests_per_s": tests_e2e / (ms_e2e * 1e-3), "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "valid_paths_gathered": num_valid_global},
            "parity": parity,
            "gather_check": {"ranks_ok": gather_ok_ranks, "ranks": world, "valid_paths_per_rank": counts,
                             "what": "after the timed loop every rank reads the all-gathered buffer back: the "
                                     "counts match its mask and its own record holds its valid paths bit for bit"},
            "gpu_launches": args.steps * ((inst or {}).get("our_launches_per_step") or launches_per_step(wl, with_vjp)),
            "clocks": clocks, "roofline": roofline,
        }
        if world > 1 and ms_weak:
            wcfg = workload_config(build_workload(args.workload, 0, world, weak=True), world)
            line["weak_scaling"] = {"what": "extra leg: every rank traces a full-size shard of its own "
                                            f"({wl['cand_global']} candidates per GPU)",
                                    "steps": args.weak_steps, "ms_per_step": ms_weak / args.weak_steps,
                                    "value": wcfg["algorithmic_tests_per_step"] * args.weak_steps / (ms_weak * 1e-3)}
        if world == 1 and not args.no_cpu:
            cb = cpu_baseline_block(wl, args.cpu_seconds)
            line["cpu_baseline"] = cb
            line["vs_cpu"] = {
                "e2e_decided_pairs_ratio": line["e2e"]["value"] / cb["value"],
                "executed_tests_ratio": line["executed_tests_per_s"] / cb["executed_tests_per_s"],
                "note": "both arms stop a candidate at its first blocker; the first ratio is a ratio of times "
                        "for the same job, the second compares Moller-Trumbore evaluations per second",
            }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout() -> None:
    """Keep stdout for the ONE JSON line: everything else that writes to file descriptor 1 (NCCL's
    version / debug lines, library banners) is sent to stderr instead."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main() -> None:
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per bounded sample")
    ap.add_argument("--e2e-steps", type=int, default=None, help="default: as many as --steps")
    ap.add_argument("--sustain-seconds", type=float, default=5.0,
                    help="extra leg: repeat the step for at least this long (0 = skip)")
    ap.add_argument("--weak-steps", type=int, default=5, help="extra weak-scaling leg for N > 1 (0 = skip)")
    ap.add_argument("--cand", type=int, default=None, help="override the workload's number of candidates")
    ap.add_argument("--emulate-shard", default=None, help="r/w: single process, the shard of rank r of w (diagnostic)")
    ap.add_argument("--profile-only", action="store_true",
                    help="run exactly --steps steps of the timed loop's body and nothing else (for ncu)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-vjp", action="store_true", help="forward only (default: forward + VJP every step)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(
            f"--gpus {args.gpus} needs one process per GPU: launch with "
            f"python -m torch.distributed.run --nnodes=1 --nproc-per-node {args.gpus} "
            "--master-addr 127.0.0.1 --master-port 29500 bench.py ..."
        )
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
