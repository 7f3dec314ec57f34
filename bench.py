#!/usr/bin/env python
"""bench.py — throughput of the DiffeRT geometric hot path (trace + validate path candidates).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path over one batch of synthetic input: trace-and-validate
(`_trace_path_candidates`, reference differt/src/differt/geometry/_solvers.py:499-770) of every
(tx, rx, candidate) of the workload, with the blockage test evaluated for every candidate like the
reference does ("dense"), the reverse mode of `vertices.sum()` w.r.t. tx, rx and the mesh vertices
(BASELINE config 3 is "with VJP"), followed by the compaction of the valid paths (`TracedPaths.masked()`)
and — for N > 1, where every rank traces its own shard of the candidates — ONE all-gather of the
survivors.  The metric is BASELINE.json's: ray–triangle tests per second, counted as SURVEY.md
§8(d) defines it — every (ray, triangle) pair the step DECIDES, rays x triangles ("algorithmic",
what the reference's dense evaluation executes) — so that `value` is whole-job throughput in the
same unit for both arms.  The Möller–Trumbore evaluations the kernel actually executed (device
counter; fewer, because an any-hit query stops at the first blocking tile) are reported next to it
as `executed_tests_per_s`, and the roofline is computed from the EXECUTED count only.

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
    # BASELINE.json configs[3] per GPU: 16 TX x 4096 RX, order 3, the 4096 candidates sharded over 8 GPUs
    "urban10k_16tx_4096rx_order3": dict(scene=("urban", 29, 29), rx=(64, 64), order=3, cand=512, ntx=16),
    # BASELINE.json configs[4] per GPU: 50k-triangle mesh, 1 TX x 16 384 RX, order 4, 2048 candidates
    "urban50k_1tx_16384rx_order4": dict(scene=("urban", 64, 65), rx=(128, 128), order=4, cand=2048),
    # small variant for quick checks (not a bench line)
    "urban10k_small": dict(scene=("urban", 29, 29), rx=(16, 16), order=3, cand=1024),
}
DEFAULT_WORKLOAD = "urban10k_1tx_4096rx_order3"


def build_workload(name: str, rank: int, world: int):
    """Seeded synthetic inputs (host, NumPy).  Every rank gets its own `cand` candidates (weak
    scaling): candidates rank*cand .. (rank+1)*cand of a global list of world*cand."""
    from differt_b200 import scenes

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
    cand_all = scenes.sampled_candidates(t.shape[0], w["order"], w["cand"] * world, seed=1234)
    start = rank * w["cand"]
    cand = np.ascontiguousarray(cand_all[start:start + w["cand"]])
    # candidates known to be valid for some receivers (found by tools/find_valid_candidates.py and
    # re-validated by the CPU oracle in tests/) lead every shard, so that the step produces real paths
    fixture = ROOT / "tests" / "golden" / "urban10k_valid_candidates.npz"
    if w["scene"] == ("urban", 29, 29) and fixture.exists():
        known = np.load(fixture)[f"order{w['order']}"]
        n = min(known.shape[0], cand.shape[0] // 4)
        cand[:n] = known[:n]
    return dict(
        name=name, vertices=v, triangles=t, tx=tx, rx=rx, order=w["order"], cand=cand,
        cand_global=w["cand"] * world, cand_start=start,
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


def cpu_step(wl: dict, cand, rx):
    """One dense (no early exit) trace + validate on the host cores → (tests, seconds, valid)."""
    from oracle import c_oracle as co

    t0 = time.perf_counter()
    _, _, mask, tests = co.trace_path_candidates(
        wl["vertices"], wl["triangles"], wl["tx"], rx, cand, early_exit=False, count_tests=True
    )
    return tests, time.perf_counter() - t0, int(mask.sum())


def cpu_calibrate(wl: dict, seconds: float):
    """Pick a sample that takes about `seconds` on this host."""
    cand, rx, _ = cpu_sample(wl, 2e8)
    cpu_step(wl, cand[:8], rx[:1])  # load + thread pool warm-up
    tests, dt, _ = cpu_step(wl, cand, rx)
    rate = tests / max(dt, 1e-6)
    return cpu_sample(wl, rate * seconds)


def run_reference(args, rank: int) -> None:
    """`--impl reference`: the reference's CPU algorithm (oracle port; JAX/Warp are not installable
    here — DESIGN.md) on all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import c_oracle as co

    wl = build_workload(args.workload, 0, 1)
    cores = cpu_threads()
    # bounded sample per step, sized so that the whole --steps K --warmup W run stays near two minutes
    per_step = max(0.5, min(args.cpu_seconds, 120.0 / max(args.steps + args.warmup, 1)))
    cand, rx, sample = cpu_calibrate(wl, per_step)
    for _ in range(args.warmup):
        cpu_step(wl, cand, rx)
    tests = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n, _, valid = cpu_step(wl, cand, rx)
        tests += n
    dt = time.perf_counter() - t0
    value = tests / dt
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(wl, 1, sample=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def launches_per_step(wl: dict, with_vjp: bool) -> int:
    tiles = -(-int(wl["triangles"].shape[0]) // 512)
    rounds = min(4, tiles - 1)
    return 1 + 2 + 1 + (3 + 1 + 4 * rounds) + -(-tiles // 8) + 3 + (1 if with_vjp else 0)


def workload_config(wl: dict, world: int, **extra) -> dict:
    T = int(wl["triangles"].shape[0])
    pairs = int(wl["tx"].shape[0] * wl["rx"].shape[0] * wl["cand"].shape[0])
    cfg = {
        "workload": wl["name"], "triangles": T, "num_tx": int(wl["tx"].shape[0]),
        "num_rx": int(wl["rx"].shape[0]), "order": int(wl["order"]),
        "candidates_per_gpu": int(wl["cand"].shape[0]), "candidate_pairs_per_gpu": pairs,
        "algorithmic_tests_per_gpu_step": pairs * (wl["order"] + 1) * T,
        "blockage": "dense: every segment of every candidate is submitted to the any-hit kernel, like "
                    "the reference; value counts rays x triangles decided, executed_tests_per_s the "
                    "Moller-Trumbore evaluations actually run",
        "parallelism": f"candidate shards x{world}, one all-gather of valid paths",
        "l2_policy": "no flush: each step writes >1.3 GB of path vertices/objects (L2 is 126 MB); "
                     "the 0.5 MB packed mesh is L2/shared-memory resident by design",
    }
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


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
        dist.init_process_group("nccl", device_id=dev)

    wl = build_workload(args.workload, rank, world)
    k = wl["order"]
    capacity = 1 << 14

    # host buffers (pinned) for the e2e leg; device-resident copies for the kernel-level leg
    host = {n: torch.from_numpy(wl[n]).pin_memory() for n in ("vertices", "triangles", "tx", "rx", "cand")}
    mesh = drt.Mesh.from_numpy(wl["vertices"], wl["triangles"])
    tx_d, rx_d, cand_d = (host[n].to(dev) for n in ("tx", "rx", "cand"))
    record = GatherRecord(capacity, k, dev)
    stats_acc = torch.zeros(4, dtype=torch.int64, device=dev)

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
        fill_record(record, paths, wl["cand_global"], wl["cand_start"])
        if world > 1:
            gathered = torch.empty(world * record.nbytes, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered, record.buffer)
        return paths

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
        fill_record(record, paths, wl["cand_global"], wl["cand_start"])
        # device → host: the mask and the gradients into pinned buffers (asynchronous), then the gather,
        # whose count read is the synchronisation point
        mask_host.copy_(paths.mask, non_blocking=True)
        grads_h = []
        for i, g in enumerate(grads):
            if i >= len(grad_host):
                grad_host.append(torch.empty(g.shape, dtype=g.dtype, pin_memory=True))
            grad_host[i].copy_(g, non_blocking=True)
            grads_h.append(grad_host[i])
        valid = gather_valid_paths(record)  # all-gather (N>1) + counts to host
        out = (valid.index.cpu(), valid.vertices.cpu(), valid.objects.cpu(), mask_host, *grads_h)
        torch.cuda.current_stream().synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_start()
    ev0.record()
    for i in range(args.steps):
        paths = step_resident(i < 64)
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    tests_local = int(stats_acc[0].item())
    valid_local = int(paths.mask.sum().item())
    import ctypes as C

    kern_ms = []
    for slot in range(_lib.lib.drt_profile_count()):
        f = C.c_float()
        _lib.check(_lib.lib.drt_profile_elapsed_ms(slot, C.byref(f)))
        kern_ms.append(f.value)
    del paths

    # ---- the API's default (pruned) mode: blockage only for candidates that pass the cheap tests -----
    for _ in range(2):
        drt.trace_path_candidates(mesh, tx_d, rx_d, cand_d)
    evp0, evp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    evp0.record()
    for _ in range(10):
        drt.trace_path_candidates(mesh, tx_d, rx_d, cand_d)
    evp1.record()
    barrier()
    ms_pruned = evp0.elapsed_time(evp1) / 10

    # ---- end-to-end leg: host buffers through the public API ---------------------------------------
    e2e_steps = max(1, min(args.steps, 3)) if args.e2e_steps is None else args.e2e_steps
    step_e2e()
    barrier()
    stats_acc.zero_()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(e2e_steps):
        out = step_e2e()
    ev3.record()
    barrier()
    ms_e2e = ev2.elapsed_time(ev3)
    tests_e2e_local = int(stats_acc[0].item())
    h2d = sum(int(h.numel() * h.element_size()) for h in host.values())
    d2h = sum(int(o.numel() * o.element_size()) for o in out) + 8 * world
    num_valid_global = int(out[0].shape[0])

    # ---- reduce over ranks: max time, summed work ----------------------------------------------------
    red = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    tot = torch.tensor([tests_local, tests_e2e_local, valid_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e = red.tolist()
    tests, tests_e2e, valid_total = tot.tolist()

    if rank == 0:
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        kms = float(np.mean(kern_ms)) if kern_ms else None
        tests_per_launch = tests_local / max(args.steps, 1)
        achieved = BYTES_PER_TEST * tests_per_launch / (kms * 1e-3) / 1e9 if kms else None
        roofline = {
            "kernel": "blockage pass = drt::path_head_kernel<order+1> (head tiles, ~3/4 of the time) + "
                      "drt::intersect_kernel<order+1, ANY, PATH> (ring pass over the survivors)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if achieved else None, "traffic": None,
            "peak_source": peak_src, "bytes_per_test": BYTES_PER_TEST,
            "tests_per_launch": tests_per_launch, "kernel_ms": kms,
            "kernel_share_of_step": kms * args.steps / ms if kms else None,
            # what actually bounds the kernel: issue slots; 63 = SASS instructions per
            # test (cuobjdump of the inner loop), peak = 148 SMs x 4 schedulers x SM clock
            "fp32_issue": {
                "sass_instructions_per_test": 63,
                "achieved_warp_instructions_per_s": 63 * tests_per_launch / 32 / (kms * 1e-3) if kms else None,
                "peak_warp_instructions_per_s": 148 * 4 * (clocks["sm_mhz"] or 1965.0) * 1e6 if clocks else None,
            },
            "note": "streamed-operand model (36 B per executed test); the packed mesh is on-chip "
                    "resident so DRAM traffic is far below it and the binding limit is FP32 issue — "
                    "see DESIGN.md and profiles/",
        }
        fi = roofline["fp32_issue"]
        if fi["achieved_warp_instructions_per_s"] and fi["peak_warp_instructions_per_s"]:
            fi["frac"] = fi["achieved_warp_instructions_per_s"] / fi["peak_warp_instructions_per_s"]
        traffic_file = ROOT / "profiles" / "traffic.json"
        if traffic_file.exists():
            roofline["traffic"] = json.loads(traffic_file.read_text()).get("dram_bytes_per_launch")
        algo_step = world * workload_config(wl, world)["algorithmic_tests_per_gpu_step"]
        line = {
            "metric": METRIC, "value": algo_step * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(
                wl, world,
                reverse_mode=("every step also runs the VJP of vertices.sum() w.r.t. tx, rx and mesh.vertices "
                              "(all-ones cotangent, 1 ms); not counted in value's tests") if with_vjp else "off"),
            "candidate_pairs_per_s": world * wl["tx"].shape[0] * wl["rx"].shape[0] * wl["cand"].shape[0]
            * args.steps / (ms * 1e-3),
            "valid_paths_per_s": valid_total * args.steps / (ms * 1e-3),
            "default_mode": {"what": "trace_path_candidates() as the API runs it by default: identical outputs, "
                                     "blockage only for candidates that pass the cheap tests (rank 0, not "
                                     "part of value)",
                             "ms_per_step": ms_pruned,
                             "candidate_pairs_per_s": wl["tx"].shape[0] * wl["rx"].shape[0] * wl["cand"].shape[0]
                             / (ms_pruned * 1e-3)},
            "valid_paths_per_step": valid_total,
            "executed_tests_per_s": tests / (ms * 1e-3),
            "executed_fraction_of_algorithmic": tests / max(args.steps * algo_step, 1),
            "e2e": {"value": algo_step * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "executed_tests_per_s": tests_e2e / (ms_e2e * 1e-3), "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "valid_paths_gathered": num_valid_global},
            # our kernels per step: pack, area keys + gather, stage A, ordering pass (hit count, iota,
            # gather, sample list, then per greedy round: resident pass on the samples, hit count, iota,
            # gather), ceil(tiles / 8) resident passes of the cascade, 3 compaction kernels, the VJP
            # (CUB's radix-sort kernels are not counted as ours)
            "gpu_launches": args.steps * launches_per_step(wl, with_vjp),
            "clocks": clocks, "roofline": roofline,
        }
        if world == 1 and not args.no_cpu:
            cores = cpu_threads()
            cand, rx, sample = cpu_calibrate(wl, args.cpu_seconds)
            n, dt, _ = cpu_step(wl, cand, rx)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample, "seconds": dt, "host_cpus": os.cpu_count()}
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
    ap.add_argument("--e2e-steps", type=int, default=None)
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
