"""Regenerate BASELINE.md §5-§6 (measured results, per-kernel table) from the JSON lines under profiles/.
Read-only for everything except BASELINE.md."""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
P = ROOT / "profiles"


def load(name):
    f = P / name
    return json.loads(f.read_text()) if f.exists() else None


def fmt(x):
    m, e = f"{x:.2e}".split("e")
    return f"{m}·10^{int(e)}"


def main():
    b = (ROOT / "BASELINE.md").read_text()
    b = b[: b.index("## 5. Measured results")]
    d = load("r2_bench_final2_n1.json")
    ref = load("r2_bench_final2_reference_arm.json")
    strong = {n: load(f"r2_bench_strong_n{n}.json") for n in (1, 2, 4, 8)}
    rows = []
    for n, s in strong.items():
        if s is None:
            continue
        eff = s["value"] / strong[1]["value"] / n
        k = s["roofline"]["kernel_ms_over_ranks"] if "kernel_ms_over_ranks" in s["roofline"] else None
        weak = (s.get("weak_scaling") or {}).get("ms_per_step")
        rows.append(f"| differt_b200 at commit b93abce (traversal before the in-plane / head-test work: 29 ms at N = 1), strong scaling (fixed config 3, receivers dealt block-cyclically) | {n} | {s['ms_per_step']:.2f} ms | "
                    f"{fmt(s['value'])} | {fmt(s['executed_tests_per_s'])} | {fmt(s['candidate_pairs_per_s'])} | {eff:.3f} | "
                    f"{(('%.2f / %.2f ms' % (k['max'], k['min'])) if k else '—')} | {('%.1f ms' % weak) if weak else '—'} |")
    cb = d["cpu_baseline"]
    sec5 = f"""## 5. Measured results (round 2, NVIDIA B200 @ 1965 MHz, `profiles/r2_*`)

Workload = config 3 (urban grid 10 094 triangles, 1 TX × 4096 RX, order 3, 4096 candidates, every candidate
blockage-decided like the reference, forward + VJP + compaction): 1.68·10^7 candidate-pairs, 6.7·10^7 rays,
6.77·10^11 (ray, triangle) pairs per step.  `python bench.py` / `torchrun ... bench.py --gpus N`; the problem is
the same for every N (strong scaling).  `value` counts the pairs DECIDED (SURVEY §8d) in BOTH arms — both stop a
candidate at its first blocker; "executed" counts the Möller–Trumbore evaluations actually run.

| arm | GPUs | step | pairs decided /s (`value`) | executed tests /s | candidate-pairs /s | efficiency | blockage kernel, slowest / fastest rank | weak-scaling leg |
|---|---|---|---|---|---|---|---|---|
| **differt_b200, final code** (`r2_bench_final2_n1.json`) | 1 | **{d['ms_per_step']:.2f} ms** | {fmt(d['value'])} | {fmt(d['executed_tests_per_s'])} | {fmt(d['candidate_pairs_per_s'])} | — | {d['roofline']['kernel_ms']:.2f} ms | — |
{chr(10).join(rows)}
| differt_b200, end to end from host buffers (`e2e`) | 1 | {d['e2e']['ms_per_step']:.2f} ms | {fmt(d['e2e']['value'])} | {fmt(d['e2e']['executed_tests_per_s'])} | — | — | — | — |
| differt_b200, API default mode (blockage only for candidates passing the cheap tests; identical outputs) | 1 | {d['default_mode']['ms_per_step']:.2f} ms | — | — | {fmt(d['default_mode']['candidate_pairs_per_s'])} | — | — | — |
| reference algorithm restated on CPU (`--impl reference`: C/OpenMP/AVX2 port, early exit like the GPU arm; JAX/Warp not installable), {ref['cpu_baseline']['cores']} host cores | 0 | — | {fmt(ref['value'])} | {fmt(ref['executed_tests_per_s'])} | — | — | — | — |
| same port, dense (no early exit: the reference's literal `fori_loop`) | 0 | — | {fmt(ref['dense_no_early_exit_tests_per_s'])} | {fmt(ref['dense_no_early_exit_tests_per_s'])} | — | — | — | — |

Ratios (same box, same run): `e2e` ÷ reference arm = **{d['e2e']['value'] / ref['value']:.0f}×** on decided pairs (a ratio of times
for the same job); executed-vs-executed Möller–Trumbore rate {d['vs_cpu']['executed_tests_ratio']:.0f}× (the GPU arm executes
{d['executed_fraction_of_algorithmic'] * 100:.2f} % of the pairs, the CPU arm {cb['executed_fraction_of_algorithmic'] * 100:.0f} %: the traversal replaces
tests by node tests).  Parity of the timed step: {d['parity']['checked_pairs']} pairs re-checked by the oracle, {d['parity']['mismatches']} mismatches.
Roofline of the blockage kernel: instruction issue, {d['roofline']['frac']:.2f} of 148 × 4 × 1.965 GHz ({d['roofline']['achieved']:.3g} warp
instructions/s, ncu-counted); DRAM {d['roofline']['traffic'] / 1e9:.2f} GB per step (the path vertices) = {d['roofline']['hbm_model']['dram_gbs_measured']:.0f} GB/s.
BASELINE.json's HBM accounting (36 B per executed test) gives {d['roofline']['hbm_model']['model_gbs']:.0f} GB/s-equivalent = {d['roofline']['hbm_model']['model_gbs'] / d['roofline']['hbm_model']['hbm_peak_gbs']:.2f} of
the measured {d['roofline']['hbm_model']['hbm_peak_gbs']:.0f} GB/s — inapplicable as a roofline (the operand never leaves the chip), reported as
`roofline.hbm_model`.

The multi-GPU rows were measured earlier in the round and NOT repeated on the final kernel: the one
`gpurun --gpus 2` call made for it (two-device pytest + `torchrun` N = 2) ran into the call's time limit without
output and used up the round's GPU budget; the sharding code did not change in between (candidate / receiver
shards + one all-gather, `differt_b200/distributed.py`), every rank runs the same kernels on its shard, and the
driver's own 1 → 8 run at round end is the measurement of record.
"""
    others = []
    for label, f1, f8 in (
        ("config 2: street canyon 986 triangles, 1 × 256 RX, order 2, one 65 536-candidate chunk", "r2_bench_cfg2_canyon1k_n1.json", None),
        ("config 4: urban 10 094 triangles, 16 TX × 4096 RX, order 3, 4096 candidates (2.7·10^8 candidate-pairs)", "r2_bench_cfg4_urban10k_16tx_n1.json", "r2_bench_cfg4_urban10k_16tx_n8.json"),
        ("config 5: urban 49 922 triangles, 1 × 16 384 RX, order 4, 2048 candidates (3.4·10^7 candidate-pairs, 8.4·10^12 pairs)", "r2_bench_cfg5_urban50k_order4_n1.json", "r2_bench_cfg5_urban50k_order4_n8_c2048.json"),
    ):
        a, c = load(f1), load(f8) if f8 else None
        fin = load(f1.replace("r2_bench_cfg", "r2_bench_final2_cfg"))
        if a is None:
            continue
        line = (f"| {label} | {('%.1f ms' % fin['ms_per_step']) if fin else '—'} | {a['ms_per_step']:.1f} ms | {fmt(a['value'])} | {fmt(a['candidate_pairs_per_s'])} | "
                f"{a['executed_fraction_of_algorithmic'] * 100:.2f} % | {a['parity']['mismatches']} / {a['parity']['checked_pairs']} |")
        if c is not None:
            line += f" {c['ms_per_step']:.1f} ms | {fmt(c['value'])} | {a['ms_per_step'] / c['ms_per_step'] / 8:.2f} | {c['parity']['mismatches']} / {c['parity']['checked_pairs']} |"
        else:
            line += " — | — | — | — |"
        others.append(line)
    sweep = []
    for cnum in (256, 512, 1024, 2048):
        s = load(f"r2_bench_cfg5_urban50k_order4_n8_c{cnum}.json")
        if s:
            sweep.append(f"C = {cnum}: {s['ms_per_step']:.2f} ms, {fmt(s['value'])} pairs/s, {fmt(s['candidate_pairs_per_s'])} candidate-pairs/s")
    sec5 += f"""
Other BASELINE configurations (`bench.py --workload …`, forward + VJP every step, same definitions; `profiles/r2_bench_cfg*`):

| configuration | **1 GPU step, final code** | 1 GPU step (b93abce) | pairs decided /s | candidate-pairs /s | executed | parity (mismatches / checked) | 8 GPUs step | pairs decided /s | efficiency | parity |
|---|---|---|---|---|---|---|---|---|---|---|
{chr(10).join(others)}

The "final code" column comes from `--no-cpu` runs (`profiles/r2_bench_final2_cfg*_n1.json`: no in-bench oracle check; the
final code's parity is covered by the GPU test-suite and by config 3's in-bench check); every other column of this table
is the run at commit b93abce, whose in-bench parity figures are shown.

Config 5 sweep over the number of candidates on 8 GPUs (fwd + bwd): {'; '.join(sweep)}.

Round 1 for comparison (same config 3, weak scaling, CPU arm dense): step 95 ms, 7.1·10^12 pairs decided /s, 6.4 % of
the pairs executed; round-2 progression on the same workload: 95 → 68 (flat two-level cull behind the ordered
rows) → 131 / 153 (thread-per-candidate and all-segments-together traversals) → 78 → 29 → 26.5 (first step on the
deepest level of ≤ 32 nodes) → 20.4 (exact treatment of in-plane and rounding-residue segments) → 17.9 (head test as one
evaluation) → **16.8 ms** (`DESIGN.md` §4).
"""
    k = load("r2_kernels.json")
    lines = ["## 6. Per-kernel timings (round 2, `tools/bench_kernels.py`, `profiles/r2_kernels.json`)", "",
             'CUDA events, 3 warm-ups, median of 10, B200 @ 1965 MHz, through the Python API unless marked "kernel only".  `GB/s` = '
             "algorithmic bytes (SURVEY §8d per-unit figures) ÷ time against the measured 6549 GB/s copy bandwidth; for the all-pairs "
             "kernels that is the streamed-operand model and exceeds 1.  `ncu --set full` of the same kernels "
             "(`profiles/r2_ncu_full_primitives.json`, `r2_ncu_full_bench_step_final.json`, DRAM bytes ÷ kernel time): K1 192 µs, 5.68 TB/s = "
             "0.87 of peak; K5 122 µs, 4.31 TB/s = 0.66; stage A 470 µs, 2.77 TB/s = 0.42 with 68 % of the issue slots busy; K6b 709 µs, "
             "1.43 TB/s = 0.22 at 68 % issue (was 877 µs); all-pairs any-hit / nearest hit at 2^18 rays 3.00 / 6.16 ms → culled traversal "
             "0.48 / 0.89 ms.", "",
             "| kernel | ms | throughput | GB/s | frac of HBM peak | note |", "|---|---|---|---|---|---|"]
    for r in k["rows"]:
        thr = [(kk, v) for kk, v in r.items() if kk.endswith("_per_s")][0]
        g, f = r.get("gbs"), r.get("frac_of_hbm_peak")
        lines.append(f"| {r['kernel']} | {r['ms']:.3f} | {thr[1]:.3g} {thr[0].replace('_per_s', '')}/s | {('%.0f' % g) if g else '—'} | "
                     f"{('%.2f' % f) if f else '—'} | {r['note']} |")
    rl = load("r1_relaxed_trace.json")
    if rl:
        lines += ["", "### Relaxed (`smoothing_factor`) trace, forward and reverse mode (`tools/bench_relaxed.py`, `profiles/r1_relaxed_trace.json`, round 1)", "",
                  "| kernel | ms | relaxed pair evaluations / s | note |", "|---|---|---|---|"]
        for r in rl["rows"]:
            lines.append(f"| {r['kernel'].strip()} | {r['ms']:.3f} | {r['relaxed_tests_per_s']:.3g} | {r['note']} |")
    (ROOT / "BASELINE.md").write_text(b + sec5 + "\n" + "\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
