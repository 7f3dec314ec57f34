"""Regenerate BASELINE.md sections 5-6 and profiles/traffic.json from the JSON artefacts under profiles/
(round-1 bookkeeping script; paths are relative to /root/repo)."""
import json,re
json.dump({
 "source": "profiles/r1_ncu_dram_traffic_bench_workload.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum; bench workload urban10k_1tx_4096rx_order3; 30 consecutive launches of the blockage kernels = a little more than one step)",
 "dram_bytes_per_launch": 2071018240 + 24238336,
 "per_kernel": {
   "path_head_kernel<4> (16 launches: cascade passes over all candidates + greedy-round passes over the samples)": {"dram_bytes": 2071018240, "ns": 89383232},
   "hit_count_kernel<4> (14 launches: ordering pass rounds)": {"dram_bytes": 24238336, "ns": 10832832}},
 "note": "the blockage pass reads the 1.0 GB of path vertices written by stage A (60 B per candidate, fetched as partial 32 B sectors) and writes the mask bytes and the survivor lists; the packed mesh never comes from DRAM",
 "dram_bytes_per_launch_note": "per step = per bracketed blockage pass (the unit roofline.achieved uses)"
}, open('/root/repo/profiles/traffic.json','w'), indent=1)
p='/root/repo/bench.py'
s=open(p).read()
s=s.replace('''            # pack, area keys + gather, stage A, head pass, ring pass, 3 compaction kernels per step
            # (+ 4 CUB radix-sort kernels, not counted as ours)
            "gpu_launches": args.steps * 9,''','''            # per step: pack, area keys + gather, stage A, hit-count + iota + gather (ordering pass),
            # head pass, ring pass, 3 compaction kernels (+ 8 CUB radix-sort kernels, not counted as ours)
            "gpu_launches": args.steps * 12,''')
open(p,'w').write(s)

b=open('/root/repo/BASELINE.md').read()
b=b[:b.index('## 5. Measured results')]
d=json.load(open('/root/repo/profiles/r1_bench_v7.json'))
n2=json.load(open('/root/repo/profiles/r1_bench_v7_n2.json'))
n8=json.load(open('/root/repo/profiles/r1_bench_v6_n8.json'))
ref=json.load(open('/root/repo/profiles/r1_bench_reference_arm.json')); ref['value']=d['cpu_baseline']['value']; ref['cpu_baseline']=d['cpu_baseline']
big=json.load(open('/root/repo/profiles/r1_bench_v3_urban50k_order4.json'))
def fmt(x):
    m,e=f"{x:.2e}".split("e"); return f"{m}·10^{int(e)}"
sec5=f'''## 5. Measured results (round 1, NVIDIA B200 @ 1965 MHz, `profiles/`)

Workload = config 3 (urban grid 10 094 triangles, 1 TX × 4096 RX, order 3, 4096 candidates per GPU,
every candidate blockage-tested like the reference): 1.68·10^7 candidate-pairs, 6.7·10^7 rays,
6.77·10^11 (ray, triangle) pairs per step per GPU.  `python bench.py` / `torchrun ... bench.py --gpus N`.
`value` counts the pairs DECIDED (SURVEY §8d); "executed" counts the Möller–Trumbore evaluations
actually run (the any-hit query stops at the first blocking row of triangles).

| arm | GPUs | step | pairs decided /s (`value`) | executed tests /s | candidate-pairs /s | executed × 36 B vs 6549 GB/s | FP32 issue slots |
|---|---|---|---|---|---|---|---|
| differt_b200, inputs resident | 1 | {d['ms_per_step']:.0f} ms | {fmt(d['value'])} | {fmt(d['executed_tests_per_s'])} | {fmt(d['candidate_pairs_per_s'])} | {d['roofline']['frac']:.2f} | {d['roofline']['fp32_issue']['frac']:.2f} |
| differt_b200, end to end from host buffers | 1 | {d['e2e']['ms_per_step']:.0f} ms | {fmt(d['e2e']['value'])} | {fmt(d['e2e']['executed_tests_per_s'])} | — | — | — |
| differt_b200, weak scaling | 2 | {n2['ms_per_step']:.0f} ms | {fmt(n2['value'])} | {fmt(n2['executed_tests_per_s'])} | {fmt(n2['candidate_pairs_per_s'])} | — | — |
| differt_b200, weak scaling | 8 | {n8['ms_per_step']:.0f} ms | {fmt(n8['value'])} | {fmt(n8['executed_tests_per_s'])} | {fmt(n8['candidate_pairs_per_s'])} | — | — |
| differt_b200, API default mode (blockage only for candidates passing the cheap tests; identical outputs) | 1 | {d['default_mode']['ms_per_step']:.2f} ms | — | — | {fmt(d['default_mode']['candidate_pairs_per_s'])} | — | — |
| reference algorithm restated on CPU (C/OpenMP/AVX2 port; JAX/Warp not installable), {ref['cpu_baseline']['cores']} host cores, dense | 0 | — | {fmt(ref['value'])} | {fmt(ref['value'])} | — | — | — |

Other configurations (`--workload`): config 5 per GPU (49 922 triangles, 1 × 16 384 RX, order 4, 2048
candidates: 3.4·10^7 candidate-pairs, 8.4·10^12 pairs per step): {big['ms_per_step']:.0f} ms per step,
{fmt(big['value'])} pairs decided /s, {fmt(big['executed_tests_per_s'])} executed tests /s (fraction
{big['executed_fraction_of_algorithmic']:.3f}).  Config 2 end to end (986 triangles, 1 × 256 RX, ALL 971 210 order-2
candidates decoded on the device, compact kernel, valid paths merged in reference order): see the
"config 2" row of §6.  Config 4 per GPU (16 TX × 4096 RX, order 3, 512 of the 4096 candidates = the
shard one of 8 GPUs gets: 3.4·10^7 candidate-pairs): 202 ms per step, 6.7·10^12 pairs decided /s,
4.3·10^11 executed tests /s, default mode 1.0 ms.

North-star floor (10^9 tests/s per B200 at ≥ 60 % of the HBM roofline ⇔ 1.09·10^11 tests/s): exceeded
≈4× on executed tests and ≈50× on decided pairs.  Progression within the round (same workload, step
time): 1198 → 649 → 576 → 288 → 138 → 127 → 117 → 109 → 97 → 95 ms (`profiles/README.md`, `DESIGN.md` §4).
'''
k=json.load(open('/root/repo/profiles/r1_kernels.json'))
lines=["## 6. Per-kernel timings (round 1, `tools/bench_kernels.py`, `profiles/r1_kernels.json`)","",
'CUDA events, 3 warm-ups, median of 10, B200 @ 1965 MHz, through the Python API unless marked "kernel only".  `GB/s` = algorithmic bytes (SURVEY §8d per-unit figures) ÷ time; fraction against the measured 6549 GB/s copy bandwidth.  The all-pairs kernels are FP32-issue bound (DESIGN.md §4): their `GB/s` is the streamed-operand model (36 B per executed test) and can exceed 1.  Kernel-only `ncu` times of the HBM-bound kernels: K1 193 µs (5.6 TB/s DRAM, 86 % of peak), K5 121 µs (66 %), stage A 461 µs (2.8 TB/s, 43 %), K6b 878 µs (compute-bound at 168 registers).',"",
"| kernel | ms | throughput | GB/s | frac of HBM peak | note |","|---|---|---|---|---|---|"]
for r in k["rows"]:
    thr=[(kk,v) for kk,v in r.items() if kk.endswith("_per_s")][0]
    g=r.get('gbs'); f=r.get('frac_of_hbm_peak')
    lines.append(f"| {r['kernel']} | {r['ms']:.3f} | {thr[1]:.3g} {thr[0].replace('_per_s','')}/s | {('%.0f'%g) if g else '—'} | {('%.2f'%f) if f else '—'} | {r['note']} |")
# the relaxed (smoothing_factor) trace, measured separately (tools/bench_relaxed.py)
rl=json.load(open('/root/repo/profiles/r1_relaxed_trace.json'))
lines+=["","### Relaxed (`smoothing_factor`) trace, forward and reverse mode (`tools/bench_relaxed.py`, `profiles/r1_relaxed_trace.json`)","",
"Street canyon (986 triangles), 1 TX × 256 RX × 4096 sampled order-2 candidates = 1.05·10⁶ paths; the relaxed blockage is a clipped SUM over all triangles, so all 3.1·10⁹ (segment, triangle) pairs are evaluated — no early exit, no ordering.  The `min` over the reference's seven sigmoids per pair is computed as the sigmoid of the min of their arguments (one `expf` + one reciprocal per pair): 40.7 → 32.6 (reciprocal instead of division) → 14.7 (one sigmoid) → 11.8 ms (NaN-propagating `min.NaN.f32`, fused by ptxas into 3-input `FMNMX3.NAN`, instead of compare/select chains).  `ncu --set full` of the blockage kernel (`profiles/r1_ncu_full_relaxed_blockage.json`): issue-slot utilisation 89 % (124 warp instructions per evaluation), DRAM traffic 56 MB per launch (the packed mesh stays in L1/L2: 99.9 % L1 hit rate) — instruction-issue bound like the hard path.","",
"| kernel | ms | relaxed pair evaluations / s | note |","|---|---|---|---|"]
for r in rl["rows"]:
    lines.append(f"| {r['kernel'].strip()} | {r['ms']:.3f} | {r['relaxed_tests_per_s']:.3g} | {r['note']} |")
open('/root/repo/BASELINE.md','w').write(b+sec5+"\n"+"\n".join(lines)+"\n")
