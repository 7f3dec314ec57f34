"""Per-kernel summary of an `ncu --set full` report → JSON under profiles/.

    ncu --set full --import-source on --clock-control none -k regex:'...' -o gpurun_out/X python ...
    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/X.csv
    python tools/ncu_summary.py /tmp/X.csv profiles/rN_ncu_X.json [--note "..."]

For every profiled launch: duration, warp instructions, issue-slot utilisation, occupancy, pipe
utilisations, DRAM bytes and achieved DRAM bandwidth against MEASURED_PEAKS.json, L1/L2 hit rates,
registers, and the stall reasons above 0.15 warps per issue.
"""

from __future__ import annotations

import argparse
import csv
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEEP = {
    "gpu__time_duration.sum": "duration",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_warp_instruction",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_rate_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_ncu_peak",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__block_size": "block_size",
    "launch__grid_size": "grid_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem",
    "launch__shared_mem_per_block_static": "static_smem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
}
UNIT_SCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
              "second": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("out")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, units = rows[0], rows[1]
    peaks = ROOT / "MEASURED_PEAKS.json"
    hbm = float(json.loads(peaks.read_text())["hbm_gbs"]) if peaks.exists() else 6650.0
    out = {"source": f"ncu --set full --clock-control none; {Path(args.csv).name}", "note": args.note,
           "hbm_peak_gbs": hbm, "kernels": []}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name", "")[:160]}
        for key, name in KEEP.items():
            if key in d and d[key] not in ("", "n/a"):
                try:
                    v = float(d[key].replace(",", ""))
                except ValueError:
                    continue
                k[name] = v * UNIT_SCALE.get(u.get(key, ""), 1.0) if name in ("duration", "dram_read", "dram_write") else v
        stalls = {}
        for key, v in d.items():
            if "issue_stalled" in key and key.endswith("per_issue_active.ratio"):
                try:
                    f = float(v)
                except ValueError:
                    continue
                if f > 0.15 and "selected" not in key.replace("not_selected", ""):
                    stalls[key.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = round(f, 2)
        k["stalls_warps_per_issue"] = stalls
        if "duration" in k and k["duration"] > 0:
            byts = k.get("dram_read", 0.0) + k.get("dram_write", 0.0)
            k["dram_bytes"] = byts
            k["dram_gbs"] = byts / k["duration"] / 1e9
            k["dram_frac_of_measured_hbm_peak"] = k["dram_gbs"] / hbm
            if "warp_instructions" in k:
                k["warp_instructions_per_s"] = k["warp_instructions"] / k["duration"]
        out["kernels"].append(k)
    Path(args.out).write_text(json.dumps(out, indent=1) + "\n")
    for k in out["kernels"]:
        print(f"{k['kernel'][:60]:60s} {k.get('duration', 0) * 1e6:10.1f} us  issue {k.get('issue_slots_busy_pct', 0):5.1f}%  "
              f"dram {k.get('dram_gbs', 0):7.1f} GB/s ({k.get('dram_frac_of_measured_hbm_peak', 0):.2f})  regs {k.get('registers_per_thread', 0):.0f}  "
              f"occ {k.get('achieved_occupancy_pct', 0):.0f}%  stalls {k['stalls_warps_per_issue']}")


if __name__ == "__main__":
    main()
