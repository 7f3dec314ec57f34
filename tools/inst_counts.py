"""Per-step kernel / warp-instruction counts of the bench step from an ncu launch list.

    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv \
        --log-file profiles/rN_launches_<name>.csv python bench.py --profile-only --steps 3 [--workload W]
    python tools/inst_counts.py profiles/rN_launches_<name>.csv --steps 3 [--workload W] [--world 1]

Writes/updates profiles/inst_counts.json: for the key "<workload>@<world>" the warp instructions per
step of the blockage kernels (what bench.py's roofline divides by the measured kernel time), the
number of our launches per step and each kernel's share of the step's GPU time.  bench.py reads this
file; it never hard-codes an instruction count.
"""

from __future__ import annotations

import argparse
import collections
import csv
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BLOCKAGE = ("path_head_kernel", "path_walk_kernel", "hit_count_kernel", "intersect_kernel")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, required=True)
    ap.add_argument("--workload", default="urban10k_1tx_4096rx_order3")
    ap.add_argument("--world", type=int, default=1)
    args = ap.parse_args()
    hdr, per = None, collections.OrderedDict()
    for r in csv.reader(open(args.csv)):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            name = d["Kernel Name"].split("(")[0].replace("void ", "")
            e = per.setdefault(name, {"launches": set(), "ns": 0.0, "inst": 0.0, "dram": 0.0})
            e["launches"].add(d["ID"])
            v = float(d["Metric Value"].replace(",", ""))
            if d["Metric Name"] == "gpu__time_duration.sum":
                e["ns"] += v
            elif d["Metric Name"] == "smsp__inst_executed.sum":
                e["inst"] += v
            elif d["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(d.get("Metric Unit", "byte"), 1.0)
                e["dram"] += v * scale
    total_ns = sum(e["ns"] for e in per.values())
    ours = {k: e for k, e in per.items() if k.startswith("drt::")}
    block = {k: e for k, e in ours.items() if any(b in k for b in BLOCKAGE)}
    entry = {
        "source": f"{Path(args.csv).name}: ncu launch list of `bench.py --profile-only --steps {args.steps}`"
                  f" ({args.workload}, {args.world} GPU), totals / {args.steps}",
        "blockage_warp_instructions_per_step": sum(e["inst"] for e in block.values()) / args.steps,
        "blockage_ns_per_step_under_ncu": sum(e["ns"] for e in block.values()) / args.steps,
        "dram_bytes_per_step": sum(e["dram"] for e in per.values()) / args.steps,
        "blockage_dram_bytes_per_step": sum(e["dram"] for e in block.values()) / args.steps,
        "our_launches_per_step": round(sum(len(e["launches"]) for e in ours.values()) / args.steps),
        "all_launches_per_step": round(sum(len(e["launches"]) for e in per.values()) / args.steps),
        "kernels": {k: {"launches_per_step": len(e["launches"]) / args.steps, "us_per_step": e["ns"] / args.steps / 1e3,
                        "warp_instructions_per_step": e["inst"] / args.steps,
                        "dram_bytes_per_step": e["dram"] / args.steps,
                        "share_of_gpu_time": e["ns"] / total_ns} for k, e in per.items()},
    }
    out = ROOT / "profiles" / "inst_counts.json"
    data = json.loads(out.read_text()) if out.exists() else {}
    data[f"{args.workload}@{args.world}"] = entry
    out.write_text(json.dumps(data, indent=1) + "\n")
    print(json.dumps({k: v for k, v in entry.items() if k != "kernels"}, indent=1))
    for k, e in entry["kernels"].items():
        print(f"{k[:70]:70s} {e['launches_per_step']:6.1f} x  {e['us_per_step']:10.1f} us  {e['share_of_gpu_time']:6.3f}")


if __name__ == "__main__":
    main()
