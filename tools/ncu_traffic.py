#!/usr/bin/env python3
"""Record ncu-measured DRAM traffic per launch in profiles/traffic.json, stamped with the
fingerprint of the kernel sources it was captured for (bench.py refuses a stale stamp).

    python tools/ncu_traffic.py <report.ncu-rep> <kernel class> [<summary .txt to write>]

Run here (no GPU needed): it only reads the report with `ncu -i`."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from new_cg_variants_b200 import build as _b  # noqa: E402

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, cls = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    lines = [f"# {os.path.basename(rep)}: {len(data)} launches of {data[0][col['Kernel Name']][:100]}"]
    for d in data:
        rd = float(d[col["dram__bytes_read.sum"]]) * to_bytes[units[col["dram__bytes_read.sum"]]]
        wr = float(d[col["dram__bytes_write.sum"]]) * to_bytes[units[col["dram__bytes_write.sum"]]]
        per.append(rd + wr)
        lines.append("  ".join(f"{k}={d[col[k]]}{units[col[k]]}" for k in KEEP if k in col))
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(d[i])
                  for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h and d[i] and float(d[i]) > 0.2}
        lines.append("    stalls (warps per issue-active cycle): " + json.dumps(stalls))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(path)) if os.path.exists(path) else {}
    t = {k: v for k, v in t.items()
         if isinstance(v, dict) and v.get("kernel_sources_sha") == _b.kernel_fingerprint(k)}   # drop stale / legacy entries
    t[cls] = {"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per),
              "kernel_sources_sha": _b.kernel_fingerprint(cls),
              "source": f"ncu --set full --clock-control none ({os.path.basename(rep)}): dram__bytes_read.sum + dram__bytes_write.sum"}
    json.dump(t, open(path, "w"), indent=1)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    print(f"{cls}: {t[cls]['dram_bytes_per_launch'] / 1e6:.1f} MB per launch")


if __name__ == "__main__":
    main()
