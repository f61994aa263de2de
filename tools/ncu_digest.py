#!/usr/bin/env python3
"""Digest an ncu report into a small text file (what profiles/ keeps): per-launch key metrics,
stall reasons, and the instructions with most stall samples.   python tools/ncu_digest.py X.ncu-rep out.txt"""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {rep}: {len(data)} launch(es) captured with ncu --set full --clock-control none"]
    for d in data:
        lines.append("kernel: " + d[col["Kernel Name"]][:160])
        lines.append("  " + "  ".join(f"{k}={d[col[k]]}{units[col[k]]}" for k in KEEP if k in col))
        stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(float(d[i]), 2)
                  for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h and d[i] and float(d[i]) > 0.2}
        lines.append("  stalls (warps per issue-active cycle): " + json.dumps(stalls))
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        body = []
        for r in src[2:]:
            if len(r) < len(h) or r[0] == "Address":
                break
            body.append(r)
        tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
        keys = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        agg = {k: sum(int(r[ix[k]]) for r in body) for k in keys}
        lines.append(f"source page (first launch): {len(body)} SASS instructions, {tot} stall samples; by reason: " +
                     json.dumps({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01}))
        for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:12]:
            why = {k: r[ix[k]] for k in keys if int(r[ix[k]]) > 0.2 * max(1, int(r[ix["# Samples"]]))}
            lines.append(f"  {int(r[ix['# Samples']]):6d}  {r[1].strip()[:70]:70s} {json.dumps(why)}")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))


if __name__ == "__main__":
    main()
