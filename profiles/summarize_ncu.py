#!/usr/bin/env python3
"""Turns ncu output brought back in gpurun_out/ into the small text summaries committed here.

  python profiles/summarize_ncu.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
  python profiles/summarize_ncu.py full gpurun_out/prof_viterbi_r01.ncu-rep > profiles/r01_viterbi_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v * scale))
    agg = OrderedDict()
    for k, ms in rows:
        name = k.split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / total:.1f}% |")
    print(f"\ntotal {total:.3f} ms over {len(rows)} launches (ncu-serialised, cold-cache: compare shares, not absolutes)")


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__inst_executed.sum",
    "smsp__inst_executed.avg.per_cycle_active", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_tensor_cycles_active", "achieved_occupancy", "sm__cycles_elapsed.avg", "lts__throughput",
    "l1tex__throughput", "smsp__cycles_active.avg",
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    if len(rd) < 3:
        print("no data")
        return
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS):
                print(f"| {h} | {row[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
