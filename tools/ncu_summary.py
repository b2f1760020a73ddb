#!/usr/bin/env python
"""Summarise gpurun_out/<tag>_launches.csv (ncu launch list) and <tag>_<kernel>.ncu-rep (ncu
--set full) into profiles/<tag>_*.txt.   python tools/ncu_summary.py r01a fps knn"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct"]


def launches(tag):
    p = os.path.join(OUT, f"{tag}_launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    dst = os.path.join(PROF, f"{tag}_launches.txt")
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none over bench.py --steps 1 --warmup 1 (+3 profile steps)\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
        f.write(f"# total {tot:.3f} ms over {sum(n for n, _ in agg.values())} launches\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{ms:10.3f} ms {100 * ms / tot:5.1f}%  n={n:4d}  avg {1e3 * ms / n:9.1f} us  {k}\n")
    print(open(dst).read())


def full(tag, name):
    rep = os.path.join(OUT, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    dst = os.path.join(PROF, f"{tag}_{name}_full.txt")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on: {len(rows) - 2} launch(es) matching {name}\n")
        for v in rows[2:]:
            for w in WANT:
                for i, x in enumerate(h):
                    if x == w:
                        f.write(f"{w} = {v[i]} {u[i]}\n")
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    for n in sys.argv[2:]:
        full(tag, n)
