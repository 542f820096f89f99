#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1d.csv            > profiles/r1_launches.txt
    python tools/ncu_summary.py kernels  gpurun_out/prof_r1d.ncu-rep            > profiles/r1_kernels.txt
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tick = {k: v for k, v in agg.items() if "dfma_kernel" not in k}
    tot = sum(a[1] for a in tick.values())
    print(f"# {path}: gpu__time_duration.sum per kernel (ncu --clock-control none; cold-cache, serialised: compare SHARES)")
    print(f"{'avg us':>10} {'count':>6} {'share':>7}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        share = f"{100 * t / tot:6.1f}%" if k in tick else "   (mb)"
        print(f"{t / c:10.1f} {c:6d} {share}  {k}")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.max"]


def kernels(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    seen = set()
    print(f"# {path}: ncu --set full --clock-control none, one launch per kernel")
    for r in rows[2:]:
        name = r[H.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        print(f"\n## {name[:110]}")
        for w in WANT:
            if w in H:
                i = H.index(w)
                print(f"  {w:70s} {r[i]:>16s} {U[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])
