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


def fp64(path, n_uavs=1048576, pattern="uav_step"):
    """Executed instructions of the stepping kernel per UAV-step, summed by opcode from the source page (needs -lineinfo + --import-source on)."""
    import json
    import re

    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{pattern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1]
    H = rows[1]
    si, ti = H.index("Source"), H.index("Thread Instructions Executed")
    per = collections.Counter()
    seen_kernels = 0
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":  # a second launch of the same kernel follows: one is enough
            break
        if len(r) <= ti:
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si])
        if not m:
            continue
        per[m.group(1).split(".")[0]] += float(r[ti])
    tot = sum(per.values())
    g = lambda k: per.get(k, 0.0) / n_uavs
    out = {"kernel": name, "per_uav_step": {k: round(g(k), 1) for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU")}, "source": f"ncu --set full --import-source on, per-instruction 'Thread Instructions Executed' of {path} summed by opcode (DFMA = 2 flop)"}
    out["per_uav_step"]["all_instructions"] = round(tot / n_uavs, 1)
    out["executed_fp64_flop_per_uav_step"] = round(2 * g("DFMA") + g("DMUL") + g("DADD"), 1)
    out["as_written_census_flop_per_uav_step"] = 2550
    print(json.dumps(out, indent=1))


def traffic(path, n_uavs=1048576, pattern="uav_step"):
    import json

    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, U = rows[0], rows[1]
    for r in rows[2:]:
        if pattern in r[H.index("Kernel Name")]:
            def val(metric):
                i = H.index(metric)
                v = float(r[i].replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[U[i]]
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            print(json.dumps({"kernel": r[H.index("Kernel Name")], "n_uavs": n_uavs, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                              "dram_bytes_per_uav_step": round((rd + wr) / n_uavs, 2), "source": f"ncu --set full --clock-control none, {path}"}, indent=1))
            return


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "fp64": fp64, "traffic": traffic}[sys.argv[1]](sys.argv[2])
