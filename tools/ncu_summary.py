#!/usr/bin/env python3
"""Summaries committed under profiles/: (1) per-kernel share of a `--metrics gpu__time_duration.sum` launch list,
(2) key metrics of an `ncu --set full` capture incl. the DRAM traffic figure bench.py reports.

    python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep ENVS PRECISION [--json profiles/ncu_traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        agg[r[ki][:90]][0] += 1
        agg[r[ki][:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms of kernel time (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{v[1]:10.3f} ms {v[0]:5d} launches {100 * v[1] / tot:6.2f}%  {k}")


def full(path, envs, precision, out_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct"]
    vals = {}
    for i, name in enumerate(h):
        if name in want or name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio"):
            vals[name] = (v[i], u[i])
    print(f"# {path}  kernel: {v[h.index('Kernel Name')] if 'Kernel Name' in h else ''}")
    for k in sorted(vals):
        try:
            if k.startswith("smsp__average_warps_issue_stalled") and float(vals[k][0].replace(",", "")) < 0.05:
                continue
        except ValueError:
            pass
        print(f"{k:90s} {vals[k][0]:>18s} {vals[k][1]}")

    def num(name):
        x, unit = vals[name]
        return float(x.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    print(f"dram traffic per launch: {traffic / 1e6:.2f} MB")
    if out_json:
        json.dump({"envs": int(envs), "precision": precision, "dram_bytes_per_launch": traffic, "source": path,
                   "kernel_ms_under_ncu": vals["gpu__time_duration.sum"][0] + " " + vals["gpu__time_duration.sum"][1]},
                  open(out_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        oj = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
        full(sys.argv[2], sys.argv[3], sys.argv[4], oj)
