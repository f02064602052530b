#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py r01            # reads gpurun_out/launches.csv and prof_*.ncu-rep
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
WANT = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, mi, ui = (hdr.index(n) for n in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    tot, n = collections.OrderedDict(), 0
    for row in r:
        if len(row) <= vi or row[mi] != "gpu__time_duration.sum":
            continue
        v = float(row[vi].replace(",", ""))
        v *= {"usecond": 1e-3, "us": 1e-3, "nsecond": 1e-6, "ns": 1e-6}.get(row[ui], 1.0)
        d = tot.setdefault(re.sub(r"\(.*", "", row[ki])[:90], [0, 0.0])
        d[0] += 1
        d[1] += v
        n += 1
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(REPO, "profiles", tag + "_launches_bench.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 2 --warmup 3 "
                "--no-cpu-baseline (PQ_BENCH_NO_AUTOTUNE=1)\n# per-launch times are cold-cache and serialised: "
                "compare SHARES, not absolutes\n# %d launches, %.2f ms total\n" % (n, total))
        f.write("%-92s %6s %10s %6s\n" % ("kernel", "count", "total_ms", "share"))
        for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-92s %6d %10.3f %5.1f%%\n" % (k, c, t, 100 * t / total))
        ours = sum(t for k, (c, t) in tot.items() if "pq::" in k)
        f.write("# kernels of this repo (pq::*): %.3f ms = %.1f%% of the profiled launches\n" % (ours, 100 * ours / total))


def _num(txt, unit):
    v = float(txt.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


def full(tag):
    import json
    traffic_path = os.path.join(REPO, "profiles", tag + "_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        with open(os.path.join(REPO, "profiles", "%s_ncu_%s.txt" % (tag, name)), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on -k regex:%s (1 launch)\n" % name)
            for vals in rows[2:]:
                for w in WANT:
                    if w in hdr:
                        i = hdr.index(w)
                        f.write("%-92s %s %s\n" % (w, vals[i], units[i]))
                f.write("\n")
                ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
                traffic[name] = {"kernel": vals[hdr.index("Kernel Name")],
                                 "dram_bytes_per_launch": _num(vals[ir], units[ir]) + _num(vals[iw], units[iw]),
                                 "ncu_time_us": float(vals[it].replace(",", "")) * {"msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3, "second": 1e6}.get(units[it], 1.0), "source": "ncu --set full, 1 launch"}
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    full(tag)
