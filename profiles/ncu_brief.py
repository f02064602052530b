#!/usr/bin/env python
"""One tracked text summary per `ncu --set full` capture: the metrics the roofline discussion uses plus the SASS
instruction mix of the kernel (ncu source page), per warp-element when the element count is given.

    python profiles/ncu_brief.py gpurun_out/prof_x.ncu-rep profiles/r02_ncu_x.txt [elements] ["command line"]
"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], stdout=subprocess.PIPE, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    elements = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] else 0.0
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[-1]
    lines = ["# ncu --set full --clock-control none --import-source on  (%s)" % cmd,
             "# kernel: %s" % vals[hdr.index("Kernel Name")][:150]]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append("%-86s %s %s" % (w, vals[i], units[i]))
    src = page(rep, "source")
    h = src[1]
    ia, ie = h.index("Source"), h.index("Instructions Executed")
    ops, tot = collections.Counter(), 0
    for r in src[2:]:
        if len(r) <= ie:
            continue
        n = int(r[ie] or 0)
        tot += n
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
        ops[".".join((m.group(1) if m else r[ia]).split(".")[:2])] += n
    we = elements / 32.0
    lines.append("# SASS instruction mix (warp instructions executed%s)" % (", per warp-element = 32 output elements" if we else ""))
    lines.append("%-28s %14d %s" % ("TOTAL", tot, ("%8.2f" % (tot / we)) if we else ""))
    for k, v in ops.most_common(26):
        lines.append("%-28s %14d %s" % (k, v, ("%8.2f" % (v / we)) if we else ""))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
