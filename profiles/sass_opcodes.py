#!/usr/bin/env python
"""Per-kernel SASS opcode census of the shipped library: proves which kernels are Blackwell-native
(UTCIMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, ATOMS/RED = shared / global atomics).

    python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "pytorch-quantity_b200", "lib", "libpq_sm100.so")
OPS = ["UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "ATOMS", "RED", "ATOMG", "REDUX",
       "LDG", "STG", "LDS", "STS", "IMAD", "VIADD", "VIMNMX", "VIADDMNMX", "I2IP", "PRMT", "FMUL", "FADD", "DADD", "DMUL",
       "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    names = list(counts)
    if names:
        dm = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    print("# cuobjdump -sass %s  (arch sm_100a)" % os.path.relpath(LIB, REPO))
    print("# instruction counts per kernel; only opcodes of interest, '_total' = all instructions")
    total = collections.Counter()
    for fn, c in counts.items():
        short = re.sub(r"\(.*", "", demangle.get(fn, fn))
        cols = " ".join("%s=%d" % (op, sum(v for k, v in c.items() if k == op or k.startswith(op + "."))) for op in OPS
                        if any(k == op or k.startswith(op + ".") for k in c))
        print("%-90s total=%-6d %s" % (short[:90], c["_total"], cols))
        for k, v in c.items():
            total[k.split(".")[0]] += v
    print("# library totals: " + " ".join("%s=%d" % (op, total[op]) for op in OPS if total[op]))


if __name__ == "__main__":
    sys.exit(main())
