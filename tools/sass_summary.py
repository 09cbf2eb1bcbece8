#!/usr/bin/env python
"""Per-kernel static summary of libmepb.so: registers / spills / shared memory (cuobjdump -res-usage) and counts of the SASS
mnemonics that show what the kernels run on (DMMA = FP64 tensor-core MMA, DFMA/DADD/DMUL = FP64 pipe, UBLKCP = TMA bulk copy,
SYNCS = mbarrier, MUFU, LDS/STS, LDG/STG).  No GPU needed.

    python tools/sass_summary.py [path/to/libmepb.so] > profiles/sass_summary.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mepinterp_b200", "libmepb.so")
MNEMONICS = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"mepb::\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return name.split("(")[0]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = OrderedDict()
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            d = dict(re.findall(r"(\w+):(\d+)", line))
            usage[cur] = d
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = Counter()
            counts[cur]["_total"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn] += 1
                    break
    names = demangle(list(usage))
    print(f"# {os.path.relpath(LIB, ROOT)}: per-kernel resources (cuobjdump -res-usage) and SASS mnemonic counts (cuobjdump -sass)")
    print("# DMMA = mma.sync f64 on the FP64 tensor cores; UBLKCP = cp.async.bulk (TMA); SYNCS = mbarrier; STL/LDL = local-memory (spill) traffic")
    hdr = ["kernel", "regs", "stack", "smem", "instr"] + MNEMONICS
    rows = []
    for fn, d in usage.items():
        c = counts.get(fn, Counter())
        rows.append([short(names.get(fn, fn)), d.get("REG", "?"), d.get("STACK", "?"), d.get("SHARED", "?"), str(c.get("_total", 0))] +
                    [str(c.get(mn, 0)) for mn in MNEMONICS])
    rows.sort(key=lambda r: r[0])
    w = [max(len(h), max((len(r[i]) for r in rows), default=0)) for i, h in enumerate(hdr)]
    print("  ".join(h.ljust(w[i]) for i, h in enumerate(hdr)))
    for r in rows:
        print("  ".join(x.ljust(w[i]) for i, x in enumerate(r)))
    tot = Counter()
    for c in counts.values():
        tot.update(c)
    print("# totals: " + ", ".join(f"{mn}={tot.get(mn, 0)}" for mn in MNEMONICS))
    for bad in ("UTCHMMA", "UTCQMMA", "LDTM", "HGMMA"):
        print(f"# {bad}: {sass.count(bad)} (tcgen05 has no FP64 kind: expected 0)")


if __name__ == "__main__":
    main()
