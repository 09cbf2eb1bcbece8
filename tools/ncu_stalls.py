#!/usr/bin/env python
"""Stall-reason totals (warp-state samples of the source page) per captured kernel of an .ncu-rep:  ncu_stalls.py <rep> [regex]"""
import collections, csv, io, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(io.StringIO(out)))
i, seen = 0, collections.Counter()
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2; data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): data.append(rows[j])
            j += 1
        i = j
        seen[name] += 1
        if seen[name] > 1 or (pat and not re.search(pat, name)): continue
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = collections.Counter()
        for r in data:
            for s in stalls: tot[s] += int(r[hdr.index(s)])
        T = max(sum(tot.values()), 1)
        ie = hdr.index("Instructions Executed")
        print(f"== {name[:100]}\n   {len(data)} SASS instructions, {sum(int(r[ie]) for r in data)} warp instructions executed, {T} samples")
        print("   " + "  ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in tot.most_common(8)))
    else:
        i += 1
