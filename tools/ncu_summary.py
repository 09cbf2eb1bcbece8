#!/usr/bin/env python
"""Summarise ncu artefacts brought back by gpurun into small text files for profiles/.

  launches  <launches.csv>                      per-kernel count / total ms / share of the step
  raw       <file.ncu-rep> [kernel-regex]       the raw-page metrics the roofline discussion uses, per captured launch
  lines     <file.ncu-rep> <cubin.sass> <mangled-substr> <source.cu> [N [kernel-regex]]   executed instructions / stall samples by source line
(sass file: `cuobjdump -xelf all libmepb.so; nvdisasm -g -c x.cubin > x.sass`)
"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        v = float(r[-1].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[hdr.index("Metric Unit")], 1e-6)
        a = agg[r[hdr.index("Kernel Name")]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms of kernel time (ncu: cold-cache, serialised — compare shares)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[0]:5d} launches {v[1]:11.3f} ms {100 * v[1] / tot:6.2f}%  {k[:110]}")


def raw(rep, pat=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not re.search(pat, name):
            continue
        print("== " + name[:120])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:70s} {r[i]:>18s} {units[i]}")


def lines(rep, sass, func, src, top=40, kernel=None, by="inst"):
    """kernel: regex on the (demangled) kernel name of the report; default = first captured launch"""
    top = int(top)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = 0
    if kernel:
        start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and re.search(kernel, r[1]))
    hdr_i = next(i for i, r in enumerate(rows) if i >= start and r and r[0] == "Address")
    hdr = rows[hdr_i]
    data = []
    for r in rows[hdr_i + 1:]:
        if r and r[0] == "Kernel Name":
            break                       # first captured launch only
        if len(r) == len(hdr):
            data.append(r)
    ie, st = hdr.index("Instructions Executed"), hdr.index("# Samples")
    cur, inf, il = None, False, []
    for ln in open(sass).read().splitlines():
        if re.match(r"^\s*\.section\s+\.text\.", ln) or re.match(r"^\.text\.", ln):
            inf = func in ln
            continue
        if not inf:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
        elif re.match(r"^\s+/\*[0-9a-f]{4,}\*/", ln):
            il.append(cur)
    n = min(len(il), len(data))
    print(f"# {func}: {len(il)} SASS instructions in the cubin, {len(data)} in the report (must match for the mapping to hold)")
    agg, smp = collections.Counter(), collections.Counter()
    for i in range(n):
        agg[il[i]] += int(data[i][ie])
        smp[il[i]] += int(data[i][st])
    tot, tots = sum(agg.values()), max(sum(smp.values()), 1)
    text = open(src).read().splitlines()
    base = src.split("/")[-1]
    print(f"# total warp instructions executed {tot}, stall samples {tots}")
    order = sorted(agg.items(), key=lambda x: -x[1]) if by == "inst" else sorted(agg.items(), key=lambda x: -smp[x[0]])
    for (f, l), v in order[:top]:
        code = text[l - 1].strip()[:90] if f and f.endswith(base) and l <= len(text) else (f or "?").split("/")[-1]
        print(f"{100 * v / tot:6.2f}% inst {100 * smp[(f, l)] / tots:6.2f}% samples  line {l:4d}  {code}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    {"launches": launches, "raw": raw, "lines": lines}[cmd](*sys.argv[2:])
