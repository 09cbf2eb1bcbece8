#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` captures:  ncu_traffic.py <workload> <file.ncu-rep> [<workload> <file.ncu-rep> ...]
Per kernel class (the names bench.py uses) the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one launch,
averaged over the captured launches of that class.  bench.py copies the dominant kernel's figure into roofline.traffic."""
import collections, csv, io, json, os, re, subprocess, sys

CLASSES = [("eig_small", "eig"), ("bands_sub|interp_small", "interp"), ("mep_sums", "mep_sums"), ("sweep_ahc", "sweep_ahc"), ("sweep_photo", "sweep_photo"),
           ("fourier|stage_[abc]|sep_", "fourier"), ("tridiag", "tridiag"), ("tql", "tql"), ("zrot", "zrot"), ("eig_rot_large", "eig_rot")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def parse(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    agg = collections.defaultdict(list)
    for r in rows[2:]:
        cls = next((c for pat, c in CLASSES if re.search(pat, r[ik])), None)
        if cls:
            agg[cls].append(float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]])
    return {k: sum(v) / len(v) for k, v in agg.items()}


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    for wl, rep in zip(sys.argv[1::2], sys.argv[2::2]):
        data.setdefault(wl, {}).update(parse(rep))
        data[wl]["_source"] = os.path.basename(rep)
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data, indent=1, sort_keys=True))
