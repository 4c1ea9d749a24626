"""tools/make_traffic_json.py -- profiles/r2_traffic.json from the ncu CSVs of tools/gpu_profiles.sh.

Each gpurun_out/traffic_r2_*.csv holds dram__bytes_read.sum, dram__bytes_write.sum and gpu__time_duration.sum of ONE launch of a
shipped whole-tree kernel (ncu --csv).  bench.py looks its roofline.traffic up in the JSON by (kernel, taxa, patterns per GPU).

Usage: python tools/make_traffic_json.py [gpurun_out] > profiles/r2_traffic.json
"""
import csv
import glob
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3, "s": 1e3}


def short_name(full):
    """void p4b::cl_tree_dna2_kernel<(int)4, (int)2, (int)8, (int)2>(...) -> cl_tree_dna2_kernel<4,2,8,2>"""
    m = re.search(r"(cl_tree_\w+)<([^>]*)>", full)
    if not m:
        return full
    args = [re.sub(r"\(int\)", "", x).strip() for x in m.group(2).split(",")]
    return "%s<%s>" % (m.group(1), ",".join(args))


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    entries = []
    for path in sorted(glob.glob(os.path.join(d, "traffic_r2_*.csv"))):
        rows = [r for r in csv.reader(open(path)) if len(r) > 10]
        if len(rows) < 2:
            continue
        hdr = rows[0]
        ix = {h: i for i, h in enumerate(hdr)}
        vals, name = {}, None
        for r in rows[1:]:
            name = r[ix["Kernel Name"]]
            vals[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        tag = os.path.basename(path)[len("traffic_r2_"):-4]          # dna_1000000 | aa_cfg3 | aa_cfg4
        if tag.startswith("dna_"):
            taxa, pats = 200, int(tag[4:])
        elif tag == "aa_cfg3":
            taxa, pats = 100, 199999
        elif tag == "aa_cfg4":
            taxa, pats = 60, 50000
        else:
            continue
        key = short_name(name)
        if key.startswith("cl_tree_aa_kernel"):
            key += " x 4 categories"         # as pf.lastCLKernelName() reports it
        entries.append({"kernel": key, "taxa": taxa, "patterns_per_gpu": pats, "dram_bytes_read": vals.get("dram__bytes_read.sum"),
                        "dram_bytes_write": vals.get("dram__bytes_write.sum"), "kernel_ms_under_ncu": vals.get("gpu__time_duration.sum"),
                        "source": "profiles/%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, this round's shipped kernel)" % os.path.basename(path)})
    json.dump({"note": "DRAM traffic per launch of the shipped whole-tree kernels, one ncu capture per shard size (tools/gpu_profiles.sh)", "entries": entries},
              sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
