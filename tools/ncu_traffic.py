"""profiles/ncu_traffic.json from an `ncu -i X.ncu-rep --page raw --csv` export: per kernel the DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), tagged with the sha256 of the sources (csrc/, the header, the nvcc flags) the
captured library was built from, so bench.py only quotes `roofline.traffic` when it is measuring a build of those sources.

    python tools/ncu_traffic.py <raw.csv> [<raw2.csv> ...] [--so herbolution_b200/libherbgpu.so] [--merge]"""
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    args = sys.argv[1:]
    so = os.path.join(ROOT, "herbolution_b200", "libherbgpu.so")
    merge = "--merge" in args
    if "--so" in args:
        so = args[args.index("--so") + 1]
    files = [a for a in args if a.endswith(".csv")]
    sys.path.insert(0, ROOT)
    from herbolution_b200 import build as b
    from bench import KERNEL_SOURCES
    hh = hashlib.sha256()
    for name in KERNEL_SOURCES:
        hh.update(name.encode() + b"\0" + open(os.path.join(b.CSRC, name), "rb").read())
    hh.update(" ".join(b.NVCC_FLAGS).encode())
    sha = hh.hexdigest()
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    out = {"src_sha256": sha, "kernels": {}}
    if merge and os.path.exists(out_path):
        old = json.load(open(out_path))
        if old.get("src_sha256") == sha:
            out = old
    for fn in files:
        rows = list(csv.reader(open(fn)))
        hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
        ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        for r in data:
            if len(r) != len(hdr):
                continue
            key = re.sub(r"\(.*", "", r[ki]).replace("hb::(anonymous namespace)::", "").replace("void ", "").replace("unnamed>::", "").strip()
            key = re.sub(r"<.*", "", key)
            rd = float(r[ri].replace(",", "")) * UNIT[units[ri]]
            wr = float(r[wi].replace(",", "")) * UNIT[units[wi]]
            out["kernels"][key] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                                   "duration_under_ncu": f"{r[ti]} {units[ti]}", "kernel": r[ki][:120],
                                   "source": f"ncu --set full, {os.path.basename(fn)}"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(out_path, sorted(out["kernels"]))


if __name__ == "__main__":
    main()
