"""Aggregate an `ncu --page source --csv` SASS export by CUDA source line, using nvdisasm -g line info.

usage: ncu_lines.py <sass_csv> <cubin> <function-substring> [topN]
"""
import csv, re, subprocess, sys
from collections import defaultdict

sass_csv, cubin, fn_sub = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find function section
line_of = {}
cur_line = None
in_fn = False
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        in_fn = fn_sub in m.group(1)
        cur_line = None
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
recs = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(recs[0]["Address"], 16)
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
keys = ["Instructions Executed", "# Samples", "L1 Wavefronts Shared", "Thread Instructions Executed"]
for r in recs:
    off = int(r["Address"], 16) - base
    ln = line_of.get(off, ("?", 0))
    for k in keys:
        v = float(r.get(k, 0) or 0)
        agg[ln][k] += v
        tot[k] += v
print({k: tot[k] for k in keys})
src_cache = {}
def src(ln):
    f, n = ln
    if f == "?": return ""
    import glob
    if f not in src_cache:
        c = glob.glob(f"/root/repo/herbolution_b200/csrc/{f}")
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["Instructions Executed"])[:topn]:
    print(f"{ln[0]}:{ln[1]:<4d} inst {100*a[keys[0]]/tot[keys[0]]:5.1f}%  samp {100*a[keys[1]]/max(tot[keys[1]],1):5.1f}%  "
          f"smemwf {100*a[keys[2]]/max(tot[keys[2]],1):5.1f}%  thr/inst {a[keys[3]]/max(a[keys[0]],1):4.1f} | {src(ln)}")
