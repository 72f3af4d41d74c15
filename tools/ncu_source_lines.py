"""Per CUDA source line: warp instructions executed, avg active lanes, stall samples -- from
    ncu -i rep --page source --csv --print-source cuda,sass --kernel-id :::K > src.csv
    python tools/ncu_source_lines.py src.csv [file-substring]"""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]; hdr = None; continue
    if r[0] == "Line No":
        hdr = r
        ci = {n: hdr.index(n) for n in ("Address", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
        continue
    if hdr is None or cur_file is None or len(r) <= ci["Thread Instructions Executed"]:
        continue
    if not r[0].strip().isdigit():  # SASS row; the CUDA row above it already aggregates
        continue
    try:
        ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]]); smp = int(r[ci["# Samples"]])
    except ValueError:
        continue
    if ie == 0 and smp == 0:
        continue
    agg[(cur_file, int(r[0]))] = (ie, te, smp, r[1])
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print(f"total warp-inst {tot}  samples {tots}")
for (f, ln), (ie, te, smp, src) in agg.items():
    if want in f:
        print(f"{f.split('/')[-1]:18s}:{ln:4d} {ie/tot:6.2%} inst  {smp/max(tots,1):6.2%} smp  lanes {te/max(ie,1):5.1f} | {src.strip()[:100]}")
