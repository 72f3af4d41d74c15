"""Dynamic instruction mix + hot SASS ranges from `ncu --page source --csv` output.
    ncu -i rep --page source --csv --kernel-id ::regex:NAME:K > src.csv ; python tools/ncu_source_hot.py src.csv"""
import csv
import collections
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples", "Avg. Threads Executed")}
tot_inst = 0
by_op = collections.defaultdict(lambda: [0, 0, 0])
lines = []
for r in rows[hi + 1:]:
    if len(r) <= ci["# Samples"]:
        continue
    src = r[ci["Source"]]
    try:
        ie = int(r[ci["Instructions Executed"]]); te = int(r[ci["Thread Instructions Executed"]]); smp = int(r[ci["# Samples"]])
    except ValueError:
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
    op = m.group(1) if m else "?"
    by_op[op][0] += ie; by_op[op][1] += te; by_op[op][2] += smp
    tot_inst += ie
    lines.append((ie, te, smp, src))
print(f"total warp instructions {tot_inst}")
print("opcode            warp-inst   share   avg lanes   stall samples")
for op, (ie, te, smp) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:14s} {ie:12d}  {ie/tot_inst:6.1%}   {te/max(ie,1):6.1f}   {smp}")
tot_s = sum(l[2] for l in lines)
print(f"\ntop stall-sample instructions (of {tot_s} samples)")
for ie, te, smp, src in sorted(lines, key=lambda l: -l[2])[:25]:
    print(f"{smp:7d} {smp/tot_s:6.1%}  exec={ie:10d} lanes={te/max(ie,1):5.1f}  {src[:90]}")
