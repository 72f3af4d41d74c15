"""profiles/trace_dram_bytes.json from the per-launch ncu csv of tools/evidence_r2.sh step 4 (second pass only).
    python tools/trace_traffic.py gpurun_out/r2_trace_dram_per_launch.csv profiles/trace_dram_bytes.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launch = {}
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    launch.setdefault(int(r[ii]), {"name": r[ki].split("(")[0]})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(launch)
half = ids[len(ids) // 2:]                     # the second of the two passes
rd = sum(launch[i]["dram__bytes_read.sum"] for i in half)
wr = sum(launch[i]["dram__bytes_write.sum"] for i in half)
l2 = sum(launch[i]["lts__t_bytes.sum"] for i in half)
ns = sum(launch[i]["gpu__time_duration.sum"] for i in half)
out = {"dram_bytes_per_launch": (rd + wr) / len(half), "dram_bytes_per_pass": rd + wr, "dram_read_bytes_per_pass": rd,
       "dram_write_bytes_per_pass": wr, "l2_bytes_per_pass": l2, "trace_launches_per_pass": len(half),
       "ncu_ms_per_pass": ns / 1e6,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum over every k_trace_*/k_tail launch of the "
                 "second C2 pass (tools/evidence_r2.sh step 4: tools/profile_pass.py --config c2 --passes 2 --no-detail); "
                 "profiles/r2_trace_dram_per_launch.csv"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out)
