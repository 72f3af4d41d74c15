"""Per-kernel totals and the launch-by-launch list of the LAST pass in an ncu `--metrics gpu__time_duration.sum --csv` log.
    python tools/launch_summary.py profiles/r2_launches_world8_share.csv [title] > profiles/..._summary.txt
A pass starts at a k_raygen launch that follows a k_splat launch (or at the first k_raygen)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v
    name = r[ki].split("(")[0].replace("void ", "").replace("bpt::", "")
    launches.append((name, us))
starts = [i for i, (n, _) in enumerate(launches) if n.startswith("k_raygen") and (i == 0 or launches[i - 1][0].startswith("k_splat") or launches[i - 1][0].startswith("k_write"))]
# group batches into passes: batches of one pass are enqueued back to back; take the trailing batches that make up the last pass
n_batches = int(sys.argv[3]) if len(sys.argv) > 3 else 1
first = starts[-n_batches] if len(starts) >= n_batches else 0
last = launches[first:]
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("Per-launch times under ncu are serialised and cold-cache: compare shares, not absolutes.\n")
tot = sum(us for _, us in last)
agg = collections.OrderedDict()
for n, us in last:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
print(f"last pass ({n_batches} batch(es)), {len(last)} launches, {tot:.1f} us of kernel time:")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:28s} {c:4d} launches {us:10.1f} us {us/tot:6.1%}")
print("\nlaunch by launch (us):")
for n, us in last:
    print(f"  {n:28s} {us:9.1f}")
