"""profiles/r2_trace_metrics.json from the `ncu --set full` captures of the traversal launches (tools/evidence_r2.sh step 2):
per launch the figures bench.py quotes in its roofline object (active lanes, issue-active), and their instruction-weighted means.
    python tools/trace_metrics.py OUT.json REP [REP ...]"""
import csv
import json
import subprocess
import sys

FIELDS = {
    "ms": "gpu__time_duration.sum",
    "active_lanes_of_32": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_rate_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "registers": "launch__registers_per_thread",
    "warp_instructions": "smsp__inst_executed.sum",
    "long_scoreboard_per_issue": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
}
launches = []
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", ""), "capture": rep.split("/")[-1]}
        for k, m in FIELDS.items():
            if m in hdr:
                v = float(r[hdr.index(m)].replace(",", ""))
                if k == "ms":
                    u = units[hdr.index(m)]
                    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
                d[k] = v
        local = 0.0
        for m in ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"):
            if m in hdr:
                local += float(r[hdr.index(m)].replace(",", ""))
        d["local_memory_sectors"] = local
        launches.append(d)
w = sum(l["warp_instructions"] for l in launches)
out = {"source": "ncu --set full of traversal launches of a C2 pass (tools/evidence_r2.sh step 2: the first four of the sky-half batch and the "
                 "first four of the ground / mesh half; profiles/r2_trace_closest_shadow_c2.txt)",
       "launches": launches,
       "active_lanes_of_32_weighted": sum(l["active_lanes_of_32"] * l["warp_instructions"] for l in launches) / w,
       "issue_active_pct_weighted": sum(l["issue_active_pct"] * l["warp_instructions"] for l in launches) / w}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps({k: out[k] for k in ("active_lanes_of_32_weighted", "issue_active_pct_weighted")}))
