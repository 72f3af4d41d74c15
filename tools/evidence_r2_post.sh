#!/bin/bash
# Run HERE after tools/evidence_r2.sh came back: summaries of the ncu captures -> profiles/ (tracked).
set -e
O=gpurun_out
python tools/ncu_summary.py $O/r2_trace_c2.ncu-rep > profiles/r2_trace_closest_shadow_c2.txt
python tools/ncu_summary.py $O/r2_trace_c2_ground.ncu-rep >> profiles/r2_trace_closest_shadow_c2.txt
python tools/trace_metrics.py profiles/r2_trace_metrics.json $O/r2_trace_c2.ncu-rep $O/r2_trace_c2_ground.ncu-rep
python tools/ncu_summary.py $O/r2_merged_tail_shade_w8.ncu-rep > profiles/r2_merged_tail_shade_raygen_splat_world8.txt
python tools/ncu_summary.py $O/r2_shade_c2.ncu-rep > profiles/r2_shade_raygen_splat_c2.txt
python tools/ncu_summary.py $O/r2_shade_c2_ground.ncu-rep >> profiles/r2_shade_raygen_splat_c2.txt
python tools/trace_traffic.py $O/r2_trace_dram_per_launch.csv profiles/trace_dram_bytes.json
cp $O/r2_trace_dram_per_launch.csv $O/r2_launches_bench.csv $O/r2_launches_world8_share.csv $O/r2_pass_period.log profiles/
for c in c1 c3 c4; do [ -s $O/r2_bench_${c}_n1.json ] && cp $O/r2_bench_${c}_n1.json profiles/; done
echo done
