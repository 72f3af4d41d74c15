#!/bin/bash
# Summaries of the ncu captures of tools/evidence_r2.sh (runs on the GPU box, where the .ncu-rep files are):
#   tools/evidence_r2_post.sh REP_DIR OUT_DIR
set -e
R=${1:-/tmp/bpt_ncu}; P=${2:-gpurun_out/profiles}; O=gpurun_out
mkdir -p $P
python tools/ncu_summary.py $R/r2_trace_c2.ncu-rep > $P/r2_trace_closest_shadow_c2.txt
python tools/ncu_summary.py $R/r2_trace_c2_ground.ncu-rep >> $P/r2_trace_closest_shadow_c2.txt
python tools/trace_metrics.py $P/r2_trace_metrics.json $R/r2_trace_c2.ncu-rep $R/r2_trace_c2_ground.ncu-rep
python tools/ncu_summary.py $R/r2_merged_tail_shade_w8.ncu-rep > $P/r2_merged_tail_shade_raygen_splat_world8.txt
python tools/ncu_summary.py $R/r2_shade_c2.ncu-rep > $P/r2_shade_raygen_splat_c2.txt
python tools/ncu_summary.py $R/r2_shade_c2_ground.ncu-rep >> $P/r2_shade_raygen_splat_c2.txt
python tools/trace_traffic.py $O/r2_trace_dram_per_launch.csv $P/trace_dram_bytes.json
cp $O/r2_trace_dram_per_launch.csv $O/r2_launches_bench.csv $O/r2_launches_world8_share.csv $O/r2_pass_period.log $P/
for c in c3 c4; do [ -s $O/r2_bench_${c}_n1.json ] && cp $O/r2_bench_${c}_n1.json $P/; done
echo done
