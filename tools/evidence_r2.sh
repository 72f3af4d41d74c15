#!/bin/bash
# Round-2 measurement evidence, one gpurun call (1 GPU).  Every ncu run is preceded by the same command without ncu.
#   tools/evidence_r2.sh          -> gpurun_out/r2_* (logs, csv lists) and gpurun_out/profiles/ (the summaries of the ncu captures, made on the
#   box by tools/evidence_r2_post.sh: the .ncu-rep files themselves stay in $R on the box -- together they exceed what gpurun brings back)
set -u
mkdir -p gpurun_out
O=gpurun_out
R=/tmp/bpt_ncu
mkdir -p $R
P="python tools/profile_pass.py"
# 1. every launch of a bench run with its device time
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $B > $O/r2_bench_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_ncu_launches.log 2>&1
# 2. ncu --set full of the dominant traversal launches (C2, full frame; one pass = two batches, enqueued one after the other:
#    the sky half first, then the ground / mesh half): bounce 0 / 1 closest hit and their shadow launches of both
C="$P --config c2 --passes 1 --no-detail"
timeout 200 $C > $O/r2_plain_c2.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 4 -f -o $R/r2_trace_c2 $C > $O/r2_ncu_trace.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace --launch-skip 24 -c 4 -f -o $R/r2_trace_c2_ground $C > $O/r2_ncu_trace_ground.log 2>&1
# 3. the merged traversal launch, the fused tail and the shade / raygen / splat kernels on an 8-rank share (small batches)
W="$P --config c2 --world 8 --passes 1 --no-detail"
timeout 200 $W > $O/r2_plain_w8.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_trace_merged|k_tail|k_shade|k_raygen|k_splat" -c 12 -f -o $R/r2_merged_tail_shade_w8 $W > $O/r2_ncu_w8.log 2>&1
# 3b. launch timeline of the same share, and the pass period of back-to-back passes (no ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_world8_share.csv $P --config c2 --world 8 --passes 2 --no-detail --sync-each > $O/r2_ncu_w8_launches.log 2>&1
for w in 8 4 2 1; do
  echo "=== world $w: back-to-back passes, then a host wait after every pass" >> $O/r2_pass_period.log
  timeout 300 $P --config c2 --world $w --passes 12 --no-detail 2>&1 | grep total_ms >> $O/r2_pass_period.log
  timeout 300 $P --config c2 --world $w --passes 6 --no-detail --sync-each 2>&1 | grep total_ms >> $O/r2_pass_period.log
done
# 4. DRAM / L2 bytes of every traversal launch of one C2 pass (roofline.traffic)
D="$P --config c2 --passes 2 --no-detail --sync-each"
timeout 200 $D > $O/r2_plain_c2x2.log 2>&1 && \
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k "regex:k_trace|k_tail" --csv --log-file $O/r2_trace_dram_per_launch.csv $D > $O/r2_ncu_dram.log 2>&1
# 5. shade / raygen / splat of the full frame: the sky half's first launches, and the ground / mesh half's first two k_shade
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_shade|k_raygen|k_splat" -c 4 -f -o $R/r2_shade_c2 $C > $O/r2_ncu_shade.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade --launch-skip 12 -c 2 -f -o $R/r2_shade_c2_ground $C > $O/r2_ncu_shade_ground.log 2>&1
# 6. per-stage times of the other BASELINE configurations (CUDA events, no ncu)
for c in c3 c4; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/r2_bench_${c}_n1.json 2> $O/r2_bench_${c}_n1.err
done
bash tools/evidence_r2_post.sh $R $O/profiles
cp $R/r2_trace_c2_ground.ncu-rep $O/ 2>/dev/null       # one capture comes back for source-level reading (~20 MB)
ls -la $O/r2_* $O/profiles | head -60
