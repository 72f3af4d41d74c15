#!/bin/bash
# Round-2 measurement evidence, one gpurun call (1 GPU).  Every ncu run is preceded by the same command without ncu.
#   tools/evidence_r2.sh          -> gpurun_out/r2_*
set -u
mkdir -p gpurun_out
O=gpurun_out
P="python tools/profile_pass.py"
# 1. every launch of a bench run with its device time
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $B > $O/r2_bench_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_ncu_launches.log 2>&1
# 2. ncu --set full of the dominant traversal launches (C2, full frame: bounce 0 / 1 closest hit and their shadow launches)
C="$P --config c2 --passes 1 --no-detail"
timeout 200 $C > $O/r2_plain_c2.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 4 -f -o $O/r2_trace_c2 $C > $O/r2_ncu_trace.log 2>&1
# 3. the merged traversal launch, the fused tail and the shade / raygen / splat kernels on an 8-rank share (small batches)
W="$P --config c2 --world 8 --passes 1 --no-detail"
timeout 200 $W > $O/r2_plain_w8.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_trace_merged|k_tail|k_shade|k_raygen|k_splat" -c 12 -f -o $O/r2_merged_tail_shade_w8 $W > $O/r2_ncu_w8.log 2>&1
# 4. DRAM / L2 bytes of every traversal launch of one C2 pass (roofline.traffic)
D="$P --config c2 --passes 2 --no-detail"
timeout 200 $D > $O/r2_plain_c2x2.log 2>&1 && \
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k "regex:k_trace|k_tail" --csv --log-file $O/r2_trace_dram_per_launch.csv $D > $O/r2_ncu_dram.log 2>&1
# 5. shade kernel of the full frame (bounce 0 and 1)
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_shade|k_raygen|k_splat" -c 4 -f -o $O/r2_shade_c2 $C > $O/r2_ncu_shade.log 2>&1
ls -la $O/r2_* | head -30
