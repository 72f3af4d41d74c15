set -u
mkdir -p gpurun_out
O=gpurun_out/x20.log
: > $O
P="python tools/profile_pass.py"
for e in "BPT_X=0" "BPT_MERGE_MAX_SLOTS=200000000" "BPT_TAIL_THRESHOLD=262144" "BPT_TAIL_THRESHOLD=1048576" "BPT_TAIL_THRESHOLD=0" "BPT_TAIL_THRESHOLD=16384"; do
  echo "=== c2 $e" | tee -a $O
  env $e timeout 300 $P --config c2 --passes 12 --no-detail 2>&1 | grep total_ms | cut -c1-120 | tee -a $O
done
for e in "BPT_X=0" "BPT_TAIL_THRESHOLD=262144" "BPT_MERGE_MAX_SLOTS=200000000"; do
  echo "=== c3 $e" | tee -a $O
  env $e timeout 300 $P --config c3 --spp 64 --passes 8 --no-detail 2>&1 | grep total_ms | cut -c1-120 | tee -a $O
done
