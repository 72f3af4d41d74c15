set -u
mkdir -p gpurun_out
O=gpurun_out/x1.log
: > $O
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) | tee -a $O
P="python tools/profile_pass.py"
for cfg in "--config c2" "--config c2 --world 8" "--config c4 --spp 64"; do
  echo "=== overlap $cfg" | tee -a $O
  timeout 300 $P $cfg --passes 6 --no-detail 2>&1 | grep total_ms | tee -a $O
  echo "=== sync-each $cfg" | tee -a $O
  timeout 300 $P $cfg --passes 6 --no-detail --sync-each 2>&1 | grep total_ms | tee -a $O
done
for rf in 8 12 20 24; do
  echo "=== refill $rf c2 detail" | tee -a $O
  BPT_REFILL=$rf timeout 300 $P --config c2 --passes 3 2>&1 | grep total_ms | tee -a $O
done
echo "=== refill 16 c2 detail" | tee -a $O
timeout 300 $P --config c2 --passes 3 2>&1 | grep total_ms | tee -a $O
