set -u
mkdir -p gpurun_out
O=gpurun_out/x3.log
: > $O
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) | tee -a $O
P="python tools/profile_pass.py"
for cfg in "--config c2 --world 8" "--config c2" "--config c4 --spp 64" "--config c3 --spp 64"; do
  for bb in 0 1 2; do
    echo "=== overlap bulk_bounce=$bb $cfg" | tee -a $O
    BPT_BULK_BOUNCE=$bb timeout 300 $P $cfg --passes 8 --no-detail 2>&1 | grep total_ms | tee -a $O
  done
  echo "=== sync-each $cfg" | tee -a $O
  timeout 300 $P $cfg --passes 6 --no-detail --sync-each 2>&1 | grep total_ms | tee -a $O
done
echo "=== world 4 / 2" | tee -a $O
timeout 300 $P --config c2 --world 4 --passes 8 --no-detail 2>&1 | grep total_ms | tee -a $O
timeout 300 $P --config c2 --world 2 --passes 8 --no-detail 2>&1 | grep total_ms | tee -a $O
timeout 600 python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>gpurun_out/x3_bench.err | tee gpurun_out/x3_bench.json | cut -c1-400
