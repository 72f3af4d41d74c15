set -u
mkdir -p gpurun_out
O=gpurun_out/x4.log
: > $O
P="python tools/profile_pass.py"
for cfg in "--config c2 --world 8" "--config c2 --world 4" "--config c2 --world 2"; do
  for bb in 0 1 2; do
    echo "=== overlap bulk_bounce=$bb $cfg" | tee -a $O
    BPT_BULK_BOUNCE=$bb timeout 300 $P $cfg --passes 12 --no-detail 2>&1 | grep total_ms | tee -a $O
  done
  echo "=== overlap 2 batches per pass $cfg" | tee -a $O
  BPT_BACK_TO_BACK=0 timeout 300 $P $cfg --passes 12 --no-detail 2>&1 | grep total_ms | tee -a $O
  echo "=== sync-each $cfg" | tee -a $O
  timeout 300 $P $cfg --passes 6 --no-detail --sync-each 2>&1 | grep total_ms | tee -a $O
done
