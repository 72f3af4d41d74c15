set -u
mkdir -p gpurun_out
O=gpurun_out/x5.log
: > $O
P="python tools/profile_pass.py"
for cfg in "--config c2 --world 8" "--config c2" "--config c2 --world 2"; do
  for st in 0 1 2; do
    echo "=== overlap stagger=$st bb=0 $cfg" | tee -a $O
    BPT_STAGGER=$st BPT_BULK_BOUNCE=0 timeout 300 $P $cfg --passes 12 --no-detail 2>&1 | grep total_ms | tee -a $O
  done
  echo "=== sync-each stagger=1 $cfg" | tee -a $O
  timeout 300 $P $cfg --passes 6 --no-detail --sync-each 2>&1 | grep total_ms | tee -a $O
done
