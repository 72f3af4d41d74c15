set -u
mkdir -p gpurun_out
O=gpurun_out
P="python tools/profile_pass.py"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 300 $B > $O/r2_bench_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_bench.csv $B > $O/r2_ncu_launches.log 2>&1
D="$P --config c2 --passes 2 --no-detail --sync-each"
timeout 200 $D > $O/r2_plain_c2x2.log 2>&1 && \
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k "regex:k_trace|k_tail" --csv --log-file $O/r2_trace_dram_per_launch.csv $D > $O/r2_ncu_dram.log 2>&1
python tools/trace_traffic.py $O/r2_trace_dram_per_launch.csv $O/trace_dram_bytes.json
rm -f $O/r2_pass_period.log
for w in 8 4 2 1; do
  echo "=== world $w: back-to-back passes, then a host wait after every pass" >> $O/r2_pass_period.log
  timeout 300 $P --config c2 --world $w --passes 12 --no-detail 2>&1 | grep total_ms >> $O/r2_pass_period.log
  timeout 300 $P --config c2 --world $w --passes 6 --no-detail --sync-each 2>&1 | grep total_ms >> $O/r2_pass_period.log
done
cut -c1-60 $O/r2_pass_period.log
python -c "import __graft_entry__ as g; g.smoke()"
