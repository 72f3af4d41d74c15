mkdir -p gpurun_out
P="python tools/profile_pass.py"
BPT_DUMP_SPANS=1 timeout 300 $P --config c2 --world 8 --passes 3 > gpurun_out/x2_w8.log 2>&1
BPT_DUMP_SPANS=1 timeout 300 $P --config c2 --passes 3 > gpurun_out/x2_n1.log 2>&1
tail -3 gpurun_out/x2_w8.log
