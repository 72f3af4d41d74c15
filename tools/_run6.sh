set -u
mkdir -p gpurun_out
O=gpurun_out/x6.log
: > $O
(timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_async.py -x -q 2>&1 | tail -8) | tee -a $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/x6_bench_n2.err | tee gpurun_out/x6_bench_n2.json | cut -c1-300
P="python tools/profile_pass.py"
for np_ in 2 3 4; do
  echo "=== pipes=$np_ world 8" | tee -a $O
  BPT_PIPES=$np_ timeout 300 $P --config c2 --world 8 --passes 12 --no-detail 2>&1 | grep total_ms | cut -c1-40 | tee -a $O
done
