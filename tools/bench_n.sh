#!/bin/bash
# bench.py on N GPUs of one box, launched the way the driver does (torchrun, one rank per GPU):  tools/bench_n.sh N  -> gpurun_out/r2_bench_nN.json
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
cut -c1-200 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
