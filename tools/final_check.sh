#!/bin/bash
# What the driver runs at round end, in one gpurun call: GPU tests, smoke(), the default bench line.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) | tee gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
cut -c1-220 gpurun_out/r2_bench_n1.json
