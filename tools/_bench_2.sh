bash tools/_bench_n.sh 2
(timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3) | tee gpurun_out/r2_pytest_multi.log
