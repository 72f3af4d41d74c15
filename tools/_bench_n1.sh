mkdir -p gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
cut -c1-250 gpurun_out/r2_bench_n1.json; cut -c1-250 gpurun_out/r2_bench_reference_n1.json
