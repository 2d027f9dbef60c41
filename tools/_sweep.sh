( time python bench.py ) > gpurun_out/r02ag_bench_1gpu.json 2> gpurun_out/r02ag_bench_1gpu.err
tail -c 600 gpurun_out/r02ag_bench_1gpu.json; tail -5 gpurun_out/r02ag_bench_1gpu.err
