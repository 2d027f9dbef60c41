( time timeout 900 python bench.py ) > gpurun_out/r02bi_bench_1gpu.json 2> gpurun_out/r02bi_bench_1gpu.err
tail -3 gpurun_out/r02bi_bench_1gpu.err
