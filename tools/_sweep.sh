set -x
( time timeout 900 python bench.py ) > gpurun_out/r02ba_bench_1gpu.json 2> gpurun_out/r02ba_bench_1gpu.err
tail -c 200 gpurun_out/r02ba_bench_1gpu.json; tail -4 gpurun_out/r02ba_bench_1gpu.err
