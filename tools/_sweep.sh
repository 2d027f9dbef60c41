set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02an_pytest_gpu.txt 2>&1; tail -5 gpurun_out/r02an_pytest_gpu.txt
( time python bench.py ) > gpurun_out/r02an_bench_1gpu.json 2> gpurun_out/r02an_bench_1gpu.err
tail -c 300 gpurun_out/r02an_bench_1gpu.json; tail -5 gpurun_out/r02an_bench_1gpu.err
python tools/time_configs.py c1 c3 c4inv c4ker > gpurun_out/r02an_time_configs.txt 2>&1; cat gpurun_out/r02an_time_configs.txt
