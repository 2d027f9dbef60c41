set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02au_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02au_pytest_gpu.txt
timeout 200 python tools/time_configs.py c1 > gpurun_out/r02au_c1.txt 2>&1; cat gpurun_out/r02au_c1.txt
timeout 300 python bench.py --workload c1 --no-cpu > gpurun_out/r02au_bench_c1.json 2> gpurun_out/r02au_bench_c1.err; tail -c 600 gpurun_out/r02au_bench_c1.json
