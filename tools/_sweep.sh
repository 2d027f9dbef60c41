set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02bb_pytest_gpu.txt 2>&1; tail -6 gpurun_out/r02bb_pytest_gpu.txt
timeout 200 python tools/time_configs.py c3 > gpurun_out/r02bb_c3.txt 2>&1; cat gpurun_out/r02bb_c3.txt
