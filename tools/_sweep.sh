set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02as_pytest.txt 2>&1; tail -4 gpurun_out/r02as_pytest.txt
timeout 200 python tools/time_configs.py c4inv c4ker > gpurun_out/r02as_c4.txt 2>&1; cat gpurun_out/r02as_c4.txt
