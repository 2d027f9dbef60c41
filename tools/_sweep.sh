python tools/time_configs.py c4ker > gpurun_out/r02w_c4ker.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02w_c4ker.txt
LSX_DISABLE_SUBWARP=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matrix_api.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02w_c4ker.txt
cat gpurun_out/r02w_c4ker.txt
