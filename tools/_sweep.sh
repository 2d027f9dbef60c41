python tools/time_configs.py c4inv c4ker > gpurun_out/r02x_tile.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02x_tile.txt
LSX_DISABLE_SUBWARP=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02x_tile.txt
cat gpurun_out/r02x_tile.txt
