echo "== ty8 default + static blocks" > gpurun_out/r02t_tile.txt
python tools/time_configs.py c4inv c4ker >> gpurun_out/r02t_tile.txt 2>&1
echo "== 256-thread shapes + static blocks" >> gpurun_out/r02t_tile.txt
LSX_TILE_TY8=0 python tools/time_configs.py c4inv c4ker >> gpurun_out/r02t_tile.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02t_tile.txt
LSX_TILE_TY8=0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02t_tile.txt
cat gpurun_out/r02t_tile.txt
