set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "in_place" > gpurun_out/r02ai_pytest_inv.txt 2>&1; tail -5 gpurun_out/r02ai_pytest_inv.txt
ncu --set full --clock-control none --import-source on -k regex:k_tile_inv -s 2 -c 1 -o gpurun_out/r02ai_tile_inv python tools/time_configs.py c4inv > gpurun_out/r02ai_ncu.log 2>&1
tail -3 gpurun_out/r02ai_ncu.log
