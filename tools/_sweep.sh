python tools/time_configs.py c1 c3 > gpurun_out/r02z_c3.txt 2>&1
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/r02z_c3.txt
cat gpurun_out/r02z_c3.txt
