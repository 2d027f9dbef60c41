set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -x -q -m gpu > gpurun_out/r02bg_pytest.txt 2>&1; tail -3 gpurun_out/r02bg_pytest.txt
timeout 200 python tools/time_configs.py c4inv c4ker > gpurun_out/r02bg_cfg.txt 2>&1; cat gpurun_out/r02bg_cfg.txt
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_tile_inv" -s 2 -c 1 --csv --log-file gpurun_out/r02bg_ncu_dram_c4inv.csv python tools/time_configs.py c4inv > /dev/null 2>&1
grep -o '"dram__bytes[^"]*","[^"]*","[^"]*"\|"gpu__time[^"]*","[^"]*","[^"]*"' gpurun_out/r02bg_ncu_dram_c4inv.csv
