set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02at_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02at_pytest_gpu.txt
( time timeout 900 python bench.py ) > gpurun_out/r02at_bench_1gpu.json 2> gpurun_out/r02at_bench_1gpu.err
tail -c 200 gpurun_out/r02at_bench_1gpu.json; tail -4 gpurun_out/r02at_bench_1gpu.err
timeout 300 ncu --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:"k_tile_reg|k_verify|k_assemble" -s 6 -c 3 --csv --log-file gpurun_out/r02at_ncu_metrics_c4ker.csv python tools/time_configs.py c4ker > gpurun_out/r02at_ncu_c4ker.log 2>&1
tail -5 gpurun_out/r02at_ncu_metrics_c4ker.csv | cut -c1-300
