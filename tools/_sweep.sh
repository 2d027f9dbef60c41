run() { K=$1; shift; env "$@" python tools/sweep_c5.py 4096 $K >> gpurun_out/r02ab_c5.jsonl 2>> gpurun_out/r02ab_c5.err; }
run 1013 LSX_TAG=auto_1013
run 1013 LSX_TC_COAL=0 LSX_TAG=rowmap_1013
run 1013 LSX_TC_COAL=1 LSX_TAG=coal_1013
run 127 LSX_TAG=auto_127
run 127 LSX_TC_COAL=0 LSX_TAG=rowmap_127
run 127 LSX_TC_COAL=1 LSX_TAG=coal_127
cat gpurun_out/r02ab_c5.jsonl; tail -3 gpurun_out/r02ab_c5.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
LSX_LARGE_STREAMS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02ab_launches_c5_127.csv python tools/sweep_c5.py 4096 127 > gpurun_out/r02ab_ncu.log 2>&1
tail -2 gpurun_out/r02ab_ncu.log
