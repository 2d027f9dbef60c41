run() { K=$1; shift; env "$@" python tools/sweep_c5.py 4096 $K >> gpurun_out/r02ad_c5.jsonl 2>> gpurun_out/r02ad_c5.err; }
run 1013 LSX_TAG=pair_1013
run 1013 LSX_LU_PAIR=0 LSX_TAG=nopair_1013
run 127 LSX_TAG=pair_127
run 127 LSX_LU_PAIR=0 LSX_TAG=nopair_127
cat gpurun_out/r02ad_c5.jsonl; tail -3 gpurun_out/r02ad_c5.err
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
cap() { # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none -f -k regex:$rx --launch-skip $skip -c $cnt -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page details > gpurun_out/${name}.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv --metrics smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fmaheavy.sum > gpurun_out/${name}_raw.csv 2>&1
  rm -f /tmp/$name.ncu-rep
}
cap r02ad_c3_k_subwarp k_subwarp 1 1 python tools/time_configs.py c3
cap r02ad_c4inv 'k_tile_reg|k_assemble' 2 2 python tools/time_configs.py c4inv
cap r02ad_c4ker 'k_tile_reg|k_assemble' 2 2 python tools/time_configs.py c4ker
LSX_LARGE_STREAMS=1 cap r02ad_c5_k_lu8 k_lu8 200 2 python tools/sweep_c5.py 4096 127
LSX_LARGE_STREAMS=1 cap r02ad_c5_k_gemm_tc k_gemm_tc 0 22 python tools/sweep_c5.py 4096 127
du -sh gpurun_out
