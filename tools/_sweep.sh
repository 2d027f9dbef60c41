run() { K=$1; shift; env "$@" python tools/sweep_c5.py 4096 $K >> gpurun_out/r02af_c5.jsonl 2>> gpurun_out/r02af_c5.err; }
run 1013 LSX_TAG=trsm64_1013
run 1013 LSX_TRSM64=0 LSX_TAG=trsm32_1013
run 1013 LSX_TAG=trsm64_1013
run 1013 LSX_TRSM64=0 LSX_TAG=trsm32_1013
run 127 LSX_TAG=trsm64_127
run 127 LSX_TRSM64=0 LSX_TAG=trsm32_127
cat gpurun_out/r02af_c5.jsonl; tail -3 gpurun_out/r02af_c5.err
for mb in 1 5 6; do echo "MINB=$mb" >> gpurun_out/r02af_c3.txt; LSX_SW_MINB=$mb python tools/time_configs.py c1 c3 >> gpurun_out/r02af_c3.txt 2>&1; done
cat gpurun_out/r02af_c3.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
LSX_SW_MINB=5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
