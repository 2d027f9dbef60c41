set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02aj_pytest_parity.txt 2>&1; tail -5 gpurun_out/r02aj_pytest_parity.txt
{
echo "== row-norm prime count (default)"; python tools/time_configs.py c4inv c4ker
echo "== LSX_NO_DATA_BOUND=1"; LSX_NO_DATA_BOUND=1 python tools/time_configs.py c4inv c4ker
} > gpurun_out/r02aj_c4_data_bound.txt 2>&1
cat gpurun_out/r02aj_c4_data_bound.txt
