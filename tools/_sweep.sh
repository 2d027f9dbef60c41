set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02ak_pytest_parity.txt 2>&1; tail -5 gpurun_out/r02ak_pytest_parity.txt
{
echo "== row-norm prime count (default)"; python tools/time_configs.py c3 c4inv c4ker
echo "== LSX_NO_DATA_BOUND=1"; LSX_NO_DATA_BOUND=1 python tools/time_configs.py c3
} > gpurun_out/r02ak_c3_c4.txt 2>&1
cat gpurun_out/r02ak_c3_c4.txt
