set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_guards.py -x -q -m gpu > gpurun_out/r02aq_pytest.txt 2>&1; tail -5 gpurun_out/r02aq_pytest.txt
{
echo "== default (rhs out of the tile, shared-memory Garner tables)"; python tools/time_configs.py c4inv c4ker
echo "== LSX_TILE_RHS_IN=1 (kernel basis in the five-block tile)"; LSX_TILE_RHS_IN=1 python tools/time_configs.py c4ker
} > gpurun_out/r02aq_c4.txt 2>&1
cat gpurun_out/r02aq_c4.txt
