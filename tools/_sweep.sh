set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "in_place or register_tiled or c4 or prime_count or edge_inverse" > gpurun_out/r02ap_pytest_inv.txt 2>&1; tail -8 gpurun_out/r02ap_pytest_inv.txt
{
for s in 0 1 2 3; do echo "== LSX_TILE_INV_SHAPE=$s (0: pair steps 5 CTAs, 1: single steps, 2: pair 4 CTAs, 3: pair 6 CTAs)"; LSX_TILE_INV_SHAPE=$s python tools/time_configs.py c4inv; done
} > gpurun_out/r02ap_c4inv_pair.txt 2>&1
cat gpurun_out/r02ap_c4inv_pair.txt
