set -x
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02ao_pytest_multi_2gpu.txt 2>&1; tail -5 gpurun_out/r02ao_pytest_multi_2gpu.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 ) > gpurun_out/r02ao_bench_2gpu.json 2> gpurun_out/r02ao_bench_2gpu.err
tail -c 300 gpurun_out/r02ao_bench_2gpu.json; tail -5 gpurun_out/r02ao_bench_2gpu.err
