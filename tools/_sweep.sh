set -x
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 ) > gpurun_out/r02aw_bench_8gpu.json 2> gpurun_out/r02aw_bench_8gpu.err
tail -c 300 gpurun_out/r02aw_bench_8gpu.json; tail -5 gpurun_out/r02aw_bench_8gpu.err
