set -x
( time timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload c1 --no-cpu ) > gpurun_out/r02bf_bench_c1_2gpu.json 2> gpurun_out/r02bf_bench_c1_2gpu.err
tail -c 700 gpurun_out/r02bf_bench_c1_2gpu.json; grep -i "capture\|error\|Traceback" gpurun_out/r02bf_bench_c1_2gpu.err | head
