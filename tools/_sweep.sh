set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02az_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02az_pytest_gpu.txt
timeout 300 python bench.py --workload c1 --no-cpu > gpurun_out/r02az_bench_c1.json 2> gpurun_out/r02az_bench_c1.err; tail -3 gpurun_out/r02az_bench_c1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02az_bench_c1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['gpu_launches'], d['steps'], d['roofline'].get('launch'), d['roofline']['kernel_launches_per_step'], d['e2e']['value'])
"
