set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ax_smoke.txt 2>&1; tail -2 gpurun_out/r02ax_smoke.txt
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02ax_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02ax_pytest_gpu.txt
nvidia-smi --query-gpu=memory.used --format=csv
