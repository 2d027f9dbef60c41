set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02bh_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r02bh_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
