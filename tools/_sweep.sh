timeout 300 python tools/graph_probe2.py > gpurun_out/r02bc_graph_probe2.txt 2>&1; tail -12 gpurun_out/r02bc_graph_probe2.txt
