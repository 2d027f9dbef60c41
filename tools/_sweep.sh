VARIANTS="128:4:1:0 128:5:1:0 128:6:1:0 64:8:1:0 64:10:1:0 64:12:1:0 96:6:1:0 96:7:1:0 256:2:1:0 128:4:0:0 128:4:2:0" bash tools/lab/run_inv8_lab.sh run gpurun_out/r02o_lab_variants.jsonl
grep bareiss gpurun_out/r02o_lab_variants.jsonl
tail -3 gpurun_out/r02o_lab_variants.jsonl.err
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_inv_tpmILi8ELi4ELb0ELi3 -s 3 -c 1 -f -o gpurun_out/r02o_inv8_bareiss tools/lab/bin/inv8_bareiss 5 ncu > gpurun_out/r02o_ncu.log 2>&1
tail -3 gpurun_out/r02o_ncu.log
