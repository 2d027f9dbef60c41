cd /root/repo
T=tools/tc_gemm_test
for args in "4096 148 0 256 6 2 0 0 0 0 0" "4096 148 0 256 22 2 0 0 0 0 0"; do
  timeout 60 $T $args | cut -c1-130,220-400; echo "rc=$?"
done
