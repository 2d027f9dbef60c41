cd /root/repo
T=tools/tc_gemm_test
for args in "520 2 64 32 0 2" "520 2 64 128 0 2" "1000 3 128 256 0 2" "4096 37 1024 256 0 3" "4096 148 0 256 0 2" "4096 148 0 64 0 2"; do
  timeout 60 $T $args; echo "rc=$?"
done
