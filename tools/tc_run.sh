cd /root/repo
T=tools/tc_gemm_test
for args in "256 1 0 64 0 1" "256 1 0 64 1 1" "512 2 64 64 0 2" "520 2 64 128 0 2" "1024 3 128 256 0 2" "4096 8 0 64 0 3" "4096 8 0 128 0 3" "4096 8 0 256 0 3" "4096 37 1024 256 0 3"; do
  timeout 60 $T $args; echo "rc=$?"
done
