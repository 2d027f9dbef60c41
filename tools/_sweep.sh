for v in base 128_1 256_0 256_1 512_1 512_2; do
  if [ $v = base ]; then L=linalg_solver_b200/liblsx.so; else L=tools/lab/so/liblsx_sw_$v.so; fi
  echo "== $v" >> gpurun_out/r02p_sw_variants.txt
  LSX_LIB_PATH=$PWD/$L python tools/time_configs.py c3 c1 >> gpurun_out/r02p_sw_variants.txt 2>&1
done
cat gpurun_out/r02p_sw_variants.txt
