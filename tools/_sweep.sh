set -x
for mb in 8192 2048; do
  for w in c4inv c4ker; do
    LSX_WS_MB=$mb timeout 300 python bench.py --workload $w --no-cpu --steps 5 > gpurun_out/r02av_ws${mb}_$w.json 2> gpurun_out/r02av_ws${mb}_$w.err
    python -c "
import json,sys
d=json.loads(open('gpurun_out/r02av_ws${mb}_$w.json').read().strip().splitlines()[-1])
print('ws_mb', $mb, '$w', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
" | tee -a gpurun_out/r02av_ws_sweep.txt
  done
done
