#!/bin/bash
# Builds (here, no GPU needed) or runs (on the GPU box) the variants of tools/lab/inv8_lab.cu.
#   tools/lab/run_inv8_lab.sh build     -> tools/lab/bin/inv8_T<threads>_B<minb>_W<window>_C<experiment switch LSX_TPM_X>
#   tools/lab/run_inv8_lab.sh run OUT   -> one JSON line per variant appended to OUT
set -u
cd "$(dirname "$0")/../.."
VARIANTS="${VARIANTS:-128:4:1:1 128:4:1:0 64:8:1:1 64:9:1:1 64:9:1:0 96:6:1:1 64:10:1:1}"
if [ "$1" = build ]; then
  mkdir -p tools/lab/bin
  for v in $VARIANTS; do
    IFS=: read t b w c <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DLSX_TPM_THREADS=$t -DLSX_TPM_MINB=$b \
      -DLSX_TPM_WINDOW=$w -DLSX_TPM_X=$c -o tools/lab/bin/inv8_T${t}_B${b}_W${w}_C${c} tools/lab/inv8_lab.cu -ldl &
  done
  wait
else
  for v in $VARIANTS; do
    IFS=: read t b w c <<< "$v"
    timeout 120 tools/lab/bin/inv8_T${t}_B${b}_W${w}_C${c} 20 T${t}_B${b}_W${w}_C${c} >> "$2" 2>> "$2.err"
  done
fi
