#!/bin/bash
mkdir -p gpurun_out
export B200_LIB_PATH=$PWD/rtp_llm_b200/lib_dev.so
for shape in "4096 4096" "4096 6144" "14336 4096" "4096 28672"; do
  echo "=== int4 B32 K N = $shape"
  ( timeout -s KILL 100 python tools/gemm_trace.py int4 $shape ) 2>&1 | grep -vE "^ ?[0-9]+  " | grep -vE "^sm |^it " | head -12
done
