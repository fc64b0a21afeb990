#!/bin/bash
for n in prod half; do
  if [ $n = prod ]; then unset B200_LIB_PATH; else export B200_LIB_PATH=$PWD/rtp_llm_b200/lib_$n.so; fi
  echo "=== $n"
  timeout 300 python tests/gpu_probe.py gemm:int4 gemm:int8 2>&1 | grep -E "SUMMARY|FAIL" | head -5
  timeout 300 python tools/kernel_bench.py gemm 2>&1 | grep -E "int4 B32 K4096 N6144 :|int4 B32 K4096 N4096 :|int4 B32 K4096 N28672 :|int4 B32 K14336|int8 B32"
done
