#!/bin/bash
# first-contact GPU run: each group in its own process + timeout so one hang / fault does not hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_smi.txt 2>&1
for grp in glue attn gemm:f16 gemm:int8 gemm:int4 gemm:big; do
  f=gpurun_out/probe_${grp/:/_}.log
  timeout -s KILL ${PROBE_TIMEOUT:-240} python tests/gpu_probe.py $grp > $f 2>&1
  echo "== $grp exit=$? ==" | tee -a $f
  grep -E "PASS|FAIL|SUMMARY|Error|error" $f | head -60
done
