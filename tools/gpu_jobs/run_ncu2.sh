#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_trace.py int4 4096 4096 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:wo_gemm -s 20 -c 1 -f -o gpurun_out/prof_gemm_int4_o \
    python tools/kernel_bench.py one_gemm int4 32 4096 4096 > gpurun_out/ncu_gemm_o.log 2>&1
echo "ncu exit=$?"
