#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:paged_decode|wo_gemm|rmsnorm|rope_append|silu_and|embedding_k|argmax_k|convert_block' \
    -s ${NCU_SKIP:-1172} -c ${NCU_COUNT:-293} --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist exit=$?"
