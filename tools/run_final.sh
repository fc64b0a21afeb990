#!/bin/bash
# final validation + evidence for profiles/
mkdir -p gpurun_out
( timeout -s KILL 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
( timeout -s KILL 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu.log
( timeout -s KILL 900 python bench.py ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-1200
( timeout -s KILL 300 python bench.py --impl reference ) > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-700
( timeout -s KILL 600 python tools/kernel_bench.py attn gemm ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"
if [ "${WITH_NCU:-1}" = "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:paged_decode|wo_gemm|rmsnorm|rope_append|silu_and|embedding_k|argmax_k|convert_block' \
    -s 1044 -c 261 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pdl 0 > gpurun_out/bench_under_ncu.log 2>&1
echo "launchlist exit=$?"
ncu --set full --clock-control none --import-source on -k regex:paged_decode_attn -s 10 -c 1 -f -o gpurun_out/prof_attn_b32_s2048 \
    python tools/kernel_bench.py one_attn 32 32 8 2048 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit=$?"
ncu --set full --clock-control none --import-source on -k regex:wo_gemm -s 20 -c 1 -f -o gpurun_out/prof_gemm_int4_w13 \
    python tools/kernel_bench.py one_gemm int4 32 4096 28672 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
fi
