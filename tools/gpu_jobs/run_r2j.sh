#!/bin/bash
# r02 single-GPU evidence: reference-kernel goldens (flashinfer glue ops), ncu launch list of the default step, ncu --set full of
# the attention kernel, the cluster GEMM (w13) and the persistent kernel (w13 as a one-op program)
mkdir -p gpurun_out
( timeout -s KILL 400 python tools/make_gpu_golden.py ) > gpurun_out/make_gpu_golden.log 2>&1; echo "golden exit=$?"; tail -2 gpurun_out/make_gpu_golden.log | cut -c1-300
( timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:paged_decode|wo_gemm|rmsnorm|rope_append|silu_and|embedding_k|argmax_k|convert_block|decode_segment' \
    -s 1044 -c 261 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pdl 0 ) > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist exit=$?"
( timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:paged_decode_attn -s 10 -c 1 -f -o gpurun_out/r02_prof_attn_b32_s2048 \
    python tools/kernel_bench.py one_attn 32 32 8 2048 ) > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit=$?"
( timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:wo_gemm -s 20 -c 1 -f -o gpurun_out/r02_prof_gemm_cluster_w13 \
    python tools/kernel_bench.py one_gemm int4 32 4096 28672 ) > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
( B200_GEMM_PERSISTENT=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:decode_segment -s 20 -c 1 -f -o gpurun_out/r02_prof_gemm_persistent_w13 \
    python tools/kernel_bench.py one_gemm int4 32 4096 28672 ) > gpurun_out/ncu_gemm_p.log 2>&1; echo "ncu persistent exit=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.npz 2>/dev/null
