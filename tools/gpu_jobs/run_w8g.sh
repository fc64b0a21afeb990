#!/bin/bash
# developer: GPU checks for the 8-bit group-wise GEMM variant
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode_step.py tests/test_gpu_loader.py tests/test_gpu_host_api.py -x -q -m gpu -k "gemm or step or loader or linear" > gpurun_out/w8g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/w8g_tests.log
tail -5 gpurun_out/w8g_tests.log
timeout 200 python bench.py --quant int8g --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/w8g_bench.log 2>&1
tail -2 gpurun_out/w8g_bench.log
