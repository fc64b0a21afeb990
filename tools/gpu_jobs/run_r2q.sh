#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
( timeout -s KILL 200 python tools/kernel_bench.py attn ) > gpurun_out/kernel_bench_attn.log 2>&1; grep -E "attn B(1|16|32 Hq32 Hkv8 S2048 :|64)" gpurun_out/kernel_bench_attn.log
( timeout -s KILL 300 python bench.py --no-cpu-baseline ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-260
