#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode_step.py tests/test_gpu_host_api.py -q -x ) > gpurun_out/pytest_rope.log 2>&1; echo "pytest exit=$?"; tail -8 gpurun_out/pytest_rope.log | cut -c1-300
( timeout -s KILL 300 python bench.py --no-cpu-baseline ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-330
( timeout -s KILL 200 python tools/kernel_bench.py attn ) > gpurun_out/kernel_bench_attn.log 2>&1; grep -E "attn B(1|16|32 Hq32 Hkv8 S2048 :|64)" gpurun_out/kernel_bench_attn.log
