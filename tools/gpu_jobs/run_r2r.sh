#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode_step.py tests/test_gpu_host_api.py tests/test_pybind_module.py -q ) > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/pytest_attn.log | cut -c1-300
( timeout -s KILL 200 python tools/kernel_bench.py attn ) > gpurun_out/kernel_bench_attn.log 2>&1; grep -E "attn B(1|16|32 Hq32 Hkv8 S2048 :|64)" gpurun_out/kernel_bench_attn.log
