#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_api.py tests/test_gpu_decode_step.py tests/test_pybind_module.py -q -x ) > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit=$?"; tail -6 gpurun_out/pytest_attn.log | cut -c1-300
( timeout -s KILL 600 python tools/attn_vs_trtllm.py ) > gpurun_out/attn_vs_trtllm.log 2>&1; echo "h2h exit=$?"; grep -E "^B|NVIDIA" gpurun_out/attn_vs_trtllm.log | cut -c1-260
( timeout -s KILL 200 python tools/kernel_bench.py attn ) > gpurun_out/kernel_bench_attn.log 2>&1; echo "kbench exit=$?"; grep -E "attn B(1|16|32 Hq32 Hkv8 S2048 :|64)" gpurun_out/kernel_bench_attn.log
( timeout -s KILL 240 python bench.py --no-cpu-baseline --batch 1 --ctx 4096 --steps 30 --warmup 5 ) 2>/dev/null | tail -1 | cut -c1-200
( timeout -s KILL 240 python bench.py --no-cpu-baseline --batch 8 --ctx 4096 --steps 30 --warmup 5 ) 2>/dev/null | tail -1 | cut -c1-200
