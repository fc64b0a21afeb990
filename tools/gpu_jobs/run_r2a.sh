#!/bin/bash
# round-2 first GPU job: state check + head-to-head vs the reference's Blackwell attention kernel
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/pytest_gpu.log
( timeout -s KILL 900 python tools/attn_vs_trtllm.py ) > gpurun_out/attn_vs_trtllm.log 2>&1; echo "h2h exit=$?"; grep -E "^B|NVIDIA" gpurun_out/attn_vs_trtllm.log | cut -c1-400
( timeout -s KILL 300 python tools/kernel_bench.py gemm ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kernel_bench.log
