#!/bin/bash
# round-2: persistent segment kernel bring-up
mkdir -p gpurun_out
( timeout -s KILL 300 python tests/gpu_probe.py gemm:int4 gemm:int8 ) > gpurun_out/probe_gemm.log 2>&1; echo "probe exit=$?"; grep -c PASS gpurun_out/probe_gemm.log; grep -E "FAIL|EXC" gpurun_out/probe_gemm.log | head -20
( timeout -s KILL 600 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu.log
( timeout -s KILL 300 python tools/kernel_bench.py gemm ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kernel_bench.log
( timeout -s KILL 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-1500
