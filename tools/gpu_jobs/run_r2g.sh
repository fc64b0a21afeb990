#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 120 python tools/program_probe.py ) > gpurun_out/program_probe.log 2>&1; echo "probe exit=$?"; grep -E "^layers|Error|error" gpurun_out/program_probe.log | cut -c1-250 | tail -4
( timeout -s KILL 200 python tests/gpu_probe.py gemm:int4 gemm:int8 ) > gpurun_out/probe_gemm.log 2>&1; echo "gemm probe exit=$?"; grep -c PASS gpurun_out/probe_gemm.log; grep -E "FAIL|EXC" gpurun_out/probe_gemm.log | head -20
( timeout -s KILL 300 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -5 gpurun_out/pytest_gpu.log
( timeout -s KILL 200 python tools/kernel_bench.py gemm ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"; head -9 gpurun_out/kernel_bench.log
( timeout -s KILL 120 python tools/program_trace.py ) > gpurun_out/program_trace.log 2>&1; echo "trace exit=$?"; sed -n 1,20p gpurun_out/program_trace.log; grep -A40 "fine timeline" gpurun_out/program_trace.log | head -60
( timeout -s KILL 200 python bench.py --no-cpu-baseline ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-260
