#!/bin/bash
mkdir -p gpurun_out
export B200_DEBUG=1
( timeout -s KILL 120 python tools/kernel_bench.py one_gemm int4 32 4096 28672 ) > gpurun_out/kb_default.log 2>&1; echo "default: $?"; grep -E "b200|gemm" gpurun_out/kb_default.log
( B200_SEG_GRID=296 timeout -s KILL 120 python tools/kernel_bench.py gemm ) > gpurun_out/kb_296.log 2>&1; echo "grid296: $?"; grep -E "b200|gemm int" gpurun_out/kb_296.log | head -12
( B200_SEG_GRID=296 B200_SK_WHOLE_TILES=1 timeout -s KILL 120 python tools/kernel_bench.py one_gemm int4 32 4096 28672 ) > gpurun_out/kb_296w.log 2>&1; echo "grid296 whole: $?"; grep -E "gemm" gpurun_out/kb_296w.log
for mr in 4 8; do ( B200_SEG_GRID=296 B200_SK_MIN_RUN=$mr timeout -s KILL 120 python tools/kernel_bench.py one_gemm int4 32 4096 4096 ) > gpurun_out/kb_296_mr$mr.log 2>&1; echo "o-proj min_run $mr: $?"; grep -E "gemm" gpurun_out/kb_296_mr$mr.log; done
( B200_SEG_GRID=296 GRID=296 timeout -s KILL 120 python tools/program_trace.py ) > gpurun_out/program_trace.log 2>&1; echo "trace exit=$?"; tail -34 gpurun_out/program_trace.log
( B200_SEG_GRID=296 timeout -s KILL 200 python bench.py --no-cpu-baseline ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-260
( B200_SEG_GRID=296 timeout -s KILL 200 python -m pytest tests/test_gpu_decode_step.py -q -x ) > gpurun_out/pytest_step.log 2>&1; echo "pytest exit=$?"; tail -2 gpurun_out/pytest_step.log
