#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 600 python tests/gpu_probe.py gemm:int4 gemm:int8 gemm:f16 ) > gpurun_out/probe_gemm.log 2>&1; echo "probe exit=$?"; grep -E "FAIL|SUMMARY" gpurun_out/probe_gemm.log | head
( timeout -s KILL 900 python tools/kernel_bench.py ${KB_ARGS:-gemm} ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kernel_bench.log | tail -70
