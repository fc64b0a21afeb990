#!/bin/bash
mkdir -p gpurun_out
( timeout -s KILL 300 python tools/program_probe.py ) > gpurun_out/program_probe.log 2>&1; echo "probe exit=$?"; grep -E "^layers|Error|error" gpurun_out/program_probe.log | cut -c1-400
( timeout -s KILL 300 python tools/kernel_bench.py gemm ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kernel_bench.log
