#!/bin/bash
mkdir -p gpurun_out
export B200_DEBUG=1
( timeout -s KILL 120 python tools/program_trace.py ) > gpurun_out/program_trace.log 2>&1; echo "trace exit=$?"; grep -A80 "fine timeline" gpurun_out/program_trace.log | head -150
