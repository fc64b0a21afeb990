#!/bin/bash
# one GPU visit: smoke, gpu tests, per-kernel timings, the bench line
mkdir -p gpurun_out
( timeout -s KILL 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -3 gpurun_out/smoke.log
( timeout -s KILL 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -15 gpurun_out/pytest_gpu.log
( timeout -s KILL 600 python tools/kernel_bench.py ${KB_ARGS:-attn} ) > gpurun_out/kernel_bench.log 2>&1; echo "kbench exit=$?"; cat gpurun_out/kernel_bench.log | tail -60
( timeout -s KILL 900 python bench.py --steps ${STEPS:-30} --warmup 5 ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -2 gpurun_out/bench.log | cut -c1-1200
( timeout -s KILL 900 python bench.py --steps ${STEPS:-30} --warmup 5 --pdl 1 --no-cpu-baseline ) > gpurun_out/bench_pdl.log 2>&1; echo "bench pdl exit=$?"; tail -2 gpurun_out/bench_pdl.log | cut -c1-700
