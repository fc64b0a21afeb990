#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
tail -2 gpurun_out/final_smoke.log
timeout 100 python bench.py --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/final_bench.log 2>&1
tail -1 gpurun_out/final_bench.log | cut -c1-400
