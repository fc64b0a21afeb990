#!/bin/bash
# single-GPU validation of everything new + BASELINE config bench lines (C1 FP16, C2 INT8, C3 INT4 batch sweep @ ctx 4096)
mkdir -p gpurun_out
( timeout -s KILL 900 python -m pytest tests -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
( timeout -s KILL 120 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
( timeout -s KILL 300 python bench.py ) > gpurun_out/bench.log 2>&1; echo "bench exit=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
: > gpurun_out/bench_configs.jsonl
for q in f16 int8; do ( timeout -s KILL 240 python bench.py --quant $q --no-cpu-baseline --steps 30 --warmup 5 ) 2>gpurun_out/bench_$q.err | tail -1 >> gpurun_out/bench_configs.jsonl; done
for b in 1 2 4 8 16 32 64; do ( timeout -s KILL 240 python bench.py --batch $b --ctx 4096 --no-cpu-baseline --steps 30 --warmup 5 ) 2>gpurun_out/bench_b$b.err | tail -1 >> gpurun_out/bench_configs.jsonl; done
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.jsonl"):
    try:
        d = json.loads(l)
    except Exception:
        print("bad line", l[:100]); continue
    print("%-60s %8.0f tok/s %7.3f ms step_frac %.3f attn_frac %.3f e2e %8.0f" % (d["config"]["workload"], d["value"], d["ms_per_step"], d["step_roofline"]["frac"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
( timeout -s KILL 200 python tools/kernel_bench.py attn ) > gpurun_out/kernel_bench_attn.log 2>&1; echo "kbench exit=$?"; head -12 gpurun_out/kernel_bench_attn.log
