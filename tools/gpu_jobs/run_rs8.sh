#!/bin/bash
# developer: TP8 A/B of the fused GEMM + reduce-scatter path (headline config), then C4
mkdir -p gpurun_out
run() {  # $1 = fuse flag, rest = bench args
  f=$1; shift
  B200_FUSE_GEMM_RS=$f timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline "$@" 2>>gpurun_out/rs8_err.log | grep '^{' | tee -a gpurun_out/rs8_lines.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('fuse_gemm_rs=$f', d['config']['workload'], d['ms_per_step'], d['value'], d['launches_per_step'], d['parity_check']['ok'], d['config'].get('gemm_reduce_scatter_fused'))
" >> gpurun_out/rs8_ab.txt
}
run 1 --steps 60 --warmup 10
run 0 --steps 60 --warmup 10
run 1 --steps 60 --warmup 10
run 1 --model qwen2-72b --batch 16 --ctx 8192 --steps 30 --warmup 5
cat gpurun_out/rs8_ab.txt
