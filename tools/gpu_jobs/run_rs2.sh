#!/bin/bash
# developer: 2-GPU checks of the fused GEMM + reduce-scatter path, then a TP2 A/B
mkdir -p gpurun_out
echo skip tests


for f in 1 0 1; do
  B200_FUSE_GEMM_RS=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu-baseline 2>>gpurun_out/rs2_err.log | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('fuse_gemm_rs=$f', d['ms_per_step'], d['value'], d['launches_per_step'], d['parity_check'], d['config'].get('gemm_reduce_scatter_fused'))
" >> gpurun_out/rs2_ab.txt
done
cat gpurun_out/rs2_ab.txt
