#!/bin/bash
# 2-GPU job: peer collectives + TP step parity, TP2 bench lines
mkdir -p gpurun_out
( timeout -s KILL 500 python -m pytest tests/test_gpu_tp.py -q -x ) > gpurun_out/pytest_tp.log 2>&1; echo "pytest tp exit=$?"; tail -25 gpurun_out/pytest_tp.log | cut -c1-300
run_bench() { tag=$1; shift; ( timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 "$@" ) > gpurun_out/bench_tp2_$tag.log 2>&1; echo "bench $tag exit=$?"; grep '"metric"' gpurun_out/bench_tp2_$tag.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  N=%d comm=%s prog=%s value=%.0f tok/s ms=%.3f launches=%d parity=%s' % (d['n_gpus'], d['config']['tp_allreduce'], d['config']['decode_program'], d['value'], d['ms_per_step'], d['launches_per_step'], d.get('parity_check')))"; }
run_bench peer --comm peer
run_bench nccl --comm nccl
run_bench peer_prog --comm peer --program 1
