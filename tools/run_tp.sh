#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py ) > gpurun_out/tp_check_$N.log 2>&1; echo "tp_check exit=$?"; grep -E "PASS|FAIL|Error|error" gpurun_out/tp_check_$N.log | head; tail -3 gpurun_out/tp_check_$N.log
( timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --pdl 1 ) > gpurun_out/bench_tp$N.log 2>&1; echo "bench exit=$?"; tail -2 gpurun_out/bench_tp$N.log | cut -c1-900
( timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --pdl 1 --comm nccl ) > gpurun_out/bench_tp${N}_nccl.log 2>&1; echo "bench nccl exit=$?"; tail -2 gpurun_out/bench_tp${N}_nccl.log | cut -c1-400
