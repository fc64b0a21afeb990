#!/bin/bash
# r02 8-GPU job: collectives at world 8, TP8 headline bench, per-op timeline at TP8, BASELINE config C4 (Qwen2-72B INT4 TP8 B16 ctx8192)
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout -s KILL 240 $TR --master-port 29541 tools/tp_collectives_check.py ) > gpurun_out/tp8_collectives.log 2>&1; echo "collectives exit=$?"; grep -cE "^\[PASS\]" gpurun_out/tp8_collectives.log; grep -E "^\[FAIL\]" gpurun_out/tp8_collectives.log | head
( timeout -s KILL 300 $TR --master-port 29542 bench.py --gpus $N --steps 30 --warmup 5 ) > gpurun_out/bench_tp8.log 2>&1; echo "bench tp8 exit=$?"; grep '"metric"' gpurun_out/bench_tp8.log | cut -c1-330
( timeout -s KILL 300 $TR --master-port 29543 tools/step_timeline.py ) > gpurun_out/tp8_timeline.log 2>&1; echo "timeline exit=$?"; grep -vE "^W|Warning|warn" gpurun_out/tp8_timeline.log | head -20
( timeout -s KILL 600 $TR --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 --model qwen2-72b --batch 16 --ctx 8192 ) > gpurun_out/bench_tp8_qwen72b.log 2>&1; echo "bench qwen exit=$?"; grep '"metric"' gpurun_out/bench_tp8_qwen72b.log | cut -c1-330; tail -3 gpurun_out/bench_tp8_qwen72b.log | cut -c1-200
( timeout -s KILL 300 $TR --master-port 29545 bench.py --gpus $N --steps 30 --warmup 5 --comm nccl ) > gpurun_out/bench_tp8_nccl.log 2>&1; echo "bench tp8 nccl exit=$?"; grep '"metric"' gpurun_out/bench_tp8_nccl.log | cut -c1-200
