#!/bin/bash
# r02 final 8-GPU job: TP step parity vs the unsharded oracle at world 8, TP8 bench (default / rope fused into attention), C4
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout -s KILL 300 $TR --master-port 29551 tools/tp_check.py ) > gpurun_out/tp8_check.log 2>&1; echo "tp_check exit=$?"; grep -E "^\[(PASS|FAIL)\]" gpurun_out/tp8_check.log
( timeout -s KILL 300 $TR --master-port 29552 bench.py --gpus $N --steps 40 --warmup 5 ) > gpurun_out/bench_tp8.log 2>&1; echo "bench tp8 exit=$?"; grep '"metric"' gpurun_out/bench_tp8.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  TP8 default: %.0f tok/s %.3f ms launches %d parity %s' % (d['value'], d['ms_per_step'], d['launches_per_step'], d.get('parity_check')))"
( B200_FUSE_ROPE=1 timeout -s KILL 300 $TR --master-port 29553 bench.py --gpus $N --steps 40 --warmup 5 ) > gpurun_out/bench_tp8_fuserope.log 2>&1; echo "bench tp8 fuse_rope exit=$?"; grep '"metric"' gpurun_out/bench_tp8_fuserope.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  TP8 rope fused: %.0f tok/s %.3f ms launches %d' % (d['value'], d['ms_per_step'], d['launches_per_step']))"
( timeout -s KILL 600 $TR --master-port 29554 bench.py --gpus $N --steps 20 --warmup 3 --model qwen2-72b --batch 16 --ctx 8192 ) > gpurun_out/bench_tp8_qwen72b.log 2>&1; echo "bench qwen exit=$?"; grep '"metric"' gpurun_out/bench_tp8_qwen72b.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  C4 Qwen2-72B TP8: %.0f tok/s %.3f ms parity %s' % (d['value'], d['ms_per_step'], d.get('parity_check')))"
