#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
if [ "${TP_CHECK:-1}" = "1" ]; then
( timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py ) > gpurun_out/tp_check_$N.log 2>&1; echo "tp_check exit=$?"; grep -E "PASS|FAIL" gpurun_out/tp_check_$N.log | sort -u | head
fi
for comm in ${COMMS:-peer nccl}; do
( timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --comm $comm ) > gpurun_out/bench_tp${N}_$comm.log 2>&1; echo "bench $comm exit=$?"; grep '"metric"' gpurun_out/bench_tp${N}_$comm.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  N=%d comm=%s value=%.0f tok/s ms=%.3f e2e=%.0f attn_frac=%.3f' % (d['n_gpus'], d['config']['tp_allreduce'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))"
done
