#!/bin/bash
for rep in 1 2; do for f in 0 1; do
  B200_FUSE_ROPE=$f timeout -s KILL 200 python bench.py --no-cpu-baseline --steps 60 --warmup 10 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('fuse_rope=$f rep=$rep  %.4f ms  %.0f tok/s  launches %d' % (d['ms_per_step'], d['value'], d['launches_per_step']))"
done; done
for f in 0 1; do B200_FUSE_ROPE=$f timeout -s KILL 200 python bench.py --no-cpu-baseline --batch 8 --ctx 4096 --steps 40 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys, json; d=json.loads(sys.stdin.read()); print('B8 ctx4096 fuse_rope=$f  %.4f ms  %.0f tok/s' % (d['ms_per_step'], d['value']))"; done
