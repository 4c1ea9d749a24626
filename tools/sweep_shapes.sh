#!/bin/bash
# tools/sweep_shapes.sh -- whole-tree kernel launch shapes against shard size (run under gpurun)
export P4B_BENCH_CACHE=/tmp/p4bcache
for pat in ${PATS:-1000000 500000 250000 125000}; do
  for v in ${VARIANTS:--1 0 1 2 3 5 6}; do
    if [ $v -ge 0 ]; then export P4B_FUSED_VARIANT=$v; else unset P4B_FUSED_VARIANT; fi
    python bench.py --patterns $pat --steps 40 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('patterns %8d variant %2s  ms %.4f  evals/s %8.1f  e2e %8.1f' % ($pat, '$v', d['ms_per_step'], d['value'], d['e2e']['value']))"
  done
done
