#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused20.py -m gpu -x -q 2>&1 | tail -5
run() { timeout 120 python tools/bench_cfg.py --cfg 3 --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1  ms %.3f  GB/s %.0f  TFLOP/s %.2f  e2e %.2f lnL %.6f' % (d['ms_per_eval'], d['algorithmic_GBps'], d['TFLOPs'], d['e2e_ms_calcLogLike'], d['lnL']))"; }
for g in 2 3 4; do P4B_AA_MT=2 P4B_AA_GROUPS=$g run "mt 2 groups $g"; done
P4B_AA_MT=2 P4B_AA_GROUPS=2 P4B_AA_MINB=2 run "mt 2 groups 2 minb 2"
P4B_AA_MT=2 P4B_AA_GROUPS=1 P4B_AA_MINB=3 run "mt 2 groups 1 minb 3"
P4B_AA_MT=4 P4B_AA_GROUPS=2 run "mt 4 groups 2"
timeout 120 python tools/bench_cfg.py --cfg 4 --steps 10 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
P4B_AA_MT=2 P4B_AA_GROUPS=4 timeout 300 ncu --set full --import-source on --clock-control none -k regex:cl_tree_aa -c 1 -o gpurun_out/prof_aa_tree_v5 python tools/bench_cfg.py --cfg 3 --steps 1 > gpurun_out/ncu_aa.log 2>&1
tail -3 gpurun_out/ncu_aa.log
