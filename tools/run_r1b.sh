#!/bin/bash
# round-1 session-2 GPU pass: tests, headline bench, MCMC config
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -c 3000 gpurun_out/bench_r1_b.json
for opt in "" "--no-batch" "--no-batch --no-defer" "--no-bulk"; do
  python tools/bench_mcmc.py --gens 100 --warmup 10 $opt 2>&1 | tail -1
done
python tools/bench_mcmc.py --gens 100 --warmup 10 --patterns 62500 2>&1 | tail -1
python tools/bench_mcmc.py --gens 100 --warmup 10 --patterns 62500 --no-batch 2>&1 | tail -1
python tools/bench_cfg.py --cfg 4 --steps 10 2>&1 | tail -1
