#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for opt in "" "--no-batch"; do
  python tools/bench_mcmc.py --gens 100 --warmup 10 $opt 2>&1 | tail -1
done
python tools/bench_mcmc.py --gens 100 --warmup 10 --patterns 62500 2>&1 | tail -1
python tools/bench_mcmc.py --gens 100 --warmup 10 --patterns 62500 --no-batch 2>&1 | tail -1
python tools/bench_mcmc.py --engine reference --gens 20 --warmup 2 --ref-patterns 2048 2>&1 | tail -1
python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -c 1500
