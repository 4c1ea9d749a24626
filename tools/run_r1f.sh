#!/bin/bash
# round-1 final single-GPU pass: tests, both bench arms, launch list, config benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 3200 gpurun_out/bench_r1_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1_final_ref.json; cut -c1-330 gpurun_out/bench_r1_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for c in 1 3 4; do timeout 300 python tools/bench_cfg.py --cfg $c --steps 20 2>&1 | tail -1 | cut -c1-300; done
timeout 600 python tools/bench_mcmc.py --gens 200 --warmup 10 2>&1 | tail -1 | cut -c1-200
