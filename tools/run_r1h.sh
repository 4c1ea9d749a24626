#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1750 -c 60 --csv --log-file gpurun_out/launches_r1_mcmc_steady.csv python tools/bench_mcmc.py --gens 30 --warmup 2 --patterns 250000 --mode batched > gpurun_out/mcmc_ncu.log 2>&1; tail -1 gpurun_out/mcmc_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --launch-skip 40 -k regex:cl_tree_dna -c 2 -o gpurun_out/prof_cl_batched_v1 python tools/bench_mcmc.py --gens 30 --warmup 2 --patterns 250000 --mode batched > gpurun_out/mcmc_ncu2.log 2>&1; tail -1 gpurun_out/mcmc_ncu2.log | cut -c1-200
