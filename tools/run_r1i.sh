#!/bin/bash
# round-1 Newton-Raphson pass: parity tests, bench_opt, launch list + one full capture of the derivative kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newt.py tests/test_gpu_opt.py tests/test_gpu_sim.py -q -m gpu > gpurun_out/newt_tests.log 2>&1; tail -45 gpurun_out/newt_tests.log | cut -c1-400
timeout 900 python tools/bench_opt.py > gpurun_out/bench_opt.json 2> gpurun_out/bench_opt.err; tail -c 2500 gpurun_out/bench_opt.json; tail -5 gpurun_out/bench_opt.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"newt_|cl_dna_kernel|transpose_deck" --launch-skip 3000 -c 500 --csv --log-file gpurun_out/launches_r1_newt.csv python tools/bench_opt.py --no-brent --no-cpu > gpurun_out/bench_opt_ncu.log 2>&1; tail -2 gpurun_out/bench_opt_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none -k regex:newt_kernel --launch-skip 200 -c 2 -o gpurun_out/prof_newt_v1 python tools/bench_opt.py --no-brent --no-cpu --patterns 500000 > gpurun_out/bench_opt_ncu2.log 2>&1; tail -2 gpurun_out/bench_opt_ncu2.log | cut -c1-300
