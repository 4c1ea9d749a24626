#!/bin/bash
# round-1 Newton-Raphson pass 3: prefetching 4-state kernel, two-pattern 20-state kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newt.py tests/test_gpu_opt.py tests/test_gpu_sim.py -q -m gpu > gpurun_out/newt_tests3.log 2>&1; tail -25 gpurun_out/newt_tests3.log | cut -c1-400
timeout 900 python tools/bench_opt.py > gpurun_out/bench_opt3.json 2> gpurun_out/bench_opt3.err; tail -c 2500 gpurun_out/bench_opt3.json; tail -5 gpurun_out/bench_opt3.err
timeout 600 python tools/bench_opt.py --cfg 3 --cpu-sample 512 > gpurun_out/bench_opt3_aa.json 2> gpurun_out/bench_opt3_aa.err; tail -c 2500 gpurun_out/bench_opt3_aa.json; tail -5 gpurun_out/bench_opt3_aa.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"newt_|cl_dna_kernel|transpose_deck" --launch-skip 3000 -c 500 --csv --log-file gpurun_out/launches_r1_newt3.csv python tools/bench_opt.py --no-brent --no-cpu > gpurun_out/bench_opt_ncu5.log 2>&1; tail -2 gpurun_out/bench_opt_ncu5.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"newt_|cl_d|transpose_deck" --launch-skip 1500 -c 400 --csv --log-file gpurun_out/launches_r1_newt3_aa.csv python tools/bench_opt.py --cfg 3 --no-brent --no-cpu > gpurun_out/bench_opt_ncu6.log 2>&1; tail -2 gpurun_out/bench_opt_ncu6.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:newt_dna_kernel --launch-skip 200 -c 2 -o gpurun_out/prof_newt_v3 python tools/bench_opt.py --no-brent --no-cpu > gpurun_out/bench_opt_ncu7.log 2>&1; tail -2 gpurun_out/bench_opt_ncu7.log | cut -c1-200
