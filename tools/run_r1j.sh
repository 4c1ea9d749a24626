#!/bin/bash
# round-1 Newton-Raphson pass 2: one-launch 4-state derivative kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_newt.py tests/test_gpu_opt.py tests/test_gpu_sim.py -q -m gpu > gpurun_out/newt_tests2.log 2>&1; tail -25 gpurun_out/newt_tests2.log | cut -c1-400
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_gpu_newt.py::test_newt_around_matches_reference" -q -m gpu -k "1-kw0" > gpurun_out/newt_racecheck.log 2>&1; tail -4 gpurun_out/newt_racecheck.log | cut -c1-300
timeout 900 python tools/bench_opt.py > gpurun_out/bench_opt2.json 2> gpurun_out/bench_opt2.err; tail -c 2500 gpurun_out/bench_opt2.json; tail -5 gpurun_out/bench_opt2.err
timeout 600 python tools/bench_opt.py --cfg 3 --cpu-sample 512 > gpurun_out/bench_opt2_aa.json 2> gpurun_out/bench_opt2_aa.err; tail -c 2500 gpurun_out/bench_opt2_aa.json; tail -5 gpurun_out/bench_opt2_aa.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"newt_|cl_dna_kernel|transpose_deck" --launch-skip 3000 -c 500 --csv --log-file gpurun_out/launches_r1_newt2.csv python tools/bench_opt.py --no-brent --no-cpu > gpurun_out/bench_opt_ncu3.log 2>&1; tail -2 gpurun_out/bench_opt_ncu3.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:newt_dna_kernel --launch-skip 200 -c 2 -o gpurun_out/prof_newt_v2 python tools/bench_opt.py --no-brent --no-cpu > gpurun_out/bench_opt_ncu4.log 2>&1; tail -2 gpurun_out/bench_opt_ncu4.log | cut -c1-300
