#!/bin/bash
# round-1 Newton-Raphson pass 4: deck kernel over categories
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_newt.py tests/test_gpu_newt_golden.py -q -m gpu > gpurun_out/newt_tests4.log 2>&1; tail -8 gpurun_out/newt_tests4.log | cut -c1-400
timeout 300 python tools/bench_opt.py --cfg 3 --no-cpu --no-brent > gpurun_out/bench_opt4_aa.json 2> gpurun_out/bench_opt4_aa.err; tail -c 1200 gpurun_out/bench_opt4_aa.json; tail -3 gpurun_out/bench_opt4_aa.err
