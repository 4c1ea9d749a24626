#!/bin/bash
# round-1 final single-GPU pass (second session): all GPU tests, smoke, both bench arms, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_final.log 2>&1; tail -6 gpurun_out/gpu_tests_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r1_final2.json 2> gpurun_out/bench_r1_final2.err; tail -c 3400 gpurun_out/bench_r1_final2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1_final2_ref.json; cut -c1-330 gpurun_out/bench_r1_final2_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_bench_final2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
