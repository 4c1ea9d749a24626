#!/bin/bash
# Round-2 opener (one B200): where the per-step floor of the whole-tree kernels comes from, at the shard sizes of 4 and 8 GPUs.
#   - bench lines at 1 M / 250 k / 125 k patterns (the kernel's time per evaluation)
#   - ncu --set full of cl_tree_dna_kernel at 125 k patterns (stall reasons of a latency-bound launch)
#   - launch list + full capture of the 20-state Newton kernel (FMA pipe 38 %)
mkdir -p gpurun_out
for n in 1000000 250000 125000; do
  timeout 300 python bench.py --patterns $n --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-400
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_tree_dna --launch-skip 3 -c 1 -o gpurun_out/prof_r2_tree_125k \
  python bench.py --patterns 125000 --steps 3 --warmup 2 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:newt_aa_kernel --launch-skip 100 -c 2 -o gpurun_out/prof_r2_newt_aa \
  python tools/bench_opt.py --cfg 3 --no-brent --no-cpu > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
