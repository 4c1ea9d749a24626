#!/bin/bash
# Round-2 evidence captures on ONE B200 (run under gpurun): DRAM traffic of the shipped whole-tree kernels at every shard
# size the bench quotes (ncu, three metrics), ncu --set full summaries of the 4-state and the 20-state kernel, the launch
# list of the bench command, and the default bench run of both arms.  Outputs under gpurun_out/ (kept well under 64 MiB:
# full reports are summarised to text on the box and only the 1 M-pattern one is brought back).
mkdir -p gpurun_out
export P4B_BENCH_CACHE=/tmp/p4b_cache
for n in 1000000 500000 250000 125000; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cl_tree_dna2 --launch-skip 3 -c 1 \
    --csv --log-file gpurun_out/traffic_r2_dna_$n.csv python bench.py --patterns $n --steps 3 --warmup 2 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "traffic dna $n rc=$?"
done
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cl_tree_aa2 --launch-skip 3 -c 1 \
  --csv --log-file gpurun_out/traffic_r2_aa_cfg3.csv python tools/bench_cfg.py --cfg 3 --steps 3 > /dev/null 2>&1; echo "traffic aa rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_dna2 --launch-skip 3 -c 1 -o gpurun_out/prof_r2_dna_1M \
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "prof dna rc=$?"
python tools/ncu_summary.py gpurun_out/prof_r2_dna_1M.ncu-rep > gpurun_out/r2_prof_dna2_1M.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_dna2 --launch-skip 3 -c 1 -o /tmp/prof_r2_dna_125k \
  python bench.py --patterns 125000 --steps 3 --warmup 2 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "prof dna 125k rc=$?"
python tools/ncu_summary.py /tmp/prof_r2_dna_125k.ncu-rep > gpurun_out/r2_prof_dna2_125k.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_aa2 --launch-skip 3 -c 1 -o /tmp/prof_r2_aa_cfg3 \
  python tools/bench_cfg.py --cfg 3 --steps 3 > /dev/null 2>&1; echo "prof aa rc=$?"
python tools/ncu_summary.py /tmp/prof_r2_aa_cfg3.ncu-rep > gpurun_out/r2_prof_aa2_cfg3.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "launches rc=$?"
timeout 1500 python bench.py > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r2_reference.json; cut -c1-300 gpurun_out/bench_r2_reference.json
du -sh gpurun_out
