#!/bin/bash
# Round-2 evidence captures on ONE B200 (run under gpurun): DRAM traffic of the shipped whole-tree kernels at every shard
# size the bench quotes (ncu, three metrics), ncu --set full summaries of the 4-state and the 20-state kernel, the launch
# list of the bench command, and the default bench run of both arms.  Outputs under gpurun_out/ (kept well under 64 MiB:
# full reports are summarised to text on the box and only the 1 M-pattern one is brought back).
# Afterwards, here:  python tools/make_traffic_json.py gpurun_out > profiles/r2_traffic.json
mkdir -p gpurun_out
export P4B_BENCH_CACHE=/tmp/p4b_cache
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
for n in 1000000 500000 250000 125000; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:cl_tree_dna2 --launch-skip 3 -c 1 \
    --csv --log-file gpurun_out/traffic_r2_dna_$n.csv python bench.py --patterns $n --steps 3 --warmup 2 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "traffic dna $n rc=$?"
done
for c in 3 4; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:cl_tree_aa_kernel --launch-skip 4 -c 1 \
    --csv --log-file gpurun_out/traffic_r2_aa_cfg$c.csv python tools/bench_cfg.py --cfg $c --steps 3 > /dev/null 2>&1; echo "traffic aa cfg$c rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_dna2 --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_dna_1M \
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "prof dna rc=$?"
python tools/ncu_summary.py gpurun_out/prof_r2_dna_1M.ncu-rep > gpurun_out/r2_prof_dna2_1M.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_dna2 --launch-skip 3 -c 1 -f -o /tmp/prof_r2_dna_125k \
  python bench.py --patterns 125000 --steps 3 --warmup 2 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "prof dna 125k rc=$?"
python tools/ncu_summary.py /tmp/prof_r2_dna_125k.ncu-rep > gpurun_out/r2_prof_dna2_125k.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_aa_kernel --launch-skip 4 -c 1 -f -o /tmp/prof_r2_aa_cfg3 \
  python tools/bench_cfg.py --cfg 3 --steps 3 > /dev/null 2>&1; echo "prof aa rc=$?"
python tools/ncu_summary.py /tmp/prof_r2_aa_cfg3.ncu-rep > gpurun_out/r2_prof_aa_cfg3.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cl_tree_dmma --launch-skip 3 -c 1 -f -o /tmp/prof_r2_dmma61 \
  python tools/sweep_aa.py --cfg 61 --steps 2 > /dev/null 2>&1; echo "prof dmma61 rc=$?"
python tools/ncu_summary.py /tmp/prof_r2_dmma61.ncu-rep > gpurun_out/r2_prof_dmma61.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > /dev/null 2>&1; echo "launches rc=$?"
du -sh gpurun_out
