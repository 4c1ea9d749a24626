#!/bin/bash
# N-GPU pass: sharded headline bench (N from $1)
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_${N}gpu.json 2> gpurun_out/bench_r1_${N}gpu.err
tail -c 2300 gpurun_out/bench_r1_${N}gpu.json; grep -i "error\|Traceback" -A5 gpurun_out/bench_r1_${N}gpu.err | head -20
