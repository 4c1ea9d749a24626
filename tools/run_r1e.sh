#!/bin/bash
# 2-GPU pass: sharded headline bench, sharded MCMC parity of lnL
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_2gpu.json 2> gpurun_out/bench_r1_2gpu.err
tail -c 2500 gpurun_out/bench_r1_2gpu.json; tail -3 gpurun_out/bench_r1_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>&1 | tail -2
