#!/bin/bash
# Final check of the round on 8 GPUs: the bench line with every leg at 8 ranks.
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 > gpurun_out/r2w_bench_8gpu.json 2> gpurun_out/r2w_bench_8gpu.err
tail -c 600 gpurun_out/r2w_bench_8gpu.json; tail -3 gpurun_out/r2w_bench_8gpu.err
