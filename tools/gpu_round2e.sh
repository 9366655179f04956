#!/bin/bash
# Final check of the round on 2 GPUs: smoke(), the bench line with every leg at 2 ranks (row shards and ensemble halves on the new kernels).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err
tail -c 1500 gpurun_out/r2v_bench_2gpu.json; tail -3 gpurun_out/r2v_bench_2gpu.err
