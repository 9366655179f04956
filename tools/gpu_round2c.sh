#!/bin/bash
# One gpurun call (one GPU): the bench line with every leg, then an ncu launch list of recorded RK4 steps at N = 4096 (the last
# 160 kernel launches of 24 steps = about three steady-state steps) for profiles/r02t_launches_N4096_summary.txt.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 3000 gpurun_out/r2t_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches_N4096.csv python tests/gpu_round2.py steps_small > gpurun_out/r2t_steps_small.log 2>&1
tail -3 gpurun_out/r2t_steps_small.log
wc -l gpurun_out/r2t_launches_N4096.csv
