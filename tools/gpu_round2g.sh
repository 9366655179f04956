#!/bin/bash
# ncu --set full of two kernels added at the end of round 2: the fused radix-8 transform at N = 4096 and the cooperative LU panel at n = 4096.
set -u
mkdir -p gpurun_out
STEPS_K=12 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft8_zphi -s 30 -c 1 -o gpurun_out/r2x_fft8_zphi_N4096 python tests/gpu_round2.py steps_small > gpurun_out/r2x_ncu_fft8.log 2>&1
REPS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lu_panel_coop -s 70 -c 1 -o gpurun_out/r2x_lu_panel_coop_n4096 python tests/gpu_lu_profile.py 4096 > gpurun_out/r2x_ncu_lu.log 2>&1
ls -la gpurun_out/r2x*
