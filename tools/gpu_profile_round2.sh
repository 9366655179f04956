#!/bin/bash
# One gpurun call that collects the round-2 ncu evidence (one GPU).  Writes gpurun_out/r2p_*.
set -u
mkdir -p gpurun_out
O=gpurun_out
# (1) launch lists: blocked LU, with the cluster panel and with the one-CTA panel
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2p_lu_launches_n2048.csv python tests/gpu_lu_profile.py 2048 > $O/r2p_lu_profile_n2048.log 2>&1
RB_LU_CLUSTER_PANEL=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2p_lu_launches_n2048_onecta.csv python tests/gpu_lu_profile.py 2048 > $O/r2p_lu_profile_n2048_onecta.log 2>&1
python tests/gpu_lu_profile.py 4096 > $O/r2p_lu_profile_n4096.log 2>&1
RB_LU_CLUSTER_PANEL=0 python tests/gpu_lu_profile.py 4096 > $O/r2p_lu_profile_n4096_onecta.log 2>&1
tail -n 4 $O/r2p_lu_profile_n2048.log $O/r2p_lu_profile_n2048_onecta.log $O/r2p_lu_profile_n4096.log $O/r2p_lu_profile_n4096_onecta.log
# (2) launch list of the recorded step at N = 65536 (the bench command) and at N = 4096
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 400 --csv --log-file $O/r2p_launches_N65536.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > $O/r2p_bench_under_ncu_N65536.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 500 --csv --log-file $O/r2p_launches_N4096.csv python bench.py --n 4096 --steps 40 --warmup 3 --no-cpu --no-extra > $O/r2p_bench_under_ncu_N4096.log 2>&1
# (3) --set full of the dominant kernel (tiled sweep, 4 rows per thread, two-level reduction) at N = 65536
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 40 -c 2 -o $O/r2p_sweep_N65536 python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $O/r2p_ncu_sweep.log 2>&1
# (4) --set full of the trailing-update GEMM (FP64 tensor path) and the cluster panel
ncu --set full --clock-control none --import-source on -k regex:lu_gemm_kernel -s 4 -c 2 -o $O/r2p_lu_gemm python tests/gpu_lu_profile.py 4096 > $O/r2p_ncu_lu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lu_panel_cluster -s 4 -c 2 -o $O/r2p_lu_panel python tests/gpu_lu_profile.py 4096 > $O/r2p_ncu_lu_panel.log 2>&1
ls -la $O | tail -20
