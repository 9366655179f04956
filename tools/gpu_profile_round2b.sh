#!/bin/bash
# Second round-2 evidence run (one GPU): full launch list of the bench command, captures of the ensemble and helium sweeps.
set -u
mkdir -p gpurun_out
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2q_launches_N65536.csv python bench.py --steps 6 --warmup 5 --no-cpu --no-extra > $O/r2q_bench_under_ncu_N65536.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2q_launches_helium.csv python tests/gpu_profile_targets.py helium > $O/r2q_helium_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2q_launches_ensemble.csv python tests/gpu_profile_targets.py ensemble > $O/r2q_ensemble_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -s 150 -c 2 -o $O/r2q_sweep2_ensemble python tests/gpu_profile_targets.py ensemble > $O/r2q_ncu_ensemble.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 300 -c 2 -o $O/r2q_sweep_helium python tests/gpu_profile_targets.py helium > $O/r2q_ncu_helium.log 2>&1
tail -n 3 $O/r2q_ncu_ensemble.log $O/r2q_ncu_helium.log
ls -la $O | grep r2q
