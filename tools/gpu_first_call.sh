#!/bin/bash
# One gpurun call that settles everything written after round 1's GPU minutes were spent.  Usage (from the repo root):
#   gpurun --timeout 900 -- 'bash tools/gpu_first_call.sh'
# Writes gpurun_out/: pending_tier.log (tests/test_zz_gpu_implicit.py, verbose, xfail markers shown as XPASS / XFAIL),
# pending_tier_blocked_lu.log (the Gauss-Legendre cases again with the blocked LU selected), implicit_report.json (where the time of
# an implicit step goes, both LU factorisations), pytest_gpu.log (the whole GPU tier), bench_1gpu.json.
set -u
mkdir -p gpurun_out
python -m superfluid_dynamics_b200.build > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_zz_gpu_implicit.py -m gpu -rA -q > gpurun_out/pending_tier.log 2>&1
tail -n 60 gpurun_out/pending_tier.log
RB_LU_BLOCKED=1 timeout 300 python -m pytest tests/test_zz_gpu_implicit.py -m gpu -rA -q -k "gl2 or integrate_simulation" > gpurun_out/pending_tier_blocked_lu.log 2>&1
tail -n 15 gpurun_out/pending_tier_blocked_lu.log
timeout 300 python tests/gpu_implicit_report.py gpurun_out/implicit_report.json > gpurun_out/implicit_report.log 2>&1
tail -n 40 gpurun_out/implicit_report.log
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
tail -n 15 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
cat gpurun_out/bench_1gpu.json
