#!/bin/bash
# One gpurun call (one GPU): the whole GPU test tier, the bench line with every leg, the implicit-tier report.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -c 600 gpurun_out/r2u_bench.json
timeout 600 python tests/gpu_implicit_report.py gpurun_out/r2u_implicit_report.json > gpurun_out/r2u_implicit_report.log 2>&1; tail -5 gpurun_out/r2u_implicit_report.log
