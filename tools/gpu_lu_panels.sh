#!/bin/bash
# One gpurun call: the blocked LU (rb_lu_solve) with and without the look-ahead schedule (RB_LU_LOOKAHEAD) and with the one-CTA /
# cooperative panel (RB_LU_PANEL), several n (wall clock around a synchronised call, 5 repetitions each;
# see profiles/r02q_lu_panels.txt).
for n in 512 1000 1536 2048 3072 4096 7000 12288; do
  for la in 0 1; do
    echo "== n=$n panel=coop lookahead=$la"; RB_LU_LOOKAHEAD=$la timeout 120 python tests/gpu_lu_profile.py $n 2>&1 | tail -6 | tr '\n' ' '; echo
  done
done
timeout 900 python -m pytest tests/test_zz_gpu_implicit.py -x -q -m gpu 2>&1 | tail -5
