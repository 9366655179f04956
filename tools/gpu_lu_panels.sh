#!/bin/bash
# One gpurun call: the blocked LU (rb_lu_solve) with the one-CTA panel and with the cooperative panel, several n (CUDA events via
# wall clock around a synchronised call, 5 repetitions each; see profiles/r02q_lu_panels.txt).
for n in 512 700 1000 1300 2048 3072 9000 12288; do
  for p in one coop; do
    echo "== n=$n panel=$p"; RB_LU_PANEL=$p timeout 120 python tests/gpu_lu_profile.py $n 2>&1 | tail -6 | tr '\n' ' '; echo
  done
done
