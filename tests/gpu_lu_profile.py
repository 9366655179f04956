"""python tests/gpu_lu_profile.py [n] -- two blocked LU solves of a Gaussian n x n matrix (first: warm-up), for ncu launch lists:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lu_launches.csv python tests/gpu_lu_profile.py 2048
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from superfluid_dynamics_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
A = rng.standard_normal((n, n)) + 4.0 * np.eye(n)
b = rng.standard_normal(n)
Af = torch.as_tensor(np.asfortranarray(A).ravel(order="F"), device=dev)
bf = torch.as_tensor(b, device=dev)
for rep in range(int(os.environ.get("REPS", "5"))):
    dA, db = Af.clone(), bf.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = api.lu_solve(dA, db, n, 1)
    torch.cuda.synchronize()
    print(f"rep {rep}: n={n} info={info} {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
x = db.cpu().numpy()
print("rel err vs LAPACK", float(np.abs(x - np.linalg.solve(A, b)).max() / np.abs(x).max()))
