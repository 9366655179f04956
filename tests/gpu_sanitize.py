"""python tests/gpu_sanitize.py -- a few small calls through every kernel added in round 2, meant to be run under compute-sanitizer:
    compute-sanitizer --tool racecheck python tests/gpu_sanitize.py
    compute-sanitizer --tool memcheck  python tests/gpu_sanitize.py
(shared-memory hazards of the fused transforms' ping-pong exchange, the warp-per-row-group sweeps' staging / shared row groups /
double-buffered members, the cooperative LU panel; out-of-bounds accesses of the padded arrays)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import roberts_oracle as ro  # noqa: E402
from superfluid_dynamics_b200 import api  # noqa: E402

dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(a, device=dev)  # noqa: E731
props = api.ProblemProperties(rho=0.0)
sizes = [int(v) for v in os.environ.get("SAN_SIZES", "300,512,1024,2048,4096").split(",")]
for N in sizes:
    c = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    y = T(ro.pack_state(*ro.trochoid(N, 0.3)))
    out = torch.zeros(2 * N, dtype=torch.complex128, device=dev)
    c.run(y, out)
    torch.cuda.synchronize()
    print("rhs", N, c.sweepPlan()["kernel"], c.solve_stats()["converged"], flush=True)
    stp = api.AutonomousRungeKuttaStepper(c, 1e-3)
    stp.initialize(y, True)
    stp.runSteps(2)
    torch.cuda.synchronize()
    print("steps", N, stp.stats()["graph_launches"], flush=True)
os.environ["RB_SWEEP_V3B"] = "1"
for N, B in ((300, 5), (512, 3)):
    members = [ro.pack_state(*ro.trochoid(N, 0.1 + 0.05 * b)) for b in range(B)]
    c = api.BaseBoundaryIntegralCalculator(N, B, props, api.WaterBoundaryProblem(props))
    y = T(api.ensemble_state(members, N))
    out = torch.zeros(2 * N * B, dtype=torch.complex128, device=dev)
    c.run(y, out)
    torch.cuda.synchronize()
    print("ensemble", N, B, c.sweepPlan()["kernel"], c.solve_stats()["converged"], flush=True)
os.environ.pop("RB_SWEEP_V3B")
for n in (300, 700):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    dA = T(np.asfortranarray(A).ravel(order="F")).clone()
    db = T(b).clone()
    info = api.lu_solve(dA, db, n, 1)
    torch.cuda.synchronize()
    print("lu", n, info, float(np.abs(db.cpu().numpy() - np.linalg.solve(A, b)).max()), flush=True)
print("done", flush=True)
