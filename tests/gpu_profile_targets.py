"""python tests/gpu_profile_targets.py ensemble|helium|n4096 -- short runs of one configuration for ncu captures (tools/gpu_profile_round2b.sh)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import roberts_oracle as ro  # noqa: E402
from superfluid_dynamics_b200 import api  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
what = sys.argv[1] if len(sys.argv) > 1 else "ensemble"
if what == "ensemble":
    N, B, dt, steps = 512, 1024, 1e-3, 24
    hs = 0.05 + 0.35 * np.arange(B) / (B - 1)
    y0 = api.ensemble_state([ro.pack_state(*ro.trochoid(N, h)) for h in hs], N)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, B, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
elif what == "helium":
    N, B, dt, steps, depth = 16384, 1, 1e-3, 24, 0.0942478
    al = 2 * np.pi * np.arange(N) / N
    y0 = ro.pack_state(al + 1j * 0.1 * depth * np.cos(al), np.zeros(N))
    props = api.ProblemProperties(rho=0.0, depth=depth)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), device=dev, guess="warm")
else:
    N, B, dt, steps = 4096, 1, 1e-3, 40
    y0 = ro.pack_state(*ro.trochoid(N, 0.4))
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
stp = api.AutonomousRungeKuttaStepper(calc, dt)
y = torch.as_tensor(y0, device=dev)
stp.initialize(y, True)
stp.runSteps(steps)
torch.cuda.synchronize()
print(what, stp.stats(), calc.solve_stats(), calc.sweepPlan(), flush=True)
