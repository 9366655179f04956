"""Growth of grid-scale noise in the (unfiltered) Roberts scheme vs N and steepness. Run under gpurun."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import roberts_oracle as ro
from superfluid_dynamics_b200 import api
dev = torch.device("cuda:0")

def hf(y, N):
    Y = y[:N].imag
    c = torch.fft.fft(Y)
    return float(c[N // 4: 3 * N // 4].abs().max() / N)

for N, h, dt, steps in ((4096, 0.4, 1e-3, 60), (16384, 0.4, 1e-3, 40), (65536, 0.4, 1e-3, 20), (65536, 0.1, 1e-3, 40),
                        (65536, 0.01, 1e-3, 40), (65536, 0.4, 1e-4, 40), (16384, 0.1, 1e-3, 60)):
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    Z, Phi = ro.trochoid(N, h)
    st = torch.as_tensor(ro.pack_state(Z, Phi), device=dev)
    stp.initialize(st, True)
    vals, its = [], []
    t0 = time.time()
    prev = calc.solve_stats()["total_iterations"]
    blew = None
    for i in range(steps):
        try:
            stp.runStep()
        except RuntimeError as e:      # strict mode: a solve on a surface that has blown up fails loudly
            blew = f"step {i + 1}: {str(e)[:110]}"
            break
        s = calc.solve_stats()
        its.append((s["total_iterations"] - prev) / 4.0); prev = s["total_iterations"]
        vals.append(hf(st, N))
        if not np.isfinite(vals[-1]) or vals[-1] > 1e-3:
            break
    el = time.time() - t0
    v = np.array(vals)
    g = np.log(v[-1] / v[max(0, len(v) // 2)]) / (dt * (len(v) - 1 - max(0, len(v) // 2))) if len(v) > 3 else float("nan")
    if blew:
        print(f"N={N} h={h} dt={dt}: BLEW UP at {blew}", flush=True)
    if len(vals) == 0:
        continue
    print(f"N={N} h={h} dt={dt}: steps={len(v)} hf[0]={v[0]:.2e} hf[-1]={v[-1]:.2e} growth_rate~{g:.1f}/unit time "
          f"iters/rhs first={its[0]:.1f} last={its[-1]:.1f} {len(v)/el:.2f} steps/s", flush=True)
    print("   hf:", " ".join(f"{x:.1e}" for x in v[::max(1, len(v)//12)]))
