"""Diagnostic sweep over the CUDA path (prints errors instead of asserting).  Run under gpurun:
    python tests/gpu_debug.py > gpurun_out/debug.log 2>&1
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import roberts_oracle as ro  # noqa: E402
from superfluid_dynamics_b200 import api  # noqa: E402

dev = torch.device("cuda:0")


def T(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def section(name, fn):
    t = time.time()
    try:
        fn()
    except Exception:
        print(f"[{name}] EXCEPTION")
        traceback.print_exc()
    print(f"[{name}] done in {time.time() - t:.2f}s", flush=True)


def dense():
    for N, h, t in ((2, 0.5, 0.0), (4, 0.5, 0.1), (8, 0.5, 0.1), (64, 0.3, 0.0)):
        Z, Phi = ro.trochoid(N, h, 10.0, t)
        Zp, Zpp, _ = ro.trochoid_derivatives(N, h, 10.0, t)
        A = torch.empty(N * N, dtype=torch.float64, device=dev)
        api.createMKernel(A, T(Z), T(Zp), T(Zpp), 0.0, N)
        print("createM", N, rel(A.cpu().numpy().reshape(N, N).T, ro.create_M(Z, Zp, Zpp, 0.0)))
        api.createFiniteDepthMKernel(A, T(Z), T(Zp), T(Zpp), 0.3, N, 1, False)
        print("createFiniteDepthM", N, rel(A.cpu().numpy().reshape(N, N).T, ro.create_finite_depth_M(Z, Zp, Zpp, 0.3)))
        V1 = torch.empty(N * N, dtype=torch.complex128, device=dev)
        V2 = torch.empty(N, dtype=torch.complex128, device=dev)
        for lower in (True, False):
            api.createVelocityMatrices(T(Z), T(Zp), T(Zpp), N, V1, V2, lower)
            e1, e2 = ro.velocity_matrices(Z, Zp, Zpp, lower)
            print("V1/V2", N, lower, rel(V1.cpu().numpy().reshape(N, N).T, e1), rel(V2.cpu().numpy(), e2))
            api.createHeliumVelocityMatrices(T(Z), T(Zp), T(Zpp), 0.3, N, V1, V2, lower)
            e1, e2 = ro.helium_velocity_matrices(Z, Zp, Zpp, 0.3, lower)
            print("helium V1/V2", N, lower, rel(V1.cpu().numpy().reshape(N, N).T, e1), rel(V2.cpu().numpy(), e2))


def derivs():
    for N in (4, 8, 64, 1024, 1000):
        Z, Phi = ro.trochoid(N, 0.5, 10.0, 0.0)
        props = api.ProblemProperties(rho=0.0)
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
        Zp, PhiP, Zpp = calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
        eZp, ePhiP, eZpp = ro.zphi_derivative(Z, Phi, ro.ProblemProperties(rho=0.0), "cuda")
        print("zphi", N, np.abs(Zp.cpu().numpy() - eZp).max(), np.abs(Zpp.cpu().numpy() - eZpp).max(),
              np.abs(PhiP.cpu().numpy() - ePhiP).max())
        x = np.cos(3 * np.arange(N) * 2 * np.pi / N) + 0.1 * np.random.default_rng(0).standard_normal(N)
        d = calc.fftDerivative(T(x.astype(np.complex128)), False, 0.5)
        print("fftDerivative", N, np.abs(d.cpu().numpy() - ro.fft_derivative(x, 0.5)).max())
        d = calc.fftDerivative(T(x.astype(np.complex128)), True, 1.0)
        print("fftDerivative2", N, np.abs(d.cpu().numpy() - ro.fft_derivative(x, 1.0, second=True)).max())


def cotsum():
    for N, h in ((64, 0.3), (256, 0.3), (1000, 0.4), (1024, 0.4), (4096, 0.4), (5000, 0.2)):
        Z, Phi = ro.trochoid(N, h)
        props = api.ProblemProperties(rho=0.0)
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
        calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
        x = np.cos(3 * 2 * np.pi * np.arange(N) / N) + 0.3 * np.sin(2 * np.pi * np.arange(N) / N) + 0.1
        S = calc.cotangentSum(T(Z), T(x)).cpu().numpy()
        rows = np.arange(N) if N <= 1024 else np.r_[0:4, N - 4:N, N // 2 - 2:N // 2 + 2, 255:258, 511:514]
        e = ro.cot_rowsum(Z, x, rows)
        print("cotsum", N, h, rel(S[rows], e), "nan" if np.isnan(S).any() else "")


def full_rhs():
    for physics, N, h, B in (("water", 64, 0.1, 1), ("water", 64, 0.4, 1), ("water", 256, 0.4, 1), ("water", 1000, 0.3, 1),
                             ("water", 1024, 0.4, 1), ("water", 2048, 0.4, 1), ("water", 128, 0.3, 3),
                             ("helium_inf", 256, 0.05, 1), ("helium", 256, 0.01, 1), ("helium", 128, 0.02, 2)):
        rho = 0.0
        depth = 0.3 if physics != "water" else 1.0
        props = api.ProblemProperties(rho=rho, depth=depth)
        oprops = ro.ProblemProperties(rho=rho, depth=depth)
        prob = {"water": api.WaterBoundaryProblem, "helium": api.HeliumBoundaryProblem,
                "helium_inf": api.HeliumInfiniteDepthBoundaryProblem}[physics](props)
        states = []
        for b in range(B):
            Z, Phi = ro.trochoid(N, h * (1 + 0.3 * b))
            states.append((Z, Phi))
        st = np.concatenate([s[0] for s in states] + [s[1].astype(np.complex128) for s in states])
        for mode in ("matrix_free", "dense_lu"):
            calc = api.BaseBoundaryIntegralCalculator(N, B, props, prob, solve_mode=mode)
            out = torch.zeros(2 * N * B, dtype=torch.complex128, device=dev)
            calc.run(T(st), out)
            torch.cuda.synchronize()
            e = ro.rhs(st, N, B, oprops, physics, "cuda")
            o = out.cpu().numpy()
            stats = calc.solve_stats()
            print("rhs", physics, N, h, B, mode, "vel", rel(o[:N * B], e[:N * B]), "dphi", rel(o[N * B:], e[N * B:]), stats)
            _, _, aux = ro.rhs_single(states[0][0], states[0][1], oprops, physics, "cuda", full=True)
            print("    a", rel(calc.getDevA().cpu().numpy()[:N], aux["a"]), "upper",
                  rel(calc.devVelocitiesUpper.cpu().numpy()[:N], aux["v_upper"]))


def rk4():
    for N, h, steps in ((64, 0.1, 100), (64, 0.4, 100), (256, 0.3, 100), (1024, 0.4, 20)):
        props = api.ProblemProperties(rho=0.0)
        oprops = ro.ProblemProperties(rho=0.0)
        Z, Phi = ro.trochoid(N, h)
        y0 = ro.pack_state(Z, Phi)
        f = lambda s: ro.rhs(s, N, 1, oprops, "water", "cuda")
        ye = y0.copy()
        for _ in range(steps):
            ye = ro.rk4_step(f, ye, 1e-3)
        for guess in ("cold", "warm"):
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess=guess)
            stp = api.AutonomousRungeKuttaStepper(calc, 1e-3)
            stp.initialize(y0, False)
            t0 = time.time()
            stp.runSteps(steps)
            y = stp.getState()
            dt = time.time() - t0
            print("rk4", N, h, steps, guess, "Z", rel(y[:N], ye[:N]), "Phi", rel(y[N:], ye[N:]), f"{steps / dt:.1f} steps/s",
                  calc.solve_stats(), stp.stats())


def steprate():
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for N, h, dt, steps in ((256, 0.3, 1e-3, 200), (1024, 0.4, 1e-3, 200), (4096, 0.4, 1e-3, 200), (16384, 0.4, 2e-4, 40),
                            (65536, 0.4, 1e-4, 20)):
        for order in (1, 2, 3, 4):
            os.environ["RB_GUESS_ORDER"] = str(order)
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
            Z, Phi = ro.trochoid(N, h)
            st = T(ro.pack_state(Z, Phi))
            stp.initialize(st, True)
            stp.runSteps(10)
            torch.cuda.synchronize()
            s0 = calc.solve_stats()
            t0 = time.time()
            stp.runSteps(steps)
            torch.cuda.synchronize()
            el = time.time() - t0
            s1 = calc.solve_stats()
            its = (s1["total_iterations"] - s0["total_iterations"]) / max(1, s1["total_solves"] - s0["total_solves"])
            print(f"steprate N={N} order={order}: {steps / el:.1f} steps/s, {its:.2f} sweeps/solve, {stp.stats()}", flush=True)


def tune():
    """sweep time vs the CTA-count target of the chunking heuristic"""
    for N in (4096, 16384, 65536):
        for target in (148 * 4, 148 * 8, 148 * 16, 148 * 32, 148 * 64):
            os.environ["RB_TARGET_CTAS"] = str(target)
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
            Z, Phi = ro.trochoid(N, 0.4)
            ms, pairs = calc.benchSweep(T(ro.pack_state(Z, Phi)), 20)
            print(f"tune N={N} target={target}: {ms * 1e3:.1f} us  {20 * pairs / (ms * 1e-3) / 1e12:.2f} TF", flush=True)
    os.environ.pop("RB_TARGET_CTAS", None)


def v2cmp():
    """tiled (v1) vs persistent (v2) sweep"""
    for N in (256, 1024, 4096, 8192, 16384, 32768, 65536):
        for v2 in ("0", "1"):
            os.environ["RB_SWEEP_V2"] = v2
            os.environ["RB_VERBOSE"] = "1" if v2 == "1" else "0"
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
            Z, Phi = ro.trochoid(N, 0.4)
            ms, pairs = calc.benchSweep(T(ro.pack_state(Z, Phi)), 50 if N <= 16384 else 10)
            print(f"v2cmp N={N} v2={v2}: {ms * 1e3:.1f} us  {20 * pairs / (ms * 1e-3) / 1e12:.2f} TF", flush=True)
    os.environ.pop("RB_SWEEP_V2", None); os.environ.pop("RB_VERBOSE", None)


def tune2():
    """persistent sweep at small N: source groups per CTA (threads = 32 * groups at R = 1), rows per thread, tiled kernel beside it"""
    for N in (1024, 2048, 4096, 8192):
        Z, Phi = ro.trochoid(N, 0.4)
        st = T(ro.pack_state(Z, Phi))
        cfgs = [dict(RB_SWEEP_V2="0")]
        for G in (4, 8, 16, 32):
            cfgs.append(dict(RB_SWEEP_V2="1", RB_V2_GROUPS=str(G), RB_V2_RB="32", RB_V2_R="1"))
        for G in (4, 8, 16):
            cfgs.append(dict(RB_SWEEP_V2="1", RB_V2_GROUPS=str(G), RB_V2_RB="64", RB_V2_R="2"))
        if N >= 4096:
            for G in (4, 8):
                cfgs.append(dict(RB_SWEEP_V2="1", RB_V2_GROUPS=str(G), RB_V2_RB="64", RB_V2_R="2", RB_V2_SPLIT="2"))
        for cfg in cfgs:
            os.environ.update(cfg)
            try:
                props = api.ProblemProperties(rho=0.0)
                calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
                ms, pairs = calc.benchSweep(st, 50)
                print(f"tune2 N={N} {cfg}: {ms * 1e3:.1f} us  {20 * pairs / (ms * 1e-3) / 1e12:.2f} TF", flush=True)
            except Exception as e:
                print(f"tune2 N={N} {cfg}: FAILED {e}", flush=True)
            for k in cfg:
                os.environ.pop(k, None)


def overlapcmp():
    """a' transform beside the combined sweep (graph fork/join) + fused geometry/guess and finish/RK update: step rate and state"""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for N, dt in ((256, 1e-3), (1024, 1e-3), (2048, 1e-3), (4096, 1e-3), (8192, 5e-4)):
        ref = None
        for ov in ("0", "1"):
            os.environ["RB_OVERLAP"] = ov
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
            Z, Phi = ro.trochoid(N, 0.4)
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
            st = T(ro.pack_state(Z, Phi))
            stp.initialize(st, True)
            stp.runSteps(20)
            torch.cuda.synchronize()
            t0 = time.time()
            stp.runSteps(300)
            torch.cuda.synchronize()
            el = time.time() - t0
            y = st.cpu().numpy()
            if ref is None:
                ref = y
            print(f"overlapcmp N={N} overlap={ov}: {300 / el:.1f} steps/s dev={rel(y, ref):.2e} {stp.stats()} {calc.solve_stats()}", flush=True)
    os.environ.pop("RB_OVERLAP", None)


def tune3():
    """persistent sweep at large N: 4 source groups (896 threads, 72 registers) against 3 (672 threads, 80 registers)"""
    for N in (65536, 32768):
        Z, Phi = ro.trochoid(N, 0.4)
        st = T(ro.pack_state(Z, Phi))
        for cfg in (dict(RB_SWEEP_V2="1"), dict(RB_SWEEP_V2="1", RB_V2_GROUPS="3"), dict(RB_SWEEP_V2="1", RB_V2_GROUPS="2"),
                    dict(RB_SWEEP_V2="0")):
            os.environ.update(cfg)
            os.environ["RB_VERBOSE"] = "1"
            try:
                props = api.ProblemProperties(rho=0.0)
                calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
                ms, pairs = calc.benchSweep(st, 10)
                print(f"tune3 N={N} {cfg}: {ms * 1e3:.1f} us  {20 * pairs / (ms * 1e-3) / 1e12:.2f} TF", flush=True)
            except Exception as e:
                print(f"tune3 N={N} {cfg}: FAILED {e}", flush=True)
            for k in list(cfg) + ["RB_VERBOSE"]:
                os.environ.pop(k, None)


def tune4():
    """tiled sweep at N = 65536: rows per thread (CTA = 256 / rows threads) against the CTA-count target"""
    N = 65536
    Z, Phi = ro.trochoid(N, 0.4)
    st = T(ro.pack_state(Z, Phi))
    for rows, target in ((4, 4736), (4, 9472), (4, 18944), (8, 4736), (8, 9472), (8, 18944), (8, 37888)):
        cfg = dict(RB_SWEEP_V2="0", RB_V1_ROWS=str(rows), RB_TARGET_CTAS=str(target))
        os.environ.update(cfg)
        try:
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
            ms, pairs = calc.benchSweep(st, 8)
            print(f"tune4 N={N} rows={rows} target={target}: {ms * 1e3:.1f} us  {20 * pairs / (ms * 1e-3) / 1e12:.2f} TF", flush=True)
        except Exception as e:
            print(f"tune4 N={N} {cfg}: FAILED {e}", flush=True)
        for k in cfg:
            os.environ.pop(k, None)


def kernelcmp():
    """step rate with the persistent (v2) against the tiled (v1) sweep kernel in the launch-bound regime"""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for N, dt in ((1024, 1e-3), (2048, 1e-3), (4096, 1e-3), (8192, 5e-4)):
        for cfg in (dict(RB_SWEEP_V2="1"), dict(RB_SWEEP_V2="0"), dict(RB_SWEEP_V2="0", RB_TARGET_CTAS="592"),
                    dict(RB_SWEEP_V2="0", RB_TARGET_CTAS="2368")):
            os.environ.update(cfg)
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
            Z, Phi = ro.trochoid(N, 0.4)
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
            st = T(ro.pack_state(Z, Phi))
            stp.initialize(st, True)
            stp.runSteps(20)
            torch.cuda.synchronize()
            t0 = time.time()
            stp.runSteps(300)
            torch.cuda.synchronize()
            el = time.time() - t0
            print(f"kernelcmp N={N} {cfg}: {300 / el:.1f} steps/s {stp.stats()}", flush=True)
            for k in cfg:
                os.environ.pop(k, None)


def ensemble():
    """BASELINE config 5, second half: 1024-member ensemble at N = 512 (replicas only across GPUs)"""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    N, B, dt, steps = 512, 1024, 1e-3, 30
    hs = 0.05 + 0.35 * np.arange(B) / (B - 1)
    Zs, Ps = zip(*(ro.trochoid(N, h) for h in hs))
    y0 = np.concatenate(list(Zs) + [p.astype(np.complex128) for p in Ps])
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, B, props, api.WaterBoundaryProblem(props), guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    st = T(y0)
    stp.initialize(st, True)
    stp.runSteps(12)
    torch.cuda.synchronize()
    s0 = calc.solve_stats()
    t0 = time.time()
    stp.runSteps(steps)
    torch.cuda.synchronize()
    el = time.time() - t0
    s1 = calc.solve_stats()
    its = (s1["total_iterations"] - s0["total_iterations"]) / max(1, s1["total_solves"] - s0["total_solves"])
    print(f"ensemble B={B} N={N}: {steps / el:.1f} steps/s = {B * steps / el:.0f} member-steps/s, {its:.2f} sweeps/solve, "
          f"{stp.stats()} finite={bool(torch.isfinite(torch.view_as_real(st)).all())}", flush=True)
    ms, pairs = calc.benchSweep(T(y0), 20)
    print(f"ensemble sweep: {ms * 1e3:.1f} us, {20 * pairs / (ms * 1e-3) / 1e12:.2f} TFLOP/s")


def optimistic():
    """first sweep already combined (verify + velocities) when the guess (extrapolation + predicted row sums) is good enough:
    sweeps per solve and step rate against the solve tolerance"""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for N, h, dt, steps in ((1024, 0.4, 1e-3, 200), (4096, 0.4, 1e-3, 200), (16384, 0.4, 2e-4, 60), (65536, 0.4, 1e-4, 30)):
        ref = None
        for tol, order, policy, predict in ((1e-13, 4, 0, 0), (1e-13, 4, 1, -1), (1e-12, 4, 1, -1), (1e-12, 5, 1, -1), (1e-11, 4, 1, -1),
                                            (1e-10, 4, 1, -1), (1e-10, 3, 1, -1)):
            props = api.ProblemProperties(rho=0.0)
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm", tolerance=tol)
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
            stp.setGuess(order, predict)
            stp.setOptimistic(policy)
            Z, Phi = ro.trochoid(N, h)
            st = T(ro.pack_state(Z, Phi))
            stp.initialize(st, True)
            stp.runSteps(12)
            torch.cuda.synchronize()
            s0 = calc.solve_stats()
            t0 = time.time()
            stp.runSteps(steps)
            torch.cuda.synchronize()
            el = time.time() - t0
            s1 = calc.solve_stats()
            its = (s1["total_iterations"] - s0["total_iterations"]) / max(1, s1["total_solves"] - s0["total_solves"])
            y = st.cpu().numpy()
            if ref is None:
                ref = y
            gs = stp.guessStats()
            print(f"optimistic N={N} dt={dt} tol={tol:g} order={order} policy={policy} predict={predict}: {steps / el:.1f} steps/s, "
                  f"{its:.2f} sweeps/solve, first_rel={['%.1e' % v for v in gs['first_rel']]} mask={gs['opt_mask']} "
                  f"one_sweep={gs['one_sweep_solves']}/{gs['optimistic_solves']} dev_vs_first={rel(y, ref):.2e} "
                  f"final_rel={s1['residual']:.1e} {stp.stats()}", flush=True)


def dmma():
    print("fp64 peak TFLOP/s", api.measure_fp64_peak())
    for k, (ms, tf) in api.measure_fp64_tensor_overlap().items():
        print(f"dmma-overlap {k}: {ms:.3f} ms, {tf:.2f} TFLOP/s", flush=True)


def speed():
    print("fp64 peak TFLOP/s", api.measure_fp64_peak(), " 3-register-operand DFMA:", api.measure_fp64_rate_3operand())
    for N in (1024, 4096, 16384, 65536):
        props = api.ProblemProperties(rho=0.0)
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
        Z, Phi = ro.trochoid(N, 0.4)
        st = T(ro.pack_state(Z, Phi))
        ms, pairs = calc.benchSweep(st, 10)
        print(f"sweep N={N}: {ms * 1e3:.1f} us, {pairs / ms / 1e9:.3f} Gpairs/ms = {20 * pairs / (ms * 1e-3) / 1e12:.2f} TFLOP/s (20 flop/pair)")
        out = torch.zeros_like(st)
        calc.run(st, out)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            calc.run(st, out)
        torch.cuda.synchronize()
        print(f"   rhs {1e3 * (time.time() - t0) / 5:.3f} ms", calc.solve_stats())


if __name__ == "__main__":
    which = sys.argv[1:] or ["dense", "derivs", "cotsum", "full_rhs", "rk4", "speed"]
    print(torch.cuda.get_device_name(0))
    for w in which:
        section(w, globals()[w])
