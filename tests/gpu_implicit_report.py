"""python tests/gpu_implicit_report.py [out.json] -- diagnostic companion of tests/test_zz_gpu_implicit.py (needs a GPU): where the
time of one implicit Gauss-Legendre step goes.  For a helium film at several N it times, with CUDA events on the launching
stream after a warm-up, (a) one real-state RHS, (b) one finite-difference Jacobian (two batched RHS evaluations, batch 3N),
(c) the 6N x 6N dense solve with the unblocked and with the blocked (DMMA trailing update) factorisation, and (d) whole steps of
the integrator with its own statistics (Newton iterations, RHS evaluations, Jacobians, linear solves per step), and checks every
figure it prints against the oracle where that is cheap (Jacobian at the smallest N, LU against LAPACK).  Writes JSON."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def film(N, depth, amp):
    a = 2 * np.pi * np.arange(N) / N
    return np.concatenate([a - 0.3 * amp * depth * np.sin(a), amp * depth * np.cos(a), 0.2 * amp * depth * np.sin(a)])


def dense_mode_rhs_ms(N=4096, reps=3):
    """BASELINE config 3 ("dense FP64 BIE solve", N = 4096): one RHS with M materialised and factorised (RB_SOLVE_DENSE_LU), through
    whichever LU the environment selects (RB_LU_BLOCKED), checked against the matrix-free RHS."""
    import torch
    from superfluid_dynamics_b200 import api
    dev = torch.device("cuda:0")
    a = 2 * np.pi * np.arange(N) / N
    y0 = np.concatenate([(a - 0.4 * np.sin(a)) + 1j * (0.4 * np.cos(a)), (0.4 * np.sin(a)).astype(np.complex128)])
    props = api.ProblemProperties(rho=0.0)
    st = torch.as_tensor(y0, device=dev)
    outs = {}
    res = {}
    for mode in ("dense_lu", "matrix_free"):
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, solve_mode=mode)
        out = torch.zeros(2 * N, dtype=torch.complex128, device=dev)
        calc.run(st, out)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            calc.run(st, out)
        torch.cuda.synchronize(dev)
        res[mode + "_rhs_ms"] = (time.perf_counter() - t0) * 1e3 / reps
        outs[mode] = out.cpu().numpy()
    res["dense_vs_matrix_free_rel"] = float(np.abs(outs["dense_lu"] - outs["matrix_free"]).max() / np.abs(outs["matrix_free"]).max())
    res["blocked_lu"] = os.environ.get("RB_LU_BLOCKED", "0")
    return res


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--dense-mode":
        print(json.dumps(dense_mode_rhs_ms()))
        return
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "implicit_report.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    import torch
    from oracle import roberts_oracle as ro
    from superfluid_dynamics_b200 import api, build
    build.build(verbose=False)
    dev = torch.device("cuda:0")
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    def T(a):
        return torch.as_tensor(np.ascontiguousarray(a), device=dev)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / reps

    report = {"lu": {}, "sizes": {}}
    rng = np.random.default_rng(0)
    for n in (384, 1536, 4096, 6144):
        A = rng.standard_normal((n, n)) + 4.0 * np.eye(n)
        b = rng.standard_normal(n)
        ref = np.linalg.solve(A, b)
        Af = T(np.asfortranarray(A).ravel(order="F"))
        entry = {}
        for name, blocked in (("unblocked", 0), ("blocked_dmma", 1)):
            if blocked == 0 and n > 4096:
                continue
            errs = []

            def once():
                dA, db = Af.clone(), T(b).clone()
                info = api.lu_solve(dA, db, n, blocked)
                errs.append((info, float(np.abs(db.cpu().numpy() - ref).max() / np.abs(ref).max())))
            t0 = time.perf_counter()
            once()
            once()
            entry[name] = {"ms_incl_copies": (time.perf_counter() - t0) * 500.0, "info": errs[-1][0], "rel_err_vs_lapack": errs[-1][1],
                           "gflops": (2.0 / 3.0) * n ** 3 / ((time.perf_counter() - t0) * 0.5) / 1e9}
        # the routine the reference's MatrixSolver calls (cusolverDnDgetrf + cusolverDnDgetrs, L/MatrixSolver.cuh:123-125), here
        # through torch.linalg (cuSOLVER backend): library baseline on the same GPU, CUDA events, matrix copy excluded from neither
        try:
            torch.backends.cuda.preferred_linalg_library("cusolver")
            At, bt = T(A), T(b).reshape(n, 1)
            def cus():
                LU, piv = torch.linalg.lu_factor(At)
                return torch.linalg.lu_solve(LU, piv, bt)
            x = cus()
            ms = timed(cus, 3)
            entry["cusolver_getrf_getrs"] = {"ms": ms, "gflops": (2.0 / 3.0) * n ** 3 / (ms * 1e-3) / 1e9,
                                             "rel_err_vs_lapack": float(np.abs(x.cpu().numpy().ravel() - ref).max() / np.abs(ref).max())}
        except Exception as e:  # noqa: BLE001
            entry["cusolver_getrf_getrs"] = {"error": repr(e)[:200]}
        # the same two factorisations timed with CUDA events (device-to-device copy of A inside: 8 n^2 bytes, ~0.02-0.1 ms)
        for name, blocked in (("unblocked", 0), ("blocked_dmma", 1)):
            if name not in entry:
                continue
            dA, db = Af.clone(), T(b).clone()
            def run():
                dA.copy_(Af)
                db.copy_(T(b))
                api.lu_solve(dA, db, n, blocked)
            ms = timed(run, 2)
            entry[name]["ms_events"] = ms
            entry[name]["gflops_events"] = (2.0 / 3.0) * n ** 3 / (ms * 1e-3) / 1e9
        report["lu"][str(n)] = entry
        print("lu", n, entry, flush=True)
        json.dump(report, open(out_path, "w"), indent=1)

    # the dense solve mode of the hot path at N = 4096 with either factorisation (the choice is read once per process)
    import subprocess
    report["dense_mode_N4096"] = {}
    for blocked in ("0", "1"):
        env = dict(os.environ, RB_LU_BLOCKED=blocked)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--dense-mode"], env=env, capture_output=True, text=True, timeout=240)
            report["dense_mode_N4096"]["blocked" if blocked == "1" else "unblocked"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            report["dense_mode_N4096"]["blocked" if blocked == "1" else "unblocked"] = {"error": repr(e)[:300]}
    print("dense mode", report["dense_mode_N4096"], flush=True)
    json.dump(report, open(out_path, "w"), indent=1)

    depth = 0.0942478   # the App's film (A/kernel.cu:79)
    for N in (64, 256, 512):
        props = api.ProblemProperties(rho=1.0, depth=depth)
        prob = api.HeliumBoundaryProblem(props)
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, prob, device=dev, guess="warm")
        real = api.RealBoundaryItegralCalculator(calc)
        jc = api.JacobianCalculator(N, props, prob, device=dev)
        y = film(N, depth, 0.1)
        dy = T(y)
        out = torch.zeros(3 * N, dtype=torch.float64, device=dev)
        J = torch.zeros(9 * N * N, dtype=torch.float64, device=dev)
        e = {"rhs_ms": timed(lambda: real.run(dy, out), 20), "jacobian_ms": timed(lambda: jc.calculateJacobian(dy, J), 3),
             "jacobian_solver": jc.solve_stats()}
        if N == 64:
            exp = ro.jacobian_fd(y, N, ro.ProblemProperties(rho=1.0, depth=depth), "helium", 1e-6)
            e["jacobian_rel_err_vs_oracle"] = float(np.abs(J.cpu().numpy().reshape(3 * N, 3 * N).T - exp).max() / np.abs(exp).max())
        h = 0.05
        gl = api.GaussLegendre2(real, jc, api.GaussLegendre2Options(stepSize=h, returnTrajectory=False))
        gl.initialize(y, False)
        t0 = time.perf_counter()
        steps = 0
        for _ in range(5):
            if not gl.step(h):
                break
            steps += 1
        el = time.perf_counter() - t0
        st = gl.stats()
        e["gl2"] = {"h": h, "steps": steps, "s_per_step": el / max(steps, 1), "stats": st,
                    "per_step": {k: st[k] / max(steps, 1) for k in ("newton_iterations", "rhs_evaluations", "jacobians", "linear_solves")}}
        report["sizes"][str(N)] = e
        print("N", N, e, flush=True)
        json.dump(report, open(out_path, "w"), indent=1)
        del gl, jc, real, calc
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
