#!/usr/bin/env python
"""bench.py -- RK4 steps/s of the Roberts boundary-integral step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 65536] [--impl native|reference]

A "step" is one classical RK4 step (4 RHS evaluations: FFT derivatives, matrix-free solve for the vortex-sheet strength,
O(N^2) velocity summation, fused stage update) of one synthetic trochoidal ("Stokes") surface, h = 0.4, dt = 1e-3.
Prints ONE JSON line.  Under torchrun (N > 1) the row blocks of every O(N^2) sweep are sharded over the ranks.
--impl reference times the CPU statement of the same path (oracle port of the reference's NumPy/LAPACK arithmetic) on a
bounded sample of the same workload on the host cores.  Both arms also report, as context, the RK4 step rate of the
reference's own CUDA path (its classes compiled unmodified into oracle/_ref/, run in a child process on the same GPU) at the
sizes it supports (N <= 16384 here; N = 65536 overflows its int indices).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_WAVE = 0.4


def time_step(N):
    """dt = 1e-3 where RK4 is stable for it.  The nonlinear (advective) part of the Lagrangian system has
    eigenvalues ~ i * U * k_max with U ~ h the orbital velocity and k_max = (N/2)/(1-h) at the crest, so RK4 needs
    dt < 2.83 (1-h) / (h N/2): 1.3e-4 at N = 65536, h = 0.4 (measured: dt = 1e-3 blows up within 4 steps there, in the
    reference's arithmetic as much as here; dt = 1e-4 is stable)."""
    return 1e-3 if N <= 8192 else 1e-4


F_PAIR = 20.0   # algorithmic flops per pair evaluation-and-accumulate


def trochoid_state(N, h=H_WAVE):
    a = 2.0 * np.pi * np.arange(N) / N
    Z = (a - h * np.sin(a)) + 1j * (h * np.cos(a))
    Phi = h * np.sin(a)
    return np.concatenate([Z, Phi.astype(np.complex128)])


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port timed on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline(N, budget_s=20.0):
    """Reference CPU arithmetic (direct 1/tan assembly of M and V1, LAPACK LU, BLAS mat-vec) per RK4 step at size N.

    N <= 1024: whole steps are run.  Larger N: the three cost components are timed on a bounded sample and scaled:
    assembly on `rows` rows of the N x N matrices, LU at n_lu and scaled by (N/n_lu)^3, mat-vec by (N/n_lu)^2."""
    from oracle import roberts_oracle as ro
    cores = os.cpu_count() or 1
    props = ro.ProblemProperties(rho=0.0)
    Z, Phi = ro.trochoid(N, H_WAVE)
    if N <= 1024:
        y = ro.pack_state(Z, Phi)
        f = lambda s: ro.rhs(s, N, 1, props, "water", "cuda")
        ro.rk4_step(f, y, time_step(N))
        t0, n = time.perf_counter(), 0
        while True:
            y = ro.rk4_step(f, y, time_step(N))
            n += 1
            if time.perf_counter() - t0 > min(budget_s, 10.0) or n >= 50:
                break
        dt = (time.perf_counter() - t0) / n
        return dict(value=1.0 / dt, unit="steps/s", cores=cores, kind="port",
                    sample=f"{n} full RK4 steps at N={N} (NumPy oracle: np.tan assembly, LAPACK LU, BLAS mat-vec)")
    rows = max(8, min(256, int(4.0e6 // N)))
    x = np.cos(2 * np.pi * np.arange(N) / N)
    r = np.arange(rows) * (N // rows)
    t0 = time.perf_counter()
    ro.cot_rowsum(Z, x, r)
    t_rows = time.perf_counter() - t0
    t_assembly = t_rows * (N / rows) * 3.0          # M, V1 lower, V1 upper are each assembled per RHS (L/BaseBoundaryIntegrator.cuh:240-242, 295)
    n_lu = 2048
    A = np.random.default_rng(0).standard_normal((n_lu, n_lu)) + n_lu * np.eye(n_lu)
    b = np.ones(n_lu)
    np.linalg.solve(A[:256, :256], b[:256])
    t0 = time.perf_counter()
    np.linalg.solve(A, b)
    t_lu = (time.perf_counter() - t0) * (N / n_lu) ** 3
    Ac = A.astype(np.complex128)
    t0 = time.perf_counter()
    Ac @ b.astype(np.complex128)
    t_mv = (time.perf_counter() - t0) * (N / n_lu) ** 2 * 2.0
    t_step = 4.0 * (t_assembly + t_lu + t_mv)
    return dict(value=1.0 / t_step, unit="steps/s", cores=cores, kind="port",
                sample=(f"extrapolated from: direct 1/tan assembly of {rows} of {N} rows ({t_rows:.2f} s), LAPACK LU at n={n_lu} "
                        f"scaled by (N/n)^3, complex mat-vec at n={n_lu} scaled by (N/n)^2; per step = 4 x (3 assemblies + LU + 2 mat-vec)"))


def reference_gpu_rates(sizes):
    """RK4 steps/s of the reference's own CUDA path (oracle/_ref, child process, same GPU) on the bench surface at each N."""
    from oracle import ref_runner
    if not ref_runner.available():
        return {}
    jobs = [dict(op="rk4", kind="water", N=n, props=dict(rho=0.0), state=trochoid_state(n), dt=time_step(n),
                 warmup=2 if n <= 4096 else 1, steps=8 if n <= 4096 else 2) for n in sizes]
    out = {}
    for j, r in zip(jobs, ref_runner.run_jobs(jobs, timeout=240)):
        if "error" in r:
            out[j["N"]] = {"error": r["error"][:200]}
        else:
            out[j["N"]] = {"steps_per_s": j["steps"] / r["seconds"], "steps": j["steps"], "warmup": j["warmup"],
                           "finite": bool(np.isfinite(r["state"]).all())}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.n
    vals = []
    base = None
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(N, budget_s=8.0)
        if i >= args.warmup:
            vals.append(base["value"])
        if time.perf_counter() - t_all > 150:
            break
    v = float(np.mean(vals)) if vals else base["value"]
    # context only (not this line's value): the reference's own CUDA path at the largest bench size it supports, if a GPU is here
    ref_cuda = None
    if not args.no_reference_gpu:
        r = reference_gpu_rates([4096]).get(4096)
        if r:
            ref_cuda = dict(r, N=4096, what="reference's own CUDA path (oracle/_ref) on this box's GPU 0; it cannot run N=65536 "
                                            "(int indices overflow at n >= 46341, L/createM.cuh:52)")
    line = dict(metric=f"RK4 steps/s at N={N}", value=v, unit="steps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 / v, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=workload_config(N),
                cpu_baseline=dict(base, value=v), reference_cuda=ref_cuda,
                e2e=dict(value=v, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(N):
    return {"workload": f"water (rho=0) trochoidal Stokes surface h={H_WAVE}, N={N}, batch=1, dt={time_step(N)}, classical RK4, "
                        f"FP64, matrix-free Richardson solve tol=1e-13 (warm start), trajectory logging off",
            "l2": "L2 flushed (256 MiB write) between timed steps; each step timed with its own CUDA event pair"}


# ---------------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 <= t <= t1:
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(l.split(",")[1]) for _, l in self.lines[-3:] if len(l.split(",")) > 2] or [float("nan")]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    from superfluid_dynamics_b200 import _lib, api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE JSON line on stdout: NCCL's version banner (printed at VERSION and at WARN level) and anything else it
        # logs go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        sys.stdout.flush()
        saved_stdout = os.dup(1)      # belt and braces: whatever a native library writes to fd 1 during the run lands on stderr
        os.dup2(2, 1)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    # a non-default stream: the library launches (and graph-captures) on torch's current stream, and the CUDA events below are
    # recorded on that same stream
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    lib = _lib.load()
    N = args.n
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
    if world > 1:
        calc.initComm(rank, world)
    stepper = api.AutonomousRungeKuttaStepper(calc, time_step(N))
    y0 = trochoid_state(N)
    state = torch.as_tensor(y0, device=dev)
    stepper.initialize(state, True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        stepper.runStep()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.rb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    it0 = calc.solve_stats()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        stepper.runStep()
        ev[i][1].record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = lib.rb_launch_count() - launches0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    final = stepper.getState()
    assert np.isfinite(final).all(), "state blew up"

    # ---- roofline of the dominant kernel (the sweep), measured live with CUDA events on the launching stream --------------
    sweep_ms, pairs = calc.benchSweep(torch.as_tensor(y0, device=dev), 30)
    peak = api.measure_fp64_peak(dev)
    achieved = F_PAIR * pairs / world / (sweep_ms * 1e-3) / 1e12
    it1 = calc.solve_stats()
    mv_per_rhs = (it1["total_iterations"] - it0["total_iterations"]) / max(1, it1["total_solves"] - it0["total_solves"])
    # all O(N^2) sweeps executed per step: solver sweeps (incl. the combined verify+velocity ones) + velocity-only sweeps
    sweeps_per_step = (it1["total_iterations"] - it0["total_iterations"] + it1["velocity_sweeps"] - it0["velocity_sweeps"]) / args.steps
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum)
    traffic = None
    tf = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tf) and world == 1:
        traffic = json.load(open(tf)).get(str(N), {}).get("dram_bytes_per_launch")
    kernel = "rb::sweep2_kernel<MV, 1> (persistent)" if N <= 1024 else \
        ("rb::sweep_kernel<MV, 4 rows/thread> (tiled)" if N >= 49152 else "rb::sweep_kernel<MV, 2 rows/thread> (tiled)")
    roofline = {"bound": "fp64", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic,
                "bound_note": "FP64 vector pipe (DFMA), the roofline north_star names for the O(N^2) summation; HBM traffic per launch "
                              "is ~1e-4 of what the HBM roofline would allow (working set lives in L2)",
                "peak_source": "measured live by this library's DFMA-only probe (MEASURED_PEAKS.json holds no FP64 figure); "
                               "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2",
                "algorithmic_flops_per_launch": F_PAIR * pairs / world, "launch_ms": sweep_ms,
                "sweeps_per_step": sweeps_per_step,
                "step_frac": F_PAIR * pairs * sweeps_per_step * value / 1e12 / (peak * world)}

    # ---- end to end through the public API with host buffers: every step H2D state, step, D2H state -------------------------
    e2e = None
    if rank == 0 or world > 1:
        host = torch.as_tensor(y0).pin_memory()
        st2 = torch.as_tensor(y0, device=dev)
        stepper.initialize(st2, True)
        for _ in range(max(2, args.warmup)):
            stepper.runStep()
        host.copy_(st2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            st2.copy_(host, non_blocking=True)
            stepper.runStep()
            host.copy_(st2, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        barrier()
        t_e2e = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        nbytes = host.numel() * 16
        e2e = {"value": args.steps / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes}

    if rank == 0:
        extra = {}
        if world == 1 and not args.no_extra:
            # one-shot C-ABI call with host buffers (solver construction + H2D + K steps + D2H inside the timed region)
            init = np.concatenate([y0[:N].real, y0[:N].imag, y0[N:].real])
            t0 = time.perf_counter()
            api.integrate_rk4_host(init, N, 1, props, "water", time_step(N), args.steps)
            extra["e2e_one_call_steps_per_s"] = args.steps / (time.perf_counter() - t0)
            # the other sizes: N = 4096 (the second size the metric names) and N = 16384, each next to the REFERENCE'S OWN CUDA path
            # (oracle/_ref/libcusuperhelium_ref.so: its classes compiled unmodified, run in a child process on this same GPU;
            # N = 65536 is beyond it: int indices overflow at n >= 46341, L/createM.cuh:52)
            ref_gpu = reference_gpu_rates([n for n in (4096, 16384) if n != N]) if not args.no_reference_gpu else {}
            for n2 in (4096, 16384, 65536):
                if n2 == N or (n2 == 65536 and N != 4096):
                    continue
                c2 = api.BaseBoundaryIntegralCalculator(n2, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
                s2 = api.AutonomousRungeKuttaStepper(c2, time_step(n2))
                s2.initialize(torch.as_tensor(trochoid_state(n2), device=dev), True)
                k2 = 100 if n2 <= 8192 else (30 if n2 <= 16384 else args.steps)
                s2.runSteps(12)
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                s2.runSteps(k2)
                b.record()
                torch.cuda.synchronize(dev)
                ms2, pr2 = c2.benchSweep(torch.as_tensor(trochoid_state(n2), device=dev), 20)
                extra[f"n{n2}"] = {"steps_per_s": k2 / (a.elapsed_time(b) * 1e-3), "steps": k2, "sweep_ms": ms2,
                                   "sweep_tflops": F_PAIR * pr2 / (ms2 * 1e-3) / 1e12,
                                   "sweep_frac_of_fp64_peak": F_PAIR * pr2 / (ms2 * 1e-3) / 1e12 / peak}
                r = ref_gpu.get(n2)
                if r and "steps_per_s" in r:
                    extra[f"n{n2}"]["reference_cuda_steps_per_s"] = r["steps_per_s"]
                    extra[f"n{n2}"]["speedup_vs_reference_cuda"] = extra[f"n{n2}"]["steps_per_s"] / r["steps_per_s"]
                elif r:
                    extra[f"n{n2}"]["reference_cuda"] = r
                del s2, c2
            # BASELINE config 5, second half: the 1024-member ensemble at N = 512 (member m: trochoid h_m = 0.05 + 0.35 m / 1023,
            # SURVEY.md section 8d), batched in one solver; across GPUs the members are replicas, no communication.  An extra: a
            # failure here is recorded, it never takes the headline down.
            try:
                Ne, Be, ke = 512, 1024, 30
                hs = 0.05 + 0.35 * np.arange(Be) / (Be - 1)
                members = [trochoid_state(Ne, h) for h in hs]
                ye = np.concatenate([m_[:Ne] for m_ in members] + [m_[Ne:] for m_ in members])   # [Z of every member | Phi of every member]
                ce = api.BaseBoundaryIntegralCalculator(Ne, Be, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
                se = api.AutonomousRungeKuttaStepper(ce, 1e-3)
                ste = torch.as_tensor(ye, device=dev)
                se.initialize(ste, True)
                se.runSteps(12)
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                se.runSteps(ke)
                b.record()
                torch.cuda.synchronize(dev)
                finite = bool(torch.isfinite(torch.view_as_real(ste)).all())
                extra["ensemble_1024xN512"] = {"steps_per_s": ke / (a.elapsed_time(b) * 1e-3),
                                               "member_steps_per_s": Be * ke / (a.elapsed_time(b) * 1e-3), "steps": ke,
                                               "finite": finite, "converged": bool(ce.solve_stats()["converged"])}
                del se, ce, ste
            except Exception as e:  # noqa: BLE001
                extra["ensemble_1024xN512"] = {"error": repr(e)[:200]}
            if ref_gpu:
                extra["reference_cuda_note"] = ("reference_cuda_* = the reference's own CUDA path (BaseBoundaryIntegralCalculator + "
                                                "AutonomousRungeKuttaStepper compiled unmodified from its sources, oracle/build_ref.py) "
                                                "on this same GPU, same surface and dt, host clock around the steps after warm-up")
        cpu = cpu_baseline(N) if (world == 1 and not args.no_cpu) else None
        line = dict(metric=f"RK4 steps/s at N={N}", value=value, unit="steps/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f64", data="synthetic", config=workload_config(N), roofline=roofline, cpu_baseline=cpu, e2e=e2e,
                    gpu_launches=int(launches), clocks=clocks, solver_iterations_per_rhs=mv_per_rhs,
                    **extra)
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the N=4096 / N=16384 and one-call extras")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the compiled reference CUDA path (oracle/_ref)")
    args = ap.parse_args()
    # the stepper tunes the number of recorded sweeps and fills its 4-step stage history during the first steps: warm up past that
    args.warmup = max(args.warmup, 12) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
