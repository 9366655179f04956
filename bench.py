#!/usr/bin/env python
"""bench.py -- RK4 steps/s of the Roberts boundary-integral step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 65536] [--impl native|reference]

A "step" is one classical RK4 step (4 RHS evaluations: FFT derivatives, matrix-free solve for the vortex-sheet strength,
O(N^2) velocity summation, fused stage update) of one synthetic trochoidal ("Stokes") surface, h = 0.4; dt = 1e-3 for N <= 8192 and
1e-4 above (RK4's stability limit for this system, see time_step()).  Prints ONE JSON line.  Under torchrun (N > 1) the row blocks of
every O(N^2) sweep are sharded over the ranks.  Besides the headline (N = 65536) the line carries, measured in the same run and at
the same rank count: N = 4096 (the other size the metric names) and N = 16384 next to the reference's own CUDA path, the helium
film of BASELINE config 4 (N = 16384, finite depth, row-sharded), the 1024 x N=512 ensemble of config 5 (members spread over the
ranks), the dense-LU solve mode at N = 4096 (config 3), HBM figures of the memory-bound kernels, and correctness evidence for the
timed run itself (replicas identical, sharded vs single-GPU difference, solve status).
--impl reference times the CPU statement of the same path (oracle port of the reference's NumPy/LAPACK arithmetic) on a bounded
sample of the same workload on the host cores.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_WAVE = 0.4
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 148 SMs x 64 DFMA / clk x 2 flop x 1.965 GHz = 37.2
HELIUM_DEPTH = 0.0942478                                  # A/kernel.cu:77-82 film depth (BASELINE config 4)


def time_step(N):
    """dt = 1e-3 where RK4 is stable for it.  The nonlinear (advective) part of the Lagrangian system has
    eigenvalues ~ i * U * k_max with U ~ h the orbital velocity and k_max = (N/2)/(1-h) at the crest, so RK4 needs
    dt < 2.83 (1-h) / (h N/2): 1.3e-4 at N = 65536, h = 0.4 (measured: dt = 1e-3 blows up within 4 steps there, in the
    reference's arithmetic as much as here; dt = 1e-4 is stable; profiles/r02_stability.log)."""
    return 1e-3 if N <= 8192 else 1e-4


F_PAIR = 20.0   # algorithmic flops per pair evaluation-and-accumulate; 40 with the finite-depth image term


def trochoid_state(N, h=H_WAVE):
    a = 2.0 * np.pi * np.arange(N) / N
    Z = (a - h * np.sin(a)) + 1j * (h * np.cos(a))
    Phi = h * np.sin(a)
    return np.concatenate([Z, Phi.astype(np.complex128)])


def film_state(N, depth=HELIUM_DEPTH, amp=0.1):
    a = 2.0 * np.pi * np.arange(N) / N
    return np.concatenate([a + 1j * amp * depth * np.cos(a), np.zeros(N, np.complex128)])


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port timed on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------------------------
def cpu_measured(budget_s=12.0):
    """MEASURED (not extrapolated) CPU legs of the same arithmetic, sized to run in seconds: BASELINE config 1 (N = 64, whole RK4
    steps) and one full RHS at N = 4096 (direct 1/tan assembly of M and both V1, LAPACK LU, BLAS mat-vec: the reference's CPU path
    as the NumPy oracle restates it)."""
    from oracle import roberts_oracle as ro
    out = {}
    props = ro.ProblemProperties(rho=0.0)
    N = 64
    y = ro.pack_state(*ro.trochoid(N, 0.1))
    f = lambda s: ro.rhs(s, N, 1, props, "water", "cuda")
    ro.rk4_step(f, y, 1e-3)
    t0, n = time.perf_counter(), 0
    while n < 100 and time.perf_counter() - t0 < 3.0:
        y = ro.rk4_step(f, y, 1e-3)
        n += 1
    out["n64_steps_per_s"] = n / (time.perf_counter() - t0)
    out["n64_sample"] = f"{n} whole RK4 steps, N=64 trochoid h=0.1, dt=1e-3 (BASELINE config 1), vectorised NumPy oracle"
    N = 4096
    Z, Phi = ro.trochoid(N, H_WAVE)
    t0 = time.perf_counter()
    ro.rhs_single(Z, Phi, props, "water", "cuda")
    t = time.perf_counter() - t0
    out["n4096_rhs_s"] = t
    out["n4096_steps_per_s"] = 1.0 / (4.0 * t)
    out["n4096_sample"] = "ONE full RHS at N=4096 (np.tan assembly of M, V1 lower and upper, LAPACK LU, BLAS mat-vec), x4 per RK4 step"
    return out


def cpu_baseline(N, budget_s=20.0, measured=True):
    """Reference CPU arithmetic (direct 1/tan assembly of M and V1, LAPACK LU, BLAS mat-vec) per RK4 step at size N.

    N <= 1024: whole steps are run.  Larger N: the three cost components are timed on a bounded sample and scaled:
    assembly on `rows` rows of the N x N matrices, LU at n_lu and scaled by (N/n_lu)^3, mat-vec by (N/n_lu)^2 -- at N = 65536 one
    real RHS of this algorithm is 1.9e14 flops of LU on a 34 GB matrix, hours of CPU time, so the headline's CPU figure is by
    construction an EXTRAPOLATION (flagged as such); the measured legs next to it (cpu_measured) are real runs."""
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
        return dict(value=1.0 / dt, unit="steps/s", cores=cores, kind="port", extrapolated=False,
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
    out = dict(value=1.0 / t_step, unit="steps/s", cores=cores, kind="port", extrapolated=True,
               sample=(f"EXTRAPOLATED from: direct 1/tan assembly of {rows} of {N} rows ({t_rows:.2f} s), LAPACK LU at n={n_lu} "
                       f"scaled by (N/n)^3, complex mat-vec at n={n_lu} scaled by (N/n)^2; per step = 4 x (3 assemblies + LU + 2 mat-vec); "
                       f"a real run of this O(N^3) algorithm at N={N} would take hours"))
    if measured:
        try:
            out["measured"] = cpu_measured()
        except Exception as e:  # noqa: BLE001
            out["measured"] = {"error": repr(e)[:200]}
    return out


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
        base = cpu_baseline(N, budget_s=8.0, measured=False)
        if i >= args.warmup:
            vals.append(base["value"])
        if time.perf_counter() - t_all > 150:
            break
    v = float(np.mean(vals)) if vals else base["value"]
    try:
        base["measured"] = cpu_measured()
    except Exception as e:  # noqa: BLE001
        base["measured"] = {"error": repr(e)[:200]}
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
                note=("value is an extrapolation of the reference's O(N^3) CPU algorithm from a bounded sample (see cpu_baseline.sample); "
                      "cpu_baseline.measured holds real runs at N=64 and N=4096"),
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
class Ctx:
    """rank / device / collectives of this process"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.saved_stdout = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # rank 0 prints ONE JSON line on stdout: NCCL's version banner (printed at VERSION and at WARN level) and anything else it
            # logs go to stderr
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
                os.environ.pop("NCCL_DEBUG")
            sys.stdout.flush()
            self.saved_stdout = os.dup(1)      # belt and braces: whatever a native library writes to fd 1 during the run lands on stderr
            os.dup2(2, 1)
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
        self.dev = torch.device(f"cuda:{self.local}")
        torch.cuda.set_device(self.dev)
        # a non-default stream: the library launches (and graph-captures) on torch's current stream, and the CUDA events below are
        # recorded on that same stream
        torch.cuda.set_stream(torch.cuda.Stream(device=self.dev))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def replica_diff(self, state):
        """max |state on this rank - state on rank 0| over all ranks (row-sharded replicas must be bit-identical: 0.0)"""
        if self.world == 1:
            return 0.0
        ref = state.clone()
        self.dist.broadcast(ref, 0)
        d = (self.torch.view_as_real(ref) - self.torch.view_as_real(state)).abs().max()
        return self.max_over_ranks(float(d.item()))

    def emit(self, line):
        if self.world > 1:
            sys.stdout.flush()
            os.dup2(self.saved_stdout, 1)
        print(json.dumps(line), flush=True)


def timed_steps(ctx, stepper, steps, flush=None):
    """K steps, each bracketed by its own CUDA event pair on the launching stream (L2 flushed in between when asked); returns
    ms per step (max over ranks) and the wall-clock window."""
    torch = ctx.torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev[i][0].record()
        stepper.runStep()
        ev[i][1].record()
    ctx.barrier()
    t1 = time.perf_counter()
    ms = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    return ms / steps, t0, t1


def batch_rate(ctx, stepper, steps):
    """steps/s of runSteps(steps) between one CUDA event pair (no flush): how a simulation actually runs"""
    torch = ctx.torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    a.record()
    stepper.runSteps(steps)
    b.record()
    ctx.barrier()
    return steps / (ctx.max_over_ranks(a.elapsed_time(b)) * 1e-3)


def status_of(calc):
    s = calc.solve_stats()
    return {k: s[k] for k in ("converged", "stagnated", "stagnated_solves", "failed_solves", "worst_residual")}


def leg_sharded_water(ctx, api, N, steps, warm=30, compare_single=True, ref_gpu=None, peak=None, dt=None):
    """Water surface of size N, row-sharded over the ranks of this run: steps/s (runSteps between one event pair), sweep roofline,
    replicas identical, difference to a single-GPU run of the same step count (rank 0)."""
    torch = ctx.torch
    props = api.ProblemProperties(rho=0.0)
    dt = time_step(N) if dt is None else dt
    y0 = trochoid_state(N)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=ctx.dev, guess="warm")
    if ctx.world > 1:
        calc.initComm(ctx.rank, ctx.world)
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    st = torch.as_tensor(y0, device=ctx.dev)
    stp.initialize(st, True)
    stp.runSteps(warm)
    it0 = calc.solve_stats()
    rate = batch_rate(ctx, stp, steps)
    it1 = calc.solve_stats()
    out = {"steps_per_s": rate, "steps": steps, "dt": dt,
           "sweeps_per_step": (it1["total_iterations"] - it0["total_iterations"] + it1["velocity_sweeps"] - it0["velocity_sweeps"]) / steps,
           "replica_max_abs_diff": ctx.replica_diff(st), "status": status_of(calc), "plan": calc.sweepPlan()}
    final = st.cpu().numpy()
    ms, pairs = calc.benchSweep(torch.as_tensor(y0, device=ctx.dev), 30 if N <= 16384 else 12)
    ms = ctx.max_over_ranks(ms)
    out["sweep_ms"] = ms
    out["sweep_tflops_per_gpu"] = F_PAIR * pairs / ctx.world / (ms * 1e-3) / 1e12
    if peak:
        out["sweep_frac_of_fp64_peak"] = out["sweep_tflops_per_gpu"] / peak
        out["sweep_frac_of_fp64_nominal"] = out["sweep_tflops_per_gpu"] / FP64_NOMINAL_TFLOPS
    del stp, calc
    if ctx.world > 1 and compare_single and ctx.rank == 0:
        c1 = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=ctx.dev, guess="warm")
        s1 = api.AutonomousRungeKuttaStepper(c1, dt)
        s1.initialize(y0, False)
        s1.runSteps(warm + steps)
        single = s1.getState()
        out["shard_vs_single_rel"] = float(np.abs(final - single).max() / np.abs(single).max())
        del s1, c1
    ctx.barrier()
    if ref_gpu and "steps_per_s" in ref_gpu:
        out["reference_cuda_steps_per_s"] = ref_gpu["steps_per_s"]
        out["speedup_vs_reference_cuda"] = rate / ref_gpu["steps_per_s"]
    elif ref_gpu:
        out["reference_cuda"] = ref_gpu
    return out


def leg_helium(ctx, api, peak, N=16384, steps=10, warm=30):
    """BASELINE config 4: helium film with van-der-Waals forcing, finite depth (image term, F_pair = 40), N = 16384, row-sharded."""
    torch = ctx.torch
    props = api.ProblemProperties(rho=0.0, depth=HELIUM_DEPTH)
    dt = 1e-3
    y0 = film_state(N)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), device=ctx.dev, guess="warm")
    if ctx.world > 1:
        calc.initComm(ctx.rank, ctx.world)
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    st = torch.as_tensor(y0, device=ctx.dev)
    stp.initialize(st, True)
    stp.runSteps(warm)
    it0 = calc.solve_stats()
    rate = batch_rate(ctx, stp, steps)
    it1 = calc.solve_stats()
    sweeps = (it1["total_iterations"] - it0["total_iterations"] + it1["velocity_sweeps"] - it0["velocity_sweeps"]) / steps
    out = {"steps_per_s": rate, "steps": steps, "dt": dt, "depth": HELIUM_DEPTH, "amplitude": 0.1 * HELIUM_DEPTH, "sweeps_per_step": sweeps,
           "replica_max_abs_diff": ctx.replica_diff(st), "status": status_of(calc), "plan": calc.sweepPlan(),
           "graph": stp.stats()}
    final = st.cpu().numpy()
    ms, pairs = calc.benchSweep(torch.as_tensor(y0, device=ctx.dev), 20)      # pairs counts the image pairs too (2 N^2)
    ms = ctx.max_over_ranks(ms)
    tf = F_PAIR * pairs / ctx.world / (ms * 1e-3) / 1e12
    out.update(sweep_ms=ms, sweep_tflops_per_gpu=tf, sweep_frac_of_fp64_peak=tf / peak, sweep_frac_of_fp64_nominal=tf / FP64_NOMINAL_TFLOPS,
               step_frac_of_fp64_peak=F_PAIR * pairs * sweeps * rate / 1e12 / (peak * ctx.world),
               flops_per_pair=2 * F_PAIR)
    del stp, calc
    if ctx.world > 1 and ctx.rank == 0:
        c1 = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), device=ctx.dev, guess="warm")
        s1 = api.AutonomousRungeKuttaStepper(c1, dt)
        s1.initialize(y0, False)
        s1.runSteps(warm + steps)
        single = s1.getState()
        out["shard_vs_single_rel"] = float(np.abs(final - single).max() / np.abs(single).max())
        del s1, c1
    ctx.barrier()
    return out


def leg_ensemble(ctx, api, peak, Ne=512, Be=1024, steps=30, warm=30):
    """BASELINE config 5, second half: the 1024-member ensemble at N = 512 (member m: trochoid h_m = 0.05 + 0.35 m / 1023,
    SURVEY.md section 8d); the members are spread over the ranks (replicas only: no communication), each rank steps its share in
    one batched solver.  Rank 0's first member is checked against the same member stepped alone."""
    torch = ctx.torch
    props = api.ProblemProperties(rho=0.0)
    lo, hi = api.ensemble_member_range(Be, ctx.rank, ctx.world)
    hs = 0.05 + 0.35 * np.arange(Be) / (Be - 1)
    members = [trochoid_state(Ne, h) for h in hs[lo:hi]]
    calc = api.BaseBoundaryIntegralCalculator(Ne, hi - lo, props, api.WaterBoundaryProblem(props), device=ctx.dev, guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, 1e-3)
    st = torch.as_tensor(api.ensemble_state(members, Ne), device=ctx.dev)
    stp.initialize(st, True)
    stp.runSteps(warm)
    it0 = calc.solve_stats()
    rate = batch_rate(ctx, stp, steps)
    it1 = calc.solve_stats()
    sweeps = (it1["total_iterations"] - it0["total_iterations"] + it1["velocity_sweeps"] - it0["velocity_sweeps"]) / steps
    finite = bool(torch.isfinite(torch.view_as_real(st)).all())
    flops_step = F_PAIR * Ne * Ne * Be * sweeps
    out = {"steps_per_s": rate, "member_steps_per_s": Be * rate, "steps": steps, "members": Be, "members_per_rank": hi - lo, "N": Ne,
           "sweeps_per_step": sweeps, "finite": finite, "status": status_of(calc), "plan": calc.sweepPlan(),
           "step_tflops_per_gpu": flops_step * rate / 1e12 / ctx.world,
           "step_frac_of_fp64_peak": flops_step * rate / 1e12 / (peak * ctx.world)}
    if ctx.rank == 0:
        got = st.cpu().numpy()
        nb = hi - lo
        mine = np.concatenate([got[:Ne], got[nb * Ne:nb * Ne + Ne]])
        alone = api.BaseBoundaryIntegralCalculator(Ne, 1, props, api.WaterBoundaryProblem(props), device=ctx.dev, guess="warm")
        s1 = api.AutonomousRungeKuttaStepper(alone, 1e-3)
        s1.initialize(members[0], False)
        s1.runSteps(warm + steps)
        out["first_member_vs_stepped_alone_rel"] = float(np.abs(mine - s1.getState()).max() / np.abs(mine).max())
        del s1, alone
    del stp, calc, st
    ctx.barrier()
    return out


def leg_dense_mode(ctx, api, N=4096, reps=3):
    """BASELINE config 3 ("dense FP64 BIE solve", N = 4096): RHS with M materialised in HBM and factorised (RB_SOLVE_DENSE_LU: assembly
    kernel + blocked LU with FP64 tensor-core trailing updates), against the matrix-free RHS and the reference's own CUDA path
    (createMKernel + cuSOLVER getrf/getrs, L/MatrixSolver.cuh:114-125)."""
    torch = ctx.torch
    props = api.ProblemProperties(rho=0.0)
    st = torch.as_tensor(trochoid_state(N), device=ctx.dev)
    res, outs = {}, {}
    for mode in ("dense_lu", "matrix_free"):
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=ctx.dev, solve_mode=mode)
        out = torch.zeros(2 * N, dtype=torch.complex128, device=ctx.dev)
        calc.run(st, out)
        torch.cuda.synchronize(ctx.dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            calc.run(st, out)
        b.record()
        torch.cuda.synchronize(ctx.dev)
        res[mode + "_rhs_ms"] = a.elapsed_time(b) / reps
        outs[mode] = out.cpu().numpy()
        del calc
    res["dense_steps_per_s_equivalent"] = 1e3 / (4.0 * res["dense_lu_rhs_ms"])
    res["dense_vs_matrix_free_rel"] = float(np.abs(outs["dense_lu"] - outs["matrix_free"]).max() / np.abs(outs["matrix_free"]).max())
    res["lu_flops_per_rhs"] = 2.0 / 3.0 * N ** 3
    return res


def leg_hbm_kernels(ctx, api, hbm_peak_gbs):
    """The HBM-bound kernels north_star asks GB/s for: materialised assembly (create_M 8 B/entry, velocity matrices 16 B/entry,
    N = 4096) and the solve-closing / stage-update kernels at the ensemble's B N = 524288 (measured through a whole ensemble RHS
    is not possible in isolation: the update kernels are timed through their C-ABI entry points on 2 B N complex values)."""
    import ctypes
    torch = ctx.torch
    lib = api._lib.load()
    out = {}
    N = 4096
    from superfluid_dynamics_b200 import api as _api
    props = _api.ProblemProperties(rho=0.0)
    calc = _api.BaseBoundaryIntegralCalculator(N, 1, props, _api.WaterBoundaryProblem(props), device=ctx.dev)
    y = torch.as_tensor(trochoid_state(N), device=ctx.dev)
    Z, Phi = y[:N].contiguous(), y[N:].contiguous()
    Zp, PhiP, Zpp = calc.zPhiDerivative(Z, Phi)
    A = torch.empty(N * N, dtype=torch.float64, device=ctx.dev)
    V1 = torch.empty(N * N, dtype=torch.complex128, device=ctx.dev)
    V2 = torch.empty(N, dtype=torch.complex128, device=ctx.dev)

    def timeit(fn, reps=20):
        fn()
        torch.cuda.synchronize(ctx.dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize(ctx.dev)
        return a.elapsed_time(b) / reps

    ms = timeit(lambda: _api.createMKernel(A, Z, Zp, Zpp, 0.0, N))
    out["create_M_N4096"] = {"ms": ms, "algorithmic_bytes": 8.0 * N * N, "gbs": 8.0 * N * N / (ms * 1e-3) / 1e9}
    ms = timeit(lambda: _api.createVelocityMatrices(Z, Zp, Zpp, N, V1, V2, True))
    out["velocity_matrices_N4096"] = {"ms": ms, "algorithmic_bytes": 16.0 * N * N, "gbs": 16.0 * N * N / (ms * 1e-3) / 1e9}
    n = 2 * 524288                       # [Z | Phi] of the 1024 x N=512 ensemble
    ys = [torch.randn(n, dtype=torch.complex128, device=ctx.dev) for _ in range(6)]
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    sp = ctypes.c_void_p(torch.cuda.current_stream(ctx.dev).cuda_stream)
    ms = timeit(lambda: lib.rb_rk4_stage_update(p(ys[5]), p(ys[0]), p(ys[1]), 0.5e-3, n, sp))
    out["stage_update_BN524288"] = {"ms": ms, "algorithmic_bytes": 3.0 * 16 * n, "gbs": 3.0 * 16 * n / (ms * 1e-3) / 1e9}
    ms = timeit(lambda: lib.rb_rk4_final_update(p(ys[0]), p(ys[1]), p(ys[2]), p(ys[3]), p(ys[4]), 1e-3, n, sp))
    out["final_update_BN524288"] = {"ms": ms, "algorithmic_bytes": 6.0 * 16 * n, "gbs": 6.0 * 16 * n / (ms * 1e-3) / 1e9}
    for v in out.values():
        v["frac_of_hbm_peak"] = v["gbs"] / hbm_peak_gbs
    out["hbm_peak_gbs"] = hbm_peak_gbs
    out["note"] = ("stage/final update working sets (48 / 100 MB) fit the 126 MB L2: figures above the HBM peak are L2 bandwidth; "
                   "the assembly kernels write 134 / 268 MB and are bound by their N^2 complex tan evaluations, not by HBM")
    return out


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        return 6446.0     # the fallback B200_PROFILING.md states


def run_native(args):
    ctx = Ctx()
    torch = ctx.torch
    from superfluid_dynamics_b200 import _lib, api
    lib = _lib.load()
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
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

    for _ in range(args.warmup):
        stepper.runStep()
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.rb_launch_count()
    it0 = calc.solve_stats()
    ms_per_step, t_wall0, t_wall1 = timed_steps(ctx, stepper, args.steps, flush)
    launches = lib.rb_launch_count() - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = 1e3 / ms_per_step
    it1 = calc.solve_stats()
    final = stepper.getState()
    assert np.isfinite(final).all(), "state blew up"
    # ---- correctness evidence for the timed run itself -----------------------------------------------------------------------
    verify = {"state_sha256": hashlib.sha256(final.tobytes()).hexdigest()[:16], "replica_max_abs_diff": ctx.replica_diff(state),
              "status": status_of(calc), "steps_total": args.warmup + args.steps}
    if world > 1 and rank == 0 and not args.no_verify:
        c1 = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
        s1 = api.AutonomousRungeKuttaStepper(c1, time_step(N))
        s1.initialize(y0, False)
        s1.runSteps(args.warmup + args.steps)
        single = s1.getState()
        verify["shard_vs_single_rel"] = float(np.abs(final - single).max() / np.abs(single).max())
        del s1, c1
    ctx.barrier()

    # ---- roofline of the dominant kernel (the sweep), measured live with CUDA events on the launching stream --------------
    sweep_ms, pairs = calc.benchSweep(torch.as_tensor(y0, device=dev), 30)
    sweep_ms = ctx.max_over_ranks(sweep_ms)
    peak = api.measure_fp64_peak(dev)
    achieved = F_PAIR * pairs / world / (sweep_ms * 1e-3) / 1e12
    mv_per_rhs = (it1["total_iterations"] - it0["total_iterations"]) / max(1, it1["total_solves"] - it0["total_solves"])
    # all O(N^2) sweeps executed per step: solver sweeps (incl. the combined verify+velocity ones) + velocity-only sweeps
    sweeps_per_step = (it1["total_iterations"] - it0["total_iterations"] + it1["velocity_sweeps"] - it0["velocity_sweeps"]) / args.steps
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum)
    traffic = None
    tf = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tf) and world == 1:
        traffic = json.load(open(tf)).get(str(N), {}).get("dram_bytes_per_launch")
    plan = calc.sweepPlan()
    kernel = (f"rb::sweep2_kernel<MV> (persistent, {plan['ctas']} CTAs x {plan['threads']} threads)" if plan["kernel"] == "persistent" else
              f"rb::sweep3_kernel<MV, {plan['rows_per_thread']} rows/warp> (one warp per row group, {plan['ctas']} CTAs x {plan['threads']} threads)"
              if plan["kernel"] == "warp_rows" else
              f"rb::sweep_kernel<MV, {plan['rows_per_thread']} rows/thread> (tiled: {plan['row_cells']} row cells x {plan['nchunks']} source chunks "
              f"= {plan['ctas']} CTAs x {plan['threads']} threads)")
    step_ms_sweeps = sweeps_per_step * sweep_ms
    roofline = {"bound": "fp64", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS, "peak_nominal": FP64_NOMINAL_TFLOPS,
                "traffic": traffic,
                "bound_note": "FP64 vector pipe (DFMA), the roofline north_star names for the O(N^2) summation; HBM traffic per launch "
                              "is ~1e-4 of what the HBM roofline would allow (working set lives in L2)",
                "peak_source": "frac: of this library's live DFMA-only probe (MEASURED_PEAKS.json holds no FP64 figure); "
                               "frac_of_nominal: of 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s",
                "algorithmic_flops_per_launch": F_PAIR * pairs / world, "launch_ms": sweep_ms,
                "sweeps_per_step": sweeps_per_step,
                "step_frac": F_PAIR * pairs * sweeps_per_step * value / 1e12 / (peak * world),
                # what is left of a step besides its sweeps: replicated O(N log N) work (FFT derivatives, element-wise kernels) and, on a
                # row-sharded run, the flag waits of the exchanges (8 per step)
                "step_ms_in_sweeps": step_ms_sweeps, "step_ms_residual": ms_per_step - step_ms_sweeps}

    # ---- end to end through the public API with host buffers: every step H2D state, step, D2H state -------------------------
    host = torch.as_tensor(y0).pin_memory()
    st2 = torch.as_tensor(y0, device=dev)
    stepper.initialize(st2, True)
    for _ in range(max(2, args.warmup)):
        stepper.runStep()
    host.copy_(st2)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st2.copy_(host, non_blocking=True)
        stepper.runStep()
        host.copy_(st2, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ctx.barrier()
    t_e2e = ctx.max_over_ranks(time.perf_counter() - t0)
    nbytes = host.numel() * 16
    e2e = {"value": args.steps / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes}
    del stepper, calc
    ctx.barrier()

    line = dict(metric=f"RK4 steps/s at N={N}", value=value, unit="steps/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="f64", data="synthetic", config=workload_config(N), roofline=roofline, cpu_baseline=None, e2e=e2e,
                gpu_launches=int(launches), clocks=clocks, solver_iterations_per_rhs=mv_per_rhs, verify=verify)

    # ---- the other configs, at this run's rank count.  Every leg is an extra: a failure is recorded, and a watchdog prints the
    # headline without them should one hang (a dead peer inside a collective), so that they can never take the headline down ----
    done = threading.Event()
    if not args.no_extra:
        def watchdog():
            if not done.wait(args.extra_timeout):
                if rank == 0:
                    line["extras_timed_out_after_s"] = args.extra_timeout
                    ctx.emit(line)
                os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()

        def leg(name, fn):
            t0 = time.perf_counter()
            try:
                r = fn()
                if rank == 0:
                    r["leg_seconds"] = time.perf_counter() - t0
                    line[name] = r
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    line[name] = {"error": repr(e)[:300]}
            if world > 1:
                # keep the ranks together whatever happened on one of them (an exception on one rank only would otherwise leave the
                # others inside the next leg's first collective)
                ctx.barrier()

        if world == 1:
            # one-shot C-ABI call with host buffers (solver construction + H2D + K steps + D2H inside the timed region)
            init = np.concatenate([y0[:N].real, y0[:N].imag, y0[N:].real])
            t0 = time.perf_counter()
            api.integrate_rk4_host(init, N, 1, props, "water", time_step(N), args.steps)
            line["e2e_one_call_steps_per_s"] = args.steps / (time.perf_counter() - t0)
        # the reference's OWN CUDA path (oracle/_ref/libcusuperhelium_ref.so: its classes compiled unmodified, run in a child process
        # on this same GPU; N = 65536 is beyond it: int indices overflow at n >= 46341, L/createM.cuh:52)
        ref_gpu = reference_gpu_rates([4096, 16384]) if (rank == 0 and world == 1 and not args.no_reference_gpu) else {}
        for n2 in (4096, 16384):
            if n2 != N and n2 // 256 >= world:
                leg(f"n{n2}", lambda n2=n2: leg_sharded_water(ctx, api, n2, 200 if n2 <= 8192 else 40, ref_gpu=ref_gpu.get(n2), peak=peak))
        if world == 1 and N == 65536:
            # the headline's dt = 1e-4 is a round number below RK4's stability limit 2.83 (1 - h) / (h N / 2) = 1.3e-4 for this surface: the
            # same run at 1.25e-4, the largest stable round step (the sweeps per RHS, hence the rate, depend on dt through the
            # extrapolated start of the solve)
            leg("n65536_dt1.25e-4", lambda: leg_sharded_water(ctx, api, 65536, 20, compare_single=False, peak=peak, dt=1.25e-4))
        leg("helium_n16384", lambda: leg_helium(ctx, api, peak))
        leg("ensemble_1024xN512", lambda: leg_ensemble(ctx, api, peak))
        if world == 1:
            leg("dense_mode_n4096", lambda: leg_dense_mode(ctx, api))
            leg("hbm_kernels", lambda: leg_hbm_kernels(ctx, api, hbm_peak()))
            if ref_gpu:
                line["reference_cuda_note"] = ("reference_cuda_* = the reference's own CUDA path (BaseBoundaryIntegralCalculator + "
                                               "AutonomousRungeKuttaStepper compiled unmodified from its sources, oracle/build_ref.py) "
                                               "on this same GPU, same surface and dt, host clock around the steps after warm-up")
    done.set()
    if rank == 0:
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(N)
        ctx.emit(line)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configs (N=4096 / 16384, helium, ensemble, dense mode, HBM kernels)")
    ap.add_argument("--no-verify", action="store_true", help="skip the single-GPU rerun the sharded state is compared with")
    ap.add_argument("--extra-timeout", type=float, default=240.0, help="seconds after which the headline is printed without the extras")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip timing the compiled reference CUDA path (oracle/_ref)")
    args = ap.parse_args()
    # the stepper fills its 4-step stage history and tunes the number of recorded sweeps during the first steps (first graph: 16 sweeps
    # per solve; after 8 steps that needed far fewer: needed + 1; after 8 more that all needed the same: exactly that many), each
    # change being a re-capture of the step's CUDA graph: warm up past that (~22 steps) so that no capture falls into the timed region
    args.warmup = max(args.warmup, 30) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
