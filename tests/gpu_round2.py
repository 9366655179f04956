"""Round-2 measurements on one B200 (diagnostics: prints, never asserts).  Run under gpurun, sections by name:
    python tests/gpu_round2.py shardtune n4096 ... > gpurun_out/round2.log 2>&1
"""
import json
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
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
RESULTS = {}


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def water(N, B=1, **kw):
    props = api.ProblemProperties(rho=0.0)
    return api.BaseBoundaryIntegralCalculator(N, B, props, api.WaterBoundaryProblem(props), device=dev, **kw)


def with_env(cfg, fn):
    os.environ.update({k: str(v) for k, v in cfg.items()})
    try:
        return fn()
    finally:
        for k in cfg:
            os.environ.pop(k, None)


def shardtune():
    """Per-rank sweep of a row-sharded run, timed on one GPU (rb_debug_set_row_range): source chunks vs time."""
    peak = api.measure_fp64_peak(dev)
    rows = []
    nc = lambda *v: [{"RB_NCHUNKS": x} if x else {} for x in v]
    plans = (
        (65536, {1: nc(0, 37, 64, 128), 2: nc(0, 37, 74, 128, 256), 4: nc(0, 74, 148, 256, 512),
                 8: nc(0, 74, 148, 256, 512) + [{"RB_NCHUNKS": 148, "RB_CHUNK_GROUP": 1}, {"RB_NCHUNKS": 512, "RB_CHUNK_GROUP": 1},
                                                {"RB_NCHUNKS": 256, "RB_CHUNK_GROUP": 8}, {"RB_NCHUNKS": 256, "RB_CHUNK_GROUP": 32}]}),
        (16384, {1: nc(0, 32, 64, 128) + [{"RB_V1_ROWS": 4, "RB_NCHUNKS": 64}, {"RB_V1_ROWS": 4, "RB_NCHUNKS": 128}],
                 2: nc(0, 32, 64, 128), 4: nc(0, 64, 128, 256) + [{"RB_V1_ROWS": 4, "RB_NCHUNKS": 128}],
                 8: nc(0, 64, 128, 256) + [{"RB_V1_ROWS": 4, "RB_NCHUNKS": 128}, {"RB_V1_ROWS": 4, "RB_NCHUNKS": 256}]}),
        (4096, {1: nc(0, 16, 32, 64) + [{"RB_NCHUNKS": 64, "RB_CHUNK_GROUP": 1}, {"RB_V1_ROWS": 4, "RB_NCHUNKS": 64}, {"RB_SWEEP_V2": 1}],
                2: nc(0, 64), 4: nc(0, 64)}),
    )
    for N, cands in plans:
        st = T(ro.pack_state(*ro.trochoid(N, 0.4)))
        ncell = N // 256
        for G, cfgs in cands.items():
            for cfg in cfgs:
                def run():
                    c = water(N)
                    if G > 1:
                        c.debugSetRowRange((G // 2) * (ncell // G), ncell // G)
                    ms, pairs = c.benchSweep(st, 10 if N >= 65536 else 40)
                    return ms, c.sweepPlan()
                try:
                    ms, plan = with_env(cfg, run)
                    tf = 20.0 * N * (N / G) / (ms * 1e-3) / 1e12
                    rows.append(dict(N=N, G=G, cfg=cfg, us=ms * 1e3, tflops=tf, frac=tf / peak, **plan))
                    print(f"shardtune N={N} G={G} {cfg}: {ms * 1e3:8.1f} us  {tf:5.2f} TF  {tf / peak:.3f}  "
                          f"tile={plan['tile']}x{plan['tiles_per_chunk']} nchunks={plan['nchunks']} ctas={plan['ctas']} rpt={plan['rows_per_thread']} {plan['kernel']}",
                          flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"shardtune N={N} G={G} {cfg}: FAILED {e}", flush=True)
    RESULTS["shardtune"] = dict(peak=peak, rows=rows)


def n4096():
    """Where the N = 4096 step goes: sweep alone (tiled / persistent, chunk candidates), whole-step rate."""
    N = 4096
    st = T(ro.pack_state(*ro.trochoid(N, 0.4)))
    peak = api.measure_fp64_peak(dev)
    for cfg in ({}, {"RB_NCHUNKS": 8}, {"RB_NCHUNKS": 16}, {"RB_NCHUNKS": 32}, {"RB_NCHUNKS": 64}, {"RB_SWEEP_V2": 1},
                {"RB_V1_ROWS": 4, "RB_NCHUNKS": 37}, {"RB_V1_ROWS": 4, "RB_NCHUNKS": 64}):
        try:
            def run():
                c = water(N)
                ms, pairs = c.benchSweep(st, 100)
                return ms, c.sweepPlan()
            ms, plan = with_env(cfg, run)
            tf = 20.0 * N * N / (ms * 1e-3) / 1e12
            print(f"n4096 sweep {cfg}: {ms * 1e3:.1f} us {tf:.2f} TF {tf / peak:.3f} {plan}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"n4096 sweep {cfg}: FAILED {e}", flush=True)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for cfg in ({}, {"RB_NCHUNKS": 16}, {"RB_NCHUNKS": 64}):
        def run():
            c = water(N, guess="warm")
            stp = api.AutonomousRungeKuttaStepper(c, 1e-3)
            y = T(ro.pack_state(*ro.trochoid(N, 0.4)))
            stp.initialize(y, True)
            stp.runSteps(20)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            stp.runSteps(300)
            b.record()
            torch.cuda.synchronize()
            return 300 / (a.elapsed_time(b) * 1e-3), stp.stats(), c.solve_stats()
        try:
            rate, ss, cs = with_env(cfg, run)
            print(f"n4096 steps {cfg}: {rate:.1f} steps/s {ss} {cs}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"n4096 steps {cfg}: FAILED {e}", flush=True)


def v3():
    """The warp-per-row-group sweep (pair_kernels3.cu) against the tiled and the persistent kernel: solver sweep alone, several N."""
    peak = api.measure_fp64_peak(dev)
    sizes = [int(v) for v in os.environ.get("V3_SIZES", "512,1024,2048,4096,8192").split(",")]
    for N in sizes:
        st = T(ro.pack_state(*ro.trochoid(N, 0.4)))
        cfgs = [{"RB_SWEEP_V3": 1}] + [{"RB_SWEEP_V3": 1, "RB_V3_R": r, "RB_V3_S": sp} for r in (1, 2, 4) for sp in (1, 2, 4)]
        cfgs += [{"RB_SWEEP_V3": 0, "RB_SWEEP_V2": 0}, {"RB_SWEEP_V3": 0, "RB_SWEEP_V2": 1}]
        for cfg in cfgs:
            try:
                def run():
                    c = water(N)
                    ms, pairs = c.benchSweep(st, 200)
                    return ms, c.sweepPlan()
                ms, plan = with_env(cfg, run)
                tf = 20.0 * N * N / (ms * 1e-3) / 1e12
                print(f"v3 sweep N={N} {cfg}: {ms * 1e3:.1f} us {tf / peak:.3f} of peak {plan['kernel']} R={plan['rows_per_thread']} ctas={plan['ctas']} thr={plan['threads']}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"v3 sweep N={N} {cfg}: not available ({str(e)[:60]})", flush=True)


def steps_small():
    """A few recorded RK4 steps at one small size (STEPS_N, default 4096), for ncu launch lists of the launch-bound regime:
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c 160 --csv --log-file ... python tests/gpu_round2.py steps_small"""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    N = int(os.environ.get("STEPS_N", "4096"))
    c = water(N, guess="warm")
    stp = api.AutonomousRungeKuttaStepper(c, 1e-3)
    y = T(ro.pack_state(*ro.trochoid(N, 0.4)))
    stp.initialize(y, True)
    stp.runSteps(int(os.environ.get("STEPS_K", "24")))
    torch.cuda.synchronize()
    print(f"steps_small N={N}: {stp.stats()} launches so far {api.launch_count() if hasattr(api, 'launch_count') else ''}", flush=True)


def kernel_times():
    """Live per-kernel durations of recorded RK4 steps (torch.profiler / CUPTI activity records, not ncu replays) at STEPS_N."""
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    N = int(os.environ.get("STEPS_N", "4096"))
    c = water(N, guess="warm")
    stp = api.AutonomousRungeKuttaStepper(c, float(os.environ.get("STEPS_DT", "1e-3")))
    y = T(ro.pack_state(*ro.trochoid(N, 0.4)))
    stp.initialize(y, True)
    stp.runSteps(60)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        stp.runSteps(40)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    agg = {}
    for e in ev:
        k = e.name.split("(")[0].replace("void ", "").replace("rb::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    tot = sum(v[1] for v in agg.values())
    t0 = min(e.time_range.start for e in ev)
    t1 = max(e.time_range.end for e in ev)
    print(f"kernel_times N={N}: 40 steps, span {(t1 - t0) / 40:.1f} us per step, kernel time {tot / 40:.1f} us per step", flush=True)
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"kernel_times   {n / 40:6.1f} per step  {t / n:8.2f} us avg  {100 * t / tot:5.1f} %  {k}", flush=True)


def steprates():
    """RK4 step rate at the bench sizes with the committed defaults."""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for N, dt, steps in ((1024, 1e-3, 300), (4096, 1e-3, 300), (16384, 1e-4, 60), (65536, 1e-4, 20)):
        c = water(N, guess="warm")
        stp = api.AutonomousRungeKuttaStepper(c, dt)
        y = T(ro.pack_state(*ro.trochoid(N, 0.4)))
        stp.initialize(y, True)
        stp.runSteps(14)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        stp.runSteps(steps)
        b.record()
        torch.cuda.synchronize()
        print(f"steprate N={N}: {steps / (a.elapsed_time(b) * 1e-3):.1f} steps/s {stp.stats()} {c.solve_stats()} {c.sweepPlan()}", flush=True)


def fftrates():
    """RK4 step rate at 1024 < N <= 8192 with the fused radix-8 transforms (default) and with the library transforms (RB_OWN_FFT=0)."""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    for N, dt, steps in ((512, 1e-3, 100), (1024, 1e-3, 100), (2048, 1e-3, 100), (4096, 1e-3, 100), (8192, 1e-3, 60), (8192, 2.5e-4, 60)):
        for own in ("1", "0"):
            os.environ["RB_OWN_FFT"] = own
            try:
                c = water(N, guess="warm")
            finally:
                os.environ.pop("RB_OWN_FFT", None)
            stp = api.AutonomousRungeKuttaStepper(c, dt)
            y = T(ro.pack_state(*ro.trochoid(N, 0.4)))
            stp.initialize(y, True)
            try:
                stp.runSteps(40)
                torch.cuda.synchronize()
                best = 0.0
                for _ in range(3):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    stp.runSteps(steps)
                    b.record()
                    torch.cuda.synchronize()
                    best = max(best, steps / (a.elapsed_time(b) * 1e-3))
                print(f"fftrate N={N} own_fft={own}: {best:.1f} steps/s {stp.stats()}", flush=True)
            except Exception as e:   # noqa: BLE001
                print(f"fftrate N={N} own_fft={own}: FAILED {str(e)[:160]} {stp.stats()}", flush=True)


def helium():
    """Helium film (finite depth, image term): recorded steps with the device-driven GMRES cycle against the host-driven solver."""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    depth = 0.0942478
    for N, steps in ((1024, 100), (4096, 60), (16384, 20)):
        al = 2 * np.pi * np.arange(N) / N
        y0 = ro.pack_state(al + 1j * 0.1 * depth * np.cos(al), np.zeros(N))
        finals = {}
        for dg in (0, 1):
            def run():
                props = api.ProblemProperties(rho=0.0, depth=depth)
                c = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), device=dev, guess="warm")
                stp = api.AutonomousRungeKuttaStepper(c, 1e-3)
                y = T(y0)
                stp.initialize(y, True)
                stp.runSteps(12)
                torch.cuda.synchronize()
                it0 = c.solve_stats()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                stp.runSteps(steps)
                b.record()
                torch.cuda.synchronize()
                it1 = c.solve_stats()
                ms, pairs = c.benchSweep(T(y0), 20)
                return (steps / (a.elapsed_time(b) * 1e-3), (it1["total_iterations"] - it0["total_iterations"]) / steps, stp.stats(), it1,
                        y.cpu().numpy(), ms)
            try:
                rate, spp, ss, cs, yf, ms = with_env({"RB_DEVICE_GMRES": dg}, run)
                finals[dg] = yf
                print(f"helium N={N} device_gmres={dg}: {rate:.1f} steps/s, {spp:.2f} sweeps/step, sweep {ms * 1e3:.1f} us, {ss} {cs}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"helium N={N} device_gmres={dg}: FAILED {e}", flush=True)
        if len(finals) == 2:
            d = np.abs(finals[0] - finals[1]).max() / np.abs(finals[0]).max()
            print(f"helium N={N}: device-driven vs host-driven final state rel diff {d:.2e}", flush=True)


def ensemble():
    """BASELINE config 5b: 1024 members x N = 512 in one batched solver; 4 rows per thread (default) against round 1's 2."""
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    peak = api.measure_fp64_peak(dev)
    for N, B in ((512, 1024), (256, 2048), (768, 512), (512, 128), (300, 37)):
        hs = 0.05 + 0.35 * np.arange(B) / (B - 1)
        sts = [ro.pack_state(*ro.trochoid(N, h)) for h in hs]
        y0 = api.ensemble_state(sts, N)
        finals = {}
        for r4 in (3, 1):   # 3: the warp-per-row-group mapping (sweep3b_kernel), 1: persistent kernel with 4 rows per thread
            def run():
                c = water(N, B, guess="warm")
                stp = api.AutonomousRungeKuttaStepper(c, 1e-3)
                y = T(y0)
                stp.initialize(y, True)
                stp.runSteps(12)
                torch.cuda.synchronize()
                it0 = c.solve_stats()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                stp.runSteps(30)
                b.record()
                torch.cuda.synchronize()
                it1 = c.solve_stats()
                ms, pairs = c.benchSweep(T(y0), 20)
                return 30 / (a.elapsed_time(b) * 1e-3), (it1["total_iterations"] - it0["total_iterations"]) / 30, ms, pairs, c.sweepPlan(), y.cpu().numpy(), it1
            try:
                rate, spp, ms, pairs, plan, yf, st = with_env({"RB_SWEEP_V3B": 1 if r4 == 3 else 0, "RB_SWEEP_V3": -1}, run)
                finals[r4] = yf
                tf = 20.0 * pairs / (ms * 1e-3) / 1e12
                print(f"ensemble {B} x N={N} R4={r4}: {rate:.1f} steps/s ({B * rate:.0f} member-steps/s), {spp:.2f} sweeps/step, sweep {ms * 1e3:.1f} us = "
                      f"{tf:.2f} TF = {tf / peak:.3f}; step {20.0 * pairs * spp * rate / 1e12 / peak:.3f} of peak; {plan} failed={st['failed_solves']}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"ensemble {B} x N={N} R4={r4}: FAILED {e}", flush=True)
        if len(finals) == 2:
            print(f"ensemble {B} x N={N}: warp-rows vs persistent final state rel diff {np.abs(finals[3] - finals[1]).max() / np.abs(finals[1]).max():.2e}", flush=True)


SECTIONS = dict(shardtune=shardtune, n4096=n4096, steprates=steprates, steps_small=steps_small, kernel_times=kernel_times, fftrates=fftrates, v3=v3, helium=helium, ensemble=ensemble)

if __name__ == "__main__":
    names = sys.argv[1:] or list(SECTIONS)
    for n in names:
        t0 = time.time()
        try:
            SECTIONS[n]()
        except Exception:  # noqa: BLE001
            print(f"[{n}] EXCEPTION")
            traceback.print_exc()
        print(f"[{n}] done in {time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(OUT, "round2_results.json"), "a") as f:
        f.write(json.dumps(RESULTS) + "\n")
