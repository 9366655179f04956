"""Row-sharded multi-GPU check, launched with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_multirank.py
Every rank also runs the single-GPU solver on its own device; the sharded state must agree with it to round-off."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import roberts_oracle as ro  # noqa: E402
from superfluid_dynamics_b200 import api  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    ok = True
    for N, h, dt, steps in ((2048, 0.4, 1e-3, 20), (4096, 0.4, 1e-3, 20), (8192, 0.3, 2.5e-4, 10), (16384, 0.3, 2e-4, 10), (65536, 0.4, 1e-4, 5)):
        props = api.ProblemProperties(rho=0.0)
        Z, Phi = ro.trochoid(N, h)
        y0 = ro.pack_state(Z, Phi)
        single = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
        s1 = api.AutonomousRungeKuttaStepper(single, dt)
        st1 = torch.as_tensor(y0, device=dev)
        s1.initialize(st1, True)
        s1.runSteps(3)
        torch.cuda.synchronize()
        t0 = time.time()
        s1.runSteps(steps)
        torch.cuda.synchronize()
        t_single = (time.time() - t0) / steps
        sharded = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
        sharded.initComm(rank, world)
        s2 = api.AutonomousRungeKuttaStepper(sharded, dt)
        st2 = torch.as_tensor(y0, device=dev)
        s2.initialize(st2, True)
        s2.runSteps(3)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.time()
        s2.runSteps(steps)
        torch.cuda.synchronize()
        dist.barrier()
        t_sh = (time.time() - t0) / steps
        a, b = st1.cpu().numpy(), st2.cpu().numpy()
        err = np.abs(a - b).max() / np.abs(a).max()
        # all ranks must hold the same replica bit for bit
        ref = st2.clone()
        dist.broadcast(ref, 0)
        same = bool((torch.view_as_real(ref) == torch.view_as_real(st2)).all().item())
        good = err <= 1e-12 and same and sharded.commError() == 0
        ok = ok and good
        print(f"[rank {rank}] N={N}: sharded vs single rel err {err:.2e}, replicas identical {same}, "
              f"{1.0 / t_single:.1f} -> {1.0 / t_sh:.1f} steps/s ({t_single / t_sh:.2f}x on {world} GPUs) "
              f"{sharded.solve_stats()} {s2.stats()} {'OK' if good else 'FAIL'}", flush=True)
        del s1, s2, single, sharded
    # BASELINE config 4: helium film with van-der-Waals forcing, finite depth d = 0.0942478, N = 16384, row-sharded (GMRES solve)
    N, depth, dt, steps = 16384, 0.0942478, 1e-3, 3
    props = api.ProblemProperties(rho=0.0, depth=depth)
    al = 2 * np.pi * np.arange(N) / N
    y0 = ro.pack_state(al + 1j * 0.1 * depth * np.cos(al), np.zeros(N))
    res = []
    for shard in (False, True):
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), device=dev, guess="warm")
        if shard:
            calc.initComm(rank, world)
        stp = api.AutonomousRungeKuttaStepper(calc, dt)
        st = torch.as_tensor(y0, device=dev)
        stp.initialize(st, True)
        stp.runSteps(1)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.time()
        stp.runSteps(steps)
        torch.cuda.synchronize()
        dist.barrier()
        res.append((st.cpu().numpy(), (time.time() - t0) / steps, calc.solve_stats()))
        del stp, calc
    err = np.abs(res[0][0] - res[1][0]).max() / np.abs(res[0][0]).max()
    good = err <= 1e-10 and res[1][2]["converged"]
    ok = ok and good
    print(f"[rank {rank}] helium N={N} d={depth}: sharded vs single rel err {err:.2e}, {1.0 / res[0][1]:.2f} -> {1.0 / res[1][1]:.2f} "
          f"steps/s on {world} GPUs, {res[1][2]} {'OK' if good else 'FAIL'}", flush=True)
    # BASELINE config 5, second half: the 1024-member ensemble at N = 512 spread over the ranks (replicas only: every rank steps its
    # own members, nothing is exchanged); member m: trochoid h_m = 0.05 + 0.35 m / 1023.  Rank 0's first member is checked against
    # the same member stepped alone.
    N, B, dt, steps = 512, 1024, 1e-3, 20
    lo, hi = api.ensemble_member_range(B, rank, world)
    hs = 0.05 + 0.35 * np.arange(B) / (B - 1)
    members = [ro.pack_state(*ro.trochoid(N, h)) for h in hs[lo:hi]]
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, hi - lo, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    st = torch.as_tensor(api.ensemble_state(members, N), device=dev)
    stp.initialize(st, True)
    stp.runSteps(5)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    stp.runSteps(steps)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    alone = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), device=dev, guess="warm")
    s1 = api.AutonomousRungeKuttaStepper(alone, dt)
    s1.initialize(members[0], False)
    s1.runSteps(5 + steps)
    got = st.cpu().numpy()
    mine = np.concatenate([got[:N], got[(hi - lo) * N:(hi - lo) * N + N]])
    err = np.abs(mine - s1.getState()).max() / np.abs(mine).max()
    good = err <= 1e-12 and bool(torch.isfinite(torch.view_as_real(st)).all().item())
    ok = ok and good
    print(f"[rank {rank}] ensemble {B} x N={N}: members [{lo}, {hi}), {B * steps / (t.item() * 1e-3):.0f} member-steps/s on {world} GPUs "
          f"(max over ranks), first member vs stepped alone {err:.2e} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
