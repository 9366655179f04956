"""Shared by tests/test_gpu_reference.py and tests/gpu_reference_report.py: the cases on which the CUDA path (through the C ABI) is
compared with the REFERENCE ITSELF -- the reference's own CUDA classes (BaseBoundaryIntegralCalculator, Water/Helium boundary
problems, AutonomousRungeKuttaStepper) compiled from /root/reference into oracle/_ref/libcusuperhelium_ref.so
(oracle/build_ref.py) and run on the same GPU in a child process (oracle/ref_runner.py).

`reference_results()` runs every reference job once (one child process, one CUDA context); `measure(api, case, ref)` runs the
same input through this repo's library and returns the relative differences.  Tolerances live in the test file."""
import numpy as np

from oracle import ref_runner
from oracle import roberts_oracle as ro


def _surface(N, h, kind="trochoid", depth=1.0):
    a = 2.0 * np.pi * np.arange(N) / N
    if kind == "trochoid":
        Z, Phi = ro.trochoid(N, h)
    elif kind == "multimode":     # no symmetry: several modes with phases
        Z = (a + 0.05 * np.sin(2 * a + 0.3) - h * np.sin(a)) + 1j * (h * np.cos(a) + 0.04 * np.sin(3 * a + 1.0) + 0.02 * np.cos(5 * a))
        Phi = h * np.sin(a) + 0.03 * np.cos(2 * a + 0.7) + 0.01 * np.sin(7 * a)
    elif kind == "film":          # thin film of mean depth `depth`, amplitude h * depth
        Z = (a - 0.3 * h * depth * np.sin(a)) + 1j * (h * depth * np.cos(a))
        Phi = 0.2 * h * depth * np.sin(a)
    else:
        raise ValueError(kind)
    return ro.pack_state(Z, np.asarray(Phi, np.float64))


def _case(name, op, physics, N, h, surface="trochoid", **kw):
    props = dict(rho=kw.pop("rho", 0.0), kappa=kw.pop("kappa", 0.0), depth=kw.pop("depth", 1.0), U=0.0,
                 use_expansions=kw.pop("use_expansions", False), expansion_order=kw.pop("expansion_order", 1),
                 infinite_depth=kw.pop("infinite_depth", False))
    c = dict(name=name, op=op, physics=physics, N=N, h=h, surface=surface, props=props)
    c.update(kw)
    return c


# BASELINE.json configs 1-3 live here as (water, N = 64 / 1024 / 4096); the helium film of config 4 at the sizes the reference
# was instantiated for.  The reference fixes N at compile time (template parameter) and supports batchSize = 1 on this path.
CASES = [
    _case("rhs_water_N64", "rhs", "water", 64, 0.3),
    _case("rhs_water_N256", "rhs", "water", 256, 0.4),
    _case("rhs_water_N256_rho0.2", "rhs", "water", 256, 0.3, rho=0.2),
    _case("rhs_water_N256_multimode", "rhs", "water", 256, 0.2, surface="multimode"),
    _case("rhs_water_N1024", "rhs", "water", 1024, 0.4),
    _case("rhs_water_N4096", "rhs", "water", 4096, 0.4),
    _case("rhs_water_N16384", "rhs", "water", 16384, 0.4),      # the largest size the reference's harness instantiates
    _case("rhs_helium_inf_N256", "rhs", "helium_inf", 256, 0.05, depth=0.3),
    _case("rhs_helium_N256", "rhs", "helium", 256, 0.01, depth=0.3),
    _case("rhs_helium_N256_kappa", "rhs", "helium", 256, 0.01, depth=0.3, kappa=0.01),
    _case("rhs_helium_N256_flag_infinite", "rhs", "helium", 256, 0.01, depth=0.3, infinite_depth=True),
    _case("rhs_helium_N256_expansion2", "rhs", "helium", 256, 0.01, depth=0.3, use_expansions=True, expansion_order=2),
    _case("rhs_helium_film_N1024", "rhs", "helium", 1024, 0.1, surface="film", depth=0.0942478),
    _case("rk4_water_N64_100", "rk4", "water", 64, 0.1, dt=1e-3, steps=100),
    _case("rk4_water_N256_100", "rk4", "water", 256, 0.3, dt=1e-3, steps=100),
    _case("rk4_water_N1024_100", "rk4", "water", 1024, 0.4, dt=1e-3, steps=100),
    _case("rk4_water_N4096_100", "rk4", "water", 4096, 0.4, dt=1e-3, steps=100),   # BASELINE config 3 / half of the metric, the full 100 steps
    _case("rk4_water_N16384_10", "rk4", "water", 16384, 0.4, dt=1e-4, steps=10),   # dt: RK4 stability limit 2.83 (1-h) / (h N/2), DESIGN.md section 5
    _case("rk4_helium_inf_N256_100", "rk4", "helium_inf", 256, 0.05, depth=0.3, dt=1e-3, steps=100),
    _case("rk4_helium_film_N256_100", "rk4", "helium", 256, 0.1, surface="film", depth=0.0942478, dt=1e-3, steps=100),
]


# the optomechanically driven film: App-like properties (A/kernel.cu:77-82: depth 0.0942478, rho = 1, base
# units 1) and variables chosen so that every term of the augmented RHS is O(1e-2 .. 1e-1): drive strength hbar G / sigma^2 ~ 5e-34
OPTO = dict(detuning=0.5, gamma=2.0, G=3.0, Tau=0.7, max_intensity=1e32, initial_time=0.0, location_x0_mode=3.0,
            sigma_optical_mode=0.8, Beta=2e-33, DampingStrength=0.01)
AUG_CASES = [
    _case("aug_rhs_N64", "aug_rhs", "helium", 64, 0.1, surface="film", depth=0.0942478, rho=1.0),
    _case("aug_rhs_N256", "aug_rhs", "helium", 256, 0.1, surface="film", depth=0.0942478, rho=1.0),
    _case("aug_rhs_N1024", "aug_rhs", "helium", 1024, 0.1, surface="film", depth=0.0942478, rho=1.0),
    _case("aug_rk4_N64_100", "aug_rk4", "helium", 64, 0.1, surface="film", depth=0.0942478, rho=1.0, dt=1e-3, steps=100),
    _case("aug_rk4_N256_100", "aug_rk4", "helium", 256, 0.1, surface="film", depth=0.0942478, rho=1.0, dt=1e-3, steps=100),
]
# the same drive in its explicitly time-dependent form (RK4_Time_Dependent.cuh): state [Z | Phi], delayed intensity inside the stepper
TIMED_CASES = [
    _case("timed_rk4_N64_100", "timed_rk4", "helium", 64, 0.1, surface="film", depth=0.0942478, rho=1.0, dt=1e-3, steps=100, t0=0.0),
    _case("timed_rk4_N256_100_t0", "timed_rk4", "helium", 256, 0.1, surface="film", depth=0.0942478, rho=1.0, dt=1e-3, steps=100, t0=0.25),
    _case("timed_rk4_N1024_20", "timed_rk4", "helium", 1024, 0.1, surface="film", depth=0.0942478, rho=1.0, dt=1e-3, steps=20, t0=0.0),
]
for _c in AUG_CASES + TIMED_CASES:
    _c["opto"] = dict(OPTO)
CASES = CASES + AUG_CASES + TIMED_CASES


def state_of(case):
    y = _surface(case["N"], case["h"], case["surface"], case["props"]["depth"])
    if case["op"].startswith("aug"):
        a = 2.0 * np.pi * np.arange(case["N"]) / case["N"]
        D = 0.02 * np.cos(a - 0.4) + 0.01          # delayed intensity block
        y = np.concatenate([y, D.astype(np.complex128)])
    return y


def reference_results(cases=CASES, timeout=420):
    """{case name: result dict} from ONE child process running the compiled reference."""
    jobs = []
    for c in cases:
        j = dict(op=c["op"], kind=c["physics"], N=c["N"], props=c["props"], state=state_of(c))
        if c["op"] in ("rk4", "aug_rk4"):
            j.update(dt=c["dt"], steps=c["steps"], warmup=0)
        if c["op"] == "timed_rk4":
            j.update(dt=c["dt"], steps=c["steps"], t0=c["t0"])
        if c["op"].startswith("aug") or c["op"] == "timed_rk4":
            j["opto"] = c["opto"]
        jobs.append(j)
    res = ref_runner.run_jobs(jobs, timeout=timeout)
    return {c["name"]: r for c, r in zip(cases, res)}


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _calc(api, case, **kw):
    p = case["props"]
    props = api.ProblemProperties(rho=p["rho"], kappa=p["kappa"], depth=p["depth"], use_expansions=p["use_expansions"],
                                  expansion_order=p["expansion_order"], infinite_depth=p["infinite_depth"])
    prob = {"water": api.WaterBoundaryProblem, "helium": api.HeliumBoundaryProblem,
            "helium_inf": api.HeliumInfiniteDepthBoundaryProblem}[case["physics"]](props)
    return api.BaseBoundaryIntegralCalculator(case["N"], 1, props, prob, **kw)


def measure(api, case, ref, torch):
    """Relative differences (max-norm, relative to the reference's max) between this repo's CUDA path and the reference's."""
    N = case["N"]
    y0 = state_of(case)
    dev = torch.device("cuda:0")
    out = {}
    if case["op"] == "timed_rk4":
        p = case["props"]
        props = api.ProblemProperties(rho=p["rho"], kappa=p["kappa"], depth=p["depth"])
        variables = api.OptomechanicalVariables(**case["opto"])
        integrator = api.TimedBoundaryIntegrator(N, 1, props, api.HeliumWithOptomechanicalDrivingProblem(props, variables), guess="warm")
        stp = api.RungeKuttaStepper(integrator, case["dt"])
        stp.initialize(y0, False)
        n = stp.runEvolution(case["t0"], case["t0"] + (case["steps"] + 0.5) * case["dt"])
        y = stp.getState()
        out["steps_taken"] = int(n)
        out["position"] = rel(y[:N], ref["state"][:N])
        out["potential"] = rel(y[N:], ref["state"][N:])
        out["reference_steps_per_s"] = case["steps"] / ref["seconds"] if ref.get("seconds") else None
        out["converged"] = bool(integrator.solve_stats()["converged"])
        return out
    if case["op"].startswith("aug"):
        p = case["props"]
        props = api.ProblemProperties(rho=p["rho"], kappa=p["kappa"], depth=p["depth"])
        variables = api.OptomechanicalVariables(**case["opto"])
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumDrivenAutonomousProblem(props, variables),
                                                  guess="warm" if case["op"] == "aug_rk4" else "cold")
        integrator = api.AugmentedBoundaryIntegrator(calc, api.DelayedIntensityIntegrator(variables))
        if case["op"] == "aug_rhs":
            rhs = torch.zeros(3 * N, dtype=torch.complex128, device=dev)
            integrator.run(torch.as_tensor(y0, device=dev), rhs)
            o = rhs.cpu().numpy()
            out["velocity"] = rel(o[:N], ref["rhs"][:N])
            out["dphi_dt"] = rel(o[N:2 * N], ref["rhs"][N:2 * N])
            out["dD_dt"] = rel(o[2 * N:], ref["rhs"][2 * N:])
        else:
            stp = api.AugmentedRungeKuttaStepper(integrator, case["dt"])
            stp.initialize(y0, False)
            stp.runSteps(case["steps"])
            y = stp.getState()
            out["position"] = rel(y[:N], ref["state"][:N])
            out["potential"] = rel(y[N:2 * N], ref["state"][N:2 * N])
            out["delayed_intensity"] = rel(y[2 * N:], ref["state"][2 * N:])
            out["reference_steps_per_s"] = case["steps"] / ref["seconds"] if ref.get("seconds") else None
        out["converged"] = bool(calc.solve_stats()["converged"])
        return out
    if case["op"] == "rhs":
        calc = _calc(api, case, compute_energies=True)
        rhs = torch.zeros(2 * N, dtype=torch.complex128, device=dev)
        calc.run(torch.as_tensor(y0, device=dev), rhs)
        o = rhs.cpu().numpy()
        out["velocity"] = rel(o[:N], ref["rhs"][:N])
        out["dphi_dt"] = rel(o[N:], ref["rhs"][N:])
        out["a"] = rel(calc.getDevA().cpu().numpy()[:N], ref["a"])
        out["vel_upper"] = rel(calc.devVelocitiesUpper.cpu().numpy()[:N], ref["vel_upper"])
        out["zp"] = rel(calc.getDevZp().cpu().numpy()[:N], ref["zp"])
        out["zpp"] = rel(calc.getDevZpp().cpu().numpy()[:N], ref["zpp"])
        out["phi_prime"] = rel(calc.devPhiPrime.cpu().numpy()[:N], ref["phi_prime"])
        e = calc.energies()
        for i, k in enumerate(("kinetic", "potential", "surface", "volume_flux")):
            out["energy_" + k] = float(abs(e[k] - ref["energies"][i]) / max(1.0, abs(ref["energies"][i])))
        out["converged"] = bool(calc.solve_stats()["converged"])
    else:
        calc = _calc(api, case, guess="warm", compute_energies=True)
        stp = api.AutonomousRungeKuttaStepper(calc, case["dt"])
        stp.initialize(y0, False)
        stp.runSteps(case["steps"])
        y = stp.getState()
        out["position"] = rel(y[:N], ref["state"][:N])
        out["potential"] = rel(y[N:], ref["state"][N:])
        out["converged"] = bool(calc.solve_stats()["converged"])
        out["reference_steps_per_s"] = case["steps"] / ref["seconds"] if ref.get("seconds") else None
        # drift of energy (kinetic + potential + surface, each evaluated by the oracle's formulas on both final states) and of
        # the volume: the reference's own drift over the same steps is the bar (north_star: "no worse than the reference")
        p = case["props"]
        oprops = ro.ProblemProperties(rho=p["rho"], kappa=p["kappa"], depth=p["depth"], use_expansions=p["use_expansions"],
                                      expansion_order=p["expansion_order"], infinite_depth=p["infinite_depth"])

        def diag(yv):
            v, _, aux = ro.rhs_single(yv[:N], yv[N:].real, oprops, case["physics"], full=True)
            en = ro.energies(yv[:N], aux["Zp"], yv[N:], v, oprops, case["physics"])
            return en["kinetic"] + en["potential"] + en["surface"], ro.volume(yv[:N], aux["Zp"])

        if N <= 1024:
            e0, v0 = diag(y0)
            er, vr = diag(ref["state"])
            em, vm = diag(y)
            out["energy_drift_ours"], out["energy_drift_reference"] = abs(em - e0), abs(er - e0)
            out["volume_drift_ours"], out["volume_drift_reference"] = abs(vm - v0), abs(vr - v0)
            out["energy_scale"] = abs(e0)
    return out
