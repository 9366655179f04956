"""Host-side mirror of the reference's solver / time-stepper interface for the Roberts BIE RK4 path.

Class and method names follow CuSuperHelium (L/ = CuSuperHelium/CuSuperHelium/):
    ProblemProperties                    L/ProblemProperties.hpp:5-32
    WaterBoundaryProblem / HeliumBoundaryProblem / HeliumInfiniteDepthBoundaryProblem
                                         L/WaterBoundaryProblem.cuh, L/HeliumBoundaryProblem.cuh
    BaseBoundaryIntegralCalculator       L/BaseBoundaryIntegrator.cuh:10-85   (run, runTimeStep, calculateVorticities, getDevA..)
    AutonomousRungeKuttaStepper          L/AutonomousRungeKuttaStepper.cuh:24-121 (initialize, runStep, runEvolution, setTimeStep)
    createMKernel, createFiniteDepthMKernel, createVelocityMatrices, createHeliumVelocityMatrices,
    compute_rhs_phi_expression, ...      the kernels the reference's tests launch by name
Device memory is held in torch tensors (complex128 / float64 on a CUDA device); every computation goes through the C ABI of
libroberts_b200.so.  N and the batch size are runtime values here (template parameters in the reference).
"""
from __future__ import annotations

import ctypes
import dataclasses

import numpy as np
import torch

from . import _lib
from ._lib import check

PHYSICS = {"water": _lib.RB_WATER, "helium": _lib.RB_HELIUM, "helium_inf": _lib.RB_HELIUM_INF}


@dataclasses.dataclass
class ProblemProperties:
    """L/ProblemProperties.hpp:5-32 (nondimensional)."""
    L: float = 1.0
    rho: float = 1.0
    U: float = 0.0
    kappa: float = 0.0
    depth: float = 1.0
    initial_amplitude: float = 1.0
    use_expansions: bool = False
    expansion_order: int = 1
    infinite_depth: bool = False


class _BoundaryProblem:
    physics = "water"

    def __init__(self, properties: ProblemProperties):
        self.properties = properties


class WaterBoundaryProblem(_BoundaryProblem):
    physics = "water"


class HeliumBoundaryProblem(_BoundaryProblem):
    physics = "helium"


class HeliumInfiniteDepthBoundaryProblem(_BoundaryProblem):
    physics = "helium_inf"


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors must be contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _DevView:
    """Zero-copy view of library-owned device memory as a torch tensor (float64)."""

    def __init__(self, ptr, n_doubles, device_index):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                         "strides": None}
        self._dev = device_index


def _view(ptr, n_doubles, device, complex_=False):
    t = torch.as_tensor(_DevView(ptr, n_doubles, device.index), device=device)
    return torch.view_as_complex(t.view(-1, 2)) if complex_ else t


class BaseBoundaryIntegralCalculator:
    """RHS assembler: BaseBoundaryIntegralCalculator<N, batchSize>(ProblemProperties&, BoundaryProblem<N,batchSize>&)."""

    def __init__(self, N: int, batchSize: int, problemProperties: ProblemProperties, boundaryProblem: _BoundaryProblem = None,
                 device=None, solve_mode: str = "matrix_free", guess: str = "cold", tolerance: float = 1e-13,
                 max_iterations: int = 200, compute_energies: bool = False):
        lib = _lib.load()
        if lib.rb_device_count() == 0:
            raise _lib.RobertsError("no CUDA device: superfluid_dynamics_b200 has no CPU path")
        self.lib = lib
        self.N, self.batchSize = int(N), int(batchSize)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.properties = problemProperties
        physics = boundaryProblem.physics if boundaryProblem is not None else "water"
        self.physics = physics
        p = _lib.rb_props()
        lib.rb_default_props(ctypes.byref(p))
        p.rho, p.U, p.kappa, p.depth = problemProperties.rho, problemProperties.U, problemProperties.kappa, problemProperties.depth
        p.use_expansions = int(problemProperties.use_expansions)
        p.expansion_order = int(problemProperties.expansion_order)
        p.infinite_depth = int(problemProperties.infinite_depth)
        p.physics = PHYSICS[physics]
        p.solve_mode = _lib.RB_SOLVE_DENSE_LU if solve_mode == "dense_lu" else _lib.RB_SOLVE_MATRIX_FREE
        p.guess_mode = _lib.RB_GUESS_WARM if guess == "warm" else _lib.RB_GUESS_COLD
        p.max_iterations = max_iterations
        p.compute_energies = int(compute_energies)
        p.tolerance = tolerance
        self._props = p
        with torch.cuda.device(self.device):
            check(lib.rb_set_device(self.device.index), "rb_set_device")
            self.handle = lib.rb_create(self.N, self.batchSize, ctypes.byref(p))
        if not self.handle:
            raise _lib.RobertsError("rb_create: " + lib.rb_last_error().decode())
        self.setStream(torch.cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_destroy(h)

    # --- AutonomousProblem<std_complex, 2N*batch> ---
    def setStream(self, stream):
        check(self.lib.rb_set_stream(self.handle, ctypes.c_void_p(int(stream))), "rb_set_stream")

    def run(self, initialState: torch.Tensor, rhs: torch.Tensor):
        check(self.lib.rb_rhs(self.handle, _ptr(initialState), _ptr(rhs)), "rb_rhs")

    runTimeStep = run

    def calculateVorticities(self, initialState: torch.Tensor):
        check(self.lib.rb_vorticities(self.handle, _ptr(initialState)), "rb_vorticities")

    def getDevA(self):
        return _view(self.lib.rb_dev_a(self.handle), self.N * self.batchSize, self.device)

    def getDevZp(self):
        return _view(self.lib.rb_dev_zp(self.handle), 2 * self.N * self.batchSize, self.device, True)

    def getDevZpp(self):
        return _view(self.lib.rb_dev_zpp(self.handle), 2 * self.N * self.batchSize, self.device, True)

    @property
    def devVelocitiesUpper(self):
        return _view(self.lib.rb_dev_velocities_upper(self.handle), 2 * self.N * self.batchSize, self.device, True)

    @property
    def devPhiPrime(self):
        return _view(self.lib.rb_dev_phi_prime(self.handle), self.N * self.batchSize, self.device)

    def synchronize(self):
        check(self.lib.rb_synchronize(self.handle), "rb_synchronize")

    def energies(self):
        out = (ctypes.c_double * 5)()
        check(self.lib.rb_energies(self.handle, out), "rb_energies")
        return dict(kinetic=out[0], potential=out[1], surface=out[2], volume_flux=out[3], volume=out[4])

    def solve_stats(self):
        out = (ctypes.c_double * 6)()
        check(self.lib.rb_solve_stats(self.handle, out), "rb_solve_stats")
        st = (ctypes.c_double * 8)()
        check(self.lib.rb_solve_status(self.handle, st), "rb_solve_status")
        return dict(iterations=int(out[0]), converged=bool(out[1]), residual=out[2], total_iterations=int(out[3]),
                    total_solves=int(out[4]), velocity_sweeps=int(out[5]), stagnated=bool(st[1]), stagnated_solves=int(st[2]),
                    failed_solves=int(st[3]), worst_residual=st[4], strict=bool(st[5]))

    def setStrict(self, strict=True):
        """strict (default): a solve that neither converges nor stagnates on the round-off floor raises; False: statistics only."""
        check(self.lib.rb_set_strict(self.handle, int(bool(strict))), "rb_set_strict")

    # --- derivatives (ZPhiDerivative / FftDerivative) ---
    def zPhiDerivative(self, Z, Phi):
        Zp, PhiP, Zpp = torch.empty_like(Z), torch.empty_like(Z), torch.empty_like(Z)
        check(self.lib.rb_zphi_derivative(self.handle, _ptr(Z), _ptr(Phi), _ptr(Zp), _ptr(PhiP), _ptr(Zpp)), "rb_zphi_derivative")
        return Zp, PhiP, Zpp

    def fftDerivative(self, x, doubleDev=False, scaling=1.0):
        out = torch.empty_like(x)
        check(self.lib.rb_fft_derivative(self.handle, _ptr(x), _ptr(out), int(doubleDev), float(scaling)), "rb_fft_derivative")
        return out

    def cotangentSum(self, Z, x):
        S = torch.empty_like(Z)
        check(self.lib.rb_cotangent_sum(self.handle, _ptr(Z), _ptr(x), _ptr(S)), "rb_cotangent_sum")
        return S

    # --- multi-GPU: row cells sharded over the ranks of a torch.distributed process group (one process per GPU, one node) ---
    def initComm(self, rank=None, world=None, group=None):
        import torch.distributed as dist
        rank = dist.get_rank(group) if rank is None else rank
        world = dist.get_world_size(group) if world is None else world
        nb = self.lib.rb_comm_handle_bytes()
        mine = ctypes.create_string_buffer(nb)
        check(self.lib.rb_comm_export(self.handle, mine), "rb_comm_export")
        blobs = exchange_handles(mine.raw, group)
        assert len(blobs) == world and all(len(b) == nb for b in blobs)
        check(self.lib.rb_comm_init(self.handle, int(rank), int(world), b"".join(blobs)), "rb_comm_init")
        dist.barrier(group)
        self.rank, self.world = rank, world

    def commError(self):
        return self.lib.rb_comm_error(self.handle)

    def benchSweep(self, state, reps=10):
        ms, pairs = ctypes.c_float(), ctypes.c_double()
        check(self.lib.rb_bench_sweep(self.handle, _ptr(state), reps, ctypes.byref(ms), ctypes.byref(pairs)), "rb_bench_sweep")
        return ms.value, pairs.value

    def debugSetRowRange(self, cell0, cells):
        """Measurement aid: sweep only the 256-row cells [cell0, cell0 + cells) -- the per-rank share of a row-sharded run, on one GPU."""
        check(self.lib.rb_debug_set_row_range(self.handle, int(cell0), int(cells)), "rb_debug_set_row_range")

    def sweepPlan(self):
        out = (ctypes.c_int * 8)()
        check(self.lib.rb_sweep_plan(self.handle, out), "rb_sweep_plan")
        return dict(kernel={1: "tiled", 2: "persistent", 3: "warp_rows"}[out[0]], rows_per_thread=out[1], tile=out[2], tiles_per_chunk=out[3],
                    nchunks=out[4], row_cells=out[5], ctas=out[6], threads=out[7])


def exchange_handles(blob: bytes, group=None):
    """All-gather one opaque handle per rank (plumbing only: torch.distributed, any backend)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, blob, group=group)
    return out


def comm_row_range(N, rank, nranks):
    """Rows [lo, hi) of the interaction operators owned by `rank` (whole 256-row cells, contiguous)."""
    out = (ctypes.c_int * 2)()
    if _lib.load().rb_comm_row_range(int(N), int(rank), int(nranks), out) != 0:
        raise ValueError("bad rank / nranks")
    return out[0], out[1]


def ensemble_member_range(members, rank, nranks):
    """Members [lo, hi) of an ensemble owned by `rank`: contiguous, sizes differing by at most one, the larger shares first.
    Ensembles are replicas only -- every rank steps its own members with its own batched solver, nothing is exchanged
    (SURVEY.md section 8e)."""
    members, rank, nranks = int(members), int(rank), int(nranks)
    if nranks < 1 or not 0 <= rank < nranks or members < 0:
        raise ValueError("bad rank / nranks")
    q, r = divmod(members, nranks)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def ensemble_state(member_states, N):
    """Pack per-member states [Z_m | Phi_m] (2N complex each) into the batched layout [Z of every member | Phi of every member]."""
    ms = [np.asarray(m, np.complex128) for m in member_states]
    return np.concatenate([m[:N] for m in ms] + [m[N:2 * N] for m in ms])


class AutonomousRungeKuttaStepper:
    """AutonomousRungeKuttaStepper<std_complex, 2N>(AutonomousProblem&, tstep, logger)."""

    def __init__(self, autonomousProblem: BaseBoundaryIntegralCalculator, tstep: float = 1e-2):
        self.problem = autonomousProblem
        self.lib = autonomousProblem.lib
        self.handle = self.lib.rb_rk4_create(autonomousProblem.handle, float(tstep))
        if not self.handle:
            raise _lib.RobertsError("rb_rk4_create: " + self.lib.rb_last_error().decode())
        self._keep = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_rk4_destroy(h)

    def setTimeStep(self, tstep):
        check(self.lib.rb_rk4_set_time_step(self.handle, float(tstep)), "rb_rk4_set_time_step")

    def setOptions(self, initial_timestep):
        self.setTimeStep(initial_timestep)

    def initialize(self, devY0, onDevice=False):
        if onDevice:
            self._keep = devY0
            check(self.lib.rb_rk4_initialize(self.handle, _ptr(devY0), 1), "rb_rk4_initialize")
        else:
            host = np.ascontiguousarray(np.asarray(devY0, dtype=np.complex128))
            check(self.lib.rb_rk4_initialize(self.handle, host.ctypes.data_as(ctypes.c_void_p), 0), "rb_rk4_initialize")

    def runStep(self, _step=0):
        check(self.lib.rb_rk4_step(self.handle), "rb_rk4_step")

    def runSteps(self, steps):
        check(self.lib.rb_rk4_run_steps(self.handle, int(steps)), "rb_rk4_run_steps")

    def runEvolution(self, startTime, endTime):
        n = ctypes.c_size_t()
        check(self.lib.rb_rk4_evolve(self.handle, float(startTime), float(endTime), ctypes.byref(n)), "rb_rk4_evolve")
        return n.value

    def getState(self):
        n = 2 * self.problem.N * self.problem.batchSize
        host = np.empty(n, np.complex128)
        check(self.lib.rb_rk4_get_state(self.handle, host.ctypes.data_as(ctypes.c_void_p)), "rb_rk4_get_state")
        return host

    def currentTime(self):
        return self.lib.rb_rk4_current_time(self.handle)

    def stats(self):
        out = (ctypes.c_double * 4)()
        check(self.lib.rb_rk4_stats(self.handle, out), "rb_rk4_stats")
        ch = (ctypes.c_double * 4)()
        check(self.lib.rb_rk4_chunk_stats(self.handle, ch), "rb_rk4_chunk_stats")
        return dict(graph_launches=int(out[0]), graph_captures=int(out[1]), fallback_steps=int(out[2]), graph_sweeps=int(out[3]),
                    chunk_steps=int(ch[0]), chunks=int(ch[1]), chunks_rolled_back=int(ch[2]),
                    tight=bool(ch[3] - int(ch[3]) > 0.25), tight_failures=int(ch[3]))

    def guessStats(self):
        out = (ctypes.c_double * 8)()
        check(self.lib.rb_rk4_guess_stats(self.handle, out), "rb_rk4_guess_stats")
        return dict(first_rel=[out[i] for i in range(4)], opt_mask=int(out[4]), optimistic_solves=int(out[5]),
                    one_sweep_solves=int(out[6]), policy=int(out[7]))

    def setOptimistic(self, policy):
        check(self.lib.rb_rk4_set_optimistic(self.handle, int(policy)), "rb_rk4_set_optimistic")

    def setGuess(self, order, predict=-1):
        check(self.lib.rb_rk4_set_guess(self.handle, int(order), int(predict)), "rb_rk4_set_guess")

    def setLogging(self, every, capacity):
        check(self.lib.rb_rk4_set_logging(self.handle, int(every), int(capacity)), "rb_rk4_set_logging")

    def copyTrajectory(self):
        tp, tc = ctypes.POINTER(ctypes.c_double)(), ctypes.c_size_t()
        sp, sc = ctypes.c_void_p(), ctypes.c_size_t()
        check(self.lib.rb_rk4_copy_trajectory(self.handle, ctypes.byref(tp), ctypes.byref(tc), ctypes.byref(sp), ctypes.byref(sc)),
              "rb_rk4_copy_trajectory")
        n = 2 * self.problem.N * self.problem.batchSize
        times = np.ctypeslib.as_array(tp, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
        if sc.value:
            buf = (ctypes.c_double * (2 * n * sc.value)).from_address(sp.value)
            states = np.frombuffer(buf, dtype=np.complex128).reshape(sc.value, n).copy()
        else:
            states = np.zeros((0, n), np.complex128)
        self.lib.rb_free(ctypes.cast(tp, ctypes.c_void_p))
        self.lib.rb_free(sp)
        return times, states


@dataclasses.dataclass
class OptomechanicalVariables:
    """L/OptomechanicalVariables.h:3-28 (nondimensional).  `drive_strength` is
    LightIntensity::get_current_intensity_drive_strength(variables, properties) (L/LightIntensity.cuh:30-33): pass it, or the base
    units through `set_drive_strength` (the reference's default base units are 1)."""
    detuning: float = 0.0
    gamma: float = 1.0
    G: float = 1.0
    Tau: float = 1.0
    max_intensity: float = 0.0
    initial_time: float = 0.0
    location_x0_mode: float = 0.0
    sigma_optical_mode: float = 1.0
    Beta: float = 0.0
    DampingStrength: float = 0.01
    drive_strength: float = None

    def set_drive_strength(self, rho, base_energy=1.0, base_time=1.0):
        self.drive_strength = _lib.load().rb_opto_drive_strength(ctypes.byref(self._c(0.0)), float(base_energy), float(base_time),
                                                                 float(rho))
        return self

    def _c(self, strength=None):
        v = _lib.rb_opto()
        for f in ("detuning", "gamma", "G", "Tau", "max_intensity", "initial_time", "location_x0_mode", "sigma_optical_mode", "Beta",
                  "DampingStrength"):
            setattr(v, f, float(getattr(self, f)))
        v.drive_strength = float(self.drive_strength if strength is None else strength)
        return v


class HeliumDrivenAutonomousProblem(HeliumBoundaryProblem):
    """HeliumDrivenAutonomousProblem<N,B>(ProblemProperties&, OptomechanicalVariables&), L/HeliumDrivenAutonomousProblem.cuh:10-26."""

    def __init__(self, properties: ProblemProperties, variables: OptomechanicalVariables):
        super().__init__(properties)
        self.variables = variables


class DelayedIntensityIntegrator:
    """DelayedIntensityIntegrator<N,B>(OptomechanicalVariables&), L/DelayedIntensityIntegrator.cuh:9-39."""

    def __init__(self, variables: OptomechanicalVariables):
        self.variables = variables


class AugmentedBoundaryIntegrator:
    """AugmentedBoundaryIntegrator<N,B>(integrator, delayedIntegrator), L/AugmentedBoundaryIntegrator.cuh:10-40: the autonomous
    system y = [Z | Phi | D] (3 N B complex).  The boundary-integral calculator is the one built on a
    HeliumDrivenAutonomousProblem; its variables and the delayed integrator's must be the same object's values."""

    def __init__(self, integrator: BaseBoundaryIntegralCalculator, delayedIntegrator: DelayedIntensityIntegrator = None,
                 variables: OptomechanicalVariables = None):
        self.integrator = integrator
        self.lib = integrator.lib
        self.variables = variables or (delayedIntegrator.variables if delayedIntegrator is not None else None)
        if self.variables is None:
            raise ValueError("AugmentedBoundaryIntegrator needs the OptomechanicalVariables")
        if self.variables.drive_strength is None:
            self.variables.set_drive_strength(integrator.properties.rho)
        self.N, self.batchSize, self.device = integrator.N, integrator.batchSize, integrator.device

    def run(self, initialState: torch.Tensor, rhs: torch.Tensor):
        v = self.variables._c()
        check(self.lib.rb_augmented_rhs(self.integrator.handle, ctypes.byref(v), _ptr(initialState), _ptr(rhs)), "rb_augmented_rhs")

    def setStream(self, stream):
        self.integrator.setStream(stream)

    def lightIntensity(self, Z: torch.Tensor):
        out = torch.empty(Z.numel(), dtype=torch.float64, device=Z.device)
        v = self.variables._c()
        check(self.lib.rb_light_intensity(_ptr(Z), _ptr(out), ctypes.byref(v), Z.numel(), _stream_ptr(Z.device)), "rb_light_intensity")
        return out


class HeliumWithOptomechanicalDrivingProblem(HeliumBoundaryProblem):
    """HeliumWithOptomechanicalDrivingProblem<N>(ProblemProperties&, OptomechanicalVariables), L/HeliumWithDrivingBoundaryProblem.cuh:7-67:
    the explicitly time-dependent drive (delayed intensity advanced by an exponential integrator, L/DelayedIntensityTerm.cuh)."""

    def __init__(self, properties: ProblemProperties, variables: OptomechanicalVariables):
        super().__init__(properties)
        self.variables = variables


class TimedBoundaryIntegrator(BaseBoundaryIntegralCalculator):
    """TimedBoundaryIntegrator<N,B>(ProblemProperties&, TimedBoundaryProblem&), L/TimedBoundaryIntegrator.cuh:8-49."""

    def __init__(self, N, batchSize, problemProperties, boundaryProblem: HeliumWithOptomechanicalDrivingProblem, **kw):
        super().__init__(N, batchSize, problemProperties, boundaryProblem, **kw)
        self.variables = boundaryProblem.variables
        if self.variables.drive_strength is None:
            self.variables.set_drive_strength(problemProperties.rho)


class RungeKuttaStepper:
    """RungeKuttaStepper<std_complex, 2N>(TimedProblem&, tstep), L/RK4_Time_Dependent.cuh:18-460.  As in the reference runStep()
    does not advance the time (runEvolution's loop does); runStep(advance=True) adds that `currentTime += timeStep`."""

    def __init__(self, timedProblem: TimedBoundaryIntegrator, tstep: float = 1e-2):
        self.problem = timedProblem
        self.lib = timedProblem.lib
        v = timedProblem.variables._c()
        self.handle = self.lib.rb_timed_rk4_create(timedProblem.handle, ctypes.byref(v), float(tstep))
        if not self.handle:
            raise _lib.RobertsError("rb_timed_rk4_create: " + self.lib.rb_last_error().decode())
        self._keep = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_timed_rk4_destroy(h)

    def setTimeStep(self, tstep):
        check(self.lib.rb_timed_rk4_set_time_step(self.handle, float(tstep)), "rb_timed_rk4_set_time_step")

    def setStartingTime(self, time):
        check(self.lib.rb_timed_rk4_set_starting_time(self.handle, float(time)), "rb_timed_rk4_set_starting_time")

    def initialize(self, devY0, onDevice=False):
        if onDevice:
            self._keep = devY0
            check(self.lib.rb_timed_rk4_initialize(self.handle, _ptr(devY0), 1), "rb_timed_rk4_initialize")
        else:
            host = np.ascontiguousarray(np.asarray(devY0, dtype=np.complex128))
            check(self.lib.rb_timed_rk4_initialize(self.handle, host.ctypes.data_as(ctypes.c_void_p), 0), "rb_timed_rk4_initialize")

    def run(self, time, saveProgress, state: torch.Tensor, rhs: torch.Tensor):
        """setCurrentTime(time); setSaveProgress(saveProgress); run(state, rhs)."""
        check(self.lib.rb_timed_rhs(self.handle, float(time), int(bool(saveProgress)), _ptr(state), _ptr(rhs)), "rb_timed_rhs")

    def runStep(self, _step=0, advance=False):
        check(self.lib.rb_timed_rk4_step(self.handle, int(bool(advance))), "rb_timed_rk4_step")

    def runEvolution(self, startTime, endTime):
        n = ctypes.c_size_t()
        check(self.lib.rb_timed_rk4_evolve(self.handle, float(startTime), float(endTime), ctypes.byref(n)), "rb_timed_rk4_evolve")
        return n.value

    def getState(self):
        host = np.empty(2 * self.problem.N * self.problem.batchSize, np.complex128)
        check(self.lib.rb_timed_rk4_get_state(self.handle, host.ctypes.data_as(ctypes.c_void_p)), "rb_timed_rk4_get_state")
        return host

    def setOptions(self, initial_timestep, returnTrajectory=True):
        """RK4Options{initial_timestep, returnTrajectory} (L/RK4Options.h)."""
        self.setTimeStep(initial_timestep)
        check(self.lib.rb_timed_rk4_set_logging(self.handle, int(bool(returnTrajectory))), "rb_timed_rk4_set_logging")

    def copyTrajectory(self):
        """copyTimesToHost + copyStatesToHost: (times, states [count x 2 N B])."""
        tp, tc = ctypes.POINTER(ctypes.c_double)(), ctypes.c_size_t()
        sp, sc = ctypes.c_void_p(), ctypes.c_size_t()
        check(self.lib.rb_timed_rk4_copy_trajectory(self.handle, ctypes.byref(tp), ctypes.byref(tc), ctypes.byref(sp), ctypes.byref(sc)),
              "rb_timed_rk4_copy_trajectory")
        n = 2 * self.problem.N * self.problem.batchSize
        times = np.ctypeslib.as_array(tp, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
        buf = ctypes.cast(sp, ctypes.POINTER(ctypes.c_double))
        states = np.ctypeslib.as_array(buf, shape=(sc.value, 2 * n)).copy().view(np.complex128) if sc.value else np.zeros((0, n), np.complex128)
        if tc.value:
            self.lib.rb_free(ctypes.cast(tp, ctypes.c_void_p))
        self.lib.rb_free(sp)
        return times, states

    def delayedIntensity(self):
        return _view(self.lib.rb_timed_rk4_dev_delayed_intensity(self.handle), self.problem.N * self.problem.batchSize,
                     self.problem.device)

    def currentTime(self):
        return self.lib.rb_timed_rk4_current_time(self.handle)


class RealBoundaryItegralCalculator:
    """RealBoundaryItegralCalculator<N>(BaseBoundaryIntegralCalculator<N,1>&), L/RealBoundaryIntegralCalculator.cuh:37-89: the RHS on
    real states [x | y | phi] (3N doubles on the device); the spelling of the class name is the reference's."""

    def __init__(self, boundaryIntegralCalculator: BaseBoundaryIntegralCalculator):
        assert boundaryIntegralCalculator.batchSize == 1
        self.calculator = boundaryIntegralCalculator
        self.lib = boundaryIntegralCalculator.lib

    def run(self, initialState: torch.Tensor, rhs: torch.Tensor):
        check(self.lib.rb_real_rhs(self.calculator.handle, _ptr(initialState), _ptr(rhs)), "rb_real_rhs")


def createInitialBatchedZ(initialState: torch.Tensor, ZBatched: torch.Tensor, eps: float, N: int):
    """createInitialBatchedZ<<<(ceil(2N/256), 3N), 256>>>(initialState, ZBatched, eps, N), L/JacobianCalculator.cuh:11-77."""
    check(_lib.load().rb_perturbed_states(_ptr(initialState), _ptr(ZBatched), float(eps), int(N), _stream_ptr(initialState.device)),
          "rb_perturbed_states")


def lu_solve(A: torch.Tensor, b: torch.Tensor, n: int, blocked: int = -1) -> int:
    """MatrixSolver<N,1>::solve (L/MatrixSolver.cuh:114-125): A (n x n column-major, destroyed) x = b in place; returns getrf's info."""
    info = ctypes.c_int()
    check(_lib.load().rb_lu_solve(_ptr(A), _ptr(b), int(n), int(blocked), ctypes.byref(info), _stream_ptr(A.device)), "rb_lu_solve")
    return info.value


class JacobianCalculator:
    """JacobianCalculator<N>(std::make_unique<BaseBoundaryIntegralCalculator<N, 3N>>(properties, problem)), L/JacobianCalculator.cuh:168-284:
    here the calculator builds its batch-3N RHS assembler itself from the properties and the physics plugin."""

    def __init__(self, N: int, problemProperties: ProblemProperties, boundaryProblem: _BoundaryProblem = None, device=None,
                 tolerance: float = 1e-13, max_iterations: int = 200):
        lib = _lib.load()
        if lib.rb_device_count() == 0:
            raise _lib.RobertsError("no CUDA device: superfluid_dynamics_b200 has no CPU path")
        self.lib, self.N = lib, int(N)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        p = _lib.rb_props()
        lib.rb_default_props(ctypes.byref(p))
        p.rho, p.U, p.kappa, p.depth = problemProperties.rho, problemProperties.U, problemProperties.kappa, problemProperties.depth
        p.use_expansions = int(problemProperties.use_expansions)
        p.expansion_order = int(problemProperties.expansion_order)
        p.infinite_depth = int(problemProperties.infinite_depth)
        p.physics = PHYSICS[boundaryProblem.physics if boundaryProblem is not None else "helium"]
        p.max_iterations = max_iterations
        p.tolerance = tolerance
        with torch.cuda.device(self.device):
            check(lib.rb_set_device(self.device.index), "rb_set_device")
            self.handle = lib.rb_jacobian_create(self.N, ctypes.byref(p))
        if not self.handle:
            raise _lib.RobertsError("rb_jacobian_create: " + lib.rb_last_error().decode())
        self.setStream(torch.cuda.current_stream(self.device).cuda_stream)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_jacobian_destroy(h)

    def setEpsilon(self, eps: float):
        check(self.lib.rb_jacobian_set_epsilon(self.handle, float(eps)), "rb_jacobian_set_epsilon")

    def setStream(self, stream):
        check(self.lib.rb_jacobian_set_stream(self.handle, ctypes.c_void_p(int(stream))), "rb_jacobian_set_stream")

    def calculateJacobian(self, devState: torch.Tensor, devJacobian: torch.Tensor):
        """devState: 3N doubles [x | y | phi]; devJacobian: 9 N^2 doubles, column-major (jac[c * 3N + r] = d f_r / d y_c)."""
        check(self.lib.rb_jacobian_calculate(self.handle, _ptr(devState), _ptr(devJacobian)), "rb_jacobian_calculate")

    def solve_stats(self):
        out = (ctypes.c_double * 6)()
        check(self.lib.rb_solve_stats(self.lib.rb_jacobian_solver(self.handle), out), "rb_solve_stats")
        return dict(iterations=int(out[0]), converged=bool(out[1]), relative_residual=out[2])


@dataclasses.dataclass
class GaussLegendre2Options:
    """GaussLegendre2Options, L/GaussLegendre.cuh:70-90."""
    stepSize: float = 0.01
    newtonTolerance: float = 1e-10
    maxNewtonIterations: int = 20
    allowSimplifiedFallback: bool = False
    returnTrajectory: bool = True
    armijo_c: float = 1e-4
    backtrack: float = 0.5
    minAlpha: float = 1e-6
    maxStepsHalves: int = 6

    def _c(self):
        o = _lib.rb_gl2_options()
        for f in dataclasses.fields(self):
            setattr(o, f.name, type(getattr(o, f.name))(getattr(self, f.name)))
        return o


class GaussLegendre2:
    """GaussLegendre2<N>(AutonomousProblem<double, 3N>& problem, JacobianCalculator<N>&, GaussLegendre2Options), L/GaussLegendre.cuh:107-612."""

    def __init__(self, problem: RealBoundaryItegralCalculator, jacobianCalculator: JacobianCalculator,
                 options: GaussLegendre2Options = None):
        self.problem, self.jacobianCalculator = problem, jacobianCalculator
        self.lib = problem.lib
        self.N = problem.calculator.N
        o = (options or GaussLegendre2Options())._c()
        self.handle = self.lib.rb_gl2_create(problem.calculator.handle, jacobianCalculator.handle, ctypes.byref(o))
        if not self.handle:
            raise _lib.RobertsError("rb_gl2_create: " + self.lib.rb_last_error().decode())
        self._keep = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_gl2_destroy(h)

    def setOptions(self, options: GaussLegendre2Options):
        o = options._c()
        check(self.lib.rb_gl2_set_options(self.handle, ctypes.byref(o)), "rb_gl2_set_options")

    def initialize(self, initialState, onDevice=False):
        if onDevice:
            self._keep = initialState
            check(self.lib.rb_gl2_initialize(self.handle, _ptr(initialState), 1), "rb_gl2_initialize")
        else:
            host = np.ascontiguousarray(np.asarray(initialState, dtype=np.float64))
            assert host.size == 3 * self.N
            check(self.lib.rb_gl2_initialize(self.handle, host.ctypes.data_as(ctypes.c_void_p), 0), "rb_gl2_initialize")

    def step(self, h: float) -> bool:
        """gaussLegendreS2Step from the current state; True (and the state advanced) when the Newton iteration converged."""
        ok = ctypes.c_int()
        check(self.lib.rb_gl2_step(self.handle, float(h), ctypes.byref(ok)), "rb_gl2_step")
        return bool(ok.value)

    def runEvolution(self, startTime: float, endTime: float):
        check(self.lib.rb_gl2_evolve(self.handle, float(startTime), float(endTime)), "rb_gl2_evolve")

    def getState(self):
        host = np.empty(3 * self.N)
        check(self.lib.rb_gl2_get_state(self.handle, _dp(host)), "rb_gl2_get_state")
        return host

    def copyTrajectory(self):
        """copyTimesToHost + copyStatesToHost: (times, states [count x 3N])."""
        tp, sp = ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)()
        tc, sc = ctypes.c_size_t(), ctypes.c_size_t()
        check(self.lib.rb_gl2_copy_trajectory(self.handle, ctypes.byref(tp), ctypes.byref(tc), ctypes.byref(sp), ctypes.byref(sc)),
              "rb_gl2_copy_trajectory")
        times = np.ctypeslib.as_array(tp, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
        states = np.ctypeslib.as_array(sp, shape=(sc.value, 3 * self.N)).copy() if sc.value else np.zeros((0, 3 * self.N))
        if tc.value:
            self.lib.rb_free(ctypes.cast(tp, ctypes.c_void_p))
        self.lib.rb_free(ctypes.cast(sp, ctypes.c_void_p))
        return times, states

    def stats(self):
        st = _lib.rb_gl2_stats()
        check(self.lib.rb_gl2_get_stats(self.handle, ctypes.byref(st)), "rb_gl2_get_stats")
        return {n: getattr(st, n) for n, _ in st._fields_}


class AugmentedRungeKuttaStepper:
    """AutonomousRungeKuttaStepper<std_complex, 3N>(AugmentedBoundaryIntegrator&, tstep) (A/kernel.cu:85-96)."""

    def __init__(self, problem: AugmentedBoundaryIntegrator, tstep: float = 1e-2):
        self.problem = problem
        self.lib = problem.lib
        v = problem.variables._c()
        self.handle = self.lib.rb_aug_rk4_create(problem.integrator.handle, ctypes.byref(v), float(tstep))
        if not self.handle:
            raise _lib.RobertsError("rb_aug_rk4_create: " + self.lib.rb_last_error().decode())
        self._keep = None

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_aug_rk4_destroy(h)

    def setTimeStep(self, tstep):
        check(self.lib.rb_aug_rk4_set_time_step(self.handle, float(tstep)), "rb_aug_rk4_set_time_step")

    def initialize(self, devY0, onDevice=False):
        if onDevice:
            self._keep = devY0
            check(self.lib.rb_aug_rk4_initialize(self.handle, _ptr(devY0), 1), "rb_aug_rk4_initialize")
        else:
            host = np.ascontiguousarray(np.asarray(devY0, dtype=np.complex128))
            check(self.lib.rb_aug_rk4_initialize(self.handle, host.ctypes.data_as(ctypes.c_void_p), 0), "rb_aug_rk4_initialize")

    def runStep(self, _step=0):
        check(self.lib.rb_aug_rk4_step(self.handle), "rb_aug_rk4_step")

    def runSteps(self, steps):
        check(self.lib.rb_aug_rk4_run_steps(self.handle, int(steps)), "rb_aug_rk4_run_steps")

    def runEvolution(self, startTime, endTime):
        n = ctypes.c_size_t()
        check(self.lib.rb_aug_rk4_evolve(self.handle, float(startTime), float(endTime), ctypes.byref(n)), "rb_aug_rk4_evolve")
        return n.value

    def getState(self):
        host = np.empty(3 * self.problem.N * self.problem.batchSize, np.complex128)
        check(self.lib.rb_aug_rk4_get_state(self.handle, host.ctypes.data_as(ctypes.c_void_p)), "rb_aug_rk4_get_state")
        return host

    def currentTime(self):
        return self.lib.rb_aug_rk4_current_time(self.handle)


# ---- the kernels the reference's tests launch by name ------------------------------------------
def createMKernel(A, Z, Zp, Zpp, rho, n, batchSize=1):
    check(_lib.load().rb_create_M(_ptr(A), _ptr(Z), _ptr(Zp), _ptr(Zpp), float(rho), int(n), int(batchSize), _stream_ptr(A.device)),
          "rb_create_M")


def createFiniteDepthMKernel(A, Z, Zp, Zpp, h, n, batchSize=1, infinite_depth=False):
    check(_lib.load().rb_create_finite_depth_M(_ptr(A), _ptr(Z), _ptr(Zp), _ptr(Zpp), float(h), int(n), int(batchSize),
                                               int(infinite_depth), _stream_ptr(A.device)), "rb_create_finite_depth_M")


def createVelocityMatrices(Z, Zp, Zpp, N, out1, out2, lower=True, batchSize=1):
    check(_lib.load().rb_velocity_matrices(_ptr(Z), _ptr(Zp), _ptr(Zpp), int(N), _ptr(out1), _ptr(out2), int(lower),
                                           int(batchSize), _stream_ptr(Z.device)), "rb_velocity_matrices")


def createHeliumVelocityMatrices(Z, Zp, Zpp, h, N, out1, out2, lower=True, batchSize=1, infinite_depth=False):
    check(_lib.load().rb_helium_velocity_matrices(_ptr(Z), _ptr(Zp), _ptr(Zpp), float(h), int(N), _ptr(out1), _ptr(out2),
                                                  int(lower), int(batchSize), int(infinite_depth), _stream_ptr(Z.device)),
          "rb_helium_velocity_matrices")


def compute_rhs_phi_expression(Z, V1, V2, result, rho, N):
    check(_lib.load().rb_rhs_phi_water(_ptr(Z), _ptr(V1), _ptr(V2), _ptr(result), float(rho), int(N), _stream_ptr(Z.device)),
          "rb_rhs_phi_water")


def compute_rhs_helium_phi_expression(Z, V1, result, h, N):
    check(_lib.load().rb_rhs_phi_helium(_ptr(Z), _ptr(V1), _ptr(result), float(h), int(N), _stream_ptr(Z.device)),
          "rb_rhs_phi_helium")


def compute_rhs_helium_phi_expression_with_surface_tension(Z, Zp, Zpp, V1, result, h, kappa, N):
    check(_lib.load().rb_rhs_phi_helium_surface_tension(_ptr(Z), _ptr(Zp), _ptr(Zpp), _ptr(V1), _ptr(result), float(h),
                                                        float(kappa), int(N), _stream_ptr(Z.device)),
          "rb_rhs_phi_helium_surface_tension")


def compute_rhs_helium_phi_expression_expansion_terms(Z, V1, result, h, N, order=2):
    check(_lib.load().rb_rhs_phi_helium_expansion(_ptr(Z), _ptr(V1), _ptr(result), float(h), int(N), int(order),
                                                  _stream_ptr(Z.device)), "rb_rhs_phi_helium_expansion")


@dataclasses.dataclass
class RK45_Options:
    """L/RK45.cuh:21-27."""
    atol: float = 1e-6
    rtol: float = 1e-3
    h_min: float = 1e-16
    h_max: float = 1e10
    initial_timestep: float = 1e-2

    def _c(self):
        return _lib.rb_rk45_options(self.atol, self.rtol, self.h_min, self.h_max, self.initial_timestep)


class RK45_std_complex:
    """RK45_std_complex<N>(AutonomousProblem&, logger, valueLoggers, tstep, h_max, h_min) (L/RK45.cuh:103-180, 333-400): adaptive
    Runge-Kutta-Fehlberg 4(5).  `autonomousProblem` is either a BaseBoundaryIntegralCalculator (the boundary-integral RHS runs
    inside the library) or any object with `run(state, rhs)` taking complex128 CUDA tensors of length `n`, the Python counterpart
    of AutonomousProblem<T,N>::run (L/AutonomousProblem.h:9-28)."""

    StepAccepted, StepRejected = "StepAccepted", "StepRejected"
    ReachedEndTime, StiffnessDetected = "ReachedEndTime", "StiffnessDetected"

    def __init__(self, autonomousProblem, tstep: float = 1e-2, h_max: float = 1e10, h_min: float = 1e-16, n: int = None,
                 device=None):
        self.lib = _lib.load()
        opt = RK45_Options(h_min=h_min, h_max=h_max, initial_timestep=tstep)._c()
        self.problem = autonomousProblem
        if isinstance(autonomousProblem, BaseBoundaryIntegralCalculator):
            self.device = autonomousProblem.device
            self.n = 2 * autonomousProblem.N * autonomousProblem.batchSize
            self.handle = self.lib.rb_rk45_create(autonomousProblem.handle, ctypes.byref(opt))
        else:
            if n is None:
                raise ValueError("a generic problem needs the state length n")
            self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            self.n = int(n)
            dev = self.device

            def _run(_user, state_ptr, rhs_ptr, _stream):
                autonomousProblem.run(_view(state_ptr, 2 * self.n, dev, True), _view(rhs_ptr, 2 * self.n, dev, True))

            self._cb = _lib.RB_RHS_FN(_run)   # keep the trampoline alive
            self.handle = self.lib.rb_rk45_create_generic(self.n, self._cb, None, ctypes.byref(opt), _stream_ptr(self.device))
        if not self.handle:
            raise _lib.RobertsError("rb_rk45_create: " + self.lib.rb_last_error().decode())

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            self.lib.rb_rk45_destroy(h)

    def setTolerance(self, atol, rtol):
        check(self.lib.rb_rk45_set_tolerance(self.handle, float(atol), float(rtol)), "rb_rk45_set_tolerance")

    def setOptions(self, options: RK45_Options):
        o = options._c()
        check(self.lib.rb_rk45_set_options(self.handle, ctypes.byref(o)), "rb_rk45_set_options")

    def setMaxRejectedSteps(self, maxRejected):
        check(self.lib.rb_rk45_set_max_rejected(self.handle, int(maxRejected)), "rb_rk45_set_max_rejected")

    def initialize(self, initialState, onDevice=False):
        if onDevice:
            check(self.lib.rb_rk45_initialize(self.handle, _ptr(initialState), 1), "rb_rk45_initialize")
        else:
            host = np.ascontiguousarray(np.asarray(initialState, dtype=np.complex128))
            check(self.lib.rb_rk45_initialize(self.handle, host.ctypes.data_as(ctypes.c_void_p), 0), "rb_rk45_initialize")

    def runStep(self, _i=0):
        acc = ctypes.c_int()
        check(self.lib.rb_rk45_step(self.handle, ctypes.byref(acc)), "rb_rk45_step")
        return self.StepAccepted if acc.value else self.StepRejected

    def runEvolution(self, startTime, endTime):
        res = ctypes.c_int()
        check(self.lib.rb_rk45_evolve(self.handle, float(startTime), float(endTime), ctypes.byref(res)), "rb_rk45_evolve")
        return self.ReachedEndTime if res.value == 0 else self.StiffnessDetected

    def getY(self):
        return _view(self.lib.rb_rk45_dev_state(self.handle), 2 * self.n, self.device, True)

    def getState(self):
        host = np.empty(self.n, np.complex128)
        check(self.lib.rb_rk45_get_state(self.handle, host.ctypes.data_as(ctypes.c_void_p)), "rb_rk45_get_state")
        return host

    def getCurrentTime(self):
        return self.lib.rb_rk45_current_time(self.handle)

    def getCurrentTimeStep(self):
        return self.lib.rb_rk45_current_timestep(self.handle)

    def stats(self):
        out = (ctypes.c_double * 4)()
        check(self.lib.rb_rk45_stats(self.handle, out), "rb_rk45_stats")
        return dict(accepted=int(out[0]), rejected=int(out[1]), rhs_evaluations=int(out[2]), scaled_error=out[3])


def measure_fp64_peak(device=None):
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    out = ctypes.c_double()
    check(_lib.load().rb_measure_fp64_peak(ctypes.byref(out), _stream_ptr(device)), "rb_measure_fp64_peak")
    return out.value


def measure_fp64_rate_3operand(device=None):
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    out = ctypes.c_double()
    check(_lib.load().rb_measure_fp64_rate_3operand(ctypes.byref(out), _stream_ptr(device)), "rb_measure_fp64_rate_3operand")
    return out.value


def measure_fp64_tensor_overlap(device=None):
    """DMMA m8n8k4 beside DFMA: {mix: (ms, TFLOP/s)} for the per-iteration mixes of mma.sync and fma instructions"""
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    out = (ctypes.c_double * 12)()
    check(_lib.load().rb_measure_fp64_tensor_overlap(out, _stream_ptr(device)), "rb_measure_fp64_tensor_overlap")
    names = ("8mma", "32fma", "8mma+32fma", "4mma+32fma", "2mma+32fma", "1mma+32fma")
    return {n: (out[2 * i], out[2 * i + 1]) for i, n in enumerate(names)}


# ---- legacy host-vector exports (L/Export.cuh) -------------------------------------------------
def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def calculateRHSFromVectors(x, y, phi, L, rho, kappa, depth, batchSize=None):
    lib = _lib.load()
    x, y, phi = (np.ascontiguousarray(v, np.float64) for v in (x, y, phi))
    vx, vy, dphi = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    if batchSize is None:
        rc = lib.calculateRHSFromVectors(_dp(x), _dp(y), _dp(phi), _dp(vx), _dp(vy), _dp(dphi), L, rho, kappa, depth, x.size)
    else:
        rc = lib.calculateRHS256FromVectorsBatched(_dp(x), _dp(y), _dp(phi), _dp(vx), _dp(vy), _dp(dphi), L, rho, kappa, depth,
                                                   int(batchSize))
    check(rc, "calculateRHSFromVectors")
    return vx, vy, dphi


def integrateSimulationRK4(initialState, simProperties: _lib.SimProperties, rkOptions: _lib.RK4SolverOptions, N):
    """Linux-safe binding of the export the reference declares (L/Export.cuh:69); structs go by reference."""
    lib = _lib.load()
    st = np.ascontiguousarray(initialState, np.float64)
    so, to = ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)()
    sc, tc = ctypes.c_size_t(), ctypes.c_size_t()
    check(lib.integrateSimulationRK4(_dp(st), ctypes.byref(so), ctypes.byref(sc), ctypes.byref(to), ctypes.byref(tc),
                                     ctypes.byref(simProperties), ctypes.byref(rkOptions), int(N)), "integrateSimulationRK4")
    states = np.ctypeslib.as_array(so, shape=(sc.value, 3 * N)).copy() if sc.value else np.zeros((0, 3 * N))
    times = np.ctypeslib.as_array(to, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
    lib.integrateSimulationRK4_freeMemory(so, to)
    return states, times


def integrate_rk4_host(initialState, N, batch, properties: ProblemProperties, physics, dt, steps, guess="warm",
                       tolerance=1e-13):
    lib = _lib.load()
    p = _lib.rb_props()
    lib.rb_default_props(ctypes.byref(p))
    p.rho, p.U, p.kappa, p.depth = properties.rho, properties.U, properties.kappa, properties.depth
    p.use_expansions, p.expansion_order = int(properties.use_expansions), int(properties.expansion_order)
    p.infinite_depth = int(properties.infinite_depth)
    p.physics = PHYSICS[physics]
    p.guess_mode = _lib.RB_GUESS_WARM if guess == "warm" else _lib.RB_GUESS_COLD
    p.tolerance = tolerance
    st = np.ascontiguousarray(initialState, np.float64)
    out = np.empty_like(st)
    check(lib.rb_integrate_rk4_host(_dp(st), _dp(out), int(N), int(batch), ctypes.byref(p), float(dt), int(steps)),
          "rb_integrate_rk4_host")
    return out


def calculateRhsAugmentedOptomechanical(state, simProperties: _lib.SimProperties, optomechanicalVariables: _lib.COptomechanicalVariables, N):
    """L/Export.cuh:78: state = [x | y | phi | D] (4N doubles, SI properties / variables) -> [vx | vy | dphi/dt | dD/dt]."""
    lib = _lib.load()
    st = np.ascontiguousarray(state, np.float64)
    out = np.empty_like(st)
    check(lib.calculateRhsAugmentedOptomechanical(_dp(st), _dp(out), ctypes.byref(simProperties),
                                                  ctypes.byref(optomechanicalVariables), int(N)), "calculateRhsAugmentedOptomechanical")
    return out


def integrateAugmentedOptomechanicalSimulationRK4(initialState, simProperties, rkOptions, optomechanicalVariables, N):
    """L/Export.cuh:75: RK4 evolution of the augmented system; returns (states [count x 4N], times)."""
    lib = _lib.load()
    st = np.ascontiguousarray(initialState, np.float64)
    so, to = ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)()
    sc, tc = ctypes.c_size_t(), ctypes.c_size_t()
    check(lib.integrateAugmentedOptomechanicalSimulationRK4(_dp(st), ctypes.byref(so), ctypes.byref(sc), ctypes.byref(to),
                                                            ctypes.byref(tc), ctypes.byref(simProperties), ctypes.byref(rkOptions),
                                                            ctypes.byref(optomechanicalVariables), int(N)),
          "integrateAugmentedOptomechanicalSimulationRK4")
    states = np.ctypeslib.as_array(so, shape=(sc.value, 4 * N)).copy() if sc.value else np.zeros((0, 4 * N))
    times = np.ctypeslib.as_array(to, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
    lib.integrateAugmentedOptomechanicalSimulationRK4_freeMemory(so, to)
    return states, times


def integrateOptomechanicalSimulationRK4(initialState, simProperties, rkOptions, optomechanicalVariables, N):
    """L/Export.cuh:72: RK4 evolution with the explicitly time-dependent drive; returns (states [count x 3N], times)."""
    lib = _lib.load()
    st = np.ascontiguousarray(initialState, np.float64)
    so, to = ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)()
    sc, tc = ctypes.c_size_t(), ctypes.c_size_t()
    check(lib.integrateOptomechanicalSimulationRK4(_dp(st), ctypes.byref(so), ctypes.byref(sc), ctypes.byref(to), ctypes.byref(tc),
                                                   ctypes.byref(simProperties), ctypes.byref(rkOptions),
                                                   ctypes.byref(optomechanicalVariables), int(N)),
          "integrateOptomechanicalSimulationRK4")
    states = np.ctypeslib.as_array(so, shape=(sc.value, 3 * N)).copy() if sc.value else np.zeros((0, 3 * N))
    times = np.ctypeslib.as_array(to, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
    lib.integrateOptomechanicalSimulationRK4_freeMemory(so, to)
    return states, times


def calculateJacobian(state, L, rho, kappa, depth, epsilon, N):
    """L/Export.cuh:50: finite-difference Jacobian of the real-state helium RHS, SI properties; returns the (3N, 3N) matrix
    J[r, c] = d f_r / d y_c (the export's buffer is its column-major flattening)."""
    st = np.ascontiguousarray(state, np.float64)
    jac = np.empty(9 * N * N)
    check(_lib.load().calculateJacobian(_dp(st), _dp(jac), float(L), float(rho), float(kappa), float(depth), float(epsilon), int(N)),
          "calculateJacobian")
    return jac.reshape(3 * N, 3 * N).T.copy()


def calculatePerturbedStates256(x, y, phi, L, rho, kappa, depth, epsilon):
    """L/Export.cuh:53: the 3N perturbed states (N = 256) as a complex array of 6 N^2 entries."""
    out = np.empty(6 * 256 * 256, np.complex128)
    xs, ys, ps = (np.ascontiguousarray(v, np.float64) for v in (x, y, phi))
    check(_lib.load().calculatePerturbedStates256(_dp(xs), _dp(ys), _dp(ps), out.ctypes.data_as(ctypes.c_void_p), float(L), float(rho),
                                                  float(kappa), float(depth), float(epsilon)), "calculatePerturbedStates256")
    return out


def integrateSimulationGL2(initialState, simProperties: _lib.SimProperties, glCOptions: _lib.GaussLegendreOptions, N):
    """L/Export.cuh:66: implicit Gauss-Legendre-2 evolution; returns (states [count x 3N], times)."""
    lib = _lib.load()
    st = np.ascontiguousarray(initialState, np.float64)
    so, to = ctypes.POINTER(ctypes.c_double)(), ctypes.POINTER(ctypes.c_double)()
    sc, tc = ctypes.c_size_t(), ctypes.c_size_t()
    check(lib.integrateSimulationGL2(_dp(st), ctypes.byref(so), ctypes.byref(sc), ctypes.byref(to), ctypes.byref(tc),
                                     ctypes.byref(simProperties), ctypes.byref(glCOptions), int(N)), "integrateSimulationGL2")
    states = np.ctypeslib.as_array(so, shape=(sc.value, 3 * N)).copy() if sc.value else np.zeros((0, 3 * N))
    times = np.ctypeslib.as_array(to, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
    lib.integrateSimulationGL2_freeMemory(so, to)
    return states, times
