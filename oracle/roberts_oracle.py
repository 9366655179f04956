"""CPU oracle for the Roberts (1983) boundary-integral RK4 time step.

TEST INFRASTRUCTURE ONLY.  This module restates, in vectorised NumPy, the
arithmetic of the reference's CUDA path (CuSuperHelium) for one RHS evaluation
and one classical RK4 step.  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product path (``superfluid_dynamics_b200``) never does.

Reference citations use the prefixes of SURVEY.md:
  L/ = CuSuperHelium/CuSuperHelium/         (CUDA library)
  T/ = CuSuperHelium/CuSuperHelium.Tests/   (gtest fixtures)
  P/ = CuSuperHelium/Python/                (NumPy statement of the water path)

Parity pinning (see tests/test_oracle.py and tests/golden/):
  * water path: pinned against P/WaterIntegralCalculator.py run in the build
    container (golden vectors committed under tests/golden/), and against the
    closed forms / host loops of T/MatrixMTests.cuh.
  * helium paths (createFiniteDepthMKernel, createHeliumVelocityMatrices,
    compute_rhs_helium_phi_expression*; infinite depth, surface tension,
    expansion terms): the reference holds no CPU statement, no test and no
    golden vector of its own for these kernels.  Pinned instead against
    OUTPUTS OF THE REFERENCE ITSELF: its CUDA classes compiled unmodified
    (oracle/build_ref.py) and run on a B200, vectors committed as
    tests/golden/ref_cuda_golden.npz with the generating script
    (tests/golden/make_ref_cuda_golden.py); the same file re-pins the water
    path at the whole-RHS and 100-RK4-step level
    (tests/test_oracle.py::test_oracle_matches_reference_cuda_golden).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

PI = math.pi
ALPHA_HAMAKER = 3.5e-24  # L/constants.cuh:11


# --------------------------------------------------------------------------
# problem description (L/ProblemProperties.hpp:5-32)
# --------------------------------------------------------------------------
@dataclasses.dataclass
class ProblemProperties:
    L: float = 1.0
    rho: float = 1.0
    U: float = 0.0
    kappa: float = 0.0
    depth: float = 1.0
    initial_amplitude: float = 1.0
    use_expansions: bool = False
    expansion_order: int = 1
    infinite_depth: bool = False
    base_length: float = 1.0
    base_time: float = 1.0
    base_energy: float = 1.0
    base_acceleration: float = 1.0


def adimensionalize_properties(props: ProblemProperties, rho_helium: float = 150.0) -> ProblemProperties:
    """SI -> nondimensional conversion of L/Export.cu:1222-1246."""
    p = dataclasses.replace(props)
    p.base_length = p.L / (2.0 * PI)
    p.base_acceleration = 3 * ALPHA_HAMAKER / p.depth ** 4
    p.base_time = math.sqrt(p.base_length / p.base_acceleration)
    p.base_energy = 3.0 * rho_helium * ALPHA_HAMAKER * p.base_length ** 4 / p.depth ** 4
    surface_tension_factor = rho_helium * p.base_length ** 3 / (p.base_time * p.base_time)
    p.kappa = p.kappa / surface_tension_factor
    p.depth = p.depth / p.base_length
    p.rho = p.rho / rho_helium
    return p


def adimensionalize_rk4_options(time_step: float, t0: float, t1: float, props: ProblemProperties):
    """L/Export.cu:1213-1220."""
    return time_step / props.base_time, t0 / props.base_time, t1 / props.base_time


# --------------------------------------------------------------------------
# initial surfaces (T/MatrixMTests.cuh:19-67, P/StokeWaves.py:3-8,
# L/SimulationFunctions.cuh:18-28)
# --------------------------------------------------------------------------
def trochoid(N: int, h: float, omega: float = 1.0, t: float = 0.0, rho: float = 0.0):
    """Z = X + iY and Phi of the analytic trochoidal wave sampled at alpha_j = 2 pi j / N."""
    a = 2.0 * PI * np.arange(N) / N
    X = a - h * np.sin(a - omega * t)
    Y = h * np.cos(a - omega * t)
    Phi = h * (1.0 + rho) * omega * np.sin(a - omega * t)
    return X + 1j * Y, Phi.astype(np.float64)


def trochoid_derivatives(N: int, h: float, omega: float = 1.0, t: float = 0.0, rho: float = 0.0):
    """Per-index analytic derivatives of the trochoid (prepareZPhi, T/MatrixMTests.cuh:51-67)."""
    a = 2.0 * PI * np.arange(N) / N
    s = 2.0 * PI / N
    Zp = (1 - h * np.cos(a - omega * t)) * s + 1j * (-h * np.sin(a - omega * t)) * s
    Zpp = (h * np.sin(a - omega * t)) * s * s + 1j * (-h * np.cos(a - omega * t)) * s * s
    PhiP = h * (1.0 + rho) * omega * np.cos(a - omega * t) * s
    return Zp, Zpp, PhiP


def sinusoid(N: int, eps: float):
    """Small-amplitude deep-water wave: X=alpha, Y=eps cos, Phi=eps sin."""
    a = 2.0 * PI * np.arange(N) / N
    return a + 1j * eps * np.cos(a), eps * np.sin(a)


def pack_state(Z: np.ndarray, Phi: np.ndarray) -> np.ndarray:
    """[Z | Phi+0i] complex128 layout of L/BaseBoundaryIntegrator.cuh:141-145."""
    return np.concatenate([np.asarray(Z, np.complex128).ravel(), np.asarray(Phi, np.float64).ravel().astype(np.complex128)])


# --------------------------------------------------------------------------
# spectral derivatives
# --------------------------------------------------------------------------
def d1_cuda(x: np.ndarray) -> np.ndarray:
    """First derivative per index, CUDA semantics (L/utilities.cuh:106-148 + L/Derivatives.cuh:190-257).

    modes 0..N/2-1: *ik ; mode N/2: *i*pi*(N/2) ; mode N/2+1: zeroed ; modes > N/2+1: *i(k-N).
    The 1/n of the unnormalised inverse cuFFT is folded into the multiply.
    Operates along the last axis (batched layout [b][N])."""
    x = np.asarray(x, np.complex128)
    n = x.shape[-1]
    c = np.fft.fft(x, axis=-1)
    i = np.arange(n)
    fac = np.where(i < n // 2, i, i - n).astype(np.float64)
    r = np.empty_like(c)
    # result.x = -i*y/n ; result.y = i*x/n   (same operation order as the kernel)
    r.real = -fac * c.imag / float(n)
    r.imag = fac * c.real / float(n)
    if n // 2 < n:
        m = n // 2
        r[..., m] = (-PI * m * c[..., m].imag / float(n)) + 1j * (PI * m * c[..., m].real / float(n))
    if n // 2 + 1 < n:
        r[..., n // 2 + 1] = 0.0
    return np.fft.ifft(r, axis=-1) * n  # unnormalised inverse


def d2_cuda(x: np.ndarray) -> np.ndarray:
    """Second derivative per index (L/utilities.cuh:178-200): *(-k^2)/n for every mode."""
    x = np.asarray(x, np.complex128)
    n = x.shape[-1]
    c = np.fft.fft(x, axis=-1)
    i = np.arange(n)
    fac = np.where(i <= n // 2, i, i - n).astype(np.float64)
    r = -(fac * fac) * c / float(n)
    return np.fft.ifft(r, axis=-1) * n


def d1_python(x: np.ndarray) -> np.ndarray:
    """First derivative, semantics of P/Derivatives.py:6-18 (differs from CUDA at the Nyquist mode only)."""
    x = np.asarray(x, np.complex128)
    n = x.shape[-1]
    idx = np.fft.fftshift(np.arange(-n / 2, n / 2))
    c = np.fft.fft(x, axis=-1)
    nyq = c[..., n // 2].copy()
    r = 1j * idx * c
    if n // 2 + 1 < n:
        r[..., n // 2 + 1] = 0
    r[..., n // 2] = -PI * np.real(nyq)
    return np.fft.ifft(r, axis=-1)


def d2_python(x: np.ndarray) -> np.ndarray:
    """P/Derivatives.py:19-27."""
    x = np.asarray(x, np.complex128)
    n = x.shape[-1]
    idx = np.power(1j * np.fft.fftshift(np.arange(-n / 2, n / 2)), 2)
    return np.fft.ifft(np.fft.fft(x, axis=-1) * idx, axis=-1)


def fft_derivative(x, scale=1.0, second=False, deriv="cuda"):
    """FftDerivative<N,B>::exec (L/Derivatives.cuh:190-257)."""
    if deriv == "cuda":
        r = d2_cuda(x) if second else d1_cuda(x)
    else:
        r = d2_python(x) if second else d1_python(x)
    return r * scale if scale != 1.0 else r


def zphi_derivative(Z, Phi, props: ProblemProperties, deriv="cuda"):
    """ZPhiDerivative<N,B>::exec (L/Derivatives.cuh:311-384).  Returns Zp, PhiPrime (complex), Zpp."""
    Z = np.asarray(Z, np.complex128)
    n = Z.shape[-1]
    j = np.arange(n, dtype=np.float64)
    zlin = 2 * PI * j / n
    philin = -(1 + props.rho) * PI * props.U / n * j
    zper = Z - zlin
    phiper = np.asarray(Phi, np.complex128) - philin
    s = 2.0 * PI / n
    Zpp = fft_derivative(zper, 4.0 * PI * PI / (n * n), second=True, deriv=deriv)
    Zp = fft_derivative(zper, s, deriv=deriv)
    PhiP = fft_derivative(phiper, s, deriv=deriv)
    Zp = Zp + s
    if props.U != 0:
        PhiP = PhiP + (-(1 + props.rho) * PI * props.U / n)
    return Zp, PhiP, Zpp


# --------------------------------------------------------------------------
# cotangent kernels
# --------------------------------------------------------------------------
def cot(z):
    """cot(z) = 1/tan(z) as in L/utilities.cuh:312-315."""
    return 1.0 / np.tan(z)


def _cot_green(Z):
    """C[k, j] = cot((Z_k - Z_j)/2) (L/utilities.cuh:347-350); diagonal set to 0."""
    d = 0.5 * (Z[:, None] - Z[None, :])
    np.fill_diagonal(d, 1.0)  # placeholder, avoids the pole
    C = cot(d)
    np.fill_diagonal(C, 0.0)
    return C


def create_M(Z, Zp, Zpp, rho: float) -> np.ndarray:
    """createMKernel (L/createM.cuh:43-63).  Returns M[k, j] (row k, col j).

    The device buffer is column-major A[k + j*n]; use ``M.T.ravel()`` /
    ``order='F'`` to compare with the flat device layout."""
    C = _cot_green(np.asarray(Z, np.complex128))
    M = 0.25 * (1 - rho) / PI * (Zp[:, None] * C).imag
    np.fill_diagonal(M, 0.5 * (1 + rho) + 0.25 * (1 - rho) / PI * (Zpp / Zp).imag)
    return M


def create_finite_depth_M(Z, Zp, Zpp, h: float, infinite_depth: bool = False) -> np.ndarray:
    """createFiniteDepthMKernel (L/createM.cuh:65-92).  The image term carries no Zp_k factor (quirk 4)."""
    Z = np.asarray(Z, np.complex128)
    C = _cot_green(Z)
    M = 0.25 / PI * (Zp[:, None] * C).imag
    diag = 0.5 + 0.25 / PI * (Zpp / Zp).imag
    if not infinite_depth:
        img = cot(0.5 * (Z[:, None] - np.conj(Z)[None, :]) + 1j * h)
        M = M - 0.25 / PI * img.imag
        diag = diag - 0.25 / PI * cot(1j * (Z.imag + h)).imag
    np.fill_diagonal(M, diag)
    return M


def velocity_matrices(Z, Zp, Zpp, lower: bool = True):
    """createVelocityMatrices (L/WaterVelocities.cuh:38-70).  Returns V1[k, j], V2[k]."""
    Z = np.asarray(Z, np.complex128)
    C = _cot_green(Z)
    V1 = 1j * (-1.0 / (4.0 * PI) * C)
    diag = 1j * (-1.0 / (4.0 * PI) * Zpp / np.power(Zp, 2.0))
    diag = diag + 1.0 / (2.0 * Zp) if lower else diag - 0.5 / Zp
    np.fill_diagonal(V1, diag)
    V2 = 1j * (1.0 / (2.0 * PI * Zp))
    return V1, V2


def helium_velocity_matrices(Z, Zp, Zpp, h: float, lower: bool = True, infinite_depth: bool = False):
    """createHeliumVelocityMatrices (L/WaterVelocities.cuh:72-107)."""
    Z = np.asarray(Z, np.complex128)
    C = _cot_green(Z)
    V1 = 1j * (-1.0 / (4.0 * PI) * C)
    diag = 1j * (-1.0 / (4.0 * PI) * Zpp / np.power(Zp, 2.0))
    if not infinite_depth:
        V1 = V1 + 1j * (1.0 / (4.0 * PI) * cot(0.5 * (Z[:, None] - np.conj(Z)[None, :]) + 1j * h))
        diag = diag + 1j * (1.0 / (4.0 * PI) * cot(1j * (Z.imag + h)))
    diag = diag + 1.0 / (2.0 * Zp) if lower else diag - 0.5 / Zp
    np.fill_diagonal(V1, diag)
    V2 = 1j * (1.0 / (2.0 * PI * Zp))
    return V1, V2


# --------------------------------------------------------------------------
# dPhi/dt
# --------------------------------------------------------------------------
def rhs_phi_water(Z, V1, V2, rho: float):
    """compute_rhs_phi_expression (L/createM.cuh:96-107) including the V1[1] quirk on the dot product."""
    v1a = V1.real ** 2 + V1.imag ** 2
    v2a = V2.real ** 2 + V2.imag ** 2
    dot = V1[1].real * V2.real + V1.imag * V2.imag
    return -(1 + rho) * Z.imag + 0.5 * v1a + 0.5 * rho * v2a - rho * dot


def rhs_phi_helium(Z, V1, h: float):
    """compute_rhs_helium_phi_expression (L/createM.cuh:109-117)."""
    vdw = h / 3.0
    return vdw * np.power(1.0 + Z.imag / h, -3.0) - vdw + 0.5 * V1.real * V1.real + 0.5 * V1.imag * V1.imag


def rhs_phi_helium_surface_tension(Z, Zp, Zpp, V1, h: float, kappa: float):
    """compute_rhs_helium_phi_expression_with_surface_tension (L/createM.cuh:195-211)."""
    curv = (Zp.real * Zpp.imag - Zp.imag * Zpp.real) / np.power(Zp.real ** 2 + Zp.imag ** 2, 1.5)
    v1a = V1.real ** 2 + V1.imag ** 2
    return 20.447761896665416 * h / 3.0 * (1.0 / np.power(1.0 + Z.imag / h, 3) - 1) + 0.5 * v1a + kappa * curv


def rhs_phi_helium_expansion(Z, V1, h: float, order: int = 2):
    """compute_rhs_helium_phi_expression_expansion_terms (L/createM.cuh:171-193); switch fall-through."""
    kin = 0.5 * V1.real * V1.real + 0.5 * V1.imag * V1.imag
    vdw = np.zeros_like(kin)
    if order == 3:
        vdw = vdw + (-10.0 / 3.0 * np.power(Z.imag, 3.0) / (h * h))
    if order in (2, 3):
        vdw = vdw + 2.0 * np.power(Z.imag, 2.0) / h
    if order in (1, 2, 3):
        vdw = vdw + (-Z.imag)
    return vdw + kin


# --------------------------------------------------------------------------
# one RHS (L/BaseBoundaryIntegrator.cuh:138-306)
# --------------------------------------------------------------------------
PHYSICS = ("water", "helium", "helium_inf")


def rhs_single(Z, Phi, props: ProblemProperties, physics: str = "water", deriv: str = "cuda", full: bool = False):
    """One RHS evaluation for one surface.  Returns (velocities u+iv, dPhi/dt[, intermediates])."""
    n = len(Z)
    Z = np.asarray(Z, np.complex128)
    Zp, PhiPc, Zpp = zphi_derivative(Z, Phi, props, deriv)
    b = PhiPc.real.copy()  # complex_to_real, L/BaseBoundaryIntegrator.cuh:299
    if physics == "water":
        M = create_M(Z, Zp, Zpp, props.rho)  # L/WaterBoundaryProblem.cuh:19
    elif physics == "helium":
        M = create_finite_depth_M(Z, Zp, Zpp, props.depth, props.infinite_depth)  # L/HeliumBoundaryProblem.cuh:17
    elif physics == "helium_inf":
        M = create_M(Z, Zp, Zpp, props.depth)  # depth in the rho slot, L/HeliumBoundaryProblem.cuh:61 (quirk 6)
    else:
        raise ValueError(physics)
    a = np.linalg.solve(M, b)  # MatrixSolver::solve, partial-pivot LU, L/MatrixSolver.cuh:114-125
    ap = fft_derivative(a.astype(np.complex128), 2.0 * PI / n, deriv=deriv)  # L/BaseBoundaryIntegrator.cuh:201-203

    def vel(lower):
        if physics == "helium":
            V1, V2 = helium_velocity_matrices(Z, Zp, Zpp, props.depth, lower, props.infinite_depth)
        else:
            V1, V2 = velocity_matrices(Z, Zp, Zpp, lower)
        w = V2 * ap + V1 @ a.astype(np.complex128)  # L/WaterVelocities.cuh:217-224
        return np.conj(w)  # :241

    v_lower = vel(True)
    v_upper = vel(False)
    if physics == "water":
        dphi = rhs_phi_water(Z, v_lower, v_upper, props.rho)
    elif physics == "helium":
        if props.use_expansions:
            dphi = rhs_phi_helium_expansion(Z, v_lower, props.depth, props.expansion_order)
        elif props.kappa != 0.0:
            dphi = rhs_phi_helium_surface_tension(Z, Zp, Zpp, v_lower, props.depth, props.kappa)
        else:
            dphi = rhs_phi_helium(Z, v_lower, props.depth)
    else:
        dphi = rhs_phi_helium(Z, v_lower, props.depth)
    if full:
        return v_lower, dphi, dict(Zp=Zp, Zpp=Zpp, PhiPrime=b, M=M, a=a, aprime=ap, v_upper=v_upper)
    return v_lower, dphi


def rhs(state: np.ndarray, N: int, batch: int, props: ProblemProperties, physics="water", deriv="cuda") -> np.ndarray:
    """AutonomousProblem::run on the packed complex state [Z_b0..Z_b(B-1) | Phi_b0..] (L/BaseBoundaryIntegrator.cuh:141-146)."""
    state = np.asarray(state, np.complex128)
    out = np.zeros_like(state)
    for b in range(batch):
        Z = state[b * N:(b + 1) * N]
        Phi = state[batch * N + b * N: batch * N + (b + 1) * N]
        v, dphi = rhs_single(Z, Phi, props, physics, deriv)
        out[b * N:(b + 1) * N] = v
        out[batch * N + b * N: batch * N + (b + 1) * N] = dphi
    return out


# --------------------------------------------------------------------------
# classical RK4 (L/AutonomousRungeKuttaStepper.cuh:124-307, 344-374, 418-437)
# --------------------------------------------------------------------------
def rk4_step(f, y0: np.ndarray, dt: float) -> np.ndarray:
    k1 = f(y0)
    y1 = y0 + (dt * 0.5) * k1
    k2 = f(y1)
    y2 = y0 + (dt * 0.5) * k2
    k3 = f(y2)
    y3 = y0 + dt * k3
    k4 = f(y3)
    ksum = k1 + 2.0 * k2 + 2.0 * k3 + k4  # add_k_vectors, L/utilities.cuh:78-83
    return y0 + (dt / 6.0) * ksum


def rk4_num_steps(t0: float, t1: float, dt: float) -> int:
    """steps = size_t((t1 - t0)/dt), truncation (L/AutonomousRungeKuttaStepper.cuh:421)."""
    return int((t1 - t0) / dt)


def rk4_evolve(f, y0: np.ndarray, t0: float, t1: float, dt: float, trajectory: bool = False):
    y = np.array(y0, np.complex128)
    steps = rk4_num_steps(t0, t1, dt)
    t = t0
    times, states = [], []
    for _ in range(steps):
        y = rk4_step(f, y, dt)
        t += dt
        if trajectory:
            times.append(t)
            states.append(y.copy())
    if trajectory:
        return y, np.array(times), np.array(states)
    return y


# --------------------------------------------------------------------------
# adaptive Runge-Kutta-Fehlberg 4(5) (L/RK45.cuh:194-330, L/RK45_Kernels.cuh:15-106)
# --------------------------------------------------------------------------
@dataclasses.dataclass
class RK45Options:
    """RK45_Options, L/RK45.cuh:21-27."""
    atol: float = 1e-6
    rtol: float = 1e-3
    h_min: float = 1e-16
    h_max: float = 1e10
    initial_timestep: float = 1e-2


# Fehlberg tableau, L/RK45_Kernels.cuh:18-45
_RKF_A = ((1.0 / 4.0,),
          (3.0 / 32.0, 9.0 / 32.0),
          (1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0),
          (439.0 / 216.0, -8.0, 3680.0 / 513.0, -845.0 / 4104.0),
          (-8.0 / 27.0, 2.0, -3544.0 / 2565.0, 1859.0 / 4104.0, -11.0 / 40.0))
_RKF_B5 = (16.0 / 135.0, 0.0, 6656.0 / 12825.0, 28561.0 / 56430.0, -9.0 / 50.0, 2.0 / 55.0)
_RKF_B4 = (25.0 / 216.0, 0.0, 1408.0 / 2565.0, 2197.0 / 4104.0, -1.0 / 5.0, 0.0)


def rk45_new_timestep(old_h: float, error: float, accepting: bool, h_min: float, h_max: float) -> float:
    """calculateNewTimestep, L/RK45.cuh:306-330."""
    safety, minfac, maxfac, expo = 0.9, 0.2, 5.0, 1.0 / 5.0
    if error == 0.0:
        return old_h * (maxfac if accepting else 1.0)
    fac = safety * error ** (-expo)
    fac = min(max(fac, minfac), maxfac if accepting else 1.0)
    return min(max(old_h * fac, h_min), h_max)


class RK45:
    """RK45Base<T,N>::runStep / runEvolution (L/RK45.cuh:194-304) for f(y) -> dy/dt on a complex vector."""

    def __init__(self, f, options: RK45Options = None, max_rejected: int = 500):
        o = options or RK45Options()
        self.f, self.atol, self.rtol, self.h_min, self.h_max = f, o.atol, o.rtol, o.h_min, o.h_max
        self.h, self.t = o.initial_timestep, 0.0
        self.max_rejected = max_rejected
        self.accepted_prev, self.k1, self.y = True, None, None
        self.n_accepted = self.n_rejected = self.n_rhs = 0
        self.scaled_error = 0.0

    def initialize(self, y0):
        self.y = np.array(y0, np.complex128)
        self.accepted_prev = True

    def _rhs(self, y):
        self.n_rhs += 1
        return self.f(y)

    def run_step(self) -> bool:
        h, y = self.h, self.y
        if self.accepted_prev:                       # a rejected attempt keeps k1 (:262-264)
            self.k1 = self._rhs(y)
        k = [self.k1]
        for row in _RKF_A:                           # calculateTempY (:341-367): c1 k1 + c2 k2 + ... + y, left to right
            acc = (row[0] * h) * k[0]
            for c, kj in zip(row[1:], k[1:]):
                acc = acc + (c * h) * kj
            k.append(self._rhs(acc + y))
        y5 = y.copy()                                # rk45_error_and_y5 (L/RK45_Kernels.cuh:63-106)
        e = np.zeros_like(y)
        for b5, b4, kj in zip(_RKF_B5, _RKF_B4, k):
            y5 = y5 + (h * b5) * kj
            e = e + (h * (b5 - b4)) * kj
        sc = np.maximum(self.atol + self.rtol * np.maximum(np.abs(y), np.abs(y5)), 1e-300)
        self.scaled_error = float(np.sqrt((1.0 / y.size) * np.sum((np.abs(e) / sc) ** 2)))   # :399
        accepted = self.scaled_error <= 1.0
        self.accepted_prev = accepted
        h_new = rk45_new_timestep(h, self.scaled_error, accepted, self.h_min, self.h_max)
        if accepted:
            self.y = y5
            self.t += h
            self.n_accepted += 1
        else:
            self.n_rejected += 1
        self.h = h_new
        return accepted

    def run_evolution(self, t0: float, t1: float) -> str:
        self.t = t0
        while True:
            if self.t >= t1:
                return "ReachedEndTime"
            self.h = min(self.h, t1 - self.t)        # never overshoot (:207-211)
            rejected = 0
            while not self.run_step():
                rejected += 1
                if rejected > self.max_rejected:
                    return "StiffnessDetected"


# --------------------------------------------------------------------------
# diagnostics (L/Energies.cuh:136-204 functors, scaling :61-128); batch 0 only
# --------------------------------------------------------------------------
def energies(Z, Zp, Phi, vel, props: ProblemProperties, physics="water") -> dict:
    Z = np.asarray(Z, np.complex128)
    Phi = np.asarray(Phi, np.complex128)
    U, rho = props.U, props.rho
    kin = (Phi.real + 0.5 * U * (1.0 + rho) * Z.real) * (-1.0 * Zp.imag * vel.real + Zp.real * vel.imag) \
        - 0.5 * U * ((vel.real + rho * vel.real) * Zp.real + (vel.imag + rho * vel.imag) * Zp.imag
                     + 0.5 * U * (1.0 - rho) * Zp.real) * Z.imag
    out = {"kinetic": kin.sum() * 0.25 / PI}
    if physics == "water":
        out["potential"] = (Z.imag * Z.imag * Zp.real).sum() * 0.25 * (1.0 + rho) / PI
    else:
        out["potential"] = (1 / np.power(1 + Z.imag / props.depth, 2) - 1.0).sum() * props.depth ** 2 / 6.0
    out["surface"] = (np.sqrt(Zp.real ** 2 + Zp.imag ** 2).sum() - 2.0 * PI) * props.kappa / (2.0 * PI)
    out["volume_flux"] = (vel.imag * Zp.real + vel.real * Zp.imag).sum() * 0.5 / PI
    return out


def volume(Z, Zp) -> float:
    """Area under the surface per period, sum Y_j X'_j (SURVEY.md section 8d 'Parity')."""
    return float((Z.imag * Zp.real).sum())


# --------------------------------------------------------------------------
# blocked row sample of the M*x / velocity operators (CPU baseline at N too large
# for a dense matrix)
# --------------------------------------------------------------------------
def cot_rowsum(Z, x, rows) -> np.ndarray:
    """S_k = sum_{j != k} cot((Z_k - Z_j)/2) x_j for the listed rows k."""
    Z = np.asarray(Z, np.complex128)
    rows = np.asarray(rows)
    d = 0.5 * (Z[rows, None] - Z[None, :])
    d[np.arange(len(rows)), rows] = 1.0
    C = cot(d)
    C[np.arange(len(rows)), rows] = 0.0
    return C @ np.asarray(x, np.complex128)


# --------------------------------------------------------------------------
# optomechanically driven film: the autonomous augmented system [Z | Phi | D]
# (L/OptomechanicalVariables.h, L/LightIntensity.cuh, L/createM.cuh:138-169,
#  L/HeliumDrivenAutonomousProblem.cuh, L/DelayedIntensityIntegrator.cuh,
#  L/AugmentedBoundaryIntegrator.cuh)
# --------------------------------------------------------------------------
HBAR = 1.054571817e-34  # L/constants.cuh:10


@dataclasses.dataclass
class OptomechanicalVariables:
    """L/OptomechanicalVariables.h:3-28."""
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


def light_intensity(height, x, v: OptomechanicalVariables):
    """LightIntensity::compute_intensity / compute_x_profile, L/LightIntensity.cuh:17-29."""
    profile = np.exp(-(x - v.location_x0_mode) ** 2 / (2 * v.sigma_optical_mode ** 2))
    delta_f = v.detuning - v.G * height
    return 0.25 * v.gamma ** 2 * v.max_intensity / (delta_f ** 2 + (v.gamma / 2) ** 2) * profile


def drive_strength(v: OptomechanicalVariables, props: ProblemProperties) -> float:
    """LightIntensity::get_current_intensity_drive_strength, L/LightIntensity.cuh:30-33."""
    return HBAR / (props.base_energy * props.base_time * props.rho) * v.G / v.sigma_optical_mode ** 2


def adimensionalize_optomechanical(v: OptomechanicalVariables, props: ProblemProperties) -> OptomechanicalVariables:
    """adimensionalizeOptomechanicalVariables, L/Export.cu:1250-1275 (props already nondimensionalised)."""
    o = dataclasses.replace(v)
    o.gamma *= props.base_time
    o.detuning *= props.base_time
    o.G *= props.base_time * props.base_length
    o.Tau /= props.base_time
    o.location_x0_mode /= props.base_length
    o.sigma_optical_mode /= props.base_length
    hbar_adim = HBAR / props.base_energy / props.base_time
    o.Beta *= hbar_adim * o.G / o.Tau / (o.sigma_optical_mode ** 2 * props.rho)
    return o


def augmented_rhs(state: np.ndarray, N: int, props: ProblemProperties, v: OptomechanicalVariables, physics: str = "helium",
                  deriv: str = "cuda") -> np.ndarray:
    """AugmentedBoundaryIntegrator::run (L/AugmentedBoundaryIntegrator.cuh:25-29) on [Z | Phi | D] (3N complex):
    the boundary-integral RHS with HeliumDrivenAutonomousProblem::CalculateRhsPhi (base dPhi/dt, then
    add_optical_field_drive_terms_no_time_depence, L/createM.cuh:138-149), then DelayedIntensityIntegrator::run
    (calculate_intensity_delayed_rhs :161-169, add_delayed_intensity_phi_rhs :151-159)."""
    state = np.asarray(state, np.complex128)
    out = np.zeros(3 * N, np.complex128)
    out[:2 * N] = rhs(state[:2 * N], N, 1, props, physics, deriv)
    Z, D, w = state[:N], state[2 * N:], out[:N]
    inten = light_intensity(Z.imag, Z.real, v)
    out[N:2 * N] += v.DampingStrength * w.imag
    out[N:2 * N] += drive_strength(v, props) * inten
    out[2 * N:] = v.Beta * inten - 1.0 / v.Tau * D
    out[N:2 * N] += D
    return out


class TimedDrive:
    """The explicitly time-dependent drive: TimedBoundaryIntegrator<N,1> over HeliumWithOptomechanicalDrivingProblem<N>
    (L/TimedBoundaryIntegrator.cuh:21-26, L/HeliumWithDrivingBoundaryProblem.cuh:42-45, kernel add_optical_field_drive_terms
    L/createM.cuh:119-136) with the exponential integrator of DelayedIntensityTermDevice (L/DelayedIntensityTerm.cuh:16-33), and
    RungeKuttaStepperBase::runStep / runEvolution (L/RK4_Time_Dependent.cuh:145-283, 307-328).  Every point sees the reference
    time from before the launch (the reference's kernel races on *prev_time; this is the intended reading)."""

    def __init__(self, N: int, props: ProblemProperties, v: OptomechanicalVariables, physics: str = "helium", deriv: str = "cuda"):
        self.N, self.props, self.v, self.physics, self.deriv = N, props, v, physics, deriv
        self.delayed = np.zeros(N)
        self.prev_time = v.initial_time
        self.t = v.initial_time

    def set_starting_time(self, time: float):
        self.t = time
        self.prev_time = time

    def rhs(self, state, time: float, save: bool):
        N, v = self.N, self.v
        out = rhs(np.asarray(state, np.complex128), N, 1, self.props, self.physics, self.deriv)
        Z, w = state[:N], out[:N]
        inten = light_intensity(Z.imag, Z.real, v)
        if time == self.prev_time:
            d = inten.copy()
        else:
            a = math.exp(-(time - self.prev_time) / v.Tau)
            d = a * self.delayed + v.Beta * v.Tau * (1 - a) * inten
        if save:
            self.delayed = d
            self.prev_time = time
        out[N:] += v.DampingStrength * w.imag
        out[N:] += v.Beta * d
        out[N:] += drive_strength(v, self.props) * inten
        return out

    def step(self, y, h: float):
        """runStep at the current time (not advanced here, as in the reference)."""
        half = h * 0.5
        k1 = self.rhs(y, self.t, True)
        k2 = self.rhs(y + half * k1, self.t + half, False)
        k3 = self.rhs(y + half * k2, self.t + half, False)
        k4 = self.rhs(y + h * k3, self.t + h, False)
        return y + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)

    def evolve(self, y, t0: float, t1: float, h: float):
        """runEvolution: returns (final state, logged times, logged states); a state is logged with the time at the START of its step."""
        self.set_starting_time(t0)
        steps = int((t1 - t0) / h)
        times, states = [], []
        for _ in range(steps):
            y = self.step(y, h)
            times.append(self.t)
            states.append(y.copy())
            self.t += h
        return y, np.array(times), states


# --------------------------------------------------------------------------
# real-state wrapper, finite-difference Jacobian, implicit Gauss-Legendre-2 integrator
# --------------------------------------------------------------------------
def real_to_complex_state(y: np.ndarray, N: int) -> np.ndarray:
    """convertToComplexStateKernel / createInitialState (L/RealBoundaryIntegralCalculator.cuh:4-12, L/JacobianCalculator.cuh:159-166):
    [x | y | phi] (3N doubles) -> [x + i y | phi + 0 i] (2N complex)."""
    y = np.asarray(y, np.float64)
    out = np.empty(2 * N, np.complex128)
    out[:N] = y[:N] + 1j * y[N:2 * N]
    out[N:] = y[2 * N:3 * N]
    return out


def complex_to_real_rhs(r: np.ndarray, N: int) -> np.ndarray:
    """convertToRealRhsKernel (L/RealBoundaryIntegralCalculator.cuh:14-23): [w | dPhi/dt] -> [Re w | Im w | Re dPhi/dt]."""
    return np.concatenate([r[:N].real, r[:N].imag, r[N:2 * N].real])


def real_rhs(y: np.ndarray, N: int, props: ProblemProperties, physics: str = "helium", deriv: str = "cuda") -> np.ndarray:
    """RealBoundaryItegralCalculator<N>::run (L/RealBoundaryIntegralCalculator.cuh:59-70)."""
    return complex_to_real_rhs(rhs(real_to_complex_state(y, N), N, 1, props, physics, deriv), N)


def perturbed_states(state: np.ndarray, N: int, eps: float) -> np.ndarray:
    """createInitialBatchedZ (L/JacobianCalculator.cuh:11-77): 3N copies of the complex state [Z | Phi] in the batched layout
    [Z of member 0 .. Z of member 3N-1 | Phi of member 0 ..]; member b = c N + j has coordinate c (0: x, 1: y, 2: phi) of
    point j moved by eps."""
    state = np.asarray(state, np.complex128)
    B = 3 * N
    out = np.empty(2 * B * N, np.complex128)
    Zb = out[:B * N].reshape(B, N)
    Pb = out[B * N:].reshape(B, N)
    Zb[:] = state[:N]
    Pb[:] = state[N:]
    j = np.arange(N)
    Zb[j, j] += eps
    Zb[N + j, j] += 1j * eps
    Pb[2 * N + j, j] += eps
    return out


def jacobian_fd(y: np.ndarray, N: int, props: ProblemProperties, physics: str = "helium", eps: float = 1e-6, deriv: str = "cuda") -> np.ndarray:
    """JacobianCalculator<N>::calculateJacobian (L/JacobianCalculator.cuh:226-284) with createJacobianMatrixFromPerturbedRhs
    (:92-156): central differences of the batched RHS (batch = 3N) at +-eps.  Returns J[r, c] = d f_r / d y_c as a (3N, 3N) array;
    the reference's buffer is its column-major flattening, `J.ravel(order="F")` (C[c * 3N + r])."""
    s = real_to_complex_state(y, N)
    B = 3 * N
    pos = rhs(perturbed_states(s, N, eps), N, B, props, physics, deriv)
    neg = rhs(perturbed_states(s, N, -eps), N, B, props, physics, deriv)
    return jacobian_from_perturbed(pos, neg, N, eps)


def jacobian_from_perturbed(pos: np.ndarray, neg: np.ndarray, N: int, eps: float) -> np.ndarray:
    """createJacobianMatrixFromPerturbedRhs (L/JacobianCalculator.cuh:92-156): pos / neg are the batched RHS [w of member 0 .. w of
    member 3N-1 | dPhi/dt of member 0 ..] at +eps / -eps; returns J[r, c] (the reference's buffer is J.ravel(order="F"))."""
    B = 3 * N
    diff = np.asarray(pos) - np.asarray(neg)
    d = diff.real / (2.0 * eps) + 1j * (diff.imag / (2.0 * eps))   # complex / real is component-wise in libcu++ (numpy's is not, to the ulp)
    J = np.empty((3 * N, 3 * N))
    w = d[:B * N].reshape(B, N)       # member c: velocity rows
    p = d[B * N:].reshape(B, N)       # member c: dPhi/dt rows
    J[:N, :] = w.real.T
    J[N:2 * N, :] = w.imag.T
    J[2 * N:, :] = p.real.T
    return J


SQRT3 = 1.7320508075688772935  # L/GLCoefficients.hpp:3
GL_A = ((0.25, 0.25 - SQRT3 / 6.0), (0.25 + SQRT3 / 6.0, 0.25))   # L/GLCoefficients.hpp:7-10
GL_B = (0.5, 0.5)


@dataclasses.dataclass
class GaussLegendre2Options:
    """GaussLegendre2Options (L/GaussLegendre.cuh:70-90; its constructor sets allowSimplifiedFallback = false) and the C struct
    GaussLegendreOptions (L/ExportTypes.cuh:20-33)."""
    stepSize: float = 0.01
    newtonTolerance: float = 1e-10
    maxNewtonIterations: int = 20
    allowSimplifiedFallback: bool = False
    returnTrajectory: bool = True
    armijo_c: float = 1e-4
    backtrack: float = 0.5
    minAlpha: float = 1e-6
    maxStepsHalves: int = 6


def gl2_newton_matrix(J1: np.ndarray, J2: np.ndarray, h: float) -> np.ndarray:
    """Jacobian of the stage residual R_i = k_i - f(y + h sum_j a_ij k_j): block (i, j) = delta_ij I - h a_ij J_i
    (P/integration/gauss_legendre.py:40-52).  The CUDA fillMMatrix (L/GaussLegendre.cuh:42-51) stores, read column-major as its
    LU does, the blocks (1,2) = -h a21 J1 and (2,1) = -h a12 J2, i.e. with the two off-diagonal coefficients exchanged (an inexact
    Newton matrix; the fixed point is the same) and is launched with grid and block exchanged (:483), which fails for 3N/16 > 32;
    the Python statement is followed here."""
    m = J1.shape[0]
    I = np.eye(m)
    M = np.empty((2 * m, 2 * m))
    M[:m, :m] = I - h * GL_A[0][0] * J1
    M[:m, m:] = -h * GL_A[0][1] * J1
    M[m:, :m] = -h * GL_A[1][0] * J2
    M[m:, m:] = I - h * GL_A[1][1] * J2
    return M


def gl2_step(f, J, y: np.ndarray, h: float, opt: GaussLegendre2Options):
    """GaussLegendre2<N>::gaussLegendreS2Step (L/GaussLegendre.cuh:441-560) == gauss_legendre_s2_step
    (P/integration/gauss_legendre.py:55-170): damped Newton on the two stage slopes with Armijo backtracking on
    phi = |R|^2 / 2 and an optional switch to a Jacobian frozen at the base point.  Where the two statements differ the Python one
    (the intended algorithm) is followed: residual and its norm over BOTH stages (the CUDA cublasDdot runs over 3N of the 6N
    entries, :606), tolerance relative to 1 + |k| (CUDA: 1 + |k|^2, :607/:469), stage states from the slopes being tested (the CUDA
    stageStates reads the member buffers k1/k2 whatever slopes residualAndPhi was given, :580-586).  Returns
    (y_next or None, info)."""
    y = np.asarray(y, np.float64)
    m = y.size
    fy = f(y)
    k = np.stack([fy, fy.copy()])

    def residual_and_phi(k_):
        y1 = y + h * (GL_A[0][0] * k_[0] + GL_A[0][1] * k_[1])
        y2 = y + h * (GL_A[1][0] * k_[0] + GL_A[1][1] * k_[1])
        R = np.concatenate([k_[0] - f(y1), k_[1] - f(y2)])
        return R, 0.5 * float(R @ R), y1, y2

    simplified, converged, freeze, Jf = False, False, False, None
    it = 0
    res_norm = float("nan")
    for it in range(opt.maxNewtonIterations):
        R, phi, y1, y2 = residual_and_phi(k)
        res_norm = math.sqrt(2.0 * phi)
        if res_norm <= opt.newtonTolerance * (1.0 + float(np.linalg.norm(k))):
            converged = True
            break
        J1, J2 = (Jf, Jf) if freeze else (J(y1), J(y2))
        dK = np.linalg.solve(gl2_newton_matrix(J1, J2, h), -R)
        alpha = 1.0
        target = phi - opt.armijo_c * alpha * res_norm ** 2
        while True:
            kt = np.stack([k[0] + alpha * dK[:m], k[1] + alpha * dK[m:]])
            _, phit, _, _ = residual_and_phi(kt)
            if phit <= target:
                k = kt
                break
            alpha *= opt.backtrack
            target = phi - opt.armijo_c * alpha * res_norm ** 2
            if alpha < opt.minAlpha:
                if not freeze and opt.allowSimplifiedFallback:
                    freeze, simplified = True, True
                    Jf = J(y)
                    break
                return None, dict(nit=it, converged=False, res_norm=res_norm, simplified_used=simplified)
    else:
        it = opt.maxNewtonIterations
    if not converged:
        R, phi, _, _ = residual_and_phi(k)
        res_norm = math.sqrt(2.0 * phi)
    y_next = y + h * (GL_B[0] * k[0] + GL_B[1] * k[1])
    return y_next, dict(nit=it, converged=converged, res_norm=res_norm, simplified_used=simplified)


def gl2_integrate(f, J, y0: np.ndarray, t0: float, t1: float, opt: GaussLegendre2Options):
    """GaussLegendre2<N>::runEvolution (L/GaussLegendre.cuh:216-299) == integrate_gl2 (P/integration/gauss_legendre.py:173-267): steps of
    min(stepSize, |t1 - t|) in the direction of t1; a step whose Newton iteration fails is retried with half the size, at most
    maxStepsHalves times and not below |t1 - t0| / 2^20, and the reduced size is kept for the following steps.  The trajectory
    starts with (t0, y0).  Returns (times, states) -- with returnTrajectory off, ([], [final state])."""
    if not opt.stepSize > 0.0:
        raise ValueError("Step size must be positive.")
    y = np.asarray(y0, np.float64).copy()
    total = abs(t1 - t0)
    hmin = total / 2.0 ** 20
    h = opt.stepSize
    T, Y = [float(t0)], [y.copy()]
    t = float(t0)
    forward = 1.0 if t1 >= t0 else -1.0
    while (t - t1) * forward < 0.0:
        htry = min(h, abs(t1 - t)) * forward
        ok, info = False, None
        for _ in range(opt.maxStepsHalves + 1):
            y_next, info = gl2_step(f, J, y, htry, opt)
            if y_next is not None and info["converged"]:
                ok = True
                break
            if abs(htry) <= hmin:
                break
            htry *= 0.5
        if not ok:
            raise RuntimeError(f"Gauss-Legendre 2nd Order method failed to converge at t~{t}; residual={info['res_norm']:.3e}; last htry={htry:.3e}")
        y = y_next
        t += htry
        h = abs(htry)
        T.append(t)
        Y.append(y.copy())
    if not opt.returnTrajectory:
        return np.zeros(0), np.array([y])
    return np.array(T), np.array(Y)
