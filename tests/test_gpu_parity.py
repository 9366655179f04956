"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle and the
reference-generated golden vectors.  Tolerances are stated per test; floating point only (FP64).

Noise floor: the FFT derivatives amplify round-off by ~N*eps (mode number times coefficient noise), in the reference as much as
here, so RHS-level agreement between two correct FP64 implementations is ~1e-16*N relative, not 1e-16."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import roberts_oracle as ro  # noqa: E402


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from superfluid_dynamics_b200 import api, build
    build.build(verbose=False)
    return api


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def colmajor(t, N):
    return t.cpu().numpy().reshape(N, N).T


# ---- the reference's own kernel tests, same fixtures and tolerances (T/MatrixMTests.cuh) --------------------------
def test_two_by_two_M_matrix(api):
    """Kernels.TwoByTwoMMatrix :160-223, tol 1e-14."""
    N, h = 2, 0.5
    Z, _ = ro.trochoid(N, h, 10.0, 0.0)
    Zp, Zpp, _ = ro.trochoid_derivatives(N, h, 10.0, 0.0)
    A = torch.empty(N * N, dtype=torch.float64, device="cuda:0")
    api.createMKernel(A, T(Z), T(Zp), T(Zpp), 0.0, N, 1)
    th = math.sinh(2 * h) / (math.cosh(2 * h) + 1.0)
    exp = np.array([[0.5 - 0.25 * h / (1.0 - h), (h - 1) / 4.0 * th], [(h + 1) / 4.0 * th, 0.5 + 0.25 * h / (1.0 + h)]])
    assert np.abs(colmajor(A, N) - exp).max() <= 1e-14


@pytest.mark.parametrize("N,t", [(4, 0.1), (8, 0.1), (64, 0.0), (300, 0.2)])
def test_M_and_velocity_matrices(api, N, t):
    """Kernels.MMatrixKernel :225-282 (1e-14) and Kernels.Velocities :396-468 (1e-12), plus larger / ragged N."""
    h = 0.5
    Z, _ = ro.trochoid(N, h, 10.0, t)
    Zp, Zpp, _ = ro.trochoid_derivatives(N, h, 10.0, t)
    A = torch.empty(N * N, dtype=torch.float64, device="cuda:0")
    api.createMKernel(A, T(Z), T(Zp), T(Zpp), 0.0, N, 1)
    assert np.abs(colmajor(A, N) - ro.create_M(Z, Zp, Zpp, 0.0)).max() <= 1e-14
    api.createMKernel(A, T(Z), T(Zp), T(Zpp), 0.3, N, 1)
    assert np.abs(colmajor(A, N) - ro.create_M(Z, Zp, Zpp, 0.3)).max() <= 1e-14
    V1 = torch.empty(N * N, dtype=torch.complex128, device="cuda:0")
    V2 = torch.empty(N, dtype=torch.complex128, device="cuda:0")
    for lower in (True, False):
        api.createVelocityMatrices(T(Z), T(Zp), T(Zpp), N, V1, V2, lower, 1)
        e1, e2 = ro.velocity_matrices(Z, Zp, Zpp, lower)
        assert np.abs(colmajor(V1, N) - e1).max() <= 1e-12 * max(1.0, np.abs(e1).max())
        assert np.abs(V2.cpu().numpy() - e2).max() <= 1e-12 * max(1.0, np.abs(e2).max())


def test_finite_depth_kernels(api):
    """createFiniteDepthMKernel / createHeliumVelocityMatrices: no reference test exists; vs the oracle here, and vs the reference's
    own CUDA path (whole RHS, finite depth) in tests/test_gpu_reference.py."""
    N, depth = 96, 0.6
    Z, _ = ro.trochoid(N, 0.2)
    Zp, Zpp, _ = ro.trochoid_derivatives(N, 0.2)
    A = torch.empty(N * N, dtype=torch.float64, device="cuda:0")
    V1 = torch.empty(N * N, dtype=torch.complex128, device="cuda:0")
    V2 = torch.empty(N, dtype=torch.complex128, device="cuda:0")
    for inf in (False, True):
        api.createFiniteDepthMKernel(A, T(Z), T(Zp), T(Zpp), depth, N, 1, inf)
        e = ro.create_finite_depth_M(Z, Zp, Zpp, depth, inf)
        assert np.abs(colmajor(A, N) - e).max() <= 1e-13 * np.abs(e).max()
        for lower in (True, False):
            api.createHeliumVelocityMatrices(T(Z), T(Zp), T(Zpp), depth, N, V1, V2, lower, 1, inf)
            e1, e2 = ro.helium_velocity_matrices(Z, Zp, Zpp, depth, lower, inf)
            assert rel(colmajor(V1, N), e1) <= 1e-13 and rel(V2.cpu().numpy(), e2) <= 1e-13


def test_batched_matrix_kernels(api):
    """grid.z = batch layout A[k + j*n + b*n*n] (L/createM.cuh:47-52)."""
    N, B = 40, 3
    Zs, Zps, Zpps = [], [], []
    for b in range(B):
        Z, _ = ro.trochoid(N, 0.1 + 0.1 * b)
        Zp, Zpp, _ = ro.trochoid_derivatives(N, 0.1 + 0.1 * b)
        Zs.append(Z); Zps.append(Zp); Zpps.append(Zpp)
    A = torch.empty(B * N * N, dtype=torch.float64, device="cuda:0")
    api.createMKernel(A, T(np.concatenate(Zs)), T(np.concatenate(Zps)), T(np.concatenate(Zpps)), 0.0, N, B)
    Ah = A.cpu().numpy().reshape(B, N, N)
    for b in range(B):
        assert np.abs(Ah[b].T - ro.create_M(Zs[b], Zps[b], Zpps[b], 0.0)).max() <= 1e-14


def test_zphi_derivatives_analytic(api):
    """Kernels.ZPhiDerivatives :477-607: N = 1024, h = 0.5, omega = 10 vs analytic derivatives, tol 1e-14."""
    N, h, omega = 1024, 0.5, 10.0
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    Z, Phi = ro.trochoid(N, h, omega, 0.0)
    eZp, eZpp, ePhiP = ro.trochoid_derivatives(N, h, omega, 0.0)
    Zp, PhiP, Zpp = calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
    assert np.abs(Zp.cpu().numpy() - eZp).max() <= 1e-14
    assert np.abs(Zpp.cpu().numpy() - eZpp).max() <= 1e-14
    assert np.abs(PhiP.cpu().numpy().real - ePhiP).max() <= 1e-14


@pytest.mark.parametrize("batch", [1, 2])
@pytest.mark.parametrize("N", [128, 256, 512, 1024, 2048, 4096, 8192])
def test_fused_shared_memory_transforms_on_rough_data(api, N, batch):
    """ZPhiDerivative::exec (L/Derivatives.cuh:311-384) through the one-CTA fused transforms of spectral.cu -- radix-2 below N = 256,
    the radix-8 register-resident transform (all three first-pass radices: log2 N mod 3 = 0, 1, 2) up to 8192 -- on data with a
    full spectrum, so every mode, the pi factor at N/2 and the zeroed mode N/2+1 count; against the oracle's numpy transforms.
    The real derivative a' = (2 pi / N) D1(a) of the same kernels is checked through the whole RHS below and in test_rhs_*."""
    rng = np.random.default_rng(N + batch)
    a = 2 * np.pi * np.arange(N) / N
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, batch, props, api.WaterBoundaryProblem(props))
    Zs = [a + 0.01 * (rng.standard_normal(N) + 1j * rng.standard_normal(N)) for _ in range(batch)]
    Phis = [(0.1 * rng.standard_normal(N)).astype(np.complex128) for _ in range(batch)]
    Zp, PhiP, Zpp = calc.zPhiDerivative(T(np.concatenate(Zs)), T(np.concatenate(Phis)))
    Zp, PhiP, Zpp = (v.cpu().numpy().reshape(batch, N) for v in (Zp, PhiP, Zpp))
    for b in range(batch):
        eZp, ePhiP, eZpp = ro.zphi_derivative(Zs[b], Phis[b], ro.ProblemProperties(rho=0.0))
        for got, exp in ((Zp[b], eZp), (PhiP[b], ePhiP), (Zpp[b], eZpp)):
            assert np.abs(got - exp).max() <= 2e-13 * max(1.0, np.abs(exp).max())


@pytest.mark.parametrize("N", [2048, 4096, 8192])
def test_rhs_with_fused_transforms_matches_the_library_transforms(api, N):
    """The whole RHS (surface derivatives and the a' of the solve) with the fused radix-8 transforms (default up to N = 8192)
    against the same solver on cuFFT (RB_OWN_FFT=0): identical to round-off.  Round-off here grows like N: the velocity contains
    V2 a' with a' the spectral derivative of the solved a, which multiplies the 1e-15 noise of a by wavenumbers up to N/2 (measured
    own-against-library: 8e-14 at N = 1024, 2.2e-13 at 2048, 5.2e-13 at 4096, 7.9e-13 at 8192, a itself to 1.5e-15)."""
    props = api.ProblemProperties(rho=0.0)
    Z, Phi = ro.trochoid(N, 0.3)
    y = T(ro.pack_state(Z, Phi))
    outs = []
    for own in ("1", "0"):
        os.environ["RB_OWN_FFT"] = own
        try:
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
        finally:
            os.environ.pop("RB_OWN_FFT", None)
        out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
        calc.run(y, out)
        assert calc.solve_stats()["converged"]
        outs.append(out.cpu().numpy())
    assert np.abs(outs[0] - outs[1]).max() <= 1e-15 * N * np.abs(outs[1]).max()


@pytest.mark.parametrize("N", [2, 4, 8, 64, 1000])
def test_derivative_nyquist_quirks(api, N):
    """Nyquist-rich data: the CUDA conventions (pi factor, zeroed mode N/2+1) must be reproduced exactly."""
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    d1 = calc.fftDerivative(T(x), False, 0.7).cpu().numpy()
    d2 = calc.fftDerivative(T(x), True, 1.0).cpu().numpy()
    e1, e2 = ro.fft_derivative(x, 0.7), ro.fft_derivative(x, 1.0, second=True)
    assert np.abs(d1 - e1).max() <= 1e-13 * max(1.0, np.abs(e1).max())
    assert np.abs(d2 - e2).max() <= 1e-13 * max(1.0, np.abs(e2).max())


def test_rhs_phi_kernels(api):
    """Kernels.RhsPhi :748-819 (1e-14) and the helium variants vs the oracle."""
    N, omega, h, t = 32, 10.0, 0.5, 0.1
    a = 2 * np.pi * np.arange(N) / N
    i = np.arange(N, dtype=np.float64)
    Yv = h * np.cos(i - omega * t)
    Zh = (a - h * np.sin(a - omega * t)) + 1j * Yv
    VL = 2 * np.pi / N * ((1 - h * np.cos(a - omega * t)) + 1j * (-h * np.sin(a - omega * t)))
    VU = 0.3 * VL[::-1].copy()
    out = torch.empty(N, dtype=torch.complex128, device="cuda:0")
    api.compute_rhs_phi_expression(T(Zh), T(VL), T(np.zeros(N, complex)), out, 0.0, N)
    exp = -Yv + 0.5 * (VL.real ** 2 + VL.imag ** 2)
    o = out.cpu().numpy()
    assert np.abs(o.real - exp).max() <= 1e-14 and np.abs(o.imag).max() == 0.0
    api.compute_rhs_phi_expression(T(Zh), T(VL), T(VU), out, 0.25, N)   # rho != 0 exercises the V1[1] quirk
    assert np.abs(out.cpu().numpy().real - ro.rhs_phi_water(Zh, VL, VU, 0.25)).max() <= 1e-14
    Zp, Zpp, _ = ro.trochoid_derivatives(N, h, omega, t)
    Zs = Zh.real + 1j * 0.1 * Yv
    api.compute_rhs_helium_phi_expression(T(Zs), T(VL), out, 0.8, N)
    assert np.abs(out.cpu().numpy().real - ro.rhs_phi_helium(Zs, VL, 0.8)).max() <= 1e-13
    api.compute_rhs_helium_phi_expression_with_surface_tension(T(Zs), T(Zp), T(Zpp), T(VL), out, 0.8, 0.05, N)
    e = ro.rhs_phi_helium_surface_tension(Zs, Zp, Zpp, VL, 0.8, 0.05)
    assert np.abs(out.cpu().numpy().real - e).max() <= 1e-13 * max(1.0, np.abs(e).max())
    for order in (1, 2, 3):
        api.compute_rhs_helium_phi_expression_expansion_terms(T(Zs), T(VL), out, 0.8, N, order)
        assert np.abs(out.cpu().numpy().real - ro.rhs_phi_helium_expansion(Zs, VL, 0.8, order)).max() <= 1e-14


# ---- the matrix-free core ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,h", [(2, 0.3), (3, 0.2), (64, 0.3), (255, 0.3), (256, 0.4), (257, 0.4), (1000, 0.4), (1024, 0.4),
                                 (1300, 0.2), (4096, 0.4)])
def test_cotangent_sum_matches_direct_evaluation(api, N, h):
    """S_k = sum_{j != k} cot((z_k - z_j)/2) x_j: exponential / cell-local form vs direct 1/tan evaluation.
    tol 2e-13 relative to max|S| (near-neighbour entries are ~N/pi, so this is ~1e-14 relative per entry)."""
    Z, Phi = ro.trochoid(N, h)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
    x = np.cos(3 * 2 * np.pi * np.arange(N) / N) + 0.3 * np.sin(2 * np.pi * np.arange(N) / N) + 0.1
    S = calc.cotangentSum(T(Z), T(x)).cpu().numpy()
    rows = np.arange(N) if N <= 1300 else np.r_[0:6, N - 6:N, N // 2 - 3:N // 2 + 3, 253:259, 509:515, 1021:1027]
    assert not np.isnan(S).any()
    assert rel(S[rows], ro.cot_rowsum(Z, x, rows)) <= 2e-13


@pytest.mark.parametrize("N,h", [(2, 0.3), (64, 0.3), (300, 0.3), (1000, 0.4), (1024, 0.4), (1300, 0.2), (1536, 0.4), (2048, 0.4), (3000, 0.3),
                                 (3072, 0.4), (4096, 0.4), (5000, 0.3), (8192, 0.4)])
def test_warp_per_row_group_sweep_matches_direct_evaluation_and_the_tiled_kernel(api, N, h):
    """sweep3_kernel (pair_kernels3.cu; default for 1024 < N <= 8192, forced here at every size): 1, 2 and 4 rows per warp, one and
    two staged tiles, partial last cell, few-cell surfaces without cell-local coordinates.  (a) the raw cotangent sum against the
    direct 1/tan evaluation of the oracle, as test_cotangent_sum_matches_direct_evaluation; (b) the whole RHS (solver sweeps,
    combined velocity sweep, convergence decision) against the same solver on the tiled kernel."""
    Z, Phi = ro.trochoid(N, h)
    props = api.ProblemProperties(rho=0.0)
    x = np.cos(3 * 2 * np.pi * np.arange(N) / N) + 0.3 * np.sin(2 * np.pi * np.arange(N) / N) + 0.1
    y = T(ro.pack_state(Z, Phi))
    outs, plans = [], []
    for v3 in ("1", "0"):
        os.environ["RB_SWEEP_V3"] = v3
        os.environ["RB_SWEEP_V2"] = "0"
        try:
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
        finally:
            os.environ.pop("RB_SWEEP_V3", None)
            os.environ.pop("RB_SWEEP_V2", None)
        plans.append(calc.sweepPlan())
        if v3 == "1":
            calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
            S = calc.cotangentSum(T(Z), T(x)).cpu().numpy()
            rows = np.arange(N) if N <= 1300 else np.r_[0:6, N - 6:N, N // 2 - 3:N // 2 + 3, 253:259, 509:515, 1021:1027]
            assert not np.isnan(S).any()
            assert rel(S[rows], ro.cot_rowsum(Z, x, rows)) <= 2e-13
        out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
        calc.run(y, out)
        assert calc.solve_stats()["converged"]
        outs.append(out.cpu().numpy())
    assert plans[0]["kernel"] == "warp_rows" and plans[1]["kernel"] == "tiled", plans
    assert rel(outs[0], outs[1]) <= 1e-12


@pytest.mark.parametrize("N,B", [(64, 7), (300, 5), (512, 3), (768, 2), (512, 160)])
def test_ensemble_sweep_one_member_per_cta_matches_the_members_stepped_alone(api, N, B):
    """sweep3b_kernel (pair_kernels3.cu; default for ensembles of >= 64 members with 256 < N <= 768, forced here): every member's
    RHS from the batched solver against the same member through a solver of its own (batch 1), ragged N, fewer and more members
    than SMs; and per-member convergence (every member's own residual)."""
    hs = 0.05 + 0.3 * np.arange(B) / max(B - 1, 1)
    members = [ro.pack_state(*ro.trochoid(N, h)) for h in hs]
    props = api.ProblemProperties(rho=0.0)
    os.environ["RB_SWEEP_V3B"] = "1"
    try:
        calc = api.BaseBoundaryIntegralCalculator(N, B, props, api.WaterBoundaryProblem(props))
    finally:
        os.environ.pop("RB_SWEEP_V3B", None)
    assert calc.sweepPlan()["kernel"] == "warp_rows"
    y = T(api.ensemble_state(members, N))
    out = torch.zeros(2 * N * B, dtype=torch.complex128, device="cuda:0")
    calc.run(y, out)
    assert calc.solve_stats()["converged"]
    got = out.cpu().numpy()
    alone = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    o1 = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
    for b in sorted({0, B // 2, B - 1}):
        alone.run(T(members[b]), o1)
        e = o1.cpu().numpy()
        mine = np.concatenate([got[b * N:(b + 1) * N], got[B * N + b * N:B * N + (b + 1) * N]])
        assert rel(mine, e) <= 1e-12, b


def test_cotangent_sum_is_linear_and_row_local_at_full_size(api):
    """Size-independent properties at N = 65536 (BASELINE config 5): linearity in x, and a spot check of rows against the
    direct evaluation (the dense oracle does not fit in host memory at this size)."""
    N = 65536
    Z, Phi = ro.trochoid(N, 0.4)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
    al = 2 * np.pi * np.arange(N) / N
    x1, x2 = np.cos(5 * al) + 0.2, np.sin(3 * al) * np.cos(al)
    S1 = calc.cotangentSum(T(Z), T(x1)).cpu().numpy()
    S2 = calc.cotangentSum(T(Z), T(x2)).cpu().numpy()
    S3 = calc.cotangentSum(T(Z), T(2.0 * x1 - 3.0 * x2)).cpu().numpy()
    assert rel(S3, 2.0 * S1 - 3.0 * S2) <= 1e-12
    rows = np.r_[0:3, N - 3:N, 32766:32770, 255:258, 20000:20002]
    # wrap-around rows carry the reference's own ~eps*N/(2 pi) cancellation error in fl(x_k - x_j); hence 5e-11
    assert rel(S1[rows], ro.cot_rowsum(Z, x1, rows)) <= 5e-11
    inner = np.r_[32766:32770, 20000:20002]
    assert rel(S1[inner], ro.cot_rowsum(Z, x1, inner)) <= 1e-12


def _spot_rows(N):
    """>= 32 scattered rows: cell boundaries (every 256), the wrap-around ends, the middle, and pseudo-random interior rows."""
    rng = np.random.default_rng(7)
    rows = np.r_[0:3, N - 3:N, 254:258, 510:514, N // 2 - 2:N // 2 + 2, N // 3, 2 * N // 3 + 1, rng.integers(0, N, 14)]
    return np.unique(rows)


@pytest.mark.parametrize("N,h,physics", [(65536, 0.4, "water"), (49152, 0.3, "water"), (16384, 0.4, "water")])
def test_full_rhs_rows_at_the_headline_size_match_direct_evaluation(api, N, h, physics):
    """The kernels the headline times (tiled sweep, 4 rows per thread at N >= 16384: MV epilogue of the solve, VEL epilogue of the
    velocities and dPhi/dt) against an independent formula AT the metric's own size, where no dense oracle fits: after rb_rhs, on
    >= 32 scattered rows k,
      (1) the solve: | b_k - (M a)_k | / max|b| <= 1e-11, with (M a)_k = Mdiag_k a_k + (1/4pi) Im(Zp_k S_k) (L/createM.cuh:43-63),
      (2) the velocity: conj[ (-i/4pi) S_k + V1diag_k a_k + V2_k a'_k ] (L/WaterVelocities.cuh:38-70, 217-241),
      (3) dPhi/dt = -Y + |w|^2 / 2 (L/createM.cuh:96-107 at rho = 0),
    S_k = sum_{j != k} cot((z_k - z_j)/2) a_j by the oracle's direct 1/tan evaluation (ro.cot_rowsum) of the GPU's own a.
    Zp, Zpp, b are the GPU's (the derivatives carry N eps .. N^2 eps / 4 of round-off noise in ANY implementation: they are pinned
    separately at sizes the reference runs); a' = (2 pi / N) D1 a by the oracle's FFT derivative.  Tolerances, relative to the max
    of the compared array: 1e-11 for (1); 5e-11 for (2) and (3) (the direct 1/tan itself loses eps N / 2 pi = 2e-12 on wrap-around
    pairs, the FFT derivative of a another N eps)."""
    Z, Phi = ro.trochoid(N, h)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    plan = calc.sweepPlan()
    assert plan["kernel"] == "tiled" and plan["rows_per_thread"] == 4, plan    # the kernel the headline times
    out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
    calc.run(T(ro.pack_state(Z, Phi)), out)
    st = calc.solve_stats()
    assert st["converged"] and st["residual"] <= 1e-13, st
    o = out.cpu().numpy()
    a = calc.getDevA().cpu().numpy()
    Zp, Zpp, b = calc.getDevZp().cpu().numpy(), calc.getDevZpp().cpu().numpy(), calc.devPhiPrime.cpu().numpy()
    rows = _spot_rows(N)
    assert len(rows) >= 32
    S = ro.cot_rowsum(Z, a, rows)
    inv4pi = 0.25 / np.pi
    Ma = (0.5 + inv4pi * (Zpp[rows] / Zp[rows]).imag) * a[rows] + inv4pi * (Zp[rows] * S).imag
    res = np.abs(b[rows] - Ma).max() / np.abs(b).max()
    assert res <= 1e-11, res
    ap = ro.fft_derivative(a.astype(np.complex128), 2.0 * np.pi / N)
    w = -1j * inv4pi * S + (-1j * inv4pi * Zpp[rows] / Zp[rows] ** 2 + 0.5 / Zp[rows]) * a[rows] + 1j / (2.0 * np.pi * Zp[rows]) * ap[rows]
    ev = np.abs(o[rows] - np.conj(w)).max() / np.abs(o[:N]).max()
    assert ev <= 5e-11, ev
    dphi = -Z.imag[rows] + 0.5 * np.abs(w) ** 2
    ed = np.abs(o[N + rows] - dphi).max() / np.abs(o[N:]).max()
    assert ed <= 5e-11, ed
    print(f"N={N}: rows {len(rows)}, solve residual {res:.2e}, velocity {ev:.2e}, dPhi/dt {ed:.2e}, {st}")


@pytest.mark.parametrize("N,dt", [(65536, 1e-4), (4096, 1e-3)])
def test_recorded_steps_equal_unrecorded_steps_at_the_headline_size(api, N, dt):
    """The kernels of a RECORDED step (solver sweep + combined verify-and-velocity sweep with the RK update folded into
    finish_solve, stage-history start) against the unrecorded sequence (solver sweeps to convergence, then the velocity sweep the
    test above pins row by row, separate update kernels) at the sizes the metric is quoted on: both integrate the same system with
    every solve verified to 1e-13, so 8 steps must agree to ~1e-12 (tolerance 1e-11 relative, position and potential)."""
    props = api.ProblemProperties(rho=0.0)
    y0 = ro.pack_state(*ro.trochoid(N, 0.4))
    states = []
    for no_graph in ("0", "1"):
        os.environ["RB_NO_GRAPH"] = no_graph
        try:
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
        finally:
            os.environ.pop("RB_NO_GRAPH", None)
        y = T(y0)
        stp.initialize(y, True)
        stp.runSteps(8)
        st, ss = calc.solve_stats(), stp.stats()
        assert st["converged"] and st["failed_solves"] == 0, st
        assert (ss["graph_launches"] >= 1) == (no_graph == "0"), ss
        states.append(y.cpu().numpy())
    assert rel(states[0][:N], states[1][:N]) <= 1e-11
    assert rel(states[0][N:], states[1][N:]) <= 1e-11


def test_tight_recording_drops_the_surplus_round_and_keeps_the_trajectory(api):
    """After a run of steps that all needed the same number of sweeps the stepper records exactly that many (no self-skipping
    surplus round); a step that then ran out of sweeps would be rolled back and redone.  The trajectory is the one of the ordinary
    recording (RB_TIGHT_GRAPH=0) to round-off: both verify every solve to 1e-13."""
    N, dt, steps = 1024, 1e-3, 80
    props = api.ProblemProperties(rho=0.0)
    y0 = ro.pack_state(*ro.trochoid(N, 0.4))
    out = {}
    for tight in ("1", "0"):
        os.environ["RB_TIGHT_GRAPH"] = tight
        try:
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
        finally:
            os.environ.pop("RB_TIGHT_GRAPH", None)
        y = T(y0)
        stp.initialize(y, True)
        stp.runSteps(steps)
        ss, st = stp.stats(), calc.solve_stats()
        assert st["converged"] and st["failed_solves"] == 0, st
        assert ss["tight"] == (tight == "1"), ss
        if tight == "1":
            assert ss["graph_sweeps"] == 2 and ss["tight_failures"] == 0, ss
        out[tight] = y.cpu().numpy()
    assert rel(out["1"], out["0"]) <= 1e-12


# ---- full RHS -------------------------------------------------------------------------------------------------------
def _split(state):
    N = len(state) // 3
    return state[:N] + 1j * state[N:2 * N], state[2 * N:]


@pytest.mark.parametrize("tag", ["N64_h0.1", "N64_h0.4", "N32_h0.25", "N128_h0.3", "N64_pert"])
@pytest.mark.parametrize("mode", ["matrix_free", "dense_lu"])
def test_rhs_matches_reference_golden(api, golden, tag, mode):
    """Outputs of the reference's P/WaterIntegralCalculator.py (tests/golden/ref_water_rhs.npz), tol 2e-12."""
    Z, Phi = _split(golden[tag + "_state"])
    N = len(Z)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), solve_mode=mode)
    out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
    calc.run(T(ro.pack_state(Z, Phi)), out)
    o = out.cpu().numpy()
    ref = golden[tag + "_rhs"]
    mine = np.hstack((o[:N].real, o[:N].imag, o[N:].real))
    assert np.abs(mine - ref).max() <= 2e-12 * max(1.0, np.abs(ref).max())
    assert np.abs(o[N:].imag).max() == 0.0
    assert np.abs(calc.getDevA().cpu().numpy() - golden[tag + "_a"]).max() <= 2e-12
    if mode == "matrix_free":
        assert calc.solve_stats()["converged"]


@pytest.mark.parametrize("physics,N,h,B,rho,depth", [
    ("water", 2, 0.2, 1, 0.0, 1.0), ("water", 6, 0.2, 1, 0.0, 1.0), ("water", 256, 0.4, 1, 0.0, 1.0),
    ("water", 1000, 0.3, 1, 0.0, 1.0), ("water", 1024, 0.4, 1, 0.0, 1.0), ("water", 4096, 0.4, 1, 0.0, 1.0),
    ("water", 128, 0.3, 3, 0.0, 1.0), ("water", 512, 0.2, 5, 0.0, 1.0), ("water", 256, 0.3, 1, 0.2, 1.0),
    ("helium_inf", 256, 0.05, 1, 0.0, 0.3), ("helium", 256, 0.01, 1, 0.0, 0.3), ("helium", 128, 0.02, 2, 0.0, 0.3),
    ("helium", 1024, 0.00942478, 1, 0.0, 0.0942478)])
def test_rhs_matches_oracle(api, physics, N, h, B, rho, depth):
    """Full RHS vs the oracle (CUDA derivative semantics): tol 1e-15*N + 1e-13 relative (FFT-derivative noise floor)."""
    props = api.ProblemProperties(rho=rho, depth=depth)
    oprops = ro.ProblemProperties(rho=rho, depth=depth)
    prob = {"water": api.WaterBoundaryProblem, "helium": api.HeliumBoundaryProblem,
            "helium_inf": api.HeliumInfiniteDepthBoundaryProblem}[physics](props)
    states = [ro.trochoid(N, h * (1 + 0.3 * b)) for b in range(B)]
    st = np.concatenate([s[0] for s in states] + [s[1].astype(np.complex128) for s in states])
    calc = api.BaseBoundaryIntegralCalculator(N, B, props, prob)
    out = torch.zeros(2 * N * B, dtype=torch.complex128, device="cuda:0")
    calc.run(T(st), out)
    o = out.cpu().numpy()
    e = ro.rhs(st, N, B, oprops, physics, "cuda")
    tol = 1e-15 * N + 1e-13
    if physics == "helium":
        tol *= 50   # cond(M) ~ N/(2 pi) for the reference's finite-depth operator
    assert rel(o[:N * B], e[:N * B]) <= tol
    assert rel(o[N * B:], e[N * B:]) <= tol
    _, _, aux = ro.rhs_single(states[0][0], states[0][1], oprops, physics, "cuda", full=True)
    assert rel(calc.devVelocitiesUpper.cpu().numpy()[:N], aux["v_upper"]) <= 3 * tol
    assert rel(calc.getDevZp().cpu().numpy()[:N], aux["Zp"]) <= tol
    assert rel(calc.devPhiPrime.cpu().numpy()[:N], aux["PhiPrime"]) <= tol


def test_helium_gmres_agrees_with_dense_lu_and_reports_convergence(api):
    """Finite-depth helium operator (spectrum between 1/2 and N/4pi): matrix-free GMRES vs the dense LU validation path."""
    N, depth = 512, 0.0942478
    props = api.ProblemProperties(rho=0.0, depth=depth)
    prob = api.HeliumBoundaryProblem(props)
    for amp in (0.1, 0.5):
        al = 2 * np.pi * np.arange(N) / N
        Z = al - 0.3 * amp * depth * np.sin(al) + 1j * amp * depth * np.cos(al)
        Phi = 0.2 * amp * depth * np.sin(al)
        st = T(ro.pack_state(Z, Phi))
        outs = []
        for mode in ("matrix_free", "dense_lu"):
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, prob, solve_mode=mode)
            out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
            calc.run(st, out)
            outs.append((out.cpu().numpy(), calc.getDevA().cpu().numpy().copy(), calc.solve_stats()))
        assert outs[0][2]["converged"] and 2 <= outs[0][2]["iterations"] <= 80
        assert rel(outs[0][1], outs[1][1]) <= 1e-10            # cond(M) ~ N/(2 pi)
        assert rel(outs[0][0], outs[1][0]) <= 1e-10


def _film(N, depth, amp):
    al = 2 * np.pi * np.arange(N) / N
    return ro.pack_state(al - 0.3 * amp * depth * np.sin(al) + 1j * amp * depth * np.cos(al), 0.2 * amp * depth * np.sin(al))


def test_recorded_helium_steps_with_the_device_driven_gmres_cycle(api):
    """Finite-depth helium film inside the RK4 stepper: recorded steps (CUDA graph) whose GMRES cycle -- Arnoldi, Givens rotations,
    per-member convergence -- runs on the device and is verified by the combined velocity sweep, against the host-driven restarted
    GMRES (RB_DEVICE_GMRES=0, unrecorded steps): both solve M a = b to 1e-13, so 30 steps agree to ~1e-13 (tolerance 1e-11); and
    against the oracle's dense-LU RK4 on the same film (north_star bar 1e-9; cond(M) ~ N / 2 pi)."""
    N, depth, dt, steps = 1024, 0.0942478, 1e-3, 30
    props = api.ProblemProperties(rho=0.0, depth=depth)
    y0 = _film(N, depth, 0.1)
    states = {}
    for dg in ("1", "0"):
        os.environ["RB_DEVICE_GMRES"] = dg
        try:
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), guess="warm")
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
        finally:
            os.environ.pop("RB_DEVICE_GMRES", None)
        y = T(y0)
        stp.initialize(y, True)
        stp.runSteps(steps)
        st, ss = calc.solve_stats(), stp.stats()
        assert st["converged"] and st["failed_solves"] == 0, st
        assert (ss["graph_launches"] >= steps // 2) == (dg == "1"), ss
        states[dg] = y.cpu().numpy()
    assert rel(states["1"], states["0"]) <= 1e-11
    oprops = ro.ProblemProperties(rho=0.0, depth=depth)
    f = lambda s: ro.rhs(s, N, 1, oprops, "helium", "cuda")
    ye = y0.copy()
    for _ in range(5):
        ye = ro.rk4_step(f, ye, dt)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    stp.initialize(y0, False)
    stp.runSteps(5)
    assert rel(stp.getState(), ye) <= 1e-9


def test_helium_ensemble_members_converge_individually(api):
    """A batch of films of very different amplitude in one recorded stepper: every member's GMRES cycle stops on its OWN residual
    against its OWN ||b|| (a member with a small right-hand side is not hidden behind the others), so each member's trajectory
    equals the one it has when stepped alone (1e-11)."""
    N, depth, dt, steps = 256, 0.0942478, 1e-3, 12
    props = api.ProblemProperties(rho=0.0, depth=depth)
    amps = (0.3, 1e-3, 0.05)
    members = [_film(N, depth, a) for a in amps]
    calc = api.BaseBoundaryIntegralCalculator(N, len(amps), props, api.HeliumBoundaryProblem(props), guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    y = T(api.ensemble_state(members, N))
    stp.initialize(y, True)
    stp.runSteps(steps)
    assert calc.solve_stats()["failed_solves"] == 0 and stp.stats()["graph_launches"] >= 1
    got = y.cpu().numpy()
    B = len(amps)
    for m, y0 in enumerate(members):
        alone = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumBoundaryProblem(props), guess="warm")
        s1 = api.AutonomousRungeKuttaStepper(alone, dt)
        s1.initialize(y0, False)
        s1.runSteps(steps)
        mine = np.concatenate([got[m * N:(m + 1) * N], got[B * N + m * N:B * N + (m + 1) * N]])
        e = s1.getState()
        # position relative to 2 pi; potential relative to ITS OWN scale (the small member's potential is 1e-5 of the large one's)
        assert rel(mine[:N], e[:N]) <= 1e-11, (m, rel(mine[:N], e[:N]))
        assert np.abs(mine[N:] - e[N:]).max() <= 1e-9 * np.abs(e[N:]).max() + 1e-16, m


def test_dense_lu_and_matrix_free_agree_at_N4096(api):
    """BASELINE config 3: steep trochoid (h = 0.4), N = 4096 -- assemble-and-factorise (the reference's way) vs the matrix-free
    iteration agree to <= 1e-12 per RHS."""
    N = 4096
    Z, Phi = ro.trochoid(N, 0.4)
    st = T(ro.pack_state(Z, Phi))
    props = api.ProblemProperties(rho=0.0)
    outs = []
    for mode in ("matrix_free", "dense_lu"):
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), solve_mode=mode)
        out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
        calc.run(st, out)
        outs.append((out.cpu().numpy(), calc.getDevA().cpu().numpy().copy()))
    assert rel(outs[0][1], outs[1][1]) <= 1e-12
    assert rel(outs[0][0], outs[1][0]) <= 1e-12


def test_small_amplitude_wave_follows_linear_dispersion(api):
    """BASELINE config 2: N = 1024 deep-water wave, g = 1, k = 1: omega^2 = k.  After t the profile is eps cos(x - t) + O(eps^2)."""
    N, eps, dt, steps = 1024, 1e-4, 1e-3, 400
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
    stp = api.AutonomousRungeKuttaStepper(calc, dt)
    Z, Phi = ro.sinusoid(N, eps)
    stp.initialize(ro.pack_state(Z, Phi), False)
    stp.runSteps(steps)
    y = stp.getState()
    t = dt * steps
    assert np.abs(y[:N].imag - eps * np.cos(y[:N].real - t)).max() <= 5 * eps * eps
    # phase speed: the crest (max of Y) moved by omega t / k = t
    k = int(np.argmax(y[:N].imag))
    assert abs(y[k].real % (2 * np.pi) - t) <= 2 * (2 * np.pi / N)


def test_ensemble_batch_1024_members_N512(api):
    """BASELINE config 5 (second half): 1024-member ensemble at N = 512, member m = trochoid h_m = 0.05 + 0.35 m / 1023
    (SURVEY.md section 8d).  A sample of members is compared with the oracle; all members must converge."""
    N, B = 512, 1024
    hs = 0.05 + 0.35 * np.arange(B) / (B - 1)
    Zs, Ps = zip(*(ro.trochoid(N, h) for h in hs))
    st = np.concatenate(list(Zs) + [p.astype(np.complex128) for p in Ps])
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, B, props, api.WaterBoundaryProblem(props))
    out = torch.zeros(2 * N * B, dtype=torch.complex128, device="cuda:0")
    calc.run(T(st), out)
    assert calc.solve_stats()["converged"]
    o = out.cpu().numpy()
    assert np.isfinite(o).all()
    oprops = ro.ProblemProperties(rho=0.0)
    for m in (0, 1, 511, 1022, 1023):
        v, dphi = ro.rhs_single(Zs[m], Ps[m], oprops, "water", "cuda")
        assert rel(o[m * N:(m + 1) * N], v) <= 1e-12
        assert rel(o[B * N + m * N: B * N + (m + 1) * N], dphi) <= 1e-12


def test_rhs_rejects_nothing_silently(api):
    """A surface with coincident points has no finite RHS: the iteration must report non-convergence, not a number."""
    N = 64
    Z, Phi = ro.trochoid(N, 0.3)
    Z = Z.copy()
    Z[10] = Z[11]
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), max_iterations=30)
    out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
    with pytest.raises(RuntimeError, match="did not converge"):   # strict (the default): the call itself fails
        calc.run(T(ro.pack_state(Z, Phi)), out)
    st = calc.solve_stats()
    assert not st["converged"] and not st["stagnated"] and st["failed_solves"] == 1
    calc.setStrict(False)                                          # statistics only: the call returns, the status still says failed
    calc.run(T(ro.pack_state(Z, Phi)), out)
    st = calc.solve_stats()
    assert not st["converged"] and not st["stagnated"] and st["failed_solves"] == 2
    # the same through the stepper: a failed step leaves the state as it was
    calc.setStrict(True)
    stp = api.AutonomousRungeKuttaStepper(calc, 1e-3)
    y0 = T(ro.pack_state(Z, Phi))
    keep = y0.clone()
    stp.initialize(y0, True)
    with pytest.raises(RuntimeError, match="did not converge"):
        stp.runStep()
    assert bool((torch.view_as_real(y0) == torch.view_as_real(keep)).all())


# ---- RK4 ------------------------------------------------------------------------------------------------------------
def test_stage_update_kernels(api):
    """cublasZaxpy / add_k_vectors replacement (L/AutonomousRungeKuttaStepper.cuh:349-361)."""
    import ctypes
    n = 1000
    rng = np.random.default_rng(1)
    y0, k1, k2, k3, k4 = (rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(5))
    lib = api._lib.load()
    out = torch.empty(n, dtype=torch.complex128, device="cuda:0")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    ty0, tk = T(y0), [T(k) for k in (k1, k2, k3, k4)]
    assert lib.rb_rk4_stage_update(p(out), p(ty0), p(tk[0]), 0.37, n, None) == 0
    torch.cuda.synchronize()
    assert np.abs(out.cpu().numpy() - (y0 + 0.37 * k1)).max() <= 1e-15
    assert lib.rb_rk4_final_update(p(ty0), p(tk[0]), p(tk[1]), p(tk[2]), p(tk[3]), 0.01, n, None) == 0
    torch.cuda.synchronize()
    assert np.abs(ty0.cpu().numpy() - (y0 + 0.01 / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4))).max() <= 1e-15


@pytest.mark.parametrize("N,h,steps,guess", [(64, 0.1, 100, "cold"), (64, 0.4, 100, "warm"), (256, 0.3, 100, "warm"),
                                             (1024, 0.4, 100, "warm")])
def test_rk4_100_steps_match_oracle(api, N, h, steps, guess):
    """north_star parity bar: <= 1e-9 relative in surface position and potential after 100 RK4 steps (dt = 1e-3);
    energy and volume drift no worse than the oracle's."""
    props = api.ProblemProperties(rho=0.0)
    oprops = ro.ProblemProperties(rho=0.0)
    Z, Phi = ro.trochoid(N, h)
    y0 = ro.pack_state(Z, Phi)
    f = lambda s: ro.rhs(s, N, 1, oprops, "water", "cuda")
    ye = y0.copy()
    for _ in range(steps):
        ye = ro.rk4_step(f, ye, 1e-3)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess=guess)
    stp = api.AutonomousRungeKuttaStepper(calc, 1e-3)
    stp.initialize(y0, False)
    stp.runSteps(steps)
    y = stp.getState()
    assert rel(y[:N], ye[:N]) <= 1e-9
    assert rel(y[N:], ye[N:]) <= 1e-9
    assert calc.solve_stats()["converged"]

    def diag(yv):
        v, dphi, aux = ro.rhs_single(yv[:N], yv[N:].real, oprops, full=True)
        e = ro.energies(yv[:N], aux["Zp"], yv[N:], v, oprops)
        return e["kinetic"] + e["potential"], ro.volume(yv[:N], aux["Zp"])

    e0, v0 = diag(y0)
    eo, vo = diag(ye)
    em, vm = diag(y)
    assert abs(em - e0) <= abs(eo - e0) + 1e-13 * abs(e0)
    assert abs(vm - v0) <= abs(vo - v0) + 1e-13


def test_rk4_device_alias_evolve_and_trajectory(api):
    """initialize(ptr, onDevice=true) aliases the caller's buffer (:312-318); runEvolution truncates the step count (:421);
    TrajectoryLogger semantics (L/TrajectoryLogger.cuh:65-73)."""
    N = 64
    props = api.ProblemProperties(rho=0.0)
    Z, Phi = ro.trochoid(N, 0.2)
    y0 = ro.pack_state(Z, Phi)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    stp = api.AutonomousRungeKuttaStepper(calc, 1e-2)
    dev = T(y0)
    stp.initialize(dev, True)
    stp.setLogging(1, 16)
    n = stp.runEvolution(0.0, 0.055)
    assert n == ro.rk4_num_steps(0.0, 0.055, 1e-2) == 5
    times, states = stp.copyTrajectory()
    assert len(times) == 5 and np.allclose(times, 0.01 * np.arange(1, 6))
    assert np.array_equal(states[-1], dev.cpu().numpy())          # the caller's buffer holds the evolved state
    ye = ro.rk4_evolve(lambda s: ro.rhs(s, N, 1, ro.ProblemProperties(rho=0.0)), y0, 0.0, 0.055, 1e-2)
    assert rel(states[-1], ye) <= 1e-12


def test_energies_match_oracle(api):
    N, h = 256, 0.3
    for physics, depth in (("water", 1.0), ("helium_inf", 0.5)):
        props = api.ProblemProperties(rho=0.0, depth=depth, kappa=0.01 if physics != "water" else 0.0)
        oprops = ro.ProblemProperties(rho=0.0, depth=depth, kappa=props.kappa)
        prob = api.WaterBoundaryProblem(props) if physics == "water" else api.HeliumInfiniteDepthBoundaryProblem(props)
        Z, Phi = ro.trochoid(N, h if physics == "water" else 0.05)
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, prob, compute_energies=True)
        out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
        calc.run(T(ro.pack_state(Z, Phi)), out)
        e = calc.energies()
        v, dphi, aux = ro.rhs_single(Z, Phi, oprops, physics, full=True)
        eo = ro.energies(Z, aux["Zp"], Phi.astype(np.complex128), v, oprops, physics)
        for k in ("kinetic", "potential", "surface", "volume_flux"):
            assert abs(e[k] - eo[k]) <= 1e-12 * max(1.0, abs(eo[k])), k
        assert abs(e["volume"] - ro.volume(Z, aux["Zp"])) <= 1e-12


# ---- legacy exports -------------------------------------------------------------------------------------------------
def test_legacy_rhs_exports(api):
    """calculateRHSFromVectors / ...Batched (L/Export.cuh:30-51): SI in, HeliumBoundaryProblem, nondimensionalised inside."""
    N, L, depth_si = 256, 1e-6, 15e-9
    op = ro.adimensionalize_properties(ro.ProblemProperties(L=L, rho=0.0, kappa=0.0, depth=depth_si))
    Z, Phi = ro.trochoid(N, 0.1 * op.depth)
    vx, vy, dphi = api.calculateRHSFromVectors(Z.real, Z.imag, Phi, L, 0.0, 0.0, depth_si)
    v, dp = ro.rhs_single(Z, Phi, op, "helium", "cuda")
    tol = 1e-10
    assert rel(vx + 1j * vy, v) <= tol and rel(dphi, dp) <= tol
    B = 3
    xs = np.tile(Z.real, B); ys = np.concatenate([Z.imag * (1 + 0.1 * b) for b in range(B)]); ps = np.tile(Phi, B)
    vx, vy, dphi = api.calculateRHSFromVectors(xs, ys, ps, L, 0.0, 0.0, depth_si, batchSize=B)
    for b in range(B):
        v, dp = ro.rhs_single(xs[b * N:(b + 1) * N] + 1j * ys[b * N:(b + 1) * N], ps[b * N:(b + 1) * N], op, "helium", "cuda")
        assert rel(vx[b * N:(b + 1) * N] + 1j * vy[b * N:(b + 1) * N], v) <= tol


def test_integrate_simulation_rk4_export(api):
    """integrateSimulationRK4 (declared L/Export.cuh:69, a stub in the reference): final-state and trajectory modes."""
    from superfluid_dynamics_b200 import _lib
    N, L, depth_si = 128, 1e-6, 15e-9
    sp = _lib.SimProperties(L, 0.0, 0.0, depth_si, False, 1, True)   # infinite depth film: iterative solve inside
    op = ro.adimensionalize_properties(ro.ProblemProperties(L=L, rho=0.0, kappa=0.0, depth=depth_si, infinite_depth=True))
    Z, Phi = ro.trochoid(N, 0.05 * op.depth)
    init = np.hstack((Z.real, Z.imag, Phi))
    dt_si = 1e-3 * op.base_time
    opts = _lib.RK4SolverOptions(dt_si, 0.0, 10.2 * dt_si, False)
    states, times = api.integrateSimulationRK4(init, sp, opts, N)
    assert states.shape == (1, 3 * N) and len(times) == 0
    f = lambda s: ro.rhs(s, N, 1, op, "helium", "cuda")
    ye = ro.rk4_evolve(f, ro.pack_state(Z, Phi), 0.0, 10.2e-3, 1e-3)
    got = states[0]
    assert rel(got[:N] + 1j * got[N:2 * N], ye[:N]) <= 1e-9 and rel(got[2 * N:], ye[N:].real) <= 1e-9
    opts = _lib.RK4SolverOptions(dt_si, 0.0, 10.2 * dt_si, True)
    states, times = api.integrateSimulationRK4(init, sp, opts, N)
    assert states.shape == (10, 3 * N) and len(times) == 10
    assert np.allclose(states[-1], got, rtol=0, atol=1e-14)


def test_cpp_compat_header_runs_reference_tests(api):
    """tests/cpp/compat_test.cu: the reference's gtest fixtures through the reference's C++ names (cusuperhelium_compat.cuh)."""
    import subprocess
    from superfluid_dynamics_b200 import build
    exe = build.build_compat_test(verbose=False)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout


def test_graph_rollback_when_recorded_sweeps_do_not_suffice(api):
    """A steep wave needs more sweeps than the first recorded graph holds (24): the step must be rolled back and redone, and the
    result must equal the run without graphs."""
    import os
    N, h, dt, steps = 256, 0.85, 1e-3, 6
    Z, Phi = ro.trochoid(N, h)
    y0 = ro.pack_state(Z, Phi)
    props = api.ProblemProperties(rho=0.0)
    res = {}
    for no_graph in ("0", "1"):
        os.environ["RB_NO_GRAPH"] = no_graph
        try:
            calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
            stp = api.AutonomousRungeKuttaStepper(calc, dt)
        finally:
            os.environ.pop("RB_NO_GRAPH", None)
        stp.initialize(y0, False)
        stp.runSteps(steps)
        res[no_graph] = (stp.getState(), stp.stats(), calc.solve_stats())
    assert res["0"][1]["fallback_steps"] >= 1 and res["0"][1]["graph_launches"] >= steps
    assert res["1"][1]["graph_launches"] == 0
    assert res["0"][2]["converged"] and res["1"][2]["converged"]
    assert rel(res["0"][0], res["1"][0]) <= 1e-12


def test_single_sweep_rhs_with_predicted_row_sums(api):
    """Optimistic stages: with the row sums of the extrapolated iterate predicted from their own history (O(N) work), the first
    O(N^2) sweep of a solve already verifies the iterate and delivers its velocities -- one sweep per RHS.  The predicted iterate
    is only ever a starting point: every solve still ends on a verified residual <= tolerance, so the trajectory must agree with
    the never-optimistic run to the solve tolerance, and with the oracle to the north_star bar."""
    N, h, dt, steps, tol = 1024, 0.3, 5e-4, 80, 1e-11
    Z, Phi = ro.trochoid(N, h)
    y0 = ro.pack_state(Z, Phi)
    props = api.ProblemProperties(rho=0.0)
    oprops = ro.ProblemProperties(rho=0.0)
    out = {}
    for policy in (0, 1):
        calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm", tolerance=tol)
        stp = api.AutonomousRungeKuttaStepper(calc, dt)
        stp.setGuess(4, 1)
        stp.setOptimistic(policy)
        stp.initialize(y0, False)
        stp.runSteps(steps)
        out[policy] = (stp.getState(), stp.guessStats(), calc.solve_stats(), stp.stats())
    gs = out[1][1]
    assert out[0][1]["optimistic_solves"] == 0
    assert gs["optimistic_solves"] > 0 and gs["one_sweep_solves"] >= 0.9 * gs["optimistic_solves"], gs
    assert max(gs["first_rel"]) <= tol, gs            # the guess verified as it stood
    assert out[1][2]["converged"] and out[0][2]["converged"]
    # fewer O(N^2) sweeps in total
    assert out[1][2]["total_iterations"] < 0.75 * out[0][2]["total_iterations"], (out[0][2], out[1][2])
    assert rel(out[1][0], out[0][0]) <= 1e-9
    f = lambda s: ro.rhs(s, N, 1, oprops, "water", "cuda")
    ye = y0.copy()
    for _ in range(steps):
        ye = ro.rk4_step(f, ye, dt)
    assert rel(out[1][0][:N], ye[:N]) <= 1e-9 and rel(out[1][0][N:], ye[N:]) <= 1e-9


def test_rk45_generic_problem_matches_reference_table_and_oracle(api):
    """TEST(ODE_Solvers, RK45) (T/ODESolverTests.cuh:248-421): dz/dt = i z through RK45_std_complex with a caller-supplied
    AutonomousProblem::run, setTolerance(1e-8, 1e-8), runEvolution(0, 10); reference table within its own 1e-2, and the oracle's
    restatement of L/RK45.cuh step for step."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ode_rotation.npz"))
    j = np.arange(256)
    z0 = 2 * np.pi * j / 256 + 1j * np.sin(2 * np.pi * j * 0.01)

    class Oscillatory:           # OscillatoryProblemStdComplex<N>, T/ODESolverTests.cuh:44-71
        def run(self, state, rhs):
            rhs.copy_(1j * state)

    stp = api.RK45_std_complex(Oscillatory(), 1e-3, n=256)
    stp.initialize(T(z0), True)
    stp.setTolerance(1e-8, 1e-8)
    assert stp.runEvolution(0.0, 10.0) == stp.ReachedEndTime
    y = stp.getState()
    assert np.abs(y.real - g["rk45_x"]).max() <= 1e-2 and np.abs(y.imag - g["rk45_y"]).max() <= 1e-2
    o = ro.RK45(lambda s: 1j * s, ro.RK45Options(atol=1e-8, rtol=1e-8, initial_timestep=1e-3))
    o.initialize(z0)
    assert o.run_evolution(0.0, 10.0) == "ReachedEndTime"
    st = stp.stats()
    assert (st["accepted"], st["rejected"], st["rhs_evaluations"]) == (o.n_accepted, o.n_rejected, o.n_rhs)
    assert np.abs(y - o.y).max() <= 1e-11 and abs(stp.getCurrentTime() - 10.0) <= 1e-12
    assert np.abs(stp.getY().cpu().numpy() - y).max() == 0.0


def test_rk45_on_the_boundary_integral_rhs_matches_oracle(api):
    """the adaptive stepper over the same RHS as the RK4 path (runSimulationWater takes RK45_Options,
    L/SimulationRunner.cuh:582-602); step sequence and final state against the oracle."""
    N, h = 64, 0.3
    Z, Phi = ro.trochoid(N, h)
    y0 = ro.pack_state(Z, Phi)
    props = api.ProblemProperties(rho=0.0)
    oprops = ro.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props), guess="warm")
    stp = api.RK45_std_complex(calc)
    stp.setOptions(api.RK45_Options(atol=1e-9, rtol=1e-9, initial_timestep=1e-3))
    stp.initialize(y0, False)
    assert stp.runStep(0) in (stp.StepAccepted, stp.StepRejected)
    assert stp.runEvolution(stp.getCurrentTime(), 0.05) == stp.ReachedEndTime
    o = ro.RK45(lambda s: ro.rhs(s, N, 1, oprops, "water", "cuda"), ro.RK45Options(atol=1e-9, rtol=1e-9, initial_timestep=1e-3))
    o.initialize(y0)
    o.run_step()
    assert o.run_evolution(o.t, 0.05) == "ReachedEndTime"
    st = stp.stats()
    assert (st["accepted"], st["rejected"]) == (o.n_accepted, o.n_rejected), (st, o.n_accepted, o.n_rejected)
    y = stp.getState()
    assert rel(y[:N], o.y[:N]) <= 1e-9 and rel(y[N:], o.y[N:]) <= 1e-9
    assert abs(stp.getCurrentTimeStep() - o.h) <= 1e-6 * o.h


@pytest.mark.parametrize("N,G", [(1024, 2), (2048, 4), (4096, 8), (4096, 3), (8192, 2), (16384, 8)])
def test_row_range_of_a_shard_gives_the_same_rows(api, N, G):
    """What rank r of a G-rank row-sharded run computes, on one GPU (rb_debug_set_row_range): the raw cotangent sum restricted to the
    rank's row cells must equal the same rows of the whole-surface sum -- for the warp-per-row-group kernel (N <= 8192: rows per
    rank from 512 to 4096, i.e. 1, 2 and 4 rows per warp, shared row groups, two staged tiles) to the last bit or to round-off
    when the rows per warp differ, for the tiled kernel (N = 16384) to round-off (its chunking changes with the row count)."""
    Z, Phi = ro.trochoid(N, 0.4)
    props = api.ProblemProperties(rho=0.0)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.WaterBoundaryProblem(props))
    calc.zPhiDerivative(T(Z), T(Phi.astype(np.complex128)))
    x = np.cos(3 * 2 * np.pi * np.arange(N) / N) + 0.3 * np.sin(2 * np.pi * np.arange(N) / N) + 0.1
    whole = calc.cotangentSum(T(Z), T(x)).cpu().numpy()
    ncell = (N + 255) // 256
    for r in sorted({0, G // 2, G - 1}):
        c0 = r * ncell // G
        c1 = (r + 1) * ncell // G
        calc.debugSetRowRange(c0, c1 - c0)
        part = calc.cotangentSum(T(Z), T(x)).cpu().numpy()
        rows = slice(c0 * 256, min(N, c1 * 256))
        assert rel(part[rows], whole[rows]) <= 1e-13, (r, calc.sweepPlan())
    calc.debugSetRowRange(0, 0)
    again = calc.cotangentSum(T(Z), T(x)).cpu().numpy()
    assert np.array_equal(again, whole)


def test_row_sharded_two_gpus_match_single_gpu(api):
    """N > 1: row-sharded run on two GPUs (torchrun, one rank per GPU) vs the single-GPU run; skipped on a one-GPU box."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "gpu_multirank.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "FAIL" not in r.stdout


# ---- optomechanically driven film ------------------------------------------------------------
def _film_state(N, depth, amp=0.1):
    al = 2 * np.pi * np.arange(N) / N
    Z = al - 0.3 * amp * depth * np.sin(al) + 1j * amp * depth * np.cos(al)
    Phi = 0.2 * amp * depth * np.sin(al)
    D = 0.02 * np.cos(al - 0.4) + 0.01
    return np.concatenate([Z, Phi.astype(np.complex128), D.astype(np.complex128)])


OPTO = dict(detuning=0.5, gamma=2.0, G=3.0, Tau=0.7, max_intensity=1e32, location_x0_mode=3.0, sigma_optical_mode=0.8, Beta=2e-33,
            DampingStrength=0.01)


@pytest.mark.parametrize("N,batch", [(64, 1), (300, 1), (128, 3)])
def test_augmented_rhs_and_light_intensity_match_oracle(api, N, batch):
    """rb_augmented_rhs / rb_light_intensity vs the oracle's restatement of L/createM.cuh:138-169 and L/LightIntensity.cuh:17-33,
    incl. a ragged N and a batch (layout [Z_b | Phi_b | D_b], 3 N B): 1e-10 relative (thin film, cond(M) ~ N / 2 pi)."""
    depth = 0.0942478
    props = api.ProblemProperties(rho=1.0, depth=depth)
    oprops = ro.ProblemProperties(rho=1.0, depth=depth)
    v = api.OptomechanicalVariables(**OPTO)
    ov = ro.OptomechanicalVariables(**OPTO)
    members = [_film_state(N, depth, 0.1 * (1 + 0.5 * b)) for b in range(batch)]
    st = np.concatenate([m[blk * N:(blk + 1) * N] for blk in range(3) for m in members])
    calc = api.BaseBoundaryIntegralCalculator(N, batch, props, api.HeliumDrivenAutonomousProblem(props, v))
    integ = api.AugmentedBoundaryIntegrator(calc, api.DelayedIntensityIntegrator(v))
    assert abs(v.drive_strength - ro.drive_strength(ov, oprops)) <= 1e-15 * abs(v.drive_strength)
    out = torch.zeros(3 * N * batch, dtype=torch.complex128, device="cuda:0")
    integ.run(T(st), out)
    o = out.cpu().numpy()
    for b, m in enumerate(members):
        e = ro.augmented_rhs(m, N, oprops, ov)
        for blk in range(3):
            got = o[(blk * batch + b) * N:(blk * batch + b + 1) * N]
            assert rel(got, e[blk * N:(blk + 1) * N]) <= 1e-10, (b, blk)
    inten = integ.lightIntensity(T(members[0][:N])).cpu().numpy()
    assert rel(inten, ro.light_intensity(members[0][:N].imag, members[0][:N].real, ov)) <= 1e-14


def test_augmented_rk4_and_legacy_exports_match_oracle(api):
    """AutonomousRungeKuttaStepper<std_complex, 3N> over the augmented system (50 steps, 1e-9), device aliasing, runEvolution's
    truncated step count; calculateRhsAugmentedOptomechanical / integrateAugmentedOptomechanicalSimulationRK4 (L/Export.cuh:75-78)
    with SI inputs against the oracle's adimensionalize_properties / adimensionalize_optomechanical (L/Export.cu:1222-1275)."""
    from superfluid_dynamics_b200 import _lib
    N, depth, dt = 128, 0.0942478, 1e-3
    props = api.ProblemProperties(rho=1.0, depth=depth)
    oprops = ro.ProblemProperties(rho=1.0, depth=depth)
    v = api.OptomechanicalVariables(**OPTO)
    ov = ro.OptomechanicalVariables(**OPTO)
    y0 = _film_state(N, depth)
    f = lambda s: ro.augmented_rhs(s, N, oprops, ov)
    ye = y0.copy()
    for _ in range(50):
        ye = ro.rk4_step(f, ye, dt)
    calc = api.BaseBoundaryIntegralCalculator(N, 1, props, api.HeliumDrivenAutonomousProblem(props, v), guess="warm")
    stp = api.AugmentedRungeKuttaStepper(api.AugmentedBoundaryIntegrator(calc, api.DelayedIntensityIntegrator(v)), dt)
    dev = T(y0)
    stp.initialize(dev, True)
    assert stp.runEvolution(0.0, 0.0505) == 50
    y = stp.getState()
    assert np.array_equal(y, dev.cpu().numpy())
    for blk in range(3):
        assert rel(y[blk * N:(blk + 1) * N], ye[blk * N:(blk + 1) * N]) <= 1e-9, blk
    assert abs(stp.currentTime() - 0.05) <= 1e-12

    # legacy exports, SI in: a 15 nm film on a 1 um period (A/kernel.cu:79), laboratory-unit optomechanical variables
    L, d_si = 1e-6, 15e-9
    sp = _lib.SimProperties(L=L, rho=150.0, kappa=0.0, depth=d_si, use_expansions=False, expansion_order=1, infinite_depth=False)
    op_si = ro.adimensionalize_properties(ro.ProblemProperties(L=L, rho=150.0, kappa=0.0, depth=d_si))
    cv = _lib.COptomechanicalVariables(detuning=0.5 / op_si.base_time, gamma=2.0 / op_si.base_time,
                                       G=3.0 / (op_si.base_time * op_si.base_length), tau=0.7 * op_si.base_time, max_intensity=5e7,
                                       initial_time=0.0, location_x0_mode=3.0 * op_si.base_length,
                                       sigma_optical_mode=0.8 * op_si.base_length, beta=1.0, damping_strength=0.01)
    ov_si = ro.adimensionalize_optomechanical(
        ro.OptomechanicalVariables(detuning=cv.detuning, gamma=cv.gamma, G=cv.G, Tau=cv.tau, max_intensity=cv.max_intensity,
                                   location_x0_mode=cv.location_x0_mode, sigma_optical_mode=cv.sigma_optical_mode, Beta=cv.beta,
                                   DampingStrength=cv.damping_strength), op_si)
    ys = _film_state(N, op_si.depth)
    flat = np.concatenate([ys[:N].real, ys[:N].imag, ys[N:2 * N].real, ys[2 * N:].real])
    got = api.calculateRhsAugmentedOptomechanical(flat, sp, cv, N)
    e = ro.augmented_rhs(ys, N, op_si, ov_si)
    exp = np.concatenate([e[:N].real, e[:N].imag, e[N:2 * N].real, e[2 * N:].real])
    for blk in range(4):
        assert rel(got[blk * N:(blk + 1) * N], exp[blk * N:(blk + 1) * N]) <= 1e-10, blk
    opts = _lib.RK4SolverOptions(timeStep=1e-3 * op_si.base_time, t0=0.0, t1=0.0105 * op_si.base_time, returnTrajectory=True)
    states, times = api.integrateAugmentedOptomechanicalSimulationRK4(flat, sp, opts, cv, N)
    assert states.shape == (10, 4 * N) and len(times) == 10
    fs = lambda s: ro.augmented_rhs(s, N, op_si, ov_si)
    yo = ys.copy()
    for _ in range(10):
        yo = ro.rk4_step(fs, yo, 1e-3)
    expf = np.concatenate([yo[:N].real, yo[:N].imag, yo[N:2 * N].real, yo[2 * N:].real])
    for blk in range(4):
        assert rel(states[-1][blk * N:(blk + 1) * N], expf[blk * N:(blk + 1) * N]) <= 1e-9, blk
    opts.returnTrajectory = False
    final, t2 = api.integrateAugmentedOptomechanicalSimulationRK4(flat, sp, opts, cv, N)
    assert final.shape == (1, 4 * N) and len(t2) == 0 and rel(final[0], states[-1]) <= 1e-13


def test_time_dependent_drive_matches_oracle(api):
    """rb_timed_* vs the oracle's TimedDrive (L/createM.cuh:119-136, L/DelayedIntensityTerm.cuh:16-33, L/RK4_Time_Dependent.cuh):
    one timed RHS with and without saving, runStep's un-advanced time, runEvolution with a starting time, the delayed-intensity
    buffer, and integrateOptomechanicalSimulationRK4 with SI inputs (times logged at the start of each step)."""
    from superfluid_dynamics_b200 import _lib
    N, depth, dt = 128, 0.0942478, 1e-3
    props = api.ProblemProperties(rho=1.0, depth=depth)
    oprops = ro.ProblemProperties(rho=1.0, depth=depth)
    v = api.OptomechanicalVariables(**OPTO)
    ov = ro.OptomechanicalVariables(**OPTO)
    y0 = _film_state(N, depth)[:2 * N]
    integ = api.TimedBoundaryIntegrator(N, 1, props, api.HeliumWithOptomechanicalDrivingProblem(props, v), guess="warm")
    stp = api.RungeKuttaStepper(integ, dt)
    td = ro.TimedDrive(N, oprops, ov)
    # RHS level: first call at t == starting time takes D = I and saves it; a later call decays it
    stp.setStartingTime(0.3)
    td.set_starting_time(0.3)
    out = torch.zeros(2 * N, dtype=torch.complex128, device="cuda:0")
    for time, save in ((0.3, True), (0.3005, False), (0.301, True), (0.3015, False)):
        stp.run(time, save, T(y0), out)
        e = td.rhs(y0, time, save)
        assert rel(out.cpu().numpy()[:N], e[:N]) <= 1e-10 and rel(out.cpu().numpy()[N:], e[N:]) <= 1e-10, (time, save)
        assert rel(stp.delayedIntensity().cpu().numpy(), td.delayed) <= 1e-14
    # runStep does not advance the time; runStep(advance=True) does
    stp.initialize(y0, False)
    stp.setStartingTime(0.0)
    td.set_starting_time(0.0)
    stp.runStep()
    assert stp.currentTime() == 0.0
    y1 = td.step(y0, dt)
    assert rel(stp.getState(), y1) <= 1e-10
    # evolution from a starting time, with the trajectory (time logged at the START of each step)
    stp.initialize(y0, False)
    stp.setOptions(dt, returnTrajectory=True)
    assert stp.runEvolution(0.25, 0.25 + 30.5 * dt) == 30
    ye, times, states = ro.TimedDrive(N, oprops, ov).evolve(y0, 0.25, 0.25 + 30.5 * dt, dt)
    y = stp.getState()
    assert rel(y[:N], ye[:N]) <= 1e-9 and rel(y[N:], ye[N:]) <= 1e-9
    assert abs(stp.currentTime() - (0.25 + 30 * dt)) <= 1e-12
    lt, ls = stp.copyTrajectory()
    assert len(lt) == 30 and np.abs(lt - times).max() <= 1e-12 and ls.shape == (30, 2 * N)
    assert rel(ls[4], states[4]) <= 1e-9 and np.array_equal(ls[-1], y)
    stp.setOptions(dt, returnTrajectory=False)
    lt, ls = stp.copyTrajectory()
    assert len(lt) == 0 and ls.shape == (1, 2 * N) and np.array_equal(ls[0], y)

    # legacy export, SI in
    L, d_si = 1e-6, 15e-9
    sp = _lib.SimProperties(L=L, rho=150.0, kappa=0.0, depth=d_si, use_expansions=False, expansion_order=1, infinite_depth=False)
    op_si = ro.adimensionalize_properties(ro.ProblemProperties(L=L, rho=150.0, kappa=0.0, depth=d_si))
    cv = _lib.COptomechanicalVariables(detuning=0.5 / op_si.base_time, gamma=2.0 / op_si.base_time,
                                       G=3.0 / (op_si.base_time * op_si.base_length), tau=0.7 * op_si.base_time, max_intensity=5e7,
                                       initial_time=0.0, location_x0_mode=3.0 * op_si.base_length,
                                       sigma_optical_mode=0.8 * op_si.base_length, beta=1.0, damping_strength=0.01)
    ov_si = ro.adimensionalize_optomechanical(
        ro.OptomechanicalVariables(detuning=cv.detuning, gamma=cv.gamma, G=cv.G, Tau=cv.tau, max_intensity=cv.max_intensity,
                                   location_x0_mode=cv.location_x0_mode, sigma_optical_mode=cv.sigma_optical_mode, Beta=cv.beta,
                                   DampingStrength=cv.damping_strength), op_si)
    ys = _film_state(N, op_si.depth)[:2 * N]
    flat = np.concatenate([ys[:N].real, ys[:N].imag, ys[N:].real])
    opts = _lib.RK4SolverOptions(timeStep=1e-3 * op_si.base_time, t0=0.1 * op_si.base_time, t1=0.1105 * op_si.base_time,
                                 returnTrajectory=True)
    states, times = api.integrateOptomechanicalSimulationRK4(flat, sp, opts, cv, N)
    yo, to, so = ro.TimedDrive(N, op_si, ov_si).evolve(ys, 0.1, 0.1105, 1e-3)
    assert states.shape == (10, 3 * N) and len(times) == 10
    assert np.abs(times - to).max() <= 1e-12 and abs(times[0] - 0.1) <= 1e-12      # time at the START of each step
    expf = np.concatenate([yo[:N].real, yo[:N].imag, yo[N:].real])
    for blk in range(3):
        assert rel(states[-1][blk * N:(blk + 1) * N], expf[blk * N:(blk + 1) * N]) <= 1e-9, blk
    opts.returnTrajectory = False
    final, t2 = api.integrateOptomechanicalSimulationRK4(flat, sp, opts, cv, N)
    assert final.shape == (1, 3 * N) and len(t2) == 0 and rel(final[0], states[-1]) <= 1e-13
