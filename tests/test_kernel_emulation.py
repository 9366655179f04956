"""CPU tier: the library's small CUDA kernels executed WITHOUT a GPU.  tests/cpp/emu_kernels.cpp compiles the shipped kernel sources
(csrc/implicit_kernels.cuh, csrc/lu_kernels.cu) with g++ against tests/cpp/cuda_emu.h, a minimal host emulation of the CUDA
execution model (one std::thread per CUDA thread, std::barrier for __syncthreads, buffered exchanges for warp shuffles, the PTX
fragment layout for the m8n8k4 FP64 mma), and launches them with the grids the library uses.  This checks index arithmetic,
shared-memory staging, barrier placement and the fragment bookkeeping of the blocked LU against the oracle / LAPACK; it says nothing
about speed and does not replace the GPU tier (tests/test_zz_gpu_implicit.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import roberts_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = ctypes.POINTER(ctypes.c_double)


def _p(a):
    return a.ctypes.data_as(D)


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "cpp", "emu_kernels.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "_build", "libemu_kernels.so")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "cuda_emu.h"), os.path.join(ROOT, "superfluid_dynamics_b200", "csrc", "lu_kernels.cu"),
            os.path.join(ROOT, "superfluid_dynamics_b200", "csrc", "implicit_kernels.cuh"),
            os.path.join(ROOT, "superfluid_dynamics_b200", "csrc", "launch.cuh")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-DRB_EMULATE", "-I",
                               os.path.join(ROOT, "tests", "cpp"), src, "-o", out])
    lib = ctypes.CDLL(out)
    lib.emu_lu_factor_blocked.restype = ctypes.c_int
    return lib


def test_state_conversions(emu):
    N = 37
    rng = np.random.default_rng(0)
    y = rng.standard_normal(3 * N)
    s = np.zeros(2 * N, np.complex128)
    emu.emu_real_to_complex_state(_p(y), _p(s.view(np.float64)), N)
    assert np.array_equal(s, ro.real_to_complex_state(y, N))
    r = rng.standard_normal(2 * N) + 1j * rng.standard_normal(2 * N)
    out = np.zeros(3 * N)
    emu.emu_complex_to_real_rhs(_p(r.view(np.float64)), _p(out), N)
    assert np.array_equal(out, ro.complex_to_real_rhs(r, N))


@pytest.mark.parametrize("N", [5, 37, 300])
def test_perturbed_states_kernel_bit_exact(emu, N):
    rng = np.random.default_rng(N)
    st = rng.standard_normal(2 * N) + 1j * np.concatenate([rng.standard_normal(N), np.zeros(N)])
    eps = 1e-6
    pos = np.full(6 * N * N, np.nan + 0j)
    neg = np.full(6 * N * N, np.nan + 0j)
    emu.emu_perturbed_states(_p(st.view(np.float64)), _p(pos.view(np.float64)), _p(neg.view(np.float64)), ctypes.c_double(eps), N)
    assert np.array_equal(pos, ro.perturbed_states(st, N, eps))
    assert np.array_equal(neg, ro.perturbed_states(st, N, -eps))
    one = np.full(6 * N * N, np.nan + 0j)
    emu.emu_perturbed_states(_p(st.view(np.float64)), _p(one.view(np.float64)), None, ctypes.c_double(-eps), N)
    assert np.array_equal(one, neg)


@pytest.mark.parametrize("N", [3, 16, 37])
def test_jacobian_from_perturbed_kernel(emu, N):
    rng = np.random.default_rng(N)
    pos = rng.standard_normal(6 * N * N) + 1j * rng.standard_normal(6 * N * N)
    neg = rng.standard_normal(6 * N * N) + 1j * rng.standard_normal(6 * N * N)
    C = np.full(9 * N * N, np.nan)
    emu.emu_jacobian_from_perturbed(_p(pos.view(np.float64)), _p(neg.view(np.float64)), _p(C), N, ctypes.c_double(1e-6))
    J = ro.jacobian_from_perturbed(pos, neg, N, 1e-6)
    assert np.array_equal(C, J.ravel(order="F"))      # column-major, every entry written


def test_gauss_legendre_elementwise_kernels(emu):
    N = 41
    n = 3 * N
    rng = np.random.default_rng(2)
    y, k, dK, fy = rng.standard_normal(n), rng.standard_normal(2 * n), rng.standard_normal(2 * n), rng.standard_normal(2 * n)
    h, alpha = 0.037, 0.25
    st = np.zeros(2 * n)
    cs = np.zeros(4 * N, np.complex128)
    emu.emu_gl2_stage_states_batched(_p(y), ctypes.c_double(h), _p(k), _p(st), _p(cs.view(np.float64)), N)
    A = ro.GL_A
    y1 = y + h * (A[0][0] * k[:n] + A[0][1] * k[n:])
    y2 = y + h * (A[1][0] * k[:n] + A[1][1] * k[n:])
    assert np.abs(st[:n] - y1).max() <= 1e-15 and np.abs(st[n:] - y2).max() <= 1e-15
    # the batch-2 state [Z_1 | Z_2 | Phi_1 | Phi_2] holds exactly the same numbers as the real stage states
    s1, s2 = ro.real_to_complex_state(st[:n], N), ro.real_to_complex_state(st[n:], N)
    assert np.array_equal(cs, np.concatenate([s1[:N], s2[:N], s1[N:], s2[N:]]))
    # and the batched RHS comes back as f(y1) | f(y2) in the real layout
    r1, r2 = rng.standard_normal(2 * N) + 1j * rng.standard_normal(2 * N), rng.standard_normal(2 * N) + 1j * rng.standard_normal(2 * N)
    crhs = np.concatenate([r1[:N], r2[:N], r1[N:], r2[N:]])
    f12 = np.zeros(2 * n)
    emu.emu_gl2_batched_rhs_to_real(_p(crhs.view(np.float64)), _p(f12), N)
    assert np.array_equal(f12, np.concatenate([ro.complex_to_real_rhs(r1, N), ro.complex_to_real_rhs(r2, N)]))
    kt = np.zeros(2 * n)
    emu.emu_gl2_trial(_p(k), ctypes.c_double(alpha), _p(dK), _p(kt), ctypes.c_size_t(2 * n))
    assert np.abs(kt - (k + alpha * dK)).max() <= 1e-15
    neg = np.zeros(2 * n)
    emu.emu_gl2_negate(_p(dK), _p(neg), ctypes.c_size_t(2 * n))
    assert np.array_equal(neg, -dK)
    nxt = np.zeros(n)
    emu.emu_gl2_next_state(_p(y), ctypes.c_double(h), _p(k), _p(nxt), ctypes.c_size_t(n))
    assert np.abs(nxt - (y + h * 0.5 * (k[:n] + k[n:]))).max() <= 1e-15
    # residual with both sums in one single-CTA pass (1024 threads, barriers)
    R, sums = np.zeros(2 * n), np.zeros(2)
    emu.emu_gl2_residual(_p(fy), _p(k), _p(R), ctypes.c_size_t(2 * n), _p(sums))
    assert np.array_equal(R, k - fy)
    assert abs(sums[0] - R @ R) <= 1e-12 * (R @ R) and abs(sums[1] - k @ k) <= 1e-12 * (k @ k)


@pytest.mark.parametrize("n", [1, 7, 40])
def test_newton_matrix_kernel(emu, n):
    rng = np.random.default_rng(n)
    J1, J2 = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    M = np.full(4 * n * n, np.nan)
    emu.emu_gl2_newton_matrix(_p(np.asfortranarray(J1).ravel(order="F")), _p(np.asfortranarray(J2).ravel(order="F")),
                              ctypes.c_double(0.05), _p(M), ctypes.c_size_t(n))
    exp = ro.gl2_newton_matrix(J1, J2, 0.05)
    assert np.abs(M.reshape(2 * n, 2 * n).T - exp).max() <= 1e-15      # column-major, every entry written


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 64, 65, 100, 131, 300])
def test_blocked_lu_kernels_against_lapack(emu, n):
    """lu_kernels.cu thread for thread: panel (pivot search by shuffles, swaps, rank-1 updates), swap, trsm, the 64 x 64 trailing
    update on the emulated m8n8k4 mma, gemv, with b eliminated on the fly; then the upper-triangular solve here."""
    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))          # not diagonally dominant: a row interchange in almost every column
    x_true = rng.standard_normal(n)
    b = A @ x_true
    Af = np.asfortranarray(A).ravel(order="F").copy()
    bf = b.copy()
    info = emu.emu_lu_factor_blocked(_p(Af), _p(bf), n)
    assert info == 0
    U = np.triu(Af.reshape(n, n).T)
    x = np.linalg.solve(U, bf)
    ref = np.linalg.solve(A, b)
    assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(A @ x - b).max() <= 1e-11 * np.abs(A).sum(axis=1).max() * np.abs(x).max()


@pytest.mark.parametrize("n", [1, 31, 32, 33, 70, 129])
def test_blocked_lu_solve_with_blocked_back_substitution(emu, n):
    """launch_lu_solve_blocked end to end: the factorisation above followed by lu_backsolve_blocked_kernel (warp-level triangle solve
    in shared memory, block-wide update of the rows above); b returns the solution."""
    rng = np.random.default_rng(200 + n)
    A = rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    Af = np.asfortranarray(A).ravel(order="F").copy()
    x = b.copy()
    emu.emu_lu_solve_blocked.restype = ctypes.c_int
    assert emu.emu_lu_solve_blocked(_p(Af), _p(x), n) == 0
    ref = np.linalg.solve(A, b)
    assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max()


def test_blocked_lu_reports_singular_column(emu):
    n = 70
    rng = np.random.default_rng(1)
    A = rng.standard_normal((n, n))
    A[:, 40] = 0.0
    Af = np.asfortranarray(A).ravel(order="F").copy()
    bf = rng.standard_normal(n)
    assert emu.emu_lu_factor_blocked(_p(Af), _p(bf), n) == 41


# ---- the host logic of csrc/implicit.cu over a mock RHS assembler (tests/cpp/emu_implicit.cpp) -----------------------------------------
@pytest.fixture(scope="module")
def impl():
    """libemu_implicit.so: implicit.cu compiled by g++ (kernels emulated, CUDA runtime stubbed), rb_rhs forwarded to the oracle."""
    from superfluid_dynamics_b200 import _lib
    src = os.path.join(ROOT, "tests", "cpp", "emu_implicit.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "_build", "libemu_implicit.so")
    csrc = os.path.join(ROOT, "superfluid_dynamics_b200", "csrc")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, os.path.join(ROOT, "tests", "cpp", "cuda_emu.h"), os.path.join(csrc, "implicit.cu"), os.path.join(csrc, "implicit_kernels.cuh"),
            os.path.join(csrc, "launch.cuh"), os.path.join(ROOT, "include", "roberts_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-DRB_EMULATE", "-I",
                               os.path.join(ROOT, "tests", "cpp"), src, "-o", out])
    lib = ctypes.CDLL(out)
    for name, (res, args) in _lib.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    physics_of = {0: "water", 1: "helium", 2: "helium_inf"}

    RHS = ctypes.CFUNCTYPE(None, ctypes.POINTER(_lib.rb_props), ctypes.c_int, ctypes.c_int, D, D)
    ADIM = ctypes.CFUNCTYPE(None, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, D)

    def rhs_cb(pp, N, B, state, out):
        p = pp.contents
        props = ro.ProblemProperties(rho=p.rho, kappa=p.kappa, depth=p.depth, use_expansions=bool(p.use_expansions),
                                     expansion_order=p.expansion_order, infinite_depth=bool(p.infinite_depth))
        st = np.ctypeslib.as_array(state, shape=(4 * B * N,)).view(np.complex128)
        r = ro.rhs(st, N, B, props, physics_of[p.physics], "cuda")
        np.ctypeslib.as_array(out, shape=(4 * B * N,))[:] = r.view(np.float64)

    def adim_cb(L, rho, kappa, depth, out):
        op = ro.adimensionalize_properties(ro.ProblemProperties(L=L, rho=rho, kappa=kappa, depth=depth))
        out[0], out[1], out[2] = op.rho, op.kappa, op.depth

    lib._keep = (RHS(rhs_cb), ADIM(adim_cb))
    lib.emu_set_callbacks(*lib._keep)
    lib._props = _lib.rb_props
    lib._lib = _lib
    return lib


def _props(impl, physics, depth, rho):
    p = impl._props()
    impl.rb_default_props(ctypes.byref(p))
    p.physics = {"water": 0, "helium": 1, "helium_inf": 2}[physics]
    p.depth, p.rho = depth, rho
    return p


def _film(N, depth, amp):
    a = 2 * np.pi * np.arange(N) / N
    return np.concatenate([a - 0.3 * amp * depth * np.sin(a), amp * depth * np.cos(a), 0.2 * amp * depth * np.sin(a)])


def test_implicit_host_logic_real_rhs_and_jacobian(impl):
    """rb_real_rhs, rb_perturbed_states and rb_jacobian_calculate as implicit.cu runs them (kernels, grids, buffers, order): with
    the oracle's RHS behind rb_rhs they must reproduce the oracle's real RHS and Jacobian exactly."""
    N, depth = 9, 0.3
    p = _props(impl, "helium", depth, 1.0)
    oprops = ro.ProblemProperties(rho=1.0, depth=depth)
    y = _film(N, depth, 0.1)
    s = impl.rb_create(N, 1, ctypes.byref(p))
    out = np.zeros(3 * N)
    assert impl.rb_real_rhs(s, _p(y), _p(out)) == 0
    assert np.array_equal(out, ro.real_rhs(y, N, oprops, "helium"))
    st = ro.real_to_complex_state(y, N)
    zb = np.zeros(6 * N * N, np.complex128)
    assert impl.rb_perturbed_states(_p(st.view(np.float64)), _p(zb.view(np.float64)), 1e-6, N, None) == 0
    assert np.array_equal(zb, ro.perturbed_states(st, N, 1e-6))
    j = impl.rb_jacobian_create(N, ctypes.byref(p))
    assert j
    assert impl.rb_jacobian_set_epsilon(j, 1e-5) == 0
    J = np.zeros(9 * N * N)
    assert impl.rb_jacobian_calculate(j, _p(y), _p(J)) == 0
    assert np.array_equal(J.reshape(3 * N, 3 * N).T, ro.jacobian_fd(y, N, oprops, "helium", 1e-5))
    assert impl.emu_rhs_calls(impl.rb_jacobian_solver(j)) == 2      # one batched evaluation per sign
    impl.rb_jacobian_destroy(j)
    impl.rb_destroy(s)


def _gl2(impl, c, trajectory=True):
    p = _props(impl, c["physics"], c["depth"], 0.0 if c["physics"] == "water" else 1.0)
    s = impl.rb_create(c["N"], 1, ctypes.byref(p))
    j = impl.rb_jacobian_create(c["N"], ctypes.byref(p))
    o = impl._lib.rb_gl2_options()
    impl.rb_gl2_default_options(ctypes.byref(o))
    o.stepSize, o.newtonTolerance, o.maxNewtonIterations = c["h"], c["tol"], c["maxit"]
    o.allowSimplifiedFallback, o.returnTrajectory, o.maxStepsHalves = int(c["fallback"]), int(trajectory), c["halves"]
    g = impl.rb_gl2_create(s, j, ctypes.byref(o))
    assert g
    return s, j, g, o


def _trajectory(impl, g, N):
    tp, sp = D(), D()
    tc, sc = ctypes.c_size_t(), ctypes.c_size_t()
    assert impl.rb_gl2_copy_trajectory(g, ctypes.byref(tp), ctypes.byref(tc), ctypes.byref(sp), ctypes.byref(sc)) == 0
    times = np.ctypeslib.as_array(tp, shape=(tc.value,)).copy() if tc.value else np.zeros(0)
    states = np.ctypeslib.as_array(sp, shape=(sc.value, 3 * N)).copy()
    impl.rb_free(ctypes.cast(tp, ctypes.c_void_p))
    impl.rb_free(ctypes.cast(sp, ctypes.c_void_p))
    return times, states


def _case(name):
    g = _gl2_golden()
    N, depth, amp, t0, t1, h, tol, maxit, fallback, halves = g[name + "/params"]
    return dict(N=int(N), depth=float(depth), t0=float(t0), t1=float(t1), h=float(h), tol=float(tol), maxit=int(maxit),
                fallback=bool(fallback), halves=int(halves), physics=str(g[name + "/physics"]), y0=g[name + "/y0"], T=g[name + "/T"],
                Y=g[name + "/Y"])


def _gl2_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_gl2.npz"))


@pytest.mark.parametrize("name", ["helium_film_N16", "helium_film_N16_backward", "helium_thin_N16_fallback", "water_N16", "helium_thin_N16_halving",
                                  "helium_thin_N16_fallback_halving"])
def test_implicit_host_logic_reproduces_reference_gl2_trajectories(impl, name):
    """rb_gl2_evolve as implicit.cu runs it (Newton iteration, Armijo search, step control, logging) over the oracle's RHS: the
    trajectory of the reference's own Python integrator (golden), to 1e-12 (the Newton systems are solved by a different LU)."""
    c = _case(name)
    s, j, g, _ = _gl2(impl, c)
    y0 = c["y0"].copy()
    assert impl.rb_gl2_initialize(g, _p(y0), 0) == 0
    assert impl.rb_gl2_evolve(g, c["t0"], c["t1"]) == 0, impl.rb_last_error()
    times, states = _trajectory(impl, g, c["N"])
    assert len(times) == len(c["T"]) and np.abs(times - c["T"]).max() <= 1e-15
    assert np.array_equal(states[0], c["y0"])
    assert np.abs(states - c["Y"]).max() <= 1e-12
    final = np.zeros(3 * c["N"])
    assert impl.rb_gl2_get_state(g, _p(final)) == 0 and np.array_equal(final, states[-1])
    st = impl._lib.rb_gl2_stats()
    assert impl.rb_gl2_get_stats(g, ctypes.byref(st)) == 0
    assert st.converged == 1 and st.steps_accepted == len(c["T"]) - 1
    assert (st.steps_halved > 0) == ("halving" in name)
    assert st.jacobians == 2 * st.linear_solves or c["fallback"]
    for h_, f_ in ((g, impl.rb_gl2_destroy), (j, impl.rb_jacobian_destroy), (s, impl.rb_destroy)):
        f_(h_)


def test_implicit_host_logic_options_steps_and_failures(impl):
    c = _case("helium_film_N16")
    # without a trajectory: the final state alone, no times; a second evolution starts again from options.stepSize
    s, j, g, o = _gl2(impl, c, trajectory=False)
    y0 = c["y0"].copy()
    assert impl.rb_gl2_initialize(g, _p(y0), 0) == 0
    assert impl.rb_gl2_evolve(g, c["t0"], c["t1"]) == 0
    times, states = _trajectory(impl, g, c["N"])
    assert len(times) == 0 and states.shape == (1, 3 * c["N"]) and np.abs(states[0] - c["Y"][-1]).max() <= 1e-12
    assert impl.rb_gl2_evolve(g, c["t1"], c["t0"]) == 0                 # symmetric scheme: back to the start
    back = np.zeros(3 * c["N"])
    impl.rb_gl2_get_state(g, _p(back))
    assert np.abs(back - c["y0"]).max() <= 1e-9
    st = impl._lib.rb_gl2_stats()
    impl.rb_gl2_get_stats(g, ctypes.byref(st))
    assert st.steps_accepted <= 2 * (len(c["T"]) - 1) + 1                # not thousands of slivers
    # single steps on caller-owned memory: converged -> advanced in place; an impossible request -> left alone
    y1 = c["y0"].copy()
    assert impl.rb_gl2_initialize(g, _p(y1), 1) == 0
    ok = ctypes.c_int()
    assert impl.rb_gl2_step(g, c["h"], ctypes.byref(ok)) == 0 and ok.value == 1
    assert np.abs(y1 - c["Y"][1]).max() <= 1e-12
    o.newtonTolerance, o.maxNewtonIterations = 1e-300, 1
    assert impl.rb_gl2_set_options(g, ctypes.byref(o)) == 0
    before = y1.copy()
    assert impl.rb_gl2_step(g, c["h"], ctypes.byref(ok)) == 0 and ok.value == 0
    assert np.array_equal(y1, before)
    # an evolution that cannot converge reports it the way the reference throws it
    assert impl.rb_gl2_evolve(g, 0.0, 0.1) == -1
    assert b"failed to converge" in impl.rb_last_error()
    # options the line search could not terminate with are refused
    o.backtrack = 1.0
    assert impl.rb_gl2_set_options(g, ctypes.byref(o)) == -1 and b"backtrack" in impl.rb_last_error()
    o.backtrack = 0.5
    # a step size that can never reach the end time is refused (the reference's `< 0` test lets 0 through, into an endless loop)
    o.stepSize = 0.0
    impl.rb_gl2_set_options(g, ctypes.byref(o))
    assert impl.rb_gl2_evolve(g, 0.0, 0.1) == -1 and b"Step size must be positive" in impl.rb_last_error()
    o.stepSize = c["h"]
    # inner solves that do not converge make every slope unacceptable: same outcome, no hang
    o.newtonTolerance, o.maxNewtonIterations = 1e-10, 20
    impl.rb_gl2_set_options(g, ctypes.byref(o))
    impl.emu_force_unconverged(1)
    try:
        assert impl.rb_gl2_evolve(g, 0.0, 0.1) == -1
    finally:
        impl.emu_force_unconverged(0)
    for h_, f_ in ((g, impl.rb_gl2_destroy), (j, impl.rb_jacobian_destroy), (s, impl.rb_destroy)):
        f_(h_)


def test_implicit_host_logic_legacy_exports(impl):
    """calculateJacobian / calculatePerturbedStates256 / integrateSimulationGL2 as implicit.cu implements them (SI in, packing, malloc'd
    outputs) over the oracle's RHS, against the oracle on the nondimensionalised problem."""
    _lib = impl._lib
    N, L, d_si, rho_si = 8, 1e-6, 15e-9, 145.0
    op = ro.adimensionalize_properties(ro.ProblemProperties(L=L, rho=rho_si, kappa=0.0, depth=d_si))
    y0 = _film(N, op.depth, 0.1)
    jac = np.zeros(9 * N * N)
    assert impl.calculateJacobian(_p(y0), _p(jac), L, rho_si, 0.0, d_si, 1e-6, N) == 0
    assert np.array_equal(jac.reshape(3 * N, 3 * N).T, ro.jacobian_fd(y0, N, op, "helium", 1e-6))
    rng = np.random.default_rng(4)
    x, y, phi = rng.standard_normal(256), rng.standard_normal(256), rng.standard_normal(256)
    zp = np.zeros(6 * 256 * 256, np.complex128)
    assert impl.calculatePerturbedStates256(_p(x), _p(y), _p(phi), zp.ctypes.data_as(ctypes.c_void_p), L, rho_si, 0.0, d_si, 1e-6) == 0
    assert np.array_equal(zp, ro.perturbed_states(np.concatenate([x + 1j * y, phi + 0j]), 256, 1e-6))
    sp = _lib.SimProperties(L=L, rho=rho_si, kappa=0.0, depth=d_si, use_expansions=False, expansion_order=1, infinite_depth=False)
    go = _lib.GaussLegendreOptions(t0=0.0, t1=0.25, stepSize=0.1, newtonTolerance=1e-10, maxNewtonIterations=20,
                                   allowSimplifiedFallback=False, returnTrajectory=True, armijo_c=1e-4, backtrack=0.5, minAlpha=1e-6,
                                   maxStepsHalves=6)
    so, to = D(), D()
    sc, tc = ctypes.c_size_t(), ctypes.c_size_t()
    init = y0.copy()
    assert impl.integrateSimulationGL2(_p(init), ctypes.byref(so), ctypes.byref(sc), ctypes.byref(to), ctypes.byref(tc),
                                       ctypes.byref(sp), ctypes.byref(go), N) == 0, impl.rb_last_error()
    states = np.ctypeslib.as_array(so, shape=(sc.value, 3 * N)).copy()
    times = np.ctypeslib.as_array(to, shape=(tc.value,)).copy()
    impl.integrateSimulationGL2_freeMemory(so, to)
    To, Yo = ro.gl2_integrate(lambda v: ro.real_rhs(v, N, op, "helium"), lambda v: ro.jacobian_fd(v, N, op, "helium", 1e-6), y0, 0.0, 0.25,
                              ro.GaussLegendre2Options(stepSize=0.1))
    assert len(times) == len(To) and np.abs(times - To).max() <= 1e-15
    assert states.shape == Yo.shape and np.abs(states - Yo).max() <= 1e-12


# ---- the named functions / kernels of include/cusuperhelium_compat.cuh (tests/cpp/emu_compat.cpp) ------------------------------------
@pytest.fixture(scope="module")
def compat():
    src = os.path.join(ROOT, "tests", "cpp", "emu_compat.cpp")
    out = os.path.join(ROOT, "tests", "cpp", "_build", "libemu_compat.so")
    deps = [src, os.path.join(ROOT, "tests", "cpp", "cuda_emu.h"), os.path.join(ROOT, "include", "cusuperhelium_compat.cuh"),
            os.path.join(ROOT, "include", "roberts_b200_device.cuh"), os.path.join(ROOT, "include", "roberts_b200.h")]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-I",
                               os.path.join(ROOT, "tests", "cpp"), "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, src, "-o", out])
    return ctypes.CDLL(out)


def _c(a):
    return np.ascontiguousarray(a, np.complex128)


def test_compat_complex_functions_with_the_reference_fixtures(compat):
    """sin / cos / cotangent_complex / cot / cotangent_green_function under the reference's names, on the fixtures of
    T/ComplexFunctionsTests.cuh (the diagonal of the first period; Zk = pi (1 + i), Zj = Zk + 0.1 (1 + i) with the reference's
    50-digit value and its 1e-13) and on known answers computed with mpmath at 40 digits."""
    N = 256
    v = 2 * np.pi / (N + 1) * (np.arange(N) + 1)
    z = _c(v + 1j * v)
    out = np.zeros_like(z)
    compat.emuc_sin(_p(z.view(np.float64)), _p(out.view(np.float64)), N)
    assert (np.abs(out - np.sin(z)) / np.abs(np.sin(z))).max() <= 4e-15
    compat.emuc_cos(_p(z.view(np.float64)), _p(out.view(np.float64)), N)
    assert (np.abs(out - np.cos(z)) / np.abs(np.cos(z))).max() <= 4e-15
    ref = np.cos(z) / np.sin(z)
    compat.emuc_cotangent_complex(_p(z.view(np.float64)), _p(out.view(np.float64)), N)
    assert (np.abs(out - ref) / np.abs(ref)).max() <= 1e-14
    out2 = np.zeros_like(z)
    compat.emuc_cot(_p(z.view(np.float64)), _p(out2.view(np.float64)), N)
    assert np.array_equal(out, out2)
    q = _c([1e-3 + 2e-3j, 0.7 - 1.3j, 3.0 + 40.0j])
    expect = np.array([199.99966666691110686 - 400.00066666662221382j, 0.14933231644907048137 + 1.0145011371447459055j,
                       -1.0086068994196989300e-35 - 1.0j])
    o3 = np.zeros_like(q)
    compat.emuc_cotangent_complex(_p(q.view(np.float64)), _p(o3.view(np.float64)), 3)
    assert (np.abs(o3 - expect) / np.abs(expect)).max() <= 1e-15
    assert abs(o3[2].real - expect[2].real) <= 1e-15 * abs(expect[2].real)          # no cancellation where cosh - cos would lose everything
    M = 16
    zk = _c(np.full(M, np.pi + 1j * np.pi))
    zj = _c(np.full(M, (np.pi + 0.1) + 1j * (np.pi + 0.1)))
    g = np.zeros(M, np.complex128)
    compat.emuc_green(_p(zk.view(np.float64)), _p(zj.view(np.float64)), _p(g.view(np.float64)), M)
    assert np.abs(g.real + 9.9833388915330681153509188355277398344111330657474).max() <= 1e-13
    assert np.abs(g.imag - 10.016672219575397493791065794281138271298250914604).max() <= 1e-13


def test_compat_precision_math_with_the_reference_fixtures(compat):
    """PrecisionMath::fastPreciseInvSub and c_twoDiff on the fixtures of T/ComplexFunctionsTests.cuh:247-405, plus what those
    fixtures do not reach: a difference with catastrophic cancellation, where the double-double path must beat plain FP64."""
    from fractions import Fraction
    M = 16
    i = np.arange(M, dtype=np.float64)
    zk = _c(i + 1j * i)
    zj = _c((i - np.pi) + 1j * (i - np.pi))
    r = np.zeros(M, np.complex128)
    compat.emuc_inv_sub(_p(zk.view(np.float64)), _p(zj.view(np.float64)), _p(r.view(np.float64)), M)
    e = 0.15915494309189533576888376337251436203445964574046
    assert np.abs(r.real - e).max() <= 1e-13 and np.abs(r.imag + e).max() <= 1e-13
    zj3 = _c((i + np.pi) + 1j * (i + np.pi))
    hi, lo = np.zeros(M, np.complex128), np.zeros(M, np.complex128)
    compat.emuc_two_diff(_p(zk.view(np.float64)), _p(zj3.view(np.float64)), _p(hi.view(np.float64)), _p(lo.view(np.float64)), M)
    assert np.array_equal(hi.real, i - (i + np.pi)) and np.array_equal(hi.imag, i - (i + np.pi))
    assert np.abs(hi.real + np.pi).max() <= 4e-15 and not lo.real.any() and not lo.imag.any()
    # hi + lo is the exact difference whatever the operands (error-free transformation): checked in rational arithmetic
    a = _c([1.0 + 3.0j, 1e16 + 1.0j, 0.1 + 0.2j])
    b = _c([1e-20 + 1e-30j, 1.0 + 1e16j, 0.3 - 0.7j])
    hi, lo = np.zeros(3, np.complex128), np.zeros(3, np.complex128)
    compat.emuc_two_diff(_p(a.view(np.float64)), _p(b.view(np.float64)), _p(hi.view(np.float64)), _p(lo.view(np.float64)), 3)
    for k in range(3):
        assert Fraction(hi[k].real) + Fraction(lo[k].real) == Fraction(a[k].real) - Fraction(b[k].real)
        assert Fraction(hi[k].imag) + Fraction(lo[k].imag) == Fraction(a[k].imag) - Fraction(b[k].imag)
    # 1 / (Z1 - Z2) for nearly equal operands whose difference is not exactly representable in one double
    z1 = _c([(1.0 + 2.0 ** -30) + (1.0 + 2.0 ** -29) * 1j])
    z2 = _c([2.0 ** -60 + (2.0 ** -61) * 1j])
    r = np.zeros(1, np.complex128)
    compat.emuc_inv_sub(_p(z1.view(np.float64)), _p(z2.view(np.float64)), _p(r.view(np.float64)), 1)
    dre = Fraction(z1[0].real) - Fraction(z2[0].real)
    dim = Fraction(z1[0].imag) - Fraction(z2[0].imag)
    den = dre * dre + dim * dim
    assert abs(Fraction(r[0].real) - dre / den) <= Fraction(1, 2 ** 51) * abs(dre / den)
    assert abs(Fraction(r[0].imag) + dim / den) <= Fraction(1, 2 ** 51) * abs(dim / den)


@pytest.mark.parametrize("N", [4, 64, 130])
def test_compat_jacobian_kernels_in_the_reference_geometry(compat, N):
    """createInitialState<<<N, 1>>>, createInitialBatchedZ<<<(ceil(2N/256), 3N), 256>>> (one thread per entry of the 2N-entry state, the
    reference's geometry, T/MatrixMTests.cuh:334-364) and createJacobianMatrixFromPerturbedRhs against the oracle, bit for bit."""
    rng = np.random.default_rng(N)
    y = rng.standard_normal(3 * N)
    st = np.zeros(2 * N, np.complex128)
    compat.emuc_create_initial_state(_p(y), _p(st.view(np.float64)), ctypes.c_size_t(N))
    assert np.array_equal(st, ro.real_to_complex_state(y, N))
    zb = np.full(6 * N * N, np.nan + 0j)
    compat.emuc_create_initial_batched_z(_p(st.view(np.float64)), _p(zb.view(np.float64)), ctypes.c_double(1e-6), ctypes.c_size_t(N))
    assert np.array_equal(zb, ro.perturbed_states(st, N, 1e-6))
    pos = rng.standard_normal(6 * N * N) + 1j * rng.standard_normal(6 * N * N)
    neg = rng.standard_normal(6 * N * N) + 1j * rng.standard_normal(6 * N * N)
    C = np.full(9 * N * N, np.nan)
    compat.emuc_jacobian_from_perturbed(_p(pos.view(np.float64)), _p(neg.view(np.float64)), _p(C), ctypes.c_size_t(N), ctypes.c_double(1e-6))
    assert np.array_equal(C, ro.jacobian_from_perturbed(pos, neg, N, 1e-6).ravel(order="F"))
