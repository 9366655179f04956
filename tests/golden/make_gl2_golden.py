"""Golden vectors for the implicit Gauss-Legendre-2 integrator from the REFERENCE's own Python statement of it.

Run in the build container only (needs /root/reference):
    python tests/golden/make_gl2_golden.py
Writes tests/golden/ref_gl2.npz.  The reference module imported (unmodified, from where it lies) is
CuSuperHelium/Python/integration/gauss_legendre.py: `integrate_gl2` / `gauss_legendre_s2_step` (damped Newton with Armijo
backtracking, simplified-Newton fallback, step halving).  In the reference its f and J call the Windows DLL
(calculate_rhs256_from_vectors / calculate_jacobian); here they are the oracle's restatements of those two exports
(oracle/roberts_oracle.py: real_rhs, jacobian_fd), so what the file pins is the oracle's restatement of the INTEGRATOR
(gl2_step / gl2_integrate) on the same f and J.  Nothing under /root/reference is read at test time.
"""
import os
import sys

import numpy as np

REF = "/root/reference/CuSuperHelium/Python/integration"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

CASES = {
    # name: (physics, N, depth, amplitude (fraction of depth), t0, t1, h, tol, maxit, fallback, max halves)
    "helium_film_N16": ("helium", 16, 0.3, 0.05, 0.0, 0.5, 0.05, 1e-10, 20, False, 6),
    "helium_film_N16_backward": ("helium", 16, 0.3, 0.05, 0.5, 0.0, 0.05, 1e-10, 20, False, 6),
    "helium_thin_N16_fallback": ("helium", 16, 0.0942478, 0.1, 0.0, 0.23, 0.1, 1e-11, 12, True, 6),
    "water_N16": ("water", 16, 1.0, 0.1, 0.0, 0.3, 0.1, 1e-10, 20, False, 6),
    # a step too long for 8 Newton iterations: the first attempt fails, the halved step converges and is kept
    "helium_thin_N16_halving": ("helium", 16, 0.0942478, 0.5, 0.0, 6.0, 6.0, 1e-11, 8, True, 6),
    # the line search gives up, the Jacobian is frozen at the base point (simplified Newton), the step is halved all the same
    "helium_thin_N16_fallback_halving": ("helium", 16, 0.0942478, 0.6, 0.0, 3.0, 3.0, 1e-11, 8, True, 6),
}


def initial_state(N, depth, amp):
    a = 2 * np.pi * np.arange(N) / N
    return np.concatenate([a - 0.3 * amp * depth * np.sin(a), amp * depth * np.cos(a), 0.2 * amp * depth * np.sin(a)])


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import gauss_legendre as gl  # noqa: the reference's module
    from oracle import roberts_oracle as ro

    out = {"names": np.array(list(CASES))}
    for name, (physics, N, depth, amp, t0, t1, h, tol, maxit, fallback, halves) in CASES.items():
        props = ro.ProblemProperties(rho=0.0 if physics == "water" else 1.0, depth=depth)
        y0 = initial_state(N, depth, amp)
        f = lambda y: ro.real_rhs(y, N, props, physics)            # noqa: E731
        J = lambda y: ro.jacobian_fd(y, N, props, physics, 1e-6)   # noqa: E731
        newton = gl.NewtonOptions(tol=tol, maxit=maxit, armijo_c=1e-4, backtrack=0.5, min_alpha=1e-6,
                                  allow_simplified_fallback=fallback)
        T, Y = gl.integrate_gl2(f, J, y0, t0, t1, h, newton, return_trajectory=True, max_step_halves=halves, show_progress=False)
        out[name + "/y0"] = y0
        out[name + "/T"] = T
        out[name + "/Y"] = Y
        out[name + "/params"] = np.array([N, depth, amp, t0, t1, h, tol, maxit, float(fallback), halves])
        out[name + "/physics"] = np.array(physics)
        print(name, "steps", len(T) - 1, "final |dy|", np.abs(Y[-1] - y0).max())
    np.savez_compressed(os.path.join(HERE, "ref_gl2.npz"), **out)


if __name__ == "__main__":
    main()
