"""Extract the known-answer tables of the reference's ODE-solver tests into a small fixture.

Run in the build container only (needs /root/reference):
    python tests/golden/make_ode_golden.py
Reads CuSuperHelium/CuSuperHelium.Tests/ODESolverTests.cuh (TEST(ODE_Solvers, RK4) :73-246 and TEST(ODE_Solvers, RK45) :248-421):
the expected_x / expected_y tables (SciPy solve_ivp values of dz/dt = i z at t = 10 for z_j(0) = 2 pi j/256 + i sin(2 pi j 0.01),
which the reference compares with EXPECT_NEAR(.., 1e-2)).  Writes tests/golden/ref_ode_rotation.npz; nothing under
/root/reference is read at test time.
"""
import os
import re

import numpy as np

SRC = "/root/reference/CuSuperHelium/CuSuperHelium.Tests/ODESolverTests.cuh"
HERE = os.path.dirname(os.path.abspath(__file__))


def tables(text):
    out = []
    for m in re.finditer(r"std::vector<double>\s+(expected_[xy])\s*=\s*\{([^}]*)\}", text):
        out.append((m.group(1), np.array([float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()])))
    return out


def main():
    text = open(SRC).read()
    i45 = text.index("TEST(ODE_Solvers, RK45)")
    rk4, rk45 = dict(tables(text[:i45])), dict(tables(text[i45:]))
    assert all(len(v) == 256 for v in list(rk4.values()) + list(rk45.values()))
    np.savez(os.path.join(HERE, "ref_ode_rotation.npz"), rk4_x=rk4["expected_x"], rk4_y=rk4["expected_y"],
             rk45_x=rk45["expected_x"], rk45_y=rk45["expected_y"])
    print("wrote ref_ode_rotation.npz", {k: v[:3] for k, v in rk45.items()})


if __name__ == "__main__":
    main()
