"""Generate golden vectors for the Roberts BIE path from the REFERENCE's own Python code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/ref_water_rhs.npz.  The reference modules imported are
CuSuperHelium/Python/WaterIntegralCalculator.py, Derivatives.py and StokeWaves.py
(unmodified, imported from where they lie).  Nothing under /root/reference is
read at test time; only the committed .npz travels.
"""
import os
import sys

import numpy as np

REF = "/root/reference/CuSuperHelium/Python"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from WaterIntegralCalculator import WaterIntegralCalculator  # noqa
    import StokeWaves  # noqa
    import Derivatives  # noqa

    out = {}
    for N, h in ((64, 0.1), (64, 0.4), (32, 0.25), (128, 0.3)):
        j = 2 * np.pi * np.arange(N) / N
        X = StokeWaves.X(j, 0.0, h, 1.0)
        Y = StokeWaves.Y(j, 0.0, h, 1.0)
        Phi = StokeWaves.Phi(j, 0.0, h, 1.0)
        w = WaterIntegralCalculator()
        rhs = w.runTimeStep(0.0, np.hstack((X, Y, Phi)))
        tag = f"N{N}_h{h}"
        out[tag + "_state"] = np.hstack((X, Y, Phi))
        out[tag + "_rhs"] = rhs
        out[tag + "_a"] = w.a
        out[tag + "_M"] = w.M
        out[tag + "_Xp"] = w.Xp
        out[tag + "_Yp"] = w.Yp
        out[tag + "_Xpp"] = w.Xpp
        out[tag + "_Ypp"] = w.Ypp
        out[tag + "_Phip"] = w.Phip
        out[tag + "_ap"] = w.ap
        V1, V2 = WaterIntegralCalculator._WaterIntegralCalculator__createVelocityMatrices(X, Y, w.Xp, w.Yp, w.Xpp, w.Ypp)
        out[tag + "_V1"] = V1
        out[tag + "_V2"] = V2
    # a perturbed, non-analytic surface (two modes + phase) so that the pin is not trochoid-only
    N = 64
    j = 2 * np.pi * np.arange(N) / N
    X = j + 0.05 * np.sin(2 * j + 0.3)
    Y = 0.2 * np.cos(j) + 0.03 * np.sin(3 * j + 1.0)
    Phi = 0.1 * np.sin(j + 0.5) + 0.02 * np.cos(2 * j)
    w = WaterIntegralCalculator()
    out["N64_pert_state"] = np.hstack((X, Y, Phi))
    out["N64_pert_rhs"] = w.runTimeStep(0.0, np.hstack((X, Y, Phi)))
    out["N64_pert_a"] = w.a
    np.savez_compressed(os.path.join(HERE, "ref_water_rhs.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_water_rhs.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
