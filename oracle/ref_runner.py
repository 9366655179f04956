"""TEST INFRASTRUCTURE (oracle/): run the compiled reference (oracle/_ref/libcusuperhelium_ref.so, the reference's own CUDA
classes, see oracle/build_ref.py) in a CHILD process and hand its outputs back as NumPy arrays.

A child process because the reference's error handling is `exit(EXIT_FAILURE)` (L/utilities.cuh:408-444): a failure inside it
must not take the test runner or bench.py down with it.  Jobs are dicts:

    {"op": "rhs", "kind": "water"|"helium"|"helium_inf", "N": 256, "props": {"rho": 0.0, ...}, "state": complex128[2N]}
        -> rhs complex128[2N], a float64[N], vel_upper, zp, zpp complex128[N], phi_prime float64[N], energies float64[4]
           (kinetic, potential, surface, volume flux: the reference's EnergyContainer / VolumeFlux after that RHS)
    {"op": "aug_rhs" | "aug_rk4", "N": .., "props": .., "opto": {OptomechanicalVariables fields}, "state": complex128[3N] [, dt, steps]}
        -> the augmented optomechanical system [Z | Phi | D] (HeliumDrivenAutonomousProblem + AugmentedBoundaryIntegrator):
           rhs complex128[3N], or the state after the steps and the seconds they took
    {"op": "timed_rk4", "N": .., "props": .., "opto": .., "state": complex128[2N], "t0": 0.0, "dt": .., "steps": ..}
        -> the explicitly time-dependent drive (HeliumWithOptomechanicalDrivingProblem + TimedBoundaryIntegrator +
           RungeKuttaStepper::runEvolution): state after the steps, seconds
    {"op": "rk4", ..., "dt": 1e-3, "steps": 100, "warmup": 0}
        -> state complex128[2N] after warmup+steps steps, seconds (host clock around the last `steps` steps, device
           synchronised on both sides), energies of the last RHS evaluated

    results = run_jobs([job, ...])      # list of dicts, {"error": "..."} for a job that did not finish

Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this module.
"""
import ctypes
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libcusuperhelium_ref.so")
KINDS = {"water": 0, "helium": 1, "helium_inf": 2}


class _Props(ctypes.Structure):
    _fields_ = [("rho", ctypes.c_double), ("kappa", ctypes.c_double), ("depth", ctypes.c_double), ("U", ctypes.c_double),
                ("L", ctypes.c_double), ("use_expansions", ctypes.c_int), ("expansion_order", ctypes.c_int),
                ("infinite_depth", ctypes.c_int)]


class _Opto(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("detuning", "gamma", "G", "Tau", "max_intensity", "initial_time", "location_x0_mode",
                                               "sigma_optical_mode", "Beta", "DampingStrength")]


def available():
    return os.path.exists(LIB)


def _props(d):
    d = dict(d or {})
    return _Props(float(d.get("rho", 0.0)), float(d.get("kappa", 0.0)), float(d.get("depth", 1.0)), float(d.get("U", 0.0)),
                  float(d.get("L", 1.0)), int(bool(d.get("use_expansions", False))), int(d.get("expansion_order", 1)),
                  int(bool(d.get("infinite_depth", False))))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _load():
    lib = ctypes.CDLL(LIB)
    D = ctypes.POINTER(ctypes.c_double)
    lib.ref_rhs.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_Props), D, D, D, D, D, D, D, D]
    lib.ref_rhs.restype = ctypes.c_int
    lib.ref_rk4.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(_Props), D, ctypes.c_double, ctypes.c_int, ctypes.c_int, D, D]
    lib.ref_rk4.restype = ctypes.c_int
    if hasattr(lib, "ref_timed"):
        lib.ref_timed.argtypes = [ctypes.c_int, ctypes.POINTER(_Props), ctypes.POINTER(_Opto), D, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_int, D]
        lib.ref_timed.restype = ctypes.c_int
    if hasattr(lib, "ref_augmented"):
        lib.ref_augmented.argtypes = [ctypes.c_int, ctypes.POINTER(_Props), ctypes.POINTER(_Opto), D, D, ctypes.c_double, ctypes.c_int,
                                      ctypes.c_int, D]
        lib.ref_augmented.restype = ctypes.c_int
    lib.ref_num_sizes.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.ref_num_sizes.restype = ctypes.c_int
    return lib


def sizes():
    """The N the reference was instantiated for (it fixes N at compile time)."""
    lib = _load()
    buf = (ctypes.c_int * 32)()
    n = lib.ref_num_sizes(buf, 32)
    return [buf[i] for i in range(n)]


def _run_one(lib, job):
    N = int(job["N"])
    kind = KINDS[job["kind"]]
    p = _props(job.get("props"))
    state = np.ascontiguousarray(job["state"], dtype=np.complex128).copy()
    if job["op"] == "timed_rk4":
        assert state.size == 2 * N
        o = _Opto(*[float(job["opto"][n]) for n, _ in _Opto._fields_])
        sec = ctypes.c_double(0.0)
        rc = lib.ref_timed(N, ctypes.byref(p), ctypes.byref(o), _dp(state.view(np.float64)), float(job.get("t0", 0.0)),
                           float(job["dt"]), int(job["steps"]), ctypes.byref(sec))
        if rc != 0:
            return {"error": f"ref_timed returned {rc}"}
        return dict(state=state, seconds=sec.value)
    if job["op"] in ("aug_rhs", "aug_rk4"):
        assert state.size == 3 * N
        o = _Opto(*[float(job["opto"][n]) for n, _ in _Opto._fields_])
        out = np.zeros(3 * N, np.complex128)
        sec = ctypes.c_double(0.0)
        steps = int(job["steps"]) if job["op"] == "aug_rk4" else 0
        rc = lib.ref_augmented(N, ctypes.byref(p), ctypes.byref(o), _dp(state.view(np.float64)), _dp(out.view(np.float64)),
                               float(job.get("dt", 0.0)), int(job.get("warmup", 0)), steps, ctypes.byref(sec))
        if rc != 0:
            return {"error": f"ref_augmented returned {rc}"}
        return dict(rhs=out) if steps == 0 else dict(state=out, seconds=sec.value)
    assert state.size == 2 * N
    en = np.zeros(4)
    if job["op"] == "rhs":
        rhs = np.zeros(2 * N, np.complex128)
        a, pp = np.zeros(N), np.zeros(N)
        vu, zp, zpp = (np.zeros(N, np.complex128) for _ in range(3))
        rc = lib.ref_rhs(kind, N, ctypes.byref(p), _dp(state.view(np.float64)), _dp(rhs.view(np.float64)), _dp(a),
                         _dp(vu.view(np.float64)), _dp(zp.view(np.float64)), _dp(zpp.view(np.float64)), _dp(pp), _dp(en))
        if rc != 0:
            return {"error": f"ref_rhs returned {rc}"}
        return dict(rhs=rhs, a=a, vel_upper=vu, zp=zp, zpp=zpp, phi_prime=pp, energies=en)
    if job["op"] == "rk4":
        sec = ctypes.c_double(0.0)
        rc = lib.ref_rk4(kind, N, ctypes.byref(p), _dp(state.view(np.float64)), float(job["dt"]), int(job.get("warmup", 0)),
                         int(job["steps"]), ctypes.byref(sec), _dp(en))
        if rc != 0:
            return {"error": f"ref_rk4 returned {rc}"}
        return dict(state=state, seconds=sec.value, energies=en)
    return {"error": "unknown op " + str(job["op"])}


def _child(inp, outdir):
    jobs = pickle.load(open(inp, "rb"))
    lib = _load()
    for i, job in enumerate(jobs):
        try:
            res = _run_one(lib, job)
        except Exception as e:  # noqa: BLE001
            res = {"error": repr(e)}
        # one file per finished job: what completed survives a later exit() inside the reference
        with open(os.path.join(outdir, f"{i}.pkl.tmp"), "wb") as f:
            pickle.dump(res, f)
        os.replace(os.path.join(outdir, f"{i}.pkl.tmp"), os.path.join(outdir, f"{i}.pkl"))


def run_jobs(jobs, timeout=600, device=None):
    if not available():
        return [{"error": "oracle/_ref/libcusuperhelium_ref.so not built (python -m oracle.build_ref needs /root/reference)"}] * len(jobs)
    with tempfile.TemporaryDirectory() as d:
        inp = os.path.join(d, "jobs.pkl")
        pickle.dump(jobs, open(inp, "wb"))
        env = dict(os.environ)
        if device is not None:
            env["CUDA_VISIBLE_DEVICES"] = str(device)
        note = ""
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", inp, d], env=env, timeout=timeout,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            if r.returncode != 0:
                note = f"child exit {r.returncode}: {(r.stderr or '')[-400:]}"
        except subprocess.TimeoutExpired:
            note = f"child timed out after {timeout} s"
        out = []
        for i in range(len(jobs)):
            f = os.path.join(d, f"{i}.pkl")
            out.append(pickle.load(open(f, "rb")) if os.path.exists(f) else {"error": note or "no result"})
        return out


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--child":
        _child(sys.argv[2], sys.argv[3])
    else:
        print("library:", LIB, "built" if available() else "absent")
        if available():
            print("sizes:", sizes())
