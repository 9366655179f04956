"""Build libroberts_b200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo snapshot).

    python -m superfluid_dynamics_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libroberts_b200.so")
SOURCES = ["pair_kernels.cu", "pair_kernels2.cu", "pair_kernels3.cu", "spectral.cu", "dense_kernels.cu", "stepper_kernels.cu", "rk45_kernels.cu", "krylov_kernels.cu", "drive_kernels.cu", "implicit.cu", "lu_kernels.cu", "solver.cu", "stepper.cu", "comm.cu", "probes.cu", "exports.cu", "drive.cu", "rk45.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-diag-suppress", "128"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, "internal.cuh"), os.path.join(CSRC, "host.cuh"), os.path.join(CSRC, "implicit_kernels.cuh"), os.path.join(CSRC, "launch.cuh"), os.path.join(HERE, "..", "include", "roberts_b200.h"),
               os.path.join(HERE, "..", "include", "roberts_b200_device.cuh")]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcufft", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


def build_compat_test(verbose=True):
    """tests/cpp/compat_test.cu: the reference's kernel tests written against include/cusuperhelium_compat.cuh."""
    root = os.path.abspath(os.path.join(HERE, ".."))
    src = os.path.join(root, "tests", "cpp", "compat_test.cu")
    exe = os.path.join(LIBDIR, "compat_test")
    deps = [src, os.path.join(root, "include", "cusuperhelium_compat.cuh"), os.path.join(root, "include", "roberts_b200.h"),
            os.path.join(root, "include", "roberts_b200_device.cuh"), LIB]
    if _stale(exe, deps):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-lineinfo", "-std=c++17", "-I",
               os.path.join(root, "include"), src, "-o", exe, "-L", LIBDIR, "-lroberts_b200", "-Xlinker", "-rpath=$ORIGIN",
               "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_compat_test())
