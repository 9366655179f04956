"""TEST INFRASTRUCTURE (oracle/): compile the reference's own CUDA path into oracle/_ref/libcusuperhelium_ref.so.

    python -m oracle.build_ref [--force]

The sources are compiled where they lie under /root/reference (read-only, -I), nothing is copied into the repo; the only
translation unit is oracle/ref_cuda/harness.cu (ours: instantiates the reference's classes for a few N, see its header for the
two accommodations a conforming compiler needs).  Output only under oracle/_ref/ (git-ignored, travels to the GPU box).  The GPU
box has no /root/reference: there the prebuilt library is used as is, and build() is a no-op when the sources are absent.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/CuSuperHelium/CuSuperHelium"
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libcusuperhelium_ref.so")
HARNESS = os.path.join(HERE, "ref_cuda", "harness.cu")
SHIMS = os.path.join(HERE, "ref_cuda", "shims")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def available():
    return os.path.exists(LIB)


def build(force=False, verbose=True):
    """Returns the library path, or None when neither the sources nor a prebuilt library exist."""
    if not os.path.isdir(REF_SRC):
        return LIB if available() else None
    deps = [HARNESS, os.path.join(SHIMS, "matplotlibcpp.h"), __file__]
    if not force and available() and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-lineinfo", "-std=c++20", "-w",
           "-I", SHIMS, "-I", REF_SRC, "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", HARNESS, "-o", LIB,
           "-lcufft", "-lcublas", "-lcusolver", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
