"""Generate golden vectors from the REFERENCE'S OWN CUDA PATH (needs a GPU and oracle/_ref/libcusuperhelium_ref.so, i.e. the
reference's classes compiled unmodified by oracle/build_ref.py):

    gpurun -- 'python tests/golden/make_ref_cuda_golden.py gpurun_out/ref_cuda_golden.npz'     (then copied to tests/golden/)

For every small case of tests/ref_cases.py (N <= 256) the file holds the input state and what the reference returned: the RHS
and its intermediates (a, upper-fluid velocity, Z', Z'', Phi', energies) or the state after the RK4 steps.  These pin the CPU
oracle -- including its finite-depth helium mode, for which the reference has no CPU statement, test or vector of its own --
in the CPU test tier (tests/test_oracle.py::test_oracle_matches_reference_cuda_golden); nothing under /root/reference or
oracle/_ref is needed at test time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import ref_cases
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_cuda_golden.npz")
    cases = [c for c in ref_cases.CASES if c["N"] <= 256]
    res = ref_cases.reference_results(cases)
    out = {}
    names = []
    for c in cases:
        r = res[c["name"]]
        if "error" in r:
            print("no result for", c["name"], r["error"])
            continue
        names.append(c["name"])
        out[c["name"] + "/state0"] = ref_cases.state_of(c)
        for k, v in r.items():
            out[c["name"] + "/" + k] = np.asarray(v)
    out["names"] = np.array(names)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path, "with", len(names), "cases")


if __name__ == "__main__":
    main()
