"""python tests/gpu_reference_report.py [out.json] -- run every case of tests/ref_cases.py (this repo's CUDA path against the
compiled reference, same GPU) and write the measured differences and the reference's own RK4 step rates as JSON.  Diagnostic
companion of tests/test_gpu_reference.py; the committed copy of its output lives under profiles/."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "reference_parity.json")
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    import torch
    import ref_cases
    from oracle import ref_runner
    from superfluid_dynamics_b200 import api, build
    build.build(verbose=False)
    report = {"library": ref_runner.LIB, "available": ref_runner.available(), "cases": {}}
    t0 = time.time()
    ref = ref_cases.reference_results()
    report["reference_child_seconds"] = time.time() - t0
    for c in ref_cases.CASES:
        r = ref[c["name"]]
        if "error" in r:
            report["cases"][c["name"]] = {"error": r["error"]}
            continue
        try:
            report["cases"][c["name"]] = ref_cases.measure(api, c, r, torch)
        except Exception as e:  # noqa: BLE001
            report["cases"][c["name"]] = {"error": "ours: " + repr(e)}
        json.dump(report, open(out_path, "w"), indent=1)
    report["total_seconds"] = time.time() - t0
    json.dump(report, open(out_path, "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
