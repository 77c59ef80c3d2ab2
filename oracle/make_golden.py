"""Generate tests/golden/*.json — run from the repo root: `python oracle/make_golden.py`.

1. Copies the reference's committed regression series (golden vectors, not code) out of its XML files
   (needs /root/reference; only available in the build container).
2. Runs the oracle on the reference's default dam-break cases and stores the energy series it produces,
   so the CPU test-suite can (a) check them against the reference series with the reference's own DTW
   criterion and (b) re-run a short prefix to prove the fixture still comes from the current oracle.
"""
import json
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from sphinxsys_b200 import cases  # noqa: E402

REF = "/root/reference/tests"
OUT = os.path.join(ROOT, "tests", "golden")


def read_series(path):
    s = open(path).read()
    vals = re.findall(r'snapshot_(\d+)="([^"]+)"', s)
    return [float(v) for _, v in sorted(vals, key=lambda x: int(x[0]))]


def read_threshold(path):
    return float(re.search(r'TotalMechanicalEnergy_0="([^"]+)"', open(path).read()).group(1))


def reference_goldens():
    out = {}
    specs = {
        "2d_dambreak_legacy": ("2d_examples/test_2d_dambreak/regression_test_tool", [0, 6, 11]),
        "3d_dambreak_ck_sycl": ("tests_sycl/3d_examples/test_3d_dambreak_sycl/regression_test_tool", [0, 5, 10]),
        "3d_dambreak_legacy": ("3d_examples/test_3d_dambreak/regression_test_tool", [0, 3, 6]),
    }
    for name, (d, runs) in specs.items():
        base = os.path.join(REF, d)
        out[name] = {
            "source": d,
            "dtw_threshold": read_threshold(os.path.join(base, "WaterBody_TotalMechanicalEnergy_dtwdistance.xml")),
            "runs": {str(r): read_series(os.path.join(base, f"WaterBody_TotalMechanicalEnergy_Run_{r}_result.xml"))
                     for r in runs},
        }
    return out


def oracle_series():
    out = {}
    t0 = time.time()
    c2 = cases.dam_break(dim=2, dp=0.025, dtype=np.float64)
    s = orc.OracleSim(c2, f64=True)
    s.exec("prepare_legacy")
    s.exec("run_legacy", 20.0, 1e9, 0.1, 200)
    t, e = s.series()
    out["2d_dambreak_legacy_f64"] = {"time": t.tolist(), "energy": e.tolist(),
                                     "args": {"dim": 2, "dp": 0.025, "f64": True, "end_time": 20.0,
                                              "output_interval": 0.1, "observe_every": 200}}
    print("2d legacy done", time.time() - t0, flush=True)
    c3 = cases.dam_break(dim=3, dp=0.05, dtype=np.float32)
    s = orc.OracleSim(c3, f64=False, correction=1)
    s.exec("prepare_ck")
    s.exec("run_ck", 20.0, 1e9, 1.0, 100)
    t, e = s.series()
    out["3d_dambreak_ck_f32_correction"] = {"time": t.tolist(), "energy": e.tolist(),
                                            "args": {"dim": 3, "dp": 0.05, "f64": False, "correction": 1,
                                                     "end_time": 20.0, "record_interval": 1.0, "sort_interval": 100}}
    print("3d ck done", time.time() - t0, flush=True)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if os.path.isdir(REF):
        json.dump(reference_goldens(), open(os.path.join(OUT, "reference_regression.json"), "w"), indent=1)
    json.dump(oracle_series(), open(os.path.join(OUT, "oracle_energy_series.json"), "w"), indent=1)
