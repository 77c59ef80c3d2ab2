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


PROBES_3D = [[5.366, y, 0.25] for y in (0.01, 0.1, 0.2, 0.24, 0.252, 0.266)]  # createObservationPoints, dambreak.cpp:54-65


def read_probe_series(path):
    """{probe: [values]} from an ObservedQuantityRecording regression file (Particle_k snapshot_i="...")."""
    s = open(path).read()
    out = []
    for k in range(64):
        m = re.search(r"<Particle_%d ([^>]*)/>" % k, s)
        if not m:
            break
        vals = re.findall(r'snapshot_(\d+)="([^"]+)"', m.group(1))
        out.append([float("%.6g" % float(v)) for _, v in sorted(vals, key=lambda x: int(x[0]))])
    return out


PROBES_2D = [[5.366, 0.2, 0.0]]  # Dambreak.cpp:27-28


def reference_pressure_goldens():
    out = {}
    specs = {"3d_dambreak_ck_sycl": ("tests_sycl/3d_examples/test_3d_dambreak_sycl/regression_test_tool", PROBES_3D, (0, 10, 20)),
             "3d_dambreak_legacy": ("3d_examples/test_3d_dambreak/regression_test_tool", PROBES_3D, (0, 9, 18)),
             "2d_dambreak_legacy": ("2d_examples/test_2d_dambreak/regression_test_tool", PROBES_2D, (0, 10, 20))}
    for name, (d, probes, runs) in specs.items():
        base = os.path.join(REF, d)
        thr = re.findall(r'Pressure_(\d+)="([^"]+)"', open(os.path.join(base, "FluidObserver_Pressure_dtwdistance.xml")).read())
        out[name] = {"source": d, "probes": probes, "dtw_threshold": [float(v) for _, v in sorted(thr, key=lambda x: int(x[0]))],
                     "runs": {str(r): read_probe_series(os.path.join(base, f"FluidObserver_Pressure_Run_{r}_result.xml")) for r in runs}}
    return out


def oracle_legacy_probe_series():
    """The two first-generation case files with their probes: test_2d_dambreak (probe sampled with the energy, iteration 0 and
    every 200th) and test_3d_dambreak (six probes every iteration)."""
    out = {}
    t0 = time.time()
    c2 = cases.dam_break(dim=2, dp=0.025, dtype=np.float64)
    s = orc.OracleSim(c2, f64=True, observers=PROBES_2D)
    s.exec("prepare_legacy")
    s.exec("run_legacy", 20.0, 1e9, 0.1, 200)
    p = s.probe_series()
    out["2d_dambreak_legacy_f64"] = {"pressure": [[float("%.6g" % v) for v in p[:, 0]]],
                                     "args": {"dim": 2, "dp": 0.025, "f64": True, "observers": PROBES_2D, "end_time": 20.0,
                                              "output_interval": 0.1, "observe_every": 200}}
    print("2d legacy probes done", time.time() - t0, p.shape, flush=True)
    c3 = cases.dam_break(dim=3, dp=0.05, dtype=np.float64)
    s = orc.OracleSim(c3, f64=True, observers=PROBES_3D)
    s.exec("prepare_legacy")
    s.exec("run_legacy", 20.0, 1e9, 1.0, -1)
    p = s.probe_series()
    t, e = s.series()
    out["3d_dambreak_legacy_f64"] = {"pressure": [[float("%.6g" % v) for v in p[:, k]] for k in range(p.shape[1])], "energy": e.tolist(),
                                     "args": {"dim": 3, "dp": 0.05, "f64": True, "observers": PROBES_3D, "end_time": 20.0,
                                              "output_interval": 1.0, "observe_every": -1, "sort_interval": 100}}
    print("3d legacy probes done", time.time() - t0, p.shape, flush=True)
    return out


def oracle_probe_series():
    """The complete reference case file on the oracle (LinearCorrection variants + FreeSurfaceIndication + observers)."""
    t0 = time.time()
    c3 = cases.dam_break(dim=3, dp=0.05, dtype=np.float32)
    s = orc.OracleSim(c3, f64=False, correction=1, surface_indicator=1, observers=PROBES_3D)
    s.exec("prepare_ck")
    s.exec("run_ck", 20.0, 1e9, 1.0, 100)
    t, e = s.series()
    p = s.probe_series()
    print("3d ck full case done", time.time() - t0, p.shape, flush=True)
    return {"3d_dambreak_ck_f32_full_case": {
        "time": t.tolist(), "energy": e.tolist(), "pressure": [[float("%.6g" % v) for v in p[:, k]] for k in range(p.shape[1])],
        "surface_particles_end": int(s.uint("Indicator").sum()),
        "args": {"dim": 3, "dp": 0.05, "f64": False, "correction": 1, "surface_indicator": 1, "observers": PROBES_3D,
                 "end_time": 20.0, "record_interval": 1.0, "sort_interval": 100}}}


def oracle_series_3d_legacy():
    """test_3d_dambreak (first-generation API, Real = double in the reference build): energy at iteration 0 and at every output
    interval (dambreak.cpp:100-146 of tests/3d_examples/test_3d_dambreak)."""
    t0 = time.time()
    c3 = cases.dam_break(dim=3, dp=0.05, dtype=np.float64)
    s = orc.OracleSim(c3, f64=True)
    s.exec("prepare_legacy")
    s.exec("run_legacy", 20.0, 1e9, 1.0, -1)
    t, e = s.series()
    print("3d legacy done", time.time() - t0, len(e), flush=True)
    return {"3d_dambreak_legacy_f64": {"time": t.tolist(), "energy": e.tolist(),
                                       "args": {"dim": 3, "dp": 0.05, "f64": True, "end_time": 20.0, "output_interval": 1.0,
                                                "observe_every": -1, "sort_interval": 100}}}


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
    if "--pressure" in sys.argv:  # only the probe fixtures (the energy fixtures are left as they are)
        if os.path.isdir(REF):
            json.dump(reference_pressure_goldens(), open(os.path.join(OUT, "reference_pressure_probes.json"), "w"))
        series = oracle_probe_series()
        series.update(oracle_legacy_probe_series())
        json.dump(series, open(os.path.join(OUT, "oracle_probe_series.json"), "w"))
        sys.exit(0)
    if "--legacy-pressure" in sys.argv:  # only the probe series of the two first-generation case files, merged into the fixtures
        if os.path.isdir(REF):
            json.dump(reference_pressure_goldens(), open(os.path.join(OUT, "reference_pressure_probes.json"), "w"))
        path = os.path.join(OUT, "oracle_probe_series.json")
        cur = json.load(open(path))
        cur.update(oracle_legacy_probe_series())
        json.dump(cur, open(path, "w"))
        sys.exit(0)
    if "--legacy-3d" in sys.argv:  # only the 3-D legacy energy series, merged into the existing fixture
        path = os.path.join(OUT, "oracle_energy_series.json")
        cur = json.load(open(path))
        cur.update(oracle_series_3d_legacy())
        json.dump(cur, open(path, "w"), indent=1)
        sys.exit(0)
    if os.path.isdir(REF):
        json.dump(reference_goldens(), open(os.path.join(OUT, "reference_regression.json"), "w"), indent=1)
    series = oracle_series()
    series.update(oracle_series_3d_legacy())
    json.dump(series, open(os.path.join(OUT, "oracle_energy_series.json"), "w"), indent=1)
