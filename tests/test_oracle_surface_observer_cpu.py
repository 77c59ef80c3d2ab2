"""CPU tests of the oracle's FreeSurfaceIndicationCK and observer-probe restatements,
pinned by properties the reference's formulas imply and by the reference's committed pressure-probe regression series.

Golden sources (relative to /root/reference):
  tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/regression_test_tool/FluidObserver_Pressure_Run_{0,10,20}_result.xml
  (2,281 / 2,285 / 2,296 records of six probes each) and FluidObserver_Pressure_dtwdistance.xml (threshold 2.5 per probe), copied as
  numbers into tests/golden/reference_pressure_probes.json by oracle/make_golden.py --pressure.
"""
import json
import os

import numpy as np
import pytest

from helpers import dtw_distance, dtw_two_sided

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
PROBES = [[5.366, y, 0.25] for y in (0.01, 0.1, 0.2, 0.24, 0.252, 0.266)]


def _sim(dim=3, dp=0.05, f64=True, **kw):
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    case = cases.dam_break(dim=dim, dp=dp, dtype=np.float64 if f64 else np.float32)
    s = orc.OracleSim(case, f64=f64, surface_indicator=1, **kw)
    s.exec("prepare_ck")
    return case, s


@pytest.mark.parametrize("dim,dp", [(3, 0.05), (2, 0.025)])
def test_position_divergence_and_indicator_on_the_lattice(dim, dp):
    """sum_j -dW_ij V_j r_ij reproduces the dimension for a full support (the quantity the threshold 0.75 * Dimensions is
    measured against, surface_indication_ck.hpp:19); particles of the outermost layers of the two free faces are flagged,
    particles more than the support radius below them are not."""
    case, s = _sim(dim, dp)
    s.exec("surface_indication")
    pd = s.real("PositionDivergence").copy()
    ind = s.uint("Indicator").copy()
    pos = s.real("Position", 3).reshape(-1, 3)
    rc = 2.6 * dp
    deep = (pos[:, 0] < case.LL - rc) & (pos[:, 1] < case.LH - rc)  # full support (walls count as neighbours)
    assert deep.sum() > 100
    assert np.all(np.abs(pd[deep] - dim) < 0.03 * dim), (pd[deep].min(), pd[deep].max())
    assert not ind[deep].any()
    away = (pos[:, 0] > rc) & (pos[:, 1] > rc)  # wall particles fill the support next to the tank walls
    if dim == 3:
        away &= (pos[:, 2] > rc) & (pos[:, 2] < case.LW - rc)
    outer = ((pos[:, 0] > case.LL - 0.9 * dp) | (pos[:, 1] > case.LH - 0.9 * dp)) & away
    assert outer.sum() > 20 and ind[outer].all()
    assert np.array_equal(s.uint("PreviousSurfaceIndicator"), ind)
    assert set(np.unique(ind).tolist()) <= {0, 1}


def test_spatial_temporal_override():
    """A particle whose divergence drops below the threshold is only accepted as a surface particle if it was one before
    or has a previous surface particle among its neighbours (surface_indication_ck.hpp:64-69): with an all-zero
    previous indicator nothing can be flagged, and the override value 2 * threshold is what gets stored."""
    case, s = _sim()
    s.exec("surface_indication")
    first = s.uint("Indicator").copy()
    pd_first = s.real("PositionDivergence").copy()
    assert first.sum() > 0
    s.uint("PreviousSurfaceIndicator")[:] = 0
    s.exec("surface_indication")
    assert s.uint("Indicator").sum() == 0
    pd = s.real("PositionDivergence")
    was_surface_like = pd_first < 2.25
    # inner divergence < threshold -> replaced by 4.5, then the wall part is added on top
    assert np.all(pd[was_surface_like] >= 4.5 - 1e-12)
    # idempotence with the proper history: same state + own previous indicator -> same answer
    s.uint("PreviousSurfaceIndicator")[:] = first
    s.exec("surface_indication")
    assert np.array_equal(s.uint("Indicator"), first)
    # (PositionDivergence itself is NOT idempotent: near a wall the inner part alone is below the threshold, so a
    # particle that is not a previous surface particle gets the override before the wall part is added)
    assert np.array_equal(s.real("PositionDivergence")[first == 1], pd_first[first == 1])


def test_sort_carries_previous_indicator_only():
    """PreviousSurfaceIndicator is an evolving (sorted) variable, Indicator is not (surface_indication_ck.hpp:42-46)."""
    case, s = _sim(f64=False)
    s.exec("run_ck", 1e9, 3, 1e9, 100)
    prev = s.uint("PreviousSurfaceIndicator").copy()
    ind = s.uint("Indicator").copy()
    s.exec("sort")
    perm = s.uint("Permutation").copy()
    assert np.array_equal(s.uint("PreviousSurfaceIndicator"), prev[perm])
    assert np.array_equal(s.uint("Indicator"), ind)


def test_observer_interpolation_properties():
    """Shepard-normalised interpolation (interpolation_dynamics.hpp:44-60): a constant field is reproduced exactly up to
    rounding, a probe without neighbours reads 0 / TinyReal = 0, and the relation is the plain cut-off set."""
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    case = cases.dam_break(dim=3, dp=0.05, dtype=np.float64)
    probes = np.array([[1.0, 0.5, 0.25], [1.97, 0.97, 0.25], [3.0, 0.5, 0.25], [0.3, 0.2, 0.1]])
    s = orc.OracleSim(case, f64=True, observers=probes)
    s.exec("prepare_ck")
    s.real("Pressure")[:] = 7.25
    s.exec("observe_pressure")
    out = s.real("Pressure", 1, body=2)
    assert np.allclose(out[[0, 1, 3]], 7.25, rtol=1e-12)
    assert out[2] == 0.0
    off, idx = s.uint("observer_offset"), s.uint("observer_index")
    pos = s.real("Position", 3).reshape(-1, 3)
    for i, x in enumerate(probes):
        want = np.nonzero(np.sum((pos - x) ** 2, axis=1) < (2.6 * 0.05) ** 2 * (1 - 1e-12))[0]
        assert sorted(idx[off[i]:off[i + 1]].tolist()) == want.tolist()
    # linear field: the Shepard interpolant of a linear function at the centre of a symmetric lattice stencil is exact
    pos_y = pos[:, 1].copy()
    s.real("Pressure")[:] = 3.0 * pos_y
    s.exec("observe_pressure")
    centre = np.array([1.0 + 0.0, 0.5, 0.25])  # probe 0 sits on a lattice plane midpoint in x, y, z (dp = 0.05)
    assert abs(s.real("Pressure", 1, body=2)[0] - 3.0 * centre[1]) < 1e-9


def _gold(name):
    return json.load(open(os.path.join(GOLD, name)))


def test_reference_probe_fixture_shape():
    g = _gold("reference_pressure_probes.json")["3d_dambreak_ck_sycl"]
    assert g["dtw_threshold"] == [2.5] * 6
    assert sorted(g["runs"]) == ["0", "10", "20"]
    for r, n in (("0", 2281), ("10", 2285), ("20", 2296)):  # records = advection steps + 1 of that reference run
        assert len(g["runs"][r]) == 6 and all(len(series) == n for series in g["runs"][r])
    assert g["probes"] == PROBES


def test_oracle_full_case_probe_series_meets_reference_dtw():
    """The oracle's run of the COMPLETE reference case file (fixture produced by oracle/make_golden.py --pressure: fp32,
    LinearCorrection variants + FreeSurfaceIndication + six probes, t = 0..20) against the reference's committed probe
    series under the reference's own acceptance test (DTW distance <= 2.5 per probe), and the energy series (<= 0.05)."""
    ref = _gold("reference_pressure_probes.json")["3d_dambreak_ck_sycl"]
    eref = _gold("reference_regression.json")["3d_dambreak_ck_sycl"]
    mine = _gold("oracle_probe_series.json")["3d_dambreak_ck_f32_full_case"]
    assert len(mine["pressure"]) == 6
    n = len(mine["pressure"][0])
    assert 0.8 * 2281 <= n <= 1.2 * 2281
    for k in range(6):
        d = [dtw_distance(run[k], mine["pressure"][k]) for run in ref["runs"].values()]
        assert min(d) <= ref["dtw_threshold"][k], (k, d)
    de = [dtw_distance(run, mine["energy"]) for run in eref["runs"].values()]
    assert max(de) <= eref["dtw_threshold"], de


def test_oracle_probe_fixture_is_reproducible_prefix():
    """The fixture comes from the current oracle: the first 40 advection steps reproduce its first 41 records."""
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    mine = _gold("oracle_probe_series.json")["3d_dambreak_ck_f32_full_case"]
    case = cases.dam_break(dim=3, dp=0.05, dtype=np.float32)
    s = orc.OracleSim(case, f64=False, correction=1, surface_indicator=1, observers=PROBES)
    s.exec("prepare_ck")
    s.exec("run_ck", 1e9, 40, 1e9, 100)
    p = s.probe_series()
    assert p.shape == (41, 6)
    want = np.array(mine["pressure"])[:, :41].T
    assert np.allclose(p, want, rtol=2e-5, atol=1e-6)


def test_oracle_legacy_probe_series_pass_reference_dtw_thresholds():
    """The wall-pressure probes of the two first-generation case files the reference holds regression data for — test_2d_dambreak
    (one probe sampled with the energy: 23 snapshots, threshold 1.078) and test_3d_dambreak (six probes written every iteration:
    2,180 records, thresholds 1.5 - 4.5) — from the oracle's full runs (oracle/make_golden.py --legacy-pressure), under the
    reference's criterion in both argument orders."""
    ref = _gold("reference_pressure_probes.json")
    ours = _gold("oracle_probe_series.json")
    g2, o2 = ref["2d_dambreak_legacy"], ours["2d_dambreak_legacy_f64"]["pressure"][0]
    assert len(o2) == 23
    for run, series in g2["runs"].items():
        d = dtw_two_sided(o2, series[0])
        assert d <= g2["dtw_threshold"][0], f"2-D probe, run {run}: DTW {d:.3f} > {g2['dtw_threshold'][0]}"
    g3, o3 = ref["3d_dambreak_legacy"], ours["3d_dambreak_legacy_f64"]["pressure"]
    assert len(o3) == 6 and 0.95 * 2180 < len(o3[0]) < 1.05 * 2180
    for k in range(6):
        for run, series in g3["runs"].items():
            d = dtw_two_sided(o3[k], series[k])
            assert d <= g3["dtw_threshold"][k], f"3-D legacy probe {k}, run {run}: DTW {d:.3f} > {g3['dtw_threshold'][k]}"


def test_oracle_reproduces_committed_2d_probe_prefix():
    """The committed 2-D probe fixture comes from the current oracle: iteration 0 and every 200th iteration up to t = 4.6."""
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    ours = _gold("oracle_probe_series.json")["2d_dambreak_legacy_f64"]
    case = cases.dam_break(dim=2, dp=0.025, dtype=np.float64)
    o = orc.OracleSim(case, f64=True, observers=ours["args"]["observers"])
    o.exec("prepare_legacy")
    o.exec("run_legacy", 4.6, 1e9, 0.1, 200)
    p = o.probe_series()
    n = p.shape[0]
    assert n >= 5 and np.allclose(p[:n, 0], ours["pressure"][0][:n], rtol=1e-5, atol=1e-12)
    assert np.max(np.abs(p[:n, 0])) > 0.1  # the water has reached the probe within the prefix
