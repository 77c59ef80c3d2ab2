"""CPU tests of the oracle's periodic-boundary path (BASELINE config 4: periodic Taylor-Green vortex).

The reference has periodic conditions on its TBB path only and commits no golden data for a 3-D periodic case, so the
oracle's restatement of PeriodicBounding / PeriodicCellLinkedList
(/root/reference/src/shared/particle_dynamics/general_dynamics/domian_bouding/domain_bounding.{h,cpp}) is pinned here by
properties with known answers: lattice neighbour counts, an O(n^2) minimum-image search, translation invariance and the
wrap rule itself.
"""
import numpy as np
import pytest

from helpers import make_oracle


def _oracle(case, f64=False):
    from oracle import oracle as orc
    return orc.OracleSim(case, f64=f64, free_surface=0)


def _rows(sim):
    off = sim.uint("inner_offset").astype(np.int64)
    idx = sim.uint("inner_index")
    return [np.sort(idx[off[i]:off[i + 1]]) for i in range(sim.n_fluid)]


@pytest.mark.parametrize("dim,n_side,expected", [(2, 20, 20), (3, 14, 80)])
def test_periodic_lattice_has_bulk_neighbour_count_everywhere(oracle_lib, dim, n_side, expected):
    """In a periodic box every lattice particle is a bulk particle: 80 neighbours in 3-D (integer points with
    0 < i^2+j^2+k^2 < 6.76), 20 in 2-D, and one common value of the kernel summation."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=dim, n_side=n_side, jitter=0.0)
    sim = _oracle(case)
    sim.exec("cell_list_fluid")
    sim.exec("relations")
    counts = np.diff(sim.uint("inner_offset").astype(np.int64))
    assert counts.min() == expected and counts.max() == expected
    sim.exec("compression_summation")
    s = sim.real("CompressionSummation")
    assert float(s.max() - s.min()) < 2e-5 * float(s.mean())
    # ghost entries exist only for particles within the cut-off radius of a face
    assert sim.uint("fluid_ext_index").size > case.n_fluid


def test_periodic_search_matches_minimum_image_brute_force(oracle_lib):
    """Neighbour sets of the cell-list search with ghost entries == O(n^2) minimum-image search (double precision,
    jittered lattice: no pair sits within rounding distance of the cut-off)."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=3, n_side=12, jitter=0.2, dtype=np.float64)
    sim = _oracle(case, f64=True)
    sim.exec("cell_list_fluid")
    sim.exec("relations")
    rows = _rows(sim)
    x = case.fluid_pos.astype(np.float64)
    L = 1.0
    rc2 = case.kernel.cutoff ** 2
    n = x.shape[0]
    for i in range(0, n, 7):
        d = x[i] - x
        d -= L * np.round(d / L)
        r2 = np.sum(d * d, axis=1)
        nb = np.nonzero((r2 < rc2) & (np.arange(n) != i))[0]
        assert np.array_equal(rows[i], nb.astype(np.uint32)), i


def test_periodic_translation_invariance(oracle_lib):
    """Shifting every particle by one vector and wrapping it back (PeriodicBounding) permutes nothing: neighbour sets
    are identical and the kernel summation agrees to rounding."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=3, n_side=12, jitter=0.2, dtype=np.float64)
    a = _oracle(case, f64=True)
    a.exec("cell_list_fluid"); a.exec("relations"); a.exec("compression_summation")
    b = _oracle(case, f64=True)
    b.real("Position", 3)[:] = (case.fluid_pos + np.array([0.37, -0.21, 0.55])).reshape(-1)
    b.exec("periodic_bounding")
    pos = b.real("Position", 3)
    assert pos.min() >= 0.0 and pos.max() <= 1.0
    b.exec("cell_list_fluid"); b.exec("relations"); b.exec("compression_summation")
    ra, rb = _rows(a), _rows(b)
    assert all(np.array_equal(p, q) for p, q in zip(ra, rb))
    assert np.max(np.abs(a.real("CompressionSummation") - b.real("CompressionSummation"))) < 1e-12


def test_periodic_bounding_rule(oracle_lib):
    """x < lower -> x + L ; x > upper -> x - L ; inside (bounds included) untouched; domain_bounding.h:98-108."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=2, n_side=16, jitter=0.0)
    sim = _oracle(case)
    pos = sim.real("Position", 3)
    pos[0:3] = (-0.25, 1.5, 0.0)
    pos[3:6] = (0.0, 1.0, 0.0)
    pos[6:9] = (1.0000001, -1e-7, 0.0)
    sim.exec("periodic_bounding")
    pos = sim.real("Position", 3)
    assert np.allclose(pos[0:3], (0.75, 0.5, 0.0))
    assert np.array_equal(pos[3:6], np.array([0.0, 1.0, 0.0], dtype=np.float32))
    assert abs(pos[6] - 1e-7) < 1e-7 and abs(pos[7] - 1.0) < 1e-6


def test_periodic_taylor_green_run_f32_vs_f64(oracle_lib):
    """A dozen advection steps with two sorts: the fp32 oracle stays within fp32 drift of the fp64 oracle, positions stay
    inside the box, mass is conserved and the kinetic energy decays (Riemann dissipation) without blowing up."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=3, n_side=16, jitter=0.05)
    out = {}
    for f64 in (False, True):
        sim = _oracle(case, f64=f64)
        sim.exec("prepare_ck")
        e0 = sim.exec("energy")
        sim.exec("run_ck", 1e9, 12, 1e9, 5)
        out[f64] = (sim.real("Position", 3).copy(), sim.real("Velocity", 3).copy(), sim.real("Density").copy(),
                    sim.exec("energy"), e0, int(sim.exec("acoustic_steps")))
    p32, v32, r32, e32, e0, n32 = out[False]
    p64, v64, r64, e64, _, n64 = out[True]
    assert n32 == n64
    assert p32.min() >= 0.0 and p32.max() <= 1.0
    assert np.max(np.abs(v32 - v64)) < 5e-4 * np.max(np.abs(v64))
    assert np.max(np.abs(r32 - r64)) < 1e-5
    assert 0.5 * e0 < e64 < e0 and abs(e32 - e64) < 1e-5 * e0
