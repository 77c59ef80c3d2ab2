"""N > 1 on the CPU: the exchange protocol of the slab-decomposed run against the single-domain oracle.

oracle/decomposed.py restates what SlabDecomposition / DamBreakCK::stepOuter do on N GPUs — plane ownership, migration
and ghost planes at the configuration update, the three ghost refreshes inside a step, own-particle reductions combined
by max — with the oracle doing the arithmetic. The bar is the one the GPU runs are held to (tests/multi_gpu_check.py):
every variable of every particle bit-identical to the undecomposed run. The ranks run as threads of this process for
the wider sweeps and as two gloo processes (one process per rank, as on the GPUs) for the world_size-2 test.
The periodic ring along x (BASELINE config 4 on N GPUs) is pinned here BEFORE its CUDA side exists: DESIGN.md §6c.
"""
import dataclasses
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import decomposed as dec  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from sphinxsys_b200 import cases  # noqa: E402


def _single(case, steps, **kw):
    g = orc.OracleSim(case, **kw)
    g.exec("prepare_ck")
    g.exec("run_ck", 1e9, steps, 1e9, 0)  # no ParticleSortCK: decomposed runs keep the initial numbering
    return g


def _mismatches(g, glob):
    bad = []
    for nm in glob:
        if nm in dec.INDICATOR_UINTS:
            if not np.array_equal(g.uint(nm), glob[nm].reshape(-1)):
                bad.append(nm)
            continue
        w = dec.WIDTH[nm]
        ref = g.real(nm, w).reshape(-1, w)
        if not np.array_equal(ref.view(np.uint32), glob[nm].view(np.uint32)):
            bad.append(nm)
    return bad


def _dam_break():
    case = cases.dam_break(dim=3, dp=0.05)
    planes = dec.x_plane(case.fluid_pos, case.mesh)
    return case, planes


def _ring_case(drift, x_scale=1, n_side=16, dim=3):
    case = cases.taylor_green(dim=dim, n_side=n_side, x_scale=x_scale)
    mesh, first, planes = dec.aligned_periodic_mesh(case)
    vel = case.fluid_vel.copy()
    vel[:, 0] += np.float32(drift)  # uniform drift: particles cross the seam between the last and the first rank
    return dataclasses.replace(case, mesh=mesh, fluid_vel=vel), first, planes


@pytest.mark.parametrize("nranks,steps", [(2, 12), (4, 60)])
def test_dam_break_slabs_bit_identical(nranks, steps):
    case, planes = _dam_break()
    cuts = dec.plan_cuts(planes, 0, case.mesh.cells[0], nranks)
    g = _single(case, steps)
    states, ranks = dec.run_threads(case, nranks, cuts, steps)
    assert all(r.acoustic_steps == int(g.exec("acoustic_steps")) for r in ranks), "ranks disagree on the time steps"
    assert _mismatches(g, dec.gather_by_gid(states, case.n_fluid)) == []
    if steps >= 60:
        assert sum(r.migrated for r in ranks) > 0, "the run should exercise migration"


def test_dam_break_cuts_do_not_matter():
    """Results do not depend on where the slabs are cut (in-cell order is by global id), slabs of ONE plane included
    (its particles are the left and the right boundary plane at once)."""
    case, _ = _dam_break()
    g = _single(case, 8)
    for cuts in ([0, 7, 52], [0, 16, 52], [0, 5, 9, 52], [0, 10, 11, 12, 13, 52]):
        states, _ = dec.run_threads(case, len(cuts) - 1, cuts, 8)
        assert _mismatches(g, dec.gather_by_gid(states, case.n_fluid)) == []


def test_dam_break_recut_keeps_bit_identity():
    """Re-balancing at the sort cadence (SlabDecomposition::recut) with the host layer's own planner: deliberately skewed
    initial cuts are pulled back over a few re-cuts, every hand-over goes to a neighbour rank (asserted inside rebuild),
    and the fields stay those of the single-domain run."""
    import threading
    from sphinxsys_b200 import host
    case, planes = _dam_break()
    nranks, steps = 3, 24
    balanced = dec.plan_cuts(planes, 0, case.mesh.cells[0], nranks)
    skewed = [balanced[0], balanced[1] + 3, balanced[2] + 6, balanced[3]]
    g = _single(case, steps)
    comms = dec.ThreadComm.make(nranks)
    ranks, errors, changes = [None] * nranks, [], [0] * nranks

    def work(r):
        try:
            sr = dec.SlabRank(case, comms[r], skewed)
            for k in range(1, steps + 1):
                sr.step_outer()
                if k % 4 == 0:
                    changes[r] += int(sr.recut(host.plan_slab_cuts, host.limit_cut_moves))
            ranks[r] = sr
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            comms[r]._s.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
    assert not errors, (real or errors)[0]
    assert changes[0] > 0 and all(r.cuts == ranks[0].cuts for r in ranks)
    own = [r.n_own for r in ranks]
    assert max(own) - min(own) < 0.35 * case.n_fluid / nranks, f"slabs not re-balanced: {own}"
    assert _mismatches(g, dec.gather_by_gid([r.own_state() for r in ranks], case.n_fluid)) == []


def test_ring_recut_keeps_bit_identity():
    """Re-cutting on a ring: the cuts inside the periodic box move (the seam stays), hand-overs go to neighbours only."""
    import threading
    from sphinxsys_b200 import host
    case, first, planes = _ring_case(1.5, x_scale=2, n_side=12)   # 9 box planes
    skewed = [first, first + 1, first + 2, first + planes]
    steps, nranks = 12, 3
    g = _single(case, steps, free_surface=0)
    comms = dec.ThreadComm.make(nranks)
    ranks, errors, changes = [None] * nranks, [], [0] * nranks

    def work(r):
        try:
            sr = dec.SlabRank(case, comms[r], skewed, ring=True, free_surface=0)
            for k in range(1, steps + 1):
                sr.step_outer()
                if k % 3 == 0:
                    changes[r] += int(sr.recut(host.plan_slab_cuts, host.limit_cut_moves))
            ranks[r] = sr
        except BaseException as e:  # noqa: BLE001
            errors.append(e)
            comms[r]._s.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
    assert not errors, (real or errors)[0]
    assert changes[0] > 0 and ranks[0].cuts[0] == first and ranks[0].cuts[-1] == first + planes
    own = [r.n_own for r in ranks]
    assert max(own) - min(own) < 0.2 * case.n_fluid, f"slabs not re-balanced: {own} with cuts {ranks[0].cuts}"
    assert _mismatches(g, dec.gather_by_gid([r.own_state() for r in ranks], case.n_fluid)) == []


@pytest.mark.parametrize("viscosity", [0.0, 0.02])
def test_dam_break_correction_variants_bit_identical(viscosity):
    """The Correction aliases the complete reference case file uses (dambreak.cpp:117-124): one more refresh, of the B
    matrix right after LinearCorrectionMatrix has rebuilt it, keeps the decomposed run bit-identical; without it, it is not."""
    case, planes = _dam_break()
    cuts = dec.plan_cuts(planes, 0, case.mesh.cells[0], 3)
    kw = dict(correction=1, viscosity=viscosity)
    g = _single(case, 20, **kw)
    states, _ = dec.run_threads(case, 3, cuts, 20, **kw)
    glob = dec.gather_by_gid(states, case.n_fluid)
    assert "LinearCorrectionMatrix" in glob and _mismatches(g, glob) == []
    if viscosity == 0.0:
        states, _ = dec.run_threads(case, 3, cuts, 6, skip_refresh=["LinearCorrectionMatrix"], **kw)
        assert _mismatches(_single(case, 6, **kw), dec.gather_by_gid(states, case.n_fluid)) != []


def test_dam_break_complete_case_dynamics_bit_identical():
    """The dynamics of the complete reference case file (dambreak.cpp:117-134,188-205: Correction aliases + free-surface
    indication) on three slabs: PositionDivergence has to be refreshed on the ghost planes between the two sweeps of the
    indication (the second reads the first's result on the neighbours); without that the indicator differs."""
    case, planes = _dam_break()
    cuts = dec.plan_cuts(planes, 0, case.mesh.cells[0], 3)
    kw = dict(correction=1, surface_indicator=1)
    g = _single(case, 30, **kw)
    states, ranks = dec.run_threads(case, 3, cuts, 30, **kw)
    glob = dec.gather_by_gid(states, case.n_fluid)
    assert "Indicator" in glob and 0 < int(glob["Indicator"].sum()) < case.n_fluid
    assert _mismatches(g, glob) == []
    assert sum(r.migrated for r in ranks) > 0
    states, _ = dec.run_threads(case, 3, cuts, 30, skip_refresh=["PositionDivergence"], **kw)
    assert "Indicator" in _mismatches(g, dec.gather_by_gid(states, case.n_fluid))


def test_dam_break_observer_probes_bit_identical():
    """The six wall-pressure probes of the reference case file (dambreak.cpp:54-65,223-224) on three slabs: every probe is
    taken from the rank that owns its cell plane (it stores the probe's whole neighbourhood); the recorded series equals
    the single-domain one bit for bit, also for probes placed right at the cuts."""
    case, planes = _dam_break()
    cuts = dec.plan_cuts(planes, 0, case.mesh.cells[0], 3)
    s = case.mesh.spacing
    at_cuts = [[case.mesh.lower[0] + c * s + dx, 0.3, 0.25] for c in cuts[1:-1] for dx in (-1e-4, 1e-4)]
    probes = [[case.DL, y, 0.5 * case.DW] for y in (0.01, 0.1, 0.2, 0.24, 0.252, 0.266)] + at_cuts + [[0.3, 0.2, 0.1], [1.7, 0.6, 0.4]]
    kw = dict(observers=probes)
    g = _single(case, 20, **kw)
    ref = g.probe_series()
    states, ranks = dec.run_threads(case, 3, cuts, 20, **kw)
    got = np.array(ranks[0].probe_series)
    assert got.shape == ref.shape == (21, len(probes))
    assert np.array_equal(got.astype(np.float32).view(np.uint32), ref.astype(np.float32).view(np.uint32))
    assert np.abs(ref[:, 6:]).max() > 0, "the probes inside the water column should see pressure"
    assert all(np.array_equal(np.array(r.probe_series), got) for r in ranks)


@pytest.mark.parametrize("stale", ["VolumetricMeasure", "Pressure", "Velocity"])
def test_every_refresh_is_needed(stale):
    """Each of the three ghost refreshes carries a value the neighbours' own particles read: dropping one breaks parity."""
    case, planes = _dam_break()
    cuts = dec.plan_cuts(planes, 0, case.mesh.cells[0], 2)
    g = _single(case, 6)
    states, _ = dec.run_threads(case, 2, cuts, 6, skip_refresh=[stale])
    assert _mismatches(g, dec.gather_by_gid(states, case.n_fluid)) != []


@pytest.mark.parametrize("nranks,steps,drift", [(2, 10, 1.0), (3, 30, 1.0), (2, 15, -1.5), (6, 10, 1.5)])
def test_periodic_ring_bit_identical(nranks, steps, drift):
    case, first, planes = _ring_case(drift)
    cuts = dec.plan_cuts(dec.x_plane(case.fluid_pos, case.mesh), first, first + planes, nranks)
    g = _single(case, steps, free_surface=0)
    states, ranks = dec.run_threads(case, nranks, cuts, steps, ring=True, free_surface=0)
    assert _mismatches(g, dec.gather_by_gid(states, case.n_fluid)) == []
    seam = ranks[-1] if drift > 0 else ranks[0]
    assert seam.wrapped > 0, "particles should have crossed the periodic seam"


def test_periodic_ring_replicated_box_bit_identical():
    """Weak-scaling shape of config 4: the box replicated along x (2 L x L x L), one copy per rank."""
    case, first, planes = _ring_case(1.5, x_scale=2, n_side=12)
    assert case.n_fluid == 2 * 12 ** 3 and case.periodic_upper[0] == 2.0
    cuts = dec.plan_cuts(dec.x_plane(case.fluid_pos, case.mesh), first, first + planes, 2)
    g = _single(case, 10, free_surface=0)
    states, ranks = dec.run_threads(case, 2, cuts, 10, ring=True, free_surface=0)
    assert _mismatches(g, dec.gather_by_gid(states, case.n_fluid)) == []
    assert ranks[-1].wrapped > 0 and abs(ranks[0].n_own - ranks[1].n_own) < 0.15 * case.n_fluid  # 9 planes: 4 + 5


def test_periodic_ring_viscous_transport_bit_identical():
    """The Taylor-Green case as the reference runs it — viscous force and transport-velocity correction
    (taylor_green.cpp:106-116) — on a ring: no refresh beyond the three of the inviscid step is needed."""
    case, first, planes = _ring_case(1.0)
    kw = dict(free_surface=0, viscosity=0.01, transport_velocity=1)
    cuts = dec.plan_cuts(dec.x_plane(case.fluid_pos, case.mesh), first, first + planes, 2)
    g = _single(case, 12, **kw)
    states, ranks = dec.run_threads(case, 2, cuts, 12, ring=True, **kw)
    glob = dec.gather_by_gid(states, case.n_fluid)
    assert "PreviousViscousForce" in glob and np.abs(glob["ViscousForce"]).max() > 0
    assert _mismatches(g, glob) == []
    assert ranks[-1].wrapped > 0


def test_periodic_ring_2d_bit_identical():
    """The 2-D Taylor-Green case (taylor_green.cpp) on a ring of three slabs."""
    case, first, planes = _ring_case(1.0, n_side=40, dim=2)
    cuts = dec.plan_cuts(dec.x_plane(case.fluid_pos, case.mesh), first, first + planes, 3)
    g = _single(case, 25, free_surface=0)
    states, ranks = dec.run_threads(case, 3, cuts, 25, ring=True, free_surface=0)
    assert _mismatches(g, dec.gather_by_gid(states, case.n_fluid)) == []
    assert ranks[-1].wrapped > 0


def test_periodic_ring_of_one_slab_bit_identical():
    """One rank that is its own neighbour on both sides (what tests/test_gpu_zz_periodic_ring.py runs on the GPU): x periodic
    through the seam exchange alone, y / z through the image entries."""
    case, first, planes = _ring_case(2.0)
    g = _single(case, 12, free_surface=0)
    sr = dec.SlabRank(case, dec.SerialComm(), [first, first + planes], ring=True, free_surface=0)
    for _ in range(12):
        sr.step_outer()
    assert sr.wrapped > 0 and sr.n_ghost[0] > 0 and sr.n_ghost[1] > 0
    assert _mismatches(g, dec.gather_by_gid([sr.own_state()], case.n_fluid)) == []


def test_aligned_periodic_mesh_keeps_neighbour_sets():
    """The aligned mesh (spacing L / floor(L / r_c) >= r_c) changes cells, not neighbours: same sorted rows as the case mesh."""
    case = cases.taylor_green(dim=3, n_side=16)
    mesh, first, planes = dec.aligned_periodic_mesh(case)
    assert mesh.spacing >= case.kernel.cutoff and planes == int(1.0 / case.kernel.cutoff)
    a = orc.OracleSim(case, free_surface=0)
    b = orc.OracleSim(dataclasses.replace(case, mesh=mesh), free_surface=0)
    rows = []
    for s in (a, b):
        s.exec("prepare_ck")
        off, idx = s.uint("inner_offset"), s.uint("inner_index")
        rows.append([tuple(sorted(idx[off[i]:off[i + 1]].tolist())) for i in range(0, case.n_fluid, 7)])
    assert rows[0] == rows[1]


# ---- one process per rank over gloo, as the GPU run has one process per GPU ----
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for name in ("dam_break", "complete", "ring"):
            if name in ("dam_break", "complete"):
                case, planes = _dam_break()
                cuts, steps, kw, ring = dec.plan_cuts(planes, 0, case.mesh.cells[0], world), 10, {}, False
                if name == "complete":  # the dynamics of the complete reference case file: correction, indication, probes
                    kw = dict(correction=1, surface_indicator=1,
                              observers=[[case.DL, y, 0.5 * case.DW] for y in (0.01, 0.1, 0.2)] + [[0.9, 0.3, 0.25]])
            else:
                case, first, nplanes = _ring_case(1.0)
                cuts = dec.plan_cuts(dec.x_plane(case.fluid_pos, case.mesh), first, first + nplanes, world)
                steps, kw, ring = 8, {"free_surface": 0}, True
            sr = dec.SlabRank(case, dec.GlooComm(), cuts, ring=ring, **kw)
            for _ in range(steps):
                sr.step_outer()
            states = [None] * world
            dist.all_gather_object(states, sr.own_state())
            if rank == 0:
                g = _single(case, steps, **kw)
                bad = _mismatches(g, dec.gather_by_gid(states, case.n_fluid))
                if "observers" in kw and not np.array_equal(np.array(sr.probe_series, dtype=np.float32), g.probe_series().astype(np.float32)):
                    bad.append("probe series")
                out[name] = (bad, sr.acoustic_steps, int(g.exec("acoustic_steps")), sr.migrated + sr.wrapped)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_domain():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for name in ("dam_break", "complete", "ring"):
        bad, ac, ac_single, moved = results[0][name]
        assert bad == [], f"{name}: variables differ from the single-domain oracle: {bad}"
        assert ac == ac_single
    assert results[0]["ring"][3] > 0


def test_protocol_fuzz():
    """Random slabs (empty and one-plane slabs included), random mixes of the optional dynamics, random drifts through the
    seam, random lengths: the decomposed run is the single-domain run, bit for bit, every time."""
    rng = np.random.default_rng(20261018)
    dam, _ = _dam_break()
    for _ in range(3):
        n = int(rng.integers(2, 6))
        cuts = [0] + np.sort(rng.choice(np.arange(5, 30), size=n - 1, replace=False)).tolist() + [dam.mesh.cells[0]]
        steps = int(rng.integers(8, 25))
        kw = dict(correction=int(rng.integers(0, 2)), surface_indicator=int(rng.integers(0, 2)))
        states, _ranks = dec.run_threads(dam, n, cuts, steps, **kw)
        assert _mismatches(_single(dam, steps, **kw), dec.gather_by_gid(states, dam.n_fluid)) == [], (cuts, steps, kw)
    for _ in range(3):
        case, first, planes = _ring_case(float(rng.uniform(-2.5, 2.5)), x_scale=int(rng.integers(1, 3)), n_side=int(rng.choice([12, 16])))
        n = int(rng.integers(2, 5))
        cuts = [first] + np.sort(rng.choice(np.arange(first + 1, first + planes), size=n - 1, replace=False)).tolist() + [first + planes]
        steps = int(rng.integers(8, 20))
        kw = dict(free_surface=0, viscosity=float(rng.choice([0.0, 0.01])), transport_velocity=int(rng.integers(0, 2)))
        states, _ranks = dec.run_threads(case, n, cuts, steps, ring=True, **kw)
        assert _mismatches(_single(case, steps, **kw), dec.gather_by_gid(states, case.n_fluid)) == [], (cuts, steps, kw)
