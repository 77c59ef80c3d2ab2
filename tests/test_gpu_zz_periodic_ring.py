"""GPU test of the periodic RING of slabs (BASELINE config 4 on N GPUs) on ONE GPU: a ring of one slab.

The rank is its own left and right neighbour (sphb200_comm_create_self + sphb200_comm_set_ring), so everything the
N-GPU path does runs here except the NCCL transport itself: the aligned mesh, plane ownership, the selection of the
boundary planes and the leavers, the seam shift of what crosses the box face (sphb200_seam_shift), the ghost planes in
the cell order, their three refreshes inside a step, the y / z images made for own particles and ghost planes alike.
Periodicity along x comes from the slab exchange alone — no x images — and is checked against the oracle's
single-domain periodic run (ghost list entries on all three axes) on the same mesh. The protocol itself is pinned bit
for bit on the CPU at 1-3 ranks (tests/test_decomposed_oracle_cpu.py).

Bar: no particle lost or owned twice; fields within the tolerances of tests/test_gpu_periodic.py (order inside a
neighbour row differs from the oracle's by construction, so sums differ in the last bits).
"""
import dataclasses
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from helpers import rel_err  # noqa: E402

REPORT = {}

def _report(key, value):
    REPORT[key] = value
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        json.dump(REPORT, open(os.path.join(out, "ring_report.json"), "w"), indent=1, default=float)
    except OSError:
        pass


def _case(n_side, drift, jitter=0.05, dim=3):
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=dim, n_side=n_side, jitter=jitter)
    vel = case.fluid_vel.copy()
    vel[:, 0] += np.float32(drift)  # uniform drift: particles cross the seam
    return dataclasses.replace(case, fluid_vel=vel)


def _own(gpu, name, n):
    """Own particles of a variable brought into the global numbering; asserts that ownership is a partition."""
    rid = gpu.download_own("ReferenceID").astype(np.int64)
    assert rid.size == n and np.array_equal(np.sort(rid), np.arange(n)), "a particle was lost or is owned twice"
    a = gpu.download_own(name)
    out = np.empty_like(a)
    out[rid] = a
    return out


def _oracle_on_gpu_mesh(case, gpu, **kw):
    from oracle import oracle as orc
    from sphinxsys_b200 import hostmath as hm
    m = gpu.mesh()
    mesh = hm.MeshSpec(tuple(float(v) for v in m.lower), float(m.spacing), tuple(int(c) for c in m.cells))
    return orc.OracleSim(dataclasses.replace(case, mesh=mesh), free_surface=0, **kw), mesh


def test_ring_of_one_slab_first_configuration():
    """After initialize(): the two seam ghost planes are the opposite boundary planes, and the density summation over
    own + seam ghosts + y/z images equals the oracle's over its periodic entry table."""
    from sphinxsys_b200.host import TaylorGreenCK
    case = _case(16, 0.0, jitter=0.1)
    gpu = TaylorGreenCK(case, ring=True)
    gpu.initialize()
    ref, mesh = _oracle_on_gpu_mesh(case, gpu)
    assert mesh.spacing >= case.kernel.cutoff
    planes = int(gpu.exec("box_planes"))
    assert planes == int(1.0 / case.kernel.cutoff)
    begin, count, stored = gpu.own_range()
    assert count == case.n_fluid
    # seam ghost planes: the particles of the first and of the last own plane, seen from the other side of the box
    from oracle import decomposed as dec
    pl = dec.x_plane(case.fluid_pos, mesh)
    expect = int((pl == pl.min()).sum() + (pl == pl.max()).sum())
    assert int(gpu.exec("plane_ghost_particles")) == expect
    _report("first_configuration", {"own": count, "plane_ghosts": expect, "images": gpu.ghost_particles, "box_planes": planes})
    ref.exec("cell_list_fluid")
    ref.exec("relations")
    gpu.exec("density_summation")
    ref.exec("compression_summation")
    e = rel_err(_own(gpu, "CompressionSummation", case.n_fluid), ref.real("CompressionSummation").copy())
    _report("summation_rel_err", e)
    assert e < 1e-5, f"CompressionSummation {e:.3e}"
    pos = _own(gpu, "Position", case.n_fluid)
    assert np.array_equal(pos, case.fluid_pos), "own positions must be untouched by the first configuration"


@pytest.mark.parametrize("n_side,n_outer,drift", [(16, 15, 2.0), (20, 12, -1.5)])
def test_ring_of_one_slab_against_single_domain_oracle(n_side, n_outer, drift):
    """The case loop with a drift through the seam: same tolerances as test_taylor_green_multi_step_drift."""
    from sphinxsys_b200.host import TaylorGreenCK
    case = _case(n_side, drift)
    gpu = TaylorGreenCK(case, ring=True)
    gpu.initialize()
    o32, _ = _oracle_on_gpu_mesh(case, gpu)
    o64, _ = _oracle_on_gpu_mesh(case, gpu, f64=True)
    for o in (o32, o64):
        o.exec("prepare_ck")
        o.exec("run_ck", 1e9, n_outer, 1e9, 0)
    n_ac = gpu.run_outer(n_outer)
    assert int(o32.exec("acoustic_steps")) == n_ac
    same_path = int(o64.exec("acoustic_steps")) == n_ac
    n = case.n_fluid
    rep = {"acoustic_steps": n_ac}
    pos = _own(gpu, "Position", n)
    ref_pos = o32.real("Position", 3).reshape(-1, 3)
    # a particle exactly at a face may sit on either side of it (plane ownership vs position test): compare modulo L
    d = pos.astype(np.float64) - ref_pos
    d -= np.round(d)  # (y and z: a particle within rounding of a face may be wrapped on one side only)
    rep["Position"] = float(np.abs(d).max())
    assert rep["Position"] < 5e-6, f"Position {rep['Position']:.3e}"
    crossed = int((np.abs(pos[:, 0].astype(np.float64) - case.fluid_pos[:, 0]) > 0.5).sum())
    rep["crossed_the_seam"] = crossed
    assert crossed > 0, "the drift should carry particles through the seam"
    assert pos[:, 0].min() >= -1e-6 and pos[:, 0].max() <= 1.0 + 1e-6
    for nm, w, tol in (("Velocity", 3, 2e-4), ("Density", 1, 2e-6), ("Compression", 1, 2e-6)):
        a, b = _own(gpu, nm, n), o32.real(nm, w).reshape(-1, w) if w > 1 else o32.real(nm, w)
        e = rel_err(a, b)
        b64 = o64.real(nm, w).reshape(-1, w) if w > 1 else o64.real(nm, w)
        noise = rel_err(b, b64) if same_path else 0.0
        rep[nm] = {"gpu_vs_oracle32": e, "oracle32_vs_oracle64": noise}
        assert e <= max(tol, 2.0 * noise), f"{nm}: {e:.3e} (fp32 noise {noise:.3e})"
    e_gpu, e_ref = gpu.energy(), o32.exec("energy")
    rep["energy"] = [e_gpu, e_ref]
    assert abs(e_gpu - e_ref) <= 1e-5 * abs(e_ref)
    _report(f"ring_of_one_{n_side}_{drift}", rep)


def test_ring_of_one_slab_viscous_transport():
    """Taylor-Green as the reference runs it (viscous force, Re = 100, + transport-velocity correction) on the ring of one
    slab: bounds of tests/test_gpu_viscous_transport.py (the fp32 oracle's own distance to the fp64 oracle, >= 2e-4)."""
    from sphinxsys_b200.host import TaylorGreenCK
    case = _case(16, 1.5)
    mu, n_outer = 0.01, 8
    gpu = TaylorGreenCK(case, ring=True, mu_f=mu, transport_velocity=True)
    gpu.initialize()
    o32, _ = _oracle_on_gpu_mesh(case, gpu, viscosity=mu, transport_velocity=1)
    o64, _ = _oracle_on_gpu_mesh(case, gpu, f64=True, viscosity=mu, transport_velocity=1)
    for o in (o32, o64):
        o.exec("prepare_ck")
        o.exec("run_ck", 1e9, n_outer, 1e9, 0)
    n_ac = gpu.run_outer(n_outer)
    assert n_ac == int(o32.exec("acoustic_steps"))
    n, rep = case.n_fluid, {}
    for nm, w in (("Position", 3), ("Velocity", 3), ("Density", 1)):
        g = _own(gpu, nm, n).astype(np.float64).reshape(-1)
        r32, r64 = o32.real(nm, w).astype(np.float64), o64.real(nm, w).astype(np.float64)
        if nm == "Position":  # compare modulo the box
            g, r32 = g + np.round(r64 - g), r32 + np.round(r64 - r32)
        scale = max(np.max(np.abs(r64)), 1e-30)
        rep[nm] = (float(np.max(np.abs(g - r64)) / scale), float(np.max(np.abs(r32 - r64)) / scale))
        assert rep[nm][0] < max(4.0 * rep[nm][1], 2e-4), (nm, rep[nm])
    e_visc = gpu.energy()
    gpu.close()
    inviscid = TaylorGreenCK(case, ring=True)
    inviscid.initialize()
    inviscid.run_outer(n_outer)
    rep["energy_viscous_vs_inviscid"] = (e_visc, inviscid.energy())
    assert e_visc < rep["energy_viscous_vs_inviscid"][1]
    _report("ring_of_one_viscous_transport", rep)


def test_ring_of_one_slab_2d():
    """The 2-D Taylor-Green case of the reference (taylor_green.cpp) on the ring of one slab: x through the seam, y by images."""
    from sphinxsys_b200.host import TaylorGreenCK
    case = _case(40, 1.5, dim=2)
    n_outer = 15
    gpu = TaylorGreenCK(case, ring=True)
    gpu.initialize()
    o32, _ = _oracle_on_gpu_mesh(case, gpu)
    o64, _ = _oracle_on_gpu_mesh(case, gpu, f64=True)
    for o in (o32, o64):
        o.exec("prepare_ck")
        o.exec("run_ck", 1e9, n_outer, 1e9, 0)
    n_ac = gpu.run_outer(n_outer)
    assert n_ac == int(o32.exec("acoustic_steps"))
    same_path = int(o64.exec("acoustic_steps")) == n_ac
    n, rep = case.n_fluid, {}
    pos = _own(gpu, "Position", n)
    d = pos.astype(np.float64) - o32.real("Position", 3).reshape(-1, 3)
    d -= np.round(d)
    rep["Position"] = float(np.abs(d).max())
    assert rep["Position"] < 5e-6
    assert int((np.abs(pos[:, 0].astype(np.float64) - case.fluid_pos[:, 0]) > 0.5).sum()) > 0
    for nm, w, tol in (("Velocity", 3, 2e-4), ("Density", 1, 2e-6)):
        b = o32.real(nm, w).reshape(-1, w) if w > 1 else o32.real(nm, w)
        b64 = o64.real(nm, w).reshape(-1, w) if w > 1 else o64.real(nm, w)
        e, noise = rel_err(_own(gpu, nm, n), b), (rel_err(b, b64) if same_path else 0.0)
        rep[nm] = {"gpu_vs_oracle32": e, "oracle32_vs_oracle64": noise}
        assert e <= max(tol, 2.0 * noise), f"{nm}: {e:.3e} (fp32 noise {noise:.3e})"
    _report("ring_of_one_2d", rep)
