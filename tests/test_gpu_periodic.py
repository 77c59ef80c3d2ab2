"""GPU parity tests of the periodic-boundary path (BASELINE config 4: periodic Taylor-Green vortex): ghost (image)
particles behind the real ones + a second cell-linked list, against the oracle's restatement of the reference's
ghost list entries (domain_bounding.cpp:18-65). Everything runs through the C++ host layer and the C ABI.

Bar: bit-exact wrapped positions, image sets and neighbour SETS (the order inside a row differs by construction:
the device walks the real list, then the image list; the oracle walks one merged list); fields within the tolerances
written next to each assertion.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from helpers import gpu_field, lists_on_oracle_positions, oracle_field, rel_err  # noqa: E402

REPORT = {}


def _report(key, value):
    REPORT[key] = value
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        json.dump(REPORT, open(os.path.join(out, "periodic_report.json"), "w"), indent=1, default=float)
    except OSError:
        pass


@pytest.fixture(scope="module")
def ctx():
    from sphinxsys_b200 import capi
    assert torch.cuda.is_available(), "these tests need a GPU"
    c = capi.Context(0)
    yield c
    c.close()


def _p(t):
    return C.c_void_p(t.data_ptr())


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _oracle(case, f64=False):
    from oracle import oracle as orc
    return orc.OracleSim(case, f64=f64, free_surface=0)


def _gpu(case, **kw):
    from sphinxsys_b200.host import TaylorGreenCK
    return TaylorGreenCK(case, **kw)


def _sorted_rows(off, idx):
    off = off.astype(np.int64)
    return [np.sort(idx[off[i]:off[i + 1]]) for i in range(off.size - 1)]


# ------------------------------------------------------------------------------------------------------
# C ABI primitives
# ------------------------------------------------------------------------------------------------------
def _box(axes, lower, upper, cutoff):
    from sphinxsys_b200 import capi
    b = capi.PeriodicT()
    for d in range(3):
        b.lower[d], b.upper[d] = lower[d], upper[d]
    b.axes, b.cutoff = axes, cutoff
    return b


def test_periodic_bounding_and_images_primitives(ctx):
    rng = np.random.default_rng(11)
    n = 200_000
    lower, upper = np.array([0.0, -1.0, 0.5], np.float32), np.array([1.0, 1.0, 2.0], np.float32)
    L = upper - lower
    cutoff = np.float32(0.13)
    pos = np.zeros((n, 4), np.float32)
    pos[:, :3] = rng.uniform(lower - 0.3, upper + 0.3, size=(n, 3)).astype(np.float32)
    pos[0, :3] = lower          # exactly on the faces: untouched, no image (strict comparisons)
    pos[1, :3] = upper
    d = torch.from_numpy(pos).cuda()
    box = _box(0b101, lower, upper, cutoff)  # periodic in x and z only
    ctx.call("sphb200_periodic_bounding", C.byref(box), _p(d), n, _s())
    ref = pos.copy()
    for a in (0, 2):
        c = ref[:, a]
        lo, up = c < lower[a], c > upper[a]
        c[lo] = c[lo] + L[a]
        c[up & ~lo] = c[up & ~lo] - L[a]
    got = d.cpu().numpy()
    assert np.array_equal(got, ref)
    # images of the wrapped set
    cap = n
    ipos = torch.zeros((cap, 4), dtype=torch.float32, device="cuda")
    isrc = torch.zeros(cap, dtype=torch.int32, device="cuda")
    cnt = C.c_uint32(0)
    ctx.call("sphb200_periodic_images", C.byref(box), _p(d), n, _p(ipos), _p(isrc), cap, C.byref(cnt), _s())
    exp_pos, exp_src = [], []
    lo_edge, up_edge = lower + cutoff, upper - cutoff
    for i in range(n):
        x = ref[i, :3]
        opts = []
        for a in range(3):
            o = [0]
            if a != 1:
                if x[a] > lower[a] and x[a] < lo_edge[a]:
                    o.append(1)
                if x[a] < upper[a] and x[a] > up_edge[a]:
                    o.append(2)
            opts.append(o)
        if all(len(o) == 1 for o in opts):
            continue
        for cx in opts[0]:
            for cy in opts[1]:
                for cz in opts[2]:
                    if cx == cy == cz == 0:
                        continue
                    y = x.copy()
                    for a, c in enumerate((cx, cy, cz)):
                        if c == 1:
                            y[a] = x[a] + L[a]
                        elif c == 2:
                            y[a] = x[a] - L[a]
                    exp_pos.append(y)
                    exp_src.append(i)
    assert cnt.value == len(exp_src) > 0
    assert np.array_equal(isrc.cpu().numpy()[: cnt.value], np.array(exp_src, np.int32))
    assert np.array_equal(ipos.cpu().numpy()[: cnt.value, :3], np.array(exp_pos, np.float32))
    # capacity too small: count still reported, nothing written, error code SPHB200_E_CAPACITY
    from sphinxsys_b200 import capi
    cnt2 = C.c_uint32(0)
    rc = ctx.lib.sphb200_periodic_images(ctx._ctx, C.byref(box), _p(d), n, _p(ipos), _p(isrc), 10, C.byref(cnt2), _s())
    assert rc == -2 and cnt2.value == cnt.value


def test_ghost_copy_primitive(ctx):
    rng = np.random.default_rng(3)
    n_real, n_ghost = 5000, 777
    src = rng.integers(0, n_real, size=n_ghost).astype(np.int32)
    rec = rng.standard_normal((n_real + n_ghost, 8)).astype(np.float32)
    sca = rng.standard_normal(n_real + n_ghost).astype(np.float32)
    d_rec, d_sca, d_src = torch.from_numpy(rec).cuda(), torch.from_numpy(sca).cuda(), torch.from_numpy(src).cuda()
    ctx.call("sphb200_ghost_copy", _p(d_rec), 32, 12, 16, _p(d_src), n_real, n_ghost, _s())  # (Vol, v) of a 32-byte record
    ctx.call("sphb200_ghost_copy", _p(d_sca), 4, 0, 4, _p(d_src), n_real, n_ghost, _s())
    exp = rec.copy()
    exp[n_real:, 3:7] = rec[src, 3:7]
    assert np.array_equal(d_rec.cpu().numpy(), exp)
    exps = sca.copy()
    exps[n_real:] = sca[src]
    assert np.array_equal(d_sca.cpu().numpy(), exps)
    d_rec2 = torch.from_numpy(rec).cuda()
    ctx.call("sphb200_ghost_copy", _p(d_rec2), 32, 16, 16, _p(d_src), n_real, n_ghost, _s())  # aligned 16-byte path
    exp2 = rec.copy()
    exp2[n_real:, 4:8] = rec[src, 4:8]
    assert np.array_equal(d_rec2.cpu().numpy(), exp2)


# ------------------------------------------------------------------------------------------------------
# neighbour machinery with images
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n_side,jitter", [(3, 16, 0.0), (3, 20, 0.2), (2, 32, 0.2)])
def test_periodic_neighbour_sets(dim, n_side, jitter):
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=dim, n_side=n_side, jitter=jitter)
    gpu = _gpu(case)
    gpu.initialize()
    ref = _oracle(case)
    ref.exec("cell_list_fluid")
    ref.exec("relations")
    # one image per ghost list entry of the reference
    assert gpu.ghost_particles == ref.uint("fluid_ext_index").size - case.n_fluid
    off, idx = gpu.export_csr()
    assert off.size == case.n_fluid + 1
    assert np.array_equal(off, ref.uint("inner_offset"))
    rows_g = _sorted_rows(off, idx)
    rows_r = _sorted_rows(ref.uint("inner_offset"), ref.uint("inner_index"))
    assert all(np.array_equal(a, b) for a, b in zip(rows_g, rows_r))
    if jitter == 0.0:
        counts = np.diff(off.astype(np.int64))
        assert counts.min() == counts.max() == (80 if dim == 3 else 20)


def test_periodic_bounding_through_host_layer():
    """Particles pushed out of the box on every side come back on the other side, bit-identical to the oracle."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=3, n_side=16, jitter=0.1)
    rng = np.random.default_rng(5)
    shifted = (case.fluid_pos.astype(np.float64) + 0.08 * rng.uniform(-1, 1, size=case.fluid_pos.shape)).astype(np.float32)
    gpu = _gpu(case)
    gpu.upload("Position", shifted)
    gpu.exec("periodic_bounding")
    ref = _oracle(case)
    ref.real("Position", 3)[:] = shifted.reshape(-1)
    ref.exec("periodic_bounding")
    assert np.array_equal(gpu_field(gpu, "Position"), oracle_field(ref, "Position", 3))
    gpu.exec("update_configuration", 0)
    ref.exec("cell_list_fluid")
    ref.exec("relations")
    off, idx = gpu.export_csr()
    assert np.array_equal(off, ref.uint("inner_offset"))
    assert all(np.array_equal(a, b) for a, b in zip(_sorted_rows(off, idx), _sorted_rows(ref.uint("inner_offset"), ref.uint("inner_index"))))


# ------------------------------------------------------------------------------------------------------
# dynamics
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n_side", [(3, 20), (2, 40)])
def test_taylor_green_one_outer_step_per_dynamics(dim, n_side):
    """Each dynamics of one advection step against the double-precision oracle: 1e-5 of the field norm
    (pressure 3e-4: its fp32 granularity is p0 ulp(1))."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=dim, n_side=n_side, jitter=0.1)
    gpu = _gpu(case, fused_time_step=False)
    gpu.initialize()
    ref = _oracle(case, f64=True)
    ref.exec("cell_list_fluid")
    ref.exec("relations")

    def check(names, tag):
        for nm, w, tol in names:
            e = rel_err(gpu_field(gpu, nm), oracle_field(ref, nm, w))
            assert e < tol, f"{tag}: {nm} {e:.3e}"

    gpu.exec("density_summation"); ref.exec("compression_summation"); ref.exec("density_regularization")
    check([("CompressionSummation", 1, 1e-5), ("Compression", 1, 1e-5), ("Density", 1, 1e-5)], "summation")
    gpu.exec("advection_setup"); ref.exec("advection_setup")
    check([("VolumetricMeasure", 1, 1e-5)], "advection setup")
    dt_g, dt_r = gpu.exec("advection_dt"), ref.exec("advection_dt")
    assert abs(dt_g - dt_r) < 1e-6 * dt_r
    for k in range(3):
        dt = gpu.exec("acoustic_dt")
        assert abs(dt - ref.exec("acoustic_dt")) < 1e-5 * dt
        gpu.exec("acoustic1", dt); ref.exec("acoustic1", dt)
        check([("Pressure", 1, 3e-4), ("Force", 3, 2e-4), ("CompressionRate", 1, 2e-4), ("Velocity", 3, 1e-5),
               ("Displacement", 3, 1e-5)], f"1st half {k}")
        gpu.exec("acoustic2", dt); ref.exec("acoustic2", dt)
        check([("Force", 3, 2e-4), ("CompressionRate", 1, 2e-4), ("Compression", 1, 1e-5), ("Density", 1, 1e-5),
               ("Displacement", 3, 1e-5)], f"2nd half {k}")
    gpu.exec("update_position"); ref.exec("update_position")
    check([("Position", 3, 1e-6)], "update position")


@pytest.mark.parametrize("dim,n_side,n_outer", [(3, 20, 12), (2, 40, 20)])
def test_taylor_green_multi_step_drift(dim, n_side, n_outer):
    """The case loop on both sides (sort every 5 advection steps, particles crossing the periodic faces): fields within
    2e-4 (positions 5e-6) of the fp32 oracle or twice the fp32 oracle's own distance to the fp64 oracle; kinetic energy
    1e-5; neighbour sets identical at the end."""
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=dim, n_side=n_side, jitter=0.05)
    gpu = _gpu(case, sort_interval=5)
    gpu.initialize()
    o32, o64 = _oracle(case, f64=False), _oracle(case, f64=True)
    for o in (o32, o64):
        o.exec("prepare_ck")
        o.exec("run_ck", 1e9, n_outer, 1e9, 5)
    n_ac = gpu.run_outer(n_outer)
    assert int(o32.exec("acoustic_steps")) == n_ac
    same_path = int(o64.exec("acoustic_steps")) == n_ac
    assert np.array_equal(gpu_field(gpu, "OriginalID"), o32.uint("OriginalID"))
    rep = {}
    for nm, w, tol in (("Position", 3, 5e-6), ("Velocity", 3, 2e-4), ("Density", 1, 2e-6), ("Compression", 1, 2e-6)):
        e = rel_err(gpu_field(gpu, nm), oracle_field(o32, nm, w))
        noise = rel_err(oracle_field(o32, nm, w), oracle_field(o64, nm, w)) if same_path else 0.0
        rep[nm] = {"gpu_vs_oracle32": e, "oracle32_vs_oracle64": noise}
        assert e <= max(tol, 2.0 * noise), f"{nm}: {e:.3e} (fp32 noise {noise:.3e})"
    pos = gpu_field(gpu, "Position")
    assert pos[:, :dim].min() >= 0.0 and pos[:, :dim].max() <= 1.0
    e_gpu, e_ref = gpu.energy(), o32.exec("energy")
    rep["energy"] = [e_gpu, e_ref]
    assert abs(e_gpu - e_ref) <= 1e-5 * abs(e_ref)
    off, idx = lists_on_oracle_positions(gpu, o32, periodic=True)  # identical inputs: the oracle's end positions
    assert np.array_equal(off, o32.uint("inner_offset"))
    assert all(np.array_equal(a, b) for a, b in zip(_sorted_rows(off, idx), _sorted_rows(o32.uint("inner_offset"), o32.uint("inner_index"))))
    _report(f"taylor_green_drift_{dim}d", rep)


def test_taylor_green_generated_case_and_ghost_update():
    """The case generated by the C++ host layer itself (lattice + analytic initial condition): bulk neighbour counts,
    and ghost_update_ (all variables) leaves every image equal to its source."""
    gpu = _gpu(None, dim=3, n_side=16, generate=True)
    gpu.initialize()
    off, _ = gpu.export_csr()
    counts = np.diff(off.astype(np.int64))
    assert gpu.n_fluid == 16 ** 3 and counts.min() == counts.max() == 80
    n_ac = gpu.run_outer(3)
    assert n_ac >= 3
    gpu.exec("ghost_update")
    e = gpu.energy()
    assert 0.0 < e < 0.125 * 1.0001  # kinetic energy of the 3-D vortex: rho U^2 / 8 per unit volume, decaying
