"""Shared helpers for the parity tests: build the oracle simulation and the GPU solver on identical inputs."""
from __future__ import annotations

import numpy as np

FLUID_REAL = ["VolumetricMeasure", "Mass", "Density", "Pressure", "Compression", "CompressionRate", "CompressionSummation"]
FLUID_VEC = ["Position", "Velocity", "Displacement", "Force", "ForcePrior"]


def rel_err(a, b):
    """max |a-b| / max |b| (field-norm relative error)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / scale


def make_oracle(case, f64=False, correction=0, riemann=1):
    from oracle import oracle as orc
    return orc.OracleSim(case, f64=f64, correction=correction, riemann=riemann)


def make_gpu(case, correction=False, fused_time_step=True, sort_interval=100, relation_stride=None, legacy=False, riemann=1):
    """The C++ host layer (include/sphinxsys_ck) on the same particle arrays the oracle gets."""
    from sphinxsys_b200.host import DamBreakCK
    return DamBreakCK(case, correction=correction, fused_time_step=fused_time_step, sort_interval=sort_interval,
                      relation_stride=relation_stride, legacy=legacy, riemann=riemann)


def oracle_field(sim, name, width=1):
    a = sim.real(name, width).copy()
    return a.reshape(-1, 3) if width == 3 else (a.reshape(-1, 9) if width == 9 else a)


def gpu_field(solver, name):
    return solver.download(name)


def perturb_state(case, seed=7, vel_scale=0.3, jitter=0.15):
    """A developed-looking state: jittered positions, smooth + random velocity field (same arrays for both sides)."""
    rng = np.random.default_rng(seed)
    pos = case.fluid_pos.astype(np.float64).copy()
    d = case.dim
    pos[:, :d] += jitter * case.dp * rng.uniform(-1, 1, size=(pos.shape[0], d))
    vel = np.zeros_like(pos)
    vel[:, 0] = vel_scale * np.sin(2.0 * pos[:, 1]) + 0.05 * rng.standard_normal(pos.shape[0])
    vel[:, 1] = -vel_scale * np.cos(1.5 * pos[:, 0]) + 0.05 * rng.standard_normal(pos.shape[0])
    if d == 3:
        vel[:, 2] = 0.1 * vel_scale * np.sin(3.0 * pos[:, 0]) + 0.05 * rng.standard_normal(pos.shape[0])
    return pos.astype(case.dtype), vel.astype(case.dtype)


def dtw_distance(a, b, window=5):
    """Restatement of RegressionTestDynamicTimeWarping::calculateDTWDistance (dynamic_time_warping_method.hpp:17-55),
    including its window handling (cells outside the band stay 0)."""
    la, lb = len(a), len(b)
    assert 0.8 * la <= lb <= 1.2 * la
    D = np.zeros((la, lb))
    D[0, 0] = abs(a[0] - b[0])
    for i in range(1, la):
        D[i, 0] = D[i - 1, 0] + abs(a[i] - b[0])
    for j in range(1, lb):
        D[0, j] = D[0, j - 1] + abs(a[0] - b[j])
    w = max(window, abs(la - lb))
    for i in range(1, la):
        for j in range(max(1, i - w), min(lb, i + w)):
            D[i, j] = abs(a[i] - b[j]) + min(D[i - 1, j], D[i, j - 1], D[i - 1, j - 1])
    return float(D[la - 1, lb - 1])


def dtw_two_sided(ours, ref_run, window=5):
    """The reference compares a new run with a stored one as calculateDTWDistance(current, stored) (updateDTWDistance,
    dynamic_time_warping_method.hpp:86-99). Its band leaves the corner cell at 0 when the first series is SHORTER than the second
    by exactly the band width (the loop bound `j != min(b_length, i + window)` is exclusive) — the check is then vacuous. Series
    of different lengths are therefore compared in both argument orders and the larger distance counts: never weaker than the
    reference's own check, never vacuous."""
    return max(dtw_distance(ours, ref_run, window), dtw_distance(ref_run, ours, window))


def lists_on_oracle_positions(gpu, o32, periodic=False):
    """Neighbour lists of the END state of a multi-step run, on IDENTICAL inputs: after many steps the two fp32 paths differ
    by rounding (a few 1e-7 in position), so a pair that sits within that distance of the cut-off may be a neighbour on one
    side and not on the other — with 10^5..10^8 pairs some always do. The bar is bit-exact sets on the same inputs: hand the
    oracle's end positions to the GPU path (reference particle order), rebuild cell list + relation there, then compare."""
    gpu.upload("Position", oracle_field(o32, "Position", 3))
    if periodic:
        gpu.exec("update_configuration", 0.0)  # cell list, periodic images, relation (positions are bounded already)
    else:
        gpu.exec("cell_list_fluid")
        gpu.exec("relations")
    return gpu.export_csr()
