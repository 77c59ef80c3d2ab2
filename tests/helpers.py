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


def make_gpu(case, correction=False, fused_time_step=True, sort_interval=100, relation_stride=None, legacy=False):
    """The C++ host layer (include/sphinxsys_ck) on the same particle arrays the oracle gets."""
    from sphinxsys_b200.host import DamBreakCK
    return DamBreakCK(case, correction=correction, fused_time_step=fused_time_step, sort_interval=sort_interval,
                      relation_stride=relation_stride, legacy=legacy)


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
