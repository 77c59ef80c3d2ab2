"""CPU tests of the oracle's ViscousForceCK / KernelGradientIntegral / TransportVelocityCorrectionCK restatements
 through properties the reference's formulas imply (viscous_force.hpp:44-103,
kernel_gradient_integral.hpp:33-78, transport_velocity_correction_ck.hpp:39-50)."""
import numpy as np
import pytest


def _tg(dim=2, n_side=24, jitter=0.05, **kw):
    from oracle import oracle as orc
    from sphinxsys_b200 import cases
    case = cases.taylor_green(dim=dim, n_side=n_side, jitter=jitter, dtype=np.float64)
    s = orc.OracleSim(case, f64=True, free_surface=0, **kw)
    s.exec("prepare_ck")
    return case, s


@pytest.mark.parametrize("dim,n_side", [(2, 24), (3, 12)])
def test_viscous_force_properties(dim, n_side):
    case, s = _tg(dim, n_side, viscosity=0.01)
    n = case.n_fluid
    # 1. a uniform velocity field feels no viscous force
    s.real("Velocity", 3)[:] = np.tile([0.3, -0.2, 0.1 if dim == 3 else 0.0], n)
    s.exec("viscous_force")
    assert np.max(np.abs(s.real("ViscousForce", 3))) < 1e-15
    # 2. pair forces are antisymmetric: total momentum is conserved in the periodic box
    s.real("Velocity", 3)[:] = case.fluid_vel.astype(np.float64).reshape(-1)
    fp0 = s.real("ForcePrior", 3).copy()
    s.exec("viscous_force")
    F = s.real("ViscousForce", 3).reshape(-1, 3).copy()
    assert np.max(np.abs(F)) > 0
    assert np.max(np.abs(F.sum(axis=0))) < 1e-12 * np.abs(F).sum()
    # 3. it dissipates: sum F.v < 0
    v = s.real("Velocity", 3).reshape(-1, 3)
    assert float(np.sum(F * v)) < 0
    # 4. ForcePriorCK difference form: first call adds F, a second call on the same state adds nothing
    assert np.allclose(s.real("ForcePrior", 3) - fp0, F.reshape(-1), atol=1e-18)
    fp1 = s.real("ForcePrior", 3).copy()
    s.exec("viscous_force")
    assert np.allclose(s.real("ForcePrior", 3), fp1, atol=1e-18)
    # 5. linear in mu: the Laplacian-like operator of the Taylor-Green field is ~ -2 (2 pi)^2 mu v (dim = 2), loosely
    if dim == 2:
        m = s.real("Mass")
        acc = F / m[:, None]
        k2 = 2 * (2 * np.pi) ** 2
        ratio = np.sum(acc * v) / np.sum(v * v) / (-0.01 * k2)
        assert 0.7 < ratio < 1.3, ratio


def test_kernel_gradient_integral_and_transport_correction():
    # perfect lattice: the kernel gradient integral vanishes by symmetry
    case, s = _tg(2, 24, jitter=0.0, transport_velocity=1)
    s.exec("kernel_gradient_integral")
    g = s.real("KernelGradientIntegral", 3)
    assert np.max(np.abs(g)) < 1e-9 * 24
    # jittered lattice: non-zero, and the correction moves particles by coefficient h^2 limiter(h^2 |g|^2) g
    case, s = _tg(2, 24, jitter=0.2, transport_velocity=1)
    s.exec("kernel_gradient_integral")
    g = s.real("KernelGradientIntegral", 3).reshape(-1, 3).copy()
    assert np.max(np.abs(g)) > 1e-3
    d0 = s.real("Displacement", 3).reshape(-1, 3).copy()
    s.exec("transport_velocity_correction", 1, 0)
    d1 = s.real("Displacement", 3).reshape(-1, 3).copy()
    h = case.kernel.h
    lim = np.minimum(100.0 * h * h * np.sum(g * g, axis=1), 1.0)
    assert np.allclose(d1 - d0, (0.2 * h * h * lim)[:, None] * g, rtol=1e-12, atol=1e-18)
    # NoLimiter spelling
    s.exec("transport_velocity_correction", 0, 0)
    d2 = s.real("Displacement", 3).reshape(-1, 3).copy()
    assert np.allclose(d2 - d1, 0.2 * h * h * g, rtol=1e-12, atol=1e-18)
    # the correction reduces the particle disorder: repeating integral -> shift lowers |g|
    pos = s.real("Position", 3)
    pos[:] += (d2 - d0).reshape(-1)
    s.exec("cell_list_fluid")
    s.exec("relations")
    s.exec("kernel_gradient_integral")
    assert np.linalg.norm(s.real("KernelGradientIntegral", 3)) < np.linalg.norm(g)
