"""Host-side construction of the PODs that the reference builds on the host and copies by value
into every computing kernel: mesh geometry, smoothing-kernel table, Riemann/EOS constants.

They are computed ONCE, in `Real` precision, and handed unchanged to whoever consumes them (the CUDA
library through the C ABI; the CPU oracle in tests) so integer outputs can be compared bit for bit
(SURVEY.md Appendix A).

Reference (paths relative to /root/reference/src/shared):
  mesh geometry ......... meshes/base_mesh.cpp:6-16, sphinxsys_system/sph_system.cpp:39,
                          meshes/cell_linked_list.cpp:14 (buffer width 2), adaptations/adaptation.cpp:80-85
  kernel table .......... shared_ck/smoothing_kernel/kernel_tabulated_ck.cpp:6-25,
                          kernels/kernel_wendland_c2.cpp:8-50, kernels/kernel_laguerre_gauss.cpp:8-50,
                          kernels/base_kernel.h:87-93
  lattice number density  adaptations/adaptation.cpp:26-60
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

KERNEL_WENDLAND_C2 = 0
KERNEL_LAGUERRE_GAUSS = 1
KERNEL_RESOLUTION = 20
TABULATED_SIZE = KERNEL_RESOLUTION + 4


@dataclass
class MeshSpec:
    """`Mesh` POD: lower bound, spacing, number of cells per axis (all_cells = all_grid_points - 1)."""
    lower: tuple
    spacing: float
    cells: tuple

    @property
    def total_cells(self) -> int:
        return int(self.cells[0]) * int(self.cells[1]) * int(self.cells[2])


def make_mesh(bounds_lower, bounds_upper, spacing, buffer_width=2, dtype=np.float32) -> MeshSpec:
    """Mesh::Mesh(tentative_bounds, grid_spacing, buffer_width) evaluated in `dtype` precision.

    2-D inputs (length-2 bounds) are embedded with one cell layer in z (lower_z = 0)."""
    R = dtype
    lo = np.asarray(bounds_lower, dtype=R)
    up = np.asarray(bounds_upper, dtype=R)
    dim = lo.shape[0]
    sp = R(spacing)
    mesh_buffer = R(buffer_width) * sp
    lower = (lo - mesh_buffer).astype(R)
    tentative = ((up + mesh_buffer).astype(R) - lower).astype(R)
    grid_pts = np.ceil((tentative / sp).astype(R)).astype(np.int64) + 1
    cells = grid_pts - 1
    if dim == 2:
        lower = np.array([lower[0], lower[1], R(0)], dtype=R)
        cells = np.array([cells[0], cells[1], 1])
    return MeshSpec(tuple(float(v) for v in lower), float(sp), tuple(int(c) for c in cells))


@dataclass
class KernelSpec:
    """KernelTabulatedCK POD plus the scalars Neighbor<SPHAdaptation,SPHAdaptation>::SmoothingKernel holds."""
    dim: int
    kind: int
    h: float
    kernel_size: float
    dimension_factor: float
    w: np.ndarray = field(repr=False)
    dw: np.ndarray = field(repr=False)

    @property
    def inv_h(self):
        return 1.0 / self.h

    @property
    def cutoff(self):
        return self.kernel_size * self.h


def _w1d(kind, q):
    if kind == KERNEL_WENDLAND_C2:
        return (1.0 - 0.5 * q) ** 4 * (1.0 + 2.0 * q)
    return (1.0 - q ** 2 + q ** 4 / 6.0) * np.exp(-(q ** 2))


def _dw1d(kind, q):
    if kind == KERNEL_WENDLAND_C2:
        return 0.625 * (q - 2.0) ** 3 * q
    return (-(q ** 5) / 3.0 + 8.0 * q ** 3 / 3.0 - 4.0 * q) * np.exp(-(q ** 2))


def _sigma(kind, dim):
    if kind == KERNEL_WENDLAND_C2:
        return {1: 3.0 / 4.0, 2: 7.0 / (4.0 * math.pi), 3: 21.0 / (16.0 * math.pi)}[dim]
    return {1: 8.0 / (5.0 * math.sqrt(math.pi)), 2: 3.0 / math.pi, 3: 8.0 / math.pi ** 1.5}[dim]


def make_kernel(h, dim, kind=KERNEL_WENDLAND_C2, dtype=np.float32) -> KernelSpec:
    """Tabulate W_1D/dW_1D at q = (k-1)*dq, k = 0..23, dq = kernel_size/20.

    The reference evaluates the analytic forms with double literals and stores the result as `Real`
    (kernel_wendland_c2.cpp:17-30), i.e. double evaluation rounded to Real; note the table deliberately
    samples the unsymmetrised polynomial at q = -dq and beyond the support (q = 2.1, 2.2)."""
    R = dtype
    h_r = R(h)
    inv_h = R(1.0) / h_r
    ks = R(2.0)
    dq = R(ks / R(KERNEL_RESOLUTION))
    q = np.array([float(R(k - 1) * dq) for k in range(TABULATED_SIZE)], dtype=np.float64)
    w = _w1d(kind, q).astype(R)
    dw = _dw1d(kind, q).astype(R)
    # factor_W_dim = inv_h^dim * sigma ; DimensionFactor = factor_W_dim * h^dim (base_kernel.h:91-93)
    factor = R(inv_h ** dim * R(_sigma(kind, dim)))
    dimension_factor = R(factor * h_r ** dim)
    return KernelSpec(dim, kind, float(h_r), float(ks), float(dimension_factor), w, dw)


def kernel_W(spec: KernelSpec, r):
    """Closed-form (not tabulated) W(r) in double — used by tests and lattice number density."""
    q = np.asarray(r, dtype=np.float64) / spec.h
    return _sigma(spec.kind, spec.dim) / spec.h ** spec.dim * _w1d(spec.kind, q)


def lattice_number_density(spec: KernelSpec, dp) -> float:
    """SPHAdaptation::computeLatticeNumberDensity (adaptation.cpp:26-60)."""
    rc = spec.cutoff
    depth = int(rc / dp) + 1
    rng = np.arange(-depth, depth + 1) * dp
    if spec.dim == 2:
        X, Y = np.meshgrid(rng, rng, indexing="ij")
        d = np.sqrt(X ** 2 + Y ** 2)
    else:
        X, Y, Z = np.meshgrid(rng, rng, rng, indexing="ij")
        d = np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    mask = d < rc
    return float(np.sum(kernel_W(spec, d[mask])))


@dataclass
class FluidSpec:
    """WeaklyCompressibleFluid + Riemann solver constants (materials/weakly_compressible_fluid.cpp:8-12,
    shared_ck/.../riemann_solver_ck.hpp:58-69)."""
    rho0: float
    c0: float
    riemann: int = 1          # 0 NoRiemann, 1 AcousticRiemann (TruncatedLinear 3.0), 2 Dissipative
    correction: int = 0       # 0 NoKernelCorrection, 1 LinearCorrection
    free_surface: int = 1
    limiter_coeff: float = 3.0
