"""Synthetic dam-break set-up: what the reference's pre-processing (lattice particle generator, body
shapes, wall normals) produces for the dam-break case files, generated directly with numpy.

Reference (relative to /root/reference):
  case parameters ...... tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/dambreak.cpp:11-51 (3-D),
                         tests/2d_examples/test_2d_dambreak/Dambreak.cpp:13-45 (2-D)
  lattice rule ......... src/for_3D_build/particle_generator/particle_generator_lattice_3d.cpp:12-25
                         (cell centres of Mesh(system bounds, dp, 0); loop order i -> j -> k)
  system bounds ........ src/shared/sphinxsys_system/sph_system.cpp:39 (case bounds expanded by 4 dp)
  box containment ...... src/shared/geometries/geometric_element.h:43-55 (|x_local| <= halfsize)
  wall normal .......... src/shared/geometries/base_geometry.cpp:45-59,118-140,
                         src/shared/geometries/geometric_element.cpp:19-61
  initial state ........ src/shared/materials/base_material.cpp:37-40 (rho = rho0, m = rho0 Vol)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import hostmath as hm


@dataclass
class DamBreakCase:
    dim: int
    dp: float
    dtype: type
    # geometry
    DL: float
    DH: float
    DW: float
    LL: float
    LH: float
    LW: float
    BW: float
    # material / numerics
    rho0: float
    gravity_g: float
    U_ref: float
    c0: float
    h: float
    # particles (packed [n, 3], z = 0 in 2-D)
    fluid_pos: np.ndarray = field(repr=False, default=None)
    wall_pos: np.ndarray = field(repr=False, default=None)
    wall_normal: np.ndarray = field(repr=False, default=None)
    vol: float = 0.0
    # derived PODs
    mesh: hm.MeshSpec = None
    kernel: hm.KernelSpec = None
    sigma0: float = 0.0
    # periodic box (Taylor-Green): bit d of periodic_axes = periodic along axis d; bounds already rounded to `dtype`
    periodic_axes: int = 0
    periodic_lower: tuple = (0.0, 0.0, 0.0)
    periodic_upper: tuple = (0.0, 0.0, 0.0)
    fluid_vel: np.ndarray = field(repr=False, default=None)
    system_lower: tuple = None
    system_upper: tuple = None

    @property
    def n_fluid(self):
        return self.fluid_pos.shape[0]

    @property
    def n_wall(self):
        return self.wall_pos.shape[0]

    @property
    def gravity(self):
        return (0.0, -self.gravity_g, 0.0)


def _lattice_axis(lower, upper, dp, R):
    """Cell-centre coordinates of Mesh(bounds, dp, buffer 0) along one axis, in Real arithmetic:
    all_cells = ceil((upper - lower)/dp); x_i = (lower + i*dp) + 0.5*dp."""
    lower = R(lower)
    upper = R(upper)
    dp = R(dp)
    n = int(np.ceil(R((upper - lower)) / dp))
    i = np.arange(n).astype(R)
    return ((lower + i * dp).astype(R) + R(0.5) * dp).astype(R)


def _box_closest_point(p, center, half):
    """GeometricBox::findClosestPoint in the box frame (geometric_element.cpp:19-61), vectorised."""
    c = p - center
    clamped = np.clip(c, -half, half)
    outside = np.any(np.abs(c) > half, axis=1)
    # inside: project to the nearest side (first axis wins ties: strict '<' in the reference loop)
    d_to_side = half - np.abs(c)
    which = np.argmin(d_to_side, axis=1)  # argmin returns the first minimum, as the reference's loop does
    proj = c.copy()
    rows = np.arange(c.shape[0])
    sign = np.where(c[rows, which] < 0, -1.0, 1.0)
    proj[rows, which] = sign * half[which]
    out = np.where(outside[:, None], clamped, proj)
    return out + center


def dam_break(dim=3, dp=0.05, dtype=np.float32, width_scale=1.0, kernel_kind=hm.KERNEL_WENDLAND_C2) -> DamBreakCase:
    """Build the dam-break particle set.

    dim=3: tank 5.366 x 2 x 0.5*width_scale, water 2 x 1 x 0.5*width_scale (dambreak.cpp:11-18).
    dim=2: tank 5.366 x 5.366, water 2 x 1 (Dambreak.cpp:13-18).
    `width_scale` stretches the z extent (DW, LW) for weak-scaling runs.
    """
    R = dtype
    DL = 5.366
    LL, LH = 2.0, 1.0
    if dim == 3:
        DH, DW, LW = 2.0, 0.5 * width_scale, 0.5 * width_scale
    else:
        DH, DW, LW = 5.366, 0.0, 0.0
    BW = dp * 4
    rho0, g = 1.0, 1.0
    U_ref = 2.0 * np.sqrt(g * LH)
    c0 = 10.0 * U_ref
    h = 1.3 * dp
    case = DamBreakCase(dim, dp, R, DL, DH, DW, LL, LH, LW, BW, rho0, g, float(U_ref), float(c0), h)

    ext = [DL, DH, DW][:dim]
    case_lo = np.array([-BW] * dim)
    case_up = np.array([e + BW for e in ext])
    sys_lo = case_lo - 4 * dp
    sys_up = case_up + 4 * dp

    axes = [_lattice_axis(sys_lo[d], sys_up[d], dp, R) for d in range(dim)]

    def in_box(axis_vals, lo, up):
        # closed-box test on the lattice coordinate
        return (axis_vals >= lo) & (axis_vals <= up)

    # ---- water block: box [0,LL] x [0,LH] (x [0,LW]) ----
    wlim = [(0.0, LL), (0.0, LH), (0.0, LW)][:dim]
    wsel = [axes[d][in_box(axes[d], *wlim[d])] for d in range(dim)]
    grids = np.meshgrid(*wsel, indexing="ij")  # i -> j -> k order == C order of 'ij' meshgrid
    fpos = np.stack([g_.reshape(-1) for g_ in grids], axis=1).astype(R)

    # ---- wall: outer box minus inner box ----
    olim = [(-BW, ext[d] + BW) for d in range(dim)]
    osel = [axes[d][in_box(axes[d], *olim[d])] for d in range(dim)]
    og = np.meshgrid(*osel, indexing="ij")
    opos = np.stack([g_.reshape(-1) for g_ in og], axis=1).astype(R)
    inner_mask = np.ones(opos.shape[0], dtype=bool)
    for d in range(dim):
        inner_mask &= (opos[:, d] >= 0.0) & (opos[:, d] <= ext[d])
    wpos = opos[~inner_mask]

    # ---- wall normals: toward/away from the closest point on {outer box, inner box} surfaces ----
    p64 = wpos.astype(np.float64)
    half_in = np.array([0.5 * e for e in ext])
    half_out = half_in + BW
    center = half_in.copy()
    cp_out = _box_closest_point(p64, center, half_out)
    cp_in = _box_closest_point(p64, center, half_in)
    d_out = np.linalg.norm(p64 - cp_out, axis=1)
    d_in = np.linalg.norm(p64 - cp_in, axis=1)
    use_in = d_in <= d_out  # later sub-shape wins ties (base_geometry.cpp:131 '<=')
    cp = np.where(use_in[:, None], cp_in, cp_out)
    disp = cp - p64
    nrm = np.linalg.norm(disp, axis=1, keepdims=True)
    nrm[nrm == 0] = 1.0
    normal = disp / nrm  # wall particles are contained in the shape: direction_to_surface

    def embed(a):
        if dim == 3:
            return np.ascontiguousarray(a.astype(R))
        out = np.zeros((a.shape[0], 3), dtype=R)
        out[:, :2] = a
        return out

    case.fluid_pos = embed(fpos)
    case.wall_pos = embed(wpos)
    case.wall_normal = embed(normal)
    case.vol = float(R(dp) ** dim)
    case.kernel = hm.make_kernel(h, dim, kernel_kind, dtype=R)
    case.mesh = hm.make_mesh(sys_lo, sys_up, case.kernel.cutoff, 2, dtype=R)
    case.sigma0 = hm.lattice_number_density(case.kernel, dp)
    return case


def fluid_block(n_side, dp=0.01, jitter=0.0, seed=1, dtype=np.float32, dim=3):
    """A jittered cubic lattice block without walls (neighbour micro-benchmark / unit tests)."""
    rng = np.random.default_rng(seed)
    ax = (np.arange(n_side) + 0.5) * dp
    if dim == 3:
        g_ = np.meshgrid(ax, ax, ax, indexing="ij")
    else:
        g_ = np.meshgrid(ax, ax, indexing="ij")
    pos = np.stack([a.reshape(-1) for a in g_], axis=1)
    pos = pos + jitter * dp * rng.uniform(-1, 1, size=pos.shape)
    if dim == 2:
        pos = np.concatenate([pos, np.zeros((pos.shape[0], 1))], axis=1)
    return np.ascontiguousarray(pos.astype(dtype))


def random_block(n, seed=1, dp=0.01, dtype=np.float32, particles_per_cell=17.6):
    """Config 5 of BASELINE.json (neighbour-search micro-benchmark): n positions uniform-random in a cube sized so that
    a cell of the cell-linked list (edge = cut-off radius 2.6 dp) holds `particles_per_cell` particles on average.
    Returned as a DamBreakCase-shaped object without walls so the oracle and the host layer can both consume it."""
    R = dtype
    h = 1.3 * dp
    kernel = hm.make_kernel(h, 3, hm.KERNEL_WENDLAND_C2, dtype=R)
    cell = kernel.cutoff
    edge = cell * (n / particles_per_cell) ** (1.0 / 3.0)
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0.0, edge, size=(n, 3)).astype(R)
    case = DamBreakCase(3, dp, R, edge, edge, edge, edge, edge, edge, 0.0, 1.0, 1.0, 2.0, 20.0, h)
    case.fluid_pos = np.ascontiguousarray(pos)
    case.wall_pos = np.zeros((0, 3), dtype=R)
    case.wall_normal = np.zeros((0, 3), dtype=R)
    case.vol = float(R(dp) ** 3)
    case.kernel = kernel
    case.mesh = hm.make_mesh(np.zeros(3), np.full(3, edge), kernel.cutoff, 2, dtype=R)
    case.sigma0 = hm.lattice_number_density(kernel, dp)
    return case


def taylor_green(dim=3, n_side=32, jitter=0.05, seed=2024, dtype=np.float32, L=1.0, U=1.0, x_scale=1) -> DamBreakCase:
    """Periodic Taylor-Green vortex (BASELINE config 4): n_side^dim particles on the lattice of the periodic box
    [0, L]^dim with a deterministic jitter of `jitter` dp, no walls, no gravity, c0 = 10 U.

    Reference: tests/2d_examples/test_2d_taylor_green/taylor_green.cpp:14-57 (box, material, initial condition; the 3-D
    velocity field is the usual extension v = U (sin 2pi x cos 2pi y cos 2pi z, -cos 2pi x sin 2pi y cos 2pi z, 0)).
    Returned in the DamBreakCase shape (without wall particles) so the oracle and the host layer both consume it.
    `x_scale` (integer) replicates the box along x: [0, x_scale L] x [0, L]^(dim-1), x_scale n_side^dim particles, the
    weak-scaling shape of config 4; the velocity field keeps its period L."""
    R = dtype
    dp = L / n_side
    h = 1.3 * dp
    Lx = L * int(x_scale)
    case = DamBreakCase(dim, dp, R, Lx, L, L if dim == 3 else 0.0, Lx, L, L if dim == 3 else 0.0, 0.0, 1.0, 0.0, float(U), 10.0 * U, h)
    ext = [Lx, L, L][:dim]
    sys_lo = np.array([0.0 - 4 * dp] * dim)
    sys_up = np.array([e + 4 * dp for e in ext])
    axes = [_lattice_axis(sys_lo[d], sys_up[d], dp, R) for d in range(dim)]
    sel = [a[(a >= 0.0) & (a <= ext[d])] for d, a in enumerate(axes)]
    grids = np.meshgrid(*sel, indexing="ij")
    pos = np.stack([g_.reshape(-1) for g_ in grids], axis=1).astype(np.float64)
    if jitter:
        rng = np.random.default_rng(seed)
        pos = pos + jitter * dp * rng.uniform(-1.0, 1.0, size=pos.shape)
    if dim == 2:
        pos = np.concatenate([pos, np.zeros((pos.shape[0], 1))], axis=1)
    pos = np.ascontiguousarray(pos.astype(R))
    x, y, z = (pos[:, k].astype(np.float64) for k in range(3))
    tp = 2.0 * np.pi
    vel = np.zeros_like(pos, dtype=np.float64)
    if dim == 2:
        vel[:, 0] = -U * np.cos(tp * x) * np.sin(tp * y)
        vel[:, 1] = U * np.sin(tp * x) * np.cos(tp * y)
    else:
        vel[:, 0] = U * np.sin(tp * x) * np.cos(tp * y) * np.cos(tp * z)
        vel[:, 1] = -U * np.cos(tp * x) * np.sin(tp * y) * np.cos(tp * z)
    case.fluid_pos = pos
    case.fluid_vel = np.ascontiguousarray(vel.astype(R))
    case.wall_pos = np.zeros((0, 3), dtype=R)
    case.wall_normal = np.zeros((0, 3), dtype=R)
    case.vol = float(R(dp) ** dim)
    case.kernel = hm.make_kernel(h, dim, hm.KERNEL_WENDLAND_C2, dtype=R)
    case.mesh = hm.make_mesh(sys_lo, sys_up, case.kernel.cutoff, 2, dtype=R)
    case.sigma0 = hm.lattice_number_density(case.kernel, dp)
    case.periodic_axes = (1 << dim) - 1
    case.periodic_lower = (0.0, 0.0, 0.0)
    case.periodic_upper = tuple(float(R(ext[d])) if d < dim else 0.0 for d in range(3))
    case.system_lower = tuple(float(R(v)) for v in sys_lo) + ((0.0,) if dim == 2 else ())
    case.system_upper = tuple(float(R(v)) for v in sys_up) + ((0.0,) if dim == 2 else ())
    return case
