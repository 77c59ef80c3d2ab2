"""ctypes binding of the CPU oracle (oracle/sph_oracle.cpp).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package `sphinxsys_b200` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "sph_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class MeshPOD(C.Structure):
    _fields_ = [("lower", C.c_double * 3), ("spacing", C.c_double), ("cells", C.c_int * 3)]


class KernelPOD(C.Structure):
    _fields_ = [("dim", C.c_int), ("kind", C.c_int), ("h", C.c_double), ("kernel_size", C.c_double),
                ("dimension_factor", C.c_double), ("w", C.c_double * 24), ("dw", C.c_double * 24)]


class ParamsPOD(C.Structure):
    _fields_ = [("dim", C.c_int), ("riemann", C.c_int), ("correction", C.c_int), ("free_surface", C.c_int),
                ("rho0", C.c_double), ("c0", C.c_double), ("gravity", C.c_double * 3), ("U_ref", C.c_double),
                ("h_min", C.c_double), ("acoustic_cfl", C.c_double), ("advection_cfl", C.c_double),
                ("correction_alpha", C.c_double), ("sigma0", C.c_double), ("wall_rho0", C.c_double),
                ("contact_depth", C.c_int), ("threads", C.c_int), ("periodic_axes", C.c_int),
                ("periodic_lower", C.c_double * 3), ("periodic_upper", C.c_double * 3), ("periodic_cutoff", C.c_double),
                ("surface_indicator", C.c_int), ("viscosity", C.c_double), ("transport_velocity", C.c_int),
                ("transport_coefficient", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.POINTER(ParamsPOD), C.POINTER(KernelPOD), C.POINTER(MeshPOD),
                                 C.POINTER(MeshPOD), C.c_uint32, C.c_uint32]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_real.restype = C.c_void_p
        L.orc_real.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_uint.restype = C.c_void_p
        L.orc_uint.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        L.orc_exec.restype = C.c_double
        L.orc_exec.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double]
        L.orc_series.restype = C.c_uint64
        L.orc_series.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_probe_series.restype = C.c_uint64
        L.orc_probe_series.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_exclusive_scan_u32.restype = C.c_uint32
        L.orc_exclusive_scan_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_cell_keys.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.POINTER(MeshPOD), C.c_void_p, C.c_void_p]
        L.orc_sort_pairs_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_kernel_eval.restype = C.c_double
        L.orc_kernel_eval.argtypes = [C.c_int, C.POINTER(KernelPOD), C.c_int, C.c_double]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def mesh_pod(mesh) -> MeshPOD:
    m = MeshPOD()
    for d in range(3):
        m.lower[d] = mesh.lower[d]
        m.cells[d] = mesh.cells[d]
    m.spacing = mesh.spacing
    return m


def kernel_pod(k) -> KernelPOD:
    p = KernelPOD()
    p.dim, p.kind, p.h, p.kernel_size, p.dimension_factor = k.dim, k.kind, k.h, k.kernel_size, k.dimension_factor
    for i in range(24):
        p.w[i] = float(k.w[i])
        p.dw[i] = float(k.dw[i])
    return p


def exclusive_scan(values: np.ndarray):
    v = np.ascontiguousarray(values, dtype=np.uint32)
    out = np.empty_like(v)
    last = lib().orc_exclusive_scan_u32(v.ctypes.data, out.ctypes.data, v.size)
    return out, int(last)


def cell_keys(pos: np.ndarray, mesh):
    pos = np.ascontiguousarray(pos)
    f64 = int(pos.dtype == np.float64)
    n = pos.shape[0]
    cell = np.empty(n, dtype=np.uint32)
    key = np.empty(n, dtype=np.uint32)
    mp = mesh_pod(mesh)
    lib().orc_cell_keys(f64, pos.ctypes.data, n, C.byref(mp), cell.ctypes.data, key.ctypes.data)
    return cell, key


def sort_pairs(keys: np.ndarray, vals: np.ndarray):
    k = np.ascontiguousarray(keys, dtype=np.uint32).copy()
    v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    lib().orc_sort_pairs_u32(k.ctypes.data, v.ctypes.data, k.size)
    return k, v


def kernel_eval(kspec, which, q, f64=False):
    kp = kernel_pod(kspec)
    return lib().orc_kernel_eval(int(f64), C.byref(kp), which, float(q))


class OracleSim:
    """One fluid body + one wall body + inner/contact relations, advanced by the oracle."""

    def __init__(self, case, f64=False, riemann=1, correction=0, free_surface=1, threads=0, contact_depth=1,
                 surface_indicator=0, observers=None, viscosity=0.0, transport_velocity=0, correction_alpha=0.5):
        self.case = case
        self.f64 = bool(f64)
        self.dtype = np.float64 if f64 else np.float32
        p = ParamsPOD()
        p.dim, p.riemann, p.correction, p.free_surface = case.dim, riemann, correction, free_surface
        p.rho0, p.c0 = case.rho0, case.c0
        for d in range(3):
            p.gravity[d] = case.gravity[d]
        p.U_ref, p.h_min = case.U_ref, case.kernel.h
        p.acoustic_cfl, p.advection_cfl, p.correction_alpha = 0.6, 0.25, float(correction_alpha)
        p.sigma0, p.wall_rho0, p.contact_depth, p.threads = case.sigma0, 1.0, contact_depth, threads
        p.periodic_axes = int(getattr(case, "periodic_axes", 0))
        if p.periodic_axes:
            for d in range(3):
                p.periodic_lower[d] = case.periodic_lower[d]
                p.periodic_upper[d] = case.periodic_upper[d]
            p.periodic_cutoff = float(self.dtype(case.kernel.cutoff))
        p.surface_indicator = int(surface_indicator)
        p.viscosity, p.transport_velocity, p.transport_coefficient = float(viscosity), int(transport_velocity), 0.2
        self._params = p
        kp, mp = kernel_pod(case.kernel), mesh_pod(case.mesh)
        self._h = lib().orc_create(int(f64), C.byref(p), C.byref(kp), C.byref(mp), C.byref(mp), case.n_fluid, case.n_wall)
        self.n_fluid, self.n_wall = case.n_fluid, case.n_wall
        # initial state (base_material.cpp:37-40, acoustic_step_1st_half.hpp:17-25)
        self.real("Position", 3)[:] = case.fluid_pos.reshape(-1)
        self.real("VolumetricMeasure")[:] = case.vol
        self.real("VolumetricMeasureRef")[:] = case.vol
        self.real("Mass")[:] = case.rho0 * case.vol
        self.real("Density")[:] = case.rho0
        self.real("Compression")[:] = 1.0
        if getattr(case, "fluid_vel", None) is not None:
            self.real("Velocity", 3)[:] = case.fluid_vel.reshape(-1)
        if case.n_wall:
            self.real("Position", 3, wall=True)[:] = case.wall_pos.reshape(-1)
            self.real("VolumetricMeasure", wall=True)[:] = case.vol
            self.real("VolumetricMeasureRef", wall=True)[:] = case.vol
            self.real("Mass", wall=True)[:] = 1.0 * case.vol
            self.real("NormalDirection", 3, wall=True)[:] = case.wall_normal.reshape(-1)

        if observers is not None and len(observers):
            obs = np.asarray(observers, dtype=self.dtype).reshape(-1, 3)
            self.n_observer = obs.shape[0]
            self.exec("set_observers", self.n_observer)
            self.real("Position", 3, body=2)[:] = obs.reshape(-1)
        else:
            self.n_observer = 0

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def real(self, name, width=1, wall=False, body=None) -> np.ndarray:
        """numpy VIEW of a named Real array (re-fetch after ops that may reallocate, e.g. sort). body: 0 fluid, 1 wall, 2 observer."""
        ln = C.c_uint64()
        ptr = lib().orc_real(self._h, int(wall) if body is None else int(body), name.encode(), width, C.byref(ln))
        ct = C.c_double if self.f64 else C.c_float
        if ln.value == 0:
            return np.empty(0, dtype=self.dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(ln.value,))

    def uint(self, name) -> np.ndarray:
        ln = C.c_uint64()
        ptr = lib().orc_uint(self._h, name.encode(), C.byref(ln))
        if ln.value == 0:
            return np.empty(0, dtype=np.uint32)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(ln.value,))

    def exec(self, op, a0=0.0, a1=0.0, a2=0.0, a3=0.0) -> float:
        r = lib().orc_exec(self._h, op.encode(), float(a0), float(a1), float(a2), float(a3))
        if r == -12345.0:
            raise ValueError(f"unknown oracle op {op}")
        return r

    def probe_series(self):
        """rows recorded by prepare_ck / run_ck: interpolated Pressure at every observer, one row per recorded step"""
        rows = int(lib().orc_probe_series(self._h, None, 0))
        out = np.empty((rows, max(self.n_observer, 1)), dtype=np.float64)
        if rows:
            lib().orc_probe_series(self._h, out.ctypes.data, rows)
        return out[:, : self.n_observer]

    def series(self):
        n = lib().orc_series(self._h, None, None, 0)
        t = np.empty(n, dtype=np.float64)
        e = np.empty(n, dtype=np.float64)
        lib().orc_series(self._h, t.ctypes.data, e.ctypes.data, n)
        return t, e
