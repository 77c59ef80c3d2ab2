"""ctypes binding of libsphb200_host.so — the C++ host layer (include/sphinxsys_ck/*.h) driven from Python.

The host code of this project is C++ (as the reference's is): bodies, relations, the dynamics classes with the
reference spellings and the dam-break case loop live in include/sphinxsys_ck/. This module only lets pytest and
bench.py create a case object, run its dynamics by name and move arrays in and out in the REFERENCE particle order.
There is no fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libsphb200_host.so")

VEC_NAMES = {"Position", "Velocity", "Displacement", "Force", "ForcePrior", "NormalDirection", "PreviousGravityForceCK",
             "AverageVelocity", "AverageAcceleration"}
UINT_NAMES = {"OriginalID", "SortedID", "ReferenceID"}
MAT_NAMES = {"LinearCorrectionMatrix"}


class Options(C.Structure):
    _fields_ = [("dim", C.c_int32), ("dp", C.c_double), ("DL", C.c_double), ("DH", C.c_double), ("DW", C.c_double),
                ("LL", C.c_double), ("LH", C.c_double), ("LW", C.c_double), ("correction", C.c_int32),
                ("fused_time_step", C.c_int32), ("fused_regularization", C.c_int32), ("sort_interval", C.c_int32),
                ("device", C.c_int32), ("relation_stride", C.c_int32), ("system_lower", C.c_double * 3),
                ("system_upper", C.c_double * 3), ("use_system_bounds", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        capi.load()  # libsphb200.so first (the host library links it)
        if not os.path.exists(HOST_LIB_PATH):
            raise capi.SphB200Error(f"{HOST_LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(HOST_LIB_PATH)
        L.sphck_last_error.restype = C.c_char_p
        L.sphck_dambreak_create.restype = C.c_void_p
        L.sphck_dambreak_create.argtypes = [C.POINTER(Options), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64]
        L.sphck_destroy.argtypes = [C.c_void_p]
        L.sphck_count.restype = C.c_uint64
        L.sphck_count.argtypes = [C.c_void_p, C.c_int]
        L.sphck_launches.restype = C.c_uint64
        L.sphck_launches.argtypes = [C.c_void_p]
        L.sphck_synchronize.argtypes = [C.c_void_p]
        L.sphck_mesh.argtypes = [C.c_void_p, C.c_int, C.POINTER(capi.MeshT)]
        L.sphck_kernel.argtypes = [C.c_void_p, C.POINTER(capi.KernelT)]
        L.sphck_exec.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.POINTER(C.c_double)]
        L.sphck_acoustic1_phase.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.sphck_download.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        L.sphck_upload.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
        L.sphck_has_variable.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.sphck_device_pointer.restype = C.c_void_p
        L.sphck_device_pointer.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.sphck_cell_offsets.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64]
        L.sphck_export_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def _kind(name):
    return 1 if name in VEC_NAMES else (2 if name in UINT_NAMES else (3 if name in MAT_NAMES else 0))


class DamBreakCK:
    """Handle of a C++ `SPH::DamBreakCK` (include/sphinxsys_ck/dambreak_case.h)."""

    def __init__(self, case=None, device_index=0, correction=False, fused_time_step=True, sort_interval=100,
                 relation_stride=None, fused_regularization=True, dim=3, dp=0.05, generate=False):
        self.lib = load()
        o = Options()
        if case is not None:
            dim, dp = case.dim, case.dp
            o.DL, o.DH, o.DW, o.LL, o.LH, o.LW = case.DL, case.DH, case.DW, case.LL, case.LH, case.LW
        else:
            o.DL, o.DH, o.DW, o.LL, o.LH, o.LW = (5.366, 2.0, 0.5, 2.0, 1.0, 0.5) if dim == 3 else (5.366, 5.366, 0.0, 2.0, 1.0, 0.0)
        o.dim, o.dp = dim, dp
        o.correction, o.fused_time_step, o.fused_regularization = int(correction), int(fused_time_step), int(fused_regularization)
        o.sort_interval, o.device = int(sort_interval), int(device_index)
        o.relation_stride = -1 if relation_stride is None else int(relation_stride)
        o.use_system_bounds = 0
        self.case = case
        if case is not None and not generate:
            fp = np.ascontiguousarray(case.fluid_pos, dtype=np.float32)
            wp = np.ascontiguousarray(case.wall_pos, dtype=np.float32)
            wn = np.ascontiguousarray(case.wall_normal, dtype=np.float32)
            self._h = self.lib.sphck_dambreak_create(C.byref(o), fp.ctypes.data, fp.shape[0], wp.ctypes.data, wn.ctypes.data, wp.shape[0])
        else:
            self._h = self.lib.sphck_dambreak_create(C.byref(o), None, 0, None, None, 0)
        if not self._h:
            raise capi.SphB200Error("sphck_dambreak_create failed: " + self.lib.sphck_last_error().decode())
        self.n_fluid = int(self.lib.sphck_count(self._h, 0))
        self.n_wall = int(self.lib.sphck_count(self._h, 1))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sphck_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise capi.SphB200Error(f"{what} failed: {self.lib.sphck_last_error().decode()}")

    def exec(self, op, a0=0.0) -> float:
        r = C.c_double(0)
        self._check(self.lib.sphck_exec(self._h, op.encode(), float(a0), C.byref(r)), op)
        return r.value

    def initialize(self):
        self.exec("initialize")

    def step_outer(self) -> int:
        return int(self.exec("step_outer"))

    def run_outer(self, n) -> int:
        return int(self.exec("run_outer", n))

    def acoustic1_phase(self, phase, dt):
        self._check(self.lib.sphck_acoustic1_phase(self._h, int(phase), float(dt)), "acoustic1_phase")

    def energy(self) -> float:
        return self.exec("energy")

    def synchronize(self):
        self._check(self.lib.sphck_synchronize(self._h), "synchronize")

    @property
    def launches(self) -> int:
        return int(self.lib.sphck_launches(self._h))

    @property
    def physical_time(self):
        return self.exec("physical_time")

    @property
    def acoustic_steps(self):
        return int(self.exec("acoustic_steps"))

    @property
    def last_acoustic_dt(self):
        return self.exec("last_acoustic_dt")

    def has_variable(self, name, wall=False):
        return bool(self.lib.sphck_has_variable(self._h, int(wall), name.encode()))

    def download(self, name, wall=False, out=None) -> np.ndarray:
        """Host copy of a variable in the reference particle order (Vecd packed as 3 floats)."""
        n = self.n_wall if wall else self.n_fluid
        k = _kind(name)
        shape, dt = {0: ((n,), np.float32), 1: ((n, 3), np.float32), 2: ((n,), np.uint32), 3: ((n, 9), np.float32)}[k]
        if out is None:
            out = np.empty(shape, dtype=dt)
        self._check(self.lib.sphck_download(self._h, int(wall), name.encode(), k, out.ctypes.data), f"download {name}")
        return out

    def upload(self, name, arr, wall=False):
        k = _kind(name)
        dt = np.uint32 if k == 2 else np.float32
        a = np.ascontiguousarray(arr, dtype=dt)
        self._check(self.lib.sphck_upload(self._h, int(wall), name.encode(), k, a.ctypes.data), f"upload {name}")

    def mesh(self, wall=False) -> capi.MeshT:
        m = capi.MeshT()
        self._check(self.lib.sphck_mesh(self._h, int(wall), C.byref(m)), "mesh")
        return m

    def kernel(self) -> capi.KernelT:
        k = capi.KernelT()
        self._check(self.lib.sphck_kernel(self._h, C.byref(k)), "kernel")
        return k

    def cell_offsets(self, wall=False) -> np.ndarray:
        m = self.mesh(wall)
        cells = int(m.cells[0]) * int(m.cells[1]) * int(m.cells[2])
        out = np.empty(cells + 1, dtype=np.uint32)
        self._check(self.lib.sphck_cell_offsets(self._h, int(wall), out.ctypes.data, cells + 1), "cell_offsets")
        return out

    def export_csr(self, contact=False):
        """(particle_offset[n+1], neighbor_index[total]) of the inner / contact relation in reference particle ids."""
        total = C.c_uint64(0)
        off = np.empty(self.n_fluid + 1, dtype=np.uint32)
        self._check(self.lib.sphck_export_csr(self._h, int(contact), off.ctypes.data, None, 0, C.byref(total)), "export_csr")
        idx = np.empty(max(int(total.value), 1), dtype=np.uint32)
        self._check(self.lib.sphck_export_csr(self._h, int(contact), off.ctypes.data, idx.ctypes.data, idx.size, C.byref(total)), "export_csr")
        return off, idx[: int(total.value)]
