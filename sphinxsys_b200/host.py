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
             "AverageVelocity", "AverageAcceleration", "ViscousForce", "PreviousViscousForce", "KernelGradientIntegral"}
UINT_NAMES = {"OriginalID", "SortedID", "ReferenceID"}
MAT_NAMES = {"LinearCorrectionMatrix"}
INT_NAMES = {"Indicator", "PreviousSurfaceIndicator"}


class Options(C.Structure):
    _fields_ = [("dim", C.c_int32), ("dp", C.c_double), ("DL", C.c_double), ("DH", C.c_double), ("DW", C.c_double),
                ("LL", C.c_double), ("LH", C.c_double), ("LW", C.c_double), ("correction", C.c_int32),
                ("fused_time_step", C.c_int32), ("fused_regularization", C.c_int32), ("sort_interval", C.c_int32),
                ("device", C.c_int32), ("relation_stride", C.c_int32), ("system_lower", C.c_double * 3),
                ("system_upper", C.c_double * 3), ("use_system_bounds", C.c_int32), ("legacy", C.c_int32), ("rank", C.c_int32),
                ("nranks", C.c_int32), ("unique_id", C.c_uint8 * 128), ("surface_indicator", C.c_int32), ("observers", C.c_int32),
                ("mu_f", C.c_double), ("transport_velocity", C.c_int32), ("serial_exchange", C.c_int32), ("recut_interval", C.c_int32), ("initial_cut_shift", C.c_int32),
                ("riemann", C.c_int32), ("kernel_kind", C.c_int32), ("full_wall", C.c_int32)]


class TaylorGreenOptions(C.Structure):
    _fields_ = [("dim", C.c_int32), ("dp", C.c_double), ("L", C.c_double), ("U_f", C.c_double), ("fused_time_step", C.c_int32),
                ("fused_regularization", C.c_int32), ("sort_interval", C.c_int32), ("device", C.c_int32),
                ("relation_stride", C.c_int32), ("system_lower", C.c_double * 3), ("system_upper", C.c_double * 3),
                ("use_system_bounds", C.c_int32), ("mu_f", C.c_double), ("transport_velocity", C.c_int32), ("x_scale", C.c_double)]


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
        L.sphck_taylor_green_create.restype = C.c_void_p
        L.sphck_taylor_green_create.argtypes = [C.POINTER(TaylorGreenOptions), C.c_void_p, C.c_void_p, C.c_uint64]
        L.sphck_aligned_periodic_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.sphck_taylor_green_create_ring.restype = C.c_void_p
        L.sphck_taylor_green_create_ring.argtypes = [C.POINTER(TaylorGreenOptions), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                                     C.c_int32, C.c_int32, C.c_char_p]
        L.sphck_destroy.argtypes = [C.c_void_p]
        L.sphck_count.restype = C.c_uint64
        L.sphck_count.argtypes = [C.c_void_p, C.c_int]
        L.sphck_record_states.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64)]
        L.sphck_device_allocations.restype = C.c_uint64
        L.sphck_device_allocations.argtypes = [C.c_void_p]
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
        L.sphck_comm_unique_id.argtypes = [C.c_void_p]
        L.sphck_own_range.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.sphck_download_raw.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.sphck_upload_raw.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint64]
        L.sphck_cuts.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.sphck_plan_slab_cuts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.sphck_limit_cut_moves.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sphck_export_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.sphck_probe_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.sphck_pipeline_create.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.sphck_pipeline_stage_uploads.argtypes = [C.c_void_p, C.c_void_p]
        L.sphck_pipeline_commit_uploads.argtypes = [C.c_void_p]
        L.sphck_pipeline_stage_downloads.argtypes = [C.c_void_p, C.c_void_p]
        L.sphck_pipeline_synchronize.argtypes = [C.c_void_p]
        L.sphck_pipeline_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = L
    return _lib


def comm_unique_id() -> bytes:
    """128-byte NCCL communicator id (call on rank 0, hand to every rank: DamBreakCK(..., unique_id=...))."""
    buf = (C.c_uint8 * 128)()
    if load().sphck_comm_unique_id(buf) != 0:
        raise capi.SphB200Error("sphck_comm_unique_id failed (NCCL not loadable?)")
    return bytes(buf)


def plan_slab_cuts(per_plane, nranks):
    """Cell-plane cuts (host-only helper of include/sphinxsys_ck/slab_decomposition.h)."""
    h = np.ascontiguousarray(per_plane, dtype=np.uint64)
    out = np.zeros(nranks + 1, dtype=np.int32)
    if load().sphck_plan_slab_cuts(h.ctypes.data, h.size, int(nranks), out.ctypes.data) != 0:
        raise capi.SphB200Error("plan_slab_cuts failed: " + load().sphck_last_error().decode())
    return out


def aligned_periodic_mesh(lower, upper, cutoff, dim=3):
    """(MeshT, SeamT) of a ring-decomposed periodic body (alignedPeriodicMesh in include/sphinxsys_ck/slab_decomposition.h)."""
    lo = (C.c_double * 3)(*[float(v) for v in lower])
    up = (C.c_double * 3)(*[float(v) for v in upper])
    mesh, seam = capi.MeshT(), capi.SeamT()
    if load().sphck_aligned_periodic_mesh(lo, up, float(cutoff), int(dim), C.byref(mesh), C.byref(seam)) != 0:
        raise capi.SphB200Error("aligned_periodic_mesh failed: " + load().sphck_last_error().decode())
    return mesh, seam


def wall_slab_plan(per_plane, depth, margin, X0, X1, bound, stored=(0, -1)):
    """(lo, hi, reload): the wall planes a rank stores for the fluid planes [X0, X1) given the planes stored now (WallSlab::plan in
    include/sphinxsys_ck/dambreak_case.h; host arithmetic only)."""
    h = np.ascontiguousarray(per_plane, dtype=np.uint64)
    below = np.concatenate([[0], np.cumsum(h)]).astype(np.uint64)
    lo_hi = np.array(stored, dtype=np.int32)
    reload = C.c_int32(0)
    fn = load().sphck_wall_slab_plan
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.POINTER(C.c_int32)]
    if fn(below.ctypes.data, int(h.size), int(depth), int(margin), int(X0), int(X1), int(bound), lo_hi.ctypes.data, C.byref(reload)) != 0:
        raise capi.SphB200Error("wall_slab_plan failed: " + load().sphck_last_error().decode())
    return int(lo_hi[0]), int(lo_hi[1]), bool(reload.value)


def limit_cut_moves(old_cuts, wanted):
    """Cuts a re-balancing step may take from `old_cuts` towards `wanted` (SlabDecomposition::recut: neighbour transfers only)."""
    o = np.ascontiguousarray(old_cuts, dtype=np.int32)
    w = np.ascontiguousarray(wanted, dtype=np.int32)
    out = np.zeros(o.size, dtype=np.int32)
    if load().sphck_limit_cut_moves(o.ctypes.data, w.ctypes.data, int(o.size - 1), out.ctypes.data) != 0:
        raise capi.SphB200Error("limit_cut_moves failed: " + load().sphck_last_error().decode())
    return out


def particle_digest(ids, pos, vel):
    """64-bit digest of a particle set that does not depend on storage order or on how the set is split over ranks:
    the sum (mod 2^64) over the particles of a mix of the particle's global number with the bit patterns of its position and
    velocity. Equal digests of two runs <=> (up to hash collisions) every particle has bit-identical position and velocity."""
    h = np.ascontiguousarray(ids, dtype=np.uint32).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    for a in (pos, vel):
        bits = np.ascontiguousarray(a[:, :3], dtype=np.float32).view(np.uint32)
        for c in range(3):
            h ^= bits[:, c].astype(np.uint64)
            h *= np.uint64(0xBF58476D1CE4E5B9)
            h ^= h >> np.uint64(29)
    return int(np.add.reduce(h, dtype=np.uint64)) if h.size else 0


def _kind(name):
    return 1 if name in VEC_NAMES else (2 if name in UINT_NAMES else (3 if name in MAT_NAMES else (4 if name in INT_NAMES else 0)))


class DamBreakCK:
    """Handle of a C++ `SPH::DamBreakCK` (include/sphinxsys_ck/dambreak_case.h)."""

    def __init__(self, case=None, device_index=0, correction=False, fused_time_step=True, sort_interval=100,
                 relation_stride=None, fused_regularization=True, dim=3, dp=0.05, generate=False, rank=0, nranks=1,
                 unique_id=None, width_scale=1.0, legacy=False, surface_indicator=False, observers=False, mu_f=0.0,
                 transport_velocity=False, serial_exchange=False, recut_interval=None, initial_cut_shift=0, riemann=1,
                 kernel_kind=None, full_wall=False):
        self.lib = load()
        o = Options()
        o.full_wall = int(bool(full_wall))
        o.riemann = int(riemann)
        # the smoothing kernel follows the case the oracle gets unless stated (0 Wendland C2, 1 Laguerre-Gauss)
        o.kernel_kind = int(kernel_kind if kernel_kind is not None else (getattr(case.kernel, "kind", 0) if case is not None else 0))
        o.serial_exchange = int(bool(serial_exchange))
        o.recut_interval = -1 if recut_interval is None else int(recut_interval)
        o.initial_cut_shift = int(initial_cut_shift)
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
        o.DW, o.LW = o.DW * width_scale, o.LW * width_scale
        o.legacy = int(bool(legacy))
        o.surface_indicator, o.observers = int(bool(surface_indicator)), int(bool(observers))
        o.mu_f, o.transport_velocity = float(mu_f), int(bool(transport_velocity))
        o.rank, o.nranks = int(rank), int(nranks)
        if nranks > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("decomposed runs need the 128-byte unique_id from comm_unique_id() on rank 0")
            for i, b in enumerate(unique_id):
                o.unique_id[i] = b
        self.rank, self.nranks = int(rank), int(nranks)
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
        self.n_fluid = int(self.lib.sphck_count(self._h, 0))  # stored fluid particles (own + ghosts when decomposed)
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
    def device_allocations(self) -> int:
        """cudaMalloc calls made through the library so far (arrays + scratch arenas)."""
        return int(self.lib.sphck_device_allocations(self._h))

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
        shape, dt = {0: ((n,), np.float32), 1: ((n, 3), np.float32), 2: ((n,), np.uint32), 3: ((n, 9), np.float32),
                     4: ((n,), np.int32)}[k]
        if out is None:
            out = np.empty(shape, dtype=dt)
        self._check(self.lib.sphck_download(self._h, int(wall), name.encode(), k, out.ctypes.data), f"download {name}")
        return out

    def upload(self, name, arr, wall=False):
        k = _kind(name)
        dt = np.uint32 if k == 2 else (np.int32 if k == 4 else np.float32)
        a = np.ascontiguousarray(arr, dtype=dt)
        self._check(self.lib.sphck_upload(self._h, int(wall), name.encode(), k, a.ctypes.data), f"upload {name}")

    def own_range(self):
        """(begin, count, stored): slots of this rank's own particles and the stored total (own + ghosts)."""
        b, c, s_ = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.sphck_own_range(self._h, C.byref(b), C.byref(c), C.byref(s_)), "own_range")
        return int(b.value), int(c.value), int(s_.value)

    def download_own(self, name) -> np.ndarray:
        """This rank's own particles of a variable in STORAGE order (pair with download_own('ReferenceID'))."""
        b, c, _ = self.own_range()
        k = _kind(name)
        shape, dt = {0: ((c,), np.float32), 1: ((c, 4), np.float32), 2: ((c,), np.uint32), 3: ((c, 9), np.float32)}[k]
        out = np.empty(shape, dtype=dt)
        self._check(self.lib.sphck_download_raw(self._h, 0, name.encode(), out.ctypes.data, b, c), f"download_raw {name}")
        return out[:, :3] if k == 1 else out

    def record_states(self, folder) -> int:
        """BodyStatesRecordingToVtpCK::writeToFile (device -> host sync of the write list + one .vtp per body); returns the
        bytes synchronised so far."""
        b = C.c_uint64(0)
        self._check(self.lib.sphck_record_states(self._h, str(folder).encode(), C.byref(b)), "record_states")
        return int(b.value)

    def step_trace_report(self, steps):
        """SPHB200_STEP_TRACE=1: print the per-stage wall-time table of the case loop to stderr and reset it."""
        self.lib.sphck_step_trace_report.argtypes = [C.c_uint64]
        self._check(self.lib.sphck_step_trace_report(int(steps)), "step_trace_report")

    def state_digest(self):
        """(own particle count, particle_digest of (ReferenceID, Position, Velocity)) of this rank's own particles; the sum of
        the digests over the ranks (mod 2^64) equals the digest of the same state held by one GPU."""
        ids = self.download_own("ReferenceID")
        with np.errstate(over="ignore"):
            return int(ids.size), particle_digest(ids, self.download_own("Position"), self.download_own("Velocity"))

    def download_own_into(self, name, out):
        """download_own into a caller buffer (e.g. pinned): device element layout, Vecd = 4 floats per particle."""
        b, c, _ = self.own_range()
        self._check(self.lib.sphck_download_raw(self._h, 0, name.encode(), out.ctypes.data, b, c), f"download_raw {name}")

    def upload_own(self, name, arr):
        """Raw upload of this rank's own particles (storage order, device element layout). Asynchronous for pinned memory."""
        b, c, _ = self.own_range()
        self._check(self.lib.sphck_upload_raw(self._h, 0, name.encode(), arr.ctypes.data, b, c), f"upload_raw {name}")

    def probe_records(self):
        """(times[rows], pressure[rows, probes]) recorded by the case's ObservedQuantityRecording so far."""
        rows, probes = int(self.exec("probe_records")), int(self.exec("probe_count"))
        t = np.zeros(rows, dtype=np.float64)
        v = np.zeros((rows, max(probes, 1)), dtype=np.float64)
        if rows:
            self._check(self.lib.sphck_probe_records(self._h, t.ctypes.data, v.ctypes.data, rows), "probe_records")
        return t, v[:, :probes]

    # ---- overlapped host <-> device transfers (HostTransferPipeline; host arrays must be pinned, reference order) ----
    def pipeline_create(self, inputs, outputs):
        self._check(self.lib.sphck_pipeline_create(self._h, ",".join(inputs).encode(), ",".join(outputs).encode()), "pipeline_create")
        self._pipe_in, self._pipe_out = list(inputs), list(outputs)

    def pipeline_stage_uploads(self, arrays):
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        self._check(self.lib.sphck_pipeline_stage_uploads(self._h, ptrs), "pipeline_stage_uploads")

    def pipeline_commit_uploads(self):
        self._check(self.lib.sphck_pipeline_commit_uploads(self._h), "pipeline_commit_uploads")

    def pipeline_stage_downloads(self, arrays):
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        self._check(self.lib.sphck_pipeline_stage_downloads(self._h, ptrs), "pipeline_stage_downloads")

    def pipeline_synchronize(self):
        self._check(self.lib.sphck_pipeline_synchronize(self._h), "pipeline_synchronize")

    def pipeline_bytes(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.sphck_pipeline_bytes(self._h, C.byref(a), C.byref(b)), "pipeline_bytes")
        return int(a.value), int(b.value)

    def cuts(self) -> np.ndarray:
        out = np.zeros(self.nranks + 1, dtype=np.int32)
        self._check(self.lib.sphck_cuts(self._h, out.ctypes.data, out.size), "cuts")
        return out

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


class TaylorGreenCK(DamBreakCK):
    """Handle of a C++ `SPH::TaylorGreenCK` (include/sphinxsys_ck/taylor_green_case.h): periodic box, no wall body.
    Shares the driving interface of DamBreakCK (exec by name, upload/download in the reference particle order)."""

    def __init__(self, case=None, device_index=0, fused_time_step=True, sort_interval=100, relation_stride=None,
                 fused_regularization=True, dim=3, n_side=32, generate=False, mu_f=0.0, transport_velocity=False,
                 ring=False, rank=0, nranks=1, unique_id=None, own=None, local_ids=None):
        """ring=True: periodic along x through a ring of slabs, one process per GPU (nranks == 1: a ring of one slab).
        `own` (indices into the case's particle arrays, ascending) are this rank's particles; default all of them.
        `local_ids`: the case holds THIS rank's particles only and these are their global numbers (large runs)."""
        self.lib = load()
        o = TaylorGreenOptions()
        o.mu_f, o.transport_velocity = float(mu_f), int(bool(transport_velocity))
        o.x_scale = 1.0
        if case is not None:
            o.dim, o.dp, o.L, o.U_f = case.dim, case.dp, case.DH, case.U_ref
            o.x_scale = case.DL / case.DH
        else:
            o.dim, o.dp, o.L, o.U_f = dim, 1.0 / n_side, 1.0, 1.0
        o.fused_time_step, o.fused_regularization = int(fused_time_step), int(fused_regularization)
        o.sort_interval, o.device = int(sort_interval), int(device_index)
        o.relation_stride = -1 if relation_stride is None else int(relation_stride)
        o.use_system_bounds = 0
        self.rank, self.nranks, self.case = int(rank), int(nranks), case
        if ring:
            if case is None:
                raise ValueError("a ring run needs the case (this rank's particles are cut out of it)")
            if local_ids is not None:
                fp = np.ascontiguousarray(case.fluid_pos, dtype=np.float32)
                fv = np.ascontiguousarray(case.fluid_vel, dtype=np.float32)
                ids = np.ascontiguousarray(local_ids, dtype=np.uint32)
                if ids.size != fp.shape[0]:
                    raise ValueError("local_ids: one global number per particle of the case")
            else:
                own = np.arange(case.n_fluid) if own is None else np.asarray(own)
                fp = np.ascontiguousarray(case.fluid_pos[own], dtype=np.float32)
                fv = np.ascontiguousarray(case.fluid_vel[own], dtype=np.float32)
                ids = np.ascontiguousarray(own, dtype=np.uint32)
            uid = bytes(unique_id) if unique_id is not None else bytes(128)
            self._h = self.lib.sphck_taylor_green_create_ring(C.byref(o), fp.ctypes.data, fv.ctypes.data, ids.ctypes.data,
                                                              fp.shape[0], int(rank), int(nranks), uid)
        elif case is not None and not generate:
            if case.system_lower is not None:
                o.use_system_bounds = 1
                for d in range(3):
                    o.system_lower[d] = case.system_lower[d]
                    o.system_upper[d] = case.system_upper[d]
            fp = np.ascontiguousarray(case.fluid_pos, dtype=np.float32)
            fv = np.ascontiguousarray(case.fluid_vel, dtype=np.float32)
            self._h = self.lib.sphck_taylor_green_create(C.byref(o), fp.ctypes.data, fv.ctypes.data, fp.shape[0])
        else:
            self._h = self.lib.sphck_taylor_green_create(C.byref(o), None, None, 0)
        if not self._h:
            raise capi.SphB200Error("sphck_taylor_green_create failed: " + self.lib.sphck_last_error().decode())
        self.n_fluid = int(self.lib.sphck_count(self._h, 0))
        self.n_wall = 0

    @property
    def ghost_particles(self) -> int:
        return int(self.exec("ghost_particles"))
