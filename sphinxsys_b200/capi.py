"""ctypes binding of libsphb200.so — the C ABI declared in include/sphb200.h.

There is NO fallback: if the shared library is missing or a call fails this module raises. PyTorch is only
used by callers for device memory and streams; every pointer crossing this boundary is a raw device pointer.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsphb200.so")


class SphB200Error(RuntimeError):
    pass


class Vec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class MeshT(C.Structure):
    _fields_ = [("lower", C.c_float * 3), ("spacing", C.c_float), ("cells", C.c_int32 * 3)]


class SeamT(C.Structure):
    """sphb200_seam_t: the periodic seam of a ring of slabs (x axis of the mesh, box planes, plane x ranges)."""
    _fields_ = [("mesh_lower", C.c_float), ("mesh_spacing", C.c_float), ("mesh_cells", C.c_int32), ("first_plane", C.c_int32),
                ("box_planes", C.c_int32), ("own_min", C.c_float), ("own_max", C.c_float), ("ghost_low_min", C.c_float),
                ("ghost_low_max", C.c_float), ("ghost_high_min", C.c_float), ("ghost_high_max", C.c_float)]


class KernelT(C.Structure):
    _fields_ = [("dim", C.c_int32), ("kind", C.c_int32), ("h", C.c_float), ("src_h", C.c_float),
                ("kernel_size", C.c_float), ("dimension_factor", C.c_float), ("w", C.c_float * 24), ("dw", C.c_float * 24)]


class FluidT(C.Structure):
    _fields_ = [("rho0", C.c_float), ("c0", C.c_float), ("riemann", C.c_int32), ("correction", C.c_int32),
                ("limiter_coeff", C.c_float), ("free_surface", C.c_int32), ("formulation", C.c_int32),
                ("sigma0", C.c_float), ("wall_rho0", C.c_float)]


_P = C.c_void_p


class FluidView(C.Structure):
    _fields_ = [("n", C.c_uint32), ("pos", _P), ("vel", _P), ("dpos", _P), ("force", _P), ("force_prior", _P),
                ("vol", _P), ("mass", _P), ("rho", _P), ("p", _P), ("compression", _P), ("compression_rate", _P),
                ("vol_ref", _P), ("compression_sum", _P), ("B", _P), ("correction_record", _P), ("posvol", _P),
                ("posvolref", _P), ("posvolvel", _P), ("active_begin", C.c_uint32), ("active_end", C.c_uint32)]


class WallView(C.Structure):
    _fields_ = [("n", C.c_uint32), ("pos", _P), ("posvol", _P), ("posvolref", _P), ("vel_ave", _P), ("acc_ave", _P), ("normal", _P),
                ("vol_ref", _P)]


class CellListT(C.Structure):
    _fields_ = [("cell_offset", _P), ("particle_index", _P), ("sorted_pos", _P)]


class RelationT(C.Structure):
    _fields_ = [("count", _P), ("slice_offset", _P), ("index", _P), ("capacity", C.c_uint64), ("order", _P),
                ("bank_aligned", C.c_int32)]


class SearchT(C.Structure):
    _fields_ = [("tar_mesh", MeshT), ("kernel", KernelT), ("src_pos", _P), ("n_src", C.c_uint32), ("src_order", _P),
                ("src_sorted_pos", _P), ("tar_pos", _P), ("tar_list", CellListT), ("is_inner", C.c_int32), ("legacy_criterion", C.c_int32),
                ("search_depth", C.c_int32), ("src_begin", C.c_uint32), ("src_end", C.c_uint32), ("cell_ordered", C.c_int32),
                ("tar2_pos", _P), ("tar2_list", CellListT), ("tar2_index_base", C.c_uint32)]


class PeriodicT(C.Structure):
    _fields_ = [("lower", C.c_float * 3), ("upper", C.c_float * 3), ("axes", C.c_int32), ("cutoff", C.c_float)]


class FluidArgs(C.Structure):
    _fields_ = [("fluid", FluidView), ("wall", WallView), ("inner", RelationT), ("contact", RelationT),
                ("kernel", KernelT), ("material", FluidT)]


# every symbol include/sphb200.h declares: name -> (restype, argtypes)
_CTX = C.c_void_p
_I, _U32, _U64, _F = C.c_int, C.c_uint32, C.c_uint64, C.c_float
SYMBOLS = {
    "sphb200_version": (_I, []),
    "sphb200_context_create": (_I, [_I, C.POINTER(_CTX)]),
    "sphb200_context_destroy": (_I, [_CTX]),
    "sphb200_last_error_string": (C.c_char_p, [_CTX]),
    "sphb200_launch_count": (_U64, [_CTX]),
    "sphb200_malloc_device": (_I, [C.POINTER(_P), C.c_size_t]),
    "sphb200_malloc_host": (_I, [C.POINTER(_P), C.c_size_t]),
    "sphb200_free_device": (_I, [_P]),
    "sphb200_free_host": (_I, [_P]),
    "sphb200_copy_h2d": (_I, [_P, _P, C.c_size_t, _P]),
    "sphb200_copy_d2h": (_I, [_P, _P, C.c_size_t, _P]),
    "sphb200_copy_d2d": (_I, [_P, _P, C.c_size_t, _P]),
    "sphb200_stream_sync": (_I, [_P]),
    "sphb200_fill_u32": (_I, [_CTX, _P, _U32, _U64, _P]),
    "sphb200_fill_f32": (_I, [_CTX, _P, _F, _U64, _P]),
    "sphb200_iota_u32": (_I, [_CTX, _P, _U64, _P]),
    "sphb200_vec3_to_vec4": (_I, [_CTX, _P, _P, _U32, _P]),
    "sphb200_vec4_to_vec3": (_I, [_CTX, _P, _P, _U32, _P]),
    "sphb200_pack_posvol": (_I, [_CTX, _P, _P, _P, _U32, _P]),
    "sphb200_pack_records": (_I, [_CTX, _U32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sphb200_exclusive_scan_u32": (_I, [_CTX, _P, _P, _U64, C.POINTER(_U32), _P]),
    "sphb200_sort_pairs_u32": (_I, [_CTX, _P, _P, _U64, _I, _P]),
    "sphb200_gather_multi": (_I, [_CTX, _I, C.POINTER(_P), C.POINTER(_P), C.POINTER(_U32), _P, _U32, _P]),
    "sphb200_morton_keys": (_I, [_CTX, C.POINTER(MeshT), _P, _U32, _P, _P, _P, _P]),
    "sphb200_update_sorted_id": (_I, [_CTX, _P, _P, _U32, _P]),
    "sphb200_cell_list_build": (_I, [_CTX, C.POINTER(MeshT), _P, _U32, CellListT, _P]),
    "sphb200_cell_list_build_reorder": (_I, [_CTX, C.POINTER(MeshT), _P, _U32, _P, CellListT, _I, C.POINTER(_P), C.POINTER(_P),
                                             C.POINTER(_U32), _P]),
    "sphb200_relation_count": (_I, [_CTX, C.POINTER(SearchT), RelationT, C.POINTER(_U64), _P]),
    "sphb200_relation_fill": (_I, [_CTX, C.POINTER(SearchT), RelationT, _P]),
    "sphb200_relation_build_fixed": (_I, [_CTX, C.POINTER(SearchT), RelationT, _U32, C.POINTER(_U32), _P]),
    "sphb200_periodic_bounding": (_I, [_CTX, C.POINTER(PeriodicT), _P, _U32, _P]),
    "sphb200_periodic_images": (_I, [_CTX, C.POINTER(PeriodicT), _P, _U32, _P, _P, _U32, C.POINTER(_U32), _P]),
    "sphb200_ghost_copy": (_I, [_CTX, _P, _U32, _U32, _U32, _P, _U32, _U32, _P]),
    "sphb200_relation_export_csr": (_I, [_CTX, RelationT, _U32, _P, _P, _P, _P, _U64, _P]),
    "sphb200_gravity_force": (_I, [_CTX, C.POINTER(FluidView), C.POINTER(_F * 3), _P, _P]),
    "sphb200_compression_summation": (_I, [_CTX, C.POINTER(FluidArgs), _I, _P]),
    "sphb200_density_regularization": (_I, [_CTX, C.POINTER(FluidArgs), _P]),
    "sphb200_advection_setup": (_I, [_CTX, C.POINTER(FluidView), _P]),
    "sphb200_update_position": (_I, [_CTX, C.POINTER(FluidView), _P]),
    "sphb200_advection_time_step": (_I, [_CTX, C.POINTER(FluidView), _F, _F, _F, C.POINTER(_F), C.POINTER(_F), _P]),
    "sphb200_acoustic_time_step": (_I, [_CTX, C.POINTER(FluidArgs), _F, _F, C.POINTER(_F), C.POINTER(_F), _P]),
    "sphb200_advection_time_step_legacy": (_I, [_CTX, C.POINTER(FluidView), _F, _F, _F, C.POINTER(_F), C.POINTER(_F), _P]),
    "sphb200_acoustic_1st_half": (_I, [_CTX, C.POINTER(FluidArgs), _F, _P]),
    "sphb200_acoustic_2nd_half": (_I, [_CTX, C.POINTER(FluidArgs), _F, _F, _P, _P]),
    "sphb200_acoustic_1st_half_initialize": (_I, [_CTX, C.POINTER(FluidArgs), _F, _P]),
    "sphb200_acoustic_1st_half_interact": (_I, [_CTX, C.POINTER(FluidArgs), _F, _I, _P]),
    "sphb200_linear_correction_matrix": (_I, [_CTX, C.POINTER(FluidArgs), _F, _P]),
    "sphb200_pack_correction_records": (_I, [_CTX, _U32, _P, _P, _P, _P]),
    "sphb200_stream_create": (_I, [C.POINTER(_P)]),
    "sphb200_slab_select": (_I, [_CTX, C.POINTER(MeshT), _P, _U32, _U32, _I, _I, _P, _P, _P, _P]),
    "sphb200_stream_create_with_priority": (_I, [C.POINTER(_P), _I]),
    "sphb200_stream_destroy": (_I, [_P]),
    "sphb200_event_create": (_I, [C.POINTER(_P)]),
    "sphb200_event_destroy": (_I, [_P]),
    "sphb200_event_record": (_I, [_P, _P]),
    "sphb200_stream_wait_event": (_I, [_P, _P]),
    "sphb200_viscous_force": (_I, [_CTX, C.POINTER(FluidArgs), _F, _F, _P, _P, _P]),
    "sphb200_kernel_gradient_integral": (_I, [_CTX, C.POINTER(FluidArgs), _P, _P]),
    "sphb200_transport_velocity_correction": (_I, [_CTX, C.POINTER(FluidView), _P, _F, _F, _I, _F, _P, _P]),
    "sphb200_free_surface_indication": (_I, [_CTX, C.POINTER(FluidArgs), _P, _P, _P, _F, _F, _P]),
    "sphb200_free_surface_indication_sweep": (_I, [_CTX, C.POINTER(FluidArgs), _P, _P, _P, _F, _F, _I, _P]),
    "sphb200_interpolate": (_I, [_CTX, C.POINTER(KernelT), _P, C.c_uint32, RelationT, _P, _P, _I, _P, _P]),
    "sphb200_interpolate_restoring": (_I, [_CTX, C.POINTER(KernelT), _P, C.c_uint32, RelationT, _P, _P, _I, _P, _P]),
    "sphb200_comm_unique_id": (_I, [_P]),
    "sphb200_comm_create": (_I, [_CTX, _I, _I, _P]),
    "sphb200_comm_destroy": (_I, [_CTX]),
    "sphb200_comm_rank": (_I, [_CTX]),
    "sphb200_comm_size": (_I, [_CTX]),
    "sphb200_comm_exchange": (_I, [_CTX, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "sphb200_comm_create_self": (_I, [_CTX]),
    "sphb200_comm_set_ring": (_I, [_CTX, _I]),
    "sphb200_comm_is_ring": (_I, [_CTX]),
    "sphb200_seam_shift": (_I, [_CTX, _P, _U32, _U32, C.c_float, _P, _P]),
    "sphb200_device_allocation_count": (C.c_uint64, []),
    "sphb200_comm_mailbox_open": (_I, [_CTX, C.c_size_t]),
    "sphb200_comm_mailbox_close": (_I, [_CTX]),
    "sphb200_comm_mailbox_peer_bytes": (C.c_size_t, [_CTX, _I]),
    "sphb200_comm_mailbox_status": (_P, [_CTX]),
    "sphb200_comm_push": (_I, [_CTX, _I, _I, _P, _P, _P, _P, C.c_uint64, _P]),
    "sphb200_comm_pull": (_I, [_CTX, _I, _I, _P, _P, _U32, _P, _U32, _P, C.c_uint64, _P]),
    "sphb200_cell_list_build_reorder_n": (_I, [_CTX, C.POINTER(MeshT), _P, _U32, _P, _P, CellListT, _I, _P, _P, _P, _P]),
    "sphb200_slab_total": (_I, [_CTX, _U32, _P, _P, _P, _P]),
    "sphb200_slab_bounds": (_I, [_CTX, _P, _P, _I, _P, _I, _P, _P]),
    "sphb200_comm_allreduce_max_f32": (_I, [_CTX, _P, _I, _P]),
    "sphb200_comm_allreduce_sum_f64": (_I, [_CTX, _P, _I, _P]),
    "sphb200_comm_allgather_u64": (_I, [_CTX, _P, _P, _I, _P]),
    "sphb200_total_mechanical_energy": (_I, [_CTX, C.POINTER(FluidView), C.POINTER(_F * 3), C.POINTER(C.c_double), _P]),
}

_lib = None


def load():
    """Load libsphb200.so and bind every declared symbol. Raises if the library or a symbol is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphB200Error(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def mesh_t(mesh) -> MeshT:
    m = MeshT()
    for d in range(3):
        m.lower[d] = mesh.lower[d]
        m.cells[d] = mesh.cells[d]
    m.spacing = mesh.spacing
    return m


def kernel_t(k, src_h=None) -> KernelT:
    t = KernelT()
    t.dim, t.kind, t.h, t.kernel_size, t.dimension_factor = k.dim, k.kind, k.h, k.kernel_size, k.dimension_factor
    t.src_h = k.h if src_h is None else src_h
    for i in range(24):
        t.w[i] = float(k.w[i])
        t.dw[i] = float(k.dw[i])
    return t


class Context:
    """sphb200_context_t wrapper: one per GPU; turns non-zero return codes into exceptions."""

    def __init__(self, device=0):
        self.lib = load()
        self._ctx = _CTX()
        rc = self.lib.sphb200_context_create(int(device), C.byref(self._ctx))
        if rc != 0:
            raise SphB200Error(f"sphb200_context_create(device={device}) failed with code {rc}")
        self.device = device

    def close(self):
        if self._ctx:
            self.lib.sphb200_context_destroy(self._ctx)
            self._ctx = _CTX()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.sphb200_launch_count(self._ctx))

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self._ctx, *args)
        if rc != 0:
            msg = self.lib.sphb200_last_error_string(self._ctx)
            raise SphB200Error(f"{name} failed: code {rc}: {msg.decode() if msg else ''}")
        return rc
