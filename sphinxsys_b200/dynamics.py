"""Host-side mirror of the SPHinXsys CK dynamics interface for the WCSPH hot path, driving libsphb200.so.

Names, argument meaning and sequencing follow the reference (paths relative to /root/reference/src/shared):
  execution policies ........ particle_dynamics/execution/execution_policy.h:37-83          (par_device)
  DiscreteVariable .......... common/sphinxsys_variable.h:196-378 (host array + device mirror, explicit sync)
  BaseParticles ............. particles/base_particles.h:80-263 (name -> variable registry, evolving variables)
  Inner / Contact ........... shared_ck/body_relation/relation_ck.h:59-174
  UpdateCellLinkedList ...... shared_ck/particle_dynamics/configuration_dynamics/update_cell_linked_list.hpp:75-106
  UpdateRelation ............ .../update_body_relation.hpp:117-164,240-288 (count -> scan -> grow -> fill)
  ParticleSortCK ............ .../particle_sort_ck.hpp:76-104
  StateDynamics / ReduceDynamicsCK ... shared_ck/particle_dynamics/simple_algorithms_ck.h:41-121
  InteractionDynamicsCK ..... shared_ck/particle_dynamics/interaction_algorithms_ck.{h,hpp,cpp}
  fluid dynamics ............ shared_ck/particle_dynamics/fluid_dynamics/*.h (see capi / sphb200.h per entry point)

The reference is compiled C++; the C++ twin of this mirror is include/sphinxsys_ck/*.h. This Python mirror exists
so that tests and bench.py read like the reference's case files. PyTorch provides device memory and streams only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from . import hostmath as hm


class ParallelDevicePolicy:
    """execution::ParallelDevicePolicy — the only policy this package implements (CUDA, sm_100a)."""


par_device = ParallelDevicePolicy()
MainExecutionPolicy = ParallelDevicePolicy


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


# ------------------------------------------------------------------------------------------------------
# data model
# ------------------------------------------------------------------------------------------------------
_KINDS = {"real": (torch.float32, ()), "vec": (torch.float32, (4,)), "mat": (torch.float32, (9,)), "uint": (torch.int32, ())}


class BaseParticles:
    """Registry of named DiscreteVariables living on the device. Vecd variables are float4 on the device."""

    def __init__(self, ctx: capi.Context, n: int, device):
        self.ctx, self.n, self.device = ctx, int(n), device
        self.vars: dict[str, torch.Tensor] = {}
        self.kinds: dict[str, str] = {}
        self.evolving: list[str] = []
        self.version = 0
        self._shadow: dict[str, torch.Tensor] = {}

    def registerStateVariable(self, name, kind="real", init=0.0):
        if name not in self.vars:
            dt, shape = _KINDS[kind]
            # relation/scan helpers read one entry past n for 'uint' counters; pad every array by one element
            t = torch.zeros((self.n + 1,) + shape, dtype=dt, device=self.device)
            if kind == "mat" and init == "identity":
                t[:, 0] = 1.0
                t[:, 4] = 1.0
                t[:, 8] = 1.0
            elif init not in (0, 0.0, "identity"):
                t.fill_(init)
            self.vars[name] = t
            self.kinds[name] = kind
        return self.vars[name]

    def getVariableByName(self, name):
        if name not in self.vars:
            raise KeyError(f"variable '{name}' is not registered")  # sphinxsys_variable.h:252-257 semantics
        return self.vars[name]

    def addEvolvingVariable(self, name):
        if name not in self.evolving:
            self.evolving.append(name)

    # --- explicit synchronisation (DiscreteVariable::synchronizeToDevice / synchronizeWithDevice) ---
    def upload(self, name, host: np.ndarray, pinned: torch.Tensor | None = None):
        t = self.vars[name]
        kind = self.kinds[name]
        n = self.n
        if kind == "vec":
            h = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32).reshape(n, 3)) if pinned is None else pinned
            staging = h.to(self.device, non_blocking=True)
            self.ctx.call("sphb200_vec3_to_vec4", _ptr(t), _ptr(staging), n, _stream())
        elif kind == "uint":
            h = torch.from_numpy(np.ascontiguousarray(host).astype(np.int32)) if pinned is None else pinned
            t[:n].copy_(h.to(self.device, non_blocking=True).reshape(t[:n].shape))
        else:
            h = torch.from_numpy(np.ascontiguousarray(host, dtype=np.float32)) if pinned is None else pinned
            t[:n].copy_(h.to(self.device, non_blocking=True).reshape(t[:n].shape))
        self.version += 1

    def download(self, name) -> np.ndarray:
        t = self.vars[name]
        n = self.n
        if self.kinds[name] == "vec":
            out = torch.empty((n, 3), dtype=torch.float32, device=self.device)
            self.ctx.call("sphb200_vec4_to_vec3", _ptr(out), _ptr(t), n, _stream())
            return out.cpu().numpy()
        a = t[:n].cpu().numpy()
        return a.view(np.uint32) if self.kinds[name] == "uint" else a


class SPHBody:
    def __init__(self, ctx, name, n, kernel: hm.KernelSpec, mesh: hm.MeshSpec, device):
        self.ctx, self.name, self.kernel, self.mesh, self.device = ctx, name, kernel, mesh, device
        self.particles = BaseParticles(ctx, n, device)
        p = self.particles
        p.registerStateVariable("Position", "vec")
        p.registerStateVariable("VolumetricMeasure", "real")
        p.registerStateVariable("PosVol", "vec")  # derived gather record
        self.cell_linked_list = None
        self.posvol_dirty = True

    @property
    def n(self):
        return self.particles.n

    def getCellLinkedList(self):
        if self.cell_linked_list is None:
            self.cell_linked_list = CellLinkedList(self)
        return self.cell_linked_list

    def refresh_posvol(self):
        if self.posvol_dirty:
            p = self.particles
            self.ctx.call("sphb200_pack_posvol", _ptr(p.vars["PosVol"]), _ptr(p.vars["Position"]),
                          _ptr(p.vars["VolumetricMeasure"]), p.n, _stream())
            self.posvol_dirty = False


class FluidBody(SPHBody):
    def __init__(self, ctx, name, n, kernel, mesh, device, rho0, c0):
        super().__init__(ctx, name, n, kernel, mesh, device)
        self.rho0, self.c0 = float(rho0), float(c0)
        p = self.particles
        # base_particles.cpp:33-36 + base_material.cpp:37-40
        p.registerStateVariable("Density", "real", rho0)
        p.registerStateVariable("Mass", "real")
        p.registerStateVariable("OriginalID", "uint")
        p.registerStateVariable("SortedID", "uint")
        ids = torch.arange(n + 1, dtype=torch.int32, device=device)
        p.vars["OriginalID"].copy_(ids)
        p.vars["SortedID"].copy_(ids)
        for v in ("Position", "VolumetricMeasure", "OriginalID"):
            p.addEvolvingVariable(v)


class SolidBody(SPHBody):
    def __init__(self, ctx, name, n, kernel, mesh, device):
        super().__init__(ctx, name, n, kernel, mesh, device)
        p = self.particles
        p.registerStateVariable("NormalDirection", "vec")
        p.registerStateVariable("VolumetricMeasureRef", "real")


class CellLinkedList:
    """CellLinkedList<SPHAdaptation>: cell_offset[cells+1], particle_index[max(n,cells)] (cell_linked_list.cpp:167-175)."""

    def __init__(self, body: SPHBody):
        self.body = body
        self.mesh_t = capi.mesh_t(body.mesh)
        cells = body.mesh.total_cells
        self.cell_offset = torch.zeros(cells + 2, dtype=torch.int32, device=body.device)
        self.particle_index = torch.zeros(max(body.n, 1) + 1, dtype=torch.int32, device=body.device)
        # cell-ordered copy of Position: contiguous candidate runs for the neighbour search (library extension)
        self.sorted_pos = torch.zeros((max(body.n, 1), 4), dtype=torch.float32, device=body.device)

    def view(self) -> capi.CellListT:
        return capi.CellListT(_ptr(self.cell_offset), _ptr(self.particle_index), _ptr(self.sorted_pos))


class _Relation:
    def __init__(self, source: SPHBody, target: SPHBody, is_inner: bool):
        self.source, self.target, self.is_inner = source, target, is_inner
        n = source.n
        dev = source.device
        self.count = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        self.slice_offset = torch.zeros((n + 31) // 32 + 2, dtype=torch.int32, device=dev)
        # initial capacity ParticlesBound + 1 as in relation_ck.hpp:17,28-31 (the first exec always grows it)
        self.capacity = n + 1
        self.index = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.total = 0
        self.version = 0
        # one-pass build: fixed slice stride (rows per slot); 0 selects the exact count -> scan -> fill build
        self.fixed_stride = 0
        self.max_count = 0
        h = max(source.kernel.h, target.kernel.h)
        self.kernel_t = capi.kernel_t(source.kernel if source.kernel.h >= target.kernel.h else target.kernel, src_h=source.kernel.h)
        self.kernel_t.h = h
        # CellLinkedList::NeighborSearch::ContactSearchBox, cell_linked_list.hpp:161-167
        cut = source.kernel.kernel_size * source.kernel.h
        sp = target.mesh.spacing
        self.search_depth = 1 if is_inner else int(np.ceil((max(sp, cut) - np.finfo(np.float32).eps) / sp))

    def view(self) -> capi.RelationT:
        order = self.source.getCellLinkedList().particle_index
        return capi.RelationT(_ptr(self.count), _ptr(self.slice_offset), _ptr(self.index), self.capacity, _ptr(order))

    def search(self) -> capi.SearchT:
        src, tar = self.source, self.target
        scl, tcl = src.getCellLinkedList(), tar.getCellLinkedList()
        return capi.SearchT(tcl.mesh_t, self.kernel_t, _ptr(src.particles.vars["Position"]), src.n, _ptr(scl.particle_index),
                            _ptr(scl.sorted_pos), _ptr(tar.particles.vars["Position"]), tcl.view(), int(self.is_inner),
                            self.search_depth)

    def export_csr(self):
        """Reference layout (particle_offset_[n+1], neighbor_index_[total]) as numpy, for parity checks."""
        ctx, n = self.source.ctx, self.source.n
        off = torch.zeros(n + 2, dtype=torch.int32, device=self.source.device)
        idx = torch.zeros(max(self.total, 1), dtype=torch.int32, device=self.source.device)
        ctx.call("sphb200_relation_export_csr", self.view(), n, _ptr(off), _ptr(idx), max(self.total, 1), _stream())
        return off[: n + 1].cpu().numpy().view(np.uint32), idx.cpu().numpy().view(np.uint32)


class Inner(_Relation):
    def __init__(self, body: SPHBody):
        super().__init__(body, body, True)


class Contact(_Relation):
    def __init__(self, source: SPHBody, contact_bodies):
        if len(contact_bodies) != 1:
            raise NotImplementedError("the hot path covers one contact (wall) body")
        super().__init__(source, contact_bodies[0], False)


# ------------------------------------------------------------------------------------------------------
# configuration dynamics
# ------------------------------------------------------------------------------------------------------
class UpdateCellLinkedList:
    def __init__(self, policy, body: SPHBody):
        self.body = body
        self.cll = body.getCellLinkedList()

    def exec(self, dt=0.0):
        b = self.body
        b.ctx.call("sphb200_cell_list_build", C.byref(self.cll.mesh_t), _ptr(b.particles.vars["Position"]), b.n,
                   self.cll.view(), _stream())


class UpdateRelation:
    """UpdateRelation<Policy, Inner<>, Contact<>>: exec() runs every relation in order."""

    def __init__(self, policy, *relations):
        self.relations = relations

    @staticmethod
    def _grow(r, entries):
        # DiscreteVariable::reallocateData: 1.25 x the required size, no copy (sphinxsys_variable.h:368-375)
        r.capacity = int(entries * 1.25) + 1
        r.index = torch.empty(r.capacity, dtype=torch.int32, device=r.source.device)
        r.version += 1

    def exec(self, dt=0.0):
        for r in self.relations:
            src = r.source
            ctx = src.ctx
            search = r.search()
            if r.fixed_stride:
                need = ((src.n + 31) // 32) * 32 * r.fixed_stride
                if need > r.capacity:
                    self._grow(r, need)
                mx = C.c_uint32(0)
                ctx.call("sphb200_relation_build_fixed", C.byref(search), r.view(), r.fixed_stride, C.byref(mx), _stream())
                r.max_count = int(mx.value)
                r.total = need
                if r.max_count <= r.fixed_stride:
                    continue
                r.fixed_stride = 0  # a row overflowed the stride: rebuild exactly, and stay exact from now on
            required = C.c_uint64(0)
            ctx.call("sphb200_relation_count", C.byref(search), r.view(), C.byref(required), _stream())
            r.total = int(required.value)
            if r.total > r.capacity:
                self._grow(r, r.total)
            ctx.call("sphb200_relation_fill", C.byref(search), r.view(), _stream())


class ParticleSortCK:
    """key build -> stable radix sort by key -> permute every evolving variable -> sorted-id rebuild."""

    def __init__(self, policy, body: FluidBody):
        self.body = body
        n = body.n
        self.sequence = torch.zeros(n + 1, dtype=torch.int32, device=body.device)
        self.permutation = torch.zeros(n + 1, dtype=torch.int32, device=body.device)

    def exec(self, dt=0.0):
        b = self.body
        p = b.particles
        ctx = b.ctx
        n = p.n
        mesh_t = b.getCellLinkedList().mesh_t
        ctx.call("sphb200_morton_keys", C.byref(mesh_t), _ptr(p.vars["Position"]), n, _ptr(self.sequence),
                 _ptr(self.permutation), C.c_void_p(0), _stream())
        ctx.call("sphb200_sort_pairs_u32", _ptr(self.sequence), _ptr(self.permutation), n, 30, _stream())
        names = list(p.evolving)
        k = len(names)
        dst = (C.c_void_p * k)()
        srcs = (C.c_void_p * k)()
        nbytes = (C.c_uint32 * k)()
        for i, nm in enumerate(names):
            t = p.vars[nm]
            if nm not in p._shadow:
                p._shadow[nm] = torch.empty_like(t)
            dst[i] = p._shadow[nm].data_ptr()
            srcs[i] = t.data_ptr()
            nbytes[i] = t.element_size() * (t[0].numel())
        ctx.call("sphb200_gather_multi", k, dst, srcs, nbytes, _ptr(self.permutation), n, _stream())
        for nm in names:
            p.vars[nm], p._shadow[nm] = p._shadow[nm], p.vars[nm]
        p.version += 1
        b.posvol_dirty = True
        ctx.call("sphb200_update_sorted_id", _ptr(p.vars["OriginalID"]), _ptr(p.vars["SortedID"]), n, _stream())


# ------------------------------------------------------------------------------------------------------
# fluid dynamics: local-dynamics descriptors + the algorithm classes that execute them
# ------------------------------------------------------------------------------------------------------
class _FluidSystem:
    """Shared argument block (the reference's per-dynamics ComputingKernel PODs) for one fluid + wall pair."""

    def __init__(self, inner: Inner, contact: Contact | None, riemann=1, correction=0, free_surface=1):
        self.inner, self.contact = inner, contact
        self.fluid: FluidBody = inner.source
        self.wall: SolidBody | None = contact.target if contact is not None else None
        self.riemann, self.correction, self.free_surface = riemann, correction, free_surface
        p = self.fluid.particles
        # AcousticStep constructor: acoustic_step_1st_half.hpp:13-39
        p.registerStateVariable("Pressure", "real")
        p.registerStateVariable("Compression", "real", 1.0)
        p.registerStateVariable("CompressionRate", "real")
        p.registerStateVariable("Velocity", "vec")
        p.registerStateVariable("Displacement", "vec")
        p.registerStateVariable("Force", "vec")
        p.registerStateVariable("ForcePrior", "vec")
        for v in ("Velocity", "Mass", "ForcePrior", "Compression", "CompressionRate"):
            p.addEvolvingVariable(v)
        # CompressionSummation constructor: density_regularization.hpp:14-27
        p.registerStateVariable("VolumetricMeasureRef", "real")
        p.registerStateVariable("CompressionSummation", "real", 1.0)
        p.addEvolvingVariable("VolumetricMeasureRef")
        if correction:
            p.registerStateVariable("LinearCorrectionMatrix", "mat", "identity")
        self._cache = None
        self._cache_key = None

    def args(self) -> capi.FluidArgs:
        f, w = self.fluid, self.wall
        key = (f.particles.version, self.inner.version, self.contact.version if self.contact else 0,
               w.particles.version if w else 0)
        if self._cache is not None and key == self._cache_key:
            return self._cache
        v = f.particles.vars
        a = capi.FluidArgs()
        a.fluid = capi.FluidView(f.n, _ptr(v["Position"]), _ptr(v["Velocity"]), _ptr(v["Displacement"]), _ptr(v["Force"]),
                                 _ptr(v["ForcePrior"]), _ptr(v["VolumetricMeasure"]), _ptr(v["Mass"]), _ptr(v["Density"]),
                                 _ptr(v["Pressure"]), _ptr(v["Compression"]), _ptr(v["CompressionRate"]),
                                 _ptr(v["VolumetricMeasureRef"]), _ptr(v["CompressionSummation"]),
                                 _ptr(v.get("LinearCorrectionMatrix")), _ptr(v["PosVol"]))
        if w is not None and w.n:
            wv = w.particles.vars
            a.wall = capi.WallView(w.n, _ptr(wv["Position"]), _ptr(wv["PosVol"]), _ptr(wv.get("AverageVelocity")),
                                   _ptr(wv.get("AverageAcceleration")), _ptr(wv["NormalDirection"]),
                                   _ptr(wv["VolumetricMeasureRef"]))
            a.contact = self.contact.view()
        else:
            a.wall = capi.WallView(0, None, None, None, None, None, None)
        a.inner = self.inner.view()
        a.kernel = self.inner.kernel_t
        a.material = capi.FluidT(f.rho0, f.c0, self.riemann, self.correction, 3.0, self.free_surface)
        self._cache, self._cache_key = a, key
        return a

    def prepare(self):
        self.fluid.refresh_posvol()
        if self.wall is not None:
            self.wall.refresh_posvol()


# local dynamics descriptors (type tags with the reference names)
class GravityForceCK: pass
class AdvectionStepSetup: pass
class UpdateParticlePosition: pass
class AdvectionTimeStepCK: pass
class AcousticTimeStepCK: pass
class DensityRegularization: pass
class CompressionSummation: pass
class AcousticStep1stHalfWithWallRiemannCK: pass
class AcousticStep2ndHalfWithWallRiemannCK: pass
class AcousticStep1stHalfWithWallRiemannCorrectionCK: pass
class AcousticStep2ndHalfWithWallRiemannCorrectionCK: pass
class LinearCorrectionMatrixComplex: pass
class TotalMechanicalEnergyCK: pass


class Gravity:
    def __init__(self, vector, reference_position=(0.0, 0.0, 0.0)):
        v = tuple(float(x) for x in vector) + (0.0,) * (3 - len(vector))
        self.vector = v
        self.c = (C.c_float * 3)(*v)


class StateDynamics:
    """StateDynamics<Policy, LocalDynamics>::exec(dt): one update(i, dt) sweep (simple_algorithms_ck.h:59-74)."""

    def __init__(self, policy, local_dynamics, system: _FluidSystem, *args):
        self.kind, self.sys, self.extra = local_dynamics, system, args
        p = system.fluid.particles
        if local_dynamics is GravityForceCK:
            p.registerStateVariable("PreviousGravityForceCK", "vec")
            p.addEvolvingVariable("ForcePrior")
            p.addEvolvingVariable("PreviousGravityForceCK")

    def exec(self, dt=0.0):
        s = self.sys
        ctx = s.fluid.ctx
        a = s.args()
        if self.kind is GravityForceCK:
            g: Gravity = self.extra[0]
            ctx.call("sphb200_gravity_force", C.byref(a.fluid), C.byref(g.c),
                     _ptr(s.fluid.particles.vars["PreviousGravityForceCK"]), _stream())
        elif self.kind is AdvectionStepSetup:
            ctx.call("sphb200_advection_setup", C.byref(a.fluid), _stream())
            s.fluid.posvol_dirty = True
        elif self.kind is UpdateParticlePosition:
            ctx.call("sphb200_update_position", C.byref(a.fluid), _stream())
            s.fluid.posvol_dirty = True
        elif self.kind is DensityRegularization:
            ctx.call("sphb200_density_regularization", C.byref(a), _stream())
        else:
            raise NotImplementedError(self.kind)


class ReduceDynamicsCK:
    """ReduceDynamicsCK<Policy, LocalDynamicsReduce>::exec() -> FinishDynamics::Result(reduced) on the host."""

    def __init__(self, policy, local_dynamics, system: _FluidSystem, *args):
        self.kind, self.sys, self.extra = local_dynamics, system, args
        self.h_min = system.fluid.kernel.h
        self.last_reduced = None

    def exec(self, dt=0.0):
        s = self.sys
        ctx = s.fluid.ctx
        a = s.args()
        red, out = C.c_float(0), C.c_float(0)
        if self.kind is AdvectionTimeStepCK:
            u_ref = float(self.extra[0])
            cfl = float(self.extra[1]) if len(self.extra) > 1 else 0.25
            ctx.call("sphb200_advection_time_step", C.byref(a.fluid), self.h_min, u_ref, cfl, C.byref(red), C.byref(out), _stream())
        elif self.kind is AcousticTimeStepCK:
            cfl = float(self.extra[0]) if self.extra else 0.6
            ctx.call("sphb200_acoustic_time_step", C.byref(a), self.h_min, cfl, C.byref(red), C.byref(out), _stream())
        elif self.kind is TotalMechanicalEnergyCK:
            g: Gravity = self.extra[0]
            e = C.c_double(0)
            ctx.call("sphb200_total_mechanical_energy", C.byref(a.fluid), C.byref(g.c), C.byref(e), _stream())
            return float(e.value)
        else:
            raise NotImplementedError(self.kind)
        self.last_reduced = float(red.value)
        return float(out.value)


class InteractionDynamicsCK:
    """InteractionDynamicsCK<Policy, Interaction<Inner<...>, Contact<...>>>::exec(dt)."""

    def __init__(self, policy, interaction, system: _FluidSystem, *args):
        self.kind, self.sys, self.extra = interaction, system, args
        corr = interaction in (AcousticStep1stHalfWithWallRiemannCorrectionCK, AcousticStep2ndHalfWithWallRiemannCorrectionCK)
        if corr and not system.correction:
            raise ValueError("Correction variants need a _FluidSystem built with correction=1")
        self.h_min = system.fluid.kernel.h
        self.next_reduced = None  # device float for the fused acoustic-dt reduction (2nd half)

    def enable_fused_time_step(self):
        self.next_reduced = torch.zeros(1, dtype=torch.float32, device=self.sys.fluid.device)
        return self.next_reduced

    def exec(self, dt=0.0):
        s = self.sys
        ctx = s.fluid.ctx
        s.prepare()
        a = s.args()
        k = self.kind
        if k is CompressionSummation:
            regularize = int(self.extra[0]) if self.extra else 0
            ctx.call("sphb200_compression_summation", C.byref(a), regularize, _stream())
        elif k in (AcousticStep1stHalfWithWallRiemannCK, AcousticStep1stHalfWithWallRiemannCorrectionCK):
            ctx.call("sphb200_acoustic_1st_half", C.byref(a), float(dt), _stream())
        elif k in (AcousticStep2ndHalfWithWallRiemannCK, AcousticStep2ndHalfWithWallRiemannCorrectionCK):
            nr = self.next_reduced
            if nr is not None:
                nr.zero_()
            ctx.call("sphb200_acoustic_2nd_half", C.byref(a), float(dt), self.h_min, _ptr(nr), _stream())
        elif k is LinearCorrectionMatrixComplex:
            alpha = float(self.extra[0]) if self.extra else 0.5
            ctx.call("sphb200_linear_correction_matrix", C.byref(a), alpha, _stream())
        else:
            raise NotImplementedError(k)
