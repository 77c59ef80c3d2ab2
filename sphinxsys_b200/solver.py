"""Dam-break time loop on the GPU, sequenced exactly like the reference case file
tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/dambreak.cpp:98-225 (definitions :98-136, loop :183-225),
built from the dynamics classes in `dynamics.py`.  The 2-D case runs through the same code (z == 0).
"""
from __future__ import annotations

import numpy as np
import torch

from . import capi
from . import cases
from . import dynamics as dyn


class DamBreakCK:
    def __init__(self, case: cases.DamBreakCase, device_index=0, correction=False, riemann=1, fused_time_step=True,
                 sort_interval=100, ctx: capi.Context | None = None, relation_stride=None):
        if not torch.cuda.is_available():
            raise capi.SphB200Error("CUDA device required: libsphb200 has no CPU path")
        self.case = case
        self.device = torch.device("cuda", device_index)
        torch.cuda.set_device(self.device)
        self.ctx = ctx or capi.Context(device_index)
        self.correction = bool(correction)
        self.sort_interval = sort_interval
        ctx_ = self.ctx
        k, mesh = case.kernel, case.mesh
        # ---- bodies (dambreak.cpp:79-91) ----
        self.water_block = dyn.FluidBody(ctx_, "WaterBody", case.n_fluid, k, mesh, self.device, case.rho0, case.c0)
        self.wall_boundary = dyn.SolidBody(ctx_, "WallBoundary", case.n_wall, k, mesh, self.device)
        # ---- relations (:98-100) ----
        self.water_block_inner = dyn.Inner(self.water_block)
        self.water_wall_contact = dyn.Contact(self.water_block, [self.wall_boundary])
        # one-pass relation build with a fixed row stride (falls back to the exact build if a row overflows);
        # relation_stride=0 forces the exact count -> scan -> fill build of the reference
        if relation_stride is None:
            relation_stride = 128 if case.dim == 3 else 40
        self.water_block_inner.fixed_stride = int(relation_stride)
        self.water_wall_contact.fixed_stride = int(relation_stride)
        self.system = dyn._FluidSystem(self.water_block_inner, self.water_wall_contact, riemann=riemann,
                                       correction=int(correction), free_surface=1)
        P = dyn.par_device
        # ---- methods (:112-136) ----
        self.water_cell_linked_list = dyn.UpdateCellLinkedList(P, self.water_block)
        self.wall_cell_linked_list = dyn.UpdateCellLinkedList(P, self.wall_boundary)
        self.water_block_update_complex_relation = dyn.UpdateRelation(P, self.water_block_inner, self.water_wall_contact)
        self.particle_sort = dyn.ParticleSortCK(P, self.water_block)
        self.gravity = dyn.Gravity(case.gravity)
        self.constant_gravity = dyn.StateDynamics(P, dyn.GravityForceCK, self.system, self.gravity)
        self.water_advection_step_setup = dyn.StateDynamics(P, dyn.AdvectionStepSetup, self.system)
        self.water_update_particle_position = dyn.StateDynamics(P, dyn.UpdateParticlePosition, self.system)
        self.fluid_linear_correction_matrix = dyn.InteractionDynamicsCK(P, dyn.LinearCorrectionMatrixComplex, self.system, 0.5)
        a1 = dyn.AcousticStep1stHalfWithWallRiemannCorrectionCK if correction else dyn.AcousticStep1stHalfWithWallRiemannCK
        a2 = dyn.AcousticStep2ndHalfWithWallRiemannCorrectionCK if correction else dyn.AcousticStep2ndHalfWithWallRiemannCK
        self.fluid_acoustic_step_1st_half = dyn.InteractionDynamicsCK(P, a1, self.system)
        self.fluid_acoustic_step_2nd_half = dyn.InteractionDynamicsCK(P, a2, self.system)
        # CompressionSummation<Inner<>,Contact<>> with DensityRegularization fused into the same launch
        self.fluid_density_summation = dyn.InteractionDynamicsCK(P, dyn.CompressionSummation, self.system, 1)
        self.fluid_advection_time_step = dyn.ReduceDynamicsCK(P, dyn.AdvectionTimeStepCK, self.system, case.U_ref)
        self.fluid_acoustic_time_step = dyn.ReduceDynamicsCK(P, dyn.AcousticTimeStepCK, self.system)
        self.record_water_mechanical_energy = dyn.ReduceDynamicsCK(P, dyn.TotalMechanicalEnergyCK, self.system, self.gravity)
        self.fused_time_step = bool(fused_time_step)
        self._next_reduced = self.fluid_acoustic_step_2nd_half.enable_fused_time_step() if fused_time_step else None
        self._fused_valid = False
        self.h_min = k.h
        # ---- counters ----
        self.physical_time = 0.0
        self.number_of_iterations = 0
        self.acoustic_steps = 0
        self.last_acoustic_dt = 0.0
        self.last_advection_dt = 0.0
        self._host_state = None

    # ------------------------------------------------------------------------------------------
    def host_initial_state(self):
        """The particle state the reference's pre-processing would hand over (host arrays, reference layout)."""
        c = self.case
        nf, nw = c.n_fluid, c.n_wall
        return {
            "fluid": {"Position": c.fluid_pos, "VolumetricMeasure": np.full(nf, c.vol, np.float32),
                      "VolumetricMeasureRef": np.full(nf, c.vol, np.float32),
                      "Mass": np.full(nf, c.rho0 * c.vol, np.float32)},
            "wall": {"Position": c.wall_pos, "VolumetricMeasure": np.full(nw, c.vol, np.float32),
                     "VolumetricMeasureRef": np.full(nw, c.vol, np.float32), "NormalDirection": c.wall_normal},
        }

    def upload_state(self, state=None):
        """DiscreteVariable::synchronizeToDevice for every variable of the initial state."""
        state = state or self.host_initial_state()
        fp, wp = self.water_block.particles, self.wall_boundary.particles
        for name, arr in state["fluid"].items():
            fp.upload(name, arr)
        for name, arr in state["wall"].items():
            wp.upload(name, arr)
        self.water_block.posvol_dirty = True
        self.wall_boundary.posvol_dirty = True

    def initialize(self, upload=True):
        """dambreak.cpp:152-160: gravity, both cell lists, relations."""
        if upload:
            self.upload_state()
        self.constant_gravity.exec()
        self.water_cell_linked_list.exec()
        self.wall_cell_linked_list.exec()
        self.water_block_update_complex_relation.exec()
        self._fused_valid = False

    def acoustic_dt(self):
        if self.fused_time_step and self._fused_valid:
            red = float(self._next_reduced.item())  # 4-byte D2H, the same sync the reference's reduce makes
            return float(np.float32(0.6) * np.float32(self.h_min) / (np.float32(red) + np.float32(2.71051e-20)))
        return self.fluid_acoustic_time_step.exec()

    def step_outer(self):
        """One advection step, dambreak.cpp:188-222."""
        self.fluid_density_summation.exec()          # summation + regularisation, one launch
        self.water_advection_step_setup.exec()
        advection_dt = self.fluid_advection_time_step.exec()
        if self.correction:
            self.fluid_linear_correction_matrix.exec()
        relaxation_time = 0.0
        acoustic_dt = 0.0
        n_inner = 0
        while relaxation_time < advection_dt:
            acoustic_dt = self.acoustic_dt()
            self.fluid_acoustic_step_1st_half.exec(acoustic_dt)
            self.fluid_acoustic_step_2nd_half.exec(acoustic_dt)
            self._fused_valid = self.fused_time_step
            relaxation_time += acoustic_dt
            self.physical_time += acoustic_dt
            n_inner += 1
        self.acoustic_steps += n_inner
        self.water_update_particle_position.exec()
        self.number_of_iterations += 1
        if self.sort_interval and self.number_of_iterations % self.sort_interval == 0 and self.number_of_iterations != 1:
            self.particle_sort.exec()
            self._fused_valid = False  # "Force" is not an evolving variable: its pairing with ForcePrior changed
        self.water_cell_linked_list.exec()
        self.water_block_update_complex_relation.exec()
        self.last_acoustic_dt, self.last_advection_dt = acoustic_dt, advection_dt
        return n_inner

    def energy(self):
        return self.record_water_mechanical_energy.exec()

    def download(self, name, wall=False):
        return (self.wall_boundary if wall else self.water_block).particles.download(name)
