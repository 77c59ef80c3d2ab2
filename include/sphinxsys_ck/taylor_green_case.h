// sphinxsys_ck/taylor_green_case.h — periodic Taylor-Green vortex (2-D and 3-D) on the CK dynamics: the periodic
// boundary path of the hot path (BASELINE config 4).
//
// The reference offers periodic conditions on its TBB path only; this case is the CK spelling of
//   tests/2d_examples/test_2d_taylor_green/taylor_green.cpp:62-200 (bodies, initial condition, loop order:
//       bounding -> cell-linked list -> periodic images -> configuration)
// with the ghost-particle form of the periodic condition used as in
//   tests/2d_examples/test_2d_throat/throat.cpp:137-138,182-184,204-205,271-278
//       (Ghost<PeriodicAlongAxis>, generateParticlesWithReserve, ghost_update_ as pre-process of the half steps)
// and the acoustic/advection sequencing of tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/dambreak.cpp:188-222.
// Inviscid, no transport-velocity correction (those dynamics are outside SURVEY.md §8a).
#ifndef SPHINXSYS_CK_TAYLOR_GREEN_CASE_H
#define SPHINXSYS_CK_TAYLOR_GREEN_CASE_H

#include <functional>

#include "sphinxsys_ck.h"

namespace SPH
{
struct TaylorGreenParameters
{
    int dim = 3;
    double dp = 1.0 / 32.0;   // global_resolution
    double L = 1.0;           // DL = DH (= DW): the periodic box is [0, L]^dim (taylor_green.cpp:14-15)
    double x_scale = 1.0;     // box [0, x_scale L] x [0, L]^(dim-1): the domain replicated along x for weak scaling
    double rho0_f = 1.0, U_f = 1.0; // :19-20; c_f = 10 U_f
    int sort_interval = 100;
    bool fused_time_step = true;
    bool fused_regularization = true;
    // ring decomposition along x (SURVEY.md §8e, config 4 on N GPUs): one process per GPU, the communicator of the
    // context must be a ring already (sphb200_comm_create or _create_self, then sphb200_comm_set_ring)
    bool ring = false;
    int rank = 0, nranks = 1;
    double mu_f = 0.0;               // > 0: Viscosity closure + ViscousForceInnerCK in the loop (taylor_green.cpp:21-22,108: rho0 U L / Re)
    bool transport_velocity = false; // KernelGradientIntegralInner + TransportVelocityCorrectionCK<SPHBody, TruncatedLinear> (:109)
};

class TaylorGreenWaterBlock : public ComplexShape
{
  public:
    TaylorGreenWaterBlock(const std::string &name, const TaylorGreenParameters &q) : ComplexShape(name)
    {
        double half[3] = {0.5 * q.L * q.x_scale, 0.5 * q.L, q.dim == 3 ? 0.5 * q.L : 0.0};
        add<GeometricShapeBox>(half, half);
    }
};

class TaylorGreenCK
{
  public:
    using P = MainExecutionPolicy;
    using GhostUpdate = PeriodicConditionUsingGhostParticles::Update;
    // ghost_update_ of a ring-decomposed run: first the ghost planes of the neighbour slabs (over the seam where needed),
    // then the images of the other axes, which copy from own particles and ghost planes alike
    class SeamGhostUpdate : public BaseDynamics<void>
    {
        SlabDecomposition &decomposition_;
        PeriodicImages &images_;
        std::vector<std::string> names_;

      public:
        SeamGhostUpdate(SlabDecomposition &d, PeriodicImages &images, std::initializer_list<const char *> names)
            : decomposition_(d), images_(images), names_(names.begin(), names.end()) {}
        void exec(Real = 0.0) override
        {
            images_.ensure(); // pending image creation changes the stored total, not the ghost planes: do it first
            decomposition_.refreshGhosts(names_);
            images_.update(names_);
        }
    };
    TaylorGreenParameters q_;
    Real U_f_, c_f_;
    SPHSystem sph_system;
    FluidBody water_block;
    std::vector<std::unique_ptr<Ghost<PeriodicAlongAxis>>> ghost_along_axis;
    std::unique_ptr<Inner<>> water_block_inner;
    std::unique_ptr<UpdateCellLinkedList<P, RealBody>> water_cell_linked_list;
    std::unique_ptr<UpdateRelation<P, Inner<>>> water_block_update_inner_relation;
    std::unique_ptr<ParticleSortCK<P>> particle_sort;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::AdvectionStepSetup>> water_advection_step_setup;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::UpdateParticlePosition>> water_update_particle_position;
    std::unique_ptr<InteractionDynamicsCK<P, fluid_dynamics::AcousticStep1stHalfInnerRiemannCK>> fluid_acoustic_step_1st_half;
    std::unique_ptr<InteractionDynamicsCK<P, fluid_dynamics::AcousticStep2ndHalfInnerRiemannCK>> fluid_acoustic_step_2nd_half;
    std::unique_ptr<InteractionDynamicsCK<P, fluid_dynamics::CompressionSummation<Inner<>>>> fluid_density_summation;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::DensityRegularization<SPHBody, WeaklyCompressibleFluid, Internal>>> fluid_density_regularization;
    std::unique_ptr<ReduceDynamicsCK<P, fluid_dynamics::AdvectionTimeStepCK>> fluid_advection_time_step;
    std::unique_ptr<ReduceDynamicsCK<P, fluid_dynamics::AcousticTimeStepCK<WeaklyCompressibleFluid>>> fluid_acoustic_time_step;
    std::vector<std::unique_ptr<PeriodicConditionUsingGhostParticles>> periodic_condition; // x, y(, z)
    std::unique_ptr<BaseDynamics<void>> volume_ghost_update, pressure_ghost_update, velocity_ghost_update;
    std::unique_ptr<SlabDecomposition> decomposition; // ring runs only
    SeamRing seam_ring;
    std::unique_ptr<InteractionDynamicsCK<P, fluid_dynamics::ViscousForceInnerCK>> viscous_force;
    std::unique_ptr<InteractionDynamicsCK<P, KernelGradientIntegralInner>> kernel_gradient_integral;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::TransportVelocityCorrectionCK<SPHBody, TruncatedLinear>>> transport_velocity_correction;
    Gravity no_gravity;
    std::unique_ptr<ReduceDynamicsCK<P, TotalMechanicalEnergyCK>> record_total_kinetic_energy; // zero gravity: kinetic part only
    SingleVariable<Real> *sv_physical_time = nullptr;
    size_t number_of_iterations = 0, acoustic_steps = 0;
    double physical_time = 0;
    Real last_acoustic_dt = 0, last_advection_dt = 0;

    static BoundingBoxd caseBounds(const TaylorGreenParameters &q)
    {
        return BoundingBoxd(Vecd(0, 0, 0), Vecd(Real(q.L * q.x_scale), Real(q.L), q.dim == 3 ? Real(q.L) : Real(0)));
    }
    // TaylorGreenInitialCondition (taylor_green.cpp:44-57) and its usual 3-D extension
    static Vecd initialVelocity(const Vecd &x, int dim, Real U)
    {
        const double two_pi = 2.0 * 3.14159265358979323846;
        if (dim == 2)
            return Vecd(Real(-U * std::cos(two_pi * x.x) * std::sin(two_pi * x.y)), Real(U * std::sin(two_pi * x.x) * std::cos(two_pi * x.y)), 0);
        return Vecd(Real(U * std::sin(two_pi * x.x) * std::cos(two_pi * x.y) * std::cos(two_pi * x.z)),
                    Real(-U * std::cos(two_pi * x.x) * std::sin(two_pi * x.y) * std::cos(two_pi * x.z)), 0);
    }

    // positions / velocities == nullptr: lattice and analytic initial condition generated here; otherwise the
    // arrays handed over (packed xyz, reference order), e.g. a jittered lattice made by the harness
    // ring runs: positions / velocities / reference_ids are THIS rank's particles (global numbering in reference_ids)
    explicit TaylorGreenCK(const TaylorGreenParameters &q, const std::vector<Vecd> *positions = nullptr,
                           const std::vector<Vecd> *velocities = nullptr, const BoundingBoxd *exact_system_bounds = nullptr,
                           const std::vector<UnsignedInt> *reference_ids = nullptr)
        : q_(q), U_f_(Real(q.U_f)), c_f_(Real(10.0) * U_f_), sph_system(caseBounds(q), Real(q.dp), q.dim),
          water_block(sph_system, makeShared<TaylorGreenWaterBlock>("WaterBody", q)), no_gravity(Vecd(0, 0, 0))
    {
        using namespace fluid_dynamics;
        if (exact_system_bounds) sph_system.setSystemDomainBoundsExact(*exact_system_bounds);
        if (q.mu_f > 0) water_block.defineClosure<WeaklyCompressibleFluid, Viscosity>(Real(q.rho0_f), c_f_, Real(q.mu_f));
        else water_block.defineMatterMaterial<WeaklyCompressibleFluid>(Real(q.rho0_f), c_f_);
        BoundingBoxd box = water_block.getSPHBodyBounds();
        for (int a = 0; a < q.dim; ++a) ghost_along_axis.emplace_back(new Ghost<PeriodicAlongAxis>(box, a));
        if (q.dim == 3) water_block.reserveFor(*ghost_along_axis[0], *ghost_along_axis[1], *ghost_along_axis[2]);
        else water_block.reserveFor(*ghost_along_axis[0], *ghost_along_axis[1]);
        std::vector<int> cuts;
        if (q.ring)
        {
            // cell planes that tile the box; equal plane shares (the lattice is uniform); storage for the ghost planes of
            // both neighbours, for migrants and for the images of the other axes of all of them
            if (!positions) throw SphError("TaylorGreenCK: a ring run takes this rank's particles from the caller");
            sphb200_mesh_t mesh = alignedPeriodicMesh(box, water_block.getSPHAdaptation().CutOffRadius(), q.dim, seam_ring);
            if (seam_ring.box_planes() < q.nranks) throw SphError("TaylorGreenCK: fewer cell planes than ranks");
            for (int r = 0; r <= q.nranks; ++r) cuts.push_back(seam_ring.first_plane() + (int)((long)seam_ring.box_planes() * r / q.nranks));
            const int own_planes = cuts[q.rank + 1] - cuts[q.rank];
            const size_t n = positions->size(), per_plane = n / (size_t)own_planes + 1;
            // images of the other axes: Ghost<PeriodicAlongAxis>::reserveSize with the slab (own + 2 ghost planes) as x extent
            size_t reserve = 0;
            for (int a = 1; a < q.dim; ++a)
            {
                double face = 1.0;
                for (int d = 0; d < q.dim; ++d)
                    if (d != a) face *= (d == 0 ? double(own_planes + 2) * double(mesh.spacing) : double(box.upper_[d] - box.lower_[d])) / q.dp + 8.0;
                reserve += (size_t)std::ceil(2.0 * 4.0 * face);
            }
            const size_t bound = n + 4 * per_plane + 2 * reserve + 4096;
            water_block.generateParticlesFromPositions(*positions, Real(std::pow(Real(q.dp), Real(q.dim))), bound, reference_ids);
            water_block.getCellLinkedList().resetMesh(mesh, water_block.getBaseParticles().ParticlesBound());
        }
        else if (positions) water_block.generateParticlesFromPositions(*positions, Real(std::pow(Real(q.dp), Real(q.dim))));
        else water_block.generateParticles<BaseParticles, Lattice>();

        water_block_inner.reset(new Inner<>(water_block));
        water_cell_linked_list.reset(new UpdateCellLinkedList<P, RealBody>(water_block));
        water_block_update_inner_relation.reset(new UpdateRelation<P, Inner<>>(*water_block_inner));
        particle_sort.reset(new ParticleSortCK<P>(water_block));
        water_advection_step_setup.reset(new StateDynamics<P, AdvectionStepSetup>(water_block));
        water_update_particle_position.reset(new StateDynamics<P, UpdateParticlePosition>(water_block));
        fluid_acoustic_step_1st_half.reset(new InteractionDynamicsCK<P, AcousticStep1stHalfInnerRiemannCK>(*water_block_inner));
        fluid_acoustic_step_2nd_half.reset(new InteractionDynamicsCK<P, AcousticStep2ndHalfInnerRiemannCK>(*water_block_inner));
        fluid_density_summation.reset(new InteractionDynamicsCK<P, CompressionSummation<Inner<>>>(*water_block_inner));
        fluid_density_regularization.reset(new StateDynamics<P, DensityRegularization<SPHBody, WeaklyCompressibleFluid, Internal>>(water_block));
        if (q.fused_regularization) fluid_density_summation->addPostStateDynamics(*fluid_density_regularization);
        fluid_advection_time_step.reset(new ReduceDynamicsCK<P, AdvectionTimeStepCK>(water_block, U_f_));
        fluid_acoustic_time_step.reset(new ReduceDynamicsCK<P, AcousticTimeStepCK<WeaklyCompressibleFluid>>(water_block));
        if (q.fused_time_step) fluid_acoustic_step_2nd_half->fuseTimeStepReduction(*fluid_acoustic_time_step);
        // ring runs: x is periodic through the slab exchange, images serve the other axes only
        for (int a = q.ring ? 1 : 0; a < q.dim; ++a)
            periodic_condition.emplace_back(new PeriodicConditionUsingGhostParticles(water_block, *ghost_along_axis[a]));
        // what the neighbours of a ghost read, refreshed where it changes (throat.cpp:183-184 queues ghost_update_ the same way)
        PeriodicImages &images = periodic_condition[0]->images();
        if (q.ring)
        {
            // viscous force, kernel gradient integral and transport correction read Position, VolumetricMeasure and Velocity of
            // the neighbours, all current on ghost planes and images where they run: the three refreshes below suffice
            // (pinned on the CPU: tests/test_decomposed_oracle_cpu.py::test_periodic_ring_viscous_transport_bit_identical)
            decomposition.reset(new SlabDecomposition(water_block, q.rank, q.nranks, cuts, seam_ring));
            fluid_advection_time_step->setDecomposition(decomposition.get());
            fluid_acoustic_time_step->setDecomposition(decomposition.get());
            volume_ghost_update.reset(new SeamGhostUpdate(*decomposition, images, {"VolumetricMeasure"}));
            pressure_ghost_update.reset(new SeamGhostUpdate(*decomposition, images, {"Pressure"}));
            velocity_ghost_update.reset(new SeamGhostUpdate(*decomposition, images, {"PosVolVel"}));
        }
        else
        {
            volume_ghost_update.reset(new GhostUpdate(images, {"VolumetricMeasure"}));
            pressure_ghost_update.reset(new GhostUpdate(images, {"Pressure"}));
            velocity_ghost_update.reset(new GhostUpdate(images, {"PosVolVel"}));
        }
        fluid_acoustic_step_1st_half->addPreContactInteraction(*pressure_ghost_update);
        fluid_acoustic_step_2nd_half->addPreContactInteraction(*velocity_ghost_update);
        if (q.mu_f > 0) viscous_force.reset(new InteractionDynamicsCK<P, ViscousForceInnerCK>(*water_block_inner));
        if (q.transport_velocity)
        {
            kernel_gradient_integral.reset(new InteractionDynamicsCK<P, KernelGradientIntegralInner>(*water_block_inner));
            transport_velocity_correction.reset(new StateDynamics<P, TransportVelocityCorrectionCK<SPHBody, TruncatedLinear>>(water_block));
        }
        record_total_kinetic_energy.reset(new ReduceDynamicsCK<P, TotalMechanicalEnergyCK>(water_block, no_gravity));
        if (decomposition) record_total_kinetic_energy->setDecomposition(decomposition.get());
        sv_physical_time = sph_system.getSystemVariableByName<Real>("PhysicalTime");

        // initial_condition.exec()
        BaseParticles &particles = water_block.getBaseParticles();
        auto *dv_vel = particles.getVariableByName<Vecd>("Velocity");
        size_t n = particles.TotalRealParticles();
        if (velocities)
        {
            if (velocities->size() != n) throw SphError("TaylorGreenCK: velocity array size mismatch");
            for (size_t i = 0; i < n; ++i) dv_vel->Data()[i] = (*velocities)[i];
        }
        else
        {
            auto *dv_pos = particles.getVariableByName<Vecd>("Position");
            dv_pos->synchronizeWithDevice();
            for (size_t i = 0; i < n; ++i) dv_vel->Data()[i] = initialVelocity(dv_pos->Data()[i], q.dim, U_f_);
        }
        dv_vel->synchronizeToDevice();
    }

    // bounding -> cell-linked list -> periodic images -> configuration (taylor_green.cpp:186-191)
    void updateConfiguration(bool bounding)
    {
        if (bounding)
            for (auto &pc : periodic_condition) SPHCK_STAGE("periodic bounding", pc->bounding_.exec());
        // ring: x is bounded by migration (what leaves the box over the seam arrives shifted on the other side)
        if (decomposition) SPHCK_STAGE("slab rebuild (migration, ghost planes, cell list)", decomposition->rebuild());
        else SPHCK_STAGE("cell list", water_cell_linked_list->exec());
        for (auto &pc : periodic_condition) SPHCK_STAGE("ghost creation", pc->ghost_creation_.exec());
        SPHCK_STAGE("relation", water_block_update_inner_relation->exec());
    }
    void initialize()
    {
        updateConfiguration(false); // taylor_green.cpp:131-134
        fluid_acoustic_time_step->setPrimed(false);
    }

    // one advection step; returns the number of acoustic sub-steps taken
    int stepOuter()
    {
        SPHCK_STAGE("density summation", fluid_density_summation->exec());
        if (!q_.fused_regularization) fluid_density_regularization->exec();
        SPHCK_STAGE("advection setup", water_advection_step_setup->exec());
        SPHCK_STAGE("volume ghost update", volume_ghost_update->exec()); // neighbours read V_j of the images
        // viscous force, kernel gradient integral, transport correction: order of lid_driven_cavity_sycl.cpp:268-276
        if (viscous_force) viscous_force->exec();
        if (kernel_gradient_integral)
        {
            kernel_gradient_integral->exec();
            transport_velocity_correction->exec();
        }
        Real advection_dt = 0;
        SPHCK_STAGE("advection dt", advection_dt = fluid_advection_time_step->exec());
        Real relaxation_time = 0, acoustic_dt = 0;
        int n_inner = 0;
        while (relaxation_time < advection_dt)
        {
            SPHCK_STAGE("acoustic dt", acoustic_dt = fluid_acoustic_time_step->exec());
            SPHCK_STAGE("1st half (+ghost pressure)", fluid_acoustic_step_1st_half->exec(acoustic_dt)); // initialize -> ghost pressure -> interact + update
            SPHCK_STAGE("2nd half (+ghost velocity)", fluid_acoustic_step_2nd_half->exec(acoustic_dt)); // ghost velocity -> one fused launch
            relaxation_time += acoustic_dt;
            physical_time += acoustic_dt;
            sv_physical_time->incrementValue(acoustic_dt);
            ++n_inner;
        }
        acoustic_steps += n_inner;
        SPHCK_STAGE("update position", water_update_particle_position->exec());
        number_of_iterations++;
        if (!decomposition && q_.sort_interval > 0 && number_of_iterations % q_.sort_interval == 0 && number_of_iterations != 1)
        {
            particle_sort->exec();
            fluid_acoustic_time_step->setPrimed(false);
        }
        updateConfiguration(true);
        last_acoustic_dt = acoustic_dt;
        last_advection_dt = advection_dt;
        return n_inner;
    }
};
} // namespace SPH
#endif
