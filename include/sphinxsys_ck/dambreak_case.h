// sphinxsys_ck/dambreak_case.h — the 3-D (and 2-D) dam-break case assembled from the host-layer classes, sequenced
// exactly like the reference case file tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/dambreak.cpp:69-225
// (definitions :98-136, preparation :152-160, main loop :183-225). examples/dambreak_ck.cpp drives it from main();
// the Python test/bench harness drives the same object through sphinxsys_b200/csrc/host_api.cpp.
#ifndef SPHINXSYS_CK_DAMBREAK_CASE_H
#define SPHINXSYS_CK_DAMBREAK_CASE_H

#include <functional>

#include "sphinxsys_ck.h"
#include "slab_decomposition.h"
#include "legacy_dynamics.h"

namespace SPH
{
struct DamBreakParameters
{
    int dim = 3;
    double dp = 0.05;          // global_resolution (dambreak.cpp:19)
    double DL = 5.366, DH = 2.0, DW = 0.5, LL = 2.0, LH = 1.0, LW = 0.5; // dambreak.cpp:13-18
    double rho0_f = 1.0, gravity_g = 1.0;                                // :22-23
    bool correction = false;   // LinearCorrectionCK variants (the reference case file uses them; the hot path is without)
    int riemann = 1;           // 0 NoRiemannSolverCK, 1 AcousticRiemannSolverCK (the case file), 2 DissipativeRiemannSolverCK
                               // (riemann_solver_ck.h:46-173; aliases acoustic_step_1st_half.h:196-201)
    int kernel_kind = 0;       // 0 KernelWendlandC2 (default), 1 resetKernel<KernelTabulated<KernelLaguerreGauss>>(20), adaptation.h:96-100
    bool surface_indicator = false; // FreeSurfaceIndicationComplexSpatialTemporalCK in the loop (dambreak.cpp:133-134,192)
    bool observers = false;         // FluidObserver pressure probes of the case file (dambreak.cpp:54-65,87-88,140-141,223-224)
    double mu_f = 0.0;              // > 0: Viscosity closure + ViscousForceWithWallCK after the advection set-up
    bool transport_velocity = false; // KernelGradientIntegral(Corrected)Complex + TransportVelocityCorrectionCK<SPHBody, TruncatedLinear>
                                     // (lid_driven_cavity_sycl.cpp:208-214,268-276)
    bool fused_time_step = true;
    bool fused_regularization = true;
    int sort_interval = 100;   // :217-220
    // legacy API + formulation (tests/2d_examples/test_2d_dambreak/Dambreak.cpp, tests/3d_examples/test_3d_dambreak):
    // Integration1stHalf/2ndHalfWithWallRiemann, DensitySummationComplexFreeSurface, state Density/DensityChangeRate
    bool legacy = false;
    // slab decomposition over the GPUs of one node (needs sphb200_comm_create on this process's context first)
    int rank = 0, nranks = 1;
    int initial_cut_shift = 0;    // decomposed runs: start with the interior cuts moved by so many planes (test hook)
    int recut_interval = 100;     // decomposed runs: re-balance the slab cuts every so many advection steps (0: never)
    bool overlap_exchange = true; // decomposed runs: hide the plane exchange behind interior compute (acousticStepOverlapped)
    bool wall_slabs = true;       // decomposed runs: every rank stores the wall planes around its slab only (WallSlab)
    static DamBreakParameters twoDimensional(double dp = 0.025)
    {
        DamBreakParameters p;
        p.dim = 2; p.dp = dp; p.DH = 5.366; p.DW = 0; p.LW = 0; // tests/2d_examples/test_2d_dambreak/Dambreak.cpp:13-18
        return p;
    }
};

class WaterBlock : public ComplexShape
{
  public:
    WaterBlock(const std::string &name, const DamBreakParameters &q) : ComplexShape(name)
    {
        double half[3] = {0.5 * q.LL, 0.5 * q.LH, 0.5 * q.LW};
        add<GeometricShapeBox>(half, half);
    }
};
class WallBoundary : public ComplexShape
{
  public:
    WallBoundary(const std::string &name, const DamBreakParameters &q) : ComplexShape(name)
    {
        double BW = 4.0 * q.dp;
        double half_in[3] = {0.5 * q.DL, 0.5 * q.DH, 0.5 * q.DW};
        double half_out[3] = {0.5 * q.DL + BW, 0.5 * q.DH + BW, q.dim == 3 ? 0.5 * q.DW + BW : 0.0};
        add<GeometricShapeBox>(half_in, half_out);
        subtract<GeometricShapeBox>(half_in, half_in);
    }
};

template <class T> struct TypeTag { using type = T; };

// Slab of the static wall (SURVEY §8e / DESIGN §6): a rank of a decomposed run stores only the wall particles of the cell
// planes its own fluid can reach — [first own plane - depth - margin, last own plane + depth + margin] on the wall's mesh —
// instead of a full copy of the wall. The master copy (positions, normals, planes; the reference's particle order) stays on
// the host; when re-cuts move the slab beyond the stored planes the subset is loaded again and the wall's cell list is
// rebuilt (the wall is static: nothing else ever changes there, dambreak.cpp:159 builds its list once too).
// Wall particles keep their GLOBAL numbers as ReferenceID, so in-cell order, neighbour rows and the CSR export are those of
// the undecomposed run; the number of wall particles in the planes below the stored ones is the wall's slot origin
// (RelationBase::view: bank-aligned contact rows depend on s_global - t_global only).
class WallSlab
{
    SolidBody &wall_;
    std::vector<Vecd> pos_, normal_;
    std::vector<int> plane_;
    std::vector<uint64_t> below_; // below_[x] = wall particles in planes < x
    int planes_ = 0, depth_ = 1, margin_ = 4;
    int lo_ = 0, hi_ = -1; // stored planes [lo_, hi_]
    uint64_t loads_ = 0;

  public:
    // `m`: the mesh of the wall's cell-linked list (handed in: the list itself is made once the storage is sized)
    WallSlab(SolidBody &wall, const sphb200_mesh_t &m, std::vector<Vecd> positions, std::vector<Vecd> normals, int search_depth)
        : wall_(wall), pos_(std::move(positions)), normal_(std::move(normals)), depth_(search_depth)
    {
        planes_ = m.cells[0];
        plane_.resize(pos_.size());
        below_.assign(planes_ + 1, 0);
        for (size_t i = 0; i < pos_.size(); ++i)
        {
            plane_[i] = hostCellCoordinate(pos_[i].x, m.lower[0], m.spacing, m.cells[0]);
            below_[plane_[i] + 1]++;
        }
        for (int x = 0; x < planes_; ++x) below_[x + 1] += below_[x];
    }
    size_t globalParticles() const { return pos_.size(); }
    uint64_t loads() const { return loads_; }
    // largest number of wall particles any window of `width` planes holds (sizes the storage once)
    size_t largestWindow(int width) const
    {
        uint64_t best = 0;
        for (int x = 0; x < planes_; ++x) best = std::max(best, below_[std::min(planes_, x + width)] - below_[x]);
        return (size_t)best;
    }
    int marginPlanes() const { return margin_; }
    // wall particles the planes around the fluid planes [X0, X1) hold, margins included
    size_t particlesAround(int X0, int X1) const
    {
        const int lo = std::max(0, X0 - depth_ - margin_), hi = std::min(planes_ - 1, X1 - 1 + depth_ + margin_);
        return (size_t)(below_[hi + 1] - below_[lo]);
    }
    // Pure planning (host arithmetic only): the planes to store for the fluid planes [X0, X1) given the planes stored now
    // ([lo, hi], empty if hi < lo) and the storage `bound`. Returns false if what is stored suffices.
    bool plan(int X0, int X1, size_t bound, int &lo, int &hi) const
    {
        return planWallPlanes(below_.data(), planes_, depth_, margin_, X0, X1, bound, lo, hi);
    }
    // the same on a plain histogram: below[x] = wall particles in planes < x, x = 0 .. planes (tests/test_host_logic_cpu.py)
    static bool planWallPlanes(const uint64_t *below, int planes, int depth, int margin, int X0, int X1, size_t bound, int &lo, int &hi)
    {
        const int need_lo = std::max(0, X0 - depth), need_hi = std::min(planes - 1, X1 - 1 + depth);
        if (hi >= lo && need_lo >= lo && need_hi <= hi) return false;
        lo = std::max(0, need_lo - margin);
        hi = std::min(planes - 1, need_hi + margin);
        // the margin only saves reloads: give it up where the storage is too small for it
        while ((size_t)(below[hi + 1] - below[lo]) > bound && (lo < need_lo || hi > need_hi))
        {
            if (lo < need_lo) ++lo;
            if (hi > need_hi) --hi;
        }
        return true;
    }
    uint64_t particlesBelow(int plane) const { return below_[std::min(std::max(plane, 0), planes_)]; }
    // make sure the wall planes around the fluid planes [X0, X1) are stored; true if the subset was (re)loaded
    bool ensure(int X0, int X1)
    {
        if (!plan(X0, X1, wall_.getBaseParticles().ParticlesBound(), lo_, hi_)) return false;
        std::vector<Vecd> p, nrm;
        std::vector<UnsignedInt> ids;
        const size_t count = (size_t)(below_[hi_ + 1] - below_[lo_]);
        p.reserve(count), nrm.reserve(count), ids.reserve(count);
        for (size_t i = 0; i < pos_.size(); ++i)
            if (plane_[i] >= lo_ && plane_[i] <= hi_)
            {
                p.push_back(pos_[i]);
                nrm.push_back(normal_[i]);
                ids.push_back((UnsignedInt)i);
            }
        wall_.loadWallSubset(p, nrm, ids);
        wall_.setSlotOrigin((uint32_t)(below_[lo_] & 0xffffffffull));
        ++loads_;
        return true;
    }
};

class DamBreakCK
{
  public:
    using P = MainExecutionPolicy;
    DamBreakParameters q_;
    Real U_f_, c_f_;
    SPHSystem sph_system;
    FluidBody water_block;
    SolidBody wall_boundary;
    std::unique_ptr<Inner<>> water_block_inner;
    std::unique_ptr<Contact<>> water_wall_contact;
    std::unique_ptr<UpdateCellLinkedList<P, RealBody>> water_cell_linked_list, wall_cell_linked_list;
    std::unique_ptr<UpdateRelation<P, Inner<>, Contact<>>> water_block_update_complex_relation;
    std::unique_ptr<ParticleSortCK<P>> particle_sort;
    Gravity gravity;
    std::unique_ptr<StateDynamics<P, GravityForceCK<Gravity>>> constant_gravity;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::AdvectionStepSetup>> water_advection_step_setup;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::UpdateParticlePosition>> water_update_particle_position;
    std::unique_ptr<InteractionDynamicsCK<P, LinearCorrectionMatrixComplex>> fluid_linear_correction_matrix;
    std::unique_ptr<InteractionDynamicsCK<P, fluid_dynamics::FreeSurfaceIndicationComplexSpatialTemporalCK>> fluid_boundary_indicator;
    std::unique_ptr<InteractionDynamicsBase> fluid_viscous_force, kernel_gradient_integral;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::TransportVelocityCorrectionCK<SPHBody, TruncatedLinear>>> transport_correction;
    std::unique_ptr<ObserverBody> fluid_observer;
    std::unique_ptr<Contact<>> fluid_observer_contact;
    std::unique_ptr<UpdateRelation<P, Contact<>>> fluid_observer_contact_relation;
    std::unique_ptr<ObservedQuantityRecording<P, Real>> fluid_observer_pressure;
    std::unique_ptr<BodyStatesRecordingToVtpCK<P>> body_states_recording; // made on demand (recordStates)
    std::unique_ptr<InteractionDynamicsBase> fluid_acoustic_step_1st_half, fluid_acoustic_step_2nd_half;
    std::unique_ptr<InteractionDynamicsBase> fluid_density_summation;
    std::unique_ptr<StateDynamics<P, fluid_dynamics::DensityRegularization<SPHBody, WeaklyCompressibleFluid, FreeSurface>>> fluid_density_regularization;
    std::unique_ptr<BaseDynamics<Real>> fluid_advection_time_step_owner, fluid_acoustic_time_step_owner;
    fluid_dynamics::AcousticTimeStepBase *fluid_acoustic_time_step = nullptr; // exec(), setPrimed(), ReducedValue()
    BaseDynamics<Real> *fluid_advection_time_step = nullptr;
    std::function<Real()> advection_reduced_value;
    // legacy-API objects (q.legacy)
    std::unique_ptr<ComplexRelation> water_wall_complex;
    fluid_dynamics::FluidDynamicsBase *cuts_adv_ = nullptr;
    std::unique_ptr<ReduceDynamicsCK<P, TotalMechanicalEnergyCK>> record_water_mechanical_energy;
    std::unique_ptr<SlabDecomposition> decomposition; // nranks > 1 only
    std::unique_ptr<WallSlab> wall_slab;              // nranks > 1 only: this rank's part of the static wall
    fluid_dynamics::AcousticStep1stHalfPhases *first_half_phases_ = nullptr;
    fluid_dynamics::AcousticStep2ndHalfPhases *second_half_phases_ = nullptr;
    SingleVariable<Real> *sv_physical_time = nullptr;
    size_t number_of_iterations = 0, acoustic_steps = 0;
    double physical_time = 0; // accumulated in double for reporting; the SingleVariable keeps the Real copy
    Real last_acoustic_dt = 0, last_advection_dt = 0;

    // createObservationPoints(), dambreak.cpp:54-65 (3-D) and Dambreak.cpp:27-28 (2-D)
    static std::vector<Vecd> createObservationPoints(const DamBreakParameters &q)
    {
        std::vector<Vecd> pts;
        if (q.dim == 2)
        {
            pts.push_back(Vecd(Real(q.DL), Real(0.2), 0));
            return pts;
        }
        for (double y : {0.01, 0.1, 0.2, 0.24, 0.252, 0.266}) pts.push_back(Vecd(Real(q.DL), Real(y), Real(0.5 * q.DW)));
        return pts;
    }
    static BoundingBoxd caseBounds(const DamBreakParameters &q)
    {
        double BW = 4.0 * q.dp;
        return BoundingBoxd(Vecd(Real(-BW), Real(-BW), q.dim == 3 ? Real(-BW) : Real(0)),
                            Vecd(Real(q.DL + BW), Real(q.DH + BW), q.dim == 3 ? Real(q.DW + BW) : Real(0)));
    }

    // positions == nullptr: generate the lattice here (generateParticles<BaseParticles, Lattice>()); otherwise use the
    // arrays handed over (packed xyz, reference order), e.g. produced by the same rule on the harness side.
    explicit DamBreakCK(const DamBreakParameters &q, const std::vector<Vecd> *fluid_positions = nullptr,
                        const std::vector<Vecd> *wall_positions = nullptr, const std::vector<Vecd> *wall_normals = nullptr,
                        const BoundingBoxd *exact_system_bounds = nullptr)
        : q_(q), U_f_(Real(2.0 * std::sqrt(q.gravity_g * q.LH))), c_f_(Real(10.0) * U_f_),
          sph_system(caseBounds(q), Real(q.dp), q.dim),
          water_block(sph_system, makeShared<WaterBlock>("WaterBody", q)),
          wall_boundary(sph_system, makeShared<WallBoundary>("WallBoundary", q)),
          gravity(Vecd(0, Real(-q.gravity_g), 0))
    {
        if (q.kernel_kind == 1)
        {
            if (q.legacy) throw SphError("legacy formulation: only the Wendland C2 kernel");
            water_block.getSPHAdaptation().resetKernel<KernelLaguerreGauss>();
            wall_boundary.getSPHAdaptation().resetKernel<KernelLaguerreGauss>();
        }
        else if (q.kernel_kind != 0) throw SphError("kernel_kind: 0 (Wendland C2) or 1 (Laguerre-Gauss)");
        {
            // system bounds = case bounds + 4 dp (sph_system.cpp:39), evaluated in double from the case file's double
            // literals and rounded once to Real
            double BW = 4.0 * q.dp, ext[3] = {q.DL, q.DH, q.DW};
            BoundingBoxd sb;
            for (int d = 0; d < q.dim; ++d)
            {
                sb.lower_[d] = Real(-BW - 4.0 * q.dp);
                sb.upper_[d] = Real(ext[d] + BW + 4.0 * q.dp);
            }
            sph_system.setSystemDomainBoundsExact(exact_system_bounds ? *exact_system_bounds : sb);
        }
        Real vol = Real(std::pow(Real(q.dp), Real(q.dim)));
        if (q.mu_f > 0) water_block.defineClosure<WeaklyCompressibleFluid, Viscosity>(Real(q.rho0_f), c_f_, Real(q.mu_f));
        else water_block.defineMatterMaterial<WeaklyCompressibleFluid>(Real(q.rho0_f), c_f_);
        std::vector<int> cuts;
        if (q.nranks > 1)
        {
            // every rank generates the global lattice, keeps the particles of its own cell planes and their global numbers
            std::vector<Vecd> all = fluid_positions ? *fluid_positions
                                                    : generateLattice(water_block.getInitialShape(), sph_system.system_domain_bounds_, Real(q.dp), q.dim);
            SPHAdaptation &ad = water_block.getSPHAdaptation();
            sphb200_mesh_t mesh = makeMesh(sph_system.system_domain_bounds_, ad.CutOffRadius(), 2, q.dim);
            std::vector<uint64_t> per_plane(mesh.cells[0], 0);
            std::vector<int> plane(all.size());
            for (size_t i = 0; i < all.size(); ++i)
            {
                plane[i] = hostCellCoordinate(all[i].x, mesh.lower[0], mesh.spacing, mesh.cells[0]);
                per_plane[plane[i]]++;
            }
            cuts = planSlabCuts(per_plane, q.nranks);
            if (q.initial_cut_shift) // deliberately unbalanced start (tests of SlabDecomposition::recut)
            {
                std::vector<int> skewed = cuts;
                for (int r = 1; r < q.nranks; ++r) skewed[r] += q.initial_cut_shift;
                cuts = limitCutMoves(cuts, skewed);
            }
            std::vector<Vecd> own;
            std::vector<UnsignedInt> ids;
            uint64_t max_plane = 0;
            for (uint64_t c : per_plane) max_plane = std::max(max_plane, c);
            for (size_t i = 0; i < all.size(); ++i)
                if (plane[i] >= cuts[q.rank] && plane[i] < cuts[q.rank + 1])
                {
                    own.push_back(all[i]);
                    ids.push_back((UnsignedInt)i);
                }
            // room for what re-cuts may bring: a rank that starts small (skewed cuts) ends with its fair share of the fluid
            const size_t share = std::max(own.size(), all.size() / (size_t)q.nranks + 1);
            size_t bound = share + share / 4 + 8 * (size_t)max_plane + 4096;
            water_block.generateParticlesFromPositions(own, vol, bound, &ids);
        }
        else if (fluid_positions) water_block.generateParticlesFromPositions(*fluid_positions, vol);
        else water_block.generateParticles<BaseParticles, Lattice>();
        wall_boundary.defineMatterMaterial<Solid>();
        if (q.nranks > 1 && q.wall_slabs)
        {
            // slab of the wall: master copy on the host, storage for the widest window this rank may ever hold
            std::vector<Vecd> all_wall = wall_positions ? *wall_positions
                                                        : generateLattice(wall_boundary.getInitialShape(), sph_system.system_domain_bounds_, Real(q.dp), q.dim);
            std::vector<Vecd> all_normals = wall_normals ? *wall_normals : wall_boundary.normalsFromBodyShape(all_wall);
            const sphb200_mesh_t wmesh = makeMesh(sph_system.system_domain_bounds_, wall_boundary.getSPHAdaptation().CutOffRadius(), 2, q.dim);
            // every rank owns about planes / nranks planes; re-cuts may widen a slab: room for twice that plus the margins
            const int width = std::min(wmesh.cells[0], 2 * (wmesh.cells[0] / q.nranks + 1) + 2 * (1 + 4) + 2);
            wall_slab.reset(new WallSlab(wall_boundary, wmesh, std::move(all_wall), std::move(all_normals), 1));
            // ... and never less than a quarter more than the first slab needs: slabs are cut by particle count, so a rank
            // downstream of the water column starts with a long, nearly empty range of planes (and most of the wall)
            const size_t first = wall_slab->particlesAround(cuts[q.rank], cuts[q.rank + 1]);
            const size_t room = std::min(wall_slab->globalParticles(), std::max(first + first / 4, wall_slab->largestWindow(width))) + 1024;
            std::vector<Vecd> none;
            wall_boundary.generateParticlesFromPositions(none, vol, room); // empty storage of that size; WallSlab::ensure fills it
            wall_boundary.registerWallVariables(nullptr);
            wall_slab->ensure(cuts[q.rank], cuts[q.rank + 1]);
        }
        else
        {
            if (wall_positions) wall_boundary.generateParticlesFromPositions(*wall_positions, vol);
            else wall_boundary.generateParticles<BaseParticles, Lattice>();
            if (wall_normals) wall_boundary.registerWallVariables(wall_normals);
            else wall_boundary.computeNormalFromBodyShape(); // NormalFromBodyShapeCK, run on the host (dambreak.cpp:119,153)
        }

        using namespace fluid_dynamics;
        water_cell_linked_list.reset(new UpdateCellLinkedList<P, RealBody>(water_block));
        wall_cell_linked_list.reset(new UpdateCellLinkedList<P, RealBody>(wall_boundary));
        particle_sort.reset(new ParticleSortCK<P>(water_block));
        constant_gravity.reset(new StateDynamics<P, GravityForceCK<Gravity>>(water_block, gravity));
        if (q.legacy)
        {
            // Dambreak.cpp:97-116 — the legacy class names over the same kernels (legacy_dynamics.h)
            if (q.correction || q.nranks > 1) throw SphError("legacy formulation: no kernel correction, no decomposition");
            auto *inner = new InnerRelation(water_block);
            auto *contact = new ContactRelation(water_block, {&wall_boundary});
            water_block_inner.reset(inner);
            water_wall_contact.reset(contact);
            water_wall_complex.reset(new ComplexRelation(*inner, *contact));
            fluid_acoustic_step_1st_half.reset(new Dynamics1Level<Integration1stHalfWithWallRiemann>(*inner, *contact));
            auto *second = new Dynamics1Level<Integration2ndHalfWithWallRiemann>(*inner, *contact);
            fluid_acoustic_step_2nd_half.reset(second);
            fluid_density_summation.reset(new InteractionWithUpdate<DensitySummationComplexFreeSurface>(*inner, *contact));
            auto *adv = new ReduceDynamics<AdvectionViscousTimeStep>(water_block, U_f_);
            fluid_advection_time_step_owner.reset(adv);
            fluid_advection_time_step = adv;
            advection_reduced_value = [adv]() { return adv->ReducedValue(); };
            auto *ac = new ReduceDynamics<AcousticTimeStep>(water_block);
            fluid_acoustic_time_step_owner.reset(ac);
            fluid_acoustic_time_step = ac;
            if (q.fused_time_step) second->fuseTimeStepReduction(*ac);
        }
        else
        {
            water_block_inner.reset(new Inner<>(water_block));
            water_wall_contact.reset(new Contact<>(water_block, {&wall_boundary}));
            water_advection_step_setup.reset(new StateDynamics<P, AdvectionStepSetup>(water_block));
            water_update_particle_position.reset(new StateDynamics<P, UpdateParticlePosition>(water_block));
            auto *ac = new ReduceDynamicsCK<P, AcousticTimeStepCK<WeaklyCompressibleFluid>>(water_block);
            fluid_acoustic_time_step_owner.reset(ac);
            fluid_acoustic_time_step = ac;
            if (q.correction)
            {
                if (q.riemann != 1) throw SphError("correction variants: the reference has aliases for the acoustic Riemann solver only");
                fluid_linear_correction_matrix.reset(new InteractionDynamicsCK<P, LinearCorrectionMatrixComplex>(
                    DynamicsArgs(*water_block_inner, 0.5), *water_wall_contact));
                fluid_acoustic_step_1st_half.reset(new InteractionDynamicsCK<P, AcousticStep1stHalfWithWallRiemannCorrectionCK>(*water_block_inner, *water_wall_contact));
                auto *a2 = new InteractionDynamicsCK<P, AcousticStep2ndHalfWithWallRiemannCorrectionCK>(*water_block_inner, *water_wall_contact);
                fluid_acoustic_step_2nd_half.reset(a2);
                if (q.fused_time_step) a2->fuseTimeStepReduction(*ac);
            }
            else
            {
                // the closed set of Riemann variants of SURVEY §8 a14 (aliases acoustic_step_1st_half.h:196-201, 2nd_half.h)
                auto make = [&](auto first, auto second) {
                    using First = typename decltype(first)::type;
                    using Second = typename decltype(second)::type;
                    fluid_acoustic_step_1st_half.reset(new InteractionDynamicsCK<P, First>(*water_block_inner, *water_wall_contact));
                    auto *a2 = new InteractionDynamicsCK<P, Second>(*water_block_inner, *water_wall_contact);
                    fluid_acoustic_step_2nd_half.reset(a2);
                    if (q.fused_time_step) a2->fuseTimeStepReduction(*ac);
                };
                if (q.riemann == 1) make(TypeTag<AcousticStep1stHalfWithWallRiemannCK>{}, TypeTag<AcousticStep2ndHalfWithWallRiemannCK>{});
                else if (q.riemann == 0) make(TypeTag<AcousticStep1stHalfWithWallNoRiemannCK>{}, TypeTag<AcousticStep2ndHalfWithWallNoRiemannCK>{});
                else if (q.riemann == 2) make(TypeTag<AcousticStep1stHalfWithWallDissipativeRiemannCK>{}, TypeTag<AcousticStep2ndHalfWithWallDissipativeRiemannCK>{});
                else throw SphError("riemann: 0 (NoRiemann), 1 (Acoustic) or 2 (Dissipative)");
            }
            auto *sum = new InteractionDynamicsCK<P, CompressionSummation<Inner<>, Contact<>>>(*water_block_inner, *water_wall_contact);
            fluid_density_summation.reset(sum);
            fluid_density_regularization.reset(new StateDynamics<P, DensityRegularization<SPHBody, WeaklyCompressibleFluid, FreeSurface>>(water_block));
            if (q.fused_regularization) sum->addPostStateDynamics(*fluid_density_regularization);
            auto *adv = new ReduceDynamicsCK<P, AdvectionTimeStepCK>(water_block, U_f_);
            fluid_advection_time_step_owner.reset(adv);
            fluid_advection_time_step = adv;
            advection_reduced_value = [adv]() { return adv->ReducedValue(); };
            cuts_adv_ = adv; // gets the decomposition below, once it exists
            if (q.mu_f > 0 || q.transport_velocity)
            {
                if (q.nranks > 1) throw SphError("slab decomposition: viscous force / transport velocity are not decomposed yet");
                if (q.mu_f > 0)
                {
                    if (q.correction)
                        fluid_viscous_force.reset(new InteractionDynamicsCK<P, ViscousForceCK<Inner<WithUpdate, Viscosity, LinearCorrectionCK>, Contact<Wall, Viscosity, LinearCorrectionCK>>>(*water_block_inner, *water_wall_contact));
                    else
                        fluid_viscous_force.reset(new InteractionDynamicsCK<P, ViscousForceWithWallCK>(*water_block_inner, *water_wall_contact));
                }
                if (q.transport_velocity)
                {
                    if (q.correction)
                        kernel_gradient_integral.reset(new InteractionDynamicsCK<P, KernelGradientIntegralCorrectedComplex>(*water_block_inner, *water_wall_contact));
                    else
                        kernel_gradient_integral.reset(new InteractionDynamicsCK<P, KernelGradientIntegralComplex>(*water_block_inner, *water_wall_contact));
                    transport_correction.reset(new StateDynamics<P, TransportVelocityCorrectionCK<SPHBody, TruncatedLinear>>(water_block));
                }
            }
            if (q.surface_indicator)
            {
                fluid_boundary_indicator.reset(new InteractionDynamicsCK<P, FreeSurfaceIndicationComplexSpatialTemporalCK>(*water_block_inner, *water_wall_contact));
            }
        }
        if (q.observers)
        {
            // dambreak.cpp:87-88,128,140-141 (CK); Dambreak.cpp:78-79,113 and test_3d_dambreak/dambreak.cpp:84-85,127 (first-generation
            // API: ObservedQuantityRecording<Real>("Pressure", fluid_observer_contact), the same interpolation)
            registerPressure();
            fluid_observer.reset(new ObserverBody(sph_system, "FluidObserver"));
            fluid_observer->generateParticles<ObserverParticles>(createObservationPoints(q));
            fluid_observer_contact.reset(new Contact<>(*fluid_observer, {&water_block}));
            fluid_observer_contact_relation.reset(new UpdateRelation<P, Contact<>>(*fluid_observer_contact));
            fluid_observer_pressure.reset(new ObservedQuantityRecording<P, Real>(*fluid_observer_contact, "Pressure"));
        }
        if (wall_slab && water_wall_contact->search_depth_ != 1) throw SphError("WallSlab: the contact search reaches further than one cell plane");
        water_block_update_complex_relation.reset(new UpdateRelation<P, Inner<>, Contact<>>(*water_block_inner, *water_wall_contact));
        record_water_mechanical_energy.reset(new ReduceDynamicsCK<P, TotalMechanicalEnergyCK>(water_block, gravity));
        sv_physical_time = sph_system.getSystemVariableByName<Real>("PhysicalTime");
        first_half_phases_ = dynamic_cast<fluid_dynamics::AcousticStep1stHalfPhases *>(fluid_acoustic_step_1st_half.get());
        second_half_phases_ = dynamic_cast<fluid_dynamics::AcousticStep2ndHalfPhases *>(fluid_acoustic_step_2nd_half.get());
        if (q.nranks > 1)
        {
            decomposition.reset(new SlabDecomposition(water_block, q.rank, q.nranks, cuts));
            cuts_adv_->setDecomposition(decomposition.get());
            fluid_acoustic_time_step->setDecomposition(decomposition.get());
            record_water_mechanical_energy->setDecomposition(decomposition.get());
            if (fluid_boundary_indicator) fluid_boundary_indicator->setDecomposition(decomposition.get());
            if (fluid_observer_pressure) fluid_observer_pressure->setDecomposition(decomposition.get());
        }
    }

    void registerPressure() { water_block.getBaseParticles().registerStateVariable<Real>("Pressure"); }

    // dambreak.cpp:141-145,177,230: the body-state output with its write list; every call synchronises Position and the
    // listed variables device -> host (BodyStatesRecordingToVtpCK::prepareToWrite) and writes one .vtp per body
    void recordStates(const std::string &folder)
    {
        if (!body_states_recording)
        {
            body_states_recording.reset(new BodyStatesRecordingToVtpCK<P>(sph_system, folder));
            body_states_recording->addToWrite<Vecd>(wall_boundary, "NormalDirection");
            body_states_recording->addToWrite<Real>(water_block, "Density");
            if (fluid_boundary_indicator)
            {
                body_states_recording->addToWrite<int>(water_block, "Indicator");
                body_states_recording->addToWrite<Real>(water_block, "PositionDivergence");
            }
        }
        body_states_recording->writeToFile(number_of_iterations);
    }

    // dambreak.cpp:152-160
    void initialize()
    {
        constant_gravity->exec();
        if (decomposition) decomposition->rebuild();
        else water_cell_linked_list->exec();
        wall_cell_linked_list->exec();
        if (q_.legacy) water_wall_complex->updateConfiguration();
        else water_block_update_complex_relation->exec();
        if (fluid_observer_contact_relation)
        {
            fluid_observer_contact_relation->exec();
            fluid_observer_pressure->writeToFile(number_of_iterations); // first output before the main loop, dambreak.cpp:177
        }
        fluid_acoustic_time_step->setPrimed(false);
    }

    // one advection step of the legacy case file, Dambreak.cpp:166-215
    int stepOuterLegacy()
    {
        Real advection_dt = fluid_advection_time_step->exec();
        fluid_density_summation->exec(); // fluid_density_by_summation
        Real relaxation_time = 0, acoustic_dt = 0;
        int n_inner = 0;
        while (relaxation_time < advection_dt)
        {
            acoustic_dt = fluid_acoustic_time_step->exec();
            fluid_acoustic_step_1st_half->exec(acoustic_dt); // fluid_pressure_relaxation
            fluid_acoustic_step_2nd_half->exec(acoustic_dt); // fluid_density_relaxation
            relaxation_time += acoustic_dt;
            physical_time += acoustic_dt;
            sv_physical_time->incrementValue(acoustic_dt);
            ++n_inner;
        }
        acoustic_steps += n_inner;
        number_of_iterations++;
        if (q_.sort_interval > 0 && number_of_iterations % q_.sort_interval == 0 && number_of_iterations != 1)
        {
            particle_sort->exec(); // particle_sorting
            fluid_acoustic_time_step->setPrimed(false);
        }
        water_cell_linked_list->exec();           // water_block.updateCellLinkedList()
        water_wall_complex->updateConfiguration(); // neighbour lists + frozen pair geometry
        if (fluid_observer_contact_relation)
        {
            // fluid_observer_contact.updateConfiguration(); write_recorded_water_pressure.writeToFile(), test_3d_dambreak/dambreak.cpp:193-194
            // (every iteration; the 2-D case file keeps every 200th of these records, Dambreak.cpp:175-180)
            fluid_observer_contact_relation->exec();
            fluid_observer_pressure->writeToFile(number_of_iterations);
        }
        last_acoustic_dt = acoustic_dt;
        last_advection_dt = advection_dt;
        return n_inner;
    }

    // One acoustic step of a decomposed run with the plane exchange hidden behind interior compute.
    // Default stream M: initialize (all own slots), then the INTERIOR launches of both halves. High-priority side
    // stream S: exchange of Pressure, 1st half on the boundary planes, exchange of the velocity records, 2nd half on
    // the boundary planes. Boundary launches run first on S because the neighbour rank waits for their results; they
    // and the NCCL kernels slot into the SMs the interior launch frees, so neither stream idles the GPU.
    //   event 0  M -> S  initialize done (the boundary 1st half reads every neighbour's Pressure)
    //   event 1  S -> M  1st half done on the boundary planes (the interior 2nd half reads their velocities)
    //   event 2  M -> S  1st half done on the interior (the boundary 2nd half reads its velocities)
    //   event 3  S -> M  2nd half done on the boundary planes: the step is complete on M
    void acousticStepOverlapped(Real dt)
    {
        using SlotRange = SlabDecomposition::SlotRange;
        SlabDecomposition &d = *decomposition;
        BaseParticles &p = water_block.getBaseParticles();
        void *M = execution_instance().stream(), *S = d.sideStream();
        const SlotRange own{d.ownBegin(), d.ownBegin() + d.ownParticles()}, in = d.interior();
        const SlotRange edge[2] = {d.leftBoundary(), d.rightBoundary()};
        auto on = [&](const SlotRange &r, auto &&launch) {
            if (r.empty()) return;
            p.setActiveRange(r.begin, r.end);
            launch();
        };
        second_half_phases_->primeFusedReduction();
        first_half_phases_->deviceInitialize(dt);
        d.signal(0, M);
        {
            StreamScope side(S);
            d.await(0, S);
            if (q_.correction) d.refreshGhosts({"Pressure", "LinearCorrectionRecord"}); // the correction variants read p_j from the record
            else d.refreshGhosts({"Pressure"});
            for (const SlotRange &r : edge) on(r, [&] { first_half_phases_->deviceInteractAndUpdate(dt); });
            d.signal(1, S);
            d.refreshGhosts({"PosVolVel"}); // the velocity records of the boundary planes are final: send them now
        }
        on(in, [&] { first_half_phases_->deviceInteractAndUpdate(dt); });
        d.signal(2, M);
        d.await(1, M);
        {
            StreamScope side(S);
            d.await(2, S);
            for (const SlotRange &r : edge) on(r, [&] { second_half_phases_->deviceLaunch(dt); });
            d.signal(3, S);
        }
        on(in, [&] { second_half_phases_->deviceLaunch(dt); });
        d.await(3, M);
        p.setActiveRange(own.begin, own.end);
    }

    // dambreak.cpp:216-224: particle sort (every sort_interval steps), cell-linked list, relations, observers
    void updateConfiguration(bool allow_sort)
    {
        // decomposed runs keep the initial global numbering (ParticleSortCK only renumbers: storage is cell ordered anyway)
        if (allow_sort && !decomposition && q_.sort_interval > 0 && number_of_iterations % q_.sort_interval == 0 && number_of_iterations != 1)
        {
            SPHCK_STAGE("particle sort", particle_sort->exec());
            fluid_acoustic_time_step->setPrimed(false); // Force/ForcePrior pairing changed (see ParticleSortCK)
        }
        if (decomposition)
        {
            // migration + ghost planes + cell-linked list; at the sort cadence the slabs are re-balanced as well
            const bool recut = allow_sort && q_.recut_interval > 0 && number_of_iterations % q_.recut_interval == 0 && number_of_iterations != 1;
            if (recut)
            {
                SPHCK_STAGE("slab re-cut", decomposition->recut());
                // the cuts moved: load the wall planes around the new slab if they are not stored yet, and list them
                if (wall_slab && wall_slab->ensure(decomposition->cuts()[q_.rank], decomposition->cuts()[q_.rank + 1])) wall_cell_linked_list->exec();
            }
            else SPHCK_STAGE("slab rebuild (migration, ghost planes, cell list)", decomposition->update());
        }
        else SPHCK_STAGE("cell list", water_cell_linked_list->exec());
        SPHCK_STAGE("relations", water_block_update_complex_relation->exec());
        if (fluid_observer_contact_relation)
        {
            fluid_observer_contact_relation->exec();
            fluid_observer_pressure->writeToFile(number_of_iterations);
        }
    }

    // Where the configuration update of an advection step runs. AfterDynamics is the reference loop (dambreak.cpp:188-224).
    // BeforeDynamics serves callers that hand the particle state in from HOST buffers every step (bench.py e2e,
    // HostTransferPipeline): the lists must be built for the state that was just uploaded, and building them again at
    // the end of the step for a state that is about to be overwritten would be wasted work. Same launches per step.
    enum class ConfigurationUpdate { AfterDynamics, BeforeDynamics };
    ConfigurationUpdate configuration_update = ConfigurationUpdate::AfterDynamics;

    // one advection step, dambreak.cpp:188-222; returns the number of acoustic sub-steps taken
    int stepOuter()
    {
        if (q_.legacy) return stepOuterLegacy();
        if (configuration_update == ConfigurationUpdate::BeforeDynamics)
        {
            updateConfiguration(false);
            fluid_acoustic_time_step->setPrimed(false); // the state came from outside: no fused reduction to reuse
        }
        SPHCK_STAGE("density summation", fluid_density_summation->exec());
        if (!q_.fused_regularization) fluid_density_regularization->exec();
        SPHCK_STAGE("advection setup", water_advection_step_setup->exec());
        if (decomposition) SPHCK_STAGE("ghost refresh (volume)", decomposition->refreshGhosts({"VolumetricMeasure"})); // neighbours read V_j of ghost particles
        if (fluid_viscous_force) fluid_viscous_force->exec(); // lid_driven_cavity_sycl.cpp:268-276
        if (kernel_gradient_integral)
        {
            kernel_gradient_integral->exec();
            transport_correction->exec();
        }
        Real advection_dt = 0;
        SPHCK_STAGE("advection dt", advection_dt = fluid_advection_time_step->exec());
        if (fluid_boundary_indicator) SPHCK_STAGE("free-surface indication", fluid_boundary_indicator->exec());
        if (q_.correction)
        {
            SPHCK_STAGE("linear correction matrix", fluid_linear_correction_matrix->exec());
            // both half steps read B of the neighbours: the one refresh the correction variants add to a decomposed step
            // (tests/test_decomposed_oracle_cpu.py::test_dam_break_correction_variants_bit_identical)
            if (decomposition) decomposition->refreshGhosts({"LinearCorrectionMatrix", "LinearCorrectionRecord"});
        }
        Real relaxation_time = 0, acoustic_dt = 0;
        int n_inner = 0;
        while (relaxation_time < advection_dt)
        {
            SPHCK_STAGE("acoustic dt", acoustic_dt = fluid_acoustic_time_step->exec()); // global max when decomposed
            if (decomposition && q_.overlap_exchange)
                SPHCK_STAGE("acoustic step (overlapped exchange)", acousticStepOverlapped(acoustic_dt));
            else if (decomposition)
            {
                // the two neighbour-read variables of the half steps are refreshed on the ghost planes in between
                first_half_phases_->deviceInitialize(acoustic_dt);
                if (q_.correction) decomposition->refreshGhosts({"Pressure", "LinearCorrectionRecord"});
                else decomposition->refreshGhosts({"Pressure"});
                first_half_phases_->deviceInteractAndUpdate(acoustic_dt);
                decomposition->refreshGhosts({"PosVolVel"}); // the 2nd half reads neighbour velocities from the gather record
                fluid_acoustic_step_2nd_half->exec(acoustic_dt);
            }
            else
            {
                SPHCK_STAGE("1st half", fluid_acoustic_step_1st_half->exec(acoustic_dt));
                SPHCK_STAGE("2nd half", fluid_acoustic_step_2nd_half->exec(acoustic_dt));
            }
            relaxation_time += acoustic_dt;
            physical_time += acoustic_dt;
            sv_physical_time->incrementValue(acoustic_dt);
            ++n_inner;
        }
        acoustic_steps += n_inner;
        SPHCK_STAGE("update position", water_update_particle_position->exec());
        number_of_iterations++;
        if (configuration_update == ConfigurationUpdate::AfterDynamics) updateConfiguration(true);
        last_acoustic_dt = acoustic_dt;
        last_advection_dt = advection_dt;
        return n_inner;
    }
};
} // namespace SPH
#endif
