// sphinxsys_ck/legacy_dynamics.h — the first-generation ("legacy") API names of the same hot path: SimpleDynamics /
// InteractionWithUpdate / Dynamics1Level / ReduceDynamics, InnerRelation / ContactRelation / ComplexRelation,
// ParticleSorting, fluid_dynamics::Integration1stHalfWithWallRiemann / Integration2ndHalfWithWallRiemann /
// DensitySummationComplexFreeSurface / AcousticTimeStep / AdvectionViscousTimeStep.
//
// Same kernels as the CK path with material.formulation == 1 (include/sphb200.h): state Density/DensityChangeRate,
// analytic Wendland kernel, neighbour criterion |d|^2 < rc^2, positions advanced inside the half steps, and pair
// geometry (W_ij, dW_ij, e_ij, r_ij) FROZEN at the last updateConfiguration() — the reference stores it per pair in
// Neighborhood (particle_neighborhood/neighborhood.cpp:26-35,84-99); here the gather records packed at that moment
// play that role and the geometry is recomputed from them.
//
// Reference (relative to /root/reference/src/shared):
//   algorithms ................. particle_dynamics/dynamics_algorithms.h:100-353
//   relations .................. body_relations/{inner,contact,complex}_body_relation.h
//   fluid integration .......... particle_dynamics/fluid_dynamics/fluid_integration.hpp:49-231
//   density summation .......... particle_dynamics/fluid_dynamics/density_summation.cpp:8-22,58-78, .hpp:28-32
//   time steps ................. particle_dynamics/fluid_dynamics/fluid_time_step.cpp:11-59
//   lattice number density ..... adaptations/adaptation.cpp:26-60
//   case file .................. tests/2d_examples/test_2d_dambreak/Dambreak.cpp:97-220
#ifndef SPHINXSYS_CK_LEGACY_DYNAMICS_H
#define SPHINXSYS_CK_LEGACY_DYNAMICS_H

#include "fluid_dynamics.h"

namespace SPH
{
// SPHAdaptation::computeLatticeNumberDensity: sum of the analytic kernel over the lattice points inside the cut-off
inline Real computeLatticeNumberDensity(const SPHAdaptation &ad)
{
    const int dim = ad.dim_;
    const double h = ad.h_ref_, dp = ad.global_resolution_, rc = double(ad.kernel_.kernel_size) * h;
    const double pi = 3.14159265358979323846;
    const double sigma = ad.kernel_kind_ == 0 ? (dim == 2 ? 7.0 / (4.0 * pi) : 21.0 / (16.0 * pi)) : (dim == 2 ? 3.0 / pi : 8.0 / std::pow(pi, 1.5));
    const int depth = int(rc / dp) + 1;
    double sum = 0;
    for (int i = -depth; i <= depth; ++i)
        for (int j = -depth; j <= depth; ++j)
            for (int k = (dim == 3 ? -depth : 0); k <= (dim == 3 ? depth : 0); ++k)
            {
                double d = std::sqrt(double(i * i + j * j + k * k)) * dp;
                if (d < rc)
                {
                    double q = d / h;
                    double w1 = ad.kernel_kind_ == 0 ? std::pow(1.0 - 0.5 * q, 4) * (1.0 + 2.0 * q)
                                                     : (1.0 - q * q + std::pow(q, 4) / 6.0) * std::exp(-q * q);
                    sum += sigma / std::pow(h, dim) * w1;
                }
            }
    return Real(sum);
}

// ---- relations ----
class InnerRelation : public Inner<>
{
  public:
    explicit InnerRelation(SPHBody &body) : Inner<>(body) { legacy_criterion_ = true; }
};
class ContactRelation : public Contact<>
{
  public:
    ContactRelation(SPHBody &body, std::initializer_list<SPHBody *> contact_bodies) : Contact<>(body, contact_bodies) { legacy_criterion_ = true; }
};
class ComplexRelation
{
    InnerRelation &inner_;
    ContactRelation &contact_;
    UpdateRelation<MainExecutionPolicy, Inner<>, Contact<>> update_;

  public:
    ComplexRelation(InnerRelation &inner, ContactRelation &contact) : inner_(inner), contact_(contact), update_(inner, contact) {}
    InnerRelation &getInnerRelation() { return inner_; }
    ContactRelation &getContactRelation() { return contact_; }
    // rebuilds both neighbour lists; the pair geometry used until the next call is that of the positions NOW
    void updateConfiguration()
    {
        update_.exec();
        inner_.source_.setPosVolDirty();
        inner_.source_.refreshPosVol();
        contact_.target_.refreshPosVol();
    }
};
inline void updateCellLinkedList(SPHBody &body)
{
    UpdateCellLinkedList<MainExecutionPolicy, RealBody> u(body);
    u.exec();
}

// ---- algorithms (dynamics_algorithms.h); the execution policy defaults to the device ----
template <class LocalDynamicsType, class ExecutionPolicy = MainExecutionPolicy> using SimpleDynamics = StateDynamics<ExecutionPolicy, LocalDynamicsType>;
template <class LocalDynamicsType, class ExecutionPolicy = MainExecutionPolicy> using ReduceDynamics = ReduceDynamicsCK<ExecutionPolicy, LocalDynamicsType>;
template <class LocalDynamicsType, class ExecutionPolicy = MainExecutionPolicy>
using InteractionWithUpdate = InteractionDynamicsCK<ExecutionPolicy, LocalDynamicsType>;
template <class LocalDynamicsType, class ExecutionPolicy = MainExecutionPolicy> using Dynamics1Level = InteractionDynamicsCK<ExecutionPolicy, LocalDynamicsType>;
template <class LocalDynamicsType, class ExecutionPolicy = MainExecutionPolicy> using InteractionDynamics = InteractionDynamicsCK<ExecutionPolicy, LocalDynamicsType>;
using ParticleSorting = ParticleSortCK<MainExecutionPolicy>;
template <class GravityType> using GravityForce = GravityForceCK<GravityType>;

namespace fluid_dynamics
{
class LegacyFluidDynamics : public FluidDynamicsBase
{
  protected:
    // FluidIntegration constructor (fluid_integration.hpp:12-26) + the sortable set ParticleSorting swaps (all of them)
    void registerLegacyVariables()
    {
        formulation_ = 1;
        BaseParticles &p = particles_;
        p.registerStateVariable<Real>("Pressure");
        p.registerStateVariable<Real>("DensityChangeRate");
        p.registerStateVariable<Real>("DensitySummation");
        p.registerStateVariable<Vecd>("Velocity");
        p.registerStateVariable<Vecd>("Force");
        p.registerStateVariable<Vecd>("ForcePrior");
        p.addEvolvingVariable<Vecd>("Velocity");
        p.addEvolvingVariable<Real>("Mass");
        p.addEvolvingVariable<Vecd>("ForcePrior");
        p.addEvolvingVariable<Vecd>("Force");
        p.addEvolvingVariable<Real>("DensityChangeRate");
        p.addEvolvingVariable<Real>("Density");
        p.addEvolvingVariable<Real>("Pressure");
        sigma0_ = computeLatticeNumberDensity(sph_body_.getSPHAdaptation());
        if (contact_)
            if (auto *solid = dynamic_cast<Solid *>(&contact_->target_.getBaseMaterial())) wall_rho0_ = solid->rho0_;
    }

  public:
    explicit LegacyFluidDynamics(SPHBody &body) : FluidDynamicsBase(body) { registerLegacyVariables(); }
    LegacyFluidDynamics(RelationBase &inner, RelationBase *contact) : FluidDynamicsBase(inner, contact) { registerLegacyVariables(); }
};

// DensitySummation<Inner<FreeSurface>, Contact<>>
class DensitySummationComplexFreeSurface : public LegacyFluidDynamics
{
  public:
    DensitySummationComplexFreeSurface(InnerRelation &inner, ContactRelation &contact) : LegacyFluidDynamics(inner, &contact) { free_surface_ = 1; }
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_compression_summation, &a, 1, execution_instance().stream());
        return post;
    }
};
class DensitySummationComplex : public DensitySummationComplexFreeSurface
{
  public:
    DensitySummationComplex(InnerRelation &inner, ContactRelation &contact) : DensitySummationComplexFreeSurface(inner, contact) { free_surface_ = 0; }
};

template <class RiemannType> class Integration1stHalfWithWall : public LegacyFluidDynamics
{
  public:
    Integration1stHalfWithWall(InnerRelation &inner, ContactRelation &contact) : LegacyFluidDynamics(inner, &contact) { riemann_ = RiemannType::kind; }
    std::vector<BaseDynamics<void> *> deviceInteract(Real dt, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_acoustic_1st_half, &a, dt, execution_instance().stream());
        return post;
    }
};
template <class RiemannType> class Integration2ndHalfWithWall : public LegacyFluidDynamics
{
    AcousticTimeStepBase *fused_time_step_ = nullptr;

  public:
    Integration2ndHalfWithWall(InnerRelation &inner, ContactRelation &contact) : LegacyFluidDynamics(inner, &contact) { riemann_ = RiemannType::kind; }
    void fuseTimeStepReduction(AcousticTimeStepBase &time_step) { fused_time_step_ = &time_step; }
    std::vector<BaseDynamics<void> *> deviceInteract(Real dt, const std::vector<BaseDynamics<void> *> &post)
    {
        ExecutionInstance &ex = execution_instance();
        sphb200_fluid_args_t a = fluidArgs();
        float *slot = nullptr;
        Real h_min = sph_body_.getSPHAdaptation().MinimumSmoothingLength();
        if (fused_time_step_)
        {
            slot = fused_time_step_->fusedSlot();
            SPHCK_CALL(sphb200_fill_f32, slot, 0.0f, 1, ex.stream());
        }
        SPHCK_CALL(sphb200_acoustic_2nd_half, &a, dt, h_min, slot, ex.stream());
        if (fused_time_step_) fused_time_step_->setPrimed(true);
        return post;
    }
};
using Integration1stHalfWithWallRiemann = Integration1stHalfWithWall<AcousticRiemannSolverCK>;
using Integration2ndHalfWithWallRiemann = Integration2ndHalfWithWall<AcousticRiemannSolverCK>;
using Integration1stHalfWithWallNoRiemann = Integration1stHalfWithWall<NoRiemannSolverCK>;
using Integration2ndHalfWithWallNoRiemann = Integration2ndHalfWithWall<NoRiemannSolverCK>;

// AcousticTimeStep: 0.6 h / (max(c0 + |v|) + tiny)
class AcousticTimeStep : public AcousticTimeStepBase
{
  public:
    explicit AcousticTimeStep(SPHBody &body, Real acousticCFL = Real(0.6)) : AcousticTimeStepBase(body, acousticCFL)
    {
        formulation_ = 1;
        particles_.registerStateVariable<Real>("DensityChangeRate");
    }
};
// AdvectionTimeStep / AdvectionViscousTimeStep (inviscid here): 0.25 h / (max(sqrt(max(|v|^2, 4 h |F + F_prior| / m)), U_ref) + tiny)
class AdvectionTimeStep : public FluidDynamicsBase
{
    Real u_ref_, cfl_, h_min_, reduced_ = 0;

  public:
    using OutputType = Real;
    AdvectionTimeStep(SPHBody &body, Real U_ref, Real advectionCFL = Real(0.25))
        : FluidDynamicsBase(body), u_ref_(U_ref), cfl_(advectionCFL), h_min_(body.getSPHAdaptation().MinimumSmoothingLength())
    {
        formulation_ = 1;
        particles_.registerStateVariable<Vecd>("Velocity");
        particles_.registerStateVariable<Vecd>("Force");
        particles_.registerStateVariable<Vecd>("ForcePrior");
    }
    Real ReducedValue() const { return reduced_; }
    Real deviceReduce(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        float dt = 0;
        SPHCK_CALL(sphb200_advection_time_step_legacy, &f, h_min_, u_ref_, cfl_, &reduced_, &dt, execution_instance().stream());
        return dt;
    }
};
using AdvectionViscousTimeStep = AdvectionTimeStep;
} // namespace fluid_dynamics
using TotalMechanicalEnergy = TotalMechanicalEnergyCK;
} // namespace SPH
#endif
