// sphinxsys_ck/fluid_dynamics.h — the algorithm drivers (StateDynamics / ReduceDynamicsCK / InteractionDynamicsCK)
// and the local dynamics of the weakly-compressible fluid hot path, each ending in one C-ABI call.
//
// Reference (relative to /root/reference/src/shared/shared_ck/particle_dynamics unless noted):
//   StateDynamics, ReduceDynamicsCK ......... simple_algorithms_ck.h:41-121
//   InteractionDynamicsCK ................... interaction_algorithms_ck.{h,hpp,cpp} (init -> pre -> interact -> post -> update)
//   GravityForceCK .......................... general_dynamics/force_prior_ck.{h,hpp}
//   AdvectionStepSetup, UpdateParticlePosition, AdvectionTimeStepCK, AcousticTimeStepCK
//                                             fluid_dynamics/fluid_time_step_ck.{h,hpp,cpp}
//   CompressionSummation, DensityRegularization  fluid_dynamics/density_regularization.{h,hpp}
//   AcousticStep1stHalf / 2ndHalf + aliases .... fluid_dynamics/acoustic_step_1st_half.{h,hpp}, acoustic_step_2nd_half.{h,hpp}
//   Riemann solvers ......................... fluid_dynamics/riemann_solver/riemann_solver_ck.h:46-173
//   LinearCorrectionMatrix .................. general_dynamics/kernel_correction_ck.{h,hpp}
//   TotalMechanicalEnergyCK ................. general_dynamics/general_reduce_ck.h:41-96
//
// The closed set of supported type tuples is SURVEY.md §8a: Inner<OneLevel, Riemann, Correction> x Contact<Wall, same, same>.
#ifndef SPHINXSYS_CK_FLUID_DYNAMICS_H
#define SPHINXSYS_CK_FLUID_DYNAMICS_H

#include <fstream>
#include <iomanip>

#include "configuration.h"
#include "slab_decomposition.h"

namespace SPH
{
// ---- type tags (spelled as in the reference) ----
// LinearCorrectionRecord: the symmetric part of LinearCorrectionMatrix as one 32-byte gather record (sphb200.h,
// sphb200_fluid_view_t::correction_record) — what AcousticStep1stHalf<..., LinearCorrectionCK> reads of its neighbours.
// The identity until LinearCorrectionMatrix has run, like the matrix itself; a registered state variable, so it is
// reordered, migrated and refreshed on ghost planes with the matrix.
inline void registerCorrectionRecord(BaseParticles &p)
{
    GatherRecord8 identity;
    identity.v[0] = identity.v[3] = identity.v[5] = Real(1);
    p.registerStateVariable<GatherRecord8>("LinearCorrectionRecord", identity);
}
struct Base {};
struct WithUpdate {};
struct WithInitialization {};
struct OneLevel {};
struct Wall {};
struct FreeSurface {};
struct Internal {};
struct Boundary {};
struct BulkParticles {};
// limiters, common/common_functors.h:69-94
struct NoLimiter { static constexpr int kind = 0; static constexpr float slope = 0.f; };
struct TruncatedLinear { static constexpr int kind = 1; static constexpr float slope = 100.f; };
struct NoKernelCorrectionCK { static constexpr int kind = 0; };
struct LinearCorrectionCK { static constexpr int kind = 1; };
namespace fluid_dynamics
{
struct NoRiemannSolverCK { static constexpr int kind = 0; };
struct AcousticRiemannSolverCK { static constexpr int kind = 1; };
struct DissipativeRiemannSolverCK { static constexpr int kind = 2; };
} // namespace fluid_dynamics

class Gravity
{
    Vecd g_;

  public:
    explicit Gravity(const Vecd &g, const Vecd &reference_position = Vecd()) : g_(g) { (void)reference_position; }
    const Vecd &InducedAcceleration() const { return g_; }
    void toArray(float out[3]) const { out[0] = g_.x; out[1] = g_.y; out[2] = g_.z; }
};

// DynamicsArgs(identifier, args...): interaction_algorithms_ck / base_local_dynamics.h
template <class Identifier, typename... Args> struct DynamicsArgsT
{
    Identifier &identifier_;
    std::tuple<Args...> others_;
};
template <class Identifier, typename... Args> DynamicsArgsT<Identifier, Args...> DynamicsArgs(Identifier &id, Args... args)
{
    return DynamicsArgsT<Identifier, Args...>{id, std::make_tuple(args...)};
}

// =========================================================================================================
// algorithm drivers
// =========================================================================================================
template <class ExecutionPolicy, class UpdateType> class StateDynamics : public UpdateType, public BaseDynamics<void>
{
  public:
    template <typename... Args> explicit StateDynamics(Args &&...args) : UpdateType(std::forward<Args>(args)...)
    {
        execution::require_device_policy<ExecutionPolicy>();
    }
    void exec(Real dt = 0.0) override { this->deviceUpdate(dt); }
};

template <class ExecutionPolicy, class ReduceType>
class ReduceDynamicsCK : public ReduceType, public BaseDynamics<typename ReduceType::OutputType>
{
  public:
    using OutputType = typename ReduceType::OutputType;
    template <typename... Args> explicit ReduceDynamicsCK(Args &&...args) : ReduceType(std::forward<Args>(args)...)
    {
        execution::require_device_policy<ExecutionPolicy>();
    }
    OutputType exec(Real dt = 0.0) override { return this->deviceReduce(dt); }
};

class InteractionDynamicsBase : public BaseDynamics<void>
{
  protected:
    std::vector<BaseDynamics<void> *> pre_processes_, post_processes_;

  public:
    // interaction_algorithms_ck.h: addPre/PostContactInteraction, addPostStateDynamics (by reference form)
    InteractionDynamicsBase &addPreContactInteraction(BaseDynamics<void> &d) { pre_processes_.push_back(&d); return *this; }
    InteractionDynamicsBase &addPostContactInteraction(BaseDynamics<void> &d) { post_processes_.push_back(&d); return *this; }
    InteractionDynamicsBase &addPostStateDynamics(BaseDynamics<void> &d) { post_processes_.push_back(&d); return *this; }
};

// dynamics whose initialize step writes something neighbours read in the interact step expose the two phases
template <class T, class = void> struct HasInitializePhase : std::false_type {};
template <class T>
struct HasInitializePhase<T, std::void_t<decltype(std::declval<T &>().deviceInitialize(Real(0))), decltype(std::declval<T &>().deviceInteractAndUpdate(Real(0)))>>
    : std::true_type {};

template <class ExecutionPolicy, class InteractionType>
class InteractionDynamicsCK : public InteractionType, public InteractionDynamicsBase
{
  public:
    template <typename... Args> explicit InteractionDynamicsCK(Args &&...args) : InteractionType(std::forward<Args>(args)...)
    {
        execution::require_device_policy<ExecutionPolicy>();
    }
    // runAllSteps (interaction_algorithms_ck.cpp:6-34): [initialize] -> pre -> interact(inner, contacts) -> post -> [update].
    // initialize/interact/update of one dynamics are fused inside the library wherever no neighbour reads the value
    // being written; a post process the library can fold into the same launch is handed to deviceInteract().
    // With pre processes queued (e.g. the ghost update of a periodic condition, throat.cpp:183-184) a dynamics whose
    // initialize step feeds its interact step runs initialize first, exactly as the reference orders them.
    void exec(Real dt = 0.0) override
    {
        if constexpr (HasInitializePhase<InteractionType>::value)
        {
            if (!pre_processes_.empty())
            {
                this->deviceInitialize(dt);
                for (auto *d : pre_processes_) d->exec(dt);
                this->deviceInteractAndUpdate(dt);
                for (auto *d : post_processes_) d->exec(dt);
                return;
            }
        }
        for (auto *d : pre_processes_) d->exec(dt);
        std::vector<BaseDynamics<void> *> remaining = this->deviceInteract(dt, post_processes_);
        for (auto *d : remaining) d->exec(dt);
    }
};

// =========================================================================================================
// shared plumbing: the argument block of one fluid body (+ wall)
// =========================================================================================================
namespace fluid_dynamics
{
class FluidDynamicsBase
{
  protected:
    SPHBody &sph_body_;
    BaseParticles &particles_;
    RelationBase *inner_ = nullptr, *contact_ = nullptr;
    int riemann_ = 1, correction_ = 0, free_surface_ = 1;
    SlabDecomposition *decomposition_ = nullptr; // reductions become global when set
    int formulation_ = 0;                        // 0 CK, 1 legacy (legacy_dynamics.h)
    Real sigma0_ = 0, wall_rho0_ = 1;            // legacy DensitySummation constants

  public:
    void setDecomposition(SlabDecomposition *d) { decomposition_ = d; }
    explicit FluidDynamicsBase(SPHBody &body) : sph_body_(body), particles_(body.getBaseParticles()) {}
    FluidDynamicsBase(RelationBase &inner, RelationBase *contact)
        : sph_body_(inner.source_), particles_(inner.source_.getBaseParticles()), inner_(&inner), contact_(contact) {}
    SPHBody &getSPHBody() { return sph_body_; }

  protected:
    // AcousticStep constructor: acoustic_step_1st_half.hpp:13-39
    void registerAcousticVariables()
    {
        BaseParticles &p = particles_;
        p.registerStateVariable<Real>("Pressure");
        p.registerStateVariable<Real>("Compression", Real(1));
        p.registerStateVariable<Real>("CompressionRate");
        p.registerStateVariable<Vecd>("Velocity");
        p.registerStateVariable<Vecd>("Displacement");
        p.registerStateVariable<Vecd>("Force");
        p.registerStateVariable<Vecd>("ForcePrior");
        p.addEvolvingVariable<Vecd>("Velocity");
        p.addEvolvingVariable<Real>("Mass");
        p.addEvolvingVariable<Vecd>("ForcePrior");
        p.addEvolvingVariable<Real>("Compression");
        p.addEvolvingVariable<Real>("CompressionRate");
    }
    // CompressionSummation constructor: density_regularization.hpp:14-27
    void registerSummationVariables()
    {
        BaseParticles &p = particles_;
        bool fresh = !p.hasVariable("VolumetricMeasureRef");
        auto *ref = p.registerStateVariable<Real>("VolumetricMeasureRef");
        if (fresh)
        {
            ExecutionInstance &ex = execution_instance();
            ex.check(sphb200_copy_d2d(ref->deviceAddress(), p.deviceData<Real>("VolumetricMeasure"),
                                      p.TotalRealParticles() * sizeof(Real), ex.stream()), "sphb200_copy_d2d");
        }
        p.registerStateVariable<Real>("CompressionSummation", Real(1));
        p.registerStateVariable<Real>("Compression", Real(1));
        p.addEvolvingVariable<Real>("VolumetricMeasureRef");
    }
    sphb200_fluid_view_t fluidView()
    {
        BaseParticles &p = particles_;
        sphb200_fluid_view_t f;
        std::memset(&f, 0, sizeof(f));
        f.n = (uint32_t)p.TotalRealParticles();
        f.pos = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("Position");
        f.vel = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("Velocity");
        f.dpos = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("Displacement");
        f.force = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("Force");
        f.force_prior = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("ForcePrior");
        f.vol = (float *)p.deviceDataOrNull<Real>("VolumetricMeasure");
        f.mass = (float *)p.deviceDataOrNull<Real>("Mass");
        f.rho = (float *)p.deviceDataOrNull<Real>("Density");
        f.p = (float *)p.deviceDataOrNull<Real>("Pressure");
        f.compression = (float *)p.deviceDataOrNull<Real>("Compression");
        f.compression_rate = (float *)p.deviceDataOrNull<Real>("CompressionRate");
        f.vol_ref = (float *)p.deviceDataOrNull<Real>("VolumetricMeasureRef");
        f.compression_sum = (float *)p.deviceDataOrNull<Real>("CompressionSummation");
        f.B = (float *)p.deviceDataOrNull<Matd>("LinearCorrectionMatrix");
        f.correction_record = p.deviceDataOrNull<GatherRecord8>("LinearCorrectionRecord");
        f.posvol = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("PosVol");
        f.posvolref = (sphb200_vec4_t *)p.deviceDataOrNull<Vecd>("PosVolRef");
        f.posvolvel = p.deviceDataOrNull<GatherRecord8>("PosVolVel");
        f.active_begin = (uint32_t)p.activeBegin();
        f.active_end = (uint32_t)p.activeEnd();
        if (formulation_ == 1)
        {
            // legacy state: Density/DensityChangeRate, positions advance inside the half steps (dpos aliases pos), pair
            // geometry comes from the gather records frozen at the last updateConfiguration()
            f.dpos = f.pos;
            f.compression = nullptr;
            f.compression_rate = (float *)p.deviceDataOrNull<Real>("DensityChangeRate");
            f.compression_sum = (float *)p.deviceDataOrNull<Real>("DensitySummation");
            f.posvolref = f.posvol; // no VolumetricMeasureRef in the legacy state: the summation ignores the weight
        }
        return f;
    }
    sphb200_fluid_t material()
    {
        WeaklyCompressibleFluid &fl = dynamic_cast<WeaklyCompressibleFluid &>(sph_body_.getBaseMaterial());
        sphb200_fluid_t m;
        m.rho0 = fl.rho0_;
        m.c0 = fl.c0_;
        m.riemann = riemann_;
        m.correction = correction_;
        m.limiter_coeff = Real(3.0); // riemann_solver_ck.h:97
        m.free_surface = free_surface_;
        m.formulation = formulation_;
        m.sigma0 = sigma0_;
        m.wall_rho0 = wall_rho0_;
        return m;
    }
    sphb200_fluid_args_t fluidArgs()
    {
        sphb200_fluid_args_t a;
        std::memset(&a, 0, sizeof(a));
        if (sph_body_.periodicImages()) SPHCK_STAGE("  args: pending periodic images", sph_body_.periodicImages()->ensure());
        SPHCK_STAGE("  args: gather records", sph_body_.refreshPosVol());
        a.fluid = fluidView();
        a.material = material();
        if (inner_)
        {
            a.inner = inner_->view();
            a.kernel = inner_->kernel_;
        }
        else
            a.kernel = sph_body_.getSPHAdaptation().kernel_;
        if (contact_ && contact_->target_.TotalRealParticles())
        {
            SPHBody &w = contact_->target_;
            w.refreshPosVol();
            BaseParticles &wp = w.getBaseParticles();
            a.wall.n = (uint32_t)wp.TotalRealParticles();
            a.wall.pos = (const sphb200_vec4_t *)wp.deviceData<Vecd>("Position");
            a.wall.posvol = (const sphb200_vec4_t *)wp.deviceData<Vecd>("PosVol");
            a.wall.posvolref = (const sphb200_vec4_t *)wp.deviceDataOrNull<Vecd>("PosVolRef");
            a.wall.vel_ave = (const sphb200_vec4_t *)wp.deviceDataOrNull<Vecd>("AverageVelocity");
            a.wall.acc_ave = (const sphb200_vec4_t *)wp.deviceDataOrNull<Vecd>("AverageAcceleration");
            a.wall.normal = (const sphb200_vec4_t *)wp.deviceDataOrNull<Vecd>("NormalDirection");
            a.wall.vol_ref = (const float *)wp.deviceDataOrNull<Real>("VolumetricMeasureRef");
            a.contact = contact_->view();
        }
        return a;
    }
};

// ---- StateDynamics local dynamics ----
class AdvectionStepSetup : public FluidDynamicsBase
{
  public:
    explicit AdvectionStepSetup(SPHBody &body) : FluidDynamicsBase(body) { particles_.registerStateVariable<Vecd>("Displacement"); }
    void deviceUpdate(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        SPHCK_CALL(sphb200_advection_setup, &f, execution_instance().stream());
        sph_body_.setPosVolDirty();
    }
};
class UpdateParticlePosition : public FluidDynamicsBase
{
  public:
    explicit UpdateParticlePosition(SPHBody &body) : FluidDynamicsBase(body) { particles_.registerStateVariable<Vecd>("Displacement"); }
    void deviceUpdate(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        SPHCK_CALL(sphb200_update_position, &f, execution_instance().stream());
        sph_body_.setPosVolDirty();
    }
};
class DensityRegularizationBase : public FluidDynamicsBase
{
  public:
    using FluidDynamicsBase::FluidDynamicsBase;
    int flowType() const { return free_surface_; }
    void deviceUpdate(Real)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_density_regularization, &a, execution_instance().stream());
    }
};
template <class BodyType, class FluidType, class FlowType> class DensityRegularization : public DensityRegularizationBase
{
  public:
    explicit DensityRegularization(SPHBody &body) : DensityRegularizationBase(body)
    {
        free_surface_ = std::is_same<FlowType, FreeSurface>::value ? 1 : 0;
        registerSummationVariables();
    }
};

// ---- ReduceDynamicsCK local dynamics ----
class AdvectionTimeStepCK : public FluidDynamicsBase
{
    Real u_ref_, cfl_, h_min_, reduced_ = 0;

  public:
    using OutputType = Real;
    AdvectionTimeStepCK(SPHBody &body, Real U_ref, Real advectionCFL = Real(0.25))
        : FluidDynamicsBase(body), u_ref_(U_ref), cfl_(advectionCFL), h_min_(body.getSPHAdaptation().MinimumSmoothingLength())
    {
        particles_.registerStateVariable<Vecd>("Velocity");
    }
    Real ReducedValue() const { return reduced_; }
    Real deviceReduce(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        float dt = 0;
        SPHCK_CALL(sphb200_advection_time_step, &f, h_min_, u_ref_, cfl_, &reduced_, &dt, execution_instance().stream());
        if (decomposition_)
        {
            reduced_ = decomposition_->allReduceMax(reduced_);
            dt = cfl_ * h_min_ / (std::max(std::sqrt(reduced_), u_ref_) + TinyReal); // fluid_time_step_ck.cpp:24-27
        }
        return dt;
    }
};

class AcousticTimeStepBase : public FluidDynamicsBase
{
  protected:
    Real cfl_, h_min_, reduced_ = 0;
    DeviceBuffer fused_slot_; // device float written by the fused 2nd-half launch
    bool primed_ = false;

  public:
    using OutputType = Real;
    AcousticTimeStepBase(SPHBody &body, Real acousticCFL) : FluidDynamicsBase(body), cfl_(acousticCFL), h_min_(body.getSPHAdaptation().MinimumSmoothingLength())
    {
        registerAcousticVariables();
    }
    Real ReducedValue() const { return reduced_; }
    Real exec(Real dt = 0.0) { return deviceReduce(dt); } // same as ReduceDynamicsCK<...>::exec through a base pointer
    Real minimumSmoothingLength() const { return h_min_; }
    // hooks used by AcousticStep2ndHalf when the reduction is folded into its launch
    float *fusedSlot()
    {
        fused_slot_.ensure(64);
        return fused_slot_.get<float>();
    }
    void setPrimed(bool v) { primed_ = v; }
    Real deviceReduce(Real)
    {
        ExecutionInstance &ex = execution_instance();
        if (primed_)
        {
            // same 4-byte device->host read the stand-alone reduction ends with (particle_iterators_sycl.h:80-105)
            primed_ = false;
            if (decomposition_) decomposition_->allReduceMaxDevice(fused_slot_.get<float>());
            ex.check(sphb200_copy_d2h(&reduced_, fused_slot_.get(), sizeof(float), ex.stream()), "sphb200_copy_d2h");
            ex.synchronize();
            return cfl_ * h_min_ / (reduced_ + TinyReal); // FinishDynamics::Result, fluid_time_step_ck.hpp:31-36
        }
        sphb200_fluid_args_t a = fluidArgs();
        float dt = 0;
        SPHCK_CALL(sphb200_acoustic_time_step, &a, h_min_, cfl_, &reduced_, &dt, ex.stream());
        if (decomposition_)
        {
            reduced_ = decomposition_->allReduceMax(reduced_);
            dt = cfl_ * h_min_ / (reduced_ + TinyReal);
        }
        return dt;
    }
};
template <class FluidType = WeaklyCompressibleFluid> class AcousticTimeStepCK : public AcousticTimeStepBase
{
  public:
    explicit AcousticTimeStepCK(SPHBody &body, Real acousticCFL = Real(0.6)) : AcousticTimeStepBase(body, acousticCFL) {}
};

// ---- InteractionDynamicsCK local dynamics ----
template <typename... T> class CompressionSummation;
template <> class CompressionSummation<Inner<>, Contact<>> : public FluidDynamicsBase
{
  public:
    CompressionSummation(Inner<> &inner, Contact<> &contact) : FluidDynamicsBase(inner, &contact) { registerSummationVariables(); }
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        // a DensityRegularization of the same body queued as post process is folded into the summation launch
        int regularize = 0;
        std::vector<BaseDynamics<void> *> remaining;
        for (auto *d : post)
        {
            auto *reg = dynamic_cast<DensityRegularizationBase *>(d);
            if (reg && &reg->getSPHBody() == &sph_body_ && !regularize)
            {
                regularize = 1;
                free_surface_ = reg->flowType();
            }
            else
                remaining.push_back(d);
        }
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_STAGE("  summation: launch", SPHCK_CALL(sphb200_compression_summation, &a, regularize, execution_instance().stream()));
        return remaining;
    }
};
template <> class CompressionSummation<Inner<>> : public FluidDynamicsBase
{
  public:
    explicit CompressionSummation(Inner<> &inner) : FluidDynamicsBase(inner, nullptr) { registerSummationVariables(); }
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        int regularize = 0;
        std::vector<BaseDynamics<void> *> remaining;
        for (auto *d : post)
        {
            auto *reg = dynamic_cast<DensityRegularizationBase *>(d);
            if (reg && &reg->getSPHBody() == &sph_body_ && !regularize)
            {
                regularize = 1;
                free_surface_ = reg->flowType();
            }
            else
                remaining.push_back(d);
        }
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_STAGE("  summation: launch", SPHCK_CALL(sphb200_compression_summation, &a, regularize, execution_instance().stream()));
        return remaining;
    }
};

// phase-granular access (initialize | interact + update), used by parity tests against the reference's phases
class AcousticStep1stHalfPhases
{
  public:
    virtual ~AcousticStep1stHalfPhases() {}
    virtual void deviceInitialize(Real dt) = 0;
    virtual void deviceInteractAndUpdate(Real dt) = 0;
};

template <class RiemannType, class CorrectionType>
class AcousticStep1stHalfWithWall : public FluidDynamicsBase, public AcousticStep1stHalfPhases
{
  public:
    // AcousticStep1stHalf<Inner<OneLevel, Riemann, Correction>> without a contact part
    explicit AcousticStep1stHalfWithWall(Inner<> &inner) : FluidDynamicsBase(inner, nullptr)
    {
        riemann_ = RiemannType::kind;
        correction_ = CorrectionType::kind;
        registerAcousticVariables();
        if (correction_) particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }
    AcousticStep1stHalfWithWall(Inner<> &inner, Contact<> &contact) : FluidDynamicsBase(inner, &contact)
    {
        riemann_ = RiemannType::kind;
        correction_ = CorrectionType::kind;
        registerAcousticVariables();
        if (correction_) particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }
    std::vector<BaseDynamics<void> *> deviceInteract(Real dt, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_acoustic_1st_half, &a, dt, execution_instance().stream());
        return post;
    }
    void deviceInitialize(Real dt) override
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_acoustic_1st_half_initialize, &a, dt, execution_instance().stream());
    }
    void deviceInteractAndUpdate(Real dt) override
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_acoustic_1st_half_interact, &a, dt, 1, execution_instance().stream());
    }
};

// launch-granular access to the 2nd half (prime the fused reduction once, then one launch per active slot range)
class AcousticStep2ndHalfPhases
{
  public:
    virtual ~AcousticStep2ndHalfPhases() {}
    virtual void primeFusedReduction() = 0;
    virtual void deviceLaunch(Real dt) = 0;
};

template <class RiemannType, class CorrectionType> class AcousticStep2ndHalfWithWall : public FluidDynamicsBase, public AcousticStep2ndHalfPhases
{
    AcousticTimeStepBase *fused_time_step_ = nullptr;

  public:
    explicit AcousticStep2ndHalfWithWall(Inner<> &inner) : FluidDynamicsBase(inner, nullptr)
    {
        riemann_ = RiemannType::kind;
        correction_ = CorrectionType::kind;
        registerAcousticVariables();
        if (correction_) particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }
    AcousticStep2ndHalfWithWall(Inner<> &inner, Contact<> &contact) : FluidDynamicsBase(inner, &contact)
    {
        riemann_ = RiemannType::kind;
        correction_ = CorrectionType::kind;
        registerAcousticVariables();
        if (correction_) particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }
    // Fold the next AcousticTimeStepCK reduction into this launch (library extension: removes one pass over the
    // particles per acoustic step; the reduced value is bit-identical to the stand-alone reduction).
    void fuseTimeStepReduction(AcousticTimeStepBase &time_step) { fused_time_step_ = &time_step; }
    void unfuseTimeStepReduction() { fused_time_step_ = nullptr; }
    std::vector<BaseDynamics<void> *> deviceInteract(Real dt, const std::vector<BaseDynamics<void> *> &post)
    {
        primeFusedReduction();
        deviceLaunch(dt);
        return post;
    }
    // the two halves of deviceInteract(), for runs that advance the active slots in several launches (interior and
    // boundary planes of a decomposed run): prime once, then one deviceLaunch() per slot range — every launch folds its
    // particles into the same reduction slot
    void primeFusedReduction() override
    {
        if (!fused_time_step_) return;
        SPHCK_CALL(sphb200_fill_f32, fused_time_step_->fusedSlot(), 0.0f, 1, execution_instance().stream());
        fused_time_step_->setPrimed(true);
    }
    void deviceLaunch(Real dt) override
    {
        sphb200_fluid_args_t a = fluidArgs();
        float *slot = fused_time_step_ ? fused_time_step_->fusedSlot() : nullptr;
        Real h_min = fused_time_step_ ? fused_time_step_->minimumSmoothingLength() : sph_body_.getSPHAdaptation().MinimumSmoothingLength();
        SPHCK_CALL(sphb200_acoustic_2nd_half, &a, dt, h_min, slot, execution_instance().stream());
    }
};

// aliases, acoustic_step_1st_half.h:196-201 / acoustic_step_2nd_half.h:183-191
using AcousticStep1stHalfWithWallRiemannCK = AcousticStep1stHalfWithWall<AcousticRiemannSolverCK, NoKernelCorrectionCK>;
using AcousticStep2ndHalfWithWallRiemannCK = AcousticStep2ndHalfWithWall<AcousticRiemannSolverCK, NoKernelCorrectionCK>;
using AcousticStep1stHalfWithWallRiemannCorrectionCK = AcousticStep1stHalfWithWall<AcousticRiemannSolverCK, LinearCorrectionCK>;
using AcousticStep2ndHalfWithWallRiemannCorrectionCK = AcousticStep2ndHalfWithWall<AcousticRiemannSolverCK, LinearCorrectionCK>;
using AcousticStep1stHalfWithWallNoRiemannCK = AcousticStep1stHalfWithWall<NoRiemannSolverCK, NoKernelCorrectionCK>;
using AcousticStep2ndHalfWithWallNoRiemannCK = AcousticStep2ndHalfWithWall<NoRiemannSolverCK, NoKernelCorrectionCK>;
using AcousticStep1stHalfWithWallDissipativeRiemannCK = AcousticStep1stHalfWithWall<DissipativeRiemannSolverCK, NoKernelCorrectionCK>;
using AcousticStep2ndHalfWithWallDissipativeRiemannCK = AcousticStep2ndHalfWithWall<DissipativeRiemannSolverCK, NoKernelCorrectionCK>;
// AcousticStep1stHalf/2ndHalf<Inner<OneLevel, Riemann, NoKernelCorrectionCK>>: the same classes built from the inner
// relation alone (bodies without walls, e.g. the periodic Taylor-Green vortex)
using AcousticStep1stHalfInnerRiemannCK = AcousticStep1stHalfWithWallRiemannCK;
using AcousticStep2ndHalfInnerRiemannCK = AcousticStep2ndHalfWithWallRiemannCK;
using AcousticStep1stHalfInnerNoRiemannCK = AcousticStep1stHalfWithWallNoRiemannCK;
using AcousticStep2ndHalfInnerNoRiemannCK = AcousticStep2ndHalfWithWallNoRiemannCK;
// ---- free-surface indication (general_dynamics/surface_indication/surface_indication_ck.h:19-170) ----
// InteractionDynamicsCK<P, FreeSurfaceIndicationCK<Inner<WithUpdate>, Contact<>>>(inner, contact).exec():
// inner interact (PositionDivergence, spatial-temporal override next to the previous surface) -> contact interact ->
// update (Indicator, PreviousSurfaceIndicator).
template <typename... RelationTypes> class FreeSurfaceIndicationCK;
template <> class FreeSurfaceIndicationCK<Inner<WithUpdate>, Contact<>> : public FluidDynamicsBase
{
    Real threshold_by_dimensions_, smoothing_length_;

  public:
    FreeSurfaceIndicationCK(Inner<> &inner, Contact<> &contact)
        : FluidDynamicsBase(inner, &contact), threshold_by_dimensions_(Real(0.75) * Real(inner.source_.getSPHSystem().dim_)),
          smoothing_length_(inner.source_.getSPHAdaptation().ReferenceSmoothingLength())
    {
        // surface_indication_ck.hpp:17-20,42-46
        particles_.registerStateVariable<int>("Indicator");
        particles_.registerStateVariable<Real>("PositionDivergence");
        particles_.registerStateVariable<int>("PreviousSurfaceIndicator", 1);
        particles_.addEvolvingVariable<int>("PreviousSurfaceIndicator");
    }
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        int32_t *indicator = (int32_t *)particles_.deviceData<int>("Indicator");
        float *position_divergence = (float *)particles_.deviceData<Real>("PositionDivergence");
        int32_t *previous = (int32_t *)particles_.deviceData<int>("PreviousSurfaceIndicator");
        if (decomposition_)
        {
            // the second sweep reads PositionDivergence of the neighbours: on the ghost planes it comes from their owners
            // (tests/test_decomposed_oracle_cpu.py::test_dam_break_complete_case_dynamics_bit_identical)
            SPHCK_CALL(sphb200_free_surface_indication_sweep, &a, indicator, position_divergence, previous, threshold_by_dimensions_,
                       smoothing_length_, 0, execution_instance().stream());
            decomposition_->refreshGhosts({"PositionDivergence"});
            SPHCK_CALL(sphb200_free_surface_indication_sweep, &a, indicator, position_divergence, previous, threshold_by_dimensions_,
                       smoothing_length_, 1, execution_instance().stream());
            return post;
        }
        SPHCK_CALL(sphb200_free_surface_indication, &a, indicator, position_divergence, previous, threshold_by_dimensions_,
                   smoothing_length_, execution_instance().stream());
        return post;
    }
};
using FreeSurfaceIndicationComplexSpatialTemporalCK = FreeSurfaceIndicationCK<Inner<WithUpdate>, Contact<>>;
// ---- viscous force (fluid_dynamics/viscous_force.h:40-131) ----
// InteractionDynamicsCK<P, ViscousForceCK<Inner<WithUpdate, Viscosity, Correction>[, Contact<Wall, Viscosity, Correction>]>>:
// inner interact -> wall interact -> ForcePriorCK update ("ViscousForce" enters ForcePrior as a difference to its previous value)
template <typename... RelationTypes> class ViscousForceCK;
class ViscousForceBase : public FluidDynamicsBase
{
  protected:
    Real mu_, smoothing_length_;
    ViscousForceBase(Inner<> &inner, Contact<> *contact, int correction) : FluidDynamicsBase(inner, contact)
    {
        correction_ = correction;
        auto *visc = dynamic_cast<Viscosity *>(&inner.source_.getBaseMaterial());
        if (!visc) throw SphError("ViscousForceCK: the fluid material carries no Viscosity (use defineClosure<WeaklyCompressibleFluid, Viscosity>)");
        mu_ = visc->ReferenceViscosity();
        smoothing_length_ = inner.source_.getSPHAdaptation().ReferenceSmoothingLength();
        particles_.registerStateVariable<Vecd>("Velocity");
        particles_.registerStateVariable<Vecd>("ViscousForce");
        particles_.registerStateVariable<Vecd>("PreviousViscousForce"); // ForcePriorCK, force_prior_ck.cpp:13-14
        particles_.registerStateVariable<Vecd>("ForcePrior");
        particles_.addEvolvingVariable<Vecd>("ForcePrior");
        particles_.addEvolvingVariable<Vecd>("PreviousViscousForce");
        if (correction) particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }

  public:
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_viscous_force, &a, mu_, smoothing_length_, (sphb200_vec4_t *)particles_.deviceData<Vecd>("ViscousForce"),
                   (sphb200_vec4_t *)particles_.deviceData<Vecd>("PreviousViscousForce"), execution_instance().stream());
        return post;
    }
};
template <class CorrectionType> class ViscousForceCK<Inner<WithUpdate, Viscosity, CorrectionType>> : public ViscousForceBase
{
  public:
    explicit ViscousForceCK(Inner<> &inner) : ViscousForceBase(inner, nullptr, CorrectionType::kind) {}
};
template <class CorrectionType>
class ViscousForceCK<Inner<WithUpdate, Viscosity, CorrectionType>, Contact<Wall, Viscosity, CorrectionType>> : public ViscousForceBase
{
  public:
    ViscousForceCK(Inner<> &inner, Contact<> &contact) : ViscousForceBase(inner, &contact, CorrectionType::kind) {}
};
using ViscousForceInnerCK = ViscousForceCK<Inner<WithUpdate, Viscosity, NoKernelCorrectionCK>>;
using ViscousForceWithWallCK = ViscousForceCK<Inner<WithUpdate, Viscosity, NoKernelCorrectionCK>, Contact<Wall, Viscosity, NoKernelCorrectionCK>>;

// ---- transport velocity correction (fluid_dynamics/transport_velocity_correction_ck.h:12-44) ----
// StateDynamics<P, TransportVelocityCorrectionCK<SPHBody, LimiterType[, BulkParticles]>>(body[, coefficient = 0.2])
template <class DynamicsIdentifier, class LimiterType, typename... ParticleScopes> class TransportVelocityCorrectionCK : public FluidDynamicsBase
{
    Real h_ref_, coefficient_;
    static constexpr bool bulk_only_ = (std::is_same<ParticleScopes, BulkParticles>::value || ... || false);

  public:
    explicit TransportVelocityCorrectionCK(SPHBody &body, Real coefficient = Real(0.2))
        : FluidDynamicsBase(body), h_ref_(body.getSPHAdaptation().ReferenceSmoothingLength()), coefficient_(coefficient)
    {
        particles_.getVariableByName<Vecd>("Displacement");
        particles_.getVariableByName<Vecd>("KernelGradientIntegral");
        if (bulk_only_) particles_.getVariableByName<int>("Indicator");
    }
    void deviceUpdate(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        SPHCK_CALL(sphb200_transport_velocity_correction, &f, (const sphb200_vec4_t *)particles_.deviceData<Vecd>("KernelGradientIntegral"),
                   coefficient_, h_ref_, LimiterType::kind, LimiterType::slope,
                   bulk_only_ ? (const int32_t *)particles_.deviceData<int>("Indicator") : nullptr, execution_instance().stream());
    }
};
} // namespace fluid_dynamics

// ---- general dynamics ----
template <class GravityType> class GravityForceCK : public fluid_dynamics::FluidDynamicsBase
{
    GravityType gravity_;

  public:
    GravityForceCK(SPHBody &body, const GravityType &gravity) : FluidDynamicsBase(body), gravity_(gravity)
    {
        // ForcePriorCK: force_prior_ck.cpp:13-14
        particles_.registerStateVariable<Vecd>("ForcePrior");
        particles_.registerStateVariable<Vecd>("PreviousGravityForceCK");
        particles_.addEvolvingVariable<Vecd>("ForcePrior");
        particles_.addEvolvingVariable<Vecd>("PreviousGravityForceCK");
    }
    void deviceUpdate(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        float g[3];
        gravity_.toArray(g);
        SPHCK_CALL(sphb200_gravity_force, &f, g, (sphb200_vec4_t *)particles_.deviceData<Vecd>("PreviousGravityForceCK"),
                   execution_instance().stream());
    }
};

class LinearCorrectionMatrixComplex : public fluid_dynamics::FluidDynamicsBase
{
    Real alpha_;

  public:
    LinearCorrectionMatrixComplex(DynamicsArgsT<Inner<>, double> args, Contact<> &contact)
        : FluidDynamicsBase(args.identifier_, &contact), alpha_(Real(std::get<0>(args.others_)))
    {
        particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }
    LinearCorrectionMatrixComplex(Inner<> &inner, Contact<> &contact, Real alpha = Real(0))
        : FluidDynamicsBase(inner, &contact), alpha_(alpha)
    {
        particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_linear_correction_matrix, &a, alpha_, execution_instance().stream());
        return post;
    }
};

// ---- kernel gradient integral (general_dynamics/kernel_gradient_integral.h:36-114) ----
template <typename... RelationTypes> class KernelGradientIntegral;
class KernelGradientIntegralBase : public fluid_dynamics::FluidDynamicsBase
{
  protected:
    KernelGradientIntegralBase(Inner<> &inner, Contact<> *contact, int correction) : FluidDynamicsBase(inner, contact)
    {
        correction_ = correction;
        particles_.registerStateVariable<Vecd>("KernelGradientIntegral");
        if (correction) particles_.registerStateVariable<Matd>("LinearCorrectionMatrix", Matd::Identity()), registerCorrectionRecord(particles_);
    }

  public:
    std::vector<BaseDynamics<void> *> deviceInteract(Real, const std::vector<BaseDynamics<void> *> &post)
    {
        sphb200_fluid_args_t a = fluidArgs();
        SPHCK_CALL(sphb200_kernel_gradient_integral, &a, (sphb200_vec4_t *)particles_.deviceData<Vecd>("KernelGradientIntegral"),
                   execution_instance().stream());
        return post;
    }
};
template <class CorrectionType> class KernelGradientIntegral<Inner<CorrectionType>> : public KernelGradientIntegralBase
{
  public:
    explicit KernelGradientIntegral(Inner<> &inner) : KernelGradientIntegralBase(inner, nullptr, CorrectionType::kind) {}
};
template <class CorrectionType>
class KernelGradientIntegral<Inner<CorrectionType>, Contact<Boundary, CorrectionType>> : public KernelGradientIntegralBase
{
  public:
    KernelGradientIntegral(Inner<> &inner, Contact<> &contact) : KernelGradientIntegralBase(inner, &contact, CorrectionType::kind) {}
};
using KernelGradientIntegralInner = KernelGradientIntegral<Inner<NoKernelCorrectionCK>>;
using KernelGradientIntegralComplex = KernelGradientIntegral<Inner<NoKernelCorrectionCK>, Contact<Boundary, NoKernelCorrectionCK>>;
using KernelGradientIntegralCorrectedComplex = KernelGradientIntegral<Inner<LinearCorrectionCK>, Contact<Boundary, LinearCorrectionCK>>;

class TotalMechanicalEnergyCK : public fluid_dynamics::FluidDynamicsBase
{
    Gravity gravity_;

  public:
    using OutputType = double;
    TotalMechanicalEnergyCK(SPHBody &body, const Gravity &gravity) : FluidDynamicsBase(body), gravity_(gravity)
    {
        particles_.registerStateVariable<Vecd>("Velocity");
    }
    double deviceReduce(Real)
    {
        sphb200_fluid_view_t f = fluidView();
        float g[3];
        gravity_.toArray(g);
        double e = 0;
        SPHCK_CALL(sphb200_total_mechanical_energy, &f, g, &e, execution_instance().stream());
        if (decomposition_) e = decomposition_->allReduceSum(e);
        return e;
    }
};

// ---- observation (general_dynamics/interpolation_dynamics.h:43-150, io_system/io_observation_ck.h:38-107) ----
// ObservingQuantityCK<P, DataType>(contact, "Name"): Interpolation<Contact<DataType>> of a variable of the observed
// body at the particles of an observer body; the result lives in the observer's variable of the same name.
struct RestoringCorrection {}; // Interpolation<Contact<DataType, RestoringCorrection>>, interpolation_dynamics.hpp:72-100
template <class... Parameters> struct UsesRestoringCorrection : std::false_type {};
template <class First, class... Rest>
struct UsesRestoringCorrection<First, Rest...>
    : std::integral_constant<bool, std::is_same<First, RestoringCorrection>::value || UsesRestoringCorrection<Rest...>::value> {};

template <class ExecutionPolicy, class DataType, class... Parameters> class ObservingQuantityCK : public BaseDynamics<void>
{
    Contact<> &contact_;
    std::string variable_name_;
    DiscreteVariable<DataType> *dv_interpolated_quantities_;

  public:
    ObservingQuantityCK(Contact<> &contact, const std::string &variable_name)
        : contact_(contact), variable_name_(variable_name),
          dv_interpolated_quantities_(contact.source_.getBaseParticles().template registerStateVariable<DataType>(variable_name))
    {
        execution::require_device_policy<ExecutionPolicy>();
        static_assert(std::is_same<DataType, Real>::value || std::is_same<DataType, Vecd>::value, "observed quantities are Real or Vecd");
        contact.target_.getBaseParticles().template getVariableByName<DataType>(variable_name); // must exist on the observed body
    }
    DiscreteVariable<DataType> *dvInterpolatedQuantities() { return dv_interpolated_quantities_; }
    void exec(Real dt = 0.0) override
    {
        SPHBody &observer = contact_.source_, &observed = contact_.target_;
        observed.refreshPosVol();
        BaseParticles &op = observer.getBaseParticles(), &tp = observed.getBaseParticles();
        const sphb200_vec4_t *src = (const sphb200_vec4_t *)op.deviceData<Vecd>("Position"), *tar = (const sphb200_vec4_t *)tp.deviceData<Vecd>("PosVol");
        const float *data = (const float *)tp.template deviceData<DataType>(variable_name_);
        float *out = (float *)dv_interpolated_quantities_->deviceAddress();
        const int width = std::is_same<DataType, Vecd>::value ? 4 : 1;
        // (observing "Position" writes the observer's own positions — the reference registers the interpolated variable under
        // the same name, interpolation_dynamics.hpp:20-27; every thread reads and writes its own observer only)
        if (UsesRestoringCorrection<Parameters...>::value)
            SPHCK_CALL(sphb200_interpolate_restoring, &contact_.kernel_, src, (uint32_t)op.TotalRealParticles(), contact_.view(), tar, data,
                       width, out, execution_instance().stream());
        else
            SPHCK_CALL(sphb200_interpolate, &contact_.kernel_, src, (uint32_t)op.TotalRealParticles(), contact_.view(), tar, data, width,
                       out, execution_instance().stream());
    }

};

// ObservedQuantityRecording<P, DataType>(contact, "Name"): writeToFile(iteration) runs the observation, brings the values
// to the host and appends one record (physical time + one value per observer); records are kept in memory and, when an
// output path is set, appended to a .dat file in the reference's column layout (io_observation_ck.h:69-94).
template <class ExecutionPolicy, class DataType, class... Parameters> class ObservedQuantityRecording
{
    SPHBody &observer_;
    ObservingQuantityCK<ExecutionPolicy, DataType, Parameters...> observation_method_;
    DiscreteVariable<DataType> *dv_interpolated_quantities_;
    size_t number_of_observe_;
    SingleVariable<Real> *sv_physical_time_;
    std::string quantity_name_, file_path_;
    bool header_written_ = false;
    std::vector<Real> times_;
    std::vector<std::vector<DataType>> records_;
    SlabDecomposition *decomposition_ = nullptr;

    // Slab-decomposed runs: every rank interpolates every probe over the particles it stores; the rank that owns the probe's
    // cell plane stores its whole neighbourhood (own planes + one ghost plane either side, all variables current right
    // after the configuration update), so its value is the single-GPU one, bit for bit. One term per probe is non-zero:
    // the sum over the ranks is exact (tests/test_decomposed_oracle_cpu.py::test_dam_break_observer_probes_bit_identical).
    void combineOverRanks()
    {
        static_assert(sizeof(DataType) % sizeof(Real) == 0, "observed quantities are made of Reals");
        constexpr size_t width = sizeof(DataType) / sizeof(Real);
        auto *dv_pos = observer_.getBaseParticles().template getVariableByName<Vecd>("Position");
        Real *v = reinterpret_cast<Real *>(dv_interpolated_quantities_->Data());
        std::vector<double> sum(number_of_observe_ * width, 0.0);
        for (size_t i = 0; i != number_of_observe_; ++i)
            if (decomposition_->ownsPlane(decomposition_->planeOf(dv_pos->Data()[i].x)))
                for (size_t c = 0; c != width; ++c) sum[i * width + c] = double(v[i * width + c]);
        for (size_t first = 0; first < sum.size(); first += 64)
            decomposition_->allReduceSum(sum.data() + first, (int)std::min<size_t>(64, sum.size() - first));
        for (size_t k = 0; k != sum.size(); ++k) v[k] = Real(sum[k]);
    }

  public:
    void setDecomposition(SlabDecomposition *d) { decomposition_ = d; }
    ObservedQuantityRecording(Contact<> &contact_relation, const std::string &variable_name)
        : observer_(contact_relation.getSPHBody()), observation_method_(contact_relation, variable_name),
          dv_interpolated_quantities_(observation_method_.dvInterpolatedQuantities()),
          number_of_observe_(observer_.getBaseParticles().TotalRealParticles()),
          sv_physical_time_(observer_.getSPHSystem().template getSystemVariableByName<Real>("PhysicalTime")), quantity_name_(variable_name) {}
    void setOutputPath(const std::string &folder) { file_path_ = folder + "/" + observer_.Name() + "_" + quantity_name_ + ".dat"; }
    void writeToFile(size_t iteration_step = 0)
    {
        observation_method_.exec();
        dv_interpolated_quantities_->synchronizeWithDevice();
        if (decomposition_) combineOverRanks();
        const DataType *v = dv_interpolated_quantities_->Data();
        times_.push_back(sv_physical_time_->getValue());
        records_.emplace_back(v, v + number_of_observe_);
        if (file_path_.empty()) return;
        if (!header_written_)
        {
            std::ofstream out(file_path_.c_str(), std::ios::out);
            out << "run_time" << "   ";
            for (size_t i = 0; i != number_of_observe_; ++i) out << quantity_name_ << "[" << i << "]" << "   ";
            out << "\n";
            header_written_ = true;
        }
        std::ofstream out(file_path_.c_str(), std::ios::app);
        out << sv_physical_time_->getValue() << "   ";
        for (size_t i = 0; i != number_of_observe_; ++i) writeValue(out, v[i]);
        out << "\n";
    }
    DataType *getObservedQuantity() { return dv_interpolated_quantities_->Data(); }
    size_t NumberOfObservedQuantity() { return number_of_observe_; }
    DiscreteVariable<DataType> &getObservedVariable() { return *dv_interpolated_quantities_; }
    const std::vector<Real> &recordedTimes() const { return times_; }
    const std::vector<std::vector<DataType>> &records() const { return records_; }

  private:
    static void writeValue(std::ostream &out, Real v) { out << std::fixed << std::setprecision(9) << v << "   "; }
    static void writeValue(std::ostream &out, const Vecd &v) { out << std::fixed << std::setprecision(9) << v.x << "   " << v.y << "   " << v.z << "   "; }
};
} // namespace SPH
#endif
