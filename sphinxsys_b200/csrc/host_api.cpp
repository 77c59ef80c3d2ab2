// host_api.cpp — C entry points that let the Python test/bench harness drive the C++ host layer
// (include/sphinxsys_ck/*.h). Builds libsphb200_host.so; links libsphb200.so (the CUDA library) and nothing else.
// All particle arrays cross this boundary in the REFERENCE particle order and in the reference's packed layout
// (Vecd = 3 floats), i.e. what a SPHinXsys user sees in DiscreteVariable::Data().
#include <cstring>
#include <string>

#include "../../include/sphinxsys_ck/dambreak_case.h"
#include "../../include/sphinxsys_ck/taylor_green_case.h"

using namespace SPH;

namespace
{
thread_local std::string g_error;

int g_ring_handles = 0; // ring runs share the communicator of the (process-wide) context: the last one gives it back

struct Handle
{
    std::unique_ptr<DamBreakCK> sim;   // dam break (CK or legacy spelling) ...
    std::unique_ptr<TaylorGreenCK> tg; // ... or the periodic Taylor-Green vortex
    std::unique_ptr<HostTransferPipeline> pipeline; // overlapped host <-> device transfers of the fluid body (bench e2e)
    bool owns_comm = false; // ring runs made the context's communicator: give the context back without it
    ~Handle()
    {
        pipeline.reset();
        tg.reset();
        sim.reset();
        if (owns_comm && --g_ring_handles == 0) sphb200_comm_destroy(execution_instance().ctx());
    }
    SPHBody &body(int which)
    {
        if (tg)
        {
            if (which) throw SphError("the Taylor-Green case has no wall body");
            return tg->water_block;
        }
        return which ? (SPHBody &)sim->wall_boundary : (SPHBody &)sim->water_block;
    }
    RelationBase &relation(int which)
    {
        if (tg)
        {
            if (which) throw SphError("the Taylor-Green case has no contact relation");
            return *tg->water_block_inner;
        }
        return which ? (RelationBase &)*sim->water_wall_contact : (RelationBase &)*sim->water_block_inner;
    }
    fluid_dynamics::AcousticTimeStepBase *acousticTimeStep() { return tg ? tg->fluid_acoustic_time_step.get() : sim->fluid_acoustic_time_step; }
};

template <class F> int guarded(F &&f)
{
    try
    {
        f();
        return 0;
    }
    catch (const std::exception &e)
    {
        g_error = e.what();
        return -1;
    }
}

std::vector<Vecd> toVecd(const float *xyz, uint64_t n)
{
    std::vector<Vecd> v(n);
    for (uint64_t i = 0; i < n; ++i) v[i] = Vecd(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    return v;
}

SPHBody &body(Handle *h, int which) { return h->body(which); }
} // namespace

extern "C"
{
    struct sphck_dambreak_options
    {
        int32_t dim;
        double dp, DL, DH, DW, LL, LH, LW;
        int32_t correction, fused_time_step, fused_regularization, sort_interval, device;
        int32_t relation_stride;  // < 0: default; 0: exact two-phase build; > 0: one-pass stride
        double system_lower[3], system_upper[3]; // exact system bounds in Real (what the harness used for its lattice)
        int32_t use_system_bounds;
        int32_t legacy;                  // legacy API/formulation (Integration1stHalf/2ndHalf, DensitySummation)
        int32_t rank, nranks;            // slab decomposition: one process per GPU
        uint8_t unique_id[128];          // communicator id from sphck_comm_unique_id on rank 0 (nranks > 1)
        int32_t surface_indicator;       // FreeSurfaceIndicationComplexSpatialTemporalCK in the loop
        int32_t observers;               // the case file's pressure probes (ObserverBody + ObservedQuantityRecording)
        double mu_f;                     // > 0: Viscosity closure + ViscousForceWithWallCK
        int32_t transport_velocity;      // KernelGradientIntegral(Corrected)Complex + TransportVelocityCorrectionCK
        int32_t serial_exchange;         // decomposed runs: 1 = plane exchange in line with the dynamics (no overlap)
        int32_t recut_interval;          // decomposed runs: re-balance the cuts every so many advection steps (< 0: default 100)
        int32_t initial_cut_shift;       // decomposed runs: unbalanced start, interior cuts moved by so many planes (test hook)
        int32_t riemann;                 // 0 NoRiemannSolverCK, 1 AcousticRiemannSolverCK, 2 DissipativeRiemannSolverCK
        int32_t kernel_kind;             // 0 Wendland C2, 1 Laguerre-Gauss (tabulated, resetKernel)
        int32_t full_wall;               // decomposed runs: 1 = every rank keeps the whole wall (default: its slab of it)
    };

    const char *sphck_last_error() { return g_error.c_str(); }
    int sphck_comm_unique_id(void *id128) { return sphb200_comm_unique_id(id128); }
    // decomposed runs: slots [begin, begin + count) hold this rank's own particles; stored = own + ghosts
    int sphck_own_range(void *hp, uint64_t *begin, uint64_t *count, uint64_t *stored)
    {
        return guarded([&] {
            BaseParticles &p = ((Handle *)hp)->body(0).getBaseParticles();
            *begin = p.activeBegin();
            *count = p.activeEnd() - p.activeBegin();
            *stored = p.TotalRealParticles();
        });
    }
    // raw copy of `count` slots starting at `begin` (storage order, device element layout: Vecd = 4 floats)
    int sphck_download_raw(void *hp, int which, const char *name, void *out, uint64_t begin, uint64_t count)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            BaseParticles &p = h->body(which).getBaseParticles();
            p.downloadRaw(p.findVariable(name), out, begin, count);
        });
    }
    // raw upload into `count` slots starting at `begin` (storage order, device element layout)
    int sphck_upload_raw(void *hp, int which, const char *name, const void *in, uint64_t begin, uint64_t count)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            SPHBody &b = h->body(which);
            BaseParticles &p = b.getBaseParticles();
            DiscreteVariableBase *v = p.findVariable(name);
            ExecutionInstance &ex = execution_instance();
            const size_t eb = v->deviceElementBytes();
            if (begin + count > p.ParticlesBound()) throw SphError("sphck_upload_raw: range outside the storage");
            if (count) ex.check(sphb200_copy_h2d((char *)v->deviceAddress() + begin * eb, in, count * eb, ex.stream()), "sphb200_copy_h2d");
            b.setPosVolDirty();
            h->acousticTimeStep()->setPrimed(false);
        });
    }
    int sphck_cuts(void *hp, int32_t *out, int capacity)
    {
        return guarded([&] {
            if (!((Handle *)hp)->sim) throw SphError("not a decomposed run");
            DamBreakCK &s = *((Handle *)hp)->sim;
            if (!s.decomposition) throw SphError("not a decomposed run");
            const std::vector<int> &c = s.decomposition->cuts();
            if ((int)c.size() > capacity) throw SphError("capacity too small");
            for (size_t i = 0; i < c.size(); ++i) out[i] = c[i];
        });
    }
    // host-only planning helper (no GPU needed): cuts from a particles-per-plane histogram
    int sphck_plan_slab_cuts(const uint64_t *per_plane, int planes, int nranks, int32_t *cuts_out)
    {
        return guarded([&] {
            std::vector<uint64_t> h(per_plane, per_plane + planes);
            std::vector<int> c = planSlabCuts(h, nranks);
            for (size_t i = 0; i < c.size(); ++i) cuts_out[i] = c[i];
        });
    }

    // host-only: how far a re-balancing may move the cuts in one go (SlabDecomposition::recut)
    // mesh and seam thresholds of a ring-decomposed periodic body (host arithmetic only: no GPU needed)
    int sphck_aligned_periodic_mesh(const double *lower, const double *upper, double cutoff, int dim, sphb200_mesh_t *mesh,
                                    sphb200_seam_t *seam)
    {
        return guarded([&] {
            const Vecd lo((Real)lower[0], (Real)lower[1], (Real)lower[2]), up((Real)upper[0], (Real)upper[1], (Real)upper[2]);
            BoundingBoxd box(lo, up);
            SeamRing ring;
            *mesh = alignedPeriodicMesh(box, Real(cutoff), dim, ring);
            *seam = ring.seam;
        });
    }
    // host-only: the wall planes a rank stores for the fluid planes [X0, X1) (WallSlab::plan). below[x] = wall particles in the
    // planes < x (planes + 1 entries); lo_hi: the planes stored now on entry (hi < lo: none), the planes to store on return;
    // *reload = 1 if the subset has to be loaded
    int sphck_wall_slab_plan(const uint64_t *below, int planes, int depth, int margin, int X0, int X1, uint64_t bound, int32_t *lo_hi,
                             int32_t *reload)
    {
        return guarded([&] {
            int lo = lo_hi[0], hi = lo_hi[1];
            *reload = WallSlab::planWallPlanes(below, planes, depth, margin, X0, X1, (size_t)bound, lo, hi) ? 1 : 0;
            lo_hi[0] = lo, lo_hi[1] = hi;
        });
    }
    int sphck_limit_cut_moves(const int32_t *old_cuts, const int32_t *wanted, int nranks, int32_t *cuts_out)
    {
        return guarded([&] {
            std::vector<int> o(old_cuts, old_cuts + nranks + 1), w(wanted, wanted + nranks + 1);
            std::vector<int> c = limitCutMoves(o, w);
            for (size_t i = 0; i < c.size(); ++i) cuts_out[i] = c[i];
        });
    }

    // fluid_xyz / wall_xyz / wall_normal_xyz may be NULL: the C++ lattice generator and shape normals are used then
    void *sphck_dambreak_create(const sphck_dambreak_options *o, const float *fluid_xyz, uint64_t n_fluid, const float *wall_xyz,
                                const float *wall_normal_xyz, uint64_t n_wall)
    {
        Handle *h = new Handle();
        int rc = guarded([&] {
            execution_instance().setDevice(o->device);
            DamBreakParameters q;
            q.dim = o->dim; q.dp = o->dp;
            q.DL = o->DL; q.DH = o->DH; q.DW = o->DW; q.LL = o->LL; q.LH = o->LH; q.LW = o->LW;
            q.correction = o->correction != 0;
            q.riemann = o->riemann;
            q.kernel_kind = o->kernel_kind;
            q.wall_slabs = o->full_wall == 0;
            q.fused_time_step = o->fused_time_step != 0;
            q.fused_regularization = o->fused_regularization != 0;
            q.sort_interval = o->sort_interval;
            q.legacy = o->legacy != 0;
            q.surface_indicator = o->surface_indicator != 0;
            q.observers = o->observers != 0;
            q.mu_f = o->mu_f;
            q.transport_velocity = o->transport_velocity != 0;
            q.rank = o->rank;
            q.overlap_exchange = o->serial_exchange == 0;
            if (o->recut_interval >= 0) q.recut_interval = o->recut_interval;
            q.initial_cut_shift = o->initial_cut_shift;
            q.nranks = o->nranks > 0 ? o->nranks : 1;
            if (q.nranks > 1)
            {
                // one communicator per context: a decomposed case made earlier in this process leaves its own behind
                if (sphb200_comm_size(execution_instance().ctx()) > 1) sphb200_comm_destroy(execution_instance().ctx());
                execution_instance().check(sphb200_comm_create(execution_instance().ctx(), q.nranks, q.rank, o->unique_id), "sphb200_comm_create");
            }
            std::vector<Vecd> fp, wp, wn;
            BoundingBoxd sb;
            if (o->use_system_bounds)
                sb = BoundingBoxd(Vecd(Real(o->system_lower[0]), Real(o->system_lower[1]), Real(o->system_lower[2])),
                                  Vecd(Real(o->system_upper[0]), Real(o->system_upper[1]), Real(o->system_upper[2])));
            if (fluid_xyz) fp = toVecd(fluid_xyz, n_fluid);
            if (wall_xyz) wp = toVecd(wall_xyz, n_wall);
            if (wall_normal_xyz) wn = toVecd(wall_normal_xyz, n_wall);
            h->sim.reset(new DamBreakCK(q, fluid_xyz ? &fp : nullptr, wall_xyz ? &wp : nullptr, wall_normal_xyz ? &wn : nullptr,
                                        o->use_system_bounds ? &sb : nullptr));
            if (o->relation_stride >= 0)
            {
                h->sim->water_block_inner->fixed_stride_ = (uint32_t)o->relation_stride;
                h->sim->water_wall_contact->fixed_stride_ = (uint32_t)o->relation_stride;
            }
        });
        if (rc)
        {
            delete h;
            return nullptr;
        }
        return h;
    }
    struct sphck_taylor_green_options
    {
        int32_t dim;
        double dp, L, U_f;
        int32_t fused_time_step, fused_regularization, sort_interval, device, relation_stride;
        double system_lower[3], system_upper[3];
        int32_t use_system_bounds;
        double mu_f;                 // > 0: viscous (ViscousForceInnerCK)
        int32_t transport_velocity;  // KernelGradientIntegralInner + TransportVelocityCorrectionCK
        double x_scale;              // box [0, x_scale L] x [0, L]^(dim-1) (0 is read as 1)
    };
    // xyz / vel_xyz may be NULL: lattice and analytic initial condition are generated by the C++ case then
    void *sphck_taylor_green_create(const sphck_taylor_green_options *o, const float *xyz, const float *vel_xyz, uint64_t n)
    {
        Handle *h = new Handle();
        int rc = guarded([&] {
            execution_instance().setDevice(o->device);
            TaylorGreenParameters q;
            q.dim = o->dim; q.dp = o->dp; q.L = o->L; q.U_f = o->U_f;
            q.x_scale = o->x_scale > 0 ? o->x_scale : 1.0;
            q.fused_time_step = o->fused_time_step != 0;
            q.fused_regularization = o->fused_regularization != 0;
            q.sort_interval = o->sort_interval;
            q.mu_f = o->mu_f;
            q.transport_velocity = o->transport_velocity != 0;
            std::vector<Vecd> pos, vel;
            if (xyz) pos = toVecd(xyz, n);
            if (vel_xyz) vel = toVecd(vel_xyz, n);
            BoundingBoxd sb;
            if (o->use_system_bounds)
                sb = BoundingBoxd(Vecd(Real(o->system_lower[0]), Real(o->system_lower[1]), Real(o->system_lower[2])),
                                  Vecd(Real(o->system_upper[0]), Real(o->system_upper[1]), Real(o->system_upper[2])));
            h->tg.reset(new TaylorGreenCK(q, xyz ? &pos : nullptr, vel_xyz ? &vel : nullptr, o->use_system_bounds ? &sb : nullptr));
            if (o->relation_stride >= 0) h->tg->water_block_inner->fixed_stride_ = (uint32_t)o->relation_stride;
        });
        if (rc)
        {
            delete h;
            return nullptr;
        }
        return h;
    }
    // Ring-decomposed Taylor-Green run (periodic along x through the slab exchange, config 4 on N GPUs): this rank's
    // particles with their global numbers. nranks == 1 makes a ring of one slab on a communicator without NCCL.
    void *sphck_taylor_green_create_ring(const sphck_taylor_green_options *o, const float *xyz, const float *vel_xyz,
                                         const uint32_t *global_ids, uint64_t n, int32_t rank, int32_t nranks, const void *unique_id)
    {
        Handle *h = new Handle();
        int rc = guarded([&] {
            if (!xyz || !vel_xyz || !global_ids) throw SphError("ring run: positions, velocities and global ids are required");
            execution_instance().setDevice(o->device);
            sphb200_context_t *ctx = execution_instance().ctx();
            if (g_ring_handles == 0)
            {
                // a communicator left behind by an earlier decomposed case of this process is not a ring: start afresh
                if (sphb200_comm_size(ctx) > 1) sphb200_comm_destroy(ctx);
                if (nranks > 1) execution_instance().check(sphb200_comm_create(ctx, nranks, rank, unique_id), "sphb200_comm_create");
                else execution_instance().check(sphb200_comm_create_self(ctx), "sphb200_comm_create_self");
                execution_instance().check(sphb200_comm_set_ring(ctx, 1), "sphb200_comm_set_ring");
            }
            else if (sphb200_comm_size(ctx) != (nranks > 1 ? nranks : 1))
                throw SphError("ring run: another ring run of a different size is alive on this context");
            ++g_ring_handles;
            h->owns_comm = true;
            TaylorGreenParameters q;
            q.dim = o->dim; q.dp = o->dp; q.L = o->L; q.U_f = o->U_f;
            q.x_scale = o->x_scale > 0 ? o->x_scale : 1.0;
            q.fused_time_step = o->fused_time_step != 0;
            q.fused_regularization = o->fused_regularization != 0;
            q.sort_interval = 0;
            q.mu_f = o->mu_f;
            q.transport_velocity = o->transport_velocity != 0;
            q.ring = true; q.rank = rank; q.nranks = nranks;
            std::vector<Vecd> pos = toVecd(xyz, n), vel = toVecd(vel_xyz, n);
            std::vector<UnsignedInt> ids(global_ids, global_ids + n);
            BoundingBoxd sb;
            if (o->use_system_bounds)
                sb = BoundingBoxd(Vecd(Real(o->system_lower[0]), Real(o->system_lower[1]), Real(o->system_lower[2])),
                                  Vecd(Real(o->system_upper[0]), Real(o->system_upper[1]), Real(o->system_upper[2])));
            h->tg.reset(new TaylorGreenCK(q, &pos, &vel, o->use_system_bounds ? &sb : nullptr, &ids));
            if (o->relation_stride >= 0) h->tg->water_block_inner->fixed_stride_ = (uint32_t)o->relation_stride;
        });
        if (rc)
        {
            delete h;
            return nullptr;
        }
        return h;
    }
    void sphck_destroy(void *hp) { delete (Handle *)hp; }

    uint64_t sphck_count(void *hp, int which)
    {
        Handle *h = (Handle *)hp;
        if (h->tg) return which ? 0 : h->tg->water_block.getBaseParticles().hostSyncCount();
        return body(h, which).TotalRealParticles();
    }
    uint64_t sphck_launches(void *) { return execution_instance().launches(); }
    uint64_t sphck_device_allocations(void *) { return sphb200_device_allocation_count(); }
    // BodyStatesRecordingToVtpCK::writeToFile of the dam-break case into `folder` (one .vtp per body); returns bytes synchronised
    int sphck_record_states(void *hp, const char *folder, uint64_t *bytes_synchronized)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            if (!h->sim) throw SphError("record_states: dam-break cases only");
            h->sim->recordStates(folder);
            if (bytes_synchronized) *bytes_synchronized = h->sim->body_states_recording->bytesSynchronized();
        });
    }
    int sphck_synchronize(void *) { return guarded([] { execution_instance().synchronize(); }); }
    // SPHB200_STEP_TRACE=1: print the per-stage table gathered so far (over `steps` advection steps) and start afresh
    int sphck_step_trace_report(uint64_t steps)
    {
        return guarded([&] {
            if (!StepTrace::enabled()) return;
            StepTrace::get().report(std::cerr, steps);
            StepTrace::get().clear();
        });
    }

    // mesh / kernel PODs as computed by the host layer (parity of the host arithmetic with the oracle's inputs)
    int sphck_mesh(void *hp, int which, sphb200_mesh_t *out)
    {
        return guarded([&] { *out = body((Handle *)hp, which).getCellLinkedList().mesh_; });
    }
    int sphck_kernel(void *hp, sphb200_kernel_t *out)
    {
        return guarded([&] { *out = ((Handle *)hp)->relation(0).kernel_; });
    }

    // one name per dynamics object of the case (tests drive them one by one)
    int sphck_exec(void *hp, const char *op_c, double a0, double *result)
    {
        Handle *h = (Handle *)hp;
        std::string op(op_c);
        double r = 0;
        if (h->tg)
        {
            TaylorGreenCK &s = *h->tg;
            int rc = guarded([&] {
                if (op == "initialize") s.initialize();
                else if (op == "step_outer") r = s.stepOuter();
                else if (op == "run_outer")
                {
                    long n = (long)a0, total = 0;
                    for (long k = 0; k < n; ++k) total += s.stepOuter();
                    r = (double)total;
                }
                else if (op == "cell_list_fluid") s.water_cell_linked_list->exec();
                else if (op == "rebuild")
                {
                    // the configuration update without the relation search: cell list (ring: migration + ghost planes) + images
                    if (s.decomposition) s.decomposition->rebuild();
                    else s.water_cell_linked_list->exec();
                    for (auto &pc : s.periodic_condition) pc->ghost_creation_.exec();
                }
                else if (op == "periodic_bounding") { for (auto &pc : s.periodic_condition) pc->bounding_.exec(); }
                else if (op == "ghost_creation") { for (auto &pc : s.periodic_condition) pc->ghost_creation_.exec(); }
                else if (op == "ghost_update") s.periodic_condition[0]->ghost_update_.exec();
                else if (op == "ghost_particles") { s.periodic_condition[0]->images().ensure(); r = (double)s.periodic_condition[0]->images().ghostParticles(); }
                else if (op == "plane_ghost_particles") r = s.decomposition ? (double)s.decomposition->ghostParticles() : 0.0;
                else if (op == "handed_over") r = s.decomposition ? (double)s.decomposition->migratedOut() : 0.0;
                else if (op == "box_planes") r = (double)s.seam_ring.box_planes();
                else if (op == "relations") s.water_block_update_inner_relation->exec();
                else if (op == "update_configuration") s.updateConfiguration(a0 != 0.0);
                else if (op == "sort") { s.particle_sort->exec(); s.fluid_acoustic_time_step->setPrimed(false); }
                else if (op == "density_summation") s.fluid_density_summation->exec();
                else if (op == "density_regularization") s.fluid_density_regularization->exec();
                else if (op == "advection_setup") { s.water_advection_step_setup->exec(); s.volume_ghost_update->exec(); }
                else if (op == "update_position") s.water_update_particle_position->exec();
                else if (op == "viscous_force") { if (!s.viscous_force) throw SphError("case built without viscosity"); s.viscous_force->exec(); }
                else if (op == "kernel_gradient_integral") { if (!s.kernel_gradient_integral) throw SphError("case built without transport velocity"); s.kernel_gradient_integral->exec(); }
                else if (op == "transport_velocity_correction") { if (!s.transport_velocity_correction) throw SphError("case built without transport velocity"); s.transport_velocity_correction->exec(); }
                else if (op == "advection_dt") r = s.fluid_advection_time_step->exec();
                else if (op == "advection_dt_reduced") r = s.fluid_advection_time_step->ReducedValue();
                else if (op == "acoustic_dt") r = s.fluid_acoustic_time_step->exec();
                else if (op == "acoustic_dt_reduced") r = s.fluid_acoustic_time_step->ReducedValue();
                else if (op == "acoustic_dt_unprime") s.fluid_acoustic_time_step->setPrimed(false);
                else if (op == "acoustic1") s.fluid_acoustic_step_1st_half->exec(Real(a0));
                else if (op == "acoustic2") s.fluid_acoustic_step_2nd_half->exec(Real(a0));
                else if (op == "energy") r = s.record_total_kinetic_energy->exec();
                else if (op == "physical_time") r = s.physical_time;
                else if (op == "acoustic_steps") r = (double)s.acoustic_steps;
                else if (op == "outer_steps") r = (double)s.number_of_iterations;
                else if (op == "last_acoustic_dt") r = s.last_acoustic_dt;
                else if (op == "inner_total") r = (double)s.water_block_inner->total_;
                else if (op == "inner_stride") r = (double)s.water_block_inner->fixed_stride_;
                else if (op == "inner_max_count") r = (double)s.water_block_inner->max_count_;
                else if (op == "set_relation_stride") s.water_block_inner->fixed_stride_ = (uint32_t)a0;
                else throw SphError("unknown op '" + op + "'");
            });
            if (result) *result = r;
            return rc;
        }
        DamBreakCK &s = *h->sim;
        int rc = guarded([&] {
            if (op == "initialize") s.initialize();
            else if (op == "step_outer") r = s.stepOuter();
            else if (op == "run_outer")
            {
                long n = (long)a0, total = 0;
                for (long k = 0; k < n; ++k) total += s.stepOuter();
                r = (double)total;
            }
            else if (op == "set_sort_interval") { s.q_.sort_interval = (int)a0; s.q_.recut_interval = (int)a0; } // ParticleSortCK / re-cut cadence
            else if (op == "configuration_before_dynamics")
                s.configuration_update = a0 != 0.0 ? DamBreakCK::ConfigurationUpdate::BeforeDynamics : DamBreakCK::ConfigurationUpdate::AfterDynamics;
            else if (op == "gravity") s.constant_gravity->exec();
            else if (op == "cell_list_fluid") s.water_cell_linked_list->exec();
            else if (op == "cell_list_wall") s.wall_cell_linked_list->exec();
            else if (op == "relations")
            {
                if (s.water_wall_complex) s.water_wall_complex->updateConfiguration();
                else s.water_block_update_complex_relation->exec();
            }
            else if (op == "sort") { s.particle_sort->exec(); s.fluid_acoustic_time_step->setPrimed(false); }
            else if (op == "density_summation") s.fluid_density_summation->exec();
            else if (op == "density_regularization") { if (s.fluid_density_regularization) s.fluid_density_regularization->exec(); }
            else if (op == "advection_setup") { if (s.water_advection_step_setup) s.water_advection_step_setup->exec(); }
            else if (op == "update_position") { if (s.water_update_particle_position) s.water_update_particle_position->exec(); }
            else if (op == "advection_dt") r = s.fluid_advection_time_step->exec();
            else if (op == "advection_dt_reduced") r = s.advection_reduced_value();
            else if (op == "acoustic_dt") r = s.fluid_acoustic_time_step->exec();
            else if (op == "acoustic_dt_reduced") r = s.fluid_acoustic_time_step->ReducedValue();
            else if (op == "acoustic_dt_unprime") s.fluid_acoustic_time_step->setPrimed(false);
            else if (op == "acoustic1") s.fluid_acoustic_step_1st_half->exec(Real(a0));
            else if (op == "acoustic2") s.fluid_acoustic_step_2nd_half->exec(Real(a0));
            else if (op == "linear_correction") { if (s.fluid_linear_correction_matrix) s.fluid_linear_correction_matrix->exec(); }
            else if (op == "surface_indication")
            {
                if (!s.fluid_boundary_indicator) throw SphError("case built without surface_indicator");
                s.fluid_boundary_indicator->exec();
            }
            else if (op == "viscous_force") { if (!s.fluid_viscous_force) throw SphError("case built without viscosity"); s.fluid_viscous_force->exec(); }
            else if (op == "kernel_gradient_integral") { if (!s.kernel_gradient_integral) throw SphError("case built without transport velocity"); s.kernel_gradient_integral->exec(); }
            else if (op == "transport_velocity_correction") { if (!s.transport_correction) throw SphError("case built without transport velocity"); s.transport_correction->exec(); }
            else if (op == "observer_relation") { if (s.fluid_observer_contact_relation) s.fluid_observer_contact_relation->exec(); }
            else if (op == "observe_pressure")
            {
                if (!s.fluid_observer_pressure) throw SphError("case built without observers");
                s.fluid_observer_pressure->writeToFile(s.number_of_iterations);
            }
            else if (op == "probe_records") r = s.fluid_observer_pressure ? (double)s.fluid_observer_pressure->records().size() : 0.0;
            else if (op == "probe_count") r = s.fluid_observer_pressure ? (double)s.fluid_observer_pressure->NumberOfObservedQuantity() : 0.0;
            else if (op == "energy") r = s.record_water_mechanical_energy->exec();
            else if (op == "physical_time") r = s.physical_time;
            else if (op == "acoustic_steps") r = (double)s.acoustic_steps;
            else if (op == "outer_steps") r = (double)s.number_of_iterations;
            else if (op == "last_acoustic_dt") r = s.last_acoustic_dt;
            else if (op == "ghost_particles") r = s.decomposition ? (double)s.decomposition->ghostParticles() : 0.0;
            else if (op == "recuts") r = s.decomposition ? (double)s.decomposition->recuts() : 0.0;
            else if (op == "rebuild") { if (s.decomposition) s.decomposition->update(); else s.water_cell_linked_list->exec(); }
            else if (op == "wall_slab_loads") r = s.wall_slab ? (double)s.wall_slab->loads() : 0.0;
            else if (op == "wall_global_particles") r = s.wall_slab ? (double)s.wall_slab->globalParticles() : (double)s.wall_boundary.TotalRealParticles();
            else if (op == "rebuild_host_syncs") r = s.decomposition ? (double)s.decomposition->hostSyncs() : 0.0;
            else if (op == "inner_total") r = (double)s.water_block_inner->total_;
            else if (op == "inner_pairs" || op == "contact_pairs")
            {
                // neighbour-list entries of the own particles (sum of the row counts), read back once
                RelationBase &rel = op == "inner_pairs" ? (RelationBase &)*s.water_block_inner : (RelationBase &)*s.water_wall_contact;
                BaseParticles &p = s.water_block.getBaseParticles();
                const size_t b = p.activeBegin(), e = p.activeEnd();
                std::vector<uint32_t> c(e - b);
                ExecutionInstance &ex = execution_instance();
                if (e > b) ex.check(sphb200_copy_d2h(c.data(), rel.count_.get<uint32_t>() + b, (e - b) * sizeof(uint32_t), ex.stream()), "sphb200_copy_d2h");
                ex.synchronize();
                uint64_t tot = 0;
                for (uint32_t v : c) tot += v;
                r = (double)tot;
            }
            else if (op == "inner_stride") r = (double)s.water_block_inner->fixed_stride_;
            else if (op == "inner_max_count") r = (double)s.water_block_inner->max_count_;
            else if (op == "set_relation_stride")
            {
                s.water_block_inner->fixed_stride_ = (uint32_t)a0;
                s.water_wall_contact->fixed_stride_ = (uint32_t)a0;
            }
            else throw SphError("unknown op '" + op + "'");
        });
        if (result) *result = r;
        return rc;
    }

    // phase-granular 1st half (parity tests against the reference's per-phase kernels)
    int sphck_acoustic1_phase(void *hp, int phase /*0 initialize, 1 interact+update*/, double dt)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            auto *ph = h->tg ? static_cast<fluid_dynamics::AcousticStep1stHalfPhases *>(h->tg->fluid_acoustic_step_1st_half.get())
                             : dynamic_cast<fluid_dynamics::AcousticStep1stHalfPhases *>(h->sim->fluid_acoustic_step_1st_half.get());
            if (!ph) throw SphError("1st half does not expose phases");
            if (phase == 0) ph->deviceInitialize(Real(dt));
            else ph->deviceInteractAndUpdate(Real(dt));
        });
    }

    // kind: 0 Real, 1 Vecd (3 floats per particle), 2 UnsignedInt, 3 Matd (9 floats). Reference particle order.
    int sphck_download(void *hp, int which, const char *name, int kind, void *out)
    {
        return guarded([&] {
            BaseParticles &p = body((Handle *)hp, which).getBaseParticles();
            if (kind == 0) p.download(p.getVariableByName<Real>(name), (Real *)out);
            else if (kind == 1) p.download(p.getVariableByName<Vecd>(name), (Vecd *)out);
            else if (kind == 2) p.download(p.getVariableByName<UnsignedInt>(name), (UnsignedInt *)out);
            else if (kind == 4) p.download(p.getVariableByName<int>(name), (int *)out);
            else p.download(p.getVariableByName<Matd>(name), (Matd *)out);
        });
    }
    int sphck_upload(void *hp, int which, const char *name, int kind, const void *in)
    {
        return guarded([&] {
            SPHBody &b = body((Handle *)hp, which);
            BaseParticles &p = b.getBaseParticles();
            if (kind == 0) p.upload(p.getVariableByName<Real>(name), (const Real *)in);
            else if (kind == 1) p.upload(p.getVariableByName<Vecd>(name), (const Vecd *)in);
            else if (kind == 2) p.upload(p.getVariableByName<UnsignedInt>(name), (const UnsignedInt *)in);
            else if (kind == 4) p.upload(p.getVariableByName<int>(name), (const int *)in);
            else p.upload(p.getVariableByName<Matd>(name), (const Matd *)in);
            b.setPosVolDirty();
            ((Handle *)hp)->acousticTimeStep()->setPrimed(false);
        });
    }
    int sphck_has_variable(void *hp, int which, const char *name) { return body((Handle *)hp, which).getBaseParticles().hasVariable(name) ? 1 : 0; }

    // raw (slot-order) device pointer of a variable, for harness-side pinned-memory transfers
    void *sphck_device_pointer(void *hp, int which, const char *name, int kind)
    {
        void *out = nullptr;
        guarded([&] {
            BaseParticles &p = body((Handle *)hp, which).getBaseParticles();
            if (kind == 0) out = p.deviceData<Real>(name);
            else if (kind == 1) out = p.deviceData<Vecd>(name);
            else if (kind == 2) out = p.deviceData<UnsignedInt>(name);
            else out = p.deviceData<Matd>(name);
        });
        return out;
    }

    // ---- overlapped transfers (HostTransferPipeline of the fluid body); names are separated by ',' ----
    int sphck_pipeline_create(void *hp, const char *inputs, const char *outputs)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            BaseParticles &p = h->body(0).getBaseParticles();
            h->pipeline.reset(new HostTransferPipeline(p));
            // decomposed bodies exchange this rank's own slots raw (storage order, device layout)
            h->pipeline->setRawOwnSlots(h->sim && h->sim->decomposition);
            auto add = [&](const char *list, bool input) {
                std::string s(list ? list : ""), name;
                size_t pos = 0;
                while (pos <= s.size())
                {
                    size_t e = s.find(',', pos);
                    if (e == std::string::npos) e = s.size();
                    name = s.substr(pos, e - pos);
                    pos = e + 1;
                    if (name.empty()) continue;
                    DiscreteVariableBase *v = p.findVariable(name);
                    if (auto *r = dynamic_cast<DiscreteVariable<Real> *>(v)) input ? h->pipeline->addInput(r) : h->pipeline->addOutput(r);
                    else if (auto *q = dynamic_cast<DiscreteVariable<Vecd> *>(v)) input ? h->pipeline->addInput(q) : h->pipeline->addOutput(q);
                    else throw SphError("pipeline: only Real and Vecd variables are supported: " + name);
                }
            };
            add(inputs, true);
            add(outputs, false);
        });
    }
    int sphck_pipeline_stage_uploads(void *hp, const void *const *pinned_host)
    {
        return guarded([&] { ((Handle *)hp)->pipeline->stageUploads(pinned_host); });
    }
    int sphck_pipeline_commit_uploads(void *hp)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            h->pipeline->commitUploads();
            h->body(0).setPosVolDirty();
            h->acousticTimeStep()->setPrimed(false);
        });
    }
    int sphck_pipeline_stage_downloads(void *hp, void *const *pinned_host)
    {
        return guarded([&] { ((Handle *)hp)->pipeline->stageDownloads(pinned_host); });
    }
    int sphck_pipeline_synchronize(void *hp)
    {
        return guarded([&] { ((Handle *)hp)->pipeline->synchronize(); });
    }
    int sphck_pipeline_bytes(void *hp, uint64_t *in_bytes, uint64_t *out_bytes)
    {
        return guarded([&] {
            *in_bytes = ((Handle *)hp)->pipeline->inputBytes();
            *out_bytes = ((Handle *)hp)->pipeline->outputBytes();
        });
    }

    // recorded probe rows (ObservedQuantityRecording::records()): times[rows], values[rows * probes]
    int sphck_probe_records(void *hp, double *times, double *values, uint64_t capacity_rows)
    {
        return guarded([&] {
            Handle *h = (Handle *)hp;
            if (!h->sim || !h->sim->fluid_observer_pressure) throw SphError("case built without observers");
            auto &rec = *h->sim->fluid_observer_pressure;
            const auto &rows = rec.records();
            for (size_t r = 0; r < rows.size() && r < capacity_rows; ++r)
            {
                times[r] = rec.recordedTimes()[r];
                for (size_t k = 0; k < rows[r].size(); ++k) values[r * rows[r].size() + k] = rows[r][k];
            }
        });
    }

    // cell-linked list (cell_offset[cells + 1]) and relation CSR in reference ids.
    int sphck_cell_offsets(void *hp, int which, uint32_t *out, uint64_t count)
    {
        return guarded([&] {
            CellLinkedList &cl = body((Handle *)hp, which).getCellLinkedList();
            ExecutionInstance &ex = execution_instance();
            ex.check(sphb200_copy_d2h(out, cl.cell_offset_.get(), count * sizeof(uint32_t), ex.stream()), "sphb200_copy_d2h");
            ex.synchronize();
        });
    }
    // relation 0 inner, 1 contact. Call with index == NULL to get the sizes first.
    int sphck_export_csr(void *hp, int relation, uint32_t *offset, uint32_t *index, uint64_t index_capacity, uint64_t *total)
    {
        return guarded([&] {
            RelationBase &r = ((Handle *)hp)->relation(relation);
            std::vector<uint32_t> off, idx;
            r.exportCSR(off, idx);
            if (total) *total = idx.size();
            if (offset) std::memcpy(offset, off.data(), off.size() * sizeof(uint32_t));
            if (index)
            {
                if (idx.size() > index_capacity) throw SphError("sphck_export_csr: index capacity too small");
                std::memcpy(index, idx.data(), idx.size() * sizeof(uint32_t));
            }
        });
    }
}
