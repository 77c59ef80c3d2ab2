// sphinxsys_ck/periodic_images.h — periodic boundary conditions by ghost (image) particles.
//
// Included by configuration.h right after CellLinkedList (it needs that class complete); not a stand-alone header.
//
// Reference (relative to /root/reference/src/shared/particle_dynamics/general_dynamics/domian_bouding):
//   PeriodicAlongAxis ........................... domain_bounding.h:48-66
//   PeriodicBounding (bounding_) ................ domain_bounding.h:91-140
//   Ghost<PeriodicAlongAxis> .................... ghost_bounding.h:41-62, ghost_bounding.cpp:6-26
//   PeriodicConditionUsingGhostParticles ........ ghost_bounding.h:64-137, ghost_bounding.cpp:28-138
//       bounding_ / ghost_creation_ / ghost_update_, used as in tests/2d_examples/test_2d_throat/throat.cpp:137-278
//   PeriodicConditionUsingCellLinkedList ........ domain_bounding.h:142-175, domain_bounding.cpp:18-65 (same neighbour sets:
//       ghost list entries (source index, translated position) instead of ghost particles)
// The reference has these on the TBB path only; the CK/device spelling below is ours, the semantics are the reference's.
//
// Storage: the images of all periodic axes of a body live cell ordered BEHIND its real particles
// (slots [n_real, n_real + n_ghost)), each with the translated position and a copy of every variable of its source;
// they have their own cell-linked list on the body's mesh, which the relation search walks after the real one.
// Dynamics run on the active range [0, n_real) and read ghosts like any other neighbour.
#ifndef SPHINXSYS_CK_PERIODIC_IMAGES_H
#define SPHINXSYS_CK_PERIODIC_IMAGES_H

namespace SPH
{
constexpr int xAxis = 0, yAxis = 1, zAxis = 2;

struct PeriodicAlongAxis
{
  protected:
    BoundingBoxd bounding_bounds_;
    int axis_;
    Vecd periodic_translation_;

  public:
    PeriodicAlongAxis(const BoundingBoxd &bounding_bounds, int axis) : bounding_bounds_(bounding_bounds), axis_(axis)
    {
        periodic_translation_[axis] = bounding_bounds.upper_[axis] - bounding_bounds.lower_[axis];
    }
    BoundingBoxd getBoundingBox() const { return bounding_bounds_; }
    int getAxis() const { return axis_; }
    Vecd getPeriodicTranslation() const { return periodic_translation_; }
};

template <> class Ghost<PeriodicAlongAxis> : public PeriodicAlongAxis
{
    bool is_particles_reserved_ = false;

  public:
    Ghost(const BoundingBoxd &bounding_bounds, int axis) : PeriodicAlongAxis(bounding_bounds, axis) {}
    // ghost_bounding.cpp:19-26 reserves 2 x ghost_width(4) x face extent / dp per side in 2-D; here the face is the
    // product of the other extents (each widened by the ghost width so edge and corner images fit)
    size_t reserveSize(Real dp, int dim) const
    {
        double face = 1.0;
        for (int d = 0; d < dim; ++d)
            if (d != axis_) face *= (double(bounding_bounds_.upper_[d]) - double(bounding_bounds_.lower_[d])) / double(dp) + 8.0;
        return (size_t)std::ceil(2.0 * 4.0 * face);
    }
    void setReserved() { is_particles_reserved_ = true; }
    void checkParticlesReserved() const
    {
        if (!is_particles_reserved_)
            throw SphError("Ghost<PeriodicAlongAxis>: ghost particles are not reserved (use generateParticlesWithReserve)");
    }
};

class PeriodicImages
{
    SPHBody &body_;
    sphb200_periodic_t box_;
    int armed_axes_ = 0;   // axes whose ghost_creation_ ran since the last cell-linked-list update
    bool valid_ = false;   // ghosts exist for the armed axes
    uint32_t n_real_ = 0, n_ghost_ = 0;
    // slab-decomposed periodic runs (SlabDecomposition with a ring): the images of the remaining axes are made for every
    // STORED particle (own + the ghost planes of the neighbour slabs), dynamics keep running on the own range
    bool decomposed_ = false;
    uint32_t own_begin_ = 0, own_end_ = 0;
    DeviceBuffer tmp_pos_, tmp_src_, ghost_src_, cell_offset_, particle_index_;
    size_t work_capacity_ = 0; // images the work arrays hold (grown with head-room: see ensure())

  public:
    explicit PeriodicImages(SPHBody &body) : body_(body)
    {
        std::memset(&box_, 0, sizeof(box_));
        box_.cutoff = body.getSPHAdaptation().CutOffRadius(); // cut_off_radius_max_, domain_bounding.h:125
        n_real_ = (uint32_t)body.getBaseParticles().hostSyncCount();
        size_t cells = body.getCellLinkedList().total_cells_;
        cell_offset_.reset((cells + 2) * sizeof(uint32_t));
    }
    void addAxis(const PeriodicAlongAxis &a)
    {
        const int k = a.getAxis();
        BoundingBoxd b = a.getBoundingBox();
        box_.lower[k] = b.lower_[k];
        box_.upper[k] = b.upper_[k];
        box_.axes |= 1 << k;
    }
    // SlabDecomposition::rebuild(): `stored` particles (own + x ghost planes) in cell order, own slots [begin, end)
    void setStoredRange(uint32_t stored, uint32_t own_begin, uint32_t own_end)
    {
        decomposed_ = true;
        n_real_ = stored;
        own_begin_ = own_begin;
        own_end_ = own_end;
        invalidate();
    }
    uint32_t realParticles() const { return n_real_; }
    uint32_t ghostParticles() const { return valid_ ? n_ghost_ : 0; }
    const uint32_t *ghostSource() const { return ghost_src_.get<uint32_t>(); }
    bool armed() const { return armed_axes_ != 0; }

    // PeriodicBounding::exec for one axis (all real particles; only those outside the box move)
    void bounding(int axis)
    {
        BaseParticles &p = body_.getBaseParticles();
        sphb200_periodic_t b = box_;
        b.axes = 1 << axis;
        // decomposed runs: the own particles only (ghost planes beyond the seam are outside the box by construction)
        const uint32_t first = decomposed_ ? own_begin_ : 0, count = decomposed_ ? own_end_ - own_begin_ : n_real_;
        SPHCK_CALL(sphb200_periodic_bounding, &b, (sphb200_vec4_t *)p.deviceData<Vecd>("Position") + first, count, execution_instance().stream());
        body_.setPosVolDirty();
    }
    // ghost_creation_.exec() of one axis: the images themselves are made once, for all armed axes together, when
    // they are first needed (ensure()), which yields the same set as the reference's axis-by-axis creation
    void arm(int axis)
    {
        armed_axes_ |= 1 << axis;
        valid_ = false;
    }
    // UpdateCellLinkedList::exec() ran: the storage order changed, ghosts are gone until ghost_creation_ runs again
    void invalidate()
    {
        armed_axes_ = 0;
        valid_ = false;
        n_ghost_ = 0;
    }
    sphb200_cell_list_t listView() const
    {
        sphb200_cell_list_t v;
        v.cell_offset = cell_offset_.get<uint32_t>();
        v.particle_index = particle_index_.get<uint32_t>();
        v.sorted_pos = nullptr;
        return v;
    }
    void ensure()
    {
        if (valid_ || !armed_axes_) return;
        ExecutionInstance &ex = execution_instance();
        BaseParticles &p = body_.getBaseParticles();
        CellLinkedList &cl = body_.getCellLinkedList();
        if (!body_.isCellOrdered()) throw SphError("periodic ghost creation needs the body's cell-linked list (UpdateCellLinkedList first)");
        const uint32_t capacity = (uint32_t)(p.ParticlesBound() - n_real_);
        // The room behind the real particles changes by a few slots every step (decomposed runs: the ghost planes do), and
        // every new maximum used to re-allocate all four work arrays — four cudaFree / cudaMalloc pairs per advection step in
        // the 256^3 ring, each synchronising the device (bench.py `config4.device_allocations_per_step_rank0`). They are
        // sized with a quarter of head-room now: allocation stops after the first steps.
        if (capacity > work_capacity_)
        {
            work_capacity_ = (size_t)capacity + capacity / 4 + 4096;
            tmp_pos_.ensure((work_capacity_ + 1) * 16);
            tmp_src_.ensure((work_capacity_ + 1) * 4);
            ghost_src_.ensure((work_capacity_ + 1) * 4);
            particle_index_.ensure((std::max<size_t>(work_capacity_, cl.total_cells_) + 2) * 4);
        }
        sphb200_periodic_t b = box_;
        b.axes = armed_axes_;
        uint32_t count = 0;
        int rc = 0;
        SPHCK_STAGE("    images: enumerate", rc = sphb200_periodic_images(ex.ctx(), &b, (const sphb200_vec4_t *)p.deviceData<Vecd>("Position"), n_real_,
                                                                        tmp_pos_.get<sphb200_vec4_t>(), tmp_src_.get<uint32_t>(), capacity, &count, ex.stream()));
        if (rc == SPHB200_E_CAPACITY)
            throw SphError("ghost particles exceed the reserve: " + std::to_string(count) + " > " + std::to_string(capacity)); // checkWithinGhostSize
        ex.check(rc, "sphb200_periodic_images");
        n_ghost_ = count;
        p.setTotalRealParticles((size_t)n_real_ + n_ghost_);
        if (decomposed_) p.setActiveRange(own_begin_, own_end_);
        else p.setActiveRange(0, n_real_);
        // images into the cell order of their own list: translated positions straight into the tail of Position
        {
            char *pos_tail = (char *)p.deviceData<Vecd>("Position") + (size_t)n_real_ * 16;
            void *dst[2] = {pos_tail, ghost_src_.get()};
            const void *src[2] = {tmp_pos_.get(), tmp_src_.get()};
            uint32_t eb[2] = {16, 4};
            SPHCK_STAGE("    images: cell list", SPHCK_CALL(sphb200_cell_list_build_reorder, &cl.mesh_, tmp_pos_.get<sphb200_vec4_t>(), n_ghost_, tmp_src_.get<uint32_t>(),
                                                             listView(), 2, dst, src, eb, ex.stream()));
        }
        // every other stored variable: copy of the source particle (BaseParticles::updateGhostParticle)
        if (n_ghost_)
        {
            std::vector<void *> dst;
            std::vector<const void *> src;
            std::vector<uint32_t> eb;
            for (DiscreteVariableBase *v : p.reorderedVariables())
            {
                if (v->Name() == "Position") continue;
                eb.push_back(v->deviceElementBytes());
                src.push_back(v->deviceAddress());
                dst.push_back((char *)v->deviceAddress() + (size_t)n_real_ * v->deviceElementBytes());
            }
            SPHCK_CALL(sphb200_gather_multi, (int)dst.size(), dst.data(), src.data(), eb.data(), ghost_src_.get<uint32_t>(), n_ghost_, ex.stream());
        }
        valid_ = true;
        body_.setPosVolDirty();
    }
    // ghost_update_: copy named variables (all stored ones when `names` is empty) from the sources to their images.
    // Positions are NOT touched (images keep their translated position until the next creation); the derived gather
    // records are patched in place: Vol of PosVol, VolRef of PosVolRef, (Vol, v) of PosVolVel.
    void update(const std::vector<std::string> &names)
    {
        ensure();
        if (!valid_ || n_ghost_ == 0) return;
        ExecutionInstance &ex = execution_instance();
        BaseParticles &p = body_.getBaseParticles();
        auto copy = [&](DiscreteVariableBase *v, uint32_t offset, uint32_t bytes) {
            SPHCK_CALL(sphb200_ghost_copy, v->deviceAddress(), v->deviceElementBytes(), offset, bytes, ghost_src_.get<uint32_t>(), n_real_,
                       n_ghost_, ex.stream());
        };
        auto one = [&](const std::string &nm) {
            if (nm == "Position") return;
            if (!p.hasVariable(nm)) throw SphError("ghost update: the variable '" + nm + "' is not registered");
            DiscreteVariableBase *v = p.findVariable(nm);
            if (body_.recordsDirty() && (nm == "PosVol" || nm == "PosVolRef" || nm == "PosVolVel")) return; // repacked from the primaries anyway
            if (nm == "PosVol" || nm == "PosVolRef") copy(v, 12, 4);
            else if (nm == "PosVolVel") copy(v, 12, 16);
            else copy(v, 0, v->deviceElementBytes());
        };
        if (names.empty())
        {
            for (DiscreteVariableBase *v : p.reorderedVariables()) one(v->Name());
            for (const char *nm : {"PosVol", "PosVolRef", "PosVolVel"})
                if (p.hasVariable(nm)) one(nm);
        }
        else
            for (const std::string &nm : names) one(nm);
    }
};

inline PeriodicImages &SPHBody::definePeriodicImages()
{
    if (!periodic_images_) periodic_images_.reset(new PeriodicImages(*this));
    return *periodic_images_;
}

// PeriodicConditionUsingGhostParticles: three dynamics per periodic axis, used exactly as in the reference case
// files (bounding_ before the cell-linked-list update, ghost_creation_ after it, ghost_update_ wherever a variable
// that neighbours read has changed, e.g. as pre-process of the acoustic half steps, throat.cpp:183-184).
class PeriodicConditionUsingGhostParticles
{
    class Bounding : public BaseDynamics<void>
    {
        PeriodicImages &images_;
        int axis_;

      public:
        Bounding(PeriodicImages &images, int axis) : images_(images), axis_(axis) {}
        void exec(Real = 0.0) override { images_.bounding(axis_); }
    };
    class Creation : public BaseDynamics<void>
    {
        PeriodicImages &images_;
        int axis_;

      public:
        Creation(PeriodicImages &images, int axis) : images_(images), axis_(axis) {}
        void exec(Real = 0.0) override { images_.arm(axis_); }
    };

  public:
    // copies the listed variables (none listed = every stored variable) from the sources to their images
    class Update : public BaseDynamics<void>
    {
        PeriodicImages &images_;
        std::vector<std::string> names_;

      public:
        explicit Update(PeriodicImages &images, std::initializer_list<const char *> names = {}) : images_(images)
        {
            for (const char *n : names) names_.push_back(n);
        }
        void exec(Real = 0.0) override { images_.update(names_); }
    };

  private:
    PeriodicImages &images_;

  public:
    Bounding bounding_;
    Creation ghost_creation_;
    Update ghost_update_;
    PeriodicConditionUsingGhostParticles(RealBody &real_body, Ghost<PeriodicAlongAxis> &ghost_boundary)
        : images_(real_body.definePeriodicImages()), bounding_(images_, ghost_boundary.getAxis()),
          ghost_creation_(images_, ghost_boundary.getAxis()), ghost_update_(images_)
    {
        ghost_boundary.checkParticlesReserved();
        images_.addAxis(ghost_boundary);
    }
    PeriodicImages &images() { return images_; }
};
} // namespace SPH
#endif
