// sphinxsys_ck/configuration.h — cell-linked list, body relations and the configuration dynamics that build them.
//
// Reference (relative to /root/reference/src/shared):
//   CellLinkedList<SPHAdaptation> ...... meshes/cell_linked_list.h:125-167, cell_linked_list.cpp:14,167-175
//   Inner<> / Contact<> / Relation ..... shared_ck/body_relation/relation_ck.h:59-174, relation_ck.hpp:9-85
//   search depth ....................... meshes/cell_linked_list.hpp:161-167
//   UpdateCellLinkedList ............... shared_ck/particle_dynamics/configuration_dynamics/update_cell_linked_list.{h,hpp}
//   UpdateRelation ..................... .../update_body_relation.{h,hpp} (count -> scan -> grow -> fill)
//   ParticleSortCK ..................... .../particle_sort_ck.hpp:61-104, base_configuration_dynamics.h:74-128
#ifndef SPHINXSYS_CK_CONFIGURATION_H
#define SPHINXSYS_CK_CONFIGURATION_H

#include "particles.h"

namespace SPH
{
class CellLinkedList
{
  public:
    sphb200_mesh_t mesh_;
    size_t total_cells_;
    DeviceBuffer cell_offset_, particle_index_;
    CellLinkedList(SPHBody &body)
    {
        SPHSystem &sys = body.getSPHSystem();
        // grid spacing = kernel cut-off radius, buffer width 2 (adaptation.cpp:80-85, cell_linked_list.cpp:14)
        mesh_ = makeMesh(sys.system_domain_bounds_, body.getSPHAdaptation().CutOffRadius(), 2, sys.dim_);
        total_cells_ = (size_t)mesh_.cells[0] * mesh_.cells[1] * mesh_.cells[2];
        size_t n = body.getBaseParticles().ParticlesBound();
        cell_offset_.reset((total_cells_ + 2) * sizeof(uint32_t));
        particle_index_.reset((std::max(n, total_cells_) + 2) * sizeof(uint32_t)); // cell_linked_list.cpp:172-174
    }
    sphb200_cell_list_t view() const
    {
        sphb200_cell_list_t v;
        v.cell_offset = cell_offset_.get<uint32_t>();
        v.particle_index = particle_index_.get<uint32_t>();
        v.sorted_pos = nullptr;
        return v;
    }
    // Another mesh for the same body (spacing >= cut-off radius keeps every neighbour set): ring-decomposed periodic
    // runs need cell planes that tile the periodic box (alignedPeriodicMesh in slab_decomposition.h). Call it before
    // anything sized by the cell count exists (periodic images, relations to this body).
    void resetMesh(const sphb200_mesh_t &mesh, size_t particles_bound)
    {
        mesh_ = mesh;
        total_cells_ = (size_t)mesh_.cells[0] * mesh_.cells[1] * mesh_.cells[2];
        cell_offset_.reset((total_cells_ + 2) * sizeof(uint32_t));
        particle_index_.reset((std::max(particles_bound, total_cells_) + 2) * sizeof(uint32_t));
    }
};
inline CellLinkedList &SPHBody::getCellLinkedList()
{
    if (!cell_linked_list_) cell_linked_list_.reset(new CellLinkedList(*this));
    return *cell_linked_list_;
}
} // namespace SPH
#include "periodic_images.h"
namespace SPH
{
inline SPHBody::~SPHBody() {}

// ---------------------------------------------------------------------------------------------------------
// relations
// ---------------------------------------------------------------------------------------------------------
class RelationBase
{
  public:
    SPHBody &source_;
    SPHBody &target_;
    bool is_inner_;
    DeviceBuffer count_, slice_offset_, index_;
    uint64_t capacity_ = 0, total_ = 0, version_ = 0;
    uint32_t fixed_stride_ = 0, max_count_ = 0; // one-pass build with a fixed row stride; 0 = exact two-phase build
    sphb200_kernel_t kernel_;
    int search_depth_ = 1;
    bool legacy_criterion_ = false; // NeighborBuilder criterion |d|^2 < rc^2 (legacy InnerRelation / ContactRelation)
    bool bank_aligned_ = false;     // rows laid out for the L1 banks by the last build (sphb200_relation_t::bank_aligned)

    RelationBase(SPHBody &source, SPHBody &target, bool is_inner) : source_(source), target_(target), is_inner_(is_inner)
    {
        size_t n = source.getBaseParticles().ParticlesBound();
        count_.reset((n + 2) * sizeof(uint32_t));
        slice_offset_.reset(((n + 31) / 32 + 2) * sizeof(uint32_t));
        capacity_ = n + 1; // relation_ck.hpp:17,28-31: the first exec always grows it
        index_.reset(capacity_ * sizeof(uint32_t));
        SPHAdaptation &sa = source.getSPHAdaptation(), &ta = target.getSPHAdaptation();
        // Neighbor<SPHAdaptation,SPHAdaptation>: inv_h = 1 / max(src_h, tar_h) (neighbor_method.hpp:73-76)
        kernel_ = sa.h_ref_ >= ta.h_ref_ ? sa.kernel_ : ta.kernel_;
        kernel_.src_h = sa.h_ref_;
        if (!is_inner)
        {
            // ContactSearchBox depth: ceil((max(target spacing, source cut-off) - eps) / target spacing)
            Real sp = target.getCellLinkedList().mesh_.spacing;
            Real cut = sa.CutOffRadius();
            Real eps = std::numeric_limits<Real>::epsilon();
            search_depth_ = (int)std::ceil((std::max(sp, cut) - eps) / sp);
        }
        fixed_stride_ = source.getSPHSystem().dim_ == 3 ? 128 : 40;
    }
    virtual ~RelationBase() {}
    SPHBody &getSPHBody() { return source_; }
    sphb200_relation_t view() const
    {
        sphb200_relation_t r;
        r.count = count_.get<uint32_t>();
        r.slice_offset = slice_offset_.get<uint32_t>();
        r.index = index_.get<uint32_t>();
        r.capacity = capacity_;
        r.order = nullptr; // storage is cell ordered: slot == particle id
        // contact relations: the target body is not decomposed with the source, see sphb200_relation_t::bank_aligned
        // (and when the target keeps only a slab of its particles as well — WallSlab — its own slot origin comes off again:
        // the layout class of an entry is s_global - t_global whatever the two bodies store)
        r.bank_aligned = bank_aligned_ ? 1 + (is_inner_ ? 0 : (int)((source_.slotOrigin() - target_.slotOrigin()) & 7u)) : 0;
        return r;
    }
    sphb200_search_t search()
    {
        sphb200_search_t s;
        std::memset(&s, 0, sizeof(s));
        if (source_.periodicImages()) source_.periodicImages()->ensure();
        if (target_.periodicImages()) target_.periodicImages()->ensure();
        CellLinkedList &tcl = target_.getCellLinkedList();
        s.tar_mesh = tcl.mesh_;
        s.kernel = kernel_;
        s.src_pos = (const sphb200_vec4_t *)source_.getBaseParticles().deviceData<Vecd>("Position");
        s.n_src = (uint32_t)source_.TotalRealParticles();
        s.tar_pos = (const sphb200_vec4_t *)target_.getBaseParticles().deviceData<Vecd>("Position");
        s.tar_list = tcl.view();
        s.is_inner = is_inner_ ? 1 : 0;
        s.legacy_criterion = legacy_criterion_ ? 1 : 0;
        s.search_depth = search_depth_;
        s.src_begin = (uint32_t)source_.getBaseParticles().activeBegin();
        s.src_end = (uint32_t)source_.getBaseParticles().activeEnd();
        s.cell_ordered = (source_.isCellOrdered() && target_.isCellOrdered()) ? 1 : 0;
        if (PeriodicImages *im = target_.periodicImages())
            if (im->ghostParticles())
            {
                // the images of the target body: second candidate set, stored behind its real particles
                if (!s.cell_ordered) throw SphError("periodic images need cell-ordered bodies");
                s.tar2_pos = s.tar_pos + im->realParticles();
                s.tar2_list = im->listView();
                s.tar2_index_base = im->realParticles();
            }
        return s;
    }
    void grow(uint64_t entries)
    {
        // DiscreteVariable::reallocateData: 1.25 x the required size, no copy (sphinxsys_variable.h:368-375)
        capacity_ = (uint64_t)(entries * 1.25) + 1;
        index_.reset(capacity_ * sizeof(uint32_t));
        ++version_;
    }
    // the reference's particle_offset_/neighbor_index_ CSR, in REFERENCE particle ids (host arrays)
    void exportCSR(std::vector<uint32_t> &offset, std::vector<uint32_t> &index)
    {
        uint32_t n = (uint32_t)source_.getBaseParticles().hostSyncCount(); // rows of image particles are empty and not exported
        ExecutionInstance &ex = execution_instance();
        DeviceBuffer d_off((n + 2) * sizeof(uint32_t)), d_idx((std::max<uint64_t>(total_, 1)) * sizeof(uint32_t));
        SPHCK_CALL(sphb200_relation_export_csr, view(), n, source_.getBaseParticles().referenceID(),
                   target_.getBaseParticles().referenceID(), d_off.get<uint32_t>(), d_idx.get<uint32_t>(),
                   std::max<uint64_t>(total_, 1), ex.stream());
        offset.assign(n + 1, 0);
        ex.check(sphb200_copy_d2h(offset.data(), d_off.get(), (n + 1) * sizeof(uint32_t), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
        index.assign(offset[n], 0);
        if (offset[n]) ex.check(sphb200_copy_d2h(index.data(), d_idx.get(), (size_t)offset[n] * sizeof(uint32_t), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
    }
};

template <typename... Parameters> class Inner;
template <typename... Parameters> class Contact;
template <> class Inner<> : public RelationBase
{
  public:
    explicit Inner(SPHBody &body) : RelationBase(body, body, true) {}
};
template <> class Contact<> : public RelationBase
{
  public:
    Contact(SPHBody &body, std::initializer_list<SPHBody *> contact_bodies) : RelationBase(body, checked(contact_bodies), false) {}

  private:
    static SPHBody &checked(std::initializer_list<SPHBody *> bodies)
    {
        if (bodies.size() != 1) throw SphError("Contact<>: the hot path covers exactly one contact body per relation");
        return **bodies.begin();
    }
};

// ---------------------------------------------------------------------------------------------------------
// UpdateCellLinkedList<Policy, RealBody>
// ---------------------------------------------------------------------------------------------------------
template <class ExecutionPolicy, class BodyType = RealBody> class UpdateCellLinkedList : public BaseDynamics<void>
{
    SPHBody &body_;

  public:
    explicit UpdateCellLinkedList(SPHBody &body) : body_(body) { execution::require_device_policy<ExecutionPolicy>(); }
    // count -> scan -> fill (update_cell_linked_list.hpp:75-106), then the storage of the body is brought into the
    // new cell order by one fused gather of every registered variable (derived records are rebuilt instead)
    void exec(Real dt = 0.0) override
    {
        BaseParticles &p = body_.getBaseParticles();
        CellLinkedList &cl = body_.getCellLinkedList();
        if (PeriodicImages *im = body_.periodicImages())
        {
            // the images of the previous configuration are dropped; ghost_creation_ makes the new ones
            p.setTotalRealParticles(im->realParticles());
            p.setActiveRange(0, im->realParticles());
            im->invalidate();
        }
        uint32_t n = (uint32_t)p.TotalRealParticles();
        std::vector<DiscreteVariableBase *> vars = p.reorderedVariables();
        std::vector<void *> dst(vars.size());
        std::vector<const void *> src(vars.size());
        std::vector<uint32_t> bytes(vars.size());
        for (size_t k = 0; k < vars.size(); ++k)
        {
            dst[k] = vars[k]->shadowAddress();
            src[k] = vars[k]->deviceAddress();
            bytes[k] = vars[k]->deviceElementBytes();
        }
        SPHCK_CALL(sphb200_cell_list_build_reorder, &cl.mesh_, (const sphb200_vec4_t *)p.deviceData<Vecd>("Position"), n,
                   p.referenceID(), cl.view(), (int)vars.size(), dst.data(), src.data(), bytes.data(),
                   execution_instance().stream());
        for (auto *v : vars) v->swapWithShadow();
        p.storageReordered();
        body_.setCellOrdered(true);
        body_.setPosVolDirty();
    }
};

// ---------------------------------------------------------------------------------------------------------
// UpdateRelation<Policy, Inner<>, Contact<>> (any number of relations, executed in order)
// ---------------------------------------------------------------------------------------------------------
// SPHB200_BANK_ALIGN=0 in the environment switches the L1-bank-aligned row layout off (A/B measurements only)
inline bool bankAlignEnabled()
{
    static const bool on = [] {
        const char *e = std::getenv("SPHB200_BANK_ALIGN");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <class ExecutionPolicy, class... RelationTypes> class UpdateRelation : public BaseDynamics<void>
{
    std::vector<RelationBase *> relations_;

  public:
    explicit UpdateRelation(RelationTypes &...relations) : relations_{&relations...} { execution::require_device_policy<ExecutionPolicy>(); }
    void exec(Real dt = 0.0) override
    {
        ExecutionInstance &ex = execution_instance();
        for (RelationBase *r : relations_)
        {
            sphb200_search_t s;
            SPHCK_STAGE("  relation: search arguments (+ pending periodic images)", s = r->search()); // creates pending periodic images first: they are stored particles too
            uint32_t n = s.n_src;
            r->bank_aligned_ = r->fixed_stride_ && s.cell_ordered && bankAlignEnabled();
            if (r->fixed_stride_)
            {
                uint64_t need = (uint64_t)((n + 31) / 32) * 32ull * r->fixed_stride_;
                if (need > r->capacity_) SPHCK_STAGE("  relation: grow the index array", r->grow(need));
                uint32_t mx = 0;
                SPHCK_STAGE("  relation: one-pass build", SPHCK_CALL(sphb200_relation_build_fixed, &s, r->view(), r->fixed_stride_, &mx, ex.stream()));
                r->max_count_ = mx;
                r->total_ = need;
                if (mx <= r->fixed_stride_) continue;
                r->fixed_stride_ = 0; // a row overflowed the stride: rebuild exactly, and stay exact from now on
                r->bank_aligned_ = false;
            }
            uint64_t required = 0;
            SPHCK_CALL(sphb200_relation_count, &s, r->view(), &required, ex.stream());
            r->total_ = required;
            if (required > r->capacity_) r->grow(required); // update_body_relation.hpp:145-155
            SPHCK_CALL(sphb200_relation_fill, &s, r->view(), ex.stream());
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// ParticleSortCK<Policy>: new reference ids = stable sort of the Morton keys of the cell indices. Storage is cell
// ordered already, so what changes here is the particle NUMBERING the host sees (ReferenceID / SortedID), plus —
// exactly as in the reference — every variable that is NOT in the evolving list keeps its old index: particle
// with new id r inherits the non-evolving values stored at old id r ("Force" is the one that carries information
// across this point: acoustic_step_2nd_half.hpp:72 -> acoustic_step_1st_half.hpp:109).
// ---------------------------------------------------------------------------------------------------------
template <class ExecutionPolicy> class ParticleSortCK : public BaseDynamics<void>
{
    SPHBody &body_;
    DeviceBuffer keys_slot_, keys_ref_, perm_, new_of_old_, new_rid_, carry_perm_;

  public:
    explicit ParticleSortCK(SPHBody &body) : body_(body)
    {
        execution::require_device_policy<ExecutionPolicy>();
        body.getBaseParticles().registerStateVariable<UnsignedInt>("SortedID"); // indexed by OriginalID, not by slot
        body.getBaseParticles().markDerived("SortedID");
    }
    void exec(Real dt = 0.0) override
    {
        ExecutionInstance &ex = execution_instance();
        BaseParticles &p = body_.getBaseParticles();
        uint32_t n = (uint32_t)p.hostSyncCount(); // real particles (images of a periodic body are not renumbered)
        if (n == 0) return;
        size_t nb = ((size_t)n + 1) * sizeof(uint32_t);
        for (DeviceBuffer *b : {&keys_slot_, &keys_ref_, &perm_, &new_of_old_, &new_rid_, &carry_perm_}) b->ensure(nb);
        void *st = ex.stream();
        CellLinkedList &cl = body_.getCellLinkedList();
        uint32_t *rid = p.referenceID();
        uint32_t *inv = p.inverseReferenceID(); // slot of old reference id r
        // prepareSequence (particle_sort_ck.hpp:61-67): keys in reference order, sequence = iota
        SPHCK_CALL(sphb200_morton_keys, &cl.mesh_, (const sphb200_vec4_t *)p.deviceData<Vecd>("Position"), n,
                   keys_slot_.get<uint32_t>(), perm_.get<uint32_t>(), nullptr, st);
        {
            void *dst[1] = {keys_ref_.get()};
            const void *src[1] = {keys_slot_.get()};
            uint32_t eb[1] = {4};
            SPHCK_CALL(sphb200_gather_multi, 1, dst, src, eb, inv, n, st);
        }
        // sort_by_key (stable; the reference's unstable sort leaves ties arbitrary): perm[new id] = old id
        SPHCK_CALL(sphb200_sort_pairs_u32, keys_ref_.get<uint32_t>(), perm_.get<uint32_t>(), (uint64_t)n, 30, st);
        // new_of_old[perm[r]] = r ; new_rid[slot] = new_of_old[rid[slot]]
        SPHCK_CALL(sphb200_update_sorted_id, perm_.get<uint32_t>(), new_of_old_.get<uint32_t>(), n, st);
        {
            void *dst[1] = {new_rid_.get()};
            const void *src[1] = {new_of_old_.get()};
            uint32_t eb[1] = {4};
            SPHCK_CALL(sphb200_gather_multi, 1, dst, src, eb, rid, n, st);
        }
        // non-evolving carry: value_new[slot] = value_old[inv[new_rid[slot]]] for every variable the reference
        // does NOT permute (everything outside the evolving list keeps its array index there)
        {
            std::vector<DiscreteVariableBase *> carried;
            for (DiscreteVariableBase *v : p.reorderedVariables())
                if (!p.isEvolving(v) && v != p.referenceIDVariable()) carried.push_back(v);
            if (!carried.empty())
            {
                {
                    void *dst[1] = {carry_perm_.get()};
                    const void *src[1] = {inv};
                    uint32_t eb[1] = {4};
                    SPHCK_CALL(sphb200_gather_multi, 1, dst, src, eb, new_rid_.get<uint32_t>(), n, st);
                }
                std::vector<void *> dst(carried.size());
                std::vector<const void *> src(carried.size());
                std::vector<uint32_t> eb(carried.size());
                for (size_t k = 0; k < carried.size(); ++k)
                {
                    dst[k] = carried[k]->shadowAddress();
                    src[k] = carried[k]->deviceAddress();
                    eb[k] = carried[k]->deviceElementBytes();
                }
                SPHCK_CALL(sphb200_gather_multi, (int)carried.size(), dst.data(), src.data(), eb.data(), carry_perm_.get<uint32_t>(), n, st);
                for (auto *v : carried) v->swapWithShadow();
            }
        }
        ex.check(sphb200_copy_d2d(rid, new_rid_.get(), (size_t)n * sizeof(uint32_t), st), "sphb200_copy_d2d");
        p.storageReordered(); // invalidates the cached inverse map
        // updateSortedID (particle_sort_ck.hpp:69-74): sorted_id[original_id[i]] = i, i = reference id
        {
            uint32_t *inv_new = p.inverseReferenceID();
            void *dst[1] = {keys_slot_.get()}; // OriginalID in reference order
            const void *src[1] = {p.deviceData<UnsignedInt>("OriginalID")};
            uint32_t eb[1] = {4};
            SPHCK_CALL(sphb200_gather_multi, 1, dst, src, eb, inv_new, n, st);
            SPHCK_CALL(sphb200_update_sorted_id, keys_slot_.get<uint32_t>(), (uint32_t *)p.deviceData<UnsignedInt>("SortedID"), n, st);
        }
    }
};
} // namespace SPH
#endif
