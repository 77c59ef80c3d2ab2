// sphinxsys_ck/slab_decomposition.h — 1-D slab domain decomposition of a fluid body over the GPUs of one node.
// NEW functionality (the reference has no distributed path). One process per GPU.
//
// * Every rank keeps the GLOBAL mesh, so cell ids are the single-GPU ones; rank r owns the cell planes
//   [cut[r], cut[r+1]) along x (the slowest cell axis) and keeps a full copy of the static wall body.
// * Storage stays cell ordered: `left ghost plane | own planes | right ghost plane`. A cell plane is therefore ONE
//   contiguous slot range of every variable array, in the same particle order on the sender and on the receiver
//   (in-cell order = ascending global ReferenceID), so halo refresh is a grouped ncclSend/ncclRecv out of and into the
//   variable arrays themselves — no pack/unpack kernels, no index maps.
// * rebuild() replaces UpdateCellLinkedList::exec() once per advection step: sort the own particles by their new cell,
//   pick the particles next to each cut (those that left the slab + the boundary plane) with one pass over the own
//   positions, hand them to the neighbour in one message per variable, append what arrives, sort everything into cell order. Particles move less than one cell per advection step (CFL),
//   so leavers are always in the plane next to the cut. Dynamics then run on the active slot range only.
// * Cuts are particle-count quantiles of the distribution along x (planSlabCuts), re-balanced at the sort cadence
//   (recut()). Periodic boxes close the chain of slabs into a ring (SeamRing below).
#ifndef SPHINXSYS_CK_SLAB_DECOMPOSITION_H
#define SPHINXSYS_CK_SLAB_DECOMPOSITION_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <utility>

#include "configuration.h"

namespace SPH
{
// Mesh::CellIndexFromPosition along one axis on the host, same arithmetic as the device (base_mesh.hxx:9-15)
inline int hostCellCoordinate(Real x, Real lower, Real spacing, int cells)
{
    Real t = x - lower;
    Real u = t / spacing;
    int k = (int)std::floor(u);
    return std::min(std::max(k, 0), cells - 1);
}

// Cell-plane cuts cut[0..nranks] with cut[0] = 0, cut[nranks] = planes: every rank gets whole planes and, as far as
// whole planes allow, the same number of particles. Ranks never get an empty plane range.
inline std::vector<int> planSlabCuts(const std::vector<uint64_t> &particles_per_plane, int nranks)
{
    const int planes = (int)particles_per_plane.size();
    if (nranks < 1 || planes < nranks) throw SphError("planSlabCuts: fewer cell planes than ranks");
    uint64_t total = 0;
    for (uint64_t c : particles_per_plane) total += c;
    std::vector<int> cut(nranks + 1, 0);
    cut[nranks] = planes;
    uint64_t cum = 0;
    int p = 0;
    for (int r = 1; r < nranks; ++r)
    {
        // smallest plane index whose cumulative count reaches r/nranks of the total
        const double target = double(total) * double(r) / double(nranks);
        while (p < planes && double(cum + particles_per_plane[p]) <= target) cum += particles_per_plane[p++];
        // choose the nearer side of the plane that straddles the target
        int c = p;
        if (p < planes && (target - double(cum)) > (double(cum + particles_per_plane[p]) - target)) c = p + 1;
        c = std::max(c, cut[r - 1] + 1);               // at least one plane per rank
        c = std::min(c, planes - (nranks - r));        // leave one plane for each remaining rank
        cut[r] = c;
        while (p < c) cum += particles_per_plane[p++];
    }
    return cut;
}

// Cut r may move at most to the far end of an adjacent slab (old cut r-1 + 1 ... old cut r+1 - 1) and cuts stay strictly
// increasing: all particles that change owner then move between NEIGHBOUR ranks only.
inline std::vector<int> limitCutMoves(const std::vector<int> &old_cuts, std::vector<int> wanted)
{
    const int n = (int)old_cuts.size() - 1;
    if ((int)wanted.size() != n + 1) throw SphError("limitCutMoves: cut vectors differ in length");
    wanted[0] = old_cuts[0];
    wanted[n] = old_cuts[n];
    for (int r = 1; r < n; ++r)
    {
        int c = std::min(std::max(wanted[r], old_cuts[r - 1] + 1), old_cuts[r + 1] - 1);
        c = std::max(c, wanted[r - 1] + 1);
        wanted[r] = c;
    }
    for (int r = n - 1; r >= 1; --r) wanted[r] = std::min(wanted[r], wanted[r + 1] - 1);
    return wanted;
}

// Periodic runs along x on a RING of slabs (BASELINE config 4 on N GPUs): the first and the last rank are neighbours and
// what crosses the seam travels with x shifted by -/+ L (sphb200_seam_shift). Ownership is by cell plane, so the planes
// tile the box: [lower, upper) = the planes [first_plane, first_plane + box_planes) up to rounding, and the shift keeps
// every particle in the plane it belongs to (sphb200_seam_t). The protocol is pinned on the CPU by oracle/decomposed.py
// (tests/test_decomposed_oracle_cpu.py).
struct SeamRing
{
    bool on = false;
    Real lower = 0, upper = 0; // the periodic box along x
    sphb200_seam_t seam;
    int first_plane() const { return seam.first_plane; }
    int box_planes() const { return seam.box_planes; }
};

// {largest x whose cell plane is < plane, smallest x whose cell plane is >= plane} by bisection over the mesh's own
// cell arithmetic (monotone in x)
inline std::pair<Real, Real> planeFace(const sphb200_mesh_t &m, int plane)
{
    const Real face = m.lower[0] + Real(plane) * m.spacing;
    Real lo = face - Real(0.5) * m.spacing, hi = face + Real(0.5) * m.spacing;
    if (!(hostCellCoordinate(lo, m.lower[0], m.spacing, m.cells[0]) < plane && hostCellCoordinate(hi, m.lower[0], m.spacing, m.cells[0]) >= plane))
        throw SphError("planeFace: the cell arithmetic does not bracket the plane face");
    for (int it = 0; it < 200; ++it)
    {
        const Real mid = lo + (hi - lo) * Real(0.5);
        if (mid <= lo || mid >= hi) break;
        if (hostCellCoordinate(mid, m.lower[0], m.spacing, m.cells[0]) < plane) lo = mid;
        else hi = mid;
    }
    return {lo, hi};
}

// Mesh for a ring-decomposed periodic body: spacing = L_x / floor(L_x / cutoff) >= cutoff (neighbour sets do not depend
// on the mesh once the spacing is at least the cut-off radius), `ghost_planes` cell layers around the box on every axis
// (room for the images of the other axes and for the seam ghost planes).
inline sphb200_mesh_t alignedPeriodicMesh(const BoundingBoxd &box, Real cutoff, int dim, SeamRing &ring, int ghost_planes = 2)
{
    const Real lo = box.lower_[0], up = box.upper_[0], L = up - lo;
    const int planes = (int)std::floor(double(L) / double(cutoff));
    if (planes < 3) throw SphError("alignedPeriodicMesh: the periodic box is narrower than three cut-off radii");
    Real spacing = L / Real(planes);
    while (spacing < cutoff) spacing = std::nextafter(spacing, std::numeric_limits<Real>::max());
    sphb200_mesh_t m;
    m.spacing = spacing;
    for (int d = 0; d < 3; ++d)
    {
        if (d >= dim)
        {
            m.lower[d] = 0;
            m.cells[d] = 1;
            continue;
        }
        m.lower[d] = box.lower_[d] - Real(ghost_planes) * spacing;
        m.cells[d] = d == 0 ? planes + 2 * ghost_planes
                            : (int)std::ceil(double(box.upper_[d] - m.lower[d]) / double(spacing)) + ghost_planes;
    }
    const int k0 = ghost_planes;
    ring.on = true;
    ring.lower = lo;
    ring.upper = up;
    sphb200_seam_t &sm = ring.seam;
    sm.mesh_lower = m.lower[0];
    sm.mesh_spacing = spacing;
    sm.mesh_cells = m.cells[0];
    sm.first_plane = k0;
    sm.box_planes = planes;
    std::pair<Real, Real> f;
    f = planeFace(m, k0 - 1);          sm.ghost_low_min = f.second;
    f = planeFace(m, k0);              sm.ghost_low_max = f.first;  sm.own_min = f.second;
    f = planeFace(m, k0 + planes);     sm.own_max = f.first;        sm.ghost_high_min = f.second;
    f = planeFace(m, k0 + planes + 1); sm.ghost_high_max = f.first;
    // the faces of the box sit on the plane faces up to rounding
    // (lower + k * spacing is rounded per plane: the distance grows with the number of planes, a few ulp of the box length)
    const Real tol = Real(1e-5) * spacing + Real(16) * std::numeric_limits<Real>::epsilon() * std::max({std::abs(lo), std::abs(up), L});
    if (std::abs(sm.own_min - lo) > tol || std::abs(sm.ghost_high_min - up) > tol)
        throw SphError("alignedPeriodicMesh: the box faces are not on cell-plane faces");
    return m;
}

class SlabDecomposition
{
    SPHBody &body_;
    int rank_, nranks_;
    SeamRing ring_;
    std::vector<int> cuts_;
    uint32_t plane_cells_; // cells per x plane (ny * nz)
    // slot offsets after the last rebuild: ghosts [0, a0) | first own plane [a0, f1) ... last own plane [l0, a1) | ghosts [a1, n)
    uint32_t a0_ = 0, f1_ = 0, l0_ = 0, a1_ = 0, n_ = 0;
    DeviceBuffer scalars_; // small device scratch for counts and reductions
    DeviceBuffer select_;  // index lists of the outgoing particles (rebuild)
    DeviceBuffer plane_scratch_; // plane-boundary offsets and the particles-per-plane histogram (recut)
    uint64_t recuts_ = 0;
    uint64_t migrated_out_ = 0, ghost_particles_ = 0;
    // peer-mailbox rebuild (rebuildPeer): sequence number of the pushes, capacity the mailboxes were opened with
    uint64_t mail_seq_ = 0;
    bool mail_open_ = false;
    bool configured_ = false; // a rebuild has run: planes and ghosts exist (every rank flips this at the same call)
    uint64_t host_syncs_ = 0; // host round trips spent in rebuilds (diagnostics)
    // overlap of the plane exchange with interior compute: a high-priority side stream carries the exchanges and the
    // boundary-plane launches, the default stream the interior launches; events order the two (dambreak_case.h)
    void *side_stream_ = nullptr;
    void *events_[4] = {nullptr, nullptr, nullptr, nullptr};

    uint32_t readOffset(uint32_t cell)
    {
        uint32_t v = 0;
        ExecutionInstance &ex = execution_instance();
        ex.check(sphb200_copy_d2h(&v, body_.getCellLinkedList().cell_offset_.get<uint32_t>() + cell, sizeof(uint32_t), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
        return v;
    }
    void readOffsets(const uint32_t *cells, uint32_t *out, int count)
    {
        ExecutionInstance &ex = execution_instance();
        for (int k = 0; k < count; ++k)
            ex.check(sphb200_copy_d2h(out + k, body_.getCellLinkedList().cell_offset_.get<uint32_t>() + cells[k], sizeof(uint32_t), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
    }
    // cell-list build + storage reorder of the slots [begin, begin + n) into [0, n); n_dev != nullptr: n is the capacity
    // the launches are sized for and the live count is read from the device
    void reorder(uint32_t begin, uint32_t n, const uint32_t *n_dev = nullptr)
    {
        BaseParticles &p = body_.getBaseParticles();
        CellLinkedList &cl = body_.getCellLinkedList();
        std::vector<DiscreteVariableBase *> vars = p.reorderedVariables();
        std::vector<void *> dst(vars.size());
        std::vector<const void *> src(vars.size());
        std::vector<uint32_t> bytes(vars.size());
        for (size_t k = 0; k < vars.size(); ++k)
        {
            bytes[k] = vars[k]->deviceElementBytes();
            dst[k] = vars[k]->shadowAddress();
            src[k] = (const char *)vars[k]->deviceAddress() + (size_t)begin * bytes[k];
        }
        SPHCK_CALL(sphb200_cell_list_build_reorder_n, &cl.mesh_, (const sphb200_vec4_t *)p.deviceData<Vecd>("Position") + begin, n, n_dev,
                   p.referenceID() + begin, cl.view(), (int)vars.size(), dst.data(), src.data(), bytes.data(),
                   execution_instance().stream());
        for (auto *v : vars) v->swapWithShadow();
        p.storageReordered();
    }

  public:
    SlabDecomposition(SPHBody &body, int rank, int nranks, const std::vector<int> &cuts, const SeamRing &ring = SeamRing())
        : body_(body), rank_(rank), nranks_(nranks), ring_(ring), cuts_(cuts), scalars_(1024 + 8 * ((size_t)nranks + 2))
    {
        const sphb200_mesh_t &m = body.getCellLinkedList().mesh_;
        plane_cells_ = (uint32_t)m.cells[1] * (uint32_t)m.cells[2];
        if (ring_.on)
        {
            if ((int)cuts.size() != nranks + 1 || cuts.front() != ring_.first_plane() || cuts.back() != ring_.first_plane() + ring_.box_planes())
                throw SphError("SlabDecomposition: the cuts of a ring must run over the cell planes of the periodic box");
            for (int r = 0; r < nranks; ++r)
                if (cuts[r + 1] <= cuts[r]) throw SphError("SlabDecomposition: every rank of a ring needs at least one plane");
            if (nranks == 1 && ring_.box_planes() < 3) throw SphError("SlabDecomposition: a ring of one slab needs three planes");
            if (!sphb200_comm_is_ring(execution_instance().ctx()))
                throw SphError("SlabDecomposition: the communicator is not a ring (sphb200_comm_set_ring)");
        }
        else if ((int)cuts.size() != nranks + 1 || cuts.front() != 0 || cuts.back() != m.cells[0])
            throw SphError("SlabDecomposition: cuts must run from 0 to the number of x planes");
        BaseParticles &p = body.getBaseParticles();
        n_ = (uint32_t)p.TotalRealParticles();
        a0_ = 0;
        a1_ = n_;
        p.setActiveRange(a0_, a1_);
    }
    ~SlabDecomposition()
    {
        if (mail_open_) sphb200_comm_mailbox_close(execution_instance().ctx());
        for (void *e : events_)
            if (e) sphb200_event_destroy(e);
        if (side_stream_) sphb200_stream_destroy(side_stream_);
    }
    SlabDecomposition(const SlabDecomposition &) = delete;
    SlabDecomposition &operator=(const SlabDecomposition &) = delete;
    int rank() const { return rank_; }
    int size() const { return nranks_; }
    bool isRing() const { return ring_.on; }
    bool hasLeft() const { return ring_.on || rank_ > 0; }
    bool hasRight() const { return ring_.on || rank_ + 1 < nranks_; }
    // --- boundary / interior split of the own slots (valid after rebuild()) ---
    // Boundary planes are the own planes a neighbour rank reads (and whose particles read ghost planes): the first own
    // plane if there is a left neighbour, the last own plane if there is a right neighbour. Interior slots read own
    // particles only, so they can be advanced while the ghost planes are in flight.
    struct SlotRange
    {
        uint32_t begin, end;
        bool empty() const { return end <= begin; }
    };
    SlotRange leftBoundary() const
    {
        if (!hasLeft()) return {a0_, a0_};
        return {a0_, std::min(f1_, a1_)};
    }
    SlotRange rightBoundary() const
    {
        if (!hasRight()) return {a1_, a1_};
        return {std::max(l0_, leftBoundary().end), a1_}; // one-plane slabs: the plane is the left boundary already
    }
    SlotRange interior() const { return {leftBoundary().end, rightBoundary().begin}; }
    void *sideStream()
    {
        if (!side_stream_) execution_instance().check(sphb200_stream_create_with_priority(&side_stream_, 1), "sphb200_stream_create_with_priority");
        return side_stream_;
    }
    void *event(int k)
    {
        if (!events_[k]) execution_instance().check(sphb200_event_create(&events_[k]), "sphb200_event_create");
        return events_[k];
    }
    // signal(k, s): event k marks everything issued so far on stream s; await(k, s): stream s does not run past this
    // point before that mark is reached (signal must be issued first in host order)
    void signal(int k, void *stream) { execution_instance().check(sphb200_event_record(event(k), stream), "sphb200_event_record"); }
    void await(int k, void *stream) { execution_instance().check(sphb200_stream_wait_event(stream, event(k)), "sphb200_stream_wait_event"); }

    const std::vector<int> &cuts() const { return cuts_; }
    uint32_t ownParticles() const { return a1_ - a0_; }
    uint32_t ownBegin() const { return a0_; }
    uint32_t storedParticles() const { return n_; }
    uint64_t ghostParticles() const { return ghost_particles_; }
    uint64_t migratedOut() const { return migrated_out_; }

    // SPHB200_PEER_REBUILD=0 keeps the NCCL exchange (count round trip + grouped send/recv) for every rebuild
    static bool peerRebuildEnabled()
    {
        static const bool on = [] {
            const char *e = std::getenv("SPHB200_PEER_REBUILD");
            return !(e && e[0] == '0');
        }();
        return on;
    }
    uint64_t hostSyncs() const { return host_syncs_; }

    // The configuration update of an advection step. Chains of slabs use the peer-mailbox rebuild (one host round trip,
    // migrants written straight into the neighbours' memory); rings, the very first rebuild (which sizes the mailboxes from
    // the planes it finds) and the hand-over rebuilds of recut() — whole slabs may change owner there — use the NCCL one.
    void update()
    {
        // the choice depends on nothing rank-local: all ranks take the same branch at the same call
        if (!ring_.on && nranks_ > 1 && peerRebuildEnabled() && configured_) rebuildPeer();
        else rebuild();
    }

    // rebuild() without host-known message sizes. Everything up to the last step is enqueued without a host round trip:
    //   select -> push to both neighbours (gather + remote stores into their mailboxes + release, ONE kernel per side)
    //   -> pull from both neighbours (device-side wait on the mailbox header, append behind the own slots, count stays on
    //   the device) -> cell-list build + reorder sized for the storage, live count read from the device -> plane offsets,
    //   counts and status gathered into one record -> all-gather of the own counts (slot origins) -> ONE copy to the host.
    void rebuildPeer()
    {
        ExecutionInstance &ex = execution_instance();
        BaseParticles &p = body_.getBaseParticles();
        CellLinkedList &cl = body_.getCellLinkedList();
        void *st = ex.stream();
        const uint32_t n_old = a1_ - a0_;
        const int X0 = cuts_[rank_], X1 = cuts_[rank_ + 1];
        std::vector<DiscreteVariableBase *> vars = p.reorderedVariables();
        const size_t k = vars.size();
        std::vector<const void *> src(k);
        std::vector<void *> dst(k);
        std::vector<uint32_t> bytes(k);
        size_t entry_bytes = 0;
        for (size_t i = 0; i < k; ++i)
        {
            bytes[i] = vars[i]->deviceElementBytes();
            src[i] = vars[i]->deviceAddress();
            dst[i] = vars[i]->deviceAddress();
            entry_bytes += bytes[i];
        }
        if (!mail_open_)
        {
            // a box holds four times the larger boundary plane (leavers are a small fraction of a plane per step); what does
            // not fit is reported by the status word, not written
            // (the largest plane of ALL ranks: a box must hold what the NEIGHBOUR sends)
            const size_t mine = std::max<size_t>(std::max(f1_ - a0_, a1_ - l0_), 4096);
            const size_t plane = (size_t)allReduceMax(Real(mine / 1024 + 1)) * 1024;
            ex.check(sphb200_comm_mailbox_open(ex.ctx(), 64 + 4 * plane * entry_bytes), "sphb200_comm_mailbox_open");
            mail_open_ = true;
        }
        const uint64_t seq = ++mail_seq_;
        select_.ensure((size_t)2 * n_old * sizeof(uint32_t) + 64);
        uint32_t *left_idx = select_.get<uint32_t>(), *right_idx = left_idx + n_old;
        // device words of this rebuild (bytes 768.. of the scratch): [0,1] send counts, [2,3] receive counts, [4] stored total,
        // [5..8] cells of the plane boundaries, [16..23] the record that goes to the host
        uint32_t *w = scalars_.get<uint32_t>() + 192;
        uint32_t *d_send = w, *d_recv = w + 2, *d_ntot = w + 4, *d_cells = w + 5, *d_rec = w + 16;
        SPHCK_STAGE("  rebuild: select", SPHCK_CALL(sphb200_slab_select, &cl.mesh_, (const sphb200_vec4_t *)p.deviceData<Vecd>("Position"), a0_, n_old,
                                                    hasLeft() ? X0 : -1, hasRight() ? X1 - 1 : -1, left_idx, right_idx, d_send, st));
        if (hasLeft()) SPHCK_STAGE("  rebuild: push left", SPHCK_CALL(sphb200_comm_push, 0, (int)k, src.data(), bytes.data(), left_idx, d_send, seq, st));
        if (hasRight()) SPHCK_STAGE("  rebuild: push right", SPHCK_CALL(sphb200_comm_push, 1, (int)k, src.data(), bytes.data(), right_idx, d_send + 1, seq, st));
        const uint32_t bound = (uint32_t)p.ParticlesBound();
        if (hasLeft()) SPHCK_STAGE("  rebuild: pull left (wait + unpack)", SPHCK_CALL(sphb200_comm_pull, 0, (int)k, dst.data(), bytes.data(), a1_, nullptr, bound, d_recv, seq, st));
        if (hasRight())
            SPHCK_STAGE("  rebuild: pull right (wait + unpack)",
                        SPHCK_CALL(sphb200_comm_pull, 1, (int)k, dst.data(), bytes.data(), a1_, hasLeft() ? d_recv : nullptr, bound, d_recv + 1, seq, st));
        SPHCK_CALL(sphb200_slab_total, n_old, hasLeft() ? d_recv : nullptr, hasRight() ? d_recv + 1 : nullptr, d_ntot, st);
        // everything into cell order at the front of the arrays; launches sized for what the storage can hold
        // ... which is an eighth above the last stored total (own + ghost planes change by well under a percent per step;
        // launching for the whole reserve — twice the own count — cost 0.1 ms of empty blocks), never beyond the storage
        const uint32_t capacity = (uint32_t)std::min<uint64_t>(bound - a0_, (uint64_t)n_ + n_ / 8 + 8192);
        p.setTotalRealParticles(bound);
        SPHCK_STAGE("  rebuild: cell list + reorder", reorder(a0_, capacity, d_ntot));
        uint32_t cells4[4] = {(uint32_t)X0 * plane_cells_, (uint32_t)(X0 + 1) * plane_cells_, (uint32_t)(X1 - 1) * plane_cells_,
                              (uint32_t)X1 * plane_cells_};
        ex.check(sphb200_copy_h2d(d_cells, cells4, sizeof(cells4), st), "sphb200_copy_h2d");
        uint64_t *d_own = scalars_.get<uint64_t>() + 128, *d_all = d_own + 1; // bytes 1024.. : behind the reduction windows
        SPHCK_CALL(sphb200_slab_bounds, cl.cell_offset_.get<uint32_t>(), d_cells, 4, d_ntot, 0 | (3 << 8), d_rec, d_own, st);
        SPHCK_STAGE("  rebuild: all-gather of the own counts", SPHCK_CALL(sphb200_comm_allgather_u64, d_own, d_all, 1, st));
        // the ONE host round trip: plane offsets, stored total, status, send counts, own counts of all ranks
        uint32_t rec[6] = {0, 0, 0, 0, 0, 0}, sent[2] = {0, 0};
        std::vector<uint64_t> all(nranks_);
        ex.check(sphb200_copy_d2h(rec, d_rec, sizeof(rec), st), "sphb200_copy_d2h");
        ex.check(sphb200_copy_d2h(sent, d_send, sizeof(sent), st), "sphb200_copy_d2h");
        ex.check(sphb200_copy_d2h(all.data(), d_all, all.size() * sizeof(uint64_t), st), "sphb200_copy_d2h");
        ex.synchronize();
        ++host_syncs_;
        if (rec[4] > capacity)
            throw SphError("SlabDecomposition::rebuildPeer: rank " + std::to_string(rank_) + " received more particles in one step (" +
                           std::to_string(rec[4]) + " stored) than the cell-list launch was sized for (" + std::to_string(capacity) + ")");
        if (rec[5])
            throw SphError("SlabDecomposition::rebuildPeer: rank " + std::to_string(rank_) + " mailbox status " + std::to_string(rec[5]) +
                           " (1: more migrants than a mailbox holds, 2: a neighbour did not deliver, 4: particle storage exhausted)");
        a0_ = rec[0];
        f1_ = rec[1];
        l0_ = rec[2];
        a1_ = rec[3];
        n_ = rec[4];
        p.setTotalRealParticles(n_);
        p.setActiveRange(a0_, a1_);
        body_.setCellOrdered(true);
        body_.setPosVolDirty();
        ghost_particles_ = (uint64_t)a0_ + (n_ - a1_);
        migrated_out_ += (uint64_t)sent[0] + sent[1];
        if (PeriodicImages *im = body_.periodicImages()) im->setStoredRange(n_, a0_, a1_);
        if (checkExchangeEnabled()) verifyGhostPlanes();
        uint64_t below = 0;
        for (int r = 0; r < rank_; ++r) below += all[r];
        body_.setSlotOrigin((uint32_t)((below - a0_) & 0xffffffffull));
    }

    // once per advection step, instead of UpdateCellLinkedList::exec(). `check_planes` = false for the first of recut()'s
    // two rebuilds: between them the former owner still holds whole planes it handed over, so "ghosts" are not one plane
    // per side yet and the SPHB200_CHECK_EXCHANGE comparison would raise a false alarm (and leave the peers in NCCL).
    void rebuild(bool check_planes = true)
    {
        ExecutionInstance &ex = execution_instance();
        BaseParticles &p = body_.getBaseParticles();
        CellLinkedList &cl = body_.getCellLinkedList();
        void *st = ex.stream();
        const uint32_t n_old = a1_ - a0_;
        const int X0 = cuts_[rank_], X1 = cuts_[rank_ + 1];
        std::vector<DiscreteVariableBase *> vars = p.reorderedVariables();
        const size_t k = vars.size();
        // 1. what the neighbours need of the own slots [a0, a1): the first / last own plane and whatever moved beyond
        //    it (particles move less than one cell per advection step, so leavers sit next to the cut). One pass over the
        //    own positions (sphb200_slab_select) instead of a cell-order sort: the receiver sorts the lot anyway.
        select_.ensure((size_t)2 * n_old * sizeof(uint32_t) + 64);
        uint32_t *left_idx = select_.get<uint32_t>(), *right_idx = left_idx + n_old;
        uint32_t *d_counts = scalars_.get<uint32_t>() + 48; // bytes 192..199 of the 256-byte scratch
        SPHCK_CALL(sphb200_slab_select, &cl.mesh_, (const sphb200_vec4_t *)p.deviceData<Vecd>("Position"), a0_, n_old,
                   hasLeft() ? X0 : -1, hasRight() ? X1 - 1 : -1, left_idx, right_idx, d_counts, st);
        // 2. sizes: own counts to the host (the gathers and sends are sized by them), then swapped with the neighbours
        uint64_t *d = scalars_.get<uint64_t>();
        uint32_t sel[2] = {0, 0};
        ex.check(sphb200_copy_d2h(sel, d_counts, sizeof(sel), st), "sphb200_copy_d2h");
        ex.synchronize();
        host_syncs_ += 4; // own counts, neighbours' counts, plane offsets, slot origins
        const uint32_t send_l = sel[0], send_r = sel[1];
        uint64_t host_counts[4] = {send_l, send_r, 0, 0};
        ex.check(sphb200_copy_h2d(d, host_counts, sizeof(host_counts), st), "sphb200_copy_h2d");
        {
            const void *sl[1] = {d}, *sr[1] = {d + 1};
            void *rl[1] = {d + 2}, *rr[1] = {d + 3};
            size_t b[1] = {sizeof(uint64_t)};
            SPHCK_CALL(sphb200_comm_exchange, 1, sl, b, rl, b, sr, b, rr, b, st);
        }
        ex.check(sphb200_copy_d2h(host_counts, d, sizeof(host_counts), st), "sphb200_copy_d2h");
        // meanwhile: stage the outgoing particles of every variable in its shadow array: [0, send_l) left, then right
        std::vector<void *> stage_l(k), stage_r(k);
        std::vector<const void *> src(k);
        std::vector<uint32_t> bytes(k);
        for (size_t i = 0; i < k; ++i)
        {
            bytes[i] = vars[i]->deviceElementBytes();
            src[i] = vars[i]->deviceAddress();
            stage_l[i] = vars[i]->shadowAddress();
            stage_r[i] = (char *)stage_l[i] + (size_t)send_l * bytes[i];
        }
        if (send_l) SPHCK_CALL(sphb200_gather_multi, (int)k, stage_l.data(), src.data(), bytes.data(), left_idx, send_l, st);
        if (send_r) SPHCK_CALL(sphb200_gather_multi, (int)k, stage_r.data(), src.data(), bytes.data(), right_idx, send_r, st);
        ex.synchronize();
        const uint32_t recv_l = hasLeft() ? (uint32_t)host_counts[2] : 0, recv_r = hasRight() ? (uint32_t)host_counts[3] : 0;
        const size_t n_tot = (size_t)n_old + recv_l + recv_r;
        p.setTotalRealParticles((size_t)a0_ + n_tot); // throws if the reserved storage is exhausted
        // 3. payload: one contiguous segment per variable and direction, received right behind the own slots (the old
        //    ghosts are dropped; particles that left stay here as ghosts of their new owner)
        {
            std::vector<const void *> sl(k), sr(k);
            std::vector<void *> rl(k), rr(k);
            std::vector<size_t> bsl(k), bsr(k), brl(k), brr(k);
            for (size_t i = 0; i < k; ++i)
            {
                const size_t eb = bytes[i];
                char *base = (char *)vars[i]->deviceAddress();
                sl[i] = stage_l[i];
                bsl[i] = send_l * eb;
                sr[i] = stage_r[i];
                bsr[i] = send_r * eb;
                rl[i] = base + (size_t)a1_ * eb;
                brl[i] = recv_l * eb;
                rr[i] = base + ((size_t)a1_ + recv_l) * eb;
                brr[i] = recv_r * eb;
            }
            SPHCK_CALL(sphb200_comm_exchange, (int)k, sl.data(), bsl.data(), rl.data(), brl.data(), sr.data(), bsr.data(), rr.data(),
                       brr.data(), st);
        }
        // ring: what came in over the seam (rank 0 from its left, the last rank from its right) is the neighbour's plane
        // and leavers seen from the other side of the box: x -/+ L
        if (ring_.on)
        {
            DiscreteVariableBase *pos = p.findVariable("Position");
            char *base = (char *)pos->deviceAddress();
            const uint32_t eb = pos->deviceElementBytes();
            const Real L = ring_.upper - ring_.lower;
            if (rank_ == 0 && recv_l)
                SPHCK_CALL(sphb200_seam_shift, base + (size_t)a1_ * eb, eb, recv_l, -L, &ring_.seam, st);
            if (rank_ == nranks_ - 1 && recv_r)
                SPHCK_CALL(sphb200_seam_shift, base + ((size_t)a1_ + recv_l) * eb, eb, recv_r, L, &ring_.seam, st);
        }
        // 4. everything into cell order at the front of the arrays; own particles are the planes [X0, X1)
        reorder(a0_, (uint32_t)n_tot);
        p.setTotalRealParticles(n_tot);
        uint32_t cells2[4] = {(uint32_t)X0 * plane_cells_, (uint32_t)(X0 + 1) * plane_cells_, (uint32_t)(X1 - 1) * plane_cells_,
                              (uint32_t)X1 * plane_cells_},
                 off2[4];
        readOffsets(cells2, off2, 4);
        a0_ = off2[0];
        f1_ = off2[1];
        l0_ = off2[2];
        a1_ = off2[3];
        n_ = (uint32_t)n_tot;
        p.setActiveRange(a0_, a1_);
        body_.setCellOrdered(true);
        body_.setPosVolDirty();
        ghost_particles_ = (uint64_t)a0_ + (n_ - a1_);
        migrated_out_ += (uint64_t)send_l + send_r; // boundary-plane particles and leavers handed to the neighbours
        // periodic images of the other axes: made for all stored particles behind them, after this configuration update
        if (PeriodicImages *im = body_.periodicImages()) im->setStoredRange(n_, a0_, a1_);
        if (check_planes && checkExchangeEnabled()) verifyGhostPlanes();
        // ring: a particle that left over the seam stays here as a ghost with the x it had on THIS side of the box, while
        // every later refresh delivers fl(fl(x -/+ L) +/- L) from its new owner. Bring Position on the ghost planes to the
        // owner's value right away, so that the relation build, the summation and both half steps see ONE position.
        if (check_planes && ring_.on && !std::getenv("SPHB200_NO_RING_POSITION_REFRESH")) refreshGhosts({"Position"});
        // 5. slot origin: the slot the first stored particle has in the undecomposed run = particles owned by the ranks
        //    below minus the left ghost plane. Relations against bodies that are NOT decomposed (the wall) lay their rows
        //    out relative to it, so that summation order does not depend on the decomposition (sphb200_relation_t::bank_aligned).
        {
            uint64_t own = a1_ - a0_, *d_own = scalars_.get<uint64_t>() + 128, *d_all = d_own + 1; // bytes 1024.. : behind the reduction windows
            std::vector<uint64_t> all(nranks_);
            ex.check(sphb200_copy_h2d(d_own, &own, sizeof(own), st), "sphb200_copy_h2d");
            SPHCK_CALL(sphb200_comm_allgather_u64, d_own, d_all, 1, st);
            ex.check(sphb200_copy_d2h(all.data(), d_all, all.size() * sizeof(uint64_t), st), "sphb200_copy_d2h");
            ex.synchronize();
            uint64_t below = 0;
            for (int r = 0; r < rank_; ++r) below += all[r];
            body_.setSlotOrigin((uint32_t)((below - a0_) & 0xffffffffull));
        }
        configured_ = true;
    }

    // SPHB200_CHECK_EXCHANGE=1: after every rebuild() the ranks tell each other how many particles their boundary planes
    // hold and compare with their ghost planes. refreshGhosts() sends and receives whole planes in place, so a mismatch
    // would leave an NCCL receive waiting for bytes that never come; with the check the run stops with a message instead
    // (for first runs of a new exchange pattern; one small exchange and one host round trip per advection step).
    static bool checkExchangeEnabled()
    {
        static const bool on = [] {
            const char *e = std::getenv("SPHB200_CHECK_EXCHANGE");
            return e && e[0] != '0' && e[0] != 0;
        }();
        return on;
    }
    void verifyGhostPlanes()
    {
        ExecutionInstance &ex = execution_instance();
        void *st = ex.stream();
        uint64_t *d = scalars_.get<uint64_t>();
        uint64_t host[4] = {uint64_t(f1_ - a0_), uint64_t(a1_ - l0_), 0, 0}; // my first plane -> left, my last plane -> right
        ex.check(sphb200_copy_h2d(d, host, sizeof(host), st), "sphb200_copy_h2d");
        {
            const void *sl[1] = {d}, *sr[1] = {d + 1};
            void *rl[1] = {d + 2}, *rr[1] = {d + 3};
            size_t b[1] = {sizeof(uint64_t)};
            SPHCK_CALL(sphb200_comm_exchange, 1, sl, b, rl, b, sr, b, rr, b, st);
        }
        ex.check(sphb200_copy_d2h(host, d, sizeof(host), st), "sphb200_copy_d2h");
        ex.synchronize();
        const uint64_t left_ghosts = a0_, right_ghosts = uint64_t(n_) - a1_;
        if ((hasLeft() && host[2] != left_ghosts) || (hasRight() && host[3] != right_ghosts))
            throw SphError("SlabDecomposition: rank " + std::to_string(rank_) + " holds " + std::to_string(left_ghosts) + " / " +
                           std::to_string(right_ghosts) + " ghost particles (left / right) but its neighbours' boundary planes hold " +
                           std::to_string(host[2]) + " / " + std::to_string(host[3]));
    }

    // Re-balance: new cuts from the CURRENT particles-per-plane histogram (all ranks), then hand the planes that changed
    // owner to the neighbour. Called instead of rebuild() every `recut_interval` advection steps (the reference re-sorts
    // at that cadence, dambreak.cpp:217-220; a dam break empties the slabs near the gate and fills the ones downstream).
    // A cut moves at most to the far end of an adjacent slab, so every transfer is between neighbours and rebuild()'s
    // exchange carries it; the second rebuild() drops the copies the former owner kept and restores one-plane ghosts.
    // Results do not depend on the cuts (in-cell order is by ReferenceID), so a run with re-cuts stays bit-identical.
    void recut()
    {
        ExecutionInstance &ex = execution_instance();
        CellLinkedList &cl = body_.getCellLinkedList();
        const int planes = cl.mesh_.cells[0];
        void *st = ex.stream();
        // own particles per plane from the cell offsets at the plane boundaries (cell order: a plane is one slot range)
        std::vector<uint32_t> at(planes + 1), off(planes + 1);
        for (int x = 0; x <= planes; ++x) at[x] = (uint32_t)x * plane_cells_;
        plane_scratch_.ensure((size_t)(planes + 1) * (2 * sizeof(uint32_t) + sizeof(double)) + 64);
        uint32_t *d_at = plane_scratch_.get<uint32_t>(), *d_off = d_at + (planes + 1);
        double *d_hist = reinterpret_cast<double *>(plane_scratch_.get<char>() + (((size_t)(planes + 1) * 2 * sizeof(uint32_t) + 63) / 64) * 64);
        ex.check(sphb200_copy_h2d(d_at, at.data(), at.size() * sizeof(uint32_t), st), "sphb200_copy_h2d");
        {
            void *dst[1] = {d_off};
            const void *src[1] = {cl.cell_offset_.get<uint32_t>()};
            uint32_t eb[1] = {sizeof(uint32_t)};
            SPHCK_CALL(sphb200_gather_multi, 1, dst, src, eb, d_at, (uint32_t)(planes + 1), st);
        }
        ex.check(sphb200_copy_d2h(off.data(), d_off, off.size() * sizeof(uint32_t), st), "sphb200_copy_d2h");
        ex.synchronize();
        std::vector<double> hist(planes, 0.0);
        for (int x = cuts_[rank_]; x < cuts_[rank_ + 1]; ++x) hist[x] = double(off[x + 1] - off[x]);
        ex.check(sphb200_copy_h2d(d_hist, hist.data(), hist.size() * sizeof(double), st), "sphb200_copy_h2d");
        SPHCK_CALL(sphb200_comm_allreduce_sum_f64, d_hist, planes, st);
        ex.check(sphb200_copy_d2h(hist.data(), d_hist, hist.size() * sizeof(double), st), "sphb200_copy_d2h");
        ex.synchronize();
        // a ring is cut over the planes of the periodic box only; its seam (the first and the last cut) stays where it is
        const int p0 = ring_.on ? ring_.first_plane() : 0, p1 = ring_.on ? p0 + ring_.box_planes() : planes;
        std::vector<uint64_t> per_plane(p1 - p0);
        for (int x = p0; x < p1; ++x) per_plane[x - p0] = (uint64_t)(hist[x] + 0.5);
        std::vector<int> wanted = planSlabCuts(per_plane, nranks_);
        for (int &c : wanted) c += p0;
        std::vector<int> next = limitCutMoves(cuts_, wanted);
        ++recuts_;
        if (next == cuts_)
        {
            update();
            return;
        }
        cuts_ = next;
        rebuild(false); // hand over: ghosts are not one plane per side until the former owner has dropped its copies
        rebuild();
    }
    uint64_t recuts() const { return recuts_; }

    // refresh named variables on the ghost planes from their owners (contiguous ranges, in place)
    void refreshGhosts(std::initializer_list<const char *> names) { refreshGhosts(std::vector<std::string>(names.begin(), names.end())); }
    void refreshGhosts(const std::vector<std::string> &names)
    {
        BaseParticles &p = body_.getBaseParticles();
        const size_t k = names.size();
        std::vector<const void *> sl(k), sr(k);
        std::vector<void *> rl(k), rr(k);
        std::vector<size_t> bsl(k), bsr(k), brl(k), brr(k);
        size_t i = 0;
        for (const std::string &nm : names)
        {
            DiscreteVariableBase *v = p.findVariable(nm);
            const size_t eb = v->deviceElementBytes();
            char *base = (char *)v->deviceAddress();
            sl[i] = base + (size_t)a0_ * eb;             // my first plane -> left neighbour's right ghosts
            bsl[i] = (size_t)(f1_ - a0_) * eb;
            sr[i] = base + (size_t)l0_ * eb;             // my last plane -> right neighbour's left ghosts
            bsr[i] = (size_t)(a1_ - l0_) * eb;
            rl[i] = base;                                // left ghosts [0, a0)
            brl[i] = (size_t)a0_ * eb;
            rr[i] = base + (size_t)a1_ * eb;             // right ghosts [a1, n)
            brr[i] = (size_t)(n_ - a1_) * eb;
            ++i;
        }
        SPHCK_CALL(sphb200_comm_exchange, (int)k, sl.data(), bsl.data(), rl.data(), brl.data(), sr.data(), bsr.data(), rr.data(),
                   brr.data(), execution_instance().stream());
        // ring: records that carry a position arrive over the seam with the owner's x; shift them as rebuild() did
        if (ring_.on && (rank_ == 0 || rank_ == nranks_ - 1))
            for (const std::string &nm : names)
            {
                if (nm != "Position" && nm != "PosVol" && nm != "PosVolRef" && nm != "PosVolVel") continue;
                DiscreteVariableBase *v = p.findVariable(nm);
                char *base = (char *)v->deviceAddress();
                const uint32_t eb = v->deviceElementBytes();
                const Real L = ring_.upper - ring_.lower;
                void *st = execution_instance().stream();
                if (rank_ == 0 && a0_) SPHCK_CALL(sphb200_seam_shift, base, eb, a0_, -L, &ring_.seam, st);
                if (rank_ == nranks_ - 1 && n_ > a1_)
                    SPHCK_CALL(sphb200_seam_shift, base + (size_t)a1_ * eb, eb, n_ - a1_, L, &ring_.seam, st);
            }
    }

    Real allReduceMax(Real v)
    {
        ExecutionInstance &ex = execution_instance();
        float *d = scalars_.get<float>() + 16;
        ex.check(sphb200_copy_h2d(d, &v, sizeof(float), ex.stream()), "sphb200_copy_h2d");
        SPHCK_CALL(sphb200_comm_allreduce_max_f32, d, 1, ex.stream());
        ex.check(sphb200_copy_d2h(&v, d, sizeof(float), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
        return v;
    }
    // in place on a device float (the fused acoustic-dt slot)
    void allReduceMaxDevice(float *dev) { SPHCK_CALL(sphb200_comm_allreduce_max_f32, dev, 1, execution_instance().stream()); }
    // cell plane of an x coordinate on the body's mesh, and whether this rank owns it (observer probes are recorded by the
    // rank that owns their plane: it stores the probe's whole neighbourhood, own planes + one ghost plane on either side)
    int planeOf(Real x) const
    {
        const sphb200_mesh_t &m = body_.getCellLinkedList().mesh_;
        return hostCellCoordinate(x, m.lower[0], m.spacing, m.cells[0]);
    }
    bool ownsPlane(int plane) const { return plane >= cuts_[rank_] && plane < cuts_[rank_ + 1]; }
    void allReduceSum(double *values, int count)
    {
        if (count <= 0) return;
        if (count > 64) throw SphError("SlabDecomposition::allReduceSum: at most 64 values per call");
        ExecutionInstance &ex = execution_instance();
        double *d = scalars_.get<double>() + 32; // bytes 256..767 of the 1024-byte scratch
        ex.check(sphb200_copy_h2d(d, values, (size_t)count * sizeof(double), ex.stream()), "sphb200_copy_h2d");
        SPHCK_CALL(sphb200_comm_allreduce_sum_f64, d, count, ex.stream());
        ex.check(sphb200_copy_d2h(values, d, (size_t)count * sizeof(double), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
    }
    double allReduceSum(double v)
    {
        ExecutionInstance &ex = execution_instance();
        double *d = scalars_.get<double>() + 16;
        ex.check(sphb200_copy_h2d(d, &v, sizeof(double), ex.stream()), "sphb200_copy_h2d");
        SPHCK_CALL(sphb200_comm_allreduce_sum_f64, d, 1, ex.stream());
        ex.check(sphb200_copy_d2h(&v, d, sizeof(double), ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
        return v;
    }
};
} // namespace SPH
#endif
