// sphinxsys_ck/particles.h — DiscreteVariable / SingleVariable, BaseParticles, bodies and the SPHSystem.
//
// Storage model: every per-particle variable has ONE device array whose order follows the body's
// cell-linked list ("slot order"); UpdateCellLinkedList re-establishes that order every time it runs by one fused
// gather of all variables. The order the reference would have (ids fixed between ParticleSortCK calls) is kept as
// the `ReferenceID` variable, and host synchronisation goes through it, so host arrays (`Data()`), I/O and parity
// checks see exactly the reference's particle numbering.
//
// Reference (relative to /root/reference/src/shared):
//   DiscreteVariable / SingleVariable ... common/sphinxsys_variable.h:50-378, src_sycl/.../sphinxsys_variable_sycl.hpp
//   BaseParticles ....................... particles/base_particles.h:80-263, base_particles.cpp:33-36,84-96
//   SPHBody / RealBody / FluidBody ...... bodies/base_body.h, bodies/solid_body.h, bodies/fluid_body.h
//   materials ........................... materials/weakly_compressible_fluid.h:40-70, base_material.cpp:37-40,67-75
//   SPHSystem ........................... sphinxsys_system/sph_system.h, sph_system.cpp:39
#ifndef SPHINXSYS_CK_PARTICLES_H
#define SPHINXSYS_CK_PARTICLES_H

#include <cstring>
#include <map>

#include "geometry.h"

namespace SPH
{
template <class T> struct DeviceType;
template <> struct DeviceType<Real> { static constexpr uint32_t bytes = 4; };
template <> struct DeviceType<UnsignedInt> { static constexpr uint32_t bytes = 4; };
template <> struct DeviceType<int> { static constexpr uint32_t bytes = 4; };
template <> struct DeviceType<Vecd> { static constexpr uint32_t bytes = 16; }; // float4 on the device
template <> struct DeviceType<Matd> { static constexpr uint32_t bytes = 36; };
struct GatherRecord8 { Real v[8] = {0, 0, 0, 0, 0, 0, 0, 0}; }; // 32-byte derived record (library extension)
template <> struct DeviceType<GatherRecord8> { static constexpr uint32_t bytes = 32; };

inline void *deviceAllocate(size_t bytes)
{
    void *p = nullptr;
    execution_instance().ctx(); // selects the device
    int rc = sphb200_malloc_device(&p, bytes);
    if (rc != 0) throw SphError("sphb200_malloc_device(" + std::to_string(bytes) + ") failed: " + std::to_string(rc));
    return p;
}

// RAII device array
class DeviceBuffer
{
    void *p_ = nullptr;
    size_t bytes_ = 0;

  public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t bytes) { reset(bytes); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer() { release(); }
    void release()
    {
        if (p_) sphb200_free_device(p_);
        p_ = nullptr;
        bytes_ = 0;
    }
    void reset(size_t bytes)
    {
        release();
        p_ = deviceAllocate(bytes ? bytes : 4);
        bytes_ = bytes;
    }
    // grow without copy (DiscreteVariable::reallocateData semantics). A buffer that has to grow AGAIN gets an eighth of
    // head-room: sizes that follow a particle count creep up by a few elements per step in decomposed and periodic runs,
    // and every cudaFree / cudaMalloc pair synchronises the device
    void ensure(size_t bytes)
    {
        if (bytes > bytes_) reset(bytes_ ? bytes + bytes / 8 : bytes);
    }
    void swap(DeviceBuffer &o)
    {
        std::swap(p_, o.p_);
        std::swap(bytes_, o.bytes_);
    }
    template <class T = void> T *get() const { return (T *)p_; }
    size_t bytes() const { return bytes_; }
};

class BaseParticles;

class DiscreteVariableBase
{
  protected:
    std::string name_;
    size_t size_;        // allocated elements (particles bound + 1)
    uint32_t elem_bytes_; // device element size
    DeviceBuffer dev_, shadow_;
    BaseParticles *particles_ = nullptr;

  public:
    DiscreteVariableBase(const std::string &name, size_t size, uint32_t elem_bytes, BaseParticles *particles)
        : name_(name), size_(size), elem_bytes_(elem_bytes), dev_(size * elem_bytes), particles_(particles) {}
    virtual ~DiscreteVariableBase() {}
    const std::string &Name() const { return name_; }
    size_t getDataSize() const { return size_; }
    uint32_t deviceElementBytes() const { return elem_bytes_; }
    void *deviceAddress() const { return dev_.get(); }
    void *shadowAddress()
    {
        shadow_.ensure(size_ * elem_bytes_);
        return shadow_.get();
    }
    void swapWithShadow() { dev_.swap(shadow_); }
    void fillDeviceZero()
    {
        std::vector<char> z(size_ * elem_bytes_, 0);
        execution_instance().check(sphb200_copy_h2d(dev_.get(), z.data(), z.size(), execution_instance().stream()), "copy_h2d");
        execution_instance().synchronize();
    }
};

template <class T> class DiscreteVariable : public DiscreteVariableBase
{
    std::vector<T> host_; // reference particle order; valid after synchronizeWithDevice()

  public:
    DiscreteVariable(const std::string &name, size_t size, BaseParticles *particles)
        : DiscreteVariableBase(name, size, DeviceType<T>::bytes, particles) {}
    // host view (created on demand), reference order
    T *Data()
    {
        if (host_.size() != size_) host_.assign(size_, T());
        return host_.data();
    }
    // device view: Real*/UnsignedInt* as is, Vecd as sphb200_vec4_t*, Matd as 9 packed floats
    template <class Policy> void *DelegatedData(const Policy &)
    {
        execution::require_device_policy<Policy>();
        return deviceAddress();
    }
    void synchronizeToDevice();   // host (reference order) -> device (slot order)
    void synchronizeWithDevice(); // device -> host
    // sphinxsys_variable.h:350-356: what the output / reload paths call; a device policy is the only one there is here
    template <class Policy> void prepareForOutput(const Policy &)
    {
        execution::require_device_policy<Policy>();
        synchronizeWithDevice();
    }
    template <class Policy> void finalizeLoadIn(const Policy &)
    {
        execution::require_device_policy<Policy>();
        synchronizeToDevice();
    }
};

// SingleVariable<T>: a named scalar living on the host, mirrored to the device on demand (sphinxsys_variable.h:50-120)
template <class T> class SingleVariable
{
    std::string name_;
    T value_;

  public:
    SingleVariable(const std::string &name, const T &v) : name_(name), value_(v) {}
    const std::string &Name() const { return name_; }
    T getValue() const { return value_; }
    void setValue(const T &v) { value_ = v; }
    void incrementValue(const T &v) { value_ += v; }
};

// ---------------------------------------------------------------------------------------------------------
class BaseParticles
{
    size_t total_real_particles_, particles_bound_;
    std::map<std::string, std::unique_ptr<DiscreteVariableBase>> all_variables_;
    std::vector<DiscreteVariableBase *> ordered_; // registration order (the order the fused gather uses)
    std::vector<DiscreteVariableBase *> evolving_variables_;
    std::vector<std::string> derived_;            // variables rebuilt after a reorder instead of being gathered
    DiscreteVariable<UnsignedInt> *dv_reference_id_ = nullptr;
    DeviceBuffer inverse_, staging_;
    uint64_t storage_version_ = 1, inverse_version_ = 0;
    bool identity_order_ = true; // ReferenceID == iota (no reorder happened yet)

    size_t active_begin_ = 0, active_end_ = 0; // slots the dynamics update (own particles of a decomposed run)

  public:
    // n particles now, room for `bound` (>= n): ghost and migrated particles of a decomposed run live behind / around
    // the real ones, as the reference keeps buffer and ghost particles behind the real ones (base_particles.h:67-72)
    explicit BaseParticles(size_t n, size_t bound = 0) : total_real_particles_(n), particles_bound_(std::max(n, bound))
    {
        active_end_ = n;
        dv_reference_id_ = registerStateVariable<UnsignedInt>("ReferenceID");
        SPHCK_CALL(sphb200_iota_u32, (uint32_t *)dv_reference_id_->deviceAddress(), particles_bound_ + 1, execution_instance().stream());
    }
    // number of stored particles (own + ghosts in a decomposed run)
    size_t TotalRealParticles() const { return total_real_particles_; }
    size_t ParticlesBound() const { return particles_bound_; }
    void setTotalRealParticles(size_t n)
    {
        if (n > particles_bound_) throw SphError("particle storage exhausted: " + std::to_string(n) + " > bound " + std::to_string(particles_bound_));
        total_real_particles_ = n;
    }
    // a new particle set in the same storage (static bodies re-cut by the slab decomposition): n particles in the order
    // the caller uploads them next, i.e. host order == slot order again until the next reorder
    void resetContents(size_t n)
    {
        setTotalRealParticles(n);
        active_begin_ = 0;
        active_end_ = n;
        identity_order_ = true;
        ++storage_version_;
    }
    size_t activeBegin() const { return active_begin_; }
    size_t activeEnd() const { return active_end_; }
    void setActiveRange(size_t b, size_t e)
    {
        active_begin_ = b;
        active_end_ = e;
    }
    // particles the host sees through DiscreteVariable::Data(): the real ones. Image (ghost) particles of a periodic
    // body are stored behind them ([activeEnd, TotalRealParticles)) and never cross to the host.
    size_t hostSyncCount() const { return active_begin_ == 0 && active_end_ <= total_real_particles_ ? active_end_ : total_real_particles_; }
    const std::vector<DiscreteVariableBase *> &allVariablesInOrder() const { return ordered_; }
    uint64_t storageVersion() const { return storage_version_; }
    bool identityOrder() const { return identity_order_; }

    template <class T> DiscreteVariable<T> *registerStateVariable(const std::string &name, const T &init = T())
    {
        auto it = all_variables_.find(name);
        if (it != all_variables_.end())
        {
            auto *v = dynamic_cast<DiscreteVariable<T> *>(it->second.get());
            if (!v) throw SphError("registerStateVariable: '" + name + "' already registered with another type");
            return v;
        }
        auto *v = new DiscreteVariable<T>(name, particles_bound_ + 1, this);
        all_variables_[name].reset(v);
        ordered_.push_back(v);
        fillDevice(v, init);
        return v;
    }
    template <class T> DiscreteVariable<T> *getVariableByName(const std::string &name)
    {
        auto it = all_variables_.find(name);
        DiscreteVariable<T> *v = it == all_variables_.end() ? nullptr : dynamic_cast<DiscreteVariable<T> *>(it->second.get());
        if (!v)
        {
            // sphinxsys_variable.h:252-257
            std::cout << "\n Error: the variable '" << name << "' is not registered!" << std::endl;
            throw SphError("the variable '" + name + "' is not registered");
        }
        return v;
    }
    bool hasVariable(const std::string &name) const { return all_variables_.count(name) != 0; }
    template <class T> void *deviceData(const std::string &name) { return getVariableByName<T>(name)->deviceAddress(); }
    template <class T> void *deviceDataOrNull(const std::string &name)
    {
        return hasVariable(name) ? getVariableByName<T>(name)->deviceAddress() : nullptr;
    }
    template <class T> void addEvolvingVariable(const std::string &name)
    {
        DiscreteVariableBase *v = getVariableByName<T>(name);
        for (auto *e : evolving_variables_)
            if (e == v) return;
        evolving_variables_.push_back(v);
    }
    const std::vector<DiscreteVariableBase *> &EvolvingVariables() const { return evolving_variables_; }
    bool isEvolving(const DiscreteVariableBase *v) const
    {
        for (auto *e : evolving_variables_)
            if (e == v) return true;
        return false;
    }
    // a derived variable is recomputed by its owner after a storage reorder (it is skipped by the gather)
    void markDerived(const std::string &name) { derived_.push_back(name); }
    std::vector<DiscreteVariableBase *> reorderedVariables() const
    {
        std::vector<DiscreteVariableBase *> out;
        for (auto *v : ordered_)
        {
            bool skip = false;
            for (auto &d : derived_)
                if (d == v->Name()) skip = true;
            if (!skip) out.push_back(v);
        }
        return out;
    }
    void storageReordered()
    {
        ++storage_version_;
        identity_order_ = false;
    }
    // raw copy of the slots [begin, begin + count) of a variable (storage order; Vecd comes as float4)
    void downloadRaw(DiscreteVariableBase *v, void *host, size_t begin, size_t count)
    {
        ExecutionInstance &ex = execution_instance();
        const uint32_t eb = v->deviceElementBytes();
        if (count) ex.check(sphb200_copy_d2h(host, (const char *)v->deviceAddress() + begin * eb, count * eb, ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
    }
    DiscreteVariableBase *findVariable(const std::string &name)
    {
        auto it = all_variables_.find(name);
        if (it == all_variables_.end()) throw SphError("the variable '" + name + "' is not registered");
        return it->second.get();
    }
    uint32_t *referenceID() { return (uint32_t *)dv_reference_id_->deviceAddress(); }
    DiscreteVariable<UnsignedInt> *referenceIDVariable() { return dv_reference_id_; }

    // slot of reference id r (inverse of ReferenceID), cached per storage version
    uint32_t *inverseReferenceID()
    {
        uint32_t n = (uint32_t)hostSyncCount();
        inverse_.ensure((n + 1) * sizeof(uint32_t));
        if (inverse_version_ != storage_version_)
        {
            SPHCK_CALL(sphb200_update_sorted_id, referenceID(), inverse_.get<uint32_t>(), n, execution_instance().stream());
            inverse_version_ = storage_version_;
        }
        return inverse_.get<uint32_t>();
    }

    // ---- host <-> device, through the reference order ----
    template <class T> void upload(DiscreteVariable<T> *v, const T *host)
    {
        uint32_t n = (uint32_t)hostSyncCount();
        if (n == 0) return;
        ExecutionInstance &ex = execution_instance();
        const uint32_t eb = v->deviceElementBytes();
        const size_t host_bytes = sizeof(T) * (size_t)n;
        const size_t raw_bytes = ((host_bytes + 63) / 64) * 64;
        staging_.ensure(raw_bytes + (size_t)eb * n + 64);
        char *raw = staging_.get<char>();
        char *conv = raw + raw_bytes;
        ex.check(sphb200_copy_h2d(raw, host, host_bytes, ex.stream()), "sphb200_copy_h2d");
        void *src = raw;
        if (std::is_same<T, Vecd>::value)
        {
            SPHCK_CALL(sphb200_vec3_to_vec4, (sphb200_vec4_t *)conv, (const float *)raw, n, ex.stream());
            src = conv;
        }
        if (identity_order_)
            ex.check(sphb200_copy_d2d(v->deviceAddress(), src, (size_t)eb * n, ex.stream()), "sphb200_copy_d2d");
        else
        {
            void *dst[1] = {v->deviceAddress()};
            const void *srcs[1] = {src};
            uint32_t bytes[1] = {eb};
            SPHCK_CALL(sphb200_gather_multi, 1, dst, srcs, bytes, referenceID(), n, ex.stream());
        }
        // matrices / pressures written by the host: their gather record follows (fluid_dynamics.h, registerCorrectionRecord)
        if (hasVariable("LinearCorrectionRecord"))
        {
            if (std::is_same<T, Matd>::value && v->Name() == "LinearCorrectionMatrix")
                SPHCK_CALL(sphb200_pack_correction_records, n, (const float *)v->deviceAddress(), nullptr,
                           deviceData<GatherRecord8>("LinearCorrectionRecord"), ex.stream());
            if (std::is_same<T, Real>::value && v->Name() == "Pressure")
                SPHCK_CALL(sphb200_pack_correction_records, n, nullptr, (const float *)v->deviceAddress(),
                           deviceData<GatherRecord8>("LinearCorrectionRecord"), ex.stream());
        }
        ex.synchronize(); // the host buffer may be pageable
    }
    template <class T> void download(DiscreteVariable<T> *v, T *host)
    {
        uint32_t n = (uint32_t)hostSyncCount();
        if (n == 0) return;
        ExecutionInstance &ex = execution_instance();
        const uint32_t eb = v->deviceElementBytes();
        const size_t a = (((size_t)eb * n + 63) / 64) * 64;
        staging_.ensure(a + sizeof(T) * (size_t)n + 64);
        char *ordered = staging_.get<char>();
        char *packed = ordered + a;
        const void *src = v->deviceAddress();
        if (!identity_order_)
        {
            void *dst[1] = {ordered};
            const void *srcs[1] = {src};
            uint32_t bytes[1] = {eb};
            SPHCK_CALL(sphb200_gather_multi, 1, dst, srcs, bytes, inverseReferenceID(), n, ex.stream());
            src = ordered;
        }
        if (std::is_same<T, Vecd>::value)
        {
            SPHCK_CALL(sphb200_vec4_to_vec3, (float *)packed, (const sphb200_vec4_t *)src, n, ex.stream());
            src = packed;
        }
        ex.check(sphb200_copy_d2h(host, src, sizeof(T) * (size_t)n, ex.stream()), "sphb200_copy_d2h");
        ex.synchronize();
    }

  private:
    template <class T> void fillDevice(DiscreteVariable<T> *v, const T &init)
    {
        std::vector<T> h(v->getDataSize(), init);
        size_t n_keep = total_real_particles_;
        // fill all allocated elements (including the padding entry) directly, order-independent
        ExecutionInstance &ex = execution_instance();
        if (std::is_same<T, Vecd>::value)
        {
            std::vector<float> h4(4 * h.size());
            for (size_t i = 0; i < h.size(); ++i)
            {
                const Vecd &p = *(const Vecd *)&h[i];
                h4[4 * i] = p.x; h4[4 * i + 1] = p.y; h4[4 * i + 2] = p.z; h4[4 * i + 3] = 0;
            }
            ex.check(sphb200_copy_h2d(v->deviceAddress(), h4.data(), h4.size() * 4, ex.stream()), "sphb200_copy_h2d");
            ex.synchronize();
        }
        else
        {
            ex.check(sphb200_copy_h2d(v->deviceAddress(), h.data(), h.size() * sizeof(T), ex.stream()), "sphb200_copy_h2d");
            ex.synchronize();
        }
        (void)n_keep;
    }
};

// ---------------------------------------------------------------------------------------------------------
// HostTransferPipeline (extension): synchronizeToDevice / synchronizeWithDevice of a fixed set of variables, split into
// an asynchronous copy on a side stream and a device-side re-ordering step on the main stream, so that the host <->
// device traffic of step s+1 / s-1 overlaps the dynamics of step s. Host buffers must be pinned and hold the variables
// in the reference particle order and layout (Vecd = 3 Reals), exactly what DiscreteVariable::Data() holds.
//   stageUploads(host...)   side stream : waits until the previous commit has consumed the staging, H2D of every input
//   commitUploads()         main stream : waits for the H2D, Vecd 3->4 conversion + gather into the slot order
//   stageDownloads(host...) main stream : gather into the reference order (+ 4->3), then side stream: D2H
//   synchronize()           host waits for the side stream
// ---------------------------------------------------------------------------------------------------------
class HostTransferPipeline
{
    struct Item
    {
        DiscreteVariableBase *v;
        bool is_vec;
        size_t host_elem_bytes;
        DeviceBuffer raw, conv;
    };
    BaseParticles &p_;
    std::vector<std::unique_ptr<Item>> ins_, outs_;
    // slab-decomposed bodies: the host holds THIS RANK's own particles in storage (slot) order — what SlabDecomposition
    // hands out with the ReferenceID array — in the reference's packed layout (Vecd = 3 floats: a quarter fewer bytes over
    // PCIe than the device's float4); the copies go between the pinned buffers and the own slot range
    // [activeBegin, activeEnd) with the 3 <-> 4 conversion on the device, no reordering passes
    bool raw_own_slots_ = false;
    void *copy_stream_ = nullptr, *ev_h2d_ = nullptr, *ev_commit_ = nullptr, *ev_out_ready_ = nullptr, *ev_d2h_ = nullptr;

    template <class T> static Item *makeItem(DiscreteVariable<T> *v)
    {
        Item *it = new Item();
        it->v = v;
        it->is_vec = std::is_same<T, Vecd>::value;
        it->host_elem_bytes = sizeof(T);
        return it;
    }

  public:
    explicit HostTransferPipeline(BaseParticles &particles) : p_(particles)
    {
        ExecutionInstance &ex = execution_instance();
        ex.ctx();
        ex.check(sphb200_stream_create(&copy_stream_), "sphb200_stream_create");
        for (void **e : {&ev_h2d_, &ev_commit_, &ev_out_ready_, &ev_d2h_}) ex.check(sphb200_event_create(e), "sphb200_event_create");
    }
    HostTransferPipeline(const HostTransferPipeline &) = delete;
    ~HostTransferPipeline()
    {
        for (void *e : {ev_h2d_, ev_commit_, ev_out_ready_, ev_d2h_})
            if (e) sphb200_event_destroy(e);
        if (copy_stream_) sphb200_stream_destroy(copy_stream_);
    }
    template <class T> void addInput(DiscreteVariable<T> *v) { ins_.emplace_back(makeItem(v)); }
    template <class T> void addOutput(DiscreteVariable<T> *v) { outs_.emplace_back(makeItem(v)); }
    size_t inputs() const { return ins_.size(); }
    size_t outputs() const { return outs_.size(); }
    void setRawOwnSlots(bool on) { raw_own_slots_ = on; }
    bool rawOwnSlots() const { return raw_own_slots_; }
    size_t count() const { return raw_own_slots_ ? p_.activeEnd() - p_.activeBegin() : p_.hostSyncCount(); }
    size_t hostElementBytes(const Item &it) const { return it.host_elem_bytes; }
    size_t inputBytes() const
    {
        size_t b = 0;
        for (auto &it : ins_) b += hostElementBytes(*it) * count();
        return b;
    }
    size_t outputBytes() const
    {
        size_t b = 0;
        for (auto &it : outs_) b += hostElementBytes(*it) * count();
        return b;
    }
    void stageUploads(const void *const *pinned_host)
    {
        ExecutionInstance &ex = execution_instance();
        const size_t n = count();
        ex.check(sphb200_stream_wait_event(copy_stream_, ev_commit_), "sphb200_stream_wait_event");
        for (size_t k = 0; k < ins_.size(); ++k)
        {
            Item &it = *ins_[k];
            const size_t eb = hostElementBytes(it);
            it.raw.ensure(eb * n + 64);
            ex.check(sphb200_copy_h2d(it.raw.get(), pinned_host[k], eb * n, copy_stream_), "sphb200_copy_h2d");
        }
        ex.check(sphb200_event_record(ev_h2d_, copy_stream_), "sphb200_event_record");
    }
    void commitUploads()
    {
        ExecutionInstance &ex = execution_instance();
        void *st = ex.stream();
        const uint32_t n = (uint32_t)count();
        ex.check(sphb200_stream_wait_event(st, ev_h2d_), "sphb200_stream_wait_event");
        if (raw_own_slots_)
        {
            // the staging already is in slot order: one conversion / device copy per variable into the own slot range
            for (auto &ip : ins_)
            {
                const size_t eb = ip->v->deviceElementBytes();
                char *own = (char *)ip->v->deviceAddress() + p_.activeBegin() * eb;
                if (ip->is_vec) SPHCK_CALL(sphb200_vec3_to_vec4, (sphb200_vec4_t *)own, (const float *)ip->raw.get(), n, st);
                else ex.check(sphb200_copy_d2d(own, ip->raw.get(), eb * n, st), "sphb200_copy_d2d");
            }
            ex.check(sphb200_event_record(ev_commit_, st), "sphb200_event_record");
            return;
        }
        std::vector<void *> dst;
        std::vector<const void *> src;
        std::vector<uint32_t> bytes;
        for (auto &ip : ins_)
        {
            Item &it = *ip;
            const void *from = it.raw.get();
            if (it.is_vec)
            {
                it.conv.ensure((size_t)16 * n + 64);
                SPHCK_CALL(sphb200_vec3_to_vec4, (sphb200_vec4_t *)it.conv.get(), (const float *)it.raw.get(), n, st);
                from = it.conv.get();
            }
            dst.push_back(it.v->deviceAddress());
            src.push_back(from);
            bytes.push_back(it.v->deviceElementBytes());
        }
        if (p_.identityOrder())
            for (size_t k = 0; k < dst.size(); ++k) ex.check(sphb200_copy_d2d(dst[k], src[k], (size_t)bytes[k] * n, st), "sphb200_copy_d2d");
        else if (!dst.empty())
            SPHCK_CALL(sphb200_gather_multi, (int)dst.size(), dst.data(), src.data(), bytes.data(), p_.referenceID(), n, st);
        ex.check(sphb200_event_record(ev_commit_, st), "sphb200_event_record");
    }
    void stageDownloads(void *const *pinned_host)
    {
        ExecutionInstance &ex = execution_instance();
        void *st = ex.stream();
        const uint32_t n = (uint32_t)count();
        ex.check(sphb200_stream_wait_event(st, ev_d2h_), "sphb200_stream_wait_event"); // the previous D2H has left the staging
        if (raw_own_slots_)
        {
            for (auto &op : outs_)
            {
                const size_t eb = op->v->deviceElementBytes();
                const char *own = (const char *)op->v->deviceAddress() + p_.activeBegin() * eb;
                op->raw.ensure(op->host_elem_bytes * n + 64);
                if (op->is_vec) SPHCK_CALL(sphb200_vec4_to_vec3, (float *)op->raw.get(), (const sphb200_vec4_t *)own, n, st);
                else ex.check(sphb200_copy_d2d(op->raw.get(), own, eb * n, st), "sphb200_copy_d2d");
            }
            ex.check(sphb200_event_record(ev_out_ready_, st), "sphb200_event_record");
            ex.check(sphb200_stream_wait_event(copy_stream_, ev_out_ready_), "sphb200_stream_wait_event");
            for (size_t k = 0; k < outs_.size(); ++k)
                ex.check(sphb200_copy_d2h(pinned_host[k], outs_[k]->raw.get(), outs_[k]->host_elem_bytes * n, copy_stream_), "sphb200_copy_d2h");
            ex.check(sphb200_event_record(ev_d2h_, copy_stream_), "sphb200_event_record");
            return;
        }
        std::vector<void *> dst;
        std::vector<const void *> src;
        std::vector<uint32_t> bytes;
        for (auto &op : outs_)
        {
            Item &it = *op;
            it.raw.ensure((size_t)it.v->deviceElementBytes() * n + 64);
            dst.push_back(it.raw.get());
            src.push_back(it.v->deviceAddress());
            bytes.push_back(it.v->deviceElementBytes());
        }
        if (p_.identityOrder())
            for (size_t k = 0; k < dst.size(); ++k) ex.check(sphb200_copy_d2d(dst[k], src[k], (size_t)bytes[k] * n, st), "sphb200_copy_d2d");
        else if (!dst.empty())
            SPHCK_CALL(sphb200_gather_multi, (int)dst.size(), dst.data(), src.data(), bytes.data(), p_.inverseReferenceID(), n, st);
        for (auto &op : outs_)
            if (op->is_vec)
            {
                op->conv.ensure((size_t)12 * n + 64);
                SPHCK_CALL(sphb200_vec4_to_vec3, (float *)op->conv.get(), (const sphb200_vec4_t *)op->raw.get(), n, st);
            }
        ex.check(sphb200_event_record(ev_out_ready_, st), "sphb200_event_record");
        ex.check(sphb200_stream_wait_event(copy_stream_, ev_out_ready_), "sphb200_stream_wait_event");
        for (size_t k = 0; k < outs_.size(); ++k)
        {
            Item &it = *outs_[k];
            ex.check(sphb200_copy_d2h(pinned_host[k], it.is_vec ? it.conv.get() : it.raw.get(), it.host_elem_bytes * n, copy_stream_), "sphb200_copy_d2h");
        }
        ex.check(sphb200_event_record(ev_d2h_, copy_stream_), "sphb200_event_record");
    }
    void synchronize() { execution_instance().check(sphb200_stream_sync(copy_stream_), "sphb200_stream_sync"); }
};

template <class T> void DiscreteVariable<T>::synchronizeToDevice() { particles_->upload(this, Data()); }
template <class T> void DiscreteVariable<T>::synchronizeWithDevice() { particles_->download(this, Data()); }

// ---------------------------------------------------------------------------------------------------------
// materials
// ---------------------------------------------------------------------------------------------------------
class BaseMaterial
{
  public:
    virtual ~BaseMaterial() {}
    virtual Real ReferenceDensity() const = 0;
};
class WeaklyCompressibleFluid : public BaseMaterial
{
  public:
    Real rho0_, c0_, p0_;
    WeaklyCompressibleFluid(Real rho0, Real c0) : rho0_(rho0), c0_(c0), p0_(rho0 * c0 * c0) {}
    Real ReferenceDensity() const override { return rho0_; }
    Real ReferenceSoundSpeed() const { return c0_; }
};
// Viscosity (materials/viscosity.h:40-67): reference viscosity mu; a fluid body carries it through
// defineClosure<WeaklyCompressibleFluid, Viscosity>(rho0, c0, mu) (base_body.h defineClosure, closure.h)
class Viscosity
{
    Real mu_;

  public:
    explicit Viscosity(Real mu) : mu_(mu) {}
    Real ReferenceViscosity() const { return mu_; }
};
template <class MaterialType, class ViscosityType> class Closure : public MaterialType, public ViscosityType
{
  public:
    template <class A, class B, class C> Closure(A rho0, B c0, C mu) : MaterialType(Real(rho0), Real(c0)), ViscosityType(Real(mu)) {}
};
class Solid : public BaseMaterial
{
  public:
    Real rho0_;
    explicit Solid(Real rho0 = 1.0) : rho0_(rho0) {}
    Real ReferenceDensity() const override { return rho0_; }
};

// ---------------------------------------------------------------------------------------------------------
// SPHSystem and bodies
// ---------------------------------------------------------------------------------------------------------
class SPHBody;
struct Lattice {};

class SPHSystem
{
  public:
    BoundingBoxd system_domain_bounds_;
    Real global_resolution_;
    int dim_;
    std::vector<SPHBody *> bodies_;
    std::map<std::string, std::unique_ptr<SingleVariable<Real>>> system_variables_;
    // sph_system.cpp:39: the case bounds are expanded by 4 dp on every side
    SPHSystem(const BoundingBoxd &case_bounds, Real resolution, int dim = 3) : global_resolution_(resolution), dim_(dim)
    {
        system_domain_bounds_ = case_bounds;
        for (int d = 0; d < dim; ++d)
        {
            // evaluated in double and rounded once, as the case files' double literals are
            system_domain_bounds_.lower_[d] = Real(double(case_bounds.lower_[d]) - 4.0 * double(resolution));
            system_domain_bounds_.upper_[d] = Real(double(case_bounds.upper_[d]) + 4.0 * double(resolution));
        }
        system_variables_["PhysicalTime"].reset(new SingleVariable<Real>("PhysicalTime", 0));
    }
    void setSystemDomainBoundsExact(const BoundingBoxd &b) { system_domain_bounds_ = b; }
    template <class T> SingleVariable<T> *getSystemVariableByName(const std::string &name)
    {
        auto it = system_variables_.find(name);
        if (it == system_variables_.end()) throw SphError("system variable '" + name + "' not found");
        return it->second.get();
    }
    int Dimensions() const { return dim_; }
};

class CellLinkedList;
class PeriodicImages;
template <typename... T> class Ghost;

class SPHBody
{
  protected:
    SPHSystem &sph_system_;
    std::string name_;
    std::shared_ptr<ComplexShape> shape_;
    std::unique_ptr<SPHAdaptation> adaptation_;
    std::unique_ptr<BaseParticles> particles_;
    std::unique_ptr<BaseMaterial> material_;
    std::unique_ptr<CellLinkedList> cell_linked_list_;
    std::unique_ptr<PeriodicImages> periodic_images_; // periodic_images.h; null for bodies without periodic conditions
    size_t particle_reserve_ = 0;                     // room for ghost particles behind the real ones
    bool posvol_dirty_ = true;
    uint32_t slot_origin_ = 0;
    bool cell_ordered_ = false;

  public:
    SPHBody(SPHSystem &system, std::shared_ptr<ComplexShape> shape)
        : sph_system_(system), name_(shape->Name()), shape_(shape),
          adaptation_(new SPHAdaptation(system.global_resolution_, system.dim_))
    {
        system.bodies_.push_back(this);
    }
    SPHBody(SPHSystem &system, const std::string &name)
        : sph_system_(system), name_(name), adaptation_(new SPHAdaptation(system.global_resolution_, system.dim_))
    {
        system.bodies_.push_back(this);
    }
    virtual ~SPHBody();
    const std::string &Name() const { return name_; }
    SPHSystem &getSPHSystem() { return sph_system_; }
    SPHAdaptation &getSPHAdaptation() { return *adaptation_; }
    BaseParticles &getBaseParticles()
    {
        if (!particles_) throw SphError("body '" + name_ + "': particles not generated");
        return *particles_;
    }
    BaseMaterial &getBaseMaterial() { return *material_; }
    ComplexShape &getInitialShape() { return *shape_; }
    size_t TotalRealParticles() { return getBaseParticles().TotalRealParticles(); }
    CellLinkedList &getCellLinkedList();
    PeriodicImages *periodicImages() { return periodic_images_.get(); }
    PeriodicImages &definePeriodicImages();
    // the bounds of the body shape (base_body.cpp: getSPHBodyBounds)
    BoundingBoxd getSPHBodyBounds() { return shape_->getBounds(); }
    void reserveParticles(size_t extra) { particle_reserve_ += extra; }
    template <class MaterialType, typename... Args> MaterialType *defineMatterMaterial(Args &&...args)
    {
        MaterialType *m = new MaterialType(std::forward<Args>(args)...);
        material_.reset(m);
        return m;
    }
    // defineClosure<WeaklyCompressibleFluid, Viscosity>(rho0, c0, mu): the fluid material with its viscosity model
    template <class MaterialType, class ViscosityType, typename... Args> Closure<MaterialType, ViscosityType> *defineClosure(Args &&...args)
    {
        auto *m = new Closure<MaterialType, ViscosityType>(std::forward<Args>(args)...);
        material_.reset(m);
        return m;
    }
    // generateParticles<BaseParticles, Lattice>(): lattice points inside the body shape
    template <class ParticlesType, class Generator> void generateParticles()
    {
        std::vector<Vecd> pos = generateLattice(*shape_, sph_system_.system_domain_bounds_, sph_system_.global_resolution_, sph_system_.dim_);
        Real vol = Real(std::pow(sph_system_.global_resolution_, Real(sph_system_.dim_)));
        generateParticlesFromPositions(pos, vol);
    }
    // generateParticlesWithReserve<BaseParticles, Lattice>(ghost_x, ghost_y, ...): room for the ghost particles of the
    // periodic conditions is reserved behind the real particles (base_body.h generateParticlesWithReserve,
    // particle_reserve.h, ghost_bounding.cpp:9-26)
    template <class ParticlesType, class Generator, class... Reserves> void generateParticlesWithReserve(Reserves &...reserves)
    {
        std::vector<Vecd> pos = generateLattice(*shape_, sph_system_.system_domain_bounds_, sph_system_.global_resolution_, sph_system_.dim_);
        Real vol = Real(std::pow(sph_system_.global_resolution_, Real(sph_system_.dim_)));
        reserveFor(reserves...);
        generateParticlesFromPositions(pos, vol);
    }
    // the same reservation for particles handed over by the caller (reload files, harness-generated lattices)
    template <class... Reserves> void reserveFor(Reserves &...reserves)
    {
        size_t dummy[] = {0, (particle_reserve_ += reserves.reserveSize(sph_system_.global_resolution_, sph_system_.dim_), size_t(0))...};
        (void)dummy;
        int dummy2[] = {0, (reserves.setReserved(), 0)...};
        (void)dummy2;
    }
    // positions handed over by the caller (e.g. a reload file): base_particles.cpp:33-36, base_material.cpp:37-40
    // `bound` reserves room for migrated and ghost particles; `reference_ids` (optional) are the global particle
    // numbers of a decomposed run (default: 0..n-1)
    void generateParticlesFromPositions(const std::vector<Vecd> &pos, Real vol, size_t bound = 0,
                                        const std::vector<UnsignedInt> *reference_ids = nullptr)
    {
        size_t n = pos.size();
        if (bound == 0 && particle_reserve_) bound = n + particle_reserve_;
        particles_.reset(new BaseParticles(n, bound));
        BaseParticles &p = *particles_;
        auto *dv_pos = p.registerStateVariable<Vecd>("Position");
        auto *dv_vol = p.registerStateVariable<Real>("VolumetricMeasure", vol);
        p.registerStateVariable<Vecd>("PosVol"); // derived gather record (x, y, z, Vol)
        p.markDerived("PosVol");
        p.registerStateVariable<Vecd>("PosVolRef"); // (x, y, z, VolRef)
        p.markDerived("PosVolRef");
        Real rho0 = material_ ? material_->ReferenceDensity() : Real(1);
        p.registerStateVariable<Real>("Density", rho0);
        p.registerStateVariable<Real>("Mass", rho0 * vol);
        auto *dv_oid = p.registerStateVariable<UnsignedInt>("OriginalID");
        SPHCK_CALL(sphb200_iota_u32, (uint32_t *)dv_oid->deviceAddress(), n + 1, execution_instance().stream());
        if (reference_ids && n)
        {
            ExecutionInstance &ex = execution_instance();
            ex.check(sphb200_copy_h2d(p.referenceID(), reference_ids->data(), n * sizeof(UnsignedInt), ex.stream()), "sphb200_copy_h2d");
            ex.check(sphb200_copy_h2d(dv_oid->deviceAddress(), reference_ids->data(), n * sizeof(UnsignedInt), ex.stream()), "sphb200_copy_h2d");
            ex.synchronize();
        }
        p.addEvolvingVariable<Vecd>("Position");
        p.addEvolvingVariable<Real>("VolumetricMeasure");
        p.addEvolvingVariable<UnsignedInt>("OriginalID");
        if (n) p.upload(dv_pos, pos.data());
        (void)dv_vol;
        posvol_dirty_ = true;
    }
    // (x, y, z, Vol) gather record; refreshed lazily whenever Position or VolumetricMeasure changed
    void setPosVolDirty() { posvol_dirty_ = true; }
    bool recordsDirty() const { return posvol_dirty_; }
    void refreshPosVol()
    {
        if (!posvol_dirty_) return;
        BaseParticles &p = getBaseParticles();
        const bool has_ref = p.hasVariable("VolumetricMeasureRef"), has_vel = p.hasVariable("Velocity");
        if (has_vel && !p.hasVariable("PosVolVel"))
        {
            p.registerStateVariable<GatherRecord8>("PosVolVel"); // (x, y, z, Vol, vx, vy, vz, -)
            p.markDerived("PosVolVel");
        }
        SPHCK_CALL(sphb200_pack_records, (uint32_t)p.TotalRealParticles(), (const sphb200_vec4_t *)p.deviceData<Vecd>("Position"),
                   (const float *)p.deviceData<Real>("VolumetricMeasure"),
                   has_ref ? (const float *)p.deviceData<Real>("VolumetricMeasureRef") : nullptr,
                   has_vel ? (const sphb200_vec4_t *)p.deviceData<Vecd>("Velocity") : nullptr,
                   (sphb200_vec4_t *)p.deviceData<Vecd>("PosVol"),
                   has_ref ? (sphb200_vec4_t *)p.deviceData<Vecd>("PosVolRef") : nullptr,
                   has_vel ? p.deviceData<GatherRecord8>("PosVolVel") : nullptr, execution_instance().stream());
        posvol_dirty_ = false;
    }
    // slab-decomposed runs: slot of the first stored particle in the undecomposed run (SlabDecomposition::rebuild)
    uint32_t slotOrigin() const { return slot_origin_; }
    void setSlotOrigin(uint32_t o) { slot_origin_ = o; }
    bool isCellOrdered() const { return cell_ordered_; }
    void setCellOrdered(bool v) { cell_ordered_ = v; }
};
using RealBody = SPHBody;

class FluidBody : public SPHBody
{
  public:
    using SPHBody::SPHBody;
    // FluidBody water_block(sph_system, initial_water_block): the shape object is copied (dambreak.cpp:80-81)
    template <class ShapeType, typename = typename std::enable_if<std::is_base_of<ComplexShape, ShapeType>::value>::type>
    FluidBody(SPHSystem &system, const ShapeType &shape) : SPHBody(system, std::shared_ptr<ComplexShape>(new ShapeType(shape))) {}
    WeaklyCompressibleFluid &getFluid() { return dynamic_cast<WeaklyCompressibleFluid &>(getBaseMaterial()); }
};

class SolidBody : public SPHBody
{
  public:
    using SPHBody::SPHBody;
    // Solid::AverageVelocity/AverageAcceleration are registered on demand by Interaction<Wall>
    // (base_material.cpp:67-75, interaction_ck.hpp:79-91); a static wall leaves them unregistered (== 0).
    void registerWallVariables(const std::vector<Vecd> *normals = nullptr)
    {
        BaseParticles &p = getBaseParticles();
        auto *dv_n = p.registerStateVariable<Vecd>("NormalDirection");
        auto *dv_vol = p.getVariableByName<Real>("VolumetricMeasure");
        auto *dv_ref = p.registerStateVariable<Real>("VolumetricMeasureRef");
        size_t n = p.TotalRealParticles();
        ExecutionInstance &ex = execution_instance();
        ex.check(sphb200_copy_d2d(dv_ref->deviceAddress(), dv_vol->deviceAddress(), n * sizeof(Real), ex.stream()), "sphb200_copy_d2d");
        if (normals && n) p.upload(dv_n, normals->data());
    }
    // Slab-decomposed runs keep only the part of the static wall a rank can reach (dambreak_case.h, WallSlab): replace the
    // stored particle set by `pos` / `normals` with the global particle numbers `ids` (in-cell order and the CSR export go
    // by them). Volumes are uniform and were filled over the whole storage at registration. The caller rebuilds the cell
    // list (which brings the storage into cell order) afterwards.
    void loadWallSubset(const std::vector<Vecd> &pos, const std::vector<Vecd> &normals, const std::vector<UnsignedInt> &ids)
    {
        BaseParticles &p = getBaseParticles();
        const size_t n = pos.size();
        if (normals.size() != n || ids.size() != n) throw SphError("loadWallSubset: array sizes differ");
        p.resetContents(n); // throws if the storage is too small
        ExecutionInstance &ex = execution_instance();
        if (n)
        {
            p.upload(p.getVariableByName<Vecd>("Position"), pos.data());
            p.upload(p.getVariableByName<Vecd>("NormalDirection"), normals.data());
            ex.check(sphb200_copy_d2d(p.deviceData<Real>("VolumetricMeasureRef"), p.deviceData<Real>("VolumetricMeasure"), n * sizeof(Real), ex.stream()), "sphb200_copy_d2d");
            ex.check(sphb200_copy_h2d(p.referenceID(), ids.data(), n * sizeof(UnsignedInt), ex.stream()), "sphb200_copy_h2d");
            ex.check(sphb200_copy_h2d(p.deviceData<UnsignedInt>("OriginalID"), ids.data(), n * sizeof(UnsignedInt), ex.stream()), "sphb200_copy_h2d");
            ex.synchronize();
        }
        setCellOrdered(false);
        setPosVolDirty();
    }
    // NormalFromBodyShapeCK for arbitrary points of the body (host)
    std::vector<Vecd> normalsFromBodyShape(const std::vector<Vecd> &pos) const
    {
        std::vector<Vecd> normals(pos.size());
        for (size_t i = 0; i < pos.size(); ++i) normals[i] = shape_->directionToSurface(pos[i], sph_system_.dim_);
        return normals;
    }
    // NormalFromBodyShapeCK (host-side in the reference case file: StateDynamics<ParallelPolicy, ...>)
    void computeNormalFromBodyShape()
    {
        BaseParticles &p = getBaseParticles();
        auto *dv_pos = p.getVariableByName<Vecd>("Position");
        dv_pos->synchronizeWithDevice();
        size_t n = p.TotalRealParticles();
        std::vector<Vecd> normals(n);
        for (size_t i = 0; i < n; ++i) normals[i] = shape_->directionToSurface(dv_pos->Data()[i], sph_system_.dim_);
        registerWallVariables(&normals);
    }
};

// ObserverBody fluid_observer(sph_system, "FluidObserver"); fluid_observer.generateParticles<ObserverParticles>(points);
// ref: bodies/base_body.h (ObserverBody), particle_generator (ObserverParticles): probe particles at given positions
struct ObserverParticles {};
class ObserverBody : public SPHBody
{
  public:
    ObserverBody(SPHSystem &system, const std::string &name) : SPHBody(system, name) {}
    template <class ParticlesType> void generateParticles(const std::vector<Vecd> &positions)
    {
        static_assert(std::is_same<ParticlesType, ObserverParticles>::value, "ObserverBody holds ObserverParticles");
        generateParticlesFromPositions(positions, Real(0)); // observers carry no volume (they are never neighbours)
    }
};

inline std::shared_ptr<ComplexShape> makeSharedShape(ComplexShape *s) { return std::shared_ptr<ComplexShape>(s); }
template <class T, typename... Args> std::shared_ptr<T> makeShared(Args &&...args) { return std::make_shared<T>(std::forward<Args>(args)...); }
} // namespace SPH
#endif
