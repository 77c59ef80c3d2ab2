// sphinxsys_ck/base.h — data types, execution policies and the device execution instance of the C++ host layer.
//
// This header set keeps the spellings of the SPHinXsys "CK" API for the weakly-compressible SPH hot path so that a
// case file written against the reference compiles against it with `par_device` mapped to the B200 library:
// every `exec()` below ends in calls to the C ABI of libsphb200.so (include/sphb200.h). There is no host/TBB
// implementation and no SYCL: instantiating an algorithm with another policy is a compile-time error.
//
// Reference (paths relative to /root/reference/src/shared):
//   Real/Vecd/UnsignedInt ........ common/base_data_type.h:50-84,204-209
//   execution policies ........... shared_ck/.../execution_policy.h:37-83
//   ExecutionInstance ............ src_sycl/.../implementation_sycl.h:43-94 (one queue, global singleton)
//   error convention ............. common/sphinxsys_variable.h:252-257 (message to std::cout, then exit(1))
#ifndef SPHINXSYS_CK_BASE_H
#define SPHINXSYS_CK_BASE_H

#include <chrono>
#include <map>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../sphb200.h"

namespace SPH
{
using Real = float;           // base_data_type.h:50-53 (SPHINXSYS_USE_FLOAT, as the SYCL build sets)
using UnsignedInt = uint32_t; // base_data_type.h:58-60
constexpr Real TinyReal = 2.71051e-20f; // base_data_type.h:207
constexpr Real Pi = Real(3.14159265358979323846);

// Vecd: always three components; 2-D cases keep z == 0 (the device kernels are dimension-agnostic that way).
struct Vecd
{
    Real x = 0, y = 0, z = 0;
    Vecd() = default;
    Vecd(Real a, Real b, Real c = 0) : x(a), y(b), z(c) {}
    Real &operator[](int d) { return d == 0 ? x : (d == 1 ? y : z); }
    Real operator[](int d) const { return d == 0 ? x : (d == 1 ? y : z); }
    Vecd operator+(const Vecd &o) const { return Vecd(x + o.x, y + o.y, z + o.z); }
    Vecd operator-(const Vecd &o) const { return Vecd(x - o.x, y - o.y, z - o.z); }
    Vecd operator*(Real s) const { return Vecd(x * s, y * s, z * s); }
    static Vecd Zero() { return Vecd(); }
};
using Vec3d = Vecd;
template <class T> using StdVec = std::vector<T>; // base_data_type.h
using Vec2d = Vecd;
inline Vecd operator*(Real s, const Vecd &v) { return v * s; }

struct Matd // row-major 3x3
{
    Real m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    static Matd Identity() { return Matd(); }
};

struct BoundingBoxd
{
    Vecd lower_, upper_;
    BoundingBoxd() = default;
    BoundingBoxd(const Vecd &lo, const Vecd &up) : lower_(lo), upper_(up) {}
};

// ---- execution policies: only ParallelDevicePolicy is implemented ----
namespace execution
{
struct SequencedPolicy {};
struct ParallelPolicy {};
struct ParallelDevicePolicy {};
inline constexpr SequencedPolicy seq{};
inline constexpr ParallelPolicy par_host{};
inline constexpr ParallelDevicePolicy par_device{};
template <class Policy> constexpr void require_device_policy()
{
    static_assert(std::is_same<Policy, ParallelDevicePolicy>::value,
                  "sphinxsys_ck on libsphb200 implements execution::ParallelDevicePolicy only (no CPU fallback)");
}
} // namespace execution
using MainExecutionPolicy = execution::ParallelDevicePolicy;

// ---- errors: the reference prints and exits; here the message is thrown so harnesses can report it ----
class SphError : public std::runtime_error
{
  public:
    explicit SphError(const std::string &what) : std::runtime_error(what) {}
};

// ---- one library context + stream per process and device (ExecutionInstance of the SYCL build) ----
class ExecutionInstance
{
    sphb200_context_t *ctx_ = nullptr;
    int device_ = 0;
    void *stream_ = nullptr; // default stream: ordered with every other user of the default stream

    ExecutionInstance() = default;

  public:
    ~ExecutionInstance()
    {
        if (ctx_) sphb200_context_destroy(ctx_);
    }
    static ExecutionInstance &get()
    {
        static ExecutionInstance instance;
        return instance;
    }
    void setDevice(int device)
    {
        if (ctx_ && device != device_) throw SphError("ExecutionInstance: device already selected");
        device_ = device;
    }
    sphb200_context_t *ctx()
    {
        if (!ctx_)
        {
            int rc = sphb200_context_create(device_, &ctx_);
            if (rc != 0)
                throw SphError("sphb200_context_create(device=" + std::to_string(device_) + ") failed with code " +
                               std::to_string(rc) + ": a CUDA device is required (libsphb200 has no CPU path)");
        }
        return ctx_;
    }
    void *stream() const { return stream_; }
    void setStream(void *s) { stream_ = s; } // see StreamScope
    int device() const { return device_; }
    void check(int rc, const char *what)
    {
        if (rc == 0) return;
        std::string msg = std::string(what) + " failed: code " + std::to_string(rc) + ": " +
                          (ctx_ ? sphb200_last_error_string(ctx_) : "");
        std::cout << "\n Error: " << msg << std::endl; // reference style: report on std::cout ...
        throw SphError(msg);                           // ... and stop (exception instead of exit(1))
    }
    void synchronize() { check(sphb200_stream_sync(stream_), "sphb200_stream_sync"); }
    uint64_t launches() { return ctx_ ? sphb200_launch_count(ctx_) : 0; }
};
inline ExecutionInstance &execution_instance() { return ExecutionInstance::get(); }

// SPHB200_STEP_TRACE=1: wall time of every named stage of a case loop, the device drained before and after each one
// (a diagnosis mode: it serialises host and device, so the totals are not the run's speed). StepTrace::report() prints
// the table; the case classes wrap their stages in SPHCK_STAGE("name", statement).
class StepTrace
{
    std::map<std::string, std::pair<double, uint64_t>> ms_;
    std::vector<std::string> order_;

  public:
    static bool enabled()
    {
        static const bool on = [] {
            const char *e = std::getenv("SPHB200_STEP_TRACE");
            return e && e[0] != '0' && e[0] != 0;
        }();
        return on;
    }
    static StepTrace &get()
    {
        static StepTrace t;
        return t;
    }
    void add(const std::string &name, double ms)
    {
        auto it = ms_.find(name);
        if (it == ms_.end())
        {
            order_.push_back(name);
            ms_[name] = {ms, 1};
        }
        else
        {
            it->second.first += ms;
            it->second.second++;
        }
    }
    void report(std::ostream &out, uint64_t steps)
    {
        double total = 0;
        for (auto &n : order_) total += ms_[n].first;
        out << "SPHB200_STEP_TRACE over " << steps << " advection steps (device drained around every stage):\n";
        for (auto &n : order_)
            out << "  " << n << ": " << ms_[n].first / double(steps ? steps : 1) << " ms per step in " << double(ms_[n].second) / double(steps ? steps : 1)
                << " calls (" << 100.0 * ms_[n].first / (total > 0 ? total : 1) << " %)\n";
        out << "  total: " << total / double(steps ? steps : 1) << " ms per step" << std::endl;
    }
    void clear()
    {
        ms_.clear();
        order_.clear();
    }
};
#define SPHCK_STAGE(name, statement)                                                                              \
    do                                                                                                            \
    {                                                                                                             \
        if (::SPH::StepTrace::enabled())                                                                          \
        {                                                                                                         \
            ::SPH::execution_instance().synchronize();                                                            \
            auto t0__ = std::chrono::steady_clock::now();                                                         \
            statement;                                                                                            \
            ::SPH::execution_instance().synchronize();                                                            \
            ::SPH::StepTrace::get().add(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0__).count()); \
        }                                                                                                         \
        else                                                                                                      \
        {                                                                                                         \
            statement;                                                                                            \
        }                                                                                                         \
    } while (0)

// Every library call made while the scope lives is issued on `stream` instead of the default stream (ordering against
// the default stream is then the caller's business: events).
class StreamScope
{
    void *saved_;

  public:
    explicit StreamScope(void *stream) : saved_(execution_instance().stream()) { execution_instance().setStream(stream); }
    ~StreamScope() { execution_instance().setStream(saved_); }
    StreamScope(const StreamScope &) = delete;
    StreamScope &operator=(const StreamScope &) = delete;
};

#define SPHCK_CALL(fn, ...) ::SPH::execution_instance().check(fn(::SPH::execution_instance().ctx(), __VA_ARGS__), #fn)

// BaseDynamics<ReturnType>: base_dynamics.h (exec(dt) interface)
template <class ReturnType = void> class BaseDynamics
{
  public:
    virtual ~BaseDynamics() {}
    virtual ReturnType exec(Real dt = 0.0) = 0;
};
} // namespace SPH
#endif
