// =====================================================================================
//  sph_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
//  A plain C++17/OpenMP restatement of the SPHinXsys weakly-compressible SPH fluid hot
//  path (neighbour machinery + density summation + acoustic 1st/2nd half with wall,
//  Riemann), written from the algorithm description of the reference sources cited on
//  each function ("ref:" comments; paths relative to /root/reference/src/shared).
//
//  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
//  legs may load this library. The product path (libsphb200.so) never links or calls it.
//
//  The reference itself cannot be compiled in this image (needs Eigen, oneTBB, Boost,
//  Simbody, spdlog, DPC++), so this restatement is pinned instead by the reference's own
//  known-answer vectors and regression series (see tests/test_oracle_*.py).
//
//  Canonical conventions:
//    * compiled with -ffp-contract=off: every fp op in index/criterion expressions is
//      rounded separately;
//    * cell lists hold particles in ascending index order inside a cell; CSR neighbour
//      rows follow the reference search order (cells x -> y -> z, then in-cell order),
//      so results are deterministic (the reference's atomics make its order arbitrary);
//    * 2-D runs through the same code with z == 0 and one cell layer in z (bit-identical
//      to the 2-D formulas: the extra terms are exact zeros).
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t u32;

namespace
{
// ------------------------------------------------------------------------------------
// small helpers; ref: common/scalar_functions.h:40-43,70-73,128-131
// ------------------------------------------------------------------------------------
template <class R> inline R SMAX(R a, R b) { return a >= b ? a : b; }
template <class R> inline R SMIN(R a, R b) { return a <= b ? a : b; }
template <class R> inline R SGN(R x) { return x < R(0) ? R(-1) : (x > R(0) ? R(1) : R(0)); }

template <class R> struct V3
{
    R x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(R a, R b, R c) : x(a), y(b), z(c) {}
    V3 operator+(const V3 &o) const { return V3(x + o.x, y + o.y, z + o.z); }
    V3 operator-(const V3 &o) const { return V3(x - o.x, y - o.y, z - o.z); }
    V3 operator-() const { return V3(-x, -y, -z); }
    V3 operator*(R s) const { return V3(x * s, y * s, z * s); }
    V3 operator/(R s) const { return V3(x / s, y / s, z / s); }
    V3 &operator+=(const V3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    V3 &operator-=(const V3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    R dot(const V3 &o) const { return (x * o.x + y * o.y) + z * o.z; }
    R squaredNorm() const { return (x * x + y * y) + z * z; }
    R norm() const { return std::sqrt(squaredNorm()); }
    // Eigen normalized(): z = squaredNorm(); z > 0 ? v / sqrt(z) : v
    V3 normalized() const { R s = squaredNorm(); return s > R(0) ? (*this) / std::sqrt(s) : *this; }
};
template <class R> inline V3<R> operator*(R s, const V3<R> &v) { return v * s; }

template <class R> struct M3 // row-major 3x3
{
    R m[9];
    static M3 Zero() { M3 a; for (R &v : a.m) v = 0; return a; }
    static M3 Identity() { M3 a = Zero(); a.m[0] = a.m[4] = a.m[8] = 1; return a; }
    M3 operator*(R s) const { M3 a; for (int k = 0; k < 9; ++k) a.m[k] = m[k] * s; return a; }
    M3 operator+(const M3 &o) const { M3 a; for (int k = 0; k < 9; ++k) a.m[k] = m[k] + o.m[k]; return a; }
    V3<R> operator*(const V3<R> &v) const
    {
        return V3<R>((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[3] * v.x + m[4] * v.y) + m[5] * v.z,
                     (m[6] * v.x + m[7] * v.y) + m[8] * v.z);
    }
    M3 operator*(const M3 &o) const
    {
        M3 a;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                a.m[3 * r + c] = (m[3 * r] * o.m[c] + m[3 * r + 1] * o.m[3 + c]) + m[3 * r + 2] * o.m[6 + c];
        return a;
    }
    M3 transpose() const
    {
        M3 a;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) a.m[3 * r + c] = m[3 * c + r];
        return a;
    }
    R determinant() const
    {
        return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
               m[2] * (m[3] * m[7] - m[4] * m[6]);
    }
    M3 inverse() const
    {
        R d = determinant();
        M3 a;
        a.m[0] = (m[4] * m[8] - m[5] * m[7]) / d;
        a.m[1] = (m[2] * m[7] - m[1] * m[8]) / d;
        a.m[2] = (m[1] * m[5] - m[2] * m[4]) / d;
        a.m[3] = (m[5] * m[6] - m[3] * m[8]) / d;
        a.m[4] = (m[0] * m[8] - m[2] * m[6]) / d;
        a.m[5] = (m[2] * m[3] - m[0] * m[5]) / d;
        a.m[6] = (m[3] * m[7] - m[4] * m[6]) / d;
        a.m[7] = (m[1] * m[6] - m[0] * m[7]) / d;
        a.m[8] = (m[0] * m[4] - m[1] * m[3]) / d;
        return a;
    }
};

// ------------------------------------------------------------------------------------
// Mesh; ref: meshes/base_mesh.cpp:6-16 (geometry is computed by the caller in Real
// precision and handed in), meshes/base_mesh.hxx:9-26,73-99
// ------------------------------------------------------------------------------------
struct MeshPOD
{
    double lower[3];
    double spacing;
    int cells[3];
};

template <class R> struct Mesh
{
    R lower[3];
    R spacing;
    int cells[3];
    u32 total() const { return u32(cells[0]) * u32(cells[1]) * u32(cells[2]); }
    void set(const MeshPOD &p)
    {
        for (int d = 0; d < 3; ++d) { lower[d] = R(p.lower[d]); cells[d] = p.cells[d]; }
        spacing = R(p.spacing);
    }
    // ref: base_mesh.hxx:9-15 — floor((x - lower) / spacing), clamp to [0, all_grid_points - 2]
    inline void cellIndex(const R *x, int *c) const
    {
        for (int d = 0; d < 3; ++d)
        {
            R t = x[d] - lower[d];
            R u = t / spacing;
            int k = (int)std::floor(u);
            k = k < 0 ? 0 : k;
            k = k > cells[d] - 1 ? cells[d] - 1 : k;
            c[d] = k;
        }
    }
    // ref: base_mesh.hxx:73-78 (z fastest)
    inline u32 linear(const int *c) const
    {
        return u32(c[0]) * u32(cells[1]) * u32(cells[2]) + u32(c[1]) * u32(cells[2]) + u32(c[2]);
    }
};

// ref: base_mesh.hxx:85-99
inline u32 mortonSpread(u32 i)
{
    u32 x = i;
    x &= 0x3ff;
    x = (x | x << 16) & 0x30000ff;
    x = (x | x << 8) & 0x300f00f;
    x = (x | x << 4) & 0x30c30c3;
    x = (x | x << 2) & 0x9249249;
    return x;
}
inline u32 mortonKey(const int *c) { return mortonSpread(c[0]) | (mortonSpread(c[1]) << 1) | (mortonSpread(c[2]) << 2); }

// ------------------------------------------------------------------------------------
// Tabulated kernel; ref: shared_ck/smoothing_kernel/kernel_tabulated_ck.h:37-72, .cpp:6-25,
// shared_ck/body_relation/neighbor_method.hpp:18-142. Table values are produced by the
// caller from the analytic W_1D/dW_1D (kernels/kernel_wendland_c2.cpp:8-50,
// kernels/kernel_laguerre_gauss.cpp:8-50) in Real precision.
// ------------------------------------------------------------------------------------
struct KernelPOD
{
    int dim;          // 2 or 3
    int kind;         // 0 Wendland C2, 1 Laguerre-Gauss (only used by the analytic legacy form)
    double h;         // reference smoothing length
    double kernel_size; // 2.0
    double dimension_factor; // factor_W_dim * h^dim  (base_kernel.h:87-93)
    double w[24], dw[24];
};

template <class R> struct Kernel
{
    int dim, kind;
    R h, inv_h, kernel_size, kernel_size_sq;
    R dimension_factor;
    R dq, delta0, delta1, delta2, delta3;
    R w[24], dw[24];
    R inv_h_pow_dim, inv_h_pow_dim1; // inv_h^dim and inv_h^(dim+1)
    // legacy analytic factors (base_kernel.cpp:21-29)
    R factor_W, factor_dW, rc_ref_sqr;

    void set(const KernelPOD &p)
    {
        dim = p.dim; kind = p.kind;
        h = R(p.h);
        inv_h = R(1.0) / h;
        kernel_size = R(p.kernel_size);
        kernel_size_sq = kernel_size * kernel_size;
        dimension_factor = R(p.dimension_factor);
        dq = kernel_size / R(20);
        delta0 = (R(-1.0) * dq) * (R(-2.0) * dq) * (R(-3.0) * dq);
        delta1 = dq * (R(-1.0) * dq) * (R(-2.0) * dq);
        delta2 = (R(2.0) * dq) * dq * (R(-1.0) * dq);
        delta3 = (R(3.0) * dq) * (R(2.0) * dq) * dq;
        for (int k = 0; k < 24; ++k) { w[k] = R(p.w[k]); dw[k] = R(p.dw[k]); }
        R ih2 = inv_h * inv_h, ih3 = ih2 * inv_h, ih4 = ih3 * inv_h;
        inv_h_pow_dim = dim == 2 ? ih2 : ih3;
        inv_h_pow_dim1 = dim == 2 ? ih3 : ih4;
        // legacy: factor_W_dim = inv_h^dim * sigma_dim ; sigma_dim = dimension_factor (up to rounding)
        factor_W = inv_h_pow_dim * dimension_factor;
        factor_dW = inv_h * factor_W;
        R rc = kernel_size * h;
        rc_ref_sqr = rc * rc;
    }
    // ref: kernel_tabulated_ck.h:45-58
    inline R interpolateCubic(const R *data, R q) const
    {
        int location = (int)std::floor(q / dq);
        int i = location + 1;
        R f1 = q - R(location) * dq;
        R f0 = f1 + dq;
        R f2 = f1 - dq;
        R f3 = f1 - 2 * dq;
        return (f1 * f2 * f3) / delta0 * data[i - 1] + (f0 * f2 * f3) / delta1 * data[i] +
               (f0 * f1 * f3) / delta2 * data[i + 1] + (f0 * f1 * f2) / delta3 * data[i + 2];
    }
    // ref: neighbor_method.hpp:24-52,103-130
    inline R W(const V3<R> &disp) const { return inv_h_pow_dim * dimension_factor * interpolateCubic(w, disp.norm() * inv_h); }
    inline R dW(const V3<R> &disp) const { return inv_h_pow_dim1 * dimension_factor * interpolateCubic(dw, disp.norm() * inv_h); }
    inline R W0() const { return inv_h_pow_dim * dimension_factor * interpolateCubic(w, R(0)); }
    // ref: neighbor_method.hpp:152-156
    inline bool criterion(const R *xi, const R *xj) const
    {
        R sx = inv_h * (xi[0] - xj[0]);
        R sy = inv_h * (xi[1] - xj[1]);
        R sz = inv_h * (xi[2] - xj[2]);
        R r2 = (sx * sx + sy * sy) + sz * sz;
        return r2 < kernel_size_sq;
    }
    // legacy analytic forms; ref: kernels/base_kernel.cpp:44-61, kernel_wendland_c2.cpp, kernel_laguerre_gauss.cpp
    inline R W1D(R q) const
    {
        if (kind == 0) return R(std::pow(1.0 - 0.5 * q, 4) * (1.0 + 2.0 * q));
        return R((1.0 - std::pow(double(q), 2) + std::pow(double(q), 4) / 6.0) * std::exp(-std::pow(double(q), 2)));
    }
    inline R dW1D(R q) const
    {
        if (kind == 0) return R(0.625 * std::pow(q - 2.0, 3) * q);
        double Q = q;
        return R((-std::pow(Q, 5) / 3.0 + 8.0 * std::pow(Q, 3) / 3.0 - 4.0 * Q) * std::exp(-Q * Q));
    }
    inline R W_analytic(R r) const { return factor_W * W1D(r * inv_h); }
    inline R dW_analytic(R r) const { return factor_dW * dW1D(r * inv_h); }
};

// ------------------------------------------------------------------------------------
// Primitives; ref: common/algorithm_primitive.h:244-250 (scan)
// ------------------------------------------------------------------------------------
u32 exclusiveScan(const u32 *in, u32 *out, u32 n)
{
    // out[0] = 0; out[i] = sum_{k<i} in[k]; the last input entry is unused. returns out[n-1].
    u32 acc = 0;
    for (u32 i = 0; i < n; ++i)
    {
        u32 v = in[i];
        out[i] = acc;
        acc += v;
    }
    return n ? out[n - 1] : 0;
}

struct CSR
{
    std::vector<u32> offset, index;
    std::vector<unsigned char> shift; // periodic runs: image code of every entry (13 = the particle itself), else empty
};

// Periodic runs follow PeriodicConditionUsingCellLinkedList (particle_dynamics/general_dynamics/domian_bouding/
// domain_bounding.cpp:18-65): besides the particles, the cell-linked list holds ghost ENTRIES (source index,
// translated position); `ext_*` is that entry table (real particles first), `particle_index` then holds entry ids.
template <class R> struct CellList
{
    Mesh<R> mesh;
    std::vector<u32> cell_offset, particle_index;
    std::vector<u32> ext_index;           // entry -> source particle
    std::vector<R> ext_pos;               // entry -> (translated) position
    std::vector<unsigned char> ext_shift; // entry -> image code (sx+1) + 3 (sy+1) + 9 (sz+1), s in {-1, 0, +1}
};

// ref: shared_ck/.../update_cell_linked_list.hpp:40-106 (count -> scan -> fill)
template <class R> void buildCellList(CellList<R> &cl, const R *pos, u32 n)
{
    u32 cells = cl.mesh.total();
    std::vector<u32> count(cells + 1, 0), cell_of(n);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i)
    {
        int c[3];
        cl.mesh.cellIndex(pos + 3 * i, c);
        cell_of[i] = cl.mesh.linear(c);
    }
    for (u32 i = 0; i < n; ++i) count[cell_of[i]]++;
    cl.cell_offset.assign(cells + 1, 0);
    exclusiveScan(count.data(), cl.cell_offset.data(), cells + 1);
    cl.particle_index.assign(std::max<u32>(n, 1), 0);
    std::vector<u32> cursor(cl.cell_offset.begin(), cl.cell_offset.end() - 1);
    for (u32 i = 0; i < n; ++i) cl.particle_index[cursor[cell_of[i]]++] = i; // ascending i inside a cell
}

// ref: meshes/cell_linked_list.hpp:120-160 + for_3D_build/meshes/mesh_iterators.hpp:18-27
template <class R, class F> inline void searchBox(const CellList<R> &cl, const R *x, int depth, const F &f)
{
    int c[3];
    cl.mesh.cellIndex(x, c);
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d)
    {
        lo[d] = std::max(0, c[d] - depth);
        hi[d] = std::min(cl.mesh.cells[d], c[d] + depth + 1);
    }
    int cc[3];
    for (cc[0] = lo[0]; cc[0] < hi[0]; ++cc[0])
        for (cc[1] = lo[1]; cc[1] < hi[1]; ++cc[1])
            for (cc[2] = lo[2]; cc[2] < hi[2]; ++cc[2])
            {
                u32 lin = cl.mesh.linear(cc);
                for (u32 n = cl.cell_offset[lin]; n < cl.cell_offset[lin + 1]; ++n) f(cl.particle_index[n]);
            }
}

// ref: shared_ck/.../update_body_relation.hpp:62-164 — the reference builds the symmetric
// inner list by a one-sided (i<j) search with atomics; the resulting neighbour SET of i is
// { j != i : criterion(i,j) } over the 3^d cell box, which is what is enumerated here.
template <class R, class Crit>
void buildInner(CSR &csr, const CellList<R> &cl, const R *pos, u32 n, const Crit &crit)
{
    std::vector<u32> count(n + 1, 0);
    const bool ext = !cl.ext_index.empty(); // periodic: the list holds entries (source, translated position)
    const R *epos = ext ? cl.ext_pos.data() : pos;
    // an entry is a neighbour unless it is the particle itself (a particle's own images count only in boxes narrower
    // than two cut-off radii, which the periodic conditions do not support)
    auto hit = [&](long i, u32 e) { return e != (u32)i && crit(pos + 3 * i, epos + 3 * e); };
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; ++i)
    {
        u32 c = 0;
        searchBox(cl, pos + 3 * i, 1, [&](u32 e) { if (hit(i, e)) ++c; });
        count[i] = c;
    }
    csr.offset.assign(n + 1, 0);
    exclusiveScan(count.data(), csr.offset.data(), n + 1);
    csr.index.assign(std::max<u32>(csr.offset[n], 1), 0);
    csr.shift.clear();
    if (ext) csr.shift.assign(csr.index.size(), 13);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; ++i)
    {
        u32 k = csr.offset[i];
        searchBox(cl, pos + 3 * i, 1, [&](u32 e) {
            if (!hit(i, e)) return;
            if (ext) { csr.index[k] = cl.ext_index[e]; csr.shift[k] = cl.ext_shift[e]; }
            else csr.index[k] = e;
            ++k;
        });
    }
}

// ref: update_body_relation.hpp:199-288 (full search, deterministic order)
template <class R, class Crit>
void buildContact(CSR &csr, const CellList<R> &tar_cl, const R *src_pos, u32 n_src, const R *tar_pos, int depth,
                  const Crit &crit)
{
    std::vector<u32> count(n_src + 1, 0);
    if (tar_cl.cell_offset.empty()) // no contact body (or its cell list was never built): empty relation
    {
        csr.offset.assign(n_src + 1, 0);
        csr.index.assign(1, 0);
        return;
    }
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n_src; ++i)
    {
        u32 c = 0;
        searchBox(tar_cl, src_pos + 3 * i, depth, [&](u32 j) { if (crit(src_pos + 3 * i, tar_pos + 3 * j)) ++c; });
        count[i] = c;
    }
    csr.offset.assign(n_src + 1, 0);
    exclusiveScan(count.data(), csr.offset.data(), n_src + 1);
    csr.index.assign(std::max<u32>(csr.offset[n_src], 1), 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n_src; ++i)
    {
        u32 k = csr.offset[i];
        searchBox(tar_cl, src_pos + 3 * i, depth,
                  [&](u32 j) { if (crit(src_pos + 3 * i, tar_pos + 3 * j)) csr.index[k++] = j; });
    }
}

// ------------------------------------------------------------------------------------
// Parameters handed in by the caller
// ------------------------------------------------------------------------------------
struct ParamsPOD
{
    int dim;                // 2 | 3
    int riemann;            // 0 NoRiemann, 1 Acoustic (TruncatedLinear 3.0), 2 Dissipative (NoLimiter)
    int correction;         // 0 none, 1 LinearCorrection (B matrix)
    int free_surface;       // density regularisation: 1 FreeSurface (max(sigma,1)), 0 Internal
    double rho0, c0;
    double gravity[3];
    double U_ref;
    double h_min;           // = h_ref
    double acoustic_cfl;    // 0.6
    double advection_cfl;   // 0.25
    double correction_alpha; // 0.5
    double sigma0;          // legacy: lattice number density (adaptation.cpp:26-60)
    double wall_rho0;       // legacy: Solid reference density (1.0)
    int contact_depth;      // search depth in the wall mesh (cell_linked_list.hpp:161-167)
    int threads;            // 0 = leave OpenMP default
    int periodic_axes;      // bit d: periodic along axis d (PeriodicAlongAxis, domain_bounding.h:48-66)
    double periodic_lower[3], periodic_upper[3]; // bounding_bounds_ (already rounded to Real by the caller)
    double periodic_cutoff; // cut_off_radius_max_
    int surface_indicator;  // 1: FreeSurfaceIndicationCK runs in the case loop (dambreak.cpp:133-134,192)
    double viscosity;       // mu of the Viscosity closure; > 0: ViscousForceCK runs in the case loop
    int transport_velocity; // 1: KernelGradientIntegral + TransportVelocityCorrectionCK<TruncatedLinear> run in the case loop
    double transport_coefficient; // 0.2 (transport_velocity_correction_ck.h:19)
};

template <class R> struct Body
{
    u32 n = 0;
    std::map<std::string, std::vector<R>> real;   // scalars (n), vectors (3n), matrices (9n)
    std::map<std::string, std::vector<u32>> uint; // ids
    std::vector<R> &r(const std::string &k, size_t width = 1, R init = R(0))
    {
        auto it = real.find(k);
        if (it == real.end()) it = real.emplace(k, std::vector<R>(size_t(n) * width, init)).first;
        return it->second;
    }
};

template <class R> struct Sim
{
    ParamsPOD P;
    Kernel<R> K;
    Body<R> fluid, wall, observer;
    CellList<R> fluid_cl, wall_cl;
    CSR inner, contact, observer_contact;
    std::vector<std::vector<double>> probe_series; // one row per recorded step: interpolated pressure at every probe
    // legacy per-pair storage (AoS Neighborhood restated in CSR order; neighborhood.h:49-66)
    std::vector<R> in_W, in_dW, in_r, in_e, ct_W, ct_dW, ct_r, ct_e;
    // riemann constants; ref: riemann_solver_ck.hpp:58-69
    R rho0, c0, p0, Z, inv_Z_sum, inv_Z_ave, Z_geo, inv_c_ave, limiter;
    double physical_time = 0;
    long acoustic_steps = 0, outer_steps = 0;
    std::vector<double> energy_series, time_series;

    void init(const ParamsPOD &p, const KernelPOD &k, const MeshPOD &fm, const MeshPOD &wm)
    {
        P = p;
        K.set(k);
        fluid_cl.mesh.set(fm);
        wall_cl.mesh.set(wm);
        rho0 = R(p.rho0); c0 = R(p.c0);
        p0 = rho0 * c0 * c0; // weakly_compressible_fluid.cpp:8-9
        Z = rho0 * c0;
        inv_Z_sum = R(1.0) / (Z + Z);
        inv_Z_ave = (Z + Z) / (Z * Z + Z * Z);
        Z_geo = R(2.0) * Z * Z * inv_Z_sum;
        inv_c_ave = R(0.5) * (rho0 + rho0) * inv_Z_ave;
        limiter = R(3.0);
#ifdef _OPENMP
        if (p.threads > 0) omp_set_num_threads(p.threads);
#endif
    }

    // ---- riemann; ref: riemann_solver_ck.hpp:17-56, common/common_functors.h:82-94 ----
    inline R AverageP(R p_i, R p_j) const { return inv_Z_sum * (p_i * Z + p_j * Z); }
    inline V3<R> AverageV(const V3<R> &vi, const V3<R> &vj) const { return (vi * Z + vj * Z) * inv_Z_sum; }
    inline R PJump(R u) const
    {
        if (P.riemann == 0) return R(0);
        R lim = P.riemann == 1 ? SMIN(limiter * (inv_c_ave * SMAX(u, R(0))), R(1)) : R(1);
        return Z_geo * u * lim;
    }
    inline R UJump(R dp) const { return P.riemann == 0 ? R(0) : dp * inv_Z_ave; }

    inline V3<R> vec(const std::vector<R> &a, u32 i) const { return V3<R>(a[3 * i], a[3 * i + 1], a[3 * i + 2]); }
    inline void setv(std::vector<R> &a, u32 i, const V3<R> &v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }
    inline M3<R> mat(const std::vector<R> &a, u32 i) const { M3<R> m; std::memcpy(m.m, &a[9 * i], 9 * sizeof(R)); return m; }

    void ensureFluidState()
    {
        u32 n = fluid.n;
        fluid.r("Position", 3); fluid.r("Velocity", 3); fluid.r("Displacement", 3);
        fluid.r("Force", 3); fluid.r("ForcePrior", 3); fluid.r("PreviousGravityForceCK", 3);
        fluid.r("VolumetricMeasure"); fluid.r("Mass"); fluid.r("Density", 1, rho0); fluid.r("Pressure");
        fluid.r("Compression", 1, R(1)); fluid.r("CompressionRate"); fluid.r("VolumetricMeasureRef");
        fluid.r("CompressionSummation", 1, R(1)); fluid.r("DensityChangeRate"); fluid.r("DensitySummation");
        if (fluid.real.find("LinearCorrectionMatrix") == fluid.real.end())
        {
            auto &B = fluid.r("LinearCorrectionMatrix", 9);
            for (u32 i = 0; i < n; ++i) B[9 * i] = B[9 * i + 4] = B[9 * i + 8] = R(1);
        }
        if (fluid.uint.find("OriginalID") == fluid.uint.end())
        {
            auto &o = fluid.uint["OriginalID"]; o.resize(n); std::iota(o.begin(), o.end(), 0u);
            auto &s = fluid.uint["SortedID"]; s.resize(n); std::iota(s.begin(), s.end(), 0u);
        }
        wall.r("Position", 3); wall.r("VolumetricMeasure"); wall.r("VolumetricMeasureRef"); wall.r("Mass");
        wall.r("Velocity", 3); wall.r("Acceleration", 3); wall.r("NormalDirection", 3);
    }

    // =================================================================================
    // configuration dynamics
    // =================================================================================
    // PeriodicBounding::checkLowerBound/checkUpperBound, axis by axis; ref: domain_bounding.h:98-108
    void periodicBounding()
    {
        std::vector<R> &pos = fluid.r("Position", 3);
        for (int a = 0; a < 3; ++a)
        {
            if (!(P.periodic_axes >> a & 1)) continue;
            const R lo = R(P.periodic_lower[a]), up = R(P.periodic_upper[a]), L = up - lo;
            for (u32 i = 0; i < fluid.n; ++i)
            {
                R &x = pos[3 * i + a];
                if (x < lo) x += L;
                else if (x > up) x -= L;
            }
        }
    }
    // UpdateCellLinkedList, then for periodic runs PeriodicCellLinkedList::exec of every periodic axis in turn:
    // every list entry (ghost entries of the earlier axes included) with lower < x < lower + cutoff gets a ghost
    // entry at x + L, every one with upper - cutoff < x < upper one at x - L; ref: domain_bounding.cpp:18-65
    void cellListFluid()
    {
        const std::vector<R> &pos = fluid.r("Position", 3);
        if (!P.periodic_axes)
        {
            fluid_cl.ext_index.clear(); fluid_cl.ext_pos.clear(); fluid_cl.ext_shift.clear();
            buildCellList(fluid_cl, pos.data(), fluid.n);
            return;
        }
        std::vector<u32> &ei = fluid_cl.ext_index;
        std::vector<R> &ep = fluid_cl.ext_pos;
        std::vector<unsigned char> &es = fluid_cl.ext_shift;
        ei.resize(fluid.n); std::iota(ei.begin(), ei.end(), 0u);
        ep.assign(pos.begin(), pos.begin() + size_t(3) * fluid.n);
        es.assign(fluid.n, 13);
        const int weight[3] = {1, 3, 9};
        for (int a = 0; a < 3; ++a)
        {
            if (!(P.periodic_axes >> a & 1)) continue;
            const R lo = R(P.periodic_lower[a]), up = R(P.periodic_upper[a]), L = up - lo, rc = R(P.periodic_cutoff);
            const size_t m = ei.size();
            for (size_t e = 0; e < m; ++e)
            {
                const R x = ep[3 * e + a];
                for (int side = 0; side < 2; ++side)
                {
                    const bool near = side == 0 ? (x > lo && x < lo + rc) : (x < up && x > up - rc);
                    if (!near) continue;
                    ei.push_back(ei[e]);
                    for (int d = 0; d < 3; ++d) ep.push_back(ep[3 * e + d]);
                    ep[ep.size() - 3 + a] = side == 0 ? x + L : x - L;
                    es.push_back((unsigned char)(es[e] + (side == 0 ? weight[a] : -weight[a])));
                }
            }
        }
        buildCellList(fluid_cl, ep.data(), (u32)ei.size());
    }
    // x_i - x_j of entry n of the inner relation; for a periodic image x_j is the TRANSLATED position, rounded as the
    // reference's ghost list entry is (particle_position +/- periodic_translation_, domain_bounding.cpp:26,45)
    inline V3<R> innerDisp(const std::vector<R> &pos, u32 i, u32 n) const
    {
        V3<R> xj = vec(pos, inner.index[n]);
        if (!inner.shift.empty() && inner.shift[n] != 13)
        {
            int code = inner.shift[n];
            const int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
            R *c[3] = {&xj.x, &xj.y, &xj.z};
            for (int a = 0; a < 3; ++a)
            {
                const R L = R(P.periodic_upper[a]) - R(P.periodic_lower[a]);
                if (s[a] > 0) *c[a] = *c[a] + L;
                else if (s[a] < 0) *c[a] = *c[a] - L;
            }
        }
        return vec(pos, i) - xj;
    }
    void cellListWall() { buildCellList(wall_cl, wall.r("Position", 3).data(), wall.n); }

    void relationsCK()
    {
        const Kernel<R> &k = K;
        auto crit = [&k](const R *a, const R *b) { return k.criterion(a, b); };
        buildInner(inner, fluid_cl, fluid.r("Position", 3).data(), fluid.n, crit);
        buildContact(contact, wall_cl, fluid.r("Position", 3).data(), fluid.n, wall.r("Position", 3).data(),
                     P.contact_depth, crit);
    }
    // legacy criterion: displacement.squaredNorm() < rc_ref_sqr ; ref: kernels/base_kernel.h:105-114,
    // particle_neighborhood/neighborhood.cpp:26-35,84-99,147-160
    void relationsLegacy()
    {
        const R rc2 = K.rc_ref_sqr;
        auto crit = [rc2](const R *a, const R *b) {
            R dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
            return (dx * dx + dy * dy) + dz * dz < rc2;
        };
        const std::vector<R> &pos = fluid.r("Position", 3), &wpos = wall.r("Position", 3);
        buildInner(inner, fluid_cl, pos.data(), fluid.n, crit);
        buildContact(contact, wall_cl, pos.data(), fluid.n, wpos.data(), P.contact_depth, crit);
        auto fill = [&](const CSR &csr, const std::vector<R> &tpos, std::vector<R> &W, std::vector<R> &dW,
                        std::vector<R> &rr, std::vector<R> &e) {
            size_t m = csr.offset[fluid.n];
            W.resize(m); dW.resize(m); rr.resize(m); e.resize(3 * m);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)fluid.n; ++i)
                for (u32 n = csr.offset[i]; n < csr.offset[i + 1]; ++n)
                {
                    u32 j = csr.index[n];
                    V3<R> d = vec(pos, i) - vec(tpos, j);
                    R dist = std::sqrt(d.squaredNorm());
                    W[n] = K.W_analytic(dist);
                    dW[n] = K.dW_analytic(dist);
                    rr[n] = dist;
                    V3<R> ee = d / (dist + R(2.71051e-20)); // base_kernel.h:99-100
                    e[3 * n] = ee.x; e[3 * n + 1] = ee.y; e[3 * n + 2] = ee.z;
                }
        };
        fill(inner, pos, in_W, in_dW, in_r, in_e);
        fill(contact, wpos, ct_W, ct_dW, ct_r, ct_e);
    }

    // ref: shared_ck/.../particle_sort_ck.hpp:61-104, base_configuration_dynamics.h:101-125.
    // The reference host sort is an unstable quicksort; any permutation with sorted keys is valid.
    // The oracle uses a stable sort (ties keep ascending index), as the device radix sort does.
    // `with_force`: the reference does NOT list "Force" among the CK evolving variables
    // (acoustic_step_1st_half.hpp:30-34) so it is left unpermuted unless asked.
    void sortParticles(bool legacy)
    {
        u32 n = fluid.n;
        std::vector<u32> keys(n), perm(n);
        const std::vector<R> &pos = fluid.r("Position", 3);
        for (u32 i = 0; i < n; ++i)
        {
            int c[3];
            fluid_cl.mesh.cellIndex(&pos[3 * i], c);
            keys[i] = mortonKey(c);
            perm[i] = i;
        }
        std::stable_sort(perm.begin(), perm.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });
        fluid.uint["SortKeys"] = keys;
        fluid.uint["Permutation"] = perm;
        std::vector<std::string> names;
        if (legacy)
            names = {"Position", "VolumetricMeasure", "Velocity", "Mass", "ForcePrior", "Force", "DensityChangeRate",
                     "Density", "Pressure"};
        else
            names = {"Position", "VolumetricMeasure", "Velocity", "Mass", "ForcePrior", "Compression", "CompressionRate",
                     "VolumetricMeasureRef", "PreviousGravityForceCK"};
        // ForcePriorCK registers the previous value of its force as evolving (force_prior_ck.cpp:13-14)
        if (!legacy && fluid.real.count("PreviousViscousForce")) names.push_back("PreviousViscousForce");
        for (const std::string &nm : names)
        {
            std::vector<R> &a = fluid.real[nm];
            size_t w = a.size() / n;
            std::vector<R> tmp(a);
            for (u32 i = 0; i < n; ++i)
                for (size_t e = 0; e < w; ++e) a[w * i + e] = tmp[w * perm[i] + e];
        }
        if (!legacy && fluid.uint.count("PreviousSurfaceIndicator")) // evolving: surface_indication_ck.hpp:43
        {
            std::vector<u32> &a = fluid.uint["PreviousSurfaceIndicator"];
            std::vector<u32> t2(a);
            for (u32 i = 0; i < n; ++i) a[i] = t2[perm[i]];
        }
        std::vector<u32> &oid = fluid.uint["OriginalID"], &sid = fluid.uint["SortedID"];
        std::vector<u32> tmp(oid);
        for (u32 i = 0; i < n; ++i) oid[i] = tmp[perm[i]];
        for (u32 i = 0; i < n; ++i) sid[oid[i]] = i;
    }

    // =================================================================================
    // CK fluid dynamics
    // =================================================================================
    // ref: general_dynamics/force_prior_ck.hpp:38-44, force_prior_ck.h:53-57
    void gravityForce()
    {
        std::vector<R> &fp = fluid.r("ForcePrior", 3), &prev = fluid.r("PreviousGravityForceCK", 3);
        const std::vector<R> &m = fluid.r("Mass");
        V3<R> g(R(P.gravity[0]), R(P.gravity[1]), R(P.gravity[2]));
        for (u32 i = 0; i < fluid.n; ++i)
        {
            V3<R> cur = m[i] * g;
            setv(fp, i, vec(fp, i) + (cur - vec(prev, i)));
            setv(prev, i, cur);
        }
    }

    // ref: fluid_dynamics/density_regularization.hpp:40-50 (inner), :74-83 (contact)
    void compressionSummation()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &Vref = fluid.r("VolumetricMeasureRef");
        const std::vector<R> &wpos = wall.r("Position", 3), &wVref = wall.r("VolumetricMeasureRef");
        std::vector<R> &sum = fluid.r("CompressionSummation");
        R W0 = K.W0();
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            R s = W0 * Vref[i];
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n)
            {
                u32 j = inner.index[n];
                s += K.W(innerDisp(pos, i, n)) * Vref[j];
            }
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n)
            {
                u32 j = contact.index[n];
                s += K.W(vec(pos, i) - vec(wpos, j)) * wVref[j];
            }
            sum[i] = s;
        }
    }
    // ref: density_regularization.hpp:109-118, density_regularization.h:145-184
    void densityRegularization()
    {
        const std::vector<R> &sum = fluid.r("CompressionSummation");
        std::vector<R> &C = fluid.r("Compression"), &rho = fluid.r("Density");
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            C[i] = P.free_surface ? SMAX(sum[i], R(1)) : sum[i];
            rho[i] = C[i] * rho0;
        }
    }
    // FreeSurfaceIndicationCK<Inner<WithUpdate>, Contact<>>; ref: general_dynamics/surface_indication/
    // surface_indication_ck.hpp:12-160 (constructor constants :19-20, inner interact :52-70, near-previous :72-87,
    // update :98-107, very-near :109-128, contact interact :149-160); sequencing interaction_algorithms_ck.cpp:6-34
    void surfaceIndication()
    {
        surfaceInteract();
        surfaceUpdate();
    }
    // inner interact (:52-70) + near-previous check (:72-87) + contact interact (:149-160)
    void surfaceInteract()
    {
        const u32 n = fluid.n;
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure");
        const std::vector<R> &wpos = wall.r("Position", 3), &wVol = wall.r("VolumetricMeasure");
        std::vector<R> &pos_div = fluid.r("PositionDivergence");
        std::vector<u32> &ind = fluid.uint["Indicator"], &prev = fluid.uint["PreviousSurfaceIndicator"];
        if (ind.size() != n) ind.assign(n, 0u);
        if (prev.size() != n) prev.assign(n, 1u); // registerStateVariable<int>("PreviousSurfaceIndicator", 1)
        const R threshold = R(0.75) * R(P.dim);
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)n; ++i)
        {
            R pd(0);
            for (u32 m = inner.offset[i]; m < inner.offset[i + 1]; ++m)
            {
                V3<R> d = innerDisp(pos, i, m);
                pd -= K.dW(d) * Vol[inner.index[m]] * d.norm();
            }
            if (pd < threshold && prev[i] != 1u)
            {
                bool near_previous = false;
                for (u32 m = inner.offset[i]; m < inner.offset[i + 1] && !near_previous; ++m)
                    near_previous = prev[inner.index[m]] == 1u;
                if (!near_previous) pd = R(2.0) * threshold;
            }
            R pw(0);
            for (u32 m = contact.offset[i]; m < contact.offset[i + 1]; ++m)
            {
                u32 j = contact.index[m];
                V3<R> d = vec(pos, i) - vec(wpos, j);
                pw -= K.dW(d) * wVol[j] * d.norm();
            }
            pos_div[i] = pd + pw;
        }
    }
    // update (:98-107) + very-near check (:109-128): reads PositionDivergence of the neighbours, i.e. the result of
    // surfaceInteract() on them (a slab-decomposed run refreshes it on the ghost planes in between)
    void surfaceUpdate()
    {
        const u32 n = fluid.n;
        const std::vector<R> &pos = fluid.r("Position", 3), &pos_div = fluid.r("PositionDivergence");
        std::vector<u32> &ind = fluid.uint["Indicator"], &prev = fluid.uint["PreviousSurfaceIndicator"];
        if (ind.size() != n) ind.assign(n, 0u);
        const R threshold = R(0.75) * R(P.dim), h = R(P.h_min);
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)n; ++i)
        {
            u32 v = 1u;
            if (pos_div[i] > threshold)
            {
                bool very_near = false;
                for (u32 m = inner.offset[i]; m < inner.offset[i + 1] && !very_near; ++m)
                {
                    u32 j = inner.index[m];
                    if (pos_div[j] < threshold && innerDisp(pos, i, m).norm() < h) very_near = true;
                }
                if (!very_near) v = 0u;
            }
            ind[i] = v;
        }
        prev = ind;
    }
    // observer probes: UpdateRelation<Contact<>> (observer -> fluid) + Interpolation<Contact<Real>> of "Pressure";
    // ref: update_body_relation.hpp:199-288, general_dynamics/interpolation_dynamics.hpp:44-60, io_observation_ck.h:69-94
    void observerRelation()
    {
        if (!observer.n) return;
        const Kernel<R> &k = K;
        auto crit = [&k](const R *a, const R *b) { return k.criterion(a, b); };
        buildContact(observer_contact, fluid_cl, observer.r("Position", 3).data(), observer.n, fluid.r("Position", 3).data(), 1, crit);
    }
    void observe(const std::string &name)
    {
        if (!observer.n) return;
        const std::vector<R> &opos = observer.r("Position", 3), &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure");
        const std::vector<R> &data = fluid.r(name);
        std::vector<R> &out = observer.r(name);
        for (u32 i = 0; i < observer.n; ++i)
        {
            R q(0), total(0);
            for (u32 m = observer_contact.offset[i]; m < observer_contact.offset[i + 1]; ++m)
            {
                u32 j = observer_contact.index[m];
                R w = K.W(vec(opos, i) - vec(pos, j)) * Vol[j];
                q += w * data[j];
                total += w;
            }
            out[i] = q / (total + R(2.71051e-20));
        }
    }
    // Interpolation<Contact<DataType, RestoringCorrection>>::InteractKernel::interact, general_dynamics/interpolation_dynamics.hpp:72-100:
    // the first-order consistent interpolation (reproduces constant and linear fields on any neighbour set whose restoring
    // matrix is regular; known answer: unit_test_interpolation_ck/2d_interpolation.cpp interpolates "Position" at a random
    // point of a randomised lattice and expects the point back to 1e-6). Matrices are (dim + 1) square; Eigen's inverse of
    // a fixed 3x3 / 4x4 matrix is the cofactor formula, of which only row 0 is needed.
    //   A(0,0) = W V, A(0,1+b) = -W V r_b, A(1+a,0) = dW V e_a, A(1+a,1+b) = -dW V r_a e_b;   restoring = Eps I + sum_j A_j
    //   prediction = sum_j A_j.col(0) data_j;   out = (restoring^-1).row(0) . prediction
    static R minorDet(const R M[4][4], int n, int skip_row, int skip_col)
    {
        int rr[3] = {0, 0, 0}, cc[3] = {0, 0, 0}, a = 0, b = 0;
        for (int k = 0; k < n; ++k)
        {
            if (k != skip_row) rr[a++] = k;
            if (k != skip_col) cc[b++] = k;
        }
        if (n == 3) return M[rr[0]][cc[0]] * M[rr[1]][cc[1]] - M[rr[0]][cc[1]] * M[rr[1]][cc[0]];
        return M[rr[0]][cc[0]] * (M[rr[1]][cc[1]] * M[rr[2]][cc[2]] - M[rr[1]][cc[2]] * M[rr[2]][cc[1]]) -
               M[rr[0]][cc[1]] * (M[rr[1]][cc[0]] * M[rr[2]][cc[2]] - M[rr[1]][cc[2]] * M[rr[2]][cc[0]]) +
               M[rr[0]][cc[2]] * (M[rr[1]][cc[0]] * M[rr[2]][cc[1]] - M[rr[1]][cc[1]] * M[rr[2]][cc[0]]);
    }
    void observeRestoring(const std::string &name, int width)
    {
        if (!observer.n) return;
        const int dim = P.dim, n = dim + 1;
        const std::vector<R> &opos = observer.r("Position", 3), &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure");
        const std::vector<R> &data = fluid.r(name, (size_t)width);
        std::vector<R> &out = observer.r("Restored" + name, (size_t)width);
        for (u32 i = 0; i < observer.n; ++i)
        {
            R M[4][4], pred[4][3];
            for (int a = 0; a < 4; ++a)
            {
                for (int b = 0; b < 4; ++b) M[a][b] = a == b ? std::numeric_limits<R>::epsilon() : R(0);
                for (int c = 0; c < 3; ++c) pred[a][c] = R(0);
            }
            for (u32 m = observer_contact.offset[i]; m < observer_contact.offset[i + 1]; ++m)
            {
                u32 j = observer_contact.index[m];
                V3<R> r = vec(opos, i) - vec(pos, j);
                V3<R> e = r.normalized();
                const R rv[3] = {r.x, r.y, r.z}, ev[3] = {e.x, e.y, e.z};
                R WV = K.W(r) * Vol[j], dWV = K.dW(r) * Vol[j];
                R col0[4];
                col0[0] = WV;
                M[0][0] += WV;
                for (int b = 0; b < dim; ++b) M[0][1 + b] += -WV * rv[b];
                for (int a = 0; a < dim; ++a)
                {
                    col0[1 + a] = dWV * ev[a];
                    M[1 + a][0] += dWV * ev[a];
                    for (int b = 0; b < dim; ++b) M[1 + a][1 + b] += -dWV * rv[a] * ev[b];
                }
                for (int a = 0; a < n; ++a)
                    for (int c = 0; c < width; ++c) pred[a][c] += col0[a] * data[(size_t)width * j + c];
            }
            // row 0 of the inverse: (M^-1)(0,k) = cofactor(k,0) / det
            R cof[4], det = R(0);
            for (int k = 0; k < n; ++k)
            {
                cof[k] = ((k & 1) ? R(-1) : R(1)) * minorDet(M, n, k, 0);
                det += M[k][0] * cof[k];
            }
            for (int c = 0; c < width; ++c)
            {
                R v = R(0);
                for (int k = 0; k < n; ++k) v += (cof[k] / det) * pred[k][c];
                out[(size_t)width * i + c] = v;
            }
        }
    }
    void recordProbes()
    {
        if (!observer.n) return;
        observe("Pressure");
        const std::vector<R> &out = observer.r("Pressure");
        probe_series.emplace_back(out.begin(), out.end());
    }
    // correction_(i): identity for NoKernelCorrectionCK, B_i for LinearCorrectionCK (kernel_correction_ck.h:42-187)
    inline M3<R> corrMat(const std::vector<R> &B, u32 i) const
    {
        if (P.correction) return mat(B, i);
        M3<R> m;
        std::memset(m.m, 0, sizeof(m.m));
        m.m[0] = m.m[4] = m.m[8] = R(1);
        return m;
    }
    // ViscousForceCK<Inner<WithUpdate, Viscosity, Correction>, Contact<Wall, ...>> + ForcePriorCK::UpdateKernel;
    // ref: fluid_dynamics/viscous_force.hpp:44-103, general_dynamics/force_prior_ck.h:53-57, materials/viscosity.h:60-66
    void viscousForce()
    {
        const u32 n = fluid.n;
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure"), &vel = fluid.r("Velocity", 3);
        const std::vector<R> &wpos = wall.r("Position", 3), &wVol = wall.r("VolumetricMeasure"), &wvel = wall.r("Velocity", 3);
        const std::vector<R> &B = fluid.r("LinearCorrectionMatrix", 9);
        std::vector<R> &F = fluid.r("ViscousForce", 3), &prev = fluid.r("PreviousViscousForce", 3), &Fp = fluid.r("ForcePrior", 3);
        const R mu = R(P.viscosity), h = R(P.h_min), eps = R(0.01) * h * h;
        const R mu_ij = R(2.0) * mu * mu / (mu + mu); // PairGeomAverageFixed
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)n; ++i)
        {
            V3<R> f, fw;
            const V3<R> vi = vec(vel, i);
            const M3<R> Bi = corrMat(B, i);
            for (u32 m = inner.offset[i]; m < inner.offset[i + 1]; ++m)
            {
                u32 j = inner.index[m];
                V3<R> d = innerDisp(pos, i, m);
                V3<R> e = d.normalized();
                R dWV = K.dW(d) * Vol[j];
                V3<R> vel_derivative = (vi - vec(vel, j)) / (d.squaredNorm() + eps);
                M3<R> Bs = Bi + corrMat(B, j);
                f += d.dot(Bs * e) * mu_ij * vel_derivative * dWV;
            }
            for (u32 m = contact.offset[i]; m < contact.offset[i + 1]; ++m)
            {
                u32 j = contact.index[m];
                V3<R> d = vec(pos, i) - vec(wpos, j);
                V3<R> e = d.normalized();
                R dWV = K.dW(d) * wVol[j];
                V3<R> vel_derivative = R(2.0) * (vi - vec(wvel, j)) / (d.squaredNorm() + eps);
                fw += R(2.0) * d.dot(Bi * e) * mu * vel_derivative * dWV;
            }
            V3<R> total = f * Vol[i] + fw * Vol[i];
            setv(F, i, total);
            setv(Fp, i, vec(Fp, i) + (total - vec(prev, i)));
            setv(prev, i, total);
        }
    }
    // KernelGradientIntegral<Inner<Correction>, Contact<Boundary, Correction>>; ref: general_dynamics/kernel_gradient_integral.hpp:33-78
    void kernelGradientIntegral()
    {
        const u32 n = fluid.n;
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure");
        const std::vector<R> &wpos = wall.r("Position", 3), &wVol = wall.r("VolumetricMeasure");
        const std::vector<R> &B = fluid.r("LinearCorrectionMatrix", 9);
        std::vector<R> &kgi = fluid.r("KernelGradientIntegral", 3);
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)n; ++i)
        {
            V3<R> g, gw;
            const M3<R> Bi = corrMat(B, i);
            for (u32 m = inner.offset[i]; m < inner.offset[i + 1]; ++m)
            {
                u32 j = inner.index[m];
                V3<R> d = innerDisp(pos, i, m);
                g -= ((Bi + corrMat(B, j)) * d.normalized()) * (K.dW(d) * Vol[j]);
            }
            for (u32 m = contact.offset[i]; m < contact.offset[i + 1]; ++m)
            {
                u32 j = contact.index[m];
                V3<R> d = vec(pos, i) - vec(wpos, j);
                gw -= (Bi * d.normalized()) * (R(2.0) * K.dW(d) * wVol[j]);
            }
            setv(kgi, i, g + gw);
        }
    }
    // TransportVelocityCorrectionCK<SPHBody, TruncatedLinear | NoLimiter, [BulkParticles]>;
    // ref: fluid_dynamics/transport_velocity_correction_ck.hpp:39-50, common/common_functors.h:69-94
    void transportVelocityCorrection(int limiter, bool bulk_only)
    {
        const std::vector<R> &kgi = fluid.r("KernelGradientIntegral", 3);
        std::vector<R> &dpos = fluid.r("Displacement", 3);
        const R h = R(P.h_min), h2 = h * h, scaling = R(P.transport_coefficient) * h2;
        const std::vector<u32> *ind = bulk_only ? &fluid.uint["Indicator"] : nullptr;
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            if (ind && (*ind)[i] != 0u) continue;
            V3<R> g = vec(kgi, i);
            R lim = limiter ? SMIN(R(100.0) * (h2 * g.squaredNorm()), R(1)) : R(1);
            setv(dpos, i, vec(dpos, i) + (scaling * lim) * g);
        }
    }
    // ref: fluid_time_step_ck.h:139-180
    void advectionSetup()
    {
        std::vector<R> &Vol = fluid.r("VolumetricMeasure"), &dpos = fluid.r("Displacement", 3);
        const std::vector<R> &m = fluid.r("Mass"), &rho = fluid.r("Density");
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            Vol[i] = m[i] / rho[i];
            dpos[3 * i] = dpos[3 * i + 1] = dpos[3 * i + 2] = R(0);
        }
    }
    void updatePosition()
    {
        std::vector<R> &pos = fluid.r("Position", 3);
        const std::vector<R> &dpos = fluid.r("Displacement", 3);
        #pragma omp parallel for schedule(static)
        for (size_t k = 0; k < size_t(3) * fluid.n; ++k) pos[k] += dpos[k];
    }
    // ref: fluid_time_step_ck.h:106-109, fluid_time_step_ck.cpp:24-27; TinyReal base_data_type.h:207
    // Slab-decomposed runs (oracle/decomposed.py): a rank stores its own particles first and the ghost planes behind
    // them; the time-step reductions then run over the own particles only and the ranks combine the raw values (max).
    long reduce_count = -1; // -1: all particles
    u32 reduceCount() const { return reduce_count < 0 ? fluid.n : std::min<u32>((u32)reduce_count, fluid.n); }
    double advectionDtReduced()
    {
        const std::vector<R> &vel = fluid.r("Velocity", 3);
        R red = std::numeric_limits<R>::lowest();
        const u32 n_red = reduceCount();
#pragma omp parallel for schedule(static) reduction(max : red)
        for (u32 i = 0; i < n_red; ++i) red = SMAX(red, vec(vel, i).squaredNorm());
        return double(red);
    }
    double advectionDt() { return advectionDtOf(advectionDtReduced()); }
    double advectionDtOf(double reduced)
    {
        R red = R(reduced);
        return double(R(P.advection_cfl) * R(P.h_min) / (SMAX(R(std::sqrt(red)), R(P.U_ref)) + R(2.71051e-20)));
    }
    // ref: fluid_time_step_ck.hpp:51-57, :31-36
    double acousticDtReduced()
    {
        const std::vector<R> &vel = fluid.r("Velocity", 3), &F = fluid.r("Force", 3), &Fp = fluid.r("ForcePrior", 3),
                             &m = fluid.r("Mass");
        R hmin = R(P.h_min);
        R red = std::numeric_limits<R>::lowest();
        const u32 n_red = reduceCount();
#pragma omp parallel for schedule(static) reduction(max : red)
        for (u32 i = 0; i < n_red; ++i)
        {
            R fn = (vec(F, i) + vec(Fp, i)).norm();
            R acc = std::sqrt(R(4.0) * hmin * fn / m[i]);
            red = SMAX(red, SMAX(c0 + vec(vel, i).norm(), acc));
        }
        return double(red);
    }
    double acousticDt() { return acousticDtOf(acousticDtReduced()); }
    double acousticDtOf(double reduced)
    {
        R red = R(reduced);
        return double(R(P.acoustic_cfl) * R(P.h_min) / (red + R(2.71051e-20)));
    }

    // ---- 1st half; ref: fluid_dynamics/acoustic_step_1st_half.hpp:66-74,89-111,122-127,157-180 ----
    void a1Init(R dt)
    {
        std::vector<R> &C = fluid.r("Compression"), &rho = fluid.r("Density"), &p = fluid.r("Pressure"),
                       &dpos = fluid.r("Displacement", 3);
        const std::vector<R> &Cd = fluid.r("CompressionRate"), &vel = fluid.r("Velocity", 3);
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            C[i] += R(0.5) * dt * Cd[i];
            rho[i] = C[i] * rho0;
            p[i] = p0 * (rho[i] / rho0 - R(1.0));
            setv(dpos, i, vec(dpos, i) + vec(vel, i) * dt * R(0.5));
        }
    }
    void a1Inner()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure"), &p = fluid.r("Pressure"),
                             &C = fluid.r("Compression"), &B = fluid.r("LinearCorrectionMatrix", 9);
        std::vector<R> &F = fluid.r("Force", 3), &Cd = fluid.r("CompressionRate");
        const bool corr = P.correction != 0;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            V3<R> fs;
            R diss(0);
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n)
            {
                u32 j = inner.index[n];
                V3<R> d = innerDisp(pos, i, n);
                R dWV = K.dW(d) * Vol[j];
                V3<R> e = d.normalized();
                if (corr)
                {
                    // AverageP on matrices: inv_sum * (B_j p_i Z + B_i p_j Z)
                    M3<R> Pm = (mat(B, j) * p[i] * Z + mat(B, i) * p[j] * Z) * inv_Z_sum;
                    fs -= (Pm * R(2.0) * dWV) * e;
                }
                else
                    fs -= AverageP(p[i], p[j]) * R(2.0) * dWV * e;
                diss += UJump(p[i] - p[j]) * dWV;
            }
            setv(F, i, vec(F, i) + fs * Vol[i]);
            Cd[i] = diss * C[i];
        }
    }
    void a1Wall()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure"), &p = fluid.r("Pressure"),
                             &C = fluid.r("Compression"), &rho = fluid.r("Density"), &m = fluid.r("Mass"),
                             &Fp = fluid.r("ForcePrior", 3), &B = fluid.r("LinearCorrectionMatrix", 9);
        const std::vector<R> &wpos = wall.r("Position", 3), &wVol = wall.r("VolumetricMeasure"),
                             &wacc = wall.r("Acceleration", 3);
        std::vector<R> &F = fluid.r("Force", 3), &Cd = fluid.r("CompressionRate");
        const bool corr = P.correction != 0;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            V3<R> fs;
            R diss(0);
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n)
            {
                u32 j = contact.index[n];
                V3<R> d = vec(pos, i) - vec(wpos, j);
                R dWV = K.dW(d) * wVol[j];
                V3<R> e = d.normalized();
                R r_ij = d.norm();
                R face_acc = (vec(Fp, i) / m[i] - vec(wacc, j)).dot(-e);
                R p_w = p[i] + rho[i] * r_ij * SMAX(R(0), face_acc);
                if (corr)
                    fs -= (mat(B, i) * (p[i] + p_w) * dWV) * e;
                else
                    fs -= (p[i] + p_w) * dWV * e;
                diss += UJump(p[i] - p_w) * dWV;
            }
            setv(F, i, vec(F, i) + fs * Vol[i]);
            Cd[i] += diss * C[i];
        }
    }
    void a1Update(R dt)
    {
        std::vector<R> &vel = fluid.r("Velocity", 3);
        const std::vector<R> &F = fluid.r("Force", 3), &Fp = fluid.r("ForcePrior", 3), &m = fluid.r("Mass");
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i) setv(vel, i, vec(vel, i) + (vec(Fp, i) + vec(F, i)) / m[i] * dt);
    }
    // ---- 2nd half; ref: fluid_dynamics/acoustic_step_2nd_half.hpp:33-38,53-73,83-89,117-137 ----
    void a2Init(R dt)
    {
        std::vector<R> &dpos = fluid.r("Displacement", 3);
        const std::vector<R> &vel = fluid.r("Velocity", 3);
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i) setv(dpos, i, vec(dpos, i) + vec(vel, i) * dt * R(0.5));
    }
    void a2Inner()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure"), &vel = fluid.r("Velocity", 3),
                             &C = fluid.r("Compression"), &B = fluid.r("LinearCorrectionMatrix", 9);
        std::vector<R> &F = fluid.r("Force", 3), &Cd = fluid.r("CompressionRate");
        const bool corr = P.correction != 0;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            R div(0);
            V3<R> pd;
            V3<R> vi = vec(vel, i);
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n)
            {
                u32 j = inner.index[n];
                V3<R> d = innerDisp(pos, i, n);
                R dWV = K.dW(d) * Vol[j];
                V3<R> e = d.normalized();
                V3<R> vj = vec(vel, j);
                V3<R> vave = AverageV(vi, vj);
                V3<R> ce = corr ? mat(B, i) * e : e;
                div += R(2.0) * (vi - vave).dot(ce) * dWV;
                R u = (vi - vj).dot(e);
                pd += PJump(u) * dWV * e;
            }
            Cd[i] += div * C[i];
            setv(F, i, pd * Vol[i]);
        }
    }
    void a2Wall()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure"), &vel = fluid.r("Velocity", 3),
                             &C = fluid.r("Compression"), &B = fluid.r("LinearCorrectionMatrix", 9);
        const std::vector<R> &wpos = wall.r("Position", 3), &wVol = wall.r("VolumetricMeasure"),
                             &wvel = wall.r("Velocity", 3), &wn = wall.r("NormalDirection", 3);
        std::vector<R> &F = fluid.r("Force", 3), &Cd = fluid.r("CompressionRate");
        const bool corr = P.correction != 0;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            R div(0);
            V3<R> pd;
            V3<R> vi = vec(vel, i);
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n)
            {
                u32 j = contact.index[n];
                V3<R> d = vec(pos, i) - vec(wpos, j);
                R dWV = K.dW(d) * wVol[j];
                V3<R> e = d.normalized();
                V3<R> vdiff = R(2.0) * (vi - vec(wvel, j));
                V3<R> ce = corr ? mat(B, i) * e : e;
                div += vdiff.dot(ce) * dWV;
                V3<R> nj = vec(wn, j);
                V3<R> nf = SGN(e.dot(nj)) * nj;
                R u = vdiff.dot(nf);
                pd += PJump(u) * dWV * nf;
            }
            Cd[i] += div * C[i];
            setv(F, i, vec(F, i) + pd * Vol[i]);
        }
    }
    void a2Update(R dt)
    {
        std::vector<R> &C = fluid.r("Compression"), &rho = fluid.r("Density");
        const std::vector<R> &Cd = fluid.r("CompressionRate");
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            C[i] += R(0.5) * dt * Cd[i];
            rho[i] = C[i] * rho0;
        }
    }
    // ref: interaction_algorithms_ck.cpp:29-34 (init -> inner -> contact -> update)
    void acoustic1(R dt) { a1Init(dt); a1Inner(); a1Wall(); a1Update(dt); }
    void acoustic2(R dt) { a2Init(dt); a2Inner(); a2Wall(); a2Update(dt); }

    // ref: general_dynamics/kernel_correction_ck.hpp:40-95; inverse with Tikhonov regularisation
    // (common/vector_functions: inverseTikhonov(B, eps) = (B^T B + eps I)^-1 B^T)
    void linearCorrection()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &Vol = fluid.r("VolumetricMeasure");
        const std::vector<R> &wpos = wall.r("Position", 3), &wVol = wall.r("VolumetricMeasure");
        std::vector<R> &B = fluid.r("LinearCorrectionMatrix", 9);
        R alpha = R(P.correction_alpha);
        const int dim = P.dim;
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            M3<R> b = M3<R>::Zero();
            auto acc = [&](const V3<R> &d, R vol) {
                V3<R> g = (K.dW(d) * d.normalized()) * vol; // nablaW_ij * V_j
                R rr[3] = {d.x, d.y, d.z}, gg[3] = {g.x, g.y, g.z};
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) b.m[3 * r + c] -= rr[r] * gg[c];
            };
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n)
            {
                u32 j = inner.index[n];
                acc(innerDisp(pos, i, n), Vol[j]);
            }
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n)
            {
                u32 j = contact.index[n];
                acc(vec(pos, i) - vec(wpos, j), wVol[j]);
            }
            if (dim == 2) b.m[8] = R(1); // embed the 2x2 block so determinant/inverse act on it alone
            R det = b.determinant();
            R det_sqr = SMAX(alpha - det, R(0));
            M3<R> bt = b.transpose();
            M3<R> btb = bt * b;
            R eps = R(1.0e-8);
            btb.m[0] += eps; btb.m[4] += eps; btb.m[8] += eps;
            M3<R> inv = btb.inverse() * bt;
            R wgt = det / (det + det_sqr);
            M3<R> out = inv * wgt + M3<R>::Identity() * (R(1.0) - wgt);
            std::memcpy(&B[9 * i], out.m, 9 * sizeof(R));
        }
    }

    // ref: general_dynamics/general_reduce_ck.h:52-88, external_force.h:53-56
    double mechanicalEnergy()
    {
        const std::vector<R> &pos = fluid.r("Position", 3), &vel = fluid.r("Velocity", 3), &m = fluid.r("Mass");
        V3<R> g(R(P.gravity[0]), R(P.gravity[1]), R(P.gravity[2]));
        R s(0);
        for (u32 i = 0; i < fluid.n; ++i)
            s += R(0.5) * m[i] * vec(vel, i).squaredNorm() + m[i] * g.dot(V3<R>() - vec(pos, i));
        return double(s);
    }

    // =================================================================================
    // legacy formulation (Integration1stHalf/2ndHalf, DensitySummation); state rho/drho_dt/pos
    // =================================================================================
    // ref: particle_dynamics/fluid_dynamics/density_summation.cpp:8-22,58-78, .hpp:28-32
    void legacyDensitySummation()
    {
        std::vector<R> &rho = fluid.r("Density"), &rsum = fluid.r("DensitySummation");
        const std::vector<R> &m = fluid.r("Mass"), &wm = wall.r("Mass");
        R W0 = K.factor_W; // Kernel::W0 = factor_W_dim (base_kernel.h:134-136)
        R inv_sigma0 = R(1.0) / R(P.sigma0);
        R inv_rho0_k = R(1.0) / R(P.wall_rho0);
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            R sigma = W0;
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n) sigma += in_W[n];
            R rs = sigma * rho0 * inv_sigma0;
            R sc(0);
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n) sc += ct_W[n] * inv_rho0_k * wm[contact.index[n]];
            rs += sc * rho0 * rho0 * inv_sigma0 / m[i];
            rsum[i] = rs;
            rho[i] = P.free_surface ? SMAX(rs, rho0) : rs;
        }
        if (!P.free_surface)
        {
            std::vector<R> &Vol = fluid.r("VolumetricMeasure");
            #pragma omp parallel for schedule(static)
            for (u32 i = 0; i < fluid.n; ++i) Vol[i] = m[i] / rho[i];
        }
    }
    // ref: particle_dynamics/fluid_dynamics/fluid_integration.hpp:49-113
    void legacy1(R dt)
    {
        std::vector<R> &rho = fluid.r("Density"), &p = fluid.r("Pressure"), &pos = fluid.r("Position", 3),
                       &drho = fluid.r("DensityChangeRate"), &F = fluid.r("Force", 3), &vel = fluid.r("Velocity", 3);
        const std::vector<R> &Vol = fluid.r("VolumetricMeasure"), &m = fluid.r("Mass"), &Fp = fluid.r("ForcePrior", 3);
        const std::vector<R> &wVol = wall.r("VolumetricMeasure"), &wacc = wall.r("Acceleration", 3);
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            rho[i] += drho[i] * dt * R(0.5);
            p[i] = p0 * (rho[i] / rho0 - R(1.0));
            setv(pos, i, vec(pos, i) + vec(vel, i) * dt * R(0.5));
        }
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            V3<R> f;
            R diss(0);
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n)
            {
                u32 j = inner.index[n];
                R dWV = in_dW[n] * Vol[j];
                V3<R> e(in_e[3 * n], in_e[3 * n + 1], in_e[3 * n + 2]);
                f -= (p[i] + p[j]) * dWV * e;
                diss += UJump(p[i] - p[j]) * dWV;
            }
            V3<R> fw;
            R dissw(0);
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n)
            {
                u32 j = contact.index[n];
                V3<R> e(ct_e[3 * n], ct_e[3 * n + 1], ct_e[3 * n + 2]);
                R dWV = ct_dW[n] * wVol[j];
                R face_acc = (vec(Fp, i) / m[i] - vec(wacc, j)).dot(-e);
                R p_w = p[i] + rho[i] * ct_r[n] * SMAX(R(0), face_acc);
                fw -= (p[i] + p_w) * dWV * e;
                dissw += UJump(p[i] - p_w) * dWV;
            }
            V3<R> Fi = vec(F, i) + f * Vol[i];
            R dr = diss * rho[i];
            Fi = Fi + fw * Vol[i];
            dr += dissw * rho[i];
            setv(F, i, Fi);
            drho[i] = dr;
        }
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i) setv(vel, i, vec(vel, i) + (vec(Fp, i) + vec(F, i)) / m[i] * dt);
    }
    // ref: fluid_integration.hpp:159-231
    void legacy2(R dt)
    {
        std::vector<R> &rho = fluid.r("Density"), &pos = fluid.r("Position", 3), &drho = fluid.r("DensityChangeRate"),
                       &F = fluid.r("Force", 3);
        const std::vector<R> &Vol = fluid.r("VolumetricMeasure"), &vel = fluid.r("Velocity", 3);
        const std::vector<R> &wVol = wall.r("VolumetricMeasure"), &wvel = wall.r("Velocity", 3),
                             &wn = wall.r("NormalDirection", 3);
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i) setv(pos, i, vec(pos, i) + vec(vel, i) * dt * R(0.5));
#pragma omp parallel for schedule(dynamic, 256)
        for (long i = 0; i < (long)fluid.n; ++i)
        {
            V3<R> vi = vec(vel, i);
            R dcr(0);
            V3<R> pd;
            for (u32 n = inner.offset[i]; n < inner.offset[i + 1]; ++n)
            {
                u32 j = inner.index[n];
                V3<R> e(in_e[3 * n], in_e[3 * n + 1], in_e[3 * n + 2]);
                R dWV = in_dW[n] * Vol[j];
                R u = (vi - vec(vel, j)).dot(e);
                dcr += u * dWV;
                pd += PJump(u) * dWV * e;
            }
            R dcrw(0);
            V3<R> pdw;
            for (u32 n = contact.offset[i]; n < contact.offset[i + 1]; ++n)
            {
                u32 j = contact.index[n];
                V3<R> e(ct_e[3 * n], ct_e[3 * n + 1], ct_e[3 * n + 2]);
                R dWV = ct_dW[n] * wVol[j];
                V3<R> nj = vec(wn, j);
                V3<R> nf = SGN(e.dot(nj)) * nj;
                V3<R> v_in_wall = R(2.0) * vec(wvel, j) - vi;
                dcrw += (vi - v_in_wall).dot(e) * dWV;
                R u = R(2.0) * (vi - vec(wvel, j)).dot(nf);
                pdw += PJump(u) * dWV * nf;
            }
            drho[i] += dcr * rho[i];
            drho[i] += dcrw * rho[i];
            setv(F, i, pd * Vol[i] + pdw * Vol[i]);
        }
        #pragma omp parallel for schedule(static)
        for (u32 i = 0; i < fluid.n; ++i) rho[i] += drho[i] * dt * R(0.5);
    }
    // ref: particle_dynamics/fluid_dynamics/fluid_time_step.cpp:21-59
    double legacyAcousticDt()
    {
        const std::vector<R> &vel = fluid.r("Velocity", 3);
        R red = std::numeric_limits<R>::lowest();
        #pragma omp parallel for schedule(static) reduction(max : red)
        for (u32 i = 0; i < fluid.n; ++i) red = SMAX(red, c0 + vec(vel, i).norm());
        return double(R(P.acoustic_cfl) * R(P.h_min) / (red + R(2.71051e-20)));
    }
    double legacyAdvectionDt()
    {
        const std::vector<R> &vel = fluid.r("Velocity", 3), &F = fluid.r("Force", 3), &Fp = fluid.r("ForcePrior", 3),
                             &m = fluid.r("Mass");
        R red = std::numeric_limits<R>::lowest();
        #pragma omp parallel for schedule(static) reduction(max : red)
        for (u32 i = 0; i < fluid.n; ++i)
        {
            R acc = R(4.0) * R(P.h_min) * (vec(F, i) + vec(Fp, i)).norm() / m[i];
            red = SMAX(red, SMAX(vec(vel, i).squaredNorm(), acc));
        }
        R speed_max = std::sqrt(red);
        return double(R(P.advection_cfl) * R(P.h_min) / (SMAX(speed_max, R(P.U_ref)) + R(2.71051e-20)));
    }

    // =================================================================================
    // time loops (host sequencing restated from the case files)
    // =================================================================================
    // ref: tests/tests_sycl/3d_examples/test_3d_dambreak_sycl/dambreak.cpp:152-225.
    // Runs until `end_time` or `max_outer` outer steps (whichever first); records total mechanical
    // energy at t=0 and every `record_interval` of physical time (as writeToFile does).
    void prepareCK()
    {
        ensureFluidState();
        gravityForce();
        cellListFluid();
        cellListWall();
        relationsCK();
        observerRelation();
        energy_series.clear(); time_series.clear(); probe_series.clear();
        energy_series.push_back(mechanicalEnergy()); time_series.push_back(physical_time);
        recordProbes(); // fluid_observer_pressure.writeToFile(number_of_iterations) before the loop, dambreak.cpp:177
    }
    long runCK(double end_time, long max_outer, double record_interval, int sort_interval)
    {
        long done = 0;
        while (physical_time < end_time && done < max_outer)
        {
            double integration_time = 0;
            while (integration_time < record_interval && done < max_outer)
            {
                compressionSummation();
                densityRegularization();
                advectionSetup();
                if (P.viscosity > 0) viscousForce(); // lid_driven_cavity_sycl.cpp:268-276: viscous force, [correction, indicator,] integral, transport
                if (P.transport_velocity)
                {
                    kernelGradientIntegral();
                    transportVelocityCorrection(1, false);
                }
                double adv_dt = advectionDt();
                if (P.surface_indicator) surfaceIndication(); // fluid_boundary_indicator.exec(), dambreak.cpp:192
                if (P.correction) linearCorrection();
                double relax = 0;
                while (relax < adv_dt)
                {
                    double dt = acousticDt();
                    acoustic1(R(dt));
                    acoustic2(R(dt));
                    relax += dt; integration_time += dt; physical_time += dt;
                    ++acoustic_steps;
                }
                updatePosition();
                ++outer_steps; ++done;
                if (sort_interval > 0 && outer_steps % sort_interval == 0 && outer_steps != 1) sortParticles(false);
                if (P.periodic_axes) periodicBounding(); // taylor_green.cpp:186-191: bounding, cell list, images, configuration
                cellListFluid();
                relationsCK();
                observerRelation(); // fluid_observer_contact_relation.exec(); fluid_observer_pressure.writeToFile(), :223-224
                recordProbes();
            }
            if (integration_time >= record_interval)
            {
                energy_series.push_back(mechanicalEnergy()); time_series.push_back(physical_time);
            }
        }
        return done;
    }
    // ref: tests/2d_examples/test_2d_dambreak/Dambreak.cpp:118-220 and
    // tests/3d_examples/test_3d_dambreak/dambreak.cpp:108-194 (same sequencing).
    // Energy is recorded at iteration 0 and every `observe_every` outer iterations (2-D case file),
    // or at every output interval when observe_every <= 0 (< 0: the 3-D case file's order, see runLegacy).
    void prepareLegacy()
    {
        ensureFluidState();
        gravityForce(); // GravityForce: force_prior = m g  (same increment form)
        cellListFluid();
        cellListWall();
        relationsLegacy();
        energy_series.clear(); time_series.clear();
        energy_series.push_back(mechanicalEnergy()); time_series.push_back(physical_time);
    }
    // observe_every >= 0: the order of the 2-D case file (Dambreak.cpp:118-220): the acoustic dt is reduced BEFORE the two half steps.
    //   > 0: energy and the probe are written at iteration 0 and every `observe_every`-th iteration (:137-139,175-180), before
    //   that iteration's configuration update (the legacy interpolation reads the kernel weights stored at the previous update;
    //   here it is evaluated on the current positions: one advection step apart, immaterial under the DTW criterion that judges
    //   it); == 0: energy at every output interval only.
    // observe_every < 0: the 3-D case file (test_3d_dambreak/dambreak.cpp:160-196): the half steps run with the dt reduced
    //   AFTER the previous pair (`Real dt = 0.0` before the loop, :151), the probes are written every iteration after the
    //   configuration update (:193-194), the energy at every output interval (:197).
    double legacy_dt = 0.0;
    long runLegacy(double end_time, long max_outer, double output_interval, int observe_every, int sort_interval)
    {
        long done = 0;
        if (observe_every > 0 && outer_steps == 0 && observer.n)
        {
            observerRelation();
            recordProbes();
        }
        while (physical_time < end_time && done < max_outer)
        {
            double integration_time = 0;
            while (integration_time < output_interval && done < max_outer)
            {
                double adv_dt = legacyAdvectionDt();
                legacyDensitySummation();
                double relax = 0;
                while (relax < adv_dt)
                {
                    double dt;
                    if (observe_every >= 0)
                    {
                        dt = legacyAcousticDt();
                        legacy1(R(dt));
                        legacy2(R(dt));
                    }
                    else
                    {
                        legacy1(R(legacy_dt));
                        legacy2(R(legacy_dt));
                        dt = legacy_dt = legacyAcousticDt();
                    }
                    relax += dt; integration_time += dt; physical_time += dt;
                    ++acoustic_steps;
                }
                if (observe_every > 0 && outer_steps % 100 == 0 && outer_steps % observe_every == 0 && outer_steps != 0)
                {
                    energy_series.push_back(mechanicalEnergy()); time_series.push_back(physical_time);
                    recordProbes();
                }
                ++outer_steps; ++done;
                if (sort_interval > 0 && outer_steps % sort_interval == 0 && outer_steps != 1) sortParticles(true);
                cellListFluid();
                relationsLegacy();
                observerRelation();
                if (observe_every < 0) recordProbes();
            }
            if (observe_every <= 0 && integration_time >= output_interval)
            {
                energy_series.push_back(mechanicalEnergy()); time_series.push_back(physical_time);
            }
        }
        return done;
    }
};

// type-erased handle
struct Handle
{
    int f64;
    Sim<float> *f;
    Sim<double> *d;
};

template <class R>
double execOp(Sim<R> &s, const std::string &op, double a0, double a1, double a2, double a3)
{
    s.ensureFluidState();
    if (op == "gravity") s.gravityForce();
    else if (op == "cell_list_fluid") s.cellListFluid();
    else if (op == "periodic_bounding") s.periodicBounding();
    else if (op == "cell_list_wall") s.cellListWall();
    else if (op == "relations") s.relationsCK();
    else if (op == "relations_legacy") s.relationsLegacy();
    else if (op == "sort") s.sortParticles(false);
    else if (op == "sort_legacy") s.sortParticles(true);
    else if (op == "compression_summation") s.compressionSummation();
    else if (op == "density_regularization") s.densityRegularization();
    else if (op == "advection_setup") s.advectionSetup();
    else if (op == "update_position") s.updatePosition();
    else if (op == "advection_dt") return s.advectionDt();
    else if (op == "advection_dt_reduced") return s.advectionDtReduced();
    else if (op == "acoustic_dt") return s.acousticDt();
    else if (op == "acoustic_dt_reduced") return s.acousticDtReduced();
    else if (op == "advection_dt_of") return s.advectionDtOf(a0);
    else if (op == "acoustic_dt_of") return s.acousticDtOf(a0);
    else if (op == "set_reduce_count") s.reduce_count = (long)a0;
    else if (op == "acoustic1") s.acoustic1(R(a0));
    else if (op == "acoustic2") s.acoustic2(R(a0));
    else if (op == "acoustic1_init") s.a1Init(R(a0));
    else if (op == "acoustic1_inner") s.a1Inner();
    else if (op == "acoustic1_wall") s.a1Wall();
    else if (op == "acoustic1_update") s.a1Update(R(a0));
    else if (op == "acoustic2_init") s.a2Init(R(a0));
    else if (op == "acoustic2_inner") s.a2Inner();
    else if (op == "acoustic2_wall") s.a2Wall();
    else if (op == "acoustic2_update") s.a2Update(R(a0));
    else if (op == "linear_correction") s.linearCorrection();
    else if (op == "surface_indication") s.surfaceIndication();
    else if (op == "surface_interact") s.surfaceInteract();
    else if (op == "surface_update") s.surfaceUpdate();
    else if (op == "viscous_force") s.viscousForce();
    else if (op == "kernel_gradient_integral") s.kernelGradientIntegral();
    else if (op == "transport_velocity_correction") s.transportVelocityCorrection((int)a0, a1 != 0.0);
    else if (op == "set_observers") { s.observer.n = (u32)a0; s.observer.real.clear(); }
    else if (op == "observer_relation") s.observerRelation();
    else if (op == "observe_pressure") s.observe("Pressure");
    else if (op == "observe_restoring_pressure") s.observeRestoring("Pressure", 1);
    else if (op == "observe_restoring_position") s.observeRestoring("Position", 3);
    else if (op == "probe_records") return (double)s.probe_series.size();
    else if (op == "energy") return s.mechanicalEnergy();
    else if (op == "legacy_density_summation") s.legacyDensitySummation();
    else if (op == "legacy1") s.legacy1(R(a0));
    else if (op == "legacy2") s.legacy2(R(a0));
    else if (op == "legacy_acoustic_dt") return s.legacyAcousticDt();
    else if (op == "legacy_advection_dt") return s.legacyAdvectionDt();
    else if (op == "prepare_ck") s.prepareCK();
    else if (op == "prepare_legacy") s.prepareLegacy();
    else if (op == "run_ck") return (double)s.runCK(a0, (long)a1, a2, (int)a3);
    else if (op == "run_legacy") return (double)s.runLegacy(a0, (long)a1, a2, (int)a3, 100);
    else if (op == "physical_time") return s.physical_time;
    else if (op == "acoustic_steps") return (double)s.acoustic_steps;
    else if (op == "outer_steps") return (double)s.outer_steps;
    else return -12345.0; // unknown op
    return 0.0;
}
} // namespace

// =====================================================================================
// C interface (ctypes)
// =====================================================================================
extern "C"
{
    void *orc_create(int f64, const ParamsPOD *p, const KernelPOD *k, const MeshPOD *fluid_mesh, const MeshPOD *wall_mesh,
                     u32 n_fluid, u32 n_wall)
    {
        Handle *h = new Handle{f64, nullptr, nullptr};
        if (f64)
        {
            h->d = new Sim<double>();
            h->d->init(*p, *k, *fluid_mesh, *wall_mesh);
            h->d->fluid.n = n_fluid; h->d->wall.n = n_wall;
        }
        else
        {
            h->f = new Sim<float>();
            h->f->init(*p, *k, *fluid_mesh, *wall_mesh);
            h->f->fluid.n = n_fluid; h->f->wall.n = n_wall;
        }
        return h;
    }
    void orc_destroy(void *hp)
    {
        Handle *h = (Handle *)hp;
        delete h->f; delete h->d; delete h;
    }
    // returns pointer to the named real array of body (0 fluid, 1 wall); creates it (zero) with `width` if absent
    void *orc_real(void *hp, int body, const char *name, int width, uint64_t *len)
    {
        Handle *h = (Handle *)hp;
        if (h->f64)
        {
            auto &b = body == 2 ? h->d->observer : (body ? h->d->wall : h->d->fluid);
            auto &v = b.r(name, width);
            *len = v.size();
            return v.data();
        }
        auto &b = body == 2 ? h->f->observer : (body ? h->f->wall : h->f->fluid);
        auto &v = b.r(name, width);
        *len = v.size();
        return v.data();
    }
    // named u32 arrays: fluid ids ("OriginalID","SortedID","SortKeys","Permutation") and structures
    // "fluid_cell_offset","fluid_particle_index","wall_cell_offset","wall_particle_index",
    // "inner_offset","inner_index","contact_offset","contact_index"
    void *orc_uint(void *hp, const char *name, uint64_t *len)
    {
        Handle *h = (Handle *)hp;
        std::string k(name);
        std::vector<u32> *v = nullptr;
#define PICK(S)                                                                                                       \
    if (k == "fluid_cell_offset") v = &S->fluid_cl.cell_offset;                                                      \
    else if (k == "fluid_particle_index") v = &S->fluid_cl.particle_index;                                           \
    else if (k == "wall_cell_offset") v = &S->wall_cl.cell_offset;                                                   \
    else if (k == "wall_particle_index") v = &S->wall_cl.particle_index;                                             \
    else if (k == "inner_offset") v = &S->inner.offset;                                                              \
    else if (k == "inner_index") v = &S->inner.index;                                                                \
    else if (k == "fluid_ext_index") v = &S->fluid_cl.ext_index;                                                     \
    else if (k == "contact_offset") v = &S->contact.offset;                                                          \
    else if (k == "contact_index") v = &S->contact.index;                                                            \
    else if (k == "observer_offset") v = &S->observer_contact.offset;                                                \
    else if (k == "observer_index") v = &S->observer_contact.index;                                                  \
    else { S->ensureFluidState(); v = &S->fluid.uint[k]; }
        if (h->f64) { PICK(h->d) } else { PICK(h->f) }
#undef PICK
        *len = v->size();
        return v->data();
    }
    double orc_exec(void *hp, const char *op, double a0, double a1, double a2, double a3)
    {
        Handle *h = (Handle *)hp;
        return h->f64 ? execOp(*h->d, op, a0, a1, a2, a3) : execOp(*h->f, op, a0, a1, a2, a3);
    }
    // recorded energy series of the last run_* call
    uint64_t orc_series(void *hp, double *times, double *energy, uint64_t cap)
    {
        Handle *h = (Handle *)hp;
        const std::vector<double> &e = h->f64 ? h->d->energy_series : h->f->energy_series;
        const std::vector<double> &t = h->f64 ? h->d->time_series : h->f->time_series;
        uint64_t m = std::min<uint64_t>(cap, e.size());
        for (uint64_t i = 0; i < m; ++i) { times[i] = t[i]; energy[i] = e[i]; }
        return e.size();
    }

    // recorded probe rows of the last prepare_ck/run_ck: out[row * n_probe + k]
    uint64_t orc_probe_series(void *hp, double *out, uint64_t cap_rows)
    {
        Handle *h = (Handle *)hp;
        const std::vector<std::vector<double>> &p = h->f64 ? h->d->probe_series : h->f->probe_series;
        for (uint64_t r = 0; r < p.size() && r < cap_rows; ++r)
            for (size_t k = 0; k < p[r].size(); ++k) out[r * p[r].size() + k] = p[r][k];
        return p.size();
    }

    // ---- stand-alone primitives (no Sim) ----
    // ref: common/algorithm_primitive.h:244-250; known answer
    // tests/unit_tests_src/.../test_exclusive_scan/test_exclusive_scan.cpp:9-27
    u32 orc_exclusive_scan_u32(const u32 *in, u32 *out, u32 n) { return exclusiveScan(in, out, n); }
    // cell ids (linear) and Morton keys of positions (packed 3 per particle), f32 or f64
    void orc_cell_keys(int f64, const void *pos, u32 n, const MeshPOD *mp, u32 *cell, u32 *key)
    {
        if (f64)
        {
            Mesh<double> m; m.set(*mp);
            const double *x = (const double *)pos;
            for (u32 i = 0; i < n; ++i) { int c[3]; m.cellIndex(x + 3 * i, c); cell[i] = m.linear(c); key[i] = mortonKey(c); }
        }
        else
        {
            Mesh<float> m; m.set(*mp);
            const float *x = (const float *)pos;
            for (u32 i = 0; i < n; ++i) { int c[3]; m.cellIndex(x + 3 * i, c); cell[i] = m.linear(c); key[i] = mortonKey(c); }
        }
    }
    // stable sort of (key, value) pairs ascending by key
    void orc_sort_pairs_u32(u32 *keys, u32 *vals, u32 n)
    {
        std::vector<u32> idx(n);
        std::iota(idx.begin(), idx.end(), 0u);
        std::stable_sort(idx.begin(), idx.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });
        std::vector<u32> k(n), v(n);
        for (u32 i = 0; i < n; ++i) { k[i] = keys[idx[i]]; v[i] = vals[idx[i]]; }
        std::memcpy(keys, k.data(), n * sizeof(u32));
        std::memcpy(vals, v.data(), n * sizeof(u32));
    }
    // tabulated kernel evaluation (for the closed-form pin); which: 0 W, 1 dW ; returns normalized value
    double orc_kernel_eval(int f64, const KernelPOD *kp, int which, double q)
    {
        if (f64) { Kernel<double> k; k.set(*kp); return k.interpolateCubic(which ? k.dw : k.w, q); }
        Kernel<float> k; k.set(*kp);
        return k.interpolateCubic(which ? k.dw : k.w, float(q));
    }
    int orc_max_threads()
    {
#ifdef _OPENMP
        return omp_get_max_threads();
#else
        return 1;
#endif
    }
}
