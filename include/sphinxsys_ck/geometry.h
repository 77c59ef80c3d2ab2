// sphinxsys_ck/geometry.h — host-side PODs the reference builds on the host and copies into every kernel object
// (mesh geometry, tabulated smoothing kernel), plus the small part of the pre-processing the dam-break case files
// use: box shapes, ComplexShape add/subtract, the lattice particle generator and wall normals.
// Everything is evaluated in `Real` exactly as written here and handed UNCHANGED to the device library (and to the
// CPU oracle in tests), so integer results can be compared bit for bit.
//
// Reference (relative to /root/reference/src):
//   Mesh ........................ shared/meshes/base_mesh.cpp:6-16, shared/meshes/cell_linked_list.cpp:14
//   system bounds +4dp .......... shared/sphinxsys_system/sph_system.cpp:39
//   SPHAdaptation ............... shared/adaptations/adaptation.cpp:12-19,80-85 (h = 1.3 dp, Wendland C2 default)
//   kernel table ................ shared/shared_ck/smoothing_kernel/kernel_tabulated_ck.cpp:6-25,
//                                 shared/kernels/kernel_wendland_c2.cpp:8-50, kernel_laguerre_gauss.cpp:8-50,
//                                 shared/kernels/base_kernel.h:87-93
//   shapes ...................... shared/geometries/geometric_element.{h,cpp}, base_geometry.cpp:45-59,118-140
//   lattice generator ........... for_3D_build/particle_generator/particle_generator_lattice_3d.cpp:12-25
#ifndef SPHINXSYS_CK_GEOMETRY_H
#define SPHINXSYS_CK_GEOMETRY_H

#include "base.h"

namespace SPH
{
// ---------------------------------------------------------------------------------------------------------
// Mesh(tentative_bounds, grid_spacing, buffer_width) in Real arithmetic
// ---------------------------------------------------------------------------------------------------------
inline sphb200_mesh_t makeMesh(const BoundingBoxd &bounds, Real spacing, int buffer_width, int dim)
{
    sphb200_mesh_t m;
    Real mesh_buffer = Real(buffer_width) * spacing;
    for (int d = 0; d < 3; ++d)
    {
        if (d >= dim)
        {
            m.lower[d] = 0;
            m.cells[d] = 1;
            continue;
        }
        Real lower = bounds.lower_[d] - mesh_buffer;
        Real tentative = (bounds.upper_[d] + mesh_buffer) - lower;
        int grid_pts = (int)std::ceil(tentative / spacing) + 1;
        m.lower[d] = lower;
        m.cells[d] = grid_pts - 1;
    }
    m.spacing = spacing;
    return m;
}

// ---------------------------------------------------------------------------------------------------------
// smoothing kernels (tabulated, 20 intervals + 4 guard entries)
// ---------------------------------------------------------------------------------------------------------
struct KernelWendlandC2 {};
struct KernelLaguerreGauss {};

inline sphb200_kernel_t makeKernel(Real h, int dim, int kind /*0 Wendland C2, 1 Laguerre-Gauss*/)
{
    sphb200_kernel_t k;
    k.dim = dim;
    k.kind = kind;
    k.h = h;
    k.src_h = h;
    k.kernel_size = Real(2.0);
    Real inv_h = Real(1.0) / h;
    Real dq = k.kernel_size / Real(20);
    for (int i = 0; i < 24; ++i)
    {
        double q = (double)(Real(i - 1) * dq); // table node in Real, analytic form in double, stored as Real
        double w, dw;
        if (kind == 0)
        {
            w = std::pow(1.0 - 0.5 * q, 4) * (1.0 + 2.0 * q);
            dw = 0.625 * std::pow(q - 2.0, 3) * q;
        }
        else
        {
            w = (1.0 - q * q + std::pow(q, 4) / 6.0) * std::exp(-(q * q));
            dw = (-std::pow(q, 5) / 3.0 + 8.0 * std::pow(q, 3) / 3.0 - 4.0 * q) * std::exp(-(q * q));
        }
        k.w[i] = (Real)w;
        k.dw[i] = (Real)dw;
    }
    const double pi = 3.14159265358979323846;
    double sigma;
    if (kind == 0)
        sigma = dim == 1 ? 3.0 / 4.0 : (dim == 2 ? 7.0 / (4.0 * pi) : 21.0 / (16.0 * pi));
    else
        sigma = dim == 1 ? 8.0 / (5.0 * std::sqrt(pi)) : (dim == 2 ? 3.0 / pi : 8.0 / std::pow(pi, 1.5));
    // factor_W_dim = inv_h^dim * sigma ; DimensionFactor = factor_W_dim * h^dim (base_kernel.h:91-93)
    Real inv_h_pow = Real(std::pow(inv_h, Real(dim)));
    Real factor = Real(inv_h_pow * Real(sigma));
    Real h_pow = Real(std::pow(h, Real(dim)));
    k.dimension_factor = Real(factor * h_pow);
    return k;
}

class SPHAdaptation
{
  public:
    Real global_resolution_, h_ref_;
    int dim_, kernel_kind_ = 0;
    sphb200_kernel_t kernel_;
    SPHAdaptation(Real resolution, int dim, Real h_spacing_ratio = Real(1.3))
        : global_resolution_(resolution), h_ref_(h_spacing_ratio * resolution), dim_(dim)
    {
        kernel_ = makeKernel(h_ref_, dim_, kernel_kind_);
    }
    Real ReferenceSmoothingLength() const { return h_ref_; }
    Real MinimumSmoothingLength() const { return h_ref_; }
    Real ReferenceSpacing() const { return global_resolution_; }
    Real CutOffRadius() const { return kernel_.kernel_size * kernel_.h; }
    // adaptation.h:96-100 resetKernel<KernelTabulated<KernelLaguerreGauss>>(20)
    template <class KernelType> void resetKernel()
    {
        kernel_kind_ = std::is_same<KernelType, KernelLaguerreGauss>::value ? 1 : 0;
        kernel_ = makeKernel(h_ref_, dim_, kernel_kind_);
    }
};

// ---------------------------------------------------------------------------------------------------------
// shapes
// ---------------------------------------------------------------------------------------------------------
struct Transform
{
    Vecd translation_;
    explicit Transform(const Vecd &t = Vecd()) : translation_(t) {}
};

class GeometricShapeBox
{
  public:
    double center_[3], halfsize_[3]; // kept in double: the case files give these as double literals
    GeometricShapeBox(const Transform &t, const Vecd &halfsize)
    {
        for (int d = 0; d < 3; ++d) { center_[d] = t.translation_[d]; halfsize_[d] = halfsize[d]; }
    }
    GeometricShapeBox(const double center[3], const double halfsize[3])
    {
        for (int d = 0; d < 3; ++d) { center_[d] = center[d]; halfsize_[d] = halfsize[d]; }
    }
    // closed-box containment on the transformed point (geometric_element.h:43-55)
    bool checkContain(const Vecd &p, int dim) const
    {
        for (int d = 0; d < dim; ++d)
            if (std::fabs(double(p[d]) - double(center_[d])) > double(halfsize_[d])) return false;
        return true;
    }
    // GeometricBox::findClosestPoint (geometric_element.cpp:19-61), evaluated in double
    void findClosestPoint(const double *p, int dim, double *out) const
    {
        double c[3], half[3];
        bool outside = false;
        for (int d = 0; d < dim; ++d)
        {
            c[d] = p[d] - double(center_[d]);
            half[d] = double(halfsize_[d]);
            if (std::fabs(c[d]) > half[d]) outside = true;
        }
        if (outside)
            for (int d = 0; d < dim; ++d) out[d] = std::fmin(std::fmax(c[d], -half[d]), half[d]) + double(center_[d]);
        else
        {
            int which = 0;
            double best = half[0] - std::fabs(c[0]);
            for (int d = 1; d < dim; ++d)
            {
                double v = half[d] - std::fabs(c[d]);
                if (v < best) { best = v; which = d; } // first axis wins ties
            }
            for (int d = 0; d < dim; ++d) out[d] = c[d] + double(center_[d]);
            out[which] = (c[which] < 0 ? -half[which] : half[which]) + double(center_[which]);
        }
    }
};

class ComplexShape
{
    struct Item
    {
        GeometricShapeBox box;
        bool add;
    };
    std::vector<Item> items_;
    std::string name_;

  public:
    explicit ComplexShape(const std::string &name) : name_(name) {}
    virtual ~ComplexShape() {}
    const std::string &Name() const { return name_; }
    template <class ShapeType> void add(const Transform &t, const Vecd &halfsize) { items_.push_back({ShapeType(t, halfsize), true}); }
    template <class ShapeType> void subtract(const Transform &t, const Vecd &halfsize) { items_.push_back({ShapeType(t, halfsize), false}); }
    template <class ShapeType> void add(const double center[3], const double halfsize[3]) { items_.push_back({ShapeType(center, halfsize), true}); }
    template <class ShapeType> void subtract(const double center[3], const double halfsize[3]) { items_.push_back({ShapeType(center, halfsize), false}); }
    // bounding box of the added sub-shapes (Shape::getBounds), rounded once to Real
    BoundingBoxd getBounds() const
    {
        double lo[3] = {1e300, 1e300, 1e300}, up[3] = {-1e300, -1e300, -1e300};
        bool any = false;
        for (const Item &it : items_)
            if (it.add)
            {
                any = true;
                for (int d = 0; d < 3; ++d)
                {
                    lo[d] = std::fmin(lo[d], it.box.center_[d] - it.box.halfsize_[d]);
                    up[d] = std::fmax(up[d], it.box.center_[d] + it.box.halfsize_[d]);
                }
            }
        if (!any) return BoundingBoxd();
        return BoundingBoxd(Vecd(Real(lo[0]), Real(lo[1]), Real(lo[2])), Vecd(Real(up[0]), Real(up[1]), Real(up[2])));
    }
    bool checkContain(const Vecd &p, int dim) const
    {
        bool in = false;
        for (const Item &it : items_)
        {
            if (it.add) in = in || it.box.checkContain(p, dim);
            else in = in && !it.box.checkContain(p, dim);
        }
        return in;
    }
    // unit vector from p to the closest point over all sub-shape surfaces; later shapes win ties ('<=')
    Vecd directionToSurface(const Vecd &p, int dim) const
    {
        double pd[3] = {p.x, p.y, p.z}, best[3] = {0, 0, 0}, best_dist = 1e300;
        for (const Item &it : items_)
        {
            double cp[3] = {0, 0, 0};
            it.box.findClosestPoint(pd, dim, cp);
            double d2 = 0;
            for (int d = 0; d < dim; ++d) d2 += (pd[d] - cp[d]) * (pd[d] - cp[d]);
            double dist = std::sqrt(d2);
            if (dist <= best_dist)
            {
                best_dist = dist;
                for (int d = 0; d < 3; ++d) best[d] = cp[d];
            }
        }
        double disp[3] = {0, 0, 0}, nrm = 0;
        for (int d = 0; d < dim; ++d)
        {
            disp[d] = best[d] - pd[d];
            nrm += disp[d] * disp[d];
        }
        nrm = std::sqrt(nrm);
        if (nrm == 0) nrm = 1.0;
        return Vecd(Real(disp[0] / nrm), Real(disp[1] / nrm), Real(disp[2] / nrm));
    }
};

// Lattice generator: cell centres of Mesh(system bounds, dp, buffer 0); loop order x -> y -> z
inline std::vector<Vecd> generateLattice(const ComplexShape &shape, const BoundingBoxd &system_bounds, Real dp, int dim)
{
    std::vector<Real> axis[3];
    for (int d = 0; d < 3; ++d)
    {
        if (d >= dim)
        {
            axis[d].push_back(0);
            continue;
        }
        Real lower = system_bounds.lower_[d], upper = system_bounds.upper_[d];
        int n = (int)std::ceil((upper - lower) / dp);
        for (int i = 0; i < n; ++i) axis[d].push_back((lower + Real(i) * dp) + Real(0.5) * dp);
    }
    std::vector<Vecd> out;
    for (Real x : axis[0])
        for (Real y : axis[1])
            for (Real z : axis[2])
            {
                Vecd p(x, y, z);
                if (shape.checkContain(p, dim)) out.push_back(p);
            }
    return out;
}
} // namespace SPH
#endif
