// fluid.cu — weakly-compressible SPH fluid dynamics on sm_100a: density (compression) summation,
// acoustic 1st/2nd half with wall + Riemann dissipation, time-step reductions, advection set-up.
//
// Replaces the device launches of (paths relative to /root/reference/src/shared/shared_ck/particle_dynamics):
//   fluid_dynamics/density_regularization.hpp:40-118     CompressionSummation<Inner/Contact>, DensityRegularization
//   fluid_dynamics/acoustic_step_1st_half.hpp:66-180     initialize / interact(inner) / interact(wall) / update
//   fluid_dynamics/acoustic_step_2nd_half.hpp:33-137     initialize / interact(inner) / interact(wall) / update
//   fluid_dynamics/fluid_time_step_ck.{h,hpp,cpp}        AcousticTimeStepCK, AdvectionTimeStepCK, AdvectionStepSetup,
//                                                        UpdateParticlePosition
//   general_dynamics/force_prior_ck.hpp:38-44            GravityForceCK
//   general_dynamics/kernel_correction_ck.hpp:40-95      LinearCorrectionMatrix<Inner<WithUpdate>,Contact<>>
//   general_dynamics/general_reduce_ck.h:52-88           TotalMechanicalEnergyCK
//
// Design: one thread per fluid particle; neighbour rows are read from the SELL-32 relation
// (one coalesced 128-byte load per warp and neighbour slot); neighbour state is gathered as float4 records
// (x, y, z, Vol) / (vx, vy, vz, -) through L1; the tabulated smoothing kernel (4-point Lagrange on a 24-entry
// table, kernel_tabulated_ck.h:45-58) is evaluated from per-interval cubic coefficients staged in shared
// memory (one LDS.128 + 3 FMA instead of 4 table loads, 4 divisions and ~25 flops); the reference's four
// launches per half step are fused wherever no neighbour reads the value being written.
#include "common.cuh"

// SPH_TRIM (tuning bit mask, scripts/gpu_variants.sh): 1 = wendland_dw with q folded into the scales, 2 = unit vector and
// impedance folded out of the 2nd-half pair loop, 4 = InvImpedanceAve folded out of the 1st-half pair loop
#ifndef SPH_TRIM
#define SPH_TRIM 6 // measured (profiles/r02_kernel_variants.jsonl): bit 1 costs 4 % in k_a1_interact, bit 2 gains 1.6 % in k_a2, bit 4 is neutral
#endif

constexpr int FL_THREADS = 128;
#ifndef SPH_FL_MIN_BLOCKS
#define SPH_FL_MIN_BLOCKS 8
#endif
constexpr int FL_MIN_BLOCKS = SPH_FL_MIN_BLOCKS; // 8: <= 64 registers, 50% occupancy with 4 gathers in flight per warp (scripts/gpu_variants.sh sweep)
constexpr int KT_INTERVALS = 21; // intervals of the 24-entry table that q in [0, 2] can select
constexpr int KT_SLOTS = 32;     // padded so that (index & 31) can never leave the table

struct KTab
{
    float4 c[KT_SLOTS];
};

struct FArgs
{
    // fluid
    u32 n;
    u32 begin, end; // active slots [begin, end): ghost slots outside are read as neighbours but never written
    float4 *pos, *vel, *dpos, *force, *force_prior, *posvol;
    const float4 *posvolref; // (x, y, z, VolRef)
    float4 *rec2;            // 32-byte records (x, y, z, Vol | vx, vy, vz, -)
    float *vol, *mass, *rho, *p, *C, *Cdot, *vol_ref, *Csum, *B;
    int p_in_brec; // 1: the correction variants run: the 1st-half initialize also keeps the pressure slot of brec current
    float4 *brec; // LinearCorrectionRecord: (Bxx, Bxy, Bxz, Byy | Byz, Bzz, p, -): the symmetric part of B and the pressure as one 32-byte gather record, or nullptr
    // wall
    u32 n_wall;
    const float4 *w_pos, *w_posvol, *w_posvolref, *w_vel, *w_acc, *w_n;
    const float *w_vol_ref;
    // relations
    const u32 *in_count, *in_slice, *in_index;
    const u32 *ct_count, *ct_slice, *ct_index;
    const u32 *order; // slot -> particle id (the relations' row order), or nullptr
    // constants
    float inv_h, inv_dq, q_scale, W0;
    // closed form of the tabulated Wendland C2 interpolant (see wendland_*): scales and error-term coefficients
    int analytic;
    float wl_w_scale, wl_w_c4, wl_w_c5, wl_four_dq, wl_dw_a, wl_dw_ah, wl_dw_c; // wl_dw_ah = wl_dw_a / h
    float rho0, c0, p0, Z, inv_Z_sum, inv_Z_ave, Z_geo, inv_c_ave, limiter;
    float lim_k; // limiter * inv_c_ave: TruncatedLinear(InvSoundSpeedAve * max(u, 0)) in one multiply
    int free_surface, dim;
    int legacy;               // 1: state is Density/DensityChangeRate (in rho / Cdot), analytic kernel
    float lg_sum_scale;       // legacy DensitySummation: rho0 / sigma0
    float lg_wall_scale;      // rho0^2 / sigma0
};

// Per-interval cubic coefficients of the 4-point Lagrange interpolant (kernel_tabulated_ck.h:45-58).
// With t = q/dq - location in [0,1) the interpolant is a0 + a1 t + a2 t^2 + a3 t^3; the table stores it
// re-expanded in the CENTRED variable s = t - 1/2 in [-1/2, 1/2] (what the magic-number rounding in eval_tab
// produces), pre-multiplied by `scale` (inv_h^dim * dimension_factor for W, inv_h^(dim+1) * ... for dW).
static void build_tab(const float *data, double scale, KTab *out)
{
    for (int loc = 0; loc < KT_SLOTS; ++loc)
    {
        int l = loc < KT_INTERVALS ? loc : KT_INTERVALS - 1;
        double d0 = data[l], d1 = data[l + 1], d2 = data[l + 2], d3 = data[l + 3];
        double a0 = d1;
        double a1 = -d0 / 3.0 - d1 / 2.0 + d2 - d3 / 6.0;
        double a2 = d0 / 2.0 - d1 + d2 / 2.0;
        double a3 = -d0 / 6.0 + d1 / 2.0 - d2 / 2.0 + d3 / 6.0;
        // t = s + 1/2
        double b0 = a0 + a1 / 2.0 + a2 / 4.0 + a3 / 8.0;
        double b1 = a1 + a2 + 3.0 * a3 / 4.0;
        double b2 = a2 + 3.0 * a3 / 2.0;
        double b3 = a3;
        out->c[loc] = make_float4((float)(b0 * scale), (float)(b1 * scale), (float)(b2 * scale), (float)(b3 * scale));
    }
}

static int make_fargs(sphb200_context *ctx, const sphb200_fluid_args_t *s, FArgs *a, KTab *wtab, KTab *dwtab)
{
    const sphb200_fluid_view_t &f = s->fluid;
    a->n = f.n;
    a->begin = f.active_end ? f.active_begin : 0u;
    a->end = f.active_end ? f.active_end : f.n;
    if (a->begin > a->end || a->end > f.n)
    {
        snprintf(ctx->err, sizeof(ctx->err), "active range [%u, %u) outside [0, %u)", a->begin, a->end, f.n);
        return SPHB200_E_INVALID;
    }
    a->pos = (float4 *)f.pos; a->vel = (float4 *)f.vel; a->dpos = (float4 *)f.dpos;
    a->force = (float4 *)f.force; a->force_prior = (float4 *)f.force_prior; a->posvol = (float4 *)f.posvol;
    a->posvolref = (const float4 *)f.posvolref; a->rec2 = (float4 *)f.posvolvel;
    a->vol = f.vol; a->mass = f.mass; a->rho = f.rho; a->p = f.p; a->C = f.compression; a->Cdot = f.compression_rate;
    a->vol_ref = f.vol_ref; a->Csum = f.compression_sum; a->B = f.B; a->brec = (float4 *)f.correction_record;
    const sphb200_wall_view_t &w = s->wall;
    a->n_wall = w.n;
    a->w_pos = (const float4 *)w.pos; a->w_posvol = (const float4 *)w.posvol; a->w_posvolref = (const float4 *)w.posvolref; a->w_vel = (const float4 *)w.vel_ave;
    a->w_acc = (const float4 *)w.acc_ave; a->w_n = (const float4 *)w.normal; a->w_vol_ref = w.vol_ref;
    a->in_count = s->inner.count; a->in_slice = s->inner.slice_offset; a->in_index = s->inner.index;
    a->ct_count = w.n ? s->contact.count : nullptr;
    a->ct_slice = w.n ? s->contact.slice_offset : nullptr;
    a->ct_index = w.n ? s->contact.index : nullptr;
    a->order = s->inner.order;
    if (w.n && s->contact.order != s->inner.order)
    {
        snprintf(ctx->err, sizeof(ctx->err), "inner and contact relations must share one slot order");
        return SPHB200_E_INVALID;
    }
    const sphb200_kernel_t &k = s->kernel;
    if (k.dim != 2 && k.dim != 3)
    {
        snprintf(ctx->err, sizeof(ctx->err), "kernel.dim must be 2 or 3");
        return SPHB200_E_UNSUPPORTED;
    }
    float inv_h = 1.0f / k.h, src_inv_h = 1.0f / k.src_h;
    double ih = inv_h, sih = src_inv_h;
    double w_scale = (k.dim == 2 ? ih * ih : ih * ih * ih) * k.dimension_factor;
    double dw_scale = w_scale * ih;
    if (wtab) build_tab(k.w, w_scale, wtab);
    if (dwtab) build_tab(k.dw, dw_scale, dwtab);
    // Wendland C2 is a polynomial (W_1D degree 5, dW_1D degree 4), so the reference's 4-point Lagrange interpolant on
    // its table equals f(q) - f[q0,q1,q2,q3,q] * prod(q - q_k) EXACTLY, with the divided difference in closed form
    // (a4 + a5 * (q0+q1+q2+q3+q)). The kernels use that form when the table handed in really is the Wendland table
    // (checked here node by node); any other table (Laguerre-Gauss, user kernels) takes the shared-memory table path.
    {
        const double dq = (double)k.kernel_size / 20.0;
        bool is_wendland = k.kernel_size == 2.0f;
        for (int i = 0; i < 24 && is_wendland; ++i)
        {
            double q = (double)((float)(i - 1) * (float)dq);
            double w = pow(1.0 - 0.5 * q, 4) * (1.0 + 2.0 * q), dw = 0.625 * pow(q - 2.0, 3) * q;
            if (fabs((double)k.w[i] - w) > 2e-6 || fabs((double)k.dw[i] - dw) > 2e-6) is_wendland = false;
        }
        a->analytic = is_wendland ? 1 : 0;
        a->legacy = s->material.formulation == 1;
        if (a->legacy && !is_wendland)
        {
            snprintf(ctx->err, sizeof(ctx->err), "legacy formulation: only the Wendland C2 kernel has a device path");
            return SPHB200_E_UNSUPPORTED;
        }
        // legacy evaluates the analytic kernel: same closed form with the interpolation-error term switched off
        const double dq4 = a->legacy ? 0.0 : dq * dq * dq * dq;
        a->wl_w_scale = (float)w_scale;
        a->wl_w_c4 = (float)((-0.9375 + 2.0 * 0.125 * dq) * dq4 * w_scale);
        a->wl_w_c5 = (float)(0.125 * dq4 * w_scale);
        a->wl_four_dq = (float)(4.0 * dq);
        a->wl_dw_a = (float)(0.625 * dw_scale);
        a->wl_dw_ah = (float)(0.625 * dw_scale * ih);
        a->wl_dw_c = (float)(0.625 * dq4 * dw_scale);
    }
    a->inv_h = inv_h;
    a->inv_dq = 20.0f / k.kernel_size;
    a->q_scale = inv_h * a->inv_dq; // r -> q / dq in one multiply
    a->W0 = (float)((k.dim == 2 ? sih * sih : sih * sih * sih) * k.dimension_factor * k.w[1]);
    a->lg_sum_scale = s->material.sigma0 > 0.f ? s->material.rho0 / s->material.sigma0 : 0.f;
    a->lg_wall_scale = s->material.sigma0 > 0.f ? s->material.rho0 * s->material.rho0 / s->material.sigma0 : 0.f;
    const sphb200_fluid_t &m = s->material;
    if (m.riemann < 0 || m.riemann > 2 || m.correction < 0 || m.correction > 1)
    {
        snprintf(ctx->err, sizeof(ctx->err), "unsupported riemann/correction kind");
        return SPHB200_E_UNSUPPORTED;
    }
    // ImpedanceModel, riemann_solver_ck.hpp:58-69 (same fluid on both sides), evaluated in Real
    a->rho0 = m.rho0; a->c0 = m.c0;
    a->p0 = m.rho0 * m.c0 * m.c0;
    a->Z = m.rho0 * m.c0;
    a->inv_Z_sum = 1.0f / (a->Z + a->Z);
    a->inv_Z_ave = (a->Z + a->Z) / (a->Z * a->Z + a->Z * a->Z);
    a->Z_geo = 2.0f * a->Z * a->Z * a->inv_Z_sum;
    a->inv_c_ave = 0.5f * (m.rho0 + m.rho0) * a->inv_Z_ave;
    a->limiter = m.limiter_coeff;
    a->lim_k = m.limiter_coeff * a->inv_c_ave;
    a->free_surface = m.free_surface;
    a->dim = k.dim;
    a->p_in_brec = m.correction && a->brec;
    return 0;
}

// slot handled by this thread: launches start at the 32-aligned slot below `begin` so that lane == slot % 32
// (the SELL-32 relation layout relies on it)
__device__ __forceinline__ u32 active_slot(const FArgs &a) { return (a.begin & ~31u) + blockIdx.x * blockDim.x + threadIdx.x; }
static inline unsigned active_blocks(const FArgs &a, unsigned threads) { return sph_blocks(a.end - (a.begin & ~31u), threads); }

__device__ __forceinline__ void stage_tab(const KTab &src, float4 *dst)
{
    if (threadIdx.x < KT_SLOTS) dst[threadIdx.x] = src.c[threadIdx.x];
    __syncthreads();
}

// Tabulated-kernel value (already scaled) at distance r: u = r * q_scale - 1/2 = q/dq - 1/2; adding 1.5 * 2^23
// rounds u to the nearest integer = floor(q/dq) (ties land on a node, where adjacent intervals agree), leaves
// that integer in the low mantissa bits, and s = u - round(u) is the centred local coordinate.
// FFMA + 2 FADD + LOP + LDS.128 + 3 FFMA, no conversion/SFU instructions.
__device__ __forceinline__ float eval_tab(const float4 *tab, float r, float q_scale)
{
    const float MAGIC = 12582912.0f; // 1.5 * 2^23
    float u = fmaf(r, q_scale, -0.5f);
    float m = u + MAGIC;
    float s = u - (m - MAGIC);
    float4 c = tab[__float_as_int(m) & (KT_SLOTS - 1)];
    return fmaf(fmaf(fmaf(c.w, s, c.z), s, c.y), s, c.x);
}

// Closed form of the same interpolant for the Wendland C2 table (no memory access): with t = q/dq - floor(q/dq) and
// s = t - 1/2, prod(q - q_k) = dq^4 (s^2 - 9/4)(s^2 - 1/4); dW_1D = 0.625 (q-2)^3 q has the constant divided
// difference 0.625, W_1D = (1-q/2)^4 (1+2q) has a4 + a5 (q0+q1+q2+q3+q) with a4 = -15/16, a5 = 1/8.
__device__ __forceinline__ float wendland_dw(const FArgs &a, float r)
{
    const float MAGIC = 12582912.0f;
    float u = fmaf(r, a.q_scale, -0.5f);
    float m = u + MAGIC;
    float s = u - (m - MAGIC);
    float s2 = s * s;
#if SPH_TRIM & 1
    float pi = fmaf(s2, s2 - 2.5f, 0.5625f); // (s2 - 9/4)(s2 - 1/4)
    float g = fmaf(r, a.inv_h, -2.0f);       // q - 2
    float poly = (g * g) * (g * (r * a.wl_dw_ah)); // 0.625 scale q (q - 2)^3
#else
    float q = r * a.inv_h;
    float pi = (s2 - 2.25f) * (s2 - 0.25f);
    float g = q - 2.0f;
    float poly = (g * g) * (g * (q * a.wl_dw_a));
#endif
    return fmaf(-a.wl_dw_c, pi, poly);
}
__device__ __forceinline__ float wendland_w(const FArgs &a, float r)
{
    const float MAGIC = 12582912.0f;
    float q = r * a.inv_h;
    float u = fmaf(r, a.q_scale, -0.5f);
    float m = u + MAGIC;
    float loc = m - MAGIC;
    float s = u - loc;
    float s2 = s * s;
    float pi = (s2 - 2.25f) * (s2 - 0.25f);
    float h = fmaf(-0.5f, q, 1.0f);
    float h2 = h * h;
    float poly = (h2 * h2) * (fmaf(2.0f, q, 1.0f) * a.wl_w_scale);
    float dd = fmaf(a.wl_w_c5, fmaf(loc, a.wl_four_dq, q), a.wl_w_c4);
    return fmaf(-dd, pi, poly);
}
template <bool ANALYTIC> __device__ __forceinline__ float kernel_dw(const FArgs &a, const float4 *tab, float r)
{
    return ANALYTIC ? wendland_dw(a, r) : eval_tab(tab, r, a.q_scale);
}
template <bool ANALYTIC> __device__ __forceinline__ float kernel_w(const FArgs &a, const float4 *tab, float r)
{
    return ANALYTIC ? wendland_w(a, r) : eval_tab(tab, r, a.q_scale);
}

// one 256-bit load (LDG.E.256 on sm_100a) of a 32-byte record that is not written during the launch
__device__ __forceinline__ void load_rec2(const float4 *rec, u32 j, float4 &a, float4 &b)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(rec + 2ull * j));
}

// 1/sqrt(x) for x > 0 (MUFU.RSQ, flush-to-zero form: no denormal fix-up code); callers clamp x away from 0
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// |d| and 1/|d| with the convention of Eigen's normalized(): a zero vector stays zero (inv_r finite, d * inv_r = 0)
__device__ __forceinline__ void dist(float r2, float &r, float &inv_r)
{
    inv_r = fast_rsqrt(fmaxf(r2, 1.0e-30f));
    r = r2 * inv_r;
}

__device__ __forceinline__ float3 mat_vec(const float *B, float3 v)
{
    return make_float3(B[0] * v.x + B[1] * v.y + B[2] * v.z, B[3] * v.x + B[4] * v.y + B[5] * v.z,
                       B[6] * v.x + B[7] * v.y + B[8] * v.z);
}
__device__ __forceinline__ void load_mat(const float *B, u32 i, float *out)
{
#pragma unroll
    for (int k = 0; k < 9; ++k) out[k] = B[9ull * i + k];
}


// -----------------------------------------------------------------------------------------------------
// Neighbour loop with explicit memory-level parallelism. A plain `for (k < cnt)` loop has a lane-dependent exit, so
// the compiler may not hoist the index load or the gather of iteration k+1 above the exit test of iteration k: every
// iteration then pays index latency (the index stream comes from DRAM) + gather latency back to back, and the kernels
// were latency bound at one outstanding gather per warp (profiles/r01_v4_*). Here the rows of a slot are walked in
// batches of U: the U indices of the NEXT batch are loaded while the current batch is gathered and evaluated, and the
// U gathers of a batch are issued together. Rows past the end are clamped to the last valid row (always a legal
// index) and masked out by `valid`, which the bodies fold into the pair weight.
//   gather(u, j)        : load what the pair needs into slot u of the body's staging registers
//   compute(u, valid)   : evaluate pair u; valid == false must contribute nothing
// -----------------------------------------------------------------------------------------------------
#ifndef SPH_NB_U
#define SPH_NB_U 4
#endif
constexpr int NB_U = SPH_NB_U;
// SPH_PREFETCH (tuning, scripts/gpu_variants.sh): 1 = prefetch.global.L2, 2 = prefetch.global.L1 of the NEXT batch's records
// as soon as its indices have arrived (costs no registers: hides part of the gather latency the 4-deep batches leave)
#ifndef SPH_PREFETCH
#define SPH_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_record(const void *p)
{
#if SPH_PREFETCH == 1
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#elif SPH_PREFETCH == 2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
struct NoPrefetch
{
    __device__ __forceinline__ void operator()(u32) const {}
};
template <int U = NB_U, class Gather, class Compute, class Prefetch = NoPrefetch>
__device__ __forceinline__ void for_neighbors(const u32 *__restrict__ idx, u32 cnt, Gather gather, Compute compute, Prefetch prefetch = Prefetch())
{
    if (cnt == 0) return;
    const u32 last = cnt - 1u;
    u32 j[U];
#pragma unroll
    for (int u = 0; u < U; ++u) j[u] = idx[32ull * min((u32)u, last)];
    for (u32 k = 0; k < cnt; k += U)
    {
        u32 jn[U];
#pragma unroll
        for (int u = 0; u < U; ++u) jn[u] = idx[32ull * min(k + U + (u32)u, last)];
#pragma unroll
        for (int u = 0; u < U; ++u) gather(u, j[u]);
#if SPH_PREFETCH
#pragma unroll
        for (int u = 0; u < U; ++u) prefetch(jn[u]);
#endif
#pragma unroll
        for (int u = 0; u < U; ++u) compute(u, k + (u32)u < cnt);
#pragma unroll
        for (int u = 0; u < U; ++u) j[u] = jn[u];
    }
}
#ifndef SPH_SUM_U
#define SPH_SUM_U 4
#endif
constexpr int SUM_U = SPH_SUM_U; // the summation kernel is light on registers (32): deeper batches are affordable there
constexpr int NB_WALL_U = 2; // wall pairs need up to three records each: smaller batches keep the kernels spill-free

// RiemannSolver<...>::ComputingKernel::DissipativePJump, riemann_solver_ck.hpp:44-49, WITHOUT its constant factor
// ImpedanceGeoAve (Z_geo): the callers multiply the accumulated sums by it once per particle
template <int RIEMANN> __device__ __forceinline__ float pjump_over_z(const FArgs &a, float u)
{
    if (RIEMANN == 0) return 0.f;
#if SPH_TRIM & 2
    if (RIEMANN == 2) return u;
    return u * fminf(a.lim_k * fmaxf(u, 0.f), 1.f);
#else
    float lim = RIEMANN == 1 ? fminf(a.limiter * (a.inv_c_ave * fmaxf(u, 0.f)), 1.f) : 1.f;
    return a.Z_geo * u * lim; // the complete DissipativePJump: the callers' factor is 1 then
#endif
}

// =====================================================================================================
// simple per-particle dynamics (run on the active slot range of the view: pointers are shifted by `b`)
// =====================================================================================================
struct Range
{
    u32 b, n;
};
static inline Range active_range(const sphb200_fluid_view_t *f)
{
    Range r;
    r.b = f->active_end ? f->active_begin : 0u;
    u32 e = f->active_end ? f->active_end : f->n;
    r.n = e > r.b ? e - r.b : 0u;
    return r;
}
__global__ void __launch_bounds__(256)
    k_gravity(u32 n, const float *__restrict__ mass, float4 *__restrict__ force_prior, float4 *__restrict__ prev, float gx,
              float gy, float gz)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = mass[i];
    float4 cur = make_float4(m * gx, m * gy, m * gz, 0.f);
    float4 fp = force_prior[i], pv = prev[i];
    fp.x += cur.x - pv.x; fp.y += cur.y - pv.y; fp.z += cur.z - pv.z;
    force_prior[i] = fp;
    prev[i] = cur;
}
extern "C" int sphb200_gravity_force(sphb200_context_t *ctx, const sphb200_fluid_view_t *f, const float g[3],
                                     sphb200_vec4_t *previous_force, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && g && previous_force && f->mass && f->force_prior, "null pointer");
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_gravity, sph_blocks(r.n, 256), 256, 0, stream, r.n, f->mass + r.b, (float4 *)f->force_prior + r.b,
                   (float4 *)previous_force + r.b, g[0], g[1], g[2]);
    return 0;
}

__global__ void __launch_bounds__(256)
    k_advection_setup(u32 n, const float *__restrict__ mass, const float *__restrict__ rho, float *__restrict__ vol,
                      float4 *__restrict__ dpos)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    vol[i] = __fdiv_rn(mass[i], rho[i]);
    dpos[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
extern "C" int sphb200_advection_setup(sphb200_context_t *ctx, const sphb200_fluid_view_t *f, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && f->mass && f->rho && f->vol && f->dpos, "null pointer");
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_advection_setup, sph_blocks(r.n, 256), 256, 0, stream, r.n, f->mass + r.b, f->rho + r.b, f->vol + r.b,
                   (float4 *)f->dpos + r.b);
    return 0;
}

__global__ void __launch_bounds__(256) k_update_position(u32 n, float4 *__restrict__ pos, const float4 *__restrict__ dpos)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 x = pos[i], d = dpos[i];
    x.x = __fadd_rn(x.x, d.x); x.y = __fadd_rn(x.y, d.y); x.z = __fadd_rn(x.z, d.z);
    pos[i] = x;
}
extern "C" int sphb200_update_position(sphb200_context_t *ctx, const sphb200_fluid_view_t *f, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && f->pos && f->dpos, "null pointer");
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_update_position, sph_blocks(r.n, 256), 256, 0, stream, r.n, (float4 *)f->pos + r.b,
                   (const float4 *)f->dpos + r.b);
    return 0;
}

// =====================================================================================================
// reductions
// =====================================================================================================
__device__ __forceinline__ float norm2_rn(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
__device__ __forceinline__ void block_max_to_global(float v, float *out)
{
    __shared__ float ws[32];
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        float t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
        t = warp_max(t);
        // all reduced quantities are >= 0, for which the unsigned bit pattern is order preserving
        if (threadIdx.x == 0) atomicMax((unsigned *)out, __float_as_uint(t));
    }
}

// AdvectionTimeStepCK::ReduceKernel::reduce = |v|^2, fluid_time_step_ck.h:106-109
__global__ void __launch_bounds__(256) k_reduce_advection(u32 n, const float4 *__restrict__ vel, float *out)
{
    float v = 0.f;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float4 u = vel[i];
        v = fmaxf(v, norm2_rn(u.x, u.y, u.z));
    }
    block_max_to_global(v, out);
}
// AcousticTimeStepCK::ReduceKernel::reduce, fluid_time_step_ck.hpp:51-57
__device__ __forceinline__ float acoustic_measure(float4 v, float4 F, float4 Fp, float m, float c0, float h_min)
{
    float fn = sqrtf(norm2_rn(__fadd_rn(F.x, Fp.x), __fadd_rn(F.y, Fp.y), __fadd_rn(F.z, Fp.z)));
    float acc = sqrtf(__fdiv_rn(__fmul_rn(__fmul_rn(4.0f, h_min), fn), m));
    float sp = __fadd_rn(c0, sqrtf(norm2_rn(v.x, v.y, v.z)));
    return fmaxf(sp, acc);
}
__global__ void __launch_bounds__(256)
    k_reduce_acoustic(u32 n, const float4 *__restrict__ vel, const float4 *__restrict__ force,
                      const float4 *__restrict__ force_prior, const float *__restrict__ mass, float c0, float h_min, float *out)
{
    float v = 0.f;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        v = fmaxf(v, acoustic_measure(vel[i], force[i], force_prior[i], mass[i], c0, h_min));
    block_max_to_global(v, out);
}
// legacy AcousticTimeStep::reduce = c0 + |v| (fluid_time_step.cpp:21-36)
__global__ void __launch_bounds__(256) k_reduce_acoustic_legacy(u32 n, const float4 *__restrict__ vel, float c0, float *out)
{
    float v = 0.f;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float4 u = vel[i];
        v = fmaxf(v, __fadd_rn(c0, sqrtf(norm2_rn(u.x, u.y, u.z))));
    }
    block_max_to_global(v, out);
}
// legacy AdvectionViscousTimeStep / AdvectionTimeStep::reduce = max(|v|^2, 4 h |F + F_prior| / m) (fluid_time_step.cpp:38-59)
__global__ void __launch_bounds__(256)
    k_reduce_advection_legacy(u32 n, const float4 *__restrict__ vel, const float4 *__restrict__ force,
                              const float4 *__restrict__ force_prior, const float *__restrict__ mass, float h_min, float *out)
{
    float v = 0.f;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float4 u = vel[i], F = force[i], Fp = force_prior[i];
        float fn = sqrtf(norm2_rn(__fadd_rn(F.x, Fp.x), __fadd_rn(F.y, Fp.y), __fadd_rn(F.z, Fp.z)));
        float acc = __fdiv_rn(__fmul_rn(__fmul_rn(4.0f, h_min), fn), mass[i]);
        v = fmaxf(v, fmaxf(norm2_rn(u.x, u.y, u.z), acc));
    }
    block_max_to_global(v, out);
}

static int read_scalar(sphb200_context *ctx, float *host, cudaStream_t st)
{
    SPH_CUDA(ctx, cudaMemcpyAsync(ctx->host_pinned, ctx->dev_scalars, sizeof(float), cudaMemcpyDeviceToHost, st));
    SPH_CUDA(ctx, cudaStreamSynchronize(st));
    *host = *(float *)ctx->host_pinned;
    return 0;
}

extern "C" int sphb200_advection_time_step(sphb200_context_t *ctx, const sphb200_fluid_view_t *f, float h_min, float u_ref,
                                           float cfl, float *reduced_host, float *dt_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && f->vel, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPH_CUDA(ctx, cudaMemsetAsync(ctx->dev_scalars, 0, sizeof(float), st));
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_reduce_advection, min(sph_blocks(r.n, 256), 148u * 8u), 256, 0, st, r.n, (const float4 *)f->vel + r.b,
                   (float *)ctx->dev_scalars);
    float red;
    int rc = read_scalar(ctx, &red, st);
    if (rc) return rc;
    if (reduced_host) *reduced_host = red;
    // FinishDynamics::Result, fluid_time_step_ck.cpp:24-27
    if (dt_host) *dt_host = cfl * h_min / (fmaxf(sqrtf(red), u_ref) + 2.71051e-20f);
    return 0;
}

extern "C" int sphb200_advection_time_step_legacy(sphb200_context_t *ctx, const sphb200_fluid_view_t *f, float h_min, float u_ref,
                                                  float cfl, float *reduced_host, float *dt_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && f->vel && f->force && f->force_prior && f->mass, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPH_CUDA(ctx, cudaMemsetAsync(ctx->dev_scalars, 0, sizeof(float), st));
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_reduce_advection_legacy, min(sph_blocks(r.n, 256), 148u * 8u), 256, 0, st, r.n, (const float4 *)f->vel + r.b,
                   (const float4 *)f->force + r.b, (const float4 *)f->force_prior + r.b, f->mass + r.b, h_min,
                   (float *)ctx->dev_scalars);
    float red;
    int rc = read_scalar(ctx, &red, st);
    if (rc) return rc;
    if (reduced_host) *reduced_host = red;
    if (dt_host) *dt_host = cfl * h_min / (fmaxf(sqrtf(red), u_ref) + 2.71051e-20f);
    return 0;
}

extern "C" int sphb200_acoustic_time_step(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float h_min, float cfl,
                                          float *reduced_host, float *dt_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s && s->fluid.vel && s->fluid.force && s->fluid.force_prior && s->fluid.mass, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const sphb200_fluid_view_t &f = s->fluid;
    SPH_CUDA(ctx, cudaMemsetAsync(ctx->dev_scalars, 0, sizeof(float), st));
    Range r = active_range(&f);
    if (r.n && s->material.formulation == 1)
        SPH_LAUNCH(ctx, k_reduce_acoustic_legacy, min(sph_blocks(r.n, 256), 148u * 8u), 256, 0, st, r.n, (const float4 *)f.vel + r.b,
                   s->material.c0, (float *)ctx->dev_scalars);
    else if (r.n)
        SPH_LAUNCH(ctx, k_reduce_acoustic, min(sph_blocks(r.n, 256), 148u * 8u), 256, 0, st, r.n, (const float4 *)f.vel + r.b,
                   (const float4 *)f.force + r.b, (const float4 *)f.force_prior + r.b, f.mass + r.b, s->material.c0, h_min,
                   (float *)ctx->dev_scalars);
    float red;
    int rc = read_scalar(ctx, &red, st);
    if (rc) return rc;
    if (reduced_host) *reduced_host = red;
    // FinishDynamics::Result, fluid_time_step_ck.hpp:31-36
    if (dt_host) *dt_host = cfl * h_min / (red + 2.71051e-20f);
    return 0;
}

__global__ void __launch_bounds__(256)
    k_energy(u32 n, const float4 *__restrict__ pos, const float4 *__restrict__ vel, const float *__restrict__ mass, float gx,
             float gy, float gz, double *out)
{
    double s = 0.0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        float4 x = pos[i], v = vel[i];
        float m = mass[i];
        // 0.5 m |v|^2 + m g.(0 - x); general_reduce_ck.h:52-88, external_force.h:53-56
        float ke = 0.5f * m * (v.x * v.x + v.y * v.y + v.z * v.z);
        float pe = m * (gx * (0.f - x.x) + gy * (0.f - x.y) + gz * (0.f - x.z));
        s += (double)ke + (double)pe;
    }
    __shared__ double ws[32];
    s = warp_sum_f64(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        double t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
        t = warp_sum_f64(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}
extern "C" int sphb200_total_mechanical_energy(sphb200_context_t *ctx, const sphb200_fluid_view_t *f, const float g[3],
                                               double *energy_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && g && energy_host && f->pos && f->vel && f->mass, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPH_CUDA(ctx, cudaMemsetAsync(ctx->dev_scalars, 0, sizeof(double), st));
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_energy, min(sph_blocks(r.n, 256), 148u * 4u), 256, 0, st, r.n, (const float4 *)f->pos + r.b,
                   (const float4 *)f->vel + r.b, f->mass + r.b, g[0], g[1], g[2], (double *)ctx->dev_scalars);
    SPH_CUDA(ctx, cudaMemcpyAsync(ctx->host_pinned, ctx->dev_scalars, sizeof(double), cudaMemcpyDeviceToHost, st));
    SPH_CUDA(ctx, cudaStreamSynchronize(st));
    *energy_host = *(double *)ctx->host_pinned;
    return 0;
}

// =====================================================================================================
// compression (density) summation + regularisation
// =====================================================================================================
template <bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS) k_compression_summation(FArgs a, KTab wtab, int regularize)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(wtab, tab);
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    float4 xi = a.pos[i];
    float s = a.legacy ? a.W0 : a.W0 * a.vol_ref[i];
    {
        u32 cnt = a.in_count[t];
        const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
        float4 xj[SUM_U];
        const bool legacy = a.legacy != 0;
        for_neighbors<SUM_U>(
            idx, cnt, [&](int u, u32 j) { xj[u] = a.posvolref[j]; },
            [&](int u, bool valid) {
                float dx = xi.x - xj[u].x, dy = xi.y - xj[u].y, dz = xi.z - xj[u].z;
                float r, inv_r;
                dist(dx * dx + dy * dy + dz * dz, r, inv_r);
                float w = kernel_w<ANALYTIC>(a, tab, r) * (legacy ? 1.0f : xj[u].w);
                s += valid ? w : 0.f;
            });
    }
    float sw = 0.f; // legacy: wall part is weighted separately (density_summation.cpp:58-78)
    if (a.n_wall)
    {
        u32 cnt = a.ct_count[t];
        const u32 *idx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
        float4 xj[NB_U];
        const bool legacy = a.legacy != 0;
        for_neighbors(
            idx, cnt, [&](int u, u32 j) { xj[u] = a.w_posvolref[j]; },
            [&](int u, bool valid) {
                float dx = xi.x - xj[u].x, dy = xi.y - xj[u].y, dz = xi.z - xj[u].z;
                float r, inv_r;
                dist(dx * dx + dy * dy + dz * dz, r, inv_r);
                float w = kernel_w<ANALYTIC>(a, tab, r) * xj[u].w;
                w = valid ? w : 0.f;
                if (legacy) sw += w; else s += w;
            });
    }
    if (a.legacy)
    {
        // DensitySummation<Inner<FreeSurface>> + <Contact<>>: rho_sum = sigma rho0/sigma0 + (sum_wall W m_j/rho0_wall) rho0^2/(sigma0 m_i)
        float rs = s * a.lg_sum_scale + sw * a.lg_wall_scale / a.mass[i];
        a.Csum[i] = rs;
        a.rho[i] = a.free_surface ? fmaxf(rs, a.rho0) : rs;
        return;
    }
    a.Csum[i] = s;
    if (regularize)
    {
        float C = a.free_surface ? fmaxf(s, 1.0f) : s; // Regularization<FreeSurface>, density_regularization.h:163-181
        a.C[i] = C;
        a.rho[i] = C * a.rho0;
    }
}

extern "C" int sphb200_compression_summation(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, int regularize, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s, "null pointer");
    FArgs a;
    KTab wtab;
    int rc = make_fargs(ctx, s, &a, &wtab, nullptr);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, a.n == 0 || (a.pos && a.posvolref && (a.vol_ref || a.legacy) && a.Csum && a.in_count && a.in_slice && a.in_index), "null fluid array");
    SPH_CHECK_ARG(ctx, a.legacy ? (a.rho && a.mass) : (!regularize || (a.C && a.rho)), "null fluid array");
    SPH_CHECK_ARG(ctx, a.n_wall == 0 || (a.w_posvolref && a.ct_count && a.ct_slice && a.ct_index), "null wall array");
    if (a.end > a.begin)
    {
        if (a.analytic) SPH_LAUNCH(ctx, k_compression_summation<true>, active_blocks(a, FL_THREADS), FL_THREADS, 0, stream, a, wtab, regularize);
        else SPH_LAUNCH(ctx, k_compression_summation<false>, active_blocks(a, FL_THREADS), FL_THREADS, 0, stream, a, wtab, regularize);
    }
    return 0;
}

__global__ void __launch_bounds__(256)
    k_density_regularization(u32 n, const float *__restrict__ Csum, float *__restrict__ C, float *__restrict__ rho, float rho0,
                             int free_surface)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = Csum[i];
    float c = free_surface ? fmaxf(s, 1.0f) : s;
    C[i] = c;
    rho[i] = c * rho0;
}
extern "C" int sphb200_density_regularization(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s && s->fluid.compression_sum && s->fluid.compression && s->fluid.rho, "null pointer");
    const sphb200_fluid_view_t &f = s->fluid;
    Range r = active_range(&f);
    if (r.n)
        SPH_LAUNCH(ctx, k_density_regularization, sph_blocks(r.n, 256), 256, 0, stream, r.n, f.compression_sum + r.b,
                   f.compression + r.b, f.rho + r.b, s->material.rho0, s->material.free_surface);
    return 0;
}

// =====================================================================================================
// acoustic step, 1st half
// =====================================================================================================
// InitializeKernel::initialize, acoustic_step_1st_half.hpp:66-74. Separate launch: interact reads p_j of
// neighbours, which must all be post-initialize.
__global__ void __launch_bounds__(256) k_a1_init(FArgs a, float dt)
{
    u32 i = a.begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.end) return;
    float rho;
    if (a.legacy) // Integration1stHalf::initialization, fluid_integration.hpp:49-56
        rho = a.rho[i] + a.Cdot[i] * dt * 0.5f;
    else
    {
        float C = a.C[i] + 0.5f * dt * a.Cdot[i];
        rho = C * a.rho0;
        a.C[i] = C;
    }
    a.rho[i] = rho;
    const float p = a.p0 * (rho / a.rho0 - 1.0f);
    a.p[i] = p;
    if (a.p_in_brec) a.brec[2ull * i + 1].z = p; // the neighbours of the correction variants read p here (k_a1_interact)
    float4 d = a.dpos[i], v = a.vel[i];
    d.x += v.x * dt * 0.5f; d.y += v.y * dt * 0.5f; d.z += v.z * dt * 0.5f;
    a.dpos[i] = d;
}

// InteractKernel::interact (inner, :89-111; wall, :157-180) + UpdateKernel::update (:122-127), one launch:
// update touches only particle i's own velocity, which no neighbour reads in this half step.
template <int RIEMANN, bool CORR, bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS, FL_MIN_BLOCKS) k_a1_interact(FArgs a, KTab dwtab, float dt, int do_update)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(dwtab, tab);
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    const float4 xi = a.posvol[i];
    const float p_i = a.p[i];
    float Bi[9];
    if (CORR) load_mat(a.B, i, Bi);

    float fx = 0.f, fy = 0.f, fz = 0.f, diss = 0.f;
    {
        u32 cnt = a.in_count[t];
        const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_U];
        float pjs[NB_U];
        if (CORR)
        {
            // sum_j dWV (p_i B_j + p_j B_i) e = p_i sum_j dWV B_j e + B_i sum_j p_j dWV e: B_i leaves the pair loop, and B_j
            // comes as ONE 32-byte record of its six distinct entries (B is the regularised inverse of the symmetric
            // sum_j r (x) r dW V / |r|: symmetric up to rounding; the record holds (B + B^T) / 2) instead of nine 4-byte gathers;
            // p_j rides in the record's seventh float (k_a1_init keeps it current): two gathers per pair, 48 bytes
            float4 bas[NB_U], bbs[NB_U];
            float ax = 0.f, ay = 0.f, az = 0.f; // sum_j dWV B_j e
            for_neighbors(
                idx, cnt,
                [&](int u, u32 j) {
                    xjs[u] = a.posvol[j];
                    load_rec2(a.brec, j, bas[u], bbs[u]);
                },
                [&](int u, bool valid) {
                    const float4 xj = xjs[u], ba = bas[u], bb = bbs[u];
                    const float p_j = bb.z; // written by k_a1_init of this acoustic step (and refreshed on ghost planes with the record)
                    float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                    float r2 = dx * dx + dy * dy + dz * dz;
                    float r, inv_r;
                    dist(r2, r, inv_r);
                    float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                    dWV = valid ? dWV : 0.f;
                    const float g = dWV * inv_r; // dWV e = g d
                    const float gx = g * dx, gy = g * dy, gz = g * dz;
                    ax += ba.x * gx + ba.y * gy + ba.z * gz;
                    ay += ba.y * gx + ba.w * gy + bb.x * gz;
                    az += ba.z * gx + bb.x * gy + bb.y * gz;
                    fx += p_j * gx; fy += p_j * gy; fz += p_j * gz; // sum_j p_j dWV e (B_i applied below)
#if SPH_TRIM & 4
                    if (RIEMANN) diss += (p_i - p_j) * dWV;
#else
                    if (RIEMANN) diss += (p_i - p_j) * a.inv_Z_ave * dWV;
#endif
                });
            const float3 bs = mat_vec(Bi, make_float3(fx, fy, fz));
            fx = -(p_i * ax + bs.x); fy = -(p_i * ay + bs.y); fz = -(p_i * az + bs.z);
        }
        else
        {
        for_neighbors(
            idx, cnt,
            [&](int u, u32 j) {
                xjs[u] = a.posvol[j];
                pjs[u] = a.p[j];
            },
            [&](int u, bool valid) {
                const float4 xj = xjs[u];
                const float p_j = pjs[u];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r2 = dx * dx + dy * dy + dz * dz;
                float r, inv_r;
                dist(r2, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                // AverageP (riemann_solver_ck.hpp:19-24) with Z_i == Z_j (one fluid): 2 * pave = p_i + p_j
                float c = (p_i + p_j) * dWV * inv_r;
                fx -= c * dx; fy -= c * dy; fz -= c * dz;
#if SPH_TRIM & 4
                if (RIEMANN) diss += (p_i - p_j) * dWV; // DissipativeUJump, :51-56 (its constant InvImpedanceAve: once per particle below)
#else
                if (RIEMANN) diss += (p_i - p_j) * a.inv_Z_ave * dWV; // DissipativeUJump, :51-56
#endif
            },
            [&](u32 j) { prefetch_record(a.posvol + j); });
        }
    }
    float wx = 0.f, wy = 0.f, wz = 0.f, wdiss = 0.f;
    const float vol_i = xi.w;
    float4 Fp = a.force_prior[i];
    const float m_i = a.mass[i];
    if (a.n_wall)
    {
        u32 cnt = a.ct_count[t];
        if (cnt)
        {
            const float rho_i = a.rho[i];
            const float ax = Fp.x / m_i, ay = Fp.y / m_i, az = Fp.z / m_i;
            const u32 *idx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
            float4 xjs[NB_WALL_U], was[NB_WALL_U];
            const bool has_acc = a.w_acc != nullptr;
            for_neighbors<NB_WALL_U>(
                idx, cnt,
                [&](int q, u32 j) {
                    xjs[q] = a.w_posvol[j];
                    if (has_acc) was[q] = a.w_acc[j];
                },
                [&](int q, bool valid) {
                const float4 xj = xjs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r2 = dx * dx + dy * dy + dz * dz;
                float r, inv_r;
                dist(r2, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                float ex = dx * inv_r, ey = dy * inv_r, ez = dz * inv_r;
                float rx = ax, ry = ay, rz = az;
                if (has_acc)
                {
                    float4 wa = was[q];
                    rx -= wa.x; ry -= wa.y; rz -= wa.z;
                }
                float face_acc = -(rx * ex + ry * ey + rz * ez);
                float p_w = p_i + rho_i * r * fmaxf(0.f, face_acc);
                float c = (p_i + p_w) * dWV;
                if (CORR)
                {
                    float3 be = mat_vec(Bi, make_float3(ex, ey, ez));
                    wx -= c * be.x; wy -= c * be.y; wz -= c * be.z;
                }
                else
                {
                    wx -= c * ex; wy -= c * ey; wz -= c * ez;
                }
#if SPH_TRIM & 4
                if (RIEMANN) wdiss += (p_i - p_w) * dWV;
#else
                if (RIEMANN) wdiss += (p_i - p_w) * a.inv_Z_ave * dWV;
#endif
                });
        }
    }
    float4 F = a.force[i];
    F.x += fx * vol_i; F.y += fy * vol_i; F.z += fz * vol_i;
    F.x += wx * vol_i; F.y += wy * vol_i; F.z += wz * vol_i;
    a.force[i] = F;
    const float C_i = a.legacy ? a.rho[i] : a.C[i];
#if SPH_TRIM & 4
    float cd = (diss * a.inv_Z_ave) * C_i;
    cd += (wdiss * a.inv_Z_ave) * C_i;
#else
    float cd = diss * C_i;
    cd += wdiss * C_i;
#endif
    a.Cdot[i] = cd;
    if (do_update)
    {
        float4 v = a.vel[i];
        v.x += (Fp.x + F.x) / m_i * dt;
        v.y += (Fp.y + F.y) / m_i * dt;
        v.z += (Fp.z + F.z) / m_i * dt;
        a.vel[i] = v;
        if (a.rec2) a.rec2[2ull * i + 1] = v; // keep the 2nd-half gather record current
    }
}

static int check_acoustic_args(sphb200_context *ctx, const FArgs &a, bool second)
{
    SPH_CHECK_ARG(ctx, a.n == 0 || (a.posvol && a.vel && a.dpos && a.force && a.force_prior && a.mass && a.rho && a.p && (a.C || a.legacy) &&
                                    a.Cdot && a.in_count && a.in_slice && a.in_index),
                  "null fluid array");
    SPH_CHECK_ARG(ctx, a.n_wall == 0 || (a.w_posvol && a.ct_count && a.ct_slice && a.ct_index), "null wall array");
    SPH_CHECK_ARG(ctx, !second || a.n_wall == 0 || a.w_n, "null wall normal");
    SPH_CHECK_ARG(ctx, !second || a.n == 0 || a.rec2, "null posvolvel record array");
    return 0;
}

extern "C" int sphb200_acoustic_1st_half_initialize(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float dt, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s, "null pointer");
    FArgs a;
    int rc = make_fargs(ctx, s, &a, nullptr, nullptr);
    if (rc) return rc;
    rc = check_acoustic_args(ctx, a, false);
    if (rc) return rc;
    if (a.end > a.begin) SPH_LAUNCH(ctx, k_a1_init, sph_blocks(a.end - a.begin, 256), 256, 0, stream, a, dt);
    return 0;
}

template <bool CORR, bool ANALYTIC>
static int launch_a1_k(sphb200_context *ctx, const FArgs &a, const KTab &t, int riemann, float dt, int upd, void *stream)
{
    unsigned g = active_blocks(a, FL_THREADS);
    switch (riemann)
    {
    case 0: SPH_LAUNCH(ctx, (k_a1_interact<0, CORR, ANALYTIC>), g, FL_THREADS, 0, stream, a, t, dt, upd); break;
    case 1: SPH_LAUNCH(ctx, (k_a1_interact<1, CORR, ANALYTIC>), g, FL_THREADS, 0, stream, a, t, dt, upd); break;
    default: SPH_LAUNCH(ctx, (k_a1_interact<2, CORR, ANALYTIC>), g, FL_THREADS, 0, stream, a, t, dt, upd); break;
    }
    return 0;
}
template <bool CORR> static int launch_a1(sphb200_context *ctx, const FArgs &a, const KTab &t, int riemann, float dt, int upd, void *stream)
{
    return a.analytic ? launch_a1_k<CORR, true>(ctx, a, t, riemann, dt, upd, stream) : launch_a1_k<CORR, false>(ctx, a, t, riemann, dt, upd, stream);
}

extern "C" int sphb200_acoustic_1st_half_interact(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float dt, int do_update,
                                                  void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s, "null pointer");
    FArgs a;
    KTab dwtab;
    int rc = make_fargs(ctx, s, &a, nullptr, &dwtab);
    if (rc) return rc;
    rc = check_acoustic_args(ctx, a, false);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, !s->material.correction || (a.B && a.brec), "LinearCorrectionCK needs fluid.B and fluid.correction_record");
    if (a.end <= a.begin) return 0;
    return s->material.correction ? launch_a1<true>(ctx, a, dwtab, s->material.riemann, dt, do_update, stream)
                                  : launch_a1<false>(ctx, a, dwtab, s->material.riemann, dt, do_update, stream);
}

extern "C" int sphb200_acoustic_1st_half(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float dt, void *stream)
{
    int rc = sphb200_acoustic_1st_half_initialize(ctx, s, dt, stream);
    if (rc) return rc;
    return sphb200_acoustic_1st_half_interact(ctx, s, dt, 1, stream);
}

// =====================================================================================================
// acoustic step, 2nd half — initialize + interact(inner) + interact(wall) + update in ONE launch: neighbours
// only read Position/Vol/Velocity here, none of which this half step writes.
// =====================================================================================================
template <int RIEMANN, bool CORR, bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS, FL_MIN_BLOCKS) k_a2(FArgs a, KTab dwtab, float dt, float h_min, float *next_reduced)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(dwtab, tab);
    u32 t = active_slot(a);
    float measure = 0.f;
    if (t >= a.begin && t < a.end)
    {
        const u32 i = a.order ? a.order[t] : t;
        const float4 xi = a.posvol[i];
        const float4 vi = a.vel[i];
        // InitializeKernel::initialize, acoustic_step_2nd_half.hpp:33-38
        {
            float4 d = a.dpos[i];
            d.x += vi.x * dt * 0.5f; d.y += vi.y * dt * 0.5f; d.z += vi.z * dt * 0.5f;
            a.dpos[i] = d;
        }
        float Bi[9];
        if (CORR) load_mat(a.B, i, Bi);
        float div = 0.f, px = 0.f, py = 0.f, pz = 0.f;
        {
            u32 cnt = a.in_count[t];
            const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
            float4 xjs[NB_U], vjs[NB_U];
            for_neighbors(
                idx, cnt, [&](int q, u32 j) { load_rec2(a.rec2, j, xjs[q], vjs[q]); },
                [&](int q, bool valid) {
                const float4 xj = xjs[q], vj = vjs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r2 = dx * dx + dy * dy + dz * dz;
                float r, inv_r;
                dist(r2, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                // AverageV (riemann_solver_ck.hpp:26-31) with Z_i == Z_j: 2 (v_i - v_ave) = v_i - v_j
                float ux = vi.x - vj.x, uy = vi.y - vj.y, uz = vi.z - vj.z;
#if SPH_TRIM & 2
                // u = (v_i - v_j) . e_ij with e_ij = d / |d| folded into the scalars (one multiply instead of three)
                float u = (ux * dx + uy * dy + uz * dz) * inv_r;
                if (CORR)
                {
                    float3 ce = mat_vec(Bi, make_float3(dx * inv_r, dy * inv_r, dz * inv_r));
                    div += (ux * ce.x + uy * ce.y + uz * ce.z) * dWV;
                }
                else
                    div += u * dWV;
                float c = pjump_over_z<RIEMANN>(a, u) * (dWV * inv_r);
                px += c * dx; py += c * dy; pz += c * dz;
#else
                float ex = dx * inv_r, ey = dy * inv_r, ez = dz * inv_r;
                float u = ux * ex + uy * ey + uz * ez;
                if (CORR)
                {
                    float3 ce = mat_vec(Bi, make_float3(ex, ey, ez));
                    div += (ux * ce.x + uy * ce.y + uz * ce.z) * dWV;
                }
                else
                    div += u * dWV;
                float c = pjump_over_z<RIEMANN>(a, u) * dWV;
                px += c * ex; py += c * ey; pz += c * ez;
#endif
                },
                [&](u32 j) { prefetch_record(a.rec2 + 2ull * j); });
        }
#if SPH_TRIM & 2
        px *= a.Z_geo; py *= a.Z_geo; pz *= a.Z_geo;
#endif
        float wdiv = 0.f, wx = 0.f, wy = 0.f, wz = 0.f;
        if (a.n_wall)
        {
            u32 cnt = a.ct_count[t];
            const u32 *idx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
            float4 xjs[NB_WALL_U], njs[NB_WALL_U], wvs[NB_WALL_U];
            const bool has_vel = a.w_vel != nullptr;
            for_neighbors<NB_WALL_U>(
                idx, cnt,
                [&](int q, u32 j) {
                    xjs[q] = a.w_posvol[j];
                    njs[q] = a.w_n[j];
                    if (has_vel) wvs[q] = a.w_vel[j];
                },
                [&](int q, bool valid) {
                const float4 xj = xjs[q], nj = njs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r2 = dx * dx + dy * dy + dz * dz;
                float r, inv_r;
                dist(r2, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                float ex = dx * inv_r, ey = dy * inv_r, ez = dz * inv_r;
                float vx = vi.x, vy = vi.y, vz = vi.z;
                if (has_vel)
                {
                    float4 wv = wvs[q];
                    vx -= wv.x; vy -= wv.y; vz -= wv.z;
                }
                vx *= 2.0f; vy *= 2.0f; vz *= 2.0f; // vel_diff = 2 (v_i - v_wall)
                float3 ce = CORR ? mat_vec(Bi, make_float3(ex, ey, ez)) : make_float3(ex, ey, ez);
                wdiv += (vx * ce.x + vy * ce.y + vz * ce.z) * dWV;
                float en = ex * nj.x + ey * nj.y + ez * nj.z;
                float sg = en < 0.f ? -1.f : (en > 0.f ? 1.f : 0.f); // SGN, scalar_functions.h:128-131
                float nx = sg * nj.x, ny = sg * nj.y, nz = sg * nj.z;
                float u = vx * nx + vy * ny + vz * nz;
                float c = pjump_over_z<RIEMANN>(a, u) * dWV;
                wx += c * nx; wy += c * ny; wz += c * nz;
                });
#if SPH_TRIM & 2
            wx *= a.Z_geo; wy *= a.Z_geo; wz *= a.Z_geo;
#endif
        }
        const float vol_i = xi.w;
        float C = a.legacy ? a.rho[i] : a.C[i];
        float cd = a.Cdot[i];
        cd += div * C;
        cd += wdiv * C;
        a.Cdot[i] = cd;
        float4 F = make_float4(px * vol_i, py * vol_i, pz * vol_i, 0.f);
        F.x += wx * vol_i; F.y += wy * vol_i; F.z += wz * vol_i;
        a.force[i] = F;
        // UpdateKernel::update, acoustic_step_2nd_half.hpp:83-89
        if (a.legacy)
            a.rho[i] = C + cd * dt * 0.5f; // Integration2ndHalf::update, fluid_integration.hpp:172-176
        else
        {
            C += 0.5f * dt * cd;
            a.C[i] = C;
            a.rho[i] = C * a.rho0;
        }
        if (next_reduced)
            measure = a.legacy ? __fadd_rn(a.c0, sqrtf(norm2_rn(vi.x, vi.y, vi.z)))
                               : acoustic_measure(vi, F, a.force_prior[i], a.mass[i], a.c0, h_min);
    }
    if (next_reduced) block_max_to_global(measure, next_reduced);
}

template <bool CORR, bool ANALYTIC>
static int launch_a2_k(sphb200_context *ctx, const FArgs &a, const KTab &t, int riemann, float dt, float h_min, float *nr, void *stream)
{
    unsigned g = active_blocks(a, FL_THREADS);
    switch (riemann)
    {
    case 0: SPH_LAUNCH(ctx, (k_a2<0, CORR, ANALYTIC>), g, FL_THREADS, 0, stream, a, t, dt, h_min, nr); break;
    case 1: SPH_LAUNCH(ctx, (k_a2<1, CORR, ANALYTIC>), g, FL_THREADS, 0, stream, a, t, dt, h_min, nr); break;
    default: SPH_LAUNCH(ctx, (k_a2<2, CORR, ANALYTIC>), g, FL_THREADS, 0, stream, a, t, dt, h_min, nr); break;
    }
    return 0;
}
template <bool CORR>
static int launch_a2(sphb200_context *ctx, const FArgs &a, const KTab &t, int riemann, float dt, float h_min, float *nr, void *stream)
{
    return a.analytic ? launch_a2_k<CORR, true>(ctx, a, t, riemann, dt, h_min, nr, stream)
                      : launch_a2_k<CORR, false>(ctx, a, t, riemann, dt, h_min, nr, stream);
}

extern "C" int sphb200_acoustic_2nd_half(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float dt, float h_min,
                                         float *next_reduced_dev, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s, "null pointer");
    FArgs a;
    KTab dwtab;
    int rc = make_fargs(ctx, s, &a, nullptr, &dwtab);
    if (rc) return rc;
    rc = check_acoustic_args(ctx, a, true);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, !s->material.correction || a.B, "LinearCorrectionCK needs fluid.B");
    if (a.end <= a.begin) return 0;
    return s->material.correction ? launch_a2<true>(ctx, a, dwtab, s->material.riemann, dt, h_min, next_reduced_dev, stream)
                                  : launch_a2<false>(ctx, a, dwtab, s->material.riemann, dt, h_min, next_reduced_dev, stream);
}

// =====================================================================================================
// linear correction matrix: B_i = -(sum_j r_ij (x) nablaW_ij V_j), then the Tikhonov-regularised,
// determinant-weighted inverse; kernel_correction_ck.hpp:40-95, common/vector_functions.h:199-203
// =====================================================================================================
__device__ __forceinline__ float det3(const float *m)
{
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ __forceinline__ void inv3(const float *m, float *o)
{
    float d = det3(m);
    o[0] = (m[4] * m[8] - m[5] * m[7]) / d; o[1] = (m[2] * m[7] - m[1] * m[8]) / d; o[2] = (m[1] * m[5] - m[2] * m[4]) / d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) / d; o[4] = (m[0] * m[8] - m[2] * m[6]) / d; o[5] = (m[2] * m[3] - m[0] * m[5]) / d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) / d; o[7] = (m[1] * m[6] - m[0] * m[7]) / d; o[8] = (m[0] * m[4] - m[1] * m[3]) / d;
}

template <bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS) k_linear_correction(FArgs a, KTab dwtab, float alpha)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(dwtab, tab);
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    const float4 xi = a.posvol[i];
    float b[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) b[k] = 0.f;
    auto accumulate = [&](float4 xj, bool valid) {
        float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        float r2 = dx * dx + dy * dy + dz * dz;
        float r, inv_r;
        dist(r2, r, inv_r);
        float g = kernel_dw<ANALYTIC>(a, tab, r) * inv_r * xj.w; // dW/r * V_j : nablaW V = g * d
        g = valid ? g : 0.f;
        float gx = g * dx, gy = g * dy, gz = g * dz;
        b[0] -= dx * gx; b[1] -= dx * gy; b[2] -= dx * gz;
        b[3] -= dy * gx; b[4] -= dy * gy; b[5] -= dy * gz;
        b[6] -= dz * gx; b[7] -= dz * gy; b[8] -= dz * gz;
    };
    {
        u32 cnt = a.in_count[t];
        const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_U];
        for_neighbors(idx, cnt, [&](int u, u32 j) { xjs[u] = a.posvol[j]; }, [&](int u, bool valid) { accumulate(xjs[u], valid); });
    }
    if (a.n_wall)
    {
        u32 cnt = a.ct_count[t];
        const u32 *idx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_U];
        for_neighbors(idx, cnt, [&](int u, u32 j) { xjs[u] = a.w_posvol[j]; }, [&](int u, bool valid) { accumulate(xjs[u], valid); });
    }
    if (a.dim == 2) b[8] = 1.0f;
    float det = det3(b);
    float det_sqr = fmaxf(alpha - det, 0.f);
    float btb[9], bt[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) bt[3 * r + c] = b[3 * c + r];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            btb[3 * r + c] = bt[3 * r] * b[c] + bt[3 * r + 1] * b[3 + c] + bt[3 * r + 2] * b[6 + c];
    const float eps = 1.0e-8f; // SqrtEps, base_data_type.h:206
    btb[0] += eps; btb[4] += eps; btb[8] += eps;
    float ib[9], inv[9];
    inv3(btb, ib);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            inv[3 * r + c] = ib[3 * r] * bt[c] + ib[3 * r + 1] * bt[3 + c] + ib[3 * r + 2] * bt[6 + c];
    float wgt = det / (det + det_sqr);
    float B[9];
#pragma unroll
    for (int k = 0; k < 9; ++k)
    {
        float id = (k == 0 || k == 4 || k == 8) ? 1.0f : 0.f;
        B[k] = wgt * inv[k] + (1.0f - wgt) * id;
        a.B[9ull * i + k] = B[k];
    }
    if (a.brec) // the gather record of the 1st-half interaction (k_a1_interact): the symmetric part, six entries
    {
        a.brec[2ull * i] = make_float4(B[0], 0.5f * (B[1] + B[3]), 0.5f * (B[2] + B[6]), B[4]);
        a.brec[2ull * i + 1] = make_float4(0.5f * (B[5] + B[7]), B[8], a.p ? a.p[i] : 0.f, 0.f);
    }
}

// the gather record of given matrices (matrices written outside the library, e.g. uploaded by the host)
__global__ void __launch_bounds__(256)
    k_pack_correction_records(u32 n, const float *__restrict__ B, const float *__restrict__ p, float4 *__restrict__ rec)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (B)
    {
        float b[9];
        load_mat(B, i, b);
        rec[2ull * i] = make_float4(b[0], 0.5f * (b[1] + b[3]), 0.5f * (b[2] + b[6]), b[4]);
        rec[2ull * i + 1].x = 0.5f * (b[5] + b[7]);
        rec[2ull * i + 1].y = b[8];
    }
    if (p) rec[2ull * i + 1].z = p[i];
}
extern "C" int sphb200_pack_correction_records(sphb200_context_t *ctx, uint32_t n, const float *B, const float *pressure,
                                               void *correction_record, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (n == 0 || ((B || pressure) && correction_record)), "null pointer");
    if (n) SPH_LAUNCH(ctx, k_pack_correction_records, sph_blocks(n, 256), 256, 0, stream, n, B, pressure, (float4 *)correction_record);
    return 0;
}

extern "C" int sphb200_linear_correction_matrix(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float alpha, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s, "null pointer");
    FArgs a;
    KTab dwtab;
    int rc = make_fargs(ctx, s, &a, nullptr, &dwtab);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, a.n == 0 || (a.posvol && a.B && a.in_count && a.in_slice && a.in_index), "null fluid array");
    SPH_CHECK_ARG(ctx, a.n_wall == 0 || (a.w_posvol && a.ct_count && a.ct_slice && a.ct_index), "null wall array");
    if (a.end > a.begin)
    {
        if (a.analytic) SPH_LAUNCH(ctx, k_linear_correction<true>, active_blocks(a, FL_THREADS), FL_THREADS, 0, stream, a, dwtab, alpha);
        else SPH_LAUNCH(ctx, k_linear_correction<false>, active_blocks(a, FL_THREADS), FL_THREADS, 0, stream, a, dwtab, alpha);
    }
    return 0;
}

// =====================================================================================================
// free-surface indication: FreeSurfaceIndicationCK<Inner<WithUpdate>, Contact<>>
// general_dynamics/surface_indication/surface_indication_ck.hpp:52-160
// =====================================================================================================
// inner InteractKernel::interact (:52-70) + isNearPreviousFreeSurface (:72-87) + contact InteractKernel::interact (:149-160),
// one launch: the contact part only adds to the particle's own PositionDivergence.
template <bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS) k_surface_interact(FArgs a, KTab dwtab, const int *__restrict__ previous,
                                                                  float *__restrict__ pos_div, float threshold)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(dwtab, tab);
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    const float4 xi = a.posvol[i];
    const u32 cnt = a.in_count[t];
    const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
    float pd = 0.f;
    {
        float4 xjs[SUM_U];
        for_neighbors<SUM_U>(
            idx, cnt, [&](int u, u32 j) { xjs[u] = a.posvol[j]; },
            [&](int u, bool valid) {
                const float4 xj = xjs[u];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r, inv_r;
                dist(dx * dx + dy * dy + dz * dz, r, inv_r);
                const float w = kernel_dw<ANALYTIC>(a, tab, r) * xj.w * r;
                pd -= valid ? w : 0.f;
            });
    }
    if (pd < threshold && previous[i] != 1)
    {
        // only particles that newly look like surface particles pay for the second sweep
        bool near_previous = false;
        for (u32 k = 0; k < cnt && !near_previous; ++k) near_previous = previous[idx[32ull * k]] == 1;
        if (!near_previous) pd = 2.0f * threshold;
    }
    float pw = 0.f;
    if (a.n_wall)
    {
        u32 wc = a.ct_count[t];
        const u32 *widx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
        float4 xjs[SUM_U];
        for_neighbors<SUM_U>(
            widx, wc, [&](int u, u32 j) { xjs[u] = a.w_posvol[j]; },
            [&](int u, bool valid) {
                const float4 xj = xjs[u];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r, inv_r;
                dist(dx * dx + dy * dy + dz * dz, r, inv_r);
                const float w = kernel_dw<ANALYTIC>(a, tab, r) * xj.w * r;
                pw -= valid ? w : 0.f;
            });
    }
    pos_div[i] = pd + pw;
}

// UpdateKernel::update (:98-107) + isVeryNearFreeSurface (:109-128). Separate launch: reads PositionDivergence of neighbours.
__global__ void __launch_bounds__(FL_THREADS) k_surface_update(FArgs a, const float *__restrict__ pos_div, int *__restrict__ indicator,
                                                                int *__restrict__ previous, float threshold, float smoothing_length)
{
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    int ind = 1;
    if (pos_div[i] > threshold)
    {
        const float4 xi = a.posvol[i];
        const u32 cnt = a.in_count[t];
        const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
        // batched like the pair loops (for_neighbors): an interior particle reads PositionDivergence of ALL its neighbours and
        // finds none below the threshold — a loop with an early exit kept one dependent gather in flight per warp
        bool very_near = false;
        float pds[SUM_U];
        u32 js[SUM_U];
        for_neighbors<SUM_U>(
            idx, cnt,
            [&](int u, u32 j) {
                pds[u] = pos_div[j];
                js[u] = j;
            },
            [&](int u, bool valid) {
                if (valid && !very_near && pds[u] < threshold)
                {
                    float4 xj = a.posvol[js[u]];
                    float dx = __fsub_rn(xi.x, xj.x), dy = __fsub_rn(xi.y, xj.y), dz = __fsub_rn(xi.z, xj.z);
                    very_near = sqrtf(norm2_rn(dx, dy, dz)) < smoothing_length; // same rounding as the CPU evaluation
                }
            });
        if (!very_near) ind = 0;
    }
    indicator[i] = ind;
    previous[i] = ind;
}

static int surface_indication_sweeps(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, int32_t *indicator, float *position_divergence,
                                     int32_t *previous_indicator, float threshold, float smoothing_length, bool first, bool second,
                                     void *stream);

extern "C" int sphb200_free_surface_indication(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, int32_t *indicator,
                                               float *position_divergence, int32_t *previous_indicator, float threshold,
                                               float smoothing_length, void *stream)
{
    return surface_indication_sweeps(ctx, s, indicator, position_divergence, previous_indicator, threshold, smoothing_length, true, true, stream);
}

// One sweep of the indication (sweep 0: interact, writes PositionDivergence; sweep 1: update, reads PositionDivergence of
// the neighbours): slab-decomposed runs refresh PositionDivergence on the ghost planes between the two.
extern "C" int sphb200_free_surface_indication_sweep(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, int32_t *indicator,
                                                     float *position_divergence, int32_t *previous_indicator, float threshold,
                                                     float smoothing_length, int sweep, void *stream)
{
    SPH_CHECK_ARG(ctx, sweep == 0 || sweep == 1, "sweep must be 0 (interact) or 1 (update)");
    return surface_indication_sweeps(ctx, s, indicator, position_divergence, previous_indicator, threshold, smoothing_length, sweep == 0,
                                     sweep == 1, stream);
}

static int surface_indication_sweeps(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, int32_t *indicator, float *position_divergence,
                                     int32_t *previous_indicator, float threshold, float smoothing_length, bool first, bool second,
                                     void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s && indicator && position_divergence && previous_indicator, "null pointer");
    FArgs a;
    KTab dwtab;
    int rc = make_fargs(ctx, s, &a, nullptr, &dwtab);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, a.n == 0 || (a.posvol && a.in_count && a.in_slice && a.in_index), "null fluid array");
    SPH_CHECK_ARG(ctx, a.n_wall == 0 || (a.w_posvol && a.ct_count && a.ct_slice && a.ct_index), "null wall array");
    if (a.end <= a.begin) return 0;
    unsigned g = active_blocks(a, FL_THREADS);
    if (first)
    {
        if (a.analytic) SPH_LAUNCH(ctx, k_surface_interact<true>, g, FL_THREADS, 0, stream, a, dwtab, previous_indicator, position_divergence, threshold);
        else SPH_LAUNCH(ctx, k_surface_interact<false>, g, FL_THREADS, 0, stream, a, dwtab, previous_indicator, position_divergence, threshold);
    }
    if (second)
        SPH_LAUNCH(ctx, k_surface_update, g, FL_THREADS, 0, stream, a, position_divergence, indicator, previous_indicator, threshold,
                   smoothing_length);
    return 0;
}

// =====================================================================================================
// observation: Interpolation<Contact<DataType>>::InteractKernel::interact, general_dynamics/interpolation_dynamics.hpp:44-60
//   out_i = sum_j W_ij V_j data_j / (sum_j W_ij V_j + TinyReal)
// One thread per observer particle (observer bodies hold a handful of probes; their rows are in plain slot order).
// =====================================================================================================
template <int WIDTH, bool ANALYTIC>
__global__ void __launch_bounds__(128) k_interpolate(FArgs a, KTab wtab, const float4 *__restrict__ src_pos, u32 n_src,
                                                     const u32 *__restrict__ count, const u32 *__restrict__ slice,
                                                     const u32 *__restrict__ index, const float4 *__restrict__ tar_posvol,
                                                     const float *__restrict__ tar_data, float *__restrict__ out)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(wtab, tab);
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_src) return;
    const float4 xi = src_pos[t];
    const u32 cnt = count[t];
    const u32 *idx = index + (u64)slice[t >> 5] + (t & 31u);
    float acc[WIDTH];
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) acc[c] = 0.f;
    float total = 0.f;
    for (u32 k = 0; k < cnt; ++k)
    {
        u32 j = idx[32ull * k];
        float4 xj = tar_posvol[j];
        float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
        float r = sqrtf(dx * dx + dy * dy + dz * dz);
        float w = kernel_w<ANALYTIC>(a, tab, r) * xj.w;
#pragma unroll
        for (int c = 0; c < WIDTH; ++c) acc[c] += w * tar_data[(u64)WIDTH * j + c];
        total += w;
    }
    const float denom = total + 2.71051e-20f; // TinyReal, base_data_type.h:207
#pragma unroll
    for (int c = 0; c < WIDTH; ++c) out[(u64)WIDTH * t + c] = acc[c] / denom;
}

extern "C" int sphb200_interpolate(sphb200_context_t *ctx, const sphb200_kernel_t *kernel, const sphb200_vec4_t *src_pos,
                                   uint32_t n_src, sphb200_relation_t rel, const sphb200_vec4_t *tar_posvol, const float *tar_data,
                                   int width, float *out, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && kernel, "null pointer");
    SPH_CHECK_ARG(ctx, width == 1 || width == 4, "width must be 1 (Real) or 4 (Vecd stored as float4)");
    if (n_src == 0) return 0;
    SPH_CHECK_ARG(ctx, src_pos && rel.count && rel.slice_offset && rel.index && tar_posvol && tar_data && out, "null pointer");
    SPH_CHECK_ARG(ctx, rel.order == nullptr, "observer rows must be in plain slot order");
    sphb200_fluid_args_t s;
    memset(&s, 0, sizeof(s));
    s.kernel = *kernel;
    s.material.rho0 = 1.f;
    s.material.c0 = 1.f;
    FArgs a;
    KTab wtab;
    int rc = make_fargs(ctx, &s, &a, &wtab, nullptr);
    if (rc) return rc;
    unsigned g = sph_blocks(n_src, 128);
    const float4 *sp = (const float4 *)src_pos, *tp = (const float4 *)tar_posvol;
    if (width == 1)
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_interpolate<1, true>), g, 128, 0, stream, a, wtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
        else SPH_LAUNCH(ctx, (k_interpolate<1, false>), g, 128, 0, stream, a, wtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
    }
    else
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_interpolate<4, true>), g, 128, 0, stream, a, wtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
        else SPH_LAUNCH(ctx, (k_interpolate<4, false>), g, 128, 0, stream, a, wtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
    }
    return 0;
}

// Interpolation<Contact<DataType, RestoringCorrection>>::InteractKernel::interact, general_dynamics/interpolation_dynamics.hpp:72-100:
// the first-order consistent interpolation — reproduces constant and linear fields on ANY neighbour set whose restoring
// matrix is regular (one-sided neighbourhoods at walls and free surfaces included). Per observer, with n = dim + 1:
//   A_j(0,0) = W V, A_j(0,1+b) = -W V r_b, A_j(1+a,0) = dW V e_a, A_j(1+a,1+b) = -dW V r_a e_b
//   restoring = Eps I + sum_j A_j;  prediction = sum_j A_j.col(0) data_j;  out = (restoring^-1).row(0) . prediction
// Row 0 of the inverse by cofactors (what Eigen's fixed-size inverse evaluates): (M^-1)(0,k) = cofactor(k,0) / det.
template <int K> __device__ __forceinline__ float minor3_col0(const float (&M)[4][4]) // n == 3: rows != K of {0,1,2}, columns 1,2
{
    constexpr int r0 = K == 0 ? 1 : 0, r1 = K == 2 ? 1 : 2;
    return M[r0][1] * M[r1][2] - M[r0][2] * M[r1][1];
}
template <int K> __device__ __forceinline__ float minor4_col0(const float (&M)[4][4]) // n == 4: rows != K of {0,1,2,3}, columns 1,2,3
{
    constexpr int r0 = K == 0 ? 1 : 0, r1 = K <= 1 ? 2 : 1, r2 = K == 3 ? 2 : 3;
    return M[r0][1] * (M[r1][2] * M[r2][3] - M[r1][3] * M[r2][2]) - M[r0][2] * (M[r1][1] * M[r2][3] - M[r1][3] * M[r2][1]) +
           M[r0][3] * (M[r1][1] * M[r2][2] - M[r1][2] * M[r2][1]);
}
template <int WIDTH, bool ANALYTIC>
__global__ void __launch_bounds__(128) k_interpolate_restoring(FArgs a, KTab wtab, KTab dwtab, const float4 *__restrict__ src_pos, u32 n_src,
                                                               const u32 *__restrict__ count, const u32 *__restrict__ slice,
                                                               const u32 *__restrict__ index, const float4 *__restrict__ tar_posvol,
                                                               const float *__restrict__ tar_data, float *__restrict__ out)
{
    __shared__ float4 tab[KT_SLOTS], dtab[KT_SLOTS];
    if (!ANALYTIC)
    {
        stage_tab(wtab, tab);
        stage_tab(dwtab, dtab);
    }
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_src) return;
    const float4 xi = src_pos[t];
    const u32 cnt = count[t];
    const u32 *idx = index + (u64)slice[t >> 5] + (t & 31u);
    const int dim = a.dim;
    float M[4][4], pred[4][WIDTH];
#pragma unroll
    for (int r = 0; r < 4; ++r)
    {
#pragma unroll
        for (int c = 0; c < 4; ++c) M[r][c] = r == c ? 1.1920929e-07f : 0.f; // Eps, base_data_type.h:205
#pragma unroll
        for (int c = 0; c < WIDTH; ++c) pred[r][c] = 0.f;
    }
    for (u32 k = 0; k < cnt; ++k)
    {
        const u32 j = idx[32ull * k];
        const float4 xj = tar_posvol[j];
        const float rv[3] = {xi.x - xj.x, xi.y - xj.y, xi.z - xj.z};
        float r, inv_r;
        dist(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2], r, inv_r);
        const float WV = kernel_w<ANALYTIC>(a, tab, r) * xj.w, dWV = kernel_dw<ANALYTIC>(a, dtab, r) * xj.w;
        float col0[4];
        col0[0] = WV;
        M[0][0] += WV;
#pragma unroll
        for (int b = 0; b < 3; ++b)
            if (b < dim) M[0][1 + b] -= WV * rv[b];
#pragma unroll
        for (int q = 0; q < 3; ++q)
        {
            const float ge = q < dim ? dWV * (rv[q] * inv_r) : 0.f; // dW V e_q
            col0[1 + q] = ge;
            M[1 + q][0] += ge;
#pragma unroll
            for (int b = 0; b < 3; ++b)
                if (q < dim && b < dim) M[1 + q][1 + b] -= dWV * rv[q] * (rv[b] * inv_r);
        }
#pragma unroll
        for (int c = 0; c < WIDTH; ++c)
        {
            const float d = tar_data[(u64)WIDTH * j + c];
#pragma unroll
            for (int q = 0; q < 4; ++q) pred[q][c] += col0[q] * d;
        }
    }
    float cof[4];
    if (dim == 2)
    {
        cof[0] = minor3_col0<0>(M); cof[1] = -minor3_col0<1>(M); cof[2] = minor3_col0<2>(M); cof[3] = 0.f;
    }
    else
    {
        cof[0] = minor4_col0<0>(M); cof[1] = -minor4_col0<1>(M); cof[2] = minor4_col0<2>(M); cof[3] = -minor4_col0<3>(M);
    }
    const float det = M[0][0] * cof[0] + M[1][0] * cof[1] + M[2][0] * cof[2] + (dim == 3 ? M[3][0] * cof[3] : 0.f);
#pragma unroll
    for (int c = 0; c < WIDTH; ++c)
    {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) v += (cof[q] / det) * pred[q][c];
        out[(u64)WIDTH * t + c] = v;
    }
}

extern "C" int sphb200_interpolate_restoring(sphb200_context_t *ctx, const sphb200_kernel_t *kernel, const sphb200_vec4_t *src_pos,
                                             uint32_t n_src, sphb200_relation_t rel, const sphb200_vec4_t *tar_posvol,
                                             const float *tar_data, int width, float *out, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && kernel, "null pointer");
    SPH_CHECK_ARG(ctx, width == 1 || width == 4, "width must be 1 (Real) or 4 (Vecd stored as float4)");
    if (n_src == 0) return 0;
    SPH_CHECK_ARG(ctx, src_pos && rel.count && rel.slice_offset && rel.index && tar_posvol && tar_data && out, "null pointer");
    SPH_CHECK_ARG(ctx, rel.order == nullptr, "observer rows must be in plain slot order");
    sphb200_fluid_args_t s;
    memset(&s, 0, sizeof(s));
    s.kernel = *kernel;
    s.material.rho0 = 1.f;
    s.material.c0 = 1.f;
    FArgs a;
    KTab wtab, dwtab;
    int rc = make_fargs(ctx, &s, &a, &wtab, &dwtab);
    if (rc) return rc;
    unsigned g = sph_blocks(n_src, 128);
    const float4 *sp = (const float4 *)src_pos, *tp = (const float4 *)tar_posvol;
    if (width == 1)
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_interpolate_restoring<1, true>), g, 128, 0, stream, a, wtab, dwtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
        else SPH_LAUNCH(ctx, (k_interpolate_restoring<1, false>), g, 128, 0, stream, a, wtab, dwtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
    }
    else
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_interpolate_restoring<4, true>), g, 128, 0, stream, a, wtab, dwtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
        else SPH_LAUNCH(ctx, (k_interpolate_restoring<4, false>), g, 128, 0, stream, a, wtab, dwtab, sp, n_src, rel.count, rel.slice_offset, rel.index, tp, tar_data, out);
    }
    return 0;
}

// =====================================================================================================
// viscous force: ViscousForceCK<Inner<WithUpdate, Viscosity, Correction>, Contact<Wall, Viscosity, Correction>> + the
// ForcePriorCK update it carries; fluid_dynamics/viscous_force.hpp:44-103, general_dynamics/force_prior_ck.h:53-57
//   inner: F_i = V_i sum_j r_ij.((B_i + B_j) e_ij) mu_ij (v_i - v_j) / (r^2 + 0.01 h^2) dW_ij V_j
//   wall : F_i += V_i sum_j 2 r_ij.(B_i e_ij) mu 2 (v_i - v_wall,j) / (r^2 + 0.01 h^2) dW_ij V_j
//   then : ForcePrior += F - Previous; Previous = F          (one launch: only the particle's own data is written)
// =====================================================================================================
template <bool CORR, bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS, FL_MIN_BLOCKS)
    k_viscous_force(FArgs a, KTab dwtab, float mu, float h2_eps, float4 *__restrict__ viscous_force, float4 *__restrict__ previous_force)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(dwtab, tab);
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    const float4 xi = a.posvol[i];
    const float4 vi = a.vel[i];
    float Bi[9];
    if (CORR) load_mat(a.B, i, Bi);
    float fx = 0.f, fy = 0.f, fz = 0.f;
    {
        u32 cnt = a.in_count[t];
        const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_U], vjs[NB_U];
        u32 js[NB_U];
        for_neighbors<(CORR ? 2 : NB_U)>(
            idx, cnt,
            [&](int q, u32 j) {
                load_rec2(a.rec2, j, xjs[q], vjs[q]);
                if (CORR) js[q] = j;
            },
            [&](int q, bool valid) {
                const float4 xj = xjs[q], vj = vjs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r2 = dx * dx + dy * dy + dz * dz;
                float r, inv_r;
                dist(r2, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                float proj; // r_ij . ((B_i + B_j) e_ij)
                if (CORR)
                {
                    float Bj[9];
                    float4 ba, bb; // B_j as one 32-byte record of its symmetric part (see k_a1_interact)
                    load_rec2(a.brec, js[q], ba, bb);
                    Bj[0] = ba.x; Bj[1] = ba.y; Bj[2] = ba.z; Bj[3] = ba.y; Bj[4] = ba.w; Bj[5] = bb.x; Bj[6] = ba.z; Bj[7] = bb.x; Bj[8] = bb.y;
#pragma unroll
                    for (int k = 0; k < 9; ++k) Bj[k] += Bi[k];
                    float3 be = mat_vec(Bj, make_float3(dx * inv_r, dy * inv_r, dz * inv_r));
                    proj = dx * be.x + dy * be.y + dz * be.z;
                }
                else
                    proj = 2.0f * r; // (1 + 1) r_ij . e_ij
                float c = proj * mu * dWV / (r2 + h2_eps);
                fx += c * (vi.x - vj.x); fy += c * (vi.y - vj.y); fz += c * (vi.z - vj.z);
            });
    }
    float wx = 0.f, wy = 0.f, wz = 0.f;
    if (a.n_wall)
    {
        u32 cnt = a.ct_count[t];
        const u32 *idx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_WALL_U], wvs[NB_WALL_U];
        const bool has_vel = a.w_vel != nullptr;
        for_neighbors<NB_WALL_U>(
            idx, cnt,
            [&](int q, u32 j) {
                xjs[q] = a.w_posvol[j];
                if (has_vel) wvs[q] = a.w_vel[j];
            },
            [&](int q, bool valid) {
                const float4 xj = xjs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r2 = dx * dx + dy * dy + dz * dz;
                float r, inv_r;
                dist(r2, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                float proj;
                if (CORR)
                {
                    float3 be = mat_vec(Bi, make_float3(dx * inv_r, dy * inv_r, dz * inv_r));
                    proj = dx * be.x + dy * be.y + dz * be.z;
                }
                else
                    proj = r;
                float ux = vi.x, uy = vi.y, uz = vi.z;
                if (has_vel) { ux -= wvs[q].x; uy -= wvs[q].y; uz -= wvs[q].z; }
                float c = 2.0f * proj * mu * dWV / (r2 + h2_eps) * 2.0f;
                wx += c * ux; wy += c * uy; wz += c * uz;
            });
    }
    const float vol_i = xi.w;
    float4 F = make_float4(fx * vol_i, fy * vol_i, fz * vol_i, 0.f);
    F.x += wx * vol_i; F.y += wy * vol_i; F.z += wz * vol_i;
    viscous_force[i] = F;
    float4 P = previous_force[i], Fp = a.force_prior[i];
    Fp.x += F.x - P.x; Fp.y += F.y - P.y; Fp.z += F.z - P.z;
    a.force_prior[i] = Fp;
    previous_force[i] = F;
}

extern "C" int sphb200_viscous_force(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, float mu, float smoothing_length,
                                     sphb200_vec4_t *viscous_force, sphb200_vec4_t *previous_viscous_force, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s && viscous_force && previous_viscous_force, "null pointer");
    FArgs a;
    KTab dwtab;
    int rc = make_fargs(ctx, s, &a, nullptr, &dwtab);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, a.n == 0 || (a.posvol && a.vel && a.rec2 && a.force_prior && a.in_count && a.in_slice && a.in_index), "null fluid array");
    SPH_CHECK_ARG(ctx, a.n_wall == 0 || (a.w_posvol && a.ct_count && a.ct_slice && a.ct_index), "null wall array");
    SPH_CHECK_ARG(ctx, !s->material.correction || (a.B && a.brec), "LinearCorrectionCK needs fluid.B and fluid.correction_record");
    if (a.end <= a.begin) return 0;
    const float h2_eps = 0.01f * smoothing_length * smoothing_length;
    unsigned g = active_blocks(a, FL_THREADS);
    float4 *vf = (float4 *)viscous_force, *pf = (float4 *)previous_viscous_force;
    if (s->material.correction)
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_viscous_force<true, true>), g, FL_THREADS, 0, stream, a, dwtab, mu, h2_eps, vf, pf);
        else SPH_LAUNCH(ctx, (k_viscous_force<true, false>), g, FL_THREADS, 0, stream, a, dwtab, mu, h2_eps, vf, pf);
    }
    else
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_viscous_force<false, true>), g, FL_THREADS, 0, stream, a, dwtab, mu, h2_eps, vf, pf);
        else SPH_LAUNCH(ctx, (k_viscous_force<false, false>), g, FL_THREADS, 0, stream, a, dwtab, mu, h2_eps, vf, pf);
    }
    return 0;
}

// =====================================================================================================
// KernelGradientIntegral<Inner<Correction>, Contact<Boundary, Correction>>; general_dynamics/kernel_gradient_integral.hpp:33-78
//   kgi_i = - sum_j (B_i + B_j) dW_ij V_j e_ij - sum_wall 2 B_i dW_ij V_j e_ij
// =====================================================================================================
template <bool CORR, bool ANALYTIC>
__global__ void __launch_bounds__(FL_THREADS, FL_MIN_BLOCKS) k_kernel_gradient_integral(FArgs a, KTab dwtab, float4 *__restrict__ kgi)
{
    __shared__ float4 tab[KT_SLOTS];
    if (!ANALYTIC) stage_tab(dwtab, tab);
    u32 t = active_slot(a);
    if (t < a.begin || t >= a.end) return;
    const u32 i = a.order ? a.order[t] : t;
    const float4 xi = a.posvol[i];
    float Bi[9];
    if (CORR) load_mat(a.B, i, Bi);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    {
        u32 cnt = a.in_count[t];
        const u32 *idx = a.in_index + (u64)a.in_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_U];
        u32 js[NB_U];
        for_neighbors(
            idx, cnt,
            [&](int q, u32 j) {
                xjs[q] = a.posvol[j];
                if (CORR) js[q] = j;
            },
            [&](int q, bool valid) {
                const float4 xj = xjs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r, inv_r;
                dist(dx * dx + dy * dy + dz * dz, r, inv_r);
                float dWV = kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                float3 e = make_float3(dx * inv_r, dy * inv_r, dz * inv_r);
                if (CORR)
                {
                    float Bj[9];
                    load_mat(a.B, js[q], Bj);
#pragma unroll
                    for (int k = 0; k < 9; ++k) Bj[k] += Bi[k];
                    e = mat_vec(Bj, e);
                }
                else
                    dWV *= 2.0f;
                gx -= dWV * e.x; gy -= dWV * e.y; gz -= dWV * e.z;
            });
    }
    if (a.n_wall)
    {
        u32 cnt = a.ct_count[t];
        const u32 *idx = a.ct_index + (u64)a.ct_slice[t >> 5] + (t & 31u);
        float4 xjs[NB_U];
        for_neighbors(
            idx, cnt, [&](int q, u32 j) { xjs[q] = a.w_posvol[j]; },
            [&](int q, bool valid) {
                const float4 xj = xjs[q];
                float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                float r, inv_r;
                dist(dx * dx + dy * dy + dz * dz, r, inv_r);
                float dWV = 2.0f * kernel_dw<ANALYTIC>(a, tab, r) * xj.w;
                dWV = valid ? dWV : 0.f;
                float3 e = make_float3(dx * inv_r, dy * inv_r, dz * inv_r);
                if (CORR) e = mat_vec(Bi, e);
                gx -= dWV * e.x; gy -= dWV * e.y; gz -= dWV * e.z;
            });
    }
    kgi[i] = make_float4(gx, gy, gz, 0.f);
}

extern "C" int sphb200_kernel_gradient_integral(sphb200_context_t *ctx, const sphb200_fluid_args_t *s, sphb200_vec4_t *kernel_gradient_integral,
                                                void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && s && kernel_gradient_integral, "null pointer");
    FArgs a;
    KTab dwtab;
    int rc = make_fargs(ctx, s, &a, nullptr, &dwtab);
    if (rc) return rc;
    SPH_CHECK_ARG(ctx, a.n == 0 || (a.posvol && a.in_count && a.in_slice && a.in_index), "null fluid array");
    SPH_CHECK_ARG(ctx, a.n_wall == 0 || (a.w_posvol && a.ct_count && a.ct_slice && a.ct_index), "null wall array");
    SPH_CHECK_ARG(ctx, !s->material.correction || a.B, "LinearCorrectionCK needs fluid.B");
    if (a.end <= a.begin) return 0;
    unsigned g = active_blocks(a, FL_THREADS);
    float4 *out = (float4 *)kernel_gradient_integral;
    if (s->material.correction)
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_kernel_gradient_integral<true, true>), g, FL_THREADS, 0, stream, a, dwtab, out);
        else SPH_LAUNCH(ctx, (k_kernel_gradient_integral<true, false>), g, FL_THREADS, 0, stream, a, dwtab, out);
    }
    else
    {
        if (a.analytic) SPH_LAUNCH(ctx, (k_kernel_gradient_integral<false, true>), g, FL_THREADS, 0, stream, a, dwtab, out);
        else SPH_LAUNCH(ctx, (k_kernel_gradient_integral<false, false>), g, FL_THREADS, 0, stream, a, dwtab, out);
    }
    return 0;
}

// TransportVelocityCorrectionCK<..., Limiter, Scopes...>::UpdateKernel::update; fluid_dynamics/transport_velocity_correction_ck.hpp:39-50
//   dpos_i += coefficient h^2 limiter(h^2 |kgi_i|^2) kgi_i   (h_ratio == 1: single resolution), inside the particle scope
__global__ void __launch_bounds__(256)
    k_transport_velocity_correction(u32 n, float4 *__restrict__ dpos, const float4 *__restrict__ kgi, float scaling, float h2, int limiter,
                                    float slope, const int *__restrict__ indicator)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (indicator && indicator[i] != 0) return; // BulkParticles: Indicator == 0 (particle_functors_ck.h:69-98)
    float4 g = kgi[i];
    float sq = g.x * g.x + g.y * g.y + g.z * g.z;
    float lim = limiter ? fminf(slope * (h2 * sq), 1.0f) : 1.0f; // TruncatedLinear / NoLimiter, common_functors.h:69-94
    float c = scaling * lim;
    float4 d = dpos[i];
    d.x += c * g.x; d.y += c * g.y; d.z += c * g.z;
    dpos[i] = d;
}
extern "C" int sphb200_transport_velocity_correction(sphb200_context_t *ctx, const sphb200_fluid_view_t *f,
                                                     const sphb200_vec4_t *kernel_gradient_integral, float coefficient, float h_ref,
                                                     int limiter, float limiter_slope, const int32_t *indicator, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && f && f->dpos && kernel_gradient_integral, "null pointer");
    SPH_CHECK_ARG(ctx, limiter == 0 || limiter == 1, "limiter: 0 NoLimiter, 1 TruncatedLinear");
    Range r = active_range(f);
    if (r.n)
        SPH_LAUNCH(ctx, k_transport_velocity_correction, sph_blocks(r.n, 256), 256, 0, stream, r.n, (float4 *)f->dpos + r.b,
                   (const float4 *)kernel_gradient_integral + r.b, coefficient * h_ref * h_ref, h_ref * h_ref, limiter, limiter_slope,
                   indicator ? indicator + r.b : nullptr);
    return 0;
}
