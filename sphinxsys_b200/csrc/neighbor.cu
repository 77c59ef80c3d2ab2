// neighbor.cu — neighbour machinery: Morton keys, cell-linked list, relation (neighbour list) build.
// Replaces the device launches of ParticleSortCK::prepareSequence/updateSortedID
// (shared_ck/.../particle_sort_ck.hpp:61-104), UpdateCellLinkedList::exec (update_cell_linked_list.hpp:75-106)
// and UpdateRelation<Inner/Contact>::exec (update_body_relation.hpp:117-164,240-288).
//
// Differences from the reference data flow (results are the same sets, in a defined order):
//   * cell list: ONE atomic pass returns each particle's arrival rank; the fill pass reuses it, and a final
//     per-particle counting pass orders every cell by ascending particle index -> deterministic lists;
//   * relations: every particle searches its full 3^d box itself (no one-sided search, no atomics); lists are
//     stored in the coalesced SELL-32 layout described in sphb200.h.
#include "common.cuh"

// =====================================================================================================
// Morton keys
// =====================================================================================================
__global__ void __launch_bounds__(256)
    k_morton_keys(DMesh m, const float4 *__restrict__ pos, u32 n, u32 *__restrict__ keys, u32 *__restrict__ perm,
                  u32 *__restrict__ cell_id)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 x = pos[i];
    int a = cell_coord(x.x, m.lx, m.spacing, m.cx);
    int b = cell_coord(x.y, m.ly, m.spacing, m.cy);
    int c = cell_coord(x.z, m.lz, m.spacing, m.cz);
    if (keys) keys[i] = morton_spread(a) | (morton_spread(b) << 1) | (morton_spread(c) << 2);
    if (perm) perm[i] = i;
    if (cell_id) cell_id[i] = cell_linear(m, a, b, c);
}

extern "C" int sphb200_morton_keys(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos, uint32_t n,
                                   uint32_t *keys, uint32_t *perm, uint32_t *cell_id, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && mesh && (pos || n == 0), "null pointer");
    if (n == 0) return 0;
    SPH_LAUNCH(ctx, k_morton_keys, sph_blocks(n, 256), 256, 0, stream, make_dmesh(mesh), (const float4 *)pos, n, keys, perm,
               cell_id);
    return 0;
}

__global__ void k_update_sorted_id(const u32 *__restrict__ original_id, u32 *__restrict__ sorted_id, u32 n)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sorted_id[original_id[i]] = i;
}
extern "C" int sphb200_update_sorted_id(sphb200_context_t *ctx, const uint32_t *original_id, uint32_t *sorted_id, uint32_t n,
                                        void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((original_id && sorted_id) || n == 0), "null pointer");
    if (n) SPH_LAUNCH(ctx, k_update_sorted_id, sph_blocks(n, 256), 256, 0, stream, original_id, sorted_id, n);
    return 0;
}

// =====================================================================================================
// cell-linked list
// =====================================================================================================
// Kernels of the cell-list build take the particle count either from the host (n_dev == nullptr) or from device memory
// (n = min(n, *n_dev): slab-decomposed runs append arrivals whose number only the device knows, slab_decomposition.h).
__device__ __forceinline__ u32 live_count(u32 n, const u32 *n_dev) { return n_dev ? min(n, *n_dev) : n; }

// One pass: cell of every particle, its arrival rank inside the cell, the cell populations. Storage is cell ordered, so the
// lanes of a warp fall into very few cells: the lanes of one cell are found with __match_any_sync and ONE of them adds the
// group's size to the counter (warp-aggregated atomics: ~2 atomics per warp instead of 32 on the same two addresses).
__global__ void __launch_bounds__(256)
    k_cell_count(DMesh m, const float4 *__restrict__ pos, u32 n, const u32 *__restrict__ n_dev, u32 *__restrict__ counts,
                 u32 *__restrict__ cell_of, u32 *__restrict__ rank)
{
    n = live_count(n, n_dev);
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    u32 c = 0xffffffffu; // lanes past the end form their own group and do nothing
    if (live)
    {
        float4 x = pos[i];
        c = cell_linear(m, cell_coord(x.x, m.lx, m.spacing, m.cx), cell_coord(x.y, m.ly, m.spacing, m.cy),
                        cell_coord(x.z, m.lz, m.spacing, m.cz));
    }
    const u32 lane = threadIdx.x & 31u;
    const u32 peers = __match_any_sync(0xffffffffu, c);
    const u32 leader = __ffs(peers) - 1u;
    u32 base = 0;
    if (live && lane == leader) base = atomicAdd(&counts[c], (u32)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (live)
    {
        cell_of[i] = c;
        rank[i] = base + __popc(peers & ((1u << lane) - 1u));
    }
}
__global__ void __launch_bounds__(256)
    k_cell_fill(const u32 *__restrict__ cell_offset, const u32 *__restrict__ cell_of, const u32 *__restrict__ rank, u32 n,
                const u32 *__restrict__ n_dev, u32 *__restrict__ list)
{
    n = live_count(n, n_dev);
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[cell_offset[cell_of[i]] + rank[i]] = i;
}
// deterministic in-cell order: the final slot of i is the number of cell-mates with a smaller sort key
// (sort_key == nullptr: the particle index itself). Keys must be unique within a cell.
__global__ void __launch_bounds__(256)
    k_cell_order(const u32 *__restrict__ cell_offset, const u32 *__restrict__ cell_of, const u32 *__restrict__ unordered,
                 const u32 *__restrict__ sort_key, u32 n, const u32 *__restrict__ n_dev, u32 *__restrict__ particle_index,
                 const float4 *__restrict__ pos, float4 *__restrict__ sorted_pos)
{
    n = live_count(n, n_dev);
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 c = cell_of[i];
    u32 b = cell_offset[c], e = cell_offset[c + 1];
    u32 smaller = 0;
    if (sort_key)
    {
        u32 ki = sort_key[i];
        for (u32 k = b; k < e; ++k) smaller += (sort_key[unordered[k]] < ki);
    }
    else
        for (u32 k = b; k < e; ++k) smaller += (unordered[k] < i);
    particle_index[b + smaller] = i;
    if (sorted_pos) sorted_pos[b + smaller] = pos[i];
}

static int cell_list_build(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos, uint32_t n,
                           const uint32_t *n_dev, const uint32_t *sort_key, sphb200_cell_list_t list, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && mesh && list.cell_offset && list.particle_index && (pos || n == 0), "null pointer");
    u64 cells = (u64)mesh->cells[0] * mesh->cells[1] * mesh->cells[2];
    SPH_CHECK_ARG(ctx, cells > 0 && cells < (1ull << 32) - 1, "bad cell count");
    cudaStream_t st = (cudaStream_t)stream;
    void *p;
    int rc = sph_scratch(ctx, 2, (cells + 1 + 3 * (size_t)n) * sizeof(u32) + 64, &p);
    if (rc) return rc;
    u32 *counts = (u32 *)p, *cell_of = counts + cells + 1, *rank = cell_of + n, *unordered = rank + n;
    SPH_CUDA(ctx, cudaMemsetAsync(counts, 0, (cells + 1) * sizeof(u32), st));
    DMesh m = make_dmesh(mesh);
    if (n) SPH_LAUNCH(ctx, k_cell_count, sph_blocks(n, 256), 256, 0, st, m, (const float4 *)pos, n, n_dev, counts, cell_of, rank);
    rc = sph_scan_u32(ctx, counts, list.cell_offset, cells + 1, 0, st);
    if (rc) return rc;
    if (n)
    {
        SPH_LAUNCH(ctx, k_cell_fill, sph_blocks(n, 256), 256, 0, st, list.cell_offset, cell_of, rank, n, n_dev, unordered);
        SPH_LAUNCH(ctx, k_cell_order, sph_blocks(n, 256), 256, 0, st, list.cell_offset, cell_of, unordered, sort_key, n, n_dev,
                   list.particle_index, (const float4 *)pos, (float4 *)list.sorted_pos);
    }
    return 0;
}

extern "C" int sphb200_cell_list_build(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos,
                                       uint32_t n, sphb200_cell_list_t list, void *stream)
{
    return cell_list_build(ctx, mesh, pos, n, nullptr, nullptr, list, stream);
}

int sph_gather_multi_n(sphb200_context_t *ctx, int count, void *const *dst, const void *const *src, const uint32_t *elem_bytes,
                       const uint32_t *perm, uint32_t n, const uint32_t *n_dev, void *stream); // primitives.cu

// UpdateCellLinkedList for a body whose STORAGE follows the cell order: builds the list from the
// current positions, gathers every listed variable into its shadow buffer by the resulting permutation
// (dst_k[slot] = src_k[particle_index[slot]]) and leaves particle_index = identity. In-cell order = ascending
// sort_key (the reference particle id), so lists and sums do not depend on the storage history.
extern "C" int sphb200_cell_list_build_reorder_n(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos,
                                                 uint32_t n, const uint32_t *n_dev, const uint32_t *sort_key,
                                                 sphb200_cell_list_t list, int count, void *const *dst, const void *const *src,
                                                 const uint32_t *elem_bytes, void *stream)
{
    int rc = cell_list_build(ctx, mesh, pos, n, n_dev, sort_key, list, stream);
    if (rc) return rc;
    if (n == 0) return 0;
    rc = sph_gather_multi_n(ctx, count, dst, src, elem_bytes, list.particle_index, n, n_dev, stream);
    if (rc) return rc;
    return sphb200_iota_u32(ctx, list.particle_index, n, stream);
}
extern "C" int sphb200_cell_list_build_reorder(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos,
                                               uint32_t n, const uint32_t *sort_key, sphb200_cell_list_t list, int count,
                                               void *const *dst, const void *const *src, const uint32_t *elem_bytes,
                                               void *stream)
{
    return sphb200_cell_list_build_reorder_n(ctx, mesh, pos, n, nullptr, sort_key, list, count, dst, src, elem_bytes, stream);
}

// Slab bookkeeping after a rebuild whose particle count lives on the device (slab_decomposition.h): out[0..k) = the cell
// offsets at `cells[0..k)` (the plane boundaries of the slab), out[k] = *n_dev, out[k+1] = the mailbox status word (or 0);
// own64[0] = out[own_hi] - out[own_lo] as a 64-bit word, ready for the all-gather of the ranks' own counts.
__global__ void k_slab_bounds(const u32 *__restrict__ cell_offset, const u32 *__restrict__ cells, int k, const u32 *__restrict__ n_dev,
                              const u32 *__restrict__ status, u32 *__restrict__ out, int own_lo, int own_hi, u64 *__restrict__ own64)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        for (int c = 0; c < k; ++c) out[c] = cell_offset[cells[c]];
        out[k] = n_dev ? *n_dev : 0u;
        out[k + 1] = status ? *status : 0u;
        if (own64) own64[0] = (u64)(out[own_hi] - out[own_lo]);
    }
}
__global__ void k_slab_total(u32 base, const u32 *a, const u32 *b, u32 *out) { *out = base + (a ? *a : 0u) + (b ? *b : 0u); }
extern "C" int sphb200_slab_total(sphb200_context_t *ctx, uint32_t base, const uint32_t *a_dev, const uint32_t *b_dev, uint32_t *out_dev,
                                  void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && out_dev, "null pointer");
    SPH_LAUNCH(ctx, k_slab_total, 1, 1, 0, stream, base, a_dev, b_dev, out_dev);
    return 0;
}
extern "C" int sphb200_slab_bounds(sphb200_context_t *ctx, const uint32_t *cell_offset, const uint32_t *cells_dev, int k,
                                   const uint32_t *n_dev, int own_lo_hi, uint32_t *out_dev, uint64_t *own64_dev, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && cell_offset && cells_dev && out_dev && k > 0 && k <= 16, "bad arguments");
    const int lo = own_lo_hi & 0xff, hi = (own_lo_hi >> 8) & 0xff;
    SPH_CHECK_ARG(ctx, lo < k && hi < k, "own_lo_hi outside the cell list");
    SPH_LAUNCH(ctx, k_slab_bounds, 1, 32, 0, stream, cell_offset, cells_dev, k, n_dev, sphb200_comm_mailbox_status(ctx), out_dev, lo, hi, own64_dev);
    return 0;
}

// =====================================================================================================
// relations (neighbour lists)
// =====================================================================================================
struct SearchArgs
{
    DMesh m;
    const float4 *src_pos;
    const float4 *src_sorted_pos;
    const u32 *src_order;
    const float4 *tar_pos;
    const float4 *tar_sorted_pos;
    const u32 *cell_offset;
    const u32 *particle_index;
    u32 n_src;
    u32 src_begin, src_end; // source slots searched
    float inv_h, ks2;
    float rc2; // legacy criterion threshold (kernel_size * h)^2
    int legacy_criterion;
    int depth;
    int cell_ordered;
    // optional second candidate set on the same mesh (periodic images stored behind the real particles)
    const float4 *tar2_pos;
    const u32 *cell_offset2;
    u32 index_base2;
};
// launches start at the 32-aligned slot below src_begin so that lane == slot % 32 (SELL-32 layout)
__device__ __forceinline__ u32 search_slot(const SearchArgs &a) { return (a.src_begin & ~31u) + blockIdx.x * blockDim.x + threadIdx.x; }
static inline unsigned search_blocks(const SearchArgs &a, unsigned threads) { return sph_blocks(a.src_end - (a.src_begin & ~31u), threads); }
// the first active lane of a warp publishes per-slice values
__device__ __forceinline__ bool slice_writer(bool active)
{
    u32 b = __ballot_sync(0xffffffffu, active);
    return active && (threadIdx.x & 31) == (u32)(__ffs(b) - 1);
}

// Neighbor<SPHAdaptation,SPHAdaptation>::NeighborCriterion, neighbor_method.hpp:152-156; every op rounded
// separately so that set membership is bit-identical to the CPU evaluation.
__device__ __forceinline__ bool within(float4 xi, float4 xj, float inv_h, float ks2)
{
    float sx = __fmul_rn(inv_h, __fsub_rn(xi.x, xj.x));
    float sy = __fmul_rn(inv_h, __fsub_rn(xi.y, xj.y));
    float sz = __fmul_rn(inv_h, __fsub_rn(xi.z, xj.z));
    float r2 = __fadd_rn(__fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy)), __fmul_rn(sz, sz));
    return r2 < ks2;
}
// legacy NeighborBuilder criterion, kernels/base_kernel.h:105-114: displacement.squaredNorm() < rc_ref_sqr
__device__ __forceinline__ bool within_legacy(float4 xi, float4 xj, float rc2)
{
    float dx = __fsub_rn(xi.x, xj.x), dy = __fsub_rn(xi.y, xj.y), dz = __fsub_rn(xi.z, xj.z);
    float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return r2 < rc2;
}
__device__ __forceinline__ bool criterion(const float4 &xi, const float4 &xj, float inv_h, float ks2, float rc2, int legacy)
{
    return legacy ? within_legacy(xi, xj, rc2) : within(xi, xj, inv_h, ks2);
}

// Enumerate candidates in the reference order: cells x -> y -> z (mesh_iterators.hpp:18-27); for fixed (x, y)
// the z cells are contiguous in the linear index, so each (x, y) column is ONE run of the cell list. With
// SORTED (cell-ordered position copy available) a run is a contiguous float4 stream and the particle id is only
// fetched for hits; otherwise positions are gathered through particle_index.
template <bool INNER, bool SORTED, class F>
__device__ __forceinline__ void for_each_neighbor(const SearchArgs &a, u32 i, float4 xi, F f)
{
    const DMesh &m = a.m;
    int ca = cell_coord(xi.x, m.lx, m.spacing, m.cx);
    int cb = cell_coord(xi.y, m.ly, m.spacing, m.cy);
    int cc = cell_coord(xi.z, m.lz, m.spacing, m.cz);
    int x0 = max(0, ca - a.depth), x1 = min(m.cx, ca + a.depth + 1);
    int y0 = max(0, cb - a.depth), y1 = min(m.cy, cb + a.depth + 1);
    int z0 = max(0, cc - a.depth), z1 = min(m.cz, cc + a.depth + 1);
    for (int x = x0; x < x1; ++x)
        for (int y = y0; y < y1; ++y)
        {
            u32 lin0 = cell_linear(m, x, y, z0);
            u32 b = a.cell_offset[lin0], e = a.cell_offset[lin0 + (u32)(z1 - z0)];
#pragma unroll 4
            for (u32 k = b; k < e; ++k)
            {
                if (SORTED)
                {
                    float4 xj = a.tar_sorted_pos[k];
                    if (criterion(xi, xj, a.inv_h, a.ks2, a.rc2, a.legacy_criterion))
                    {
                        u32 j = a.particle_index[k];
                        if (!(INNER && j == i)) f(j);
                    }
                }
                else
                {
                    u32 j = a.particle_index[k];
                    if (INNER && j == i) continue;
                    float4 xj = a.tar_pos[j];
                    if (criterion(xi, xj, a.inv_h, a.ks2, a.rc2, a.legacy_criterion)) f(j);
                }
            }
        }
}

__device__ __forceinline__ void load_source(const SearchArgs &a, u32 t, u32 &i, float4 &xi)
{
    i = a.src_order ? a.src_order[t] : t;
    xi = a.src_sorted_pos ? a.src_sorted_pos[t] : a.src_pos[i];
}

// MODE 0: count only (exact two-phase build, phase 1)
// MODE 1: fill only  (phase 2; slice offsets come from the scan)
// MODE 2: one pass, fixed slice stride: count + fill + running max of the row length
template <bool INNER, bool SORTED, int MODE>
__global__ void __launch_bounds__(128)
    k_relation(SearchArgs a, u32 *__restrict__ count, u32 *__restrict__ slice, u32 *__restrict__ index, u64 capacity, u32 stride,
               u32 *__restrict__ max_count)
{
    u32 t = search_slot(a);
    u32 c = 0;
    const bool active = t >= a.src_begin && t < a.src_end;
    if (active)
    {
        u32 i;
        float4 xi;
        load_source(a, t, i, xi);
        if (MODE == 0)
            for_each_neighbor<INNER, SORTED>(a, i, xi, [&](u32) { ++c; });
        else
        {
            u64 base = (MODE == 1 ? (u64)slice[t >> 5] : (u64)(t >> 5) * 32ull * stride) + (t & 31u);
            u32 limit = MODE == 2 ? stride : 0xffffffffu;
            for_each_neighbor<INNER, SORTED>(a, i, xi, [&](u32 j) {
                u64 pos = base + 32ull * c;
                if (c < limit && pos < capacity) index[pos] = j;
                ++c;
            });
        }
        if (MODE != 1) count[t] = c;
    }
    if (MODE == 0)
    {
        u32 mx = warp_max_u32(c);
        if ((threadIdx.x & 31) == 0 && t < a.n_src) slice[t >> 5] = mx * 32u;
    }
    if (MODE == 2)
    {
        u32 mx = warp_max_u32(c);
        if ((threadIdx.x & 31) == 0 && t < a.n_src)
        {
            slice[t >> 5] = (t >> 5) * 32u * stride;
            if (mx > 0) atomicMax(max_count, mx);
        }
    }
}

// One lane's column of a SELL slice brought into the bank-aligned layout (see "L1-bank-aligned rows" below): `col` is the
// lane's first entry (entries 32 apart), `tile` its column of a [32][32] shared-memory tile (entries 32 apart: bank ==
// lane), c its row count, cmax the warp's. Every lane only touches its own column of the tile and of the slice, so no
// warp synchronisation is needed — also not against the stores the same lane made to `col` earlier in the kernel.
constexpr int BA_GROUP = 32;
__device__ __forceinline__ void bank_align_column(u32 *__restrict__ col, u32 *__restrict__ tile, u32 c, u32 cmax, u32 t, u32 origin)
{
    for (u32 g = 0; g < cmax; g += BA_GROUP)
    {
        // rows [g, g + 32) of the warp: coalesced 128-byte loads, all independent (rows < cmax exist for every lane)
        u32 sv[BA_GROUP];
#pragma unroll
        for (int u = 0; u < BA_GROUP; ++u) sv[u] = g + u < cmax ? col[32ull * (g + u)] : 0u;
        const u32 n = c > g ? min(c - g, (u32)BA_GROUP) : 0u; // entries of this lane in the group
        // Entry u wants a position p with (g + p) = (s - t) mod 8, i.e. p = p0 + 8 q, p0 = (s - t - g) mod 8, and takes
        // the first free one of its class. Positions of a class fill in the order q = 0, 1, 2, 3, so "first free" is a
        // per-class counter: eight 4-bit counters in one word — one short add chain instead of a used-mask/ffs chain
        // per entry. Entries whose class is full (p >= n) wait and fill the holes afterwards, lowest hole first.
        const u32 shift0 = t + origin + g;
        u32 taken = 0, waiting = 0; // 4-bit count per class / entries not yet placed
#pragma unroll
        for (int u = 0; u < BA_GROUP; ++u)
        {
            if ((u32)u >= n) continue;
            const u32 p0 = (sv[u] - shift0) & 7u;
            const u32 q = (taken >> (4u * p0)) & 15u;
            const u32 p = p0 + 8u * q;
            if (p < n)
            {
                tile[32u * p] = sv[u];
                taken += 1u << (4u * p0);
            }
            else
                waiting |= 1u << u;
        }
        if (waiting)
        {
            // positions taken so far: class p0 holds p0, p0 + 8, ... for its count
            u32 used = 0;
#pragma unroll
            for (u32 p0 = 0; p0 < 8u; ++p0)
            {
                const u32 k = (taken >> (4u * p0)) & 15u;
                used |= (0x01010101u & ((k >= 4u ? 0xffffffffu : (1u << (8u * k)) - 1u))) << p0;
            }
#pragma unroll
            for (int u = 0; u < BA_GROUP; ++u)
            {
                if (!((waiting >> u) & 1u)) continue;
                const u32 p = __ffs(~used) - 1u; // lowest free position (< n: as many free positions as waiting entries)
                tile[32u * p] = sv[u];
                used |= 1u << p;
            }
        }
#pragma unroll 8
        for (u32 p = 0; p < n; ++p) col[32ull * (g + p)] = tile[32u * p];
    }
}

// packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2 — two IEEE fp32 operations per issue slot)
__device__ __forceinline__ unsigned long long pk2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// -----------------------------------------------------------------------------------------------------
// Warp-uniform search for CELL-ORDERED storage (slot == particle id on both sides).
// The 32 lanes of a warp are storage neighbours, i.e. they sit in the same (x, y) cell column and in a few
// consecutive z cells. Lanes of one column form a group; for each of the (2d+1)^2 neighbouring columns the group
// walks ONE contiguous run of target slots (cells z_min-d .. z_max+d), so control flow is uniform and every
// candidate load is a broadcast. A lane accepts candidate k only inside its own clamped cell window
// [lo, hi) — exactly the cells the reference visits (cell_linked_list.hpp:113-167) — and rows come out in the
// reference order (cells x -> y -> z, then in-cell order).
// The criterion is decided by the sign of a fused-arithmetic estimate; a lane whose chunk holds a candidate within
// 2e-5 (relative) of the threshold re-decides that chunk with the separately rounded reference expression (within()),
// so set membership stays bit-identical. Per candidate: one broadcast LDS.128, 3 FADD, 3 FFMA, one funnel shift (sign
// bit into the hit mask) and one FMNMX (ambiguity tracker) — the integer/predicate pipe runs at half the FP32 rate on
// sm_100 (scripts/microbench/ffma2.cu), which is what bounded the two-threshold FSETP/SEL form used before.
// -----------------------------------------------------------------------------------------------------
// tuning (scripts/gpu_variants.sh): chunk masks a lane may hold back, minimum resident blocks (0: the compiler's choice)
#ifndef SPH_REL_NW
#define SPH_REL_NW 32
#endif
#ifndef SPH_REL_MINB
#define SPH_REL_MINB 0
#endif
#if SPH_REL_MINB > 0
#define SPH_REL_BOUNDS __launch_bounds__(128, SPH_REL_MINB)
#else
#define SPH_REL_BOUNDS __launch_bounds__(128)
#endif
template <bool INNER, int MODE, bool TWO>
__global__ void SPH_REL_BOUNDS
    k_relation_ordered(SearchArgs a, u32 *__restrict__ count, u32 *__restrict__ slice, u32 *__restrict__ index, u64 capacity,
                       u32 stride, u32 *__restrict__ max_count, int align_origin)
{
    constexpr int CH = 32; // candidates tested per chunk; hits of a chunk are collected in a per-lane bit mask
    constexpr int NW = SPH_REL_NW; // chunk masks a lane may hold back before its hits are written out
    __shared__ __align__(16) float rel_tile[4][2][3][CH]; // per warp: two tiles of x[CH], y[CH], z[CH]
    __shared__ u32 rel_mask[4][NW][32];
    __shared__ u32 rel_base[4][NW];
    const u32 lane = threadIdx.x & 31u;
    float(*const warp_tile)[3][CH] = rel_tile[threadIdx.x >> 5];
    u32 *const my_mask = &rel_mask[threadIdx.x >> 5][0][lane]; // word w of this lane at my_mask[32 w] (bank == lane)
    u32 *const warp_base = rel_base[threadIdx.x >> 5];
    u32 buf = 0;
    u32 nw = 0, nzw = 0, pend = 0; // held-back words, bitmap of the non-empty ones, hits in them
    const float4 far = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f); // stands in for candidates past the stored range
    const u32 t = search_slot(a);
    const bool active = t >= a.src_begin && t < a.src_end;
    const DMesh &m = a.m;
    float4 xi = active ? a.src_pos[t] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int ca = cell_coord(xi.x, m.lx, m.spacing, m.cx);
    const int cb = cell_coord(xi.y, m.ly, m.spacing, m.cy);
    const int cc = cell_coord(xi.z, m.lz, m.spacing, m.cz);
    const int d = a.depth;
    // The fused estimate s = |xi - xj|^2 - thr (thr = the support radius squared, unscaled) decides a candidate by its
    // sign; `amb` tracks min |s| over the chunk. Only when some candidate of the chunk lies within `band` of the threshold
    // (relative 2e-5: two orders above the rounding difference between the estimate and the reference expression) is the
    // chunk of that lane re-decided with the separately rounded reference expression (criterion()).
    const float h2 = 1.0f / (a.inv_h * a.inv_h);
    const float thr = a.legacy_criterion ? a.rc2 : a.ks2 * h2, band = 2.0e-5f * thr;
    const unsigned long long nthr2 = pk2(-thr, -thr), nxi_x = pk2(-xi.x, -xi.x), nxi_y = pk2(-xi.y, -xi.y), nxi_z = pk2(-xi.z, -xi.z);
    // entries of this slot live at index[off], index[off + 32], ...; off_end bounds what may be written
    u64 base64 = (MODE == 1 ? (u64)(active ? slice[t >> 5] : 0u) : (u64)(t >> 5) * 32ull * stride) + (t & 31u);
    u64 room = capacity > base64 ? (capacity - base64 + 31ull) / 32ull : 0ull; // rows that fit below `capacity`
    if (MODE == 2 && room > stride) room = stride;
    u32 *const out = index + base64;
    const u32 row_limit = (u32)(room > 0xffffffffull ? 0xffffffffull : room);
    u32 c = 0;
    // Deferred emission. Writing the hits of each chunk at once costs max-over-lanes(hits in the chunk) iterations per
    // chunk — 3.4x the mean, because the lanes' hits peak in different chunks (measured: 293 iterations per warp for 85
    // rows, half of the kernel's stall samples). The chunk masks are held back in shared memory instead and written in
    // ONE flat loop whose trip count is the largest pending hit count of the warp (~100): every iteration each lane
    // takes its next hit — lowest word first, lowest bit first, i.e. still the reference search order.
    auto flush = [&]() {
        __syncwarp();
        const u32 trip = warp_max_u32(pend);
        const u32 todo_w = c < row_limit ? min(pend, row_limit - c) : 0u; // rows beyond the limit are only counted
        u32 bits = 0, cur = 0, cw = c;
        for (u32 k = 0; k < trip; ++k)
            if (k < todo_w)
            {
                if (bits == 0)
                {
                    const u32 w = __ffs(nzw) - 1;
                    nzw &= nzw - 1;
                    bits = my_mask[32u * w];
                    cur = warp_base[w];
                }
                const u32 b = __ffs(bits) - 1;
                bits &= bits - 1;
                out[32ull * cw] = cur + b;
                ++cw;
            }
        c += pend;
        nw = 0, nzw = 0, pend = 0;
        __syncwarp();
    };
    for (int pass = 0; pass < (TWO ? 2 : 1); ++pass)
    {
    // pass 1 walks the second candidate set (periodic images): same mesh, its own cell list, no self exclusion
    const float4 *__restrict__ tpos = TWO && pass ? a.tar2_pos : a.tar_pos;
    const u32 *__restrict__ coff = TWO && pass ? a.cell_offset2 : a.cell_offset;
    const u32 ibase = TWO && pass ? a.index_base2 : 0u;
    const bool self_excl = INNER && !(TWO && pass);
    const u32 n_tar = coff[(u32)m.cx * (u32)m.cy * (u32)m.cz]; // stored candidates: full chunks may read (not accept) past a run
    u32 todo = __ballot_sync(0xffffffffu, active);
    while (todo) // the WHOLE warp walks the runs of every column group (uniform control flow); non-members accept nothing
    {
        const int leader = __ffs(todo) - 1;
        const int la = __shfl_sync(0xffffffffu, ca, leader), lb = __shfl_sync(0xffffffffu, cb, leader);
        const bool in = active && ca == la && cb == lb;
        todo &= ~__ballot_sync(0xffffffffu, in);
        const int zmin = __reduce_min_sync(0xffffffffu, in ? cc : 0x7fffffff), zmax = __reduce_max_sync(0xffffffffu, in ? cc : -1);
        const int x0 = max(0, la - d), x1 = min(m.cx, la + d + 1);
        const int y0 = max(0, lb - d), y1 = min(m.cy, lb + d + 1);
        const int z0 = max(0, zmin - d), z1 = min(m.cz, zmax + d + 1);
        const int wz0 = max(0, cc - d), wz1 = min(m.cz, cc + d + 1);
        for (int x = x0; x < x1; ++x)
            for (int y = y0; y < y1; ++y)
            {
                const u32 col = cell_linear(m, x, y, 0);
                const u32 rb = coff[col + z0], re = coff[col + z1];
                const u32 lo = in ? coff[col + wz0] : 0u, hi = in ? coff[col + wz1] : 0u;
                // chunk staging: lane l fetches candidate kb + l (one coalesced 512-byte load per chunk), the warp
                // exchanges the chunk through its shared-memory tile and every lane reads candidate b with a broadcast
                // LDS.128. (A broadcast LDG.128 costs 4 L1 data-pipe wavefronts — it is served per quarter warp even when
                // all lanes read the same 16 bytes — which bounded this kernel at 83 % L1 before; profiles/r01_v5_*.)
                // Two tiles per warp: one __syncwarp per chunk orders the stores of chunk i+2 behind the reads of chunk i.
                float4 nxt = rb + lane < n_tar ? tpos[rb + lane] : far;
                for (u32 kb = rb; kb < re; kb += CH)
                {
                    float(*tile)[CH] = warp_tile[buf];
                    buf ^= 1u;
                    tile[0][lane] = nxt.x, tile[1][lane] = nxt.y, tile[2][lane] = nxt.z;
                    __syncwarp();
                    if (kb + CH < re) nxt = kb + CH + lane < n_tar ? tpos[kb + CH + lane] : far;
                    // hits: the sign bit of s is shifted into the mask with one funnel shift per candidate (candidate b
                    // ends at bit 31-b, reversed afterwards); candidates past the run are dropped by the window mask.
                    // Four candidates per step: three broadcast LDS.128 (x, y, z of candidates 4q..4q+3), then two packed
                    // evaluations of s = |xj - xi|^2 - thr (3 FADD2 + 3 FFMA2 each).
                    u32 hits = 0;
                    float amb = 3.0e38f;
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q)
                    {
                        const float4 X = *reinterpret_cast<const float4 *>(&tile[0][4 * q]);
                        const float4 Y = *reinterpret_cast<const float4 *>(&tile[1][4 * q]);
                        const float4 Z = *reinterpret_cast<const float4 *>(&tile[2][4 * q]);
                        const unsigned long long dx0 = add2(pk2(X.x, X.y), nxi_x), dx1 = add2(pk2(X.z, X.w), nxi_x);
                        const unsigned long long dy0 = add2(pk2(Y.x, Y.y), nxi_y), dy1 = add2(pk2(Y.z, Y.w), nxi_y);
                        const unsigned long long dz0 = add2(pk2(Z.x, Z.y), nxi_z), dz1 = add2(pk2(Z.z, Z.w), nxi_z);
                        const unsigned long long s01 = fma2(dz0, dz0, fma2(dy0, dy0, fma2(dx0, dx0, nthr2)));
                        const unsigned long long s23 = fma2(dz1, dz1, fma2(dy1, dy1, fma2(dx1, dx1, nthr2)));
                        float s0, s1, s2, s3;
                        upk2(s01, s0, s1);
                        upk2(s23, s2, s3);
                        hits = __funnelshift_l(__float_as_uint(s0), hits, 1);
                        hits = __funnelshift_l(__float_as_uint(s1), hits, 1);
                        hits = __funnelshift_l(__float_as_uint(s2), hits, 1);
                        hits = __funnelshift_l(__float_as_uint(s3), hits, 1);
                        amb = fminf(fminf(amb, fminf(fabsf(s0), fabsf(s1))), fminf(fabsf(s2), fabsf(s3)));
                    }
                    hits = __brev(hits);
                    // the lane's own cell window [lo, hi) as a bit range of this chunk; INNER: not itself
                    const u32 blo = lo > kb ? min(lo - kb, 32u) : 0u, bhi = hi > kb ? min(hi - kb, 32u) : 0u;
                    u32 wmask = (bhi >= 32u ? 0xffffffffu : (1u << bhi) - 1u) & ~(blo >= 32u ? 0xffffffffu : (1u << blo) - 1u);
                    if (self_excl && t - kb < (u32)CH) wmask &= ~(1u << (t - kb));
                    if (amb <= band && wmask) // rare: some candidate within 2e-5 of the threshold -> reference expression
                    {
                        hits = 0;
                        for (u32 w = wmask; w;)
                        {
                            const u32 b = __ffs(w) - 1;
                            w &= w - 1;
                            if (criterion(xi, make_float4(tile[0][b], tile[1][b], tile[2][b], 0.f), a.inv_h, a.ks2, a.rc2, a.legacy_criterion)) hits |= 1u << b;
                        }
                    }
                    hits &= wmask;
                    if (MODE == 0)
                        c += __popc(hits);
                    else
                    {
                        if (nw == NW) flush();
                        my_mask[32u * nw] = hits;
                        warp_base[nw] = ibase + kb; // same value from every lane
                        nzw |= (hits != 0u ? 1u : 0u) << nw;
                        pend += __popc(hits);
                        ++nw;
                    }
                }
            }
    }
    }
    if (MODE != 0) flush();
    // bank-aligned rows (align_origin >= 0): the warp re-lays out the rows it has just written while they are still in
    // L2 — the stand-alone pass (k_bank_align) streamed the whole index array from and to DRAM once more
    if (MODE == 2 && align_origin >= 0)
    {
        const u32 cw = active ? min(c, row_limit) : 0u;
        bank_align_column(out, my_mask, cw, warp_max_u32(cw), t, (u32)align_origin);
    }
    if (MODE != 1 && active) count[t] = c;
    if (MODE == 0)
    {
        u32 mx = warp_max_u32(c);
        if (slice_writer(active)) slice[t >> 5] = mx * 32u;
    }
    if (MODE == 2)
    {
        u32 mx = warp_max_u32(c);
        if (slice_writer(active))
        {
            slice[t >> 5] = (t >> 5) * 32u * stride;
            if (mx > 0) atomicMax(max_count, mx);
        }
    }
}


// -----------------------------------------------------------------------------------------------------
// L1-bank-aligned rows (sphb200_relation_t::bank_aligned). A per-lane gather of 16-byte records is served per quarter
// warp, and its cost is the bank-conflict degree of the 8 addresses, not the number of cache lines touched (measured:
// scripts/microbench/l1_gather.cu, profiles/r01_l1_gather_microbench.txt — 11.5 data-pipe wavefronts per warp gather for
// random slots, 5.75 when (slot mod 8) is distinct inside every quarter warp, 5.2 fully coalesced). The rows of a slot
// are therefore permuted so that, as far as possible, the entry s in row k of slot t satisfies (s - t - k) mod 8 == 0:
// in row k lane l then reads a record of bank class (l + k) mod 8, distinct across the 8 lanes of a quarter warp.
// The permutation is local: each group of 32 consecutive rows (entries in ascending order) is re-laid out on its own —
// an entry takes the first free position of its class inside the group, the leftovers fill the remaining positions —
// which aligns ~87 % of the entries (a global layout reaches 92 %) but needs only a 4 KB shared-memory tile per warp
// and two coalesced passes over the slice. Entry SET and count of every row are untouched; the class depends only on
// s - t, i.e. for an INNER relation not on where the slab of a decomposed run starts, so summation order and every result
// bit are the same on 1 and on N GPUs. For a CONTACT relation the target slots do not move with the source slab, so the
// caller passes the slab's slot origin (the slot its first stored particle has in the undecomposed run, mod 8;
// sphb200_relation_t::bank_aligned = 1 + origin) and the class is taken from s - (t + origin).
// -----------------------------------------------------------------------------------------------------
constexpr int BA_WARPS = 4;
__global__ void __launch_bounds__(32 * BA_WARPS)
    k_bank_align(u32 *__restrict__ index, const u32 *__restrict__ count, const u32 *__restrict__ slice, u32 first_slice, u32 n_slices,
                 u32 src_begin, u32 src_end, u32 stride, u32 origin)
{
    __shared__ u32 ba_tile[BA_WARPS][BA_GROUP][32];
    const u32 lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const u32 sl = first_slice + blockIdx.x * BA_WARPS + w;
    if (sl >= first_slice + n_slices) return;
    const u32 t = sl * 32u + lane;
    const bool active = t >= src_begin && t < src_end;
    const u32 c = active ? min(count[t], stride) : 0u;
    bank_align_column(index + (u64)slice[sl] + lane, &ba_tile[w][0][lane], c, warp_max_u32(c), t, origin);
}

static int bank_align(sphb200_context *ctx, const SearchArgs &a, u32 *count, u32 *slice, u32 *index, u32 stride, u32 origin, cudaStream_t st)
{
    if (a.src_end <= a.src_begin) return 0;
    const u32 first = a.src_begin >> 5, n_slices = ((a.src_end + 31u) >> 5) - first;
    SPH_LAUNCH(ctx, k_bank_align, sph_blocks(n_slices, BA_WARPS), 32 * BA_WARPS, 0, st, index, count, slice, first, n_slices, a.src_begin,
               a.src_end, stride, origin);
    return 0;
}

static int make_search(sphb200_context *ctx, const sphb200_search_t *s, SearchArgs *a)
{
    SPH_CHECK_ARG(ctx, s->search_depth >= 1 && s->search_depth <= 4, "search depth out of range");
    a->m = make_dmesh(&s->tar_mesh);
    a->src_pos = (const float4 *)s->src_pos;
    a->src_sorted_pos = (const float4 *)s->src_sorted_pos;
    a->src_order = s->src_order;
    a->tar_pos = (const float4 *)s->tar_pos;
    a->tar_sorted_pos = (const float4 *)s->tar_list.sorted_pos;
    a->cell_offset = s->tar_list.cell_offset;
    a->particle_index = s->tar_list.particle_index;
    a->n_src = s->n_src;
    a->src_begin = s->src_end ? s->src_begin : 0u;
    a->src_end = s->src_end ? s->src_end : s->n_src;
    SPH_CHECK_ARG(ctx, a->src_begin <= a->src_end && a->src_end <= s->n_src, "source slot range outside [0, n_src)");
    a->inv_h = 1.0f / s->kernel.h; // inv_h_ = 1 / max(src_h, tar_h), neighbor_method.hpp:73-76
    a->ks2 = s->kernel.kernel_size * s->kernel.kernel_size;
    {
        float rc = s->kernel.kernel_size * s->kernel.h;
        a->rc2 = rc * rc;
    }
    a->legacy_criterion = s->legacy_criterion;
    a->depth = s->search_depth;
    a->cell_ordered = s->cell_ordered;
    a->tar2_pos = (const float4 *)s->tar2_pos;
    a->cell_offset2 = s->tar2_list.cell_offset;
    a->index_base2 = s->tar2_index_base;
    SPH_CHECK_ARG(ctx, !s->tar2_pos || (s->cell_ordered && s->tar2_list.cell_offset), "a second candidate set needs a cell_ordered search and its cell list");
    if (s->cell_ordered && s->n_src)
    {
        SPH_CHECK_ARG(ctx, a->src_pos && a->tar_pos && a->cell_offset && !a->src_order, "cell_ordered search needs src_pos, tar_pos, cell_offset and no src_order");
        return 0;
    }
    if (s->n_src)
    {
        SPH_CHECK_ARG(ctx, (a->src_pos || (a->src_sorted_pos && a->src_order)) && a->tar_pos && a->cell_offset && a->particle_index,
                      "null pointer");
        SPH_CHECK_ARG(ctx, !s->is_inner || a->src_pos || a->src_order, "null pointer");
    }
    return 0;
}

template <int MODE>
static int launch_relation(sphb200_context *ctx, const SearchArgs &a, bool inner, u32 *count, u32 *slice, u32 *index, u64 cap,
                           u32 stride, u32 *max_count, cudaStream_t st, int align_origin = -1, bool *aligned = nullptr)
{
    if (aligned) *aligned = a.cell_ordered && MODE == 2 && align_origin >= 0;
    if (a.src_end <= a.src_begin) return 0;
    unsigned g = search_blocks(a, 128);
    bool sorted = a.tar_sorted_pos != nullptr;
    if (a.cell_ordered)
    {
        if (a.tar2_pos)
        {
            if (inner) SPH_LAUNCH(ctx, (k_relation_ordered<true, MODE, true>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count, align_origin);
            else SPH_LAUNCH(ctx, (k_relation_ordered<false, MODE, true>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count, align_origin);
        }
        else if (inner) SPH_LAUNCH(ctx, (k_relation_ordered<true, MODE, false>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count, align_origin);
        else SPH_LAUNCH(ctx, (k_relation_ordered<false, MODE, false>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count, align_origin);
        return 0;
    }
    if (inner && sorted) SPH_LAUNCH(ctx, (k_relation<true, true, MODE>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count);
    else if (inner) SPH_LAUNCH(ctx, (k_relation<true, false, MODE>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count);
    else if (sorted) SPH_LAUNCH(ctx, (k_relation<false, true, MODE>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count);
    else SPH_LAUNCH(ctx, (k_relation<false, false, MODE>), g, 128, 0, st, a, count, slice, index, cap, stride, max_count);
    return 0;
}

extern "C" int sphb200_relation_count(sphb200_context_t *ctx, const sphb200_search_t *search, sphb200_relation_t rel,
                                      uint64_t *required_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && search && rel.count && rel.slice_offset, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    u32 n_src = search->n_src;
    u32 nslices = (n_src + 31) / 32;
    if (n_src == 0)
    {
        SPH_CUDA(ctx, cudaMemsetAsync(rel.slice_offset, 0, sizeof(u32), st));
        if (required_host) *required_host = 0;
        return 0;
    }
    SearchArgs a;
    int rc = make_search(ctx, search, &a);
    if (rc) return rc;
    void *p;
    rc = sph_scratch(ctx, 3, ((size_t)nslices + 1) * sizeof(u32) + 64, &p);
    if (rc) return rc;
    u32 *slice_len = (u32 *)p;
    SPH_CUDA(ctx, cudaMemsetAsync(slice_len, 0, ((size_t)nslices + 1) * sizeof(u32), st));
    rc = launch_relation<0>(ctx, a, search->is_inner != 0, rel.count, slice_len, nullptr, 0, 0, nullptr, st);
    if (rc) return rc;
    rc = sph_scan_u32(ctx, slice_len, rel.slice_offset, (u64)nslices + 1, 0, st);
    if (rc) return rc;
    if (required_host)
    {
        SPH_CUDA(ctx, cudaMemcpyAsync(ctx->host_pinned, rel.slice_offset + nslices, sizeof(u32), cudaMemcpyDeviceToHost, st));
        SPH_CUDA(ctx, cudaStreamSynchronize(st));
        *required_host = *(u32 *)ctx->host_pinned;
    }
    return 0;
}

extern "C" int sphb200_relation_fill(sphb200_context_t *ctx, const sphb200_search_t *search, sphb200_relation_t rel, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && search && rel.count && rel.slice_offset && rel.index, "null pointer");
    if (search->n_src == 0) return 0;
    SearchArgs a;
    int rc = make_search(ctx, search, &a);
    if (rc) return rc;
    return launch_relation<1>(ctx, a, search->is_inner != 0, rel.count, rel.slice_offset, rel.index, rel.capacity, 0, nullptr,
                              (cudaStream_t)stream);
}

extern "C" int sphb200_relation_build_fixed(sphb200_context_t *ctx, const sphb200_search_t *search, sphb200_relation_t rel,
                                            uint32_t stride, uint32_t *max_count_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && search && rel.count && rel.slice_offset && rel.index, "null pointer");
    SPH_CHECK_ARG(ctx, stride > 0, "stride must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    u32 n_src = search->n_src;
    u64 nslices = ((u64)n_src + 31) / 32;
    if (rel.capacity < nslices * 32ull * stride)
    {
        snprintf(ctx->err, sizeof(ctx->err), "relation_build_fixed: capacity %llu < %llu", (unsigned long long)rel.capacity,
                 (unsigned long long)(nslices * 32ull * stride));
        return SPHB200_E_CAPACITY;
    }
    SPH_CHECK_ARG(ctx, nslices * 32ull * stride < (1ull << 32), "fixed-stride relation exceeds 2^32 entries");
    u32 *dmax = (u32 *)ctx->dev_scalars + 8;
    SPH_CUDA(ctx, cudaMemsetAsync(dmax, 0, sizeof(u32), st));
    if (n_src)
    {
        SearchArgs a;
        int rc = make_search(ctx, search, &a);
        if (rc) return rc;
        bool aligned = false;
        rc = launch_relation<2>(ctx, a, search->is_inner != 0, rel.count, rel.slice_offset, rel.index, rel.capacity, stride, dmax, st,
                                rel.bank_aligned ? (int)((u32)(rel.bank_aligned - 1) & 7u) : -1, &aligned);
        if (rc) return rc;
        if (rel.bank_aligned && !aligned) // generic (index-indirected) search: stand-alone layout pass
        {
            rc = bank_align(ctx, a, rel.count, rel.slice_offset, rel.index, stride, (u32)(rel.bank_aligned - 1) & 7u, st);
            if (rc) return rc;
        }
    }
    if (max_count_host)
    {
        SPH_CUDA(ctx, cudaMemcpyAsync((u32 *)ctx->host_pinned + 8, dmax, sizeof(u32), cudaMemcpyDeviceToHost, st));
        SPH_CUDA(ctx, cudaStreamSynchronize(st));
        *max_count_host = *((u32 *)ctx->host_pinned + 8);
    }
    return 0;
}

// SELL-32 (slot order) -> CSR by particle id
__global__ void __launch_bounds__(256) k_counts_by_id(const u32 *__restrict__ count, const u32 *__restrict__ order, u32 n, u32 *__restrict__ by_id)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) by_id[order ? order[t] : t] = count[t];
}
// rows of a bank-aligned relation leave in ascending target slot order (= the reference search order for cell-ordered
// bodies): repeated minimum selection, quadratic in the row length, which is fine for an export-only path
__global__ void __launch_bounds__(128)
    k_export_csr_sorted(const u32 *__restrict__ count, const u32 *__restrict__ slice_offset, const u32 *__restrict__ index,
                        const u32 *__restrict__ order, const u32 *__restrict__ tar_ids, u32 n, const u32 *__restrict__ particle_offset,
                        u32 *__restrict__ neighbor_index)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    u32 i = order ? order[t] : t;
    u64 src = (u64)slice_offset[t >> 5] + (t & 31u);
    u32 dst = particle_offset[i];
    u32 c = count[t];
    long long last = -1;
    for (u32 k = 0; k < c; ++k)
    {
        u32 best = 0xffffffffu;
        for (u32 q = 0; q < c; ++q)
        {
            u32 j = index[src + 32ull * q];
            if ((long long)j > last && j < best) best = j;
        }
        neighbor_index[dst + k] = tar_ids ? tar_ids[best] : best;
        last = best;
    }
}
__global__ void __launch_bounds__(128)
    k_export_csr(const u32 *__restrict__ count, const u32 *__restrict__ slice_offset, const u32 *__restrict__ index,
                 const u32 *__restrict__ order, const u32 *__restrict__ tar_ids, u32 n, const u32 *__restrict__ particle_offset,
                 u32 *__restrict__ neighbor_index)
{
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    u32 i = order ? order[t] : t;
    u64 src = (u64)slice_offset[t >> 5] + (t & 31u);
    u32 dst = particle_offset[i];
    u32 c = count[t];
    for (u32 k = 0; k < c; ++k)
    {
        u32 j = index[src + 32ull * k];
        neighbor_index[dst + k] = tar_ids ? tar_ids[j] : j;
    }
}

extern "C" int sphb200_relation_export_csr(sphb200_context_t *ctx, sphb200_relation_t rel, uint32_t n, const uint32_t *src_ids,
                                           const uint32_t *tar_ids, uint32_t *particle_offset, uint32_t *neighbor_index,
                                           uint64_t index_capacity, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && rel.count && rel.slice_offset && particle_offset, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const u32 *order = src_ids ? src_ids : rel.order;
    void *p;
    int rc = sph_scratch(ctx, 3, ((size_t)n + 1) * sizeof(u32) + 64, &p);
    if (rc) return rc;
    u32 *by_id = (u32 *)p;
    SPH_CUDA(ctx, cudaMemsetAsync(by_id + n, 0, sizeof(u32), st));
    if (n) SPH_LAUNCH(ctx, k_counts_by_id, sph_blocks(n, 256), 256, 0, st, rel.count, order, n, by_id);
    // particle_offset = exclusive scan of the per-particle counts over n+1 entries
    rc = sph_scan_u32(ctx, by_id, particle_offset, (u64)n + 1, 0, st);
    if (rc) return rc;
    if (n == 0 || !neighbor_index) return 0;
    SPH_CUDA(ctx, cudaMemcpyAsync(ctx->host_pinned, particle_offset + n, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SPH_CUDA(ctx, cudaStreamSynchronize(st));
    u32 total = *(u32 *)ctx->host_pinned;
    if (total > index_capacity)
    {
        snprintf(ctx->err, sizeof(ctx->err), "export_csr: need %u entries, capacity %llu", total,
                 (unsigned long long)index_capacity);
        return SPHB200_E_CAPACITY;
    }
    if (rel.bank_aligned)
        SPH_LAUNCH(ctx, k_export_csr_sorted, sph_blocks(n, 128), 128, 0, st, rel.count, rel.slice_offset, rel.index, order, tar_ids, n,
                   particle_offset, neighbor_index);
    else
        SPH_LAUNCH(ctx, k_export_csr, sph_blocks(n, 128), 128, 0, st, rel.count, rel.slice_offset, rel.index, order, tar_ids, n,
                   particle_offset, neighbor_index);
    return 0;
}

// =====================================================================================================
// periodic boundary: bounding (wrap) and image particles
// ref: particle_dynamics/general_dynamics/domian_bouding/domain_bounding.h:48-175, domain_bounding.cpp:18-65.
// The reference inserts, axis by axis, ghost ENTRIES (source index, translated position) into the cell-linked list;
// here the images become ghost PARTICLES stored cell ordered behind the real ones (the storage convention of the
// slab-decomposed runs), so every kernel reads them like any other neighbour.
// =====================================================================================================
struct DPeriodic
{
    float lower[3], upper[3], shift[3], cutoff;
    int axes;
};
static inline DPeriodic make_dperiodic(const sphb200_periodic_t *b)
{
    DPeriodic d;
    for (int k = 0; k < 3; ++k)
    {
        d.lower[k] = b->lower[k];
        d.upper[k] = b->upper[k];
        d.shift[k] = b->upper[k] - b->lower[k]; // periodic_translation_, in Real (domain_bounding.h:56-57)
    }
    d.cutoff = b->cutoff;
    d.axes = b->axes;
    return d;
}
__device__ __forceinline__ float &comp(float4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

__global__ void __launch_bounds__(256) k_periodic_bounding(DPeriodic b, float4 *__restrict__ pos, u32 n)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 x = pos[i];
    bool changed = false;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (b.axes >> k & 1)
        {
            float &c = comp(x, k);
            if (c < b.lower[k]) { c = __fadd_rn(c, b.shift[k]); changed = true; }      // checkLowerBound
            else if (c > b.upper[k]) { c = __fsub_rn(c, b.shift[k]); changed = true; } // checkUpperBound
        }
    if (changed) pos[i] = x;
}
extern "C" int sphb200_periodic_bounding(sphb200_context_t *ctx, const sphb200_periodic_t *box, sphb200_vec4_t *pos, uint32_t n,
                                         void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && box && (pos || n == 0), "null pointer");
    if (n) SPH_LAUNCH(ctx, k_periodic_bounding, sph_blocks(n, 256), 256, 0, stream, make_dperiodic(box), (float4 *)pos, n);
    return 0;
}

// per axis: bit 0 = image at +L (particle near the lower face), bit 1 = image at -L (near the upper face)
__device__ __forceinline__ int image_options(const DPeriodic &b, float c, int k)
{
    if (!(b.axes >> k & 1)) return 0;
    int o = 0;
    if (c > b.lower[k] && c < __fadd_rn(b.lower[k], b.cutoff)) o |= 1; // InsertListDataNearLowerBound
    if (c < b.upper[k] && c > __fsub_rn(b.upper[k], b.cutoff)) o |= 2; // InsertListDataNearUpperBound
    return o;
}
__device__ __forceinline__ int option_count(int o) { return 1 + (o & 1) + (o >> 1 & 1); }
// MODE 0: number of images per particle; MODE 1: write them at offset[i]
template <int MODE>
__global__ void __launch_bounds__(256)
    k_periodic_images(DPeriodic b, const float4 *__restrict__ pos, u32 n, u32 *__restrict__ count, const u32 *__restrict__ offset,
                      float4 *__restrict__ image_pos, u32 *__restrict__ image_src)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 x = pos[i];
    const int ox = image_options(b, x.x, 0), oy = image_options(b, x.y, 1), oz = image_options(b, x.z, 2);
    if (MODE == 0)
    {
        count[i] = (u32)(option_count(ox) * option_count(oy) * option_count(oz) - 1);
        return;
    }
    if ((ox | oy | oz) == 0) return;
    u32 k = offset[i];
    // choice 0: stay, 1: +L, 2: -L ; the all-stay combination is the particle itself
    for (int cx = 0; cx < 3; ++cx)
    {
        if (cx && !(ox >> (cx - 1) & 1)) continue;
        for (int cy = 0; cy < 3; ++cy)
        {
            if (cy && !(oy >> (cy - 1) & 1)) continue;
            for (int cz = 0; cz < 3; ++cz)
            {
                if (cz && !(oz >> (cz - 1) & 1)) continue;
                if ((cx | cy | cz) == 0) continue;
                float4 y = x;
                if (cx) y.x = cx == 1 ? __fadd_rn(x.x, b.shift[0]) : __fsub_rn(x.x, b.shift[0]);
                if (cy) y.y = cy == 1 ? __fadd_rn(x.y, b.shift[1]) : __fsub_rn(x.y, b.shift[1]);
                if (cz) y.z = cz == 1 ? __fadd_rn(x.z, b.shift[2]) : __fsub_rn(x.z, b.shift[2]);
                image_pos[k] = y;
                image_src[k] = i;
                ++k;
            }
        }
    }
}
extern "C" int sphb200_periodic_images(sphb200_context_t *ctx, const sphb200_periodic_t *box, const sphb200_vec4_t *pos, uint32_t n,
                                       sphb200_vec4_t *image_pos, uint32_t *image_src, uint32_t capacity, uint32_t *count_host,
                                       void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && box && count_host && (pos || n == 0), "null pointer");
    *count_host = 0;
    if (n == 0 || box->axes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    void *p;
    int rc = sph_scratch(ctx, 2, 2 * ((size_t)n + 1) * sizeof(u32) + 64, &p);
    if (rc) return rc;
    u32 *cnt = (u32 *)p, *off = cnt + n + 1;
    DPeriodic b = make_dperiodic(box);
    SPH_CUDA(ctx, cudaMemsetAsync(cnt + n, 0, sizeof(u32), st));
    SPH_LAUNCH(ctx, k_periodic_images<0>, sph_blocks(n, 256), 256, 0, st, b, (const float4 *)pos, n, cnt, nullptr, nullptr, nullptr);
    rc = sph_scan_u32(ctx, cnt, off, (u64)n + 1, 0, st);
    if (rc) return rc;
    SPH_CUDA(ctx, cudaMemcpyAsync(ctx->host_pinned, off + n, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SPH_CUDA(ctx, cudaStreamSynchronize(st));
    const u32 total = *(u32 *)ctx->host_pinned;
    *count_host = total;
    if (total > capacity)
    {
        snprintf(ctx->err, sizeof(ctx->err), "periodic_images: %u images, capacity %u", total, capacity);
        return SPHB200_E_CAPACITY;
    }
    if (total == 0) return 0;
    SPH_CHECK_ARG(ctx, image_pos && image_src, "null output");
    SPH_LAUNCH(ctx, k_periodic_images<1>, sph_blocks(n, 256), 256, 0, st, b, (const float4 *)pos, n, nullptr, off, (float4 *)image_pos,
               image_src);
    return 0;
}

// word-granular copy of a slice of every ghost element from its source element
__global__ void __launch_bounds__(256)
    k_ghost_copy(u32 *__restrict__ array, u32 elem_words, u32 offset_words, u32 copy_words, const u32 *__restrict__ ghost_src,
                 u32 n_real, u64 total_words)
{
    u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= total_words) return;
    u32 g = (u32)(w / copy_words), k = (u32)(w % copy_words);
    array[((u64)n_real + g) * elem_words + offset_words + k] = array[(u64)ghost_src[g] * elem_words + offset_words + k];
}
__global__ void __launch_bounds__(256)
    k_ghost_copy16(uint4 *__restrict__ array, u32 elem_q, u32 offset_q, const u32 *__restrict__ ghost_src, u32 n_real, u32 n_ghost)
{
    u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_ghost) return;
    array[((u64)n_real + g) * elem_q + offset_q] = array[(u64)ghost_src[g] * elem_q + offset_q];
}
extern "C" int sphb200_ghost_copy(sphb200_context_t *ctx, void *array, uint32_t elem_bytes, uint32_t offset_bytes, uint32_t copy_bytes,
                                  const uint32_t *ghost_src, uint32_t n_real, uint32_t n_ghost, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((array && ghost_src) || n_ghost == 0), "null pointer");
    SPH_CHECK_ARG(ctx, elem_bytes % 4 == 0 && offset_bytes % 4 == 0 && copy_bytes % 4 == 0 && copy_bytes > 0 &&
                           offset_bytes + copy_bytes <= elem_bytes,
                  "sizes must be multiples of 4 with offset + copy <= element");
    if (n_ghost == 0) return 0;
    if (copy_bytes == 16 && elem_bytes % 16 == 0 && offset_bytes % 16 == 0)
        SPH_LAUNCH(ctx, k_ghost_copy16, sph_blocks(n_ghost, 256), 256, 0, stream, (uint4 *)array, elem_bytes / 16, offset_bytes / 16,
                   ghost_src, n_real, n_ghost);
    else
    {
        u64 total = (u64)n_ghost * (copy_bytes / 4);
        SPH_LAUNCH(ctx, k_ghost_copy, sph_blocks(total, 256), 256, 0, stream, (u32 *)array, elem_bytes / 4, offset_bytes / 4,
                   copy_bytes / 4, ghost_src, n_real, total);
    }
    return 0;
}

// =====================================================================================================
// slab decomposition: which own particles does a neighbour rank need? (slab_decomposition.h)
// =====================================================================================================
// Slot i of [begin, begin + n) goes on the LEFT list if its x cell plane is <= plane_left (the first own plane and
// whatever moved below it), on the RIGHT list if it is >= plane_right (the last own plane and above); plane < 0 switches
// a side off. Lists are appended with warp-aggregated atomics (order irrelevant: the receiver brings everything into
// cell order, in-cell by ReferenceID). counts[0], counts[1] must be zero on entry.
__global__ void __launch_bounds__(256)
    k_slab_select(DMesh m, const float4 *__restrict__ pos, u32 begin, u32 n, int plane_left, int plane_right,
                  u32 *__restrict__ left_idx, u32 *__restrict__ right_idx, u32 *__restrict__ counts)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    bool l = false, r = false;
    if (k < n)
    {
        const int cx = cell_coord(pos[begin + k].x, m.lx, m.spacing, m.cx);
        l = plane_left >= 0 && cx <= plane_left;
        r = plane_right >= 0 && cx >= plane_right;
    }
    const u32 lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
    const u32 bl = __ballot_sync(0xffffffffu, l), br = __ballot_sync(0xffffffffu, r);
    u32 base_l = 0, base_r = 0;
    if (lane == 0)
    {
        if (bl) base_l = atomicAdd(counts, __popc(bl));
        if (br) base_r = atomicAdd(counts + 1, __popc(br));
    }
    base_l = __shfl_sync(0xffffffffu, base_l, 0);
    base_r = __shfl_sync(0xffffffffu, base_r, 0);
    if (l) left_idx[base_l + __popc(bl & below)] = begin + k;
    if (r) right_idx[base_r + __popc(br & below)] = begin + k;
}

extern "C" int sphb200_slab_select(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos, uint32_t begin,
                                   uint32_t n, int plane_left, int plane_right, uint32_t *left_idx, uint32_t *right_idx,
                                   uint32_t *counts, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && mesh && counts && (n == 0 || (pos && left_idx && right_idx)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPH_CUDA(ctx, cudaMemsetAsync(counts, 0, 2 * sizeof(u32), st));
    if (n == 0) return 0;
    SPH_LAUNCH(ctx, k_slab_select, sph_blocks(n, 256), 256, 0, st, make_dmesh(mesh), (const float4 *)pos, begin, n, plane_left,
               plane_right, left_idx, right_idx, counts);
    return 0;
}
