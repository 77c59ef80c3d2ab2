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
__global__ void __launch_bounds__(256)
    k_cell_count(DMesh m, const float4 *__restrict__ pos, u32 n, u32 *__restrict__ counts, u32 *__restrict__ cell_of,
                 u32 *__restrict__ rank)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 x = pos[i];
    u32 c = cell_linear(m, cell_coord(x.x, m.lx, m.spacing, m.cx), cell_coord(x.y, m.ly, m.spacing, m.cy),
                        cell_coord(x.z, m.lz, m.spacing, m.cz));
    cell_of[i] = c;
    rank[i] = atomicAdd(&counts[c], 1u);
}
__global__ void __launch_bounds__(256)
    k_cell_fill(const u32 *__restrict__ cell_offset, const u32 *__restrict__ cell_of, const u32 *__restrict__ rank, u32 n,
                u32 *__restrict__ list)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[cell_offset[cell_of[i]] + rank[i]] = i;
}
// deterministic in-cell order: the final slot of i is the number of cell-mates with a smaller index
__global__ void __launch_bounds__(256)
    k_cell_order(const u32 *__restrict__ cell_offset, const u32 *__restrict__ cell_of, const u32 *__restrict__ unordered,
                 u32 n, u32 *__restrict__ particle_index)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 c = cell_of[i];
    u32 b = cell_offset[c], e = cell_offset[c + 1];
    u32 smaller = 0;
    for (u32 k = b; k < e; ++k) smaller += (unordered[k] < i);
    particle_index[b + smaller] = i;
}

extern "C" int sphb200_cell_list_build(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos,
                                       uint32_t n, sphb200_cell_list_t list, void *stream)
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
    if (n) SPH_LAUNCH(ctx, k_cell_count, sph_blocks(n, 256), 256, 0, st, m, (const float4 *)pos, n, counts, cell_of, rank);
    rc = sph_scan_u32(ctx, counts, list.cell_offset, cells + 1, 0, st);
    if (rc) return rc;
    if (n)
    {
        SPH_LAUNCH(ctx, k_cell_fill, sph_blocks(n, 256), 256, 0, st, list.cell_offset, cell_of, rank, n, unordered);
        SPH_LAUNCH(ctx, k_cell_order, sph_blocks(n, 256), 256, 0, st, list.cell_offset, cell_of, unordered, n,
                   list.particle_index);
    }
    return 0;
}

// =====================================================================================================
// relations (neighbour lists)
// =====================================================================================================
struct SearchArgs
{
    DMesh m;
    const float4 *src_pos;
    const float4 *tar_pos;
    const u32 *cell_offset;
    const u32 *particle_index;
    u32 n_src;
    float inv_h, ks2;
    int depth;
};

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

// enumerate candidates in the reference order: cells x -> y -> z (mesh_iterators.hpp:18-27); for fixed (x, y)
// the z cells are contiguous in the linear index, so each (x, y) column is one run of the particle list.
template <bool INNER, class F> __device__ __forceinline__ void for_each_neighbor(const SearchArgs &a, u32 i, float4 xi, F f)
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
            for (u32 k = b; k < e; ++k)
            {
                u32 j = a.particle_index[k];
                if (INNER && j == i) continue;
                float4 xj = a.tar_pos[j];
                if (within(xi, xj, a.inv_h, a.ks2)) f(j);
            }
        }
}

template <bool INNER>
__global__ void __launch_bounds__(128) k_relation_count(SearchArgs a, u32 *__restrict__ count, u32 *__restrict__ slice_len)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u32 c = 0;
    if (i < a.n_src)
    {
        float4 xi = a.src_pos[i];
        for_each_neighbor<INNER>(a, i, xi, [&](u32) { ++c; });
        count[i] = c;
    }
    u32 mx = warp_max_u32(c);
    if ((threadIdx.x & 31) == 0 && (i >> 5) <= ((a.n_src - 1) >> 5)) slice_len[i >> 5] = mx * 32u;
}

template <bool INNER>
__global__ void __launch_bounds__(128)
    k_relation_fill(SearchArgs a, const u32 *__restrict__ slice_offset, u32 *__restrict__ index, u64 capacity)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_src) return;
    float4 xi = a.src_pos[i];
    u64 pos = (u64)slice_offset[i >> 5] + (i & 31u);
    for_each_neighbor<INNER>(a, i, xi, [&](u32 j) {
        if (pos < capacity) index[pos] = j;
        pos += 32;
    });
}

static int make_search(sphb200_context *ctx, const sphb200_mesh_t *mesh, const sphb200_kernel_t *kernel,
                       const sphb200_vec4_t *src_pos, u32 n_src, const sphb200_vec4_t *tar_pos, sphb200_cell_list_t list,
                       int depth, SearchArgs *a)
{
    a->m = make_dmesh(mesh);
    a->src_pos = (const float4 *)src_pos;
    a->tar_pos = (const float4 *)tar_pos;
    a->cell_offset = list.cell_offset;
    a->particle_index = list.particle_index;
    a->n_src = n_src;
    a->inv_h = 1.0f / kernel->h; // inv_h_ = 1 / max(src_h, tar_h), neighbor_method.hpp:73-76
    a->ks2 = kernel->kernel_size * kernel->kernel_size;
    a->depth = depth;
    return 0;
}

extern "C" int sphb200_relation_count(sphb200_context_t *ctx, const sphb200_mesh_t *tar_mesh, const sphb200_kernel_t *kernel,
                                      const sphb200_vec4_t *src_pos, uint32_t n_src, const sphb200_vec4_t *tar_pos,
                                      sphb200_cell_list_t tar_list, int is_inner, int search_depth, sphb200_relation_t rel,
                                      uint64_t *required_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && tar_mesh && kernel && rel.count && rel.slice_offset, "null pointer");
    SPH_CHECK_ARG(ctx, search_depth >= 1 && search_depth <= 4, "search depth out of range");
    cudaStream_t st = (cudaStream_t)stream;
    u32 nslices = (n_src + 31) / 32;
    if (n_src == 0)
    {
        SPH_CUDA(ctx, cudaMemsetAsync(rel.slice_offset, 0, sizeof(u32), st));
        if (required_host) *required_host = 0;
        return 0;
    }
    SPH_CHECK_ARG(ctx, src_pos && tar_pos && tar_list.cell_offset && tar_list.particle_index, "null pointer");
    SearchArgs a;
    make_search(ctx, tar_mesh, kernel, src_pos, n_src, tar_pos, tar_list, search_depth, &a);
    void *p;
    int rc = sph_scratch(ctx, 3, ((size_t)nslices + 1) * sizeof(u32) + 64, &p);
    if (rc) return rc;
    u32 *slice_len = (u32 *)p;
    SPH_CUDA(ctx, cudaMemsetAsync(slice_len + nslices, 0, sizeof(u32), st));
    if (is_inner)
        SPH_LAUNCH(ctx, k_relation_count<true>, sph_blocks(n_src, 128), 128, 0, st, a, rel.count, slice_len);
    else
        SPH_LAUNCH(ctx, k_relation_count<false>, sph_blocks(n_src, 128), 128, 0, st, a, rel.count, slice_len);
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

extern "C" int sphb200_relation_fill(sphb200_context_t *ctx, const sphb200_mesh_t *tar_mesh, const sphb200_kernel_t *kernel,
                                     const sphb200_vec4_t *src_pos, uint32_t n_src, const sphb200_vec4_t *tar_pos,
                                     sphb200_cell_list_t tar_list, int is_inner, int search_depth, sphb200_relation_t rel,
                                     void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && tar_mesh && kernel && rel.count && rel.slice_offset && rel.index, "null pointer");
    SPH_CHECK_ARG(ctx, search_depth >= 1 && search_depth <= 4, "search depth out of range");
    if (n_src == 0) return 0;
    SearchArgs a;
    make_search(ctx, tar_mesh, kernel, src_pos, n_src, tar_pos, tar_list, search_depth, &a);
    cudaStream_t st = (cudaStream_t)stream;
    if (is_inner)
        SPH_LAUNCH(ctx, k_relation_fill<true>, sph_blocks(n_src, 128), 128, 0, st, a, rel.slice_offset, rel.index, rel.capacity);
    else
        SPH_LAUNCH(ctx, k_relation_fill<false>, sph_blocks(n_src, 128), 128, 0, st, a, rel.slice_offset, rel.index, rel.capacity);
    return 0;
}

// SELL-32 -> CSR
__global__ void __launch_bounds__(128)
    k_export_csr(const u32 *__restrict__ count, const u32 *__restrict__ slice_offset, const u32 *__restrict__ index, u32 n,
                 const u32 *__restrict__ particle_offset, u32 *__restrict__ neighbor_index)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 src = (u64)slice_offset[i >> 5] + (i & 31u);
    u32 dst = particle_offset[i];
    u32 c = count[i];
    for (u32 k = 0; k < c; ++k) neighbor_index[dst + k] = index[src + 32ull * k];
}

extern "C" int sphb200_relation_export_csr(sphb200_context_t *ctx, sphb200_relation_t rel, uint32_t n, uint32_t *particle_offset,
                                           uint32_t *neighbor_index, uint64_t index_capacity, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && rel.count && rel.slice_offset && particle_offset, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // particle_offset = exclusive scan of count over n+1 entries (the last input is unused)
    int rc = sph_scan_u32(ctx, rel.count, particle_offset, (u64)n + 1, 0, st);
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
    SPH_LAUNCH(ctx, k_export_csr, sph_blocks(n, 128), 128, 0, st, rel.count, rel.slice_offset, rel.index, n, particle_offset,
               neighbor_index);
    return 0;
}
