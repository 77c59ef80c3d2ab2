// common.cuh — context, launch bookkeeping and device helpers shared by the libsphb200 translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/sphb200.h"

typedef uint32_t u32;
typedef uint64_t u64;

struct sphb200_context
{
    int device;
    char err[512];
    u64 launches;
    // grow-on-demand scratch arenas (sort double buffers, histograms, cell ranks, ...)
    void *scratch[4];
    size_t scratch_bytes[4];
    // small pinned host block + device scalar block for value-returning calls
    void *host_pinned; // 256 B
    void *dev_scalars; // 256 B
    // slab-decomposed runs: NCCL communicator (comm.cu), null for single-GPU use
    void *comm;
    int rank, nranks;
    int ring;      // 1: the slab chain is closed (periodic along x): rank 0 and rank nranks-1 are neighbours
    int self_comm; // 1: communicator of ONE rank without NCCL (a ring of one slab: exchanges are device copies)
    // peer mailboxes (comm.cu, sphb200_comm_mailbox_*): migrants and boundary planes are written straight into the
    // neighbour's memory over NVLink (CUDA IPC mapping) by the kernel that gathers them; no host-known message sizes
    void *mailbox;              // this rank's boxes: [parity 0/1][from-left, from-right], each `mailbox_box_bytes`
    void *peer_mailbox[2];      // the left / right neighbour's boxes mapped into this process (nullptr: no neighbour)
    int peer_mapped[2];         // 1: peer_mailbox[s] came from cudaIpcOpenMemHandle (to be closed)
    size_t mailbox_box_bytes;   // bytes of one box of THIS rank (64-byte header + payload)
    size_t peer_box_bytes[2];   // ... of the neighbours' boxes
    unsigned *mailbox_dev;      // device words: [0..1] block tickets of the two push launches, [2] status flags
};

#define SPH_CHECK_ARG(ctx, cond, msg)                                                              \
    do                                                                                             \
    {                                                                                              \
        if (!(cond))                                                                               \
        {                                                                                          \
            if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s", __func__, msg);            \
            return SPHB200_E_INVALID;                                                              \
        }                                                                                          \
    } while (0)

#define SPH_CUDA(ctx, expr)                                                                        \
    do                                                                                             \
    {                                                                                              \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
        {                                                                                          \
            if (ctx)                                                                               \
                snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s -> %s", __func__, #expr,         \
                         cudaGetErrorString(e__));                                                 \
            return (int)e__;                                                                       \
        }                                                                                          \
    } while (0)

// kernel launch with bookkeeping: counts the launch and surfaces launch-configuration errors
#define SPH_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                                    \
    do                                                                                             \
    {                                                                                              \
        kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                  \
        (ctx)->launches++;                                                                         \
        SPH_CUDA(ctx, cudaGetLastError());                                                         \
    } while (0)

void sph_count_allocation(); // primitives.cu
static inline int sph_scratch(sphb200_context *ctx, int slot, size_t bytes, void **out)
{
    if (ctx->scratch_bytes[slot] < bytes)
    {
        sph_count_allocation();
        if (ctx->scratch[slot]) SPH_CUDA(ctx, cudaFree(ctx->scratch[slot]));
        ctx->scratch[slot] = nullptr;
        ctx->scratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4 + 256;
        SPH_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return 0;
}

static inline unsigned sph_blocks(u64 n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// ---------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------
struct DMesh
{
    float lx, ly, lz, spacing;
    int cx, cy, cz;
};
static inline DMesh make_dmesh(const sphb200_mesh_t *m)
{
    DMesh d;
    d.lx = m->lower[0]; d.ly = m->lower[1]; d.lz = m->lower[2];
    d.spacing = m->spacing;
    d.cx = m->cells[0]; d.cy = m->cells[1]; d.cz = m->cells[2];
    return d;
}

// Mesh::CellIndexFromPosition, base_mesh.hxx:9-15 — IEEE subtract, IEEE divide, floor, clamp; bit-identical
// to the oracle (no reciprocal multiply, no contraction).
__device__ __forceinline__ int cell_coord(float x, float lower, float spacing, int cells)
{
    float t = __fsub_rn(x, lower);
    float u = __fdiv_rn(t, spacing);
    int k = (int)floorf(u);
    k = max(k, 0);
    k = min(k, cells - 1);
    return k;
}
// Mesh::transferMeshIndexTo1D, base_mesh.hxx:73-78 (z fastest)
__device__ __forceinline__ u32 cell_linear(const DMesh &m, int a, int b, int c)
{
    return (u32)a * (u32)m.cy * (u32)m.cz + (u32)b * (u32)m.cz + (u32)c;
}
// Mesh::MortonCode, base_mesh.hxx:90-99 (10 bits per axis)
__device__ __forceinline__ u32 morton_spread(u32 x)
{
    x &= 0x3ff;
    x = (x | x << 16) & 0x30000ff;
    x = (x | x << 8) & 0x300f00f;
    x = (x | x << 4) & 0x30c30c3;
    x = (x | x << 2) & 0x9249249;
    return x;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ u32 warp_max_u32(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// internal cross-TU entry points
int sph_scan_u32(sphb200_context *ctx, const u32 *in, u32 *out, u64 n, int scratch_slot, cudaStream_t stream);
