// primitives.cu — context/memory entry points and the data-parallel primitives of the hot path:
// exclusive scan, stable LSD radix sort by key, fused multi-array gather, fills and Vecd layout conversion.
// Replaces (not ports) src_sycl/shared/common/algorithm_primitive_sycl.{h,hpp,cpp} and the USM helpers in
// src_sycl/shared/particle_dynamics/implementation_sycl.h:99-160.
#include "common.cuh"

// =====================================================================================================
// context + memory
// =====================================================================================================
extern "C" int sphb200_version(void) { return SPHB200_VERSION; }

extern "C" int sphb200_context_create(int device, sphb200_context_t **out)
{
    if (!out) return SPHB200_E_INVALID;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return (int)e;
    if (device < 0 || device >= count) return SPHB200_E_INVALID;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    sphb200_context *ctx = new sphb200_context();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->nranks = 1;
    e = cudaMallocHost(&ctx->host_pinned, 256);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dev_scalars, 256);
    if (e != cudaSuccess)
    {
        delete ctx;
        return (int)e;
    }
    *out = ctx;
    return SPHB200_OK;
}

extern "C" int sphb200_context_destroy(sphb200_context_t *ctx)
{
    if (!ctx) return SPHB200_E_INVALID;
    if (ctx->comm) sphb200_comm_destroy(ctx);
    for (int s = 0; s < 4; ++s)
        if (ctx->scratch[s]) cudaFree(ctx->scratch[s]);
    if (ctx->host_pinned) cudaFreeHost(ctx->host_pinned);
    if (ctx->dev_scalars) cudaFree(ctx->dev_scalars);
    delete ctx;
    return SPHB200_OK;
}

extern "C" const char *sphb200_last_error_string(const sphb200_context_t *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t sphb200_launch_count(const sphb200_context_t *ctx) { return ctx ? ctx->launches : 0; }

// device allocations made so far by this process through the library (arrays of the host layer + scratch arenas): a
// steady-state loop must not allocate — cudaMalloc / cudaFree synchronise the device and cost milliseconds
static unsigned long long g_device_allocations = 0;
void sph_count_allocation() { ++g_device_allocations; }
extern "C" uint64_t sphb200_device_allocation_count(void) { return g_device_allocations; }
extern "C" int sphb200_malloc_device(void **ptr, size_t bytes)
{
    ++g_device_allocations;
    return (int)cudaMalloc(ptr, bytes ? bytes : 1);
}
extern "C" int sphb200_malloc_host(void **ptr, size_t bytes) { return (int)cudaMallocHost(ptr, bytes ? bytes : 1); }
extern "C" int sphb200_free_device(void *ptr) { return (int)cudaFree(ptr); }
extern "C" int sphb200_free_host(void *ptr) { return (int)cudaFreeHost(ptr); }
extern "C" int sphb200_copy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
extern "C" int sphb200_copy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
extern "C" int sphb200_copy_d2d(void *dst, const void *src, size_t bytes, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
extern "C" int sphb200_stream_sync(void *stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }
// a second, non-blocking stream and events: host <-> device transfers that overlap the dynamics running on the
// caller's main stream (the SYCL build has one in-order queue; overlap is an extension of this library)
extern "C" int sphb200_stream_create(void **stream)
{
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    *stream = (void *)s;
    return (int)e;
}
extern "C" int sphb200_stream_create_with_priority(void **stream, int high_priority)
{
    int lo = 0, hi = 0; // numerically lower == higher priority
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t s = nullptr;
    e = cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo);
    *stream = (void *)s;
    return (int)e;
}
extern "C" int sphb200_stream_destroy(void *stream) { return (int)cudaStreamDestroy((cudaStream_t)stream); }
extern "C" int sphb200_event_create(void **event)
{
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    *event = (void *)ev;
    return (int)e;
}
extern "C" int sphb200_event_destroy(void *event) { return (int)cudaEventDestroy((cudaEvent_t)event); }
extern "C" int sphb200_event_record(void *event, void *stream) { return (int)cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream); }
extern "C" int sphb200_stream_wait_event(void *stream, void *event) { return (int)cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0); }

// =====================================================================================================
// fills, layout conversion
// =====================================================================================================
template <class T> __global__ void k_fill(T *dst, T v, u64 n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = v;
}
static inline unsigned fill_grid(u64 n) { u64 b = (n + 255) / 256; return (unsigned)(b < 148u * 16u ? b : 148u * 16u); }

extern "C" int sphb200_fill_u32(sphb200_context_t *ctx, uint32_t *dst, uint32_t value, uint64_t n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (dst || n == 0), "null pointer");
    if (n == 0) return 0;
    SPH_LAUNCH(ctx, k_fill<u32>, fill_grid(n), 256, 0, stream, dst, value, n);
    return 0;
}
extern "C" int sphb200_fill_f32(sphb200_context_t *ctx, float *dst, float value, uint64_t n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (dst || n == 0), "null pointer");
    if (n == 0) return 0;
    SPH_LAUNCH(ctx, k_fill<float>, fill_grid(n), 256, 0, stream, dst, value, n);
    return 0;
}

__global__ void k_iota(u32 *dst, u64 n)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = (u32)i;
}
extern "C" int sphb200_iota_u32(sphb200_context_t *ctx, uint32_t *dst, uint64_t n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (dst || n == 0), "null pointer");
    if (n == 0) return 0;
    SPH_LAUNCH(ctx, k_iota, fill_grid(n), 256, 0, stream, dst, n);
    return 0;
}

__global__ void k_vec3_to_vec4(float4 *dst, const float *src, u32 n)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float4(src[3 * (u64)i], src[3 * (u64)i + 1], src[3 * (u64)i + 2], 0.f);
}
__global__ void k_vec4_to_vec3(float *dst, const float4 *src, u32 n)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        float4 v = src[i];
        dst[3 * (u64)i] = v.x; dst[3 * (u64)i + 1] = v.y; dst[3 * (u64)i + 2] = v.z;
    }
}
__global__ void k_pack_posvol(float4 *posvol, const float4 *pos, const float *vol, u32 n)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
    {
        float4 p = pos[i];
        p.w = vol[i];
        posvol[i] = p;
    }
}
__global__ void __launch_bounds__(256)
    k_pack_records(u32 n, const float4 *__restrict__ pos, const float *__restrict__ vol, const float *__restrict__ vol_ref,
                   const float4 *__restrict__ vel, float4 *__restrict__ posvol, float4 *__restrict__ posvolref, float4 *__restrict__ posvolvel)
{
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    if (posvol || posvolvel)
    {
        p.w = vol[i];
        if (posvol) posvol[i] = p;
        if (posvolvel)
        {
            posvolvel[2ull * i] = p;
            float4 v = vel[i];
            v.w = 0.f;
            posvolvel[2ull * i + 1] = v;
        }
    }
    if (posvolref)
    {
        p.w = vol_ref[i];
        posvolref[i] = p;
    }
}
extern "C" int sphb200_pack_records(sphb200_context_t *ctx, uint32_t n, const sphb200_vec4_t *pos, const float *vol, const float *vol_ref,
                                    const sphb200_vec4_t *vel, sphb200_vec4_t *posvol, sphb200_vec4_t *posvolref, void *posvolvel,
                                    void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (pos || n == 0), "null pointer");
    SPH_CHECK_ARG(ctx, n == 0 || !(posvol || posvolvel) || vol, "posvol / posvolvel need vol");
    SPH_CHECK_ARG(ctx, n == 0 || !posvolref || vol_ref, "posvolref needs vol_ref");
    SPH_CHECK_ARG(ctx, n == 0 || !posvolvel || vel, "posvolvel needs vel");
    if (n) SPH_LAUNCH(ctx, k_pack_records, sph_blocks(n, 256), 256, 0, stream, n, (const float4 *)pos, vol, vol_ref,
                      (const float4 *)vel, (float4 *)posvol, (float4 *)posvolref, (float4 *)posvolvel);
    return 0;
}

extern "C" int sphb200_vec3_to_vec4(sphb200_context_t *ctx, sphb200_vec4_t *dst, const float *src3, uint32_t n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((dst && src3) || n == 0), "null pointer");
    if (n) SPH_LAUNCH(ctx, k_vec3_to_vec4, sph_blocks(n, 256), 256, 0, stream, (float4 *)dst, src3, n);
    return 0;
}
extern "C" int sphb200_vec4_to_vec3(sphb200_context_t *ctx, float *dst3, const sphb200_vec4_t *src, uint32_t n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((dst3 && src) || n == 0), "null pointer");
    if (n) SPH_LAUNCH(ctx, k_vec4_to_vec3, sph_blocks(n, 256), 256, 0, stream, dst3, (const float4 *)src, n);
    return 0;
}
extern "C" int sphb200_pack_posvol(sphb200_context_t *ctx, sphb200_vec4_t *posvol, const sphb200_vec4_t *pos,
                                   const float *vol, uint32_t n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((posvol && pos && vol) || n == 0), "null pointer");
    if (n) SPH_LAUNCH(ctx, k_pack_posvol, sph_blocks(n, 256), 256, 0, stream, (float4 *)posvol, (const float4 *)pos, vol, n);
    return 0;
}

// =====================================================================================================
// exclusive scan (u32): tile reduce -> recursive scan of tile sums -> tile scan with carry-in.
// Tiles of 4096 (512 threads x 8 striped items) keep every global access coalesced.
// =====================================================================================================
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u32 block_excl_scan_512(u32 v, u32 *ws /*[34]*/, u32 &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) ws[wid] = incl;
    __syncthreads();
    if (wid == 0)
    {
        u32 w = lane < (SCAN_THREADS / 32) ? ws[lane] : 0;
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            u32 t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane == (SCAN_THREADS / 32) - 1) ws[33] = wi;
        __syncwarp();
        if (lane < (SCAN_THREADS / 32)) ws[lane] = wi - w;
    }
    __syncthreads();
    u32 excl = incl - v + ws[wid];
    total = ws[33];
    __syncthreads();
    return excl;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const u32 *__restrict__ in, u32 *__restrict__ sums, u64 n)
{
    __shared__ u32 ws[SCAN_THREADS / 32];
    u64 base = (u64)blockIdx.x * SCAN_TILE;
    u32 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        u64 idx = base + (u64)k * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        u32 t = threadIdx.x < SCAN_THREADS / 32 ? ws[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan_apply(const u32 *in, u32 *out, const u32 *__restrict__ carry, u64 n)
{
    __shared__ u32 ws[34];
    u64 base = (u64)blockIdx.x * SCAN_TILE;
    u32 running = carry ? carry[blockIdx.x] : 0;
#pragma unroll 1
    for (int k = 0; k < SCAN_ITEMS; ++k)
    {
        u64 idx = base + (u64)k * SCAN_THREADS + threadIdx.x;
        u32 v = idx < n ? in[idx] : 0;
        u32 total;
        u32 excl = block_excl_scan_512(v, ws, total);
        if (idx < n) out[idx] = excl + running;
        running += total;
    }
}

int sph_scan_u32(sphb200_context *ctx, const u32 *in, u32 *out, u64 n, int slot, cudaStream_t stream)
{
    if (n == 0) return 0;
    // level sizes
    u64 sizes[6];
    int levels = 0;
    u64 m = n;
    size_t total = 0;
    while (m > SCAN_TILE)
    {
        m = (m + SCAN_TILE - 1) / SCAN_TILE;
        sizes[levels++] = m;
        total += m;
    }
    u32 *sums = nullptr;
    if (levels)
    {
        void *p;
        int rc = sph_scratch(ctx, slot, total * sizeof(u32), &p);
        if (rc) return rc;
        sums = (u32 *)p;
    }
    // reduce up
    const u32 *src = in;
    u64 cur = n;
    u32 *lvl[6];
    u32 *ptr = sums;
    for (int l = 0; l < levels; ++l)
    {
        lvl[l] = ptr;
        SPH_LAUNCH(ctx, k_scan_reduce, (unsigned)sizes[l], SCAN_THREADS, 0, stream, src, lvl[l], cur);
        src = lvl[l];
        cur = sizes[l];
        ptr += sizes[l];
    }
    // scan down
    for (int l = levels - 1; l >= 0; --l)
    {
        const u32 *carry = (l == levels - 1) ? nullptr : lvl[l + 1];
        unsigned grid = (unsigned)((sizes[l] + SCAN_TILE - 1) / SCAN_TILE);
        SPH_LAUNCH(ctx, k_scan_apply, grid, SCAN_THREADS, 0, stream, lvl[l], lvl[l], carry, sizes[l]);
    }
    unsigned grid = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    SPH_LAUNCH(ctx, k_scan_apply, grid, SCAN_THREADS, 0, stream, in, out, levels ? lvl[0] : (const u32 *)nullptr, n);
    return 0;
}

extern "C" int sphb200_exclusive_scan_u32(sphb200_context_t *ctx, const uint32_t *in, uint32_t *out, uint64_t n,
                                          uint32_t *last_host, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((in && out) || n == 0), "null pointer");
    if (n == 0)
    {
        if (last_host) *last_host = 0;
        return 0;
    }
    int rc = sph_scan_u32(ctx, in, out, n, 0, (cudaStream_t)stream);
    if (rc) return rc;
    if (last_host)
    {
        SPH_CUDA(ctx, cudaMemcpyAsync(ctx->host_pinned, out + (n - 1), sizeof(u32), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        SPH_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
        *last_host = *(u32 *)ctx->host_pinned;
    }
    return 0;
}

// =====================================================================================================
// stable LSD radix sort of (key, value) pairs, 8-bit digits.
// per pass: tile histograms [digit][tile] -> exclusive scan -> stable scatter (warp match + per-warp counts).
// =====================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS)
    k_rs_hist(const u32 *__restrict__ keys, u64 n, int shift, u32 *__restrict__ hist, u32 ntiles)
{
    __shared__ u32 cnt[256];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k)
    {
        u64 idx = base + (u64)k * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&cnt[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(u64)threadIdx.x * ntiles + blockIdx.x] = cnt[threadIdx.x];
}

// Stable scatter of one tile (RS_TILE pairs). The tile is first brought into digit order in shared memory and then
// written out by sorted position: pairs of one digit leave as one contiguous run (16 pairs = 64 bytes on average) instead
// of one 4-byte store per 32-byte sector, which is what bounded the direct scatter at large sizes (33 -> 14 ms at 268 M
// pairs; profiles/r01_config5_microbench.json). Ranking needs no block barrier per round: warp w owns the 512
// consecutive pairs [512 w, 512 w + 512) of the tile, ranks them against its own digit counters (warp match + a running
// per-warp count), and one pass over the 8 x 256 counters turns them into the warps' start positions inside the tile.
#ifndef SPH_RS_MIN_BLOCKS
#define SPH_RS_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(RS_THREADS, SPH_RS_MIN_BLOCKS)
    k_rs_scatter(const u32 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u32 *__restrict__ keys_out,
                 u32 *__restrict__ vals_out, u64 n, int shift, const u32 *__restrict__ hist, u32 ntiles)
{
    __shared__ u32 gbase[256];           // where the tile's pairs of digit d start in the output
    __shared__ u32 lstart[256];          // exclusive tile histogram
    __shared__ u32 wcnt[RS_WARPS][256];  // per-warp digit counts, then the warp's start position for the digit
    __shared__ u32 wtot[RS_WARPS];
    __shared__ u32 skey[RS_TILE], sval[RS_TILE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const u64 tile_base = (u64)blockIdx.x * RS_TILE;
    const u32 in_tile = (u32)min((u64)RS_TILE, n - tile_base);
    {
        // tile histogram from the scanned global one: hist is [digit][tile], exclusive over the whole array
        const u64 at = (u64)threadIdx.x * ntiles + blockIdx.x;
        const u32 here = hist[at];
        const u32 next = at + 1 < (u64)256 * ntiles ? hist[at + 1] : (u32)n;
        gbase[threadIdx.x] = here;
        const u32 c = next - here;
        // exclusive scan of the 256 counts: warp scan + scan of the 8 warp totals
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            u32 v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wtot[wid] = incl;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) wcnt[w][threadIdx.x] = 0;
        __syncthreads();
        u32 off = 0;
        for (int w = 0; w < wid; ++w) off += wtot[w];
        lstart[threadIdx.x] = off + incl - c;
    }
    // phase 1: rank every pair among the pairs of its digit inside the warp's 512 (warp barriers only)
    u32 key[RS_ITEMS], val[RS_ITEMS], lrank[RS_ITEMS];
    const u32 warp_first = (u32)wid * (RS_ITEMS * 32u);
    u32 *mine = wcnt[wid];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const u32 k = warp_first + (u32)r * 32u + (u32)lane;
        const bool valid = k < in_tile;
        key[r] = valid ? keys_in[tile_base + k] : 0xffffffffu;
        val[r] = valid ? vals_in[tile_base + k] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
    {
        const bool valid = warp_first + (u32)r * 32u + (u32)lane < in_tile;
        const u32 d = (key[r] >> shift) & 255u;
        const u32 peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | (u32)lane));
        const u32 rank = __popc(peers & ((1u << lane) - 1u));
        const u32 prev = valid ? mine[d] : 0u;
        __syncwarp();
        if (valid && rank == 0) mine[d] = prev + __popc(peers);
        __syncwarp();
        lrank[r] = prev + rank;
    }
    __syncthreads();
    // phase 2: counts -> start position of (warp, digit) inside the tile
    {
        u32 run = lstart[threadIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w)
        {
            const u32 c = wcnt[w][threadIdx.x];
            wcnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: into digit order
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
        if (warp_first + (u32)r * 32u + (u32)lane < in_tile)
        {
            const u32 at = mine[(key[r] >> shift) & 255u] + lrank[r];
            skey[at] = key[r];
            sval[at] = val[r];
        }
    __syncthreads();
    // phase 4: write out by sorted position: consecutive threads -> consecutive output addresses inside a digit run
#pragma unroll 4
    for (u32 k = threadIdx.x; k < in_tile; k += RS_THREADS)
    {
        const u32 kk = skey[k];
        const u32 d = (kk >> shift) & 255u;
        const u32 pos = gbase[d] + (k - lstart[d]);
        keys_out[pos] = kk;
        vals_out[pos] = sval[k];
    }
}

extern "C" int sphb200_sort_pairs_u32(sphb200_context_t *ctx, uint32_t *keys, uint32_t *values, uint64_t n, int key_bits,
                                      void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ((keys && values) || n == 0), "null pointer");
    SPH_CHECK_ARG(ctx, key_bits > 0 && key_bits <= 32, "key_bits out of range");
    SPH_CHECK_ARG(ctx, n < (1ull << 32), "n too large");
    if (n <= 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    u32 ntiles = (u32)((n + RS_TILE - 1) / RS_TILE);
    void *p;
    int rc = sph_scratch(ctx, 1, (size_t)n * 8 + (size_t)256 * ntiles * 4 + 64, &p);
    if (rc) return rc;
    u32 *k2 = (u32 *)p, *v2 = k2 + n, *hist = v2 + n;
    u32 *kin = keys, *vin = values, *kout = k2, *vout = v2;
    int passes = (key_bits + 7) / 8;
    for (int pass = 0; pass < passes; ++pass)
    {
        int shift = pass * 8;
        SPH_LAUNCH(ctx, k_rs_hist, ntiles, RS_THREADS, 0, st, kin, n, shift, hist, ntiles);
        rc = sph_scan_u32(ctx, hist, hist, (u64)256 * ntiles, 0, st);
        if (rc) return rc;
        SPH_LAUNCH(ctx, k_rs_scatter, ntiles, RS_THREADS, 0, st, kin, vin, kout, vout, n, shift, hist, ntiles);
        u32 *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    if (kin != keys)
    {
        SPH_CUDA(ctx, cudaMemcpyAsync(keys, kin, n * 4, cudaMemcpyDeviceToDevice, st));
        SPH_CUDA(ctx, cudaMemcpyAsync(values, vin, n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// =====================================================================================================
// fused multi-array gather: dst_k[i] = src_k[perm[i]] for up to 16 arrays in one pass over perm
// =====================================================================================================
struct GatherArgs
{
    void *dst[16];
    const void *src[16];
    u32 bytes[16];
    int count;
};

__global__ void __launch_bounds__(256) k_gather_multi(GatherArgs a, const u32 *__restrict__ perm, u32 n, const u32 *__restrict__ n_dev)
{
    if (n_dev) n = min(n, *n_dev); // count known on the device only (slab_decomposition.h)
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 p = perm[i];
#pragma unroll 1
    for (int k = 0; k < a.count; ++k)
    {
        u32 b = a.bytes[k];
        if (b == 4)
            ((u32 *)a.dst[k])[i] = ((const u32 *)a.src[k])[p];
        else if (b == 16)
            ((float4 *)a.dst[k])[i] = ((const float4 *)a.src[k])[p];
        else
        {
            const u32 *s = (const u32 *)a.src[k] + (u64)p * (b / 4);
            u32 *d = (u32 *)a.dst[k] + (u64)i * (b / 4);
            for (u32 e = 0; e < b / 4; ++e) d[e] = s[e];
        }
    }
}

int sph_gather_multi_n(sphb200_context_t *ctx, int count, void *const *dst, const void *const *src, const uint32_t *elem_bytes,
                       const uint32_t *perm, uint32_t n, const uint32_t *n_dev, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && dst && src && elem_bytes && (perm || n == 0), "null pointer");
    SPH_CHECK_ARG(ctx, count >= 0, "negative count");
    if (n == 0) return 0;
    for (int start = 0; start < count; start += 16)
    {
        GatherArgs a;
        a.count = min(16, count - start);
        for (int k = 0; k < a.count; ++k)
        {
            a.dst[k] = dst[start + k];
            a.src[k] = src[start + k];
            a.bytes[k] = elem_bytes[start + k];
            SPH_CHECK_ARG(ctx, a.dst[k] && a.src[k] && a.dst[k] != a.src[k], "gather needs distinct non-null dst/src");
            SPH_CHECK_ARG(ctx, a.bytes[k] >= 4 && a.bytes[k] % 4 == 0, "elem_bytes must be a multiple of 4");
        }
        SPH_LAUNCH(ctx, k_gather_multi, sph_blocks(n, 256), 256, 0, stream, a, perm, n, n_dev);
    }
    return 0;
}
extern "C" int sphb200_gather_multi(sphb200_context_t *ctx, int count, void *const *dst, const void *const *src,
                                    const uint32_t *elem_bytes, const uint32_t *perm, uint32_t n, void *stream)
{
    return sph_gather_multi_n(ctx, count, dst, src, elem_bytes, perm, n, nullptr, stream);
}
