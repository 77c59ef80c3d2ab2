// comm.cu — neighbour exchange and small collectives for slab-decomposed runs (one process per GPU).
// New functionality (the reference has no distributed path): contiguous-range halo exchange over NVLink/NVSwitch.
// Because storage follows the cell order and x is the slowest cell axis, the cell planes a neighbour rank needs are
// contiguous slot ranges of every variable array, so an exchange is a grouped ncclSend/ncclRecv straight out of and
// into the variable arrays — no pack/unpack kernels.
// NCCL is loaded lazily with dlopen (libnccl.so.2: the copy already mapped by the host process, e.g. torch's, or the
// system one), so single-GPU users of libsphb200.so do not need NCCL at all.
#include <dlfcn.h>

#include "common.cuh"

namespace
{
typedef struct
{
    char internal[128];
} nccl_unique_id;
typedef void *nccl_comm_t;
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_UINT64 = 5, NCCL_FLOAT32 = 7 }; // nccl.h ncclDataType_t
enum { NCCL_SUM = 0, NCCL_MAX = 2 };                                                           // nccl.h ncclRedOp_t

struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl(char *err, size_t errlen)
{
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names)
    {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib)
    {
        snprintf(err, errlen, "cannot load libnccl.so.2: %s", dlerror());
        return false;
    }
#define LOAD(field, sym)                                                                     \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, sym);                                      \
    if (!g_nccl.field)                                                                       \
    {                                                                                        \
        snprintf(err, errlen, "libnccl: missing symbol %s", sym);                            \
        return false;                                                                        \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.ok = true;
    return true;
}
} // namespace

#define SPH_NCCL(ctx, expr)                                                                              \
    do                                                                                                   \
    {                                                                                                    \
        int r__ = (expr);                                                                                \
        if (r__ != 0)                                                                                    \
        {                                                                                                \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s -> nccl error %d (%s)", __func__, #expr, r__, \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?");                          \
            return SPHB200_E_COMM;                                                                       \
        }                                                                                                \
    } while (0)

extern "C" int sphb200_comm_unique_id(void *id128)
{
    char err[256];
    if (!id128) return SPHB200_E_INVALID;
    if (!load_nccl(err, sizeof(err))) return SPHB200_E_COMM;
    nccl_unique_id id;
    if (g_nccl.GetUniqueId(&id) != 0) return SPHB200_E_COMM;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int sphb200_comm_create(sphb200_context_t *ctx, int nranks, int rank, const void *id128)
{
    SPH_CHECK_ARG(ctx, ctx && id128 && nranks >= 1 && rank >= 0 && rank < nranks, "bad communicator arguments");
    if (!load_nccl(ctx->err, sizeof(ctx->err))) return SPHB200_E_COMM;
    SPH_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_unique_id id;
    memcpy(&id, id128, sizeof(id));
    nccl_comm_t comm = nullptr;
    SPH_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return 0;
}

extern "C" int sphb200_comm_destroy(sphb200_context_t *ctx)
{
    SPH_CHECK_ARG(ctx, ctx, "null context");
    if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy((nccl_comm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->ring = 0;
    ctx->self_comm = 0;
    ctx->nranks = 1;
    ctx->rank = 0;
    return 0;
}

// Communicator of a single rank, no NCCL: with sphb200_comm_set_ring(ctx, 1) the rank is its own left and right
// neighbour (a periodic box in one slab) and sphb200_comm_exchange becomes two device-to-device copies per segment;
// the reductions and the all-gather are identities.
extern "C" int sphb200_comm_create_self(sphb200_context_t *ctx)
{
    SPH_CHECK_ARG(ctx, ctx, "null context");
    SPH_CHECK_ARG(ctx, !ctx->comm, "the context already has an NCCL communicator");
    ctx->self_comm = 1;
    ctx->rank = 0;
    ctx->nranks = 1;
    return 0;
}

// ring != 0: the neighbours of sphb200_comm_exchange are (rank - 1) mod nranks and (rank + 1) mod nranks
extern "C" int sphb200_comm_set_ring(sphb200_context_t *ctx, int ring)
{
    SPH_CHECK_ARG(ctx, ctx && (ctx->comm || ctx->self_comm), "no communicator (call sphb200_comm_create or sphb200_comm_create_self)");
    ctx->ring = ring ? 1 : 0;
    return 0;
}

extern "C" int sphb200_comm_rank(const sphb200_context_t *ctx) { return ctx ? ctx->rank : 0; }
extern "C" int sphb200_comm_size(const sphb200_context_t *ctx) { return ctx && ctx->comm ? ctx->nranks : 1; }
extern "C" int sphb200_comm_is_ring(const sphb200_context_t *ctx) { return ctx && (ctx->comm || ctx->self_comm) ? ctx->ring : 0; }

// One grouped exchange with the left (rank-1) and right (rank+1) neighbour: `count` segments per direction.
// send_left[k]/send_left_bytes[k] go to rank-1 and arrive in ITS recv_right[k]; symmetric for the other direction.
// Ranks at the ends of the slab chain skip the missing side; on a ring (sphb200_comm_set_ring) the chain has no ends.
// All pointers are device pointers.
extern "C" int sphb200_comm_exchange(sphb200_context_t *ctx, int count, const void *const *send_left, const size_t *send_left_bytes,
                                     void *const *recv_left, const size_t *recv_left_bytes, const void *const *send_right,
                                     const size_t *send_right_bytes, void *const *recv_right, const size_t *recv_right_bytes,
                                     void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (ctx->comm || ctx->self_comm), "no communicator (call sphb200_comm_create)");
    SPH_CHECK_ARG(ctx, count >= 0, "negative count");
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->self_comm)
    {
        // one rank: without the ring there is no neighbour at all; with it, what goes left comes back in from the right
        if (!ctx->ring) return 0;
        for (int k = 0; k < count; ++k)
        {
            const size_t sl = send_left ? send_left_bytes[k] : 0, sr = send_right ? send_right_bytes[k] : 0;
            const size_t rl = recv_left ? recv_left_bytes[k] : 0, rr = recv_right ? recv_right_bytes[k] : 0;
            SPH_CHECK_ARG(ctx, sl == rr && sr == rl, "ring of one rank: send and receive sizes of a segment differ");
            if (sl) SPH_CUDA(ctx, cudaMemcpyAsync(recv_right[k], send_left[k], sl, cudaMemcpyDeviceToDevice, st));
            if (sr) SPH_CUDA(ctx, cudaMemcpyAsync(recv_left[k], send_right[k], sr, cudaMemcpyDeviceToDevice, st));
        }
        return 0;
    }
    nccl_comm_t comm = (nccl_comm_t)ctx->comm;
    int left = ctx->rank - 1, right = ctx->rank + 1;
    if (ctx->ring)
    {
        left = (left + ctx->nranks) % ctx->nranks;
        right = right % ctx->nranks;
    }
    const bool has_left = left >= 0, has_right = right < ctx->nranks;
    // Order inside the group: NCCL matches the sends and receives between one pair of ranks in the order they are
    // posted. In a ring of two ranks both neighbours are the same peer: its first send (to ITS left) is my segment from
    // the right, so receives are posted right before left while sends go left before right.
    SPH_NCCL(ctx, g_nccl.GroupStart());
    for (int k = 0; k < count; ++k)
    {
        if (has_left && send_left && send_left_bytes[k]) SPH_NCCL(ctx, g_nccl.Send(send_left[k], send_left_bytes[k], NCCL_UINT8, left, comm, st));
        if (has_right && send_right && send_right_bytes[k]) SPH_NCCL(ctx, g_nccl.Send(send_right[k], send_right_bytes[k], NCCL_UINT8, right, comm, st));
        if (has_right && recv_right && recv_right_bytes[k]) SPH_NCCL(ctx, g_nccl.Recv(recv_right[k], recv_right_bytes[k], NCCL_UINT8, right, comm, st));
        if (has_left && recv_left && recv_left_bytes[k]) SPH_NCCL(ctx, g_nccl.Recv(recv_left[k], recv_left_bytes[k], NCCL_UINT8, left, comm, st));
    }
    SPH_NCCL(ctx, g_nccl.GroupEnd());
    ctx->launches++; // the grouped exchange runs as one NCCL kernel
    return 0;
}

extern "C" int sphb200_comm_allreduce_max_f32(sphb200_context_t *ctx, float *dev_inout, int n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (ctx->comm || ctx->self_comm) && dev_inout && n > 0, "bad arguments");
    if (ctx->self_comm) return 0;
    SPH_NCCL(ctx, g_nccl.AllReduce(dev_inout, dev_inout, (size_t)n, NCCL_FLOAT32, NCCL_MAX, (nccl_comm_t)ctx->comm, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

extern "C" int sphb200_comm_allreduce_sum_f64(sphb200_context_t *ctx, double *dev_inout, int n, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (ctx->comm || ctx->self_comm) && dev_inout && n > 0, "bad arguments");
    if (ctx->self_comm) return 0;
    SPH_NCCL(ctx, g_nccl.AllReduce(dev_inout, dev_inout, (size_t)n, 8 /* ncclFloat64 */, NCCL_SUM, (nccl_comm_t)ctx->comm, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

extern "C" int sphb200_comm_allgather_u64(sphb200_context_t *ctx, const uint64_t *dev_send, uint64_t *dev_recv, int n_per_rank,
                                          void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && (ctx->comm || ctx->self_comm) && dev_send && dev_recv && n_per_rank > 0, "bad arguments");
    if (ctx->self_comm)
    {
        SPH_CUDA(ctx, cudaMemcpyAsync(dev_recv, dev_send, (size_t)n_per_rank * sizeof(uint64_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return 0;
    }
    SPH_NCCL(ctx, g_nccl.AllGather(dev_send, dev_recv, (size_t)n_per_rank, NCCL_UINT64, (nccl_comm_t)ctx->comm, (cudaStream_t)stream));
    ctx->launches++;
    return 0;
}

// Positions that crossed the periodic seam of a ring of slabs: x += delta (delta = +L for what the last rank receives
// from rank 0, -L the other way round), the arithmetic of the reference's ghost list entry (domain_bounding.cpp:26,45).
// A segment holds the sender's boundary plane (ghosts here) and its leavers (own here), told apart by the cell plane
// the sender saw them in. Ownership is by cell plane, so in the rounding cases where the sum would fall into the wrong
// plane it is held at the edge of the right one: leavers inside the box planes, the boundary plane in the one ghost
// plane beyond the face (no particle lost or owned twice, ghost order = the owner's plane order).
__global__ void __launch_bounds__(256) k_seam_shift(char *__restrict__ base, u32 stride, u32 n, float delta, sphb200_seam_t sm)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float *p = reinterpret_cast<float *>(base + (size_t)k * stride);
    const float x = *p;
    const int plane = cell_coord(x, sm.mesh_lower, sm.mesh_spacing, sm.mesh_cells);
    float y = __fadd_rn(x, delta);
    if (delta > 0.f)
    {
        if (plane < sm.first_plane) y = fminf(y, sm.own_max);                          // leaver of rank 0
        else y = fminf(fmaxf(y, sm.ghost_high_min), sm.ghost_high_max);                // its first plane: my upper ghost plane
    }
    else
    {
        if (plane >= sm.first_plane + sm.box_planes) y = fmaxf(y, sm.own_min);         // leaver of the last rank
        else y = fmaxf(fminf(y, sm.ghost_low_max), sm.ghost_low_min);                  // its last plane: my lower ghost plane
    }
    *p = y;
}

extern "C" int sphb200_seam_shift(sphb200_context_t *ctx, void *base, uint32_t stride_bytes, uint32_t n, float delta,
                                  const sphb200_seam_t *seam, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && seam && (n == 0 || base) && stride_bytes >= 12 && stride_bytes % 4 == 0, "bad arguments");
    if (n == 0) return 0;
    SPH_LAUNCH(ctx, k_seam_shift, sph_blocks(n, 256), 256, 0, (cudaStream_t)stream, (char *)base, stride_bytes, n, delta, *seam);
    return 0;
}
