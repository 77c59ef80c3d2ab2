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
    if (ctx->mailbox) sphb200_comm_mailbox_close(ctx);
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


// =====================================================================================================
// Peer mailboxes: migration and boundary planes written straight into the neighbour's memory (see sphb200.h)
// =====================================================================================================
namespace
{
constexpr int MAIL_MAX_VARS = 32;
constexpr unsigned MAIL_HEADER = 64; // bytes: {u64 seq, u32 count}
struct MailVars
{
    void *ptr[MAIL_MAX_VARS];  // push: source arrays; pull: destination arrays
    u32 bytes[MAIL_MAX_VARS];  // element size
    u64 offset[MAIL_MAX_VARS]; // byte offset of the variable's records inside the payload
    int count;
};
struct MailIdentity
{
    cudaIpcMemHandle_t handle; // 64 bytes
    u64 box_bytes;
};

__device__ __forceinline__ void copy_record(char *dst, const char *src, u32 bytes)
{
    if (bytes == 16) *(float4 *)dst = *(const float4 *)src;
    else if (bytes == 4) *(u32 *)dst = *(const u32 *)src;
    else
        for (u32 w = 0; w < bytes; w += 4) *(u32 *)(dst + w) = *(const u32 *)(src + w);
}

// gather + remote store + release. Grid-stride over the list (its length lives in device memory).
__global__ void __launch_bounds__(256)
    k_mail_push(MailVars a, const u32 *__restrict__ idx, const u32 *__restrict__ n_dev, u32 capacity, char *box, u64 seq,
                unsigned *ticket, unsigned *status)
{
    u32 n = *n_dev;
    const bool overflow = n > capacity;
    if (overflow) n = capacity;
    char *payload = box + MAIL_HEADER;
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
    {
        const u32 p = idx[e];
#pragma unroll 1
        for (int k = 0; k < a.count; ++k)
        {
            const u32 b = a.bytes[k];
            copy_record(payload + a.offset[k] + (u64)e * b, (const char *)a.ptr[k] + (u64)p * b, b);
        }
    }
    __threadfence_system(); // this thread's remote stores are visible system-wide before its block takes a ticket
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1)
        {
            *ticket = 0; // ready for the next launch (stream ordered)
            if (overflow) atomicOr(status, 1u);
            __threadfence_system();
            *(volatile u32 *)(box + 8) = overflow ? 0xffffffffu : n;
            __threadfence_system();
            *(volatile u64 *)box = seq; // release: the receiver polls this word
        }
    }
}

__device__ __forceinline__ u64 global_timer_ns()
{
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// one thread polls the header of the local box until the neighbour has released `seq`
__global__ void k_mail_wait(const char *box, u64 seq, u32 *count_dev, unsigned *status, u64 timeout_ns)
{
    const u64 t0 = global_timer_ns();
    const volatile u64 *flag = (const volatile u64 *)box;
    u32 n = 0;
    bool ok = false;
    for (;;)
    {
        if (*flag == seq)
        {
            ok = true;
            break;
        }
        if (global_timer_ns() - t0 > timeout_ns) break;
        __nanosleep(200);
    }
    __threadfence_system(); // acquire side: the payload reads of the following launch come after the flag read
    if (ok)
    {
        n = *(const volatile u32 *)(box + 8);
        if (n == 0xffffffffu)
        {
            atomicOr(status, 1u);
            n = 0;
        }
    }
    else
        atomicOr(status, 2u);
    *count_dev = n;
}

__global__ void __launch_bounds__(256)
    k_mail_unpack(MailVars a, const char *box, const u32 *__restrict__ count_dev, u32 dst_begin, const u32 *__restrict__ extra_dev,
                  u32 dst_end, unsigned *status)
{
    const u32 n = *count_dev;
    const u64 base = (u64)dst_begin + (extra_dev ? *extra_dev : 0u);
    const char *payload = box + MAIL_HEADER;
    if (base + n > dst_end)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(status, 4u);
        return; // nothing is written: the host raises the storage-exhausted error
    }
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
    {
#pragma unroll 1
        for (int k = 0; k < a.count; ++k)
        {
            const u32 b = a.bytes[k];
            copy_record((char *)a.ptr[k] + (base + e) * b, payload + a.offset[k] + (u64)e * b, b);
        }
    }
}

int mail_vars(sphb200_context *ctx, int count, void *const *ptr, const uint32_t *elem_bytes, size_t box_bytes, MailVars *out, u32 *capacity)
{
    SPH_CHECK_ARG(ctx, count > 0 && count <= MAIL_MAX_VARS, "1..32 variables per mailbox transfer");
    u64 per_entry = 0;
    for (int k = 0; k < count; ++k)
    {
        SPH_CHECK_ARG(ctx, ptr[k] && elem_bytes[k] >= 4 && elem_bytes[k] % 4 == 0, "null array or element size not a multiple of 4");
        per_entry += elem_bytes[k];
    }
    SPH_CHECK_ARG(ctx, box_bytes > MAIL_HEADER + per_entry, "mailbox smaller than one entry");
    // 16-byte aligned sections: capacity is a multiple of 4 entries
    const u64 cap = ((box_bytes - MAIL_HEADER) / per_entry) & ~3ull;
    SPH_CHECK_ARG(ctx, cap > 0, "mailbox smaller than four entries");
    u64 off = 0;
    out->count = count;
    for (int k = 0; k < count; ++k)
    {
        out->ptr[k] = ptr[k];
        out->bytes[k] = elem_bytes[k];
        out->offset[k] = off;
        off += cap * elem_bytes[k];
    }
    *capacity = (u32)(cap > 0xfffffff0ull ? 0xfffffff0ull : cap);
    return 0;
}
inline int neighbour_rank(const sphb200_context *ctx, int side)
{
    int r = ctx->rank + (side ? 1 : -1);
    if (ctx->ring) r = (r + ctx->nranks) % ctx->nranks;
    return (r < 0 || r >= ctx->nranks) ? -1 : r;
}
} // namespace

extern "C" int sphb200_comm_mailbox_open(sphb200_context_t *ctx, size_t box_bytes)
{
    SPH_CHECK_ARG(ctx, ctx && (ctx->comm || ctx->self_comm), "no communicator (call sphb200_comm_create)");
    SPH_CHECK_ARG(ctx, !ctx->mailbox, "mailboxes are open already");
    SPH_CHECK_ARG(ctx, box_bytes >= 4096, "box_bytes too small");
    box_bytes = (box_bytes + 255) & ~(size_t)255;
    SPH_CUDA(ctx, cudaSetDevice(ctx->device));
    SPH_CUDA(ctx, cudaMalloc(&ctx->mailbox, 4 * box_bytes));
    SPH_CUDA(ctx, cudaMemset(ctx->mailbox, 0, 4 * box_bytes)); // seq 0 is never used by a push
    SPH_CUDA(ctx, cudaMalloc((void **)&ctx->mailbox_dev, 64));
    SPH_CUDA(ctx, cudaMemset(ctx->mailbox_dev, 0, 64));
    ctx->mailbox_box_bytes = box_bytes;
    ctx->peer_mailbox[0] = ctx->peer_mailbox[1] = nullptr;
    ctx->peer_mapped[0] = ctx->peer_mapped[1] = 0;
    if (ctx->self_comm)
    {
        if (ctx->ring) // a ring of one slab: the rank is its own neighbour on both sides
            for (int s = 0; s < 2; ++s) ctx->peer_mailbox[s] = ctx->mailbox, ctx->peer_box_bytes[s] = box_bytes;
        return 0;
    }
    // identities of all ranks: IPC handle + box size, gathered through the communicator
    const int n = ctx->nranks;
    MailIdentity mine;
    memset(&mine, 0, sizeof(mine));
    SPH_CUDA(ctx, cudaIpcGetMemHandle(&mine.handle, ctx->mailbox));
    mine.box_bytes = box_bytes;
    char *d_all = nullptr;
    SPH_CUDA(ctx, cudaMalloc((void **)&d_all, (size_t)(n + 1) * sizeof(MailIdentity)));
    SPH_CUDA(ctx, cudaMemcpy(d_all + (size_t)n * sizeof(MailIdentity), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    SPH_NCCL(ctx, g_nccl.AllGather(d_all + (size_t)n * sizeof(MailIdentity), d_all, sizeof(MailIdentity), NCCL_UINT8, (nccl_comm_t)ctx->comm, 0));
    SPH_CUDA(ctx, cudaStreamSynchronize(0));
    MailIdentity *all = new MailIdentity[n];
    cudaError_t e = cudaMemcpy(all, d_all, (size_t)n * sizeof(MailIdentity), cudaMemcpyDeviceToHost);
    cudaFree(d_all);
    if (e != cudaSuccess)
    {
        delete[] all;
        SPH_CUDA(ctx, e);
    }
    for (int s = 0; s < 2 && e == cudaSuccess; ++s)
    {
        const int r = neighbour_rank(ctx, s);
        if (r < 0) continue;
        if (r == ctx->rank) ctx->peer_mailbox[s] = ctx->mailbox;
        else if (s == 1 && r == neighbour_rank(ctx, 0)) ctx->peer_mailbox[1] = ctx->peer_mailbox[0]; // ring of two: one peer
        else
        {
            e = cudaIpcOpenMemHandle(&ctx->peer_mailbox[s], all[r].handle, cudaIpcMemLazyEnablePeerAccess);
            ctx->peer_mapped[s] = e == cudaSuccess;
        }
        ctx->peer_box_bytes[s] = all[r].box_bytes;
    }
    delete[] all;
    if (e != cudaSuccess)
    {
        snprintf(ctx->err, sizeof(ctx->err), "sphb200_comm_mailbox_open: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" int sphb200_comm_mailbox_close(sphb200_context_t *ctx)
{
    SPH_CHECK_ARG(ctx, ctx, "null context");
    if (!ctx->mailbox) return 0;
    cudaDeviceSynchronize();
    for (int s = 0; s < 2; ++s)
    {
        if (ctx->peer_mapped[s] && ctx->peer_mailbox[s]) cudaIpcCloseMemHandle(ctx->peer_mailbox[s]);
        ctx->peer_mailbox[s] = nullptr;
        ctx->peer_mapped[s] = 0;
    }
    cudaFree(ctx->mailbox);
    cudaFree(ctx->mailbox_dev);
    ctx->mailbox = nullptr;
    ctx->mailbox_dev = nullptr;
    return 0;
}

extern "C" size_t sphb200_comm_mailbox_peer_bytes(const sphb200_context_t *ctx, int side)
{
    return ctx && ctx->mailbox && ctx->peer_mailbox[side ? 1 : 0] ? ctx->peer_box_bytes[side ? 1 : 0] : 0;
}
extern "C" const uint32_t *sphb200_comm_mailbox_status(const sphb200_context_t *ctx) { return ctx && ctx->mailbox_dev ? ctx->mailbox_dev + 2 : nullptr; }

extern "C" int sphb200_comm_push(sphb200_context_t *ctx, int side, int count, const void *const *src, const uint32_t *elem_bytes,
                                 const uint32_t *idx, const uint32_t *n_dev, uint64_t seq, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ctx->mailbox && src && elem_bytes && idx && n_dev && seq > 0, "bad arguments (mailboxes open?)");
    side = side ? 1 : 0;
    SPH_CHECK_ARG(ctx, ctx->peer_mailbox[side], "no neighbour on that side");
    MailVars a;
    u32 capacity = 0;
    int rc = mail_vars(ctx, count, (void *const *)src, elem_bytes, ctx->peer_box_bytes[side], &a, &capacity);
    if (rc) return rc;
    // what I send to my LEFT neighbour arrives "from the right" there (box 1), and the other way round
    char *box = (char *)ctx->peer_mailbox[side] + ((size_t)(seq & 1ull) * 2 + (side ? 0 : 1)) * ctx->peer_box_bytes[side];
    SPH_LAUNCH(ctx, k_mail_push, 148 * 4, 256, 0, stream, a, idx, n_dev, capacity, box, (u64)seq, ctx->mailbox_dev + side, ctx->mailbox_dev + 2);
    return 0;
}

extern "C" int sphb200_comm_pull(sphb200_context_t *ctx, int side, int count, void *const *dst, const uint32_t *elem_bytes,
                                 uint32_t dst_begin, const uint32_t *dst_extra_dev, uint32_t dst_end, uint32_t *count_dev,
                                 uint64_t seq, void *stream)
{
    SPH_CHECK_ARG(ctx, ctx && ctx->mailbox && dst && elem_bytes && count_dev && seq > 0, "bad arguments (mailboxes open?)");
    side = side ? 1 : 0;
    SPH_CHECK_ARG(ctx, ctx->peer_mailbox[side], "no neighbour on that side");
    MailVars a;
    u32 capacity = 0;
    int rc = mail_vars(ctx, count, dst, elem_bytes, ctx->mailbox_box_bytes, &a, &capacity);
    if (rc) return rc;
    const char *box = (const char *)ctx->mailbox + ((size_t)(seq & 1ull) * 2 + side) * ctx->mailbox_box_bytes;
    SPH_LAUNCH(ctx, k_mail_wait, 1, 1, 0, stream, box, (u64)seq, count_dev, ctx->mailbox_dev + 2, (u64)20000000000ull);
    SPH_LAUNCH(ctx, k_mail_unpack, 148 * 4, 256, 0, stream, a, box, (const u32 *)count_dev, dst_begin, dst_extra_dev, dst_end, ctx->mailbox_dev + 2);
    return 0;
}
