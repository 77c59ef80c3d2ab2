/* ====================================================================================================
 *  sphb200.h — C ABI of libsphb200.so: the B200-native (sm_100a) weakly-compressible SPH hot path.
 *
 *  This is the drop-in boundary for what SPHinXsys reaches today through `execution::par_device`
 *  (SYCL).  Generic lambdas cannot cross a C ABI, so the seam sits one level up, at the `exec()` of each
 *  dynamics class: every entry point below replaces the device launches one reference `exec()` makes.
 *  Citations are `path:line` relative to /root/reference/src/shared unless a longer path is given.
 *
 *  Conventions
 *    - every function returns int: 0 = OK, < 0 = SPHB200_E_* below, > 0 = a cudaError_t value;
 *      nothing throws or exits; sphb200_last_error_string() describes the last failure of a context.
 *    - all pointers inside *_view_t / *_args_t are DEVICE pointers unless the field says "host";
 *      the library never owns particle data, only its own scratch (sort buffers, histograms).
 *    - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it except where a value is
 *      returned to the host (scan total, reductions, required capacities), exactly where the reference
 *      synchronises too (particle_iterators_sycl.h:80-105, algorithm_primitive_sycl.h:96-122).
 *    - Vecd variables live on the device as float4 (x, y, z, w); `w` is padding unless stated. Host
 *      `Vecd*` arrays (packed 3 floats) are converted with sphb200_vec3_to_vec4 / sphb200_vec4_to_vec3.
 *      2-D cases use the same kernels with z == 0 and one cell layer in z.
 *    - Real = float, UnsignedInt = uint32_t (base_data_type.h:50-60 under SPHINXSYS_USE_SYCL).
 * ==================================================================================================== */
#ifndef SPHB200_H
#define SPHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHB200_VERSION 100

enum
{
    SPHB200_OK = 0,
    SPHB200_E_INVALID = -1,     /* bad argument (null pointer, unsupported enum, n too large) */
    SPHB200_E_CAPACITY = -2,    /* output buffer smaller than required; required size is returned */
    SPHB200_E_UNSUPPORTED = -3, /* type tuple outside the closed set listed in DESIGN.md */
    SPHB200_E_NOMEM = -4,
    SPHB200_E_COMM = -5         /* NCCL missing or a collective failed (see sphb200_last_error_string) */
};

typedef struct sphb200_context sphb200_context_t; /* opaque; one per GPU, one driving host thread */

typedef struct { float x, y, z, w; } sphb200_vec4_t;

/* Mesh POD; ref: meshes/base_mesh.h, base_mesh.cpp:6-16 (computed by the host in Real precision). */
typedef struct
{
    float lower[3];   /* mesh_lower_bound_ */
    float spacing;    /* grid_spacing_ */
    int32_t cells[3]; /* all_cells_ = all_grid_points_ - 1 ; cells[2] == 1 in 2-D */
} sphb200_mesh_t;

/* KernelTabulatedCK + Neighbor<SPHAdaptation,SPHAdaptation>::SmoothingKernel scalars;
 * ref: shared_ck/smoothing_kernel/kernel_tabulated_ck.h:37-72, shared_ck/body_relation/neighbor_method.hpp:18-142 */
typedef struct
{
    int32_t dim;            /* 2 | 3 */
    int32_t kind;           /* 0 Wendland C2, 1 Laguerre-Gauss (informational; the table defines the kernel) */
    float h;                /* max(src_h, tar_h): inv_h_ = 1/h */
    float src_h;            /* source smoothing length: W0 uses src_inv_h_ */
    float kernel_size;      /* 2.0 */
    float dimension_factor; /* Kernel::DimensionFactor{2,3}D(), base_kernel.h:91-93 */
    float w[24];            /* w_1d[k]  = W_1D((k-1) dq), dq = kernel_size/20 */
    float dw[24];           /* dw_1d[k] = dW_1D((k-1) dq) */
} sphb200_kernel_t;

/* WeaklyCompressibleFluid::EosKernel + RiemannSolver::ComputingKernel selection;
 * ref: materials/weakly_compressible_fluid.h:40-70, shared_ck/.../riemann_solver/riemann_solver_ck.h:46-173 */
typedef struct
{
    float rho0, c0;
    int32_t riemann;       /* 0 NoRiemannSolverCK, 1 AcousticRiemannSolverCK, 2 DissipativeRiemannSolverCK */
    int32_t correction;    /* 0 NoKernelCorrectionCK, 1 LinearCorrectionCK (needs fluid.B) */
    float limiter_coeff;   /* 3.0 (riemann_solver_ck.h:97) */
    int32_t free_surface;  /* DensityRegularization flow type: 1 FreeSurface, 0 Internal */
    int32_t formulation;   /* 0: CK (state Compression/CompressionRate, tabulated kernel, current positions);
                              1: legacy Integration1stHalf/2ndHalf + DensitySummation (state Density/DensityChangeRate
                                 passed in rho / compression_rate, analytic kernel, pair geometry FROZEN at the last
                                 configuration update: keep the gather records packed at that time, pass dpos == pos).
                              ref: particle_dynamics/fluid_dynamics/fluid_integration.hpp:49-231, density_summation.cpp:8-78 */
    float sigma0;          /* legacy DensitySummation: lattice number density (adaptation.cpp:26-60) */
    float wall_rho0;       /* legacy DensitySummation: reference density of the contact (wall) material */
} sphb200_fluid_t;

/* Device views of one fluid body (the DiscreteVariables the acoustic steps register,
 * acoustic_step_1st_half.hpp:13-39, density_regularization.hpp:14-27, fluid_time_step_ck.cpp:34-49). */
typedef struct
{
    uint32_t n;                   /* TotalRealParticles */
    sphb200_vec4_t *pos;          /* Position */
    sphb200_vec4_t *vel;          /* Velocity */
    sphb200_vec4_t *dpos;         /* Displacement */
    sphb200_vec4_t *force;        /* Force */
    sphb200_vec4_t *force_prior;  /* ForcePrior */
    float *vol;                   /* VolumetricMeasure */
    float *mass;                  /* Mass */
    float *rho;                   /* Density */
    float *p;                     /* Pressure */
    float *compression;           /* Compression */
    float *compression_rate;      /* CompressionRate */
    float *vol_ref;               /* VolumetricMeasureRef */
    float *compression_sum;       /* CompressionSummation */
    float *B;                     /* LinearCorrectionMatrix, 9 floats row-major per particle, or NULL */
    void *correction_record;      /* 32-byte records (Bxx, Bxy, Bxz, Byy, Byz, Bzz, p, -): the symmetric part of B and the
                                     pressure as ONE gather per neighbour for the 1st-half interaction with LinearCorrectionCK
                                     (which reads B and p of the neighbour); B part also for ViscousForceCK with correction.
                                     sphb200_linear_correction_matrix writes the B part next to B, the 1st-half initialize
                                     keeps the pressure slot current (sphb200_pack_correction_records for values written
                                     elsewhere); required whenever material.correction is set for those dynamics */
    sphb200_vec4_t *posvol;       /* derived gather records, one load per neighbour instead of two; refresh with
                                     sphb200_pack_records whenever their sources changed outside the library:
                                     posvol = (x, y, z, Vol)                         [1st half, correction matrix] */
    sphb200_vec4_t *posvolref;    /* (x, y, z, VolRef)                             [compression summation] */
    void *posvolvel;              /* 32-byte records (x, y, z, Vol, vx, vy, vz, -) [2nd half: one 256-bit load];
                                     the 1st-half update keeps the velocity part current */
    uint32_t active_begin;        /* dynamics update the slots [active_begin, active_end) only; slots outside (ghost */
    uint32_t active_end;          /* particles of a decomposed run) are read as neighbours. active_end == 0: [0, n) */
} sphb200_fluid_view_t;

/* Device views of one wall (contact) body; ref: interaction_ck.hpp:79-91 (Interaction<Wall>). */
typedef struct
{
    uint32_t n;
    const sphb200_vec4_t *pos;      /* Position */
    const sphb200_vec4_t *posvol;   /* (x, y, z, Vol) */
    const sphb200_vec4_t *posvolref;/* (x, y, z, VolRef) */
    const sphb200_vec4_t *vel_ave;  /* AverageVelocity, or NULL == 0 */
    const sphb200_vec4_t *acc_ave;  /* AverageAcceleration, or NULL == 0 */
    const sphb200_vec4_t *normal;   /* NormalDirection */
    const float *vol_ref;           /* VolumetricMeasureRef */
} sphb200_wall_view_t;

/* CellLinkedList<SPHAdaptation> storage; ref: meshes/cell_linked_list.cpp:167-175 */
typedef struct
{
    uint32_t *cell_offset;      /* [total_cells + 1] */
    uint32_t *particle_index;   /* [max(n, 1)] ; inside a cell, ascending particle index (deterministic) */
    sphb200_vec4_t *sorted_pos; /* optional [n]: sorted_pos[k] = pos[particle_index[k]] (contiguous candidate runs for
                                   the neighbour search); NULL = not kept */
} sphb200_cell_list_t;

/* Relation<...>::NeighborList in the library's coalesced layout ("SELL-32").
 * Rows are stored per SLOT t; slot t holds source particle i = order[t] (order == NULL: i = t). The library
 * uses the source body's cell-list order (order = particle_index) so that the 32 rows of a slice belong to
 * spatially adjacent particles whatever the storage order is. Slots are grouped in slices of 32; entry k of
 * slot t is index[slice_offset[t/32] + 32*k + (t%32)], k < count[t]; entries are TARGET PARTICLE ids.
 * Row order = reference search order (cells x->y->z, then in-cell order), so sums are reproducible.
 * sphb200_relation_export_csr() converts to the reference's particle_offset_/neighbor_index_ form indexed by
 * particle id (shared_ck/body_relation/relation_ck.h:89-102). */
typedef struct
{
    uint32_t *count;        /* [n + 1] neighbours of slot t (entry n is scratch) */
    uint32_t *slice_offset; /* [ceil(n/32) + 1] */
    uint32_t *index;        /* [capacity] */
    uint64_t capacity;      /* entries allocated for `index` */
    const uint32_t *order;  /* [n] slot -> source particle id, or NULL (identity) */
    int32_t bank_aligned;   /* 1 (cell-ordered one-pass builds only): after the search the entries of every row are permuted
                               inside groups of 32 so that entry s in row k of slot t has (s - t - k) mod 8 == 0 wherever
                               possible — in every row the 8 lanes of a quarter warp then gather records from 8 different
                               L1 bank groups. Entry SET and count are unchanged; sphb200_relation_export_csr sorts such rows
                               back into ascending target order. 0: rows in search order (cells x -> y -> z, in-cell order).
                               Values 1 + o, o in 0..7: t is taken as t + o — a slab-decomposed run passes the slot its first
                               stored particle has in the undecomposed run (mod 8) for CONTACT relations, whose target slots
                               do not move with the slab, so that rows come out in the single-GPU order. */
} sphb200_relation_t;

/* One neighbour search: which particles look (src, in slot order), where they look (tar body + its cell list) */
typedef struct
{
    sphb200_mesh_t tar_mesh;
    sphb200_kernel_t kernel;
    const sphb200_vec4_t *src_pos;        /* [n_src (+ghosts)] source Position */
    uint32_t n_src;                       /* number of source slots */
    const uint32_t *src_order;            /* slot -> source particle id (NULL: identity) */
    const sphb200_vec4_t *src_sorted_pos; /* optional: source positions in slot order */
    const sphb200_vec4_t *tar_pos;        /* target Position */
    sphb200_cell_list_t tar_list;         /* target cell-linked list */
    int32_t is_inner;                     /* 1: Inner<> (exclude j == i), 0: Contact<> */
    int32_t legacy_criterion;             /* 0: |inv_h d|^2 < kernel_size^2 (neighbor_method.hpp:152-156);
                                             1: |d|^2 < (kernel_size h)^2, NeighborBuilder (kernels/base_kernel.h:105-114) */
    int32_t search_depth;                 /* cells each side: 1 inner; contact: cell_linked_list.hpp:161-167 */
    uint32_t src_begin, src_end;          /* source slots searched: [src_begin, src_end); src_end == 0: [0, n_src) */
    int32_t cell_ordered;                 /* 1: src_pos and tar_pos are STORED in the cell order of their own lists
                                             (slot == particle id, particle_index == identity, src_order == NULL; see
                                             sphb200_cell_list_build_reorder): selects the warp-uniform search */
    /* optional SECOND candidate set on the same mesh, searched after the first and appended to the same rows
     * (cell_ordered searches only). Used for the periodic images of the target body, which are stored cell ordered
     * behind its real particles: hits are recorded as tar2_index_base + (slot in tar2_pos). tar2_pos == NULL: none. */
    const sphb200_vec4_t *tar2_pos;
    sphb200_cell_list_t tar2_list;
    uint32_t tar2_index_base;
} sphb200_search_t;

/* Periodic box of a body; ref: particle_dynamics/general_dynamics/domian_bouding/domain_bounding.h:48-66
 * (PeriodicAlongAxis: bounding_bounds_, axis_, periodic_translation_ = upper - lower), one bit per periodic axis. */
typedef struct
{
    float lower[3], upper[3]; /* bounding_bounds_ */
    int32_t axes;             /* bit d: periodic along axis d */
    float cutoff;             /* cut_off_radius_max_: images are made for particles closer than this to a periodic face */
} sphb200_periodic_t;

/* ---------------------------------------------------------------------------------------------------
 * context, diagnostics, memory  (replaces implementation_sycl.h:43-160 ExecutionInstance + USM helpers)
 * ------------------------------------------------------------------------------------------------- */
int sphb200_version(void);
int sphb200_context_create(int device, sphb200_context_t **out);
int sphb200_context_destroy(sphb200_context_t *ctx);
const char *sphb200_last_error_string(const sphb200_context_t *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t sphb200_launch_count(const sphb200_context_t *ctx);

int sphb200_malloc_device(void **ptr, size_t bytes);                 /* allocateDeviceOnly, implementation_sycl.h:99-105 */
int sphb200_malloc_host(void **ptr, size_t bytes);                   /* allocateHostStaging (pinned) */
int sphb200_free_device(void *ptr);
/* device allocations made so far through the library (arrays + scratch arenas): a steady-state loop must not add any */
uint64_t sphb200_device_allocation_count(void);
int sphb200_free_host(void *ptr);
int sphb200_copy_h2d(void *dst, const void *src, size_t bytes, void *stream); /* copyToDevice   :117-127 */
int sphb200_copy_d2h(void *dst, const void *src, size_t bytes, void *stream); /* copyFromDevice :129-139 */
int sphb200_copy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int sphb200_stream_sync(void *stream);
/* transfer overlap (extension; the SYCL build has a single in-order queue): a non-blocking side stream and events, so
 * that sphb200_copy_h2d / _d2h on pinned host memory run while dynamics execute on the main stream */
int sphb200_stream_create(void **stream);
/* side stream whose kernels are scheduled ahead of (high_priority != 0) or behind the other streams' pending blocks:
 * the halo exchange and the boundary-plane launches of a decomposed run (slab_decomposition.h) */
int sphb200_stream_create_with_priority(void **stream, int high_priority);
int sphb200_stream_destroy(void *stream);
int sphb200_event_create(void **event);
int sphb200_event_destroy(void *event);
int sphb200_event_record(void *event, void *stream);
int sphb200_stream_wait_event(void *stream, void *event);
int sphb200_fill_u32(sphb200_context_t *ctx, uint32_t *dst, uint32_t value, uint64_t n, void *stream);
int sphb200_fill_f32(sphb200_context_t *ctx, float *dst, float value, uint64_t n, void *stream);
int sphb200_iota_u32(sphb200_context_t *ctx, uint32_t *dst, uint64_t n, void *stream); /* dst[i] = i */
/* Vecd layout conversion (device to device): packed 3-float AoS <-> float4 */
int sphb200_vec3_to_vec4(sphb200_context_t *ctx, sphb200_vec4_t *dst, const float *src3, uint32_t n, void *stream);
int sphb200_vec4_to_vec3(sphb200_context_t *ctx, float *dst3, const sphb200_vec4_t *src, uint32_t n, void *stream);
int sphb200_pack_posvol(sphb200_context_t *ctx, sphb200_vec4_t *posvol, const sphb200_vec4_t *pos, const float *vol,
                        uint32_t n, void *stream);
/* all gather records of a body in one pass; any output (and the inputs only it needs) may be NULL */
int sphb200_pack_records(sphb200_context_t *ctx, uint32_t n, const sphb200_vec4_t *pos, const float *vol, const float *vol_ref,
                         const sphb200_vec4_t *vel, sphb200_vec4_t *posvol, sphb200_vec4_t *posvolref, void *posvolvel,
                         void *stream);

/* ---------------------------------------------------------------------------------------------------
 * primitives  (replaces algorithm_primitive_sycl.h:44-137)
 * ------------------------------------------------------------------------------------------------- */
/* exclusive_scan(policy, first, d_first, n, plus): out[0]=0, out[i]=sum_{k<i} in[k]; the last input is not
 * read into the result; returns out[n-1] through *last_host when non-NULL (synchronises the stream).
 * ref: common/algorithm_primitive.h:244-279, src_sycl/.../algorithm_primitive_sycl.h:96-122 */
int sphb200_exclusive_scan_u32(sphb200_context_t *ctx, const uint32_t *in, uint32_t *out, uint64_t n,
                               uint32_t *last_host, void *stream);
/* RadixSort<...>::sort_by_key: stable ascending LSD radix sort of (key, value) pairs, in place.
 * ref: src_sycl/.../algorithm_primitive_sycl.hpp:90-238 */
int sphb200_sort_pairs_u32(sphb200_context_t *ctx, uint32_t *keys, uint32_t *values, uint64_t n, int key_bits,
                           void *stream);
/* UpdateSortableVariables: dst[i] = src[perm[i]] for `count` arrays in one launch; elem_bytes[k] in {4, 16, 36}.
 * ref: shared_ck/.../base_configuration_dynamics.h:74-128 (the reference copies then gathers, per variable) */
int sphb200_gather_multi(sphb200_context_t *ctx, int count, void *const *dst, const void *const *src,
                         const uint32_t *elem_bytes, const uint32_t *perm, uint32_t n, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * configuration dynamics
 * ------------------------------------------------------------------------------------------------- */
/* ParticleSortCK::prepareSequence: key = Morton(cell(pos)), perm[i] = i; ref: particle_sort_ck.hpp:61-67,
 * meshes/base_mesh.hxx:9-15,85-99.  `cell_id` (linear cell index, base_mesh.hxx:73-78) optional. */
int sphb200_morton_keys(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos, uint32_t n,
                        uint32_t *keys, uint32_t *perm, uint32_t *cell_id, void *stream);
/* ParticleSortCK::updateSortedID: sorted_id[original_id[i]] = i; ref: particle_sort_ck.hpp:69-74 */
int sphb200_update_sorted_id(sphb200_context_t *ctx, const uint32_t *original_id, uint32_t *sorted_id, uint32_t n,
                             void *stream);
/* UpdateCellLinkedList::exec: count -> scan -> fill (+ in-cell ordering); ref: update_cell_linked_list.hpp:40-106 */
int sphb200_cell_list_build(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos, uint32_t n,
                            sphb200_cell_list_t list, void *stream);
/* UpdateCellLinkedList::exec for a body whose storage FOLLOWS the cell order: builds the list from `pos`, then
 * gathers `count` variables into their shadow buffers by the new permutation (dst_k[slot] = src_k[old index],
 * elem_bytes as in sphb200_gather_multi; `pos` must be among them) and leaves particle_index == identity.
 * Inside a cell particles are ordered by ascending sort_key[i] (unique keys, e.g. the reference particle id;
 * NULL = the old index), so lists and summation order do not depend on the storage history. The caller swaps
 * each variable with its shadow afterwards. */
int sphb200_cell_list_build_reorder(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos,
                                    uint32_t n, const uint32_t *sort_key, sphb200_cell_list_t list, int count,
                                    void *const *dst, const void *const *src, const uint32_t *elem_bytes, void *stream);
/* The same with the particle count taken from DEVICE memory: n is the capacity the launches are sized for and
 * min(n, *n_dev) particles are processed (n_dev == NULL: n). For slab-decomposed runs, whose arrivals are counted on the
 * device (sphb200_comm_pull), so that a rebuild needs no host round trip before its last step. */
int sphb200_cell_list_build_reorder_n(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos,
                                      uint32_t n, const uint32_t *n_dev, const uint32_t *sort_key, sphb200_cell_list_t list,
                                      int count, void *const *dst, const void *const *src, const uint32_t *elem_bytes,
                                      void *stream);
/* Slab bookkeeping in one device record (new): out_dev[0..k) = cell_offset[cells_dev[0..k)], out_dev[k] = *n_dev,
 * out_dev[k + 1] = the mailbox status word; own64_dev[0] = out[hi] - out[lo] with own_lo_hi = lo | hi << 8 (the rank's own
 * particle count, as the 64-bit word the all-gather of the slot origins sends). */
/* *out_dev = base + *a_dev + *b_dev (NULL terms are 0): the stored-particle count after the arrivals of a rebuild. */
int sphb200_slab_total(sphb200_context_t *ctx, uint32_t base, const uint32_t *a_dev, const uint32_t *b_dev, uint32_t *out_dev, void *stream);
int sphb200_slab_bounds(sphb200_context_t *ctx, const uint32_t *cell_offset, const uint32_t *cells_dev, int k,
                        const uint32_t *n_dev, int own_lo_hi, uint32_t *out_dev, uint64_t *own64_dev, void *stream);
/* UpdateRelation<Inner<>>::exec and <Contact<>>::exec, split as the reference splits them so the host can
 * grow `index` between the two phases (update_body_relation.hpp:117-164, 240-288):
 *   *_count : fills rel.count and rel.slice_offset, returns the required capacity (entries) in *required_host
 *   *_fill  : fills rel.index (returns SPHB200_E_CAPACITY if rel.capacity is too small)
 * Neighbour criterion: |inv_h (x_i - x_j)|^2 < kernel_size^2, strict, ops rounded separately
 * (neighbor_method.hpp:152-156); inner additionally j != i.  `tar_*` describe the searched body. */
int sphb200_relation_count(sphb200_context_t *ctx, const sphb200_search_t *search, sphb200_relation_t rel,
                           uint64_t *required_host, void *stream);
int sphb200_relation_fill(sphb200_context_t *ctx, const sphb200_search_t *search, sphb200_relation_t rel, void *stream);
/* One-pass variant: every slice gets the fixed length 32*stride (slice_offset[s] = 32*stride*s, which needs
 * rel.capacity >= 32*stride*ceil(n/32)), so count and fill happen in a single search. Rows longer than `stride`
 * are truncated in `index` but counted in full: *max_count_host > stride tells the host to fall back to
 * sphb200_relation_count/_fill (nothing is silently dropped). */
int sphb200_relation_build_fixed(sphb200_context_t *ctx, const sphb200_search_t *search, sphb200_relation_t rel,
                                 uint32_t stride, uint32_t *max_count_host, void *stream);
/* PeriodicBounding::checkLowerBound/checkUpperBound for every periodic axis: x < lower -> x += L; x > upper -> x -= L.
 * ref: domain_bounding.h:98-108 */
int sphb200_periodic_bounding(sphb200_context_t *ctx, const sphb200_periodic_t *box, sphb200_vec4_t *pos, uint32_t n,
                              void *stream);
/* PeriodicCellLinkedList::exec for all periodic axes at once: every particle with lower < x < lower + cutoff gets an
 * image at x + L, every particle with upper - cutoff < x < upper one at x - L, and the combinations over the axes
 * (what the reference's axis-by-axis insertion of ghost list entries produces, domain_bounding.cpp:18-65).
 * Images are written in ascending source order: image_pos[k] (translated position), image_src[k] (source particle).
 * The number of images is returned in *count_host (synchronises); if it exceeds `capacity` nothing is written and
 * SPHB200_E_CAPACITY is returned so the caller can grow the buffers. */
int sphb200_periodic_images(sphb200_context_t *ctx, const sphb200_periodic_t *box, const sphb200_vec4_t *pos, uint32_t n,
                            sphb200_vec4_t *image_pos, uint32_t *image_src, uint32_t capacity, uint32_t *count_host,
                            void *stream);
/* Refresh of image (ghost) particles stored behind the real ones: for g < n_ghost copies `copy_bytes` bytes at
 * `offset_bytes` inside element ghost_src[g] of `array` to the same place inside element n_real + g.
 * elem_bytes, offset_bytes and copy_bytes are multiples of 4. */
int sphb200_ghost_copy(sphb200_context_t *ctx, void *array, uint32_t elem_bytes, uint32_t offset_bytes, uint32_t copy_bytes,
                       const uint32_t *ghost_src, uint32_t n_real, uint32_t n_ghost, void *stream);
/* SELL-32 (slot order) -> reference CSR indexed by particle id (particle_offset_[n+1], neighbor_index_[total]).
 * src_ids: slot -> id of the source particle (NULL: rel.order, else identity); tar_ids: stored target index -> id
 * (NULL: identity). Row order is preserved (rel.bank_aligned: rows are emitted in ascending stored target index, which
 * is the reference search order for cell-ordered bodies). */
int sphb200_relation_export_csr(sphb200_context_t *ctx, sphb200_relation_t rel, uint32_t n, const uint32_t *src_ids,
                                const uint32_t *tar_ids, uint32_t *particle_offset, uint32_t *neighbor_index,
                                uint64_t index_capacity, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * fluid dynamics
 * ------------------------------------------------------------------------------------------------- */
typedef struct
{
    sphb200_fluid_view_t fluid;
    sphb200_wall_view_t wall;       /* wall.n == 0: no contact body */
    sphb200_relation_t inner;       /* inner.order defines the slot order the launch iterates in */
    sphb200_relation_t contact;     /* must have been built with the same src_order as `inner` */
    sphb200_kernel_t kernel;
    sphb200_fluid_t material;
} sphb200_fluid_args_t;

/* GravityForceCK<Gravity>::update (+ForcePriorCK::update); ref: general_dynamics/force_prior_ck.hpp:38-44 */
int sphb200_gravity_force(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid, const float gravity[3],
                          sphb200_vec4_t *previous_force, void *stream);
/* InteractionDynamicsCK<CompressionSummation<Inner<>,Contact<>>>::exec, optionally followed in the same launch
 * by DensityRegularization::update (regularize != 0).  ref: fluid_dynamics/density_regularization.hpp:40-118 */
int sphb200_compression_summation(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, int regularize, void *stream);
int sphb200_density_regularization(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, void *stream);
/* StateDynamics<AdvectionStepSetup>, <UpdateParticlePosition>; ref: fluid_time_step_ck.h:139-180 */
int sphb200_advection_setup(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid, void *stream);
int sphb200_update_position(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid, void *stream);
/* ReduceDynamicsCK<AdvectionTimeStepCK>, <AcousticTimeStepCK>: return the REDUCED value (max) and the
 * finished dt, as FinishDynamics::Result does.  ref: fluid_time_step_ck.hpp:31-57, fluid_time_step_ck.cpp:12-27 */
int sphb200_advection_time_step(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid, float h_min, float u_ref,
                                float cfl, float *reduced_host, float *dt_host, void *stream);
int sphb200_acoustic_time_step(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float h_min, float cfl,
                               float *reduced_host, float *dt_host, void *stream);
/* legacy ReduceDynamics<AdvectionViscousTimeStep>: max(|v|^2, 4 h |F + F_prior| / m); the legacy AcousticTimeStep
 * (max(c0 + |v|)) is sphb200_acoustic_time_step with material.formulation == 1.
 * ref: particle_dynamics/fluid_dynamics/fluid_time_step.cpp:21-59 */
int sphb200_advection_time_step_legacy(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid, float h_min, float u_ref,
                                       float cfl, float *reduced_host, float *dt_host, void *stream);
/* InteractionDynamicsCK<AcousticStep1stHalf<Inner<OneLevel,..>,Contact<Wall,..>>>::exec(dt):
 * initialize -> interact(inner) -> interact(wall) -> update, ref: acoustic_step_1st_half.hpp:66-180,
 * interaction_algorithms_ck.cpp:29-34.  Two launches (initialize; fused interact+update). */
int sphb200_acoustic_1st_half(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float dt, void *stream);
/* ...AcousticStep2ndHalf...::exec(dt), one fused launch; ref: acoustic_step_2nd_half.hpp:33-137.
 * When next_reduced_dev != NULL the launch also leaves max_i AcousticTimeStepCK::reduce(i) of the NEW state
 * in *next_reduced_dev (device float, must be zeroed by the caller) so the next step needs no extra pass. */
int sphb200_acoustic_2nd_half(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float dt, float h_min,
                              float *next_reduced_dev, void *stream);
/* phase-granular entry points (parity tests against the reference's per-phase kernels) */
int sphb200_acoustic_1st_half_initialize(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float dt, void *stream);
int sphb200_acoustic_1st_half_interact(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float dt, int do_update,
                                       void *stream);
/* LinearCorrectionMatrix<Inner<WithUpdate>,Contact<>>; ref: general_dynamics/kernel_correction_ck.hpp:40-95.
 * Writes fluid.B and, when given, fluid.correction_record. */
int sphb200_linear_correction_matrix(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float alpha, void *stream);
/* correction_record (see sphb200_fluid_view_t) of n particles whose matrices (B) and / or pressures were written outside the
 * library (host uploads); a NULL source leaves its part of the records as it is */
int sphb200_pack_correction_records(sphb200_context_t *ctx, uint32_t n, const float *B, const float *pressure,
                                    void *correction_record, void *stream);
/* InteractionDynamicsCK<FreeSurfaceIndicationCK<Inner<WithUpdate>, Contact<>>>::exec: inner interact (position
 * divergence, spatial-temporal override next to the previous surface) -> contact interact -> update (Indicator,
 * PreviousSurfaceIndicator). threshold = 0.75 * Dimensions, smoothing_length = ReferenceSmoothingLength().
 * Two launches (update reads the neighbours' PositionDivergence).
 * ref: general_dynamics/surface_indication/surface_indication_ck.hpp:12-160 */
int sphb200_free_surface_indication(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, int32_t *indicator,
                                    float *position_divergence, int32_t *previous_indicator, float threshold,
                                    float smoothing_length, void *stream);
/* one sweep of the same (sweep 0: interact, writes PositionDivergence; sweep 1: update + very-near check, reads
 * PositionDivergence of the neighbours; surface_indication_ck.hpp:52-87,149-160 and :98-128): slab-decomposed runs refresh
 * PositionDivergence on the ghost planes between the two (new; sphb200_free_surface_indication = sweep 0 then sweep 1) */
int sphb200_free_surface_indication_sweep(sphb200_context_t *ctx, const sphb200_fluid_args_t *fluid, int32_t *indicator,
                                          float *position_divergence, int32_t *previous_indicator, float threshold,
                                          float smoothing_length, int sweep, void *stream);
/* Interpolation<Contact<DataType>>::InteractKernel::interact for an observer body:
 * out[i] = sum_j W_ij V_j data[j] / (sum_j W_ij V_j + TinyReal) over the rows of `rel` (observer -> observed body).
 * width 1: Real data; width 4: Vecd data as stored on the device (float4, all four lanes are interpolated).
 * ref: general_dynamics/interpolation_dynamics.hpp:44-60 */
int sphb200_interpolate(sphb200_context_t *ctx, const sphb200_kernel_t *kernel, const sphb200_vec4_t *src_pos, uint32_t n_src,
                        sphb200_relation_t rel, const sphb200_vec4_t *tar_posvol, const float *tar_data, int width, float *out,
                        void *stream);
/* Interpolation<Contact<DataType, RestoringCorrection>>::InteractKernel::interact: the first-order consistent interpolation
 * (constant and linear fields are reproduced on any neighbour set with a regular restoring matrix), same arguments;
 * out[i] = (restoring_i^-1).row(0) . prediction_i with restoring_i = Eps I + sum_j A_ij, prediction_i = sum_j A_ij.col(0) data[j].
 * ref: general_dynamics/interpolation_dynamics.hpp:72-100; known answer unit_test_interpolation_ck/2d_interpolation.cpp */
int sphb200_interpolate_restoring(sphb200_context_t *ctx, const sphb200_kernel_t *kernel, const sphb200_vec4_t *src_pos,
                                  uint32_t n_src, sphb200_relation_t rel, const sphb200_vec4_t *tar_posvol, const float *tar_data,
                                  int width, float *out, void *stream);
/* InteractionDynamicsCK<ViscousForceCK<Inner<WithUpdate, Viscosity, Correction>, Contact<Wall, Viscosity, Correction>>>::exec:
 * inner interact -> wall interact -> ForcePriorCK update (ForcePrior += F - Previous; Previous = F), one launch.
 * Needs fluid.posvolvel (the 32-byte gather record) and fluid.force_prior; material.correction selects the B-matrix form.
 * ref: fluid_dynamics/viscous_force.hpp:44-103, general_dynamics/force_prior_ck.h:53-57, materials/viscosity.h:40-67 */
int sphb200_viscous_force(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, float mu, float smoothing_length,
                          sphb200_vec4_t *viscous_force, sphb200_vec4_t *previous_viscous_force, void *stream);
/* KernelGradientIntegral<Inner<Correction>, Contact<Boundary, Correction>>: kgi_i = -sum (B_i + B_j) dW V_j e_ij
 * - sum_wall 2 B_i dW V_j e_ij; ref: general_dynamics/kernel_gradient_integral.hpp:33-78 */
int sphb200_kernel_gradient_integral(sphb200_context_t *ctx, const sphb200_fluid_args_t *a, sphb200_vec4_t *kernel_gradient_integral,
                                     void *stream);
/* StateDynamics<TransportVelocityCorrectionCK<SPHBody, Limiter, Scopes...>>: dpos_i += coefficient h^2 limiter(h^2 |kgi|^2) kgi_i.
 * limiter 0 NoLimiter, 1 TruncatedLinear(slope); indicator != NULL restricts the update to BulkParticles (Indicator == 0).
 * ref: fluid_dynamics/transport_velocity_correction_ck.hpp:39-50, common/common_functors.h:69-94 */
int sphb200_transport_velocity_correction(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid,
                                          const sphb200_vec4_t *kernel_gradient_integral, float coefficient, float h_ref, int limiter,
                                          float limiter_slope, const int32_t *indicator, void *stream);
/* ReduceDynamicsCK<TotalMechanicalEnergyCK>; ref: general_dynamics/general_reduce_ck.h:52-88 */
int sphb200_total_mechanical_energy(sphb200_context_t *ctx, const sphb200_fluid_view_t *fluid, const float gravity[3],
                                    double *energy_host, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * slab-decomposed multi-GPU runs (new: the reference has no distributed path). One process and one context per
 * GPU; the communicator is NCCL over NVLink/NVSwitch. Bootstrap: rank 0 calls sphb200_comm_unique_id and hands
 * the 128 bytes to the other ranks by whatever the launcher offers (torch.distributed, MPI, a file).
 * Storage follows the cell order with x the slowest axis, so the cell planes a neighbouring slab needs are
 * contiguous ranges of every variable array: exchanges send and receive those ranges in place.
 * ------------------------------------------------------------------------------------------------- */
#define SPHB200_UNIQUE_ID_BYTES 128
/* Slab decomposition (new): indices of the slots in [begin, begin + n) whose x cell plane
 * (Mesh::CellIndexFromPosition, base_mesh.hxx:9-15) is <= plane_left (left list) or >= plane_right (right list); a
 * negative plane switches the side off. counts[0..1] (device) receive the list lengths. List order is unspecified. */
int sphb200_slab_select(sphb200_context_t *ctx, const sphb200_mesh_t *mesh, const sphb200_vec4_t *pos, uint32_t begin,
                        uint32_t n, int plane_left, int plane_right, uint32_t *left_idx, uint32_t *right_idx,
                        uint32_t *counts, void *stream);
int sphb200_comm_unique_id(void *id128);
int sphb200_comm_create(sphb200_context_t *ctx, int nranks, int rank, const void *id128);
int sphb200_comm_destroy(sphb200_context_t *ctx);
int sphb200_comm_rank(const sphb200_context_t *ctx);
int sphb200_comm_size(const sphb200_context_t *ctx);
/* one grouped send/recv with rank-1 ("left") and rank+1 ("right"): `count` device segments per direction; what a
 * rank sends left arrives in its left neighbour's recv_right segments of the same index. Missing neighbours are skipped;
 * in a ring (sphb200_comm_set_ring) there are none missing: rank 0's left neighbour is rank nranks-1. */
int sphb200_comm_exchange(sphb200_context_t *ctx, int count, const void *const *send_left, const size_t *send_left_bytes,
                          void *const *recv_left, const size_t *recv_left_bytes, const void *const *send_right,
                          const size_t *send_right_bytes, void *const *recv_right, const size_t *recv_right_bytes, void *stream);
/* Communicator of ONE rank without NCCL; with the ring switched on the rank is its own neighbour on both sides (a
 * periodic box in one slab) and sphb200_comm_exchange copies device to device; reductions and all-gather are identities. */
int sphb200_comm_create_self(sphb200_context_t *ctx);
/* ring != 0 closes the slab chain (periodic along x, BASELINE config 4 on N GPUs): the neighbours of
 * sphb200_comm_exchange become (rank - 1) mod nranks and (rank + 1) mod nranks. */
int sphb200_comm_set_ring(sphb200_context_t *ctx, int ring);
int sphb200_comm_is_ring(const sphb200_context_t *ctx);
/* The periodic seam of a ring of slabs along x: the x axis of the body's mesh, the cell planes of the periodic box
 * [first_plane, first_plane + box_planes), and the x ranges (inclusive, in the mesh's own cell arithmetic,
 * base_mesh.hxx:9-15) of the box planes and of the one ghost plane on either side of the box. */
typedef struct
{
    float mesh_lower, mesh_spacing;
    int32_t mesh_cells, first_plane, box_planes;
    float own_min, own_max;
    float ghost_low_min, ghost_low_max;
    float ghost_high_min, ghost_high_max;
} sphb200_seam_t;
/* Positions that crossed the seam: x of n records `stride_bytes` apart (x = the first float of a record: Position,
 * PosVol, PosVolVel) += delta, delta = +/-(box_upper - box_lower): the translated position of the reference's periodic
 * ghost entry, particle_dynamics/general_dynamics/domian_bouding/domain_bounding.cpp:26,45. Records whose plane was
 * outside the box planes on the sender (leavers) end inside them, the others (its boundary plane) in the ghost plane. */
int sphb200_seam_shift(sphb200_context_t *ctx, void *base, uint32_t stride_bytes, uint32_t n, float delta,
                       const sphb200_seam_t *seam, void *stream);
/* Peer mailboxes (new): every rank owns four boxes of `box_bytes` (two parities x {from the left, from the right}); the
 * neighbours map them over CUDA IPC (NVLink peer access) and WRITE into them directly: sphb200_comm_push gathers the listed
 * slots of `count` variables — the list length is read from DEVICE memory, so no message size crosses the host — stores
 * them into the neighbour's box of parity (seq & 1) and releases the box by writing {count, seq} into its header
 * (system-scope fence, last block of the launch). sphb200_comm_pull waits on the device for the header of `seq`, appends the
 * box's records behind `dst_begin` (+ *dst_extra_dev) of each variable and leaves the count in *count_dev.
 * side: 0 = the left neighbour, 1 = the right neighbour (push: where it goes; pull: where it comes from).
 * Status bits (sphb200_comm_mailbox_status, device word): 1 = a push had more entries than the neighbour's box holds,
 * 2 = a pull timed out waiting for its neighbour, 4 = a pull ran out of room behind dst_begin. open/close are collective. */
int sphb200_comm_mailbox_open(sphb200_context_t *ctx, size_t box_bytes);
int sphb200_comm_mailbox_close(sphb200_context_t *ctx);
size_t sphb200_comm_mailbox_peer_bytes(const sphb200_context_t *ctx, int side);
const uint32_t *sphb200_comm_mailbox_status(const sphb200_context_t *ctx);
int sphb200_comm_push(sphb200_context_t *ctx, int side, int count, const void *const *src, const uint32_t *elem_bytes,
                      const uint32_t *idx, const uint32_t *n_dev, uint64_t seq, void *stream);
int sphb200_comm_pull(sphb200_context_t *ctx, int side, int count, void *const *dst, const uint32_t *elem_bytes,
                      uint32_t dst_begin, const uint32_t *dst_extra_dev, uint32_t dst_end, uint32_t *count_dev, uint64_t seq,
                      void *stream);
int sphb200_comm_allreduce_max_f32(sphb200_context_t *ctx, float *dev_inout, int n, void *stream);
int sphb200_comm_allreduce_sum_f64(sphb200_context_t *ctx, double *dev_inout, int n, void *stream);
int sphb200_comm_allgather_u64(sphb200_context_t *ctx, const uint64_t *dev_send, uint64_t *dev_recv, int n_per_rank, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SPHB200_H */
