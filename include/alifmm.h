/*
 * alifmm.h -- C ABI of the B200 ALI-FMM travel-time-field + ray-tracing path.
 *
 * The reference (WiPi-UoS/ALI-FMM-and-ray-tracing, Anis_TTF_rays.py = "ATR") has no FFI
 * layer: its hot path is numba-jitted Python called from class ALI_FMM (ATR:3789-4705).
 * This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md).  Every entry point names the reference interface it replaces.
 *
 * Conventions: all pointers are caller-owned HOST memory, C-contiguous, row-major
 * [z][x]; sizes are element counts; functions return 0 on success and a negative
 * ALIFMM_E_* code on failure, with a message available from alifmm_last_error()
 * (thread-local).  A context is bound to one CUDA device and must be used by one
 * thread at a time.  There is no CPU fallback: without a usable device
 * alifmm_create() fails.
 */
#ifndef ALIFMM_H
#define ALIFMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALIFMM_OK 0
#define ALIFMM_E_INVALID (-1)   /* bad argument */
#define ALIFMM_E_CUDA (-2)      /* CUDA runtime error (no device, out of memory, launch failure) */
#define ALIFMM_E_CAPACITY (-3)  /* an internal work list overflowed (see alifmm_set_option) */
#define ALIFMM_E_STATE (-4)     /* call order (e.g. rays before any field was computed) */

typedef struct alifmm_ctx alifmm_ctx;

/* The model arrays ALI_FMM.__init__ / update take (ATR:3793, 3870):
 *   veln     orientation in degrees              float64 [nz*nx]
 *   velpn    material id, 0 = stiffness tensors  int32   [nz*nx]
 *   vel_map  velocity scale                      float64 [nz*nx]
 *   stif_den (c22,c23,c33,c44 [MPa], rho)        int64   [nz*nx*5] or NULL
 *   has_stif the reference's "stif_den is not None" as seen by the jitted code (the
 *            class substitutes zeros for None, ATR:3890-3891, so it is normally 1)
 *   group_vel / phase_vel  velocity tables        float64 [361*n_cols], column 0 = angle */
typedef struct {
    int32_t nz, nx;
    double dnx;
    const double *veln;
    const int32_t *velpn;
    const double *vel_map;
    const int64_t *stif_den;
    int32_t has_stif;
    const double *group_vel;
    const double *phase_vel;
    int32_t n_cols;
} alifmm_model_desc;

typedef struct {
    int64_t node_solves;      /* final travel times produced (nodes x sources) */
    int64_t seq_pops;         /* nodes accepted by the sequential near-source replica */
    int64_t seq_evals;        /* ALI update evaluations in the sequential replica */
    int64_t band_rounds;      /* sum over sources of band-march rounds */
    int64_t band_rounds_max;  /* slowest source */
    int64_t band_evals;       /* ALI update evaluations in the band march */
    int64_t fallback_evals;   /* fouds18_A fallback evaluations (ATR:240) */
    int64_t max_band;         /* largest narrow band seen */
    int64_t rays;             /* rays traced by the last alifmm_rays call */
    int64_t ray_points;       /* path points written by the last alifmm_rays call */
    int64_t kernel_launches;  /* kernels launched by the last ttf / rays call */
    double ms_seq;            /* device time of the last call's kernels (CUDA events) */
    double ms_march;
    double ms_finalize;
    double ms_rays;
    double vmax;              /* model-wide phase-velocity bound used for delta */
    double delta;             /* acceptance band of the last field batch, seconds */
    int64_t cluster_size;     /* CTAs (SMs) per source in the band march of the last batch */
    int64_t seq_threads;      /* threads per source in the sequential near-source phase */
    double seq_mcycles_min;   /* fastest / slowest source of the last batch, SM cycles x 1e6: */
    double seq_mcycles_max;   /*   sequential phase */
    double march_mcycles_min; /*   band march (sum of its phases) */
    double march_mcycles_max;
} alifmm_counters_t;

/* Number of CUDA devices visible to the process (0 when there is none). */
int alifmm_device_count(void);

/* Uploads the model to `device` and keeps it resident.  Replaces the model hand-over
 * of ALI_FMM.update / update_i / find_all_TTF_rays (ATR:3889-3900, 4073-4076). */
int alifmm_create(const alifmm_model_desc *desc, int device, alifmm_ctx **out);
void alifmm_destroy(alifmm_ctx *ctx);

/* Options: "delta_frac" (acceptance band as a fraction of dnx/vmax, default 0.35, maximum 0.4: 0.1 ... 0.35 give
 * the same bits on every test model -- the reference algorithm's own solution on a correctly ordered heap,
 * PARITY.md --, 0.4 differs by up to 5e-12, 0.45 by 1e-7, 0.5 by 1e-4: tests/test_kernel_replay.py), "handover_margin" (nodes the sequential replica runs past the last
 * refined source box, default 27), "band_capacity_factor" (narrow-band list capacity as
 * a multiple of nz+nx of the solved grid, default 6), "threads_per_source" (CTA size of
 * the band march: 256, 512, 640, 768, 896 or 1024; default 768 = 80 registers per thread, measured
 * fastest on B200), "cluster_size" (CTAs = SMs per source in the band march: 0 = automatic, the largest of
 * 8 ... 1 that keeps every source of the batch resident at once -- 16 sources: 6, 32: 4, 64: 2; results do not depend on it),
 * "cluster_threads" (CTA size of the cluster march, 512 or 768; 0 = automatic: 768 for pairs, else 512), "ray_min_blocks"
 * (CTAs of four ray warps per SM the ray kernel is compiled for: 4 = 126 registers, the default; 5, 6 spill), "seq_threads" (CTA size of the sequential
 * near-source kernel: 32 = one warp per source, the default and measured fastest; 64, 128 or 256 evaluate one
 * speculated node per warp between two CTA barriers -- same result, ~8 % slower on B200), "band_smem_kb" (shared memory for the band lists, 0..180, default 0 = keep L1
 * for the field gathers), "resort_every" (re-order the band
 * list along the front every this many rounds so that warps coalesce, default 8; 0 = never). */
int alifmm_set_option(alifmm_ctx *ctx, const char *name, double value);

/* Runs the library's kernels on the caller's CUDA stream (a cudaStream_t passed as
 * void*); NULL restores the context's own stream. */
int alifmm_set_stream(alifmm_ctx *ctx, void *cuda_stream);

/* Travel-time fields of n_src sources at once.  Source k sits on coarse node
 * (src_iz[k], src_ix[k]) (the reference rounds scx/dnx, scz/dnx: ATR:1509-1510).
 * subgrid == 1 replaces travel() (ATR:1463); odd subgrid > 1 replaces
 * travel_finer_grid() (ATR:2120) and returns fields of (subgrid*(nz-1)+1) x
 * (subgrid*(nx-1)+1) nodes.  The fields stay resident on the device as slots
 * 0..n_src-1 until the next call; out_host (n_src fields back to back) may be NULL. */
int alifmm_ttf(alifmm_ctx *ctx, int32_t n_src, const int32_t *src_iz, const int32_t *src_ix, int32_t subgrid,
               double *out_host);

/* Copies one resident field to the host. */
int alifmm_ttf_fetch(alifmm_ctx *ctx, int32_t slot, double *out_host);

/* Extents of the resident fields. */
int alifmm_ttf_shape(alifmm_ctx *ctx, int32_t *n_slots, int32_t *fz, int32_t *fx, int32_t *subgrid);

/* Traces n_rays rays through resident fields; replaces find_ray() (ATR:3104) as called
 * from find_all_TTF_rays (ATR:4349) / parallel_TTF_rays (ATR:3728).  Ray r starts on
 * coarse node (src_iz[r], src_ix[r]) and runs to the source node of field rec_slot[r].
 * Outputs, per ray: out_x / out_y [capacity] path points in FINE-grid units (the caller
 * divides by subgrid, ATR:3729-3730), out_len, out_time (seconds), out_flag (bit 0: the
 * reference's "Travel time to receiver increasing" early exit, bit 1: plane left the
 * grid, bit 2: empty plane, bit 3: capacity reached).  capacity is the reference's
 * 5*(nz+nx) unless the caller wants less. */
int alifmm_rays(alifmm_ctx *ctx, int32_t n_rays, const int32_t *src_iz, const int32_t *src_ix,
                const int32_t *rec_slot, int32_t capacity, double *out_x, double *out_y, int32_t *out_len,
                double *out_time, int32_t *out_flag);

/* Same rays, delivered in the reference's dense layout (ALI_FMM.ray_paths_x / ray_paths_y,
 * ATR:4306-4307): ray r is written to base_x + row[r] * capacity (its first out_len[r] points,
 * each divided by `divisor` -- the reference's callers divide the fine-grid coordinates by
 * subgrid_size, ATR:3729-3730, 4355-4356); the rest of the row is left untouched.  Only the used
 * points cross PCIe (packed on the device, one copy through a pinned staging buffer). */
int alifmm_rays_into(alifmm_ctx *ctx, int32_t n_rays, const int32_t *src_iz, const int32_t *src_ix,
                     const int32_t *rec_slot, int32_t capacity, double divisor, const int64_t *row, double *base_x,
                     double *base_y, int32_t *out_len, double *out_time, int32_t *out_flag);

/* Device and pinned buffers of destroyed contexts are kept for the next context of the same
 * process (ALI_FMM creates one per call).  alifmm_trim(device) releases them (device < 0: all). */
int alifmm_trim(int device);

/* Free / total bytes of the context's device (used by the host side to size batches); buffers held
 * for reuse count as free. */
int alifmm_mem_info(alifmm_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes);

/* Work counters and device timings of the last calls (feeds bench.py's roofline). */
int alifmm_counters(alifmm_ctx *ctx, alifmm_counters_t *out);

/* Christoffel curves of one material, 361 samples at 1 degree; replaces
 * ALI_FMM.generate_group_vel / generate_phase_vel (ATR:4112-4206).  Stiffness in Pa. */
int alifmm_velocity_curves(alifmm_ctx *ctx, double c22, double c23, double c33, double c44, double density,
                           double *group_out, double *phase_out);

/* Model sanity scan; replaces min_max_vel (ATR:3736-3787). */
int alifmm_min_max_vel(alifmm_ctx *ctx, double *min_vel, double *max_vel);

/* Christoffel curves of n_mat materials in one launch and one copy each way; replaces the per-material
 * loop of ALI_FMM.add_materials (ATR:4208-4256 calling ATR:4112-4206).  props[n_mat][5] =
 * (c22, c23, c33, c44 [Pa], density); outputs [n_mat][361] (row m = the 361 samples of material m).
 * Needs no context (the class builds its tables before any model is uploaded). */
int alifmm_velocity_curves_batch(int device, int32_t n_mat, const double *props, double *group_out, double *phase_out);

/* Node-level operator check (used by the parity tests): runs the device restatements of update()
 * (ATR:904-1410, with wavefront_angle_dist ATR:1413-1460) and of fouds18_A() (ATR:240-901) on n
 * independent caller-supplied states.  State c is a grid of nz x nx nodes: model arrays
 * veln/velpn/vel_map [n][nz*nx] (+ stif_den [n][nz*nx][5] or NULL), travel times ttn [n][nz*nx],
 * statuses nsts [n][nz*nx] (-1 far, 0 alive, >0 narrow band, ATR:103) and the node pos[c] = (iz, ix).
 * Outputs per state: out_update (-1.0 = no stencil, ATR:1408-1410), out_fouds (the value
 * fouds18_A returns, including its min with ttn[iz, ix], ATR:898-899), out_stencil (may be NULL;
 * 0-15 = stencil used, -1/-2 none as in ATR:989-1366). */
int alifmm_eval_nodes(int device, int32_t n, int32_t nz, int32_t nx, double dnx, const double *veln,
                      const int32_t *velpn, const double *vel_map, const int64_t *stif_den, int32_t has_stif,
                      const double *group_vel, const double *phase_vel, int32_t n_cols, const double *ttn,
                      const int32_t *nsts, const int32_t *pos, double *out_update, double *out_fouds,
                      int32_t *out_stencil);

/* One travel-time field (travel(), ATR:1463; subgrid 1) decomposed into row strips on 2 ... 8 GPUs (BASELINE config 5): each
 * device holds its strip of the model and of the field plus a halo of one tile row either side (the stencil reaches two
 * nodes, ATR:940-987); values published next to a boundary and claims of nodes across it go to the neighbouring strip,
 * the per-round minimum / band length / termination to every strip, all through peer-mapped memory over NVLink (no NCCL
 * on the data path).  The reference has no counterpart (one field is one heap); the result is bit-identical to alifmm_ttf()
 * on one device.  The value is capacity, not speed.  devices: n_dev distinct, mutually peer-accessible devices, strip k
 * (rows from the top) on devices[k]; split_row <= 0: automatic (equal shares, boundaries on multiples of 4, kept away
 * from the source so that its refined neighbourhood lies inside one strip), > 0 only with two devices (rounded down to a
 * multiple of 4; an inadmissible row is an error, not adjusted); out_host
 * [nz * nx]; counters may be NULL. */
int alifmm_ttf_split(const alifmm_model_desc *desc, int32_t n_dev, const int32_t *devices, int32_t src_iz, int32_t src_ix,
                     int32_t split_row, double *out_host, alifmm_counters_t *counters);

/* The rows alifmm_ttf_split gives to its strips: rows[k] ... rows[k + 1] - 1 for strip k, rows[0] = 0, rows[n_dev] = nz
 * (rows: n_dev + 1 entries).  Host arithmetic only; the same errors as alifmm_ttf_split for an inadmissible request. */
int alifmm_split_rows(int32_t nz, int32_t n_dev, int32_t src_iz, int32_t split_row, int32_t *rows);

const char *alifmm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* ALIFMM_H */
