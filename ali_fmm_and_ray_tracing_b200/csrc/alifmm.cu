// alifmm.cu -- kernels and C ABI (include/alifmm.h) of the B200 ALI-FMM path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//             --shared -Xcompiler -fPIC  (see __graft_entry__.build()).
//
// Kernels
//   ali_vmax_kernel      model-wide phase-velocity bound (for the acceptance band delta)
//   ali_seq_kernel       one warp per source: exact sequential replica of the reference's
//                        nested near-source grids + main-grid start (ali_seq.cuh)
//   ali_march_kernel     one CTA per source: band-synchronous narrow-band march
//                        (ali_band.cuh), lists compacted with warp ballots
//   ali_finalize_kernel  fine path: T / subgrid (ATR:2832)
//   ali_rays_kernel      one warp per ray: plane-marching Fermat search (ali_ray.cuh)
//   ali_curves_kernel / ali_minmax_kernel   material curves and model sanity scan
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <algorithm>
#include <mutex>
#include <thread>

#include "../../include/alifmm.h"
#include "ali_core.cuh"
#include "ali_seq.cuh"
#include "ali_band.cuh"
#include "ali_ray.cuh"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(ALIFMM_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
    } while (0)

// ---------------------------------------------------------------------------
// device-side per-source records
// ---------------------------------------------------------------------------
struct AliSourceRec {
    int src_iz, src_ix;     // coarse node
    AliSeqResult seq;       // window + counters of the sequential phase
    long long rounds, band_evals, band_fallbacks, max_band;
    long long cycles[4];    // SM cycles spent in phases A0 / A1 / B / C (thread 0's view)
    int overflow;           // 1: sequential heap/window, 2: band list
};

struct AliBatch {
    AliModel m;
    const AliModel *m_dev;  // the same struct in device memory (for out-of-line slow paths)
    int sg;
    int nz, nx;             // extents of the solved grid
    int margin;
    double delta;
    double *T;              // [n_src][nz*nx] result, row-major (written by the finalize kernel)
    double *Tt;             // [n_src][tn] field the band march works on, 4 x 4-node tiles (ali_band.cuh)
    uint8_t *st;            // [n_src][tn] alive flags, same indexing
    size_t tn;              // ali_field_nodes_tiled(nz, nx)
    // sequential scratch, per source
    double *seq_t;          // [n_src][2*seq_cap]
    int32_t *seq_s;         // [n_src][2*seq_cap]
    int32_t *seq_heap;      // [n_src][2*heap_cap]
    double *seq_hkey;       // [n_src][heap_cap] travel times of the heap entries
    double *seq_cval;       // [n_src][seq_cap] evaluation cache of the cooperative march
    uint8_t *seq_cflag;     // [n_src][seq_cap]
    size_t seq_cap;
    int heap_cap;
    // band lists, per source
    unsigned *lists;        // [n_src][2*band_cap] packed (iz << 16 | ix)
    double *stage;          // [n_src][band_cap]
    int band_cap;
    int resort_every;       // re-sort the band list along the front every this many rounds (0: never)
    AliSourceRec *rec;      // [n_src]
};

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// Builds the 64-byte per-node material records from the caller's arrays.
__global__ void ali_records_kernel(int n, const double *veln, const int32_t *velpn, const double *vel_map,
                                   const long long *stif, AliMatRec *rec, int *bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        AliMatRec r;
        r.veln = veln[i]; r.vel_map = vel_map[i]; r.velpn = velpn[i]; r.pad = 0;
        if (bad && (!isfinite(r.veln) || !isfinite(r.vel_map))) atomicOr(bad, 1);
        for (int k = 0; k < 5; k++) r.s[k] = stif ? (double)stif[(size_t)5 * i + k] : 0.0;
        rec[i] = r;
    }
}

__global__ void ali_vmax_kernel(AliModel m, unsigned long long *out_bits)
{
    const size_t n = (size_t)m.nz * m.nx;
    double best = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int iz = (int)(i / m.nx), ix = (int)(i % m.nx);
        double v = ali_node_vmax(m, iz, ix);
        if (v > best) best = v;
    }
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(best));
}

// One CTA per source: thread 0 walks the reference's heap loop, warp 0 finds the evaluations worth doing
// ahead, and every warp evaluates one of them per step (ali_seq_march_coop).  With blockDim.x == 32 the
// one-warp form of round 1 runs instead.
#define ALI_SEQ_MAX_THREADS 256
__global__ void __launch_bounds__(ALI_SEQ_MAX_THREADS, 1) ali_seq_kernel(AliBatch b)
{
    const int src = blockIdx.x;
    const int lane = threadIdx.x;
    AliSourceRec &rec = b.rec[src];
    AliSourcePlan p;
    ali_make_plan(p, b.m, rec.src_iz, rec.src_ix, b.sg, b.margin);
    AliSeqScratch sc;
    sc.tA = b.seq_t + (size_t)src * 2 * b.seq_cap;
    sc.tB = sc.tA + b.seq_cap;
    sc.sA = b.seq_s + (size_t)src * 2 * b.seq_cap;
    sc.sB = sc.sA + b.seq_cap;
    sc.heap = reinterpret_cast<AliHeapEnt *>(b.seq_heap) + (size_t)src * b.heap_cap;
    sc.hkey = b.seq_hkey + (size_t)src * ALI_HKEY_SLOTS(b.heap_cap);
    sc.cval = b.seq_cval + (size_t)src * b.seq_cap;
    sc.cflag = b.seq_cflag + (size_t)src * b.seq_cap;
    sc.heap_cap = b.heap_cap;
    sc.status_cap = b.seq_cap;
    AliSeqResult res;
    ali_seq_source(b.m, p, sc, res, lane, (int)blockDim.x);
    if (lane == 0) {
        rec.seq = res;
        rec.overflow = res.overflow ? 1 : 0;
    }
}

// Warp-aggregated append of k (0..4) items per lane to a list with a shared counter.
__device__ __forceinline__ int ali_warp_reserve(int k, int *counter)
{
    const unsigned lane = threadIdx.x & 31;
    int incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - k;
}

// Two warp-aggregated appends at once (next-list slot and work-list slot).
__device__ __forceinline__ void ali_warp_reserve2(int k, int kw, int *counter, int *wcounter, int &pos, int &wpos)
{
    const unsigned lane = threadIdx.x & 31;
    int incl = k | (kw << 16); // k <= 4, kw <= 4: both prefix sums in one scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += v;
    }
    int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0, wbase = 0;
    if (lane == 31) {
        if (total & 0xffff) base = atomicAdd(counter, total & 0xffff);
        if (total >> 16) wbase = atomicAdd(wcounter, total >> 16);
    }
    base = __shfl_sync(0xffffffffu, base, 31);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    pos = base + (incl & 0xffff) - k;
    wpos = wbase + (incl >> 16) - kw;
}

// Band-synchronous march of one source per CTA (ali_band.cuh).  The narrow band lives in
// shared memory when it fits (packed entries and the round's work list, 16 bytes per band
// node; their values stay in global memory so that L1 keeps room for the T / status
// gathers), else in the global buffers of the batch.
//   round r:  A  evaluate the work list (band nodes whose window changed) -> value[i]
//             -- barrier --
//             B  publish those values (T, status, dirty marks); tmin = min(all band values)
//             -- barrier --
//             C  accept value <= tmin + delta (alive, enlist far neighbours), compact the
//                survivors into the other buffer, build the next work list and the minimum
//                of the values that will not be re-evaluated
//             -- barrier --
// Angular bin of a band node around the source: entries sorted by it are ordered along the
// front, so the lanes of a warp work on neighbouring nodes and their gathers coalesce.
#define ALI_SORT_BINS 2048
__device__ __forceinline__ int ali_sort_bin(unsigned e, int isz, int isx)
{
    float a = atan2f((float)(ALI_PACK_Z(e) - isz), (float)(ALI_PACK_X(e) - isx)); // (-pi, pi]
    int k = (int)((a + 3.14159265f) * (float)(ALI_SORT_BINS / 6.2831853f));
    return k < 0 ? 0 : (k >= ALI_SORT_BINS ? ALI_SORT_BINS - 1 : k);
}

// "Window changed" marks of one round live in shared memory as a bitmap over the grid folded
// modulo 512 x 512 nodes (32 KB).  A node that publishes a new value sets the bits of its 12
// window neighbours; phase C tests one bit per surviving band node.  Two nodes 512 apart share
// a bit: such a false mark only costs a re-evaluation, which returns the same value (the update
// is a pure function of the window).  This replaces 12 scattered global byte stores per
// published node and a global load + store per band node per round.
#define ALI_DMAP_BITS 9
#define ALI_DMAP_WORDS ((1 << (2 * ALI_DMAP_BITS)) / 32)
#define ALI_MARCH_SMEM_DMAP (ALI_DMAP_WORDS * 4)
#define ALI_MARCH_SMEM_FIXED (ALI_MARCH_SMEM_DMAP + ALI_MT_WORDS * 8)
__device__ __forceinline__ void ali_dmap_row(unsigned *dmap, int z, int x0, unsigned pattern)
{
    const unsigned m = (1u << ALI_DMAP_BITS) - 1u;
    const unsigned row = ((unsigned)z & m) << (ALI_DMAP_BITS - 5);
    const unsigned xs = (unsigned)x0 & m;
    const unsigned long long bits = (unsigned long long)pattern << (xs & 31u);
    const unsigned w0 = xs >> 5;
    atomicOr(&dmap[row | w0], (unsigned)bits);
    const unsigned hi = (unsigned)(bits >> 32);
    if (hi) atomicOr(&dmap[row | ((w0 + 1u) & ((1u << (ALI_DMAP_BITS - 5)) - 1u))], hi);
}

// Sets the bits of the 12 window neighbours of (iz, ix) (slot layout of ali_core.cuh).
__device__ __forceinline__ void ali_dmap_mark(unsigned *dmap, int iz, int ix)
{
    ali_dmap_row(dmap, iz - 2, ix - 2, 0x04u);
    ali_dmap_row(dmap, iz - 1, ix - 2, 0x0eu);
    ali_dmap_row(dmap, iz, ix - 2, 0x1bu);
    ali_dmap_row(dmap, iz + 1, ix - 2, 0x0eu);
    ali_dmap_row(dmap, iz + 2, ix - 2, 0x04u);
}

__device__ __forceinline__ bool ali_dmap_test(const unsigned *dmap, int iz, int ix)
{
    const unsigned m = (1u << ALI_DMAP_BITS) - 1u;
    const unsigned xs = (unsigned)ix & m;
    return (dmap[(((unsigned)iz & m) << (ALI_DMAP_BITS - 5)) | (xs >> 5)] >> (xs & 31u)) & 1u;
}

template <int NT>
__global__ void __launch_bounds__(NT) ali_march_kernel(AliBatch b, int smem_cap)
{
    // dynamic shared memory: window-change bitmap | sin/cos + atan tables (ali_glibcmath.cuh) | optional band lists
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned *s_dmap = reinterpret_cast<unsigned *>(s_raw);
    uint64_t *s_sincos = reinterpret_cast<uint64_t *>(s_raw + ALI_MARCH_SMEM_DMAP);
    unsigned char *s_lists = s_raw + ALI_MARCH_SMEM_FIXED;
    const int src = blockIdx.x;
    const int tid = threadIdx.x;
    AliSourceRec &rec = b.rec[src];
    __shared__ int s_count[2];
    __shared__ int s_nwork[2];
    __shared__ unsigned long long s_evalmin[2], s_basemin[2];
    __shared__ int s_overflow;
    __shared__ unsigned long long s_evals, s_fbs;
    __shared__ AliBandGrid s_grid;   // copy for the out-of-line FD fallback
    __shared__ int s_bins[ALI_SORT_BINS];
    __shared__ int s_wsum[32];
    __shared__ int s_force[2];   // by round parity: re-evaluate the whole band (FD fallback outside the register path)
    __shared__ long long s_cyc[4], s_tprev;   // phase timing, kept by thread 0 in shared memory (no registers in the hot loop)
    const int isz = (b.sg > 1 ? b.sg : 1) * rec.src_iz, isx = (b.sg > 1 ? b.sg : 1) * rec.src_ix;

    AliBandGrid g;
    g.nz = b.nz; g.nx = b.nx;
    g.T = b.Tt + (size_t)src * b.tn;
    g.st = b.st + (size_t)src * b.tn;
    g.t4x = (b.nx + 3) >> 2;
    g.dirty = nullptr;   // window-changed marks live in shared memory here (s_dmap); the byte map is the host replay's
    g.tiles_x = 0;
    g.dnx = b.m.dnx;
    g.mv = ali_band_view(b.sg);

    // band storage: value[2][cap] f64 (global), entry[2][cap] u32, work[2][cap] (index into entry)
    const int cap = b.band_cap;
    double *val0 = b.stage + (size_t)src * 2 * cap, *val1 = val0 + cap;
    unsigned *ent0, *ent1, *wrk0, *wrk1;
    if (smem_cap >= cap) {
        ent0 = (unsigned *)s_lists; ent1 = ent0 + cap;
        wrk0 = ent1 + cap; wrk1 = wrk0 + cap;
    } else {
        ent0 = b.lists + (size_t)src * 4 * cap; ent1 = ent0 + cap;
        wrk0 = ent1 + cap; wrk1 = wrk0 + cap;
    }

    for (int q = tid; q < ALI_MT_WORDS; q += NT) s_sincos[q] = q < ALI_GL_SINCOSTAB_COUNT ? ali_gl_sincostab[q] : ali_gl_atan_cij[q - ALI_GL_SINCOSTAB_COUNT];
    if (tid == 0) {
        s_grid = g;
        s_count[0] = 0; s_count[1] = 0; s_nwork[0] = 0; s_nwork[1] = 0;
        s_evalmin[0] = ~0ull; s_evalmin[1] = ~0ull; s_basemin[0] = ~0ull; s_basemin[1] = ~0ull;
        s_overflow = rec.overflow;
        s_evals = 0; s_fbs = 0;
        s_force[0] = 0; s_force[1] = 0;
        s_cyc[0] = s_cyc[1] = s_cyc[2] = s_cyc[3] = 0;
    }
    __syncthreads();
    if (s_overflow) return;

    // hand-over: window of the sequential phase (statuses + travel times in its level buffers) ->
    // tiled field + alive flags + first band list; every band node starts in the work list
    {
        const AliSeqResult w = rec.seq;
        const int nlev = b.sg > 1 ? 2 : 3;
        const size_t woff = (size_t)src * 2 * b.seq_cap + ((((nlev - 1) & 1) == 0) ? b.seq_cap : 0);
        const int32_t *wst = b.seq_s + woff;
        const double *wt = b.seq_t + woff;
        const int wn = w.wnz * w.wnx;
        for (int base = 0; base < wn; base += NT) {
            int i = base + tid;
            int k = 0;
            unsigned entry = 0;
            double tv = 0.0;
            if (i < wn) {
                int z = i / w.wnx, x = i - z * w.wnx;
                int32_t s = wst[i];
                if (s >= 0) {   // far nodes keep the field's pre-filled "far" word
                    tv = wt[i];
                    const size_t node = g.ti(w.wz0 + z, w.wx0 + x);
                    g.T[node] = tv;
                    if (s == 0) g.st[node] = ALI_ST_ALIVE;
                    else { k = 1; entry = ALI_PACK(w.wz0 + z, w.wx0 + x); }
                }
            }
            int pos = ali_warp_reserve(k, &s_count[0]);
            if (k) {
                if (pos < cap) { ent0[pos] = entry; wrk0[pos] = (unsigned)pos; val0[pos] = tv; }
                else s_overflow = 2;
            }
        }
    }
    __syncthreads();
    if (tid == 0) s_nwork[0] = s_count[0];
    __syncthreads();

    int rounds = 0, max_band = 0;
    unsigned my_evals = 0, my_fbs = 0;   // per thread: a field has fewer than 2^31 nodes
    int cur = 0;
    while (true) {
        const int n = s_count[cur];
        if (n == 0 || s_overflow) break;
        const int nwork = s_nwork[cur];
        double *val = cur == 0 ? val0 : val1, *nval = cur == 0 ? val1 : val0;
        unsigned *ent = cur == 0 ? ent0 : ent1, *nent = cur == 0 ? ent1 : ent0;
        unsigned *wrk = cur == 0 ? wrk0 : wrk1, *nwrk = cur == 0 ? wrk1 : wrk0;
        rounds++;
        if (n > max_band) max_band = n;
        if (tid == 0) s_tprev = clock64();
        // phase A: evaluate the work list from the round's snapshot.  The first two items of a thread
        // stay in registers for phase B (a round rarely has more than 2 * NT items); the minimum of
        // the new values is reduced here so that phase B has nothing to wait for.
        if (tid == 0) { s_count[cur ^ 1] = 0; s_nwork[cur ^ 1] = 0; s_basemin[cur ^ 1] = ~0ull; s_evalmin[cur ^ 1] = ~0ull; }
        for (int q = tid; q < ALI_DMAP_WORDS / 4; q += NT) reinterpret_cast<uint4 *>(s_dmap)[q] = make_uint4(0u, 0u, 0u, 0u);
        double lmin = 1e300;
        unsigned pe0 = 0, pe1 = 0;
        double pv0 = 0.0, pv1 = 0.0;
        int pmask = 0, it = 0;
        for (int q = tid; q < nwork; q += NT, it++) {
            const int i = (int)wrk[q];
            const unsigned e = ent[i];
            const int iz = ALI_PACK_Z(e), ix = ALI_PACK_X(e);
            int fb = 0;
            const double vold = val[i];   // the value this node last published (0: none yet)
            double v = ali_band_eval(b.m, b.m_dev, g, &s_grid, iz, ix, &fb, s_sincos);
            if (v != v) v = __longlong_as_double(ALI_T_NAN_VALUE_BITS);   // never the far / enlisted codes
            val[i] = v;
            my_evals++;
            my_fbs += fb;
            lmin = fmin(lmin, v);
            if (it == 0) { pe0 = e; pv0 = v; pmask |= (v != vold ? 1 : 0) | (fb ? 4 : 0); }
            else if (it == 1) { pe1 = e; pv1 = v; pmask |= (v != vold ? 2 : 0) | (fb ? 8 : 0); }
            else {
                // rare: more than 2 * NT items in one round; published from memory in phase B
                if (v != vold) val[i] = -v;          // sign marks "changed" until phase B
                if (fb) s_force[rounds & 1] = 1;
            }
        }
        for (int o = 16; o > 0; o >>= 1) lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        if ((tid & 31) == 0 && lmin < 1e300)
            atomicMin(&s_evalmin[cur], (unsigned long long)__double_as_longlong(lmin));
        __syncthreads();
        if (tid == 0) { const long long now = clock64(); s_cyc[0] += now - s_tprev; s_tprev = now; }
        // phase B: publish the re-evaluated values that changed and mark their window neighbours.
        // A node the FD fallback evaluated also depends on alive flags: it is re-evaluated every
        // round (its own bit), like the reference re-evaluates it on every neighbouring pop.
        if (tid == 0) s_force[(rounds + 1) & 1] = 0;
        if (pmask & 1) { g.T[g.ti(ALI_PACK_Z(pe0), ALI_PACK_X(pe0))] = pv0; ali_dmap_mark(s_dmap, ALI_PACK_Z(pe0), ALI_PACK_X(pe0)); }
        if (pmask & 2) { g.T[g.ti(ALI_PACK_Z(pe1), ALI_PACK_X(pe1))] = pv1; ali_dmap_mark(s_dmap, ALI_PACK_Z(pe1), ALI_PACK_X(pe1)); }
        if (pmask & 4) ali_dmap_row(s_dmap, ALI_PACK_Z(pe0), ALI_PACK_X(pe0), 1u);
        if (pmask & 8) ali_dmap_row(s_dmap, ALI_PACK_Z(pe1), ALI_PACK_X(pe1), 1u);
        for (int q = tid + 2 * NT; q < nwork; q += NT) {
            const int i = (int)wrk[q];
            const unsigned e = ent[i];
            const double v = val[i];
            if (v < 0.0) {
                val[i] = -v;
                g.T[g.ti(ALI_PACK_Z(e), ALI_PACK_X(e))] = -v;
                ali_dmap_mark(s_dmap, ALI_PACK_Z(e), ALI_PACK_X(e));
            }
        }
        __syncthreads();
        if (tid == 0) { const long long now = clock64(); s_cyc[1] += now - s_tprev; s_tprev = now; }
        // phase C: accept + extend the band; compact survivors; next work list (deferred to the
        // re-sort pass on the rounds that re-order the list along the front)
        const bool resort = b.resort_every > 0 && (rounds % b.resort_every) == 0;
        const int force = s_force[rounds & 1];
        const unsigned long long tminb = s_evalmin[cur] < s_basemin[cur] ? s_evalmin[cur] : s_basemin[cur];
        const double thr = __longlong_as_double((long long)tminb) + b.delta;
        double bmin = 1e300;
        // software pipeline over a thread's entries: (entry, value) are loaded two passes ahead and the four
        // neighbour state words of an entry that will be accepted one pass ahead, so that the L2 round trips of
        // consecutive passes overlap instead of forming one chain
        unsigned e1 = 0, e2 = 0;
        double v1 = 0.0, v2 = 0.0;
        unsigned long long sw = 0ull, se = 0ull, sn = 0ull, ss = 0ull;
        if (tid < n) { e1 = ent[tid]; v1 = val[tid]; }
        if (tid + NT < n) { e2 = ent[tid + NT]; v2 = val[tid + NT]; }
        if (tid < n && !(v1 > thr)) ali_band_accept_peek(g, ALI_PACK_Z(e1), ALI_PACK_X(e1), sw, se, sn, ss);
        for (int base = 0; base < n; base += NT) {
            const int i = base + tid;
            int k = 0, kw = 0;
            unsigned out[4];
            double v = 0.0;
            const unsigned e = e1;
            const double v_mine = v1;
            const unsigned long long cw_ = sw, ce_ = se, cn_ = sn, cs_ = ss;
            e1 = e2; v1 = v2;
            if (i + 2 * NT < n) { e2 = ent[i + 2 * NT]; v2 = val[i + 2 * NT]; }
            if (i + NT < n && !(v1 > thr)) ali_band_accept_peek(g, ALI_PACK_Z(e1), ALI_PACK_X(e1), sw, se, sn, ss);
            if (i < n) {
                v = v_mine;
                const int iz = ALI_PACK_Z(e), ix = ALI_PACK_X(e);
                if (!(v > thr)) {   // also a NaN value (degenerate material): the reference pops it too; never loops
                    k = ali_band_accept_peeked(g, iz, ix, cw_, ce_, cn_, cs_, out); // new nodes: always evaluated next round
                    kw = k;
                    v = 0.0;
                } else {
                    out[0] = e; k = 1;
                    if (!resort) {
                        if (force || ali_dmap_test(s_dmap, iz, ix)) kw = 1;
                        else bmin = fmin(bmin, v);
                    }
                }
            }
            int pos, wpos;
            ali_warp_reserve2(k, kw, &s_count[cur ^ 1], &s_nwork[cur ^ 1], pos, wpos);
            if (pos + k <= cap) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q < k) {
                        nent[pos + q] = out[q];
                        nval[pos + q] = v;
                        if (q < kw) nwrk[wpos + q] = (unsigned)(pos + q);
                    }
            } else if (k) {
                s_overflow = 2;
            }
        }
        for (int o = 16; o > 0; o >>= 1) bmin = fmin(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
        if ((tid & 31) == 0 && bmin < 1e300)
            atomicMin(&s_basemin[cur ^ 1], (unsigned long long)__double_as_longlong(bmin));
        __syncthreads();
        if (tid == 0) { const long long now = clock64(); s_cyc[2] += now - s_tprev; s_tprev = now; }
        if (resort && !s_overflow) {   // (any list length: phase C left the work list / base minimum to this pass)
            // counting sort of the new list by angular bin, from the "next" buffers back into the
            // current ones (free now); then the work list / base minimum on the sorted order
            const int nn = s_count[cur ^ 1];
            for (int q = tid; q < ALI_SORT_BINS; q += NT) s_bins[q] = 0;
            __syncthreads();
            for (int i = tid; i < nn; i += NT) atomicAdd(&s_bins[ali_sort_bin(nent[i], isz, isx)], 1);
            __syncthreads();
            {   // exclusive scan of the bins (ALI_SORT_BINS / NT bins per thread)
                constexpr int PER = (ALI_SORT_BINS + NT - 1) / NT;
                int loc[PER];
                int sum = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    int idx = tid * PER + q;
                    loc[q] = idx < ALI_SORT_BINS ? s_bins[idx] : 0;
                    sum += loc[q];
                }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((tid & 31) >= o) incl += v;
                }
                if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
                __syncthreads();
                if (tid < 32) {
                    int w = tid < NT / 32 ? s_wsum[tid] : 0;
                    int wi = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        int v = __shfl_up_sync(0xffffffffu, wi, o);
                        if (tid >= o) wi += v;
                    }
                    s_wsum[tid] = wi - w;
                }
                __syncthreads();
                int run = s_wsum[tid >> 5] + incl - sum;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    int idx = tid * PER + q;
                    if (idx < ALI_SORT_BINS) s_bins[idx] = run;
                    run += loc[q];
                }
            }
            if (tid == 0) { s_nwork[cur] = 0; s_basemin[cur] = ~0ull; }
            __syncthreads();
            for (int i = tid; i < nn; i += NT) {
                const unsigned e = nent[i];
                const int pos = atomicAdd(&s_bins[ali_sort_bin(e, isz, isx)], 1);
                ent[pos] = e;
                val[pos] = nval[i];
            }
            __syncthreads();
            double bm = 1e300;
            for (int base = 0; base < nn; base += NT) {
                const int i = base + tid;
                int kw = 0;
                if (i < nn) {
                    const unsigned e = ent[i];
                    const double v = val[i];
                    if (v == 0.0) kw = 1;           // new node: no estimate yet
                    else {
                        if (force || ali_dmap_test(s_dmap, ALI_PACK_Z(e), ALI_PACK_X(e))) kw = 1;
                        else bm = fmin(bm, v);
                    }
                }
                int wpos = ali_warp_reserve(kw, &s_nwork[cur]);
                if (kw) wrk[wpos] = (unsigned)i;
            }
            for (int o = 16; o > 0; o >>= 1) bm = fmin(bm, __shfl_xor_sync(0xffffffffu, bm, o));
            if ((tid & 31) == 0 && bm < 1e300)
                atomicMin(&s_basemin[cur], (unsigned long long)__double_as_longlong(bm));
            if (tid == 0) { s_count[cur] = nn; s_evalmin[cur] = ~0ull; } // the list stays in the current buffers
            __syncthreads();
        } else {
            cur ^= 1;
        }
        if (tid == 0) {
            const long long now = clock64(); s_cyc[3] += now - s_tprev;
        }
    }
    atomicAdd(&s_evals, (unsigned long long)my_evals);
    atomicAdd(&s_fbs, (unsigned long long)my_fbs);
    __syncthreads();
    if (tid == 0) {
        rec.rounds = rounds;
        rec.max_band = max_band;
        rec.band_evals = (long long)s_evals;
        rec.band_fallbacks = (long long)s_fbs;
        for (int q = 0; q < 4; q++) rec.cycles[q] = s_cyc[q];
        if (s_overflow) rec.overflow = s_overflow;
    }
}

// ---------------------------------------------------------------------------
// The same march with a thread-block CLUSTER per source (2, 4 or 8 CTAs = SMs): for batches with
// fewer sources than SMs (128 sources sharded over several GPUs leave 9 SMs per source on each).
// Every CTA works on a slice of the source's band list each phase; the three barriers of a round
// become cluster barriers (barrier.cluster arrive.release / wait.acquire, which also invalidates
// L1, so plain loads after it see the other CTAs' stores).  Everything the CTAs share lives in
// global memory (L2: 234-262 cycles, the same as a remote shared-memory access on this part):
// the band lists and their values as before, and a small control block per source -- list
// counters, the round's minima, the window-change bitmap and the sort bins -- touched only with
// atomics and volatile loads.  The set of nodes evaluated, published, accepted and enlisted in a
// round does not depend on how the lists are split, so the field is bit-identical to the
// one-CTA kernel's (tests/test_gpu_parity.py::test_cluster_march_equals_single_cta).
struct AliClusterCtl {
    int count[2], nwork[2];
    unsigned long long evalmin[2], basemin[2];
    int force[2];
    int overflow, pad;
    int bins[ALI_SORT_BINS];
    unsigned dmap[ALI_DMAP_WORDS];
};

__device__ __forceinline__ int ali_ldv(const int *p) { return *(const volatile int *)p; }
__device__ __forceinline__ unsigned long long ali_ldv(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
__device__ __forceinline__ bool ali_dmap_test_g(const unsigned *dmap, int iz, int ix)
{
    const unsigned m = (1u << ALI_DMAP_BITS) - 1u;
    const unsigned xs = (unsigned)ix & m;
    return (*(const volatile unsigned *)&dmap[(((unsigned)iz & m) << (ALI_DMAP_BITS - 5)) | (xs >> 5)] >> (xs & 31u)) & 1u;
}

__device__ __forceinline__ void ali_cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned ali_cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned ali_cluster_size()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}

template <int NT>
__global__ void __launch_bounds__(NT) ali_march_cluster_kernel(AliBatch b, AliClusterCtl *ctl_all)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t *s_sincos = reinterpret_cast<uint64_t *>(s_raw);
    const int C = (int)ali_cluster_size(), rank = (int)ali_cluster_rank();
    const int src = blockIdx.x / C;
    const int tid = threadIdx.x, gtid = rank * NT + tid, GT = C * NT;
    AliSourceRec &rec = b.rec[src];
    AliClusterCtl *ctl = ctl_all + src;
    __shared__ unsigned long long s_evals, s_fbs;
    __shared__ AliBandGrid s_grid;   // copy for the out-of-line FD fallback
    __shared__ int s_wsum[32];
    __shared__ long long s_cyc[4], s_tprev;
    const int isz = (b.sg > 1 ? b.sg : 1) * rec.src_iz, isx = (b.sg > 1 ? b.sg : 1) * rec.src_ix;

    AliBandGrid g;
    g.nz = b.nz; g.nx = b.nx;
    g.T = b.Tt + (size_t)src * b.tn;
    g.st = b.st + (size_t)src * b.tn;
    g.t4x = (b.nx + 3) >> 2;
    g.dirty = nullptr;
    g.tiles_x = 0;
    g.dnx = b.m.dnx;
    g.mv = ali_band_view(b.sg);

    const int cap = b.band_cap;
    double *val0 = b.stage + (size_t)src * 2 * cap, *val1 = val0 + cap;
    unsigned *ent0 = b.lists + (size_t)src * 4 * cap, *ent1 = ent0 + cap, *wrk0 = ent1 + cap, *wrk1 = wrk0 + cap;

    for (int q = tid; q < ALI_MT_WORDS; q += NT) s_sincos[q] = q < ALI_GL_SINCOSTAB_COUNT ? ali_gl_sincostab[q] : ali_gl_atan_cij[q - ALI_GL_SINCOSTAB_COUNT];
    if (tid == 0) {
        s_grid = g;
        s_evals = 0; s_fbs = 0;
        s_cyc[0] = s_cyc[1] = s_cyc[2] = s_cyc[3] = 0;
    }
    if (rec.overflow) return;   // (the sequential phase's verdict: the same for every CTA of the cluster)
    if (gtid == 0) {            // the control block arrives zeroed (host memset)
        ctl->evalmin[0] = ~0ull; ctl->evalmin[1] = ~0ull; ctl->basemin[0] = ~0ull; ctl->basemin[1] = ~0ull;
    }
    ali_cluster_sync();

    // hand-over of the sequential phase's window (see ali_march_kernel), shared by the CTAs
    {
        const AliSeqResult w = rec.seq;
        const int nlev = b.sg > 1 ? 2 : 3;
        const size_t woff = (size_t)src * 2 * b.seq_cap + ((((nlev - 1) & 1) == 0) ? b.seq_cap : 0);
        const int32_t *wst = b.seq_s + woff;
        const double *wt = b.seq_t + woff;
        const int wn = w.wnz * w.wnx;
        for (int base = rank * NT; base < wn; base += GT) {
            int i = base + tid;
            int k = 0;
            unsigned entry = 0;
            double tv = 0.0;
            if (i < wn) {
                int z = i / w.wnx, x = i - z * w.wnx;
                int32_t s = wst[i];
                if (s >= 0) {
                    tv = wt[i];
                    const size_t node = g.ti(w.wz0 + z, w.wx0 + x);
                    g.T[node] = tv;
                    if (s == 0) g.st[node] = ALI_ST_ALIVE;
                    else { k = 1; entry = ALI_PACK(w.wz0 + z, w.wx0 + x); }
                }
            }
            int pos = ali_warp_reserve(k, &ctl->count[0]);
            if (k) {
                if (pos < cap) { ent0[pos] = entry; wrk0[pos] = (unsigned)pos; val0[pos] = tv; }
                else ctl->overflow = 2;
            }
        }
    }
    ali_cluster_sync();
    if (gtid == 0) ctl->nwork[0] = ali_ldv(&ctl->count[0]);
    ali_cluster_sync();

    int rounds = 0, max_band = 0;
    unsigned my_evals = 0, my_fbs = 0;
    int cur = 0;
    while (true) {
        // (read after a cluster barrier, not written before the next one: the same for every thread of the cluster)
        const int n = ali_ldv(&ctl->count[cur]);
        if (n == 0 || ali_ldv(&ctl->overflow)) break;
        const int nwork = ali_ldv(&ctl->nwork[cur]);
        double *val = cur == 0 ? val0 : val1, *nval = cur == 0 ? val1 : val0;
        unsigned *ent = cur == 0 ? ent0 : ent1, *nent = cur == 0 ? ent1 : ent0;
        unsigned *wrk = cur == 0 ? wrk0 : wrk1, *nwrk = cur == 0 ? wrk1 : wrk0;
        rounds++;
        if (n > max_band) max_band = n;
        if (tid == 0) s_tprev = clock64();
        const bool resort = b.resort_every > 0 && (rounds % b.resort_every) == 0;
        // phase A
        if (gtid == 0) { ctl->count[cur ^ 1] = 0; ctl->nwork[cur ^ 1] = 0; ctl->basemin[cur ^ 1] = ~0ull; ctl->evalmin[cur ^ 1] = ~0ull; }
        for (int q = gtid; q < ALI_DMAP_WORDS / 4; q += GT) reinterpret_cast<uint4 *>(ctl->dmap)[q] = make_uint4(0u, 0u, 0u, 0u);
        if (resort)
            for (int q = gtid; q < ALI_SORT_BINS; q += GT) ctl->bins[q] = 0;
        double lmin = 1e300;
        unsigned pe0 = 0, pe1 = 0;
        double pv0 = 0.0, pv1 = 0.0;
        int pmask = 0, it = 0;
        // (warp w of CTA r takes the items of warp slot w * C + r: every CTA of the cluster gets the same
        // number of warps' worth of work, also when the list is shorter than the cluster has threads)
        const int q0 = (((tid >> 5) * C + rank) << 5) | (tid & 31);
        for (int q = q0; q < nwork; q += GT, it++) {
            const int i = (int)wrk[q];
            const unsigned e = ent[i];
            const int iz = ALI_PACK_Z(e), ix = ALI_PACK_X(e);
            int fb = 0;
            const double vold = val[i];
            double v = ali_band_eval(b.m, b.m_dev, g, &s_grid, iz, ix, &fb, s_sincos);
            if (v != v) v = __longlong_as_double(ALI_T_NAN_VALUE_BITS);
            val[i] = v;
            my_evals++;
            my_fbs += fb;
            lmin = fmin(lmin, v);
            if (it == 0) { pe0 = e; pv0 = v; pmask |= (v != vold ? 1 : 0) | (fb ? 4 : 0); }
            else if (it == 1) { pe1 = e; pv1 = v; pmask |= (v != vold ? 2 : 0) | (fb ? 8 : 0); }
            else {
                if (v != vold) val[i] = -v;
                if (fb) ctl->force[rounds & 1] = 1;
            }
        }
        for (int o = 16; o > 0; o >>= 1) lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        if ((tid & 31) == 0 && lmin < 1e300)
            atomicMin(&ctl->evalmin[cur], (unsigned long long)__double_as_longlong(lmin));
        ali_cluster_sync();
        if (tid == 0) { const long long now = clock64(); s_cyc[0] += now - s_tprev; s_tprev = now; }
        // phase B
        if (gtid == 0) ctl->force[(rounds + 1) & 1] = 0;
        if (pmask & 1) { g.T[g.ti(ALI_PACK_Z(pe0), ALI_PACK_X(pe0))] = pv0; ali_dmap_mark(ctl->dmap, ALI_PACK_Z(pe0), ALI_PACK_X(pe0)); }
        if (pmask & 2) { g.T[g.ti(ALI_PACK_Z(pe1), ALI_PACK_X(pe1))] = pv1; ali_dmap_mark(ctl->dmap, ALI_PACK_Z(pe1), ALI_PACK_X(pe1)); }
        if (pmask & 4) ali_dmap_row(ctl->dmap, ALI_PACK_Z(pe0), ALI_PACK_X(pe0), 1u);
        if (pmask & 8) ali_dmap_row(ctl->dmap, ALI_PACK_Z(pe1), ALI_PACK_X(pe1), 1u);
        for (int q = q0 + 2 * GT; q < nwork; q += GT) {
            const int i = (int)wrk[q];
            const unsigned e = ent[i];
            const double v = val[i];
            if (v < 0.0) {
                val[i] = -v;
                g.T[g.ti(ALI_PACK_Z(e), ALI_PACK_X(e))] = -v;
                ali_dmap_mark(ctl->dmap, ALI_PACK_Z(e), ALI_PACK_X(e));
            }
        }
        ali_cluster_sync();
        if (tid == 0) { const long long now = clock64(); s_cyc[1] += now - s_tprev; s_tprev = now; }
        // phase C
        const int force = ali_ldv(&ctl->force[rounds & 1]);
        const unsigned long long em = ali_ldv(&ctl->evalmin[cur]), bmn = ali_ldv(&ctl->basemin[cur]);
        const double thr = __longlong_as_double((long long)(em < bmn ? em : bmn)) + b.delta;
        double bmin = 1e300;
        for (int i = q0; i - (tid & 31) < n; i += GT) {   // (whole warps stay in the loop: the reservation shuffles)
            int k = 0, kw = 0;
            unsigned out[4];
            double v = 0.0;
            if (i < n) {
                const unsigned e = ent[i];
                v = val[i];
                const int iz = ALI_PACK_Z(e), ix = ALI_PACK_X(e);
                if (!(v > thr)) {
                    k = ali_band_accept(g, iz, ix, out);
                    kw = k;
                    v = 0.0;
                } else {
                    out[0] = e; k = 1;
                    if (!resort) {
                        if (force || ali_dmap_test_g(ctl->dmap, iz, ix)) kw = 1;
                        else bmin = fmin(bmin, v);
                    }
                }
            }
            int pos, wpos;
            ali_warp_reserve2(k, kw, &ctl->count[cur ^ 1], &ctl->nwork[cur ^ 1], pos, wpos);
            if (pos + k <= cap) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (q < k) {
                        nent[pos + q] = out[q];
                        nval[pos + q] = v;
                        if (q < kw) nwrk[wpos + q] = (unsigned)(pos + q);
                    }
            } else if (k) {
                ctl->overflow = 2;
            }
        }
        for (int o = 16; o > 0; o >>= 1) bmin = fmin(bmin, __shfl_xor_sync(0xffffffffu, bmin, o));
        if ((tid & 31) == 0 && bmin < 1e300)
            atomicMin(&ctl->basemin[cur ^ 1], (unsigned long long)__double_as_longlong(bmin));
        ali_cluster_sync();
        if (tid == 0) { const long long now = clock64(); s_cyc[2] += now - s_tprev; s_tprev = now; }
        if (resort && !ali_ldv(&ctl->overflow)) {
            // counting sort of the new list by angular bin through the shared bins: histogram (every
            // CTA), exclusive scan (CTA 0), scatter back into the current buffers, work list + base minimum
            const int nn = ali_ldv(&ctl->count[cur ^ 1]);
            for (int i = gtid; i < nn; i += GT) atomicAdd(&ctl->bins[ali_sort_bin(nent[i], isz, isx)], 1);
            ali_cluster_sync();
            if (rank == 0) {
                constexpr int PER = (ALI_SORT_BINS + NT - 1) / NT;
                int loc[PER];
                int sum = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    int idx = tid * PER + q;
                    loc[q] = idx < ALI_SORT_BINS ? ali_ldv(&ctl->bins[idx]) : 0;
                    sum += loc[q];
                }
                int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((tid & 31) >= o) incl += v;
                }
                if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
                __syncthreads();
                if (tid < 32) {
                    int w = tid < NT / 32 ? s_wsum[tid] : 0;
                    int wi = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        int v = __shfl_up_sync(0xffffffffu, wi, o);
                        if (tid >= o) wi += v;
                    }
                    s_wsum[tid] = wi - w;
                }
                __syncthreads();
                int run = s_wsum[tid >> 5] + incl - sum;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    int idx = tid * PER + q;
                    if (idx < ALI_SORT_BINS) ctl->bins[idx] = run;
                    run += loc[q];
                }
                if (tid == 0) { ctl->nwork[cur] = 0; ctl->basemin[cur] = ~0ull; }
            }
            ali_cluster_sync();
            for (int i = gtid; i < nn; i += GT) {
                const unsigned e = nent[i];
                const int pos = atomicAdd(&ctl->bins[ali_sort_bin(e, isz, isx)], 1);
                ent[pos] = e;
                val[pos] = nval[i];
            }
            ali_cluster_sync();
            double bm = 1e300;
            for (int i = q0; i - (tid & 31) < nn; i += GT) {
                int kw = 0;
                if (i < nn) {
                    const unsigned e = ent[i];
                    const double v = val[i];
                    if (v == 0.0) kw = 1;
                    else {
                        if (force || ali_dmap_test_g(ctl->dmap, ALI_PACK_Z(e), ALI_PACK_X(e))) kw = 1;
                        else bm = fmin(bm, v);
                    }
                }
                int wpos = ali_warp_reserve(kw, &ctl->nwork[cur]);
                if (kw) wrk[wpos] = (unsigned)i;
            }
            for (int o = 16; o > 0; o >>= 1) bm = fmin(bm, __shfl_xor_sync(0xffffffffu, bm, o));
            if ((tid & 31) == 0 && bm < 1e300)
                atomicMin(&ctl->basemin[cur], (unsigned long long)__double_as_longlong(bm));
            if (gtid == 0) { ctl->count[cur] = nn; ctl->evalmin[cur] = ~0ull; }
            ali_cluster_sync();
        } else {
            cur ^= 1;
        }
        if (tid == 0) {
            const long long now = clock64(); s_cyc[3] += now - s_tprev;
        }
    }
    atomicAdd(&s_evals, (unsigned long long)my_evals);
    atomicAdd(&s_fbs, (unsigned long long)my_fbs);
    __syncthreads();
    if (tid == 0) {
        atomicAdd((unsigned long long *)&rec.band_evals, s_evals);
        atomicAdd((unsigned long long *)&rec.band_fallbacks, s_fbs);
        if (rank == 0) {
            rec.rounds = rounds;
            rec.max_band = max_band;
            for (int q = 0; q < 4; q++) rec.cycles[q] = s_cyc[q];
            const int ovf = ali_ldv(&ctl->overflow);
            if (ovf) rec.overflow = ovf;
        }
    }
}

// Tiled march field -> the caller's row-major field, T / subgrid (ATR:2832); nodes the march never
// reached get the reference's 0.  A warp reads 8 sectors of 8 tiles and writes 256 contiguous bytes;
// the other three rows of those tiles are read by the next rows' warps out of L2.
__global__ void ali_finalize_kernel(const double *Tt, const uint8_t *st, double *T, int n_src, int nz, int nx, size_t tn, int sg)
{
    const size_t N = (size_t)nz * nx, total = (size_t)n_src * N;
    const size_t t4x = (size_t)((nx + 3) >> 2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t src = i / N, r = i - src * N;
        const int z = (int)(r / nx), x = (int)(r - (size_t)z * nx);
        const size_t node = src * tn + ((((size_t)(z >> 2) * t4x + (size_t)(x >> 2)) << 4) | (size_t)(((z & 3) << 2) | (x & 3)));
        const double v = Tt[node];
        // an accepted node without a number (NaN material / velocity) stays NaN, as in the reference
        T[i] = (v >= 0.0) ? v / sg : ((v != v && st[node] == ALI_ST_ALIVE) ? __longlong_as_double(ALI_T_NAN_VALUE_BITS) : 0.0);
    }
}

struct AliRayJob {
    int src_iz, src_ix; // coarse node the ray starts from
    int rec_slot;       // resident field it is traced through
};

struct AliRayArgs {
    AliModel m;
    int sg, fz, fx;
    const double *T;            // resident fields
    const AliSourceRec *rec;    // their sources
    const AliRayJob *jobs;
    const int *order;           // ray indices, longest expected path first
    int *next;                  // work queue: next position in `order`
    int n_rays, cap;
    double *out_x, *out_y, *out_time;
    int *out_len, *out_flag;
    int maxc;
};

#define ALI_RAY_WARPS 4
template <int MINB>
__global__ void __launch_bounds__(32 * ALI_RAY_WARPS, MINB) ali_rays_kernel(AliRayArgs a)
{
    extern __shared__ double s_buf[];
    __shared__ AliRayState s_state[ALI_RAY_WARPS];
    __shared__ AliRayPlane s_plane[ALI_RAY_WARPS];
    __shared__ int s_go[ALI_RAY_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The warps take rays from a queue ordered longest-first (paths differ 2x in length: a fixed ray -> warp
    // assignment left the SMs half empty while the longest rays of the last wave finished).
    for (;;) {
        int ray = 0;
        if (lane == 0) {
            const int q = atomicAdd(a.next, 1);
            ray = q < a.n_rays ? a.order[q] : -1;
        }
        ray = __shfl_sync(0xffffffffu, ray, 0);
        if (ray < 0) break;
        double *TT = s_buf + (size_t)warp * 3 * a.maxc;
        double *vals = TT + a.maxc;
        double *poss = vals + a.maxc;
        AliRayState &s = s_state[warp];
        AliRayPlane &pl = s_plane[warp];
        const AliRayJob job = a.jobs[ray];
        const double *rec = a.T + (size_t)job.rec_slot * a.fz * a.fx;
        double *ray_x = a.out_x + (size_t)ray * a.cap;
        double *ray_y = a.out_y + (size_t)ray * a.cap;
        if (lane == 0) {
            const AliSourceRec &r = a.rec[job.rec_slot];
            s.last_x = (double)(a.sg * job.src_ix); s.last_y = (double)(a.sg * job.src_iz);
            s.rx = (double)(a.sg * r.src_ix); s.ry = (double)(a.sg * r.src_iz);
            s.lvx = s.rx - s.last_x; s.lvy = s.ry - s.last_y;
            s.len = 1; s.flag = 0; s.done = 0;
            ray_x[0] = s.last_x; ray_y[0] = s.last_y;
        }
        __syncwarp();
        while (true) {
            if (lane == 0) {
                int go = 0;
                if (ali_ray_continue(s, a.sg)) {
                    if (s.len >= a.cap - 1) s.flag |= ALI_RAY_CAPACITY;
                    else go = ali_ray_choose_plane(s, a.sg, a.fz, a.fx, pl) ? 1 : 0;
                }
                s_go[warp] = go;
            }
            __syncwarp();
            if (!s_go[warp]) break;
            const double lx = s.last_x, ly = s.last_y;
            for (int i = lane; i < pl.len; i += 32) TT[i] = ali_ray_candidate_time(a.m, rec, a.fx, pl, i, lx, ly, a.sg);
            __syncwarp();
            for (int j = 1 + lane; j < pl.len - 1; j += 32) vals[j] = ali_ray_local_min(TT, j, poss[j]);
            __syncwarp();
            if (lane == 0) {
                double min_i = ali_ray_select(TT, vals, poss, pl.len);
                s_go[warp] = ali_ray_advance(s, pl, min_i, rec, a.fx, ray_x, ray_y) ? 1 : 0;
            }
            __syncwarp();
            if (!s_go[warp]) break;
        }
        if (lane == 0) {
            ray_x[s.len] = s.rx; ray_y[s.len] = s.ry;
            s.len += 1;
        }
        __syncwarp();
        // ray_time (ATR:2992-3022): segment times in parallel, summed in path order
        const int nseg = s.len - 1;
        double acc = 0.0;
        for (int base = 0; base < nseg; base += 32) {
            int i = base + lane;
            double v = 0.0;
            if (i < nseg) v = ali_time_between_points(a.m, ray_x[i], ray_x[i + 1], ray_y[i], ray_y[i + 1], a.sg, 1 << 20);
            TT[lane] = v;
            __syncwarp();
            if (lane == 0) {
                int cnt = nseg - base < 32 ? nseg - base : 32;
                for (int q = 0; q < cnt; q++) acc += TT[q];
            }
            __syncwarp();
        }
        if (lane == 0) {
            a.out_len[ray] = s.len;
            a.out_time[ray] = acc;
            a.out_flag[ray] = s.flag;
        }
        __syncwarp();
    }
}

// generate_group_vel / generate_phase_vel (ATR:4112-4206) at one whole degree a in 0..360 (stiffness in
// Pa, no factor 1000; exact special cases at multiples of 90; 180..360 mirror 0..180, ATR:4154, 4200).
__device__ __forceinline__ void ali_curve_point(int a, double c22, double c23, double c33, double c44, double rho,
                                                double &gv, double &pv)
{
    int angle = a < 180 ? a : a - 180;
    if (angle >= 180) angle -= 180;   // 360 -> 180 -> index 0
    if (angle % 90 == 0) {
        const double lam = (angle % 180 == 90) ? c33 : c22;
        gv = sqrt(lam / rho);
        pv = gv;
        return;
    }
    const double rad = ALI_DEG2RAD * angle;
    const double t = ALI_TAN(rad);
    const double A = c22 + c33 - 2 * c44;
    const double B = (c23 + c44) * (t - 1 / t);
    const double C = c22 - c33;
    const double disc = sqrt(B * B + A * A - C * C);
    double ph;
    if (angle < 90) ph = ali_pymod(ALI_ATAN((-B - disc) / (C - A)), ALI_PI);
    else ph = ali_pymod(ALI_ATAN((-B + disc) / (C - A)), ALI_PI);
    const double lam = 0.5 * (ALI_COS(2 * ph) * (c22 - c44) + ALI_SIN(2 * ph) * (c23 + c44) * t + c22 + c44);
    gv = sqrt(lam / rho) / ALI_COS(rad - ph);
    const double cs = ALI_COS(rad), sn = ALI_SIN(rad);
    const double A2 = cs * cs * c22 + sn * sn * c44;
    const double B2 = cs * sn * (c23 + c44);
    const double C2 = cs * cs * c44 + sn * sn * c33;
    pv = sqrt((A2 + C2 + sqrt((A2 - C2) * (A2 - C2) + 4 * (B2 * B2))) / (2 * rho));
}

__global__ void ali_curves_kernel(double c22, double c23, double c33, double c44, double rho, double *group,
                                  double *phase)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a > 360) return;
    ali_curve_point(a, c22, c23, c33, c44, rho, group[a], phase[a]);
}

// The same curves for a batch of materials (add_materials, ATR:4208-4256): props[n_mat][5] in Pa,
// outputs [n_mat][361].  One thread per (material, degree).
__global__ void ali_curves_batch_kernel(int n_mat, const double *props, double *group, double *phase)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_mat * 361) return;
    const int mat = t / 361, a = t - mat * 361;
    const double c22 = props[5 * mat], c23 = props[5 * mat + 1], c33 = props[5 * mat + 2], c44 = props[5 * mat + 3],
                 rho = props[5 * mat + 4];
    ali_curve_point(a, c22, c23, c33, c44, rho, group[t], phase[t]);
}

// Node-level operator check (test entry alifmm_eval_nodes): one thread per caller-supplied state
// runs the ALI update (ATR:904-1410) and the FD fallback (ATR:240-901) exactly as the kernels do.
struct AliNodeState {
    int nz, nx;
    const double *t;
    const int32_t *st;
    ALI_DEV bool in(int z, int x) const { return z >= 0 && z < nz && x >= 0 && x < nx; }
    ALI_DEV bool avail(int z, int x) const { return in(z, x) && st[z * nx + x] >= 0; }
    ALI_DEV bool alive(int z, int x) const { return in(z, x) && st[z * nx + x] == 0; }
    ALI_DEV double tt(int z, int x) const { return t[z * nx + x]; }
};

__global__ void ali_eval_nodes_kernel(int n, AliModel m0, const double *ttn, const int32_t *nsts, const int32_t *pos,
                                      double *out_update, double *out_fouds, int32_t *out_stencil)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int nn = m0.nz * m0.nx;
    AliModel m = m0;
    m.rec = m0.rec + (size_t)c * nn;
    AliNodeState st;
    st.nz = m.nz; st.nx = m.nx; st.t = ttn + (size_t)c * nn; st.st = nsts + (size_t)c * nn;
    const int iz = pos[2 * c], ix = pos[2 * c + 1];
    const AliMatView idv = ali_view_identity();
    AliMat mat;
    AliWindow w;
    ali_fetch_mat(m, idv, iz, ix, mat);
    ali_gather(st, iz, ix, m.nz, m.nx, w);
    int stencil = 0;
    out_update[c] = ali_update_window(m, mat, w, iz, ix, m.nz, m.nx, m.dnx, &stencil);
    if (out_stencil) out_stencil[c] = stencil;
    out_fouds[c] = ali_fouds18(m, mat, st, iz, ix, m.dnx, m.dnx, m.nx, m.nz);
}

// min_max_vel (ATR:3736-3787): group velocity at 0/45/90/135 degrees per node (Christoffel
// models, decided by velpn[0,0] as in the reference) or table column extrema.
__global__ void ali_minmax_kernel(AliModel m, int first_velpn, const double *col_min, const double *col_max,
                                  unsigned long long *min_bits, unsigned long long *max_bits)
{
    const size_t n = (size_t)m.nz * m.nx;
    double lo = 1e300, hi = 0.0;
    const AliMatView idv = ali_view_identity();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int iz = (int)(i / m.nx), ix = (int)(i % m.nx);
        AliMat mat;
        ali_fetch_mat(m, idv, iz, ix, mat);
        if (first_velpn == 0) {
            for (int q = 0; q < 4; q++) {
                double v = ali_christoffel_group(45.0 * q, mat.s, mat.vel_map);
                lo = fmin(lo, v); hi = fmax(hi, v);
            }
        } else {
            lo = fmin(lo, mat.vel_map * col_min[mat.velpn]);
            hi = fmax(hi, mat.vel_map * col_max[mat.velpn]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(min_bits, (unsigned long long)__double_as_longlong(lo));
        atomicMax(max_bits, (unsigned long long)__double_as_longlong(hi));
    }
}

// ---------------------------------------------------------------------------
// host side: context
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct alifmm_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    AliModel m{};           // device pointers
    const AliModel *m_dev = nullptr;
    int first_velpn = 0;
    int nz = 0, nx = 0;
    std::vector<DevBuf> model_allocs;   // pooled like the batch buffers (cudaFree stalls for up to 0.8 s now and then)
    double vmax = 0.0;
    // options
    double delta_frac = 0.35;   // rounds ~ 1 / delta_frac.  0.1 ... 0.35 give the same bits on every test model (host replay + B200); 0.4: 5e-12; 0.45: 1e-7; 0.5: 1e-4
    int margin = 27;
    double band_cap_factor = 6.0;
    int threads_per_source = 768;   // 80 registers per thread: fewer spills than 1024 x 64, more warps than 512 x 128 (measured)
    int resort_every = 8;
    int ray_min_blocks = 4;     // CTAs of 4 warps per SM the ray kernel is compiled for (4: 126 registers, 5: 102, 6: 85)
    int band_smem_bytes = 0;   // measured on B200: L1 for the T / status gathers is worth more than smem lists
    int seq_threads = 32;      // CTA size of the sequential near-source kernel: 32 = one warp (measured fastest); 64..256 = one candidate per warp
    int cluster_size = 0;      // CTAs (SMs) per source in the band march: 0 = as many (1, 2, 4, 8) as keep the batch in one wave
    int cluster_threads = 0;   // CTA size of the cluster march: 0 = 512 (128 registers, no spills), else 512 or 768
    // resident batch
    int n_slots = 0, sg = 0, fz = 0, fx = 0;
    std::vector<int> slot_iz, slot_ix;   // source node of every resident field
    DevBuf T, Tt, st, seq_t, seq_s, seq_heap, seq_hkey, seq_cval, seq_cflag, lists, stage, rec, jobs, ray_q, ray_x, ray_y, ray_time, ray_len, ray_flag, ray_off, pack, misc, ctl;
    alifmm_counters_t cnt{};
};

// Large device buffers are recycled through a per-device pool: the reference-facing API creates a
// context per call (it re-takes the model every time, ATR:3889-3900), and cudaFree / cudaMalloc of
// the tens of GB a headline batch needs cost 0.7 s per call otherwise.  alifmm_trim() empties it.
#define ALI_POOL_MIN_BYTES ((size_t)4096)
struct PoolEntry { void *p; size_t bytes; int device; };
static std::mutex g_pool_mutex;
static std::vector<PoolEntry> g_pool;

static size_t pool_bytes(int device)
{
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    size_t n = 0;
    for (const PoolEntry &e : g_pool) if (e.device == device) n += e.bytes;
    return n;
}

static void pool_trim(int device)   // device < 0: all devices
{
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    int cur = -1;
    cudaGetDevice(&cur);
    for (size_t i = 0; i < g_pool.size();) {
        if (device < 0 || g_pool[i].device == device) {
            cudaSetDevice(g_pool[i].device);
            cudaFree(g_pool[i].p);
            g_pool[i] = g_pool.back();
            g_pool.pop_back();
        } else {
            i++;
        }
    }
    if (cur >= 0) cudaSetDevice(cur);
}

static void dev_give_back(DevBuf &b, int device)
{
    if (!b.p) return;
    if (b.bytes >= ALI_POOL_MIN_BYTES) {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        g_pool.push_back(PoolEntry{b.p, b.bytes, device});
    } else {
        cudaFree(b.p);
    }
    b.p = nullptr; b.bytes = 0;
}

// The calling thread's current device must be the context's.
static int dev_reserve(DevBuf &b, size_t bytes)
{
    if (b.bytes >= bytes && b.p) return ALIFMM_OK;
    int device = 0;
    cudaGetDevice(&device);
    dev_give_back(b, device);
    if (bytes >= ALI_POOL_MIN_BYTES) {
        // smallest pooled buffer that fits and is not wastefully large
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        int best = -1;
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i].device == device && g_pool[i].bytes >= bytes && g_pool[i].bytes <= bytes + bytes / 4 + (1 << 20) &&
                (best < 0 || g_pool[i].bytes < g_pool[best].bytes))
                best = (int)i;
        if (best >= 0) {
            b.p = g_pool[best].p; b.bytes = g_pool[best].bytes;
            g_pool[best] = g_pool.back();
            g_pool.pop_back();
            return ALIFMM_OK;
        }
    }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        // the pool may be holding what this allocation needs
        cudaGetLastError();
        pool_trim(device);
        e = cudaMalloc(&b.p, bytes);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(ALIFMM_E_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    b.bytes = bytes;
    return ALIFMM_OK;
}

static void dev_release(DevBuf &b, int device) { dev_give_back(b, device); }

// Pinned staging buffers are recycled like the device buffers (cudaHostAlloc costs tens of ms).
static std::vector<PoolEntry> g_pin_pool;
static void *pin_take(size_t bytes, size_t *got)
{
    {
        // best fit, and never a buffer more than twice the request (a big staging buffer handed out for a
        // 64 MB chunk would force another big cudaHostAlloc later)
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        int best = -1;
        for (size_t i = 0; i < g_pin_pool.size(); i++)
            if (g_pin_pool[i].bytes >= bytes && g_pin_pool[i].bytes <= 2 * bytes + (1 << 20) &&
                (best < 0 || g_pin_pool[i].bytes < g_pin_pool[best].bytes))
                best = (int)i;
        if (best >= 0) {
            void *p = g_pin_pool[best].p;
            *got = g_pin_pool[best].bytes;
            g_pin_pool[best] = g_pin_pool.back();
            g_pin_pool.pop_back();
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got = bytes;
    return p;
}
#define ALI_PIN_POOL_MAX_BYTES ((size_t)1 << 30)   // page-locked memory kept for reuse, at most
static void pin_give(void *p, size_t bytes)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mutex);
        size_t held = 0;
        for (const PoolEntry &e : g_pin_pool) held += e.bytes;
        if (held + bytes <= ALI_PIN_POOL_MAX_BYTES) {
            g_pin_pool.push_back(PoolEntry{p, bytes, -1});
            return;
        }
    }
    cudaFreeHost(p);
}

// Field results -> caller's (pageable) memory.  A plain cudaMemcpy into pageable memory runs at
// ~5 GB/s (the driver stages it, and a fresh destination takes a page fault per 4 KB); here the
// copy is pipelined through two pinned buffers and the host side of it is shared by a few threads.
static int copy_to_host(alifmm_ctx *c, void *dst, const void *src_dev, size_t bytes)
{
    cudaStream_t s = c->stream;
    const size_t CH = (size_t)64 << 20;
    size_t g0 = 0, g1 = 0;
    char *pin[2] = {nullptr, nullptr};
    if (bytes >= ((size_t)16 << 20)) {
        pin[0] = (char *)pin_take(CH, &g0);
        pin[1] = pin[0] ? (char *)pin_take(CH, &g1) : nullptr;
    }
    if (!pin[0] || !pin[1]) {
        if (pin[0]) pin_give(pin[0], g0);
        CUDA_TRY(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        return ALIFMM_OK;
    }
    cudaEvent_t ev[2];
    cudaError_t e = cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    const size_t nch = (bytes + CH - 1) / CH;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * CH, n = bytes - off < CH ? bytes - off : CH;
        cudaError_t r = cudaMemcpyAsync(pin[i & 1], (const char *)src_dev + off, n, cudaMemcpyDeviceToHost, s);
        if (r == cudaSuccess) r = cudaEventRecord(ev[i & 1], s);
        return r;
    };
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
    if (e == cudaSuccess) e = issue(0);
    for (size_t i = 0; i < nch && e == cudaSuccess; i++) {
        if (i + 1 < nch) e = issue(i + 1);          // the other buffer was consumed in the previous iteration
        if (e == cudaSuccess) e = cudaEventSynchronize(ev[i & 1]);
        if (e != cudaSuccess) break;
        const size_t off = i * CH, n = bytes - off < CH ? bytes - off : CH;
        const size_t part = ((n / nt) + 4095) & ~(size_t)4095;
        auto work = [&](unsigned t) {
            const size_t a = (size_t)t * part;
            if (a >= n) return;
            const size_t m = n - a < part ? n - a : part;
            memcpy((char *)dst + off + a, pin[i & 1] + a, m);
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nt; t++) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
    }
    if (e != cudaSuccess) cudaStreamSynchronize(s);
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    pin_give(pin[0], g0); pin_give(pin[1], g1);
    if (e != cudaSuccess) return fail(ALIFMM_E_CUDA, std::string("copy_to_host: ") + cudaGetErrorString(e));
    return ALIFMM_OK;
}

template <class Tp>
static int upload(alifmm_ctx *c, const Tp *host, size_t n, const Tp **dev_out)
{
    DevBuf buf;
    int rc = dev_reserve(buf, n * sizeof(Tp));
    if (rc != ALIFMM_OK) return rc;
    c->model_allocs.push_back(buf);
    CUDA_TRY(cudaMemcpyAsync(buf.p, host, n * sizeof(Tp), cudaMemcpyHostToDevice, c->stream));
    *dev_out = (const Tp *)buf.p;
    return ALIFMM_OK;
}

extern "C" const char *alifmm_last_error(void) { return g_last_error.c_str(); }

extern "C" int alifmm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void alifmm_destroy(alifmm_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (DevBuf &mb : c->model_allocs) dev_release(mb, c->device);
    DevBuf *bufs[] = {&c->T, &c->Tt, &c->st, &c->seq_t, &c->seq_s, &c->seq_heap, &c->seq_hkey, &c->seq_cval, &c->seq_cflag, &c->lists, &c->stage, &c->rec, &c->jobs, &c->ray_q,
                      &c->ray_x, &c->ray_y, &c->ray_time, &c->ray_len, &c->ray_flag, &c->ray_off, &c->pack, &c->misc, &c->ctl};
    for (DevBuf *b : bufs) dev_release(*b, c->device);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int alifmm_create(const alifmm_model_desc *d, int device, alifmm_ctx **out)
{
    if (!d || !out) return fail(ALIFMM_E_INVALID, "alifmm_create: null argument");
    if (d->nz < 1 || d->nx < 1 || !(d->dnx > 0) || !d->veln || !d->velpn || !d->vel_map || !d->group_vel ||
        !d->phase_vel || d->n_cols < 1)
        return fail(ALIFMM_E_INVALID, "alifmm_create: bad model descriptor");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ALIFMM_E_CUDA, "alifmm_create: no CUDA device available (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(ALIFMM_E_INVALID, "alifmm_create: device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    alifmm_ctx *c = new alifmm_ctx();
    c->device = device;
    int rc = ALIFMM_OK;
    auto bail = [&](int code) { alifmm_destroy(c); return code; };
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(ALIFMM_E_CUDA, "cudaStreamCreate failed"));
    c->stream = c->own_stream;
    for (auto &ev : c->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) return bail(fail(ALIFMM_E_CUDA, "cudaEventCreate failed"));
    const size_t n = (size_t)d->nz * d->nx;
    c->nz = d->nz; c->nx = d->nx;
    c->m.nz = d->nz; c->m.nx = d->nx; c->m.dnx = d->dnx; c->m.ncol = d->n_cols;
    c->m.has_stif = d->has_stif ? 1 : 0;
    // material ids must index the tables (an out-of-range id would read outside them on the device)
    for (size_t i = 0; i < n; i++)
        if (d->velpn[i] < 0 || d->velpn[i] >= d->n_cols)
            return bail(fail(ALIFMM_E_INVALID, "alifmm_create: velpn holds a material id outside the velocity tables"));
    c->first_velpn = d->velpn[0];
    {   // raw arrays -> 64-byte records (the raw device copies are released afterwards)
        const double *dv = nullptr, *dm = nullptr;
        const int32_t *dp = nullptr;
        const long long *ds = nullptr;
        if ((rc = upload(c, d->veln, n, &dv)) != 0) return bail(rc);
        if ((rc = upload(c, d->velpn, n, &dp)) != 0) return bail(rc);
        if ((rc = upload(c, d->vel_map, n, &dm)) != 0) return bail(rc);
        if (d->stif_den && (rc = upload(c, (const long long *)d->stif_den, n * 5, &ds)) != 0) return bail(rc);
        DevBuf recb;
        if ((rc = dev_reserve(recb, n * sizeof(AliMatRec))) != 0) return bail(rc);
        void *recp = recb.p;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        if ((rc = dev_reserve(c->misc, 64)) != 0) { dev_release(recb, c->device); return bail(rc); }
        int bad = 0;
        if (cudaMemsetAsync(c->misc.p, 0, 64, c->stream) != cudaSuccess) { dev_release(recb, c->device); return bail(fail(ALIFMM_E_CUDA, "memset failed")); }
        ali_records_kernel<<<blocks, 256, 0, c->stream>>>((int)n, dv, dp, dm, ds, (AliMatRec *)recp, (int *)c->misc.p);
        if (cudaMemcpyAsync(&bad, c->misc.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaStreamSynchronize(c->stream) != cudaSuccess) {
            dev_release(recb, c->device);
            return bail(fail(ALIFMM_E_CUDA, std::string("alifmm_create: records kernel failed: ") + cudaGetErrorString(cudaGetLastError())));
        }
        if (bad) {
            // a NaN / infinite orientation or velocity scale gives the reference a NaN field; here it is an error
            dev_release(recb, c->device);
            return bail(fail(ALIFMM_E_INVALID, "alifmm_create: veln / vel_map hold non-finite values"));
        }
        for (DevBuf &q : c->model_allocs) dev_release(q, c->device);
        c->model_allocs.clear();
        c->model_allocs.push_back(recb);
        c->m.rec = (const AliMatRec *)recp;
    }
    if ((rc = upload(c, d->group_vel, (size_t)361 * d->n_cols, &c->m.group_tab)) != 0) return bail(rc);
    if ((rc = upload(c, d->phase_vel, (size_t)361 * d->n_cols, &c->m.phase_tab)) != 0) return bail(rc);
    {   // the model struct itself, in device memory, for out-of-line slow paths
        DevBuf mb;
        if ((rc = dev_reserve(mb, sizeof(AliModel) < 4096 ? 4096 : sizeof(AliModel))) != 0) return bail(rc);
        void *mp = mb.p;
        c->model_allocs.push_back(mb);
        if (cudaMemcpyAsync(mp, &c->m, sizeof(AliModel), cudaMemcpyHostToDevice, c->stream) != cudaSuccess)
            return bail(fail(ALIFMM_E_CUDA, "alifmm_create: model upload failed"));
        c->m_dev = (const AliModel *)mp;
    }
    // model-wide phase-velocity bound
    if ((rc = dev_reserve(c->misc, 64)) != 0) return bail(rc);
    if (cudaMemsetAsync(c->misc.p, 0, 64, c->stream) != cudaSuccess) return bail(fail(ALIFMM_E_CUDA, "memset failed"));
    {
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        ali_vmax_kernel<<<blocks, 256, 0, c->stream>>>(c->m, (unsigned long long *)c->misc.p);
    }
    unsigned long long bits = 0;
    if (cudaMemcpyAsync(&bits, c->misc.p, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) {
        cudaError_t le = cudaGetLastError();
        return bail(fail(ALIFMM_E_CUDA, std::string("alifmm_create: vmax kernel failed: ") + cudaGetErrorString(le)));
    }
    memcpy(&c->vmax, &bits, 8);
    if (!(c->vmax > 0) || !isfinite(c->vmax))
        return bail(fail(ALIFMM_E_INVALID, "alifmm_create: model has no positive finite phase velocity"));
    c->cnt.vmax = c->vmax;
    *out = c;
    return ALIFMM_OK;
}

extern "C" int alifmm_set_option(alifmm_ctx *c, const char *name, double value)
{
    if (!c || !name) return fail(ALIFMM_E_INVALID, "alifmm_set_option: null argument");
    if (!strcmp(name, "delta_frac")) {
        if (!(value > 0 && value <= 0.4)) return fail(ALIFMM_E_INVALID, "delta_frac must be in (0, 0.4]");
        c->delta_frac = value;
    } else if (!strcmp(name, "handover_margin")) {
        if (value < 4 || value > 100000) return fail(ALIFMM_E_INVALID, "handover_margin must be >= 4");
        c->margin = (int)value;
    } else if (!strcmp(name, "band_capacity_factor")) {
        if (!(value >= 1 && value <= 1024)) return fail(ALIFMM_E_INVALID, "band_capacity_factor must be in [1, 1024]");
        c->band_cap_factor = value;
    } else if (!strcmp(name, "threads_per_source")) {
        int t = (int)value;
        if (t != 256 && t != 512 && t != 640 && t != 768 && t != 896 && t != 1024)
            return fail(ALIFMM_E_INVALID, "threads_per_source must be 256, 512, 640, 768, 896 or 1024");
        c->threads_per_source = t;
    } else if (!strcmp(name, "ray_min_blocks")) {
        if (value != 4 && value != 5 && value != 6) return fail(ALIFMM_E_INVALID, "alifmm_set_option: ray_min_blocks is 4, 5 or 6");
        c->ray_min_blocks = (int)value;
    } else if (!strcmp(name, "resort_every")) {
        if (value < 0 || value > 1000000) return fail(ALIFMM_E_INVALID, "resort_every must be >= 0");
        c->resort_every = (int)value;
    } else if (!strcmp(name, "seq_threads")) {
        int t = (int)value;
        if (t != 32 && t != 64 && t != 128 && t != 256) return fail(ALIFMM_E_INVALID, "seq_threads must be 32, 64, 128 or 256");
        c->seq_threads = t;
    } else if (!strcmp(name, "cluster_size")) {
        int t = (int)value;
        if (t < 0 || t > 8) return fail(ALIFMM_E_INVALID, "cluster_size must be 0 (auto) or 1 ... 8");
        c->cluster_size = t;
    } else if (!strcmp(name, "cluster_threads")) {
        int t = (int)value;
        if (t != 0 && t != 512 && t != 768) return fail(ALIFMM_E_INVALID, "cluster_threads must be 0 (auto), 512 or 768");
        c->cluster_threads = t;
    } else if (!strcmp(name, "band_smem_kb")) {
        if (value < 0 || value > 180) return fail(ALIFMM_E_INVALID, "band_smem_kb must be in [0, 180]");
        c->band_smem_bytes = (int)value * 1024;
    } else {
        return fail(ALIFMM_E_INVALID, std::string("unknown option ") + name);
    }
    return ALIFMM_OK;
}

extern "C" int alifmm_set_stream(alifmm_ctx *c, void *s)
{
    if (!c) return fail(ALIFMM_E_INVALID, "alifmm_set_stream: null context");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return ALIFMM_OK;
}

static int ttf_attempt(alifmm_ctx *c, int32_t n_src, const int32_t *src_iz, const int32_t *src_ix, int32_t sg,
                       double *out_host, int band_cap_override, int *band_overflow);

extern "C" int alifmm_ttf(alifmm_ctx *c, int32_t n_src, const int32_t *src_iz, const int32_t *src_ix, int32_t sg,
                          double *out_host)
{
    // first try a narrow-band capacity that fits in shared memory; a band that outgrows it
    // (closed fronts of interior sources on large grids) is re-run from global buffers
    int overflow = 0;
    if (c && c->band_smem_bytes >= 16 * 2048) {
        int rc = ttf_attempt(c, n_src, src_iz, src_ix, sg, out_host, c->band_smem_bytes / 16, &overflow);
        if (rc != ALIFMM_E_CAPACITY || !overflow) return rc;
    }
    return ttf_attempt(c, n_src, src_iz, src_ix, sg, out_host, 0, &overflow);
}

static int ttf_attempt(alifmm_ctx *c, int32_t n_src, const int32_t *src_iz, const int32_t *src_ix, int32_t sg,
                       double *out_host, int band_cap_override, int *band_overflow)
{
    if (!c || !src_iz || !src_ix) return fail(ALIFMM_E_INVALID, "alifmm_ttf: null argument");
    if (n_src < 1) return fail(ALIFMM_E_INVALID, "alifmm_ttf: n_src must be >= 1");
    if (sg < 1 || (sg & 1) == 0) return fail(ALIFMM_E_INVALID, "alifmm_ttf: subgrid must be odd and >= 1");
    for (int k = 0; k < n_src; k++)
        if (src_iz[k] < 0 || src_iz[k] >= c->nz || src_ix[k] < 0 || src_ix[k] >= c->nx)
            return fail(ALIFMM_E_INVALID, "alifmm_ttf: source node outside the grid");
    CUDA_TRY(cudaSetDevice(c->device));
    const int fz = sg > 1 ? sg * (c->nz - 1) + 1 : c->nz;
    const int fx = sg > 1 ? sg * (c->nx - 1) + 1 : c->nx;
    if ((long long)fz * fx > 2147483000LL) return fail(ALIFMM_E_INVALID, "alifmm_ttf: grid has more than 2^31 nodes");
    const size_t N = (size_t)fz * fx;

    AliBatch b;
    if (fz > 65535 || fx > 65535) return fail(ALIFMM_E_INVALID, "alifmm_ttf: grid side exceeds 65535 nodes");
    b.m = c->m; b.m_dev = c->m_dev; b.sg = sg; b.nz = fz; b.nx = fx; b.margin = c->margin;
    b.delta = c->delta_frac * c->m.dnx / c->vmax;
    {   // scratch sizes from the plan (same for every source)
        AliSourcePlan p;
        ali_make_plan(p, c->m, 0, 0, sg, c->margin);
        size_t lvl = ali_plan_max_level_nodes(p);
        size_t w = (size_t)(2 * (p.stop_r + 4) + 1);
        size_t cap = lvl > w * w ? lvl : w * w;
        b.seq_cap = cap;
        b.heap_cap = (int)(cap / 2 + 64);
    }
    b.band_cap = (int)(c->band_cap_factor * (double)(fz + fx)) + 1024;
    if (band_cap_override > 0 && band_cap_override < b.band_cap) b.band_cap = band_cap_override;
    *band_overflow = 0;
    b.resort_every = c->resort_every;
    int rc;
    b.tn = ali_field_nodes_tiled(fz, fx);
    if ((rc = dev_reserve(c->T, (size_t)n_src * N * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->Tt, (size_t)n_src * b.tn * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->st, (size_t)n_src * b.tn + 16)) != 0) return rc;
    if ((rc = dev_reserve(c->seq_t, (size_t)n_src * 2 * b.seq_cap * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->seq_s, (size_t)n_src * 2 * b.seq_cap * sizeof(int32_t))) != 0) return rc;
    if ((rc = dev_reserve(c->seq_heap, (size_t)n_src * 2 * b.heap_cap * sizeof(int32_t))) != 0) return rc;
    if ((rc = dev_reserve(c->seq_hkey, (size_t)n_src * ALI_HKEY_SLOTS(b.heap_cap) * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->seq_cval, (size_t)n_src * b.seq_cap * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->seq_cflag, (size_t)n_src * b.seq_cap + 16)) != 0) return rc;
    if ((rc = dev_reserve(c->lists, (size_t)n_src * 4 * b.band_cap * sizeof(unsigned))) != 0) return rc;
    if ((rc = dev_reserve(c->stage, (size_t)n_src * 2 * b.band_cap * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->rec, (size_t)n_src * sizeof(AliSourceRec))) != 0) return rc;
    b.T = (double *)c->T.p; b.Tt = (double *)c->Tt.p; b.st = (uint8_t *)c->st.p;
    b.seq_t = (double *)c->seq_t.p; b.seq_s = (int32_t *)c->seq_s.p; b.seq_heap = (int32_t *)c->seq_heap.p;
    b.seq_hkey = (double *)c->seq_hkey.p; b.seq_cval = (double *)c->seq_cval.p; b.seq_cflag = (uint8_t *)c->seq_cflag.p;
    b.lists = (unsigned *)c->lists.p; b.stage = (double *)c->stage.p; b.rec = (AliSourceRec *)c->rec.p;

    std::vector<AliSourceRec> recs(n_src);
    memset(recs.data(), 0, recs.size() * sizeof(AliSourceRec));
    for (int k = 0; k < n_src; k++) { recs[k].src_iz = src_iz[k]; recs[k].src_ix = src_ix[k]; }
    c->n_slots = 0;
    c->slot_iz.assign(src_iz, src_iz + n_src); c->slot_ix.assign(src_ix, src_ix + n_src);
    cudaStream_t s = c->stream;
    CUDA_TRY(cudaMemcpyAsync(b.rec, recs.data(), recs.size() * sizeof(AliSourceRec), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(c->ev[0], s));
    CUDA_TRY(cudaMemsetAsync(b.Tt, ALI_T_UNSET_BYTE, (size_t)n_src * b.tn * sizeof(double), s)); // NaN = far
    CUDA_TRY(cudaMemsetAsync(b.st, 0, (size_t)n_src * b.tn, s));
    ali_seq_kernel<<<n_src, c->seq_threads, 0, s>>>(b);
    c->cnt.seq_threads = c->seq_threads;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev[1], s));
    // CTAs per source: a cluster when the batch leaves SMs idle (sources sharded over several GPUs)
    // threads per CTA of a cluster: 512 (128 registers, no spills) unless asked otherwise; a pair of CTAs does better
    // with 768 each (64 sources: 440 against 484 ms), decided below once the cluster size is known
    int csize = 1, cthreads = c->cluster_threads ? c->cluster_threads : 512;
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            if (c->cluster_threads != 0 || csize != 2) break;
            cthreads = 768;   // re-check residency with the larger CTAs
            csize = 1;
        }
        const size_t csmem = (size_t)ALI_MT_WORDS * 8;
        for (int cand : {8, 7, 6, 5, 4, 3, 2}) {
            if (c->cluster_size != 0 && c->cluster_size != cand) continue;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(n_src * cand)); cfg.blockDim = dim3((unsigned)cthreads); cfg.dynamicSmemBytes = csmem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)cand; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nclusters = 0;
            cudaError_t qe = cthreads == 768 ? cudaOccupancyMaxActiveClusters(&nclusters, ali_march_cluster_kernel<768>, &cfg)
                                             : cudaOccupancyMaxActiveClusters(&nclusters, ali_march_cluster_kernel<512>, &cfg);
            if (qe != cudaSuccess) { cudaGetLastError(); continue; }
            if (getenv("ALIFMM_DEBUG")) fprintf(stderr, "[alifmm] clusters of %d x %d threads resident at once: %d (batch: %d sources)\n", cand, cthreads, nclusters, n_src);
            // automatic choice: every source's cluster resident at once (a second wave would double the time)
            if (nclusters >= n_src || (c->cluster_size == cand && nclusters > 0)) { csize = cand; break; }
        }
        if (c->cluster_size == 1) csize = 1;
        if (pass == 1 && csize != 2) { csize = 2; cthreads = 512; break; }
    }
    c->cnt.cluster_size = csize;
    if (csize > 1) {
        int rc2;
        if ((rc2 = dev_reserve(c->ctl, (size_t)n_src * sizeof(AliClusterCtl))) != 0) return rc2;
        CUDA_TRY(cudaMemsetAsync(c->ctl.p, 0, (size_t)n_src * sizeof(AliClusterCtl), s));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(n_src * csize)); cfg.blockDim = dim3((unsigned)cthreads);
        cfg.dynamicSmemBytes = (size_t)ALI_MT_WORDS * 8; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        AliClusterCtl *ctlp = (AliClusterCtl *)c->ctl.p;
        if (cthreads == 768) CUDA_TRY(cudaLaunchKernelEx(&cfg, ali_march_cluster_kernel<768>, b, ctlp));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, ali_march_cluster_kernel<512>, b, ctlp));
    } else {
        // band entries + work lists in shared memory when they fit in c->band_smem_bytes
        size_t need = (size_t)b.band_cap * 16;
        int smem_cap = need <= (size_t)c->band_smem_bytes ? b.band_cap : 0;
        size_t smem = (smem_cap ? need : 0) + ALI_MARCH_SMEM_FIXED;
        if (c->threads_per_source >= 1024) {
            CUDA_TRY(cudaFuncSetAttribute(ali_march_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ali_march_kernel<1024><<<n_src, 1024, smem, s>>>(b, smem_cap);
        } else if (c->threads_per_source >= 896) {
            CUDA_TRY(cudaFuncSetAttribute(ali_march_kernel<896>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ali_march_kernel<896><<<n_src, 896, smem, s>>>(b, smem_cap);
        } else if (c->threads_per_source >= 768) {
            CUDA_TRY(cudaFuncSetAttribute(ali_march_kernel<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ali_march_kernel<768><<<n_src, 768, smem, s>>>(b, smem_cap);
        } else if (c->threads_per_source >= 640) {
            CUDA_TRY(cudaFuncSetAttribute(ali_march_kernel<640>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ali_march_kernel<640><<<n_src, 640, smem, s>>>(b, smem_cap);
        } else if (c->threads_per_source >= 512) {
            CUDA_TRY(cudaFuncSetAttribute(ali_march_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ali_march_kernel<512><<<n_src, 512, smem, s>>>(b, smem_cap);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(ali_march_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ali_march_kernel<256><<<n_src, 256, smem, s>>>(b, smem_cap);
        }
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev[2], s));
    int launches = 3;
    {
        size_t total = (size_t)n_src * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        ali_finalize_kernel<<<blocks, 256, 0, s>>>(b.Tt, b.st, b.T, n_src, fz, fx, b.tn, sg);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], s));
    CUDA_TRY(cudaMemcpyAsync(recs.data(), b.rec, recs.size() * sizeof(AliSourceRec), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));

    alifmm_counters_t &cn = c->cnt;
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); cn.ms_seq = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); cn.ms_march = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); cn.ms_finalize = ms;
    cn.node_solves = (int64_t)n_src * (int64_t)N;
    cn.seq_pops = cn.seq_evals = cn.band_rounds = cn.band_rounds_max = cn.band_evals = cn.fallback_evals = cn.max_band = 0;
    cn.kernel_launches = launches;
    cn.delta = b.delta;
    int overflow = 0;
    cn.seq_mcycles_min = cn.march_mcycles_min = 1e300;
    cn.seq_mcycles_max = cn.march_mcycles_max = 0.0;
    for (const AliSourceRec &r : recs) {
        const double sc = 1e-6 * (double)r.seq.cnt.cyc_total;
        const double mc = 1e-6 * (double)(r.cycles[0] + r.cycles[1] + r.cycles[2] + r.cycles[3]);
        if (sc < cn.seq_mcycles_min) cn.seq_mcycles_min = sc;
        if (sc > cn.seq_mcycles_max) cn.seq_mcycles_max = sc;
        if (mc < cn.march_mcycles_min) cn.march_mcycles_min = mc;
        if (mc > cn.march_mcycles_max) cn.march_mcycles_max = mc;
        cn.seq_pops += r.seq.cnt.pops;
        cn.seq_evals += r.seq.cnt.evals;
        cn.fallback_evals += r.seq.cnt.fallbacks + r.band_fallbacks;
        cn.band_rounds += r.rounds;
        if (r.rounds > cn.band_rounds_max) cn.band_rounds_max = r.rounds;
        cn.band_evals += r.band_evals;
        if (r.max_band > cn.max_band) cn.max_band = r.max_band;
        overflow |= r.overflow;
    }
    if (getenv("ALIFMM_DEBUG")) {
        const AliSourceRec &r = recs[0];
        fprintf(stderr, "[alifmm] source 0 seq: pops %lld evals %lld, steps %lld computed %lld, cycles/pop walk %.0f, cycles/step %.0f; Mcycles total %.0f fills %.0f starts %.0f\n", r.seq.cnt.pops,
                r.seq.cnt.evals, r.seq.cnt.steps, r.seq.cnt.computed, (double)r.seq.cnt.cyc_heap / (r.seq.cnt.pops + 1e-9),
                (double)r.seq.cnt.cyc_eval / (r.seq.cnt.steps + 1e-9), r.seq.cnt.cyc_total * 1e-6, r.seq.cnt.cyc_fill * 1e-6,
                r.seq.cnt.cyc_start * 1e-6);
        {
            int kmax = 0, kmin = 0;
            for (size_t k = 0; k < recs.size(); k++) {
                if (recs[k].seq.cnt.cyc_total > recs[kmax].seq.cnt.cyc_total) kmax = (int)k;
                if (recs[k].seq.cnt.cyc_total < recs[kmin].seq.cnt.cyc_total) kmin = (int)k;
            }
            for (int k : {kmin, kmax})
                fprintf(stderr, "[alifmm] seq %s source %d (z=%d,x=%d): Mcycles %.0f, pops %lld steps %lld, walk/pop %.0f, cycles/step %.0f\n",
                        k == kmin ? "fastest" : "slowest", k, recs[k].src_iz, recs[k].src_ix, recs[k].seq.cnt.cyc_total * 1e-6,
                        recs[k].seq.cnt.pops, recs[k].seq.cnt.steps, (double)recs[k].seq.cnt.cyc_heap / (recs[k].seq.cnt.pops + 1e-9),
                        (double)recs[k].seq.cnt.cyc_eval / (recs[k].seq.cnt.steps + 1e-9));
        }
        fprintf(stderr, "[alifmm] source 0: rounds %lld, cycles/round A %.0f B %.0f C %.0f sort %.0f, evals/round %.0f, band max %lld\n",
                r.rounds, (double)r.cycles[0] / (r.rounds + 1e-9), (double)r.cycles[1] / (r.rounds + 1e-9),
                (double)r.cycles[2] / (r.rounds + 1e-9), (double)r.cycles[3] / (r.rounds + 1e-9),
                (double)r.band_evals / (r.rounds + 1e-9), r.max_band);
    }
    if (overflow & 2) *band_overflow = 1;
    if (overflow & 2)
        return fail(ALIFMM_E_CAPACITY, "alifmm_ttf: narrow-band list overflowed; raise option band_capacity_factor");
    if (overflow & 1)
        return fail(ALIFMM_E_CAPACITY, "alifmm_ttf: sequential near-source scratch overflowed");
    c->n_slots = n_src; c->sg = sg; c->fz = fz; c->fx = fx;
    if (out_host) return copy_to_host(c, out_host, b.T, (size_t)n_src * N * sizeof(double));
    return ALIFMM_OK;
}

extern "C" int alifmm_ttf_shape(alifmm_ctx *c, int32_t *n_slots, int32_t *fz, int32_t *fx, int32_t *sg)
{
    if (!c) return fail(ALIFMM_E_INVALID, "alifmm_ttf_shape: null context");
    if (n_slots) *n_slots = c->n_slots;
    if (fz) *fz = c->fz;
    if (fx) *fx = c->fx;
    if (sg) *sg = c->sg;
    return ALIFMM_OK;
}

extern "C" int alifmm_ttf_fetch(alifmm_ctx *c, int32_t slot, double *out_host)
{
    if (!c || !out_host) return fail(ALIFMM_E_INVALID, "alifmm_ttf_fetch: null argument");
    if (slot < 0 || slot >= c->n_slots) return fail(ALIFMM_E_STATE, "alifmm_ttf_fetch: no such resident field");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t N = (size_t)c->fz * c->fx;
    return copy_to_host(c, out_host, (double *)c->T.p + (size_t)slot * N, N * sizeof(double));
}

// Validates the jobs, launches the ray kernel and leaves its outputs in the context's device buffers.
static int rays_launch(alifmm_ctx *c, int32_t n_rays, const int32_t *src_iz, const int32_t *src_ix,
                       const int32_t *rec_slot, int32_t cap, AliRayArgs &a, bool zero_paths = false)
{
    if (c->n_slots < 1) return fail(ALIFMM_E_STATE, "alifmm_rays: no resident travel-time fields (call alifmm_ttf first)");
    if (n_rays < 1) return fail(ALIFMM_E_INVALID, "alifmm_rays: n_rays must be >= 1");
    if (cap < 4) return fail(ALIFMM_E_INVALID, "alifmm_rays: capacity too small");
    std::vector<AliRayJob> jobs(n_rays);
    for (int r = 0; r < n_rays; r++) {
        if (rec_slot[r] < 0 || rec_slot[r] >= c->n_slots) return fail(ALIFMM_E_INVALID, "alifmm_rays: rec_slot out of range");
        if (src_iz[r] < 0 || src_iz[r] >= c->nz || src_ix[r] < 0 || src_ix[r] >= c->nx)
            return fail(ALIFMM_E_INVALID, "alifmm_rays: source node outside the grid");
        jobs[r].src_iz = src_iz[r]; jobs[r].src_ix = src_ix[r]; jobs[r].rec_slot = rec_slot[r];
    }
    // longest expected path first (distance source -> receiver)
    std::vector<int> order(n_rays + 1);
    {
        std::vector<long long> key(n_rays);
        for (int r = 0; r < n_rays; r++) {
            const long long dz = src_iz[r] - c->slot_iz[rec_slot[r]], dx = src_ix[r] - c->slot_ix[rec_slot[r]];
            key[r] = dz * dz + dx * dx;
            order[r + 1] = r;
        }
        std::stable_sort(order.begin() + 1, order.end(), [&](int p, int q) { return key[p] > key[q]; });
        order[0] = 0;   // the queue's counter
    }
    CUDA_TRY(cudaSetDevice(c->device));
    int rc;
    if ((rc = dev_reserve(c->jobs, (size_t)n_rays * sizeof(AliRayJob))) != 0) return rc;
    if ((rc = dev_reserve(c->ray_q, (size_t)(n_rays + 1) * sizeof(int))) != 0) return rc;
    if ((rc = dev_reserve(c->ray_x, (size_t)n_rays * cap * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->ray_y, (size_t)n_rays * cap * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->ray_time, (size_t)n_rays * sizeof(double))) != 0) return rc;
    if ((rc = dev_reserve(c->ray_len, (size_t)n_rays * sizeof(int))) != 0) return rc;
    if ((rc = dev_reserve(c->ray_flag, (size_t)n_rays * sizeof(int))) != 0) return rc;
    a.m = c->m; a.sg = c->sg; a.fz = c->fz; a.fx = c->fx;
    a.T = (const double *)c->T.p; a.rec = (const AliSourceRec *)c->rec.p; a.jobs = (const AliRayJob *)c->jobs.p;
    a.next = (int *)c->ray_q.p; a.order = a.next + 1;
    a.n_rays = n_rays; a.cap = cap;
    a.out_x = (double *)c->ray_x.p; a.out_y = (double *)c->ray_y.p; a.out_time = (double *)c->ray_time.p;
    a.out_len = (int *)c->ray_len.p; a.out_flag = (int *)c->ray_flag.p;
    a.maxc = ali_ray_max_candidates(c->sg);
    if (a.maxc < 32) a.maxc = 32;
    cudaStream_t s = c->stream;
    CUDA_TRY(cudaMemcpyAsync(c->jobs.p, jobs.data(), jobs.size() * sizeof(AliRayJob), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(c->ray_q.p, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    // (pageable source: the call returns once `jobs` has been staged)
    const size_t smem = (size_t)ALI_RAY_WARPS * 3 * a.maxc * sizeof(double);
    if (smem > 200 * 1024) return fail(ALIFMM_E_INVALID, "alifmm_rays: subgrid too large for the ray kernel's shared memory");
    void (*kern)(AliRayArgs) = c->ray_min_blocks >= 6 ? ali_rays_kernel<6> : c->ray_min_blocks == 5 ? ali_rays_kernel<5> : ali_rays_kernel<4>;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (zero_paths) {   // the whole [n_rays][cap] block goes to the caller: no stale bytes of a recycled buffer behind a path
        CUDA_TRY(cudaMemsetAsync(a.out_x, 0, (size_t)n_rays * cap * sizeof(double), s));
        CUDA_TRY(cudaMemsetAsync(a.out_y, 0, (size_t)n_rays * cap * sizeof(double), s));
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], s));
    int blocks = (n_rays + ALI_RAY_WARPS - 1) / ALI_RAY_WARPS;
    {   // no more CTAs than fit at once: the rest of the rays come through the queue
        int per_sm = 0, sms = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * ALI_RAY_WARPS, smem) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device) == cudaSuccess && per_sm > 0 && sms > 0) {
            if (blocks > per_sm * sms) blocks = per_sm * sms;
        } else {
            cudaGetLastError();
        }
    }
    kern<<<blocks, 32 * ALI_RAY_WARPS, smem, s>>>(a);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev[4], s));
    return ALIFMM_OK;
}

static void rays_counters(alifmm_ctx *c, int32_t n_rays, const int32_t *out_len, int launches)
{
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]);
    c->cnt.ms_rays = ms;
    c->cnt.rays = n_rays;
    c->cnt.ray_points = 0;
    for (int r = 0; r < n_rays; r++) c->cnt.ray_points += out_len[r];
    c->cnt.kernel_launches = launches;
}

extern "C" int alifmm_rays(alifmm_ctx *c, int32_t n_rays, const int32_t *src_iz, const int32_t *src_ix,
                           const int32_t *rec_slot, int32_t cap, double *out_x, double *out_y, int32_t *out_len,
                           double *out_time, int32_t *out_flag)
{
    if (!c || !src_iz || !src_ix || !rec_slot || !out_len || !out_time)
        return fail(ALIFMM_E_INVALID, "alifmm_rays: null argument");
    AliRayArgs a;
    int rc = rays_launch(c, n_rays, src_iz, src_ix, rec_slot, cap, a, out_x != nullptr || out_y != nullptr);
    if (rc != ALIFMM_OK) return rc;
    cudaStream_t s = c->stream;
    CUDA_TRY(cudaMemcpyAsync(out_len, a.out_len, (size_t)n_rays * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_time, a.out_time, (size_t)n_rays * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_flag) CUDA_TRY(cudaMemcpyAsync(out_flag, a.out_flag, (size_t)n_rays * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (out_x) CUDA_TRY(cudaMemcpyAsync(out_x, a.out_x, (size_t)n_rays * cap * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_y) CUDA_TRY(cudaMemcpyAsync(out_y, a.out_y, (size_t)n_rays * cap * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    rays_counters(c, n_rays, out_len, 1);
    return ALIFMM_OK;
}

// Packs the used part of every ray (its first len[r] points, divided by `divisor`) back to back.
__global__ void ali_ray_pack_kernel(const double *x, const double *y, const int *len, const long long *off, int cap,
                                    double divisor, double *px, double *py)
{
    const int r = blockIdx.x;
    const int n = len[r];
    const double *sx = x + (size_t)r * cap, *sy = y + (size_t)r * cap;
    double *dx = px + off[r], *dy = py + off[r];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        dx[k] = sx[k] / divisor;
        dy[k] = sy[k] / divisor;
    }
}

extern "C" int alifmm_rays_into(alifmm_ctx *c, int32_t n_rays, const int32_t *src_iz, const int32_t *src_ix,
                                const int32_t *rec_slot, int32_t cap, double divisor, const int64_t *row, double *base_x,
                                double *base_y, int32_t *out_len, double *out_time, int32_t *out_flag)
{
    if (!c || !src_iz || !src_ix || !rec_slot || !out_len || !out_time || !row || !base_x || !base_y)
        return fail(ALIFMM_E_INVALID, "alifmm_rays_into: null argument");
    if (!(divisor > 0)) return fail(ALIFMM_E_INVALID, "alifmm_rays_into: divisor must be positive");
    for (int r = 0; r < n_rays; r++)
        if (row[r] < 0) return fail(ALIFMM_E_INVALID, "alifmm_rays_into: negative row");
    AliRayArgs a;
    int rc = rays_launch(c, n_rays, src_iz, src_ix, rec_slot, cap, a);
    if (rc != ALIFMM_OK) return rc;
    cudaStream_t s = c->stream;
    CUDA_TRY(cudaMemcpyAsync(out_len, a.out_len, (size_t)n_rays * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_time, a.out_time, (size_t)n_rays * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_flag) CUDA_TRY(cudaMemcpyAsync(out_flag, a.out_flag, (size_t)n_rays * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    std::vector<long long> off(n_rays + 1);
    off[0] = 0;
    for (int r = 0; r < n_rays; r++) {
        if (out_len[r] < 0 || out_len[r] > cap) return fail(ALIFMM_E_STATE, "alifmm_rays_into: ray length out of range");
        off[r + 1] = off[r] + out_len[r];
    }
    const size_t total = (size_t)off[n_rays];
    if (total > 0) {
        if ((rc = dev_reserve(c->ray_off, (size_t)(n_rays + 1) * sizeof(long long))) != 0) return rc;
        if ((rc = dev_reserve(c->pack, 2 * total * sizeof(double))) != 0) return rc;
        double *px = (double *)c->pack.p, *py = px + total;
        CUDA_TRY(cudaMemcpyAsync(c->ray_off.p, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, s));
        ali_ray_pack_kernel<<<n_rays, 128, 0, s>>>(a.out_x, a.out_y, a.out_len, (const long long *)c->ray_off.p, cap, divisor, px, py);
        CUDA_TRY(cudaGetLastError());
        size_t pin_bytes = 0;
        double *pin = (double *)pin_take(2 * total * sizeof(double), &pin_bytes);
        if (!pin) return fail(ALIFMM_E_CUDA, "alifmm_rays_into: cannot allocate the pinned staging buffer");
        cudaError_t e = cudaMemcpyAsync(pin, px, 2 * total * sizeof(double), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            pin_give(pin, pin_bytes);
            return fail(ALIFMM_E_CUDA, std::string("alifmm_rays_into: ") + cudaGetErrorString(e));
        }
        // scatter into the caller's rows; mostly first-touch page faults of a fresh destination, so a few
        // host threads share it
        {
            unsigned nt = std::thread::hardware_concurrency();
            nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
            if (n_rays < 256) nt = 1;
            auto work = [&](int t) {
                for (int r = t; r < n_rays; r += (int)nt) {
                    const size_t n = (size_t)out_len[r];
                    memcpy(base_x + (size_t)row[r] * cap, pin + off[r], n * sizeof(double));
                    memcpy(base_y + (size_t)row[r] * cap, pin + total + off[r], n * sizeof(double));
                }
            };
            std::vector<std::thread> th;
            for (unsigned t = 1; t < nt; t++) th.emplace_back(work, (int)t);
            work(0);
            for (auto &x : th) x.join();
        }
        pin_give(pin, pin_bytes);
    }
    rays_counters(c, n_rays, out_len, total > 0 ? 2 : 1);
    return ALIFMM_OK;
}

extern "C" int alifmm_mem_info(alifmm_ctx *c, int64_t *free_bytes, int64_t *total_bytes)
{
    if (!c) return fail(ALIFMM_E_INVALID, "alifmm_mem_info: null context");
    CUDA_TRY(cudaSetDevice(c->device));
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)(f + pool_bytes(c->device));   // pooled buffers are reusable
    if (total_bytes) *total_bytes = (int64_t)t;
    return ALIFMM_OK;
}

extern "C" int alifmm_trim(int device)
{
    pool_trim(device);
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (PoolEntry &e : g_pin_pool) cudaFreeHost(e.p);
    g_pin_pool.clear();
    return ALIFMM_OK;
}

extern "C" int alifmm_counters(alifmm_ctx *c, alifmm_counters_t *out)
{
    if (!c || !out) return fail(ALIFMM_E_INVALID, "alifmm_counters: null argument");
    *out = c->cnt;
    return ALIFMM_OK;
}

extern "C" int alifmm_velocity_curves(alifmm_ctx *c, double c22, double c23, double c33, double c44, double density,
                                      double *group_out, double *phase_out)
{
    if (!c || !group_out || !phase_out) return fail(ALIFMM_E_INVALID, "alifmm_velocity_curves: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc;
    if ((rc = dev_reserve(c->misc, 2 * 361 * sizeof(double) + 64)) != 0) return rc;
    double *g = (double *)((char *)c->misc.p + 64), *p = g + 361;
    ali_curves_kernel<<<3, 128, 0, c->stream>>>(c22, c23, c33, c44, density, g, p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(group_out, g, 361 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(phase_out, p, 361 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return ALIFMM_OK;
}

extern "C" int alifmm_min_max_vel(alifmm_ctx *c, double *min_vel, double *max_vel)
{
    if (!c || !min_vel || !max_vel) return fail(ALIFMM_E_INVALID, "alifmm_min_max_vel: null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    // column extrema of the group table over all 361 rows (ATR:3755-3759) on the host: tiny
    std::vector<double> tab((size_t)361 * c->m.ncol), cmin(c->m.ncol), cmax(c->m.ncol);
    CUDA_TRY(cudaMemcpy(tab.data(), c->m.group_tab, tab.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int k = 0; k < c->m.ncol; k++) {
        double lo = tab[k], hi = tab[k];
        for (int a = 1; a < 361; a++) {
            double v = tab[(size_t)a * c->m.ncol + k];
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
        cmin[k] = lo; cmax[k] = hi;
    }
    const int first = c->first_velpn;
    int rc;
    const size_t need = 64 + 2 * (size_t)c->m.ncol * sizeof(double);
    if ((rc = dev_reserve(c->misc, need > 2 * 361 * sizeof(double) + 64 ? need : 2 * 361 * sizeof(double) + 64)) != 0) return rc;
    unsigned long long init[2] = {~0ull, 0ull};
    double *dmin = (double *)((char *)c->misc.p + 64), *dmax = dmin + c->m.ncol;
    CUDA_TRY(cudaMemcpyAsync(c->misc.p, init, 16, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dmin, cmin.data(), cmin.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dmax, cmax.data(), cmax.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const size_t n = (size_t)c->nz * c->nx;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ali_minmax_kernel<<<blocks, 256, 0, c->stream>>>(c->m, first, dmin, dmax, (unsigned long long *)c->misc.p,
                                                     (unsigned long long *)c->misc.p + 1);
    CUDA_TRY(cudaGetLastError());
    unsigned long long res[2];
    CUDA_TRY(cudaMemcpyAsync(res, c->misc.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(min_vel, &res[0], 8);
    memcpy(max_vel, &res[1], 8);
    return ALIFMM_OK;
}

// ---------------------------------------------------------------------------
// context-free entries: material tables for a batch of materials, node-level operator check
// ---------------------------------------------------------------------------
static int pick_device(int device, const char *who)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ALIFMM_E_CUDA, std::string(who) + ": no CUDA device available (this library has no CPU path)");
    }
    if (device < 0 || device >= ndev) return fail(ALIFMM_E_INVALID, std::string(who) + ": device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    return ALIFMM_OK;
}

extern "C" int alifmm_velocity_curves_batch(int device, int32_t n_mat, const double *props, double *group_out,
                                            double *phase_out)
{
    if (!props || !group_out || !phase_out) return fail(ALIFMM_E_INVALID, "alifmm_velocity_curves_batch: null argument");
    if (n_mat < 1 || n_mat > 1000000) return fail(ALIFMM_E_INVALID, "alifmm_velocity_curves_batch: n_mat out of range");
    int rc = pick_device(device, "alifmm_velocity_curves_batch");
    if (rc != ALIFMM_OK) return rc;
    const size_t np = (size_t)n_mat * 5, no = (size_t)n_mat * 361;
    DevBuf buf;
    if ((rc = dev_reserve(buf, (np + 2 * no) * sizeof(double))) != 0) return rc;
    double *dp = (double *)buf.p, *dg = dp + np, *dph = dg + no;
    cudaError_t e = cudaMemcpy(dp, props, np * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        ali_curves_batch_kernel<<<(int)((no + 127) / 128), 128>>>(n_mat, dp, dg, dph);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(group_out, dg, no * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(phase_out, dph, no * sizeof(double), cudaMemcpyDeviceToHost);
    dev_release(buf, device);
    if (e != cudaSuccess) return fail(ALIFMM_E_CUDA, std::string("alifmm_velocity_curves_batch: ") + cudaGetErrorString(e));
    return ALIFMM_OK;
}

extern "C" int alifmm_eval_nodes(int device, int32_t n, int32_t nz, int32_t nx, double dnx, const double *veln,
                                 const int32_t *velpn, const double *vel_map, const int64_t *stif_den, int32_t has_stif,
                                 const double *group_vel, const double *phase_vel, int32_t n_cols, const double *ttn,
                                 const int32_t *nsts, const int32_t *pos, double *out_update, double *out_fouds,
                                 int32_t *out_stencil)
{
    if (!veln || !velpn || !vel_map || !group_vel || !phase_vel || !ttn || !nsts || !pos || !out_update || !out_fouds)
        return fail(ALIFMM_E_INVALID, "alifmm_eval_nodes: null argument");
    if (n < 1 || nz < 1 || nx < 1 || n_cols < 1 || !(dnx > 0)) return fail(ALIFMM_E_INVALID, "alifmm_eval_nodes: bad sizes");
    const size_t nn = (size_t)nz * nx, tot = (size_t)n * nn;
    if (tot > 100000000u) return fail(ALIFMM_E_INVALID, "alifmm_eval_nodes: too many nodes");
    for (size_t i = 0; i < tot; i++)
        if (velpn[i] < 0 || velpn[i] >= n_cols) return fail(ALIFMM_E_INVALID, "alifmm_eval_nodes: velpn outside the velocity tables");
    for (int c = 0; c < n; c++)
        if (pos[2 * c] < 0 || pos[2 * c] >= nz || pos[2 * c + 1] < 0 || pos[2 * c + 1] >= nx)
            return fail(ALIFMM_E_INVALID, "alifmm_eval_nodes: node outside its grid");
    int rc = pick_device(device, "alifmm_eval_nodes");
    if (rc != ALIFMM_OK) return rc;
    std::vector<DevBuf> bufs;
    auto cleanup = [&]() { for (DevBuf &b : bufs) dev_release(b, device); };
    auto up = [&](const void *host, size_t bytes, void **out) -> int {
        DevBuf b;
        int r = dev_reserve(b, bytes < 8 ? 8 : bytes);
        if (r != ALIFMM_OK) return r;
        bufs.push_back(b);
        if (host && cudaMemcpy(b.p, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(ALIFMM_E_CUDA, "alifmm_eval_nodes: upload failed");
        *out = b.p;
        return ALIFMM_OK;
    };
    void *dv = nullptr, *dpn = nullptr, *dm = nullptr, *ds = nullptr, *dg = nullptr, *dph = nullptr, *dt = nullptr, *dst = nullptr,
         *dpos = nullptr, *drec = nullptr, *dou = nullptr, *dof = nullptr, *dos = nullptr;
    if ((rc = up(veln, tot * 8, &dv)) || (rc = up(velpn, tot * 4, &dpn)) || (rc = up(vel_map, tot * 8, &dm)) ||
        (stif_den && (rc = up(stif_den, tot * 40, &ds))) || (rc = up(group_vel, (size_t)361 * n_cols * 8, &dg)) ||
        (rc = up(phase_vel, (size_t)361 * n_cols * 8, &dph)) || (rc = up(ttn, tot * 8, &dt)) || (rc = up(nsts, tot * 4, &dst)) ||
        (rc = up(pos, (size_t)n * 8, &dpos)) || (rc = up(nullptr, tot * sizeof(AliMatRec), &drec)) ||
        (rc = up(nullptr, (size_t)n * 8, &dou)) || (rc = up(nullptr, (size_t)n * 8, &dof)) || (rc = up(nullptr, (size_t)n * 4, &dos))) {
        cleanup();
        return rc;
    }
    ali_records_kernel<<<(int)((tot + 255) / 256), 256>>>((int)tot, (const double *)dv, (const int32_t *)dpn, (const double *)dm,
                                                         (const long long *)ds, (AliMatRec *)drec, nullptr);
    AliModel m{};
    m.nz = nz; m.nx = nx; m.rec = (const AliMatRec *)drec; m.has_stif = has_stif ? 1 : 0;
    m.group_tab = (const double *)dg; m.phase_tab = (const double *)dph; m.ncol = n_cols; m.dnx = dnx;
    ali_eval_nodes_kernel<<<(n + 63) / 64, 64>>>(n, m, (const double *)dt, (const int32_t *)dst, (const int32_t *)dpos,
                                                (double *)dou, (double *)dof, (int32_t *)dos);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(out_update, dou, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_fouds, dof, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_stencil) e = cudaMemcpy(out_stencil, dos, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return fail(ALIFMM_E_CUDA, std::string("alifmm_eval_nodes: ") + cudaGetErrorString(e));
    return ALIFMM_OK;
}

// ---------------------------------------------------------------------------
// one field on two GPUs: row strips with a 2-row halo over peer memory (ali_strip.cuh)
// ---------------------------------------------------------------------------
#include "ali_strip.cuh"

namespace {
struct StripDev {
    int device = -1;
    int zlo = 0, zhi = 0, za = 0, zb = 0;   // rows owned [zlo, zhi); rows allocated [za, zb) (multiples of 4, halo included)
    cudaStream_t stream = nullptr;
    std::vector<void *> allocs;
    AliModel m{};
    const AliModel *m_dev = nullptr;
    double *Tt = nullptr; uint8_t *st = nullptr;      // allocation starts (not offset)
    double *out = nullptr;
    unsigned *lists = nullptr; double *stage = nullptr;
    AliClusterCtl *ctl = nullptr; AliStripXchg *xchg = nullptr; AliSourceRec *rec = nullptr;
    double *seq_t = nullptr; int32_t *seq_s = nullptr; int32_t *seq_heap = nullptr; double *seq_hkey = nullptr, *seq_cval = nullptr;
    uint8_t *seq_cflag = nullptr;
    double vmax = 0.0;
};
}

// Rows owned by the strips of alifmm_ttf_split: strip k owns [rows[k], rows[k + 1]).  Equal shares, boundaries on multiples
// of 4 (tile rows); a boundary closer to the source row than the sequential phase's window (hand-over radius + 8 rows)
// is pushed away from it.  Pure host arithmetic (no device needed).
extern "C" int alifmm_split_rows(int32_t nz, int32_t n_dev, int32_t src_iz, int32_t split_row, int32_t *rows)
{
    if (!rows) return fail(ALIFMM_E_INVALID, "alifmm_split_rows: null argument");
    if (n_dev < 2 || n_dev > ALI_MAX_STRIPS) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: 2 ... 8 devices (a chain of row strips)");
    if (nz < 64 || nz > 65535) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: 64 ... 65535 rows");
    if (src_iz < 0 || src_iz >= nz) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: source node outside the grid");
    if (split_row > 0 && n_dev != 2) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: an explicit split row only with two devices");
    AliSourcePlan plan;
    AliModel pm{}; pm.nz = nz; pm.nx = 64;
    ali_make_plan(plan, pm, src_iz, 0, 1, 27);
    const int keep = plan.stop_r + 8;   // the sequential phase's window (stop_r + 4 each way) must lie inside one strip
    int *zb = rows;
    zb[0] = 0; zb[n_dev] = nz;
    for (int k = 1; k < n_dev; k++) zb[k] = (int)((long long)nz * k / n_dev) & ~3;
    if (split_row > 0) zb[1] = split_row & ~3;
    else
        for (int k = 1; k < n_dev; k++) {
            if (zb[k] <= src_iz && src_iz - zb[k] + 1 < keep) zb[k] = (src_iz - keep) & ~3;         // boundary above the source: up
            else if (zb[k] > src_iz && zb[k] - src_iz < keep) zb[k] = (src_iz + keep + 3) & ~3;      // below: down
        }
    for (int k = 1; k < n_dev; k++) {
        const bool near = zb[k] <= src_iz ? (src_iz - zb[k] + 1 < keep) : (zb[k] - src_iz < keep);
        if (zb[k] < 8 || zb[k] > nz - 8 || zb[k] - zb[k - 1] < 16 || near)
            return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: no admissible split rows (strips of at least 16 rows, the source's refined neighbourhood inside one strip)");
    }
    if (nz - zb[n_dev - 1] < 8) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: no admissible split rows");
    return ALIFMM_OK;
}

extern "C" int alifmm_ttf_split(const alifmm_model_desc *d, int32_t n_dev, const int32_t *devices, int32_t src_iz,
                                int32_t src_ix, int32_t split_row, double *out_host, alifmm_counters_t *counters)
{
    if (!d || !devices || !out_host) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: null argument");
    if (n_dev < 2 || n_dev > ALI_MAX_STRIPS) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: 2 ... 8 devices (a chain of row strips)");
    if (d->nz < 64 || d->nx < 1 || !(d->dnx > 0) || !d->veln || !d->velpn || !d->vel_map || !d->group_vel || !d->phase_vel || d->n_cols < 1)
        return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: bad model descriptor (at least 64 rows)");
    if (d->nz > 65535 || d->nx > 65535) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: grid side exceeds 65535 nodes");
    if (src_iz < 0 || src_iz >= d->nz || src_ix < 0 || src_ix >= d->nx) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: source node outside the grid");
    if (split_row > 0 && n_dev != 2) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: an explicit split row only with two devices");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < n_dev) { cudaGetLastError(); return fail(ALIFMM_E_CUDA, "alifmm_ttf_split: needs as many CUDA devices as strips"); }
    for (int k = 0; k < n_dev; k++) {
        if (devices[k] < 0 || devices[k] >= ndev) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: device index out of range");
        for (int j = 0; j < k; j++)
            if (devices[j] == devices[k]) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: the devices must differ");
    }
    const int nz = d->nz, nx = d->nx, margin = 27;
    AliSourcePlan plan;
    AliModel pm{}; pm.nz = nz; pm.nx = nx;
    ali_make_plan(plan, pm, src_iz, src_ix, 1, margin);
    int zb[ALI_MAX_STRIPS + 1];
    {
        const int rrc = alifmm_split_rows(nz, n_dev, src_iz, split_row, zb);
        if (rrc != ALIFMM_OK) return rrc;
    }
    for (size_t i = 0; i < (size_t)nz * nx; i++)
        if (d->velpn[i] < 0 || d->velpn[i] >= d->n_cols) return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: velpn holds a material id outside the velocity tables");

    StripDev S[ALI_MAX_STRIPS];
    std::string err;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int ev_device = -1;
    auto cleanup = [&]() {
        if (ev_device >= 0) {
            cudaSetDevice(ev_device);
            for (auto &e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        }
        for (StripDev &s : S) {
            if (s.device < 0) continue;
            cudaSetDevice(s.device);
            if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
            for (void *p : s.allocs) cudaFree(p);
        }
    };
#define STRIP_TRY(expr)                                                                             \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e__); cleanup(); return fail(ALIFMM_E_CUDA, "alifmm_ttf_split: " + err); } \
    } while (0)
    const size_t t4x = (size_t)((nx + 3) >> 2);
    const int band_cap = (int)(6.0 * (double)(nz + nx)) + 1024;
    int owner = 0;
    for (int k = 0; k < n_dev; k++)
        if (src_iz >= zb[k] && src_iz < zb[k + 1]) owner = k;
    size_t seq_cap = 0; int heap_cap = 0;
    {
        size_t lvl = ali_plan_max_level_nodes(plan);
        size_t w = (size_t)(2 * (plan.stop_r + 4) + 1);
        seq_cap = lvl > w * w ? lvl : w * w;
        heap_cap = (int)(seq_cap / 2 + 64);
    }
    // every strip is set up by its own host thread: the uploads of the strips' model rows (pageable host memory, one
    // PCIe link per GPU) run side by side
    std::string serr[ALI_MAX_STRIPS];
    int src_codes[ALI_MAX_STRIPS] = {0};
#define STRIP_TRY_K(expr)                                                                           \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) { serr[k] = std::string(#expr) + ": " + cudaGetErrorString(e__); return (int)ALIFMM_E_CUDA; } \
    } while (0)
    auto setup = [&](int k) -> int {
        StripDev &s = S[k];
        s.device = devices[k];
        s.zlo = zb[k]; s.zhi = zb[k + 1];
        s.za = k == 0 ? 0 : s.zlo - 4; s.zb = k == n_dev - 1 ? ((nz + 3) & ~3) : s.zhi + 4;
        STRIP_TRY_K(cudaSetDevice(s.device));
        for (int j = 0; j < n_dev; j++) {
            if (j == k) continue;
            int can = 0;
            STRIP_TRY_K(cudaDeviceCanAccessPeer(&can, s.device, devices[j]));
            if (!can) { serr[k] = "the devices cannot access each other's memory"; return (int)ALIFMM_E_CUDA; }
            cudaError_t pe = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) STRIP_TRY_K(pe);
            cudaGetLastError();
        }
        STRIP_TRY_K(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        auto alloc = [&](size_t bytes, void **out) -> cudaError_t {
            cudaError_t e = cudaMalloc(out, bytes < 256 ? 256 : bytes);
            if (e == cudaSuccess) s.allocs.push_back(*out);
            return e;
        };
        // model rows [za, min(zb, nz)) -> records
        const int r0 = s.za, r1 = s.zb < nz ? s.zb : nz;
        const size_t nn = (size_t)(r1 - r0) * nx, off = (size_t)r0 * nx;
        void *dv, *dp, *dm, *ds = nullptr, *drec, *dg, *dph, *dmod, *dbad;
        STRIP_TRY_K(alloc(nn * 8, &dv)); STRIP_TRY_K(alloc(nn * 4, &dp)); STRIP_TRY_K(alloc(nn * 8, &dm));
        if (d->stif_den) STRIP_TRY_K(alloc(nn * 40, &ds));
        STRIP_TRY_K(alloc(nn * sizeof(AliMatRec), &drec)); STRIP_TRY_K(alloc(64, &dbad));
        STRIP_TRY_K(cudaMemcpyAsync(dv, d->veln + off, nn * 8, cudaMemcpyHostToDevice, s.stream));
        STRIP_TRY_K(cudaMemcpyAsync(dp, d->velpn + off, nn * 4, cudaMemcpyHostToDevice, s.stream));
        STRIP_TRY_K(cudaMemcpyAsync(dm, d->vel_map + off, nn * 8, cudaMemcpyHostToDevice, s.stream));
        if (ds) STRIP_TRY_K(cudaMemcpyAsync(ds, d->stif_den + off * 5, nn * 40, cudaMemcpyHostToDevice, s.stream));
        STRIP_TRY_K(cudaMemsetAsync(dbad, 0, 64, s.stream));
        {
            int blocks = (int)((nn + 255) / 256);
            if (blocks > 148 * 8) blocks = 148 * 8;
            ali_records_kernel<<<blocks, 256, 0, s.stream>>>((int)nn, (const double *)dv, (const int32_t *)dp, (const double *)dm,
                                                            (const long long *)ds, (AliMatRec *)drec, (int *)dbad);
        }
        STRIP_TRY_K(alloc((size_t)361 * d->n_cols * 8, &dg)); STRIP_TRY_K(alloc((size_t)361 * d->n_cols * 8, &dph));
        STRIP_TRY_K(cudaMemcpyAsync(dg, d->group_vel, (size_t)361 * d->n_cols * 8, cudaMemcpyHostToDevice, s.stream));
        STRIP_TRY_K(cudaMemcpyAsync(dph, d->phase_vel, (size_t)361 * d->n_cols * 8, cudaMemcpyHostToDevice, s.stream));
        // the strip's own maximum of the phase velocity (on its rows, with a strip-local model view)
        AliModel ml{};
        ml.nz = r1 - r0; ml.nx = nx; ml.rec = (const AliMatRec *)drec; ml.has_stif = d->has_stif ? 1 : 0;
        ml.group_tab = (const double *)dg; ml.phase_tab = (const double *)dph; ml.ncol = d->n_cols; ml.dnx = d->dnx;
        unsigned long long *dvmax = (unsigned long long *)dbad + 1;
        {
            int blocks = (int)((nn + 255) / 256);
            if (blocks > 148 * 8) blocks = 148 * 8;
            ali_vmax_kernel<<<blocks, 256, 0, s.stream>>>(ml, dvmax);
        }
        unsigned long long hb[2] = {0, 0};
        STRIP_TRY_K(cudaMemcpyAsync(hb, dbad, 16, cudaMemcpyDeviceToHost, s.stream));
        STRIP_TRY_K(cudaStreamSynchronize(s.stream));
        if ((int)hb[0]) { serr[k] = "veln / vel_map hold non-finite values"; return (int)ALIFMM_E_INVALID; }
        memcpy(&s.vmax, &hb[1], 8);
        // the march's model: full-grid extents, records offset so that (z * nx + x) indexes the strip's rows
        s.m = ml; s.m.nz = nz; s.m.rec = (const AliMatRec *)drec - off;
        STRIP_TRY_K(alloc(sizeof(AliModel), &dmod));
        STRIP_TRY_K(cudaMemcpyAsync(dmod, &s.m, sizeof(AliModel), cudaMemcpyHostToDevice, s.stream));
        s.m_dev = (const AliModel *)dmod;
        // field strip (tiled), alive flags, result rows, lists, control / exchange blocks, record
        const size_t tn = (size_t)((s.zb - s.za) >> 2) * t4x * 16;
        STRIP_TRY_K(alloc(tn * 8, (void **)&s.Tt)); STRIP_TRY_K(alloc(tn + 16, (void **)&s.st));
        STRIP_TRY_K(alloc((size_t)(s.zhi - s.zlo) * nx * 8, (void **)&s.out));
        STRIP_TRY_K(alloc((size_t)4 * band_cap * 4, (void **)&s.lists)); STRIP_TRY_K(alloc((size_t)2 * band_cap * 8, (void **)&s.stage));
        STRIP_TRY_K(alloc(sizeof(AliClusterCtl), (void **)&s.ctl)); STRIP_TRY_K(alloc(ALI_MAX_STRIPS * sizeof(AliStripXchg), (void **)&s.xchg));
        STRIP_TRY_K(alloc(sizeof(AliSourceRec), (void **)&s.rec));
        STRIP_TRY_K(cudaMemsetAsync(s.Tt, ALI_T_UNSET_BYTE, tn * 8, s.stream));
        STRIP_TRY_K(cudaMemsetAsync(s.st, 0, tn + 16, s.stream));
        STRIP_TRY_K(cudaMemsetAsync(s.lists, 0, (size_t)4 * band_cap * 4, s.stream));   // (stale slots must stay valid indices after an overflow)
        STRIP_TRY_K(cudaMemsetAsync(s.stage, 0, (size_t)2 * band_cap * 8, s.stream));
        STRIP_TRY_K(cudaMemsetAsync(s.ctl, 0, sizeof(AliClusterCtl), s.stream));
        STRIP_TRY_K(cudaMemsetAsync(s.xchg, 0, ALI_MAX_STRIPS * sizeof(AliStripXchg), s.stream));
        AliSourceRec hr;
        memset(&hr, 0, sizeof hr);
        hr.src_iz = src_iz; hr.src_ix = src_ix;
        STRIP_TRY_K(cudaMemcpyAsync(s.rec, &hr, sizeof hr, cudaMemcpyHostToDevice, s.stream));
        if (k == owner) {
            STRIP_TRY_K(alloc(2 * seq_cap * 8, (void **)&s.seq_t)); STRIP_TRY_K(alloc(2 * seq_cap * 4, (void **)&s.seq_s));
            STRIP_TRY_K(alloc((size_t)2 * heap_cap * 4, (void **)&s.seq_heap)); STRIP_TRY_K(alloc(ALI_HKEY_SLOTS(heap_cap) * 8, (void **)&s.seq_hkey));
            STRIP_TRY_K(alloc(seq_cap * 8, (void **)&s.seq_cval)); STRIP_TRY_K(alloc(seq_cap + 16, (void **)&s.seq_cflag));
        }
        STRIP_TRY_K(cudaStreamSynchronize(s.stream));
    
        return (int)ALIFMM_OK;
    };
    {
        std::vector<std::thread> workers;
        for (int k = 0; k < n_dev; k++) workers.emplace_back([&, k]() { src_codes[k] = setup(k); });
        for (std::thread &w : workers) w.join();
    }
#undef STRIP_TRY_K
    for (int k = 0; k < n_dev; k++)
        if (src_codes[k] != ALIFMM_OK) { cleanup(); return fail(src_codes[k], "alifmm_ttf_split: " + serr[k]); }
    double vmax = 0.0;
    for (int k = 0; k < n_dev; k++) vmax = S[k].vmax > vmax ? S[k].vmax : vmax;
    if (!(vmax > 0) || !isfinite(vmax)) { cleanup(); return fail(ALIFMM_E_INVALID, "alifmm_ttf_split: model has no positive finite phase velocity"); }
    AliStripArgs A[ALI_MAX_STRIPS];
    for (int k = 0; k < n_dev; k++) {
        StripDev &s = S[k];
        memset(&A[k], 0, sizeof A[k]);
        AliBatch &b = A[k].b;
        memset(&b, 0, sizeof b);
        b.m = s.m; b.m_dev = s.m_dev; b.sg = 1; b.nz = nz; b.nx = nx; b.margin = margin;
        b.delta = 0.35 * d->dnx / vmax;
        b.Tt = s.Tt - (size_t)(s.za >> 2) * t4x * 16;
        b.st = s.st - (size_t)(s.za >> 2) * t4x * 16;
        b.tn = 0;
        b.seq_t = s.seq_t; b.seq_s = s.seq_s; b.seq_heap = s.seq_heap; b.seq_hkey = s.seq_hkey; b.seq_cval = s.seq_cval; b.seq_cflag = s.seq_cflag;
        b.seq_cap = seq_cap; b.heap_cap = heap_cap;
        b.lists = s.lists; b.stage = s.stage; b.band_cap = band_cap; b.resort_every = 8; b.rec = s.rec;
        A[k].ctl = s.ctl; A[k].xl = s.xchg; A[k].zlo = s.zlo; A[k].zhi = s.zhi; A[k].has_seq = k == owner;
        A[k].n_strips = n_dev; A[k].me = k;
        A[k].spin_limit = 400000000LL;   // ~ a minute of 64 ns sleeps: a dead peer ends the kernel instead of hanging the GPU
    }
    for (int k = 0; k < n_dev; k++) {
        for (int j = 0; j < n_dev; j++) A[k].px[j] = S[j].xchg;
        for (int side = 0; side < 2; side++) {
            const int j = side == 0 ? k - 1 : k + 1;
            if (j < 0 || j >= n_dev) continue;
            const StripDev &p = S[j];
            A[k].nb[side].T = p.Tt - (size_t)(p.za >> 2) * t4x * 16;
            A[k].nb[side].st = p.st - (size_t)(p.za >> 2) * t4x * 16;
            A[k].nb[side].ctl = p.ctl; A[k].nb[side].lists = p.lists; A[k].nb[side].stage = p.stage;
        }
    }
    STRIP_TRY(cudaSetDevice(S[owner].device));
    ev_device = S[owner].device;
    for (auto &e : ev) STRIP_TRY(cudaEventCreate(&e));
    STRIP_TRY(cudaEventRecord(ev[0], S[owner].stream));
    ali_seq_kernel<<<1, 32, 0, S[owner].stream>>>(A[owner].b);
    STRIP_TRY(cudaGetLastError());
    STRIP_TRY(cudaEventRecord(ev[1], S[owner].stream));
    STRIP_TRY(cudaStreamSynchronize(S[owner].stream));
    for (int k = 0; k < n_dev; k++) {
        STRIP_TRY(cudaSetDevice(S[k].device));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8); cfg.blockDim = dim3(512);
        cfg.dynamicSmemBytes = (size_t)ALI_MT_WORDS * 8; cfg.stream = S[k].stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        STRIP_TRY(cudaLaunchKernelEx(&cfg, ali_march_strip_kernel<512>, A[k]));
    }
    AliSourceRec recs[ALI_MAX_STRIPS];
    for (int k = 0; k < n_dev; k++) {
        STRIP_TRY(cudaSetDevice(S[k].device));
        if (k == owner) STRIP_TRY(cudaEventRecord(ev[2], S[k].stream));
        const size_t total = (size_t)(S[k].zhi - S[k].zlo) * nx;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        ali_finalize_rows_kernel<<<blocks, 256, 0, S[k].stream>>>(A[k].b.Tt, A[k].b.st, S[k].out, S[k].zlo, S[k].zhi, nx);
        STRIP_TRY(cudaMemcpyAsync(&recs[k], S[k].rec, sizeof(AliSourceRec), cudaMemcpyDeviceToHost, S[k].stream));
    }
    for (int k = 0; k < n_dev; k++) {
        STRIP_TRY(cudaSetDevice(S[k].device));
        STRIP_TRY(cudaStreamSynchronize(S[k].stream));
    }
    int ovf = 0;
    for (int k = 0; k < n_dev; k++) ovf |= recs[k].overflow;
    if (ovf) {
        cleanup();
        if (ovf & 4) return fail(ALIFMM_E_CUDA, "alifmm_ttf_split: a GPU stopped answering the per-round exchange");
        if (ovf & 2) return fail(ALIFMM_E_CAPACITY, "alifmm_ttf_split: narrow-band list overflowed");
        return fail(ALIFMM_E_CAPACITY, "alifmm_ttf_split: sequential near-source scratch overflowed");
    }
    {   // the strips' rows come back side by side as well
        cudaError_t cerr[ALI_MAX_STRIPS];
        std::vector<std::thread> workers;
        for (int k = 0; k < n_dev; k++)
            workers.emplace_back([&, k]() {
                cerr[k] = cudaSetDevice(S[k].device);
                if (cerr[k] == cudaSuccess)
                    cerr[k] = cudaMemcpy(out_host + (size_t)S[k].zlo * nx, S[k].out, (size_t)(S[k].zhi - S[k].zlo) * nx * 8, cudaMemcpyDeviceToHost);
            });
        for (std::thread &w : workers) w.join();
        for (int k = 0; k < n_dev; k++) STRIP_TRY(cerr[k]);
    }
    if (counters) {
        memset(counters, 0, sizeof *counters);
        float ms = 0;
        cudaSetDevice(S[owner].device);
        cudaEventElapsedTime(&ms, ev[0], ev[1]); counters->ms_seq = ms;
        cudaEventElapsedTime(&ms, ev[1], ev[2]); counters->ms_march = ms;
        counters->node_solves = (int64_t)nz * nx;
        counters->seq_pops = recs[owner].seq.cnt.pops; counters->seq_evals = recs[owner].seq.cnt.evals;
        counters->band_rounds = counters->band_rounds_max = recs[0].rounds;
        for (int k = 0; k < n_dev; k++) {
            counters->band_evals += recs[k].band_evals;
            counters->fallback_evals += recs[k].band_fallbacks;
            counters->max_band += recs[k].max_band;
        }
        counters->fallback_evals += recs[owner].seq.cnt.fallbacks;
        counters->kernel_launches = 1 + 2 * n_dev;
        counters->vmax = vmax; counters->delta = 0.35 * d->dnx / vmax;
        counters->cluster_size = 8; counters->seq_threads = 32;
    }
    cleanup();
#undef STRIP_TRY
    return ALIFMM_OK;
}
