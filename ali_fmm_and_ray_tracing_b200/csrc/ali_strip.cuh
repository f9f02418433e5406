// ali_strip.cuh -- one travel-time field decomposed into row strips on 2 ... 8 GPUs (BASELINE config 5).
//
// The reference has no counterpart (a single field is one heap, ATR:2055-2102); SURVEY.md 8(e) asks for row
// strips with a 2-row halo (the stencil reaches two nodes, ATR:940-987) exchanged over NVLink and a scalar
// exchange per band round, with no NCCL on the data path.  The value is CAPACITY: each GPU holds only its
// strip of the model (64 B per node) and of the field; a single source does not get faster (the band is
// O(grid side) nodes per round, and every round now costs three inter-GPU barriers).
//
// How.  Each GPU runs the cluster form of the band march (ali_march_cluster_kernel) on the band nodes of its
// rows, with full-grid coordinates: its buffers are offset so that the tiled index of (z, x) works unchanged.
// Everything that crosses the strip boundary goes through peer-mapped memory (cudaDeviceEnablePeerAccess):
//   * a node published within two rows of the boundary is also stored into the peer's halo copy of the field,
//     and marks the peer's window-change bitmap (plain stores / RED over NVLink);
//   * an accepted node on the boundary row claims its neighbour across the boundary with an atomicCAS on the
//     PEER's field word (the owner's copy is the authority) and appends it to the PEER's band list (atomicAdd on
//     the peer's counters);
//   * the round's minimum, the band length (termination) and the overflow / force flags are written into the
//     peer's exchange block;
//   * the three barriers of a round become: cluster barrier, system fence, epoch flag stored to the peer and
//     awaited from the peer, cluster barrier.
// With more than two strips a GPU has a peer above and a peer below for the halo traffic, and the scalars of a round
// (minimum, band length, flags, barrier epoch) go to EVERY other GPU's exchange block, one slot per writer.
// The set of nodes evaluated, published, accepted and enlisted per round is that of the one-GPU kernel, so the
// field is bit-identical to it (tests/test_gpu_parity.py::test_two_gpu_strips_equal_one_gpu, ..._four_...).
#pragma once

#define ALI_MAX_STRIPS 8

struct AliStripXchg {                 // slot w of a GPU's exchange block: written by strip w over NVLink, read locally
    unsigned long long flag;          // inter-GPU barrier epoch
    unsigned long long evalmin[2], basemin[2];   // the writer's minima, by round parity
    int count[2];                     // the writer's band length at the start of the round, by round parity
    int force[2];
    int overflow, pad;
};

struct AliStripPeer {           // the strip above (index 0) / below (index 1): field, alive flags, control block, lists
    double *T;
    uint8_t *st;
    AliClusterCtl *ctl;
    unsigned *lists;            // ent0 | ent1 | wrk0 | wrk1, band_cap entries each
    double *stage;              // val0 | val1
};

struct AliStripArgs {
    AliBatch b;                 // nz / nx: the FULL grid; Tt, st and the model records are offset to full-grid indexing
    AliClusterCtl *ctl;         // local control block (also written by the neighbours: list counters, bitmap)
    AliStripXchg *xl;           // local exchange block: ALI_MAX_STRIPS slots, slot w written by strip w
    int n_strips, me;
    int zlo, zhi;               // rows this GPU owns
    int has_seq;                // this GPU ran the sequential near-source phase (it owns the source)
    AliStripPeer nb[2];         // above (rows < zlo), below (rows >= zhi); null pointers at the ends of the chain
    AliStripXchg *px[ALI_MAX_STRIPS];   // every strip's exchange block (px[me] == xl)
    long long spin_limit;       // iterations a GPU waits for a peer before giving up (overflow code 4)
};

// Inter-GPU barrier.  Every thread fences its remote stores at system scope, the cluster meets, thread w (w < n_strips,
// w != me) announces the epoch to strip w and waits for strip w's, the cluster meets again (its acquire side also
// invalidates L1, so plain loads see what the peers stored into local memory).
__device__ __forceinline__ bool ali_strip_sync(const AliStripArgs &a, unsigned long long &epoch, int gtid)
{
    __threadfence_system();
    ali_cluster_sync();
    epoch++;
    if (gtid < a.n_strips && gtid != a.me) {
        *(volatile unsigned long long *)&a.px[gtid][a.me].flag = epoch;
        long long spins = 0;
        while (*(volatile unsigned long long *)&a.xl[gtid].flag < epoch) {
            if (++spins > a.spin_limit) { a.ctl->overflow = 4; break; }
            __nanosleep(64);
        }
        __threadfence_system();
    }
    ali_cluster_sync();
    return ali_ldv(&a.ctl->overflow) != 4;
}

template <int NT>
__global__ void __launch_bounds__(NT) ali_march_strip_kernel(AliStripArgs a)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint64_t *s_sincos = reinterpret_cast<uint64_t *>(s_raw);
    const AliBatch &b = a.b;
    const int C = (int)ali_cluster_size(), rank = (int)ali_cluster_rank();
    const int tid = threadIdx.x, gtid = rank * NT + tid, GT = C * NT;
    AliSourceRec &rec = b.rec[0];
    AliClusterCtl *ctl = a.ctl;
    __shared__ unsigned long long s_evals, s_fbs;
    __shared__ AliBandGrid s_grid;
    __shared__ int s_wsum[32];
    const int isz = rec.src_iz, isx = rec.src_ix;

    AliBandGrid g;
    g.nz = b.nz; g.nx = b.nx;
    g.T = b.Tt; g.st = b.st;
    g.t4x = (b.nx + 3) >> 2;
    g.dirty = nullptr; g.tiles_x = 0;
    g.dnx = b.m.dnx;
    g.mv = ali_band_view(1);
    const int cap = b.band_cap;
    double *val0 = b.stage, *val1 = val0 + cap;
    unsigned *ent0 = b.lists, *ent1 = ent0 + cap, *wrk0 = ent1 + cap, *wrk1 = wrk0 + cap;
    const bool peer_above = a.me > 0, peer_below = a.me < a.n_strips - 1;

    for (int q = tid; q < ALI_MT_WORDS; q += NT) s_sincos[q] = q < ALI_GL_SINCOSTAB_COUNT ? ali_gl_sincostab[q] : ali_gl_atan_cij[q - ALI_GL_SINCOSTAB_COUNT];
    if (tid == 0) { s_grid = g; s_evals = 0; s_fbs = 0; }
    if (gtid == 0) { ctl->evalmin[0] = ~0ull; ctl->evalmin[1] = ~0ull; ctl->basemin[0] = ~0ull; ctl->basemin[1] = ~0ull; }
    ali_cluster_sync();
    if (a.has_seq && !rec.overflow) {   // hand-over of the sequential phase's window: all of it lies in this strip
        const AliSeqResult w = rec.seq;
        const size_t woff = b.seq_cap;   // (subgrid 1: three levels, the main-grid window lives in the second buffer pair)
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
    if (a.has_seq && rec.overflow && gtid == 0) ctl->overflow = 1;
    ali_cluster_sync();
    if (gtid == 0) ctl->nwork[0] = ali_ldv(&ctl->count[0]);
    unsigned long long epoch = 0;
    bool alive = ali_strip_sync(a, epoch, gtid);

    int rounds = 0, max_band = 0;
    unsigned my_evals = 0, my_fbs = 0;
    int cur = 0;
    while (alive) {
        // (after a list overflow the counters may exceed the capacity: the strips leave the loop together in phase C of
        // this round, once the flag has travelled; until then stay inside the arrays)
        const int n = min(ali_ldv(&ctl->count[cur]), cap);
        const int nwork = min(ali_ldv(&ctl->nwork[cur]), cap);
        double *val = cur == 0 ? val0 : val1, *nval = cur == 0 ? val1 : val0;
        unsigned *ent = cur == 0 ? ent0 : ent1, *nent = cur == 0 ? ent1 : ent0;
        unsigned *wrk = cur == 0 ? wrk0 : wrk1, *nwrk = cur == 0 ? wrk1 : wrk0;
        rounds++;
        const int par = rounds & 1;
        if (n > max_band) max_band = n;
        const bool resort = b.resort_every > 0 && (rounds % b.resort_every) == 0;
        // ---- phase A: evaluate the local work list
        if (gtid == 0) { ctl->count[cur ^ 1] = 0; ctl->nwork[cur ^ 1] = 0; ctl->basemin[cur ^ 1] = ~0ull; ctl->evalmin[cur ^ 1] = ~0ull; }
        for (int q = gtid; q < ALI_DMAP_WORDS / 4; q += GT) reinterpret_cast<uint4 *>(ctl->dmap)[q] = make_uint4(0u, 0u, 0u, 0u);
        if (resort)
            for (int q = gtid; q < ALI_SORT_BINS; q += GT) ctl->bins[q] = 0;
        double lmin = 1e300;
        unsigned pe0 = 0, pe1 = 0;
        double pv0 = 0.0, pv1 = 0.0;
        int pmask = 0, it = 0;
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
                if (fb) ctl->force[par] = 1;
            }
        }
        for (int o = 16; o > 0; o >>= 1) lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        if ((tid & 31) == 0 && lmin < 1e300)
            atomicMin(&ctl->evalmin[cur], (unsigned long long)__double_as_longlong(lmin));
        if (!(alive = ali_strip_sync(a, epoch, gtid))) break;
        // ---- phase B: publish (also into the peer's halo rows), mark, tell the peer this strip's minimum and length
        if (gtid == 0) ctl->force[(rounds + 1) & 1] = 0;
        if (gtid < a.n_strips && gtid != a.me) {
            AliStripXchg *x = &a.px[gtid][a.me];
            x->evalmin[par] = ali_ldv(&ctl->evalmin[cur]);
            x->basemin[par] = ali_ldv(&ctl->basemin[cur]);
            x->count[par] = n;
            x->force[par] = ali_ldv(&ctl->force[par]);
            x->overflow = ali_ldv(&ctl->overflow);
        }
        auto publish = [&](unsigned e, double v) {
            const int z = ALI_PACK_Z(e), x = ALI_PACK_X(e);
            const size_t node = g.ti(z, x);
            g.T[node] = v;
            ali_dmap_mark(ctl->dmap, z, x);
            if (peer_above && z < a.zlo + 2) { a.nb[0].T[node] = v; ali_dmap_mark(a.nb[0].ctl->dmap, z, x); }
            if (peer_below && z >= a.zhi - 2) { a.nb[1].T[node] = v; ali_dmap_mark(a.nb[1].ctl->dmap, z, x); }
        };
        if (pmask & 1) publish(pe0, pv0);
        if (pmask & 2) publish(pe1, pv1);
        if (pmask & 4) ali_dmap_row(ctl->dmap, ALI_PACK_Z(pe0), ALI_PACK_X(pe0), 1u);
        if (pmask & 8) ali_dmap_row(ctl->dmap, ALI_PACK_Z(pe1), ALI_PACK_X(pe1), 1u);
        for (int q = q0 + 2 * GT; q < nwork; q += GT) {
            const int i = (int)wrk[q];
            const unsigned e = ent[i];
            const double v = val[i];
            if (v < 0.0) { val[i] = -v; publish(e, -v); }
        }
        if (!(alive = ali_strip_sync(a, epoch, gtid))) break;
        // ---- phase C: accept against the minimum of BOTH strips; neighbours across the boundary are claimed from,
        // and appended to, the peer
        int total = n, ovf = ali_ldv(&ctl->overflow), force = ali_ldv(&ctl->force[par]);
        unsigned long long tm = ali_ldv(&ctl->evalmin[cur]);
        { const unsigned long long o1 = ali_ldv(&ctl->basemin[cur]); tm = tm < o1 ? tm : o1; }
        for (int w = 0; w < a.n_strips; w++) {
            if (w == a.me) continue;
            const AliStripXchg *x = &a.xl[w];
            total += *(volatile const int *)&x->count[par];
            ovf |= *(volatile const int *)&x->overflow;
            force |= *(volatile const int *)&x->force[par];
            const unsigned long long o2 = *(volatile const unsigned long long *)&x->evalmin[par], o3 = *(volatile const unsigned long long *)&x->basemin[par];
            tm = tm < o2 ? tm : o2; tm = tm < o3 ? tm : o3;
        }
        if (total == 0 || ovf) { if (total == 0) rounds--; break; }   // (the round that finds every list empty did no work)
        const double thr = __longlong_as_double((long long)tm) + b.delta;
        double bmin = 1e300;
        for (int i = q0; i - (tid & 31) < n; i += GT) {
            int k = 0, kw = 0;
            unsigned out[4];
            double v = 0.0;
            if (i < n) {
                const unsigned e = ent[i];
                v = val[i];
                const int iz = ALI_PACK_Z(e), ix = ALI_PACK_X(e);
                if (!(v > thr)) {
                    const size_t me = g.ti(iz, ix);
                    g.st[me] = ALI_ST_ALIVE;
                    if (peer_above && iz < a.zlo + 2) a.nb[0].st[me] = ALI_ST_ALIVE;
                    if (peer_below && iz >= a.zhi - 2) a.nb[1].st[me] = ALI_ST_ALIVE;
                    unsigned long long *cw = (unsigned long long *)g.T;
                    const volatile unsigned long long *tw = (const volatile unsigned long long *)g.T;
#pragma unroll
                    for (int dir = 0; dir < 4; dir++) {
                        const int z = iz + (dir == 2 ? -1 : dir == 3 ? 1 : 0), x = ix + (dir == 0 ? -1 : dir == 1 ? 1 : 0);
                        if (z < 0 || z >= b.nz || x < 0 || x >= b.nx) continue;
                        const size_t nb = g.ti(z, x);
                        if (z >= a.zlo && z < a.zhi) {
                            if (tw[nb] == ALI_T_FAR_BITS && atomicCAS(cw + nb, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
                                out[k++] = ALI_PACK(z, x);
                        } else {
                            const AliStripPeer &pr = a.nb[z < a.zlo ? 0 : 1];
                            if (atomicCAS((unsigned long long *)pr.T + nb, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS) {
                                // the neighbour owns it: into ITS next list, as a new node (evaluated next round)
                                const int pos = atomicAdd(&pr.ctl->count[cur ^ 1], 1);
                                if (pos < cap) {
                                    const int wpos = atomicAdd(&pr.ctl->nwork[cur ^ 1], 1);
                                    unsigned *pl = pr.lists + (cur == 0 ? cap : 0);            // its next entry list
                                    unsigned *pw = pr.lists + 2 * cap + (cur == 0 ? cap : 0);  // its next work list
                                    double *pv = pr.stage + (cur == 0 ? cap : 0);
                                    pl[pos] = ALI_PACK(z, x);
                                    pv[pos] = 0.0;
                                    pw[wpos] = (unsigned)pos;
                                } else {
                                    pr.ctl->overflow = 2;
                                    ctl->overflow = 2;
                                }
                            }
                        }
                    }
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
        if (!(alive = ali_strip_sync(a, epoch, gtid))) break;
        if (resort && !ali_ldv(&ctl->overflow)) {
            // the strip's own list, sorted along the front as in the cluster kernel (local barriers only)
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
            int ovf = ali_ldv(&ctl->overflow);
            for (int w = 0; w < a.n_strips; w++)
                if (w != a.me) ovf |= *(volatile const int *)&a.xl[w].overflow;
            if (ovf) rec.overflow = ovf;
        }
    }
}

// Rows [zlo, zhi) of a strip's tiled field -> row-major, T / 1 (subgrid 1); never-reached nodes get 0.
__global__ void ali_finalize_rows_kernel(const double *Tt, const uint8_t *st, double *out, int zlo, int zhi, int nx)
{
    const size_t t4x = (size_t)((nx + 3) >> 2);
    const size_t total = (size_t)(zhi - zlo) * nx;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int z = zlo + (int)(i / nx), x = (int)(i % nx);
        const size_t node = (((size_t)(z >> 2) * t4x + (size_t)(x >> 2)) << 4) | (size_t)(((z & 3) << 2) | (x & 3));
        const double v = Tt[node];
        out[i] = (v >= 0.0) ? v : ((v != v && st[node] == ALI_ST_ALIVE) ? __longlong_as_double(ALI_T_NAN_VALUE_BITS) : 0.0);
    }
}
