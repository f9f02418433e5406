// ali_seq.cuh -- exact sequential replica of the reference's near-source treatment.
//
// The reference seeds every travel-time field with a heap-ordered march on nested
// refined grids around the source (ATR:1508-2052 for travel(), ATR:2155-2759 for
// travel_finer_grid()), then continues on the main grid.  The discrete solution near the
// source depends on the exact pop order of its (quirky) binary heap, so this part is
// replayed literally: one lane per source walks the same heap, the other lanes of the
// warp only help with the embarrassingly parallel fills (grid reset, analytic seed).
// The march is continued on the main grid until the front is `stop_r` nodes away from
// the source, where the band-synchronous kernel (ali_kernels.cu) takes over.
#pragma once
#include "ali_core.cuh"
#include <string.h>

#if defined(__CUDACC__)
#define ALI_SYNCWARP() __syncwarp()
#define ALI_SYNCCTA() __syncthreads()   /* the sequential kernel runs one source per CTA (one or more warps) */
#else
#define ALI_SYNCWARP() ((void)0)
#define ALI_SYNCCTA() ((void)0)
#endif

// Node state of one grid during the sequential march.  Status follows the reference:
// -1 far, 0 alive, >0 position in the heap (ATR:103).  Status is stored for a window
// [wz0, wz0+wnz) x [wx0, wx0+wnx) only; everything outside it is far by construction.
// Heap entry: the node as (iz << 16) | ix plus its index in the status window, so that the status
// updates of the heap operations need no address arithmetic.  One 8-byte word.
struct alignas(8) AliHeapEnt {
    unsigned zx, wi;
};
#define ALI_ENT_Z(e) ((int)((e).zx >> 16))
#define ALI_ENT_X(e) ((int)((e).zx & 0xffffu))

struct AliSeqGrid {
    int nz, nx;           // full extents of this grid (edge logic, absolute coordinates)
    double *t;            // T(z, x) = t[z * t_stride + x - toff]: a buffer of the status window's extent
    long long t_stride, toff;
    int32_t *st;
    int wz0, wx0, wnz, wnx;
    AliHeapEnt *heap;     // 1-indexed (ATR:118-119): packed node + its window index
    double *hkey;         // hkey[pos] == T of the node at heap position pos: saves the dependent T load per comparison
    int ndup;             // nodes with two heap entries (seed corners, ATR:1601-1612): both entries follow T
    unsigned dup0, dup1, dup2, dup3;   // packed (iz << 16 | ix); scalars so that a by-value copy of the grid stays in registers
    double *cv;           // evaluation cache, window-indexed like st (cooperative march only)
    uint8_t *cf;          // 1: cv holds the update of this node for the current state of its window
    int ntr, heap_cap;
    int overflow;
    AliMatView mv;
    double dnx;

    ALI_DEV bool in_win(int z, int x) const
    {
        return z >= wz0 && z < wz0 + wnz && x >= wx0 && x < wx0 + wnx;
    }
    ALI_DEV size_t widx(int z, int x) const { return (size_t)(z - wz0) * wnx + (x - wx0); }
    ALI_DEV int32_t &s(int z, int x) const { return st[(size_t)(z - wz0) * wnx + (x - wx0)]; }
    ALI_DEV int32_t status(int z, int x) const { return in_win(z, x) ? s(z, x) : -1; }
    // A node has an estimate (reference: nsts >= 0) exactly when its T is not NaN: the buffers are
    // reset to NaN and values are only ever written together with a status >= 0.  Outside the window
    // (also when level 1 of travel() asks with a wrong nnz, ATR:1645) there is no estimate.
    ALI_DEV bool avail(int z, int x) const { return in_win(z, x) && t[(long long)z * t_stride + x - toff] >= 0.0; }
    ALI_DEV bool alive(int z, int x) const { return in_win(z, x) && s(z, x) == 0; }
    ALI_DEV double &tref(int z, int x) const { return t[(long long)z * t_stride + x - toff]; }
    ALI_DEV double tt(int z, int x) const { return t[(long long)z * t_stride + x - toff]; }
};

// Python round(k / 2): round-half-to-even (ATR:123, 135, 160, 172).
ALI_DEV int ali_half_round(int k)
{
    int h = k >> 1;
    if (k & 1) return (h & 1) ? h + 1 : h;
    return h;
}

// Keys beyond the heap's last entry hold +infinity (ali_seq_clear fills the whole array; a slot that
// leaves the heap is reset), so the sift-down can load a level's four grandchild keys without
// comparing indices with ntr: an absent child never wins.  The key array has ALI_HKEY_SLOTS(heap_cap)
// entries so that the look-ahead of the last level stays inside it.
#define ALI_HKEY_SLOTS(cap) (2 * (size_t)(cap) + 8)
#if defined(__CUDA_ARCH__)
#define ALI_KEY_NONE (__longlong_as_double(0x7ff0000000000000LL))
#else
#define ALI_KEY_NONE (HUGE_VAL)
#endif

// Sift-up shared by addtree / updtree (ATR:122-137, 159-174): the reference swaps the entry at
// tpc with its parent while T(iz, ix) is smaller than the parent's time.  Here the parents move
// down into the hole and the entry is stored once at the end (same final arrangement as the chain
// of swaps).  The reference rewrites the status of the moving node at every swap; only the last
// value survives unless a node sits in the heap twice (ndup > 0: the seed corners, level 0 only),
// where the order of the status writes matters and is kept.  e0 / k0: the entry at tpc and its key.
ALI_DEV void ali_sift_up_core(AliSeqGrid &g, unsigned wi_me, double tv, AliHeapEnt e0, double k0, int tpc)
{
    int tpp = ali_half_round(tpc);
    bool moved = false;
    const bool dups = g.ndup > 0;
    while (tpp > 0) {
        const double pk = g.hkey[tpp];
        if (tv < pk) {
            const AliHeapEnt pe = g.heap[tpp];
            if (dups) g.st[wi_me] = tpp;
            g.st[pe.wi] = tpc;
            g.heap[tpc] = pe;
            g.hkey[tpc] = pk;
            tpc = tpp;
            tpp = ali_half_round(tpc);
            moved = true;
        } else {
            tpp = 0;
        }
    }
    if (moved) {
        g.heap[tpc] = e0; g.hkey[tpc] = k0;
        if (!dups) g.st[wi_me] = tpc;
    }
}

// updtree (ATR:141-175): called right after T(iz, ix) was rewritten, so the keys of the node's heap
// entries are refreshed first.  The reference compares through ttn[btg[...]]; keys that always equal
// T give the same order.
ALI_DEV void ali_updtree_w(AliSeqGrid &g, int iz, int ix, unsigned wi_me)
{
    const unsigned me = ((unsigned)iz << 16) | (unsigned)ix;
    const int tpc = g.st[wi_me];
    const double tv = g.t[wi_me];   // (T is window-indexed: t[z * t_stride + x - toff] == t[wi])
    const AliHeapEnt e0 = g.heap[tpc];
    double k0;
    if (e0.zx == me) { k0 = tv; g.hkey[tpc] = tv; }
    else k0 = g.hkey[tpc];   // only after the reference's own duplicate-entry mix-ups
    if (g.ndup > 0) {
        const bool is_dup = (g.dup0 == me) || (g.ndup > 1 && g.dup1 == me) || (g.ndup > 2 && g.dup2 == me) ||
                            (g.ndup > 3 && g.dup3 == me);
        if (is_dup)
            for (int q = 1; q <= g.ntr; q++)
                if (g.heap[q].zx == me) g.hkey[q] = tv;
    }
    ali_sift_up_core(g, wi_me, tv, e0, k0, tpc);
}

ALI_DEV void ali_updtree(AliSeqGrid &g, int iz, int ix) { ali_updtree_w(g, iz, ix, (unsigned)g.widx(iz, ix)); }

// addtree (ATR:106-138).
ALI_DEV void ali_addtree_w(AliSeqGrid &g, int iz, int ix, unsigned wi_me)
{
    if (g.ntr + 2 >= g.heap_cap) { g.overflow = 1; return; }
    const unsigned me = ((unsigned)iz << 16) | (unsigned)ix;
    const double tv = g.t[wi_me];
    if (g.st[wi_me] > 0) {   // second heap entry for a node (the reference pushes the seed corners twice)
        if (g.ndup == 0) g.dup0 = me; else if (g.ndup == 1) g.dup1 = me; else if (g.ndup == 2) g.dup2 = me;
        else if (g.ndup == 3) g.dup3 = me; else { g.overflow = 1; return; }
        g.ndup++;
        for (int q = 1; q <= g.ntr; q++)
            if (g.heap[q].zx == me) g.hkey[q] = tv;
    }
    g.ntr += 1;
    AliHeapEnt e0;
    e0.zx = me; e0.wi = wi_me;
    g.st[wi_me] = g.ntr;
    g.heap[g.ntr] = e0;
    g.hkey[g.ntr] = tv;
    ali_sift_up_core(g, wi_me, tv, e0, tv, g.ntr);
}

ALI_DEV void ali_addtree(AliSeqGrid &g, int iz, int ix) { ali_addtree_w(g, iz, ix, (unsigned)g.widx(iz, ix)); }

// downtree (ATR:178-237): the last entry replaces the root and sinks; children move up into the
// hole, the entry is stored once at the end.  The reference's two cases -- both children present
// (ATR:196-221), only one (ATR:222-236) -- are one loop here: the absent sibling's key is +infinity.
ALI_DEV void ali_downtree(AliSeqGrid &g)
{
    int ntr = g.ntr;
    if (ntr == 1) { g.hkey[1] = ALI_KEY_NONE; g.ntr = 0; return; }
    const AliHeapEnt le = g.heap[ntr];
    const double kv = g.hkey[ntr];
    g.hkey[ntr] = ALI_KEY_NONE;
    g.st[le.wi] = 1;
    g.heap[1] = le;
    g.hkey[1] = kv;
    ntr -= 1;
    int tpp = 1, tpc = 2;
    bool moved = false;
    const bool dups = g.ndup > 0;
    // The keys of a level's pair are loaded one level ahead (for both candidates), so that the chain
    // per level is compare + select instead of load + compare.  Only the hole position is written
    // on the way down, never a position below it: what was loaded ahead stays valid.
    double ka = g.hkey[2], kb = g.hkey[3];
    while (tpc <= ntr) {
        const int ga = 2 * tpc;
        const double gaa = g.hkey[ga], gab = g.hkey[ga + 1], gba = g.hkey[ga + 2], gbb = g.hkey[ga + 3];
        const bool second = ka > kb;
        const double rd1 = second ? kb : ka;
        tpc += second ? 1 : 0;
        if (rd1 < kv) {
            const AliHeapEnt ce = g.heap[tpc];
            if (dups) g.st[le.wi] = tpc;
            g.st[ce.wi] = tpp;
            g.heap[tpp] = ce;
            g.hkey[tpp] = rd1;
            tpp = tpc;
            tpc = 2 * tpp;
            ka = second ? gba : gaa;
            kb = second ? gbb : gab;
            moved = true;
        } else {
            break;
        }
    }
    if (moved) {
        g.heap[tpp] = le; g.hkey[tpp] = kv;
        if (!dups) g.st[le.wi] = tpp;
    }
    g.ntr = ntr;
}

// One evaluation of a node of a sequential grid (ali_eval_node with a branch-free gather for nodes
// two or more inside the status window, where every window node is inside the grid as well).
ALI_DEV double ali_seq_eval(const AliModel &m, const AliSeqGrid &g, int iz, int ix, int nnz_logic, int *used_fallback)
{
    if (nnz_logic == g.nz && iz - 2 >= g.wz0 && iz + 2 < g.wz0 + g.wnz && ix - 2 >= g.wx0 && ix + 2 < g.wx0 + g.wnx) {
        AliMat mat;
        AliWindow w;
        ali_fetch_mat(m, g.mv, iz, ix, mat);
        const double *tp = g.t + ((long long)iz * g.t_stride + ix - g.toff);
        const long long st = g.t_stride;
        unsigned av = 0;
#pragma unroll
        for (int k = 0; k < 12; k++) w.t[k] = tp[ALI_W_DZ(k) * st + ALI_W_DX(k)];
#pragma unroll
        for (int k = 0; k < 12; k++) {
            if (w.t[k] >= 0.0) av |= 1u << k;
            else w.t[k] = 0.0;   // as ali_gather leaves it
        }
        w.avail = av;
        double v = ali_update_window(m, mat, w, iz, ix, g.nz, g.nx, g.dnx, nullptr);
        if (v == -1.0) {
            v = ali_fouds18(m, mat, g, iz, ix, g.dnx, g.dnx, g.nx, g.nz);
            if (used_fallback) *used_fallback = 1;
        }
        return v;
    }
    return ali_eval_node(m, g.mv, g, iz, ix, nnz_logic, g.nx, g.nz, g.nx, g.dnx, used_fallback);
}

struct AliSeqCounters {
    long long pops, evals, fallbacks;
    long long cyc_heap, cyc_eval;   // device builds: SM cycles in heap operations / evaluations
    long long steps, computed;      // cooperative march: warp-wide evaluation steps, evaluations executed in them
    long long cyc_total, cyc_fill, cyc_start;   // device builds: whole source / fills + seeds / perimeter push + hand-offs
};

#if defined(__CUDA_ARCH__)
#define ALI_CLOCK() clock64()
#else
#define ALI_CLOCK() 0ll
#endif

// One heap-ordered march (ATR:1621-1674 and its copies ATR:1787-1844, 1937-1993,
// 2055-2102, 2292-2346, 2460-2504, 2775-2817).
//   max_dist >= 0 : stop once a popped node has an out-of-grid neighbour exactly
//                   max_dist + 1 from the centre (the refined box was not clipped there);
//   nnz_bug       : level 1 of travel() passes nnx as nnz for x-direction updates of
//                   close nodes (ATR:1645);
//   stop_r >= 0   : (main grid only) stop after popping a node whose Chebyshev distance
//                   from the centre reaches stop_r -- hand-over point to the band march.
// Returns why it stopped: 0 heap empty, 1 front left the refined box, 2 stop_r reached,
// 3 scratch exhausted / window too small.
#define ALI_SEQ_EMPTY 0
#define ALI_SEQ_BOX 1
#define ALI_SEQ_HANDOVER 2
#define ALI_SEQ_LIMIT 3
ALI_DEV int ali_seq_march(AliSeqGrid &g, const AliModel &m, int cx, int cz, int max_dist, int nnz_bug, int stop_r,
                          AliSeqCounters &cnt)
{
    bool finished = false;
    int why = ALI_SEQ_EMPTY;
    const int nnx = g.nx, nnz = g.nz;
    while (g.ntr > 0 && !finished) {
        const int ix = ALI_ENT_X(g.heap[1]), iz = ALI_ENT_Z(g.heap[1]);
        long long c0 = ALI_CLOCK();
        g.s(iz, ix) = 0;
        ali_downtree(g);
        cnt.pops++;
        cnt.cyc_heap += ALI_CLOCK() - c0;
        for (int s = 0; s < 4; s++) {
            int z = iz, x = ix;
            if (s == 0) x = ix - 1; else if (s == 1) x = ix + 1; else if (s == 2) z = iz - 1; else z = iz + 1;
            bool inside = (s < 2) ? (0 <= x && x <= nnx - 1) : (0 <= z && z <= nnz - 1);
            if (inside) {
                if (!g.in_win(z, x)) { finished = true; why = ALI_SEQ_LIMIT; continue; } // window too small: hand over early
                int32_t stv = g.s(z, x);
                if (stv != 0) {
                    int nnz_l = (nnz_bug && s < 2 && stv > 0) ? nnx : nnz;
                    int fb = 0;
                    long long c1 = ALI_CLOCK();
                    double v = ali_seq_eval(m, g, z, x, nnz_l, &fb);
                    long long c2 = ALI_CLOCK();
                    cnt.evals++;
                    cnt.fallbacks += fb;
                    g.tref(z, x) = v;
                    if (stv == -1) ali_addtree(g, z, x);
                    else ali_updtree(g, z, x);
                    cnt.cyc_eval += c2 - c1;
                    cnt.cyc_heap += ALI_CLOCK() - c2;
                }
            } else if (max_dist >= 0) {
                int d = (s < 2) ? (cx - x) : (cz - z);
                if (d < 0) d = -d;
                if (d == max_dist + 1) { finished = true; why = ALI_SEQ_BOX; }
            }
        }
        if (stop_r >= 0 && why != ALI_SEQ_BOX) {
            int dz = iz - cz, dx = ix - cx;
            if (dz < 0) dz = -dz;
            if (dx < 0) dx = -dx;
            if ((dz > dx ? dz : dx) >= stop_r) { finished = true; if (why == ALI_SEQ_EMPTY) why = ALI_SEQ_HANDOVER; }
        }
        if (g.overflow) { finished = true; why = ALI_SEQ_LIMIT; }
    }
    if (!finished) why = ALI_SEQ_EMPTY;
    return why;
}

// ---------------------------------------------------------------------------
// Cooperative form of the same march: identical result, shorter critical path.
//
// The heap order is inherently sequential, but the expensive part of a pop -- the ALI update
// of the popped node's neighbours -- is a pure function of the neighbour's 12-node window
// (values + availability), its material and the grid edges.  So its result can be computed
// ahead of time and kept until a node of that window changes:
//   * cv/cf cache the update of a node for the current state of its window.  Whenever a node
//     receives a new value (or its first one) the flags of its 12 window neighbours are
//     cleared.  A re-evaluation the reference performs on an unchanged window (97 % of its
//     re-evaluations on the weld) is answered from the cache.
//   * Lane 0 walks the reference's loop.  When it needs an update that is not cached it stops,
//     and the whole warp evaluates in one step: lane 0 the missing node, the other lanes the
//     not-yet-cached neighbours of the nodes in the first heap positions (the next pops).
// Results of the FD fallback are never cached (it also reads alive flags), nor are the
// evaluations level 1 of travel() makes with the wrong nnz (ATR:1645).
// ---------------------------------------------------------------------------
struct AliCoopState {
    int have_p, iz, ix, s;       // popped node whose neighbours are being visited, next neighbour
    unsigned wi;                 // its window index
    int miss, mz, mx, mnnz;      // evaluation lane 0 is waiting for
    int miss_ready, miss_fb;
    double miss_v;
    int finished, why;
};

ALI_DEV void ali_seq_invalidate(const AliSeqGrid &g, int z, int x, unsigned wi)
{
    if (z - 2 >= g.wz0 && z + 2 < g.wz0 + g.wnz && x - 2 >= g.wx0 && x + 2 < g.wx0 + g.wnx) {
        uint8_t *f = g.cf + wi;
        const int w = g.wnx;
        f[-2 * w] = 0; f[-w - 1] = 0; f[-w] = 0; f[-w + 1] = 0;
        f[-2] = 0; f[-1] = 0; f[1] = 0; f[2] = 0;
        f[w - 1] = 0; f[w] = 0; f[w + 1] = 0; f[2 * w] = 0;
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) {
            const int zz = z + ALI_W_DZ(k), xx = x + ALI_W_DX(k);
            if (g.in_win(zz, xx)) g.cf[g.widx(zz, xx)] = 0;
        }
    }
}

// Speculation looks at the neighbours of the first ALI_COOP_HEAP_POSITIONS heap entries (the
// likely next pops).  wanted(hp): 4-bit mask of the directions (x-1, x+1, z-1, z+1) whose node is
// not alive and has no valid cached update.
// (Ranking the candidates of the first 32 entries with ballots / __fns was measured too: 3 % fewer
// steps, each 35 % dearer.)
#define ALI_COOP_HEAP_POSITIONS 8
ALI_DEV unsigned ali_coop_wanted(const AliSeqGrid &g, int hp, int ntr)
{
    unsigned mask = 0;
    if (hp <= ntr && hp <= ALI_COOP_HEAP_POSITIONS) {
        const AliHeapEnt he = g.heap[hp];
        const int pz = ALI_ENT_Z(he), px = ALI_ENT_X(he);
#pragma unroll
        for (int dir = 0; dir < 4; dir++) {
            const int cz = pz + (dir == 2 ? -1 : dir == 3 ? 1 : 0), cx = px + (dir == 0 ? -1 : dir == 1 ? 1 : 0);
            const unsigned wi = he.wi + (unsigned)(dir == 0 ? -1 : dir == 1 ? 1 : dir == 2 ? -g.wnx : g.wnx);
            if (cz >= 0 && cz < g.nz && cx >= 0 && cx < g.nx && g.in_win(cz, cx) && g.st[wi] != 0 && g.cf[wi] == 0)
                mask |= 1u << dir;
        }
    }
    return mask;
}

// True if (cz, cx) is worth evaluating ahead: inside, not alive, no valid cached update.
ALI_DEV bool ali_coop_wanted_node(const AliSeqGrid &g, int cz, int cx)
{
    return cz >= 0 && cz < g.nz && cx >= 0 && cx < g.nx && g.in_win(cz, cx) && g.s(cz, cx) != 0 &&
           g.cf[g.widx(cz, cx)] == 0;
}

// One lane's evaluation in a step: lane 0 serves the pending miss (serve), the others evaluate
// the node they were assigned, if any, into the cache.
ALI_DEV int ali_coop_step(const AliSeqGrid &g, const AliModel &m, AliCoopState &cs, bool serve, bool has_cand, int cz,
                          int cx)
{
    const int nnz = g.nz;
    int cnnz = nnz;
    if (serve) {
        cz = cs.mz; cx = cs.mx; cnnz = cs.mnnz;
    } else if (!has_cand) {
        return 0;
    }
    int fb = 0;
    const double v = ali_seq_eval(m, g, cz, cx, cnnz, &fb);
    if (serve) {
        cs.miss_v = v; cs.miss_fb = fb; cs.miss_ready = 1; cs.miss = 0;
    } else if (!fb) {
        g.cv[g.widx(cz, cx)] = v;
        g.cf[g.widx(cz, cx)] = 1;
    }
    return 1;
}

// Lane 0: the reference's loop (same order of heap operations as ali_seq_march) until it ends
// (returns 1) or needs an update that is not cached (returns 0 with cs.miss set).
ALI_DEV int ali_coop_advance(AliSeqGrid &g, AliCoopState &cs, int cx, int cz, int max_dist, int nnz_bug, int stop_r,
                             AliSeqCounters &cnt)
{
    const int nnx = g.nx, nnz = g.nz;
    for (;;) {
        if (!cs.have_p) {
            if (g.ntr <= 0 || cs.finished) return 1;
            const AliHeapEnt top = g.heap[1];
            cs.ix = ALI_ENT_X(top); cs.iz = ALI_ENT_Z(top); cs.wi = top.wi;
            g.st[top.wi] = 0;
            ali_downtree(g);
            cnt.pops++;
            cs.s = 0;
            cs.have_p = 1;
        }
        const int iz = cs.iz, ix = cs.ix;
        for (; cs.s < 4; cs.s++) {
            const int s = cs.s;
            int z = iz, x = ix;
            if (s == 0) x = ix - 1; else if (s == 1) x = ix + 1; else if (s == 2) z = iz - 1; else z = iz + 1;
            const bool inside = (s < 2) ? (0 <= x && x <= nnx - 1) : (0 <= z && z <= nnz - 1);
            if (inside) {
                // A neighbour of a popped node is always inside the status window: level grids store every node,
                // and on the main grid the march stops at the first pop stop_r nodes from the source while the
                // window reaches stop_r + 4 (ali_src_main_geometry).  The host replay keeps the check.
#if !defined(__CUDA_ARCH__)
                if (!g.in_win(z, x)) { cs.finished = 1; cs.why = ALI_SEQ_LIMIT; continue; }
#endif
                const unsigned wi = cs.wi + (unsigned)(s == 0 ? -1 : s == 1 ? 1 : s == 2 ? -g.wnx : g.wnx);
                // the neighbour's state in one go (independent loads; T is stored window-indexed like the rest:
                // t[z * t_stride + x - toff] == t[wi] for every grid the march runs on)
                const int32_t stv = g.st[wi];
                const uint8_t cfv = g.cf[wi];
                const double cvv = g.cv[wi];
                const double told = g.t[wi];
                if (stv != 0) {
                    const int nnz_l = (nnz_bug && s < 2 && stv > 0) ? nnx : nnz;
                    double v;
                    int fb = 0;
                    if (cs.miss_ready) { v = cs.miss_v; fb = cs.miss_fb; cs.miss_ready = 0; }
                    else if (nnz_l == nnz && cfv) v = cvv;
                    else { cs.miss = 1; cs.mz = z; cs.mx = x; cs.mnnz = nnz_l; return 0; }
                    cnt.evals++;
                    cnt.fallbacks += fb;
                    const bool changed = (stv == -1) || !(told == v);
                    g.t[wi] = v;
                    if (changed) ali_seq_invalidate(g, z, x, wi);
                    if (nnz_l == nnz && !fb) { g.cv[wi] = v; g.cf[wi] = 1; }
                    else g.cf[wi] = 0;
                    if (stv == -1) ali_addtree_w(g, z, x, wi);
                    else ali_updtree_w(g, z, x, wi);
                }
            } else if (max_dist >= 0) {
                int d = (s < 2) ? (cx - x) : (cz - z);
                if (d < 0) d = -d;
                if (d == max_dist + 1) { cs.finished = 1; cs.why = ALI_SEQ_BOX; }
            }
        }
        cs.have_p = 0;
        if (stop_r >= 0 && cs.why != ALI_SEQ_BOX) {
            int dz = iz - cz, dx = ix - cx;
            if (dz < 0) dz = -dz;
            if (dx < 0) dx = -dx;
            if ((dz > dx ? dz : dx) >= stop_r) { cs.finished = 1; if (cs.why == ALI_SEQ_EMPTY) cs.why = ALI_SEQ_HANDOVER; }
        }
        if (g.overflow) { cs.finished = 1; cs.why = ALI_SEQ_LIMIT; }
    }
}

// Same contract as ali_seq_march; every lane of the warp calls it (the host replay passes
// nlanes and plays the lanes one after the other).  Only lane 0's g / cnt are meaningful.
ALI_DEV int ali_seq_march_coop(AliSeqGrid &g_mem, const AliModel &m, int cx, int cz, int max_dist, int nnz_bug,
                               int stop_r, AliSeqCounters &cnt_mem, int lane, int nlanes)
{
    // the caller's grid sits in an array of the source state (local memory): work on register copies
    AliSeqGrid g = g_mem;
    AliSeqCounters cnt = cnt_mem;
    AliCoopState cs;
    cs.have_p = 0; cs.iz = cs.ix = cs.s = 0;
    cs.miss = 0; cs.mz = cs.mx = cs.mnnz = 0;
    cs.miss_ready = 0; cs.miss_fb = 0; cs.miss_v = 0.0;
    cs.finished = 0; cs.why = ALI_SEQ_EMPTY;
#if defined(__CUDA_ARCH__)
    if (nlanes > 32) {
        // CTA of several warps: warp 0 is the walker (lane 0) and, when the walk stops at a miss, finds
        // the step's candidates exactly as the one-warp form does; the candidates are then evaluated ONE
        // PER WARP (a lone lane runs the update at single-thread latency, where 32 different nodes in
        // one warp serialise their divergent paths), between two CTA barriers.
        __shared__ int s_nc, s_done, s_serve, s_miss_fb;
        __shared__ int s_cz[32], s_cx[32], s_cnnz[32];
        __shared__ double s_miss_v;
        const int warp = lane >> 5, wl = lane & 31, nw = nlanes >> 5;
        if (lane == 0) s_done = 0;
        for (;;) {
            if (warp == 0) {
                const int ntr = __shfl_sync(0xffffffffu, g.ntr, 0);
                const int serve = __shfl_sync(0xffffffffu, cs.miss, 0);
                const int pz = __shfl_sync(0xffffffffu, cs.iz, 0), px = __shfl_sync(0xffffffffu, cs.ix, 0);
                const int ps = __shfl_sync(0xffffffffu, cs.s, 0);
                const int mz = __shfl_sync(0xffffffffu, cs.mz, 0), mx = __shfl_sync(0xffffffffu, cs.mx, 0);
                const int mnnz = __shfl_sync(0xffffffffu, cs.mnnz, 0);
                bool has = false;
                int cz_c = 0, cx_c = 0, cnnz = g.nz;
                if (serve && wl == 0) { has = true; cz_c = mz; cx_c = mx; cnnz = mnnz; }
                if (serve && wl >= 1 && wl <= 3) {
                    const int d = ps + wl;
                    if (d < 4) {
                        cz_c = pz + (d == 2 ? -1 : d == 3 ? 1 : 0); cx_c = px + (d == 0 ? -1 : d == 1 ? 1 : 0);
                        has = ali_coop_wanted_node(g, cz_c, cx_c);
                    }
                }
                {
                    const int slot = serve ? wl - 4 : wl;
                    if (slot >= 0 && !has) {
                        const int hp = 1 + (slot >> 2), dir = slot & 3;
                        if (hp <= ntr) {
                            const AliHeapEnt he = g.heap[hp];
                            cz_c = ALI_ENT_Z(he) + (dir == 2 ? -1 : dir == 3 ? 1 : 0);
                            cx_c = ALI_ENT_X(he) + (dir == 0 ? -1 : dir == 1 ? 1 : 0);
                            has = ali_coop_wanted_node(g, cz_c, cx_c);
                        }
                    }
                }
                // two lanes may name the same node (neighbouring heap entries): keep the first
                const unsigned long long key = ((unsigned long long)(unsigned)cz_c << 32) | (unsigned)cx_c;
                const unsigned same = __match_any_sync(0xffffffffu, has ? key : (0xffffffff00000000ull | (unsigned)wl));
                if (has && (same & ((1u << wl) - 1u))) has = false;
                const unsigned mask = __ballot_sync(0xffffffffu, has);
                if (has) {
                    const int k = __popc(mask & ((1u << wl) - 1u));
                    s_cz[k] = cz_c; s_cx[k] = cx_c; s_cnnz[k] = cnnz;
                }
                if (wl == 0) { s_nc = __popc(mask); s_serve = serve; }
            }
            __syncthreads();
            if (s_done) break;
            {
                const int nc = s_nc, serve = s_serve;
                const int k = warp + wl * nw;   // candidate k -> warp k % nw, lane k / nw
                if (k < nc) {
                    int fb = 0;
                    const int cz_c = s_cz[k], cx_c = s_cx[k];
                    const double v = ali_seq_eval(m, g, cz_c, cx_c, s_cnnz[k], &fb);
                    if (k == 0 && serve) { s_miss_v = v; s_miss_fb = fb; }
                    else if (!fb) { g.cv[g.widx(cz_c, cx_c)] = v; g.cf[g.widx(cz_c, cx_c)] = 1; }
                }
            }
            __syncthreads();
            if (lane == 0) {
                if (cs.miss) { cs.miss_v = s_miss_v; cs.miss_fb = s_miss_fb; cs.miss_ready = 1; cs.miss = 0; }
                cnt.steps++;
                cnt.computed += s_nc;
                s_done = ali_coop_advance(g, cs, cx, cz, max_dist, nnz_bug, stop_r, cnt);
            }
        }
        g_mem.ntr = g.ntr; g_mem.overflow = g.overflow; g_mem.ndup = g.ndup;
        g_mem.dup0 = g.dup0; g_mem.dup1 = g.dup1; g_mem.dup2 = g.dup2; g_mem.dup3 = g.dup3;
        cnt_mem = cnt;
        if (!cs.finished) cs.why = ALI_SEQ_EMPTY;
        __syncthreads();   // (s_done is re-armed by the next march)
        return cs.why;
    }
#endif
    for (;;) {
        int done = 0;
#if defined(__CUDA_ARCH__)
        const long long c0 = ALI_CLOCK();
        const int ntr = __shfl_sync(0xffffffffu, g.ntr, 0);
        const int serve = __shfl_sync(0xffffffffu, cs.miss, 0);
        // a step that serves a miss reserves lanes 1..3 for the neighbours of the popped node that
        // are still to be visited (the popped node has left the heap)
        const int pz = __shfl_sync(0xffffffffu, cs.iz, 0), px = __shfl_sync(0xffffffffu, cs.ix, 0);
        const int ps = __shfl_sync(0xffffffffu, cs.s, 0);
        bool has = false;
        int cz_c = 0, cx_c = 0;
        if (serve && lane >= 1 && lane <= 3) {
            const int d = ps + lane;
            if (d < 4) {
                cz_c = pz + (d == 2 ? -1 : d == 3 ? 1 : 0); cx_c = px + (d == 0 ? -1 : d == 1 ? 1 : 0);
                has = ali_coop_wanted_node(g, cz_c, cx_c);
            }
        }
        // fixed assignment: lane -> (heap position, direction) of the first 7-8 heap entries
        {
            const int slot = serve ? lane - 4 : lane;
            if (slot >= 0 && !has) {
                const int hp = 1 + (slot >> 2), dir = slot & 3;
                if (hp <= ntr) {
                    const AliHeapEnt he = g.heap[hp];
                    cz_c = ALI_ENT_Z(he) + (dir == 2 ? -1 : dir == 3 ? 1 : 0);
                    cx_c = ALI_ENT_X(he) + (dir == 0 ? -1 : dir == 1 ? 1 : 0);
                    has = ali_coop_wanted_node(g, cz_c, cx_c);
                }
            }
        }
        const int did = ali_coop_step(g, m, cs, serve && lane == 0, has, cz_c, cx_c);
        const unsigned mask = __ballot_sync(0xffffffffu, did);
        if (lane == 0) {
            const long long c1 = ALI_CLOCK();
            cnt.steps++;
            cnt.computed += __popc(mask);
            done = ali_coop_advance(g, cs, cx, cz, max_dist, nnz_bug, stop_r, cnt);
            cnt.cyc_eval += c1 - c0;
            cnt.cyc_heap += ALI_CLOCK() - c1;
        }
        __syncwarp();
        done = __shfl_sync(0xffffffffu, done, 0);
#else
        (void)lane;
        cnt.steps++;
        {
            // the lanes pick their nodes from the same snapshot: collect first, evaluate after
            int cz_c[32], cx_c[32], nc = 0, used = 0;
            const bool serve = cs.miss != 0;
            if (serve) {
                used = 4;
                for (int d = cs.s + 1; d < 4; d++) {
                    const int z = cs.iz + (d == 2 ? -1 : d == 3 ? 1 : 0), x = cs.ix + (d == 0 ? -1 : d == 1 ? 1 : 0);
                    if (ali_coop_wanted_node(g, z, x)) { cz_c[nc] = z; cx_c[nc] = x; nc++; }
                }
            }
            int taken = 0;
            for (int hp = 1; hp <= ALI_COOP_HEAP_POSITIONS - (serve ? 1 : 0) && used + taken < nlanes; hp++) {
                const unsigned wm = ali_coop_wanted(g, hp, g.ntr);
                for (int dir = 0; dir < 4 && used + taken < nlanes; dir++)
                    if (wm & (1u << dir)) {
                        cz_c[nc] = ALI_ENT_Z(g.heap[hp]) + (dir == 2 ? -1 : dir == 3 ? 1 : 0);
                        cx_c[nc] = ALI_ENT_X(g.heap[hp]) + (dir == 0 ? -1 : dir == 1 ? 1 : 0);
                        nc++; taken++;
                    }
            }
            if (serve) cnt.computed += ali_coop_step(g, m, cs, true, false, 0, 0);
            for (int q = 0; q < nc; q++) cnt.computed += ali_coop_step(g, m, cs, false, true, cz_c[q], cx_c[q]);
        }
        done = ali_coop_advance(g, cs, cx, cz, max_dist, nnz_bug, stop_r, cnt);
#endif
        if (done) break;
    }
    g_mem.ntr = g.ntr; g_mem.overflow = g.overflow; g_mem.ndup = g.ndup;
    g_mem.dup0 = g.dup0; g_mem.dup1 = g.dup1; g_mem.dup2 = g.dup2; g_mem.dup3 = g.dup3;
    cnt_mem = cnt;
    if (!cs.finished) cs.why = ALI_SEQ_EMPTY;
    return cs.why;
}

// Resets a level grid: T = 0, status = far.  Cooperative over `nlanes` lanes.
ALI_DEV void ali_seq_clear(AliSeqGrid &g, bool clear_t, int lane, int nlanes)
{
    size_t n = (size_t)g.wnz * g.wnx;
    for (size_t i = lane; i < n; i += nlanes) g.st[i] = -1;
    if (g.cf)
        for (size_t i = lane; i < n; i += nlanes) g.cf[i] = 0;
    if (clear_t)
        for (size_t i = lane; i < n; i += nlanes) {   // level grids only (window == grid); NaN = no estimate
            const unsigned long long nanbits = 0xFFFFFFFFFFFFFFFFull;
            memcpy(&g.t[i], &nanbits, sizeof(double));
        }
    if (g.hkey) {
        const size_t nk = ALI_HKEY_SLOTS(g.heap_cap);
        for (size_t i = lane; i < nk; i += nlanes) g.hkey[i] = ALI_KEY_NONE;
    }
    g.ntr = 0;
    g.overflow = 0;
    g.ndup = 0;
}

// Analytic straight-ray seed of the source's own coarse cell (ATR:1546-1590 with
// sign = -1; ATR:2223-2267 with sign = +1).  Cooperative; followed by ali_seq_seed_push.
ALI_DEV void ali_seq_seed(AliSeqGrid &g1, const AliModel &m, const AliMatView &src_view, int isz, int isx, int cz1,
                          int cx1, int side1, double sign, int lane, int nlanes)
{
    AliMat mat;
    ali_fetch_mat(m, src_view, isz, isx, mat);
    const int w = 2 * side1 + 1;
    for (int q = lane; q < w * w; q += nlanes) {
        int i = q / w - side1, j = q % w - side1;
        if (!(0 <= cz1 + i && cz1 + i <= g1.nz - 1)) continue;
        if (!(0 <= cx1 + j && cx1 + j <= g1.nx - 1)) continue;
        double angle;
        if (j == 0) angle = 90.0;
        else angle = ALI_RAD2DEG * ALI_ATAN((double)i / (double)j);
        double eff = ali_pymod(mat.veln + sign * angle, 180.0);
        double vel;
        if (mat.velpn != 0) vel = ali_table_vel(m.group_tab, m.ncol, eff, mat.velpn, mat.vel_map);
        else vel = ali_christoffel_group(eff, mat.s, mat.vel_map);
        double length = g1.dnx * sqrt((double)(i * i + j * j));
        g1.tref(cz1 + i, cx1 + j) = length / vel;
        g1.s(cz1 + i, cx1 + j) = 0;
    }
}

// Pushes the perimeter of the seed square on the heap in the reference's order
// (ATR:1601-1612); corners are pushed twice, as there.  Single lane.
ALI_DEV void ali_seq_seed_push(AliSeqGrid &g1, int cz1, int cx1, int side1)
{
    int xa = cx1 - side1 > 0 ? cx1 - side1 : 0;
    int xb = cx1 + side1 < g1.nx - 1 ? cx1 + side1 : g1.nx - 1;
    int za = cz1 - side1 > 0 ? cz1 - side1 : 0;
    int zb = cz1 + side1 < g1.nz - 1 ? cz1 + side1 : g1.nz - 1;
    if (cz1 - side1 >= 0) for (int i = xa; i <= xb; i++) ali_addtree(g1, cz1 - side1, i);
    if (cz1 + side1 <= g1.nz - 1) for (int i = xa; i <= xb; i++) ali_addtree(g1, cz1 + side1, i);
    if (cx1 - side1 >= 0) for (int i = za; i <= zb; i++) ali_addtree(g1, i, cx1 - side1);
    if (cx1 + side1 <= g1.nx - 1) for (int i = za; i <= zb; i++) ali_addtree(g1, i, cx1 + side1);
}

// Every-third-node injection of grid a into the 3x coarser grid b (ATR:1719-1753,
// 1887-1921, 2006-2040, 2391-2425, 2725-2759).  Single lane: the heap insertion order
// is part of the result.
ALI_DEV void ali_seq_handoff(const AliSeqGrid &a_mem, int cza, int cxa, AliSeqGrid &b_mem, int czb, int cxb)
{
    const AliSeqGrid a = a_mem;   // register copies (see ali_seq_march_coop)
    AliSeqGrid b = b_mem;
    for (int i = 0; i <= a.nz - 1; i += 3) {
        for (int j = 0; j <= a.nx - 1; j += 3) {
            int pz = czb + (i - cza) / 3;
            int px = cxb + (j - cxa) / 3;
            b.tref(pz, px) = a.tt(i, j);
            int32_t stv = a.s(i, j);
            if (stv == 0) {
                bool outer = false;
                b.s(pz, px) = 0;
                if (i - 3 >= 0) { if (a.s(i - 3, j) == -1) outer = true; } else outer = true;
                if (i + 3 <= a.nz - 1) { if (a.s(i + 3, j) == -1) outer = true; } else outer = true;
                if (j - 3 >= 0) { if (a.s(i, j - 3) == -1) outer = true; } else outer = true;
                if (j + 3 <= a.nx - 1) { if (a.s(i, j + 3) == -1) outer = true; } else outer = true;
                if (outer) ali_addtree(b, pz, px);
            } else if (stv > 0) {
                ali_addtree(b, pz, px);
            }
        }
    }
    b_mem.ntr = b.ntr; b_mem.overflow = b.overflow; b_mem.ndup = b.ndup;
    b_mem.dup0 = b.dup0; b_mem.dup1 = b.dup1; b_mem.dup2 = b.dup2; b_mem.dup3 = b.dup3;
}

// ---------------------------------------------------------------------------
// Plan of one source: the nested levels and the main-grid window.
// ---------------------------------------------------------------------------
#define ALI_MAX_LEVELS 3
struct AliSourcePlan {
    int fine;              // 0: travel() (ATR:1463), 1: travel_finer_grid() (ATR:2120)
    int sg;                // subgrid_size (1 for travel)
    int nz, nx;            // main grid extents (fine extents when fine)
    int isz, isx;          // source node on the main grid
    int nlev;
    int scale[ALI_MAX_LEVELS], size[ALI_MAX_LEVELS];
    int stop_r;            // hand-over radius on the main grid
};

ALI_DEV int ali_imax(int a, int b) { return a > b ? a : b; }
ALI_DEV int ali_imin(int a, int b) { return a < b ? a : b; }

ALI_HD void ali_make_plan(AliSourcePlan &p, const AliModel &m, int src_iz, int src_ix, int sg, int handover_margin)
{
    p.sg = sg;
    p.fine = sg > 1;
    if (!p.fine) {
        p.nz = m.nz; p.nx = m.nx; p.isz = src_iz; p.isx = src_ix;
        p.nlev = 3;
        p.scale[0] = 27; p.size[0] = 2;   // ATR:1513-1514
        p.scale[1] = 9;  p.size[1] = 6;   // ATR:1685-1686
        p.scale[2] = 3;  p.size[2] = 13;  // ATR:1852-1853
    } else {
        p.nz = sg * (m.nz - 1) + 1; p.nx = sg * (m.nx - 1) + 1;
        p.isz = sg * src_iz; p.isx = sg * src_ix;
        p.nlev = 2;
        p.scale[0] = 9; p.size[0] = 2 * sg + (sg - 1) / 2;  // ATR:2188-2189
        p.scale[1] = 3; p.size[1] = p.size[0] + 3 * sg;     // ATR:2355-2356
        p.scale[2] = 1; p.size[2] = 0;
    }
    p.stop_r = p.size[p.nlev - 1] + handover_margin;
}

// Largest level grid (nodes) any source of this plan family can need.
ALI_HD size_t ali_plan_max_level_nodes(const AliSourcePlan &p)
{
    size_t best = 0;
    for (int l = 0; l < p.nlev; l++) {
        size_t w = (size_t)(2 * p.size[l] * p.scale[l] + 1);
        if (w * w > best) best = w * w;
    }
    return best;
}

struct AliSeqScratch {
    double *tA, *tB;     // level T buffers (ping-pong), each max_level_nodes
    int32_t *sA, *sB;    // level status buffers; sB is re-used for the main-grid window
    AliHeapEnt *heap;    // heap_cap entries
    double *hkey;        // heap_cap
    double *cval;        // evaluation cache of the grid being marched (status_cap entries), or nullptr
    uint8_t *cflag;
    int heap_cap;
    size_t status_cap;   // entries available in sA / sB
};

// Result of the sequential phase on the main grid.
struct AliSeqResult {
    int wz0, wx0, wnz, wnx; // window of the main grid whose statuses live in scratch.sB
    int overflow;
    AliSeqCounters cnt;
};

// ---------------------------------------------------------------------------
// One source, step by step.  The driver (ali_seq_source below for the all-sequential form,
// the ttf kernel / host replay for the hybrid form) calls, per level:
//   ali_src_level_geometry   every lane: extents, buffers, view of level l
//   ali_src_level_fill       cooperative: reset + analytic seed (level 0)
//   ali_src_level_start      one lane: perimeter push (level 0) or hand-off from level l-1
//   ali_src_level_seq        one lane: heap-ordered march, optionally only up to stop_r
// and the same four for the main grid.
// ---------------------------------------------------------------------------
struct AliSrcState {
    AliSeqGrid lv[2];
    AliSeqGrid mg;
    int cz[2], cx[2];
    AliSeqCounters cnt;
    int overflow;
    AliMatView base;   // view of the grid the levels refine: the coarse model (travel) or the sg-refined one
};

ALI_DEV void ali_src_begin(AliSrcState &s, const AliSourcePlan &p)
{
    s.cz[0] = s.cz[1] = s.cx[0] = s.cx[1] = 0;
    s.cnt.pops = s.cnt.evals = s.cnt.fallbacks = 0;
    s.cnt.cyc_heap = s.cnt.cyc_eval = 0;
    s.cnt.steps = s.cnt.computed = 0;
    s.cnt.cyc_total = s.cnt.cyc_fill = s.cnt.cyc_start = 0;
    s.overflow = 0;
    s.base = ali_make_view(1, 0, 0, p.fine ? p.sg : 1, p.fine ? 1 : 0);
}

ALI_DEV void ali_src_level_geometry(AliSrcState &s, const AliModel &m, const AliSourcePlan &p,
                                    const AliSeqScratch &sc, int l)
{
    const int cur = l & 1;
    AliSeqGrid &g = s.lv[cur];
    const int size = p.size[l], scl = p.scale[l];
    const int left = ali_imax(0, p.isx - size), right = ali_imin(p.nx - 1, p.isx + size);
    const int bottom = ali_imax(0, p.isz - size), top = ali_imin(p.nz - 1, p.isz + size);
    g.nz = scl * (top - bottom) + 1;
    g.nx = scl * (right - left) + 1;
    g.t = cur == 0 ? sc.tA : sc.tB;
    g.st = cur == 0 ? sc.sA : sc.sB;
    g.t_stride = g.nx; g.toff = 0;
    g.wz0 = 0; g.wx0 = 0; g.wnz = g.nz; g.wnx = g.nx;
    g.heap = sc.heap; g.heap_cap = sc.heap_cap; g.hkey = sc.hkey;
    g.cv = sc.cval; g.cf = sc.cflag;
    g.mv = ali_make_view(scl, bottom, left, p.fine ? p.sg : 1, 1);
    g.dnx = m.dnx / scl;
    s.cx[cur] = scl * (p.isx - left);
    s.cz[cur] = scl * (p.isz - bottom);
}

// Half-side of the level's initial alive square: the analytic seed on level 0, the previous
// level's box (in this level's nodes) otherwise.
ALI_HD int ali_src_level_ring(const AliSourcePlan &p, int l)
{
    if (l == 0) return p.fine ? (4 + 9 * ((p.sg - 1) / 2)) : 13;
    return p.scale[l] * p.size[l - 1];
}

ALI_DEV void ali_src_level_fill(AliSrcState &s, const AliModel &m, const AliSourcePlan &p, int l, int lane,
                                int nlanes, bool sync_between)
{
    const int cur = l & 1;
    AliSeqGrid &g = s.lv[cur];
    ali_seq_clear(g, true, lane, nlanes);
    if (l == 0) {
        if (sync_between) ALI_SYNCCTA();
        ali_seq_seed(g, m, s.base, p.isz, p.isx, s.cz[cur], s.cx[cur], ali_src_level_ring(p, 0),
                     p.fine ? 1.0 : -1.0, lane, nlanes);
    }
}

ALI_DEV void ali_src_level_start(AliSrcState &s, const AliSourcePlan &p, int l)
{
    const int cur = l & 1;
    if (l == 0) ali_seq_seed_push(s.lv[cur], s.cz[cur], s.cx[cur], ali_src_level_ring(p, 0));
    else ali_seq_handoff(s.lv[cur ^ 1], s.cz[cur ^ 1], s.cx[cur ^ 1], s.lv[cur], s.cz[cur], s.cx[cur]);
}

ALI_DEV int ali_src_level_seq(AliSrcState &s, const AliModel &m, const AliSourcePlan &p, int l, int stop_r)
{
    const int cur = l & 1;
    int why = ali_seq_march(s.lv[cur], m, s.cx[cur], s.cz[cur], p.scale[l] * p.size[l],
                            (!p.fine && l == 0) ? 1 : 0, stop_r, s.cnt);
    s.overflow |= s.lv[cur].overflow;
    return why;
}

ALI_DEV int ali_src_level_seq_coop(AliSrcState &s, const AliModel &m, const AliSourcePlan &p, int l, int lane, int nlanes)
{
    const int cur = l & 1;
    int why = ali_seq_march_coop(s.lv[cur], m, s.cx[cur], s.cz[cur], p.scale[l] * p.size[l],
                                 (!p.fine && l == 0) ? 1 : 0, -1, s.cnt, lane, nlanes);
    s.overflow |= s.lv[cur].overflow;
    return why;
}

// Main grid: statuses and travel times of a window around the source, in the level buffers that
// the last level does not use.  The band march copies the window into its own field afterwards.
ALI_DEV void ali_src_main_geometry(AliSrcState &s, const AliModel &m, const AliSourcePlan &p, const AliSeqScratch &sc)
{
    AliSeqGrid &mg = s.mg;
    const int last = (p.nlev - 1) & 1;
    const int half = p.stop_r + 4;
    mg.nz = p.nz; mg.nx = p.nx;
    mg.st = last == 0 ? sc.sB : sc.sA;
    mg.t = last == 0 ? sc.tB : sc.tA;
    mg.wz0 = ali_imax(0, p.isz - half); mg.wx0 = ali_imax(0, p.isx - half);
    mg.wnz = ali_imin(p.nz - 1, p.isz + half) - mg.wz0 + 1;
    mg.wnx = ali_imin(p.nx - 1, p.isx + half) - mg.wx0 + 1;
    mg.t_stride = mg.wnx;
    mg.toff = (long long)mg.wz0 * mg.wnx + mg.wx0;
    mg.heap = sc.heap; mg.heap_cap = sc.heap_cap; mg.hkey = sc.hkey;
    mg.cv = sc.cval; mg.cf = sc.cflag;
    mg.mv = s.base;
    mg.dnx = m.dnx;
    if ((size_t)mg.wnz * mg.wnx > sc.status_cap) s.overflow = 1;
}

ALI_DEV void ali_src_main_start_and_seq(AliSrcState &s, const AliModel &m, const AliSourcePlan &p)
{
    const int last = (p.nlev - 1) & 1;
    ali_seq_handoff(s.lv[last], s.cz[last], s.cx[last], s.mg, p.isz, p.isx);
    ali_seq_march(s.mg, m, p.isx, p.isz, -1, 0, p.stop_r, s.cnt);
    s.overflow |= s.mg.overflow;
}

// Levels + main-grid start for one source (result: statuses + travel times of res' window).  With an
// evaluation cache in the scratch (sc.cval) the marches run in their cooperative form, else lane
// 0 alone walks them and the other lanes only take part in the fills.
ALI_DEV void ali_seq_source(const AliModel &m, const AliSourcePlan &p, const AliSeqScratch &sc, AliSeqResult &res,
                            int lane, int nlanes)
{
    AliSrcState s;
    const bool coop = sc.cval != nullptr;
    const long long t_begin = ALI_CLOCK();
    ali_src_begin(s, p);
    for (int l = 0; l < p.nlev; l++) {
        long long t0 = ALI_CLOCK();
        ali_src_level_geometry(s, m, p, sc, l);
        ali_src_level_fill(s, m, p, l, lane, nlanes, true);
        ALI_SYNCCTA();
        long long t1 = ALI_CLOCK();
        if (lane == 0) ali_src_level_start(s, p, l);
        ALI_SYNCCTA();
        s.cnt.cyc_fill += t1 - t0;
        s.cnt.cyc_start += ALI_CLOCK() - t1;
        if (coop) ali_src_level_seq_coop(s, m, p, l, lane, nlanes);
        else if (lane == 0) ali_src_level_seq(s, m, p, l, -1);
        ALI_SYNCCTA();
    }
    ali_src_main_geometry(s, m, p, sc);
#if defined(__CUDA_ARCH__)
    {   // thread 0's verdict for everybody
        __shared__ int s_ovf;
        if (lane == 0) s_ovf = s.overflow;
        __syncthreads();
        s.overflow = s_ovf;
    }
#endif
    if (!s.overflow) {
        long long t0 = ALI_CLOCK();
        ali_seq_clear(s.mg, true, lane, nlanes);
        ALI_SYNCCTA();
        long long t1 = ALI_CLOCK();
        const int last = (p.nlev - 1) & 1;
        if (lane == 0) ali_seq_handoff(s.lv[last], s.cz[last], s.cx[last], s.mg, p.isz, p.isx);
        ALI_SYNCCTA();
        s.cnt.cyc_fill += t1 - t0;
        s.cnt.cyc_start += ALI_CLOCK() - t1;
        if (coop) ali_seq_march_coop(s.mg, m, p.isx, p.isz, -1, 0, p.stop_r, s.cnt, lane, nlanes);
        else if (lane == 0) ali_seq_march(s.mg, m, p.isx, p.isz, -1, 0, p.stop_r, s.cnt);
        s.overflow |= s.mg.overflow;
        ALI_SYNCCTA();
    }
    s.cnt.cyc_total = ALI_CLOCK() - t_begin;
    res.wz0 = s.mg.wz0; res.wx0 = s.mg.wx0; res.wnz = s.mg.wnz; res.wnx = s.mg.wnx;
    res.overflow = s.overflow;
    res.cnt = s.cnt;
}
