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

#if defined(__CUDACC__)
#define ALI_SYNCWARP() __syncwarp()
#else
#define ALI_SYNCWARP() ((void)0)
#endif

// Node state of one grid during the sequential march.  Status follows the reference:
// -1 far, 0 alive, >0 position in the heap (ATR:103).  Status is stored for a window
// [wz0, wz0+wnz) x [wx0, wx0+wnx) only; everything outside it is far by construction.
struct AliSeqGrid {
    int nz, nx;           // full extents of this grid (edge logic, absolute coordinates)
    double *t;            // T(z, x) = t[z * t_stride + x]
    long long t_stride;
    int32_t *st;
    int wz0, wx0, wnz, wnx;
    int32_t *heap;        // (iz, ix) pairs, 1-indexed (ATR:118-119)
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
    ALI_DEV bool avail(int z, int x) const { return in_win(z, x) && s(z, x) >= 0; }
    ALI_DEV bool alive(int z, int x) const { return in_win(z, x) && s(z, x) == 0; }
    ALI_DEV double &tref(int z, int x) const { return t[(long long)z * t_stride + x]; }
    ALI_DEV double tt(int z, int x) const { return t[(long long)z * t_stride + x]; }
};

// Python round(k / 2): round-half-to-even (ATR:123, 135, 160, 172).
ALI_DEV int ali_half_round(int k)
{
    int h = k >> 1;
    if (k & 1) return (h & 1) ? h + 1 : h;
    return h;
}

ALI_DEV void ali_heap_swap(AliSeqGrid &g, int a, int b)
{
    int32_t e0 = g.heap[2 * a], e1 = g.heap[2 * a + 1];
    g.heap[2 * a] = g.heap[2 * b]; g.heap[2 * a + 1] = g.heap[2 * b + 1];
    g.heap[2 * b] = e0; g.heap[2 * b + 1] = e1;
}

// Sift-up shared by addtree / updtree (ATR:122-137, 159-174).
ALI_DEV void ali_sift_up(AliSeqGrid &g, int iz, int ix, int tpc)
{
    int tpp = ali_half_round(tpc);
    double tv = g.tt(iz, ix);
    while (tpp > 0) {
        int aa = g.heap[2 * tpp], bb = g.heap[2 * tpp + 1];
        if (tv < g.tt(aa, bb)) {
            g.s(iz, ix) = tpp;
            g.s(aa, bb) = tpc;
            ali_heap_swap(g, tpc, tpp);
            tpc = tpp;
            tpp = ali_half_round(tpc);
        } else {
            tpp = 0;
        }
    }
}

ALI_DEV void ali_addtree(AliSeqGrid &g, int iz, int ix)
{
    if (g.ntr + 2 >= g.heap_cap) { g.overflow = 1; return; }
    g.ntr += 1;
    g.s(iz, ix) = g.ntr;
    g.heap[2 * g.ntr] = iz;
    g.heap[2 * g.ntr + 1] = ix;
    ali_sift_up(g, iz, ix, g.ntr);
}

ALI_DEV void ali_updtree(AliSeqGrid &g, int iz, int ix) { ali_sift_up(g, iz, ix, g.s(iz, ix)); }

// ATR:178-237.
ALI_DEV void ali_downtree(AliSeqGrid &g)
{
    int ntr = g.ntr;
    if (ntr == 1) { g.ntr = 0; return; }
    g.s(g.heap[2 * ntr], g.heap[2 * ntr + 1]) = 1;
    g.heap[2] = g.heap[2 * ntr];
    g.heap[3] = g.heap[2 * ntr + 1];
    ntr -= 1;
    int tpp = 1, tpc = 2;
    while (tpc < ntr) {
        double rd1 = g.tt(g.heap[2 * tpc], g.heap[2 * tpc + 1]);
        double rd2 = g.tt(g.heap[2 * tpc + 2], g.heap[2 * tpc + 3]);
        if (rd1 > rd2) { tpc += 1; rd1 = rd2; }
        rd2 = g.tt(g.heap[2 * tpp], g.heap[2 * tpp + 1]);
        if (rd1 < rd2) {
            g.s(g.heap[2 * tpp], g.heap[2 * tpp + 1]) = tpc;
            g.s(g.heap[2 * tpc], g.heap[2 * tpc + 1]) = tpp;
            ali_heap_swap(g, tpc, tpp);
            tpp = tpc;
            tpc = 2 * tpp;
        } else {
            tpc = ntr + 1;
        }
    }
    if (tpc == ntr) {
        double rd1 = g.tt(g.heap[2 * tpc], g.heap[2 * tpc + 1]);
        double rd2 = g.tt(g.heap[2 * tpp], g.heap[2 * tpp + 1]);
        if (rd1 < rd2) {
            g.s(g.heap[2 * tpp], g.heap[2 * tpp + 1]) = tpc;
            g.s(g.heap[2 * tpc], g.heap[2 * tpc + 1]) = tpp;
            ali_heap_swap(g, tpc, tpp);
        }
    }
    g.ntr = ntr;
}

struct AliSeqCounters {
    long long pops, evals, fallbacks;
    long long cyc_heap, cyc_eval;   // device builds: SM cycles in heap operations / evaluations
    long long steps, computed;      // cooperative march: warp-wide evaluation steps, evaluations executed in them
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
        const int ix = g.heap[3], iz = g.heap[2];
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
                    double v = ali_eval_node(m, g.mv, g, z, x, nnz_l, nnx, nnz, nnx, g.dnx, &fb);
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
    int miss, mz, mx, mnnz;      // evaluation lane 0 is waiting for
    int miss_ready, miss_fb;
    double miss_v;
    int finished, why;
};

ALI_DEV void ali_seq_invalidate(const AliSeqGrid &g, int z, int x)
{
    if (z - 2 >= g.wz0 && z + 2 < g.wz0 + g.wnz && x - 2 >= g.wx0 && x + 2 < g.wx0 + g.wnx) {
        uint8_t *f = g.cf + g.widx(z, x);
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

#define ALI_COOP_HEAP_POSITIONS 8
// One lane's share of an evaluation step.  ntr: current heap size (lane 0's value).
ALI_DEV int ali_coop_step(const AliSeqGrid &g, const AliModel &m, AliCoopState &cs, int lane, int ntr)
{
    const int nnx = g.nx, nnz = g.nz;
    int cz = 0, cx = 0, cnnz = nnz;
    bool want = false;
    const bool serve = (lane == 0 && cs.miss);
    if (serve) {
        cz = cs.mz; cx = cs.mx; cnnz = cs.mnnz;
        want = true;
    } else {
        const int slot = lane == 0 ? 4 * ALI_COOP_HEAP_POSITIONS - 1 : lane - 1;
        const int hp = 1 + (slot >> 2), dir = slot & 3;
        if (hp <= ntr && hp <= ALI_COOP_HEAP_POSITIONS) {
            cz = g.heap[2 * hp]; cx = g.heap[2 * hp + 1];
            if (dir == 0) cx -= 1; else if (dir == 1) cx += 1; else if (dir == 2) cz -= 1; else cz += 1;
            if (cz >= 0 && cz < nnz && cx >= 0 && cx < nnx && g.in_win(cz, cx))
                want = g.s(cz, cx) != 0 && g.cf[g.widx(cz, cx)] == 0;
        }
    }
    if (want) {
        int fb = 0;
        const double v = ali_eval_node(m, g.mv, g, cz, cx, cnnz, nnx, nnz, nnx, g.dnx, &fb);
        if (serve) {
            cs.miss_v = v; cs.miss_fb = fb; cs.miss_ready = 1; cs.miss = 0;
        } else if (!fb) {
            g.cv[g.widx(cz, cx)] = v;
            g.cf[g.widx(cz, cx)] = 1;
        }
    }
    return want ? 1 : 0;
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
            cs.ix = g.heap[3]; cs.iz = g.heap[2];
            g.s(cs.iz, cs.ix) = 0;
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
                if (!g.in_win(z, x)) { cs.finished = 1; cs.why = ALI_SEQ_LIMIT; continue; }
                const int32_t stv = g.s(z, x);
                if (stv != 0) {
                    const int nnz_l = (nnz_bug && s < 2 && stv > 0) ? nnx : nnz;
                    const size_t wi = g.widx(z, x);
                    double v;
                    int fb = 0;
                    if (cs.miss_ready) { v = cs.miss_v; fb = cs.miss_fb; cs.miss_ready = 0; }
                    else if (nnz_l == nnz && g.cf[wi]) v = g.cv[wi];
                    else { cs.miss = 1; cs.mz = z; cs.mx = x; cs.mnnz = nnz_l; return 0; }
                    cnt.evals++;
                    cnt.fallbacks += fb;
                    const bool changed = (stv == -1) || !(g.tt(z, x) == v);
                    g.tref(z, x) = v;
                    if (changed) ali_seq_invalidate(g, z, x);
                    if (nnz_l == nnz && !fb) { g.cv[wi] = v; g.cf[wi] = 1; }
                    else g.cf[wi] = 0;
                    if (stv == -1) ali_addtree(g, z, x);
                    else ali_updtree(g, z, x);
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
ALI_DEV int ali_seq_march_coop(AliSeqGrid &g, const AliModel &m, int cx, int cz, int max_dist, int nnz_bug,
                               int stop_r, AliSeqCounters &cnt, int lane, int nlanes)
{
    AliCoopState cs;
    cs.have_p = 0; cs.iz = cs.ix = cs.s = 0;
    cs.miss = 0; cs.mz = cs.mx = cs.mnnz = 0;
    cs.miss_ready = 0; cs.miss_fb = 0; cs.miss_v = 0.0;
    cs.finished = 0; cs.why = ALI_SEQ_EMPTY;
    for (;;) {
        int done = 0;
#if defined(__CUDA_ARCH__)
        const int ntr = __shfl_sync(0xffffffffu, g.ntr, 0);
        const int did = ali_coop_step(g, m, cs, lane, ntr);
        const unsigned mask = __ballot_sync(0xffffffffu, did);
        if (lane == 0) {
            cnt.steps++;
            cnt.computed += __popc(mask);
            done = ali_coop_advance(g, cs, cx, cz, max_dist, nnz_bug, stop_r, cnt);
        }
        __syncwarp();
        done = __shfl_sync(0xffffffffu, done, 0);
#else
        (void)lane;
        cnt.steps++;
        for (int l = 0; l < nlanes; l++) cnt.computed += ali_coop_step(g, m, cs, l, g.ntr);
        done = ali_coop_advance(g, cs, cx, cz, max_dist, nnz_bug, stop_r, cnt);
#endif
        if (done) break;
    }
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
        for (size_t i = lane; i < n; i += nlanes) g.t[i] = 0.0; // level grids only (window == grid)
    g.ntr = 0;
    g.overflow = 0;
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
ALI_DEV void ali_seq_handoff(const AliSeqGrid &a, int cza, int cxa, AliSeqGrid &b, int czb, int cxb)
{
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
    int32_t *heap;       // 2 * heap_cap
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
    g.t_stride = g.nx;
    g.wz0 = 0; g.wx0 = 0; g.wnz = g.nz; g.wnx = g.nx;
    g.heap = sc.heap; g.heap_cap = sc.heap_cap;
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
        if (sync_between) ALI_SYNCWARP();
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

ALI_DEV void ali_src_main_geometry(AliSrcState &s, const AliModel &m, const AliSourcePlan &p, const AliSeqScratch &sc,
                                   double *T)
{
    AliSeqGrid &mg = s.mg;
    const int last = (p.nlev - 1) & 1;
    const int half = p.stop_r + 4;
    mg.nz = p.nz; mg.nx = p.nx;
    mg.t = T; mg.t_stride = p.nx;
    mg.st = last == 0 ? sc.sB : sc.sA;
    mg.wz0 = ali_imax(0, p.isz - half); mg.wx0 = ali_imax(0, p.isx - half);
    mg.wnz = ali_imin(p.nz - 1, p.isz + half) - mg.wz0 + 1;
    mg.wnx = ali_imin(p.nx - 1, p.isx + half) - mg.wx0 + 1;
    mg.heap = sc.heap; mg.heap_cap = sc.heap_cap;
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

// Levels + main-grid start for one source.  `T` is the source's main-grid field.  With an
// evaluation cache in the scratch (sc.cval) the marches run in their cooperative form, else lane
// 0 alone walks them and the other lanes only take part in the fills.
ALI_DEV void ali_seq_source(const AliModel &m, const AliSourcePlan &p, const AliSeqScratch &sc, double *T,
                            AliSeqResult &res, int lane, int nlanes)
{
    AliSrcState s;
    const bool coop = sc.cval != nullptr;
    ali_src_begin(s, p);
    for (int l = 0; l < p.nlev; l++) {
        ali_src_level_geometry(s, m, p, sc, l);
        ali_src_level_fill(s, m, p, l, lane, nlanes, true);
        ALI_SYNCWARP();
        if (lane == 0) ali_src_level_start(s, p, l);
        ALI_SYNCWARP();
        if (coop) ali_src_level_seq_coop(s, m, p, l, lane, nlanes);
        else if (lane == 0) ali_src_level_seq(s, m, p, l, -1);
        ALI_SYNCWARP();
    }
    ali_src_main_geometry(s, m, p, sc, T);
#if defined(__CUDA_ARCH__)
    s.overflow = __shfl_sync(0xffffffffu, s.overflow, 0);
#endif
    if (!s.overflow) {
        ali_seq_clear(s.mg, false, lane, nlanes);
        ALI_SYNCWARP();
        const int last = (p.nlev - 1) & 1;
        if (lane == 0) ali_seq_handoff(s.lv[last], s.cz[last], s.cx[last], s.mg, p.isz, p.isx);
        ALI_SYNCWARP();
        if (coop) ali_seq_march_coop(s.mg, m, p.isx, p.isz, -1, 0, p.stop_r, s.cnt, lane, nlanes);
        else if (lane == 0) ali_seq_march(s.mg, m, p.isx, p.isz, -1, 0, p.stop_r, s.cnt);
        s.overflow |= s.mg.overflow;
        ALI_SYNCWARP();
    }
    res.wz0 = s.mg.wz0; res.wx0 = s.mg.wx0; res.wnz = s.mg.wnz; res.wnx = s.mg.wnx;
    res.overflow = s.overflow;
    res.cnt = s.cnt;
}
