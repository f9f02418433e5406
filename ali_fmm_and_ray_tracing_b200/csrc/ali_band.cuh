// ali_band.cuh -- per-node phases of the band-synchronous narrow-band march.
//
// The reference pops one node at a time from a heap (ATR:2055-2102, 2775-2817).  The
// march here advances the whole narrow band in rounds that reproduce the same discrete
// solution wherever that solution does not hinge on the heap's pop timing (SURVEY.md 7.3,
// DESIGN.md "Parity"):
//   A  every band node whose 12-neighbour window changed since its last evaluation
//      re-evaluates the ALI update from the state at the start of the round (alive values
//      + last round's published tentatives) into a staging slot -- the update is a pure
//      function of that window, so skipping unchanged windows equals re-evaluating all;
//   B  changed values are published (T, status QUEUED -> BAND), their neighbours are marked
//      dirty, and tmin is reduced;
//   C  nodes with T <= tmin + delta become alive; their far 4-neighbours join the band
//      as QUEUED (in the list, not yet visible to stencils, like a node the reference
//      has not computed yet).
// delta = frac * dnx / vmax with frac <= 0.4.
#pragma once
#include "ali_core.cuh"

// Node state of the band march lives in the travel-time word itself:
//   far       NaN, all bits set (the field is pre-filled with 0xFF bytes)
//   enlisted  NaN, lowest bit clear: in the band list, no estimate published yet
//   estimate  any value >= 0: available to stencils (reference: nsts >= 0), tentative or final
// so the 12-neighbour gather and the enlisting of far neighbours touch one array (whose
// sectors the gathers keep in L1).  A separate byte per node records "alive"; it is written
// once per node and read only by the FD fallback.  Dirty flags are bytes in 16 x 8-node tiles
// (one 128-byte line each) so that vertical band segments do not take one line per node.
#define ALI_ST_FAR 0
#define ALI_ST_ALIVE 3
#define ALI_T_UNSET_BYTE 0xFF
#define ALI_T_FAR_BITS 0xFFFFFFFFFFFFFFFFull
#define ALI_T_ENLISTED_BITS 0xFFFFFFFFFFFFFFFEull
#define ALI_T_NAN_VALUE_BITS 0x7FF8000000000001ll   /* a computed NaN (degenerate material), canonicalised */

// Band list entries pack the node as (iz << 16) | ix (grids up to 65535 x 65535).
#define ALI_PACK(iz, ix) (((unsigned)(iz) << 16) | (unsigned)(ix))
#define ALI_PACK_Z(p) ((int)((p) >> 16))
#define ALI_PACK_X(p) ((int)((p) & 0xffffu))

// Field layout.  The kernel marches on 4 x 4-node tiles (one 128-byte line per tile): a warp
// works on 32 neighbouring front nodes, and on a row-major field every node of a steep front
// segment sits in a line of its own, so each of the 12 gather instructions costs 32 L1 tag
// look-ups -- the bound of phase A.  Tiles make that 8-9 for any front direction.  The host
// replay keeps the row-major layout (t4x == 0); the finalize kernel writes the caller's
// row-major field.
struct AliBandGrid {
    int nz, nx;
    double *T;        // travel times of this source (seconds * sg on the fine path), layout per t4x
    uint8_t *st;      // same indexing; ALI_ST_ALIVE once accepted
    uint8_t *dirty;   // host replay only: tiled [tiles_z][tiles_x][8][16]; 1: window changed since the last evaluation
    int tiles_x;
    int t4x;          // 0: row-major [nz][nx]; else the number of 4-node tiles per row, (nx + 3) / 4
    AliMatView mv;
    double dnx;
    ALI_DEV size_t ti(int z, int x) const
    {
        // 32-bit arithmetic: a field has fewer than 2^31 nodes (checked by alifmm_ttf), tile padding adds < 1 %
        if (t4x) return (size_t)(((((unsigned)z >> 2) * (unsigned)t4x + ((unsigned)x >> 2)) << 4) | (((unsigned)z & 3u) << 2) | ((unsigned)x & 3u));
        return (size_t)z * nx + x;
    }
    ALI_DEV bool avail(int z, int x) const { return T[ti(z, x)] >= 0.0; } // false for NaN
    ALI_DEV bool alive(int z, int x) const { return st[ti(z, x)] == ALI_ST_ALIVE; }
    ALI_DEV double tt(int z, int x) const { return T[ti(z, x)]; }
};

ALI_HD size_t ali_field_nodes_tiled(int nz, int nx) { return (size_t)((nz + 3) >> 2) * (size_t)((nx + 3) >> 2) * 16; }

ALI_HD int ali_dirty_tiles_x(int nx) { return (nx + 15) >> 4; }
ALI_HD size_t ali_dirty_bytes(int nz, int nx) { return (size_t)((nz + 7) >> 3) * (size_t)ali_dirty_tiles_x(nx) * 128; }
ALI_DEV size_t ali_dirty_index(const AliBandGrid &g, int iz, int ix)
{
    return ((size_t)(iz >> 3) * g.tiles_x + (size_t)(ix >> 4)) * 128 + (size_t)(((iz & 7) << 4) | (ix & 15));
}

ALI_HD AliMatView ali_band_view(int sg)
{
    return ali_make_view(1, 0, 0, sg > 1 ? sg : 1, sg > 1 ? 1 : 0);
}

// 12-neighbour gather; interior nodes take a branch-free path with independent loads.
ALI_DEV void ali_band_gather(const AliBandGrid &g, int iz, int ix, AliWindow &w)
{
    if (iz >= 2 && iz < g.nz - 2 && ix >= 2 && ix < g.nx - 2) {
        unsigned av = 0;
        if (g.t4x) {
            // tiled address = row term + column term
            unsigned rt[5], ct[5];
#pragma unroll
            for (int q = 0; q < 5; q++) {
                const int z = iz + q - 2, x = ix + q - 2;
                rt[q] = (((unsigned)(z >> 2) * (unsigned)g.t4x) << 4) | (unsigned)((z & 3) << 2);
                ct[q] = ((unsigned)(x >> 2) << 4) | (unsigned)(x & 3);
            }
#pragma unroll
            for (int k = 0; k < 12; k++) w.t[k] = g.T[rt[ALI_W_DZ(k) + 2] + ct[ALI_W_DX(k) + 2]];
        } else {
            const double *tp = g.T + ((size_t)iz * g.nx + ix);
            const int nx = g.nx;
#pragma unroll
            for (int k = 0; k < 12; k++) w.t[k] = tp[ALI_W_DZ(k) * nx + ALI_W_DX(k)];
        }
#pragma unroll
        for (int k = 0; k < 12; k++)
            if (w.t[k] >= 0.0) av |= 1u << k;
        w.avail = av;
    } else {
        ali_gather(g, iz, ix, g.nz, g.nx, w);
    }
}

// FD fallback of a band node, out of line and through pointers so that the hot path keeps
// its state in registers.  m_dev / g_mem point to copies of the model / grid structs in
// memory (device global resp. shared).
ALI_DEV_NOINLINE double ali_band_fouds_slow(const AliModel *m_dev, const AliBandGrid *g_mem, int iz, int ix)
{
    AliModel m = *m_dev;
    AliBandGrid g = *g_mem;
    AliMat mat;
    ali_fetch_mat(m, g.mv, iz, ix, mat);
    return ali_fouds18(m, mat, g, iz, ix, g.dnx, g.dnx, g.nx, g.nz);
}

// Phase A: value the reference would store for this node given the current state.
ALI_DEV double ali_band_eval(const AliModel &m, const AliModel *m_dev, const AliBandGrid &g,
                             const AliBandGrid *g_mem, int iz, int ix, int *fallback, const uint64_t *mt = nullptr,
                             bool level_view = false)
{
    AliMat mat;
    AliWindow w;
    // the kernel's grid is the model refined by subgrid only (ali_band_view); the host replay can also
    // run band rounds on a source level, whose view has a scale and an origin of its own
    if (level_view) ali_fetch_mat(m, g.mv, iz, ix, mat);
    else ali_fetch_mat_refined(m, g.mv.scale0, g.mv.side0, g.mv.mul0, g.mv.cast, iz, ix, mat);
    ali_band_gather(g, iz, ix, w);
    double v = ali_update_window(m, mat, w, iz, ix, g.nz, g.nx, g.dnx, nullptr, mt);
    if (v == -1.0) {
        v = ali_band_fouds_slow(m_dev, g_mem, iz, ix);
        *fallback = 1;
    }
    return v;
}

// Marks the 12 window neighbours of a node dirty (the window relation is symmetric).
ALI_DEV void ali_band_mark_dirty(const AliBandGrid &g, int iz, int ix)
{
    if (iz >= 2 && iz < g.nz - 2 && ix >= 2 && ix < g.nx - 2) {
        // tiled address = row term + column term
        size_t rt[5], ct[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int z = iz + q - 2, x = ix + q - 2;
            rt[q] = (size_t)(z >> 3) * g.tiles_x * 128 + (size_t)((z & 7) << 4);
            ct[q] = (size_t)(x >> 4) * 128 + (size_t)(x & 15);
        }
#pragma unroll
        for (int k = 0; k < 12; k++) g.dirty[rt[ALI_W_DZ(k) + 2] + ct[ALI_W_DX(k) + 2]] = 1;
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) {
            int z = iz + ALI_W_DZ(k), x = ix + ALI_W_DX(k);
            if (z >= 0 && z < g.nz && x >= 0 && x < g.nx) g.dirty[ali_dirty_index(g, z, x)] = 1;
        }
    }
}

// Phase B: make the staged value visible if it changed anything.
ALI_DEV bool ali_band_changed(const AliBandGrid &g, int iz, int ix, double v)
{
    return !(g.T[g.ti(iz, ix)] == v);   // also true for a node without an estimate yet (NaN)
}

ALI_DEV void ali_band_store(const AliBandGrid &g, int iz, int ix, double v)
{
    g.T[g.ti(iz, ix)] = v;
    ali_band_mark_dirty(g, iz, ix);
}

ALI_DEV void ali_band_publish(const AliBandGrid &g, int iz, int ix, double v)
{
    if (ali_band_changed(g, iz, ix, v)) ali_band_store(g, iz, ix, v);
}

// Claims a far node for the band list; returns true for exactly one caller.
ALI_DEV bool ali_band_claim(const AliBandGrid &g, size_t node)
{
    unsigned long long *w = (unsigned long long *)(g.T + node);
#if defined(__CUDA_ARCH__)
    if (*(volatile unsigned long long *)w != ALI_T_FAR_BITS) return false;
    return atomicCAS(w, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS;
#else
    if (*w != ALI_T_FAR_BITS) return false;
    *w = ALI_T_ENLISTED_BITS;
    return true;
#endif
}

// Phase C for an accepted node: alive + enlist far 4-neighbours (the reference's visiting
// order, ATR:2065-2102, is irrelevant here: new nodes are evaluated next round from a
// common snapshot).  Returns the packed new entries.
ALI_DEV int ali_band_accept(const AliBandGrid &g, int iz, int ix, unsigned *nb)
{
    int cnt = 0;
    const size_t me = g.ti(iz, ix);
    g.st[me] = ALI_ST_ALIVE;
    const bool hw = ix > 0, he = ix < g.nx - 1, hn = iz > 0, hs = iz < g.nz - 1;
    size_t nw, ne, nn, ns;
    if (g.t4x) {   // step inside the tile, or to the facing edge of the next tile
        const size_t trow = (size_t)g.t4x << 4;
        nw = (ix & 3) ? me - 1 : me - 13;
        ne = ((ix & 3) != 3) ? me + 1 : me + 13;
        nn = (iz & 3) ? me - 4 : me - trow + 12;
        ns = ((iz & 3) != 3) ? me + 4 : me + trow - 12;
    } else {
        nw = me - 1; ne = me + 1; nn = me - g.nx; ns = me + g.nx;
    }
#if defined(__CUDA_ARCH__)
    // The four state words are loaded together (one memory latency), then only far ones are claimed.  Plain
    // (L1-cached) loads: the lines are usually still there from phase A's gathers, and a stale word can only
    // read "far" for a node that has been claimed since (states never return to far) -- the CAS settles that.
    const unsigned long long *tw = (const unsigned long long *)g.T;
    const unsigned long long vw = hw ? tw[nw] : 0ull, ve = he ? tw[ne] : 0ull;
    const unsigned long long vn = hn ? tw[nn] : 0ull, vs = hs ? tw[ns] : 0ull;
    unsigned long long *cw = (unsigned long long *)g.T;
    if (vw == ALI_T_FAR_BITS && atomicCAS(cw + nw, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz, ix - 1);
    if (ve == ALI_T_FAR_BITS && atomicCAS(cw + ne, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz, ix + 1);
    if (vn == ALI_T_FAR_BITS && atomicCAS(cw + nn, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz - 1, ix);
    if (vs == ALI_T_FAR_BITS && atomicCAS(cw + ns, ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz + 1, ix);
#else
    if (hw && ali_band_claim(g, nw)) nb[cnt++] = ALI_PACK(iz, ix - 1);
    if (he && ali_band_claim(g, ne)) nb[cnt++] = ALI_PACK(iz, ix + 1);
    if (hn && ali_band_claim(g, nn)) nb[cnt++] = ALI_PACK(iz - 1, ix);
    if (hs && ali_band_claim(g, ns)) nb[cnt++] = ALI_PACK(iz + 1, ix);
#endif
    return cnt;
}

#if defined(__CUDACC__)
// ali_band_accept in two steps, so that the state words of the NEXT entry a thread will accept are in flight
// while it claims for the current one (tiled fields only).  A word loaded early can only be stale as "far".
ALI_DEV void ali_band_accept_peek(const AliBandGrid &g, int iz, int ix, unsigned long long &vw, unsigned long long &ve,
                                  unsigned long long &vn, unsigned long long &vs)
{
    const size_t me = g.ti(iz, ix), trow = (size_t)g.t4x << 4;
    const unsigned long long *tw = (const unsigned long long *)g.T;
    vw = ix > 0 ? tw[(ix & 3) ? me - 1 : me - 13] : 0ull;
    ve = ix < g.nx - 1 ? tw[((ix & 3) != 3) ? me + 1 : me + 13] : 0ull;
    vn = iz > 0 ? tw[(iz & 3) ? me - 4 : me - trow + 12] : 0ull;
    vs = iz < g.nz - 1 ? tw[((iz & 3) != 3) ? me + 4 : me + trow - 12] : 0ull;
}

ALI_DEV int ali_band_accept_peeked(const AliBandGrid &g, int iz, int ix, unsigned long long vw, unsigned long long ve,
                                   unsigned long long vn, unsigned long long vs, unsigned *nb)
{
    int cnt = 0;
    const size_t me = g.ti(iz, ix), trow = (size_t)g.t4x << 4;
    g.st[me] = ALI_ST_ALIVE;
    unsigned long long *cw = (unsigned long long *)g.T;
    if (vw == ALI_T_FAR_BITS && atomicCAS(cw + ((ix & 3) ? me - 1 : me - 13), ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz, ix - 1);
    if (ve == ALI_T_FAR_BITS && atomicCAS(cw + (((ix & 3) != 3) ? me + 1 : me + 13), ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz, ix + 1);
    if (vn == ALI_T_FAR_BITS && atomicCAS(cw + ((iz & 3) ? me - 4 : me - trow + 12), ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz - 1, ix);
    if (vs == ALI_T_FAR_BITS && atomicCAS(cw + (((iz & 3) != 3) ? me + 4 : me + trow - 12), ALI_T_FAR_BITS, ALI_T_ENLISTED_BITS) == ALI_T_FAR_BITS)
        nb[cnt++] = ALI_PACK(iz + 1, ix);
    return cnt;
}
#endif

// Largest phase velocity a coarse node can produce (1-degree sampling), for delta.
ALI_DEV double ali_node_vmax(const AliModel &m, int iz, int ix)
{
    AliMat mat;
    const AliMatView idv = ali_view_identity();
    ali_fetch_mat(m, idv, iz, ix, mat);
    double best = 0.0;
    if (mat.velpn != 0 || !m.has_stif) {
        for (int a = 0; a < 180; a++) {
            double v = mat.vel_map * m.phase_tab[(size_t)a * m.ncol + mat.velpn];
            if (v > best) best = v;
        }
    } else {
        for (int a = 0; a < 180; a++) {
            double v = ali_christoffel_phase((double)a, mat.s, mat.vel_map);
            if (v > best) best = v;
        }
    }
    return best;
}

// ---------------------------------------------------------------------------
// Hybrid source levels: a refined level grid can be marched by band rounds too, between a
// sequential warm-up (ali_seq.cuh) and the moment its front leaves the refined box.
// ---------------------------------------------------------------------------
// Box edges whose crossing ends a level (the edges the domain did not clip; ATR:1651-1652,
// 1673-1674): bit 0 left (x = 0), 1 right, 2 top (z = 0), 3 bottom.
ALI_HD int ali_level_stop_mask(int nz, int nx, int cz, int cx, int max_dist)
{
    int mask = 0;
    if (cx == max_dist) mask |= 1;
    if (nx - 1 - cx == max_dist) mask |= 2;
    if (cz == max_dist) mask |= 4;
    if (nz - 1 - cz == max_dist) mask |= 8;
    return mask;
}

ALI_DEV bool ali_level_on_stop_edge(int mask, int nz, int nx, int iz, int ix)
{
    return ((mask & 1) && ix == 0) || ((mask & 2) && ix == nx - 1) || ((mask & 4) && iz == 0) ||
           ((mask & 8) && iz == nz - 1);
}
