// ali_band.cuh -- per-node phases of the band-synchronous narrow-band march.
//
// The reference pops one node at a time from a heap (ATR:2055-2102, 2775-2817).  The
// march here advances the whole narrow band in rounds that reproduce the same discrete
// solution:
//   A  every band node re-evaluates the ALI update from the state at the start of the
//      round (alive values + last round's published tentatives) into a staging slot;
//   B  staged values are published (T, status QUEUED -> BAND) and tmin is reduced;
//   C  nodes with T <= tmin + delta become alive; their far 4-neighbours join the band
//      as QUEUED (in the list, not yet visible to stencils, like a node the reference
//      has not computed yet).
// delta = frac * dnx / vmax with frac <= 0.4 keeps the result identical to heap order.
#pragma once
#include "ali_core.cuh"

// Node status of the band march.  avail (reference: nsts >= 0) <=> status >= BAND.
#define ALI_ST_FAR 0
#define ALI_ST_QUEUED 1
#define ALI_ST_BAND 2
#define ALI_ST_ALIVE 3

struct AliBandGrid {
    int nz, nx;
    double *T;        // [nz*nx] travel times of this source (seconds * sg on the fine path)
    uint8_t *st;      // [nz*nx]
    AliMatView mv;
    double dnx;
    ALI_DEV bool avail(int z, int x) const { return st[(size_t)z * nx + x] >= ALI_ST_BAND; }
    ALI_DEV bool alive(int z, int x) const { return st[(size_t)z * nx + x] == ALI_ST_ALIVE; }
    ALI_DEV double tt(int z, int x) const { return T[(size_t)z * nx + x]; }
};

// Phase A: value the reference would store for this node given the current state.
ALI_DEV double ali_band_eval(const AliModel &m, const AliBandGrid &g, int node, int *fallback)
{
    int iz = node / g.nx, ix = node - iz * g.nx;
    return ali_eval_node(m, g.mv, g, iz, ix, g.nz, g.nx, g.nz, g.nx, g.dnx, fallback);
}

// Phase B: make the staged value visible.
ALI_DEV void ali_band_publish(const AliBandGrid &g, int node, double v)
{
    g.T[node] = v;
    if (g.st[node] == ALI_ST_QUEUED) g.st[node] = ALI_ST_BAND;
}

// Claims a far node for the band list; returns true for exactly one caller.
ALI_DEV bool ali_band_claim(const AliBandGrid &g, int node)
{
    if (g.st[node] != ALI_ST_FAR) return false;
#if defined(__CUDA_ARCH__)
    // byte-wide claim through a 32-bit atomicOr on the containing word: within phase C a
    // FAR byte can only turn QUEUED, so OR-ing bit 0 never corrupts another state.
    unsigned *word = (unsigned *)((uintptr_t)(g.st + node) & ~(uintptr_t)3);
    unsigned shift = 8u * (unsigned)((uintptr_t)(g.st + node) & 3);
    unsigned old = atomicOr(word, (unsigned)ALI_ST_QUEUED << shift);
    return ((old >> shift) & 0xffu) == ALI_ST_FAR;
#else
    g.st[node] = ALI_ST_QUEUED;
    return true;
#endif
}

// Phase C for an accepted node: alive + enlist far 4-neighbours (ATR:2065-2102 order is
// irrelevant here: new nodes are evaluated next round from a common snapshot).
ALI_DEV int ali_band_accept(const AliBandGrid &g, int node, int *nb)
{
    int iz = node / g.nx, ix = node - iz * g.nx;
    int cnt = 0;
    g.st[node] = ALI_ST_ALIVE;
    if (ix > 0 && ali_band_claim(g, node - 1)) nb[cnt++] = node - 1;
    if (ix < g.nx - 1 && ali_band_claim(g, node + 1)) nb[cnt++] = node + 1;
    if (iz > 0 && ali_band_claim(g, node - g.nx)) nb[cnt++] = node - g.nx;
    if (iz < g.nz - 1 && ali_band_claim(g, node + g.nx)) nb[cnt++] = node + g.nx;
    return cnt;
}

// Largest phase velocity a coarse node can produce (1-degree sampling), for delta.
ALI_DEV double ali_node_vmax(const AliModel &m, int iz, int ix)
{
    AliMat mat;
    const AliMatView idv = ali_view_identity();
    ali_fetch_mat(m, idv, iz, ix, mat, true);
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
