/*
 * ali_oracle.c -- CPU restatement of the ALI-FMM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path.  It is imported by tests/,
 * by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs, and by nothing else; the product (ali_fmm_and_ray_tracing_b200/) never
 * links or calls it.
 *
 * It restates, in plain sequential C, what the reference's numba functions in
 * Anis_TTF_rays.py ("ATR") compute, keeping the reference's evaluation order
 * (no FMA contraction: build with -ffp-contract=off) so that results agree with
 * the reference bit for bit on the same libm.  Pinning: tests/golden/ holds
 * fields, rays and node-level operator outputs produced by importing the real
 * reference in the build container (tests/golden/make_golden.py); the CPU test
 * suite checks this file against them.
 *
 * Each function cites the reference lines it follows.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static const double RAD2DEG = 180.0 / M_PI; /* numba math.degrees: x * (180/pi) */
static const double DEG2RAD = M_PI / 180.0; /* numba math.radians: x * (pi/180) */

/* Python float modulo for a positive divisor (numba real_divmod_func_body). */
static double pymod(double a, double w)
{
    double m = fmod(a, w);
    if (m != 0.0) {
        if (m < 0.0) m += w;
    } else {
        m = 0.0;
    }
    return m;
}

/* Diagnostic switch (default 0 = the reference's behaviour).  With 1 the narrow band is a
 * correct min-heap: the parent of k is k/2 (the reference's round(k/2) picks a cousin for
 * k = 3 mod 4, ATR:123) and a re-evaluated node is sifted down as well as up (the reference
 * only sifts up, ATR:141-175, although its update may INCREASE a value).  Used by the tests
 * to show that the CUDA path equals the reference algorithm with pops in true time order, and
 * that it differs from the reference itself only downstream of heap mis-orderings. */
static int g_true_heap = 0;
void ali_oracle_set_true_heap(int on) { g_true_heap = on; }
/* Diagnostic 2: the reference's own heap (quirks included) on the refined source levels and on the main grid
 * until a popped node is `stop_r` nodes (Chebyshev) from the source, a CORRECT heap from then on -- the point
 * where the CUDA path hands over from its sequential replica to the band march.  stop_r < 0: off. */
static int g_true_after = -1, g_src_z = 0, g_src_x = 0;
void ali_oracle_set_true_heap_after(int stop_r) { g_true_after = stop_r; }

/* Python round(k / 2) for a positive int k: round-half-to-even (ATR:123,135,160,172). */
static int half_round(int k)
{
    if (g_true_heap) return k >> 1;
    int h = k >> 1;
    if (k & 1) return (h & 1) ? h + 1 : h;
    return h;
}

/* ------------------------------------------------------------------------- */
/* A grid with its material arrays and march state.                           */
/* veln / vel_map are held as double AFTER whatever cast the reference applied */
/* (int32 truncation / float32 rounding on refined grids, ATR:1527-1529).      */
/* stif holds the five int64 entries converted to double (exact below 2^53).   */
/* ------------------------------------------------------------------------- */
typedef struct {
    int nz, nx;
    double *veln;
    int32_t *velpn;
    double *vel_map;
    double *stif;   /* [nz*nx*5] or NULL */
    int has_stif;   /* "stif_den is not None" in the reference */
    double *ttn;
    int32_t *nsts;
    int32_t *btg;   /* heap: (iz, ix) pairs, 1-indexed */
    int ntr;
    int own_mat;    /* material arrays owned by this grid */
} Grid;

typedef struct {
    const double *group; /* avlist2 / velocity_dat: [361][ncol] */
    const double *phase; /* phase_vel: [361][ncol] */
    int ncol;
} Tables;

#define AT(g, iz, ix) ((size_t)(iz) * (size_t)(g)->nx + (size_t)(ix))

/* ---- heap (ATR:94-237) --------------------------------------------------- */
static void sift_down_from(Grid *g, int tpp);
static void addtree(Grid *g, int iz, int ix)
{
    int tpc, tpp;
    g->ntr += 1;
    g->nsts[AT(g, iz, ix)] = g->ntr;
    g->btg[2 * g->ntr + 1] = ix;
    g->btg[2 * g->ntr + 0] = iz;
    tpc = g->ntr;
    tpp = half_round(tpc);
    while (tpp > 0) {
        int aa = g->btg[2 * tpp], bb = g->btg[2 * tpp + 1];
        if (g->ttn[AT(g, iz, ix)] < g->ttn[AT(g, aa, bb)]) {
            int e0, e1;
            g->nsts[AT(g, iz, ix)] = tpp;
            g->nsts[AT(g, aa, bb)] = tpc;
            e0 = g->btg[2 * tpc]; e1 = g->btg[2 * tpc + 1];
            g->btg[2 * tpc] = g->btg[2 * tpp]; g->btg[2 * tpc + 1] = g->btg[2 * tpp + 1];
            g->btg[2 * tpp] = e0; g->btg[2 * tpp + 1] = e1;
            tpc = tpp;
            tpp = half_round(tpc);
        } else {
            tpp = 0;
        }
    }
}

static void updtree(Grid *g, int iz, int ix)
{
    int tpc = g->nsts[AT(g, iz, ix)];
    int tpp = half_round(tpc);
    while (tpp > 0) {
        int aa = g->btg[2 * tpp], bb = g->btg[2 * tpp + 1];
        if (g->ttn[AT(g, iz, ix)] < g->ttn[AT(g, aa, bb)]) {
            int e0, e1;
            g->nsts[AT(g, iz, ix)] = tpp;
            g->nsts[AT(g, aa, bb)] = tpc;
            e0 = g->btg[2 * tpc]; e1 = g->btg[2 * tpc + 1];
            g->btg[2 * tpc] = g->btg[2 * tpp]; g->btg[2 * tpc + 1] = g->btg[2 * tpp + 1];
            g->btg[2 * tpp] = e0; g->btg[2 * tpp + 1] = e1;
            tpc = tpp;
            tpp = half_round(tpc);
        } else {
            tpp = 0;
        }
    }
    if (g_true_heap) sift_down_from(g, g->nsts[AT(g, iz, ix)]);
}

static void sift_down_from(Grid *g, int tpp)
{
    int ntr = g->ntr;
    for (;;) {
        int tpc = 2 * tpp, e0, e1;
        double rc, rp;
        if (tpc > ntr) break;
        if (tpc + 1 <= ntr && g->ttn[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])] >
                                  g->ttn[AT(g, g->btg[2 * tpc + 2], g->btg[2 * tpc + 3])])
            tpc += 1;
        rc = g->ttn[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])];
        rp = g->ttn[AT(g, g->btg[2 * tpp], g->btg[2 * tpp + 1])];
        if (!(rc < rp)) break;
        g->nsts[AT(g, g->btg[2 * tpp], g->btg[2 * tpp + 1])] = tpc;
        g->nsts[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])] = tpp;
        e0 = g->btg[2 * tpc]; e1 = g->btg[2 * tpc + 1];
        g->btg[2 * tpc] = g->btg[2 * tpp]; g->btg[2 * tpc + 1] = g->btg[2 * tpp + 1];
        g->btg[2 * tpp] = e0; g->btg[2 * tpp + 1] = e1;
        tpp = tpc;
    }
}

static void downtree(Grid *g)
{
    int tpp, tpc;
    int ntr = g->ntr;
    if (ntr == 1) {
        g->ntr = 0;
        return;
    }
    g->nsts[AT(g, g->btg[2 * ntr], g->btg[2 * ntr + 1])] = 1;
    g->btg[2] = g->btg[2 * ntr];
    g->btg[3] = g->btg[2 * ntr + 1];
    ntr -= 1;
    tpp = 1;
    tpc = 2;
    while (tpc < ntr) {
        double rd1 = g->ttn[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])];
        double rd2 = g->ttn[AT(g, g->btg[2 * (tpc + 1)], g->btg[2 * (tpc + 1) + 1])];
        if (rd1 > rd2) tpc += 1;
        rd1 = g->ttn[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])];
        rd2 = g->ttn[AT(g, g->btg[2 * tpp], g->btg[2 * tpp + 1])];
        if (rd1 < rd2) {
            int e0, e1;
            g->nsts[AT(g, g->btg[2 * tpp], g->btg[2 * tpp + 1])] = tpc;
            g->nsts[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])] = tpp;
            e0 = g->btg[2 * tpc]; e1 = g->btg[2 * tpc + 1];
            g->btg[2 * tpc] = g->btg[2 * tpp]; g->btg[2 * tpc + 1] = g->btg[2 * tpp + 1];
            g->btg[2 * tpp] = e0; g->btg[2 * tpp + 1] = e1;
            tpp = tpc;
            tpc = 2 * tpp;
        } else {
            tpc = ntr + 1;
        }
    }
    if (tpc == ntr) {
        double rd1 = g->ttn[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])];
        double rd2 = g->ttn[AT(g, g->btg[2 * tpp], g->btg[2 * tpp + 1])];
        if (rd1 < rd2) {
            int e0, e1;
            g->nsts[AT(g, g->btg[2 * tpp], g->btg[2 * tpp + 1])] = tpc;
            g->nsts[AT(g, g->btg[2 * tpc], g->btg[2 * tpc + 1])] = tpp;
            e0 = g->btg[2 * tpc]; e1 = g->btg[2 * tpc + 1];
            g->btg[2 * tpc] = g->btg[2 * tpp]; g->btg[2 * tpc + 1] = g->btg[2 * tpp + 1];
            g->btg[2 * tpp] = e0; g->btg[2 * tpp + 1] = e1;
        }
    }
    g->ntr = ntr;
}

/* ---- velocities ----------------------------------------------------------- */
/* 1-degree table interpolation (ATR:1371-1375, 1559-1563, 2951-2954). */
static double table_vel(const double *tab, int ncol, double eff, int col, double vm)
{
    int a1 = (int)floor(eff);
    int a2 = (a1 + 1) % 180;
    double rem = eff - a1;
    return vm * ((1 - rem) * tab[(size_t)a1 * ncol + col] + rem * tab[(size_t)a2 * ncol + col]);
}

/* Christoffel group velocity, stiffness in MPa (ATR:3542-3558 and its inlined copies). */
static double christoffel_group(double eff, const double *s, double vm)
{
    double m90 = pymod(eff, 90.0);
    if (m90 < 0.01 || m90 > 90 - 0.01) {
        double lam;
        if (fabs(pymod(eff, 180.0) - 90) < 1) lam = s[2]; else lam = s[0];
        return 1000 * vm * sqrt(lam / s[4]);
    } else {
        double c22 = s[0], c23 = s[1], c33 = s[2], c44 = s[3];
        double t = tan(DEG2RAD * eff);
        double A = c22 + c33 - 2 * c44;
        double B = (c23 + c44) * (t - 1 / t);
        double C = c22 - c33;
        double disc = sqrt(B * B + A * A - C * C);
        double ph;
        double lam;
        if (eff < 90)
            ph = pymod(atan((-B - disc) / (C - A)), M_PI);
        else
            ph = pymod(atan((-B + disc) / (C - A)), M_PI);
        lam = 0.5 * (cos(2 * ph) * (c22 - c44) + sin(2 * ph) * (c23 + c44) * t + c22 + c44);
        return 1000 * vm * sqrt(lam / s[4]) / cos(DEG2RAD * eff - ph);
    }
}

/* Christoffel phase velocity (ATR:1400-1406). */
static double christoffel_phase(double eff, const double *s, double vm)
{
    double c = cos(DEG2RAD * eff);
    double sn = sin(DEG2RAD * eff);
    double A = c * c * s[0] + sn * sn * s[3];
    double B = c * sn * (s[1] + s[3]);
    double C = c * c * s[3] + sn * sn * s[2];
    return 1000 * vm * sqrt((A + C + sqrt((A - C) * (A - C) + 4 * (B * B))) / (2 * s[4]));
}

static double group_velocity_at(const Grid *g, const Tables *t, int iz, int ix, double eff)
{
    size_t p = AT(g, iz, ix);
    if (g->velpn[p] != 0 || !g->has_stif)
        return table_vel(t->group, t->ncol, eff, g->velpn[p], g->vel_map[p]);
    return christoffel_group(eff, g->stif + 5 * p, g->vel_map[p]);
}

/* ---- wavefront_angle_dist (ATR:1413-1460) --------------------------------- */
static void wavefront_angle_dist(int ix, int iz, int x1, int x2, int x3, int z1, int z2, int z3,
                                 double y1, double y2, double y3, double *angle, double *dist)
{
    double a, xpos, zpos, dx, dz;
    if (y3 != y1) {
        a = (y2 - y1) / (y3 - y1);
    } else {
        *angle = 0.0;
        *dist = -1.0;
        return;
    }
    xpos = (1 - a) * x1 + a * x3;
    zpos = (1 - a) * z1 + a * z3;
    dx = x2 - xpos;
    dz = z2 - zpos;
    if (dx == 0)
        *angle = 0.0;
    else
        *angle = pymod(RAD2DEG * atan(dz / dx) + 90, 180.0);
    *dist = fabs(dz * (x2 - ix) - dx * (z2 - iz)) / sqrt(dx * dx + dz * dz);
}

/* ---- update (ATR:904-1410) ------------------------------------------------ */
/* nnz / nnx are the LOGICAL extents the caller passes (the reference passes a
 * wrong nnz once, ATR:1645); reads outside the real array are treated as far. */
static int st_ok(const Grid *g, int iz, int ix)
{
    if (iz < 0 || ix < 0 || iz >= g->nz || ix >= g->nx) return 0;
    return g->nsts[AT(g, iz, ix)] >= 0;
}
#define TT(dz, dx) (g->ttn[AT(g, iz + (dz), ix + (dx))])
#define WAD(xa, xb, xc, za, zb, zc, ya, yb, yc) \
    wavefront_angle_dist(ix, iz, xa, xb, xc, za, zb, zc, ya, yb, yc, &angle, &dist)

static double ali_update(const Grid *g, const Tables *t, int iz, int ix, double dnx, int nnz, int nnx,
                         int *stencil_out)
{
    int sp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int stencil_no = -1;
    double min_diff = 1000000.0, diff;
    double angle = 0.0, dist = -1.0, wt = 0.0;
    const double r2 = sqrt(2.0);

    if (ix > 1) { if (st_ok(g, iz, ix - 2)) sp[3] += 1; }
    if (ix > 0) {
        if (st_ok(g, iz, ix - 1)) { sp[4] += 1; sp[7] += 1; }
        if (iz > 0) { if (st_ok(g, iz - 1, ix - 1)) { sp[0] += 1; sp[3] += 1; sp[4] += 1; } }
        if (iz < nnz - 1) { if (st_ok(g, iz + 1, ix - 1)) { sp[2] += 1; sp[3] += 1; sp[7] += 1; } }
    }
    if (ix < nnx - 2) { if (st_ok(g, iz, ix + 2)) sp[1] += 1; }
    if (ix < nnx - 1) {
        if (st_ok(g, iz, ix + 1)) { sp[5] += 1; sp[6] += 1; }
        if (iz > 0) { if (st_ok(g, iz - 1, ix + 1)) { sp[0] += 1; sp[1] += 1; sp[5] += 1; } }
        if (iz < nnz - 1) { if (st_ok(g, iz + 1, ix + 1)) { sp[1] += 1; sp[2] += 1; sp[6] += 1; } }
    }
    if (iz > 1) { if (st_ok(g, iz - 2, ix)) sp[0] += 1; }
    if (iz > 0) { if (st_ok(g, iz - 1, ix)) { sp[4] += 1; sp[5] += 1; } }
    if (iz < nnz - 2) { if (st_ok(g, iz + 2, ix)) sp[2] += 1; }
    if (iz < nnz - 1) { if (st_ok(g, iz + 1, ix)) { sp[6] += 1; sp[7] += 1; } }

    if (sp[0] == 3) { diff = fabs(TT(-1, -1) - TT(-1, 1)); if (diff < min_diff) { stencil_no = 0; min_diff = diff; } }
    if (sp[1] == 3) { diff = fabs(TT(-1, 1) - TT(1, 1));   if (diff < min_diff) { stencil_no = 1; min_diff = diff; } }
    if (sp[2] == 3) { diff = fabs(TT(1, -1) - TT(1, 1));   if (diff < min_diff) { stencil_no = 2; min_diff = diff; } }
    if (sp[3] == 3) { diff = fabs(TT(-1, -1) - TT(1, -1)); if (diff < min_diff) { stencil_no = 3; min_diff = diff; } }
    if (sp[4] == 3) { diff = fabs(TT(0, -1) - TT(-1, 0));  if (diff < min_diff) { stencil_no = 4; min_diff = diff; } }
    if (sp[5] == 3) { diff = fabs(TT(-1, 0) - TT(0, 1));   if (diff < min_diff) { stencil_no = 5; min_diff = diff; } }
    if (sp[6] == 3) { diff = fabs(TT(1, 0) - TT(0, 1));    if (diff < min_diff) { stencil_no = 6; min_diff = diff; } }
    if (sp[7] == 3) { diff = fabs(TT(0, -1) - TT(1, 0));   if (diff < min_diff) { stencil_no = 7; min_diff = diff; } }

    if (stencil_no != -1) {
        switch (stencil_no) {
        case 0: /* ATR:1040-1058 */
            if (TT(-1, -1) < TT(-1, 1)) { WAD(ix, ix - 1, ix + 1, iz - 2, iz - 1, iz - 1, TT(-2, 0), TT(-1, -1), TT(-1, 1)); wt = TT(-1, -1); }
            else                        { WAD(ix, ix + 1, ix - 1, iz - 2, iz - 1, iz - 1, TT(-2, 0), TT(-1, 1), TT(-1, -1)); wt = TT(-1, 1); }
            break;
        case 1: /* ATR:1059-1077 */
            if (TT(-1, 1) < TT(1, 1)) { WAD(ix + 2, ix + 1, ix + 1, iz, iz - 1, iz + 1, TT(0, 2), TT(-1, 1), TT(1, 1)); wt = TT(-1, 1); }
            else                      { WAD(ix + 2, ix + 1, ix + 1, iz, iz + 1, iz - 1, TT(0, 2), TT(1, 1), TT(-1, 1)); wt = TT(1, 1); }
            break;
        case 2: /* ATR:1078-1096 */
            if (TT(1, -1) < TT(1, 1)) { WAD(ix, ix - 1, ix + 1, iz + 2, iz + 1, iz + 1, TT(2, 0), TT(1, -1), TT(1, 1)); wt = TT(1, -1); }
            else                      { WAD(ix, ix + 1, ix - 1, iz + 2, iz + 1, iz + 1, TT(2, 0), TT(1, 1), TT(1, -1)); wt = TT(1, 1); }
            break;
        case 3: /* ATR:1097-1115 */
            if (TT(-1, -1) < TT(1, -1)) { WAD(ix - 2, ix - 1, ix - 1, iz, iz - 1, iz + 1, TT(0, -2), TT(-1, -1), TT(1, -1)); wt = TT(-1, -1); }
            else                        { WAD(ix - 2, ix - 1, ix - 1, iz, iz + 1, iz - 1, TT(0, -2), TT(1, -1), TT(-1, -1)); wt = TT(1, -1); }
            break;
        case 4: /* ATR:1116-1122 */
            if (TT(0, -1) < TT(-1, 0)) { WAD(ix - 1, ix - 1, ix, iz - 1, iz, iz - 1, TT(-1, -1), TT(0, -1), TT(-1, 0)); wt = TT(0, -1); }
            else                       { WAD(ix - 1, ix, ix - 1, iz - 1, iz - 1, iz, TT(-1, -1), TT(-1, 0), TT(0, -1)); wt = TT(-1, 0); }
            break;
        case 5: /* ATR:1123-1129 */
            if (TT(-1, 0) < TT(0, 1)) { WAD(ix + 1, ix, ix + 1, iz - 1, iz - 1, iz, TT(-1, 1), TT(-1, 0), TT(0, 1)); wt = TT(-1, 0); }
            else                      { WAD(ix + 1, ix + 1, ix, iz - 1, iz, iz - 1, TT(-1, 1), TT(0, 1), TT(-1, 0)); wt = TT(0, 1); }
            break;
        case 6: /* ATR:1130-1136 */
            if (TT(1, 0) < TT(0, 1)) { WAD(ix + 1, ix, ix + 1, iz + 1, iz + 1, iz, TT(1, 1), TT(1, 0), TT(0, 1)); wt = TT(1, 0); }
            else                     { WAD(ix + 1, ix + 1, ix, iz + 1, iz, iz + 1, TT(1, 1), TT(0, 1), TT(1, 0)); wt = TT(0, 1); }
            break;
        default: /* 7, ATR:1137-1143 */
            if (TT(0, -1) < TT(1, 0)) { WAD(ix - 1, ix - 1, ix, iz + 1, iz, iz + 1, TT(1, -1), TT(0, -1), TT(1, 0)); wt = TT(0, -1); }
            else                      { WAD(ix - 1, ix, ix - 1, iz + 1, iz + 1, iz, TT(1, -1), TT(1, 0), TT(0, -1)); wt = TT(1, 0); }
            break;
        }
    }

    if (stencil_no == -1 || ix == 0 || ix == nnx - 1 || iz == 0 || iz == nnz - 1) { /* ATR:1146 */
        int q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const double w1 = r2 - 1, w2 = 2 - r2;
        if (ix > 1) { if (st_ok(g, iz, ix - 2)) { q[4] += 1; q[7] += 1; } }
        if (ix > 0) {
            if (st_ok(g, iz, ix - 1)) { q[4] += 1; q[7] += 1; }
            if (iz > 0) { if (st_ok(g, iz - 1, ix - 1)) { q[2] += 1; q[7] += 1; } }
            if (iz < nnz - 1) { if (st_ok(g, iz + 1, ix - 1)) { q[3] += 1; q[4] += 1; } }
        }
        if (ix < nnx - 2) { if (st_ok(g, iz, ix + 2)) { q[5] += 1; q[6] += 1; } }
        if (ix < nnx - 1) {
            if (st_ok(g, iz, ix + 1)) { q[5] += 1; q[6] += 1; }
            if (iz > 0) { if (st_ok(g, iz - 1, ix + 1)) { q[1] += 1; q[6] += 1; } }
            if (iz < nnz - 1) { if (st_ok(g, iz + 1, ix + 1)) { q[0] += 1; q[5] += 1; } }
        }
        if (iz > 1) { if (st_ok(g, iz - 2, ix)) { q[1] += 1; q[2] += 1; } }
        if (iz > 0) { if (st_ok(g, iz - 1, ix)) { q[1] += 1; q[2] += 1; } }
        if (iz < nnz - 2) { if (st_ok(g, iz + 2, ix)) { q[0] += 1; q[3] += 1; } }
        if (iz < nnz - 1) { if (st_ok(g, iz + 1, ix)) { q[0] += 1; q[3] += 1; } }

        if (stencil_no == -1) min_diff = 1000000.0;
        stencil_no = -2;
#define TRI(k, az, ax, bz, bx, cz, cx)                                              \
        if (q[k] == 3) {                                                            \
            if (TT(az, ax) < fmin(TT(bz, bx), TT(cz, cx))) {                        \
                diff = fabs(w1 * TT(az, ax) + w2 * TT(bz, bx) - TT(cz, cx));        \
                if (diff < min_diff) { stencil_no = k; min_diff = diff; }           \
            }                                                                       \
        }
        TRI(0, 2, 0, 1, 0, 1, 1)
        TRI(1, -2, 0, -1, 0, -1, 1)
        TRI(2, -2, 0, -1, 0, -1, -1)
        TRI(3, 2, 0, 1, 0, 1, -1)
        TRI(4, 0, -2, 0, -1, 1, -1)
        TRI(5, 0, 2, 0, 1, 1, 1)
        TRI(6, 0, 2, 0, 1, -1, 1)
        TRI(7, 0, -2, 0, -1, -1, -1)
#undef TRI
        if (stencil_no != -2) {
            switch (stencil_no) {
            case 0: /* ATR:1263-1274 */
                if (TT(1, 0) < TT(1, 1)) {
                    if (ix == 0) { angle = 90.; dist = 1.; }
                    else WAD(ix, ix, ix + 1, iz + 2, iz + 1, iz + 1, TT(2, 0), TT(1, 0), TT(1, 1));
                } else {
                    WAD(ix, ix + 1, ix, iz + 2, iz + 1, iz + 1, TT(2, 0), TT(1, 1), TT(1, 0));
                }
                wt = TT(1, 1);
                break;
            case 1: /* ATR:1275-1287 */
                if (TT(-1, 0) < TT(-1, 1)) {
                    if (ix == 0) { angle = 90.; dist = 1.; }
                    else WAD(ix, ix, ix + 1, iz - 2, iz - 1, iz - 1, TT(-2, 0), TT(-1, 0), TT(-1, 1));
                    wt = TT(-1, 0);
                } else {
                    WAD(ix, ix + 1, ix, iz - 2, iz - 1, iz - 1, TT(-2, 0), TT(-1, 1), TT(-1, 0));
                    wt = TT(-1, 1);
                }
                break;
            case 2: /* ATR:1288-1300 */
                if (TT(-1, 0) < TT(-1, -1)) {
                    if (ix == nnx - 1) { angle = 90.; dist = 1.; }
                    else WAD(ix, ix, ix - 1, iz - 2, iz - 1, iz - 1, TT(-2, 0), TT(-1, 0), TT(-1, -1));
                    wt = TT(-1, 0);
                } else {
                    WAD(ix, ix - 1, ix, iz - 2, iz - 1, iz - 1, TT(-2, 0), TT(-1, -1), TT(-1, 0));
                    wt = TT(-1, -1);
                }
                break;
            case 3: /* ATR:1301-1313 */
                if (TT(1, 0) < TT(1, -1)) {
                    if (ix == nnx - 1) { angle = 90.; dist = 1.; }
                    else WAD(ix, ix, ix - 1, iz + 2, iz + 1, iz + 1, TT(2, 0), TT(1, 0), TT(1, -1));
                    wt = TT(1, 0);
                } else {
                    WAD(ix, ix - 1, ix, iz + 2, iz + 1, iz + 1, TT(2, 0), TT(1, -1), TT(1, 0));
                    wt = TT(1, -1);
                }
                break;
            case 4: /* ATR:1314-1326 */
                if (TT(0, -1) < TT(1, -1)) {
                    if (iz == 0) { angle = 0.; dist = 1.; }
                    else WAD(ix - 2, ix - 1, ix - 1, iz, iz, iz + 1, TT(0, -2), TT(0, -1), TT(1, -1));
                    wt = TT(0, -1);
                } else {
                    WAD(ix - 2, ix - 1, ix - 1, iz, iz + 1, iz, TT(0, -2), TT(1, -1), TT(0, -1));
                    wt = TT(1, -1);
                }
                break;
            case 5: /* ATR:1327-1339 */
                if (TT(0, 1) < TT(1, 1)) {
                    if (iz == 0) { angle = 0.; dist = 1.; }
                    else WAD(ix + 2, ix + 1, ix + 1, iz, iz, iz + 1, TT(0, 2), TT(0, 1), TT(1, 1));
                    wt = TT(0, 1);
                } else {
                    WAD(ix + 2, ix + 1, ix + 1, iz, iz + 1, iz, TT(0, 2), TT(1, 1), TT(0, 1));
                    wt = TT(1, 1);
                }
                break;
            case 6: /* ATR:1340-1352 */
                if (TT(0, 1) < TT(-1, 1)) {
                    if (iz == nnz - 1) { angle = 0.; dist = 1.; }
                    else WAD(ix + 2, ix + 1, ix + 1, iz, iz, iz - 1, TT(0, 2), TT(0, 1), TT(-1, 1));
                    wt = TT(0, 1);
                } else {
                    WAD(ix + 2, ix + 1, ix + 1, iz, iz - 1, iz, TT(0, 2), TT(-1, 1), TT(0, 1));
                    wt = TT(-1, 1);
                }
                break;
            default: /* 7, ATR:1353-1365 */
                if (TT(0, -1) < TT(-1, -1)) {
                    if (iz == nnz - 1) { angle = 0.; dist = 1.; }
                    else WAD(ix - 2, ix - 1, ix - 1, iz, iz, iz - 1, TT(0, -2), TT(0, -1), TT(-1, -1));
                    wt = TT(0, -1);
                } else {
                    WAD(ix - 2, ix - 1, ix - 1, iz, iz - 1, iz, TT(0, -2), TT(-1, -1), TT(0, -1));
                    wt = TT(-1, -1);
                }
                break;
            }
            stencil_no += 8;
        }
    }
    if (stencil_out) *stencil_out = stencil_no;
    if (dist != -1.0) {
        size_t p = AT(g, iz, ix);
        double eff = pymod(g->veln[p] - angle, 180.0);
        double vel;
        if (g->velpn[p] != 0 || !g->has_stif)
            vel = table_vel(t->phase, t->ncol, eff, g->velpn[p], g->vel_map[p]);
        else
            vel = christoffel_phase(eff, g->stif + 5 * p, g->vel_map[p]);
        return wt + (dist * dnx / vel);
    }
    return -1.0;
}
#undef WAD

/* ---- fouds18_A (ATR:240-901) ---------------------------------------------- */
#define NS(k, j) (g->nsts[AT(g, k, j)])
#define TN(k, j) (g->ttn[AT(g, k, j)])
static double ali_fouds18(const Grid *g, const Tables *t, int iz, int ix, double dnx, double dnz, int nnx,
                          int nnz)
{
    size_t p = AT(g, iz, ix);
    int tsw1 = 0, tsw2 = 0, tsw3 = 0, tsw4 = 0;
    double travm = 0, travmd = 0, travmt = 0, travms = 0;
    double wave_ang, eff, slown, mf2;
    double a = 0, b = 0, c = 0, tref = 0, tdiv = 1, u, v, em, rd1, tdsh, trav;
    int jn, kn, lp;
    (void)v;

    /* 0-degree stencil, ATR:281-459 */
    wave_ang = 0;
    eff = pymod(wave_ang - g->veln[p], 180.0);
    slown = 1.0 / group_velocity_at(g, t, iz, ix, eff);
    for (jn = 0; jn < 2; jn++) {
        int j = jn == 0 ? ix - 1 : ix + 1;
        int j2 = 0, swj;
        if (!(0 <= j && j <= nnx - 1)) continue;
        swj = -1;
        if (j == ix - 1) { j2 = j - 1; if (j2 >= 0) { if (NS(iz, j2) == 0) swj = 0; } }
        else             { j2 = j + 1; if (j2 <= nnx - 1) { if (NS(iz, j2) == 0) swj = 0; } }
        if (NS(iz, j) == 0 && swj == 0) { swj = -1; if (TN(iz, j) >= TN(iz, j2)) swj = 0; }
        else swj = -1;
        for (kn = 0; kn < 2; kn++) {
            int k = kn == 0 ? iz - 1 : iz + 1;
            int k2 = 0, swk, swsol;
            if (!(0 <= k && k <= nnz - 1)) continue;
            swk = -1;
            if (k == iz - 1) { k2 = k - 1; if (k2 >= 0) { if (NS(k2, ix) == 0) swk = 0; } }
            else             { k2 = k + 1; if (k2 <= nnz - 1) { if (NS(k2, ix) == 0) swk = 0; } }
            if (NS(k, ix) == 0 && swk == 0) { swk = -1; if (TN(k, ix) >= TN(k2, ix)) swk = 0; }
            else swk = -1;
            swsol = 0;
            if (swj == 0) {
                swsol = 1;
                if (swk == 0) {
                    double e1 = 4.0 * TN(iz, j) - TN(iz, j2), e2 = 4.0 * TN(k, ix) - TN(k2, ix);
                    u = 2.0 * dnx; a = 18;
                    b = -6 * (4.0 * TN(iz, j) - TN(iz, j2) + 4.0 * TN(k, ix) - TN(k2, ix));
                    c = e1 * e1 + e2 * e2 - 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else if (NS(k, ix) == 0) {
                    double e1 = 3.0 * TN(k, ix), e2 = 4.0 * TN(iz, j) - TN(iz, j2);
                    v = 2.0 * dnx; a = 18;
                    b = -6.0 * (3.0 * TN(k, ix) + 4.0 * TN(iz, j) - TN(iz, j2));
                    c = e1 * e1 + e2 * e2 - 4 * (v * v) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else {
                    u = 2.0 * dnx; a = 1.0; b = 0.0;
                    c = -(u * u) * (slown * slown);
                    tref = 4.0 * TN(iz, j) - TN(iz, j2);
                    tdiv = 3.0;
                    tdiv = 1.0; /* ATR:395 overrides ATR:389 */
                }
            } else if (NS(iz, j) == 0) {
                swsol = 1;
                if (swk == 0) {
                    double e1 = 3.0 * TN(iz, j), e2 = 4.0 * TN(k, ix) - TN(k2, ix);
                    u = dnx;
                    em = 3.0 * TN(iz, j) + 4.0 * TN(k, ix) - TN(k2, ix);
                    a = 18; b = -6.0 * em;
                    c = e1 * e1 + e2 * e2 - 3 * 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else if (NS(k, ix) == 0) {
                    double e3 = u = dnx;
                    e3 = u * slown;
                    a = 2; b = -2 * (TN(k, ix) + TN(iz, j));
                    c = TN(k, ix) * TN(k, ix) + TN(iz, j) * TN(iz, j) - e3 * e3;
                    tref = 0.0; tdiv = 1.0;
                } else {
                    double e3 = TN(iz, j) + slown * dnx;
                    a = 1.0; b = 0.0; c = -(e3 * e3);
                    tref = 0.0; tdiv = 1.0;
                }
            } else {
                if (swk == 0) {
                    swsol = 1;
                    u = 2.0 * dnz; a = 1.0; b = 0.0;
                    c = -(u * u) * (slown * slown);
                    tref = 4.0 * TN(k, ix) - TN(k2, ix);
                    tdiv = 3.0;
                } else if (NS(k, ix) == 0) {
                    double e3 = TN(k, ix) + slown * dnz;
                    swsol = 1;
                    a = 1.0; b = 0.0; c = -(e3 * e3);
                    tref = 0.0; tdiv = 1.0;
                }
            }
            if (swsol == 1) {
                rd1 = b * b - 4.0 * a * c;
                if (rd1 < 0) rd1 = 0;
                tdsh = (-b + sqrt(rd1)) / (2.0 * a);
                trav = (tref + tdsh) / tdiv;
                if (tsw1 == 1) travm = fmin(trav, travm);
                else { travm = trav; tsw1 = 1; }
            }
        }
    }

    /* 45-degree stencil, ATR:467-696 */
    wave_ang = 45;
    eff = rint(pymod(wave_ang - g->veln[p], 180.0));
    slown = 1.0 / group_velocity_at(g, t, iz, ix, eff);
    mf2 = sqrt(2.0);
    for (jn = 0; jn < 2; jn++) {
        int j = jn == 0 ? ix - 1 : ix + 1;
        int k = jn == 0 ? iz + 1 : iz - 1;
        int j2 = 0, k2 = 0, swdiag, jjn;
        if (!(0 <= j && j <= nnx - 1 && 0 <= k && k <= nnz - 1)) continue;
        swdiag = -1;
        if (j == ix - 1) { j2 = j - 1; k2 = k + 1; if (j2 >= 0 && k2 <= nnz - 1) { if (NS(k2, j2) == 0) swdiag = 0; } }
        else             { j2 = j + 1; k2 = k - 1; if (j2 <= nnx - 1 && k2 >= 0) { if (NS(k2, j2) == 0) swdiag = 0; } }
        if (NS(k, j) == 0 && swdiag == 0) { swdiag = -1; if (TN(k, j) >= TN(k2, j2)) swdiag = 0; }
        else swdiag = -1;
        for (jjn = 0; jjn < 2; jjn++) {
            int jj = jjn == 0 ? ix - 1 : ix + 1;
            int kk = jjn == 0 ? iz - 1 : iz + 1;
            int jj2 = 0, kk2 = 0, swskew, swsol;
            if (!(0 <= jj && jj <= nnx - 1 && 0 <= kk && kk <= nnz - 1)) continue;
            swskew = -1;
            if (jj == ix - 1) { jj2 = jj - 1; kk2 = kk - 1; if (jj2 >= 0 && kk2 >= 0) { if (NS(kk2, jj2) == 0) swskew = 0; } }
            else              { jj2 = jj + 1; kk2 = kk + 1; if (jj2 <= nnx - 1 && kk2 <= nnz - 1) { if (NS(kk2, jj2) == 0) swskew = 0; } }
            if (NS(kk, jj) == 0 && swskew == 0) { swskew = -1; if (TN(kk, jj) >= TN(kk2, jj2)) swskew = 0; }
            else swskew = -1;
            swsol = 0;
            if (swdiag == 0) {
                swsol = 1;
                if (swskew == 0) {
                    double e1 = 4.0 * TN(k, j) - TN(k2, j2), e2 = 4.0 * TN(kk, jj) - TN(kk2, jj2);
                    u = 2.0 * mf2 * dnx; a = 18.0;
                    b = -6.0 * (4.0 * TN(k, j) - TN(k2, j2) + 4.0 * TN(kk, jj) - TN(kk2, jj2));
                    c = e1 * e1 + e2 * e2 - 4 * (u * u) * (slown * slown);
                    tref = 0; tdiv = 1.0;
                } else if (NS(kk, jj) == 0) {
                    double e1 = 3.0 * TN(kk, jj), e2 = 4.0 * TN(k, j) - TN(k2, j2);
                    v = 2.0 * mf2 * dnx; a = 18;
                    b = -6.0 * (3.0 * TN(kk, jj) + 4.0 * TN(k, j) - TN(k2, j2));
                    c = e1 * e1 + e2 * e2 - 4 * (v * v) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else {
                    double e3;
                    u = mf2 * 2.0 * dnx; a = 1.0; b = 0.0;
                    e3 = u * slown;
                    c = -1.0 * (e3 * e3);
                    tref = (4.0 * TN(k, j) - TN(k2, j2));
                    tdiv = 3.0;
                }
            } else if (NS(k, j) == 0) {
                swsol = 1;
                if (swskew == 0) {
                    double e1 = 3.0 * TN(k, j), e2 = 4.0 * TN(kk, jj) - TN(kk2, jj2);
                    u = mf2 * dnx;
                    em = 3.0 * TN(k, j) + 4.0 * TN(kk, jj) - TN(kk2, jj2);
                    a = 18; b = -6.0 * em;
                    c = e1 * e1 + e2 * e2 - 3 * 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else if (NS(kk, jj) == 0) {
                    double e3;
                    u = mf2 * dnx;
                    e3 = u * slown;
                    a = 2; b = -2 * (TN(kk, jj) + TN(k, j));
                    c = TN(kk, jj) * TN(kk, jj) + TN(k, j) * TN(k, j) - 4.0 / 9.0 * (e3 * e3);
                    tref = 0.0; tdiv = 1.0;
                } else {
                    double e3;
                    u = mf2 * dnx;
                    e3 = TN(k, j) + slown * u;
                    a = 1.0; b = 0.0; c = -(e3 * e3);
                    tref = 0; tdiv = 1.0;
                }
            } else {
                if (swskew == 0) {
                    swsol = 1;
                    u = 2.0 * mf2 * dnz; a = 1.0; b = 0.0;
                    c = -(u * u) * (slown * slown);
                    tref = 4.0 * TN(kk, jj) - TN(kk2, jj2);
                    tdiv = 3.0;
                } else if (NS(kk, jj) == 0) {
                    swsol = 1;
                    u = mf2 * dnx; a = 1.0; b = 0.0;
                    c = -(slown * slown) * (u * u);
                    tref = TN(kk, jj);
                    tdiv = 1.0;
                }
            }
            if (swsol == 1) {
                rd1 = b * b - 4.0 * a * c;
                if (rd1 > 0) {
                    tdsh = (-b + sqrt(rd1)) / (2.0 * a);
                    trav = (tref + tdsh) / tdiv;
                    if (tsw2 == 1) travmd = fmin(trav, travmd);
                    else { travmd = trav; tsw2 = 1; }
                }
            }
        }
    }
    if (travmd != 0) travmd = fmin(travm, travmd);
    else travmd = travm;

    /* 26.6 / 63.4-degree stencils, ATR:698-897 */
    wave_ang = rint(RAD2DEG * atan(0.5));
    for (lp = 0; lp < 2; lp++) {
        int jv[5], kv[5], q;
        int *tsw = lp == 0 ? &tsw3 : &tsw4;
        double *acc = lp == 0 ? &travmt : &travms;
        if (lp == 0) {
            int jv0[5] = {ix - 1, ix + 2, ix + 1, ix - 2, ix - 1};
            int kv0[5] = {iz - 2, iz - 1, iz + 2, iz + 1, iz - 2};
            memcpy(jv, jv0, sizeof jv); memcpy(kv, kv0, sizeof kv);
            eff = pymod(-wave_ang - g->veln[p], 180.0);
        } else {
            int jv1[5] = {ix + 1, ix + 2, ix - 1, ix - 2, ix + 1};
            int kv1[5] = {iz - 2, iz + 1, iz + 2, iz - 1, iz - 2};
            memcpy(jv, jv1, sizeof jv); memcpy(kv, kv1, sizeof kv);
            eff = pymod(wave_ang - g->veln[p], 180.0);
        }
        slown = 1.0 / group_velocity_at(g, t, iz, ix, eff);
        mf2 = sqrt(5.0);
        for (q = 0; q < 4; q++) {
            int j = jv[q], k = kv[q], jj = jv[q + 1], kk = kv[q + 1];
            int swsol = 0;
            if (!(0 <= j && j <= nnx - 1)) continue;
            if (!(0 <= k && k <= nnz - 1)) continue;
            if (!(0 <= jj && jj <= nnx - 1)) continue;
            if (!(0 <= kk && kk <= nnz - 1)) continue;
            if (NS(k, j) == 0) {
                swsol = 1;
                if (NS(kk, jj) == 0) {
                    double e3;
                    u = mf2 * dnx;
                    e3 = u * slown;
                    a = 2; b = -2 * (TN(kk, jj) + TN(k, j));
                    c = TN(kk, jj) * TN(kk, jj) + TN(k, j) * TN(k, j) - 2 * (e3 * e3);
                    tref = 0.0;
                } else {
                    double e3;
                    u = mf2 * dnx;
                    e3 = slown * u;
                    a = 1; b = 0; c = -(e3 * e3);
                    tref = TN(k, j);
                }
            } else if (NS(kk, jj) == 0) {
                double e3;
                swsol = 1;
                u = mf2 * dnx;
                e3 = slown * u;
                a = 1; b = 0; c = -(e3 * e3);
                tref = TN(kk, jj);
            }
            if (swsol == 1) {
                rd1 = b * b - 4 * a * c;
                if (rd1 < 0) rd1 = 0;
                tdsh = (-b + sqrt(rd1)) / (2.0 * a);
                trav = tref + tdsh;
                if (*tsw == 1) *acc = fmin(trav, *acc);
                else { *acc = trav; *tsw = 1; }
            }
        }
        if (lp == 0) {
            if (travmt != 0) travmt = fmin(travmt, travmd);
            else travmt = travmd;
        } else {
            if (travms != 0) travms = fmin(travmt, travms);
            else travms = travmt;
        }
    }
    if (TN(iz, ix) != 0) travms = fmin(travms, TN(iz, ix));
    return travms;
}
#undef NS
#undef TN
#undef TT

/* ---- grid refinement (ATR:26-91) ------------------------------------------ */
/* Builds the refined material arrays of the window [z0..z1] x [x0..x1] of src. */
static void grid_free(Grid *g)
{
    if (g->own_mat) { free(g->veln); free(g->velpn); free(g->vel_map); free(g->stif); }
    free(g->ttn); free(g->nsts); free(g->btg);
    memset(g, 0, sizeof *g);
}

static void grid_alloc_state(Grid *g, int zero_ttn)
{
    size_t n = (size_t)g->nz * g->nx, i;
    if (zero_ttn) g->ttn = (double *)calloc(n, sizeof(double));
    g->nsts = (int32_t *)malloc(n * sizeof(int32_t));
    for (i = 0; i < n; i++) g->nsts[i] = -1;
    g->btg = (int32_t *)calloc(2 * (n + 2), sizeof(int32_t));
    g->ntr = 0;
}

static void refine(const Grid *src, int z0, int z1, int x0, int x1, int scale, Grid *dst)
{
    int wz = z1 - z0 + 1, wx = x1 - x0 + 1;
    int nz = scale * (wz - 1) + 1, nx = scale * (wx - 1) + 1;
    int side = (scale - 1) / 2, i, j;
    size_t n = (size_t)nz * nx;
    memset(dst, 0, sizeof *dst);
    dst->nz = nz; dst->nx = nx; dst->own_mat = 1; dst->has_stif = src->has_stif;
    dst->veln = (double *)malloc(n * sizeof(double));
    dst->velpn = (int32_t *)malloc(n * sizeof(int32_t));
    dst->vel_map = (double *)malloc(n * sizeof(double));
    dst->stif = src->stif ? (double *)malloc(n * 5 * sizeof(double)) : NULL;
    for (i = 0; i < nz; i++) {
        int ci = z0 + (i + side) / scale;
        for (j = 0; j < nx; j++) {
            int cj = x0 + (j + side) / scale;
            size_t s = AT(src, ci, cj), d = (size_t)i * nx + j;
            dst->veln[d] = (double)(int32_t)src->veln[s];      /* int32 truncation */
            dst->velpn[d] = src->velpn[s];
            dst->vel_map[d] = (double)(float)src->vel_map[s];  /* float32 rounding */
            if (dst->stif) memcpy(dst->stif + 5 * d, src->stif + 5 * s, 5 * sizeof(double));
        }
    }
}

/* ---- one FMM march loop (ATR:1621-1674 and its five copies) ---------------- */
/* cx/cz + max_dist give the "front left the box" stop test (max_dist < 0: none);
 * nnz_bug: the level-1 loop of travel() passes nnx as nnz for x-direction close
 * updates (ATR:1645). */
static long g_update_calls = 0, g_fouds_calls = 0;

static double eval_node(Grid *g, const Tables *t, int iz, int ix, double dnx, int nnz, int nnx)
{
    double v = ali_update(g, t, iz, ix, dnx, nnz, nnx, NULL);
    g_update_calls++;
    if (v == -1.0) {
        v = ali_fouds18(g, t, iz, ix, dnx, dnx, g->nx, g->nz);
        g_fouds_calls++;
    }
    return v;
}

static void march(Grid *g, const Tables *t, double dnx, int cx, int cz, int max_dist, int nnz_bug)
{
    int finished = 0;
    int nnx = g->nx, nnz = g->nz, s;
    int switch_pending = 0, switched = 0;
    while (g->ntr > 0 && !finished) {
        int ix = g->btg[3], iz = g->btg[2];
        if (switch_pending) {   /* diagnostic 2: from here on a correct min-heap (heapified once) */
            int k;
            g_true_heap = 1;
            for (k = g->ntr / 2; k >= 1; k--) sift_down_from(g, k);
            switch_pending = 0; switched = 1;
            ix = g->btg[3]; iz = g->btg[2];
        }
        if (g_true_after >= 0 && max_dist < 0 && !switched && !g_true_heap) {
            int dz = abs(iz - g_src_z), dx = abs(ix - g_src_x);
            if ((dz > dx ? dz : dx) >= g_true_after) switch_pending = 1;   /* after this pop's neighbours */
        }
        g->nsts[AT(g, iz, ix)] = 0;
        downtree(g);
        for (s = 0; s < 2; s++) {
            int i = s == 0 ? ix - 1 : ix + 1;
            if (0 <= i && i <= nnx - 1) {
                int32_t st = g->nsts[AT(g, iz, i)];
                if (st == -1) {
                    g->ttn[AT(g, iz, i)] = eval_node(g, t, iz, i, dnx, nnz, nnx);
                    addtree(g, iz, i);
                } else if (st > 0) {
                    g->ttn[AT(g, iz, i)] = eval_node(g, t, iz, i, dnx, nnz_bug ? nnx : nnz, nnx);
                    updtree(g, iz, i);
                }
            } else if (max_dist >= 0 && abs(cx - i) == max_dist + 1) {
                finished = 1;
            }
        }
        for (s = 0; s < 2; s++) {
            int i = s == 0 ? iz - 1 : iz + 1;
            if (0 <= i && i <= nnz - 1) {
                int32_t st = g->nsts[AT(g, i, ix)];
                if (st == -1) {
                    g->ttn[AT(g, i, ix)] = eval_node(g, t, i, ix, dnx, nnz, nnx);
                    addtree(g, i, ix);
                } else if (st > 0) {
                    g->ttn[AT(g, i, ix)] = eval_node(g, t, i, ix, dnx, nnz, nnx);
                    updtree(g, i, ix);
                }
            } else if (max_dist >= 0 && abs(cz - i) == max_dist + 1) {
                finished = 1;
            }
        }
    }
    if (switched) g_true_heap = 0;
}

/* Analytic straight-ray seed of the source's own coarse cell + perimeter push
 * (ATR:1546-1612 with sign = -1; ATR:2223-2288 with sign = +1). */
static void seed_source(Grid *g1, const Tables *t, const Grid *src, int isz, int isx, int cz1, int cx1,
                        int side1, double dnx1, double sign)
{
    int i, j;
    size_t ps = AT(src, isz, isx);
    for (i = -side1; i <= side1; i++) {
        if (!(0 <= cz1 + i && cz1 + i <= g1->nz - 1)) continue;
        for (j = -side1; j <= side1; j++) {
            double angle, eff, vel, length;
            if (!(0 <= cx1 + j && cx1 + j <= g1->nx - 1)) continue;
            if (j == 0) angle = 90.0;
            else angle = RAD2DEG * atan((double)i / (double)j);
            eff = pymod(src->veln[ps] + sign * angle, 180.0);
            if (src->velpn[ps] != 0)
                vel = table_vel(t->group, t->ncol, eff, src->velpn[ps], src->vel_map[ps]);
            else
                vel = christoffel_group(eff, src->stif + 5 * ps, src->vel_map[ps]);
            length = dnx1 * sqrt((double)(i * i + j * j));
            g1->ttn[AT(g1, cz1 + i, cx1 + j)] = length / vel;
            g1->nsts[AT(g1, cz1 + i, cx1 + j)] = 0;
        }
    }
    {
        int xa = cx1 - side1 > 0 ? cx1 - side1 : 0;
        int xb = cx1 + side1 < g1->nx - 1 ? cx1 + side1 : g1->nx - 1;
        int za = cz1 - side1 > 0 ? cz1 - side1 : 0;
        int zb = cz1 + side1 < g1->nz - 1 ? cz1 + side1 : g1->nz - 1;
        if (cz1 - side1 >= 0) for (i = xa; i <= xb; i++) addtree(g1, cz1 - side1, i);
        if (cz1 + side1 <= g1->nz - 1) for (i = xa; i <= xb; i++) addtree(g1, cz1 + side1, i);
        if (cx1 - side1 >= 0) for (i = za; i <= zb; i++) addtree(g1, i, cx1 - side1);
        if (cx1 + side1 <= g1->nx - 1) for (i = za; i <= zb; i++) addtree(g1, i, cx1 + side1);
    }
}

/* Every-third-node injection into the next (3x coarser) grid (ATR:1719-1753). */
static void handoff(const Grid *a, int cza, int cxa, Grid *b, int czb, int cxb)
{
    int i, j;
    for (i = 0; i < a->nz + 1; i += 3) {
        for (j = 0; j < a->nx + 1; j += 3) {
            int pz, px;
            int32_t st;
            if (i > a->nz - 1 || j > a->nx - 1) continue; /* never taken: extents are 3k+1 */
            pz = czb + (i - cza) / 3;
            px = cxb + (j - cxa) / 3;
            b->ttn[AT(b, pz, px)] = a->ttn[AT(a, i, j)];
            st = a->nsts[AT(a, i, j)];
            if (st == 0) {
                int outer = 0;
                b->nsts[AT(b, pz, px)] = 0;
                if (i - 3 >= 0) { if (a->nsts[AT(a, i - 3, j)] == -1) outer = 1; } else outer = 1;
                if (i + 3 <= a->nz - 1) { if (a->nsts[AT(a, i + 3, j)] == -1) outer = 1; } else outer = 1;
                if (j - 3 >= 0) { if (a->nsts[AT(a, i, j - 3)] == -1) outer = 1; } else outer = 1;
                if (j + 3 <= a->nx - 1) { if (a->nsts[AT(a, i, j + 3)] == -1) outer = 1; } else outer = 1;
                if (outer) addtree(b, pz, px);
            }
            if (st > 0) addtree(b, pz, px);
        }
    }
}

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* ---- travel (ATR:1463-2117) ------------------------------------------------ */
static void travel_core(Grid *m, const Tables *t, double scx, double scz, double dnx)
{
    /* m: main grid with materials, ttn (caller-owned, NOT zeroed here), nsts/btg fresh. */
    int isx = (int)rint(scx / dnx), isz = (int)rint(scz / dnx);
    int nnx = m->nx, nnz = m->nz;
    static const int sizes[3] = {2, 6, 13};
    static const int scales[3] = {27, 9, 3};
    Grid lv[3];
    int cz[3], cx[3], l;
    for (l = 0; l < 3; l++) {
        int size = sizes[l], sc = scales[l];
        int left = imax(0, isx - size), right = imin(nnx - 1, isx + size);
        int bottom = imax(0, isz - size), top = imin(nnz - 1, isz + size);
        refine(m, bottom, top, left, right, sc, &lv[l]);
        grid_alloc_state(&lv[l], 1);
        cx[l] = sc * (isx - left);
        cz[l] = sc * (isz - bottom);
        if (l == 0)
            seed_source(&lv[0], t, m, isz, isx, cz[0], cx[0], (sc - 1) / 2, dnx / sc, -1.0);
        else
            handoff(&lv[l - 1], cz[l - 1], cx[l - 1], &lv[l], cz[l], cx[l]);
        march(&lv[l], t, dnx / sc, cx[l], cz[l], sc * size, l == 0);
    }
    handoff(&lv[2], cz[2], cx[2], m, isz, isx);
    g_src_z = isz; g_src_x = isx;
    march(m, t, dnx, 0, 0, -1, 0);
    for (l = 0; l < 3; l++) grid_free(&lv[l]);
}

/* ---- travel_finer_grid (ATR:2120-2832) ------------------------------------- */
static void travel_finer_core(const Grid *m0, const Tables *t, double scx, double scz, double dnx, int sg,
                              double *out)
{
    Grid f, l1, l2;
    int isx0 = (int)rint(scx / dnx), isz0 = (int)rint(scz / dnx);
    int isx, isz, nnx, nnz;
    int size1, side1, size2, left, right, bottom, top, cx1, cz1, cx2, cz2;
    size_t n, i;
    Grid m0s = *m0;
    double *zero_stif = NULL;
    if (!m0->stif) { /* ATR:2159-2160: zeros when None, and "not None" from here on */
        zero_stif = (double *)calloc((size_t)m0->nz * m0->nx * 5, sizeof(double));
        m0s.stif = zero_stif;
    }
    refine(&m0s, 0, m0->nz - 1, 0, m0->nx - 1, sg, &f);
    f.has_stif = 1;
    free(zero_stif);
    nnz = f.nz; nnx = f.nx;
    n = (size_t)nnz * nnx;
    f.ttn = out;
    for (i = 0; i < n; i++) out[i] = 0.0;
    grid_alloc_state(&f, 0);
    isx = sg * isx0; isz = sg * isz0;

    size1 = 2 * sg + (sg - 1) / 2;
    side1 = (9 - 1) / 2 + 9 * ((sg - 1) / 2);
    left = imax(0, isx - size1); right = imin(nnx - 1, isx + size1);
    bottom = imax(0, isz - size1); top = imin(nnz - 1, isz + size1);
    refine(&f, bottom, top, left, right, 9, &l1);
    grid_alloc_state(&l1, 1);
    cx1 = 9 * (isx - left); cz1 = 9 * (isz - bottom);
    seed_source(&l1, t, &f, isz, isx, cz1, cx1, side1, dnx / 9, +1.0);
    march(&l1, t, dnx / 9, cx1, cz1, 9 * size1, 0);

    size2 = size1 + 3 * sg;
    left = imax(0, isx - size2); right = imin(nnx - 1, isx + size2);
    bottom = imax(0, isz - size2); top = imin(nnz - 1, isz + size2);
    refine(&f, bottom, top, left, right, 3, &l2);
    grid_alloc_state(&l2, 1);
    cx2 = 3 * (isx - left); cz2 = 3 * (isz - bottom);
    handoff(&l1, cz1, cx1, &l2, cz2, cx2);
    march(&l2, t, dnx / 3, cx2, cz2, 3 * size2, 0);

    handoff(&l2, cz2, cx2, &f, isz, isx);
    g_src_z = isz; g_src_x = isx;
    march(&f, t, dnx, 0, 0, -1, 0);
    for (i = 0; i < n; i++) out[i] = out[i] / sg;
    f.ttn = NULL;
    grid_free(&l1); grid_free(&l2); grid_free(&f);
}

/* ---- time_between_points / ray_time / find_ray (ATR:2835-3465) ------------- */
static double time_between_points(double x1, double x2, double y1, double y2, double dnx, int sg,
                                  const Grid *m, const Tables *t)
{
    double section_time = 0.0, start_x, end_x, start_y, end_y, prev_x, prev_y, angle, mm = 0, cc = 0;
    double next_x, next_y, next_x_val, next_y_val;
    int finished_x = 0, finished_y = 0, dir_x, dir_y;
    x1 = x1 / sg; x2 = x2 / sg; y1 = y1 / sg; y2 = y2 / sg;
    start_x = x1; end_x = x2; start_y = y1; end_y = y2; prev_x = x1; prev_y = y1;
    if (x1 == x2) angle = 0;
    else angle = RAD2DEG * atan((y2 - y1) / (x2 - x1));
    if (end_x != start_x) {
        mm = (end_y - start_y) / (end_x - start_x);
        cc = start_y - mm * start_x;
    }
    dir_x = start_x < end_x ? 1 : -1;
    dir_y = start_y < end_y ? 1 : -1;
    next_x = rint(start_x) + dir_x * 0.5;
    next_y = rint(start_y) + dir_y * 0.5;
    while (!(finished_x && finished_y)) {
        int x_pos, y_pos;
        double eff, distance, vel;
        if (((next_x > end_x && dir_x == 1) || (next_x < end_x && dir_x == -1)) && !finished_x) {
            finished_x = 1; next_x = end_x;
        }
        if (((next_y > end_y && dir_y == 1) || (next_y < end_y && dir_y == -1)) && !finished_y) {
            finished_y = 1; next_y = end_y;
        }
        if (end_x == start_x) {
            next_x_val = start_x; next_y_val = next_y; next_y += dir_y;
        } else {
            double next_x_yval = mm * next_x + cc;
            if (mm != 0) {
                double next_y_xval = (next_y - cc) / mm;
                double dA = (start_x - next_x) * (start_x - next_x) + (start_y - next_x_yval) * (start_y - next_x_yval);
                double dB = (start_x - next_y_xval) * (start_x - next_y_xval) + (start_y - next_y) * (start_y - next_y);
                if (dA < dB) { next_x_val = next_x; next_y_val = next_x_yval; next_x += dir_x; }
                else { next_x_val = next_y_xval; next_y_val = next_y; next_y += dir_y; }
            } else {
                next_x_val = next_x; next_y_val = next_x_yval; next_x += dir_x;
            }
        }
        x_pos = (int)rint((prev_x + next_x_val) / 2);
        y_pos = (int)rint((prev_y + next_y_val) / 2);
        eff = pymod(m->veln[AT(m, y_pos, x_pos)] - angle, 180.0);
        distance = dnx * sqrt((prev_x - next_x_val) * (prev_x - next_x_val) + (prev_y - next_y_val) * (prev_y - next_y_val));
        vel = group_velocity_at(m, t, y_pos, x_pos, eff);
        section_time += distance * (1.0 / vel);
        prev_x = next_x_val; prev_y = next_y_val;
    }
    return section_time;
}

static double ray_time(const double *rx, const double *ry, int len, double dnx, int sg, const Grid *m,
                       const Tables *t)
{
    double tt = 0.0;
    int i;
    for (i = 0; i < len - 1; i++) tt += time_between_points(rx[i], rx[i + 1], ry[i], ry[i + 1], dnx, sg, m, t);
    return tt;
}

/* Quadratic local-minimum search shared by the four plane families (ATR:3191-3218). */
static double plane_min(const double *TTv, int len)
{
    double minimum, min_i;
    int j;
    if (TTv[0] < TTv[len - 1]) { minimum = TTv[0]; min_i = 0; }
    else { minimum = TTv[len - 1]; min_i = len - 1; }
    for (j = 1; j < len - 1; j++) {
        double t1 = TTv[j - 1], t2 = TTv[j], t3 = TTv[j + 1];
        if (t1 >= t2 && t2 <= t3) {
            double a = (t1 + t3 - 2 * t2) / 2, b = (t3 - t1) / 2, pos, val;
            if (a != 0) {
                pos = -b / (2 * a);
                val = a * (pos * pos) + b * pos + t2;
                pos += j;
            } else {
                pos = j; val = t2;
            }
            if (val < minimum) { min_i = pos; minimum = val; }
        }
    }
    return min_i;
}

#define FLAG_TTF_INCREASING 1
#define FLAG_LEFT_GRID 2
#define FLAG_EMPTY_PLANE 4

static int find_ray(double dnx, const Tables *t, double sx, double sy, double rx_, double ry_,
                    const double *rec, int fz, int fx, const Grid *m, int sg, double *ray_x, double *ray_y,
                    int cap, double *time_out, int *flag_out)
{
    const int search_dist = 3 * sg + 1, search_dist_2 = 2 * sg + 1;
    double last_x = sx, last_y = sy, lvx = rx_ - sx, lvy = ry_ - sy;
    int ray_len = 1, flag = 0;
    int nnx = fz, nnz = fx; /* reference's swapped names: nnx = rows, nnz = cols (ATR:3151-3152) */
    double *TTv = (double *)malloc(sizeof(double) * (size_t)(2 * search_dist + 8));
    const double r2 = sqrt(2.0);
#define REC(y, x) rec[(size_t)(y) * (size_t)fx + (size_t)(x)]
    ray_x[0] = sx; ray_y[0] = sy;
    while ((last_x - rx_) * (last_x - rx_) + (last_y - ry_) * (last_y - ry_) > (1.6 * sg) * (1.6 * sg)) {
        double v[4], best;
        int dir = 0, k, c_value, len, i;
        double nx_, ny_;
        if (ray_len >= cap - 1) break;
        if ((last_x - rx_) * (last_x - rx_) + (last_y - ry_) * (last_y - ry_) < (double)((4 * sg) * (4 * sg))) {
            lvx = rx_ - last_x; lvy = ry_ - last_y;
        }
        v[0] = fabs(lvx); v[1] = fabs(lvx + lvy) / r2; v[2] = fabs(lvy); v[3] = fabs(lvx - lvy) / r2;
        best = v[0];
        for (k = 1; k < 4; k++) if (v[k] > best) { best = v[k]; dir = k; }
        if (dir == 0) {
            int min_val, max_val;
            c_value = (int)rint(last_x);
            if (lvx > 0) c_value += sg; else c_value -= sg;
            if (c_value < 0 || c_value >= nnz) { flag |= FLAG_LEFT_GRID; break; }
            min_val = imax(0, (int)rint(last_y) - search_dist);
            max_val = imin(nnx - 1, (int)rint(last_y) + search_dist);
            len = max_val - min_val + 1;
            if (len < 1) { flag |= FLAG_EMPTY_PLANE; break; }
            for (i = 0; i < len; i++) {
                int xv = i + min_val;
                TTv[i] = REC(xv, c_value) + time_between_points(last_x, c_value, last_y, xv, dnx, sg, m, t);
            }
            nx_ = c_value; ny_ = plane_min(TTv, len) + min_val;
        } else if (dir == 1) {
            int min_x, max_x;
            c_value = (int)rint(last_x) + (int)rint(last_y);
            if (lvx > 0) {
                c_value += sg;
                min_x = imax(imax(0, c_value - (nnx - 1)), (int)rint(last_x) - search_dist_2);
                max_x = imin(imin(nnz - 1, c_value), c_value - (int)rint(last_y) + search_dist_2);
            } else {
                c_value -= sg;
                min_x = imax(imax(0, c_value - (nnx - 1)), c_value - (int)rint(last_y) - search_dist_2);
                max_x = imin(imin(nnz - 1, c_value), (int)rint(last_x) + search_dist_2);
            }
            len = max_x - min_x + 1;
            if (len < 1) { flag |= FLAG_EMPTY_PLANE; break; }
            for (i = 0; i < len; i++) {
                int xc = min_x + i, yc = -xc + c_value;
                TTv[i] = REC(yc, xc) + time_between_points(last_x, xc, last_y, yc, dnx, sg, m, t);
            }
            nx_ = min_x + plane_min(TTv, len); ny_ = c_value - nx_;
        } else if (dir == 2) {
            int min_val, max_val;
            c_value = (int)rint(last_y);
            if (lvy > 0) c_value += sg; else c_value -= sg;
            if (c_value < 0 || c_value >= nnx) { flag |= FLAG_LEFT_GRID; break; }
            min_val = imax(0, (int)rint(last_x) - search_dist);
            max_val = imin(nnz - 1, (int)rint(last_x) + search_dist);
            len = max_val - min_val + 1;
            if (len < 1) { flag |= FLAG_EMPTY_PLANE; break; }
            for (i = 0; i < len; i++) {
                int yv = i + min_val;
                TTv[i] = REC(c_value, yv) + time_between_points(last_x, yv, last_y, c_value, dnx, sg, m, t);
            }
            nx_ = plane_min(TTv, len) + min_val; ny_ = c_value;
        } else {
            int min_x, max_x;
            c_value = (int)rint(last_y) - (int)rint(last_x);
            if (lvx < 0) {
                c_value += sg;
                min_x = imax(imax(0, -c_value), (int)rint(last_y) - c_value - search_dist_2);
                max_x = imin(imin(nnz - 1, (nnx - 1) - c_value), (int)rint(last_x) + search_dist_2);
            } else {
                c_value -= sg;
                min_x = imax(imax(0, -c_value), (int)rint(last_x) - search_dist_2);
                max_x = imin(imin(nnz - 1, (nnx - 1) - c_value), (int)rint(last_y) - c_value + search_dist_2);
            }
            len = max_x - min_x + 1;
            if (len < 1) { flag |= FLAG_EMPTY_PLANE; break; }
            for (i = 0; i < len; i++) {
                int xc = min_x + i, yc = xc + c_value;
                TTv[i] = REC(yc, xc) + time_between_points(last_x, xc, last_y, yc, dnx, sg, m, t);
            }
            nx_ = min_x + plane_min(TTv, len); ny_ = nx_ + c_value;
        }
        ray_x[ray_len] = nx_; ray_y[ray_len] = ny_;
        if (REC((int)rint(last_y), (int)rint(last_x)) < REC((int)rint(ny_), (int)rint(nx_))) {
            flag |= FLAG_TTF_INCREASING; /* reference prints a warning here (ATR:3407) */
            break;
        }
        lvx = nx_ - last_x; last_x = nx_;
        lvy = ny_ - last_y; last_y = ny_;
        ray_len += 1;
    }
#undef REC
    ray_x[ray_len] = rx_; ray_y[ray_len] = ry_;
    ray_len += 1;
    *time_out = ray_time(ray_x, ray_y, ray_len, dnx, sg, m, t);
    if (flag_out) *flag_out = flag;
    free(TTv);
    return ray_len;
}

/* ------------------------------------------------------------------------- */
/* Exported C entry points (bound with ctypes from oracle/ali_oracle.py).      */
/* ------------------------------------------------------------------------- */
static void model_view(Grid *m, int nz, int nx, const double *veln, const int32_t *velpn,
                       const double *vel_map, const int64_t *stif, int has_stif)
{
    memset(m, 0, sizeof *m);
    m->nz = nz; m->nx = nx;
    m->veln = (double *)veln; m->velpn = (int32_t *)velpn; m->vel_map = (double *)vel_map;
    m->has_stif = has_stif;
    m->own_mat = 0;
    if (stif) {
        size_t n = (size_t)nz * nx * 5, i;
        m->stif = (double *)malloc(n * sizeof(double));
        for (i = 0; i < n; i++) m->stif[i] = (double)stif[i];
    }
}

int ali_oracle_travel(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                      const int64_t *stif, int has_stif, const double *group_tab, const double *phase_tab,
                      int ncol, double scx, double scz, double dnx, double *ttn)
{
    Grid m;
    Tables t = {group_tab, phase_tab, ncol};
    model_view(&m, nz, nx, veln, velpn, vel_map, stif, has_stif);
    m.ttn = ttn;
    grid_alloc_state(&m, 0);
    travel_core(&m, &t, scx, scz, dnx);
    free(m.nsts); free(m.btg); free(m.stif);
    return 0;
}

int ali_oracle_travel_finer(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                            const int64_t *stif, const double *group_tab, const double *phase_tab, int ncol,
                            double scx, double scz, double dnx, int sg, double *out)
{
    Grid m;
    Tables t = {group_tab, phase_tab, ncol};
    model_view(&m, nz, nx, veln, velpn, vel_map, stif, 1);
    travel_finer_core(&m, &t, scx, scz, dnx, sg, out);
    free(m.stif);
    return 0;
}

int ali_oracle_find_ray(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                        const int64_t *stif, int has_stif, const double *group_tab, int ncol, double dnx,
                        int sg, const double *rec_ttf, int fz, int fx, double sx, double sy, double rx,
                        double ry, double *ray_x, double *ray_y, int cap, double *time_out, int *flag_out)
{
    Grid m;
    Tables t = {group_tab, group_tab, ncol};
    int len;
    model_view(&m, nz, nx, veln, velpn, vel_map, stif, has_stif);
    len = find_ray(dnx, &t, sx, sy, rx, ry, rec_ttf, fz, fx, &m, sg, ray_x, ray_y, cap, time_out, flag_out);
    free(m.stif);
    return len;
}

double ali_oracle_time_between_points(int nz, int nx, const double *veln, const int32_t *velpn,
                                      const double *vel_map, const int64_t *stif, int has_stif,
                                      const double *group_tab, int ncol, double dnx, int sg, double x1,
                                      double x2, double y1, double y2)
{
    Grid m;
    Tables t = {group_tab, group_tab, ncol};
    double r;
    model_view(&m, nz, nx, veln, velpn, vel_map, stif, has_stif);
    r = time_between_points(x1, x2, y1, y2, dnx, sg, &m, &t);
    free(m.stif);
    return r;
}

/* Node-level operators on a caller-provided (ttn, nsts) state, for unit parity. */
double ali_oracle_update_node(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                              const int64_t *stif, int has_stif, const double *phase_tab, int ncol,
                              const double *ttn, const int32_t *nsts, int iz, int ix, double dnx,
                              int *stencil_out)
{
    Grid m;
    Tables t = {phase_tab, phase_tab, ncol};
    double r;
    model_view(&m, nz, nx, veln, velpn, vel_map, stif, has_stif);
    m.ttn = (double *)ttn; m.nsts = (int32_t *)nsts;
    r = ali_update(&m, &t, iz, ix, dnx, nz, nx, stencil_out);
    free(m.stif);
    return r;
}

double ali_oracle_fouds_node(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                             const int64_t *stif, int has_stif, const double *group_tab, int ncol,
                             const double *ttn, const int32_t *nsts, int iz, int ix, double dnx)
{
    Grid m;
    Tables t = {group_tab, group_tab, ncol};
    double r;
    model_view(&m, nz, nx, veln, velpn, vel_map, stif, has_stif);
    m.ttn = (double *)ttn; m.nsts = (int32_t *)nsts;
    r = ali_fouds18(&m, &t, iz, ix, dnx, dnx, nx, nz);
    free(m.stif);
    return r;
}

/* The same operators at ABSOLUTE grid coordinates of a large grid of which the caller holds only the rows
 * [z0, z0 + rows) (full width): update() interpolates in absolute coordinates (ATR:1444-1450) and tests the
 * grid edges, so a window cut out of the grid and renumbered can differ in the last ulp.  All arrays are
 * slabs [rows][nx] (stif [rows][nx][5]); nothing outside rows iz-2 .. iz+2 is read.  Test infrastructure
 * (tests/parity_tools.py). */
double ali_oracle_update_node_slab(int nz, int nx, int z0, int rows, const double *veln, const int32_t *velpn,
                                   const double *vel_map, const int64_t *stif, int has_stif, const double *phase_tab,
                                   const double *group_tab, int ncol, const double *ttn, const int32_t *nsts, int iz,
                                   int ix, double dnx, int *used_fouds)
{
    Grid m;
    Tables t = {group_tab, phase_tab, ncol};
    double r, *sd = NULL;
    const ptrdiff_t off = (ptrdiff_t)z0 * nx;
    size_t n = (size_t)rows * nx * 5, i;
    memset(&m, 0, sizeof m);
    m.nz = nz; m.nx = nx;
    m.veln = (double *)veln - off; m.velpn = (int32_t *)velpn - off; m.vel_map = (double *)vel_map - off;
    m.has_stif = has_stif;
    if (stif) {
        sd = (double *)malloc(n * sizeof(double));
        for (i = 0; i < n; i++) sd[i] = (double)stif[i];
        m.stif = sd - off * 5;
    }
    m.ttn = (double *)ttn - off; m.nsts = (int32_t *)nsts - off;
    r = ali_update(&m, &t, iz, ix, dnx, nz, nx, NULL);
    if (used_fouds) *used_fouds = 0;
    if (r == -1.0) {
        r = ali_fouds18(&m, &t, iz, ix, dnx, dnx, nx, nz);
        if (used_fouds) *used_fouds = 1;
    }
    free(sd);
    return r;
}

double ali_oracle_group_vel(double angle, double c22, double c23, double c33, double c44, double sigma,
                            double vel_scale)
{
    double s[5] = {c22, c23, c33, c44, sigma};
    return christoffel_group(angle, s, vel_scale);
}

double ali_oracle_phase_vel(double angle, double c22, double c23, double c33, double c44, double sigma,
                            double vel_scale)
{
    double s[5] = {c22, c23, c33, c44, sigma};
    return christoffel_phase(angle, s, vel_scale);
}

void ali_oracle_counters(long *update_calls, long *fouds_calls, int reset)
{
    if (update_calls) *update_calls = g_update_calls;
    if (fouds_calls) *fouds_calls = g_fouds_calls;
    if (reset) { g_update_calls = 0; g_fouds_calls = 0; }
}
