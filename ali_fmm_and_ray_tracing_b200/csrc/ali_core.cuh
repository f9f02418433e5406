// ali_core.cuh -- device-side building blocks of the ALI-FMM hot path.
//
// Everything here is a pure device function: the ALI local update operator, the
// multi-stencil FD fallback, the Christoffel / table velocities, the fused
// nearest-neighbour material fetch and the straight-ray DDA integrator.
// Reference behaviour: /root/reference/Anis_TTF_rays.py ("ATR"), cited per function.
//
// Arithmetic is fp64 with the reference's operation order; the translation unit is
// compiled with -fmad=false so that a*b+c rounds twice exactly as numba/LLVM does.
//
// ALI_DEV is __device__ under nvcc.  tests/emu compiles the same header with a
// host compiler (ALI_DEV empty) to replay the kernels' logic on the CPU test box;
// the shipped library contains device code only.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define ALI_DEV __device__ __forceinline__
#define ALI_DEV_NOINLINE __device__ __noinline__
#define ALI_HD __host__ __device__ inline /* sizing helpers only (no arithmetic of the path) */
#else
#define ALI_DEV inline
#define ALI_DEV_NOINLINE inline
#define ALI_HD inline
#endif

// Transcendentals: on the device the accurate versions of ali_crmath.cuh (they agree with the
// glibc results the reference sees in 99.9 % of the calls; CUDA's libm does not).  The host replay
// (tests/emu) routes them through hooks so that it can use glibc (the reference's libm), the same
// accurate versions (bit-identical to the device), or inject last-ulp noise to measure how
// sensitive a solution is.
#include "ali_crmath.cuh"
#if defined(ALI_EMU_NOISE) && !defined(__CUDACC__)
double ali_emu_atan(double x);
double ali_emu_sin(double x);
double ali_emu_cos(double x);
double ali_emu_tan(double x);
#define ALI_ATAN(x) ali_emu_atan(x)
#define ALI_SIN(x) ali_emu_sin(x)
#define ALI_COS(x) ali_emu_cos(x)
#define ALI_TAN(x) ali_emu_tan(x)
#define ALI_SINCOS(tab, x, s, c) do { (void)(tab); (c) = ali_emu_cos(x); (s) = ali_emu_sin(x); } while (0)
#elif defined(__CUDACC__) && !defined(ALI_LIBM_MATH)
#define ALI_ATAN(x) ali_cr_atan(x)
#define ALI_SIN(x) ali_cr_sin(x)
#define ALI_COS(x) ali_cr_cos(x)
#define ALI_TAN(x) ali_cr_tan(x)
#define ALI_SINCOS(tab, x, s, c) ali_cr_sincos((tab), (x), (s), (c))
#else
#define ALI_ATAN(x) atan(x)
#define ALI_SIN(x) sin(x)
#define ALI_COS(x) cos(x)
#define ALI_TAN(x) tan(x)
#define ALI_SINCOS(tab, x, s, c) do { (void)(tab); (c) = cos(x); (s) = sin(x); } while (0)
#endif

#define ALI_PI 3.14159265358979323846
#define ALI_RAD2DEG (180.0 / ALI_PI)
#define ALI_DEG2RAD (ALI_PI / 180.0)

// ---------------------------------------------------------------------------
// Model resident in HBM.  One 64-byte record per COARSE node, built once per model from
// the caller's arrays (veln f64, velpn i32, vel_map f64, stif_den i64[5]): the int64
// stiffness entries are converted to fp64 once (exact below 2^53) instead of per evaluation.
// ---------------------------------------------------------------------------
struct AliMatRec {   // two 32-byte sectors: tabulated materials only touch the first
    double veln;     // orientation, degrees (as passed by the caller)
    double vel_map;  // velocity scale (as passed by the caller)
    int velpn;       // material id (0 = Christoffel from s[])
    int pad;
    double s[5];     // c22, c23, c33, c44 [MPa], rho; zeros when the model has no stif_den
};

struct AliModel {
    int nz, nx;
    const AliMatRec *rec;    // [nz*nx]
    int has_stif;            // reference's "stif_den is not None"
    const double *group_tab; // [361*ncol]
    const double *phase_tab; // [361*ncol]
    int ncol;
    double dnx;
};

// Maps a node of a (possibly twice-refined) grid to its coarse model node, fusing the
// reference's finer_grid_n / finer_grid_n_2 (ATR:26-91) into the fetch: the refined
// model is never materialised.  level index -> parent = o + (i + side1)/scale1;
// parent -> coarse = (p + side0)/scale0.  `cast` applies the int32 truncation of the
// orientation and the float32 rounding of vel_map that refined grids carry
// (ATR:1527-1529, 2156-2158).  mul1/mul0 are the 2^32/scale multipliers that replace the
// integer divisions (exact for indices below 2^32/scale).
struct AliMatView {
    int scale1, side1, z0, x0;
    int scale0, side0;
    int cast;
    unsigned mul1, mul0;
};

struct AliMat {
    double veln, vel_map;
    int velpn;
    const double *s;   // -> AliMatRec::s of the coarse node
};

ALI_HD unsigned ali_div_magic(int scale) { return scale <= 1 ? 0u : (unsigned)(4294967296ull / (unsigned)scale) + 1u; }

ALI_DEV int ali_div_by(int x, int scale, unsigned mul)
{
    if (scale <= 1) return x;
#if defined(__CUDA_ARCH__)
    return (int)__umulhi((unsigned)x, mul);
#else
    return (int)(((unsigned long long)(unsigned)x * mul) >> 32);
#endif
}

ALI_HD AliMatView ali_make_view(int scale1, int z0, int x0, int scale0, int cast)
{
    AliMatView v;
    v.scale1 = scale1; v.side1 = (scale1 - 1) / 2; v.z0 = z0; v.x0 = x0;
    v.scale0 = scale0; v.side0 = (scale0 - 1) / 2; v.cast = cast;
    v.mul1 = ali_div_magic(scale1); v.mul0 = ali_div_magic(scale0);
    return v;
}

ALI_DEV AliMatView ali_view_identity() { return ali_make_view(1, 0, 0, 1, 0); }

ALI_DEV void ali_fetch_mat(const AliModel &m, const AliMatView &v, int iz, int ix, AliMat &out)
{
    int pz = v.z0 + ali_div_by(iz + v.side1, v.scale1, v.mul1);
    int px = v.x0 + ali_div_by(ix + v.side1, v.scale1, v.mul1);
    int cz = ali_div_by(pz + v.side0, v.scale0, v.mul0);
    int cx = ali_div_by(px + v.side0, v.scale0, v.mul0);
    const AliMatRec *r = m.rec + ((size_t)cz * (size_t)m.nx + (size_t)cx);
#if defined(__CUDA_ARCH__)
    // one 16-byte load for (veln, vel_map) instead of two 8-byte ones
    const double2 vv = __ldg(reinterpret_cast<const double2 *>(r));
    double vn = vv.x, vm = vv.y;
#else
    double vn = r->veln, vm = r->vel_map;
#endif
    if (v.cast) {
        vn = (double)(int)vn;
        vm = (double)(float)vm;
    }
    out.veln = vn;
    out.vel_map = vm;
    out.velpn = r->velpn;
    out.s = r->s;
}

// The same fetch for a grid that is the coarse model refined by `scale0` only (the band march's view:
// no level, no origin), so that the hot loop keeps four values of the view instead of nine.
ALI_DEV void ali_fetch_mat_refined(const AliModel &m, int scale0, int side0, unsigned mul0, int cast, int iz, int ix,
                                   AliMat &out)
{
    const int cz = ali_div_by(iz + side0, scale0, mul0);
    const int cx = ali_div_by(ix + side0, scale0, mul0);
    const AliMatRec *r = m.rec + ((size_t)cz * (size_t)m.nx + (size_t)cx);
#if defined(__CUDA_ARCH__)
    const double2 vv = __ldg(reinterpret_cast<const double2 *>(r));
    double vn = vv.x, vm = vv.y;
#else
    double vn = r->veln, vm = r->vel_map;
#endif
    if (cast) {
        vn = (double)(int)vn;
        vm = (double)(float)vm;
    }
    out.veln = vn;
    out.vel_map = vm;
    out.velpn = r->velpn;
    out.s = r->s;
}

ALI_DEV int ali_imax2(int a, int b) { return a > b ? a : b; }
ALI_DEV int ali_imin2(int a, int b) { return a < b ? a : b; }

// Python float modulo for a positive divisor (numba real_divmod semantics).
ALI_DEV double ali_pymod(double a, double w)
{
    double m = fmod(a, w);
    if (m != 0.0) {
        if (m < 0.0) m += w;
    } else {
        m = 0.0;
    }
    return m;
}

// a mod 180 with the same result as ali_pymod(a, 180): the common ranges avoid fmod()
// (x - 180 is exact for 180 <= x < 360; for -180 <= x < 0 Python adds 180 with one rounding).
ALI_DEV double ali_pymod180(double a)
{
    if (a >= 0.0) {
        if (a < 180.0) return a;
        if (a < 360.0) return a - 180.0;
    } else if (a >= -180.0) {
        return a + 180.0;
    }
    return ali_pymod(a, 180.0);
}

// ---- velocities -------------------------------------------------------------
// 1-degree table interpolation (ATR:1371-1375).
ALI_DEV double ali_table_vel(const double *tab, int ncol, double eff, int col, double vm)
{
    int a1 = (int)floor(eff);
    int a2 = (a1 + 1) % 180;
    double rem = eff - a1;
    return vm * ((1 - rem) * tab[(size_t)a1 * ncol + col] + rem * tab[(size_t)a2 * ncol + col]);
}

// Christoffel phase velocity, stiffness in MPa (ATR:1400-1406).
ALI_DEV double ali_christoffel_phase(double eff, const double *s, double vm, const double *sincos_tab = nullptr)
{
    double c, sn;
    ALI_SINCOS(sincos_tab, ALI_DEG2RAD * eff, sn, c);
    double A = c * c * s[0] + sn * sn * s[3];
    double B = c * sn * (s[1] + s[3]);
    double C = c * c * s[3] + sn * sn * s[2];
    return 1000 * vm * sqrt((A + C + sqrt((A - C) * (A - C) + 4 * (B * B))) / (2 * s[4]));
}

// Christoffel group velocity (ATR:3542-3558 and inlined copies ATR:293-315, 1565-1587,
// 2241-2263, 2956-2978).
ALI_DEV double ali_christoffel_group(double eff, const double *s, double vm)
{
    double m90 = ali_pymod(eff, 90.0);
    if (m90 < 0.01 || m90 > 90 - 0.01) {
        double lam;
        if (fabs(ali_pymod(eff, 180.0) - 90) < 1) lam = s[2]; else lam = s[0];
        return 1000 * vm * sqrt(lam / s[4]);
    }
    double c22 = s[0], c23 = s[1], c33 = s[2], c44 = s[3];
    double t = ALI_TAN(ALI_DEG2RAD * eff);
    double A = c22 + c33 - 2 * c44;
    double B = (c23 + c44) * (t - 1 / t);
    double C = c22 - c33;
    double disc = sqrt(B * B + A * A - C * C);
    double ph;
    if (eff < 90)
        ph = ali_pymod(ALI_ATAN((-B - disc) / (C - A)), ALI_PI);
    else
        ph = ali_pymod(ALI_ATAN((-B + disc) / (C - A)), ALI_PI);
    double c2, s2;
    ALI_SINCOS(nullptr, 2 * ph, s2, c2);
    double lam = 0.5 * (c2 * (c22 - c44) + s2 * (c23 + c44) * t + c22 + c44);
    return 1000 * vm * sqrt(lam / s[4]) / ALI_COS(ALI_DEG2RAD * eff - ph);
}

ALI_DEV double ali_phase_velocity(const AliModel &m, const AliMat &mat, double eff, const double *sincos_tab = nullptr)
{
    if (mat.velpn != 0 || !m.has_stif) return ali_table_vel(m.phase_tab, m.ncol, eff, mat.velpn, mat.vel_map);
    return ali_christoffel_phase(eff, mat.s, mat.vel_map, sincos_tab);
}

ALI_DEV double ali_group_velocity(const AliModel &m, const AliMat &mat, double eff)
{
    if (mat.velpn != 0 || !m.has_stif) return ali_table_vel(m.group_tab, m.ncol, eff, mat.velpn, mat.vel_map);
    return ali_christoffel_group(eff, mat.s, mat.vel_map);
}

// ---- wavefront_angle_dist (ATR:1413-1460) -----------------------------------
// Coordinates are the grid's own absolute indices: the reference interpolates in
// absolute coordinates, and the rounding of (1-a)*x1 + a*x3 depends on them.
ALI_DEV void ali_wad(int ix, int iz, int x1, int x2, int x3, int z1, int z2, int z3, double y1, double y2,
                     double y3, double &angle, double &dist)
{
    if (y3 == y1) {
        angle = 0.0;
        dist = -1.0;
        return;
    }
    double a = (y2 - y1) / (y3 - y1);
    double xpos = (1 - a) * x1 + a * x3;
    double zpos = (1 - a) * z1 + a * z3;
    double dx = x2 - xpos;
    double dz = z2 - zpos;
    if (dx == 0)
        angle = 0.0;
    else
        angle = ali_pymod180(ALI_RAD2DEG * ALI_ATAN(dz / dx) + 90);
    dist = fabs(dz * (x2 - ix) - dx * (z2 - iz)) / sqrt(dx * dx + dz * dz);
}

// ---- the 12-neighbour window --------------------------------------------------
// Slot order: 0:(-2,0) 1:(-1,-1) 2:(-1,0) 3:(-1,+1) 4:(0,-2) 5:(0,-1) 6:(0,+1) 7:(0,+2)
//             8:(+1,-1) 9:(+1,0) 10:(+1,+1) 11:(+2,0)
#define ALI_W_DZ(s) ((s) == 0 ? -2 : (s) <= 3 ? -1 : (s) <= 7 ? 0 : (s) <= 10 ? 1 : 2)
#define ALI_W_DX(s) ((s) == 0 ? 0 : (s) == 1 ? -1 : (s) == 2 ? 0 : (s) == 3 ? 1 : (s) == 4 ? -2 : (s) == 5 ? -1 : \
                     (s) == 6 ? 1 : (s) == 7 ? 2 : (s) == 8 ? -1 : (s) == 9 ? 0 : (s) == 10 ? 1 : 0)
// (dz+2) | (dx+2) << 3 of a slot, and the three slots of a stencil packed as a | b<<6 | c<<12
#define ALI_W_CODE(s) ((unsigned)((ALI_W_DZ(s) + 2) | ((ALI_W_DX(s) + 2) << 3)))
#define ALI_STENCIL_CODE(a, b, c) (ALI_W_CODE(a) | (ALI_W_CODE(b) << 6) | (ALI_W_CODE(c) << 12))

struct AliWindow {
    double t[12];
    unsigned avail; // bit s: node in slot s is inside the grid and has an estimate (nsts >= 0)
};

// The ALI local update (ATR:904-1410) on a gathered window.
//   nnz/nnx : logical extents used for the edge tests (ATR:1146, 1265, 1316 ...).
//   dnx     : spacing of THIS grid (dnx/27, dnx/9, dnx/3 on the source levels).
//   returns -1.0 when no stencil gives a solution (caller falls back to fouds18).
// The window is indexed with compile-time slots only, so it lives in registers.
ALI_DEV double ali_update_window(const AliModel &m, const AliMat &mat, const AliWindow &w, int iz, int ix,
                                 int nnz, int nnx, double dnx, int *stencil_out, const double *sincos_tab = nullptr)
{
    const unsigned av = w.avail;
    int stencil_no = -1;
    double min_diff = 1000000.0;
    double angle = 0.0, dist = -1.0, wt = 0.0;
    double ta = 0.0, tb = 0.0, tc = 0.0; // apex, wing1/axial, wing2/diagonal of the chosen stencil
    unsigned code = 0;

    // phase 1: square/diamond stencils 0..7, smallest |T(wing1) - T(wing2)| wins, lowest index
    // on ties (ATR:989-1033)
#define ALI_P1(k, A, B, C)                                                                   \
    if ((av & ((1u << A) | (1u << B) | (1u << C))) == ((1u << A) | (1u << B) | (1u << C))) { \
        double diff = fabs(w.t[B] - w.t[C]);                                                 \
        if (diff < min_diff) {                                                               \
            stencil_no = k; min_diff = diff;                                                 \
            ta = w.t[A]; tb = w.t[B]; tc = w.t[C];                                           \
            code = ALI_STENCIL_CODE(A, B, C);                                                \
        }                                                                                    \
    }
    ALI_P1(0, 0, 1, 3)
    ALI_P1(1, 7, 3, 10)
    ALI_P1(2, 11, 8, 10)
    ALI_P1(3, 4, 1, 8)
    ALI_P1(4, 1, 5, 2)
    ALI_P1(5, 3, 2, 6)
    ALI_P1(6, 10, 9, 6)
    ALI_P1(7, 8, 5, 9)
#undef ALI_P1
    if (stencil_no != -1) {
        // B = the wing with the strictly smaller time, C the other (ATR:1040-1143)
        unsigned cb = (code >> 6) & 63u, cc = (code >> 12) & 63u, ca = code & 63u;
        if (!(tb < tc)) {
            double tmp = tb; tb = tc; tc = tmp;
            unsigned ct = cb; cb = cc; cc = ct;
        }
        ali_wad(ix, iz, ix + (int)(ca >> 3) - 2, ix + (int)(cb >> 3) - 2, ix + (int)(cc >> 3) - 2,
                iz + (int)(ca & 7u) - 2, iz + (int)(cb & 7u) - 2, iz + (int)(cc & 7u) - 2, ta, tb, tc, angle, dist);
        wt = tb;
    }

    if (stencil_no == -1 || ix == 0 || ix == nnx - 1 || iz == 0 || iz == nnz - 1) { // ATR:1146
        // phase 2: triangular stencils 8..15 (apex A two nodes away, axial n1, diagonal n2); A must
        // be the earliest; weighted difference criterion (ATR:1205-1260)
        const double r2 = sqrt(2.0);
        const double w1 = r2 - 1, w2 = 2 - r2;
        if (stencil_no == -1) min_diff = 1000000.0;
        stencil_no = -2;
#define ALI_P2(k, A, B, C)                                                                   \
    if ((av & ((1u << A) | (1u << B) | (1u << C))) == ((1u << A) | (1u << B) | (1u << C))) { \
        if (w.t[A] < fmin(w.t[B], w.t[C])) {                                                 \
            double diff = fabs(w1 * w.t[A] + w2 * w.t[B] - w.t[C]);                          \
            if (diff < min_diff) {                                                           \
                stencil_no = k; min_diff = diff;                                             \
                ta = w.t[A]; tb = w.t[B]; tc = w.t[C];                                       \
                code = ALI_STENCIL_CODE(A, B, C);                                            \
            }                                                                                \
        }                                                                                    \
    }
        ALI_P2(0, 11, 9, 10)
        ALI_P2(1, 0, 2, 3)
        ALI_P2(2, 0, 2, 1)
        ALI_P2(3, 11, 9, 8)
        ALI_P2(4, 4, 5, 8)
        ALI_P2(5, 7, 6, 10)
        ALI_P2(6, 7, 6, 3)
        ALI_P2(7, 4, 5, 1)
#undef ALI_P2
        if (stencil_no != -2) {
            const unsigned ca = code & 63u, cb = (code >> 6) & 63u, cc = (code >> 12) & 63u;
            const int xa = ix + (int)(ca >> 3) - 2, za = iz + (int)(ca & 7u) - 2;
            const int xb = ix + (int)(cb >> 3) - 2, zb = iz + (int)(cb & 7u) - 2;
            const int xc = ix + (int)(cc >> 3) - 2, zc = iz + (int)(cc & 7u) - 2;
            if (tb < tc) {
                // on the listed grid edge the wavefront is forced (ATR:1265-1267, 1316-1318)
                bool forced;
                double fangle;
                if (stencil_no <= 1) { forced = (ix == 0); fangle = 90.; }
                else if (stencil_no <= 3) { forced = (ix == nnx - 1); fangle = 90.; }
                else if (stencil_no <= 5) { forced = (iz == 0); fangle = 0.; }
                else { forced = (iz == nnz - 1); fangle = 0.; }
                if (forced) { angle = fangle; dist = 1.; }
                else ali_wad(ix, iz, xa, xb, xc, za, zb, zc, ta, tb, tc, angle, dist);
                wt = tb;
            } else {
                ali_wad(ix, iz, xa, xc, xb, za, zc, zb, ta, tc, tb, angle, dist);
                wt = tc;
            }
            if (stencil_no == 0) wt = tc; // stencil 8 always takes T(iz+1, ix+1) (ATR:1274)
            stencil_no += 8;
        }
    }
    if (stencil_out) *stencil_out = stencil_no;
    if (dist != -1.0) {
        double eff = ali_pymod180(mat.veln - angle);
        double vel = ali_phase_velocity(m, mat, eff, sincos_tab);   // sincos_tab: optional copy of the sin/cos table in faster memory
        return wt + (dist * dnx / vel);
    }
    return -1.0;
}

// Gathers the 12-neighbour window from a state accessor S providing
//   bool avail(int z, int x)  -- stored status is band/alive (node inside S's storage)
//   double tt(int z, int x)
// Nodes beyond the LOGICAL extents nnz x nnx are absent, as in ATR:940-987; S::avail must
// itself reject coordinates outside its storage.
template <class S>
ALI_DEV void ali_gather(const S &s, int iz, int ix, int nnz, int nnx, AliWindow &w)
{
    unsigned av = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const int dz = ALI_W_DZ(k), dx = ALI_W_DX(k);
        int z = iz + dz, x = ix + dx;
        double t = 0.0;
        // the reference tests only the side the offset points to (ATR:940-987), which
        // matters when it is handed a wrong nnz (ATR:1645)
        bool zin = dz < 0 ? (z >= 0) : (dz > 0 ? (z < nnz) : true);
        bool xin = dx < 0 ? (x >= 0) : (dx > 0 ? (x < nnx) : true);
        if (zin && xin && s.avail(z, x)) {
            av |= 1u << k;
            t = s.tt(z, x);
        }
        w.t[k] = t;
    }
    w.avail = av;
}

// ---- fouds18_A (ATR:240-901): multi-stencil FD fallback ------------------------
// Used only when the ALI update finds no stencil (ATR:2069-2070).  S additionally
// provides bool alive(int z, int x) (nsts == 0).  nnx/nnz are the grid's real extents.
template <class S>
ALI_DEV_NOINLINE double ali_fouds18(const AliModel &m, const AliMat &mat, const S &st, int iz, int ix, double dnx,
                                    double dnz, int nnx, int nnz)
{
    int tsw1 = 0, tsw2 = 0, tsw3 = 0, tsw4 = 0;
    double travm = 0, travmd = 0, travmt = 0, travms = 0;
    double wave_ang, eff, slown, mf2;
    double a = 0, b = 0, c = 0, tref = 0, tdiv = 1, u, em, rd1, tdsh, trav;
#define NSA(k, j) (st.alive((k), (j)))
#define TN(k, j) (st.tt((k), (j)))
    // 0-degree stencil (ATR:281-459)
    wave_ang = 0;
    eff = ali_pymod(wave_ang - mat.veln, 180.0);
    slown = 1.0 / ali_group_velocity(m, mat, eff);
    for (int jn = 0; jn < 2; jn++) {
        int j = jn == 0 ? ix - 1 : ix + 1;
        int j2 = 0, swj;
        if (!(0 <= j && j <= nnx - 1)) continue;
        swj = -1;
        if (j == ix - 1) { j2 = j - 1; if (j2 >= 0) { if (NSA(iz, j2)) swj = 0; } }
        else             { j2 = j + 1; if (j2 <= nnx - 1) { if (NSA(iz, j2)) swj = 0; } }
        if (NSA(iz, j) && swj == 0) { swj = -1; if (TN(iz, j) >= TN(iz, j2)) swj = 0; }
        else swj = -1;
        for (int kn = 0; kn < 2; kn++) {
            int k = kn == 0 ? iz - 1 : iz + 1;
            int k2 = 0, swk, swsol;
            if (!(0 <= k && k <= nnz - 1)) continue;
            swk = -1;
            if (k == iz - 1) { k2 = k - 1; if (k2 >= 0) { if (NSA(k2, ix)) swk = 0; } }
            else             { k2 = k + 1; if (k2 <= nnz - 1) { if (NSA(k2, ix)) swk = 0; } }
            if (NSA(k, ix) && swk == 0) { swk = -1; if (TN(k, ix) >= TN(k2, ix)) swk = 0; }
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
                } else if (NSA(k, ix)) {
                    double e1 = 3.0 * TN(k, ix), e2 = 4.0 * TN(iz, j) - TN(iz, j2);
                    u = 2.0 * dnx; a = 18;
                    b = -6.0 * (3.0 * TN(k, ix) + 4.0 * TN(iz, j) - TN(iz, j2));
                    c = e1 * e1 + e2 * e2 - 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else {
                    u = 2.0 * dnx; a = 1.0; b = 0.0;
                    c = -(u * u) * (slown * slown);
                    tref = 4.0 * TN(iz, j) - TN(iz, j2);
                    tdiv = 1.0; // ATR:395 overrides the 3.0 of ATR:389
                }
            } else if (NSA(iz, j)) {
                swsol = 1;
                if (swk == 0) {
                    double e1 = 3.0 * TN(iz, j), e2 = 4.0 * TN(k, ix) - TN(k2, ix);
                    u = dnx;
                    em = 3.0 * TN(iz, j) + 4.0 * TN(k, ix) - TN(k2, ix);
                    a = 18; b = -6.0 * em;
                    c = e1 * e1 + e2 * e2 - 3 * 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else if (NSA(k, ix)) {
                    double e3 = dnx * slown;
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
                } else if (NSA(k, ix)) {
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

    // 45-degree stencil (ATR:467-696)
    wave_ang = 45;
    eff = rint(ali_pymod(wave_ang - mat.veln, 180.0));
    slown = 1.0 / ali_group_velocity(m, mat, eff);
    mf2 = sqrt(2.0);
    for (int jn = 0; jn < 2; jn++) {
        int j = jn == 0 ? ix - 1 : ix + 1;
        int k = jn == 0 ? iz + 1 : iz - 1;
        int j2 = 0, k2 = 0, swdiag;
        if (!(0 <= j && j <= nnx - 1 && 0 <= k && k <= nnz - 1)) continue;
        swdiag = -1;
        if (j == ix - 1) { j2 = j - 1; k2 = k + 1; if (j2 >= 0 && k2 <= nnz - 1) { if (NSA(k2, j2)) swdiag = 0; } }
        else             { j2 = j + 1; k2 = k - 1; if (j2 <= nnx - 1 && k2 >= 0) { if (NSA(k2, j2)) swdiag = 0; } }
        if (NSA(k, j) && swdiag == 0) { swdiag = -1; if (TN(k, j) >= TN(k2, j2)) swdiag = 0; }
        else swdiag = -1;
        for (int jjn = 0; jjn < 2; jjn++) {
            int jj = jjn == 0 ? ix - 1 : ix + 1;
            int kk = jjn == 0 ? iz - 1 : iz + 1;
            int jj2 = 0, kk2 = 0, swskew, swsol;
            if (!(0 <= jj && jj <= nnx - 1 && 0 <= kk && kk <= nnz - 1)) continue;
            swskew = -1;
            if (jj == ix - 1) { jj2 = jj - 1; kk2 = kk - 1; if (jj2 >= 0 && kk2 >= 0) { if (NSA(kk2, jj2)) swskew = 0; } }
            else              { jj2 = jj + 1; kk2 = kk + 1; if (jj2 <= nnx - 1 && kk2 <= nnz - 1) { if (NSA(kk2, jj2)) swskew = 0; } }
            if (NSA(kk, jj) && swskew == 0) { swskew = -1; if (TN(kk, jj) >= TN(kk2, jj2)) swskew = 0; }
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
                } else if (NSA(kk, jj)) {
                    double e1 = 3.0 * TN(kk, jj), e2 = 4.0 * TN(k, j) - TN(k2, j2);
                    u = 2.0 * mf2 * dnx; a = 18;
                    b = -6.0 * (3.0 * TN(kk, jj) + 4.0 * TN(k, j) - TN(k2, j2));
                    c = e1 * e1 + e2 * e2 - 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else {
                    double e3;
                    u = mf2 * 2.0 * dnx; a = 1.0; b = 0.0;
                    e3 = u * slown;
                    c = -1.0 * (e3 * e3);
                    tref = (4.0 * TN(k, j) - TN(k2, j2));
                    tdiv = 3.0;
                }
            } else if (NSA(k, j)) {
                swsol = 1;
                if (swskew == 0) {
                    double e1 = 3.0 * TN(k, j), e2 = 4.0 * TN(kk, jj) - TN(kk2, jj2);
                    u = mf2 * dnx;
                    em = 3.0 * TN(k, j) + 4.0 * TN(kk, jj) - TN(kk2, jj2);
                    a = 18; b = -6.0 * em;
                    c = e1 * e1 + e2 * e2 - 3 * 4 * (u * u) * (slown * slown);
                    tref = 0.0; tdiv = 1.0;
                } else if (NSA(kk, jj)) {
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
                } else if (NSA(kk, jj)) {
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

    // 26.6 / 63.4-degree stencils (ATR:698-897)
    wave_ang = rint(ALI_RAD2DEG * ALI_ATAN(0.5));
    for (int lp = 0; lp < 2; lp++) {
        if (lp == 0) eff = ali_pymod(-wave_ang - mat.veln, 180.0);
        else eff = ali_pymod(wave_ang - mat.veln, 180.0);
        slown = 1.0 / ali_group_velocity(m, mat, eff);
        mf2 = sqrt(5.0);
        for (int q = 0; q < 4; q++) {
            // lp 0: j_vec = [ix-1, ix+2, ix+1, ix-2, ix-1], k_vec = [iz-2, iz-1, iz+2, iz+1, iz-2]
            // lp 1: j_vec = [ix+1, ix+2, ix-1, ix-2, ix+1], k_vec = [iz-2, iz+1, iz+2, iz-1, iz-2]
            int q1 = (q + 1) & 3;
            int dj0 = (q == 0) ? -1 : (q == 1) ? 2 : (q == 2) ? 1 : -2;
            int dk0 = (q == 0) ? -2 : (q == 1) ? -1 : (q == 2) ? 2 : 1;
            int dj1 = (q1 == 0) ? -1 : (q1 == 1) ? 2 : (q1 == 2) ? 1 : -2;
            int dk1 = (q1 == 0) ? -2 : (q1 == 1) ? -1 : (q1 == 2) ? 2 : 1;
            if (lp == 1) {
                // mirror: j offsets (+1, +2, -1, -2), k offsets (-2, +1, +2, -1)
                dj0 = (q == 0) ? 1 : (q == 1) ? 2 : (q == 2) ? -1 : -2;
                dk0 = (q == 0) ? -2 : (q == 1) ? 1 : (q == 2) ? 2 : -1;
                dj1 = (q1 == 0) ? 1 : (q1 == 1) ? 2 : (q1 == 2) ? -1 : -2;
                dk1 = (q1 == 0) ? -2 : (q1 == 1) ? 1 : (q1 == 2) ? 2 : -1;
            }
            int j = ix + dj0, k = iz + dk0, jj = ix + dj1, kk = iz + dk1;
            int swsol = 0;
            if (!(0 <= j && j <= nnx - 1)) continue;
            if (!(0 <= k && k <= nnz - 1)) continue;
            if (!(0 <= jj && jj <= nnx - 1)) continue;
            if (!(0 <= kk && kk <= nnz - 1)) continue;
            if (NSA(k, j)) {
                swsol = 1;
                if (NSA(kk, jj)) {
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
            } else if (NSA(kk, jj)) {
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
                if (lp == 0) {
                    if (tsw3 == 1) travmt = fmin(trav, travmt);
                    else { travmt = trav; tsw3 = 1; }
                } else {
                    if (tsw4 == 1) travms = fmin(trav, travms);
                    else { travms = trav; tsw4 = 1; }
                }
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
    {
        double self = TN(iz, ix);
        if (self != 0) travms = fmin(travms, self);
    }
#undef NSA
#undef TN
    return travms;
}

// One evaluation of a node: ALI update, FD fallback when it has no solution
// (ATR:2068-2071).  Returns the value the reference would store.
template <class S>
ALI_DEV double ali_eval_node(const AliModel &m, const AliMatView &mv, const S &st, int iz, int ix, int nnz_logic,
                             int nnx_logic, int nnz_real, int nnx_real, double dnx, int *used_fallback)
{
    AliMat mat;
    AliWindow w;
    ali_fetch_mat(m, mv, iz, ix, mat);
    ali_gather(st, iz, ix, nnz_logic, nnx_logic, w);
    double v = ali_update_window(m, mat, w, iz, ix, nnz_logic, nnx_logic, dnx, nullptr);
    if (v == -1.0) {
        v = ali_fouds18(m, mat, st, iz, ix, dnx, dnx, nnx_real, nnz_real);
        if (used_fallback) *used_fallback = 1;
    }
    return v;
}

// ---- time_between_points (ATR:2835-2989) ---------------------------------------
// Straight-segment travel time through the COARSE model; coordinates are fine-grid
// indices (divided by sg here, as the reference does).  `max_pieces` bounds the DDA
// (the reference has no bound; a well-posed segment needs < 2*(|dx|+|dy|)+4 pieces).
ALI_DEV double ali_time_between_points(const AliModel &m, double x1, double x2, double y1, double y2, int sg,
                                       int max_pieces)
{
    double section_time = 0.0, angle, mm = 0, cc = 0;
    double next_x, next_y, next_x_val, next_y_val;
    bool finished_x = false, finished_y = false;
    x1 = x1 / sg; x2 = x2 / sg; y1 = y1 / sg; y2 = y2 / sg;
    const double start_x = x1, end_x = x2, start_y = y1, end_y = y2;
    double prev_x = x1, prev_y = y1;
    if (x1 == x2) angle = 0;
    else angle = ALI_RAD2DEG * ALI_ATAN((y2 - y1) / (x2 - x1));
    if (end_x != start_x) {
        mm = (end_y - start_y) / (end_x - start_x);
        cc = start_y - mm * start_x;
    }
    const int dir_x = start_x < end_x ? 1 : -1;
    const int dir_y = start_y < end_y ? 1 : -1;
    next_x = rint(start_x) + dir_x * 0.5;
    next_y = rint(start_y) + dir_y * 0.5;
    const AliMatView idv = ali_view_identity();
    int pieces = 0;
    while (!(finished_x && finished_y) && pieces < max_pieces) {
        pieces++;
        if (((next_x > end_x && dir_x == 1) || (next_x < end_x && dir_x == -1)) && !finished_x) {
            finished_x = true; next_x = end_x;
        }
        if (((next_y > end_y && dir_y == 1) || (next_y < end_y && dir_y == -1)) && !finished_y) {
            finished_y = true; next_y = end_y;
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
        int x_pos = (int)rint((prev_x + next_x_val) / 2);
        int y_pos = (int)rint((prev_y + next_y_val) / 2);
        // the reference has no bounds check here; clamp so a stray index cannot fault
        x_pos = x_pos < 0 ? 0 : (x_pos > m.nx - 1 ? m.nx - 1 : x_pos);
        y_pos = y_pos < 0 ? 0 : (y_pos > m.nz - 1 ? m.nz - 1 : y_pos);
        AliMat mat;
        ali_fetch_mat(m, idv, y_pos, x_pos, mat);
        double eff = ali_pymod180(mat.veln - angle);
        double distance = m.dnx * sqrt((prev_x - next_x_val) * (prev_x - next_x_val) +
                                       (prev_y - next_y_val) * (prev_y - next_y_val));
        double vel = ali_group_velocity(m, mat, eff);
        section_time += distance * (1.0 / vel);
        prev_x = next_x_val; prev_y = next_y_val;
    }
    return section_time;
}
