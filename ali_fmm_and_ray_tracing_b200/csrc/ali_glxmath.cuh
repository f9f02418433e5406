// ali_glxmath.cuh -- branch-light forms of glibc's sin / cos / tan / atan for SIMT hardware.
//
// csrc/ali_glibcmath.cuh restates glibc's routines instruction by instruction, including their control
// flow: one block of code per argument range.  On a GPU the lanes of a warp hold different angles, so
// a warp executes every range's block in turn (measured on B200, tools/bench/math_bench.cu: sin + cos
// of one argument 2300-2500 cycles with mixed ranges against 500-860 with a uniform one; atan 800-1000
// against 320-430).  The functions below compute THE SAME ARITHMETIC -- every multiply, add and fma of
// the range an argument falls in, in glibc's order, same constants, same tables -- but choose between
// the ranges with selects instead of branches, and share what sin and cos of the same argument have in
// common (the quadrant reduction).  They are not a new approximation: tests/test_kernel_replay.py
// (test_device_math_agrees_with_glibc) requires the bits of the running libm for every argument, exactly
// as for the literal restatement, and arguments outside the hot range fall back to it.
//
// Algorithm (glibc 2.39 sysdeps/ieee754/dbl-64/s_sin.c, s_atan.c, as compiled with FMA contraction;
// read off the machine code through ali_glibcmath.cuh):
//   sin / cos:  |x| < 0.855469: arguments (x, 0);  < 2.426265: pi/2 - |x| in two words;  else x - n pi/2
//               in two words (reduce_sincos).  Then do_sin / do_cos on the reduced pair: a table of
//               sin, cos at multiples of 1/128 (each in two words) plus short polynomials of the
//               remainder; TAYLOR_SIN below 0.126.
//   atan:       |x| < 1/16: odd polynomial;  < 1: table of degree-6 expansions around 241 points;
//               < 16: the same table on 1/|x| with an exact residual of the division;  larger: pi/2
//               minus an odd polynomial in 1/|x|.
#pragma once
#include "ali_glibcmath.cuh"

#if defined(__CUDACC__)
#define ALI_GX_DEV __device__ __forceinline__
#define ALI_GX_HI(d) ((uint32_t)__double2hiint(d))
#define ALI_GX_LO(d) ((uint32_t)__double2loint(d))
#define ALI_GX_COPYSIGN(m, s) copysign((m), (s))
#else
#define ALI_GX_DEV static inline
#define ALI_GX_HI(d) ((uint32_t)(ALI_GL_B(d) >> 32))
#define ALI_GX_LO(d) ((uint32_t)ALI_GL_B(d))
#define ALI_GX_COPYSIGN(m, s) copysign((m), (s))
#endif
#define ALI_GX_C(bits) ALI_GL_D(bits##ull)

// do_sin / do_cos / TAYLOR_SIN of glibc on a reduced pair (a, da), |a| < 0.86, selected without branches:
//   is_cos:  do_cos(a, da)                               (s_sin.c: do_cos)
//   else:    |a| < 0.126 ? TAYLOR_SIN(a*a, a, da) : do_sin(a, da)   (s_sin.c: do_sin)
ALI_GX_DEV double ali_gx_sincos_core(bool is_cos, double a, double da, const uint64_t *tab)
{
    const double big = ALI_GX_C(0x42c8000000000000);   // 1.5 * 2^45: rounds |a| to a multiple of 1/128
    const double sn3 = ALI_GX_C(0xbfc5555555555515), sn5 = ALI_GX_C(0x3f811110e829872f);
    const double cs2 = ALI_GX_C(0x3fe0000000000000), cs4 = ALI_GX_C(0xbfa5555555555535), cs6 = ALI_GX_C(0x3f56c16bedd9e239);
    const double aa = fabs(a);
    const double dxs = (a <= 0.0) ? -da : da;           // do_sin: if (x <= 0) dx = -dx
    const double dxc = (a < 0.0) ? -da : da;            // do_cos: if (x < 0) dx = -dx
    const double u = aa + big;
    const uint32_t k = ALI_GX_LO(u) << 2;
    const double rb = aa - (u - big);
    const double r = is_cos ? rb + dxc : rb;
    const double xx = r * r;
    const double p = ALI_GL_FMA(xx, sn5, sn3);
    const double t = r * xx;
    double q = ALI_GL_FMA(xx, cs6, cs4);
    q = ALI_GL_FMA(xx, q, cs2);
    const double cq = xx * q;
    const double sn = ALI_GL_D(tab[k]), ssn = ALI_GL_D(tab[k + 1]), cs = ALI_GL_D(tab[k + 2]), ccs = ALI_GL_D(tab[k + 3]);
    // do_sin
    const double s_s = r + ALI_GL_FMA(t, p, dxs);
    const double c_s = ALI_GL_FMA(dxs, r, cq);
    double res_s = sn + ALI_GL_FMA(s_s, cs, ALI_GL_FMA(-c_s, sn, ALI_GL_FMA(s_s, ccs, ssn)));
    res_s = ALI_GX_COPYSIGN(res_s, a);
    // TAYLOR_SIN
    const double xa = a * a;
    double tp = ALI_GL_FMA(xa, ALI_GX_C(0xbe5addffc2fcdf59), ALI_GX_C(0x3ec71de27b9a7ed9));
    tp = ALI_GL_FMA(xa, tp, ALI_GX_C(0xbf2a01a019db08b8));
    tp = ALI_GL_FMA(xa, tp, ALI_GX_C(0x3f81111111110ece));
    tp = ALI_GL_FMA(xa, tp, ALI_GX_C(0xbfc5555555555555));
    tp = ALI_GL_FMA(tp, a, -(da * 0.5));
    const double res_t = ALI_GL_FMA(xa, tp, da) + a;
    // do_cos
    const double s_c = ALI_GL_FMA(t, p, r);
    const double res_c = cs + ALI_GL_FMA(-s_c, sn, ALI_GL_FMA(-cq, cs, ALI_GL_FMA(-s_c, ssn, ccs)));
    return is_cos ? res_c : (aa < ALI_GX_C(0x3fc020c49ba5e354) ? res_t : res_s);
}

// sin(x) and cos(x) with glibc's bits.  `tab`: ali_gl_sincostab or a copy of it.
ALI_GX_DEV void ali_gx_sincos(double x, const uint64_t *tab, double &s_out, double &c_out)
{
    const uint32_t hx = ALI_GX_HI(x) & 0x7fffffffu;
    if (hx <= 0x3e3fffffu) { s_out = x; c_out = 1.0; return; }            // |x| < 2^-27 (both routines return early)
    if (hx <= 0x3e4fffffu || hx > 0x419921fau) {                          // 2^-27 .. 2^-26, huge, inf, nan: literal restatement
        s_out = ali_glibc_sin_t(x, tab); c_out = ali_glibc_cos_t(x, tab);
        return;
    }
    const double ax = fabs(x);
    const bool r1 = hx <= 0x3feb5fffu;                  // |x| < 0.855469
    const bool r2 = !r1 && hx <= 0x400368fcu;           // |x| < 2.426265
    // reduce_sincos: x - n * pi/2 in two words (used beyond 2.426265; computed for every lane)
    const double toint = ALI_GX_C(0x4338000000000000);
    const double tq = ALI_GL_FMA(x, ALI_GX_C(0x3fe45f306dc9c883), toint);
    const double xn = tq - toint;
    const uint32_t n = ALI_GX_LO(tq);
    double y = ALI_GL_FMA(-xn, ALI_GX_C(0x3ff921fb58000000), x);
    y = ALI_GL_FMA(-xn, ALI_GX_C(0xbe4dde973c000000), y);
    const double pp3 = ALI_GX_C(0xbc8cb3b398000000), pp4 = ALI_GX_C(0xbacd747f23e32ed7);
    const double t2 = ALI_GL_FMA(-xn, pp3, y);
    const double d0 = ALI_GL_FMA(-pp3, xn, y - t2);
    const double a3 = ALI_GL_FMA(-xn, pp4, t2);
    const double e3 = ALI_GL_FMA(-xn, pp4, t2 - a3);
    const double da3 = d0 + e3;
    // pi/2 - |x| (used below 2.426265)
    const double hp0 = ALI_GX_C(0x3ff921fb54442d18), hp1 = ALI_GX_C(0x3c91a62633145c07);
    const double yy = hp0 - ax;
    const double a2c = yy + hp1;
    const double da2c = (yy - a2c) + hp1;
    // sin:  r1: do_sin(x, 0)   r2: copysign(do_cos(yy, hp1), x)   r3: n odd ? do_cos(a3, da3) : do_sin(a3, da3), negated if n & 2
    const bool sin_is_cos = r1 ? false : (r2 ? true : (n & 1u) != 0u);
    const double sa = r1 ? x : (r2 ? yy : a3), sda = r1 ? 0.0 : (r2 ? hp1 : da3);
    double sv = ali_gx_sincos_core(sin_is_cos, sa, sda, tab);
    if (r2) sv = ALI_GX_COPYSIGN(sv, x);
    if (!r1 && !r2 && (n & 2u)) sv = -sv;
    // cos:  r1: do_cos(x, 0)   r2: do_sin(a2c, da2c)   r3: the same with n + 1
    const uint32_t n1 = n + 1u;
    const bool cos_is_cos = r1 ? true : (r2 ? false : (n1 & 1u) != 0u);
    const double ca = r1 ? x : (r2 ? a2c : a3), cda = r1 ? 0.0 : (r2 ? da2c : da3);
    double cv = ali_gx_sincos_core(cos_is_cos, ca, cda, tab);
    if (!r1 && !r2 && (n1 & 2u)) cv = -cv;
    s_out = sv;
    c_out = cv;
}

// One of the two (which != 0: cos), at the cost of one core evaluation instead of two.
ALI_GX_DEV double ali_gx_sin_or_cos(int which, double x, const uint64_t *tab)
{
    const uint32_t hx = ALI_GX_HI(x) & 0x7fffffffu;
    if (hx <= 0x3e3fffffu) return which ? 1.0 : x;
    if (hx <= 0x3e4fffffu || hx > 0x419921fau) return which ? ali_glibc_cos_t(x, tab) : ali_glibc_sin_t(x, tab);
    const double ax = fabs(x);
    const bool r1 = hx <= 0x3feb5fffu;
    const bool r2 = !r1 && hx <= 0x400368fcu;
    const double toint = ALI_GX_C(0x4338000000000000);
    const double tq = ALI_GL_FMA(x, ALI_GX_C(0x3fe45f306dc9c883), toint);
    const double xn = tq - toint;
    const uint32_t n = ALI_GX_LO(tq) + (which ? 1u : 0u);
    double y = ALI_GL_FMA(-xn, ALI_GX_C(0x3ff921fb58000000), x);
    y = ALI_GL_FMA(-xn, ALI_GX_C(0xbe4dde973c000000), y);
    const double pp3 = ALI_GX_C(0xbc8cb3b398000000), pp4 = ALI_GX_C(0xbacd747f23e32ed7);
    const double t2 = ALI_GL_FMA(-xn, pp3, y);
    const double d0 = ALI_GL_FMA(-pp3, xn, y - t2);
    const double a3 = ALI_GL_FMA(-xn, pp4, t2);
    const double e3 = ALI_GL_FMA(-xn, pp4, t2 - a3);
    const double da3 = d0 + e3;
    const double hp0 = ALI_GX_C(0x3ff921fb54442d18), hp1 = ALI_GX_C(0x3c91a62633145c07);
    const double yy = hp0 - ax;
    const double a2c = yy + hp1;
    const double da2c = (yy - a2c) + hp1;
    // sin: r1 do_sin(x, 0), r2 do_cos(yy, hp1) with the sign of x;  cos: r1 do_cos(x, 0), r2 do_sin(a2c, da2c);  r3: by quadrant
    const bool is_cos = r1 ? (which != 0) : (r2 ? (which == 0) : (n & 1u) != 0u);
    const double a = r1 ? x : (r2 ? (which ? a2c : yy) : a3);
    const double da = r1 ? 0.0 : (r2 ? (which ? da2c : hp1) : da3);
    double v = ali_gx_sincos_core(is_cos, a, da, tab);
    if (r2 && !which) v = ALI_GX_COPYSIGN(v, x);
    if (!r1 && !r2 && (n & 2u)) v = -v;
    return v;
}

ALI_GX_DEV double ali_gx_sin(double x, const uint64_t *tab) { return ali_gx_sin_or_cos(0, x, tab); }
ALI_GX_DEV double ali_gx_cos(double x, const uint64_t *tab) { return ali_gx_sin_or_cos(1, x, tab); }

// atan(x) with glibc's bits.  `tab`: ali_gl_atan_cij (241 rows of 7) or a copy of it.
ALI_GX_DEV double ali_gx_atan(double x, const uint64_t *tab)
{
    const double u = fabs(x);
    // below 2^-27-ish glibc returns x; beyond 16 the series in 1/x; nan: all through the literal restatement
    if (!(u >= ALI_GX_C(0x3e4bb67a00000000)) || !(u < 16.0)) return ali_glibc_atan_t(x, tab);
    const double hp0 = ALI_GX_C(0x3ff921fb54442d18), hp1 = ALI_GX_C(0x3c91a62633145c07);
    // A: |x| < 1/16
    const double v = x * x;
    double pa = ALI_GL_FMA(v, ALI_GX_C(0x3fb375f08b31cbce), ALI_GX_C(0xbfb7458022b13c25));
    pa = ALI_GL_FMA(v, pa, ALI_GX_C(0x3fbc71c6e5129a3b));
    pa = ALI_GL_FMA(v, pa, ALI_GX_C(0xbfc24924923f7603));
    pa = ALI_GL_FMA(v, pa, ALI_GX_C(0x3fc99999999997fd));
    pa = ALI_GL_FMA(v, pa, ALI_GX_C(0xbfd5555555555555));
    const double res_a = ALI_GL_FMA(x * v, pa, x);
    // B (|x| < 1) on u, C (|x| < 16) on w = 1 / u with the exact residual of the division
    const bool big = !(u < 1.0);
    const double w = 1.0 / u;
    const double pw = w * u;
    const double ew = ALI_GL_FMA(u, w, -pw);
    const double resid = (1.0 - pw) - ew;
    const double vv = big ? w : u;
    const double two52 = ALI_GX_C(0x4330000000000000);
    const int i = (int)(ALI_GL_FMA(vv, 256.0, two52) - two52) - 16;
    const uint64_t *row = tab + 7 * (i < 0 ? 0 : i);           // (i < 0 only on the |x| < 1/16 lanes, whose result is res_a)
    const double z0 = vv - ALI_GL_D(row[0]);
    const double z = big ? ALI_GL_FMA(resid, w, z0) : z0;
    double p = ALI_GL_FMA(z, ALI_GL_D(row[6]), ALI_GL_D(row[5]));
    p = ALI_GL_FMA(z, p, ALI_GL_D(row[4]));
    p = ALI_GL_FMA(z, p, ALI_GL_D(row[3]));
    p = ALI_GL_FMA(z, p, ALI_GL_D(row[2]));
    const double res_b = ALI_GL_FMA(p, z, ALI_GL_D(row[1]));
    const double res_c = (hp0 - ALI_GL_D(row[1])) + ALI_GL_FMA(-p, z, hp1);
    const double m = big ? res_c : res_b;
    return (u < 0.0625) ? res_a : ALI_GX_COPYSIGN(m, x);
}

// tan(x) with glibc's bits (s_tan.c).  `tab`: ali_gl_tan_xfg (186 rows of 4) or a copy of it.
//   |x| <= 0.0608: odd polynomial;  <= 0.787: table of (x_i, tan x_i, cot x_i) at steps of 1/256 and a short
//   series of the remainder, one division;  <= 25: x - n pi/2 in two words, then the same two cases on the
//   reduced pair -- tan for even n, -cot for odd n (the polynomial case of -cot divides in double-double).
// The table case -- 92 % of the angles the ray integrator produces -- is one straight line of code for every
// range and both parities; the polynomial case (within 3.5 degrees of a multiple of 90) is a branch.
ALI_GX_DEV double ali_gx_tan(double x, const uint64_t *tab)
{
    const double w = fabs(x);
    if (!(w > ALI_GX_C(0x3e4b096c00000000)) || !(w <= 25.0)) return ali_glibc_tan_t(x, tab);   // tiny, beyond 25, nan
    const bool small = w <= ALI_GX_C(0x3fe92f1a00000000);
    const double toint = ALI_GX_C(0x4338000000000000);
    const double tq = ALI_GL_FMA(x, ALI_GX_C(0x3fe45f306dc9c883), toint);
    const double xn = tq - toint;
    double y0 = ALI_GL_FMA(-xn, ALI_GX_C(0x3ff921fb58000000), x);
    y0 = ALI_GL_FMA(-xn, ALI_GX_C(0xbe4dde973c000000), y0);
    const double mp3 = ALI_GX_C(0xbc8cb3b399d747f2);
    const double ar = ALI_GL_FMA(-xn, mp3, y0);
    const double dar = ALI_GL_FMA(-xn, mp3, y0 - ar);
    const double a = small ? x : ar, da = small ? 0.0 : dar;
    const bool odd = !small && (ALI_GX_LO(tq) & 1u);
    const double ya = fabs(a);
    const double yya = (a < 0.0) ? -da : da, sy = (a < 0.0) ? -1.0 : 1.0;
    if (ya <= ALI_GX_C(0x3faf212d00000000)) {
        const double a2 = a * a;
        double p = ALI_GL_FMA(a2, ALI_GX_C(0x3f82385a3cf2e4ea), ALI_GX_C(0x3f9664ed49cfc666));
        p = ALI_GL_FMA(a2, p, ALI_GX_C(0x3faba1ba1cdb8745));
        p = ALI_GL_FMA(a2, p, ALI_GX_C(0x3fc11111111107c6));
        p = ALI_GL_FMA(a2, p, ALI_GX_C(0x3fd5555555555555));
        const double a3 = a * a2;
        if (small) return ALI_GL_FMA(a3, p, a);
        const double t2 = ALI_GL_FMA(a3, p, da);
        const double y = a + t2;
        if (!odd) return y;
        const double yy = (fabs(a) > fabs(t2)) ? (a - y) + t2 : (t2 - y) + a;
        const double r = 1.0 / y;
        const double pr = r * y;
        const double e = ALI_GL_FMA(r, y, -pr);
        const double s1 = ((1.0 - pr) - e) + 0.0;
        const double q = ALI_GL_FMA(-yy, r, s1) / y;
        const double h = r + q;
        return -(((r - h) + q) + h);
    }
    const int i = (int)ALI_GL_FMA(ya, 256.0, -15.5);
    const uint64_t *row = tab + 4 * i;
    const double z = (ya - ALI_GL_D(row[0])) + yya;
    const double z2 = z * z;
    const double t = ALI_GL_FMA(z * z2, ALI_GL_FMA(z2, ALI_GX_C(0x3fc11112e0a6b45f), ALI_GX_C(0x3fd5555555554dbd)), z);
    const double fi = ALI_GL_D(row[1]), gi = ALI_GL_D(row[2]);
    const double d = ((fi + gi) * t) / (odd ? t + fi : gi - t);
    return odd ? (gi - d) * -sy : (d + fi) * sy;
}
