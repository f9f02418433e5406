// ali_emu.cpp -- host replay of the CUDA kernels' logic.  TEST TOOL ONLY.
//
// Compiles the product's device headers (csrc/ali_core.cuh, ali_seq.cuh, ali_band.cuh)
// with a host compiler and runs the phases of the kernels sequentially, so that the
// algorithm (sequential near-source replica + band-synchronous march + ray tracer)
// can be checked against the oracle on the GPU-less test box.  Never loaded by the
// product package.
#include <vector>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#define ALI_EMU_NOISE 1
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_core.cuh"
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_seq.cuh"
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_band.cuh"
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_ray.cuh"

// Last-ulp noise on transcendental results (probability g_noise_p, +-1 ulp), to measure how
// sensitive a solution is to the libm in use (GPU libm and glibc differ in the last ulps).
static double g_noise_p = 0.0;
static unsigned long long g_noise_state = 88172645463325252ull;
double ali_emu_noise(double v)
{
    if (g_noise_p <= 0.0) return v;
    g_noise_state ^= g_noise_state << 13; g_noise_state ^= g_noise_state >> 7; g_noise_state ^= g_noise_state << 17;
    double u = (double)(g_noise_state >> 11) / 9007199254740992.0;
    if (u >= g_noise_p) return v;
    return nextafter(v, (g_noise_state & 1) ? 1e300 : -1e300);
}
extern "C" void emu_set_noise(double p, unsigned long long seed) { g_noise_p = p; if (seed) g_noise_state = seed; }

// Which sin / cos / tan / atan the replay uses: 0 = the running libm (what the reference runs on), 1 = the
// device's functions (csrc/ali_glibcmath.cuh, the restatement of glibc's routines the kernels run).  The two
// return the same bits; the switch exists to prove it (tests/test_kernel_replay.py).
static int g_crmath = 0;
extern "C" void emu_set_crmath(int on) { g_crmath = on; }
// fn: 0 atan, 1 sin, 2 cos, 3 tan: the literal restatements (ali_glibcmath.cuh);
//     4 atan, 5 sin, 6 cos, 7 tan: the branch-light forms the kernels call (ali_glxmath.cuh)
static double device_math(int fn, double x)
{
    switch (fn) {
    case 0: return ali_glibc_atan(x);
    case 1: return ali_glibc_sin(x);
    case 2: return ali_glibc_cos(x);
    case 3: return ali_glibc_tan(x);
    case 4: return ali_gx_atan(x, ali_gl_atan_cij);
    case 5: return ali_gx_sin(x, ali_gl_sincostab);
    case 6: return ali_gx_cos(x, ali_gl_sincostab);
    default: return ali_gx_tan(x, ali_gl_tan_xfg);
    }
}
extern "C" double emu_crmath_eval(int fn, double x) { return device_math(fn, x); }
// Bulk comparison of the device's functions with the running libm: n arguments, returns the number of results
// whose bits differ (NaN == NaN), and the first offending argument.
extern "C" long long emu_math_mismatches(int fn, const double *x, long long n, double *first_bad)
{
    long long bad = 0;
    for (long long i = 0; i < n; i++) {
        const double a = device_math(fn, x[i]);
        const int lf = fn >= 4 ? fn - 4 : fn;   // 4..7 -> atan, sin, cos, tan
        const double b = lf == 0 ? atan(x[i]) : lf == 1 ? sin(x[i]) : lf == 2 ? cos(x[i]) : tan(x[i]);
        if (std::memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) {
            if (bad == 0 && first_bad) *first_bad = x[i];
            bad++;
        }
    }
    return bad;
}
double ali_emu_atan(double x) { return ali_emu_noise(g_crmath ? ali_gx_atan(x, ali_gl_atan_cij) : atan(x)); }
double ali_emu_sin(double x) { return ali_emu_noise(g_crmath ? ali_gx_sin(x, ali_gl_sincostab) : sin(x)); }
double ali_emu_cos(double x) { return ali_emu_noise(g_crmath ? ali_gx_cos(x, ali_gl_sincostab) : cos(x)); }
double ali_emu_tan(double x) { return ali_emu_noise(g_crmath ? ali_gx_tan(x, ali_gl_tan_xfg) : tan(x)); }

struct HostModel {
    std::vector<AliMatRec> rec;
    AliModel m;
};

// Host replay of ali_records_kernel.
static void make_model(HostModel &h, int nz, int nx, const double *veln, const int32_t *velpn,
                       const double *vel_map, const long long *stif, int has_stif, const double *group_tab,
                       const double *phase_tab, int ncol, double dnx)
{
    size_t n = (size_t)nz * nx;
    h.rec.resize(n);
    for (size_t i = 0; i < n; i++) {
        AliMatRec &r = h.rec[i];
        r.veln = veln[i]; r.vel_map = vel_map[i]; r.velpn = velpn[i]; r.pad = 0;
        for (int k = 0; k < 5; k++) r.s[k] = stif ? (double)stif[5 * i + k] : 0.0;
    }
    AliModel &m = h.m;
    m.nz = nz; m.nx = nx; m.rec = h.rec.data();
    m.has_stif = has_stif; m.group_tab = group_tab; m.phase_tab = phase_tab; m.ncol = ncol; m.dnx = dnx;
}


// Sequential replay of ali_rays_kernel: `nlanes` lanes evaluate candidates, lane 0 selects.
static int ali_emu_trace_ray(const AliModel &m, int sg, const double *rec, int fz, int fx, double sx, double sy,
                             double rx, double ry, double *ray_x, double *ray_y, int cap, double *time_out,
                             int *flag_out, int nlanes)
{
    AliRayState s;
    s.last_x = sx; s.last_y = sy; s.lvx = rx - sx; s.lvy = ry - sy; s.rx = rx; s.ry = ry;
    s.len = 1; s.flag = 0; s.done = 0;
    ray_x[0] = sx; ray_y[0] = sy;
    const int maxc = ali_ray_max_candidates(sg);
    std::vector<double> TT(maxc), vals(maxc), poss(maxc);
    while (ali_ray_continue(s, sg)) {
        AliRayPlane pl;
        if (s.len >= cap - 1) { s.flag |= ALI_RAY_CAPACITY; break; }
        if (!ali_ray_choose_plane(s, sg, fz, fx, pl)) break;
        for (int lane = 0; lane < nlanes; lane++)
            for (int i = lane; i < pl.len; i += nlanes)
                TT[i] = ali_ray_candidate_time(m, rec, fx, pl, i, s.last_x, s.last_y, sg);
        for (int lane = 0; lane < nlanes; lane++)
            for (int j = 1 + lane; j < pl.len - 1; j += nlanes) vals[j] = ali_ray_local_min(TT.data(), j, poss[j]);
        double min_i = ali_ray_select(TT.data(), vals.data(), poss.data(), pl.len);
        if (!ali_ray_advance(s, pl, min_i, rec, fx, ray_x, ray_y)) break;
    }
    ray_x[s.len] = rx; ray_y[s.len] = ry;
    s.len += 1;
    double tt = 0.0; // ray_time (ATR:2992-3022): in-order sum of the segment times
    for (int i = 0; i < s.len - 1; i++)
        tt += ali_time_between_points(m, ray_x[i], ray_x[i + 1], ray_y[i], ray_y[i + 1], sg, 1 << 20);
    *time_out = tt;
    if (flag_out) *flag_out = s.flag;
    return s.len;
}

extern "C" double emu_model_vmax(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                                 const long long *stif, int has_stif, const double *group_tab,
                                 const double *phase_tab, int ncol, double dnx)
{
    HostModel hm;
    make_model(hm, nz, nx, veln, velpn, vel_map, stif, has_stif, group_tab, phase_tab, ncol, dnx);
    const AliModel &m = hm.m;
    double best = 0.0;
    for (int iz = 0; iz < nz; iz++)
        for (int ix = 0; ix < nx; ix++) {
            double v = ali_node_vmax(m, iz, ix);
            if (v > best) best = v;
        }
    return best;
}

// Band rounds on one grid (host replay of the kernel's round loop).
//   stop_mask != 0: a refined level -- end when the first node on an unclipped box edge is
//   accepted (ATR:1651-1652), accepting only nodes up to that node's time in the last round and
//   evaluating the nodes that round enlisted.
struct BandStats { long long rounds = 0, evals = 0, fbs = 0, maxlist = 0; };

static void band_run(const AliModel &m, AliBandGrid &bg, std::vector<unsigned> &list, double delta, int stop_mask,
                     bool eager, BandStats &st)
{
    std::vector<unsigned> next;
    std::vector<double> tnew;
    bool last_round = false;
    while (!list.empty()) {
        st.rounds++;
        if ((long long)list.size() > st.maxlist) st.maxlist = (long long)list.size();
        tnew.resize(list.size());
        for (size_t i = 0; i < list.size(); i++) {
            const int iz = ALI_PACK_Z(list[i]), ix = ALI_PACK_X(list[i]);
            const size_t node = bg.ti(iz, ix);
            const size_t di = ali_dirty_index(bg, iz, ix);
            if (bg.dirty[di] || eager) {
                int fb = 0;
                bg.dirty[di] = 0;
                tnew[i] = ali_band_eval(m, &m, bg, &bg, iz, ix, &fb, nullptr, stop_mask != 0);
                if (fb) bg.dirty[di] = 1;
                st.evals++; st.fbs += fb;
            } else {
                tnew[i] = bg.T[node];
            }
        }
        double tmin = 1e300;
        for (size_t i = 0; i < list.size(); i++) {
            ali_band_publish(bg, ALI_PACK_Z(list[i]), ALI_PACK_X(list[i]), tnew[i]);
            if (tnew[i] < tmin) tmin = tnew[i];
        }
        if (last_round) break;   // the extra evaluation round after the stop
        double thr = tmin + delta;
        if (stop_mask) {
            double tstar = 1e300;
            for (size_t i = 0; i < list.size(); i++)
                if (tnew[i] <= thr && ali_level_on_stop_edge(stop_mask, bg.nz, bg.nx, ALI_PACK_Z(list[i]), ALI_PACK_X(list[i])) &&
                    tnew[i] < tstar)
                    tstar = tnew[i];
            if (tstar < 1e300) { thr = tstar; last_round = true; }
        }
        next.clear();
        for (size_t i = 0; i < list.size(); i++) {
            if (tnew[i] <= thr) {
                unsigned nb[4];
                int cnt = ali_band_accept(bg, ALI_PACK_Z(list[i]), ALI_PACK_X(list[i]), nb);
                for (int k = 0; k < cnt; k++) {
                    next.push_back(nb[k]);   // new nodes are always evaluated next round (kernel: work list)
                    bg.dirty[ali_dirty_index(bg, ALI_PACK_Z(nb[k]), ALI_PACK_X(nb[k]))] = 1;
                }
            } else {
                next.push_back(list[i]);
            }
        }
        list.swap(next);
    }
}

// Sequential state of a grid window -> band representation (far = NaN bits, alive byte, list).
static void seq_to_band(const AliSeqGrid &g, AliBandGrid &bg, std::vector<unsigned> &list)
{
    for (int z = 0; z < g.wnz; z++)
        for (int x = 0; x < g.wnx; x++) {
            int32_t s = g.st[(size_t)z * g.wnx + x];
            const int az = g.wz0 + z, ax = g.wx0 + x;
            size_t node = bg.ti(az, ax);
            if (s >= 0) bg.T[node] = g.tt(az, ax);   // the sequential phase keeps its window in its own buffer
            if (s == 0) bg.st[node] = ALI_ST_ALIVE;
            else if (s > 0) { list.push_back(ALI_PACK(az, ax)); bg.dirty[ali_dirty_index(bg, az, ax)] = 1; }
            else { unsigned long long bits = ALI_T_FAR_BITS; std::memcpy(&bg.T[node], &bits, 8); }
        }
}

// Band representation of a level -> sequential statuses for the hand-off (-1 far, 0 alive, 1 band).
static void band_to_seq(const AliBandGrid &bg, AliSeqGrid &g)
{
    for (int z = 0; z < g.nz; z++)
        for (int x = 0; x < g.nx; x++) {
            size_t node = (size_t)z * g.nx + x;
            if (bg.st[node] == ALI_ST_ALIVE) g.st[node] = 0;
            else if (bg.T[node] >= 0.0) g.st[node] = 1;
            else g.st[node] = -1;
        }
}

// counters: [0] seq pops, [1] seq evals, [2] seq fallbacks, [3] band rounds, [4] band evals,
//           [5] band fallbacks, [6] max list length, [7] overflow flag, [8] level band rounds,
//           [9] level band evals, [10] cooperative steps, [11] evaluations executed in them
// Lanes of the cooperative sequential march (ali_seq.cuh); 0 = lane 0 alone walks the reference's loop.
static int g_coop_lanes = 32;
static int g_tiled = 0;   // replay the band march on the kernel's 4 x 4-tiled field layout
extern "C" void emu_set_tiled(int on) { g_tiled = on; }
extern "C" void emu_set_coop(int nlanes) { g_coop_lanes = nlanes; }

extern "C" int emu_ttf(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                       const long long *stif, int has_stif, const double *group_tab, const double *phase_tab,
                       int ncol, double dnx, int src_iz, int src_ix, int sg, int margin, double delta_frac,
                       double vmax, int eager, int level_margin, double *T, long long *counters)
{
    HostModel hm;
    make_model(hm, nz, nx, veln, velpn, vel_map, stif, has_stif, group_tab, phase_tab, ncol, dnx);
    const AliModel &m = hm.m;
    AliSourcePlan p;
    ali_make_plan(p, m, src_iz, src_ix, sg, margin);
    const size_t n = (size_t)p.nz * p.nx;
    size_t lvl = ali_plan_max_level_nodes(p);
    size_t win = (size_t)(2 * (p.stop_r + 4) + 1) * (2 * (p.stop_r + 4) + 1);
    size_t cap = lvl > win ? lvl : win;
    std::vector<double> tA(cap), tB(cap);
    std::vector<int32_t> sA(cap), sB(cap), heap(2 * (cap / 2 + 64));
    AliSeqScratch sc;
    sc.tA = tA.data(); sc.tB = tB.data(); sc.sA = sA.data(); sc.sB = sB.data();
    sc.heap = reinterpret_cast<AliHeapEnt *>(heap.data()); sc.heap_cap = (int)(cap / 2 + 64); sc.status_cap = cap;
    std::vector<double> hkey(ALI_HKEY_SLOTS(cap / 2 + 64));
    sc.hkey = hkey.data();
    std::vector<double> cval(g_coop_lanes > 0 ? cap : 0);
    std::vector<uint8_t> cflag(g_coop_lanes > 0 ? cap : 0);
    sc.cval = g_coop_lanes > 0 ? cval.data() : nullptr;
    sc.cflag = g_coop_lanes > 0 ? cflag.data() : nullptr;
    const bool coop = g_coop_lanes > 0 && level_margin < 0;
    std::memset(T, ALI_T_UNSET_BYTE, n * sizeof(double)); // NaN = no estimate
    const double delta = delta_frac * dnx / vmax;   // the levels scale it with their spacing

    AliSrcState s;
    ali_src_begin(s, p);
    BandStats lst, mst;
    std::vector<uint8_t> lalive(lvl), ldirty(ali_dirty_bytes(2 * 4096, 2 * 4096) > 0 ? 0 : 0);
    for (int l = 0; l < p.nlev; l++) {
        ali_src_level_geometry(s, m, p, sc, l);
        ali_src_level_fill(s, m, p, l, 0, 1, false);
        ali_src_level_start(s, p, l);
        AliSeqGrid &g = s.lv[l & 1];
        const int ring = ali_src_level_ring(p, l);
        const int stop_r = level_margin >= 0 ? ring + level_margin : -1;
        int why = coop ? ali_src_level_seq_coop(s, m, p, l, 0, g_coop_lanes) : ali_src_level_seq(s, m, p, l, stop_r);
        if (why == ALI_SEQ_HANDOVER) {
            // band rounds on the level grid until the front leaves the refined box
            AliBandGrid bg;
            bg.nz = g.nz; bg.nx = g.nx; bg.T = g.t; bg.dnx = g.dnx; bg.mv = g.mv;
            std::fill(lalive.begin(), lalive.begin() + (size_t)g.nz * g.nx, (uint8_t)ALI_ST_FAR);
            ldirty.assign(ali_dirty_bytes(g.nz, g.nx), 0);
            bg.st = lalive.data(); bg.dirty = ldirty.data(); bg.tiles_x = ali_dirty_tiles_x(g.nx); bg.t4x = 0;
            std::vector<unsigned> list;
            seq_to_band(g, bg, list);
            int mask = ali_level_stop_mask(g.nz, g.nx, s.cz[l & 1], s.cx[l & 1], p.scale[l] * p.size[l]);
            band_run(m, bg, list, delta / p.scale[l], mask, eager != 0, lst);
            band_to_seq(bg, g);
        }
        if (const char *dbg = getenv("ALI_EMU_DUMP")) {
            char fn[256];
            snprintf(fn, sizeof fn, "%s_L%d.bin", dbg, l);
            FILE *f = fopen(fn, "wb");
            int hdr[4] = {g.nz, g.nx, s.cz[l & 1], s.cx[l & 1]};
            fwrite(hdr, 4, 4, f);
            fwrite(g.t, 8, (size_t)g.nz * g.nx, f);
            fwrite(g.st, 4, (size_t)g.nz * g.nx, f);
            fclose(f);
        }
    }
    ali_src_main_geometry(s, m, p, sc);
    counters[7] = s.overflow;
    if (s.overflow) return -1;
    ali_seq_clear(s.mg, true, 0, 1);
    if (coop) {
        const int last = (p.nlev - 1) & 1;
        ali_seq_handoff(s.lv[last], s.cz[last], s.cx[last], s.mg, p.isz, p.isx);
        ali_seq_march_coop(s.mg, m, p.isx, p.isz, -1, 0, p.stop_r, s.cnt, 0, g_coop_lanes);
        s.overflow |= s.mg.overflow;
    } else {
        ali_src_main_start_and_seq(s, m, p);
    }
    counters[10] = s.cnt.steps; counters[11] = s.cnt.computed;
    counters[0] = s.cnt.pops; counters[1] = s.cnt.evals; counters[2] = s.cnt.fallbacks;
    counters[7] = s.overflow;
    if (s.overflow) return -1;

    // ---- band-synchronous march of the main grid (replay of the kernel's round loop) ----
    // the kernel's tiled field layout can be replayed too (g_tiled): same result, un-tiled at the end
    const size_t nt = g_tiled ? ali_field_nodes_tiled(p.nz, p.nx) : n;
    std::vector<uint8_t> status(nt, ALI_ST_FAR), dirty(ali_dirty_bytes(p.nz, p.nx), 0);
    std::vector<double> Tt;
    AliBandGrid bg;
    bg.nz = p.nz; bg.nx = p.nx; bg.st = status.data(); bg.dirty = dirty.data(); bg.dnx = m.dnx;
    bg.tiles_x = ali_dirty_tiles_x(p.nx);
    bg.t4x = g_tiled ? (p.nx + 3) / 4 : 0;
    if (g_tiled) {
        Tt.resize(nt);
        std::memset(Tt.data(), ALI_T_UNSET_BYTE, nt * sizeof(double));
        bg.T = Tt.data();
    } else {
        bg.T = T;
    }
    bg.mv = ali_band_view(sg);
    std::vector<unsigned> list;
    seq_to_band(s.mg, bg, list);
    band_run(m, bg, list, delta, 0, eager != 0, mst);
    if (g_tiled)
        for (int z = 0; z < p.nz; z++)
            for (int x = 0; x < p.nx; x++) T[(size_t)z * p.nx + x] = Tt[bg.ti(z, x)];
    for (size_t i = 0; i < n; i++) T[i] = (T[i] >= 0.0) ? T[i] / p.sg : 0.0; // ATR:2832
    counters[3] = mst.rounds; counters[4] = mst.evals; counters[5] = mst.fbs + lst.fbs; counters[6] = mst.maxlist;
    counters[8] = lst.rounds; counters[9] = lst.evals;
    return 0;
}

extern "C" int emu_find_ray(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                            const long long *stif, int has_stif, const double *group_tab, int ncol, double dnx,
                            int sg, const double *rec_ttf, int fz, int fx, double sx, double sy, double rx,
                            double ry, double *ray_x, double *ray_y, int cap, double *time_out, int *flag_out,
                            int nlanes)
{
    HostModel hm;
    make_model(hm, nz, nx, veln, velpn, vel_map, stif, has_stif, group_tab, group_tab, ncol, dnx);
    const AliModel &m = hm.m;
    return ali_emu_trace_ray(m, sg, rec_ttf, fz, fx, sx, sy, rx, ry, ray_x, ray_y, cap, time_out, flag_out, nlanes);
}
