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

// counters: [0] seq pops, [1] seq evals, [2] seq fallbacks, [3] band rounds, [4] band evals,
//           [5] band fallbacks, [6] max list length, [7] overflow flag
extern "C" int emu_ttf(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                       const long long *stif, int has_stif, const double *group_tab, const double *phase_tab,
                       int ncol, double dnx, int src_iz, int src_ix, int sg, int margin, double delta, int eager,
                       double *T, long long *counters)
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
    sc.heap = heap.data(); sc.heap_cap = (int)(cap / 2 + 64); sc.status_cap = cap;
    std::memset(T, ALI_T_UNSET_BYTE, n * sizeof(double)); // NaN = no estimate
    AliSeqResult res;
    ali_seq_source(m, p, sc, T, res, 0, 1);
    counters[0] = res.cnt.pops; counters[1] = res.cnt.evals; counters[2] = res.cnt.fallbacks;
    counters[7] = res.overflow;
    if (res.overflow) return -1;

    // ---- band-synchronous march (replay of ali_march_kernel) ----
    std::vector<uint8_t> status(n, ALI_ST_FAR), dirty(ali_dirty_bytes(p.nz, p.nx), 0);
    std::vector<unsigned> list, next;
    const int32_t *wst = ((p.nlev - 1) & 1) == 0 ? sc.sB : sc.sA;
    for (int z = 0; z < res.wnz; z++)
        for (int x = 0; x < res.wnx; x++) {
            int32_t s = wst[(size_t)z * res.wnx + x];
            size_t node = (size_t)(res.wz0 + z) * p.nx + (res.wx0 + x);
            if (s == 0) status[node] = ALI_ST_ALIVE;
            else if (s > 0) { list.push_back(ALI_PACK(res.wz0 + z, res.wx0 + x)); }
            else { unsigned long long bits = ALI_T_FAR_BITS; std::memcpy(&T[node], &bits, 8); }
        }
    AliBandGrid bg;
    bg.nz = p.nz; bg.nx = p.nx; bg.T = T; bg.st = status.data(); bg.dirty = dirty.data(); bg.dnx = m.dnx;
    bg.tiles_x = ali_dirty_tiles_x(p.nx);
    for (size_t i = 0; i < list.size(); i++) dirty[ali_dirty_index(bg, ALI_PACK_Z(list[i]), ALI_PACK_X(list[i]))] = 1;
    bg.mv = ali_band_view(sg);
    std::vector<double> tnew;
    long long rounds = 0, evals = 0, fbs = 0, maxlist = 0;
    while (!list.empty()) {
        rounds++;
        if ((long long)list.size() > maxlist) maxlist = (long long)list.size();
        tnew.resize(list.size());
        for (size_t i = 0; i < list.size(); i++) {
            const int iz = ALI_PACK_Z(list[i]), ix = ALI_PACK_X(list[i]);
            const size_t node = (size_t)iz * p.nx + ix;
            const size_t di = ali_dirty_index(bg, iz, ix);
            if (dirty[di] || eager) {
                int fb = 0;
                dirty[di] = 0;
                tnew[i] = ali_band_eval(m, &m, bg, sg, iz, ix, &fb);
                if (fb) dirty[di] = 1;
                evals++; fbs += fb;
            } else {
                tnew[i] = T[node];
            }
        }
        double tmin = 1e300;
        for (size_t i = 0; i < list.size(); i++) {
            ali_band_publish(bg, ALI_PACK_Z(list[i]), ALI_PACK_X(list[i]), tnew[i]);
            if (tnew[i] < tmin) tmin = tnew[i];
        }
        const double thr = tmin + delta;
        next.clear();
        for (size_t i = 0; i < list.size(); i++) {
            if (tnew[i] <= thr) {
                unsigned nb[4];
                int cnt = ali_band_accept(bg, ALI_PACK_Z(list[i]), ALI_PACK_X(list[i]), nb);
                for (int k = 0; k < cnt; k++) {
                    next.push_back(nb[k]);   // new nodes are always evaluated next round (kernel: work list)
                    dirty[ali_dirty_index(bg, ALI_PACK_Z(nb[k]), ALI_PACK_X(nb[k]))] = 1;
                }
            } else {
                next.push_back(list[i]);
            }
        }
        list.swap(next);
    }
    for (size_t i = 0; i < n; i++) T[i] = (T[i] >= 0.0) ? T[i] / p.sg : 0.0; // ATR:2832
    counters[3] = rounds; counters[4] = evals; counters[5] = fbs; counters[6] = maxlist;
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
