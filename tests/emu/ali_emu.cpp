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
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_core.cuh"
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_seq.cuh"
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_band.cuh"
#include "../../ali_fmm_and_ray_tracing_b200/csrc/ali_ray.cuh"

static AliModel make_model(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                           const long long *stif, int has_stif, const double *group_tab, const double *phase_tab,
                           int ncol, double dnx)
{
    AliModel m;
    m.nz = nz; m.nx = nx; m.veln = veln; m.velpn = velpn; m.vel_map = vel_map; m.stif = stif;
    m.has_stif = has_stif; m.group_tab = group_tab; m.phase_tab = phase_tab; m.ncol = ncol; m.dnx = dnx;
    return m;
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
    AliModel m = make_model(nz, nx, veln, velpn, vel_map, stif, has_stif, group_tab, phase_tab, ncol, dnx);
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
                       int ncol, double dnx, int src_iz, int src_ix, int sg, int margin, double delta, double *T,
                       long long *counters)
{
    AliModel m = make_model(nz, nx, veln, velpn, vel_map, stif, has_stif, group_tab, phase_tab, ncol, dnx);
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
    std::memset(T, 0, n * sizeof(double));
    AliSeqResult res;
    ali_seq_source(m, p, sc, T, res, 0, 1);
    counters[0] = res.cnt.pops; counters[1] = res.cnt.evals; counters[2] = res.cnt.fallbacks;
    counters[7] = res.overflow;
    if (res.overflow) return -1;

    // ---- band-synchronous march (replay of ali_march_kernel) ----
    std::vector<uint8_t> status(n, ALI_ST_FAR);
    std::vector<int> list, next;
    const int32_t *wst = ((p.nlev - 1) & 1) == 0 ? sc.sB : sc.sA;
    for (int z = 0; z < res.wnz; z++)
        for (int x = 0; x < res.wnx; x++) {
            int32_t s = wst[(size_t)z * res.wnx + x];
            size_t node = (size_t)(res.wz0 + z) * p.nx + (res.wx0 + x);
            if (s == 0) status[node] = ALI_ST_ALIVE;
            else if (s > 0) { status[node] = ALI_ST_BAND; list.push_back((int)node); }
        }
    AliBandGrid bg;
    bg.nz = p.nz; bg.nx = p.nx; bg.T = T; bg.st = status.data(); bg.dnx = m.dnx;
    bg.mv.scale1 = 1; bg.mv.side1 = 0; bg.mv.z0 = 0; bg.mv.x0 = 0;
    bg.mv.scale0 = p.fine ? p.sg : 1; bg.mv.side0 = p.fine ? (p.sg - 1) / 2 : 0; bg.mv.cast = p.fine ? 1 : 0;
    std::vector<double> tnew;
    long long rounds = 0, evals = 0, fbs = 0, maxlist = 0;
    while (!list.empty()) {
        rounds++;
        if ((long long)list.size() > maxlist) maxlist = (long long)list.size();
        tnew.resize(list.size());
        for (size_t i = 0; i < list.size(); i++) {
            int fb = 0;
            tnew[i] = ali_band_eval(m, bg, list[i], &fb);
            evals++; fbs += fb;
        }
        double tmin = 1e300;
        for (size_t i = 0; i < list.size(); i++) {
            ali_band_publish(bg, list[i], tnew[i]);
            if (tnew[i] < tmin) tmin = tnew[i];
        }
        const double thr = tmin + delta;
        next.clear();
        for (size_t i = 0; i < list.size(); i++) {
            int node = list[i];
            if (tnew[i] <= thr) {
                int nb[4];
                int cnt = ali_band_accept(bg, node, nb);
                for (int k = 0; k < cnt; k++) next.push_back(nb[k]);
            } else {
                next.push_back(node);
            }
        }
        list.swap(next);
    }
    if (p.fine)
        for (size_t i = 0; i < n; i++) T[i] = T[i] / p.sg; // ATR:2832
    counters[3] = rounds; counters[4] = evals; counters[5] = fbs; counters[6] = maxlist;
    return 0;
}

extern "C" int emu_find_ray(int nz, int nx, const double *veln, const int32_t *velpn, const double *vel_map,
                            const long long *stif, int has_stif, const double *group_tab, int ncol, double dnx,
                            int sg, const double *rec_ttf, int fz, int fx, double sx, double sy, double rx,
                            double ry, double *ray_x, double *ray_y, int cap, double *time_out, int *flag_out,
                            int nlanes)
{
    AliModel m = make_model(nz, nx, veln, velpn, vel_map, stif, has_stif, group_tab, group_tab, ncol, dnx);
    return ali_emu_trace_ray(m, sg, rec_ttf, fz, fx, sx, sy, rx, ry, ray_x, ray_y, cap, time_out, flag_out, nlanes);
}
