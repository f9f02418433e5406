// ali_ray.cuh -- building blocks of the plane-marching Fermat ray tracer.
//
// Reference: find_ray (ATR:3104-3465).  A ray is advanced one coarse cell per step: a
// search plane (x = c, y = c or a diagonal) is placed sg fine nodes ahead of the last
// point; for every integer node of the plane within reach the receiver-sourced travel
// time plus the straight-ray time from the last point is evaluated; a quadratic through
// each local minimum gives the next (fractional) point.  The candidates of a plane are
// independent (one lane each); the plane choice and the selection are uniform.
#pragma once
#include "ali_core.cuh"

#define ALI_RAY_TTF_INCREASING 1 // reference prints "Travel time to receiver increasing" (ATR:3407)
#define ALI_RAY_LEFT_GRID 2      // plane left the grid (ATR:3172-3173, 3294-3295)
#define ALI_RAY_EMPTY_PLANE 4    // no candidate on the plane (reference: undefined behaviour)
#define ALI_RAY_CAPACITY 8       // path buffer full (reference: no check)

struct AliRayState {
    double last_x, last_y, lvx, lvy; // last point and last segment vector (fine-grid units)
    double rx, ry;                   // receiver
    int len;                         // points stored so far
    int flag;
    int done;
};

struct AliRayPlane {
    int dir;     // 0: x = c, 1: y = -x + c, 2: y = c, 3: y = x + c
    int c_value;
    int lo;      // first candidate coordinate along the plane (row for dir 0, x otherwise)
    int len;     // number of candidates
};

ALI_HD int ali_ray_max_candidates(int sg) { return 6 * sg + 3; }

// True while the tracer must take another step (ATR:3156).
ALI_DEV bool ali_ray_continue(const AliRayState &s, int sg)
{
    double dx = s.last_x - s.rx, dy = s.last_y - s.ry;
    return dx * dx + dy * dy > (1.6 * sg) * (1.6 * sg);
}

// Chooses the search plane and its candidate range (ATR:3157-3184, 3222-3233, 3286-3306,
// 3343-3354).  fz/fx: rows/cols of the fine field (the reference calls them nnx/nnz).
// Returns false when the march ends here (flag set).
ALI_DEV bool ali_ray_choose_plane(AliRayState &s, int sg, int fz, int fx, AliRayPlane &pl)
{
    const int search_dist = 3 * sg + 1, search_dist_2 = 2 * sg + 1;
    const int nnx = fz, nnz = fx;
    double dx = s.last_x - s.rx, dy = s.last_y - s.ry;
    if (dx * dx + dy * dy < (double)((4 * sg) * (4 * sg))) {
        s.lvx = s.rx - s.last_x;
        s.lvy = s.ry - s.last_y;
    }
    const double r2 = sqrt(2.0);
    double v0 = fabs(s.lvx), v1 = fabs(s.lvx + s.lvy) / r2, v2 = fabs(s.lvy), v3 = fabs(s.lvx - s.lvy) / r2;
    int dir = 0;
    double best = v0;
    if (v1 > best) { best = v1; dir = 1; }
    if (v2 > best) { best = v2; dir = 2; }
    if (v3 > best) { best = v3; dir = 3; }
    const int rlx = (int)rint(s.last_x), rly = (int)rint(s.last_y);
    int c, lo, hi;
    if (dir == 0) {
        c = rlx;
        if (s.lvx > 0) c += sg; else c -= sg;
        if (c < 0 || c >= nnz) { s.flag |= ALI_RAY_LEFT_GRID; return false; }
        lo = ali_imax2(0, rly - search_dist);
        hi = ali_imin2(nnx - 1, rly + search_dist);
    } else if (dir == 1) {
        c = rlx + rly;
        if (s.lvx > 0) {
            c += sg;
            lo = ali_imax2(ali_imax2(0, c - (nnx - 1)), rlx - search_dist_2);
            hi = ali_imin2(ali_imin2(nnz - 1, c), c - rly + search_dist_2);
        } else {
            c -= sg;
            lo = ali_imax2(ali_imax2(0, c - (nnx - 1)), c - rly - search_dist_2);
            hi = ali_imin2(ali_imin2(nnz - 1, c), rlx + search_dist_2);
        }
    } else if (dir == 2) {
        c = rly;
        if (s.lvy > 0) c += sg; else c -= sg;
        if (c < 0 || c >= nnx) { s.flag |= ALI_RAY_LEFT_GRID; return false; }
        lo = ali_imax2(0, rlx - search_dist);
        hi = ali_imin2(nnz - 1, rlx + search_dist);
    } else {
        c = rly - rlx;
        if (s.lvx < 0) {
            c += sg;
            lo = ali_imax2(ali_imax2(0, -c), rly - c - search_dist_2);
            hi = ali_imin2(ali_imin2(nnz - 1, (nnx - 1) - c), rlx + search_dist_2);
        } else {
            c -= sg;
            lo = ali_imax2(ali_imax2(0, -c), rlx - search_dist_2);
            hi = ali_imin2(ali_imin2(nnz - 1, (nnx - 1) - c), rly - c + search_dist_2);
        }
    }
    pl.dir = dir; pl.c_value = c; pl.lo = lo; pl.len = hi - lo + 1;
    if (pl.len < 1) { s.flag |= ALI_RAY_EMPTY_PLANE; return false; }
    return true;
}

// Fine-grid coordinates of candidate i of a plane.
ALI_DEV void ali_ray_candidate_xy(const AliRayPlane &pl, int i, int &cx, int &cy)
{
    int v = pl.lo + i;
    if (pl.dir == 0) { cx = pl.c_value; cy = v; }
    else if (pl.dir == 1) { cx = v; cy = -v + pl.c_value; }
    else if (pl.dir == 2) { cx = v; cy = pl.c_value; }
    else { cx = v; cy = v + pl.c_value; }
}

// TT[i] = rec_TTF[candidate] + straight-ray time from the last point (ATR:3189, 3252, 3312, 3373).
ALI_DEV double ali_ray_candidate_time(const AliModel &m, const double *rec, int fx, const AliRayPlane &pl, int i,
                                      double last_x, double last_y, int sg)
{
    int cx, cy;
    ali_ray_candidate_xy(pl, i, cx, cy);
    double t = rec[(size_t)cy * fx + cx];
    return t + ali_time_between_points(m, last_x, (double)cx, last_y, (double)cy, sg, 16 * sg + 64);
}

// Quadratic fit at interior candidate j (ATR:3199-3214): value of the fitted minimum, or
// +inf when j is not a local minimum.  pos is the fractional index of the minimum.
ALI_DEV double ali_ray_local_min(const double *TT, int j, double &pos)
{
    double t1 = TT[j - 1], t2 = TT[j], t3 = TT[j + 1];
    if (!(t1 >= t2 && t2 <= t3)) { pos = 0; return 1e300; }
    double a = (t1 + t3 - 2 * t2) / 2, b = (t3 - t1) / 2, val;
    if (a != 0) {
        pos = -b / (2 * a);
        val = a * (pos * pos) + b * pos + t2;
        pos += j;
    } else {
        pos = j; val = t2;
    }
    return val;
}

// Final selection (ATR:3192-3218): end points first, then the first strictly smaller
// fitted minimum in index order.
ALI_DEV double ali_ray_select(const double *TT, const double *vals, const double *poss, int len)
{
    double minimum, min_i;
    if (TT[0] < TT[len - 1]) { minimum = TT[0]; min_i = 0; }
    else { minimum = TT[len - 1]; min_i = len - 1; }
    for (int j = 1; j < len - 1; j++)
        if (vals[j] < minimum) { min_i = poss[j]; minimum = vals[j]; }
    return min_i;
}

// Applies the selected position: new point, TTF monotonicity test (ATR:3406-3428).
// Returns false when the ray ends early.  The candidate point is written at ray[len]
// either way, as in the reference (it is overwritten by the receiver on early exit).
ALI_DEV bool ali_ray_advance(AliRayState &s, const AliRayPlane &pl, double min_i, const double *rec, int fx,
                             double *ray_x, double *ray_y)
{
    double nx_, ny_;
    if (pl.dir == 0) { nx_ = pl.c_value; ny_ = min_i + pl.lo; }
    else if (pl.dir == 1) { nx_ = pl.lo + min_i; ny_ = pl.c_value - nx_; }
    else if (pl.dir == 2) { nx_ = min_i + pl.lo; ny_ = pl.c_value; }
    else { nx_ = pl.lo + min_i; ny_ = nx_ + pl.c_value; }
    ray_x[s.len] = nx_;
    ray_y[s.len] = ny_;
    if (rec[(size_t)((int)rint(s.last_y)) * fx + (int)rint(s.last_x)] <
        rec[(size_t)((int)rint(ny_)) * fx + (int)rint(nx_)]) {
        s.flag |= ALI_RAY_TTF_INCREASING;
        return false;
    }
    s.lvx = nx_ - s.last_x; s.last_x = nx_;
    s.lvy = ny_ - s.last_y; s.last_y = ny_;
    s.len += 1;
    return true;
}
