"""Dev probe: runs weld fields / rays through the C ABI and prints timings, counters, parity."""
import argparse
import sys
import time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import build as _b
if os.environ.get("ALIFMM_LIB"):
    _b.LIB_PATH = os.path.abspath(os.environ["ALIFMM_LIB"])
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models

ap = argparse.ArgumentParser()
ap.add_argument("--sg", type=int, default=9)
ap.add_argument("--nsrc", type=int, default=8)
ap.add_argument("--frac", type=float, default=0.35)
ap.add_argument("--margin", type=int, default=27)
ap.add_argument("--threads", type=int, default=768)
ap.add_argument("--check", type=int, default=1, help="number of fields compared with the oracle")
ap.add_argument("--rays", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--smemkb", type=int, default=-1)
ap.add_argument("--resort", type=int, default=-1)
ap.add_argument("--cluster", type=int, default=-1)
ap.add_argument("--cthreads", type=int, default=-1)
ap.add_argument("--seqthreads", type=int, default=-1)
a = ap.parse_args()

w = models.weld()
scx, scz, pairs = models.weld_headline()
n = a.nsrc
sel = np.linspace(0, len(scx) - 1, n).round().astype(int) if n < len(scx) else np.arange(len(scx))
iz = np.round(scz[sel] / w["dnx"]).astype(np.int32)
ix = np.round(scx[sel] / w["dnx"]).astype(np.int32)
g = np.ones((361, 2)); g[:, 0] = np.arange(361)
t0 = time.time()
ctx = _capi.Context(w["veln"], w["velpn"], w["vel_map"], w["stif_den"], True, g, g.copy(), w["dnx"])
ctx.set_option("delta_frac", a.frac); ctx.set_option("handover_margin", a.margin); ctx.set_option("threads_per_source", a.threads)
if a.smemkb >= 0: ctx.set_option("band_smem_kb", a.smemkb)
if a.resort >= 0: ctx.set_option("resort_every", a.resort)
if a.cluster >= 0: ctx.set_option("cluster_size", a.cluster)
if a.cthreads >= 0: ctx.set_option("cluster_threads", a.cthreads)
if a.seqthreads >= 0: ctx.set_option("seq_threads", a.seqthreads)
print("create %.3f s, mem" % (time.time() - t0), ctx.mem_info())
for rep in range(a.reps):
    t0 = time.time()
    ctx.ttf(iz, ix, a.sg, fetch=False)
    dt = time.time() - t0
    c = ctx.counters()
    print("cluster %d | seq Mcyc %.0f..%.0f march Mcyc %.0f..%.0f" % (c["cluster_size"], c["seq_mcycles_min"], c["seq_mcycles_max"], c["march_mcycles_min"], c["march_mcycles_max"]))
    print("ttf wall %.3f s: seq %.1f ms march %.1f ms fin %.1f ms | node_solves %.3e -> %.3e /s | seq_pops %d band_rounds_max %d band_evals %.3e (%.2f/node) max_band %d fallbacks %d" % (
        dt, c["ms_seq"], c["ms_march"], c["ms_finalize"], c["node_solves"], c["node_solves"] / dt, c["seq_pops"], c["band_rounds_max"],
        c["band_evals"], c["band_evals"] / c["node_solves"], c["max_band"], c["fallback_evals"]), flush=True)
if a.check:
    from oracle import ali_oracle as orc
    om = orc.Model(w["veln"], w["velpn"], w["vel_map"], w["stif_den"])
    for k in range(min(a.check, n)):
        T = ctx.ttf_fetch(k)
        t0 = time.time()
        ref = orc.travel_finer_grid(om, scx[sel[k]], scz[sel[k]], w["dnx"], a.sg) if a.sg > 1 else orc.travel(om, scx[sel[k]], scz[sel[k]], w["dnx"])
        e = models.rel_err(ref, T)
        print("field %d src (z=%d,x=%d): oracle %.1f s; max rel %.3e; >1e-5 %.5f; >1e-7 %.5f; >1e-12 %.5f; bitexact %.4f" % (
            k, iz[k], ix[k], time.time() - t0, e.max(), (e > 1e-5).mean(), (e > 1e-7).mean(), (e > 1e-12).mean(), (ref == T).mean()), flush=True)
if a.rays:
    rng = np.random.default_rng(0)
    rs = rng.integers(0, n, a.rays); rr = (rs + 1 + rng.integers(0, n - 1, a.rays)) % n
    keep = None
    for mb in (4, 5, 6, 4):
        ctx.set_option("ray_min_blocks", mb)
        t0 = time.time()
        x, y, ln, tm, fl = ctx.rays(iz[rs], ix[rs], rr.astype(np.int32))
        dt = time.time() - t0
        c = ctx.counters()
        same = "" if keep is None else " same as first: %s" % (np.array_equal(keep[0], x) and np.array_equal(keep[1], tm) and np.array_equal(keep[2], ln))
        if keep is None:
            keep = (x.copy(), tm.copy(), ln.copy())
        print("rays (min blocks %d) wall %.3f s kernel %.1f ms: %d rays, %d points, flags %s%s" % (mb, dt, c["ms_rays"], a.rays, ln.sum(), np.bincount(fl), same), flush=True)
    if a.check:
        for r in range(min(3, a.rays)):
            T = ctx.ttf_fetch(int(rr[r]))
            ox, oy, ot, of = orc.find_ray(om, w["dnx"], (a.sg * ix[rs[r]], a.sg * iz[rs[r]]), (a.sg * ix[rr[r]], a.sg * iz[rr[r]]), T, a.sg)
            dev = models.polyline_distance(x[r, :ln[r]] / a.sg, y[r, :ln[r]] / a.sg, ox / a.sg, oy / a.sg)
            print("ray %d: len %d vs %d, time %.9e vs %.9e, dev %.3e cells, flags %d/%d" % (r, ln[r], len(ox), tm[r], ot, dev, fl[r], of))
ctx.close()
