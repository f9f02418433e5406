"""Dev probe: is the field independent of the acceptance band?  GPU fields at delta_frac 0.35 / 0.4 vs 0.3, bitwise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models
from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import ALI_FMM

def run(name, m, srcs, sg=1):
    g, p = (m["group_vel"], m["phase_vel"]) if m.get("group_vel") is not None else (None, None)
    if g is None:
        g = np.ones((361, 2)); g[:, 0] = np.arange(361); p = g.copy()
    stif = m["stif_den"] if m["stif_den"] is not None else np.zeros(m["veln"].shape + (5,), dtype=np.int64)
    ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], stif, True, g, p, m["dnx"])
    iz = np.array([s[0] for s in srcs], dtype=np.int32); ix = np.array([s[1] for s in srcs], dtype=np.int32)
    out = {}
    for f in (0.3, 0.2, 0.35, 0.4):
        ctx.set_option("delta_frac", f)
        out[f] = ctx.ttf(iz, ix, sg)
        c = ctx.counters()
        print("%-10s frac %.2f rounds_max %d march %.1f ms" % (name, f, c["band_rounds_max"], c["ms_march"]), end=" | ")
        print(" ".join("%s:%s/%.0e" % (srcs[k], "same" if np.array_equal(out[f][k], out[0.3][k]) else "DIFF", models.rel_err(out[0.3][k], out[f][k]).max()) for k in range(len(srcs))) if f != 0.3 else "", flush=True)
    ctx.close()

m = models.notebook_christoffel(101); m["veln"] = 35.0 * np.ones((101, 101))
run("christ101", m, [(0, 50), (100, 50), (50, 0), (50, 100), (1, 50), (50, 1), (99, 99), (2, 97), (100, 0)])
run("nb_table", models.notebook_table(ALI_FMM), [(100, 1), (140, 199)])
run("nb_christ", models.notebook_christoffel(), [(100, 1), (140, 199), (1, 100)])
run("nb_grad", models.notebook_gradient(), [(30, 1), (180, 199)])
n = 768; v = models.voronoi(n, n * n // 4096, 1234)
scx, scz = models.lattice_sources(n, v["dnx"], rows=4, cols=2)
run("voronoi768", v, [(int(round(z / v["dnx"])), int(round(x / v["dnx"]))) for x, z in zip(scx, scz)])
run("weldcrop3", models.weld_crop(120, 160), [(0, 40), (119, 100), (60, 80)], sg=3)
run("weld sg1", models.weld(), [(0, 27), (423, 300), (0, 250), (423, 100)])
w = models.weld(); sx, sz, _ = models.weld_headline()
run("weld sg9", w, [(0, 27), (423, 27 + 7 * 30)], sg=9)
