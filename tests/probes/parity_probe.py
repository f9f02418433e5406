"""Dev probe: GPU fields vs the oracle (glibc) and vs the host replay run on the device's math (bit-exactness)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models
from tests.emu import emu
from oracle import ali_oracle as orc
from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import ALI_FMM

def run(name, m, srcs, sg=1):
    g, p = (m["group_vel"], m["phase_vel"]) if m.get("group_vel") is not None else (None, None)
    if g is None:
        g = np.ones((361, 2)); g[:, 0] = np.arange(361); p = g.copy()
    stif = m["stif_den"] if m["stif_den"] is not None else np.zeros(m["veln"].shape + (5,), dtype=np.int64)
    ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], stif, True, g, p, m["dnx"])
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], stif, m.get("group_vel"), m.get("phase_vel"))
    iz = np.array([s[0] for s in srcs], dtype=np.int32); ix = np.array([s[1] for s in srcs], dtype=np.int32)
    T = ctx.ttf(iz, ix, sg)
    for k, s in enumerate(srcs):
        ref = orc.travel_finer_grid(om, m["dnx"] * s[1], m["dnx"] * s[0], m["dnx"], sg) if sg > 1 else orc.travel(om, m["dnx"] * s[1], m["dnx"] * s[0], m["dnx"])
        emu.set_crmath(True); R, _, _ = emu.ttf(om, m["dnx"], s[0], s[1], sg); emu.set_crmath(False)
        e = models.rel_err(ref, T[k]); d = models.rel_err(R, T[k])
        print("%-10s src %-10s vs oracle: <=1e-5 %.5f max %.1e bitexact %.4f | vs replay(device math): bitexact %.6f max %.1e" % (
            name, s, (e <= 1e-5).mean(), e.max(), (ref == T[k]).mean(), (R == T[k]).mean(), d.max()), flush=True)
    ctx.close()

m = models.notebook_christoffel(101); m["veln"] = 35.0 * np.ones((101, 101))
run("christ101", m, [(0, 50), (50, 0), (99, 99), (2, 97), (100, 0), (50, 50)])
run("nb_table", models.notebook_table(ALI_FMM), [(100, 1), (140, 199)])
run("nb_christ", models.notebook_christoffel(), [(100, 1), (140, 199), (1, 100)])
run("nb_grad", models.notebook_gradient(), [(30, 1), (180, 199)])
n = 768; v = models.voronoi(n, n * n // 4096, 1234)
scx, scz = models.lattice_sources(n, v["dnx"], rows=4, cols=2)
run("voronoi768", v, [(int(round(z / v["dnx"])), int(round(x / v["dnx"]))) for x, z in zip(scx, scz)])
c = models.weld_crop(120, 160)
run("weldcrop", c, [(0, 40), (119, 100)], sg=3)
run("weld sg1", models.weld(), [(0, 27), (423, 300)])
