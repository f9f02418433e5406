"""Dev probe: Voronoi 768 lattice sources, GPU vs oracle, optional library override / options."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import build as _b
if os.environ.get("ALIFMM_LIB"):
    _b.LIB_PATH = os.path.abspath(os.environ["ALIFMM_LIB"])
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models
from oracle import ali_oracle as orc
n = 768
m = models.voronoi(n, n * n // 4096, 1234)
scx, scz = models.lattice_sources(n, m["dnx"], rows=4, cols=2)
iz = np.round(scz / m["dnx"]).astype(np.int32); ix = np.round(scx / m["dnx"]).astype(np.int32)
g = np.ones((361, 2)); g[:, 0] = np.arange(361)
ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, g.copy(), m["dnx"])
om = orc.Model(m["veln"], m["velpn"], m["vel_map"], m["stif_den"])
refs = [orc.travel(om, scx[k], scz[k], m["dnx"]) for k in range(8)]
print("lib", _capi.library_path())
for opts in ({}, {"delta_frac": 0.25}, {"delta_frac": 0.1}, {"resort_every": 0}, {"threads_per_source": 256}):
    for k, v in opts.items():
        ctx.set_option(k, v)
    T = ctx.ttf(iz, ix, 1)
    S = ctx.ttf(iz[5:6], ix[5:6], 1)[0]
    out = []
    for k in range(8):
        e = models.rel_err(refs[k], T[k])
        out.append("%d:%.4f/%.1e" % (k, (e <= 1e-5).mean(), e.max()))
    print(opts, " ".join(out), "| single==batch", np.array_equal(S, T[5]))
