"""Dev probe: edge/corner sources in the homogeneous Christoffel medium at several delta_frac; optional library override."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import build as _b
if os.environ.get("ALIFMM_LIB"):
    _b.LIB_PATH = os.path.abspath(os.environ["ALIFMM_LIB"])
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models
from oracle import ali_oracle as orc
m = models.notebook_christoffel(101)
m["veln"] = 35.0 * np.ones((101, 101))
pts = [(0, 50), (100, 50), (50, 0), (50, 100), (1, 50), (50, 1), (99, 99), (2, 97), (100, 0)]
g = np.ones((361, 2)); g[:, 0] = np.arange(361)
ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, g.copy(), m["dnx"])
iz = np.array([p[0] for p in pts], dtype=np.int32); ix = np.array([p[1] for p in pts], dtype=np.int32)
om = orc.Model(m["veln"], m["velpn"], m["vel_map"], m["stif_den"])
refs = [orc.travel(om, m["dnx"] * ix[k], m["dnx"] * iz[k], m["dnx"]) for k in range(len(pts))]
print("lib", _capi.library_path())
for frac in (0.25, 0.3, 0.35, 0.4):
    ctx.set_option("delta_frac", frac)
    T = ctx.ttf(iz, ix, 1)
    for k in range(len(pts)):
        e = models.rel_err(refs[k], T[k])
        print("frac %.2f src %s: <=1e-5 %.4f  max %.2e  median %.1e bitexact %.3f" % (frac, pts[k], (e <= 1e-5).mean(), e.max(), np.median(e), (refs[k] == T[k]).mean()))
