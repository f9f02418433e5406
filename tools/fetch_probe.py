"""Dev probe: cost of returning fields to the host (alifmm_ttf with an output pointer)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models
w = models.weld(); scx, scz, pairs = models.weld_headline()
iz = np.round(scz / w["dnx"]).astype(np.int32)[:16]; ix = np.round(scx / w["dnx"]).astype(np.int32)[:16]
g = np.ones((361, 2)); g[:, 0] = np.arange(361)
ctx = _capi.Context(w["veln"], w["velpn"], w["vel_map"], w["stif_den"], True, g, g.copy(), w["dnx"])
for k in range(3):
    t0 = time.perf_counter(); ctx.ttf(iz, ix, 9, fetch=False); t1 = time.perf_counter()
    T = ctx.ttf(iz, ix, 9); t2 = time.perf_counter()
    print("16 fields (%.2f GB): resident %.3f s, with fetch %.3f s -> D2H+alloc %.3f s (%.1f GB/s)" % (T.nbytes / 1e9, t1 - t0, t2 - t1, (t2 - t1) - (t1 - t0), T.nbytes / 1e9 / ((t2 - t1) - (t1 - t0))))
    del T
