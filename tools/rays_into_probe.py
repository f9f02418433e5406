"""Dev probe: where alifmm_rays_into spends its host time (fresh vs touched destination arrays)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ali_fmm_and_ray_tracing_b200 import _capi
from tests import models
w = models.weld(); scx, scz, pairs = models.weld_headline()
iz = np.round(scz / w["dnx"]).astype(np.int32); ix = np.round(scx / w["dnx"]).astype(np.int32)
g = np.ones((361, 2)); g[:, 0] = np.arange(361)
ctx = _capi.Context(w["veln"], w["velpn"], w["vel_map"], w["stif_den"], True, g, g.copy(), w["dnx"])
ctx.ttf(iz, ix, 9, fetch=False)
ri, rj = np.nonzero(pairs)
cap = 5 * (424 + 500)
rows = ri * 128 + rj
print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
for trial in range(3):
    t0 = time.perf_counter(); bx = np.zeros((128, 128, cap)); by = np.zeros((128, 128, cap)); t1 = time.perf_counter()
    ln, tm, fl = ctx.rays_into(iz[ri], ix[ri], rj.astype(np.int32), cap, 9, rows, bx, by); t2 = time.perf_counter()
    ln, tm, fl = ctx.rays_into(iz[ri], ix[ri], rj.astype(np.int32), cap, 9, rows, bx, by); t3 = time.perf_counter()
    print("zeros %.3f s | rays_into fresh %.3f s | again (touched) %.3f s | kernel %.1f ms" % (t1 - t0, t2 - t1, t3 - t2, ctx.counters()["ms_rays"]))
    del bx, by
