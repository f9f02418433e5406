"""Dev probe: per-call breakdown of ALI_FMM.find_all_TTF_rays_parallel over consecutive calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import models
from ali_fmm_and_ray_tracing_b200 import _capi
import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
shim.tqdm_disable = True
acc = {}
def timed(cls, name):
    f = getattr(cls, name)
    def g(*a, **k):
        t = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
    setattr(cls, name, g)
for n in ("__init__", "ttf", "rays_into", "close", "counters", "mem_info"):
    timed(_capi.Context, n)
zs = shim._zeros_sparse
def zs_t(shape):
    t = time.perf_counter(); r = zs(shape); acc["zeros"] = acc.get("zeros", 0.0) + time.perf_counter() - t; return r
shim._zeros_sparse = zs_t
w = models.weld(); scx, scz, pairs = models.weld_headline()
fm = shim.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
for k in range(6):
    acc.clear()
    t0 = time.perf_counter()
    fm.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], subgrid_size=9, trans_pairs=pairs, stif_den=w["stif_den"], n_threads=8)
    dt = time.perf_counter() - t0
    print("call %d: %.3f s | " % (k, dt) + " ".join("%s %.3f" % (n, v) for n, v in sorted(acc.items(), key=lambda kv: -kv[1])) + " | other %.3f" % (dt - sum(acc.values())), flush=True)
