"""Dev probe: where the end-to-end time of ALI_FMM.find_all_TTF_rays_parallel goes (headline workload)."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import models
import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
shim.tqdm_disable = True
w = models.weld()
scx, scz, pairs = models.weld_headline()
fm = shim.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
def call():
    return fm.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], subgrid_size=9, trans_pairs=pairs, stif_den=w["stif_den"], n_threads=8)
for k in range(5):
    if k == 1:
        pr1 = cProfile.Profile(); pr1.enable()
    t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
    if k == 1:
        pr1.disable(); pstats.Stats(pr1).sort_stats("tottime").print_stats(8)
    c = fm.last_counters[0]
    print("call %d: %.3f s (device: seq %.0f march %.0f fin %.0f rays %.0f ms)" % (k, dt, c["ms_seq"], c["ms_march"], c["ms_finalize"], c["ms_rays"]), flush=True)
pr = cProfile.Profile(); pr.enable(); call(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
c = fm.last_counters[0]
print({k: c[k] for k in c if k.startswith("ms_")})
