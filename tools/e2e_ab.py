"""Dev probe: end-to-end call time with the two ways of allocating the dense ray arrays, alternating."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import models
import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
shim.tqdm_disable = True
w = models.weld(); scx, scz, pairs = models.weld_headline()
fm = shim.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
sparse = shim._zeros_sparse
def call():
    return fm.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], subgrid_size=9, trans_pairs=pairs, stif_den=w["stif_den"], n_threads=8)
call(); call()
for k in range(10):
    shim._zeros_sparse = sparse if k % 2 == 0 else (lambda shape: np.zeros(shape))
    t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
    print("call %d %s: %.3f s" % (k, "mmap-nohuge" if k % 2 == 0 else "np.zeros", dt), flush=True)
