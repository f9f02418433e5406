"""Dev probe: FMC all-pairs on the weld at subgrid 3 against the oracle (times and path distances)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests import models
from oracle import ali_oracle as orc
import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
shim.tqdm_disable = True
w = models.weld(); dnx = w["dnx"]
xs = [33, 243, 467]
scx = np.array([dnx * x for x in xs] * 2); scz = np.array([0.0] * 3 + [dnx * 423] * 3)
sg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
fm = shim.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=dnx)
pairs = np.ones((6, 6)) - np.eye(6)
times = fm.find_all_TTF_rays(w["veln"], w["velpn"], w["vel_map"], subgrid_size=sg, trans_pairs=pairs, stif_den=w["stif_den"])
om = orc.Model(w["veln"], w["velpn"], w["vel_map"], w["stif_den"])
for j in range(6):
    ref_T = orc.travel_finer_grid(om, scx[j], scz[j], dnx, sg)
    for i in range(6):
        if i == j: continue
        rx, ry, rt, fl = orc.find_ray(om, dnx, (sg * fm.isx[i], sg * fm.isz[i]), (sg * fm.isx[j], sg * fm.isz[j]), ref_T, sg)
        x, y = fm.ray_path(i, j)
        print(i, j, "len", len(x), len(rx), "t rel %.2e" % (abs(times[i, j] - rt) / rt), "dist %.3f" % models.polyline_distance(x, y, rx / sg, ry / sg), "flags", fm.ray_flags[i, j], fl)
