"""Generates tests/golden/golden_weld_rays.npz by running the REAL reference on the Weld_rays.py
workload (Weld_rays.py:15-61: 31 + 31 transducers, 31 receiver fields at subgrid 9, 961 rays) through
its own parallel driver, ALI_FMM.find_all_TTF_rays_parallel.  Build container only (needs
/root/reference and numba); ~5 minutes on 8 cores after the first-call JIT.

    python tests/golden/make_golden_weld_rays.py

Stored: times [62, 62] float64, ray_len [62, 62] int32, and the 961 paths packed back to back as float32
(coarse-cell coordinates < 500: 3e-5 cells resolution, the parity gate is 0.1 cell) with their offsets.
BASELINE.md section 2 quotes the scalars of this run (times.sum() = 0.01527291403909612, ...).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from _refload import load_reference  # noqa: E402
import models  # noqa: E402


def main():
    ref = load_reference("/root/reference")
    w = models.weld()
    scx, scz, pairs = models.weld_rays_py()
    obj = ref.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
    t0 = time.time()
    times = obj.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], stif_den=w["stif_den"],
                                           n_threads=os.cpu_count() or 2, trans_pairs=pairs)
    print("reference run: %.1f s" % (time.time() - t0))
    ln = np.asarray(obj.ray_len, dtype=np.int32)
    ii, jj = np.nonzero(ln)
    off = np.zeros(len(ii) + 1, dtype=np.int64)
    off[1:] = np.cumsum(ln[ii, jj])
    px = np.zeros(off[-1], dtype=np.float32)
    py = np.zeros(off[-1], dtype=np.float32)
    for k, (i, j) in enumerate(zip(ii, jj)):
        px[off[k]:off[k + 1]] = obj.ray_paths_x[i, j, :ln[i, j]]
        py[off[k]:off[k + 1]] = obj.ray_paths_y[i, j, :ln[i, j]]
    np.savez_compressed(os.path.join(HERE, "golden_weld_rays.npz"), times=times, ray_len=ln, pair_i=ii.astype(np.int32),
                        pair_j=jj.astype(np.int32), offsets=off, path_x=px, path_y=py)
    print("times.sum() %.17g, nonzero %d, ray_len.sum() %d, min/max %d/%d, times[0,31] %.17g times[30,31] %.17g times[15,46] %.17g" % (
        times.sum(), (times > 0).sum(), ln.sum(), ln[ln > 0].min(), ln.max(), times[0, 31], times[30, 31], times[15, 46]))


if __name__ == "__main__":
    main()
