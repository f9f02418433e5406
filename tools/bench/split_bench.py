"""Config 5 (one source on a 16384 x 16384 grid) as two row strips on two GPUs against one GPU.

    python tools/bench/split_bench.py [--n 16384] [--out gpurun_out/split_bench.json]

Needs two GPUs.  Prints one JSON line: the device times of both forms (sequential near-source phase + band march, CUDA
events), wall time of the calls (model upload included), bitwise equality of the two fields and the device memory either
form holds per GPU.  The decomposition buys capacity (each GPU holds half of the model and of the field), not speed."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import models  # noqa: E402
from ali_fmm_and_ray_tracing_b200 import _capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--out", default="")
    ap.add_argument("--gpus", type=int, default=2)
    a = ap.parse_args()
    n = a.n
    m = models.voronoi(n, max(64, (n // 64) ** 2), 1235)
    g = np.ones((361, 2))
    g[:, 0] = np.arange(361)
    src = (n // 4, n // 2)   # (the centre row would make the automatic split asymmetric: keep the source in the upper strip)
    ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, g.copy(), m["dnx"])
    t0 = time.perf_counter()
    one = ctx.ttf(np.array([src[0]], dtype=np.int32), np.array([src[1]], dtype=np.int32), 1)[0]
    w1 = time.perf_counter() - t0
    c1 = ctx.counters()
    ctx.close()
    _capi.trim(0)
    res = []
    for rep in range(2):
        t0 = time.perf_counter()
        two, c2 = _capi.ttf_split(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, g.copy(), m["dnx"], src[0], src[1],
                                  devices=tuple(range(a.gpus)))
        res.append((time.perf_counter() - t0, c2))
    w2, c2 = res[-1]
    nodes = n * n
    per_node_one = 64 + 8 + 8 + 1    # model records + tiled field + row-major result + alive flags
    line = {"workload": "%d x %d Voronoi-grain Christoffel grid, one source at (%d, %d), subgrid 1" % (n, n, src[0], src[1]),
            "bitwise_equal": bool(np.array_equal(one, two)),
            "one_gpu": {"ms_seq": c1["ms_seq"], "ms_march": c1["ms_march"], "rounds": c1["band_rounds"], "cluster_size": c1["cluster_size"],
                        "wall_s_first_call_without_model_upload": w1, "device_bytes": nodes * per_node_one},
            "strips": a.gpus, "two_gpu_strips": {"ms_seq": c2["ms_seq"], "ms_march": c2["ms_march"], "rounds": c2["band_rounds"], "cluster_size": 8,
                               "wall_s_with_model_upload": w2, "device_bytes_per_gpu": (nodes // a.gpus) * per_node_one,
                               "us_per_round": 1e3 * c2["ms_march"] / max(1, c2["band_rounds"])},
            "node_solves_per_s_one_gpu": nodes / (1e-3 * (c1["ms_seq"] + c1["ms_march"])),
            "node_solves_per_s_two_gpu": nodes / (1e-3 * (c2["ms_seq"] + c2["ms_march"]))}
    s = json.dumps(line)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
