#!/usr/bin/env python
"""Parity survey (GPU box): every BASELINE config's fields / rays against the oracle, with the
deviation classification of tests/parity_tools.py.  Writes gpurun_out/parity_survey.json; PARITY.md is
written from it.  Test infrastructure: uses the oracle as the checker.

    python tools/parity_survey.py [--quick]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import models, parity_tools  # noqa: E402
from oracle import ali_oracle as orc  # noqa: E402
from ali_fmm_and_ray_tracing_b200 import _capi  # noqa: E402


def tables(m):
    if m.get("group_vel") is not None:
        return m["group_vel"], m["phase_vel"]
    g = np.ones((361, 2))
    g[:, 0] = np.arange(361)
    return g, g.copy()


def omodel(m):
    stif = m["stif_den"] if m["stif_den"] is not None else np.zeros(m["veln"].shape + (5,), dtype=np.int64)
    return orc.Model(m["veln"], m["velpn"], m["vel_map"], stif, m.get("group_vel"), m.get("phase_vel"))


_W = {}


def _oracle_field(args):
    """(the shipped reference's field, the reference algorithm's field on a heap that orders correctly beyond the
    hand-over radius -- oracle.set_true_heap_after)"""
    name, sz, sx, sg = args
    m = _W[name]
    om = omodel(m)
    run = (lambda: orc.travel_finer_grid(om, m["dnx"] * sx, m["dnx"] * sz, m["dnx"], sg)) if sg > 1 else \
        (lambda: orc.travel(om, m["dnx"] * sx, m["dnx"] * sz, m["dnx"]))
    ref = run()
    orc.set_true_heap_after((13 if sg == 1 else 5 * sg + (sg - 1) // 2) + 27)
    try:
        fixed = run()
    finally:
        orc.set_true_heap_after(-1)
    return ref, fixed


def field_stats(m, both, T, sg, src):
    ref, fixed = both
    e = models.rel_err(ref, T)
    ef = models.rel_err(fixed, T)
    box = 13 if sg == 1 else (5 * sg + (sg - 1) // 2)
    cls = parity_tools.classify_deviations(orc, m, ref, T, sg=sg, source=(src[0] * (sg if sg > 1 else 1), src[1] * (sg if sg > 1 else 1)), box=box)
    return {"source_zx": [int(src[0]), int(src[1])], "nodes": int(ref.size), "bit_equal": float((ref == T).mean()),
            "frac_gt_1e-5": float((e > 1e-5).mean()), "frac_gt_1e-9": float((e > 1e-9).mean()), "p99": float(np.quantile(e, 0.99)),
            "max": float(e.max()), "bit_equal_correct_heap": float((fixed == T).mean()), "max_correct_heap": float(ef.max()),
            "patches": cls.get("patches", 0), "roots_checked": cls.get("roots_checked", 0),
            "roots_glitch": cls["roots_glitch"], "roots_in_source_box": cls["roots_in_source_box"], "unexplained": cls["unexplained"],
            "unexplained_samples": cls.get("unexplained_samples", [])}


def survey_fields(name, m, sources, sg, pool):
    assert _W[name] is m   # (registered before the worker pool forks)
    g, p = tables(m)
    ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, p, m["dnx"])
    iz = np.array([s[0] for s in sources], dtype=np.int32)
    ix = np.array([s[1] for s in sources], dtype=np.int32)
    t0 = time.time()
    T = ctx.ttf(iz, ix, sg)
    t_gpu = time.time() - t0
    ctx.close()
    t0 = time.time()
    refs = pool.map(_oracle_field, [(name, int(s[0]), int(s[1]), sg) for s in sources]) if pool else \
        [_oracle_field((name, int(s[0]), int(s[1]), sg)) for s in sources]
    t_cpu = time.time() - t0
    out = {"config": name, "subgrid": sg, "gpu_s": t_gpu, "oracle_s": t_cpu, "fields": [field_stats(m, refs[k], T[k], sg, sources[k]) for k in range(len(sources))]}
    print(json.dumps(out), flush=True)
    return out


def _oracle_rays(args):
    mname, j, srcs, sg, scx, scz = args
    m = _W[mname]
    om = omodel(m)
    T = orc.travel_finer_grid(om, scx[j], scz[j], m["dnx"], sg)
    rec = (sg * round(scx[j] / m["dnx"]), sg * round(scz[j] / m["dnx"]))
    out = []
    for i in srcs:
        src = (sg * round(scx[i] / m["dnx"]), sg * round(scz[i] / m["dnx"]))
        rx, ry, t, fl = orc.find_ray(om, m["dnx"], src, rec, T, sg)
        out.append((int(i), rx / sg, ry / sg, t))
    return j, out


def survey_rays(name, mname, scx, scz, pairs, sg, pool, golden=None):
    """find_all_TTF_rays_parallel through the class against oracle rays through ORACLE fields (end to end)."""
    from Anis_TTF_rays import ALI_FMM
    import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
    shim.tqdm_disable = True
    m = _W[mname]   # (models are registered before the worker pool forks)
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], scx, scz, group_vel=m.get("group_vel"), phase_vel=m.get("phase_vel"),
                 stif_den=m["stif_den"], dnx=m["dnx"])
    t0 = time.time()
    times = fm.find_all_TTF_rays_parallel(m["veln"], m["velpn"], m["vel_map"], subgrid_size=sg, trans_pairs=pairs,
                                          stif_den=m["stif_den"], n_threads=8)
    t_gpu = time.time() - t0
    recs = [j for j in range(len(scx)) if pairs[:, j].sum() > 0]
    jobs = [(mname, j, [int(i) for i in np.nonzero(pairs[:, j])[0] if i != j], sg, np.asarray(scx), np.asarray(scz)) for j in recs]
    t0 = time.time()
    res = pool.map(_oracle_rays, jobs) if pool else [_oracle_rays(a) for a in jobs]
    t_cpu = time.time() - t0
    devs, trel, lens = [], [], 0
    ref_sum = 0.0
    for j, rays in res:
        for i, rx, ry, t in rays:
            x, y = fm.ray_path(i, j)
            devs.append(models.polyline_distance(x, y, rx, ry))
            trel.append(abs(times[i, j] - t) / t)
            lens += len(rx)
            ref_sum += t
    devs, trel = np.array(devs), np.array(trel)
    out = {"config": name, "subgrid": sg, "rays": int(len(devs)), "gpu_s": t_gpu, "oracle_s": t_cpu,
           "within_0.1_cell": int((devs <= 0.1).sum()), "within_0.01_cell": int((devs <= 0.01).sum()), "max_dev_cells": float(devs.max()),
           "median_dev_cells": float(np.median(devs)), "max_time_rel": float(trel.max()), "times_sum": float(times.sum()),
           "oracle_times_sum": ref_sum, "ray_len_sum": int(fm.ray_len.sum()), "oracle_ray_len_sum": int(lens),
           "ray_len_min": int(fm.ray_len[fm.ray_len > 0].min()), "ray_len_max": int(fm.ray_len.max())}
    if golden:
        out["golden"] = {k: float(times[idx]) for k, idx in golden.items()}
    print(json.dumps(out), flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_survey.json"))
    args = ap.parse_args()
    orc.build()
    from Anis_TTF_rays import ALI_FMM
    cores = min(os.cpu_count() or 1, 32)
    results = {"fields": [], "rays": [], "cores": cores}
    w = models.weld()
    nb = {"nb_gradient": models.notebook_gradient(), "nb_christoffel": models.notebook_christoffel(), "nb_table": models.notebook_table(ALI_FMM)}
    vor = models.voronoi(768, 144, 1234)
    sx, sz = models.lattice_sources(768, vor["dnx"], rows=4, cols=4)
    vsrc = [(int(round(z / vor["dnx"])), int(round(x / vor["dnx"]))) for x, z in zip(sx, sz)]
    _W.update(nb)
    _W.update(weld=w, vor768=vor)
    big = None
    if not args.quick:
        big = models.voronoi(4096, 4096, 1234)
        _W["vor4096"] = big
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_kernel_replay as tk
    strip = tk._strip_model(16384, 192, 11)
    _W["strip16384"] = strip
    with mp.get_context("fork").Pool(cores) as pool:
        for name, m in nb.items():
            src = [(int(round(z / m["dnx"])), int(round(x / m["dnx"]))) for x, z in zip(m["scx"], m["scz"])]
            results["fields"].append(survey_fields(name, m, src, 1, pool))
            results["fields"].append(survey_fields(name, m, src[:2], 9, pool))
        results["fields"].append(survey_fields("weld", w, [(0, 25), (0, 250), (200, 250), (423, 160)], 1, pool))
        results["fields"].append(survey_fields("weld", w, [(423, 160), (0, 27), (0, 250), (423, 468), (0, 475), (423, 27), (0, 132), (423, 300)], 9, pool))
        results["fields"].append(survey_fields("vor768", vor, vsrc, 1, pool))
        results["fields"].append(survey_fields("strip16384", strip, [(8192, 96)], 1, pool))
        if big is not None:
            bx, bz = models.lattice_sources(4096, big["dnx"])
            bsrc = [(int(round(z / big["dnx"])), int(round(x / big["dnx"]))) for x, z in zip(bx, bz)]
            results["fields"].append(survey_fields("vor4096", big, bsrc[::16], 1, pool))
        # rays, end to end
        scx, scz, pairs = models.weld_rays_py()
        results["rays"].append(survey_rays("weld_rays_py", "weld", scx, scz, pairs, 9, pool,
                                           golden={"times[0,31]": (0, 31), "times[30,31]": (30, 31), "times[15,46]": (15, 46)}))
        m = nb["nb_christoffel"]
        results["rays"].append(survey_rays("nb_christoffel", "nb_christoffel", m["scx"], m["scz"], np.triu(np.ones((3, 3)), 1), 9, pool))
        m = nb["nb_gradient"]
        results["rays"].append(survey_rays("nb_gradient", "nb_gradient", m["scx"], m["scz"], np.triu(np.ones((2, 2)), 1), 9, pool))
        m = nb["nb_table"]
        results["rays"].append(survey_rays("nb_table", "nb_table", m["scx"], m["scz"], np.ones((2, 2)) - np.eye(2), 9, pool))
        if not args.quick:
            fx, fz = models.weld_array(32, 33, 14)
            results["rays"].append(survey_rays("fmc64", "weld", fx, fz, np.ones((64, 64)) - np.eye(64), 9, pool))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
