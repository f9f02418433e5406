#!/usr/bin/env python
"""bench.py -- benchmark of the B200 ALI-FMM path (BASELINE.json metric and configs).

Headline workload ("step"): the weld grid (424 x 500, Weld_rays.py model) at subgrid 9 with 128
transducers (64 top + 64 bottom, x = 27 + 7k): 128 receiver travel-time fields of
3808 x 4492 nodes + all 8192 top<->bottom rays.  One step = one full pass over the workload.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores
    python bench.py --config fmc64|weld1|nb|vor4096|big16384   # the other BASELINE configs

Multi-GPU (default, "strong"): the metric is "128 sources at 1/2/4/8 GPUs" -- the workload's
receivers are SHARDED over the ranks (sharding.rank_pairs: contiguous blocks, the rays into a
receiver go with it), no collective on the data path; value = the whole workload's node-solves /
max-over-ranks device time.  ``--scaling weak`` replicates the workload on every rank instead
(round 1's mode); at N > 1 one weak step is timed as well and reported as config.weak_value.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tests import models  # noqa: E402
from ali_fmm_and_ray_tracing_b200 import sharding  # noqa: E402

B_ALG = 36.0  # algorithmic bytes per node-solve: T store + T load + veln + velpn + vel_map


# ----------------------------------------------------------------------------- workloads
def make_config(name, sources=128):
    """BASELINE.json configs.  kind "rays": receiver fields + rays (find_all_TTF_rays*);
    kind "fields": fields only (update*)."""
    if name == "headline":
        w = models.weld()
        n_side = sources // 2
        if n_side == 64:
            scx, scz, pairs = models.weld_headline()
        else:  # reduced variants for quick checks (--sources)
            scx, scz = models.weld_array(n_side, 27, max(1, min(7 * 64 // n_side, (500 - 54) // max(1, n_side - 1))))
            pairs = np.zeros((2 * n_side, 2 * n_side))
            pairs[:n_side, n_side:] = 1
            pairs[n_side:, :n_side] = 1
        n = len(scx)
        return dict(name=name, model=w, scx=scx, scz=scz, pairs=pairs, sg=9, kind="rays",
                    workload="weld 424x500 (Weld_rays.py model, synthetic constant stif_den) subgrid 9, %d transducers: %d "
                             "receiver fields of 3808x4492 nodes + %d rays" % (n, n, int(pairs.sum())))
    if name == "weld1":   # config 2: one receiver field (transducer 40 of Weld_rays.py) + the 31 rays from the top row
        w = models.weld()
        scx, scz, _ = models.weld_rays_py()
        pairs = np.zeros((62, 62))
        pairs[:31, 40] = 1
        return dict(name=name, model=w, scx=scx, scz=scz, pairs=pairs, sg=9, kind="rays",
                    workload="weld 424x500 subgrid 9, Weld_rays.py array: receiver 40 field (3808x4492) + 31 rays")
    if name == "fmc64":   # config 3: 64-element full matrix capture
        w = models.weld()
        scx, scz = models.weld_array(32, 33, 14)
        pairs = np.ones((64, 64)) - np.eye(64)
        return dict(name=name, model=w, scx=scx, scz=scz, pairs=pairs, sg=9, kind="rays",
                    workload="weld 424x500 subgrid 9, 64-element FMC (32 top + 32 bottom, x = 33 + 14k): 64 fields of "
                             "3808x4492 nodes + 4032 rays")
    if name == "nb":      # config 1: notebook cells 34-40 (run-time Christoffel, veln = 20), 3 transducers
        m = models.notebook_christoffel()
        pairs = np.triu(np.ones((3, 3)), 1)
        return dict(name=name, model=m, scx=m["scx"], scz=m["scz"], pairs=pairs, sg=9, kind="rays",
                    workload="notebook cells 34-40: 201x201 homogeneous anisotropic medium subgrid 9, 3 transducers: 2 "
                             "receiver fields of 1801x1801 nodes + 3 rays")
    if name == "vor4096":  # config 4
        m = models.voronoi(4096, 4096, 1234)
        scx, scz = models.lattice_sources(4096, m["dnx"])
        return dict(name=name, model=m, scx=scx, scz=scz, pairs=None, sg=1, kind="fields",
                    workload="synthetic 4096x4096 Voronoi-grain Christoffel grid (4096 grains, seed 1234), 128 sources on "
                             "a 16x8 lattice, subgrid 1 (travel)")
    if name == "big16384":  # config 5 on ONE GPU (the field fits; domain decomposition is capacity, not speed)
        m = models.voronoi(16384, 65536, 1235)
        scx = np.array([m["dnx"] * 8192])
        scz = np.array([m["dnx"] * 8192])
        return dict(name=name, model=m, scx=scx, scz=scz, pairs=None, sg=1, kind="fields",
                    workload="synthetic 16384x16384 Voronoi-grain Christoffel grid (65536 grains, seed 1235), one source "
                             "at the centre, subgrid 1 (travel)")
    raise ValueError("unknown config " + name)


def field_nodes(cfg):
    nz, nx = cfg["model"]["veln"].shape
    sg = cfg["sg"]
    return (sg * (nz - 1) + 1) * (sg * (nx - 1) + 1) if sg > 1 else nz * nx


def tables(m):
    if m.get("group_vel") is not None:
        return m["group_vel"], m["phase_vel"]
    g = np.ones((361, 2))
    g[:, 0] = np.arange(361)
    return g, g.copy()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
_CPU = {}


def _cpu_worker(args):
    """One field (+ a few of its rays) with the oracle on one host core."""
    j, with_rays = args
    from oracle import ali_oracle as orc
    cfg = _CPU["cfg"]
    m, scx, scz, sg = cfg["model"], cfg["scx"], cfg["scz"], cfg["sg"]
    stif = m["stif_den"] if m["stif_den"] is not None else np.zeros(m["veln"].shape + (5,), dtype=np.int64)
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], stif, m.get("group_vel"), m.get("phase_vel"))
    t0 = time.perf_counter()
    T = orc.travel_finer_grid(om, scx[j], scz[j], m["dnx"], sg) if sg > 1 else orc.travel(om, scx[j], scz[j], m["dnx"])
    t1 = time.perf_counter()
    n_rays = 0
    if with_rays and cfg["pairs"] is not None:
        rec = (sg * round(scx[j] / m["dnx"]), sg * round(scz[j] / m["dnx"]))
        for i in np.nonzero(cfg["pairs"][:, j])[0][:with_rays]:
            src = (sg * round(scx[i] / m["dnx"]), sg * round(scz[i] / m["dnx"]))
            orc.find_ray(om, m["dnx"], src, rec, T, sg)
            n_rays += 1
    t2 = time.perf_counter()
    return T.size, t1 - t0, n_rays, t2 - t1


def cpu_sample(cfg, cores):
    """Times a bounded sample of the workload's fields (+ 2 rays each) with the oracle, one process per
    core on ``cores`` host cores.  Returns the throughput and a description of the sample."""
    import multiprocessing as mp
    from oracle import ali_oracle as orc
    orc.build()
    note = ""
    if cfg["name"] == "big16384":
        # one 268 M-node field is ~2 minutes on a core: time the model's central 4096 x 4096 crop instead
        m = cfg["model"]
        sl = (slice(6144, 10240), slice(6144, 10240))
        sub = dict(veln=np.ascontiguousarray(m["veln"][sl]), velpn=np.ascontiguousarray(m["velpn"][sl]),
                   vel_map=np.ascontiguousarray(m["vel_map"][sl]), stif_den=np.ascontiguousarray(m["stif_den"][sl]),
                   dnx=m["dnx"])
        cfg = dict(cfg, model=sub, scx=np.array([m["dnx"] * 2048]), scz=np.array([m["dnx"] * 2048]))
        note = " (the model's central 4096x4096 crop, source at its centre; per-node cost of the heap march grows with log N)"
    recs = sharding.receivers_of(cfg["pairs"]) if cfg["pairs"] is not None else list(range(len(cfg["scx"])))
    n_fields = min(cores, len(recs))
    js = [recs[int(round(k * (len(recs) - 1) / max(1, n_fields - 1)))] if n_fields > 1 else recs[len(recs) // 2]
          for k in range(n_fields)]
    _CPU["cfg"] = cfg
    t0 = time.perf_counter()
    if n_fields > 1:
        with mp.get_context("fork").Pool(n_fields) as pool:
            res = pool.map(_cpu_worker, [(j, 2) for j in js])
    else:
        res = [_cpu_worker((js[0], 2))]
    wall = time.perf_counter() - t0
    nodes = sum(r[0] for r in res)
    rays = sum(r[2] for r in res)
    return {"wall_s": wall, "node_solves": nodes, "rays": rays, "node_solves_per_s": nodes / wall, "cores": n_fields,
            "ttf_core_s": sum(r[1] for r in res), "ray_core_s": sum(r[3] for r in res),
            "sample": "%d of the workload's %d fields (%.1f M nodes each) + 2 rays each%s, oracle/ali_oracle.c (C port of the "
                      "reference, bit-identical to it in the build container), one process per core, %d cores" % (
                          n_fields, len(recs), res[0][0] / 1e6, note, n_fields)}


def base_line(cfg, args, world, scaling):
    return {
        "metric": "ttf_node_solves_per_s", "unit": "node-solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "name": cfg["name"]},
    }


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is a numba
    module that cannot travel to the GPU box) on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_config(args.config, args.sources)
    cores = os.cpu_count() or 1
    per_step = []
    for s in range(args.warmup + args.steps):
        r = cpu_sample(cfg, cores)
        if s >= args.warmup:
            per_step.append(r)
    wall = sum(r["wall_s"] for r in per_step)
    nodes = sum(r["node_solves"] for r in per_step)
    val = nodes / wall
    line = base_line(cfg, args, args.gpus, "strong")
    line.update({
        "impl": "reference", "value": val, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "cpu_baseline": {"value": val, "unit": "node-solves/s", "cores": per_step[-1]["cores"], "kind": "port",
                         "sample": per_step[-1]["sample"] + " (per step)"},
        "e2e": {"value": val, "unit": "node-solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ali_fmm_and_ray_tracing_b200 import _capi
    from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import ALI_FMM, set_devices
    import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config == "big16384" and world > 1 and rank != 0:
        # (the decomposed single field is driven by rank 0 alone: no 17 GB model per rank)
        dist.barrier()
        dist.barrier()
        dist.destroy_process_group()
        return
    cfg = make_config(args.config, args.sources)
    m, scx, scz, sg = cfg["model"], cfg["scx"], cfg["scz"], cfg["sg"]
    dnx = m["dnx"]
    n_trans = len(scx)
    N = field_nodes(cfg)
    all_pairs = cfg["pairs"]
    all_fields = sharding.receivers_of(all_pairs) if all_pairs is not None else list(range(n_trans))
    g_tab, p_tab = tables(m)
    iz_all = np.round(scz / dnx).astype(np.int32)
    ix_all = np.round(scx / dnx).astype(np.int32)
    shim.tqdm_disable = True
    set_devices([local])
    if cfg["name"] == "big16384" and world > 1:
        # BASELINE config 5 as stated: ONE field, domain-decomposed into row strips over the N GPUs (NVLink halo
        # exchange, alifmm_ttf_split).  Rank 0 drives all N devices through the library; the other ranks only keep the
        # launch contract (barrier, exit).
        run_split(args, cfg, world, rank, local, dist, torch)
        return

    def shard(scaling):
        """This rank's receivers and rays: (field transducer ids, pair matrix or source mask)."""
        w_, r_ = (world, rank) if scaling == "strong" else (1, 0)
        if all_pairs is not None:
            mine = sharding.rank_pairs(all_pairs, w_, r_)
            return sharding.receivers_of(mine), mine
        ids = [all_fields[k] for k in sharding.shard_indices(len(all_fields), w_, r_)]
        mask = np.zeros(n_trans, dtype=int)
        mask[ids] = 1
        return ids, mask

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.current_stream()

    def resident_run(scaling, steps, warmup, sample_clocks):
        """value: model uploaded once, inputs resident in HBM, fields stay in HBM, ray times come back."""
        fields, sel = shard(scaling)
        ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g_tab, p_tab, dnx, device=local)
        ctx.set_stream(stream.cuda_stream)
        for k, v in (("delta_frac", args.delta_frac), ("threads_per_source", args.threads), ("cluster_size", args.cluster),
                     ("cluster_threads", args.cluster_threads), ("seq_threads", args.seq_threads)):
            if v:
                ctx.set_option(k, v)
        iz, ix = iz_all[fields], ix_all[fields]
        ray_src, ray_slot = [], []
        if all_pairs is not None:
            for slot, j in enumerate(fields):
                for i in np.nonzero(sel[:, j])[0]:
                    ray_src.append(i)
                    ray_slot.append(slot)
        ray_src = np.array(ray_src, dtype=int)
        ray_slot = np.array(ray_slot, dtype=np.int32)
        batch = max(1, len(fields))
        if len(fields):
            free, _ = ctx.mem_info()
            batch = max(1, min(len(fields), int(free * 0.8 // (N * 17 + (64 << 20)))))

        def step():
            cs = []
            for pos in range(0, len(fields), batch):
                ctx.ttf(iz[pos:pos + batch], ix[pos:pos + batch], sg, fetch=False)
                c1 = ctx.counters()
                c2 = None
                sel_r = (ray_slot >= pos) & (ray_slot < pos + batch)
                if sel_r.any():
                    ctx.rays(iz_all[ray_src[sel_r]], ix_all[ray_src[sel_r]], ray_slot[sel_r] - pos, want_paths=False)
                    c2 = ctx.counters()
                cs.append((c1, c2))
            return cs

        for _ in range(warmup):
            step()
        sampler = ClockSampler(local) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        acc = {"ms_seq": [], "ms_march": [], "ms_rays": [], "ms_finalize": [], "launches": 0, "ray_points": 0}
        last = None
        for _ in range(steps):
            cs = step()
            for key in ("ms_seq", "ms_march", "ms_finalize"):
                acc[key].append(sum(c1[key] for c1, _ in cs))
            acc["ms_rays"].append(sum(c2["ms_rays"] for _, c2 in cs if c2))
            acc["launches"] += sum(c1["kernel_launches"] + (c2["kernel_launches"] if c2 else 0) for c1, c2 in cs)
            last = cs
        ev1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = allmax(ev0.elapsed_time(ev1))
        ctx.close()
        return {"ms_total": ms, "acc": acc, "last": last, "clocks": clocks, "fields": len(fields), "rays": len(ray_src),
                "batches": (len(fields) + batch - 1) // batch if len(fields) else 0}

    def e2e_run(scaling, steps):
        """e2e: the reference-facing API with HOST buffers (model upload, fields / ray paths back)."""
        fields, sel = shard(scaling)
        fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], scx, scz, group_vel=m.get("group_vel"), phase_vel=m.get("phase_vel"),
                     stif_den=m["stif_den"], dnx=dnx)
        for k, v in (("delta_frac", args.delta_frac), ("cluster_size", args.cluster), ("seq_threads", args.seq_threads)):
            if v:
                fm.options[k] = v

        def call():
            if not fields:
                return None
            if cfg["kind"] == "rays":
                return fm.find_all_TTF_rays_parallel(m["veln"], m["velpn"], m["vel_map"], subgrid_size=sg, trans_pairs=sel,
                                                     stif_den=m["stif_den"], n_threads=8)
            return fm.update_parallel(m["veln"], m["velpn"], m["vel_map"], stif_den=m["stif_den"], subgrid_size=sg,
                                      sources=sel, n_threads=8)

        call()   # warm-up (allocations, page faults)
        barrier()
        t0 = time.perf_counter()
        calls = []
        out = None
        for _ in range(steps):
            tc = time.perf_counter()
            out = call()
            calls.append(time.perf_counter() - tc)
        torch.cuda.synchronize()
        secs = allmax(time.perf_counter() - t0)
        h2d = int(m["veln"].nbytes + m["veln"].size * 4 + m["vel_map"].nbytes + (m["stif_den"].nbytes if m["stif_den"] is not None else 0)
                  + 2 * g_tab.nbytes + len(fields) * 8)
        if cfg["kind"] == "rays":
            n_r = int(sel.sum())
            assert out is None or (out > 0).sum() == n_r
            pts = int(fm.last_counters[0]["ray_points"]) if fields and fm.last_counters[0] else 0
            h2d += n_r * 12
            d2h = int(pts * 16 + n_r * 16 + len(fields) * 200)
        else:
            d2h = int(len(fields) * N * 8)
        return {"s": secs, "calls": calls, "h2d": h2d, "d2h": d2h}

    scaling = args.scaling
    res = resident_run(scaling, args.steps, args.warmup, True)
    whole_fields = len(all_fields) * (world if scaling == "weak" else 1)
    whole_rays = (int(all_pairs.sum()) if all_pairs is not None else 0) * (world if scaling == "weak" else 1)
    node_solves_step = whole_fields * N
    value = args.steps * node_solves_step / (res["ms_total"] * 1e-3)
    rays_per_s = args.steps * whole_rays / (res["ms_total"] * 1e-3)
    weak = None
    if world > 1 and scaling == "strong" and not args.no_weak:
        wr = resident_run("weak", 1, 1, False)
        weak = world * len(all_fields) * N / (wr["ms_total"] * 1e-3)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e = e2e_run(scaling, e2e_steps)
    e2e_value = e2e_steps * node_solves_step / e2e["s"]

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        acc = res["acc"]
        march_ms = statistics.mean(acc["ms_march"]) if acc["ms_march"] else 0.0
        seq_ms = statistics.mean(acc["ms_seq"]) if acc["ms_seq"] else 0.0
        rays_ms = statistics.mean(acc["ms_rays"]) if acc["ms_rays"] else 0.0
        my_nodes = res["fields"] * N           # node-solves rank 0 produces per step (= per march launch sequence)
        achieved = my_nodes * B_ALG / (march_ms * 1e-3) / 1e9 if march_ms > 0 else 0.0
        achieved_ttf = my_nodes * B_ALG / ((march_ms + seq_ms) * 1e-3) / 1e9 if march_ms + seq_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "march_traffic.json")
        if os.path.exists(tp) and cfg["name"] == "headline" and world == 1:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        c1 = res["last"][0][0] if res["last"] else None
        c2 = res["last"][0][1] if res["last"] else None
        line = base_line(cfg, args, world, scaling)
        line.update({"value": value, "ms_per_step": res["ms_total"] / args.steps})
        line["config"].update({
            "fields_total": whole_fields, "rays_total": whole_rays, "fields_rank0": res["fields"], "rays_rank0": res["rays"],
            "batches_per_step_rank0": res["batches"], "rays_per_s": rays_per_s,
            "l2_policy": "inputs_larger_than_l2 (%.2f GB of field state per rank and step)" % (my_nodes * 17 / 1e9),
            "ms_seq_kernel": seq_ms, "ms_march_kernel": march_ms, "ms_rays_kernel": rays_ms,
            "ms_finalize_kernel": statistics.mean(acc["ms_finalize"]) if acc["ms_finalize"] else 0.0,
            "delta_frac": args.delta_frac or 0.35,
        })
        if c1:
            line["config"].update({
                "march_ctas_per_source": c1["cluster_size"], "seq_threads_per_source": c1["seq_threads"],
                "band_rounds_max": c1["band_rounds_max"], "fallback_evals": c1["fallback_evals"],
                "update_evals_per_node_solve": (c1["band_evals"] + c1["seq_evals"]) / max(1, c1["node_solves"]),
                "per_source_spread": {"seq_ms_min": c1["seq_mcycles_min"] / 1.965, "seq_ms_max": c1["seq_mcycles_max"] / 1.965,
                                      "march_ms_min": c1["march_mcycles_min"] / 1.965, "march_ms_max": c1["march_mcycles_max"] / 1.965,
                                      "note": "SM cycles of the fastest / slowest source at 1965 MHz (first batch of the last step)"},
            })
        if c2:
            ray_steps = c2["ray_points"]
            line["config"]["ray_steps_per_s"] = ray_steps / (c2["ms_rays"] * 1e-3) if c2["ms_rays"] > 0 else None
        if weak is not None:
            line["config"]["weak_value"] = weak
            line["config"]["weak_note"] = "every rank solves the whole workload (round 1's mode), one step"
        line["roofline"] = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": "ali_march_cluster_kernel" if c1 and c1["cluster_size"] > 1 else "ali_march_kernel", "peak_source": peak_src,
            "algorithmic_bytes_per_launch": my_nodes * B_ALG,
            "frac_seq_plus_march": achieved_ttf / peak,
            "note": "per rank (rank 0).  The march is round-latency bound (one barrier-separated round per 0.35 dnx/vmax of "
                    "travel time), not bandwidth bound; frac_seq_plus_march counts the sequential near-source kernel too; "
                    "see DESIGN.md"}
        line["e2e"] = {"value": e2e_value, "unit": "node-solves/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "api": "ALI_FMM.find_all_TTF_rays_parallel" if cfg["kind"] == "rays" else "ALI_FMM.update_parallel",
                       "steps": e2e_steps, "s_per_step": e2e["s"] / e2e_steps, "s_per_call": [round(t, 4) for t in e2e["calls"]],
                       "bytes_note": "rank 0's copies"}
        line["gpu_launches"] = res["acc"]["launches"]
        line["clocks"] = res["clocks"]
        if world == 1 and not args.no_cpu:
            r = cpu_sample(cfg, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": r["node_solves_per_s"], "unit": "node-solves/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"] + ", %.1f s wall; single-core field time %.1f s" % (
                                        r["wall_s"], r["ttf_core_s"] / r["cores"]),
                                    "rays_per_s_per_core": r["rays"] / r["ray_core_s"] if r["ray_core_s"] > 0 else None}
        if args.parity:
            line["parity"] = parity_check(cfg, args, local)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_split(args, cfg, world, rank, local, dist, torch):
    from ali_fmm_and_ray_tracing_b200 import _capi
    m = cfg["model"]
    dnx = m["dnx"]
    nz, nx = m["veln"].shape
    g_tab, p_tab = tables(m)
    iz = int(round(float(cfg["scz"][0]) / dnx))
    ix = int(round(float(cfg["scx"][0]) / dnx))
    line = None
    if rank == 0:
        devices = tuple(range(world))

        def call():
            t0 = time.perf_counter()
            _, c = _capi.ttf_split(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g_tab, p_tab, dnx, iz, ix, devices=devices)
            return time.perf_counter() - t0, c

        for _ in range(args.warmup):
            call()
        sampler = ClockSampler(local)
        sampler.start()
        walls, dev_ms, seq_ms, march_ms, launches, rounds = [], [], [], [], 0, 0
        for _ in range(args.steps):
            w, c = call()
            walls.append(w)
            seq_ms.append(c["ms_seq"]); march_ms.append(c["ms_march"]); dev_ms.append(c["ms_seq"] + c["ms_march"])
            launches += c["kernel_launches"]
            rounds = c["band_rounds"]
        clocks = sampler.stop()
        nodes = nz * nx
        ms = statistics.mean(dev_ms)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = (nodes / world) * B_ALG / (statistics.mean(march_ms) * 1e-3) / 1e9
        line = base_line(cfg, args, world, "strong")
        line.update({"value": nodes / (ms * 1e-3), "ms_per_step": ms})
        line["config"].update({
            "decomposition": "%d row strips, one per GPU; halo, claims and per-round minimum over NVLink peer memory (alifmm_ttf_split)" % world,
            "fields_total": 1, "rays_total": 0, "ms_seq_kernel": statistics.mean(seq_ms), "ms_march_kernel": statistics.mean(march_ms),
            "band_rounds_max": rounds, "march_ctas_per_source": 8, "device_bytes_per_gpu": (nodes // world) * 81,
            "l2_policy": "inputs_larger_than_l2 (%.2f GB of field + model state per GPU)" % ((nodes // world) * 81 / 1e9),
            "timing_note": "value: CUDA events around the sequential phase and the strip kernels on the source's device "
                           "(all strips run in lockstep: three inter-GPU barriers per round); e2e: wall time of the call with host arrays "
                           "(model strips up, field back)"})
        line["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                            "kernel": "ali_march_strip_kernel", "peak_source": peak_src, "algorithmic_bytes_per_launch": (nodes / world) * B_ALG,
                            "note": "per GPU.  Round-latency bound: %d rounds of three NVLink barriers each; the decomposition buys capacity, not speed" % rounds}
        h2d = int(m["veln"].nbytes + m["veln"].size * 4 + m["vel_map"].nbytes + m["stif_den"].nbytes + 2 * g_tab.nbytes * world)
        line["e2e"] = {"value": nodes / statistics.mean(walls), "unit": "node-solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": nodes * 8,
                       "api": "alifmm_ttf_split (ALI_FMM.update_i_split)", "steps": args.steps, "s_per_step": statistics.mean(walls),
                       "s_per_call": [round(t, 4) for t in walls]}
        line["gpu_launches"] = launches
        line["clocks"] = clocks
    if world > 1:
        dist.barrier()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_check(cfg, args, device):
    """--parity: compares up to ``args.parity`` of the workload's fields (and 2 rays each) with the oracle."""
    from ali_fmm_and_ray_tracing_b200 import _capi
    from oracle import ali_oracle as orc
    orc.build()
    m, scx, scz, sg = cfg["model"], cfg["scx"], cfg["scz"], cfg["sg"]
    g_tab, p_tab = tables(m)
    recs = sharding.receivers_of(cfg["pairs"]) if cfg["pairs"] is not None else list(range(len(scx)))
    js = [recs[int(round(k * (len(recs) - 1) / max(1, args.parity - 1)))] for k in range(min(args.parity, len(recs)))]
    js = sorted(set(js))
    dnx = m["dnx"]
    iz = np.round(scz / dnx).astype(np.int32)
    ix = np.round(scx / dnx).astype(np.int32)
    ctx = _capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g_tab, p_tab, dnx, device=device)
    stif = m["stif_den"] if m["stif_den"] is not None else np.zeros(m["veln"].shape + (5,), dtype=np.int64)
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], stif, m.get("group_vel"), m.get("phase_vel"))
    ctx.ttf(iz[js], ix[js], sg, fetch=False)
    out = {"fields": [], "rays": []}
    for slot, j in enumerate(js):
        T = ctx.ttf_fetch(slot)
        ref = orc.travel_finer_grid(om, scx[j], scz[j], dnx, sg) if sg > 1 else orc.travel(om, scx[j], scz[j], dnx)
        e = models.rel_err(ref, T)
        out["fields"].append({"transducer": int(j), "frac_gt_1e-5": float((e > 1e-5).mean()), "frac_gt_1e-9": float((e > 1e-9).mean()),
                              "p99": float(np.quantile(e, 0.99)), "max": float(e.max()), "bit_equal": float((ref == T).mean())})
        if cfg["pairs"] is not None:
            srcs = np.nonzero(cfg["pairs"][:, j])[0][:2]
            # rays through the ORACLE's field would test the tracer alone; these go through the GPU field end to end
            x, y, ln, tm, fl = ctx.rays(iz[srcs], ix[srcs], np.full(len(srcs), slot, dtype=np.int32))
            for r, i in enumerate(srcs):
                ox, oy, ot, _ = orc.find_ray(om, dnx, (sg * ix[i], sg * iz[i]), (sg * ix[j], sg * iz[j]), ref, sg)
                dev = models.polyline_distance(x[r, :ln[r]] / sg, y[r, :ln[r]] / sg, ox / sg, oy / sg)
                out["rays"].append({"pair": [int(i), int(j)], "dev_cells": float(dev), "time_rel": float(abs(tm[r] - ot) / ot)})
    ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="headline", choices=["headline", "nb", "weld1", "fmc64", "vor4096", "big16384"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--sources", type=int, default=128, help="headline config: transducers (default 128)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--delta-frac", type=float, default=0.0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--cluster", type=int, default=0, help="CTAs per source in the band march (0 = automatic)")
    ap.add_argument("--cluster-threads", type=int, default=0)
    ap.add_argument("--seq-threads", type=int, default=0)
    ap.add_argument("--parity", type=int, default=0, help="also compare this many fields (+ 2 rays each) with the oracle")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra replicated (weak) step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
