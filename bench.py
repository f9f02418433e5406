#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ALI-FMM path (BASELINE.json metric).

Workload ("step"): the weld grid (424 x 500, Weld_rays.py model) at subgrid 9 with 128
transducers (64 top + 64 bottom, x = 27 + 7k): 128 receiver travel-time fields of
3808 x 4492 nodes + all 8192 top<->bottom rays.  One step = one full pass.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU path (oracle port) on the host cores

Multi-GPU: sources are independent (the reference shards them over processes too), so every
rank solves its own 128-source batch on its own GPU with no collective on the data path
("weak" scaling); value = all ranks' node-solves / max-over-ranks device time.
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

SG = 9
B_ALG = 36.0  # algorithmic bytes per node-solve: T store + T load + veln + velpn + vel_map


def workload(n_per_side):
    w = models.weld()
    if n_per_side == 64:
        scx, scz, pairs = models.weld_headline()
    else:  # reduced variants for quick checks (--sources)
        first = 27
        pitch = max(1, (500 - 2 * first) // max(1, n_per_side - 1)) if n_per_side > 1 else 1
        scx, scz = models.weld_array(n_per_side, first, min(pitch, 7 * 64 // n_per_side))
        n = 2 * n_per_side
        pairs = np.zeros((n, n))
        pairs[:n_per_side, n_per_side:] = 1
        pairs[n_per_side:, :n_per_side] = 1
    return w, scx, scz, pairs


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
def _cpu_worker(args):
    """One receiver field (+ its rays) with the oracle on one host core."""
    j, with_rays = args
    from oracle import ali_oracle as orc
    w, scx, scz, pairs = _CPU["w"], _CPU["scx"], _CPU["scz"], _CPU["pairs"]
    om = orc.Model(w["veln"], w["velpn"], w["vel_map"], w["stif_den"])
    t0 = time.perf_counter()
    T = orc.travel_finer_grid(om, scx[j], scz[j], w["dnx"], SG)
    t1 = time.perf_counter()
    n_rays = 0
    if with_rays:
        rec = (SG * round(scx[j] / w["dnx"]), SG * round(scz[j] / w["dnx"]))
        for i in np.nonzero(pairs[:, j])[0][:with_rays]:
            src = (SG * round(scx[i] / w["dnx"]), SG * round(scz[i] / w["dnx"]))
            orc.find_ray(om, w["dnx"], src, rec, T, SG)
            n_rays += 1
    t2 = time.perf_counter()
    return T.size, t1 - t0, n_rays, t2 - t1


_CPU = {}


def cpu_sample(n_fields, rays_per_field, cores):
    """Times ``n_fields`` receiver fields (+ rays) of the workload on ``cores`` host cores."""
    import multiprocessing as mp
    w, scx, scz, pairs = workload(64)
    _CPU.update(w=w, scx=scx, scz=scz, pairs=pairs)
    from oracle import ali_oracle as orc
    orc.build()
    js = [int(round(k * 127 / max(1, n_fields - 1))) if n_fields > 1 else 40 for k in range(n_fields)]
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(j, rays_per_field) for j in js])
    else:
        res = [_cpu_worker((j, rays_per_field)) for j in js]
    wall = time.perf_counter() - t0
    nodes = sum(r[0] for r in res)
    rays = sum(r[2] for r in res)
    return {"wall_s": wall, "node_solves": nodes, "rays": rays, "node_solves_per_s": nodes / wall,
            "ttf_core_s": sum(r[1] for r in res), "ray_core_s": sum(r[3] for r in res)}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is a numba
    module that cannot travel to the GPU box) on all host cores, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    use = min(cores, 128)
    per_step = []
    for s in range(args.warmup + args.steps):
        r = cpu_sample(use, 2, use)
        if s >= args.warmup:
            per_step.append(r)
    wall = sum(r["wall_s"] for r in per_step)
    nodes = sum(r["node_solves"] for r in per_step)
    val = nodes / wall
    line = {
        "impl": "reference", "metric": "ttf_node_solves_per_s", "value": val, "unit": "node-solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "weld 424x500 subgrid 9, 128 transducers: receiver fields 3808x4492 + top<->bottom rays",
                   "sample_per_step": "%d receiver fields + 2 rays each, one per host core" % use},
        "cpu_baseline": {"value": val, "unit": "node-solves/s", "cores": use, "kind": "port",
                         "sample": "%d fields of 17.1 M nodes per step, oracle/ali_oracle.c (C port of the reference, "
                                   "bit-identical to it in the build container), one process per core" % use},
        "e2e": {"value": val, "unit": "node-solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from ali_fmm_and_ray_tracing_b200 import _capi
    from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import ALI_FMM, set_devices

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or _capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_side = args.sources // 2
    w, scx, scz, pairs = workload(n_side)
    n_src = len(scx)
    dnx = w["dnx"]
    iz = np.round(scz / dnx).astype(np.int32)
    ix = np.round(scx / dnx).astype(np.int32)
    ray_src, ray_slot = [], []
    for j in range(n_src):
        for i in np.nonzero(pairs[:, j])[0]:
            ray_src.append(i)
            ray_slot.append(j)
    ray_src = np.array(ray_src)
    ray_slot = np.array(ray_slot, dtype=np.int32)
    n_rays = len(ray_src)
    g = np.ones((361, 2))
    g[:, 0] = np.arange(361)

    # ---- resident-input timing (value): model uploaded once, fields stay in HBM
    ctx = _capi.Context(w["veln"], w["velpn"], w["vel_map"], w["stif_den"], True, g, g.copy(), dnx, device=local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if args.delta_frac:
        ctx.set_option("delta_frac", args.delta_frac)
    if args.threads:
        ctx.set_option("threads_per_source", args.threads)

    def step():
        ctx.ttf(iz, ix, SG, fetch=False)
        c1 = ctx.counters()
        ctx.rays(iz[ray_src], ix[ray_src], ray_slot, want_paths=False)
        c2 = ctx.counters()
        return c1, c2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    ms_march, ms_seq, ms_rays, launches = [], [], [], 0
    c1 = c2 = None
    for _ in range(args.steps):
        c1, c2 = step()
        ms_seq.append(c1["ms_seq"])
        ms_march.append(c1["ms_march"])
        ms_rays.append(c2["ms_rays"])
        launches += c1["kernel_launches"] + c2["kernel_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    node_solves_step = int(c1["node_solves"])
    value = world * args.steps * node_solves_step / (ms_total_max * 1e-3)
    rays_per_s = world * args.steps * n_rays / (ms_total_max * 1e-3)

    # ---- end to end through the reference-facing API with host buffers (e2e)
    set_devices([local])
    fm = ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=dnx)
    import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
    shim.tqdm_disable = True
    if args.delta_frac:
        fm.options["delta_frac"] = args.delta_frac
    ctx.close()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    fm.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], subgrid_size=SG, trans_pairs=pairs,
                                  stif_den=w["stif_den"], n_threads=8)   # warm-up (allocations, page faults)
    barrier()
    t0 = time.perf_counter()
    e2e_calls = []
    for _ in range(e2e_steps):
        tc = time.perf_counter()
        times = fm.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], subgrid_size=SG, trans_pairs=pairs,
                                              stif_den=w["stif_den"], n_threads=8)
        e2e_calls.append(time.perf_counter() - tc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * e2e_steps * node_solves_step / e2e_s
    cap = 5 * (w["veln"].shape[0] + w["veln"].shape[1])
    h2d = int(w["veln"].nbytes + w["veln"].size * 4 + w["vel_map"].nbytes + w["stif_den"].nbytes + 2 * g.nbytes
              + n_src * 8 + n_rays * 12)
    # device -> host per step: the used points of every ray (x and y, packed on the device), their
    # lengths / times / flags, and the per-source records
    ray_points = int(fm.last_counters[0]["ray_points"])
    d2h = int(ray_points * 16 + n_rays * 16 + n_src * 200)
    assert (times > 0).sum() == n_rays

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        march_ms = statistics.mean(ms_march)
        achieved = node_solves_step * B_ALG / (march_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "march_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = min(os.cpu_count() or 1, 16)
            r = cpu_sample(cores, 2, cores)
            cpu = {"value": r["node_solves_per_s"], "unit": "node-solves/s", "cores": cores, "kind": "port",
                   "sample": "%d of the 128 receiver fields (17.1 M nodes each) + 2 rays each, oracle/ali_oracle.c, one "
                             "process per core, %.1f s wall; single-core field time %.1f s" % (
                                 cores, r["wall_s"], r["ttf_core_s"] / cores),
                   "rays_per_s_per_core": r["rays"] / r["ray_core_s"] if r["ray_core_s"] > 0 else None}
        line = {
            "metric": "ttf_node_solves_per_s", "value": value, "unit": "node-solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "weld 424x500 (Weld_rays.py model, synthetic constant stif_den) subgrid 9, %d transducers: %d "
                            "receiver fields of 3808x4492 nodes + %d rays per GPU" % (n_src, n_src, n_rays),
                "sources_per_gpu": n_src, "rays_per_gpu": n_rays, "rays_per_s": rays_per_s,
                "l2_policy": "inputs_larger_than_l2 (%.1f GB of fields per step)" % (n_src * node_solves_step / n_src * 10 / 1e9),
                "ms_seq_kernel": statistics.mean(ms_seq), "ms_march_kernel": march_ms, "ms_rays_kernel": statistics.mean(ms_rays),
                "band_rounds_max": c1["band_rounds_max"], "update_evals_per_node_solve": (c1["band_evals"] + c1["seq_evals"]) / node_solves_step,
                "fallback_evals": c1["fallback_evals"], "delta_frac": args.delta_frac or 0.3,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "ali_march_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": node_solves_step * B_ALG,
                         "note": "the march is round-latency bound (one barrier-separated round per 0.3 dnx/vmax of "
                                 "travel time), not bandwidth bound; see DESIGN.md"},
            "e2e": {"value": e2e_value, "unit": "node-solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "ALI_FMM.find_all_TTF_rays_parallel", "steps": e2e_steps, "s_per_step": e2e_s / e2e_steps,
                    "s_per_call": [round(t, 4) for t in e2e_calls]},
            "gpu_launches": launches, "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sources", type=int, default=128, help="transducers per GPU (default: the headline 128)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--delta-frac", type=float, default=0.0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
