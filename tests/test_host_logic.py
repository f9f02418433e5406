"""CPU: the host-side mirror of the reference interface, the C ABI surface and the sharding."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import models

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_library):
    """Every function include/alifmm.h declares is exported by libalifmm.so (no compute call)."""
    hdr = open(os.path.join(ROOT, "include", "alifmm.h")).read()
    declared = set(re.findall(r"\b(alifmm_[a-z_]+)\s*\(", hdr))
    declared.discard("alifmm_ctx")
    assert len(declared) >= 14
    lib = ctypes.CDLL(built_library)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    from ali_fmm_and_ray_tracing_b200 import _capi
    assert set(_capi.EXPORTS) == declared


def test_no_gpu_means_loud_failure(built_library):
    """There is no CPU fallback: without a device the compute entry points raise."""
    from ali_fmm_and_ray_tracing_b200 import _capi
    if _capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    g = np.ones((361, 2))
    with pytest.raises(_capi.AlifmmError) as ei:
        _capi.Context(np.zeros((8, 8)), np.ones((8, 8), dtype=int), np.ones((8, 8)), None, True, g, g, 1e-3)
    assert "no CUDA device" in str(ei.value)
    from Anis_TTF_rays import ALI_FMM
    m = models.notebook_gradient(21)
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"][:1] * 0.1, m["scz"][:1] * 0.1)
    with pytest.raises(_capi.AlifmmError):
        fm.update(m["veln"], m["velpn"], m["vel_map"])
    with pytest.raises(_capi.AlifmmError):   # the two-GPU strip solve needs two devices, and says so
        _capi.ttf_split(np.zeros((128, 16)), np.ones((128, 16), dtype=int), np.ones((128, 16)), None, True, g, g, 1e-3, 10, 5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ali_fmm_and_ray_tracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ali_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_constructor_mirrors_reference_errors():
    from Anis_TTF_rays import ALI_FMM
    m = models.notebook_gradient(11)
    with pytest.raises(TypeError):  # velpn must be integer (ATR:3834-3838)
        ALI_FMM(m["veln"], m["velpn"].astype(float), m["vel_map"], m["scx"], m["scz"])
    with pytest.raises(TypeError):  # stif_den must be int64 (ATR:3821-3822)
        ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"], stif_den=np.zeros((11, 11, 5), dtype=np.int32))
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], np.array([0.0012, 0.0095]), np.array([0.003, 0.0085]))
    assert fm.velocity_dat.shape == (361, 2) and np.all(fm.velocity_dat[:, 0] == np.arange(361))
    assert list(fm.isx) == [1.0, 10.0] and list(fm.isz) == [3.0, 8.0]  # banker's round as the reference
    assert fm.nnx == 11 and fm.nnz == 11 and fm.nsrc == 2 and fm.dnz == fm.dnx == 1e-3
    assert fm.ray_paths_x is None and fm.ray_len is None
    with pytest.raises(ValueError):  # ATR:4573-4574
        fm.find_all_TTF_rays_parallel(m["veln"], m["velpn"], m["vel_map"], n_threads=1)


def test_material_tables_match_reference():
    """generate_group_vel / generate_phase_vel / add_materials (ATR:4112-4256) vs the fixture
    produced by the reference's own class."""
    from Anis_TTF_rays import ALI_FMM
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_fields.npz"))
    g = ALI_FMM.generate_group_vel(None, *models.STEEL_PA, False)
    p = ALI_FMM.generate_phase_vel(None, *models.STEEL_PA, False)
    assert np.array_equal(g, z["nb2_group"][:, 1]) and np.array_equal(p, z["nb2_phase"][:, 1])
    m = models.notebook_gradient(11)
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    fm.add_materials(np.array(models.STEEL_PA))
    assert fm.velocity_dat.shape == (361, 2) and np.array_equal(fm.velocity_dat[:, 1], g)
    fm.add_materials(np.array(models.STEEL_PA), keep_materials=True)
    assert fm.velocity_dat.shape == (361, 3) and np.array_equal(fm.phase_vel[:, 2], p)
    two = np.array([models.STEEL_PA, models.STEEL_PA], dtype=float)
    fm.add_materials(two, keep_materials=True)  # the reference sizes by materials.shape[1] (ATR:4228)
    assert fm.velocity_dat.shape == (361, 3 + 5)


def test_shard_bounds_cover_and_balance():
    from ali_fmm_and_ray_tracing_b200.sharding import shard_bounds, split_list
    for n in (0, 1, 7, 128, 129):
        for w in (1, 2, 3, 8):
            parts = [shard_bounds(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
            assert sum(split_list(list(range(n)), w), []) == list(range(n))


def test_rank_pairs_partition_the_workload():
    """bench.py's strong-scaling split (and the class's per-device split): every receiver and every ray of
    the headline workload belongs to exactly one rank."""
    from ali_fmm_and_ray_tracing_b200.sharding import rank_pairs, receivers_of
    _, _, pairs = models.weld_headline()
    assert receivers_of(pairs) == list(range(128))
    for world in (1, 2, 4, 8, 3):
        parts = [rank_pairs(pairs, world, r) for r in range(world)]
        assert np.array_equal(sum(parts), pairs)
        recs = [receivers_of(p) for p in parts]
        assert sorted(sum(recs, [])) == list(range(128)) and max(map(len, recs)) - min(map(len, recs)) <= 1
        assert all(int(p.sum()) == 64 * len(r) for p, r in zip(parts, recs))      # the rays into a receiver go with it


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    from ali_fmm_and_ray_tracing_b200.sharding import rank_pairs, receivers_of
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    # the product's split of a 13-transducer workload (what bench.py and ALI_FMM do per rank / device): each rank
    # "solves" its receivers (here: records which ones) -- no collective on the data path; only the bench-style
    # reduction of counters / times crosses ranks
    pairs = np.triu(np.ones((13, 13)), 1) + np.tril(np.ones((13, 13)), -1)
    mine = rank_pairs(pairs, world, rank)
    owner = torch.full((13,), -1, dtype=torch.int64)
    owner[receivers_of(mine)] = rank
    gathered = [torch.empty_like(owner) for _ in range(world)]
    dist.all_gather(gathered, owner)
    t = torch.tensor([float(mine.sum())])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([1.0 + rank])
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        merged = torch.stack(gathered).max(dim=0).values
        q.put((merged.tolist(), t.item(), tmax.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_sharding_over_gloo():
    """N > 1 path on CPU: receivers (and the rays into them) are partitioned by the product's sharding
    functions, nothing is exchanged but counters / times."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, total, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert merged == [0] * 7 + [1] * 6 and total == 13.0 * 12 and tmax == 2.0


def test_dense_ray_arrays_behave_like_numpy_zeros():
    """The class allocates ray_paths_x / ray_paths_y (mostly empty, hundreds of MB) from a mapping
    that opts out of transparent huge pages; callers must not notice (Weld_rays.py slices and saves them)."""
    import io
    from ali_fmm_and_ray_tracing_b200.Anis_TTF_rays import _zeros_sparse
    small = _zeros_sparse((3, 3, 10))
    big = _zeros_sparse((40, 40, 6000))           # 76.8 MB: the mapped path
    for a in (small, big):
        assert isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable
        assert not a.any()
    big[1, 2, :5] = np.arange(5)
    view = big[:, :, 0:7]
    buf = io.BytesIO()
    np.save(buf, view)
    buf.seek(0)
    back = np.load(buf)
    assert back.shape == (40, 40, 7) and np.array_equal(back[1, 2, :5], np.arange(5)) and back.sum() == 10
    copy = big.copy()
    del big, view                                  # the mapping goes away with the last reference
    assert copy[1, 2, 4] == 4.0


def test_strip_rows_of_the_domain_decomposition(built_library):
    """alifmm_split_rows (host arithmetic of alifmm_ttf_split): equal shares on multiples of 4 rows, boundaries kept
    away from the source so that its sequential near-source window (hand-over radius 40 + 8 rows) lies in one strip."""
    from ali_fmm_and_ray_tracing_b200 import _capi
    assert _capi.split_rows(16384, 2, 4096) == [0, 8192, 16384]
    assert _capi.split_rows(16384, 8, 100) == [0] + [2048 * k for k in range(1, 8)] + [16384]
    assert _capi.split_rows(768, 2, 384) == [0, 336, 768]            # source on the first row below the boundary: pushed up
    assert _capi.split_rows(768, 2, 383) == [0, 432, 768]            # source just above the boundary: pushed down
    for nz, n, src in ((1536, 4, 385), (1536, 8, 1535), (4096, 3, 0), (640, 4, 320)):
        rows = _capi.split_rows(nz, n, src)
        assert rows[0] == 0 and rows[-1] == nz and len(rows) == n + 1
        assert all(b % 4 == 0 for b in rows[1:-1]) and all(b - a >= 16 for a, b in zip(rows, rows[1:]))
        owner = max(k for k in range(n) if rows[k] <= src)
        assert src - rows[owner] + 1 >= 48 or owner == 0
        assert rows[owner + 1] - src >= 48 or owner == n - 1
    assert _capi.split_rows(768, 2, 10, split_row=601) == [0, 600, 768]
    for bad in ((768, 2, 384, 380), (768, 1, 10, -1), (768, 9, 10, -1), (768, 4, 10, 300), (200, 8, 100, -1)):
        with pytest.raises(_capi.AlifmmError):
            _capi.split_rows(*bad[:3], split_row=bad[3])


def test_bench_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference`` (the CPU arm the driver times next to the GPU arm) on the smallest config: one JSON
    line with the contract's keys, the oracle port as ``cpu_baseline`` and an ``e2e`` object without copies."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "nb", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "ttf_node_solves_per_s" and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["value"] > 1e5 and "workload" in line["config"]
