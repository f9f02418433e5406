"""GPU: parity of the CUDA path (through the C ABI) with the oracle and with the reference's
golden outputs.  Tolerances: BASELINE.json's north_star asks for per-node travel times within
1e-5 relative and ray paths within 0.1 grid cell.  The CUDA path reproduces the reference to
rounding (<= 1e-9) except downstream of nodes whose reference value depends on the pop timing
of the reference's heap (DESIGN.md "Parity"); tests on such inputs bound the affected fraction."""
import os

import numpy as np
import pytest

from tests import models

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_NODE = 1e-5     # north_star
TOL_EXACT = 1e-9    # rounding-level agreement
TOL_BITS = 1e-13    # "the same computation": a few ulps at most (expected: every bit equal)
TOL_CELL = 0.1      # north_star, coarse grid cells


@pytest.fixture(scope="module")
def capi(built_library):
    from ali_fmm_and_ray_tracing_b200 import _capi
    if _capi.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device (the product has no CPU fallback)")
    return _capi


@pytest.fixture(scope="module")
def orc():
    from oracle import ali_oracle
    return ali_oracle


def _load(name):
    z = np.load(os.path.join(G, name))
    return {k: z[k] for k in z.files}


def _tables(m):
    if m.get("group_vel") is not None:
        return m["group_vel"], m["phase_vel"]
    g = np.ones((361, 2))
    g[:, 0] = np.arange(361)
    return g, g.copy()


def _ctx(capi, m):
    g, p = _tables(m)
    return capi.Context(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, p, m["dnx"])


def _omodel(orc, m):
    stif = m["stif_den"] if m["stif_den"] is not None else np.zeros(m["veln"].shape + (5,), dtype=np.int64)
    return orc.Model(m["veln"], m["velpn"], m["vel_map"], stif, m.get("group_vel"), m.get("phase_vel"))


def _check_field(ref, got, frac=0.999, worst=2e-3, med=1e-9, what=""):
    """Field parity: at least ``frac`` of the nodes within the north-star tolerance, the median at
    rounding level and the worst node bounded.  A handful of nodes may differ more: where two
    stencils tie to the last ulp the reference's own value flips with the libm in use
    (tests/test_kernel_replay.py::test_replay_last_ulp_sensitivity)."""
    e = models.rel_err(ref, got)
    assert np.isfinite(e).all(), what
    assert (e <= TOL_NODE).mean() >= frac, (what, (e <= TOL_NODE).mean(), e.max())
    assert np.median(e) <= med, (what, np.median(e))
    assert e.max() <= worst, (what, e.max())
    return e


def _explained(orc, m, ref, T, sg, src, what=""):
    """Where the CUDA field differs from the reference's, why?  The device runs the reference's operator on
    the reference's libm bits, so only the ORDER in which nodes see each other can differ -- and the reference's
    order is that of a heap that mis-orders (parent index round(k/2) with banker's rounding, ATR:123; updates
    only sift up although they may raise a value, ATR:141-175).  The oracle can run the reference's algorithm
    with that heap ordering correctly from the hand-over radius on (ali_oracle.set_true_heap_after: the refined
    source levels and the main grid up to the hand-over keep the reference's own heap, which the sequential
    replica reproduces).  The CUDA field must equal THAT field: bit for bit, or -- where the reference's lazy
    re-evaluation trigger (only when a 4-neighbour pops, ATR:2065-2102) leaves a stale last digit -- within 1e-12.
    So every deviation from the shipped reference beyond 1e-12 is the reference's heap mis-ordering."""
    stop_r = (13 if sg == 1 else 5 * sg + (sg - 1) // 2) + 27      # last refined box + handover_margin (alifmm_set_option)
    om = _omodel(orc, m)
    orc.set_true_heap_after(stop_r)
    try:
        fixed = orc.travel(om, m["dnx"] * src[1], m["dnx"] * src[0], m["dnx"]) if sg == 1 else \
            orc.travel_finer_grid(om, m["dnx"] * src[1], m["dnx"] * src[0], m["dnx"], sg)
    finally:
        orc.set_true_heap_after(-1)
    e = models.rel_err(fixed, T)
    d = models.rel_err(ref, T)
    print("%s: differs from the reference on %.4f of the nodes (%.5f beyond 1e-5, max %.1e); from the reference on a "
          "correct heap on %.6f (max %.1e)" % (what, (ref != T).mean(), (d > TOL_NODE).mean(), d.max(), (fixed != T).mean(), e.max()))
    assert e.max() <= 1e-12, (what, e.max(), (fixed != T).mean())
    return {"bit_equal_fixed": float((fixed == T).mean()), "max_fixed": float(e.max()), "deviating": float((ref != T).mean())}


def _nodes(m, scx, scz):
    return (np.round(np.asarray(scz) / m["dnx"]).astype(np.int32), np.round(np.asarray(scx) / m["dnx"]).astype(np.int32))


# ----------------------------------------------------------------------------- fields, travel()
@pytest.mark.parametrize("name", ["gradient", "christoffel", "table"])
def test_notebook_fields_match_oracle_and_golden(capi, orc, name):
    from Anis_TTF_rays import ALI_FMM
    m = {"gradient": models.notebook_gradient, "christoffel": models.notebook_christoffel,
         "table": lambda: models.notebook_table(ALI_FMM)}[name]()
    ctx = _ctx(capi, m)
    iz, ix = _nodes(m, m["scx"], m["scz"])
    T = ctx.ttf(iz, ix, 1)
    om = _omodel(orc, m)
    # The device runs glibc's own sin / cos / atan (csrc/ali_glibcmath.cuh), so stencil ties resolve exactly
    # as in the reference -- also on the homogeneous, axis-aligned tabulated medium, where a last-ulp change of
    # atan moves 40 % of the nodes (tests/test_kernel_replay.py::test_replay_last_ulp_sensitivity_of_symmetric_media;
    # round 1's correctly rounded functions left 6 % of the second field beyond 1e-5).
    # Measured on B200 (gpurun_out/parity_survey.json, PARITY.md): 5 of the 7 fields equal the reference's bit for
    # bit; the Christoffel field of source 1 differs by <= 2.5e-14 on 4 % of its nodes and the tabulated field of
    # source 1 by up to 2.9e-2 on 5.9 % -- both behind ONE node each where the reference's heap popped late
    # (_explained).  The tabulated medium is homogeneous and axis-aligned: stencils tie everywhere and a single
    # different choice moves everything downstream (test_replay_last_ulp_sensitivity_of_symmetric_media).
    exact = {"gradient": (0, 1), "christoffel": (0, 2), "table": (0,)}[name]
    for k in range(len(iz)):
        ref = orc.travel(om, m["scx"][k], m["scz"][k], m["dnx"])
        e = models.rel_err(ref, T[k])
        print("%s field %d: max rel err %.3e, bit-equal nodes %.6f" % (name, k, e.max(), (ref == T[k]).mean()))
        if k in exact:
            assert np.array_equal(ref, T[k]), (name, k, e.max())
        else:
            _explained(orc, m, ref, T[k], 1, (iz[k], ix[k]), what="%s field %d" % (name, k))
            if name == "christoffel":
                assert e.max() <= 1e-13
            else:
                assert (e > TOL_NODE).mean() <= 0.12 and e.max() <= 6e-2   # (2 x measured)
        assert T[k][iz[k], ix[k]] == 0.0
    gold = _load("golden_fields.npz")
    if name == "gradient":   # straight from the reference (notebook cell 12)
        assert np.array_equal(gold["nb1_T0"], T[0])
        assert abs(T[0].sum() - 1.3402942867072842) <= 1e-12
    elif name == "christoffel":
        assert models.rel_err(gold["nb3_sub"], T[:, ::4, ::4]).max() <= 1e-13
        assert np.array_equal(gold["nb3_sub"][0], T[0, ::4, ::4]) and np.array_equal(gold["nb3_T2"], T[2])
    else:
        assert np.array_equal(gold["nb2_sub"][0], T[0, ::4, ::4])
    ctx.close()


def test_weld_coarse_fields_match_reference_golden(capi, orc):
    """Weld model, travel(): edge and interior sources, against the reference's own output."""
    w = models.weld()
    gold = _load("golden_fields.npz")
    ctx = _ctx(capi, w)
    src = gold["weld1_src"]
    T = ctx.ttf(src[:, 1].astype(np.int32), src[:, 0].astype(np.int32), 1)
    for k in range(len(src)):
        # measured: sources 0, 1, 3 bit-identical to the reference; source 2 (interior, z=200, x=250) within 8e-13
        # (6 % of its nodes, behind one reference heap glitch: test_deviations_are_downstream_of_reference_heap_glitches)
        e = models.rel_err(gold["weld1_sub"][k], T[k][::4, ::4])
        print("weld coarse source %d: max rel err %.2e" % (k, e.max()))
        assert e.max() <= 2e-12, (k, e.max())
        assert abs(T[k].sum() - gold["weld1_sum"][k]) <= 1e-12 * gold["weld1_sum"][k]
    c = ctx.counters()
    assert c["node_solves"] == 4 * 424 * 500 and c["band_rounds_max"] > 100 and c["kernel_launches"] == 3
    ctx.close()


def test_edge_and_corner_sources(capi, orc):
    """Sources on / next to every edge: clipped source boxes, triangular edge stencils and the
    reference's wrong-nnz call (ATR:1645)."""
    m = models.notebook_christoffel(101)
    m["veln"] = 35.0 * np.ones((101, 101))
    pts = [(0, 50), (100, 50), (50, 0), (50, 100), (1, 50), (50, 1), (99, 99), (2, 97), (100, 0)]
    ctx = _ctx(capi, m)
    iz = np.array([p[0] for p in pts], dtype=np.int32)
    ix = np.array([p[1] for p in pts], dtype=np.int32)
    T = ctx.ttf(iz, ix, 1)
    om = _omodel(orc, m)
    for k in range(len(pts)):
        ref = orc.travel(om, m["dnx"] * ix[k], m["dnx"] * iz[k], m["dnx"])
        # homogeneous medium, exact stencil ties: identical to the reference to rounding level since the
        # device's atan / sin / cos agree with glibc (with CUDA's libm two of these sources moved by 1e-4)
        assert models.rel_err(ref, T[k]).max() <= 1e-12, pts[k]
    ctx.close()


def test_kernels_equal_host_replay_bit_for_bit(capi, orc):
    """The kernels against the host replay of the same source files run on the same accurate
    sin / cos / tan / atan (csrc/ali_crmath.cuh; fma() is exact on both sides): every bit equal --
    the sequential phase, the band march on the tiled field, the finalize pass.  This separates
    'the GPU implements the algorithm' (here, exact) from 'the algorithm reproduces the reference'
    (the replay tests on glibc, exact away from reference heap glitches)."""
    from tests.emu import emu
    cases = []
    v = models.voronoi(512, 64, 1234)
    sx, sz = models.lattice_sources(512, v["dnx"], rows=2, cols=2)
    cases.append((v, [(int(round(z / v["dnx"])), int(round(x / v["dnx"]))) for x, z in zip(sx, sz)], 1))
    cases.append((models.weld_crop(120, 160), [(0, 40), (119, 100), (60, 80)], 3))
    h = models.notebook_christoffel(101)
    h["veln"] = 35.0 * np.ones((101, 101))
    cases.append((h, [(0, 50), (99, 99), (2, 97), (50, 50)], 1))
    try:
        emu.set_crmath(True)
        for m, srcs, sg in cases:
            ctx = _ctx(capi, m)
            om = _omodel(orc, m)
            T = ctx.ttf(np.array([s[0] for s in srcs], dtype=np.int32), np.array([s[1] for s in srcs], dtype=np.int32), sg)
            for k, s in enumerate(srcs):
                R, _, rc = emu.ttf(om, m["dnx"], s[0], s[1], sg)
                assert rc == 0 and np.array_equal(R, T[k]), (s, sg, models.rel_err(R, T[k]).max())
            ctx.close()
    finally:
        emu.set_crmath(False)


# ----------------------------------------------------------------------------- BASELINE configs 3, 4, 5
def test_voronoi_config4_fields_match_oracle(capi, orc):
    """BASELINE config 4 at a size the oracle finishes in seconds: randomly oriented Voronoi grains
    (Christoffel steel), sources on the lattice, travel() (subgrid 1), one batch."""
    n = 768
    m = models.voronoi(n, n * n // 4096, 1234)
    scx, scz = models.lattice_sources(n, m["dnx"], rows=4, cols=2)
    iz, ix = _nodes(m, scx, scz)
    ctx = _ctx(capi, m)
    T = ctx.ttf(iz, ix, 1)
    om = _omodel(orc, m)
    from tests.emu import emu
    for k in range(len(iz)):
        ref = orc.travel(om, scx[k], scz[k], m["dnx"])
        # 0-2 % of the nodes (depending on the source; measured 0.9800 ... 1.0000 within 1e-5) sit behind a heap glitch of the reference (DESIGN.md 3,
        # tests/test_kernel_replay.py::test_replay_deviations_start_at_reference_glitches) ...
        _check_field(ref, T[k], frac=0.97, med=1e-12, worst=5e-3, what=("voronoi", k))
        _explained(orc, m, ref, T[k], 1, (iz[k], ix[k]), what="voronoi 768 source %d" % k)
        if k in (0, 5):
            # ... and the kernel equals the host replay of its algorithm (on the running libm) bit for bit
            R, _, rc = emu.ttf(om, m["dnx"], int(iz[k]), int(ix[k]), 1)
            assert rc == 0 and np.array_equal(R, T[k]), models.rel_err(R, T[k]).max()
    ctx.close()


def test_voronoi_config4_full_size(capi, orc):
    """BASELINE config 4 at full size (4096 x 4096, 4096 grains): four of the 128 lattice sources in
    one batch; one field against the oracle, the others through size-independent properties (a
    source solved alone gives the same bits; the field is a causal fixed point of the operator)."""
    n = 4096
    m = models.voronoi(n, 4096, 1234)
    scx, scz = models.lattice_sources(n, m["dnx"])
    sel = [0, 37, 90, 127]
    iz, ix = _nodes(m, scx[sel], scz[sel])
    ctx = _ctx(capi, m)
    T = ctx.ttf(iz, ix, 1)
    assert T.shape == (4, n, n) and np.isfinite(T).all() and (T >= 0).all()
    c = ctx.counters()
    assert c["node_solves"] == 4 * n * n
    om = _omodel(orc, m)
    ref = orc.travel(om, scx[sel[1]], scz[sel[1]], m["dnx"])
    _check_field(ref, T[1], frac=0.99, med=1e-12, worst=5e-3, what="voronoi 4096")
    _explained(orc, m, ref, T[1], 1, (iz[1], ix[1]), what="voronoi 4096")
    alone = ctx.ttf(iz[2:3], ix[2:3], 1)[0]
    assert np.array_equal(alone, T[2])
    rng = np.random.default_rng(5)
    ok = tot = 0
    F = T[3]
    for _ in range(300):
        z, x = int(rng.integers(2, n - 2)), int(rng.integers(2, n - 2))
        if max(abs(z - iz[3]), abs(x - ix[3])) < 48:
            continue
        z0, x0 = z - 2, x - 2                          # the operator only looks at the 5 x 5 window
        win = np.ascontiguousarray(F[z0:z0 + 5, x0:x0 + 5])
        sub = dict(veln=m["veln"][z0:z0 + 5, x0:x0 + 5], velpn=m["velpn"][z0:z0 + 5, x0:x0 + 5],
                   vel_map=m["vel_map"][z0:z0 + 5, x0:x0 + 5], stif_den=m["stif_den"][z0:z0 + 5, x0:x0 + 5])
        osub = _omodel(orc, {k: np.ascontiguousarray(v) for k, v in sub.items()})
        nsts = np.where(win < win[2, 2], 0, -1).astype(np.int32)
        v, _ = orc.update_node(osub, win, nsts, 2, 2, m["dnx"])
        tot += 1
        ok += abs(v - win[2, 2]) <= 1e-9 * win[2, 2]
    assert tot > 200 and ok / tot >= 0.97, (ok, tot)
    ctx.close()


def test_long_grid_config5_proxy(capi, orc):
    """Config 5's extent on one axis (16384 nodes) on a strip the oracle can solve: exercises the
    packed 16-bit coordinates and the band capacity on a long front."""
    nz, nx = 16384, 192
    rng = np.random.default_rng(11)
    veln = np.repeat(np.repeat(rng.uniform(0, 180, (nz // 64, nx // 64)), 64, axis=0), 64, axis=1)
    m = dict(veln=veln, velpn=np.zeros((nz, nx), dtype=int), vel_map=np.ones((nz, nx)),
             stif_den=models.const_stif((nz, nx)), dnx=1e-4)
    ctx = _ctx(capi, m)
    T = ctx.ttf(np.array([8192], dtype=np.int32), np.array([96], dtype=np.int32), 1)[0]
    om = _omodel(orc, m)
    ref = orc.travel(om, m["dnx"] * 96, m["dnx"] * 8192, m["dnx"])
    e = models.rel_err(ref, T)
    # a long channel carries every reference heap glitch to its end: 26 % of the nodes end up
    # 1.1e-5 ... 2e-5 from the reference, none beyond 2.4e-3 (see the replay test named above) ...
    assert (e <= 2e-5).mean() >= 0.9 and (e <= 1e-4).mean() >= 0.9999 and e.max() <= 5e-3 and np.median(e) <= 1e-8
    # ... while the kernel and the host replay of its algorithm (on the running libm) give the same bits, and all of
    # it starts at reference heap glitches
    from tests.emu import emu
    R, _, rc = emu.ttf(om, m["dnx"], 8192, 96, 1)
    assert rc == 0 and np.array_equal(R, T), models.rel_err(R, T).max()
    _explained(orc, m, ref, T, 1, (8192, 96), what="16384 x 192 strip")
    ctx.close()


def test_fmc_config3_all_pairs(capi, orc):
    """BASELINE config 3 in small: full-matrix capture on the weld (subgrid 3) -- every i != j
    pair of 3 top + 3 bottom elements, including same-side pairs (rays along the surface), through
    the class API; times and paths against the oracle's find_ray through the oracle's fields."""
    from Anis_TTF_rays import ALI_FMM
    w = models.weld()
    dnx = w["dnx"]
    xs = [33, 243, 467]
    scx = np.array([dnx * x for x in xs] * 2)
    scz = np.array([0.0] * 3 + [dnx * 423] * 3)
    sg = 3
    fm = ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=dnx)
    pairs = np.ones((6, 6)) - np.eye(6)
    times = fm.find_all_TTF_rays(w["veln"], w["velpn"], w["vel_map"], subgrid_size=sg, trans_pairs=pairs,
                                 stif_den=w["stif_den"])
    assert (times > 0).sum() == 30 and not times.diagonal().any()
    om = _omodel(orc, w)
    good = 0
    for j in range(6):
        ref_T = orc.travel_finer_grid(om, scx[j], scz[j], dnx, sg)
        for i in range(6):
            if i == j:
                continue
            rx, ry, rt, _ = orc.find_ray(om, dnx, (sg * fm.isx[i], sg * fm.isz[i]), (sg * fm.isx[j], sg * fm.isz[j]), ref_T, sg)
            x, y = fm.ray_path(i, j)
            assert (x[0], y[0]) == (fm.isx[i], fm.isz[i]) and (x[-1], y[-1]) == (fm.isx[j], fm.isz[j])
            same_path = models.polyline_distance(x, y, rx / sg, ry / sg) <= TOL_CELL
            # a ray may settle on a neighbouring branch of nearly equal time where the two fields differ
            # by a reference heap glitch (measured: 1 of the 30, 1.4 cells, 2.8e-4 in time)
            assert abs(times[i, j] - rt) <= (1e-6 if same_path else 1e-3) * rt, (i, j, times[i, j], rt)
            good += same_path
    assert good >= 28, good


# ----------------------------------------------------------------------------- fields, travel_finer_grid()
@pytest.mark.parametrize("sg", [3, 5])
def test_weld_crop_fine_fields_match_reference_golden(capi, sg):
    c = models.weld_crop(60, 80)
    gold = _load("golden_fields.npz")
    ctx = _ctx(capi, c)
    src = gold["crop_src"]
    T = ctx.ttf(src[:, 1].astype(np.int32), src[:, 0].astype(np.int32), sg)
    assert T.shape == (4, sg * 59 + 1, sg * 79 + 1)
    for k in range(4):
        e = models.rel_err(gold["crop_sg%d_sub" % sg][k], T[k][::3, ::3])
        assert (e <= TOL_NODE).mean() >= 0.99 and np.median(e) <= 1e-13, (k, e.max())
    ctx.close()


def test_weld_crop_sg9_field_matches_reference_golden(capi):
    c = models.weld_crop(30, 40)
    gold = _load("golden_fields.npz")
    ctx = _ctx(capi, c)
    T = ctx.ttf(np.array([0], dtype=np.int32), np.array([5], dtype=np.int32), 9)[0]
    e = models.rel_err(gold["crop9_T"], T)
    assert (e <= TOL_NODE).mean() >= 0.99 and np.median(e) <= 1e-13
    ctx.close()


def test_weld_sg9_headline_field_against_reference(capi):
    """The headline grid (3808 x 4492 per field): transducer 40 of Weld_rays.py against the
    reference's own field (sub-sampled fixture + BASELINE.md scalars)."""
    w = models.weld()
    gold = _load("golden_fields.npz")
    ctx = _ctx(capi, w)
    T = ctx.ttf(np.array([423], dtype=np.int32), np.array([160], dtype=np.int32), 9)[0]
    assert T.shape == (3808, 4492)
    # measured on B200: all 17,105,536 nodes equal the oracle's bit for bit (PARITY.md); against the reference's
    # own sub-sampled output and checksums (round 1's correctly rounded device math left 0.44 % beyond 1e-5):
    assert np.array_equal(gold["weld9_sub"], T[::16, ::16]), models.rel_err(gold["weld9_sub"], T[::16, ::16]).max()
    assert abs(T.sum() - gold["weld9_stats"][0]) <= 1e-13 * gold["weld9_stats"][0]
    assert T.max() == gold["weld9_stats"][1]
    c = ctx.counters()
    assert c["node_solves"] == 3808 * 4492
    ctx.close()


# ----------------------------------------------------------------------------- invariants at full size
def test_batch_equals_single_and_is_deterministic(capi):
    """Fields do not depend on which other sources share the batch, nor on the run."""
    c = models.weld_crop(60, 80)
    ctx = _ctx(capi, c)
    iz = np.array([0, 59, 30, 0], dtype=np.int32)
    ix = np.array([10, 70, 40, 0], dtype=np.int32)
    A = ctx.ttf(iz, ix, 3)
    B = ctx.ttf(iz, ix, 3)
    assert np.array_equal(A, B)
    for k in range(4):
        S = ctx.ttf(iz[k:k + 1], ix[k:k + 1], 3)
        assert np.array_equal(S[0], A[k])
    for t in (256, 64):          # nor does the CTA size of the sequential near-source kernel (one warp / one candidate per warp)
        ctx.set_option("seq_threads", t)
        assert np.array_equal(ctx.ttf(iz, ix, 3), A)
    ctx.set_option("seq_threads", 32)
    ctx.set_option("cluster_size", 1)
    for t in (256, 512, 1024):   # CTA size does not change results
        ctx.set_option("threads_per_source", t)
        assert np.array_equal(ctx.ttf(iz, ix, 3), A)
    ctx.close()


def test_cluster_march_equals_single_cta(capi):
    """The band march with a thread-block cluster per source (2 / 4 / 8 CTAs, 512 or 768 threads each)
    against the one-CTA kernel: every bit equal, on the weld at subgrid 3 / 9 and on a Voronoi grid
    (closed fronts around interior sources), and the same work counters."""
    cases = [(models.weld_crop(60, 80), [(0, 10), (59, 70), (30, 40), (0, 0)], 3),
             (models.weld_crop(40, 56), [(0, 20), (39, 30)], 9),
             (models.voronoi(640, 100, 1234), [(320, 320), (40, 600), (639, 0)], 1)]
    for m, srcs, sg in cases:
        iz = np.array([p[0] for p in srcs], dtype=np.int32)
        ix = np.array([p[1] for p in srcs], dtype=np.int32)
        ctx = _ctx(capi, m)
        ctx.set_option("cluster_size", 1)
        A = ctx.ttf(iz, ix, sg)
        ca = ctx.counters()
        assert ca["cluster_size"] == 1
        for cs, ct in ((2, 768), (2, 0), (3, 768), (4, 512), (5, 512), (6, 512), (7, 768), (8, 512), (8, 768), (0, 0)):
            ctx.set_option("cluster_size", cs)
            ctx.set_option("cluster_threads", ct)
            B = ctx.ttf(iz, ix, sg)
            cb = ctx.counters()
            assert cb["cluster_size"] == (cs if cs else 8), cb["cluster_size"]
            assert np.array_equal(A, B), (sg, cs, ct, models.rel_err(A, B).max())
            assert cb["band_rounds"] == ca["band_rounds"] and cb["max_band"] == ca["max_band"]
        ctx.close()


def test_field_is_a_causal_fixed_point(capi, orc):
    """Oracle-free invariant: re-evaluating a node of the finished field with
    only its earlier neighbours available reproduces its value (nearly everywhere)."""
    c = models.weld_crop(60, 80)
    ctx = _ctx(capi, c)
    T = ctx.ttf(np.array([0], dtype=np.int32), np.array([40], dtype=np.int32), 1)[0]
    om = _omodel(orc, c)
    rng = np.random.default_rng(3)
    ok = tot = 0
    for _ in range(400):
        z, x = rng.integers(2, 58), rng.integers(2, 78)
        if max(abs(z - 0), abs(x - 40)) < 16:
            continue
        nsts = np.where(T < T[z, x], 0, -1).astype(np.int32)
        v, _ = orc.update_node(om, T, nsts, z, x, c["dnx"])
        tot += 1
        ok += abs(v - T[z, x]) <= 1e-9 * T[z, x]
    assert tot > 200 and ok / tot >= 0.97
    ctx.close()


def test_delta_fraction_does_not_change_the_solution(capi):
    """Acceptance band 0.1 ... 0.35 (default) of dnx/vmax: the same bits; 0.4 (the maximum the option accepts):
    within 1e-10 (measured 5e-12)."""
    for m, src, sg in ((models.weld_crop(60, 80), (0, 40), 3), (models.voronoi(512, 64, 7), (256, 200), 1)):
        ctx = _ctx(capi, m)
        iz, ix = np.array([src[0]], dtype=np.int32), np.array([src[1]], dtype=np.int32)
        out = {}
        for f in (0.1, 0.25, 0.3, 0.35, 0.4):
            ctx.set_option("delta_frac", f)
            out[f] = ctx.ttf(iz, ix, sg)[0]
            assert ctx.counters()["delta"] > 0
        for f in (0.1, 0.25, 0.3):
            assert np.array_equal(out[0.35], out[f]), (f, models.rel_err(out[0.35], out[f]).max())
        assert models.rel_err(out[0.35], out[0.4]).max() <= 1e-10
        ctx.close()


# ----------------------------------------------------------------------------- rays
def test_notebook_rays_match_reference_golden(capi):
    """find_all_TTF_rays on the notebook media (cells 16, 40): times and paths from the reference."""
    from Anis_TTF_rays import ALI_FMM
    gold = _load("golden_rays.npz")
    m = models.notebook_gradient()
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    times = fm.find_all_TTF_rays(m["veln"], m["velpn"], m["vel_map"], subgrid_size=9)
    assert times.shape == (2, 2) and times[1, 0] == 0 and times[0, 0] == 0
    assert abs(times[0, 1] - gold["nb1_times"][0, 1]) <= 1e-6 * gold["nb1_times"][0, 1]
    assert abs(times[0, 1] - 5.08845096e-05) <= 1e-10   # value printed in the notebook
    x, y = fm.ray_path(0, 1)
    assert fm.ray_len[0, 1] == len(x) == 341
    assert (x[0], y[0]) == (1.0, 30.0) and (x[-1], y[-1]) == (199.0, 180.0)
    assert models.polyline_distance(x, y, gold["nb1_ray_x"], gold["nb1_ray_y"]) <= TOL_CELL
    assert np.abs(x - gold["nb1_ray_x"]).max() <= 1e-4
    m = models.notebook_christoffel()
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"], stif_den=m["stif_den"])
    times = fm.find_all_TTF_rays(m["veln"], m["velpn"], m["vel_map"], stif_den=m["stif_den"])
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        assert abs(times[i, j] - gold["nb3_times"][i, j]) <= 1e-6 * gold["nb3_times"][i, j]
        x, y = fm.ray_path(i, j)
        assert len(x) == gold["nb3_len"][i, j]
        assert models.polyline_distance(x, y, gold["nb3_ray_x_%d%d" % (i, j)], gold["nb3_ray_y_%d%d" % (i, j)]) <= TOL_CELL
    assert fm.ray_path(2, 0) == (None, None)


def test_rays_through_same_field_match_oracle(capi, orc):
    """The ray kernel alone: rays traced by the GPU and by the oracle through the SAME field."""
    c = models.weld_crop(60, 80)
    ctx = _ctx(capi, c)
    sg = 9
    rec = (59, 70)
    ctx.ttf(np.array([rec[0]], dtype=np.int32), np.array([rec[1]], dtype=np.int32), sg, fetch=False)
    T = ctx.ttf_fetch(0)
    srcs = [(0, 10), (0, 40), (30, 0), (0, 79), (59, 0), (20, 35)]
    x, y, ln, tm, fl = ctx.rays([s[0] for s in srcs], [s[1] for s in srcs], [0] * len(srcs))
    om = _omodel(orc, c)
    for r, s in enumerate(srcs):
        ox, oy, ot, of = orc.find_ray(om, c["dnx"], (sg * s[1], sg * s[0]), (sg * rec[1], sg * rec[0]), T, sg)
        assert ln[r] == len(ox) and fl[r] == of
        assert np.abs(x[r, :ln[r]] - ox).max() <= 1e-6 and np.abs(y[r, :ln[r]] - oy).max() <= 1e-6
        assert abs(tm[r] - ot) <= 1e-10 * ot
    # the work queue hands the rays out longest-first whatever the caller's order, and the register-budget variants of
    # the kernel (option ray_min_blocks) run the same arithmetic: a permuted, much longer job list gives the same rays
    rng = np.random.default_rng(5)
    pick = rng.integers(0, len(srcs), 700)
    for mb in (4, 5, 6):
        ctx.set_option("ray_min_blocks", mb)
        x2, y2, ln2, tm2, fl2 = ctx.rays([srcs[q][0] for q in pick], [srcs[q][1] for q in pick], [0] * len(pick))
        assert np.array_equal(ln2, ln[pick]) and np.array_equal(tm2, tm[pick]) and np.array_equal(fl2, fl[pick])
        assert np.array_equal(x2, x[pick]) and np.array_equal(y2, y[pick])
    ctx.close()


def test_rays_into_dense_layout_equals_plain_rays(capi):
    """alifmm_rays_into (paths packed on the device, scattered into the reference's dense
    [n, n, 5(nz+nx)] arrays, coordinates divided by subgrid_size) == alifmm_rays / subgrid_size."""
    c = models.weld_crop(60, 80)
    ctx = _ctx(capi, c)
    sg = 3
    ctx.ttf(np.array([59, 0], dtype=np.int32), np.array([70, 5], dtype=np.int32), sg, fetch=False)
    srcs = [(0, 10), (0, 40), (30, 0), (0, 79), (59, 0), (20, 35)]
    siz, six = [s[0] for s in srcs], [s[1] for s in srcs]
    slots = [0, 1, 0, 1, 1, 0]
    cap = 5 * (60 + 80)
    x, y, ln, tm, fl = ctx.rays(siz, six, slots, cap)
    for r in range(len(srcs)):
        assert not x[r, ln[r]:].any() and not y[r, ln[r]:].any()   # zeros behind a path, also from a recycled buffer
    bx, by = np.full((4, 5, cap), -7.0), np.full((4, 5, cap), -7.0)
    rows = np.array([3, 19, 0, 7, 12, 8])
    ln2, tm2, fl2 = ctx.rays_into(siz, six, slots, cap, sg, rows, bx, by)
    assert np.array_equal(ln, ln2) and np.array_equal(tm, tm2) and np.array_equal(fl, fl2)
    fx, fy = bx.reshape(20, cap), by.reshape(20, cap)
    for r, row in enumerate(rows):
        assert np.array_equal(fx[row, :ln[r]], x[r, :ln[r]] / sg) and np.array_equal(fy[row, :ln[r]], y[r, :ln[r]] / sg)
        assert (fx[row, ln[r]:] == -7.0).all()            # the rest of the row is left untouched
    untouched = np.setdiff1d(np.arange(20), rows)
    assert (fx[untouched] == -7.0).all() and (fy[untouched] == -7.0).all()
    with pytest.raises(ValueError):
        ctx.rays_into(siz, six, slots, cap, sg, rows + 20, bx, by)
    ctx.close()
    capi.trim()


def test_weld_sg9_rays_against_reference(capi):
    """Headline-size end to end: field + rays of receiver 40 on the GPU against the reference's
    ray paths and times (Weld_rays.py geometry)."""
    from Anis_TTF_rays import ALI_FMM
    w = models.weld()
    gold = _load("golden_rays.npz")
    srcx = gold["weld9_ray_srcx"]
    scx = w["dnx"] * np.concatenate([srcx, [160]])
    scz = w["dnx"] * np.concatenate([np.zeros(len(srcx)), [423]])
    pairs = np.zeros((5, 5))
    pairs[:4, 4] = 1
    fm = ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
    times = fm.find_all_TTF_rays(w["veln"], w["velpn"], w["vel_map"], trans_pairs=pairs, stif_den=w["stif_den"])
    # the field of receiver 40 equals the reference's bit for bit (test_weld_sg9_headline_field_against_reference),
    # so its rays do too: every point, every time
    for k in range(4):
        x, y = fm.ray_path(k, 4)
        gx, gy = gold["weld9_ray_x_%d" % k], gold["weld9_ray_y_%d" % k]
        assert len(x) == len(gx) and np.abs(x - gx).max() <= 1e-9 and np.abs(y - gy).max() <= 1e-9, k
        assert abs(times[k, 4] - gold["weld9_ray_times"][k]) <= 1e-12 * gold["weld9_ray_times"][k]


def test_weld_rays_py_script_workflow(capi, tmp_path):
    """BASELINE config 2: the reference's own driver script (Weld_rays.py:38-72) on the drop-in class
    -- 31 + 31 transducers, find_all_TTF_rays_parallel at the default subgrid 9 (31 receiver fields of
    17.1 M nodes, 961 rays), the max_len slicing and the four .npy files plot_rays.py reads back."""
    from Anis_TTF_rays import ALI_FMM
    w = models.weld()
    scx, scz, pairs = models.weld_rays_py()
    fm = ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
    trav_times = fm.find_all_TTF_rays_parallel(w["veln"], w["velpn"], w["vel_map"], stif_den=w["stif_den"], n_threads=8,
                                               trans_pairs=pairs)
    assert trav_times.shape == (62, 62) and (trav_times > 0).sum() == 961
    assert not trav_times[31:, :].any() and not trav_times[:, :31].any()
    assert (fm.ray_len > 0).sum() == 961 and fm.ray_paths_x.shape == (62, 62, 5 * (424 + 500))
    max_len = np.max(fm.ray_len)
    assert 800 < max_len < 1200
    rx, ry = fm.ray_paths_x[:, :, 0:max_len], fm.ray_paths_y[:, :, 0:max_len]
    for name, arr in (("trav_times", trav_times), ("ray_paths_x", rx), ("ray_paths_y", ry), ("ray_len", fm.ray_len)):
        np.save(tmp_path / (name + ".npy"), arr)
        assert np.array_equal(np.load(tmp_path / (name + ".npy")), arr)
    # every path starts on its source, ends on its receiver and stays inside the plate
    for i in range(31):
        for j in (31, 40, 61):
            n = fm.ray_len[i, j]
            assert (rx[i, j, 0], ry[i, j, 0]) == (fm.isx[i], fm.isz[i]) and (rx[i, j, n - 1], ry[i, j, n - 1]) == (fm.isx[j], fm.isz[j])
            assert rx[i, j, :n].min() >= 0 and rx[i, j, :n].max() <= 499 and ry[i, j, :n].min() >= 0 and ry[i, j, :n].max() <= 423
            assert not rx[i, j, n:].any()
    # The whole run against the REAL reference's output of the same script (tests/golden/golden_weld_rays.npz,
    # make_golden_weld_rays.py; BASELINE.md section 2 quotes its scalars).  Measured on B200: ray_len.sum() equal,
    # times.sum() to 1e-7, 944 of the 961 paths within 0.1 cell (935 within 0.01) -- the other 17 run through
    # one of the fields that differ from the reference behind a reference heap glitch (PARITY.md) and settle on
    # a neighbouring branch of nearly equal time (worst 4.2 cells, 5e-4 in time).
    g = _load("golden_weld_rays.npz")
    assert abs(g["times"].sum() - 0.01527291403909612) <= 1e-15 and g["ray_len"].sum() == 524461
    assert abs(trav_times.sum() - g["times"].sum()) <= 1e-5 * g["times"].sum()
    assert abs(int(fm.ray_len.sum()) - 524461) <= 524      # within 0.1 %
    for i, j, v in ((0, 31, 1.461554973641925e-05), (30, 31, 2.0367740326699965e-05), (15, 46, 1.5397920119148576e-05)):
        assert abs(trav_times[i, j] - v) <= 1e-4 * v, (i, j, trav_times[i, j])
    assert int(fm.ray_len[fm.ray_len > 0].min()) == 424 and abs(int(fm.ray_len.max()) - 873) <= 2
    rel = np.abs(trav_times - g["times"])[g["times"] > 0] / g["times"][g["times"] > 0]
    assert rel.max() <= 1e-3 and np.median(rel) <= 1e-12, (rel.max(), np.median(rel))
    near = exact = 0
    for k in range(len(g["pair_i"])):
        i, j = int(g["pair_i"][k]), int(g["pair_j"][k])
        gx = g["path_x"][g["offsets"][k]:g["offsets"][k + 1]].astype(np.float64)
        gy = g["path_y"][g["offsets"][k]:g["offsets"][k + 1]].astype(np.float64)
        x, y = fm.ray_path(i, j)
        d = models.polyline_distance(x, y, gx, gy)
        near += d <= TOL_CELL
        exact += d <= 1e-3     # (the fixture stores float32 coordinates)
    print("Weld_rays.py run: %d of 961 paths within 0.1 cell of the reference's, %d within 0.001; times.sum() rel %.2e, worst time rel %.2e" % (
        near, exact, abs(trav_times.sum() - g["times"].sum()) / g["times"].sum(), rel.max()))
    assert near >= 927 and exact >= 900      # (measured 944 / 935; twice the misses allowed)


def test_parallel_methods_over_two_devices_equal_serial(capi):
    """The thread-per-device path of the ``*_parallel`` methods (receivers sharded over the visible GPUs by
    sharding.split_list) on TWO devices: fields, times and paths bitwise equal to the one-device run.  Needs a
    box with at least two GPUs (gpurun --gpus 2); skipped on one."""
    if capi.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from Anis_TTF_rays import ALI_FMM
    import ali_fmm_and_ray_tracing_b200.Anis_TTF_rays as shim
    c = models.weld_crop(60, 80)
    dnx = c["dnx"]
    scx = dnx * np.array([5.0, 30.0, 60.0, 75.0, 20.0, 50.0])
    scz = dnx * np.array([0.0, 0.0, 0.0, 59.0, 59.0, 59.0])
    pairs = np.zeros((6, 6))
    pairs[:3, 3:] = 1
    pairs[3:, :3] = 1
    out = {}
    try:
        for devs in ([0], [0, 1]):
            shim.set_devices(devs)
            fm = ALI_FMM(c["veln"], c["velpn"], c["vel_map"], scx, scz, stif_den=c["stif_den"], dnx=dnx)
            T = fm.update_parallel(c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"], subgrid_size=3)
            times = fm.find_all_TTF_rays_parallel(c["veln"], c["velpn"], c["vel_map"], subgrid_size=3, trans_pairs=pairs,
                                                  stif_den=c["stif_den"], n_threads=2)
            assert len(fm.last_counters) == len(devs) and all(k is not None for k in fm.last_counters)
            out[len(devs)] = (T, times, np.array(fm.ray_paths_x), np.array(fm.ray_paths_y), np.array(fm.ray_len))
    finally:
        shim.set_devices(None)
    for a, b in zip(out[1], out[2]):
        assert np.array_equal(a, b)


def test_two_gpu_strips_equal_one_gpu(capi):
    """one coarse field decomposed into two row strips on two GPUs (2-row halo, claims and the
    per-round minimum exchanged through peer memory) against the same field solved on one GPU: every bit equal.
    Sources in either strip, automatic and explicit split rows, a Voronoi grid (closed fronts crossing the boundary
    both ways) and config 5's long extent.  Needs two GPUs (gpurun --gpus 2); skipped on one."""
    if capi.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    nz, nx = 16384, 192
    rng = np.random.default_rng(11)
    veln = np.repeat(np.repeat(rng.uniform(0, 180, (nz // 64, nx // 64)), 64, axis=0), 64, axis=1)
    long_m = dict(veln=veln, velpn=np.zeros((nz, nx), dtype=int), vel_map=np.ones((nz, nx)),
                  stif_den=models.const_stif((nz, nx)), dnx=1e-4)
    cases = [(models.voronoi(768, 120, 77), [((100, 300), -1), ((700, 40), -1), ((384, 384), -1), ((10, 10), 600), ((767, 767), 64)]),
             (models.weld_crop(200, 260), [((0, 130), -1), ((199, 20), 100)]),
             (long_m, [((8192, 96), -1), ((100, 3), 12000)])]
    for m, runs in cases:
        g, p = _tables(m)
        ctx = _ctx(capi, m)
        for (sz, sx), split in runs:
            one = ctx.ttf(np.array([sz], dtype=np.int32), np.array([sx], dtype=np.int32), 1)[0]
            c1 = ctx.counters()
            try:
                two, c2 = capi.ttf_split(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, p, m["dnx"], sz, sx,
                                         devices=(0, 1), split_row=split)
            except capi.AlifmmError as e:
                if "cannot access each other" in str(e):
                    pytest.skip("the two devices are not peer-accessible")
                raise
            assert np.array_equal(one, two), (m["veln"].shape, sz, sx, split, models.rel_err(one, two).max())
            # same rounds; never more evaluations (the window-change bitmap is folded 512 x 512, and a strip's bitmap
            # only receives its own marks and the boundary's: fewer aliases, so fewer value-preserving re-evaluations)
            assert c2["band_rounds"] == c1["band_rounds"], (c1, c2)
            assert 0.99 * c1["band_evals"] <= c2["band_evals"] <= c1["band_evals"], (c1["band_evals"], c2["band_evals"])
        ctx.close()
    # the class method (devices from set_devices / ALIFMM_DEVICES)
    from Anis_TTF_rays import ALI_FMM
    c = cases[1][0]
    fm = ALI_FMM(c["veln"], c["velpn"], c["vel_map"], c["dnx"] * np.array([130.0]), np.array([0.0]), stif_den=c["stif_den"], dnx=c["dnx"])
    a = fm.update_i(0, c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"])
    b = fm.update_i_split(0, c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"], devices=[0, 1])
    assert np.array_equal(a, b)
    with pytest.raises(RuntimeError):   # the source's refined neighbourhood would straddle the boundary
        m = cases[0][0]
        g, p = _tables(m)
        capi.ttf_split(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, p, m["dnx"], 384, 384, split_row=380)


def test_four_and_eight_gpu_strips_equal_one_gpu(capi):
    """The same decomposition as a chain of 4 (and, where the box has them, 8) strips: halo traffic to the strip above
    and below, the round's scalars to every strip.  Bit-equal to one GPU, same number of rounds."""
    n_gpu = capi.device_count()
    if n_gpu < 4:
        pytest.skip("needs four CUDA devices")
    m = models.voronoi(1536, 300, 78)
    g, p = _tables(m)
    ctx = _ctx(capi, m)
    for n_strips in (4, 8):
        if n_gpu < n_strips:
            continue
        for sz, sx in ((100, 700), (800, 50), (1535, 1535), (385, 768)):   # (the last one sits next to an even boundary)
            one = ctx.ttf(np.array([sz], dtype=np.int32), np.array([sx], dtype=np.int32), 1)[0]
            c1 = ctx.counters()
            try:
                two, c2 = capi.ttf_split(m["veln"], m["velpn"], m["vel_map"], m["stif_den"], True, g, p, m["dnx"], sz, sx,
                                         devices=tuple(range(n_strips)))
            except capi.AlifmmError as e:
                if "cannot access each other" in str(e):
                    pytest.skip("the devices are not peer-accessible")
                raise
            assert np.array_equal(one, two), (n_strips, sz, sx, models.rel_err(one, two).max())
            assert c2["band_rounds"] == c1["band_rounds"], (c1, c2)
    ctx.close()


# ----------------------------------------------------------------------------- reference-facing API
def test_class_api_shapes_and_conventions(capi, tmp_path, monkeypatch):
    from Anis_TTF_rays import ALI_FMM
    c = models.weld_crop(40, 50)
    scx = c["dnx"] * np.array([5.0, 25.0, 45.0])
    scz = c["dnx"] * np.array([0.0, 39.0, 0.0])
    fm = ALI_FMM(c["veln"], c["velpn"], c["vel_map"], scx, scz, stif_den=c["stif_den"], dnx=c["dnx"])
    T = fm.update(c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"], sources=np.array([1, 0, 1]))
    assert T.shape == (3, 40, 50) and not T[1].any() and T[0].any() and T[2].any()
    Ti = fm.update_i(2, c["veln"], c["velpn"], c["vel_map"], c["stif_den"])
    assert np.array_equal(Ti, T[2])
    Tf = fm.update(c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"], subgrid_size=3)
    assert Tf.shape == (3, 118, 148)
    Tp = fm.update_parallel(c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"], subgrid_size=3, n_threads=4)
    assert np.array_equal(Tp, Tf)
    monkeypatch.chdir(tmp_path)
    assert fm.update_parallel(c["veln"], c["velpn"], c["vel_map"], stif_den=c["stif_den"], low_mem=True) is None
    assert np.array_equal(np.load("temp_TTF_2.npy"), T[2])   # ATR:3614
    times = fm.find_all_TTF_rays_parallel(c["veln"], c["velpn"], c["vel_map"], subgrid_size=3, stif_den=c["stif_den"],
                                          n_threads=2)
    serial = fm.find_all_TTF_rays(c["veln"], c["velpn"], c["vel_map"], subgrid_size=3, stif_den=c["stif_den"],
                                  save_rays=False)
    assert np.array_equal(times, serial)
    assert (times > 0).sum() == 3 and times[1, 0] == 0      # default pairs: upper triangle (ATR:4293-4297)
    assert fm.ray_paths_x.shape == (3, 3, 5 * 90)


def test_errors_are_reported_not_swallowed(capi):
    c = models.weld_crop(40, 50)
    ctx = _ctx(capi, c)
    with pytest.raises(capi.AlifmmError) as e:
        ctx.ttf([0], [0], 2)            # even subgrid
    assert e.value.code == -1
    with pytest.raises(capi.AlifmmError):
        ctx.ttf([40], [0], 1)           # source outside the grid
    with pytest.raises(capi.AlifmmError) as e:
        ctx.rays([0], [0], [0])         # no resident field yet
    assert e.value.code == -4
    ctx.set_option("band_capacity_factor", 1.0)
    ctx.ttf([0], [25], 1)               # small grids still fit (capacity has a 1024 floor)
    with pytest.raises(capi.AlifmmError):
        ctx.set_option("delta_frac", 0.6)
    ctx.close()
    g = np.ones((361, 2))
    with pytest.raises(capi.AlifmmError):  # material id outside the tables
        capi.Context(np.zeros((8, 8)), 3 * np.ones((8, 8), dtype=int), np.ones((8, 8)), None, True, g, g, 1e-3)


def test_material_curves_and_model_scan_on_device(capi, orc):
    """alifmm_velocity_curves / alifmm_min_max_vel (ATR:4112-4206, 3736-3787)."""
    from Anis_TTF_rays import ALI_FMM
    w = models.weld()
    ctx = _ctx(capi, w)
    g, p = ctx.velocity_curves(*models.STEEL_PA)
    gr = ALI_FMM.generate_group_vel(None, *models.STEEL_PA, False)
    pr = ALI_FMM.generate_phase_vel(None, *models.STEEL_PA, False)
    assert np.abs(g / gr - 1).max() <= 1e-12 and np.abs(p / pr - 1).max() <= 1e-12
    # weld: velpn[0, 0] == 1, so the reference scans table columns for EVERY node, including the
    # angle column for the Christoffel nodes (ATR:3784-3786): min 0, max 5790
    lo, hi = ctx.min_max_vel()
    assert lo == 0.0 and hi == 5790.0
    ctx.close()
    m = models.notebook_christoffel(41)
    ctx = _ctx(capi, m)
    lo, hi = ctx.min_max_vel()
    vs = [orc.group_vel(a, *models.STEEL_MPA) for a in (0, 45, 90, 135)]
    assert abs(lo - min(vs)) <= 1e-9 * lo and abs(hi - max(vs)) <= 1e-9 * hi
    ctx.close()


def test_cuda_equals_the_reference_algorithm_on_a_correct_heap(capi, orc):
    """The parity statement of this repo (PARITY.md): the CUDA path computes, bit for bit, what the reference's
    own algorithm computes when its narrow-band heap orders correctly beyond the hand-over radius (_explained).
    Models on which the shipped reference deviates -- the notebook's tabulated medium (5.9 % of the nodes beyond
    1e-5), a 16384-long strip (26 %), Voronoi grains, the weld at subgrid 1 and 9 (a headline-size field with 40 %
    of its nodes different from the shipped reference, 3.8 % beyond 1e-5)."""
    from Anis_TTF_rays import ALI_FMM
    import tests.test_kernel_replay as tk
    vor = models.voronoi(768, 144, 1234)
    cases = [("notebook table medium, source 1", models.notebook_table(ALI_FMM), (140, 199), 1),
             ("notebook Christoffel medium, source 1", models.notebook_christoffel(), (140, 199), 1),
             ("notebook gradient medium, subgrid 9", models.notebook_gradient(), (30, 1), 9),
             ("weld coarse, interior source", models.weld(), (200, 250), 1),
             ("voronoi 768", vor, (480, 288), 1), ("voronoi 768", vor, (480, 480), 1), ("voronoi 768", vor, (96, 96), 1),
             ("2048 x 192 strip of 64 x 64 blocks", tk._strip_model(2048, 192, 11), (1024, 96), 1),
             ("weld subgrid 9 (headline grid), transducer at x=300 bottom", models.weld(), (423, 300), 9)]
    deviating = exact = 0
    for what, m, (sz, sx), sg in cases:
        ctx = _ctx(capi, m)
        T = ctx.ttf(np.array([sz], dtype=np.int32), np.array([sx], dtype=np.int32), sg)[0]
        ctx.close()
        om = _omodel(orc, m)
        ref = orc.travel(om, m["dnx"] * sx, m["dnx"] * sz, m["dnx"]) if sg == 1 else orc.travel_finer_grid(om, m["dnx"] * sx, m["dnx"] * sz, m["dnx"], sg)
        r = _explained(orc, m, ref, T, sg, (sz, sx), what=what)
        deviating += r["deviating"] > 0
        exact += r["bit_equal_fixed"] == 1.0
    assert deviating >= 7      # (the check is not vacuous: these fields do differ from the shipped reference)
    assert exact >= 7          # (measured: all but the two stale-last-digit cases, <= 8e-13)


# ----------------------------------------------------------------------------- node-level operators (rows a1-a3, f3)
def test_node_level_update_and_fouds_on_device(capi):
    """Device update() (ATR:904-1410) and fouds18_A() (ATR:240-901) on the 4000 committed states the
    REAL reference evaluated (tests/golden/golden_ops.npz, make_golden.py): same no-solution (-1.0)
    pattern, values equal to the reference's bit for bit."""
    ops = _load("golden_ops.npz")
    upd, fou, sten = capi.eval_nodes(ops["veln"], ops["velpn"], ops["vel_map"], ops["stif"], True, ops["group_tab"],
                                     ops["phase_tab"], float(ops["dnx"]), ops["ttn"], ops["nsts"], ops["pos"])
    ref_u, ref_f = ops["out_update"], ops["out_fouds"]
    none = ref_u == -1.0
    assert none.sum() > 100 and np.array_equal(upd == -1.0, none)           # identical no-stencil cases
    eu = models.rel_err(ref_u[~none], upd[~none])
    ef = models.rel_err(ref_f, fou)
    ulp = 2.0 ** -52
    print("update: max %.2e, >1ulp %d of %d; fouds: max %.2e, >1ulp %d" % (eu.max(), (eu > ulp).sum(), eu.size, ef.max(),
                                                                            (ef > ulp).sum()))
    # measured on B200: every bit equal (the device's atan / sin / cos / tan are glibc's own routines)
    assert eu.max() == 0.0 and ef.max() == 0.0


def test_fouds_fallback_fires_where_the_oracle_fires(capi, orc):
    """Models on which the reference's update() finds no stencil for some evaluations and falls back to
    fouds18_A (counted by the oracle): a two-row strip and a source in the corner of the weld crop
    (coarse and subgrid 3).  The CUDA path must take the fallback too (its count is at least the
    oracle's: forced nodes are re-evaluated every round) and reproduce the field."""
    strip = models.notebook_christoffel(64)
    for k in ("veln", "velpn", "vel_map", "stif_den"):
        strip[k] = np.ascontiguousarray(strip[k][:2])
    crop = models.weld_crop(60, 80)
    for m, (sz, sx), sg in ((strip, (1, 5), 1), (crop, (0, 0), 1), (crop, (0, 0), 3)):
        om = _omodel(orc, m)
        orc.counters(reset=True)
        ref = orc.travel(om, m["dnx"] * sx, m["dnx"] * sz, m["dnx"]) if sg == 1 else \
            orc.travel_finer_grid(om, m["dnx"] * sx, m["dnx"] * sz, m["dnx"], sg)
        _, n_fouds = orc.counters()
        assert n_fouds > 0
        ctx = _ctx(capi, m)
        T = ctx.ttf(np.array([sz], dtype=np.int32), np.array([sx], dtype=np.int32), sg)[0]
        c = ctx.counters()
        ctx.close()
        e = models.rel_err(ref, T)
        print("fallback: oracle %d, device %d, max rel err %.2e" % (n_fouds, c["fallback_evals"], e.max()))
        assert c["fallback_evals"] >= n_fouds
        assert e.max() == 0.0


@pytest.mark.parametrize("shape,sg", [((1024, 40), 1), ((1024, 48), 1), ((160, 14), 3), ((40, 300), 1)])
def test_resort_on_narrow_bands_does_not_change_the_field(capi, shape, sg):
    """Bands of 64 nodes or fewer on a re-sort round (narrow strips): the field must not depend on
    ``resort_every`` and must equal the host replay of the algorithm bit for bit (the sort
    pass used to be skipped for short lists, dropping that round's window-change marks)."""
    from tests.emu import emu
    from oracle import ali_oracle as orc
    nz, nx = shape
    v = models.voronoi(max(nz, nx), 24, 99)
    m = dict(veln=np.ascontiguousarray(v["veln"][:nz, :nx]), velpn=np.zeros((nz, nx), dtype=int),
             vel_map=np.ones((nz, nx)), stif_den=models.const_stif((nz, nx)), dnx=1e-4)
    src = (nz // 2, min(nx // 2, 7))
    fields = []
    for every in (0, 8, 1):
        ctx = _ctx(capi, m)
        ctx.set_option("resort_every", every)
        fields.append(ctx.ttf(np.array([src[0]], dtype=np.int32), np.array([src[1]], dtype=np.int32), sg)[0])
        ctx.close()
    assert np.array_equal(fields[0], fields[1]) and np.array_equal(fields[0], fields[2])
    try:
        emu.set_crmath(True)
        R, _, rc = emu.ttf(_omodel(orc, m), m["dnx"], src[0], src[1], sg)
    finally:
        emu.set_crmath(False)
    assert rc == 0 and np.array_equal(R, fields[1]), models.rel_err(R, fields[1]).max()


def test_non_finite_models_fail_or_finish(capi):
    """NaN in the model is an error (ALIFMM_E_INVALID); nodes whose velocity is NaN for other reasons
    (velpn == 0 without stif_den: Christoffel on zeros, as in the reference) are accepted like any
    other node and the march ends (it used to spin)."""
    m = models.notebook_gradient(48)
    bad = dict(m)
    bad["vel_map"] = m["vel_map"].copy()
    bad["vel_map"][5, 5] = np.nan
    with pytest.raises(capi.AlifmmError) as ei:
        _ctx(capi, bad)
    assert ei.value.code == -1
    hole = dict(m)
    hole["velpn"] = m["velpn"].copy()
    hole["velpn"][20:24, 20:24] = 0      # no stif_den: velocity NaN on these nodes
    ctx = _ctx(capi, hole)
    T = ctx.ttf(np.array([10], dtype=np.int32), np.array([10], dtype=np.int32), 1)[0]
    c = ctx.counters()
    ctx.close()
    assert c["band_rounds_max"] < 100000
    assert np.isnan(T[20:24, 20:24]).any() and np.isfinite(T[:15, :15]).all()


def test_material_curves_batch_on_device(capi):
    """alifmm_velocity_curves_batch (row f4): many materials, one launch; against the host tables
    (ATR:4112-4206 arithmetic on glibc)."""
    from Anis_TTF_rays import ALI_FMM
    rng = np.random.default_rng(3)
    mats = np.array([[249e9, 133e9, 205e9, 125e9, 7850]] + [[rng.uniform(200e9, 300e9), rng.uniform(100e9, 150e9),
                                                             rng.uniform(180e9, 260e9), rng.uniform(90e9, 140e9),
                                                             rng.uniform(7000, 9000)] for _ in range(40)])
    g, p = capi.velocity_curves_batch(mats)
    for k in range(len(mats)):
        gr = ALI_FMM.generate_group_vel(None, *mats[k], False)
        pr = ALI_FMM.generate_phase_vel(None, *mats[k], False)
        assert np.abs(g[k] / gr - 1).max() <= 1e-12 and np.abs(p[k] / pr - 1).max() <= 1e-12
    m = models.notebook_gradient(11)
    fm = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    host = ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    fm.options["tables_on_device"] = True
    fm.add_materials(mats[:7], keep_materials=False)    # 2-D, replaces: the reference fills shape[1] = 5 columns
    host.add_materials(mats[:7], keep_materials=False)
    assert fm.velocity_dat.shape == host.velocity_dat.shape == (361, 6)
    assert np.abs(fm.velocity_dat[:, 1:] / host.velocity_dat[:, 1:] - 1).max() <= 1e-12
    assert np.abs(fm.phase_vel[:, 1:] / host.phase_vel[:, 1:] - 1).max() <= 1e-12
