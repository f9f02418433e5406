"""Generates the committed golden fixtures by running the REAL reference.

Run in the build container only (needs /root/reference, numba):

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/weld_model.npz   the reference's weld example model (weld_veln/velpn/vel_map.npy,
                                  Weld_rays.py:9-11) in compressed form; weld_stif_den.npy is
                                  missing from the reference checkout (.MISSING_LARGE_BLOBS) and
                                  is synthesised at load time (tests/models.py)
    tests/golden/golden_ops.npz   node-level outputs of update / fouds18_A / time_between_points /
                                  group_vel on random states
    tests/golden/golden_fields.npz travel / travel_finer_grid fields (full for small grids,
                                  sub-sampled + checksums for the weld)
    tests/golden/golden_rays.npz  find_ray paths and times, find_all_TTF_rays times (notebook cells
                                  16 / 30 / 40)
Nothing at test or bench time reads /root/reference.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from _refload import load_reference  # noqa: E402
import models  # noqa: E402  (tests/models.py)

REF = "/root/reference"


def md5(a):
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_weld_model():
    veln = np.load(os.path.join(REF, "weld_veln.npy"))
    velpn = np.load(os.path.join(REF, "weld_velpn.npy"))
    vel_map = np.load(os.path.join(REF, "weld_vel_map.npy"))
    np.savez_compressed(os.path.join(HERE, "weld_model.npz"), veln=veln, velpn=velpn.astype(np.int8), vel_map=vel_map)


def make_ops(ref):
    rng = np.random.default_rng(20261018)
    n_cases = 4000
    G = 6  # grid side of each random state
    steel = np.array([249000, 133000, 205000, 125000, 7850], dtype=np.int64)
    gt = np.ones((361, 3))
    gt[:, 0] = np.arange(361)
    gt[:, 2] = ref.ALI_FMM.generate_group_vel(None, 249e9, 133e9, 205e9, 125e9, 7850, False)
    pt = gt.copy()
    pt[:, 2] = ref.ALI_FMM.generate_phase_vel(None, 249e9, 133e9, 205e9, 125e9, 7850, False)
    ttn = np.zeros((n_cases, G, G))
    nsts = np.zeros((n_cases, G, G), dtype=np.int32)
    veln = np.zeros((n_cases, G, G))
    velpn = np.zeros((n_cases, G, G), dtype=np.int64)
    vel_map = np.zeros((n_cases, G, G))
    stif = np.zeros((n_cases, G, G, 5), dtype=np.int64)
    pos = np.zeros((n_cases, 2), dtype=np.int32)
    out_upd = np.zeros(n_cases)
    out_fou = np.zeros(n_cases)
    dnx = 2e-4
    for c in range(n_cases):
        # a plausible local field: plane-ish wave + noise, random availability
        th = rng.uniform(0, 2 * np.pi)
        zz, xx = np.mgrid[0:G, 0:G]
        base = 1e-5 + (np.cos(th) * xx + np.sin(th) * zz) * dnx / 5800.0
        ttn[c] = base * (1 + 0.02 * rng.standard_normal((G, G)))
        if c % 7 == 0:  # exact ties / degenerate stencils
            ttn[c] = np.round(ttn[c] / 2e-8) * 2e-8
        p_avail = rng.choice([0.3, 0.6, 0.9])
        st = np.where(rng.random((G, G)) < p_avail, rng.integers(0, 3, (G, G)), -1)
        nsts[c] = st
        veln[c] = rng.choice([0.0, 20.0, 255.8, 272.2, 352.3, -30.5], size=(G, G))
        kind = c % 3
        if kind == 0:
            velpn[c] = 0
            vel_map[c] = 1.0
        elif kind == 1:
            velpn[c] = 1
            vel_map[c] = rng.uniform(3000, 7000, (G, G))
        else:
            velpn[c] = 2
            vel_map[c] = 1.0
        stif[c] = steel
        iz, ix = rng.integers(0, G, 2)
        pos[c] = (iz, ix)
        st2 = nsts[c].copy()
        st2[iz, ix] = -1 if c % 2 == 0 else 1
        nsts[c] = st2
        if st2[iz, ix] == -1 and c % 4 == 0:
            ttn[c, iz, ix] = 0.0
        out_upd[c] = ref.update(veln[c], velpn[c], vel_map[c], nsts[c], ttn[c], iz, ix, dnx, G, G, pt, stif[c])
        out_fou[c] = ref.fouds18_A(iz, ix, nsts[c], ttn[c], dnx, dnx, G, G, veln[c], velpn[c], vel_map[c], gt, stif[c])
    # time_between_points on a random 12x14 model
    nz, nx = 12, 14
    mv = rng.choice([0.0, 20.0, 255.8, 272.2, 352.3], size=(nz, nx))
    mp = rng.choice([0, 0, 1, 2], size=(nz, nx)).astype(np.int64)
    mm = np.where(mp == 1, rng.uniform(3000, 7000, (nz, nx)), 1.0)
    ms = np.zeros((nz, nx, 5), dtype=np.int64)
    ms[:, :] = steel
    sg = 9
    n_seg = 2000
    seg = np.zeros((n_seg, 4))
    seg_t = np.zeros(n_seg)
    for c in range(n_seg):
        x1 = rng.uniform(0.6, (nx - 1.6)) * sg
        y1 = rng.uniform(0.6, (nz - 1.6)) * sg
        if c % 5 == 0:
            x1, y1 = float(round(x1)), float(round(y1))
        ln = rng.uniform(0.5, 4.0) * sg
        a = rng.uniform(0, 2 * np.pi)
        x2 = float(np.clip(x1 + ln * np.cos(a), 0.6 * sg, (nx - 1.6) * sg))
        y2 = float(np.clip(y1 + ln * np.sin(a), 0.6 * sg, (nz - 1.6) * sg))
        if c % 9 == 0:
            x2, y2 = float(round(x2)), float(round(y2))
        if c % 17 == 0:
            y2 = y1  # horizontal
        seg[c] = (x1, x2, y1, y2)
        seg_t[c] = ref.time_between_points(x1, x2, y1, y2, dnx, sg, gt, mv, mp, mm, ms)
    ang = np.concatenate([np.linspace(0, 180, 721)[:-1], [0.005, 89.995, 90.004, 179.999]])
    gv = np.array([ref.group_vel(a, 249000, 133000, 205000, 125000, 7850, 1.0) for a in ang])
    np.savez_compressed(os.path.join(HERE, "golden_ops.npz"), dnx=dnx, group_tab=gt, phase_tab=pt, ttn=ttn, nsts=nsts,
                        veln=veln, velpn=velpn, vel_map=vel_map, stif=stif, pos=pos, out_update=out_upd,
                        out_fouds=out_fou, tbp_veln=mv, tbp_velpn=mp, tbp_vel_map=mm, tbp_stif=ms, tbp_sg=sg,
                        tbp_seg=seg, tbp_time=seg_t, gv_angle=ang, gv_value=gv)
    print("ops: update -1 count", int((out_upd == -1.0).sum()), "of", n_cases)


def make_fields(ref):
    out = {}
    # notebook config (i): isotropic gradient (cells 6-12)
    m = models.notebook_gradient()
    obj = ref.ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    T = obj.update(m["veln"], m["velpn"], m["vel_map"])
    out["nb1_T0"] = T[0]
    out["nb1_T1_sub"] = T[1][::4, ::4]
    out["nb1_sum"] = np.array([T[0].sum(), T[1].sum(), T[0].max(), T[1].max()])
    # notebook config (iii): run-time Christoffel (cells 34-40)
    m = models.notebook_christoffel()
    obj = ref.ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"], stif_den=m["stif_den"])
    T = obj.update(m["veln"], m["velpn"], m["vel_map"], stif_den=m["stif_den"])
    out["nb3_T2"] = T[2]
    out["nb3_sub"] = T[:, ::4, ::4]
    # notebook config (ii): table material (cells 26-30)
    m = models.notebook_table(ref.ALI_FMM)
    obj = ref.ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    obj.velocity_dat, obj.phase_vel = m["group_vel"], m["phase_vel"]
    T = obj.update(m["veln"], m["velpn"], m["vel_map"])
    out["nb2_sub"] = T[:, ::4, ::4]
    out["nb2_group"] = m["group_vel"]
    out["nb2_phase"] = m["phase_vel"]
    # weld, coarse (Weld_rays.py model), 4 sources incl. edges and interior
    w = models.weld()
    srcs = [(25, 0), (250, 0), (250, 200), (160, 423)]
    scx = np.array([w["dnx"] * s[0] for s in srcs])
    scz = np.array([w["dnx"] * s[1] for s in srcs])
    obj = ref.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], scx, scz, stif_den=w["stif_den"], dnx=w["dnx"])
    T = obj.update(w["veln"], w["velpn"], w["vel_map"], stif_den=w["stif_den"])
    out["weld1_src"] = np.array(srcs)
    out["weld1_sub"] = T[:, ::4, ::4]
    out["weld1_sum"] = T.sum(axis=(1, 2))
    out["weld1_md5"] = np.array([md5(T[k]) for k in range(len(srcs))])
    # weld crop, fine grids
    c = models.weld_crop(60, 80)
    srcs = [(10, 0), (70, 59), (40, 30), (0, 0)]
    scx = np.array([c["dnx"] * s[0] for s in srcs])
    scz = np.array([c["dnx"] * s[1] for s in srcs])
    obj = ref.ALI_FMM(c["veln"], c["velpn"], c["vel_map"], scx, scz, stif_den=c["stif_den"], dnx=c["dnx"])
    out["crop_src"] = np.array(srcs)
    for sg in (3, 5):
        T = np.stack([obj.update_i(k, c["veln"], c["velpn"], c["vel_map"], c["stif_den"], subgrid_size=sg) for k in range(4)])
        out["crop_sg%d_sub" % sg] = T[:, ::3, ::3]
        out["crop_sg%d_sum" % sg] = T.sum(axis=(1, 2))
        out["crop_sg%d_md5" % sg] = np.array([md5(T[k]) for k in range(4)])
    c2 = models.weld_crop(30, 40)
    obj = ref.ALI_FMM(c2["veln"], c2["velpn"], c2["vel_map"], np.array([c2["dnx"] * 5.0]), np.array([0.0]),
                      stif_den=c2["stif_den"], dnx=c2["dnx"])
    T = obj.update_i(0, c2["veln"], c2["velpn"], c2["vel_map"], c2["stif_den"], subgrid_size=9)
    out["crop9_T"] = T
    # weld sg=9, transducer 40 of Weld_rays.py (x=160, z=423): the headline field
    t0 = time.time()
    obj = ref.ALI_FMM(w["veln"], w["velpn"], w["vel_map"], np.array([w["dnx"] * 160]), np.array([w["dnx"] * 423]),
                      stif_den=w["stif_den"], dnx=w["dnx"])
    T = obj.update_i(0, w["veln"], w["velpn"], w["vel_map"], w["stif_den"], subgrid_size=9)
    print("weld sg=9 reference field: %.1f s" % (time.time() - t0))
    out["weld9_sub"] = T[::16, ::16]
    out["weld9_stats"] = np.array([T.sum(), T.max(), T[1904, 2246]])
    out["weld9_md5"] = np.array([md5(T)])
    np.save("/tmp/weld9_rec40.npy", T)  # reused by make_rays in the same run
    np.savez_compressed(os.path.join(HERE, "golden_fields.npz"), **out)
    for k, v in out.items():
        print("fields:", k, v.shape)


def make_rays(ref):
    out = {}
    # notebook cell 16: isotropic gradient, sg=9, ray 0 -> 1
    m = models.notebook_gradient()
    obj = ref.ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    times = obj.find_all_TTF_rays(m["veln"], m["velpn"], m["vel_map"], subgrid_size=9)
    out["nb1_times"] = times
    x, y = obj.ray_path(0, 1)
    out["nb1_ray_x"], out["nb1_ray_y"] = np.array(x), np.array(y)
    # notebook cell 40: Christoffel, 3 transducers
    m = models.notebook_christoffel()
    obj = ref.ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"], stif_den=m["stif_den"])
    times = obj.find_all_TTF_rays(m["veln"], m["velpn"], m["vel_map"], stif_den=m["stif_den"])
    out["nb3_times"] = times
    out["nb3_len"] = obj.ray_len.copy()
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        x, y = obj.ray_path(i, j)
        out["nb3_ray_x_%d%d" % (i, j)], out["nb3_ray_y_%d%d" % (i, j)] = np.array(x), np.array(y)
    # notebook cell 30: table material, both directions
    m = models.notebook_table(ref.ALI_FMM)
    obj = ref.ALI_FMM(m["veln"], m["velpn"], m["vel_map"], m["scx"], m["scz"])
    obj.velocity_dat, obj.phase_vel = m["group_vel"], m["phase_vel"]
    pairs = np.array([[0, 1], [1, 0]])
    out["nb2_times"] = obj.find_all_TTF_rays(m["veln"], m["velpn"], m["vel_map"], trans_pairs=pairs)
    # weld sg=9: rays into transducer 40 from four top transducers (Weld_rays.py geometry)
    w = models.weld()
    T = np.load("/tmp/weld9_rec40.npy")
    rec = np.array([9.0 * 160, 9.0 * 423])
    gv = np.ones((361, 2))
    gv[:, 0] = np.arange(361)
    srcx = [25, 175, 325, 475]
    out["weld9_ray_srcx"] = np.array(srcx)
    tt = []
    for k, sx in enumerate(srcx):
        rx, ry, t = ref.find_ray(w["dnx"], gv, np.array([9.0 * sx, 0.0]), rec, T, w["veln"], w["velpn"], w["vel_map"],
                                 w["stif_den"], 9)
        out["weld9_ray_x_%d" % k], out["weld9_ray_y_%d" % k] = rx / 9, ry / 9
        tt.append(t)
    out["weld9_ray_times"] = np.array(tt)
    np.savez_compressed(os.path.join(HERE, "golden_rays.npz"), **out)
    for k, v in out.items():
        print("rays:", k, v.shape)


if __name__ == "__main__":
    t0 = time.time()
    make_weld_model()
    ref = load_reference(REF)
    make_ops(ref)
    make_fields(ref)
    make_rays(ref)
    print("done in %.0f s" % (time.time() - t0))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
