"""Model / workload definitions shared by tests, bench.py and the golden generator.

These are the BASELINE.json configurations: the notebook's three
media, the weld example of Weld_rays.py and the synthetic Voronoi-grain grids.
"""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STEEL_MPA = (249000, 133000, 205000, 125000, 7850)   # notebook cell 34 (MPa, kg/m^3)
STEEL_PA = (249e9, 133e9, 205e9, 125e9, 7850)        # notebook cell 20/26 (Pa)


def const_stif(shape, vals=STEEL_MPA):
    s = np.zeros((shape[0], shape[1], 5), dtype=np.int64)
    s[:, :] = vals
    return s


def notebook_gradient(n=201):
    """Notebook cells 6-16: isotropic medium, velocity 3000 + 21 j m/s, two transducers."""
    veln = np.zeros((n, n))
    velpn = np.ones((n, n), dtype=int)
    vel_map = np.zeros((n, n))
    for j in range(n):
        vel_map[:, j] = 3000 + 21 * j
    return dict(veln=veln, velpn=velpn, vel_map=vel_map, stif_den=None, dnx=1e-3,
                scx=np.array([1e-3, 199e-3]), scz=np.array([30e-3, 180e-3]))


def notebook_christoffel(n=201):
    """Notebook cells 34-40: run-time Christoffel, orientation 20 degrees, three transducers."""
    return dict(veln=20 * np.ones((n, n)), velpn=np.zeros((n, n), dtype=int), vel_map=np.ones((n, n)),
                stif_den=const_stif((n, n)), dnx=1e-3,
                scx=np.array([1e-3, 199e-3, 100e-3]), scz=np.array([100e-3, 140e-3, 1e-3]))


def notebook_table(cls, n=201):
    """Notebook cells 26-30: tabulated material from the steel constants (Pa); ``cls`` provides
    generate_group_vel / generate_phase_vel (the reference's or this repo's ALI_FMM)."""
    g = np.ones((361, 2))
    g[:, 0] = np.arange(361)
    p = g.copy()
    g[:, 1] = cls.generate_group_vel(None, *STEEL_PA, False)
    p[:, 1] = cls.generate_phase_vel(None, *STEEL_PA, False)
    return dict(veln=np.zeros((n, n)), velpn=np.ones((n, n), dtype=int), vel_map=np.ones((n, n)), stif_den=None,
                dnx=1e-3, scx=np.array([1e-3, 199e-3]), scz=np.array([100e-3, 140e-3]), group_vel=g, phase_vel=p)


def weld():
    """Weld example (Weld_rays.py:9-13).  weld_stif_den.npy is absent from the reference
    checkout; the notebook's austenitic constants stand in at every node."""
    z = np.load(os.path.join(HERE, "golden", "weld_model.npz"))
    veln = z["veln"].astype(np.float64)
    velpn = z["velpn"].astype(int)
    vel_map = z["vel_map"].astype(np.float64)
    return dict(veln=veln, velpn=velpn, vel_map=vel_map, stif_den=const_stif(veln.shape), dnx=2e-4)


def weld_crop(nz, nx, z0=0, x0=170):
    w = weld()
    sl = (slice(z0, z0 + nz), slice(x0, x0 + nx))
    return dict(veln=w["veln"][sl].copy(), velpn=w["velpn"][sl].copy(), vel_map=w["vel_map"][sl].copy(),
                stif_den=w["stif_den"][sl].copy(), dnx=w["dnx"])


def weld_array(n_per_side, first_x, pitch, dnx=2e-4, nnz=424):
    """Transducer coordinates (metres): n_per_side on the top row (z = 0) followed by
    n_per_side on the bottom row (z = nnz - 1), x = first_x + pitch k."""
    xs = dnx * (first_x + pitch * np.arange(n_per_side))
    scx = np.concatenate([xs, xs])
    scz = np.concatenate([np.zeros(n_per_side), dnx * (nnz - 1) * np.ones(n_per_side)])
    return scx, scz


def weld_rays_py():
    """Weld_rays.py:15-36: 31 + 31 transducers, x = 25 + 15 k, pairs top -> bottom."""
    scx, scz = weld_array(31, 25, 15)
    pairs = np.zeros((62, 62))
    pairs[:31, 31:] = 1
    return scx, scz, pairs


def weld_headline():
    """Headline workload: 64 + 64 transducers at x = 27 + 7 k, all top <-> bottom rays
    (128 receiver fields, 2 * 64 * 64 = 8192 rays)."""
    scx, scz = weld_array(64, 27, 7)
    pairs = np.zeros((128, 128))
    pairs[:64, 64:] = 1
    pairs[64:, :64] = 1
    return scx, scz, pairs


def voronoi(n, n_seeds, seed, dnx=1e-4):
    """Synthetic randomly oriented anisotropic grid (configs 4/5): Voronoi grains with
    orientation ~ U[0, 180), Christoffel steel everywhere."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0, n, size=(n_seeds, 2))
    ori = rng.uniform(0, 180, size=n_seeds)
    from scipy.spatial import cKDTree
    tree = cKDTree(pts)
    veln = np.empty((n, n))
    rows = max(1, (1 << 24) // n)          # nearest seed of every node, a block of rows at a time (bounded memory)
    xx = np.arange(n, dtype=np.float64)
    for z0 in range(0, n, rows):
        z1 = min(n, z0 + rows)
        q = np.empty(((z1 - z0) * n, 2))
        q[:, 0] = np.repeat(np.arange(z0, z1, dtype=np.float64), n)
        q[:, 1] = np.tile(xx, z1 - z0)
        _, idx = tree.query(q, workers=-1)
        veln[z0:z1] = ori[idx].reshape(z1 - z0, n)
    return dict(veln=veln, velpn=np.zeros((n, n), dtype=np.int32), vel_map=np.ones((n, n)), stif_den=const_stif((n, n)),
                dnx=dnx)


def lattice_sources(n, dnx, rows=16, cols=8):
    """Config 4: sources on a rows x cols lattice at (z, x) = (n/32 + n/16 a, n/16 + n/8 b)."""
    zs = [n // (2 * rows) + (n // rows) * a for a in range(rows)]
    xs = [n // (2 * cols) + (n // cols) * b for b in range(cols)]
    scz = np.array([dnx * z for z in zs for _ in xs])
    scx = np.array([dnx * x for _ in zs for x in xs])
    return scx, scz


def rel_err(ref, got):
    """Per-node relative error |got - ref| / |ref| (0 where ref == 0 == got)."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    d = np.abs(got - ref)
    out = np.zeros_like(d)
    nz = ref != 0
    out[nz] = d[nz] / np.abs(ref[nz])
    out[(~nz) & (d != 0)] = np.inf
    return out


def polyline_distance(px, py, qx, qy):
    """Largest distance from the points (px, py) to the polyline (qx, qy)."""
    px, py, qx, qy = (np.asarray(a, dtype=np.float64) for a in (px, py, qx, qy))
    ax, ay = qx[:-1], qy[:-1]
    bx, by = qx[1:], qy[1:]
    dx, dy = bx - ax, by - ay
    l2 = dx * dx + dy * dy
    l2 = np.where(l2 == 0, 1.0, l2)
    worst = 0.0
    for x, y in zip(px, py):
        t = np.clip(((x - ax) * dx + (y - ay) * dy) / l2, 0.0, 1.0)
        d = np.hypot(x - (ax + t * dx), y - (ay + t * dy))
        worst = max(worst, float(d.min()))
    return worst
