"""ctypes wrapper around oracle/libali_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

The functions mirror the reference's numba functions (Anis_TTF_rays.py, "ATR"):
``travel`` (ATR:1463), ``travel_finer_grid`` (ATR:2120), ``find_ray`` (ATR:3104),
``time_between_points`` (ATR:2835), ``update`` (ATR:904), ``fouds18_A`` (ATR:240),
``group_vel`` (ATR:3520).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libali_oracle.so")
_lib = None

_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "ali_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libali_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ali_oracle_time_between_points.restype = ctypes.c_double
        _lib.ali_oracle_update_node.restype = ctypes.c_double
        _lib.ali_oracle_fouds_node.restype = ctypes.c_double
        _lib.ali_oracle_group_vel.restype = ctypes.c_double
        _lib.ali_oracle_phase_vel.restype = ctypes.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Model:
    """Coarse model arrays in the dtypes the oracle's C entry points take."""

    def __init__(self, veln, velpn, vel_map=None, stif_den=None, group_vel=None, phase_vel=None):
        self.veln = np.ascontiguousarray(veln, dtype=np.float64)
        self.velpn = np.ascontiguousarray(velpn, dtype=np.int32)
        self.nz, self.nx = self.veln.shape
        if vel_map is None:
            vel_map = np.ones(self.veln.shape)
        self.vel_map = np.ascontiguousarray(vel_map, dtype=np.float64)
        self.has_stif = stif_den is not None
        self.stif = None if stif_den is None else np.ascontiguousarray(stif_den, dtype=np.int64)
        if group_vel is None:
            group_vel = np.ones((361, 2))
            group_vel[:, 0] = np.arange(361)
            phase_vel = group_vel.copy()
        self.group = np.ascontiguousarray(group_vel, dtype=np.float64)
        self.phase = np.ascontiguousarray(phase_vel, dtype=np.float64)
        assert self.group.shape == self.phase.shape and self.group.shape[0] == 361
        self.ncol = self.group.shape[1]

    def _margs(self):
        return (self.nz, self.nx, _p(self.veln, _f64p), _p(self.velpn, _i32p), _p(self.vel_map, _f64p),
                _p(self.stif, _i64p))


def travel(m, scx, scz, dnx, ttn=None):
    """ATR:1463 travel().  ``ttn`` is the caller-owned in/out array (zeros when omitted)."""
    if ttn is None:
        ttn = np.zeros((m.nz, m.nx))
    assert ttn.dtype == np.float64 and ttn.flags.c_contiguous and ttn.shape == (m.nz, m.nx)
    lib().ali_oracle_travel(*m._margs(), int(m.has_stif), _p(m.group, _f64p), _p(m.phase, _f64p), m.ncol,
                            ctypes.c_double(scx), ctypes.c_double(scz), ctypes.c_double(dnx), _p(ttn, _f64p))
    return ttn


def travel_finer_grid(m, scx, scz, dnx, sg):
    """ATR:2120 travel_finer_grid()."""
    out = np.empty((sg * (m.nz - 1) + 1, sg * (m.nx - 1) + 1))
    lib().ali_oracle_travel_finer(*m._margs(), _p(m.group, _f64p), _p(m.phase, _f64p), m.ncol,
                                  ctypes.c_double(scx), ctypes.c_double(scz), ctypes.c_double(dnx), int(sg),
                                  _p(out, _f64p))
    return out


def find_ray(m, dnx, source, receiver, rec_ttf, sg):
    """ATR:3104 find_ray(): returns (ray_x, ray_y, time, flag) in fine-grid coordinates."""
    rec_ttf = np.ascontiguousarray(rec_ttf, dtype=np.float64)
    cap = 5 * (m.nz + m.nx)
    rx = np.zeros(cap)
    ry = np.zeros(cap)
    t = ctypes.c_double(0)
    flag = ctypes.c_int(0)
    n = lib().ali_oracle_find_ray(*m._margs(), int(m.has_stif), _p(m.group, _f64p), m.ncol,
                                  ctypes.c_double(dnx), int(sg), _p(rec_ttf, _f64p), rec_ttf.shape[0],
                                  rec_ttf.shape[1], ctypes.c_double(source[0]), ctypes.c_double(source[1]),
                                  ctypes.c_double(receiver[0]), ctypes.c_double(receiver[1]), _p(rx, _f64p),
                                  _p(ry, _f64p), cap, ctypes.byref(t), ctypes.byref(flag))
    return rx[:n].copy(), ry[:n].copy(), t.value, flag.value


def time_between_points(m, x1, x2, y1, y2, dnx, sg):
    """ATR:2835 time_between_points()."""
    return lib().ali_oracle_time_between_points(*m._margs(), int(m.has_stif), _p(m.group, _f64p), m.ncol,
                                                ctypes.c_double(dnx), int(sg), ctypes.c_double(x1),
                                                ctypes.c_double(x2), ctypes.c_double(y1), ctypes.c_double(y2))


def update_node(m, ttn, nsts, iz, ix, dnx):
    """ATR:904 update() on a caller-provided (ttn, nsts) state.  Returns (value, stencil_no)."""
    ttn = np.ascontiguousarray(ttn, dtype=np.float64)
    nsts = np.ascontiguousarray(nsts, dtype=np.int32)
    st = ctypes.c_int(0)
    v = lib().ali_oracle_update_node(*m._margs(), int(m.has_stif), _p(m.phase, _f64p), m.ncol, _p(ttn, _f64p),
                                     _p(nsts, _i32p), int(iz), int(ix), ctypes.c_double(dnx), ctypes.byref(st))
    return v, st.value


def fouds_node(m, ttn, nsts, iz, ix, dnx):
    """ATR:240 fouds18_A() on a caller-provided (ttn, nsts) state."""
    ttn = np.ascontiguousarray(ttn, dtype=np.float64)
    nsts = np.ascontiguousarray(nsts, dtype=np.int32)
    return lib().ali_oracle_fouds_node(*m._margs(), int(m.has_stif), _p(m.group, _f64p), m.ncol, _p(ttn, _f64p),
                                       _p(nsts, _i32p), int(iz), int(ix), ctypes.c_double(dnx))


def update_node_slab(m, nz, z0, ttn, nsts, iz, ix, dnx):
    """update() (fouds18_A on -1.0) at the ABSOLUTE node (iz, ix) of an nz-row grid of which ``m`` / ``ttn`` /
    ``nsts`` hold only the rows [z0, z0 + m.nz), full width.  Returns (value, used_fouds)."""
    ttn = np.ascontiguousarray(ttn, dtype=np.float64)
    nsts = np.ascontiguousarray(nsts, dtype=np.int32)
    assert ttn.shape == (m.nz, m.nx) == nsts.shape and z0 <= iz - 2 + 2 and iz < z0 + m.nz
    fb = ctypes.c_int(0)
    f = lib().ali_oracle_update_node_slab
    f.restype = ctypes.c_double
    v = f(int(nz), m.nx, int(z0), m.nz, _p(m.veln, _f64p), _p(m.velpn, _i32p), _p(m.vel_map, _f64p), _p(m.stif, _i64p),
          int(m.has_stif), _p(m.phase, _f64p), _p(m.group, _f64p), m.ncol, _p(ttn, _f64p), _p(nsts, _i32p), int(iz), int(ix),
          ctypes.c_double(dnx), ctypes.byref(fb))
    return v, fb.value


def group_vel(angle, c22, c23, c33, c44, sigma, vel_scale=1.0):
    """ATR:3520 group_vel() (stiffness in MPa)."""
    return lib().ali_oracle_group_vel(*[ctypes.c_double(float(v)) for v in (angle, c22, c23, c33, c44, sigma, vel_scale)])


def phase_vel(angle, c22, c23, c33, c44, sigma, vel_scale=1.0):
    """Christoffel phase velocity used inside update() (ATR:1400-1406)."""
    return lib().ali_oracle_phase_vel(*[ctypes.c_double(float(v)) for v in (angle, c22, c23, c33, c44, sigma, vel_scale)])


def set_true_heap(on):
    """Diagnostic: make the narrow band a correct min-heap (see ali_oracle.c, g_true_heap)."""
    lib().ali_oracle_set_true_heap(int(bool(on)))


def set_true_heap_after(stop_r):
    """Diagnostic: the reference's own heap up to the moment a popped main-grid node is ``stop_r`` nodes
    (Chebyshev) from the source, a correct min-heap from then on (see ali_oracle.c); ``stop_r`` < 0 = off."""
    lib().ali_oracle_set_true_heap_after(int(stop_r))


def counters(reset=False):
    u = ctypes.c_long(0)
    f = ctypes.c_long(0)
    lib().ali_oracle_counters(ctypes.byref(u), ctypes.byref(f), int(reset))
    return u.value, f.value
