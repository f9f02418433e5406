"""ctypes binding of libalifmm.so (include/alifmm.h).

The library is the only compute path of this package: when it is missing or no CUDA
device is usable every entry point raises -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


class AlifmmError(RuntimeError):
    """Raised when a C-ABI call reports an error (code, message)."""

    def __init__(self, code, msg):
        super().__init__("alifmm error %d: %s" % (code, msg))
        self.code = code


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("nz", ctypes.c_int32), ("nx", ctypes.c_int32), ("dnx", ctypes.c_double),
        ("veln", _f64p), ("velpn", _i32p), ("vel_map", _f64p), ("stif_den", _i64p),
        ("has_stif", ctypes.c_int32), ("group_vel", _f64p), ("phase_vel", _f64p), ("n_cols", ctypes.c_int32),
    ]


class Counters(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in (
        "node_solves", "seq_pops", "seq_evals", "band_rounds", "band_rounds_max", "band_evals", "fallback_evals",
        "max_band", "rays", "ray_points", "kernel_launches")] + [(n, ctypes.c_double) for n in (
            "ms_seq", "ms_march", "ms_finalize", "ms_rays", "vmax", "delta")] + [
        ("cluster_size", ctypes.c_int64), ("seq_threads", ctypes.c_int64)] + [(n, ctypes.c_double) for n in (
            "seq_mcycles_min", "seq_mcycles_max", "march_mcycles_min", "march_mcycles_max")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


EXPORTS = [
    "alifmm_device_count", "alifmm_create", "alifmm_destroy", "alifmm_set_option", "alifmm_set_stream",
    "alifmm_ttf", "alifmm_ttf_fetch", "alifmm_ttf_shape", "alifmm_rays", "alifmm_rays_into", "alifmm_trim",
    "alifmm_mem_info", "alifmm_counters",
    "alifmm_velocity_curves", "alifmm_min_max_vel", "alifmm_last_error",
    "alifmm_velocity_curves_batch", "alifmm_eval_nodes", "alifmm_ttf_split", "alifmm_split_rows",
]

_lib = None


def library_path():
    return _build.LIB_PATH


def load():
    """Loads libalifmm.so; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise AlifmmError(-2, "%s is missing: build it with ali_fmm_and_ray_tracing_b200.build.build_library() "
                              "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    vp = ctypes.c_void_p
    lib.alifmm_device_count.restype = ctypes.c_int
    lib.alifmm_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int, ctypes.POINTER(vp)]
    lib.alifmm_destroy.argtypes = [vp]
    lib.alifmm_destroy.restype = None
    lib.alifmm_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_double]
    lib.alifmm_set_stream.argtypes = [vp, vp]
    lib.alifmm_ttf.argtypes = [vp, ctypes.c_int32, _i32p, _i32p, ctypes.c_int32, _f64p]
    lib.alifmm_ttf_fetch.argtypes = [vp, ctypes.c_int32, _f64p]
    lib.alifmm_ttf_shape.argtypes = [vp, _i32p, _i32p, _i32p, _i32p]
    lib.alifmm_rays.argtypes = [vp, ctypes.c_int32, _i32p, _i32p, _i32p, ctypes.c_int32, _f64p, _f64p, _i32p, _f64p,
                                _i32p]
    lib.alifmm_rays_into.argtypes = [vp, ctypes.c_int32, _i32p, _i32p, _i32p, ctypes.c_int32, ctypes.c_double, _i64p, _f64p,
                                     _f64p, _i32p, _f64p, _i32p]
    lib.alifmm_trim.argtypes = [ctypes.c_int]
    lib.alifmm_mem_info.argtypes = [vp, _i64p, _i64p]
    lib.alifmm_counters.argtypes = [vp, ctypes.POINTER(Counters)]
    lib.alifmm_velocity_curves.argtypes = [vp] + [ctypes.c_double] * 5 + [_f64p, _f64p]
    lib.alifmm_min_max_vel.argtypes = [vp, _f64p, _f64p]
    lib.alifmm_last_error.restype = ctypes.c_char_p
    lib.alifmm_velocity_curves_batch.argtypes = [ctypes.c_int, ctypes.c_int32, _f64p, _f64p, _f64p]
    lib.alifmm_eval_nodes.argtypes = [ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, _f64p,
                                      _i32p, _f64p, _i64p, ctypes.c_int32, _f64p, _f64p, ctypes.c_int32, _f64p, _i32p,
                                      _i32p, _f64p, _f64p, _i32p]
    lib.alifmm_ttf_split.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int32, _i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                     _f64p, ctypes.POINTER(Counters)]
    lib.alifmm_split_rows.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _i32p]
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise AlifmmError(rc, load().alifmm_last_error().decode("utf-8", "replace"))


def trim(device=-1):
    """Releases the device / pinned buffers the library keeps for reuse between contexts."""
    load().alifmm_trim(int(device))


def device_count():
    return load().alifmm_device_count()


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def velocity_curves_batch(materials, device=0):
    """Group and phase velocity curves of a batch of materials on the device: ``materials`` [n, 5] =
    (c22, c23, c33, c44 [Pa], density); returns (group, phase), float64 [n, 361] each."""
    props = np.ascontiguousarray(np.atleast_2d(materials), dtype=np.float64)
    if props.ndim != 2 or props.shape[1] != 5:
        raise ValueError("materials must have shape (n, 5)")
    n = props.shape[0]
    g = np.zeros((n, 361))
    p = np.zeros((n, 361))
    _check(load().alifmm_velocity_curves_batch(int(device), n, _ptr(props, _f64p), _ptr(g, _f64p), _ptr(p, _f64p)))
    return g, p


def split_rows(nz, n_dev, src_iz, split_row=-1):
    """Rows of the strips ``ttf_split`` would use: n_dev + 1 boundaries (alifmm_split_rows; needs no device)."""
    rows = np.zeros(int(n_dev) + 1 if 0 < int(n_dev) < 64 else 65, dtype=np.int32)
    _check(load().alifmm_split_rows(int(nz), int(n_dev), int(src_iz), int(split_row), _ptr(rows, _i32p)))
    return rows[:int(n_dev) + 1].tolist()


def ttf_split(veln, velpn, vel_map, stif_den, has_stif, group_vel, phase_vel, dnx, src_iz, src_ix, devices=(0, 1), split_row=-1):
    """One coarse field (travel(), subgrid 1) decomposed into row strips on len(devices) = 2 ... 8 GPUs
    (alifmm_ttf_split).  Returns (field float64 [nz, nx], counters dict)."""
    veln = np.ascontiguousarray(veln, dtype=np.float64)
    velpn = np.ascontiguousarray(velpn, dtype=np.int32)
    vel_map = np.ascontiguousarray(vel_map, dtype=np.float64)
    stif = None if stif_den is None else np.ascontiguousarray(stif_den, dtype=np.int64)
    group = np.ascontiguousarray(group_vel, dtype=np.float64)
    phase = np.ascontiguousarray(phase_vel, dtype=np.float64)
    nz, nx = veln.shape
    d = ModelDesc(nz, nx, float(dnx), _ptr(veln, _f64p), _ptr(velpn, _i32p), _ptr(vel_map, _f64p), _ptr(stif, _i64p),
                  int(bool(has_stif)), _ptr(group, _f64p), _ptr(phase, _f64p), group.shape[1])
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    out = np.empty((nz, nx))
    c = Counters()
    _check(load().alifmm_ttf_split(ctypes.byref(d), len(dev), _ptr(dev, _i32p), int(src_iz), int(src_ix), int(split_row),
                                   _ptr(out, _f64p), ctypes.byref(c)))
    return out, c.as_dict()


def eval_nodes(veln, velpn, vel_map, stif_den, has_stif, group_vel, phase_vel, dnx, ttn, nsts, pos, device=0):
    """Device update() / fouds18_A() on n independent states ([n, nz, nx] arrays, pos [n, 2] = (iz, ix)).
    Returns (out_update, out_fouds, stencil_no)."""
    veln = np.ascontiguousarray(veln, dtype=np.float64)
    n, nz, nx = veln.shape
    velpn = np.ascontiguousarray(velpn, dtype=np.int32)
    vel_map = np.ascontiguousarray(vel_map, dtype=np.float64)
    stif = None if stif_den is None else np.ascontiguousarray(stif_den, dtype=np.int64)
    group = np.ascontiguousarray(group_vel, dtype=np.float64)
    phase = np.ascontiguousarray(phase_vel, dtype=np.float64)
    ttn = np.ascontiguousarray(ttn, dtype=np.float64)
    nsts = np.ascontiguousarray(nsts, dtype=np.int32)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    assert velpn.shape == veln.shape == vel_map.shape == ttn.shape == nsts.shape and pos.shape == (n, 2)
    assert group.shape == phase.shape and group.shape[0] == 361
    ou = np.zeros(n)
    of = np.zeros(n)
    os_ = np.zeros(n, dtype=np.int32)
    _check(load().alifmm_eval_nodes(int(device), n, nz, nx, float(dnx), _ptr(veln, _f64p), _ptr(velpn, _i32p),
                                    _ptr(vel_map, _f64p), _ptr(stif, _i64p), int(bool(has_stif)), _ptr(group, _f64p),
                                    _ptr(phase, _f64p), group.shape[1], _ptr(ttn, _f64p), _ptr(nsts, _i32p),
                                    _ptr(pos, _i32p), _ptr(ou, _f64p), _ptr(of, _f64p), _ptr(os_, _i32p)))
    return ou, of, os_


class Context:
    """One model resident on one CUDA device (alifmm_ctx)."""

    def __init__(self, veln, velpn, vel_map, stif_den, has_stif, group_vel, phase_vel, dnx, device=0):
        lib = load()
        self._lib = lib
        # keep the converted host arrays alive for the duration of the upload
        self.veln = np.ascontiguousarray(veln, dtype=np.float64)
        self.velpn = np.ascontiguousarray(velpn, dtype=np.int32)
        self.vel_map = np.ascontiguousarray(vel_map, dtype=np.float64)
        self.stif = None if stif_den is None else np.ascontiguousarray(stif_den, dtype=np.int64)
        self.group = np.ascontiguousarray(group_vel, dtype=np.float64)
        self.phase = np.ascontiguousarray(phase_vel, dtype=np.float64)
        if self.veln.ndim != 2 or self.velpn.shape != self.veln.shape or self.vel_map.shape != self.veln.shape:
            raise ValueError("veln, velpn and vel_map must be 2-D arrays of the same shape")
        if self.stif is not None and self.stif.shape != self.veln.shape + (5,):
            raise ValueError("stif_den must have shape (nz, nx, 5)")
        if self.group.ndim != 2 or self.group.shape[0] != 361 or self.phase.shape != self.group.shape:
            raise ValueError("velocity tables must have shape (361, n_materials + 1)")
        self.nz, self.nx = self.veln.shape
        self.dnx = float(dnx)
        self.device = int(device)
        d = ModelDesc(self.nz, self.nx, self.dnx, _ptr(self.veln, _f64p), _ptr(self.velpn, _i32p),
                      _ptr(self.vel_map, _f64p), _ptr(self.stif, _i64p), int(bool(has_stif)), _ptr(self.group, _f64p),
                      _ptr(self.phase, _f64p), self.group.shape[1])
        h = ctypes.c_void_p()
        _check(lib.alifmm_create(ctypes.byref(d), self.device, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.alifmm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        _check(self._lib.alifmm_set_option(self._h, name.encode(), float(value)))

    def set_stream(self, cuda_stream):
        _check(self._lib.alifmm_set_stream(self._h, ctypes.c_void_p(cuda_stream) if cuda_stream else None))

    def field_shape(self, subgrid):
        if subgrid > 1:
            return subgrid * (self.nz - 1) + 1, subgrid * (self.nx - 1) + 1
        return self.nz, self.nx

    def ttf(self, src_iz, src_ix, subgrid=1, out=None, fetch=True):
        """Fields of all sources; returns float64 [n_src, fz, fx] (or None when fetch=False)."""
        src_iz = np.ascontiguousarray(src_iz, dtype=np.int32)
        src_ix = np.ascontiguousarray(src_ix, dtype=np.int32)
        n = len(src_iz)
        fz, fx = self.field_shape(int(subgrid))
        if fetch and out is None:
            out = np.empty((n, fz, fx))
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == n * fz * fx
        _check(self._lib.alifmm_ttf(self._h, n, _ptr(src_iz, _i32p), _ptr(src_ix, _i32p), int(subgrid),
                                    _ptr(out, _f64p) if out is not None else None))
        return out

    def ttf_fetch(self, slot, out=None):
        n, fz, fx, _ = self.ttf_shape()
        if out is None:
            out = np.empty((fz, fx))
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != fz * fx:
            raise ValueError("ttf_fetch: destination must be C-contiguous float64 of the field's size")
        _check(self._lib.alifmm_ttf_fetch(self._h, int(slot), _ptr(out, _f64p)))
        return out

    def ttf_shape(self):
        v = [ctypes.c_int32() for _ in range(4)]
        _check(self._lib.alifmm_ttf_shape(self._h, *[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def rays(self, src_iz, src_ix, rec_slot, capacity=None, want_paths=True):
        """Traces rays through the resident fields.  Returns (x, y, length, time, flag); x / y are
        float64 [n_rays, capacity] in fine-grid units (None when want_paths is False)."""
        src_iz = np.ascontiguousarray(src_iz, dtype=np.int32)
        src_ix = np.ascontiguousarray(src_ix, dtype=np.int32)
        rec_slot = np.ascontiguousarray(rec_slot, dtype=np.int32)
        n = len(src_iz)
        if capacity is None:
            capacity = 5 * (self.nz + self.nx)
        x = np.zeros((n, capacity)) if want_paths else None
        y = np.zeros((n, capacity)) if want_paths else None
        ln = np.zeros(n, dtype=np.int32)
        tm = np.zeros(n)
        fl = np.zeros(n, dtype=np.int32)
        _check(self._lib.alifmm_rays(self._h, n, _ptr(src_iz, _i32p), _ptr(src_ix, _i32p), _ptr(rec_slot, _i32p),
                                     int(capacity), _ptr(x, _f64p), _ptr(y, _f64p), _ptr(ln, _i32p), _ptr(tm, _f64p),
                                     _ptr(fl, _i32p)))
        return x, y, ln, tm, fl

    def rays_into(self, src_iz, src_ix, rec_slot, capacity, divisor, rows, base_x, base_y):
        """Traces rays and writes ray r (its points divided by ``divisor``) into row ``rows[r]`` of the
        caller's dense float64 arrays ``base_x`` / ``base_y`` ([..., capacity], C-contiguous).
        Returns (length, time, flag)."""
        src_iz = np.ascontiguousarray(src_iz, dtype=np.int32)
        src_ix = np.ascontiguousarray(src_ix, dtype=np.int32)
        rec_slot = np.ascontiguousarray(rec_slot, dtype=np.int32)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        n = len(src_iz)
        for a in (base_x, base_y):
            if a.dtype != np.float64 or not a.flags.c_contiguous or a.shape[-1] != capacity:
                raise ValueError("rays_into: destination must be C-contiguous float64 with last dimension = capacity")
            if n and int(rows.max()) >= a.size // capacity:
                raise ValueError("rays_into: row index beyond the destination")
        ln = np.zeros(n, dtype=np.int32)
        tm = np.zeros(n)
        fl = np.zeros(n, dtype=np.int32)
        _check(self._lib.alifmm_rays_into(self._h, n, _ptr(src_iz, _i32p), _ptr(src_ix, _i32p), _ptr(rec_slot, _i32p),
                                          int(capacity), float(divisor), _ptr(rows, _i64p), _ptr(base_x, _f64p),
                                          _ptr(base_y, _f64p), _ptr(ln, _i32p), _ptr(tm, _f64p), _ptr(fl, _i32p)))
        return ln, tm, fl

    def mem_info(self):
        f = ctypes.c_int64()
        t = ctypes.c_int64()
        _check(self._lib.alifmm_mem_info(self._h, ctypes.byref(f), ctypes.byref(t)))
        return f.value, t.value

    def counters(self):
        c = Counters()
        _check(self._lib.alifmm_counters(self._h, ctypes.byref(c)))
        return c.as_dict()

    def velocity_curves(self, c22, c23, c33, c44, density):
        g = np.zeros(361)
        p = np.zeros(361)
        _check(self._lib.alifmm_velocity_curves(self._h, c22, c23, c33, c44, density, _ptr(g, _f64p), _ptr(p, _f64p)))
        return g, p

    def min_max_vel(self):
        lo = ctypes.c_double()
        hi = ctypes.c_double()
        _check(self._lib.alifmm_min_max_vel(self._h, ctypes.byref(lo), ctypes.byref(hi)))
        return lo.value, hi.value
