"""ctypes wrapper for tests/emu/libali_emu.so (host replay of the kernels' logic; test tool only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libali_emu.so")
_lib = None
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libali_emu.so"])
        _lib = ctypes.CDLL(_LIB)
        _lib.emu_model_vmax.restype = ctypes.c_double
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _margs(m, dnx):
    return (m.nz, m.nx, _p(m.veln, _f64p), _p(m.velpn, _i32p), _p(m.vel_map, _f64p), _p(m.stif, _i64p),
            int(m.has_stif), _p(m.group, _f64p), _p(m.phase, _f64p), m.ncol, ctypes.c_double(dnx))


def set_noise(p, seed=0):
    """Last-ulp noise on sin/cos/tan/atan results with probability p (sensitivity studies)."""
    lib().emu_set_noise(ctypes.c_double(p), ctypes.c_ulonglong(seed))


def set_coop(nlanes):
    """Lanes of the cooperative sequential march the replay plays (kernel: 32); 0 = the plain sequential loop."""
    lib().emu_set_coop(int(nlanes))


def set_crmath(on):
    """sin / cos / tan / atan of the replay: False = the running libm (the reference's, default), True = the
    device's functions (csrc/ali_glibcmath.cuh, glibc's own routines restated) -- same bits either way."""
    lib().emu_set_crmath(int(bool(on)))


def math_mismatches(fn, x):
    """Number of arguments in ``x`` for which the device's function (0 atan, 1 sin, 2 cos, 3 tan) and the
    running libm return different bits, and the first such argument."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    first = ctypes.c_double(0.0)
    f = lib().emu_math_mismatches
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_int, _f64p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_double)]
    return int(f(int(fn), _p(x, _f64p), x.size, ctypes.byref(first))), first.value


def set_tiled(on):
    """Replays the band march on the kernel's 4 x 4-tiled field layout (default: row-major)."""
    lib().emu_set_tiled(int(bool(on)))


def model_vmax(m, dnx):
    return lib().emu_model_vmax(*_margs(m, dnx))


def ttf(m, dnx, src_iz, src_ix, sg=1, margin=27, frac=0.35, vmax=None, eager=False, level_margin=-1):
    """Replays seq-init + band march for one source; m is an oracle.ali_oracle.Model."""
    if vmax is None:
        vmax = model_vmax(m, dnx)
    nz = sg * (m.nz - 1) + 1 if sg > 1 else m.nz
    nx = sg * (m.nx - 1) + 1 if sg > 1 else m.nx
    T = np.zeros((nz, nx))
    cnt = np.zeros(12, dtype=np.int64)
    rc = lib().emu_ttf(*_margs(m, dnx), int(src_iz), int(src_ix), int(sg), int(margin), ctypes.c_double(frac),
                       ctypes.c_double(vmax), int(eager), int(level_margin), _p(T, _f64p),
                       cnt.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    names = ["seq_pops", "seq_evals", "seq_fallbacks", "rounds", "band_evals", "band_fallbacks", "max_list", "overflow",
             "level_rounds", "level_evals", "coop_steps", "coop_computed"]
    return T, dict(zip(names, cnt.tolist())), rc


def find_ray(m, dnx, source, receiver, rec_ttf, sg, nlanes=32):
    rec_ttf = np.ascontiguousarray(rec_ttf, dtype=np.float64)
    cap = 5 * (m.nz + m.nx)
    rx = np.zeros(cap)
    ry = np.zeros(cap)
    t = ctypes.c_double(0)
    flag = ctypes.c_int(0)
    n = lib().emu_find_ray(m.nz, m.nx, _p(m.veln, _f64p), _p(m.velpn, _i32p), _p(m.vel_map, _f64p),
                           _p(m.stif, _i64p), int(m.has_stif), _p(m.group, _f64p), m.ncol, ctypes.c_double(dnx),
                           int(sg), _p(rec_ttf, _f64p), rec_ttf.shape[0], rec_ttf.shape[1],
                           ctypes.c_double(source[0]), ctypes.c_double(source[1]), ctypes.c_double(receiver[0]),
                           ctypes.c_double(receiver[1]), _p(rx, _f64p), _p(ry, _f64p), cap, ctypes.byref(t),
                           ctypes.byref(flag), int(nlanes))
    return rx[:n].copy(), ry[:n].copy(), t.value, flag.value
