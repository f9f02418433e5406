"""Drop-in ``Anis_TTF_rays`` module: the reference's ``ALI_FMM`` class on B200 CUDA kernels.

Mirrors the public surface of the reference's Anis_TTF_rays.py ("ATR", class ALI_FMM at
ATR:3789-4705): same method names, argument meaning, return conventions and error
behaviour.  The two hot paths -- the anisotropic travel-time-field solve (travel /
travel_finer_grid, ATR:1463-2832) and the ray tracing through it (find_ray,
ATR:3104-3465) -- run in libalifmm.so (csrc/alifmm.cu) through the C ABI of
include/alifmm.h.  There is no CPU fallback: without the built library and a CUDA
device the compute methods raise.

The reference parallelises over sources with ``multiprocessing`` workers
(ATR:3560-3733); here sources are batched on a GPU and sharded over the visible GPUs
(``n_threads`` is validated as in the reference but is not a thread count any more).
"""
import math
import os
import threading

import numpy as np

from . import _capi
from .sharding import split_list as _split

try:  # progress bars are optional, as is the reference's module-level switch (ATR:22-24)
    from tqdm.auto import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None

tqdm_disable = False

__all__ = ["ALI_FMM", "tqdm_disable", "set_devices"]

_devices = None


def set_devices(devices):
    """Selects the CUDA devices the ``*_parallel`` methods shard sources over (default: env
    ALIFMM_DEVICES as a comma list, else every visible device).  The serial methods use
    the first one."""
    global _devices
    _devices = None if devices is None else [int(d) for d in devices]


def _device_list():
    if _devices is not None:
        return list(_devices)
    env = os.environ.get("ALIFMM_DEVICES")
    if env:
        return [int(t) for t in env.split(",") if t.strip() != ""]
    n = _capi.device_count()
    if n < 1:
        raise _capi.AlifmmError(-2, "no CUDA device available (this module has no CPU path)")
    return list(range(n))


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


class _Bar:
    """tqdm wrapper honouring the module's ``tqdm_disable`` (ATR:24)."""

    def __init__(self, total, desc):
        self.bar = None
        if _tqdm is not None and not tqdm_disable:
            self.bar = _tqdm(total=int(total), desc=desc, ncols=100, colour="green",
                             bar_format="{l_bar} {bar} | {n_fmt}/{total_fmt} [{elapsed}]")

    def update(self, n=1):
        if self.bar is not None and n:
            self.bar.update(int(n))

    def close(self):
        if self.bar is not None:
            self.bar.close()


def _zeros_sparse(shape):
    """np.zeros for a large array of which only a small part will be written (the dense ray-path
    arrays: 2 x 605 MB for 128 transducers, ~6 % used).  With transparent huge pages set to
    "always" the first touch of every 2 MB region zeroes all of it (0.1 s per call for the
    headline workload); an anonymous mapping that opts out of huge pages only pays for the 4 KB
    pages that are written."""
    import mmap
    n = int(np.prod(shape)) * 8
    if n < (64 << 20) or not hasattr(mmap, "MADV_NOHUGEPAGE"):
        return np.zeros(shape)
    mm = mmap.mmap(-1, n)
    try:
        mm.madvise(mmap.MADV_NOHUGEPAGE)
    except (OSError, ValueError):
        pass
    return np.frombuffer(mm, dtype=np.float64).reshape(shape)


def _axis_velocity(a, c_22, c_33, density):
    """Velocity along a symmetry axis (a = 0 or 90 degrees): sqrt(c / rho), exact special case of the
    reference's curves (ATR:4134-4139, 4183-4188)."""
    return math.sqrt((c_33 if a == 90 else c_22) / density)


def _group_sample(a, c_22, c_23, c_33, c_44, density):
    """Christoffel group velocity at a whole degree 0 <= a < 180 (closed form of ATR:4141-4152:
    phase angle from the quadratic in tan, then the group speed along the ray direction)."""
    if a % 90 == 0:
        return _axis_velocity(a, c_22, c_33, density)
    rad = math.radians(a)
    t = math.tan(rad)
    A = c_22 + c_33 - 2 * c_44
    B = (c_23 + c_44) * (t - 1 / t)
    C = c_22 - c_33
    root = math.sqrt(B ** 2 + A ** 2 - C ** 2)
    phi = math.atan((-B - root if a < 90 else -B + root) / (C - A)) % math.pi
    lam = 0.5 * (math.cos(2 * phi) * (c_22 - c_44) + math.sin(2 * phi) * (c_23 + c_44) * t + c_22 + c_44)
    return math.sqrt(lam / density) / math.cos(rad - phi)


def _phase_sample(a, c_22, c_23, c_33, c_44, density):
    """Christoffel phase velocity at a whole degree 0 <= a < 180 (ATR:4190-4197)."""
    if a % 90 == 0:
        return _axis_velocity(a, c_22, c_33, density)
    co = math.cos(math.radians(a))
    si = math.sin(math.radians(a))
    A = co ** 2 * c_22 + si ** 2 * c_44
    B = co * si * (c_23 + c_44)
    C = co ** 2 * c_44 + si ** 2 * c_33
    return math.sqrt((A + C + math.sqrt((A - C) ** 2 + 4 * B ** 2)) / (2 * density))


def _mirror_curve(half):
    """361 samples from the 180 of the upper half plane: the curves are pi-periodic (ATR:4154, 4200)."""
    curve = np.empty(361)
    curve[0:180] = half
    curve[180:360] = half
    curve[360] = half[0]
    return curve


def _polar_plot(curve, title):
    plt = _plt()
    if plt is not None:
        plt.polar(math.pi / 180 * np.arange(0, 361), curve)
        plt.title(title)
        plt.show()


class ALI_FMM:
    """Travel time fields and ray tracing in anisotropic media (reference: ATR:3789)."""

    # bytes of device memory one resident field node costs (T fp64 + status byte)
    _BYTES_PER_NODE = 17   # result field f64 + tiled march field f64 + alive byte
    # fraction of the free device memory a batch of fields may take
    _MEM_FRACTION = 0.8

    def __init__(self, veln, velpn, vel_map, scx, scz, group_vel=None, phase_vel=None, stif_den=None, dnx=1e-3):
        # ATR:3818-3867
        self.stif_den = stif_den
        if type(stif_den) != type(None):
            if type(stif_den[0, 0, 0]) != np.int64:
                raise TypeError("Stifness tensors and density array must have the type np.int64. 32bit integers will not work correctly.")
            elif stif_den[0, 0, 0] > 1e9:
                print("Warning: Stifness tensors must be in MPa, due to 64 bit integer limitations when solving the christoffel equation")
        if type(group_vel) == type(None):
            self.velocity_dat = 1 * np.ones((361, 2))
            self.velocity_dat[:, 0] = np.arange(0, 361)
            self.phase_vel = np.copy(self.velocity_dat)
        else:
            self.velocity_dat = group_vel
            self.phase_vel = phase_vel
        self.veln = veln
        self.velpn = velpn
        try:
            if np.issubdtype(velpn[0, 0], np.integer) == False:
                raise TypeError("velpn must be a numpy array of integers")
        except:
            raise TypeError("velpn must be a numpy array of integers")
        self.vel_map = vel_map
        self.dnx = dnx
        self.dnz = dnx
        self.nnx = veln.shape[1]
        self.nnz = veln.shape[0]
        self.ttn = np.zeros(veln.shape)
        self.scx = scx
        self.scz = scz
        self.gox = 0
        self.goz = 0
        self.isx = np.zeros(len(scx))
        self.isz = np.zeros(len(scx))
        for i in range(len(scx)):
            self.isx[i] = round((scx[i] - self.gox) / self.dnx)
            self.isz[i] = round((scz[i] - self.goz) / self.dnz)
        self.ntr = 0
        self.nsrc = len(scx)
        self.ray_paths_x = None
        self.ray_paths_y = None
        self.ray_len = None
        # B200 additions (not part of the reference surface)
        self.ray_flags = None        # per-ray status bits instead of the reference's print (ATR:3407)
        self.last_counters = None    # work counters / device timings of the last call, per device
        self.options = {}            # alifmm_set_option() overrides, e.g. {"delta_frac": 0.25}
        self.model_velocity_range = None   # (min, max) group velocity found by find_all_TTF_rays_parallel's model scan

    # ------------------------------------------------------------------ internals
    def _source_nodes(self, indices):
        """Coarse nodes of the given transducers: round(scx/dnx) as travel() does (ATR:1509-1510)."""
        iz = np.array([int(round((float(self.scz[i]) - self.goz) / self.dnz)) for i in indices], dtype=np.int32)
        ix = np.array([int(round((float(self.scx[i]) - self.gox) / self.dnx)) for i in indices], dtype=np.int32)
        return iz, ix

    def _context(self, veln, velpn, vel_map, stif_den, device):
        """Uploads the model; ``stif_den`` None means the zeros the reference substitutes (ATR:3890)."""
        ctx = _capi.Context(veln, velpn, vel_map, stif_den, True, self.velocity_dat, self.phase_vel, self.dnx,
                            device=device)
        for k, v in self.options.items():
            if k != "tables_on_device":   # (a host-side switch of add_materials, not a library option)
                ctx.set_option(k, v)
        return ctx

    def _fields_per_batch(self, ctx, subgrid):
        fz, fx = ctx.field_shape(subgrid)
        free, _ = ctx.mem_info()
        per = fz * fx * self._BYTES_PER_NODE + (64 << 20)
        return max(1, int(free * self._MEM_FRACTION // per))

    def _ttf_on_devices(self, veln, velpn, vel_map, stif_den, subgrid_size, indices, devices, dest, done=None):
        """Computes the fields of transducers ``indices`` sharded over ``devices``.  Field i is
        copied from the device straight into ``dest(i)`` (a C-contiguous float64 array of the
        field's shape; fields stay in HBM until then -- no second host copy), then ``done(i, array)``
        is called."""
        indices = list(indices)
        devices = devices[:max(1, min(len(devices), len(indices)))]
        shards = _split(indices, len(devices))
        errors = []
        counters = [None] * len(devices)
        lock = threading.Lock()

        def work(d, shard):
            try:
                ctx = self._context(veln, velpn, vel_map, stif_den, devices[d])
                try:
                    step = self._fields_per_batch(ctx, subgrid_size)
                    for pos in range(0, len(shard), step):
                        part = shard[pos:pos + step]
                        iz, ix = self._source_nodes(part)
                        ctx.ttf(iz, ix, subgrid_size, fetch=False)
                        counters[d] = _merge_counters(counters[d], ctx.counters(), None)
                        for slot, i in enumerate(part):
                            with lock:
                                out = dest(i)
                            ctx.ttf_fetch(slot, out=out)
                            if done is not None:
                                with lock:
                                    done(i, out)
                finally:
                    ctx.close()
            except BaseException as e:  # re-raised in the caller's thread
                errors.append(e)

        if len(devices) == 1:
            work(0, shards[0])
        else:
            threads = [threading.Thread(target=work, args=(d, shards[d])) for d in range(len(devices))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        if errors:
            raise errors[0]
        self.last_counters = counters

    # ------------------------------------------------------------------ travel time fields
    def update(self, veln, velpn, vel_map=None, stif_den=None, subgrid_size=1, sources=None):
        """Travel time fields of all (selected) sources: float64 [nsrc, nz', nx'] (ATR:3870)."""
        if type(stif_den) == type(None):
            self.stif_den = np.zeros((veln.shape[0], veln.shape[1], 5))
        else:
            self.stif_den = stif_den
        self.veln = veln
        self.velpn = velpn
        if type(vel_map) == type(None):
            self.vel_map = np.ones(veln.shape)
        else:
            self.vel_map = vel_map
        if type(sources) == type(None):
            sources = np.ones(len(self.scx))
        return self._update_impl(stif_den, subgrid_size, sources, _device_list()[:1], False)

    def update_parallel(self, veln, velpn, vel_map=None, stif_den=None, subgrid_size=1, sources=None, n_threads=2,
                        low_mem=False):
        """As ``update`` but sharded over the visible GPUs (reference: worker processes,
        ATR:3938).  ``low_mem`` spills each field to ``temp_TTF_<i>.npy`` and returns None
        (ATR:3612-3615, 3660-3670)."""
        if type(stif_den) == type(None):
            self.stif_den = np.zeros((veln.shape[0], veln.shape[1], 5))
        else:
            self.stif_den = stif_den
        self.veln = veln
        self.velpn = velpn
        if type(vel_map) == type(None):
            self.vel_map = np.ones(veln.shape)
        else:
            self.vel_map = vel_map
        if type(sources) == type(None):
            sources = np.ones(len(self.scx), dtype=int)
        return self._update_impl(stif_den, subgrid_size, sources, _device_list(), low_mem)

    def _update_impl(self, stif_den, subgrid_size, sources, devices, low_mem):
        selected = [i for i in range(self.nsrc) if sources[i] == 1]
        if subgrid_size == 1:
            shape = (self.veln.shape[0], self.veln.shape[1])
        else:
            shape = (subgrid_size * (self.veln.shape[0] - 1) + 1, subgrid_size * (self.veln.shape[1] - 1) + 1)
        if low_mem:
            travel_time_field = None
            spill = {}

            def dest(i):
                # one buffer per worker thread, reused for every field it spills (ATR:3612-3615, 3660-3670)
                return spill.setdefault(threading.get_ident(), np.empty(shape))

            def done(i, field):
                np.save("temp_TTF_" + str(i) + ".npy", field)
        else:
            if subgrid_size != 1 and not selected:
                # the reference sizes its result from the first computed field (ATR:3928-3936)
                raise UnboundLocalError("travel_time_field: no source selected")
            travel_time_field = np.zeros((self.nsrc, shape[0], shape[1]))

            def dest(i):
                return travel_time_field[i]

            def done(i, field):
                pass
        if selected:
            bar = _Bar(len(selected), "Finished TTF's")

            def counting_done(i, field):
                done(i, field)
                bar.update(1)

            try:
                self._ttf_on_devices(self.veln, self.velpn, self.vel_map, stif_den, subgrid_size, selected, devices,
                                     dest, counting_done)
            finally:
                bar.close()
        return travel_time_field

    def update_i(self, source_i, veln, velpn, vel_map, stif_den=None, subgrid_size=1):
        """Travel time field of one source (ATR:4053)."""
        if type(vel_map) == type(None):
            vel_map = np.ones(veln.shape)
        ctx = self._context(veln, velpn, vel_map, stif_den, _device_list()[0])
        try:
            fz, fx = ctx.field_shape(subgrid_size)
            out = np.empty((fz, fx))
            iz, ix = self._source_nodes([source_i])
            ctx.ttf(iz, ix, subgrid_size, fetch=False)
            ctx.ttf_fetch(0, out=out)
            self.last_counters = [ctx.counters()]
        finally:
            ctx.close()
        return out

    def update_i_split(self, source_i, veln, velpn, vel_map, stif_den=None, devices=None):
        """Travel time field of one source (``update_i`` at subgrid 1) on a grid too large for one GPU: the field
        is cut into row strips, one per device (2 ... 8; default: the visible devices), which exchange their halo
        and the per-round minimum over NVLink (``alifmm_ttf_split``; no counterpart in the reference, same bits as
        ``update_i``)."""
        if type(vel_map) == type(None):
            vel_map = np.ones(veln.shape)
        devices = list(_device_list() if devices is None else devices)[:8]
        iz, ix = self._source_nodes([source_i])
        out, counters = _capi.ttf_split(veln, velpn, vel_map, stif_den, True, self.velocity_dat, self.phase_vel, self.dnx,
                                        int(iz[0]), int(ix[0]), devices=devices)
        self.last_counters = [counters]
        return out

    # ------------------------------------------------------------------ material tables
    def plot_phase(self, material_index=1):
        """Polar plot of a tabulated phase velocity curve (ATR:4090)."""
        plt = _plt()
        if plt is None:
            raise ImportError("matplotlib is required for plot_phase")
        plt.polar(math.pi / 180 * self.velocity_dat[:, 0], self.phase_vel[:, material_index])
        plt.show()

    def plot_group(self, material_index=1):
        """Polar plot of a tabulated group velocity curve (ATR:4101)."""
        plt = _plt()
        if plt is None:
            raise ImportError("matplotlib is required for plot_group")
        plt.polar(math.pi / 180 * self.velocity_dat[:, 0], self.velocity_dat[:, material_index])
        plt.show()

    def generate_group_vel(self, c_22, c_23, c_33, c_44, density, plot=True):
        """Group velocity curve of a material: 361 samples at 1 degree, stiffness in Pa (ATR:4112).
        Callable with ``None`` as self, as the reference's documentation does."""
        curve = _mirror_curve([_group_sample(a, c_22, c_23, c_33, c_44, density) for a in range(180)])
        if plot == True:
            _polar_plot(curve, "Group Velocity")
        return curve

    def generate_phase_vel(self, c_22, c_23, c_33, c_44, density, plot=True):
        """Phase velocity curve of a material: 361 samples at 1 degree, stiffness in Pa (ATR:4162)."""
        curve = _mirror_curve([_phase_sample(a, c_22, c_23, c_33, c_44, density) for a in range(180)])
        if plot == True:
            _polar_plot(curve, "Phase Velocity")
        return curve

    def add_materials(self, materials, keep_materials=False):
        """Replaces (default) or extends (``keep_materials``) the velocity tables with the curves of
        ``materials`` -- one (c22, c23, c33, c44 [Pa], density) row, or a 2-D array of rows (ATR:4208).
        The table widths follow the reference, including its sizing of 2-D input by
        ``materials.shape[1]`` (ATR:4228-4231, 4243-4252).  With ``options["tables_on_device"]`` the
        curves of all rows come from one alifmm_velocity_curves_batch launch."""
        materials = np.asarray(materials)
        rows = materials[None, :] if materials.ndim == 1 else materials
        old_g, old_p = self.velocity_dat, self.phase_vel
        n_old = old_g.shape[1]
        if keep_materials == True:
            first = n_old
            n_fill = 1 if materials.ndim == 1 else materials.shape[0]
            width = n_old + (1 if materials.ndim == 1 else materials.shape[1])
        else:
            first = 1
            n_fill = 1 if materials.ndim == 1 else materials.shape[1]   # (the reference loops over shape[1] here)
            width = 1 + n_fill
        if rows.shape[0] < n_fill or first + n_fill > width:
            # the reference indexes past its arrays in these cases (ATR:4232-4235, 4249-4252)
            raise IndexError("add_materials: %d material rows do not fit the reference's table sizing" % rows.shape[0])
        curves_g, curves_p = self._material_curves(rows[:n_fill])
        tables = []
        for old, curves in ((old_g, curves_g), (old_p, curves_p)):
            tab = np.zeros((361, width))
            if keep_materials == True:
                tab[:, :n_old] = old
            else:
                tab[:, 0] = np.arange(0, 361)
            tab[:, first:first + n_fill] = curves.T
            tables.append(tab)
        if keep_materials == True:
            if materials.ndim == 1:
                print("material id of new material is " + str(n_old))
            else:
                print("material id's of new materials are " + str(n_old) + " - " + str(n_old + materials.shape[0] - 1))
        self.velocity_dat, self.phase_vel = tables

    def _material_curves(self, rows):
        """Group / phase curves of the material rows, float64 [n, 361] each."""
        if getattr(self, "options", None) and self.options.get("tables_on_device"):
            return _capi.velocity_curves_batch(np.asarray(rows, dtype=np.float64), device=_device_list()[0])
        g = np.array([ALI_FMM.generate_group_vel(self, r[0], r[1], r[2], r[3], r[4], False) for r in rows])
        p = np.array([ALI_FMM.generate_phase_vel(self, r[0], r[1], r[2], r[3], r[4], False) for r in rows])
        return g.reshape(len(rows), 361), p.reshape(len(rows), 361)

    # ------------------------------------------------------------------ rays
    def _default_pairs(self, n_trans):
        trans_pairs = np.zeros((n_trans, n_trans))
        for i in range(n_trans):
            for j in range(n_trans):
                if i < j:
                    trans_pairs[i, j] = 1
        return trans_pairs

    def _ttf_rays_impl(self, veln, velpn, vel_map, subgrid_size, trans_pairs, stif_den, save_rays, devices,
                       scan_model=False):
        n_trans = len(self.isx)
        cap = 5 * (veln.shape[0] + veln.shape[1])
        if save_rays:
            self.ray_paths_x = _zeros_sparse((n_trans, n_trans, cap))
            self.ray_paths_y = _zeros_sparse((n_trans, n_trans, cap))
            self.ray_len = np.zeros((n_trans, n_trans), dtype=int)
        self.ray_flags = np.zeros((n_trans, n_trans), dtype=int)
        if type(trans_pairs) == type(None):
            trans_pairs = self._default_pairs(n_trans)
        receivers = [j for j in range(n_trans) if np.sum(trans_pairs[:, j]) > 0]
        times = np.zeros((n_trans, n_trans))
        if not receivers:
            return times
        # rays of receiver j: every source i != j with trans_pairs[i, j] == 1 (ATR:4340-4342)
        pairs_of = {j: [i for i in range(n_trans) if i != j and trans_pairs[i, j] == 1] for j in receivers}
        n_rays = sum(len(v) for v in pairs_of.values())
        bar_ttf = _Bar(len(receivers), "Finished TTF's    ")
        bar_ray = _Bar(n_rays, "Finished ray paths")
        devices = devices[:max(1, min(len(devices), len(receivers)))]
        shards = _split(receivers, len(devices))
        errors = []
        counters = [None] * len(devices)
        lock = threading.Lock()

        def work(d, shard):
            try:
                ctx = self._context(veln, velpn, vel_map, stif_den, devices[d])
                try:
                    if scan_model and d == 0:
                        # the reference's model sanity scan (ATR:4583-4587; device reduction
                        # alifmm_min_max_vel).  Like the reference -- whose Warning objects are built
                        # but never raised -- nothing is printed; the range is kept for the caller.
                        self.model_velocity_range = ctx.min_max_vel()
                    step = self._fields_per_batch(ctx, subgrid_size)
                    agg = None
                    for pos in range(0, len(shard), step):
                        part = shard[pos:pos + step]
                        iz, ix = self._source_nodes(part)
                        ctx.ttf(iz, ix, subgrid_size, fetch=False)   # fields stay in HBM
                        c_ttf = ctx.counters()
                        with lock:
                            bar_ttf.update(len(part))
                        ray_i, ray_slot = [], []
                        for slot, j in enumerate(part):
                            for i in pairs_of[j]:
                                ray_i.append(i)
                                ray_slot.append(slot)
                        c_ray = None
                        if ray_i:
                            siz, six = self._source_nodes(ray_i)
                            ri = np.asarray(ray_i)
                            rj = np.asarray(part)[np.asarray(ray_slot)]
                            if save_rays:
                                # the library writes each path straight into ray_paths_x / ray_paths_y
                                # (coarse-cell units, ATR:4355-4356)
                                ln, tm, fl = ctx.rays_into(siz, six, ray_slot, cap, subgrid_size, ri * n_trans + rj,
                                                           self.ray_paths_x, self.ray_paths_y)
                            else:
                                _, _, ln, tm, fl = ctx.rays(siz, six, ray_slot, cap, want_paths=False)
                            c_ray = ctx.counters()
                            with lock:
                                times[ri, rj] = tm
                                self.ray_flags[ri, rj] = fl
                                if save_rays:
                                    self.ray_len[ri, rj] = ln
                                bar_ray.update(len(ray_i))
                        agg = _merge_counters(agg, c_ttf, c_ray)
                    counters[d] = agg
                finally:
                    ctx.close()
            except BaseException as e:
                errors.append(e)

        try:
            if len(devices) == 1:
                work(0, shards[0])
            else:
                threads = [threading.Thread(target=work, args=(d, shards[d])) for d in range(len(devices))]
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
        finally:
            bar_ttf.close()
            bar_ray.close()
        if errors:
            raise errors[0]
        self.last_counters = counters
        return times

    def find_all_TTF_rays(self, veln, velpn, vel_map=None, subgrid_size=9, trans_pairs=None, stif_den=None,
                          save_rays=True):
        """Receiver travel time fields + ray paths for all requested pairs; returns the travel
        times [n_trans, n_trans] (ATR:4258).  Paths are read with ``ray_path``."""
        if type(vel_map) == type(None):
            vel_map = np.ones(veln.shape)
        return self._ttf_rays_impl(veln, velpn, vel_map, subgrid_size, trans_pairs, stif_den, save_rays,
                                   _device_list()[:1])

    def find_all_TTF_rays_parallel(self, veln, velpn, vel_map=None, subgrid_size=9, trans_pairs=None, stif_den=None,
                                   n_threads=2, save_rays=True):
        """As ``find_all_TTF_rays`` with receivers sharded over the visible GPUs (ATR:4550)."""
        if n_threads == 1:
            raise ValueError("n_threads should not equal one. Use find_all_TTF_rays for single process.")
        if type(vel_map) == type(None):
            vel_map = np.ones(veln.shape)
        return self._ttf_rays_impl(veln, velpn, vel_map, subgrid_size, trans_pairs, stif_den, save_rays, _device_list(),
                                   scan_model=True)

    def ray_path(self, i, j):
        """Ray path from transducer i to j computed by find_all_TTF_rays* (ATR:4687)."""
        if self.ray_len[i, j] == 0:
            print("Ray path has not been calculated")
            return None, None
        else:
            ray_len = self.ray_len[i, j]
            return self.ray_paths_x[i, j, 0:ray_len], self.ray_paths_y[i, j, 0:ray_len]


def _merge_counters(agg, c_ttf, c_ray):
    """Sums the per-batch counters of one device."""
    out = dict(agg) if agg else {}
    for src, keys in ((c_ttf, ("node_solves", "seq_pops", "seq_evals", "band_rounds", "band_evals", "fallback_evals",
                               "ms_seq", "ms_march", "ms_finalize")),
                      (c_ray, ("rays", "ray_points", "ms_rays"))):
        if src is None:
            continue
        for k in keys:
            out[k] = out.get(k, 0) + src[k]
        for k in ("band_rounds_max", "max_band"):
            out[k] = max(out.get(k, 0), src[k])
        for k in ("vmax", "delta"):
            out[k] = src[k]
    return out
