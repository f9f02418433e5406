"""Test infrastructure: explains WHERE a computed travel-time field differs from the reference's.

The CUDA path runs the reference's operator on the reference's libm bits (csrc/ali_glibcmath.cuh), so
it can only differ from the reference where the band march and the reference's heap disagree about
what a node sees when it is computed.  ``classify_deviations`` checks that claim node by node:

* deviating nodes form *patches*: connected sets under "within two 12-neighbour windows of each other"
  (a node's value depends on the values -- final or still tentative -- of its window, and a tentative value
  on that node's window in turn);
* the *root* of a patch is its earliest node (smallest reference time): nothing deviating lies upstream
  of it.  A root is a *reference heap glitch* when the reference's own value there is NOT what its update
  operator gives from the reference's own earlier neighbours (the node was popped late by the mis-ordered
  heap -- parent index round(k/2), ATR:123 -- and re-evaluated from a non-causal state), while the computed
  field holds exactly that causal value.

Every deviating node belongs to a patch, and every patch's root must be a glitch: then 100 % of the
deviations are explained by a property of the reference.  Uses the oracle (tests only)."""
import numpy as np

W_OFF = [(-2, 0), (-1, -1), (-1, 0), (-1, 1), (0, -2), (0, -1), (0, 1), (0, 2), (1, -1), (1, 0), (1, 1), (2, 0)]


def _shift(a, dz, dx, fill):
    """b[z, x] = a[z + dz, x + dx] (``fill`` outside)."""
    out = np.full_like(a, fill)
    nz, nx = a.shape
    zs, ze = max(0, -dz), min(nz, nz - dz)
    xs, xe = max(0, -dx), min(nx, nx - dx)
    out[zs:ze, xs:xe] = a[zs + dz:ze + dz, xs + dx:xe + dx]
    return out


def causal_value(orc, m, field, z, x, sg=1):
    """The reference operator's value at (z, x) from the nodes of ``field`` with a smaller time
    (update(), ATR:904-1410; fouds18_A on -1.0, ATR:2069-2070), evaluated at the node's ABSOLUTE
    coordinates on a slab of the rows around it (oracle.update_node_slab: the operator interpolates in
    absolute coordinates, so a renumbered window could differ in the last ulp).  ``field`` is the
    reference's output: seconds, already divided by ``sg`` on the fine path (ATR:2832) -- it is
    multiplied back, which is exact only for sg == 1."""
    nz, nx = field.shape
    z0, z1 = max(0, z - 3), min(nz, z + 4)
    t = np.ascontiguousarray(field[z0:z1] * sg)
    nsts = np.where(t < t[z - z0, x], 0, -1).astype(np.int32)
    if sg > 1:   # material of a fine node: nearest coarse node, orientation truncated, vel_map in float32 (ATR:2156-2163)
        cz = (np.arange(z0, z1) + (sg - 1) // 2) // sg
        cx = (np.arange(nx) + (sg - 1) // 2) // sg
        veln = np.trunc(m["veln"][np.ix_(cz, cx)])
        vel_map = m["vel_map"][np.ix_(cz, cx)].astype(np.float32).astype(np.float64)
        velpn = m["velpn"][np.ix_(cz, cx)]
        stif = m["stif_den"][np.ix_(cz, cx)] if m["stif_den"] is not None else None
    else:
        veln, vel_map, velpn = m["veln"][z0:z1], m["vel_map"][z0:z1], m["velpn"][z0:z1]
        stif = m["stif_den"][z0:z1] if m["stif_den"] is not None else None
    if stif is None:
        stif = np.zeros(veln.shape + (5,), dtype=np.int64)
    om = orc.Model(np.ascontiguousarray(veln), np.ascontiguousarray(velpn), np.ascontiguousarray(vel_map),
                   np.ascontiguousarray(stif), m.get("group_vel"), m.get("phase_vel"))
    v, _ = orc.update_node_slab(om, nz, z0, t, nsts, z, x, m["dnx"])
    return v / sg


def classify_deviations(orc, m, ref, got, sg=1, source=None, box=0, max_roots=400):
    """See the module docstring.  ``source`` / ``box``: fine-grid node and half-width of the last refined
    source box; roots inside it would be deviations of the sequential replica (there must be none)."""
    ref = np.asarray(ref)
    got = np.asarray(got)
    dev = ref != got
    n_dev = int(dev.sum())
    out = {"nodes": int(ref.size), "deviating": n_dev, "frac_gt_1e-5": float((np.abs(got - ref) > 1e-5 * np.abs(ref)).mean())}
    if n_dev == 0:
        out.update(roots=0, roots_glitch=0, roots_in_source_box=0, unexplained=0, explained_frac=1.0)
        return out
    from scipy import ndimage
    labels, n_patch = ndimage.label(ndimage.binary_dilation(dev, structure=np.ones((5, 5), dtype=bool), iterations=2))
    lab = np.where(dev, labels, 0)
    # earliest deviating node of every patch
    flat = np.flatnonzero(dev)
    order = flat[np.argsort(ref.ravel()[flat], kind="stable")]
    first = {}
    for i in order:
        k = int(lab.ravel()[i])
        if k not in first:
            first[k] = i
            if len(first) == n_patch:
                break
    roots = np.array([divmod(int(i), ref.shape[1]) for i in sorted(first.values(), key=lambda j: ref.ravel()[j])])
    out["patches"] = int(n_patch)
    out["roots"] = int(len(roots))
    glitch = in_box = unexplained = 0
    tol = 0.0 if sg == 1 else 1e-14   # (fine path: the field was divided by sg, see causal_value)
    worst = []
    for z, x in roots[:max_roots]:
        z, x = int(z), int(x)
        if source is not None and max(abs(z - source[0]), abs(x - source[1])) <= box:
            in_box += 1
            continue
        c = causal_value(orc, m, ref, z, x, sg)
        ref_is_causal = abs(c - ref[z, x]) <= tol * abs(c)
        got_is_causal = abs(c - got[z, x]) <= tol * abs(c)
        if (not ref_is_causal) and got_is_causal:
            glitch += 1
        else:
            unexplained += 1
            worst.append((z, x, float(ref[z, x]), float(got[z, x]), float(c)))
    checked = min(len(roots), max_roots)
    out.update(roots_checked=checked, roots_glitch=glitch, roots_in_source_box=in_box, unexplained=unexplained,
               explained_frac=float(glitch / max(1, checked - in_box)) if checked > in_box else 1.0, unexplained_samples=worst[:5])
    return out
