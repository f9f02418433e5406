"""CPU: pins the oracle (oracle/ali_oracle.c) against fixtures produced by the REAL reference
(tests/golden/make_golden.py).  Tolerance 1e-12 relative: the oracle restates the reference's
arithmetic in the same order on the same libm (bit-identical in the build container)."""
import os

import numpy as np
import pytest

from oracle import ali_oracle as orc
from tests import models

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12


def close(a, b, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)))


@pytest.fixture(scope="module")
def ops():
    z = np.load(os.path.join(G, "golden_ops.npz"))
    return {k: z[k] for k in z.files}  # materialise once: NpzFile re-inflates on every access


@pytest.fixture(scope="module")
def fields():
    z = np.load(os.path.join(G, "golden_fields.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def rays():
    z = np.load(os.path.join(G, "golden_rays.npz"))
    return {k: z[k] for k in z.files}


def test_update_operator_matches_reference(ops):
    """update() (ATR:904) on 4000 random states, incl. edge nodes, ties and -1.0 results."""
    n = len(ops["out_update"])
    dnx = float(ops["dnx"])
    bad = 0
    for c in range(n):
        m = orc.Model(ops["veln"][c], ops["velpn"][c], ops["vel_map"][c], ops["stif"][c], ops["group_tab"], ops["phase_tab"])
        v, _ = orc.update_node(m, ops["ttn"][c], ops["nsts"][c], ops["pos"][c][0], ops["pos"][c][1], dnx)
        if not close(v, ops["out_update"][c]):
            bad += 1
    assert bad == 0
    assert (ops["out_update"] == -1.0).sum() > 100  # the no-stencil path is exercised


def test_fouds_operator_matches_reference(ops):
    """fouds18_A() (ATR:240) on the same states."""
    n = len(ops["out_fouds"])
    dnx = float(ops["dnx"])
    for c in range(n):
        m = orc.Model(ops["veln"][c], ops["velpn"][c], ops["vel_map"][c], ops["stif"][c], ops["group_tab"], ops["phase_tab"])
        v = orc.fouds_node(m, ops["ttn"][c], ops["nsts"][c], ops["pos"][c][0], ops["pos"][c][1], dnx)
        assert close(v, ops["out_fouds"][c]), c


def test_time_between_points_matches_reference(ops):
    m = orc.Model(ops["tbp_veln"], ops["tbp_velpn"], ops["tbp_vel_map"], ops["tbp_stif"], ops["group_tab"], ops["phase_tab"])
    sg = int(ops["tbp_sg"])
    dnx = float(ops["dnx"])
    for (x1, x2, y1, y2), t in zip(ops["tbp_seg"], ops["tbp_time"]):
        assert close(orc.time_between_points(m, x1, x2, y1, y2, dnx, sg), t)


def test_group_velocity_matches_reference(ops):
    for a, v in zip(ops["gv_angle"], ops["gv_value"]):
        assert close(orc.group_vel(a, 249000, 133000, 205000, 125000, 7850), v)


def test_notebook_gradient_fields(fields):
    """Notebook cell 12 (travel, ATR:1463); BASELINE.md golden scalars."""
    m = models.notebook_gradient()
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], np.zeros((201, 201, 5), dtype=np.int64))
    T0 = orc.travel(om, m["scx"][0], m["scz"][0], m["dnx"])
    T1 = orc.travel(om, m["scx"][1], m["scz"][1], m["dnx"])
    assert close(T0, fields["nb1_T0"])
    assert close(T1[::4, ::4], fields["nb1_T1_sub"])
    assert close([T0.sum(), T1.sum(), T0.max(), T1.max()], fields["nb1_sum"])
    assert close([T0.max(), T1.max(), T0.sum(), T1.sum()],
                 [5.387580497828128e-05, 5.4978153508260094e-05, 1.3402942867072842, 0.9845241576063767], 1e-9)


def test_notebook_christoffel_and_table_fields(fields):
    m = models.notebook_christoffel()
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], m["stif_den"])
    T = np.stack([orc.travel(om, m["scx"][k], m["scz"][k], m["dnx"]) for k in range(3)])
    assert close(T[2], fields["nb3_T2"])
    assert close(T[:, ::4, ::4], fields["nb3_sub"])
    om = orc.Model(np.zeros((201, 201)), np.ones((201, 201), dtype=int), np.ones((201, 201)),
                   np.zeros((201, 201, 5), dtype=np.int64), fields["nb2_group"], fields["nb2_phase"])
    T = np.stack([orc.travel(om, x, z, 1e-3) for x, z in ((1e-3, 100e-3), (199e-3, 140e-3))])
    assert close(T[:, ::4, ::4], fields["nb2_sub"])


def test_weld_coarse_fields(fields):
    """Weld model, travel(): sources on the top edge, interior and bottom edge."""
    w = models.weld()
    om = orc.Model(w["veln"], w["velpn"], w["vel_map"], w["stif_den"])
    for k, (sx, sz) in enumerate(fields["weld1_src"]):
        T = orc.travel(om, w["dnx"] * sx, w["dnx"] * sz, w["dnx"])
        assert close(T[::4, ::4], fields["weld1_sub"][k])
        assert close(T.sum(), fields["weld1_sum"][k])
    # weld sg=1, source (x=25, z=0)
    assert close(fields["weld1_sum"][0], 2.4407884315144712, 1e-9)


@pytest.mark.parametrize("sg", [3, 5])
def test_weld_crop_fine_fields(fields, sg):
    """travel_finer_grid() (ATR:2120) on a 60 x 80 crop of the weld, four source positions."""
    c = models.weld_crop(60, 80)
    om = orc.Model(c["veln"], c["velpn"], c["vel_map"], c["stif_den"])
    for k, (sx, sz) in enumerate(fields["crop_src"]):
        T = orc.travel_finer_grid(om, c["dnx"] * sx, c["dnx"] * sz, c["dnx"], sg)
        assert close(T[::3, ::3], fields["crop_sg%d_sub" % sg][k])
        assert close(T.sum(), fields["crop_sg%d_sum" % sg][k])


def test_weld_crop_sg9_field(fields):
    c = models.weld_crop(30, 40)
    om = orc.Model(c["veln"], c["velpn"], c["vel_map"], c["stif_den"])
    T = orc.travel_finer_grid(om, c["dnx"] * 5.0, 0.0, c["dnx"], 9)
    assert close(T, fields["crop9_T"])


def test_notebook_rays(rays, fields):
    """find_ray() + find_all_TTF_rays times: notebook cells 16, 30, 40."""
    m = models.notebook_gradient()
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], np.zeros((201, 201, 5), dtype=np.int64))
    T = orc.travel_finer_grid(om, m["scx"][1], m["scz"][1], m["dnx"], 9)
    rx, ry, t, flag = orc.find_ray(om, m["dnx"], (9 * 1, 9 * 30), (9 * 199, 9 * 180), T, 9)
    assert len(rx) == len(rays["nb1_ray_x"]) == 341
    assert close(rx / 9, rays["nb1_ray_x"]) and close(ry / 9, rays["nb1_ray_y"])
    assert close(t, rays["nb1_times"][0, 1])
    assert abs(t - 5.08845096e-05) < 1e-13  # value stored in the notebook (cell 16)
    m = models.notebook_christoffel()
    om = orc.Model(m["veln"], m["velpn"], m["vel_map"], m["stif_den"])
    T2 = orc.travel_finer_grid(om, m["scx"][2], m["scz"][2], m["dnx"], 9)
    rx, ry, t, flag = orc.find_ray(om, m["dnx"], (9 * 199, 9 * 140), (9 * 100, 9 * 1), T2, 9)
    assert close(rx / 9, rays["nb3_ray_x_12"]) and close(ry / 9, rays["nb3_ray_y_12"])
    assert close(t, rays["nb3_times"][1, 2])
    assert abs(t - 2.76255662e-05) < 3e-12  # notebook cell 40 print-out (2e-8 .. 3e-7 version drift)


def test_weld_sg9_rays_through_reference_field(rays):
    """find_ray() through a 3808 x 4492 field: recompute the field with the oracle (16 s) and
    retrace the four golden rays of receiver 40."""
    w = models.weld()
    om = orc.Model(w["veln"], w["velpn"], w["vel_map"], w["stif_den"])
    T = orc.travel_finer_grid(om, w["dnx"] * 160, w["dnx"] * 423, w["dnx"], 9)
    for k, sx in enumerate(rays["weld9_ray_srcx"]):
        rx, ry, t, flag = orc.find_ray(om, w["dnx"], (9 * int(sx), 0), (9 * 160, 9 * 423), T, 9)
        assert close(rx / 9, rays["weld9_ray_x_%d" % k]) and close(ry / 9, rays["weld9_ray_y_%d" % k])
        assert close(t, rays["weld9_ray_times"][k])
