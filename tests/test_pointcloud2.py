"""sensor_msgs/PointCloud2 ingest ("next" row of SURVEY.md §8f: the caller-side data format).
CPU part: the oracle's restatement of nanopcl::from(msg) (bridge/ros/impl.hpp:180-270 — the
reference has no test for it, so these cases pin the restatement against the source's stated
behaviour) and the host-side field parsing.  GPU part: integrate(msg) with the unpacking done
in the first kernel == integrate(from(msg)) on the oracle."""
import numpy as np
import pytest

import oracle_binding as ob
from fastdem_b200.api import (PF_FLOAT32, PF_FLOAT64, PF_UINT8, PF_UINT16, PointCloud2)


def _msg(n=50, intensity_type=PF_FLOAT32, with_rgb=True, pad=0, seed=0):
    rng = np.random.RandomState(seed)
    xyz = rng.uniform(-4, 4, size=(n, 3)).astype(np.float32)
    xyz[3, 0] = np.nan
    xyz[7, 1] = np.inf
    xyz[11, 2] = -np.inf
    if intensity_type in (PF_UINT8, PF_UINT16):
        inten = rng.randint(0, 200, size=n)
    else:
        inten = rng.uniform(0, 1, size=n)
    rgb = rng.randint(0, 256, size=(n, 3)) if with_rgb else None
    return PointCloud2.from_arrays(xyz, inten, rgb, intensity_type=intensity_type, pad=pad), xyz, inten, rgb


@pytest.mark.parametrize("itype", [PF_UINT8, PF_UINT16, PF_FLOAT32, PF_FLOAT64])
def test_from_pointcloud2_drops_nonfinite_and_converts(itype):
    msg, xyz, inten, rgb = _msg(intensity_type=itype, pad=5)
    lo = msg.layout()
    assert (lo.off_x, lo.off_y, lo.off_z) == (0, 4, 8) and lo.off_intensity >= 12 and lo.off_rgb > lo.off_intensity
    assert msg.point_step % 4 == 0
    p, i, c = ob.from_pointcloud2(msg.data, msg.size(), lo)
    keep = np.isfinite(xyz).all(axis=1)
    assert keep.sum() == len(xyz) - 3 and p.shape[0] == keep.sum()
    assert np.array_equal(p[:, :3], xyz[keep]) and (p[:, 3] == 1.0).all()
    want_i = np.asarray(inten)[keep]
    if itype == PF_FLOAT32:
        want_i = want_i.astype(np.float32)
    elif itype == PF_FLOAT64:
        want_i = want_i.astype(np.float64).astype(np.float32)
    assert np.array_equal(i, want_i.astype(np.float32))
    assert np.array_equal(c, np.asarray(rgb)[keep].astype(np.uint8))


def test_from_pointcloud2_without_xyz_or_points_is_empty():
    msg, *_ = _msg()
    lo = msg.layout()
    lo.off_y = -1
    p, i, c = ob.from_pointcloud2(msg.data, msg.size(), lo)
    assert p.shape[0] == 0
    lo = msg.layout()
    p, i, c = ob.from_pointcloud2(msg.data, 0, lo)
    assert p.shape[0] == 0


def test_field_parsing_rgba_and_unknown_fields():
    m = PointCloud2(b"\x00" * 32, 1, 1, 32, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("ring", 12, 4),
                                             ("rgba", 16, 6), ("t", 20, 6), ("intensity", 24, 2)])
    lo = m.layout()
    assert (lo.off_rgb, lo.off_intensity, lo.intensity_type) == (16, 24, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("itype,with_rgb", [(PF_FLOAT32, False), (PF_UINT16, True), (PF_FLOAT64, True), (PF_UINT8, False),
                                            (None, True), (None, False)])   # 16-byte points: the 128-bit load path
def test_integrate_pointcloud2_matches_oracle(fdem, itype, with_rgb):
    from fastdem_b200 import synthetic as syn
    from parity_utils import compare_maps
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    gdem = fdem.FastDEM(gmap, cfg)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    rng = np.random.RandomState(1)
    for k in range(5):
        s = syn.make_scan(wl, k)
        xyz = s["xyzw"][:, :3].copy()
        bad = rng.choice(len(xyz), 40, replace=False)
        xyz[bad[:20], 0] = np.nan
        xyz[bad[20:], 2] = np.inf          # +inf survives the default range crop: must go in ingest
        inten = None if itype is None else s["intensity"] * (200 if itype in (PF_UINT8, PF_UINT16) else 1)
        rgb = rng.randint(0, 256, size=(len(xyz), 3)) if with_rgb else None
        pad = 0 if itype is None else k % 3 * 4
        if itype is None and not with_rgb:
            pad = 4   # x, y, z + 4 bytes of padding: still a 16-byte point, no channel
        msg = PointCloud2.from_arrays(xyz, inten, rgb, intensity_type=itype or PF_FLOAT32, pad=pad)
        st = gdem.integrate_pointcloud2(msg, s["T_base_sensor"], s["T_world_base"])
        p, i, c = ob.from_pointcloud2(msg.data, msg.size(), msg.layout())
        ok, ost, _ = odem.integrate(p, s["T_base_sensor"], s["T_world_base"], i, c)
        assert st.n_input == p.shape[0] == len(xyz) - 40
        assert (st.integrated, st.n_kept, st.n_cells) == (int(ok), ost.n_kept, ost.n_cells)
    compare_maps(gmap, omap)
    # the same through FastDEM.integrate(msg, ...) and with the body already on the device
    import torch
    dmsg = PointCloud2(torch.from_numpy(np.ascontiguousarray(msg.data)).cuda(), msg.width, msg.height,
                       msg.point_step, msg.fields)
    assert gdem.integrate(dmsg, s["T_base_sensor"], s["T_world_base"])
    odem.integrate(p, s["T_base_sensor"], s["T_world_base"], i, c)
    compare_maps(gmap, omap)


@pytest.mark.gpu
def test_integrate_pointcloud2_degenerate_messages(fdem):
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5)
    dem = fdem.FastDEM(gmap)
    allnan = PointCloud2.from_arrays(np.full((8, 3), np.nan, np.float32))
    assert not dem.integrate(allnan, np.eye(4), np.eye(4)) and gmap.isEmpty()
    empty = PointCloud2(np.zeros(0, np.uint8), 0, 1, 16, [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)])
    assert not dem.integrate(empty, np.eye(4), np.eye(4))
    noxyz = PointCloud2(np.zeros(64, np.uint8), 4, 1, 16, [("x", 0, 7), ("intensity", 12, 7)])
    assert not dem.integrate(noxyz, np.eye(4), np.eye(4)) and gmap.isEmpty()
    with pytest.raises(fdem.FdemError):  # a float field that is not 4-byte aligned inside the point
        bad = PointCloud2(np.zeros(64, np.uint8), 4, 1, 16, [("x", 1, 7), ("y", 5, 7), ("z", 9, 7)])
        dem.integrate(bad, np.eye(4), np.eye(4))


@pytest.mark.gpu
def test_submit_pointcloud2_stream_equals_sync(fdem):
    """submit_pointcloud2(k+1); collect(k) — the streaming form, two messages in flight — must
    leave the same map as integrate_pointcloud2 one message at a time."""
    from fastdem_b200 import synthetic as syn
    from parity_utils import compare_layer
    wl = syn.WORKLOADS["tiny"]
    maps = []
    for streaming in (False, True):
        gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
        gdem = fdem.FastDEM(gmap, wl.config())
        msgs = []
        for k in range(7):
            s = syn.make_scan(wl, k)
            msgs.append((PointCloud2.from_arrays(s["xyzw"][:, :3], s["intensity"]), s["T_base_sensor"], s["T_world_base"]))
        cells = []
        if streaming:
            prev = None
            for m, a, b in msgs:
                t = gdem.submit_pointcloud2(m, a, b)
                if prev is not None:
                    cells.append(gdem.collect(prev).n_cells)
                prev = t
            cells.append(gdem.collect(prev).n_cells)
        else:
            for m, a, b in msgs:
                cells.append(gdem.integrate_pointcloud2(m, a, b).n_cells)
        maps.append((gmap, cells))
    assert maps[0][1] == maps[1][1] and min(maps[0][1]) > 0
    for name in maps[0][0].getLayers():
        compare_layer(name, maps[1][0].get(name), maps[0][0].get(name), rtol=0, atol=0)
    assert msgs[0][0].point_step == 16   # x, y, z, intensity: 16 bytes per point on the wire
